// kmb_tu_wide.cu -- instantiates the two-word (K <= 64) engines on both geometries.
#include "kmb_launch.h"

namespace kmb {
namespace {
template <int NW32>
static cudaError_t launch_wide(bool validate, bool digest, bool hash, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                               cudaStream_t st, const EncDesc& enc, const WideParams& ep) {
    if (!hash && !digest)  // canonical words only (BASELINE config 3): the hash arithmetic is compiled out
        return validate ? launch_eng<WideEng<NW32, true, false, false>>(fg, cg, l, st, enc, ep)
                        : launch_eng<WideEng<NW32, false, false, false>>(fg, cg, l, st, enc, ep);
    if (validate) return digest ? launch_eng<WideEng<NW32, true, true>>(fg, cg, l, st, enc, ep)
                                : launch_eng<WideEng<NW32, true, false>>(fg, cg, l, st, enc, ep);
    return digest ? launch_eng<WideEng<NW32, false, true>>(fg, cg, l, st, enc, ep)
                  : launch_eng<WideEng<NW32, false, false>>(fg, cg, l, st, enc, ep);
}
}  // namespace

cudaError_t launch_wide(int nw32, bool validate, bool digest, bool hash, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                        cudaStream_t st, const EncDesc& enc, const WideParams& ep) {
    return nw32 == 2 ? launch_wide<2>(validate, digest, hash, fg, cg, l, st, enc, ep)
         : nw32 == 3 ? launch_wide<3>(validate, digest, hash, fg, cg, l, st, enc, ep)
                     : launch_wide<4>(validate, digest, hash, fg, cg, l, st, enc, ep);
}

}  // namespace kmb
