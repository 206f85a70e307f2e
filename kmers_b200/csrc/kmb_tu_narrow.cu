// kmb_tu_narrow.cu -- instantiates the K <= 32 materialising engines (NarrowEng, MODE 0) on both geometries.
#include "kmb_launch.h"

namespace kmb {
namespace {
// Template dispatch: VALIDATE x DIGEST x FWRC x KHI for one MODE.
template <int MODE>
static cudaError_t launch_narrow(bool validate, bool digest, bool fwrc, bool khi, bool hash, const FixedGeom* fg, const CsrGeom* cg,
                                 const Launch& l, cudaStream_t st, const EncDesc& enc, const NarrowParams& ep) {
#define KMB_CASE(V, D, F, H) \
    if (validate == V && digest == D && fwrc == F && khi == H) return launch_eng<NarrowEng<V, D, F, MODE, H>>(fg, cg, l, st, enc, ep);
    if (MODE == 0 && !hash && !digest && !fwrc) {  // canonical words only: the hash arithmetic is compiled out
#define KMB_NOHASH(V, H) \
        if (validate == V && khi == H) return launch_eng<NarrowEng<V, false, false, 0, H, false>>(fg, cg, l, st, enc, ep);
        KMB_NOHASH(true, true) KMB_NOHASH(true, false) KMB_NOHASH(false, true) KMB_NOHASH(false, false)
#undef KMB_NOHASH
    }
    KMB_CASE(true, false, false, true) KMB_CASE(true, false, false, false)
    KMB_CASE(true, true, false, true) KMB_CASE(true, true, false, false)
    KMB_CASE(false, false, false, true) KMB_CASE(false, false, false, false)
    KMB_CASE(false, true, false, true) KMB_CASE(false, true, false, false)
    if (MODE == 0) {
        KMB_CASE(true, false, (MODE == 0), true) KMB_CASE(true, false, (MODE == 0), false)
        KMB_CASE(true, true, (MODE == 0), true) KMB_CASE(true, true, (MODE == 0), false)
        KMB_CASE(false, false, (MODE == 0), true) KMB_CASE(false, false, (MODE == 0), false)
        KMB_CASE(false, true, (MODE == 0), true) KMB_CASE(false, true, (MODE == 0), false)
    }
#undef KMB_CASE
    return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t launch_narrow_materialise(bool validate, bool digest, bool fwrc, bool khi, bool hash, const FixedGeom* fg, const CsrGeom* cg,
                                      const Launch& l, cudaStream_t st, const EncDesc& enc, const NarrowParams& ep) {
    return launch_narrow<0>(validate, digest, fwrc, khi, hash, fg, cg, l, st, enc, ep);
}

}  // namespace kmb
