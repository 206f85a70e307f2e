// kmb_tu_hist.cu -- instantiates the fused-histogram engines: NarrowEng MODE 1 (global-atomic bins) and
// MODE 2 (16-bit shared-memory bins, persistent grid).
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

// MODE 2 (shared-memory bins): persistent grid, opt-in dynamic shared memory above 48 KiB
template <class Eng>
static cudaError_t launch_hist_eng(const FixedGeom* fg, const CsrGeom* cg, unsigned grid, size_t smem, uint32_t n_tiles,
                                   uint32_t tile_words, uint32_t n_bins, cudaStream_t st, const EncDesc& enc, const NarrowParams& ep) {
    cudaError_t e;
    if (fg) {
        e = cudaFuncSetAttribute(hist_fixed_kernel<Eng>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hist_fixed_kernel<Eng><<<grid, kHistThreads, smem, st>>>(*fg, enc, ep, n_tiles, tile_words, n_bins);
    } else {
        e = cudaFuncSetAttribute(hist_csr_kernel<Eng>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        hist_csr_kernel<Eng><<<grid, kHistThreads, smem, st>>>(*cg, enc, ep, n_tiles, n_bins);
    }
    return cudaGetLastError();
}
}  // namespace

cudaError_t launch_narrow_hist_global(bool validate, bool digest, bool khi, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                                      cudaStream_t st, const EncDesc& enc, const NarrowParams& ep) {
    return launch_narrow<1>(validate, digest, false, khi, true, fg, cg, l, st, enc, ep);
}

cudaError_t launch_hist_smem(bool validate, bool digest, bool khi, const FixedGeom* fg, const CsrGeom* cg, unsigned grid,
                                    size_t smem, uint32_t n_tiles, uint32_t tile_words, uint32_t n_bins, cudaStream_t st,
                                    const EncDesc& enc, const NarrowParams& ep) {
    const bool hibin = khi && ep.out.hist_hi_shift != 0xFFFFFFFFu;  // the bin is a field of the high word (NarrowEng::kHiBin)
#define KMB_CASE(V, D, H) \
    if (validate == V && digest == D && khi == H && !hibin) return launch_hist_eng<NarrowEng<V, D, false, 2, H, true>>(fg, cg, grid, smem, n_tiles, tile_words, n_bins, st, enc, ep);
    KMB_CASE(true, false, true) KMB_CASE(true, false, false) KMB_CASE(true, true, true) KMB_CASE(true, true, false)
    KMB_CASE(false, false, true) KMB_CASE(false, false, false) KMB_CASE(false, true, true) KMB_CASE(false, true, false)
#undef KMB_CASE
#define KMB_CASE(V, D) \
    if (validate == V && digest == D && hibin) return launch_hist_eng<NarrowEng<V, D, false, 2, true, false>>(fg, cg, grid, smem, n_tiles, tile_words, n_bins, st, enc, ep);
    KMB_CASE(true, false) KMB_CASE(true, true) KMB_CASE(false, false) KMB_CASE(false, true)
#undef KMB_CASE
    return cudaErrorInvalidValue;
}

}  // namespace kmb
