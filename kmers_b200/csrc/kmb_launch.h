// kmb_launch.h -- host-side launch interface between the C-ABI translation unit (kmers_b200.cu) and the
// translation units that instantiate the kernels (kmb_tu_*.cu).  Internal: nothing here is exported.
#pragma once
#include <cuda_runtime.h>

#include "kmb_extract.cuh"
#include "kmb_extract_wide.cuh"
#include "kmb_compact.cuh"
#include "kmb_minimizer.cuh"

namespace kmb {

struct Launch {
    unsigned grid = 0;
    size_t smem = 0;
};

// one engine on the fixed-length or the ragged (CSR) geometry
template <class Eng>
inline cudaError_t launch_eng(const FixedGeom* fg, const CsrGeom* cg, const Launch& l, cudaStream_t st, const EncDesc& enc,
                              const typename Eng::Params& ep) {
    if (l.smem > 36 * 1024) {  // with the kernels' static shared memory that may exceed the 48 KiB a launch gets without opting in
        const cudaError_t e = fg ? cudaFuncSetAttribute(fixed_kernel<Eng>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem)
                                 : cudaFuncSetAttribute(csr_kernel<Eng>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)l.smem);
        if (e != cudaSuccess) return e;
    }
    if (fg) fixed_kernel<Eng><<<l.grid, kExtractThreads, l.smem, st>>>(*fg, enc, ep);
    else csr_kernel<Eng><<<l.grid, kExtractThreads, l.smem, st>>>(*cg, enc, ep);
    return cudaGetLastError();
}

// kmb_tu_narrow.cu: K <= 32, canonical / hash / fw / rc arrays (+ digest)
cudaError_t launch_narrow_materialise(bool validate, bool digest, bool fwrc, bool khi, bool hash, const FixedGeom* fg, const CsrGeom* cg,
                                      const Launch& l, cudaStream_t st, const EncDesc& enc, const NarrowParams& ep);
// kmb_tu_hist.cu: fused histogram, global-atomic bins / shared-memory bins (persistent grid)
cudaError_t launch_narrow_hist_global(bool validate, bool digest, bool khi, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                                      cudaStream_t st, const EncDesc& enc, const NarrowParams& ep);
cudaError_t launch_hist_smem(bool validate, bool digest, bool khi, const FixedGeom* fg, const CsrGeom* cg, unsigned grid, size_t smem,
                             uint32_t n_tiles, uint32_t tile_words, uint32_t n_bins, cudaStream_t st, const EncDesc& enc,
                             const NarrowParams& ep);
// kmb_tu_compact.cu: iterator-identical compacted stream
cudaError_t launch_compact(bool count_only, bool validate, bool khi, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                           cudaStream_t st, const EncDesc& enc, const CompactParams& ep);
// fixed-length reads, emit launch: persistent software-pipelined kernel (l.smem = ONE tile buffer, l.grid = tiles)
size_t compact_pipe_tile_budget();  // bytes one tile buffer may take so that KMB_COMPACT_MINCTAS CTAs fit an SM
cudaError_t launch_compact_pipe(bool validate, bool khi, const FixedGeom& fg, const Launch& l, int device, cudaStream_t st, const EncDesc& enc,
                                const CompactParams& ep);
cudaError_t launch_compact_fixup(const uint64_t* win_offsets, uint64_t W, uint64_t n_reads, uint64_t slots_per_cta, const unsigned long long* desc,
                                 uint64_t* emit_offsets, cudaStream_t st);
cudaError_t launch_compact_backfill(const uint64_t* win_offsets, uint64_t n_reads, const unsigned long long* total_emitted,
                                    uint64_t* emit_offsets, cudaStream_t st);
// kmb_tu_minimizer.cu
cudaError_t launch_minimizers(bool validate, const FixedGeom* fg, const CsrGeom* cg, const Launch& l, cudaStream_t st,
                              const EncDesc& enc, const MinParams& ep);
cudaError_t launch_minimizer_words(const uint64_t* in, uint64_t n, uint32_t k, uint32_t w, uint32_t hash_k, uint64_t* mmer_out,
                                   uint32_t* offset_out, cudaStream_t st);
// kmb_tu_wide.cu: two-word k-mers (K <= 64), nw32 = live 32-bit words of a k-mer (2..4)
cudaError_t launch_wide(int nw32, bool validate, bool digest, bool hash, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                        cudaStream_t st, const EncDesc& enc, const WideParams& ep);

}  // namespace kmb
