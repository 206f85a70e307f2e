// kmb_tu_compact.cu -- instantiates the compaction engines (sizing count; single-pass emit) on both geometries.
#include <algorithm>

#include "kmb_launch.h"

namespace kmb {
namespace {
template <bool COUNT_ONLY>
static cudaError_t launch_compact(bool validate, bool khi, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                                  cudaStream_t st, const EncDesc& enc, const CompactParams& ep) {
    // dynamic shared memory = the geometry's tile (+ tables), 16-byte aligned, then CompactShared: above 48 KiB -> opt in
    const uint32_t tile_bytes = (uint32_t)((l.smem + 15) & ~(size_t)15);
    const size_t smem = tile_bytes + (fg ? (COUNT_ONLY ? kCompactCountBytes : sizeof(CompactShared))
                                         : (COUNT_ONLY ? offsetof(CompactSharedT<kCompactCsrItems>, canon) : sizeof(CompactSharedT<kCompactCsrItems>)));
#define KMB_CASE(V, H)                                                                                                  \
    if (validate == V && khi == H) {                                                                                    \
        cudaError_t e;                                                                                                  \
        if (fg) {                                                                                                       \
            e = cudaFuncSetAttribute(compact_fixed_kernel<CompactEng<V, H, COUNT_ONLY>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                             \
            compact_fixed_kernel<CompactEng<V, H, COUNT_ONLY>><<<l.grid, kExtractThreads, smem, st>>>(*fg, enc, ep, tile_bytes); \
        } else {                                                                                                        \
            e = cudaFuncSetAttribute(compact_csr_kernel<CompactEng<V, H, COUNT_ONLY, false, kCompactCsrItems>>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                                             \
            compact_csr_kernel<CompactEng<V, H, COUNT_ONLY, false, kCompactCsrItems>><<<l.grid, kExtractThreads, smem, st>>>(*cg, enc, ep, tile_bytes); \
        }                                                                                                               \
        return cudaGetLastError();                                                                                      \
    }
    KMB_CASE(true, true) KMB_CASE(true, false) KMB_CASE(false, true) KMB_CASE(false, false)
#undef KMB_CASE
    return cudaErrorInvalidValue;
}
}  // namespace

cudaError_t launch_compact(bool count_only, bool validate, bool khi, const FixedGeom* fg, const CsrGeom* cg, const Launch& l,
                           cudaStream_t st, const EncDesc& enc, const CompactParams& ep) {
    return count_only ? launch_compact<true>(validate, khi, fg, cg, l, st, enc, ep)
                      : launch_compact<false>(validate, khi, fg, cg, l, st, enc, ep);
}

// 228 KiB of shared memory per SM, 1 KiB of it reserved per resident CTA: what one of KMB_COMPACT_MINCTAS CTAs can have
static constexpr size_t kPipeCtaSmem = (228 * 1024) / KMB_COMPACT_MINCTAS - 1024;

size_t compact_pipe_tile_budget() { return ((kPipeCtaSmem - sizeof(CompactPipeShared)) / 2) & ~(size_t)15; }

cudaError_t launch_compact_pipe(bool validate, bool khi, const FixedGeom& fg, const Launch& l, int device, cudaStream_t st, const EncDesc& enc,
                                const CompactParams& ep) {
    const uint32_t tile_bytes = (uint32_t)((l.smem + 15) & ~(size_t)15);
    const size_t smem = 2 * (size_t)tile_bytes + sizeof(CompactPipeShared);
    int sms = 0;
    cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
#define KMB_CASE(V, H)                                                                                                  \
    if (validate == V && khi == H) {                                                                                    \
        auto kern = compact_fixed_pipe_kernel<CompactEng<V, H, false, true, kCompactPipeItems>>;                                           \
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                         \
        if (e != cudaSuccess) return e;                                                                                 \
        int per_sm = 0;                                                                                                 \
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kExtractThreads, smem);                        \
        if (e != cudaSuccess) return e;                                                                                 \
        /* resident CTAs only: each walks the tiles by ticket (correct for any grid; more CTAs than fit would just queue) */ \
        const unsigned grid = (unsigned)std::min<uint64_t>(l.grid, (uint64_t)sms * (uint64_t)std::max(per_sm, 1));     \
        kern<<<grid, kExtractThreads, smem, st>>>(fg, enc, ep, tile_bytes, l.grid);                                     \
        return cudaGetLastError();                                                                                      \
    }
    KMB_CASE(true, true) KMB_CASE(true, false) KMB_CASE(false, true) KMB_CASE(false, false)
#undef KMB_CASE
    return cudaErrorInvalidValue;
}

// The emit kernel wrote every read's first-entry index counted from its TILE's first entry (the tile's start is the last thing
// a CTA learns); after the launch every descriptor holds the tile's inclusive prefix, so the start of tile t is desc[t - 1].
// win_offsets == NULL: fixed-length reads, read r's first slot is r * W.
__global__ void __launch_bounds__(256) compact_fixup_kernel(const uint64_t* win_offsets, uint64_t W, uint64_t n_reads, uint64_t slots_per_cta,
                                                            const unsigned long long* desc, uint64_t* emit_offsets) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    uint64_t slot;
    if (win_offsets) {
        slot = win_offsets[r];
        if (win_offsets[r + 1] == slot) return;  // no window, no entry: compact_backfill_kernel fills these in
    } else {
        slot = r * W;
    }
    const uint64_t tile = slot / slots_per_cta;
    if (tile) emit_offsets[r] += desc[tile - 1] & kDescValue;
}

cudaError_t launch_compact_fixup(const uint64_t* win_offsets, uint64_t W, uint64_t n_reads, uint64_t slots_per_cta, const unsigned long long* desc,
                                 uint64_t* emit_offsets, cudaStream_t st) {
    compact_fixup_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(win_offsets, W, n_reads, slots_per_cta, desc, emit_offsets);
    return cudaGetLastError();
}

// Reads without a window (shorter than k) open no entry: their emit offset is that of the next read that has
// windows (or the total).  win_offsets[r + 1] == win_offsets[r] identifies them.
__global__ void __launch_bounds__(256) compact_backfill_kernel(const uint64_t* win_offsets, uint64_t n_reads,
                                                               const unsigned long long* total_emitted_ptr, uint64_t* emit_offsets) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n_reads) return;
    const uint64_t total_emitted = *total_emitted_ptr;
    if (win_offsets[r + 1] != win_offsets[r]) return;
    // first read after r whose window offset exceeds win_offsets[r] - 1 ... i.e. first r' > r with windows
    uint64_t lo = r + 1, hi = n_reads;  // answer in [lo, hi]; n_reads = none
    const uint64_t v = win_offsets[r];
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (win_offsets[mid + 1] > v) hi = mid; else lo = mid + 1;
    }
    emit_offsets[r] = lo < n_reads ? emit_offsets[lo] : total_emitted;
}

cudaError_t launch_compact_backfill(const uint64_t* win_offsets, uint64_t n_reads, const unsigned long long* total_emitted, uint64_t* emit_offsets,
                                    cudaStream_t st) {
    compact_backfill_kernel<<<(unsigned)((n_reads + 255) / 256), 256, 0, st>>>(win_offsets, n_reads, total_emitted, emit_offsets);
    return cudaGetLastError();
}

}  // namespace kmb
