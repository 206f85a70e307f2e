// kmb_compact.cuh -- iterator-identical (compacted) output for K <= 32.
//
// CanonicalKmerIterator yields CanonicalKmerPos{km, pos} only for the windows that hold no invalid
// base, in increasing pos (naive_impl/canonical_kmer_iterator.rs:13-16, 42-70).  This engine writes
// exactly that sequence for every read of the batch, back to back in read order:
//   pos_out[i]   = pos (i32, as the reference)         canon_out[i] = get_canonical_word()
//   hash_out[i]  = hash_one(LexHasherState(k), canon)  emit_offsets[r] = index of read r's first entry
// Two launches of the same geometry: COUNT_ONLY counts the valid windows of every CTA (reads the bases
// only); after an exclusive scan of the CTA counts the emit launch re-derives the per-item offsets with a
// CTA-wide scan and stores each valid window at its final index.
#pragma once
#include <cstddef>

#include "kmb_extract.cuh"

namespace kmb {

struct CompactOut {
    uint64_t* canon;
    uint64_t* hash;
    int32_t* pos;
    uint64_t* emit_offsets;              // n_reads + 1 (entry n_reads is written by the host)
    unsigned long long* cta_counts;      // COUNT_ONLY: valid windows per CTA (out); emit: exclusive scan of them (in)
};

struct CompactParams {
    WinConst wc;
    CompactOut out;
};

constexpr int kCompactRound = kExtractThreads * kRun;  // entries one round of items can emit (2048)

// Shared memory of the compaction kernels beyond the tile: per-item counts, and the staging buffers through which
// a round's entries reach global memory as coalesced stores (each thread's <= 8 entries land at arbitrary,
// unaligned indices; storing them directly would touch every 32-byte sector 4-8 times).
struct CompactShared {
    uint32_t cnt[kItemsPerCta + 1];
    uint32_t round_off[kItemsPerCta / kExtractThreads + 2];  // exclusive offset of every round's first item, and the total
    uint32_t warp_tot[kExtractThreads / 32];
    // staging buffers of the emit launch; the counting launch allocates the struct only up to here (kCompactCountBytes)
    // (no hash buffer: LexHash(canon) is one pair reversal, computed when the entry leaves -- cheaper than staging it)
    alignas(16) uint64_t canon[kCompactRound];
    int32_t pos[kCompactRound];
};
constexpr size_t kCompactCountBytes = offsetof(CompactShared, canon);

template <bool VALIDATE, bool KHI, bool COUNT_ONLY>
struct CompactEng {
    using Params = CompactParams;
    using Span = kmb::Span;
    static constexpr bool kValidate = VALIDATE;
    using Shape = ShapeRun;
    static constexpr bool kTwoPhase = true, kCountOnly = COUNT_ONLY;
    static constexpr int kSpanEntries = 4;
    const CompactParams& p;
    CompactShared& sh;
    uint64_t pass_base = 0;      // valid windows of this CTA's passes so far
    uint64_t cur_pass_base = 0;  // ... before the current pass
    uint64_t cta_base = 0;       // valid windows of all earlier CTAs

    __device__ CompactEng(const CompactParams& params, CompactShared& shared) : p(params), sh(shared) {
        if (!COUNT_ONLY) cta_base = p.out.cta_counts[blockIdx.x];
    }
    __device__ __forceinline__ uint32_t K() const { return p.wc.K; }
    __device__ __forceinline__ Span load(const uint2* tile, uint32_t rel) const { return load_span<VALIDATE>(tile, rel, p.wc); }
    __device__ __forceinline__ bool dirty(const Span& s) const { return s.inv != 0ull; }

    __device__ __forceinline__ void begin_pass(uint32_t n_items) {
        for (uint32_t i = threadIdx.x; i <= n_items; i += blockDim.x) sh.cnt[i] = 0;
    }
    __device__ __forceinline__ void count(uint32_t li, uint32_t c) { sh.cnt[li] += c; }  // one thread owns item li

    // exclusive scan of cnt[0 .. n_items) in place; cnt[n_items] = total.  Called by all threads between barriers.
    __device__ __forceinline__ void scan(uint32_t n_items) {
        constexpr int PER = (kItemsPerCta + kExtractThreads - 1) / kExtractThreads;  // consecutive items per thread
        const uint32_t base = threadIdx.x * PER;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = (base + i < n_items) ? sh.cnt[base + i] : 0u; sum += v[i]; }
        uint32_t incl = sum;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) sh.warp_tot[warp] = incl;
        __syncthreads();
        uint32_t before = 0, total = 0;
        for (int w = 0; w < kExtractThreads / 32; ++w) { const uint32_t t = sh.warp_tot[w]; if (w < warp) before += t; total += t; }
        uint32_t run = before + incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) { if (base + i < n_items) sh.cnt[base + i] = run; run += v[i]; }
        if (threadIdx.x == 0) sh.cnt[n_items] = total;
        __syncthreads();
        // offsets at which the rounds of items start (the emit phase bumps cnt[] for single-window items)
        const uint32_t rounds = (n_items + kExtractThreads - 1) / kExtractThreads;
        if (threadIdx.x <= rounds) sh.round_off[threadIdx.x] = sh.cnt[min(threadIdx.x * kExtractThreads, n_items)];
        cur_pass_base = pass_base;
        pass_base += total;
    }

    // Where entry i of a round sits in the staging buffers.  A thread stages its <= 8 entries at consecutive indices, so
    // lanes are 8 entries apart: unswizzled, 16 lanes of a warp would hit the same pair of banks (ncu: 72 % of the
    // shared-memory wavefronts were conflicts).  XORing the low 4 index bits with the next 4 spreads both this
    // stride-8 write pattern and the linear read-out of round_end over all banks.
    __device__ static __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 4) & 15u); }

    // stage one entry of the current round (local index = its offset inside the round)
    __device__ __forceinline__ void put(uint32_t local, const Window& w, uint64_t pos) const {
        const uint32_t i = swz(local);
        sh.canon[i] = w.canon;
        sh.pos[i] = (int32_t)pos;
    }

    template <bool TWO, bool CHECK>
    __device__ __forceinline__ void run(const Span& a, const Span& b, uint32_t n_first, uint64_t, uint32_t nwin, const ItemCtx& ic) {
        const uint32_t off = sh.cnt[ic.li];
        uint32_t local = off - sh.round_off[ic.li / kExtractThreads];
        const uint64_t o0 = cta_base + cur_pass_base + off;  // global index of this item's first entry
        if (ic.pos_a == 0 && p.out.emit_offsets) p.out.emit_offsets[ic.r_a] = o0;  // this item opens read r_a
        uint32_t emitted = 0;
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            if ((uint32_t)j < nwin) {
                const bool second = TWO && (uint32_t)j >= n_first;
                if (TWO && (uint32_t)j == n_first && p.out.emit_offsets) p.out.emit_offsets[ic.r_b] = o0 + emitted;  // opens read r_b
                const Span s = second ? b : a;
                const bool ok = !CHECK || (((uint32_t)(s.inv >> j)) & p.wc.kmask) == 0u;
                if (ok) {
                    put(local + emitted, make_window<KHI>(s, j, p.wc), second ? (uint64_t)(j - n_first) : ic.pos_a + j);
                    ++emitted;
                }
            }
        }
    }
    __device__ __forceinline__ void single(const uint2* tile, uint32_t rel, uint64_t, const ItemCtx& ic) {
        // windows of one item arrive in order; the running index lives in cnt[li] (owned by this thread)
        const uint32_t off = sh.cnt[ic.li];
        if (ic.pos_a == 0 && p.out.emit_offsets) p.out.emit_offsets[ic.r_a] = cta_base + cur_pass_base + off;
        const Span s = load_span<VALIDATE>(tile, rel, p.wc);
        const bool ok = !VALIDATE || (((uint32_t)s.inv) & p.wc.kmask) == 0u;
        if (ok) {
            put(off - sh.round_off[ic.li / kExtractThreads], make_window<KHI>(s, 0, p.wc), ic.pos_a);
            sh.cnt[ic.li] = off + 1;
        }
    }
    // all threads, once per round: the staged entries of the round leave as coalesced stores
    __device__ __forceinline__ void round_end(uint32_t q_round, uint32_t) {
        __syncthreads();
        const uint32_t lo = sh.round_off[q_round], n = sh.round_off[q_round + 1] - lo;
        const uint64_t g0 = cta_base + cur_pass_base + lo;
        for (uint32_t e = threadIdx.x; e < n; e += blockDim.x) {
            const uint32_t i = swz(e);
            const uint64_t c = sh.canon[i];
            if (p.out.canon) p.out.canon[g0 + e] = c;
            if (p.out.hash) p.out.hash[g0 + e] = pair_reverse64(c) >> (2 * (32 - p.wc.K));  // LexHasher::write_u64, hash.rs:60-71
            if (p.out.pos) p.out.pos[g0 + e] = sh.pos[i];
        }
        __syncthreads();  // the next round re-uses the staging buffers
    }
    __device__ __forceinline__ void finish() {
        if (COUNT_ONLY && threadIdx.x == 0) p.out.cta_counts[blockIdx.x] = pass_base;
    }
};

// dynamic shared memory: [tile (+ CSR tables)] then CompactShared
template <class Eng>
__global__ void __launch_bounds__(kExtractThreads) compact_fixed_kernel(const FixedGeom g, const EncDesc enc, const CompactParams ep,
                                                                         uint32_t tile_bytes) {
    extern __shared__ uint2 tile[];
    CompactShared& sh = *reinterpret_cast<CompactShared*>(reinterpret_cast<unsigned char*>(tile) + tile_bytes);
    Eng eng(ep, sh);
    fixed_body(g, enc, eng, tile, blockIdx.x);
    eng.finish();
}

template <class Eng>
__global__ void __launch_bounds__(kExtractThreads) compact_csr_kernel(const CsrGeom g, const EncDesc enc, const CompactParams ep,
                                                                       uint32_t tile_bytes) {
    extern __shared__ uint2 tile[];
    __shared__ CsrPass pass;
    uint64_t* c_off = reinterpret_cast<uint64_t*>(tile + g.tile_entries);
    uint64_t* c_win = c_off + (kCsrCache + 2);
    CompactShared& sh = *reinterpret_cast<CompactShared*>(reinterpret_cast<unsigned char*>(tile) + tile_bytes);
    Eng eng(ep, sh);
    csr_body(g, enc, eng, tile, c_off, c_win, &pass, blockIdx.x);
    eng.finish();
}

}  // namespace kmb
