// kmb_compact.cuh -- iterator-identical (compacted) output for K <= 32.
//
// CanonicalKmerIterator yields CanonicalKmerPos{km, pos} only for the windows that hold no invalid
// base, in increasing pos (naive_impl/canonical_kmer_iterator.rs:13-16, 42-70).  This engine writes
// exactly that sequence for every read of the batch, back to back in read order:
//   pos_out[i]   = pos (i32, as the reference)         canon_out[i] = get_canonical_word()
//   hash_out[i]  = hash_one(LexHasherState(k), canon)  emit_offsets[r] = index of read r's first entry
// ONE launch, the bases are read once: a CTA stages its tile, counts the valid windows of its items from the staged
// invalid masks, scans them CTA-wide, learns where its entries start in the output from a decoupled look-back over
// the earlier tiles (single-pass chained scan: every tile publishes {aggregate | inclusive prefix} in one 64-bit
// descriptor; tiles are handed out by an atomic ticket so a tile only ever waits for tiles that are already running),
// and stores each valid window at its final index.  COUNT_ONLY (the sizing call) just adds up the CTA totals.
// emit_offsets are written tile-local by the kernel and shifted by the tiles' starts afterwards (compact_fixup_kernel).
#pragma once
#include <cstddef>

#include "kmb_extract.cuh"

namespace kmb {

struct CompactOut {
    uint64_t* canon;
    uint64_t* hash;
    int32_t* pos;
    uint64_t* emit_offsets;              // n_reads + 1 (entry n_reads is written by the host)
    unsigned long long* desc;            // look-back descriptors, one per tile: status << 62 | value (zeroed before the launch)
    unsigned long long* ticket;          // next tile to hand out (zeroed before the launch)
    unsigned long long* total;           // COUNT_ONLY: += every CTA's count; emit: the last tile stores the grand total
    uint64_t capacity;                   // entries the output arrays hold: nothing is written at or beyond it
    uint32_t vec16;                      // canon / hash are 16-byte aligned and pos 8-byte aligned: pairs leave as vectors
};

constexpr unsigned long long kDescAggregate = 1ull << 62, kDescPrefix = 2ull << 62, kDescValue = (1ull << 62) - 1;
__device__ __forceinline__ unsigned long long desc_load(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void desc_store(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

struct CompactParams {
    WinConst wc;
    CompactOut out;
};

constexpr int kCompactRound = kExtractThreads * kRun;  // entries one round of items can emit (2048): 256 per warp
constexpr int kCompactWarps = kExtractThreads / 32;
constexpr int kWarpSlice = 32 * kRun;

// Shared memory of the compaction kernels beyond the tile: per-item counts, and the staging buffers through which
// a round's entries reach global memory as coalesced stores (each thread's <= 8 entries land at arbitrary,
// unaligned indices; storing them directly would touch every 32-byte sector 4-8 times).
struct CompactShared {
    uint32_t cnt[kItemsPerCta + 1];
    uint32_t wr_off[kItemsPerCta / 32 + 2];  // exclusive offset of the first item of every warp-round (32 items), and the total
    uint32_t warp_tot[kExtractThreads / 32];
    uint32_t tile_id;                 // this CTA's ticket
    uint32_t lb_has[kExtractThreads / 32];          // look-back: warp w's window holds a tile that knows its prefix
    unsigned long long lb_sum[kExtractThreads / 32];  // ... and the values of its window up to that tile
    // staging buffers of the emit launch; the counting launch allocates the struct only up to here (kCompactCountBytes).
    // One slice of kWarpSlice entries per warp: the 256 entries its 32 items can emit in a round.
    alignas(16) uint64_t canon[kCompactWarps * kWarpSlice];
    uint64_t hash[kCompactWarps * kWarpSlice];
    int32_t pos[kCompactWarps * kWarpSlice];
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
    uint64_t cta_base = 0;       // valid windows of all earlier tiles
    uint32_t tile_id = 0;
    bool placed = false;         // cta_base is known
    bool final_pass = false;     // thread 0: the tile's last pass has been counted and published

    __device__ CompactEng(const CompactParams& params, CompactShared& shared) : p(params), sh(shared) {}

    // Take the next tile.  Called by all threads at the top of the kernel; the ticket order is the scheduling order, so
    // every tile with a smaller id has started before this one and look-back cannot wait for a CTA that is not resident.
    __device__ __forceinline__ uint32_t take_tile() {
        if (COUNT_ONLY) { tile_id = blockIdx.x; return tile_id; }
        if (threadIdx.x == 0) sh.tile_id = (uint32_t)atomicAdd(p.out.ticket, 1ull);
        __syncthreads();
        tile_id = sh.tile_id;
        return tile_id;
    }

    // After a pass's scan (pass_base = this CTA's count so far, incl. the pass): tell the later tiles.  A tile whose start is
    // not known yet publishes its aggregate (single-pass tiles: all of them except CSR tiles whose reads are mostly shorter
    // than k; a multi-pass tile can only do so once its last pass is counted -- its successors wait); one that knows its
    // start publishes the inclusive prefix.
    __device__ __forceinline__ void publish(bool last_pass) {
        if (COUNT_ONLY || !last_pass || threadIdx.x != 0) return;
        if (!placed) {
            if (tile_id > 0) desc_store(p.out.desc + tile_id, kDescAggregate | pass_base);
        } else {
            desc_store(p.out.desc + tile_id, kDescPrefix | (cta_base + pass_base));
            if (tile_id + 1 == gridDim.x) *p.out.total = cta_base + pass_base;
        }
        final_pass = true;  // (thread 0 only: it is the one that publishes again in place())
    }

    // Where this tile's entries start in the output: decoupled look-back over the earlier tiles, by the whole CTA at once --
    // thread i reads the descriptor of tile (tile_id - 1 - i), so one round covers kExtractThreads predecessors with a
    // single global-load latency.  Called from the first round_end of the tile, i.e. AFTER the first round of windows has
    // been computed and staged: the tiles just ahead of this one, which started moments earlier, have had that long to
    // count and publish, so the CTA rarely has to wait here (looking back right after the scan cost 22 % of the kernel in
    // spinning: ncu kernels_r02f).  Nothing before the write-out needs the result.
    __device__ __forceinline__ void place() {
        const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
        unsigned long long excl = 0;
        int64_t idx = (int64_t)tile_id - 1;
        for (;;) {  // CTA-uniform
            if (idx < 0) break;
            const int64_t mine = idx - (int64_t)threadIdx.x;
            unsigned long long d = kDescPrefix;  // before tile 0: "prefix 0"
            if (mine >= 0) {
                do { d = desc_load(p.out.desc + mine); } while ((d >> 62) == 0ull);
            }
            const unsigned pf = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
            // value of this warp's window up to (and including) its nearest prefix holder
            const unsigned upto = pf ? (unsigned)__ffs(pf) - 1u : 31u;
            unsigned long long v = lane <= upto ? (d & kDescValue) : 0ull;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) { sh.lb_sum[warp] = v; sh.lb_has[warp] = pf != 0u; }
            __syncthreads();
            bool found = false;
            for (int w = 0; w < kExtractThreads / 32 && !found; ++w) {  // nearest warps first; stop at the first prefix
                excl += sh.lb_sum[w];
                found = sh.lb_has[w] != 0u;
            }
            __syncthreads();
            if (found) break;
            idx -= kExtractThreads;
        }
        cta_base = excl;
        placed = true;
        if (final_pass) {  // thread 0 of a tile whose last pass is the current one: the count it published was complete
            desc_store(p.out.desc + tile_id, kDescPrefix | (cta_base + pass_base));
            if (tile_id + 1 == gridDim.x) *p.out.total = cta_base + pass_base;
        }
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
        // offsets at which the warp-rounds (32 consecutive items = one warp in one round) start; a copy, because the emit
        // phase bumps cnt[] for single-window items
        const uint32_t wrs = (n_items + 31) / 32;
        if (threadIdx.x <= wrs) sh.wr_off[threadIdx.x] = sh.cnt[min(threadIdx.x * 32u, n_items)];
        cur_pass_base = pass_base;
        pass_base += total;
    }

    // Where entry i of a round sits in the staging buffers.  A thread stages its <= 8 entries at consecutive indices, so
    // lanes are 8 entries apart: unswizzled, 16 lanes of a warp would hit the same pair of banks (ncu: 72 % of the
    // shared-memory wavefronts were conflicts).  XORing the low 4 index bits with the next 4 spreads both this
    // stride-8 write pattern and the linear read-out of round_end over all banks.
    __device__ static __forceinline__ uint32_t swz(uint32_t i) { return i ^ ((i >> 4) & 15u); }

    // stage one entry of this warp's current round (local index = its offset inside the warp-round, < 256): every warp
    // owns a 256-entry slice of the staging buffers, so a round needs no CTA-wide barrier
    __device__ __forceinline__ void put(uint32_t local, const Window& w, uint64_t pos) const {
        const uint32_t i = (threadIdx.x >> 5) * kWarpSlice + swz(local);
        sh.canon[i] = w.canon;
        sh.hash[i] = w.hash;  // the XOR fold of make_window: two LOP3, against ten for a pair reversal when the entry leaves
        sh.pos[i] = (int32_t)pos;
    }

    template <bool TWO, bool CHECK>
    __device__ __forceinline__ void run(const Span& a, const Span& b, uint32_t n_first, uint64_t, uint32_t nwin, const ItemCtx& ic) {
        const uint32_t off = sh.cnt[ic.li];
        const uint32_t wlo = sh.wr_off[ic.li >> 5];
        // index of this item's first entry, counted from the tile's first: the tile's own start is not known yet (place()) and
        // is added to emit_offsets afterwards (compact_fixup_kernel)
        const uint64_t o0 = cur_pass_base + off;
        uint32_t local = off - wlo;  // position inside the warp's staging slice
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
        if (ic.pos_a == 0 && p.out.emit_offsets) p.out.emit_offsets[ic.r_a] = cur_pass_base + off;  // tile-local, see run()
        const Span s = load_span<VALIDATE>(tile, rel, p.wc);
        const bool ok = !VALIDATE || (((uint32_t)s.inv) & p.wc.kmask) == 0u;
        if (ok) {
            put(off - sh.wr_off[ic.li >> 5], make_window<KHI>(s, 0, p.wc), ic.pos_a);
            sh.cnt[ic.li] = off + 1;
        }
    }
    // all threads, once per round: every warp writes the entries its 32 items staged -- one contiguous run of the output --
    // two entries per lane and step: an aligned pair of the output arrays leaves as one 16-byte store per 8-byte array and
    // one 8-byte store for the positions (the pair's two entries are wherever the parity of the run's first index puts them
    // in the slice).  The tile's start is looked up on the way into the first round's write-out.
    __device__ __forceinline__ void round_end(uint32_t q_round, uint32_t n_items) {
        if (!placed) place();  // CTA-uniform: contains barriers
        __syncwarp();
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        const uint32_t wr = q_round * kCompactWarps + warp;  // this warp-round's index among the pass's
        if (wr * 32u < n_items) {
            const uint32_t lo = sh.wr_off[wr], n = sh.wr_off[wr + 1] - lo;
            const uint64_t g0 = cta_base + cur_pass_base + lo;
            // the caller's arrays may be too small: the host reports it, nothing is overrun
            const uint32_t n_ok = g0 >= p.out.capacity ? 0u : (uint32_t)min((uint64_t)n, p.out.capacity - g0);
            const uint32_t par = (uint32_t)(g0 & 1ull);
            const uint64_t* sc = sh.canon + warp * kWarpSlice;
            const uint64_t* shh = sh.hash + warp * kWarpSlice;
            const int32_t* sp = sh.pos + warp * kWarpSlice;
            // aligned position i <-> global index gb + i <-> staged entry i - par; entries [0, n_ok) exist
            const uint64_t gb = g0 - par;
            uint64_t* gc = p.out.canon ? p.out.canon + gb : nullptr;
            uint64_t* gh = p.out.hash ? p.out.hash + gb : nullptr;
            int32_t* gp = p.out.pos ? p.out.pos + gb : nullptr;
            const uint32_t end = par + n_ok;
            // full aligned pairs: positions [i_lo, i_hi); at most one lone entry on either side of them
            const uint32_t i_lo = 2 * par, i_hi = end & ~1u;
            if (p.out.vec16) {  // the output arrays are 16-byte (positions: 8-byte) aligned
                for (uint32_t i = i_lo + 2 * lane; i < i_hi; i += 64) {
                    const uint32_t i0 = swz(i - par), i1 = swz(i + 1 - par);
                    if (gc) *reinterpret_cast<ulonglong2*>(gc + i) = make_ulonglong2(sc[i0], sc[i1]);
                    if (gh) *reinterpret_cast<ulonglong2*>(gh + i) = make_ulonglong2(shh[i0], shh[i1]);
                    if (gp) *reinterpret_cast<int2*>(gp + i) = make_int2(sp[i0], sp[i1]);
                }
            } else {
                for (uint32_t i = i_lo + lane; i < i_hi; i += 32) {
                    const uint32_t i0 = swz(i - par);
                    if (gc) gc[i] = sc[i0];
                    if (gh) gh[i] = shh[i0];
                    if (gp) gp[i] = sp[i0];
                }
            }
            // the lone entries: position 1 when the run starts on an odd index, position end - 1 when it ends on an even one
            uint32_t lone = 0xFFFFFFFFu;
            if (lane == 0 && par && end > 1u) lone = 1u;
            if (lane == 1 && (end & 1u) && end - 1u >= i_lo && end - 1u >= par) lone = end - 1u;
            if (lone != 0xFFFFFFFFu) {
                const uint32_t i0 = swz(lone - par);
                if (gc) gc[lone] = sc[i0];
                if (gh) gh[lone] = shh[i0];
                if (gp) gp[lone] = sp[i0];
            }
        }
        __syncwarp();  // the next round re-uses this warp's slice
    }
    __device__ __forceinline__ void finish() {
        if (COUNT_ONLY && threadIdx.x == 0 && pass_base) atomicAdd(p.out.total, (unsigned long long)pass_base);
    }
};

// dynamic shared memory: [tile (+ CSR tables)] then CompactShared
template <class Eng>
__global__ void __launch_bounds__(kExtractThreads) compact_fixed_kernel(const FixedGeom g, const EncDesc enc, const CompactParams ep,
                                                                         uint32_t tile_bytes) {
    extern __shared__ uint2 tile[];
    CompactShared& sh = *reinterpret_cast<CompactShared*>(reinterpret_cast<unsigned char*>(tile) + tile_bytes);
    Eng eng(ep, sh);
    fixed_body(g, enc, eng, tile, eng.take_tile());
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
    csr_body(g, enc, eng, tile, c_off, c_win, &pass, eng.take_tile());
    eng.finish();
}

}  // namespace kmb
