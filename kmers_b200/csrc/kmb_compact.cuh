// kmb_compact.cuh -- iterator-identical (compacted) output for K <= 32.
//
// CanonicalKmerIterator yields CanonicalKmerPos{km, pos} only for the windows that hold no invalid
// base, in increasing pos (naive_impl/canonical_kmer_iterator.rs:13-16, 42-70).  This engine writes
// exactly that sequence for every read of the batch, back to back in read order:
//   pos_out[i]   = pos (i32, as the reference)         canon_out[i] = get_canonical_word()
//   hash_out[i]  = hash_one(LexHasherState(k), canon)  emit_offsets[r] = index of read r's first entry
// ONE launch, the bases are read once: a CTA stages a tile, counts the valid windows of its items from the staged
// invalid masks, scans them CTA-wide, learns where its entries start in the output from a decoupled look-back over
// the earlier tiles (single-pass chained scan: every tile publishes {aggregate | inclusive prefix} in one 64-bit
// descriptor; tiles are handed out by an atomic ticket so a tile only ever waits for tiles that are already running),
// and stores each valid window at its final index.  COUNT_ONLY (the sizing call) just adds up the CTA totals.
//
// Two kernels share the engine:
//   compact_fixed_pipe_kernel (fixed-length reads, PIPE): persistent CTAs, software-pipelined over tiles -- a CTA counts tile
//       i+1 and publishes its aggregate BEFORE it emits tile i, and one of its warps looks back for tile i+1 while the others
//       emit tile i.  Every descriptor a tile needs has been published a whole emit phase earlier, so nobody waits, and a
//       tile's start is known before its first entry is staged.
//   compact_fixed_kernel / compact_csr_kernel (one tile per CTA): the tile's aggregate goes out after the scan and the CTA
//       looks back on the way into its first write-out.  Descriptors of tiles that started moments earlier are often not
//       there yet: measured 18-24 % of the stall samples at that point (ncu kernels_r02h/r02i), which is what PIPE removes.
//       emit_offsets are written tile-local and shifted by the tiles' starts afterwards (compact_fixup_kernel).
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
    uint32_t all_vec;                    // ... and all three arrays are asked for
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

#ifndef KMB_COMPACT_MINCTAS
#define KMB_COMPACT_MINCTAS 3
#endif
constexpr int kCompactRound = kExtractThreads * kRun;  // entries one round of items can emit (2048): 256 per warp
constexpr int kCompactWarps = kExtractThreads / 32;
constexpr int kWarpSlice = 32 * kRun + 2;  // the 256 entries a warp's 32 items can emit in a round, + 1 for the parity shift (put)
constexpr int kLookBack = 8;               // PIPE: descriptors per lane and look-back step (256 earlier tiles per load latency)
// Items per tile of the pipelined kernel.  It runs three CTAs per SM (the 64 registers that four allow cost more in spills and
// rebuilt addresses than the fourth CTA brings: 2.133 against 2.164 ms per 4 M reads), and the shared memory that frees goes
// into larger tiles: what a CTA does once per tile -- ticket, staging loads, scan, the barriers between them -- is 16 % of
// the instructions but a third of the stall samples.  Per 4 M reads: 1024 items 2.128 ms, 1536 2.080, 2048 2.014, 2304 (the
// most that fits three CTAs) 1.981; two CTAs per SM with 3072 / 4096 items 2.115 / 2.144.
#ifndef KMB_COMPACT_PIPE_ITEMS
#define KMB_COMPACT_PIPE_ITEMS 2304
#endif
constexpr int kCompactPipeItems = KMB_COMPACT_PIPE_ITEMS;
#ifndef KMB_COMPACT_CSR_ITEMS
#define KMB_COMPACT_CSR_ITEMS 1536  // ragged reads, one tile per CTA: as the dense ragged kernels (kCsrItemsPerCta); still three CTAs per SM
#endif
constexpr int kCompactCsrItems = KMB_COMPACT_CSR_ITEMS;
static_assert(kItemsPerCta * kRun < 65536 && kCompactPipeItems * kRun < 65536, "per-tile offsets are 16-bit");
static_assert(kCompactPipeItems % kExtractThreads == 0, "whole rounds");

// per-tile counters: PIPE keeps two (the tile being emitted and the one counted ahead)
template <int ITEMS>
struct CompactBufT {
    uint16_t cnt[ITEMS + 2];          // per item: count, then (scan) exclusive offset inside the pass; [n_items] = total
    uint32_t wr_off[ITEMS / 32 + 2];  // exclusive offset of the first item of every warp-round (32 items), and the total
};

// Shared memory of the compaction kernels beyond the tile: per-item counts, and the staging buffers through which
// a round's entries reach global memory as coalesced stores (each thread's <= 8 entries land at arbitrary,
// unaligned indices; storing them directly would touch every 32-byte sector 4-8 times).
template <int ITEMS>
struct CompactSharedT {
    CompactBufT<ITEMS> buf;
    uint32_t warp_tot[kCompactWarps];
    uint32_t tile_id;                            // the ticket just taken
    uint32_t lb_has[kCompactWarps];              // CTA-wide look-back: warp w's window holds a tile that knows its prefix
    unsigned long long lb_sum[kCompactWarps];    // ... and the values of its window up to that tile
    // staging buffers of the emit launch; the counting launch allocates the struct only up to here (kCompactCountBytes).
    // One slice of kWarpSlice entries per warp.
    alignas(16) uint64_t canon[kCompactWarps * kWarpSlice];
    uint64_t hash[kCompactWarps * kWarpSlice];
    int32_t pos[kCompactWarps * kWarpSlice];
};
using CompactShared = CompactSharedT<kItemsPerCta>;
constexpr size_t kCompactCountBytes = offsetof(CompactShared, canon);
static_assert((kWarpSlice * sizeof(uint64_t)) % 16 == 0 && (kWarpSlice * sizeof(int32_t)) % 8 == 0, "slices stay aligned for vector loads");

// PIPE: the second counter set and what the CTA remembers about its two tiles
struct CompactPipeShared {
    CompactSharedT<kCompactPipeItems> s;
    CompactBufT<kCompactPipeItems> buf1;
    FixedTile ft[2];
    unsigned long long base[2];   // where the tile's entries start
    uint32_t next_tile;           // the tile counted ahead, its count and its buffer: what warp 0 looks back for during an emit phase
    uint32_t next_total;
    uint32_t next_b;
};

template <bool VALIDATE, bool KHI, bool COUNT_ONLY, bool PIPE = false, int ITEMS = kItemsPerCta>
struct CompactEng {
    using Shared = CompactSharedT<ITEMS>;
    using Buf = CompactBufT<ITEMS>;
    using Params = CompactParams;
    using Span = kmb::Span;
    static constexpr bool kValidate = VALIDATE;
    using Shape = ShapeRun;
    static constexpr bool kTwoPhase = true, kCountOnly = COUNT_ONLY;
    static constexpr int kSpanEntries = 4;
    const CompactParams& p;
    Shared& sh;
    Buf* cb;                     // the counters of the tile at hand
    uint64_t pass_base = 0;      // valid windows of this tile's passes so far
    uint64_t cur_pass_base = 0;  // ... before the current pass
    uint64_t cta_base = 0;       // valid windows of all earlier tiles
    uint32_t tile_id = 0;
    uint32_t n_tiles = 0;
    uint32_t sp = 0;             // this round's entries are staged one slot up (see round_begin)
    uint32_t lb_round = 0xFFFFFFFFu;  // PIPE: the round of the emit phase after which warp 0 looks back for the next tile
    bool placed = false;         // cta_base is known
    bool fits = false;           // PIPE: the whole tile lies below p.out.capacity
    bool final_pass = false;     // thread 0: the tile's last pass has been counted and published

    __device__ CompactEng(const CompactParams& params, Shared& shared) : p(params), sh(shared), cb(&shared.buf), n_tiles(gridDim.x) {}

    // Take the next tile.  Called by all threads; the ticket order is the scheduling order, so every tile with a smaller
    // id has been taken by a CTA that is running, and look-back cannot wait for a CTA that is not resident.
    __device__ __forceinline__ uint32_t take_tile() {
        if (COUNT_ONLY) { tile_id = blockIdx.x; return tile_id; }
        if (threadIdx.x == 0) sh.tile_id = (uint32_t)atomicAdd(p.out.ticket, 1ull);
        __syncthreads();
        tile_id = sh.tile_id;
        return tile_id;
    }

    // ---- one tile per CTA
    // After a pass's scan (pass_base = this CTA's count so far, incl. the pass): tell the later tiles.  A tile whose start is
    // not known yet publishes its aggregate (single-pass tiles: all of them except CSR tiles whose reads are mostly shorter
    // than k; a multi-pass tile can only do so once its last pass is counted -- its successors wait); one that knows its
    // start publishes the inclusive prefix.
    __device__ __forceinline__ void publish(bool last_pass) {
        if (COUNT_ONLY || PIPE || !last_pass || threadIdx.x != 0) return;
        if (!placed) {
            if (tile_id > 0) desc_store(p.out.desc + tile_id, kDescAggregate | pass_base);
        } else {
            desc_store(p.out.desc + tile_id, kDescPrefix | (cta_base + pass_base));
            if (tile_id + 1 == n_tiles) *p.out.total = cta_base + pass_base;
        }
        final_pass = true;  // (thread 0 only: it is the one that publishes again in place())
    }

    // Where this tile's entries start in the output: decoupled look-back over the earlier tiles, by the whole CTA at once --
    // thread i reads the descriptor of tile (tile_id - 1 - i), so one round covers kExtractThreads predecessors with a
    // single global-load latency.  Called from the first round_end of the tile, i.e. AFTER the first round of windows has
    // been computed and staged: the tiles just ahead of this one, which started moments earlier, have had that long to
    // count and publish (looking back right after the scan cost 22 % of the kernel in spinning, ncu kernels_r02f; one warp
    // looking back early while the others compute, and handing the result over, was slower still: kernels_r02i).
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
            for (int w = 0; w < kCompactWarps && !found; ++w) {  // nearest warps first; stop at the first prefix
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
            if (tile_id + 1 == n_tiles) *p.out.total = cta_base + pass_base;
        }
    }

    // ---- PIPE
    // the counters and identity of the tile the next count / emit phase works on
    __device__ __forceinline__ void bind(Buf* buf, uint32_t tile) { cb = buf; tile_id = tile; pass_base = 0; cur_pass_base = 0; }
    // after count_pass of a tile counted ahead: its aggregate goes out at once (tile 0 starts at 0 and only ever holds a prefix)
    __device__ __forceinline__ void publish_count() const {
        if (threadIdx.x == 0 && tile_id > 0) desc_store(p.out.desc + tile_id, kDescAggregate | pass_base);
    }
    // One warp: where tile t (total entries `total`) starts.  Lane i reads the descriptors of tiles (t - 1 - i - 32 m), all
    // kLookBack loads in flight at once: the tiles that have not been placed themselves -- those counted within the last few
    // microseconds -- cost one load latency.  Publishes t's inclusive prefix and leaves the start in *slot.
    __device__ __forceinline__ void place_tile(uint32_t t, uint32_t total, unsigned long long* slot) const {
        const unsigned lane = threadIdx.x & 31u;
        unsigned long long excl = 0;
        int64_t idx = (int64_t)t - 1;
        bool found = idx < 0;
        // one descriptor per lane: wait until its tile has counted, then add what it contributes; true = one held a prefix
        auto take = [&](unsigned long long d, int64_t mine) {
            while ((d >> 62) == 0ull) d = desc_load(p.out.desc + mine);
            const unsigned pf = __ballot_sync(0xffffffffu, (d >> 62) == 2ull);
            const unsigned upto = pf ? (unsigned)__ffs(pf) - 1u : 32u;  // nearest tile that knows its prefix
            // aggregates of the tiles nearer than that one: at most 32 tiles of kItemsPerCta * kRun windows each
            uint32_t v = lane < upto ? (uint32_t)d : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            excl += v;
            if (pf) excl += __shfl_sync(0xffffffffu, d, (int)upto) & kDescValue;
            return pf != 0u;
        };
        while (!found) {  // warp-uniform
            unsigned long long d[kLookBack];
#pragma unroll
            for (int m = 0; m < kLookBack; ++m) {
                const int64_t mine = idx - (int64_t)lane - 32 * m;
                d[m] = mine >= 0 ? desc_load(p.out.desc + mine) : kDescPrefix;  // before tile 0: "prefix 0"
            }
#pragma unroll
            for (int m = 0; m < kLookBack; ++m)
                if (!found) found = take(d[m], idx - (int64_t)lane - 32 * m);
            idx -= 32 * kLookBack;
        }
        if (lane == 0) {
            *slot = excl;
            desc_store(p.out.desc + t, kDescPrefix | (excl + total));
            if (t + 1 == n_tiles) *p.out.total = excl + total;
        }
    }
    // ahead of a tile's emit phase: its count, its start, and after which round warp 0 places the next tile
    __device__ __forceinline__ void begin_emit(uint32_t total, unsigned long long base, uint32_t look_back_round) {
        pass_base = total; cur_pass_base = 0; cta_base = base; placed = true; lb_round = look_back_round;
        fits = base + total <= p.out.capacity;
    }

    __device__ __forceinline__ uint32_t K() const { return p.wc.K; }
    __device__ __forceinline__ Span load(const uint2* tile, uint32_t rel) const { return load_span<VALIDATE>(tile, rel, p.wc); }
    __device__ __forceinline__ bool dirty(const Span& s) const { return s.inv != 0ull; }

    __device__ __forceinline__ void begin_pass(uint32_t n_items) {
        for (uint32_t i = threadIdx.x; i <= n_items; i += blockDim.x) cb->cnt[i] = 0;
    }
    __device__ __forceinline__ void count(uint32_t li, uint32_t c) { cb->cnt[li] += (uint16_t)c; }  // one thread owns item li

    // exclusive scan of cnt[0 .. n_items) in place; cnt[n_items] = total.  Called by all threads between barriers.
    __device__ __forceinline__ void scan(uint32_t n_items) {
        constexpr int PER = (ITEMS + kExtractThreads - 1) / kExtractThreads;  // consecutive items per thread
        const uint32_t base = threadIdx.x * PER;
        uint32_t v[PER], sum = 0;
#pragma unroll
        for (int i = 0; i < PER; ++i) { v[i] = (base + i < n_items) ? cb->cnt[base + i] : 0u; sum += v[i]; }
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
        for (int w = 0; w < kCompactWarps; ++w) { const uint32_t t = sh.warp_tot[w]; if (w < warp) before += t; total += t; }
        uint32_t run = before + incl - sum;
#pragma unroll
        for (int i = 0; i < PER; ++i) { if (base + i < n_items) cb->cnt[base + i] = (uint16_t)run; run += v[i]; }
        if (threadIdx.x == 0) cb->cnt[n_items] = (uint16_t)total;
        __syncthreads();
        // offsets at which the warp-rounds (32 consecutive items = one warp in one round) start; a copy, because the emit
        // phase bumps cnt[] for single-window items
        const uint32_t wrs = (n_items + 31) / 32;
        if (threadIdx.x <= wrs) cb->wr_off[threadIdx.x] = cb->cnt[min(threadIdx.x * 32u, n_items)];
        cur_pass_base = pass_base;
        pass_base += total;
    }

    // all threads, ahead of a round's items.  Where the warp's run of entries starts in the output decides how its slice is
    // laid out: with the first entry on an odd index everything is staged one slot up, so that an aligned pair of the output
    // arrays is an aligned pair of the slice and leaves with one 16-byte load and one 16-byte store per array.  A CTA that
    // does not know the tile's start yet (one tile per CTA: its first round) stages unshifted and pays for it in round_end
    // if the index turns out odd.
    __device__ __forceinline__ void round_begin(uint32_t q_round) {
        if (COUNT_ONLY) return;
        const uint32_t wr = q_round * kCompactWarps + (threadIdx.x >> 5);
        sp = placed ? (uint32_t)((cta_base + cur_pass_base + cb->wr_off[wr]) & 1ull) : 0u;
    }

    // Where slot i of a warp's slice sits in the staging buffers.  A thread stages its <= 8 entries at consecutive slots, so
    // lanes are 8 slots apart: unswizzled, 16 lanes of a warp would hit the same pair of banks (ncu: 72 % of the
    // shared-memory wavefronts were conflicts).  XORing bits 1-3 of the slot with the next three spreads both this
    // stride-8 write pattern and the linear read-out of round_end over all banks, and keeps the two slots of an aligned pair
    // next to each other (bit 0 is left alone): the pair is one 16-byte load.
    __device__ static __forceinline__ uint32_t swz(uint32_t i) { return i ^ (((i >> 4) & 7u) << 1); }

    // stage one entry of this warp's current round (local index = its offset inside the warp-round, < 256): every warp
    // owns a slice of the staging buffers, so a round needs no CTA-wide barrier
    __device__ __forceinline__ void put(uint32_t local, const Window& w, uint64_t pos) const {
        const uint32_t i = (threadIdx.x >> 5) * kWarpSlice + swz(local + sp);
        sh.canon[i] = w.canon;
        sh.hash[i] = w.hash;  // the XOR fold of make_window: two LOP3, against ten for a pair reversal when the entry leaves
        sh.pos[i] = (int32_t)pos;
    }

    template <bool TWO, bool CHECK>
    __device__ __forceinline__ void run(const Span& a, const Span& b, uint32_t n_first, uint64_t, uint32_t nwin, const ItemCtx& ic) {
        const uint32_t off = cb->cnt[ic.li];
        const uint32_t wlo = cb->wr_off[ic.li >> 5];
        // index of this item's first entry.  One tile per CTA: counted from the tile's first -- the tile's own start is not known
        // yet in the first round and is added to emit_offsets afterwards (compact_fixup_kernel).  PIPE: final.
        const uint64_t o0 = (PIPE ? cta_base : 0ull) + cur_pass_base + off;
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
        const uint32_t off = cb->cnt[ic.li];
        if (ic.pos_a == 0 && p.out.emit_offsets) p.out.emit_offsets[ic.r_a] = (PIPE ? cta_base : 0ull) + cur_pass_base + off;  // see run()
        const Span s = load_span<VALIDATE>(tile, rel, p.wc);
        const bool ok = !VALIDATE || (((uint32_t)s.inv) & p.wc.kmask) == 0u;
        if (ok) {
            put(off - cb->wr_off[ic.li >> 5], make_window<KHI>(s, 0, p.wc), ic.pos_a);
            cb->cnt[ic.li] = (uint16_t)(off + 1);
        }
    }

    // the aligned pairs of a warp-round whose slice is laid out like the output (sp == par): slot i <-> output index gb + i.
    // Four steps of 32 pairs cover the slice; the swizzle of a lane's pair only differs by a constant between the steps.
    template <bool ALL>
    __device__ __forceinline__ void write_pairs(uint32_t lane, uint32_t i_lo, uint32_t i_hi, const uint64_t* sc, const uint64_t* shh,
                                                const int32_t* spos, uint64_t* gc, uint64_t* gh, int32_t* gp) const {
        const uint32_t b0 = (2u * lane) ^ (((lane >> 3) & 7u) << 1), b1 = b0 ^ 8u;
        // this lane's first pair in the three arrays; the steps are immediate offsets from here.  Made opaque to the compiler,
        // which otherwise rebuilds all three 64-bit addresses from (gb + i) in every step to save six registers: 10 of the 21
        // instructions of a step (ncu kernels_r02n)
        uint64_t* pc = gc + 2u * lane;
        uint64_t* ph = gh + 2u * lane;
        int32_t* pp = gp + 2u * lane;
        asm volatile("" : "+l"(pc), "+l"(ph), "+l"(pp));
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const uint32_t i = 2u * lane + 64u * s;
            if (i >= i_lo && i < i_hi) {
                const uint32_t q = ((s & 1) ? b1 : b0) + 64u * s;  // = swz(i)
                if (ALL || gc) *reinterpret_cast<ulonglong2*>(pc + 64 * s) = *reinterpret_cast<const ulonglong2*>(sc + q);
                if (ALL || gh) *reinterpret_cast<ulonglong2*>(ph + 64 * s) = *reinterpret_cast<const ulonglong2*>(shh + q);
                if (ALL || gp) *reinterpret_cast<int2*>(pp + 64 * s) = *reinterpret_cast<const int2*>(spos + q);
            }
        }
    }

    // all threads, once per round: every warp writes the entries its 32 items staged -- one contiguous run of the output --
    // two entries per lane and step: an aligned pair of the output arrays leaves as one 16-byte store per 8-byte array and
    // one 8-byte store for the positions.
    __device__ __forceinline__ void round_end(uint32_t q_round, uint32_t n_items) {
        if (!PIPE && !placed) place();  // CTA-uniform: contains barriers
        __syncwarp();
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        const uint32_t wr = q_round * kCompactWarps + warp;  // this warp-round's index among the pass's
        if (wr * 32u < n_items) {
            const uint32_t lo = cb->wr_off[wr], n = cb->wr_off[wr + 1] - lo;
            const uint64_t g0 = cta_base + cur_pass_base + lo;
            // the caller's arrays may be too small: the host reports it, nothing is overrun
            const uint32_t n_ok = (PIPE && fits) ? n : g0 >= p.out.capacity ? 0u : (uint32_t)min((uint64_t)n, p.out.capacity - g0);
            const uint32_t par = (uint32_t)(g0 & 1ull);
            const uint64_t* sc = sh.canon + warp * kWarpSlice;
            const uint64_t* shh = sh.hash + warp * kWarpSlice;
            const int32_t* spos = sh.pos + warp * kWarpSlice;
            // aligned position i <-> output index gb + i <-> entry i - par <-> slot i - par + sp; entries [0, n_ok) exist
            const uint64_t gb = g0 - par;
            const uint32_t end = par + n_ok;
            const uint32_t shift = sp - par;  // 0, or -1 (mod 2^32) for a round staged before the tile's start was known
            // full aligned pairs: positions [i_lo, i_hi); at most one lone entry on either side of them:
            // position 1 when the run starts on an odd index, position end - 1 when it ends on an even one
            const uint32_t i_lo = 2 * par, i_hi = end & ~1u;
            uint32_t lone = 0xFFFFFFFFu;
            if (lane == 0 && par && end > 1u) lone = 1u;
            if (lane == 1 && (end & 1u) && end - 1u >= i_lo && end - 1u >= par) lone = end - 1u;
            if (p.out.all_vec && shift == 0u) {  // the common case: all three arrays asked for and aligned, slice laid out like the output
                uint64_t* gc = p.out.canon + gb;
                uint64_t* gh = p.out.hash + gb;
                int32_t* gp = p.out.pos + gb;
                write_pairs<true>(lane, i_lo, i_hi, sc, shh, spos, gc, gh, gp);
                if (lone != 0xFFFFFFFFu) {
                    const uint32_t i0 = swz(lone);
                    gc[lone] = sc[i0];
                    gh[lone] = shh[i0];
                    gp[lone] = spos[i0];
                }
            } else {
                uint64_t* gc = p.out.canon ? p.out.canon + gb : nullptr;
                uint64_t* gh = p.out.hash ? p.out.hash + gb : nullptr;
                int32_t* gp = p.out.pos ? p.out.pos + gb : nullptr;
                if (p.out.vec16 && shift == 0u) {
                    write_pairs<false>(lane, i_lo, i_hi, sc, shh, spos, gc, gh, gp);
                } else if (p.out.vec16) {  // the output arrays are 16-byte (positions: 8-byte) aligned
                    for (uint32_t i = i_lo + 2 * lane; i < i_hi; i += 64) {
                        const uint32_t i0 = swz(i + shift), i1 = swz(i + 1 + shift);
                        if (gc) *reinterpret_cast<ulonglong2*>(gc + i) = make_ulonglong2(sc[i0], sc[i1]);
                        if (gh) *reinterpret_cast<ulonglong2*>(gh + i) = make_ulonglong2(shh[i0], shh[i1]);
                        if (gp) *reinterpret_cast<int2*>(gp + i) = make_int2(spos[i0], spos[i1]);
                    }
                } else {
                    for (uint32_t i = i_lo + lane; i < i_hi; i += 32) {
                        const uint32_t i0 = swz(i + shift);
                        if (gc) gc[i] = sc[i0];
                        if (gh) gh[i] = shh[i0];
                        if (gp) gp[i] = spos[i0];
                    }
                }
                if (lone != 0xFFFFFFFFu) {
                    const uint32_t i0 = swz(lone + shift);
                    if (gc) gc[lone] = sc[i0];
                    if (gh) gh[lone] = shh[i0];
                    if (gp) gp[lone] = spos[i0];
                }
            }
        }
        __syncwarp();  // the next round re-uses this warp's slice
        if (PIPE && q_round == lb_round && threadIdx.x < 32u) {
            // the tile this CTA counted before the current emit phase: every earlier tile was counted about then and has
            // published long since; placing it now (not at the start of its own emit phase) keeps the look-back of the
            // other CTAs short -- only the tiles counted within the last few microseconds hold no prefix
            CompactPipeShared& ps = reinterpret_cast<CompactPipeShared&>(sh);
            place_tile(ps.next_tile, ps.next_total, &ps.base[ps.next_b]);
        }
    }
    __device__ __forceinline__ void finish() {
        if (COUNT_ONLY && threadIdx.x == 0 && pass_base) atomicAdd(p.out.total, (unsigned long long)pass_base);
    }
};

// dynamic shared memory: [tile (+ CSR tables)] then CompactShared
template <class Eng>
__global__ void __launch_bounds__(kExtractThreads, 4) compact_fixed_kernel(const FixedGeom g, const EncDesc enc, const CompactParams ep,
                                                                         uint32_t tile_bytes) {
    extern __shared__ uint2 tile[];
    CompactShared& sh = *reinterpret_cast<CompactShared*>(reinterpret_cast<unsigned char*>(tile) + tile_bytes);
    Eng eng(ep, sh);
    fixed_body(g, enc, eng, tile, eng.take_tile());
    eng.finish();
}

// Fixed-length reads, emit launch: persistent CTAs, software-pipelined over the tiles (see the head of this file).
// Dynamic shared memory: [tile 0][tile 1] (tile_bytes each) then CompactPipeShared.  n_tiles descriptors.
template <class Eng>
__global__ void __launch_bounds__(kExtractThreads, KMB_COMPACT_MINCTAS) compact_fixed_pipe_kernel(const FixedGeom g, const EncDesc enc, const CompactParams ep,
                                                                               uint32_t tile_bytes, uint32_t n_tiles) {
    extern __shared__ uint2 tile[];
    CompactPipeShared& ps = *reinterpret_cast<CompactPipeShared*>(reinterpret_cast<unsigned char*>(tile) + 2 * (size_t)tile_bytes);
    Eng eng(ep, ps.s);
    eng.n_tiles = n_tiles;
    const uint32_t K = eng.K();
    using Shape = typename Eng::Shape;
    const uint32_t tile_entries = tile_bytes / (uint32_t)sizeof(uint2);

    // stage + count tile t into buffer b, publish its aggregate; returns its count.  Starts behind a barrier (take_tile).
    auto count_ahead = [&](uint32_t t, uint32_t b) {
        uint2* tl = tile + b * tile_entries;
        const FixedTile ft = fixed_stage<Eng>(g, enc, K, tl, t);
        if (threadIdx.x == 0) ps.ft[b] = ft;
        eng.bind(b ? &ps.buf1 : &ps.s.buf, t);
        auto item = [&](uint32_t li, auto&& one, auto&& two, auto&& single) { fixed_item<Shape, true>(g, ft, li, one, two, single); };
        count_pass(eng, tl, K, ft.n_items, item);  // (its first barrier is the one behind the staging)
        eng.publish_count();
        return (uint32_t)eng.pass_base;
    };

    uint32_t cur = eng.take_tile();
    if (cur >= n_tiles) return;
    uint32_t b = 0;
    uint32_t cur_total = count_ahead(cur, 0);
    if (threadIdx.x < 32u) eng.place_tile(cur, cur_total, &ps.base[0]);  // the CTA's first tile: nobody has looked back for it
    for (;;) {
        // The ticket is taken right before the tile is counted, never earlier: a CTA that holds an uncounted ticket through an
        // emit phase (to hide the atomic's round trip, or to prefetch the tile) makes every later tile's look-back wait for that
        // phase, the waits add up from CTA to CTA and the whole grid falls into lock-step -- measured 2.1x slower; taking it
        // only one write-out early was no gain either (kernels_r02i).
        const uint32_t nxt = eng.take_tile();  // barrier: the previous emit phase is over, ps.base[] is visible
        const bool has_next = nxt < n_tiles;
        uint32_t nxt_total = 0;
        if (has_next) {
            nxt_total = count_ahead(nxt, b ^ 1u);
            if (threadIdx.x == 0) { ps.next_tile = nxt; ps.next_total = nxt_total; ps.next_b = b ^ 1u; }
            __syncwarp();
        }
        {
            const FixedTile ft = ps.ft[b];
            const uint32_t rounds = (ft.n_items + kExtractThreads - 1) / kExtractThreads;
            eng.bind(b ? &ps.buf1 : &ps.s.buf, cur);
            eng.begin_emit(cur_total, ps.base[b], has_next ? min(1u, rounds - 1u) : 0xFFFFFFFFu);
            auto item = [&](uint32_t li, auto&& one, auto&& two, auto&& single) { fixed_item<Shape, true>(g, ft, li, one, two, single); };
            emit_pass(eng, tile + b * tile_entries, ft.n_items, item, [](uint32_t) {});
        }
        if (!has_next) break;
        cur = nxt; cur_total = nxt_total; b ^= 1u;
    }
}

template <class Eng>
__global__ void __launch_bounds__(kExtractThreads, 3) compact_csr_kernel(const CsrGeom g, const EncDesc enc, const CompactParams ep,
                                                                       uint32_t tile_bytes) {
    extern __shared__ uint2 tile[];
    __shared__ CsrPass pass;
    uint64_t* c_off = reinterpret_cast<uint64_t*>(tile + g.tile_entries);
    uint64_t* c_win = c_off + (kCsrCache + 2);
    typename Eng::Shared& sh = *reinterpret_cast<typename Eng::Shared*>(reinterpret_cast<unsigned char*>(tile) + tile_bytes);
    Eng eng(ep, sh);
    csr_body(g, enc, eng, tile, c_off, c_win, &pass, eng.take_tile());
    eng.finish();
}

}  // namespace kmb
