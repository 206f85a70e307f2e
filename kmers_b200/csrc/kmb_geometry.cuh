// kmb_geometry.cuh -- how a batch of reads is cut into CTA tiles and per-thread work items, shared by every
// extraction-type engine: K <= 32 (kmb_extract.cuh), two-word (kmb_extract_wide.cuh), compaction (kmb_compact.cuh),
// minimizers (kmb_minimizer.cuh).
//
// Work is cut in OUTPUT-slot space: an item is kRun = 8 windows of the dense result arrays (SURVEY.md 8d layout) whose
// slots are fixed by the item's index alone (ShapeRun: 8 consecutive slots; ShapePair: two lanes interleaved over 16
// slots), so an item's stores are aligned whatever the read lengths are.  A slot maps back to (read, window position):
//   fixed-length reads : read = slot / W, pos = slot % W          (W = L - K + 1)
//   ragged (CSR) reads : read = last r with win_offsets[r] <= slot, pos = slot - win_offsets[r]
// An item lies inside one read (one span), straddles one read boundary (two spans), or -- only for reads with fewer
// windows than an item spans -- covers several reads (window-by-window path).
//
// A CTA first stages the stretch of the flat read stream its windows cover into shared memory (2 bits/base + 1 invalid
// bit/base, kmb_device.cuh), then runs its items from that tile: the common case (one clean span) in a first sweep, the
// few two-span / dirty items densely in a second one (run_pass).
#pragma once
#ifndef KMB_CSR_NARROW
#define KMB_CSR_NARROW 1
#endif
#include "kmb_device.cuh"

namespace kmb {

#ifndef KMB_EXTRACT_THREADS
#define KMB_EXTRACT_THREADS 256
#endif
#ifndef KMB_ITEMS_PER_CTA
#define KMB_ITEMS_PER_CTA 1024
#endif
constexpr int kExtractThreads = KMB_EXTRACT_THREADS;
constexpr int kItemsPerCta = KMB_ITEMS_PER_CTA;  // default 1024 items = 8192 slots per CTA
constexpr int kMaxItemsPerCta = 4 * kItemsPerCta;  // the 1024-thread histogram CTAs take 4096 items per tile (fixed-length reads)
constexpr int kStageBatch = 3;                   // 16-byte loads a thread keeps in flight while staging

// ---------------------------------------------------------------------------
// phase 1: stage a stretch of the read stream into shared memory as
// {packed bits, invalid mask} entries, 16 bases each.  Entry 0 starts at
// first_al (16-byte aligned, at or below the first base needed).
// ---------------------------------------------------------------------------
template <bool VALIDATE, int BATCH = kStageBatch>
__device__ __forceinline__ void stage_tile(const uint8_t* bases, uint64_t n_bytes, const uint8_t* first_al,
                                           uint32_t n_entries, const EncDesc& enc, uint2* tile) {
    // CTA-uniform: does the whole stretch lie inside the batch?  (all but the edge CTAs)
    const bool inside = first_al >= bases && first_al + (size_t)n_entries * 16 <= bases + n_bytes;
    if (inside) {
        const uint4* src = reinterpret_cast<const uint4*>(first_al);
        for (uint32_t v0 = threadIdx.x; v0 < n_entries; v0 += BATCH * blockDim.x) {
            uint4 raw[BATCH];
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {  // all loads first: BATCH requests in flight per thread
                const uint32_t v = v0 + b * blockDim.x;
                if (v < n_entries) raw[b] = ld_stream_v4(src + v);
            }
#pragma unroll
            for (int b = 0; b < BATCH; ++b) {
                const uint32_t v = v0 + b * blockDim.x;
                if (v < n_entries) {
                    PackedWord pw = pack16<VALIDATE>(raw[b]);
                    tile[v] = make_uint2(apply_encoding(pw.bits, enc), pw.inv);
                }
            }
        }
    } else {
#pragma unroll 1
        for (uint32_t v = threadIdx.x; v < n_entries; v += blockDim.x) {
            PackedWord pw = pack16<VALIDATE>(load16_guarded(bases, n_bytes, first_al + (size_t)v * 16));
            tile[v] = make_uint2(apply_encoding(pw.bits, enc), pw.inv);
        }
    }
}

// Where an item (or a single window) sits in the batch: handed to the engine with every call so that
// engines producing per-read or position-bearing output (kmb_compact.cuh) know what they are emitting.
struct ItemCtx {
    uint32_t li;     // item index inside the CTA's current pass
    uint64_t r_a;    // read of span A's windows
    uint64_t pos_a;  // position inside read r_a of span A's window 0
    uint64_t r_b;    // read of span B's windows (two-span items; they start at position 0 of r_b)
};

// Which slots a work item covers, counted from its first slot.  An item always has kRun windows; `nwin` handed to
// the engine is the number of SLOTS that exist from the item's first slot on (capped at kSpanSlots), so window j is
// present iff off(j) < nwin, and comes from span B of a two-span item iff off(j) >= n_first.
//   ShapeRun : kRun consecutive slots (64 B of an 8-byte array: one lane's two 32-byte stores fill half a line).
//   ShapePair: for 16-byte slots.  kRun consecutive slots would be 128 B per lane, and a warp-wide 32-byte store would
//              touch 32 different lines -- measured 5.1 TB/s instead of 6.1 (scripts/micro/store_patterns.cu; ncu: the
//              L1 LSU data pipe is 93 % busy).  So lanes 2i and 2i+1 share 16 slots: lane 2i takes slots
//              {0,1, 4,5, 8,9, 12,13}, lane 2i+1 the same + 2, and every store instruction covers 64 contiguous bytes per
//              lane pair.  Still one span and kRun windows per lane.
struct ShapeRun {
    static constexpr int kSpanSlots = kRun;  // slots from the first to the last window, inclusive
    static constexpr int kAlign = kRun;      // a pass must start at a multiple of this many slots
    __device__ static __forceinline__ uint32_t first(uint32_t li) { return li * kRun; }
    __device__ static __forceinline__ constexpr int off(int j) { return j; }
    __device__ static __forceinline__ bool owns(uint32_t) { return true; }
    __device__ static __forceinline__ uint32_t n_items(uint32_t n_slots) { return (n_slots + kRun - 1) / kRun; }
};
struct ShapePair {
    static constexpr int kSpanSlots = 2 * kRun - 2;
    static constexpr int kAlign = 2 * kRun;
    __device__ static __forceinline__ uint32_t first(uint32_t li) { return (li >> 1) * (2 * kRun) + (li & 1u) * 2u; }
    __device__ static __forceinline__ constexpr int off(int j) { return (j >> 1) * 4 + (j & 1); }
    __device__ static __forceinline__ bool owns(uint32_t s) { return (s & 2u) == 0u; }
    __device__ static __forceinline__ uint32_t n_items(uint32_t n_slots) { return ((n_slots + 2 * kRun - 1) / (2 * kRun)) * 2; }
};

// Number of windows j in [j0, j1) starting at tile base rel + j whose K bases are all valid
// (K <= 64, j1 <= kRun).  NE = tile entries that may be read (the engine's span size).
template <int NE>
__device__ __forceinline__ uint32_t count_valid_windows(const uint2* tile, uint32_t rel, uint32_t K, uint32_t j0, uint32_t j1) {
    const uint32_t e = rel >> 4, o = rel & 15u;
    uint32_t y[6] = {0, 0, 0, 0, 0, 0};
    uint32_t any = 0;
#pragma unroll
    for (int i = 0; i < NE && i < 6; ++i) { y[i] = tile[e + i].y; any |= y[i]; }
    if (any == 0u) return j1 - j0;
    const uint64_t m0 = (uint64_t)y[0] | ((uint64_t)y[1] << 16) | ((uint64_t)y[2] << 32) | ((uint64_t)y[3] << 48);
    const uint64_t m1 = (uint64_t)y[4] | ((uint64_t)y[5] << 16);
    const uint64_t lo = o ? ((m0 >> o) | (m1 << (64 - o))) : m0;
    const uint64_t hi = m1 >> o;
    const uint64_t kmask = K >= 64 ? ~0ull : ((1ull << K) - 1ull);
    uint32_t n = 0;
    for (uint32_t j = j0; j < j1; ++j) {
        const uint64_t x = j ? ((lo >> j) | (hi << (64 - j))) : lo;
        n += (x & kmask) == 0ull;
    }
    return n;
}

// 2-bit packed input (the layout kmb_pack writes with Naive::ACGT and u64 words = SeqVector's words,
// naive_impl/seq_vector.rs:230-242): the tile entries are the packed 32-bit words themselves.
// A packed store cannot hold an invalid base, so the invalid masks are zero.
// With `inv` (the flat packed staging format the host packer writes, kmb_hostpack.h): one 16-bit invalid mask per
// word travels with the bits, so validation works as on ASCII input; words past the end read as invalid.
__device__ __forceinline__ void stage_packed(const uint32_t* words, const uint16_t* inv, uint64_t n_words32, uint64_t first_word,
                                             uint32_t n_entries, const EncDesc& enc, uint2* tile) {
    for (uint32_t v = threadIdx.x; v < n_entries; v += blockDim.x) {
        const uint64_t i = first_word + v;
        uint32_t w = i < n_words32 ? __ldg(words + i) : 0u;
        // the store holds A0 C1 G2 T3; x ^ (x >> 1) per field is its own inverse and leads back to the internal code
        if (!enc.is_acgt) w = apply_encoding(w ^ ((w >> 1) & 0x55555555u), enc);
        uint32_t m = 0u;
        if (inv) m = i < n_words32 ? (uint32_t)__ldg(inv + i) : 0xFFFFu;
        tile[v] = make_uint2(w, m);
    }
}

// Stage the stretch that starts at flat base index g_start; returns the offset of that base inside tile entry 0.
template <bool VALIDATE>
__device__ __forceinline__ uint32_t stage_stretch(const uint8_t* bases, uint64_t n_bytes, uint32_t packed, uint64_t g_start,
                                                  uint32_t span, uint32_t span_entries, const EncDesc& enc, uint2* tile,
                                                  const uint16_t* inv = nullptr) {
    if (packed) {
        const uint32_t mis = (uint32_t)(g_start & 15u);
        stage_packed(reinterpret_cast<const uint32_t*>(bases), inv, n_bytes >> 2, g_start >> 4, ((span + mis + 15) >> 4) + span_entries - 1, enc, tile);
        return mis;
    }
    const uint8_t* first = bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    stage_tile<VALIDATE>(bases, n_bytes, first - mis, ((span + mis + 15) >> 4) + span_entries - 1, enc, tile);
    return mis;
}

// ---------------------------------------------------------------------------
// fixed-length reads
// ---------------------------------------------------------------------------
struct FixedGeom {
    const uint8_t* bases;   // flat read stream
    uint64_t n_bytes;       // bytes of the buffer behind `bases`
    uint64_t L;             // distance between read starts, in bases (= read length; packed: padded to 32)
    uint64_t W;             // windows (= output slots) per read = L - K + 1
    uint64_t total_slots;   // n_reads * W
    uint64_t w_magic64;     // floor(2^64 / W) + 1 (W >= 2), 0 = divide
    uint32_t L32;           // L mod 2^32 (only differences inside a tile are formed)
    uint32_t W32;           // W (< 2^32, checked on the host)
    uint32_t w_magic;       // floor(2^32 / W) + 1
    uint32_t items_per_cta; // host-chosen so the staged stretch fits shared memory
    uint32_t packed;        // 0: ASCII; 1: 2-bit packed words, SeqVector layout; 2: flat 2-bit stream + invalid masks (`inv`)
    const uint16_t* inv;    // packed == 2: one 16-bit invalid mask per 32-bit word of `bases`
    uint32_t unified;       // W >= kRun and not a multiple of it: every read has a straddling item (see fixed_item)
};

// u / W for a small u (u < W + slots per CTA): 0/1 when W is large, else multiply-high
__device__ __forceinline__ uint32_t div_w(uint32_t u, const FixedGeom& g, uint32_t slots_per_cta) {
    if (g.W32 >= slots_per_cta) return (u >= g.W32) ? 1u : 0u;  // u < W + slots_per_cta <= 2W
    if (g.W32 == 1) return u;
    return __umulhi(u, g.w_magic);
}

// The item loop.  Plain engines stride over the items.  Two-phase engines need every thread of the CTA to reach
// round_end (it holds barriers) the same number of times and from ONE call site (`__syncthreads` is an aligned
// barrier: all lanes of a warp must execute the same instruction), so they run CTA-uniform rounds.
template <bool ROUNDS, class Item, class RoundBegin, class RoundEnd>
__device__ __forceinline__ void for_each_item(uint32_t n_items, Item&& item, RoundBegin&& round_begin, RoundEnd&& round_end) {
    if constexpr (ROUNDS) {
        const uint32_t rounds = (n_items + blockDim.x - 1) / blockDim.x;
        for (uint32_t q_round = 0; q_round < rounds; ++q_round) {
            const uint32_t li = q_round * blockDim.x + threadIdx.x;
            round_begin(q_round);
            if (li < n_items) item(li);
            __syncwarp();
            round_end(q_round);
        }
    } else {
        for (uint32_t li = threadIdx.x; li < n_items; li += blockDim.x) item(li);
    }
}

// Items off the common path -- two-span items, spans that hold an invalid base -- are few, but a warp that meets one
// executes that path for a handful of lanes at the full instruction cost of the common path (measured: +60..100 %
// ALU work whenever W is not a multiple of the item size).  Plain engines therefore only note such items during the
// first sweep and run them afterwards, densely packed: 32 of them per warp instruction.
struct Deferred {
    uint32_t n;
    uint16_t li[kMaxItemsPerCta];
};
static_assert(kMaxItemsPerCta <= 65536, "deferred item indices are 16-bit");
__device__ __forceinline__ Deferred& deferred() {
    __shared__ Deferred d;
    return d;
}
// called by the bodies ahead of the barrier that ends staging
__device__ __forceinline__ void deferred_reset() {
    if (threadIdx.x == 0) deferred().n = 0;
}
__device__ __forceinline__ void defer(uint32_t li) {
    Deferred& d = deferred();
    d.li[atomicAdd(&d.n, 1u)] = (uint16_t)li;
}

// Two-phase engines (compaction), first half of a pass: count the valid windows of every item of the staged tile and scan the
// counts CTA-wide.  Ends behind a barrier.
template <class Eng, class Item>
__device__ __forceinline__ void count_pass(Eng& eng, const uint2* tile, uint32_t K, uint32_t n_items, Item&& item) {
    eng.begin_pass(n_items);
    __syncthreads();
    auto count_one = [&](uint32_t rel, uint64_t, uint32_t nwin, const ItemCtx& ic) {
        eng.count(ic.li, count_valid_windows<Eng::kSpanEntries>(tile, rel, K, 0, nwin));
    };
    auto count_two = [&](uint32_t rel_a, uint32_t rel_b, uint32_t left, uint64_t, uint32_t nwin, const ItemCtx& ic) {
        uint32_t c = count_valid_windows<Eng::kSpanEntries>(tile, rel_a, K, 0, left);
        if (left < nwin) c += count_valid_windows<Eng::kSpanEntries>(tile, rel_b, K, left, nwin);
        eng.count(ic.li, c);
    };
    auto count_single = [&](uint32_t rel, uint64_t, const ItemCtx& ic) {
        eng.count(ic.li, count_valid_windows<Eng::kSpanEntries>(tile, rel, K, 0, 1));
    };
    for_each_item<true>(n_items, [&](uint32_t li) { item(li, count_one, count_two, count_single); }, [](uint32_t) {}, [](uint32_t) {});
    __syncthreads();
    eng.scan(n_items);
    __syncthreads();
}

// ... second half: emit at the scanned offsets, in CTA-uniform rounds.  Two-phase engines cannot put their dirty items off to
// a second sweep (their output order is fixed by the scan), so they run every item through the checking variant: a few
// instructions per window instead of both variants per warp.
template <class Eng, class Item, class AfterRound>
__device__ __forceinline__ void emit_pass(Eng& eng, const uint2* tile, uint32_t n_items, Item&& item, AfterRound&& after_round) {
    constexpr uint32_t kAll = (uint32_t)Eng::Shape::kSpanSlots;
    auto emit_one = [&](uint32_t rel, uint64_t slot0, uint32_t nwin, const ItemCtx& ic) {
        const typename Eng::Span s = eng.load(tile, rel);
        if (Eng::kValidate) eng.template run<false, true>(s, s, kAll, slot0, nwin, ic);
        else eng.template run<false, false>(s, s, kAll, slot0, nwin, ic);
    };
    auto emit_two = [&](uint32_t rel_a, uint32_t rel_b, uint32_t left, uint64_t slot0, uint32_t nwin, const ItemCtx& ic) {
        const typename Eng::Span a = eng.load(tile, rel_a);
        typename Eng::Span b = a;
        if (rel_b != rel_a) b = eng.load(tile, rel_b);  // (unified items: a one-span item comes with rel_b == rel_a)
        if (Eng::kValidate) eng.template run<true, true>(a, b, left, slot0, nwin, ic);
        else eng.template run<true, false>(a, b, left, slot0, nwin, ic);
    };
    auto emit_single = [&](uint32_t rel, uint64_t slot, const ItemCtx& ic) { eng.single(tile, rel, slot, ic); };
    for_each_item<true>(n_items, [&](uint32_t li) { item(li, emit_one, emit_two, emit_single); },
                        [&](uint32_t q_round) { eng.round_begin(q_round); },
                        [&](uint32_t q_round) { eng.round_end(q_round, n_items); after_round(q_round); });
}

// One pass over the items of a staged tile.  `item(li, one, two, single)` classifies item li and calls exactly one of
// the three handlers (single: once per window).  Two-phase engines (compaction) first count the valid windows of
// every item, scan the counts CTA-wide, then emit at the scanned offsets.
template <class Eng, class Item>
__device__ __forceinline__ void run_pass(Eng& eng, const uint2* tile, uint32_t K, uint32_t n_items, Item&& item, bool last_pass = true) {
    if constexpr (Eng::kTwoPhase) {
        count_pass(eng, tile, K, n_items, item);
        if (Eng::kCountOnly) return;
        eng.publish(last_pass);  // the tile's count goes out at once
        emit_pass(eng, tile, n_items, item, [](uint32_t) {});
    } else {
        constexpr uint32_t kAll = (uint32_t)Eng::Shape::kSpanSlots;
        auto emit_one = [&](uint32_t rel, uint64_t slot0, uint32_t nwin, const ItemCtx& ic) {
            const typename Eng::Span s = eng.load(tile, rel);
            if (Eng::kValidate && eng.dirty(s)) eng.template run<false, true>(s, s, kAll, slot0, nwin, ic);
            else eng.template run<false, false>(s, s, kAll, slot0, nwin, ic);
        };
        auto emit_two = [&](uint32_t rel_a, uint32_t rel_b, uint32_t left, uint64_t slot0, uint32_t nwin, const ItemCtx& ic) {
            const typename Eng::Span a = eng.load(tile, rel_a);
            const typename Eng::Span b = eng.load(tile, rel_b);
            if (Eng::kValidate && (eng.dirty(a) || eng.dirty(b))) eng.template run<true, true>(a, b, left, slot0, nwin, ic);
            else eng.template run<true, false>(a, b, left, slot0, nwin, ic);
        };
        auto emit_single = [&](uint32_t rel, uint64_t slot, const ItemCtx& ic) { eng.single(tile, rel, slot, ic); };
        // sweep 1: the common path (one clean span); everything else is noted
        auto fast_one = [&](uint32_t rel, uint64_t slot0, uint32_t nwin, const ItemCtx& ic) {
            const typename Eng::Span s = eng.load(tile, rel);
            if (Eng::kValidate && eng.dirty(s)) defer(ic.li);
            else eng.template run<false, false>(s, s, kAll, slot0, nwin, ic);
        };
        auto later_two = [&](uint32_t, uint32_t, uint32_t, uint64_t, uint32_t, const ItemCtx& ic) { defer(ic.li); };
        for_each_item<false>(n_items, [&](uint32_t li) { item(li, fast_one, later_two, emit_single); }, [](uint32_t) {}, [](uint32_t) {});
        __syncthreads();
        // sweep 2: the noted items, one per lane
        const Deferred& d = deferred();
        const uint32_t n_def = d.n;
        for (uint32_t i = threadIdx.x; i < n_def; i += blockDim.x) item(d.li[i], emit_one, emit_two, emit_single);
    }
}

// What a CTA knows about one tile of the fixed geometry (CTA-uniform)
struct FixedTile {
    uint64_t slot_base;  // first output slot
    uint64_t r_first;    // read of that slot
    uint32_t n_slots;
    uint32_t p_first;    // position of the tile's first window in read r_first
    uint32_t mis;        // offset of the tile's first base inside tile entry 0
    uint32_t n_items;
};

// the stretch of the flat stream that holds tile tile_idx's windows: first base and length in bases; fills t except mis / n_items
__device__ __forceinline__ void fixed_tile_range(const FixedGeom& g, uint32_t K, uint32_t tile_idx, FixedTile& t, uint64_t& g_start, uint32_t& span) {
    const uint32_t slots_per_cta = g.items_per_cta * kRun;
    t.slot_base = (uint64_t)tile_idx * slots_per_cta;
    t.n_slots = (uint32_t)min((uint64_t)slots_per_cta, g.total_slots - t.slot_base);
    if (g.W == 1) t.r_first = t.slot_base;
    else if (g.w_magic64) t.r_first = div_magic64(t.slot_base, g.w_magic64);
    else t.r_first = t.slot_base / g.W;
    t.p_first = (uint32_t)(t.slot_base - t.r_first * g.W);
    g_start = t.r_first * g.L + t.p_first;
    const uint32_t u_last = t.p_first + t.n_slots - 1;
    const uint32_t q_last = div_w(u_last, g, slots_per_cta);
    // bases from the first window's first base to the last window's last base (mod 2^32 exact: small)
    span = q_last * g.L32 + (u_last - q_last * g.W32) - t.p_first + K;
}

// phase 1: pack the stretch of the flat stream that holds the tile's windows (no barrier)
template <class Eng>
__device__ __forceinline__ FixedTile fixed_stage(const FixedGeom& g, const EncDesc& enc, uint32_t K, uint2* tile, uint32_t tile_idx) {
    FixedTile t;
    uint64_t g_start;
    uint32_t span;
    fixed_tile_range(g, K, tile_idx, t, g_start, span);
    t.mis = stage_stretch<Eng::kValidate>(g.bases, g.n_bytes, g.packed, g_start, span, Eng::kSpanEntries, enc, tile, g.inv);
    t.n_items = Eng::Shape::n_items(t.n_slots);
    return t;
}

// phase 2: every item is one span, two spans (straddles a read boundary) or, for reads with fewer than kRun windows, a run
// of single windows
template <class Shape, bool UNIFIED_OK, class One, class Two, class Single>
__device__ __forceinline__ void fixed_item(const FixedGeom& g, const FixedTile& t, uint32_t li, One&& one, Two&& two, Single&& single) {
    const uint32_t slots_per_cta = g.items_per_cta * kRun;
    const uint32_t fs = Shape::first(li);              // the item's first slot, counted from the CTA's
    if (fs >= t.n_slots) return;
    const uint32_t u = t.p_first + fs;                 // ... and from window 0 of read r_first
    const uint32_t q = div_w(u, g, slots_per_cta);     // reads crossed since r_first
    const uint32_t pos = u - q * g.W32;                 // window position inside its read
    const uint64_t slot0 = t.slot_base + fs;
    const uint32_t nwin = min((uint32_t)Shape::kSpanSlots, t.n_slots - fs);
    const uint32_t rel = q * g.L32 + pos - t.p_first + t.mis;  // first base, relative to tile entry 0
    const uint32_t left = g.W32 - pos;                   // windows left in this read (>= 1)
    const ItemCtx ic{li, t.r_first + q, pos, t.r_first + q + 1};
    if (UNIFIED_OK && g.unified) {
        // two-phase engines when every read has a straddling item: one call site for both kinds of item (as csr_item32)
        const bool straddles = left < nwin;
        two(rel, straddles ? (q + 1) * g.L32 - t.p_first + t.mis - left : rel, straddles ? left : nwin, slot0, nwin, ic);
        return;
    }
    if (left >= (uint32_t)Shape::kSpanSlots || left >= nwin) {
        one(rel, slot0, nwin, ic);
    } else if (g.W32 >= (uint32_t)Shape::kSpanSlots) {
        // straddles exactly one boundary: slots s >= left start read q+1 at position s - left
        two(rel, (q + 1) * g.L32 - t.p_first + t.mis - left, left, slot0, nwin, ic);
    } else {
        for (uint32_t s = 0; s < nwin; ++s) {
            if (!Shape::owns(s)) continue;
            const uint32_t us = u + s, qs = div_w(us, g, slots_per_cta), ps = us - qs * g.W32;
            single(qs * g.L32 + ps - t.p_first + t.mis, slot0 + s, ItemCtx{li, t.r_first + qs, ps, 0});
        }
    }
}

template <class Eng>
__device__ __forceinline__ void fixed_body(const FixedGeom& g, const EncDesc& enc, Eng& eng, uint2* tile, uint32_t tile_idx) {
    const uint32_t K = eng.K();
    const FixedTile t = fixed_stage<Eng>(g, enc, K, tile, tile_idx);
    if constexpr (!Eng::kTwoPhase) deferred_reset();
    __syncthreads();
    using Shape = typename Eng::Shape;
    auto item = [&](uint32_t li, auto&& one, auto&& two, auto&& single) { fixed_item<Shape, Eng::kTwoPhase>(g, t, li, one, two, single); };
    run_pass(eng, tile, K, t.n_items, item);
}

// ---------------------------------------------------------------------------
// ragged reads (CSR offsets)
// ---------------------------------------------------------------------------
constexpr int kCsrCache = 512;   // reads whose offsets a CTA keeps in shared memory (more: read from global)
constexpr int kCsrGroup = 8;     // items per entry of the per-pass owner table

// What a CTA of the ragged kernels needs to know before it can fetch anything, precomputed per tile by csr_index_kernel
// for the common case that the tile's whole stretch fits one staged pass: the CTA then starts staging at once, with no
// serial set-up by thread 0 and one barrier less.
struct alignas(16) CsrTileDesc {
    uint64_t g0;      // flat base index of the first base the tile needs
    uint64_t r_lo;    // read owning the tile's first slot
    uint32_t d_last;  // read owning the tile's last slot, counted from r_lo
    uint32_t d_hi;    // read owning the NEXT tile's first slot (the last slot's, for the last tile), counted from r_lo
    uint32_t span;    // bases from g0 to the last window's last base; 0xFFFFFFFF: more than one pass (set up in the kernel)
    uint32_t narrow;  // bit 31: the tile's reads span < 2^31 bases and windows, so everything inside the tile is 32-bit arithmetic
                      // relative to read r_lo; bits 30..0 then hold the position of the tile's first slot inside read r_lo
};
constexpr uint32_t kCsrNarrow = 0x80000000u;

struct CsrGeom {
    const uint8_t* bases;
    uint64_t n_bytes;             // bytes of the buffer behind `bases`
    uint64_t n_bases;             // length of the flat base index space (ASCII: n_bytes; packed: 4 * n_bytes)
    const uint64_t* offsets;      // n_reads + 1: flat base index of every read's first base
    const uint64_t* win_offsets;  // n_reads + 1, exclusive prefix of per-read window counts
    const CsrTileDesc* tile_desc; // grid entries (csr_index_kernel): everything a CTA needs before its first fetch, in ONE load
    uint64_t n_reads;
    uint64_t total_slots;
    uint32_t items_per_cta;
    uint32_t tile_entries;        // shared-memory capacity of the staged tile, in 16-base entries
    uint32_t packed;              // as FixedGeom::packed (offsets then index the padded, packed space)
};

// largest r in [lo, hi] with a[r] <= v   (a[lo] <= v guaranteed)
__device__ __forceinline__ uint64_t last_le(const uint64_t* a, uint64_t lo, uint64_t hi, uint64_t v) {
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo + 1) >> 1);
        if (a[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ... the same on 32-bit tables
__device__ __forceinline__ uint32_t last_le32(const uint32_t* a, uint32_t lo, uint32_t hi, uint32_t v) {
    while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        if (a[mid] <= v) lo = mid; else hi = mid - 1;
    }
    return lo;
}

struct CsrPass {  // one staged stretch: slots [slot_lo, slot_hi) of reads [r_lo, r_hi]
    uint64_t slot_lo, slot_hi, r_lo, r_hi, g0;
    uint32_t span;
};

// One work item of a ragged pass: which read(s) its slots belong to, then one span / two spans / window by window.
template <class Shape, bool UNIFIED, class One, class Two, class Single>
__device__ __forceinline__ void csr_item(uint32_t li, uint32_t n_slots, const CsrPass& ps, const uint64_t* off, const uint64_t* win,
                                         const uint64_t* grp, uint32_t mis, One&& one, Two&& two, Single&& single) {
    const uint32_t fs = Shape::first(li);
    if (fs >= n_slots) return;
    const uint64_t slot0 = ps.slot_lo + fs;
    const uint32_t nwin = min((uint32_t)Shape::kSpanSlots, n_slots - fs);
    uint64_t r = last_le(win, grp[li / kCsrGroup], grp[li / kCsrGroup + 1], slot0);
    const uint64_t pos = slot0 - win[r];
    const uint64_t left = win[r + 1] - win[r] - pos;  // windows left in read r (>= 1)
    const uint32_t rel = (uint32_t)(off[r] + pos - ps.g0) + mis;
    if constexpr (UNIFIED) {
        // one call site for one-span and two-span items (see csr_item32)
        uint64_t r2 = r;
        uint32_t rel_b = rel, n_first = nwin;
        bool fits = true;
        if (left < nwin) {
            r2 = r + 1;
            while (win[r2 + 1] == win[r2]) ++r2;
            fits = left + (win[r2 + 1] - win[r2]) >= nwin;
            rel_b = (uint32_t)(off[r2] - ps.g0) + mis - (uint32_t)left;
            n_first = (uint32_t)left;
        }
        if (fits) {
            two(rel, rel_b, n_first, slot0, nwin, ItemCtx{li, r, pos, r2});
            return;
        }
    } else {
        if (left >= nwin) {
            one(rel, slot0, nwin, ItemCtx{li, r, pos, 0});
            return;
        }
        uint64_t r2 = r + 1;
        while (win[r2 + 1] == win[r2]) ++r2;  // next read that has windows (exists: nwin > left)
        if (left + (win[r2 + 1] - win[r2]) >= nwin) {
            two(rel, (uint32_t)(off[r2] - ps.g0) + mis - (uint32_t)left, (uint32_t)left, slot0, nwin, ItemCtx{li, r, pos, r2});
            return;
        }
    }
    // several short reads inside one item: window by window
    uint64_t p = pos, w_r = win[r + 1] - win[r];
    for (uint32_t s = 0; s < nwin; ++s) {
        while (p >= w_r) { ++r; p = 0; w_r = win[r + 1] - win[r]; }
        if (Shape::owns(s)) single((uint32_t)(off[r] + p - ps.g0) + mis, slot0 + s, ItemCtx{li, r, p, 0});
        ++p;
    }
}

// The same for a "narrow" tile (CsrTileDesc::narrow): w32[i] / o32[i] = window / base offset of read r_lo + i counted from read
// r_lo's, p0 = position of the tile's first slot inside read r_lo, grp32 = owner (relative to r_lo) of every group's first
// slot.  Slots are compared in the biased coordinate fs + p0.  No 64-bit compare, subtract or shared-memory load is left on
// the per-item path (the absolute read numbers of ItemCtx are only formed by the engines that use them).
template <class Shape, bool UNIFIED, class One, class Two, class Single>
__device__ __forceinline__ void csr_item32(uint32_t li, uint32_t n_slots, uint32_t p0, uint64_t slot_begin, uint64_t r_lo, const uint32_t* w32,
                                           const uint32_t* o32, const uint32_t* grp32, uint32_t mis, One&& one, Two&& two, Single&& single) {
    const uint32_t fs = Shape::first(li);
    if (fs >= n_slots) return;
    const uint32_t fb = fs + p0;
    const uint64_t slot0 = slot_begin + fs;
    const uint32_t nwin = min((uint32_t)Shape::kSpanSlots, n_slots - fs);
    uint32_t r = last_le32(w32, grp32[li / kCsrGroup], grp32[li / kCsrGroup + 1], fb);
    const uint32_t pos = fb - w32[r];
    const uint32_t left = w32[r + 1] - w32[r] - pos;  // windows left in read r (>= 1)
    const uint32_t rel = o32[r] + pos - p0 + mis;     // the tile's stretch starts p0 bases into read r_lo
    if constexpr (UNIFIED) {
        // Engines that cannot put two-span items off to a second sweep (compaction: the scan fixes the order) would run the
        // one-span AND the two-span handler in every warp that holds a straddling item -- on ragged reads that is every warp,
        // in both sweeps.  Here both kinds go through ONE call of the two-span handler: a one-span item is a two-span item
        // whose second span is its first and never used (n_first = nwin).
        uint32_t r2 = r, rel_b = rel, n_first = nwin;
        bool fits = true;
        if (left < nwin) {
            r2 = r + 1;
            while (w32[r2 + 1] == w32[r2]) ++r2;  // next read that has windows (exists: nwin > left)
            fits = left + (w32[r2 + 1] - w32[r2]) >= nwin;
            rel_b = o32[r2] - p0 + mis - left;
            n_first = left;
        }
        if (fits) {
            two(rel, rel_b, n_first, slot0, nwin, ItemCtx{li, r_lo + r, pos, r_lo + r2});
            return;
        }
    } else {
        if (left >= nwin) {
            one(rel, slot0, nwin, ItemCtx{li, r_lo + r, pos, 0});
            return;
        }
        uint32_t r2 = r + 1;
        while (w32[r2 + 1] == w32[r2]) ++r2;  // next read that has windows (exists: nwin > left)
        if (left + (w32[r2 + 1] - w32[r2]) >= nwin) {
            two(rel, o32[r2] - p0 + mis - left, left, slot0, nwin, ItemCtx{li, r_lo + r, pos, r_lo + r2});
            return;
        }
    }
    // several short reads inside one item: window by window
    uint32_t p = pos, w_r = w32[r + 1] - w32[r];
    for (uint32_t s = 0; s < nwin; ++s) {
        while (p >= w_r) { ++r; p = 0; w_r = w32[r + 1] - w32[r]; }
        if (Shape::owns(s)) single(o32[r] + p - p0 + mis, slot0 + s, ItemCtx{li, r_lo + r, p, 0});
        ++p;
    }
}

template <class Eng>
__device__ __forceinline__ void csr_body(const CsrGeom& g, const EncDesc& enc, Eng& eng, uint2* tile, uint64_t* c_off,
                                         uint64_t* c_win, CsrPass* pass, uint32_t tile_idx) {
    using Shape = typename Eng::Shape;
    constexpr bool kUnified = Eng::kTwoPhase;  // (see csr_item32)
    const uint32_t K = eng.K();
    const uint64_t slots_per_cta = (uint64_t)g.items_per_cta * kRun;
    const uint64_t slot_begin = (uint64_t)tile_idx * slots_per_cta;
    const uint64_t slot_end = min(g.total_slots, slot_begin + slots_per_cta);
    const uint32_t tile_bases = (g.tile_entries - Eng::kSpanEntries - 1) * 16;  // bases one pass can stage
    __shared__ uint64_t grp[kMaxItemsPerCta / kCsrGroup + 1];

    // reads this CTA can touch; their offsets go to shared memory when they fit (the common case)
    const uint4 td_a = __ldg(reinterpret_cast<const uint4*>(g.tile_desc + tile_idx));
    const uint4 td_b = __ldg(reinterpret_cast<const uint4*>(g.tile_desc + tile_idx) + 1);
    CsrTileDesc td;
    td.g0 = mk64(td_a.x, td_a.y); td.r_lo = mk64(td_a.z, td_a.w); td.d_last = td_b.x; td.d_hi = td_b.y; td.span = td_b.z;
    td.narrow = td_b.w;
    const uint64_t R_lo = td.r_lo, R_hi = td.r_lo + td.d_hi;
    if ((td.narrow & kCsrNarrow) && KMB_CSR_NARROW) {
        // The common case in 32 bits: one pass, the tile's reads cached as offsets relative to read R_lo.
        const uint32_t p0 = td.narrow & ~kCsrNarrow;
        const uint64_t win_lo = slot_begin - p0, off_lo = td.g0 - p0;  // = win_offsets[R_lo], offsets[R_lo]
        uint32_t* w32 = reinterpret_cast<uint32_t*>(c_win);
        uint32_t* o32 = reinterpret_cast<uint32_t*>(c_off);
        uint32_t* grp32 = reinterpret_cast<uint32_t*>(grp);
        const uint32_t n = td.d_hi + 2;
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            o32[i] = (uint32_t)(g.offsets[R_lo + i] - off_lo);
            w32[i] = (uint32_t)(g.win_offsets[R_lo + i] - win_lo);
        }
        const uint32_t n_slots = (uint32_t)(slot_end - slot_begin);
        const uint32_t n_items = Shape::n_items(n_slots);
        const uint32_t n_groups = (n_items + kCsrGroup - 1) / kCsrGroup;
        const uint32_t mis = stage_stretch<Eng::kValidate>(g.bases, g.n_bytes, g.packed, td.g0, td.span, Eng::kSpanEntries, enc, tile);
        if constexpr (!Eng::kTwoPhase) deferred_reset();
        __syncthreads();  // the offsets and the tile
        constexpr uint32_t kGroupSlots = kCsrGroup * kRun;
        for (uint32_t r = threadIdx.x; r <= td.d_last; r += blockDim.x) {  // every read marks the groups whose first slot it owns
            const uint32_t a = max(w32[r], p0), b = min(w32[r + 1], p0 + n_slots);
            if (b > a) {
                const uint32_t t1 = (b - 1 - p0) / kGroupSlots;
                for (uint32_t t = (a - p0 + kGroupSlots - 1) / kGroupSlots; t <= t1; ++t) grp32[t] = r;
            }
        }
        if (threadIdx.x == 0) grp32[n_groups] = td.d_last;
        __syncthreads();
        auto item = [&](uint32_t li, auto&& one, auto&& two, auto&& single) {
            csr_item32<Shape, kUnified>(li, n_slots, p0, slot_begin, R_lo, w32, o32, grp32, mis, one, two, single);
        };
        run_pass(eng, tile, K, n_items, item, true);
        return;
    }
    const uint64_t* off = g.offsets;  // tables indexed by absolute read number
    const uint64_t* win = g.win_offsets;
    if (R_hi - R_lo + 2 <= (uint64_t)kCsrCache + 2) {
        const uint32_t n = (uint32_t)(R_hi - R_lo + 2);
        for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
            c_off[i] = g.offsets[R_lo + i];
            c_win[i] = g.win_offsets[R_lo + i];
        }
        off = c_off - R_lo;
        win = c_win - R_lo;
    }

    uint64_t cur = slot_begin, r_cur = R_lo;
    // The common case: the offsets of the tile's reads are in shared memory and the whole tile is one pass whose stretch
    // csr_index_kernel already worked out.  Staging starts at once, next to the loads of the offsets (one barrier for
    // both), and the per-group owner table is filled by the reads themselves -- every read marks the groups whose first
    // slot it owns -- instead of one binary search per group.
    if (off != g.offsets && td.span != 0xFFFFFFFFu) {
        CsrPass ps;
        ps.slot_lo = slot_begin; ps.slot_hi = slot_end; ps.r_lo = R_lo; ps.r_hi = R_lo + td.d_last; ps.g0 = td.g0; ps.span = td.span;
        const uint32_t n_slots = (uint32_t)(ps.slot_hi - ps.slot_lo);
        const uint32_t n_items = Shape::n_items(n_slots);
        const uint32_t n_groups = (n_items + kCsrGroup - 1) / kCsrGroup;
        const uint32_t mis = stage_stretch<Eng::kValidate>(g.bases, g.n_bytes, g.packed, ps.g0, ps.span, Eng::kSpanEntries, enc, tile);
        if constexpr (!Eng::kTwoPhase) deferred_reset();
        __syncthreads();  // the offsets (loaded above) and the tile
        constexpr uint32_t kGroupSlots = kCsrGroup * kRun;
        for (uint64_t r = R_lo + threadIdx.x; r <= ps.r_hi; r += blockDim.x) {
            const uint64_t a = max(win[r], ps.slot_lo), b = min(win[r + 1], ps.slot_hi);
            if (b > a) {
                const uint32_t t1 = (uint32_t)((b - 1 - ps.slot_lo) / kGroupSlots);
                for (uint32_t t = (uint32_t)((a - ps.slot_lo + kGroupSlots - 1) / kGroupSlots); t <= t1; ++t) grp[t] = r;
            }
        }
        if (threadIdx.x == 0) grp[n_groups] = ps.r_hi;
        __syncthreads();
        auto item = [&](uint32_t li, auto&& one, auto&& two, auto&& single) {
            csr_item<Shape, kUnified>(li, n_slots, ps, off, win, grp, mis, one, two, single);
        };
        run_pass(eng, tile, K, n_items, item, true);
        return;
    }
    __syncthreads();  // the cached offsets
    while (cur < slot_end) {  // CTA-uniform; one pass unless a stretch of very short reads overflows the tile
        if (threadIdx.x == 0) {
            CsrPass ps;
            ps.slot_lo = cur;
            ps.r_lo = cur == slot_begin ? R_lo : last_le(win, r_cur, R_hi, cur);  // owner of slot `cur`
            ps.g0 = off[ps.r_lo] + (cur - win[ps.r_lo]);
            // The common case needs no search: everything up to the CTA's last slot fits one tile.  That slot's owner is
            // R_hi or, when R_hi starts exactly at slot_end, the nearest earlier read that has windows.
            uint64_t r_last = R_hi;
            while (win[r_last] > slot_end - 1) --r_last;
            const uint64_t g_end_all = off[r_last] + (slot_end - 1 - win[r_last]) + K;
            if (g_end_all - ps.g0 <= (uint64_t)tile_bases) {
                ps.slot_hi = slot_end;
                ps.r_hi = r_last;
                ps.span = (uint32_t)(g_end_all - ps.g0);
            } else {
                // windows starting before g_lim fit wholly in a tile that starts at g0
                const uint64_t g_lim = ps.g0 + tile_bases - K + 1;
                uint64_t lim = slot_end;
                if (g_lim < g.n_bases) {
                    const uint64_t r = last_le(off, ps.r_lo, R_hi, g_lim);
                    const uint64_t w_r = win[r + 1] - win[r];
                    lim = min(lim, win[r] + min(g_lim - off[r], w_r));  // slots whose window starts before g_lim
                }
                if (lim < slot_end && lim - cur >= (uint64_t)Shape::kAlign) lim = cur + ((lim - cur) / Shape::kAlign) * Shape::kAlign;  // keep items aligned
                ps.slot_hi = lim;
                ps.r_hi = last_le(win, ps.r_lo, R_hi, lim - 1);
                const uint64_t g_end = off[ps.r_hi] + (lim - 1 - win[ps.r_hi]) + K;
                ps.span = (uint32_t)(g_end - ps.g0);
            }
            *pass = ps;
        }
        __syncthreads();
        const CsrPass ps = *pass;
        const uint32_t n_slots = (uint32_t)(ps.slot_hi - ps.slot_lo);
        const uint32_t n_items = Shape::n_items(n_slots);
        // owner of the first slot of every group of kCsrGroup items: an item then searches a handful of reads, not the pass
        const uint32_t n_groups = (n_items + kCsrGroup - 1) / kCsrGroup;
        for (uint32_t t = threadIdx.x; t <= n_groups; t += blockDim.x)
            grp[t] = last_le(win, ps.r_lo, ps.r_hi, min(ps.slot_lo + (uint64_t)t * (kCsrGroup * kRun), ps.slot_hi - 1));
        const uint32_t mis = stage_stretch<Eng::kValidate>(g.bases, g.n_bytes, g.packed, ps.g0, ps.span, Eng::kSpanEntries, enc, tile);
        if constexpr (!Eng::kTwoPhase) deferred_reset();
        __syncthreads();

        auto item = [&](uint32_t li, auto&& one, auto&& two, auto&& single) {
            csr_item<Shape, kUnified>(li, n_slots, ps, off, win, grp, mis, one, two, single);
        };
        run_pass(eng, tile, K, n_items, item, ps.slot_hi == slot_end);
        __syncthreads();  // the next pass overwrites the tile
        cur = ps.slot_hi;
        r_cur = ps.r_hi;
    }
}

}  // namespace kmb
