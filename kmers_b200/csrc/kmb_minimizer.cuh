// kmb_minimizer.cuh -- "next" row N1: minimizers.
//
//   Kmer::minimizer_word (naive_impl/kmer.rs:170-191)            -> minimizer_words_kernel (element-wise)
//   SeqVecMinimizerIter  (naive_impl/seq_vector/minimizers.rs:38-142) -> MinimizerEng (one (lmer, pos) per k-mer window)
//
// Both pick the LEFTMOST w-mer of minimum hash_one(LexHasherState(hash_k), lmer: u64) inside a k-mer
// (kmer.rs:183 strict '<'; minimizers.rs:72-78 keeps the older deque entry on ties).  The reference's monotone
// deque is inherently sequential; on the GPU every work item (8 consecutive k-mer windows) splits its positions
// into [suffix of the first 7 | part common to all 8 windows | prefix of the last 7] (van Herk / Gil-Werman), so
// a window costs ~(m + 23) / 8 compare steps instead of m = k - w + 1.
// LexHash of an lmer is its lexicographic rank, which is simply the lmer read from the PAIR-REVERSED span
// (hash.rs:60-71), so no per-lmer bit reversal is needed: the span is reversed once per item.
// The kernel is ALU-bound, so the candidate compare is specialised on the lmer width (CLS):
//   0: w <= 13  rank and position share one 32-bit key (rank << 6 | position): one min per candidate
//   1: w <= 15  32-bit rank + position
//   2: w <= 32  64-bit rank + position
#pragma once
#include "kmb_extract.cuh"

namespace kmb {

struct MinConst {
    WinConst wc;
    uint32_t w;            // minimizer width
    uint32_t m;            // lmers per k-mer = K - w + 1
    uint32_t hmask32;      // CLS 0/1: the compared bits of an lmer rank (low 2w bits without the low 2*(w - hash_k)); CLS 0: << 6
    uint64_t hmask64;      // CLS 2: hash_k < w keeps only the first hash_k bases (hash.rs:69), comparing the masked ranks is
                           //        comparing the shifted ones
    uint32_t vm0, vm1, vm2;  // complement constant over the valid bits of the shifted reverse-complement span
    uint64_t wmask;        // low 2w bits
};

struct MinOut {
    uint64_t* mmer;  // lmer word (forward strand, base 0 in bits 1:0) per slot
    uint32_t* pos;   // position of the lmer inside its read per slot
    uint32_t vec_ok;
};

struct MinParams {
    MinConst mc;
    MinOut out;
};

// bits [s, s + 64) of the 96-bit value x2:x1:x0
__device__ __forceinline__ uint64_t bits96(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t s) {
    const uint32_t ws = s >> 5, b = s & 31u;
    const uint32_t w0 = ws == 0 ? x0 : (ws == 1 ? x1 : (ws == 2 ? x2 : 0u));
    const uint32_t w1 = ws == 0 ? x1 : (ws == 1 ? x2 : 0u);
    const uint32_t w2 = ws == 0 ? x2 : 0u;
    return mk64(__funnelshift_r(w0, w1, b), __funnelshift_r(w1, w2, b));
}

struct Cand {
    uint64_t h;
    uint32_t p;
};

// Leftmost-minimum lmer position (relative to the span's first base) for the 8 windows of a span.
__device__ __forceinline__ void span_minimizers64(const Span& s, const MinConst& mc, uint32_t (&best)[kRun]) {
    // P = pair reversal of the span's first kRun+K-1 bases: field f <-> base kRun+K-2-f, so the lmer at base p,
    // read as a number, is its lexicographic rank: v_p = (P >> 2*(kRun+K-1-w-p)) & wmask.
    const uint32_t p0 = s.d0 ^ mc.vm0, p1 = s.d1 ^ mc.vm1, p2 = s.d2 ^ mc.vm2;
    const uint32_t m = mc.m;
    if (m >= (uint32_t)kRun) {
        // positions m-1 .. 0 by walking a register down 2 bits at a time (p = m-1 sits at bit 2*(kRun-1))
        uint64_t r_lo = bits96(p0, p1, p2, 2 * (kRun - 1));
        uint32_t r_hi = p2 >> (2 * (kRun - 1));  // bits 64+ of P >> 14, for the walk
        Cand c{~0ull, 0};
        for (uint32_t p = m - 1; p >= (uint32_t)(kRun - 1); --p) {  // common part [kRun-1, m-1], right to left: '<=' keeps the left
            const uint64_t hv = r_lo & mc.hmask64;
            if (hv <= c.h) { c.h = hv; c.p = p; }
            r_lo = (r_lo >> 2) | ((uint64_t)r_hi << 62);
            r_hi >>= 2;
        }
        Cand suf[kRun];  // suf[j] = leftmost min over [j, kRun-2]
        suf[kRun - 1] = Cand{~0ull, 0};
        Cand cur{~0ull, 0};
#pragma unroll
        for (int p = kRun - 2; p >= 0; --p) {
            const uint64_t hv = r_lo & mc.hmask64;
            if (hv <= cur.h) { cur.h = hv; cur.p = (uint32_t)p; }
            suf[p] = cur;
            r_lo = (r_lo >> 2) | ((uint64_t)r_hi << 62);
            r_hi >>= 2;
        }
        Cand pre{~0ull, 0};  // leftmost min over [m, m+j-1], grown to the right: strict '<'
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            Cand r = suf[j];
            if (c.h < r.h) r = c;
            if (j > 0) {
                const uint64_t hv = bits96(p0, p1, p2, 2 * (kRun - 1 - j)) & mc.hmask64;  // position m + j - 1
                if (hv < pre.h) { pre.h = hv; pre.p = m + j - 1; }
                if (pre.h < r.h) r = pre;
            }
            best[j] = r.p;
        }
    } else {
        // few lmers per k-mer: plain scan of each window, left to right
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            Cand r{~0ull, (uint32_t)j};
            for (uint32_t q = 0; q < m; ++q) {
                const uint32_t p = j + q;
                const uint64_t hv = bits96(p0, p1, p2, 2 * (kRun + mc.wc.K - 1 - mc.w - p)) & mc.hmask64;
                if (q == 0 || hv < r.h) { r.h = hv; r.p = p; }
            }
            best[j] = r.p;
        }
    }
}

// ---- candidate keys ------------------------------------------------------------------------------------------
// Leftmost-minimum bookkeeping: take_left(cur, c) merges a candidate that lies LEFT of everything in cur (ties go to
// the candidate), take_right one that lies to the right (ties keep cur).
template <int CLS>
struct MinKey;
template <>
struct MinKey<0> {  // positions make every key distinct, so both merges are a plain min
    uint32_t k;
    __device__ static __forceinline__ MinKey worst() { return {0xFFFFFFFFu}; }
    __device__ static __forceinline__ MinKey make(uint32_t x, uint32_t p, const MinConst& mc) { return {(x & mc.hmask32) | p}; }
    __device__ __forceinline__ void take_left(const MinKey& c) { k = min(k, c.k); }
    __device__ __forceinline__ void take_right(const MinKey& c) { k = min(k, c.k); }
    __device__ __forceinline__ uint32_t pos() const { return k & 63u; }
};
template <>
struct MinKey<1> {
    uint32_t h, p;
    __device__ static __forceinline__ MinKey worst() { return {0xFFFFFFFFu, 0u}; }  // ranks are < 2^30
    __device__ static __forceinline__ MinKey make(uint32_t x, uint32_t p, const MinConst& mc) { return {x & mc.hmask32, p}; }
    __device__ __forceinline__ void take_left(const MinKey& c) { if (c.h <= h) *this = c; }
    __device__ __forceinline__ void take_right(const MinKey& c) { if (c.h < h) *this = c; }
    __device__ __forceinline__ uint32_t pos() const { return p; }
};

// 32 bits of x2:x1:x0 from bit sh (< 64) on
__device__ __forceinline__ uint32_t bits32(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t sh) {
    return sh < 32u ? __funnelshift_r(x0, x1, sh) : __funnelshift_r(x1, x2, sh - 32u);
}

// Leftmost-minimum lmer position (relative to the span's first base) for the kRun windows of a span, w <= 15.
template <int CLS>
__device__ __forceinline__ void span_minimizers32(const Span& s, const MinConst& mc, uint32_t (&best)[kRun]) {
    using Key = MinKey<CLS>;
    // P = pair reversal of the span's first kRun+K-1 bases: the lmer at base p, read as a number, is its rank:
    // v_p = (P >> 2*(kRun+K-1-w-p)) & wmask = (P >> 2*(m+kRun-2-p)) & wmask.  CLS 0 keeps P << 6 so that an extract
    // leaves room for the position in the low 6 bits.
    uint32_t p0 = s.d0 ^ mc.vm0, p1 = s.d1 ^ mc.vm1, p2 = s.d2 ^ mc.vm2;
    if (CLS == 0) { p2 = __funnelshift_l(p1, p2, 6); p1 = __funnelshift_l(p0, p1, 6); p0 <<= 6; }
    const uint32_t m = mc.m;
    if (m >= (uint32_t)kRun) {
        // common part [kRun-1, m-1], right to left: position m-1-t sits at bit 2*(kRun-1) + 2t
        Key c = Key::worst();
        const uint32_t tmax = m - (uint32_t)kRun;  // <= 24
#pragma unroll
        for (uint32_t t = 0; t <= 8u; ++t) {  // compile-time shifts; the kernel-uniform guard becomes a uniform branch
            if (t > tmax) break;
            c.take_left(Key::make(__funnelshift_r(p0, p1, 2 * (kRun - 1) + 2 * t), m - 1 - t, mc));
        }
        for (uint32_t t = 9; t <= tmax; ++t) c.take_left(Key::make(__funnelshift_r(p1, p2, 2 * t - 18), m - 1 - t, mc));
        // suffixes of [0, kRun-2]: Q = P >> 2m puts position p at bit 2*(kRun-2-p)
        uint32_t q0, q1, q2;
        shr96(q0, q1, q2, p0, p1, p2, 2 * m);
        Key suf[kRun];
        suf[kRun - 1] = Key::worst();
        Key cur = Key::worst();
#pragma unroll
        for (int p = kRun - 2; p >= 0; --p) {
            cur.take_left(Key::make(__funnelshift_r(q0, q1, 2 * (kRun - 2 - p)), (uint32_t)p, mc));
            suf[p] = cur;
        }
        // prefixes of [m, m+kRun-2]: position m+j-1 sits at bit 2*(kRun-1-j)
        Key pre = Key::worst();
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            Key r = suf[j];
            r.take_right(c);
            if (j > 0) {
                pre.take_right(Key::make(__funnelshift_r(p0, p1, 2 * (kRun - 1 - j)), m + j - 1, mc));
                r.take_right(pre);
            }
            best[j] = r.pos();
        }
    } else {
        // few lmers per k-mer: plain scan of each window, left to right (shift 2*(m+kRun-2-p) < 32)
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            Key r = Key::make(__funnelshift_r(p0, p1, 2 * (m + kRun - 2 - j)), (uint32_t)j, mc);
#pragma unroll
            for (int q = 1; q < kRun - 1; ++q)
                if ((uint32_t)q < m) r.take_right(Key::make(__funnelshift_r(p0, p1, 2 * (m + kRun - 2 - j - q)), (uint32_t)(j + q), mc));
            best[j] = r.pos();
        }
    }
}

template <int CLS>
__device__ __forceinline__ void span_minimizers(const Span& s, const MinConst& mc, uint32_t (&best)[kRun]) {
    if constexpr (CLS == 2) span_minimizers64(s, mc, best);
    else span_minimizers32<CLS>(s, mc, best);
}

// the lmer word at base bp of the forward span
template <int CLS>
__device__ __forceinline__ uint64_t lmer_at(const Span& s, uint32_t bp, const MinConst& mc) {
    if constexpr (CLS == 2) {
        return bits96(s.a0, s.a1, s.a2, 2 * bp) & mc.wmask;
    } else {  // w <= 15: one 32-bit extract
        const uint32_t sh = 2 * bp, ws = sh >> 5;
        const uint32_t lo = ws == 0 ? s.a0 : (ws == 1 ? s.a1 : s.a2);
        const uint32_t hi = ws == 0 ? s.a1 : (ws == 1 ? s.a2 : 0u);
        return (uint64_t)(__funnelshift_r(lo, hi, sh) & (uint32_t)mc.wmask);
    }
}

template <bool VALIDATE, int CLS>
struct MinimizerEng {
    using Params = MinParams;
    using Span = kmb::Span;
    static constexpr bool kValidate = VALIDATE;
    using Shape = ShapeRun;
    static constexpr bool kTwoPhase = false, kCountOnly = false;
    static constexpr int kSpanEntries = 4;
    static constexpr int kMinCtas = 4, kMinCtasCsr = 4;
    const MinParams& p;
    __device__ explicit MinimizerEng(const MinParams& params) : p(params) {}
    __device__ __forceinline__ uint32_t K() const { return p.mc.wc.K; }
    __device__ __forceinline__ Span load(const uint2* tile, uint32_t rel) const { return load_span<VALIDATE>(tile, rel, p.mc.wc); }
    __device__ __forceinline__ bool dirty(const Span& s) const { return s.inv != 0ull; }

    template <bool TWO, bool CHECK>
    __device__ __forceinline__ void run(const Span& a, const Span& b, uint32_t n_first, uint64_t slot0, uint32_t nwin, const ItemCtx& ic) {
        uint32_t ba[kRun], bb[kRun];
        span_minimizers<CLS>(a, p.mc, ba);
        if (TWO) span_minimizers<CLS>(b, p.mc, bb);
        uint64_t om[kRun];
        uint32_t op[kRun];
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            const bool second = TWO && (uint32_t)j >= n_first;
            const Span& s = second ? b : a;
            const uint32_t bp = second ? bb[j] : ba[j];
            bool ok = true;
            if (CHECK) ok = (((uint32_t)(s.inv >> j)) & p.mc.wc.kmask) == 0u;
            om[j] = ok ? lmer_at<CLS>(s, bp, p.mc) : ~0ull;
            // span B was loaded n_first bases before its read's first base
            op[j] = ok ? (second ? bp - n_first : (uint32_t)ic.pos_a + bp) : 0xFFFFFFFFu;
        }
        if (nwin == kRun && p.out.vec_ok && (slot0 & 7ull) == 0ull) {
            if (p.out.mmer) {
                st_stream_v4u64(p.out.mmer + slot0, om[0], om[1], om[2], om[3]);
                st_stream_v4u64(p.out.mmer + slot0 + 4, om[4], om[5], om[6], om[7]);
            }
            if (p.out.pos)
                st_stream_v4u64(reinterpret_cast<uint64_t*>(p.out.pos + slot0), mk64(op[0], op[1]), mk64(op[2], op[3]),
                                mk64(op[4], op[5]), mk64(op[6], op[7]));
        } else {
#pragma unroll
            for (int j = 0; j < kRun; ++j) {
                if ((uint32_t)j < nwin) {
                    if (p.out.mmer) st_stream_u64(p.out.mmer + slot0 + j, om[j]);
                    if (p.out.pos) p.out.pos[slot0 + j] = op[j];
                }
            }
        }
    }
    __device__ __forceinline__ void single(const uint2* tile, uint32_t rel, uint64_t slot, const ItemCtx& ic) {
        const Span s = load_span<VALIDATE>(tile, rel, p.mc.wc);
        uint32_t best[kRun];
        span_minimizers<CLS>(s, p.mc, best);
        const bool ok = !VALIDATE || (((uint32_t)s.inv) & p.mc.wc.kmask) == 0u;
        if (p.out.mmer) st_stream_u64(p.out.mmer + slot, ok ? lmer_at<CLS>(s, best[0], p.mc) : ~0ull);
        if (p.out.pos) p.out.pos[slot] = ok ? (uint32_t)ic.pos_a + best[0] : 0xFFFFFFFFu;
    }
    __device__ __forceinline__ void finish(unsigned long long (&)[3][32]) {}
};

}  // namespace kmb
