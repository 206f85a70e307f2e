// kmb_extract_wide.cuh -- EXTENSION: canonical k-mers for 1 <= K <= 64 as two
// u64 words (little-endian word order), any of the 24 Naive encodings / Xor10.
//
// The reference defines no canonical form above 32 bases (naive_impl caps at
// 32, naive_impl/kmer.rs:211-213; kmer::Kmer<P,K,B> derives only Debug,
// kmer.rs:11).  The pieces that ARE pinned: the packed layout of
// Encoding::encode (encoding/naive.rs:116-124, goldens :388-445) and
// Encoding::rev_comp::<K> (naive.rs:138-154).  canonical = unsigned min of the
// 2K-bit integers, hash = 2K-bit pair reversal -- the natural extension of
// naive_impl/canonical_kmer.rs:113-119 and naive_impl/hash.rs:60-71.
//
// Same structure as kmb_extract.cuh: slot-space work items (8 windows per thread,
// 32-byte stores), a span of NW32+1 packed words per item, one multiword compare
// per window, hash by XOR fold.  The items have ShapePair (kmb_geometry.cuh): two
// neighbouring lanes interleave their windows so that a store covers 64 contiguous
// bytes per lane pair instead of 32 bytes in each of 32 different lines.
// NW32 = live 32-bit words of a k-mer = ceil(2K / 32) in {2, 3, 4} (K <= 32 uses 2).
#pragma once
#include "kmb_extract.cuh"

namespace kmb {

struct WideConst {
    uint32_t K;
    uint32_t shiftD;     // 2 * (16 * (NW32 + 1) - (ShapePair::kSpanSlots + K - 1))
    uint32_t mask_a;     // mask of word NW32 - 2
    uint32_t mask_b;     // mask of word NW32 - 1 (the top live word)
    uint32_t cmask;      // complement constant replicated over 16 fields
    uint32_t cm_a, cm_b; // cmask & mask_a / mask_b
    uint64_t kmask;      // low K bits (window validity)
};

struct WideOut {
    uint64_t* canon;  // 2 words per slot
    uint64_t* hash;   // 2 words per slot
    unsigned long long* digest;
    uint32_t vec_ok;
};

struct WideParams {
    WideConst wc;
    WideOut out;
};

template <int NW32>
struct WideSpan {
    uint32_t a[NW32 + 1];  // forward span, 16 * (NW32 + 1) bases
    uint32_t d[NW32 + 1];  // reverse complement of its first ShapePair::kSpanSlots + K - 1 bases, at bit 0
    uint64_t inv_lo;       // invalid-base bits 0..63 of the span
    uint32_t inv_hi;       // bits 64..95
};

template <int NA, int WS>
__device__ __forceinline__ void shr_words(uint32_t (&d)[NA], const uint32_t (&c)[NA], uint32_t s) {
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const uint32_t lo = (i + WS) < NA ? c[i + WS] : 0u;
        const uint32_t hi = (i + WS + 1) < NA ? c[i + WS + 1] : 0u;
        d[i] = __funnelshift_r(lo, hi, s);
    }
}

template <int NW32, bool VALIDATE>
__device__ __forceinline__ WideSpan<NW32> load_wide_span(const uint2* tile, uint32_t rel, const WideConst& wc) {
    constexpr int NA = NW32 + 1;
    const uint32_t e = rel >> 4, o2 = (rel & 15u) * 2;
    uint2 t[NA + 1];
#pragma unroll
    for (int i = 0; i < NA + 1; ++i) t[i] = tile[e + i];
    WideSpan<NW32> s;
#pragma unroll
    for (int i = 0; i < NA; ++i) s.a[i] = __funnelshift_r(t[i].x, t[i + 1].x, o2);
    uint32_t c[NA];
#pragma unroll
    for (int i = 0; i < NA; ++i) c[i] = pair_reverse32(s.a[NA - 1 - i] ^ wc.cmask);
    const uint32_t sh = wc.shiftD & 31u;
    switch (wc.shiftD >> 5) {  // kernel-uniform
        case 0: shr_words<NA, 0>(s.d, c, sh); break;
        case 1: shr_words<NA, 1>(s.d, c, sh); break;
        default: shr_words<NA, 2>(s.d, c, sh); break;
    }
    s.inv_lo = 0; s.inv_hi = 0;
    if (VALIDATE) {
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < NA + 1; ++i) any |= t[i].y;
        if (any != 0u) {
            // 16 * (NA + 1) <= 96 mask bits, shifted down by the item's offset inside entry e
            const uint64_t m0 = (uint64_t)t[0].y | ((uint64_t)t[1].y << 16) | ((uint64_t)t[2].y << 32) | ((uint64_t)t[3].y << 48);
            uint64_t m1 = 0;
            if constexpr (NA + 1 > 4) m1 |= (uint64_t)t[4].y;
            if constexpr (NA + 1 > 5) m1 |= (uint64_t)t[5].y << 16;
            const uint32_t o = o2 >> 1;
            s.inv_lo = o ? ((m0 >> o) | (m1 << (64 - o))) : m0;
            s.inv_hi = (uint32_t)(m1 >> o);
        }
    }
    return s;
}

struct WideWindow {
    uint64_t c0, c1, h0, h1;
};

// the window that starts j bases (= slots) after the span's first base
template <int NW32>
__device__ __forceinline__ WideWindow wide_window(const WideSpan<NW32>& s, int j, const WideConst& wc) {
    uint32_t f[4] = {0, 0, 0, 0}, r[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NW32; ++i) {
        f[i] = __funnelshift_r(s.a[i], s.a[i + 1], 2 * j);
        r[i] = __funnelshift_r(s.d[i], s.d[i + 1], 2 * (ShapePair::kSpanSlots - 1 - j));
    }
    f[NW32 - 2] &= wc.mask_a; r[NW32 - 2] &= wc.mask_a;
    f[NW32 - 1] &= wc.mask_b; r[NW32 - 1] &= wc.mask_b;
    // unsigned compare of the 2K-bit integers (strict '<', as canonical_kmer.rs:114): fw < rc <=> fw - rc borrows.
    // One subtract-with-borrow chain (5 instructions); the compiler's own 128-bit compare costs ~14.
    uint32_t t_, borrow;
    asm("sub.cc.u32 %0, %2, %6;\n\t"
        "subc.cc.u32 %0, %3, %7;\n\t"
        "subc.cc.u32 %0, %4, %8;\n\t"
        "subc.cc.u32 %0, %5, %9;\n\t"
        "subc.u32 %1, 0, 0;"
        : "=r"(t_), "=r"(borrow)
        : "r"(f[0]), "r"(f[1]), "r"(f[2]), "r"(f[3]), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]));
    const bool fw_less = borrow != 0u;
    uint32_t c[4] = {0, 0, 0, 0}, h[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < NW32; ++i) {
        c[i] = fw_less ? f[i] : r[i];
        const uint32_t cm = (i == NW32 - 1) ? wc.cm_b : ((i == NW32 - 2) ? wc.cm_a : wc.cmask);
        h[i] = (f[i] ^ r[i] ^ cm) ^ c[i];  // other strand ^ complement constant = pair reversal of the canonical strand
    }
    WideWindow w;
    w.c0 = mk64(c[0], c[1]); w.c1 = mk64(c[2], c[3]);
    w.h0 = mk64(h[0], h[1]); w.h1 = mk64(h[2], h[3]);
    return w;
}

template <int NW32>
__device__ __forceinline__ bool wide_ok(const WideSpan<NW32>& s, int j, const WideConst& wc) {
    const uint64_t x = j ? ((s.inv_lo >> j) | ((uint64_t)s.inv_hi << (64 - j))) : s.inv_lo;
    return (x & wc.kmask) == 0ull;
}

struct WideAcc {
    uint64_t canon = 0, hash = 0;
    uint32_t valid = 0;
};

// One item = 4 chunks of two neighbouring slots (ShapePair: slots 4i, 4i+1 from the item's first).  TWO / CHECK as in
// emit_run; a slot s exists iff s < nwin and belongs to span B iff s >= n_first.
template <int NW32, bool TWO, bool CHECK, bool DIGEST, bool HASH>
__device__ __forceinline__ void emit_wide_run(const WideSpan<NW32>& A, const WideSpan<NW32>& B, uint32_t n_first,
                                              const WideConst& wc, const WideOut& o, uint64_t slot0, uint32_t nwin,
                                              WideAcc& acc) {
#pragma unroll
    for (int i = 0; i < kRun / 2; ++i) {  // two windows = one 32-byte store per array
        WideWindow w[2];
        bool ok[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int s_off = ShapePair::off(2 * i + t);
            WideSpan<NW32> s = A;
            if (TWO && (uint32_t)s_off >= n_first) s = B;
            w[t] = wide_window<NW32>(s, s_off, wc);
            ok[t] = true;
            if (CHECK) ok[t] = wide_ok<NW32>(s, s_off, wc);
            if (DIGEST && ok[t] && (uint32_t)s_off < nwin) {
                acc.canon += w[t].c0 + w[t].c1; acc.hash += w[t].h0 + w[t].h1; acc.valid += 1;
            }
            if (CHECK && !ok[t]) { w[t].c0 = w[t].c1 = w[t].h0 = w[t].h1 = ~0ull; }
        }
        const int s0 = ShapePair::off(2 * i);
        const uint64_t slot = slot0 + s0;
        if ((uint32_t)s0 + 1 < nwin && o.vec_ok && (slot & 1ull) == 0ull) {
            if (o.canon) st_stream_v4u64(o.canon + 2 * slot, w[0].c0, w[0].c1, w[1].c0, w[1].c1);
            if (HASH && o.hash) st_stream_v4u64(o.hash + 2 * slot, w[0].h0, w[0].h1, w[1].h0, w[1].h1);
        } else {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                if ((uint32_t)(s0 + t) < nwin) {
                    if (o.canon) st_stream_v2u64(o.canon + 2 * (slot + t), w[t].c0, w[t].c1);
                    if (HASH && o.hash) st_stream_v2u64(o.hash + 2 * (slot + t), w[t].h0, w[t].h1);
                }
            }
        }
    }
}

template <int NW32, bool VALIDATE, bool DIGEST>
__device__ __forceinline__ void emit_wide_single(const uint2* tile, uint32_t rel, const WideConst& wc, const WideOut& o,
                                                 uint64_t slot, WideAcc& acc) {
    const WideSpan<NW32> s = load_wide_span<NW32, VALIDATE>(tile, rel, wc);
    WideWindow w = wide_window<NW32>(s, 0, wc);
    const bool ok = !VALIDATE || wide_ok<NW32>(s, 0, wc);
    if (DIGEST && ok) { acc.canon += w.c0 + w.c1; acc.hash += w.h0 + w.h1; acc.valid += 1; }
    if (!ok) { w.c0 = w.c1 = w.h0 = w.h1 = ~0ull; }
    if (o.canon) st_stream_v2u64(o.canon + 2 * slot, w.c0, w.c1);
    if (o.hash) st_stream_v2u64(o.hash + 2 * slot, w.h0, w.h1);
}

template <bool DIGEST>
__device__ __forceinline__ void wide_reduce(unsigned long long (&red)[3][32], unsigned long long* digest,
                                            const WideAcc& acc) {
    if (!DIGEST) return;
    const uint64_t v = warp_sum64(acc.valid), c = warp_sum64(acc.canon), h = warp_sum64(acc.hash);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = v; red[1][warp] = c; red[2][warp] = h; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long s = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        atomicAdd(digest + threadIdx.x, s);
    }
}

// The two-word engine plugged into the geometry of kmb_geometry.cuh (kernels: fixed_kernel / csr_kernel).
template <int NW32, bool VALIDATE, bool DIGEST, bool HASH = true>
struct WideEng {
    using Params = WideParams;
    using Span = WideSpan<NW32>;
    static constexpr bool kValidate = VALIDATE;
    using Shape = ShapePair;
    static constexpr bool kTwoPhase = false, kCountOnly = false;
    static constexpr int kSpanEntries = NW32 + 2;  // tile entries one span reads
    static constexpr int kMinCtas = 0, kMinCtasCsr = DIGEST ? 3 : 4;  // resident CTAs/SM the register allocation must allow (0: compiler's choice)
    const WideParams& p;
    WideAcc acc;
    __device__ explicit WideEng(const WideParams& params) : p(params) {}
    __device__ __forceinline__ uint32_t K() const { return p.wc.K; }
    __device__ __forceinline__ Span load(const uint2* tile, uint32_t rel) const { return load_wide_span<NW32, VALIDATE>(tile, rel, p.wc); }
    __device__ __forceinline__ bool dirty(const Span& s) const { return (s.inv_lo | (uint64_t)s.inv_hi) != 0ull; }
    template <bool TWO, bool CHECK>
    __device__ __forceinline__ void run(const Span& a, const Span& b, uint32_t n_first, uint64_t slot0, uint32_t nwin, const ItemCtx&) {
        emit_wide_run<NW32, TWO, CHECK, DIGEST, HASH>(a, b, n_first, p.wc, p.out, slot0, nwin, acc);
    }
    __device__ __forceinline__ void single(const uint2* tile, uint32_t rel, uint64_t slot, const ItemCtx&) {
        emit_wide_single<NW32, VALIDATE, DIGEST>(tile, rel, p.wc, p.out, slot, acc);
    }
    __device__ __forceinline__ void finish(unsigned long long (&red)[3][32]) {
        wide_reduce<DIGEST>(red, p.out.digest, acc);
    }
};

}  // namespace kmb
