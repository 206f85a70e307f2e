// kmb_tu_minimizer.cu -- instantiates the minimizer engines on both geometries, and Kmer::minimizer_word.
#include <type_traits>

#include "kmb_launch.h"

namespace kmb {

cudaError_t launch_minimizers(bool validate, const FixedGeom* fg, const CsrGeom* cg, const Launch& l, cudaStream_t st,
                              const EncDesc& enc, const MinParams& ep) {
    const int cls = ep.mc.w <= 13 ? 0 : (ep.mc.w <= 15 ? 1 : 2);  // width class of the candidate compare (kmb_minimizer.cuh)
#define KMB_CASE(V, C) \
    if (validate == V && cls == C) return launch_eng<MinimizerEng<V, C>>(fg, cg, l, st, enc, ep);
    KMB_CASE(true, 0) KMB_CASE(true, 1) KMB_CASE(true, 2) KMB_CASE(false, 0) KMB_CASE(false, 1) KMB_CASE(false, 2)
#undef KMB_CASE
    return cudaErrorInvalidValue;
}

// Kmer::minimizer_word (naive_impl/kmer.rs:170-191) on every word: leftmost width-mer of minimum LexHash.
//
// LexHasher::write_u64 (hash.rs:60-71) of an lmer = its first hash_k bases read as a number, base 0 most significant.
// Pair-reversing the whole k-mer ONCE puts base 0 in the top field, so the rank of the lmer at `pos` is a plain field
// extract of R = pair_reverse(word) >> 2 (32 - k):  v = (R >> 2 (m - 1 - pos)) & wmask, m = k - width + 1 lmers; the hash is
// v >> 2 (width - hash_k) for hash_k < width (dropping low bits = masking them for the purpose of comparing) and
// v << 2 (hash_k - width) otherwise, which orders and ties exactly like v.
// The kernel is issue-bound (m candidates per word against 20 bytes of traffic), so the candidate step is specialised
// on the width like MinimizerEng's (kmb_minimizer.cuh):
//   CLS 0  width <= 13: rank and position share one 32-bit key (rank << 6 | pos) -- one SHF, one LOP3, one MIN per candidate
//   CLS 1  width <= 16: 32-bit rank, candidates walked right to left so that '<=' keeps the leftmost minimum
//   CLS 2  wider      : 64-bit ranks (at most 16 candidates)
// and every thread works on kMinWordsPerThread words at once.
// words per thread (% of the copy peak at 10^8 words, k=31 w=15): 2 70, 4 75, 6 79, 8 77, 12 77, 16 78
#ifndef KMB_MINWORDS
#define KMB_MINWORDS 6
#endif
constexpr int kMinWordsPerThread = KMB_MINWORDS;

// 2^i as a constant-bank operand.  Integer shifts run on the ALU pipe, which this kernel saturates; IMAD runs on the FMA
// pipe, which it leaves idle (both issue one warp instruction per 2 cycles per SM sub-partition).  A shift written as a
// multiplication by a power of two that ptxas cannot see through (it comes from constant memory) is issued as
// IMAD / IMAD.HI and moves to the idle pipe:  x >> s == umulhi(x, 2^(32-s)),  x << s == x * 2^s.
__constant__ uint32_t kPow2[32] = {1u << 0,  1u << 1,  1u << 2,  1u << 3,  1u << 4,  1u << 5,  1u << 6,  1u << 7,  1u << 8,  1u << 9,  1u << 10,
                                   1u << 11, 1u << 12, 1u << 13, 1u << 14, 1u << 15, 1u << 16, 1u << 17, 1u << 18, 1u << 19, 1u << 20, 1u << 21,
                                   1u << 22, 1u << 23, 1u << 24, 1u << 25, 1u << 26, 1u << 27, 1u << 28, 1u << 29, 1u << 30, 1u << 31};

// low 32 bits of (hi:lo) >> S, S a compile-time constant in [0, 64), through the FMA pipe
template <int S>
__device__ __forceinline__ uint32_t shr64_lo_fma(uint32_t lo, uint32_t hi) {
    if constexpr (S == 0) return lo;
    else if constexpr (S < 32) return hi * kPow2[32 - S] + __umulhi(lo, kPow2[32 - S]);
    else if constexpr (S == 32) return hi;
    else return __umulhi(hi, kPow2[64 - S]);
}
template <int S>
__device__ __forceinline__ uint32_t shr64_lo_alu(uint32_t lo, uint32_t hi) {
    if constexpr (S == 0) return lo;
    else if constexpr (S < 32) return __funnelshift_r(lo, hi, S);
    else return hi >> (S - 32);
}

// Widths 14 and 15: a rank has 2w = 28 / 30 bits, which leaves SPARE = 4 / 2 bits of a 32-bit key -- not enough for a
// position (up to 19 candidates), but enough for the position INSIDE a group of G = 2^SPARE neighbours.  Inside a group
// one unsigned min per candidate on (rank << SPARE | j) keeps the leftmost minimum; across groups, left to right, a group
// only replaces the best so far when its rank is strictly smaller: key_new < (key_best & ~(G-1)).  ~1.8 ALU operations per
// candidate instead of 5 (extract, mask, min, compare, select), the key construction going through the FMA pipe for
// three candidates out of four.  hmask (hash_k < width) is folded into the key mask.
template <int SPARE, int M, bool MASKED, int P>
__device__ __forceinline__ uint32_t group_key(uint32_t lo, uint32_t hi, uint32_t keymask) {
    constexpr int G = 1 << SPARE, J = P % G, T = M - 1 - P, S = 2 * T;  // the candidate's rank = low 2w bits of R >> S
    uint32_t key;
    if constexpr (S == 0 && !MASKED) {
        key = lo * kPow2[SPARE] + (uint32_t)J;
    } else if constexpr (S < SPARE) {  // the net shift is to the left; bits of the base below land in the J field and are masked off
        key = ((lo * kPow2[SPARE - S]) & keymask) | (uint32_t)J;
    } else if constexpr (MASKED || (T % 4) == 3) {
        // ALU form: bits [S - SPARE, ...) so that the rank already sits above the SPARE low bits, which one LOP3 replaces by J
        key = (shr64_lo_alu<S - SPARE>(lo, hi) & keymask) | (uint32_t)J;
    } else {
        key = shr64_lo_fma<S>(lo, hi) * kPow2[SPARE] + (uint32_t)J;  // the multiplication also drops the bits above the rank
    }
    return key;
}

template <int I, int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

template <int SPARE, int M, bool MASKED>
__device__ __forceinline__ uint32_t leftmost_min_grouped(uint64_t word, uint32_t k, uint32_t keymask) {
    const uint64_t R = pair_reverse64(word) >> (2 * (32 - k));
    const uint32_t lo = (uint32_t)R, hi = (uint32_t)(R >> 32);
    constexpr int G = 1 << SPARE, NG = (M + G - 1) / G;
    constexpr uint32_t JM = (uint32_t)G - 1u;
    uint32_t best = 0xFFFFFFFFu, bg = 0;
    static_for<0, NG>([&](auto gi) {
        constexpr int g = decltype(gi)::value;
        uint32_t gm = 0xFFFFFFFFu;
        static_for<0, G>([&](auto ji) {
            constexpr int P = g * G + decltype(ji)::value;
            if constexpr (P < M) gm = min(gm, group_key<SPARE, M, MASKED, P>(lo, hi, keymask));
        });
        if constexpr (g == 0) {
            best = gm;
        } else {
            const bool take = gm < (best & ~JM);  // strictly smaller rank: ties stay with the group on the left
            best = take ? gm : best;
            bg = take ? (uint32_t)g : bg;
        }
    });
    return bg * (uint32_t)G + (best & JM);
}

// M = number of candidate lmers (k - width + 1), a template parameter so that every shift amount is an immediate and the
// candidate loop has no counters, guards or branches: 3 instructions per candidate for CLS 0, 5 for CLS 1.
template <int CLS, int M>
__device__ __forceinline__ uint32_t leftmost_min_pos(uint64_t word, uint32_t k, uint32_t hmask32, uint64_t hmask64) {
    const uint64_t R = pair_reverse64(word) >> (2 * (32 - k));
    const uint32_t lo = (uint32_t)R, hi = (uint32_t)(R >> 32);
    if constexpr (CLS == 0) {
        // key(t) = ((R >> 2t) << 6) & hmask32 | pos, pos = M - 1 - t; for t < 3 the net shift is to the left
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int t = 0; t < M; ++t) {
            const uint32_t x = t < 3 ? lo << (6 - 2 * t) : (t < 19 ? __funnelshift_r(lo, hi, 2 * t - 6) : hi >> (2 * t - 38));
            best = min(best, (x & hmask32) | (uint32_t)(M - 1 - t));
        }
        return best & 63u;
    } else if constexpr (CLS == 1) {
        uint32_t best = 0xFFFFFFFFu, pos = 0;
#pragma unroll
        for (int t = 0; t < M; ++t) {  // right to left: '<=' keeps the leftmost minimum
            const uint32_t h = (t < 16 ? __funnelshift_r(lo, hi, 2 * t) : hi >> (2 * t - 32)) & hmask32;
            const bool le = h <= best;
            best = le ? h : best;
            pos = le ? (uint32_t)(M - 1 - t) : pos;
        }
        return pos;
    } else {
        uint64_t best = ~0ull, r = R;
        uint32_t pos = 0;
#pragma unroll
        for (int t = 0; t < M; ++t, r >>= 2) {
            const uint64_t h = r & hmask64;
            if (h <= best) { best = h; pos = (uint32_t)(M - 1 - t); }
        }
        return pos;
    }
}

// CLS: 0 / 1 / 2 as above; 3 = grouped keys, width 15 (SPARE 2); 4 = width 14 (SPARE 4); 5, 6 = the same with hash_k < width
template <int CLS, int M>
__global__ void __launch_bounds__(256) minimizer_words_kernel(const uint64_t* __restrict__ in, uint64_t n, uint32_t k, uint32_t width,
                                                              uint32_t hmask32, uint64_t hmask64, uint64_t* __restrict__ mmer_out,
                                                              uint32_t* __restrict__ offset_out) {
    const uint64_t base = (uint64_t)blockIdx.x * (256 * kMinWordsPerThread) + threadIdx.x;
    const uint64_t wmask = width >= 32 ? ~0ull : ((1ull << (2 * width)) - 1ull);
    uint64_t word[kMinWordsPerThread];
#pragma unroll
    for (int u = 0; u < kMinWordsPerThread; ++u) {
        const uint64_t i = base + (uint64_t)u * 256;
        word[u] = i < n ? __ldg(in + i) : 0ull;
    }
#pragma unroll
    for (int u = 0; u < kMinWordsPerThread; ++u) {
        const uint64_t i = base + (uint64_t)u * 256;
        uint32_t off;
        if constexpr (CLS == 3) off = leftmost_min_grouped<2, M, false>(word[u], k, hmask32);
        else if constexpr (CLS == 4) off = leftmost_min_grouped<4, M, false>(word[u], k, hmask32);
        else if constexpr (CLS == 5) off = leftmost_min_grouped<2, M, true>(word[u], k, hmask32);
        else if constexpr (CLS == 6) off = leftmost_min_grouped<4, M, true>(word[u], k, hmask32);
        else off = leftmost_min_pos<CLS, M>(word[u], k, hmask32, hmask64);
        if (i < n) {
            if (mmer_out) mmer_out[i] = (word[u] >> (2 * off)) & wmask;  // sub_kmer_word, kmer.rs:155-161
            if (offset_out) offset_out[i] = off;
        }
    }
}

template <int CLS, int M>
static cudaError_t launch_mw(const uint64_t* in, uint64_t n, uint32_t k, uint32_t w, uint32_t hmask32, uint64_t hmask64, uint64_t* mmer_out,
                             uint32_t* offset_out, cudaStream_t st) {
    const unsigned grid = (unsigned)((n + 256 * kMinWordsPerThread - 1) / (256 * kMinWordsPerThread));
    minimizer_words_kernel<CLS, M><<<grid, 256, 0, st>>>(in, n, k, w, hmask32, hmask64, mmer_out, offset_out);
    return cudaGetLastError();
}

template <int CLS>
static cudaError_t launch_mw_cls(uint32_t m, const uint64_t* in, uint64_t n, uint32_t k, uint32_t w, uint32_t hmask32, uint64_t hmask64,
                                 uint64_t* mmer_out, uint32_t* offset_out, cudaStream_t st) {
    switch (m) {
#define KMB_M(M) case M: return launch_mw<CLS, M>(in, n, k, w, hmask32, hmask64, mmer_out, offset_out, st);
        KMB_M(1) KMB_M(2) KMB_M(3) KMB_M(4) KMB_M(5) KMB_M(6) KMB_M(7) KMB_M(8) KMB_M(9) KMB_M(10) KMB_M(11) KMB_M(12) KMB_M(13) KMB_M(14)
        KMB_M(15) KMB_M(16) KMB_M(17) KMB_M(18) KMB_M(19)
#undef KMB_M
    }
    if constexpr (CLS == 0) {  // width <= 13 leaves up to 32 candidates
        switch (m) {
#define KMB_M(M) case M: return launch_mw<0, M>(in, n, k, w, hmask32, hmask64, mmer_out, offset_out, st);
            KMB_M(20) KMB_M(21) KMB_M(22) KMB_M(23) KMB_M(24) KMB_M(25) KMB_M(26) KMB_M(27) KMB_M(28) KMB_M(29) KMB_M(30) KMB_M(31) KMB_M(32)
#undef KMB_M
        }
    }
    return cudaErrorInvalidValue;
}

cudaError_t launch_minimizer_words(const uint64_t* in, uint64_t n, uint32_t k, uint32_t w, uint32_t hash_k, uint64_t* mmer_out,
                                   uint32_t* offset_out, cudaStream_t st) {
    const uint64_t wmask = w >= 32 ? ~0ull : ((1ull << (2 * w)) - 1ull);
    const uint32_t hs = hash_k < w ? 2 * (w - hash_k) : 0;  // the hash keeps only the first hash_k bases (hash.rs:69)
    const uint64_t hmask64 = wmask & ~((1ull << hs) - 1ull);
    const uint32_t m = k - w + 1;  // 1 .. 32; width >= 14 -> m <= 19, width >= 17 -> m <= 16
    if (w <= 13) return launch_mw_cls<0>(m, in, n, k, w, (uint32_t)hmask64 << 6, hmask64, mmer_out, offset_out, st);
    if (w == 15) return hs ? launch_mw_cls<5>(m, in, n, k, w, (uint32_t)hmask64 << 2, hmask64, mmer_out, offset_out, st)
                           : launch_mw_cls<3>(m, in, n, k, w, (uint32_t)hmask64 << 2, hmask64, mmer_out, offset_out, st);
    if (w == 14) return hs ? launch_mw_cls<6>(m, in, n, k, w, (uint32_t)hmask64 << 4, hmask64, mmer_out, offset_out, st)
                           : launch_mw_cls<4>(m, in, n, k, w, (uint32_t)hmask64 << 4, hmask64, mmer_out, offset_out, st);
    if (w <= 16) return launch_mw_cls<1>(m, in, n, k, w, (uint32_t)hmask64, hmask64, mmer_out, offset_out, st);
    return launch_mw_cls<2>(m, in, n, k, w, 0u, hmask64, mmer_out, offset_out, st);
}

}  // namespace kmb
