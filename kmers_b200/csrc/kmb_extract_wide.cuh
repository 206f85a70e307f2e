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
// Same tiling as kmb_extract.cuh, with a 5-word (80-base) span per item.
#pragma once
#include "kmb_extract.cuh"

namespace kmb {

constexpr int kWideA = 5;  // 32-bit words of forward span: kRun + 64 - 1 = 71 bases <= 80

struct WideParams {
    // geometry (fixed-length batches use the first block, CSR the second)
    const uint8_t* bases;
    uint64_t n_bytes;
    uint64_t L, W;
    uint32_t L32, rpr, rpr_magic;
    uint64_t total_items;
    const uint64_t* offsets;
    const uint64_t* win_offsets;
    uint64_t n_reads;
    uint32_t K;
    uint32_t shiftD;     // 2 * (80 - (kRun + K - 1))
    uint32_t mask[4];    // low 2K bits over four 32-bit words
    uint64_t* canon;     // 2 words per slot
    uint64_t* hash;      // 2 words per slot
    unsigned long long* digest;
    EncDesc enc;
};

template <int WS>
__device__ __forceinline__ void shr_words(uint32_t (&d)[kWideA + 1], const uint32_t (&c)[kWideA], uint32_t s) {
#pragma unroll
    for (int i = 0; i < kWideA + 1; ++i) {
        const uint32_t lo = (i + WS) < kWideA ? c[i + WS] : 0u;
        const uint32_t hi = (i + WS + 1) < kWideA ? c[i + WS + 1] : 0u;
        d[i] = __funnelshift_r(lo, hi, s);
    }
}

struct WideSpan {
    uint32_t a[kWideA + 1];  // forward span (a[kWideA] = 0 pad)
    uint32_t d[kWideA + 1];  // reverse complement of the first kRun+K-1 bases, at bit 0
    uint64_t inv_lo, inv_hi; // invalid-base bits of the span
};

template <bool VALIDATE>
__device__ __forceinline__ void load_wide_span(const uint2* tile, uint32_t rel, uint32_t cmask, uint32_t shiftD,
                                               WideSpan& s) {
    const uint32_t e = rel >> 4, o2 = (rel & 15u) * 2;
    uint2 t[kWideA + 1];
#pragma unroll
    for (int i = 0; i < kWideA + 1; ++i) t[i] = tile[e + i];
#pragma unroll
    for (int i = 0; i < kWideA; ++i) s.a[i] = __funnelshift_r(t[i].x, t[i + 1].x, o2);
    s.a[kWideA] = 0;
    uint32_t c[kWideA];
#pragma unroll
    for (int i = 0; i < kWideA; ++i) c[i] = pair_reverse32(s.a[kWideA - 1 - i] ^ cmask);
    const uint32_t sh = shiftD & 31u;
    switch (shiftD >> 5) {
        case 0: shr_words<0>(s.d, c, sh); break;
        case 1: shr_words<1>(s.d, c, sh); break;
        case 2: shr_words<2>(s.d, c, sh); break;
        case 3: shr_words<3>(s.d, c, sh); break;
        default: shr_words<4>(s.d, c, sh); break;
    }
    s.inv_lo = 0; s.inv_hi = 0;
    if (VALIDATE) {
        const uint64_t mlo = (uint64_t)t[0].y | ((uint64_t)t[1].y << 16) | ((uint64_t)t[2].y << 32) | ((uint64_t)t[3].y << 48);
        const uint64_t mhi = (uint64_t)t[4].y | ((uint64_t)t[5].y << 16);
        const uint32_t o = o2 >> 1;
        s.inv_lo = o ? ((mlo >> o) | (mhi << (64 - o))) : mlo;
        s.inv_hi = mhi >> o;
    }
}

struct WideWindow {
    uint64_t canon[2], hash[2];
    bool ok;
};

template <bool VALIDATE>
__device__ __forceinline__ WideWindow wide_window(const WideSpan& s, int j, const WideParams& p) {
    uint32_t f[4], r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[i] = __funnelshift_r(s.a[i], s.a[i + 1], 2 * j) & p.mask[i];
        r[i] = __funnelshift_r(s.d[i], s.d[i + 1], 2 * (kRun - 1 - j)) & p.mask[i];
    }
    const uint64_t f0 = mk64(f[0], f[1]), f1 = mk64(f[2], f[3]);
    const uint64_t r0 = mk64(r[0], r[1]), r1 = mk64(r[2], r[3]);
    const bool fw_less = (f1 < r1) || (f1 == r1 && f0 < r0);
    const uint64_t cm = mk64(p.enc.cmask, p.enc.cmask);
    WideWindow w;
    w.canon[0] = fw_less ? f0 : r0;
    w.canon[1] = fw_less ? f1 : r1;
    // pair reversal of the canonical strand == complement-constant XOR of the other strand
    w.hash[0] = ((fw_less ? r0 : f0) ^ cm) & mk64(p.mask[0], p.mask[1]);
    w.hash[1] = ((fw_less ? r1 : f1) ^ cm) & mk64(p.mask[2], p.mask[3]);
    w.ok = true;
    if (VALIDATE) {
        const uint64_t x = j ? ((s.inv_lo >> j) | (s.inv_hi << (64 - j))) : s.inv_lo;
        const uint64_t km = p.K >= 64 ? ~0ull : ((1ull << p.K) - 1ull);
        w.ok = (x & km) == 0ull;
    }
    return w;
}

__device__ __forceinline__ void wide_store(uint64_t* base, uint64_t slot, const uint64_t (&v)[2]) {
    st_stream_v2u64(base + 2 * slot, v[0], v[1]);
}

template <bool VALIDATE, bool DIGEST>
__device__ __forceinline__ void wide_emit(const WideParams& p, const WideWindow& w, uint64_t slot, uint64_t& acc_c,
                                          uint64_t& acc_h, uint32_t& acc_v) {
    if (DIGEST && w.ok) { acc_c += w.canon[0] + w.canon[1]; acc_h += w.hash[0] + w.hash[1]; acc_v += 1; }
    const uint64_t ones[2] = {~0ull, ~0ull};
    if (p.canon) wide_store(p.canon, slot, w.ok ? w.canon : ones);
    if (p.hash) wide_store(p.hash, slot, w.ok ? w.hash : ones);
}

template <bool DIGEST>
__device__ __forceinline__ void wide_reduce(unsigned long long (&red)[3][kExtractThreads / 32], unsigned long long* digest,
                                            uint32_t acc_v, uint64_t acc_c, uint64_t acc_h) {
    if (!DIGEST) return;
    uint64_t v = warp_sum64(acc_v), c = warp_sum64(acc_c), h = warp_sum64(acc_h);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = v; red[1][warp] = c; red[2][warp] = h; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long s = 0;
        for (int w = 0; w < kExtractThreads / 32; ++w) s += red[threadIdx.x][w];
        atomicAdd(digest + threadIdx.x, s);
    }
}

template <bool VALIDATE, bool DIGEST>
__global__ void __launch_bounds__(kExtractThreads) extract_wide_fixed_kernel(const WideParams p) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][kExtractThreads / 32];
    const uint64_t item0 = (uint64_t)blockIdx.x * kItemsPerCta;
    const uint32_t n_items = (uint32_t)min((uint64_t)kItemsPerCta, p.total_items - item0);
    const uint64_t r_first = item0 / p.rpr;
    const uint32_t run_first = (uint32_t)(item0 - r_first * p.rpr);
    const uint64_t g_start = r_first * p.L + (uint64_t)run_first * kRun;
    const uint64_t last = item0 + n_items - 1;
    const uint64_t r_last = last / p.rpr;
    const uint32_t run_last = (uint32_t)(last - r_last * p.rpr);
    uint64_t g_end = r_last * p.L + (uint64_t)run_last * kRun + kRun + p.K - 1;
    if (g_end > p.n_bytes) g_end = p.n_bytes;
    const uint8_t* first = p.bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    const uint32_t n_entries = (uint32_t)((g_end - g_start + mis + 15) >> 4) + kWideA;
    stage_tile<VALIDATE>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    __syncthreads();

    uint64_t acc_c = 0, acc_h = 0;
    uint32_t acc_v = 0;
    for (uint32_t li = threadIdx.x; li < n_items; li += kExtractThreads) {
        const uint32_t gi = run_first + li;
        uint32_t q;
        if (p.rpr >= (uint32_t)kItemsPerCta) q = (gi >= p.rpr) ? 1u : 0u;
        else if (p.rpr == 1) q = gi;
        else q = __umulhi(gi, p.rpr_magic);
        const uint32_t run = gi - q * p.rpr;
        const uint32_t p0 = run * kRun;
        const uint64_t slot0 = (r_first + q) * p.W + p0;
        const uint32_t nwin = (uint32_t)min((uint64_t)kRun, p.W - p0);
        WideSpan s;
        load_wide_span<VALIDATE>(tile, q * p.L32 + p0 - run_first * kRun + mis, p.enc.cmask, p.shiftD, s);
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            if ((uint32_t)j < nwin) {
                WideWindow w = wide_window<VALIDATE>(s, j, p);
                wide_emit<VALIDATE, DIGEST>(p, w, slot0 + j, acc_c, acc_h, acc_v);
            }
        }
    }
    wide_reduce<DIGEST>(red, p.digest, acc_v, acc_c, acc_h);
}

template <bool VALIDATE, bool DIGEST>
__global__ void __launch_bounds__(kExtractThreads) extract_wide_csr_kernel(const WideParams p) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][kExtractThreads / 32];
    __shared__ uint64_t s_rlo, s_rhi;
    const uint64_t g_start = (uint64_t)blockIdx.x * kCsrTileBases;
    const uint64_t g_stop = min(g_start + (uint64_t)kCsrTileBases, p.n_bytes);
    const uint64_t g_end = min(g_stop + p.K - 1, p.n_bytes);
    const uint8_t* first = p.bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    const uint32_t n_entries = (uint32_t)((g_end - g_start + mis + 15) >> 4) + kWideA;
    stage_tile<VALIDATE>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    if (threadIdx.x == 0) {
        s_rlo = find_read(p.offsets, 0, p.n_reads - 1, g_start);
        s_rhi = find_read(p.offsets, s_rlo, p.n_reads - 1, g_stop - 1);
    }
    __syncthreads();

    uint64_t acc_c = 0, acc_h = 0;
    uint32_t acc_v = 0;
    const uint32_t n_items = (uint32_t)((g_stop - g_start + kRun - 1) / kRun);
    for (uint32_t li = threadIdx.x; li < n_items; li += kExtractThreads) {
        const uint64_t g0 = g_start + (uint64_t)li * kRun;
        WideSpan s;
        load_wide_span<VALIDATE>(tile, li * kRun + mis, p.enc.cmask, p.shiftD, s);
        uint64_t r = find_read(p.offsets, s_rlo, s_rhi, g0);
        uint64_t r_beg = __ldg(p.offsets + r), r_end = __ldg(p.offsets + r + 1);
        uint64_t w_off = __ldg(p.win_offsets + r);
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            const uint64_t g = g0 + j;
            if (g >= g_stop) break;
            while (g >= r_end) {
                ++r;
                r_beg = r_end;
                r_end = __ldg(p.offsets + r + 1);
                w_off = __ldg(p.win_offsets + r);
            }
            if (g + p.K > r_end) continue;
            WideWindow w = wide_window<VALIDATE>(s, j, p);
            wide_emit<VALIDATE, DIGEST>(p, w, w_off + (g - r_beg), acc_c, acc_h, acc_v);
        }
    }
    wide_reduce<DIGEST>(red, p.digest, acc_v, acc_c, acc_h);
}

}  // namespace kmb
