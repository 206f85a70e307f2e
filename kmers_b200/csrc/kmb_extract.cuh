// kmb_extract.cuh -- the hot path: batched CanonicalKmerIterator for K <= 32.
//
// Replaces, in batched form (file:line under /root/reference/src):
//   naive_impl/canonical_kmer_iterator.rs:42-101  which windows exist, in what order
//   naive_impl/mod.rs:40-50                        ASCII -> 2-bit, validity
//   naive_impl/kmer.rs:91-102, 124-147             rolling fw / rc words, reverse complement
//   naive_impl/canonical_kmer.rs:113-119           canonical = min(fw, rc)
//   naive_impl/hash.rs:60-71                       LexHasher
//
// Formulation: an emitted window's fw/rc words depend only on
// its own K bytes, so no rolling state is carried.  A CTA packs its stretch of
// the flat read stream once into shared memory (2 bits/base + 1 invalid
// bit/base); a work item is one thread x kRun(=8) consecutive windows of one
// read: it pulls 4 packed words, and every window is two funnel-shift extracts
// (forward strand, and the reverse-complemented span computed once per item).
//   canonical = min(fw, rc);   LexHash(canonical) = ~max(fw, rc) & mask
// (pair-reversing fw gives ~rc, hash.rs:62-68 vs kmer.rs:125-133).
// Each lane stores 2 x 32 B per output array: full sectors, no read-for-ownership.
#pragma once
#include "kmb_device.cuh"

namespace kmb {

constexpr int kExtractThreads = 256;
constexpr int kItemsPerCta = 1024;  // 8192 windows, 128 KiB of output per CTA

struct ExtractParams {
    const uint8_t* bases;   // flat read stream
    uint64_t n_bytes;       // n_reads * L
    uint32_t L32;           // L mod 2^32 (only differences inside a tile are formed)
    uint64_t W;             // windows per read = L - K + 1
    uint32_t rpr;           // work items (runs of kRun windows) per read = ceil(W / kRun)
    uint32_t rpr_magic;     // floor(2^32 / rpr) + 1, used when rpr < kItemsPerCta
    uint64_t L;             // read length
    uint64_t total_items;   // n_reads * rpr
    uint32_t K;
    uint32_t shiftD;        // 2 * (48 - (kRun + K - 1))
    uint32_t mask_lo, mask_hi;  // low 2K bits
    uint32_t vec_ok;        // output pointers are 32-byte aligned
    uint64_t* canon;
    uint64_t* hash;
    uint64_t* fw;
    uint64_t* rc;
    unsigned long long* digest;  // n_valid, checksum_canon, checksum_hash
    unsigned long long* hist;    // fused histogram mode
    uint32_t hist_shift;         // 2K - hist_bits
    EncDesc enc;
};

// One 64-bit window -> the four output words.
struct WindowOut {
    uint64_t canon, hash, fw, rc;
    bool valid;
};

// Stage the CTA's stretch of the read stream into shared memory as
// {packed bits, invalid mask} entries, 16 bases each.  Entry 0 starts at the
// 16-byte aligned address at or below `first`.
template <bool VALIDATE>
__device__ __forceinline__ void stage_tile(const uint8_t* bases, uint64_t n_bytes, const uint8_t* first_al,
                                           uint32_t n_entries, const EncDesc& enc, uint2* tile) {
    for (uint32_t v = threadIdx.x; v < n_entries; v += blockDim.x) {
        uint4 raw = load16_guarded(bases, n_bytes, first_al + (size_t)v * 16);
        PackedWord pw = pack16<VALIDATE>(raw);
        tile[v] = make_uint2(apply_encoding(pw.bits, enc), pw.inv);
    }
}

// MODE: 0 = materialise (canon/hash [+fw/rc]), 1 = fused histogram (nothing materialised)
template <bool VALIDATE, bool DIGEST, bool FWRC, int MODE>
__global__ void __launch_bounds__(kExtractThreads) extract_fixed_kernel(const ExtractParams p) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][kExtractThreads / 32];

    const uint64_t item0 = (uint64_t)blockIdx.x * kItemsPerCta;
    const uint32_t n_items = (uint32_t)min((uint64_t)kItemsPerCta, p.total_items - item0);
    const uint64_t r_first = item0 / p.rpr;
    const uint32_t run_first = (uint32_t)(item0 - r_first * p.rpr);

    // ---- phase 1: pack the stretch [g_start, g_end) of the flat stream
    const uint64_t g_start = r_first * p.L + (uint64_t)run_first * kRun;
    const uint64_t last = item0 + n_items - 1;
    const uint64_t r_last = last / p.rpr;
    const uint32_t run_last = (uint32_t)(last - r_last * p.rpr);
    uint64_t g_end = r_last * p.L + (uint64_t)run_last * kRun + kRun + p.K - 1;
    if (g_end > p.n_bytes) g_end = p.n_bytes;
    const uint8_t* first = p.bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    const uint32_t n_entries = (uint32_t)((g_end - g_start + mis + 15) >> 4) + 3;  // +3: items read 4 entries
    stage_tile<VALIDATE>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    __syncthreads();

    // ---- phase 2: one item = kRun windows of one read
    uint64_t acc_canon = 0, acc_hash = 0;
    uint32_t acc_valid = 0;
    const uint32_t cmask = p.enc.cmask;
    const uint32_t kmask = (p.K >= 32) ? 0xFFFFFFFFu : ((1u << p.K) - 1u);

    for (uint32_t li = threadIdx.x; li < n_items; li += kExtractThreads) {
        const uint32_t gi = run_first + li;
        uint32_t q;  // reads crossed since r_first
        if (p.rpr >= (uint32_t)kItemsPerCta) q = (gi >= p.rpr) ? 1u : 0u;
        else if (p.rpr == 1) q = gi;
        else q = __umulhi(gi, p.rpr_magic);
        const uint32_t run = gi - q * p.rpr;
        const uint32_t p0 = run * kRun;
        const uint64_t r = r_first + q;
        const uint64_t slot0 = r * p.W + p0;
        const uint32_t nwin = (uint32_t)min((uint64_t)kRun, p.W - p0);
        // position of the item's first base relative to tile entry 0 (mod 2^32 exact)
        const uint32_t rel = q * p.L32 + p0 - run_first * kRun + mis;
        const uint32_t e = rel >> 4, o2 = (rel & 15u) * 2;

        const uint2 t0 = tile[e], t1 = tile[e + 1], t2 = tile[e + 2], t3 = tile[e + 3];
        // forward span, 48 bases from the item's first base
        const uint32_t a0 = __funnelshift_r(t0.x, t1.x, o2);
        const uint32_t a1 = __funnelshift_r(t1.x, t2.x, o2);
        const uint32_t a2 = __funnelshift_r(t2.x, t3.x, o2);
        // reverse complement of the first kRun+K-1 bases of the span, at bit 0
        uint32_t d0, d1, d2;
        shr96(d0, d1, d2, pair_reverse32(a2 ^ cmask), pair_reverse32(a1 ^ cmask), pair_reverse32(a0 ^ cmask),
              p.shiftD);
        // invalid-base bits of the span (bit i = base i of the span)
        uint64_t inv = 0;
        if (VALIDATE) {
            if ((t0.y | t1.y | t2.y | t3.y) != 0u) {
                uint64_t m = (uint64_t)t0.y | ((uint64_t)t1.y << 16) | ((uint64_t)t2.y << 32) | ((uint64_t)t3.y << 48);
                inv = m >> (o2 >> 1);
            }
        }

        uint64_t oc[kRun], oh[kRun], ofw[FWRC ? kRun : 1], orc[FWRC ? kRun : 1];
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            const uint32_t flo = __funnelshift_r(a0, a1, 2 * j) & p.mask_lo;
            const uint32_t fhi = __funnelshift_r(a1, a2, 2 * j) & p.mask_hi;
            const uint32_t rlo = __funnelshift_r(d0, d1, 2 * (kRun - 1 - j)) & p.mask_lo;
            const uint32_t rhi = __funnelshift_r(d1, d2, 2 * (kRun - 1 - j)) & p.mask_hi;
            const uint64_t fw = mk64(flo, fhi), rc = mk64(rlo, rhi);
            const bool fw_less = fw < rc;  // canonical_kmer.rs:114, strict '<'
            uint64_t canon = fw_less ? fw : rc;
            const uint64_t other = fw_less ? rc : fw;
            uint64_t h = (other ^ mk64(cmask, cmask)) & mk64(p.mask_lo, p.mask_hi);
            bool ok = true;
            if (VALIDATE) ok = (((uint32_t)(inv >> j)) & kmask) == 0u;
            if (DIGEST || MODE == 1) {
                const bool counted = ok && (uint32_t)j < nwin;
                if (DIGEST && counted) { acc_canon += canon; acc_hash += h; acc_valid += 1; }
                if (MODE == 1 && counted) atomicAdd(p.hist + (h >> p.hist_shift), 1ull);
            }
            if (MODE == 0) {
                if (VALIDATE && !ok) { canon = ~0ull; h = ~0ull; }
                oc[j] = canon; oh[j] = h;
                if (FWRC) { ofw[j] = ok ? fw : ~0ull; orc[j] = ok ? rc : ~0ull; }
            }
        }

        if (MODE == 0) {
            if (nwin == kRun && p.vec_ok && (slot0 & 3ull) == 0ull) {
                if (p.canon) {
                    st_stream_v4u64(p.canon + slot0, oc[0], oc[1], oc[2], oc[3]);
                    st_stream_v4u64(p.canon + slot0 + 4, oc[4], oc[5], oc[6], oc[7]);
                }
                if (p.hash) {
                    st_stream_v4u64(p.hash + slot0, oh[0], oh[1], oh[2], oh[3]);
                    st_stream_v4u64(p.hash + slot0 + 4, oh[4], oh[5], oh[6], oh[7]);
                }
                if (FWRC) {
                    if (p.fw) {
                        st_stream_v4u64(p.fw + slot0, ofw[0], ofw[1], ofw[2], ofw[3]);
                        st_stream_v4u64(p.fw + slot0 + 4, ofw[4], ofw[5], ofw[6], ofw[7]);
                    }
                    if (p.rc) {
                        st_stream_v4u64(p.rc + slot0, orc[0], orc[1], orc[2], orc[3]);
                        st_stream_v4u64(p.rc + slot0 + 4, orc[4], orc[5], orc[6], orc[7]);
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < kRun; ++j) {
                    if ((uint32_t)j < nwin) {
                        if (p.canon) st_stream_u64(p.canon + slot0 + j, oc[j]);
                        if (p.hash) st_stream_u64(p.hash + slot0 + j, oh[j]);
                        if (FWRC) {
                            if (p.fw) st_stream_u64(p.fw + slot0 + j, ofw[j]);
                            if (p.rc) st_stream_u64(p.rc + slot0 + j, orc[j]);
                        }
                    }
                }
            }
        }
    }

    if (DIGEST) {
        uint64_t v = warp_sum64(acc_valid), c = warp_sum64(acc_canon), h = warp_sum64(acc_hash);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { red[0][warp] = v; red[1][warp] = c; red[2][warp] = h; }
        __syncthreads();
        if (threadIdx.x < 3) {
            unsigned long long s = 0;
            for (int w = 0; w < kExtractThreads / 32; ++w) s += red[threadIdx.x][w];
            atomicAdd(p.digest + threadIdx.x, s);
        }
    }
}

// ---------------------------------------------------------------------------
// Ragged (CSR) batches: tiles are fixed stretches of the flat stream; an item
// is kRun consecutive flat positions and every window checks its own read.
// Slower stores (8 B each) -- this is the generality path, not the headline.
// ---------------------------------------------------------------------------
constexpr int kCsrTileBases = 4096;

struct CsrParams {
    const uint8_t* bases;
    uint64_t n_bytes;
    const uint64_t* offsets;      // n_reads + 1
    const uint64_t* win_offsets;  // n_reads + 1 exclusive prefix of window counts
    uint64_t n_reads;
    uint32_t K;
    uint32_t shiftD;
    uint32_t mask_lo, mask_hi;
    uint64_t* canon;
    uint64_t* hash;
    uint64_t* fw;
    uint64_t* rc;
    unsigned long long* digest;
    unsigned long long* hist;
    uint32_t hist_shift;
    EncDesc enc;
};

// largest r in [lo, hi] with offsets[r] <= g   (offsets[lo] <= g guaranteed)
__device__ __forceinline__ uint64_t find_read(const uint64_t* offsets, uint64_t lo, uint64_t hi, uint64_t g) {
    while (lo < hi) {
        uint64_t mid = lo + ((hi - lo + 1) >> 1);
        if (__ldg(offsets + mid) <= g) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <bool VALIDATE, bool DIGEST, bool FWRC, int MODE>
__global__ void __launch_bounds__(kExtractThreads) extract_csr_kernel(const CsrParams p) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][kExtractThreads / 32];
    __shared__ uint64_t s_rlo, s_rhi;

    const uint64_t g_start = (uint64_t)blockIdx.x * kCsrTileBases;
    const uint64_t g_stop = min(g_start + (uint64_t)kCsrTileBases, p.n_bytes);  // windows start in [g_start, g_stop)
    const uint64_t g_end = min(g_stop + p.K - 1, p.n_bytes);
    const uint8_t* first = p.bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    const uint32_t n_entries = (uint32_t)((g_end - g_start + mis + 15) >> 4) + 3;
    stage_tile<VALIDATE>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    if (threadIdx.x == 0) {
        s_rlo = find_read(p.offsets, 0, p.n_reads - 1, g_start);
        s_rhi = find_read(p.offsets, s_rlo, p.n_reads - 1, g_stop - 1);
    }
    __syncthreads();

    uint64_t acc_canon = 0, acc_hash = 0;
    uint32_t acc_valid = 0;
    const uint32_t cmask = p.enc.cmask;
    const uint32_t kmask = (p.K >= 32) ? 0xFFFFFFFFu : ((1u << p.K) - 1u);
    const uint32_t n_items = (uint32_t)((g_stop - g_start + kRun - 1) / kRun);

    for (uint32_t li = threadIdx.x; li < n_items; li += kExtractThreads) {
        const uint64_t g0 = g_start + (uint64_t)li * kRun;
        const uint32_t rel = li * kRun + mis;
        const uint32_t e = rel >> 4, o2 = (rel & 15u) * 2;
        const uint2 t0 = tile[e], t1 = tile[e + 1], t2 = tile[e + 2], t3 = tile[e + 3];
        const uint32_t a0 = __funnelshift_r(t0.x, t1.x, o2);
        const uint32_t a1 = __funnelshift_r(t1.x, t2.x, o2);
        const uint32_t a2 = __funnelshift_r(t2.x, t3.x, o2);
        uint32_t d0, d1, d2;
        shr96(d0, d1, d2, pair_reverse32(a2 ^ cmask), pair_reverse32(a1 ^ cmask), pair_reverse32(a0 ^ cmask),
              p.shiftD);
        uint64_t inv = 0;
        if (VALIDATE) {
            uint64_t m = (uint64_t)t0.y | ((uint64_t)t1.y << 16) | ((uint64_t)t2.y << 32) | ((uint64_t)t3.y << 48);
            inv = m >> (o2 >> 1);
        }
        uint64_t r = find_read(p.offsets, s_rlo, s_rhi, g0);
        uint64_t r_beg = __ldg(p.offsets + r), r_end = __ldg(p.offsets + r + 1);
        uint64_t w_off = __ldg(p.win_offsets + r);
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            const uint64_t g = g0 + j;
            if (g >= g_stop) break;
            while (g >= r_end) {  // step over read boundaries (and empty reads)
                ++r;
                r_beg = r_end;
                r_end = __ldg(p.offsets + r + 1);
                w_off = __ldg(p.win_offsets + r);
            }
            if (g + p.K > r_end) continue;  // no window starts here
            const uint32_t flo = __funnelshift_r(a0, a1, 2 * j) & p.mask_lo;
            const uint32_t fhi = __funnelshift_r(a1, a2, 2 * j) & p.mask_hi;
            const uint32_t rlo = __funnelshift_r(d0, d1, 2 * (kRun - 1 - j)) & p.mask_lo;
            const uint32_t rhi = __funnelshift_r(d1, d2, 2 * (kRun - 1 - j)) & p.mask_hi;
            const uint64_t fw = mk64(flo, fhi), rc = mk64(rlo, rhi);
            const bool fw_less = fw < rc;
            uint64_t canon = fw_less ? fw : rc;
            const uint64_t other = fw_less ? rc : fw;
            uint64_t h = (other ^ mk64(cmask, cmask)) & mk64(p.mask_lo, p.mask_hi);
            bool ok = true;
            if (VALIDATE) ok = (((uint32_t)(inv >> j)) & kmask) == 0u;
            if (DIGEST && ok) { acc_canon += canon; acc_hash += h; acc_valid += 1; }
            if (MODE == 1) {
                if (ok) atomicAdd(p.hist + (h >> p.hist_shift), 1ull);
            } else {
                const uint64_t slot = w_off + (g - r_beg);
                if (p.canon) st_stream_u64(p.canon + slot, ok ? canon : ~0ull);
                if (p.hash) st_stream_u64(p.hash + slot, ok ? h : ~0ull);
                if (FWRC) {
                    if (p.fw) st_stream_u64(p.fw + slot, ok ? fw : ~0ull);
                    if (p.rc) st_stream_u64(p.rc + slot, ok ? rc : ~0ull);
                }
            }
        }
    }

    if (DIGEST) {
        uint64_t v = warp_sum64(acc_valid), c = warp_sum64(acc_canon), h = warp_sum64(acc_hash);
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        if (lane == 0) { red[0][warp] = v; red[1][warp] = c; red[2][warp] = h; }
        __syncthreads();
        if (threadIdx.x < 3) {
            unsigned long long s = 0;
            for (int w = 0; w < kExtractThreads / 32; ++w) s += red[threadIdx.x][w];
            atomicAdd(p.digest + threadIdx.x, s);
        }
    }
}

}  // namespace kmb
