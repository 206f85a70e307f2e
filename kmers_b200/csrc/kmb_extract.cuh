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
// (pair-reversing fw gives ~rc, hash.rs:62-68 vs kmer.rs:125-133), computed as
// fw ^ rc ^ canonical ^ (cmask & mask) so one predicate serves both outputs.
// Each lane stores 2 x 32 B per output array: full sectors, no read-for-ownership.
#pragma once
#include "kmb_device.cuh"

namespace kmb {

#ifndef KMB_EXTRACT_THREADS
#define KMB_EXTRACT_THREADS 256
#endif
#ifndef KMB_ITEMS_PER_CTA
#define KMB_ITEMS_PER_CTA 1024
#endif
constexpr int kExtractThreads = KMB_EXTRACT_THREADS;
constexpr int kItemsPerCta = KMB_ITEMS_PER_CTA;  // default 1024 items = 8192 windows, 128 KiB of output per CTA
constexpr int kStageBatch = 3;      // 16-byte loads a thread keeps in flight while staging

// What every window needs besides its span.
struct WinConst {
    uint32_t K;
    uint32_t shiftD;            // 2 * (48 - (kRun + K - 1))
    uint32_t mask_lo, mask_hi;  // low 2K bits
    uint32_t cm_lo, cm_hi;      // complement constant & mask (LexHash fold)
    uint32_t cmask;             // complement constant replicated over 16 fields
    uint32_t kmask;             // low K bits (window validity)
};

struct OutPtrs {
    uint64_t* canon;
    uint64_t* hash;
    uint64_t* fw;
    uint64_t* rc;
    unsigned long long* digest;  // n_valid, checksum_canon, checksum_hash
    unsigned long long* hist;    // fused histogram mode
    uint32_t hist_shift;         // 2K - hist_bits
    uint32_t vec_ok;             // output pointers are 32-byte aligned
};

struct ExtractParams {
    const uint8_t* bases;   // flat read stream
    uint64_t n_bytes;       // n_reads * L
    uint64_t L;             // read length
    uint64_t W;             // windows (= output slots) per read = L - K + 1
    uint64_t total_slots;   // n_reads * W
    uint64_t w_magic64;     // floor(2^64 / W) + 1 (W >= 2), 0 = divide
    uint32_t L32;           // L mod 2^32 (only differences inside a tile are formed)
    uint32_t W32;           // min(W, 2^32 - 1)
    uint32_t w_magic;       // floor(2^32 / W) + 1, used when 1 < W < slots per CTA + W
    uint32_t items_per_cta; // work items per CTA (host-chosen so the staged stretch fits shared memory)
    WinConst wc;
    OutPtrs out;
    EncDesc enc;
};

// ---------------------------------------------------------------------------
// phase 1: stage a stretch of the read stream into shared memory as
// {packed bits, invalid mask} entries, 16 bases each.  Entry 0 starts at
// first_al (16-byte aligned, at or below the first base needed).
// ---------------------------------------------------------------------------
template <bool VALIDATE>
__device__ __forceinline__ void stage_tile(const uint8_t* bases, uint64_t n_bytes, const uint8_t* first_al,
                                           uint32_t n_entries, const EncDesc& enc, uint2* tile) {
    // CTA-uniform: does the whole stretch lie inside the batch?  (all but the edge CTAs)
    const bool inside = first_al >= bases && first_al + (size_t)n_entries * 16 <= bases + n_bytes;
    if (inside) {
        const uint4* src = reinterpret_cast<const uint4*>(first_al);
        for (uint32_t v0 = threadIdx.x; v0 < n_entries; v0 += kStageBatch * blockDim.x) {
            uint4 raw[kStageBatch];
#pragma unroll
            for (int b = 0; b < kStageBatch; ++b) {  // all loads first: kStageBatch requests in flight per thread
                const uint32_t v = v0 + b * blockDim.x;
                if (v < n_entries) raw[b] = ld_stream_v4(src + v);
            }
#pragma unroll
            for (int b = 0; b < kStageBatch; ++b) {
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

// ---------------------------------------------------------------------------
// phase 2 building blocks
// ---------------------------------------------------------------------------
struct Span {
    uint32_t a0, a1, a2;  // forward span: 48 bases from the item's first base
    uint32_t d0, d1, d2;  // reverse complement of its first kRun+K-1 bases, at bit 0
    uint64_t inv;         // invalid-base bits of the span (bit i = base i); 0 when none
};

template <bool VALIDATE>
__device__ __forceinline__ Span load_span(const uint2* tile, uint32_t rel, const WinConst& wc) {
    const uint32_t e = rel >> 4, o2 = (rel & 15u) * 2;
    const uint2 t0 = tile[e], t1 = tile[e + 1], t2 = tile[e + 2], t3 = tile[e + 3];
    Span s;
    s.a0 = __funnelshift_r(t0.x, t1.x, o2);
    s.a1 = __funnelshift_r(t1.x, t2.x, o2);
    s.a2 = __funnelshift_r(t2.x, t3.x, o2);
    shr96(s.d0, s.d1, s.d2, pair_reverse32(s.a2 ^ wc.cmask), pair_reverse32(s.a1 ^ wc.cmask),
          pair_reverse32(s.a0 ^ wc.cmask), wc.shiftD);
    s.inv = 0;
    if (VALIDATE) {
        if ((t0.y | t1.y | t2.y | t3.y) != 0u) {
            const uint64_t m = (uint64_t)t0.y | ((uint64_t)t1.y << 16) | ((uint64_t)t2.y << 32) | ((uint64_t)t3.y << 48);
            s.inv = (m >> (o2 >> 1)) & ((1ull << (kRun + wc.K - 1)) - 1ull);  // only the bases this item's windows cover
        }
    }
    return s;
}

struct Window {
    uint64_t fw, rc, canon, hash;
};

// KHI: K > 16 (two 32-bit halves live); else everything fits the low half.
template <bool KHI>
__device__ __forceinline__ Window make_window(const Span& s, int j, const WinConst& wc) {
    Window w;
    if (KHI) {
        const uint32_t flo = __funnelshift_r(s.a0, s.a1, 2 * j);
        const uint32_t fhi = __funnelshift_r(s.a1, s.a2, 2 * j) & wc.mask_hi;
        const uint32_t rlo = __funnelshift_r(s.d0, s.d1, 2 * (kRun - 1 - j));
        const uint32_t rhi = __funnelshift_r(s.d1, s.d2, 2 * (kRun - 1 - j)) & wc.mask_hi;
        w.fw = mk64(flo, fhi);
        w.rc = mk64(rlo, rhi);
        const bool fw_less = w.fw < w.rc;  // canonical_kmer.rs:114, strict '<'
        const uint32_t clo = fw_less ? flo : rlo, chi = fw_less ? fhi : rhi;
        w.canon = mk64(clo, chi);
        // the other strand, complemented and masked = pair reversal of the canonical strand (hash.rs:60-71)
        w.hash = mk64((flo ^ rlo ^ wc.cm_lo) ^ clo, (fhi ^ rhi ^ wc.cm_hi) ^ chi);
    } else {
        const uint32_t flo = __funnelshift_r(s.a0, s.a1, 2 * j) & wc.mask_lo;
        const uint32_t rlo = __funnelshift_r(s.d0, s.d1, 2 * (kRun - 1 - j)) & wc.mask_lo;
        w.fw = flo;
        w.rc = rlo;
        const uint32_t clo = flo < rlo ? flo : rlo;
        w.canon = clo;
        w.hash = (flo ^ rlo ^ wc.cm_lo) ^ clo;
    }
    return w;
}

struct Acc {
    uint64_t canon = 0, hash = 0;
    uint32_t valid = 0;
};

// The kRun windows of one work item = kRun consecutive, 64-byte-aligned output slots.
// TWO: the item straddles a read boundary: windows j < n_first come from span A (the tail of
//      one read), the rest from span B (the head of the next; B is loaded n_first bases early so
//      the same index j addresses it).
// CHECK: some base of a span is invalid -> per-window validity + sentinel.
// nwin: slots of this item that exist (kRun except at the very end of the batch).
template <bool TWO, bool CHECK, bool DIGEST, bool FWRC, int MODE, bool KHI>
__device__ __forceinline__ void emit_run(const Span& A, const Span& B, uint32_t n_first, const WinConst& wc,
                                         const OutPtrs& o, uint64_t slot0, uint32_t nwin, Acc& acc) {
    uint64_t oc[kRun], oh[kRun], ofw[FWRC ? kRun : 1], orc[FWRC ? kRun : 1];
#pragma unroll
    for (int j = 0; j < kRun; ++j) {
        Span s = A;
        if (TWO && (uint32_t)j >= n_first) s = B;
        Window w = make_window<KHI>(s, j, wc);
        bool ok = true;
        if (CHECK) ok = (((uint32_t)(s.inv >> j)) & wc.kmask) == 0u;
        if (DIGEST || MODE == 1) {
            const bool counted = ok && (uint32_t)j < nwin;
            if (DIGEST && counted) { acc.canon += w.canon; acc.hash += w.hash; acc.valid += 1; }
            if (MODE == 1 && counted) atomicAdd(o.hist + (w.hash >> o.hist_shift), 1ull);
        }
        if (MODE == 0) {
            oc[j] = (CHECK && !ok) ? ~0ull : w.canon;
            oh[j] = (CHECK && !ok) ? ~0ull : w.hash;
            if (FWRC) { ofw[j] = (CHECK && !ok) ? ~0ull : w.fw; orc[j] = (CHECK && !ok) ? ~0ull : w.rc; }
        }
    }
    if (MODE != 0) return;
    if (nwin == kRun && o.vec_ok) {
        if (o.canon) {
            st_stream_v4u64(o.canon + slot0, oc[0], oc[1], oc[2], oc[3]);
            st_stream_v4u64(o.canon + slot0 + 4, oc[4], oc[5], oc[6], oc[7]);
        }
        if (o.hash) {
            st_stream_v4u64(o.hash + slot0, oh[0], oh[1], oh[2], oh[3]);
            st_stream_v4u64(o.hash + slot0 + 4, oh[4], oh[5], oh[6], oh[7]);
        }
        if (FWRC) {
            if (o.fw) {
                st_stream_v4u64(o.fw + slot0, ofw[0], ofw[1], ofw[2], ofw[3]);
                st_stream_v4u64(o.fw + slot0 + 4, ofw[4], ofw[5], ofw[6], ofw[7]);
            }
            if (o.rc) {
                st_stream_v4u64(o.rc + slot0, orc[0], orc[1], orc[2], orc[3]);
                st_stream_v4u64(o.rc + slot0 + 4, orc[4], orc[5], orc[6], orc[7]);
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < kRun; ++j) {
            if ((uint32_t)j < nwin) {
                if (o.canon) st_stream_u64(o.canon + slot0 + j, oc[j]);
                if (o.hash) st_stream_u64(o.hash + slot0 + j, oh[j]);
                if (FWRC) {
                    if (o.fw) st_stream_u64(o.fw + slot0 + j, ofw[j]);
                    if (o.rc) st_stream_u64(o.rc + slot0 + j, orc[j]);
                }
            }
        }
    }
}

// One window on its own (reads with fewer than kRun windows: an item then spans several reads).
template <bool VALIDATE, bool DIGEST, bool FWRC, int MODE, bool KHI>
__device__ __forceinline__ void emit_single(const uint2* tile, uint32_t rel, const WinConst& wc, const OutPtrs& o,
                                            uint64_t slot, Acc& acc) {
    const Span s = load_span<VALIDATE>(tile, rel, wc);
    const Window w = make_window<KHI>(s, 0, wc);
    const bool ok = !VALIDATE || (((uint32_t)s.inv) & wc.kmask) == 0u;
    if (DIGEST && ok) { acc.canon += w.canon; acc.hash += w.hash; acc.valid += 1; }
    if (MODE == 1) {
        if (ok) atomicAdd(o.hist + (w.hash >> o.hist_shift), 1ull);
        return;
    }
    if (o.canon) st_stream_u64(o.canon + slot, ok ? w.canon : ~0ull);
    if (o.hash) st_stream_u64(o.hash + slot, ok ? w.hash : ~0ull);
    if (FWRC) {
        if (o.fw) st_stream_u64(o.fw + slot, ok ? w.fw : ~0ull);
        if (o.rc) st_stream_u64(o.rc + slot, ok ? w.rc : ~0ull);
    }
}

template <bool DIGEST>
__device__ __forceinline__ void reduce_digest(unsigned long long (&red)[3][kExtractThreads / 32],
                                              unsigned long long* digest, const Acc& acc) {
    if (!DIGEST) return;
    const uint64_t v = warp_sum64(acc.valid), c = warp_sum64(acc.canon), h = warp_sum64(acc.hash);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = v; red[1][warp] = c; red[2][warp] = h; }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long s = 0;
        for (int w = 0; w < kExtractThreads / 32; ++w) s += red[threadIdx.x][w];
        atomicAdd(digest + threadIdx.x, s);
    }
}

// ---------------------------------------------------------------------------
// fixed-length reads.  MODE: 0 = materialise, 1 = fused histogram
//
// Work is cut in OUTPUT-slot space: item i = slots [8i, 8i+8), so every item's
// stores are 64-byte aligned whatever W = L-K+1 is.  Slot s belongs to read
// s / W at position s % W.  An item lies inside one read (fast path), or
// straddles one read boundary (two spans), or -- only when W < 8 -- several.
// ---------------------------------------------------------------------------
// u / W for a small u (u < W + slots per CTA): 0/1 when W is large, else multiply-high
__device__ __forceinline__ uint32_t div_w(uint32_t u, const ExtractParams& p, uint32_t slots_per_cta) {
    if (p.W32 >= slots_per_cta) return (u >= p.W32) ? 1u : 0u;  // u < W + slots_per_cta <= 2W
    if (p.W32 == 1) return u;
    return __umulhi(u, p.w_magic);
}

template <bool VALIDATE, bool DIGEST, bool FWRC, int MODE, bool KHI>
__global__ void __launch_bounds__(kExtractThreads) extract_fixed_kernel(const ExtractParams p) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][kExtractThreads / 32];

    const uint32_t slots_per_cta = p.items_per_cta * kRun;
    const uint64_t slot_base = (uint64_t)blockIdx.x * slots_per_cta;
    const uint32_t n_slots = (uint32_t)min((uint64_t)slots_per_cta, p.total_slots - slot_base);
    uint64_t r_first;
    if (p.W == 1) r_first = slot_base;
    else if (p.w_magic64) r_first = div_magic64(slot_base, p.w_magic64);
    else r_first = slot_base / p.W;
    const uint32_t p_first = (uint32_t)(slot_base - r_first * p.W);  // position of the CTA's first window in its read

    // ---- phase 1: pack the stretch of the flat stream that holds the CTA's windows
    const uint64_t g_start = r_first * p.L + p_first;
    const uint32_t u_last = p_first + n_slots - 1;
    const uint32_t q_last = div_w(u_last, p, slots_per_cta);
    // bases from the first window's first base to the last window's last base (mod 2^32 exact: small)
    const uint32_t span = q_last * p.L32 + (u_last - q_last * p.W32) - p_first + p.wc.K;
    const uint8_t* first = p.bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    const uint32_t n_entries = ((span + mis + 15) >> 4) + 3;  // +3: a span reads 4 entries
    stage_tile<VALIDATE>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    __syncthreads();

    // ---- phase 2
    Acc acc;
    const uint32_t n_items = (n_slots + kRun - 1) / kRun;
    for (uint32_t li = threadIdx.x; li < n_items; li += kExtractThreads) {
        const uint32_t u = p_first + li * kRun;            // first slot, counted from window 0 of read r_first
        const uint32_t q = div_w(u, p, slots_per_cta);      // reads crossed since r_first
        const uint32_t pos = u - q * p.W32;                  // window position inside its read
        const uint64_t slot0 = slot_base + (uint64_t)li * kRun;
        const uint32_t nwin = min((uint32_t)kRun, n_slots - li * kRun);
        const uint32_t rel = q * p.L32 + pos - p_first + mis;  // first base, relative to tile entry 0
        const uint32_t left = p.W32 - pos;                    // windows left in this read (>= 1)
        if (left >= (uint32_t)kRun || left >= nwin) {
            const Span s = load_span<VALIDATE>(tile, rel, p.wc);
            if (VALIDATE && s.inv != 0ull) emit_run<false, true, DIGEST, FWRC, MODE, KHI>(s, s, kRun, p.wc, p.out, slot0, nwin, acc);
            else emit_run<false, false, DIGEST, FWRC, MODE, KHI>(s, s, kRun, p.wc, p.out, slot0, nwin, acc);
        } else if (p.W32 >= (uint32_t)kRun) {
            // straddles exactly one boundary: windows j >= left start read q+1 at position j - left
            const Span a = load_span<VALIDATE>(tile, rel, p.wc);
            const Span b = load_span<VALIDATE>(tile, (q + 1) * p.L32 - p_first + mis - left, p.wc);
            if (VALIDATE && (a.inv | b.inv) != 0ull) emit_run<true, true, DIGEST, FWRC, MODE, KHI>(a, b, left, p.wc, p.out, slot0, nwin, acc);
            else emit_run<true, false, DIGEST, FWRC, MODE, KHI>(a, b, left, p.wc, p.out, slot0, nwin, acc);
        } else {
            // reads with fewer than kRun windows: window by window
            for (uint32_t j = 0; j < nwin; ++j) {
                const uint32_t uj = u + j, qj = div_w(uj, p, slots_per_cta);
                emit_single<VALIDATE, DIGEST, FWRC, MODE, KHI>(tile, qj * p.L32 + (uj - qj * p.W32) - p_first + mis, p.wc, p.out,
                                                               slot0 + j, acc);
            }
        }
    }
    reduce_digest<DIGEST>(red, p.out.digest, acc);
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
    WinConst wc;
    OutPtrs out;
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

template <bool VALIDATE, bool DIGEST, bool FWRC, int MODE, bool KHI>
__global__ void __launch_bounds__(kExtractThreads) extract_csr_kernel(const CsrParams p) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][kExtractThreads / 32];
    __shared__ uint64_t s_rlo, s_rhi;

    const uint64_t g_start = (uint64_t)blockIdx.x * kCsrTileBases;
    const uint64_t g_stop = min(g_start + (uint64_t)kCsrTileBases, p.n_bytes);  // windows start in [g_start, g_stop)
    const uint64_t g_end = min(g_stop + p.wc.K - 1, p.n_bytes);
    const uint8_t* first = p.bases + g_start;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u);
    const uint32_t n_entries = (uint32_t)((g_end - g_start + mis + 15) >> 4) + 3;
    stage_tile<VALIDATE>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    if (threadIdx.x == 0) {
        s_rlo = find_read(p.offsets, 0, p.n_reads - 1, g_start);
        s_rhi = find_read(p.offsets, s_rlo, p.n_reads - 1, g_stop - 1);
    }
    __syncthreads();

    Acc acc;
    const OutPtrs& o = p.out;
    const uint32_t n_items = (uint32_t)((g_stop - g_start + kRun - 1) / kRun);
    for (uint32_t li = threadIdx.x; li < n_items; li += kExtractThreads) {
        const uint64_t g0 = g_start + (uint64_t)li * kRun;
        const Span s = load_span<VALIDATE>(tile, li * kRun + mis, p.wc);
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
            if (g + p.wc.K > r_end) continue;  // no window starts here
            const Window w = make_window<KHI>(s, j, p.wc);
            bool ok = true;
            if (VALIDATE) ok = (((uint32_t)(s.inv >> j)) & p.wc.kmask) == 0u;
            if (DIGEST && ok) { acc.canon += w.canon; acc.hash += w.hash; acc.valid += 1; }
            if (MODE == 1) {
                if (ok) atomicAdd(o.hist + (w.hash >> o.hist_shift), 1ull);
            } else {
                const uint64_t slot = w_off + (g - r_beg);
                if (o.canon) st_stream_u64(o.canon + slot, ok ? w.canon : ~0ull);
                if (o.hash) st_stream_u64(o.hash + slot, ok ? w.hash : ~0ull);
                if (FWRC) {
                    if (o.fw) st_stream_u64(o.fw + slot, ok ? w.fw : ~0ull);
                    if (o.rc) st_stream_u64(o.rc + slot, ok ? w.rc : ~0ull);
                }
            }
        }
    }
    reduce_digest<DIGEST>(red, o.digest, acc);
}

}  // namespace kmb
