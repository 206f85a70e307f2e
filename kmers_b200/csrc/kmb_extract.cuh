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
#include "kmb_geometry.cuh"

namespace kmb {

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
    unsigned long long* hist;    // fused histogram mode: global u64 bins
    unsigned int* s_hist;        // MODE 2: this CTA's bins in shared memory, two 16-bit counters per word (smem_hist_add)
    uint32_t hist_shift;         // 2K - hist_bits
    uint32_t vec_ok;             // output pointers are 32-byte aligned
    uint32_t one;                // the constant 1, from the parameter bank: a multiplier ptxas cannot fold, so that
                                 // 64-bit accumulation is issued as IMAD.WIDE on the FMA pipe (see HistAcc)
    uint32_t hist_hi_shift;      // MODE 2 fast path: bin = (~max_hi & mask_hi) >> hist_hi_shift (hist_shift >= 32), else 0xFFFFFFFF
    uint32_t bin_mask;           // (1 << hist_bits) - 1
};

struct NarrowParams {
    WinConst wc;
    OutPtrs out;
};

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

// MODE 2: one count into this CTA's shared-memory histogram.  Two 16-bit counters share a 32-bit word.  The
// increment that takes a counter to 2^15 (its atomicAdd returned 0x7FFF in that half: exactly one thread per lap sees
// that) moves those 2^15 to the global bin at once.  A counter therefore stays far below 2^16 -- it would take another
// 32768 increments of the same bin between that thread's two consecutive atomics to overflow it -- so no carry ever
// crosses into the neighbouring counter and the shared counts stay exact.
// The fused-histogram kernels are bound by the integer ALU pipe, so the bookkeeping around the atomic is written to go
// through the FMA pipe where it can (IMAD with multipliers from the parameter bank): one LOP3 for the half, one for the
// word address, one for the overflow test.
__device__ __forceinline__ void smem_hist_add(const OutPtrs& o, uint32_t bin) {
    const uint32_t inc = 1u << ((bin * (16u * o.one)) & 16u);  // 1 or 0x10000: the counter's half     (IMAD, LOP3, SHF)
    const uint32_t m = inc * 0x7FFFu;                          // the counter's low 15 bits            (IMAD)
    const uint32_t off = (bin * (2u * o.one)) & ~3u;           // byte offset of the word              (IMAD, LOP3)
    unsigned int* w = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(o.s_hist) + off);
    const uint32_t old = atomicAdd(w, inc);
    if ((~old & m) == 0u) {                                     // the half held 0x7FFF before this increment
        atomicSub(w, m + inc);                                  // take 2^15 out of it ...
        atomicAdd(o.hist + bin, 32768ull);                      // ... and give them to the global bin
    }
}

// Digest of the fused-histogram kernels, kept off the ALU pipe.  With mx = max(fw, rc):
//   canonical = fw + rc - mx,   LexHash(canonical) = ~mx & mask = mask - mx      (kmb_extract.cuh header)
// so three running sums -- of fw + rc, of mx, and the window count -- give both checksums, and each 32-bit half is added
// into a 64-bit accumulator by one IMAD.WIDE (x * counted + acc: the multiplier doubles as the predicate).
struct HistAcc {
    uint64_t s_lo = 0, m_lo = 0;  // sums of the low words of fw and rc / of max(fw, rc): the carries out of bit 31 matter
    uint32_t s_hi = 0, m_hi = 0;  // sums of the high words: only their low 32 bits reach a checksum mod 2^64
    uint32_t valid = 0;
    __device__ __forceinline__ void add(uint32_t flo, uint32_t fhi, uint32_t rlo, uint32_t rhi, uint32_t mlo, uint32_t mhi, uint32_t counted) {
        s_lo += (uint64_t)flo * counted; s_lo += (uint64_t)rlo * counted;  // IMAD.WIDE
        s_hi += fhi * counted; s_hi += rhi * counted;                      // IMAD
        m_lo += (uint64_t)mlo * counted; m_hi += mhi * counted;
        valid += counted;
    }
    __device__ __forceinline__ void to(Acc& a, uint64_t mask) const {
        const uint64_t mx = m_lo + ((uint64_t)m_hi << 32);
        a.valid = valid;
        a.canon = s_lo + ((uint64_t)s_hi << 32) - mx;
        a.hash = (uint64_t)valid * mask - mx;
    }
};

// One window of the fused histogram (MODE 2): only what the bin and the digest need.  The bin is the top hist_bits of
// LexHash(canonical) = ~max(fw, rc) & mask, i.e. (for 2K - hist_bits >= 32) a field of the HIGH word of the larger strand:
// no canonical word, no hash word, no 64-bit shift.
// HIBIN: 2K - hist_bits >= 32, the bin is a field of the high word (a template parameter: as a run-time test both variants
// would be issued, predicated).  GUARD: the window may be missing or invalid (`counted`); without it every window counts.
template <bool KHI, bool DIGEST, bool HIBIN, bool GUARD>
__device__ __forceinline__ void hist_window(const Span& s, int j, const WinConst& wc, const OutPtrs& o, bool counted, HistAcc& hacc) {
    uint32_t flo = __funnelshift_r(s.a0, s.a1, 2 * j), rlo = __funnelshift_r(s.d0, s.d1, 2 * (kRun - 1 - j));
    uint32_t fhi = 0, rhi = 0;
    bool fw_less;
    if (KHI) {
        fhi = __funnelshift_r(s.a1, s.a2, 2 * j) & wc.mask_hi;
        rhi = __funnelshift_r(s.d1, s.d2, 2 * (kRun - 1 - j)) & wc.mask_hi;
        fw_less = mk64(flo, fhi) < mk64(rlo, rhi);
    } else {
        flo &= wc.mask_lo;
        rlo &= wc.mask_lo;
        fw_less = flo < rlo;
    }
    const uint32_t mx_hi = fw_less ? rhi : fhi;
    uint32_t bin;
    const uint32_t mult = GUARD ? (counted ? o.one : 0u) : o.one;
    if (KHI && HIBIN) {
        bin = (~mx_hi & wc.mask_hi) >> o.hist_hi_shift;
        if (DIGEST) hacc.add(flo, fhi, rlo, rhi, fw_less ? rlo : flo, mx_hi, mult);
    } else {
        const uint32_t mx_lo = fw_less ? rlo : flo;
        const uint64_t hash = ~mk64(mx_lo, mx_hi) & mk64(wc.mask_lo, wc.mask_hi);
        bin = (uint32_t)(hash >> o.hist_shift);
        if (DIGEST) hacc.add(flo, fhi, rlo, rhi, mx_lo, mx_hi, mult);
    }
    if (!GUARD || counted) smem_hist_add(o, bin);
}

// the kRun windows of one work item into the shared-memory histogram (MODE 2)
template <bool TWO, bool CHECK, bool DIGEST, bool KHI, bool HIBIN>
__device__ __forceinline__ void emit_run_hist(const Span& A, const Span& B, uint32_t n_first, const WinConst& wc, const OutPtrs& o,
                                              uint32_t nwin, HistAcc& hacc) {
    if (!CHECK && nwin == kRun) {  // the common case: a full item without an invalid base -- no per-window guards at all
#pragma unroll
        for (int j = 0; j < kRun; ++j) hist_window<KHI, DIGEST, HIBIN, false>((TWO && (uint32_t)j >= n_first) ? B : A, j, wc, o, true, hacc);
        return;
    }
#pragma unroll
    for (int j = 0; j < kRun; ++j) {
        const Span& s = (TWO && (uint32_t)j >= n_first) ? B : A;
        bool ok = true;
        if (CHECK) ok = (((uint32_t)(s.inv >> j)) & wc.kmask) == 0u;
        hist_window<KHI, DIGEST, HIBIN, true>(s, j, wc, o, ok && (uint32_t)j < nwin, hacc);
    }
}

// The kRun windows of one work item = kRun consecutive, 64-byte-aligned output slots.
// TWO: the item straddles a read boundary: windows j < n_first come from span A (the tail of
//      one read), the rest from span B (the head of the next; B is loaded n_first bases early so
//      the same index j addresses it).
// CHECK: some base of a span is invalid -> per-window validity + sentinel.
// nwin: slots of this item that exist (kRun except at the very end of the batch).
template <bool TWO, bool CHECK, bool DIGEST, bool FWRC, int MODE, bool KHI, bool HASH>
__device__ __forceinline__ void emit_run(const Span& A, const Span& B, uint32_t n_first, const WinConst& wc,
                                         const OutPtrs& o, uint64_t slot0, uint32_t nwin, Acc& acc) {
    uint64_t oc[kRun], oh[HASH ? kRun : 1], ofw[FWRC ? kRun : 1], orc[FWRC ? kRun : 1];  // HASH == false: hash_out is NULL
#pragma unroll
    for (int j = 0; j < kRun; ++j) {
        Span s = A;
        if (TWO && (uint32_t)j >= n_first) s = B;
        Window w = make_window<KHI>(s, j, wc);
        bool ok = true;
        if (CHECK) ok = (((uint32_t)(s.inv >> j)) & wc.kmask) == 0u;
        if (DIGEST || MODE != 0) {
            const bool counted = ok && (uint32_t)j < nwin;
            if (DIGEST && counted) { acc.canon += w.canon; acc.hash += w.hash; acc.valid += 1; }
            if (MODE == 1 && counted) atomicAdd(o.hist + (w.hash >> o.hist_shift), 1ull);
            if (MODE == 2 && counted) smem_hist_add(o, (uint32_t)(w.hash >> o.hist_shift));
        }
        if (MODE == 0) {
            oc[j] = (CHECK && !ok) ? ~0ull : w.canon;
            if (HASH) oh[j] = (CHECK && !ok) ? ~0ull : w.hash;
            if (FWRC) { ofw[j] = (CHECK && !ok) ? ~0ull : w.fw; orc[j] = (CHECK && !ok) ? ~0ull : w.rc; }
        }
    }
    if (MODE != 0) return;
    if (nwin == kRun && o.vec_ok && (slot0 & 3ull) == 0ull) {
        if (o.canon) {
            st_stream_v4u64(o.canon + slot0, oc[0], oc[1], oc[2], oc[3]);
            st_stream_v4u64(o.canon + slot0 + 4, oc[4], oc[5], oc[6], oc[7]);
        }
        if (HASH && o.hash) {
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
                if (HASH && o.hash) st_stream_u64(o.hash + slot0 + j, oh[j]);
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
__device__ __forceinline__ void reduce_digest(unsigned long long (&red)[3][32],
                                              unsigned long long* digest, const Acc& acc) {
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

// ---------------------------------------------------------------------------
// The K <= 32 engine plugged into the geometry of kmb_geometry.cuh.
// MODE: 0 = materialise, 1 = fused histogram (nothing materialised).
// ---------------------------------------------------------------------------
template <bool VALIDATE, bool DIGEST, bool FWRC, int MODE, bool KHI, bool HASH = true>
struct NarrowEng {
    using Params = NarrowParams;
    using Span = kmb::Span;
    static constexpr bool kValidate = VALIDATE;
    // MODE 2 re-uses the HASH parameter (it materialises nothing): HASH == false selects the variant whose bin is a field of
    // the high word of the larger strand (2K - hist_bits >= 32)
    static constexpr bool kHiBin = MODE == 2 && KHI && !HASH;
    using Shape = ShapeRun;
    static constexpr bool kTwoPhase = false, kCountOnly = false;
    static constexpr int kSpanEntries = 4;  // tile entries one span reads
    // Resident CTAs per SM the register allocation must allow; 0 = left to the compiler, which lands on 40-63 registers for
    // the fixed-length kernels (an explicit 1 made it take 80-90 and halve the occupancy; an explicit 6 for the
    // canonical-words-only variant made it spill).  The ragged kernels need the cap: 73 registers -> 64 took them 86 % -> 95 %.
    static constexpr int kMinCtas = 0, kMinCtasCsr = FWRC ? 2 : 4;
    const NarrowParams& p;
    Acc acc;
    HistAcc hacc;  // MODE 2 only
    __device__ explicit NarrowEng(const NarrowParams& params) : p(params) {}
    __device__ __forceinline__ uint32_t K() const { return p.wc.K; }
    __device__ __forceinline__ Span load(const uint2* tile, uint32_t rel) const { return load_span<VALIDATE>(tile, rel, p.wc); }
    __device__ __forceinline__ bool dirty(const Span& s) const { return s.inv != 0ull; }
    template <bool TWO, bool CHECK>
    __device__ __forceinline__ void run(const Span& a, const Span& b, uint32_t n_first, uint64_t slot0, uint32_t nwin, const ItemCtx&) {
        if constexpr (MODE == 2) emit_run_hist<TWO, CHECK, DIGEST, KHI, kHiBin>(a, b, n_first, p.wc, p.out, nwin, hacc);
        else emit_run<TWO, CHECK, DIGEST, FWRC, MODE, KHI, HASH>(a, b, n_first, p.wc, p.out, slot0, nwin, acc);
    }
    __device__ __forceinline__ void single(const uint2* tile, uint32_t rel, uint64_t slot, const ItemCtx&) {
        if constexpr (MODE == 2) {
            const Span s = load_span<VALIDATE>(tile, rel, p.wc);
            hist_window<KHI, DIGEST, kHiBin, true>(s, 0, p.wc, p.out, !VALIDATE || (((uint32_t)s.inv) & p.wc.kmask) == 0u, hacc);
        } else {
            emit_single<VALIDATE, DIGEST, FWRC, MODE, KHI>(tile, rel, p.wc, p.out, slot, acc);
        }
    }
    __device__ __forceinline__ void finish(unsigned long long (&red)[3][32]) {
        if (MODE == 2 && DIGEST) hacc.to(acc, mk64(p.wc.mask_lo, p.wc.mask_hi));
        reduce_digest<DIGEST>(red, p.out.digest, acc);
    }
};

template <class Eng>
__global__ void __launch_bounds__(kExtractThreads, Eng::kMinCtas) fixed_kernel(const FixedGeom g, const EncDesc enc, const typename Eng::Params ep) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][32];
    Eng eng(ep);
    fixed_body(g, enc, eng, tile, blockIdx.x);
    eng.finish(red);
}

// dynamic shared memory: [tile_entries x uint2][kCsrCache + 2 offsets][kCsrCache + 2 window offsets]
template <class Eng>
__global__ void __launch_bounds__(kExtractThreads, Eng::kMinCtasCsr) csr_kernel(const CsrGeom g, const EncDesc enc, const typename Eng::Params ep) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][32];
    __shared__ CsrPass pass;
    uint64_t* c_off = reinterpret_cast<uint64_t*>(tile + g.tile_entries);
    uint64_t* c_win = c_off + (kCsrCache + 2);
    Eng eng(ep);
    csr_body(g, enc, eng, tile, c_off, c_win, &pass, blockIdx.x);
    eng.finish(red);
}

// ---------------------------------------------------------------------------
// Fused histogram with shared-memory bins (hist_bits <= 16): a persistent grid (a few CTAs per SM) walks the
// tiles; every CTA counts into its own 2^hist_bits x 16-bit shared histogram and adds it to the global u64 bins
// once at the end -- global atomics drop from one per k-mer to bins per CTA.
// dynamic shared memory: [tile (+ CSR tables)] then the histogram words.
// ---------------------------------------------------------------------------
constexpr int kHistThreads = 1024;  // one big CTA per SM shares the 128 KiB histogram: occupancy comes from its 32 warps

template <class Eng>
__device__ __forceinline__ void smem_hist_flush(unsigned int* s_hist, unsigned long long* hist, uint32_t n_bins) {
    for (uint32_t w = threadIdx.x; w < (n_bins + 1) / 2; w += blockDim.x) {
        const uint32_t v = s_hist[w];
        if (v & 0xFFFFu) atomicAdd(hist + 2 * w, (unsigned long long)(v & 0xFFFFu));
        if (v >> 16) atomicAdd(hist + 2 * w + 1, (unsigned long long)(v >> 16));
    }
}

template <class Eng>
__global__ void __launch_bounds__(kHistThreads) hist_fixed_kernel(const FixedGeom g, const EncDesc enc, typename Eng::Params ep,
                                                                      uint32_t n_tiles, uint32_t tile_words, uint32_t n_bins) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][32];
    unsigned int* s_hist = reinterpret_cast<unsigned int*>(tile + tile_words);
    for (uint32_t w = threadIdx.x; w < (n_bins + 1) / 2; w += blockDim.x) s_hist[w] = 0;
    ep.out.s_hist = s_hist;
    __syncthreads();
    Eng eng(ep);
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        fixed_body(g, enc, eng, tile, t);
        __syncthreads();  // the next tile overwrites the staged stretch
    }
    smem_hist_flush<Eng>(s_hist, ep.out.hist, n_bins);
    eng.finish(red);
}

template <class Eng>
__global__ void __launch_bounds__(kHistThreads) hist_csr_kernel(const CsrGeom g, const EncDesc enc, typename Eng::Params ep,
                                                                    uint32_t n_tiles, uint32_t n_bins) {
    extern __shared__ uint2 tile[];
    __shared__ unsigned long long red[3][32];
    __shared__ CsrPass pass;
    uint64_t* c_off = reinterpret_cast<uint64_t*>(tile + g.tile_entries);
    uint64_t* c_win = c_off + (kCsrCache + 2);
    unsigned int* s_hist = reinterpret_cast<unsigned int*>(c_win + (kCsrCache + 2));
    for (uint32_t w = threadIdx.x; w < (n_bins + 1) / 2; w += blockDim.x) s_hist[w] = 0;
    ep.out.s_hist = s_hist;
    __syncthreads();
    Eng eng(ep);
    for (uint32_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        csr_body(g, enc, eng, tile, c_off, c_win, &pass, t);
        __syncthreads();
    }
    smem_hist_flush<Eng>(s_hist, ep.out.hist, n_bins);
    eng.finish(red);
}

}  // namespace kmb
