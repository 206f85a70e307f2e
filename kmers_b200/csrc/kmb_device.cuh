// kmb_device.cuh -- device-side building blocks shared by the kernels.
// sm_100a only.  Everything here is integer / byte work; tensor cores are
// deliberately unused (HBM-bound path).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace kmb {

constexpr int kRun = 8;             // windows per work item (one thread, 64 B of each output array)

// ---------------------------------------------------------------- memory ops
// Streaming 16-byte read of the read bytes: read once, keep out of L1.
__device__ __forceinline__ uint4 ld_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// Output stores.  The result arrays are written once and never re-read by the
// kernel: KMB_ST selects the cache operator (default .cs = streaming / evict-first).
#ifndef KMB_ST
#define KMB_ST ".cs"
#endif
// 32-byte (256-bit, sm_100+) store: one full sector per lane.
__device__ __forceinline__ void st_stream_v4u64(uint64_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    asm volatile("st.global" KMB_ST ".v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}
__device__ __forceinline__ void st_stream_v2u64(uint64_t* p, uint64_t a, uint64_t b) {
    asm volatile("st.global" KMB_ST ".v2.b64 [%0], {%1,%2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void st_stream_u64(uint64_t* p, uint64_t a) {
    asm volatile("st.global" KMB_ST ".b64 [%0], %1;" ::"l"(p), "l"(a) : "memory");
}

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// ---------------------------------------------------------------- bit tricks
// Reverse the order of the sixteen 2-bit fields of a 32-bit word: bit
// reversal (BREV) then swap the two bits inside every pair.  Equivalent to the
// five mask-and-shift stages of naive_impl/kmer.rs:125-130 on half a word.
__device__ __forceinline__ uint32_t pair_reverse32(uint32_t x) {
    uint32_t y = __brev(x);
    return ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
}
__device__ __forceinline__ uint64_t pair_reverse64(uint64_t x) {
    uint32_t lo = pair_reverse32((uint32_t)x);
    uint32_t hi = pair_reverse32((uint32_t)(x >> 32));
    return ((uint64_t)lo << 32) | hi;
}

// Encoding description resolved on the host from the Naive discriminant
// (encoding/naive.rs:49-74, nuc2bits :78-86).  The tile is first packed in the
// "internal" code x = (c >> 1) & 3 (A0 C1 T2 G3, naive.rs:14-16 == Xor10,
// xor10.rs:21); any of the 24 encodings is then a per-field permutation,
// applied bit-sliced on 16 fields at once through its algebraic normal form:
//   out_b = k0_b ^ (k1_b & x0) ^ (k2_b & x1) ^ (k3_b & x0 & x1),  b in {0,1}.
// The complement under every variant is XOR with one constant pair
// (naive.rs:98-110; a^t == c^g for all 24), replicated in cmask.
struct EncDesc {
    uint32_t k0[2], k1[2], k2[2], k3[2];  // each 0 or 0x55555555
    uint32_t cmask;                       // complement constant replicated over 16 fields
    uint32_t is_acgt;                     // fast path: code = x ^ (x >> 1)
};

__device__ __forceinline__ uint32_t apply_encoding(uint32_t internal, const EncDesc& e) {
    if (e.is_acgt) return internal ^ ((internal >> 1) & 0x55555555u);
    uint32_t x0 = internal & 0x55555555u;
    uint32_t x1 = (internal >> 1) & 0x55555555u;
    uint32_t x01 = x0 & x1;
    uint32_t o0 = e.k0[0] ^ (e.k1[0] & x0) ^ (e.k2[0] & x1) ^ (e.k3[0] & x01);
    uint32_t o1 = e.k0[1] ^ (e.k1[1] & x0) ^ (e.k2[1] & x1) ^ (e.k3[1] & x01);
    return o0 | (o1 << 1);
}

// 4 ASCII bytes -> their four internal 2-bit codes gathered into bits 31:24
// (field order = byte order).  One multiply does the gather: the fields sit
// at bits 0,8,16,24 and land on 24,26,28,30 without carries.
__device__ __forceinline__ uint32_t pack4_top(uint32_t w) {
    uint32_t x = (w >> 1) & 0x03030303u;
    return x * 0x01041040u;
}

// 4 ASCII bytes -> per-byte difference from the letter the byte claims to be;
// a non-zero byte = not one of ACGTacgt (naive_impl/mod.rs:40-50).  The 2-bit
// code already says which letter the byte must be; rebuild that letter (upper
// case) from bits 2:1 and compare with the case-folded byte.
__device__ __forceinline__ uint32_t letter_diff4(uint32_t w) {
    // expected upper-case byte from bits 2:1 : A 0x41, C 0x43, G 0x47, T 0x45^0x11
    uint32_t e0 = (w & 0x06060606u) | 0x41414141u;
    uint32_t t = (w >> 2) & ~(w >> 1) & 0x01010101u;  // the byte claims to be T
    return ((w & 0xDFDFDFDFu) ^ e0) ^ (t * 0x11u);
}
// per-byte non-zero -> 4-bit mask in bits 31:28 (bit 28+i = byte i)
__device__ __forceinline__ uint32_t nonzero4_top(uint32_t diff) {
    uint32_t nz = (((diff & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | diff) & 0x80808080u;  // non-zero byte -> its bit 7
    return nz * 0x00204081u;                                                   // bits 7,15,23,31 -> 28..31
}

struct PackedWord {
    uint32_t bits;  // 16 bases, internal code, base i at bits 2i+1:2i
    uint32_t inv;   // bit i = base i invalid (low 16 bits)
};

// 16 ASCII bytes -> packed word (+ invalid mask when VALIDATE)
template <bool VALIDATE>
__device__ __forceinline__ PackedWord pack16(uint4 v) {
    uint32_t m0 = pack4_top(v.x), m1 = pack4_top(v.y), m2 = pack4_top(v.z), m3 = pack4_top(v.w);
    PackedWord r;
    r.bits = __byte_perm(__byte_perm(m0, m1, 0x0073), __byte_perm(m2, m3, 0x0073), 0x5410);
    r.inv = 0;
    if (VALIDATE) {
        const uint32_t d0 = letter_diff4(v.x), d1 = letter_diff4(v.y), d2 = letter_diff4(v.z), d3 = letter_diff4(v.w);
        if ((d0 | d1 | d2 | d3) != 0u) {  // rare: only then locate the offending bytes
            uint32_t acc = nonzero4_top(d3) >> 28;
            acc = __funnelshift_l(nonzero4_top(d2), acc, 4);
            acc = __funnelshift_l(nonzero4_top(d1), acc, 4);
            acc = __funnelshift_l(nonzero4_top(d0), acc, 4);
            r.inv = acc;
        }
    }
    return r;
}

// Guarded 16-byte fetch of the flat read stream: vector load when the chunk
// lies wholly inside [base, base+n), else byte-wise with zeros outside (a zero
// byte is an invalid base, so nothing outside the batch can form a window).
static __device__ __noinline__ uint4 load16_guarded(const uint8_t* base, uint64_t n_bytes, const uint8_t* p) {
    if (p >= base && p + 16 <= base + n_bytes) return ld_stream_v4(p);
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint8_t* q = p + i;
        uint32_t b = (q >= base && q < base + n_bytes) ? (uint32_t)(*q) : 0u;
        w[i >> 2] |= b << ((i & 3) * 8);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// 96-bit logical right shift by t in [0, 96)
__device__ __forceinline__ void shr96(uint32_t& d0, uint32_t& d1, uint32_t& d2, uint32_t c0, uint32_t c1,
                                      uint32_t c2, uint32_t t) {
    uint32_t s = t & 31u;
    if (t >= 64u) {
        d0 = c2 >> s; d1 = 0; d2 = 0;
    } else if (t >= 32u) {
        d0 = __funnelshift_r(c1, c2, s); d1 = c2 >> s; d2 = 0;
    } else {
        d0 = __funnelshift_r(c0, c1, s); d1 = __funnelshift_r(c1, c2, s); d2 = c2 >> s;
    }
}

// n / d through a precomputed magic = floor(2^64 / d) + 1 (d >= 2); exact while n * d < 2^64
__device__ __forceinline__ uint64_t div_magic64(uint64_t n, uint64_t magic) { return __umul64hi(n, magic); }

__device__ __forceinline__ uint64_t mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// warp-wide wrapping sum of a u64
__device__ __forceinline__ uint64_t warp_sum64(uint64_t v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace kmb
