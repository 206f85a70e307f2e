// kmb_encoding.cuh -- batched Encoding<P,B> (encoding/mod.rs:14-23) and the
// element-wise naive_impl::Kmer word operations, plus the synthetic generator.
#pragma once
#include "kmb_device.cuh"
#include "kmb_geometry.cuh"

namespace kmb {

// ---------------------------------------------------------------- synthetic reads
// base i = "ACGT"[splitmix64(seed + first + i) >> 62], 'N' under n_thresh20
__global__ void __launch_bounds__(256) generate_kernel(uint8_t* out, uint64_t n, uint64_t seed_plus_first,
                                                       uint32_t n_thresh20) {
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;  // 16 bases each
    const uint64_t i0 = chunk * 16;
    if (i0 >= n) return;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const uint64_t x = splitmix64(seed_plus_first + i0 + i);
        uint32_t c = (0x54474341u >> ((uint32_t)(x >> 62) * 8)) & 0xFFu;  // 'A','C','G','T'
        if (((uint32_t)(x >> 20) & 0xFFFFFu) < n_thresh20) c = 'N';
        w[i >> 2] |= c << ((i & 3) * 8);
    }
    if (i0 + 16 <= n && (reinterpret_cast<uintptr_t>(out + i0) & 15u) == 0) {
        *reinterpret_cast<uint4*>(out + i0) = make_uint4(w[0], w[1], w[2], w[3]);
    } else {
        for (int i = 0; i < 16 && i0 + i < n; ++i) out[i0 + i] = (uint8_t)(w[i >> 2] >> ((i & 3) * 8));
    }
}

// ---------------------------------------------------------------- Encoding::encode (bulk pack)
// encoding/naive.rs:116-124, xor10.rs:52-60.  Flat little-endian bit layout:
// output byte b of a read holds bases 4b..4b+3, zero beyond the read; the
// word width only decides how many padding bytes end the read's region.
// One thread = 16 bases of one read = one 32-bit store.
struct PackParams {
    const uint8_t* bases;
    const uint64_t* offsets;       // CSR or nullptr
    const uint64_t* word_offsets;  // CSR in words, or nullptr (fixed)
    uint64_t n_reads;
    uint64_t L;                    // fixed length (when offsets == nullptr)
    uint64_t out_bytes_per_read;   // fixed: ceil(L / bpw) * word_bytes
    uint32_t word_bytes;
    uint32_t bases_per_word;
    uint8_t* out;
    EncDesc enc;
};

__device__ __forceinline__ uint32_t pack16_bytes(const uint8_t* src, uint32_t n_avail, const EncDesc& enc) {
    // n_avail in [1,16] bases really present; missing ones pack as zero bits
    uint32_t w[4] = {0, 0, 0, 0};
    if (n_avail == 16 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0) {
        uint4 v = ld_stream_v4(src);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i)
            if ((uint32_t)i < n_avail) w[i >> 2] |= (uint32_t)__ldg(src + i) << ((i & 3) * 8);
    }
    PackedWord pw = pack16<false>(make_uint4(w[0], w[1], w[2], w[3]));
    uint32_t bits = apply_encoding(pw.bits, enc);
    if (n_avail < 16) bits &= (1u << (2 * n_avail)) - 1u;  // padding is zero bits, not code(0x00)
    return bits;
}

__device__ __forceinline__ void store_packed32(uint8_t* dst, uint64_t byte_off, uint64_t region_bytes, uint32_t bits) {
    // the read's region may end inside this 32-bit group (word_bits 8/16)
    if (byte_off + 4 <= region_bytes && (reinterpret_cast<uintptr_t>(dst + byte_off) & 3u) == 0) {
        *reinterpret_cast<uint32_t*>(dst + byte_off) = bits;
    } else {
        for (int b = 0; b < 4 && byte_off + b < region_bytes; ++b) dst[byte_off + b] = (uint8_t)(bits >> (8 * b));
    }
}

__global__ void __launch_bounds__(256) pack_fixed_kernel(const PackParams p, uint64_t groups_per_read) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t r = t / groups_per_read;
    if (r >= p.n_reads) return;
    const uint64_t g = t - r * groups_per_read;  // 16-base group inside the read's region
    const uint64_t b0 = g * 16;
    uint32_t bits = 0;
    if (b0 < p.L) bits = pack16_bytes(p.bases + r * p.L + b0, (uint32_t)min((uint64_t)16, p.L - b0), p.enc);
    store_packed32(p.out + r * p.out_bytes_per_read, g * 4, p.out_bytes_per_read, bits);
}

// Fixed-length reads: the same two-phase shape as the extraction kernel, in OUTPUT space.  A CTA owns kPackGroups
// consecutive 32-bit words of the packed image (coalesced 4-byte stores); it stages the stretch of reads they cover
// into shared memory with aligned 16-byte loads, then a word is one funnel-shift extract of two packed entries --
// or, where a read's region ends inside the word (u8 / u16 word types: regions are whole words of P, not of 32 bits),
// four one-byte extracts.
// (scripts/build_variant.sh, u64 / u8 words, % of the copy peak: 1024 groups with 5 loads in flight 66 / 57, 2048 with 3 loads
//  70 / 61, 2048 with 8 74 / 62, 3072 67 / 61, 4096 with 8 loads at four CTAs per SM 76 / 63, 6144 at three and 8192 at two
//  CTAs 68 / 62)
#ifndef KMB_PACK_GROUPS
#define KMB_PACK_GROUPS 4096
#endif
#ifndef KMB_PACK_BATCH
#define KMB_PACK_BATCH 8
#endif
#ifndef KMB_PACK_MINCTAS
#define KMB_PACK_MINCTAS 4
#endif
constexpr int kPackGroups = KMB_PACK_GROUPS;

struct PackTileParams {
    const uint8_t* bases;
    uint64_t n_bytes;
    uint64_t L;
    uint64_t total_bytes;   // n_reads * obr
    uint64_t obr_magic64;   // floor(2^64 / obr) + 1 (obr >= 2), 0 = divide
    uint32_t L32;
    uint32_t obr;           // output bytes per read = ceil(L / bases per word) * word bytes
    uint32_t obr_magic;     // floor(2^32 / obr) + 1
    uint8_t* out;           // 4-byte aligned
    EncDesc enc;
};

// u / obr for u < obr + 4 * kPackGroups
__device__ __forceinline__ uint32_t div_obr(uint32_t u, const PackTileParams& p) {
    if (p.obr >= 4u * kPackGroups) return (u >= p.obr) ? 1u : 0u;
    if (p.obr == 1) return u;
    return __umulhi(u, p.obr_magic);
}

// MODE 0: regions are whole 32-bit words (u32 / u64 / u128 words of P, or lucky lengths): one extract per word
// MODE 1: regions of >= 4 bytes that may end inside a word: two masked extracts per word, branch-free
// MODE 2: regions of 1..3 bytes: byte by byte
template <int MODE>
__global__ void __launch_bounds__(256, KMB_PACK_MINCTAS) pack_tile_kernel(const PackTileParams p) {
    extern __shared__ uint2 tile[];
    const uint64_t byte_base = (uint64_t)blockIdx.x * (4u * kPackGroups);
    const uint32_t n_out = (uint32_t)min((uint64_t)(4u * kPackGroups), p.total_bytes - byte_base);  // output bytes of this CTA
    uint64_t r_first;
    if (p.obr == 1) r_first = byte_base;
    else if (p.obr_magic64) r_first = div_magic64(byte_base, p.obr_magic64);
    else r_first = byte_base / p.obr;
    const uint32_t b_first = (uint32_t)(byte_base - r_first * p.obr);
    // bases from the first byte's first base to the last byte's last base (a byte holds 4 bases; clipped to the read)
    const uint64_t s_start = r_first * p.L + min((uint64_t)b_first * 4, p.L);
    const uint32_t u_last = b_first + n_out - 1, q_last = div_obr(u_last, p);
    const uint64_t s_end = (r_first + q_last) * p.L + min((uint64_t)(u_last - q_last * p.obr) * 4 + 4, p.L);
    const uint8_t* first = p.bases + s_start;
    // MODE 1: one entry of margin in front -- the extract for the second read of a word starts before that read's first base
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(first) & 15u) + (MODE == 1 ? 16u : 0u);
    const uint32_t n_entries = (uint32_t)((s_end - s_start + mis + 15) >> 4) + 2;  // an extract reads 2 entries, possibly from the stretch's very end
    stage_tile<false, KMB_PACK_BATCH>(p.bases, p.n_bytes, first - mis, n_entries, p.enc, tile);
    __syncthreads();
    // the 16 bases from tile position rel (relative to the staged stretch)
    auto extract = [&](uint32_t rel) -> uint32_t {
        const uint32_t e = rel >> 4, o2 = (rel & 15u) * 2;
        return __funnelshift_r(tile[e].x, tile[e + 1].x, o2);
    };
    auto low_bits = [](uint32_t n) -> uint32_t { return n >= 32u ? 0xFFFFFFFFu : ((1u << n) - 1u); };
    // Everything below is 32-bit and relative to the CTA (the host sends reads of 2^30 bases or more elsewhere): base x of
    // read r_first + q sits at tile position q * L + x - s_first + mis.
    const uint32_t L = p.L32, s_first = min(b_first * 4, L), bias = mis - s_first;
    const uint32_t n_words = (n_out + 3) / 4;
    for (uint32_t li = threadIdx.x; li < n_words; li += blockDim.x) {
        const uint32_t u = b_first + 4 * li, q = div_obr(u, p), b = u - q * p.obr;  // first byte of the word inside its read
        const uint32_t here = min(4u, n_out - 4 * li);                              // bytes of this word that exist
        const uint32_t b0 = b * 4;
        uint32_t bits = 0;
        if (MODE == 0) {
            if (b0 < L) bits = extract(q * L + b0 + bias) & low_bits(2 * min(16u, L - b0));
        } else if (MODE == 1) {
            // A word holds bytes of at most two reads: n_a from read q (bases 4b ..), the rest from read q+1 (from its base
            // 0 on).  Both parts are one 16-base extract, masked to the bases that exist -- the same instructions for every
            // lane, whether its word straddles a region boundary or not (a branch here would cost every warp both paths).
            const uint32_t n_a = min(here, p.obr - b);
            const uint32_t have_a = b0 < L ? min(4 * n_a, L - b0) : 0u;  // bases of read q in this word
            const uint32_t rel_a = q * L + min(b0, L) + bias;
            bits = extract(rel_a) & low_bits(2 * have_a);
            const bool two = n_a < here;
            const uint32_t have_b = two ? min(4 * (here - n_a), L) : 0u;
            // read q+1's base 0 lands at base 4 * n_a of the word: start 4 * n_a bases early (inside the margin at worst)
            const uint32_t rel_b = two ? (q + 1) * L + bias - 4 * n_a : rel_a;
            bits |= extract(rel_b) & (low_bits(8 * n_a + 2 * have_b) & ~low_bits(8 * n_a));
        } else {
            for (uint32_t j = 0; j < here; ++j) {  // regions of 1..3 bytes: byte by byte
                const uint32_t uj = u + j, qj = div_obr(uj, p), bj0 = (uj - qj * p.obr) * 4;
                if (bj0 < L) bits |= (extract(qj * L + bj0 + bias) & low_bits(2 * min(4u, L - bj0))) << (8 * j);
            }
        }
        uint8_t* dst = p.out + byte_base + 4 * li;
        if (here == 4) {
            *reinterpret_cast<uint32_t*>(dst) = bits;
        } else {
            for (uint32_t j = 0; j < here; ++j) dst[j] = (uint8_t)(bits >> (8 * j));  // the image's last, partial word
        }
    }
}

// CSR: one warp per read
__global__ void __launch_bounds__(256) pack_csr_kernel(const PackParams p) {
    const uint64_t r = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint32_t lane = threadIdx.x & 31;
    if (r >= p.n_reads) return;
    const uint64_t beg = p.offsets[r], len = p.offsets[r + 1] - beg;
    const uint64_t region = (p.word_offsets[r + 1] - p.word_offsets[r]) * p.word_bytes;
    uint8_t* dst = p.out + p.word_offsets[r] * p.word_bytes;
    const uint64_t groups = (region + 3) / 4;
    for (uint64_t g = lane; g < groups; g += 32) {
        const uint64_t b0 = g * 16;
        uint32_t bits = 0;
        if (b0 < len) bits = pack16_bytes(p.bases + beg + b0, (uint32_t)min((uint64_t)16, len - b0), p.enc);
        store_packed32(dst, g * 4, region, bits);
    }
}

// ---------------------------------------------------------------- Encoding::decode (bulk unpack)
// encoding/naive.rs:126-136 (bits2nuc :88-96).  dec = the four ASCII letters
// for codes 0..3 packed in one u32.  One thread = one input byte = 4 bases.
__global__ void __launch_bounds__(256) unpack_kernel(const uint8_t* in, uint64_t n_items, uint32_t in_bytes_per_item,
                                                     uint32_t bases_per_item, uint32_t dec, uint8_t* out) {
    const uint32_t groups = (bases_per_item + 3) / 4;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t item = t / groups;
    if (item >= n_items) return;
    const uint32_t g = (uint32_t)(t - item * groups);
    const uint32_t byte = in[item * in_bytes_per_item + g];
    uint8_t* dst = out + item * bases_per_item + (uint64_t)g * 4;
    uint32_t chars = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) chars |= ((dec >> (((byte >> (2 * i)) & 3u) * 8)) & 0xFFu) << (8 * i);
    const uint32_t n = min(4u, bases_per_item - g * 4);
    if (n == 4 && (reinterpret_cast<uintptr_t>(dst) & 3u) == 0) {
        *reinterpret_cast<uint32_t*>(dst) = chars;
    } else {
        for (uint32_t i = 0; i < n; ++i) dst[i] = (uint8_t)(chars >> (8 * i));
    }
}

// The same through shared memory: a CTA decodes `ipc` consecutive items (<= kUnpackTile letters) into a staging buffer
// laid out with the misalignment of its slice of the text, then writes the slice with aligned 16-byte stores.
// (Items are bases_per_item letters long -- 31 for a 31-mer -- so per-item stores would be byte-granular.)
constexpr int kUnpackTile = 16384;
__global__ void __launch_bounds__(256) unpack_tile_kernel(const uint8_t* in, uint64_t n_items, uint32_t in_bytes_per_item,
                                                          uint32_t bases_per_item, uint32_t dec, uint8_t* out, uint32_t ipc) {
    __shared__ __align__(16) uint8_t text[kUnpackTile + 32];
    const uint64_t i0 = (uint64_t)blockIdx.x * ipc;
    const uint32_t ni = (uint32_t)min((uint64_t)ipc, n_items - i0);
    const uint64_t g0 = i0 * bases_per_item;  // first letter of this CTA's slice
    const uint32_t n_bytes = ni * bases_per_item;
    const uint32_t mis = (uint32_t)((reinterpret_cast<uintptr_t>(out) + g0) & 15u);
    const uint32_t groups = (bases_per_item + 3) / 4;  // input bytes of an item that hold letters
    for (uint32_t t = threadIdx.x; t < ni * groups; t += blockDim.x) {
        const uint32_t it = t / groups, g = t - it * groups;
        const uint32_t byte = in[(i0 + it) * in_bytes_per_item + g];
        const uint32_t n = min(4u, bases_per_item - 4 * g);
        uint8_t* d = text + mis + it * bases_per_item + 4 * g;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j)
            if (j < n) d[j] = (uint8_t)(dec >> (((byte >> (2 * j)) & 3u) * 8));
    }
    __syncthreads();
    uint8_t* dst = out + g0;
    const uint32_t head = min(n_bytes, (16u - mis) & 15u);  // letters before the first 16-byte boundary
    const uint32_t n_vec = (n_bytes - head) / 16;
    for (uint32_t j = threadIdx.x; j < head; j += blockDim.x) dst[j] = text[mis + j];
    for (uint32_t c = threadIdx.x; c < n_vec; c += blockDim.x)
        *reinterpret_cast<uint4*>(dst + head + 16 * c) = *reinterpret_cast<const uint4*>(text + mis + head + 16 * c);
    for (uint32_t j = head + 16 * n_vec + threadIdx.x; j < n_bytes; j += blockDim.x) dst[j] = text[mis + j];
}

// ---------------------------------------------------------------- Encoding::decode, output-space form
// One thread = one 16-byte aligned block of the OUTPUT text, so every store is a full 16-byte vector whatever the item
// length is (31 letters for a 31-mer); the <= 16 letters of a block come from one or two items (more only for items
// shorter than 16 letters), each fetched as a bit field of the packed input, and four letters at a time are decoded by
// one PRMT: the 2-bit codes are spread to nibbles (three shift-or-mask steps per 8 codes) and select bytes of the
// four-letter table `dec`.  ~3 instructions per letter instead of ~9 (the per-input-byte kernel above: 25 % of the copy peak).
// All per-thread arithmetic is 32-bit and relative to the CTA's first item; SAFE = the CTA's whole input stretch (plus
// the word a funnel shift may touch beyond it) lies inside the buffer, so loads need no bounds checks (all CTAs but the last).
template <bool SAFE>
__device__ __forceinline__ uint32_t fetch_bits32(const uint32_t* __restrict__ words, uint32_t bit, uint64_t words_left_bytes) {
    const uint32_t widx = bit >> 5, sh = bit & 31u;
    uint32_t lo, hi;
    if (SAFE) {
        lo = __ldg(words + widx);
        hi = __ldg(words + widx + 1);
    } else {  // byte-granular guards at the very end of the buffer
        lo = hi = 0;
        const uint8_t* b = reinterpret_cast<const uint8_t*>(words);
#pragma unroll
        for (uint32_t i = 0; i < 4; ++i) {
            if ((uint64_t)widx * 4 + i < words_left_bytes) lo |= (uint32_t)__ldg(b + (uint64_t)widx * 4 + i) << (8 * i);
            if ((uint64_t)widx * 4 + 4 + i < words_left_bytes) hi |= (uint32_t)__ldg(b + (uint64_t)widx * 4 + 4 + i) << (8 * i);
        }
    }
    return __funnelshift_r(lo, hi, sh);
}
// 8 two-bit codes (16 bits) -> 8 letters (two registers)
__device__ __forceinline__ void decode8(uint32_t c16, uint32_t dec, uint32_t& out0, uint32_t& out1) {
    uint32_t x = c16 & 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;  // code j in nibble j
    out0 = __byte_perm(dec, 0u, x & 0xFFFFu);
    out1 = __byte_perm(dec, 0u, x >> 16);
}

// 16-byte blocks per thread (scripts/build_variant.sh; % of the copy peak at 50 M 31-mers): 2 71, 4 81, 8 91, 12 87, 16 87, 32 77
#ifndef KMB_UNPACK_BLOCKS
#define KMB_UNPACK_BLOCKS 8
#endif
constexpr int kUnpackBlocksPerThread = KMB_UNPACK_BLOCKS;

// the <= 16 codes of letters [rel, rel + n) of the CTA's item stream (rel counted from the first letter of the CTA's first item)
template <bool SAFE>
__device__ __forceinline__ uint32_t gather_codes(const uint32_t* __restrict__ cta_words, uint64_t cta_bytes_left, uint32_t bit_base,
                                                 uint32_t rel, uint32_t n, uint32_t item_bits, uint32_t bases_per_item, uint32_t bpi_magic) {
    uint32_t it = bases_per_item == 1 ? rel : __umulhi(rel, bpi_magic);  // exact: rel * bases_per_item < 2^32
    const uint32_t o = rel - it * bases_per_item;
    uint32_t take = min(n, bases_per_item - o);
    uint32_t codes = fetch_bits32<SAFE>(cta_words, bit_base + it * item_bits + 2u * o, cta_bytes_left);
    if (take < n) {  // the block runs into the next item(s)
        codes &= (1u << (2 * take)) - 1u;
        uint32_t done = take;
        do {
            ++it;
            take = min(n - done, bases_per_item);
            uint32_t bits = fetch_bits32<SAFE>(cta_words, bit_base + it * item_bits, cta_bytes_left);
            if (take < 16u) bits &= (1u << (2 * take)) - 1u;
            codes |= bits << (2 * done);
            done += take;
        } while (done < n);
    }
    return codes;
}

// INTERIOR: every block of the CTA is a full 16 letters inside the text (all CTAs but the first and the last): no per-block
// bounds arithmetic, everything 32-bit and relative to the CTA.
template <bool SAFE, bool INTERIOR>
__device__ __forceinline__ void unpack_flat_body(const uint32_t* __restrict__ cta_words, uint64_t cta_bytes_left, uint32_t bit_base, uint32_t cta_off,
                                                 int64_t cta_p0, uint32_t item_bits, uint32_t bases_per_item, uint32_t bpi_magic,
                                                 uint32_t dec, uint8_t* __restrict__ out, uint64_t total_letters, uint32_t n_blocks_cta) {
#pragma unroll
    for (int u = 0; u < kUnpackBlocksPerThread; ++u) {
        const uint32_t qrel = (uint32_t)u * 256u + threadIdx.x;
        if (!INTERIOR && qrel >= n_blocks_cta) continue;
        uint32_t j0 = 0, n = 16;
        uint32_t rel = 16u * qrel + cta_off;  // INTERIOR: cta_off = offset of letter cta_p0 inside the CTA's first item
        if (!INTERIOR) {
            const int64_t p0 = cta_p0 + 16 * (int64_t)qrel;
            const uint64_t lo = p0 < 0 ? 0ull : (uint64_t)p0;
            const uint64_t hi = min((uint64_t)(p0 + 16), total_letters);
            j0 = (uint32_t)((int64_t)lo - p0);  // non-zero only in the very first block of the text
            n = (uint32_t)(hi - lo);
            rel = (uint32_t)(lo - (uint64_t)(cta_p0 < 0 ? 0 : cta_p0)) + cta_off;
        }
        uint32_t codes = gather_codes<SAFE>(cta_words, cta_bytes_left, bit_base, rel, n, item_bits, bases_per_item, bpi_magic);
        if (!INTERIOR) codes <<= 2 * j0;
        uint32_t c0, c1, c2, c3;
        decode8(codes, dec, c0, c1);
        decode8(codes >> 16, dec, c2, c3);
        uint8_t* dst = out + (cta_p0 + 16 * (int64_t)qrel);  // 16-byte aligned by construction
        if (INTERIOR || n == 16u) {
            *reinterpret_cast<uint4*>(dst) = make_uint4(c0, c1, c2, c3);
        } else {
#pragma unroll
            for (uint32_t j = 0; j < 16; ++j) {
                const uint32_t c = j < 4 ? c0 : (j < 8 ? c1 : (j < 12 ? c2 : c3));
                if (j >= j0 && j < j0 + n) dst[j] = (uint8_t)(c >> (8 * (j & 3u)));
            }
        }
    }
}

__global__ void __launch_bounds__(256) unpack_flat_kernel(const uint8_t* __restrict__ in, uint64_t in_total_bytes, uint64_t n_items,
                                                          uint32_t in_bytes_per_item, uint32_t bases_per_item, uint32_t bpi_magic,
                                                          uint64_t bpi_magic64, uint32_t dec, uint8_t* __restrict__ out, uint64_t total_letters,
                                                          uint64_t n_blocks) {
    constexpr uint32_t kBlocksPerCta = 256 * kUnpackBlocksPerThread;
    const uint32_t mis = (uint32_t)(reinterpret_cast<uintptr_t>(out) & 15u);  // block q covers letters [16 q - mis, 16 q - mis + 16)
    const uint64_t q0 = (uint64_t)blockIdx.x * kBlocksPerCta;
    const int64_t cta_p0 = (int64_t)(q0 * 16) - (int64_t)mis;  // first letter of the CTA's first block (negative only in CTA 0)
    // first item of the CTA: one 64-bit division (by multiplication), everything after it in 32 bits relative to that item
    const uint64_t cta_p = cta_p0 < 0 ? 0ull : (uint64_t)cta_p0;
    const uint64_t cta_item = bases_per_item == 1 ? cta_p : __umul64hi(cta_p, bpi_magic64);
    const uint32_t cta_off = (uint32_t)(cta_p - cta_item * bases_per_item);
    const uint64_t cta_byte0 = cta_item * in_bytes_per_item;  // u8 / u16 word types: not always a multiple of 4 -> bit_base
    const uint32_t* cta_words = reinterpret_cast<const uint32_t*>(in + (cta_byte0 & ~3ull));
    const uint32_t bit_base = (uint32_t)(cta_byte0 & 3ull) * 8u;
    const uint64_t left = in_total_bytes - (cta_byte0 & ~3ull);
    // input the CTA can touch: its letters span at most (letters / bases_per_item + 2) items, plus the words a funnel shift may touch
    const uint64_t span_items = (uint64_t)(kBlocksPerCta * 16 + cta_off) / bases_per_item + 2;
    const bool safe = span_items * in_bytes_per_item + 12 <= left;
    const bool interior = cta_p0 >= 0 && (uint64_t)cta_p0 + kBlocksPerCta * 16ull <= total_letters;
    const uint32_t n_blocks_cta = (uint32_t)min((uint64_t)kBlocksPerCta, n_blocks - q0);
    const uint32_t item_bits = in_bytes_per_item * 8u;
    if (safe && interior)
        unpack_flat_body<true, true>(cta_words, left, bit_base, cta_off, cta_p0, item_bits, bases_per_item, bpi_magic, dec, out, total_letters, n_blocks_cta);
    else if (safe)
        unpack_flat_body<true, false>(cta_words, left, bit_base, cta_off, cta_p0, item_bits, bases_per_item, bpi_magic, dec, out, total_letters, n_blocks_cta);
    else
        unpack_flat_body<false, false>(cta_words, left, bit_base, cta_off, cta_p0, item_bits, bases_per_item, bpi_magic, dec, out, total_letters, n_blocks_cta);
}


// how an item is read and written: widest natural vector when item size and addresses allow, else words, else bytes
template <int NW32>
struct ItemIo {
    static constexpr int VB = NW32 >= 4 ? 16 : 4 * NW32;
    bool aligned, vec;
    __device__ __forceinline__ ItemIo(const uint8_t* src, const uint8_t* dst, uint32_t item_bytes) {
        aligned = (item_bytes == 4u * NW32) && ((reinterpret_cast<uintptr_t>(src) & 3u) == 0) && ((reinterpret_cast<uintptr_t>(dst) & 3u) == 0);
        vec = aligned && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & (VB - 1)) == 0;
    }
    __device__ __forceinline__ void load(const uint8_t* src, uint32_t item_bytes, uint32_t (&w)[NW32]) const {
        if (vec && NW32 >= 4) {
#pragma unroll
            for (int i = 0; i < NW32 / 4; ++i) {
                const uint4 v = reinterpret_cast<const uint4*>(src)[i];
                w[4 * i] = v.x; w[4 * i + 1] = v.y; w[4 * i + 2] = v.z; w[4 * i + 3] = v.w;
            }
        } else if (vec && NW32 == 2) {
            const uint2 v = *reinterpret_cast<const uint2*>(src);
            w[0] = v.x; w[1] = v.y;
        } else if (aligned) {
#pragma unroll
            for (int i = 0; i < NW32; ++i) w[i] = reinterpret_cast<const uint32_t*>(src)[i];
        } else {
#pragma unroll
            for (int i = 0; i < NW32; ++i) {
                uint32_t v = 0;
                for (int b = 0; b < 4; ++b)
                    if ((uint32_t)(4 * i + b) < item_bytes) v |= (uint32_t)src[4 * i + b] << (8 * b);
                w[i] = v;
            }
        }
    }
    __device__ __forceinline__ void store(uint8_t* dst, uint32_t item_bytes, const uint32_t (&o)[NW32]) const {
        if (vec && NW32 >= 4) {
#pragma unroll
            for (int i = 0; i < NW32 / 4; ++i)
                reinterpret_cast<uint4*>(dst)[i] = make_uint4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        } else if (vec && NW32 == 2) {
            *reinterpret_cast<uint2*>(dst) = make_uint2(o[0], o[1]);
        } else if (aligned) {
#pragma unroll
            for (int i = 0; i < NW32; ++i) reinterpret_cast<uint32_t*>(dst)[i] = o[i];
        } else {
#pragma unroll
            for (int i = 0; i < NW32; ++i)
                for (int b = 0; b < 4; ++b)
                    if ((uint32_t)(4 * i + b) < item_bytes) dst[4 * i + b] = (uint8_t)(o[i] >> (8 * b));
        }
    }
};

// Encoding::rev_comp::<K> (encoding/naive.rs:138-154: swap base slots i and K-1-i, complement both; fields >= K keep what
// they held) on the NW32 32-bit words of one array.  Reversing the whole array field by field puts field f at
// 16 NW32 - 1 - f; K - 1 - f is that shifted right by s = 2 (16 NW32 - K) bits, and the shift leaves exactly the low 2K bits
// occupied.  s is the same for every item: its word part WO selects the code, so every index below is a constant
// (a run-time word index into a register array costs a select chain per access: this kernel ran at 72 % ALU, 77 % of
// the copy peak, with one get32 per output word).
template <int NW32, int WO>
__device__ __forceinline__ void revcomp_array(const uint32_t (&w)[NW32], uint32_t (&o)[NW32], uint32_t K, uint32_t cmask, uint32_t sb) {
    uint32_t r[NW32];
#pragma unroll
    for (int j = 0; j < NW32; ++j) r[j] = pair_reverse32(w[NW32 - 1 - j] ^ cmask);
#pragma unroll
    for (int i = 0; i < NW32; ++i) {
        const uint32_t lo = i + WO < NW32 ? r[i + WO] : 0u, hi = i + WO + 1 < NW32 ? r[i + WO + 1] : 0u;
        const int nvalid = max(0, min(16, (int)K - 16 * i));  // fields of this word below K
        const uint32_t vmask = nvalid == 16 ? 0xFFFFFFFFu : ((1u << (2 * nvalid)) - 1u);
        o[i] = __funnelshift_r(lo, hi, sb) | (w[i] & ~vmask);
    }
}
template <int NW32, int WO = 0>
__device__ __forceinline__ void revcomp_dispatch(const uint32_t (&w)[NW32], uint32_t (&o)[NW32], uint32_t K, uint32_t cmask, uint32_t wo,
                                                 uint32_t sb) {
    if constexpr (WO + 1 < NW32) {
        if (wo != (uint32_t)WO) return revcomp_dispatch<NW32, WO + 1>(w, o, K, cmask, wo, sb);
    }
    revcomp_array<NW32, WO>(w, o, K, cmask, sb);
}

// kRevItems items per thread, a warp's items interleaved (item = base + u * 32 + lane) so that every load and store
// instruction stays coalesced while each thread keeps kRevItems independent loads in flight.
constexpr int kRevItems = 4;
template <int NW32>
__global__ void __launch_bounds__(256) revcomp_items_kernel(const uint8_t* in, uint8_t* out, uint64_t n_items,
                                                            uint32_t item_bytes, uint32_t K, uint32_t cmask) {
    const uint64_t warp0 = ((uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u)) * kRevItems + (threadIdx.x & 31u);
    const ItemIo<NW32> io(in, out, item_bytes);  // item stride keeps every item's alignment class equal to the first's
    const uint32_t s = 2u * (16u * NW32 - K);    // 2K <= 8 item_bytes <= 32 NW32 (checked on the host)
    uint32_t w[kRevItems][NW32];
#pragma unroll
    for (int u = 0; u < kRevItems; ++u) {
        const uint64_t item = warp0 + (uint64_t)u * 32;
        if (item < n_items) io.load(in + item * item_bytes, item_bytes, w[u]);
    }
#pragma unroll
    for (int u = 0; u < kRevItems; ++u) {
        const uint64_t item = warp0 + (uint64_t)u * 32;
        if (item >= n_items) continue;
        uint32_t o[NW32];
        revcomp_dispatch<NW32>(w[u], o, K, cmask, s >> 5, s & 31u);
        io.store(out + item * item_bytes, item_bytes, o);
    }
}

// ---------------------------------------------------------------- naive_impl::Kmer word ops
// OP 0: get_reverse_complement_word (kmer.rs:138-147)
// OP 1: to_canonical + is_canonical  (kmer.rs:55-58, 68-74)
// OP 2: LexHasher                    (hash.rs:60-71)
// OP 3: get_word_equivalency         (canonical_kmer.rs:42-52, 152-161)
__device__ __forceinline__ uint64_t rc_word(uint64_t w, uint32_t k) { return pair_reverse64(~w) >> (2 * (32 - k)); }

template <int OP>
__device__ __forceinline__ void word_op_one(uint64_t w, uint64_t o, uint32_t k, uint64_t mask, uint64_t& r64, uint8_t& r8) {
    r64 = 0; r8 = 0;
    if (OP == 0) {
        r64 = rc_word(w, k);
    } else if (OP == 1) {
        const uint64_t rc = rc_word(w, k);
        const bool is_canon = w <= rc;  // kmer.rs:57  *self <= rc
        r64 = is_canon ? w : rc;
        r8 = is_canon ? 1 : 0;
    } else if (OP == 2) {
        r64 = pair_reverse64(w) >> (2 * (32 - k));
    } else {
        const uint64_t fw = w & mask;  // Kmer::from_u64 masks (kmer.rs:45-48), intended mask at k == 32
        const uint64_t rc = rc_word(fw, k);
        r8 = (fw == o) ? 1 : ((rc == o) ? 2 : 0);
    }
}

// One thread = 4 consecutive words: 32-byte loads and stores when the arrays are 32-byte aligned.
template <int OP>
__global__ void __launch_bounds__(256) word_op_kernel(const uint64_t* in, const uint64_t* other, uint64_t* out,
                                                      uint8_t* out8, uint64_t n, uint32_t k, uint64_t mask, uint32_t vec_ok) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    uint64_t w[4], o[4] = {0, 0, 0, 0}, r64[4];
    uint8_t r8[4];
    const bool full = vec_ok && i0 + 4 <= n;
    if (full) {
        const ulonglong4 v = *reinterpret_cast<const ulonglong4*>(in + i0);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        if (OP == 3) {
            const ulonglong4 u = *reinterpret_cast<const ulonglong4*>(other + i0);
            o[0] = u.x; o[1] = u.y; o[2] = u.z; o[3] = u.w;
        }
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            w[t] = (i0 + t < n) ? in[i0 + t] : 0;
            if (OP == 3) o[t] = (i0 + t < n) ? other[i0 + t] : 0;
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) word_op_one<OP>(w[t], o[t], k, mask, r64[t], r8[t]);
    if (full) {
        if (OP != 3 && out) st_stream_v4u64(out + i0, r64[0], r64[1], r64[2], r64[3]);
        if ((OP == 1 || OP == 3) && out8)
            *reinterpret_cast<uint32_t*>(out8 + i0) = (uint32_t)r8[0] | ((uint32_t)r8[1] << 8) | ((uint32_t)r8[2] << 16) | ((uint32_t)r8[3] << 24);
    } else {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (i0 + t < n) {
                if (OP != 3 && out) out[i0 + t] = r64[t];
                if ((OP == 1 || OP == 3) && out8) out8[i0 + t] = r8[t];
            }
        }
    }
}

// ---------------------------------------------------------------- the small Kmer / CanonicalKmer accessors, batched
// naive_impl/mod.rs:40-50 on one byte: A/a 0, C/c 1, G/g 2, T/t 3, anything else u64::MAX (no guard in the *_u8 shifts)
__device__ __forceinline__ uint64_t encode_binary_u8_dev(uint32_t c) {
    const uint32_t u = c & 0xDFu;
    return u == 'A' ? 0ull : (u == 'C' ? 1ull : (u == 'G' ? 2ull : (u == 'T' ? 3ull : ~0ull)));
}
// naive_impl/kmer.rs:76-102.  `mask` = intended MASK_TABLE[k] (all ones at k == 32).
__device__ __forceinline__ uint64_t append_base_dev(uint64_t& data, uint64_t c, uint32_t k) {
    const uint64_t r = data & 3ull;
    data = (data >> 2) | (c << (2 * k - 2));
    return r;
}
__device__ __forceinline__ uint64_t prepend_base_dev(uint64_t& data, uint64_t c, uint32_t k, uint64_t mask) {
    const uint64_t r = (data >> (2 * k - 2)) & 3ull;
    data = mask & ((data << 2) | c);
    return r;
}

// OP 0: Kmer::sub_kmer_word (kmer.rs:150-161), a = pos, b_mask = MASK_TABLE[width]
// OP 1: Kmer::append_base[_u8]   OP 2: Kmer::prepend_base[_u8]   (kmer.rs:76-102)  -> shifted word + the base shifted off
// OP 3: CanonicalKmer::append_base[_u8]  OP 4: CanonicalKmer::prepend_base[_u8]  (canonical_kmer.rs:70-100): fw and rc words
// OP 5: CanonicalKmer::is_fw_canonical (canonical_kmer.rs:67-69): fw < rc
// bases: one per element -- 2-bit codes (Base) or, when ascii != 0, letters through encode_binary_u8
template <int OP>
__global__ void __launch_bounds__(256) kmer_shift_kernel(const uint64_t* __restrict__ fw_in, const uint64_t* __restrict__ rc_in,
                                                         const uint8_t* __restrict__ bases, uint32_t ascii, uint64_t* __restrict__ fw_out,
                                                         uint64_t* __restrict__ rc_out, uint8_t* __restrict__ out8, uint64_t n, uint32_t k,
                                                         uint64_t mask, uint32_t a, uint64_t b_mask) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t fw = fw_in[i];
    if (OP == 0) {
        fw_out[i] = (fw >> (2 * a)) & b_mask;
        return;
    }
    if (OP == 5) {
        out8[i] = fw < rc_in[i] ? 1 : 0;
        return;
    }
    const uint32_t raw = bases[i];
    const uint64_t c = ascii ? encode_binary_u8_dev(raw) : (uint64_t)raw;
    uint64_t r;
    if (OP == 1) r = append_base_dev(fw, c, k);
    if (OP == 2) r = prepend_base_dev(fw, c, k, mask);
    if (OP == 3 || OP == 4) {
        uint64_t rc = rc_in[i];
        const uint64_t cb = 3ull - c;  // complement_base (naive_impl/mod.rs:81-84); wraps like a release build for an invalid byte
        if (OP == 3) { r = append_base_dev(fw, c, k); prepend_base_dev(rc, cb, k, mask); }
        else { r = prepend_base_dev(fw, c, k, mask); append_base_dev(rc, cb, k); }
        if (rc_out) rc_out[i] = rc;
    }
    if (fw_out) fw_out[i] = fw;
    if (out8) out8[i] = (uint8_t)r;
}

// kmer::Kmer<P,K,B>::get (kmer.rs:46-48): the 2-bit field `index` of every [P; B] array (byte image)
__global__ void __launch_bounds__(256) kmer_get_kernel(const uint8_t* __restrict__ in, uint64_t n_items, uint32_t item_bytes, uint32_t index,
                                                       uint8_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_items) return;
    out[i] = (in[i * item_bytes + (index >> 2)] >> (2 * (index & 3u))) & 3u;
}
// kmer::Kmer<P,K,B>::get_prefix (kmer.rs:50-52): bits 0 ..= 2 len of the array -- 2 len + 1 bits, the reference's inclusive
// range -- as one word of P (word_bytes bytes, little endian); one thread per output byte
__global__ void __launch_bounds__(256) kmer_get_prefix_kernel(const uint8_t* __restrict__ in, uint64_t n_items, uint32_t item_bytes,
                                                              uint32_t word_bytes, uint32_t n_bits, uint8_t* __restrict__ out) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_items * word_bytes) return;
    const uint64_t i = t / word_bytes;
    const uint32_t b = (uint32_t)(t - i * word_bytes);
    uint32_t v = 0;
    if (8 * b < n_bits) {
        v = in[i * item_bytes + b];
        if (n_bits - 8 * b < 8) v &= (1u << (n_bits - 8 * b)) - 1u;
    }
    out[t] = (uint8_t)v;
}

// per-read window / word counts for the CSR prefix sums
__global__ void __launch_bounds__(256) read_counts_kernel(const uint64_t* offsets, uint64_t n_reads, uint32_t k_minus_1,
                                                          uint32_t div, uint64_t* counts) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    if (r == n_reads) { counts[r] = 0; return; }
    const uint64_t len = offsets[r + 1] - offsets[r];
    if (div == 0) counts[r] = len > k_minus_1 ? len - k_minus_1 : 0;  // windows
    else counts[r] = (len + div - 1) / div;                             // packed words
}

// One thread per CTA of the extraction grid: the read owning the tile's first slot (a log2(n_reads)-deep search: all of them
// run concurrently here instead of serially at the head of every extraction CTA), the reads owning its last slot and the
// next tile's first, and the tile's staged stretch when the whole tile fits one pass of tile_bases bases -- everything a
// CTA needs before its first fetch, in one 32-byte descriptor.
__global__ void __launch_bounds__(256) csr_index_kernel(const uint64_t* offsets, const uint64_t* win_offsets, uint64_t n_reads, uint64_t total_slots,
                                                        uint64_t slots_per_cta, uint64_t grid, uint32_t k, uint32_t tile_bases,
                                                        CsrTileDesc* tile_desc) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= grid) return;
    const uint64_t slot = b * slots_per_cta;
    const uint64_t r_lo = last_le(win_offsets, 0, n_reads - 1, slot);  // skips window-less reads: takes the last equal entry
    const uint64_t slot_end = min(total_slots, slot + slots_per_cta);
    const uint64_t r_last = last_le(win_offsets, r_lo, n_reads - 1, slot_end - 1);
    const uint64_t r_hi = slot_end < total_slots ? last_le(win_offsets, r_last, n_reads - 1, slot_end) : r_last;
    CsrTileDesc td;
    td.g0 = offsets[r_lo] + (slot - win_offsets[r_lo]);
    td.r_lo = r_lo;
    td.d_last = (uint32_t)(r_last - r_lo);
    td.d_hi = (uint32_t)min(r_hi - r_lo, (uint64_t)0xFFFFFFFFu);
    const uint64_t g_end = offsets[r_last] + (slot_end - 1 - win_offsets[r_last]) + k;
    td.span = g_end - td.g0 <= (uint64_t)tile_bases ? (uint32_t)(g_end - td.g0) : 0xFFFFFFFFu;
    // narrow: one pass, the reads' offsets fit the CTA's cache, and they span less than 2^31 bases and windows
    const uint64_t p0 = slot - win_offsets[r_lo];
    const bool narrow = td.span != 0xFFFFFFFFu && r_hi - r_lo + 2 <= (uint64_t)kCsrCache + 2 &&
                        win_offsets[r_hi + 1] - win_offsets[r_lo] < (1ull << 31) && offsets[r_hi + 1] - offsets[r_lo] < (1ull << 31);
    td.narrow = narrow ? (kCsrNarrow | (uint32_t)p0) : 0u;
    tile_desc[b] = td;
}

}  // namespace kmb
