/*
 * kmers_oracle.c -- CPU ORACLE (test infrastructure, NOT the product).
 * See kmers_oracle.h for scope, parity status and who may load this.
 *
 * Plain-C restatement of COMBINE-lab/kmers' CPU algorithms for the hot path.
 * Citations are file:line under /root/reference/src.
 */
#include "kmers_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ======================================================================= */
/* Path N : naive_impl                                                      */
/* ======================================================================= */

/* naive_impl/mod.rs:40-50 : match on b'A'|b'a' .. b'T'|b't', else u64::MAX */
uint64_t ko_encode_binary_u8(uint8_t c) {
    switch (c) {
    case 'A': case 'a': return 0; /* mod.rs:21 A */
    case 'C': case 'c': return 1; /* mod.rs:22 C */
    case 'G': case 'g': return 2; /* mod.rs:23 G */
    case 'T': case 't': return 3; /* mod.rs:24 T */
    default: return KO_INVALID_BASE;
    }
}

/* naive_impl/mod.rs:27-37 : same table, panic!() otherwise */
int ko_encode_binary(uint8_t c, uint64_t *out) {
    uint64_t b = ko_encode_binary_u8(c);
    if (b == KO_INVALID_BASE) return KO_PANIC;
    *out = b;
    return KO_OK;
}

/* naive_impl/mod.rs:81-84 : 3 - b (wrapping for invalid codes, release build) */
uint64_t ko_complement_base(uint64_t b) { return (uint64_t)3 - b; }

/* naive_impl/mod.rs:87-89 */
int ko_is_valid_nuc(uint64_t b) { return b < 4; }

/* naive_impl/kmer.rs:30-32 bitmask(pos) = (1<<pos)-1 ; :584-618 MASK_TABLE,
 * whose last entry (k == 32) is the literal 0, not all-ones. */
uint64_t ko_mask_table(unsigned k, int strict) {
    if (k < 32) return (((uint64_t)1) << (2 * k)) - 1;
    return strict ? 0 : UINT64_MAX;
}

/* naive_impl/kmer.rs:234-251 : iterate the slice in reverse, w <<= 2; w |= code */
int ko_kmer_from_bytes(const uint8_t *s, size_t len, ko_kmer *out) {
    if (len > 32) return KO_PANIC; /* kmer.rs:236-238 */
    uint64_t w = 0;
    for (size_t n = len; n > 0; --n) {
        uint64_t b;
        if (ko_encode_binary(s[n - 1], &b) != KO_OK) return KO_PANIC;
        w <<= 2;
        w |= b;
    }
    out->k = (uint8_t)len;
    out->data = w;
    return KO_OK;
}

/* naive_impl/kmer.rs:45-48 */
ko_kmer ko_kmer_from_u64(uint64_t data, uint8_t k, int strict) {
    ko_kmer km;
    km.k = k;
    km.data = data & ko_mask_table(k, strict);
    return km;
}

/* naive_impl/kmer.rs:196-207, BASE_TABLE :24 */
void ko_kmer_to_string(ko_kmer km, char *out) {
    static const char base_table[4] = {'a', 'c', 'g', 't'};
    uint64_t w = km.data;
    for (unsigned i = 0; i < km.k; ++i) {
        out[i] = base_table[w & 3u];
        w >>= 2;
    }
    out[km.k] = '\0';
}

/* naive_impl/kmer.rs:98-102 */
uint64_t ko_kmer_append_base(ko_kmer *km, uint64_t c) {
    uint64_t r = km->data & 0x03;
    km->data = (km->data >> 2) | (c << (2 * (unsigned)km->k - 2));
    return r;
}

/* naive_impl/kmer.rs:91-95 */
uint64_t ko_kmer_prepend_base(ko_kmer *km, uint64_t c, int strict) {
    uint64_t r = (km->data >> (2 * (unsigned)km->k - 2)) & 0x03;
    km->data = ko_mask_table(km->k, strict) & ((km->data << 2) | c);
    return r;
}

/* naive_impl/kmer.rs:83-88 */
uint64_t ko_kmer_append_base_u8(ko_kmer *km, uint8_t c) {
    return ko_kmer_append_base(km, ko_encode_binary_u8(c));
}

/* naive_impl/kmer.rs:76-81 */
uint64_t ko_kmer_prepend_base_u8(ko_kmer *km, uint8_t c, int strict) {
    return ko_kmer_prepend_base(km, ko_encode_binary_u8(c), strict);
}

/* The five mask-and-shift exchange stages shared by kmer.rs:125-130,
 * kmer.rs:139-144 and hash.rs:62-67 : adjacent 2-bit pairs, then nibbles,
 * bytes, half-words, words. */
static inline uint64_t pair_reverse64(uint64_t r) {
    r = ((r >> 2) & 0x3333333333333333ull) | ((r & 0x3333333333333333ull) << 2);
    r = ((r >> 4) & 0x0F0F0F0F0F0F0F0Full) | ((r & 0x0F0F0F0F0F0F0F0Full) << 4);
    r = ((r >> 8) & 0x00FF00FF00FF00FFull) | ((r & 0x00FF00FF00FF00FFull) << 8);
    r = ((r >> 16) & 0x0000FFFF0000FFFFull) | ((r & 0x0000FFFF0000FFFFull) << 16);
    r = ((r >> 32) & 0x00000000FFFFFFFFull) | ((r & 0x00000000FFFFFFFFull) << 32);
    return r;
}

/* naive_impl/kmer.rs:138-147 : !w, exchange stages, >> 2*(32-k) */
uint64_t ko_reverse_complement_word(uint64_t w, unsigned k) {
    uint64_t res = pair_reverse64(~w);
    return res >> (2 * (32 - k));
}

/* naive_impl/kmer.rs:124-136 */
ko_kmer ko_kmer_to_reverse_complement(ko_kmer km) {
    ko_kmer rc;
    rc.k = km.k;
    rc.data = ko_reverse_complement_word(km.data, km.k);
    return rc;
}

/* #[derive(Ord, PartialOrd)] on struct {k, data} (kmer.rs:6-10): compare k
 * first, then data, both unsigned. */
int ko_kmer_cmp(ko_kmer a, ko_kmer b) {
    if (a.k != b.k) return a.k < b.k ? -1 : 1;
    if (a.data != b.data) return a.data < b.data ? -1 : 1;
    return 0;
}

/* naive_impl/kmer.rs:55-58 : *self <= rc */
int ko_kmer_is_canonical(ko_kmer km) {
    return ko_kmer_cmp(km, ko_kmer_to_reverse_complement(km)) <= 0;
}

/* naive_impl/kmer.rs:68-74 */
ko_kmer ko_kmer_to_canonical(ko_kmer km) {
    return ko_kmer_is_canonical(km) ? km : ko_kmer_to_reverse_complement(km);
}

/* naive_impl/kmer.rs:155-161 */
int ko_sub_kmer_word(uint64_t word, size_t k, size_t pos, size_t width, int strict, uint64_t *out) {
    if (!(pos < k)) return KO_PANIC;         /* kmer.rs:156 */
    if (!(pos + width <= k)) return KO_PANIC; /* kmer.rs:157 */
    uint64_t w = word >> (pos * 2);
    *out = w & ko_mask_table((unsigned)width, strict);
    return KO_OK;
}

/* naive_impl/hash.rs:4-8 (Hash for Kmer = write_u64(data)), :60-71
 * (write_u64: exchange stages without the NOT, >>= (32-k)*2), :56-58 finish */
uint64_t ko_lexhash_word(uint64_t word, unsigned k) {
    uint64_t res = pair_reverse64(word);
    res >>= (32 - k) * 2;
    return res;
}

/* ---- CanonicalKmer ---- */

/* canonical_kmer.rs:22-29 */
ko_canonical_kmer ko_ck_blank_of_size(uint8_t k) {
    ko_canonical_kmer ck;
    ck.fw.k = k;
    ck.fw.data = 0;
    ck.rc.k = k;
    ck.rc.data = UINT64_MAX;
    return ck;
}

/* canonical_kmer.rs:42-52 */
ko_canonical_kmer ko_ck_from_u64(uint64_t data, uint8_t k, int strict) {
    ko_canonical_kmer ck;
    ck.fw = ko_kmer_from_u64(data, k, strict);
    ck.rc = ko_kmer_to_reverse_complement(ck.fw);
    return ck;
}

/* canonical_kmer.rs:188-196 */
int ko_ck_from_bytes(const uint8_t *s, size_t len, ko_canonical_kmer *out) {
    if (ko_kmer_from_bytes(s, len, &out->fw) != KO_OK) return KO_PANIC;
    out->rc = ko_kmer_to_reverse_complement(out->fw);
    return KO_OK;
}

/* canonical_kmer.rs:62-65 : swaps the data words only */
void ko_ck_swap(ko_canonical_kmer *ck) {
    uint64_t t = ck->fw.data;
    ck->fw.data = ck->rc.data;
    ck->rc.data = t;
}

/* canonical_kmer.rs:67-70 */
int ko_ck_is_fw_canonical(const ko_canonical_kmer *ck) { return ck->fw.data < ck->rc.data; }

/* canonical_kmer.rs:90-94 */
uint64_t ko_ck_append_base(ko_canonical_kmer *ck, uint64_t b, int strict) {
    uint64_t r = ko_kmer_append_base(&ck->fw, b);
    ko_kmer_prepend_base(&ck->rc, ko_complement_base(b), strict);
    return r;
}

/* canonical_kmer.rs:97-101 */
uint64_t ko_ck_prepend_base(ko_canonical_kmer *ck, uint64_t b, int strict) {
    uint64_t r = ko_kmer_prepend_base(&ck->fw, b, strict);
    ko_kmer_append_base(&ck->rc, ko_complement_base(b));
    return r;
}

/* canonical_kmer.rs:72-79 */
uint64_t ko_ck_append_base_u8(ko_canonical_kmer *ck, uint8_t c, int strict) {
    return ko_ck_append_base(ck, ko_encode_binary_u8(c), strict);
}

/* canonical_kmer.rs:81-88 */
uint64_t ko_ck_prepend_base_u8(ko_canonical_kmer *ck, uint8_t c, int strict) {
    return ko_ck_prepend_base(ck, ko_encode_binary_u8(c), strict);
}

/* canonical_kmer.rs:113-119 : strict '<', tie takes rc (same value) */
uint64_t ko_ck_get_canonical_word(const ko_canonical_kmer *ck) {
    return ck->fw.data < ck->rc.data ? ck->fw.data : ck->rc.data;
}

/* canonical_kmer.rs:152-161 */
int ko_ck_get_word_equivalency(const ko_canonical_kmer *ck, uint64_t other) {
    if (ck->fw.data == other) return KO_IDENTITY_MATCH;
    if (ck->rc.data == other) return KO_TWIN_MATCH;
    return KO_NO_MATCH;
}

/* ---- CanonicalKmerIterator ---- */

/* canonical_kmer_iterator.rs:42-70 */
static void iter_find_next(ko_ck_iter *it, int32_t ii, int32_t jj) {
    int32_t i = ii + 1;
    int32_t j = jj + 1;
    int32_t seq_len = (int32_t)it->seq_len;

    for (int32_t l = j; l < seq_len; ++l) {
        uint64_t b = ko_encode_binary_u8(it->seq[l]);
        if (b < 4) {
            ko_ck_append_base(&it->km, b, it->strict);
            if ((l - it->last_invalid) >= it->k) {
                it->pos = i;
                return;
            }
        } else {
            it->last_invalid = l;
            i = l + 1;
        }
    }
    it->invalid = 1;
}

/* canonical_kmer_iterator.rs:72-83 (+ CanonicalKmerPos::new :19-26) */
void ko_iter_from_u8_slice(ko_ck_iter *it, const uint8_t *s, size_t len, uint8_t k, int strict) {
    it->seq = s;
    it->seq_len = len;
    it->km = ko_ck_blank_of_size(k);
    it->pos = -1;
    it->invalid = 0;
    it->last_invalid = -1;
    it->k = (int32_t)k;
    it->strict = strict;
    iter_find_next(it, -1, -1);
}

/* canonical_kmer_iterator.rs:88-90 */
int ko_iter_exhausted(const ko_ck_iter *it) { return it->invalid; }

/* canonical_kmer_iterator.rs:93-101 */
int ko_iter_inc(ko_ck_iter *it) {
    int32_t lpos = it->pos + it->k;
    it->invalid = it->invalid || (lpos >= (int32_t)it->seq_len);
    if (!it->invalid) iter_find_next(it, it->pos, lpos - 1);
    return !it->invalid;
}

/* canonical_kmer_iterator.rs:104-111 */
int ko_iter_inc_by(ko_ck_iter *it, size_t count) {
    int v = !it->invalid;
    while (count > 0 && v) {
        v = ko_iter_inc(it);
        count -= 1;
    }
    return v;
}

/* ======================================================================= */
/* Path E : encoding + generic Kmer<P,K,B>                                  */
/* ======================================================================= */

/* bit_field 0.10 BitArray on [P;B] (not vendored; pinned by the goldens cited
 * in the header): flat bit i = bit (i % w) of word (i / w).  With the array
 * held as its little-endian byte image that is bit (i % 8) of byte (i / 8)
 * for every word width w in {8,16,32,64,128}; 2-bit fields sit at even
 * offsets so they never straddle a byte. */
static inline uint8_t get2(const uint8_t *a, size_t bitpos) {
    return (uint8_t)((a[bitpos >> 3] >> (bitpos & 7)) & 3u);
}
static inline void set2(uint8_t *a, size_t bitpos, uint8_t v) {
    uint8_t sh = (uint8_t)(bitpos & 7);
    a[bitpos >> 3] = (uint8_t)((a[bitpos >> 3] & ~(3u << sh)) | ((v & 3u) << sh));
}

/* encoding/naive.rs:14-16 */
static inline uint8_t nuc2internal(uint8_t nuc) { return (nuc >> 1) & 3u; }

/* encoding/naive.rs:19 INTERNAL2NUC ; xor10.rs:10 BITS2NUC (same table) */
static const uint8_t k_internal2nuc[4] = {'A', 'C', 'T', 'G'};

/* encoding/naive.rs:29-39 */
uint8_t ko_rev_encoding(uint8_t enc) {
    uint8_t rev = 0;
    rev ^= (uint8_t)(0u << (6 - ((enc >> 6) * 2)));
    rev ^= (uint8_t)(1u << (6 - (((enc >> 4) & 3u) * 2)));
    rev ^= (uint8_t)(2u << (6 - (((enc >> 2) & 3u) * 2)));
    rev ^= (uint8_t)(3u << (6 - ((enc & 3u) * 2)));
    return rev;
}

/* encoding/naive.rs:78-86 ; xor10.rs:17-22 */
uint8_t ko_nuc2bits(int enc, uint8_t nuc) {
    if (enc == KO_XOR10) return (nuc >> 1) & 3u;
    unsigned index = 6 - nuc2internal(nuc) * 2;
    return (uint8_t)((((uint8_t)enc) >> index) & 3u);
}

/* encoding/naive.rs:88-96 ; xor10.rs:26-31 */
uint8_t ko_bits2nuc(int enc, uint8_t bits) {
    if (enc == KO_XOR10) return k_internal2nuc[bits & 3u];
    uint8_t rev = ko_rev_encoding((uint8_t)enc);
    return k_internal2nuc[(rev >> (6 - (bits & 3u) * 2)) & 3u];
}

/* encoding/naive.rs:98-110 ; xor10.rs:35-40 */
uint8_t ko_complement_bits(int enc, uint8_t bits) {
    if (enc == KO_XOR10) return (uint8_t)((bits ^ 2u) & 0xFFu);
    uint8_t rev = ko_rev_encoding((uint8_t)enc);
    uint8_t internal = (rev >> (6 - (bits & 3u) * 2)) & 3u;
    uint8_t comp_internal = (internal ^ 2u) & 3u;
    return (uint8_t)((((uint8_t)enc) >> (6 - comp_internal * 2)) & 3u);
}

/* encoding/naive.rs:116-124 ; xor10.rs:52-60 */
int ko_encode(int enc, const uint8_t *seq, size_t len, unsigned word_bits, size_t n_words,
              uint8_t *array_out) {
    size_t total_bits = n_words * word_bits;
    memset(array_out, 0, total_bits / 8); /* mem::zeroed(), naive.rs:117 */
    for (size_t idx = 0; idx < len; ++idx) {
        if (idx * 2 + 2 > total_bits) return KO_PANIC; /* set_bits range assert */
        set2(array_out, idx * 2, ko_nuc2bits(enc, seq[idx]));
    }
    return KO_OK;
}

/* encoding/naive.rs:126-136 ; xor10.rs:62-72 */
void ko_decode(int enc, const uint8_t *array, unsigned word_bits, size_t n_words, uint8_t *seq_out) {
    size_t n = n_words * word_bits / 2;
    for (size_t idx = 0; idx < n; ++idx) seq_out[idx] = ko_bits2nuc(enc, get2(array, idx * 2));
}

/* two-pointer exchange loop, encoding/naive.rs:138-154 ; xor10.rs:86-103 */
static int rev_comp_loop(int enc, unsigned k, size_t total_bits, uint8_t *array, int strict) {
    if (k == 0) return KO_PANIC; /* K*2-2 underflows (usize) */
    size_t i = 0;
    size_t j = (size_t)k * 2 - 2;
    if (j + 2 > total_bits) return KO_PANIC; /* get_bits range assert */
    for (;;) {
        if (!(i <= j)) break;
        uint8_t comp_i = ko_complement_bits(enc, get2(array, i));
        uint8_t comp_j = ko_complement_bits(enc, get2(array, j));
        set2(array, i, comp_j);
        set2(array, j, comp_i);
        i += 2;
        if (j < 2) {
            /* `j -= 2` on usize: only reachable for K == 1:
             * debug panics, release wraps and the next get_bits asserts. */
            if (strict) return KO_PANIC;
            break;
        }
        j -= 2;
    }
    return KO_OK;
}

/* encoding/naive.rs:138-154 ; xor10.rs:74-104 */
int ko_rev_comp(int enc, unsigned k, unsigned word_bits, size_t n_words, uint8_t *array, int strict) {
    size_t total_bits = n_words * word_bits;
    if (enc == KO_XOR10 && strict) {
        /* impl exists only for P: From<u64>, i.e. u64 / u128 (xor10.rs:50) */
        if (word_bits != 64 && word_bits != 128) return KO_PANIC;
        if (n_words == 1) {
            /* xor10.rs:75-85 : array[0].to_u64().unwrap(), five exchange
             * stages, then array[0] = (8*size_of::<P>() - kmer*2).into()
             * -- which is not a reverse complement.  Release
             * (wrapping) arithmetic. */
            uint64_t kmer;
            memcpy(&kmer, array, 8);
            if (word_bits == 128) {
                uint64_t hi;
                memcpy(&hi, array + 8, 8);
                if (hi != 0) return KO_PANIC; /* to_u64() -> None -> unwrap */
            }
            kmer = pair_reverse64(kmer);
            uint64_t v = (uint64_t)word_bits - kmer * 2ull;
            memset(array, 0, word_bits / 8);
            memcpy(array, &v, 8);
            return KO_OK;
        }
    }
    return rev_comp_loop(enc, k, total_bits, array, strict);
}

/* kmer.rs:67-69 */
size_t ko_word_for_k(unsigned word_bits, size_t k) {
    size_t bases_per_word = (size_t)word_bits / 2; /* size_of::<P>()*8/2 */
    return (bases_per_word + k - 1) / bases_per_word;
}

/* kmer.rs:41-43 */
size_t ko_num_bytes(unsigned word_bits, size_t k) {
    return ((size_t)word_bits / 8) * ko_word_for_k(word_bits, k);
}

/* kmer.rs:46-48 */
uint8_t ko_kmer_get(const uint8_t *array, size_t index) { return get2(array, index * 2); }

/* kmer.rs:50-52 : get_bits(0..=(len*2)) -- an inclusive range, 2*len+1 bits */
uint64_t ko_kmer_get_prefix(const uint8_t *array, size_t len) {
    size_t nbits = len * 2 + 1;
    uint64_t v = 0;
    for (size_t b = 0; b < nbits && b < 64; ++b)
        v |= ((uint64_t)((array[b >> 3] >> (b & 7)) & 1u)) << b;
    return v;
}

/* kmer.rs:71-91 : fixed A0 C1 G2 T3 whatever the encoder was */
void ko_bitmer_to_bytes(uint64_t mer, size_t len, uint8_t *out) {
    static const uint8_t tbl[4] = {'A', 'C', 'G', 'T'};
    uint64_t m = mer;
    for (size_t i = 0; i < (size_t)(uint32_t)len; ++i) {
        out[i] = tbl[m & 3u];
        m >>= 2;
    }
}

/* ======================================================================= */
/* batch drivers                                                            */
/* ======================================================================= */

uint64_t ko_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

void ko_generate_bases(uint64_t seed, uint64_t first_index, size_t n, uint32_t n_thresh20,
                       uint8_t *out) {
    static const uint8_t acgt[4] = {'A', 'C', 'G', 'T'};
    for (size_t i = 0; i < n; ++i) {
        uint64_t x = ko_splitmix64(seed + first_index + i);
        uint8_t c = acgt[x >> 62];
        if (((x >> 20) & 0xFFFFFu) < n_thresh20) c = 'N';
        out[i] = c;
    }
}

static inline uint64_t read_begin(const uint64_t *offsets, uint64_t fixed_len, size_t r) {
    return offsets ? offsets[r] : (uint64_t)r * fixed_len;
}
static inline uint64_t n_windows(uint64_t len, unsigned k) { return len >= k ? len - k + 1 : 0; }

uint64_t ko_count_slots(const uint64_t *offsets, size_t n_reads, uint64_t fixed_len, unsigned k) {
    uint64_t total = 0;
    for (size_t r = 0; r < n_reads; ++r)
        total += n_windows(read_begin(offsets, fixed_len, r + 1) - read_begin(offsets, fixed_len, r), k);
    return total;
}

typedef struct {
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t fixed_len;
    size_t r0, r1;
    uint64_t slot0;
    unsigned k;
    int strict;
    int faithful; /* 0 = iterator path, 1 = bench-faithful per-window path */
    uint64_t *canon_out, *hash_out, *fw_out, *rc_out;
    uint64_t *hist;
    unsigned hist_bits;
    ko_digest digest;
    int status;
} job_t;

static void fill_sentinel(job_t *jb, uint64_t from, uint64_t to) {
    for (uint64_t s = from; s < to; ++s) {
        if (jb->canon_out) jb->canon_out[s] = KO_SENTINEL;
        if (jb->hash_out) jb->hash_out[s] = KO_SENTINEL;
        if (jb->fw_out) jb->fw_out[s] = KO_SENTINEL;
        if (jb->rc_out) jb->rc_out[s] = KO_SENTINEL;
    }
}

static void run_iterator_job(job_t *jb) {
    uint64_t slot = jb->slot0;
    unsigned k = jb->k;
    unsigned hshift = jb->hist ? (2 * k - jb->hist_bits) : 0;
    for (size_t r = jb->r0; r < jb->r1; ++r) {
        uint64_t b = read_begin(jb->offsets, jb->fixed_len, r);
        uint64_t e = read_begin(jb->offsets, jb->fixed_len, r + 1);
        uint64_t w = n_windows(e - b, k);
        /* the caller loop of SURVEY 3 S4: while !exhausted { get(); inc(); };
         * positions come out increasing, the gaps between them (windows the
         * iterator skipped) are filled with the sentinel. */
        ko_ck_iter it;
        ko_iter_from_u8_slice(&it, jb->bases + b, (size_t)(e - b), (uint8_t)k, jb->strict);
        uint64_t next = 0; /* first slot of this read not yet written */
        while (!ko_iter_exhausted(&it)) {
            uint64_t canon = ko_ck_get_canonical_word(&it.km);
            uint64_t h = ko_lexhash_word(canon, k);
            uint64_t pos = (uint64_t)it.pos;
            fill_sentinel(jb, slot + next, slot + pos);
            next = pos + 1;
            uint64_t s = slot + pos;
            if (jb->canon_out) jb->canon_out[s] = canon;
            if (jb->hash_out) jb->hash_out[s] = h;
            if (jb->fw_out) jb->fw_out[s] = it.km.fw.data;
            if (jb->rc_out) jb->rc_out[s] = it.km.rc.data;
            if (jb->hist) jb->hist[h >> hshift] += 1;
            jb->digest.n_valid += 1;
            jb->digest.checksum_canon += canon;
            jb->digest.checksum_hash += h;
            ko_iter_inc(&it);
        }
        fill_sentinel(jb, slot + next, slot + w);
        slot += w;
    }
}

static void run_faithful_job(job_t *jb) {
    uint64_t slot = jb->slot0;
    unsigned k = jb->k;
    for (size_t r = jb->r0; r < jb->r1; ++r) {
        uint64_t b = read_begin(jb->offsets, jb->fixed_len, r);
        uint64_t e = read_begin(jb->offsets, jb->fixed_len, r + 1);
        uint64_t w = n_windows(e - b, k);
        /* b.windows(K).map(|x| Kmer::from(x) ...) simple_benchmark.rs:14-44 */
        for (uint64_t p = 0; p < w; ++p) {
            ko_kmer km;
            if (ko_kmer_from_bytes(jb->bases + b + p, k, &km) != KO_OK) {
                jb->status = KO_PANIC;
                return;
            }
            ko_kmer canon = ko_kmer_to_canonical(km);
            uint64_t h = ko_lexhash_word(canon.data, k);
            if (jb->canon_out) jb->canon_out[slot + p] = canon.data;
            if (jb->hash_out) jb->hash_out[slot + p] = h;
            jb->digest.n_valid += 1;
            jb->digest.checksum_canon += canon.data;
            jb->digest.checksum_hash += h;
        }
        slot += w;
    }
}

static void *job_main(void *arg) {
    job_t *jb = (job_t *)arg;
    if (jb->faithful)
        run_faithful_job(jb);
    else
        run_iterator_job(jb);
    return NULL;
}

static int run_batch(int faithful, const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                     uint64_t fixed_len, unsigned k, int strict, uint64_t *canon_out,
                     uint64_t *hash_out, uint64_t *fw_out, uint64_t *rc_out, uint64_t *hist,
                     unsigned hist_bits, ko_digest *digest, int n_threads) {
    if (k < 1 || k > 32) return KO_PANIC;
    if (hist && (hist_bits < 1 || hist_bits > 2 * k || hist_bits > 30)) return KO_PANIC;
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_reads) n_threads = n_reads ? (int)n_reads : 1;

    job_t *jobs = (job_t *)calloc((size_t)n_threads, sizeof(job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    size_t nbins = hist ? ((size_t)1 << hist_bits) : 0;
    uint64_t slot = 0;
    for (int t = 0; t < n_threads; ++t) {
        job_t *jb = &jobs[t];
        jb->bases = bases;
        jb->offsets = offsets;
        jb->fixed_len = fixed_len;
        jb->r0 = n_reads * (size_t)t / (size_t)n_threads;
        jb->r1 = n_reads * (size_t)(t + 1) / (size_t)n_threads;
        jb->slot0 = slot;
        jb->k = k;
        jb->strict = strict;
        jb->faithful = faithful;
        jb->canon_out = canon_out;
        jb->hash_out = hash_out;
        jb->fw_out = fw_out;
        jb->rc_out = rc_out;
        jb->hist_bits = hist_bits;
        jb->hist = NULL;
        if (hist) jb->hist = (t == 0) ? hist : (uint64_t *)calloc(nbins, sizeof(uint64_t));
        if (offsets) {
            for (size_t r = jb->r0; r < jb->r1; ++r) slot += n_windows(offsets[r + 1] - offsets[r], k);
        } else {
            slot += (uint64_t)(jb->r1 - jb->r0) * n_windows(fixed_len, k);
        }
    }
    if (hist) memset(hist, 0, nbins * sizeof(uint64_t));
    if (n_threads == 1) {
        job_main(&jobs[0]);
    } else {
        for (int t = 0; t < n_threads; ++t) pthread_create(&tids[t], NULL, job_main, &jobs[t]);
        for (int t = 0; t < n_threads; ++t) pthread_join(tids[t], NULL);
    }
    int status = KO_OK;
    ko_digest d = {0, 0, 0};
    for (int t = 0; t < n_threads; ++t) {
        if (jobs[t].status != KO_OK) status = jobs[t].status;
        d.n_valid += jobs[t].digest.n_valid;
        d.checksum_canon += jobs[t].digest.checksum_canon;
        d.checksum_hash += jobs[t].digest.checksum_hash;
        if (hist && t > 0) {
            for (size_t b = 0; b < nbins; ++b) hist[b] += jobs[t].hist[b];
            free(jobs[t].hist);
        }
    }
    if (digest) *digest = d;
    free(jobs);
    free(tids);
    return status;
}

int ko_extract_canonical(const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                         uint64_t fixed_len, unsigned k, int strict, uint64_t *canon_out,
                         uint64_t *hash_out, uint64_t *fw_out, uint64_t *rc_out, uint64_t *hist,
                         unsigned hist_bits, ko_digest *digest, int n_threads) {
    return run_batch(0, bases, offsets, n_reads, fixed_len, k, strict, canon_out, hash_out, fw_out,
                     rc_out, hist, hist_bits, digest, n_threads);
}

int ko_bench_windows(const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                     uint64_t fixed_len, unsigned k, uint64_t *canon_out, uint64_t *hash_out,
                     ko_digest *digest, int n_threads) {
    return run_batch(1, bases, offsets, n_reads, fixed_len, k, 0, canon_out, hash_out, NULL, NULL,
                     NULL, 0, digest, n_threads);
}

/* EXTENSION -- see header.  Built only from the pinned Path-E primitives:
 * Encoding::encode (naive.rs:116-124) and Encoding::rev_comp (naive.rs:138-154)
 * on a 2 x u64 array, followed by an unsigned 128-bit compare. */
int ko_extract_canonical_wide(const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                              uint64_t fixed_len, unsigned k, int enc, int validate,
                              uint64_t *canon_out, uint64_t *hash_out, ko_digest *digest) {
    if (k < 1 || k > 64) return KO_PANIC;
    ko_digest d = {0, 0, 0};
    uint64_t slot = 0;
    for (size_t r = 0; r < n_reads; ++r) {
        uint64_t b = read_begin(offsets, fixed_len, r);
        uint64_t e = read_begin(offsets, fixed_len, r + 1);
        uint64_t w = n_windows(e - b, k);
        for (uint64_t p = 0; p < w; ++p, ++slot) {
            const uint8_t *win = bases + b + p;
            int ok = 1;
            if (validate)
                for (unsigned i = 0; i < k; ++i)
                    if (ko_encode_binary_u8(win[i]) == KO_INVALID_BASE) ok = 0;
            uint64_t cw[2] = {KO_SENTINEL, KO_SENTINEL}, hw[2] = {KO_SENTINEL, KO_SENTINEL};
            if (ok) {
                uint8_t fw[16], rc[16], hs[16];
                if (ko_encode(enc, win, k, 64, 2, fw) != KO_OK) return KO_PANIC;
                memcpy(rc, fw, 16);
                if (k == 1) { /* rev_comp::<1> is the plain complement (Q9) */
                    set2(rc, 0, ko_complement_bits(enc, get2(fw, 0)));
                } else if (ko_rev_comp(enc, k, 64, 2, rc, 0) != KO_OK) {
                    return KO_PANIC;
                }
                uint64_t f[2], c[2];
                memcpy(f, fw, 16);
                memcpy(c, rc, 16);
                int fw_less = (f[1] < c[1]) || (f[1] == c[1] && f[0] < c[0]);
                const uint8_t *canon = fw_less ? fw : rc;
                memcpy(cw, canon, 16);
                memset(hs, 0, 16);
                for (unsigned i = 0; i < k; ++i) set2(hs, 2 * (size_t)(k - 1 - i), get2(canon, 2 * (size_t)i));
                memcpy(hw, hs, 16);
                d.n_valid += 1;
                d.checksum_canon += cw[0] + cw[1];
                d.checksum_hash += hw[0] + hw[1];
            }
            if (canon_out) { canon_out[2 * slot] = cw[0]; canon_out[2 * slot + 1] = cw[1]; }
            if (hash_out) { hash_out[2 * slot] = hw[0]; hash_out[2 * slot + 1] = hw[1]; }
        }
    }
    if (digest) *digest = d;
    return KO_OK;
}

/* the same over n_threads threads: reads are cut into contiguous ranges, every range runs the function above
 * on its own slice of the outputs (test infrastructure: gives config 3 a CPU figure on all cores) */
typedef struct {
    const uint8_t *bases; const uint64_t *offsets; size_t r0, r1; uint64_t fixed_len; unsigned k; int enc, validate;
    uint64_t *canon_out, *hash_out; uint64_t slot0; ko_digest d; int status;
} wide_job_t;

static void *wide_job_main(void *arg) {
    wide_job_t *jb = (wide_job_t *)arg;
    /* a range of a CSR batch is a CSR batch again once its offsets are rebased; fixed-length ranges just shift */
    const size_t n = jb->r1 - jb->r0;
    uint64_t *canon = jb->canon_out ? jb->canon_out + 2 * jb->slot0 : NULL;
    uint64_t *hash = jb->hash_out ? jb->hash_out + 2 * jb->slot0 : NULL;
    if (!jb->offsets) {
        jb->status = ko_extract_canonical_wide(jb->bases + jb->r0 * jb->fixed_len, NULL, n, jb->fixed_len, jb->k, jb->enc,
                                               jb->validate, canon, hash, &jb->d);
        return NULL;
    }
    uint64_t *offs = (uint64_t *)malloc((n + 1) * sizeof(uint64_t));
    if (!offs) { jb->status = KO_PANIC; return NULL; }
    for (size_t i = 0; i <= n; ++i) offs[i] = jb->offsets[jb->r0 + i] - jb->offsets[jb->r0];
    jb->status = ko_extract_canonical_wide(jb->bases + jb->offsets[jb->r0], offs, n, 0, jb->k, jb->enc, jb->validate, canon, hash, &jb->d);
    free(offs);
    return NULL;
}

int ko_extract_canonical_wide_mt(const uint8_t *bases, const uint64_t *offsets, size_t n_reads, uint64_t fixed_len, unsigned k,
                                 int enc, int validate, uint64_t *canon_out, uint64_t *hash_out, ko_digest *digest, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if ((size_t)n_threads > n_reads) n_threads = n_reads ? (int)n_reads : 1;
    if (n_threads == 1) return ko_extract_canonical_wide(bases, offsets, n_reads, fixed_len, k, enc, validate, canon_out, hash_out, digest);
    wide_job_t *jobs = (wide_job_t *)calloc((size_t)n_threads, sizeof(wide_job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)n_threads, sizeof(pthread_t));
    if (!jobs || !tids) { free(jobs); free(tids); return KO_PANIC; }
    uint64_t slot = 0;
    for (int t = 0; t < n_threads; ++t) {
        wide_job_t *jb = &jobs[t];
        jb->bases = bases; jb->offsets = offsets; jb->fixed_len = fixed_len; jb->k = k; jb->enc = enc; jb->validate = validate;
        jb->canon_out = canon_out; jb->hash_out = hash_out;
        jb->r0 = n_reads * (size_t)t / (size_t)n_threads;
        jb->r1 = n_reads * (size_t)(t + 1) / (size_t)n_threads;
        jb->slot0 = slot;
        for (size_t r = jb->r0; r < jb->r1; ++r) slot += n_windows(read_begin(offsets, fixed_len, r + 1) - read_begin(offsets, fixed_len, r), k);
    }
    for (int t = 0; t < n_threads; ++t) pthread_create(&tids[t], NULL, wide_job_main, &jobs[t]);
    for (int t = 0; t < n_threads; ++t) pthread_join(tids[t], NULL);
    ko_digest d = {0, 0, 0};
    int status = KO_OK;
    for (int t = 0; t < n_threads; ++t) {
        if (jobs[t].status != KO_OK) status = jobs[t].status;
        d.n_valid += jobs[t].d.n_valid; d.checksum_canon += jobs[t].d.checksum_canon; d.checksum_hash += jobs[t].d.checksum_hash;
    }
    free(jobs);
    free(tids);
    if (digest) *digest = d;
    return status;
}

/* ======================================================================= */
/* "next" rows: minimizers + SeqVector                                      */
/* ======================================================================= */

/* naive_impl/kmer.rs:170-191 */
int ko_minimizer_word(uint64_t word, size_t k, size_t width, unsigned hash_k, int strict, uint64_t *mmer, size_t *offset) {
    uint64_t min_mmer;
    if (width > k) return KO_PANIC; /* k - width + 1 underflows / sub_kmer_word asserts */
    if (ko_sub_kmer_word(word, k, 0, width, strict, &min_mmer) != KO_OK) return KO_PANIC; /* kmer.rs:176 */
    uint64_t min_hash = UINT64_MAX;
    size_t off = 0;
    for (size_t pos = 0; pos < k - width + 1; ++pos) {
        uint64_t m;
        if (ko_sub_kmer_word(word, k, pos, width, strict, &m) != KO_OK) return KO_PANIC;
        uint64_t h = ko_lexhash_word(m, hash_k); /* hash_one(state, mmer: u64) -> write_u64, hash.rs:10-20,60-71 */
        if (h < min_hash) { /* strict: the leftmost minimum wins */
            min_mmer = m;
            min_hash = h;
            off = pos;
        }
    }
    *mmer = min_mmer;
    *offset = off;
    return KO_OK;
}

/* naive_impl/seq_vector.rs:230-242 */
int ko_sv_from_bytes(const uint8_t *s, size_t len, uint64_t *words_out) {
    size_t nw = 0;
    for (size_t i = 0; i < len; i += 32) {
        size_t n = len - i < 32 ? len - i : 32;
        ko_kmer km;
        if (ko_kmer_from_bytes(s + i, n, &km) != KO_OK) return KO_PANIC;
        words_out[nw++] = km.data;
    }
    return KO_OK;
}

/* naive_impl/seq_vector.rs:96-99 ; simple-sds RawVector::int(bit_offset, bit_len) */
int ko_sv_get_kmer_u64(const uint64_t *words, size_t len, size_t pos, size_t k, uint64_t *out) {
    if (!(pos < len)) return KO_PANIC; /* seq_vector.rs:97 */
    size_t bit = pos * 2, nbits = k * 2;
    size_t wi = bit / 64, sh = bit % 64;
    size_t n_words = (len * 2 + 63) / 64;
    uint64_t v = words[wi] >> sh;
    if (sh != 0 && sh + nbits > 64 && wi + 1 < n_words) v |= words[wi + 1] << (64 - sh);
    if (nbits < 64) v &= (((uint64_t)1) << nbits) - 1;
    *out = v;
    return KO_OK;
}

typedef struct { uint64_t lmer, pos, hash; } dqmer_t; /* minimizers.rs:8-12 */

/* naive_impl/seq_vector/minimizers.rs:38-142 */
int ko_sv_minimizers(const uint64_t *words, size_t len, size_t k, size_t w, unsigned hash_k, uint64_t *mm_words,
                     uint64_t *mm_pos) {
    if (!(len >= k)) return KO_PANIC; /* minimizers.rs:102 */
    if (w > k || w == 0) return KO_PANIC;
    size_t cap = k - w + 3;
    dqmer_t *dq = (dqmer_t *)malloc(cap * sizeof(dqmer_t));
    size_t head = 0, n = 0; /* ring buffer: front at head */
    size_t curr = 0;
    int status = KO_OK;
#define DQ(i) dq[(head + (i)) % cap]
    /* enqueue_dqmer, minimizers.rs:60-81 */
#define ENQUEUE(m)                                                            \
    do {                                                                      \
        if (n > 0 && DQ(0).pos < curr) { head = (head + 1) % cap; n--; }      \
        while (n > 0) {                                                       \
            if (DQ(n - 1).hash <= (m).hash) break;                            \
            n--;                                                              \
        }                                                                     \
        DQ(n) = (m);                                                          \
        n++;                                                                  \
    } while (0)
    for (size_t i = 0; i < k - w; ++i) { /* minimizers.rs:114-121 : lmers of the k-1 prefix */
        dqmer_t m;
        if (ko_sv_get_kmer_u64(words, len, i, w, &m.lmer) != KO_OK) { status = KO_PANIC; goto done; }
        m.pos = i;
        m.hash = ko_lexhash_word(m.lmer, hash_k);
        ENQUEUE(m);
    }
    for (; curr < len - k + 1; ++curr) { /* Iterator::next, minimizers.rs:127-142 */
        dqmer_t m; /* next_dqmer :84-90 */
        m.pos = curr + k - w;
        if (ko_sv_get_kmer_u64(words, len, m.pos, w, &m.lmer) != KO_OK) { status = KO_PANIC; goto done; }
        m.hash = ko_lexhash_word(m.lmer, hash_k);
        ENQUEUE(m);
        mm_words[curr] = DQ(0).lmer;
        mm_pos[curr] = DQ(0).pos;
    }
#undef ENQUEUE
#undef DQ
done:
    free(dq);
    return status;
}

int ko_minimizers_batch(const uint8_t *bases, const uint64_t *offsets, size_t n_reads, uint64_t fixed_len, unsigned k,
                        unsigned w, unsigned hash_k, uint64_t *mm_out, uint32_t *pos_out) {
    if (k < 1 || k > 32 || w < 1 || w > k) return KO_PANIC;
    uint64_t slot = 0;
    for (size_t r = 0; r < n_reads; ++r) {
        uint64_t b = read_begin(offsets, fixed_len, r), e = read_begin(offsets, fixed_len, r + 1);
        uint64_t len = e - b, nwin = n_windows(len, k);
        for (uint64_t p = 0; p < nwin; ++p) { mm_out[slot + p] = KO_SENTINEL; pos_out[slot + p] = UINT32_MAX; }
        uint64_t i = 0;
        while (i < len) { /* maximal runs of valid bases */
            if (ko_encode_binary_u8(bases[b + i]) == KO_INVALID_BASE) { ++i; continue; }
            uint64_t j = i;
            while (j < len && ko_encode_binary_u8(bases[b + j]) != KO_INVALID_BASE) ++j;
            uint64_t seg = j - i;
            if (seg >= k) {
                uint64_t *words = (uint64_t *)malloc(((seg + 31) / 32) * 8);
                uint64_t *mw = (uint64_t *)malloc((seg - k + 1) * 8), *mp = (uint64_t *)malloc((seg - k + 1) * 8);
                int st = ko_sv_from_bytes(bases + b + i, seg, words);
                if (st == KO_OK) st = ko_sv_minimizers(words, seg, k, w, hash_k, mw, mp);
                if (st == KO_OK)
                    for (uint64_t q = 0; q < seg - k + 1; ++q) {
                        mm_out[slot + i + q] = mw[q];
                        pos_out[slot + i + q] = (uint32_t)(mp[q] + i);
                    }
                free(words); free(mw); free(mp);
                if (st != KO_OK) return st;
            }
            i = j;
        }
        slot += nwin;
    }
    return KO_OK;
}
