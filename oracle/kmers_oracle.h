/*
 * kmers_oracle.h -- CPU ORACLE (test infrastructure, NOT the product).
 *
 * A plain-C restatement of the CPU algorithms of COMBINE-lab/kmers for the
 * hot path named in BASELINE.json (pack / window extraction / reverse
 * complement / canonical-min / LexHasher).  Every function cites the
 * reference file:line (paths relative to /root/reference/src) it follows.
 *
 * PARITY STATUS: PINNED.  The reference is a Rust crate and there is no Rust
 * toolchain in this image, so it cannot be compiled into oracle/_ref; the
 * oracle is instead checked against every golden vector / known-answer test
 * the reference's own unit tests hold for this path (tests/test_oracle_golden.py
 * transcribes them with file:line).  Two corners are UNPINNED because the
 * reference itself does not pin them: (1) Xor10 single-word rev_comp
 * (encoding/xor10.rs:75-85, all tests commented out) and (2) canonical-min /
 * hash for K > 32 (not defined by the reference; "extension" below).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.  The product
 * (kmers_b200/, include/kmers_b200.h) never links or calls it.
 */
#ifndef KMERS_ORACLE_H
#define KMERS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KO_OK 0
#define KO_PANIC (-1) /* the reference would panic!/assert! here */

#define KO_INVALID_BASE UINT64_MAX /* naive_impl/mod.rs:48 */
#define KO_SENTINEL UINT64_MAX     /* dense-slot filler for skipped windows */

/* ---------------- Path N: naive_impl (u64 k-mers, K <= 32) ---------------- */

/* naive_impl/mod.rs:40-50 */
uint64_t ko_encode_binary_u8(uint8_t c);
/* naive_impl/mod.rs:27-37 (panics on non-ACGT -> KO_PANIC) */
int ko_encode_binary(uint8_t c, uint64_t *out);
/* naive_impl/mod.rs:81-84 */
uint64_t ko_complement_base(uint64_t b);
/* naive_impl/mod.rs:87-89 */
int ko_is_valid_nuc(uint64_t b);
/* naive_impl/kmer.rs:584-618; strict!=0 reproduces MASK_TABLE[32]==0 */
uint64_t ko_mask_table(unsigned k, int strict);

/* naive_impl/kmer.rs:6-10 */
typedef struct {
    uint8_t k;
    uint64_t data;
} ko_kmer;

/* naive_impl/kmer.rs:234-251 (From<&[u8]>); >32 bases or non-ACGT -> KO_PANIC */
int ko_kmer_from_bytes(const uint8_t *s, size_t len, ko_kmer *out);
/* naive_impl/kmer.rs:45-48 */
ko_kmer ko_kmer_from_u64(uint64_t data, uint8_t k, int strict);
/* naive_impl/kmer.rs:196-207 (From<Kmer> for String), lower-case acgt */
void ko_kmer_to_string(ko_kmer km, char *out /* k+1 bytes */);
/* naive_impl/kmer.rs:98-102 / 91-95 ; return the shifted-off base */
uint64_t ko_kmer_append_base(ko_kmer *km, uint64_t c);
uint64_t ko_kmer_prepend_base(ko_kmer *km, uint64_t c, int strict);
/* naive_impl/kmer.rs:83-88 / 76-81 (no guard on invalid bytes) */
uint64_t ko_kmer_append_base_u8(ko_kmer *km, uint8_t c);
uint64_t ko_kmer_prepend_base_u8(ko_kmer *km, uint8_t c, int strict);
/* naive_impl/kmer.rs:138-147 */
uint64_t ko_reverse_complement_word(uint64_t w, unsigned k);
/* naive_impl/kmer.rs:124-136 */
ko_kmer ko_kmer_to_reverse_complement(ko_kmer km);
/* naive_impl/kmer.rs:55-58 ; derived Ord on (k, data) kmer.rs:6 */
int ko_kmer_is_canonical(ko_kmer km);
/* naive_impl/kmer.rs:68-74 */
ko_kmer ko_kmer_to_canonical(ko_kmer km);
/* derived Ord: -1 / 0 / +1 comparing (k, data) */
int ko_kmer_cmp(ko_kmer a, ko_kmer b);
/* naive_impl/kmer.rs:155-161 ; assert failures -> KO_PANIC */
int ko_sub_kmer_word(uint64_t word, size_t k, size_t pos, size_t width, int strict, uint64_t *out);

/* naive_impl/hash.rs:60-71 + finish :56-58, i.e. hash_one(&LexHasherState(k), kmer) */
uint64_t ko_lexhash_word(uint64_t word, unsigned k);

/* naive_impl/canonical_kmer.rs:14-18 */
typedef struct {
    ko_kmer fw;
    ko_kmer rc;
} ko_canonical_kmer;

#define KO_NO_MATCH 0       /* canonical_kmer.rs:8-12 MatchType::NoMatch */
#define KO_IDENTITY_MATCH 1 /* MatchType::IdentityMatch */
#define KO_TWIN_MATCH 2     /* MatchType::TwinMatch */

/* canonical_kmer.rs:22-29 */
ko_canonical_kmer ko_ck_blank_of_size(uint8_t k);
/* canonical_kmer.rs:42-52 */
ko_canonical_kmer ko_ck_from_u64(uint64_t data, uint8_t k, int strict);
/* canonical_kmer.rs:188-196 (From<&[u8]>) */
int ko_ck_from_bytes(const uint8_t *s, size_t len, ko_canonical_kmer *out);
/* canonical_kmer.rs:62-65 */
void ko_ck_swap(ko_canonical_kmer *ck);
/* canonical_kmer.rs:67-70 */
int ko_ck_is_fw_canonical(const ko_canonical_kmer *ck);
/* canonical_kmer.rs:90-94 / 97-101 */
uint64_t ko_ck_append_base(ko_canonical_kmer *ck, uint64_t b, int strict);
uint64_t ko_ck_prepend_base(ko_canonical_kmer *ck, uint64_t b, int strict);
/* canonical_kmer.rs:72-79 / 81-88 */
uint64_t ko_ck_append_base_u8(ko_canonical_kmer *ck, uint8_t c, int strict);
uint64_t ko_ck_prepend_base_u8(ko_canonical_kmer *ck, uint8_t c, int strict);
/* canonical_kmer.rs:113-119 */
uint64_t ko_ck_get_canonical_word(const ko_canonical_kmer *ck);
/* canonical_kmer.rs:152-161 */
int ko_ck_get_word_equivalency(const ko_canonical_kmer *ck, uint64_t other);

/* naive_impl/canonical_kmer_iterator.rs:13-39 */
typedef struct {
    const uint8_t *seq;
    size_t seq_len;
    ko_canonical_kmer km;
    int32_t pos;
    int invalid;
    int32_t last_invalid;
    int32_t k;
    int strict;
} ko_ck_iter;

/* canonical_kmer_iterator.rs:72-83 */
void ko_iter_from_u8_slice(ko_ck_iter *it, const uint8_t *s, size_t len, uint8_t k, int strict);
/* canonical_kmer_iterator.rs:88-90 */
int ko_iter_exhausted(const ko_ck_iter *it);
/* canonical_kmer_iterator.rs:93-101 ; returns !invalid */
int ko_iter_inc(ko_ck_iter *it);
/* canonical_kmer_iterator.rs:104-111 */
int ko_iter_inc_by(ko_ck_iter *it, size_t count);

/* ---------------- Path E: encoding + generic Kmer<P,K,B> ---------------- */
/* Arrays [P;B] are passed as their little-endian byte image (B*word_bits/8
 * bytes).  bit_field 0.10 BitArray semantics (un-vendored dependency, pinned by
 * the numeric goldens encoding/naive.rs:300,321,342,363,394,425): flat bit i
 * lives in word i / word_bits at bit i % word_bits. */

#define KO_XOR10 0x100 /* selects struct Xor10 instead of a Naive discriminant */

/* encoding/naive.rs:78-86 ; xor10.rs:17-22 */
uint8_t ko_nuc2bits(int enc, uint8_t nuc);
/* encoding/naive.rs:88-96 ; xor10.rs:26-31 */
uint8_t ko_bits2nuc(int enc, uint8_t bits);
/* encoding/naive.rs:98-110 ; xor10.rs:35-40 */
uint8_t ko_complement_bits(int enc, uint8_t bits);
/* encoding/naive.rs:29-39 */
uint8_t ko_rev_encoding(uint8_t enc);

/* encoding/naive.rs:116-124 ; xor10.rs:52-60.  KO_PANIC when the sequence
 * does not fit (bit_field set_bits range assert). */
int ko_encode(int enc, const uint8_t *seq, size_t len, unsigned word_bits, size_t n_words,
              uint8_t *array_out);
/* encoding/naive.rs:126-136 ; xor10.rs:62-72.  Emits n_words*word_bits/2
 * bytes (padding positions included). */
void ko_decode(int enc, const uint8_t *array, unsigned word_bits, size_t n_words, uint8_t *seq_out);
/* encoding/naive.rs:138-154 ; xor10.rs:74-104.  In place.  strict!=0
 * reproduces the Xor10 B==1 arithmetic of xor10.rs:75-85 (release-mode
 * wrapping) and the K==1 underflow panic; strict==0 gives the swap-loop
 * ("intended") result for every case. */
int ko_rev_comp(int enc, unsigned k, unsigned word_bits, size_t n_words, uint8_t *array, int strict);

/* kmer.rs:67-69 */
size_t ko_word_for_k(unsigned word_bits, size_t k);
/* kmer.rs:41-43 */
size_t ko_num_bytes(unsigned word_bits, size_t k);
/* kmer.rs:46-48 */
uint8_t ko_kmer_get(const uint8_t *array, size_t index);
/* kmer.rs:50-52 : bits 0..=(2*len), i.e. 2*len+1 bits, as u64 (len*2 < 64) */
uint64_t ko_kmer_get_prefix(const uint8_t *array, size_t len);
/* kmer.rs:71-91 */
void ko_bitmer_to_bytes(uint64_t mer, size_t len, uint8_t *out);

/* ---------------- batch drivers (dense-slot layout of SURVEY 8d) ---------------- */

/* Synthetic reads (counter-based, so host and device produce identical bytes
 * with no transfer): x = splitmix64(seed + first_index + i); base =
 * "ACGT"[x >> 62]; the base becomes 'N' when ((x >> 20) & 0xFFFFF) < n_thresh20
 * (n_thresh20 = 0 -> pure ACGT; 1049 ~ 0.1 %). */
uint64_t ko_splitmix64(uint64_t x);
void ko_generate_bases(uint64_t seed, uint64_t first_index, size_t n, uint32_t n_thresh20,
                       uint8_t *out);

/* Number of dense slots: sum over reads of max(0, L_r - k + 1).  offsets has
 * n_reads+1 entries, or is NULL for fixed_len reads. */
uint64_t ko_count_slots(const uint64_t *offsets, size_t n_reads, uint64_t fixed_len, unsigned k);

typedef struct {
    uint64_t n_valid;        /* windows the iterator emits */
    uint64_t checksum_canon; /* wrapping sum of canonical words of emitted windows */
    uint64_t checksum_hash;  /* wrapping sum of LexHash of emitted windows */
} ko_digest;

/* CanonicalKmerIterator + get_canonical_word + hash_one(LexHasherState(k))
 * over every read (canonical_kmer_iterator.rs:42-101, canonical_kmer.rs:113-119,
 * hash.rs:60-71).  Slot (win_off[r] + pos) receives the emitted window, every
 * other slot KO_SENTINEL.  canon_out / hash_out / fw_out / rc_out may be NULL.
 * hist (1<<hist_bits u64 bins, may be NULL) counts emitted windows by the top
 * hist_bits of the 2k-bit LexHash.  Multi-threaded over reads when n_threads>1
 * (host plumbing only; each read is processed by the sequential iterator). */
int ko_extract_canonical(const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                         uint64_t fixed_len, unsigned k, int strict, uint64_t *canon_out,
                         uint64_t *hash_out, uint64_t *fw_out, uint64_t *rc_out, uint64_t *hist,
                         unsigned hist_bits, ko_digest *digest, int n_threads);

/* The reference's own bench workload (benches/simple_benchmark.rs:14-44),
 * bench-faithful: per window Kmer::from(&[u8]) (O(K) re-encode) +
 * to_reverse_complement + canonical min + LexHash; the sums fold every result
 * so nothing is optimised away.  Input must be pure ACGT
 * (encode_binary panics otherwise -> KO_PANIC). */
int ko_bench_windows(const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                     uint64_t fixed_len, unsigned k, uint64_t *canon_out, uint64_t *hash_out,
                     ko_digest *digest, int n_threads);

/* EXTENSION (parity unpinned; not defined by the reference): canonical-min
 * for 32 < K <= 64 as two u64 words, little-endian word order, built from
 * Encoding::encode + Encoding::rev_comp (encoding/naive.rs:116-154) of each
 * window, compare = unsigned 2K-bit integer, word 1 most significant.
 * hash = 2K-bit pair-reversal (lexicographic rank), two words.  Windows are
 * the iterator's (skip any window holding a non-ACGTacgt byte) when
 * validate!=0, else every window.  Works for k <= 32 too (one word used,
 * second word 0 for canon / hash). */
int ko_extract_canonical_wide(const uint8_t *bases, const uint64_t *offsets, size_t n_reads,
                              uint64_t fixed_len, unsigned k, int enc, int validate,
                              uint64_t *canon_out /* 2 words per slot */,
                              uint64_t *hash_out /* 2 words per slot, may be NULL */,
                              ko_digest *digest);
/* the same on n_threads threads (reads cut into contiguous ranges) */
int ko_extract_canonical_wide_mt(const uint8_t *bases, const uint64_t *offsets, size_t n_reads, uint64_t fixed_len, unsigned k,
                                 int enc, int validate, uint64_t *canon_out, uint64_t *hash_out, ko_digest *digest, int n_threads);

/* ---------------- "next" rows: minimizers + packed sequence store ---------------- */

/* Kmer::minimizer_word (naive_impl/kmer.rs:170-191) with state = LexHasherState::new(hash_k):
 * the first (leftmost) width-mer of minimum hash_one(state, mmer: u64).  KO_PANIC on the
 * sub_kmer_word asserts (pos < k, pos + width <= k) or width > k. */
int ko_minimizer_word(uint64_t word, size_t k, size_t width, unsigned hash_k, int strict, uint64_t *mmer, size_t *offset);

/* SeqVector::from(&[u8]) (naive_impl/seq_vector.rs:230-242): 32 bases per u64 through Kmer::from
 * (panics on non-ACGT -> KO_PANIC).  words_out has ceil(len / 32) entries. */
int ko_sv_from_bytes(const uint8_t *s, size_t len, uint64_t *words_out);
/* SeqVector::get_kmer_u64 (seq_vector.rs:96-99) = RawVector::int(2*pos, 2*k) of simple-sds (un-vendored git
 * dependency; semantics pinned by seq_vector.rs:304-321 and minimizers.rs:221-290): 2k bits from bit 2*pos,
 * LSB first across the u64 words.  KO_PANIC on assert!(pos < len). */
int ko_sv_get_kmer_u64(const uint64_t *words, size_t len, size_t pos, size_t k, uint64_t *out);
/* SeqVecMinimizerIter (naive_impl/seq_vector/minimizers.rs:38-142), the monotone-deque algorithm restated
 * literally, hash = hash_one(LexHasherState(hash_k), lmer: u64).  Writes len - k + 1 (word, pos) pairs.
 * KO_PANIC on assert!(sv.len() >= k). */
int ko_sv_minimizers(const uint64_t *words, size_t len, size_t k, size_t w, unsigned hash_k, uint64_t *mm_words,
                     uint64_t *mm_pos);
/* Batch driver, dense slots like ko_extract_canonical: slot of window `pos` of read r holds the minimizer
 * (lmer word, position inside the read) of that k-mer.  EXTENSION: a SeqVector cannot hold non-ACGT bases;
 * here every maximal run of valid bases is one SeqVector, and windows holding an invalid base get
 * KO_SENTINEL / UINT32_MAX. */
int ko_minimizers_batch(const uint8_t *bases, const uint64_t *offsets, size_t n_reads, uint64_t fixed_len, unsigned k,
                        unsigned w, unsigned hash_k, uint64_t *mm_out, uint32_t *pos_out);

#ifdef __cplusplus
}
#endif
#endif
