/*
 * kmers_b200.h -- C ABI of the B200-native hot path of COMBINE-lab/kmers.
 *
 * The reference is a pure-Rust crate with no FFI of its own;
 * its boundary for this path is the Rust API.  Each entry point below is the
 * *batched* form of one reference item, cited as file:line under
 * /root/reference/src.  A Rust maintainer binds these with a build.rs +
 * `extern "C"` block (INTEGRATION.md shows the stub).
 *
 * Conventions
 *   - every function returns int32_t: KMB_OK (0) or a negative KMB_ERR_*;
 *     nothing unwinds or aborts across the boundary.  kmb_last_error(ctx)
 *     returns the message of the last failure on that ctx.
 *   - where the reference would panic!/assert! (k > 32, k == 0 ...) the call
 *     returns KMB_ERR_PANIC instead.
 *   - a kmb_ctx owns one CUDA stream, pinned staging buffers and the
 *     device-resident read batch.  One ctx per (host thread, GPU); a ctx is
 *     not thread-safe; distinct ctxs are independent.
 *   - pointers named *_out / in may be DEVICE or HOST memory; the library
 *     looks the pointer up (cudaPointerGetAttributes) and stages host
 *     pointers through device scratch.  Device outputs stay resident.
 *   - there is NO CPU fallback: without a CUDA device every compute call
 *     fails with KMB_ERR_NO_DEVICE.
 *   - k-mer words: base 0 in bits 1:0 (naive_impl/kmer.rs:234-251); multi-word
 *     arrays are little-endian word order, flat bit i = bit i%w of word i/w
 *     (encoding/naive.rs:297-445 goldens).
 */
#ifndef KMERS_B200_H
#define KMERS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KMB_VERSION 100

#define KMB_OK 0
#define KMB_ERR_INVALID_ARG (-1)
#define KMB_ERR_CUDA (-2)
#define KMB_ERR_NO_DEVICE (-3)
#define KMB_ERR_STATE (-4)   /* e.g. no batch loaded */
#define KMB_ERR_PANIC (-5)   /* the reference would panic here */
#define KMB_ERR_NOMEM (-6)

/* dense-slot filler for windows CanonicalKmerIterator skips
 * (canonical_kmer_iterator.rs:55-66); unambiguous because a canonical word /
 * LexHash of k <= 31 is < 2^62 (and the top word for k <= 63). */
#define KMB_SENTINEL UINT64_MAX

/* encoding selector: a Naive discriminant byte (encoding/naive.rs:49-74) or
 * KMB_ENC_XOR10 (struct Xor10, encoding/xor10.rs:12 == Naive::ACTG). */
#define KMB_ENC_ACGT 0x1E /* Naive::ACGT == naive_impl's A0 C1 G2 T3 (naive_impl/mod.rs:21-24) */
#define KMB_ENC_ACTG 0x1B /* Naive::ACTG */
#define KMB_ENC_XOR10 0x100

/* flags for kmb_extract_canonical* */
#define KMB_F_NO_VALIDATE 0x1u /* Path-E semantics: every window is kept, bytes map by (c>>1)&3 */
/* kmb_histogram only: accumulate {n_valid, checksum_canon, checksum_hash} into hist_out[n_bins .. n_bins + 2] instead of
 * reading a digest back: [bins | digest] is then one buffer of n_bins + 3 words that the ranks all-reduce in place
 * (SURVEY 8e), and the call stays asynchronous. */
#define KMB_F_DIGEST_IN_HIST 0x2u

/* MatchType, naive_impl/canonical_kmer.rs:7-12 */
#define KMB_NO_MATCH 0
#define KMB_IDENTITY_MATCH 1
#define KMB_TWIN_MATCH 2

typedef struct kmb_ctx kmb_ctx;

/* order-independent digest of one extraction (mirrors the `.sum()` of
 * benches/simple_benchmark.rs:21, wrapping). */
typedef struct {
    uint64_t n_valid;        /* windows the reference iterator emits */
    uint64_t checksum_canon; /* wrapping sum of their canonical words (all words for k > 32) */
    uint64_t checksum_hash;  /* wrapping sum of their LexHash words */
} kmb_digest;

/* ---- library / context ------------------------------------------------ */
int32_t kmb_version(void);
/* number of CUDA devices visible, 0 if none / no driver */
int32_t kmb_device_count(void);
/* device < 0: current device.  stream == NULL: the ctx creates its own
 * non-blocking stream; otherwise it borrows the caller's cudaStream_t (so the
 * ctx can share a stream with torch / other CUDA code).  To borrow the legacy
 * default stream pass cudaStreamLegacy ((void*)1), not NULL. */
int32_t kmb_ctx_create(int32_t device, void *cuda_stream, kmb_ctx **out);
int32_t kmb_ctx_destroy(kmb_ctx *ctx);
const char *kmb_last_error(const kmb_ctx *ctx); /* ctx may be NULL: last ctx-less error of this thread */
int32_t kmb_ctx_sync(kmb_ctx *ctx);
void *kmb_ctx_stream(kmb_ctx *ctx);
/* number of kernels this ctx has launched since creation (bench accounting) */
uint64_t kmb_ctx_launch_count(const kmb_ctx *ctx);

/* raw memory helpers so an FFI caller needs no CUDA binding of its own */
int32_t kmb_device_alloc(kmb_ctx *ctx, size_t bytes, void **out);
int32_t kmb_device_free(kmb_ctx *ctx, void *ptr);
int32_t kmb_host_alloc_pinned(kmb_ctx *ctx, size_t bytes, void **out);
int32_t kmb_host_free_pinned(kmb_ctx *ctx, void *ptr);
/* async on the ctx stream, either direction; kmb_ctx_sync to wait */
int32_t kmb_memcpy(kmb_ctx *ctx, void *dst, const void *src, size_t bytes);

/* ---- device-resident read batch --------------------------------------- */
/* A batch is n_reads reads concatenated without separators.  Either
 * fixed_len > 0 and offsets == NULL (read r = bytes [r*L, (r+1)*L)), or
 * offsets has n_reads+1 ascending u64 entries (CSR), offsets[0] == 0.
 * Replaces the `&[u8]` a caller hands to CanonicalKmerIterator::from_u8_slice
 * (canonical_kmer_iterator.rs:72-83) / Encoding::encode (encoding/mod.rs:16). */

/* host -> pinned staging -> device, async on the ctx stream */
int32_t kmb_batch_upload(kmb_ctx *ctx, const uint8_t *bases, uint64_t n_bytes, const uint64_t *offsets,
                         uint64_t n_reads, uint64_t fixed_len);
/* zero-copy: borrow caller-owned DEVICE memory (must outlive the batch) */
int32_t kmb_batch_attach(kmb_ctx *ctx, const uint8_t *dev_bases, uint64_t n_bytes,
                         const uint64_t *dev_offsets, uint64_t n_reads, uint64_t fixed_len);
/* synthetic fixed-length reads generated on the device:
 * x = splitmix64(seed + first_index + i); base i = "ACGT"[x >> 62], or 'N'
 * when ((x >> 20) & 0xFFFFF) < n_thresh20. */
int32_t kmb_batch_generate(kmb_ctx *ctx, uint64_t seed, uint64_t first_index, uint64_t n_reads,
                           uint64_t fixed_len, uint32_t n_thresh20);
/* copy the resident bases back (host or device dst) */
int32_t kmb_batch_download(kmb_ctx *ctx, uint8_t *dst, uint64_t n_bytes);
int32_t kmb_batch_info(const kmb_ctx *ctx, uint64_t *n_bytes, uint64_t *n_reads, uint64_t *fixed_len);
/* dense slots for window length k: sum over reads of max(0, L_r - k + 1) */
int32_t kmb_batch_num_slots(kmb_ctx *ctx, uint32_t k, uint64_t *n_slots);
/* CSR batches: exclusive prefix of per-read window counts (n_reads+1 u64,
 * host or device dst) -- slot of window `pos` of read r = win_offsets[r] + pos */
int32_t kmb_batch_window_offsets(kmb_ctx *ctx, uint32_t k, uint64_t *win_offsets_out);

/* ---- the hot path: batched CanonicalKmerIterator ---------------------- */
/* For every read of the batch, every window position `pos` in increasing
 * order (canonical_kmer_iterator.rs:42-101): slot = win_offsets[r] + pos.
 *   canon_out[slot] = CanonicalKmer::get_canonical_word (canonical_kmer.rs:113-119)
 *   hash_out[slot]  = hash_one(&LexHasherState::new(k), canonical kmer) (hash.rs:10-20, 60-71)
 *   fw_out / rc_out = get_fw_word / get_rc_word (canonical_kmer.rs:131-139)
 * Windows the iterator skips (a byte outside ACGTacgt, naive_impl/mod.rs:40-50)
 * hold KMB_SENTINEL in every output.  Any output pointer may be NULL.
 * digest (HOST pointer, may be NULL) receives the order-independent digest;
 * passing it makes the call synchronous.  1 <= k <= 32; k == 32 uses the
 * intended all-ones mask, not MASK_TABLE[32] == 0. */
int32_t kmb_extract_canonical(kmb_ctx *ctx, uint32_t k, uint32_t flags, uint64_t *canon_out,
                              uint64_t *hash_out, uint64_t *fw_out, uint64_t *rc_out,
                              kmb_digest *digest);

/* Compacted, iterator-identical form: exactly the sequence of CanonicalKmerPos{km, pos}
 * (canonical_kmer_iterator.rs:13-16) that `while !it.exhausted() { it.get(); it.inc(); }` yields for
 * every read, reads back to back in batch order:
 *   pos_out[i] = pos (i32, as the reference), canon_out[i], hash_out[i] as in kmb_extract_canonical,
 *   emit_offsets_out[r] = index of read r's first entry (n_reads + 1 entries; [n_reads] = total).
 * *n_emitted receives the number of emitted k-mers.  Size the arrays either by calling once with every output pointer
 * NULL (a counting pass over the bases) or by giving them the worst case, one entry per slot (kmb_batch_num_slots): the
 * emit call is ONE kernel launch that reads the bases once (it counts from its staged tiles and places every tile by a
 * decoupled look-back over the earlier ones), so a caller with worst-case arrays never pays for the counting pass.
 * Nothing is written at or beyond `capacity`; if *n_emitted > capacity the call fails with KMB_ERR_INVALID_ARG after
 * reporting the needed size.  Outputs may be host or device memory; any of them may be NULL.  Synchronous. */
int32_t kmb_extract_compact(kmb_ctx *ctx, uint32_t k, uint32_t flags, uint64_t *canon_out, uint64_t *hash_out,
                            int32_t *pos_out, uint64_t *emit_offsets_out, uint64_t capacity, uint64_t *n_emitted);

/* EXTENSION (not defined by the reference, parity unpinned): 1 <= k <= 64,
 * two u64 words per slot (canon_out[2*slot], [2*slot+1]; word 1 most
 * significant), any encoding.  Built from Encoding::encode + rev_comp::<K>
 * (encoding/naive.rs:116-154) per window; canonical = unsigned 2k-bit min;
 * hash = 2k-bit pair reversal (lexicographic rank).  Validation as above
 * unless KMB_F_NO_VALIDATE. */
int32_t kmb_extract_canonical_wide(kmb_ctx *ctx, uint32_t k, int32_t enc, uint32_t flags,
                                   uint64_t *canon_out, uint64_t *hash_out, kmb_digest *digest);

/* Fused, nothing materialised: histogram of emitted windows by the top
 * hist_bits of the 2k-bit LexHash (1 << hist_bits u64 bins, device or host
 * dst, ACCUMULATED into when accumulate != 0) + digest.  Config 5 of
 * BASELINE.json; the bins are what ranks all-reduce.  With KMB_F_DIGEST_IN_HIST hist_out has 3 more words, which
 * receive the digest (digest must be NULL). */
int32_t kmb_histogram(kmb_ctx *ctx, uint32_t k, uint32_t flags, uint32_t hist_bits, uint64_t *hist_out,
                      int32_t accumulate, kmb_digest *digest);

/* One-shot end-to-end form for reads that live in HOST memory (pinned or pageable; fixed-length reads): the same
 * outputs as kmb_extract_canonical for n_reads reads of fixed_len bases at host_bases.
 *   out_canon / out_hash: DEVICE arrays for the whole batch (n_reads * (fixed_len - k + 1) words each; written in place,
 *   the results stay resident), HOST arrays of that size (copied back chunk by chunk on a stream of their own while
 *   later chunks upload and compute), or NULL (that array is not produced; both NULL = digest only).
 * The call chunks the reads and overlaps, on three streams: host-side packing of the ASCII bytes to 2 bits + 1 validity
 * bit per base by a pool of worker threads (3 bits/base cross PCIe instead of 8; for pageable input this doubles as the
 * staging copy), H2D, the extraction kernel, and the D2H of results.  When the input is pinned, chunks the packers have
 * not reached are also sent as raw ASCII whenever the link would otherwise idle (and with fewer than 6 worker threads --
 * many ranks sharing one host, whose memory system is then the bottleneck -- pinned input is not packed at all).  Only format conversion and copies run
 * on the host: every k-mer is computed by the GPU kernel.  Synchronous.  This is the device-resident read-batch buffer
 * with pinned-host staging that BASELINE.json's north_star names. */
int32_t kmb_extract_canonical_host(kmb_ctx *ctx, const uint8_t *host_bases, uint64_t n_reads,
                                   uint64_t fixed_len, uint32_t k, uint32_t flags, uint64_t *out_canon,
                                   uint64_t *out_hash, kmb_digest *digest);
/* The same for reads the caller already holds 2-bit packed in host memory: host_bits[i] = bases 16i .. 16i+15 of the
 * concatenated reads (base j at bits 2j+1:2j, A0 C1 G2 T3 -- SeqVector's bit layout, naive_impl/seq_vector.rs:230-242,
 * without per-read padding), host_inv[i] bit j = base 16i+j is not one of ACGTacgt (NULL: no invalid base, as in a
 * SeqVector).  kmb_host_pack writes both.  0.25 - 0.375 B/base cross PCIe and no host thread touches the data. */
int32_t kmb_extract_canonical_host_packed(kmb_ctx *ctx, const uint32_t *host_bits, const uint16_t *host_inv,
                                          uint64_t n_reads, uint64_t fixed_len, uint32_t k, uint32_t flags,
                                          uint64_t *out_canon, uint64_t *out_hash, kmb_digest *digest);
/* ASCII -> the packed staging format above, on the host (SIMD, single-threaded; needs no GPU): the host twin of
 * SeqVector::from(&[u8]) (seq_vector.rs:230-242) plus the validity mask a SeqVector cannot hold (naive_impl/mod.rs:40-50).
 * bits_out / inv_out receive ceil(n_bases / 16) entries; bytes outside ACGTacgt encode by (c >> 1) & 3 like
 * Encoding::encode and set their invalid bit. */
int32_t kmb_host_pack(const uint8_t *bases, uint64_t n_bases, uint32_t *bits_out, uint16_t *inv_out);
/* which implementation kmb_host_pack runs on this CPU: "avx512gfni", "avx512bw", "avx2" or "swar" */
const char *kmb_host_pack_isa(void);
/* diagnostic: time (seconds) n_threads host threads take to read n_bytes at buf once -- the floor of any path that has to
 * stream the caller's reads out of host memory (bench.py reports the e2e number against it) */
int32_t kmb_host_read_probe(const uint8_t *buf, uint64_t n_bytes, uint32_t n_threads, double *seconds_out);
/* worker threads of the host pipeline (0 = default: the CPUs the process may run on, at most 32; one process per GPU on
 * a shared host should divide the cores between the ranks) */
int32_t kmb_ctx_set_host_threads(kmb_ctx *ctx, uint32_t n_threads);
/* what the last kmb_extract_canonical_host* call did: stats4 = {chunks, chunks sent as raw ASCII, H2D bytes, D2H bytes} */
int32_t kmb_ctx_host_stats(const kmb_ctx *ctx, uint64_t *stats4);

/* ---- "next" row N1: minimizers ------------------------------------------ */
/* One (lmer, pos) per k-mer window of every read, dense slots as kmb_extract_canonical: the LEFTMOST w-mer of
 * minimum hash_one(&LexHasherState::new(hash_k), lmer) inside the k-mer -- the sequence SeqVecMinimizerIter yields
 * (naive_impl/seq_vector/minimizers.rs:38-142; ties keep the older entry, :72-78).  mmer_out[slot] = the lmer word
 * (forward strand, SeqVector::get_kmer_u64, seq_vector.rs:96-99), pos_out[slot] = its position inside the read.
 * EXTENSION: a SeqVector cannot hold non-ACGT bases; windows holding one get KMB_SENTINEL / UINT32_MAX.
 * 1 <= w <= k <= 32, 1 <= hash_k <= 32.
 * The reference's iterator is generic over the hasher (SeqVecMinimizerIter<T: BuildHasher>, minimizers.rs:38-60); the
 * hasher here is always LexHasherState(hash_k), the only one whose arithmetic is in the reference tree (std's SipHash
 * DefaultHasher is randomly keyed and cannot be pinned). */
int32_t kmb_minimizers(kmb_ctx *ctx, uint32_t k, uint32_t w, uint32_t hash_k, uint32_t flags, uint64_t *mmer_out,
                       uint32_t *pos_out);
/* Kmer::minimizer_word (naive_impl/kmer.rs:170-191) with LexHasherState(hash_k) on n k-mer words:
 * mmer_out[i] = leftmost width-mer of minimum hash, offset_out[i] = its offset inside the k-mer. */
int32_t kmb_minimizer_words(kmb_ctx *ctx, uint32_t k, uint32_t w, uint32_t hash_k, const uint64_t *words, uint64_t n,
                            uint64_t *mmer_out, uint32_t *offset_out);

/* ---- "next" row N2: the batch as a 2-bit packed sequence store ------------- */
/* SeqVector twin (naive_impl/seq_vector.rs:18-258): 32 bases per u64 word, base i of a read at bits 2i+1:2i of its
 * region, A0 C1 G2 T3 -- the layout SeqVector::from(&[u8]) builds (:230-242) and kmb_pack(KMB_ENC_ACGT, 64) writes;
 * every read starts on a word boundary.  Once the batch is packed, kmb_extract_canonical (fw_out = iter_kmers,
 * :117-124), kmb_extract_canonical_wide, kmb_extract_compact, kmb_histogram and kmb_minimizers (= iter_minimizers,
 * :126-139) read 0.25 B/base instead of 1 B/base.  A packed store holds no invalid base, so every window is emitted. */
/* pack the resident ASCII batch on the device and switch the batch to the packed copy.  strict != 0: fail with
 * KMB_ERR_PANIC if any byte is outside ACGTacgt, as SeqVector::from would panic; strict == 0: such bytes encode
 * by (c >> 1) & 3 like Encoding::encode. */
int32_t kmb_batch_repack(kmb_ctx *ctx, int32_t strict);
/* borrow caller-owned packed DEVICE memory.  Fixed-length: dev_offsets == dev_word_offsets == NULL and
 * n_words == n_reads * ceil(fixed_len / 32).  Ragged: dev_offsets = base offsets (n_reads + 1), dev_word_offsets =
 * u64-word offsets (n_reads + 1), both as kmb_pack reports them. */
int32_t kmb_batch_attach_packed(kmb_ctx *ctx, const uint64_t *dev_words, uint64_t n_words, const uint64_t *dev_offsets,
                                const uint64_t *dev_word_offsets, uint64_t n_reads, uint64_t fixed_len);
/* SeqVector::get_kmer_u64 (seq_vector.rs:96-99) for n (read, pos) pairs (reads == NULL: read 0): out[i] = the k bases
 * at pos[i] of read reads[i], or KMB_SENTINEL where the reference's assert!(pos < len) / the read's end is violated. */
int32_t kmb_packed_get_kmers(kmb_ctx *ctx, uint32_t k, const uint64_t *reads, const uint64_t *pos, uint64_t n, uint64_t *out);

/* SeqVector::with_capacity + push_chars (seq_vector.rs:135-161): an empty packed sequence owned by the context, grown by
 * appending ASCII bases (host or device memory).  A byte outside ACGTacgt is KMB_ERR_PANIC and nothing is appended (push_chars
 * goes through Kmer::from, which panics).  push_chars needs a context-owned packed batch of ONE sequence: kmb_batch_new_packed,
 * or kmb_batch_repack of a one-read batch.  Every extraction op then sees the grown sequence. */
int32_t kmb_batch_new_packed(kmb_ctx *ctx, uint64_t capacity_bases);
int32_t kmb_packed_push_chars(kmb_ctx *ctx, const uint8_t *bases, uint64_t n);
/* SeqVector::slice / SeqVectorSlice (seq_vector.rs:24-90): turn the resident packed batch into a one-read VIEW of bases
 * [start, start + len) of `read` -- no copy; kmb_extract_canonical (fw_out = SeqVectorSlice::iter_kmers), kmb_minimizers
 * (= iter_minimizers), kmb_extract_compact, kmb_histogram and kmb_packed_get_kmers (= SeqVectorSlice::get_kmer_u64) then
 * work on the view.  start / len are always counted in the parent read, also when a view is already in place (which is what
 * the reference's nested slice() does: it stores `start` as the new start_pos, seq_vector.rs:57-64).  A range beyond the read
 * is KMB_ERR_PANIC (assert!(end <= self.len())).  kmb_batch_unslice restores the whole batch. */
int32_t kmb_batch_slice(kmb_ctx *ctx, uint64_t read, uint64_t start, uint64_t len);
int32_t kmb_batch_unslice(kmb_ctx *ctx);

/* ---- final reduction across GPUs ------------------------------- */
/* One process driving several GPUs: in-place element-wise wrapping-u64 sum of dev_bufs[i] (count words in the memory of
 * ctxs[i]'s GPU, e.g. [histogram | n_valid | checksum_canon | checksum_hash]) over the n_ctx contexts -- NCCL
 * ncclAllReduce(ncclUint64, ncclSum) over NVLink, enqueued on each context's stream behind its kernels; returns after
 * every stream has drained.  One context per GPU.  libnccl.so.2 is loaded on first use (KMB_ERR_STATE if absent).  Hosts
 * that run one process per GPU reduce with their own communicator instead (kmers_b200/dist.py). */
int32_t kmb_allreduce_u64(kmb_ctx *const *ctxs, int32_t n_ctx, uint64_t *const *dev_bufs, uint64_t count);

/* ---- "next" row N4: host ingest -------------------------------------------- */
/* FASTA ('>' records, multi-line sequences) or FASTQ ('@' four-line records) text in host memory -> the
 * concatenated bases and their CSR offsets (n_reads + 1 entries), exactly the arguments of kmb_batch_upload.
 * Bases are copied verbatim (case, N, IUPAC kept; line ends and '\r' dropped).  Call with bases_out == offsets_out ==
 * NULL to size the outputs.  Pure host code, needs no GPU.  Errors (malformed input) return KMB_ERR_INVALID_ARG with
 * kmb_last_error(NULL). */
int32_t kmb_parse_fastx(const char *text, uint64_t n_bytes, uint8_t *bases_out, uint64_t bases_cap, uint64_t *offsets_out,
                        uint64_t reads_cap, uint64_t *n_reads, uint64_t *n_bases);
/* parse into pinned host memory and upload as the context's (ragged) read batch */
int32_t kmb_batch_ingest_fastx(kmb_ctx *ctx, const char *text, uint64_t n_bytes, uint64_t *n_reads_out, uint64_t *n_bases_out);

/* ---- batched Encoding<P,B> (encoding/mod.rs:14-23) --------------------- */
/* Encoding::encode of every read of the batch (encoding/naive.rs:116-124,
 * xor10.rs:52-60): read r becomes ceil(L_r / (word_bits/2)) words of
 * word_bits in {8,16,32,64,128}, unused high bits 0, written from byte
 * word_offsets[r] * word_bits/8 of words_out.  No validation.
 * word_offsets_out (n_reads+1 u64, may be NULL) receives the CSR in words. */
int32_t kmb_pack(kmb_ctx *ctx, int32_t enc, uint32_t word_bits, void *words_out, uint64_t *word_offsets_out);
int32_t kmb_pack_num_words(kmb_ctx *ctx, uint32_t word_bits, uint64_t *n_words);
/* Encoding::decode (encoding/naive.rs:126-136): n_items arrays of
 * words_per_item words each -> bases_per_item ASCII bytes each, upper case.
 * bases_per_item == words_per_item*word_bits/2 reproduces the reference's
 * padding positions. */
int32_t kmb_unpack(kmb_ctx *ctx, int32_t enc, uint32_t word_bits, const void *words_in, uint64_t n_items,
                   uint32_t words_per_item, uint32_t bases_per_item, uint8_t *bases_out);
/* `impl From<Kmer> for String` (naive_impl/kmer.rs:196-207) on n k-mer words: k lower-case letters each (BASE_TABLE,
 * kmer.rs:24), base 0 first, written back to back (no terminators).  1 <= k <= 32. */
int32_t kmb_words_to_strings(kmb_ctx *ctx, uint32_t k, const uint64_t *words, uint64_t n, uint8_t *bases_out);
/* Encoding::rev_comp::<K> (encoding/naive.rs:138-154; xor10.rs:86-103 -- the
 * swap-loop result for every B, NOT the arithmetic of xor10.rs:75-85, SURVEY
 * Q2) on n_items arrays of words_per_item words; bits >= 2k are preserved.
 * k >= 1 (k == 1 is the plain complement; the reference underflows, Q9).
 * in == out allowed. */
int32_t kmb_revcomp_words(kmb_ctx *ctx, int32_t enc, uint32_t k, uint32_t word_bits, uint32_t words_per_item,
                          const void *words_in, void *words_out, uint64_t n_items);

/* ---- batched naive_impl::Kmer word ops (u64, k <= 32) ------------------- */
/* Kmer::get_reverse_complement_word (naive_impl/kmer.rs:138-147) */
int32_t kmb_reverse_complement_words(kmb_ctx *ctx, uint32_t k, const uint64_t *in, uint64_t *out, uint64_t n);
/* Kmer::to_canonical (kmer.rs:68-74); is_canonical_out (u8, may be NULL) = Kmer::is_canonical (kmer.rs:55-58) */
int32_t kmb_canonical_words(kmb_ctx *ctx, uint32_t k, const uint64_t *in, uint64_t *canon_out,
                            uint8_t *is_canonical_out, uint64_t n);
/* hash_one(&LexHasherState::new(k), Kmer) (naive_impl/hash.rs:10-20, 60-71) */
int32_t kmb_lexhash_words(kmb_ctx *ctx, uint32_t k, const uint64_t *in, uint64_t *out, uint64_t n);
/* CanonicalKmer::from_u64(words[i], k).get_word_equivalency(others[i])
 * (canonical_kmer.rs:42-52, 152-161) -> KMB_*_MATCH as u8 */
int32_t kmb_match_words(kmb_ctx *ctx, uint32_t k, const uint64_t *words, const uint64_t *others,
                        uint8_t *match_out, uint64_t n);


/* ---- the small accessors of naive_impl::Kmer / CanonicalKmer and kmer::Kmer<P,K,B>, batched ------------------------- */
/* Kmer::sub_kmer_word (naive_impl/kmer.rs:150-161): out[i] = (in[i] >> 2 pos) & MASK_TABLE[width]; pos < k and
 * pos + width <= k or KMB_ERR_PANIC (the reference asserts).  width == 32 uses the intended all-ones mask. */
int32_t kmb_sub_kmer_words(kmb_ctx *ctx, uint32_t k, uint32_t pos, uint32_t width, const uint64_t *in, uint64_t *out, uint64_t n);
/* Kmer::append_base / append_base_u8 (naive_impl/kmer.rs:83-102; slide right along the read): out[i] = (in[i] >> 2) |
 * (c << (2k - 2)), dropped_out[i] = the base shifted off (low two bits of in[i]).  bases[i] is a 2-bit code (Base), or a
 * letter when bases_are_ascii != 0 -- encoded by encode_binary_u8 WITHOUT a guard, exactly like the reference: a byte
 * outside ACGTacgt ORs u64::MAX << (2k - 2) into the word.  out / dropped_out may be NULL. */
int32_t kmb_append_base_words(kmb_ctx *ctx, uint32_t k, const uint64_t *in, const uint8_t *bases, int32_t bases_are_ascii,
                              uint64_t *out, uint8_t *dropped_out, uint64_t n);
/* Kmer::prepend_base / prepend_base_u8 (naive_impl/kmer.rs:76-95): out[i] = MASK_TABLE[k] & ((in[i] << 2) | c),
 * dropped_out[i] = the top base of in[i].  k == 32 uses the intended all-ones mask. */
int32_t kmb_prepend_base_words(kmb_ctx *ctx, uint32_t k, const uint64_t *in, const uint8_t *bases, int32_t bases_are_ascii,
                               uint64_t *out, uint8_t *dropped_out, uint64_t n);
/* CanonicalKmer::append_base[_u8] / prepend_base[_u8] (naive_impl/canonical_kmer.rs:70-100) on (fw, rc) word pairs:
 * append: fw.append_base(b), rc.prepend_base(complement_base(b)); prepend the other way round; dropped_out = what fw lost. */
int32_t kmb_canonical_append_base_words(kmb_ctx *ctx, uint32_t k, const uint64_t *fw_in, const uint64_t *rc_in, const uint8_t *bases,
                                        int32_t bases_are_ascii, uint64_t *fw_out, uint64_t *rc_out, uint8_t *dropped_out, uint64_t n);
int32_t kmb_canonical_prepend_base_words(kmb_ctx *ctx, uint32_t k, const uint64_t *fw_in, const uint64_t *rc_in, const uint8_t *bases,
                                         int32_t bases_are_ascii, uint64_t *fw_out, uint64_t *rc_out, uint8_t *dropped_out, uint64_t n);
/* CanonicalKmer::is_fw_canonical (canonical_kmer.rs:67-69): out[i] = fw[i] < rc[i].  (CanonicalKmer::swap, :62-65, is the
 * caller exchanging its two pointers; Kmer::orientation, naive_impl/kmer.rs:60-66, is kmb_canonical_words' is_canonical_out:
 * 1 = IsCanonical, 0 = NotCanononical.) */
int32_t kmb_is_fw_canonical_words(kmb_ctx *ctx, const uint64_t *fw, const uint64_t *rc, uint8_t *out, uint64_t n);
/* kmer::Kmer<P,K,B>::get (kmer.rs:46-48): codes_out[i] = the 2-bit field `index` of array i (arrays = n_items byte images of
 * words_per_item words of word_bits, as kmb_pack writes them).  An index beyond the array is KMB_ERR_PANIC (get_bits asserts). */
int32_t kmb_kmer_get(kmb_ctx *ctx, uint32_t word_bits, uint32_t words_per_item, const void *arrays, uint64_t n_items, uint32_t index,
                     uint8_t *codes_out);
/* kmer::Kmer<P,K,B>::get_prefix (kmer.rs:50-52): words_out[i] (one word of word_bits) = bits 0 ..= 2 len of array i -- the
 * reference's INCLUSIVE range, 2 len + 1 bits, matched as it is.  More bits than a P holds: KMB_ERR_PANIC. */
int32_t kmb_kmer_get_prefix(kmb_ctx *ctx, uint32_t word_bits, uint32_t words_per_item, const void *arrays, uint64_t n_items, uint32_t len,
                            void *words_out);
/* bitmer_to_bytes (kmer.rs:71-91) on n u64 words: len upper-case letters each, base 0 first, A0 C1 G2 T3 whatever encoder
 * produced the word (the reference hard-codes the table).  len <= 32. */
int32_t kmb_bitmer_to_bytes(kmb_ctx *ctx, uint32_t len, const uint64_t *mers, uint64_t n, uint8_t *bases_out);

#ifdef __cplusplus
}
#endif
#endif
