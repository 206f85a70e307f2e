//! UNVERIFIED SOURCE (never compiled: no Rust toolchain in the build image).  `extern "C"` view of include/kmers_b200.h,
//! one line per entry point; the block is GENERATED from the header by scripts/gen_rust_sys.py and
//! tests/test_abi.py fails when the two drift apart.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_void};

pub const KMB_OK: i32 = 0;
pub const KMB_ERR_INVALID_ARG: i32 = -1;
pub const KMB_ERR_CUDA: i32 = -2;
pub const KMB_ERR_NO_DEVICE: i32 = -3;
pub const KMB_ERR_STATE: i32 = -4;
pub const KMB_ERR_PANIC: i32 = -5;
pub const KMB_ERR_NOMEM: i32 = -6;
pub const KMB_SENTINEL: u64 = u64::MAX;
pub const KMB_ENC_XOR10: i32 = 0x100;
pub const KMB_F_NO_VALIDATE: u32 = 1;
pub const KMB_F_DIGEST_IN_HIST: u32 = 2;

#[repr(C)]
pub struct kmb_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug, PartialEq, Eq)]
pub struct kmb_digest {
    pub n_valid: u64,
    pub checksum_canon: u64,
    pub checksum_hash: u64,
}

extern "C" {
    // BEGIN GENERATED (scripts/gen_rust_sys.py)
    pub fn kmb_version() -> i32;
    pub fn kmb_device_count() -> i32;
    pub fn kmb_ctx_create(device: i32, cuda_stream: *mut c_void, out: *mut *mut kmb_ctx) -> i32;
    pub fn kmb_ctx_destroy(ctx: *mut kmb_ctx) -> i32;
    pub fn kmb_last_error(ctx: *const kmb_ctx) -> *const c_char;
    pub fn kmb_ctx_sync(ctx: *mut kmb_ctx) -> i32;
    pub fn kmb_ctx_stream(ctx: *mut kmb_ctx) -> *mut c_void;
    pub fn kmb_ctx_launch_count(ctx: *const kmb_ctx) -> u64;
    pub fn kmb_device_alloc(ctx: *mut kmb_ctx, bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn kmb_device_free(ctx: *mut kmb_ctx, ptr: *mut c_void) -> i32;
    pub fn kmb_host_alloc_pinned(ctx: *mut kmb_ctx, bytes: usize, out: *mut *mut c_void) -> i32;
    pub fn kmb_host_free_pinned(ctx: *mut kmb_ctx, ptr: *mut c_void) -> i32;
    pub fn kmb_memcpy(ctx: *mut kmb_ctx, dst: *mut c_void, src: *const c_void, bytes: usize) -> i32;
    pub fn kmb_batch_upload(ctx: *mut kmb_ctx, bases: *const u8, n_bytes: u64, offsets: *const u64, n_reads: u64, fixed_len: u64) -> i32;
    pub fn kmb_batch_attach(ctx: *mut kmb_ctx, dev_bases: *const u8, n_bytes: u64, dev_offsets: *const u64, n_reads: u64, fixed_len: u64) -> i32;
    pub fn kmb_batch_generate(ctx: *mut kmb_ctx, seed: u64, first_index: u64, n_reads: u64, fixed_len: u64, n_thresh20: u32) -> i32;
    pub fn kmb_batch_download(ctx: *mut kmb_ctx, dst: *mut u8, n_bytes: u64) -> i32;
    pub fn kmb_batch_info(ctx: *const kmb_ctx, n_bytes: *mut u64, n_reads: *mut u64, fixed_len: *mut u64) -> i32;
    pub fn kmb_batch_num_slots(ctx: *mut kmb_ctx, k: u32, n_slots: *mut u64) -> i32;
    pub fn kmb_batch_window_offsets(ctx: *mut kmb_ctx, k: u32, win_offsets_out: *mut u64) -> i32;
    pub fn kmb_extract_canonical(ctx: *mut kmb_ctx, k: u32, flags: u32, canon_out: *mut u64, hash_out: *mut u64, fw_out: *mut u64, rc_out: *mut u64, digest: *mut kmb_digest) -> i32;
    pub fn kmb_extract_compact(ctx: *mut kmb_ctx, k: u32, flags: u32, canon_out: *mut u64, hash_out: *mut u64, pos_out: *mut i32, emit_offsets_out: *mut u64, capacity: u64, n_emitted: *mut u64) -> i32;
    pub fn kmb_extract_canonical_wide(ctx: *mut kmb_ctx, k: u32, enc: i32, flags: u32, canon_out: *mut u64, hash_out: *mut u64, digest: *mut kmb_digest) -> i32;
    pub fn kmb_histogram(ctx: *mut kmb_ctx, k: u32, flags: u32, hist_bits: u32, hist_out: *mut u64, accumulate: i32, digest: *mut kmb_digest) -> i32;
    pub fn kmb_extract_canonical_host(ctx: *mut kmb_ctx, host_bases: *const u8, n_reads: u64, fixed_len: u64, k: u32, flags: u32, out_canon: *mut u64, out_hash: *mut u64, digest: *mut kmb_digest) -> i32;
    pub fn kmb_extract_canonical_host_packed(ctx: *mut kmb_ctx, host_bits: *const u32, host_inv: *const u16, n_reads: u64, fixed_len: u64, k: u32, flags: u32, out_canon: *mut u64, out_hash: *mut u64, digest: *mut kmb_digest) -> i32;
    pub fn kmb_host_pack(bases: *const u8, n_bases: u64, bits_out: *mut u32, inv_out: *mut u16) -> i32;
    pub fn kmb_host_pack_isa() -> *const c_char;
    pub fn kmb_host_read_probe(buf: *const u8, n_bytes: u64, n_threads: u32, seconds_out: *mut f64) -> i32;
    pub fn kmb_ctx_set_host_threads(ctx: *mut kmb_ctx, n_threads: u32) -> i32;
    pub fn kmb_ctx_host_stats(ctx: *const kmb_ctx, stats4: *mut u64) -> i32;
    pub fn kmb_minimizers(ctx: *mut kmb_ctx, k: u32, w: u32, hash_k: u32, flags: u32, mmer_out: *mut u64, pos_out: *mut u32) -> i32;
    pub fn kmb_minimizer_words(ctx: *mut kmb_ctx, k: u32, w: u32, hash_k: u32, words: *const u64, n: u64, mmer_out: *mut u64, offset_out: *mut u32) -> i32;
    pub fn kmb_batch_repack(ctx: *mut kmb_ctx, strict: i32) -> i32;
    pub fn kmb_batch_attach_packed(ctx: *mut kmb_ctx, dev_words: *const u64, n_words: u64, dev_offsets: *const u64, dev_word_offsets: *const u64, n_reads: u64, fixed_len: u64) -> i32;
    pub fn kmb_packed_get_kmers(ctx: *mut kmb_ctx, k: u32, reads: *const u64, pos: *const u64, n: u64, out: *mut u64) -> i32;
    pub fn kmb_batch_new_packed(ctx: *mut kmb_ctx, capacity_bases: u64) -> i32;
    pub fn kmb_packed_push_chars(ctx: *mut kmb_ctx, bases: *const u8, n: u64) -> i32;
    pub fn kmb_batch_slice(ctx: *mut kmb_ctx, read: u64, start: u64, len: u64) -> i32;
    pub fn kmb_batch_unslice(ctx: *mut kmb_ctx) -> i32;
    pub fn kmb_allreduce_u64(ctxs: *const *mut kmb_ctx, n_ctx: i32, dev_bufs: *const *mut u64, count: u64) -> i32;
    pub fn kmb_parse_fastx(text: *const c_char, n_bytes: u64, bases_out: *mut u8, bases_cap: u64, offsets_out: *mut u64, reads_cap: u64, n_reads: *mut u64, n_bases: *mut u64) -> i32;
    pub fn kmb_batch_ingest_fastx(ctx: *mut kmb_ctx, text: *const c_char, n_bytes: u64, n_reads_out: *mut u64, n_bases_out: *mut u64) -> i32;
    pub fn kmb_pack(ctx: *mut kmb_ctx, enc: i32, word_bits: u32, words_out: *mut c_void, word_offsets_out: *mut u64) -> i32;
    pub fn kmb_pack_num_words(ctx: *mut kmb_ctx, word_bits: u32, n_words: *mut u64) -> i32;
    pub fn kmb_unpack(ctx: *mut kmb_ctx, enc: i32, word_bits: u32, words_in: *const c_void, n_items: u64, words_per_item: u32, bases_per_item: u32, bases_out: *mut u8) -> i32;
    pub fn kmb_words_to_strings(ctx: *mut kmb_ctx, k: u32, words: *const u64, n: u64, bases_out: *mut u8) -> i32;
    pub fn kmb_revcomp_words(ctx: *mut kmb_ctx, enc: i32, k: u32, word_bits: u32, words_per_item: u32, words_in: *const c_void, words_out: *mut c_void, n_items: u64) -> i32;
    pub fn kmb_reverse_complement_words(ctx: *mut kmb_ctx, k: u32, input: *const u64, out: *mut u64, n: u64) -> i32;
    pub fn kmb_canonical_words(ctx: *mut kmb_ctx, k: u32, input: *const u64, canon_out: *mut u64, is_canonical_out: *mut u8, n: u64) -> i32;
    pub fn kmb_lexhash_words(ctx: *mut kmb_ctx, k: u32, input: *const u64, out: *mut u64, n: u64) -> i32;
    pub fn kmb_match_words(ctx: *mut kmb_ctx, k: u32, words: *const u64, others: *const u64, match_out: *mut u8, n: u64) -> i32;
    pub fn kmb_sub_kmer_words(ctx: *mut kmb_ctx, k: u32, pos: u32, width: u32, input: *const u64, out: *mut u64, n: u64) -> i32;
    pub fn kmb_append_base_words(ctx: *mut kmb_ctx, k: u32, input: *const u64, bases: *const u8, bases_are_ascii: i32, out: *mut u64, dropped_out: *mut u8, n: u64) -> i32;
    pub fn kmb_prepend_base_words(ctx: *mut kmb_ctx, k: u32, input: *const u64, bases: *const u8, bases_are_ascii: i32, out: *mut u64, dropped_out: *mut u8, n: u64) -> i32;
    pub fn kmb_canonical_append_base_words(ctx: *mut kmb_ctx, k: u32, fw_in: *const u64, rc_in: *const u64, bases: *const u8, bases_are_ascii: i32, fw_out: *mut u64, rc_out: *mut u64, dropped_out: *mut u8, n: u64) -> i32;
    pub fn kmb_canonical_prepend_base_words(ctx: *mut kmb_ctx, k: u32, fw_in: *const u64, rc_in: *const u64, bases: *const u8, bases_are_ascii: i32, fw_out: *mut u64, rc_out: *mut u64, dropped_out: *mut u8, n: u64) -> i32;
    pub fn kmb_is_fw_canonical_words(ctx: *mut kmb_ctx, fw: *const u64, rc: *const u64, out: *mut u8, n: u64) -> i32;
    pub fn kmb_kmer_get(ctx: *mut kmb_ctx, word_bits: u32, words_per_item: u32, arrays: *const c_void, n_items: u64, index: u32, codes_out: *mut u8) -> i32;
    pub fn kmb_kmer_get_prefix(ctx: *mut kmb_ctx, word_bits: u32, words_per_item: u32, arrays: *const c_void, n_items: u64, len: u32, words_out: *mut c_void) -> i32;
    pub fn kmb_bitmer_to_bytes(ctx: *mut kmb_ctx, len: u32, mers: *const u64, n: u64, bases_out: *mut u8) -> i32;
    // END GENERATED
}
