// kmers_b200.hpp -- C++17 host-side mirror of the reference's interface for the hot path,
// in batched form, over the C ABI (include/kmers_b200.h).  Header-only; link libkmers_b200.so.
//
// The reference is a Rust crate (compiled code) and this image has no Rust toolchain, so the
// host side above the C ABI is C++ and keeps the reference's names and argument meaning:
//   encoding::Naive / encoding::Xor10 / Encoding::{encode, decode, rev_comp<K>}   encoding/mod.rs:14-23
//   Kmer<P,K,B>, word_for_k<P,K>()                                                   kmer.rs:12-69
//   naive_impl::{reverse_complement, to_canonical, LexHasherState, MatchType}        naive_impl/*.rs
//   naive_impl::CanonicalKmerIterator (as a whole-batch extraction)                  canonical_kmer_iterator.rs
// Error behaviour: where the reference panics, these throw kmers_b200::Panic; other failures
// throw kmers_b200::Error.  Nothing here computes on the CPU.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "kmers_b200.h"

namespace kmers_b200 {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error("kmers_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
struct Panic : Error {  // the reference would panic!/assert! on these arguments
    using Error::Error;
};

namespace detail {
inline void check(const kmb_ctx* ctx, int32_t rc) {
    if (rc == KMB_OK) return;
    const char* m = kmb_last_error(ctx);
    if (rc == KMB_ERR_PANIC) throw Panic(rc, m ? m : "");
    throw Error(rc, m ? m : "");
}
}  // namespace detail

// ------------------------------------------------------------------ encoding
namespace encoding {
// encoding/naive.rs:49-74 : the discriminant is the code table (A,C,T,G in bits 7-6,5-4,3-2,1-0)
enum class Naive : uint8_t {
    ACTG = 0b00011011, ACGT = 0b00011110, ATCG = 0b00100111, ATGC = 0b00110110, AGCT = 0b00101101, AGTC = 0b00111001,
    CATG = 0b01001011, CAGT = 0b01001110, CTAG = 0b10000111, CTGA = 0b11000110, CGAT = 0b10001101, CGTA = 0b11001001,
    TACG = 0b01100011, TAGC = 0b01110010, TCAG = 0b10010011, TCGA = 0b11010010, TGAC = 0b10110001, TGCA = 0b11100001,
    GACT = 0b01101100, GATC = 0b01111000, GCAT = 0b10011100, GCTA = 0b11011000, GTAC = 0b10110100, GTCA = 0b11100100,
};
struct Xor10 {};  // encoding/xor10.rs:12
inline int32_t id(Naive e) { return static_cast<int32_t>(e); }
inline int32_t id(Xor10) { return KMB_ENC_XOR10; }
}  // namespace encoding

// kmer.rs:67-69
template <class P, size_t K>
constexpr size_t word_for_k() {
    return (sizeof(P) * 8 / 2 + K - 1) / (sizeof(P) * 8 / 2);
}

// naive_impl/canonical_kmer.rs:7-12
enum class MatchType : uint8_t { NoMatch = KMB_NO_MATCH, IdentityMatch = KMB_IDENTITY_MATCH, TwinMatch = KMB_TWIN_MATCH };

struct Digest {
    uint64_t n_valid = 0, checksum_canon = 0, checksum_hash = 0;
};

// Dense-slot result of a batched CanonicalKmerIterator run, on the host.
struct CanonicalKmers {
    uint32_t k = 0;
    std::vector<uint64_t> canon, hash, fw, rc;  // slot = win_offset[read] + pos ; KMB_SENTINEL = skipped window
    Digest digest;
};

// Compacted iterator output: entry i is the i-th emitted k-mer; emit_offsets[r] = first entry of read r.
struct CompactKmers {
    std::vector<int32_t> pos;
    std::vector<uint64_t> canon, hash, emit_offsets;
};

class Context;

// The device-resident read batch of a Context (stand-in for the &[u8] handed to
// CanonicalKmerIterator::from_u8_slice / Encoding::encode).
class ReadBatch {
  public:
    uint64_t num_slots(uint32_t k) const;
    // CanonicalKmerIterator + get_canonical_word + hash_one(LexHasherState(k)) over every read
    CanonicalKmers canonical_kmers(uint32_t k, bool want_fw_rc = false, bool validate = true) const;
    // same, results left in caller-provided DEVICE (or host) buffers; digest optional
    void canonical_kmers_into(uint32_t k, uint64_t* canon, uint64_t* hash, Digest* digest = nullptr, bool validate = true) const;
    // EXTENSION: 1 <= k <= 64, two words per slot
    template <class Enc>
    std::vector<std::array<uint64_t, 2>> canonical_kmers_wide(uint32_t k, Enc enc, Digest* digest = nullptr) const;
    // fused histogram by the top hist_bits of the LexHash + digest
    std::vector<uint64_t> histogram(uint32_t k, uint32_t hist_bits, Digest* digest = nullptr) const;
    // exactly what `while !it.exhausted() { it.get(); it.inc(); }` yields, read after read
    // (CanonicalKmerPos{km, pos}, canonical_kmer_iterator.rs:13-16)
    CompactKmers canonical_kmer_positions(uint32_t k) const;
    // SeqVecMinimizerIter in batch (seq_vector/minimizers.rs:38-142): (lmer word, pos) per k-mer window
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> minimizers(uint32_t k, uint32_t w, uint32_t hash_k) const;
    // switch the batch to the 2-bit packed store (SeqVector layout, seq_vector.rs:230-242)
    void to_packed(bool strict = true) const { detail::check(ctx_, kmb_batch_repack(ctx_, strict ? 1 : 0)); }
    // SeqVector::get_kmer_u64 (seq_vector.rs:96-99) at (read, pos) pairs of a packed batch
    std::vector<uint64_t> get_kmers_u64(uint32_t k, const std::vector<uint64_t>& reads, const std::vector<uint64_t>& pos) const {
        std::vector<uint64_t> out(pos.size());
        detail::check(ctx_, kmb_packed_get_kmers(ctx_, k, reads.empty() ? nullptr : reads.data(), pos.data(), pos.size(), out.data()));
        return out;
    }
    // Encoding::encode of every read -> byte image
    template <class Enc>
    std::vector<uint8_t> pack(Enc enc, uint32_t word_bits) const;

  private:
    friend class Context;
    explicit ReadBatch(kmb_ctx* c) : ctx_(c) {}
    kmb_ctx* ctx_;
};

class Context {
  public:
    explicit Context(int device = -1, void* cuda_stream = nullptr) { detail::check(nullptr, kmb_ctx_create(device, cuda_stream, &ctx_)); }
    ~Context() { kmb_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    kmb_ctx* raw() const { return ctx_; }
    void sync() { detail::check(ctx_, kmb_ctx_sync(ctx_)); }

    ReadBatch upload(const uint8_t* bases, uint64_t n_bytes, uint64_t n_reads, uint64_t fixed_len) {
        detail::check(ctx_, kmb_batch_upload(ctx_, bases, n_bytes, nullptr, n_reads, fixed_len));
        return ReadBatch(ctx_);
    }
    ReadBatch upload(const std::vector<std::string>& reads) {  // ragged
        std::vector<uint64_t> offs(reads.size() + 1, 0);
        std::string flat;
        for (size_t i = 0; i < reads.size(); ++i) { flat += reads[i]; offs[i + 1] = flat.size(); }
        detail::check(ctx_, kmb_batch_upload(ctx_, reinterpret_cast<const uint8_t*>(flat.data()), flat.size(), offs.data(),
                                             reads.size(), 0));
        return ReadBatch(ctx_);
    }
    ReadBatch ingest_fastx(const std::string& text) {  // FASTA / FASTQ text -> ragged batch
        detail::check(ctx_, kmb_batch_ingest_fastx(ctx_, text.data(), text.size(), nullptr, nullptr));
        return ReadBatch(ctx_);
    }
    ReadBatch generate(uint64_t seed, uint64_t n_reads, uint64_t fixed_len, uint32_t n_thresh20 = 0, uint64_t first_index = 0) {
        detail::check(ctx_, kmb_batch_generate(ctx_, seed, first_index, n_reads, fixed_len, n_thresh20));
        return ReadBatch(ctx_);
    }

    // ---- batched naive_impl::Kmer word operations (host vectors in, host vectors out)
    std::vector<uint64_t> reverse_complement(const std::vector<uint64_t>& words, uint32_t k) {  // kmer.rs:138-147
        std::vector<uint64_t> out(words.size());
        detail::check(ctx_, kmb_reverse_complement_words(ctx_, k, words.data(), out.data(), words.size()));
        return out;
    }
    std::vector<uint64_t> to_canonical(const std::vector<uint64_t>& words, uint32_t k, std::vector<uint8_t>* is_canonical = nullptr) {
        std::vector<uint64_t> out(words.size());
        if (is_canonical) is_canonical->resize(words.size());
        detail::check(ctx_, kmb_canonical_words(ctx_, k, words.data(), out.data(), is_canonical ? is_canonical->data() : nullptr,
                                                words.size()));
        return out;
    }
    // hash_one(&LexHasherState::new(k), kmer), naive_impl/hash.rs:10-20
    std::vector<uint64_t> lex_hash(const std::vector<uint64_t>& words, uint32_t k) {
        std::vector<uint64_t> out(words.size());
        detail::check(ctx_, kmb_lexhash_words(ctx_, k, words.data(), out.data(), words.size()));
        return out;
    }
    // CanonicalKmer::from_u64(w, k).get_word_equivalency(other), canonical_kmer.rs:152-161
    std::vector<MatchType> get_word_equivalency(const std::vector<uint64_t>& words, const std::vector<uint64_t>& others, uint32_t k) {
        std::vector<MatchType> out(words.size());
        detail::check(ctx_, kmb_match_words(ctx_, k, words.data(), others.data(), reinterpret_cast<uint8_t*>(out.data()), words.size()));
        return out;
    }

    // Kmer::minimizer_word with LexHasherState(hash_k) (naive_impl/kmer.rs:170-191) -> (mmer, offset) per word
    std::pair<std::vector<uint64_t>, std::vector<uint32_t>> minimizer_word(const std::vector<uint64_t>& words, uint32_t k, uint32_t width,
                                                                          uint32_t hash_k) {
        std::vector<uint64_t> mm(words.size());
        std::vector<uint32_t> off(words.size());
        detail::check(ctx_, kmb_minimizer_words(ctx_, k, width, hash_k, words.data(), words.size(), mm.data(), off.data()));
        return {std::move(mm), std::move(off)};
    }

    // Kmer::sub_kmer_word (naive_impl/kmer.rs:150-161)
    std::vector<uint64_t> sub_kmer(const std::vector<uint64_t>& words, uint32_t k, uint32_t pos, uint32_t width) {
        std::vector<uint64_t> out(words.size());
        detail::check(ctx_, kmb_sub_kmer_words(ctx_, k, pos, width, words.data(), out.data(), words.size()));
        return out;
    }
    // Kmer::append_base_u8 (naive_impl/kmer.rs:83-102) / prepend_base_u8 (:76-95): shifted words + the bases shifted off
    std::pair<std::vector<uint64_t>, std::vector<uint8_t>> append_base_u8(const std::vector<uint64_t>& words, const std::string& bases, uint32_t k) {
        std::vector<uint64_t> out(words.size());
        std::vector<uint8_t> dropped(words.size());
        detail::check(ctx_, kmb_append_base_words(ctx_, k, words.data(), reinterpret_cast<const uint8_t*>(bases.data()), 1, out.data(),
                                                  dropped.data(), words.size()));
        return {std::move(out), std::move(dropped)};
    }
    std::pair<std::vector<uint64_t>, std::vector<uint8_t>> prepend_base_u8(const std::vector<uint64_t>& words, const std::string& bases, uint32_t k) {
        std::vector<uint64_t> out(words.size());
        std::vector<uint8_t> dropped(words.size());
        detail::check(ctx_, kmb_prepend_base_words(ctx_, k, words.data(), reinterpret_cast<const uint8_t*>(bases.data()), 1, out.data(),
                                                   dropped.data(), words.size()));
        return {std::move(out), std::move(dropped)};
    }
    // bitmer_to_bytes (kmer.rs:71-91): upper case, A0 C1 G2 T3
    std::vector<std::string> bitmer_to_bytes(const std::vector<uint64_t>& mers, uint32_t len) {
        std::string flat(mers.size() * len, '\0');
        detail::check(ctx_, kmb_bitmer_to_bytes(ctx_, len, mers.data(), mers.size(), reinterpret_cast<uint8_t*>(&flat[0])));
        std::vector<std::string> out(mers.size());
        for (size_t i = 0; i < mers.size(); ++i) out[i] = flat.substr(i * len, len);
        return out;
    }
    // Kmer::<P,K,B>::get / get_prefix (kmer.rs:46-53) on every array
    template <class P, size_t B>
    std::vector<uint8_t> get(const std::vector<std::array<P, B>>& arrays, uint32_t index) {
        std::vector<uint8_t> out(arrays.size());
        detail::check(ctx_, kmb_kmer_get(ctx_, sizeof(P) * 8, B, arrays.data(), arrays.size(), index, out.data()));
        return out;
    }
    template <class P, size_t B>
    std::vector<P> get_prefix(const std::vector<std::array<P, B>>& arrays, uint32_t len) {
        std::vector<P> out(arrays.size());
        detail::check(ctx_, kmb_kmer_get_prefix(ctx_, sizeof(P) * 8, B, arrays.data(), arrays.size(), len, out.data()));
        return out;
    }

    // One call on reads in host memory (packed by worker threads, H2D / kernel / D2H overlapped): CanonicalKmerIterator +
    // get_canonical_word + LexHasher for every read, results in host vectors
    CanonicalKmers canonical_kmers_host(const uint8_t* bases, uint64_t n_reads, uint64_t fixed_len, uint32_t k) {
        CanonicalKmers r;
        const uint64_t n = fixed_len >= k ? n_reads * (fixed_len - k + 1) : 0;
        r.canon.resize(n);
        r.hash.resize(n);
        kmb_digest d{};
        detail::check(ctx_, kmb_extract_canonical_host(ctx_, bases, n_reads, fixed_len, k, 0, r.canon.data(), r.hash.data(), &d));
        r.digest = {d.n_valid, d.checksum_canon, d.checksum_hash};
        return r;
    }

    // ---- batched Encoding<P,B> on arrays [P;B]
    // Encoding::encode / Kmer::<P,K,B>::new for n k-mers of K ASCII bytes each (encoding/naive.rs:116-124)
    template <class P, size_t K, class Enc>
    std::vector<std::array<P, word_for_k<P, K>()>> encode(Enc enc, const uint8_t* seqs, uint64_t n) {
        constexpr size_t B = word_for_k<P, K>();
        std::vector<std::array<P, B>> out(n);
        detail::check(ctx_, kmb_batch_upload(ctx_, seqs, n * K, nullptr, n, K));
        detail::check(ctx_, kmb_pack(ctx_, encoding::id(enc), sizeof(P) * 8, out.data(), nullptr));
        return out;
    }
    // Encoding::decode (encoding/naive.rs:126-136): every position of the array, padding included
    template <class P, size_t B, class Enc>
    std::vector<std::string> decode(Enc enc, const std::vector<std::array<P, B>>& arrays) {
        const uint32_t per = B * sizeof(P) * 4;
        std::string flat(arrays.size() * per, '\0');
        detail::check(ctx_, kmb_unpack(ctx_, encoding::id(enc), sizeof(P) * 8, arrays.data(), arrays.size(), B, per,
                                       reinterpret_cast<uint8_t*>(&flat[0])));
        std::vector<std::string> out(arrays.size());
        for (size_t i = 0; i < arrays.size(); ++i) out[i] = flat.substr(i * per, per);
        return out;
    }
    // String::from(Kmer) (naive_impl/kmer.rs:196-207): lower-case letters, base 0 first
    std::vector<std::string> to_strings(const std::vector<uint64_t>& words, uint32_t k) {
        std::string flat(words.size() * k, '\0');
        detail::check(ctx_, kmb_words_to_strings(ctx_, k, words.data(), words.size(), reinterpret_cast<uint8_t*>(&flat[0])));
        std::vector<std::string> out(words.size());
        for (size_t i = 0; i < words.size(); ++i) out[i] = flat.substr(i * k, k);
        return out;
    }
    // Encoding::rev_comp::<K> (encoding/naive.rs:138-154)
    template <size_t K, class P, size_t B, class Enc>
    std::vector<std::array<P, B>> rev_comp(Enc enc, const std::vector<std::array<P, B>>& arrays) {
        std::vector<std::array<P, B>> out(arrays.size());
        detail::check(ctx_, kmb_revcomp_words(ctx_, encoding::id(enc), K, sizeof(P) * 8, B, arrays.data(), out.data(), arrays.size()));
        return out;
    }

  private:
    kmb_ctx* ctx_ = nullptr;
};

inline uint64_t ReadBatch::num_slots(uint32_t k) const {
    uint64_t n = 0;
    detail::check(ctx_, kmb_batch_num_slots(ctx_, k, &n));
    return n;
}

inline CanonicalKmers ReadBatch::canonical_kmers(uint32_t k, bool want_fw_rc, bool validate) const {
    CanonicalKmers r;
    r.k = k;
    const uint64_t n = num_slots(k);
    r.canon.resize(n);
    r.hash.resize(n);
    if (want_fw_rc) { r.fw.resize(n); r.rc.resize(n); }
    kmb_digest d{};
    detail::check(ctx_, kmb_extract_canonical(ctx_, k, validate ? 0u : KMB_F_NO_VALIDATE, r.canon.data(), r.hash.data(),
                                              want_fw_rc ? r.fw.data() : nullptr, want_fw_rc ? r.rc.data() : nullptr, &d));
    r.digest = {d.n_valid, d.checksum_canon, d.checksum_hash};
    return r;
}

inline void ReadBatch::canonical_kmers_into(uint32_t k, uint64_t* canon, uint64_t* hash, Digest* digest, bool validate) const {
    kmb_digest d{};
    detail::check(ctx_, kmb_extract_canonical(ctx_, k, validate ? 0u : KMB_F_NO_VALIDATE, canon, hash, nullptr, nullptr,
                                              digest ? &d : nullptr));
    if (digest) *digest = {d.n_valid, d.checksum_canon, d.checksum_hash};
}

template <class Enc>
std::vector<std::array<uint64_t, 2>> ReadBatch::canonical_kmers_wide(uint32_t k, Enc enc, Digest* digest) const {
    std::vector<std::array<uint64_t, 2>> out(num_slots(k));
    kmb_digest d{};
    detail::check(ctx_, kmb_extract_canonical_wide(ctx_, k, encoding::id(enc), 0, reinterpret_cast<uint64_t*>(out.data()), nullptr,
                                                   digest ? &d : nullptr));
    if (digest) *digest = {d.n_valid, d.checksum_canon, d.checksum_hash};
    return out;
}

inline std::vector<uint64_t> ReadBatch::histogram(uint32_t k, uint32_t hist_bits, Digest* digest) const {
    std::vector<uint64_t> out(size_t(1) << hist_bits);
    kmb_digest d{};
    detail::check(ctx_, kmb_histogram(ctx_, k, 0, hist_bits, out.data(), 0, digest ? &d : nullptr));
    if (digest) *digest = {d.n_valid, d.checksum_canon, d.checksum_hash};
    return out;
}

inline CompactKmers ReadBatch::canonical_kmer_positions(uint32_t k) const {
    CompactKmers r;
    uint64_t n = 0, nb = 0, nr = 0, fl = 0;
    detail::check(ctx_, kmb_extract_compact(ctx_, k, 0, nullptr, nullptr, nullptr, nullptr, 0, &n));
    kmb_batch_info(ctx_, &nb, &nr, &fl);
    r.pos.resize(n); r.canon.resize(n); r.hash.resize(n); r.emit_offsets.resize(nr + 1);
    detail::check(ctx_, kmb_extract_compact(ctx_, k, 0, r.canon.data(), r.hash.data(), r.pos.data(), r.emit_offsets.data(), n, &n));
    return r;
}

inline std::pair<std::vector<uint64_t>, std::vector<uint32_t>> ReadBatch::minimizers(uint32_t k, uint32_t w, uint32_t hash_k) const {
    const uint64_t n = num_slots(k);
    std::vector<uint64_t> mm(n);
    std::vector<uint32_t> pos(n);
    detail::check(ctx_, kmb_minimizers(ctx_, k, w, hash_k, 0, mm.data(), pos.data()));
    return {std::move(mm), std::move(pos)};
}

template <class Enc>
std::vector<uint8_t> ReadBatch::pack(Enc enc, uint32_t word_bits) const {
    uint64_t n = 0;
    detail::check(ctx_, kmb_pack_num_words(ctx_, word_bits, &n));
    std::vector<uint8_t> out(n * word_bits / 8);
    std::vector<uint64_t> woff;
    uint64_t nb = 0, nr = 0, fl = 0;
    kmb_batch_info(ctx_, &nb, &nr, &fl);
    if (fl == 0) woff.resize(nr + 1);
    detail::check(ctx_, kmb_pack(ctx_, encoding::id(enc), word_bits, out.data(), fl == 0 ? woff.data() : nullptr));
    return out;
}

// final reduction across the GPUs one process drives: in-place wrapping-u64 sum of dev_bufs[i] (device memory of
// ctxs[i]'s GPU) through NCCL (kmb_allreduce_u64)
inline void allreduce_u64(const std::vector<Context*>& ctxs, const std::vector<uint64_t*>& dev_bufs, uint64_t count) {
    std::vector<kmb_ctx*> raw;
    for (Context* c : ctxs) raw.push_back(c->raw());
    detail::check(raw.empty() ? nullptr : raw[0], kmb_allreduce_u64(raw.data(), (int32_t)raw.size(), dev_bufs.data(), count));
}

}  // namespace kmers_b200
