// kmb_tu_minimizer.cu -- instantiates the minimizer engines on both geometries, and Kmer::minimizer_word.
#include "kmb_launch.h"

namespace kmb {

cudaError_t launch_minimizers(bool validate, const FixedGeom* fg, const CsrGeom* cg, const Launch& l, cudaStream_t st,
                              const EncDesc& enc, const MinParams& ep) {
    const int cls = ep.mc.w <= 13 ? 0 : (ep.mc.w <= 15 ? 1 : 2);  // width class of the candidate compare (kmb_minimizer.cuh)
#define KMB_CASE(V, C) \
    if (validate == V && cls == C) return launch_eng<MinimizerEng<V, C>>(fg, cg, l, st, enc, ep);
    KMB_CASE(true, 0) KMB_CASE(true, 1) KMB_CASE(true, 2) KMB_CASE(false, 0) KMB_CASE(false, 1) KMB_CASE(false, 2)
#undef KMB_CASE
    return cudaErrorInvalidValue;
}

// Kmer::minimizer_word (naive_impl/kmer.rs:170-191) on every word: leftmost width-mer of minimum LexHash.
__global__ void __launch_bounds__(256) minimizer_words_kernel(const uint64_t* in, uint64_t n, uint32_t k, uint32_t width,
                                                              uint32_t hash_k, uint64_t* mmer_out, uint32_t* offset_out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t word = in[i];
    const uint64_t wmask = width >= 32 ? ~0ull : ((1ull << (2 * width)) - 1ull);
    // LexHasher::write_u64 (hash.rs:60-71) of an lmer = its first hash_k bases read as a number, base 0 most significant.
    // Pair-reversing the whole k-mer ONCE puts base 0 in the top field, so the rank of the lmer at `pos` is a plain
    // field extract: v = (rev >> 2 (k - pos - width)) & wmask; the hash is v >> 2 (width - hash_k) for hash_k < width and
    // v << 2 (hash_k - width) otherwise -- the latter orders and ties exactly like v itself.
    const uint64_t rev = pair_reverse64(word) >> (2 * (32 - k));
    const uint32_t hs = hash_k < width ? 2 * (width - hash_k) : 0;
    uint64_t min_mmer = word & wmask, min_hash = ~0ull;
    uint32_t off = 0;
    for (uint32_t pos = 0; pos + width <= k; ++pos) {                    // sub_kmer_word, kmer.rs:155-161
        const uint64_t h = ((rev >> (2 * (k - pos - width))) & wmask) >> hs;
        if (h < min_hash) { min_hash = h; off = pos; }                   // strict '<': the leftmost minimum (kmer.rs:183)
    }
    if (off) min_mmer = (word >> (2 * off)) & wmask;
    if (mmer_out) mmer_out[i] = min_mmer;
    if (offset_out) offset_out[i] = off;
}

cudaError_t launch_minimizer_words(const uint64_t* in, uint64_t n, uint32_t k, uint32_t w, uint32_t hash_k, uint64_t* mmer_out,
                                   uint32_t* offset_out, cudaStream_t st) {
    minimizer_words_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in, n, k, w, hash_k, mmer_out, offset_out);
    return cudaGetLastError();
}

}  // namespace kmb
