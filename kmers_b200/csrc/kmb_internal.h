// kmb_internal.h -- what the translation units behind the C ABI share: the context, error plumbing and the one
// extraction launcher.  Internal (hidden visibility): nothing here is exported.
#pragma once
#include "../../include/kmers_b200.h"

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#define KMB_HIDDEN __attribute__((visibility("hidden")))

struct kmb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    uint64_t launches = 0;

    // resident batch
    const uint8_t* d_bases = nullptr;
    const uint64_t* d_offsets = nullptr;
    uint8_t* own_bases = nullptr;
    size_t own_bases_cap = 0;
    uint64_t* own_offsets = nullptr;
    size_t own_offsets_cap = 0;
    uint64_t n_bytes = 0, n_reads = 0, fixed_len = 0;
    bool have_batch = false;
    // 2-bit packed batch (SeqVector layout): d_bases points at u64 words; reads start on word boundaries
    bool packed = false;
    uint64_t stride_len = 0;                  // fixed-length: bases between read starts (fixed_len padded to 32)
    const uint64_t* d_base_starts = nullptr;  // ragged: flat (padded) base index of every read's first base, n_reads + 1
    uint64_t n_bases_flat = 0;                // size of the flat base index space
    uint64_t* own_packed = nullptr;
    size_t own_packed_cap = 0;
    uint64_t* own_base_starts = nullptr;
    size_t own_base_starts_cap = 0;
    // a view of one read of the packed store (SeqVectorSlice): the parent's description while the view is in place
    bool sliced = false;
    uint64_t* own_slice = nullptr;  // 4 words: offsets [0, len], base starts [flat start, flat end]
    const uint64_t* parent_offsets = nullptr;
    const uint64_t* parent_base_starts = nullptr;
    uint64_t parent_n_reads = 0, parent_fixed_len = 0;

    // CSR window-offset cache (per k)
    uint64_t* d_win_offsets = nullptr;
    size_t win_cap = 0;
    uint32_t win_k = 0;
    uint64_t win_total = 0;
    bool win_valid = false;

    uint64_t* d_first_read = nullptr;  // per-CTA first read of the CSR kernels
    size_t first_read_cap = 0;
    unsigned long long* d_cta_counts = nullptr;  // compaction: per-tile look-back descriptors, ticket, total
    size_t cta_counts_cap = 0;

    // scratch
    unsigned long long* d_digest = nullptr;  // 3 words (+1 spare)
    unsigned long long* h_digest = nullptr;  // pinned, 4 words
    void* d_scratch[4] = {nullptr, nullptr, nullptr, nullptr};
    size_t scratch_cap[4] = {0, 0, 0, 0};
    void* d_cub = nullptr;
    size_t cub_cap = 0;
    uint8_t* h_stage[2] = {nullptr, nullptr};
    size_t stage_cap = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};

    // pipelined host path (kmb_hostpipe.cu): streams, rings, worker pool; created on first use
    void* hostpipe = nullptr;
    uint32_t host_threads = 0;  // 0 = default (kmb_ctx_set_host_threads)
};

KMB_HIDDEN int32_t kmb_i_fail(kmb_ctx* ctx, int32_t code, const char* fmt, ...);

#define CK(ctx, call)                                                                                    \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            cudaGetLastError();                                                                          \
            return kmb_i_fail(ctx, e_ == cudaErrorMemoryAllocation ? KMB_ERR_NOMEM : KMB_ERR_CUDA,       \
                              "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                                \
    } while (0)

#define NEED_CTX(ctx) \
    do { if (!(ctx)) return kmb_i_fail(nullptr, KMB_ERR_INVALID_ARG, "ctx is NULL"); } while (0)

KMB_HIDDEN int32_t kmb_i_bind(kmb_ctx* ctx);
#define BIND(ctx) do { int32_t b_ = kmb_i_bind(ctx); if (b_ != KMB_OK) return b_; } while (0)

KMB_HIDDEN bool kmb_i_is_device_ptr(const void* p);
KMB_HIDDEN bool kmb_i_is_pinned_ptr(const void* p);
// (re)allocate device memory to hold `need` bytes (syncs the ctx stream before freeing the old block)
KMB_HIDDEN int32_t kmb_i_grow(kmb_ctx* ctx, void** ptr, size_t* cap, size_t need);
KMB_HIDDEN int32_t kmb_i_digest_begin(kmb_ctx* ctx);
KMB_HIDDEN int32_t kmb_i_digest_end(kmb_ctx* ctx, kmb_digest* digest);

// How the bytes behind `d_bases` are laid out for one extraction launch
enum kmb_i_layout : uint32_t {
    KMB_I_ASCII = 0,       // 1 byte per base
    KMB_I_SEQVECTOR = 1,   // 2-bit packed, every read padded to a u64 word (SeqVector twin; no invalid bases)
    KMB_I_FLAT_PACKED = 2  // 2-bit packed flat stream + 16-bit invalid masks (kmb_hostpack.h), reads back to back
};

// One extraction launch over (d_bases, fixed_len) on stream st -- or over the ctx's resident CSR batch when csr.
// canon / hash / fw / rc: dense-slot device arrays (may be NULL); hist != NULL: fused histogram instead.
KMB_HIDDEN int32_t kmb_i_run_extract(kmb_ctx* ctx, const uint8_t* d_bases, bool csr, uint64_t n_bytes, uint64_t n_reads,
                                     uint64_t fixed_len, uint32_t k, uint32_t flags, uint64_t* canon, uint64_t* hash, uint64_t* fw,
                                     uint64_t* rc, bool want_digest, unsigned long long* hist, uint32_t hist_bits, cudaStream_t st,
                                     uint64_t stride, uint32_t layout, const uint16_t* d_inv, unsigned long long* d_digest = nullptr);
// d_digest: where the three digest words accumulate when want_digest (default: the context's scratch words)

KMB_HIDDEN void kmb_i_hostpipe_destroy(kmb_ctx* ctx);
