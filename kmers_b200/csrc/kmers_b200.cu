// kmers_b200.cu -- C ABI (include/kmers_b200.h) over the sm_100a kernels.
// The extraction engines are instantiated in their own translation units (kmb_tu_*.cu, declared in kmb_launch.h)
// so that the library builds in parallel.
// No torch types, no CPU fallback: every compute entry point needs a device.
#include "../../include/kmers_b200.h"

#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include <cub/device/device_scan.cuh>

#include "kmb_encoding.cuh"
#include "kmb_internal.h"
#include "kmb_launch.h"

using namespace kmb;

// ======================================================================= ctx (struct kmb_ctx: kmb_internal.h)
static thread_local std::string g_err;

int32_t kmb_i_fail(kmb_ctx* ctx, int32_t code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf; else g_err = buf;
    return code;
}
#define fail kmb_i_fail

int32_t kmb_i_bind(kmb_ctx* ctx) { CK(ctx, cudaSetDevice(ctx->device)); return KMB_OK; }

bool kmb_i_is_device_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
bool kmb_i_is_pinned_ptr(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}
#define is_device_ptr kmb_i_is_device_ptr
#define is_pinned_ptr kmb_i_is_pinned_ptr

int32_t kmb_i_grow(kmb_ctx* ctx, void** ptr, size_t* cap, size_t need) {
    if (*cap >= need && *ptr) return KMB_OK;
    if (*ptr) { CK(ctx, cudaStreamSynchronize(ctx->stream)); CK(ctx, cudaFree(*ptr)); *ptr = nullptr; *cap = 0; }
    size_t bytes = need + 256;
    CK(ctx, cudaMalloc(ptr, bytes));
    *cap = need;
    return KMB_OK;
}
#define grow kmb_i_grow

extern "C" int32_t kmb_version(void) { return KMB_VERSION; }

extern "C" int32_t kmb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

extern "C" int32_t kmb_ctx_create(int32_t device, void* cuda_stream, kmb_ctx** out) {
    if (!out) return fail(nullptr, KMB_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int n = kmb_device_count();
    if (n <= 0) return fail(nullptr, KMB_ERR_NO_DEVICE, "no CUDA device: kmers_b200 has no CPU fallback");
    if (device < 0) {
        if (cudaGetDevice(&device) != cudaSuccess) { cudaGetLastError(); device = 0; }
    }
    if (device >= n) return fail(nullptr, KMB_ERR_INVALID_ARG, "device %d out of range (%d visible)", device, n);
    kmb_ctx* ctx = new (std::nothrow) kmb_ctx();
    if (!ctx) return fail(nullptr, KMB_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) {
        // (void*)1 / (void*)2 are cudaStreamLegacy / cudaStreamPerThread: valid borrowed handles
        if (cuda_stream) { ctx->stream = (cudaStream_t)cuda_stream; ctx->own_stream = false; }
        else { e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking); ctx->own_stream = true; }
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&ctx->d_digest, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_digest, 4 * sizeof(unsigned long long), cudaHostAllocDefault);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        int32_t rc = fail(nullptr, KMB_ERR_CUDA, "ctx creation failed: %s", cudaGetErrorString(e));
        cudaGetLastError();
        kmb_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return KMB_OK;
}

extern "C" int32_t kmb_ctx_destroy(kmb_ctx* ctx) {
    if (!ctx) return KMB_OK;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    kmb_i_hostpipe_destroy(ctx);
    cudaFree(ctx->own_bases);
    cudaFree(ctx->own_offsets);
    cudaFree(ctx->own_packed);
    cudaFree(ctx->own_base_starts);
    cudaFree(ctx->own_slice);
    cudaFree(ctx->d_win_offsets);
    cudaFree(ctx->d_first_read);
    cudaFree(ctx->d_cta_counts);
    cudaFree(ctx->d_digest);
    cudaFreeHost(ctx->h_digest);
    for (auto& s : ctx->d_scratch) cudaFree(s);
    cudaFree(ctx->d_cub);
    for (int i = 0; i < 2; ++i) {
        cudaFreeHost(ctx->h_stage[i]);
        if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    cudaGetLastError();
    delete ctx;
    return KMB_OK;
}

extern "C" const char* kmb_last_error(const kmb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_err.c_str(); }

extern "C" int32_t kmb_ctx_sync(kmb_ctx* ctx) {
    NEED_CTX(ctx);
    BIND(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" void* kmb_ctx_stream(kmb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
extern "C" uint64_t kmb_ctx_launch_count(const kmb_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int32_t kmb_device_alloc(kmb_ctx* ctx, size_t bytes, void** out) {
    NEED_CTX(ctx);
    if (!out) return fail(ctx, KMB_ERR_INVALID_ARG, "out is NULL");
    BIND(ctx);
    CK(ctx, cudaMalloc(out, bytes ? bytes : 1));
    return KMB_OK;
}
extern "C" int32_t kmb_device_free(kmb_ctx* ctx, void* ptr) {
    NEED_CTX(ctx);
    BIND(ctx);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    CK(ctx, cudaFree(ptr));
    return KMB_OK;
}
extern "C" int32_t kmb_host_alloc_pinned(kmb_ctx* ctx, size_t bytes, void** out) {
    NEED_CTX(ctx);
    if (!out) return fail(ctx, KMB_ERR_INVALID_ARG, "out is NULL");
    BIND(ctx);
    CK(ctx, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return KMB_OK;
}
extern "C" int32_t kmb_host_free_pinned(kmb_ctx* ctx, void* ptr) {
    NEED_CTX(ctx);
    BIND(ctx);
    CK(ctx, cudaFreeHost(ptr));
    return KMB_OK;
}
extern "C" int32_t kmb_memcpy(kmb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (bytes == 0) return KMB_OK;
    if (!dst || !src) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
    return KMB_OK;
}

// ======================================================================= encodings
static bool make_enc(int32_t enc, EncDesc* d, uint32_t* dec_letters) {
    if (enc == KMB_ENC_XOR10) enc = KMB_ENC_ACTG;  // xor10.rs:17-22 == Naive::ACTG
    if (enc < 0 || enc > 255) return false;
    // code of internal x (0=A 1=C 2=T 3=G, encoding/naive.rs:14-16, 78-86)
    uint32_t code[4];
    uint32_t seen = 0;
    for (int x = 0; x < 4; ++x) { code[x] = ((uint32_t)enc >> (6 - 2 * x)) & 3u; seen |= 1u << code[x]; }
    if (seen != 0xF) return false;  // not one of the 24 permutations
    for (int b = 0; b < 2; ++b) {
        uint32_t f00 = (code[0] >> b) & 1, f01 = (code[1] >> b) & 1, f10 = (code[2] >> b) & 1, f11 = (code[3] >> b) & 1;
        d->k0[b] = f00 ? 0x55555555u : 0u;
        d->k1[b] = (f01 ^ f00) ? 0x55555555u : 0u;
        d->k2[b] = (f10 ^ f00) ? 0x55555555u : 0u;
        d->k3[b] = (f11 ^ f10 ^ f01 ^ f00) ? 0x55555555u : 0u;
    }
    d->cmask = (code[0] ^ code[2]) * 0x55555555u;  // complement = XOR constant (naive.rs:98-110)
    d->is_acgt = (enc == KMB_ENC_ACGT) ? 1u : 0u;
    if (dec_letters) {
        static const char letters[4] = {'A', 'C', 'T', 'G'};  // INTERNAL2NUC, naive.rs:19
        uint32_t dec = 0;
        for (int x = 0; x < 4; ++x) dec |= (uint32_t)letters[x] << (8 * code[x]);
        *dec_letters = dec;
    }
    return true;
}

// ======================================================================= batch
static void drop_batch(kmb_ctx* ctx) {
    ctx->have_batch = false;
    ctx->sliced = false;
    ctx->win_valid = false;
    ctx->d_bases = nullptr;
    ctx->d_offsets = nullptr;
    ctx->packed = false;
    ctx->d_base_starts = nullptr;
}

static int32_t check_shape(kmb_ctx* ctx, uint64_t n_bytes, bool has_offsets, uint64_t n_reads, uint64_t fixed_len) {
    if (has_offsets == (fixed_len > 0) && !(n_reads == 0 && !has_offsets))
        return fail(ctx, KMB_ERR_INVALID_ARG, "give either offsets or fixed_len > 0");
    if (!has_offsets && n_reads * fixed_len != n_bytes)
        return fail(ctx, KMB_ERR_INVALID_ARG, "n_bytes (%llu) != n_reads * fixed_len (%llu)",
                    (unsigned long long)n_bytes, (unsigned long long)(n_reads * fixed_len));
    return KMB_OK;
}

static int32_t stage_h2d(kmb_ctx* ctx, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return KMB_OK;
    if (is_pinned_ptr(src) || is_device_ptr(src)) {
        CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, ctx->stream));
        return KMB_OK;
    }
    const size_t chunk = (size_t)32 << 20;
    if (ctx->stage_cap < chunk) {
        for (int i = 0; i < 2; ++i) {
            if (ctx->h_stage[i]) { CK(ctx, cudaFreeHost(ctx->h_stage[i])); ctx->h_stage[i] = nullptr; }
            CK(ctx, cudaHostAlloc((void**)&ctx->h_stage[i], chunk, cudaHostAllocDefault));
        }
        ctx->stage_cap = chunk;
    }
    size_t done = 0;
    int buf = 0;
    while (done < bytes) {
        size_t n = bytes - done < chunk ? bytes - done : chunk;
        CK(ctx, cudaEventSynchronize(ctx->stage_ev[buf]));  // previous use of this staging buffer drained
        memcpy(ctx->h_stage[buf], (const uint8_t*)src + done, n);
        CK(ctx, cudaMemcpyAsync((uint8_t*)dst + done, ctx->h_stage[buf], n, cudaMemcpyHostToDevice, ctx->stream));
        CK(ctx, cudaEventRecord(ctx->stage_ev[buf], ctx->stream));
        done += n;
        buf ^= 1;
    }
    return KMB_OK;
}

extern "C" int32_t kmb_batch_upload(kmb_ctx* ctx, const uint8_t* bases, uint64_t n_bytes, const uint64_t* offsets,
                                    uint64_t n_reads, uint64_t fixed_len) {
    NEED_CTX(ctx);
    BIND(ctx);
    int32_t rc = check_shape(ctx, n_bytes, offsets != nullptr, n_reads, fixed_len);
    if (rc) return rc;
    if (n_bytes && !bases) return fail(ctx, KMB_ERR_INVALID_ARG, "bases is NULL");
    if (offsets) {  // checked before any state changes: a bad table leaves the resident batch as it was
        if (offsets[0] != 0 || offsets[n_reads] != n_bytes)
            return fail(ctx, KMB_ERR_INVALID_ARG, "offsets must start at 0 and end at n_bytes");
        for (uint64_t r = 0; r < n_reads; ++r)
            if (offsets[r + 1] < offsets[r])
                return fail(ctx, KMB_ERR_INVALID_ARG, "offsets must be ascending (offsets[%llu] > offsets[%llu])", (unsigned long long)r,
                            (unsigned long long)(r + 1));
    }
    drop_batch(ctx);
    if ((rc = grow(ctx, (void**)&ctx->own_bases, &ctx->own_bases_cap, n_bytes + 64))) return rc;
    if ((rc = stage_h2d(ctx, ctx->own_bases, bases, n_bytes))) return rc;
    ctx->d_bases = ctx->own_bases;
    if (offsets) {
        if ((rc = grow(ctx, (void**)&ctx->own_offsets, &ctx->own_offsets_cap, (n_reads + 1) * 8))) return rc;
        if ((rc = stage_h2d(ctx, ctx->own_offsets, offsets, (n_reads + 1) * 8))) return rc;
        ctx->d_offsets = ctx->own_offsets;
    }
    ctx->n_bytes = n_bytes; ctx->n_reads = n_reads; ctx->fixed_len = fixed_len;
    ctx->stride_len = ctx->fixed_len; ctx->n_bases_flat = ctx->n_bytes;
    ctx->have_batch = true;
    return KMB_OK;
}

// borrowed CSR tables live in device memory: one cheap pass flags a table that is not ascending from 0 to `last`
// (and, for packed batches, word offsets too small for their reads)
__global__ void __launch_bounds__(256) check_offsets_kernel(const uint64_t* offsets, const uint64_t* word_offsets, uint64_t n_reads,
                                                            uint64_t last, unsigned long long* bad) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r > n_reads) return;
    bool ok = true;
    if (r == 0) ok = offsets[0] == 0 && (!word_offsets || word_offsets[0] == 0);
    if (r == n_reads) ok = ok && offsets[r] == last;
    if (r < n_reads) {
        ok = ok && offsets[r + 1] >= offsets[r];
        if (word_offsets && ok)
            ok = word_offsets[r + 1] >= word_offsets[r] && (word_offsets[r + 1] - word_offsets[r]) >= (offsets[r + 1] - offsets[r] + 31) / 32;
    }
    if (!ok) atomicAdd(bad, 1ull);
}

static int32_t check_device_offsets(kmb_ctx* ctx, const uint64_t* d_offsets, const uint64_t* d_word_offsets, uint64_t n_reads, uint64_t last,
                                    bool check_last) {
    CK(ctx, cudaMemsetAsync(ctx->d_digest + 3, 0, 8, ctx->stream));
    if (!check_last) {  // the last entry is whatever it is: read it back instead of comparing
        CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, d_offsets + n_reads, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        last = ctx->h_digest[3];
    }
    check_offsets_kernel<<<(unsigned)((n_reads + 1 + 255) / 256), 256, 0, ctx->stream>>>(d_offsets, d_word_offsets, n_reads, last, ctx->d_digest + 3);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, ctx->d_digest + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_digest[3] != 0)
        return fail(ctx, KMB_ERR_INVALID_ARG, "offsets must ascend from 0 to the batch size (%llu entries violate that)", ctx->h_digest[3]);
    return KMB_OK;
}

extern "C" int32_t kmb_batch_attach(kmb_ctx* ctx, const uint8_t* dev_bases, uint64_t n_bytes,
                                    const uint64_t* dev_offsets, uint64_t n_reads, uint64_t fixed_len) {
    NEED_CTX(ctx);
    BIND(ctx);
    int32_t rc = check_shape(ctx, n_bytes, dev_offsets != nullptr, n_reads, fixed_len);
    if (rc) return rc;
    if (n_bytes && !is_device_ptr(dev_bases)) return fail(ctx, KMB_ERR_INVALID_ARG, "dev_bases is not device memory");
    if (dev_offsets && !is_device_ptr(dev_offsets)) return fail(ctx, KMB_ERR_INVALID_ARG, "dev_offsets is not device memory");
    if (dev_offsets && (rc = check_device_offsets(ctx, dev_offsets, nullptr, n_reads, n_bytes, true))) return rc;
    drop_batch(ctx);
    ctx->d_bases = dev_bases; ctx->d_offsets = dev_offsets;
    ctx->n_bytes = n_bytes; ctx->n_reads = n_reads; ctx->fixed_len = fixed_len;
    ctx->stride_len = ctx->fixed_len; ctx->n_bases_flat = ctx->n_bytes;
    ctx->have_batch = true;
    return KMB_OK;
}

extern "C" int32_t kmb_batch_generate(kmb_ctx* ctx, uint64_t seed, uint64_t first_index, uint64_t n_reads,
                                      uint64_t fixed_len, uint32_t n_thresh20) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (fixed_len == 0) return fail(ctx, KMB_ERR_INVALID_ARG, "fixed_len must be > 0");
    const uint64_t n = n_reads * fixed_len;
    drop_batch(ctx);
    int32_t rc = grow(ctx, (void**)&ctx->own_bases, &ctx->own_bases_cap, n + 64);
    if (rc) return rc;
    if (n) {
        const uint64_t chunks = (n + 15) / 16;
        generate_kernel<<<(unsigned)((chunks + 255) / 256), 256, 0, ctx->stream>>>(ctx->own_bases, n, seed + first_index,
                                                                                  n_thresh20);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
    }
    ctx->d_bases = ctx->own_bases; ctx->d_offsets = nullptr;
    ctx->n_bytes = n; ctx->n_reads = n_reads; ctx->fixed_len = fixed_len;
    ctx->stride_len = ctx->fixed_len; ctx->n_bases_flat = ctx->n_bytes;
    ctx->have_batch = true;
    return KMB_OK;
}

#define NEED_BATCH(ctx) \
    do { if (!(ctx)->have_batch) return fail(ctx, KMB_ERR_STATE, "no read batch resident (upload / attach / generate first)"); } while (0)

extern "C" int32_t kmb_batch_download(kmb_ctx* ctx, uint8_t* dst, uint64_t n_bytes) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (n_bytes > ctx->n_bytes) return fail(ctx, KMB_ERR_INVALID_ARG, "n_bytes exceeds the batch");  // packed: the packed bytes
    if (n_bytes == 0) return KMB_OK;
    CK(ctx, cudaMemcpyAsync(dst, ctx->d_bases, n_bytes, cudaMemcpyDefault, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_batch_info(const kmb_ctx* ctx, uint64_t* n_bytes, uint64_t* n_reads, uint64_t* fixed_len) {
    if (!ctx) return fail(nullptr, KMB_ERR_INVALID_ARG, "ctx is NULL");
    if (!ctx->have_batch) return KMB_ERR_STATE;
    if (n_bytes) *n_bytes = ctx->n_bytes;
    if (n_reads) *n_reads = ctx->n_reads;
    if (fixed_len) *fixed_len = ctx->fixed_len;
    return KMB_OK;
}

// exclusive prefix (n_reads + 1 entries) of per-read counts; div == 0 -> windows of length k
static int32_t scan_counts(kmb_ctx* ctx, uint32_t k, uint32_t div, uint64_t* d_out) {
    const uint64_t n = ctx->n_reads + 1;
    read_counts_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->d_offsets, ctx->n_reads, k - 1, div, d_out);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    size_t need = 0;
    CK(ctx, cub::DeviceScan::ExclusiveSum(nullptr, need, d_out, d_out, (long long)n, ctx->stream));
    int32_t rc = grow(ctx, &ctx->d_cub, &ctx->cub_cap, need);
    if (rc) return rc;
    CK(ctx, cub::DeviceScan::ExclusiveSum(ctx->d_cub, need, d_out, d_out, (long long)n, ctx->stream));
    ctx->launches++;
    return KMB_OK;
}

// makes ctx->d_win_offsets / win_total valid for k (CSR batches)
static int32_t ensure_win_offsets(kmb_ctx* ctx, uint32_t k) {
    if (ctx->win_valid && ctx->win_k == k) return KMB_OK;
    int32_t rc = grow(ctx, (void**)&ctx->d_win_offsets, &ctx->win_cap, (ctx->n_reads + 1) * 8);
    if (rc) return rc;
    if ((rc = scan_counts(ctx, k, 0, ctx->d_win_offsets))) return rc;
    CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, ctx->d_win_offsets + ctx->n_reads, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->win_total = ctx->h_digest[3];
    ctx->win_k = k;
    ctx->win_valid = true;
    return KMB_OK;
}

static int32_t num_slots(kmb_ctx* ctx, uint32_t k, uint64_t* out) {
    if (ctx->d_offsets) {
        if (ctx->n_reads == 0) { *out = 0; return KMB_OK; }
        int32_t rc = ensure_win_offsets(ctx, k);
        if (rc) return rc;
        *out = ctx->win_total;
    } else {
        *out = ctx->fixed_len >= k ? ctx->n_reads * (ctx->fixed_len - k + 1) : 0;
    }
    return KMB_OK;
}

extern "C" int32_t kmb_batch_num_slots(kmb_ctx* ctx, uint32_t k, uint64_t* n_slots) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (!n_slots) return fail(ctx, KMB_ERR_INVALID_ARG, "n_slots is NULL");
    if (k < 1) return fail(ctx, KMB_ERR_PANIC, "k = 0: a window needs at least one base (2*k-2 underflows, SURVEY Q14)");
    return num_slots(ctx, k, n_slots);
}

extern "C" int32_t kmb_batch_window_offsets(kmb_ctx* ctx, uint32_t k, uint64_t* out) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (k < 1 || !out) return fail(ctx, KMB_ERR_INVALID_ARG, "k must be >= 1 and out non-NULL");
    if (!ctx->d_offsets) return fail(ctx, KMB_ERR_STATE, "fixed-length batch: window offset of read r is r * (L - k + 1)");
    if (ctx->n_reads == 0) {
        CK(ctx, cudaMemsetAsync(ctx->d_digest + 3, 0, 8, ctx->stream));
        CK(ctx, cudaMemcpyAsync(out, ctx->d_digest + 3, 8, cudaMemcpyDefault, ctx->stream));
    } else {
        int32_t rc = ensure_win_offsets(ctx, k);
        if (rc) return rc;
        CK(ctx, cudaMemcpyAsync(out, ctx->d_win_offsets, (ctx->n_reads + 1) * 8, cudaMemcpyDefault, ctx->stream));
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// ======================================================================= output staging
// A caller pointer is either device memory (used in place) or host memory
// (device scratch now, copied back after the kernels).
struct OutBuf {
    void* user = nullptr;
    void* dev = nullptr;
    size_t bytes = 0;
    bool host = false;
};

static int32_t out_prepare(kmb_ctx* ctx, int slot, void* user, size_t bytes, OutBuf* ob) {
    ob->user = user; ob->bytes = bytes; ob->dev = nullptr; ob->host = false;
    if (!user || bytes == 0) return KMB_OK;
    if (is_device_ptr(user)) { ob->dev = user; return KMB_OK; }
    int32_t rc = grow(ctx, &ctx->d_scratch[slot], &ctx->scratch_cap[slot], bytes);
    if (rc) return rc;
    ob->dev = ctx->d_scratch[slot];
    ob->host = true;
    return KMB_OK;
}
static int32_t out_finish(kmb_ctx* ctx, const OutBuf& ob) {
    if (ob.host && ob.bytes) CK(ctx, cudaMemcpyAsync(ob.user, ob.dev, ob.bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return KMB_OK;
}

int32_t kmb_i_digest_begin(kmb_ctx* ctx) {
    CK(ctx, cudaMemsetAsync(ctx->d_digest, 0, 3 * sizeof(unsigned long long), ctx->stream));
    return KMB_OK;
}
int32_t kmb_i_digest_end(kmb_ctx* ctx, kmb_digest* digest) {
    CK(ctx, cudaMemcpyAsync(ctx->h_digest, ctx->d_digest, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    digest->n_valid = ctx->h_digest[0];
    digest->checksum_canon = ctx->h_digest[1];
    digest->checksum_hash = ctx->h_digest[2];
    return KMB_OK;
}

#define digest_begin kmb_i_digest_begin
#define digest_end kmb_i_digest_end

// ======================================================================= geometry (shared by both engines)
static uint32_t mask32(uint32_t nbits) { return nbits >= 32 ? 0xFFFFFFFFu : ((1u << nbits) - 1u); }

// fixed-length reads: slot-space geometry of fixed_kernel
static bool make_fixed_geom(const uint8_t* d_bases, uint64_t n_bytes, uint64_t n_reads, uint64_t L, uint32_t k,
                            uint32_t span_entries, FixedGeom* g, Launch* l, uint64_t stride = 0, uint32_t layout = KMB_I_ASCII,
                            uint32_t ipc_max = kItemsPerCta, size_t smem_max = 36 * 1024, const uint16_t* d_inv = nullptr,
                            bool fine_steps = false) {
    if (stride == 0) stride = L;
    g->bases = d_bases; g->n_bytes = n_bytes; g->L = stride; g->L32 = (uint32_t)stride; g->packed = layout; g->inv = d_inv;
    g->W = L - k + 1;
    if (g->W > 0xFFFF0000ull) return false;  // one read of > 4.29 Gbases: split it (window positions are 32-bit inside a CTA)
    g->W32 = (uint32_t)g->W;
    g->unified = (g->W >= (uint64_t)kRun && g->W % kRun != 0) ? 1u : 0u;
    g->total_slots = n_reads * g->W;
    g->w_magic = g->W32 > 1 ? (uint32_t)((1ull << 32) / g->W32 + 1) : 0;
    // slot / W by multiplication is exact while slot * W < 2^64; otherwise the kernel divides
    const bool magic_ok = g->W > 1 && (double)g->total_slots * (double)g->W < 9.0e18;
    g->w_magic64 = magic_ok ? (~0ull / g->W + 1) : 0;
    // items per CTA: as many as keep the staged stretch of reads under smem_max (default 36 KiB: with the kernels' ~10 KiB of
    // static shared memory that stays below the 48 KiB a launch gets without opting in)
    uint32_t ipc = ipc_max;
    for (;;) {
        const uint64_t slots = (uint64_t)ipc * kRun;
        const uint64_t crossings = slots / g->W + 2;
        const uint64_t span = slots + crossings * (k - 1 + (stride - L)) + k + 32;
        l->smem = (size_t)((span + 15) / 16 + span_entries + 2) * sizeof(uint2);
        if (l->smem <= smem_max || ipc <= 64) break;
        // whole rounds of kExtractThreads items while that is possible (fine_steps), else halves
        ipc = fine_steps && ipc > 2 * kExtractThreads ? ipc - kExtractThreads : ipc / 2;
    }
    g->items_per_cta = ipc;
    const uint64_t items = (g->total_slots + kRun - 1) / kRun;
    const uint64_t ctas = (items + ipc - 1) / ipc;
    if (ctas > 0x7FFFFFFFull) return false;
    l->grid = (unsigned)ctas;
    return true;
}

// ragged reads: window offsets (cached per k), per-CTA first-read index, shared-memory layout of csr_kernel
// Ragged reads: 1536 items per tile in the same 36.8 K-base staging buffer (reads of >= 24 bases on average fit one pass;
// shorter ones take the multi-pass path as before) -- a CTA's serial head (descriptor, offsets, staging, owner table: two
// dependent global loads and three barriers) is paid per tile.  Reads of 100..150 bp, K=31 canon+hash, ms per 4 M reads:
// 1024 items 1.144, 1280 1.129, 1536 1.106, 1792 1.116, 2016 1.125, 2560 1.143 (a larger staging buffer changes nothing);
// K=63 88.9 instead of 82.5 % of the copy peak, minimizers 72.8 instead of 69.5 %.  On fixed-length reads 1024 stays best.
constexpr uint32_t kCsrItemsPerCta = 1536;
static int32_t make_csr_geom(kmb_ctx* ctx, uint32_t k, uint32_t span_entries, CsrGeom* g, Launch* l, uint32_t ipc = kCsrItemsPerCta) {
    int32_t rc = ensure_win_offsets(ctx, k);
    if (rc) return rc;
    g->bases = ctx->d_bases; g->n_bytes = ctx->n_bytes; g->n_bases = ctx->n_bases_flat; g->packed = ctx->packed ? 1u : 0u;
    g->offsets = ctx->packed ? ctx->d_base_starts : ctx->d_offsets; g->win_offsets = ctx->d_win_offsets;
    g->n_reads = ctx->n_reads; g->total_slots = ctx->win_total; g->items_per_cta = ipc;
    g->tile_entries = 2304 * (ipc / kItemsPerCta) + span_entries;  // ~36.8 K bases per pass and 1024 (up to 2047) items
    const uint64_t slots_per_cta = (uint64_t)ipc * kRun;
    const uint64_t ctas = (g->total_slots + slots_per_cta - 1) / slots_per_cta;
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    l->grid = (unsigned)ctas;
    l->smem = (size_t)g->tile_entries * sizeof(uint2) + 2 * (size_t)(kCsrCache + 2) * sizeof(uint64_t);
    if ((rc = grow(ctx, (void**)&ctx->d_first_read, &ctx->first_read_cap, (ctas + 1) * sizeof(CsrTileDesc)))) return rc;
    CsrTileDesc* d_desc = reinterpret_cast<CsrTileDesc*>(ctx->d_first_read);
    g->tile_desc = d_desc;
    if (ctas) {
        const uint32_t tile_bases = (g->tile_entries - span_entries - 1) * 16;  // as csr_body computes it
        csr_index_kernel<<<(unsigned)((ctas + 255) / 256), 256, 0, ctx->stream>>>(g->offsets, ctx->d_win_offsets, ctx->n_reads, g->total_slots,
                                                                               slots_per_cta, ctas, k, tile_bases, d_desc);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
    }
    return KMB_OK;
}

// ======================================================================= extract (K <= 32)
static WinConst make_winconst(uint32_t k, const EncDesc& enc) {
    WinConst wc{};
    wc.K = k;
    wc.shiftD = 2 * (48 - (kRun + k - 1));
    wc.mask_lo = mask32(2 * k);
    wc.mask_hi = k > 16 ? mask32(2 * k - 32) : 0u;
    wc.cmask = enc.cmask;
    wc.cm_lo = enc.cmask & wc.mask_lo;
    wc.cm_hi = enc.cmask & wc.mask_hi;
    wc.kmask = mask32(k);
    return wc;
}

// One extraction launch over (d_bases, fixed_len) -- or over the ctx's resident CSR batch when fixed_len == 0.
int32_t kmb_i_run_extract(kmb_ctx* ctx, const uint8_t* d_bases, bool csr, uint64_t n_bytes, uint64_t n_reads,
                          uint64_t fixed_len, uint32_t k, uint32_t flags, uint64_t* canon, uint64_t* hash, uint64_t* fw,
                          uint64_t* rc, bool want_digest, unsigned long long* hist, uint32_t hist_bits, cudaStream_t st,
                          uint64_t stride, uint32_t layout, const uint16_t* d_inv, unsigned long long* d_digest) {
    EncDesc enc;
    make_enc(KMB_ENC_ACGT, &enc, nullptr);
    // a SeqVector-packed store holds no invalid base; the flat packed staging format carries its invalid masks
    const bool validate = !(flags & KMB_F_NO_VALIDATE) && layout != KMB_I_SEQVECTOR;
    const bool fwrc = fw || rc;
    const bool khi = k > 16;
    NarrowParams ep{};
    ep.wc = make_winconst(k, enc);
    ep.out.canon = canon; ep.out.hash = hash; ep.out.fw = fw; ep.out.rc = rc;
    ep.out.digest = d_digest ? d_digest : ctx->d_digest; ep.out.hist = hist; ep.out.hist_shift = 2 * k - hist_bits;
    ep.out.vec_ok = ((((uintptr_t)canon | (uintptr_t)hash | (uintptr_t)fw | (uintptr_t)rc) & 31u) == 0) ? 1u : 0u;
    ep.out.one = 1u;
    ep.out.bin_mask = hist ? (uint32_t)((1ull << hist_bits) - 1ull) : 0u;
    ep.out.hist_hi_shift = (hist && ep.out.hist_shift >= 32u) ? ep.out.hist_shift - 32u : 0xFFFFFFFFu;
    FixedGeom fg{};
    CsrGeom cg{};
    Launch l;
    // shared-memory histogram: 1024-thread CTAs, so four times the items per tile (as many per thread as elsewhere)
    const bool big = hist && hist_bits <= 16 && !getenv("KMB_HIST_GLOBAL");
    if (!csr) {
        if (fixed_len < k || n_reads == 0) return KMB_OK;
        if (!make_fixed_geom(d_bases, n_bytes, n_reads, fixed_len, k, 4, &fg, &l, stride, layout, big ? kMaxItemsPerCta : kItemsPerCta,
                             big ? 64 * 1024 : 36 * 1024, d_inv))
            return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch (or a read above 4.29 Gbases)");
    } else {
        if (n_bytes == 0 || n_reads == 0) return KMB_OK;
        int32_t r = make_csr_geom(ctx, k, 4, &cg, &l, big ? kMaxItemsPerCta : kCsrItemsPerCta);
        if (r) return r;
        if (cg.total_slots == 0) return KMB_OK;
    }
    cudaError_t e;
    if (hist && hist_bits <= 16 && !getenv("KMB_HIST_GLOBAL")) {
        // shared-memory bins: a persistent grid of as many CTAs as fit (tile + 2 bytes per bin of shared memory each)
        const uint32_t n_bins = 1u << hist_bits;
        const size_t smem = l.smem + (size_t)n_bins * 2 + 16;
        int sms = 148, max_smem = 227 * 1024;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
        cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device);
        if (smem + 2048 <= (size_t)max_smem) {
            int per_sm = (int)((size_t)(228 * 1024) / (smem + 2048));
            if (per_sm < 1) per_sm = 1;
            if (per_sm > 2048 / kHistThreads) per_sm = 2048 / kHistThreads;
            const unsigned grid = (unsigned)std::min<uint64_t>(l.grid, (uint64_t)sms * per_sm);
            e = launch_hist_smem(validate, want_digest, khi, csr ? nullptr : &fg, csr ? &cg : nullptr, grid, smem, l.grid,
                                 (uint32_t)(l.smem / sizeof(uint2)), n_bins, st, enc, ep);
            CK(ctx, e);
            ctx->launches++;
            return KMB_OK;
        }
    }
    e = hist ? launch_narrow_hist_global(validate, want_digest, khi, csr ? nullptr : &fg, csr ? &cg : nullptr, l, st, enc, ep)
             : launch_narrow_materialise(validate, want_digest, fwrc, khi, hash != nullptr, csr ? nullptr : &fg, csr ? &cg : nullptr, l, st, enc, ep);
    CK(ctx, e);
    ctx->launches++;
    return KMB_OK;
}

extern "C" int32_t kmb_extract_canonical(kmb_ctx* ctx, uint32_t k, uint32_t flags, uint64_t* canon_out,
                                         uint64_t* hash_out, uint64_t* fw_out, uint64_t* rc_out, kmb_digest* digest) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    // naive_impl/kmer.rs:236-238 panics above 32 bases; k == 0 underflows 2*k-2
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    uint64_t n_slots = 0;
    int32_t rc = num_slots(ctx, k, &n_slots);
    if (rc) return rc;
    OutBuf ob[4];
    void* user[4] = {canon_out, hash_out, fw_out, rc_out};
    for (int i = 0; i < 4; ++i)
        if ((rc = out_prepare(ctx, i, user[i], n_slots * 8, &ob[i]))) return rc;
    if (digest && (rc = digest_begin(ctx))) return rc;
    if (n_slots) {
        rc = kmb_i_run_extract(ctx, ctx->d_bases, ctx->d_offsets != nullptr, ctx->n_bytes, ctx->n_reads, ctx->fixed_len, k, flags,
                         (uint64_t*)ob[0].dev, (uint64_t*)ob[1].dev, (uint64_t*)ob[2].dev, (uint64_t*)ob[3].dev,
                         digest != nullptr, nullptr, 0, ctx->stream, ctx->stride_len, ctx->packed ? KMB_I_SEQVECTOR : KMB_I_ASCII, nullptr);
        if (rc) return rc;
    }
    bool any_host = false;
    for (int i = 0; i < 4; ++i) { if ((rc = out_finish(ctx, ob[i]))) return rc; any_host |= ob[i].host; }
    if (digest) return digest_end(ctx, digest);
    if (any_host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_histogram(kmb_ctx* ctx, uint32_t k, uint32_t flags, uint32_t hist_bits, uint64_t* hist_out,
                                 int32_t accumulate, kmb_digest* digest) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    if (hist_bits < 1 || hist_bits > 2 * k || hist_bits > 28 || !hist_out)
        return fail(ctx, KMB_ERR_INVALID_ARG, "hist_bits must be in [1, min(2k, 28)] and hist_out non-NULL");
    // KMB_F_DIGEST_IN_HIST: the three digest words are accumulated behind the bins (hist_out[n_bins .. n_bins + 2]) instead of
    // being read back, so that [bins | digest] is ONE buffer the ranks all-reduce in place, and the call stays asynchronous
    const bool tail = (flags & KMB_F_DIGEST_IN_HIST) != 0;
    if (tail && digest) return fail(ctx, KMB_ERR_INVALID_ARG, "KMB_F_DIGEST_IN_HIST: the digest goes to hist_out, pass digest = NULL");
    const size_t n_bins = (size_t)1 << hist_bits;
    const size_t bytes = (n_bins + (tail ? 3 : 0)) * 8;
    OutBuf ob;
    int32_t rc = out_prepare(ctx, 0, hist_out, bytes, &ob);
    if (rc) return rc;
    if (ob.host && accumulate) CK(ctx, cudaMemcpyAsync(ob.dev, hist_out, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (!accumulate) CK(ctx, cudaMemsetAsync(ob.dev, 0, bytes, ctx->stream));
    if (digest && (rc = digest_begin(ctx))) return rc;
    rc = kmb_i_run_extract(ctx, ctx->d_bases, ctx->d_offsets != nullptr, ctx->n_bytes, ctx->n_reads, ctx->fixed_len, k, flags, nullptr,
                           nullptr, nullptr, nullptr, digest != nullptr || tail, (unsigned long long*)ob.dev, hist_bits, ctx->stream,
                           ctx->stride_len, ctx->packed ? KMB_I_SEQVECTOR : KMB_I_ASCII, nullptr,
                           tail ? (unsigned long long*)ob.dev + n_bins : nullptr);
    if (rc) return rc;
    if ((rc = out_finish(ctx, ob))) return rc;
    if (digest) return digest_end(ctx, digest);
    if (ob.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// ======================================================================= compacted (iterator-identical) output
extern "C" int32_t kmb_extract_compact(kmb_ctx* ctx, uint32_t k, uint32_t flags, uint64_t* canon_out, uint64_t* hash_out,
                                       int32_t* pos_out, uint64_t* emit_offsets_out, uint64_t capacity, uint64_t* n_emitted) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    if (!n_emitted) return fail(ctx, KMB_ERR_INVALID_ARG, "n_emitted is NULL");
    // the reference's pos is i32 (canonical_kmer_iterator.rs:15); longer reads cannot be represented
    if (!ctx->d_offsets && ctx->fixed_len > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "reads above 2^31-1 bases: pos is i32");
    *n_emitted = 0;
    EncDesc enc;
    make_enc(KMB_ENC_ACGT, &enc, nullptr);
    const bool validate = !(flags & KMB_F_NO_VALIDATE) && !ctx->packed, khi = k > 16, csr = ctx->d_offsets != nullptr;
    uint64_t n_slots = 0;
    int32_t rc = num_slots(ctx, k, &n_slots);
    if (rc) return rc;
    OutBuf oe;
    if ((rc = out_prepare(ctx, 3, emit_offsets_out, (ctx->n_reads + 1) * 8, &oe))) return rc;
    if (n_slots == 0) {
        if (oe.dev) CK(ctx, cudaMemsetAsync(oe.dev, 0, (ctx->n_reads + 1) * 8, ctx->stream));
        if ((rc = out_finish(ctx, oe))) return rc;
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        return KMB_OK;
    }
    FixedGeom fg{};
    CsrGeom cg{};
    Launch l;
    if (!csr) {
        if (!make_fixed_geom(ctx->d_bases, ctx->n_bytes, ctx->n_reads, ctx->fixed_len, k, 4, &fg, &l, ctx->stride_len, ctx->packed ? KMB_I_SEQVECTOR : KMB_I_ASCII))
            return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch (or a read above 4.29 Gbases)");
    } else if ((rc = make_csr_geom(ctx, k, 4, &cg, &l, kCompactCsrItems))) {  // (the engine's shared-memory counters hold that many items)
        return rc;
    }
    const FixedGeom* pf = csr ? nullptr : &fg;
    const CsrGeom* pc = csr ? &cg : nullptr;
    // device words behind the launch: [grid look-back descriptors | ticket | total], zeroed before every launch
    if ((rc = grow(ctx, (void**)&ctx->d_cta_counts, &ctx->cta_counts_cap, ((size_t)l.grid + 2) * 8))) return rc;
    unsigned long long* d_ticket = ctx->d_cta_counts + l.grid;
    unsigned long long* d_total = d_ticket + 1;
    CompactParams ep{};
    ep.wc = make_winconst(k, enc);
    ep.out.desc = ctx->d_cta_counts; ep.out.ticket = d_ticket; ep.out.total = d_total;
    const bool counting_call = !canon_out && !hash_out && !pos_out && !emit_offsets_out;
    // fixed-length reads are emitted by the persistent pipelined kernel: two tile buffers per CTA, so the tiles may be smaller
    const bool pipe = !csr && !counting_call;
    if (pipe && !make_fixed_geom(ctx->d_bases, ctx->n_bytes, ctx->n_reads, ctx->fixed_len, k, 4, &fg, &l, ctx->stride_len,
                                 ctx->packed ? KMB_I_SEQVECTOR : KMB_I_ASCII, kCompactPipeItems, compact_pipe_tile_budget(), nullptr, true))
        return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch (or a read above 4.29 Gbases)");
    if (pipe && (rc = grow(ctx, (void**)&ctx->d_cta_counts, &ctx->cta_counts_cap, ((size_t)l.grid + 2) * 8))) return rc;
    if (pipe) {
        d_ticket = ctx->d_cta_counts + l.grid; d_total = d_ticket + 1;
        ep.out.desc = ctx->d_cta_counts; ep.out.ticket = d_ticket; ep.out.total = d_total;
    }
    if (counting_call) {
        // sizing: count only (reads the bases, writes nothing but the total)
        CK(ctx, cudaMemsetAsync(d_total, 0, 8, ctx->stream));
        CK(ctx, launch_compact(true, validate, khi, pf, pc, l, ctx->stream, enc, ep));
        ctx->launches++;
        CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        *n_emitted = ctx->h_digest[3];
        return KMB_OK;
    }
    // ONE launch: stage, count from the staged masks, look back over the earlier tiles, emit.  The arrays may be sized by a
    // counting call or simply hold the worst case (one entry per slot); nothing is written at or beyond `capacity`.
    const uint64_t cap = capacity < n_slots ? capacity : n_slots;
    OutBuf oc, oh, op;
    if ((rc = out_prepare(ctx, 0, canon_out, cap * 8, &oc))) return rc;
    if ((rc = out_prepare(ctx, 1, hash_out, cap * 8, &oh))) return rc;
    if ((rc = out_prepare(ctx, 2, pos_out, cap * 4, &op))) return rc;
    ep.out.canon = (uint64_t*)oc.dev; ep.out.hash = (uint64_t*)oh.dev; ep.out.pos = (int32_t*)op.dev;
    ep.out.emit_offsets = (uint64_t*)oe.dev;
    ep.out.capacity = cap;
    ep.out.vec16 = ((((uintptr_t)oc.dev | (uintptr_t)oh.dev) & 15u) == 0 && ((uintptr_t)op.dev & 7u) == 0) ? 1u : 0u;
    ep.out.all_vec = (ep.out.vec16 && oc.dev && oh.dev && op.dev) ? 1u : 0u;
    CK(ctx, cudaMemsetAsync(ctx->d_cta_counts, 0, ((size_t)l.grid + 2) * 8, ctx->stream));
    if (pipe) CK(ctx, launch_compact_pipe(validate, khi, fg, l, ctx->device, ctx->stream, enc, ep));
    else CK(ctx, launch_compact(false, validate, khi, pf, pc, l, ctx->stream, enc, ep));
    ctx->launches++;
    if (oe.dev) {
        if (!pipe) {
            // the kernel wrote tile-local first-entry indices: add the tiles' starts (their inclusive prefixes are in the descriptors
            // now).  The pipelined kernel knows a tile's start before it emits and writes final indices itself.
            const uint64_t slots_per_cta = (uint64_t)(csr ? cg.items_per_cta : fg.items_per_cta) * kRun;
            CK(ctx, launch_compact_fixup(csr ? ctx->d_win_offsets : nullptr, csr ? 0 : fg.W, ctx->n_reads, slots_per_cta, ctx->d_cta_counts,
                                         (uint64_t*)oe.dev, ctx->stream));
            ctx->launches++;
        }
        CK(ctx, cudaMemcpyAsync((uint64_t*)oe.dev + ctx->n_reads, d_total, 8, cudaMemcpyDeviceToDevice, ctx->stream));
        if (csr) {
            CK(ctx, launch_compact_backfill(ctx->d_win_offsets, ctx->n_reads, d_total, (uint64_t*)oe.dev, ctx->stream));
            ctx->launches++;
        }
    }
    CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    const uint64_t total = ctx->h_digest[3];
    *n_emitted = total;
    if (total > capacity) return fail(ctx, KMB_ERR_INVALID_ARG, "capacity %llu < %llu emitted k-mers", (unsigned long long)capacity,
                                      (unsigned long long)total);
    oc.bytes = oc.host ? total * 8 : 0; oh.bytes = oh.host ? total * 8 : 0; op.bytes = op.host ? total * 4 : 0;  // only what was emitted goes back
    if ((rc = out_finish(ctx, oc)) || (rc = out_finish(ctx, oh)) || (rc = out_finish(ctx, op)) || (rc = out_finish(ctx, oe))) return rc;
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

static int32_t in_prepare(kmb_ctx* ctx, int slot, const void* user, size_t bytes, const void** dev);

// ======================================================================= minimizers ("next" row N1)
extern "C" int32_t kmb_minimizers(kmb_ctx* ctx, uint32_t k, uint32_t w, uint32_t hash_k, uint32_t flags, uint64_t* mmer_out,
                                  uint32_t* pos_out) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    // get_kmer_u64 reads at most 64 bits (seq_vector.rs:96-99); k - w underflows for w > k (minimizers.rs:104); LexHasher
    // shifts by (32 - hash_k) * 2 (hash.rs:69)
    if (k < 1 || k > 32 || w < 1 || w > k || hash_k < 1 || hash_k > 32)
        return fail(ctx, KMB_ERR_PANIC, "need 1 <= w <= k <= 32 and 1 <= hash_k <= 32 (k=%u w=%u hash_k=%u)", k, w, hash_k);
    EncDesc enc;
    make_enc(KMB_ENC_ACGT, &enc, nullptr);
    uint64_t n_slots = 0;
    int32_t rc = num_slots(ctx, k, &n_slots);
    if (rc) return rc;
    OutBuf om, op;
    if ((rc = out_prepare(ctx, 0, mmer_out, n_slots * 8, &om))) return rc;
    if ((rc = out_prepare(ctx, 1, pos_out, n_slots * 4, &op))) return rc;
    if (n_slots) {
        MinParams ep{};
        ep.mc.wc = make_winconst(k, enc);
        ep.mc.w = w; ep.mc.m = k - w + 1;
        ep.mc.wmask = w >= 32 ? ~0ull : ((1ull << (2 * w)) - 1ull);
        const uint32_t hshift = hash_k < w ? 2 * (w - hash_k) : 0;  // the hash keeps only the first hash_k bases (hash.rs:69)
        ep.mc.hmask64 = ep.mc.wmask & ~((1ull << hshift) - 1ull);
        ep.mc.hmask32 = w <= 13 ? (uint32_t)ep.mc.hmask64 << 6 : (uint32_t)ep.mc.hmask64;
        const unsigned __int128 c96 = ((unsigned __int128)enc.cmask) | ((unsigned __int128)enc.cmask << 32) | ((unsigned __int128)enc.cmask << 64);
        const unsigned __int128 vm = c96 >> ep.mc.wc.shiftD;
        ep.mc.vm0 = (uint32_t)vm; ep.mc.vm1 = (uint32_t)(vm >> 32); ep.mc.vm2 = (uint32_t)(vm >> 64);
        ep.out.mmer = (uint64_t*)om.dev; ep.out.pos = (uint32_t*)op.dev;
        ep.out.vec_ok = ((((uintptr_t)om.dev | (uintptr_t)op.dev) & 31u) == 0) ? 1u : 0u;
        FixedGeom fg{};
        CsrGeom cg{};
        Launch l;
        const bool csr = ctx->d_offsets != nullptr, validate = !(flags & KMB_F_NO_VALIDATE) && !ctx->packed;
        if (!csr) {
            // (2048 items per tile: 86.4 instead of 85.3 % of the copy peak; the materialising K <= 32 kernels are best at 1024)
            if (!make_fixed_geom(ctx->d_bases, ctx->n_bytes, ctx->n_reads, ctx->fixed_len, k, 4, &fg, &l, ctx->stride_len, ctx->packed ? KMB_I_SEQVECTOR : KMB_I_ASCII,
                                 2 * kItemsPerCta))
                return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch (or a read above 4.29 Gbases)");
        } else if ((rc = make_csr_geom(ctx, k, 4, &cg, &l))) {
            return rc;
        }
        CK(ctx, launch_minimizers(validate, csr ? nullptr : &fg, csr ? &cg : nullptr, l, ctx->stream, enc, ep));
        ctx->launches++;
    }
    if ((rc = out_finish(ctx, om)) || (rc = out_finish(ctx, op))) return rc;
    if (om.host || op.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_minimizer_words(kmb_ctx* ctx, uint32_t k, uint32_t w, uint32_t hash_k, const uint64_t* words, uint64_t n,
                                       uint64_t* mmer_out, uint32_t* offset_out) {
    NEED_CTX(ctx);
    BIND(ctx);
    // sub_kmer_word asserts pos + width <= k (kmer.rs:155-157); LexHasher shift (hash.rs:69)
    if (k < 1 || k > 32 || w < 1 || w > k || hash_k < 1 || hash_k > 32)
        return fail(ctx, KMB_ERR_PANIC, "need 1 <= w <= k <= 32 and 1 <= hash_k <= 32 (k=%u w=%u hash_k=%u)", k, w, hash_k);
    if (n == 0) return KMB_OK;
    if (!words) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL input");
    const void* d_in;
    int32_t rc = in_prepare(ctx, 2, words, n * 8, &d_in);
    if (rc) return rc;
    OutBuf om, oo;
    if ((rc = out_prepare(ctx, 0, mmer_out, n * 8, &om))) return rc;
    if ((rc = out_prepare(ctx, 1, offset_out, n * 4, &oo))) return rc;
    const uint64_t ctas = (n + 255) / 256;
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    CK(ctx, launch_minimizer_words((const uint64_t*)d_in, n, k, w, hash_k, (uint64_t*)om.dev, (uint32_t*)oo.dev, ctx->stream));
    ctx->launches++;
    if ((rc = out_finish(ctx, om)) || (rc = out_finish(ctx, oo))) return rc;
    if (om.host || oo.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// ======================================================================= extract wide (extension, K <= 64)
extern "C" int32_t kmb_extract_canonical_wide(kmb_ctx* ctx, uint32_t k, int32_t enc_id, uint32_t flags,
                                              uint64_t* canon_out, uint64_t* hash_out, kmb_digest* digest) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (k < 1 || k > 64) return fail(ctx, KMB_ERR_INVALID_ARG, "k = %u: the two-word path supports 1 <= k <= 64", k);
    EncDesc enc;
    if (!make_enc(enc_id, &enc, nullptr)) return fail(ctx, KMB_ERR_INVALID_ARG, "enc 0x%x is not a Naive discriminant / Xor10", enc_id);
    uint64_t n_slots = 0;
    int32_t rc = num_slots(ctx, k, &n_slots);
    if (rc) return rc;
    OutBuf oc, oh;
    if ((rc = out_prepare(ctx, 0, canon_out, n_slots * 16, &oc))) return rc;
    if ((rc = out_prepare(ctx, 1, hash_out, n_slots * 16, &oh))) return rc;
    if (digest && (rc = digest_begin(ctx))) return rc;
    const bool validate = !(flags & KMB_F_NO_VALIDATE) && !ctx->packed;  // a packed store holds no invalid base
    if (n_slots) {
        const int nw32 = k <= 32 ? 2 : (k <= 48 ? 3 : 4);  // live 32-bit words of a k-mer
        uint32_t mask[4];
        for (int i = 0; i < 4; ++i) mask[i] = 2 * k > 32u * i ? mask32(2 * k - 32 * i) : 0u;
        WideParams ep{};
        ep.wc.K = k;
        ep.wc.shiftD = 2 * (16 * (nw32 + 1) - (ShapePair::kSpanSlots + k - 1));
        ep.wc.mask_a = mask[nw32 - 2]; ep.wc.mask_b = mask[nw32 - 1];
        ep.wc.cmask = enc.cmask; ep.wc.cm_a = enc.cmask & ep.wc.mask_a; ep.wc.cm_b = enc.cmask & ep.wc.mask_b;
        ep.wc.kmask = k >= 64 ? ~0ull : ((1ull << k) - 1ull);
        ep.out.canon = (uint64_t*)oc.dev; ep.out.hash = (uint64_t*)oh.dev; ep.out.digest = ctx->d_digest;
        ep.out.vec_ok = ((((uintptr_t)oc.dev | (uintptr_t)oh.dev) & 31u) == 0) ? 1u : 0u;
        FixedGeom fg{};
        CsrGeom cg{};
        Launch l;
        const bool csr = ctx->d_offsets != nullptr;
        if (!csr) {
            if (!make_fixed_geom(ctx->d_bases, ctx->n_bytes, ctx->n_reads, ctx->fixed_len, k, nw32 + 2, &fg, &l, ctx->stride_len, ctx->packed ? KMB_I_SEQVECTOR : KMB_I_ASCII))
                return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch (or a read above 4.29 Gbases)");
        } else if ((rc = make_csr_geom(ctx, k, nw32 + 2, &cg, &l))) {
            return rc;
        }
        const FixedGeom* pf = csr ? nullptr : &fg;
        const CsrGeom* pc = csr ? &cg : nullptr;
        const bool want_hash = oh.dev != nullptr;
        CK(ctx, launch_wide(nw32, validate, digest != nullptr, want_hash, pf, pc, l, ctx->stream, enc, ep));
        ctx->launches++;
    }
    if ((rc = out_finish(ctx, oc))) return rc;
    if ((rc = out_finish(ctx, oh))) return rc;
    if (digest) return digest_end(ctx, digest);
    if (oc.host || oh.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// ======================================================================= packed sequence store ("next" row N2)
__global__ void __launch_bounds__(256) base_starts_kernel(const uint64_t* word_offsets, uint64_t n, uint64_t* base_starts) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) base_starts[i] = word_offsets[i] * 32;  // 32 bases per u64 word
}

// SeqVector::get_kmer_u64 (naive_impl/seq_vector.rs:96-99) on (read, pos) pairs of a packed batch
__global__ void __launch_bounds__(256) packed_get_kmers_kernel(const uint64_t* words, uint64_t n_words, const uint64_t* offsets,
                                                               const uint64_t* base_starts, uint64_t fixed_len, uint64_t stride,
                                                               uint64_t n_reads, uint32_t k, const uint64_t* reads,
                                                               const uint64_t* pos, uint64_t n, uint64_t* out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = reads ? reads[i] : 0, p = pos[i];
    uint64_t v = ~0ull;
    if (r < n_reads) {
        const uint64_t len = offsets ? offsets[r + 1] - offsets[r] : fixed_len;
        if (p <= len && k <= len - p) {  // p + k <= len without the wrap-around of a huge p
            const uint64_t bit = ((base_starts ? base_starts[r] : r * stride) + p) * 2;
            const uint64_t wi = bit >> 6, sh = bit & 63;
            v = words[wi] >> sh;
            if (sh != 0 && wi + 1 < n_words) v |= words[wi + 1] << (64 - sh);
            if (k < 32) v &= (1ull << (2 * k)) - 1ull;
        }
    }
    out[i] = v;
}

static int32_t set_packed(kmb_ctx* ctx, const uint64_t* d_words, uint64_t n_words, const uint64_t* d_word_offsets) {
    if (ctx->d_offsets) {  // ragged: flat index of every read's first base in the padded space
        int32_t rc = grow(ctx, (void**)&ctx->own_base_starts, &ctx->own_base_starts_cap, (ctx->n_reads + 1) * 8);
        if (rc) return rc;
        base_starts_kernel<<<(unsigned)((ctx->n_reads + 1 + 255) / 256), 256, 0, ctx->stream>>>(d_word_offsets, ctx->n_reads + 1,
                                                                                             ctx->own_base_starts);
        CK(ctx, cudaGetLastError());
        ctx->launches++;
        ctx->d_base_starts = ctx->own_base_starts;
    }
    ctx->d_bases = (const uint8_t*)d_words;
    ctx->n_bytes = n_words * 8;
    ctx->n_bases_flat = n_words * 32;
    ctx->stride_len = (ctx->fixed_len + 31) / 32 * 32;
    ctx->packed = true;
    return KMB_OK;
}

extern "C" int32_t kmb_batch_repack(kmb_ctx* ctx, int32_t strict) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (ctx->packed) return KMB_OK;
    if (strict && ctx->n_bytes) {
        // SeqVector::from goes through Kmer::from, which panics on a non-ACGT byte (seq_vector.rs:236, naive_impl/mod.rs:35)
        uint64_t n_valid = 0;
        int32_t rc = kmb_extract_compact(ctx, 1, 0, nullptr, nullptr, nullptr, nullptr, 0, &n_valid);
        if (rc) return rc;
        if (n_valid != ctx->n_bytes) return fail(ctx, KMB_ERR_PANIC, "%llu bases are not ACGTacgt: SeqVector::from would panic",
                                                 (unsigned long long)(ctx->n_bytes - n_valid));
    }
    uint64_t n_words = 0;
    int32_t rc = kmb_pack_num_words(ctx, 64, &n_words);
    if (rc) return rc;
    if ((rc = grow(ctx, (void**)&ctx->own_packed, &ctx->own_packed_cap, n_words * 8 + 64))) return rc;
    uint64_t* d_woff = nullptr;
    if (ctx->d_offsets) {
        if ((rc = grow(ctx, &ctx->d_scratch[2], &ctx->scratch_cap[2], (ctx->n_reads + 1) * 8))) return rc;
        d_woff = (uint64_t*)ctx->d_scratch[2];
    }
    if ((rc = kmb_pack(ctx, KMB_ENC_ACGT, 64, ctx->own_packed, d_woff))) return rc;
    return set_packed(ctx, ctx->own_packed, n_words, d_woff);
}

extern "C" int32_t kmb_batch_attach_packed(kmb_ctx* ctx, const uint64_t* dev_words, uint64_t n_words, const uint64_t* dev_offsets,
                                           const uint64_t* dev_word_offsets, uint64_t n_reads, uint64_t fixed_len) {
    NEED_CTX(ctx);
    BIND(ctx);
    if ((dev_offsets != nullptr) == (fixed_len > 0) && n_reads) return fail(ctx, KMB_ERR_INVALID_ARG, "give either offsets or fixed_len > 0");
    if ((dev_offsets != nullptr) != (dev_word_offsets != nullptr)) return fail(ctx, KMB_ERR_INVALID_ARG, "ragged packed batches need offsets and word_offsets");
    if (n_words && !is_device_ptr(dev_words)) return fail(ctx, KMB_ERR_INVALID_ARG, "dev_words is not device memory");
    if (!dev_offsets && n_reads * ((fixed_len + 31) / 32) != n_words) return fail(ctx, KMB_ERR_INVALID_ARG, "n_words != n_reads * ceil(fixed_len / 32)");
    if (dev_offsets) {
        if (!is_device_ptr(dev_offsets) || !is_device_ptr(dev_word_offsets)) return fail(ctx, KMB_ERR_INVALID_ARG, "offsets are not device memory");
        int32_t rc = check_device_offsets(ctx, dev_word_offsets, nullptr, n_reads, n_words, true);  // word offsets: 0 .. n_words
        if (rc == KMB_OK) rc = check_device_offsets(ctx, dev_offsets, dev_word_offsets, n_reads, 0, false);
        if (rc) return rc;
    }
    drop_batch(ctx);
    ctx->d_offsets = dev_offsets; ctx->n_reads = n_reads; ctx->fixed_len = fixed_len;
    ctx->have_batch = true;
    return set_packed(ctx, dev_words, n_words, dev_word_offsets);
}

extern "C" int32_t kmb_packed_get_kmers(kmb_ctx* ctx, uint32_t k, const uint64_t* reads, const uint64_t* pos, uint64_t n, uint64_t* out) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (!ctx->packed) return fail(ctx, KMB_ERR_STATE, "the resident batch is not packed (kmb_batch_repack first)");
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: get_kmer_u64 reads at most 32 bases", k);
    if (n == 0) return KMB_OK;
    if (!pos || !out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    const void *d_reads = nullptr, *d_pos;
    int32_t rc;
    if (reads && (rc = in_prepare(ctx, 2, reads, n * 8, &d_reads))) return rc;
    if ((rc = in_prepare(ctx, 3, pos, n * 8, &d_pos))) return rc;
    OutBuf ob;
    if ((rc = out_prepare(ctx, 0, out, n * 8, &ob))) return rc;
    packed_get_kmers_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(
        (const uint64_t*)ctx->d_bases, ctx->n_bytes / 8, ctx->d_offsets, ctx->d_base_starts, ctx->fixed_len, ctx->stride_len, ctx->n_reads,
        k, (const uint64_t*)d_reads, (const uint64_t*)d_pos, n, (uint64_t*)ob.dev);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, ob))) return rc;
    if (ob.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// ----------------------------------------------------------------------- SeqVector growth and views
// SeqVector::with_capacity (seq_vector.rs:135-139): an empty packed sequence the context owns
extern "C" int32_t kmb_batch_new_packed(kmb_ctx* ctx, uint64_t capacity_bases) {
    NEED_CTX(ctx);
    BIND(ctx);
    drop_batch(ctx);
    int32_t rc = grow(ctx, (void**)&ctx->own_packed, &ctx->own_packed_cap, (capacity_bases + 31) / 32 * 8 + 64);
    if (rc) return rc;
    ctx->d_bases = (const uint8_t*)ctx->own_packed; ctx->d_offsets = nullptr;
    ctx->n_bytes = 0; ctx->n_reads = 1; ctx->fixed_len = 0; ctx->stride_len = 0; ctx->n_bases_flat = 0;
    ctx->packed = true; ctx->have_batch = true;
    return KMB_OK;
}

__global__ void __launch_bounds__(256) count_invalid_kernel(const uint8_t* chars, uint64_t n, unsigned long long* bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool inv = i < n && encode_binary_u8_dev(chars[i]) == ~0ull;
    const unsigned b = __ballot_sync(0xffffffffu, inv);
    if ((threadIdx.x & 31) == 0 && b) atomicAdd(bad, (unsigned long long)__popc(b));
}
// one thread per 64-bit word of the store that receives new bases: the word that was partly filled keeps its low bits
__global__ void __launch_bounds__(256) push_chars_kernel(uint64_t* words, uint64_t old_len, const uint8_t* chars, uint64_t n) {
    const uint64_t w0 = old_len / 32, wi = w0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t new_len = old_len + n;
    if (wi * 32 >= new_len) return;
    const uint64_t lo = max(wi * 32, old_len), hi = min(wi * 32 + 32, new_len);
    uint64_t v = (wi == w0 && (old_len & 31)) ? words[wi] : 0ull;
    for (uint64_t p = lo; p < hi; ++p) v |= (encode_binary_u8_dev(chars[p - old_len]) & 3ull) << (2 * (p & 31));
    words[wi] = v;
}

// SeqVector::push_chars (seq_vector.rs:141-161)
extern "C" int32_t kmb_packed_push_chars(kmb_ctx* ctx, const uint8_t* bases, uint64_t n) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (!ctx->packed || ctx->sliced || ctx->d_offsets || ctx->n_reads != 1 || ctx->d_bases != (const uint8_t*)ctx->own_packed)
        return fail(ctx, KMB_ERR_STATE, "push_chars needs a context-owned packed batch of ONE sequence (kmb_batch_new_packed, or kmb_batch_repack of one read)");
    if (n == 0) return KMB_OK;
    if (!bases) return fail(ctx, KMB_ERR_INVALID_ARG, "bases is NULL");
    const void* d_chars;
    int32_t rc = in_prepare(ctx, 1, bases, n, &d_chars);
    if (rc) return rc;
    // Kmer::from panics on a byte outside ACGTacgt (naive_impl/mod.rs:35): checked before anything is written
    CK(ctx, cudaMemsetAsync(ctx->d_digest + 3, 0, 8, ctx->stream));
    count_invalid_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>((const uint8_t*)d_chars, n, ctx->d_digest + 3);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, ctx->d_digest + 3, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->h_digest[3]) return fail(ctx, KMB_ERR_PANIC, "%llu bases are not ACGTacgt: SeqVector::push_chars would panic", ctx->h_digest[3]);
    const uint64_t old_len = ctx->fixed_len, new_len = old_len + n;
    const uint64_t old_words = (old_len + 31) / 32, new_words = (new_len + 31) / 32;
    if (ctx->own_packed_cap < new_words * 8 + 64) {  // grow geometrically, keep what is there
        uint64_t* bigger = nullptr;
        const size_t cap = std::max<size_t>(new_words * 8 + 64, 2 * ctx->own_packed_cap);
        CK(ctx, cudaMalloc((void**)&bigger, cap + 256));
        if (old_words) CK(ctx, cudaMemcpyAsync(bigger, ctx->own_packed, old_words * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaFree(ctx->own_packed));
        ctx->own_packed = bigger; ctx->own_packed_cap = cap;
    }
    const uint64_t touched = new_words - old_len / 32;
    push_chars_kernel<<<(unsigned)((touched + 255) / 256), 256, 0, ctx->stream>>>(ctx->own_packed, old_len, (const uint8_t*)d_chars, n);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    ctx->d_bases = (const uint8_t*)ctx->own_packed;
    ctx->fixed_len = new_len; ctx->stride_len = new_words * 32; ctx->n_bytes = new_words * 8; ctx->n_bases_flat = new_words * 32;
    ctx->win_valid = false;
    return KMB_OK;
}

// SeqVector::slice / SeqVectorSlice (seq_vector.rs:24-90): the batch becomes a view of bases [start, start + len) of one read
extern "C" int32_t kmb_batch_slice(kmb_ctx* ctx, uint64_t read, uint64_t start, uint64_t len) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (!ctx->packed) return fail(ctx, KMB_ERR_STATE, "the resident batch is not packed (kmb_batch_repack first)");
    if (!ctx->sliced) {
        ctx->parent_offsets = ctx->d_offsets; ctx->parent_base_starts = ctx->d_base_starts;
        ctx->parent_n_reads = ctx->n_reads; ctx->parent_fixed_len = ctx->fixed_len;
    }
    if (read >= ctx->parent_n_reads) return fail(ctx, KMB_ERR_INVALID_ARG, "read %llu out of range", (unsigned long long)read);
    uint64_t read_len = ctx->parent_fixed_len, flat0 = read * ctx->stride_len;
    if (ctx->parent_offsets) {  // ragged parent: this read's length and padded start live on the device
        uint64_t h[3];
        CK(ctx, cudaMemcpyAsync(h, ctx->parent_offsets + read, 16, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaMemcpyAsync(h + 2, ctx->parent_base_starts + read, 8, cudaMemcpyDeviceToHost, ctx->stream));
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        read_len = h[1] - h[0];
        flat0 = h[2];
    }
    // assert!(end <= self.len()) (seq_vector.rs:58); end - start underflows for start > end
    if (start > read_len || len > read_len - start)
        return fail(ctx, KMB_ERR_PANIC, "slice [%llu, %llu) exceeds the %llu bases of read %llu", (unsigned long long)start,
                    (unsigned long long)(start + len), (unsigned long long)read_len, (unsigned long long)read);
    if (!ctx->own_slice) CK(ctx, cudaMalloc((void**)&ctx->own_slice, 4 * 8 + 256));
    const uint64_t h[4] = {0, len, flat0 + start, flat0 + start + len};
    CK(ctx, cudaMemcpyAsync(ctx->own_slice, h, sizeof h, cudaMemcpyHostToDevice, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));  // h is a stack array
    ctx->d_offsets = ctx->own_slice; ctx->d_base_starts = ctx->own_slice + 2;
    ctx->n_reads = 1; ctx->fixed_len = 0;
    ctx->sliced = true; ctx->win_valid = false;
    return KMB_OK;
}

extern "C" int32_t kmb_batch_unslice(kmb_ctx* ctx) {
    NEED_CTX(ctx);
    NEED_BATCH(ctx);
    if (!ctx->sliced) return KMB_OK;
    ctx->d_offsets = ctx->parent_offsets; ctx->d_base_starts = ctx->parent_base_starts;
    ctx->n_reads = ctx->parent_n_reads; ctx->fixed_len = ctx->parent_fixed_len;
    ctx->sliced = false; ctx->win_valid = false;
    return KMB_OK;
}

// ======================================================================= final reduction across GPUs
// One process driving several GPUs (what a Rust host would do): NCCL all-reduce of u64 words over the contexts'
// devices.  libnccl is loaded on first use, so the library itself has no link-time dependency on it.
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*CommInitAll)(void**, int, const int*) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    std::vector<int> devices;  // the communicator clique currently held
    std::vector<void*> comms;
    std::mutex mu;
};
NcclApi g_nccl;
constexpr int kNcclUint64 = 5, kNcclSum = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h; stable ABI)

bool nccl_load(NcclApi& a) {
    if (a.lib) return true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (a.lib) break;
    }
    if (!a.lib) return false;
    a.CommInitAll = (decltype(a.CommInitAll))dlsym(a.lib, "ncclCommInitAll");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(a.lib, "ncclCommDestroy");
    a.AllReduce = (decltype(a.AllReduce))dlsym(a.lib, "ncclAllReduce");
    a.GroupStart = (decltype(a.GroupStart))dlsym(a.lib, "ncclGroupStart");
    a.GroupEnd = (decltype(a.GroupEnd))dlsym(a.lib, "ncclGroupEnd");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(a.lib, "ncclGetErrorString");
    if (a.CommInitAll && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString) return true;
    dlclose(a.lib);
    a.lib = nullptr;
    return false;
}
}  // namespace

extern "C" int32_t kmb_allreduce_u64(kmb_ctx* const* ctxs, int32_t n_ctx, uint64_t* const* dev_bufs, uint64_t count) {
    if (n_ctx < 1 || !ctxs || !dev_bufs) return fail(nullptr, KMB_ERR_INVALID_ARG, "need n_ctx >= 1 contexts and buffers");
    for (int i = 0; i < n_ctx; ++i)
        if (!ctxs[i] || (count && !dev_bufs[i])) return fail(nullptr, KMB_ERR_INVALID_ARG, "NULL context or buffer at index %d", i);
    kmb_ctx* c0 = ctxs[0];
    if (count == 0) return KMB_OK;
    if (n_ctx == 1) { BIND(c0); CK(c0, cudaStreamSynchronize(c0->stream)); return KMB_OK; }
    std::vector<int> devs(n_ctx);
    for (int i = 0; i < n_ctx; ++i) {
        devs[i] = ctxs[i]->device;
        for (int j = 0; j < i; ++j)
            if (devs[j] == devs[i]) return fail(c0, KMB_ERR_INVALID_ARG, "contexts %d and %d share device %d: one context per GPU", j, i, devs[i]);
    }
    std::lock_guard<std::mutex> lock(g_nccl.mu);
    if (!nccl_load(g_nccl)) return fail(c0, KMB_ERR_STATE, "libnccl.so.2 could not be loaded: %s", dlerror());
#define NK(call)                                                                                          \
    do {                                                                                                  \
        int r_ = (call);                                                                                  \
        if (r_ != 0) return fail(c0, KMB_ERR_CUDA, "%s failed: %s", #call, g_nccl.GetErrorString(r_));   \
    } while (0)
    if (g_nccl.devices != devs) {  // a new clique: (re)build the communicators
        for (void* c : g_nccl.comms) g_nccl.CommDestroy(c);
        g_nccl.comms.assign(n_ctx, nullptr);
        g_nccl.devices.clear();
        NK(g_nccl.CommInitAll(g_nccl.comms.data(), n_ctx, devs.data()));
        g_nccl.devices = devs;
    }
    NK(g_nccl.GroupStart());
    for (int i = 0; i < n_ctx; ++i) {
        cudaSetDevice(devs[i]);
        int r_ = g_nccl.AllReduce(dev_bufs[i], dev_bufs[i], count, kNcclUint64, kNcclSum, g_nccl.comms[i], ctxs[i]->stream);
        if (r_ != 0) { g_nccl.GroupEnd(); return fail(c0, KMB_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r_)); }
        ctxs[i]->launches++;
    }
    NK(g_nccl.GroupEnd());
#undef NK
    for (int i = 0; i < n_ctx; ++i) {
        cudaSetDevice(devs[i]);
        CK(ctxs[i], cudaStreamSynchronize(ctxs[i]->stream));
    }
    return KMB_OK;
}

// ======================================================================= host ingest ("next" row N4)
// FASTA / FASTQ text -> concatenated bases + CSR offsets.  Host-side I/O plumbing on the caller's side of the
// boundary (no k-mer arithmetic happens here); the bases go to the device untouched, so lower case, N and IUPAC
// codes are handled by the kernels exactly as the reference handles them.
namespace {
struct FastxSink {
    uint8_t* bases;      // may be NULL (counting pass)
    uint64_t* offsets;   // may be NULL
    uint64_t bases_cap, reads_cap;
    uint64_t n_bases = 0, n_reads = 0;
    bool overflow = false;
    uint64_t shift = 0;  // added to the offsets written (a chunk of a text parsed in parallel)
    void begin_read() {
        if (offsets) { if (n_reads < reads_cap) offsets[n_reads] = n_bases + shift; else overflow = true; }
        ++n_reads;
    }
    void append(const char* p, size_t n) {
        if (bases) { if (n_bases + n <= bases_cap) memcpy(bases + n_bases, p, n); else overflow = true; }
        n_bases += n;
    }
};

inline const char* line_end(const char* p, const char* end) {
    const void* q = memchr(p, '\n', (size_t)(end - p));
    return q ? (const char*)q : end;
}
inline size_t trim_cr(const char* p, const char* e) { return (e > p && e[-1] == '\r') ? (size_t)(e - p - 1) : (size_t)(e - p); }

// returns NULL on success, else a message
const char* parse_fastx(const char* text, uint64_t n, FastxSink& out) {
    const char *p = text, *end = text + n;
    while (p < end && (*p == '\n' || *p == '\r' || *p == ' ' || *p == '\t')) ++p;
    if (p == end) return nullptr;
    if (*p == '>') {  // FASTA: header line, then sequence lines up to the next '>'
        while (p < end) {
            if (*p != '>') return "FASTA: expected '>' at the start of a record";
            p = line_end(p, end);
            if (p < end) ++p;
            out.begin_read();
            while (p < end && *p != '>') {
                const char* e = line_end(p, end);
                out.append(p, trim_cr(p, e));
                p = e < end ? e + 1 : end;
            }
        }
    } else if (*p == '@') {  // FASTQ: four lines per record
        while (p < end) {
            if (*p == '\n' || *p == '\r') { ++p; continue; }  // blank lines between / after records
            if (*p != '@') return "FASTQ: expected '@' at the start of a record";
            p = line_end(p, end);
            if (p == end) return "FASTQ: truncated record (no sequence line)";
            ++p;
            const char* e = line_end(p, end);
            const size_t len = trim_cr(p, e);
            out.begin_read();
            out.append(p, len);
            if (e == end) return "FASTQ: truncated record (no '+' line)";
            p = e + 1;
            if (p >= end || *p != '+') return "FASTQ: expected '+' on the third line of a record";
            p = line_end(p, end);
            if (p == end) return "FASTQ: truncated record (no quality line)";
            ++p;
            e = line_end(p, end);
            if (trim_cr(p, e) != len) return "FASTQ: quality and sequence lengths differ";
            p = e < end ? e + 1 : end;
        }
    } else {
        return "neither FASTA ('>') nor FASTQ ('@')";
    }
    return nullptr;
}

// First record start at or after p (NULL: none before `end`).  FASTA: a line that starts with '>'.  FASTQ: a line that
// starts with '@' AND whose second-next line starts with '+' -- a quality line may start with '@' too, but then the
// second-next line is the following record's sequence, which never starts with '+'.
const char* next_record(const char* text, const char* p, const char* end, char mark) {
    if (p > text) {  // move to the start of the next line
        p = line_end(p - 1, end);
        if (p == end) return nullptr;
        ++p;
    }
    while (p < end) {
        if (*p == mark) {
            if (mark == '>') return p;
            const char* l2 = line_end(p, end);
            const char* l3 = l2 < end ? line_end(l2 + 1, end) : end;
            if (l3 < end && l3 + 1 < end && l3[1] == '+') return p;
        }
        p = line_end(p, end);
        if (p < end) ++p;
    }
    return nullptr;
}

// Parse with up to n_threads threads: the text is cut at record starts, every chunk is parsed twice (count, then fill at
// the prefix-summed positions).  Identical output to the single-threaded parse.
const char* parse_fastx_parallel(const char* text, uint64_t n, FastxSink& out, unsigned n_threads) {
    const char *p0 = text, *end = text + n;
    while (p0 < end && (*p0 == '\n' || *p0 == '\r' || *p0 == ' ' || *p0 == '\t')) ++p0;
    if (n_threads < 2 || p0 == end || (*p0 != '>' && *p0 != '@')) return parse_fastx(text, n, out);
    const char mark = *p0;
    std::vector<const char*> cut{p0};
    for (unsigned t = 1; t < n_threads; ++t) {
        const char* want = text + n / n_threads * t;
        if (want <= cut.back()) continue;
        const char* q = next_record(text, want, end, mark);
        if (q && q > cut.back()) cut.push_back(q);
    }
    cut.push_back(end);
    const size_t nc = cut.size() - 1;
    if (nc < 2) return parse_fastx(text, n, out);
    std::vector<FastxSink> part(nc, FastxSink{nullptr, nullptr, 0, 0});
    std::vector<const char*> err(nc, nullptr);
    // A std::thread constructor can throw (std::system_error) after others have started: whatever happens, every started
    // thread is joined before the vector goes away (a joinable thread's destructor would call std::terminate); the chunks
    // whose thread never started run here.
    auto run = [&](auto&& fn) {
        std::vector<std::thread> th;
        th.reserve(nc);
        size_t started = 1;
        try {
            for (; started < nc; ++started) th.emplace_back(fn, started);
        } catch (...) {
        }
        fn(0);
        for (size_t c = started; c < nc; ++c) fn(c);
        for (auto& t : th) t.join();
    };
    run([&](size_t c) { err[c] = parse_fastx(cut[c], (uint64_t)(cut[c + 1] - cut[c]), part[c]); });
    for (size_t c = 0; c < nc; ++c)
        if (err[c]) return err[c];
    std::vector<uint64_t> base0(nc + 1, 0), read0(nc + 1, 0);
    for (size_t c = 0; c < nc; ++c) { base0[c + 1] = base0[c] + part[c].n_bases; read0[c + 1] = read0[c] + part[c].n_reads; }
    out.n_bases = base0[nc];
    out.n_reads = read0[nc];
    if ((out.bases && out.n_bases > out.bases_cap) || (out.offsets && out.n_reads > out.reads_cap)) { out.overflow = true; return nullptr; }
    if (!out.bases && !out.offsets) return nullptr;  // sizing call
    run([&](size_t c) {
        FastxSink s{out.bases ? out.bases + base0[c] : nullptr, out.offsets ? out.offsets + read0[c] : nullptr, part[c].n_bases,
                    part[c].n_reads};
        s.shift = base0[c];
        parse_fastx(cut[c], (uint64_t)(cut[c + 1] - cut[c]), s);
    });
    return nullptr;
}
}  // namespace

extern "C" int32_t kmb_parse_fastx(const char* text, uint64_t n_bytes, uint8_t* bases_out, uint64_t bases_cap, uint64_t* offsets_out,
                                   uint64_t reads_cap, uint64_t* n_reads, uint64_t* n_bases) {
    if (n_bytes && !text) return fail(nullptr, KMB_ERR_INVALID_ARG, "text is NULL");
    FastxSink sink{bases_out, offsets_out, bases_cap, reads_cap};
    // one thread per ~4 MiB of text, at most the hardware's; KMB_PARSE_THREADS overrides (tests)
    unsigned n_threads = (unsigned)std::min<uint64_t>(std::max(1u, std::thread::hardware_concurrency()), n_bytes >> 22);
    if (const char* env = getenv("KMB_PARSE_THREADS")) n_threads = (unsigned)atoi(env);
    const char* err = nullptr;
    try {  // nothing may unwind across the C ABI (std::bad_alloc from the bookkeeping vectors)
        err = parse_fastx_parallel(text, n_bytes, sink, std::min(n_threads, 64u));
    } catch (const std::bad_alloc&) {
        return fail(nullptr, KMB_ERR_NOMEM, "out of host memory while parsing");
    } catch (...) {
        return fail(nullptr, KMB_ERR_INVALID_ARG, "unexpected failure while parsing");
    }
    if (err) return fail(nullptr, KMB_ERR_INVALID_ARG, "%s", err);
    if (n_reads) *n_reads = sink.n_reads;
    if (n_bases) *n_bases = sink.n_bases;
    if (offsets_out) { if (sink.n_reads < reads_cap) offsets_out[sink.n_reads] = sink.n_bases; else sink.overflow = true; }
    if (sink.overflow) return fail(nullptr, KMB_ERR_INVALID_ARG, "output capacity too small: %llu reads, %llu bases",
                                   (unsigned long long)sink.n_reads, (unsigned long long)sink.n_bases);
    return KMB_OK;
}

extern "C" int32_t kmb_batch_ingest_fastx(kmb_ctx* ctx, const char* text, uint64_t n_bytes, uint64_t* n_reads_out, uint64_t* n_bases_out) {
    NEED_CTX(ctx);
    BIND(ctx);
    uint64_t n_reads = 0, n_bases = 0;
    if (kmb_parse_fastx(text, n_bytes, nullptr, 0, nullptr, 0, &n_reads, &n_bases) != KMB_OK) return fail(ctx, KMB_ERR_INVALID_ARG, "%s", g_err.c_str());
    // parse straight into pinned memory, then one DMA each for the bases and the offsets
    uint8_t* h_bases = nullptr;
    uint64_t* h_offs = nullptr;
    CK(ctx, cudaHostAlloc((void**)&h_bases, n_bases ? n_bases : 1, cudaHostAllocDefault));
    cudaError_t e = cudaHostAlloc((void**)&h_offs, (n_reads + 1) * 8, cudaHostAllocDefault);
    if (e != cudaSuccess) { cudaFreeHost(h_bases); CK(ctx, e); }
    int32_t rc = kmb_parse_fastx(text, n_bytes, h_bases, n_bases, h_offs, n_reads + 1, nullptr, nullptr);
    if (rc == KMB_OK) rc = kmb_batch_upload(ctx, h_bases, n_bases, h_offs, n_reads, 0);
    if (rc == KMB_OK) { cudaError_t s = cudaStreamSynchronize(ctx->stream); if (s != cudaSuccess) rc = fail(ctx, KMB_ERR_CUDA, "%s", cudaGetErrorString(s)); }
    cudaFreeHost(h_bases);
    cudaFreeHost(h_offs);
    if (rc != KMB_OK) return rc;
    if (n_reads_out) *n_reads_out = n_reads;
    if (n_bases_out) *n_bases_out = n_bases;
    return KMB_OK;
}

// ======================================================================= batched Encoding<P,B>
static bool word_bits_ok(uint32_t wb) { return wb == 8 || wb == 16 || wb == 32 || wb == 64 || wb == 128; }

static int32_t pack_total_words(kmb_ctx* ctx, uint32_t word_bits, uint64_t* total, uint64_t** d_word_offsets) {
    const uint32_t bpw = word_bits / 2;
    if (!ctx->d_offsets) {
        *total = ctx->n_reads * ((ctx->fixed_len + bpw - 1) / bpw);
        if (d_word_offsets) *d_word_offsets = nullptr;
        return KMB_OK;
    }
    int32_t rc = grow(ctx, &ctx->d_scratch[3], &ctx->scratch_cap[3], (ctx->n_reads + 1) * 8);
    if (rc) return rc;
    uint64_t* d = (uint64_t*)ctx->d_scratch[3];
    if ((rc = scan_counts(ctx, 1, bpw, d))) return rc;
    CK(ctx, cudaMemcpyAsync(ctx->h_digest + 3, d + ctx->n_reads, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    *total = ctx->h_digest[3];
    if (d_word_offsets) *d_word_offsets = d;
    return KMB_OK;
}

extern "C" int32_t kmb_pack_num_words(kmb_ctx* ctx, uint32_t word_bits, uint64_t* n_words) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (!word_bits_ok(word_bits) || !n_words) return fail(ctx, KMB_ERR_INVALID_ARG, "word_bits must be 8/16/32/64/128");
    return pack_total_words(ctx, word_bits, n_words, nullptr);
}

extern "C" int32_t kmb_pack(kmb_ctx* ctx, int32_t enc_id, uint32_t word_bits, void* words_out, uint64_t* word_offsets_out) {
    NEED_CTX(ctx);
    BIND(ctx);
    NEED_BATCH(ctx);
    if (!word_bits_ok(word_bits)) return fail(ctx, KMB_ERR_INVALID_ARG, "word_bits must be 8/16/32/64/128");
    if (ctx->packed) return fail(ctx, KMB_ERR_STATE, "the resident batch is already packed");
    PackParams p{};
    if (!make_enc(enc_id, &p.enc, nullptr)) return fail(ctx, KMB_ERR_INVALID_ARG, "enc 0x%x is not a Naive discriminant / Xor10", enc_id);
    uint64_t total = 0;
    uint64_t* d_woff = nullptr;
    int32_t rc = pack_total_words(ctx, word_bits, &total, &d_woff);
    if (rc) return rc;
    const uint32_t word_bytes = word_bits / 8, bpw = word_bits / 2;
    OutBuf ob;
    if ((rc = out_prepare(ctx, 0, words_out, total * word_bytes, &ob))) return rc;
    if (total && words_out) {
        p.bases = ctx->d_bases; p.offsets = ctx->d_offsets; p.word_offsets = d_woff; p.n_reads = ctx->n_reads;
        p.L = ctx->fixed_len; p.word_bytes = word_bytes; p.bases_per_word = bpw; p.out = (uint8_t*)ob.dev;
        const uint64_t obr = ((ctx->fixed_len + bpw - 1) / bpw) * word_bytes;  // bytes of one read's packed region
        if (!ctx->d_offsets && ((uintptr_t)ob.dev & 3u) == 0 && ctx->fixed_len < (1ull << 30)) {
            // tiled kernel (aligned loads, coalesced 4-byte stores)
            PackTileParams t{};
            t.bases = ctx->d_bases; t.n_bytes = ctx->n_bytes; t.L = ctx->fixed_len; t.L32 = (uint32_t)ctx->fixed_len;
            t.obr = (uint32_t)obr;
            t.total_bytes = ctx->n_reads * obr;
            t.obr_magic = t.obr > 1 ? (uint32_t)((1ull << 32) / t.obr + 1) : 0;
            t.obr_magic64 = (t.obr > 1 && (double)t.total_bytes * (double)t.obr < 9.0e18) ? (~0ull / t.obr + 1) : 0;
            t.out = (uint8_t*)ob.dev; t.enc = p.enc;
            const uint64_t ctas = (t.total_bytes + 4 * kPackGroups - 1) / (4 * kPackGroups);
            if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
            const size_t smem = (size_t)(kPackGroups + 10) * sizeof(uint2);
            auto kern = (t.obr & 3u) == 0 ? pack_tile_kernel<0> : t.obr >= 4 ? pack_tile_kernel<1> : pack_tile_kernel<2>;
            if (smem > 48 * 1024) CK(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)ctas, 256, smem, ctx->stream>>>(t);
        } else if (!ctx->d_offsets) {
            p.out_bytes_per_read = ((ctx->fixed_len + bpw - 1) / bpw) * word_bytes;
            const uint64_t gpr = (p.out_bytes_per_read + 3) / 4;
            const uint64_t threads = ctx->n_reads * gpr;
            const uint64_t ctas = (threads + 255) / 256;
            if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
            pack_fixed_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>(p, gpr);
        } else {
            const uint64_t ctas = (ctx->n_reads * 32 + 255) / 256;
            if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
            pack_csr_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>(p);
        }
        CK(ctx, cudaGetLastError());
        ctx->launches++;
    }
    if ((rc = out_finish(ctx, ob))) return rc;
    if (word_offsets_out) {
        if (d_woff) {
            CK(ctx, cudaMemcpyAsync(word_offsets_out, d_woff, (ctx->n_reads + 1) * 8, cudaMemcpyDefault, ctx->stream));
        } else {
            return fail(ctx, KMB_ERR_STATE, "fixed-length batch: word offset of read r is r * ceil(L / (word_bits/2))");
        }
    }
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// in-buffer: device pointer used in place, host pointer staged into scratch
static int32_t in_prepare(kmb_ctx* ctx, int slot, const void* user, size_t bytes, const void** dev) {
    *dev = user;
    if (!user || bytes == 0 || is_device_ptr(user)) return KMB_OK;
    int32_t rc = grow(ctx, &ctx->d_scratch[slot], &ctx->scratch_cap[slot], bytes);
    if (rc) return rc;
    if ((rc = stage_h2d(ctx, ctx->d_scratch[slot], user, bytes))) return rc;
    *dev = ctx->d_scratch[slot];
    return KMB_OK;
}

static int32_t unpack_impl(kmb_ctx* ctx, int32_t enc_id, uint32_t word_bits, const void* words_in, uint64_t n_items,
                           uint32_t words_per_item, uint32_t bases_per_item, uint8_t* bases_out, uint32_t case_bit);

extern "C" int32_t kmb_unpack(kmb_ctx* ctx, int32_t enc_id, uint32_t word_bits, const void* words_in, uint64_t n_items,
                              uint32_t words_per_item, uint32_t bases_per_item, uint8_t* bases_out) {
    return unpack_impl(ctx, enc_id, word_bits, words_in, n_items, words_per_item, bases_per_item, bases_out, 0u);
}

// `impl From<Kmer> for String` (naive_impl/kmer.rs:196-207): BASE_TABLE = ['a','c','g','t'] (kmer.rs:24), base 0 first
extern "C" int32_t kmb_words_to_strings(kmb_ctx* ctx, uint32_t k, const uint64_t* words, uint64_t n, uint8_t* bases_out) {
    NEED_CTX(ctx);
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    return unpack_impl(ctx, KMB_ENC_ACGT, 64, words, n, 1, k, bases_out, 0x20202020u);
}

static int32_t unpack_impl(kmb_ctx* ctx, int32_t enc_id, uint32_t word_bits, const void* words_in, uint64_t n_items,
                           uint32_t words_per_item, uint32_t bases_per_item, uint8_t* bases_out, uint32_t case_bit) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (!word_bits_ok(word_bits)) return fail(ctx, KMB_ERR_INVALID_ARG, "word_bits must be 8/16/32/64/128");
    EncDesc enc;
    uint32_t dec = 0;
    if (!make_enc(enc_id, &enc, &dec)) return fail(ctx, KMB_ERR_INVALID_ARG, "enc 0x%x is not a Naive discriminant / Xor10", enc_id);
    dec |= case_bit;  // the four letters of the decode table, upper case by default
    const uint64_t in_bytes_per_item = (uint64_t)words_per_item * (word_bits / 8);
    if (bases_per_item > in_bytes_per_item * 4) return fail(ctx, KMB_ERR_INVALID_ARG, "bases_per_item exceeds the array capacity");
    if (n_items == 0 || bases_per_item == 0) return KMB_OK;
    if (!words_in || !bases_out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    const void* d_in;
    int32_t rc = in_prepare(ctx, 1, words_in, n_items * in_bytes_per_item, &d_in);
    if (rc) return rc;
    OutBuf ob;
    if ((rc = out_prepare(ctx, 0, bases_out, n_items * bases_per_item, &ob))) return rc;
    if (bases_per_item <= (uint32_t)kUnpackTile && ((uintptr_t)d_in & 3u) == 0 && !getenv("KMB_UNPACK_TILE")) {
        // output-space kernel: one aligned 16-byte block of text per thread step, four letters per PRMT
        const uint64_t total = n_items * bases_per_item;
        const uint64_t n_blocks = (((uintptr_t)ob.dev & 15u) + total + 15) / 16;
        const uint64_t ctas = (n_blocks + 256 * kUnpackBlocksPerThread - 1) / (256 * kUnpackBlocksPerThread);
        if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
        const uint32_t magic = bases_per_item > 1 ? (uint32_t)((1ull << 32) / bases_per_item + 1) : 0u;
        const uint64_t magic64 = bases_per_item > 1 ? ~0ull / bases_per_item + 1 : 0ull;  // exact while letters * bases_per_item < 2^64
        unpack_flat_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint8_t*)d_in, n_items * in_bytes_per_item, n_items,
                                                                   (uint32_t)in_bytes_per_item, bases_per_item, magic, magic64, dec,
                                                                   (uint8_t*)ob.dev, total, n_blocks);
    } else if (bases_per_item <= (uint32_t)kUnpackTile) {
        // staged through shared memory: aligned 16-byte stores whatever the item length
        const uint32_t ipc = (uint32_t)kUnpackTile / bases_per_item;
        const uint64_t ctas = (n_items + ipc - 1) / ipc;
        if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
        unpack_tile_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint8_t*)d_in, n_items, (uint32_t)in_bytes_per_item,
                                                                   bases_per_item, dec, (uint8_t*)ob.dev, ipc);
    } else {
        const uint64_t threads = n_items * ((bases_per_item + 3) / 4);
        const uint64_t ctas = (threads + 255) / 256;
        if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
        unpack_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint8_t*)d_in, n_items, (uint32_t)in_bytes_per_item,
                                                              bases_per_item, dec, (uint8_t*)ob.dev);
    }
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, ob))) return rc;
    if (ob.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_revcomp_words(kmb_ctx* ctx, int32_t enc_id, uint32_t k, uint32_t word_bits, uint32_t words_per_item,
                                     const void* words_in, void* words_out, uint64_t n_items) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (!word_bits_ok(word_bits)) return fail(ctx, KMB_ERR_INVALID_ARG, "word_bits must be 8/16/32/64/128");
    EncDesc enc;
    if (!make_enc(enc_id, &enc, nullptr)) return fail(ctx, KMB_ERR_INVALID_ARG, "enc 0x%x is not a Naive discriminant / Xor10", enc_id);
    const uint64_t item_bytes = (uint64_t)words_per_item * (word_bits / 8);
    // K == 0 underflows K*2-2 and 2K > capacity trips the get_bits range assert (encoding/naive.rs:138-147)
    if (k < 1 || 2ull * k > item_bytes * 8) return fail(ctx, KMB_ERR_PANIC, "k = %u does not fit %u x u%u", k, words_per_item, word_bits);
    if (item_bytes > 32) return fail(ctx, KMB_ERR_INVALID_ARG, "arrays above 256 bits are not supported");
    if (n_items == 0) return KMB_OK;
    if (!words_in || !words_out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    const void* d_in;
    int32_t rc = in_prepare(ctx, 1, words_in, n_items * item_bytes, &d_in);
    if (rc) return rc;
    OutBuf ob;
    if ((rc = out_prepare(ctx, 0, words_out, n_items * item_bytes, &ob))) return rc;
    const uint64_t ctas = (n_items + 256 * kRevItems - 1) / (256 * kRevItems);
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    const unsigned grid = (unsigned)ctas;
    const uint8_t* in8 = (const uint8_t*)d_in;
    uint8_t* out8 = (uint8_t*)ob.dev;
    const uint32_t ib = (uint32_t)item_bytes;
    if (item_bytes <= 4) revcomp_items_kernel<1><<<grid, 256, 0, ctx->stream>>>(in8, out8, n_items, ib, k, enc.cmask);
    else if (item_bytes <= 8) revcomp_items_kernel<2><<<grid, 256, 0, ctx->stream>>>(in8, out8, n_items, ib, k, enc.cmask);
    else if (item_bytes <= 16) revcomp_items_kernel<4><<<grid, 256, 0, ctx->stream>>>(in8, out8, n_items, ib, k, enc.cmask);
    else revcomp_items_kernel<8><<<grid, 256, 0, ctx->stream>>>(in8, out8, n_items, ib, k, enc.cmask);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, ob))) return rc;
    if (ob.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// ======================================================================= naive_impl::Kmer word ops
template <int OP>
static int32_t word_op(kmb_ctx* ctx, uint32_t k, const uint64_t* in, const uint64_t* other, uint64_t* out, uint8_t* out8,
                       uint64_t n) {
    NEED_CTX(ctx);
    BIND(ctx);
    // 2*(32-k) shifts: k == 0 overflows, k > 32 is unrepresentable (naive_impl/kmer.rs:133)
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    if (n == 0) return KMB_OK;
    if (!in) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL input");
    const void *d_in, *d_other = nullptr;
    int32_t rc = in_prepare(ctx, 2, in, n * 8, &d_in);
    if (rc) return rc;
    if (other && (rc = in_prepare(ctx, 3, other, n * 8, &d_other))) return rc;
    OutBuf o64, o8;
    if ((rc = out_prepare(ctx, 0, out, n * 8, &o64))) return rc;
    if ((rc = out_prepare(ctx, 1, out8, n, &o8))) return rc;
    const uint64_t ctas = ((n + 3) / 4 + 255) / 256;
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint32_t vec_ok = ((((uintptr_t)d_in | (uintptr_t)d_other | (uintptr_t)o64.dev) & 31u) == 0 && ((uintptr_t)o8.dev & 3u) == 0) ? 1u : 0u;
    word_op_kernel<OP><<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint64_t*)d_in, (const uint64_t*)d_other,
                                                               (uint64_t*)o64.dev, (uint8_t*)o8.dev, n, k, mask, vec_ok);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, o64))) return rc;
    if ((rc = out_finish(ctx, o8))) return rc;
    if (o64.host || o8.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_reverse_complement_words(kmb_ctx* ctx, uint32_t k, const uint64_t* in, uint64_t* out, uint64_t n) {
    if (ctx && n && !out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL output");
    return word_op<0>(ctx, k, in, nullptr, out, nullptr, n);
}
extern "C" int32_t kmb_canonical_words(kmb_ctx* ctx, uint32_t k, const uint64_t* in, uint64_t* canon_out,
                                       uint8_t* is_canonical_out, uint64_t n) {
    return word_op<1>(ctx, k, in, nullptr, canon_out, is_canonical_out, n);
}
extern "C" int32_t kmb_lexhash_words(kmb_ctx* ctx, uint32_t k, const uint64_t* in, uint64_t* out, uint64_t n) {
    if (ctx && n && !out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL output");
    return word_op<2>(ctx, k, in, nullptr, out, nullptr, n);
}
extern "C" int32_t kmb_match_words(kmb_ctx* ctx, uint32_t k, const uint64_t* words, const uint64_t* others,
                                   uint8_t* match_out, uint64_t n) {
    if (ctx && n && (!others || !match_out)) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    return word_op<3>(ctx, k, words, others, nullptr, match_out, n);
}

// ======================================================================= small Kmer / CanonicalKmer accessors, batched
template <int OP>
static int32_t shift_op(kmb_ctx* ctx, uint32_t k, const uint64_t* fw_in, const uint64_t* rc_in, const uint8_t* bases, int32_t ascii,
                        uint64_t* fw_out, uint64_t* rc_out, uint8_t* out8, uint64_t n, uint32_t a, uint64_t b_mask) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    if (n == 0) return KMB_OK;
    if (!fw_in) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL input");
    const void *d_fw, *d_rc = nullptr, *d_bases = nullptr;
    int32_t rc = in_prepare(ctx, 2, fw_in, n * 8, &d_fw);
    if (rc) return rc;
    if (rc_in && (rc = in_prepare(ctx, 3, rc_in, n * 8, &d_rc))) return rc;
    OutBuf o_fw, o_rc, o8;
    if ((rc = out_prepare(ctx, 0, fw_out, n * 8, &o_fw))) return rc;
    if ((rc = out_prepare(ctx, 1, rc_out, n * 8, &o_rc))) return rc;
    // bases (n bytes in) and the u8 output share one scratch block when both live on the host: [bases | out8]
    uint8_t* d_o8 = out8;
    const bool bases_host = bases && !is_device_ptr(bases), o8_host = out8 && !is_device_ptr(out8);
    if (bases_host || o8_host) {
        void* blk = nullptr;
        CK(ctx, cudaMallocAsync(&blk, 2 * n + 16, ctx->stream));
        if (bases_host) { rc = stage_h2d(ctx, blk, bases, n); d_bases = blk; }
        if (o8_host) d_o8 = (uint8_t*)blk + n;
        if (rc) { cudaFreeAsync(blk, ctx->stream); return rc; }
        o8.user = blk;  // remembered for the release below
    }
    if (bases && !bases_host) d_bases = bases;
    const uint64_t mask = k >= 32 ? ~0ull : ((1ull << (2 * k)) - 1ull);
    const uint64_t ctas = (n + 255) / 256;
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    kmer_shift_kernel<OP><<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint64_t*)d_fw, (const uint64_t*)d_rc, (const uint8_t*)d_bases,
                                                                  (uint32_t)(ascii != 0), (uint64_t*)o_fw.dev, (uint64_t*)o_rc.dev, d_o8, n, k,
                                                                  mask, a, b_mask);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, o_fw)) || (rc = out_finish(ctx, o_rc))) return rc;
    if (o8_host) CK(ctx, cudaMemcpyAsync(out8, d_o8, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (o8.user) CK(ctx, cudaFreeAsync(o8.user, ctx->stream));
    if (o_fw.host || o_rc.host || o8_host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_sub_kmer_words(kmb_ctx* ctx, uint32_t k, uint32_t pos, uint32_t width, const uint64_t* in, uint64_t* out, uint64_t n) {
    // assert!(pos < k); assert!(pos + width <= k);  (naive_impl/kmer.rs:155-157)
    if (ctx && !(pos < k && pos + width <= k)) return fail(ctx, KMB_ERR_PANIC, "sub_kmer: need pos < k and pos + width <= k (k=%u pos=%u width=%u)", k, pos, width);
    if (ctx && n && !out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL output");
    const uint64_t wmask = width >= 32 ? ~0ull : ((1ull << (2 * width)) - 1ull);  // intended MASK_TABLE[width] (SURVEY Q1 at 32)
    return shift_op<0>(ctx, k, in, nullptr, nullptr, 0, out, nullptr, nullptr, n, pos, wmask);
}
extern "C" int32_t kmb_append_base_words(kmb_ctx* ctx, uint32_t k, const uint64_t* in, const uint8_t* bases, int32_t bases_are_ascii,
                                         uint64_t* out, uint8_t* dropped_out, uint64_t n) {
    if (ctx && n && !bases) return fail(ctx, KMB_ERR_INVALID_ARG, "bases is NULL");
    return shift_op<1>(ctx, k, in, nullptr, bases, bases_are_ascii, out, nullptr, dropped_out, n, 0, 0);
}
extern "C" int32_t kmb_prepend_base_words(kmb_ctx* ctx, uint32_t k, const uint64_t* in, const uint8_t* bases, int32_t bases_are_ascii,
                                          uint64_t* out, uint8_t* dropped_out, uint64_t n) {
    if (ctx && n && !bases) return fail(ctx, KMB_ERR_INVALID_ARG, "bases is NULL");
    return shift_op<2>(ctx, k, in, nullptr, bases, bases_are_ascii, out, nullptr, dropped_out, n, 0, 0);
}
extern "C" int32_t kmb_canonical_append_base_words(kmb_ctx* ctx, uint32_t k, const uint64_t* fw_in, const uint64_t* rc_in, const uint8_t* bases,
                                                   int32_t bases_are_ascii, uint64_t* fw_out, uint64_t* rc_out, uint8_t* dropped_out, uint64_t n) {
    if (ctx && n && (!bases || !rc_in)) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    return shift_op<3>(ctx, k, fw_in, rc_in, bases, bases_are_ascii, fw_out, rc_out, dropped_out, n, 0, 0);
}
extern "C" int32_t kmb_canonical_prepend_base_words(kmb_ctx* ctx, uint32_t k, const uint64_t* fw_in, const uint64_t* rc_in, const uint8_t* bases,
                                                    int32_t bases_are_ascii, uint64_t* fw_out, uint64_t* rc_out, uint8_t* dropped_out, uint64_t n) {
    if (ctx && n && (!bases || !rc_in)) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    return shift_op<4>(ctx, k, fw_in, rc_in, bases, bases_are_ascii, fw_out, rc_out, dropped_out, n, 0, 0);
}
extern "C" int32_t kmb_is_fw_canonical_words(kmb_ctx* ctx, const uint64_t* fw, const uint64_t* rc, uint8_t* out, uint64_t n) {
    if (ctx && n && (!rc || !out)) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    return shift_op<5>(ctx, 1, fw, rc, nullptr, 0, nullptr, nullptr, out, n, 0, 0);
}

extern "C" int32_t kmb_kmer_get(kmb_ctx* ctx, uint32_t word_bits, uint32_t words_per_item, const void* arrays, uint64_t n_items, uint32_t index,
                                uint8_t* codes_out) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (!word_bits_ok(word_bits) || words_per_item < 1) return fail(ctx, KMB_ERR_INVALID_ARG, "word_bits must be 8/16/32/64/128 and words_per_item >= 1");
    const uint64_t item_bytes = (uint64_t)words_per_item * (word_bits / 8);
    // BitArray::get_bits asserts the range lies inside the array (bit_field 0.10)
    if (2ull * index + 1 >= item_bytes * 8) return fail(ctx, KMB_ERR_PANIC, "get(%u): beyond the %llu bits of the array", index, (unsigned long long)(item_bytes * 8));
    if (n_items == 0) return KMB_OK;
    if (!arrays || !codes_out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    const void* d_in;
    int32_t rc = in_prepare(ctx, 1, arrays, n_items * item_bytes, &d_in);
    if (rc) return rc;
    OutBuf ob;
    if ((rc = out_prepare(ctx, 0, codes_out, n_items, &ob))) return rc;
    const uint64_t ctas = (n_items + 255) / 256;
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    kmer_get_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint8_t*)d_in, n_items, (uint32_t)item_bytes, index, (uint8_t*)ob.dev);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, ob))) return rc;
    if (ob.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

extern "C" int32_t kmb_kmer_get_prefix(kmb_ctx* ctx, uint32_t word_bits, uint32_t words_per_item, const void* arrays, uint64_t n_items, uint32_t len,
                                       void* words_out) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (!word_bits_ok(word_bits) || words_per_item < 1) return fail(ctx, KMB_ERR_INVALID_ARG, "word_bits must be 8/16/32/64/128 and words_per_item >= 1");
    const uint64_t item_bytes = (uint64_t)words_per_item * (word_bits / 8);
    const uint64_t n_bits = 2ull * len + 1;  // 0..=(len*2): the reference's inclusive range (kmer.rs:51)
    // get_bits asserts the range fits one P and lies inside the array
    if (n_bits > word_bits || n_bits > item_bytes * 8)
        return fail(ctx, KMB_ERR_PANIC, "get_prefix(%u): %llu bits do not fit a u%u / the array", len, (unsigned long long)n_bits, word_bits);
    if (n_items == 0) return KMB_OK;
    if (!arrays || !words_out) return fail(ctx, KMB_ERR_INVALID_ARG, "NULL pointer");
    const void* d_in;
    int32_t rc = in_prepare(ctx, 1, arrays, n_items * item_bytes, &d_in);
    if (rc) return rc;
    const uint32_t word_bytes = word_bits / 8;
    OutBuf ob;
    if ((rc = out_prepare(ctx, 0, words_out, n_items * word_bytes, &ob))) return rc;
    const uint64_t ctas = (n_items * word_bytes + 255) / 256;
    if (ctas > 0x7FFFFFFFull) return fail(ctx, KMB_ERR_INVALID_ARG, "batch too large for one launch");
    kmer_get_prefix_kernel<<<(unsigned)ctas, 256, 0, ctx->stream>>>((const uint8_t*)d_in, n_items, (uint32_t)item_bytes, word_bytes, (uint32_t)n_bits,
                                                                   (uint8_t*)ob.dev);
    CK(ctx, cudaGetLastError());
    ctx->launches++;
    if ((rc = out_finish(ctx, ob))) return rc;
    if (ob.host) CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

// kmer.rs:71-91: always A0 C1 G2 T3, upper case, whatever encoder produced the word
extern "C" int32_t kmb_bitmer_to_bytes(kmb_ctx* ctx, uint32_t len, const uint64_t* mers, uint64_t n, uint8_t* bases_out) {
    NEED_CTX(ctx);
    if (len > 32) return fail(ctx, KMB_ERR_INVALID_ARG, "len = %u: a u64 holds 32 bases", len);
    return unpack_impl(ctx, KMB_ENC_ACGT, 64, mers, n, 1, len, bases_out, 0u);
}
