// kmb_hostpipe.cu -- the end-to-end form of the hot path for callers whose reads live in HOST memory:
// kmb_extract_canonical_host / kmb_extract_canonical_host_packed (include/kmers_b200.h).
//
// PCIe, not the kernel, bounds this path (a B200 extracts 386 G k-mers/s from resident reads; 1 B/base over a
// ~55 GB/s link feeds 44 G k-mers/s), so the pipeline is built around the bytes that cross the link:
//   * worker threads pack the caller's ASCII reads into the flat 2-bit + invalid-mask staging format
//     (kmb_hostpack.h: 3 bits/base instead of 8) straight into pinned ring buffers -- for pageable input this is also
//     the staging copy, there is no separate memcpy;
//   * while the packers are busy the DMA engine is not left idle: when the input is pinned and no packed chunk is
//     waiting, the orchestrating thread sends a chunk from the BACK of the batch as raw ASCII.  Packers eat the batch
//     from the front, the raw path from the back, they meet where their rates say they should -- no calibration;
//   * three streams: H2D, compute (the context's), D2H.  Copies in the two directions run concurrently (PCIe is
//     full duplex); a chunk's kernel waits only for its own H2D, its D2H only for its own kernel;
//   * outputs: caller's DEVICE arrays for the whole batch (the kernel writes them in place: results stay resident),
//     caller's HOST arrays (ring of device chunk buffers -> D2H stream), or none (digest only).
// The k-mer arithmetic itself -- windows, reverse complement, canonical min, LexHash -- runs only on the GPU
// (kmb_i_run_extract); the host threads do format conversion and copies.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <thread>

#include "kmb_hostpack.h"
#include "kmb_internal.h"

#define fail kmb_i_fail

namespace {

constexpr int kSlots = 8;  // packed staging slots (pinned host + device twin)
constexpr int kRaw = 3;    // device buffers of the raw-ASCII path
constexpr int kOut = 3;    // device chunk buffers per output array when the results go back to the host
constexpr size_t kPackJob = (size_t)1 << 20;  // bases per pack job (multiple of 64)

struct HostPipe {
    cudaStream_t h2d = nullptr, d2h = nullptr;
    kmbhost::Pool* pool = nullptr;
    uint8_t* h_pack[kSlots] = {};
    uint8_t* d_pack[kSlots] = {};
    size_t pack_cap = 0;
    cudaEvent_t ev_h2d[kSlots] = {}, ev_kern[kSlots] = {};
    uint8_t* d_raw[kRaw] = {};
    size_t raw_cap = 0;
    cudaEvent_t ev_raw_h2d[kRaw] = {}, ev_raw_kern[kRaw] = {};
    uint64_t* d_out[2][kOut] = {};
    size_t out_cap = 0;
    cudaEvent_t ev_out_kern[kOut] = {}, ev_out_d2h[kOut] = {};
    cudaEvent_t ev_entry = nullptr;
    // job bookkeeping shared with the workers
    std::mutex mu;
    std::condition_variable cv;
    std::atomic<uint64_t> jobs_outstanding{0};
    std::unique_ptr<std::atomic<int>[]> jobs_left;
    size_t jobs_left_cap = 0;
    // statistics of the last call (kmb_ctx_host_stats)
    uint64_t stat_chunks = 0, stat_raw_chunks = 0, stat_h2d_bytes = 0, stat_d2h_bytes = 0;
};

HostPipe* pipe_of(kmb_ctx* ctx) { return static_cast<HostPipe*>(ctx->hostpipe); }

int32_t pipe_create(kmb_ctx* ctx) {
    if (ctx->hostpipe) return KMB_OK;
    HostPipe* p = new (std::nothrow) HostPipe();
    if (!p) return fail(ctx, KMB_ERR_NOMEM, "out of host memory");
    ctx->hostpipe = p;
    CK(ctx, cudaStreamCreateWithFlags(&p->h2d, cudaStreamNonBlocking));
    CK(ctx, cudaStreamCreateWithFlags(&p->d2h, cudaStreamNonBlocking));
    auto mk = [&](cudaEvent_t* e) { return cudaEventCreateWithFlags(e, cudaEventDisableTiming); };
    for (int i = 0; i < kSlots; ++i) { CK(ctx, mk(&p->ev_h2d[i])); CK(ctx, mk(&p->ev_kern[i])); }
    for (int i = 0; i < kRaw; ++i) { CK(ctx, mk(&p->ev_raw_h2d[i])); CK(ctx, mk(&p->ev_raw_kern[i])); }
    for (int i = 0; i < kOut; ++i) { CK(ctx, mk(&p->ev_out_kern[i])); CK(ctx, mk(&p->ev_out_d2h[i])); }
    CK(ctx, mk(&p->ev_entry));
    return KMB_OK;
}

unsigned default_threads(const kmb_ctx* ctx) {
    if (ctx->host_threads) return ctx->host_threads;
    if (const char* e = getenv("KMB_HOST_THREADS")) { const int n = atoi(e); if (n > 0) return (unsigned)n; }
    const unsigned cpus = kmbhost::usable_cpus();  // one of them drives the pipeline (measured: 15 packers beat 16 on 16 cores)
    return std::min(cpus > 1 ? cpus - 1 : 1u, 32u);
}

int32_t ensure_pool(kmb_ctx* ctx, HostPipe* p) {
    const unsigned want = default_threads(ctx);
    if (p->pool && p->pool->size() == want) return KMB_OK;
    try {
        delete p->pool;
        p->pool = nullptr;
        p->pool = new kmbhost::Pool(want);
    } catch (...) {
        return fail(ctx, KMB_ERR_NOMEM, "could not start the host worker pool");
    }
    return KMB_OK;
}

// wait until every job handed to the pool has finished (they reference the caller's buffers)
void drain(HostPipe* p) {
    std::unique_lock<std::mutex> lk(p->mu);
    p->cv.wait(lk, [&] { return p->jobs_outstanding.load(std::memory_order_acquire) == 0; });
}

size_t env_mb(const char* name, size_t dflt) {
    if (const char* e = getenv(name)) { const long v = atol(e); if (v > 0) return (size_t)v; }
    return dflt;
}
bool env_flag(const char* name, bool dflt) {
    if (const char* e = getenv(name)) return atoi(e) != 0;
    return dflt;
}

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Plan {
    const uint8_t* ascii = nullptr;    // host ASCII reads (NULL for the pre-packed entry point)
    const uint32_t* pre_bits = nullptr;  // pre-packed host input
    const uint16_t* pre_inv = nullptr;
    uint64_t n_reads = 0, L = 0, W = 0, rpc = 0, n_chunks = 0;
    uint32_t k = 0, flags = 0;
    uint64_t* out[2] = {nullptr, nullptr};  // canon, hash as the caller gave them
    bool out_dev[2] = {false, false};
    bool want_digest = false;
};

// the pipeline proper; every exit path leaves no job running and no copy in flight
int32_t run_pipeline(kmb_ctx* ctx, HostPipe* p, const Plan& pl) {
    const uint64_t L = pl.L, W = pl.W, rpc = pl.rpc, n_chunks = pl.n_chunks;
    const bool prepacked = pl.ascii == nullptr;
    const bool pinned_in = kmb_i_is_pinned_ptr(prepacked ? (const void*)pl.pre_bits : (const void*)pl.ascii);
    // Packing pays when the link is the bottleneck and there are threads to feed it.  With a handful of threads per rank --
    // many ranks sharing one host -- the host's memory system is the bottleneck instead, and packing (read 1 B/base, write
    // and re-read 0.375) only adds traffic: measured on 8 ranks x 3 threads, 77 ms per step against 65 ms for raw copies
    // alone.  Pinned input then goes out as it is (pageable input still has to be staged by someone: the packers).
    const bool pack_default = default_threads(ctx) >= 6;
    const bool allow_pack = !prepacked && (env_flag("KMB_PIPE_PACK", pack_default) || !pinned_in);
    const bool allow_raw = !prepacked && pinned_in && (env_flag("KMB_PIPE_RAW", true) || !allow_pack);
    const bool host_out = (pl.out[0] && !pl.out_dev[0]) || (pl.out[1] && !pl.out_dev[1]);
    int32_t rc;

    // ---- buffers
    const uint64_t chunk_bases = rpc * L;
    const size_t chunk_words = (size_t)((chunk_bases + 15) / 16);
    const size_t inv_off_cap = round_up(chunk_words * 4, 256);
    const size_t pack_bytes = inv_off_cap + round_up(chunk_words * 2, 256);
    if (allow_pack || prepacked) {
        if (p->pack_cap < pack_bytes) {
            CK(ctx, cudaStreamSynchronize(p->h2d));
            CK(ctx, cudaStreamSynchronize(ctx->stream));
            for (int i = 0; i < kSlots; ++i) {
                if (p->h_pack[i]) { CK(ctx, cudaFreeHost(p->h_pack[i])); p->h_pack[i] = nullptr; }
                if (p->d_pack[i]) { CK(ctx, cudaFree(p->d_pack[i])); p->d_pack[i] = nullptr; }
            }
            p->pack_cap = 0;
            for (int i = 0; i < kSlots; ++i) {
                if (!prepacked) CK(ctx, cudaHostAlloc((void**)&p->h_pack[i], pack_bytes, cudaHostAllocDefault));
                CK(ctx, cudaMalloc((void**)&p->d_pack[i], pack_bytes + 256));
            }
            p->pack_cap = pack_bytes;
        } else if (!prepacked && !p->h_pack[0]) {
            for (int i = 0; i < kSlots; ++i) CK(ctx, cudaHostAlloc((void**)&p->h_pack[i], p->pack_cap, cudaHostAllocDefault));
        }
    }
    if (allow_raw && p->raw_cap < chunk_bases + 64) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        for (int i = 0; i < kRaw; ++i) {
            if (p->d_raw[i]) { CK(ctx, cudaFree(p->d_raw[i])); p->d_raw[i] = nullptr; }
        }
        p->raw_cap = 0;
        for (int i = 0; i < kRaw; ++i) CK(ctx, cudaMalloc((void**)&p->d_raw[i], chunk_bases + 64 + 256));
        p->raw_cap = chunk_bases + 64;
    }
    const size_t out_bytes_cap = (size_t)(rpc * W * 8);
    if (host_out && p->out_cap < out_bytes_cap) {
        CK(ctx, cudaStreamSynchronize(ctx->stream));
        CK(ctx, cudaStreamSynchronize(p->d2h));
        for (int a = 0; a < 2; ++a)
            for (int i = 0; i < kOut; ++i) {
                if (p->d_out[a][i]) { CK(ctx, cudaFree(p->d_out[a][i])); p->d_out[a][i] = nullptr; }
            }
        p->out_cap = 0;
        for (int a = 0; a < 2; ++a)
            for (int i = 0; i < kOut; ++i) CK(ctx, cudaMalloc((void**)&p->d_out[a][i], out_bytes_cap + 256));
        p->out_cap = out_bytes_cap;
    }
    if (allow_pack) {
        if ((rc = ensure_pool(ctx, p))) return rc;
        if (p->jobs_left_cap < n_chunks) {
            p->jobs_left.reset(new (std::nothrow) std::atomic<int>[n_chunks]);
            if (!p->jobs_left) { p->jobs_left_cap = 0; return fail(ctx, KMB_ERR_NOMEM, "out of host memory"); }
            p->jobs_left_cap = n_chunks;
        }
    }
    p->stat_chunks = n_chunks; p->stat_raw_chunks = 0; p->stat_h2d_bytes = 0; p->stat_d2h_bytes = 0;

    // the copy streams must not overtake work the caller queued on the context's stream (e.g. buffers being produced)
    CK(ctx, cudaEventRecord(p->ev_entry, ctx->stream));
    CK(ctx, cudaStreamWaitEvent(p->h2d, p->ev_entry, 0));
    CK(ctx, cudaStreamWaitEvent(p->d2h, p->ev_entry, 0));

    std::deque<cudaEvent_t> h2d_inflight;  // H2D copies enqueued and not yet seen complete, oldest first
    auto h2d_depth = [&]() {
        while (!h2d_inflight.empty() && cudaEventQuery(h2d_inflight.front()) == cudaSuccess) h2d_inflight.pop_front();
        return h2d_inflight.size();
    };
    uint64_t out_seq = 0, raw_seq = 0;

    // kernel + (for host outputs) D2H of one chunk whose input sits in device memory
    auto launch_chunk = [&](uint64_t c, const uint8_t* d_in, uint64_t in_bytes, uint32_t layout, const uint16_t* d_inv) -> int32_t {
        const uint64_t r0 = c * rpc, nr = std::min(rpc, pl.n_reads - r0), slot0 = r0 * W;
        uint64_t* dst[2];
        const int o = (int)(out_seq % kOut);
        for (int a = 0; a < 2; ++a) dst[a] = !pl.out[a] ? nullptr : (pl.out_dev[a] ? pl.out[a] + slot0 : p->d_out[a][o]);
        if (host_out && out_seq >= (uint64_t)kOut) CK(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_out_d2h[o], 0));
        int32_t r = kmb_i_run_extract(ctx, d_in, false, in_bytes, nr, L, pl.k, pl.flags, dst[0], dst[1], nullptr, nullptr, pl.want_digest,
                                      nullptr, 0, ctx->stream, L, layout, d_inv);
        if (r) return r;
        if (host_out) {
            CK(ctx, cudaEventRecord(p->ev_out_kern[o], ctx->stream));
            CK(ctx, cudaStreamWaitEvent(p->d2h, p->ev_out_kern[o], 0));
            for (int a = 0; a < 2; ++a)
                if (pl.out[a] && !pl.out_dev[a]) {
                    CK(ctx, cudaMemcpyAsync(pl.out[a] + slot0, p->d_out[a][o], nr * W * 8, cudaMemcpyDeviceToHost, p->d2h));
                    p->stat_d2h_bytes += nr * W * 8;
                }
            CK(ctx, cudaEventRecord(p->ev_out_d2h[o], p->d2h));
            ++out_seq;
        }
        return KMB_OK;
    };

    auto submit_packed = [&](uint64_t c) -> int32_t {
        const int s = (int)(c % kSlots);
        const uint64_t r0 = c * rpc, nr = std::min(rpc, pl.n_reads - r0), nb = nr * L;
        const size_t nw = (size_t)((nb + 15) / 16), inv_off = round_up(nw * 4, 256);
        if (c >= (uint64_t)kSlots) CK(ctx, cudaStreamWaitEvent(p->h2d, p->ev_kern[s], 0));  // the kernel that read d_pack[s] is done
        const uint16_t* d_inv = nullptr;
        if (prepacked) {
            const uint64_t w0 = r0 * L / 16;  // chunks start on word boundaries (rpc is a multiple of 16)
            CK(ctx, cudaMemcpyAsync(p->d_pack[s], pl.pre_bits + w0, nw * 4, cudaMemcpyHostToDevice, p->h2d));
            p->stat_h2d_bytes += nw * 4;
            if (pl.pre_inv) {
                CK(ctx, cudaMemcpyAsync(p->d_pack[s] + inv_off, pl.pre_inv + w0, nw * 2, cudaMemcpyHostToDevice, p->h2d));
                p->stat_h2d_bytes += nw * 2;
                d_inv = reinterpret_cast<const uint16_t*>(p->d_pack[s] + inv_off);
            }
        } else {
            CK(ctx, cudaMemcpyAsync(p->d_pack[s], p->h_pack[s], inv_off + nw * 2, cudaMemcpyHostToDevice, p->h2d));
            p->stat_h2d_bytes += inv_off + nw * 2;
            d_inv = reinterpret_cast<const uint16_t*>(p->d_pack[s] + inv_off);
        }
        CK(ctx, cudaEventRecord(p->ev_h2d[s], p->h2d));
        h2d_inflight.push_back(p->ev_h2d[s]);
        CK(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_h2d[s], 0));
        // a pre-packed store without masks holds no invalid base (SeqVector semantics): layout 2 with inv == NULL
        int32_t r = launch_chunk(c, p->d_pack[s], nw * 4, KMB_I_FLAT_PACKED, d_inv);
        if (r) return r;
        CK(ctx, cudaEventRecord(p->ev_kern[s], ctx->stream));
        return KMB_OK;
    };

    auto submit_raw = [&](uint64_t c) -> int32_t {
        const int s = (int)(raw_seq % kRaw);
        const uint64_t r0 = c * rpc, nr = std::min(rpc, pl.n_reads - r0), nb = nr * L;
        if (raw_seq >= (uint64_t)kRaw) CK(ctx, cudaStreamWaitEvent(p->h2d, p->ev_raw_kern[s], 0));
        CK(ctx, cudaMemcpyAsync(p->d_raw[s], pl.ascii + r0 * L, nb, cudaMemcpyHostToDevice, p->h2d));
        p->stat_h2d_bytes += nb;
        CK(ctx, cudaEventRecord(p->ev_raw_h2d[s], p->h2d));
        h2d_inflight.push_back(p->ev_raw_h2d[s]);
        CK(ctx, cudaStreamWaitEvent(ctx->stream, p->ev_raw_h2d[s], 0));
        int32_t r = launch_chunk(c, p->d_raw[s], nb, KMB_I_ASCII, nullptr);
        if (r) return r;
        CK(ctx, cudaEventRecord(p->ev_raw_kern[s], ctx->stream));
        ++raw_seq;
        ++p->stat_raw_chunks;
        return KMB_OK;
    };

    auto hand_to_packers = [&](uint64_t c) {
        const int s = (int)(c % kSlots);
        const uint64_t r0 = c * rpc, nr = std::min(rpc, pl.n_reads - r0), nb = nr * L;
        const size_t nw = (size_t)((nb + 15) / 16), inv_off = round_up(nw * 4, 256);
        const size_t n_jobs = (size_t)((nb + kPackJob - 1) / kPackJob);
        p->jobs_left[c].store((int)n_jobs, std::memory_order_release);
        const uint8_t* src = pl.ascii + r0 * L;
        uint32_t* bits = reinterpret_cast<uint32_t*>(p->h_pack[s]);
        uint16_t* inv = reinterpret_cast<uint16_t*>(p->h_pack[s] + inv_off);
        for (size_t j = 0; j < n_jobs; ++j) {
            const size_t b0 = j * kPackJob, n = (size_t)std::min<uint64_t>(kPackJob, nb - b0);
            p->jobs_outstanding.fetch_add(1, std::memory_order_acq_rel);
            std::atomic<int>* left = &p->jobs_left[c];
            p->pool->submit([p, src, b0, n, bits, inv, left] {
                kmbhost::pack_ascii(src + b0, n, bits + b0 / 16, inv + b0 / 16);
                left->fetch_sub(1, std::memory_order_acq_rel);
                p->jobs_outstanding.fetch_sub(1, std::memory_order_acq_rel);
                std::lock_guard<std::mutex> lk(p->mu);
                p->cv.notify_all();
            });
        }
    };

    uint64_t handed = 0, submitted = 0, back = n_chunks;
    rc = KMB_OK;
    try {
        if (prepacked || !allow_pack) {
            // no host work: the chunks go out in order (pre-packed words, or raw ASCII when packing is switched off)
            for (uint64_t c = 0; c < n_chunks && rc == KMB_OK; ++c) {
                while (h2d_depth() >= 4) std::this_thread::yield();  // bounded queue: keeps the event bookkeeping small
                rc = prepacked ? submit_packed(c) : submit_raw(c);
            }
        } else {
            while (submitted < back && rc == KMB_OK) {
                bool progress = false;
                // (1) hand chunks to the packers while staging slots are free
                while (handed < back && handed - submitted < (uint64_t)kSlots) {
                    const int s = (int)(handed % kSlots);
                    if (handed >= (uint64_t)kSlots && cudaEventQuery(p->ev_h2d[s]) != cudaSuccess) break;  // slot still being read by the DMA
                    hand_to_packers(handed++);
                    progress = true;
                }
                // (2) packed chunks that are complete go out, in order
                while (submitted < handed && rc == KMB_OK && p->jobs_left[submitted].load(std::memory_order_acquire) == 0) {
                    rc = submit_packed(submitted++);
                    progress = true;
                }
                // (3) the link would go idle: send a chunk from the back as it is
                if (rc == KMB_OK && allow_raw && back > handed && h2d_depth() < 2) {
                    rc = submit_raw(--back);
                    progress = true;
                }
                if (!progress) {
                    std::unique_lock<std::mutex> lk(p->mu);
                    p->cv.wait_for(lk, std::chrono::microseconds(50));
                }
            }
        }
    } catch (const std::bad_alloc&) {
        rc = fail(ctx, KMB_ERR_NOMEM, "out of host memory in the host pipeline");
    } catch (...) {
        rc = fail(ctx, KMB_ERR_CUDA, "unexpected failure in the host pipeline");
    }
    drain(p);
    cudaError_t e1 = cudaStreamSynchronize(p->h2d), e2 = cudaStreamSynchronize(ctx->stream), e3 = cudaStreamSynchronize(p->d2h);
    if (rc == KMB_OK && (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)) {
        const cudaError_t e = e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3);
        cudaGetLastError();
        rc = fail(ctx, KMB_ERR_CUDA, "host pipeline failed: %s", cudaGetErrorString(e));
    }
    return rc;
}

int32_t host_common(kmb_ctx* ctx, const uint8_t* ascii, const uint32_t* pre_bits, const uint16_t* pre_inv, uint64_t n_reads,
                    uint64_t fixed_len, uint32_t k, uint32_t flags, uint64_t* out_canon, uint64_t* out_hash, kmb_digest* digest) {
    if (k < 1 || k > 32) return fail(ctx, KMB_ERR_PANIC, "k = %u: kmers longer than 32 bases not supported (and k >= 1)", k);
    if (fixed_len == 0) return fail(ctx, KMB_ERR_INVALID_ARG, "fixed_len must be > 0");
    if (n_reads && !ascii && !pre_bits) return fail(ctx, KMB_ERR_INVALID_ARG, "the input pointer is NULL");
    if (n_reads && kmb_i_is_device_ptr(ascii ? (const void*)ascii : (const void*)pre_bits))
        return fail(ctx, KMB_ERR_INVALID_ARG, "the reads are in device memory: use kmb_batch_attach + kmb_extract_canonical");
    int32_t rc;
    if (digest && (rc = kmb_i_digest_begin(ctx))) return rc;
    const uint64_t W = fixed_len >= k ? fixed_len - k + 1 : 0;
    if (W && n_reads) {
        if ((rc = pipe_create(ctx))) return rc;
        Plan pl;
        pl.ascii = ascii; pl.pre_bits = pre_bits; pl.pre_inv = pre_inv;
        pl.n_reads = n_reads; pl.L = fixed_len; pl.W = W; pl.k = k; pl.flags = flags;
        pl.out[0] = out_canon; pl.out[1] = out_hash;
        for (int a = 0; a < 2; ++a) pl.out_dev[a] = pl.out[a] && kmb_i_is_device_ptr(pl.out[a]);
        pl.want_digest = digest != nullptr;
        // chunk: ~8 MiB of reads (3 bits/base of it cross the link when packed) and, when the results go back to the
        // host, no more output than the ring buffers should hold; a multiple of 16 reads keeps chunk starts on packed-word
        // boundaries and the output slots 32-byte aligned
        // Reads that arrive packed only cross the link: larger chunks get closer to the copy floor (10^7 x 150 bp, ms per call:
        // 4 MiB 13.7, 8 MiB 12.0, 16 MiB 11.1, 32 MiB 10.8, 64 MiB 10.6; floor 10.1) -- as long as there are enough chunks left
        // to overlap; ASCII reads also go through the packers, where 8 MiB is best (4 MiB 14.0, 8 MiB 13.3, 16 MiB 14.6, 32 MiB 16.5).
        uint64_t chunk_bases = env_mb("KMB_PIPE_CHUNK_MB", 0) << 20;
        if (chunk_bases == 0) {
            chunk_bases = 8ull << 20;
            if (!ascii) chunk_bases = std::min<uint64_t>(32ull << 20, std::max<uint64_t>(4ull << 20, n_reads * fixed_len / 16));
        }
        uint64_t rpc = chunk_bases / fixed_len;
        const bool host_out = (out_canon && !pl.out_dev[0]) || (out_hash && !pl.out_dev[1]);
        if (host_out) rpc = std::min<uint64_t>(rpc, (env_mb("KMB_PIPE_OUT_MB", 64) << 20) / (W * 8));
        rpc = std::max<uint64_t>(16, rpc / 16 * 16);
        pl.rpc = rpc;
        pl.n_chunks = (n_reads + rpc - 1) / rpc;
        if ((rc = run_pipeline(ctx, pipe_of(ctx), pl))) return rc;
    }
    if (digest) return kmb_i_digest_end(ctx, digest);
    CK(ctx, cudaStreamSynchronize(ctx->stream));
    return KMB_OK;
}

}  // namespace

void kmb_i_hostpipe_destroy(kmb_ctx* ctx) {
    HostPipe* p = pipe_of(ctx);
    if (!p) return;
    if (p->pool) { drain(p); delete p->pool; }
    if (p->h2d) { cudaStreamSynchronize(p->h2d); cudaStreamDestroy(p->h2d); }
    if (p->d2h) { cudaStreamSynchronize(p->d2h); cudaStreamDestroy(p->d2h); }
    for (int i = 0; i < kSlots; ++i) {
        cudaFreeHost(p->h_pack[i]);
        cudaFree(p->d_pack[i]);
        if (p->ev_h2d[i]) cudaEventDestroy(p->ev_h2d[i]);
        if (p->ev_kern[i]) cudaEventDestroy(p->ev_kern[i]);
    }
    for (int i = 0; i < kRaw; ++i) {
        cudaFree(p->d_raw[i]);
        if (p->ev_raw_h2d[i]) cudaEventDestroy(p->ev_raw_h2d[i]);
        if (p->ev_raw_kern[i]) cudaEventDestroy(p->ev_raw_kern[i]);
    }
    for (int i = 0; i < kOut; ++i) {
        cudaFree(p->d_out[0][i]);
        cudaFree(p->d_out[1][i]);
        if (p->ev_out_kern[i]) cudaEventDestroy(p->ev_out_kern[i]);
        if (p->ev_out_d2h[i]) cudaEventDestroy(p->ev_out_d2h[i]);
    }
    if (p->ev_entry) cudaEventDestroy(p->ev_entry);
    cudaGetLastError();
    delete p;
    ctx->hostpipe = nullptr;
}

extern "C" int32_t kmb_extract_canonical_host(kmb_ctx* ctx, const uint8_t* host_bases, uint64_t n_reads, uint64_t fixed_len,
                                              uint32_t k, uint32_t flags, uint64_t* out_canon, uint64_t* out_hash, kmb_digest* digest) {
    NEED_CTX(ctx);
    BIND(ctx);
    return host_common(ctx, host_bases, nullptr, nullptr, n_reads, fixed_len, k, flags, out_canon, out_hash, digest);
}

extern "C" int32_t kmb_extract_canonical_host_packed(kmb_ctx* ctx, const uint32_t* host_bits, const uint16_t* host_inv, uint64_t n_reads,
                                                     uint64_t fixed_len, uint32_t k, uint32_t flags, uint64_t* out_canon,
                                                     uint64_t* out_hash, kmb_digest* digest) {
    NEED_CTX(ctx);
    BIND(ctx);
    if (n_reads && !host_bits) return fail(ctx, KMB_ERR_INVALID_ARG, "host_bits is NULL");
    return host_common(ctx, nullptr, host_bits, host_inv, n_reads, fixed_len, k, flags, out_canon, out_hash, digest);
}

extern "C" int32_t kmb_ctx_set_host_threads(kmb_ctx* ctx, uint32_t n_threads) {
    NEED_CTX(ctx);
    ctx->host_threads = n_threads;
    return KMB_OK;
}

extern "C" int32_t kmb_ctx_host_stats(const kmb_ctx* ctx, uint64_t* stats4) {
    if (!ctx || !stats4) return kmb_i_fail(nullptr, KMB_ERR_INVALID_ARG, "NULL argument");
    const HostPipe* p = static_cast<const HostPipe*>(ctx->hostpipe);
    stats4[0] = p ? p->stat_chunks : 0;
    stats4[1] = p ? p->stat_raw_chunks : 0;
    stats4[2] = p ? p->stat_h2d_bytes : 0;
    stats4[3] = p ? p->stat_d2h_bytes : 0;
    return KMB_OK;
}

extern "C" int32_t kmb_host_pack(const uint8_t* bases, uint64_t n_bases, uint32_t* bits_out, uint16_t* inv_out) {
    if (n_bases && (!bases || !bits_out || !inv_out)) return kmb_i_fail(nullptr, KMB_ERR_INVALID_ARG, "NULL pointer");
    kmbhost::pack_ascii(bases, (size_t)n_bases, bits_out, inv_out);
    return KMB_OK;
}

extern "C" const char* kmb_host_pack_isa(void) { return kmbhost::pack_isa(); }

extern "C" int32_t kmb_host_read_probe(const uint8_t* buf, uint64_t n_bytes, uint32_t n_threads, double* seconds_out) {
    if (!buf || !seconds_out || n_threads < 1) return kmb_i_fail(nullptr, KMB_ERR_INVALID_ARG, "NULL pointer / no threads");
    try {
        kmbhost::Pool pool(n_threads);
        std::mutex mu;
        std::condition_variable cv;
        const uint64_t job = (uint64_t)1 << 20;
        const uint64_t n_jobs = (n_bytes + job - 1) / job;
        std::atomic<uint64_t> left{n_jobs}, sink{0};
        const auto t0 = std::chrono::steady_clock::now();
        for (uint64_t j = 0; j < n_jobs; ++j)
            pool.submit([&, j] {
                sink.fetch_xor(kmbhost::read_all(buf + j * job, (size_t)std::min(job, n_bytes - j * job)), std::memory_order_relaxed);
                if (left.fetch_sub(1, std::memory_order_acq_rel) == 1) { std::lock_guard<std::mutex> lk(mu); cv.notify_all(); }
            });
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return left.load(std::memory_order_acquire) == 0; });
        }
        *seconds_out = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    } catch (...) {
        return kmb_i_fail(nullptr, KMB_ERR_NOMEM, "could not run the read probe");
    }
    return KMB_OK;
}
