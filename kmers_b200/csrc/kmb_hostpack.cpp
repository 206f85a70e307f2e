// kmb_hostpack.cpp -- see kmb_hostpack.h.  Compiled by g++ (no CUDA); the SIMD variants carry their own target
// attributes and are selected at run time, so the object itself needs no -m flags.
#include "kmb_hostpack.h"

#include <immintrin.h>
#include <sched.h>

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace kmbhost {
namespace {

// ---------------------------------------------------------------- portable path
// lut[c] = 2-bit code | invalid << 2.  code = x ^ (x >> 1) with x = (c >> 1) & 3: A0 C1 G2 T3 for the letters
// (naive_impl/mod.rs:40-50), the Path-E field permuted the same way for every other byte (encoding/naive.rs:14-16).
struct Lut {
    uint8_t v[256];
    Lut() {
        for (int c = 0; c < 256; ++c) {
            const unsigned x = ((unsigned)c >> 1) & 3u;
            const unsigned u = (unsigned)c & 0xDFu;
            const bool ok = u == 'A' || u == 'C' || u == 'G' || u == 'T';
            v[c] = (uint8_t)((x ^ (x >> 1)) | (ok ? 0u : 4u));
        }
    }
};
const Lut g_lut;

// one entry (<= 16 bases); bases that do not exist read as code 0 / invalid
inline void pack_entry(const uint8_t* src, size_t n, uint32_t* bits, uint16_t* inv) {
    uint32_t b = 0, m = n < 16 ? (0xFFFFu << n) & 0xFFFFu : 0u;
    for (size_t j = 0; j < n; ++j) {
        const unsigned t = g_lut.v[src[j]];
        b |= (t & 3u) << (2 * j);
        m |= (t >> 2) << j;
    }
    *bits = b;
    *inv = (uint16_t)m;
}

void pack_swar(const uint8_t* src, size_t n_bases, uint32_t* bits, uint16_t* inv) {
    const size_t full = n_bases / 16;
    for (size_t e = 0; e < full; ++e) pack_entry(src + 16 * e, 16, bits + e, inv + e);
    if (n_bases % 16) pack_entry(src + 16 * full, n_bases % 16, bits + full, inv + full);
}

// ---------------------------------------------------------------- AVX2 + BMI2: 32 bases per iteration
// bit 1 / bit 2 of every byte through movemask (a 16-bit lane shift brings them to bit 7 of their own byte), the two
// bit planes interleaved by pdep; validity = the byte, case-folded, equals the letter its low nibble claims (pshufb LUT).
__attribute__((target("avx2,bmi2"), always_inline)) inline void pack32_avx2(const uint8_t* src, uint32_t* bits, uint16_t* inv, const __m256i table,
                                                                           const __m256i m0f, const __m256i mdf) {
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src));
    const uint32_t m1 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 6));
    const uint32_t m2 = (uint32_t)_mm256_movemask_epi8(_mm256_slli_epi16(v, 5));
    const uint64_t w = _pdep_u64(m1 ^ m2, 0x5555555555555555ull) | _pdep_u64(m2, 0xAAAAAAAAAAAAAAAAull);
    const __m256i expect = _mm256_shuffle_epi8(table, _mm256_and_si256(v, m0f));
    const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, _mm256_and_si256(v, mdf)));
    std::memcpy(bits, &w, 8);
    std::memcpy(inv, &bad, 4);
}

__attribute__((target("avx2,bmi2"))) void pack_avx2(const uint8_t* src, size_t n_bases, uint32_t* bits, uint16_t* inv) {
    const __m256i table = _mm256_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1,
                                           -1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1);
    const __m256i m0f = _mm256_set1_epi8(0x0F), mdf = _mm256_set1_epi8((char)0xDF);
    constexpr size_t kStreams = 4;  // see pack_avx512
    size_t done = 0;
    if (n_bases >= 64 * 1024) {
        const size_t per = n_bases / kStreams / 32;
        for (size_t b = 0; b < per; ++b)
            for (size_t s = 0; s < kStreams; ++s) {
                const size_t blk = s * per + b;
                pack32_avx2(src + 32 * blk, bits + 2 * blk, inv + 2 * blk, table, m0f, mdf);
            }
        done = kStreams * per;
    }
    const size_t blocks = n_bases / 32;
    for (size_t b = done; b < blocks; ++b) pack32_avx2(src + 32 * b, bits + 2 * b, inv + 2 * b, table, m0f, mdf);
    if (n_bases % 32) pack_swar(src + 32 * blocks, n_bases % 32, bits + 2 * blocks, inv + 2 * blocks);
}

// ---------------------------------------------------------------- AVX-512BW + BMI2: 64 bases per iteration
__attribute__((target("avx512f,avx512bw,bmi2"), always_inline)) inline void pack64_avx512(const uint8_t* src, uint32_t* bits, uint16_t* inv,
                                                                                         const __m512i table, const __m512i b1, const __m512i b2,
                                                                                         const __m512i m0f, const __m512i mdf) {
    const __m512i v = _mm512_loadu_si512(src);
    const uint64_t m1 = _mm512_test_epi8_mask(v, b1), m2 = _mm512_test_epi8_mask(v, b2);
    const uint64_t lo = m1 ^ m2;
    const uint64_t w0 = _pdep_u64(lo, 0x5555555555555555ull) | _pdep_u64(m2, 0xAAAAAAAAAAAAAAAAull);
    const uint64_t w1 = _pdep_u64(lo >> 32, 0x5555555555555555ull) | _pdep_u64(m2 >> 32, 0xAAAAAAAAAAAAAAAAull);
    const __m512i expect = _mm512_shuffle_epi8(table, _mm512_and_si512(v, m0f));
    const uint64_t bad = ~(uint64_t)_mm512_cmpeq_epi8_mask(expect, _mm512_and_si512(v, mdf));
    std::memcpy(bits, &w0, 8);
    std::memcpy(bits + 2, &w1, 8);
    std::memcpy(inv, &bad, 8);
}

// A core streams memory faster from several places at once than from one (each stream gets its own hardware prefetcher
// and the misses overlap; measured 8.7 -> 14.4 GB/s per core from one to four streams), and this loop is bound by exactly
// that, so large inputs are walked as kStreams interleaved quarters.
__attribute__((target("avx512f,avx512bw,bmi2"))) void pack_avx512(const uint8_t* src, size_t n_bases, uint32_t* bits, uint16_t* inv) {
    const __m512i table = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1));
    const __m512i b1 = _mm512_set1_epi8(0x02), b2 = _mm512_set1_epi8(0x04);
    const __m512i m0f = _mm512_set1_epi8(0x0F), mdf = _mm512_set1_epi8((char)0xDF);
    constexpr size_t kStreams = 4;
    size_t done = 0;
    if (n_bases >= 64 * 1024) {
        const size_t per = n_bases / kStreams / 64;  // 64-base blocks per stream
        for (size_t b = 0; b < per; ++b)
            for (size_t s = 0; s < kStreams; ++s) {
                const size_t blk = s * per + b;
                pack64_avx512(src + 64 * blk, bits + 4 * blk, inv + 4 * blk, table, b1, b2, m0f, mdf);
            }
        done = kStreams * per;
    }
    const size_t blocks = n_bases / 64;
    for (size_t b = done; b < blocks; ++b) pack64_avx512(src + 64 * b, bits + 4 * b, inv + 4 * b, table, b1, b2, m0f, mdf);
    if (n_bases % 64) pack_swar(src + 64 * blocks, n_bases % 64, bits + 4 * blocks, inv + 4 * blocks);
}

// ---------------------------------------------------------------- AVX-512BW + GFNI: 64 bases per iteration, no mask round trip
// The 2-bit code is a GF(2)-linear function of the byte (low = bit1 ^ bit2, high = bit2): one GF2P8AFFINEQB yields it in
// every byte; two multiply-adds (x1,x4 then x1,x16) fold four codes into one byte, VPMOVDB compacts the 16 bytes.
__attribute__((target("avx512f,avx512bw,gfni"), always_inline)) inline void pack64_gfni(const uint8_t* src, uint32_t* bits, uint16_t* inv,
                                                                                      const __m512i table, const __m512i aff, const __m512i mul4,
                                                                                      const __m512i mul16, const __m512i m0f, const __m512i mdf) {
    const __m512i v = _mm512_loadu_si512(src);
    const __m512i code = _mm512_gf2p8affine_epi64_epi8(v, aff, 0);
    const __m512i nib = _mm512_maddubs_epi16(code, mul4);   // c0 + 4 c1 per 16-bit lane
    const __m512i byt = _mm512_madd_epi16(nib, mul16);      // + 16 (c2 + 4 c3) per 32-bit lane
    const __m128i packed = _mm512_cvtepi32_epi8(byt);
    const __m512i expect = _mm512_shuffle_epi8(table, _mm512_and_si512(v, m0f));
    const uint64_t bad = ~(uint64_t)_mm512_cmpeq_epi8_mask(expect, _mm512_and_si512(v, mdf));
    _mm_storeu_si128(reinterpret_cast<__m128i*>(bits), packed);
    std::memcpy(inv, &bad, 8);
}

__attribute__((target("avx512f,avx512bw,avx512vl,gfni"))) void pack_gfni(const uint8_t* src, size_t n_bases, uint32_t* bits, uint16_t* inv) {
    const __m512i table = _mm512_broadcast_i32x4(_mm_setr_epi8(-1, 0x41, -1, 0x43, 0x54, -1, -1, 0x47, -1, -1, -1, -1, -1, -1, -1, -1));
    // output bit i of every byte = parity(row[7 - i] & byte): bit 0 <- bits 1,2 (0x06); bit 1 <- bit 2 (0x04)
    const __m512i aff = _mm512_set1_epi64((long long)((0x06ull << 56) | (0x04ull << 48)));
    const __m512i mul4 = _mm512_set1_epi16(0x0401), mul16 = _mm512_set1_epi32(0x00100001);
    const __m512i m0f = _mm512_set1_epi8(0x0F), mdf = _mm512_set1_epi8((char)0xDF);
    constexpr size_t kStreams = 4;  // see pack_avx512
    size_t done = 0;  // in 64-base blocks
    // (Non-temporal stores for the staging rings were tried and lost: the rings stay cache-resident between the packer and
    // the DMA engine -- 16 threads packed 80 GB/s with streaming stores against 100 GB/s with ordinary ones.)
    if (n_bases >= 64 * 1024) {
        const size_t per = n_bases / kStreams / 64;
        for (size_t b = 0; b < per; ++b)
            for (size_t s = 0; s < kStreams; ++s) {
                const size_t blk = s * per + b;
                pack64_gfni(src + 64 * blk, bits + 4 * blk, inv + 4 * blk, table, aff, mul4, mul16, m0f, mdf);
            }
        done = kStreams * per;
    }
    const size_t blocks = n_bases / 64;
    for (size_t b = done; b < blocks; ++b) pack64_gfni(src + 64 * b, bits + 4 * b, inv + 4 * b, table, aff, mul4, mul16, m0f, mdf);
    if (n_bases % 64) pack_swar(src + 64 * blocks, n_bases % 64, bits + 4 * blocks, inv + 4 * blocks);
}

using PackFn = void (*)(const uint8_t*, size_t, uint32_t*, uint16_t*);
struct Choice {
    PackFn fn;
    const char* name;
};

Choice choose(int which) {
    __builtin_cpu_init();
    const bool bmi2 = __builtin_cpu_supports("bmi2");
    const bool avx2 = bmi2 && __builtin_cpu_supports("avx2");
    const bool avx512 = bmi2 && __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    const bool gfni = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                      __builtin_cpu_supports("gfni");
    if ((which == 0 || which == 4) && gfni) return {pack_gfni, "avx512gfni"};
    if ((which == 0 || which == 3 || which == 4) && avx512) return {pack_avx512, "avx512bw"};
    if ((which == 0 || which == 2 || which == 3 || which == 4) && avx2) return {pack_avx2, "avx2"};
    return {pack_swar, "swar"};
}
std::atomic<int> g_forced{0};
int env_isa() {  // KMB_HOST_PACK_ISA=swar|avx2|avx512bw caps the implementation (tests run every variant)
    const char* e = getenv("KMB_HOST_PACK_ISA");
    if (!e) return 0;
    return !strcmp(e, "swar") ? 1 : (!strcmp(e, "avx2") ? 2 : (!strcmp(e, "avx512bw") ? 3 : (!strcmp(e, "avx512gfni") ? 4 : 0)));
}
Choice current() {
    static const Choice best = choose(env_isa());
    const int f = g_forced.load(std::memory_order_relaxed);
    return f == 0 ? best : choose(f);
}

}  // namespace

void pack_ascii(const uint8_t* src, size_t n_bases, uint32_t* bits, uint16_t* inv) {
    if (n_bases) current().fn(src, n_bases, bits, inv);
}
const char* pack_isa() { return current().name; }
void pack_force_isa(int which) { g_forced.store(which < 0 || which > 4 ? 0 : which, std::memory_order_relaxed); }

namespace {
__attribute__((target("avx2"))) uint64_t read_all_avx2(const uint8_t* p, size_t n) {
    __m256i a0 = _mm256_setzero_si256(), a1 = a0, a2 = a0, a3 = a0;
    size_t i = 0;
    for (; i + 128 <= n; i += 128) {
        a0 = _mm256_xor_si256(a0, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + i)));
        a1 = _mm256_xor_si256(a1, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + i + 32)));
        a2 = _mm256_xor_si256(a2, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + i + 64)));
        a3 = _mm256_xor_si256(a3, _mm256_loadu_si256(reinterpret_cast<const __m256i*>(p + i + 96)));
    }
    a0 = _mm256_xor_si256(_mm256_xor_si256(a0, a1), _mm256_xor_si256(a2, a3));
    uint64_t w[4];
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(w), a0);
    uint64_t r = w[0] ^ w[1] ^ w[2] ^ w[3];
    for (; i < n; ++i) r ^= p[i];
    return r;
}
}  // namespace

uint64_t read_all(const uint8_t* p, size_t n) {
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx2")) return read_all_avx2(p, n);
    uint64_t r = 0;
    size_t i = 0;
    for (; i + 8 <= n; i += 8) { uint64_t v; std::memcpy(&v, p + i, 8); r ^= v; }
    for (; i < n; ++i) r ^= p[i];
    return r;
}

unsigned usable_cpus() {
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0) {
        const int n = CPU_COUNT(&set);
        if (n > 0) return (unsigned)n;
    }
    const unsigned h = std::thread::hardware_concurrency();
    return h ? h : 1u;
}

// ---------------------------------------------------------------- worker pool
struct Pool::Impl {
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> q;
    std::vector<std::thread> threads;
    bool stop = false;
    void loop() {
        for (;;) {
            std::function<void()> job;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return stop || !q.empty(); });
                if (q.empty()) return;  // stop requested and nothing left
                job = std::move(q.front());
                q.pop_front();
            }
            job();
        }
    }
};

Pool::Pool(unsigned n_threads) : impl_(new Impl) {
    if (n_threads < 1) n_threads = 1;
    impl_->threads.reserve(n_threads);
    try {
        for (unsigned i = 0; i < n_threads; ++i) impl_->threads.emplace_back([this] { impl_->loop(); });
    } catch (...) {
        // fewer threads than asked for is fine as long as there is one; with none, jobs run inline in submit()
    }
}

Pool::~Pool() {
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->stop = true;
    }
    impl_->cv.notify_all();
    for (auto& t : impl_->threads) t.join();
    delete impl_;
}

unsigned Pool::size() const { return (unsigned)impl_->threads.size(); }

void Pool::submit(std::function<void()> job) {
    if (impl_->threads.empty()) {  // thread creation failed altogether
        job();
        return;
    }
    {
        std::lock_guard<std::mutex> lk(impl_->mu);
        impl_->q.push_back(std::move(job));
    }
    impl_->cv.notify_one();
}

}  // namespace kmbhost
