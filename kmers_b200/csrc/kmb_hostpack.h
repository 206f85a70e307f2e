// kmb_hostpack.h -- host side of the pinned staging path: ASCII reads -> 2 bits/base + 1 invalid bit/base, and the
// worker pool that runs it next to the DMA engine.  Plain C++ (g++), no CUDA: the .cu translation units only see this
// interface.  Internal: nothing here is exported except through include/kmers_b200.h (kmb_host_pack).
//
// Staged format ("flat packed", the tile layout of kmb_device.cuh written by the host instead of stage_tile):
//   bits[i] : 16 bases, base j at bits 2j+1:2j, code A0 C1 G2 T3 = the SeqVector / Kmer::from layout
//             (naive_impl/seq_vector.rs:230-242, naive_impl/kmer.rs:234-251); any other byte maps by (c >> 1) & 3 first
//             (encoding/naive.rs:14-16), exactly like the device packer, so KMB_F_NO_VALIDATE keeps Path-E semantics
//   inv[i]  : bit j = base j is not one of ACGTacgt (naive_impl/mod.rs:40-50)
// 3 bits per base cross PCIe instead of 8.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>

namespace kmbhost {

// Pack n_bases ASCII bytes.  bits / inv receive ceil(n_bases / 16) entries; in the last, partial entry the missing
// bases read as code 0 / invalid.  Dispatches once to AVX-512BW+GFNI, AVX-512BW+BMI2, AVX2+BMI2 or a portable loop.
void pack_ascii(const uint8_t* src, size_t n_bases, uint32_t* bits, uint16_t* inv);
// which implementation pack_ascii resolved to: "avx512gfni", "avx512bw", "avx2" or "swar"
const char* pack_isa();
// force an implementation (tests): 0 = best available, 1 = swar, 2 = avx2, 3 = avx512bw, 4 = avx512gfni (if supported)
void pack_force_isa(int which);

// Fixed-size pool of worker threads with a FIFO of jobs; the pipelined host path hands it pack / copy jobs.
class Pool {
public:
    explicit Pool(unsigned n_threads);
    ~Pool();
    Pool(const Pool&) = delete;
    Pool& operator=(const Pool&) = delete;
    unsigned size() const;
    void submit(std::function<void()> job);  // may throw std::bad_alloc
private:
    struct Impl;
    Impl* impl_;
};

// CPUs this process may run on (sched_getaffinity), at least 1
unsigned usable_cpus();

// XOR of every 8-byte word of [p, p + n): the cheapest possible consumer of a buffer, for measuring how fast this host can
// stream its memory (the floor of any path that must read the caller's ASCII reads once)
uint64_t read_all(const uint8_t* p, size_t n);

}  // namespace kmbhost
