#!/usr/bin/env python
"""bench.py -- canonical k-mers/s (K=31, 150 bp reads) on N B200s, next to the CPU path.

One "step" = one pass of the hot path (pack -> windows -> reverse complement ->
canonical min -> LexHash, materialised) over one resident batch of synthetic
reads: BASELINE.json configs[1], 10^7 x 150 bp per GPU (weak scaling: reads shard
by batch, no data-path collective).  Prints ONE JSON line (rank 0) that also carries

  e2e            the same metric through the C-ABI host call (host reads in, H2D inside the timed region); the product
                 of every variant is named: device arrays | host arrays | digest only | pre-packed input
  configs        every other BASELINE.json config and "next" row (K=63, long reads with the full invalid-base mixture,
                 ragged reads, minimizers, packed store, compacted output) timed under this same clock, each with its
                 algorithmic bytes, roofline fraction and a parity check against the oracle
  config5_strong configs[4]: ONE 1 Gbp sequence cut over the N ranks (K-1 halo), fused histogram + digest, and the
                 [bins | digest] ncclAllReduce INSIDE the timed region (strong scaling; the one collective of the path)
  parity_multi_gpu  (N > 1) the sharded histogram on real ranks against the oracle, and kmb_allreduce_u64 on 2 contexts

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

--impl reference times the reference's own CPU algorithm (the C restatement in
oracle/ -- the Rust crate cannot be built in this image) on all host threads.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 31
READ_LEN = 150
READS_PER_GPU = 10_000_000
SEED = 42
METRIC = "canonical_kmers_per_sec_k31_150bp"
UNIT = "kmers/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md, used only when MEASURED_PEAKS.json is absent
G5, SEED5, THRESH5, HIST_BITS = 1_000_000_000, 44, 105, 16  # config 5: 1 Gbp, N-rate 0.01 %, 2^16 bins


def algorithmic_bytes(n_reads: int, read_len: int, k: int) -> int:
    """L bytes read + (L-K+1) x (8 B canonical + 8 B hash) written, per read."""
    return n_reads * (read_len + (read_len - k + 1) * 16)


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def mem_available_gb() -> float:
    try:
        with open("/proc/meminfo") as f:
            for line in f:
                if line.startswith("MemAvailable:"):
                    return int(line.split()[1]) / 1e6
    except Exception:
        pass
    return 0.0


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        super().__init__(daemon=True)
        self.gpu_index = gpu_index
        self.rows = []
        self._stop = threading.Event()
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self._stop.is_set():
                    break
        except Exception:
            pass

    def stop(self):
        self._stop.set()
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(names, r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle port; the only places bench.py executes oracle/ besides the untimed parity checks)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, sample_reads: int):
    """The reference's CPU path (oracle port) on all host threads, on a bounded sample of the workload."""
    import oracle as ko
    import numpy as np
    cores = len(os.sched_getaffinity(0)) or 1
    bases = ko.generate_bases(SEED, 0, sample_reads * READ_LEN)
    n_kmers = sample_reads * (READ_LEN - K + 1)
    canon = np.zeros(n_kmers, dtype=np.uint64)  # allocated and touched once, like the GPU's resident outputs
    hsh = np.zeros(n_kmers, dtype=np.uint64)

    def one_pass():
        t0 = time.perf_counter()
        r = ko.extract_canonical(bases, K, n_reads=sample_reads, fixed_len=READ_LEN, n_threads=cores,
                                 canon_out=canon, hash_out=hsh)
        return time.perf_counter() - t0, r

    for _ in range(warmup):
        one_pass()
    times = []
    for _ in range(steps):
        dt, r = one_pass()
        times.append(dt)
    assert r["n_valid"] == n_kmers
    ms = 1e3 * sum(times) / len(times)
    return {"value": n_kmers / (ms / 1e3), "ms_per_step": ms, "cores": cores, "sample_reads": sample_reads,
            "n_kmers": n_kmers}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle as ko
    build = ko.use_native_build()
    # the GPU arm's own configuration (10 M reads per step) when the host has the memory for the 19.2 GB of result arrays
    # the CPU path materialises; a 2 M-read prefix of the same data otherwise
    need_gb = READS_PER_GPU * ((READ_LEN - K + 1) * 16 + READ_LEN) / 1e9
    sample = READS_PER_GPU if mem_available_gb() > 2.0 * need_gb else 2_000_000
    sample = min(sample, args.reads)
    r = cpu_reference(args.steps, args.warmup, sample)
    sample_desc = (f"{'all' if sample == READS_PER_GPU else 'first'} {sample} of the {READS_PER_GPU} synthetic 150 bp reads (seed {SEED}), "
                   f"CanonicalKmerIterator + get_canonical_word + LexHasher, canon+hash materialised to host arrays, {r['cores']} threads, {build}")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[1]: K=31 canonical k-mer extraction + LexHash over 10M synthetic 150bp reads per GPU",
                   "k": K, "read_len": READ_LEN, "reads_per_gpu": READS_PER_GPU, "reads_per_step": sample},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": sample_desc},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "C restatement of COMBINE-lab/kmers CPU path (oracle/); the Rust crate cannot be built in this image",
    }
    print(json.dumps(line), flush=True)


ORIG_AFFINITY = None  # the affinity this process started with (restored for the CPU baseline leg)


def bind_to_gpu_numa_node(local_rank: int):
    """Pin this rank to the CPU cores next to its GPU (NVML's CPU affinity) BEFORE any pinned host memory is allocated, so
    that the staging buffers land on the GPU's own NUMA node: with 8 ranks pulling 50+ GB/s each over PCIe, buffers on the
    far socket would all squeeze through the inter-socket link.  Returns a short description for the JSON line."""
    global ORIG_AFFINITY
    try:
        ORIG_AFFINITY = os.sched_getaffinity(0)
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        n_words = (os.cpu_count() + 63) // 64
        words = pynvml.nvmlDeviceGetCpuAffinity(h, n_words)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus near gpu {local_rank}"
    except Exception as ex:  # no NVML / restricted container: keep the inherited affinity
        return f"inherited ({type(ex).__name__})"
    return "inherited"


# ----------------------------------------------------------------------------------------------------------------------
# helpers of the GPU legs
# ----------------------------------------------------------------------------------------------------------------------
def event_times(torch, stream, fn, reps, warmup):
    """avg / best CUDA-event time of fn() on `stream` (the stream the kernels are launched on)."""
    for _ in range(warmup):
        fn()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    torch.cuda.synchronize()
    for a, b in evs:
        a.record(stream)
        fn()
        b.record(stream)
    torch.cuda.synchronize()
    ts = [a.elapsed_time(b) for a, b in evs]
    return sum(ts) / len(ts), min(ts)


def c4_mixture(torch, np, dev_bases, n_reads: int, L: int, seed: int = 43):
    """SURVEY 8(d) config-4 invalid-base mixture applied IN PLACE to uniform reads (which already carry the 0.1 % isolated
    N): per read one run of N of 1..200 bases, 5 % of the reads soft-masked to lower case over a 500-base span (stays
    valid), and on every 37th read one IUPAC letter and one newline.  Counter-based, so any prefix can be rebuilt."""
    import oracle as ko
    sm = np.array([ko.lib().ko_splitmix64(seed * 1_000_003 + r) for r in range(min(n_reads, 4096))], dtype=np.uint64)
    if n_reads > sm.size:  # cheap vectorised continuation of the same recipe (splitmix64 in numpy u64 arithmetic)
        r = np.arange(n_reads, dtype=np.uint64) + np.uint64(seed * 1_000_003)
        with np.errstate(over="ignore"):
            z = r + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
        assert np.array_equal(z[:sm.size], sm), "numpy splitmix64 != oracle splitmix64"
        sm = z
    run_len = 1 + (sm % np.uint64(200)).astype(np.int64)
    run_off = ((sm >> np.uint64(8)) % np.uint64(L - 200)).astype(np.int64)
    rows = torch.arange(n_reads, device="cuda", dtype=torch.int64) * L
    j = torch.arange(200, device="cuda", dtype=torch.int64)[None, :]
    rl, ro = torch.from_numpy(run_len).cuda()[:, None], torch.from_numpy(run_off).cuda()[:, None]
    idx = (rows[:, None] + ro + j)[j < rl]
    dev_bases[idx] = ord("N")
    soft = torch.from_numpy(((sm >> np.uint64(20)) % np.uint64(20) == 0)).cuda()
    so = torch.from_numpy(((sm >> np.uint64(28)) % np.uint64(L - 500)).astype(np.int64)).cuda()
    j5 = torch.arange(500, device="cuda", dtype=torch.int64)[None, :]
    idx = (rows[soft][:, None] + so[soft][:, None] + j5).reshape(-1)
    dev_bases[idx] |= 0x20
    iu = torch.from_numpy(((np.arange(n_reads) % 37) == 0)).cuda()
    io = torch.from_numpy(((sm >> np.uint64(40)) % np.uint64(L - 1)).astype(np.int64)).cuda()
    letters = torch.tensor(list(b"RYKM"), dtype=torch.uint8, device="cuda")
    dev_bases[rows[iu] + io[iu]] = letters[torch.from_numpy((sm % np.uint64(4)).astype(np.int64)).cuda()[iu]]
    dev_bases[rows[iu] + (io[iu] + 977) % L] = ord("\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=READS_PER_GPU, help="reads per GPU (default: configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config list and config 5")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import kmers_b200 as kb
    from kmers_b200 import _native as nv
    from kmers_b200.context import _ptr
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    host_affinity = bind_to_gpu_numa_node(local_rank) if world > 1 else "inherited (single rank)"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")  # keep NCCL's version banner off stdout: rank 0 prints exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n_reads, L = args.reads, READ_LEN
    W = L - K + 1
    n_slots = n_reads * W
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)  # torch work, the kernels and the timing events all share this stream
    ctx = kb.Context(local_rank, stream=stream.cuda_stream)
    cores = len(os.sched_getaffinity(0)) or 1
    host_threads = max(1, cores // world - 1)  # the ranks of one box share its cores; one of a rank's cores drives the pipeline
    ctx.set_host_threads(host_threads)
    peak, peak_src = hbm_peak()
    i64 = lambda n: torch.empty(n, dtype=torch.int64, device="cuda")
    ko = None
    if rank == 0:
        import oracle as ko  # the checker (untimed parity) and, below, the cpu_baseline leg
        cpu_build = ko.use_native_build()

    # rank r owns reads [r*n_reads, (r+1)*n_reads) of one synthetic data set
    batch = ctx.generate(SEED, n_reads, L, n_thresh20=0, first_index=rank * n_reads * L)
    out = kb.CanonicalKmers(k=K, n_slots=n_slots)
    out.canon = i64(n_slots)
    out.hash = i64(n_slots)

    # correctness of this very configuration: digest + a prefix against the oracle (outside the timed region)
    res = batch.extract_canonical(K, digest=True, out=out)
    digest = res.digest
    assert digest[0] == n_slots, digest
    csum = int(out.canon.sum().item()) % 2**64
    hsum = int(out.hash.sum().item()) % 2**64
    assert (csum, hsum) == (digest[1] % 2**64, digest[2] % 2**64), "digest != sum of the materialised arrays"
    parity = "digest==sum(arrays)"
    npre = min(n_reads, 100_000)
    if rank == 0:
        ref = ko.extract_canonical(ko.generate_bases(SEED, 0, npre * L), K, n_reads=npre, fixed_len=L, n_threads=cores)
        assert np.array_equal(out.canon[:npre * W].cpu().numpy().view(np.uint64), ref["canon"]), "canonical words != oracle"
        assert np.array_equal(out.hash[:npre * W].cpu().numpy().view(np.uint64), ref["hash"]), "hashes != oracle"
        parity += f"; first {npre} reads bit-exact vs oracle"

    # ---- device-resident timing (the clock sampler starts first: no idle gap between the warm-up and the timed steps)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    for _ in range(args.warmup):
        batch.extract_canonical(K, out=out)
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches0 = ctx.launch_count
    t_all0, t_all1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    for a, b in evs:
        a.record(stream)
        batch.extract_canonical(K, out=out)
        b.record(stream)
    t_all1.record(stream)
    barrier()
    launches = ctx.launch_count - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    kernel_ms = [a.elapsed_time(b) for a, b in evs]
    time.sleep(0.2)
    sampler.stop()
    total_ms = max_over_ranks(total_ms)
    ms_per_step = total_ms / args.steps
    value = world * n_slots / (ms_per_step / 1e3)

    # ---- roofline of the dominant (only) kernel
    avg_kernel_ms = sum(kernel_ms) / len(kernel_ms)
    alg = algorithmic_bytes(n_reads, L, K)
    achieved = alg / (avg_kernel_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "kmb::fixed_kernel<NarrowEng<validate=1,digest=0,fwrc=0,mode=materialise,K>16,hash=1>>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_kernel_ms, "best_launch_ms": min(kernel_ms),
                "launch_ms_all": [round(x, 4) for x in kernel_ms],
                "frac_of_nominal_8TBs": achieved / 8000.0}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                t = json.load(f)
            roofline["traffic"] = t.get("dram_bytes_per_launch_at_bench_size")
            roofline["traffic_source"] = t.get("source")
            roofline["traffic_kernel_source_sha1"] = t.get("kernel_source_sha1")  # stale if != the current sources' hash
            roofline["kernel_source_sha1_now"] = kernel_source_sha1()
        except Exception:
            pass

    # ---- e2e: host buffers through the C-ABI one-shot call (packing + H2D inside the timed region)
    e2e = None
    if not args.no_e2e:
        e2e = e2e_legs(args, torch, np, dist, kb, ctx, batch, out, digest, n_reads, L, W, world, rank, host_affinity, host_threads, barrier,
                       max_over_ranks, ko)

    # ---- the other configs under this clock (one GPU: rank 0 of a single-rank run)
    configs = None
    if not args.no_configs and world == 1:
        out = res = None
        torch.cuda.empty_cache()
        configs = config_list(torch, np, kb, nv, C, _ptr, ctx, stream, peak, ko, cores)
    # ---- config 5, strong-scaled, all-reduce inside the timed region
    config5 = None
    if not args.no_configs:
        config5 = config5_strong(args, torch, np, dist, kb, local_rank, stream, rank, world, barrier, max_over_ranks, peak, ko, cores)
    parity_multi = None
    if world > 1:
        parity_multi = multi_gpu_parity(torch, np, dist, kb, ctx, rank, world, local_rank, ko)

    # ---- final reduction across ranks of the config-1 digest (checksum / count), NCCL, outside the timed region
    global_digest = list(digest)
    if world > 1:
        from kmers_b200.dist import allreduce_histogram
        _, gd = allreduce_histogram(torch.zeros(1, dtype=torch.int64, device="cuda"), digest)  # wrapping u64 sums over ranks
        global_digest = list(gd)

    cpu = None
    if rank == 0 and not args.no_cpu:
        if ORIG_AFFINITY:
            os.sched_setaffinity(0, ORIG_AFFINITY)  # the CPU path gets every core the job was given, not just the GPU's node
        r = cpu_reference(10, 1, 2_000_000)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "build": cpu_build,
               "sample": f"first {r['sample_reads']} reads of the same workload, iterator + canonical + LexHash materialised, "
                         f"{r['cores']} threads, mean of 10 passes (~{10 * r['ms_per_step'] * r['cores'] / 1e3:.0f} core-seconds)"}
        try:  # SURVEY 8(d) variant 1: what benches/simple_benchmark.rs does per window (O(K) re-encode + rev-comp), canonicalised
            nf = 500_000
            fb = ko.generate_bases(SEED, 0, nf * READ_LEN)
            t0 = time.perf_counter()
            ko.bench_windows(fb, K, n_reads=nf, fixed_len=READ_LEN, n_threads=r["cores"], materialize=False)
            cpu["bench_faithful"] = {"value": nf * (READ_LEN - K + 1) / (time.perf_counter() - t0), "unit": UNIT, "cores": r["cores"],
                                     "sample": f"first {nf} reads, per-window Kmer::from + to_reverse_complement + canonical-min, folded"}
        except Exception as ex:
            cpu["bench_faithful"] = {"unavailable": str(ex)[:200]}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64",
            "data": "synthetic",
            "config": {"workload": "configs[1]: K=31 canonical k-mer extraction + LexHash over 10M synthetic 150bp reads per GPU",
                       "k": K, "read_len": READ_LEN, "reads_per_gpu": n_reads, "kmers_per_gpu": n_slots,
                       "bases_per_sec": world * n_reads * L / (ms_per_step / 1e3),
                       "l2_policy": "inputs+outputs (20.7 GB) larger than L2, no flush needed", "parity": parity,
                       "parallelism": f"reads sharded by batch over {world} GPU(s), no data-path collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
            "clocks": sampler.summary(), "digest": global_digest, "configs": configs, "config5_strong": config5,
            "parity_multi_gpu": parity_multi,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def kernel_source_sha1() -> str:
    """Hash of the CUDA sources the dominant kernel is built from: profiles/traffic.json is stamped with it, so a stale
    capture (kernel changed since) is visible in the bench line."""
    import hashlib
    h = hashlib.sha1()
    for f in ("kmb_device.cuh", "kmb_geometry.cuh", "kmb_extract.cuh", "kmb_tu_narrow.cu"):
        with open(os.path.join(ROOT, "kmers_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


# ----------------------------------------------------------------------------------------------------------------------
def e2e_legs(args, torch, np, dist, kb, ctx, batch, out, digest, n_reads, L, W, world, rank, host_affinity, host_threads, barrier,
             max_over_ranks, ko):
    """Everything a caller with HOST reads sees, each variant with its product named."""
    n_slots = n_reads * W
    host = torch.empty(n_reads * L, dtype=torch.uint8, pin_memory=True)
    host_np = host.numpy()
    host_np[:] = batch.download()
    steps = max(3, min(args.steps, 10))

    def timed(fn, reps, warm=2):
        for _ in range(warm):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            r = fn()
        barrier()
        return max_over_ranks((time.perf_counter() - t0) / reps), r

    # transfer floors of this box, measured the same way: one plain cudaMemcpyAsync of the same bytes, nothing else
    dev_tmp = torch.empty(n_reads * L, dtype=torch.uint8, device="cuda")
    h2d_raw_s, _ = timed(lambda: dev_tmp.copy_(host, non_blocking=True), 3, 1)
    del dev_tmp
    # ... and how fast this host can stream the reads out of its own memory at all (every path has to read them once)
    # (all ranks at once, like everything else here: the ranks of one box share its memory system)
    host_read_s, _ = timed(lambda: kb.host_read_probe(host_np, host_threads), 3, 1)

    # (1) headline: pinned host ASCII reads -> canonical + hash arrays for the whole batch, resident on the device, + digest
    out.canon.fill_(0)
    out.hash.fill_(0)
    dt, d = timed(lambda: ctx.extract_canonical_host(host_np, n_reads, L, K, out_canon=out.canon, out_hash=out.hash), steps)
    st = ctx.host_stats()
    assert d == digest, "e2e digest differs from the resident run"
    assert (int(out.canon.sum().item()) % 2**64, int(out.hash.sum().item()) % 2**64) == (digest[1], digest[2]), "e2e device arrays != digest"
    floor_s = st["h2d_bytes"] / (n_reads * L) * h2d_raw_s  # the bytes this call moved, at the raw copy's rate
    e2e = {"value": world * n_slots / dt, "unit": UNIT, "ms_per_step": 1e3 * dt, "steps": steps,
           "product": "device arrays: canonical words + LexHashes of the WHOLE batch (2 x 9.6 GB per GPU) written in place into the caller's "
                      "device arrays and left resident; 24 B digest read back",
           "h2d_bytes_per_step": st["h2d_bytes"], "d2h_bytes_per_step": 24, "input": "pinned host ASCII, 1 B/base",
           "host_threads": host_threads, "host_pack_isa": kb.host_pack_isa(), "chunks": st["chunks"], "raw_ascii_chunks": st["raw_chunks"],
           "host_affinity": host_affinity, "h2d_raw_ascii_copy_only_ms": 1e3 * h2d_raw_s,
           "speedup_over_raw_ascii_copy": h2d_raw_s / dt,
           "host_read_floor_ms": 1e3 * host_read_s, "frac_of_host_read_floor": host_read_s / dt,
           "host_read_floor_what": f"{host_threads} threads per rank reading the rank's 1.5 GB of reads once, all {world} rank(s) at the same time "
                                   f"(kmb_host_read_probe): {world * n_reads * L / host_read_s / 1e9:.0f} GB/s aggregate",
           "aggregate_host_GBps": {"ascii_consumed_by_this_call": world * n_reads * L / dt / 1e9,
                                   "raw_h2d_copy": world * n_reads * L / h2d_raw_s / 1e9,
                                   "host_read_probe": world * n_reads * L / host_read_s / 1e9,
                                   "note": "every path has to stream the ASCII reads out of host DRAM once (1 B/base), by the packer threads or by "
                                           "the DMA engine; when the ranks of one box together ask for more than its memory system delivers, that -- not "
                                           "PCIe, not the GPU -- is the ceiling of the ASCII-in figure (prepacked_input needs 0.375 B/base)"},
           "what": "kmb_extract_canonical_host: worker threads pack the reads to 2 bit + 1 validity bit per base into pinned rings "
                   "(front of the batch) while the DMA engine also takes raw ASCII chunks from the back; H2D, kernel and D2H on three streams"}
    # (2) nothing materialised: digest only
    dt, d = timed(lambda: ctx.extract_canonical_host(host_np, n_reads, L, K), steps)
    assert d == digest
    e2e["digest_only"] = {"value": world * n_slots / dt, "unit": UNIT, "ms_per_step": 1e3 * dt, "product": "24 B digest; no array is written",
                          "h2d_bytes_per_step": ctx.host_stats()["h2d_bytes"], "d2h_bytes_per_step": 24}
    # (3) reads the caller keeps 2-bit packed on the host (packed once, outside the timed region): no host work per step
    try:
        nw = (n_reads * L + 15) // 16
        hb = torch.empty(nw, dtype=torch.int32, pin_memory=True)
        hi = torch.empty(nw, dtype=torch.int16, pin_memory=True)
        hb_np, hi_np = hb.numpy().view(np.uint32), hi.numpy().view(np.uint16)
        kb.host_pack(host_np, hb_np, hi_np)
        dt, d = timed(lambda: ctx.extract_canonical_host_packed(hb_np, hi_np, n_reads, L, K, out_canon=out.canon, out_hash=out.hash), steps)
        assert d == digest
        e2e["prepacked_input"] = {"value": world * n_slots / dt, "unit": UNIT, "ms_per_step": 1e3 * dt,
                                  "product": "device arrays + digest", "input": "pinned host 2 bit + 1 validity bit per base (kmb_host_pack, done once)",
                                  "h2d_bytes_per_step": ctx.host_stats()["h2d_bytes"], "d2h_bytes_per_step": 24,
                                  "frac_of_transfer_floor": (ctx.host_stats()["h2d_bytes"] / (n_reads * L) * h2d_raw_s) / dt}
        del hb, hi, hb_np, hi_np
    except Exception as ex:
        e2e["prepacked_input"] = {"unavailable": str(ex)[:200]}
    e2e["frac_of_transfer_floor"] = floor_s / (e2e["ms_per_step"] / 1e3)  # bytes actually moved, at the raw copy's rate, over the call's time
    # (4) like for like with the CPU arm: both result arrays back in HOST memory (16 B per k-mer over PCIe: D2H-bound)
    try:
        budget = 0.35 * mem_available_gb() * 1e9 / world
        ns = n_reads if n_reads * W * 16 <= budget else max(16, int(budget // (W * 16)) // 16 * 16)
        hc = torch.empty(ns * W, dtype=torch.int64, pin_memory=True)
        hh = torch.empty(ns * W, dtype=torch.int64, pin_memory=True)
        hc_np, hh_np = hc.numpy().view(np.uint64), hh.numpy().view(np.uint64)
        # D2H floor: the same bytes through plain copies of device-resident arrays
        d2h_s, _ = timed(lambda: (hc.copy_(out.canon[:ns * W], non_blocking=True), hh.copy_(out.hash[:ns * W], non_blocking=True)), 2, 1)
        dtm, d = timed(lambda: ctx.extract_canonical_host(host_np[:ns * L], ns, L, K, out_canon=hc_np, out_hash=hh_np), 3, 1)
        stm = ctx.host_stats()
        if ns == n_reads:
            assert d == digest
        assert (int(hc_np.sum(dtype=np.uint64)), int(hh_np.sum(dtype=np.uint64))) == (d[1], d[2]), "host arrays != digest"
        e2e["host_arrays"] = {"value": world * ns * W / dtm, "unit": UNIT, "reads_per_step": ns, "ms_per_step": 1e3 * dtm,
                              "product": "host arrays: canonical words + LexHashes in pinned HOST memory (what the CPU arm produces)",
                              "h2d_bytes_per_step": stm["h2d_bytes"], "d2h_bytes_per_step": stm["d2h_bytes"],
                              "d2h_copy_only_ms": 1e3 * d2h_s, "frac_of_d2h_floor": d2h_s / dtm,
                              "what": "same call with pinned host canon+hash arrays: D2H of 16 B per k-mer is the bound (PCIe)"}
        del hc, hh, hc_np, hh_np
    except Exception as ex:  # pinned allocation can fail on a small host
        e2e["host_arrays"] = {"unavailable": str(ex)[:200]}
    del host, host_np
    return e2e


# ----------------------------------------------------------------------------------------------------------------------
def config_list(torch, np, kb, nv, C, _ptr, ctx, stream, peak, ko, cores):
    """BASELINE.json configs[2..3] and the "next" rows on one GPU: CUDA-event time of the launch(es), algorithmic bytes
    (SURVEY 8d), fraction of the measured copy peak, and a parity check of a prefix against the oracle."""
    rows = []
    i64 = lambda n: torch.empty(n, dtype=torch.int64, device="cuda")
    i32 = lambda n: torch.empty(n, dtype=torch.int32, device="cuda")
    u64 = lambda t: t.cpu().numpy().view(np.uint64)

    def add(name, units, unit, alg, ms, best, parity, **extra):
        rows.append(dict({"name": name, "units": units, "unit": unit + "/s", "value": units / (ms / 1e3), "ms": ms, "best_ms": best,
                          "algorithmic_bytes": alg, "achieved_GBps": alg / (ms / 1e3) / 1e9, "frac": alg / (ms / 1e3) / 1e9 / peak,
                          "bound": "hbm", "parity": parity}, **extra))

    def guarded(name, fn):
        try:
            fn()
        except Exception as ex:  # one config must not take the headline down with it
            rows.append({"name": name, "error": f"{type(ex).__name__}: {str(ex)[:300]}"})
        torch.cuda.empty_cache()

    n, L = READS_PER_GPU, READ_LEN
    reads = ctx.generate(SEED, n, L)
    npre = 2000
    pre = ko.generate_bases(SEED, 0, npre * L)

    # ---- config 3: K=63, two words per k-mer (extension above K=32: parity pinned for encode + rev_comp only)
    def config3():
        k, w = 63, L - 63 + 1
        canon, hsh = i64(2 * n * w), i64(2 * n * w)
        ref = ko.extract_canonical_wide(pre, k, n_reads=npre, fixed_len=L, n_threads=cores)
        for want_hash in (False, True):
            run = lambda: ctx._ck(ctx._lib.kmb_extract_canonical_wide(ctx._h, k, kb.ENC_ACGT, 0, _ptr(canon), _ptr(hsh) if want_hash else None, None))
            ms, best = event_times(torch, stream, run, 5, 2)
            ok = np.array_equal(u64(canon[:2 * npre * w]).reshape(-1, 2), ref["canon"])
            if want_hash:
                ok = ok and np.array_equal(u64(hsh[:2 * npre * w]).reshape(-1, 2), ref["hash"])
            assert ok, "config 3 differs from the oracle"
            t0 = time.perf_counter()
            ko.extract_canonical_wide(ko.generate_bases(SEED, 0, 40_000 * L), k, n_reads=40_000, fixed_len=L, n_threads=cores, want_hash=want_hash)
            cpu = 40_000 * w / (time.perf_counter() - t0)
            add(f"config3 K=63 2xu64 canon{'+hash' if want_hash else ' only'} over the 10M 150bp reads", n * w, "kmers",
                n * (L + w * (32 if want_hash else 16)), ms, best, f"first {npre} reads bit-exact vs oracle (extension above K=32)",
                cpu_port={"value": cpu, "unit": UNIT, "cores": cores, "sample": "40000 reads, Encoding::encode + swap-loop rev_comp per window"})
    guarded("config3", config3)

    # ---- "next" rows on the config-2 reads: minimizers (N1), packed store (N2)
    def minimizers():
        w = L - K + 1
        mm, mp = i64(n * w), i32(n * w)
        run = lambda: ctx._ck(ctx._lib.kmb_minimizers(ctx._h, 31, 15, 15, 0, _ptr(mm), _ptr(mp)))
        ms, best = event_times(torch, stream, run, 5, 2)
        rm, rp = ko.minimizers_batch(pre, 31, 15, 15, n_reads=npre, fixed_len=L)
        assert np.array_equal(u64(mm[:npre * w]), rm) and np.array_equal(mp[:npre * w].cpu().numpy().view(np.uint32), rp), "minimizers differ from the oracle"
        add("N1 minimizers k=31 w=15 (lmer word + pos per k-mer) over the 10M reads", n * w, "kmers", n * (L + w * 12), ms, best,
            f"first {npre} reads bit-exact vs the restated SeqVecMinimizerIter deque")
    guarded("minimizers", minimizers)

    def packed():
        w = L - K + 1
        canon, hsh = i64(n * w), i64(n * w)
        o = kb.CanonicalKmers(k=K, n_slots=n * w, canon=canon, hash=hsh)
        pb = reads.to_packed()
        ms, best = event_times(torch, stream, lambda: pb.extract_canonical(K, out=o), 5, 2)
        ref = ko.extract_canonical(pre, K, n_reads=npre, fixed_len=L, n_threads=cores)
        assert np.array_equal(u64(canon[:npre * w]), ref["canon"]) and np.array_equal(u64(hsh[:npre * w]), ref["hash"]), "packed-store extraction differs"
        add("N2 extraction from the 2-bit packed store (SeqVector twin), K=31 canon+hash", n * w, "kmers", n * (40 + w * 16), ms, best,
            f"first {npre} reads bit-exact vs oracle")
    guarded("packed", packed)

    # ---- N3 compacted, iterator-identical output on reads with 0.1 % N
    def compact():
        b = ctx.generate(SEED, n, L, n_thresh20=1049)
        cnt = C.c_uint64()
        ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, None, None, None, None, 0, C.byref(cnt)))
        m = int(cnt.value)
        cc, ch, cp, ce = i64(m), i64(m), i32(m), i64(n + 1)
        run = lambda: ctx._ck(ctx._lib.kmb_extract_compact(ctx._h, K, 0, _ptr(cc), _ptr(ch), _ptr(cp), _ptr(ce), m, C.byref(cnt)))
        ms, best = event_times(torch, stream, run, 5, 2)
        ref = ko.extract_canonical(ko.generate_bases(SEED, 0, npre * L, 1049), K, n_reads=npre, fixed_len=L, n_threads=cores)
        keep = ref["canon"] != np.uint64(kb.SENTINEL)
        e_end = int(ce[npre].item())
        assert e_end == int(keep.sum()) and np.array_equal(u64(cc[:e_end]), ref["canon"][keep]) and np.array_equal(u64(ch[:e_end]), ref["hash"][keep])
        pos = np.tile(np.arange(L - K + 1, dtype=np.int32), npre)[keep]
        assert np.array_equal(cp[:e_end].cpu().numpy(), pos), "compacted positions differ"
        add("N3 compacted (pos,canon,hash) iterator-identical stream, 10M reads + 0.1% N", m, "kmers", n * L + m * 20 + (n + 1) * 8, ms, best,
            f"first {npre} reads: exactly the restated CanonicalKmerIterator sequence")
    guarded("compact", compact)

    # ---- ragged reads of 100..150 bp
    def ragged():
        lens = torch.from_numpy(np.random.default_rng(1).integers(100, 151, size=n)).cuda()
        offs = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
        offs[1:] = torch.cumsum(lens, 0)
        nb, ns = int(offs[-1].item()), int((lens - (K - 1)).sum().item())
        flat = torch.empty(nb, dtype=torch.uint8, device="cuda")
        g = ctx.generate(SEED, 1, nb)
        ctx._ck(ctx._lib.kmb_batch_download(ctx._h, _ptr(flat), nb))
        rb = ctx.attach(flat, dev_offsets=offs)
        canon, hsh = i64(ns), i64(ns)
        o = kb.CanonicalKmers(k=K, n_slots=ns, canon=canon, hash=hsh)
        ms, best = event_times(torch, stream, lambda: rb.extract_canonical(K, out=o), 5, 2)
        ho = offs[:npre + 1].cpu().numpy().astype(np.uint64)
        ref = ko.extract_canonical(ko.generate_bases(SEED, 0, int(ho[-1])), K, offsets=ho, n_threads=cores)
        m = ref["canon"].size
        assert np.array_equal(u64(canon[:m]), ref["canon"]) and np.array_equal(u64(hsh[:m]), ref["hash"]), "ragged extraction differs"
        add("ragged reads of 100..150 bp (CSR offsets), K=31 canon+hash, 10M reads", ns, "kmers", nb + ns * 16 + n * 16, ms, best,
            f"first {npre} reads bit-exact vs oracle")
    guarded("ragged", ragged)

    # ---- config 4: 10 kbp reads with the full invalid-base mixture (window resets)
    def config4():
        n4, L4 = 100_000, 10_000
        w4 = L4 - K + 1
        flat = torch.empty(n4 * L4, dtype=torch.uint8, device="cuda")
        ctx.generate(43, n4, L4, n_thresh20=1049)
        ctx._ck(ctx._lib.kmb_batch_download(ctx._h, _ptr(flat), n4 * L4))
        c4_mixture(torch, np, flat, n4, L4)
        b4 = ctx.attach(flat, fixed_len=L4)
        canon, hsh = i64(n4 * w4), i64(n4 * w4)
        o = kb.CanonicalKmers(k=K, n_slots=n4 * w4, canon=canon, hash=hsh)
        ms, best = event_times(torch, stream, lambda: b4.extract_canonical(K, out=o), 5, 2)
        d = b4.extract_canonical(K, out=o, digest=True).digest
        p4 = 300
        hostpre = flat[:p4 * L4].cpu().numpy()
        ref = ko.extract_canonical(hostpre, K, n_reads=p4, fixed_len=L4, n_threads=cores)
        assert np.array_equal(u64(canon[:p4 * w4]), ref["canon"]) and np.array_equal(u64(hsh[:p4 * w4]), ref["hash"]), "config 4 differs from the oracle"
        n_lower = int(((hostpre >= 97) & (hostpre <= 122)).sum())
        n_bad = int((~np.isin(hostpre & 0xDF, [65, 67, 71, 84])).sum())
        t0 = time.perf_counter()
        s4 = 20_000
        rd = ko.extract_canonical(flat[:s4 * L4].cpu().numpy(), K, n_reads=s4, fixed_len=L4, n_threads=cores, materialize=False)
        cpu = s4 * w4 / (time.perf_counter() - t0)
        add("config4 K=31, 100k x 10 kbp reads, N runs + 0.1% isolated N + soft-masked spans + IUPAC + newline", n4 * w4, "slots",
            n4 * (L4 + w4 * 16), ms, best,
            f"first {p4} reads bit-exact vs oracle ({n_bad} invalid bytes, {n_lower} lower-case bytes in that prefix); {d[0]} of {n4 * w4} windows valid",
            cpu_port={"value": cpu, "unit": "slots/s", "cores": cores, "sample": f"{s4} reads, iterator, digest only"})
    guarded("config4", config4)
    return rows


# ----------------------------------------------------------------------------------------------------------------------
def config5_strong(args, torch, np, dist, kb, local_rank, stream, rank, world, barrier, max_over_ranks, peak, ko, cores):
    """configs[4]: one 1 Gbp sequence, cut into `world` contiguous ranges with a K-1 halo (a window belongs to the range
    holding its first base), fused histogram + digest per rank, then ONE in-place ncclAllReduce of the
    [65536 bins | n_valid | checksum_canon | checksum_hash] buffer -- kernel AND collective inside the timed region.
    The fold is the reference bench's wrapping `.sum()` (benches/simple_benchmark.rs:14-22) widened to a histogram."""
    from kmers_b200.dist import shard_sequence, sharded_histogram_fused
    steps = max(5, min(args.steps, 20))
    ctx5 = kb.Context(local_rank, stream=stream.cuda_stream)
    try:
        start, stop, load_stop = shard_sequence(G5, K, rank, world)
        mine = ctx5.generate(SEED5, 1, load_stop - start, n_thresh20=THRESH5, first_index=start)
        nb = (1 << HIST_BITS) + 3
        buf = torch.zeros(nb, dtype=torch.int64, device="cuda")
        for _ in range(3):
            sharded_histogram_fused(ctx5, mine, K, HIST_BITS, buf)
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(steps)]
        barrier()
        l0 = ctx5.launch_count
        for e0, e1, e2 in ev:
            e0.record(stream)
            mine.histogram(K, HIST_BITS, hist=buf, accumulate=False, digest_in_hist=True)
            e1.record(stream)
            if world > 1:
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            e2.record(stream)
        barrier()
        launches = ctx5.launch_count - l0
        k_ms = [a.elapsed_time(b) for a, b, _ in ev]
        ar_us = [1e3 * b.elapsed_time(c) for _, b, c in ev]
        step_ms = max_over_ranks(sum(a.elapsed_time(c) for a, _, c in ev) / steps)
        kern_ms = max_over_ranks(sum(k_ms) / steps)
        n_win = G5 - K + 1
        g = buf.cpu().numpy().view(np.uint64)
        gdig = [int(x) for x in g[-3:]]
        out = {"workload": "configs[4]: 1 Gbp synthetic sequence (seed 44, 0.01 % N), K=31, LexHash-prefix histogram 2^16 bins + digest, "
                           f"cut over {world} GPU(s) with a K-1 halo, [bins|digest] all-reduced in place (ncclAllReduce, u64 sum)",
               "scaling": "strong", "n_gpus": world, "steps": steps, "ms_per_step": step_ms, "value": n_win / (step_ms / 1e3), "unit": UNIT,
               "kernel_ms": kern_ms, "kernel_ms_rank0_best": min(k_ms), "allreduce_us_rank0_avg": sum(ar_us) / steps, "allreduce_us_rank0_best": min(ar_us),
               "allreduce_bytes": nb * 8, "launches_per_step": launches / steps, "global_digest": gdig,
               "hbm": {"algorithmic_bytes_per_gpu": (load_stop - start) + 8 * (1 << HIST_BITS),
                       "frac": ((load_stop - start) + 8 * (1 << HIST_BITS)) / (kern_ms / 1e3) / 1e9 / peak,
                       "note": "issue/atomic-bound by construction: 1 B/base in, nothing materialised"}}
        # the same sequence on ONE GPU in this same run (rank 0): the denominator of the strong-scaling efficiency, and parity
        if rank == 0:
            whole = ctx5.generate(SEED5, 1, G5, n_thresh20=THRESH5, first_index=0)
            b1 = torch.zeros(nb, dtype=torch.int64, device="cuda")
            t1, t1_best = event_times(torch, stream, lambda: whole.histogram(K, HIST_BITS, hist=b1, accumulate=False, digest_in_hist=True), steps, 2)
            same = bool(np.array_equal(b1.cpu().numpy().view(np.uint64), g))
            out["one_gpu_ms"] = t1
            out["one_gpu_value"] = n_win / (t1 / 1e3)
            out["efficiency_vs_one_gpu"] = t1 / (world * step_ms)
            out["kernel_efficiency_vs_one_gpu"] = t1 / (world * kern_ms)
            # issue-slot roofline of the fused kernel: thread instructions per window (from the ncu capture of this build,
            # profiles/hist_issue.json) x windows / (SMs x 4 schedulers x 32 lanes x clock)
            try:
                with open(os.path.join(ROOT, "profiles", "hist_issue.json")) as f:
                    hi = json.load(f)
                props = torch.cuda.get_device_properties(local_rank)
                issue_peak = props.multi_processor_count * 4 * 32 * hi["sm_clock_hz"]  # thread-instructions per second
                out["roofline"] = {"bound": "issue", "unit": "thread-instructions/s", "peak": issue_peak,
                                   "achieved": hi["thread_inst_per_window"] * n_win / (t1 / 1e3),
                                   "frac": hi["thread_inst_per_window"] * n_win / (t1 / 1e3) / issue_peak,
                                   "thread_inst_per_window": hi["thread_inst_per_window"], "source": hi.get("source")}
            except Exception:
                out["roofline"] = None
            # oracle on a prefix of the sequence (4 Mbp): bins + digest bit-exact
            npref = 4_000_000
            small = ctx5.generate(SEED5, 1, npref, n_thresh20=THRESH5, first_index=0)
            hb, dg = small.histogram(K, HIST_BITS, to="host")
            ref = ko.extract_canonical(ko.generate_bases(SEED5, 0, npref, THRESH5), K, n_reads=1, fixed_len=npref, hist_bits=HIST_BITS, materialize=False)
            ok = bool(np.array_equal(hb, ref["hist"])) and dg == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
            out["parity"] = (("sharded+all-reduced == one GPU (bins and digest); " if same else "SHARDED RESULT DIFFERS FROM ONE GPU; ") +
                             ("first 4 Mbp bit-exact vs oracle (bins + digest)" if ok else "ORACLE MISMATCH on the 4 Mbp prefix"))
            assert same and ok, out["parity"]
            s5 = 200_000_000
            hb5 = ko.generate_bases(SEED5, 0, s5, THRESH5)
            t0 = time.perf_counter()
            ko.extract_canonical(hb5, K, n_reads=1000, fixed_len=s5 // 1000, n_threads=cores, materialize=False, hist_bits=HIST_BITS)
            out["cpu_port"] = {"value": (s5 - 30 * 1000) / (time.perf_counter() - t0), "unit": UNIT, "cores": cores,
                               "sample": "200 Mbp as 1000 chunks, iterator + 2^16-bin histogram"}
        barrier()
        return out if rank == 0 else None
    finally:
        ctx5.close()


def multi_gpu_parity(torch, np, dist, kb, ctx, rank, world, local_rank, ko):
    """NCCL on real ranks, checked against the oracle: (1) dist.sharded_histogram (torch.distributed / NCCL, one process per
    GPU) on a 3 Mbp sequence with N; (2) kmb_allreduce_u64 -- the single-process form a Rust host uses -- on 2 contexts on
    2 GPUs, from a side process of rank 0."""
    from kmers_b200.dist import sharded_histogram
    res = {}
    n, k, bits = 3_000_017, 31, 16
    ctxp = kb.Context(local_rank)
    try:
        g_hist, g_dig, _ = sharded_histogram(ctxp, 99, n, k, bits, rank, world, n_thresh20=3000)
        ctxp.sync()
        if rank == 0:
            ref = ko.extract_canonical(ko.generate_bases(99, 0, n, 3000), k, n_reads=1, fixed_len=n, hist_bits=bits, materialize=False)
            ok = bool(np.array_equal(g_hist.cpu().numpy().view(np.uint64), ref["hist"])) and tuple(g_dig) == (ref["n_valid"], ref["checksum_canon"], ref["checksum_hash"])
            res["sharded_histogram_nccl"] = f"{'ok' if ok else 'MISMATCH'}: {world} ranks, 3 Mbp with N, reduced bins + digest vs oracle"
            assert ok, res
    finally:
        ctxp.close()
    if rank == 0:
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {ROOT!r})\n"
            "import kmers_b200 as kb\n"
            "from kmers_b200.dist import allreduce_single_process, shard_sequence\n"
            "import oracle as ko\n"
            "n, k, bits = 2_000_003, 31, 16\n"
            "ctxs = [kb.Context(d) for d in (0, 1)]\n"
            "bufs = []\n"
            "for r, c in enumerate(ctxs):\n"
            "    torch.cuda.set_device(r)\n"
            "    s, e, le = shard_sequence(n, k, r, 2)\n"
            "    b = c.generate(7, 1, le - s, n_thresh20=2000, first_index=s)\n"
            "    buf = torch.zeros((1 << bits) + 3, dtype=torch.int64, device=f'cuda:{r}')\n"
            "    b.histogram(k, bits, hist=buf, accumulate=False, digest_in_hist=True)\n"
            "    c.sync(); bufs.append(buf)\n"
            "allreduce_single_process(ctxs, bufs)\n"
            "ref = ko.extract_canonical(ko.generate_bases(7, 0, n, 2000), k, n_reads=1, fixed_len=n, hist_bits=bits, materialize=False)\n"
            "want = np.concatenate([ref['hist'], np.array([ref['n_valid'], ref['checksum_canon'], ref['checksum_hash']], dtype=np.uint64)])\n"
            "ok = all(np.array_equal(b.cpu().numpy().view(np.uint64), want) for b in bufs)\n"
            "print('ok' if ok else 'MISMATCH')\n")
        try:
            p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=240,
                               env={k_: v for k_, v in os.environ.items() if k_ not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_PORT")})
            last = (p.stdout.strip().splitlines() or ["no output"])[-1]
            res["kmb_allreduce_u64_2_contexts"] = (f"{last}: one process, 2 contexts on GPUs 0 and 1, fused [bins|digest] buffers reduced in place "
                                                   "through the C ABI, vs oracle") if p.returncode == 0 else f"failed rc={p.returncode}: {p.stderr[-300:]}"
        except Exception as ex:
            res["kmb_allreduce_u64_2_contexts"] = f"failed: {type(ex).__name__}: {str(ex)[:200]}"
    if world > 1:
        dist.barrier()
    return res if rank == 0 else None


if __name__ == "__main__":
    main()
