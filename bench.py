#!/usr/bin/env python
"""bench.py -- canonical k-mers/s (K=31, 150 bp reads) on N B200s, next to the CPU path.

One "step" = one pass of the hot path (pack -> windows -> reverse complement ->
canonical min -> LexHash, materialised) over one resident batch of synthetic
reads: BASELINE.json configs[1], 10^7 x 150 bp per GPU (weak scaling: reads shard
by batch, no data-path collective).  Prints ONE JSON line (rank 0).

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


def cpu_reference(steps: int, warmup: int, sample_reads: int):
    """The reference's CPU path (oracle port) on all host threads, on a bounded sample of the workload."""
    import oracle as ko
    cores = len(os.sched_getaffinity(0)) or 1
    bases = ko.generate_bases(SEED, 0, sample_reads * READ_LEN)
    n_kmers = sample_reads * (READ_LEN - K + 1)

    import numpy as np
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
    sample = 2_000_000
    r = cpu_reference(args.steps, min(args.warmup, 2), sample)
    sample_desc = (f"first {sample} of the {READS_PER_GPU} synthetic 150 bp reads (seed {SEED}), CanonicalKmerIterator + "
                   f"get_canonical_word + LexHasher, canon+hash materialised to host arrays, {r['cores']} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "configs[1]: K=31 canonical k-mer extraction + LexHash over synthetic 150bp reads",
                   "k": K, "read_len": READ_LEN, "reads_per_step": sample},
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--reads", type=int, default=READS_PER_GPU, help="reads per GPU (default: configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the e2e leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    if args.impl == "reference":
        run_reference_arm(args)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import kmers_b200 as kb

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

    n_reads, L = args.reads, READ_LEN
    W = L - K + 1
    n_slots = n_reads * W
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)  # torch work, the kernels and the timing events all share this stream
    ctx = kb.Context(local_rank, stream=stream.cuda_stream)
    # rank r owns reads [r*n_reads, (r+1)*n_reads) of one synthetic data set
    batch = ctx.generate(SEED, n_reads, L, n_thresh20=0, first_index=rank * n_reads * L)
    out = kb.CanonicalKmers(k=K, n_slots=n_slots)
    out.canon = torch.empty(n_slots, dtype=torch.int64, device="cuda")
    out.hash = torch.empty(n_slots, dtype=torch.int64, device="cuda")

    # correctness of this very configuration: digest + a prefix against the oracle (outside the timed region)
    res = batch.extract_canonical(K, digest=True, out=out)
    digest = res.digest
    assert digest[0] == n_slots, digest
    csum = int(out.canon.sum().item()) % 2**64
    hsum = int(out.hash.sum().item()) % 2**64
    assert (csum, hsum) == (digest[1] % 2**64, digest[2] % 2**64), "digest != sum of the materialised arrays"
    parity = "digest==sum(arrays)"
    if rank == 0:
        import oracle as ko
        npre = min(n_reads, 100_000)
        ref = ko.extract_canonical(ko.generate_bases(SEED, 0, npre * L), K, n_reads=npre, fixed_len=L, n_threads=os.cpu_count() or 1)
        assert np.array_equal(out.canon[:npre * W].cpu().numpy().view(np.uint64), ref["canon"]), "canonical words != oracle"
        assert np.array_equal(out.hash[:npre * W].cpu().numpy().view(np.uint64), ref["hash"]), "hashes != oracle"
        parity += f"; first {npre} reads bit-exact vs oracle"

    # ---- device-resident timing
    for _ in range(args.warmup):
        batch.extract_canonical(K, out=out)
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
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
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    value = world * n_slots / (ms_per_step / 1e3)

    # ---- roofline of the dominant (only) kernel
    peak, peak_src = hbm_peak()
    avg_kernel_ms = sum(kernel_ms) / len(kernel_ms)
    alg = algorithmic_bytes(n_reads, L, K)
    achieved = alg / (avg_kernel_ms / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "kmb::fixed_kernel<NarrowEng<validate=1,digest=0,fwrc=0,mode=materialise,K>16,hash=1>>", "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg, "avg_launch_ms": avg_kernel_ms, "best_launch_ms": min(kernel_ms),
                "frac_of_nominal_8TBs": achieved / 8000.0}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            with open(prof) as f:
                t = json.load(f)
            roofline["traffic"] = t.get("dram_bytes_per_launch_at_bench_size")
            roofline["traffic_source"] = t.get("source")
        except Exception:
            pass

    # ---- e2e: host buffers through the C-ABI one-shot call (H2D inside, digest read back)
    e2e = None
    if not args.no_e2e:
        host = torch.empty(n_reads * L, dtype=torch.uint8, pin_memory=True)
        host_np = host.numpy()
        host_np[:] = batch.download()
        for _ in range(2):
            ctx.extract_canonical_host(host_np, n_reads, L, K)
        # the transfer floor of this step: the same pinned buffer through one plain cudaMemcpyAsync, nothing else
        dev_tmp = torch.empty(n_reads * L, dtype=torch.uint8, device="cuda")
        dev_tmp.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            dev_tmp.copy_(host, non_blocking=True)
        torch.cuda.synchronize()
        h2d_only_ms = 1e3 * (time.perf_counter() - t0) / 3
        del dev_tmp
        e2e_steps = max(3, min(args.steps, 10))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            d = ctx.extract_canonical_host(host_np, n_reads, L, K)
        barrier()
        dt = time.perf_counter() - t0
        assert d == digest, "e2e digest differs from the resident run"
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * n_slots / (dt / e2e_steps), "unit": UNIT, "h2d_bytes_per_step": n_reads * L, "host_affinity": host_affinity,
               "h2d_copy_only_ms": h2d_only_ms, "frac_of_transfer_floor": h2d_only_ms / (1e3 * dt / e2e_steps),
               "d2h_bytes_per_step": 24, "steps": e2e_steps, "ms_per_step": 1e3 * dt / e2e_steps,
               "what": "kmb_extract_canonical_host: pinned host reads -> chunked H2D overlapped with the kernel; canonical+hash "
                       "arrays stay device-resident, the (n_valid, checksum_canon, checksum_hash) digest is read back"}
        # for transparency: the same call when the caller also wants both result arrays back in HOST memory
        # (16 B per k-mer over PCIe; bounded sample so the pinned buffers stay small)
        try:
            ns = min(n_reads, 2_000_000)
            hc = torch.empty(ns * W, dtype=torch.int64, pin_memory=True)
            hh = torch.empty(ns * W, dtype=torch.int64, pin_memory=True)
            hc_np, hh_np = hc.numpy().view(np.uint64), hh.numpy().view(np.uint64)
            ctx.extract_canonical_host(host_np[:ns * L], ns, L, K, host_canon=hc_np, host_hash=hh_np)
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.extract_canonical_host(host_np[:ns * L], ns, L, K, host_canon=hc_np, host_hash=hh_np)
            barrier()
            dtm = (time.perf_counter() - t0) / 3
            e2e["materialised_to_host"] = {"value": world * ns * W / dtm, "unit": UNIT, "reads_per_step": ns, "ms_per_step": 1e3 * dtm,
                                           "h2d_bytes_per_step": ns * L, "d2h_bytes_per_step": ns * W * 16,
                                           "what": "same call with pinned host canon+hash arrays: D2H of 16 B per k-mer dominates (PCIe)"}
            del hc, hh, hc_np, hh_np
        except Exception as ex:  # pinned allocation can fail on a small host
            e2e["materialised_to_host"] = {"unavailable": str(ex)[:200]}
        del host, host_np

    # ---- optional final reduction across ranks (checksum / count), NCCL, outside the timed region
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
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
               "sample": f"first {r['sample_reads']} reads of the same workload, iterator + canonical + LexHash materialised, "
                         f"{r['cores']} threads, mean of 10 passes (~{10 * r['ms_per_step'] * r['cores'] / 1e3:.0f} core-seconds)"}
        try:  # SURVEY 8(d) variant 1: what benches/simple_benchmark.rs does per window (O(K) re-encode + rev-comp), canonicalised
            import oracle as ko
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
            "clocks": sampler.summary(), "digest": global_digest,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
