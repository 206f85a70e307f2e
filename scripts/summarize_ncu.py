#!/usr/bin/env python
"""Turn gpurun_out/prof_<tag>.ncu-rep + launches_<tag>.csv into the tracked summaries under profiles/
(raw metric csv, launch list, traffic.json used by bench.py's roofline.traffic).  Runs on the CPU box."""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
raw = os.path.join(ROOT, "profiles", f"ncu_full_fixed_kernel_{tag}_raw.csv")
pre = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_raw.csv")  # exported on the GPU box by scripts/profile.sh
if os.path.exists(pre):
    with open(pre) as f, open(raw, "w") as g:
        g.write(f.read())
    det = os.path.join(ROOT, "gpurun_out", f"prof_{tag}_details.txt")
    if os.path.exists(det):
        with open(det) as f, open(os.path.join(ROOT, "profiles", f"ncu_full_fixed_kernel_{tag}_details.txt"), "w") as g:
            g.write(f.read())
else:
    rep = os.path.join(ROOT, "gpurun_out", f"prof_{tag}.ncu-rep")
    with open(raw, "w") as f:
        subprocess.check_call(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL)
src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
if os.path.exists(src):
    with open(src) as f, open(os.path.join(ROOT, "profiles", f"launches_{tag}.csv"), "w") as g:
        g.write(f.read())
rows = list(csv.reader(open(raw)))
h, u, v = rows[0], rows[1], rows[2]
d = {h[i]: (u[i], v[i]) for i in range(len(h))}
scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
rd = float(d["dram__bytes_read.sum"][1]) * scale[d["dram__bytes_read.sum"][0]]
wr = float(d["dram__bytes_write.sum"][1]) * scale[d["dram__bytes_write.sum"][0]]
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_sha1: the hash of the CUDA sources the captured kernel was built from)
out = {"dram_bytes_per_launch_at_bench_size": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
       "algorithmic_bytes": 20_700_000_000, "kernel_source_sha1": bench.kernel_source_sha1(),
       "source": f"profiles/{os.path.basename(raw)} (ncu --set full, one launch of the bench kernel, 10^7 x 150 bp, K=31): "
                 "dram__bytes_read.sum + dram__bytes_write.sum"}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes.sum.per_second", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__grid_size",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in keys:
    if k in d:
        print(f"{k:70s} {d[k][1]} {d[k][0]}")
print("kernel:", rows[2][h.index("Kernel Name")] if "Kernel Name" in h else "?")
