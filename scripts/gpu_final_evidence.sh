#!/bin/bash
# final evidence of round 2 on ONE GPU (after the compaction / rev_comp / pack changes): GPU tests, launch list + ncu --set full of
# the bench kernel, per-config timings, secondary-kernel captures, the bench line itself
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_final.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_r02_final.log
bash scripts/profile.sh r02 > gpurun_out/r02_profile_sh.log 2>&1; tail -3 gpurun_out/r02_profile_sh.log
timeout 900 python scripts/bench_configs.py --out gpurun_out/configs_r02.json > gpurun_out/configs_r02.log 2>&1; echo "configs rc=$?"
bash scripts/profile_kernels.sh r02m "compact1 revcomp pack64" > gpurun_out/r02_prof_m.log 2>&1; tail -3 gpurun_out/r02_prof_m.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench rc=$?"; head -c 300 gpurun_out/bench_r02_final.json
