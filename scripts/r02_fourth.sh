#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_d.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_d.log
tail -5 gpurun_out/r02_pytest_d.log
timeout 600 python scripts/bench_configs.py --gpu-only --out gpurun_out/configs_r02d.json > gpurun_out/configs_r02d.log 2>&1
grep -E "minimizer_word|decode|compacted|histogram|K=31 150bp canon" gpurun_out/configs_r02d.log
bash scripts/profile_kernels.sh r02d "hist minword unpack" > gpurun_out/r02_prof_d.log 2>&1
tail -3 gpurun_out/r02_prof_d.log
