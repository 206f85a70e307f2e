#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02w.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02w.log
for w in compact1 compact; do python scripts/prof_one.py $w --time; done
timeout 600 python scripts/bench_configs.py --gpu-only > gpurun_out/configs_r02e.json 2> gpurun_out/configs_r02e.log; echo "configs rc=$?"; tail -5 gpurun_out/configs_r02e.log
bash scripts/profile_kernels.sh r02j "compact1"
