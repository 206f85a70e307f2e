#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r02_pytest_b.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_b.log
tail -5 gpurun_out/r02_pytest_b.log
timeout 600 python scripts/bench_configs.py --gpu-only --out gpurun_out/configs_r02b.json > gpurun_out/configs_r02b.log 2>&1
cat gpurun_out/configs_r02b.log | tail -40
timeout 300 python scripts/e2e_sweep.py > gpurun_out/r02_sweep_b.log 2>&1
head -40 gpurun_out/r02_sweep_b.log
