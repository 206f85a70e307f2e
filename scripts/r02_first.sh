#!/bin/bash
# first GPU call of round 2: host probe, new-path tests, knob sweep, bench
mkdir -p gpurun_out
{ lscpu | head -25; nproc; free -g | head -2; nvidia-smi --query-gpu=name,pcie.link.gen.current,pcie.link.width.current --format=csv; } > gpurun_out/r02_host.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_errors.py tests/test_gpu_compact.py tests/test_gpu_packed.py tests/test_gpu_dist.py -m gpu -x -q > gpurun_out/r02_pytest_a.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_a.log
tail -5 gpurun_out/r02_pytest_a.log
timeout 600 python scripts/e2e_sweep.py > gpurun_out/r02_sweep.log 2>&1
tail -40 gpurun_out/r02_sweep.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err
echo "bench rc=$?"; tail -c 3000 gpurun_out/bench_r02a.err
