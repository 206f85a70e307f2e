#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py tests/test_gpu_packed.py -m gpu -x -q > gpurun_out/r02_pytest_g.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_g.log
tail -4 gpurun_out/r02_pytest_g.log
python scripts/prof_one.py compact1 --time
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_r02b.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_r02b_ref.json 2> gpurun_out/bench_r02b_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_r02b_ref.json | cut -c1-600
