#!/bin/bash
# one-GPU verification pass: GPU tests, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02v.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_r02v.log
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r02v.json 2> gpurun_out/bench_r02v.err; echo "bench rc=$?"; head -c 600 gpurun_out/bench_r02v.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02v_ref.json 2> gpurun_out/bench_r02v_ref.err; echo "ref rc=$?"; head -c 300 gpurun_out/bench_r02v_ref.json
