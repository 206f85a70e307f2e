#!/bin/bash
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_accessors.py tests/test_gpu_goldens.py tests/test_gpu_packed.py tests/test_gpu_minimizers.py tests/test_gpu_errors.py -m gpu -x -q > gpurun_out/memcheck_r02_a.log 2>&1; echo "memcheck A rc=$?"; tail -3 gpurun_out/memcheck_r02_a.log
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/memcheck_r02_b.log 2>&1; echo "memcheck B rc=$?"; tail -3 gpurun_out/memcheck_r02_b.log
