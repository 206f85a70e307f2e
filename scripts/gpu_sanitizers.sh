#!/bin/bash
mkdir -p gpurun_out
timeout 1700 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py tests/test_gpu_parity.py tests/test_gpu_minimizers.py tests/test_gpu_goldens.py -m gpu -x -q > gpurun_out/memcheck_r02_final.log 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/memcheck_r02_final.log
timeout 1700 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/racecheck_r02_final.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/racecheck_r02_final.log
