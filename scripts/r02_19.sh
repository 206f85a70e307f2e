#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_compact.py -m gpu -x -q > gpurun_out/memcheck_compact_r02.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/memcheck_compact_r02.log
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_compact.py -m gpu -x -q -k "not fullsize" > gpurun_out/racecheck_compact_r02.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/racecheck_compact_r02.log
