#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_packed.py tests/test_gpu_accessors.py -m gpu -x -q > gpurun_out/r02_pytest_i.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_i.log
tail -30 gpurun_out/r02_pytest_i.log
