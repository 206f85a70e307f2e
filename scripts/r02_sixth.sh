#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py tests/test_gpu_minimizers.py tests/test_gpu_parity.py tests/test_gpu_goldens.py tests/test_gpu_accessors.py -m gpu -x -q > gpurun_out/r02_pytest_f.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_f.log
tail -4 gpurun_out/r02_pytest_f.log
for w in minword unpack compact1; do python scripts/prof_one.py $w --time; done
bash scripts/profile_kernels.sh r02f "compact1" > gpurun_out/r02_prof_f.log 2>&1
