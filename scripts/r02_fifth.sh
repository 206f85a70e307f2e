#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dist.py tests/test_gpu_fullsize.py tests/test_gpu_fuzz.py -m gpu -x -q -k "hist or shard or fuzz" > gpurun_out/r02_pytest_e.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_e.log
tail -4 gpurun_out/r02_pytest_e.log
for w in hist histd minword unpack compact1; do python scripts/prof_one.py $w --time; done
bash scripts/profile_kernels.sh r02e "histd" > gpurun_out/r02_prof_e.log 2>&1
