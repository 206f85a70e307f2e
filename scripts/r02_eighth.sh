#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_gpu_fullsize.py > gpurun_out/r02_pytest_h.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_h.log
tail -4 gpurun_out/r02_pytest_h.log
for w in csr csr_var csr_wide csr_min; do python scripts/prof_one.py $w --time; done
