#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
for w in csr csr_var csr_wide csr_min; do python scripts/prof_one.py $w --time; done
bash scripts/profile_kernels.sh r02k "csr_var"
