#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_goldens.py tests/test_gpu_packed.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
for w in pack64 pack8 revcomp; do python scripts/prof_one.py $w --time; done
bash scripts/profile_kernels.sh r02l "pack64 revcomp"
