#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_goldens.py tests/test_gpu_parity.py tests/test_gpu_errors.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
python scripts/prof_one.py revcomp --time
bash scripts/profile_kernels.sh r02l "revcomp"
