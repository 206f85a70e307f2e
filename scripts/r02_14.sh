#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py tests/test_fastx.py -m gpu -x -q 2>&1 | tail -3
for w in compact1 compact; do python scripts/prof_one.py $w --time; done
bash scripts/profile_kernels.sh r02i "compact1"
