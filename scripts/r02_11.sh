#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_minimizers.py tests/test_gpu_packed.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for w in minimizers csr_min; do python scripts/prof_one.py $w --time; done
