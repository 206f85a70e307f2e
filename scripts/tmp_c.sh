#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py tests/test_fastx.py -m gpu -x -q 2>&1 | tail -2
for w in csr_var_compact csr_compact; do python scripts/prof_one.py $w --time | cut -c1-170; done
