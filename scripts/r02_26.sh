#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_compact.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -2
python scripts/prof_one.py compact1 --time
python scripts/prof_one.py compact1 --time
