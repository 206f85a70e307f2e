timeout 600 python -m pytest tests/test_gpu_compact.py -x -q -m gpu 2>&1 | tail -2
timeout 120 python scripts/prof_one.py compact --time 2>&1 | tail -1
