timeout 600 python -m pytest tests/test_gpu_minimizers.py -x -q -m gpu 2>&1 | tail -2
for k in minimizers csr_min hist; do timeout 120 python scripts/prof_one.py $k --time 2>&1 | tail -1; done
bash scripts/profile_kernels.sh r01h "hist" 2>&1 | tail -1
