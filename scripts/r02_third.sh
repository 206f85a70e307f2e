#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compact.py tests/test_gpu_minimizers.py tests/test_gpu_parity.py tests/test_gpu_goldens.py tests/test_gpu_packed.py tests/test_gpu_fuzz.py -m gpu -x -q > gpurun_out/r02_pytest_c.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_c.log
tail -5 gpurun_out/r02_pytest_c.log
timeout 600 python scripts/bench_configs.py --gpu-only --out gpurun_out/configs_r02c.json > gpurun_out/configs_r02c.log 2>&1
grep -E "minimizer_word|decode|compacted|ragged|histogram|pack " gpurun_out/configs_r02c.log
bash scripts/profile_kernels.sh r02c "compact1 hist csr_var minword unpack" > gpurun_out/r02_prof_c.log 2>&1
tail -3 gpurun_out/r02_prof_c.log
timeout 300 python scripts/e2e_sweep.py > gpurun_out/r02_sweep_c.log 2>&1
grep -E "host_pack|hybrid|packed only" gpurun_out/r02_sweep_c.log | head -30
