timeout 900 python scripts/bench_configs.py --out gpurun_out/configs_r01.json > gpurun_out/configs_r01.log 2>&1; grep -E "config2" gpurun_out/configs_r01.log | cut -c1-140
