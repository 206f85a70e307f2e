timeout 600 bash scripts/profile.sh r01z
timeout 600 python bench.py > gpurun_out/bench_r01z.json 2> gpurun_out/bench_r01z.err; tail -c 400 gpurun_out/bench_r01z.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
