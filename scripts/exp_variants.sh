#!/bin/bash
# Kernel experiments: run bench.py's device-resident leg with alternative builds of the library.
# usage (on a GPU box): bash scripts/exp_variants.sh build/exp/*.so
mkdir -p gpurun_out
python - <<'PY'
import torch, time
n = 1 << 31  # 2 GiB
a = torch.empty(n, dtype=torch.uint8, device="cuda"); b = torch.empty(n, dtype=torch.uint8, device="cuda")
def t(f, reps=10):
    f(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); f(); e.record(); torch.cuda.synchronize(); best = min(best, s.elapsed_time(e))
    return best
ms = t(lambda: a.fill_(7)); print(f"torch fill_ (pure write) : {n/ms/1e6:.1f} GB/s")
ms = t(lambda: b.copy_(a)); print(f"torch copy_ (read+write) : {2*n/ms/1e6:.1f} GB/s")
ms = t(lambda: a.sum()); print(f"torch sum   (pure read)  : {n/ms/1e6:.1f} GB/s")
PY
for so in "" "$@"; do
  KMERS_B200_SO=$so python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('%-28s avg %.4f ms best %.4f ms  %.1f GB/s  frac %.4f' % ('${so:-default}', r['avg_launch_ms'], r['best_launch_ms'], r['achieved'], r['frac']))"
done
