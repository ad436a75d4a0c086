#!/bin/bash
set +e
python - <<'PY'
import torch, time
x = torch.empty(64*150*512, dtype=torch.float32).pin_memory()
d = torch.empty_like(x, device='cuda')
for _ in range(3): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(50): d.copy_(x, non_blocking=True)
torch.cuda.synchronize(); el=time.perf_counter()-t0
print('raw H2D pinned: %.1f GB/s  -> max %.0f q/s' % (50*x.numel()*4/el/1e9, 64*50/el))
PY
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], "e2e", d["e2e"]["value"])'
for ns in 2 4 8 12; do
echo -n "streams $ns: "; timeout -s KILL 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-train --streams $ns 2>&1 | tail -1 | python -c "$P"
done
echo -n "streams 4 latency-mode knobs: "; timeout -s KILL 300 python bench.py --steps 1000 --warmup 50 --no-cpu-baseline --no-train --streams 4 --tree-cluster 4 --proj-ctas 0 2>&1 | tail -1 | python -c "$P"
