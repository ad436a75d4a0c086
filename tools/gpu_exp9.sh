#!/bin/bash
set +e
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="bench.py --steps 4000 --warmup 100 --no-cpu-baseline --no-train --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["kernel_us"], d["roofline"]["frac"], d["roofline"].get("full_grid"))'
for st in 0 1; do
echo -n "fp32_stencil=$st default pool: "; N2NMN_FP32_STENCIL=$st timeout -s KILL 300 python $B > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
done
for st in 0 1; do
echo -n "fp32_stencil=$st 1 stream (latency mode): "; N2NMN_FP32_STENCIL=$st timeout -s KILL 300 python $B --streams 1 > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
done
echo -n "e2e default: "; timeout -s KILL 300 python bench.py --steps 2000 --warmup 100 --no-cpu-baseline --no-train 2>&1 | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["e2e"])'
