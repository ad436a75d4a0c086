#!/bin/bash
set +e
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
B="bench.py --steps 2000 --warmup 50 --no-cpu-baseline --no-train"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "host", d["host_enqueue_ms_per_step"], "e2e", d["e2e"] and d["e2e"]["value"], d.get("kernel_us"))'
for ns in 1 4 6 8; do
echo -n "N=1 streams $ns: "; timeout -s KILL 300 python $B --streams $ns > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
done
