#!/bin/bash
set +e
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_timeline.so timeout -s KILL 300 python tools/timeline.py 2>&1 | grep -E "^transform|^pooled|cycles per node" -A4 | head -24
B="bench.py --steps 2000 --warmup 50 --no-cpu-baseline --no-train --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "host", d["host_enqueue_ms_per_step"], d.get("kernel_us"))'
for ns in 1 4 8; do
echo -n "N=1 streams $ns: "; timeout -s KILL 300 python $B --streams $ns > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
done
