#!/bin/bash
set +e
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
B="bench.py --steps 8000 --warmup 100 --no-cpu-baseline --no-train --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["kernel_us"])'
for tc in 0 2 1; do
echo -n "text_ctas $tc: "; N2NMN_TEXT_CTAS=$tc timeout -s KILL 300 python $B > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
done
echo -n "text_ctas 1, streams 16: "; N2NMN_TEXT_CTAS=1 timeout -s KILL 300 python $B --streams 16 > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
