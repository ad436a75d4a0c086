#!/bin/bash
set +e
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
B="bench.py --steps 8000 --warmup 100 --no-cpu-baseline --no-train --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], d["kernel_us"])'
echo -n "default: "; timeout -s KILL 300 python $B > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
echo -n "latency mode (1 stream): "; timeout -s KILL 300 python $B --streams 1 > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
