#!/bin/bash
set +e
B="bench.py --steps 4000 --warmup 100 --no-cpu-baseline --no-train --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"], "host", d["host_enqueue_ms_per_step"])'
export N2NMN_TREE_CLUSTER=1
for pc in 24 37 50 74; do for ns in 12 16 24; do
echo -n "cluster 1 proj_ctas $pc streams $ns: "; N2NMN_PROJ_CTAS=$pc timeout -s KILL 300 python $B --streams $ns > gpurun_out/b.log 2>&1; tail -1 gpurun_out/b.log | python -c "$P" || tail -20 gpurun_out/b.log
done; done
