#!/bin/bash
# Throughput of the pool as a function of how NARROW each batch runs:
# CTAs per question x contraction grid cap x streams. One line per point.
set +e
B="python bench.py --steps 4000 --warmup 100 --no-cpu-baseline --no-e2e --no-train"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])'
for cs in 4 2 1; do for pc in 0 50 32; do for ns in 4 8 12; do
  echo -n "cluster $cs proj_ctas $pc streams $ns: "
  timeout -s KILL 300 $B --streams $ns --tree-cluster $cs --proj-ctas $pc 2>&1 | tail -1 | python -c "$P"
done; done; done
