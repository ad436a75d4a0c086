#!/bin/bash
# cluster size x stream count (throughput, not latency)
set +e
B="python bench.py --steps 1500 --warmup 50 --no-cpu-baseline --no-e2e --no-train"
for cs in 1 2 4; do for ns in 2 4 6 8; do
  echo -n "cluster $cs streams $ns: "; N2NMN_TREE_CLUSTER=$cs timeout -s KILL 300 $B --streams $ns 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done; done
