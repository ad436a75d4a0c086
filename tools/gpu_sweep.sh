#!/bin/bash
# bench variants (cluster size, PDL on/off, executor). Each prints one JSON line.
set +e
mkdir -p gpurun_out
B="python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-e2e"
for ns in 1 2 3 4 6; do
  echo "== streams $ns"; timeout -s KILL 300 $B --streams $ns 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'])"
done
for cs in 1 2 4; do
  echo "== cluster $cs (1 stream)"; N2NMN_TREE_CLUSTER=$cs timeout -s KILL 300 $B --streams 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_us'])"
done
echo "== no PDL (1 stream)"; N2NMN_NO_PDL=1 timeout -s KILL 300 $B --streams 1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_us'])"
echo "== wave"; timeout -s KILL 300 $B --wave 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_us'])"
for l in random deep; do
echo "== layouts $l"; timeout -s KILL 300 $B --layouts $l 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['kernel_us'], d['config']['nodes_per_batch'])"
done
