#!/bin/bash
set +e
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for v in "" a3b3; do
  if [ -z "$v" ]; then unset N2NMN_LIB; else export N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_$v.so; fi
  for lay in find expert; do PB_LAYOUT=$lay timeout -s KILL 200 python tools/proj_bench.py 2>&1 | tail -1 | cut -c1-220; done
  for bsz in 256 1024; do PB_BATCH=$bsz PB_LAYOUT=expert timeout -s KILL 300 python tools/proj_bench.py 2>&1 | tail -1| cut -c1-220; done
done
unset N2NMN_LIB
B="python bench.py --steps 1500 --warmup 50 --no-cpu-baseline --no-e2e --no-train"
for ns in 1 4 6; do
  echo -n "streams $ns: "; timeout -s KILL 300 $B --streams $ns 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d.get('kernel_us'))"
done
