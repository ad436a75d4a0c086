#!/bin/bash
set +e
for v in "" st3; do
  if [ -z "$v" ]; then unset N2NMN_LIB; else export N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_$v.so; fi
  for bsz in 64 256 1024; do
    PB_BATCH=$bsz PB_LAYOUT=expert timeout -s KILL 300 python tools/proj_bench.py 2>&1 | tail -1
  done
done
