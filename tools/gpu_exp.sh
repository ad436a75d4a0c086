#!/bin/bash
set +e
for v in "" skipA skipB skipMMA; do
  if [ -z "$v" ]; then unset N2NMN_LIB; else export N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_$v.so; fi
  timeout -s KILL 200 python tools/proj_bench.py 2>&1 | tail -1
done
unset N2NMN_LIB
PB_LAYOUT=expert timeout -s KILL 200 python tools/proj_bench.py 2>&1 | tail -1
PB_BATCH=128 timeout -s KILL 200 python tools/proj_bench.py 2>&1 | tail -1
PB_BATCH=32 timeout -s KILL 200 python tools/proj_bench.py 2>&1 | tail -1
