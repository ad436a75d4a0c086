#!/bin/bash
# epilogue attribution: kernel-only time of the projection kernel per experiment build
set +e
for v in "" e_nostore e_nomath e_nostmath e_none; do
  if [ -z "$v" ]; then unset N2NMN_LIB; else export N2NMN_LIB=$PWD/n2nmn_b200/lib/libn2nmn_b200_$v.so; fi
  for lay in find expert; do
    PB_LAYOUT=$lay timeout -s KILL 200 python tools/proj_bench.py 2>&1 | tail -1 | cut -c1-200
  done
done
