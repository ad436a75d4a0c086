#!/bin/bash
set +e
mkdir -p gpurun_out
TAG=${1:-x}
ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,dram__bytes_read.sum --clock-control none -s 60 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e --no-train --streams 1 > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
