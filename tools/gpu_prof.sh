#!/bin/bash
# ncu launch list of bench.py + full capture of the forward kernels as launched for a group of 16
# batches (tools/group_bench.py). Numbers printed under ncu are never bench values; this only
# produces the inputs of tools/ncu_summary.py <tag> 16 -> profiles/<tag>.md.
set +e
mkdir -p gpurun_out
TAG=${1:-r2}
K='regex:text_proj|quad_kernel|proj_umma|tree_kernel|pool_kernel|head_kernel'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -s 60 -c 120 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 40 --warmup 10 --min-seconds 0.02 --trials 1 \
    --no-cpu-baseline --no-e2e --no-train --no-other-configs --no-other-sets > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
GB_ONLY=16 GB_ITERS=4 ncu --set full --clock-control none --import-source on -k "$K" -s 24 -c 6 \
    -o gpurun_out/prof_$TAG -f python tools/group_bench.py > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la gpurun_out | tail -5
