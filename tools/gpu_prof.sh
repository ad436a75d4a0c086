#!/bin/bash
# ncu launch list + full capture of the three hot kernels. Numbers printed under ncu are never
# bench values; this only produces profiles/.
set +e
mkdir -p gpurun_out
TAG=${1:-r1}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"text_proj|proj_umma|tree_kernel" -s 30 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e --no-train > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"proj_umma|tree_kernel|text_proj" -s 30 -c 6 \
    -o gpurun_out/prof_$TAG -f python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-train > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full rc=$?"
ls -la gpurun_out
