#!/bin/bash
# end-of-round verification: GPU parity tests, smoke, default bench line, launch list
set +e
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 1500 gpurun_out/bench_default.json; tail -2 gpurun_out/bench_default.err
timeout -s KILL 600 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
TAG=${1:-final}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"text_proj|proj_umma|tree_kernel" -s 30 -c 60 --csv \
    --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 30 --warmup 10 --no-cpu-baseline --no-e2e --no-train > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list rc=$?"
