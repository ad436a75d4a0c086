timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
GB_ONLY=8 timeout 300 python tools/group_bench.py 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-other-configs --no-cpu-baseline > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -c 300 gpurun_out/bench_now.err
