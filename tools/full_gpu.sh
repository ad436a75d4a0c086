timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -c 400 gpurun_out/bench_now.err
