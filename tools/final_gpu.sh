timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_final_ref.json 2>/dev/null
bash tools/gpu_prof2.sh r2g 2>&1 | tail -3
