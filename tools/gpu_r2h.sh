#!/bin/bash
# validation of the 16-batches-per-launch build: GPU tests, smoke, default bench line, the ncu
# launch list + --set full capture (tools/gpu_prof.sh), and a short sweep of the pool's knobs.
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout -s KILL 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_tests.log 2>&1; tail -3 gpurun_out/r2h_tests.log
echo "tests done at $((SECONDS-T0)) s"
timeout -s KILL 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 600 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; tail -c 600 gpurun_out/r2h_bench.err
echo "bench done at $((SECONDS-T0)) s"
bash tools/gpu_prof.sh r2h 2>&1 | grep rc=
echo "prof done at $((SECONDS-T0)) s"
Q="--steps 20 --warmup 5 --min-seconds 0.3 --trials 1 --no-cpu-baseline --no-e2e --no-train --no-seq2seq --no-other-sets --no-other-configs"
for V in "--streams 6" "--proj-ctas 132"; do
  echo "== $V"; timeout -s KILL 120 python bench.py $Q $V 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['roofline']['frac'], d['kernel_us'])"
done
echo "sweep done at $((SECONDS-T0)) s"
