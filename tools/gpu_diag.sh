#!/bin/bash
# GPU check: full GPU test-suite, smoke, short bench. Logs in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
echo "== GPU tests"
timeout -s KILL 1200 python -m pytest tests -m gpu -q -s -x > gpurun_out/t_gpu.log 2>&1
echo "rc=$?"; grep -E "err|passed|failed|Error|error" gpurun_out/t_gpu.log | tail -40
echo "== smoke"
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/smoke.log
echo "== bench (short)"
timeout -s KILL 600 python bench.py --steps 100 --warmup 10 --cpu-seconds 5 > gpurun_out/bench_short.log 2>&1
echo "rc=$?"; tail -2 gpurun_out/bench_short.log
