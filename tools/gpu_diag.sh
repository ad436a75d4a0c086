#!/bin/bash
# First-contact GPU run: staged so that a crash or hang in one stage still leaves logs from the
# others. Everything lands in gpurun_out/.
set +e
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
echo "== stage 1: fp32 CUDA-core projection path (tree + wave)"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -s -k "goldens and (1] or 3])" > gpurun_out/t_simt.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/t_simt.log
echo "== stage 2: tcgen05 projection path"
timeout -s KILL 600 python -m pytest tests/test_gpu_parity.py -q -s -k "goldens and (0] or 2])" > gpurun_out/t_umma.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/t_umma.log
echo "== stage 3: remaining GPU tests"
timeout -s KILL 900 python -m pytest tests/test_gpu_parity.py -q -s -k "not goldens" > gpurun_out/t_rest.log 2>&1
echo "rc=$?"; tail -25 gpurun_out/t_rest.log
echo "== stage 4: smoke"
timeout -s KILL 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/smoke.log
echo "== stage 5: bench (short)"
timeout -s KILL 600 python bench.py --steps 50 --warmup 5 --cpu-seconds 5 > gpurun_out/bench_short.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/bench_short.log
