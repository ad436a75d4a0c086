#!/bin/bash
# Train-step and seq2seq evidence: ncu launch lists + `--set full` captures of the backward kernels
# of one step and of the seq2seq step kernels. Inputs of tools/ncu_summary2.py.
set +e
mkdir -p gpurun_out
TAG=${1:-r2}
M='--metrics gpu__time_duration.sum --clock-control none --cache-control none'
ncu $M --launch-skip 150 -c 120 --csv --log-file gpurun_out/launches_${TAG}_train.csv \
    python tools/train_prof.py 6 > gpurun_out/ncu_train_list.log 2>&1
echo "train list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:tree_bwd|xtb_mma|text_xgrad|wgrad_umma|bmap_colsum' \
    --launch-skip 80 -c 10 -f -o gpurun_out/prof_${TAG}_train python tools/train_prof.py 6 > gpurun_out/ncu_train_full.log 2>&1
echo "train full rc=$?"
ncu $M --launch-skip 900 -c 180 --csv --log-file gpurun_out/launches_${TAG}_s2s.csv \
    python tools/seq2seq_bench.py > gpurun_out/ncu_s2s_list.log 2>&1
echo "s2s list rc=$?"
ncu --set full --clock-control none --import-source on -k 'regex:lstm_step|dec_attn|s2s_gemm' \
    --launch-skip 1000 -c 5 -f -o gpurun_out/prof_${TAG}_s2s python tools/seq2seq_bench.py > gpurun_out/ncu_s2s_full.log 2>&1
echo "s2s full rc=$?"
ls -la gpurun_out | tail -8
