#!/bin/bash
# validation of the sampled decoding and the fp16 host-feature path: GPU tests, smoke, a short
# bench line with both end-to-end legs
set +e
mkdir -p gpurun_out
T0=$SECONDS
timeout -s KILL 400 python -m pytest tests -m gpu -x -q > gpurun_out/r2i_tests.log 2>&1; tail -4 gpurun_out/r2i_tests.log
echo "tests done at $((SECONDS-T0)) s"
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 200 python bench.py --no-cpu-baseline --no-train --no-seq2seq --no-other-sets --no-other-configs \
    > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err; tail -c 400 gpurun_out/r2i_bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2i_bench.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['achieved_h2d_gbs_per_gpu'])
e = d['e2e_f16_host_features']
print('e2e f16', e['value'], e['achieved_h2d_gbs_per_gpu'], e['max_abs_score_diff_vs_f32_device_path'])
PY
echo "bench done at $((SECONDS-T0)) s"
