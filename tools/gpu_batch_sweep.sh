#!/bin/bash
set +e
for b in 64 128 256 512 1024; do
  echo "== batch $b"; timeout -s KILL 400 python bench.py --batch $b --pool 6 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-train --streams 2 2>&1 | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); r=d['roofline']
print('q/s %.0f  ms/step %.4f  proj us %.1f  hbm_frac %.3f  tensor_frac %.3f  kernel_us %s' % (d['value'], d['ms_per_step'], r['avg_launch_us'], r['hbm_frac'], r['tensor_frac_of_tf32_peak'], d['kernel_us']))"
done
