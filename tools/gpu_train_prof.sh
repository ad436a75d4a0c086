#!/bin/bash
# launch list of the train step (fwd + bwd + clip + Adam + re-pack): which kernels the 2 ms go to
set +e
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/train_launches.csv python tools/train_prof.py > gpurun_out/train_prof.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/train_launches.csv')) if len(r) > 10]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
acc = collections.OrderedDict()
for r in rows[1:]:
    if r[ix['Metric Name']] != 'gpu__time_duration.sum': continue
    name = r[ix['Kernel Name']].split('(')[0][-60:]
    v = float(r[ix['Metric Value']].replace(',', '')); u = r[ix['Metric Unit']]
    us = v / 1000 if u.startswith('n') else v
    a = acc.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(a[1] for a in acc.values())
for k, (n, us) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print('%-62s n=%4d total %.1f us (%.1f%%)' % (k, n, us, 100 * us / tot))
print('total', tot, 'us over the captured launches')
PY
