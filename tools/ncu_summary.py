#!/usr/bin/env python
"""Summarise gpurun_out/launches_<tag>.csv and prof_<tag>.ncu-rep into profiles/<tag>.md
(run here, no GPU needed; ncu reads the report offline)."""
import collections
import csv
import statistics
import subprocess
import sys

tag = sys.argv[1]
group = int(sys.argv[2]) if len(sys.argv) > 2 else 16
out = open('profiles/%s.md' % tag, 'w')
out.write('# ncu summary %s\n\n' % tag)
out.write('Command: `bash tools/gpu_prof.sh %s` (launch list: bench.py under `ncu --metrics '
          'gpu__time_duration.sum --clock-control none`, the pool batching %d batches per launch '
          'set; `--set full`: one set of launches of `GB_ONLY=%d tools/group_bench.py`, i.e. '
          'n2nmn_forward_group over %d batches). Per-launch times are cold-cache and serialised: '
          'compare shares.\n\n' % (tag, group, group, group))
rows = [r for r in csv.reader(open('gpurun_out/launches_%s.csv' % tag)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h, data = rows[hdr], rows[hdr + 1:]
ki, vi, mi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name')
d = collections.OrderedDict()
for r in data:
    if r[mi] == 'gpu__time_duration.sum':
        d.setdefault(r[ki].split('(')[0], []).append(float(r[vi].replace(',', '')) / 1e3)
tot = sum(statistics.mean(v) for v in d.values())
out.write('## launch list (gpu__time_duration.sum)\n\n| kernel | launches | mean us | min | max | share of step |\n|---|---|---|---|---|---|\n')
for k, v in d.items():
    out.write('| %s | %d | %.2f | %.2f | %.2f | %.1f%% |\n' % (k, len(v), statistics.mean(v), min(v), max(v), 100 * statistics.mean(v) / tot))
raw = subprocess.run(['ncu', '-i', 'gpurun_out/prof_%s.ncu-rep' % tag, '--page', 'raw', '--csv'],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__cluster_size', 'sm__cycles_elapsed.max',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_bytes.sum']
out.write('\n## --set full (one row per captured launch)\n\n')
for r in rows[2:]:
    out.write('- **%s**\n' % r[h.index('Kernel Name')].split('(')[0])
    for w in want:
        if w in h:
            out.write('  - %s = %s %s\n' % (w, r[h.index(w)], rows[1][h.index(w)]))
traffic = {}
for r in rows[2:]:
    k = r[h.index('Kernel Name')].split('(')[0].replace('void ', '').split('<')[0]
    def val(name):
        v, u = float(r[h.index(name)].replace(',', '')), rows[1][h.index(name)]
        return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    traffic.setdefault(k, []).append((val('dram__bytes_read.sum'), val('dram__bytes_write.sum')))
import json
if traffic:   # a launch-list-only run keeps the previous --set full figures
  json.dump({k: {'dram_read_bytes': sum(x[0] for x in v) / len(v),
               'dram_write_bytes': sum(x[1] for x in v) / len(v), 'launches': len(v),
               'batches_per_launch': group, 'capture': 'profiles/%s.md' % tag}
             for k, v in traffic.items()},
          open('profiles/ncu_traffic.json', 'w'), indent=1)
out.close()
print(open('profiles/%s.md' % tag).read())
