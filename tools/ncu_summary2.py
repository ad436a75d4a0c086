#!/usr/bin/env python
"""profiles/<out>.md from an ncu launch list (csv) and a `--set full` report of some of its kernels
(train step / seq2seq captures of tools/gpu_prof2.sh). Run here, no GPU needed.
usage: ncu_summary2.py <title> <launches.csv> <report.ncu-rep> <out.md> [marker kernel]"""
import collections
import csv
import statistics
import subprocess
import sys

title, lcsv, rep, outp = sys.argv[1:5]
marker = sys.argv[5] if len(sys.argv) > 5 else None
out = open(outp, 'w')
out.write('# %s\n\nProduced by `tools/gpu_prof2.sh` + `tools/ncu_summary2.py`. Launch list: '
          '`ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none` '
          '(serialised launches: compare shares, not the sum with the step time). `--set full` rows: '
          'one per captured launch.\n\n' % title)
rows = [r for r in csv.reader(open(lcsv)) if len(r) > 5]
hdr = [i for i, r in enumerate(rows) if r[0] == 'ID'][0]
h, data = rows[hdr], rows[hdr + 1:]
ki, vi, gi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size')
seq = [(r[ki].split('(')[0].replace('void ', '').replace('n2nmn::', '').replace('<unnamed>::', ''),
        float(r[vi].replace(',', '')) / 1e3, r[gi]) for r in data]
if marker:   # one step = from one marker launch to the next
    st = [i for i, s in enumerate(seq) if marker in s[0]]
    seq = seq[st[-2]:st[-1]]
    out.write('## one step, launch by launch\n\n| kernel | grid | us |\n|---|---|---|\n')
    for n, t, g in seq:
        out.write('| %s | %s | %.1f |\n' % (n, g, t))
    out.write('\nsum %.1f us over %d launches\n\n' % (sum(t for _, t, _ in seq), len(seq)))
d = collections.OrderedDict()
for n, t, g in seq:
    d.setdefault(n, []).append(t)
tot = sum(sum(v) for v in d.values())
out.write('## by kernel\n\n| kernel | launches | mean us | total us | share |\n|---|---|---|---|---|\n')
for k, v in d.items():
    out.write('| %s | %d | %.1f | %.1f | %.1f%% |\n' % (k, len(v), statistics.mean(v), sum(v), 100 * sum(v) / tot))
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
if len(rows) > 2:
    h = rows[0]
    want = ['launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
            'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor',
            'gpu__time_duration.sum', 'sm__cycles_elapsed.max',
            'sm__throughput.avg.pct_of_peak_sustained_elapsed',
            'sm__warps_active.avg.pct_of_peak_sustained_active',
            'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
            'sm__inst_executed_pipe_tensor.sum', 'smsp__inst_executed.sum',
            'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
            'dram__bytes_read.sum', 'dram__bytes_write.sum',
            'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_op_red.sum',
            'lts__t_sectors_op_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
            'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
            'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio']
    out.write('\n## --set full\n\n')
    for r in rows[2:]:
        out.write('- **%s** grid %s\n' % (r[h.index('Kernel Name')].split('(')[0],
                                         r[h.index('launch__grid_size')] if 'launch__grid_size' in h else '?'))
        for w in want:
            if w in h:
                out.write('  - %s = %s %s\n' % (w, r[h.index(w)], rows[1][h.index(w)]))
out.close()
print(open(outp).read()[:3000])
