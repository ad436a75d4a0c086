"""Per-kernel device times of ONE train step from an ncu launch list (tools/train_gpu.sh)."""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); gi = hdr.index('Grid Size')
seq = [(r[ki].split('(')[0][-60:], float(r[vi]) / 1e3, r[gi]) for r in rows[1:]]
starts = [i for i, s in enumerate(seq) if 'text_proj_kernel' in s[0]]
a, b = starts[-2], starts[-1]
tot = 0
for n, t, g in seq[a:b]:
    print('%-62s %8.1f us %s' % (n, t, g)); tot += t
print('sum %.1f us over %d launches' % (tot, b - a))
