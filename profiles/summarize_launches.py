"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals + top launches.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h, data = rows[hi], rows[hi + 1:]
kn, mv, mu = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Unit')
agg, seq = collections.defaultdict(lambda: [0, 0.0]), []
for r in data:
    if len(r) <= mv:
        continue
    t = float(r[mv].replace(',', ''))
    t = t / 1e3 if r[mu] == 'ns' else t * 1e3 if r[mu] == 'ms' else t
    name = r[kn].split('(')[0]
    agg[name][0] += 1
    agg[name][1] += t
    seq.append((name, t))
tot = sum(v[1] for v in agg.values())
print('launches %d, sum of kernel durations %.1f us (cold-cache, serialised: compare shares)' % (len(seq), tot))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-44s n=%4d  %10.1f us  %5.1f%%' % (k[:44], v[0], v[1], 100 * v[1] / tot))
print('--- 20 longest launches (index, kernel, us)')
for i, (k, t) in sorted(enumerate(seq), key=lambda x: -x[1][1])[:20]:
    print(i, k[:44], round(t, 1))
