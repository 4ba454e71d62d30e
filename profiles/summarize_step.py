"""Summarise an ncu CSV of one profiled step taken with
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
        --nvtx --nvtx-include "frost_step" --csv --log-file gpurun_out/step.csv python bench.py --quick --steps 1
Per kernel: launches, summed device time (cold-cache, serialised: compare SHARES), share, DRAM bytes per launch.
usage: python profiles/summarize_step.py gpurun_out/step.csv [traffic.json to write] > profiles/rNN_step_summary.txt"""
import collections
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
h, data = rows[hi], rows[hi + 1:]
kid, kn, mn, mv, mu = h.index('ID'), h.index('Kernel Name'), h.index('Metric Name'), h.index('Metric Value'), h.index('Metric Unit')
launch = collections.OrderedDict()
for r in data:
    if len(r) <= mv:
        continue
    v = float(r[mv].replace(',', ''))
    u = r[mu]
    if r[mn] == 'gpu__time_duration.sum':
        v = v / 1e3 if u in ('ns', 'nsecond') else v * 1e3 if u in ('ms', 'msecond') else v       # -> us
    else:
        v = v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
    launch.setdefault(r[kid], {'name': r[kn].split('(')[0].replace('void ', '')})[r[mn]] = v
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for L in launch.values():
    a = agg[L['name']]
    a[0] += 1
    a[1] += L.get('gpu__time_duration.sum', 0.0)
    a[2] += L.get('dram__bytes_read.sum', 0.0) + L.get('dram__bytes_write.sum', 0.0)
tot = sum(v[1] for v in agg.values())
print('launches %d, sum of kernel durations %.1f us (under ncu: cold caches, serialised - compare shares, not absolutes), DRAM traffic %.2f GB' % (len(launch), tot, sum(v[2] for v in agg.values()) / 1e9))
print('%-46s %5s %12s %7s %16s %10s' % ('kernel', 'n', 'time us', 'share', 'DRAM MB/launch', 'GB/s'))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-46s %5d %12.1f %6.1f%% %16.2f %10.0f' % (k[:46], v[0], v[1], 100 * v[1] / tot, v[2] / v[0] / 1e6, v[2] / v[1] / 1e3 if v[1] else 0))
if len(sys.argv) > 2:
    src = sys.argv[1].replace('gpurun_out/', 'profiles/')
    note = '%s (ncu dram__bytes_read.sum+dram__bytes_write.sum, mean over the launches of one bs=256 step)' % src
    out = {}
    merged = {}
    for k, v in agg.items():
        name = k.replace('frost::', '')
        if v[2] > 0:                                   # every instantiation under its own name: pw_fused_kernel<0>, <1>, <2> ...
            out[name] = {'dram_bytes_per_launch': v[2] / v[0], 'launches': v[0], 'source': note}
        m = merged.setdefault(name.split('<')[0], [0, 0.0])
        m[0] += v[0]
        m[1] += v[2]
    for b, m in merged.items():                        # ... and merged over the instantiations of a template
        if m[1] > 0 and b not in out:
            out[b] = {'dram_bytes_per_launch': m[1] / m[0], 'launches': m[0], 'source': note}
    total_bytes = sum(v[2] for v in agg.values())
    out['step_total'] = {'dram_bytes': total_bytes, 'launches': len(launch), 'kernel_time_us': tot,
                         'source': '%s (ncu dram__bytes_read.sum+dram__bytes_write.sum summed over every launch of one bs=256 step)' % src}
    json.dump(out, open(sys.argv[2], 'w'), indent=1)
