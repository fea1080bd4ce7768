"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-kernel launches and time per step.
usage: python tools/launch_summary.py <launches.csv> [steps-marker-kernel]"""
import collections, csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ki, vi, mi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name')
L = []
for r in rows[hdr + 1:]:
    if len(r) <= vi or r[mi] != 'gpu__time_duration.sum':
        continue
    try:
        v = float(r[vi].replace(',', '')) / 1e6
    except ValueError:
        continue
    L.append((r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', ''), v))
marker = sys.argv[2] if len(sys.argv) > 2 else 'floss_finish'
nsteps = max(1, sum(1 for k, _ in L if k.startswith(marker)))
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in L:
    agg[k[:56]][0] += 1
    agg[k[:56]][1] += v
tot = sum(v[1] for v in agg.values())
print('%d launches, %d steps, %.3f ms of kernel time per step' % (len(L), nsteps, tot / nsteps))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print('%-56s %7.1f/step %8.3f ms/step %5.1f%%' % (k, v[0] / nsteps, v[1] / nsteps, 100 * v[1] / tot))
