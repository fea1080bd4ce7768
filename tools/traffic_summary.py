"""Summarise an ncu per-launch DRAM-traffic list of the tcgen05 conv / wgrad launches (tools/gpu_r2g.sh, conv_traffic.csv):
bytes read + written per launch over the LAST captured step (115 launches), and the entry of profiles/conv_traffic.json that
bench.py reports as `roofline.traffic`.
usage: python tools/traffic_summary.py <conv_traffic.csv> <workload> [launches_per_step=115] [--write]"""
import csv, json, os, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
h = rows[hdr]
ii, ki, vi, mi, ui = h.index('ID'), h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Metric Unit')
per = {}
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(',', ''))
    u = r[ui].lower()
    scale = {'byte': 1.0, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(u, 1.0)
    per.setdefault(int(r[ii]), {})[r[mi]] = v * scale
ids = sorted(per)
L = int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 115
last = ids[-L:]
rd = sum(per[i].get('dram__bytes_read.sum', 0.0) for i in last)
wr = sum(per[i].get('dram__bytes_write.sum', 0.0) for i in last)
print('%d launches captured; last step (%d launches): read %.2f GB, written %.2f GB, %.1f MB per launch'
      % (len(ids), L, rd / 1e9, wr / 1e9, (rd + wr) / L / 1e6))
if '--write' in sys.argv:
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'conv_traffic.json')
    d = json.load(open(path))
    d[sys.argv[2]] = {
        'source': 'ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:conv3x3_tc|wgrad_tc '
                  'python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-dropin (%s), last captured step' % os.path.basename(sys.argv[1]),
        'launches_per_step': L, 'dram_read_bytes_per_step': rd, 'dram_write_bytes_per_step': wr,
        'bytes_per_launch': (rd + wr) / L}
    json.dump(d, open(path, 'w'), indent=1)
    print('wrote', path)
