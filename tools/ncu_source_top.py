"""Top stall sites of one kernel from an `ncu -i rep --page source --csv` export (possibly holding several launches).
usage: python tools/ncu_source_top.py <source.csv[.gz]> [launch index=0] [top=50]"""
import csv, gzip, sys
path = sys.argv[1]
fh = gzip.open(path, 'rt') if path.endswith('.gz') else open(path)
rows = list(csv.reader(fh))
starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
k = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top_n = int(sys.argv[3]) if len(sys.argv) > 3 else 50
a = starts[k]
b = starts[k + 1] if k + 1 < len(starts) else len(rows)
print(len(starts), 'launches in file;', rows[a][1][:90])
h = rows[a + 1]
si, src, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
stalls = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
data = [r for r in rows[a + 2:b] if len(r) > max(stalls)]
tot = sum(int(r[si]) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = {}
for r in data:
    for j in stalls:
        agg[h[j]] = agg.get(h[j], 0) + int(r[j])
print('stall totals:', sorted(agg.items(), key=lambda kv: -kv[1])[:8])
top = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top_n]
for i in sorted(top):
    r = data[i]
    st = sorted([(int(r[j]), h[j]) for j in stalls], reverse=True)[:2]
    print('%5d %6s %9s  %-72s %s' % (i, r[si], r[ie], r[src].strip()[:72], st))
