"""Aggregate an ncu --csv launch list (gpu__time_duration.sum) by kernel name."""
import csv, re, sys, collections
f = sys.argv[1]
with open(f) as fh:
    lines = [l for l in fh if l.startswith('"')]
r = csv.reader(lines)
hdr = next(r)
ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit'); gi = hdr.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0
for row in r:
    if len(row) <= vi: continue
    n = re.sub(r'\(.*', '', row[ki])[:80]
    if len(sys.argv) > 2 and sys.argv[2] == 'grid': n += ' ' + row[gi]
    v = float(row[vi].replace(',', ''))
    u = row[ui]
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'nsecond': 1e-6, 'msecond': 1.0}.get(u, 1e-6)
    agg[n][0] += 1; agg[n][1] += v; tot += v
for n, cv in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{cv[1]:9.3f} ms {cv[0]:5d} {100*cv[1]/tot:5.1f}%  {n}")
print(f"total {tot:.3f} ms")
