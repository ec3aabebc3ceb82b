"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, total and mean
time, and each kernel's share; optional [first, last) launch range.  usage: launch_summary.py csv [first last]"""
import csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
scale = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1.0, 'msecond': 1e3}
data = []
for r in rows[hi + 1:]:
    if len(r) <= mv or not r[mv]:
        continue
    name = re.sub(r'\(.*', '', r[kn]).replace('void ', '').replace('fdb::<unnamed>::', '').replace('<unnamed>::', '').strip()
    data.append((name, float(r[mv].replace(',', '')) * scale.get(r[mu], 1e-3)))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi2 = int(sys.argv[3]) if len(sys.argv) > 3 else len(data)
sel = data[lo:hi2]
tot = sum(v for _, v in sel)
agg = {}
for n, v in sel:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
print("launches %d..%d of %d, total %.1f us" % (lo, hi2, len(data), tot))
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("%-70s n=%5d total=%10.1f us mean=%9.2f us share=%5.1f%%" % (n[:70], c, v, v / c, 100 * v / tot))
