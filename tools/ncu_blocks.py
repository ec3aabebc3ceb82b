"""Group the SASS lines of `ncu --page source --csv` into runs with equal execution counts
(basic blocks) and print each run's share of instructions and of stall samples.
usage: ncu_blocks.py source.csv [kernel segment index] [min share %]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if 'Instructions Executed' in r][0]
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
segs, cur = [], []
for r in rows[hi + 1:]:
    try:
        ex = int(r[ix['Instructions Executed']])
    except (ValueError, IndexError):
        if cur:
            segs.append(cur)
            cur = []
        continue
    cur.append((ex, r[ix['Source']].strip(), int(r[ix['Warp Stall Sampling (All Samples)']] or 0)))
if cur:
    segs.append(cur)
print("segments (lines, instructions):", [(len(s), sum(d[0] for d in s)) for s in segs])
data = segs[int(sys.argv[2]) if len(sys.argv) > 2 else 0]
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.6
tot = sum(d[0] for d in data)
ts = max(1, sum(d[2] for d in data))
i = 0
while i < len(data):
    j = i
    while j + 1 < len(data) and abs(data[j + 1][0] - data[i][0]) <= 0.02 * data[i][0]:
        j += 1
    sub = sum(d[0] for d in data[i:j + 1])
    ss = sum(d[2] for d in data[i:j + 1])
    if sub > thr / 100 * tot:
        print("lines %4d-%4d n=%3d ex=%9d instr=%5.1f%% samples=%5.1f%%  %s | %s" %
              (i, j, j - i + 1, data[i][0], 100 * sub / tot, 100 * ss / ts, data[i][1][:40], data[j][1][:40]))
    i = j + 1
