"""Summarise `ncu --page source --csv` (SASS view): hottest instructions by executed count
and by stall samples, with the dominant stall reason."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
data = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr) - 5 or not r[ix["Instructions Executed"]].isdigit():
        continue
    ex = int(r[ix["Instructions Executed"]]); sm = int(r[ix["Warp Stall Sampling (All Samples)"]] or 0)
    top = max(stalls, key=lambda s: int(r[ix[s]] or 0))
    data.append((n, ex, sm, top if int(r[ix[top]] or 0) else "-", r[ix["Source"]].strip()))
tex = sum(d[1] for d in data); tsm = sum(d[2] for d in data)
print("instructions executed (warp-level):", tex, " stall samples:", tsm)
mode = sys.argv[2] if len(sys.argv) > 2 else "samples"
key = (lambda d: -d[2]) if mode == "samples" else (lambda d: -d[1])
for d in sorted(data, key=key)[: int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print("%5d ex=%9d (%4.1f%%) smp=%6d (%4.1f%%) %-16s %s" % (d[0], d[1], 100.0 * d[1] / tex, d[2], 100.0 * d[2] / max(tsm, 1), d[3], d[4][:90]))
if mode == "dump":
    for d in data:
        print("%5d ex=%9d smp=%6d %-16s %s" % (d[0], d[1], d[2], d[3], d[4][:100]))
