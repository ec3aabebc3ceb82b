"""profiles/ncu_traffic.json from an `ncu --set full` report: DRAM bytes (read + write) per launch of every
kernel, averaged over the captured launches.  usage: ncu -i rep.ncu-rep --page raw --csv | python tools/ncu_traffic.py"""
import csv, json, os, re, sys
rows = list(csv.reader(sys.stdin))
h = rows[0]
units = rows[1]
kn, rd, wr = h.index('Kernel Name'), h.index('dram__bytes_read.sum'), h.index('dram__bytes_write.sum')
scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
acc = {}
for r in rows[2:]:
    name = re.sub(r'<.*', '', r[kn].split('(')[0]).split('::')[-1].replace('void ', '').strip()
    b = float(r[rd].replace(',', '')) * scale[units[rd]] + float(r[wr].replace(',', '')) * scale[units[wr]]
    acc.setdefault(name, []).append(b)
out = {k: sum(v) / len(v) for k, v in acc.items()}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'profiles', 'ncu_traffic.json')
old = json.load(open(path)) if os.path.exists(path) else {}
old.update(out)
json.dump(old, open(path, 'w'), indent=1, sort_keys=True)
print(json.dumps(out, indent=1))
