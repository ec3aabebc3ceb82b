"""Code scan on lists that exceed L2 (BASELINE.json configs[4] scaled to one GPU): a synthesised index
of M x 12 u8 codes in P lists, nq queries, k = 10.  Times the scan phase in both scan modes.
usage: prof_scan_large.py [nq nprobe M P [modes]]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine

a = sys.argv[1:]
nq, nprobe = (int(a[0]), int(a[1])) if len(a) >= 2 else (8192, 16)
m, p = (int(a[2]), int(a[3])) if len(a) >= 4 else (40_000_000, 4096)
modes = a[4].split(",") if len(a) >= 5 else ["query", "partition"]
n, d, cn, k = 96, 12, 256, 10
rng = np.random.default_rng(7)
ctx = engine.Context(0)
coarse = rng.random((p, n), dtype=np.float32)
cbs = rng.random((d, cn, n // d), dtype=np.float32) - np.float32(0.5)
sizes = rng.multinomial(m, np.ones(p) / p)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
codes = rng.integers(0, 256, (m, d), dtype=np.uint8)
ix = engine.Index.create(ctx, coarse, cbs, off, codes)
del codes
d_q = ctx.alloc(nq * n * 4)
ctx.fill_uniform(d_q, nq * n, 0xF1EC4A5D0002 + 99)
outs = [ctx.alloc(nq * k * 4) for _ in range(3)] + [ctx.alloc(nq * 4)]
ix.set_timing(True)
res = {}
for mode in modes:
    os.environ["FDB_FILTER_SCAN"] = mode
    ms, tot = [], []
    for it in range(4):
        ctx.flush_l2()
        ctx.timer_start()
        ix.query_device(d_q, nq, k, nprobe, *outs)
        t = ctx.timer_stop()
        phases, nbytes = ix.last_timing()
        if it >= 1:
            ms.append(float(phases[4])); tot.append(t)
    part = np.zeros((nq, k), np.uint32); 
    import ctypes
    from flechasdb_b200 import _capi as capi
    scan = sum(ms) / len(ms)
    print("%-9s nq=%d nprobe=%d M=%d P=%d: scan %.3f ms -> %.0f GB/s algorithmic (%.3f of 6496.8), query %.3f ms, %.0f q/s, stats %s"
          % (mode, nq, nprobe, m, p, scan, nbytes / scan / 1e6, nbytes / scan / 1e6 / 6496.8, sum(tot) / len(tot),
             nq / (sum(tot) / len(tot) * 1e-3), ix.last_stats()), flush=True)
