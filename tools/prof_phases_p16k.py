import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
from flechasdb_b200 import engine
nq, nprobe, m, p = 10000, int(sys.argv[1]), 20_000_000, 16384
n, d, cn, k = 96, 12, 256, 10
rng = np.random.default_rng(7)
ctx = engine.Context(0)
coarse = rng.random((p, n), dtype=np.float32)
cbs = rng.random((d, cn, n // d), dtype=np.float32) - np.float32(0.5)
sizes = rng.multinomial(m, np.ones(p) / p)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
codes = rng.integers(0, 256, (m, d), dtype=np.uint8)
ix = engine.Index.create(ctx, coarse, cbs, off, codes)
d_q = ctx.alloc(nq * n * 4)
ctx.fill_uniform(d_q, nq * n, 5)
outs = [ctx.alloc(nq * k * 4) for _ in range(3)] + [ctx.alloc(nq * 4)]
ix.set_timing(True)
for mode in (0, 1):
    for it in range(3):
        ctx.timer_start()
        ix.query_device(d_q, nq, k, nprobe, *outs, mode=mode)
        t = ctx.timer_stop()
    print("mode", mode, "nprobe", nprobe, "total ms", round(t, 3), "phases", [round(float(x), 3) for x in ix.last_timing()[0]], ix.last_stats())
