"""A BASELINE.json configs[2]-shaped build on one GPU (M = 1M, N = 768, D = 48, P = 1024, C = 256) with
phase times, then a query batch answered by the filter path and by the exact pipeline (must be equal)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine
from flechasdb_b200.db import DatabaseBuilder

M, N, P, D, CN = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (1000000, 768, 1024, 48, 256)
NQ, K, NPROBE = 4096, 10, 16


class Seeds:
    def __init__(self):
        self.rng = np.random.default_rng(3)

    def first(self, n, nb):
        return self.rng.integers(0, n, nb).astype(np.uint32)

    def draws(self, nb, count):
        return self.rng.random((nb, count)).astype(np.float32)


ctx = engine.Context(0)
vs = engine.VectorSet.generate(ctx, M, N, 0xF1EC4A5D0001)
prof = {}
ctx.timer_start()
t0 = time.perf_counter()
db = DatabaseBuilder(vs, ctx=ctx, seeds=Seeds(), profile=prof).with_partitions(P).with_divisions(D) \
    .with_clusters(CN).build()
dev_ms = ctx.timer_stop()
print("build M=%d N=%d P=%d D=%d C=%d: %.3f s on the device (%.3f s wall), launches %d" %
      (M, N, P, D, CN, dev_ms * 1e-3, time.perf_counter() - t0, ctx.launches))
print("phases", {k: round(v, 3) for k, v in prof.items()})
ix = db.index
q = np.random.default_rng(5).random((NQ, N), dtype=np.float32)
for _ in range(2):
    got = ix.query(q, K, NPROBE)
t0 = time.perf_counter()
got = ix.query(q, K, NPROBE)
dt = time.perf_counter() - t0
print("query %d x (k=%d, nprobe=%d), host buffers: %.3f ms -> %.0f queries/s, stats %s" %
      (NQ, K, NPROBE, dt * 1e3, NQ / dt, ix.last_stats()))
os.environ["FDB_QUERY_EXACT"] = "1"
t0 = time.perf_counter()
want = ix.query(q, K, NPROBE)
print("exact pipeline: %.3f ms, stats %s" % ((time.perf_counter() - t0) * 1e3, ix.last_stats()))
print("filter path == exact pipeline:", all((g == w).all() for g, w in zip(got, want)))
