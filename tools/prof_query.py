"""Profiling driver: a synthetic index of the benchmark shape + a few query batches
(no k-means), so that ncu sees only the query kernels."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine

M, N, P, D, CN, NQ, K, NPROBE = 100000, 1536, 100, 12, 256, 10000, 10, 5
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
rng = np.random.default_rng(0)
ctx = engine.Context(0)
# centroids like k-means leaves them on uniform data: means of ~M/P vectors
coarse = (0.5 + rng.normal(0.0, (1.0 / (12.0 * M / P)) ** 0.5, (P, N))).astype(np.float32)
cbs = rng.random((D, CN, N // D), dtype=np.float32) - np.float32(0.5)
sizes = rng.multinomial(M, np.ones(P) / P)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
codes = rng.integers(0, CN, (M, D)).astype(np.uint8)
ix = engine.Index.create(ctx, coarse, cbs, off, codes)
d_q = ctx.alloc(NQ * N * 4)
ctx.fill_uniform(d_q, NQ * N, 2)
outs = [ctx.alloc(NQ * K * 4) for _ in range(3)] + [ctx.alloc(NQ * 4)]
ix.set_timing(True)
for _ in range(steps):
    ix.query_device(d_q, NQ, K, NPROBE, *outs)
    ctx.sync()
print("phase ms", ix.last_timing())
print("stats (filter, exact, exact candidates, scanned vectors)", ix.last_stats())
