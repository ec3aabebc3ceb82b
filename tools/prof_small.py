"""Per-call latency of small query batches (device-resident queries): phases and stats."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine

M, N, P, D, CN, K, NPROBE = 100000, 1536, 100, 12, 256, 10, 5
rng = np.random.default_rng(0)
ctx = engine.Context(0)
coarse = (0.5 + rng.normal(0.0, (1.0 / (12.0 * M / P)) ** 0.5, (P, N))).astype(np.float32)
cbs = rng.random((D, CN, N // D), dtype=np.float32) - np.float32(0.5)
sizes = rng.multinomial(M, np.ones(P) / P)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
codes = rng.integers(0, CN, (M, D)).astype(np.uint8)
ix = engine.Index.create(ctx, coarse, cbs, off, codes)
for NQ in (1, 64, 1024, 2048, 4096, 10000):
    d_q = ctx.alloc(NQ * N * 4)
    ctx.fill_uniform(d_q, NQ * N, 2)
    outs = [ctx.alloc(NQ * K * 4) for _ in range(3)] + [ctx.alloc(NQ * 4)]
    ix.set_timing(True)
    for _ in range(3):
        ix.query_device(d_q, NQ, K, NPROBE, *outs)
    ph, _ = ix.last_timing()
    ix.set_timing(False)
    ctx.sync()
    t0 = time.perf_counter()
    for _ in range(10):
        ix.query_device(d_q, NQ, K, NPROBE, *outs)
    ctx.sync()
    dt = (time.perf_counter() - t0) / 10
    print("nq %6d  wall %.3f ms  phases %s  stats %s" % (NQ, dt * 1e3, np.round(ph, 3), ix.last_stats()))
