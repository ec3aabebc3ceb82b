"""Per-round wall/device timing of Lloyd rounds (coarse and PQ) on the benchmark shape."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine
M, N, P, D, CN = 100000, 1536, 100, 12, 256
ctx = engine.Context(0)
vs = engine.VectorSet.generate(ctx, M, N, 1)
rng = np.random.default_rng(0)
def rounds(km, name, n=6):
    for r in range(n):
        ctx.sync(); t0 = time.perf_counter(); ctx.timer_start()
        km.update(); t1 = time.perf_counter()
        km.reassign(); ms = ctx.timer_stop(); t2 = time.perf_counter()
        print("%s round %d: update %.2f ms  reassign %.2f ms  (device total %.2f ms)" % (name, r, (t1-t0)*1e3, (t2-t1)*1e3, ms))
ckm = engine.KMeans(vs, P)
ckm.seed_run([5], rng.random((1, P - 1), dtype=np.float32))
rounds(ckm, "coarse")
vs.subtract_assigned(ckm)
pkm = engine.KMeans(vs, CN, dim=N // D, nb=D)
pkm.seed_run(rng.integers(0, M, D), rng.random((D, CN - 1), dtype=np.float32))
rounds(pkm, "pq")
ctx.timer_start(); pkm.run(max_rounds=5); print("pq run(5) device ms", ctx.timer_stop())
