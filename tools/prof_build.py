"""Profiling driver: seeding rounds + Lloyd rounds of the coarse quantiser and of the
PQ codebooks on the benchmark shape, so that ncu sees each build kernel a few times."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine

M, N, P, D, CN = 100000, 1536, 100, 12, 256
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
ctx = engine.Context(0)
vs = engine.VectorSet.generate(ctx, M, N, 1)
rng = np.random.default_rng(0)
ckm = engine.KMeans(vs, P)
ckm.seed_run([5], rng.random((1, P - 1), dtype=np.float32))
ctx.timer_start(); g = ckm.run(max_rounds=rounds); print("coarse", rounds, "rounds ms", ctx.timer_stop())
vs.subtract_assigned(ckm)
pkm = engine.KMeans(vs, CN, dim=N // D, nb=D)
ctx.timer_start()
pkm.seed_run(rng.integers(0, M, D), rng.random((D, CN - 1), dtype=np.float32))
print("pq seeding ms", ctx.timer_stop())
ctx.timer_start(); g = pkm.run(max_rounds=rounds); print("pq", rounds, "rounds ms", ctx.timer_stop())
ix = engine.Index.from_build(ctx, ckm, pkm)
print("launches", ctx.launches)
