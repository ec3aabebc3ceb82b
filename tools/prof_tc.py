"""Profiling driver for the tensor-core assignment: a few steady-state Lloyd rounds of the coarse
quantiser (100k x 1536, k=100) and of the PQ codebooks (12 x 100k x 128, k=256)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine
M, N, P, D, CN = 100000, 1536, 100, 12, 256
ctx = engine.Context(0)
vs = engine.VectorSet.generate(ctx, M, N, 1)
rng = np.random.default_rng(0)
ckm = engine.KMeans(vs, P)
ckm.seed_chosen(rng.choice(M, P, replace=False)[None, :])
for r in range(3):
    ckm.update(); ckm.reassign()
vs.subtract_assigned(ckm)
pkm = engine.KMeans(vs, CN, dim=N // D, nb=D)
pkm.seed_chosen(np.stack([rng.choice(M, CN, replace=False) for _ in range(D)]))
for r in range(3):
    pkm.update(); ctx.timer_start(); pkm.reassign(); print("pq reassign ms", ctx.timer_stop())
