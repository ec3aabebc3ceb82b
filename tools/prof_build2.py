"""Profiling driver: BASELINE.json configs[2] (N=768 D=48 P=1024 C=256) through the sharded builder at world = 1,
on M rows (default 250k) and a few Lloyd rounds, so that an ncu launch list shows where a build round goes.
usage: prof_build2.py [M rounds]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine, sharded

M = int(sys.argv[1]) if len(sys.argv) > 1 else 250_000
rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 5
N, P, D, C = 768, 1024, 48, 256


class Seeds:
    def __init__(self):
        self.rng = np.random.default_rng(1)

    def first(self, n, nb):
        return self.rng.integers(0, n, nb).astype(np.uint32)

    def draws(self, nb, count):
        return self.rng.random((nb, count)).astype(np.float32)


ctx = engine.Context(0)
comm = engine.Comm(ctx, 1, 0)
vs = engine.VectorSet.generate(ctx, M, N, 7)
b = sharded.ShardedDatabaseBuilder(vs, comm, M, Seeds()).with_partitions(P).with_divisions(D).with_clusters(C)
b.max_rounds = rounds
marks = []
ctx.timer_start()
res = b.build(tick=lambda name: marks.append((name, ctx.timer_stop())))
prev = 0.0
for name, t in marks:
    print("%-16s %9.2f ms" % (name, t - prev))
    prev = t
print("launches", ctx.launches)
