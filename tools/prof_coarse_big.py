"""Coarse quantiser with more than 256 centroids (BASELINE.json configs[2] shape: 1M x 768, P = 1024):
device time of Lloyd reassignments, tensor-core column tiles vs the exact fp32 kernel."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine
M, N, P = [int(a) for a in sys.argv[1:4]] if len(sys.argv) >= 4 else (1000000, 768, 1024)
ctx = engine.Context(0)
vs = engine.VectorSet.generate(ctx, M, N, 1)
rng = np.random.default_rng(0)
km = engine.KMeans(vs, P)
km.seed_chosen(rng.choice(M, P, replace=False)[None, :])
for r in range(4):
    ctx.timer_start(); km.update(); tu = ctx.timer_stop()
    ctx.timer_start(); km.reassign(); tr = ctx.timer_stop()
    print("round %d: update %.2f ms  reassign %.2f ms  %s" % (r, tu, tr, km.last_assign_info()))
_, idx = km.get()
print("checksum", int(idx.astype(np.int64).sum()), int((idx[0] * np.arange(1, M + 1, dtype=np.int64) % 1000003).sum()))
