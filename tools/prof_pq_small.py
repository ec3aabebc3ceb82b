"""PQ codebooks with small sub-vectors (configs[2]: 48 divisions of 16 dims, 1M rows): device time per phase."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine
M, N, D, CN = [int(a) for a in sys.argv[1:5]] if len(sys.argv) >= 5 else (1000000, 768, 48, 256)
ctx = engine.Context(0)
vs = engine.VectorSet.generate(ctx, M, N, 1)
rng = np.random.default_rng(0)
km = engine.KMeans(vs, CN, dim=N // D, nb=D)
km.seed_chosen(np.stack([rng.choice(M, CN, replace=False) for _ in range(D)]))
for r in range(4):
    ctx.timer_start(); km.update(); tu = ctx.timer_stop()
    ctx.timer_start(); km.reassign(); tr = ctx.timer_stop()
    print("round %d: update %.2f ms  reassign %.2f ms  %s" % (r, tu, tr, km.last_assign_info()))
