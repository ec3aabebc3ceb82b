"""configs[2] phases on ONE GPU for a shard of `rows` rows (N=768, P=1024, D=48, C=256) through the sharded
builder at world 1 (the library's comm without NCCL): where a rank's time goes at 1/2/4/8 GPUs without the
collectives.  usage: prof_cfg2.py [rows [max_rounds [builds]]]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flechasdb_b200 import engine, sharded

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
max_rounds = int(sys.argv[2]) if len(sys.argv) > 2 else 100
builds = int(sys.argv[3]) if len(sys.argv) > 3 else 2
N, P, D, C = 768, 1024, 48, 256


class Seeds:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def first(self, n, nb):
        return self.rng.integers(0, n, nb).astype(np.uint32)

    def draws(self, nb, count):
        return (self.rng.integers(0, 1 << 23, (nb, count), dtype=np.uint32).astype(np.float32) * np.float32(2.0 ** -23))


ctx = engine.Context(0)
comm = engine.Comm(ctx, 1, 0, None)
for it in range(builds):
    vs = engine.VectorSet.generate(ctx, rows, N, 0xF1EC4A5D0003)
    marks = []
    l0 = ctx.launches
    ctx.timer_start()
    b = sharded.ShardedDatabaseBuilder(vs, comm, rows, Seeds(3))
    b.max_rounds = max_rounds
    sb = b.with_partitions(P).with_divisions(D).with_clusters(C).build(tick=lambda name: marks.append((name, ctx.timer_stop())))
    t = [m for _, m in marks]
    ph = np.diff([0.0] + t)
    print("rows %d build %d: total %.1f ms; " % (rows, it, t[-1]) + ", ".join("%s %.1f" % (n, v) for (n, _), v in zip(marks, ph)) +
          "; launches %d, rounds coarse %d pq %d" % (ctx.launches - l0, sb.stats["rounds_coarse"], max(sb.stats["rounds_pq"])))
    sb.close()
    vs.close()
