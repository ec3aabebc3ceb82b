"""Profiling driver for the host-buffer query path (fdb_index_query): a synthetic index of the benchmark shape,
pinned host queries, the library's own trace (FDB_QUERY_TRACE) for several slice sizes, and the raw pinned
host->device copy time of the same bytes beside it.
usage: prof_e2e.py [reps]"""
import ctypes as C
import os
import sys
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flechasdb_b200 import engine
from flechasdb_b200 import _capi as capi

M, N, P, D, CN, NQ, K, NPROBE = 100000, 1536, 100, 12, 256, 10000, 10, 5
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
rng = np.random.default_rng(0)
ctx = engine.Context(0)
coarse = (0.5 + rng.normal(0.0, (1.0 / (12.0 * M / P)) ** 0.5, (P, N))).astype(np.float32)
cbs = rng.random((D, CN, N // D), dtype=np.float32) - np.float32(0.5)
sizes = rng.multinomial(M, np.ones(P) / P)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
codes = rng.integers(0, CN, (M, D)).astype(np.uint8)
ix = engine.Index.create(ctx, coarse, cbs, off, codes)

tq = torch.empty((NQ, N), dtype=torch.float32, pin_memory=True)
q = tq.numpy()
q[:] = rng.random((NQ, N), dtype=np.float32)
dev = torch.empty((NQ, N), dtype=torch.float32, device="cuda:0")
for _ in range(3):
    dev.copy_(tq, non_blocking=True)
torch.cuda.synchronize()
ts = []
for _ in range(reps):
    t0 = time.perf_counter()
    dev.copy_(tq, non_blocking=True)
    torch.cuda.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
print("raw pinned H2D of %.1f MB: min %.3f ms, median %.3f ms -> %.1f GB/s" % (
    q.nbytes / 1e6, min(ts), sorted(ts)[len(ts) // 2], q.nbytes / 1e6 / min(ts)))

# device-resident reference
d_q = ctx.alloc(NQ * N * 4)
ctx.fill_uniform(d_q, NQ * N, 2)
outs = [ctx.alloc(NQ * K * 4) for _ in range(3)] + [ctx.alloc(NQ * 4)]
for _ in range(3):
    ix.query_device(d_q, NQ, K, NPROBE, *outs)
ctx.sync()
ctx.timer_start()
for _ in range(reps):
    ix.query_device(d_q, NQ, K, NPROBE, *outs)
print("device-resident batch: %.3f ms" % (ctx.timer_stop() / reps))

pin = [torch.empty((NQ, K), dtype=torch.int32, pin_memory=True) for _ in range(2)] + \
      [torch.empty((NQ, K), dtype=torch.float32, pin_memory=True), torch.empty((NQ,), dtype=torch.int32, pin_memory=True)]
o = [t.numpy() for t in pin]
o = [o[0].view(np.uint32), o[1].view(np.uint32), o[2], o[3].view(np.uint32)]


def host_query():
    capi.check(capi.lib().fdb_index_query(ix.h, capi.f32p(q), NQ, K, NPROBE, capi.QUERY_STORED, capi.u32p(o[0]),
                                          capi.u32p(o[1]), capi.f32p(o[2]), capi.u32p(o[3])))


for slice_q in [None] + [int(s) for s in os.environ.get("SLICES", "1000,1250,1667,2000,2500,3334,5000").split(",")]:
    if slice_q is None:
        os.environ.pop("FDB_QUERY_HOST_SLICE", None)
    else:
        os.environ["FDB_QUERY_HOST_SLICE"] = str(slice_q)
    os.environ.pop("FDB_QUERY_TRACE", None)
    for _ in range(3):
        host_query()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        host_query()
        ts.append((time.perf_counter() - t0) * 1e3)
    print("slice %s: min %.3f ms, median %.3f ms" % (slice_q, min(ts), sorted(ts)[len(ts) // 2]), flush=True)
    os.environ["FDB_QUERY_TRACE"] = "1"
    host_query()
    sys.stderr.flush()

# tapered plans (per cent of the batch per slice) x scan kernel of the slices
os.environ.pop("FDB_QUERY_HOST_SLICE", None)
for scan in (None, "query"):
    if scan:
        os.environ["FDB_FILTER_SCAN"] = scan
    else:
        os.environ.pop("FDB_FILTER_SCAN", None)
    for plan in os.environ.get("PLANS", "25,25,25,25;10,30,30,20,10;10,25,25,25,15;15,35,35,15;5,20,25,25,15,10;20,30,30,20;12,22,22,22,22").split(";"):
        os.environ["FDB_QUERY_HOST_PLAN"] = plan
        os.environ.pop("FDB_QUERY_TRACE", None)
        for _ in range(3):
            host_query()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            host_query()
            ts.append((time.perf_counter() - t0) * 1e3)
        print("scan %s plan %s: min %.3f ms, median %.3f ms" % (scan or "default", plan, min(ts), sorted(ts)[len(ts) // 2]), flush=True)
        os.environ["FDB_QUERY_TRACE"] = "1"
        host_query()
        sys.stderr.flush()
os.environ.pop("FDB_QUERY_HOST_PLAN", None)
os.environ.pop("FDB_FILTER_SCAN", None)
