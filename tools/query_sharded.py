"""BASELINE.json configs[4] (scaled by the first argument): query sweep over NPROBE with the code lists sharded
over the ranks and the cross-GPU top-k merge on the device.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29561 \
        tools/query_sharded.py [M P NQ]        (default 100M vectors, 16384 lists, 10000 queries; N 96, D 12, C 256)
The index is synthesised (uniform coarse centroids and code vectors, uniform u8 codes, multinomial list sizes);
every rank keeps the lists it owns (size-balanced greedy).  Rank 0 also holds the whole index and checks the
merged result of the first 512 queries against the unsharded query.  One JSON line per NPROBE on rank 0."""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from flechasdb_b200 import engine, dist as fd

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
dev = "cuda:%d" % local
a = sys.argv[1:]
M, P, NQ = (int(a[0]), int(a[1]), int(a[2])) if len(a) >= 3 else (100_000_000, 16384, 10000)
N, D, CN, K = 96, 12, 256, 10
ctx = engine.Context(local)
comm = fd.Comm(dist, dev)
rng = np.random.default_rng(11)
coarse = rng.random((P, N), dtype=np.float32)
cbs = rng.random((D, CN, N // D), dtype=np.float32) - np.float32(0.5)
sizes = rng.multinomial(M, np.ones(P) / P)
off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
owner = fd.owned_partitions(sizes, world)


def codes_of(parts):
    """the lists of `parts` (ascending), generated list by list so that every rank makes the same bytes"""
    out = np.empty((int(sizes[parts].sum()), D), np.uint8)
    o = 0
    for p in parts:
        n = int(sizes[p])
        out[o:o + n] = np.random.default_rng(1_000_003 + int(p)).integers(0, 256, (n, D), dtype=np.uint8)
        o += n
    return out


mine = np.flatnonzero(owner == rank)
my_sizes = np.where(owner == rank, sizes, 0)
my_off = np.concatenate([[0], np.cumsum(my_sizes)]).astype(np.uint64)
t0 = time.perf_counter()
six = engine.Index.create(ctx, coarse, cbs, my_off, codes_of(mine))
t_index = time.perf_counter() - t0
d_q = ctx.alloc(NQ * N * 4)
ctx.fill_uniform(d_q, NQ * N, 0xF1EC4A5D0002)          # the same batch on every rank
NCHECK = min(512, NQ)
full = engine.Index.create(ctx, coarse, cbs, off, codes_of(np.arange(P))) if rank == 0 and world > 1 else None
wouts = [ctx.alloc(NCHECK * K * 4) for _ in range(3)] + [ctx.alloc(NCHECK * 4)]

for nprobe in [int(x) for x in os.environ.get("NPROBES", "8,16,32,64,128").split(",")]:
    for _ in range(2):
        res = fd.sharded_query_device(comm, six, d_q, NQ, K, nprobe)
    torch.cuda.synchronize(); comm.barrier()
    times = []
    for _ in range(3):
        ctx.flush_l2()
        torch.cuda.synchronize(); comm.barrier(); t0 = time.perf_counter()
        res = fd.sharded_query_device(comm, six, d_q, NQ, K, nprobe)
        ctx.sync(); torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    stats = six.last_stats()
    sc = torch.tensor([stats[3], stats[1]], dtype=torch.int64, device=dev)
    dist.all_reduce(sc)
    ok = None
    if full is not None:
        full.query_device(d_q, NCHECK, K, nprobe, *wouts, mode=1)
        ctx.sync()
        got = [x[:NCHECK].cpu().numpy() for x in res]
        wantn = [fd.device_tensor(getattr(ptr, "value", ptr), NCHECK * K if i < 3 else NCHECK, dev, "<f4" if i == 2 else "<i4").cpu().numpy()
                 for i, ptr in enumerate(wouts)]
        cnt = wantn[3]
        ok = bool((got[3] == cnt).all())
        # partition ids agree; vector indices are positions inside a list (the same in both layouts)
        for i in range(3):
            g, w = got[i].reshape(NCHECK, K), wantn[i].reshape(NCHECK, K)
            ok = ok and all((g[q, :cnt[q]] == w[q, :cnt[q]]).all() for q in range(NCHECK))
    if rank == 0:
        sec = float(t[0])
        print(json.dumps({
            "workload": "configs[4]: M=%d N=%d D=%d P=%d C=%d, %d queries, k=%d, nprobe=%d, code lists sharded x%d"
                        % (M, N, D, P, CN, NQ, K, nprobe, world),
            "n_gpus": world, "nprobe": nprobe, "ms_per_batch": sec * 1e3, "queries_per_s": NQ / sec,
            "scanned_vectors_all_ranks": int(sc[0]), "scan_gbs_algorithmic_all_ranks": int(sc[0]) * D / sec / 1e9,
            "exact_pipeline_queries_all_ranks": int(sc[1]), "equals_unsharded_first_%d" % NCHECK: ok,
            "index_upload_sec": round(t_index, 2)}), flush=True)
dist.barrier()
dist.destroy_process_group()
