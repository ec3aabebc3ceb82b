"""2+ GPU check of the multi-GPU entry points of the C ABI (fdb_comm: NCCL inside the library), against the
CPU oracle.  Run under torchrun (torch.distributed/gloo only carries the NCCL id and the gathered samples):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29512 tools/dist_check2.py
Rank 0 prints one line of booleans:
  seed_state   after fdb_kmeans_seed_run_sharded: centroids / assignments == oracle k-means++ with the same picks
  lloyd_close  after fdb_kmeans_run_sharded: centroids within 1e-4 relative of the oracle's Lloyd loop
  assign_exact assignments == oracle reassignment given the GPU's centroids
  query_build / query_stored  fdb_index_query_sharded == oracle query on the whole index (ids, distances, counts),
               on uniform data and on data with many exactly tied distances (few distinct codes)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from flechasdb_b200 import engine, sharded, _capi as capi
from oracle import pyoracle as oracle

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    dist.init_process_group("gloo")


def exchange(ident):
    box = [ident]
    dist.broadcast_object_list(box, src=0)
    return box[0]


def gather(obj):
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


ctx = engine.Context(local)
comm = engine.Comm(ctx, world, rank, exchange if world > 1 else None)
SEED = 0xF1EC4A5D0001
res = {}
# ---- build: two problems side by side over row shards --------------------------------------------------
n, N, nb, k, rounds = 20000, 128, 2, 48, 4
m = N // nb
lo, hi = engine.shard_rows(n, world, rank)
vs = engine.VectorSet.generate(ctx, hi - lo, N, SEED, start=lo * N)
km = engine.KMeans(vs, k, dim=m, nb=nb)
rng = np.random.default_rng(1)
picked = km.seed_run_sharded(comm, n, rng.integers(0, n, nb), rng.random((nb, k - 1)).astype(np.float32))
c0, i0 = km.get()
grads, nrounds, reas = km.run_sharded(comm, rounds)
c1, i1 = km.get()
parts = gather((i0, i1, picked))
if rank == 0:
    x = oracle.fill_uniform(n * N, SEED).reshape(n, N)
    gi0 = np.concatenate([p[0] for p in parts], axis=1)
    gi1 = np.concatenate([p[1] for p in parts], axis=1)
    res["picks_same_on_all_ranks"] = all((p[2] == picked).all() for p in parts)
    ok_seed = ok_lloyd = ok_assign = True
    for b in range(nb):
        rc, oc0, oi0, _, _ = oracle.kmeans_init(x, k, int(picked[b, 0]), chosen=picked[b, 1:], off=b * m, dim=m)
        ok_seed &= rc == 0 and bool((oc0 == c0[b]).all()) and bool((oi0 == gi0[b]).all())
        rc, oc, oi, og, _ = oracle.kmeans_lloyd(x, k, oc0, oi0, max_rounds=rounds, off=b * m, dim=m)
        ok_lloyd &= rc == 0 and float(np.abs(oc - c1[b]).max() / np.abs(oc).max()) <= 1e-4
        rc, want = oracle.kmeans_reassign(x, k, c1[b], off=b * m, dim=m)
        ok_assign &= rc == 0 and bool((want == gi1[b]).all())
    res.update(seed_state=ok_seed, lloyd_close=ok_lloyd, assign_exact=ok_assign)
km.close()
vs.close()

# ---- query: code lists sharded by partition, both semantics, with and without ties ---------------------
def check_query(dup, tag):
    Nq, P, D, Cn, M, kq, nprobe, nq = 64, 40, 4, 64, 30000, 10, 6, 300
    r = np.random.default_rng(4)
    s = Nq // D
    coarse = oracle.fill_uniform(P * Nq, SEED + 5).reshape(P, Nq)
    cbs = (oracle.fill_uniform(D * Cn * s, SEED + 6) - np.float32(0.5)).reshape(D, Cn, s)
    sizes = r.multinomial(M, np.ones(P) / P)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = r.integers(0, 3 if dup else Cn, (M, D)).astype(np.uint32)
    owner = sharded.owned_partitions(sizes, world)
    soff = sharded.shard_offsets(sizes, owner, rank)
    mine = [codes[int(off[p]):int(off[p + 1])] for p in range(P) if owner[p] == rank]
    scodes = np.concatenate(mine) if mine else np.zeros((0, D), np.uint32)
    ix = engine.Index.create(ctx, coarse, cbs, soff, scodes.astype(np.uint8))
    q = oracle.fill_uniform(nq * Nq, SEED + 77).reshape(nq, Nq)
    d_q = ctx.alloc(q.nbytes)
    ctx.upload(d_q, q)
    outs = [ctx.alloc(nq * kq * 4) for _ in range(3)] + [ctx.alloc(nq * 4)]
    oix = oracle.QueryIndex(coarse, cbs, off, codes)
    for mode, name in ((capi.QUERY_BUILD, "build"), (capi.QUERY_STORED, "stored")):
        ix.query_sharded(comm, d_q, nq, kq, nprobe, *outs, mode=mode)
        got = [ctx.download(outs[0], (nq, kq), np.uint32), ctx.download(outs[1], (nq, kq), np.uint32),
               ctx.download(outs[2], (nq, kq), np.float32), ctx.download(outs[3], (nq,), np.uint32)]
        rc, wp, wv, wd, wc = oix.query(q, kq, nprobe, mode)
        ok = rc == 0 and bool((got[3] == wc).all())
        for qi in range(nq):
            c = int(wc[qi])
            ok &= bool((got[0][qi, :c] == wp[qi, :c]).all() and (got[1][qi, :c] == wv[qi, :c]).all()
                       and (got[2][qi, :c] == wd[qi, :c]).all())
        res["query_%s_%s" % (name, tag)] = ok
        res["ties_%s_%s" % (name, tag)] = ix.last_sharded_ties()
    ix.close()
    for h in [d_q] + outs:
        ctx.free(h)


check_query(False, "uniform")
check_query(True, "tied")
res["collectives"] = comm.collectives
comm.close()
ctx.close()
if rank == 0:
    print("dist_check2 world=%d " % world + " ".join("%s=%s" % kv for kv in sorted(res.items())), flush=True)
    bad = [k_ for k_, v in res.items() if v is False]
    print("ALL_OK" if not bad else "FAILED: %s" % bad, flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
