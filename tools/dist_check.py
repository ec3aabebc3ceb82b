"""2+ GPU check of the sharded build and the partition-sharded query (run under torchrun).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/dist_check.py [M N P D C]
Checks (rank 0 prints): picks identical on all ranks; one sharded update step matches the
single-GPU update within 1e-5 relative; assignments bit-exact given the same centroids;
partition-sharded query == unsharded query (mode build).  Then times a sharded build."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from flechasdb_b200 import engine, dist as fd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda:%d" % local))
dev = "cuda:%d" % local
M, N, P, D, CN = [int(a) for a in sys.argv[1:6]] if len(sys.argv) >= 6 else (40000, 256, 32, 4, 64)
SEED = 0xF1EC4A5D0001
ctx = engine.Context(local)
comm = fd.Comm(dist, dev)
lo, hi = fd.shard_rows(M, world, rank)
# this rank's rows: the same counter-based generator, started at the shard's first element
vs = engine.VectorSet.generate(ctx, hi - lo, N, SEED, start=lo * N)
view = lambda ptr, n: fd.device_tensor(ptr, n, dev)
rng = np.random.default_rng(1)
first = rng.integers(0, M, 1)
u = rng.random((1, P - 1)).astype(np.float32)

ckm = engine.KMeans(vs, P)
sk = fd.ShardedKMeans(comm, ckm, lambda li: vs.download(li, 1)[0], M, partial_view=view)
# host-driven seeding (three host round trips per round) and the device-driven one: same picks
torch.cuda.synchronize(); comm.barrier(); t0 = time.perf_counter()
picked_host = sk.seed(first, u)
torch.cuda.synchronize(); t_seed_host = time.perf_counter() - t0
comm.barrier(); t0 = time.perf_counter()
picked = sk.seed_device(first, u)
torch.cuda.synchronize(); t_seed = time.perf_counter() - t0
ok_same = bool((picked == picked_host).all())
allp = comm.all_gather(picked.astype(np.int32))
ok_picks = all((allp[r] == allp[0]).all() for r in range(world)) and ok_same
# one update step, compared with a single-GPU engine over all rows (rank 0 only)
ptr, nfl = ckm.update_partial()
comm.all_reduce_sum_tensor(view(ptr, nfl)); torch.cuda.synchronize()
g = ckm.update_finish()
cent, _ = ckm.get()
ok_update = ok_assign = True
if rank == 0:
    full = engine.VectorSet.generate(ctx, M, N, SEED)
    fkm = engine.KMeans(full, P)
    fkm.seed_chosen(picked.astype(np.uint32))
    g1 = fkm.update()
    c1, _ = fkm.get()
    rel = np.abs(c1 - cent).max() / np.abs(c1).max()
    ok_update = rel < 1e-5 and abs(g1[0] - g[0]) <= 1e-4 * max(g1[0], 1e-30)
    fkm.set_state(cent)
    fkm.reassign()
    full_idx = fkm.get()[1][0]
ckm.reassign()
idx = ckm.get()[1][0]
gi = comm.all_gather(np.pad(idx, (0, (M + world - 1) // world + 1 - len(idx))).astype(np.int32))
if rank == 0:
    stitched = np.concatenate([gi[r][:fd.shard_rows(M, world, r)[1] - fd.shard_rows(M, world, r)[0]] for r in range(world)])
    ok_assign = bool((stitched == full_idx).all())
# full sharded Lloyd loop, timed
c_start, i_start = ckm.get()
comm.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
grads, reas = sk.run(max_rounds=100)
torch.cuda.synchronize(); t_lloyd_host = time.perf_counter() - t0
c_host, i_host = ckm.get()
# the same loop without host round trips (kernels + NCCL on the engine's stream): identical trajectory
ckm.set_state(c_start, i_start)
comm.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
grads_d, reas_d = sk.run_device(max_rounds=100)
torch.cuda.synchronize(); t_lloyd = time.perf_counter() - t0
c_dev, i_dev = ckm.get()
ok_loop = bool((c_dev == c_host).all() and (i_dev == i_host).all() and len(grads) == len(grads_d)
               and all((a == b).all() for a, b in zip(grads, grads_d)) and (reas == reas_d).all())
# partition-sharded query against an index built on rank 0's single-GPU path
ok_query = True
if True:
    rngq = np.random.default_rng(9)
    Pq, Dq, Cq, Mq, Nq = 24, 4, 32, 20000, 64
    coarse = rngq.random((Pq, Nq), dtype=np.float32)
    cbs = rngq.random((Dq, Cq, Nq // Dq), dtype=np.float32) - np.float32(0.5)
    sizes = rngq.multinomial(Mq, np.ones(Pq) / Pq)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = rngq.integers(0, 4, (Mq, Dq)).astype(np.uint8)
    q = rngq.random((200, Nq), dtype=np.float32)
    owner = fd.owned_partitions(sizes, world)
    so, sc = fd.shard_index_arrays(off, codes, owner, rank)
    six = engine.Index.create(ctx, coarse, cbs, so, sc)
    mp_, mv, md, mc = fd.sharded_query(comm, six, q, 10, 6, mode=1)
    if rank == 0:
        fix = engine.Index.create(ctx, coarse, cbs, off, codes)
        wp, wv, wd, wc = fix.query(q, 10, 6, 1)
        ok_query = bool((mc == wc).all() and (mp_ == wp).all() and (mv == wv).all() and (md == wd).all())
if rank == 0:
    print("dist_check world=%d M=%d N=%d P=%d: picks_equal=%s update_close=%s assign_exact=%s query_equal=%s "
          "loop_equal=%s seed_s=%.3f (host-driven %.3f) lloyd_s=%.3f (host-driven %.3f) rounds=%d" % (
              world, M, N, P, ok_picks, ok_update, ok_assign, ok_query, ok_loop, t_seed, t_seed_host, t_lloyd,
              t_lloyd_host, len(grads)), flush=True)
    assert ok_picks and ok_update and ok_assign and ok_query and ok_loop
dist.barrier()
dist.destroy_process_group()
