"""World-size-2 gloo tests of the multi-GPU host logic (flechasdb_b200/dist.py) on CPU."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from flechasdb_b200 import dist as fd
    from fake_engine import FakeKMeans
    from oracle import pyoracle as oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = fd.Comm(dist, "cpu")
        n, N, nb, k = 240, 16, 2, 5
        m = N // nb
        x = oracle.fill_uniform(n * N, 7).reshape(n, N)
        lo, hi = fd.shard_rows(n, world, rank)
        km = FakeKMeans(x[lo:hi], k, dim=m, nb=nb)
        sk = fd.ShardedKMeans(comm, km, lambda li: x[lo + li], n,
                              partial_view=lambda buf, nfl: torch.from_numpy(buf))
        rng = np.random.default_rng(3)
        first = rng.integers(0, n, nb)
        u = rng.random((nb, k - 1)).astype(np.float32)
        picked = sk.seed(first, u)
        grads, reas = sk.run(max_rounds=6)
        cent, idx = km.get()
        gathered = comm.all_gather(np.pad(idx, ((0, 0), (0, n - idx.shape[1]))))
        ret[rank] = dict(picked=picked, grads=np.array(grads), reas=reas, cent=cent,
                         idx=idx, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def _run(world):
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000) + world
    mgr = mp.Manager()
    ret = mgr.dict()
    if world == 1:
        _worker_single(ret)
    else:
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    return dict(ret)


def _worker_single(ret):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from flechasdb_b200 import dist as fd
    from fake_engine import FakeKMeans
    from oracle import pyoracle as oracle
    import torch
    comm = fd.Comm(None)
    n, N, nb, k = 240, 16, 2, 5
    m = N // nb
    x = oracle.fill_uniform(n * N, 7).reshape(n, N)
    km = FakeKMeans(x, k, dim=m, nb=nb)
    sk = fd.ShardedKMeans(comm, km, lambda li: x[li], n,
                          partial_view=lambda buf, nfl: torch.from_numpy(buf))
    rng = np.random.default_rng(3)
    first = rng.integers(0, n, nb)
    u = rng.random((nb, k - 1)).astype(np.float32)
    picked = sk.seed(first, u)
    grads, reas = sk.run(max_rounds=6)
    cent, idx = km.get()
    ret[0] = dict(picked=picked, grads=np.array(grads), reas=reas, cent=cent, idx=idx, lo=0, hi=n)


@pytest.mark.timeout(300)
def test_sharded_kmeans_world2_matches_world1():
    one = _run(1)[0]
    two = _run(2)
    assert set(two) == {0, 1}
    # every rank made the same picks, and they are the single-process picks
    assert (two[0]["picked"] == two[1]["picked"]).all()
    assert (two[0]["picked"] == one["picked"]).all()
    # replicated centroids agree across ranks bit for bit, and with world=1 within summation order
    assert (two[0]["cent"] == two[1]["cent"]).all()
    assert np.allclose(two[0]["cent"], one["cent"], rtol=1e-5, atol=1e-6)
    assert np.allclose(two[0]["grads"], one["grads"], rtol=1e-3, atol=1e-6)
    # assignments of the shards, stitched together, equal the single-process assignments
    idx = np.concatenate([two[0]["idx"], two[1]["idx"]], axis=1)
    assert (idx == one["idx"]).mean() > 0.99


def test_split_sample_and_ownership():
    from flechasdb_b200 import dist as fd
    assert [fd.shard_rows(10, 3, r) for r in range(3)] == [(0, 3), (3, 6), (6, 10)]
    assert fd.owner_of(5, 10, 3) == (1, 2)
    t = np.array([1.0, 0.0, 3.0], np.float32)
    assert fd.split_sample(0.0, t)[0] == 0
    r, v = fd.split_sample(0.5, t)          # draw 2.0 -> third shard, 1.0 into it
    assert r == 2 and abs(v - 1.0) < 1e-6
    r, v = fd.split_sample(0.999999, t)
    assert r == 2
    with pytest.raises(ArithmeticError):
        fd.split_sample(0.3, np.zeros(3, np.float32))


def test_partition_sharding_and_merge_equal_unsharded(oracle):
    from flechasdb_b200 import dist as fd
    rng = np.random.default_rng(5)
    N, P, D, Cn, M, k, nprobe = 32, 9, 4, 16, 700, 6, 4
    coarse = oracle.fill_uniform(P * N, 1).reshape(P, N)
    cbs = (oracle.fill_uniform(D * Cn * (N // D), 2) - np.float32(0.5)).reshape(D, Cn, N // D)
    sizes = rng.multinomial(M, np.ones(P) / P)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = rng.integers(0, 3, (M, D)).astype(np.uint32)     # few codes: many exact ties
    q = oracle.fill_uniform(20 * N, 3).reshape(20, N)
    full = oracle.QueryIndex(coarse, cbs, off, codes)
    rc, wp, wv, wd, wc = full.query(q, k, nprobe, 1)
    world = 3
    owner = fd.owned_partitions(sizes, world)
    assert sorted(np.bincount(owner, weights=sizes, minlength=world))[0] > 0.6 * M / world
    res = []
    for r in range(world):
        so, sc = fd.shard_index_arrays(off, codes, owner, r)
        ix = oracle.QueryIndex(coarse, cbs, so, sc)
        res.append(ix.query(q, k, nprobe, 1)[1:])
    probes = np.stack([full.probe(q[i], nprobe, 1)[1] for i in range(len(q))])
    gp, gv, gd, gc = (np.stack([r[i] for r in res]) for i in range(4))
    mp_, mv, md, mc = fd.merge_topk(gp, gv, gd, gc, probes, k)
    assert (mc == wc).all() and (mp_ == wp).all() and (mv == wv).all() and (md == wd).all()


def test_partition_ownership_is_balanced_and_consistent():
    """sharded.owned_partitions / shard_offsets (code lists sharded by partition): every partition
    has exactly one owner, the greedy assignment is size balanced, and a rank's offsets hold exactly its own lists."""
    import numpy as np
    from flechasdb_b200 import sharded
    rng = np.random.default_rng(3)
    for world in (1, 2, 4, 8):
        sizes = rng.multinomial(1_000_000, np.ones(4096) / 4096)
        sizes[:5] = [0, 0, 50_000, 1, 0]                       # empty lists and one very long one
        owner = sharded.owned_partitions(sizes, world)
        assert owner.shape == sizes.shape and owner.min() >= 0 and owner.max() < world
        load = np.bincount(owner, weights=sizes, minlength=world)
        assert load.sum() == sizes.sum()
        assert load.max() - load.min() <= sizes.max()          # greedy on descending sizes: within one list of each other
        total = 0
        for r in range(world):
            off = sharded.shard_offsets(sizes, owner, r)
            mine = np.diff(off.astype(np.int64))
            assert (mine[owner == r] == sizes[owner == r]).all() and (mine[owner != r] == 0).all()
            total += int(off[-1])
        assert total == int(sizes.sum())
    assert (sharded.owned_partitions(sizes, 1) == 0).all()


def test_shard_rows_cover_the_rows_once():
    from flechasdb_b200 import engine
    for n in (1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            cuts = [engine.shard_rows(n, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            assert all(lo == n * r // world for r, (lo, _) in enumerate(cuts))     # fdb_kmeans_seed_run_sharded's rule
