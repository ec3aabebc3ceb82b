"""Multi-GPU host mirror: DatabaseBuilder::build over row shards and Database::query over sharded
code lists, one rank per GPU.

Everything that moves data between GPUs happens inside libflechasdb_b200.so (fdb_comm: NCCL on the
context's stream); this file only sequences the same calls the single-GPU mirror (db.py) makes:

  DatabaseBuilder::build_with_events   src/db/build.rs:78-129
    partition  (src/partitions.rs:115-144): coarse k-means over the row shards, residues row-local
    divide     (src/vector.rs:154-174):     strided views, no copy
    codebooks  (src/db/build.rs:110-118):   the D sub-vector k-means side by side over the row shards
  Database::query                      src/db/stored.rs:315-389 / src/db/build.rs:294-340
    code lists sharded by partition (size-balanced greedy), coarse centroids and codebooks replicated.
"""
import numpy as np

from . import _capi as capi
from .engine import Comm, Index, KMeans, VectorSet, shard_rows   # noqa: F401  (re-exported)


def owned_partitions(sizes, world):
    """size-balanced greedy assignment of partitions (code lists) to ranks: (P,) owner ids"""
    order = np.argsort(-np.asarray(sizes, np.int64), kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.zeros(len(sizes), np.int64)
    for p in order:
        r = int(np.argmin(load))
        owner[p] = r
        load[r] += sizes[p]
    return owner


def shard_offsets(sizes, owner, rank):
    """offsets [P+1] of the index a rank holds: its own lists, the other partitions empty"""
    mine = np.where(np.asarray(owner) == rank, np.asarray(sizes, np.int64), 0)
    return np.concatenate([[0], np.cumsum(mine)]).astype(np.uint64)


class ShardedBuild:
    """What a rank holds after a sharded build: the replicated quantisers (coarse centroids [P][N], codebooks
    [D][C][N/D]) and, for its own rows, the partition and the PQ codes."""

    def __init__(self, vs, ckm, pkm, n_global, lo, hi, stats):
        self.vs, self.ckm, self.pkm = vs, ckm, pkm
        self.n_global, self.lo, self.hi = n_global, lo, hi
        self.stats = stats

    def quantisers(self):
        coarse, part = self.ckm.get()
        cbs, codes = self.pkm.get()
        return coarse[0], cbs, part[0], codes          # codes [D][rows of this rank]

    def close(self):
        for h in (self.pkm, self.ckm):
            h.close()


class ShardedDatabaseBuilder:
    """DatabaseBuilder (src/db/build.rs:23-130) with the rows sharded over the ranks of `comm`.

    vs        engine.VectorSet holding THIS rank's rows [n_global*rank/world, n_global*(rank+1)/world)
    seeds     db.SeedSource-like object; every rank must be given the same draws (same seed)
    """

    def __init__(self, vs, comm, n_global, seeds):
        self.vs, self.comm, self.n_global, self.seeds = vs, comm, n_global, seeds
        self.num_partitions, self.num_divisions, self.num_clusters = 10, 8, 16   # src/db/build.rs:48-50
        self.max_rounds = capi.KMEANS_MAX_ROUNDS

    def with_partitions(self, p):
        self.num_partitions = int(p)
        return self

    def with_divisions(self, d):
        self.num_divisions = int(d)
        return self

    def with_clusters(self, c):
        self.num_clusters = int(c)
        return self

    def build(self, event=lambda e: None, tick=None):
        """tick(name) is called after every phase (the bench reads the device clock there)"""
        tick = tick or (lambda name: None)
        vs, comm, M = self.vs, self.comm, self.n_global
        P, D, Cn = self.num_partitions, self.num_divisions, self.num_clusters
        N = vs.vector_size
        lo, hi = shard_rows(M, comm.world, comm.rank)
        if hi - lo != len(vs):
            raise ValueError("rank %d holds %d rows, its shard is [%d, %d)" % (comm.rank, len(vs), lo, hi))
        if N % D != 0:
            from .db import Error
            raise Error("InvalidArgs", "vector size (%d) is not divisible by %d" % (N, D))
        event(("StartingPartitioning",))
        ckm = KMeans(vs, P)
        picks_c = ckm.seed_run_sharded(comm, M, self.seeds.first(M, 1), self.seeds.draws(1, P - 1))
        tick("coarse_seeding")
        g_c, rounds_c, reas_c = ckm.run_sharded(comm, self.max_rounds)
        tick("coarse_lloyd")
        vs.subtract_assigned(ckm)                      # v_j -= centroid[indices[j]]: row-local
        tick("residues")
        event(("FinishedPartitioning",))
        pkm = KMeans(vs, Cn, col_off=0, dim=N // D, nb=D)
        picks_p = pkm.seed_run_sharded(comm, M, self.seeds.first(M, D), self.seeds.draws(D, Cn - 1))
        tick("pq_seeding")
        g_p, rounds_p, reas_p = pkm.run_sharded(comm, self.max_rounds)
        tick("pq_lloyd")
        stats = dict(picks_coarse=picks_c, picks_pq=picks_p, rounds_coarse=int(rounds_c[0]),
                     rounds_pq=[int(r) for r in rounds_p], gradients_coarse=g_c[0],
                     reassignments=int(reas_c[0]) + int(np.sum(reas_p)))
        return ShardedBuild(vs, ckm, pkm, M, lo, hi, stats)
