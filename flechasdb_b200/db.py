"""Host-side mirror of the reference's database API for the IVF-PQ path, on top of the C ABI.

Mirrors (names, argument meaning, event order, error behaviour):
  DatabaseBuilder::new / with_partitions / with_divisions / with_clusters / build[_with_events]
      src/db/build.rs:23-130
  build::Database::query[_with_events]           src/db/build.rs:294-340
  stored::Database::query[_with_events]          src/db/stored.rs:315-389  (mode="stored")
  ClusterEvent / BuildEvent / QueryEvent         src/kmeans.rs:72-88, src/db/build.rs:134-153,487-500

All arithmetic happens in libflechasdb_b200.so; this file only sequences calls and
keeps the host-side pieces of the reference (vector ids, the RNG, event callbacks).
"""
import uuid
from dataclasses import dataclass

import numpy as np

from . import _capi as capi
from .engine import Context, Index, KMeans, VectorSet


class Error(Exception):
    """flechasdb::error::Error (src/error.rs:5-18); .kind is the variant name."""

    def __init__(self, kind, message):
        super().__init__(message)
        self.kind = kind


def _invalid_args(msg):
    return Error("InvalidArgs", msg)


@dataclass
class QueryResult:
    """src/db/stored.rs:601-612 / src/db/build.rs:577-587"""
    partition_index: int
    vector_id: uuid.UUID
    vector_index: int
    squared_distance: float
    db: object = None        # the stored database the result came from (src/db/stored.rs:603)

    def get_attribute(self, key):
        """QueryResult::get_attribute (src/db/stored.rs:621-634): loads only the attributes log of the result's
        partition on first use"""
        if self.db is None:
            raise Error("InvalidContext", "the result does not belong to a stored database")
        return self.db.get_attribute_in_partition(self.partition_index, self.vector_id, key)


class SeedSource:
    """The draws the reference takes from rand::thread_rng() (src/kmeans.rs:148,172,202):
    gen_range(0..n) for the first centre and one UniformFloat<f32> draw per further centre
    ((u32 >> 9) * 2^-23, rand 0.8.5).  Seedable, so builds are reproducible."""

    def __init__(self, seed=None):
        self.rng = np.random.default_rng(seed)

    def first(self, n, nb):
        return self.rng.integers(0, n, nb).astype(np.uint32)

    def draws(self, nb, count):
        bits = self.rng.integers(0, 1 << 23, (nb, count), dtype=np.uint32)
        return bits.astype(np.float32) * np.float32(2.0 ** -23)


def _replay_cluster_events(event, wrap, grads, reassigns):
    """ClusterEvent order of cluster_with_events (src/kmeans.rs:121-137) for one problem."""
    event(wrap(("StartingCentroidInitialization",)))
    event(wrap(("FinishedCentroidInitialization",)))
    for r, g in enumerate(grads):
        event(wrap(("StartingCentroidUpdate", r)))
        event(wrap(("FinishedCentroidUpdate", r, float(g))))
        if r < reassigns:
            event(wrap(("StartingCentroidReassignment", r)))
            event(wrap(("FinishedCentroidReassignment", r)))


class DatabaseBuilder:
    """src/db/build.rs:23-130.  `vs` is a float32 (M, N) array (BlockVectorSet) or an
    engine.VectorSet already resident in HBM; it is consumed (the residues reuse its
    buffer, src/partitions.rs:17-22,120)."""

    def __init__(self, vs, ctx=None, seeds=None, exact_sampler=False, profile=None, live_events=False):
        self.vs = vs
        # live_events: every k-means is driven round by round from the host (fdb_kmeans_update / _reassign), so each
        # ClusterEvent fires when its phase starts / ends like the reference's (src/kmeans.rs:121-137) and callers that
        # time the phases between events (src/main.rs:53-94) see real durations; the divisions then run one after the
        # other like src/db/build.rs:110-118.  Default: whole loops on the device, events replayed afterwards.
        self.live_events = live_events
        self.profile = profile   # optional dict: phase name -> seconds (adds a sync per phase)
        self.ctx = ctx
        self.seeds = seeds if seeds is not None else SeedSource()
        self.exact_sampler = exact_sampler
        self.num_partitions = 10   # defaults: src/db/build.rs:48-50
        self.num_divisions = 8
        self.num_clusters = 16

    def with_partitions(self, p):
        if p <= 0:
            raise ValueError("NonZeroUsize")
        self.num_partitions = int(p)
        return self

    def with_divisions(self, d):
        if d <= 0:
            raise ValueError("NonZeroUsize")
        self.num_divisions = int(d)
        return self

    def with_clusters(self, c):
        if c <= 0:
            raise ValueError("NonZeroUsize")
        self.num_clusters = int(c)
        return self

    def build(self):
        return self.build_with_events(lambda e: None)

    def build_with_events(self, event):
        own_ctx = self.ctx is None
        ctx = self.ctx if self.ctx is not None else Context(0)
        self._made = []     # handles created by this build: closed again if it fails
        try:
            return self._build(ctx, own_ctx, event)
        except BaseException as e:
            for h in reversed(self._made):
                h.close()
            if own_ctx:
                ctx.close()
            if isinstance(e, capi.FdbError) and e.code == capi.ERR_INVALID_ARGS:
                raise _invalid_args(e.message) from e
            raise

    def _tick(self, ctx, name, t0):
        if self.profile is not None:
            import time
            ctx.sync()
            t1 = time.perf_counter()
            self.profile[name] = self.profile.get(name, 0.0) + (t1 - t0)
            return t1
        return t0

    def _cluster_live(self, km, first, u01, event):
        """cluster_with_events (src/kmeans.rs:104-139), one problem, events fired as the phases happen"""
        event(("ClusterEvent", ("StartingCentroidInitialization",)))
        km.seed_run(first, u01, self.exact_sampler)
        event(("ClusterEvent", ("FinishedCentroidInitialization",)))
        grads, reas = [], 0
        for r in range(capi.KMEANS_MAX_ROUNDS):
            event(("ClusterEvent", ("StartingCentroidUpdate", r)))
            g = float(km.update()[0])
            grads.append(g)
            event(("ClusterEvent", ("FinishedCentroidUpdate", r, g)))
            if g < capi.KMEANS_EPSILON:
                break
            event(("ClusterEvent", ("StartingCentroidReassignment", r)))
            km.reassign()
            reas += 1
            event(("ClusterEvent", ("FinishedCentroidReassignment", r)))
        return grads, reas

    def _build(self, ctx, own_ctx, event):
        import time
        t = time.perf_counter()
        P, D, Cn = self.num_partitions, self.num_divisions, self.num_clusters
        vs = self.vs if isinstance(self.vs, VectorSet) else VectorSet.upload(ctx, self.vs)
        if vs is not self.vs:
            self._made.append(vs)
        M, N = len(vs), vs.vector_size
        # assigns IDs to vectors: Uuid::new_v4() per vector (src/db/build.rs:86-91)
        event(("StartingIdAssignment",))
        raw = np.frombuffer(np.random.bytes(16 * M), np.uint8).reshape(M, 16).copy()
        raw[:, 6] = (raw[:, 6] & 0x0F) | 0x40   # version 4
        raw[:, 8] = (raw[:, 8] & 0x3F) | 0x80   # RFC 4122 variant
        event(("FinishedIdAssignment",))
        t = self._tick(ctx, "upload_and_ids", t)
        # partitions all the data (src/db/build.rs:93-98 -> src/partitions.rs:119-143)
        event(("StartingPartitioning",))
        ckm = KMeans(vs, P)
        self._made.append(ckm)
        if self.live_events:
            self._cluster_live(ckm, self.seeds.first(M, 1), self.seeds.draws(1, P - 1), event)
        else:
            ckm.seed_run(self.seeds.first(M, 1), self.seeds.draws(1, P - 1), self.exact_sampler)
            t = self._tick(ctx, "coarse_seeding", t)
            grads, _, reas = ckm.run()
            t = self._tick(ctx, "coarse_lloyd", t)
            _replay_cluster_events(event, lambda e: ("ClusterEvent", e), grads[0], int(reas[0]))
        vs.subtract_assigned(ckm)
        event(("FinishedPartitioning",))
        t = self._tick(ctx, "residues", t)
        # divides residual vectors (src/db/build.rs:100-105): strided views, no copy
        event(("StartingSubvectorDivision",))
        if N % D != 0:
            raise _invalid_args("vector size (%d) is not divisible by %d" % (N, D))
        event(("FinishedSubvectorDivision",))
        # builds codebooks for residues (src/db/build.rs:110-118): all divisions side by side
        pkm = KMeans(vs, Cn, col_off=0, dim=N // D, nb=D)
        self._made.append(pkm)
        first_p, u_p = self.seeds.first(M, D), self.seeds.draws(D, Cn - 1)
        if self.live_events:
            # one division after the other, each on its own strided view; the finished codebooks and codes are
            # then handed to the batched handle (fdb_kmeans_set_state) that the index is built from
            s = N // D
            cents = np.zeros((D, Cn, s), np.float32)
            idx = np.zeros((D, M), np.uint32)
            for di in range(D):
                event(("StartingQuantization", di))
                one = KMeans(vs, Cn, col_off=di * s, dim=s, nb=1)
                try:
                    self._cluster_live(one, first_p[di:di + 1], u_p[di:di + 1], event)
                    c, i = one.get()
                    cents[di], idx[di] = c[0], i[0]
                finally:
                    one.close()
                event(("FinishedQuantization", di))
            pkm.set_state(cents, idx)
        else:
            pkm.seed_run(first_p, u_p, self.exact_sampler)
            t = self._tick(ctx, "pq_seeding", t)
            grads, _, reas = pkm.run()
            t = self._tick(ctx, "pq_lloyd", t)
            for di in range(D):
                event(("StartingQuantization", di))
                _replay_cluster_events(event, lambda e: ("ClusterEvent", e), grads[di], int(reas[di]))
                event(("FinishedQuantization", di))
        index = Index.from_build(ctx, ckm, pkm)
        t = self._tick(ctx, "events_and_index", t)
        return Database(ctx, own_ctx, vs, ckm, pkm, index, raw, P, D, Cn)


class Database:
    """build::Database (src/db/build.rs:156-176) with its codes regrouped by partition
    on the device (Partition::new, :446-482) so that queries never filter all M indices."""

    def __init__(self, ctx, own_ctx, vs, ckm, pkm, index, id_bytes, P, D, Cn):
        self.ctx, self._own_ctx = ctx, own_ctx
        self.vs, self.ckm, self.pkm, self.index = vs, ckm, pkm, index
        self._id_bytes = id_bytes
        self._num_partitions, self._num_divisions, self._num_clusters = P, D, Cn
        self._order = None
        self._offsets = None
        self._attribute_table = {}      # vector id (16 bytes) -> {name: str | int}

    def num_vectors(self):
        return self.index.num_vectors

    def vector_size(self):
        return self.index.N

    def num_partitions(self):
        return self._num_partitions

    def num_divisions(self):
        return self._num_divisions

    def subvector_size(self):
        return self.index.N // self._num_divisions

    def num_clusters(self):
        return self._num_clusters

    def vector_ids(self):
        return (uuid.UUID(bytes=bytes(b)) for b in self._id_bytes)

    # ---- attributes (src/db/build.rs:225-285): host-side hash maps, never on the device ----
    def get_attribute(self, vector_id, key):
        """Database::get_attribute (src/db/build.rs:228-245): the value (str or int) or None; fails with
        InvalidArgs when no vector has this id"""
        idb = vector_id.bytes if isinstance(vector_id, uuid.UUID) else bytes(vector_id)
        if idb not in self._attribute_table:
            # (the reference looks the id up in attribute_table only: a vector without attributes is "no such id")
            raise _invalid_args("no such vector ID: %s" % uuid.UUID(bytes=idb))
        return self._attribute_table[idb].get(key)

    def set_attribute_at(self, i, attribute):
        """Database::set_attribute_at (src/db/build.rs:252-285): attribute = (key, value), value a str or a u64;
        replaces an existing value; InvalidArgs when i is out of bounds"""
        if not 0 <= int(i) < len(self._id_bytes):
            raise _invalid_args("vector index out of bounds: %d" % i)
        key, value = attribute
        if not isinstance(value, str):
            value = int(value)
            if not 0 <= value < (1 << 64):
                raise _invalid_args("attribute value does not fit a u64: %d" % value)
        self._attribute_table.setdefault(bytes(self._id_bytes[int(i)]), {})[str(key)] = value

    def _layout(self):
        if self._order is None:
            self._offsets, self._order, _ = self.index.layout(order=True)
        return self._offsets, self._order

    def _vector_id(self, part, vidx):
        off, order = self._layout()
        return uuid.UUID(bytes=bytes(self._id_bytes[order[int(off[part]) + vidx]]))

    def query(self, v, k, nprobe, mode="build"):
        return self.query_with_events(v, k, nprobe, lambda e: None, mode)

    def query_with_events(self, v, k, nprobe, event, mode="build"):
        """One query vector -> Vec<QueryResult> (src/db/build.rs:307-340; mode="stored":
        src/db/stored.rs:331-389).  QueryEvent order is the reference's."""
        v = np.ascontiguousarray(v, np.float32).reshape(1, -1)
        m = capi.QUERY_BUILD if mode == "build" else capi.QUERY_STORED
        try:
            if mode != "build":   # stored::Database loads its codebooks on the first query (src/db/stored.rs:343-360)
                event(("StartingQueryInitialization",))
                event(("FinishedQueryInitialization",))
            event(("StartingPartitionSelection",))
            probes, _ = self.index.probe(v, nprobe, m)
            event(("FinishedPartitionSelection",))
            part, vidx, dist, cnt = self.index.query(v, k, nprobe, m)
        except capi.FdbError as e:
            if e.code == capi.ERR_INVALID_ARGS:
                raise _invalid_args(e.message) from e
            raise
        for p in probes[0]:
            event(("StartingPartitionQuery", int(p)))
            event(("FinishedPartitionQuery", int(p)))
        event(("StartingResultSelection",))
        out = [QueryResult(int(part[0, i]), self._vector_id(int(part[0, i]), int(vidx[0, i])),
                           int(vidx[0, i]), float(dist[0, i])) for i in range(int(cnt[0]))]
        event(("FinishedResultSelection",))
        return out

    def query_batch(self, queries, k, nprobe, mode="stored"):
        """Batched form (no reference analogue: it only has the single-vector call):
        returns (partition_index, vector_index, squared_distance, count) arrays."""
        m = capi.QUERY_BUILD if mode == "build" else capi.QUERY_STORED
        return self.index.query(queries, k, nprobe, m)

    def close(self):
        for h in (self.index, self.pkm, self.ckm, self.vs):
            h.close()
        if self._own_ctx:
            self.ctx.close()
