"""Async twin of the stored database: asyncdb::stored::Database (src/asyncdb/stored.rs) and its query future
(src/asyncdb/stored/query.rs:221-355) on top of the same GPU index.

What the reference overlaps, this mirror overlaps too:

  * load_database reads only the header (src/asyncdb/stored.rs: load_database); the partition centroids and the
    codebooks are loaded -- concurrently -- by the first query (query.rs:231-300);
  * the probed partitions' files are read concurrently (query.rs:242-255: one PartitionQuery future per selected
    partition); here every file is read, inflated and parsed on a worker thread, and a partition's codes are
    uploaded to the device (fdb_index_set_partition) as soon as ITS file has arrived, while the other files are
    still being read;
  * the ADC scan of the probed partitions and the k-NN selection (query.rs:301-345) run as ONE device call once the
    last probed partition is resident -- the per-partition "execution" events are fired around it in probe order.

Attribute logs stay host-side hash maps, loaded per partition on first use (src/asyncdb/stored/get_attribute.rs).
All arithmetic happens in libflechasdb_b200.so; this file sequences I/O and calls.
"""
import asyncio
import uuid

import numpy as np

from . import _capi as capi
from . import stored
from .db import QueryResult


class AsyncQueryResult(QueryResult):
    """asyncdb::stored::QueryResult: get_attribute is a coroutine (src/asyncdb/stored.rs: QueryResult::get_attribute)"""

    async def get_attribute(self, key):
        return await self.db.get_attribute_in_partition(self.partition_index, self.vector_id, key)


class AsyncStoredDatabase:
    """asyncdb::stored::Database<f32, FS>: nothing but the header is read up front."""

    def __init__(self, ctx, base, N, P, D, C, partition_ids, centroids_id, codebook_ids, attributes_log_ids, attribute_names):
        if C > 256:
            raise stored.Error("InvalidData", "num_codes > 256 is not supported by the u8 device layout")
        self.ctx, self.base = ctx, base
        self.vector_size, self.num_partitions, self.num_divisions, self.num_codes = N, P, D, C
        self.partition_ids, self.centroids_id, self.codebook_ids = partition_ids, centroids_id, codebook_ids
        self.attributes_log_ids, self.attribute_names = attributes_log_ids, attribute_names
        self.index = None                  # created when centroids and codebooks have arrived
        self.ids = [None] * P
        self.partition_loads = 0
        self.attribute_table = None
        self.attributes_log_load_flags = [False] * P
        self._partition_tasks = {}         # partition -> the task that loads it (shared by concurrent queries)
        self._init_task = None

    @classmethod
    async def load_database(cls, ctx, base, path):
        """asyncdb::stored::LoadDatabase::load_database: header only, validated like the sync loader
        (src/db/stored.rs:671-706)"""
        loop = asyncio.get_running_loop()
        hdr = await loop.run_in_executor(None, lambda: stored.parse(stored._open(base, path, True)))

        def u(field):
            v = hdr.get(field)
            return int(v[0][1]) if v else 0

        def strs(field):
            return [v.decode() for _, v in hdr.get(field, [])]

        N, P, D, C = u(1), u(2), u(3), u(4)
        for name, v in (("vector_size", N), ("num_divisions", D), ("num_partitions", P), ("num_codes", C)):
            if v == 0:
                raise stored.Error("InvalidData", "%s is zero" % name)
        if N % D:
            raise stored.Error("InvalidData", "vector_size %d is not multiple of num_divisions %d" % (N, D))
        pids, cids = strs(10), strs(12)
        if len(pids) != P:
            raise stored.Error("InvalidData", "num_partitions %d and partition_ids.len() %d do not match" % (P, len(pids)))
        if len(cids) != D:
            raise stored.Error("InvalidData", "num_divisions %d and codebook_ids.len() %d do not match" % (D, len(cids)))
        cen = strs(11)
        if not cen:
            raise stored.Error("InvalidData", "partition_centroids_id is missing")
        return cls(ctx, base, N, P, D, C, pids, cen[0], cids, strs(13), strs(14))

    # ---- lazy pieces -------------------------------------------------------------------------------------------
    def _read_centroids(self):
        raw = stored._open(self.base, "partitions/%s.%s" % (self.centroids_id, stored.EXT), False, verify=False)
        coarse = stored._parse_floats(raw, 10)
        if coarse.size != self.num_partitions * self.vector_size:
            raise stored.Error("InvalidData", "partition centroids data length mismatch: expected %d, got %d"
                               % (self.num_partitions, coarse.size // self.vector_size))
        return coarse.reshape(self.num_partitions, self.vector_size)

    def _read_codebooks(self):
        s = self.vector_size // self.num_divisions
        cbs = np.zeros((self.num_divisions, self.num_codes, s), np.float32)
        for d, cid in enumerate(self.codebook_ids):
            data = stored._parse_floats(stored._open(self.base, "codebooks/%s.%s" % (cid, stored.EXT), False), 10)
            if data.size != self.num_codes * s:
                raise stored.Error("InvalidData", "codebook %d has %d elements, expected %d" % (d, data.size, self.num_codes * s))
            cbs[d] = data.reshape(self.num_codes, s)
        return cbs

    async def _initialize(self, event):
        """partition centroids and codebooks, read concurrently (query.rs:257-300), then the empty device index"""
        from .engine import Index
        loop = asyncio.get_running_loop()
        event(("StartingLoadingPartitionCentroids",))
        f_cen = loop.run_in_executor(None, self._read_centroids)
        event(("StartingLoadingCodebooks",))
        f_cbs = loop.run_in_executor(None, self._read_codebooks)
        coarse = await f_cen
        event(("FinishedLoadingPartitionCentroids",))
        cbs = await f_cbs
        event(("FinishedLoadingCodebooks",))
        self.index = Index.create_lazy(self.ctx, coarse, cbs)

    async def _load_partition(self, p, event):
        """one partition: file -> codes + ids on a worker thread, upload on the caller's thread"""
        loop = asyncio.get_running_loop()
        event(("StartingLoadingPartition", p))
        try:
            codes, ids = await loop.run_in_executor(
                None, stored.load_partition, self.base, self.partition_ids[p], p, self.vector_size, self.num_divisions,
                self.num_codes)
            self.index.set_partition(p, codes.astype(np.uint8))     # the other probed partitions are still being read
        except BaseException:
            self._partition_tasks.pop(p, None)      # a failed load is not cached: the next query tries again
            raise
        self.ids[p] = ids
        self.partition_loads += 1
        event(("FinishedLoadingPartition", p))

    def _partition_task(self, p, event):
        if p not in self._partition_tasks:
            self._partition_tasks[p] = asyncio.ensure_future(self._load_partition(p, event))
        return self._partition_tasks[p]

    # ---- query (src/asyncdb/stored/query.rs:221-355) -----------------------------------------------------------
    async def query(self, v, k, nprobe, event=lambda e: None):
        if k <= 0 or nprobe <= 0:
            raise stored.Error("InvalidArgs", "NonZeroUsize")
        if self._init_task is None:
            self._init_task = asyncio.ensure_future(self._initialize(event))
        await self._init_task
        v = np.asarray(v, np.float32).reshape(1, -1)
        event(("StartingPartitionSelection",))
        try:
            probes, _ = self.index.probe(v, nprobe, capi.QUERY_STORED)
        except capi.FdbError as e:
            if e.code == capi.ERR_INVALID_ARGS:
                raise stored.Error("InvalidArgs", e.message) from e
            raise
        event(("FinishedPartitionSelection",))
        plist = [int(p) for p in probes[0]]
        if not plist:
            raise stored.Error("InvalidContext", "no partitions selected for query")
        # (every load is awaited before an error is reported: no load of this query is left running behind it)
        loaded = await asyncio.gather(*[self._partition_task(p, event) for p in plist], return_exceptions=True)
        for r in loaded:
            if isinstance(r, BaseException):
                raise r
        for p in plist:
            event(("StartingPartitionQueryExecution", p))
        part, vi, d, c = self.index.query(v, k, nprobe, capi.QUERY_STORED)
        for p in plist:
            event(("FinishedPartitionQueryExecution", p))
        event(("StartingKNNSelection",))
        out = [AsyncQueryResult(int(part[0, i]), uuid.UUID(bytes=bytes(self.ids[int(part[0, i])][int(vi[0, i])])),
                                int(vi[0, i]), float(d[0, i]), self) for i in range(int(c[0]))]
        event(("FinishedKNNSelection",))
        return out

    # ---- attributes (src/asyncdb/stored/get_attribute.rs) ------------------------------------------------------
    async def load_attributes_log(self, p):
        if self.attributes_log_load_flags[p]:
            return
        if self._init_task is None:
            self._init_task = asyncio.ensure_future(self._initialize(lambda e: None))
        await self._init_task
        await self._partition_task(p, lambda e: None)
        if p >= len(self.attributes_log_ids):
            raise stored.Error("InvalidData", "no attributes log for partition %d" % p)
        loop = asyncio.get_running_loop()
        entries = await loop.run_in_executor(None, stored.load_attributes_log, self.base, self.attributes_log_ids[p],
                                             self.partition_ids[p], p, self.attribute_names)
        if self.attribute_table is None:
            self.attribute_table = {}
        for idb, name, value in entries:
            self.attribute_table.setdefault(idb, {})[name] = value
        for row in self.ids[p]:
            self.attribute_table.setdefault(bytes(row), {})
        self.attributes_log_load_flags[p] = True

    async def get_attribute_in_partition(self, partition_index, vector_id, key):
        await self.load_attributes_log(partition_index)
        return self._get_attribute_internal(vector_id, key)

    async def get_attribute(self, vector_id, key):
        """every log, read concurrently, then the look-up (an addition: the reference's async database only has the
        per-result form, src/asyncdb/stored/get_attribute.rs)"""
        await asyncio.gather(*[self.load_attributes_log(p) for p in range(self.num_partitions)])
        return self._get_attribute_internal(vector_id, key)

    def _get_attribute_internal(self, vector_id, key):
        idb = vector_id.bytes if isinstance(vector_id, uuid.UUID) else bytes(vector_id)
        attrs = (self.attribute_table or {}).get(idb)
        if attrs is None:
            raise stored.Error("InvalidArgs", "no such vector ID: %s" % uuid.UUID(bytes=idb))
        return attrs.get(key)

    def close(self):
        if self.index is not None:
            self.index.close()
            self.index = None
