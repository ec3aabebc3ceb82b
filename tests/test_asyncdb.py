"""asyncdb mirror (flechasdb_b200/asyncdb.py): event order, lazy + concurrent loading and attributes, on the CPU
with the oracle standing in for the device index; the GPU twin of this test is in test_gpu_parity.py."""
import asyncio
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flechasdb_b200 import asyncdb, stored  # noqa: E402
from oracle import pyoracle as oracle  # noqa: E402


class OracleIndex:
    """stand-in for engine.Index.create_lazy: partitions arrive one by one, queries run on the CPU oracle"""

    def __init__(self, coarse, cbs):
        self.coarse, self.cbs = coarse, cbs
        self.P, self.D = coarse.shape[0], cbs.shape[0]
        self.parts = [np.zeros((0, self.D), np.uint32) for _ in range(self.P)]
        self.uploads = []

    def set_partition(self, p, codes):
        self.parts[p] = codes.astype(np.uint32)
        self.uploads.append(p)

    def _ix(self):
        off = np.concatenate([[0], np.cumsum([len(c) for c in self.parts])]).astype(np.uint64)
        return oracle.QueryIndex(self.coarse, self.cbs, off, np.concatenate(self.parts))

    def probe(self, v, nprobe, mode):
        rc, p, d = self._ix().probe(v[0], nprobe, 0)
        if rc != 0:
            from flechasdb_b200 import _capi as capi
            raise capi.FdbError(capi.ERR_INVALID_ARGS, "nprobe exceeds the number of partitions")
        return p[None, :], d[None, :]

    def query(self, v, k, nprobe, mode):
        rc, wp, wv, wd, wc = self._ix().query(v, k, nprobe, 0)
        assert rc == 0
        return wp, wv, wd, wc

    def close(self):
        pass


def _db(tmp_path, with_attrs=True):
    rng = np.random.default_rng(21)
    N, P, D, C, M = 32, 6, 4, 16, 300
    coarse = rng.random((P, N), dtype=np.float32)
    cbs = rng.random((D, C, N // D), dtype=np.float32) - np.float32(0.5)
    sizes = rng.multinomial(M, np.ones(P) / P)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
    codes = rng.integers(0, C, (M, D)).astype(np.uint32)
    ids = rng.integers(0, 256, (M, 16)).astype(np.uint8)
    table = {bytes(ids[i]): {"n": i, "tag": "t%d" % (i % 4)} for i in range(0, M, 2)} if with_attrs else None
    base = str(tmp_path / "adb")
    h = stored.serialize_arrays(base, coarse, cbs, off, codes, ids, table)
    return base, h, (coarse, cbs, off, codes, ids)


def test_async_query_event_order_lazy_loading_and_attributes(tmp_path, monkeypatch):
    from flechasdb_b200 import engine
    made = []
    monkeypatch.setattr(engine.Index, "create_lazy", staticmethod(lambda ctx, coarse, cbs: made.append(OracleIndex(coarse, cbs)) or made[-1]))
    base, h, (coarse, cbs, off, codes, ids) = _db(tmp_path)
    want_ix = oracle.QueryIndex(coarse, cbs, off, codes)
    rng = np.random.default_rng(22)
    q = rng.random((3, 32), dtype=np.float32)

    async def run():
        db = await asyncdb.AsyncStoredDatabase.load_database(None, base, h + ".binpb")
        assert db.index is None and db.partition_loads == 0          # header only
        ev = []
        res = await db.query(q[0], 5, 3, ev.append)
        names = [e[0] for e in ev]
        # query.rs:231-300: both lazy loads start before either finishes; selection follows
        assert names[:4] == ["StartingLoadingPartitionCentroids", "StartingLoadingCodebooks",
                             "FinishedLoadingPartitionCentroids", "FinishedLoadingCodebooks"]
        assert names[4:6] == ["StartingPartitionSelection", "FinishedPartitionSelection"]
        probed = [e[1] for e in ev if e[0] == "StartingLoadingPartition"]
        assert len(probed) == 3 and names[6:9] == ["StartingLoadingPartition"] * 3      # all three loads start together
        assert sorted(e[1] for e in ev if e[0] == "FinishedLoadingPartition") == sorted(probed)
        assert [e[1] for e in ev if e[0] == "StartingPartitionQueryExecution"] == probed
        assert names[-2:] == ["StartingKNNSelection", "FinishedKNNSelection"]
        assert db.partition_loads == 3 and sorted(made[0].uploads) == sorted(probed)
        rc, wp, wv, wd, wc = want_ix.query(q[:1], 5, 3, 0)
        assert [(r.partition_index, r.vector_index, r.squared_distance) for r in res] == \
               [(int(wp[0, i]), int(wv[0, i]), float(wd[0, i])) for i in range(int(wc[0]))]
        for r in res:
            gi = int(off[r.partition_index]) + r.vector_index
            assert r.vector_id.bytes == bytes(ids[gi])
            assert await r.get_attribute("n") == (gi if gi % 2 == 0 else None)
        assert sum(db.attributes_log_load_flags) <= 3
        # a second query: no initialisation events, only the partitions it has not seen are loaded
        ev2 = []
        await db.query(q[1], 5, 3, ev2.append)
        assert ev2[0] == ("StartingPartitionSelection",)
        assert all(e[1] not in probed for e in ev2 if e[0] == "StartingLoadingPartition")
        # concurrent queries share partition loads
        before = db.partition_loads
        await asyncio.gather(db.query(q[2], 5, 6), db.query(q[2], 5, 6))
        assert db.partition_loads == 6 and db.partition_loads - before <= 6
        assert await db.get_attribute(bytes(ids[4]), "tag") == "t0" and await db.get_attribute(bytes(ids[1]), "tag") is None
        with pytest.raises(stored.Error) as e:
            await db.get_attribute(bytes(16), "tag")
        assert e.value.kind == "InvalidArgs"
        with pytest.raises(stored.Error) as e:
            await db.query(q[0], 5, 7)                                # nprobe > P
        assert e.value.kind == "InvalidArgs"
        db.close()

    asyncio.run(run())


def test_async_query_reports_a_broken_partition_file_and_retries(tmp_path, monkeypatch):
    """a partition file that fails verification is an error of the query that probes it (src/io.rs:283-299), not a
    cached state: once the file is whole again the next query loads it"""
    from flechasdb_b200 import engine
    monkeypatch.setattr(engine.Index, "create_lazy", staticmethod(lambda ctx, coarse, cbs: OracleIndex(coarse, cbs)))
    base, h, _ = _db(tmp_path, with_attrs=False)
    q = np.random.default_rng(23).random((1, 32), dtype=np.float32)
    hdr = stored.parse(stored._open(base, h + ".binpb", True))
    pids = [v.decode() for _, v in hdr[10]]

    async def run():
        db = await asyncdb.AsyncStoredDatabase.load_database(None, base, h + ".binpb")
        files = {pid: os.path.join(base, "partitions", pid + ".binpb") for pid in pids}
        saved = {pid: open(f, "rb").read() for pid, f in files.items()}
        for pid, f in files.items():                    # flip a byte in every partition file
            raw = bytearray(saved[pid])
            raw[len(raw) // 2] ^= 0x55
            open(f, "wb").write(bytes(raw))
        with pytest.raises(Exception):
            await db.query(q[0], 5, 6)
        assert db.partition_loads == 0
        for pid, f in files.items():
            open(f, "wb").write(saved[pid])
        res = await db.query(q[0], 5, 6)
        assert len(res) == 5 and db.partition_loads == 6
        db.close()

    asyncio.run(run())
