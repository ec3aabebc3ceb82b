"""Thin object wrappers over the C ABI handles (fdb_ctx / fdb_vs / fdb_km / fdb_index).

Nothing is computed here: every method is one call into libflechasdb_b200.so.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from ._capi import check, f32p, u32p, u64p, u8p, as_f32, as_u32, VP


class Context:
    def __init__(self, device=0):
        self.h = VP()
        check(capi.lib().fdb_ctx_create(device, C.byref(self.h)))
        self.device = device

    def sync(self):
        check(capi.lib().fdb_ctx_sync(self.h))

    def timer_start(self):
        check(capi.lib().fdb_ctx_timer_start(self.h))

    def timer_stop(self):
        ms = C.c_float()
        check(capi.lib().fdb_ctx_timer_stop(self.h, C.byref(ms)))
        return float(ms.value)

    @property
    def launches(self):
        return int(capi.lib().fdb_ctx_launch_count(self.h))

    def flush_l2(self):
        check(capi.lib().fdb_device_flush_l2(self.h))

    def alloc(self, nbytes):
        p = VP()
        check(capi.lib().fdb_device_alloc(self.h, nbytes, C.byref(p)))
        return p

    def free(self, p):
        check(capi.lib().fdb_device_free(self.h, p))

    def host_register(self, array):
        """page-locks a numpy array in place (asynchronous copies from / to it); undo with host_unregister"""
        check(capi.lib().fdb_host_register(self.h, array.ctypes.data_as(VP), array.nbytes))

    def host_unregister(self, array):
        check(capi.lib().fdb_host_unregister(self.h, array.ctypes.data_as(VP)))

    def upload(self, dptr, array):
        a = np.ascontiguousarray(array)
        check(capi.lib().fdb_device_upload(self.h, dptr, a.ctypes.data_as(VP), a.nbytes))

    def download(self, dptr, shape, dtype):
        out = np.empty(shape, dtype)
        check(capi.lib().fdb_device_download(self.h, out.ctypes.data_as(VP), dptr, out.nbytes))
        return out

    def fill_uniform(self, dptr, count, seed, start=0):
        check(capi.lib().fdb_device_fill_uniform(self.h, dptr, count, seed, start))

    def close(self):
        if self.h:
            capi.lib().fdb_ctx_destroy(self.h)
            self.h = VP()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class Comm:
    """One rank of the library's NCCL communicator (fdb_comm), bound to a Context.

    `exchange(id_bytes_or_None) -> id_bytes` hands rank 0's id to the other ranks; any transport
    will do (bench.py uses a torch.distributed broadcast, the tests a file)."""

    def __init__(self, ctx, world, rank, exchange=None):
        self.ctx, self.world, self.rank = ctx, world, rank
        self.h = VP()
        ident = None
        if world > 1:
            ident = np.zeros(capi.COMM_ID_BYTES, np.uint8)
            if rank == 0:
                check(capi.lib().fdb_comm_unique_id(u8p(ident)))
            ident = np.frombuffer(exchange(ident.tobytes() if rank == 0 else None), np.uint8).copy()
        check(capi.lib().fdb_comm_create(ctx.h, world, rank, None if ident is None else u8p(ident), C.byref(self.h)))

    @property
    def collectives(self):
        return int(capi.lib().fdb_comm_collective_count(self.h))

    def max_f64(self, values):
        """max over the ranks of a few host doubles (device timings); also a barrier"""
        v = np.ascontiguousarray(np.atleast_1d(values), np.float64).copy()
        check(capi.lib().fdb_comm_max_f64(self.h, v.ctypes.data_as(C.POINTER(C.c_double)), len(v)))
        return v

    def barrier(self):
        self.max_f64([0.0])

    def close(self):
        if self.h:
            capi.lib().fdb_comm_destroy(self.h)
            self.h = VP()


def shard_rows(n, world, rank):
    """contiguous row shard [lo, hi) of `rank` (the convention of fdb_kmeans_*_sharded)"""
    return n * rank // world, n * (rank + 1) // world


class VectorSet:
    """BlockVectorSet<f32> in HBM (src/vector.rs:28-100)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = handle

    @classmethod
    def upload(cls, ctx, rows):
        rows = as_f32(rows)
        assert rows.ndim == 2
        h = VP()
        check(capi.lib().fdb_vs_upload(ctx.h, f32p(rows), rows.shape[0], rows.shape[1], C.byref(h)))
        return cls(ctx, h)

    @classmethod
    def generate(cls, ctx, n, dim, seed, start=0):
        h = VP()
        check(capi.lib().fdb_vs_generate(ctx.h, n, dim, seed, start, C.byref(h)))
        return cls(ctx, h)

    def device_ptr(self):
        """Device address of the rows (for the *_device entry points)."""
        return capi.lib().fdb_vs_device_ptr(self.h)

    def __len__(self):
        return int(capi.lib().fdb_vs_len(self.h))

    @property
    def vector_size(self):
        return int(capi.lib().fdb_vs_vector_size(self.h))

    def download(self, first=0, nrows=None):
        nrows = len(self) - first if nrows is None else nrows
        out = np.empty((nrows, self.vector_size), np.float32)
        check(capi.lib().fdb_vs_download_rows(self.h, first, nrows, f32p(out)))
        return out

    def subtract_assigned(self, km):
        check(capi.lib().fdb_vs_subtract_assigned(self.h, km.h))

    def close(self):
        if self.h:
            capi.lib().fdb_vs_destroy(self.h)
            self.h = VP()


class KMeans:
    """nb side-by-side k-means problems over strided sub-vector views (src/kmeans.rs)."""

    def __init__(self, vs, k, col_off=0, dim=None, nb=1):
        dim = vs.vector_size if dim is None else dim
        self.vs, self.k, self.dim, self.nb, self.n = vs, k, dim, nb, len(vs)
        self.h = VP()
        check(capi.lib().fdb_kmeans_begin(vs.h, col_off, dim, nb, k, C.byref(self.h)))

    def seed_first(self, ci):
        ci = as_u32(np.atleast_1d(ci))
        check(capi.lib().fdb_kmeans_seed_first(self.h, u32p(ci)))

    def seed_total(self):
        t = np.zeros(self.nb, np.float32)
        check(capi.lib().fdb_kmeans_seed_total(self.h, f32p(t)))
        return t

    def seed_pick(self, u01, exact=False):
        u = as_f32(np.atleast_1d(u01))
        out = np.zeros(self.nb, np.uint32)
        check(capi.lib().fdb_kmeans_seed_pick(self.h, f32p(u), int(exact), u32p(out)))
        return out

    def seed_add(self, i, ci, exact=False):
        ci = as_u32(np.atleast_1d(ci))
        check(capi.lib().fdb_kmeans_seed_add(self.h, i, u32p(ci), int(exact)))

    def seed_run(self, first, u01, exact=False):
        first = as_u32(np.atleast_1d(first))
        u = as_f32(u01).reshape(self.nb, max(self.k - 1, 0))
        picked = np.zeros((self.nb, self.k), np.uint32)
        check(capi.lib().fdb_kmeans_seed_run(self.h, u32p(first), f32p(u), int(exact), u32p(picked)))
        return picked

    def seed_round_ext(self, i, centres, local_ci):
        c = as_f32(centres).reshape(self.nb, self.dim)
        l = as_u32(np.atleast_1d(local_ci))
        check(capi.lib().fdb_kmeans_seed_round_ext(self.h, i, f32p(c), u32p(l)))

    def seed_pick_value(self, values):
        v = as_f32(np.atleast_1d(values))
        out = np.zeros(self.nb, np.uint32)
        check(capi.lib().fdb_kmeans_seed_pick_value(self.h, f32p(v), u32p(out)))
        return out

    def seed_chosen(self, chosen):
        ch = as_u32(chosen).reshape(self.nb, self.k)
        check(capi.lib().fdb_kmeans_seed_chosen(self.h, u32p(ch)))

    def set_state(self, centroids, indices=None):
        c = as_f32(centroids).reshape(self.nb, self.k, self.dim)
        i = None if indices is None else as_u32(indices).reshape(self.nb, self.n)
        check(capi.lib().fdb_kmeans_set_state(self.h, f32p(c), None if i is None else u32p(i)))

    def _active(self, active):
        if active is None:
            return None, None
        a = np.ascontiguousarray(active, np.uint8)
        return a, u8p(a)

    def update(self, active=None):
        g = np.zeros(self.nb, np.float32)
        keep, a = self._active(active)
        check(capi.lib().fdb_kmeans_update(self.h, a, f32p(g)))
        return g

    def reassign(self, active=None):
        keep, a = self._active(active)
        check(capi.lib().fdb_kmeans_reassign(self.h, a))

    def run(self, max_rounds=capi.KMEANS_MAX_ROUNDS, eps=capi.KMEANS_EPSILON):
        g = np.zeros((self.nb, max_rounds), np.float32)
        rounds = np.zeros(self.nb, np.uint32)
        reas = np.zeros(self.nb, np.uint32)
        check(capi.lib().fdb_kmeans_run(self.h, max_rounds, eps, f32p(g), u32p(rounds), u32p(reas)))
        return [g[b, :rounds[b]].copy() for b in range(self.nb)], rounds, reas

    def seed_run_sharded(self, comm, n_global, first_global, u01):
        """k-means++ over row shards, one packed all-gather per round; returns the picked global rows [nb][k]"""
        first = as_u32(np.atleast_1d(first_global))
        u = as_f32(u01).reshape(self.nb, max(self.k - 1, 0))
        picked = np.zeros((self.nb, self.k), np.uint32)
        check(capi.lib().fdb_kmeans_seed_run_sharded(self.h, comm.h, n_global, u32p(first), f32p(u), u32p(picked)))
        return picked

    def run_sharded(self, comm, max_rounds=capi.KMEANS_MAX_ROUNDS, eps=capi.KMEANS_EPSILON):
        """the Lloyd loop over row shards, one all-reduce per round; returns what run() returns"""
        g = np.zeros((self.nb, max_rounds), np.float32)
        rounds = np.zeros(self.nb, np.uint32)
        reas = np.zeros(self.nb, np.uint32)
        check(capi.lib().fdb_kmeans_run_sharded(self.h, comm.h, max_rounds, eps, f32p(g), u32p(rounds), u32p(reas)))
        return [g[b, :rounds[b]].copy() for b in range(self.nb)], rounds, reas

    def last_assign_info(self):
        a = (C.c_uint32 * 3)()
        check(capi.lib().fdb_kmeans_last_assign_info(
            self.h, C.cast(C.byref(a, 0), capi.U32P), C.cast(C.byref(a, 4), capi.U32P),
            C.cast(C.byref(a, 8), capi.U32P)))
        return dict(tensor_cores=bool(a[0]), rechecked=int(a[1]), overflow=int(a[2]))

    def get(self):
        c = np.zeros((self.nb, self.k, self.dim), np.float32)
        i = np.zeros((self.nb, self.n), np.uint32)
        check(capi.lib().fdb_kmeans_get(self.h, f32p(c), u32p(i)))
        return c, i

    def weights(self):
        w = np.zeros((self.nb, self.n), np.float32)
        check(capi.lib().fdb_kmeans_get_weights(self.h, f32p(w)))
        return w

    def update_partial(self):
        p, n = VP(), C.c_size_t()
        check(capi.lib().fdb_kmeans_update_partial(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def update_finish(self):
        g = np.zeros(self.nb, np.float32)
        check(capi.lib().fdb_kmeans_update_finish(self.h, f32p(g)))
        return g

    def close(self):
        if self.h:
            capi.lib().fdb_kmeans_destroy(self.h)
            self.h = VP()


class Index:
    """Queryable IVF-PQ index resident in HBM (stored::Database, src/db/stored.rs:41-57)."""

    def __init__(self, ctx, handle, N, P, D, Cn):
        self.ctx, self.h = ctx, handle
        self.N, self.P, self.D, self.C = N, P, D, Cn

    @classmethod
    def create(cls, ctx, coarse, codebooks, offsets, codes):
        coarse, codebooks = as_f32(coarse), as_f32(codebooks)
        P, N = coarse.shape
        D, Cn, _ = codebooks.shape
        offsets = np.ascontiguousarray(offsets, np.uint64)
        codes = np.ascontiguousarray(codes, np.uint8)
        h = VP()
        check(capi.lib().fdb_index_create(ctx.h, N, P, D, Cn, f32p(coarse), f32p(codebooks),
                                          u64p(offsets), u8p(codes), C.byref(h)))
        return cls(ctx, h, N, P, D, Cn)

    @classmethod
    def create_lazy(cls, ctx, coarse, codebooks):
        """partition centroids + codebooks only; the code lists arrive with set_partition (src/db/stored.rs:269-293)"""
        coarse, codebooks = as_f32(coarse), as_f32(codebooks)
        P, N = coarse.shape
        D, Cn, _ = codebooks.shape
        h = VP()
        check(capi.lib().fdb_index_create_lazy(ctx.h, N, P, D, Cn, f32p(coarse), f32p(codebooks), C.byref(h)))
        return cls(ctx, h, N, P, D, Cn)

    def set_partition(self, p, codes):
        codes = np.ascontiguousarray(codes, np.uint8).reshape(-1, self.D)
        check(capi.lib().fdb_index_set_partition(self.h, int(p), u8p(codes), codes.shape[0]))

    def partition_loaded(self, p):
        return bool(capi.lib().fdb_index_partition_loaded(self.h, int(p)))

    def missing_partitions(self, q, nprobe, mode=capi.QUERY_STORED):
        q = as_f32(q).reshape(-1, self.N)
        out = np.zeros(self.P, np.uint32)
        n = C.c_size_t()
        check(capi.lib().fdb_index_missing_partitions(self.h, f32p(q), q.shape[0], nprobe, mode, u32p(out), self.P, C.byref(n)))
        return out[:n.value].copy()

    @classmethod
    def from_build(cls, ctx, coarse_km, pq_km):
        h = VP()
        check(capi.lib().fdb_index_from_build(ctx.h, coarse_km.h, pq_km.h, C.byref(h)))
        return cls(ctx, h, coarse_km.dim, coarse_km.k, pq_km.nb, pq_km.k)

    @property
    def num_vectors(self):
        return int(capi.lib().fdb_index_num_vectors(self.h))

    def layout(self, order=True):
        M = self.num_vectors
        off = np.zeros(self.P + 1, np.uint64)
        od = np.zeros(M, np.uint32) if order else None
        codes = np.zeros((M, self.D), np.uint8)
        check(capi.lib().fdb_index_get_layout(self.h, u64p(off), u32p(od) if order else None,
                                              u8p(codes)))
        return off, od, codes

    def query(self, q, k, nprobe, mode=capi.QUERY_STORED):
        q = as_f32(q).reshape(-1, self.N)
        nq = q.shape[0]
        part = np.zeros((nq, k), np.uint32)
        vidx = np.zeros((nq, k), np.uint32)
        dist = np.zeros((nq, k), np.float32)
        cnt = np.zeros(nq, np.uint32)
        check(capi.lib().fdb_index_query(self.h, f32p(q), nq, k, nprobe, mode, u32p(part),
                                         u32p(vidx), f32p(dist), u32p(cnt)))
        return part, vidx, dist, cnt

    def query_device(self, d_q, nq, k, nprobe, d_part, d_vidx, d_dist, d_cnt,
                     mode=capi.QUERY_STORED):
        check(capi.lib().fdb_index_query_device(self.h, d_q, nq, k, nprobe, mode, d_part, d_vidx,
                                                d_dist, d_cnt))

    def query_sharded(self, comm, d_q, nq, k, nprobe, d_part, d_vidx, d_dist, d_cnt, mode=capi.QUERY_STORED):
        """code lists sharded over the ranks of comm: same batch on every rank, one packed all-gather, merged
        result (identical on every rank) left in the device buffers"""
        check(capi.lib().fdb_index_query_sharded(self.h, comm.h, d_q, nq, k, nprobe, mode, d_part, d_vidx, d_dist, d_cnt))

    def last_sharded_ties(self):
        t = C.c_uint32()
        check(capi.lib().fdb_index_last_sharded_ties(self.h, C.byref(t)))
        return int(t.value)

    def probe(self, q, nprobe, mode=capi.QUERY_STORED):
        q = as_f32(q).reshape(-1, self.N)
        nq = q.shape[0]
        part = np.zeros((nq, nprobe), np.uint32)
        dist = np.zeros((nq, nprobe), np.float32)
        check(capi.lib().fdb_index_probe(self.h, f32p(q), nq, nprobe, mode, u32p(part), f32p(dist)))
        return part, dist

    def table(self, q, partition):
        q = as_f32(q).reshape(self.N)
        t = np.zeros((self.D, self.C), np.float32)
        check(capi.lib().fdb_index_table(self.h, f32p(q), int(partition), f32p(t)))
        return t

    def set_timing(self, on):
        check(capi.lib().fdb_index_set_timing(self.h, int(on)))

    def last_timing(self):
        ms = np.zeros(6, np.float32)
        b = C.c_uint64()
        check(capi.lib().fdb_index_last_timing(self.h, f32p(ms), C.byref(b)))
        return ms, int(b.value)

    def last_stats(self):
        """(queries on the ADC filter path, queries on the exact pipeline, candidates evaluated
        exactly, code vectors scanned) of the last query call."""
        st = np.zeros(4, np.uint64)
        check(capi.lib().fdb_index_last_stats(self.h, u64p(st)))
        return tuple(int(x) for x in st)

    SCAN_KERNELS = ("none (exact pipeline)", "fscan_kernel (query-major, f32 table per query)",
                    "pscan_kernel (partition-major, f32 tables, 16 queries per item)",
                    "pscan16_kernel (partition-major, 16-bit tables, 32 queries per item)",
                    "vscan_kernel (vector-lane, packed 16-bit tables, 8 queries per 128-bit look-up)")

    def last_scan_kernel(self):
        """name of the code-scan kernel the last query call ran"""
        kind = C.c_int()
        check(capi.lib().fdb_index_last_scan_kernel(self.h, C.byref(kind), None))
        return self.SCAN_KERNELS[kind.value]

    def last_scan_kernel_ms(self):
        """device milliseconds of the code-scan kernel's launches alone in the last query call (timing enabled)"""
        kind, ms = C.c_int(), C.c_float()
        check(capi.lib().fdb_index_last_scan_kernel(self.h, C.byref(kind), C.byref(ms)))
        return float(ms.value)

    def debug_band(self, nq, nprobe):
        """Test hook (after query_device): error bound E[q], the candidates' approximate distances and
        flat positions, their count, and the probe lists the filter path scanned."""
        E = np.zeros(nq, np.float32)
        approx = np.zeros((nq, 32), np.float32)
        flat = np.zeros((nq, 32), np.uint32)
        cnt = np.zeros(nq, np.uint32)
        probes = np.zeros((nq, nprobe), np.uint32)
        check(capi.lib().fdb_index_debug_band(self.h, nq, nprobe, f32p(E), f32p(approx), u32p(flat),
                                              u32p(cnt), u32p(probes)))
        return E, approx, flat, cnt, probes

    def close(self):
        if self.h:
            capi.lib().fdb_index_destroy(self.h)
            self.h = VP()
