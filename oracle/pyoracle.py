"""ctypes face of the CPU oracle (oracle/flechas_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's CPU legs.  The product package (flechasdb_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

OK = 0
ERR_INVALID_ARGS = -1
ERR_PANIC_EMPTY_CLUSTER = -4
ERR_PANIC_WEIGHTS = -5
ERR_PANIC_NAN = -6

F32P = C.POINTER(C.c_float)
U32P = C.POINTER(C.c_uint32)
U64P = C.POINTER(C.c_uint64)


class View(C.Structure):
    _fields_ = [("base", F32P), ("n", C.c_size_t), ("stride", C.c_size_t),
                ("off", C.c_size_t), ("dim", C.c_size_t)]


class WI(C.Structure):
    _fields_ = [("weights", F32P), ("n", C.c_size_t), ("total", C.c_float),
                ("scale", C.c_float)]


class Item(C.Structure):
    _fields_ = [("key", C.c_float), ("a", C.c_uint32), ("b", C.c_uint32)]


class NBest(C.Structure):
    _fields_ = [("items", C.POINTER(Item)), ("n", C.c_size_t), ("len", C.c_size_t)]


class Index(C.Structure):
    _fields_ = [("N", C.c_size_t), ("P", C.c_size_t), ("D", C.c_size_t), ("C", C.c_size_t),
                ("coarse", F32P), ("codebooks", F32P), ("offsets", U64P), ("codes_pm", U32P)]


def build(native=False):
    """Compile the oracle with the committed Makefile (gcc only)."""
    target = "native" if native else "all"
    subprocess.run(["make", "-s", "-C", HERE, target], check=True)
    return os.path.join(HERE, "liboracle_native.so" if native else "liboracle.so")


_libs = {}


def lib(native=False):
    if native in _libs:
        return _libs[native]
    path = os.path.join(HERE, "liboracle_native.so" if native else "liboracle.so")
    src = os.path.join(HERE, "flechas_oracle.c")
    if (not os.path.exists(path)) or os.path.getmtime(path) < os.path.getmtime(src):
        build(native)
    L = C.CDLL(path)
    sz = C.c_size_t
    L.fo_dot.restype = C.c_float
    L.fo_dot.argtypes = [F32P, F32P, sz]
    L.fo_dot_naive.restype = C.c_float
    L.fo_dot_naive.argtypes = [F32P, F32P, sz]
    L.fo_norm2.restype = C.c_float
    L.fo_norm2.argtypes = [F32P, sz]
    L.fo_sum.restype = C.c_float
    L.fo_sum.argtypes = [F32P, sz]
    L.fo_min.restype = C.c_int
    L.fo_min.argtypes = [F32P, sz, F32P]
    L.fo_max_abs.restype = C.c_int
    L.fo_max_abs.argtypes = [F32P, sz, F32P]
    L.fo_add_in.argtypes = [F32P, F32P, sz]
    L.fo_subtract.argtypes = [F32P, F32P, F32P, sz]
    L.fo_subtract_in.argtypes = [F32P, F32P, sz]
    L.fo_scale_in.argtypes = [F32P, C.c_float, sz]
    L.fo_sqdist.restype = C.c_float
    L.fo_sqdist.argtypes = [F32P, F32P, sz, F32P]
    L.fo_chunk_check.argtypes = [sz, sz]
    L.fo_divide.argtypes = [C.POINTER(View), sz, C.POINTER(View)]
    L.fo_wi_new.argtypes = [C.POINTER(WI), F32P, sz]
    L.fo_wi_free.argtypes = [C.POINTER(WI)]
    L.fo_wi_update.argtypes = [C.POINTER(WI), C.POINTER(sz), F32P, sz]
    L.fo_wi_get_weight.restype = C.c_float
    L.fo_wi_get_weight.argtypes = [C.POINTER(WI), sz]
    L.fo_wi_pick.restype = sz
    L.fo_wi_pick.argtypes = [C.POINTER(WI), C.c_float]
    L.fo_wi_sample_value.restype = C.c_float
    L.fo_wi_sample_value.argtypes = [C.POINTER(WI), C.c_float]
    L.fo_kmeans_init.argtypes = [C.POINTER(View), sz, sz, U32P, F32P, F32P, U32P, F32P, U32P]
    L.fo_kmeans_update.argtypes = [C.POINTER(View), sz, F32P, U32P, F32P]
    L.fo_kmeans_reassign.argtypes = [C.POINTER(View), sz, F32P, U32P, C.c_int]
    L.fo_kmeans_lloyd.argtypes = [C.POINTER(View), sz, F32P, U32P, sz, C.c_float, F32P,
                                  C.POINTER(sz), C.POINTER(sz), C.c_int]
    L.fo_residues.argtypes = [C.POINTER(View), sz, F32P, U32P]
    L.fo_extract_partitions.argtypes = [sz, sz, sz, U32P, U32P, U64P, U32P, U32P]
    L.fo_nbest_push.argtypes = [C.POINTER(NBest), Item]
    L.fo_query.argtypes = [C.POINTER(Index), F32P, sz, sz, sz, C.c_int, U32P, U32P, F32P, U32P,
                           C.c_int]
    L.fo_query_probe.argtypes = [C.POINTER(Index), F32P, sz, C.c_int, U32P, F32P]
    L.fo_query_table.argtypes = [C.POINTER(Index), F32P, C.c_uint32, F32P]
    L.fo_fill_uniform.argtypes = [F32P, sz, C.c_uint64, C.c_uint64]
    L.fo_splitmix64.restype = C.c_uint64
    L.fo_splitmix64.argtypes = [C.c_uint64, C.c_uint64]
    _libs[native] = L
    return L


def _f(a):
    return a.ctypes.data_as(F32P)


def _u(a):
    return a.ctypes.data_as(U32P)


def f32(x):
    return np.ascontiguousarray(x, dtype=np.float32)


# ---- linalg -----------------------------------------------------------------
def dot(x, y):
    x, y = f32(x), f32(y)
    assert x.size == y.size
    return float(lib().fo_dot(_f(x), _f(y), x.size))


def norm2(x):
    x = f32(x)
    return float(lib().fo_norm2(_f(x), x.size))


def sum_(x):
    x = f32(x)
    return float(lib().fo_sum(_f(x), x.size))


def min_(x):
    x = f32(x)
    o = C.c_float()
    return float(o.value) if lib().fo_min(_f(x), x.size, C.byref(o)) else None


def max_abs(x):
    x = f32(x)
    o = C.c_float()
    return float(o.value) if lib().fo_max_abs(_f(x), x.size, C.byref(o)) else None


def add_in(l, r):
    l, r = f32(l).copy(), f32(r)
    lib().fo_add_in(_f(l), _f(r), l.size)
    return l


def subtract(l, r):
    l, r = f32(l), f32(r)
    o = np.empty_like(l)
    lib().fo_subtract(_f(l), _f(r), _f(o), l.size)
    return o


def subtract_in(l, r):
    l, r = f32(l).copy(), f32(r)
    lib().fo_subtract_in(_f(l), _f(r), l.size)
    return l


def scale_in(x, a):
    x = f32(x).copy()
    lib().fo_scale_in(_f(x), a, x.size)
    return x


def sqdist(v, c):
    v, c = f32(v), f32(c)
    buf = np.empty_like(v)
    return float(lib().fo_sqdist(_f(v), _f(c), v.size, _f(buf)))


# ---- views ----------------------------------------------------------------------
def view(block, off=0, dim=None):
    """block: C-contiguous float32 (n, stride).  The caller keeps `block` alive."""
    assert block.dtype == np.float32 and block.flags.c_contiguous and block.ndim == 2
    n, stride = block.shape
    return View(_f(block), n, stride, off, stride if dim is None else dim)


def divide(block, d):
    v = view(block)
    out = (View * d)()
    rc = lib().fo_divide(C.byref(v), d, out)
    return rc, list(out)


# ---- distribution ----------------------------------------------------------------
class WeightedIndex:
    def __init__(self, weights):
        w = f32(weights)
        self._wi = WI()
        self.rc = lib().fo_wi_new(C.byref(self._wi), _f(w), w.size)
        self.ok = self.rc == OK

    def update(self, pairs):
        idx = (C.c_size_t * len(pairs))(*[p[0] for p in pairs])
        w = f32([p[1] for p in pairs])
        return lib().fo_wi_update(C.byref(self._wi), idx, _f(w), len(pairs))

    def get_weight(self, i):
        return float(lib().fo_wi_get_weight(C.byref(self._wi), i))

    def pick(self, sample):
        return int(lib().fo_wi_pick(C.byref(self._wi), sample))

    def sample_value(self, u01):
        return float(lib().fo_wi_sample_value(C.byref(self._wi), u01))

    @property
    def total(self):
        return float(self._wi.total)

    def __del__(self):
        try:
            lib().fo_wi_free(C.byref(self._wi))
        except Exception:
            pass


# ---- kmeans ------------------------------------------------------------------------
def kmeans_init(block, k, first, chosen=None, u01=None, off=0, dim=None):
    v = view(block, off, dim)
    n, m = v.n, v.dim
    cent = np.zeros((k, m), np.float32)
    idx = np.zeros(n, np.uint32)
    w = np.zeros(n, np.float32)
    picked = np.zeros(k, np.uint32)
    ch = None if chosen is None else np.ascontiguousarray(chosen, np.uint32)
    u = None if u01 is None else f32(u01)
    rc = lib().fo_kmeans_init(C.byref(v), k, first, None if ch is None else _u(ch),
                              None if u is None else _f(u), _f(cent), _u(idx), _f(w), _u(picked))
    return rc, cent, idx, w, picked


def kmeans_update(block, k, centroids, indices, off=0, dim=None):
    v = view(block, off, dim)
    cent = f32(centroids).copy()
    idx = np.ascontiguousarray(indices, np.uint32)
    g = C.c_float()
    rc = lib().fo_kmeans_update(C.byref(v), k, _f(cent), _u(idx), C.byref(g))
    return rc, cent, float(g.value)


def kmeans_reassign(block, k, centroids, off=0, dim=None, nthreads=1, native=False):
    v = view(block, off, dim)
    cent = f32(centroids)
    idx = np.zeros(v.n, np.uint32)
    rc = lib(native).fo_kmeans_reassign(C.byref(v), k, _f(cent), _u(idx), nthreads)
    return rc, idx


def kmeans_lloyd(block, k, centroids, indices, max_rounds=100, eps=1e-6, off=0, dim=None,
                 nthreads=1, native=False):
    v = view(block, off, dim)
    cent = f32(centroids).copy()
    idx = np.ascontiguousarray(indices, np.uint32).copy()
    grads = np.zeros(max_rounds, np.float32)
    nu, nr = C.c_size_t(), C.c_size_t()
    rc = lib(native).fo_kmeans_lloyd(C.byref(v), k, _f(cent), _u(idx), max_rounds, eps, _f(grads),
                                     C.byref(nu), C.byref(nr), nthreads)
    return rc, cent, idx, grads[:nu.value].copy(), nr.value


def residues(block, centroids, indices):
    """In place, like Partitioning::partition_with_events."""
    v = view(block)
    cent = f32(centroids)
    idx = np.ascontiguousarray(indices, np.uint32)
    lib().fo_residues(C.byref(v), cent.shape[0], _f(cent), _u(idx))


def extract_partitions(part_idx, codes_div_major, P):
    part_idx = np.ascontiguousarray(part_idx, np.uint32)
    codes = np.ascontiguousarray(codes_div_major, np.uint32)
    D, M = codes.shape
    offsets = np.zeros(P + 1, np.uint64)
    order = np.zeros(M, np.uint32)
    pm = np.zeros((M, D), np.uint32)
    lib().fo_extract_partitions(M, P, D, _u(part_idx), _u(codes),
                                offsets.ctypes.data_as(U64P), _u(order), _u(pm))
    return offsets, order, pm


def nbest(keys, n):
    """Push items (key, position) in order; return surviving positions in slot order."""
    items = (Item * max(n, 1))()
    nb = NBest(items, n, 0)
    for i, k in enumerate(keys):
        lib().fo_nbest_push(C.byref(nb), Item(float(k), i, 0))
    return [int(items[i].a) for i in range(nb.len)]


# ---- build (src/db/build.rs:78-129), driven step by step -----------------------------
def build_database(data, P, D, Cn, seeds, max_rounds=100, nthreads=1, native=False):
    """seeds: dict with 'coarse': (first, u01[P-1]) and 'pq': [(first, u01[C-1])]*D.
    Returns dict(coarse, part_idx, residues, codebooks[D,C,s], codes[D,M], events)."""
    data = f32(data).copy()
    M, N = data.shape
    events = {}
    first, u = seeds["coarse"]
    rc, cent, idx, _, _ = kmeans_init(data, P, first, u01=u)
    if rc:
        return rc, None
    rc, cent, idx, grads, nr = kmeans_lloyd(data, P, cent, idx, max_rounds, nthreads=nthreads,
                                            native=native)
    if rc:
        return rc, None
    events["coarse"] = (grads, nr)
    residues(data, cent, idx)
    s = N // D
    cbs = np.zeros((D, Cn, s), np.float32)
    codes = np.zeros((D, M), np.uint32)
    events["pq"] = []
    for di in range(D):
        first, u = seeds["pq"][di]
        rc, c, i, _, _ = kmeans_init(data, Cn, first, u01=u, off=di * s, dim=s)
        if rc:
            return rc, None
        rc, c, i, g, nr = kmeans_lloyd(data, Cn, c, i, max_rounds, off=di * s, dim=s,
                                       nthreads=nthreads, native=native)
        if rc:
            return rc, None
        cbs[di], codes[di] = c, i
        events["pq"].append((g, nr))
    return OK, dict(coarse=cent, part_idx=idx, residues=data, codebooks=cbs, codes=codes,
                    events=events)


# ---- query ----------------------------------------------------------------------------
class QueryIndex:
    def __init__(self, coarse, codebooks, offsets, codes_pm):
        self.coarse = f32(coarse)
        self.codebooks = f32(codebooks)
        self.offsets = np.ascontiguousarray(offsets, np.uint64)
        self.codes_pm = np.ascontiguousarray(codes_pm, np.uint32)
        P, N = self.coarse.shape
        D, Cn, _ = self.codebooks.shape
        self.ix = Index(N, P, D, Cn, _f(self.coarse), _f(self.codebooks),
                        self.offsets.ctypes.data_as(U64P), _u(self.codes_pm))
        self.N, self.P, self.D, self.C = N, P, D, Cn

    def query(self, q, k, nprobe, mode=0, nthreads=1, native=False):
        q = f32(q).reshape(-1, self.N)
        nq = q.shape[0]
        part = np.zeros((nq, k), np.uint32)
        vidx = np.zeros((nq, k), np.uint32)
        dist = np.zeros((nq, k), np.float32)
        cnt = np.zeros(nq, np.uint32)
        rc = lib(native).fo_query(C.byref(self.ix), _f(q), nq, k, nprobe, mode, _u(part), _u(vidx),
                                  _f(dist), _u(cnt), nthreads)
        return rc, part, vidx, dist, cnt

    def probe(self, q, nprobe, mode=0):
        q = f32(q).reshape(self.N)
        part = np.zeros(nprobe, np.uint32)
        dist = np.zeros(nprobe, np.float32)
        rc = lib().fo_query_probe(C.byref(self.ix), _f(q), nprobe, mode, _u(part), _f(dist))
        return rc, part, dist

    def table(self, q, part):
        q = f32(q).reshape(self.N)
        t = np.zeros((self.D, self.C), np.float32)
        lib().fo_query_table(C.byref(self.ix), _f(q), int(part), _f(t))
        return t


# ---- synthetic data -------------------------------------------------------------------
def fill_uniform(count, seed, start=0):
    out = np.empty(count, np.float32)
    lib().fo_fill_uniform(_f(out), count, seed, start)
    return out


def fill_uniform_np(count, seed, start=0):
    """numpy twin of fo_fill_uniform (same bits), used to cross-check the C code."""
    i = np.arange(start, start + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (i + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)).astype(np.float32)
