"""ctypes bindings of include/flechasdb_b200.h (libflechasdb_b200.so).

This is the only way Python reaches the engine: the same C symbols the Rust
`extern "C"` block of INTEGRATION.md binds.  There is no CPU fallback -- if the
shared library is missing, loading raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libflechasdb_b200.so")

OK = 0
ERR_INVALID_ARGS = -1
ERR_INVALID_DATA = -2
ERR_INVALID_CONTEXT = -3
ERR_EMPTY_CLUSTER = -4
ERR_WEIGHTS = -5
ERR_NAN = -6
ERR_CUDA = -7
ERR_UNSUPPORTED = -8
ERR_NCCL = -9
COMM_ID_BYTES = 128

QUERY_STORED = 0
QUERY_BUILD = 1
KMEANS_MAX_ROUNDS = 100
KMEANS_EPSILON = 1e-6

F32P = C.POINTER(C.c_float)
U8P = C.POINTER(C.c_uint8)
U32P = C.POINTER(C.c_uint32)
U64P = C.POINTER(C.c_uint64)
VP = C.c_void_p
SZ = C.c_size_t

# every symbol include/flechasdb_b200.h declares: (restype, argtypes)
SIGNATURES = {
    "fdb_last_error": (C.c_char_p, []),
    "fdb_version": (C.c_int, []),
    "fdb_device_count": (C.c_int, []),
    "fdb_ctx_create": (C.c_int, [C.c_int, C.POINTER(VP)]),
    "fdb_ctx_destroy": (None, [VP]),
    "fdb_ctx_sync": (C.c_int, [VP]),
    "fdb_ctx_timer_start": (C.c_int, [VP]),
    "fdb_ctx_timer_stop": (C.c_int, [VP, F32P]),
    "fdb_ctx_launch_count": (C.c_uint64, [VP]),
    "fdb_vs_upload": (C.c_int, [VP, F32P, SZ, SZ, C.POINTER(VP)]),
    "fdb_vs_from_device": (C.c_int, [VP, VP, SZ, SZ, C.POINTER(VP)]),
    "fdb_vs_generate": (C.c_int, [VP, SZ, SZ, C.c_uint64, C.c_uint64, C.POINTER(VP)]),
    "fdb_vs_download": (C.c_int, [VP, F32P]),
    "fdb_vs_download_rows": (C.c_int, [VP, SZ, SZ, F32P]),
    "fdb_vs_len": (SZ, [VP]),
    "fdb_vs_vector_size": (SZ, [VP]),
    "fdb_vs_device_ptr": (VP, [VP]),
    "fdb_vs_destroy": (None, [VP]),
    "fdb_vs_subtract_assigned": (C.c_int, [VP, VP]),
    "fdb_kmeans_begin": (C.c_int, [VP, SZ, SZ, SZ, SZ, C.POINTER(VP)]),
    "fdb_kmeans_destroy": (None, [VP]),
    "fdb_kmeans_seed_first": (C.c_int, [VP, U32P]),
    "fdb_kmeans_seed_total": (C.c_int, [VP, F32P]),
    "fdb_kmeans_seed_pick": (C.c_int, [VP, F32P, C.c_int, U32P]),
    "fdb_kmeans_seed_add": (C.c_int, [VP, SZ, U32P, C.c_int]),
    "fdb_kmeans_seed_run": (C.c_int, [VP, U32P, F32P, C.c_int, U32P]),
    "fdb_kmeans_seed_chosen": (C.c_int, [VP, U32P]),
    "fdb_kmeans_seed_round_ext": (C.c_int, [VP, SZ, F32P, U32P]),
    "fdb_kmeans_seed_pick_value": (C.c_int, [VP, F32P, U32P]),
    "fdb_ctx_stream": (VP, [VP]),
    "fdb_kmeans_seed_sharded_begin": (C.c_int, [VP, C.POINTER(VP), C.POINTER(VP), C.POINTER(VP), C.POINTER(VP),
                                                C.POINTER(VP)]),
    "fdb_kmeans_seed_sharded_first": (C.c_int, [VP, U32P]),
    "fdb_kmeans_seed_sharded_total": (C.c_int, [VP]),
    "fdb_kmeans_seed_sharded_pick": (C.c_int, [VP, VP, C.c_int, C.c_int]),
    "fdb_kmeans_seed_sharded_round": (C.c_int, [VP, SZ, VP, VP, C.c_int, C.c_int, SZ]),
    "fdb_kmeans_seed_sharded_finish": (C.c_int, [VP, U32P]),
    "fdb_kmeans_set_state": (C.c_int, [VP, F32P, U32P]),
    "fdb_kmeans_update": (C.c_int, [VP, U8P, F32P]),
    "fdb_kmeans_reassign": (C.c_int, [VP, U8P]),
    "fdb_kmeans_run": (C.c_int, [VP, SZ, C.c_float, F32P, U32P, U32P]),
    "fdb_kmeans_last_assign_info": (C.c_int, [VP, U32P, U32P, U32P]),
    "fdb_kmeans_get": (C.c_int, [VP, F32P, U32P]),
    "fdb_kmeans_get_weights": (C.c_int, [VP, F32P]),
    "fdb_kmeans_update_partial": (C.c_int, [VP, C.POINTER(VP), C.POINTER(SZ)]),
    "fdb_kmeans_sharded_loop_begin": (C.c_int, [VP]),
    "fdb_kmeans_sharded_partial_async": (C.c_int, [VP, C.POINTER(VP), C.POINTER(SZ)]),
    "fdb_kmeans_sharded_finish_async": (C.c_int, [VP, C.c_float]),
    "fdb_kmeans_sharded_poll": (C.c_int, [VP, U8P]),
    "fdb_kmeans_sharded_loop_end": (C.c_int, [VP, F32P, U32P, U32P]),
    "fdb_kmeans_update_finish": (C.c_int, [VP, F32P]),
    "fdb_index_probe_device": (C.c_int, [VP, VP, SZ, SZ, C.c_int, VP]),
    "fdb_index_last_probes_device": (C.c_int, [VP, SZ, SZ, VP]),
    "fdb_merge_topk_device": (C.c_int, [VP, C.c_int, SZ, SZ, SZ, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP]),
    "fdb_index_create": (C.c_int, [VP, SZ, SZ, SZ, SZ, F32P, F32P, U64P, U8P, C.POINTER(VP)]),
    "fdb_index_create_lazy": (C.c_int, [VP, SZ, SZ, SZ, SZ, F32P, F32P, C.POINTER(VP)]),
    "fdb_index_set_partition": (C.c_int, [VP, SZ, U8P, SZ]),
    "fdb_index_partition_loaded": (C.c_int, [VP, SZ]),
    "fdb_index_missing_partitions": (C.c_int, [VP, F32P, SZ, SZ, C.c_int, U32P, SZ, C.POINTER(SZ)]),
    "fdb_index_from_build": (C.c_int, [VP, VP, VP, C.POINTER(VP)]),
    "fdb_index_get_layout": (C.c_int, [VP, U64P, U32P, U8P]),
    "fdb_index_num_vectors": (SZ, [VP]),
    "fdb_index_destroy": (None, [VP]),
    "fdb_index_query": (C.c_int, [VP, F32P, SZ, SZ, SZ, C.c_int, U32P, U32P, F32P, U32P]),
    "fdb_index_query_device": (C.c_int, [VP, VP, SZ, SZ, SZ, C.c_int, VP, VP, VP, VP]),
    "fdb_index_probe": (C.c_int, [VP, F32P, SZ, SZ, C.c_int, U32P, F32P]),
    "fdb_index_table": (C.c_int, [VP, F32P, C.c_uint32, F32P]),
    "fdb_index_set_timing": (C.c_int, [VP, C.c_int]),
    "fdb_index_last_timing": (C.c_int, [VP, F32P, U64P]),
    "fdb_index_last_stats": (C.c_int, [VP, U64P]),
    "fdb_index_last_scan_kernel": (C.c_int, [VP, C.POINTER(C.c_int), F32P]),
    "fdb_index_debug_band": (C.c_int, [VP, SZ, SZ, F32P, F32P, U32P, U32P, U32P]),
    "fdb_comm_unique_id": (C.c_int, [U8P]),
    "fdb_comm_create": (C.c_int, [VP, C.c_int, C.c_int, U8P, C.POINTER(VP)]),
    "fdb_comm_destroy": (None, [VP]),
    "fdb_comm_world": (C.c_int, [VP]),
    "fdb_comm_rank": (C.c_int, [VP]),
    "fdb_comm_collective_count": (C.c_uint64, [VP]),
    "fdb_comm_allreduce_device": (C.c_int, [VP, VP, SZ]),
    "fdb_comm_allgather_device": (C.c_int, [VP, VP, VP, SZ]),
    "fdb_comm_max_f64": (C.c_int, [VP, C.POINTER(C.c_double), SZ]),
    "fdb_kmeans_seed_run_sharded": (C.c_int, [VP, VP, SZ, U32P, F32P, U32P]),
    "fdb_kmeans_run_sharded": (C.c_int, [VP, VP, SZ, C.c_float, F32P, U32P, U32P]),
    "fdb_index_query_sharded": (C.c_int, [VP, VP, VP, SZ, SZ, SZ, C.c_int, VP, VP, VP, VP]),
    "fdb_index_last_sharded_ties": (C.c_int, [VP, U32P]),
    "fdb_host_register": (C.c_int, [VP, VP, SZ]),
    "fdb_host_unregister": (C.c_int, [VP, VP]),
    "fdb_device_alloc": (C.c_int, [VP, SZ, C.POINTER(VP)]),
    "fdb_device_free": (C.c_int, [VP, VP]),
    "fdb_device_upload": (C.c_int, [VP, VP, VP, SZ]),
    "fdb_device_download": (C.c_int, [VP, VP, VP, SZ]),
    "fdb_device_fill_uniform": (C.c_int, [VP, VP, SZ, C.c_uint64, C.c_uint64]),
    "fdb_device_flush_l2": (C.c_int, [VP]),
}

_lib = None


class FdbError(RuntimeError):
    """A non-zero return of the C ABI: .code is the FDB_ERR_* value."""

    def __init__(self, code, message):
        super().__init__("fdb error %d: %s" % (code, message))
        self.code = code
        self.message = message


def lib():
    """Loads libflechasdb_b200.so.  Raises if it has not been built: no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m flechasdb_b200.build` "
            "(the engine is CUDA only, there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != OK:
        raise FdbError(rc, lib().fdb_last_error().decode(errors="replace"))


def f32p(a):
    return a.ctypes.data_as(F32P)


def u32p(a):
    return a.ctypes.data_as(U32P)


def u64p(a):
    return a.ctypes.data_as(U64P)


def u8p(a):
    return a.ctypes.data_as(U8P)


def as_f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def as_u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)
