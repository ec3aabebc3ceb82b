//! `extern "C"` bindings of include/flechasdb_b200.h (new file `src/ffi.rs` in the reference
//! crate).  NOT COMPILED HERE (no Rust toolchain in this environment); kept in lock-step with the
//! header by hand -- `tests/test_abi.py` checks that every symbol named here is exported.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_float, c_int};

#[repr(C)] pub struct fdb_ctx { _p: [u8; 0] }
#[repr(C)] pub struct fdb_vs { _p: [u8; 0] }
#[repr(C)] pub struct fdb_km { _p: [u8; 0] }
#[repr(C)] pub struct fdb_index { _p: [u8; 0] }

pub const FDB_OK: c_int = 0;
pub const FDB_ERR_INVALID_ARGS: c_int = -1;
pub const FDB_ERR_INVALID_DATA: c_int = -2;
pub const FDB_ERR_INVALID_CONTEXT: c_int = -3;
pub const FDB_ERR_EMPTY_CLUSTER: c_int = -4;
pub const FDB_ERR_WEIGHTS: c_int = -5;
pub const FDB_ERR_NAN: c_int = -6;
pub const FDB_QUERY_STORED: c_int = 0;
pub const FDB_QUERY_BUILD: c_int = 1;
pub const FDB_KMEANS_MAX_ROUNDS: usize = 100;

extern "C" {
    pub fn fdb_last_error() -> *const c_char;
    pub fn fdb_ctx_create(device: c_int, out: *mut *mut fdb_ctx) -> c_int;
    pub fn fdb_ctx_destroy(ctx: *mut fdb_ctx);
    pub fn fdb_vs_upload(ctx: *mut fdb_ctx, rows: *const c_float, n: usize, dim: usize, out: *mut *mut fdb_vs) -> c_int;
    pub fn fdb_vs_download(vs: *mut fdb_vs, rows: *mut c_float) -> c_int;
    pub fn fdb_vs_destroy(vs: *mut fdb_vs);
    pub fn fdb_vs_subtract_assigned(vs: *mut fdb_vs, km: *const fdb_km) -> c_int;
    pub fn fdb_kmeans_begin(vs: *mut fdb_vs, col_off: usize, dim: usize, nb: usize, k: usize, out: *mut *mut fdb_km) -> c_int;
    pub fn fdb_kmeans_destroy(km: *mut fdb_km);
    pub fn fdb_kmeans_seed_first(km: *mut fdb_km, ci: *const u32) -> c_int;
    pub fn fdb_kmeans_seed_pick(km: *mut fdb_km, u01: *const c_float, exact: c_int, ci_out: *mut u32) -> c_int;
    pub fn fdb_kmeans_seed_add(km: *mut fdb_km, i: usize, ci: *const u32, exact: c_int) -> c_int;
    pub fn fdb_kmeans_seed_run(km: *mut fdb_km, first: *const u32, u01: *const c_float, exact: c_int, picked: *mut u32) -> c_int;
    pub fn fdb_kmeans_update(km: *mut fdb_km, active: *const u8, gradients: *mut c_float) -> c_int;
    pub fn fdb_kmeans_reassign(km: *mut fdb_km, active: *const u8) -> c_int;
    pub fn fdb_kmeans_run(km: *mut fdb_km, max_rounds: usize, epsilon: c_float, gradients: *mut c_float, rounds: *mut u32, reassigns: *mut u32) -> c_int;
    pub fn fdb_kmeans_get(km: *mut fdb_km, centroids: *mut c_float, indices: *mut u32) -> c_int;
    pub fn fdb_index_create(ctx: *mut fdb_ctx, n: usize, p: usize, d: usize, c: usize, coarse: *const c_float,
                            codebooks: *const c_float, offsets: *const u64, codes: *const u8, out: *mut *mut fdb_index) -> c_int;
    pub fn fdb_index_from_build(ctx: *mut fdb_ctx, coarse: *const fdb_km, pq: *const fdb_km, out: *mut *mut fdb_index) -> c_int;
    pub fn fdb_index_get_layout(ix: *mut fdb_index, offsets: *mut u64, order: *mut u32, codes: *mut u8) -> c_int;
    pub fn fdb_index_query(ix: *mut fdb_index, queries: *const c_float, nq: usize, k: usize, nprobe: usize, mode: c_int,
                           out_partition: *mut u32, out_vector_index: *mut u32, out_sqdist: *mut c_float, out_count: *mut u32) -> c_int;
    pub fn fdb_index_probe(ix: *mut fdb_index, queries: *const c_float, nq: usize, nprobe: usize, mode: c_int,
                           out_partition: *mut u32, out_sqdist: *mut c_float) -> c_int;
    pub fn fdb_index_destroy(ix: *mut fdb_index);

    // rows sharded over several GPUs (one process each): stage functions enqueued on fdb_ctx_stream, the
    // caller enqueues ncclAllGather / ncclAllReduce on the same stream in between
    pub fn fdb_kmeans_seed_sharded_first(km: *mut fdb_km, local_first: *const u32) -> c_int;
    pub fn fdb_kmeans_seed_sharded_total(km: *mut fdb_km) -> c_int;
    pub fn fdb_kmeans_seed_sharded_pick(km: *mut fdb_km, d_all_totals: *const c_float, world: c_int, rank: c_int) -> c_int;
    pub fn fdb_kmeans_seed_sharded_finish(km: *mut fdb_km, picked_global: *mut u32) -> c_int;
    pub fn fdb_kmeans_sharded_loop_begin(km: *mut fdb_km) -> c_int;
    pub fn fdb_kmeans_sharded_partial_async(km: *mut fdb_km, device_buf: *mut *mut c_float, nfloats: *mut usize) -> c_int;
    pub fn fdb_kmeans_sharded_finish_async(km: *mut fdb_km, epsilon: c_float) -> c_int;
    pub fn fdb_kmeans_sharded_poll(km: *mut fdb_km, active: *mut u8) -> c_int;
    pub fn fdb_kmeans_sharded_loop_end(km: *mut fdb_km, gradients: *mut c_float, rounds: *mut u32, reassigns: *mut u32) -> c_int;

    // code lists sharded over several GPUs: per-rank query on device buffers, all-gather, merge on the device
    pub fn fdb_index_query_device(ix: *mut fdb_index, d_queries: *const c_float, nq: usize, k: usize, nprobe: usize, mode: c_int,
                                  d_partition: *mut u32, d_vector_index: *mut u32, d_sqdist: *mut c_float, d_count: *mut u32) -> c_int;
    pub fn fdb_index_probe_device(ix: *mut fdb_index, d_queries: *const c_float, nq: usize, nprobe: usize, mode: c_int, d_partition: *mut u32) -> c_int;
    pub fn fdb_index_last_probes_device(ix: *mut fdb_index, nq: usize, nprobe: usize, d_partition: *mut u32) -> c_int;
    pub fn fdb_merge_topk_device(ctx: *mut fdb_ctx, world: c_int, nq: usize, k: usize, nprobe: usize, d_partition: *const u32,
                                 d_vector_index: *const u32, d_sqdist: *const c_float, d_count: *const u32, d_probes: *const u32,
                                 d_out_partition: *mut u32, d_out_vector_index: *mut u32, d_out_sqdist: *mut c_float,
                                 d_out_count: *mut u32, d_tie_flag: *mut u32) -> c_int;
}

/// Maps a status of the C ABI onto the reference's error convention: `Err(Error::…)` for
/// argument / data problems, a panic for the invariant violations the reference panics on.
pub fn check(rc: c_int) -> Result<(), crate::error::Error> {
    use crate::error::Error;
    if rc == FDB_OK { return Ok(()); }
    let msg = unsafe { std::ffi::CStr::from_ptr(fdb_last_error()) }.to_string_lossy().into_owned();
    match rc {
        FDB_ERR_INVALID_ARGS => Err(Error::InvalidArgs(msg)),
        FDB_ERR_INVALID_DATA => Err(Error::InvalidData(msg)),
        FDB_ERR_INVALID_CONTEXT => Err(Error::InvalidContext(msg)),
        FDB_ERR_EMPTY_CLUSTER | FDB_ERR_WEIGHTS | FDB_ERR_NAN => panic!("{}", msg), // as the reference does
        _ => Err(Error::InvalidContext(msg)),
    }
}
