//! Replacement body for `stored::Database::<f32, FS>::query_with_events` (reference src/db/stored.rs:331-389) and its
//! lazy `get_partition` (src/db/stored.rs:269-293): the partition centroids and the codebooks go to the device once
//! (what load_partition_centroids / load_codebook return), a partition's codes when a query first probes it.
//! Attributes (src/db/stored.rs:108-260) are untouched: they never leave the host.
//! Public signatures are unchanged.  NOT COMPILED HERE (no Rust toolchain); the same call sequence is compiled and
//! tested in C++ (flechasdb_b200/host/flechasdb_stored.hpp) and Python (flechasdb_b200/stored.py).
use core::cell::{Cell, RefCell};

use crate::db::stored::{Database, LoadCodebook, LoadPartition, LoadPartitionCentroids, QueryEvent, QueryResult};
use crate::error::Error;
use crate::ffi::*;
use crate::io::FileSystem;
use crate::vector::VectorSet;

/// Device side of a stored database: lives next to `partitions: RefCell<Vec<Option<Partition<f32>>>>`.
pub struct DeviceIndex {
    pub ctx: *mut fdb_ctx,
    pub index: Cell<*mut fdb_index>,          // null until the first query (lazy, like the codebooks)
    pub vector_ids: RefCell<Vec<Option<Vec<uuid::Uuid>>>>,   // per partition, filled when the partition is uploaded
}

impl<FS> Database<f32, FS>
where
    FS: FileSystem,
    Self: LoadPartition<f32> + LoadCodebook<f32> + LoadPartitionCentroids<f32>,
{
    /// src/db/stored.rs:343-360 -- the lazy loads, plus the upload of what they return
    fn device_index(&self, dev: &DeviceIndex) -> Result<*mut fdb_index, Error> {
        if !dev.index.get().is_null() {
            return Ok(dev.index.get());
        }
        let centroids = self.load_partition_centroids()?;            // BlockVectorSet<f32>, P x N
        let mut codebooks: Vec<f32> = Vec::with_capacity(self.num_divisions() * self.num_codes() * self.subvector_size());
        for di in 0..self.num_divisions() {
            let cb = self.load_codebook(di)?;                         // BlockVectorSet<f32>, C x N/D
            for ci in 0..cb.len() { codebooks.extend_from_slice(cb.get(ci)); }
        }
        let mut flat: Vec<f32> = Vec::with_capacity(self.num_partitions() * self.vector_size());
        for pi in 0..centroids.len() { flat.extend_from_slice(centroids.get(pi)); }
        let mut ix = core::ptr::null_mut();
        check(unsafe { fdb_index_create_lazy(dev.ctx, self.vector_size(), self.num_partitions(), self.num_divisions(),
                                             self.num_codes(), flat.as_ptr(), codebooks.as_ptr(), &mut ix) })?;
        dev.index.set(ix);
        Ok(ix)
    }

    /// get_partition (src/db/stored.rs:269-293): read, validate and upload partition `pi` once
    fn upload_partition(&self, dev: &DeviceIndex, ix: *mut fdb_index, pi: usize) -> Result<(), Error> {
        if unsafe { fdb_index_partition_loaded(ix, pi) } != 0 {
            return Ok(());
        }
        let partition = self.load_partition(pi)?;                    // src/db/stored.rs:800-880
        let n = partition.encoded_vectors.len();
        let mut codes: Vec<u8> = Vec::with_capacity(n * self.num_divisions());
        for vi in 0..n {
            for &c in partition.encoded_vectors.get(vi) {
                if c as usize >= self.num_codes() || c > 255 {
                    return Err(Error::InvalidData(format!("partition {} holds the code {}", pi, c)));
                }
                codes.push(c as u8);
            }
        }
        check(unsafe { fdb_index_set_partition(ix, pi, codes.as_ptr(), n) })?;
        dev.vector_ids.borrow_mut()[pi] = Some(partition.vector_ids.clone());
        Ok(())
    }

    pub fn query_with_events_b200<'a, EH>(&'a self, dev: &DeviceIndex, v: &[f32], k: usize, nprobe: usize, mut event: EH)
        -> Result<Vec<QueryResult<'a, f32, FS>>, Error>
    where EH: FnMut(QueryEvent) -> ()
    {
        event(QueryEvent::StartingQueryInitialization);
        let ix = self.device_index(dev)?;
        event(QueryEvent::FinishedQueryInitialization);
        event(QueryEvent::StartingPartitionSelection);
        let mut probes = vec![0u32; nprobe];
        // Err(InvalidArgs) "nprobe {} exceeds the number of partitions {}" (src/db/stored.rs:403-409) comes back from the library
        check(unsafe { fdb_index_probe(ix, v.as_ptr(), 1, nprobe, FDB_QUERY_STORED, probes.as_mut_ptr(), core::ptr::null_mut()) })?;
        event(QueryEvent::FinishedPartitionSelection);
        for &p in &probes {
            event(QueryEvent::StartingPartitionQuery(p as usize));
            self.upload_partition(dev, ix, p as usize)?;             // lazy: the partition's file is read on its first probe
            event(QueryEvent::FinishedPartitionQuery(p as usize));
        }
        let (mut part, mut vidx, mut dist) = (vec![0u32; k], vec![0u32; k], vec![0f32; k]);
        let mut count = 0u32;
        check(unsafe { fdb_index_query(ix, v.as_ptr(), 1, k, nprobe, FDB_QUERY_STORED, part.as_mut_ptr(), vidx.as_mut_ptr(),
                                       dist.as_mut_ptr(), &mut count) })?;
        event(QueryEvent::StartingResultSelection);
        let ids = dev.vector_ids.borrow();
        let out = (0..count as usize).map(|i| QueryResult::new(
            self,
            part[i] as usize,
            ids[part[i] as usize].as_ref().expect("probed partitions are loaded")[vidx[i] as usize],
            vidx[i] as usize,
            dist[i],
        )).collect();
        event(QueryEvent::FinishedResultSelection);
        Ok(out)
    }
}

/// Rows sharded over several GPUs (no reference analogue): one rank per GPU, the library issues
/// the NCCL collectives itself.  `id` is made by rank 0 with fdb_comm_unique_id and handed over by the host.
pub fn cluster_sharded(km: *mut fdb_km, ctx: *mut fdb_ctx, world: i32, rank: i32, id: &[u8; FDB_COMM_ID_BYTES],
                       n_global: usize, first: &[u32], u01: &[f32], k: usize, nb: usize)
    -> Result<(Vec<u32>, Vec<f32>, Vec<u32>), Error>
{
    let mut comm = core::ptr::null_mut();
    check(unsafe { fdb_comm_create(ctx, world, rank, id.as_ptr(), &mut comm) })?;
    let mut picked = vec![0u32; nb * k];
    check(unsafe { fdb_kmeans_seed_run_sharded(km, comm, n_global, first.as_ptr(), u01.as_ptr(), picked.as_mut_ptr()) })?;
    let mut grads = vec![0f32; nb * FDB_KMEANS_MAX_ROUNDS];
    let (mut rounds, mut reas) = (vec![0u32; nb], vec![0u32; nb]);
    check(unsafe { fdb_kmeans_run_sharded(km, comm, FDB_KMEANS_MAX_ROUNDS, FDB_KMEANS_EPSILON, grads.as_mut_ptr(),
                                          rounds.as_mut_ptr(), reas.as_mut_ptr()) })?;
    unsafe { fdb_comm_destroy(comm) };
    Ok((picked, grads, rounds))
}
