//! Replacement bodies for `DatabaseBuilder::build_with_events` (reference src/db/build.rs:78-129),
//! `Partitioning::partition_with_events` (src/partitions.rs:119-143) and
//! `build::Database::query_with_events` (src/db/build.rs:307-340) for f32 / BlockVectorSet<f32>.
//! Public signatures are unchanged.  NOT COMPILED HERE (no Rust toolchain).
use crate::db::build::{BuildEvent, QueryEvent, QueryResult};
use crate::error::Error;
use crate::ffi::*;
use crate::kmeans_b200::{cluster_device, DeviceVectorSet};

pub fn build_with_events_b200<EH>(data: Vec<f32>, vector_size: usize, p: usize, d: usize, c: usize,
                                  mut event: EH) -> Result<Built, Error>
where EH: FnMut(BuildEvent<'_, f32>) -> ()
{
    let m = data.len() / vector_size;
    event(BuildEvent::StartingIdAssignment);
    let vector_ids: Vec<uuid::Uuid> = (0..m).map(|_| uuid::Uuid::new_v4()).collect();
    event(BuildEvent::FinishedIdAssignment);
    let mut ctx = core::ptr::null_mut();
    check(unsafe { fdb_ctx_create(0, &mut ctx) })?;
    let mut vs = core::ptr::null_mut();
    check(unsafe { fdb_vs_upload(ctx, data.as_ptr(), m, vector_size, &mut vs) })?;
    let dvs = DeviceVectorSet { ctx, vs, n: m, dim: vector_size };
    // partitions all the data: k-means with P clusters, then residues in place
    event(BuildEvent::StartingPartitioning);
    let (coarse_km, coarse) = cluster_device(&dvs, 0, vector_size, 1, p.try_into().unwrap(),
                                             |_, e| event(BuildEvent::ClusterEvent(e)))?;
    check(unsafe { fdb_vs_subtract_assigned(vs, coarse_km) })?;
    event(BuildEvent::FinishedPartitioning);
    event(BuildEvent::StartingSubvectorDivision);
    if vector_size % d != 0 {
        return Err(Error::InvalidArgs(format!("vector size ({}) is not divisible by {}", vector_size, d)));
    }
    event(BuildEvent::FinishedSubvectorDivision);
    // all D divisions side by side; events are replayed per division in the reference's order
    let mut last = usize::MAX;
    let (pq_km, codebooks) = cluster_device(&dvs, 0, vector_size / d, d, c.try_into().unwrap(), |di, e| {
        if di != last {
            if last != usize::MAX { event(BuildEvent::FinishedQuantization(last)); }
            event(BuildEvent::StartingQuantization(di));
            last = di;
        }
        event(BuildEvent::ClusterEvent(e));
    })?;
    if last != usize::MAX { event(BuildEvent::FinishedQuantization(last)); }
    let mut index = core::ptr::null_mut();
    check(unsafe { fdb_index_from_build(ctx, coarse_km, pq_km, &mut index) })?;
    Ok(Built { ctx, vs, coarse_km, pq_km, index, vector_ids, coarse, codebooks })
}

pub struct Built {
    pub ctx: *mut fdb_ctx, pub vs: *mut fdb_vs, pub coarse_km: *mut fdb_km, pub pq_km: *mut fdb_km,
    pub index: *mut fdb_index, pub vector_ids: Vec<uuid::Uuid>,
    pub coarse: Vec<crate::kmeans::Codebook<f32>>, pub codebooks: Vec<crate::kmeans::Codebook<f32>>,
}

pub fn query_with_events_b200<EH>(built: &Built, offsets: &[u64], order: &[u32], v: &[f32], k: usize,
                                  nprobe: usize, mode: i32, mut event: EH)
    -> Result<Vec<QueryResult<f32>>, Error>
where EH: FnMut(QueryEvent) -> ()
{
    event(QueryEvent::StartingPartitionSelection);
    let mut probes = vec![0u32; nprobe];
    // Err(InvalidArgs) "nprobe {} exceeds the number of partitions {}" comes back from the library
    check(unsafe { fdb_index_probe(built.index, v.as_ptr(), 1, nprobe, mode, probes.as_mut_ptr(), core::ptr::null_mut()) })?;
    event(QueryEvent::FinishedPartitionSelection);
    let (mut part, mut vidx, mut dist) = (vec![0u32; k], vec![0u32; k], vec![0f32; k]);
    let mut count = 0u32;
    check(unsafe { fdb_index_query(built.index, v.as_ptr(), 1, k, nprobe, mode, part.as_mut_ptr(),
                                   vidx.as_mut_ptr(), dist.as_mut_ptr(), &mut count) })?;
    for &p in &probes {
        event(QueryEvent::StartingPartitionQuery(p as usize));
        event(QueryEvent::FinishedPartitionQuery(p as usize));
    }
    event(QueryEvent::StartingResultSelection);
    let out = (0..count as usize).map(|i| QueryResult {
        partition_index: part[i] as usize,
        vector_id: built.vector_ids[order[offsets[part[i] as usize] as usize + vidx[i] as usize] as usize],
        vector_index: vidx[i] as usize,
        squared_distance: dist[i],
    }).collect();
    event(QueryEvent::FinishedResultSelection);
    Ok(out)
}
