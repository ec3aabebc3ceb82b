//! Replacement body of `kmeans::cluster_with_events` for `T = f32, VS = BlockVectorSet<f32>`
//! (reference src/kmeans.rs:104-139).  Same signature, same event sequence, same errors; the
//! loops run on the GPU through the C ABI.  NOT COMPILED HERE (no Rust toolchain).
use core::num::NonZeroUsize;
use rand::Rng;

use crate::error::Error;
use crate::ffi::*;
use crate::kmeans::{ClusterEvent, Codebook};
use crate::vector::BlockVectorSet;

pub struct DeviceVectorSet { pub ctx: *mut fdb_ctx, pub vs: *mut fdb_vs, pub n: usize, pub dim: usize }

/// `nb` problems over the strided views [col_off + b*dim, +dim) — nb = 1 is cluster_with_events,
/// nb = D runs the per-division loop of DatabaseBuilder::build_with_events in one go.
pub fn cluster_device<EV>(dvs: &DeviceVectorSet, col_off: usize, dim: usize, nb: usize,
                          k: NonZeroUsize, mut event_handler: EV)
    -> Result<(*mut fdb_km, Vec<Codebook<f32>>), Error>
where EV: FnMut(usize, ClusterEvent<'_, f32>) -> ()
{
    let k = k.get();
    let mut km: *mut fdb_km = core::ptr::null_mut();
    // Err(InvalidArgs) "vs has fewer vectors than k" comes back from the library (:116-120)
    check(unsafe { fdb_kmeans_begin(dvs.vs, col_off, dim, nb, k, &mut km) })?;
    // the draws initialize_centroids takes from thread_rng (:148,172,202)
    let mut rng = rand::thread_rng();
    let first: Vec<u32> = (0..nb).map(|_| rng.gen_range(0..dvs.n) as u32).collect();
    let u01: Vec<f32> = (0..nb * (k - 1)).map(|_| (rng.gen::<u32>() >> 9) as f32 * (1.0 / 8388608.0)).collect();
    check(unsafe { fdb_kmeans_seed_run(km, first.as_ptr(), u01.as_ptr(), 0, core::ptr::null_mut()) })?;
    let mut grads = vec![0f32; nb * FDB_KMEANS_MAX_ROUNDS];
    let (mut rounds, mut reassigns) = (vec![0u32; nb], vec![0u32; nb]);
    check(unsafe { fdb_kmeans_run(km, FDB_KMEANS_MAX_ROUNDS, 1e-6, grads.as_mut_ptr(),
                                  rounds.as_mut_ptr(), reassigns.as_mut_ptr()) })?;
    let mut centroids = vec![0f32; nb * k * dim];
    let mut indices = vec![0u32; nb * dvs.n];
    check(unsafe { fdb_kmeans_get(km, centroids.as_mut_ptr(), indices.as_mut_ptr()) })?;
    let mut out = Vec::with_capacity(nb);
    for b in 0..nb {
        // replay the reference's event order for problem b (:121-137)
        event_handler(b, ClusterEvent::StartingCentroidInitialization);
        event_handler(b, ClusterEvent::FinishedCentroidInitialization);
        for r in 0..rounds[b] as usize {
            event_handler(b, ClusterEvent::StartingCentroidUpdate(r));
            event_handler(b, ClusterEvent::FinishedCentroidUpdate(r, &grads[b * FDB_KMEANS_MAX_ROUNDS + r]));
            if r < reassigns[b] as usize {
                event_handler(b, ClusterEvent::StartingCentroidReassignment(r));
                event_handler(b, ClusterEvent::FinishedCentroidReassignment(r));
            }
        }
        out.push(Codebook {
            centroids: BlockVectorSet::chunk(centroids[b * k * dim..(b + 1) * k * dim].to_vec(),
                                             dim.try_into().unwrap()).unwrap(),
            indices: indices[b * dvs.n..(b + 1) * dvs.n].iter().map(|&i| i as usize).collect(),
        });
    }
    Ok((km, out))
}
