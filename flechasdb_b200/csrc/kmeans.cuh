// Device state of nb side-by-side k-means problems (Codebook<f32>, src/kmeans.rs:62-68).
#pragma once
#include "common.cuh"

namespace fdb { struct TcState; }

struct fdb_km {
    fdb_ctx *ctx = nullptr;
    fdb_vs *vs = nullptr;
    size_t col_off = 0, m = 0, nb = 0, k = 0, n = 0;
    size_t max_rounds = FDB_KMEANS_MAX_ROUNDS;
    fdb::DevBuf<float> centroids, old_centroids;   // [nb][k][m]
    fdb::DevBuf<uint32_t> indices;                  // [nb][n]
    fdb::DevBuf<float> weights, weights_new;        // [nb][n] k-means++ D^2 weights (ping-pong)
    fdb::DevBuf<uint8_t> chosen;                    // [nb][n]
    fdb::DevBuf<uint32_t> members, members_tmp;     // [nb][n] rows grouped by cluster
    fdb::DevBuf<uint32_t> hist;                     // radix histograms
    fdb::DevBuf<uint32_t> cl_off;                   // [nb][k+1]
    fdb::DevBuf<float> cnorm, cdist;                // [nb][k]
    fdb::DevBuf<float> grad, grad_hist, total;      // [nb], [nb][max_rounds], [nb]
    fdb::DevBuf<uint32_t> rounds, reassigns, ci;    // [nb]
    fdb::DevBuf<int> active, step_active;           // [nb]
    fdb::DevBuf<float> partial;                     // multi-GPU sums ++ counts
    fdb::DevBuf<float> u01;                         // seeding draws staged on the device
    fdb::DevBuf<float> centre;                      // [nb][m] externally supplied centres (sharded rows)
    fdb::DevBuf<float> centre_send, shard_values;   // sharded seeding: this rank's picked rows, split draw
    fdb::DevBuf<int> shard_owner;                   // [nb] rank that owns the round's draw
    fdb::DevBuf<uint32_t> picked;                   // [k][nb] picks of seed_run
    fdb::DevBuf<double> pick_bsum;                  // [nb][blocks] block sums of the weights (two-level sampler)
    fdb::DevBuf<unsigned> pick_blast;               // [nb][blocks] last positive weight of every block
    fdb::TcState *tc = nullptr;                     // tensor-core assignment state (tc_assign.cu)
    int last_assign_tc = 0;                         // 1 if the last reassignment ran on the tensor pipe
};

namespace fdb {
int km_sort_members(fdb_km *km, const int *d_active);
int km_update(fdb_km *km, const int *d_active, int loop_mode, float eps, size_t max_rounds);
int km_reassign(fdb_km *km, const int *d_active);
int km_seed_round(fdb_km *km, uint32_t round, int exact, const float *d_centre = nullptr);
int km_seed_pick(fdb_km *km, const float *d_u01, size_t u_stride, size_t u_off, int exact,
                 int absolute = 0);
int km_total_fast(fdb_km *km);
int km_fill_int(fdb_ctx *ctx, int *p, size_t n, int v);
int km_update_partial(fdb_km *km);
int km_update_finish(fdb_km *km);
int km_update_finish_loop(fdb_km *km, const int *d_active, int loop_mode, float eps);
int km_residuals(fdb_vs *vs, const fdb_km *km);
bool tc_eligible(const fdb_km *km);
int tc_reassign(fdb_km *km, const int *d_active);
int tc_last_stats(const fdb_km *km, unsigned out[3]);
void tc_free(fdb_km *km);
}  // namespace fdb
