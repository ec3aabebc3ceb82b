// Multi-GPU build behind the C ABI: the library owns the NCCL communicator and runs
// the collectives itself on the context's stream, between its kernels.
//
//   k-means++ (src/kmeans.rs:142-229), rows sharded: ONE packed all-gather per round.  Every rank draws a row
//   from its own shard (probability ~ its D^2 weight) and publishes [shard total | local pick | that row]; all
//   ranks then choose the shard that owns the round's draw u (probability ~ shard total) and adopt its row.
//   Two-stage sampling -- shard ~ total_r / sum, row ~ w_j / total_r -- is the reference's distribution
//   (WeightedIndex, src/distribution.rs:104-121); the draw inside the shard is a hash of (u, rank), so a rank
//   does not have to wait for the other shards' totals before it picks (the r01 scheme: three dependent
//   collectives per round).  Like the single-GPU parallel sampler the picks are distribution-equal, not equal,
//   to the reference's sequential scan; with the picks injected (fdb_kmeans_seed_chosen) everything is bit-exact.
//
//   Lloyd (src/kmeans.rs:125-137), rows sharded, centroids replicated: per round one all-reduce of
//   [nb*k*m sums || nb*k counts]; gradient and convergence flags are computed redundantly on every rank (the
//   all-reduced buffer is bit-identical everywhere, so every rank takes the same decisions).  The host looks at
//   the flags of round r - LAG while round r is being enqueued: deterministic on every rank, never stalls.
#include "comm.cuh"
#include "kmeans.cuh"

#include <dlfcn.h>

#include <algorithm>
#include <memory>
#include <vector>

namespace fdb {
namespace {

// ---- the few NCCL entry points, resolved at run time ------------------------------------------------
struct NcclId {
    char internal[128];
};
typedef int (*GetUniqueIdFn)(NcclId *);
typedef int (*CommInitRankFn)(void **, int, NcclId, int);
typedef int (*CommDestroyFn)(void *);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*AllGatherFn)(const void *, void *, size_t, int, void *, cudaStream_t);
typedef const char *(*GetErrorStringFn)(int);
typedef int (*AsyncErrorFn)(void *, int *);
constexpr int NCCL_UINT8 = 1, NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2;

struct Nccl {
    void *handle = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    AllReduceFn all_reduce = nullptr;
    AllGatherFn all_gather = nullptr;
    GetErrorStringFn error_string = nullptr;
    AsyncErrorFn async_error = nullptr;
    std::string why;
};

Nccl *nccl() {
    static Nccl n;
    static bool tried = false;
    if (tried) return &n;
    tried = true;
    const char *names[] = {getenv("FDB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    // a copy that is already in the process (torch's) is shared; otherwise the loader's search path
    for (const char *nm : names)
        if (nm && !n.handle) n.handle = dlopen(nm, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    for (const char *nm : names)
        if (nm && !n.handle) n.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (!n.handle) {
        n.why = std::string("libnccl.so.2 not found (set FDB_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
        return &n;
    }
    n.get_unique_id = (GetUniqueIdFn)dlsym(n.handle, "ncclGetUniqueId");
    n.comm_init_rank = (CommInitRankFn)dlsym(n.handle, "ncclCommInitRank");
    n.comm_destroy = (CommDestroyFn)dlsym(n.handle, "ncclCommDestroy");
    n.all_reduce = (AllReduceFn)dlsym(n.handle, "ncclAllReduce");
    n.all_gather = (AllGatherFn)dlsym(n.handle, "ncclAllGather");
    n.error_string = (GetErrorStringFn)dlsym(n.handle, "ncclGetErrorString");
    n.async_error = (AsyncErrorFn)dlsym(n.handle, "ncclCommGetAsyncError");
    if (!n.get_unique_id || !n.comm_init_rank || !n.comm_destroy || !n.all_reduce || !n.all_gather || !n.error_string) {
        n.why = "libnccl lacks an expected symbol";
        n.handle = nullptr;
    }
    return &n;
}

#define FDB_NCCL(expr)                                                                             \
    do {                                                                                           \
        int e__ = (expr);                                                                          \
        if (e__ != 0) {                                                                            \
            fdb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, nccl()->error_string(e__)); \
            return FDB_ERR_NCCL;                                                                   \
        }                                                                                          \
    } while (0)

#define ARG(cond, ...)                   \
    do {                                 \
        if (!(cond)) {                   \
            set_error(__VA_ARGS__);      \
            return FDB_ERR_INVALID_ARGS; \
        }                                \
    } while (0)

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
// the draw a rank uses inside its own shard: a hash of the round's draw and the rank, uniform in [0, 1)
__device__ __forceinline__ float local_draw(float u, int rank) {
    const uint64_t h = mix64(((uint64_t)__float_as_uint(u) << 20) + (uint64_t)(rank + 1) * 0x9E3779B97F4A7C15ULL);
    return (float)(h >> 40) * 5.9604644775390625e-08f;
}

// pack[b] = [total | local pick | picked row] of problem b; layout [nb totals][nb picks][nb][m].  first != 0:
// round 0, the pick is given (ci[b] or NONE) and the "total" says who owns it
__global__ void shard_pack_kernel(const float *x, size_t ldx, size_t col_off, size_t nb, size_t m, const uint32_t *ci,
                                  const float *total, int first, float *pack) {
    const size_t b = blockIdx.x;
    const uint32_t c = ci[b];
    const bool have = first ? c != 0xFFFFFFFFu : total[b] > 0.0f;
    if (threadIdx.x == 0) {
        pack[b] = first ? (have ? 1.0f : 0.0f) : total[b];
        reinterpret_cast<uint32_t *>(pack)[nb + b] = have ? c : 0xFFFFFFFFu;
    }
    float *row = pack + 2 * nb + b * m;
    for (size_t e = threadIdx.x; e < m; e += blockDim.x) row[e] = have ? x[(size_t)c * ldx + col_off + b * m + e] : 0.0f;
}

// after the all-gather: the shard that owns the round's draw (WeightedIndex::sample over the shard totals) gives
// the round's centre on every rank; also stages the next round's local draw
__global__ void shard_adopt_kernel(const float *all, size_t pack_words, int world, int rank, size_t nb, size_t m,
                                   size_t n_global, const float *u01, size_t u_stride, size_t u_off, int have_u,
                                   const float *u_next, size_t un_off, int have_next, uint32_t *ci, float *centre,
                                   uint32_t *picked_global, float *uloc, unsigned *gflag) {
    const size_t b = blockIdx.x;
    __shared__ int own_s;
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int r = 0; r < world; ++r) tsum += (double)all[(size_t)r * pack_words + b];
        const float total = (float)tsum;
        const double sample = have_u ? (double)__fmul_rn(u01[b * u_stride + u_off], total) : 0.0;
        double cum = 0.0;
        int last = -1, own = -1;
        for (int r = 0; r < world; ++r) {
            const double t = (double)all[(size_t)r * pack_words + b];
            if (t > 0.0) {
                last = r;
                if (cum + t > sample) {
                    own = r;
                    break;
                }
                cum += t;
            }
        }
        if (own < 0) own = last;   // rounding pushed the draw past the end: the last shard with weight
        own_s = own;
        if (own < 0) {             // the total weight is zero: WeightedIndex fails in the reference
            atomicOr(gflag, 1u);
            ci[b] = 0xFFFFFFFFu;
            picked_global[b] = 0;
        } else {
            const uint32_t li = reinterpret_cast<const uint32_t *>(all + (size_t)own * pack_words)[nb + b];
            ci[b] = own == rank ? li : 0xFFFFFFFFu;
            picked_global[b] = (uint32_t)(n_global * (size_t)own / (size_t)world) + li;   // the shard's first row + local
        }
        if (have_next) uloc[b] = local_draw(u_next[b * u_stride + un_off], rank);
    }
    __syncthreads();
    const int own = own_s;
    if (own < 0) return;
    const float *row = all + (size_t)own * pack_words + 2 * nb + b * m;
    for (size_t e = threadIdx.x; e < m; e += blockDim.x) centre[b * m + e] = row[e];
}

}  // namespace

int comm_allreduce_sum_f32(fdb_comm *c, float *d_buf, size_t n) {
    if (c->world == 1 || n == 0) return FDB_OK;
    FDB_NCCL(nccl()->all_reduce(d_buf, d_buf, n, NCCL_FLOAT32, NCCL_SUM, c->nccl, c->ctx->stream));
    c->collectives++;
    return FDB_OK;
}

int comm_allgather(fdb_comm *c, const void *d_send, void *d_recv, size_t bytes) {
    if (bytes == 0) return FDB_OK;
    if (c->world == 1) {
        if (d_send != d_recv) FDB_CUDA(cudaMemcpyAsync(d_recv, d_send, bytes, cudaMemcpyDeviceToDevice, c->ctx->stream));
        return FDB_OK;
    }
    FDB_NCCL(nccl()->all_gather(d_send, d_recv, bytes, NCCL_UINT8, c->nccl, c->ctx->stream));
    c->collectives++;
    return FDB_OK;
}

int comm_check(fdb_comm *c) {
    if (c->world == 1 || !nccl()->async_error) return FDB_OK;
    int err = 0;
    FDB_NCCL(nccl()->async_error(c->nccl, &err));
    FDB_NCCL(err);
    return FDB_OK;
}

}  // namespace fdb

using namespace fdb;

extern "C" {

int fdb_comm_unique_id(uint8_t *id) {
    ARG(id, "id is null");
    Nccl *n = nccl();
    if (!n->handle) {
        set_error("%s", n->why.c_str());
        return FDB_ERR_NCCL;
    }
    NcclId u;
    memset(&u, 0, sizeof(u));
    FDB_NCCL(n->get_unique_id(&u));
    memcpy(id, u.internal, FDB_COMM_ID_BYTES);
    return FDB_OK;
}

int fdb_comm_create(fdb_ctx *ctx, int world, int rank, const uint8_t *id, fdb_comm **out) {
    ARG(ctx && out, "null argument");
    ARG(world >= 1 && rank >= 0 && rank < world, "rank %d of %d", rank, world);
    *out = nullptr;
    FDB_TRY(ctx->use());
    std::unique_ptr<fdb_comm> c(new fdb_comm);
    c->ctx = ctx;
    c->world = world;
    c->rank = rank;
    if (world > 1) {
        ARG(id, "id is null");
        Nccl *n = nccl();
        if (!n->handle) {
            set_error("%s", n->why.c_str());
            return FDB_ERR_NCCL;
        }
        NcclId u;
        memcpy(u.internal, id, FDB_COMM_ID_BYTES);
        FDB_NCCL(n->comm_init_rank(&c->nccl, world, u, rank));
    }
    *out = c.release();
    return FDB_OK;
}

void fdb_comm_destroy(fdb_comm *c) {
    if (!c) return;
    if (c->nccl) {
        cudaSetDevice(c->ctx->device);
        cudaStreamSynchronize(c->ctx->stream);
        nccl()->comm_destroy(c->nccl);
    }
    delete c;
}

int fdb_comm_world(const fdb_comm *c) { return c ? c->world : 0; }
int fdb_comm_rank(const fdb_comm *c) { return c ? c->rank : -1; }
uint64_t fdb_comm_collective_count(const fdb_comm *c) { return c ? c->collectives : 0; }

int fdb_comm_allreduce_device(fdb_comm *c, float *d_buf, size_t n) {
    ARG(c && (d_buf || n == 0), "null argument");
    FDB_TRY(c->ctx->use());
    return comm_allreduce_sum_f32(c, d_buf, n);
}

int fdb_comm_allgather_device(fdb_comm *c, const void *d_send, void *d_recv, size_t bytes_per_rank) {
    ARG(c && ((d_send && d_recv) || bytes_per_rank == 0), "null argument");
    FDB_TRY(c->ctx->use());
    return comm_allgather(c, d_send, d_recv, bytes_per_rank);
}

/* max over the ranks of n host doubles (timings): one small all-reduce + synchronisation; doubles as a barrier */
int fdb_comm_max_f64(fdb_comm *c, double *values, size_t n) {
    ARG(c && values && n > 0 && n <= 64, "invalid argument");
    fdb_ctx *ctx = c->ctx;
    FDB_TRY(ctx->use());
    if (c->world > 1) {
        FDB_TRY(c->scratch.ensure(64 * sizeof(double)));
        FDB_CUDA(cudaMemcpyAsync(c->scratch.p, values, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        FDB_NCCL(nccl()->all_reduce(c->scratch.p, c->scratch.p, n, NCCL_FLOAT64, NCCL_MAX, c->nccl, ctx->stream));
        c->collectives++;
        FDB_CUDA(cudaMemcpyAsync(values, c->scratch.p, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    }
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return comm_check(c);
}

/* ---- k-means++ over row shards: one packed all-gather per round ------------------------------------ */
int fdb_kmeans_seed_run_sharded(fdb_km *km, fdb_comm *c, size_t n_global, const uint32_t *first_global, const float *u01,
                                uint32_t *picked_global) {
    ARG(km && c && first_global && (u01 || km->k == 1), "null argument");
    ARG(km->ctx == c->ctx, "the communicator belongs to another context");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb, k = km->k, m = km->m;
    const int world = c->world, rank = c->rank;
    const size_t lo = n_global * (size_t)rank / (size_t)world, hi = n_global * (size_t)(rank + 1) / (size_t)world;
    ARG(hi - lo == km->n, "this rank holds %zu rows, its shard of %zu rows over %d ranks is %zu", km->n, n_global, world,
        hi - lo);
    ARG(k <= n_global && n_global < (1ull << 32), "k = %zu, %zu rows", k, n_global);
    cudaStream_t st = ctx->stream;
    // the seeding state (weights, chosen marks, ci, total, centre, picks): what the host-driven sharded calls use
    float *d_tot = nullptr, *d_send = nullptr, *d_u = nullptr;
    uint32_t *d_pick = nullptr, *d_picked = nullptr;
    FDB_TRY(fdb_kmeans_seed_sharded_begin(km, &d_tot, &d_pick, &d_send, &d_u, &d_picked));
    (void)d_send;
    const size_t pack_words = nb * (2 + m);
    FDB_TRY(c->send.ensure(pack_words * 4));
    FDB_TRY(c->recv.ensure(pack_words * 4 * (size_t)world));
    FDB_TRY(c->scratch.ensure(std::max<size_t>(64 * sizeof(double), sizeof(unsigned))));
    unsigned *gflag = reinterpret_cast<unsigned *>(c->scratch.p);
    FDB_CUDA(cudaMemsetAsync(gflag, 0, sizeof(unsigned), st));
    float *pack = reinterpret_cast<float *>(c->send.p), *all = reinterpret_cast<float *>(c->recv.p);
    // round 0: the first centre is given (gen_range(0..n), src/kmeans.rs:172); its owner publishes it
    std::vector<uint32_t> local_first(nb);
    for (size_t b = 0; b < nb; ++b) {
        ARG(first_global[b] < n_global, "first index out of range");
        local_first[b] = (first_global[b] >= lo && first_global[b] < hi) ? (uint32_t)(first_global[b] - lo) : 0xFFFFFFFFu;
    }
    FDB_CUDA(cudaMemcpyAsync(d_pick, local_first.data(), nb * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    FDB_CUDA(cudaStreamSynchronize(st));   // local_first is a stack-lifetime staging buffer
    if (k > 1) {
        FDB_TRY(km->u01.ensure(nb * (k - 1)));   // (seed_sharded_begin sized it for one round)
        FDB_CUDA(cudaMemcpyAsync(km->u01.p, u01, nb * (k - 1) * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    float *uloc = km->shard_values.p;
    for (size_t i = 0; i < k; ++i) {
        if (i > 0) FDB_TRY(km_seed_pick(km, uloc, 1, 0, 0, 0));   // local pick ~ local weights; also leaves the shard total
        shard_pack_kernel<<<(unsigned)nb, 128, 0, st>>>(km->vs->d, km->vs->dim, km->col_off, nb, m, d_pick, d_tot, i == 0, pack);
        FDB_TRY(comm_allgather(c, pack, all, pack_words * 4));
        shard_adopt_kernel<<<(unsigned)nb, 128, 0, st>>>(all, pack_words, world, rank, nb, m, n_global, km->u01.p, k - 1,
                                                         i ? i - 1 : 0, i > 0, km->u01.p, i, i + 1 < k, d_pick,
                                                         km->centre.p, d_picked + i * nb, uloc, gflag);
        ctx->launches += 2;
        FDB_CHECK_LAUNCH();
        FDB_TRY(km_seed_round(km, (uint32_t)i, 0, km->centre.p));
    }
    std::vector<uint32_t> stage(nb * k);
    unsigned hflag = 0;
    FDB_CUDA(cudaMemcpyAsync(stage.data(), d_picked, stage.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaMemcpyAsync(&hflag, gflag, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
    FDB_CUDA(cudaStreamSynchronize(st));
    if (picked_global)
        for (size_t b = 0; b < nb; ++b)
            for (size_t i = 0; i < k; ++i) picked_global[b * k + i] = stage[i * nb + b];
    unsigned f = 0;
    FDB_TRY(ctx->check_flags(&f));
    FDB_TRY(comm_check(c));
    f &= ~FLAG_WEIGHTS;                 // a shard's own total may legitimately be zero ...
    if (hflag) f |= FLAG_WEIGHTS;       // ... the total over all shards may not (WeightedIndex fails in the reference)
    return map_flags(f);
}

/* ---- the Lloyd loop over row shards: one all-reduce per round --------------------------------------- */
int fdb_kmeans_run_sharded(fdb_km *km, fdb_comm *c, size_t max_rounds, float epsilon, float *gradients, uint32_t *rounds,
                           uint32_t *reassigns) {
    ARG(km && c, "null argument");
    ARG(km->ctx == c->ctx, "the communicator belongs to another context");
    ARG(max_rounds <= km->max_rounds, "max_rounds exceeds %zu", km->max_rounds);
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb;
    cudaStream_t st = ctx->stream;
    FDB_TRY(fdb_kmeans_sharded_loop_begin(km));
    // the flags of round r are copied to pinned memory behind an event; the host reads the copy of round
    // r - LAG before it enqueues round r: the same decision on every rank, and the stream never drains
    constexpr size_t LAG = 3;
    const size_t slots = LAG + 1;
    ARG(slots * nb * sizeof(int) <= ctx->h_pinned_bytes / 2, "too many problems for the staging buffer");
    int *h_active = static_cast<int *>(ctx->h_pinned);
    std::vector<cudaEvent_t> ev(slots, nullptr);
    for (auto &e : ev) FDB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    int rc = FDB_OK;
    auto body = [&]() -> int {
        for (size_t r = 0; r < max_rounds; ++r) {
            if (r >= LAG) {
                const size_t s = (r - LAG) % slots;
                FDB_CUDA(cudaEventSynchronize(ev[s]));
                bool any = false;
                for (size_t b = 0; b < nb; ++b) any |= h_active[s * nb + b] != 0;
                if (!any) break;   // every problem had converged LAG rounds ago: the rounds since changed nothing
            }
            float *buf = nullptr;
            size_t nfl = 0;
            FDB_TRY(fdb_kmeans_sharded_partial_async(km, &buf, &nfl));
            FDB_TRY(comm_allreduce_sum_f32(c, buf, nfl));
            FDB_TRY(fdb_kmeans_sharded_finish_async(km, epsilon));
            const size_t s = r % slots;
            FDB_CUDA(cudaMemcpyAsync(h_active + s * nb, km->active.p, nb * sizeof(int), cudaMemcpyDeviceToHost, st));
            FDB_CUDA(cudaEventRecord(ev[s], st));
        }
        return FDB_OK;
    };
    rc = body();
    for (auto &e : ev) cudaEventDestroy(e);
    FDB_TRY(rc);
    std::vector<float> gh(nb * km->max_rounds);
    std::vector<uint32_t> rr(nb), ra(nb);
    FDB_TRY(fdb_kmeans_sharded_loop_end(km, gh.data(), rr.data(), ra.data()));
    FDB_TRY(comm_check(c));
    for (size_t b = 0; b < nb; ++b) {
        if (gradients) memcpy(gradients + b * max_rounds, gh.data() + b * km->max_rounds, max_rounds * sizeof(float));
        if (rounds) rounds[b] = rr[b];
        if (reassigns) reassigns[b] = ra[b];
    }
    return FDB_OK;
}

}  // extern "C"
