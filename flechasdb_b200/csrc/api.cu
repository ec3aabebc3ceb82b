// The C ABI of include/flechasdb_b200.h: context, vector sets, k-means entry points.
// (index / query entry points live in query.cu)
#include "kmeans.cuh"

#include <memory>

namespace fdb {

static thread_local std::string g_last_error;

void set_error(const char *fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_last_error = buf;
}

int map_flags(unsigned f) {
    if (f & FLAG_EMPTY_CLUSTER) {
        set_error("empty cluster: assertion `count != 0` failed (src/kmeans.rs:259)");
        return FDB_ERR_EMPTY_CLUSTER;
    }
    if (f & FLAG_WEIGHTS) {
        set_error("k-means++ weights exhausted: WeightedIndex update/sample failed "
                  "(src/kmeans.rs:199,207,216)");
        return FDB_ERR_WEIGHTS;
    }
    if (f & (FLAG_NO_ARGMIN | FLAG_NAN)) {
        set_error("non-finite distances: Option::unwrap on None / partial_cmp "
                  "(src/kmeans.rs:304, src/db/stored.rs:385,426)");
        return FDB_ERR_NAN;
    }
    return FDB_OK;
}

namespace {

__device__ __forceinline__ uint64_t splitmix64(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

__global__ void fill_uniform_kernel(float *out, size_t count, uint64_t seed, uint64_t start) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = (float)(splitmix64(seed, start + i) >> 40) * 5.9604644775390625e-08f;
}

__global__ void flush_kernel(char *p, size_t bytes) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * 16;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 16; i + 16 <= bytes; i += stride)
        *reinterpret_cast<uint4 *>(p + i) = make_uint4(1, 2, 3, 4);
}

}  // namespace
}  // namespace fdb

using namespace fdb;

int fdb_ctx::use() const {
    FDB_CUDA(cudaSetDevice(device));
    return FDB_OK;
}

int fdb_ctx::check_flags(unsigned *out) {
    FDB_CUDA(cudaMemcpyAsync(h_flags, d_flags, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    FDB_CUDA(cudaStreamSynchronize(stream));
    *out = *h_flags;
    if (*h_flags) FDB_CUDA(cudaMemsetAsync(d_flags, 0, sizeof(unsigned), stream));
    return FDB_OK;
}

#define ARG(cond, ...)                   \
    do {                                 \
        if (!(cond)) {                   \
            set_error(__VA_ARGS__);      \
            return FDB_ERR_INVALID_ARGS; \
        }                                \
    } while (0)

extern "C" {

const char *fdb_last_error(void) { return g_last_error.c_str(); }
int fdb_version(void) { return 100; }

int fdb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int fdb_ctx_create(int device, fdb_ctx **out) {
    ARG(out, "out is null");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error("no CUDA device: %s (this library has no CPU fallback)",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return FDB_ERR_CUDA;
    }
    ARG(device >= 0 && device < n, "device %d out of range (%d devices)", device, n);
    std::unique_ptr<fdb_ctx> c(new fdb_ctx);
    c->device = device;
    FDB_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    FDB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                  prop.minor);
        return FDB_ERR_CUDA;
    }
    c->sm_count = prop.multiProcessorCount;
    FDB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    FDB_CUDA(cudaEventCreate(&c->ev0));
    FDB_CUDA(cudaEventCreate(&c->ev1));
    FDB_CUDA(cudaMalloc((void **)&c->d_flags, sizeof(unsigned)));
    FDB_CUDA(cudaMemset(c->d_flags, 0, sizeof(unsigned)));
    FDB_CUDA(cudaMallocHost((void **)&c->h_flags, sizeof(unsigned)));
    c->h_pinned_bytes = 1 << 20;
    FDB_CUDA(cudaMallocHost(&c->h_pinned, c->h_pinned_bytes));
    *out = c.release();
    return FDB_OK;
}

void fdb_ctx_destroy(fdb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    ctx->flush_buf.release();
    fdb::dev_trim();   // the cache of freed blocks goes back to the driver with the context
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int fdb_ctx_sync(fdb_ctx *ctx) {
    ARG(ctx, "ctx is null");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}

int fdb_ctx_timer_start(fdb_ctx *ctx) {
    ARG(ctx, "ctx is null");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    return FDB_OK;
}

int fdb_ctx_timer_stop(fdb_ctx *ctx, float *ms) {
    ARG(ctx && ms, "null argument");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    FDB_CUDA(cudaEventSynchronize(ctx->ev1));
    FDB_CUDA(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return FDB_OK;
}

uint64_t fdb_ctx_launch_count(const fdb_ctx *ctx) { return ctx ? ctx->launches : 0; }

/* ---- raw device buffers ----------------------------------------------------------- */
int fdb_device_alloc(fdb_ctx *ctx, size_t bytes, void **out) {
    ARG(ctx && out, "null argument");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaMalloc(out, bytes ? bytes : 16));
    return FDB_OK;
}

int fdb_device_free(fdb_ctx *ctx, void *p) {
    ARG(ctx, "ctx is null");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (p) FDB_CUDA(cudaFree(p));
    return FDB_OK;
}

/* page-locks caller memory (a Rust Vec<f32>, a numpy array) so that host <-> device copies of it are asynchronous
 * DMA transfers: fdb_index_query then overlaps the copy of a batch with answering it */
int fdb_host_register(fdb_ctx *ctx, void *p, size_t bytes) {
    ARG(ctx && p && bytes, "null argument");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return FDB_OK;
}

int fdb_host_unregister(fdb_ctx *ctx, void *p) {
    ARG(ctx && p, "null argument");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaHostUnregister(p));
    return FDB_OK;
}

int fdb_device_upload(fdb_ctx *ctx, void *d_dst, const void *src, size_t bytes) {
    ARG(ctx && (bytes == 0 || (d_dst && src)), "null argument");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaMemcpyAsync(d_dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}

int fdb_device_download(fdb_ctx *ctx, void *dst, const void *d_src, size_t bytes) {
    ARG(ctx && (bytes == 0 || (dst && d_src)), "null argument");
    FDB_TRY(ctx->use());
    FDB_CUDA(cudaMemcpyAsync(dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    return FDB_OK;
}

int fdb_device_fill_uniform(fdb_ctx *ctx, float *d, size_t count, uint64_t seed, uint64_t start) {
    ARG(ctx && (d || !count), "null argument");
    FDB_TRY(ctx->use());
    if (!count) return FDB_OK;
    fill_uniform_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(d, count, seed, start);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int fdb_device_flush_l2(fdb_ctx *ctx) {
    ARG(ctx, "ctx is null");
    FDB_TRY(ctx->use());
    const size_t bytes = (size_t)256 << 20;  // 2x the 126 MB L2
    FDB_TRY(ctx->flush_buf.ensure(bytes));
    flush_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->flush_buf.p, bytes);
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

/* ---- vector sets ------------------------------------------------------------------- */
static int vs_new(fdb_ctx *ctx, size_t n, size_t dim, fdb_vs **out) {
    ARG(ctx && out, "null argument");
    ARG(dim > 0, "vector size must be non-zero (NonZeroUsize, src/vector.rs:42)");
    FDB_TRY(ctx->use());
    std::unique_ptr<fdb_vs> v(new fdb_vs);
    v->ctx = ctx;
    v->n = n;
    v->dim = dim;
    v->owned = true;
    FDB_TRY(fdb::dev_alloc((void **)&v->d, std::max<size_t>(n * dim, 4) * sizeof(float)));
    *out = v.release();
    return FDB_OK;
}

int fdb_vs_upload(fdb_ctx *ctx, const float *rows, size_t n, size_t dim, fdb_vs **out) {
    ARG(rows || n == 0, "rows is null");
    FDB_TRY(vs_new(ctx, n, dim, out));
    if (n) {
        cudaError_t e = cudaMemcpyAsync((*out)->d, rows, n * dim * sizeof(float),
                                        cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) {
            set_error("upload failed: %s", cudaGetErrorString(e));
            fdb_vs_destroy(*out);
            *out = nullptr;
            return FDB_ERR_CUDA;
        }
    }
    return FDB_OK;
}

int fdb_vs_from_device(fdb_ctx *ctx, float *device_rows, size_t n, size_t dim, fdb_vs **out) {
    ARG(ctx && out && device_rows, "null argument");
    ARG(dim > 0, "vector size must be non-zero");
    std::unique_ptr<fdb_vs> v(new fdb_vs);
    v->ctx = ctx;
    v->n = n;
    v->dim = dim;
    v->owned = false;
    v->d = device_rows;
    *out = v.release();
    return FDB_OK;
}

int fdb_vs_generate(fdb_ctx *ctx, size_t n, size_t dim, uint64_t seed, uint64_t start, fdb_vs **out) {
    FDB_TRY(vs_new(ctx, n, dim, out));
    return fdb_device_fill_uniform(ctx, (*out)->d, n * dim, seed, start);
}

int fdb_vs_download_rows(fdb_vs *vs, size_t first_row, size_t nrows, float *rows) {
    ARG(vs && (rows || !nrows), "null argument");
    ARG(first_row + nrows <= vs->n, "row range out of bounds");
    FDB_TRY(vs->ctx->use());
    if (!nrows) return FDB_OK;
    FDB_CUDA(cudaMemcpyAsync(rows, vs->d + first_row * vs->dim, nrows * vs->dim * sizeof(float),
                             cudaMemcpyDeviceToHost, vs->ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(vs->ctx->stream));
    return FDB_OK;
}

int fdb_vs_download(fdb_vs *vs, float *rows) {
    ARG(vs, "vs is null");
    return fdb_vs_download_rows(vs, 0, vs->n, rows);
}

size_t fdb_vs_len(const fdb_vs *vs) { return vs ? vs->n : 0; }
size_t fdb_vs_vector_size(const fdb_vs *vs) { return vs ? vs->dim : 0; }
float *fdb_vs_device_ptr(fdb_vs *vs) { return vs ? vs->d : nullptr; }

void fdb_vs_destroy(fdb_vs *vs) {
    if (!vs) return;
    cudaSetDevice(vs->ctx->device);
    cudaStreamSynchronize(vs->ctx->stream);
    if (vs->owned && vs->d) fdb::dev_free(vs->d);
    delete vs;
}

int fdb_vs_subtract_assigned(fdb_vs *vs, const fdb_km *km) {
    ARG(vs && km, "null argument");
    ARG(km->vs == vs && km->nb == 1 && km->col_off == 0 && km->m == vs->dim,
        "km must be a single problem over the full vectors of vs");
    FDB_TRY(vs->ctx->use());
    return km_residuals(vs, km);
}

/* ---- k-means ------------------------------------------------------------------------- */
int fdb_kmeans_begin(fdb_vs *vs, size_t col_off, size_t dim, size_t nb, size_t k, fdb_km **out) {
    ARG(vs && out, "null argument");
    *out = nullptr;
    ARG(k > 0 && dim > 0 && nb > 0, "k, dim and nb must be non-zero");
    ARG(col_off + nb * dim <= vs->dim, "sub-vector views exceed the vector size: %zu + %zu*%zu > %zu",
        col_off, nb, dim, vs->dim);
    /* cluster_with_events, src/kmeans.rs:116-120 */
    ARG(vs->n >= k, "vs has fewer vectors than k: %zu < %zu", vs->n, k);
    if (vs->n >= (1ull << 31) || k > 65535) {
        set_error("unsupported size: n=%zu k=%zu", vs->n, k);
        return FDB_ERR_UNSUPPORTED;
    }
    FDB_TRY(vs->ctx->use());
    std::unique_ptr<fdb_km> km(new fdb_km);
    km->ctx = vs->ctx;
    km->vs = vs;
    km->col_off = col_off;
    km->m = dim;
    km->nb = nb;
    km->k = k;
    km->n = vs->n;
    const size_t n = vs->n;
    FDB_TRY(km->centroids.alloc(nb * k * dim));
    FDB_TRY(km->old_centroids.alloc(nb * k * dim));
    FDB_TRY(km->indices.alloc(nb * n));
    FDB_TRY(km->cnorm.alloc(nb * k));
    FDB_TRY(km->cdist.alloc(nb * k));
    FDB_TRY(km->grad.alloc(nb));
    FDB_TRY(km->grad_hist.alloc(nb * km->max_rounds));
    FDB_TRY(km->total.alloc(nb));
    FDB_TRY(km->rounds.alloc(nb));
    FDB_TRY(km->reassigns.alloc(nb));
    FDB_TRY(km->ci.alloc(nb));
    FDB_TRY(km->active.alloc(nb));
    FDB_TRY(km->step_active.alloc(nb));
    FDB_CUDA(cudaMemsetAsync(km->indices.p, 0, nb * n * sizeof(uint32_t), km->ctx->stream));
    FDB_CUDA(cudaMemsetAsync(km->centroids.p, 0, nb * k * dim * sizeof(float), km->ctx->stream));
    *out = km.release();
    return FDB_OK;
}

void fdb_kmeans_destroy(fdb_km *km) {
    if (!km) return;
    cudaSetDevice(km->ctx->device);
    cudaStreamSynchronize(km->ctx->stream);
    tc_free(km);
    delete km;
}

static int km_finish_call(fdb_km *km) {
    unsigned f = 0;
    FDB_TRY(km->ctx->check_flags(&f));
    return map_flags(f);
}

static int km_seed_alloc(fdb_km *km) {
    const size_t cnt = km->nb * km->n;
    FDB_TRY(km->weights.ensure(cnt));
    FDB_TRY(km->weights_new.ensure(cnt));
    FDB_TRY(km->chosen.ensure(cnt));
    return FDB_OK;
}

static int upload_small(fdb_ctx *ctx, void *dst, const void *src, size_t bytes) {
    FDB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));  // src may be pageable / go out of scope
    return FDB_OK;
}

// k == n: every vector is its own centre (src/kmeans.rs:158-170)
static int km_seed_identity(fdb_km *km) {
    std::vector<uint32_t> idx(km->nb * km->n);
    for (size_t b = 0; b < km->nb; ++b)
        for (size_t i = 0; i < km->n; ++i) idx[b * km->n + i] = (uint32_t)i;
    FDB_TRY(upload_small(km->ctx, km->indices.p, idx.data(), idx.size() * sizeof(uint32_t)));
    for (size_t b = 0; b < km->nb; ++b)
        FDB_CUDA(cudaMemcpy2DAsync(km->centroids.p + b * km->k * km->m, km->m * sizeof(float),
                                   km->vs->d + km->col_off + b * km->m, km->vs->dim * sizeof(float),
                                   km->m * sizeof(float), km->n, cudaMemcpyDeviceToDevice,
                                   km->ctx->stream));
    return FDB_OK;
}

int fdb_kmeans_seed_first(fdb_km *km, const uint32_t *ci) {
    ARG(km && ci, "null argument");
    FDB_TRY(km->ctx->use());
    for (size_t b = 0; b < km->nb; ++b) ARG(ci[b] < km->n, "first index out of range");
    if (km->k == km->n) return km_seed_identity(km);
    FDB_TRY(km_seed_alloc(km));
    FDB_CUDA(cudaMemsetAsync(km->chosen.p, 0, km->nb * km->n, km->ctx->stream));
    FDB_TRY(upload_small(km->ctx, km->ci.p, ci, km->nb * sizeof(uint32_t)));
    // exact total is cheap for round 0 (16-lane sum); keep it always
    FDB_TRY(km_seed_round(km, 0, 1));
    if (km->k == 1) {  // :176-184 returns before WeightedIndex::new
        unsigned f = 0;
        FDB_TRY(km->ctx->check_flags(&f));
        return map_flags(f & ~FLAG_WEIGHTS);
    }
    return km_finish_call(km);
}

int fdb_kmeans_seed_total(fdb_km *km, float *totals) {
    ARG(km && totals, "null argument");
    FDB_TRY(km->ctx->use());
    FDB_CUDA(cudaMemcpyAsync(totals, km->total.p, km->nb * sizeof(float), cudaMemcpyDeviceToHost,
                             km->ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(km->ctx->stream));
    return FDB_OK;
}

int fdb_kmeans_seed_pick(fdb_km *km, const float *u01, int exact, uint32_t *ci_out) {
    ARG(km && u01 && ci_out, "null argument");
    ARG(km->weights.p, "seed_first has not been called");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km->u01.ensure(km->nb));
    FDB_TRY(upload_small(km->ctx, km->u01.p, u01, km->nb * sizeof(float)));
    FDB_TRY(km_seed_pick(km, km->u01.p, 1, 0, exact));
    FDB_CUDA(cudaMemcpyAsync(ci_out, km->ci.p, km->nb * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             km->ctx->stream));
    return km_finish_call(km);
}

int fdb_kmeans_seed_add(fdb_km *km, size_t i, const uint32_t *ci, int exact) {
    ARG(km && ci, "null argument");
    ARG(km->weights.p, "seed_first has not been called");
    ARG(i >= 1 && i < km->k, "round %zu out of range 1..%zu", i, km->k);
    for (size_t b = 0; b < km->nb; ++b) ARG(ci[b] < km->n, "chosen index out of range");
    FDB_TRY(km->ctx->use());
    FDB_TRY(upload_small(km->ctx, km->ci.p, ci, km->nb * sizeof(uint32_t)));
    FDB_TRY(km_seed_round(km, (uint32_t)i, exact));
    if (!exact) FDB_TRY(km_total_fast(km));
    return km_finish_call(km);
}

/* ---- sharded rows (multi-GPU build): the chosen vector may live on another rank ------------ */
int fdb_kmeans_seed_pick_value(fdb_km *km, const float *sample_values, uint32_t *ci_out) {
    ARG(km && sample_values && ci_out, "null argument");
    ARG(km->weights.p, "seeding has not started");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km->u01.ensure(km->nb));
    FDB_TRY(upload_small(km->ctx, km->u01.p, sample_values, km->nb * sizeof(float)));
    FDB_TRY(km_seed_pick(km, km->u01.p, 1, 0, 0, 1));
    FDB_CUDA(cudaMemcpyAsync(ci_out, km->ci.p, km->nb * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             km->ctx->stream));
    unsigned f = 0;
    FDB_TRY(km->ctx->check_flags(&f));
    return map_flags(f & ~FLAG_WEIGHTS);  // an empty shard is not an error; the caller sums the totals
}

int fdb_kmeans_seed_round_ext(fdb_km *km, size_t i, const float *centres, const uint32_t *local_ci) {
    ARG(km && centres && local_ci, "null argument");
    ARG(i < km->k, "round %zu out of range", i);
    for (size_t b = 0; b < km->nb; ++b)
        ARG(local_ci[b] == 0xFFFFFFFFu || local_ci[b] < km->n, "local index out of range");
    FDB_TRY(km->ctx->use());
    if (i == 0) {
        FDB_TRY(km_seed_alloc(km));
        FDB_CUDA(cudaMemsetAsync(km->chosen.p, 0, km->nb * km->n, km->ctx->stream));
    } else {
        ARG(km->weights.p, "round 0 has not run");
    }
    FDB_TRY(km->centre.ensure(km->nb * km->m));
    FDB_TRY(upload_small(km->ctx, km->centre.p, centres, km->nb * km->m * sizeof(float)));
    FDB_TRY(upload_small(km->ctx, km->ci.p, local_ci, km->nb * sizeof(uint32_t)));
    FDB_TRY(km_seed_round(km, (uint32_t)i, 0, km->centre.p));
    FDB_TRY(km_total_fast(km));
    unsigned f = 0;
    FDB_TRY(km->ctx->check_flags(&f));
    return map_flags(f & ~FLAG_WEIGHTS);  // a shard's own total may legitimately be zero
}

// ---- sharded k-means++ without host round trips (rows sharded over `world` ranks) -------------
// One round = local totals -> all-gather -> split of the global draw + local pick -> all-gather of
// the picks and of the picked rows -> D^2 pass.  The three stage functions below only enqueue work
// on the context's stream; the caller runs the all-gathers (NCCL through torch.distributed) on the
// same stream in between.
namespace fdb_shard {
using namespace fdb;

// WeightedIndex::sample over the concatenated shards (src/distribution.rs:104-121): the draw lands in
// the shard whose cumulative range of totals contains it.  values[b] = the part of the draw inside
// this shard (absolute sample for the local pick), or -1 when another shard owns it.
__global__ void shard_split_kernel(const float *all_totals, int world, int rank, size_t nb, const float *u01,
                                   float *values, int *owner) {
    const size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    double tsum = 0.0;
    for (int r = 0; r < world; ++r) tsum += (double)all_totals[(size_t)r * nb + b];
    const float total = (float)tsum;
    const double sample = (double)__fmul_rn(u01[b], total);
    double cum = 0.0;
    int last = -1, own = -1;
    float val = 0.0f;
    for (int r = 0; r < world; ++r) {
        const double t = (double)all_totals[(size_t)r * nb + b];
        if (t > 0.0) {
            last = r;
            if (cum + t > sample) {
                own = r;
                val = (float)(sample - cum);
                break;
            }
            cum += t;
        }
    }
    if (own < 0 && last >= 0) {   // rounding pushed the draw past the end: the last shard with weight
        own = last;
        val = all_totals[(size_t)last * nb + b];
    }
    owner[b] = own;               // -1: the total weight is zero (WeightedIndex fails in the reference)
    values[b] = own == rank ? val : -1.0f;
}

// after the local pick: ranks that do not own the draw publish NONE, the owner its row
__global__ void shard_publish_kernel(const int *owner, int rank, size_t nb, size_t m, const float *x, size_t ldx,
                                     size_t col_off, uint32_t *ci, float *centre_send) {
    const size_t b = blockIdx.x;
    const bool mine = owner[b] == rank;
    if (!mine && threadIdx.x == 0) ci[b] = 0xFFFFFFFFu;
    __syncthreads();
    const uint32_t c = mine ? ci[b] : 0u;
    for (size_t e = threadIdx.x; e < m; e += blockDim.x)
        centre_send[b * m + e] = mine ? x[(size_t)c * ldx + col_off + b * m + e] : 0.0f;
}

// after the all-gathers: the owner's row becomes the round's centre on every rank
__global__ void shard_adopt_kernel(const uint32_t *all_picks, const float *all_centres, int world, int rank,
                                   size_t nb, size_t m, size_t n_global, uint32_t *ci, float *centre,
                                   uint32_t *picked_global, unsigned *flags) {
    const size_t b = blockIdx.x;
    int own = -1;
    for (int r = 0; r < world; ++r)
        if (all_picks[(size_t)r * nb + b] != 0xFFFFFFFFu) {
            own = r;
            break;
        }
    if (own < 0) {
        if (threadIdx.x == 0) {
            atomicOr(flags, FLAG_WEIGHTS);
            ci[b] = 0xFFFFFFFFu;
            picked_global[b] = 0;
        }
        return;
    }
    const uint32_t li = all_picks[(size_t)own * nb + b];
    if (threadIdx.x == 0) {
        ci[b] = own == rank ? li : 0xFFFFFFFFu;
        picked_global[b] = (uint32_t)(n_global * (size_t)own / (size_t)world) + li;   // shard_rows().lo + local
    }
    for (size_t e = threadIdx.x; e < m; e += blockDim.x)
        centre[b * m + e] = all_centres[((size_t)own * nb + b) * m + e];
}

}  // namespace fdb_shard

static int seed_loop(fdb_km *km, const uint32_t *first, const float *u01, const uint32_t *chosen,
                     int exact, uint32_t *picked) {
    const size_t nb = km->nb, k = km->k;
    FDB_TRY(km->ctx->use());
    if (k == km->n) {
        FDB_TRY(km_seed_identity(km));
        if (picked)
            for (size_t b = 0; b < nb; ++b)
                for (size_t i = 0; i < k; ++i) picked[b * k + i] = (uint32_t)i;
        return fdb_ctx_sync(km->ctx);
    }
    FDB_TRY(km_seed_alloc(km));
    FDB_CUDA(cudaMemsetAsync(km->chosen.p, 0, nb * km->n, km->ctx->stream));
    FDB_TRY(km->picked.ensure(nb * k));
    // picked[b][i] laid out [i][b] on the device so that round i reads a contiguous [nb] slice
    std::vector<uint32_t> stage(nb * k, 0);
    for (size_t b = 0; b < nb; ++b) {
        const uint32_t f = chosen ? chosen[b * k] : first[b];
        ARG(f < km->n, "first index out of range");
        stage[b] = f;
        if (chosen)
            for (size_t i = 1; i < k; ++i) {
                ARG(chosen[b * k + i] < km->n, "chosen index out of range");
                stage[i * nb + b] = chosen[b * k + i];
            }
    }
    FDB_TRY(upload_small(km->ctx, km->picked.p, stage.data(), stage.size() * sizeof(uint32_t)));
    if (u01 && k > 1) {
        FDB_TRY(km->u01.ensure(nb * (k - 1)));
        FDB_TRY(upload_small(km->ctx, km->u01.p, u01, nb * (k - 1) * sizeof(float)));
    }
    uint32_t *saved_ci = km->ci.p;
    int rc = FDB_OK;
    for (size_t i = 0; i < k && rc == FDB_OK; ++i) {
        km->ci.p = km->picked.p + i * nb;  // round i reads (and a pick writes) its own slice
        if (i > 0 && !chosen) rc = km_seed_pick(km, km->u01.p, k - 1, i - 1, exact);
        if (rc == FDB_OK) rc = km_seed_round(km, (uint32_t)i, i == 0 ? 1 : exact);
        if (rc == FDB_OK && i > 0 && !exact && !chosen && i + 1 < k) {
            /* the next pick recomputes the total itself */
        }
    }
    km->ci.p = saved_ci;
    FDB_TRY(rc);
    if (!exact && k > 1) FDB_TRY(km_total_fast(km));  // detects "total weight becomes zero"
    if (picked) {
        FDB_CUDA(cudaMemcpyAsync(stage.data(), km->picked.p, stage.size() * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, km->ctx->stream));
        FDB_CUDA(cudaStreamSynchronize(km->ctx->stream));
        for (size_t b = 0; b < nb; ++b)
            for (size_t i = 0; i < k; ++i) picked[b * k + i] = stage[i * nb + b];
    }
    unsigned f = 0;
    FDB_TRY(km->ctx->check_flags(&f));
    if (k == 1) f &= ~FLAG_WEIGHTS;
    return map_flags(f);
}

void *fdb_ctx_stream(fdb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int fdb_kmeans_seed_sharded_begin(fdb_km *km, float **d_totals, uint32_t **d_pick, float **d_centre_send,
                                  float **d_u01, uint32_t **d_picked_global) {
    ARG(km && d_totals && d_pick && d_centre_send && d_u01 && d_picked_global, "null argument");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km_seed_alloc(km));
    FDB_CUDA(cudaMemsetAsync(km->chosen.p, 0, km->nb * km->n, km->ctx->stream));
    FDB_TRY(km->centre.ensure(km->nb * km->m));
    FDB_TRY(km->centre_send.ensure(km->nb * km->m));
    FDB_TRY(km->u01.ensure(km->nb));
    FDB_TRY(km->shard_values.ensure(km->nb));
    FDB_TRY(km->shard_owner.ensure(km->nb));
    FDB_TRY(km->picked.ensure(km->nb * km->k));
    *d_totals = km->total.p;
    *d_pick = km->ci.p;
    *d_centre_send = km->centre_send.p;
    *d_u01 = km->u01.p;
    *d_picked_global = km->picked.p;
    return FDB_OK;
}

int fdb_kmeans_seed_sharded_total(fdb_km *km) {
    ARG(km, "km is null");
    FDB_TRY(km->ctx->use());
    return km_total_fast(km);
}

int fdb_kmeans_seed_sharded_pick(fdb_km *km, const float *d_all_totals, int world, int rank) {
    ARG(km && d_all_totals && world >= 1 && rank >= 0 && rank < world, "invalid argument");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb;
    fdb_shard::shard_split_kernel<<<(unsigned)((nb + 127) / 128), 128, 0, ctx->stream>>>(
        d_all_totals, world, rank, nb, km->u01.p, km->shard_values.p, km->shard_owner.p);
    ctx->launches++;
    FDB_TRY(km_seed_pick(km, km->shard_values.p, 1, 0, 0, 1));
    fdb_shard::shard_publish_kernel<<<(unsigned)nb, 128, 0, ctx->stream>>>(km->shard_owner.p, rank, nb, km->m, km->vs->d,
                                                               km->vs->dim, km->col_off, km->ci.p,
                                                               km->centre_send.p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int fdb_kmeans_seed_sharded_first(fdb_km *km, const uint32_t *local_first) {
    ARG(km && local_first, "null argument");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb;
    std::vector<int> own(nb);
    for (size_t b = 0; b < nb; ++b) {
        ARG(local_first[b] == 0xFFFFFFFFu || local_first[b] < km->n, "local index out of range");
        own[b] = local_first[b] == 0xFFFFFFFFu ? -1 : 0;
    }
    FDB_TRY(upload_small(ctx, km->ci.p, local_first, nb * sizeof(uint32_t)));
    FDB_TRY(upload_small(ctx, km->shard_owner.p, own.data(), nb * sizeof(int)));
    fdb_shard::shard_publish_kernel<<<(unsigned)nb, 128, 0, ctx->stream>>>(km->shard_owner.p, 0, nb, km->m, km->vs->d,
                                                               km->vs->dim, km->col_off, km->ci.p,
                                                               km->centre_send.p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int fdb_kmeans_seed_sharded_round(fdb_km *km, size_t i, const uint32_t *d_all_picks, const float *d_all_centres,
                                  int world, int rank, size_t n_global) {
    ARG(km && d_all_picks && d_all_centres && world >= 1 && rank >= 0 && rank < world, "invalid argument");
    ARG(i < km->k, "round %zu out of range", i);
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    fdb_shard::shard_adopt_kernel<<<(unsigned)km->nb, 128, 0, ctx->stream>>>(d_all_picks, d_all_centres, world, rank, km->nb,
                                                                 km->m, n_global, km->ci.p, km->centre.p,
                                                                 km->picked.p + i * km->nb, ctx->d_flags);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return km_seed_round(km, (uint32_t)i, 0, km->centre.p);
}

int fdb_kmeans_seed_sharded_finish(fdb_km *km, uint32_t *picked_global) {
    ARG(km && picked_global, "null argument");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb, k = km->k;
    std::vector<uint32_t> stage(nb * k);
    FDB_CUDA(cudaMemcpyAsync(stage.data(), km->picked.p, stage.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                             ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t b = 0; b < nb; ++b)
        for (size_t i = 0; i < k; ++i) picked_global[b * k + i] = stage[i * nb + b];
    unsigned f = 0;
    FDB_TRY(ctx->check_flags(&f));
    return map_flags(f & ~FLAG_WEIGHTS);  // a shard's own total may legitimately be zero
}

int fdb_kmeans_seed_run(fdb_km *km, const uint32_t *first, const float *u01, int exact,
                        uint32_t *picked) {
    ARG(km && first && (u01 || km->k == 1 || km->k == km->n), "null argument");
    return seed_loop(km, first, u01, nullptr, exact, picked);
}

int fdb_kmeans_seed_chosen(fdb_km *km, const uint32_t *chosen) {
    ARG(km && chosen, "null argument");
    return seed_loop(km, nullptr, nullptr, chosen, 0, nullptr);
}

int fdb_kmeans_set_state(fdb_km *km, const float *centroids, const uint32_t *indices) {
    ARG(km && centroids, "null argument");
    FDB_TRY(km->ctx->use());
    FDB_TRY(upload_small(km->ctx, km->centroids.p, centroids, km->nb * km->k * km->m * sizeof(float)));
    if (indices) {
        for (size_t i = 0; i < km->nb * km->n; ++i) ARG(indices[i] < km->k, "index out of range");
        FDB_TRY(upload_small(km->ctx, km->indices.p, indices, km->nb * km->n * sizeof(uint32_t)));
    }
    return FDB_OK;
}

static int upload_active(fdb_km *km, const uint8_t *active, const int **d_active) {
    *d_active = nullptr;
    if (!active) return FDB_OK;
    std::vector<int> a(km->nb);
    for (size_t b = 0; b < km->nb; ++b) a[b] = active[b] ? 1 : 0;
    FDB_TRY(upload_small(km->ctx, km->step_active.p, a.data(), km->nb * sizeof(int)));
    *d_active = km->step_active.p;
    return FDB_OK;
}

int fdb_kmeans_update(fdb_km *km, const uint8_t *active, float *gradients) {
    ARG(km && gradients, "null argument");
    FDB_TRY(km->ctx->use());
    const int *d_active;
    FDB_TRY(upload_active(km, active, &d_active));
    FDB_TRY(km_update(km, d_active, 0, -1.0f, km->max_rounds));
    FDB_CUDA(cudaMemcpyAsync(gradients, km->grad.p, km->nb * sizeof(float), cudaMemcpyDeviceToHost,
                             km->ctx->stream));
    return km_finish_call(km);
}

int fdb_kmeans_reassign(fdb_km *km, const uint8_t *active) {
    ARG(km, "km is null");
    FDB_TRY(km->ctx->use());
    const int *d_active;
    FDB_TRY(upload_active(km, active, &d_active));
    FDB_TRY(km_reassign(km, d_active));
    return km_finish_call(km);
}

int fdb_kmeans_run(fdb_km *km, size_t max_rounds, float epsilon, float *gradients, uint32_t *rounds,
                   uint32_t *reassigns) {
    ARG(km, "km is null");
    ARG(max_rounds <= km->max_rounds, "max_rounds exceeds %zu", km->max_rounds);
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb;
    FDB_TRY(km_fill_int(ctx, km->active.p, nb, 1));
    FDB_CUDA(cudaMemsetAsync(km->rounds.p, 0, nb * sizeof(uint32_t), ctx->stream));
    FDB_CUDA(cudaMemsetAsync(km->reassigns.p, 0, nb * sizeof(uint32_t), ctx->stream));
    FDB_CUDA(cudaMemsetAsync(km->grad_hist.p, 0, nb * km->max_rounds * sizeof(float), ctx->stream));
    // The host looks at the convergence flags of round r - LAG while it enqueues round r: the stream never drains
    // between rounds.  A converged problem clears its own flag on the device and is skipped by every later kernel,
    // so the rounds enqueued past its convergence change nothing.
    constexpr size_t LAG = 3, SLOTS = LAG + 1;
    int *h_active = (int *)ctx->h_pinned;
    if (SLOTS * nb * sizeof(int) > ctx->h_pinned_bytes / 2) {
        set_error("too many problems for the staging buffer");
        return FDB_ERR_UNSUPPORTED;
    }
    cudaEvent_t ev[SLOTS] = {nullptr, nullptr, nullptr, nullptr};
    for (auto &e : ev) FDB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    // One round = update + reassignment: ~16 stream operations, several of them shorter than a launch.  Round 0
    // runs eagerly (it allocates), round 1 is captured into a CUDA graph, the later rounds replay it.
    cudaGraphExec_t round_graph = nullptr;
    uint64_t graph_launches = 0;
    const bool use_graph = !getenv("FDB_NO_GRAPH") && !getenv("FDB_TC_STATS");
    auto one_round = [&]() -> int {
        // update; a problem whose gradient < epsilon clears its own active flag on the device,
        // so the reassignment that follows skips it (src/kmeans.rs:130-132)
        FDB_TRY(km_update(km, km->active.p, 1, epsilon, km->max_rounds));
        FDB_TRY(km_reassign(km, km->active.p));
        return FDB_OK;
    };
    int rc_loop = FDB_OK;
    for (size_t r = 0; r < max_rounds; ++r) {
        if (r >= LAG) {
            const size_t s = (r - LAG) % SLOTS;
            if (cudaEventSynchronize(ev[s]) != cudaSuccess) {
                set_error("cudaEventSynchronize failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc_loop = FDB_ERR_CUDA;
                break;
            }
            bool any = false;
            for (size_t b = 0; b < nb; ++b) any |= h_active[s * nb + b] != 0;
            if (!any) break;
        }
        if (round_graph) {
            if (cudaGraphLaunch(round_graph, ctx->stream) != cudaSuccess) {
                set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                rc_loop = FDB_ERR_CUDA;
                break;
            }
            ctx->launches += graph_launches;
        } else if (use_graph && r == 1) {
            const uint64_t l0 = ctx->launches;
            cudaGraph_t g = nullptr;
            bool ok = cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
            if (ok) {
                const int rc = one_round();
                const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
                ok = rc == FDB_OK && e == cudaSuccess && g != nullptr;
            }
            if (ok) ok = cudaGraphInstantiate(&round_graph, g, 0) == cudaSuccess;
            if (g) cudaGraphDestroy(g);
            if (!ok) {  // not capturable here: run the round the ordinary way
                cudaGetLastError();
                round_graph = nullptr;
                ctx->launches = l0;
                if ((rc_loop = one_round()) != FDB_OK) break;
            } else {
                graph_launches = ctx->launches - l0;
                if (cudaGraphLaunch(round_graph, ctx->stream) != cudaSuccess) {
                    set_error("cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                    rc_loop = FDB_ERR_CUDA;
                    break;
                }
            }
        } else {
            if ((rc_loop = one_round()) != FDB_OK) break;
        }
        const size_t s = r % SLOTS;
        if (cudaMemcpyAsync(h_active + s * nb, km->active.p, nb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
            cudaEventRecord(ev[s], ctx->stream) != cudaSuccess) {
            set_error("flag read-back failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc_loop = FDB_ERR_CUDA;
            break;
        }
    }
    for (auto &e : ev) cudaEventDestroy(e);
    if (round_graph) cudaGraphExecDestroy(round_graph);
    FDB_TRY(rc_loop);
    std::vector<float> gh(nb * km->max_rounds);
    if (gradients) {
        FDB_CUDA(cudaMemcpyAsync(gh.data(), km->grad_hist.p, gh.size() * sizeof(float),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (rounds)
        FDB_CUDA(cudaMemcpyAsync(rounds, km->rounds.p, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                 ctx->stream));
    if (reassigns)
        FDB_CUDA(cudaMemcpyAsync(reassigns, km->reassigns.p, nb * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, ctx->stream));
    int rc = km_finish_call(km);
    if (gradients)
        for (size_t b = 0; b < nb; ++b)
            memcpy(gradients + b * max_rounds, gh.data() + b * km->max_rounds, max_rounds * sizeof(float));
    return rc;
}

int fdb_kmeans_get(fdb_km *km, float *centroids, uint32_t *indices) {
    ARG(km, "km is null");
    FDB_TRY(km->ctx->use());
    if (centroids)
        FDB_CUDA(cudaMemcpyAsync(centroids, km->centroids.p, km->nb * km->k * km->m * sizeof(float),
                                 cudaMemcpyDeviceToHost, km->ctx->stream));
    if (indices)
        FDB_CUDA(cudaMemcpyAsync(indices, km->indices.p, km->nb * km->n * sizeof(uint32_t),
                                 cudaMemcpyDeviceToHost, km->ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(km->ctx->stream));
    return FDB_OK;
}

int fdb_kmeans_get_weights(fdb_km *km, float *weights) {
    ARG(km && weights, "null argument");
    ARG(km->weights.p, "seeding has not run");
    FDB_TRY(km->ctx->use());
    FDB_CUDA(cudaMemcpyAsync(weights, km->weights.p, km->nb * km->n * sizeof(float),
                             cudaMemcpyDeviceToHost, km->ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(km->ctx->stream));
    return FDB_OK;
}

int fdb_kmeans_last_assign_info(fdb_km *km, uint32_t *used_tensor_cores, uint32_t *rechecked_rows,
                                uint32_t *overflow_rows) {
    ARG(km, "km is null");
    if (used_tensor_cores) *used_tensor_cores = (uint32_t)km->last_assign_tc;
    unsigned st[3] = {0, 0, 0};
    if (km->last_assign_tc) tc_last_stats(km, st);
    if (rechecked_rows) *rechecked_rows = st[2];
    if (overflow_rows) *overflow_rows = st[1];
    return FDB_OK;
}

int fdb_kmeans_update_partial(fdb_km *km, float **device_buf, size_t *nfloats) {
    ARG(km && device_buf && nfloats, "null argument");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km_update_partial(km));
    FDB_CUDA(cudaStreamSynchronize(km->ctx->stream));  // the caller's collective runs on its own stream
    *device_buf = km->partial.p;
    *nfloats = km->nb * km->k * km->m + km->nb * km->k;
    return FDB_OK;
}

int fdb_kmeans_update_finish(fdb_km *km, float *gradients) {
    ARG(km && gradients, "null argument");
    ARG(km->partial.p, "update_partial has not been called");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km_update_finish(km));
    FDB_CUDA(cudaMemcpyAsync(gradients, km->grad.p, km->nb * sizeof(float), cudaMemcpyDeviceToHost,
                             km->ctx->stream));
    return km_finish_call(km);
}

/* ---- sharded Lloyd loop without host round trips -------------------------------------------------
 * begin; per round: partial (enqueue) -> all-reduce of the buffer on the context's stream -> finish
 * (divide, gradient, device-side convergence flags, reassignment of the still active problems);
 * poll reads the flags back (the only synchronisation, the caller chooses how often); end returns the
 * gradient history like fdb_kmeans_run. */
int fdb_kmeans_sharded_loop_begin(fdb_km *km) {
    ARG(km, "km is null");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb;
    FDB_TRY(km_fill_int(ctx, km->active.p, nb, 1));
    FDB_CUDA(cudaMemsetAsync(km->rounds.p, 0, nb * sizeof(uint32_t), ctx->stream));
    FDB_CUDA(cudaMemsetAsync(km->reassigns.p, 0, nb * sizeof(uint32_t), ctx->stream));
    FDB_CUDA(cudaMemsetAsync(km->grad_hist.p, 0, nb * km->max_rounds * sizeof(float), ctx->stream));
    return FDB_OK;
}

int fdb_kmeans_sharded_partial_async(fdb_km *km, float **device_buf, size_t *nfloats) {
    ARG(km && device_buf && nfloats, "null argument");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km_update_partial(km));
    *device_buf = km->partial.p;
    *nfloats = km->nb * km->k * km->m + km->nb * km->k;
    return FDB_OK;
}

int fdb_kmeans_sharded_finish_async(fdb_km *km, float epsilon) {
    ARG(km, "km is null");
    ARG(km->partial.p, "partial has not been called");
    FDB_TRY(km->ctx->use());
    FDB_TRY(km_update_finish_loop(km, km->active.p, 1, epsilon));
    return km_reassign(km, km->active.p);
}

int fdb_kmeans_sharded_poll(fdb_km *km, uint8_t *active) {
    ARG(km && active, "null argument");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    std::vector<int> a(km->nb);
    FDB_CUDA(cudaMemcpyAsync(a.data(), km->active.p, km->nb * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t b = 0; b < km->nb; ++b) active[b] = a[b] ? 1 : 0;
    return FDB_OK;
}

int fdb_kmeans_sharded_loop_end(fdb_km *km, float *gradients, uint32_t *rounds, uint32_t *reassigns) {
    ARG(km && gradients && rounds && reassigns, "null argument");
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(ctx->use());
    const size_t nb = km->nb;
    FDB_CUDA(cudaMemcpyAsync(gradients, km->grad_hist.p, nb * km->max_rounds * sizeof(float), cudaMemcpyDeviceToHost,
                             ctx->stream));
    FDB_CUDA(cudaMemcpyAsync(rounds, km->rounds.p, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    FDB_CUDA(cudaMemcpyAsync(reassigns, km->reassigns.p, nb * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return km_finish_call(km);
}

}  // extern "C"
