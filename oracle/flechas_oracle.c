/*
 * flechas_oracle.c -- CPU restatement of flechasdb's IVF-PQ build/query path.
 * TEST INFRASTRUCTURE ONLY (see flechas_oracle.h).  Parity for kmeans /
 * partitions / nbest / db is UNPINNED by the reference's own tests (it has none
 * for them); linalg / distribution / vector are pinned by the reference's KATs.
 *
 * Build: see oracle/Makefile (-ffp-contract=off, no fast-math).
 */
#include "flechas_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

/* minimal static-partition parallel-for (libgomp is not in this image).  Rows /
 * queries are independent, so results do not depend on the thread count. */
typedef void (*fo_range_fn)(size_t lo, size_t hi, void *arg);
typedef struct { fo_range_fn fn; size_t lo, hi; void *arg; } fo_job;
static void *fo_job_main(void *p) {
    fo_job *j = (fo_job *)p;
    j->fn(j->lo, j->hi, j->arg);
    return NULL;
}
static void parallel_for(size_t n, int nthreads, fo_range_fn fn, void *arg) {
    if (nthreads <= 1 || n < 2) { fn(0, n, arg); return; }
    if ((size_t)nthreads > n) nthreads = (int)n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    fo_job *jobs = (fo_job *)malloc(sizeof(fo_job) * (size_t)nthreads);
    for (int t = 0; t < nthreads; ++t) {
        jobs[t].fn = fn; jobs[t].arg = arg;
        jobs[t].lo = n * (size_t)t / (size_t)nthreads;
        jobs[t].hi = n * (size_t)(t + 1) / (size_t)nthreads;
        pthread_create(&th[t], NULL, fo_job_main, &jobs[t]);
    }
    for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}

#define UNROLL 16 /* src/linalg.rs:7 */

/* ======================================================================
 * src/linalg.rs
 * ==================================================================== */

/* sum_naive, src/linalg.rs:238-247 */
static float sum_naive(const float *x, size_t n) {
    float ans = 0.0f;
    for (size_t i = 0; i < n; ++i) ans += x[i];
    return ans;
}

/* dot_naive, src/linalg.rs:43-53 */
float fo_dot_naive(const float *x, const float *y, size_t n) {
    float ans = 0.0f;
    for (size_t i = 0; i < n; ++i) ans += x[i] * y[i];
    return ans;
}

/* dot, src/linalg.rs:12-40: 16 accumulators; the first len%16 elements seed
 * lanes 0..r-1 by assignment, full blocks of 16 follow, lanes are summed in
 * order at the end. */
float fo_dot(const float *x, const float *y, size_t n) {
    if (n < UNROLL) return fo_dot_naive(x, y, n);
    float acc[UNROLL];
    for (int j = 0; j < UNROLL; ++j) acc[j] = 0.0f;
    size_t r = n % UNROLL;
    for (size_t i = 0; i < r; ++i) acc[i] = x[i] * y[i];
    x += r;
    y += r;
    n -= r;
    for (size_t i = 0; i + UNROLL <= n; i += UNROLL) {
        for (int j = 0; j < UNROLL; ++j) acc[j] += x[i + j] * y[i + j];
    }
    return sum_naive(acc, UNROLL);
}

/* max_abs_naive :348-364 / max_abs :306-345 (value is order independent) */
int fo_max_abs(const float *x, size_t n, float *out) {
    if (n == 0) return 0;
    if (n < UNROLL) {
        float mx = fabsf(x[0]);
        for (size_t i = 1; i < n; ++i)
            if (fabsf(x[i]) > mx) mx = fabsf(x[i]);
        *out = mx;
        return 1;
    }
    float acc[UNROLL];
    for (int i = 0; i < UNROLL; ++i) acc[i] = fabsf(x[i]);
    x += UNROLL;
    n -= UNROLL;
    size_t r = n % UNROLL;
    for (size_t i = 0; i < r; ++i)
        if (fabsf(x[i]) > acc[i]) acc[i] = fabsf(x[i]);
    x += r;
    n -= r;
    for (size_t i = 0; i + UNROLL <= n; i += UNROLL)
        for (int j = 0; j < UNROLL; ++j)
            if (fabsf(x[i + j]) > acc[j]) acc[j] = fabsf(x[i + j]);
    float mx = acc[0];
    for (int i = 1; i < UNROLL; ++i)
        if (mx < acc[i]) mx = acc[i];
    *out = mx;
    return 1;
}

/* norm2_scaled_naive :108-118 */
static float norm2_scaled_naive(const float *x, size_t n, float a) {
    float acc = 0.0f;
    for (size_t i = 0; i < n; ++i) {
        float scaled = x[i] * a;
        acc += scaled * scaled;
    }
    return sqrtf(acc);
}

/* norm2_scaled :78-105 */
static float norm2_scaled(const float *x, size_t n, float a) {
    if (n < UNROLL) return norm2_scaled_naive(x, n, a);
    float acc[UNROLL];
    for (int j = 0; j < UNROLL; ++j) acc[j] = 0.0f;
    size_t r = n % UNROLL;
    for (size_t i = 0; i < r; ++i) {
        float scaled = a * x[i];
        acc[i] += scaled * scaled;
    }
    x += r;
    n -= r;
    for (size_t i = 0; i + UNROLL <= n; i += UNROLL)
        for (int j = 0; j < UNROLL; ++j) {
            float scaled = a * x[i + j];
            acc[j] += scaled * scaled;
        }
    return sqrtf(sum_naive(acc, UNROLL));
}

/* norm2 :61-75 */
float fo_norm2(const float *x, size_t n) {
    float mx;
    if (!fo_max_abs(x, n, &mx)) return 0.0f;
    if (mx == 0.0f) return 0.0f;
    float mx_sqrt = sqrtf(mx);
    return norm2_scaled(x, n, 1.0f / mx_sqrt) * mx_sqrt;
}

/* sum :208-235 */
float fo_sum(const float *x, size_t n) {
    if (n < UNROLL) return sum_naive(x, n);
    float acc[UNROLL];
    memcpy(acc, x, sizeof(acc));
    x += UNROLL;
    n -= UNROLL;
    size_t r = n % UNROLL;
    for (size_t i = 0; i < r; ++i) acc[i] += x[i];
    x += r;
    n -= r;
    for (size_t i = 0; i + UNROLL <= n; i += UNROLL)
        for (int j = 0; j < UNROLL; ++j) acc[j] += x[i + j];
    return sum_naive(acc, UNROLL);
}

/* min :252-283 / min_naive :286-303 */
int fo_min(const float *x, size_t n, float *out) {
    if (n == 0) return 0;
    if (n < UNROLL) {
        float mn = x[0];
        for (size_t i = 1; i < n; ++i)
            if (x[i] < mn) mn = x[i];
        *out = mn;
        return 1;
    }
    float acc[UNROLL];
    memcpy(acc, x, sizeof(acc));
    x += UNROLL;
    n -= UNROLL;
    size_t r = n % UNROLL;
    for (size_t i = 0; i < r; ++i)
        if (x[i] < acc[i]) acc[i] = x[i];
    x += r;
    n -= r;
    for (size_t i = 0; i + UNROLL <= n; i += UNROLL)
        for (int j = 0; j < UNROLL; ++j)
            if (x[i + j] < acc[j]) acc[j] = x[i + j];
    float mn = acc[0];
    for (int i = 1; i < UNROLL; ++i)
        if (acc[i] < mn) mn = acc[i];
    *out = mn;
    return 1;
}

void fo_add_in(float *l, const float *r, size_t n) {
    for (size_t i = 0; i < n; ++i) l[i] += r[i];
}
void fo_subtract(const float *l, const float *r, float *o, size_t n) {
    for (size_t i = 0; i < n; ++i) o[i] = l[i] - r[i];
}
void fo_subtract_in(float *l, const float *r, size_t n) {
    for (size_t i = 0; i < n; ++i) l[i] -= r[i];
}
void fo_scale_in(float *x, float a, size_t n) {
    for (size_t i = 0; i < n; ++i) x[i] *= a;
}

/* the pattern of every hot loop: subtract(v,c,d); dot(d,d)
 * (src/kmeans.rs:193-195,211-213,297-298; src/db/stored.rs:421-422,569-570) */
float fo_sqdist(const float *v, const float *c, size_t n, float *buf) {
    fo_subtract(v, c, buf, n);
    return fo_dot(buf, buf, n);
}

/* ======================================================================
 * src/vector.rs
 * ==================================================================== */

/* BlockVectorSet::chunk :40-57 */
int fo_chunk_check(size_t data_len, size_t vector_size) {
    if (vector_size == 0) return FO_ERR_INVALID_ARGS; /* NonZeroUsize */
    if (data_len == 0 || data_len % vector_size == 0) return FO_OK;
    return FO_ERR_INVALID_ARGS;
}

/* divide_vector_set :154-174 */
int fo_divide(const fo_view *vs, size_t d, fo_view *out) {
    if (d == 0 || vs->dim % d != 0) return FO_ERR_INVALID_ARGS;
    size_t m = vs->dim / d;
    for (size_t i = 0; i < d; ++i) {
        out[i] = *vs;
        out[i].off = vs->off + i * m;
        out[i].dim = m;
    }
    return FO_OK;
}

static inline const float *row(const fo_view *vs, size_t i) {
    return vs->base + i * vs->stride + vs->off;
}

/* ======================================================================
 * src/distribution.rs
 * ==================================================================== */

/* rand 0.8.5 UniformFloat<f32>::new(low=0, high): scale = high - low, then
 * decreased one ulp at a time while scale*max_rand+low >= high
 * (rand-0.8.5/src/distributions/uniform.rs, uniform_float_impl!::new). */
static float uniform_scale(float high) {
    const float max_rand = 1.0f - 1.1920929e-07f; /* (u32::MAX>>9) as [1,2) float - 1 */
    float scale = high - 0.0f;
    for (;;) {
        if (!(scale * max_rand + 0.0f >= high)) break;
        scale = nextafterf(scale, -INFINITY);
    }
    return scale;
}

int fo_wi_new(fo_wi *wi, const float *weights, size_t n) {
    memset(wi, 0, sizeof(*wi));
    if (n == 0) return FO_ERR_INVALID_ARGS; /* :36-38 */
    float mn;
    fo_min(weights, n, &mn);
    if (mn < 0.0f) return FO_ERR_INVALID_ARGS; /* :39-44 */
    float total = fo_sum(weights, n);          /* :45 */
    if (total <= 0.0f) return FO_ERR_INVALID_ARGS; /* :46-48 */
    wi->weights = (float *)malloc(n * sizeof(float));
    memcpy(wi->weights, weights, n * sizeof(float));
    wi->n = n;
    wi->total = total;
    wi->scale = uniform_scale(total);
    return FO_OK;
}

void fo_wi_free(fo_wi *wi) {
    free(wi->weights);
    wi->weights = NULL;
}

/* update :63-91 */
int fo_wi_update(fo_wi *wi, const size_t *idx, const float *w, size_t m) {
    float nt = wi->total;
    for (size_t j = 0; j < m; ++j) {
        if (idx[j] >= wi->n) return FO_ERR_INVALID_ARGS;
        if (w[j] < 0.0f) return FO_ERR_INVALID_ARGS;
        nt -= wi->weights[idx[j]];
        nt += w[j];
    }
    if (nt <= 0.0f) return FO_ERR_INVALID_ARGS;
    for (size_t j = 0; j < m; ++j) wi->weights[idx[j]] = w[j];
    wi->total = nt;
    wi->scale = uniform_scale(nt);
    return FO_OK;
}

float fo_wi_get_weight(const fo_wi *wi, size_t i) { return wi->weights[i]; }

/* sample :104-121, after the draw */
size_t fo_wi_pick(const fo_wi *wi, float sample) {
    float cum = 0.0f;
    size_t last = (size_t)-1;
    for (size_t i = 0; i < wi->n; ++i) {
        if (wi->weights[i] > 0.0f) {
            last = i;
            cum += wi->weights[i];
            if (cum > sample) break;
        }
    }
    return last; /* (size_t)-1 == the reference's unwrap() panic */
}

/* UniformFloat::sample: value0_1 * scale + low */
float fo_wi_sample_value(const fo_wi *wi, float u01) {
    return u01 * wi->scale + 0.0f;
}

/* ======================================================================
 * src/kmeans.rs
 * ==================================================================== */

int fo_kmeans_init(const fo_view *vs, size_t k, size_t first,
                   const uint32_t *chosen_in, const float *u01,
                   float *centroids, uint32_t *indices,
                   float *weights_out, uint32_t *picked_out) {
    size_t n = vs->n, m = vs->dim;
    if (n < k || k == 0) return FO_ERR_INVALID_ARGS; /* :116-120 */
    memset(indices, 0, n * sizeof(uint32_t));
    if (k == n) { /* :158-170 */
        for (size_t i = 0; i < n; ++i) {
            memcpy(centroids + i * m, row(vs, i), m * sizeof(float));
            indices[i] = (uint32_t)i;
            if (picked_out) picked_out[i] = (uint32_t)i;
        }
        return FO_OK;
    }
    unsigned char *chosen = (unsigned char *)calloc(n, 1);
    float *buf = (float *)malloc((m ? m : 1) * sizeof(float));
    float *weights = (float *)malloc(n * sizeof(float));
    int rc = FO_OK;
    size_t ci = first; /* :172 */
    chosen[ci] = 1;
    memcpy(centroids, row(vs, ci), m * sizeof(float));
    if (picked_out) picked_out[0] = (uint32_t)ci;
    if (k == 1) goto done; /* :176-184 */
    {
        const float *nc = row(vs, ci);
        for (size_t i = 0; i < n; ++i) { /* :188-198 */
            if (chosen[i]) weights[i] = 0.0f;
            else weights[i] = fo_sqdist(row(vs, i), nc, m, buf);
        }
    }
    fo_wi wi;
    if (fo_wi_new(&wi, weights, n) != FO_OK) { /* :199 unwrap */
        rc = FO_ERR_PANIC_WEIGHTS;
        goto done;
    }
    for (size_t i = 1; i < k; ++i) { /* :201-221 */
        if (chosen_in) ci = chosen_in[i - 1];
        else ci = fo_wi_pick(&wi, fo_wi_sample_value(&wi, u01[i - 1]));
        if (ci >= n) { rc = FO_ERR_PANIC_WEIGHTS; break; }
        chosen[ci] = 1;
        indices[ci] = (uint32_t)i;
        if (picked_out) picked_out[i] = (uint32_t)ci;
        const float *nc = row(vs, ci);
        memcpy(centroids + i * m, nc, m * sizeof(float));
        float zero = 0.0f;
        if (fo_wi_update(&wi, &ci, &zero, 1) != FO_OK) { /* :207 unwrap */
            rc = FO_ERR_PANIC_WEIGHTS;
            break;
        }
        for (size_t j = 0; j < n; ++j) {
            if (!chosen[j]) {
                float nw = fo_sqdist(row(vs, j), nc, m, buf);
                if (nw < fo_wi_get_weight(&wi, j)) {
                    if (fo_wi_update(&wi, &j, &nw, 1) != FO_OK) { /* :216 unwrap */
                        rc = FO_ERR_PANIC_WEIGHTS;
                        break;
                    }
                    indices[j] = (uint32_t)i;
                }
            }
        }
        if (rc != FO_OK) break;
    }
    if (weights_out) memcpy(weights_out, wi.weights, n * sizeof(float));
    fo_wi_free(&wi);
done:
    free(chosen);
    free(buf);
    free(weights);
    return rc;
}

/* update_centroids :232-276 */
int fo_kmeans_update(const fo_view *vs, size_t k, float *centroids,
                     const uint32_t *indices, float *gradient) {
    size_t n = vs->n, m = vs->dim;
    float *old = (float *)malloc((m ? m : 1) * sizeof(float));
    float max_distance = 0.0f, max_norm2 = 0.0f;
    /* O(n) bucket pass instead of the reference's O(k*n) filter; members are
     * still visited in ascending j per cluster, which is all that matters. */
    size_t *start = (size_t *)calloc(k + 1, sizeof(size_t));
    uint32_t *members = (uint32_t *)malloc((n ? n : 1) * sizeof(uint32_t));
    for (size_t j = 0; j < n; ++j) start[indices[j] + 1]++;
    for (size_t i = 0; i < k; ++i) start[i + 1] += start[i];
    {
        size_t *fill = (size_t *)malloc((k ? k : 1) * sizeof(size_t));
        memcpy(fill, start, k * sizeof(size_t));
        for (size_t j = 0; j < n; ++j) members[fill[indices[j]]++] = (uint32_t)j;
        free(fill);
    }
    int rc = FO_OK;
    for (size_t i = 0; i < k; ++i) {
        float *nc = centroids + i * m;
        memcpy(old, nc, m * sizeof(float));
        for (size_t e = 0; e < m; ++e) nc[e] = 0.0f;
        size_t count = start[i + 1] - start[i];
        for (size_t t = start[i]; t < start[i + 1]; ++t)
            fo_add_in(nc, row(vs, members[t]), m);
        if (count == 0) { rc = FO_ERR_PANIC_EMPTY_CLUSTER; break; } /* :259 */
        fo_scale_in(nc, 1.0f / (float)count, m);                    /* :260 */
        float cn = fo_norm2(nc, m);
        if (max_norm2 < cn) max_norm2 = cn;
        fo_subtract_in(old, nc, m);
        float dist = fo_norm2(old, m);
        if (max_distance < dist) max_distance = dist;
    }
    free(old);
    free(start);
    free(members);
    if (rc != FO_OK) return rc;
    *gradient = (max_norm2 != 0.0f) ? max_distance / max_norm2 : 0.0f;
    return FO_OK;
}

/* reassign_centroids :279-306 */
typedef struct {
    const fo_view *vs; size_t k; const float *centroids; uint32_t *indices; int bad;
} reassign_arg;
static void reassign_range(size_t lo, size_t hi, void *p) {
    reassign_arg *a = (reassign_arg *)p;
    size_t m = a->vs->dim;
    float *buf = (float *)malloc((m ? m : 1) * sizeof(float));
    for (size_t i = lo; i < hi; ++i) {
        const float *v = row(a->vs, i);
        float min_distance = INFINITY;
        long long min_index = -1;
        for (size_t j = 0; j < a->k; ++j) {
            float d = fo_sqdist(v, a->centroids + j * m, m, buf);
            if (d < min_distance) {
                min_distance = d;
                min_index = (long long)j;
            }
        }
        if (min_index < 0) __atomic_store_n(&a->bad, 1, __ATOMIC_RELAXED);
        else a->indices[i] = (uint32_t)min_index;
    }
    free(buf);
}
int fo_kmeans_reassign(const fo_view *vs, size_t k, const float *centroids,
                       uint32_t *indices, int nthreads) {
    reassign_arg a = {vs, k, centroids, indices, 0};
    parallel_for(vs->n, nthreads, reassign_range, &a);
    return a.bad ? FO_ERR_PANIC_NAN : FO_OK; /* :304 unwrap */
}

/* the loop of cluster_with_events :125-137 */
int fo_kmeans_lloyd(const fo_view *vs, size_t k, float *centroids,
                    uint32_t *indices, size_t max_rounds, float epsilon,
                    float *gradients, size_t *rounds, size_t *reassigns,
                    int nthreads) {
    size_t nu = 0, nr = 0;
    for (size_t r = 0; r < max_rounds; ++r) {
        float g;
        int rc = fo_kmeans_update(vs, k, centroids, indices, &g);
        if (rc != FO_OK) return rc;
        if (gradients) gradients[nu] = g;
        nu++;
        if (g < epsilon) break;
        rc = fo_kmeans_reassign(vs, k, centroids, indices, nthreads);
        if (rc != FO_OK) return rc;
        nr++;
    }
    if (rounds) *rounds = nu;
    if (reassigns) *reassigns = nr;
    return FO_OK;
}

/* ======================================================================
 * src/partitions.rs:128-138 -- v_j -= centroid[indices[j]] in place
 * ==================================================================== */
void fo_residues(fo_view *vs, size_t p, const float *centroids,
                 const uint32_t *indices) {
    (void)p;
    for (size_t j = 0; j < vs->n; ++j)
        fo_subtract_in(vs->base + j * vs->stride + vs->off,
                       centroids + (size_t)indices[j] * vs->dim, vs->dim);
}

/* ======================================================================
 * src/db/build.rs:446-482 -- Partition::new for every partition
 * ==================================================================== */
void fo_extract_partitions(size_t M, size_t P, size_t D,
                           const uint32_t *part_idx, const uint32_t *codes_dm,
                           uint64_t *offsets, uint32_t *order, uint32_t *codes_pm) {
    for (size_t p = 0; p <= P; ++p) offsets[p] = 0;
    for (size_t v = 0; v < M; ++v) offsets[part_idx[v] + 1]++;
    for (size_t p = 0; p < P; ++p) offsets[p + 1] += offsets[p];
    uint64_t *fill = (uint64_t *)malloc((P ? P : 1) * sizeof(uint64_t));
    memcpy(fill, offsets, P * sizeof(uint64_t));
    for (size_t v = 0; v < M; ++v) { /* ascending global index inside a partition */
        uint64_t pos = fill[part_idx[v]]++;
        order[pos] = (uint32_t)v;
        for (size_t di = 0; di < D; ++di)
            codes_pm[pos * D + di] = codes_dm[di * M + v]; /* codebooks[di].indices[vi] :468-470 */
    }
    free(fill);
}

/* ======================================================================
 * src/nbest.rs:52-64
 * ==================================================================== */
void fo_nbest_push(fo_nbest *nb, fo_item cand) {
    if (nb->len < nb->n) {
        nb->items[nb->len++] = cand;
        return;
    }
    for (;;) {
        size_t s = 0;
        for (; s < nb->len; ++s)
            if (cand.key < nb->items[s].key) break; /* first slot the candidate beats */
        if (s == nb->len) break;
        fo_item t = nb->items[s];
        nb->items[s] = cand;
        cand = t;
    }
}

/* stable merge sort by key == slice::sort_by(partial_cmp) */
static void stable_sort(fo_item *a, size_t n, fo_item *tmp) {
    if (n < 2) return;
    if (n <= 16) {
        for (size_t i = 1; i < n; ++i) {
            fo_item x = a[i];
            size_t j = i;
            while (j > 0 && x.key < a[j - 1].key) {
                a[j] = a[j - 1];
                --j;
            }
            a[j] = x;
        }
        return;
    }
    size_t h = n / 2;
    stable_sort(a, h, tmp);
    stable_sort(a + h, n - h, tmp);
    size_t i = 0, j = h, o = 0;
    while (i < h && j < n) tmp[o++] = (a[j].key < a[i].key) ? a[j++] : a[i++];
    while (i < h) tmp[o++] = a[i++];
    while (j < n) tmp[o++] = a[j++];
    memcpy(a, tmp, n * sizeof(fo_item));
}

static int any_nan(const fo_item *a, size_t n) {
    if (n < 2) return 0; /* no comparison happens, no unwrap() */
    for (size_t i = 0; i < n; ++i)
        if (a[i].key != a[i].key) return 1;
    return 0;
}

/* ======================================================================
 * query: src/db/stored.rs:394-442 (mode 0), src/db/build.rs:345-382 (mode 1)
 * ==================================================================== */
int fo_query_probe(const fo_index *ix, const float *q, size_t nprobe, int mode,
                   uint32_t *probe_part, float *probe_dist) {
    size_t N = ix->N, P = ix->P;
    if (nprobe > P) return FO_ERR_INVALID_ARGS;
    float *buf = (float *)malloc((N ? N : 1) * sizeof(float));
    size_t cap = mode ? P : nprobe;
    fo_item *items = (fo_item *)malloc((cap ? cap : 1) * sizeof(fo_item));
    fo_item *tmp = (fo_item *)malloc((cap ? cap : 1) * sizeof(fo_item));
    fo_nbest nb = {items, nprobe, 0};
    size_t len = 0;
    for (size_t pi = 0; pi < P; ++pi) {
        float d = fo_sqdist(q, ix->coarse + pi * N, N, buf);
        fo_item it = {d, (uint32_t)pi, 0};
        if (mode) items[len++] = it;
        else fo_nbest_push(&nb, it);
    }
    if (!mode) len = nb.len;
    int rc = FO_OK;
    if (any_nan(items, len)) rc = FO_ERR_PANIC_NAN;
    else {
        stable_sort(items, len, tmp);
        for (size_t i = 0; i < nprobe; ++i) {
            probe_part[i] = items[i].a;
            if (probe_dist) probe_dist[i] = items[i].key;
        }
    }
    free(buf);
    free(items);
    free(tmp);
    return rc;
}

/* distance table: src/db/stored.rs:556-573 == src/db/build.rs:525-541 */
void fo_query_table(const fo_index *ix, const float *q, uint32_t part, float *table) {
    size_t N = ix->N, D = ix->D, C = ix->C, s = N / D;
    float *loc = (float *)malloc((N ? N : 1) * sizeof(float));
    float *buf = (float *)malloc((s ? s : 1) * sizeof(float));
    fo_subtract(q, ix->coarse + (size_t)part * N, loc, N); /* localized :421 */
    for (size_t di = 0; di < D; ++di)
        for (size_t ci = 0; ci < C; ++ci)
            table[di * C + ci] =
                fo_sqdist(loc + di * s, ix->codebooks + (di * C + ci) * s, s, buf);
    free(loc);
    free(buf);
}

static int query_one(const fo_index *ix, const float *q, size_t k, size_t nprobe,
                     int mode, uint32_t *out_part, uint32_t *out_vidx,
                     float *out_dist, uint32_t *out_count) {
    size_t D = ix->D, C = ix->C;
    uint32_t *probes = (uint32_t *)malloc((nprobe ? nprobe : 1) * sizeof(uint32_t));
    int rc = fo_query_probe(ix, q, nprobe, mode, probes, NULL);
    if (rc != FO_OK) { free(probes); return rc; }
    float *table = (float *)malloc(D * C * sizeof(float));
    size_t total = 0;
    for (size_t i = 0; i < nprobe; ++i)
        total += (size_t)(ix->offsets[probes[i] + 1] - ix->offsets[probes[i]]);
    size_t cap = mode ? total : nprobe * k;
    fo_item *all = (fo_item *)malloc((cap ? cap : 1) * sizeof(fo_item));
    fo_item *tmp = (fo_item *)malloc(((cap > k ? cap : k) + 1) * sizeof(fo_item));
    fo_item *part_items = (fo_item *)malloc((k ? k : 1) * sizeof(fo_item));
    size_t nall = 0;
    for (size_t i = 0; i < nprobe; ++i) {
        uint32_t p = probes[i];
        fo_query_table(ix, q, p, table);
        size_t np = (size_t)(ix->offsets[p + 1] - ix->offsets[p]);
        const uint32_t *codes = ix->codes_pm + ix->offsets[p] * D;
        fo_nbest nb = {part_items, k, 0};
        for (size_t vi = 0; vi < np; ++vi) {
            float dist = 0.0f; /* :582-587 sequential f32 adds over divisions */
            for (size_t di = 0; di < D; ++di) dist += table[di * C + codes[vi * D + di]];
            fo_item it = {dist, p, (uint32_t)vi};
            if (mode) all[nall++] = it;
            else fo_nbest_push(&nb, it);
        }
        if (!mode)
            for (size_t t = 0; t < nb.len; ++t) all[nall++] = part_items[t];
    }
    size_t nres;
    if (mode) { /* build.rs:334-337: stable sort everything, truncate */
        if (any_nan(all, nall)) rc = FO_ERR_PANIC_NAN;
        else stable_sort(all, nall, tmp);
        nres = nall < k ? nall : k;
    } else { /* stored.rs:379-386: n_best_by_key then stable sort */
        fo_item *fin = tmp; /* reuse: needs k slots + scratch */
        fo_nbest nb = {part_items, k, 0};
        for (size_t t = 0; t < nall; ++t) fo_nbest_push(&nb, all[t]);
        nres = nb.len;
        if (any_nan(part_items, nres)) rc = FO_ERR_PANIC_NAN;
        else stable_sort(part_items, nres, fin);
        memcpy(all, part_items, nres * sizeof(fo_item));
    }
    if (rc == FO_OK) {
        for (size_t t = 0; t < nres; ++t) {
            out_part[t] = all[t].a;
            out_vidx[t] = all[t].b;
            out_dist[t] = all[t].key;
        }
        *out_count = (uint32_t)nres;
    }
    free(probes);
    free(table);
    free(all);
    free(tmp);
    free(part_items);
    return rc;
}

typedef struct {
    const fo_index *ix; const float *q; size_t k, nprobe; int mode;
    uint32_t *out_part, *out_vidx; float *out_dist; uint32_t *out_count; int rc;
} query_arg;
static void query_range(size_t lo, size_t hi, void *p) {
    query_arg *a = (query_arg *)p;
    for (size_t qi = lo; qi < hi; ++qi) {
        int rc = query_one(a->ix, a->q + qi * a->ix->N, a->k, a->nprobe, a->mode,
                           a->out_part + qi * a->k, a->out_vidx + qi * a->k,
                           a->out_dist + qi * a->k, a->out_count + qi);
        if (rc != FO_OK) __atomic_store_n(&a->rc, rc, __ATOMIC_RELAXED);
    }
}
int fo_query(const fo_index *ix, const float *q, size_t nq, size_t k, size_t nprobe,
             int mode, uint32_t *out_part, uint32_t *out_vidx, float *out_dist,
             uint32_t *out_count, int nthreads) {
    if (nprobe > ix->P) return FO_ERR_INVALID_ARGS; /* stored.rs:403-409 */
    query_arg a = {ix, q, k, nprobe, mode, out_part, out_vidx, out_dist, out_count, FO_OK};
    parallel_for(nq, nthreads, query_range, &a);
    return a.rc;
}

/* ======================================================================
 * synthetic data: counter-based uniform [0,1) with 24-bit mantissa, the
 * shape rand's rng.fill(&mut [f32]) produces (examples/build-random/src/main.rs:17-20)
 * ==================================================================== */
uint64_t fo_splitmix64(uint64_t seed, uint64_t i) {
    uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void fo_fill_uniform(float *out, size_t count, uint64_t seed, uint64_t start) {
    for (size_t i = 0; i < count; ++i)
        out[i] = (float)(fo_splitmix64(seed, start + i) >> 40) * 5.9604644775390625e-08f;
}
