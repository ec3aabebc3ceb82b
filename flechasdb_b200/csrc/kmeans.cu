// k-means on the device: k-means++ seeding, Lloyd update / reassignment.
// Replaces src/kmeans.rs:104-306 and src/distribution.rs:35-121 of the reference.
//
// All kernels work on `nb` independent problems side by side (the strided
// SubVectorSet views of one row-major vector set, src/vector.rs:103-174).
//
// Bit-exactness: distances use the reference's 16-lane summation order (see
// exact_dist.cu); centroid sums add the members of a cluster in ascending vector
// index (src/kmeans.rs:251-258) -- a stable radix sort groups the rows, then one thread
// per (cluster, dimension) adds them in order; the convergence value uses the
// max-abs-scaled norm2 of src/linalg.rs:61-105 lane for lane.  No float atomics.
#include "kmeans.cuh"

#include <algorithm>

namespace fdb {

namespace {

__device__ __forceinline__ float sq_acc(float acc, float x, float c) {
    float d = __fsub_rn(x, c);
    return __fadd_rn(acc, __fmul_rn(d, d));
}
__host__ __device__ __forceinline__ size_t minz(size_t a, size_t b) { return a < b ? a : b; }

// reference-order distance by a single thread (dot: src/linalg.rs:12-53)
__device__ float sqdist_thread(const float *__restrict__ x, const float *__restrict__ c, size_t m) {
    if (m < 16) {
        float a = 0.0f;
        for (size_t e = 0; e < m; ++e) a = sq_acc(a, x[e], c[e]);
        return a;
    }
    float acc[16];
#pragma unroll
    for (int l = 0; l < 16; ++l) acc[l] = 0.0f;
    const size_t r = m & 15;
#pragma unroll
    for (int l = 0; l < 16; ++l)
        if ((size_t)l < r) {
            float d = __fsub_rn(x[l], c[l]);
            acc[l] = __fmul_rn(d, d);
        }
    for (size_t base = r; base < m; base += 16) {
#pragma unroll
        for (int l = 0; l < 16; ++l) acc[l] = sq_acc(acc[l], x[base + l], c[base + l]);
    }
    float s = 0.0f;
#pragma unroll
    for (int l = 0; l < 16; ++l) s = __fadd_rn(s, acc[l]);
    return s;
}

// rand 0.8.5 UniformFloat<f32>::new(0, high).scale
__device__ float uniform_scale(float high) {
    const float max_rand = 1.0f - 1.1920929e-07f;
    float scale = high;
    while (__fadd_rn(__fmul_rn(scale, max_rand), 0.0f) >= high && scale > 0.0f)
        scale = __int_as_float(__float_as_int(scale) - 1);
    return scale;
}

struct SeedParams {
    const float *x;
    size_t n, ldx, col_off, m, nb, k;
    const uint32_t *ci;      // [nb] chosen vector of this round (0xFFFFFFFF: not in this shard)
    const float *centre;     // [nb][m] the chosen vectors when they may live on another rank, else null
    float *centroids;        // [nb][k][m]
    const float *w_old;      // [nb][n]
    float *w_new;            // [nb][n]
    uint32_t *indices;       // [nb][n]
    uint8_t *chosen;         // [nb][n]
    uint32_t round;          // i (0 = first centre)
};

// One k-means++ round (src/kmeans.rs:188-198 for round 0, :203-220 afterwards):
// distance of every vector to the newly chosen centre, weight = min(weight, distance).
// VEC: a quad of threads per (row, problem), float4 loads (m%16==0, aligned);
// otherwise one thread per (row, problem).
template <bool VEC>
__global__ void __launch_bounds__(256) seed_round_kernel(SeedParams p) {
    // block 0 also copies the chosen rows into the centroid table (:205-206)
    if (blockIdx.x == 0) {
        for (size_t t = threadIdx.x; t < p.nb * p.m; t += blockDim.x) {
            const size_t b = t / p.m, e = t - b * p.m;
            p.centroids[(b * p.k + p.round) * p.m + e] =
                p.centre ? p.centre[b * p.m + e] : p.x[(size_t)p.ci[b] * p.ldx + p.col_off + b * p.m + e];
        }
    }
    const size_t per = VEC ? 4 : 1;
    const size_t g = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / per;
    const bool valid = g < p.n * p.nb;
    const size_t row = valid ? g / p.nb : 0, b = valid ? g % p.nb : 0;
    const uint32_t ci = p.ci[b];
    const float *x = p.x + row * p.ldx + p.col_off + b * p.m;
    const float *c = p.centre ? p.centre + b * p.m : p.x + (size_t)ci * p.ldx + p.col_off + b * p.m;
    float d;
    bool writer;
    if (VEC) {
        const int tq = threadIdx.x & 3;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if (valid) {
            for (size_t e = 4 * tq; e < p.m; e += 16) {
                const float4 xv = *reinterpret_cast<const float4 *>(x + e);
                const float4 cv = *reinterpret_cast<const float4 *>(c + e);
                a0 = sq_acc(a0, xv.x, cv.x);
                a1 = sq_acc(a1, xv.y, cv.y);
                a2 = sq_acc(a2, xv.z, cv.z);
                a3 = sq_acc(a3, xv.w, cv.w);
            }
        }
        const int qbase = (threadIdx.x & 31) & ~3;
        float s = 0.0f;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (tq == t) {
                s = __fadd_rn(s, a0);
                s = __fadd_rn(s, a1);
                s = __fadd_rn(s, a2);
                s = __fadd_rn(s, a3);
            }
            s = __shfl_sync(0xffffffffu, s, qbase + t);
        }
        d = s;
        writer = valid && tq == 0;
    } else {
        d = valid ? sqdist_thread(x, c, p.m) : 0.0f;
        writer = valid;
    }
    if (!writer) return;
    const size_t o = b * p.n + row;
    if (row == ci) {  // chosen[ci] = true; indices[ci] = i; weight -> 0 (:203-207)
        p.chosen[o] = 1;
        p.indices[o] = p.round;
        p.w_new[o] = 0.0f;
        return;
    }
    if (p.round == 0) {
        p.indices[o] = 0;
        p.w_new[o] = d;  // :192-196 (not chosen)
        return;
    }
    const float w = p.w_old[o];
    if (!p.chosen[o] && d < w) {  // :208-219
        p.w_new[o] = d;
        p.indices[o] = p.round;
    } else {
        p.w_new[o] = w;
    }
}

// The same round for MANY SHORT problems side by side (PQ divisions of 16..64 elements: 48 x 16 at BASELINE configs[2]).
// The per-(problem, row) state lives in [nb][n] arrays, so the kernel above -- consecutive quads = consecutive problems of
// one row -- touches a separate 32-byte sector of w_old / w_new / indices / chosen for every 64 bytes of row it reads
// (measured 1.6 TB/s of row bytes).  Here a CTA takes R rows x all problems: distances go to a shared-memory tile, then
// the state is read and written with consecutive threads on consecutive rows of one problem (coalesced).
constexpr int SEED_TILE_R = 32;
__global__ void __launch_bounds__(256) seed_round_tile_kernel(SeedParams p) {
    extern __shared__ float dtile[];   // [nb][R + 1]
    constexpr int R = SEED_TILE_R;
    if (blockIdx.x == 0) {
        for (size_t t = threadIdx.x; t < p.nb * p.m; t += blockDim.x) {
            const size_t b = t / p.m, e = t - b * p.m;
            p.centroids[(b * p.k + p.round) * p.m + e] =
                p.centre ? p.centre[b * p.m + e] : p.x[(size_t)p.ci[b] * p.ldx + p.col_off + b * p.m + e];
        }
    }
    const size_t row0 = (size_t)blockIdx.x * R;
    const int nrows = (int)min((size_t)R, p.n - row0);
    const int nb = (int)p.nb, npairs = nrows * nb;
    // One THREAD per (row, problem): its 16 lanes of the reference's dot (src/linalg.rs:12-40) are 16 registers --
    // no shuffles, a fifth of the instructions of the quad-per-pair kernel.  Consecutive threads take consecutive
    // problems of a row, so a warp's loads cover 32 x m contiguous floats (every sector is used, through L1).
    const int blocks16 = (int)(p.m >> 4);
#pragma unroll 2
    for (int pair = threadIdx.x; pair < npairs; pair += blockDim.x) {
        const int r = pair / nb, b = pair - r * nb;
        const float4 *x4 = reinterpret_cast<const float4 *>(p.x + (row0 + r) * p.ldx + p.col_off + (size_t)b * p.m);
        const float4 *c4 = reinterpret_cast<const float4 *>(
            p.centre ? p.centre + (size_t)b * p.m : p.x + (size_t)p.ci[b] * p.ldx + p.col_off + (size_t)b * p.m);
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
        for (int blk = 0; blk < blocks16; ++blk) {
            float4 xv[4], cv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) xv[i] = x4[4 * blk + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) cv[i] = c4[4 * blk + i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                acc[4 * i + 0] = sq_acc(acc[4 * i + 0], xv[i].x, cv[i].x);
                acc[4 * i + 1] = sq_acc(acc[4 * i + 1], xv[i].y, cv[i].y);
                acc[4 * i + 2] = sq_acc(acc[4 * i + 2], xv[i].z, cv[i].z);
                acc[4 * i + 3] = sq_acc(acc[4 * i + 3], xv[i].w, cv[i].w);
            }
        }
        float sum = 0.0f;   // sum_naive over the 16 accumulators, src/linalg.rs:39
#pragma unroll
        for (int j = 0; j < 16; ++j) sum = __fadd_rn(sum, acc[j]);
        dtile[b * (R + 1) + r] = sum;
    }
    __syncthreads();
    // state update, consecutive threads on consecutive rows of one problem; V outputs per thread and step, their
    // loads first
    constexpr int V = 4;
    const int nout = nb * R;
    for (int t0 = threadIdx.x; t0 < nout; t0 += blockDim.x * V) {
        float w[V];
        uint8_t ch[V];
        uint32_t ci[V];
        size_t o[V];
        bool act[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const int t = t0 + v * (int)blockDim.x;
            const int b = t / R, r = t - b * R;
            act[v] = t < nout && r < nrows;
            o[v] = act[v] ? (size_t)b * p.n + row0 + r : 0;
            ci[v] = act[v] ? p.ci[b] : 0u;
            w[v] = (act[v] && p.round != 0) ? p.w_old[o[v]] : 0.0f;
            ch[v] = (act[v] && p.round != 0) ? p.chosen[o[v]] : (uint8_t)0;
        }
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (!act[v]) continue;
            const int t = t0 + v * (int)blockDim.x;
            const int b = t / R, r = t - b * R;
            const size_t row = row0 + r;
            const float d = dtile[b * (R + 1) + r];
            if (row == ci[v]) {  // chosen[ci] = true; indices[ci] = i; weight -> 0 (:203-207)
                p.chosen[o[v]] = 1;
                p.indices[o[v]] = p.round;
                p.w_new[o[v]] = 0.0f;
            } else if (p.round == 0) {
                p.indices[o[v]] = 0;
                p.w_new[o[v]] = d;  // :192-196 (not chosen)
            } else if (!ch[v] && d < w[v]) {  // :208-219
                p.w_new[o[v]] = d;
                p.indices[o[v]] = p.round;
            } else {
                p.w_new[o[v]] = w[v];
            }
        }
    }
}

// ---- WeightedIndex, exact mode: the reference's sequential f32 arithmetic ---------
// sum(): src/linalg.rs:208-235 (16 lanes, first 16 elements seed them)
__global__ void total_init_exact_kernel(const float *w, size_t n, float *total, unsigned *flags) {
    const size_t b = blockIdx.x;
    const float *x = w + b * n;
    __shared__ float acc[16];
    const int l = threadIdx.x;  // 16 threads
    if (n < 16) {
        if (l == 0) {
            float s = 0.0f;
            for (size_t i = 0; i < n; ++i) s = __fadd_rn(s, x[i]);
            total[b] = s;
            if (!(s > 0.0f)) atomicOr(flags, FLAG_WEIGHTS);
        }
        return;
    }
    float a = x[l];
    const size_t rest = n - 16, r = rest & 15;
    if ((size_t)l < r) a = __fadd_rn(a, x[16 + l]);
    for (size_t e = 16 + r + l; e < n; e += 16) a = __fadd_rn(a, x[e]);
    acc[l] = a;
    __syncthreads();
    if (l == 0) {
        float s = 0.0f;
        for (int j = 0; j < 16; ++j) s = __fadd_rn(s, acc[j]);
        total[b] = s;
        if (!(s > 0.0f)) atomicOr(flags, FLAG_WEIGHTS);  // WeightedIndex::new fails, :46-48
    }
}

// update(): total -= old; total += new, in call order (src/distribution.rs:65-77);
// calls happen for ci first (src/kmeans.rs:207) then for ascending j (:208-219)
__global__ void total_chain_exact_kernel(const float *w_old, const float *w_new, size_t n,
                                         const uint32_t *ci, float *total, unsigned *flags) {
    const size_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= gridDim.x * blockDim.x) return;
    const float *wo = w_old + b * n, *wn = w_new + b * n;
    float t = total[b];
    const uint32_t c = ci[b];
    t = __fsub_rn(t, wo[c]);
    t = __fadd_rn(t, 0.0f);
    bool bad = !(t > 0.0f);
    for (size_t j = 0; j < n; ++j) {
        if (j == c) continue;
        const float o = wo[j], nw = wn[j];
        if (nw < o) {
            t = __fsub_rn(t, o);
            t = __fadd_rn(t, nw);
            if (!(t > 0.0f)) bad = true;
        }
    }
    total[b] = t;
    if (bad) atomicOr(flags, FLAG_WEIGHTS);
}

// sample(): src/distribution.rs:104-121 with the draw u*scale
__global__ void pick_exact_kernel(const float *w, size_t n, const float *total, const float *u01,
                                  size_t u_stride, size_t u_off, uint32_t *ci, unsigned *flags) {
    const size_t b = blockIdx.x * blockDim.x + threadIdx.x;
    const float *x = w + b * n;
    const float sample = __fadd_rn(__fmul_rn(u01[b * u_stride + u_off], uniform_scale(total[b])), 0.0f);
    float cum = 0.0f;
    uint32_t last = 0xFFFFFFFFu;
    for (size_t i = 0; i < n; ++i) {
        const float v = x[i];
        if (v > 0.0f) {
            last = (uint32_t)i;
            cum = __fadd_rn(cum, v);
            if (cum > sample) break;
        }
    }
    if (last == 0xFFFFFFFFu) {
        atomicOr(flags, FLAG_WEIGHTS);
        last = 0;
    }
    ci[b] = last;
}

// ---- WeightedIndex, fast mode: deterministic parallel scan in double ----------------
constexpr int PICK_THREADS = 1024;
__global__ void __launch_bounds__(PICK_THREADS) pick_fast_kernel(const float *w, size_t n,
                                                                 const float *u01, size_t u_stride,
                                                                 size_t u_off, uint32_t *ci,
                                                                 float *total_out, unsigned *flags,
                                                                 int pick, int absolute) {
    const size_t b = blockIdx.x;
    const float *x = w + b * n;
    const int t = threadIdx.x;
    const size_t seg = (n + PICK_THREADS - 1) / PICK_THREADS;
    const size_t lo = minz((size_t)t * seg, n), hi = minz(lo + seg, n);
    double s = 0.0;
    for (size_t i = lo; i < hi; ++i) s += (double)x[i];
    __shared__ double part[PICK_THREADS];
    __shared__ double excl[PICK_THREADS];
    __shared__ int first_t;
    __shared__ unsigned last_nz, pick_s;
    part[t] = s;
    if (t == 0) {
        first_t = PICK_THREADS;
        last_nz = 0xFFFFFFFFu;
        pick_s = 0xFFFFFFFFu;
    }
    __syncthreads();
    // inclusive Hillis-Steele scan (fixed order => deterministic)
    for (int off = 1; off < PICK_THREADS; off <<= 1) {
        double v = part[t];
        if (t >= off) v += part[t - off];
        __syncthreads();
        part[t] = v;
        __syncthreads();
    }
    excl[t] = part[t] - s;
    const double total = part[PICK_THREADS - 1];
    const float total_f = (float)total;
    if (t == 0) {
        total_out[b] = total_f;
        if (!(total_f > 0.0f)) atomicOr(flags, FLAG_WEIGHTS);
    }
    if (!pick) return;
    // absolute: the caller passes the sample value itself (multi-GPU: the part of the global
    // draw that falls into this shard); a negative value means "not this shard"
    const float sample = absolute ? u01[b * u_stride + u_off]
                                  : __fadd_rn(__fmul_rn(u01[b * u_stride + u_off], uniform_scale(total_f)), 0.0f);
    const double sd = (double)sample;
    if (hi > lo && excl[t] + s > sd) atomicMin(&first_t, t);
    // last positive weight, for the case the scan never exceeds the sample
    unsigned ln = 0xFFFFFFFFu;
    for (size_t i = hi; i > lo; --i)
        if (x[i - 1] > 0.0f) {
            ln = (unsigned)(i - 1);
            break;
        }
    if (ln != 0xFFFFFFFFu) atomicMax((int *)&last_nz, (int)ln);  // indices < 2^31
    __syncthreads();
    if (t == first_t) {
        double cum = excl[t];
        uint32_t pickd = 0xFFFFFFFFu;
        for (size_t i = lo; i < hi; ++i) {
            const float v = x[i];
            if (v > 0.0f) {
                pickd = (uint32_t)i;
                cum += (double)v;
                if (cum > sd) break;
            }
        }
        pick_s = pickd;
    }
    __syncthreads();
    if (t == 0) {
        unsigned r = pick_s;
        if (r == 0xFFFFFFFFu) r = last_nz;  // the scan never exceeded the sample: last positive weight
        if (r == 0xFFFFFFFFu) {
            atomicOr(flags, FLAG_WEIGHTS);
            r = 0;
        }
        ci[b] = r;
    }
}

// Large n: the same sampler in two levels, so that the pass over the weights uses the whole GPU.
// weight_blocks_kernel: sum (double, fixed order) and last positive index of every block of
// PICK_BLOCK weights.  pick_blocks_kernel (one CTA per problem): scan of the block sums -> total and
// the block in which the running sum first exceeds the sample -> the element inside that block.
constexpr int PICK_BLOCK = 4096;
__global__ void __launch_bounds__(256) weight_blocks_kernel(const float *w, size_t n, size_t nblk, double *bsum,
                                                            unsigned *blast) {
    const size_t b = blockIdx.y, blk = blockIdx.x;
    const float *x = w + b * n;
    const size_t lo = blk * PICK_BLOCK, hi = minz(lo + PICK_BLOCK, n);
    double s = 0.0;
    unsigned last = 0xFFFFFFFFu;
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
        const float v = x[i];
        s += (double)v;
        if (v > 0.0f) last = (unsigned)i;
    }
    __shared__ double red[256];
    __shared__ int lastmax;
    red[threadIdx.x] = s;
    if (threadIdx.x == 0) lastmax = -1;
    __syncthreads();
    if (last != 0xFFFFFFFFu) atomicMax(&lastmax, (int)last);  // indices < 2^31
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        bsum[b * nblk + blk] = red[0];
        blast[b * nblk + blk] = lastmax < 0 ? 0xFFFFFFFFu : (unsigned)lastmax;
    }
}

__global__ void __launch_bounds__(PICK_THREADS) pick_blocks_kernel(const float *w, size_t n, size_t nblk,
                                                                   const double *bsum, const unsigned *blast,
                                                                   const float *u01, size_t u_stride, size_t u_off,
                                                                   uint32_t *ci, float *total_out, unsigned *flags,
                                                                   int pick, int absolute) {
    const size_t b = blockIdx.x;
    const float *x = w + b * n;
    const double *bs = bsum + b * nblk;
    const unsigned *bl = blast + b * nblk;
    const int t = threadIdx.x;
    const size_t seg = (nblk + PICK_THREADS - 1) / PICK_THREADS;
    const size_t lo = minz((size_t)t * seg, nblk), hi = minz(lo + seg, nblk);
    double s = 0.0;
    unsigned ln = 0xFFFFFFFFu;
    for (size_t i = lo; i < hi; ++i) {
        s += bs[i];
        if (bl[i] != 0xFFFFFFFFu) ln = bl[i];
    }
    __shared__ double part[PICK_THREADS];
    __shared__ int first_t;
    __shared__ unsigned last_nz, pick_s, blk_s;
    __shared__ double cum_s;
    part[t] = s;
    if (t == 0) {
        first_t = PICK_THREADS;
        last_nz = 0xFFFFFFFFu;
        pick_s = 0xFFFFFFFFu;
        blk_s = 0xFFFFFFFFu;
        cum_s = 0.0;
    }
    __syncthreads();
    for (int off = 1; off < PICK_THREADS; off <<= 1) {   // inclusive Hillis-Steele scan, fixed order
        double v = part[t];
        if (t >= off) v += part[t - off];
        __syncthreads();
        part[t] = v;
        __syncthreads();
    }
    const double excl = part[t] - s;
    const double total = part[PICK_THREADS - 1];
    const float total_f = (float)total;
    if (t == 0) {
        total_out[b] = total_f;
        if (!(total_f > 0.0f)) atomicOr(flags, FLAG_WEIGHTS);
    }
    if (!pick) return;
    const float sample = absolute ? u01[b * u_stride + u_off]
                                  : __fadd_rn(__fmul_rn(u01[b * u_stride + u_off], uniform_scale(total_f)), 0.0f);
    const double sd = (double)sample;
    if (hi > lo && excl + s > sd) atomicMin(&first_t, t);
    if (ln != 0xFFFFFFFFu) atomicMax((int *)&last_nz, (int)ln);
    __syncthreads();
    if (t == first_t) {   // the block in which the running sum first exceeds the sample
        double cum = excl;
        for (size_t i = lo; i < hi; ++i) {
            if (cum + bs[i] > sd) {
                blk_s = (unsigned)i;
                cum_s = cum;
                break;
            }
            cum += bs[i];
        }
    }
    __syncthreads();
    if (blk_s != 0xFFFFFFFFu) {
        // inside the block: 4 consecutive weights per thread, scan, then the sequential rule
        const size_t e0 = (size_t)blk_s * PICK_BLOCK + (size_t)t * (PICK_BLOCK / PICK_THREADS);
        const size_t e1 = minz(e0 + PICK_BLOCK / PICK_THREADS, n);
        double ls = 0.0;
        for (size_t i = e0; i < e1; ++i) ls += (double)x[i];
        __syncthreads();
        part[t] = ls;
        if (t == 0) first_t = PICK_THREADS;
        __syncthreads();
        for (int off = 1; off < PICK_THREADS; off <<= 1) {
            double v = part[t];
            if (t >= off) v += part[t - off];
            __syncthreads();
            part[t] = v;
            __syncthreads();
        }
        const double ex2 = cum_s + part[t] - ls;
        if (e1 > e0 && ex2 + ls > sd) atomicMin(&first_t, t);
        __syncthreads();
        if (t == first_t) {
            double cum = ex2;
            uint32_t pickd = 0xFFFFFFFFu;
            for (size_t i = e0; i < e1; ++i) {
                const float v = x[i];
                if (v > 0.0f) {
                    pickd = (uint32_t)i;
                    cum += (double)v;
                    if (cum > sd) break;
                }
            }
            pick_s = pickd;
        }
        __syncthreads();
        // (the element-wise sums may stay a hair below the block-wise ones: then the block's last
        //  positive weight is the one that crosses)
        if (t == 0 && pick_s == 0xFFFFFFFFu) pick_s = bl[blk_s];
    }
    __syncthreads();
    if (t == 0) {
        unsigned r = pick_s;
        if (r == 0xFFFFFFFFu) r = last_nz;  // the scan never exceeded the sample: last positive weight
        if (r == 0xFFFFFFFFu) {
            atomicOr(flags, FLAG_WEIGHTS);
            r = 0;
        }
        ci[b] = r;
    }
}

// ---- stable grouping of rows by cluster (LSD radix, 8-bit digits) -------------------
constexpr int SORT_CHUNK = 1024;

struct SortParams {
    const uint32_t *keys;     // [nb][n] cluster of each row
    const uint32_t *src;      // [nb][n] row order of the previous pass, or null = identity
    uint32_t *dst;            // [nb][n]
    uint32_t *hist;           // [nb][256][nchunks]
    const int *active;
    size_t n, nchunks;
    int shift;
};

__global__ void __launch_bounds__(256) sort_hist_kernel(SortParams p) {
    const size_t b = blockIdx.y, chunk = blockIdx.x;
    if (p.active && !p.active[b]) return;
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const size_t lo = chunk * SORT_CHUNK, hi = minz(lo + SORT_CHUNK, p.n);
    for (size_t i = lo + threadIdx.x; i < hi; i += 256) {
        const uint32_t row = p.src ? p.src[b * p.n + i] : (uint32_t)i;
        atomicAdd(&h[(p.keys[b * p.n + row] >> p.shift) & 255u], 1u);
    }
    __syncthreads();
    p.hist[(b * 256 + threadIdx.x) * p.nchunks + chunk] = h[threadIdx.x];
}

// exclusive scan over [digit][chunk] (digit-major) of one problem; one CTA per problem
__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t *hist, size_t count,
                                                         const int *active) {
    const size_t b = blockIdx.x;
    if (active && !active[b]) return;
    uint32_t *h = hist + b * count;
    __shared__ unsigned warp_sums[32];
    __shared__ unsigned carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (size_t base = 0; base < count; base += 1024) {
        const size_t i = base + threadIdx.x;
        const unsigned v = i < count ? h[i] : 0u;
        unsigned s = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned o = __shfl_up_sync(0xffffffffu, s, off);
            if (lane >= off) s += o;
        }
        if (lane == 31) warp_sums[warp] = s;
        __syncthreads();
        if (warp == 0) {
            unsigned ws = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned o = __shfl_up_sync(0xffffffffu, ws, off);
                if (lane >= off) ws += o;
            }
            warp_sums[lane] = ws;
        }
        __syncthreads();
        const unsigned before = carry + (warp ? warp_sums[warp - 1] : 0u) + s - v;
        if (i < count) h[i] = before;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + v;
        __syncthreads();
    }
}

// one warp walks its chunk in order; equal digits inside a group of 32 are ranked by lane
__global__ void __launch_bounds__(32) sort_scatter_kernel(SortParams p) {
    const size_t b = blockIdx.y, chunk = blockIdx.x;
    if (p.active && !p.active[b]) return;
    __shared__ unsigned off[256];
    const int lane = threadIdx.x;
    for (int d = lane; d < 256; d += 32) off[d] = p.hist[(b * 256 + d) * p.nchunks + chunk];
    __syncwarp();
    const size_t lo = chunk * SORT_CHUNK, hi = minz(lo + SORT_CHUNK, p.n);
    for (size_t base = lo; base < hi; base += 32) {
        const size_t i = base + lane;
        const bool valid = i < hi;
        uint32_t row = 0, d = 256 + lane;  // invalid lanes get unique pseudo digits
        if (valid) {
            row = p.src ? p.src[b * p.n + i] : (uint32_t)i;
            d = (p.keys[b * p.n + row] >> p.shift) & 255u;
        }
        const unsigned peers = __match_any_sync(0xffffffffu, d);
        const unsigned rank = __popc(peers & ((1u << lane) - 1u));
        unsigned basepos = 0;
        if (valid) basepos = off[d];
        __syncwarp();
        if (valid) {
            p.dst[b * p.n + basepos + rank] = row;
            if (rank == 0) off[d] = basepos + __popc(peers);
        }
        __syncwarp();
    }
}

// cluster sizes: a shared-memory histogram per CTA (k <= 4096), flushed with one global atomic per
// non-empty bin (one atomic per row on 256 hot counters is an order of magnitude slower)
constexpr int COUNT_ROWS_PER_CTA = 4096;
__global__ void __launch_bounds__(256) cluster_count_kernel(const uint32_t *keys, size_t n, size_t k, uint32_t *cnt,
                                                            const int *active) {
    extern __shared__ unsigned hist_s[];
    const size_t b = blockIdx.y;
    if (active && !active[b]) return;
    const size_t lo = (size_t)blockIdx.x * COUNT_ROWS_PER_CTA;
    const size_t hi = lo + COUNT_ROWS_PER_CTA < n ? lo + COUNT_ROWS_PER_CTA : n;
    if (k <= 4096) {
        for (size_t j = threadIdx.x; j < k; j += blockDim.x) hist_s[j] = 0;
        __syncthreads();
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&hist_s[keys[b * n + i]], 1u);
        __syncthreads();
        for (size_t j = threadIdx.x; j < k; j += blockDim.x)
            if (hist_s[j]) atomicAdd(&cnt[b * (k + 1) + j], hist_s[j]);
    } else {
        for (size_t i = lo + threadIdx.x; i < hi; i += blockDim.x) atomicAdd(&cnt[b * (k + 1) + keys[b * n + i]], 1u);
    }
}

// ---- update_centroids (src/kmeans.rs:232-276) -----------------------------------------
struct UpdateParams {
    const float *x;
    size_t n, ldx, col_off, m, nb, k;
    const uint32_t *members;   // [nb][n] rows grouped by cluster, ascending inside
    const uint32_t *cl_off;    // [nb][k+1]
    float *centroids;          // [nb][k][m] in: old, out: new
    float *old_centroids;      // [nb][k][m]
    float *partial;            // multi-GPU: sums [nb][k][m] ++ counts [nb][k]; else null
    const int *active;
    unsigned *flags;
};

// One thread per (cluster, dimension): the members are added in ascending row order, so the chain is
// sequential by construction.  m >= 128: blockIdx.x tiles the dimensions of one cluster; smaller m: a
// CTA takes 128 / mpad clusters (mpad = m rounded up to a power of two) so that all its lanes work.
__global__ void __launch_bounds__(128) accumulate_kernel(UpdateParams p, unsigned mpad) {
    const size_t b = blockIdx.z;
    if (p.active && !p.active[b]) return;
    size_t i, e;
    if (mpad >= 128) {
        i = blockIdx.y;
        e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    } else {
        i = (size_t)blockIdx.y * (128 / mpad) + threadIdx.x / mpad;
        e = threadIdx.x % mpad;
        if (i >= p.k) return;
    }
    if (e >= p.m) return;
    const uint32_t *mem = p.members + b * p.n;
    const size_t start = p.cl_off[b * (p.k + 1) + i], end = p.cl_off[b * (p.k + 1) + i + 1];
    const float *xb = p.x + p.col_off + b * p.m + e;
    float s = 0.0f;  // new_centroid.fill(0) then add_in per member, ascending j (:249-258)
    size_t t = start;
    // (tried in round 2: 16 members in flight with the next batch's indices requested ahead -- the README build went
    //  from 0.151 to 0.166 s; with the last < 16 members batched under predicates as well the kernel took 1.27 ms
    //  instead of 0.22 ms: the straightforward batches of 8 stay)
    for (; t + 8 <= end; t += 8) {
        uint32_t r[8];
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = mem[t + u];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = xb[(size_t)r[u] * p.ldx];
#pragma unroll
        for (int u = 0; u < 8; ++u) s = __fadd_rn(s, v[u]);
    }
    for (; t < end; ++t) s = __fadd_rn(s, xb[(size_t)mem[t] * p.ldx]);
    const size_t o = (b * p.k + i) * p.m + e;
    const size_t count = end - start;
    if (p.partial) {
        p.partial[o] = s;
        if (e == 0) p.partial[p.nb * p.k * p.m + b * p.k + i] = (float)count;
        return;
    }
    if (count == 0 && e == 0) atomicOr(p.flags, FLAG_EMPTY_CLUSTER);  // assert_ne!(count, 0) :259
    const float inv = __fdiv_rn(1.0f, (float)count);                  // T::one() / T::from_as(count) :260
    p.old_centroids[o] = p.centroids[o];
    p.centroids[o] = __fmul_rn(s, inv);
}

// multi-GPU: after the all-reduce of sums/counts
__global__ void finish_partial_kernel(UpdateParams p) {
    const size_t b = blockIdx.z, i = blockIdx.y;
    if (p.active && !p.active[b]) return;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.m) return;
    const size_t o = (b * p.k + i) * p.m + e;
    const float count = p.partial[p.nb * p.k * p.m + b * p.k + i];
    if (count == 0.0f && e == 0) atomicOr(p.flags, FLAG_EMPTY_CLUSTER);
    const float inv = __fdiv_rn(1.0f, count);
    p.old_centroids[o] = p.centroids[o];
    p.centroids[o] = __fmul_rn(p.partial[o], inv);
}

// norm2 (src/linalg.rs:61-105) by a half warp: lane l owns accumulator l
__device__ float norm2_half(const float *x, const float *sub, size_t m, int l, unsigned mask) {
    // value(e) = x[e] (- sub[e] when sub: subtract_in(old, new), src/kmeans.rs:265)
    auto val = [&](size_t e) { return sub ? __fsub_rn(x[e], sub[e]) : x[e]; };
    float mx = 0.0f;
    for (size_t e = l; e < m; e += 16) mx = fmaxf(mx, fabsf(val(e)));
#pragma unroll
    for (int off = 8; off >= 1; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(mask, mx, off));
    if (m == 0 || mx == 0.0f) return 0.0f;
    const float mx_sqrt = __fsqrt_rn(mx);
    const float a = __fdiv_rn(1.0f, mx_sqrt);
    float total = 0.0f;
    if (m < 16) {  // norm2_scaled_naive :108-118
        if ((l & 15) == 0) {
            float acc = 0.0f;
            for (size_t e = 0; e < m; ++e) {
                const float sc = __fmul_rn(val(e), a);
                acc = __fadd_rn(acc, __fmul_rn(sc, sc));
            }
            total = acc;
        }
        total = __shfl_sync(mask, total, (threadIdx.x & 16));
    } else {
        const size_t r = m & 15;
        float acc = 0.0f;
        if ((size_t)l < r) {
            const float sc = __fmul_rn(a, val(l));
            acc = __fadd_rn(acc, __fmul_rn(sc, sc));
        }
        for (size_t e = r + l; e < m; e += 16) {
            const float sc = __fmul_rn(a, val(e));
            acc = __fadd_rn(acc, __fmul_rn(sc, sc));
        }
        float s = 0.0f;
#pragma unroll
        for (int j = 0; j < 16; ++j) s = __fadd_rn(s, __shfl_sync(mask, acc, (threadIdx.x & 16) + j));
        total = s;
    }
    return __fmul_rn(__fsqrt_rn(total), mx_sqrt);
}

__global__ void __launch_bounds__(32) norms_kernel(const float *cent, const float *old, size_t k,
                                                   size_t m, float *cnorm, float *cdist,
                                                   const int *active) {
    const size_t b = blockIdx.y, i = blockIdx.x;
    if (active && !active[b]) return;
    const float *nc = cent + (b * k + i) * m;
    const float *oc = old + (b * k + i) * m;
    const int lane = threadIdx.x, l = lane & 15;
    const unsigned mask = lane < 16 ? 0x0000ffffu : 0xffff0000u;
    float v;
    if (lane < 16) v = norm2_half(nc, nullptr, m, l, mask);  // norm2(new_centroid) :261
    else v = norm2_half(oc, nc, m, l, mask);                   // norm2(old - new)    :265-266
    if (lane == 0) cnorm[b * k + i] = v;
    if (lane == 16) cdist[b * k + i] = v;
}

// gradient = max distance / max norm (src/kmeans.rs:262-275) and the loop bookkeeping
// of cluster_with_events (:125-137): a problem whose gradient < epsilon stops.
__global__ void __launch_bounds__(256) gradient_kernel(const float *cnorm, const float *cdist,
                                                       size_t k, float eps, int loop_mode,
                                                       size_t max_rounds, float *grad,
                                                       float *grad_hist, uint32_t *rounds,
                                                       uint32_t *reassigns, int *active) {
    const size_t b = blockIdx.x;
    if (active && !active[b]) return;
    float mn = 0.0f, md = 0.0f;
    for (size_t i = threadIdx.x; i < k; i += 256) {
        mn = fmaxf(mn, cnorm[b * k + i]);
        md = fmaxf(md, cdist[b * k + i]);
    }
    __shared__ float smn[256], smd[256];
    smn[threadIdx.x] = mn;
    smd[threadIdx.x] = md;
    __syncthreads();
    for (int off = 128; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) {
            smn[threadIdx.x] = fmaxf(smn[threadIdx.x], smn[threadIdx.x + off]);
            smd[threadIdx.x] = fmaxf(smd[threadIdx.x], smd[threadIdx.x + off]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float g = smn[0] != 0.0f ? __fdiv_rn(smd[0], smn[0]) : 0.0f;
        grad[b] = g;
        if (loop_mode) {
            const uint32_t r = rounds[b];
            grad_hist[b * max_rounds + r] = g;
            rounds[b] = r + 1;
            if (g < eps) active[b] = 0;       // break before the reassignment :130-132
            else reassigns[b] += 1;           // the reassignment that follows
        }
    }
}

__global__ void residual_kernel(float *x, size_t n, size_t ldx, size_t m, const float *cent,
                                const uint32_t *idx) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * m) return;
    const size_t row = t / m, e = t - row * m;
    // subtract_in(v, centroid), src/partitions.rs:135-136
    x[row * ldx + e] = __fsub_rn(x[row * ldx + e], cent[(size_t)idx[row] * m + e]);
}

__global__ void fill_int_kernel(int *p, size_t n, int v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// ---------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------

int km_sort_members(fdb_km *km, const int *d_active) {
    fdb_ctx *ctx = km->ctx;
    const size_t n = km->n, nb = km->nb, k = km->k;
    const size_t nchunks = (n + SORT_CHUNK - 1) / SORT_CHUNK;
    FDB_TRY(km->hist.ensure(nb * 256 * nchunks));
    FDB_TRY(km->members.ensure(nb * n));
    FDB_TRY(km->cl_off.ensure(nb * (k + 1)));
    int passes = 1;
    while (passes < 4 && (k - 1) >> (8 * passes)) passes++;
    if (passes > 1) FDB_TRY(km->members_tmp.ensure(nb * n));
    // cluster offsets = exclusive scan of the per-cluster counts
    FDB_CUDA(cudaMemsetAsync(km->cl_off.p, 0, nb * (k + 1) * sizeof(uint32_t), ctx->stream));
    {
        dim3 grid((unsigned)((n + COUNT_ROWS_PER_CTA - 1) / COUNT_ROWS_PER_CTA), (unsigned)nb);
        cluster_count_kernel<<<grid, 256, k <= 4096 ? k * sizeof(unsigned) : 0, ctx->stream>>>(
            km->indices.p, n, k, km->cl_off.p, d_active);
        ctx->launches++;
        sort_scan_kernel<<<(unsigned)nb, 1024, 0, ctx->stream>>>(km->cl_off.p, k + 1, d_active);
        ctx->launches++;
    }
    // ping-pong so that the last pass lands in km->members
    uint32_t *bufs[2] = {km->members.p, km->members_tmp.p};
    int cur = (passes & 1) ? 0 : 1;  // destination of pass 0
    const uint32_t *src = nullptr;
    for (int ps = 0; ps < passes; ++ps) {
        SortParams p;
        p.keys = km->indices.p;
        p.src = src;
        p.dst = bufs[cur];
        p.hist = km->hist.p;
        p.active = d_active;
        p.n = n;
        p.nchunks = nchunks;
        p.shift = 8 * ps;
        dim3 grid((unsigned)nchunks, (unsigned)nb);
        sort_hist_kernel<<<grid, 256, 0, ctx->stream>>>(p);
        sort_scan_kernel<<<(unsigned)nb, 1024, 0, ctx->stream>>>(km->hist.p, 256 * nchunks, d_active);
        sort_scatter_kernel<<<grid, 32, 0, ctx->stream>>>(p);
        ctx->launches += 3;
        src = bufs[cur];
        cur ^= 1;
    }
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

struct AccGrid {
    dim3 grid;
    unsigned mpad;
};
static AccGrid acc_grid(const fdb_km *km) {
    AccGrid a;
    if (km->m >= 128) {
        a.mpad = 128;
        a.grid = dim3((unsigned)((km->m + 127) / 128), (unsigned)km->k, (unsigned)km->nb);
    } else {
        unsigned mp = 1;
        while (mp < km->m) mp <<= 1;
        a.mpad = mp;
        const unsigned cpb = 128 / mp;
        a.grid = dim3(1, (unsigned)((km->k + cpb - 1) / cpb), (unsigned)km->nb);
    }
    return a;
}

static UpdateParams make_update_params(fdb_km *km, const int *d_active, float *partial) {
    UpdateParams p;
    p.x = km->vs->d;
    p.n = km->n;
    p.ldx = km->vs->dim;
    p.col_off = km->col_off;
    p.m = km->m;
    p.nb = km->nb;
    p.k = km->k;
    p.members = km->members.p;
    p.cl_off = km->cl_off.p;
    p.centroids = km->centroids.p;
    p.old_centroids = km->old_centroids.p;
    p.partial = partial;
    p.active = d_active;
    p.flags = km->ctx->d_flags;
    return p;
}

static int km_norms_and_gradient(fdb_km *km, const int *d_active, int loop_mode, float eps,
                                 size_t max_rounds) {
    fdb_ctx *ctx = km->ctx;
    dim3 gn((unsigned)km->k, (unsigned)km->nb);
    norms_kernel<<<gn, 32, 0, ctx->stream>>>(km->centroids.p, km->old_centroids.p, km->k, km->m,
                                             km->cnorm.p, km->cdist.p, d_active);
    gradient_kernel<<<(unsigned)km->nb, 256, 0, ctx->stream>>>(
        km->cnorm.p, km->cdist.p, km->k, eps, loop_mode, max_rounds, km->grad.p, km->grad_hist.p,
        km->rounds.p, km->reassigns.p, const_cast<int *>(d_active));
    ctx->launches += 2;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

// update_centroids for the active problems; loop_mode: also do the loop bookkeeping
int km_update(fdb_km *km, const int *d_active, int loop_mode, float eps, size_t max_rounds) {
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(km_sort_members(km, d_active));
    UpdateParams p = make_update_params(km, d_active, nullptr);
    const AccGrid ag = acc_grid(km);
    accumulate_kernel<<<ag.grid, 128, 0, ctx->stream>>>(p, ag.mpad);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return km_norms_and_gradient(km, d_active, loop_mode, eps, max_rounds);
}

int km_reassign(fdb_km *km, const int *d_active) {
    // tensor-core filter + exact re-check when the shape allows it, identical results
    km->last_assign_tc = 0;
    if (tc_eligible(km)) {
        km->last_assign_tc = 1;
        return tc_reassign(km, d_active);
    }
    DistProblem q;
    q.x = km->vs->d;
    q.n = km->n;
    q.ldx = km->vs->dim;
    q.col_off = km->col_off;
    q.m = km->m;
    q.nb = km->nb;
    q.c = km->centroids.p;
    q.k = km->k;
    q.active = d_active;
    return launch_exact_argmin(km->ctx, q, km->indices.p, km->n);
}

int km_seed_round(fdb_km *km, uint32_t round, int exact, const float *d_centre) {
    fdb_ctx *ctx = km->ctx;
    SeedParams p;
    p.centre = d_centre;
    p.x = km->vs->d;
    p.n = km->n;
    p.ldx = km->vs->dim;
    p.col_off = km->col_off;
    p.m = km->m;
    p.nb = km->nb;
    p.k = km->k;
    p.ci = km->ci.p;
    p.centroids = km->centroids.p;
    p.w_old = km->weights.p;
    p.w_new = km->weights_new.p;
    p.indices = km->indices.p;
    p.chosen = km->chosen.p;
    p.round = round;
    const bool vec = (km->m % 16 == 0) && (p.ldx % 4 == 0) && (p.col_off % 4 == 0) &&
                     ((uintptr_t)p.x % 16 == 0) && (!d_centre || (uintptr_t)d_centre % 16 == 0);
    const size_t threads = km->n * km->nb * (vec ? 4 : 1);
    const unsigned grid = (unsigned)((threads + 255) / 256);
    const size_t tile_smem = km->nb * (SEED_TILE_R + 1) * sizeof(float);
    static const bool no_tile = getenv("FDB_SEED_NO_TILE") != nullptr;
    if (vec && km->nb >= 2 && km->m <= 64 && tile_smem <= 48 * 1024 && !no_tile)
        seed_round_tile_kernel<<<(unsigned)((km->n + SEED_TILE_R - 1) / SEED_TILE_R), 256, tile_smem, ctx->stream>>>(p);
    else if (vec) seed_round_kernel<true><<<grid, 256, 0, ctx->stream>>>(p);
    else seed_round_kernel<false><<<grid, 256, 0, ctx->stream>>>(p);
    ctx->launches++;
    if (exact) {
        if (round == 0) {
            total_init_exact_kernel<<<(unsigned)km->nb, 16, 0, ctx->stream>>>(
                km->weights_new.p, km->n, km->total.p, ctx->d_flags);
        } else {
            total_chain_exact_kernel<<<(unsigned)km->nb, 1, 0, ctx->stream>>>(
                km->weights.p, km->weights_new.p, km->n, km->ci.p, km->total.p, ctx->d_flags);
        }
        ctx->launches++;
    }
    std::swap(km->weights.p, km->weights_new.p);
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

// the parallel sampler / total: one CTA per problem for small n, two levels for large n
static int launch_pick_fast(fdb_km *km, const float *d_u01, size_t u_stride, size_t u_off, int pick, int absolute) {
    fdb_ctx *ctx = km->ctx;
    if (km->n < 65536) {
        pick_fast_kernel<<<(unsigned)km->nb, PICK_THREADS, 0, ctx->stream>>>(
            km->weights.p, km->n, d_u01, u_stride, u_off, km->ci.p, km->total.p, ctx->d_flags, pick, absolute);
        ctx->launches++;
        return FDB_OK;
    }
    const size_t nblk = (km->n + PICK_BLOCK - 1) / PICK_BLOCK;
    FDB_TRY(km->pick_bsum.ensure(km->nb * nblk));
    FDB_TRY(km->pick_blast.ensure(km->nb * nblk));
    dim3 grid((unsigned)nblk, (unsigned)km->nb);
    weight_blocks_kernel<<<grid, 256, 0, ctx->stream>>>(km->weights.p, km->n, nblk, km->pick_bsum.p, km->pick_blast.p);
    pick_blocks_kernel<<<(unsigned)km->nb, PICK_THREADS, 0, ctx->stream>>>(
        km->weights.p, km->n, nblk, km->pick_bsum.p, km->pick_blast.p, d_u01, u_stride, u_off, km->ci.p, km->total.p,
        ctx->d_flags, pick, absolute);
    ctx->launches += 2;
    return FDB_OK;
}

int km_seed_pick(fdb_km *km, const float *d_u01, size_t u_stride, size_t u_off, int exact,
                 int absolute) {
    fdb_ctx *ctx = km->ctx;
    if (exact && !absolute) {
        pick_exact_kernel<<<(unsigned)km->nb, 1, 0, ctx->stream>>>(
            km->weights.p, km->n, km->total.p, d_u01, u_stride, u_off, km->ci.p, ctx->d_flags);
        ctx->launches++;
    } else {
        FDB_TRY(launch_pick_fast(km, d_u01, u_stride, u_off, 1, absolute));
    }
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int km_total_fast(fdb_km *km) {
    FDB_TRY(launch_pick_fast(km, nullptr, 0, 0, 0, 0));
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int km_fill_int(fdb_ctx *ctx, int *p, size_t n, int v) {
    if (!n) return FDB_OK;
    fill_int_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, n, v);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int km_update_partial(fdb_km *km) {
    fdb_ctx *ctx = km->ctx;
    FDB_TRY(km->partial.ensure(km->nb * km->k * km->m + km->nb * km->k));
    FDB_TRY(km_sort_members(km, nullptr));
    UpdateParams p = make_update_params(km, nullptr, km->partial.p);
    const AccGrid ag = acc_grid(km);
    accumulate_kernel<<<ag.grid, 128, 0, ctx->stream>>>(p, ag.mpad);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

int km_update_finish(fdb_km *km) { return km_update_finish_loop(km, nullptr, 0, -1.0f); }

// loop_mode: the gradient kernel keeps the per-problem active flags, round counters and gradient
// history on the device (problems that converged are frozen: their flag is already 0)
int km_update_finish_loop(fdb_km *km, const int *d_active, int loop_mode, float eps) {
    fdb_ctx *ctx = km->ctx;
    UpdateParams p = make_update_params(km, d_active, km->partial.p);
    dim3 grid((unsigned)((km->m + 127) / 128), (unsigned)km->k, (unsigned)km->nb);
    finish_partial_kernel<<<grid, 128, 0, ctx->stream>>>(p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return km_norms_and_gradient(km, d_active, loop_mode, eps, loop_mode ? km->max_rounds : 1);
}

int km_residuals(fdb_vs *vs, const fdb_km *km) {
    fdb_ctx *ctx = vs->ctx;
    const size_t total = vs->n * vs->dim;
    if (!total) return FDB_OK;
    residual_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(
        vs->d, vs->n, vs->dim, vs->dim, km->centroids.p, km->indices.p);
    vs->version++;
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

}  // namespace fdb
