// Exact squared distances in the reference's order of floating-point operations.
//
// Every distance on the reference's hot path is
//     subtract(v, c, d); dot(d, d)          (src/linalg.rs:158-165, :12-40)
// i.e. d_e = fl(v_e - c_e), p_e = fl(d_e*d_e) and 16 running sums acc[(e-r)%16]
// (r = m%16, the first r elements seed lanes 0..r-1), added up in lane order at the
// end; for m < 16 a single running sum (dot_naive, :43-53).  Rust never contracts to
// FMA, so every kernel here uses __fsub_rn/__fmul_rn/__fadd_rn (nvcc would otherwise
// fuse).  Callers: reassign_centroids (src/kmeans.rs:279-306), query_partitions
// (src/db/stored.rs:413-424), the ADC table (src/db/stored.rs:556-573).
//
// Two kernels:
//  * exact_tile_kernel  -- m % 16 == 0 and 16-byte aligned rows: a 32x32 (rows x
//    centroids) tile per CTA, K staged through shared memory with cp.async double
//    buffering.  A quad of 4 threads owns one (4 rows x 4 centroids) micro tile; thread
//    tq of the quad keeps lanes 4tq..4tq+3 of the 16 accumulators, so 64 independent
//    FADD chains per thread hide the 4-cycle ALU latency.  The lane sums are chained
//    through the quad with shuffles in lane order.  Bound: fp32 ALU (3 instructions
//    per element pair).
//  * exact_generic_kernel -- any m / alignment, one thread per (row, problem).
#include "common.cuh"

namespace fdb {

namespace {

constexpr uint32_t NONE = 0xFFFFFFFFu;

__device__ __forceinline__ float sq_acc(float acc, float x, float c) {
    float d = __fsub_rn(x, c);
    return __fadd_rn(acc, __fmul_rn(d, d));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

struct TileParams {
    const float *x;
    const float *c;
    const int *active;
    unsigned *flags;
    uint32_t *out_idx;
    float *out_mat;
    size_t n, ldx, col_off, m, nb, k, idx_stride;
};

constexpr int TM = 32, TN = 32, TILE_THREADS = 256;

template <int MODE, int KC>
__global__ void __launch_bounds__(TILE_THREADS, 2) exact_tile_kernel(TileParams p) {
    constexpr int PITCH = KC + 4;
    constexpr int F4_PER_ROW = KC / 4;
    constexpr int LOADS = (TM * F4_PER_ROW + TILE_THREADS - 1) / TILE_THREADS;  // float4 per thread per operand
    __shared__ __align__(16) float Xs[2][TM][PITCH];
    __shared__ __align__(16) float Cs[2][TN][PITCH];

    const int b = blockIdx.y;
    if (p.active && !p.active[b]) return;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int tq = lane & 3, pg = lane >> 2;
    const int rg = pg >> 2, cg = pg & 3;
    const int wr = warp >> 1, wc = warp & 1;
    const int trow = wr * 8 + rg * 4;   // first of this thread's 4 rows in the tile
    const int tcen = wc * 16 + cg * 4;  // first of its 4 centroids in the tile

    const size_t row0 = (size_t)blockIdx.x * TM;
    const float *xb = p.x + p.col_off + (size_t)b * p.m;
    const float *cb = p.c + (size_t)b * p.k * p.m;
    const int nkc = (int)(p.m / KC);
    const int ntn = (int)((p.k + TN - 1) / TN);
    const int total = nkc * ntn;

    auto prefetch = [&](int step, int buf) {
        const int tn = step / nkc, kc = step - tn * nkc;
        const size_t k0 = (size_t)kc * KC;
#pragma unroll
        for (int u = 0; u < LOADS; ++u) {
            const int f = tid + u * TILE_THREADS;
            if (f >= TM * F4_PER_ROW) break;
            const int r = f / F4_PER_ROW, c4 = f - r * F4_PER_ROW;
            size_t grow = row0 + r;
            if (grow >= p.n) grow = p.n - 1;
            cp_async16(&Xs[buf][r][c4 * 4], xb + grow * p.ldx + k0 + c4 * 4);
            size_t gc = (size_t)tn * TN + r;
            if (gc >= p.k) gc = p.k - 1;
            cp_async16(&Cs[buf][r][c4 * 4], cb + gc * p.m + k0 + c4 * 4);
        }
        cp_async_commit();
    };

    float acc[4][4][4];
    float bestd[4];
    uint32_t besti[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        bestd[i] = __int_as_float(0x7f800000);
        besti[i] = NONE;
    }

    prefetch(0, 0);
    for (int step = 0; step < total; ++step) {
        const int buf = step & 1;
        const int tn = step / nkc, kc = step - tn * nkc;
        if (kc == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int l = 0; l < 4; ++l) acc[i][j][l] = 0.0f;
        }
        if (step + 1 < total) {
            prefetch(step + 1, buf ^ 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < KC; kk += 16) {
            float4 xv[4], cv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
                xv[i] = *reinterpret_cast<const float4 *>(&Xs[buf][trow + i][kk + 4 * tq]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                cv[j] = *reinterpret_cast<const float4 *>(&Cs[buf][tcen + j][kk + 4 * tq]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j][0] = sq_acc(acc[i][j][0], xv[i].x, cv[j].x);
                    acc[i][j][1] = sq_acc(acc[i][j][1], xv[i].y, cv[j].y);
                    acc[i][j][2] = sq_acc(acc[i][j][2], xv[i].z, cv[j].z);
                    acc[i][j][3] = sq_acc(acc[i][j][3], xv[i].w, cv[j].w);
                }
        }
        __syncthreads();
        if (kc == nkc - 1) {
            // sum_naive over the 16 lanes (src/linalg.rs:39): chained through the quad
            const int qbase = lane & ~3;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float s4[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float s = 0.0f;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        if (tq == t) {
                            s = __fadd_rn(s, acc[i][j][0]);
                            s = __fadd_rn(s, acc[i][j][1]);
                            s = __fadd_rn(s, acc[i][j][2]);
                            s = __fadd_rn(s, acc[i][j][3]);
                        }
                        s = __shfl_sync(0xffffffffu, s, qbase + t);
                    }
                    s4[j] = s;
                }
                const size_t cen0 = (size_t)tn * TN + tcen;
                if (MODE == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (cen0 + j < p.k && s4[j] < bestd[i]) {
                            bestd[i] = s4[j];
                            besti[i] = (uint32_t)(cen0 + j);
                        }
                } else {
                    const size_t grow = row0 + trow + i;
                    if (tq == 0 && grow < p.n) {
                        float *o = p.out_mat + (grow * p.nb + b) * p.k + cen0;
                        if ((p.k & 3) == 0 && cen0 + 3 < p.k) {
                            *reinterpret_cast<float4 *>(o) = make_float4(s4[0], s4[1], s4[2], s4[3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (cen0 + j < p.k) o[j] = s4[j];
                        }
                    }
                }
            }
        }
    }

    if (MODE == 0) {
        // merge the 8 threads (2 warp columns x 4 centroid groups) that share a row:
        // lexicographic (distance, index) minimum == first strict minimum in index order
        float *red_d = &Xs[0][0][0];
        uint32_t *red_i = reinterpret_cast<uint32_t *>(&Cs[0][0][0]);
        if (tq == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                red_d[(trow + i) * 8 + wc * 4 + cg] = bestd[i];
                red_i[(trow + i) * 8 + wc * 4 + cg] = besti[i];
            }
        }
        __syncthreads();
        if (tid < TM) {
            float bd = __int_as_float(0x7f800000);
            uint32_t bi = NONE;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float d = red_d[tid * 8 + s];
                const uint32_t ix = red_i[tid * 8 + s];
                if (ix != NONE && (d < bd || (d == bd && ix < bi))) {
                    bd = d;
                    bi = ix;
                }
            }
            const size_t grow = row0 + tid;
            if (grow < p.n) {
                if (bi == NONE) atomicOr(p.flags, FLAG_NO_ARGMIN);
                else p.out_idx[(size_t)b * p.idx_stride + grow] = bi;
            }
        }
    }
}

// one thread per (row, problem): any m, any alignment
__device__ __forceinline__ float exact_sqdist_thread(const float *__restrict__ x,
                                                     const float *__restrict__ c, size_t m) {
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

template <int MODE>
__global__ void exact_generic_kernel(TileParams p) {
    const int b = blockIdx.y;
    if (p.active && !p.active[b]) return;
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= p.n) return;
    const float *x = p.x + row * p.ldx + p.col_off + (size_t)b * p.m;
    const float *cb = p.c + (size_t)b * p.k * p.m;
    float bd = __int_as_float(0x7f800000);
    uint32_t bi = NONE;
    for (size_t j = 0; j < p.k; ++j) {
        const float d = exact_sqdist_thread(x, cb + j * p.m, p.m);
        if (MODE == 0) {
            if (d < bd) {
                bd = d;
                bi = (uint32_t)j;
            }
        } else {
            p.out_mat[(row * p.nb + b) * p.k + j] = d;
        }
    }
    if (MODE == 0) {
        if (bi == NONE) atomicOr(p.flags, FLAG_NO_ARGMIN);
        else p.out_idx[(size_t)b * p.idx_stride + row] = bi;
    }
}

// few rows (the queries the filter path hands back): one quad of lanes per (row, problem,
// centroid) item instead of a 32-row tile per CTA, so that even a single row spreads over the
// whole GPU.  m % 16 == 0; lane tq of the quad owns accumulators 4tq..4tq+3.
template <bool PIPE>
__global__ void __launch_bounds__(256) exact_pairs_kernel(TileParams p) {
    const size_t item = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int lane = threadIdx.x & 31, tq = lane & 3, qbase = lane & ~3;
    const size_t total = p.n * p.nb * p.k;
    const bool act = item < total;
    const size_t it = act ? item : 0;
    const size_t row = it / (p.nb * p.k), rem = it - row * p.nb * p.k;
    const size_t b = rem / p.k, j = rem - b * p.k;
    const float *x = p.x + row * p.ldx + p.col_off + b * p.m;
    const float *c = p.c + (b * p.k + j) * p.m;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    if (PIPE) {
        // long rows: software pipeline, the next PF float4 pairs are in flight while the current
        // ones are summed (the sums are sequential per accumulator, the loads are not)
        constexpr int PF = 8;
        float4 xb[PF], cb[PF];
        const size_t iters = p.m / 16;
#pragma unroll
        for (int u = 0; u < PF; ++u)
            if ((size_t)u < iters) {
                xb[u] = __ldg(reinterpret_cast<const float4 *>(x + 4 * tq + 16 * u));
                cb[u] = __ldg(reinterpret_cast<const float4 *>(c + 4 * tq + 16 * u));
            }
        for (size_t i0 = 0; i0 < iters; i0 += PF) {
            float4 xn[PF], cn[PF];
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (i0 + PF + u < iters) {
                    xn[u] = __ldg(reinterpret_cast<const float4 *>(x + 4 * tq + 16 * (i0 + PF + u)));
                    cn[u] = __ldg(reinterpret_cast<const float4 *>(c + 4 * tq + 16 * (i0 + PF + u)));
                }
#pragma unroll
            for (int u = 0; u < PF; ++u)
                if (i0 + u < iters) {
                    a0 = sq_acc(a0, xb[u].x, cb[u].x);
                    a1 = sq_acc(a1, xb[u].y, cb[u].y);
                    a2 = sq_acc(a2, xb[u].z, cb[u].z);
                    a3 = sq_acc(a3, xb[u].w, cb[u].w);
                }
#pragma unroll
            for (int u = 0; u < PF; ++u) {
                xb[u] = xn[u];
                cb[u] = cn[u];
            }
        }
    } else {
        for (size_t e = 4 * tq; e < p.m; e += 16) {
            const float4 xv = *reinterpret_cast<const float4 *>(x + e);
            const float4 cv = *reinterpret_cast<const float4 *>(c + e);
            a0 = sq_acc(a0, xv.x, cv.x);
            a1 = sq_acc(a1, xv.y, cv.y);
            a2 = sq_acc(a2, xv.z, cv.z);
            a3 = sq_acc(a3, xv.w, cv.w);
        }
    }
    float sum = 0.0f;  // sum_naive over the 16 lanes (src/linalg.rs:39), chained through the quad
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (tq == t) {
            sum = __fadd_rn(sum, a0);
            sum = __fadd_rn(sum, a1);
            sum = __fadd_rn(sum, a2);
            sum = __fadd_rn(sum, a3);
        }
        sum = __shfl_sync(0xffffffffu, sum, qbase + t);
    }
    if (act && tq == 0) p.out_mat[(row * p.nb + b) * p.k + j] = sum;
}

template <int MODE>
int launch(fdb_ctx *ctx, const DistProblem &q, uint32_t *d_idx, size_t idx_stride, float *d_out) {
    if (q.n == 0 || q.k == 0 || q.nb == 0) return FDB_OK;
    TileParams p;
    p.x = q.x;
    p.c = q.c;
    p.active = q.active;
    p.flags = ctx->d_flags;
    p.out_idx = d_idx;
    p.out_mat = d_out;
    p.n = q.n;
    p.ldx = q.ldx;
    p.col_off = q.col_off;
    p.m = q.m;
    p.nb = q.nb;
    p.k = q.k;
    p.idx_stride = idx_stride;
    if (q.nb > 65535) {
        set_error("too many side-by-side problems: %zu", q.nb);
        return FDB_ERR_UNSUPPORTED;
    }
    const bool aligned = (q.m % 16 == 0) && (q.ldx % 4 == 0) && (q.col_off % 4 == 0) &&
                         ((uintptr_t)q.x % 16 == 0) && ((uintptr_t)q.c % 16 == 0);
    const size_t tile_ctas = ((q.n + TM - 1) / TM) * q.nb, items = q.n * q.nb * q.k;
    if (MODE == 1 && aligned && !q.active && tile_ctas * 2 < (size_t)ctx->sm_count && items <= (1u << 22)) {
        if (q.m >= 512) exact_pairs_kernel<true><<<(unsigned)((items * 4 + 255) / 256), 256, 0, ctx->stream>>>(p);
        else exact_pairs_kernel<false><<<(unsigned)((items * 4 + 255) / 256), 256, 0, ctx->stream>>>(p);
    } else if (aligned) {
        dim3 grid((unsigned)((q.n + TM - 1) / TM), (unsigned)q.nb);
        if (q.m % 64 == 0) exact_tile_kernel<MODE, 64><<<grid, TILE_THREADS, 0, ctx->stream>>>(p);
        else if (q.m % 32 == 0) exact_tile_kernel<MODE, 32><<<grid, TILE_THREADS, 0, ctx->stream>>>(p);
        else exact_tile_kernel<MODE, 16><<<grid, TILE_THREADS, 0, ctx->stream>>>(p);
    } else {
        dim3 grid((unsigned)((q.n + 127) / 128), (unsigned)q.nb);
        exact_generic_kernel<MODE><<<grid, 128, 0, ctx->stream>>>(p);
    }
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

}  // namespace

int launch_exact_argmin(fdb_ctx *ctx, const DistProblem &p, uint32_t *d_idx, size_t idx_stride) {
    return launch<0>(ctx, p, d_idx, idx_stride, nullptr);
}
int launch_exact_matrix(fdb_ctx *ctx, const DistProblem &p, float *d_out) {
    return launch<1>(ctx, p, nullptr, 0, d_out);
}

}  // namespace fdb
