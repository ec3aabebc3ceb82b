// Lloyd assignment on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), sm_100a.
// Replaces the O(n*k*m) loop of reassign_centroids (src/kmeans.rs:279-306) without changing
// its result:
//
//   1. GEMM filter.  s_ij = x_i . c_j - |c_j|^2/2  (argmin_j |x_i-c_j|^2 == argmax_j s_ij) is
//      evaluated on the tensor pipe with an fp32-accurate 3-term split: every f32 value v is
//      stored as two bf16 pieces v1 = bf16(v), v2 = bf16(v - v1) (|v - v1 - v2| <= 2^-18 |v|)
//      and  x.c ~= x1.c1 + x1.c2 + x2.c1  accumulates in fp32 in TMEM (kind::f16 UMMA,
//      M=128, N<=256, K=16; operands staged by TMA into 128B-swizzled shared memory).
//      The rows' pieces are written once per k-means problem (the rows do not change
//      between Lloyd rounds) and cost the same 4 bytes/element as the f32 rows.
//      Both operands are first shifted by the column mean mu of the rows (x' = fl(x - mu),
//      c' = fl(c - mu)): distances are shift invariant, while the GEMM error bound scales with
//      |x'||c'| instead of |x||c| (20x smaller on non-centred data such as uniform [0,1)).
//   2. Band.  With E_i >= |s~_ij - s_ij| (split + accumulation error, proportional to
//      |x_i| max_j|c_j|) and eta >= the relative rounding error of the reference's own f32
//      evaluation, the reference's argmin j* satisfies s~_ij* >= max_j s~_ij - band_i,
//      band_i = 2 (2 E_i + 1.01 eta d~_min).  The epilogue collects every j inside the band.
//   3. Exact re-check.  Rows with one candidate are final; the others (a few %) are
//      evaluated for their candidates only, in the reference's summation order and with its
//      tie rule (lowest index), by recheck_kernel.  Rows with more than CAP candidates fall
//      back to all k centroids.  Result: bit-identical indices at tensor-core speed.
//
// Kernel anatomy (persistent, one CTA per SM, 384 threads):
//   warp 0   TMA producer   cp.async.bulk.tensor 2D, SWIZZLE_128B, mbarrier complete_tx
//   warp 1   MMA issuer     one elected lane, tcgen05.mma.cta_group::1.kind::f16
//   warp 2   TMEM allocator 512 columns = two accumulator stages of up to 256 columns
//   warps 4-11 epilogue     tcgen05.ld 32x32b.x32, two passes (row max, band collection)
#include "kmeans.cuh"
#include "tc_gemm.cuh"

#include <cstdlib>

namespace fdb {

namespace {

constexpr int BM = 128;          // rows per tile (UMMA M)
constexpr int BK = 64;           // bf16 elements per K chunk = one 128-byte swizzle row
constexpr int CAP = 8;           // candidates kept per row before falling back to all k
constexpr unsigned ALL_CAP = 1u << 18;   // rows decided over all k that get a CTA of their own (the others: a warp)
constexpr int TC_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile, 128-byte rows, SWIZZLE_128B: 8-row groups are 1024 bytes apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO [16,30), SBO [32,46), version=1 [46,48),
//  layout_type=2 (SWIZZLE_128B) [61,64))
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, int bk) {
    // bk = 64: 128-byte rows, SWIZZLE_128B (layout 2), 8-row groups 1024 bytes apart;
    // bk = 16:  32-byte rows, SWIZZLE_32B  (layout 6), 8-row groups  256 bytes apart
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                                   // LBO (unused for swizzled K-major) = 1
    d |= (uint64_t)((bk == 64 ? 1024 : 256) >> 4) << 32;      // SBO
    d |= (uint64_t)1 << 46;                                   // descriptor version (Blackwell)
    d |= (uint64_t)(bk == 64 ? 2 : 6) << 61;                  // layout type
    return d;
}
// issue without waiting: the registers are valid only after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, float (&v)[32]) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t va_bits(const float (&v)[32], int i) { return __float_as_uint(v[i]); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- operand preparation -----------------------------------------------------------------
// rows: f32 -> two bf16 pieces (full row width, every division at once) + |x|^2 per (problem,row)
// (the pieces of problem b start at column out_off + b * out_m of rows out_ld wide: the layout of the
//  rows themselves, or a compact one padded with zeros when m is not a multiple of 16)
__global__ void __launch_bounds__(256) split_rows_kernel(const float *x, size_t n, size_t ldx,
                                                         size_t col_off, size_t m, size_t nb,
                                                         const float *mu, __nv_bfloat16 *x1,
                                                         __nv_bfloat16 *x2, float *xn2, size_t out_ld,
                                                         size_t out_off, size_t out_m, size_t out_bstride = 0) {
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n * nb) return;
    const size_t row = warp / nb, b = warp - row * nb;
    const float *xr = x + row * ldx + col_off + b * m;
    const float *mr = mu + col_off + b * m;
    __nv_bfloat16 *o1 = x1 + b * out_bstride + row * out_ld + out_off + b * out_m;
    __nv_bfloat16 *o2 = x2 + b * out_bstride + row * out_ld + out_off + b * out_m;
    double acc = 0.0;
    // four elements per lane (one 16-byte load, two 8-byte stores) when everything is aligned
    const bool vec = m % 4 == 0 && ((uintptr_t)xr % 16 == 0) && ((uintptr_t)mr % 16 == 0) && ((uintptr_t)o1 % 8 == 0) &&
                     ((uintptr_t)o2 % 8 == 0);
    if (vec) {
        for (size_t e = 4 * (size_t)lane; e < m; e += 128) {
            const float4 xv = *reinterpret_cast<const float4 *>(xr + e);
            const float4 mv = __ldg(reinterpret_cast<const float4 *>(mr + e));
            const float v[4] = {__fsub_rn(xv.x, mv.x), __fsub_rn(xv.y, mv.y), __fsub_rn(xv.z, mv.z), __fsub_rn(xv.w, mv.w)};
            __align__(8) __nv_bfloat16 h[4];
            __align__(8) __nv_bfloat16 l[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                h[i] = __float2bfloat16_rn(v[i]);
                l[i] = __float2bfloat16_rn(__fsub_rn(v[i], __bfloat162float(h[i])));
                acc += (double)v[i] * (double)v[i];
            }
            *reinterpret_cast<uint2 *>(o1 + e) = *reinterpret_cast<const uint2 *>(h);
            *reinterpret_cast<uint2 *>(o2 + e) = *reinterpret_cast<const uint2 *>(l);
        }
    } else {
        for (size_t e = lane; e < m; e += 32) {
            const float v = __fsub_rn(xr[e], mr[e]);
            const __nv_bfloat16 h = __float2bfloat16_rn(v);
            const float r = __fsub_rn(v, __bfloat162float(h));
            o1[e] = h;
            o2[e] = __float2bfloat16_rn(r);
            acc += (double)v * (double)v;
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) xn2[b * n + row] = __double2float_ru(acc);
}

// column means of the rows, deterministic two-level reduction (slices of rows, then slices)
constexpr int MEAN_SLICES = 64;
__global__ void __launch_bounds__(128) col_partial_kernel(const float *x, size_t n, size_t ldx, size_t c0,
                                                          size_t ncols, double *partial) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    const size_t per = (n + MEAN_SLICES - 1) / MEAN_SLICES;
    const size_t lo = blockIdx.y * per, hi = lo + per < n ? lo + per : n;
    double acc = 0.0;
    for (size_t r = lo; r < hi; ++r) acc += (double)x[r * ldx + c0 + col];
    partial[(size_t)blockIdx.y * ncols + col] = acc;
}
__global__ void col_mean_kernel(const double *partial, size_t n, size_t c0, size_t ncols, float *mu) {
    const size_t col = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= ncols) return;
    double acc = 0.0;
    for (int sl = 0; sl < MEAN_SLICES; ++sl) acc += partial[(size_t)sl * ncols + col];
    const float v = (float)(acc / (double)n);
    mu[c0 + col] = (v == v && fabsf(v) < 3.0e38f) ? v : 0.0f;
}

// centroids: f32 -> two bf16 pieces [nb*k][m], h_j = |c_j|^2/2, cmax_b = max_j |c_j| (rounded up)
__global__ void __launch_bounds__(128) prep_centroids_kernel(const float *c, size_t k, size_t m, size_t np,
                                                             const float *mu, size_t col_off, size_t mu_stride,
                                                             size_t ktotal, size_t out_m,
                                                             __nv_bfloat16 *c1, __nv_bfloat16 *c2,
                                                             float *h, unsigned *cmax2_bits,
                                                             const int *active) {
    const size_t b = blockIdx.y, j = blockIdx.x;
    if (active && !active[b]) return;
    if (j >= k || b * k + j >= ktotal) {  // padded columns never win: s = acc - inf
        if (threadIdx.x == 0) h[b * np + j] = __int_as_float(0x7f800000);
        return;
    }
    const float *cr = c + (b * k + j) * m;
    const float *mr = mu + col_off + b * mu_stride;
    double acc = 0.0;
    for (size_t e = threadIdx.x; e < m; e += blockDim.x) {
        const float v = __fsub_rn(cr[e], mr[e]);
        const __nv_bfloat16 hh = __float2bfloat16_rn(v);
        c1[(b * k + j) * out_m + e] = hh;
        c2[(b * k + j) * out_m + e] = __float2bfloat16_rn(__fsub_rn(v, __bfloat162float(hh)));
        acc += (double)v * (double)v;
    }
    __shared__ double red[128];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int off = 64; off >= 1; off >>= 1) {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        h[b * np + j] = (float)(0.5 * red[0]);
        atomicMax(&cmax2_bits[b], __float_as_uint(__double2float_ru(red[0])));  // >= 0: bit order == value order
    }
}

// shared memory plan: when both centroid pieces of a problem (all K chunks) fit beside two stages of row pieces they
// stay resident -- the rows are the only operand that streams (the PQ shape: 128 KB of centroids, read once per
// problem instead of once per 128-row tile)
static inline void tc_smem_plan(size_t np, size_t bk, size_t m, int *stages, int *bres, size_t *smem) {
    const size_t a2 = 2 * (size_t)BM * bk * 2, b2 = 2 * np * bk * 2, kchunks = m / bk;
    const size_t budget = 200 * 1024;
    const bool off = getenv("FDB_TC_NO_BRES") != nullptr;
    if (!off && kchunks * b2 + 2 * a2 <= budget) {
        *bres = 1;
        *stages = (int)std::min<size_t>(4, (budget - kchunks * b2) / a2);
        *smem = kchunks * b2 + (size_t)*stages * a2 + 1024;
    } else {
        *bres = 0;
        *stages = (int)std::min<size_t>(4, budget / (a2 + b2));
        *smem = (size_t)*stages * (a2 + b2) + 1024;
    }
}

struct TcParams {
    size_t n, m, nb, k, col_off;
    size_t xcol_stride;         // operand columns of problem b start at col_off + b * xcol_stride
    size_t xrow_stride;         // ... and its rows at b * xrow_stride (problem-major pieces: every problem's rows contiguous)
    size_t crow_stride;         // centroid rows of problem b start at b * crow_stride
    int mode;                   // 0: assignment epilogue, 1: raw scores out = alpha * acc - (sub_h ? h : 0),
                                // 2: the three largest scores of every (row, column tile) for tc_combine_kernel
    float *tile_v;              // mode 2: [n][nb][4] scores v1 >= v2 >= v3 of the tile (problem b = column tile)
    uint16_t *tile_i;           //         [n][nb][2] columns of v1, v2 inside the tile
    float *out;                 // mode 1: out[row * out_row_stride + b * out_b_stride + col]
    size_t out_row_stride, out_b_stride;
    float alpha;
    int sub_h;
    int np;                     // padded N (multiple of 64, <= 256)
    int bk;                     // bf16 elements per K chunk: 64 (m % 64 == 0) or 16
    int row_tiles;              // ceil(n / 128)
    int stages;
    int bres;                   // 1: the centroid pieces of a problem stay in shared memory (all K chunks), only the rows stream
    const float *h;             // [nb][np]
    const float *xn2;           // [nb][n]
    const unsigned *cmax2_bits; // [nb]
    const int *active;
    float gamma1, eta;
    uint32_t *indices;          // [nb][n]
    unsigned *work_count;       // rows that need the exact re-check
    uint32_t *work_rows;        // [cap] b * n + row
    uint16_t *work_cand;        // [cap][CAP]
    uint8_t *work_cnt;          // [cap] number of candidates (CAP+1 = overflow)
    unsigned work_cap;
    unsigned *stats;            // [2]: (unused), rows that overflowed CAP
    unsigned long long *dbg;    // FDB_TC_DEBUG & 16: cycle counters of the assignment epilogue [8]
    int debug;                  // FDB_TC_DEBUG (timing experiments, WRONG results): 1 = the producer skips the loads, 2 = the
                                // assignment epilogue only releases the accumulator, 4 = ... only loads it, 8 = ... does its
                                // arithmetic on registers and writes nothing, 16 = cycle counters (dbg)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
tc_assign_kernel(const __grid_constant__ CUtensorMap map_x1, const __grid_constant__ CUtensorMap map_x2,
                 const __grid_constant__ CUtensorMap map_c1, const __grid_constant__ CUtensorMap map_c2,
                 TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char *smem = reinterpret_cast<unsigned char *>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[4], empty_bar[4], tmem_full[2], tmem_empty[2], b_full, b_free;
    __shared__ uint32_t tmem_base_s;
    __shared__ float rowmax_s[2][BM], rowsec_s[2][BM];
    __shared__ uint16_t rowarg_s[2][BM];
    __shared__ __align__(16) float h_s[2][256];          // mode 0: one copy per epilogue group
    __shared__ uint16_t cand_s[2][BM][CAP];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NP = p.np;
    const int BKr = p.bk;                             // 64, or 16 when m is not a multiple of 64
    const uint32_t a_bytes = BM * BKr * 2;            // 16 KB per piece (bk = 64)
    const uint32_t b_bytes = (uint32_t)NP * BKr * 2;  // NP * 128 B per piece
    const int S = p.stages;
    const int kchunks = (int)(p.m / BKr);
    // B resident: [kchunks][c1 | c2] of the current problem at the front, then the stages hold the two row pieces only
    const bool bres = p.bres != 0;
    const uint32_t stage_bytes = bres ? 2 * a_bytes : 2 * a_bytes + 2 * b_bytes;
    unsigned char *bsm = smem;
    if (bres) smem += (size_t)kchunks * 2 * b_bytes;
    const int total_tiles = (int)p.nb * p.row_tiles;
    // contiguous tile ranges per CTA (a CTA mostly stays inside one problem: its centroids stay hot in L2)
    const int per_cta = (total_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = (int)blockIdx.x * per_cta;
    const int t_end = min(total_tiles, t_begin + per_cta);

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tmem_full[s], 1);
            mbar_init(&tmem_empty[s], p.mode == 0 ? 4 : EPI_THREADS / 32);   // mode 0: a stage belongs to one group of 4 warps
        }
        mbar_init(&b_full, 1);
        mbar_init(&b_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                     "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0, cur_b = -1;
            uint32_t phase = 0, bphase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int b = t / p.row_tiles, rt = t - b * p.row_tiles;
                if (p.active && !p.active[p.mode == 2 ? 0 : b]) continue;
                const int row0 = rt * BM + (int)((size_t)b * p.xrow_stride);
                const int kcol0 = (int)(p.col_off + (size_t)b * p.xcol_stride);
                const int crow0 = (int)((size_t)b * p.crow_stride);
                if (bres && b != cur_b) {
                    // the MMAs of the previous problem have retired (the issuer commits b_free when it moves on)
                    mbar_wait(&b_free, bphase ^ 1);
                    bphase ^= 1;
                    mbar_expect_tx(&b_full, (uint32_t)kchunks * 2 * b_bytes);
                    for (int kc = 0; kc < kchunks; ++kc) {
                        tma_load_2d(bsm + (size_t)(2 * kc) * b_bytes, &map_c1, kc * BKr, crow0, &b_full);
                        tma_load_2d(bsm + (size_t)(2 * kc + 1) * b_bytes, &map_c2, kc * BKr, crow0, &b_full);
                    }
                    cur_b = b;
                }
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    unsigned char *st = smem + (size_t)stage * stage_bytes;
                    if ((p.debug & 1) && t >= t_begin + 2) {
                        mbar_arrive(&full_bar[stage]);
                        if (++stage == S) {
                            stage = 0;
                            phase ^= 1;
                        }
                        continue;
                    }
                    mbar_expect_tx(&full_bar[stage], stage_bytes);
                    tma_load_2d(st, &map_x1, kcol0 + kc * BKr, row0, &full_bar[stage]);
                    tma_load_2d(st + a_bytes, &map_x2, kcol0 + kc * BKr, row0, &full_bar[stage]);
                    if (!bres) {
                        tma_load_2d(st + 2 * a_bytes, &map_c1, kc * BKr, crow0, &full_bar[stage]);
                        tma_load_2d(st + 2 * a_bytes + b_bytes, &map_c2, kc * BKr, crow0, &full_bar[stage]);
                    }
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            // kind::f16 instruction descriptor: D=f32 (bit 4), A=B=bf16 (bits 7, 10), K-major A and B,
            // N>>3 at [17,23), M>>4 at [24,29)   (cute::UMMA::InstrDescriptor)
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NP >> 3) << 17) |
                                   ((uint32_t)(BM >> 4) << 24);
            int stage = 0, as = 0, cur_b = -1;
            uint32_t phase = 0, aphase = 0, bphase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                const int b = t / p.row_tiles;
                if (p.active && !p.active[p.mode == 2 ? 0 : b]) continue;
                if (bres && b != cur_b) {
                    if (cur_b >= 0) umma_commit(&b_free);   // arrives when the previous problem's MMAs have retired
                    mbar_wait(&b_full, bphase);
                    bphase ^= 1;
                    tc_fence_after();
                    cur_b = b;
                }
                mbar_wait(&tmem_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(as * NP);
                for (int kc = 0; kc < kchunks; ++kc) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
                    const uint64_t a1 = make_smem_desc(sa, BKr), a2 = make_smem_desc(sa + a_bytes, BKr);
                    const uint32_t sb = bres ? smem_u32(bsm + (size_t)(2 * kc) * b_bytes) : sa + 2 * a_bytes;
                    const uint64_t b1 = make_smem_desc(sb, BKr);
                    const uint64_t b2 = make_smem_desc(sb + b_bytes, BKr);
                    for (int ks = 0; ks < BKr / 16; ++ks) {
                        const uint64_t off = (uint64_t)(ks * 32 >> 4);  // 16 bf16 = 32 bytes along K
                        umma_bf16(d_tmem, a2 + off, b1 + off, idesc, (kc | ks) != 0);  // small terms first
                        umma_bf16(d_tmem, a1 + off, b2 + off, idesc, 1);
                        umma_bf16(d_tmem, a1 + off, b1 + off, idesc, 1);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem stage when these MMAs retire
                    if (++stage == S) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tmem_full[as]);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
            }
        }
    } else if (warp >= EPI_WARP0 && p.mode == 0) {
        // ===== assignment epilogue: two groups of four warps, group g owns accumulator stage g (every other tile);
        //       one thread per row over all NP columns -- the warps share nothing, so a tile costs no CTA barrier,
        //       a warp that needs the (rare) second pass delays nobody, and while one group reduces tile t the other
        //       reduces tile t + 1 =====
        const int ew = warp - EPI_WARP0;
        const int q = ew & 3, g = ew >> 2;
        const int row = 32 * q + lane;
        const int gt = threadIdx.x - (EPI_WARP0 + 4 * g) * 32;   // 0..127 inside the group
        float *hg = h_s[g];
        const float NEG_INF = -__int_as_float(0x7f800000);
        int h_loaded = -1, seq = 0;
        uint32_t aphase = 0;
        unsigned tk_end = p.dbg ? (unsigned)clock() : 0u;
        int b = t_begin / p.row_tiles, rt = t_begin - b * p.row_tiles - 1;
        bool b_on = !p.active || (t_begin < t_end && p.active[b] != 0);
        for (int t = t_begin; t < t_end; ++t) {
            if (++rt == p.row_tiles) {
                rt = 0;
                ++b;
                b_on = !p.active || p.active[b] != 0;
            }
            if (!b_on) continue;
            if ((seq++ & 1) != g) continue;
            const size_t grow = (size_t)rt * BM + row;
            const bool valid = grow < p.n;
            // h_j = |c'_j|^2/2 of this problem in the group's shared copy (reloaded when the problem changes)
            if (b != h_loaded) {
                asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");  // nobody in the group still reads the old values
                for (int j = gt; j < NP; j += 128) hg[j] = p.h ? p.h[(size_t)b * NP + j] : 0.0f;
                h_loaded = b;
                asm volatile("bar.sync %0, 128;" ::"r"(1 + g) : "memory");
            }
            // the row's band inputs travel while the tile is still being multiplied
            float xn2 = valid ? __ldg(p.xn2 + (size_t)b * p.n + grow) : 0.0f;
            float cmax2 = __uint_as_float(__ldg(p.cmax2_bits + b));
            const unsigned tk0 = p.dbg ? (unsigned)clock() : 0u;
            mbar_wait(&tmem_full[g], aphase);
            aphase ^= 1;
            tc_fence_after();
            const unsigned tk1 = p.dbg ? (unsigned)clock() : 0u;
            const uint32_t taddr = tmem_base + (uint32_t)(g * NP) + ((uint32_t)(32 * q) << 16);
            if (p.debug & 2) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[g]);
                continue;
            }
            if (p.debug & 4) {   // the TMEM loads alone
                float va[32];
                uint32_t x = 0;
                for (int c0 = 0; c0 < NP; c0 += 32) {
                    tmem_ld32_issue(taddr + c0, va);
                    tmem_ld_wait();
                    x ^= va_bits(va, 0) ^ va_bits(va, 31);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[g]);
                if (x == 0x12345678u) p.indices[0] = x;
                continue;
            }
            // pass 1 (software pipelined TMEM loads): the two largest s = acc - h and the argmax, as two independent
            // chains (even / odd columns).  Measured with FDB_TC_DEBUG=16: ~4.4k cycles per 32 rows x 256 columns, of
            // which ~1.3k are the shared-memory reads of h (they compete with the tensor pipe's operand fetches); four
            // chains, or a block-of-four top-2 network with fewer ALU-pipe instructions, measured the same.
            float m1a = NEG_INF, m2a = NEG_INF, m1b = NEG_INF, m2b = NEG_INF;
            int ia = 0, ib = 1;
            {
                float va[32], vb[32];
                const bool ld = !(p.debug & 8);
                if (!ld) {
#pragma unroll
                    for (int i = 0; i < 32; ++i) va[i] = vb[i] = (float)(lane + i);
                }
                if (ld) tmem_ld32_issue(taddr, va);
                for (int c0 = 0; c0 < NP; c0 += 64) {   // NP is a multiple of 64
                    tmem_ld_wait();
                    if (ld) tmem_ld32_issue(taddr + c0 + 32, vb);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float s0 = __uint_as_float(va_bits(va, i)) - hg[c0 + i];
                        const float s1 = __uint_as_float(va_bits(va, i + 1)) - hg[c0 + i + 1];
                        m2a = fmaxf(m2a, fminf(s0, m1a));
                        ia = s0 > m1a ? c0 + i : ia;
                        m1a = fmaxf(m1a, s0);
                        m2b = fmaxf(m2b, fminf(s1, m1b));
                        ib = s1 > m1b ? c0 + i + 1 : ib;
                        m1b = fmaxf(m1b, s1);
                    }
                    tmem_ld_wait();
                    if (ld && c0 + 64 < NP) tmem_ld32_issue(taddr + c0 + 64, va);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const float s0 = __uint_as_float(va_bits(vb, i)) - hg[c0 + 32 + i];
                        const float s1 = __uint_as_float(va_bits(vb, i + 1)) - hg[c0 + 32 + i + 1];
                        m2a = fmaxf(m2a, fminf(s0, m1a));
                        ia = s0 > m1a ? c0 + 32 + i : ia;
                        m1a = fmaxf(m1a, s0);
                        m2b = fmaxf(m2b, fminf(s1, m1b));
                        ib = s1 > m1b ? c0 + 32 + i + 1 : ib;
                        m1b = fmaxf(m1b, s1);
                    }
                }
            }
            const float smax = fmaxf(m1a, m1b);
            const float second = fmaxf(fminf(m1a, m1b), fmaxf(m2a, m2b));
            const int imax = m1a > m1b ? ia : (m1b > m1a ? ib : min(ia, ib));
            const unsigned tk2 = p.dbg ? (unsigned)clock() : 0u;
            // band (see the header of this file).  (The empty asm pins the first use of the two loads here, after
            // pass 1: otherwise the arithmetic that depends on them alone is hoisted above the wait for the
            // accumulator and the warp sits on the load latency before it even starts waiting.)
            asm volatile("" : "+f"(xn2), "+f"(cmax2));
            const float E = p.gamma1 * sqrtf(xn2 * cmax2) * 1.0001f + 1.2e-7f * (0.5f * cmax2);
            const float dmin = fmaxf(0.0f, xn2 - 2.0f * smax + 2.0f * E);
            // rounding of the shift: |d(x',c') - d(x,c)| <= 2 sqrt(d) 2^-24 (|x'| + |c'|)
            const float shift = 1.3e-7f * sqrtf(dmin) * (sqrtf(xn2) + sqrtf(cmax2));
            const float band = 2.0f * (2.0f * E + 1.01f * p.eta * dmin + shift);
            const float thresh = smax - band;
            // a single column inside the band <=> the runner-up is below the threshold (the
            // negation also catches NaN scores, which must go to the exact kernel)
            bool single = valid && (second < thresh) && (smax > NEG_INF);
            bool need2 = valid && !single;
            if (p.debug & 8) single = need2 = (smax == 1.2345f);   // (timing experiment: nothing is written)
            // pass 2 (rare, this warp only): the columns inside the band as one bit mask per 32 columns (three
            // instructions per column, no branch), then the positions of the few set bits in ascending order
            unsigned c = 0;
            if (__any_sync(0xffffffffu, need2)) {
                float v[32];
                uint16_t *cs = cand_s[g][row];
                tmem_ld32_issue(taddr, v);
                for (int c0 = 0; c0 < NP; c0 += 32) {
                    tmem_ld_wait();
                    uint32_t hits = 0u;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float sv = __uint_as_float(va_bits(v, i)) - hg[c0 + i];
                        hits |= sv >= thresh ? (1u << i) : 0u;
                    }
                    if (c0 + 32 < NP) tmem_ld32_issue(taddr + c0 + 32, v);
                    if (!need2) hits = 0u;
                    while (hits) {   // (a handful of bits on a handful of lanes)
                        const int i = __ffs(hits) - 1;
                        hits &= hits - 1u;
                        if (c < CAP) cs[c] = (uint16_t)(c0 + i);
                        ++c;
                    }
                }
            }
            // the accumulator stage is free again
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[g]);
            const unsigned tk3 = p.dbg ? (unsigned)clock() : 0u;
            if (p.dbg && lane == 0) {
                atomicAdd(p.dbg + 5, (unsigned long long)(tk0 - tk_end));   // end of the previous tile -> ready to wait
                atomicAdd(p.dbg + 0, (unsigned long long)(tk1 - tk0));   // waiting for the accumulator
                atomicAdd(p.dbg + 1, (unsigned long long)(tk2 - tk1));   // pass 1
                atomicAdd(p.dbg + 2, (unsigned long long)(tk3 - tk2));   // band + pass 2
                atomicAdd(p.dbg + 3, 1ull);                               // warp-tiles
                if (c) atomicAdd(p.dbg + 4, 1ull);
            }
            // rows with exactly one column inside the band are final; the others go to the
            // re-check list (c < 2: non-finite scores, c > CAP: too many -> all k centroids)
            const unsigned bal = __ballot_sync(0xffffffffu, need2);
            if (single) p.indices[(size_t)b * p.n + grow] = (uint32_t)imax;
            if (bal) {
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(p.work_count, (unsigned)__popc(bal));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (need2) {
                    const unsigned slot = base + __popc(bal & ((1u << lane) - 1u));
                    if (slot < p.work_cap) {
                        p.work_rows[slot] = (uint32_t)((size_t)b * p.n + grow);
                        const unsigned cc = (c < 2 || c > CAP) ? CAP + 1 : c;
                        p.work_cnt[slot] = (uint8_t)cc;
                        uint16_t cv[CAP];
#pragma unroll
                        for (unsigned u = 0; u < CAP; ++u) cv[u] = u < c ? cand_s[g][row][u] : (uint16_t)0;
                        uint4 pk;
                        pk.x = (uint32_t)cv[0] | ((uint32_t)cv[1] << 16);
                        pk.y = (uint32_t)cv[2] | ((uint32_t)cv[3] << 16);
                        pk.z = (uint32_t)cv[4] | ((uint32_t)cv[5] << 16);
                        pk.w = (uint32_t)cv[6] | ((uint32_t)cv[7] << 16);
                        *reinterpret_cast<uint4 *>(p.work_cand + (size_t)slot * CAP) = pk;
                        if (cc > CAP) atomicAdd(&p.stats[1], 1u);
                    }
                }
            }
            if (p.dbg) {
                tk_end = (unsigned)clock();
                if (lane == 0) atomicAdd(p.dbg + 6, (unsigned long long)(tk_end - tk3));   // stores and appends
            }
        }
    } else if (warp >= EPI_WARP0) {
        // ===== epilogue of the raw / tile modes: 2 threads per row (column halves) =====
        const int ew = warp - EPI_WARP0;
        const int q = ew & 3, hf = ew >> 2;
        const int row = 32 * q + lane;
        const int half = NP >> 1;
        const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..255
        int as = 0, h_loaded = -1;
        uint32_t aphase = 0;
        for (int t = t_begin; t < t_end; ++t) {
            const int b = t / p.row_tiles, rt = t - b * p.row_tiles;
            if (p.active && !p.active[p.mode == 2 ? 0 : b]) continue;
            const size_t grow = (size_t)rt * BM + row;
            const bool valid = grow < p.n;
            // h_j = |c'_j|^2/2 of this problem in shared memory (reloaded when the problem changes)
            if (b != h_loaded) {
                asm volatile("bar.sync 1, 256;" ::: "memory");  // nobody still reads the old values
                for (int j = et; j < NP; j += EPI_THREADS) h_s[0][j] = p.h ? p.h[(size_t)b * NP + j] : 0.0f;
                h_loaded = b;
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            const float *hb = h_s[0] + hf * half;
            mbar_wait(&tmem_full[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)(as * NP + hf * half) + ((uint32_t)(32 * q) << 16);
            if (p.mode == 1) {
                // raw scores: every thread streams its half row out of TMEM, 32 columns at a time
                float *orow = p.out + grow * p.out_row_stride + (size_t)b * p.out_b_stride + hf * half;
                float va[32];
                for (int c0 = 0; c0 < half; c0 += 32) {
                    tmem_ld32_issue(taddr + c0, va);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 32; i += 4) {
                            float4 o;
                            o.x = p.alpha * va[i] - (p.sub_h ? hb[c0 + i] : 0.0f);
                            o.y = p.alpha * va[i + 1] - (p.sub_h ? hb[c0 + i + 1] : 0.0f);
                            o.z = p.alpha * va[i + 2] - (p.sub_h ? hb[c0 + i + 2] : 0.0f);
                            o.w = p.alpha * va[i + 3] - (p.sub_h ? hb[c0 + i + 3] : 0.0f);
                            *reinterpret_cast<float4 *>(orow + c0 + i) = o;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
                continue;
            }
            if (p.mode == 2) {
                // the three largest scores of this thread's half of the tile's columns, with the
                // columns of the first two (strict >: the lower column wins ties)
                const float NI = -__int_as_float(0x7f800000);
                float v1 = NI, v2 = NI, v3 = NI;
                int i1 = 0, i2 = 0;
                float va[32];
                for (int c0 = 0; c0 < half; c0 += 32) {
                    tmem_ld32_issue(taddr + c0, va);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float sv = va[i] - hb[c0 + i];
                        const int c = hf * half + c0 + i;
                        const bool g1 = sv > v1, g2 = sv > v2, g3 = sv > v3;
                        v3 = g2 ? v2 : (g3 ? sv : v3);
                        i2 = g1 ? i1 : (g2 ? c : i2);
                        v2 = g1 ? v1 : (g2 ? sv : v2);
                        i1 = g1 ? c : i1;
                        v1 = g1 ? sv : v1;
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[as]);
                // merge the two halves of the row (half 1 hands its triple to half 0)
                if (hf == 1) {
                    rowmax_s[1][row] = v1;
                    rowsec_s[1][row] = v2;
                    rowmax_s[0][row] = v3;
                    rowarg_s[1][row] = (uint16_t)i1;
                    rowarg_s[0][row] = (uint16_t)i2;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (hf == 0 && valid) {
                    const float w[3] = {rowmax_s[1][row], rowsec_s[1][row], rowmax_s[0][row]};
                    const int wi[2] = {rowarg_s[1][row], rowarg_s[0][row]};
#pragma unroll
                    for (int u = 0; u < 3; ++u) {   // columns of half 1 are higher: strict > keeps ties low
                        const float sv = w[u];
                        const int c = u < 2 ? wi[u] : 0;
                        const bool g1 = sv > v1, g2 = sv > v2, g3 = sv > v3;
                        v3 = g2 ? v2 : (g3 ? sv : v3);
                        i2 = g1 ? i1 : (g2 ? c : i2);
                        v2 = g1 ? v1 : (g2 ? sv : v2);
                        i1 = g1 ? c : i1;
                        v1 = g1 ? sv : v1;
                    }
                    const size_t o = grow * p.nb + b;
                    *reinterpret_cast<float4 *>(p.tile_v + 4 * o) = make_float4(v1, v2, v3, 0.0f);
                    *reinterpret_cast<uint32_t *>(p.tile_i + 2 * o) = (uint32_t)i1 | ((uint32_t)i2 << 16);
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                if (++as == 2) {
                    as = 0;
                    aphase ^= 1;
                }
                continue;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

// ---- exact re-check of the rows with several candidates --------------------------------------
__device__ __forceinline__ float sq_acc(float acc, float x, float c) {
    float d = __fsub_rn(x, c);
    return __fadd_rn(acc, __fmul_rn(d, d));
}

// one warp per row; each quad evaluates one candidate in the reference's order (m % 16 == 0)
__global__ void __launch_bounds__(256) recheck_kernel(const float *x, size_t n, size_t ldx, size_t col_off,
                                                      size_t m, size_t k, const float *cent,
                                                      const unsigned *work_count, unsigned work_cap,
                                                      const uint32_t *work_rows, const uint16_t *work_cand,
                                                      const uint8_t *work_cnt, uint32_t *indices,
                                                      unsigned *flags, unsigned *all_count, uint32_t *all_list, unsigned all_cap) {
    const unsigned total = min(*work_count, work_cap);
    const int lane = threadIdx.x & 31, quad = lane >> 2, tq = lane & 3;
    const int qbase = lane & ~3;
    for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
         w += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t br = work_rows[w];
        const size_t b = br / n, row = br - b * n;
        const unsigned cnt = work_cnt[w];
        const bool all = cnt > CAP;
        if (all && all_list) {   // uniform over the warp: a whole CTA takes such a row (recheck_all_kernel)
            unsigned slot = 0;
            if (lane == 0) slot = atomicAdd(all_count, 1u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < all_cap) {
                if (lane == 0) all_list[slot] = w;
                continue;
            }
        }
        const unsigned ncand = all ? (unsigned)k : cnt;
        const float *xr = x + row * ldx + col_off + b * m;
        const float *cb = cent + b * k * m;
        float bd = __int_as_float(0x7f800000);
        uint32_t bi = 0xFFFFFFFFu;
        for (unsigned c0 = 0; c0 < ncand; c0 += 8) {
            const unsigned ci = c0 + quad;
            const bool act = ci < ncand;
            const uint32_t j = act ? (all ? ci : (uint32_t)work_cand[(size_t)w * CAP + ci]) : 0;
            const float *cr = cb + (size_t)j * m;
            float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
            if (act) {
                for (size_t e = 4 * tq; e < m; e += 16) {
                    const float4 xv = *reinterpret_cast<const float4 *>(xr + e);
                    const float4 cv = *reinterpret_cast<const float4 *>(cr + e);
                    a0 = sq_acc(a0, xv.x, cv.x);
                    a1 = sq_acc(a1, xv.y, cv.y);
                    a2 = sq_acc(a2, xv.z, cv.z);
                    a3 = sq_acc(a3, xv.w, cv.w);
                }
            }
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
            if (act && (s < bd || (s == bd && j < bi))) {
                bd = s;
                bi = j;
            }
        }
        // lexicographic (distance, index) minimum over the 8 quads == first strict minimum in index order
#pragma unroll
        for (int off = 4; off <= 16; off <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, off);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || od < bd || (od == bd && oi < bi))) {
                bd = od;
                bi = oi;
            }
        }
        if (lane == 0) {
            if (bi == 0xFFFFFFFFu) atomicOr(flags, FLAG_NO_ARGMIN);
            else indices[b * n + row] = bi;
        }
    }
}

// any m / alignment: one warp per row, one lane per candidate, the reference's summation order
// (16 accumulators, the first m % 16 elements seed them; fewer than 16 elements: one running sum)
__device__ float sqdist_generic(const float *__restrict__ x, const float *__restrict__ c, size_t m) {
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
            const float d = __fsub_rn(x[l], c[l]);
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

__global__ void __launch_bounds__(256) recheck_generic_kernel(const float *x, size_t n, size_t ldx, size_t col_off,
                                                              size_t m, size_t k, const float *cent,
                                                              const unsigned *work_count, unsigned work_cap,
                                                              const uint32_t *work_rows, const uint16_t *work_cand,
                                                              const uint8_t *work_cnt, uint32_t *indices,
                                                              unsigned *flags, unsigned *all_count, uint32_t *all_list, unsigned all_cap) {
    const unsigned total = min(*work_count, work_cap);
    const int lane = threadIdx.x & 31;
    for (unsigned w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; w < total;
         w += (gridDim.x * blockDim.x) >> 5) {
        const uint32_t br = work_rows[w];
        const size_t b = br / n, row = br - b * n;
        const unsigned cnt = work_cnt[w];
        const bool all = cnt > CAP;
        if (all && all_list) {   // uniform over the warp: a whole CTA takes such a row (recheck_all_kernel)
            unsigned slot = 0;
            if (lane == 0) slot = atomicAdd(all_count, 1u);
            slot = __shfl_sync(0xffffffffu, slot, 0);
            if (slot < all_cap) {
                if (lane == 0) all_list[slot] = w;
                continue;
            }
        }
        const unsigned ncand = all ? (unsigned)k : cnt;
        const float *xr = x + row * ldx + col_off + b * m;
        const float *cb = cent + b * k * m;
        float bd = __int_as_float(0x7f800000);
        uint32_t bi = 0xFFFFFFFFu;
        for (unsigned ci = lane; ci < ncand; ci += 32) {
            const uint32_t j = all ? ci : (uint32_t)work_cand[(size_t)w * CAP + ci];
            const float d = sqdist_generic(xr, cb + (size_t)j * m, m);
            if (d < bd || (d == bd && j < bi)) {
                bd = d;
                bi = j;
            }
        }
        // lexicographic (distance, index) minimum == first strict minimum in index order
#pragma unroll
        for (int off = 1; off <= 16; off <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, off);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || od < bd || (od == bd && oi < bi))) {
                bd = od;
                bi = oi;
            }
        }
        if (lane == 0) {
            if (bi == 0xFFFFFFFFu) atomicOr(flags, FLAG_NO_ARGMIN);
            else indices[b * n + row] = bi;
        }
    }
}

// Rows whose band holds more than CAP columns are decided over all k centroids.  One warp per such row is a
// tail of hundreds of microseconds (k = 1024, m = 768: 786k subtract-multiply-adds in one warp), so a whole CTA takes
// the row: the warps share the centroids, every distance in the reference's order, lexicographic (distance, index)
// minimum == first strict minimum in index order.
template <bool ALIGNED>
__global__ void __launch_bounds__(256) recheck_all_kernel(const float *x, size_t n, size_t ldx, size_t col_off, size_t m,
                                                          size_t k, const float *cent, const unsigned *all_count,
                                                          unsigned all_cap, const uint32_t *all_list,
                                                          const uint32_t *work_rows, uint32_t *indices, unsigned *flags) {
    __shared__ float sd[8];
    __shared__ uint32_t si[8];
    const unsigned total = min(*all_count, all_cap);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, quad = lane >> 2, tq = lane & 3;
    const int qbase = lane & ~3;
    for (unsigned e = blockIdx.x; e < total; e += gridDim.x) {
        const uint32_t br = work_rows[all_list[e]];
        const size_t b = br / n, row = br - b * n;
        const float *xr = x + row * ldx + col_off + b * m;
        const float *cb = cent + b * k * m;
        float bd = __int_as_float(0x7f800000);
        uint32_t bi = 0xFFFFFFFFu;
        if (ALIGNED) {
            for (unsigned c0 = 8 * warp; c0 < k; c0 += 64) {
                const uint32_t j = c0 + quad;
                const bool act = j < k;
                const float *cr = cb + (size_t)(act ? j : 0) * m;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                if (act) {
                    for (size_t el = 4 * tq; el < m; el += 16) {
                        const float4 xv = *reinterpret_cast<const float4 *>(xr + el);
                        const float4 cv = *reinterpret_cast<const float4 *>(cr + el);
                        a0 = sq_acc(a0, xv.x, cv.x);
                        a1 = sq_acc(a1, xv.y, cv.y);
                        a2 = sq_acc(a2, xv.z, cv.z);
                        a3 = sq_acc(a3, xv.w, cv.w);
                    }
                }
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
                if (act && (s < bd || (s == bd && j < bi))) {
                    bd = s;
                    bi = j;
                }
            }
        } else {
            for (uint32_t j = threadIdx.x; j < k; j += 256) {
                const float d = sqdist_generic(xr, cb + (size_t)j * m, m);
                if (d < bd || (d == bd && j < bi)) {
                    bd = d;
                    bi = j;
                }
            }
        }
#pragma unroll
        for (int off = 1; off <= 16; off <<= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, bd, off);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, bi, off);
            if (oi != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || od < bd || (od == bd && oi < bi))) {
                bd = od;
                bi = oi;
            }
        }
        if (lane == 0) sd[warp] = bd, si[warp] = bi;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < 8; ++w)
                if (si[w] != 0xFFFFFFFFu && (bi == 0xFFFFFFFFu || sd[w] < bd || (sd[w] == bd && si[w] < bi))) {
                    bd = sd[w];
                    bi = si[w];
                }
            if (bi == 0xFFFFFFFFu) atomicOr(flags, FLAG_NO_ARGMIN);
            else indices[b * n + row] = bi;
        }
        __syncthreads();
    }
}

// ---- k > 256: column tiles of 256 centroids ------------------------------------------------------
// tc_assign_kernel (mode 2) leaves the three largest scores of every (row, tile); one thread per row
// puts them together: the global maximum, the band, and the columns that can be inside it.  A tile
// whose third score is inside the band may hold more: the row is then decided over all k centroids.
__global__ void __launch_bounds__(256) tc_combine_kernel(TcParams p, size_t ktotal) {
    const size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p.active && !p.active[0]) return;
    if (row >= p.n) return;
    const float NI = -__int_as_float(0x7f800000);
    const int nct = (int)p.nb;
    float smax = NI;
    for (int ct = 0; ct < nct; ++ct) smax = fmaxf(smax, p.tile_v[4 * (row * nct + ct)]);
    unsigned cb = 0;
    for (int ct = 0; ct < nct; ++ct) cb = max(cb, p.cmax2_bits[ct]);
    const float cmax2 = __uint_as_float(cb);
    const float xn2 = p.xn2[row];
    const float E = p.gamma1 * sqrtf(xn2 * cmax2) * 1.0001f + 1.2e-7f * (0.5f * cmax2);
    const float dmin = fmaxf(0.0f, xn2 - 2.0f * smax + 2.0f * E);
    const float shift = 1.3e-7f * sqrtf(dmin) * (sqrtf(xn2) + sqrtf(cmax2));
    const float band = 2.0f * (2.0f * E + 1.01f * p.eta * dmin + shift);
    const float thresh = smax - band;
    // every column inside the band (the negated comparisons also catch NaN scores)
    uint32_t cand[CAP];
    unsigned c = 0;
    bool all = !(smax > NI) || !(thresh == thresh);
    for (int ct = 0; ct < nct && !all; ++ct) {
        const float4 v = *reinterpret_cast<const float4 *>(p.tile_v + 4 * (row * nct + ct));
        const uint32_t ii = *reinterpret_cast<const uint32_t *>(p.tile_i + 2 * (row * nct + ct));
        if (v.z >= thresh) all = true;   // a third (and maybe more) inside the band
        if (v.x >= thresh) {
            if (c < CAP) cand[c] = (uint32_t)ct * 256u + (ii & 0xffffu);
            ++c;
        }
        if (v.y >= thresh) {
            if (c < CAP) cand[c] = (uint32_t)ct * 256u + (ii >> 16);
            ++c;
        }
    }
    if (c > CAP) all = true;
    if (!all && c == 1) {
        p.indices[row] = cand[0];
        return;
    }
    const unsigned slot = atomicAdd(p.work_count, 1u);
    if (slot < p.work_cap) {
        p.work_rows[slot] = (uint32_t)row;
        p.work_cnt[slot] = (uint8_t)((all || c < 2) ? CAP + 1 : c);
        for (unsigned u = 0; u < CAP; ++u) p.work_cand[(size_t)slot * CAP + u] = (uint16_t)(u < c && !all ? cand[u] : 0);
        if (all || c < 2) atomicAdd(&p.stats[1], 1u);
    }
    (void)ktotal;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

int make_map(CUtensorMap *map, void *base, uint64_t inner, uint64_t outer, uint32_t box_outer, int bk = BK) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available");
        return FDB_ERR_CUDA;
    }
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {inner * 2};
    cuuint32_t box[2] = {(cuuint32_t)bk, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_32B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed: %d", (int)r);
        return FDB_ERR_CUDA;
    }
    return FDB_OK;
}

}  // namespace

struct TcState {
    DevBuf<__nv_bfloat16> x1, x2, c1, c2;
    DevBuf<float> xn2, h, mu;
    DevBuf<double> mean_partial;
    DevBuf<unsigned> cmax2, work_count, stats;
    DevBuf<unsigned long long> dbg;
    DevBuf<float> tile_v;       // k > 256: [n][tiles][4] largest scores per (row, column tile)
    DevBuf<uint16_t> tile_i;    //          [n][tiles][2] their columns
    size_t pnb = 0;
    DevBuf<uint32_t> work_rows, all_list;
    DevBuf<uint16_t> work_cand;
    DevBuf<uint8_t> work_cnt;
    CUtensorMap map_x1, map_x2, map_c1, map_c2;
    uint64_t rows_version = ~0ull;
    int np = 0;
    unsigned last_stats[3] = {0, 0, 0};
};

bool tc_eligible(const fdb_km *km) {
    if (getenv("FDB_DISABLE_TC")) return false;
    const size_t ld = km->vs->dim;
    (void)ld;
    return (km->k <= 256 || (km->nb == 1 && km->k <= 65535)) && km->m >= 1 &&
           km->n >= 1 && km->n < (1ull << 31) && km->nb * km->n < (1ull << 32);
}

void tc_free(fdb_km *km) {
    delete km->tc;
    km->tc = nullptr;
}

int tc_last_stats(const fdb_km *km, unsigned out[3]) {
    if (!km->tc) return FDB_ERR_INVALID_CONTEXT;
    for (int i = 0; i < 3; ++i) out[i] = km->tc->last_stats[i];
    return FDB_OK;
}

int tc_reassign(fdb_km *km, const int *d_active) {
    fdb_ctx *ctx = km->ctx;
    const size_t n = km->n, m = km->m, nb = km->nb, k = km->k, ld = km->vs->dim;
    if (!km->tc) km->tc = new TcState;
    TcState *tc = km->tc;
    // k > 256 (one problem): column tiles of 256 centroids are the kernel's "problems"; they share the
    // operand columns, every tile leaves its three largest scores per row, tc_combine_kernel decides
    const bool tiled = k > 256;
    const size_t pk = tiled ? 256 : k, pnb = tiled ? (k + 255) / 256 : nb;
    const int np = tiled ? 256 : (int)((k + 63) / 64 * 64);
    // Piece layout: the rows' own layout when every problem's columns start on a 16-byte boundary and
    // m is a multiple of 16; else a compact copy, every problem padded with zeros to mp columns
    // Several problems side by side (the PQ divisions): problem-major pieces [nb][n][mp] -- a 128-row tile of a problem
    // is one contiguous block of HBM (the rows' own layout hands TMA 128 rows x 256 bytes at a 3 KB stride: measured
    // 1.3 TB/s instead of 4.9 on the PQ shape of the README database)
    const size_t mp = (m + 15) / 16 * 16;
    const bool prob_major = !tiled && nb > 1 && !getenv("FDB_TC_ROW_MAJOR");
    const bool rows_aligned = m % 16 == 0 && ld % 8 == 0 && km->col_off % 8 == 0;
    const bool own_layout = !prob_major && rows_aligned;
    const size_t ldp = prob_major ? mp : own_layout ? ld : nb * mp, coff = own_layout ? km->col_off : 0;
    const size_t piece_elems = prob_major ? (nb * n + BM) * mp : n * ldp;
    const int bk = mp % BK == 0 ? BK : 16;
    cudaStream_t st = ctx->stream;
    if (tc->rows_version != km->vs->version || tc->np != np || tc->pnb != pnb) {
        FDB_TRY(tc->x1.ensure(piece_elems));
        FDB_TRY(tc->x2.ensure(piece_elems));
        FDB_TRY(tc->xn2.ensure(nb * n));
        FDB_TRY(tc->c1.ensure(pnb * pk * mp + 256 * mp));  // slack: the last problem's box reads past its rows
        FDB_TRY(tc->c2.ensure(pnb * pk * mp + 256 * mp));
        FDB_TRY(tc->h.ensure(pnb * np));
        FDB_TRY(tc->cmax2.ensure(pnb));
        if (tiled) {
            FDB_TRY(tc->tile_v.ensure(n * pnb * 4));
            FDB_TRY(tc->tile_i.ensure(n * pnb * 2));
        }
        FDB_TRY(tc->work_count.ensure(2));   // [0] rows to re-check, [1] of those: rows decided over all k
        FDB_TRY(tc->all_list.ensure(ALL_CAP));
        FDB_TRY(tc->stats.ensure(2));
        FDB_TRY(tc->work_rows.ensure(nb * n));
        FDB_TRY(tc->work_cand.ensure(nb * n * CAP));
        FDB_TRY(tc->work_cnt.ensure(nb * n));
        FDB_CUDA(cudaMemsetAsync(tc->c1.p, 0, (pnb * pk * mp + 256 * mp) * 2, st));
        FDB_CUDA(cudaMemsetAsync(tc->c2.p, 0, (pnb * pk * mp + 256 * mp) * 2, st));
        if (!own_layout) {
            FDB_CUDA(cudaMemsetAsync(tc->x1.p, 0, piece_elems * 2, st));
            FDB_CUDA(cudaMemsetAsync(tc->x2.p, 0, piece_elems * 2, st));
        }
        // mu = column means of this problem's columns (any mu is valid; the mean minimises |x'|)
        const size_t ncols = nb * m;
        FDB_TRY(tc->mu.ensure(ld));
        FDB_TRY(tc->mean_partial.ensure((size_t)MEAN_SLICES * ncols));
        FDB_CUDA(cudaMemsetAsync(tc->mu.p, 0, ld * sizeof(float), st));
        {
            dim3 grid((unsigned)((ncols + 127) / 128), MEAN_SLICES);
            col_partial_kernel<<<grid, 128, 0, st>>>(km->vs->d, n, ld, km->col_off, ncols, tc->mean_partial.p);
            col_mean_kernel<<<(unsigned)((ncols + 127) / 128), 128, 0, st>>>(tc->mean_partial.p, n, km->col_off,
                                                                          ncols, tc->mu.p);
            ctx->launches += 2;
        }
        const size_t warps = n * nb;
        split_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(
            km->vs->d, n, ld, km->col_off, m, nb, tc->mu.p, tc->x1.p, tc->x2.p, tc->xn2.p, ldp, coff,
            prob_major ? 0 : own_layout ? m : mp, prob_major ? n * mp : 0);
        ctx->launches++;
        FDB_CHECK_LAUNCH();
        FDB_TRY(make_map(&tc->map_x1, tc->x1.p, ldp, prob_major ? nb * n + BM : n, BM, bk));
        FDB_TRY(make_map(&tc->map_x2, tc->x2.p, ldp, prob_major ? nb * n + BM : n, BM, bk));
        FDB_TRY(make_map(&tc->map_c1, tc->c1.p, mp, pnb * pk + 256, (uint32_t)np, bk));
        FDB_TRY(make_map(&tc->map_c2, tc->c2.p, mp, pnb * pk + 256, (uint32_t)np, bk));
        tc->rows_version = km->vs->version;
        tc->np = np;
        tc->pnb = pnb;
    }
    FDB_CUDA(cudaMemsetAsync(tc->cmax2.p, 0, pnb * sizeof(unsigned), st));
    FDB_CUDA(cudaMemsetAsync(tc->work_count.p, 0, 2 * sizeof(unsigned), st));
    FDB_CUDA(cudaMemsetAsync(tc->stats.p, 0, 2 * sizeof(unsigned), st));
    {
        dim3 grid((unsigned)np, (unsigned)pnb);
        // (tiled: every tile is centred by the same columns, rows past k are padding; the active flag of
        // the one problem is checked by the kernels that follow)
        prep_centroids_kernel<<<grid, 128, 0, st>>>(km->centroids.p, pk, m, (size_t)np, tc->mu.p, km->col_off,
                                                    tiled ? 0 : m, nb * k, mp, tc->c1.p, tc->c2.p, tc->h.p,
                                                    tc->cmax2.p, tiled ? nullptr : d_active);
        ctx->launches++;
        FDB_CHECK_LAUNCH();
    }
    TcParams p;
    p.n = n;
    p.m = mp;                   // K extent of the GEMM (zero padded); the bounds below use the true m
    p.nb = pnb;
    p.k = pk;
    p.col_off = coff;
    p.xcol_stride = (tiled || prob_major) ? 0 : (own_layout ? m : mp);
    p.xrow_stride = prob_major ? n : 0;
    p.crow_stride = pk;
    p.mode = tiled ? 2 : 0;
    p.tile_v = tc->tile_v.p;
    p.tile_i = tc->tile_i.p;
    p.out = nullptr;
    p.out_row_stride = p.out_b_stride = 0;
    p.alpha = 1.0f;
    p.sub_h = 1;
    p.np = np;
    p.row_tiles = (int)((n + BM - 1) / BM);
    p.bk = bk;
    size_t smem = 0;
    tc_smem_plan((size_t)np, (size_t)bk, mp, &p.stages, &p.bres, &smem);
    p.h = tc->h.p;
    p.xn2 = tc->xn2.p;
    p.cmax2_bits = tc->cmax2.p;
    p.active = d_active;
    // split error 3*2^-18 + fp32 accumulation in the tensor pipe, (3m/16) MMAs at <= 2^-21 each
    p.gamma1 = 3.0f * 3.8146973e-06f + (3.0f * (float)m / 16.0f) * 4.7683716e-07f;
    // the reference's own f32 evaluation: (m/16 + 20) roundings in the longest chain
    p.eta = ((float)m / 16.0f + 20.0f) * 5.9604645e-08f;
    p.indices = km->indices.p;
    p.work_count = tc->work_count.p;
    p.work_rows = tc->work_rows.p;
    p.work_cand = tc->work_cand.p;
    p.work_cnt = tc->work_cnt.p;
    p.work_cap = (unsigned)(nb * n);
    p.stats = tc->stats.p;
    p.debug = getenv("FDB_TC_DEBUG") ? atoi(getenv("FDB_TC_DEBUG")) : 0;
    p.dbg = nullptr;
    if (p.debug & 16) {
        FDB_TRY(tc->dbg.ensure(8));
        FDB_CUDA(cudaMemsetAsync(tc->dbg.p, 0, 8 * sizeof(unsigned long long), st));
        p.dbg = tc->dbg.p;
    }
    FDB_CUDA(cudaFuncSetAttribute(tc_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int total_tiles = (int)pnb * p.row_tiles;
    const int grid = std::min(total_tiles, ctx->sm_count);
    tc_assign_kernel<<<grid, TC_THREADS, smem, st>>>(tc->map_x1, tc->map_x2, tc->map_c1, tc->map_c2, p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    if (tiled) {
        tc_combine_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(p, k);
        ctx->launches++;
        FDB_CHECK_LAUNCH();
    }
    const bool rc_aligned = rows_aligned && (uintptr_t)km->vs->d % 16 == 0 && ld % 4 == 0;
    unsigned *all_count = tc->work_count.p + 1;
    if (rc_aligned)
        recheck_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(km->vs->d, n, ld, km->col_off, m, k, km->centroids.p,
                                                          tc->work_count.p, p.work_cap, tc->work_rows.p,
                                                          tc->work_cand.p, tc->work_cnt.p, km->indices.p,
                                                          ctx->d_flags, all_count, tc->all_list.p, ALL_CAP);
    else
        recheck_generic_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(km->vs->d, n, ld, km->col_off, m, k,
                                                                  km->centroids.p, tc->work_count.p, p.work_cap,
                                                                  tc->work_rows.p, tc->work_cand.p, tc->work_cnt.p,
                                                                  km->indices.p, ctx->d_flags, all_count, tc->all_list.p,
                                                                  ALL_CAP);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    // the rows that go over all k centroids: one CTA each (a no-op launch when there are none)
    if (rc_aligned)
        recheck_all_kernel<true><<<ctx->sm_count * 2, 256, 0, st>>>(km->vs->d, n, ld, km->col_off, m, k, km->centroids.p,
                                                                    all_count, ALL_CAP, tc->all_list.p, tc->work_rows.p,
                                                                    km->indices.p, ctx->d_flags);
    else
        recheck_all_kernel<false><<<ctx->sm_count * 2, 256, 0, st>>>(km->vs->d, n, ld, km->col_off, m, k, km->centroids.p,
                                                                     all_count, ALL_CAP, tc->all_list.p, tc->work_rows.p,
                                                                     km->indices.p, ctx->d_flags);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    if (p.dbg) {
        unsigned long long hd[8];
        FDB_CUDA(cudaMemcpyAsync(hd, tc->dbg.p, sizeof(hd), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        const double wt = hd[3] ? (double)hd[3] : 1.0;
        fprintf(stderr, "[fdb tc dbg] warp-tiles %llu: head %.0f, wait %.0f, pass 1 %.0f, band + pass 2 %.0f, tail %.0f cycles per warp-tile; warps with candidates %llu\n",
                hd[3], hd[5] / wt, hd[0] / wt, hd[1] / wt, hd[2] / wt, hd[6] / wt, hd[4]);
    }
    if (getenv("FDB_TC_STATS")) {
        unsigned hs[3];
        FDB_CUDA(cudaMemcpyAsync(hs, tc->stats.p, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaMemcpyAsync(hs + 2, tc->work_count.p, sizeof(unsigned), cudaMemcpyDeviceToHost, st));
        FDB_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i < 3; ++i) tc->last_stats[i] = hs[i];
        fprintf(stderr, "[fdb tc] rows=%zu recheck=%u overflow=%u\n", nb * n, hs[2], hs[1]);
    }
    return FDB_OK;
}

// ---- query-side GEMMs (tc_gemm.cuh) ---------------------------------------------------------
bool tc_shape_ok(size_t k, size_t m, size_t ld) {
    return !getenv("FDB_DISABLE_TC") && k >= 1 && k <= 256 && m % 16 == 0 && ld % 8 == 0;
}

float tc_gamma(size_t m) {
    // split error 3*2^-18 + fp32 accumulation in the tensor pipe, (3m/16) MMAs at <= 2^-21 each
    return 3.0f * 3.8146973e-06f + (3.0f * (float)m / 16.0f) * 4.7683716e-07f;
}

int tc_prepare_centroids(fdb_ctx *ctx, const float *c, size_t nb, size_t k, size_t m, const float *mu,
                         size_t mu_stride, size_t ktotal, TcCentroids *out) {
    cudaStream_t st = ctx->stream;
    const int np = (int)((k + 63) / 64 * 64);
    out->nb = nb;
    out->k = k;
    out->m = m;
    out->np = np;
    const size_t rows = nb * k + 256;  // slack: the last problem's box reads past its rows
    FDB_TRY(out->c1.alloc(rows * m));
    FDB_TRY(out->c2.alloc(rows * m));
    FDB_TRY(out->h.alloc(nb * np));
    FDB_TRY(out->cmax2.alloc(nb));
    FDB_CUDA(cudaMemsetAsync(out->c1.p, 0, rows * m * 2, st));
    FDB_CUDA(cudaMemsetAsync(out->c2.p, 0, rows * m * 2, st));
    FDB_CUDA(cudaMemsetAsync(out->cmax2.p, 0, nb * sizeof(unsigned), st));
    dim3 grid((unsigned)np, (unsigned)nb);
    prep_centroids_kernel<<<grid, 128, 0, st>>>(c, k, m, (size_t)np, mu, 0, mu_stride, ktotal, m, out->c1.p,
                                                out->c2.p, out->h.p, out->cmax2.p, nullptr);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    // K chunks of 64 bf16 (128-byte swizzle) when m allows, else of 16 (32-byte swizzle)
    FDB_TRY(make_map(&out->map1, out->c1.p, m, rows, (uint32_t)np, m % BK == 0 ? BK : 16));
    FDB_TRY(make_map(&out->map2, out->c2.p, m, rows, (uint32_t)np, m % BK == 0 ? BK : 16));
    return FDB_OK;
}

int tc_prepare_rows(fdb_ctx *ctx, const float *x, size_t n, size_t ld, size_t m, size_t nb, const float *mu,
                    TcRows *out) {
    cudaStream_t st = ctx->stream;
    const bool grown = out->x1.n < n * ld;
    FDB_TRY(out->x1.ensure(n * ld));
    FDB_TRY(out->x2.ensure(n * ld));
    FDB_TRY(out->xn2.ensure(nb * n));
    const size_t warps = n * nb;
    split_rows_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, st>>>(x, n, ld, 0, m, nb, mu, out->x1.p,
                                                                          out->x2.p, out->xn2.p, ld, 0, m);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    if (grown || out->n != n || out->ld != ld) {
        // (the chunk size follows the row length: every GEMM that uses these rows has m % 64 == 0 iff ld % 64 == 0,
        //  see the gates in filter_prepare)
        FDB_TRY(make_map(&out->map1, out->x1.p, ld, n, BM, ld % BK == 0 ? BK : 16));
        FDB_TRY(make_map(&out->map2, out->x2.p, ld, n, BM, ld % BK == 0 ? BK : 16));
        out->n = n;
        out->ld = ld;
    }
    return FDB_OK;
}

int tc_gemm_raw(fdb_ctx *ctx, const TcRows &rows, const TcCentroids &cent, size_t xcol_stride, float alpha,
                int sub_h, float *out, size_t out_row_stride, size_t out_b_stride) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.n = rows.n;
    p.m = cent.m;
    p.nb = cent.nb;
    p.k = cent.k;
    p.col_off = 0;
    p.xcol_stride = xcol_stride;
    p.xrow_stride = 0;
    p.crow_stride = cent.k;
    p.mode = 1;
    p.out = out;
    p.out_row_stride = out_row_stride;
    p.out_b_stride = out_b_stride;
    p.alpha = alpha;
    p.sub_h = sub_h;
    p.np = cent.np;
    p.row_tiles = (int)((rows.n + BM - 1) / BM);
    const int bk = cent.m % BK == 0 ? BK : 16;
    p.bk = bk;
    size_t smem = 0;
    tc_smem_plan((size_t)cent.np, (size_t)bk, p.m, &p.stages, &p.bres, &smem);
    p.h = cent.h.p;
    FDB_CUDA(cudaFuncSetAttribute(tc_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int total_tiles = (int)p.nb * p.row_tiles;
    if (total_tiles == 0) return FDB_OK;
    const int grid = std::min(total_tiles, ctx->sm_count);
    tc_assign_kernel<<<grid, TC_THREADS, smem, ctx->stream>>>(rows.map1, rows.map2, cent.map1, cent.map2, p);
    ctx->launches++;
    FDB_CHECK_LAUNCH();
    return FDB_OK;
}

}  // namespace fdb
