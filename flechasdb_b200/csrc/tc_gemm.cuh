// fp32-accurate GEMM on the tensor pipe for the query path: the tcgen05/TMEM/TMA kernel of
// tc_assign.cu with a raw-score epilogue.  Operands are the two-piece bf16 split of
// x' = fl(x - mu) (rows) and c' = fl(c - mu) (centroid rows), see the header of tc_assign.cu.
#pragma once
#include "common.cuh"

#include <cuda.h>
#include <cuda_bf16.h>

namespace fdb {

// B operand: nb problems of k centroid rows x m columns, pieces stored [nb * k][m]
struct TcCentroids {
    DevBuf<__nv_bfloat16> c1, c2;
    DevBuf<float> h;          // [nb][np]  |c'_j|^2 / 2, +inf for padded columns
    DevBuf<unsigned> cmax2;   // [nb] bits of max_j |c'_j|^2 (rounded up)
    CUtensorMap map1, map2;
    size_t nb = 0, k = 0, m = 0;
    int np = 0;               // k padded to a multiple of 64 (<= 256)
};
// A operand: rows x ld columns, split once per batch
struct TcRows {
    DevBuf<__nv_bfloat16> x1, x2;
    DevBuf<float> xn2;        // [nb][n] |x'|^2 of the problem's columns (rounded up)
    CUtensorMap map1, map2;
    size_t n = 0, ld = 0;
};

// shapes the kernel takes: m % 16 == 0 (K chunks of 64, or of 16 when m % 64 != 0), k <= 256 per problem,
// 16-byte aligned rows
bool tc_shape_ok(size_t k, size_t m, size_t ld);
// centroid rows c[nb*k][m]; problem b is centred by mu[b * mu_stride .. + m); rows at or
// beyond ktotal (column tiling of one long list of centroids) are padding
int tc_prepare_centroids(fdb_ctx *ctx, const float *c, size_t nb, size_t k, size_t m, const float *mu,
                         size_t mu_stride, size_t ktotal, TcCentroids *out);
// rows x[n][ld] seen as nb problems of m columns (nb * m <= ld), centred by mu[ld]
int tc_prepare_rows(fdb_ctx *ctx, const float *x, size_t n, size_t ld, size_t m, size_t nb, const float *mu,
                    TcRows *out);
// out[row * out_row_stride + b * out_b_stride + j] = alpha * (x'_row,b . c'_b,j) - (sub_h ? h_b,j : 0)
// for j < np; problem b reads operand columns [b * xcol_stride, + m)
int tc_gemm_raw(fdb_ctx *ctx, const TcRows &rows, const TcCentroids &cent, size_t xcol_stride, float alpha,
                int sub_h, float *out, size_t out_row_stride, size_t out_b_stride);
// error of one accumulated dot product relative to |x'| |c'| (split + accumulation in the tensor pipe)
float tc_gamma(size_t m);

}  // namespace fdb
