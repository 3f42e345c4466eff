// fused_ops.cuh -- launchers of the channels-last kernels that surround the tcgen05 GEMMs in the fused
// PVConv pipeline (pvconv_pipeline.cu).  All tensors are fp32, channels-last ("cl"): a grid is
// [B*R^3, C] and a point set is [B*N, C] with C padded to a multiple of 4 (pad columns are zero).
#pragma once
#include "common.cuh"

namespace pvb {

// per-channel BatchNorm coefficients living in device memory
struct BnCoef {
  float *mean;     // [C] batch (or running) mean
  float *invstd;   // [C] 1/sqrt(var + eps)
  float *scale;    // [C] gamma * invstd
  float *shift;    // [C] beta - mean * scale
};

// [B,C,N] -> [B*N, Cp] (+ lo = x - trunc_tf32(x)); pad columns zeroed
int launch_points_to_cl(int b, int c, int n, int cp, const float *x, float *xcl, float *xcl_lo, cudaStream_t s);

// voxel index / count (int32, reference semantics) from integer coords [B,3,N]
int launch_vox_index_count(int b, int n, int r, const int *coords, int *ind, int *cnt, cudaStream_t s);

// scatter-mean of point rows into the (pre-zeroed) grid [B*R^3, Cp]; warp-aggregated per voxel
int launch_voxelize_cl(int b, int n, int r3, int cp, const int *ind, const int *cnt, const float *xcl, float *grid,
                       cudaStream_t s);
// lo = g - trunc(g) at occupied voxels only (grid_lo pre-zeroed)
int launch_grid_lo_at_points(int b, int n, int r3, int cp, const int *ind, const float *grid, float *grid_lo,
                             cudaStream_t s);

// per-channel sum / sum of squares over rows -> partials[blocks][2][cp]; returns #blocks used via *nblocks
int launch_bn_stats(long long rows, int cp, const float *y, float *partials, int *nblocks, cudaStream_t s);
// reduce partials (fp64), produce mean/invstd/scale/shift, update running stats (momentum, unbiased var)
int launch_bn_finalize(int nblocks, int c, int cp, long long rows, float eps, float momentum, const float *partials,
                       const float *gamma, const float *beta, float *running_mean, float *running_var, BnCoef coef,
                       cudaStream_t s, long long *num_batches_tracked = nullptr);
// eval mode: coefficients from running statistics
int launch_bn_coef_from_running(int c, float eps, const float *gamma, const float *beta, const float *running_mean,
                                const float *running_var, BnCoef coef, cudaStream_t s);

// z = leaky(y*scale+shift) (+ z_lo)
int launch_bn_apply_leaky(long long rows, int cp, float slope, const float *y, BnCoef coef, float *z, float *z_lo,
                          cudaStream_t s);

// out[b,c,i] = sum_k w_k * leaky(bn2(Y2[b, idx_k, c])) + relu(bnp(P[b*N+i, c]))
int launch_devox_fused(int b, int n, int c, int cp, int r, float slope, const float *norm_coords, const float *y2,
                       BnCoef bn2, const float *p, BnCoef bnp, const float *se_s /*[B,cp] or NULL*/, float *out, cudaStream_t s);

// ---- backward ----
// stage 1 over points: relu-masked point-branch grad (ga_cl), BN reductions of both branches, and the
// scatter of the voxel-branch gradient (already multiplied by leaky') into d2 (pre-zeroed)
int launch_bwd_points(int b, int n, int c, int cp, int r, float slope, const float *grad_out,
                      const float *norm_coords, const float *y2, BnCoef bn2, const float *p, BnCoef bnp,
                      float *ga_cl, float *d2, float *partials /*[blocks][4][cp]*/, int *nblocks,
                      const float *se_s /*or NULL*/, float *ds_partials /*[blocks][cp] or NULL*/, cudaStream_t s);
// sums[j] = sum over blocks (fp64) of partials[block][j], j < ncols
int launch_reduce_partials(int nblocks, int ncols, const float *partials, float *sums, cudaStream_t s);
// out = scale * (mask(y)*g - s1/rows - xhat(y)*s2/rows) (+ out_lo); s1/s2 = RAW column sums of g' and
// g'*xhat; mask = leaky'(bn(y)) when use_mask; also emits column sums of out (conv-bias gradient)
int launch_bn_bwd_apply(long long rows, int cp, int use_mask, float slope, const float *g, const float *y, BnCoef coef,
                        const float *s1, const float *s2, float *out, float *out_lo, float *colsum_partials,
                        int *nblocks, cudaStream_t s, const float *extra = nullptr, long long rows_per_sample = 0);
// dense reductions for BN1 backward: U1 = sum leaky'(bn(y))*g, U2 = sum leaky'(..)*g*xhat
int launch_bn_bwd_reduce(long long rows, int cp, float slope, const float *g, const float *y, BnCoef coef,
                         float *partials /*[blocks][2][cp]*/, int *nblocks, cudaStream_t s);
// grad_features[b,c,i] = gG0[b*R^3 + ind_i, c] / cnt + gFpt[b*N+i, c]
int launch_bwd_final(int b, int n, int c, int cp, int r3, const int *ind, const int *cnt, const float *gg0,
                     const float *gfpt, float *grad_features, cudaStream_t s);

// ---- SE3d inside the fused block (see fused_ops.cu) ----
int launch_se_pool(int b, long long rows_per_sample, int cp, float slope, const float *y, BnCoef coef, float *partials,
                   float *pooled3 /*[b][3][cp]: sum leaky(bn(y)), sum leaky', sum leaky'*xhat*/, cudaStream_t s);
int launch_se_fc(int b, int c, int cp, int hid, long long rows_per_sample, const float *pooled3, const float *w1,
                 const float *w2, float *mean_out, float *hidden_out, float *gate, cudaStream_t s);
int launch_reduce_partials_batched(int b, int nblocks_per_sample, int ncols, const float *partials, float *sums,
                                   cudaStream_t s);
int launch_se_backward(int nb, int c, int cp, int hid, long long rows_per_sample, const float *dgate_sum,
                       const float *gate, const float *hidden, const float *mean, const float *pooled3,
                       const float *w1, const float *w2, float *dw1, float *dw2, float *extra, float *t1, float *t2,
                       cudaStream_t s);

// ---- activity-driven tile skipping (see fused_ops.cu "Activity bookkeeping") ----
int launch_build_activity(int nb, int r, int ty, int wg_bz, int wg_by, const int *cnt, int *counts /*[8]*/,
                          unsigned char *occ /*[nb*r*r]*/, unsigned char *act1, unsigned char *act_dg,
                          unsigned char *fwd2_flag, unsigned char *wg1_flag, unsigned char *wg2_flag,
                          unsigned char *dg2_flag, int4 *fwd1, int4 *dgrad1, int4 *fwd2, int4 *wg1, int4 *wg2, int4 *dg2,
                          int *chunk_counts /*[6 * ceil(max(units, k-tiles) / 1024)]*/, cudaStream_t s);
// 27 boundary-class column sums of g: [0] over the k-tiles not in kt_active, [1] over all voxels
int launch_class_sums(int nb, int r, int cp, int by, int bz, const unsigned char *kt_active, const float *g,
                      float *classsum2 /*[2][27][cp]*/, cudaStream_t s);
// dW[co][ci][tap] += c1[ci] * sum_{classes where tap valid} classsum_inactive[cls][co]   (constant-input region)
int launch_wgrad_const_update(int cin, int cout, int cp, float slope, const float *classsum_inactive,
                              const float *bias1, BnCoef bn1, float *dw, cudaStream_t s);
// total[ci] = sum over all voxels of conv^T(g)[.,ci]  (closed form from the all-voxel class sums)
int launch_conv_grad_total(int cin, int cout, int cp, const float *w, const float *classsum_all, float *total,
                           cudaStream_t s);
// BatchNorm backward on a unit list (explicit) + constant region (closed form); emits u1/u2 (raw reductions =
// dbeta/dgamma), out(+lo) on the listed units and the full conv-bias gradient db
int launch_bn_bwd_units(int max_units, int r, int ty, int c, int cp, float slope, long long rows_total,
                        const int4 *units, const int *count, const float *g, const float *y, BnCoef coef,
                        const float *total, const float *bias_prev, float *partials, float *sums_g /*[4][cp]*/,
                        float *u1, float *u2, float *out, float *out_lo, float *colsum_g /*[cp]*/, float *db,
                        cudaStream_t s);
int launch_fill_bias_rows(long long rows, int c, int cp, const float *bias, float *out, cudaStream_t s);
int launch_fill_const_conv(int nb, int r, int cin, int cout, int cp_out, float slope, const float *w, const float *bias2,
                           const float *bias1, BnCoef bn1, float *classsum /*[27][cp_out]*/, float *tapsum /*[27][cout]*/, float *out,
                           cudaStream_t s, int tables_ready = 0);

// small helpers
int launch_memset_f32(float *p, long long n, cudaStream_t s);

#ifdef __CUDACC__
__device__ __forceinline__ float4 ld4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ void st4(float *p, float4 v) { *reinterpret_cast<float4 *>(p) = v; }

// --------------------------------------------------------------------------------------------
// Column reductions over [rows, cp]: thread -> (channel quad c4 = t % cp4, row lane = t / cp4).
// --------------------------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_MAX_BLOCKS = kNumSMs * 4;

template <int NSETS, typename F>
__device__ __forceinline__ void column_reduce(long long rows, int cp, float *partials, F &&body,
                                              long long row_begin = 0, long long part_block = -1) {
  __shared__ float4 red[NSETS][RED_THREADS];
  const int cp4 = cp >> 2;
  const int rl = RED_THREADS / cp4;  // row lanes (cp4 <= 256)
  const int c4 = threadIdx.x % cp4, lane_r = threadIdx.x / cp4;
  float4 acc[NSETS];
#pragma unroll
  for (int k = 0; k < NSETS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (part_block < 0) part_block = blockIdx.x;
  if (lane_r < rl)
    for (long long r = (long long)blockIdx.x * rl + lane_r; r < rows; r += (long long)gridDim.x * rl)
      body(row_begin + r, c4, acc);
#pragma unroll
  for (int k = 0; k < NSETS; ++k) red[k][threadIdx.x] = acc[k];
  __syncthreads();
  if (threadIdx.x < cp4) {
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
      float4 s = red[k][threadIdx.x];
      for (int j = 1; j < rl; ++j) {
        const float4 v = red[k][threadIdx.x + j * cp4];
        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
      }
      st4(partials + ((size_t)part_block * NSETS + k) * cp + threadIdx.x * 4, s);
    }
  }
}

#endif  // __CUDACC__

}  // namespace pvb
