/*
 * pvcnn_b200.h -- C ABI of the B200-native PVConv operator stack.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point replaces one launcher of the reference's native layer
 * (mit-han-lab/pvcnn, modules/functional/src/<dir>/<file>.cuh) that the reference's C++
 * wrappers (<file>.cpp) call; argument order and meaning follow the reference launcher, plus
 *   - a trailing `void *stream` (cudaStream_t; the reference launches K1-K5 and K11 on the
 *     legacy default stream, SURVEY.md 8b "Threading / streams"),
 *   - an `int` return value: 0 on success, otherwise a cudaError_t (or PVCNN_E_* below)
 *     -- the reference prints and exit(-1)s (cuda_utils.cuh:28-37); we report instead.
 * All pointers are DEVICE pointers on the current device unless stated otherwise.
 * Unlike the reference (which relies on torch::zeros in the .cpp wrappers, e.g.
 * vox.cpp:33-38) the callee initialises every output itself; callers may pass
 * uninitialised memory.
 *
 * Layouts ("classic" ops): exactly the reference's -- features [B,C,N], grids [B,C,R^3],
 * coords [B,3,N], all contiguous fp32 / int32.
 */
#ifndef PVCNN_B200_H_
#define PVCNN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVCNN_B200_ABI_VERSION 1

#if defined(__GNUC__)
#define PVCNN_API __attribute__((visibility("default")))
#else
#define PVCNN_API
#endif

#define PVCNN_E_BADARG 100001   /* invalid size / null pointer                              */
#define PVCNN_E_UNSUPPORTED 100002 /* shape outside what the sm_100a kernels were built for */

/* ABI version + human-readable build string ("sm_100a, nvcc 12.9, ...") */
PVCNN_API int pvcnn_abi_version(void);
PVCNN_API const char *pvcnn_build_info(void);
/* Number of kernels this library has launched since load (bench.py "gpu_launches"). */
PVCNN_API unsigned long long pvcnn_launch_count(void);

/* ---- coordinate normalisation: modules/voxelization.py:16-25 (the reference runs ~9 ATen
 *      kernels here; we fuse them).  coords [B,3,N] -> norm_coords [B,3,N] fp32 (clamped to
 *      [0,r-1], NOT rounded; this is what devoxelize consumes) and vox_coords [B,3,N] int32
 *      (round-half-even).  Single-kernel variant: the mean is the fp64 sum rounded once (oracle/pvcnn_oracle.c),
 *      which agrees with the reference's torch reduction except where an ulp flips a .5 rounding tie; the Python
 *      layer therefore defaults to pvcnn_voxelize_denom/_apply below (PVCNN_B200_VOX=fused selects this one). */
PVCNN_API int pvcnn_voxelize_coords(int b, int n, int r, int normalize, float eps, const float *coords,
                          float *norm_coords, int *vox_coords, void *stream);

/* ---- reference-exact coordinate normalisation (default of the Python layer).  `mean` [B,3] is produced by the same
 *      ATen reduction the reference calls (coords.mean(2), modules/voxelization.py:18: its summation order belongs to
 *      torch); the rest of modules/voxelization.py:19-24 is element-wise IEEE arithmetic plus an order-independent max
 *      and is reproduced bit for bit.  pvcnn_voxelize_denom: denom[b] = max_i ||coords[b,:,i] - mean[b]||_2 * 2 + eps
 *      (:20); pvcnn_voxelize_apply: the element-wise tail (:20-24); denom may be NULL when normalize == 0. */
PVCNN_API int pvcnn_voxelize_denom(int b, int n, float eps, const float *coords, const float *mean, float *denom,
                                   void *stream);
PVCNN_API int pvcnn_voxelize_apply(int b, int n, int r, int normalize, const float *coords, const float *mean,
                                   const float *denom, float *norm_coords, int *vox_coords, void *stream);

/* ---- replaces avg_voxelize(...)       voxelization/vox.cuh:5-6  (kernels vox.cu:18-72) */
PVCNN_API int pvcnn_avg_voxelize(int b, int c, int n, int r, int r2, int r3, const int *coords,
                       const float *feat, int *ind, int *cnt, float *out, void *stream);
/* ---- replaces avg_voxelize_grad(...)  voxelization/vox.cuh:7-8  (kernel vox.cu:86-110) */
PVCNN_API int pvcnn_avg_voxelize_grad(int b, int c, int n, int s, const int *ind, const int *cnt,
                            const float *grad_y, float *grad_x, void *stream);

/* ---- replaces trilinear_devoxelize(...)       interpolate/trilinear_devox.cuh:5-8
 *      inds/wgts [B,8,N] are written only when training != 0 (trilinear_devox.cpp:45-53);
 *      they may be NULL otherwise. */
PVCNN_API int pvcnn_trilinear_devoxelize(int b, int c, int n, int r, int r2, int r3, int training,
                               const float *coords, const float *feat, int *inds, float *wgts,
                               float *outs, void *stream);
/* ---- replaces trilinear_devoxelize_grad(...)  interpolate/trilinear_devox.cuh:9-11 */
PVCNN_API int pvcnn_trilinear_devoxelize_grad(int b, int c, int n, int r3, const int *inds,
                                    const float *wgts, const float *grad_y, float *grad_x,
                                    void *stream);

/* ---- replaces ball_query(...)  ball_query/ball_query.cuh:4-6; r2 = radius*radius in float
 *      (ball_query.cpp:24).  neighbors_indices [B,M,U]. */
PVCNN_API int pvcnn_ball_query(int b, int n, int m, float r2, int u, const float *centers_coords,
                     const float *points_coords, int *neighbors_indices, void *stream);

/* ---- fused BallQuery grouping: replaces the sequence of modules/ball_query.py:16-30
 *      (grouping(points_coords) - centers_coords, grouping(points_features), torch.cat) built on
 *      grouping/grouping.cuh:4-7.  out [B,3+C,M,U]: channels 0..2 = neighbour xyz - centre xyz,
 *      channels 3.. = neighbour features (c may be 0, then features may be NULL).
 *      _grad: grad_y [B,3+C,M,U] -> grad_features [B,C,N]; grad_points_coords [B,3,N] and
 *      grad_centers_coords [B,3,M] are optional (NULL = not needed).  Outputs are zeroed here. */
PVCNN_API int pvcnn_group_concat(int b, int c, int n, int m, int u, const float *points_coords,
                       const float *centers_coords, const float *features, const int *indices,
                       float *out, void *stream);
PVCNN_API int pvcnn_group_concat_grad(int b, int c, int n, int m, int u, const float *grad_y,
                            const int *indices, float *grad_features, float *grad_points_coords,
                            float *grad_centers_coords, void *stream);

/* ---- replaces grouping(...) / grouping_grad(...)  grouping/grouping.cuh:4-7 */
PVCNN_API int pvcnn_grouping(int b, int c, int n, int m, int u, const float *features, const int *indices,
                   float *out, void *stream);
PVCNN_API int pvcnn_grouping_grad(int b, int c, int n, int m, int u, const float *grad_y,
                        const int *indices, float *grad_x, void *stream);

/* ---- replaces gather_features(...) / gather_features_grad(...)  sampling/sampling.cuh:4-7 */
PVCNN_API int pvcnn_gather_features(int b, int c, int n, int m, const float *features, const int *indices,
                          float *out, void *stream);
PVCNN_API int pvcnn_gather_features_grad(int b, int c, int n, int m, const float *grad_y,
                               const int *indices, float *grad_x, void *stream);

/* ---- replaces furthest_point_sampling(...)  sampling/sampling.cuh:8-9.  `distances` is the
 *      reference's [B,N] scratch (sampling.cpp:53-54); it may be NULL -- we keep the running
 *      distances in registers and only use it when N exceeds the register-resident limit. */
PVCNN_API int pvcnn_furthest_point_sampling(int b, int n, int m, const float *coords, float *distances,
                                  int *indices, void *stream);

/* ---- replaces three_nearest_neighbors_interpolate(...) / _grad(...)
 *      interpolate/neighbor_interpolate.cuh:4-14 */
PVCNN_API int pvcnn_three_nearest_neighbors_interpolate(int b, int c, int m, int n,
                                              const float *points_coords,
                                              const float *centers_coords,
                                              const float *centers_features, int *indices,
                                              float *weights, float *out, void *stream);
PVCNN_API int pvcnn_three_nearest_neighbors_interpolate_grad(int b, int c, int n, int m,
                                                   const float *grad_y, const int *indices,
                                                   const float *weights, float *grad_x,
                                                   void *stream);


/* =====================================================================================
 * Dense tcgen05 primitives (channels-last).  These replace the reference's cuDNN/cuBLAS calls:
 * nn.Conv3d(k=3,pad=1) of modules/pvconv.py:21,24 and the 1x1 nn.Conv1d of
 * modules/shared_mlp.py:10 (forward and data gradient); there is no reference launcher to cite
 * because the reference delegates them to torch.
 * ===================================================================================== */

/* x -> (hi, lo): hi = x with the 13 low mantissa bits cleared (exact in tf32), lo = tf32(x - hi).
 * n must be a multiple of 4.  hi may be NULL: kind::tf32 truncates its fp32 operands (measured),
 * so x itself can be passed wherever an a_hi / w_hi operand is expected. */
PVCNN_API int pvcnn_split_tf32(long long n, const float *x, float *hi, float *lo, void *stream);

/* Conv weight [cout][cin][ntaps] (torch layout, taps flattened kd*9+kh*3+kw; ntaps 1 or 27) ->
 * GEMM "B" operand [ntaps][rows][ld], split into hi/lo.
 *   mode 0 (forward):       rows = cout, K = cin,  wr[t][co][ci] = w[co][ci][t]
 *   mode 1 (data gradient): rows = cin,  K = cout, wr[t][ci][co] = w[co][ci][ntaps-1-t]   */
PVCNN_API int pvcnn_conv_weight_prep(int cout, int cin, int ntaps, int mode, int ld, const float *w,
                                     float *w_hi, float *w_lo, void *stream);

/* out[b,x,y,z,n] = bias[n] + sum_{tap,c} a[b, (x,y,z)+off(tap), c] * w[tap][n][c]   (zero padding)
 *   a_hi/a_lo : [nb,sx,sy,sz,lda] channels-last, k valid channels (lda % 4 == 0)
 *   w_hi/w_lo : [ntaps][cout][ldw] from pvcnn_conv_weight_prep;  out: [nb,sx,sy,sz,ldo]
 *   npass = 3: error-compensated 3xTF32 (fp32-faithful, ~1e-6 rel);  npass = 1: plain TF32 (lo unused)
 * A plain GEMM [M,K]x[N,K]^T is nb=sx=sy=1, sz=M, ntaps=1. */
PVCNN_API int pvcnn_igemm_conv(int nb, int sx, int sy, int sz, int k, int cout, int ntaps,
                               const float *a_hi, const float *a_lo, int lda, const float *w_hi,
                               const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                               int npass, void *stream);
PVCNN_API int pvcnn_igemm_last_error(int *host_code);
/* PVCNN_STALL_PROFILE=1: stall-cycle counters of CTA 0 of the last tensor-core kernel (diagnostic) */
PVCNN_API int pvcnn_stall_profile_read(long long *host8);

/* dW[co][ci][tap] = sum_v g[v][co] * x[v+off(tap)][ci]   (weight gradient of pvcnn_igemm_conv; torch
 * weight layout).  x: layer input [nb,sx,sy,sz,ldx], g: output gradient [nb,sx,sy,sz,ldg]; lo
 * operands as for pvcnn_igemm_conv.  dw is overwritten.  cout <= 128 in this version. */
PVCNN_API int pvcnn_conv_wgrad(int nb, int sx, int sy, int sz, int cin, int cout, int ntaps,
                               const float *x_hi, const float *x_lo, int ldx, const float *g_hi,
                               const float *g_lo, int ldg, float *dw, int npass, void *stream);

/* =====================================================================================
 * Fused PVConv block: replaces modules.PVConv.forward (modules/pvconv.py:33-39) and its autograd
 * backward, i.e. Voxelization (modules/voxelization.py:16-25) -> Conv3d/BN3d/LeakyReLU x2
 * (modules/pvconv.py:20-27) -> trilinear_devoxelize, plus SharedMLP (modules/shared_mlp.py:29-33)
 * and the residual add, as ONE call each way; SE3d (modules/se.py:6-17) is folded in when desc.with_se is set.
 *
 * Sizes below use  Mv = b*r^3, Mp = b*n, ci = pad4(cin), co = pad4(cout), ld(x) = roundup(x,32).
 * ===================================================================================== */
typedef struct {
  int b, n, cin, cout, r;
  int normalize;      /* Voxelization(normalize=...)                       */
  float eps;          /* Voxelization eps                                  */
  int training;       /* BN batch statistics + running-stat update         */
  int npass;          /* 3 = fp32-faithful 3xTF32, 1 = plain TF32          */
  float bn_eps_vox;   /* 1e-4 (modules/pvconv.py:22)                       */
  float bn_eps_pt;    /* 1e-5 (nn.BatchNorm1d default)                     */
  float momentum;     /* 0.1                                               */
  float slope;        /* LeakyReLU 0.1                                     */
  int with_se;        /* SE3d after the second conv block (modules/se.py), hidden = cout / 8 */
  int vox_stats;      /* 0: fused fp64-mean normalisation; 1: ws->vox_mean given (reference-exact), denom computed;
                         2: ws->vox_mean and ws->vox_denom given */
  int prepared;       /* inference with ws->prep: 1 = ws->prep already holds this block's GEMM operands, BatchNorm
                         coefficients and conv2 constants for the current parameters (skip rebuilding them) */
} pvcnn_pvconv_desc;

typedef struct { /* parameters in torch layouts; running stats are updated in training mode */
  const float *w1, *b1, *g1, *be1; float *rm1, *rv1;   /* voxel_layers.0 / .1 */
  const float *w2, *b2, *g2, *be2; float *rm2, *rv2;   /* voxel_layers.3 / .4 */
  const float *wp, *bp, *gp, *bep; float *rmp, *rvp;   /* point_features.layers.0 / .1 */
  const float *se_w1, *se_w2;                          /* voxel_layers.6.fc.{0,2}.weight (with_se) */
  long long *nbt1, *nbt2, *nbtp;                       /* num_batches_tracked of the three BatchNorms (int64, may be NULL):
                                                          incremented in training mode by the statistics kernels */
} pvcnn_pvconv_params;

typedef struct { /* parameter gradients (same shapes as the parameters) */
  float *w1, *b1, *g1, *be1, *w2, *b2, *g2, *be2, *wp, *bp, *gp, *bep;
  float *se_w1, *se_w2;
} pvcnn_pvconv_grads;

typedef struct { /* caller-allocated device buffers (element counts in comments) */
  float *nc;        /* b*3*n   normalised coords (also an output of the forward pass)   */
  int *vc;          /* b*3*n   */
  int *ind;         /* b*n     */
  int *cnt;         /* b*r^3   */
  float *fcl, *fcl_lo;      /* Mp*ci */
  float *g0, *g0_lo;        /* Mv*ci */
  float *y1, *z1, *z1_lo;   /* Mv*co */
  float *y2;                /* Mv*co */
  float *p;                 /* Mp*co */
  float *coef;              /* 12*co  BN coefficients (mean, invstd, scale, shift) x 3 */
  float *wprep;             /* pvcnn_pvconv_wprep_floats()    */
  float *partials;          /* pvcnn_pvconv_partials_floats() */
  float *sums;              /* 16*co  */
  /* backward-only scratch */
  float *ga, *gpp, *gpp_lo; /* Mp*co */
  float *gfpt;              /* Mp*ci */
  float *d2;                /* Mv*max(ci,co) */
  float *gy2, *gy2_lo, *gy1, *gy1_lo; /* Mv*co */
  int *sparse;              /* pvcnn_pvconv_sparse_ints(): activity lists for tile skipping (NULL = dense) */
  float *se;                /* with_se: b*(7*co + cout/8)  (pooled sums, mean, hidden, gate, d gate, dense term) */
  float *vox_mean;          /* b*3   per-cloud coordinate mean (desc.vox_stats >= 1) */
  float *vox_denom;         /* b     per-cloud normalisation denominator (desc.vox_stats >= 1, normalize) */
  float *prep;              /* pvcnn_pvconv_prep_floats(desc), or NULL: per-block buffer that survives between inference
                               calls (training == 0); see desc.prepared */
} pvcnn_pvconv_ws;

/* 1 when the grid-sized `lo` buffers (g0_lo, z1_lo, gy2_lo, gy1_lo) must be provided (3xTF32 mode: the weight-
 * gradient kernel TMA-loads them); otherwise those pointers may be NULL. */
PVCNN_API int pvcnn_pvconv_needs_grid_lo(const pvcnn_pvconv_desc *d);
PVCNN_API long long pvcnn_pvconv_sparse_ints(const pvcnn_pvconv_desc *d);
PVCNN_API long long pvcnn_pvconv_wprep_floats(const pvcnn_pvconv_desc *d);
PVCNN_API long long pvcnn_pvconv_prep_floats(const pvcnn_pvconv_desc *d);
PVCNN_API long long pvcnn_pvconv_partials_floats(const pvcnn_pvconv_desc *d);
/* features [b,cin,n], coords [b,3,n] -> out [b,cout,n] */
PVCNN_API int pvcnn_pvconv_forward(const pvcnn_pvconv_desc *d, const float *features, const float *coords,
                                   const pvcnn_pvconv_params *prm, const pvcnn_pvconv_ws *ws, float *out,
                                   void *stream);
/* grad_out [b,cout,n] -> grad_features [b,cin,n] + parameter gradients; needs the ws of the forward */
PVCNN_API int pvcnn_pvconv_backward(const pvcnn_pvconv_desc *d, const float *grad_out,
                                    const pvcnn_pvconv_params *prm, const pvcnn_pvconv_ws *ws,
                                    float *grad_features, const pvcnn_pvconv_grads *grads, void *stream);

/* ---- device-side resampling of modules/functional/sampling.py:66-82 (`logits_mask`): the reference loops over the
 *      batch on the host (mask[i].nonzero() sync, np.random.choice/shuffle).  mask: uint8/bool [b,n]; picks int32 [b,k]:
 *      a uniform k-subset of the foreground indices in random order when there are >= k of them, otherwise every
 *      foreground index k/nc times plus k%nc distinct extras, shuffled; all zeros when there is none.  Counter-based
 *      generator keyed by `seed` (oracle: oracle.logits_mask_sample, bit-identical).  n, k <= 8192. */
PVCNN_API int pvcnn_logits_mask_sample(int b, int n, int k, unsigned long long seed, const unsigned char *mask,
                                       int *picks, void *stream);

/* =====================================================================================================
 * SharedMLP on the tensor-core path: replaces nn.Conv1d/Conv2d(k=1) + nn.BatchNorm1d/2d + nn.ReLU of
 * modules/shared_mlp.py:6-33 (cuDNN + ATen in the reference) layer by layer, on channels-last rows
 * [rows, pad4(C)] (rows = B*N for dim=1, B*M*U for dim=2).  See pvcnn_b200/csrc/mlp_pipeline.cu.
 * ===================================================================================================== */
/* layout converters at the module boundary: [b,c,n] <-> [b*n, pad4(c)] (xcl_lo = x - trunc_tf32(x), may be NULL) */
PVCNN_API int pvcnn_points_to_cl(int b, int c, int n, const float *x, float *xcl, float *xcl_lo, void *stream);
PVCNN_API int pvcnn_cl_to_points(int b, int c, int n, const float *xcl, float *x, void *stream);
/* workspace sizes (floats) and the row-segment count of the pooled variant */
PVCNN_API long long pvcnn_mlp_partials_floats(int cout);
PVCNN_API long long pvcnn_mlp_wprep_floats(int cin, int cout);
PVCNN_API int pvcnn_mlp_pool_segments(long long groups, int u);
/* one layer forward: y = x W^T + bias (saved), BatchNorm (batch statistics + running update when training, running
 * statistics otherwise; coef[4*pad4(cout)] = mean, invstd, scale, shift), then either z (+ z_lo) = relu(bn(y)) or, with
 * pool_u > 0, pooled/argmax [rows/pool_u, pad4(cout)] = max over groups of pool_u consecutive rows
 * (modules/pointnet.py:87 `.max(dim=-1).values`); pool_tmp: 2*segments*(rows/pool_u)*pad4(cout) floats if segments > 1.
 * group_bias [rows / group_rows, group_ld] (NULL: none) is added to y before the BatchNorm: see
 * pvcnn_mlp_layer_forward_eval */
PVCNN_API int pvcnn_mlp_layer_forward(long long rows, int cin, int cout, int training, int npass, float bn_eps,
                                      float momentum, const float *x, const float *x_lo, const float *w,
                                      const float *bias, const float *gamma, const float *beta, float *running_mean,
                                      float *running_var, long long *num_batches_tracked, float *wprep, float *partials,
                                      float *coef, float *y, float *z, float *z_lo, int pool_u, float *pooled, int *argmax,
                                      float *pool_tmp, long long group_rows, const float *group_bias, int group_ld,
                                      void *stream);
/* Inference form of one SharedMLP layer (modules/shared_mlp.py:6-33 under model.eval()): prepare once per parameter
 * version, then one fused GEMM per forward (bias + BatchNorm(running stats) + ReLU in the epilogue; y never written).
 * group_bias [rows / group_rows, group_ld]: per-cloud additive term for input channels that are constant over a cloud
 * (models/shapenet/pvcnn.py:40-42, models/s3dis/pvcnn.py:44-46 repeat them over the points before the concat). */
PVCNN_API int pvcnn_mlp_layer_prepare(int cin, int cout, float bn_eps, const float *w, const float *gamma, const float *beta,
                            const float *running_mean, const float *running_var, float *wprep, float *coef,
                            void *stream);
PVCNN_API int pvcnn_mlp_layer_forward_eval(long long rows, int cin, int cout, int npass, const float *x, const float *x_lo,
                                 const float *wprep, const float *bias, const float *coef, long long group_rows,
                                 const float *group_bias, int group_ld, float *y, float *z, float *z_lo, int pool_u,
                                 float *pooled, int *argmax, float *pool_tmp, void *stream);

/* dense gradient gz [groups*u, pad4(cout)] of a pooled output */
PVCNN_API int pvcnn_mlp_pool_backward(long long groups, int u, int cout, const float *gpool, const int *argmax,
                                      float *gz, void *stream);
/* one layer backward: gz = d relu(bn(y)); writes dgamma/dbeta/dbias [cout], dw [cout,cin], gx [rows,pad4(cin)] (NULL to
 * skip); gy (+gy_lo) [rows,pad4(cout)] and sums [4*pad4(cout)] are scratch; d_group_bias [rows / group_rows, pad4(cout)]
 * (NULL unless the forward had a group_bias) = per-cloud column sums of the conv-output gradient */
PVCNN_API int pvcnn_mlp_layer_backward(long long rows, int cin, int cout, int npass, const float *gz, const float *x,
                                       const float *x_lo, const float *w, const float *y, const float *coef,
                                       float *wprep, float *partials, float *sums, float *gy, float *gy_lo, float *gx,
                                       float *dw, float *dbias, float *dgamma, float *dbeta, long long group_rows,
                                       float *d_group_bias, void *stream);
/* model-level glue: channel concatenation written straight into channels-last rows (the 1472-channel
 * torch.cat + repeat of models/s3dis/pvcnn.py:44-46 never exists in [B,C,N] form); src_n == 1 broadcasts a per-cloud
 * vector over the points.  pvcnn_cl_slice_to_points is its gradient (and the generic [rows] -> [B,C,N] slice reader). */
PVCNN_API int pvcnn_cat_to_cl(int b, int c, int n, int src_n, const float *x, int ld, int col0, float *rows,
                              float *rows_lo, void *stream);
PVCNN_API int pvcnn_cl_slice_to_points(int b, int c, int n, const float *rows, int ld, int col0, float *x, void *stream);
/* plain 1x1 convolution on channels-last rows (a classifier's last layer, models/utils.py:43: no BatchNorm / ReLU):
 * y = x W^T + bias on igemm_conv_kernel; backward: dw (conv_wgrad_kernel), dbias (column sums), gx (may be NULL) */
PVCNN_API int pvcnn_linear_cl_forward(long long rows, int cin, int cout, int npass, const float *x, const float *x_lo,
                                      const float *w, const float *bias, float *wprep, float *y, void *stream);
PVCNN_API int pvcnn_linear_cl_backward(long long rows, int cin, int cout, int npass, const float *gy, const float *gy_lo,
                                       const float *x, const float *x_lo, const float *w, float *wprep, float *partials,
                                       float *gx, float *dw, float *dbias, void *stream);
/* modules/ball_query.py:16-30 (grouping + centre subtraction + concat) with channels-last output rows
 * [b*m*u, pad4(3+c)] feeding pvcnn_mlp_layer_forward directly: the reference's [B,3+C,M,U] is never materialised */
PVCNN_API int pvcnn_group_concat_cl(int b, int c, int n, int m, int u, const float *points_coords,
                                    const float *centers_coords, const float *features, const int *indices,
                                    float *out, float *out_lo, void *stream);
PVCNN_API int pvcnn_group_concat_cl_grad(int b, int c, int n, int m, int u, const float *grad_rows, const int *indices,
                                         float *grad_features, float *grad_points_coords, float *grad_centers_coords,
                                         void *stream);

/* Same as pvcnn_pvconv_backward, split for data-parallel training: phase 1 = every kernel that produces a parameter
 * gradient (ends with the conv1 weight gradient), phase 2 = the input gradient (conv1 data gradient + scatter to the
 * points); phase 0 = both.  The caller launches its single gradient all-reduce between the phases (pvcnn_b200/parallel.py)
 * so the collective overlaps phase 2. */
PVCNN_API int pvcnn_pvconv_backward_phase(const pvcnn_pvconv_desc *d, const float *grad_out,
                                          const pvcnn_pvconv_params *prm, const pvcnn_pvconv_ws *ws,
                                          float *grad_features, const pvcnn_pvconv_grads *grads, int phase, void *stream);

/* =====================================================================================================
 * Test-time voting on the device: the host loops of evaluate/s3dis/eval.py:149-215,
 * evaluate/shapenet/eval.py:149-201 and the window sampling of datasets/s3dis.py:86-89.  The reference builds the
 * voted inputs with numpy (tile / np.random.shuffle / fancy indexing per window), copies confidences and predictions
 * of every batch back to the host and merges them with numba loops; here every step is a kernel
 * (pvcnn_b200/csrc/eval_voting.cu) and only the final [3, classes] counters are read back.  Random choices use a
 * counter-based pseudo-random permutation (balanced Feistel network + cycle walking) keyed by (seed, window): the
 * result does not depend on how windows are batched; oracle/eval_voting.py restates it bit for bit.
 * ===================================================================================================== */
/* eval.py:161-164 (shapenet eval.py:151-154): indices[w, :] = tile(arange(num_points[w]))[:nv], shuffled; int32 [b, nv].
 * Window w of the call is stream first_window + w of the generator.  num_points[w] <= 0 gives zeros. */
PVCNN_API int pvcnn_vote_indices(int b, int nv, unsigned long long seed, int first_window, const int *num_points,
                                 int *indices, void *stream);
/* datasets/s3dis.py:86-87: np.random.choice(num_points[w], k, replace=(num_points[w] < k)) -> int32 [b, k] */
PVCNN_API int pvcnn_window_indices(int b, int k, unsigned long long seed, int first_window, const int *num_points,
                                   int *indices, void *stream);
/* eval.py:166-171: out[(w*extra + e), c, j] = src[w, indices[w, e*npo + j], c]; src is [b, p, ch] (channels_last = 1:
 * the h5 window layout) or [b, ch, p] (channels_last = 0: shapenet eval.py:158-160); out [b*extra, ch, npo] is the
 * network input.  labels [b, p] -> out_labels [b, extra*npo] (datasets/s3dis.py:89), both NULL or both given. */
PVCNN_API int pvcnn_vote_gather(int b, int ch, int p, int extra, int npo, int channels_last, const float *src,
                                const int *indices, float *out, const int *labels, int *out_labels, void *stream);
/* eval.py:173: F.softmax(logits, dim=1)[:, c0:c1].max(dim=1) (c0 = 0, c1 = c for S3DIS; the shape's part classes for
 * ShapeNet, eval.py:162-165).  logits [b, c, n] -> conf fp32 [b, n], pred int32 [b, n] = absolute class index, first
 * maximal class on ties. */
PVCNN_API int pvcnn_softmax_max(int b, int c, int n, int c0, int c1, const float *logits, float *conf, int *pred,
                                void *stream);
/* scene state: keys uint64 [scene_points] (confidence bits << 32 | ~sequence number), scene_pred int32 [scene_points].
 * pvcnn_vote_reset = eval.py:136-137 (confidence 0, prediction -1). */
PVCNN_API int pvcnn_vote_reset(long long scene_points, unsigned long long *keys, int *scene_pred, void *stream);
/* update_scene_predictions (eval.py:189-204) / update_shape_predictions (shapenet eval.py:177-185): vote (w, p) with
 * confidence conf[w, p] and class pred[w, p] goes to scene point mapping[w, indices[w, p]] (mapping int32 [b, p], the
 * rows of `indices_split_to_full` for these windows; NULL: the point is indices[w, p] itself) and replaces the entry
 * iff its confidence is strictly larger -- among equal confidences the earliest vote in (call, w, p) order wins, as in
 * the sequential loop: order_base = number of votes merged into this scene by earlier calls. */
PVCNN_API int pvcnn_vote_merge(int b, int nv, int p, long long scene_points, unsigned int order_base, const float *conf,
                               const int *pred, const int *indices, const int *mapping, unsigned long long *keys,
                               int *scene_pred, void *stream);
/* the scene's confidences (eval.py:136 `confidences`) as fp32 [scene_points] */
PVCNN_API int pvcnn_vote_confidences(long long scene_points, const unsigned long long *keys, float *conf, void *stream);
/* update_stats (eval.py:207-215): stats uint64 [3, num_classes] += (ground-truth count, prediction count, agreement).
 * wrap_unvoted = 1: a point without a vote (prediction -1) is counted in the last class of row 1, as numba's
 * wrap-around indexing does in the reference; 0: it is counted nowhere (shapenet eval.py:188-201 tests
 * `predictions == i`; its IoU per class is stats[2] / (stats[0] + stats[1] - stats[2])).
 * num_classes <= 2048.  The caller zeroes `stats` once per scene. */
PVCNN_API int pvcnn_vote_stats(long long n, int num_classes, int wrap_unvoted, const int *gt, const int *pred,
                               unsigned long long *stats, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PVCNN_B200_H_ */
