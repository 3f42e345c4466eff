// mlp_pipeline.cu -- SharedMLP (modules/shared_mlp.py:6-33: Conv1d/Conv2d k=1 + BatchNorm + ReLU) on the tcgen05 path.
//
// One layer = one C-ABI call each way.  Activations are channels-last fp32 [rows, Cp] (rows = B*N for dim=1,
// B*M*U for dim=2; Cp = C padded to 4), so the 1x1 convolution is a plain GEMM on igemm_conv_kernel
// (conv_igemm.cu, ntaps = 1), the BatchNorm statistics / apply / backward passes are the streaming kernels of
// fused_ops.cu (LeakyReLU slope 0 == ReLU), and the weight gradient is conv_wgrad_kernel.  Replaces the cuDNN
// Conv1d/Conv2d + ATen BatchNorm + ReLU launches of the reference's SharedMLP everywhere it is used on its own
// (model heads, cloud MLPs, PointNet++ SA / FP modules: modules/pointnet.py:26,75,100).
//
// The last layer of a set-abstraction MLP can fold `max over the U neighbours` (modules/pointnet.py:87:
// `.max(dim=-1).values`) into its BatchNorm-apply pass: the [rows, C] activation is then never written, only the
// pooled [rows/U, C] tensor and the arg-max rows needed by the backward.
#include "fused_ops.cuh"

namespace pvb {
int igemm_launch(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                 int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                 int npass, cudaStream_t stream);
int igemm_launch_ep(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                    int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                    int npass, cudaStream_t stream, const IgemmEpilogue *ep);
int wgrad_launch(int nb, int sx, int sy, int sz, int cin, int cout, int ntaps, const float *x_hi, const float *x_lo,
                 int ldx, const float *g_hi, const float *g_lo, int ldg, float *dw, int npass, cudaStream_t s,
                 const int4 *ktile_list, const int *ktile_count, int *bz_out, int *by_out);

static inline int mp_pad4(int x) { return (x + 3) / 4 * 4; }
static inline int mp_ld32(int x) { return (x + 31) / 32 * 32; }

// [B*N, Cp] -> [B, C, N]   (inverse of points_to_cl_kernel: 32x32 tiles through shared memory)
__global__ void __launch_bounds__(256) cl_to_points_kernel(int c, int n, int cp, const float *__restrict__ xcl,
                                                           float *__restrict__ x) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = n0 + ty + 8 * j, cc = c0 + tx;
    tile[ty + 8 * j][tx] = (p < n && cc < cp) ? xcl[((size_t)b * n + p) * cp + cc] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c0 + ty + 8 * j, i = n0 + tx;
    if (cc < c && i < n) x[((size_t)b * c + cc) * n + i] = tile[tx][ty + 8 * j];
  }
}

// Channel concatenation written straight into channels-last rows (model-level glue: the 1472-channel `torch.cat` of
// models/s3dis/pvcnn.py:44-46 and its `repeat` of the cloud feature never exist in [B,C,N] form):
//   out[(b, i)][col0 + ch] = x[b][ch][src_n == 1 ? 0 : i]      (+ lo), 32 x 32 tiles through shared memory
__global__ void __launch_bounds__(256) cat_to_cl_kernel(int c, int n, int src_n, int ld, int col0,
                                                        const float *__restrict__ x, float *__restrict__ out,
                                                        float *__restrict__ out_lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c0 + ty + 8 * j, i = n0 + tx;
    tile[ty + 8 * j][tx] = (cc < c && i < n) ? x[((size_t)b * c + cc) * src_n + (src_n == 1 ? 0 : i)] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = n0 + ty + 8 * j, cc = c0 + tx;
    if (p < n && cc < c) {
      const float v = tile[tx][ty + 8 * j];
      const size_t o = ((size_t)b * n + p) * ld + col0 + cc;
      out[o] = v;
      if (out_lo) out_lo[o] = __fsub_rn(v, __uint_as_float(__float_as_uint(v) & 0xFFFFE000u));
    }
  }
}

// inverse: x[b][ch][i] = rows[(b, i)][col0 + ch]
__global__ void __launch_bounds__(256) cl_slice_to_points_kernel(int c, int n, int ld, int col0,
                                                                 const float *__restrict__ rows, float *__restrict__ x) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = n0 + ty + 8 * j, cc = c0 + tx;
    tile[ty + 8 * j][tx] = (p < n && cc < c) ? rows[((size_t)b * n + p) * ld + col0 + cc] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c0 + ty + 8 * j, i = n0 + tx;
    if (cc < c && i < n) x[((size_t)b * c + cc) * n + i] = tile[tx][ty + 8 * j];
  }
}

// column sums of [rows, cp] (bias gradient of a plain linear layer)
__global__ void __launch_bounds__(RED_THREADS) colsum_kernel(long long rows, int cp, const float *__restrict__ g,
                                                             float *__restrict__ partials) {
  column_reduce<1>(rows, cp, partials, [&](long long r, int c4, float4 *acc) {
    const float4 v = ldg_stream4(g + (size_t)r * cp + c4 * 4);
    acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
  });
}

// per-cloud column sums: out[(b * segs + s), cp] = sum of rows [b*n + s*per, min(n, (s+1)*per)) of g  (gradient of a per-cloud
// bias); thread -> (channel quad, row lane) as in column_reduce
__global__ void __launch_bounds__(RED_THREADS) group_colsum_kernel(long long n, int segs, int cp,
                                                                   const float *__restrict__ g, float *__restrict__ out) {
  __shared__ float4 red[RED_THREADS];
  const int cp4 = cp >> 2, rl = RED_THREADS / cp4;
  const int c4 = threadIdx.x % cp4, lane_r = threadIdx.x / cp4;
  const int seg = blockIdx.x, b = blockIdx.y;
  const long long per = (n + segs - 1) / segs;
  const long long r0 = seg * per, r1 = min(n, r0 + per);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane_r < rl)
    for (long long r = r0 + lane_r; r < r1; r += rl) {
      const float4 v = ldg_stream4(g + ((size_t)b * n + r) * cp + c4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < cp4) {
    float4 t = red[threadIdx.x];
    for (int j = 1; j < rl; ++j) {
      const float4 v = red[threadIdx.x + j * cp4];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    st4(out + ((size_t)b * segs + seg) * cp + threadIdx.x * 4, t);
  }
}
__global__ void __launch_bounds__(256) group_colsum_finish_kernel(long long total, int segs, int cp,
                                                                  const float *__restrict__ part, float *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long b = i / cp;
    const int c = (int)(i % cp);
    float t = 0.f;
    for (int k = 0; k < segs; ++k) t += part[((size_t)b * segs + k) * cp + c];
    out[i] = t;
  }
}

// relu(bn(y)) followed by max over groups of U consecutive rows; one CTA per (group, row segment).
// thread -> (channel quad c4 = t % cp4, row lane = t / cp4); partial (max, argmax) per segment, combined below.
constexpr int PL_THREADS = 256;
__global__ void __launch_bounds__(PL_THREADS) bn_relu_pool_kernel(int u, int segs, int cp, const float *__restrict__ y,
                                                                  BnCoef coef, float *__restrict__ pmax /*[G][segs][cp]*/,
                                                                  int *__restrict__ parg) {
  __shared__ float4 smax[PL_THREADS];
  __shared__ int4 sarg[PL_THREADS];
  const int cp4 = cp >> 2, rl = PL_THREADS / cp4;
  const int c4 = threadIdx.x % cp4, lane_r = threadIdx.x / cp4;
  const int g = blockIdx.x, seg = blockIdx.y;
  const int per = (u + segs - 1) / segs;
  const int r0 = seg * per, r1 = min(u, r0 + per);
  float4 best = make_float4(-1.f, -1.f, -1.f, -1.f);
  int4 arg = make_int4(r0, r0, r0, r0);
  if (lane_r < rl) {
    const float4 sc = *reinterpret_cast<const float4 *>(coef.scale + c4 * 4);
    const float4 sh = *reinterpret_cast<const float4 *>(coef.shift + c4 * 4);
    for (int r = r0 + lane_r; r < r1; r += rl) {
      const float4 v = ldg_stream4(y + ((size_t)g * u + r) * cp + c4 * 4);
      const float zx = fmaxf(fmaf(v.x, sc.x, sh.x), 0.f), zy = fmaxf(fmaf(v.y, sc.y, sh.y), 0.f);
      const float zz = fmaxf(fmaf(v.z, sc.z, sh.z), 0.f), zw = fmaxf(fmaf(v.w, sc.w, sh.w), 0.f);
      if (zx > best.x) { best.x = zx; arg.x = r; }
      if (zy > best.y) { best.y = zy; arg.y = r; }
      if (zz > best.z) { best.z = zz; arg.z = r; }
      if (zw > best.w) { best.w = zw; arg.w = r; }
    }
  }
  smax[threadIdx.x] = best;
  sarg[threadIdx.x] = arg;
  __syncthreads();
  if (threadIdx.x < cp4) {
    float4 m = smax[threadIdx.x];
    int4 a = sarg[threadIdx.x];
    for (int j = 1; j < rl; ++j) {  // ties: the smallest row index wins (rows of a lane ascend, lanes ascend)
      const float4 v = smax[threadIdx.x + j * cp4];
      const int4 w = sarg[threadIdx.x + j * cp4];
      if (v.x > m.x || (v.x == m.x && w.x < a.x)) { m.x = v.x; a.x = w.x; }
      if (v.y > m.y || (v.y == m.y && w.y < a.y)) { m.y = v.y; a.y = w.y; }
      if (v.z > m.z || (v.z == m.z && w.z < a.z)) { m.z = v.z; a.z = w.z; }
      if (v.w > m.w || (v.w == m.w && w.w < a.w)) { m.w = v.w; a.w = w.w; }
    }
    const size_t o = ((size_t)g * segs + seg) * cp + threadIdx.x * 4;
    *reinterpret_cast<float4 *>(pmax + o) = m;
    *reinterpret_cast<int4 *>(parg + o) = a;
  }
}

__global__ void __launch_bounds__(256) pool_combine_kernel(long long total, int segs, int cp,
                                                           const float *__restrict__ pmax, const int *__restrict__ parg,
                                                           float *__restrict__ pooled, int *__restrict__ argmax) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const long long g = t / cp;
    const int c = (int)(t - g * cp);
    float m = -1.f;
    int a = 0;
    for (int s = 0; s < segs; ++s) {
      const float v = pmax[((size_t)g * segs + s) * cp + c];
      if (v > m) { m = v; a = parg[((size_t)g * segs + s) * cp + c]; }
    }
    pooled[t] = fmaxf(m, 0.f);
    argmax[t] = a;
  }
}

// gz[g*U + r][c] = (r == argmax[g][c]) ? gpool[g][c] : 0   (dense, so the ordinary BatchNorm backward can follow)
__global__ void __launch_bounds__(256) pool_scatter_kernel(long long total4, int u, int cp,
                                                           const float *__restrict__ gpool, const int *__restrict__ argmax,
                                                           float *__restrict__ gz) {
  const int cp4 = cp >> 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / cp4;
    const int c4 = (int)(t - row * cp4);
    const long long g = row / u;
    const int r = (int)(row - g * u);
    const int4 a = *reinterpret_cast<const int4 *>(argmax + (size_t)g * cp + c4 * 4);
    const float4 v = *reinterpret_cast<const float4 *>(gpool + (size_t)g * cp + c4 * 4);
    float4 o;
    o.x = a.x == r ? v.x : 0.f; o.y = a.y == r ? v.y : 0.f; o.z = a.z == r ? v.z : 0.f; o.w = a.w == r ? v.w : 0.f;
    stg_stream4(gz + t * 4, o);
  }
}

// BallQuery grouping + centre subtraction + concat (modules/ball_query.py:16-30) written straight into channels-last
// rows:  out[(b,m,u)][0:3] = coords[b,:,idx] - centers[b,:,m],  out[..][3:3+C] = features[b,:,idx],  pad -> 0  (+ lo).
// 32 rows x 32 channels per CTA: gather with lanes along the rows, transposed through shared memory, stored with lanes
// along the channels (128-byte rows).  The [B,3+C,M,U] tensor of the reference is never materialised.
__global__ void __launch_bounds__(256) group_concat_cl_kernel(int c, int n, int m, int u, int cp,
                                                              const float *__restrict__ coords,
                                                              const float *__restrict__ centers,
                                                              const float *__restrict__ feat, const int *__restrict__ idx,
                                                              float *__restrict__ out, float *__restrict__ out_lo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, e0 = blockIdx.x * 32, mu = m * u, ct = c + 3;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int e = e0 + tx;
  const int src = e < mu ? __ldg(idx + (size_t)b * mu + e) : 0;
  const int ctr = e < mu ? e / u : 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = c0 + ty + 8 * j;
    float v = 0.0f;
    if (e < mu && ch < ct) {
      if (ch < 3)
        v = __fsub_rn(__ldg(coords + ((size_t)b * 3 + ch) * n + src), __ldg(centers + ((size_t)b * 3 + ch) * m + ctr));
      else
        v = __ldg(feat + ((size_t)b * c + (ch - 3)) * n + src);
    }
    tile[ty + 8 * j][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = e0 + ty + 8 * j, ch = c0 + tx;
    if (row < mu && ch < cp) {
      const float v = tile[tx][ty + 8 * j];
      const size_t o = ((size_t)b * mu + row) * cp + ch;
      out[o] = v;
      if (out_lo) out_lo[o] = __fsub_rn(v, __uint_as_float(__float_as_uint(v) & 0xFFFFE000u));
    }
  }
}

// gradient of the above: scatter-add of the channels-last row gradients into [B,C,N] features / [B,3,N] coords,
// minus the per-centre sums into [B,3,M] (all pre-zeroed)
__global__ void __launch_bounds__(256) group_concat_cl_grad_kernel(int c, int n, int m, int u, int cp,
                                                                   const float *__restrict__ g,
                                                                   const int *__restrict__ idx,
                                                                   float *__restrict__ grad_feat,
                                                                   float *__restrict__ grad_coords,
                                                                   float *__restrict__ grad_centers) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, e0 = blockIdx.x * 32, mu = m * u, ct = c + 3;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int row = e0 + ty + 8 * j, ch = c0 + tx;
    tile[ty + 8 * j][tx] = (row < mu && ch < ct) ? g[((size_t)b * mu + row) * cp + ch] : 0.0f;
  }
  __syncthreads();
  const int e = e0 + tx;
  if (e >= mu) return;
  const int dst = __ldg(idx + (size_t)b * mu + e);
  const int ctr = e / u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int ch = c0 + ty + 8 * j;
    if (ch >= ct) continue;
    const float v = tile[tx][ty + 8 * j];
    if (ch < 3) {
      if (grad_coords) atomicAdd(grad_coords + ((size_t)b * 3 + ch) * n + dst, v);
      if (grad_centers) atomicAdd(grad_centers + ((size_t)b * 3 + ch) * m + ctr, -v);
    } else if (grad_feat) {
      atomicAdd(grad_feat + ((size_t)b * c + (ch - 3)) * n + dst, v);
    }
  }
}

static BnCoef mlp_coef(float *base, int cp) { return BnCoef{base, base + cp, base + 2 * cp, base + 3 * cp}; }
}  // namespace pvb

using namespace pvb;

extern "C" int pvcnn_conv_weight_prep(int cout, int cin, int ntaps, int mode, int ld, const float *w, float *w_hi,
                                      float *w_lo, void *stream);

#define MLP_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

extern "C" {

int pvcnn_points_to_cl(int b, int c, int n, const float *x, float *xcl, float *xcl_lo, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && x && xcl);
  return launch_points_to_cl(b, c, n, mp_pad4(c), x, xcl, xcl_lo, (cudaStream_t)stream);
}

int pvcnn_cl_to_points(int b, int c, int n, const float *xcl, float *x, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && x && xcl);
  const int cp = mp_pad4(c);
  PVB_LAUNCH(cl_to_points_kernel, dim3(ceil_div(n, 32), ceil_div(c, 32), b), 256, 0, stream, c, n, cp, xcl, x);
  return 0;
}

/* modules/ball_query.py:16-30 fused, channels-last output [b*m*u, pad4(3+c)] (+ lo = x - trunc_tf32(x), may be NULL);
 * features may be NULL (c == 0). */
int pvcnn_group_concat_cl(int b, int c, int n, int m, int u, const float *points_coords, const float *centers_coords,
                          const float *features, const int *indices, float *out, float *out_lo, void *stream) {
  PVB_CHECK_ARG(b > 0 && c >= 0 && n > 0 && m > 0 && u > 0 && points_coords && centers_coords && indices && out);
  PVB_CHECK_ARG(c == 0 || features != nullptr);
  const int cp = mp_pad4(c + 3);
  PVB_LAUNCH(group_concat_cl_kernel, dim3(ceil_div((long long)m * u, 32), ceil_div(cp, 32), b), 256, 0, stream, c, n, m, u,
             cp, points_coords, centers_coords, features, indices, out, out_lo);
  return 0;
}

int pvcnn_group_concat_cl_grad(int b, int c, int n, int m, int u, const float *grad_rows, const int *indices,
                               float *grad_features, float *grad_points_coords, float *grad_centers_coords,
                               void *stream) {
  PVB_CHECK_ARG(b > 0 && c >= 0 && n > 0 && m > 0 && u > 0 && grad_rows && indices);
  cudaStream_t s = (cudaStream_t)stream;
  if (grad_features) PVB_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)b * c * n, s));
  if (grad_points_coords) PVB_CUDA(cudaMemsetAsync(grad_points_coords, 0, sizeof(float) * (size_t)b * 3 * n, s));
  if (grad_centers_coords) PVB_CUDA(cudaMemsetAsync(grad_centers_coords, 0, sizeof(float) * (size_t)b * 3 * m, s));
  const int cp = mp_pad4(c + 3);
  PVB_LAUNCH(group_concat_cl_grad_kernel, dim3(ceil_div((long long)m * u, 32), ceil_div(c + 3, 32), b), 256, 0, s, c, n, m,
             u, cp, grad_rows, indices, grad_features, grad_points_coords, grad_centers_coords);
  return 0;
}

/* channel concatenation in channels-last form: source x [b,c,n] (or [b,c,1]: broadcast over the points, src_n = 1) ->
 * columns [col0, col0+c) of rows [b*n, ld] (+ lo rows, may be NULL).  Pad columns of the destination are the caller's
 * to zero. */
int pvcnn_cat_to_cl(int b, int c, int n, int src_n, const float *x, int ld, int col0, float *rows, float *rows_lo,
                    void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && (src_n == n || src_n == 1) && x && rows && ld >= col0 + c && col0 >= 0);
  PVB_LAUNCH(cat_to_cl_kernel, dim3(ceil_div(n, 32), ceil_div(c, 32), b), 256, 0, stream, c, n, src_n, ld, col0, x, rows,
             rows_lo);
  return 0;
}

int pvcnn_cl_slice_to_points(int b, int c, int n, const float *rows, int ld, int col0, float *x, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && x && rows && ld >= col0 + c && col0 >= 0);
  PVB_LAUNCH(cl_slice_to_points_kernel, dim3(ceil_div(n, 32), ceil_div(c, 32), b), 256, 0, stream, c, n, ld, col0, rows, x);
  return 0;
}

/* plain 1x1 convolution (no BatchNorm / ReLU: the classifier's last layer, models/utils.py:43) on channels-last rows */
int pvcnn_linear_cl_forward(long long rows, int cin, int cout, int npass, const float *x, const float *x_lo,
                            const float *w, const float *bias, float *wprep, float *y, void *stream) {
  PVB_CHECK_ARG(rows > 0 && rows < (1LL << 31) && cin > 0 && cout > 0 && (npass == 1 || npass == 3) && x && w && wprep && y);
  PVB_CHECK_ARG(npass == 1 || x_lo);
  const int ci = mp_pad4(cin), co = mp_pad4(cout);
  const long long nf = (long long)cout * mp_ld32(cin);
  MLP_TRY(pvcnn_conv_weight_prep(cout, cin, 1, 0, mp_ld32(cin), w, wprep, wprep + nf, stream));
  if (co != cout) MLP_TRY(launch_memset_f32(y, rows * co, (cudaStream_t)stream));
  return igemm_launch(1, 1, 1, (int)rows, cin, cout, 1, x, x_lo, ci, wprep, wprep + nf, mp_ld32(cin), bias, y, co, npass,
                      (cudaStream_t)stream);
}

int pvcnn_linear_cl_backward(long long rows, int cin, int cout, int npass, const float *gy, const float *gy_lo,
                             const float *x, const float *x_lo, const float *w, float *wprep, float *partials,
                             float *gx, float *dw, float *dbias, void *stream) {
  PVB_CHECK_ARG(rows > 0 && rows < (1LL << 31) && cin > 0 && cout > 0 && (npass == 1 || npass == 3));
  PVB_CHECK_ARG(gy && x && w && wprep && partials && dw && (npass == 1 || (gy_lo && x_lo)));
  cudaStream_t s = (cudaStream_t)stream;
  const int ci = mp_pad4(cin), co = mp_pad4(cout);
  PVB_CHECK_ARG(co / 4 <= RED_THREADS);
  if (dbias) {
    const int rl = RED_THREADS / (co / 4);
    long long g = (rows + rl * 8 - 1) / (rl * 8);
    if (g > RED_MAX_BLOCKS) g = RED_MAX_BLOCKS;
    PVB_LAUNCH(colsum_kernel, (int)g, RED_THREADS, 0, s, rows, co, gy, partials);
    MLP_TRY(launch_reduce_partials((int)g, co, partials, partials + (size_t)g * co, s));
    PVB_CUDA(cudaMemcpyAsync(dbias, partials + (size_t)g * co, sizeof(float) * (size_t)cout, cudaMemcpyDeviceToDevice, s));
  }
  MLP_TRY(wgrad_launch(1, 1, 1, (int)rows, cin, cout, 1, x, x_lo, ci, gy, gy_lo, co, dw, npass, s, nullptr, nullptr, nullptr, nullptr));
  if (gx) {
    const long long nd = (long long)cin * mp_ld32(cout);
    MLP_TRY(pvcnn_conv_weight_prep(cout, cin, 1, 1, mp_ld32(cout), w, wprep, wprep + nd, stream));
    if (ci != cin) MLP_TRY(launch_memset_f32(gx, rows * ci, s));
    MLP_TRY(igemm_launch(1, 1, 1, (int)rows, cout, cin, 1, gy, gy_lo, co, wprep, wprep + nd, mp_ld32(cout), nullptr, gx, ci, npass, s));
  }
  return 0;
}

long long pvcnn_mlp_partials_floats(int cout) { return (long long)kNumSMs * 4 * 2 * mp_pad4(cout) + 64; }
long long pvcnn_mlp_wprep_floats(int cin, int cout) {
  const long long f = (long long)cout * mp_ld32(cin), d = (long long)cin * mp_ld32(cout);
  return 2 * (f > d ? f : d);
}
int pvcnn_mlp_pool_segments(long long groups, int u) {
  if (u <= 512) return 1;
  long long s = (2LL * kNumSMs + groups - 1) / groups;  // enough CTAs to cover the chip even with B groups
  if (s > (u + 63) / 64) s = (u + 63) / 64;
  if (s > 64) s = 64;
  return s < 1 ? 1 : (int)s;
}

/* One SharedMLP layer forward on channels-last rows.
 *   x [rows, pad4(cin)] (+ x_lo for npass == 3), w [cout, cin] (Conv1d/Conv2d k=1 weight), bias/gamma/beta [cout]
 *   y [rows, pad4(cout)]  pre-BatchNorm output (saved for the backward), coef [4 * pad4(cout)]
 *   pool_u == 0: z (+ z_lo) [rows, pad4(cout)] = relu(bn(y))
 *   pool_u  > 0: pooled [rows / pool_u, pad4(cout)] = max over groups of pool_u rows, argmax likewise (int32 row in
 *                the group); pool_tmp holds segs * 2 * (rows / pool_u) * pad4(cout) floats when segs > 1
 *   training: batch statistics (+ running-stat update); else running statistics                                      */
int pvcnn_mlp_layer_forward(long long rows, int cin, int cout, int training, int npass, float bn_eps, float momentum,
                            const float *x, const float *x_lo, const float *w, const float *bias, const float *gamma,
                            const float *beta, float *running_mean, float *running_var, long long *num_batches_tracked,
                            float *wprep, float *partials, float *coef, float *y, float *z, float *z_lo, int pool_u,
                            float *pooled, int *argmax, float *pool_tmp, long long group_rows, const float *group_bias,
                            int group_ld, void *stream) {
  PVB_CHECK_ARG(rows > 0 && rows < (1LL << 31) && cin > 0 && cout > 0 && (npass == 1 || npass == 3));
  PVB_CHECK_ARG(x && w && gamma && beta && wprep && partials && coef && y && (npass == 1 || x_lo));
  PVB_CHECK_ARG(group_bias == nullptr || (group_rows > 0 && group_rows < (1LL << 31) && rows % group_rows == 0));
  PVB_CHECK_ARG(pool_u > 0 ? (pooled && argmax && rows % pool_u == 0) : (z != nullptr));
  cudaStream_t s = (cudaStream_t)stream;
  const int ci = mp_pad4(cin), co = mp_pad4(cout);
  PVB_CHECK_ARG(co / 4 <= 256);
  const long long nf = (long long)cout * mp_ld32(cin);
  MLP_TRY(pvcnn_conv_weight_prep(cout, cin, 1, 0, mp_ld32(cin), w, wprep, wprep + nf, stream));
  if (co != cout) MLP_TRY(launch_memset_f32(y, rows * co, s));  // pad columns feed the BN passes: keep them finite
  IgemmEpilogue ep;
  ep.group_bias = group_bias;
  ep.group_rows = (int)group_rows;
  ep.group_ld = group_ld;
  MLP_TRY(igemm_launch_ep(1, 1, 1, (int)rows, cin, cout, 1, x, x_lo, ci, wprep, wprep + nf, mp_ld32(cin), bias, y, co, npass,
                          s, group_bias ? &ep : nullptr));
  BnCoef bn = mlp_coef(coef, co);
  if (training) {
    int nblk = 0;
    MLP_TRY(launch_bn_stats(rows, co, y, partials, &nblk, s));
    MLP_TRY(launch_bn_finalize(nblk, cout, co, rows, bn_eps, momentum, partials, gamma, beta, running_mean, running_var, bn, s, num_batches_tracked));
  } else {
    PVB_CHECK_ARG(running_mean && running_var);
    MLP_TRY(launch_bn_coef_from_running(cout, bn_eps, gamma, beta, running_mean, running_var, bn, s));
  }
  if (pool_u == 0) {
    MLP_TRY(launch_bn_apply_leaky(rows, co, 0.0f, y, bn, z, z_lo, s));
  } else {
    const long long groups = rows / pool_u;
    const int segs = pvcnn_mlp_pool_segments(groups, pool_u);
    PVB_CHECK_ARG(segs == 1 || pool_tmp);
    float *pm = segs == 1 ? pooled : pool_tmp;
    int *pa = segs == 1 ? argmax : reinterpret_cast<int *>(pool_tmp + (size_t)groups * segs * co);
    PVB_LAUNCH(bn_relu_pool_kernel, dim3((unsigned)groups, segs), PL_THREADS, 0, s, pool_u, segs, co, y, bn, pm, pa);
    if (segs > 1) {
      const long long total = groups * co;
      long long grid = (total + 255) / 256;
      if (grid > kNumSMs * 8) grid = kNumSMs * 8;
      PVB_LAUNCH(pool_combine_kernel, (int)grid, 256, 0, s, total, segs, co, pm, pa, pooled, argmax);
    }
  }
  return 0;
}

/* Inference form of a SharedMLP layer (eval-mode BatchNorm: running statistics), in two calls.
 *
 * pvcnn_mlp_layer_prepare: everything that depends only on the parameters -- the hi/lo GEMM operand of the weight
 *   (wprep, pvcnn_mlp_wprep_floats(cin, cout) floats) and the BatchNorm coefficients (coef, 4 * pad4(cout) floats:
 *   mean, invstd, scale = gamma * invstd, shift = beta - mean * scale).  The caller keeps both for as long as the
 *   parameters do not change, so a forward pass of a frozen network launches neither.
 *
 * pvcnn_mlp_layer_forward_eval: z = relu(bn(conv(x))) as ONE GEMM whose epilogue applies bias, BatchNorm and ReLU and
 *   writes z (+ z_lo, the lo operand of the next 3xTF32 layer; NULL when not needed): the pre-activation tensor y is
 *   never written (the two-kernel form writes y, reads it back, writes z and z_lo).  Same arithmetic as
 *   pvcnn_mlp_layer_forward(training = 0): y = acc + bias in fp32, z = max(fma(y, scale, shift), 0).
 *   group_bias [rows / group_rows, group_ld] (or NULL) is added to y before the BatchNorm: the contribution of input
 *   channels that are constant over each cloud of group_rows consecutive rows (a max-pooled cloud feature or a one-hot
 *   class vector repeated over the points, models/shapenet/pvcnn.py:40-42), computed by the caller as a
 *   [clouds, cout] GEMM instead of being concatenated to every row.
 *   pool_u > 0: as in pvcnn_mlp_layer_forward (y is then needed as scratch: [rows, pad4(cout)]). */
int pvcnn_mlp_layer_prepare(int cin, int cout, float bn_eps, const float *w, const float *gamma, const float *beta,
                            const float *running_mean, const float *running_var, float *wprep, float *coef,
                            void *stream) {
  PVB_CHECK_ARG(cin > 0 && cout > 0 && w && gamma && beta && running_mean && running_var && wprep && coef);
  const long long nf = (long long)cout * mp_ld32(cin);
  MLP_TRY(pvcnn_conv_weight_prep(cout, cin, 1, 0, mp_ld32(cin), w, wprep, wprep + nf, stream));
  return launch_bn_coef_from_running(cout, bn_eps, gamma, beta, running_mean, running_var, mlp_coef(coef, mp_pad4(cout)),
                                     (cudaStream_t)stream);
}

int pvcnn_mlp_layer_forward_eval(long long rows, int cin, int cout, int npass, const float *x, const float *x_lo,
                                 const float *wprep, const float *bias, const float *coef, long long group_rows,
                                 const float *group_bias, int group_ld, float *y, float *z, float *z_lo, int pool_u,
                                 float *pooled, int *argmax, float *pool_tmp, void *stream) {
  PVB_CHECK_ARG(rows > 0 && rows < (1LL << 31) && cin > 0 && cout > 0 && (npass == 1 || npass == 3));
  PVB_CHECK_ARG(x && wprep && coef && (npass == 1 || x_lo));
  PVB_CHECK_ARG(pool_u > 0 ? (pooled && argmax && y && rows % pool_u == 0) : (z != nullptr));
  PVB_CHECK_ARG(group_bias == nullptr || (group_rows > 0 && group_rows < (1LL << 31) && rows % group_rows == 0));
  cudaStream_t s = (cudaStream_t)stream;
  const int ci = mp_pad4(cin), co = mp_pad4(cout);
  PVB_CHECK_ARG(co / 4 <= 256);
  const long long nf = (long long)cout * mp_ld32(cin);
  BnCoef bn = mlp_coef(const_cast<float *>(coef), co);
  IgemmEpilogue ep;
  ep.group_bias = group_bias;
  ep.group_rows = (int)group_rows;
  ep.group_ld = group_ld;
  if (pool_u == 0) {
    if (co != cout) {   // pad columns are K columns of the next layer: keep them finite
      MLP_TRY(launch_memset_f32(z, rows * co, s));
      if (z_lo) MLP_TRY(launch_memset_f32(z_lo, rows * co, s));
    }
    ep.scale = bn.scale;
    ep.shift = bn.shift;
    ep.slope = 0.0f;
    ep.out_lo = z_lo;
    return igemm_launch_ep(1, 1, 1, (int)rows, cin, cout, 1, x, x_lo, ci, wprep, wprep + nf, mp_ld32(cin), bias, z, co,
                           npass, s, &ep);
  }
  if (co != cout) MLP_TRY(launch_memset_f32(y, rows * co, s));
  MLP_TRY(igemm_launch_ep(1, 1, 1, (int)rows, cin, cout, 1, x, x_lo, ci, wprep, wprep + nf, mp_ld32(cin), bias, y, co,
                          npass, s, group_bias ? &ep : nullptr));
  const long long groups = rows / pool_u;
  const int segs = pvcnn_mlp_pool_segments(groups, pool_u);
  PVB_CHECK_ARG(segs == 1 || pool_tmp);
  float *pm = segs == 1 ? pooled : pool_tmp;
  int *pa = segs == 1 ? argmax : reinterpret_cast<int *>(pool_tmp + (size_t)groups * segs * co);
  PVB_LAUNCH(bn_relu_pool_kernel, dim3((unsigned)groups, segs), PL_THREADS, 0, s, pool_u, segs, co, y, bn, pm, pa);
  if (segs > 1) {
    const long long total = groups * co;
    long long grid = (total + 255) / 256;
    if (grid > kNumSMs * 8) grid = kNumSMs * 8;
    PVB_LAUNCH(pool_combine_kernel, (int)grid, 256, 0, s, total, segs, co, pm, pa, pooled, argmax);
  }
  return 0;
}

/* dense gradient of the pooled output: gz [groups * u, cp] from gpool / argmax [groups, cp] */
int pvcnn_mlp_pool_backward(long long groups, int u, int cout, const float *gpool, const int *argmax, float *gz, void *stream) {
  PVB_CHECK_ARG(groups > 0 && u > 0 && cout > 0 && gpool && argmax && gz);
  const int co = mp_pad4(cout);
  const long long total4 = groups * u * (co / 4);
  long long grid = (total4 + 1023) / 1024;
  if (grid > kNumSMs * 8) grid = kNumSMs * 8;
  PVB_LAUNCH(pool_scatter_kernel, (int)grid, 256, 0, stream, total4, u, co, gpool, argmax, gz);
  return 0;
}

/* One SharedMLP layer backward.  gz [rows, pad4(cout)] = gradient of relu(bn(y)); produces
 *   dgamma, dbeta, dbias [cout], dw [cout, cin], gx [rows, pad4(cin)] (may be NULL: first layer of a net whose input needs
 *   no gradient), using gy (+ gy_lo) [rows, pad4(cout)] as scratch for the conv-output gradient; sums: 4 * pad4(cout). */
int pvcnn_mlp_layer_backward(long long rows, int cin, int cout, int npass, const float *gz, const float *x,
                             const float *x_lo, const float *w, const float *y, const float *coef, float *wprep,
                             float *partials, float *sums, float *gy, float *gy_lo, float *gx, float *dw, float *dbias,
                             float *dgamma, float *dbeta, long long group_rows, float *d_group_bias, void *stream) {
  PVB_CHECK_ARG(rows > 0 && rows < (1LL << 31) && cin > 0 && cout > 0 && (npass == 1 || npass == 3));
  PVB_CHECK_ARG(d_group_bias == nullptr || (group_rows > 0 && rows % group_rows == 0));
  PVB_CHECK_ARG(gz && x && w && y && coef && wprep && partials && sums && gy && dw && dbias && dgamma && dbeta);
  PVB_CHECK_ARG(npass == 1 || (x_lo && gy_lo));
  cudaStream_t s = (cudaStream_t)stream;
  const int ci = mp_pad4(cin), co = mp_pad4(cout);
  BnCoef bn = mlp_coef(const_cast<float *>(coef), co);
  int nblk = 0;
  const size_t cb = sizeof(float) * (size_t)cout;
  // ReLU mask + BatchNorm reductions: S1 = sum g', S2 = sum g' * xhat  (= dbeta, dgamma)
  MLP_TRY(launch_bn_bwd_reduce(rows, co, 0.0f, gz, y, bn, partials, &nblk, s));
  MLP_TRY(launch_reduce_partials(nblk, 2 * co, partials, sums, s));
  PVB_CUDA(cudaMemcpyAsync(dbeta, sums, cb, cudaMemcpyDeviceToDevice, s));
  PVB_CUDA(cudaMemcpyAsync(dgamma, sums + co, cb, cudaMemcpyDeviceToDevice, s));
  // conv-output gradient (+ its column sums = conv-bias gradient)
  MLP_TRY(launch_bn_bwd_apply(rows, co, 1, 0.0f, gz, y, bn, sums, sums + co, gy, npass > 1 ? gy_lo : nullptr, partials, &nblk, s));
  MLP_TRY(launch_reduce_partials(nblk, co, partials, sums + 2 * co, s));
  PVB_CUDA(cudaMemcpyAsync(dbias, sums + 2 * co, cb, cudaMemcpyDeviceToDevice, s));
  if (d_group_bias) {   // gradient of the per-cloud bias [rows / group_rows, co]: column sums of gy over each cloud
    const long long groups = rows / group_rows;
    PVB_CHECK_ARG(groups < 65536 && co / 4 <= RED_THREADS);
    const long long cap = (pvcnn_mlp_partials_floats(cout) - 64) / co;   // rows of `partials` available
    long long segs = (group_rows + 511) / 512;
    if (segs * groups > cap) segs = cap / groups;
    if (segs <= 1) {
      PVB_LAUNCH(group_colsum_kernel, dim3(1, (unsigned)groups), RED_THREADS, 0, s, group_rows, 1, co, gy, d_group_bias);
    } else {
      PVB_LAUNCH(group_colsum_kernel, dim3((unsigned)segs, (unsigned)groups), RED_THREADS, 0, s, group_rows, (int)segs, co,
                 gy, partials);
      const long long total = groups * co;
      PVB_LAUNCH(group_colsum_finish_kernel, (int)min((total + 255) / 256, (long long)kNumSMs * 4), 256, 0, s, total,
                 (int)segs, co, partials, d_group_bias);
    }
  }
  // weight gradient, then the data gradient
  MLP_TRY(wgrad_launch(1, 1, 1, (int)rows, cin, cout, 1, x, x_lo, ci, gy, gy_lo, co, dw, npass, s, nullptr, nullptr, nullptr, nullptr));
  if (gx) {
    const long long nd = (long long)cin * mp_ld32(cout);
    MLP_TRY(pvcnn_conv_weight_prep(cout, cin, 1, 1, mp_ld32(cout), w, wprep, wprep + nd, stream));
    if (ci != cin) MLP_TRY(launch_memset_f32(gx, rows * ci, s));
    MLP_TRY(igemm_launch(1, 1, 1, (int)rows, cout, cin, 1, gy, gy_lo, co, wprep, wprep + nd, mp_ld32(cout), nullptr, gx, ci, npass, s));
  }
  return 0;
}

}  // extern "C"
