// fused_ops.cu -- channels-last HBM-bound kernels of the fused PVConv pipeline (see fused_ops.cuh).
//
// These replace, in one pass each, chains of separate kernels in the reference:
//   voxelize_cl           <- avg_voxelize_kernel (vox.cu:48-72) but channels-last, warp-aggregated
//   bn_stats/bn_apply     <- nn.BatchNorm3d + nn.LeakyReLU (modules/pvconv.py:22-26)
//   devox_fused           <- BatchNorm3d-apply + LeakyReLU + trilinear_devoxelize_kernel
//                            (trilinear_devox.cu:21-105) + BatchNorm1d-apply + ReLU + add
//                            (modules/pvconv.py:36-38, modules/shared_mlp.py:11-12)
//   bwd_points            <- add-backward, ReLU/BN1d backward reductions, trilinear_devoxelize_grad_kernel
//                            (trilinear_devox.cu:119-162), LeakyReLU backward, BN3d backward reductions
//   bn_bwd_apply          <- BatchNorm backward (input gradient) + conv-bias gradient
//   bwd_final             <- avg_voxelize_grad_kernel (vox.cu:86-110) + add of the point-branch gradient
// All of them stream rows of C contiguous floats with 128-bit accesses.
#include "fused_ops.cuh"

namespace pvb {

__device__ __forceinline__ float tf32_lo(float x) {
  const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __fsub_rn(x, hi);  // <= 13 significant bits; the tensor core truncates it to tf32 itself
}
__device__ __forceinline__ float4 tf32_lo4(float4 v) {
  return make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
}
__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.0f ? z : z * slope; }

#define PVB_TRY_LAUNCH(expr)     \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

int launch_reduce_partials(int nblocks, int ncols, const float *partials, float *sums, cudaStream_t s);

static inline int grid_for(long long work_items, int per_block, int max_blocks) {
  long long g = (work_items + per_block - 1) / per_block;
  if (g > max_blocks) g = max_blocks;
  if (g < 1) g = 1;
  return (int)g;
}

// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) points_to_cl_kernel(int c, int n, int cp, const float *__restrict__ x,
                                                           float *__restrict__ xcl, float *__restrict__ xlo) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.y * 32, n0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int cc = c0 + ty + 8 * j, i = n0 + tx;
    tile[ty + 8 * j][tx] = (cc < c && i < n) ? x[((size_t)b * c + cc) * n + i] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int p = n0 + ty + 8 * j, cc = c0 + tx;
    if (p < n && cc < cp) {
      const float v = tile[tx][ty + 8 * j];
      const size_t o = ((size_t)b * n + p) * cp + cc;
      xcl[o] = v;
      if (xlo) xlo[o] = tf32_lo(v);
    }
  }
}

int launch_points_to_cl(int b, int c, int n, int cp, const float *x, float *xcl, float *xcl_lo, cudaStream_t s) {
  PVB_LAUNCH(points_to_cl_kernel, dim3(ceil_div(n, 32), ceil_div(cp, 32), b), 256, 0, s, c, n, cp, x, xcl, xcl_lo);
  return 0;
}

// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) vox_index_count_cl_kernel(int n, int r, int r3, long long total,
                                                                 const int *__restrict__ coords,
                                                                 int *__restrict__ ind, int *__restrict__ cnt) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / n), i = (int)(t % n);
    const int *co = coords + (size_t)b * 3 * n;
    const int v = co[i] * r * r + co[i + n] * r + co[i + 2 * n];  // vox.cu:31
    ind[t] = v;
    atomicAdd(cnt + (size_t)b * r3 + v, 1);
  }
}

int launch_vox_index_count(int b, int n, int r, const int *coords, int *ind, int *cnt, cudaStream_t s) {
  const int r3 = r * r * r;
  PVB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)b * r3, s));
  const long long total = (long long)b * n;
  PVB_LAUNCH(vox_index_count_cl_kernel, grid_for(total, 256, kNumSMs * 8), 256, 0, s, n, r, r3, total, coords, ind,
             cnt);
  return 0;
}

// One warp owns 32 consecutive points.  Points of the warp that fall into the same voxel form a
// group (match.any); the warp walks the groups, sums each group's rows in registers (coalesced
// 16-byte loads, ascending point order) and issues ONE vector reduction per (voxel, 4 channels).
__global__ void __launch_bounds__(256) voxelize_cl_kernel(int n, int r3, int cp, const int *__restrict__ ind,
                                                          const int *__restrict__ cnt,
                                                          const float *__restrict__ xcl, float *__restrict__ grid) {
  const int b = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = blockIdx.x * 256 + warp * 32;
  if (base >= n) return;
  const int i = base + lane;
  const bool valid = i < n;
  const int pos = valid ? ind[(size_t)b * n + i] : -1 - lane;
  const unsigned grp = __match_any_sync(0xffffffffu, pos);
  const bool leader = valid && ((__ffs(grp) - 1) == lane);
  float inv = 0.0f;
  if (valid) inv = (float)(1.0 / (double)(float)cnt[(size_t)b * r3 + pos]);  // vox.cu:66
  const unsigned leaders = __ballot_sync(0xffffffffu, leader);
  const int cp4 = cp >> 2;
  const float *rows = xcl + ((size_t)b * n + base) * cp;
  for (unsigned lm = leaders; lm; lm &= lm - 1) {
    const int L = __ffs(lm) - 1;
    const unsigned gmask = __shfl_sync(0xffffffffu, grp, L);
    const int gpos = __shfl_sync(0xffffffffu, pos, L);
    const float ginv = __shfl_sync(0xffffffffu, inv, L);
    float *dst = grid + ((size_t)b * r3 + gpos) * cp;
    for (int c4 = lane; c4 < cp4; c4 += 32) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (unsigned mm = gmask; mm; mm &= mm - 1) {
        const int src = __ffs(mm) - 1;
        const float4 v = ld4(rows + (size_t)src * cp + c4 * 4);
        acc.x = __fadd_rn(acc.x, __fmul_rn(v.x, ginv));
        acc.y = __fadd_rn(acc.y, __fmul_rn(v.y, ginv));
        acc.z = __fadd_rn(acc.z, __fmul_rn(v.z, ginv));
        acc.w = __fadd_rn(acc.w, __fmul_rn(v.w, ginv));
      }
      red_add4(dst + c4 * 4, acc);
    }
  }
}

int launch_voxelize_cl(int b, int n, int r3, int cp, const int *ind, const int *cnt, const float *xcl, float *grid,
                       cudaStream_t s) {
  PVB_LAUNCH(voxelize_cl_kernel, dim3(ceil_div(n, 256), b), 256, 0, s, n, r3, cp, ind, cnt, xcl, grid);
  return 0;
}

__global__ void __launch_bounds__(256) grid_lo_at_points_kernel(int n, int r3, int cp, long long total,
                                                                const int *__restrict__ ind,
                                                                const float *__restrict__ grid,
                                                                float *__restrict__ grid_lo) {
  const int cp4 = cp >> 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const long long p = t / cp4;
    const int c4 = (int)(t % cp4);
    const int b = (int)(p / n);
    const size_t o = ((size_t)b * r3 + ind[p]) * cp + c4 * 4;
    st4(grid_lo + o, tf32_lo4(ld4(grid + o)));
  }
}

int launch_grid_lo_at_points(int b, int n, int r3, int cp, const int *ind, const float *grid, float *grid_lo,
                             cudaStream_t s) {
  const long long total = (long long)b * n * (cp / 4);
  PVB_LAUNCH(grid_lo_at_points_kernel, grid_for(total, 256, kNumSMs * 16), 256, 0, s, n, r3, cp, total, ind, grid,
             grid_lo);
  return 0;
}

__global__ void __launch_bounds__(RED_THREADS) bn_stats_kernel(long long rows, int cp, const float *__restrict__ y,
                                                               float *__restrict__ partials) {
  column_reduce<2>(rows, cp, partials, [&](long long r, int c4, float4 *acc) {
    const float4 v = ld4(y + (size_t)r * cp + c4 * 4);
    acc[0].x += v.x; acc[0].y += v.y; acc[0].z += v.z; acc[0].w += v.w;
    acc[1].x = fmaf(v.x, v.x, acc[1].x); acc[1].y = fmaf(v.y, v.y, acc[1].y);
    acc[1].z = fmaf(v.z, v.z, acc[1].z); acc[1].w = fmaf(v.w, v.w, acc[1].w);
  });
}

int launch_bn_stats(long long rows, int cp, const float *y, float *partials, int *nblocks, cudaStream_t s) {
  PVB_CHECK_ARG(cp % 4 == 0 && cp / 4 <= RED_THREADS);
  const int rl = RED_THREADS / (cp / 4);
  const int g = grid_for(rows, rl * 8, RED_MAX_BLOCKS);
  *nblocks = g;
  PVB_LAUNCH(bn_stats_kernel, g, RED_THREADS, 0, s, rows, cp, y, partials);
  return 0;
}

// sums[j] = sum over blocks (fp64) of partials[block][j]; one warp per column, lanes stride over blocks
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) reduce_partials_kernel(int nblocks, int ncols, const float *__restrict__ partials,
                                                              float *__restrict__ sums) {
  // one CTA per column: 256 threads stride over the partial rows (4 independent fp64 accumulators for MLP)
  __shared__ double red[8];
  const int c = blockIdx.x;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  int k = threadIdx.x;
  for (; k + 768 < nblocks; k += 1024) {
    s0 += (double)partials[(size_t)k * ncols + c];
    s1 += (double)partials[(size_t)(k + 256) * ncols + c];
    s2 += (double)partials[(size_t)(k + 512) * ncols + c];
    s3 += (double)partials[(size_t)(k + 768) * ncols + c];
  }
  for (; k < nblocks; k += 256) s0 += (double)partials[(size_t)k * ncols + c];
  double s = warp_sum_d((s0 + s1) + (s2 + s3));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    sums[c] = (float)t;
  }
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(int nblocks, int c, int cp, double rows, float eps,
                                                          float momentum, const float *__restrict__ partials,
                                                          const float *__restrict__ gamma,
                                                          const float *__restrict__ beta, float *running_mean,
                                                          float *running_var, BnCoef coef,
                                                          long long *num_batches_tracked) {
  __shared__ double red[2][8];
  const int ch = blockIdx.x;  // one CTA per channel
  // nn.BatchNorm's step counter (modules built on torch increment it with one ATen launch per layer and step)
  if (ch == 0 && threadIdx.x == 0 && num_batches_tracked) *num_batches_tracked += 1;
  if (ch >= c) {
    if (threadIdx.x == 0) { coef.mean[ch] = 0.f; coef.invstd[ch] = 0.f; coef.scale[ch] = 0.f; coef.shift[ch] = 0.f; }
    return;
  }
  double s = 0.0, ss = 0.0;
  for (int k = threadIdx.x; k < nblocks; k += blockDim.x) {
    s += (double)partials[((size_t)k * 2 + 0) * cp + ch];
    ss += (double)partials[((size_t)k * 2 + 1) * cp + ch];
  }
  s = warp_sum_d(s);
  ss = warp_sum_d(ss);
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s; red[1][threadIdx.x >> 5] = ss; }
  __syncthreads();
  if (threadIdx.x != 0) return;
  s = 0.0; ss = 0.0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { s += red[0][w]; ss += red[1][w]; }
  const double mean = s / rows;
  double var = ss / rows - mean * mean;
  if (var < 0.0) var = 0.0;
  const float invstd = (float)(1.0 / sqrt(var + (double)eps));
  const float scale = gamma[ch] * invstd;
  coef.mean[ch] = (float)mean;
  coef.invstd[ch] = invstd;
  coef.scale[ch] = scale;
  coef.shift[ch] = beta[ch] - (float)mean * scale;
  if (running_mean) {  // torch: running = (1-m)*running + m*batch, unbiased variance
    running_mean[ch] = (1.0f - momentum) * running_mean[ch] + momentum * (float)mean;
    const double unbiased = rows > 1.0 ? var * rows / (rows - 1.0) : var;
    running_var[ch] = (1.0f - momentum) * running_var[ch] + momentum * (float)unbiased;
  }
}

int launch_bn_finalize(int nblocks, int c, int cp, long long rows, float eps, float momentum, const float *partials,
                       const float *gamma, const float *beta, float *running_mean, float *running_var, BnCoef coef,
                       cudaStream_t s, long long *num_batches_tracked) {
  PVB_LAUNCH(bn_finalize_kernel, cp, 256, 0, s, nblocks, c, cp, (double)rows, eps, momentum, partials,
             gamma, beta, running_mean, running_var, coef, num_batches_tracked);
  return 0;
}

__global__ void __launch_bounds__(256) bn_coef_running_kernel(int c, int cp, float eps, const float *__restrict__ gamma,
                                                              const float *__restrict__ beta,
                                                              const float *__restrict__ rm,
                                                              const float *__restrict__ rv, BnCoef coef) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= cp) return;
  if (ch >= c) {
    coef.mean[ch] = 0.f; coef.invstd[ch] = 0.f; coef.scale[ch] = 0.f; coef.shift[ch] = 0.f;
    return;
  }
  const float invstd = 1.0f / sqrtf(rv[ch] + eps);
  const float scale = gamma[ch] * invstd;
  coef.mean[ch] = rm[ch];
  coef.invstd[ch] = invstd;
  coef.scale[ch] = scale;
  coef.shift[ch] = beta[ch] - rm[ch] * scale;
}

int launch_bn_coef_from_running(int c, float eps, const float *gamma, const float *beta, const float *running_mean,
                                const float *running_var, BnCoef coef, cudaStream_t s) {
  const int cp = (c + 3) / 4 * 4;
  PVB_LAUNCH(bn_coef_running_kernel, ceil_div(cp, 256), 256, 0, s, c, cp, eps, gamma, beta, running_mean,
             running_var, coef);
  return 0;
}

// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bn_apply_leaky_kernel(long long total4, int cp, float slope,
                                                             const float *__restrict__ y, BnCoef coef,
                                                             float *__restrict__ z, float *__restrict__ zlo) {
  extern __shared__ float sco[];  // [2][cp]
  for (int c = threadIdx.x; c < cp; c += blockDim.x) {
    sco[c] = coef.scale[c];
    sco[cp + c] = coef.shift[c];
  }
  __syncthreads();
  const int cp4 = cp >> 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4;
       t += (long long)gridDim.x * blockDim.x) {
    const int c4 = (int)(t % cp4);
    const float4 v = ldg_stream4(y + t * 4);
    const float4 sc = ld4(sco + c4 * 4), sh = ld4(sco + cp + c4 * 4);
    float4 o;
    o.x = leaky(fmaf(v.x, sc.x, sh.x), slope);
    o.y = leaky(fmaf(v.y, sc.y, sh.y), slope);
    o.z = leaky(fmaf(v.z, sc.z, sh.z), slope);
    o.w = leaky(fmaf(v.w, sc.w, sh.w), slope);
    st4(z + t * 4, o);
    if (zlo) st4(zlo + t * 4, tf32_lo4(o));
  }
}

int launch_bn_apply_leaky(long long rows, int cp, float slope, const float *y, BnCoef coef, float *z, float *z_lo,
                          cudaStream_t s) {
  const long long total4 = rows * (cp / 4);
  PVB_LAUNCH(bn_apply_leaky_kernel, grid_for(total4, 256 * 4, kNumSMs * 8), 256, 2 * cp * sizeof(float), s, total4,
             cp, slope, y, coef, z, z_lo);
  return 0;
}

// --------------------------------------------------------------------------------------------
// trilinear corner setup, identical arithmetic to trilinear_devox.cu:36-75
// --------------------------------------------------------------------------------------------
struct Corners {
  float w[8];
  int idx[8];
};
__device__ __forceinline__ void corner_setup(float x, float y, float z, int r, Corners &k) {
  const int r2 = r * r;
  const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  const float xd1 = __fsub_rn(x, xl), yd1 = __fsub_rn(y, yl), zd1 = __fsub_rn(z, zl);
  const float xd0 = __fsub_rn(1.0f, xd1), yd0 = __fsub_rn(1.0f, yd1), zd0 = __fsub_rn(1.0f, zd1);
  const float a00 = __fmul_rn(xd0, yd0), a01 = __fmul_rn(xd0, yd1), a10 = __fmul_rn(xd1, yd0),
              a11 = __fmul_rn(xd1, yd1);
  k.w[0] = __fmul_rn(a00, zd0); k.w[1] = __fmul_rn(a00, zd1);
  k.w[2] = __fmul_rn(a01, zd0); k.w[3] = __fmul_rn(a01, zd1);
  k.w[4] = __fmul_rn(a10, zd0); k.w[5] = __fmul_rn(a10, zd1);
  k.w[6] = __fmul_rn(a11, zd0); k.w[7] = __fmul_rn(a11, zd1);
  const int dz = zd1 > 0 ? 1 : 0, dy = yd1 > 0 ? r : 0, dx = xd1 > 0 ? r2 : 0;
  k.idx[0] = (int)xl * r2 + (int)yl * r + (int)zl;
  k.idx[1] = k.idx[0] + dz;
  k.idx[2] = k.idx[0] + dy;
  k.idx[3] = k.idx[2] + dz;
  k.idx[4] = k.idx[0] + dx;
  k.idx[5] = k.idx[4] + dz;
  k.idx[6] = k.idx[4] + dy;
  k.idx[7] = k.idx[6] + dz;
}

constexpr int PT_TILE = 32;   // points per CTA
constexpr int PT_WARPS = 8;   // each warp walks PT_TILE / PT_WARPS points

// smem: tile[cp][33] floats
__global__ void __launch_bounds__(256) devox_fused_kernel(int n, int c, int cp, int r, float slope,
                                                          const float *__restrict__ nc,
                                                          const float *__restrict__ y2, BnCoef bn2,
                                                          const float *__restrict__ p, BnCoef bnp,
                                                          const float *__restrict__ se_s, float *__restrict__ out) {
  extern __shared__ float tile[];
  const int b = blockIdx.y, i0 = blockIdx.x * PT_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r3 = r * r * r, cp4 = cp >> 2;
  const float *co = nc + (size_t)b * 3 * n;
  // a row of cp <= 64 channels needs only 16 lanes (16 bytes each): two points per warp pass, so all 32 lanes load
  const int lpp = cp4 <= 16 ? 16 : 32, sub = lane / lpp, cl = lane % lpp;
  for (int q = sub; q < PT_TILE / PT_WARPS; q += 32 / lpp) {
    const int pt = warp * (PT_TILE / PT_WARPS) + q;
    const int i = i0 + pt;
    if (i >= n) continue;
    Corners k;
    corner_setup(co[i], co[i + n], co[i + 2 * n], r, k);
    const float *prow = p + ((size_t)b * n + i) * cp;
    for (int c4 = cl; c4 < cp4; c4 += lpp) {
      const float4 sc = ld4(bn2.scale + c4 * 4), sh = ld4(bn2.shift + c4 * 4);
      float4 acc;
      {
        const float4 v = ld4(y2 + ((size_t)b * r3 + k.idx[0]) * cp + c4 * 4);
        acc.x = __fmul_rn(k.w[0], leaky(fmaf(v.x, sc.x, sh.x), slope));
        acc.y = __fmul_rn(k.w[0], leaky(fmaf(v.y, sc.y, sh.y), slope));
        acc.z = __fmul_rn(k.w[0], leaky(fmaf(v.z, sc.z, sh.z), slope));
        acc.w = __fmul_rn(k.w[0], leaky(fmaf(v.w, sc.w, sh.w), slope));
      }
#pragma unroll
      for (int j = 1; j < 8; ++j) {
        const float4 v = ld4(y2 + ((size_t)b * r3 + k.idx[j]) * cp + c4 * 4);
        acc.x = fmaf(k.w[j], leaky(fmaf(v.x, sc.x, sh.x), slope), acc.x);
        acc.y = fmaf(k.w[j], leaky(fmaf(v.y, sc.y, sh.y), slope), acc.y);
        acc.z = fmaf(k.w[j], leaky(fmaf(v.z, sc.z, sh.z), slope), acc.z);
        acc.w = fmaf(k.w[j], leaky(fmaf(v.w, sc.w, sh.w), slope), acc.w);
      }
      if (se_s) {  // SE3d gate: a per-(sample, channel) scale commutes with the (linear) trilinear gather
        const float4 sv = ld4(se_s + (size_t)b * cp + c4 * 4);
        acc.x *= sv.x; acc.y *= sv.y; acc.z *= sv.z; acc.w *= sv.w;
      }
      const float4 pv = ld4(prow + c4 * 4);
      const float4 ps = ld4(bnp.scale + c4 * 4), ph = ld4(bnp.shift + c4 * 4);
      tile[(c4 * 4 + 0) * 33 + pt] = acc.x + fmaxf(fmaf(pv.x, ps.x, ph.x), 0.0f);
      tile[(c4 * 4 + 1) * 33 + pt] = acc.y + fmaxf(fmaf(pv.y, ps.y, ph.y), 0.0f);
      tile[(c4 * 4 + 2) * 33 + pt] = acc.z + fmaxf(fmaf(pv.z, ps.z, ph.z), 0.0f);
      tile[(c4 * 4 + 3) * 33 + pt] = acc.w + fmaxf(fmaf(pv.w, ps.w, ph.w), 0.0f);
    }
  }
  __syncthreads();
  const int i = i0 + lane;
  if (i < n)
    for (int ch = warp; ch < c; ch += PT_WARPS) out[((size_t)b * c + ch) * n + i] = tile[ch * 33 + lane];
}

int launch_devox_fused(int b, int n, int c, int cp, int r, float slope, const float *norm_coords, const float *y2,
                       BnCoef bn2, const float *p, BnCoef bnp, const float *se_s, float *out, cudaStream_t s) {
  const size_t smem = (size_t)cp * 33 * sizeof(float);
  if (smem > 40 * 1024)  // static shared memory counts against the 48 KB default too
    PVB_CUDA(cudaFuncSetAttribute(devox_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PVB_LAUNCH(devox_fused_kernel, dim3(ceil_div(n, PT_TILE), b), 256, smem, s, n, c, cp, r, slope, norm_coords, y2, bn2,
             p, bnp, se_s, out);
  return 0;
}

// --------------------------------------------------------------------------------------------
// backward stage 1 over points
// smem: gtile[cp][33] + red[PT_WARPS][4][cp]
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_points_kernel(int n, int c, int cp, int r, float slope,
                                                         const float *__restrict__ gout,
                                                         const float *__restrict__ nc,
                                                         const float *__restrict__ y2, BnCoef bn2,
                                                         const float *__restrict__ p, BnCoef bnp,
                                                         float *__restrict__ ga_cl, float *__restrict__ d2,
                                                         float *__restrict__ partials,
                                                         const float *__restrict__ se_s,
                                                         float *__restrict__ ds_partials) {
  extern __shared__ float sm[];
  float *gtile = sm;                 // [cp][33]
  float *red = sm + (size_t)cp * 33;  // [PT_WARPS][5][cp]
  const int b = blockIdx.y, i0 = blockIdx.x * PT_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int r3 = r * r * r, cp4 = cp >> 2;
  {
    const int i = i0 + lane;
    for (int ch = warp; ch < cp; ch += PT_WARPS)
      gtile[ch * 33 + lane] = (ch < c && i < n) ? gout[((size_t)b * c + ch) * n + i] : 0.0f;
  }
  __syncthreads();
  const float *co = nc + (size_t)b * 3 * n;
  // cp <= 64: a row needs 16 lanes, so the two half-warps walk different points (all 32 lanes load / scatter) and
  // their reduction partials are combined with one shuffle at the end
  const int lpp = cp4 <= 16 ? 16 : 32, sub = lane / lpp, cl = lane % lpp;
  const unsigned fold_mask = __ballot_sync(0xffffffffu, cl < cp4);  // lanes that run the (single, when lpp == 16) pass
  for (int c4 = cl; c4 < cp4; c4 += lpp) {
    const float4 sc2 = ld4(bn2.scale + c4 * 4), sh2 = ld4(bn2.shift + c4 * 4);
    const float4 mu2 = ld4(bn2.mean + c4 * 4), is2 = ld4(bn2.invstd + c4 * 4);
    const float4 scp = ld4(bnp.scale + c4 * 4), shp = ld4(bnp.shift + c4 * 4);
    const float4 mup = ld4(bnp.mean + c4 * 4), isp = ld4(bnp.invstd + c4 * 4);
    float4 S1 = make_float4(0.f, 0.f, 0.f, 0.f), S2 = S1, T1 = S1, T2 = S1, DS = S1;
    float4 sv = make_float4(1.f, 1.f, 1.f, 1.f);
    if (se_s) sv = ld4(se_s + (size_t)b * cp + c4 * 4);
    for (int q = sub; q < PT_TILE / PT_WARPS; q += 32 / lpp) {
      const int pt = warp * (PT_TILE / PT_WARPS) + q;
      const int i = i0 + pt;
      if (i >= n) continue;
      const float4 g = make_float4(gtile[(c4 * 4 + 0) * 33 + pt], gtile[(c4 * 4 + 1) * 33 + pt],
                                   gtile[(c4 * 4 + 2) * 33 + pt], gtile[(c4 * 4 + 3) * 33 + pt]);
      // ---- point branch: ReLU mask, BN1d reductions
      const size_t prow = ((size_t)b * n + i) * cp + c4 * 4;
      const float4 pv = ld4(p + prow);
      float4 ga;
      ga.x = fmaf(pv.x, scp.x, shp.x) > 0.f ? g.x : 0.f;
      ga.y = fmaf(pv.y, scp.y, shp.y) > 0.f ? g.y : 0.f;
      ga.z = fmaf(pv.z, scp.z, shp.z) > 0.f ? g.z : 0.f;
      ga.w = fmaf(pv.w, scp.w, shp.w) > 0.f ? g.w : 0.f;
      st4(ga_cl + prow, ga);
      S1.x += ga.x; S1.y += ga.y; S1.z += ga.z; S1.w += ga.w;
      S2.x = fmaf(ga.x, (pv.x - mup.x) * isp.x, S2.x);
      S2.y = fmaf(ga.y, (pv.y - mup.y) * isp.y, S2.y);
      S2.z = fmaf(ga.z, (pv.z - mup.z) * isp.z, S2.z);
      S2.w = fmaf(ga.w, (pv.w - mup.w) * isp.w, S2.w);
      // ---- voxel branch: trilinear scatter of leaky'(.) * w * g, BN3d reductions
      Corners k;
      corner_setup(co[i], co[i + n], co[i + 2 * n], r, k);
      const float4 gs = make_float4(g.x * sv.x, g.y * sv.y, g.z * sv.z, g.w * sv.w);  // gradient behind the SE gate
      float4 V = make_float4(0.f, 0.f, 0.f, 0.f);  // un-gated voxel-branch output (needed for d gate)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const size_t vrow = ((size_t)b * r3 + k.idx[j]) * cp + c4 * 4;
        const float4 v = ld4(y2 + vrow);
        const float zx = fmaf(v.x, sc2.x, sh2.x), zy = fmaf(v.y, sc2.y, sh2.y), zz = fmaf(v.z, sc2.z, sh2.z),
                    zw = fmaf(v.w, sc2.w, sh2.w);
        if (se_s) {
          V.x = fmaf(k.w[j], leaky(zx, slope), V.x); V.y = fmaf(k.w[j], leaky(zy, slope), V.y);
          V.z = fmaf(k.w[j], leaky(zz, slope), V.z); V.w = fmaf(k.w[j], leaky(zw, slope), V.w);
        }
        float4 gk;
        gk.x = k.w[j] * gs.x * (zx > 0.f ? 1.0f : slope);
        gk.y = k.w[j] * gs.y * (zy > 0.f ? 1.0f : slope);
        gk.z = k.w[j] * gs.z * (zz > 0.f ? 1.0f : slope);
        gk.w = k.w[j] * gs.w * (zw > 0.f ? 1.0f : slope);
        T1.x += gk.x; T1.y += gk.y; T1.z += gk.z; T1.w += gk.w;
        T2.x = fmaf(gk.x, (v.x - mu2.x) * is2.x, T2.x);
        T2.y = fmaf(gk.y, (v.y - mu2.y) * is2.y, T2.y);
        T2.z = fmaf(gk.z, (v.z - mu2.z) * is2.z, T2.z);
        T2.w = fmaf(gk.w, (v.w - mu2.w) * is2.w, T2.w);
        if (k.w[j] != 0.0f) red_add4(d2 + vrow, gk);
      }
      DS.x = fmaf(g.x, V.x, DS.x); DS.y = fmaf(g.y, V.y, DS.y); DS.z = fmaf(g.z, V.z, DS.z); DS.w = fmaf(g.w, V.w, DS.w);
    }
    if (lpp == 16) {  // fold the other half-warp's points in (same channel quad, lanes l and l ^ 16)
#define PVB_FOLD4(v)                                      \
  v.x += __shfl_xor_sync(fold_mask, v.x, 16);             \
  v.y += __shfl_xor_sync(fold_mask, v.y, 16);             \
  v.z += __shfl_xor_sync(fold_mask, v.z, 16);             \
  v.w += __shfl_xor_sync(fold_mask, v.w, 16);
      PVB_FOLD4(S1) PVB_FOLD4(S2) PVB_FOLD4(T1) PVB_FOLD4(T2) PVB_FOLD4(DS)
#undef PVB_FOLD4
    }
    if (sub == 0) {
      st4(red + ((size_t)warp * 5 + 0) * cp + c4 * 4, S1);
      st4(red + ((size_t)warp * 5 + 1) * cp + c4 * 4, S2);
      st4(red + ((size_t)warp * 5 + 2) * cp + c4 * 4, T1);
      st4(red + ((size_t)warp * 5 + 3) * cp + c4 * 4, T2);
      st4(red + ((size_t)warp * 5 + 4) * cp + c4 * 4, DS);
    }
  }
  __syncthreads();
  const size_t blk = (size_t)blockIdx.y * gridDim.x + blockIdx.x;  // blocks of one sample are contiguous
  for (int t = threadIdx.x; t < 5 * cp; t += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < PT_WARPS; ++w) s += red[(size_t)w * 5 * cp + t];
    if (t < 4 * cp) partials[blk * 4 * cp + t] = s;
    else if (ds_partials) ds_partials[blk * cp + (t - 4 * cp)] = s;
  }
}

int launch_bwd_points(int b, int n, int c, int cp, int r, float slope, const float *grad_out,
                      const float *norm_coords, const float *y2, BnCoef bn2, const float *p, BnCoef bnp, float *ga_cl,
                      float *d2, float *partials, int *nblocks, const float *se_s, float *ds_partials, cudaStream_t s) {
  const size_t smem = ((size_t)cp * 33 + (size_t)PT_WARPS * 5 * cp) * sizeof(float);
  if (smem > 40 * 1024)  // static shared memory counts against the 48 KB default too
    PVB_CUDA(cudaFuncSetAttribute(bwd_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid(ceil_div(n, PT_TILE), b);
  *nblocks = (int)(grid.x * grid.y);
  PVB_LAUNCH(bwd_points_kernel, grid, 256, smem, s, n, c, cp, r, slope, grad_out, norm_coords, y2, bn2, p, bnp, ga_cl,
             d2, partials, se_s, ds_partials);
  return 0;
}

int launch_reduce_partials(int nblocks, int ncols, const float *partials, float *sums, cudaStream_t s) {
  PVB_LAUNCH(reduce_partials_kernel, ncols, 256, 0, s, nblocks, ncols, partials, sums);
  return 0;
}

// --------------------------------------------------------------------------------------------
// BatchNorm backward, input gradient:  dx = scale * (g' - mean(g') - xhat * mean(g' * xhat)),
// g' = g (use_mask = 0) or leaky'(bn(y)) * g (use_mask = 1); s1/s2 are the RAW column sums of g' and
// g'*xhat, inv_count = 1/rows.  Also emits column sums of dx (the conv-bias gradient).
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) bn_bwd_apply_kernel(long long rows, int cp, int use_mask, float slope,
                                                                   float inv_count, const float *__restrict__ g,
                                                                   const float *__restrict__ y, BnCoef coef,
                                                                   const float *__restrict__ s1,
                                                                   const float *__restrict__ s2,
                                                                   float *__restrict__ out,
                                                                   float *__restrict__ out_lo,
                                                                   float *__restrict__ partials,
                                                                   const float *__restrict__ extra,
                                                                   long long rows_per_sample) {
  column_reduce<1>(rows, cp, partials, [&](long long r, int c4, float4 *acc) {
    const size_t o = (size_t)r * cp + c4 * 4;
    const float4 gv = ldg_stream4(g + o), yv = ldg_stream4(y + o);
    const float4 sc = ld4(coef.scale + c4 * 4), sh = ld4(coef.shift + c4 * 4);
    const float4 mu = ld4(coef.mean + c4 * 4), is = ld4(coef.invstd + c4 * 4);
    const float4 a1 = ld4(s1 + c4 * 4), a2 = ld4(s2 + c4 * 4);
    float4 ex = make_float4(0.f, 0.f, 0.f, 0.f);  // dense SE term: d(mean over voxels) reaches every voxel
    if (extra) ex = ld4(extra + (size_t)(r / rows_per_sample) * cp + c4 * 4);
    float4 d;
#define PVB_BWD1(f)                                                                                   \
  {                                                                                                   \
    float gg = gv.f;                                                                                  \
    if (use_mask) gg *= (fmaf(yv.f, sc.f, sh.f) > 0.f ? 1.0f : slope);                                \
    if (extra) gg += (fmaf(yv.f, sc.f, sh.f) > 0.f ? 1.0f : slope) * ex.f;                            \
    const float xh = (yv.f - mu.f) * is.f;                                                            \
    d.f = sc.f * (gg - a1.f * inv_count - xh * (a2.f * inv_count));                                   \
    acc[0].f += d.f;                                                                                  \
  }
    PVB_BWD1(x) PVB_BWD1(y) PVB_BWD1(z) PVB_BWD1(w)
#undef PVB_BWD1
    st4(out + o, d);
    if (out_lo) st4(out_lo + o, tf32_lo4(d));
  });
}

int launch_bn_bwd_apply(long long rows, int cp, int use_mask, float slope, const float *g, const float *y,
                        BnCoef coef, const float *s1, const float *s2, float *out, float *out_lo,
                        float *colsum_partials, int *nblocks, cudaStream_t s, const float *extra,
                        long long rows_per_sample) {
  PVB_CHECK_ARG(cp % 4 == 0 && cp / 4 <= RED_THREADS);
  const int rl = RED_THREADS / (cp / 4);
  const int gsz = grid_for(rows, rl * 8, RED_MAX_BLOCKS);
  *nblocks = gsz;
  PVB_LAUNCH(bn_bwd_apply_kernel, gsz, RED_THREADS, 0, s, rows, cp, use_mask, slope, (float)(1.0 / (double)rows), g, y,
             coef, s1, s2, out, out_lo, colsum_partials, extra, rows_per_sample > 0 ? rows_per_sample : 1);
  return 0;
}

__global__ void __launch_bounds__(RED_THREADS) bn_bwd_reduce_kernel(long long rows, int cp, float slope,
                                                                    const float *__restrict__ g,
                                                                    const float *__restrict__ y, BnCoef coef,
                                                                    float *__restrict__ partials) {
  column_reduce<2>(rows, cp, partials, [&](long long r, int c4, float4 *acc) {
    const size_t o = (size_t)r * cp + c4 * 4;
    const float4 gv = ld4(g + o), yv = ld4(y + o);
    const float4 sc = ld4(coef.scale + c4 * 4), sh = ld4(coef.shift + c4 * 4);
    const float4 mu = ld4(coef.mean + c4 * 4), is = ld4(coef.invstd + c4 * 4);
#define PVB_RED1(f)                                                            \
  {                                                                            \
    const float gg = gv.f * (fmaf(yv.f, sc.f, sh.f) > 0.f ? 1.0f : slope);     \
    acc[0].f += gg;                                                            \
    acc[1].f = fmaf(gg, (yv.f - mu.f) * is.f, acc[1].f);                       \
  }
    PVB_RED1(x) PVB_RED1(y) PVB_RED1(z) PVB_RED1(w)
#undef PVB_RED1
  });
}

int launch_bn_bwd_reduce(long long rows, int cp, float slope, const float *g, const float *y, BnCoef coef,
                         float *partials, int *nblocks, cudaStream_t s) {
  PVB_CHECK_ARG(cp % 4 == 0 && cp / 4 <= RED_THREADS);
  const int rl = RED_THREADS / (cp / 4);
  const int gsz = grid_for(rows, rl * 8, RED_MAX_BLOCKS);
  *nblocks = gsz;
  PVB_LAUNCH(bn_bwd_reduce_kernel, gsz, RED_THREADS, 0, s, rows, cp, slope, g, y, coef, partials);
  return 0;
}


// --------------------------------------------------------------------------------------------
// SE3d (modules/se.py:6-17) inside the fused block.
//   forward : pooled[b,c] = mean_v leaky(bn2(Y2)), gate s = sigmoid(W2 relu(W1 pooled)); the gate multiplies the
//             devoxelized voxel branch (a per-(sample,channel) scale commutes with the linear gather)
//   backward: d gate from the un-gated voxel output; d pooled is a dense 1/R^3 term on every voxel, folded into the
//             BN2 backward reductions (via A1 = sum leaky', A2 = sum leaky'*xhat per sample) and bn_bwd_apply.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RED_THREADS) se_pool_kernel(long long rows_per_sample, int cp, float slope,
                                                              const float *__restrict__ y, BnCoef coef,
                                                              float *__restrict__ partials) {
  const long long b = blockIdx.y;
  column_reduce<3>(rows_per_sample, cp, partials, [&](long long r, int c4, float4 *acc) {
    const float4 v = ld4(y + (size_t)r * cp + c4 * 4);
    const float4 sc = ld4(coef.scale + c4 * 4), sh = ld4(coef.shift + c4 * 4);
    const float4 mu = ld4(coef.mean + c4 * 4), is = ld4(coef.invstd + c4 * 4);
#define PVB_SE1(f)                                        \
  {                                                       \
    const float z = fmaf(v.f, sc.f, sh.f);                \
    const float dl = z > 0.f ? 1.0f : slope;              \
    acc[0].f += z * dl;                                   \
    acc[1].f += dl;                                       \
    acc[2].f = fmaf(dl, (v.f - mu.f) * is.f, acc[2].f);   \
  }
    PVB_SE1(x) PVB_SE1(y) PVB_SE1(z) PVB_SE1(w)
#undef PVB_SE1
  }, b * rows_per_sample, b * gridDim.x + blockIdx.x);
}

// sums[b][j] = sum over the sample's blocks of partials[b][blk][j]
__global__ void __launch_bounds__(256) reduce_partials_batched_kernel(int nblocks, int ncols,
                                                                      const float *__restrict__ partials,
                                                                      float *__restrict__ sums) {
  const int c = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, b = blockIdx.y;
  if (c >= ncols) return;
  double s = 0.0;
  for (int k = lane; k < nblocks; k += 32) s += (double)partials[((size_t)b * nblocks + k) * ncols + c];
  s = warp_sum_d(s);
  if (lane == 0) sums[(size_t)b * ncols + c] = (float)s;
}

int launch_se_pool(int b, long long rows_per_sample, int cp, float slope, const float *y, BnCoef coef, float *partials,
                   float *pooled3 /*[b][3][cp]*/, cudaStream_t s) {
  PVB_CHECK_ARG(cp % 4 == 0 && cp / 4 <= RED_THREADS);
  const int rl = RED_THREADS / (cp / 4);
  int gx = grid_for(rows_per_sample, rl * 8, max(1, RED_MAX_BLOCKS / b));
  PVB_LAUNCH(se_pool_kernel, dim3(gx, b), RED_THREADS, 0, s, rows_per_sample, cp, slope, y, coef, partials);
  PVB_LAUNCH(reduce_partials_batched_kernel, dim3(ceil_div(3 * cp, 8), b), 256, 0, s, gx, 3 * cp, partials, pooled3);
  return 0;
}

int launch_reduce_partials_batched(int b, int nblocks_per_sample, int ncols, const float *partials, float *sums,
                                   cudaStream_t s) {
  PVB_LAUNCH(reduce_partials_batched_kernel, dim3(ceil_div(ncols, 8), b), 256, 0, s, nblocks_per_sample, ncols,
             partials, sums);
  return 0;
}

// one CTA per sample: gate = sigmoid(W2 relu(W1 mean));  smem: mean[c] + hidden[h]
__global__ void __launch_bounds__(128) se_fc_kernel(int c, int cp, int hid, float inv_rows,
                                                    const float *__restrict__ pooled3, const float *__restrict__ w1,
                                                    const float *__restrict__ w2, float *__restrict__ mean_out,
                                                    float *__restrict__ hidden_out, float *__restrict__ gate) {
  extern __shared__ float sm[];
  float *m = sm, *h = sm + cp;
  const int b = blockIdx.x;
  for (int i = threadIdx.x; i < cp; i += blockDim.x) {
    const float v = i < c ? pooled3[((size_t)b * 3 + 0) * cp + i] * inv_rows : 0.f;
    m[i] = v;
    mean_out[(size_t)b * cp + i] = v;
  }
  __syncthreads();
  for (int j = threadIdx.x; j < hid; j += blockDim.x) {
    float a = 0.f;
    for (int i = 0; i < c; ++i) a = fmaf(w1[(size_t)j * c + i], m[i], a);
    a = fmaxf(a, 0.f);
    h[j] = a;
    hidden_out[(size_t)b * hid + j] = a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cp; i += blockDim.x) {
    float a = 0.f;
    if (i < c) {
      for (int j = 0; j < hid; ++j) a = fmaf(w2[(size_t)i * hid + j], h[j], a);
      a = 1.0f / (1.0f + expf(-a));
    }
    gate[(size_t)b * cp + i] = a;
  }
}

int launch_se_fc(int b, int c, int cp, int hid, long long rows_per_sample, const float *pooled3, const float *w1,
                 const float *w2, float *mean_out, float *hidden_out, float *gate, cudaStream_t s) {
  PVB_LAUNCH(se_fc_kernel, b, 128, (cp + hid) * sizeof(float), s, c, cp, hid, (float)(1.0 / (double)rows_per_sample),
             pooled3, w1, w2, mean_out, hidden_out, gate);
  return 0;
}

// single CTA: backward through sigmoid / FC2 / ReLU / FC1 for every sample; emits dW1, dW2 and
// extra[b][c] = d pooled[b][c] / rows_per_sample (the dense per-voxel gradient of the mean)
__global__ void __launch_bounds__(256) se_fc_bwd_kernel(int nb, int c, int cp, int hid, float inv_rows,
                                                        const float *__restrict__ dgate_sum /*[b][cp]*/,
                                                        const float *__restrict__ gate,
                                                        const float *__restrict__ hidden,
                                                        const float *__restrict__ mean, const float *__restrict__ w1,
                                                        const float *__restrict__ w2, float *__restrict__ dw1,
                                                        float *__restrict__ dw2, float *__restrict__ extra) {
  extern __shared__ float sm[];
  float *dsig = sm, *dh = sm + cp;
  for (int i = threadIdx.x; i < hid * c; i += blockDim.x) { dw1[i] = 0.f; dw2[i] = 0.f; }
  __syncthreads();
  for (int b = 0; b < nb; ++b) {
    for (int i = threadIdx.x; i < cp; i += blockDim.x) {
      const float sg = gate[(size_t)b * cp + i];
      dsig[i] = i < c ? dgate_sum[(size_t)b * cp + i] * sg * (1.0f - sg) : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < hid; j += blockDim.x) {
      float a = 0.f;
      for (int i = 0; i < c; ++i) a = fmaf(w2[(size_t)i * hid + j], dsig[i], a);
      dh[j] = hidden[(size_t)b * hid + j] > 0.f ? a : 0.f;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < c * hid; t += blockDim.x) {
      const int i = t / hid, j = t % hid;                      // dw2[i][j], dw1[j][i]
      dw2[(size_t)i * hid + j] += dsig[i] * hidden[(size_t)b * hid + j];
      dw1[(size_t)j * c + i] += dh[j] * mean[(size_t)b * cp + i];
    }
    for (int i = threadIdx.x; i < cp; i += blockDim.x) {
      float a = 0.f;
      if (i < c)
        for (int j = 0; j < hid; ++j) a = fmaf(w1[(size_t)j * c + i], dh[j], a);
      extra[(size_t)b * cp + i] = a * inv_rows;
    }
    __syncthreads();
  }
}

// T1[c] += sum_b extra[b][c] * A1[b][c],  T2[c] += sum_b extra[b][c] * A2[b][c]
__global__ void __launch_bounds__(256) se_fix_sums_kernel(int nb, int cp, const float *__restrict__ extra,
                                                          const float *__restrict__ pooled3, float *t1, float *t2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cp) return;
  float a1 = 0.f, a2 = 0.f;
  for (int b = 0; b < nb; ++b) {
    const float e = extra[(size_t)b * cp + i];
    a1 = fmaf(e, pooled3[((size_t)b * 3 + 1) * cp + i], a1);
    a2 = fmaf(e, pooled3[((size_t)b * 3 + 2) * cp + i], a2);
  }
  t1[i] += a1;
  t2[i] += a2;
}

int launch_se_backward(int nb, int c, int cp, int hid, long long rows_per_sample, const float *dgate_sum,
                       const float *gate, const float *hidden, const float *mean, const float *pooled3,
                       const float *w1, const float *w2, float *dw1, float *dw2, float *extra, float *t1, float *t2,
                       cudaStream_t s) {
  PVB_LAUNCH(se_fc_bwd_kernel, 1, 256, (cp + hid) * sizeof(float), s, nb, c, cp, hid,
             (float)(1.0 / (double)rows_per_sample), dgate_sum, gate, hidden, mean, w1, w2, dw1, dw2, extra);
  PVB_LAUNCH(se_fix_sums_kernel, ceil_div(cp, 256), 256, 0, s, nb, cp, extra, pooled3, t1, t2);
  return 0;
}

// --------------------------------------------------------------------------------------------
// grad_features[b,c,i] = gG0[b*R^3 + ind_i, c] / cnt + gFpt[b*N+i, c]      (vox.cu:86-110 + add)
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bwd_final_kernel(int n, int c, int cp, int r3, const int *__restrict__ ind,
                                                        const int *__restrict__ cnt,
                                                        const float *__restrict__ gg0,
                                                        const float *__restrict__ gfpt,
                                                        float *__restrict__ gfeat) {
  extern __shared__ float tile[];  // [cp][33]
  const int b = blockIdx.y, i0 = blockIdx.x * PT_TILE;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int cp4 = cp >> 2;
  for (int q = 0; q < PT_TILE / PT_WARPS; ++q) {
    const int pt = warp * (PT_TILE / PT_WARPS) + q;
    const int i = i0 + pt;
    if (i >= n) break;
    const int pos = ind[(size_t)b * n + i];
    const int cur = cnt[(size_t)b * r3 + pos];
    const float inv = cur > 0 ? (float)(1.0 / (double)(float)cur) : 0.0f;
    for (int c4 = lane; c4 < cp4; c4 += 32) {
      const float4 gv = ld4(gg0 + ((size_t)b * r3 + pos) * cp + c4 * 4);
      const float4 pv = ld4(gfpt + ((size_t)b * n + i) * cp + c4 * 4);
      tile[(c4 * 4 + 0) * 33 + pt] = fmaf(gv.x, inv, pv.x);
      tile[(c4 * 4 + 1) * 33 + pt] = fmaf(gv.y, inv, pv.y);
      tile[(c4 * 4 + 2) * 33 + pt] = fmaf(gv.z, inv, pv.z);
      tile[(c4 * 4 + 3) * 33 + pt] = fmaf(gv.w, inv, pv.w);
    }
  }
  __syncthreads();
  const int i = i0 + lane;
  if (i < n)
    for (int ch = warp; ch < c; ch += PT_WARPS) gfeat[((size_t)b * c + ch) * n + i] = tile[ch * 33 + lane];
}

int launch_bwd_final(int b, int n, int c, int cp, int r3, const int *ind, const int *cnt, const float *gg0,
                     const float *gfpt, float *grad_features, cudaStream_t s) {
  const size_t smem = (size_t)cp * 33 * sizeof(float);
  if (smem > 40 * 1024)  // static shared memory counts against the 48 KB default too
    PVB_CUDA(cudaFuncSetAttribute(bwd_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PVB_LAUNCH(bwd_final_kernel, dim3(ceil_div(n, PT_TILE), b), 256, smem, s, n, c, cp, r3, ind, cnt, gg0, gfpt,
             grad_features);
  return 0;
}


// --------------------------------------------------------------------------------------------
// Activity (sparsity) bookkeeping.  Voxelization(normalize=True) maps every cloud into the sphere of diameter 1
// inscribed in the unit cube (modules/voxelization.py:19), so at most pi/6 of the grid can ever be occupied -- in practice
// 10-20 %.  conv1's input is exactly zero elsewhere, conv2's input is a per-channel constant there.  We therefore build
// compact lists of the tiles that can differ from that closed form and let the tensor-core kernels walk only those.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) colocc_kernel(int r, long long ncols, const int *__restrict__ cnt,
                                                     unsigned char *__restrict__ occ) {
  const long long col = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= ncols) return;
  const int *c = cnt + col * r;
  int any = 0;
  for (int z = 0; z < r; ++z) any |= c[z];
  occ[col] = any != 0;
}

__device__ __forceinline__ bool occ_any(const unsigned char *occ, int b, int r, int xa, int xb, int ya, int yb) {
  xa = max(xa, 0); ya = max(ya, 0); xb = min(xb, r - 1); yb = min(yb, r - 1);
  for (int x = xa; x <= xb; ++x)
    for (int y = ya; y <= yb; ++y)
      if (occ[((size_t)b * r + x) * r + y]) return true;
  return false;
}

// Flag kernels mark, per halo-conv unit and per wgrad k-tile, whether it has to be computed; compact_lists_kernel then
// turns each flag array into an index-ordered coordinate list (ordered = neighbouring tiles run concurrently and share
// their halos in L2; also makes the traversal deterministic).
//   lists / counts: [0] conv1-forward units, [1] conv1-dgrad units, [2] conv2-forward units, [3] conv1-wgrad k-tiles,
//                   [4] conv2-wgrad k-tiles, [5] region-G units (conv2 dgrad + BN1 backward)
__global__ void __launch_bounds__(256) flag_stage1_kernel(int nb, int r, int ty, int wg_bz, int wg_by,
                                                          const unsigned char *__restrict__ occ,
                                                          unsigned char *__restrict__ act1,
                                                          unsigned char *__restrict__ act_dg,
                                                          unsigned char *__restrict__ wg1_flag) {
  const int pairs_x = (r + 1) / 2, tiles_y = (r + ty - 1) / ty;
  const int n_units = nb * pairs_x * tiles_y;
  const int wg_ty = (r + wg_by - 1) / wg_by, wg_tz = (r + wg_bz - 1) / wg_bz;
  const int n_kt = nb * r * wg_ty * wg_tz;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_units) {
    int u = t;
    const int yt = u % tiles_y; u /= tiles_y;
    const int xp = u % pairs_x; u /= pairs_x;
    const int b = u, x0 = xp * 2, y0 = yt * ty;
    act1[t] = occ_any(occ, b, r, x0 - 1, x0 + 2, y0 - 1, y0 + ty);     // halo of the two tiles sees a point
    act_dg[t] = occ_any(occ, b, r, x0, x0 + 1, y0, y0 + ty - 1);       // the tiles themselves contain a point
  }
  if (t < n_kt) {
    int u = t;
    u /= wg_tz;
    const int tyi = u % wg_ty; u /= wg_ty;
    const int x = u % r; u /= r;
    const int b = u, y0 = tyi * wg_by;
    // X shifted by (dx,dy) in [-1,1] must be non-zero somewhere in the tile's rows (occupancy is per (x,y) column)
    wg1_flag[t] = occ_any(occ, b, r, x - 1, x + 1, y0 - 1, y0 + wg_by);
  }
}

// Y1 (conv1 output) equals its bias exactly wherever no occupied voxel lies within one voxel, so Z1 = leaky(bn1(Y1)) is
// the constant c1 outside the occupancy dilated by 1, and conv2 / its weight gradient only need the tiles that can see
// the occupancy dilated by 2 (column granularity; occupancy is tracked per (x,y) column).
__global__ void __launch_bounds__(256) flag_stage2_kernel(int nb, int r, int ty, int wg_bz, int wg_by,
                                                          const unsigned char *__restrict__ occ,
                                                          const unsigned char *__restrict__ act_dg,
                                                          unsigned char *__restrict__ fwd2_flag,
                                                          unsigned char *__restrict__ wg2_flag,
                                                          unsigned char *__restrict__ dg2_flag) {
  const int pairs_x = (r + 1) / 2, tiles_y = (r + ty - 1) / ty;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  {  // conv2 weight gradient: a tap-shifted row (+-1) can see a non-constant Z1 voxel (occupancy dilated by 1)
    const int wg_ty = (r + wg_by - 1) / wg_by, wg_tz = (r + wg_bz - 1) / wg_bz;
    if (t < nb * r * wg_ty * wg_tz) {
      int u = t;
      u /= wg_tz;
      const int tyi = u % wg_ty; u /= wg_ty;
      const int x = u % r; u /= r;
      const int b = u, y0 = tyi * wg_by;
      wg2_flag[t] = occ_any(occ, b, r, x - 2, x + 2, y0 - 2, y0 + wg_by + 1);
    }
  }
  if (t >= nb * pairs_x * tiles_y) return;
  int u = t;
  const int yt = u % tiles_y; u /= tiles_y;
  const int xp = u % pairs_x; u /= pairs_x;
  const int b = u, x0 = xp * 2, y0 = yt * ty;
  // conv2 forward: the unit's halo (+-1) touches a non-constant Z1 voxel (occupancy dilated by 1)
  fwd2_flag[t] = occ_any(occ, b, r, x0 - 2, x0 + 3, y0 - 2, y0 + ty + 1);
  // region G: units whose gY1 is consumed (halo of a conv1-dgrad unit, rows of conv1-wgrad k-tiles) = 3x3 unit
  // dilation of the units that contain occupied columns; conv2's data gradient and BN1-backward run on G only
  bool gq = false;
  for (int i = max(0, xp - 1); i <= min(pairs_x - 1, xp + 1); ++i)
    for (int j = max(0, yt - 1); j <= min(tiles_y - 1, yt + 1); ++j) gq = gq || act_dg[((size_t)b * pairs_x + i) * tiles_y + j];
  dg2_flag[t] = gq;
}

struct CompactJob {
  const unsigned char *flags;
  int4 *list;
  int n;
  int is_ktile;
};
struct CompactJobs { CompactJob j[6]; };

// Ordered stream compaction of the six flag arrays into coordinate lists, spread over the chip:
//   pass 1: one CTA per (1024-flag chunk, list) counts its set flags;
//   pass 2: the same grid: a chunk's base = sum of the earlier chunks' counts, ballot scan inside the chunk, write.
// (r01: a single 1024-thread CTA per list walked up to 16 chunks serially: 32 us.)
__global__ void __launch_bounds__(1024) compact_count_kernel(CompactJobs jobs, int chunks, int *__restrict__ chunk_counts) {
  __shared__ int warp_tot[32];
  const CompactJob jb = jobs.j[blockIdx.y];
  const int i = blockIdx.x * 1024 + threadIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool f = i < jb.n && jb.flags[i];
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  if (warp == 0) {
    int v = warp_tot[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) chunk_counts[blockIdx.y * chunks + blockIdx.x] = v;
  }
}

__global__ void __launch_bounds__(1024) compact_write_kernel(CompactJobs jobs, int chunks, int r, int ty, int wg_bz,
                                                             int wg_by, const int *__restrict__ chunk_counts,
                                                             int *__restrict__ counts) {
  __shared__ int warp_tot[32];
  const CompactJob jb = jobs.j[blockIdx.y];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pairs_x = (r + 1) / 2, tiles_y = (r + ty - 1) / ty;
  const int wg_ty = (r + wg_by - 1) / wg_by, wg_tz = (r + wg_bz - 1) / wg_bz;
  int base = 0;
  for (int c = 0; c < (int)blockIdx.x; ++c) base += __ldg(chunk_counts + blockIdx.y * chunks + c);
  const int i = blockIdx.x * 1024 + threadIdx.x;
  const bool f = i < jb.n && jb.flags[i];
  const unsigned bal = __ballot_sync(0xffffffffu, f);
  if (lane == 0) warp_tot[warp] = __popc(bal);
  __syncthreads();
  int woff = 0, tot = 0;
  for (int w = 0; w < 32; ++w) {
    const int v = warp_tot[w];
    if (w < warp) woff += v;
    tot += v;
  }
  if (f) {
    const int pos = base + woff + __popc(bal & ((1u << lane) - 1));
    int u = i;
    if (jb.is_ktile) {
      const int tz = u % wg_tz; u /= wg_tz;
      const int tyi = u % wg_ty; u /= wg_ty;
      const int x = u % r; u /= r;
      jb.list[pos] = make_int4(tz * wg_bz, tyi * wg_by, x, u);
    } else {
      const int yt = u % tiles_y; u /= tiles_y;
      const int xp = u % pairs_x; u /= pairs_x;
      jb.list[pos] = make_int4(xp * 2, yt * ty, u, 0);
    }
  }
  // the chunk that holds the list's last flag publishes the total
  if (threadIdx.x == 0 && (int)blockIdx.x == (jb.n - 1) / 1024) counts[blockIdx.y] = base + tot;
}

int launch_build_activity(int nb, int r, int ty, int wg_bz, int wg_by, const int *cnt, int *counts, unsigned char *occ,
                          unsigned char *act1, unsigned char *act_dg, unsigned char *fwd2_flag, unsigned char *wg1_flag,
                          unsigned char *wg2_flag, unsigned char *dg2_flag, int4 *fwd1, int4 *dgrad1, int4 *fwd2,
                          int4 *wg1, int4 *wg2, int4 *dg2, int *chunk_counts, cudaStream_t s) {
  const long long ncols = (long long)nb * r * r;
  PVB_LAUNCH(colocc_kernel, ceil_div(ncols, 256), 256, 0, s, r, ncols, cnt, occ);
  const int n_units = nb * ((r + 1) / 2) * ((r + ty - 1) / ty);
  const int n_kt = nb * r * ((r + wg_by - 1) / wg_by) * ((r + wg_bz - 1) / wg_bz);
  const int nmax = max(n_units, n_kt);
  PVB_LAUNCH(flag_stage1_kernel, ceil_div(nmax, 256), 256, 0, s, nb, r, ty, wg_bz, wg_by, occ, act1, act_dg, wg1_flag);
  PVB_LAUNCH(flag_stage2_kernel, ceil_div(nmax, 256), 256, 0, s, nb, r, ty, wg_bz, wg_by, occ, act_dg, fwd2_flag, wg2_flag,
             dg2_flag);
  CompactJobs jobs;
  jobs.j[0] = CompactJob{act1, fwd1, n_units, 0};
  jobs.j[1] = CompactJob{act_dg, dgrad1, n_units, 0};
  jobs.j[2] = CompactJob{fwd2_flag, fwd2, n_units, 0};
  jobs.j[3] = CompactJob{wg1_flag, wg1, n_kt, 1};
  jobs.j[4] = CompactJob{wg2_flag, wg2, n_kt, 1};
  jobs.j[5] = CompactJob{dg2_flag, dg2, n_units, 0};
  const int chunks = ceil_div(nmax, 1024);
  PVB_LAUNCH(compact_count_kernel, dim3(chunks, 6), 1024, 0, s, jobs, chunks, chunk_counts);
  PVB_LAUNCH(compact_write_kernel, dim3(chunks, 6), 1024, 0, s, jobs, chunks, r, ty, wg_bz, wg_by, chunk_counts, counts);
  return 0;
}

// out[row][c] = bias[c]   (conv of an all-zero neighbourhood)
__global__ void __launch_bounds__(256) fill_bias_rows_kernel(long long total4, int c, int cp,
                                                             const float *__restrict__ bias, float *__restrict__ out) {
  const int cp4 = cp >> 2;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4; t += (long long)gridDim.x * blockDim.x) {
    const int c0 = (int)(t % cp4) * 4;
    float4 v;
    v.x = c0 + 0 < c ? __ldg(bias + c0 + 0) : 0.f;
    v.y = c0 + 1 < c ? __ldg(bias + c0 + 1) : 0.f;
    v.z = c0 + 2 < c ? __ldg(bias + c0 + 2) : 0.f;
    v.w = c0 + 3 < c ? __ldg(bias + c0 + 3) : 0.f;
    stg_stream4(out + t * 4, v);
  }
}

int launch_fill_bias_rows(long long rows, int c, int cp, const float *bias, float *out, cudaStream_t s) {
  const long long total4 = rows * (cp / 4);
  PVB_LAUNCH(fill_bias_rows_kernel, grid_for(total4, 256 * 4, kNumSMs * 8), 256, 0, s, total4, c, cp, bias, out);
  return 0;
}

// conv of a per-channel constant input c1[ci] = leaky(bn1(b1[ci])): 27 boundary classes (lo edge / interior / hi edge
// per axis decide which taps fall inside the grid).  classsum[cls][co] = b2[co] + sum_{valid taps} T[tap][co],
// T[tap][co] = sum_ci w[co][ci][tap] c1[ci]
__device__ __forceinline__ bool tap_valid_in_class(int tap, int cls) {
  const int dx = tap / 9 - 1, dy = (tap / 3) % 3 - 1, dz = tap % 3 - 1;
  const int cx = cls / 9, cy = (cls / 3) % 3, cz = cls % 3;
  return !(cx == 0 && dx < 0) && !(cx == 2 && dx > 0) && !(cy == 0 && dy < 0) && !(cy == 2 && dy > 0) &&
         !(cz == 0 && dz < 0) && !(cz == 2 && dz > 0);
}

__global__ void __launch_bounds__(128) conv_const_taps_kernel(int cin, int cout, float slope,
                                                              const float *__restrict__ w /*[co][ci][27]*/,
                                                              const float *__restrict__ bias1, BnCoef bn1,
                                                              float *__restrict__ tapsum /*[27][cout]*/) {
  extern __shared__ float c1[];
  for (int i = threadIdx.x; i < cin; i += blockDim.x) c1[i] = leaky(fmaf(bias1[i], bn1.scale[i], bn1.shift[i]), slope);
  __syncthreads();
  const int tap = blockIdx.x;
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float t = 0.f;
    for (int ci = 0; ci < cin; ++ci) t = fmaf(w[((size_t)co * cin + ci) * 27 + tap], c1[ci], t);
    tapsum[(size_t)tap * cout + co] = t;
  }
}

__global__ void __launch_bounds__(128) conv_const_classes_kernel(int cout, int cp_out, const float *__restrict__ bias2,
                                                                 const float *__restrict__ tapsum,
                                                                 float *__restrict__ classsum) {
  const int cls = blockIdx.x;
  for (int co = threadIdx.x; co < cp_out; co += blockDim.x) {
    float acc = 0.f;
    if (co < cout) {
      acc = bias2[co];
      for (int tap = 0; tap < 27; ++tap)
        if (tap_valid_in_class(tap, cls)) acc += tapsum[(size_t)tap * cout + co];
    }
    classsum[(size_t)cls * cp_out + co] = acc;
  }
}

// one CTA per (b, x, y) line of r voxels: the boundary class only varies with z inside the line
__global__ void __launch_bounds__(256) fill_class_rows_kernel(int r, int cp, const float *__restrict__ classsum,
                                                              float *__restrict__ out) {
  const int cp4 = cp >> 2;
  long long line = blockIdx.x;
  const int y = (int)(line % r), x = (int)((line / r) % r);
  const int cxy = ((x == 0 ? 0 : (x == r - 1 ? 2 : 1)) * 3 + (y == 0 ? 0 : (y == r - 1 ? 2 : 1))) * 3;
  float *dst = out + (size_t)line * r * cp;
  for (int t = threadIdx.x; t < r * cp4; t += blockDim.x) {
    const int z = t / cp4, c4 = t - z * cp4;
    const int cls = cxy + (z == 0 ? 0 : (z == r - 1 ? 2 : 1));
    stg_stream4(dst + (size_t)t * 4, ld4(classsum + (size_t)cls * cp + c4 * 4));
  }
}

int launch_fill_const_conv(int nb, int r, int cin, int cout, int cp_out, float slope, const float *w, const float *bias2,
                           const float *bias1, BnCoef bn1, float *classsum, float *tapsum, float *out, cudaStream_t s,
                           int tables_ready) {
  if (!tables_ready) {   // the 27 class constants depend on the parameters only
    PVB_LAUNCH(conv_const_taps_kernel, 27, 128, cin * sizeof(float), s, cin, cout, slope, w, bias1, bn1, tapsum);
    PVB_LAUNCH(conv_const_classes_kernel, 27, 128, 0, s, cout, cp_out, bias2, tapsum, classsum);
  }
  PVB_LAUNCH(fill_class_rows_kernel, nb * r * r, 256, 0, s, r, cp_out, classsum, out);
  return 0;
}

// --------------------------------------------------------------------------------------------
// Weight gradient of conv2 over the region where its input is the constant c1 (k-tiles whose whole tap
// neighbourhood is constant):  dW2[co][ci][tap] += c1[ci] * sum_{v inactive, tap valid at v} gY2[v][co].
// Pass 1 reduces gY2 over the inactive k-tiles into the 27 boundary classes; pass 2 is the rank-1 update.
// --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) class_colsum_kernel(int r, int cp, int by, int bz, long long n_kt,
                                                           const unsigned char *__restrict__ kt_active,
                                                           const float *__restrict__ g,
                                                           float *__restrict__ classsum /*[2][27][cp] zeroed: inactive, all*/) {
  extern __shared__ float acc[];  // [2][27][cp]
  for (int i = threadIdx.x; i < 2 * 27 * cp; i += blockDim.x) acc[i] = 0.f;
  __syncthreads();
  const int cp4 = cp >> 2, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int tz_n = (r + bz - 1) / bz, ty_n = (r + by - 1) / by;
  float4 ri[3], ra[3];  // interior (x, y) classes in registers: [z class] of the inactive-only / all-voxel sets
#pragma unroll
  for (int k = 0; k < 3; ++k) ri[k] = ra[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  int reg_c4 = -1;      // a lane always works on the same channel quad when cp4 <= 32
  for (long long kt = (long long)blockIdx.x * nwarp + warp; kt < n_kt; kt += (long long)gridDim.x * nwarp) {
    const bool inactive = !kt_active[kt];
    long long u = kt;
    const int tz = (int)(u % tz_n); u /= tz_n;
    const int tyi = (int)(u % ty_n); u /= ty_n;
    const int x = (int)(u % r); u /= r;
    const int b = (int)u;
    const int cx = x == 0 ? 0 : (x == r - 1 ? 2 : 1);
    for (int yy = 0; yy < by; ++yy) {
      const int y = tyi * by + yy;
      if (y >= r) break;
      const int cy = y == 0 ? 0 : (y == r - 1 ? 2 : 1);
      // lanes: (z parity, channel quad) -- 32 lanes cover two z rows of 16 quads; loads unrolled for memory-level parallelism
      const int zpar = cp4 <= 16 ? (lane >> 4) : 0, zstep = cp4 <= 16 ? 2 : 1;
      for (int c4 = cp4 <= 16 ? (lane & 15) : lane; c4 < cp4; c4 += (cp4 <= 16 ? 16 : 32)) {
        // s1 collects EVERY row of the tile (straight-line adds, 8 loads in flight); the two boundary rows z = 0 and
        // z = r-1 are fetched once more (L1 hits) and moved to their own classes.  A three-way accumulator select
        // (`float4 &d = cond ? s0 : ...`) made ptxas spill the accumulators to local memory (r01: 1.5 TB/s).
        float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0, s2 = s0;
        const float *row = g + ((((size_t)b * r + x) * r + y) * r + (size_t)tz * bz) * cp + c4 * 4;
        const int zend = min(bz, r - tz * bz);
#pragma unroll 8
        for (int zz = zpar; zz < zend; zz += zstep) {
          const float4 v = ldg_stream4(row + (size_t)zz * cp);
          s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
        }
        if (tz == 0 && zpar == 0) {                       // z = 0 belongs to this lane's z parity
          s0 = ld4(row);
          s1.x -= s0.x; s1.y -= s0.y; s1.z -= s0.z; s1.w -= s0.w;
        }
        {
          const int zl = r - 1 - tz * bz;                 // local index of z = r-1, if it lies in this tile
          if (zl >= 0 && zl < zend && zl > 0 && (zl % zstep) == zpar) {
            s2 = ld4(row + (size_t)zl * cp);
            s1.x -= s2.x; s1.y -= s2.y; s1.z -= s2.z; s1.w -= s2.w;
          }
        }
        if (cx == 1 && cy == 1 && cp4 <= 32) {
          // interior (x, y): ~88 % of the lines share the classes (1,1,*) -> keep them in registers (r02: the 12-24
          // shared-memory atomics per lane and k-tile, not the 134 MB read, were what this kernel spent its time on)
#define PVB_ACC4(d, v) d.x += v.x; d.y += v.y; d.z += v.z; d.w += v.w;
          if (inactive) { PVB_ACC4(ri[0], s0) PVB_ACC4(ri[1], s1) PVB_ACC4(ri[2], s2) }
          PVB_ACC4(ra[0], s0) PVB_ACC4(ra[1], s1) PVB_ACC4(ra[2], s2)
#undef PVB_ACC4
          reg_c4 = c4;
        } else {
          for (int set = inactive ? 0 : 1; set < 2; ++set) {
            float *a0 = acc + ((size_t)set * 27 + (cx * 3 + cy) * 3) * cp + c4 * 4;
            float *a1 = a0 + cp, *a2 = a0 + 2 * cp;
            atomicAdd(a0 + 0, s0.x); atomicAdd(a0 + 1, s0.y); atomicAdd(a0 + 2, s0.z); atomicAdd(a0 + 3, s0.w);
            atomicAdd(a1 + 0, s1.x); atomicAdd(a1 + 1, s1.y); atomicAdd(a1 + 2, s1.z); atomicAdd(a1 + 3, s1.w);
            atomicAdd(a2 + 0, s2.x); atomicAdd(a2 + 1, s2.y); atomicAdd(a2 + 2, s2.z); atomicAdd(a2 + 3, s2.w);
          }
        }
      }
    }
  }
  if (reg_c4 >= 0) {  // flush the interior-class registers: one round of atomics per lane for the whole kernel
    for (int set = 0; set < 2; ++set) {
      const float4 *src = set == 0 ? ri : ra;
      float *a0 = acc + ((size_t)set * 27 + (1 * 3 + 1) * 3) * cp + reg_c4 * 4;
      for (int k = 0; k < 3; ++k) {
        atomicAdd(a0 + k * cp + 0, src[k].x); atomicAdd(a0 + k * cp + 1, src[k].y);
        atomicAdd(a0 + k * cp + 2, src[k].z); atomicAdd(a0 + k * cp + 3, src[k].w);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * 27 * cp; i += blockDim.x)
    if (acc[i] != 0.f) atomicAdd(classsum + i, acc[i]);
}

// total[ci] = sum over ALL voxels of the data gradient conv^T(g) = sum_tap sum_co w[co][ci][tap] * S_tap[co],
// S_tap[co] = sum_{classes where the tap stays in the grid} classsum_all[cls][co]
__global__ void __launch_bounds__(128) conv_grad_total_kernel(int cin, int cout, int cp,
                                                              const float *__restrict__ w /*[co][ci][27]*/,
                                                              const float *__restrict__ classsum_all,
                                                              float *__restrict__ total /*[cp], zeroed*/) {
  extern __shared__ float stap[];  // [cout]
  const int tap = blockIdx.x;
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float r = 0.f;
    for (int cls = 0; cls < 27; ++cls)
      if (tap_valid_in_class(tap, cls)) r += classsum_all[(size_t)cls * cp + co];
    stap[co] = r;
  }
  __syncthreads();
  for (int ci = threadIdx.x; ci < cin; ci += blockDim.x) {
    float t = 0.f;
    for (int co = 0; co < cout; ++co) t = fmaf(w[((size_t)co * cin + ci) * 27 + tap], stap[co], t);
    atomicAdd(total + ci, t);
  }
}

__global__ void __launch_bounds__(256) wgrad_const_update_kernel(int cin, int cout, int cp, float slope,
                                                                 const float *__restrict__ classsum,
                                                                 const float *__restrict__ bias1, BnCoef bn1,
                                                                 float *__restrict__ dw /*[co][ci][27]*/) {
  extern __shared__ float rsum[];  // [cout]: sum over the classes in which this tap stays inside the grid
  const int tap = blockIdx.x;
  for (int co = threadIdx.x; co < cout; co += blockDim.x) {
    float r = 0.f;
    for (int cls = 0; cls < 27; ++cls)
      if (tap_valid_in_class(tap, cls)) r += classsum[(size_t)cls * cp + co];
    rsum[co] = r;
  }
  __syncthreads();
  for (int t = threadIdx.x; t < cout * cin; t += blockDim.x) {
    const int co = t / cin, ci = t - co * cin;
    const float c1 = leaky(fmaf(bias1[ci], bn1.scale[ci], bn1.shift[ci]), slope);
    dw[((size_t)co * cin + ci) * 27 + tap] += c1 * rsum[co];
  }
}

int launch_class_sums(int nb, int r, int cp, int by, int bz, const unsigned char *kt_active, const float *g,
                      float *classsum2 /*[2][27][cp]: inactive k-tiles, all*/, cudaStream_t s) {
  PVB_CUDA(cudaMemsetAsync(classsum2, 0, sizeof(float) * 2 * 27 * (size_t)cp, s));
  const long long n_kt = (long long)nb * r * ((r + by - 1) / by) * ((r + bz - 1) / bz);
  const size_t smem = sizeof(float) * 2 * 27 * (size_t)cp;
  if (smem > 40 * 1024)  // static shared memory counts against the 48 KB default too
    PVB_CUDA(cudaFuncSetAttribute(class_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PVB_LAUNCH(class_colsum_kernel, kNumSMs * 2, 256, smem, s, r, cp, by, bz, n_kt, kt_active, g, classsum2);
  return 0;
}

int launch_wgrad_const_update(int cin, int cout, int cp, float slope, const float *classsum_inactive,
                              const float *bias1, BnCoef bn1, float *dw, cudaStream_t s) {
  PVB_LAUNCH(wgrad_const_update_kernel, 27, 256, cout * sizeof(float), s, cin, cout, cp, slope, classsum_inactive, bias1,
             bn1, dw);
  return 0;
}

int launch_conv_grad_total(int cin, int cout, int cp, const float *w, const float *classsum_all, float *total,
                           cudaStream_t s) {
  PVB_CUDA(cudaMemsetAsync(total, 0, sizeof(float) * (size_t)cp, s));
  PVB_LAUNCH(conv_grad_total_kernel, 27, 128, cout * sizeof(float), s, cin, cout, cp, w, classsum_all, total);
  return 0;
}

// --------------------------------------------------------------------------------------------
// BatchNorm1 backward restricted to a list of units (region G); the rest of the grid is the constant region
// (Y1 == b1), whose contribution is closed-form:   sum_C g = total - sum_G g,  leaky' = d_c,  xhat = xh_c.
// A unit = HC_TX(2) x-planes x ty y-rows x r z-rows (the halo conv kernel's output unit).
// --------------------------------------------------------------------------------------------
template <int NSETS, typename F>
__device__ __forceinline__ void unit_column_reduce(int r, int ty, int cp, const int4 *units, const int *count,
                                                   float *partials, F &&body) {
  __shared__ float4 red[NSETS][RED_THREADS];
  const int cp4 = cp >> 2, rl = RED_THREADS / cp4;
  const int c4 = threadIdx.x % cp4, lane_r = threadIdx.x / cp4;
  float4 acc[NSETS];
#pragma unroll
  for (int k = 0; k < NSETS; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  if ((int)blockIdx.x < __ldg(count) && lane_r < rl) {
    const int4 uc = __ldg(units + blockIdx.x);
    for (int xx = 0; xx < 2; ++xx) {
      const int x = uc.x + xx;
      if (x >= r) break;
      const int ny = min(ty, r - uc.y);
      const long long row0 = (((long long)uc.z * r + x) * r + uc.y) * r;
      for (int i = lane_r; i < ny * r; i += rl) body(row0 + i, c4, acc);
    }
  }
#pragma unroll
  for (int k = 0; k < NSETS; ++k) red[k][threadIdx.x] = acc[k];
  __syncthreads();
  if (threadIdx.x < cp4) {
#pragma unroll
    for (int k = 0; k < NSETS; ++k) {
      float4 sum = red[k][threadIdx.x];
      for (int j = 1; j < rl; ++j) {
        const float4 v = red[k][threadIdx.x + j * cp4];
        sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
      }
      st4(partials + ((size_t)blockIdx.x * NSETS + k) * cp + threadIdx.x * 4, sum);
    }
  }
}

// sets: 0 sum leaky'*g, 1 sum leaky'*g*xhat, 2 sum g, 3 number of rows
__global__ void __launch_bounds__(RED_THREADS) bn_bwd_reduce_units_kernel(int r, int ty, int cp, float slope,
                                                                          const int4 *__restrict__ units,
                                                                          const int *__restrict__ count,
                                                                          const float *__restrict__ g,
                                                                          const float *__restrict__ y, BnCoef coef,
                                                                          float *__restrict__ partials) {
  unit_column_reduce<4>(r, ty, cp, units, count, partials, [&](long long row, int c4, float4 *acc) {
    const size_t o = (size_t)row * cp + c4 * 4;
    const float4 gv = ld4(g + o), yv = ld4(y + o);
    const float4 sc = ld4(coef.scale + c4 * 4), sh = ld4(coef.shift + c4 * 4);
    const float4 mu = ld4(coef.mean + c4 * 4), is = ld4(coef.invstd + c4 * 4);
#define PVB_RU1(f)                                                             \
  {                                                                            \
    const float gg = gv.f * (fmaf(yv.f, sc.f, sh.f) > 0.f ? 1.0f : slope);     \
    acc[0].f += gg;                                                            \
    acc[1].f = fmaf(gg, (yv.f - mu.f) * is.f, acc[1].f);                       \
    acc[2].f += gv.f;                                                          \
    acc[3].f += 1.0f;                                                          \
  }
    PVB_RU1(x) PVB_RU1(y) PVB_RU1(z) PVB_RU1(w)
#undef PVB_RU1
  });
}

// U1 = sum_G m g + d_c (total - sum_G g),  U2 = sum_G m g xhat + d_c xh_c (total - sum_G g)
__global__ void __launch_bounds__(256) bn_bwd_combine_kernel(int c, int cp, float slope, const float *__restrict__ sums_g
                                                             /*[4][cp]*/, const float *__restrict__ total,
                                                             const float *__restrict__ bias_prev, BnCoef coef,
                                                             float *__restrict__ u1, float *__restrict__ u2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cp) return;
  if (i >= c) { u1[i] = 0.f; u2[i] = 0.f; return; }
  const float zc = fmaf(bias_prev[i], coef.scale[i], coef.shift[i]);
  const float dc = zc > 0.f ? 1.0f : slope;
  const float xhc = (bias_prev[i] - coef.mean[i]) * coef.invstd[i];
  const float rest = total[i] - sums_g[2 * cp + i];
  u1[i] = sums_g[0 * cp + i] + dc * rest;
  u2[i] = sums_g[1 * cp + i] + dc * xhc * rest;
}

__global__ void __launch_bounds__(RED_THREADS) bn_bwd_apply_units_kernel(int r, int ty, int cp, float slope,
                                                                         float inv_count,
                                                                         const int4 *__restrict__ units,
                                                                         const int *__restrict__ count,
                                                                         const float *__restrict__ g,
                                                                         const float *__restrict__ y, BnCoef coef,
                                                                         const float *__restrict__ s1,
                                                                         const float *__restrict__ s2,
                                                                         float *__restrict__ out,
                                                                         float *__restrict__ out_lo,
                                                                         float *__restrict__ partials) {
  unit_column_reduce<1>(r, ty, cp, units, count, partials, [&](long long row, int c4, float4 *acc) {
    const size_t o = (size_t)row * cp + c4 * 4;
    const float4 gv = ld4(g + o), yv = ld4(y + o);
    const float4 sc = ld4(coef.scale + c4 * 4), sh = ld4(coef.shift + c4 * 4);
    const float4 mu = ld4(coef.mean + c4 * 4), is = ld4(coef.invstd + c4 * 4);
    const float4 a1 = ld4(s1 + c4 * 4), a2 = ld4(s2 + c4 * 4);
    float4 d;
#define PVB_AU1(f)                                                             \
  {                                                                            \
    const float gg = gv.f * (fmaf(yv.f, sc.f, sh.f) > 0.f ? 1.0f : slope);     \
    const float xh = (yv.f - mu.f) * is.f;                                     \
    d.f = sc.f * (gg - a1.f * inv_count - xh * (a2.f * inv_count));            \
    acc[0].f += d.f;                                                           \
  }
    PVB_AU1(x) PVB_AU1(y) PVB_AU1(z) PVB_AU1(w)
#undef PVB_AU1
    st4(out + o, d);
    if (out_lo) st4(out_lo + o, tf32_lo4(d));
  });
}

// conv-bias gradient over the whole grid: explicit part (region G) + closed form on the constant region
__global__ void __launch_bounds__(256) bn_bwd_bias_total_kernel(int c, int cp, float slope, float rows_total,
                                                                const float *__restrict__ colsum_g /*sum_G dx*/,
                                                                const float *__restrict__ sums_g /*[4][cp]*/,
                                                                const float *__restrict__ total,
                                                                const float *__restrict__ bias_prev, BnCoef coef,
                                                                const float *__restrict__ u1,
                                                                const float *__restrict__ u2, float *__restrict__ db) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float zc = fmaf(bias_prev[i], coef.scale[i], coef.shift[i]);
  const float dc = zc > 0.f ? 1.0f : slope;
  const float xhc = (bias_prev[i] - coef.mean[i]) * coef.invstd[i];
  const float rest_g = total[i] - sums_g[2 * cp + i];
  const float rows_c = rows_total - sums_g[3 * cp + i];
  db[i] = colsum_g[i] + coef.scale[i] * (dc * rest_g - rows_c * (u1[i] / rows_total + xhc * (u2[i] / rows_total)));
}

int launch_bn_bwd_units(int max_units, int r, int ty, int c, int cp, float slope, long long rows_total,
                        const int4 *units, const int *count, const float *g, const float *y, BnCoef coef,
                        const float *total, const float *bias_prev, float *partials, float *sums_g /*[4][cp]*/,
                        float *u1, float *u2, float *out, float *out_lo, float *colsum_g /*[cp]*/, float *db,
                        cudaStream_t s) {
  PVB_CHECK_ARG(cp % 4 == 0 && cp / 4 <= RED_THREADS);
  PVB_LAUNCH(bn_bwd_reduce_units_kernel, max_units, RED_THREADS, 0, s, r, ty, cp, slope, units, count, g, y, coef,
             partials);
  PVB_TRY_LAUNCH(launch_reduce_partials(max_units, 4 * cp, partials, sums_g, s));
  PVB_LAUNCH(bn_bwd_combine_kernel, ceil_div(cp, 256), 256, 0, s, c, cp, slope, sums_g, total, bias_prev, coef, u1, u2);
  PVB_LAUNCH(bn_bwd_apply_units_kernel, max_units, RED_THREADS, 0, s, r, ty, cp, slope, (float)(1.0 / (double)rows_total),
             units, count, g, y, coef, u1, u2, out, out_lo, partials);
  PVB_TRY_LAUNCH(launch_reduce_partials(max_units, cp, partials, colsum_g, s));
  PVB_LAUNCH(bn_bwd_bias_total_kernel, ceil_div(c, 256), 256, 0, s, c, cp, slope, (float)rows_total, colsum_g, sums_g,
             total, bias_prev, coef, u1, u2, db);
  return 0;
}

int launch_memset_f32(float *p, long long n, cudaStream_t s) {
  PVB_CUDA(cudaMemsetAsync(p, 0, sizeof(float) * (size_t)n, s));
  return 0;
}

}  // namespace pvb
