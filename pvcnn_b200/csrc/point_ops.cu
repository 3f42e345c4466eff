// point_ops.cu -- sm_100a kernels behind the reference's launcher-level interface
// (mit-han-lab/pvcnn modules/functional/src/*/*.cuh), in the reference's own tensor layouts.
//
// Design differences from the reference kernels (which all launch <<<B, <=512>>>, i.e. use at most
// B of the 148 SMs): every kernel here tiles (batch x points x channels) so that the
// grid covers the whole chip, reads/writes are coalesced along the point dimension, scatter
// reductions are warp-aggregated before they reach L2, zero-fills are issued by the callee, and
// everything runs on the caller's stream.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace pvb {
std::atomic<unsigned long long> g_launches{0};

// =====================================================================================
// Coordinate normalisation  (modules/voxelization.py:16-25)
// One CTA per batch element: fp64 block reduction for the mean, NaN-propagating max-norm.
// =====================================================================================
template <typename T, typename Op>
__device__ __forceinline__ T block_allreduce(T v, Op op, T *smem /* >= 32 */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  T r = smem[0];
  for (int w = 1; w < nwarp; ++w) r = op(r, smem[w]);
  return r;
}

struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };
struct OpMaxNan {
  __device__ float operator()(float a, float b) const { return (a != a) ? a : ((b != b) ? b : fmaxf(a, b)); }
};

__global__ void __launch_bounds__(512) voxelize_coords_kernel(int n, int r, int normalize, float eps,
                                                              const float *__restrict__ coords,
                                                              float *__restrict__ norm_coords,
                                                              int *__restrict__ vox_coords) {
  __shared__ double sd[32];
  __shared__ float sf[32];
  const int b = blockIdx.x;
  const float *c = coords + (size_t)b * 3 * n;
  float mean[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)c[a * n + i];
    s = block_allreduce(s, OpAddD(), sd);
    mean[a] = (float)(s / (double)n);
  }
  float denom = 1.0f;
  if (normalize) {
    float mx = 0.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      float x = __fsub_rn(c[i], mean[0]), y = __fsub_rn(c[n + i], mean[1]), z = __fsub_rn(c[2 * n + i], mean[2]);
      float s = __fmul_rn(x, x);
      s = __fadd_rn(s, __fmul_rn(y, y));
      s = __fadd_rn(s, __fmul_rn(z, z));
      mx = OpMaxNan()(__fsqrt_rn(s), mx);
    }
    mx = block_allreduce(mx, OpMaxNan(), sf);
    denom = __fadd_rn(__fmul_rn(mx, 2.0f), eps);
  }
  const float rf = (float)r, hi = (float)(r - 1);
  for (int idx = threadIdx.x; idx < 3 * n; idx += blockDim.x) {
    const int a = idx / n;
    float v = __fsub_rn(c[idx], a == 0 ? mean[0] : (a == 1 ? mean[1] : mean[2]));
    if (normalize) v = __fadd_rn(__fdiv_rn(v, denom), 0.5f);
    else v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
    v = __fmul_rn(v, rf);
    if (v < 0.0f) v = 0.0f;
    if (v > hi) v = hi;
    norm_coords[(size_t)b * 3 * n + idx] = v;
    vox_coords[(size_t)b * 3 * n + idx] = __float2int_rn(v);
  }
}

// ---- reference-exact variant (default): the per-cloud mean comes from the SAME ATen reduction the reference runs
// (coords.mean(2), modules/voxelization.py:18 -- its summation order is an implementation detail of torch), everything
// after it is element-wise IEEE arithmetic or an order-independent max, reproduced here op by op without contraction:
//   norm = sqrt((x*x + y*y) + z*z)  (ATen's 3-element reduce: one accumulator per element, combined left to right)
//   denom = max_N(norm) * 2 + eps ;  v = (c - mean) / denom + 0.5  |  (c - mean + 1) * 0.5 ;  clamp(v * r, 0, r-1)
// One CTA per cloud for the max; the element-wise tail runs over the whole chip.
__global__ void __launch_bounds__(512) voxelize_denom_kernel(int n, float eps, const float *__restrict__ coords,
                                                             const float *__restrict__ mean,
                                                             float *__restrict__ denom) {
  __shared__ float sf[32];
  const int b = blockIdx.x;
  const float *c = coords + (size_t)b * 3 * n;
  const float m0 = mean[b * 3 + 0], m1 = mean[b * 3 + 1], m2 = mean[b * 3 + 2];
  float mx = -INFINITY;  // torch.max over non-negative norms; NaN propagates
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float x = __fsub_rn(c[i], m0), y = __fsub_rn(c[n + i], m1), z = __fsub_rn(c[2 * n + i], m2);
    float s = __fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y));
    s = __fadd_rn(s, __fmul_rn(z, z));
    mx = OpMaxNan()(__fsqrt_rn(s), mx);
  }
  mx = block_allreduce(mx, OpMaxNan(), sf);
  if (threadIdx.x == 0) denom[b] = __fadd_rn(__fmul_rn(mx, 2.0f), eps);
}

__global__ void __launch_bounds__(256) voxelize_apply_kernel(int n, int r, int normalize, long long total,
                                                             const float *__restrict__ coords,
                                                             const float *__restrict__ mean,
                                                             const float *__restrict__ denom,
                                                             float *__restrict__ norm_coords,
                                                             int *__restrict__ vox_coords) {
  const float rf = (float)r, hi = (float)(r - 1);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int ba = (int)(t / n);  // b * 3 + axis
    float v = __fsub_rn(coords[t], __ldg(mean + ba));
    if (normalize) v = __fadd_rn(__fdiv_rn(v, __ldg(denom + ba / 3)), 0.5f);
    else v = __fmul_rn(__fadd_rn(v, 1.0f), 0.5f);
    v = __fmul_rn(v, rf);
    if (v < 0.0f) v = 0.0f;   // torch.clamp: NaN stays NaN
    if (v > hi) v = hi;
    norm_coords[t] = v;
    vox_coords[t] = __float2int_rn(v);  // torch.round (half to even) + .to(int32)
  }
}

// =====================================================================================
// avg_voxelize  (vox.cu:18-72).  K1: voxel index + count.  K2: scatter-mean.
// =====================================================================================
__global__ void __launch_bounds__(256) vox_index_count_kernel(int n, int r, int r2, int r3, long long total,
                                                              const int *__restrict__ coords,
                                                              int *__restrict__ ind, int *__restrict__ cnt) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int b = (int)(t / n), i = (int)(t % n);
    const int *co = coords + (size_t)b * 3 * n;
    const int v = co[i] * r2 + co[i + n] * r + co[i + 2 * n];
    ind[t] = v;
    atomicAdd(cnt + (size_t)b * r3 + v, 1);
  }
}

// Sum of v over the lanes of `grp` in ascending lane order; every lane of the group gets the total.
__device__ __forceinline__ float group_sum_ordered(unsigned grp, float v) {
  float acc = 0.0f;
  for (unsigned m = grp; m; m &= m - 1) acc = __fadd_rn(acc, __shfl_sync(grp, v, __ffs(m) - 1));
  return acc;
}

template <int CT>
__global__ void __launch_bounds__(128) vox_scatter_kernel(int c, int n, int r3, const int *__restrict__ ind,
                                                          const int *__restrict__ cnt,
                                                          const float *__restrict__ feat,
                                                          float *__restrict__ out) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT, lane = threadIdx.x & 31;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = i < n;
  const int pos = valid ? ind[(size_t)b * n + i] : -1 - lane;
  // lanes that hit the same voxel are pre-reduced in registers; one atomic per (voxel, channel)
  const unsigned grp = __match_any_sync(0xffffffffu, pos);
  const bool leader = (__ffs(grp) - 1) == lane;
  const bool multi = __any_sync(0xffffffffu, grp != (1u << lane));
  float inv = 0.0f;
  if (valid) inv = (float)(1.0 / (double)(float)cnt[(size_t)b * r3 + pos]);  // vox.cu:66
  const float *f = feat + ((size_t)b * c + c0) * n + i;
  float *o = out + ((size_t)b * c + c0) * r3 + pos;
  const int cmax = min(CT, c - c0);
  for (int j = 0; j < cmax; ++j) {
    float v = valid ? __fmul_rn(f[(size_t)j * n], inv) : 0.0f;
    if (multi) v = group_sum_ordered(grp, v);
    if (valid && leader) atomicAdd(o + (size_t)j * r3, v);
  }
}

template <int CT>
__global__ void __launch_bounds__(128) vox_grad_kernel(int c, int n, int r3, const int *__restrict__ ind,
                                                       const int *__restrict__ cnt,
                                                       const float *__restrict__ grad_y,
                                                       float *__restrict__ grad_x) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int pos = ind[(size_t)b * n + i];
  const int cur = cnt[(size_t)b * r3 + pos];
  const float inv = cur > 0 ? (float)(1.0 / (double)(float)cur) : 0.0f;
  const float *gy = grad_y + ((size_t)b * c + c0) * r3 + pos;
  float *gx = grad_x + ((size_t)b * c + c0) * n + i;
  const int cmax = min(CT, c - c0);
#pragma unroll 4
  for (int j = 0; j < cmax; ++j) gx[(size_t)j * n] = cur > 0 ? __fmul_rn(__ldg(gy + (size_t)j * r3), inv) : 0.0f;
}

// =====================================================================================
// trilinear_devoxelize  (trilinear_devox.cu:21-162)
// =====================================================================================
struct Corner8 {
  float w[8];
  int idx[8];
};
__device__ __forceinline__ void devox_setup(float x, float y, float z, int r, int r2, Corner8 &k) {
  const float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  const float xd1 = __fsub_rn(x, xl), yd1 = __fsub_rn(y, yl), zd1 = __fsub_rn(z, zl);
  const float xd0 = __fsub_rn(1.0f, xd1), yd0 = __fsub_rn(1.0f, yd1), zd0 = __fsub_rn(1.0f, zd1);
  const float a00 = __fmul_rn(xd0, yd0), a01 = __fmul_rn(xd0, yd1), a10 = __fmul_rn(xd1, yd0),
              a11 = __fmul_rn(xd1, yd1);
  k.w[0] = __fmul_rn(a00, zd0); k.w[1] = __fmul_rn(a00, zd1);
  k.w[2] = __fmul_rn(a01, zd0); k.w[3] = __fmul_rn(a01, zd1);
  k.w[4] = __fmul_rn(a10, zd0); k.w[5] = __fmul_rn(a10, zd1);
  k.w[6] = __fmul_rn(a11, zd0); k.w[7] = __fmul_rn(a11, zd1);
  const int xlo = (int)xl, ylo = (int)yl, zlo = (int)zl;
  const int dz = zd1 > 0 ? 1 : 0, dy = yd1 > 0 ? r : 0, dx = xd1 > 0 ? r2 : 0;  // trilinear_devox.cu:64-75
  k.idx[0] = xlo * r2 + ylo * r + zlo;
  k.idx[1] = k.idx[0] + dz;
  k.idx[2] = k.idx[0] + dy;
  k.idx[3] = k.idx[2] + dz;
  k.idx[4] = k.idx[0] + dx;
  k.idx[5] = k.idx[4] + dz;
  k.idx[6] = k.idx[4] + dy;
  k.idx[7] = k.idx[6] + dz;
}

template <int CT>
__global__ void __launch_bounds__(128) devox_kernel(int c, int n, int r, int r2, int r3, int training,
                                                    const float *__restrict__ coords,
                                                    const float *__restrict__ feat, int *__restrict__ inds,
                                                    float *__restrict__ wgts, float *__restrict__ outs) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float *co = coords + (size_t)b * 3 * n;
  Corner8 k;
  devox_setup(co[i], co[i + n], co[i + 2 * n], r, r2, k);
  if (training && blockIdx.y == 0) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      wgts[((size_t)b * 8 + q) * n + i] = k.w[q];
      inds[((size_t)b * 8 + q) * n + i] = k.idx[q];
    }
  }
  const float *f = feat + ((size_t)b * c + c0) * r3;
  float *o = outs + ((size_t)b * c + c0) * n + i;
  const int cmax = min(CT, c - c0);
#pragma unroll 2
  for (int j = 0; j < cmax; ++j) {
    const float *fj = f + (size_t)j * r3;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(fj + k.idx[q]);
    float acc = __fmul_rn(k.w[0], v[0]);
#pragma unroll
    for (int q = 1; q < 8; ++q) acc = __fmaf_rn(k.w[q], v[q], acc);
    o[(size_t)j * n] = acc;
  }
}

template <int CT>
__global__ void __launch_bounds__(128) devox_grad_kernel(int c, int n, int r3, const int *__restrict__ inds,
                                                         const float *__restrict__ wgts,
                                                         const float *__restrict__ grad_y,
                                                         float *__restrict__ grad_x) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int idx[8];
  float w[8];
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    idx[q] = inds[((size_t)b * 8 + q) * n + i];
    w[q] = wgts[((size_t)b * 8 + q) * n + i];
  }
  const float *gy = grad_y + ((size_t)b * c + c0) * n + i;
  float *gx = grad_x + ((size_t)b * c + c0) * r3;
  const int cmax = min(CT, c - c0);
  for (int j = 0; j < cmax; ++j) {
    const float g = gy[(size_t)j * n];
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(gx + (size_t)j * r3 + idx[q], __fmul_rn(w[q], g));
  }
}

// =====================================================================================
// ball_query  (ball_query.cu:19-50): warp per query over a shared-memory point tile.
// Ordered ballot compaction keeps the reference's "first U by index" semantics; the final
// row is hits[0..cnt) followed by hits[0] repeated (or zeros when there is no hit).
// =====================================================================================
constexpr int BQ_WARPS = 8, BQ_QPW = 4, BQ_TILE = 2048;

__global__ void __launch_bounds__(BQ_WARPS * 32) ball_query_kernel(int n, int m, float r2, int u,
                                                                   const float *__restrict__ centers,
                                                                   const float *__restrict__ points,
                                                                   int *__restrict__ out) {
  __shared__ float sp[3][BQ_TILE];
  const int b = blockIdx.y, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float *p = points + (size_t)b * 3 * n;
  const float *ce = centers + (size_t)b * 3 * m;
  const int q0 = (blockIdx.x * BQ_WARPS + warp) * BQ_QPW;
  float cx[BQ_QPW], cy[BQ_QPW], cz[BQ_QPW];
  int cnt[BQ_QPW], first[BQ_QPW];
#pragma unroll
  for (int q = 0; q < BQ_QPW; ++q) {
    const int j = q0 + q;
    cx[q] = j < m ? ce[j] : 0.f;
    cy[q] = j < m ? ce[j + m] : 0.f;
    cz[q] = j < m ? ce[j + 2 * m] : 0.f;
    cnt[q] = j < m ? 0 : u;  // out-of-range queries are "done"
    first[q] = 0;
  }
  for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
    bool warp_done = true;
#pragma unroll
    for (int q = 0; q < BQ_QPW; ++q) warp_done = warp_done && (cnt[q] >= u);
    if (__syncthreads_and(warp_done)) break;  // also fences the previous tile's readers
    const int tn = min(BQ_TILE, n - t0);
    for (int k = threadIdx.x; k < tn; k += blockDim.x) {
      sp[0][k] = p[t0 + k];
      sp[1][k] = p[t0 + k + n];
      sp[2][k] = p[t0 + k + 2 * n];
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BQ_QPW; ++q) {
      if (cnt[q] >= u) continue;  // warp-uniform
      int *row = out + ((size_t)b * m + (q0 + q)) * u;
      for (int base = 0; base < tn && cnt[q] < u; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < tn) hit = sqdist(cx[q] - sp[0][k], cy[q] - sp[1][k], cz[q] - sp[2][k]) < r2;
        const unsigned bal = __ballot_sync(0xffffffffu, hit);
        if (bal) {
          if (cnt[q] == 0) first[q] = t0 + base + __ffs(bal) - 1;
          const int slot = cnt[q] + __popc(bal & ((1u << lane) - 1));
          if (hit && slot < u) row[slot] = t0 + k;
          cnt[q] += __popc(bal);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < BQ_QPW; ++q) {
    const int j = q0 + q;
    if (j >= m) continue;
    int *row = out + ((size_t)b * m + j) * u;
    const int have = min(cnt[q], u);
    for (int v = have + lane; v < u; v += 32) row[v] = first[q];  // first==0 when no hit at all
  }
}

// =====================================================================================
// grouping / gather  (grouping.cu:18-77, sampling.cu:17-66)
// =====================================================================================
template <int CT>
__global__ void __launch_bounds__(256) grouping_kernel(int c, int n, int mu, const float *__restrict__ feat,
                                                       const int *__restrict__ idx, float *__restrict__ out) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= mu) return;
  const int src = idx[(size_t)b * mu + e];
  const float *f = feat + ((size_t)b * c + c0) * n + src;
  float *o = out + ((size_t)b * c + c0) * mu + e;
  const int cmax = min(CT, c - c0);
#pragma unroll 4
  for (int j = 0; j < cmax; ++j) o[(size_t)j * mu] = __ldg(f + (size_t)j * n);
}

template <int CT>
__global__ void __launch_bounds__(256) grouping_grad_kernel(int c, int n, int mu, const float *__restrict__ grad_y,
                                                            const int *__restrict__ idx,
                                                            float *__restrict__ grad_x) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT, lane = threadIdx.x & 31;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e < mu;
  const int dst = valid ? idx[(size_t)b * mu + e] : -1 - lane;
  // ball_query pads rows with the first hit, so neighbouring lanes very often share dst
  const unsigned grp = __match_any_sync(0xffffffffu, dst);
  const bool leader = (__ffs(grp) - 1) == lane;
  const bool multi = __any_sync(0xffffffffu, grp != (1u << lane));
  const float *gy = grad_y + ((size_t)b * c + c0) * mu + e;
  float *gx = grad_x + ((size_t)b * c + c0) * n + dst;
  const int cmax = min(CT, c - c0);
  for (int j = 0; j < cmax; ++j) {
    float v = valid ? gy[(size_t)j * mu] : 0.0f;
    if (multi) v = group_sum_ordered(grp, v);
    if (valid && leader) atomicAdd(gx + (size_t)j * n, v);
  }
}

// =====================================================================================
// furthest point sampling  (sampling.cu:86-167)
// One CTA of 1024 threads per batch element; points and running distances live in registers
// (PPT points per thread), one __syncthreads per round, integer REDUX reductions.
// Tie-break = the reference's: max distance, then smallest (k mod 512), then smallest k.
// =====================================================================================
constexpr int FPS_THREADS = 1024;

__device__ __forceinline__ unsigned fps_tie_key(int k) { return ((unsigned)(k & 511) << 22) | ((unsigned)k >> 9); }
__device__ __forceinline__ int fps_key_to_k(unsigned key) { return (int)(((key & 0x3FFFFFu) << 9) | (key >> 22)); }

template <int PPT, int NT>
__global__ void __launch_bounds__(NT, 1) fps_kernel(int n, int m, int coords_in_smem,
                                                             const float *__restrict__ coords,
                                                             int *__restrict__ indices) {
  extern __shared__ float s_xyz[];  // [3][n] when coords_in_smem
  __shared__ unsigned s_d[2][32], s_k[2][32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *co = coords + (size_t)b * 3 * n;
  int *out = indices + (size_t)b * m;
  float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int k = tid + p * NT;
    const bool v = k < n;
    px[p] = v ? co[k] : 0.f;
    py[p] = v ? co[k + n] : 0.f;
    pz[p] = v ? co[k + 2 * n] : 0.f;
    dist[p] = v ? 1e38f : 0.f;   // slots past the end never win (see the round loop)
    if (coords_in_smem && v) {
      s_xyz[k] = px[p];
      s_xyz[n + k] = py[p];
      s_xyz[2 * n + k] = pz[p];
    }
  }
  if (tid < 64) { s_d[tid >> 5][tid & 31] = 0u; s_k[tid >> 5][tid & 31] = 0xffffffffu; }  // rows of absent warps never win
  if (tid == 0) out[0] = 0;
  __syncthreads();
  int old = 0;
  for (int j = 1; j < m; ++j) {
    float x1, y1, z1;
    if (coords_in_smem) {
      x1 = s_xyz[old]; y1 = s_xyz[n + old]; z1 = s_xyz[2 * n + old];
    } else {
      x1 = __ldg(co + old); y1 = __ldg(co + old + n); z1 = __ldg(co + old + 2 * n);
    }
    // Per-thread best under (dist desc, tie key asc).  Slots past the end of the cloud carry distance 0 and can never
    // win the strict '>' below, so the loop has no bounds test.  A thread's points share k mod 512 (the stride is a
    // multiple of 512) and k ascends with p, so strict '>' keeps its smallest k  (sampling.cu:141-144) and the tie key
    // of slot p is key(slot 0) + p * (stride >> 9).
    unsigned bd;
    int bp = 0;
    {
      const float d2 = fminf(sqdist(px[0] - x1, py[0] - y1, pz[0] - z1), dist[0]);
      dist[0] = d2;
      bd = __float_as_uint(d2);  // d2 >= 0 (or +0): monotone as unsigned
    }
#pragma unroll
    for (int p = 1; p < PPT; ++p) {
      const float d2 = fminf(sqdist(px[p] - x1, py[p] - y1, pz[p] - z1), dist[p]);
      dist[p] = d2;
      const unsigned db = __float_as_uint(d2);
      if (db > bd) { bd = db; bp = p; }
    }
    const bool any = tid < n;
    const unsigned bk = fps_tie_key(tid) + (unsigned)bp * (NT >> 9);
    // warp argmax: max distance bits, then min tie key among the maxima
    unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    unsigned wk = __reduce_min_sync(0xffffffffu, (bd == wd && any) ? bk : 0xffffffffu);
    const int buf = j & 1;
    if (lane == 0) { s_d[buf][warp] = wd; s_k[buf][warp] = wk; }
    __syncthreads();
    const unsigned d_l = s_d[buf][lane], k_l = s_k[buf][lane];
    const unsigned gd = __reduce_max_sync(0xffffffffu, d_l);
    const unsigned gk = __reduce_min_sync(0xffffffffu, d_l == gd ? k_l : 0xffffffffu);
    old = (gk == 0xffffffffu) ? 0 : fps_key_to_k(gk);
    if (tid == 0) out[j] = old;
  }
}

// generic fallback for very large N: distances in global scratch
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_kernel_big(int n, int m, const float *__restrict__ coords,
                                                                 float *__restrict__ distances,
                                                                 int *__restrict__ indices) {
  __shared__ unsigned s_d[2][32], s_k[2][32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *co = coords + (size_t)b * 3 * n;
  float *dist = distances + (size_t)b * n;
  int *out = indices + (size_t)b * m;
  for (int k = tid; k < n; k += FPS_THREADS) dist[k] = 1e38f;
  if (tid == 0) out[0] = 0;
  __syncthreads();
  int old = 0;
  for (int j = 1; j < m; ++j) {
    const float x1 = __ldg(co + old), y1 = __ldg(co + old + n), z1 = __ldg(co + old + 2 * n);
    unsigned bd = 0u, bk = fps_tie_key(0);
    bool any = false;
    for (int k = tid; k < n; k += FPS_THREADS) {
      const float d = sqdist(__ldg(co + k) - x1, __ldg(co + k + n) - y1, __ldg(co + k + 2 * n) - z1);
      const float d2 = fminf(d, dist[k]);
      dist[k] = d2;
      const unsigned db = __float_as_uint(d2);
      if (!any || db > bd) { bd = db; bk = fps_tie_key(k); any = true; }
    }
    unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    unsigned wk = __reduce_min_sync(0xffffffffu, (bd == wd && any) ? bk : 0xffffffffu);
    const int buf = j & 1;
    if (lane == 0) { s_d[buf][warp] = wd; s_k[buf][warp] = wk; }
    __syncthreads();
    const unsigned d_l = s_d[buf][lane], k_l = s_k[buf][lane];
    const unsigned gd = __reduce_max_sync(0xffffffffu, d_l);
    const unsigned gk = __reduce_min_sync(0xffffffffu, d_l == gd ? k_l : 0xffffffffu);
    old = (gk == 0xffffffffu) ? 0 : fps_key_to_k(gk);
    if (tid == 0) out[j] = old;
  }
}

// =====================================================================================
// three-NN search + interpolation  (neighbor_interpolate.cu:20-170)
// =====================================================================================
constexpr int NN_TILE = 2048;
__global__ void __launch_bounds__(128) three_nn_kernel(int n, int m, const float *__restrict__ points,
                                                       const float *__restrict__ centers,
                                                       float *__restrict__ weights, int *__restrict__ indices) {
  __shared__ float sc[3][NN_TILE];
  const int b = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const float *p = points + (size_t)b * 3 * n;
  const float *ce = centers + (size_t)b * 3 * m;
  const bool valid = j < n;
  const float ux = valid ? p[j] : 0.f, uy = valid ? p[j + n] : 0.f, uz = valid ? p[j + 2 * n] : 0.f;
  double best0 = 1e40, best1 = 1e40, best2 = 1e40;
  int i0 = 0, i1 = 0, i2 = 0;
  for (int t0 = 0; t0 < m; t0 += NN_TILE) {
    const int tn = min(NN_TILE, m - t0);
    __syncthreads();
    for (int k = threadIdx.x; k < tn; k += blockDim.x) {
      sc[0][k] = ce[t0 + k];
      sc[1][k] = ce[t0 + k + m];
      sc[2][k] = ce[t0 + k + 2 * m];
    }
    __syncthreads();
    for (int k = 0; k < tn; ++k) {
      const float d = sqdist(ux - sc[0][k], uy - sc[1][k], uz - sc[2][k]);
      const double dd = (double)d;
      if (dd < best2) {
        best2 = dd; i2 = t0 + k;
        if (dd < best1) {
          best2 = best1; i2 = i1; best1 = dd; i1 = t0 + k;
          if (dd < best0) { best1 = best0; i1 = i0; best0 = dd; i0 = t0 + k; }
        }
      }
    }
  }
  if (!valid) return;
  best0 = fmax(fmin((double)1e10f, best0), (double)1e-10f);
  best1 = fmax(fmin((double)1e10f, best1), (double)1e-10f);
  best2 = fmax(fmin((double)1e10f, best2), (double)1e-10f);
  const float d0d1 = (float)(best0 * best1), d0d2 = (float)(best0 * best2), d1d2 = (float)(best1 * best2);
  const float inv = __fdiv_rn(1.0f, __fadd_rn(__fadd_rn(d0d1, d0d2), d1d2));
  float *w = weights + (size_t)b * 3 * n;
  int *id = indices + (size_t)b * 3 * n;
  w[j] = __fmul_rn(d1d2, inv);          id[j] = i0;
  w[j + n] = __fmul_rn(d0d2, inv);      id[j + n] = i1;
  w[j + 2 * n] = __fmul_rn(d0d1, inv);  id[j + 2 * n] = i2;
}

template <int CT>
__global__ void __launch_bounds__(128) three_nn_interp_kernel(int c, int m, int n, const float *__restrict__ cfeat,
                                                              const int *__restrict__ indices,
                                                              const float *__restrict__ weights,
                                                              float *__restrict__ out) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *id = indices + (size_t)b * 3 * n;
  const float *w = weights + (size_t)b * 3 * n;
  const int i1 = id[j], i2 = id[j + n], i3 = id[j + 2 * n];
  const float w1 = w[j], w2 = w[j + n], w3 = w[j + 2 * n];
  const float *f = cfeat + ((size_t)b * c + c0) * m;
  float *o = out + ((size_t)b * c + c0) * n + j;
  const int cmax = min(CT, c - c0);
#pragma unroll 4
  for (int l = 0; l < cmax; ++l) {
    const float *fl = f + (size_t)l * m;
    o[(size_t)l * n] = __fmaf_rn(__ldg(fl + i3), w3, __fmaf_rn(__ldg(fl + i2), w2, __fmul_rn(__ldg(fl + i1), w1)));
  }
}

template <int CT>
__global__ void __launch_bounds__(128) three_nn_interp_grad_kernel(int c, int n, int m,
                                                                   const float *__restrict__ grad_y,
                                                                   const int *__restrict__ indices,
                                                                   const float *__restrict__ weights,
                                                                   float *__restrict__ grad_x) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  const int *id = indices + (size_t)b * 3 * n;
  const float *w = weights + (size_t)b * 3 * n;
  const int i1 = id[j], i2 = id[j + n], i3 = id[j + 2 * n];
  const float w1 = w[j], w2 = w[j + n], w3 = w[j + 2 * n];
  const float *gy = grad_y + ((size_t)b * c + c0) * n + j;
  float *gx = grad_x + ((size_t)b * c + c0) * m;
  const int cmax = min(CT, c - c0);
  for (int l = 0; l < cmax; ++l) {
    const float g = gy[(size_t)l * n];
    float *gl = gx + (size_t)l * m;
    atomicAdd(gl + i1, __fmul_rn(g, w1));
    atomicAdd(gl + i2, __fmul_rn(g, w2));
    atomicAdd(gl + i3, __fmul_rn(g, w3));
  }
}


// =====================================================================================
// logits_mask resampling on the device (modules/functional/sampling.py:66-82).  The reference loops over the
// batch on the host: mask[i].nonzero() (a device->host sync per sample), np.random.choice / shuffle, a copy
// back.  Here one CTA per sample does: ordered compaction of the foreground candidates (block scan), a random
// k-subset / repeat-fill + shuffle with a counter-based generator (splitmix64 of (seed, sample, stream, index)),
// realised as bitonic sorts of 64-bit (random key, index) pairs in shared memory.  Same distribution as the
// reference's numpy calls (uniform subsets / permutations); no host round trip.
//   smem: keys[P] (uint64) + cand[n] (int) + extras[n] (int)
// =====================================================================================
__device__ __forceinline__ unsigned long long lm_mix(unsigned long long seed, unsigned long long b,
                                                     unsigned long long stream, unsigned long long j) {
  unsigned long long x = seed ^ (b * 0x9E3779B97F4A7C15ull) ^ (stream * 0xBF58476D1CE4E5B9ull) ^ (j * 0x94D049BB133111EBull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

__device__ void lm_bitonic_sort(unsigned long long *keys, int p2) {
  for (int size = 2; size <= p2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < (p2 >> 1); t += blockDim.x) {
        const int lo = 2 * t - (t & (stride - 1));  // index of the lower element of pair t
        const int hi = lo + stride;
        const bool up = (lo & size) == 0;
        const unsigned long long a = keys[lo], c = keys[hi];
        if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(1024) logits_mask_sample_kernel(int n, int k, unsigned long long seed,
                                                                  const unsigned char *__restrict__ mask,
                                                                  int *__restrict__ picks) {
  extern __shared__ __align__(16) unsigned char lm_smem[];
  __shared__ int warp_tot[32];
  __shared__ int s_nc;
  int p2max = 1;
  while (p2max < max(n, k)) p2max <<= 1;
  unsigned long long *keys = reinterpret_cast<unsigned long long *>(lm_smem);
  int *cand = reinterpret_cast<int *>(keys + p2max);
  int *extras = cand + n;
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned char *mk = mask + (size_t)b * n;
  // ---- ordered compaction: thread t owns the contiguous chunk [t*per, (t+1)*per)
  const int per = (n + blockDim.x - 1) / blockDim.x;
  const int i0 = tid * per, i1 = min(n, i0 + per);
  int local = 0;
  for (int i = i0; i < i1; ++i) local += mk[i] != 0;
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int v = lane < (int)(blockDim.x >> 5) ? warp_tot[lane] : 0;
    int inc2 = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int w = __shfl_up_sync(0xffffffffu, inc2, o);
      if (lane >= o) inc2 += w;
    }
    warp_tot[lane] = inc2 - v;  // exclusive warp offsets
    if (lane == 31) s_nc = inc2;
  }
  __syncthreads();
  int pos = warp_tot[warp] + incl - local;
  for (int i = i0; i < i1; ++i)
    if (mk[i]) cand[pos++] = i;
  const int nc = s_nc;
  __syncthreads();
  int *out = picks + (size_t)b * k;
  if (nc == 0) {  // sampling.py:66: selected_indices stays zero
    for (int t = tid; t < k; t += blockDim.x) out[t] = 0;
    return;
  }
  // ---- stream 0: random order of the candidates (its prefix is a uniform subset without replacement)
  int p2 = 1;
  while (p2 < nc) p2 <<= 1;
  for (int j = tid; j < p2; j += blockDim.x)
    keys[j] = j < nc ? ((lm_mix(seed, b, 0, j) & 0xFFFFFFFF00000000ull) | (unsigned long long)j) : ~0ull;
  lm_bitonic_sort(keys, p2);
  if (nc >= k) {  // sampling.py:71-73
    for (int t = tid; t < k; t += blockDim.x) out[t] = cand[(int)(keys[t] & 0xFFFFFFFFull)];
    return;
  }
  // ---- fewer candidates than k (sampling.py:74-80): every candidate k/nc times, k%nc distinct extras, shuffled
  const int rep = k / nc, rem = k - rep * nc;
  for (int t = tid; t < rem; t += blockDim.x) extras[t] = (int)(keys[t] & 0xFFFFFFFFull);
  __syncthreads();
  p2 = 1;
  while (p2 < k) p2 <<= 1;
  for (int t = tid; t < p2; t += blockDim.x)
    keys[t] = t < k ? ((lm_mix(seed, b, 1, t) & 0xFFFFFFFF00000000ull) | (unsigned long long)t) : ~0ull;
  lm_bitonic_sort(keys, p2);
  for (int t = tid; t < k; t += blockDim.x) {
    const int src = (int)(keys[t] & 0xFFFFFFFFull);                 // position in the un-shuffled list
    const int ci = src < rep * nc ? src / rep : extras[src - rep * nc];  // arange(nc).repeat(rep) ++ extras
    out[t] = cand[ci];
  }
}
}  // namespace pvb

// =====================================================================================
// C ABI
// =====================================================================================
using namespace pvb;

// Cluster variant: CS CTAs (one thread-block cluster) share a cloud.  Every CTA keeps the whole cloud's coordinates in its
// shared memory (to read the last pick) and owns n/CS of the running distances in registers.  A round has NO block or
// cluster barrier: each warp reduces its own points with REDUX, lane r (< CS) sends the warp's (distance bits, tie key)
// to CTA r with st.async, which lands the 8 bytes in that CTA's candidate table and signals its transaction mbarrier;
// every warp then waits on the LOCAL mbarrier (32*CS candidates = 256*CS bytes per round), reads the table (CS entries per
// lane) and reduces again, so all warps of all CTAs pick the same winner.  Tables and mbarriers are double buffered; a
// warp can only be one round ahead of any other warp of the cluster, because it needs everybody's candidate to go on.
// (the M-round chain is the largest non-conv cost of PVCNN++.)  The result is index-identical to fps_kernel:
// (max distance, min tie key) does not depend on how the points are partitioned.
__device__ __forceinline__ uint32_t fps_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}

template <int PPT, int CS, int NT>
__global__ void __launch_bounds__(NT, 1) fps_cluster_kernel(int n, int m, const float *__restrict__ coords,
                                                                     int *__restrict__ indices) {
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  constexpr int NW = NT / 32;   // warps per CTA = candidates per CTA and round
  extern __shared__ float s_xyz[];  // [3][n]
  __shared__ __align__(8) uint2 s_cand[2][NW * CS];   // [parity][source rank * 32 + source warp] = (distance bits, tie key)
  __shared__ __align__(8) uint64_t s_bar[2];
  static_assert((CS * NT) % 512 == 0 && NW * CS <= 32 * 8, "tie rule needs a stride that is a multiple of 512");
  const uint32_t rank = cluster.block_rank();
  const int b = blockIdx.x / CS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float *co = coords + (size_t)b * 3 * n;
  int *out = indices + (size_t)b * m;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&s_bar[0]);
  const uint32_t cand0 = (uint32_t)__cvta_generic_to_shared(&s_cand[0][0]);
  constexpr uint32_t ROUND_BYTES = (uint32_t)NW * CS * 8u;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // rounds 0 and 1 are armed here, round t + 2 right after round t completed
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0), "r"(ROUND_BYTES) : "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8), "r"(ROUND_BYTES) : "memory");
  }
  for (int k = tid; k < n; k += NT) {
    s_xyz[k] = co[k];
    s_xyz[n + k] = co[k + n];
    s_xyz[2 * n + k] = co[k + 2 * n];
  }
  float px[PPT], py[PPT], pz[PPT], dist[PPT];
#pragma unroll
  for (int p = 0; p < PPT; ++p) {
    const int k = (int)rank * NT + tid + p * (CS * NT);  // stride is a multiple of 512: see the tie rule
    const bool v = k < n;
    px[p] = v ? co[k] : 0.f;
    py[p] = v ? co[k + n] : 0.f;
    pz[p] = v ? co[k + 2 * n] : 0.f;
    dist[p] = v ? 1e38f : 0.f;
  }
  const int k0 = (int)rank * NT + tid;
  const bool any = k0 < n;
  const unsigned key0 = fps_tie_key(k0);
  if (rank == 0 && tid == 0) out[0] = 0;
  __syncthreads();
  cluster.sync();   // every CTA of the cluster is resident and its mbarriers are initialised before the first st.async
  // lane r < CS addresses CTA r: this warp's slot in that CTA's table, and that CTA's mbarriers
  const uint32_t peer_cand = fps_mapa(cand0 + (rank * (uint32_t)NW + (uint32_t)warp) * 8u, lane < CS ? lane : 0);
  const uint32_t peer_bar = fps_mapa(bar0, lane < CS ? lane : 0);
  int old = 0;
  for (int t = 0; t + 1 < m; ++t) {
    const uint32_t buf = t & 1, parity = (t >> 1) & 1;
    const float x1 = s_xyz[old], y1 = s_xyz[n + old], z1 = s_xyz[2 * n + old];
    unsigned bd;          // same per-thread rule as fps_kernel: no bounds test, strict '>' keeps the smallest k
    int bp = 0;
    {
      const float d2 = fminf(sqdist(px[0] - x1, py[0] - y1, pz[0] - z1), dist[0]);
      dist[0] = d2;
      bd = __float_as_uint(d2);
    }
#pragma unroll
    for (int p = 1; p < PPT; ++p) {
      const float d2 = fminf(sqdist(px[p] - x1, py[p] - y1, pz[p] - z1), dist[p]);
      dist[p] = d2;
      const unsigned db = __float_as_uint(d2);
      if (db > bd) { bd = db; bp = p; }
    }
    const unsigned bk = key0 + (unsigned)bp * ((CS * NT) >> 9);
    const unsigned wd = __reduce_max_sync(0xffffffffu, bd);
    const unsigned wk = __reduce_min_sync(0xffffffffu, (bd == wd && any) ? bk : 0xffffffffu);
    if (lane < CS) {
      const unsigned long long v = ((unsigned long long)wk << 32) | wd;   // uint2 {x = distance bits, y = key}
      asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];" ::"r"(
                       peer_cand + buf * ROUND_BYTES),
                   "l"(v), "r"(peer_bar + buf * 8u)
                   : "memory");
    }
    {  // wait for the 32*CS candidates of this round
      uint32_t done;
      do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar0 + buf * 8u), "r"(parity)
            : "memory");
      } while (!done);
    }
    unsigned ld = 0u, lk = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < (NW * CS + 31) / 32; ++i) {
      const int e = i * 32 + lane;
      if (e < NW * CS) {
        const uint2 c = s_cand[buf][e];
        if (c.y != 0xffffffffu && (lk == 0xffffffffu || c.x > ld || (c.x == ld && c.y < lk))) { ld = c.x; lk = c.y; }
      }
    }
    const unsigned gd = __reduce_max_sync(0xffffffffu, ld);
    const unsigned gk = __reduce_min_sync(0xffffffffu, ld == gd ? lk : 0xffffffffu);
    old = (gk == 0xffffffffu) ? 0 : fps_key_to_k(gk);
    if (rank == 0 && tid == 0) out[t + 1] = old;
    // Re-arm this buffer for round t + 2.  Thread 0 is past the wait, so the phase is complete; peers may already be
    // sending round t + 2 only after they received OUR round t + 1 candidates, which this thread has not sent yet.
    if (tid == 0 && t + 3 < m)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + buf * 8u), "r"(ROUND_BYTES)
                   : "memory");
  }
  cluster.sync();     // no CTA exits while a peer may still write into its shared memory
}

template <int PPT, int CS, int NT>
static int launch_fps_cluster(int b, int n, int m, const float *coords, int *indices, cudaStream_t s) {
  const size_t smem = sizeof(float) * 3 * (size_t)n;
  PVB_CUDA((cudaFuncSetAttribute(fps_cluster_kernel<PPT, CS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(b * CS));
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  PVB_CUDA((cudaLaunchKernelEx(&cfg, fps_cluster_kernel<PPT, CS, NT>, n, m, coords, indices)));
  ++pvb::g_launches;
  return 0;
}

template <int PPT, int NT>
static int launch_fps(int b, int n, int m, const float *coords, int *indices, cudaStream_t s) {
  size_t smem = sizeof(float) * 3 * (size_t)n;
  int in_smem = smem <= 200 * 1024;
  if (!in_smem) smem = 0;
  if (smem + 1024 > 48 * 1024)   // the static tables count against the 48 KB default as well (n = 4096 is exactly 48 KB)
    PVB_CUDA((cudaFuncSetAttribute(fps_kernel<PPT, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
  PVB_LAUNCH((fps_kernel<PPT, NT>), b, NT, smem, s, n, m, in_smem, coords, indices);
  return 0;
}

// =====================================================================================
// BallQuery grouping in one pass  (modules/ball_query.py:16-30: grouping(coords) - centre, grouping(features), cat)
// out [B, 3+C, M, U]: channels 0..2 = neighbour coordinate - centre coordinate, channels 3.. = neighbour features.
// The reference materialises [B,3,M,U] twice and [B,C,M,U] once before writing the concatenation.
// =====================================================================================
template <int CT>
__global__ void __launch_bounds__(256) group_concat_kernel(int c, int n, int m, int u, const float *__restrict__ coords,
                                                           const float *__restrict__ centers,
                                                           const float *__restrict__ feat, const int *__restrict__ idx,
                                                           float *__restrict__ out) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT, mu = m * u, ct = c + 3;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= mu) return;
  const int src = idx[(size_t)b * mu + e];
  const int ctr = e / u;
  float *o = out + ((size_t)b * ct + c0) * mu + e;
  const int cmax = min(CT, ct - c0);
#pragma unroll 4
  for (int j = 0; j < cmax; ++j) {
    const int ch = c0 + j;
    float v;
    if (ch < 3)
      v = __fsub_rn(__ldg(coords + ((size_t)b * 3 + ch) * n + src), __ldg(centers + ((size_t)b * 3 + ch) * m + ctr));
    else
      v = __ldg(feat + ((size_t)b * c + (ch - 3)) * n + src);
    o[(size_t)j * mu] = v;
  }
}

// grad_y [B,3+C,M,U] -> grad_features [B,C,N] (+)= channels 3.., grad_coords [B,3,N] (+)= channels 0..2 (optional),
// grad_centers [B,3,M] = - sum_u channels 0..2 (optional; one warp-segment reduction per centre via atomics)
template <int CT>
__global__ void __launch_bounds__(256) group_concat_grad_kernel(int c, int n, int m, int u,
                                                                const float *__restrict__ grad_y,
                                                                const int *__restrict__ idx,
                                                                float *__restrict__ grad_feat,
                                                                float *__restrict__ grad_coords,
                                                                float *__restrict__ grad_centers) {
  const int b = blockIdx.z, c0 = blockIdx.y * CT, lane = threadIdx.x & 31, mu = m * u, ct = c + 3;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const bool valid = e < mu;
  const int dst = valid ? idx[(size_t)b * mu + e] : -1 - lane;
  const int ctr = valid ? e / u : -1 - lane;
  const unsigned grp = __match_any_sync(0xffffffffu, dst);
  const bool leader = (__ffs(grp) - 1) == lane;
  const bool multi = __any_sync(0xffffffffu, grp != (1u << lane));
  const unsigned cgrp = __match_any_sync(0xffffffffu, ctr);
  const bool cleader = (__ffs(cgrp) - 1) == lane;
  const float *gy = grad_y + ((size_t)b * ct + c0) * mu + e;
  const int cmax = min(CT, ct - c0);
  for (int j = 0; j < cmax; ++j) {
    const int ch = c0 + j;
    const float g = valid ? gy[(size_t)j * mu] : 0.0f;
    float *base = ch < 3 ? grad_coords : grad_feat;   // warp-uniform
    if (base) {
      float v = g;
      if (multi) v = group_sum_ordered(grp, v);
      const size_t off = ch < 3 ? ((size_t)b * 3 + ch) * n : ((size_t)b * c + (ch - 3)) * n;
      if (valid && leader) atomicAdd(base + off + dst, v);
    }
    if (ch < 3 && grad_centers) {
      const float v = group_sum_ordered(cgrp, g);
      if (valid && cleader) atomicAdd(grad_centers + ((size_t)b * 3 + ch) * m + ctr, -v);
    }
  }
}

extern "C" {

int pvcnn_abi_version(void) { return PVCNN_B200_ABI_VERSION; }
const char *pvcnn_build_info(void) {
  return "pvcnn_b200 sm_100a (nvcc " __DATE__ " " __TIME__ ")";
}
unsigned long long pvcnn_launch_count(void) { return pvb::g_launches.load(); }

int pvcnn_voxelize_coords(int b, int n, int r, int normalize, float eps, const float *coords, float *norm_coords,
                          int *vox_coords, void *stream) {
  PVB_CHECK_ARG(b > 0 && n > 0 && r > 0 && coords && norm_coords && vox_coords);
  PVB_LAUNCH(voxelize_coords_kernel, b, 512, 0, stream, n, r, normalize, eps, coords, norm_coords, vox_coords);
  return 0;
}

int pvcnn_voxelize_denom(int b, int n, float eps, const float *coords, const float *mean, float *denom, void *stream) {
  PVB_CHECK_ARG(b > 0 && n > 0 && coords && mean && denom);
  PVB_LAUNCH(voxelize_denom_kernel, b, 512, 0, stream, n, eps, coords, mean, denom);
  return 0;
}

int pvcnn_voxelize_apply(int b, int n, int r, int normalize, const float *coords, const float *mean,
                         const float *denom, float *norm_coords, int *vox_coords, void *stream) {
  PVB_CHECK_ARG(b > 0 && n > 0 && r > 0 && coords && mean && (denom || !normalize) && norm_coords && vox_coords);
  const long long total = (long long)b * 3 * n;
  long long grid = (total + 255) / 256;
  if (grid > pvb::kNumSMs * 8) grid = pvb::kNumSMs * 8;
  PVB_LAUNCH(voxelize_apply_kernel, (int)grid, 256, 0, stream, n, r, normalize, total, coords, mean, denom, norm_coords,
             vox_coords);
  return 0;
}

int pvcnn_avg_voxelize(int b, int c, int n, int r, int r2, int r3, const int *coords, const float *feat, int *ind,
                       int *cnt, float *out, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && r > 0 && r2 == r * r && r3 == r2 * r);
  PVB_CHECK_ARG(coords && feat && ind && cnt && out);
  cudaStream_t s = (cudaStream_t)stream;
  PVB_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * (size_t)b * r3, s));
  PVB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)b * c * r3, s));
  const long long total = (long long)b * n;
  PVB_LAUNCH(vox_index_count_kernel, min(ceil_div(total, 256), kNumSMs * 8), 256, 0, s, n, r, r2, r3, total, coords,
             ind, cnt);
  constexpr int CT = 16;
  PVB_LAUNCH(vox_scatter_kernel<CT>, dim3(ceil_div(n, 128), ceil_div(c, CT), b), 128, 0, s, c, n, r3, ind, cnt, feat,
             out);
  return 0;
}

int pvcnn_avg_voxelize_grad(int b, int c, int n, int s_, const int *ind, const int *cnt, const float *grad_y,
                            float *grad_x, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && s_ > 0 && ind && cnt && grad_y && grad_x);
  constexpr int CT = 16;
  PVB_LAUNCH(vox_grad_kernel<CT>, dim3(ceil_div(n, 128), ceil_div(c, CT), b), 128, 0, stream, c, n, s_, ind, cnt,
             grad_y, grad_x);
  return 0;
}

int pvcnn_trilinear_devoxelize(int b, int c, int n, int r, int r2, int r3, int training, const float *coords,
                               const float *feat, int *inds, float *wgts, float *outs, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && r > 0 && r2 == r * r && r3 == r2 * r && coords && feat && outs);
  PVB_CHECK_ARG(!training || (inds && wgts));
  constexpr int CT = 16;
  PVB_LAUNCH(devox_kernel<CT>, dim3(ceil_div(n, 128), ceil_div(c, CT), b), 128, 0, stream, c, n, r, r2, r3, training,
             coords, feat, inds, wgts, outs);
  return 0;
}

int pvcnn_trilinear_devoxelize_grad(int b, int c, int n, int r3, const int *inds, const float *wgts,
                                    const float *grad_y, float *grad_x, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && r3 > 0 && inds && wgts && grad_y && grad_x);
  cudaStream_t s = (cudaStream_t)stream;
  PVB_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * r3, s));
  constexpr int CT = 16;
  PVB_LAUNCH(devox_grad_kernel<CT>, dim3(ceil_div(n, 128), ceil_div(c, CT), b), 128, 0, s, c, n, r3, inds, wgts,
             grad_y, grad_x);
  return 0;
}

int pvcnn_ball_query(int b, int n, int m, float r2, int u, const float *centers_coords, const float *points_coords,
                     int *neighbors_indices, void *stream) {
  PVB_CHECK_ARG(b > 0 && n > 0 && m > 0 && u > 0 && centers_coords && points_coords && neighbors_indices);
  PVB_LAUNCH(ball_query_kernel, dim3(ceil_div(m, BQ_WARPS * BQ_QPW), b), BQ_WARPS * 32, 0, stream, n, m, r2, u,
             centers_coords, points_coords, neighbors_indices);
  return 0;
}

int pvcnn_grouping(int b, int c, int n, int m, int u, const float *features, const int *indices, float *out,
                   void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && m > 0 && u > 0 && features && indices && out);
  constexpr int CT = 8;
  PVB_LAUNCH(grouping_kernel<CT>, dim3(ceil_div((long long)m * u, 256), ceil_div(c, CT), b), 256, 0, stream, c, n,
             m * u, features, indices, out);
  return 0;
}

int pvcnn_grouping_grad(int b, int c, int n, int m, int u, const float *grad_y, const int *indices, float *grad_x,
                        void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && n > 0 && m > 0 && u > 0 && grad_y && indices && grad_x);
  cudaStream_t s = (cudaStream_t)stream;
  PVB_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * n, s));
  constexpr int CT = 8;
  PVB_LAUNCH(grouping_grad_kernel<CT>, dim3(ceil_div((long long)m * u, 256), ceil_div(c, CT), b), 256, 0, s, c, n,
             m * u, grad_y, indices, grad_x);
  return 0;
}

int pvcnn_group_concat(int b, int c, int n, int m, int u, const float *points_coords, const float *centers_coords,
                       const float *features, const int *indices, float *out, void *stream) {
  PVB_CHECK_ARG(b > 0 && c >= 0 && n > 0 && m > 0 && u > 0 && points_coords && centers_coords && indices && out);
  PVB_CHECK_ARG(c == 0 || features != nullptr);
  constexpr int CT = 8;
  PVB_LAUNCH(group_concat_kernel<CT>, dim3(ceil_div((long long)m * u, 256), ceil_div(c + 3, CT), b), 256, 0, stream, c,
             n, m, u, points_coords, centers_coords, features, indices, out);
  return 0;
}

int pvcnn_group_concat_grad(int b, int c, int n, int m, int u, const float *grad_y, const int *indices,
                            float *grad_features, float *grad_points_coords, float *grad_centers_coords,
                            void *stream) {
  PVB_CHECK_ARG(b > 0 && c >= 0 && n > 0 && m > 0 && u > 0 && grad_y && indices);
  PVB_CHECK_ARG(c == 0 || grad_features != nullptr);
  cudaStream_t s = (cudaStream_t)stream;
  if (c > 0) PVB_CUDA(cudaMemsetAsync(grad_features, 0, sizeof(float) * (size_t)b * c * n, s));
  if (grad_points_coords) PVB_CUDA(cudaMemsetAsync(grad_points_coords, 0, sizeof(float) * (size_t)b * 3 * n, s));
  if (grad_centers_coords) PVB_CUDA(cudaMemsetAsync(grad_centers_coords, 0, sizeof(float) * (size_t)b * 3 * m, s));
  constexpr int CT = 8;
  // channel tiles: tile 0 holds the 3 coordinate channels (+ the first 5 feature channels)
  PVB_LAUNCH(group_concat_grad_kernel<CT>, dim3(ceil_div((long long)m * u, 256), ceil_div(c + 3, CT), b), 256, 0, s, c,
             n, m, u, grad_y, indices, c > 0 ? grad_features : nullptr, grad_points_coords, grad_centers_coords);
  return 0;
}

int pvcnn_gather_features(int b, int c, int n, int m, const float *features, const int *indices, float *out,
                          void *stream) {
  return pvcnn_grouping(b, c, n, m, 1, features, indices, out, stream);
}

int pvcnn_gather_features_grad(int b, int c, int n, int m, const float *grad_y, const int *indices, float *grad_x,
                               void *stream) {
  return pvcnn_grouping_grad(b, c, n, m, 1, grad_y, indices, grad_x, stream);
}

int pvcnn_furthest_point_sampling(int b, int n, int m, const float *coords, float *distances, int *indices,
                                  void *stream) {
  PVB_CHECK_ARG(b > 0 && n > 0 && coords && indices);
  if (m <= 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  // Default: a cluster of 4 CTAs per cloud (128 threads each up to 8192 points, 256 above) when the cloud fits in shared
  // memory, else one CTA of 512 threads.  PVCNN_B200_FPS: "cta" = single-CTA kernels only, "t1024" = 1024-thread single-CTA
  // kernels (the round-1 shape), "cNNN" = cluster of 4 CTAs x NNN threads (128, 256, 1024).
  const char *e = getenv("PVCNN_B200_FPS");
  const char *emin = getenv("PVCNN_B200_FPS_NMIN");
  const int nmin = emin ? atoi(emin) : 2048;
  const bool fits = sizeof(float) * 3 * (size_t)n <= 200 * 1024;
  const bool single = e && (e[0] == 't' || (e[0] == 'c' && e[1] == 't'));
  if (!single && fits && n >= nmin) {
    const int nt = (e && e[0] == 'c') ? atoi(e + 1) : (n <= 8192 ? 128 : 256);
    const int pp = ceil_div(n, 4 * nt);
#define PVB_FPS_C(PPT, NT) return launch_fps_cluster<PPT, 4, NT>(b, n, m, coords, indices, s)
    if (nt == 1024) { if (pp <= 1) PVB_FPS_C(1, 1024); if (pp <= 2) PVB_FPS_C(2, 1024); if (pp <= 4) PVB_FPS_C(4, 1024); }
    if (nt == 256) { if (pp <= 1) PVB_FPS_C(1, 256); if (pp <= 2) PVB_FPS_C(2, 256); if (pp <= 4) PVB_FPS_C(4, 256); if (pp <= 8) PVB_FPS_C(8, 256); if (pp <= 16) PVB_FPS_C(16, 256); }
    if (nt == 128) { if (pp <= 1) PVB_FPS_C(1, 128); if (pp <= 2) PVB_FPS_C(2, 128); if (pp <= 4) PVB_FPS_C(4, 128); if (pp <= 8) PVB_FPS_C(8, 128); if (pp <= 16) PVB_FPS_C(16, 128); if (pp <= 32) PVB_FPS_C(32, 128); }
#undef PVB_FPS_C
  }
  if (!(e && e[0] == 't' && e[1] == '1')) {   // 512 threads, up to 32 points each
    const int pp = ceil_div(n, 512);
    if (pp <= 1) return launch_fps<1, 512>(b, n, m, coords, indices, s);
    if (pp <= 2) return launch_fps<2, 512>(b, n, m, coords, indices, s);
    if (pp <= 4) return launch_fps<4, 512>(b, n, m, coords, indices, s);
    if (pp <= 8) return launch_fps<8, 512>(b, n, m, coords, indices, s);
    if (pp <= 16) return launch_fps<16, 512>(b, n, m, coords, indices, s);
    if (pp <= 32) return launch_fps<32, 512>(b, n, m, coords, indices, s);
  }
  const int ppt = ceil_div(n, FPS_THREADS);
  if (ppt <= 1) return launch_fps<1, 1024>(b, n, m, coords, indices, s);
  if (ppt <= 2) return launch_fps<2, 1024>(b, n, m, coords, indices, s);
  if (ppt <= 4) return launch_fps<4, 1024>(b, n, m, coords, indices, s);
  if (ppt <= 8) return launch_fps<8, 1024>(b, n, m, coords, indices, s);
  if (n <= 32 * 512) return launch_fps<32, 512>(b, n, m, coords, indices, s);
  float *scratch = distances;
  if (!scratch) PVB_CUDA(cudaMallocAsync((void **)&scratch, sizeof(float) * (size_t)b * n, s));
  PVB_LAUNCH(fps_kernel_big, b, FPS_THREADS, 0, s, n, m, coords, scratch, indices);
  if (!distances) PVB_CUDA(cudaFreeAsync(scratch, s));
  return 0;
}

int pvcnn_three_nearest_neighbors_interpolate(int b, int c, int m, int n, const float *points_coords,
                                              const float *centers_coords, const float *centers_features,
                                              int *indices, float *weights, float *out, void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && m > 0 && n > 0 && points_coords && centers_coords && centers_features);
  PVB_CHECK_ARG(indices && weights && out);
  PVB_LAUNCH(three_nn_kernel, dim3(ceil_div(n, 128), b), 128, 0, stream, n, m, points_coords, centers_coords, weights,
             indices);
  constexpr int CT = 16;
  PVB_LAUNCH(three_nn_interp_kernel<CT>, dim3(ceil_div(n, 128), ceil_div(c, CT), b), 128, 0, stream, c, m, n,
             centers_features, indices, weights, out);
  return 0;
}

int pvcnn_three_nearest_neighbors_interpolate_grad(int b, int c, int n, int m, const float *grad_y,
                                                   const int *indices, const float *weights, float *grad_x,
                                                   void *stream) {
  PVB_CHECK_ARG(b > 0 && c > 0 && m > 0 && n > 0 && grad_y && indices && weights && grad_x);
  cudaStream_t s = (cudaStream_t)stream;
  PVB_CUDA(cudaMemsetAsync(grad_x, 0, sizeof(float) * (size_t)b * c * m, s));
  constexpr int CT = 16;
  PVB_LAUNCH(three_nn_interp_grad_kernel<CT>, dim3(ceil_div(n, 128), ceil_div(c, CT), b), 128, 0, s, c, n, m, grad_y,
             indices, weights, grad_x);
  return 0;
}

int pvcnn_logits_mask_sample(int b, int n, int k, unsigned long long seed, const unsigned char *mask, int *picks,
                             void *stream) {
  PVB_CHECK_ARG(b > 0 && n > 0 && k > 0 && mask && picks);
  int p2 = 1;
  while (p2 < (n > k ? n : k)) p2 <<= 1;
  const size_t smem = (size_t)p2 * 8 + (size_t)n * 8;
  if (smem > 200 * 1024) return PVCNN_E_UNSUPPORTED;  // n, k <= 8192 (the reference's largest use: N = 1024, k = 512)
  if (smem > 40 * 1024)  // static shared memory counts against the 48 KB default too
    PVB_CUDA(cudaFuncSetAttribute(pvb::logits_mask_sample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PVB_LAUNCH(pvb::logits_mask_sample_kernel, b, 1024, smem, stream, n, k, seed, mask, picks);
  return 0;
}

}  // extern "C"
