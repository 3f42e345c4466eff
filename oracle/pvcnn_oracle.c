/*
 * pvcnn_oracle.c -- CPU restatement of the reference's custom-op arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (pvcnn_b200/, modules/) may
 * import, link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, as the checker.
 *
 * Every function restates one kernel of mit-han-lab/pvcnn (paths relative to
 * /root/reference/modules/functional/src/).  The reference has no CPU
 * implementation (utils.hpp:7 rejects CPU tensors), so this file follows the
 * CUDA kernels line by line, including their quirks:
 *   - products are contracted into FMAs exactly where nvcc contracts them
 *     (verified on the sm_100a SASS of the reference build);
 *     this file must therefore be compiled with -ffp-contract=off so that the
 *     explicit fmaf() calls below are the only fused operations;
 *   - atomics make the reference's summation order nondeterministic; the oracle
 *     sums in ascending point order (one of the orders the reference can take).
 *
 * Parity pinning: the reference has no tests / golden vectors.  The
 * oracle is pinned against the reference's own kernels, compiled unmodified into
 * oracle/_ref/ by oracle/build_ref.py and run on a B200: their outputs are committed as
 * tests/golden/ref_ops_golden.npz (generator: tests/golden/make_golden.py) and checked on CPU by
 * tests/test_golden_cpu.py; tests/test_ops_gpu.py additionally compares oracle, our kernels and the live
 * reference .so on the GPU box.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORACLE_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------
 * Coordinate normalisation: modules/voxelization.py:16-25 (torch tensor ops).
 *   c0 = c - mean_N(c); normalize: nc = c0 / (max_N ||c0||_2 * 2 + eps) + 0.5
 *   else nc = (c0 + 1) / 2;  nc = clamp(nc * r, 0, r-1);  vc = round_half_even(nc)
 * torch does not specify its reduction order; the oracle (and our kernel) define the
 * mean as the double-precision sum rounded once to float.
 * ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_voxelize_coords(int b, int n, int r, int normalize, float eps,
                                       const float *coords, float *norm_coords, int *vox_coords) {
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const float *c = coords + (size_t)bi * 3 * n;
    float *nc = norm_coords + (size_t)bi * 3 * n;
    int *vc = vox_coords + (size_t)bi * 3 * n;
    float mean[3];
    for (int a = 0; a < 3; ++a) {
      double s = 0.0;
      for (int i = 0; i < n; ++i) s += (double)c[a * n + i];
      mean[a] = (float)(s / (double)n);
    }
    float maxnorm = 0.0f;
    if (normalize) {
      for (int i = 0; i < n; ++i) {
        float x = c[i] - mean[0], y = c[n + i] - mean[1], z = c[2 * n + i] - mean[2];
        float s = x * x;
        s = s + y * y;
        s = s + z * z;
        float nr = sqrtf(s);
        if (nr > maxnorm || nr != nr) maxnorm = nr;
      }
    }
    float denom = maxnorm * 2.0f + eps;
    for (int a = 0; a < 3; ++a) {
      for (int i = 0; i < n; ++i) {
        float v = c[a * n + i] - mean[a];
        if (normalize) v = v / denom + 0.5f;
        else v = (v + 1.0f) / 2.0f;
        v = v * (float)r;
        /* torch.clamp propagates NaN */
        if (v < 0.0f) v = 0.0f;
        if (v > (float)(r - 1)) v = (float)(r - 1);
        nc[a * n + i] = v;
        vc[a * n + i] = (int)rintf(v); /* round-half-even under the default rounding mode */
      }
    }
  }
}

/* modules/voxelization.py:17-24 with the per-cloud mean GIVEN (the result of the reference's own
 * `coords.mean(2, keepdim=True)`, :18, whose summation order is an implementation detail of torch).  Everything
 * after the mean is element-wise IEEE arithmetic or an order-independent max:
 *   norm = sqrt((x*x + y*y) + z*z)            (ATen's 3-element reduce: one accumulator per element, combined in order)
 *   denom = max_N(norm) * 2 + eps             (:20)
 *   v = (c - mean) / denom + 0.5  |  (c - mean + 1) / 2  ;  clamp(v * r, 0, r-1) ;  round-half-even -> int32   (:20-24) */
ORACLE_API void oracle_voxelize_coords_given(int b, int n, int r, int normalize, float eps, const float *coords,
                                             const float *mean_b3, float *norm_coords, int *vox_coords) {
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const float *c = coords + (size_t)bi * 3 * n;
    const float *mean = mean_b3 + (size_t)bi * 3;
    float *nc = norm_coords + (size_t)bi * 3 * n;
    int *vc = vox_coords + (size_t)bi * 3 * n;
    float maxnorm = -INFINITY;
    if (normalize) {
      for (int i = 0; i < n; ++i) {
        float x = c[i] - mean[0], y = c[n + i] - mean[1], z = c[2 * n + i] - mean[2];
        float s = x * x + y * y;
        s = s + z * z;
        float nr = sqrtf(s);
        if (maxnorm != maxnorm) continue; /* NaN sticks, like torch.max */
        if (nr > maxnorm || nr != nr) maxnorm = nr;
      }
    }
    float denom = maxnorm * 2.0f + eps;
    for (int a = 0; a < 3; ++a) {
      for (int i = 0; i < n; ++i) {
        float v = c[a * n + i] - mean[a];
        if (normalize) v = v / denom + 0.5f;
        else v = (v + 1.0f) / 2.0f;
        v = v * (float)r;
        if (v < 0.0f) v = 0.0f;
        if (v > (float)(r - 1)) v = (float)(r - 1);
        nc[a * n + i] = v;
        vc[a * n + i] = (int)rintf(v);
      }
    }
  }
}

/* ------------------------------------------------------------------------------------
 * avg_voxelize forward: voxelization/vox.cu:18-34 (grid_stats_kernel) + :48-72
 * (avg_voxelize_kernel); allocation / zero-init rules from vox.cpp:17-43.
 * ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_avg_voxelize(int b, int c, int n, int r, const int *coords, const float *feat,
                                    int *ind, int *cnt, float *out) {
  const int r2 = r * r, r3 = r2 * r;
  memset(cnt, 0, sizeof(int) * (size_t)b * r3);
  memset(out, 0, sizeof(float) * (size_t)b * c * r3);
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const int *co = coords + (size_t)bi * 3 * n;
    int *in_ = ind + (size_t)bi * n;
    int *cn = cnt + (size_t)bi * r3;
    for (int i = 0; i < n; ++i) { /* vox.cu:31-32 */
      in_[i] = co[i] * r2 + co[i + n] * r + co[i + 2 * n];
      cn[in_[i]] += 1;
    }
    const float *f = feat + (size_t)bi * c * n;
    float *o = out + (size_t)bi * c * r3;
    for (int i = 0; i < n; ++i) { /* vox.cu:60-71 */
      int pos = in_[i];
      int cur = cn[pos];
      if (cur > 0) {
        float inv = (float)(1.0 / (double)(float)cur); /* vox.cu:66: double divide, rounded to float */
        for (int j = 0; j < c; ++j) o[(size_t)j * r3 + pos] += f[(size_t)j * n + i] * inv;
      }
    }
  }
}

/* avg_voxelize backward: vox.cu:86-110, vox.cpp:54-76 */
ORACLE_API void oracle_avg_voxelize_grad(int b, int c, int n, int r3, const int *ind, const int *cnt,
                                         const float *grad_y, float *grad_x) {
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const int *in_ = ind + (size_t)bi * n;
    const int *cn = cnt + (size_t)bi * r3;
    const float *gy = grad_y + (size_t)bi * c * r3;
    float *gx = grad_x + (size_t)bi * c * n;
    for (int j = 0; j < c; ++j)
      for (int i = 0; i < n; ++i) {
        int pos = in_[i];
        int cur = cn[pos];
        float v = 0.0f;
        if (cur > 0) {
          float inv = (float)(1.0 / (double)(float)cur);
          v = gy[(size_t)j * r3 + pos] * inv;
        }
        gx[(size_t)j * n + i] = v;
      }
  }
}

/* ------------------------------------------------------------------------------------
 * trilinear_devoxelize forward: interpolate/trilinear_devox.cu:21-105; inds/wgts are
 * only written when is_training (trilinear_devox.cpp:45-53).
 * The 8-term sum is compiled by nvcc to: acc = w000*f000; acc = fma(w001,f001,acc); ...
 * (SASS: one FMUL followed by seven FFMA in corner order 000..111).
 * ------------------------------------------------------------------------------------ */
static inline void devox_setup(float x, float y, float z, int r, int r2, float w[8], int idx[8]) {
  float xl = floorf(x), yl = floorf(y), zl = floorf(z);
  float xd1 = x - xl, yd1 = y - yl, zd1 = z - zl;
  float xd0 = 1.0f - xd1, yd0 = 1.0f - yd1, zd0 = 1.0f - zd1;
  w[0] = xd0 * yd0 * zd0; w[1] = xd0 * yd0 * zd1; w[2] = xd0 * yd1 * zd0; w[3] = xd0 * yd1 * zd1;
  w[4] = xd1 * yd0 * zd0; w[5] = xd1 * yd0 * zd1; w[6] = xd1 * yd1 * zd0; w[7] = xd1 * yd1 * zd1;
  int xlo = (int)xl, ylo = (int)yl, zlo = (int)zl;
  int xhi = (xd1 > 0) ? -1 : 0, yhi = (yd1 > 0) ? -1 : 0, zhi = (zd1 > 0) ? 1 : 0;
  idx[0] = xlo * r2 + ylo * r + zlo;  /* trilinear_devox.cu:68-75 */
  idx[1] = idx[0] + zhi;
  idx[2] = idx[0] + (yhi & r);
  idx[3] = idx[2] + zhi;
  idx[4] = idx[0] + (xhi & r2);
  idx[5] = idx[4] + zhi;
  idx[6] = idx[4] + (yhi & r);
  idx[7] = idx[6] + zhi;
}

ORACLE_API void oracle_trilinear_devoxelize(int b, int c, int n, int r, int is_training, const float *coords,
                                            const float *feat, int *inds, float *wgts, float *outs) {
  const int r2 = r * r, r3 = r2 * r;
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    const float *co = coords + (size_t)bi * 3 * n;
    const float *f = feat + (size_t)bi * c * r3;
    float *o = outs + (size_t)bi * c * n;
    for (int i = 0; i < n; ++i) {
      float w[8];
      int idx[8];
      devox_setup(co[i], co[i + n], co[i + 2 * n], r, r2, w, idx);
      if (is_training) {
        for (int k = 0; k < 8; ++k) {
          wgts[(size_t)bi * 8 * n + (size_t)k * n + i] = w[k];
          inds[(size_t)bi * 8 * n + (size_t)k * n + i] = idx[k];
        }
      }
      for (int j = 0; j < c; ++j) {
        const float *fj = f + (size_t)j * r3;
        float acc = w[0] * fj[idx[0]];
        for (int k = 1; k < 8; ++k) acc = fmaf(w[k], fj[idx[k]], acc);
        o[(size_t)j * n + i] = acc;
      }
    }
  }
}

/* trilinear_devoxelize backward: trilinear_devox.cu:119-162, trilinear_devox.cpp:67-91 */
ORACLE_API void oracle_trilinear_devoxelize_grad(int b, int c, int n, int r3, const int *inds, const float *wgts,
                                                 const float *grad_y, float *grad_x) {
  memset(grad_x, 0, sizeof(float) * (size_t)b * c * r3);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < c; ++j) {
      const int *in_ = inds + (size_t)bi * 8 * n;
      const float *wg = wgts + (size_t)bi * 8 * n;
      const float *gy = grad_y + ((size_t)bi * c + j) * n;
      float *gx = grad_x + ((size_t)bi * c + j) * r3;
      for (int i = 0; i < n; ++i) {
        float g = gy[i];
        for (int k = 0; k < 8; ++k) gx[in_[(size_t)k * n + i]] += wg[(size_t)k * n + i] * g;
      }
    }
}

/* ------------------------------------------------------------------------------------
 * ball_query: ball_query/ball_query.cu:19-50; radius squared in float by the caller
 * (ball_query.cpp:24).  d2 = fmaf(dz,dz, fmaf(dx,dx, dy*dy))  (SASS: FMUL dy, FFMA dx, FFMA dz)
 * ------------------------------------------------------------------------------------ */
static inline float sqdist(float dx, float dy, float dz) { return fmaf(dz, dz, fmaf(dx, dx, dy * dy)); }

ORACLE_API void oracle_ball_query(int b, int n, int m, float r2, int u, const float *centers, const float *points,
                                  int *out) {
  memset(out, 0, sizeof(int) * (size_t)b * m * u); /* ball_query.cpp:20-22 torch::zeros */
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < m; ++j) {
      const float *p = points + (size_t)bi * 3 * n;
      const float *ce = centers + (size_t)bi * 3 * m;
      int *o = out + ((size_t)bi * m + j) * u;
      float cx = ce[j], cy = ce[j + m], cz = ce[j + 2 * m];
      int cnt = 0;
      for (int k = 0; k < n && cnt < u; ++k) {
        float d2 = sqdist(cx - p[k], cy - p[k + n], cz - p[k + 2 * n]);
        if (d2 < r2) {
          if (cnt == 0)
            for (int v = 0; v < u; ++v) o[v] = k;
          o[cnt] = k;
          ++cnt;
        }
      }
    }
}

/* grouping forward / backward: grouping/grouping.cu:18-36, :58-77 */
ORACLE_API void oracle_grouping(int b, int c, int n, int m, int u, const float *feat, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *f = feat + ((size_t)bi * c + l) * n;
      const int *id = idx + (size_t)bi * m * u;
      float *o = out + ((size_t)bi * c + l) * m * u;
      for (int j = 0; j < m * u; ++j) o[j] = f[id[j]];
    }
}

ORACLE_API void oracle_grouping_grad(int b, int c, int n, int m, int u, const float *grad_y, const int *idx,
                                     float *grad_x) {
  memset(grad_x, 0, sizeof(float) * (size_t)b * c * n);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      float *gx = grad_x + ((size_t)bi * c + l) * n;
      const int *id = idx + (size_t)bi * m * u;
      const float *gy = grad_y + ((size_t)bi * c + l) * m * u;
      for (int j = 0; j < m * u; ++j) gx[id[j]] += gy[j];
    }
}

/* gather forward / backward: sampling/sampling.cu:17-31, :52-66 */
ORACLE_API void oracle_gather(int b, int c, int n, int m, const float *feat, const int *idx, float *out) {
  oracle_grouping(b, c, n, m, 1, feat, idx, out);
}
ORACLE_API void oracle_gather_grad(int b, int c, int n, int m, const float *grad_y, const int *idx, float *grad_x) {
  oracle_grouping_grad(b, c, n, m, 1, grad_y, idx, grad_x);
}

/* ------------------------------------------------------------------------------------
 * furthest_point_sampling: sampling/sampling.cu:86-167 with its fixed 512-thread block.
 * Thread t owns k = t, t+512, ...; per-thread strict '>' argmax (best=-1, besti=0), then
 * the strided 512-slot tree (replace iff dists[i1] < dists[i2]).  Reproduced literally so
 * that ties break exactly as on the GPU.  Scratch distances start at 1e38 (sampling.cpp:53).
 * ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_furthest_point_sampling(int b, int n, int m, const float *coords, int *indices) {
  if (m <= 0) return;
  memset(indices, 0, sizeof(int) * (size_t)b * m);
#pragma omp parallel for
  for (int bi = 0; bi < b; ++bi) {
    enum { BS = 512 };
    const float *co = coords + (size_t)bi * 3 * n;
    int *out = indices + (size_t)bi * m;
    float *dist = (float *)malloc(sizeof(float) * (size_t)n);
    float dists[BS];
    int dists_i[BS];
    for (int k = 0; k < n; ++k) dist[k] = 1e38f;
    int old = 0;
    out[0] = 0;
    for (int j = 1; j < m; ++j) {
      float x1 = co[old], y1 = co[old + n], z1 = co[old + 2 * n];
      for (int t = 0; t < BS; ++t) {
        int besti = 0;
        float best = -1.0f;
        for (int k = t; k < n; k += BS) {
          float d = sqdist(co[k] - x1, co[k + n] - y1, co[k + 2 * n] - z1);
          float d2 = fminf(d, dist[k]);
          if (d2 != dist[k]) dist[k] = d2;
          if (d2 > best) { best = d2; besti = k; }
        }
        dists[t] = best;
        dists_i[t] = besti;
      }
      for (int u = 0; (1 << u) < BS; ++u)
        for (int t = 0; t < (BS >> (u + 1)); ++t) {
          int i1 = (t * 2) << u, i2 = (t * 2 + 1) << u;
          if (dists[i1] < dists[i2]) { dists[i1] = dists[i2]; dists_i[i1] = dists_i[i2]; }
        }
      old = dists_i[0];
      out[j] = old;
    }
    free(dist);
  }
}

/* ------------------------------------------------------------------------------------
 * three-NN search + interpolation: interpolate/neighbor_interpolate.cu:20-75 (search),
 * :90-116 (interpolate), :145-170 (grad).  float d compared against double bests.
 * ------------------------------------------------------------------------------------ */
ORACLE_API void oracle_three_nn(int b, int n, int m, const float *points, const float *centers, float *weights,
                                int *indices) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int j = 0; j < n; ++j) {
      const float *p = points + (size_t)bi * 3 * n;
      const float *ce = centers + (size_t)bi * 3 * m;
      float *w = weights + (size_t)bi * 3 * n;
      int *id = indices + (size_t)bi * 3 * n;
      float ux = p[j], uy = p[j + n], uz = p[j + 2 * n];
      double best0 = 1e40, best1 = 1e40, best2 = 1e40;
      int i0 = 0, i1 = 0, i2 = 0;
      for (int k = 0; k < m; ++k) {
        float d = sqdist(ux - ce[k], uy - ce[k + m], uz - ce[k + 2 * m]);
        if (d < best2) {
          best2 = d; i2 = k;
          if (d < best1) {
            best2 = best1; i2 = i1; best1 = d; i1 = k;
            if (d < best0) { best1 = best0; i1 = i0; best0 = d; i0 = k; }
          }
        }
      }
      /* neighbor_interpolate.cu:61-63: max(min(1e10f, best), 1e-10f) evaluated in double */
      best0 = fmax(fmin((double)1e10f, best0), (double)1e-10f);
      best1 = fmax(fmin((double)1e10f, best1), (double)1e-10f);
      best2 = fmax(fmin((double)1e10f, best2), (double)1e-10f);
      float d0d1 = (float)(best0 * best1), d0d2 = (float)(best0 * best2), d1d2 = (float)(best1 * best2);
      float inv = 1.0f / (d0d1 + d0d2 + d1d2);
      w[j] = d1d2 * inv;          id[j] = i0;
      w[j + n] = d0d2 * inv;      id[j + n] = i1;
      w[j + 2 * n] = d0d1 * inv;  id[j + 2 * n] = i2;
    }
}

ORACLE_API void oracle_three_nn_interpolate(int b, int c, int m, int n, const float *cfeat, const int *indices,
                                            const float *weights, float *out) {
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *f = cfeat + ((size_t)bi * c + l) * m;
      const int *id = indices + (size_t)bi * 3 * n;
      const float *w = weights + (size_t)bi * 3 * n;
      float *o = out + ((size_t)bi * c + l) * n;
      for (int j = 0; j < n; ++j) {
        /* nvcc: a*w1 + b*w2 + c*w3 -> fma(c,w3, fma(b,w2, a*w1)) */
        o[j] = fmaf(f[id[j + 2 * n]], w[j + 2 * n], fmaf(f[id[j + n]], w[j + n], f[id[j]] * w[j]));
      }
    }
}

ORACLE_API void oracle_three_nn_interpolate_grad(int b, int c, int n, int m, const float *grad_y, const int *indices,
                                                 const float *weights, float *grad_x) {
  memset(grad_x, 0, sizeof(float) * (size_t)b * c * m);
#pragma omp parallel for collapse(2)
  for (int bi = 0; bi < b; ++bi)
    for (int l = 0; l < c; ++l) {
      const float *gy = grad_y + ((size_t)bi * c + l) * n;
      const int *id = indices + (size_t)bi * 3 * n;
      const float *w = weights + (size_t)bi * 3 * n;
      float *gx = grad_x + ((size_t)bi * c + l) * m;
      for (int j = 0; j < n; ++j) {
        gx[id[j]] += gy[j] * w[j];
        gx[id[j + n]] += gy[j] * w[j + n];
        gx[id[j + 2 * n]] += gy[j] * w[j + 2 * n];
      }
    }
}

ORACLE_API int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
