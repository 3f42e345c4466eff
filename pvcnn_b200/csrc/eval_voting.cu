// eval_voting.cu -- the reference's test-time voting loops on the device.
//
// evaluate/s3dis/eval.py:149-179 and evaluate/shapenet/eval.py:149-168 do, per batch of windows / per shape, on the
// host: tile arange(num_points_in_window) up to `total_num_voted_points`, np.random.shuffle it, gather the points,
// run the network, softmax -> max over the classes, copy confidences and predictions back, and merge them into the
// scene with a numba loop (`update_scene_predictions` :189-204 / `update_shape_predictions` :177-185: a vote replaces
// the scene's entry iff its confidence is strictly larger, so among equal confidences the EARLIEST vote in (window,
// position) order wins); `update_stats` (:207-215) is a confusion histogram.  datasets/s3dis.py:86-89 draws the
// training sample of a window with np.random.choice.  Here every one of those steps is a kernel and nothing but the
// final [3, classes] counters goes back to the host:
//   vote_indices / window_indices   counter-based pseudo-random permutation (balanced Feistel network + cycle walking)
//   vote_gather                     rows -> [batch*extra, C, num_points] network input
//   softmax_max                     F.softmax(logits, 1)[:, c0:c1].max(1)
//   vote_merge + vote_assign        64-bit atomicMax on (confidence bits | ~order), then the winner writes its class
//   vote_stats                      shared-memory histogram of (ground truth, prediction)
// All of it is HBM / atomic bound integer and byte work; no tensor cores.
#include "common.cuh"

namespace pvb {

// ---------------------------------------------------------------------------------------------------------------
// Pseudo-random permutation of [0, n), n <= 2^31: six rounds of a balanced Feistel network over 2h bits (the
// smallest even width that covers n), walked along its cycle until the value falls back into [0, n).  A Feistel
// network is a bijection of [0, 2^(2h)) whatever the round function, and cycle walking restricts a bijection of a
// superset to a bijection of the subset; 2^(2h) < 4n, so the walk takes < 4 steps on average.  The round keys come
// from splitmix64(seed, stream, round).  oracle/eval_voting.py::feistel_perm restates it bit for bit.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kVtRounds = 6;

struct VtPerm {
  uint32_t key[kVtRounds];
  uint32_t h, mask, n;
};

__device__ __forceinline__ unsigned long long vt_mix(unsigned long long seed, unsigned long long stream,
                                                     unsigned long long j) {
  unsigned long long x = seed ^ (stream * 0x9E3779B97F4A7C15ull) ^ (j * 0x94D049BB133111EBull);
  x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
  x ^= x >> 27; x *= 0x94D049BB133111EBull;
  x ^= x >> 31;
  return x;
}

__device__ __forceinline__ VtPerm vt_perm_init(unsigned long long seed, unsigned long long stream, uint32_t n) {
  VtPerm p;
  p.n = n;
  uint32_t h = 1;
  while ((1ull << (2 * h)) < (unsigned long long)n) ++h;
  p.h = h;
  p.mask = (1u << h) - 1u;   // h <= 16
#pragma unroll
  for (int r = 0; r < kVtRounds; ++r) p.key[r] = (uint32_t)(vt_mix(seed, stream, (unsigned long long)r) >> 32);
  return p;
}

__device__ __forceinline__ uint32_t vt_round(uint32_t r, uint32_t key) {
  uint32_t x = r * 0x9E3779B1u + key;
  x ^= x >> 16; x *= 0x85EBCA6Bu;
  x ^= x >> 13; x *= 0xC2B2AE35u;
  x ^= x >> 16;
  return x;
}

__device__ __forceinline__ uint32_t vt_perm(const VtPerm &p, uint32_t x) {
  do {
    uint32_t l = x >> p.h, r = x & p.mask;
#pragma unroll
    for (int k = 0; k < kVtRounds; ++k) {
      const uint32_t t = l ^ (vt_round(r, p.key[k]) & p.mask);
      l = r;
      r = t;
    }
    x = (l << p.h) | r;
  } while (x >= p.n);
  return x;
}

// eval.py:161-164: indices[w, p] = shuffle(tile(arange(n_w)))[:nv][p] = perm_nv(p) mod n_w  (tile(...)[q] = q mod n_w)
__global__ void __launch_bounds__(256) vote_indices_kernel(int nv, unsigned long long seed, int first_window,
                                                           const int *__restrict__ num_points,
                                                           int *__restrict__ indices) {
  const int w = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nv) return;
  const int n = num_points[w];
  int v = 0;
  if (n > 0) {
    const VtPerm pm = vt_perm_init(seed, (unsigned long long)(unsigned int)(first_window + w), (uint32_t)nv);
    v = (int)(vt_perm(pm, (uint32_t)p) % (uint32_t)n);
  }
  indices[(size_t)w * nv + p] = v;
}

// datasets/s3dis.py:86-87: np.random.choice(n_w, k, replace=(n_w < k)).  Without replacement: the first k entries of a
// random permutation of [0, n_w); with replacement: k independent uniform draws (multiply-high of a 32-bit variate).
__global__ void __launch_bounds__(256) window_indices_kernel(int k, unsigned long long seed, int first_window,
                                                             const int *__restrict__ num_points,
                                                             int *__restrict__ indices) {
  const int w = blockIdx.y;
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= k) return;
  const int n = num_points[w];
  const unsigned long long stream = (unsigned long long)(unsigned int)(first_window + w);
  int v = 0;
  if (n >= k) {
    const VtPerm pm = vt_perm_init(seed, stream, (uint32_t)n);
    v = (int)vt_perm(pm, (uint32_t)j);
  } else if (n > 0) {
    const uint32_t u = (uint32_t)(vt_mix(seed ^ 0xD1B54A32D192ED03ull, stream, (unsigned long long)j) >> 32);
    v = (int)(((unsigned long long)u * (unsigned long long)(uint32_t)n) >> 32);
  }
  indices[(size_t)w * k + j] = v;
}

// eval.py:166-171 (channels_last = 1: src [b, P, ch], one row per point) / shapenet eval.py:158-160 (channels_last = 0:
// src [b, ch, P]): out[(w*extra + e), c, j] = src[w, indices[w, e*npo + j], c].  One thread per voted point: the
// writes of a warp are 128 contiguous bytes per channel; the reads are one short row per thread (L2 resident windows).
// Optional labels [b, P] -> out_labels [b, extra*npo] (datasets/s3dis.py:89).
__global__ void __launch_bounds__(256) vote_gather_kernel(int ch, int P, int extra, int npo, int channels_last,
                                                          const float *__restrict__ src,
                                                          const int *__restrict__ indices,
                                                          float *__restrict__ out, const int *__restrict__ labels,
                                                          int *__restrict__ out_labels) {
  const int w = blockIdx.y;
  const int nv = extra * npo;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nv) return;
  int i = indices[(size_t)w * nv + p];
  i = min(max(i, 0), P - 1);
  const int e = p / npo, j = p - e * npo;
  float *o = out + ((size_t)(w * extra + e) * ch) * npo + j;
  if (channels_last) {
    const float *s = src + ((size_t)w * P + i) * ch;
    for (int c = 0; c < ch; ++c) o[(size_t)c * npo] = __ldg(s + c);
  } else {
    const float *s = src + (size_t)w * ch * P + i;
    for (int c = 0; c < ch; ++c) o[(size_t)c * npo] = __ldg(s + (size_t)c * P);
  }
  if (labels != nullptr) out_labels[(size_t)w * nv + p] = labels[(size_t)w * P + i];
}

// eval.py:173: F.softmax(model(inputs), dim=1).max(dim=1)  /  shapenet eval.py:162-165 with the class range of the
// shape.  softmax as torch evaluates it: exp(x - max) / sum; the arg-max keeps the FIRST maximal class (torch.max).
// logits [b, c, n] -> conf [b, n], pred [b, n] (class index in [c0, c1)).  One thread per point, coalesced over n.
__global__ void __launch_bounds__(256) softmax_max_kernel(int c, int n, int c0, int c1,
                                                          const float *__restrict__ logits,
                                                          float *__restrict__ conf, int *__restrict__ pred) {
  const int bi = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float *x = logits + (size_t)bi * c * n + p;
  float m = x[0];
  for (int k = 1; k < c; ++k) m = fmaxf(m, x[(size_t)k * n]);
  float s = 0.0f;
  for (int k = 0; k < c; ++k) s += expf(x[(size_t)k * n] - m);
  int best = c0;
  float bv = expf(x[(size_t)c0 * n] - m) / s;
  for (int k = c0 + 1; k < c1; ++k) {
    const float v = expf(x[(size_t)k * n] - m) / s;
    if (v > bv) { bv = v; best = k; }
  }
  conf[(size_t)bi * n + p] = bv;
  pred[(size_t)bi * n + p] = best;
}

__global__ void __launch_bounds__(256) vote_reset_kernel(long long n, unsigned long long *__restrict__ keys,
                                                         int *__restrict__ pred) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  keys[i] = 0ull;   // confidence 0 (eval.py:136)
  pred[i] = -1;     // eval.py:137
}

// key of a vote: positive floats order like their bit patterns; the complemented sequence number makes the earliest
// vote the largest among equal confidences (the numba loop only replaces on strictly larger confidence)
__device__ __forceinline__ unsigned long long vt_key(float conf, unsigned int order) {
  return ((unsigned long long)__float_as_uint(conf) << 32) | (unsigned long long)(0xFFFFFFFFu - order);
}

__device__ __forceinline__ long long vt_scene_index(int w, int p, int nv, int P, const int *indices,
                                                    const int *mapping) {
  const int i = indices[(size_t)w * nv + p];
  if (mapping == nullptr) return (long long)i;                 // shapenet: the shuffled index is the point
  if (i < 0 || i >= P) return -1;
  return (long long)mapping[(size_t)w * P + i];                // eval.py:199
}

// ASSIGN = false: eval.py:200-202 as an atomic max;  ASSIGN = true (launched after it): the vote whose key won the
// point writes its class (:203); keys are unique per vote, so there is exactly one writer per point and launch.
template <bool ASSIGN>
__global__ void __launch_bounds__(256) vote_merge_kernel(int nv, int P, long long scene_points, unsigned int order_base,
                                                         const float *__restrict__ conf,
                                                         const int *__restrict__ pred,
                                                         const int *__restrict__ indices,
                                                         const int *__restrict__ mapping,
                                                         unsigned long long *__restrict__ keys,
                                                         int *__restrict__ scene_pred) {
  const int w = blockIdx.y;
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nv) return;
  const float cf = conf[(size_t)w * nv + p];
  if (!(cf > 0.0f)) return;   // never beats the initial confidence 0 (also NaN)
  const long long pt = vt_scene_index(w, p, nv, P, indices, mapping);
  if (pt < 0 || pt >= scene_points) return;
  const unsigned long long key = vt_key(cf, order_base + (unsigned int)w * (unsigned int)nv + (unsigned int)p);
  if (!ASSIGN) {
    atomicMax(keys + pt, key);
  } else {
    if (keys[pt] == key) scene_pred[pt] = pred[(size_t)w * nv + p];
  }
}

__global__ void __launch_bounds__(256) vote_confidences_kernel(long long n,
                                                               const unsigned long long *__restrict__ keys,
                                                               float *__restrict__ conf) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  conf[i] = __uint_as_float((unsigned int)(keys[i] >> 32));
}

// eval.py:207-215: stats[0, gt]++, stats[1, pd]++, stats[2, gt]++ iff gt == pd.  A point that never received a vote
// keeps pd = -1, which numba's wrap-around indexing counts in the LAST class of row 1; reproduced when wrap_unvoted
// is set (shapenet's update_stats, eval.py:188-201, tests `predictions == i` and so ignores it: wrap_unvoted = 0).
// Grid-stride CTAs with a shared-memory histogram, flushed once per CTA.
__global__ void __launch_bounds__(256) vote_stats_kernel(long long n, int num_classes, int wrap_unvoted,
                                                         const int *__restrict__ gt,
                                                         const int *__restrict__ pred,
                                                         unsigned long long *__restrict__ stats) {
  extern __shared__ unsigned int vt_hist[];   // [3 * num_classes]
  for (int t = threadIdx.x; t < 3 * num_classes; t += blockDim.x) vt_hist[t] = 0u;
  __syncthreads();
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int g = gt[i];
    int pd = pred[i];
    if (g >= 0 && g < num_classes) {
      atomicAdd(&vt_hist[g], 1u);
      if (g == pd) atomicAdd(&vt_hist[2 * num_classes + g], 1u);
    }
    if (pd < 0 && wrap_unvoted) pd += num_classes;
    if (pd >= 0 && pd < num_classes) atomicAdd(&vt_hist[num_classes + pd], 1u);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < 3 * num_classes; t += blockDim.x) {
    const unsigned int v = vt_hist[t];
    if (v) atomicAdd(stats + t, (unsigned long long)v);
  }
}

}  // namespace pvb

extern "C" {

int pvcnn_vote_indices(int b, int nv, unsigned long long seed, int first_window, const int *num_points, int *indices,
                       void *stream) {
  PVB_CHECK_ARG(b > 0 && b <= 65535 && nv > 0 && first_window >= 0 && num_points && indices);
  PVB_LAUNCH(pvb::vote_indices_kernel, dim3(pvb::ceil_div(nv, 256), b), 256, 0, stream, nv, seed, first_window,
             num_points, indices);
  return 0;
}

int pvcnn_window_indices(int b, int k, unsigned long long seed, int first_window, const int *num_points, int *indices,
                         void *stream) {
  PVB_CHECK_ARG(b > 0 && b <= 65535 && k > 0 && first_window >= 0 && num_points && indices);
  PVB_LAUNCH(pvb::window_indices_kernel, dim3(pvb::ceil_div(k, 256), b), 256, 0, stream, k, seed, first_window,
             num_points, indices);
  return 0;
}

int pvcnn_vote_gather(int b, int ch, int p, int extra, int npo, int channels_last, const float *src,
                      const int *indices, float *out, const int *labels, int *out_labels, void *stream) {
  PVB_CHECK_ARG(b > 0 && b <= 65535 && ch > 0 && p > 0 && extra > 0 && npo > 0 && src && indices && out);
  PVB_CHECK_ARG((labels == nullptr) == (out_labels == nullptr));
  PVB_CHECK_ARG((long long)extra * npo <= 0x7fffffffLL && (long long)b * extra <= 0x7fffffffLL);
  PVB_LAUNCH(pvb::vote_gather_kernel, dim3(pvb::ceil_div((long long)extra * npo, 256), b), 256, 0, stream, ch, p,
             extra, npo, channels_last, src, indices, out, labels, out_labels);
  return 0;
}

int pvcnn_softmax_max(int b, int c, int n, int c0, int c1, const float *logits, float *conf, int *pred,
                      void *stream) {
  PVB_CHECK_ARG(b > 0 && b <= 65535 && c > 0 && n > 0 && c0 >= 0 && c0 < c1 && c1 <= c && logits && conf && pred);
  PVB_LAUNCH(pvb::softmax_max_kernel, dim3(pvb::ceil_div(n, 256), b), 256, 0, stream, c, n, c0, c1, logits, conf, pred);
  return 0;
}

int pvcnn_vote_reset(long long scene_points, unsigned long long *keys, int *scene_pred, void *stream) {
  PVB_CHECK_ARG(scene_points > 0 && scene_points <= 0x7fffffffLL * 256 && keys && scene_pred);
  PVB_LAUNCH(pvb::vote_reset_kernel, pvb::ceil_div(scene_points, 256), 256, 0, stream, scene_points, keys, scene_pred);
  return 0;
}

int pvcnn_vote_merge(int b, int nv, int p, long long scene_points, unsigned int order_base, const float *conf,
                     const int *pred, const int *indices, const int *mapping, unsigned long long *keys,
                     int *scene_pred, void *stream) {
  PVB_CHECK_ARG(b > 0 && b <= 65535 && nv > 0 && p > 0 && scene_points > 0 && conf && pred && indices && keys &&
                scene_pred);
  // sequence numbers order_base .. order_base + b*nv - 1 must fit 32 bits
  PVB_CHECK_ARG((unsigned long long)order_base + (unsigned long long)b * (unsigned long long)nv <= 0x100000000ull);
  const dim3 grid(pvb::ceil_div(nv, 256), b);
  PVB_LAUNCH(pvb::vote_merge_kernel<false>, grid, 256, 0, stream, nv, p, scene_points, order_base, conf, pred, indices,
             mapping, keys, scene_pred);
  PVB_LAUNCH(pvb::vote_merge_kernel<true>, grid, 256, 0, stream, nv, p, scene_points, order_base, conf, pred, indices,
             mapping, keys, scene_pred);
  return 0;
}

int pvcnn_vote_confidences(long long scene_points, const unsigned long long *keys, float *conf, void *stream) {
  PVB_CHECK_ARG(scene_points > 0 && scene_points <= 0x7fffffffLL * 256 && keys && conf);
  PVB_LAUNCH(pvb::vote_confidences_kernel, pvb::ceil_div(scene_points, 256), 256, 0, stream, scene_points, keys, conf);
  return 0;
}

int pvcnn_vote_stats(long long n, int num_classes, int wrap_unvoted, const int *gt, const int *pred,
                     unsigned long long *stats, void *stream) {
  PVB_CHECK_ARG(n > 0 && num_classes > 0 && num_classes <= 2048 && gt && pred && stats);
  const size_t smem = sizeof(unsigned int) * 3 * (size_t)num_classes;   // <= 24 KB
  long long ctas = (n + 256 * 16 - 1) / (256 * 16);                     // >= 16 points per thread before the flush
  if (ctas > pvb::kNumSMs * 8) ctas = pvb::kNumSMs * 8;
  if (ctas < 1) ctas = 1;
  PVB_LAUNCH(pvb::vote_stats_kernel, (int)ctas, 256, smem, stream, n, num_classes, wrap_unvoted, gt, pred,
             stats);
  return 0;
}

}  // extern "C"
