// pvconv_pipeline.cu -- the fused PVConv block (forward + backward) as two C-ABI calls.
// Orchestrates the channels-last kernels of fused_ops.cu around the tcgen05 GEMMs of
// conv_igemm.cu / conv_wgrad.cu.  See include/pvcnn_b200.h for the contract and DESIGN.md for the
// dataflow; reference: modules/pvconv.py:33-39 (forward wiring) and torch autograd (backward).
#include <cstdlib>

#include "fused_ops.cuh"

namespace pvb {
int igemm_launch(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                 int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                 int npass, cudaStream_t stream);
int conv_halo_launch(int nb, int sx, int sy, int sz, int k, int cout, const float *a, int lda, const float *w_hi,
                     const float *w_lo, int ldw, const float *bias, float *out, int ldo, int npass, cudaStream_t stream,
                     const int4 *unit_list, const int *unit_count);
int wgrad_launch(int nb, int sx, int sy, int sz, int cin, int cout, int ntaps, const float *x_hi, const float *x_lo,
                 int ldx, const float *g_hi, const float *g_lo, int ldg, float *dw, int npass, cudaStream_t s,
                 const int4 *ktile_list, const int *ktile_count, int *bz_out, int *by_out);

bool conv_halo_supported(int sx, int sy, int sz, int cout);  // conv_halo.cu
int conv_halo_ty(int sz);                                    // conv_halo.cu

// `lo` GRID tensors (x - trunc_tf32(x)) are materialised in 3xTF32 mode: the wgrad kernel is shared-memory-bandwidth
// bound and runs 1.6x faster when it TMA-loads lo instead of converting it in the kernel; the v1 conv kernel
// (odd resolutions, > 64 channels) needs them too.  The halo conv kernel ignores them.
static bool needs_grid_lo(const pvcnn_pvconv_desc *d) { return d->npass > 1; }

// ---- activity-driven skipping: layout of ws->sparse (ints) --------------------------------------------
struct SparseBuf {
  int *counts;               // [8]
  unsigned char *occ, *act1;
  int4 *fwd1, *dgrad1, *fwd2, *wg1, *wg2, *dg2;
  unsigned char *act_dg;     // per unit: contains an occupied column
  unsigned char *fwd2_flag, *dg2_flag, *wg1_flag;
  int max_units;
  unsigned char *wg2_flag;   // per k-tile: conv2-wgrad tile is computed on the tensor cores (else closed form)
  float *classsum;           // [27][co]  conv2 forward constants
  float *tapsum;             // [27][co]
  float *classsum_g;         // [2][27][co]  class sums of gY2: constant-region k-tiles, all voxels
  int *chunk_counts;         // [6][chunks]  per-1024-flag chunk counts of the list compaction
  int ty, wg_bz, wg_by;
  long long total_ints;
};
static inline long long up4(long long x) { return (x + 3) / 4 * 4; }
static SparseBuf sparse_at(int *base, int b, int r, int co) {
  SparseBuf v{};
  v.ty = conv_halo_ty(r);            // y rows per halo-conv tile (conv_halo.cu)
  v.wg_bz = ((r < 32 ? r : 32) + 7) / 8 * 8;   // must mirror conv_wgrad.cu's k-tile box
  v.wg_by = 32 / v.wg_bz > 0 ? 32 / v.wg_bz : 1;
  if (v.wg_by > r) v.wg_by = r;
  const long long units = (long long)b * ((r + 1) / 2) * ((r + v.ty - 1) / v.ty);
  const long long kt = (long long)b * r * ((r + v.wg_by - 1) / v.wg_by) * ((r + v.wg_bz - 1) / v.wg_bz);
  long long o = 0;
  v.counts = base + o; o += 8;
  v.occ = reinterpret_cast<unsigned char *>(base + o); o += up4((long long)b * r * r) / 4 + 4;
  v.act1 = reinterpret_cast<unsigned char *>(base + o); o += up4(units) / 4 + 4;
  v.act_dg = reinterpret_cast<unsigned char *>(base + o); o += up4(units) / 4 + 4;
  v.fwd2_flag = reinterpret_cast<unsigned char *>(base + o); o += up4(units) / 4 + 4;
  v.dg2_flag = reinterpret_cast<unsigned char *>(base + o); o += up4(units) / 4 + 4;
  v.wg1_flag = reinterpret_cast<unsigned char *>(base + o); o += up4(kt) / 4 + 4;
  v.max_units = (int)units;
  o = up4(o);
  v.fwd1 = reinterpret_cast<int4 *>(base + o); o += 4 * units;
  v.dgrad1 = reinterpret_cast<int4 *>(base + o); o += 4 * units;
  v.fwd2 = reinterpret_cast<int4 *>(base + o); o += 4 * units;
  v.dg2 = reinterpret_cast<int4 *>(base + o); o += 4 * units;
  v.wg1 = reinterpret_cast<int4 *>(base + o); o += 4 * kt;
  v.wg2 = reinterpret_cast<int4 *>(base + o); o += 4 * kt;
  v.wg2_flag = reinterpret_cast<unsigned char *>(base + o); o += up4(kt) / 4 + 4;
  o = up4(o);
  v.classsum = reinterpret_cast<float *>(base + o); o += 27LL * co;
  v.tapsum = reinterpret_cast<float *>(base + o); o += 27LL * co;
  v.classsum_g = reinterpret_cast<float *>(base + o); o += 2 * 27LL * co;
  v.chunk_counts = base + o; o += 6 * (((units > kt ? units : kt) + 1023) / 1024) + 4;
  v.total_ints = o;
  return v;
}
// Skipping applies when every 3x3x3 conv of the block runs on the halo kernel (its unit geometry) and r >= 4.
static bool sparse_enabled(const pvcnn_pvconv_desc *d) {
  const char *e = getenv("PVCNN_B200_SPARSE");
  if (e && e[0] == '0') return false;
  const char *v = getenv("PVCNN_B200_CONV");
  if (v && v[0] == 'v' && v[1] == '1') return false;
  return d->r >= 4 && conv_halo_supported(d->r, d->r, d->r, d->cout) && conv_halo_supported(d->r, d->r, d->r, d->cin);
}

static inline int pad4(int x) { return (x + 3) / 4 * 4; }
static inline int ld32(int x) { return (x + 31) / 32 * 32; }

struct WPrep {  // offsets (floats) into ws->wprep; each entry holds hi then lo
  long long w1f, w2f, wpf, w1d, w2d, wpd, total;
  long long n1f, n2f, npf, n1d, n2d, npd;  // floats of ONE copy (hi)
};
static WPrep wprep_layout(const pvcnn_pvconv_desc *d) {
  WPrep w{};
  w.n1f = 27LL * d->cout * ld32(d->cin);
  w.n2f = 27LL * d->cout * ld32(d->cout);
  w.npf = 1LL * d->cout * ld32(d->cin);
  w.n1d = 27LL * d->cin * ld32(d->cout);
  w.n2d = 27LL * d->cout * ld32(d->cout);
  w.npd = 1LL * d->cin * ld32(d->cout);
  long long o = 0;
  w.w1f = o; o += 2 * w.n1f;
  w.w2f = o; o += 2 * w.n2f;
  w.wpf = o; o += 2 * w.npf;
  w.w1d = o; o += 2 * w.n1d;
  w.w2d = o; o += 2 * w.n2d;
  w.wpd = o; o += 2 * w.npd;
  w.total = o;
  return w;
}

struct SeBuf {  // views into ws->se
  float *pooled3, *mean, *hidden, *gate, *dgate, *extra;
  int hid;
};
static SeBuf se_at(float *base, int b, int co, int cout) {
  SeBuf v;
  v.hid = cout / 8;
  v.pooled3 = base;
  v.mean = v.pooled3 + (size_t)b * 3 * co;
  v.gate = v.mean + (size_t)b * co;
  v.dgate = v.gate + (size_t)b * co;
  v.extra = v.dgate + (size_t)b * co;
  v.hidden = v.extra + (size_t)b * co;
  return v;
}

static BnCoef coef_at(float *base, int idx, int co) {
  float *p = base + (size_t)idx * 4 * co;
  return BnCoef{p, p + co, p + 2 * co, p + 3 * co};
}
}  // namespace pvb

using namespace pvb;

extern "C" int pvcnn_conv_weight_prep(int cout, int cin, int ntaps, int mode, int ld, const float *w, float *w_hi,
                                      float *w_lo, void *stream);

#define PVB_TRY(expr)            \
  do {                           \
    int rc__ = (expr);           \
    if (rc__ != 0) return rc__;  \
  } while (0)

extern "C" {

long long pvcnn_pvconv_wprep_floats(const pvcnn_pvconv_desc *d) { return wprep_layout(d).total; }

/* ws->prep (inference): [3 BatchNorm coefficient sets: 12*co][conv2 class constants 27*co][tap sums 27*co][forward GEMM
 * operands of conv1, conv2 and the point branch] */
long long pvcnn_pvconv_prep_floats(const pvcnn_pvconv_desc *d) {
  const long long co = (d->cout + 3) / 4 * 4;
  return 66 * co + wprep_layout(d).w1d;
}

int pvcnn_pvconv_needs_grid_lo(const pvcnn_pvconv_desc *d) { return needs_grid_lo(d) ? 1 : 0; }

long long pvcnn_pvconv_sparse_ints(const pvcnn_pvconv_desc *d) {
  return sparse_at(nullptr, d->b, d->r, (d->cout + 3) / 4 * 4).total_ints;
}

long long pvcnn_pvconv_partials_floats(const pvcnn_pvconv_desc *d) {
  const long long co = pad4(d->cout > d->cin ? d->cout : d->cin);
  const long long blocks_pts = (long long)d->b * ((d->n + 31) / 32);
  long long blocks = blocks_pts > kNumSMs * 4 ? blocks_pts : kNumSMs * 4;
  const long long units = sparse_at(nullptr, d->b, d->r, (int)co).max_units;  // unit-list reductions use one block per unit
  if (units > blocks) blocks = units;
  return blocks * 5 * co;  // 4 reduction sets + the per-sample SE gate partials
}

int pvcnn_pvconv_forward(const pvcnn_pvconv_desc *d, const float *features, const float *coords,
                         const pvcnn_pvconv_params *prm, const pvcnn_pvconv_ws *ws, float *out, void *stream) {
  PVB_CHECK_ARG(d && features && coords && prm && ws && out);
  PVB_CHECK_ARG(d->b > 0 && d->n > 0 && d->cin > 0 && d->cout > 0 && d->r > 1 && (d->npass == 1 || d->npass == 3));
  cudaStream_t s = (cudaStream_t)stream;
  const int b = d->b, n = d->n, r = d->r, r3 = r * r * r;
  const int ci = pad4(d->cin), co = pad4(d->cout);
  const long long Mv = (long long)b * r3, Mp = (long long)b * n;
  const bool lo = d->npass > 1;        // point-set lo tensors (1x1 GEMM runs on the v1 kernel)
  const bool glo = needs_grid_lo(d);   // grid lo tensors
  PVB_CHECK_ARG(!lo || (ws->fcl_lo != nullptr));
  PVB_CHECK_ARG(!glo || (ws->g0_lo && ws->z1_lo));
  const WPrep W = wprep_layout(d);
  float *wp = ws->wprep;
  int nblk = 0;

  // 1. coordinates -> voxel indices                               (modules/voxelization.py:17-24, vox.cu:18-34)
  if (d->vox_stats >= 1) {  // reference-exact: the mean is torch's own reduction result
    PVB_CHECK_ARG(ws->vox_mean && (ws->vox_denom || !d->normalize));
    if (d->normalize && d->vox_stats == 1) PVB_TRY(pvcnn_voxelize_denom(b, n, d->eps, coords, ws->vox_mean, ws->vox_denom, stream));
    PVB_TRY(pvcnn_voxelize_apply(b, n, r, d->normalize, coords, ws->vox_mean, ws->vox_denom, ws->nc, ws->vc, stream));
  } else {
    PVB_TRY(pvcnn_voxelize_coords(b, n, r, d->normalize, d->eps, coords, ws->nc, ws->vc, stream));
  }
  PVB_TRY(launch_vox_index_count(b, n, r, ws->vc, ws->ind, ws->cnt, s));
  const bool sparse = sparse_enabled(d) && ws->sparse != nullptr;
  SparseBuf sp{};
  if (sparse) {  // which tiles can differ from the closed form (zero / constant input)?
    sp = sparse_at(ws->sparse, b, r, co);
    PVB_TRY(launch_build_activity(b, r, sp.ty, sp.wg_bz, sp.wg_by, ws->cnt, sp.counts, sp.occ, sp.act1, sp.act_dg, sp.fwd2_flag,
                                  sp.wg1_flag, sp.wg2_flag, sp.dg2_flag, sp.fwd1, sp.dgrad1, sp.fwd2, sp.wg1, sp.wg2, sp.dg2,
                                  sp.chunk_counts, s));
  }
  // 2. points to channels-last, scatter-mean into the grid        (vox.cu:48-72)
  PVB_TRY(launch_points_to_cl(b, d->cin, n, ci, features, ws->fcl, lo ? ws->fcl_lo : nullptr, s));
  PVB_TRY(launch_memset_f32(ws->g0, Mv * ci, s));
  PVB_TRY(launch_voxelize_cl(b, n, r3, ci, ws->ind, ws->cnt, ws->fcl, ws->g0, s));
  if (glo) {
    PVB_TRY(launch_memset_f32(ws->g0_lo, Mv * ci, s));
    PVB_TRY(launch_grid_lo_at_points(b, n, r3, ci, ws->ind, ws->g0, ws->g0_lo, s));
  }
  // 3. weights -> GEMM operands.  Inference with ws->prep: operands, BatchNorm coefficients and the conv2 constant tables
  //    live in the caller's per-block buffer and are rebuilt only when desc.prepared == 0 (parameters changed).
  const bool frozen = !d->training && ws->prep != nullptr;
  const bool have_prep = frozen && d->prepared;
  float *coef_base = frozen ? ws->prep : ws->coef;
  float *cls_tab = frozen ? ws->prep + 12 * (size_t)co : sp.classsum;
  float *tap_tab = frozen ? ws->prep + 39 * (size_t)co : sp.tapsum;
  if (frozen) wp = ws->prep + 66 * (size_t)co;
  if (!have_prep) {
    PVB_TRY(pvcnn_conv_weight_prep(d->cout, d->cin, 27, 0, ld32(d->cin), prm->w1, wp + W.w1f, wp + W.w1f + W.n1f, stream));
    PVB_TRY(pvcnn_conv_weight_prep(d->cout, d->cout, 27, 0, ld32(d->cout), prm->w2, wp + W.w2f, wp + W.w2f + W.n2f, stream));
    PVB_TRY(pvcnn_conv_weight_prep(d->cout, d->cin, 1, 0, ld32(d->cin), prm->wp, wp + W.wpf, wp + W.wpf + W.npf, stream));
  }

  BnCoef bn1 = coef_at(coef_base, 0, co), bn2 = coef_at(coef_base, 1, co), bnp = coef_at(coef_base, 2, co);
  // 4. conv1 -> BN1 -> LeakyReLU                                   (modules/pvconv.py:21-23)
  if (sparse) {  // zero neighbourhood -> conv1 = bias; only the listed units run on the tensor cores
    PVB_TRY(launch_fill_bias_rows(Mv, d->cout, co, prm->b1, ws->y1, s));
    PVB_TRY(conv_halo_launch(b, r, r, r, d->cin, d->cout, ws->g0, ci, wp + W.w1f, wp + W.w1f + W.n1f, ld32(d->cin), prm->b1,
                             ws->y1, co, d->npass, s, sp.fwd1, sp.counts + 0));
  } else {
    PVB_TRY(igemm_launch(b, r, r, r, d->cin, d->cout, 27, ws->g0, ws->g0_lo, ci, wp + W.w1f, wp + W.w1f + W.n1f,
                         ld32(d->cin), prm->b1, ws->y1, co, d->npass, s));
  }
  if (d->training) {
    PVB_TRY(launch_bn_stats(Mv, co, ws->y1, ws->partials, &nblk, s));
    PVB_TRY(launch_bn_finalize(nblk, d->cout, co, Mv, d->bn_eps_vox, d->momentum, ws->partials, prm->g1, prm->be1,
                               prm->rm1, prm->rv1, bn1, s, prm->nbt1));
  } else if (!have_prep) {
    PVB_TRY(launch_bn_coef_from_running(d->cout, d->bn_eps_vox, prm->g1, prm->be1, prm->rm1, prm->rv1, bn1, s));
  }
  PVB_TRY(launch_bn_apply_leaky(Mv, co, d->slope, ws->y1, bn1, ws->z1, glo ? ws->z1_lo : nullptr, s));
  // 5. conv2 -> BN2 statistics (BN2-apply + LeakyReLU are folded into the devoxelize gather)
  if (sparse) {  // constant neighbourhood -> conv2 = one of 27 boundary-class constants
    PVB_TRY(launch_fill_const_conv(b, r, d->cout, d->cout, co, d->slope, prm->w2, prm->b2, prm->b1, bn1, cls_tab, tap_tab,
                                   ws->y2, s, have_prep ? 1 : 0));
    PVB_TRY(conv_halo_launch(b, r, r, r, d->cout, d->cout, ws->z1, co, wp + W.w2f, wp + W.w2f + W.n2f, ld32(d->cout),
                             prm->b2, ws->y2, co, d->npass, s, sp.fwd2, sp.counts + 2));
  } else {
    PVB_TRY(igemm_launch(b, r, r, r, d->cout, d->cout, 27, ws->z1, ws->z1_lo, co, wp + W.w2f, wp + W.w2f + W.n2f,
                         ld32(d->cout), prm->b2, ws->y2, co, d->npass, s));
  }
  if (d->training) {
    PVB_TRY(launch_bn_stats(Mv, co, ws->y2, ws->partials, &nblk, s));
    PVB_TRY(launch_bn_finalize(nblk, d->cout, co, Mv, d->bn_eps_vox, d->momentum, ws->partials, prm->g2, prm->be2,
                               prm->rm2, prm->rv2, bn2, s, prm->nbt2));
  } else if (!have_prep) {
    PVB_TRY(launch_bn_coef_from_running(d->cout, d->bn_eps_vox, prm->g2, prm->be2, prm->rm2, prm->rv2, bn2, s));
  }
  // 6. point branch: 1x1 conv as a GEMM over all points             (modules/shared_mlp.py:10-12)
  PVB_TRY(igemm_launch(1, 1, 1, (int)Mp, d->cin, d->cout, 1, ws->fcl, ws->fcl_lo, ci, wp + W.wpf, wp + W.wpf + W.npf,
                       ld32(d->cin), prm->bp, ws->p, co, d->npass, s));
  if (d->training) {
    PVB_TRY(launch_bn_stats(Mp, co, ws->p, ws->partials, &nblk, s));
    PVB_TRY(launch_bn_finalize(nblk, d->cout, co, Mp, d->bn_eps_pt, d->momentum, ws->partials, prm->gp, prm->bep,
                               prm->rmp, prm->rvp, bnp, s, prm->nbtp));
  } else if (!have_prep) {
    PVB_TRY(launch_bn_coef_from_running(d->cout, d->bn_eps_pt, prm->gp, prm->bep, prm->rmp, prm->rvp, bnp, s));
  }
  // 6b. SE3d gate from the pooled (post BN2 + LeakyReLU) grid                  (modules/se.py:16-17)
  const float *gate = nullptr;
  if (d->with_se) {
    PVB_CHECK_ARG(ws->se && prm->se_w1 && prm->se_w2 && d->cout >= 8);
    SeBuf se = se_at(ws->se, b, co, d->cout);
    PVB_TRY(launch_se_pool(b, r3, co, d->slope, ws->y2, bn2, ws->partials, se.pooled3, s));
    PVB_TRY(launch_se_fc(b, d->cout, co, se.hid, r3, se.pooled3, prm->se_w1, prm->se_w2, se.mean, se.hidden, se.gate, s));
    gate = se.gate;
  }
  // 7. BN2-apply + LeakyReLU + trilinear devoxelize (+ SE gate) + BN1d-apply + ReLU + add + transpose
  PVB_TRY(launch_devox_fused(b, n, d->cout, co, r, d->slope, ws->nc, ws->y2, bn2, ws->p, bnp, gate, out, s));
  return 0;
}

int pvcnn_pvconv_backward(const pvcnn_pvconv_desc *d, const float *grad_out, const pvcnn_pvconv_params *prm,
                          const pvcnn_pvconv_ws *ws, float *grad_features, const pvcnn_pvconv_grads *gr,
                          void *stream) {
  return pvcnn_pvconv_backward_phase(d, grad_out, prm, ws, grad_features, gr, 0, stream);
}

int pvcnn_pvconv_backward_phase(const pvcnn_pvconv_desc *d, const float *grad_out, const pvcnn_pvconv_params *prm,
                                const pvcnn_pvconv_ws *ws, float *grad_features, const pvcnn_pvconv_grads *gr,
                                int phase, void *stream) {
  PVB_CHECK_ARG(d && grad_out && prm && ws && grad_features && gr && d->training && phase >= 0 && phase <= 2);
  cudaStream_t s = (cudaStream_t)stream;
  const int b = d->b, n = d->n, r = d->r, r3 = r * r * r;
  const int ci = pad4(d->cin), co = pad4(d->cout);
  const long long Mv = (long long)b * r3, Mp = (long long)b * n;
  const bool lo = d->npass > 1;
  const bool glo = needs_grid_lo(d);
  PVB_CHECK_ARG(!glo || (ws->gy2_lo && ws->gy1_lo));
  const WPrep W = wprep_layout(d);
  float *wp = ws->wprep;
  BnCoef bn1 = coef_at(ws->coef, 0, co), bn2 = coef_at(ws->coef, 1, co), bnp = coef_at(ws->coef, 2, co);
  const bool sparse = sparse_enabled(d) && ws->sparse != nullptr;  // lists were built by the forward pass
  SparseBuf sp{};
  if (sparse) sp = sparse_at(ws->sparse, b, r, co);
  float *S = ws->sums;  // [16][co]: 0 S1, 1 S2, 2 T1, 3 T2, 4 U1, 5 U2, 6.. column sums
  int nblk = 0;
  const size_t cb = sizeof(float) * (size_t)d->cout;

  if (phase != 2) {  // ---- phase 1: everything that produces a PARAMETER gradient (the last one is dW1)
  // 1. per-point stage: point-branch ReLU mask + reductions; voxel-branch scatter (x leaky') + reductions
  PVB_TRY(launch_memset_f32(ws->d2, Mv * co, s));
  SeBuf se{};
  float *ds_partials = nullptr;
  if (d->with_se) {
    PVB_CHECK_ARG(ws->se && prm->se_w1 && prm->se_w2 && gr->se_w1 && gr->se_w2);
    se = se_at(ws->se, b, co, d->cout);
    ds_partials = ws->partials + (size_t)b * ((n + 31) / 32) * 4 * co;
  }
  PVB_TRY(launch_bwd_points(b, n, d->cout, co, r, d->slope, grad_out, ws->nc, ws->y2, bn2, ws->p, bnp, ws->ga, ws->d2,
                            ws->partials, &nblk, d->with_se ? se.gate : nullptr, ds_partials, s));
  PVB_TRY(launch_reduce_partials(nblk, 4 * co, ws->partials, S, s));
  const float *extra = nullptr;
  if (d->with_se) {  // d gate -> FC backward -> dense d(mean) term folded into the BN2 reductions
    PVB_TRY(launch_reduce_partials_batched(b, (n + 31) / 32, co, ds_partials, se.dgate, s));
    PVB_TRY(launch_se_backward(b, d->cout, co, se.hid, r3, se.dgate, se.gate, se.hidden, se.mean, se.pooled3, prm->se_w1,
                               prm->se_w2, gr->se_w1, gr->se_w2, se.extra, S + 2 * co, S + 3 * co, s));
    extra = se.extra;
  }
  PVB_CUDA(cudaMemcpyAsync(gr->bep, S + 0 * co, cb, cudaMemcpyDeviceToDevice, s));
  PVB_CUDA(cudaMemcpyAsync(gr->gp, S + 1 * co, cb, cudaMemcpyDeviceToDevice, s));
  PVB_CUDA(cudaMemcpyAsync(gr->be2, S + 2 * co, cb, cudaMemcpyDeviceToDevice, s));
  PVB_CUDA(cudaMemcpyAsync(gr->g2, S + 3 * co, cb, cudaMemcpyDeviceToDevice, s));
  // 2. BN1d / BN3d(2) input gradients (+ conv-bias gradients as column sums)
  PVB_TRY(launch_bn_bwd_apply(Mp, co, 0, d->slope, ws->ga, ws->p, bnp, S + 0 * co, S + 1 * co, ws->gpp,
                              lo ? ws->gpp_lo : nullptr, ws->partials, &nblk, s));
  PVB_TRY(launch_reduce_partials(nblk, co, ws->partials, S + 6 * co, s));
  PVB_CUDA(cudaMemcpyAsync(gr->bp, S + 6 * co, cb, cudaMemcpyDeviceToDevice, s));
  PVB_TRY(launch_bn_bwd_apply(Mv, co, 0, d->slope, ws->d2, ws->y2, bn2, S + 2 * co, S + 3 * co, ws->gy2,
                              glo ? ws->gy2_lo : nullptr, ws->partials, &nblk, s, extra, r3));
  PVB_TRY(launch_reduce_partials(nblk, co, ws->partials, S + 7 * co, s));
  PVB_CUDA(cudaMemcpyAsync(gr->b2, S + 7 * co, cb, cudaMemcpyDeviceToDevice, s));
  // 3. data-gradient weights
  PVB_TRY(pvcnn_conv_weight_prep(d->cout, d->cin, 27, 1, ld32(d->cout), prm->w1, wp + W.w1d, wp + W.w1d + W.n1d, stream));
  PVB_TRY(pvcnn_conv_weight_prep(d->cout, d->cout, 27, 1, ld32(d->cout), prm->w2, wp + W.w2d, wp + W.w2d + W.n2d, stream));
  PVB_TRY(pvcnn_conv_weight_prep(d->cout, d->cin, 1, 1, ld32(d->cout), prm->wp, wp + W.wpd, wp + W.wpd + W.npd, stream));
  // 4. point branch: dgrad + wgrad
  PVB_TRY(igemm_launch(1, 1, 1, (int)Mp, d->cout, d->cin, 1, ws->gpp, ws->gpp_lo, co, wp + W.wpd, wp + W.wpd + W.npd,
                       ld32(d->cout), nullptr, ws->gfpt, ci, d->npass, s));
  PVB_TRY(wgrad_launch(1, 1, 1, (int)Mp, d->cin, d->cout, 1, ws->fcl, ws->fcl_lo, ci, ws->gpp, ws->gpp_lo, co, gr->wp,
                       d->npass, s, nullptr, nullptr, nullptr, nullptr));
  // 5. conv2: wgrad (needs z1, gy2) then dgrad into the d2 buffer (d2 was consumed in step 2)
  // conv2 weight gradient: tensor cores only on the k-tiles that can see a non-constant input voxel; the rest of
  // the grid (input == c1) contributes a rank-1 term computed from 27 boundary-class sums of gY2
  if (sparse) PVB_TRY(launch_class_sums(b, r, co, sp.wg_by, sp.wg_bz, sp.wg2_flag, ws->gy2, sp.classsum_g, s));
  PVB_TRY(wgrad_launch(b, r, r, r, d->cout, d->cout, 27, ws->z1, ws->z1_lo, co, ws->gy2, ws->gy2_lo, co, gr->w2,
                       d->npass, s, sparse ? sp.wg2 : nullptr, sparse ? sp.counts + 4 : nullptr, nullptr, nullptr));
  if (sparse)
    PVB_TRY(launch_wgrad_const_update(d->cout, d->cout, co, d->slope, sp.classsum_g, prm->b1, bn1, gr->w2, s));
  // conv2 data gradient: only on region G (units whose gY1 is consumed); BN1-backward treats the rest in closed form
  float *gz1 = ws->d2;
  if (sparse) {
    PVB_TRY(conv_halo_launch(b, r, r, r, d->cout, d->cout, ws->gy2, co, wp + W.w2d, wp + W.w2d + W.n2d, ld32(d->cout),
                             nullptr, gz1, co, d->npass, s, sp.dg2, sp.counts + 5));
    PVB_TRY(launch_conv_grad_total(d->cout, d->cout, co, prm->w2, sp.classsum_g + 27 * (size_t)co, S + 13 * co, s));
    // 6. LeakyReLU' + BN1 backward on G (explicit) + constant region (closed form)
    PVB_TRY(launch_bn_bwd_units(sp.max_units, r, sp.ty, d->cout, co, d->slope, Mv, sp.dg2, sp.counts + 5, gz1, ws->y1, bn1,
                                S + 13 * co, prm->b1, ws->partials, S + 9 * co, S + 4 * co, S + 5 * co, ws->gy1,
                                glo ? ws->gy1_lo : nullptr, S + 14 * co, S + 8 * co, s));
    PVB_CUDA(cudaMemcpyAsync(gr->be1, S + 4 * co, cb, cudaMemcpyDeviceToDevice, s));
    PVB_CUDA(cudaMemcpyAsync(gr->g1, S + 5 * co, cb, cudaMemcpyDeviceToDevice, s));
    PVB_CUDA(cudaMemcpyAsync(gr->b1, S + 8 * co, cb, cudaMemcpyDeviceToDevice, s));
  } else {
    PVB_TRY(igemm_launch(b, r, r, r, d->cout, d->cout, 27, ws->gy2, ws->gy2_lo, co, wp + W.w2d, wp + W.w2d + W.n2d,
                         ld32(d->cout), nullptr, gz1, co, d->npass, s));
    // 6. LeakyReLU' + BN1 backward
    PVB_TRY(launch_bn_bwd_reduce(Mv, co, d->slope, gz1, ws->y1, bn1, ws->partials, &nblk, s));
    PVB_TRY(launch_reduce_partials(nblk, 2 * co, ws->partials, S + 4 * co, s));
    PVB_CUDA(cudaMemcpyAsync(gr->be1, S + 4 * co, cb, cudaMemcpyDeviceToDevice, s));
    PVB_CUDA(cudaMemcpyAsync(gr->g1, S + 5 * co, cb, cudaMemcpyDeviceToDevice, s));
    PVB_TRY(launch_bn_bwd_apply(Mv, co, 1, d->slope, gz1, ws->y1, bn1, S + 4 * co, S + 5 * co, ws->gy1,
                                glo ? ws->gy1_lo : nullptr, ws->partials, &nblk, s));
    PVB_TRY(launch_reduce_partials(nblk, co, ws->partials, S + 8 * co, s));
    PVB_CUDA(cudaMemcpyAsync(gr->b1, S + 8 * co, cb, cudaMemcpyDeviceToDevice, s));
  }
  // 7. conv1: wgrad, dgrad (into the d2 buffer again: gz1 is dead)
  // conv1: its input G0 is zero outside the occupied columns -> only the listed k-tiles contribute to dW1, and the
  // data gradient is only consumed at occupied voxels -> only the listed units are produced (the rest of gg0 is never read)
  int wbz = 0, wby = 0;
  PVB_TRY(wgrad_launch(b, r, r, r, d->cin, d->cout, 27, ws->g0, ws->g0_lo, ci, ws->gy1, ws->gy1_lo, co, gr->w1,
                       d->npass, s, sparse ? sp.wg1 : nullptr, sparse ? sp.counts + 3 : nullptr, &wbz, &wby));
  if (sparse) PVB_CHECK_ARG(wbz == sp.wg_bz && wby == sp.wg_by);
  }  // phase 1
  if (phase == 1) return 0;
  // ---- phase 2: the input gradient (conv1 dgrad + scatter back to the points); a data-parallel caller launches its
  //      gradient all-reduce between the phases so that it overlaps these kernels
  float *gg0 = ws->d2;
  if (sparse) {
    PVB_TRY(conv_halo_launch(b, r, r, r, d->cout, d->cin, ws->gy1, co, wp + W.w1d, wp + W.w1d + W.n1d, ld32(d->cout),
                             nullptr, gg0, ci, d->npass, s, sp.dgrad1, sp.counts + 1));
  } else {
    PVB_TRY(igemm_launch(b, r, r, r, d->cout, d->cin, 27, ws->gy1, ws->gy1_lo, co, wp + W.w1d, wp + W.w1d + W.n1d,
                         ld32(d->cout), nullptr, gg0, ci, d->npass, s));
  }
  // 8. avg_voxelize backward + point-branch gradient              (vox.cu:86-110)
  PVB_TRY(launch_bwd_final(b, n, d->cin, ci, r3, ws->ind, ws->cnt, gg0, ws->gfpt, grad_features, s));
  return 0;
}

}  // extern "C"
