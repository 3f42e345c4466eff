// conv_halo.cu -- second-generation 3x3x3 implicit-GEMM convolution (forward / data gradient) for sm_100a.
//
// v1 (conv_igemm.cu) fetches one [128 voxel x 32 channel] tile per (tap, chunk) and is bound by the
// L2->SMEM pipe (ncu: 10.9 GB pulled from L2 per launch, tensor pipe 38 % active).  This kernel keeps a HALO
// block of the input in shared memory and serves the 9 (dy,dx) taps of a z-shift from it:
//
//   * a CTA produces 2 output tiles (two x-planes of [TY x BZ] = 128 voxels) per iteration;
//   * per phase (dz in {-1,0,1}, 16-channel chunk) ONE 5-D TMA box brings the (TX+2) x (TY+2) x BZ halo
//     (z pre-shifted by dz, out-of-bounds rows zero-filled = conv padding) into 64-byte-swizzled rows;
//     a tap (dy,dx) of tile t is then just a row offset of the UMMA descriptor (multiples of 8 rows, so the
//     swizzle phase is preserved) -- 18 tile-operands from one 768-row load instead of 18 loads of 128 rows;
//   * the `lo` halves of the 3xTF32 split are computed IN the kernel by 4 converter warps
//     (lo = x - trunc_tf32(x), generic-proxy writes + fence.proxy.async), so activations need no `lo`
//     tensor in HBM and the A-side L2 traffic halves again;
//   * weights stream through a 4-deep ring of [Cout x 16] tiles (hi, lo), one per tap and phase;
//   * accumulators: per tile two main chains (phases 0-5 / 6-11: <= 108 MMAs each, the tensor core
//     truncates when accumulating) and one chain for the correction terms; summed in fp32 RN in the epilogue.
//
// L2->SMEM traffic per conv at the metric shape: ~3.1 GB (v1: 10.9 GB).
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace pvb {
using namespace umma;

constexpr int HC_THREADS = 512;   // w0 TMA(A), w1 MMA, w2 TMEM alloc, w3 TMA(B), w4-11 epilogue (one tile per 4 warps), w12-15 converters
constexpr int HC_KC = 16;         // channels per phase (64-byte rows, SWIZZLE_64B)
constexpr int HC_TX = 2;          // output x-planes (tiles) per CTA iteration
constexpr int HC_BSTAGES = 4;     // weight-tile ring
constexpr uint32_t kLayoutSW64 = 4;

struct HaloParams {
  int nb, sx, sy, sz;             // sz = BZ (full z rows), 128 % sz == 0, sz % 8 == 0
  int ty;                         // y rows per tile = 128 / sz
  int tiles_y, pairs_x;           // sy / ty (ceil), sx / 2 (ceil)
  int num_units;                  // nb * pairs_x * tiles_y
  int kchunks;                    // ceil(k / 16)
  int cout, block_n;              // block_n = cout padded to 16 (<= 64)
  int npass;
  int ldo;
  uint32_t a_rows, a_bytes;       // halo rows, bytes of one halo copy (rows * 64)
  uint32_t b_bytes;               // bytes of one weight tile copy (block_n * 64)
  const float *bias;
  float *out;
  int *err;
  const int4 *unit_list;          // optional compact list of (x0, y0, b, -) units to compute; others are skipped
  const int *unit_count;          //   (device count) -- activity-driven tile skipping, see pvconv_pipeline.cu
  long long *dbg;                 // optional stall counters of CTA 0 (PVCNN_STALL_PROFILE=1)
  int exp;                        // experiment bits (PVCNN_HALO_EXP): 1 skip lo conversion, 4 skip MMA3, 8 no A loads, 16 no B loads, 32 no epilogue stores
};

__global__ void __launch_bounds__(HC_THREADS, 1)
    conv_halo_v2_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w_hi,
                     const __grid_constant__ CUtensorMap map_w_lo, const HaloParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t a_full[2], a_ready[2], a_empty[2], b_full[HC_BSTAGES], b_empty[HC_BSTAGES], acc_full, acc_empty;
  __shared__ uint32_t tmem_base_smem;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool three = p.npass > 1;
  // smem carve-up: [A buf0: hi, lo][A buf1: hi, lo][B ring: (hi, lo) x stages]
  const uint32_t a_buf_bytes = p.a_bytes * 2;
  uint8_t *smem_b = smem + 2 * a_buf_bytes;
  const uint32_t b_stage_bytes = p.b_bytes * 2;
  const int nphases = 3 * p.kchunks;
  const int half_phase = (nphases + 1) / 2;
  const uint32_t tmem_cols = 512;  // 2 tiles x (main0, main1, corr) x block_n <= 384 -> allocate all
  const int num_units = p.unit_list ? __ldg(p.unit_count) : p.num_units;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_w_hi);
    if (three) prefetch_tensormap(&map_w_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < HC_BSTAGES; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
    mbar_init(&acc_full, 1);
    mbar_init(&acc_empty, 256);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================================ TMA producer: activation halos ================================
    // (own thread, so the next phase's halo is requested as soon as its buffer frees up, independent of
    //  how far the weight ring has advanced)
    if (elect_one()) {
      int abuf = 0;
      uint32_t aphase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        int x0, y0, b;
        if (p.unit_list) {
          const int4 uc = __ldg(p.unit_list + unit);
          x0 = uc.x; y0 = uc.y; b = uc.z;
        } else {
          int u = unit;
          y0 = (u % p.tiles_y) * p.ty; u /= p.tiles_y;
          x0 = (u % p.pairs_x) * HC_TX; u /= p.pairs_x;
          b = u;
        }
        int dz = -1, cc = 0;
        for (int ph = 0; ph < nphases; ++ph) {
          mbar_wait(&a_empty[abuf], aphase ^ 1, p.err, 21);
          if (p.exp & 8) {  // experiment: no activation traffic
            mbar_arrive(&a_full[abuf]);
          } else {
            mbar_arrive_expect_tx(&a_full[abuf], p.a_bytes);
            tma_load_5d(smem + (size_t)abuf * a_buf_bytes, &map_a, &a_full[abuf], cc * HC_KC, dz, y0 - 1, x0 - 1, b);
          }
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
          if (++cc == p.kchunks) { cc = 0; ++dz; }
        }
      }
    }
  } else if (warp == 3) {
    // ================================ TMA producer: weight tiles ================================
    if (elect_one()) {
      int bst = 0;
      uint32_t bphase = 0;
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        int dz = -1, cc = 0;
        for (int ph = 0; ph < nphases; ++ph) {
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {  // taps (dx, dy) of this dz
            const int dx = t9 / 3 - 1, dy = t9 % 3 - 1;
            const int tap = (dx + 1) * 9 + (dy + 1) * 3 + (dz + 1);
            mbar_wait(&b_empty[bst], bphase ^ 1, p.err, 22);
            uint8_t *sb = smem_b + (size_t)bst * b_stage_bytes;
            if (p.exp & 16) {  // experiment: no weight traffic
              mbar_arrive(&b_full[bst]);
            } else {
              mbar_arrive_expect_tx(&b_full[bst], three ? 2 * p.b_bytes : p.b_bytes);
              tma_load_3d(sb, &map_w_hi, &b_full[bst], cc * HC_KC, 0, tap);
              if (three) tma_load_3d(sb + p.b_bytes, &map_w_lo, &b_full[bst], cc * HC_KC, 0, tap);
            }
            if (++bst == HC_BSTAGES) { bst = 0; bphase ^= 1; }
          }
          if (++cc == p.kchunks) { cc = 0; ++dz; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, p.block_n, 0, 0);
      constexpr uint32_t dhi = desc_hi32(512, kLayoutSW64);
      // descriptor offsets (16-byte units) of tile t / tap t9 inside the halo: rows are (x_local, y_local, z)
      uint32_t tap_off[9][HC_TX];
#pragma unroll
      for (int t9 = 0; t9 < 9; ++t9)
#pragma unroll
        for (int t = 0; t < HC_TX; ++t)
          tap_off[t9][t] = (uint32_t)(((t + t9 / 3) * (p.ty + 2) + (t9 % 3)) * p.sz) * (HC_KC * 4 / 16);
      const uint32_t bn = (uint32_t)p.block_n;
      int abuf = 0, bst = 0, it = 0;
      uint32_t aphase = 0, bphase = 0;
      long long st_acc = 0, st_a = 0, st_b = 0;
      const long long t_begin = clock64();
      for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
        mbar_wait_t(&acc_empty, (uint32_t)((it & 1) ^ 1), p.err, 23, st_acc);
        tc_fence_after();
        for (int ph = 0; ph < nphases; ++ph) {
          mbar_wait_t(&a_ready[abuf], aphase, p.err, 24, st_a);
          tc_fence_after();
          const uint32_t a_hi = desc_lo32(smem_u32(smem + (size_t)abuf * a_buf_bytes), 0);
          const uint32_t a_lo = a_hi + (p.a_bytes >> 4);
          const uint32_t slot_col = ph < half_phase ? 0u : bn;
          const uint32_t fresh_main = ((ph == 0) || (ph == half_phase)) ? 0u : 1u;
          const uint32_t fresh_corr = ph == 0 ? 0u : 1u;
          // software-pipelined barrier polling: the try_wait for the NEXT weight tile is issued before this
          // tile's MMAs, so its latency hides behind the MMA issue instead of sitting on the critical path
          bool b_ready = mbar_try_wait(&b_full[bst], bphase);
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            if (!b_ready) mbar_wait_t(&b_full[bst], bphase, p.err, 25, st_b);
            {
              const int nst = (bst + 1 == HC_BSTAGES) ? 0 : bst + 1;
              const uint32_t nph = (bst + 1 == HC_BSTAGES) ? (bphase ^ 1) : bphase;
              b_ready = mbar_try_wait(&b_full[nst], nph);
            }
            tc_fence_after();
            const uint32_t b_hi = desc_lo32(smem_u32(smem_b + (size_t)bst * b_stage_bytes), 0);
            const uint32_t b_lo = b_hi + (p.b_bytes >> 4);
#pragma unroll
            for (int t = 0; t < HC_TX; ++t) {
              const uint32_t d_main = tmem_base + (uint32_t)t * 3u * bn + slot_col;
              const uint32_t d_corr = tmem_base + (uint32_t)t * 3u * bn + 2u * bn;
#pragma unroll
              for (int ks = 0; ks < HC_KC / 8; ++ks) {
                const uint32_t ao = tap_off[t9][t] + (uint32_t)ks * 2u;  // +32 bytes per k-step
                const uint32_t first = (t9 == 0 && ks == 0) ? 1u : 0u;
                if (three) {
                  if (!(p.exp & 2)) {  // A_hi is fetched from shared memory once and reused from the collector (-3 %)
                    mma_tf32_lo32_c<kCollFill>(d_main, a_hi + ao, b_hi + ks * 2u, dhi, idesc, first ? fresh_main : 1u);
                    mma_tf32_lo32_c<kCollLastUse>(d_corr, a_hi + ao, b_lo + ks * 2u, dhi, idesc, first ? fresh_corr : 1u);
                  } else {
                    mma_tf32_lo32(d_main, a_hi + ao, b_hi + ks * 2u, dhi, idesc, first ? fresh_main : 1u);
                    mma_tf32_lo32(d_corr, a_hi + ao, b_lo + ks * 2u, dhi, idesc, first ? fresh_corr : 1u);
                  }
                  if (!(p.exp & 4)) mma_tf32_lo32(d_corr, a_lo + ao, b_hi + ks * 2u, dhi, idesc, 1u);
                } else {
                  mma_tf32_lo32(d_main, a_hi + ao, b_hi + ks * 2u, dhi, idesc, first ? fresh_main : 1u);
                }
              }
            }
            mma_commit(&b_empty[bst]);
            if (++bst == HC_BSTAGES) { bst = 0; bphase ^= 1; }
          }
          mma_commit(&a_empty[abuf]);
          if (++abuf == 2) { abuf = 0; aphase ^= 1; }
        }
        mma_commit(&acc_full);
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[0] = st_a; p.dbg[1] = st_b; p.dbg[2] = st_acc; p.dbg[3] = clock64() - t_begin; p.dbg[4] = it; }
    }
  } else if (warp >= 12) {
    // ================================ converters: lo = x - trunc_tf32(x) ================================
    const int tid = threadIdx.x - 12 * 32;  // 0..127
    int abuf = 0;
    uint32_t aphase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      for (int ph = 0; ph < nphases; ++ph) {
        mbar_wait(&a_full[abuf], aphase, p.err, 26);
        if (three && !(p.exp & 1)) {
          const float4 *src = reinterpret_cast<const float4 *>(smem + (size_t)abuf * a_buf_bytes);
          float4 *dst = reinterpret_cast<float4 *>(smem + (size_t)abuf * a_buf_bytes + p.a_bytes);
          const int n16 = (int)(p.a_bytes >> 4);
          // elementwise on the swizzled bytes: hi and lo share the same layout
          for (int i = tid; i < n16; i += 128) {
            const float4 v = src[i];
            float4 l;
            l.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
            l.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
            l.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
            l.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
            dst[i] = l;
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
        }
        mbar_arrive(&a_ready[abuf]);
        if (++abuf == 2) { abuf = 0; aphase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    const int we = (warp - 4) & 3;   // TMEM lane quarter
    const int my_t = (warp - 4) >> 2;  // the output tile (x-plane) this warp drains: the two tiles drain in parallel
    const int m = we * 32 + lane;
    const int lz = m % p.sz, ly = m / p.sz;
    int it = 0;
    long long st_full = 0;
    const long long t_begin = clock64();
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      int x0, y, b;
      if (p.unit_list) {
        const int4 uc = __ldg(p.unit_list + unit);
        x0 = uc.x; y = uc.y + ly; b = uc.z;
      } else {
        int u = unit;
        y = (u % p.tiles_y) * p.ty + ly; u /= p.tiles_y;
        x0 = (u % p.pairs_x) * HC_TX; u /= p.pairs_x;
        b = u;
      }
      mbar_wait_t(&acc_full, (uint32_t)(it & 1), p.err, 27, st_full);
      tc_fence_after();
      {
        const int t = my_t;
        const int x = x0 + t;
        const bool valid = x < p.sx && y < p.sy;
        float *orow = p.out + ((((size_t)b * p.sx + x) * p.sy + y) * p.sz + lz) * p.ldo;
        const uint32_t taddr = tmem_base + ((uint32_t)(we * 32) << 16) + (uint32_t)(t * 3 * p.block_n);
        for (int c0 = 0; c0 < p.block_n; c0 += 32) {
          // all three partial accumulators of a 32-column slab are requested before a single wait
          uint32_t ra[32], rb[32], rc[32];
          float v[32];
          if (c0 + 32 <= p.block_n) {
            tmem_ld32_nowait(taddr + c0, ra);
            tmem_ld32_nowait(taddr + p.block_n + c0, rb);
            if (three) tmem_ld32_nowait(taddr + 2 * p.block_n + c0, rc);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              v[i] = __uint_as_float(ra[i]) + __uint_as_float(rb[i]);
              if (three) v[i] += __uint_as_float(rc[i]);
            }
          } else {  // block_n == 16 or 48: 16-column tail
            float w[16];
            tmem_ld16(taddr + c0, v);
            tmem_ld16(taddr + p.block_n + c0, w);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += w[i];
            if (three) {
              tmem_ld16(taddr + 2 * p.block_n + c0, w);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += w[i];
            }
#pragma unroll
            for (int i = 16; i < 32; ++i) v[i] = 0.f;
          }
          if (valid && c0 < p.cout) {
            if (p.bias) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c0 + i < p.cout) v[i] += __ldg(p.bias + c0 + i);
            }
            if (c0 + 32 <= p.cout) {
#pragma unroll
              for (int i = 0; i < 32; i += 4)
                *reinterpret_cast<float4 *>(orow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            } else {
              for (int i = 0; i < 32 && c0 + i < p.cout; ++i) orow[c0 + i] = v[i];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&acc_empty);
    }
    if (p.dbg && blockIdx.x == 0 && threadIdx.x == 4 * 32) { p.dbg[5] = st_full; p.dbg[6] = clock64() - t_begin; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int encode_map_generic(CUtensorMap *map, const void *ptr, int rank, const unsigned long long *gdim,
                       const unsigned long long *gstride_bytes, const unsigned *box, int swizzle_kind);  // conv_igemm.cu

// shape envelope of this kernel (also used by the pipeline to decide whether `lo` grids are needed at all)
bool conv_halo_v2_supported(int sx, int sy, int sz, int cout) {
  (void)sy;
  if (!(sz % 8 == 0 && sz <= 128 && 128 % sz == 0 && cout <= 64 && sx >= 2)) return false;
  const int ty = 128 / sz;
  const size_t a_bytes = (size_t)sz * (ty + 2) * (HC_TX + 2) * HC_KC * 4;
  const int bn = max(16, ((cout + 15) / 16) * 16);
  const size_t smem = 2 * a_bytes * 2 + (size_t)HC_BSTAGES * bn * HC_KC * 4 * 2 + 1024;
  return a_bytes % 1024 == 0 && smem <= 227 * 1024 - 512;
}

long long *stall_profile_buffer();  // conv_igemm.cu

int *device_error_flag(int slot);  // conv_igemm.cu

// Returns PVCNN_E_UNSUPPORTED when the shape is outside this kernel's envelope (caller falls back to v1).
int conv_halo_v2_launch(int nb, int sx, int sy, int sz, int k, int cout, const float *a, int lda, const float *w_hi,
                     const float *w_lo, int ldw, const float *bias, float *out, int ldo, int npass, cudaStream_t stream,
                     const int4 *unit_list, const int *unit_count) {
  if (!conv_halo_v2_supported(sx, sy, sz, cout)) return PVCNN_E_UNSUPPORTED;
  PVB_CHECK_ARG(a && w_hi && out && (npass == 1 || w_lo) && lda % 4 == 0 && ldw % 4 == 0 && ldo % 4 == 0);
  int *g_halo2_err = device_error_flag(3);
  PVB_CHECK_ARG(g_halo2_err != nullptr);
  HaloParams p{};
  p.nb = nb; p.sx = sx; p.sy = sy; p.sz = sz;
  p.ty = 128 / sz;
  p.tiles_y = ceil_div(sy, p.ty);
  p.pairs_x = ceil_div(sx, HC_TX);
  p.num_units = nb * p.pairs_x * p.tiles_y;
  p.kchunks = ceil_div(k, HC_KC);
  p.cout = cout;
  p.block_n = max(16, ((cout + 15) / 16) * 16);
  p.npass = npass;
  p.ldo = ldo;
  p.a_rows = (uint32_t)(sz * (p.ty + 2) * (HC_TX + 2));
  p.a_bytes = p.a_rows * HC_KC * 4;
  p.b_bytes = (uint32_t)p.block_n * HC_KC * 4;
  p.bias = bias; p.out = out; p.err = g_halo2_err;
  p.unit_list = unit_list; p.unit_count = unit_count;
  p.dbg = stall_profile_buffer();
  { const char *e = getenv("PVCNN_HALO_EXP"); p.exp = e ? atoi(e) : 0; }
  if (p.a_bytes % 1024 != 0) return PVCNN_E_UNSUPPORTED;

  CUtensorMap ma, mw_hi, mw_lo;
  {
    unsigned long long gdim[5] = {(unsigned long long)k, (unsigned long long)sz, (unsigned long long)sy,
                                  (unsigned long long)sx, (unsigned long long)nb};
    unsigned long long gstr[4] = {(unsigned long long)lda * 4, (unsigned long long)sz * lda * 4,
                                  (unsigned long long)sy * sz * lda * 4, (unsigned long long)sx * sy * sz * lda * 4};
    unsigned box[5] = {(unsigned)HC_KC, (unsigned)sz, (unsigned)(p.ty + 2), (unsigned)(HC_TX + 2), 1};
    int rc = encode_map_generic(&ma, a, 5, gdim, gstr, box, 64);
    if (rc) return rc;
  }
  {
    unsigned long long gdim[3] = {(unsigned long long)k, (unsigned long long)cout, 27ull};
    unsigned long long gstr[2] = {(unsigned long long)ldw * 4, (unsigned long long)cout * ldw * 4};
    unsigned box[3] = {(unsigned)HC_KC, (unsigned)p.block_n, 1};
    int rc = encode_map_generic(&mw_hi, w_hi, 3, gdim, gstr, box, 64);
    if (rc) return rc;
    rc = encode_map_generic(&mw_lo, npass > 1 ? w_lo : w_hi, 3, gdim, gstr, box, 64);
    if (rc) return rc;
  }
  const size_t smem = 2 * (size_t)p.a_bytes * 2 + (size_t)HC_BSTAGES * p.b_bytes * 2 + 1024;
  if (smem > 227 * 1024 - 512) return PVCNN_E_UNSUPPORTED;
  PVB_CUDA(cudaFuncSetAttribute(conv_halo_v2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = min(kNumSMs, p.num_units);
  PVB_LAUNCH(conv_halo_v2_kernel, grid, HC_THREADS, smem, stream, ma, mw_hi, mw_lo, p);
  return 0;
}

}  // namespace pvb
