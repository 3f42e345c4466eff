// conv_halo.cu -- third-generation 3x3x3 implicit-GEMM convolution (forward / data gradient) for sm_100a.
// Replaces cuDNN's Conv3d forward / dgrad behind nn.Conv3d (modules/pvconv.py:21,24).
//
// v2 (conv_halo_v2.cu) loads one z-pre-shifted halo per (dz, 16-channel chunk): 3 TMA boxes and 3 `lo` conversions
// per chunk, 12 short phases per unit, TMEM chains that grow with Cin and an epilogue that serialises with the
// next unit's MMAs.  v3 moves the z shift to the OUTPUT side:
//
//   * per 16-channel chunk ONE 5-D TMA box brings the (TX+2) x (TY+2) x BZ halo (no z shift) into 64B-swizzled
//     rows; all 27 taps are served from it.  A (dx,dy) tap of tile t is a row offset of the UMMA descriptor
//     (multiple of 8 rows: swizzle phase preserved); the dz taps accumulate into three separate TMEM
//     accumulators P_dz[row] = sum_{dx,dy,c} W[dx,dy,dz,c] X[x+dx, y+dy, row]  (row = un-shifted z);
//   * the epilogue warps drain the three accumulators after EVERY chunk into fp32 registers and apply the shift
//     there: out[z] += P_-1[z-1] + P_0[z] + P_+1[z+1]  (TMEM lane = z, so the shift is a warp shuffle by one
//     lane; segment edges contribute zero = the conv's z padding).  Consequences:
//       - L2->SMEM halo traffic and `lo` conversions drop 3x; a chunk phase carries 324 MMAs (3xTF32), so the
//         2-deep halo ring is deeply hidden;
//       - TMEM chains are 54 MMAs long for ANY Cin (the tensor core truncates when it accumulates; partial sums
//         are combined in round-to-nearest fp32 registers), so Cin is unbounded;
//       - the final store of unit i overlaps the MMAs of unit i+1 (accumulators are already in registers);
//   * because the z shift is applied in the epilogue, the three dz taps of a (dx,dy) pair read the SAME halo rows: their
//     weight tiles are loaded as ONE box [3 taps x Cout x KC] and issued as ONE tcgen05.mma with N = 3 * Cout (192),
//     whose TMEM columns are exactly [P_-1 | P_0 | P_+1].  The A operand -- 2/3 of a N=64 MMA's shared-memory bytes
//     -- is fetched once per three taps, which moves the kernel from the tensor core's shared-memory operand pipe
//     (r02 ncu: 62 % at N=64, the MMA issue loop saturated) onto the tensor pipe itself;
//   * Cout > 64 runs as N-blocks of 64 (work item = (unit, n-block)); z extents that are not 8/16/32 are padded
//     (R=12 -> BZ=16: the TMA box zero-fills rows z >= sz, the epilogue masks them).
//   * `lo = x - trunc_tf32(x)` of the 3xTF32 split is computed in the kernel by 4 converter warps (once per chunk);
//     in single-pass TF32 mode the freed shared memory becomes a 4-deep halo ring.
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace pvb {
using namespace umma;

constexpr int H3_THREADS = 512;  // w0 TMA(A), w1 MMA, w2 TMEM alloc, w3 TMA(B), w4-11 epilogue (tile = (w-4)/4), w12-15 converters
// channels per chunk: 16 in single-pass TF32 mode (64-byte rows, SWIZZLE_64B), 8 in 3xTF32 mode (32-byte rows,
// SWIZZLE_32B: hi + lo copies of a 3-deep halo ring and a 6-deep ring of 3-tap weight boxes then fit in 227 KB)
constexpr int H3_TX = 2;         // output x-planes (tiles) per work item
constexpr int H3_MAX_A = 4;      // halo ring depth (2 in 3xTF32 mode at R=32)
constexpr int H3_MAX_B = 12;     // weight-box ring depth
constexpr uint32_t kH3LayoutSW64 = 4, kH3LayoutSW32 = 6;

struct Halo3Params {
  int nb, sx, sy, sz;
  int bz;                         // z rows per line in shared memory: 8 / 16 / 32 (>= sz)
  int ty;                         // y rows per tile = 128 / bz
  int tiles_y, pairs_x;
  int num_units;                  // nb * pairs_x * tiles_y (dense walk)
  int kchunks;                    // ceil(k / KC)
  int drain;                      // chunks accumulated in TMEM between two drains (chain <= ~108 MMAs per accumulator)
  int cout, nblocks, block_n;     // N-blocks of block_n (<= 64, multiple of 16) output channels
  int npass, ldo;
  int a_stages, b_stages;
  uint32_t a_bytes, b_bytes;      // bytes of ONE copy (hi) of a halo / a 3-tap weight box
  const float *bias;
  float *out;
  int *err;
  const int4 *unit_list;          // optional compact list of (x0, y0, b, -) units to compute; others are skipped
  const int *unit_count;          //   (device count) -- activity-driven tile skipping, see pvconv_pipeline.cu
  long long *dbg;                 // optional stall counters of CTA 0 (PVCNN_STALL_PROFILE=1)
};

__device__ __forceinline__ void h3_decode(const Halo3Params &p, int unit, int &x0, int &y0, int &b) {
  if (p.unit_list) {
    const int4 uc = __ldg(p.unit_list + unit);
    x0 = uc.x; y0 = uc.y; b = uc.z;
  } else {
    int u = unit;
    y0 = (u % p.tiles_y) * p.ty; u /= p.tiles_y;
    x0 = (u % p.pairs_x) * H3_TX; u /= p.pairs_x;
    b = u;
  }
}

template <bool THREE, int BZ, int KC>
__global__ void __launch_bounds__(H3_THREADS, 1)
    conv_halo_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w_hi,
                     const __grid_constant__ CUtensorMap map_w_lo, const Halo3Params p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t a_full[H3_MAX_A], a_ready[H3_MAX_A], a_empty[H3_MAX_A], b_full[H3_MAX_B], b_empty[H3_MAX_B],
      acc_full, acc_empty;
  __shared__ uint32_t tmem_base_smem;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool three = THREE;   // 3xTF32 (hi/lo split) or single-pass TF32: compile-time, so the MMA issue loop carries no dead path
  const uint32_t a_stage_bytes = p.a_bytes * (three ? 2u : 1u);
  const uint32_t b_stage_bytes = p.b_bytes * (three ? 2u : 1u);
  uint8_t *smem_b = smem + (size_t)p.a_stages * a_stage_bytes;
  const uint32_t tmem_cols = 512;  // 2 tiles x 3 dz x block_n <= 384 columns; one CTA per SM -> take all
  const int num_units = p.unit_list ? __ldg(p.unit_count) : p.num_units;
  const int num_items = num_units * p.nblocks;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a);
    prefetch_tensormap(&map_w_hi);
    if (three) prefetch_tensormap(&map_w_lo);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < p.a_stages; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_ready[i], 128);
      mbar_init(&a_empty[i], 1);
    }
    for (int i = 0; i < p.b_stages; ++i) { mbar_init(&b_full[i], 1); mbar_init(&b_empty[i], 1); }
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
    // ================================ TMA producer: activation halos (one per chunk) ================================
    if (elect_one()) {
      int ast = 0;
      uint32_t aph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        int x0, y0, b;
        h3_decode(p, item / p.nblocks, x0, y0, b);
        for (int cc = 0; cc < p.kchunks; ++cc) {
          mbar_wait(&a_empty[ast], aph ^ 1, p.err, 21);
          mbar_arrive_expect_tx(&a_full[ast], p.a_bytes);
          tma_load_5d(smem + (size_t)ast * a_stage_bytes, &map_a, &a_full[ast], cc * KC, 0, y0 - 1, x0 - 1, b);
          if (++ast == p.a_stages) { ast = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ================================ TMA producer: weight boxes (9 per chunk, 3 dz taps each) ================================
    if (elect_one()) {
      int bst = 0;
      uint32_t bph = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        const int n0 = (item % p.nblocks) * p.block_n;
        for (int cc = 0; cc < p.kchunks; ++cc) {
#pragma unroll 1
          for (int t9 = 0; t9 < 9; ++t9) {  // t9 = (dx+1)*3 + (dy+1); torch tap index = t9*3 + (dz+1): the 3 dz taps are adjacent
            mbar_wait(&b_empty[bst], bph ^ 1, p.err, 22);
            uint8_t *sb = smem_b + (size_t)bst * b_stage_bytes;
            mbar_arrive_expect_tx(&b_full[bst], b_stage_bytes);
            tma_load_3d(sb, &map_w_hi, &b_full[bst], cc * KC, n0, t9 * 3);
            if (three) tma_load_3d(sb + p.b_bytes, &map_w_lo, &b_full[bst], cc * KC, n0, t9 * 3);
            if (++bst == p.b_stages) { bst = 0; bph ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    // (Tried in r02: letting the whole warp walk this loop so that the descriptor arithmetic stays on the uniform datapath
    //  and only the tcgen05.mma / commit are predicated on the elected lane -- it removes the R2UR moves but measured
    //  slower in 3xTF32 mode (0.515 vs 0.490 ms dense) and equal in TF32 mode, so the single elected thread stays.)
    if (elect_one()) {
      const uint32_t bn = (uint32_t)p.block_n;
      const uint32_t idesc = make_idesc_tf32(128, 3 * p.block_n, 0, 0);   // N = [P_-1 | P_0 | P_+1]
      constexpr uint32_t dhi = KC == 16 ? desc_hi32(512, kH3LayoutSW64) : desc_hi32(256, kH3LayoutSW32);
      // One elected thread issues every MMA of the CTA; every instruction next to a tcgen05.mma counts.  The halo geometry
      // is a template parameter so that every tap's descriptor offset is an immediate:
      //   rows are (x_local, y_local, z); 16-byte units: one row = KC/4, a y step = BZ rows, an x step = (TY + 2) * BZ rows
      constexpr uint32_t YS = (uint32_t)BZ * (KC * 4 / 16);
      constexpr uint32_t XS = (uint32_t)(128 / BZ + 2) * YS;
      uint64_t *a_bar = three ? a_ready : a_full;  // single pass: nothing to convert, consume the TMA data directly
      int ast = 0, bst = 0;
      uint32_t aph = 0, bph = 0, gd = 0;  // gd: drain periods issued so far (accumulator hand-shake parity)
      long long st_acc = 0, st_a = 0, st_b = 0;
      int items_done = 0;
      const uint32_t b_base = desc_lo32(smem_u32(smem_b), 0);
      const uint32_t b_step = b_stage_bytes >> 4, b_lo_off = p.b_bytes >> 4, a_lo_off = p.a_bytes >> 4;
      const uint32_t d0 = tmem_base, d1 = tmem_base + 3u * bn;   // tile 0 / tile 1: 3*bn columns each
      const long long t_begin = clock64();
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++items_done) {
        int in_period = 0;
        for (int cc = 0; cc < p.kchunks; ++cc) {
          if (in_period == 0) {
            // the epilogue has copied the previous drain period's accumulators (both tiles) into registers
            mbar_wait_t(&acc_empty, (gd & 1u) ^ 1u, p.err, 23, st_acc);
            tc_fence_after();
          }
          mbar_wait_t(&a_bar[ast], aph, p.err, 24, st_a);
          tc_fence_after();
          const uint32_t a_hi = desc_lo32(smem_u32(smem + (size_t)ast * a_stage_bytes), 0);
          const uint32_t a_lo = a_hi + a_lo_off;
          // software-pipelined barrier polling: the try_wait for the NEXT weight box is issued before this box's MMAs,
          // so its latency hides behind the MMA issue instead of sitting on the critical path
          bool b_ready = mbar_try_wait(&b_full[bst], bph);
#pragma unroll
          for (int t9 = 0; t9 < 9; ++t9) {
            if (!b_ready) mbar_wait_t(&b_full[bst], bph, p.err, 25, st_b);
            const uint32_t b_hi = b_base + (uint32_t)bst * b_step;
            const uint64_t *done_bar = &b_empty[bst];
            if (++bst == p.b_stages) { bst = 0; bph ^= 1; }
            b_ready = mbar_try_wait(&b_full[bst], bph);
            tc_fence_after();
#pragma unroll
            for (int ks = 0; ks < KC / 8; ++ks) {
#pragma unroll
              for (int t = 0; t < H3_TX; ++t) {  // the two tiles' chains alternate (independent accumulators)
                const uint32_t ao = (uint32_t)(t + t9 / 3) * XS + (uint32_t)(t9 % 3) * YS + (uint32_t)ks * 2u;  // immediate
                const uint32_t d = t == 0 ? d0 : d1;
                const uint32_t acc = (in_period == 0 && t9 == 0 && ks == 0) ? 0u : 1u;  // a drain period starts a fresh chain
                if (three) {
                  // A_hi is fetched from shared memory once and reused from the collector for the B_lo product
                  mma_tf32_lo32_c<kCollFill>(d, a_hi + ao, b_hi + ks * 2u, dhi, idesc, acc);
                  mma_tf32_lo32_c<kCollLastUse>(d, a_hi + ao, b_hi + b_lo_off + ks * 2u, dhi, idesc, 1u);
                  mma_tf32_lo32(d, a_lo + ao, b_hi + ks * 2u, dhi, idesc, 1u);
                } else {
                  mma_tf32_lo32(d, a_hi + ao, b_hi + ks * 2u, dhi, idesc, acc);
                }
              }
            }
            mma_commit(const_cast<uint64_t *>(done_bar));
          }
          mma_commit(&a_empty[ast]);
          if (++ast == p.a_stages) { ast = 0; aph ^= 1; }
          if (++in_period == p.drain || cc == p.kchunks - 1) {
            mma_commit(&acc_full);
            in_period = 0;
            ++gd;
          }
        }
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[0] = st_a; p.dbg[1] = st_b; p.dbg[2] = st_acc; p.dbg[3] = clock64() - t_begin; p.dbg[4] = items_done; }
    }
  } else if (warp >= 12) {
    // ================================ converters: lo = x - trunc_tf32(x) ================================
    if (three) {
      const int tid = threadIdx.x - 12 * 32;  // 0..127
      int ast = 0;
      uint32_t aph = 0;
      const int n16 = (int)(p.a_bytes >> 4);
      for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
        for (int cc = 0; cc < p.kchunks; ++cc) {
          mbar_wait(&a_full[ast], aph, p.err, 26);
          const float4 *src = reinterpret_cast<const float4 *>(smem + (size_t)ast * a_stage_bytes);
          float4 *dst = reinterpret_cast<float4 *>(smem + (size_t)ast * a_stage_bytes + p.a_bytes);
          // elementwise on the swizzled bytes: hi and lo share the same layout
#pragma unroll 4
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
          mbar_arrive(&a_ready[ast]);
          if (++ast == p.a_stages) { ast = 0; aph ^= 1; }
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue: drain per chunk, z shift, bias, store ================================
    const int q = (warp - 4) & 3;      // TMEM lane quarter
    const int my_t = (warp - 4) >> 2;  // the output tile (x-plane) this warp drains: the two tiles drain in parallel
    const int m = q * 32 + lane;
    const int lz = m % BZ, ly = m / BZ;
    const bool z_first = lz == 0, z_last = lz == BZ - 1;
    const int bn = p.block_n;
    uint32_t gd = 0;
    long long st_full = 0;
    const long long t_begin = clock64();
    for (int item = blockIdx.x; item < num_items; item += gridDim.x) {
      float acc[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
      for (int c0p = 0; c0p < p.kchunks; c0p += p.drain, ++gd) {   // one iteration per drain period
        mbar_wait_t(&acc_full, gd & 1u, p.err, 27, st_full);
        tc_fence_after();
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
          const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)((my_t * 3 + dz) * bn);
#pragma unroll
          for (int c0 = 0; c0 < 64; c0 += 32) {
            if (c0 < bn) {  // warp-uniform
              uint32_t r[32];
              if (c0 + 32 <= bn) {
                tmem_ld32_nowait(taddr + c0, r);
                tmem_ld_wait();
              } else {  // block_n == 16 or 48: 16-column tail
                float w[16];
                tmem_ld16(taddr + c0, w);
#pragma unroll
                for (int i = 0; i < 16; ++i) { r[i] = __float_as_uint(w[i]); r[16 + i] = 0u; }
              }
              if (dz == 2 && c0 + 32 >= bn) {  // the last slab of this tile is in registers: hand TMEM back to the MMA warp
                tc_fence_before();
                mbar_arrive(&acc_empty);
              }
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                float v = __uint_as_float(r[i]);
                if (dz == 0) {         // P_-1: out[z] takes row z-1
                  v = __shfl_up_sync(0xffffffffu, v, 1);
                  if (z_first) v = 0.0f;
                } else if (dz == 2) {  // P_+1: out[z] takes row z+1
                  v = __shfl_down_sync(0xffffffffu, v, 1);
                  if (z_last) v = 0.0f;
                }
                acc[c0 + i] = __fadd_rn(acc[c0 + i], v);
              }
            }
          }
        }
      }
      // ---- bias + store (the MMA warp is already working on the next item)
      int x0, y0, b;
      const int unit = item / p.nblocks;
      h3_decode(p, unit, x0, y0, b);
      const int n0 = (item - unit * p.nblocks) * bn;
      const int x = x0 + my_t, y = y0 + ly;
      const int ncols = min(bn, p.cout - n0);
      if (x < p.sx && y < p.sy && lz < p.sz) {
        float *orow = p.out + ((((size_t)b * p.sx + x) * p.sy + y) * p.sz + lz) * p.ldo + n0;
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          if (i < ncols) {
            float4 v = make_float4(acc[i], acc[i + 1], acc[i + 2], acc[i + 3]);
            if (p.bias) {
              v.x += __ldg(p.bias + n0 + i);
              if (i + 1 < ncols) v.y += __ldg(p.bias + n0 + i + 1);
              if (i + 2 < ncols) v.z += __ldg(p.bias + n0 + i + 2);
              if (i + 3 < ncols) v.w += __ldg(p.bias + n0 + i + 3);
            }
            if (i + 4 <= ncols) {
              *reinterpret_cast<float4 *>(orow + i) = v;
            } else {
              orow[i] = v.x;
              if (i + 1 < ncols) orow[i + 1] = v.y;
              if (i + 2 < ncols) orow[i + 2] = v.z;
            }
          }
        }
      }
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
long long *stall_profile_buffer();                                                                       // conv_igemm.cu
int *device_error_flag(int slot);                                                                        // conv_igemm.cu

int conv_halo_v2_launch(int nb, int sx, int sy, int sz, int k, int cout, const float *a, int lda, const float *w_hi,
                        const float *w_lo, int ldw, const float *bias, float *out, int ldo, int npass, cudaStream_t stream,
                        const int4 *unit_list, const int *unit_count);  // conv_halo_v2.cu
bool conv_halo_v2_supported(int sx, int sy, int sz, int cout);

static bool halo_use_v2() {
  const char *e = getenv("PVCNN_B200_CONV");
  return e && e[0] == 'v' && e[1] == '2';
}

// z rows per shared-memory line (the tile is 128 = bz * ty rows); 0 = outside the envelope
int conv_halo_bz(int sz) { return sz <= 8 ? 8 : (sz <= 16 ? 16 : (sz <= 32 ? 32 : 0)); }
// y rows per output tile of the kernel that will run this shape (the activity lists are built in these units)
int conv_halo_ty(int sz) {
  if (halo_use_v2()) return 128 / sz > 0 ? 128 / sz : 1;
  const int bz = conv_halo_bz(sz);
  return bz ? 128 / bz : 1;
}

// shape envelope of this kernel (the pipeline enables activity skipping when every 3x3x3 conv of a block is inside it)
bool conv_halo_supported(int sx, int sy, int sz, int cout) {
  if (halo_use_v2()) return conv_halo_v2_supported(sx, sy, sz, cout);
  (void)sy; (void)cout;
  return conv_halo_bz(sz) != 0 && sx >= 2;
}

// Returns PVCNN_E_UNSUPPORTED when the shape is outside this kernel's envelope (caller falls back to v1).
int conv_halo_launch(int nb, int sx, int sy, int sz, int k, int cout, const float *a, int lda, const float *w_hi,
                     const float *w_lo, int ldw, const float *bias, float *out, int ldo, int npass, cudaStream_t stream,
                     const int4 *unit_list, const int *unit_count) {
  if (halo_use_v2())
    return conv_halo_v2_launch(nb, sx, sy, sz, k, cout, a, lda, w_hi, w_lo, ldw, bias, out, ldo, npass, stream, unit_list,
                               unit_count);
  if (!conv_halo_supported(sx, sy, sz, cout)) return PVCNN_E_UNSUPPORTED;
  PVB_CHECK_ARG(a && w_hi && out && (npass == 1 || w_lo) && lda % 4 == 0 && ldw % 4 == 0 && ldo % 4 == 0);
  Halo3Params p{};
  p.nb = nb; p.sx = sx; p.sy = sy; p.sz = sz;
  p.bz = conv_halo_bz(sz);
  p.ty = 128 / p.bz;
  p.tiles_y = ceil_div(sy, p.ty);
  p.pairs_x = ceil_div(sx, H3_TX);
  p.num_units = nb * p.pairs_x * p.tiles_y;
  const int kc = npass > 1 ? 8 : 16;
  p.kchunks = ceil_div(k, kc);
  // TMEM chains stay <= 108 MMAs per accumulator (the tensor core truncates when it accumulates): 9 (dx,dy) x (KC/8)
  // k-steps x (3 products | 1) MMAs per chunk
  p.drain = max(1, 108 / (9 * (kc / 8) * (npass > 1 ? 3 : 1)));
  p.cout = cout;
  p.block_n = cout >= 64 ? 64 : max(16, ((cout + 15) / 16) * 16);
  p.nblocks = ceil_div(cout, p.block_n);
  p.npass = npass;
  p.ldo = ldo;
  p.a_bytes = (uint32_t)(p.bz * (p.ty + 2) * (H3_TX + 2)) * kc * 4;
  p.b_bytes = 3u * (uint32_t)p.block_n * kc * 4;   // one box = the 3 dz taps of a (dx,dy) pair
  if (p.a_bytes % 512 != 0 || p.b_bytes % 512 != 0) return PVCNN_E_UNSUPPORTED;
  const uint32_t a_stage = p.a_bytes * (npass > 1 ? 2 : 1), b_stage = p.b_bytes * (npass > 1 ? 2 : 1);
  const int budget = 227 * 1024 - 1024 - 1024;  // alignment slack + static shared memory (barriers)
  // a chunk's MMAs (>= 3 k cycles) hide a halo load with 2 stages; the weight boxes are the latency-critical stream (a box
  // is consumed in ~400 cycles, a TMA round trip takes ~2 k: measured 34 % of the issue time stalled on them with a
  // 6-deep ring), so everything else goes to the weight ring
  p.a_stages = 2;
  { const char *e = getenv("PVCNN_HALO_ASTAGES"); if (e && e[0] >= '2' && e[0] <= '4') p.a_stages = e[0] - '0'; }
  if ((budget - 4 * (int)b_stage) / (int)a_stage < p.a_stages) p.a_stages = (budget - 4 * (int)b_stage) / (int)a_stage;
  if (p.a_stages < 2) return PVCNN_E_UNSUPPORTED;
  p.b_stages = min(H3_MAX_B, (budget - p.a_stages * (int)a_stage) / (int)b_stage);
  p.bias = bias; p.out = out;
  p.err = device_error_flag(1);
  PVB_CHECK_ARG(p.err != nullptr);
  p.unit_list = unit_list; p.unit_count = unit_count;
  p.dbg = stall_profile_buffer();

  CUtensorMap ma, mw_hi, mw_lo;
  {
    unsigned long long gdim[5] = {(unsigned long long)k, (unsigned long long)sz, (unsigned long long)sy,
                                  (unsigned long long)sx, (unsigned long long)nb};
    unsigned long long gstr[4] = {(unsigned long long)lda * 4, (unsigned long long)sz * lda * 4,
                                  (unsigned long long)sy * sz * lda * 4, (unsigned long long)sx * sy * sz * lda * 4};
    unsigned box[5] = {(unsigned)kc, (unsigned)p.bz, (unsigned)(p.ty + 2), (unsigned)(H3_TX + 2), 1};
    int rc = encode_map_generic(&ma, a, 5, gdim, gstr, box, kc == 16 ? 64 : 320);
    if (rc) return rc;
  }
  {
    unsigned long long gdim[3] = {(unsigned long long)k, (unsigned long long)cout, 27ull};
    unsigned long long gstr[2] = {(unsigned long long)ldw * 4, (unsigned long long)cout * ldw * 4};
    unsigned box[3] = {(unsigned)kc, (unsigned)p.block_n, 3};
    int rc = encode_map_generic(&mw_hi, w_hi, 3, gdim, gstr, box, kc == 16 ? 64 : 320);
    if (rc) return rc;
    rc = encode_map_generic(&mw_lo, npass > 1 ? w_lo : w_hi, 3, gdim, gstr, box, kc == 16 ? 64 : 320);
    if (rc) return rc;
  }
  const size_t smem = (size_t)p.a_stages * a_stage + (size_t)p.b_stages * b_stage + 1024;
  const int grid = min(kNumSMs, p.num_units * p.nblocks);
#define PVB_HALO_LAUNCH(T3, BZV, KCV)                                                                                          \
  do {                                                                                                                         \
    PVB_CUDA(cudaFuncSetAttribute(conv_halo_kernel<T3, BZV, KCV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    PVB_LAUNCH((conv_halo_kernel<T3, BZV, KCV>), grid, H3_THREADS, smem, stream, ma, mw_hi, mw_lo, p);                        \
  } while (0)
  if (npass > 1) {
    if (p.bz == 32) PVB_HALO_LAUNCH(true, 32, 8); else if (p.bz == 16) PVB_HALO_LAUNCH(true, 16, 8); else PVB_HALO_LAUNCH(true, 8, 8);
  } else {
    if (p.bz == 32) PVB_HALO_LAUNCH(false, 32, 16); else if (p.bz == 16) PVB_HALO_LAUNCH(false, 16, 16); else PVB_HALO_LAUNCH(false, 8, 16);
  }
#undef PVB_HALO_LAUNCH
  return 0;
}

}  // namespace pvb
