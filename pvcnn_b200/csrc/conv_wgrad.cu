// conv_wgrad.cu -- weight gradient of the 3x3x3 / 1x1 convolutions on tcgen05 tensor cores (sm_100a).
//
//   dW[co][ci][tap] = sum_v  gY[v][co] * X[v + off(tap)][ci]            (reduction over ALL voxels / points)
//
// Replaces cuDNN's wgrad for nn.Conv3d (modules/pvconv.py:21,24) and nn.Conv1d (modules/shared_mlp.py:10).
//
//   * Both operands are read straight from the channels-last activations, i.e. with the reduction
//     index (voxel) as the ROW of the shared-memory tile and 32 channels per 128-byte row: that is the
//     MN-major operand layout of tcgen05.mma (kind::tf32 supports it), so no transposed copies exist.
//   * A = X tiles, shifted per tap by the TMA box coordinates (OOB zero fill = padding).  Four
//     (tap, 32-channel) blocks form one M=128 operand; B = the gY tile (N = cout).
//   * The tensor core truncates when it accumulates (see conv_igemm.cu), and this reduction is 524 288
//     long at the metric shape.  So TMEM only holds SHORT chains (<= 64 MMAs): the epilogue
//     warps drain each finished chain into fp32 REGISTER accumulators (round-to-nearest adds) while the
//     MMA warp fills the other TMEM buffer.  The 3xTF32 correction terms use their own accumulators.
//   * split-K over the SMs; partial results are combined with fp32 reductions into the zeroed dW.
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace pvb {
using namespace umma;

constexpr int WG_THREADS = 512;       // warp0 TMA, warp1 MMA, warp2 TMEM alloc, warps 4-11 drain, warps 12-15 lo converters
constexpr int WG_ROWS = 32;           // voxels per k-tile (box rows), 4 MMA K-steps
constexpr uint32_t WG_BLK = WG_ROWS * 128;  // bytes of one [32 rows x 32 channels] block
constexpr int WG_MAX_STAGES = 4;      // TMA ring depth (bytes in flight hide the ~3000-cycle load latency)
constexpr int WG_DRAIN_TILES = 16;    // k-tiles per TMEM chain: 16 * 4 = 64 MMAs

struct WgradParams {
  int nb, sx, sy, sz;
  int bz, by;               // k-tile box (bz * by = rows <= 32, multiple of 8)
  int tz, ty;               // k-tiles per dim (x is walked one slice at a time)
  long long num_ktiles;
  int cin, cout, ntaps;
  int chunks_in, chunks_out;
  int n_ablocks;            // ntaps * chunks_in
  int groups_per_cta;       // G: M=128 operand groups handled by one CTA
  int num_sets;             // ceil(ceil(n_ablocks/4) / G)
  int ksplit;               // CTAs per set
  int block_n;              // cout padded to 16 (<= 128)
  int npass;
  int ksteps;               // rows / 8
  uint32_t stage_bytes, g_bytes;  // g_bytes: bytes of the gY part of a stage (hi [+lo])
  uint32_t box_bytes;             // bytes one TMA box really delivers (rows * 128)
  int stages;
  int lo_in_kernel;               // 1: converter warps derive lo from hi in shared memory; 0: lo tensors are TMA-loaded
  float *dw;                // [cout][cin][ntaps], zero-initialised by the launcher
  int *err;
  const int4 *ktile_list;   // optional compact list of (z0, y0, x0, b) k-tiles that can contribute (activity skipping)
  const int *ktile_count;
  long long *dbg;           // optional [8] stall-cycle counters of CTA 0 (PVCNN_STALL_PROFILE)
};

template <int G, bool THREE>
__global__ void __launch_bounds__(WG_THREADS, 1)
    conv_wgrad_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                      const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                      const WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[WG_MAX_STAGES], ready_bar[WG_MAX_STAGES], empty_bar[WG_MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  const int set = blockIdx.x / p.ksplit, split = blockIdx.x % p.ksplit;
  const int ab0 = set * G * 4;                                  // first A-block of this CTA
  const int nab = min(G * 4, p.n_ablocks - ab0);                // A-blocks that really exist
  const int ngroups = (nab + 3) / 4;
  const uint32_t passes = p.npass > 1 ? 2u : 1u;                // hi [+ lo] copies of each operand
  const uint32_t tmem_cols_per_buf = (uint32_t)(G * p.block_n * (p.npass > 1 ? 2 : 1));
  uint32_t tmem_cols = 32;
  while (tmem_cols < 2 * tmem_cols_per_buf) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x_hi);
    prefetch_tensormap(&map_g_hi);
    if (p.npass > 1) { prefetch_tensormap(&map_x_lo); prefetch_tensormap(&map_g_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 256); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  // contiguous range of k-tiles walked by this CTA (coordinates advance incrementally: no div/mod per stage)
  const long long num_ktiles = p.ktile_list ? (long long)__ldg(p.ktile_count) : p.num_ktiles;
  const long long per_cta = (num_ktiles + p.ksplit - 1) / p.ksplit;
  const long long kt_begin = (long long)split * per_cta;
  const long long my_tiles = max(0LL, min(num_ktiles, kt_begin + per_cta) - kt_begin);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      const uint32_t tx_bytes = p.box_bytes * (uint32_t)(p.chunks_out + nab) * ((p.npass > 1 && !p.lo_in_kernel) ? 2u : 1u);
      const bool load_lo = p.npass > 1 && !p.lo_in_kernel;
      long long stall = 0;
      const long long t_begin = clock64();
      int stage = 0;
      uint32_t phase = 0;
      // this CTA's operand blocks never change: decode (tap, chunk) -> box offsets once
      int blk_c[G * 4], blk_dx[G * 4], blk_dy[G * 4], blk_dz[G * 4];
#pragma unroll
      for (int a = 0; a < G * 4; ++a) {
        const int ab = min(ab0 + a, p.n_ablocks - 1);
        const int tap = ab / p.chunks_in;
        blk_c[a] = (ab - tap * p.chunks_in) * 32;
        blk_dx[a] = p.ntaps == 27 ? tap / 9 - 1 : 0;
        blk_dy[a] = p.ntaps == 27 ? (tap / 3) % 3 - 1 : 0;
        blk_dz[a] = p.ntaps == 27 ? tap % 3 - 1 : 0;
      }
      int kt = (int)kt_begin;
      int tzi = kt % p.tz; kt /= p.tz;
      int tyi = kt % p.ty; kt /= p.ty;
      int x0 = kt % p.sx; kt /= p.sx;
      int b = kt;
      for (long long t = 0; t < my_tiles; ++t) {
        int z0 = tzi * p.bz, y0 = tyi * p.by;
        if (p.ktile_list) {
          const int4 kc = __ldg(p.ktile_list + kt_begin + t);
          z0 = kc.x; y0 = kc.y; x0 = kc.z; b = kc.w;
        }
        mbar_wait_t(&empty_bar[stage], phase ^ 1, p.err, 11, stall);
        uint8_t *st = smem + (size_t)stage * p.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        for (int cc = 0; cc < p.chunks_out; ++cc) {
          tma_load_5d(st + (size_t)cc * WG_BLK, &map_g_hi, &full_bar[stage], cc * 32, z0, y0, x0, b);
          if (load_lo)
            tma_load_5d(st + (size_t)(p.chunks_out + cc) * WG_BLK, &map_g_lo, &full_bar[stage], cc * 32, z0, y0, x0, b);
        }
        uint8_t *sa = st + p.g_bytes;
#pragma unroll
        for (int a = 0; a < G * 4; ++a)
          if (a < nab) {
            tma_load_5d(sa + (size_t)a * WG_BLK, &map_x_hi, &full_bar[stage], blk_c[a], z0 + blk_dz[a], y0 + blk_dy[a],
                        x0 + blk_dx[a], b);
            if (load_lo)
              tma_load_5d(sa + (size_t)(G * 4 + a) * WG_BLK, &map_x_lo, &full_bar[stage], blk_c[a], z0 + blk_dz[a],
                          y0 + blk_dy[a], x0 + blk_dx[a], b);
          }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++tzi == p.tz) { tzi = 0; if (++tyi == p.ty) { tyi = 0; if (++x0 == p.sx) { x0 = 0; ++b; } } }
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[0] = stall; p.dbg[1] = clock64() - t_begin; }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, p.block_n, /*a MN-major*/ 1, /*b MN-major*/ 1);
      int stage = 0;
      uint32_t phase = 0;
      int chain = 0;  // index of the current TMEM chain
      long long stall_t = 0, stall_r = 0;
      const long long t_begin = clock64();
      for (long long t = 0; t < my_tiles; ++t) {
        const int pos_in_chain = (int)(t % WG_DRAIN_TILES);
        const int buf = chain & 1;
        if (pos_in_chain == 0) {
          mbar_wait_t(&tmem_empty_bar[buf], ((chain >> 1) & 1) ^ 1, p.err, 12, stall_t);
          tc_fence_after();
        }
        mbar_wait_t(&ready_bar[stage], phase, p.err, 13, stall_r);
        tc_fence_after();
        // MN-major tf32 operands: 128B swizzle with 32-byte atoms (4-row repeat); LBO = stride between
        // 32-channel blocks, SBO = stride between 4-row K groups; +1024 bytes (64 units) per 8-voxel k-step
        constexpr uint32_t dhi = desc_hi32(512, kLayoutSW128Base32);
        const uint32_t st = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t g_hi = desc_lo32(st, WG_BLK);
        const uint32_t g_lo = g_hi + (uint32_t)p.chunks_out * (WG_BLK >> 4);
        const uint32_t a_base = g_hi + (p.g_bytes >> 4);
        const uint32_t first = pos_in_chain == 0 ? 0u : 1u;
        // The single issuing thread is issue-bound (see conv_halo.cu): precision mode and group count are template
        // parameters, the usual 4 k-steps are unrolled, and A_hi (the X block) is fetched from shared memory once for its
        // two products (collector fill / lastuse).
#pragma unroll
        for (int g = 0; g < G; ++g) {
          if (g < ngroups) {
            const uint32_t a_hi = a_base + (uint32_t)(g * 4) * (WG_BLK >> 4);
            const uint32_t a_lo = a_base + (uint32_t)(G * 4 + g * 4) * (WG_BLK >> 4);
            const uint32_t d_main = tmem_base + (uint32_t)buf * tmem_cols_per_buf + (uint32_t)(g * p.block_n);
            const uint32_t d_corr = d_main + (uint32_t)(G * p.block_n);
            auto kstep = [&](int ks) {
              const uint32_t ko = (uint32_t)ks * 64u;
              const uint32_t accum = ks == 0 ? first : 1u;
              if (THREE) {
                mma_tf32_lo32_c<kCollFill>(d_main, a_hi + ko, g_hi + ko, dhi, idesc, accum);
                mma_tf32_lo32_c<kCollLastUse>(d_corr, a_hi + ko, g_lo + ko, dhi, idesc, accum);
                mma_tf32_lo32(d_corr, a_lo + ko, g_hi + ko, dhi, idesc, 1u);
              } else {
                mma_tf32_lo32(d_main, a_hi + ko, g_hi + ko, dhi, idesc, accum);
              }
            };
            if (p.ksteps == 4) {
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) kstep(ks);
            } else {
              for (int ks = 0; ks < p.ksteps; ++ks) kstep(ks);
            }
          }
        }
        mma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (pos_in_chain == WG_DRAIN_TILES - 1 || t == my_tiles - 1) {
          mma_commit(&tmem_full_bar[buf]);
          ++chain;
        }
      }
      if (p.dbg && blockIdx.x == 0) { p.dbg[2] = stall_r; p.dbg[3] = stall_t; p.dbg[4] = clock64() - t_begin; }
    }
  } else if (warp >= 12) {
    // ================================ converters: lo = x - trunc_tf32(x) ================================
    // (activations and gradients carry no `lo` tensors in HBM; elementwise on the swizzled bytes)
    const int tid = threadIdx.x - 12 * 32;
    int stage = 0;
    uint32_t phase = 0;
    long long stall_c = 0, fence_cycles = 0;
    const long long t_begin = clock64();
    for (long long t = 0; t < my_tiles; ++t) {
      mbar_wait_t(&full_bar[stage], phase, p.err, 15, stall_c);
      if (p.npass > 1 && p.lo_in_kernel) {
        uint8_t *st = smem + (size_t)stage * p.stage_bytes;
        const int ng16 = p.chunks_out * (int)(WG_BLK >> 4), na16 = nab * (int)(WG_BLK >> 4);
        const float4 *g_src = reinterpret_cast<const float4 *>(st);
        float4 *g_dst = reinterpret_cast<float4 *>(st + (size_t)p.chunks_out * WG_BLK);
        const float4 *a_src = reinterpret_cast<const float4 *>(st + p.g_bytes);
        float4 *a_dst = reinterpret_cast<float4 *>(st + p.g_bytes + (size_t)(G * 4) * WG_BLK);
        const int total16 = ng16 + na16;
        for (int i0 = tid; i0 < total16; i0 += 4 * 128) {
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 128;
            if (i < total16) v[u] = (i < ng16) ? g_src[i] : a_src[i - ng16];
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 128;
            if (i < total16) {
              float4 l;
              l.x = __fsub_rn(v[u].x, __uint_as_float(__float_as_uint(v[u].x) & 0xFFFFE000u));
              l.y = __fsub_rn(v[u].y, __uint_as_float(__float_as_uint(v[u].y) & 0xFFFFE000u));
              l.z = __fsub_rn(v[u].z, __uint_as_float(__float_as_uint(v[u].z) & 0xFFFFE000u));
              l.w = __fsub_rn(v[u].w, __uint_as_float(__float_as_uint(v[u].w) & 0xFFFFE000u));
              if (i < ng16) g_dst[i] = l; else a_dst[i - ng16] = l;
            }
          }
        }
        const long long tf0 = clock64();
        fence_proxy_async();
        fence_cycles += clock64() - tf0;
      }
      mbar_arrive(&ready_bar[stage]);
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
    if (p.dbg && blockIdx.x == 0 && tid == 0) { p.dbg[5] = stall_c; p.dbg[6] = clock64() - t_begin; p.dbg[7] = fence_cycles; }
  } else if (warp >= 4) {
    // ================================ drain / epilogue ================================
    const int e = warp - 4;
    const int q = e & 3;   // TMEM lane quarter
    const int h = e >> 2;  // column half
    const int cols = p.block_n / 2;            // columns owned per accumulator
    constexpr int CMAX = 64 / G;               // register columns per group (G=2: 32, G=1: 64)
    float acc[G][CMAX];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < CMAX; ++i) acc[g][i] = 0.0f;
    const long long nchains = (my_tiles + WG_DRAIN_TILES - 1) / WG_DRAIN_TILES;
    for (long long ch = 0; ch < nchains; ++ch) {
      const int buf = (int)(ch & 1);
      mbar_wait(&tmem_full_bar[buf], (uint32_t)((ch >> 1) & 1), p.err, 14);
      tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * tmem_cols_per_buf;
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int c0 = 0; c0 < CMAX; c0 += 16) {
          if (g < ngroups && c0 < cols) {  // warp-uniform
            float v[16];
            tmem_ld16(tb + (uint32_t)(g * p.block_n + h * cols + c0), v);
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[g][c0 + i] += v[i];
            if (p.npass > 1) {
              tmem_ld16(tb + (uint32_t)(G * p.block_n + g * p.block_n + h * cols + c0), v);
#pragma unroll
              for (int i = 0; i < 16; ++i) acc[g][c0 + i] += v[i];
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[buf]);
    }
    // scatter-add this CTA's partial dW:  D row m = (A-block m/32, channel m%32), column = cout index
    const int m = q * 32 + lane;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const int a = g * 4 + (m >> 5);
      if (g < ngroups && a < nab) {
        const int ab = ab0 + a;
        const int tap = ab / p.chunks_in, cc = ab - tap * p.chunks_in;
        const int ci = cc * 32 + (m & 31);
        if (ci < p.cin) {
#pragma unroll
          for (int c = 0; c < CMAX; ++c) {
            const int co = h * cols + c;
            if (c < cols && co < p.cout) atomicAdd(p.dw + ((size_t)co * p.cin + ci) * p.ntaps + tap, acc[g][c]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

int encode_map_5d_cl(CUtensorMap *map, const float *ptr, int k, int ld, int nb, int sx, int sy, int sz, int bz, int by,
                     int bx, bool atom32);  // conv_igemm.cu

long long *stall_profile_buffer();  // conv_igemm.cu

int *device_error_flag(int slot);  // conv_igemm.cu

// x: layer input [nb,sx,sy,sz,ldx] (cin valid), g: output gradient [nb,sx,sy,sz,ldg] (cout <= 128 valid)
static int wgrad_launch_block(int nb, int sx, int sy, int sz, int cin, int cout, int ntaps, const float *x_hi, const float *x_lo,
                 int ldx, const float *g_hi, const float *g_lo, int ldg, float *dw, int npass, cudaStream_t s,
                 const int4 *ktile_list, const int *ktile_count, int *bz_out, int *by_out) {
  PVB_CHECK_ARG(nb > 0 && sx > 0 && sy > 0 && sz > 0 && cin > 0 && cout > 0 && (ntaps == 1 || ntaps == 27));
  PVB_CHECK_ARG(x_hi && g_hi && dw && ldx % 4 == 0 && ldg % 4 == 0);
  PVB_CHECK_ARG(cout <= 128);
  int *g_wg_err = device_error_flag(2);
  PVB_CHECK_ARG(g_wg_err != nullptr);
  WgradParams p{};
  p.nb = nb; p.sx = sx; p.sy = sy; p.sz = sz;
  p.bz = ((min(sz, WG_ROWS) + 7) / 8) * 8;
  p.by = max(1, WG_ROWS / p.bz);
  if (p.by > sy) p.by = sy;
  const int rows = p.bz * p.by;
  p.ksteps = rows / 8;
  p.tz = ceil_div(sz, p.bz);
  p.ty = ceil_div(sy, p.by);
  p.num_ktiles = (long long)nb * sx * p.ty * p.tz;
  p.cin = cin; p.cout = cout; p.ntaps = ntaps;
  p.chunks_in = ceil_div(cin, 32);
  p.chunks_out = ceil_div(cout, 32);
  p.n_ablocks = ntaps * p.chunks_in;
  p.block_n = max(16, ((cout + 15) / 16) * 16);
  // Measured on B200 (tools/stall_profile.py, 3xTF32, metric shape):
  //   G=2 + lo tensors TMA-loaded 0.83 ms | G=1 + loaded lo 1.04 ms | G=2 + in-kernel lo 1.38 ms | G=1 + in-kernel 1.73 ms.
  // The N=64 tf32 MMAs are bound by shared-memory bandwidth (6 KB of operands per MMA); converting lo in the
  // kernel adds 2x its bytes to the same pipe, so wgrad reads materialised lo tensors when the caller has them.
  p.groups_per_cta = p.block_n <= 64 ? 2 : 1;
  { const char *e = getenv("PVCNN_WGRAD_G"); if (e && e[0] == '1') p.groups_per_cta = 1; }
  p.lo_in_kernel = (x_lo == nullptr || g_lo == nullptr) ? 1 : 0;
  { const char *e = getenv("PVCNN_WGRAD_LO"); if (e && e[0] == 'k') p.lo_in_kernel = 1; }
  const int ngroups = ceil_div(p.n_ablocks, 4);
  p.num_sets = ceil_div(ngroups, p.groups_per_cta);
  p.ksplit = max(1, kNumSMs / p.num_sets);
  if ((long long)p.ksplit > p.num_ktiles) p.ksplit = (int)p.num_ktiles;
  p.npass = npass;
  const uint32_t passes = npass > 1 ? 2 : 1;
  p.g_bytes = (uint32_t)p.chunks_out * WG_BLK * passes;
  p.stage_bytes = p.g_bytes + (uint32_t)(p.groups_per_cta * 4) * WG_BLK * passes;
  p.box_bytes = (uint32_t)rows * 128u;
  p.dw = dw;
  p.err = g_wg_err;
  p.dbg = pvb::stall_profile_buffer();
  p.ktile_list = ktile_list; p.ktile_count = ktile_count;
  if (bz_out) *bz_out = p.bz;
  if (by_out) *by_out = p.by;
  p.stages = min(WG_MAX_STAGES, (int)((227 * 1024 - 2048) / p.stage_bytes));
  PVB_CHECK_ARG(p.stages >= 2 && p.num_ktiles < (1LL << 31));
  PVB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)cout * cin * ntaps, s));

  CUtensorMap mx_hi, mx_lo, mg_hi, mg_lo;
  int rc;
  if ((rc = encode_map_5d_cl(&mx_hi, x_hi, cin, ldx, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  if ((rc = encode_map_5d_cl(&mx_lo, (npass > 1 && !p.lo_in_kernel) ? x_lo : x_hi, cin, ldx, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  if ((rc = encode_map_5d_cl(&mg_hi, g_hi, cout, ldg, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  if ((rc = encode_map_5d_cl(&mg_lo, (npass > 1 && !p.lo_in_kernel) ? g_lo : g_hi, cout, ldg, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
#define PVB_WG_LAUNCH(GV, T3)                                                                                         \
  do {                                                                                                                \
    PVB_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<GV, T3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    PVB_LAUNCH((conv_wgrad_kernel<GV, T3>), p.num_sets * p.ksplit, WG_THREADS, smem, s, mx_hi, mx_lo, mg_hi, mg_lo, p); \
  } while (0)
  if (p.groups_per_cta == 2) {
    if (npass > 1) PVB_WG_LAUNCH(2, true); else PVB_WG_LAUNCH(2, false);
  } else {
    if (npass > 1) PVB_WG_LAUNCH(1, true); else PVB_WG_LAUNCH(1, false);
  }
#undef PVB_WG_LAUNCH
  return 0;
}

int wgrad_dz_launch(int nb, int sx, int sy, int sz, int cin, int cout, const float *x_hi, const float *x_lo, int ldx,
                    const float *g_hi, const float *g_lo, int ldg, float *dw, int npass, cudaStream_t s,
                    const int4 *ktile_list, const int *ktile_count, int *bz_out, int *by_out);  // conv_wgrad_dz.cu

// Any cout: output channels are walked in blocks of 128 (the N extent of one TMEM accumulator set); a block reads its
// own channel slice of g (pointer offset inside the channels-last rows) and writes its own rows of dW.
int wgrad_launch(int nb, int sx, int sy, int sz, int cin, int cout, int ntaps, const float *x_hi, const float *x_lo,
                 int ldx, const float *g_hi, const float *g_lo, int ldg, float *dw, int npass, cudaStream_t s,
                 const int4 *ktile_list, const int *ktile_count, int *bz_out, int *by_out) {
  PVB_CHECK_ARG(cout > 0 && g_hi && dw && cin > 0 && (ntaps == 1 || ntaps == 27));
  // 3x3x3: the dz-merged kernel (conv_wgrad_dz.cu: one N = 3*Cout MMA per (dx,dy) pair, X blocks loaded once for three
  // taps) in blocks of 64 output channels; PVCNN_B200_WGRAD=v1 forces the per-tap kernel below (A/B, diagnostics)
  const char *e_wg = getenv("PVCNN_B200_WGRAD");
  // (measured: 0.77 -> 0.59 ms dense at 64->64 R=32; wider layers would need several 64-channel passes that each re-read X
  //  and are faster on the per-tap kernel's N=128 tiles: 64->128@16 0.18 vs 0.21 ms, so they stay there)
  if (ntaps == 27 && cout <= 64 && !(e_wg && e_wg[0] == 'v' && e_wg[1] == '1')) {
    bool ok = true;
    for (int n0 = 0; n0 < cout && ok; n0 += 64) {
      const int rc = wgrad_dz_launch(nb, sx, sy, sz, cin, min(64, cout - n0), x_hi, x_lo, ldx, g_hi + n0,
                                     g_lo ? g_lo + n0 : nullptr, ldg, dw + (size_t)n0 * cin * ntaps, npass, s, ktile_list,
                                     ktile_count, bz_out, by_out);
      if (rc == PVCNN_E_UNSUPPORTED) ok = false;
      else if (rc != 0) return rc;
    }
    if (ok) return 0;
  }
  for (int n0 = 0; n0 < cout; n0 += 128) {
    const int nblk = min(128, cout - n0);
    const int rc = wgrad_launch_block(nb, sx, sy, sz, cin, nblk, ntaps, x_hi, x_lo, ldx, g_hi + n0, g_lo ? g_lo + n0 : nullptr,
                                      ldg, dw + (size_t)n0 * cin * ntaps, npass, s, ktile_list, ktile_count, bz_out, by_out);
    if (rc != 0) return rc;
  }
  return 0;
}

}  // namespace pvb

extern "C" int pvcnn_conv_wgrad(int nb, int sx, int sy, int sz, int cin, int cout, int ntaps, const float *x_hi,
                                const float *x_lo, int ldx, const float *g_hi, const float *g_lo, int ldg, float *dw,
                                int npass, void *stream) {
  return pvb::wgrad_launch(nb, sx, sy, sz, cin, cout, ntaps, x_hi, x_lo, ldx, g_hi, g_lo, ldg, dw, npass,
                           (cudaStream_t)stream, nullptr, nullptr, nullptr, nullptr);
}
