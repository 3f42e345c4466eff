// conv_wgrad_dz.cu -- 3x3x3 weight gradient with the three dz taps of a (dx,dy) pair merged into ONE N = 3*Cout MMA.
//
//   dW[(dx,dy,dz)][ci][co] = sum_v X[v + (dx,dy,dz)][ci] * gY[v][co]
//                          = sum_u X[u + (dx,dy,0)][ci] * gY[u - (0,0,dz)][co]            (u = v shifted along z)
//
// so for a fixed (dx,dy) the X operand (the wide one: M = 128 = 4 blocks of 32 input channels) is THE SAME for dz = -1, 0,
// +1, and the three taps differ only in which z-shifted copy of the gY line they multiply.  The producer loads the gY line
// three times with the TMA z coordinate pre-shifted by +1 / 0 / -1 (out-of-range rows zero-filled = the conv's padding)
// side by side, and one tcgen05.mma [M=128, N=3*Cout, K=8] produces [dW(.,.,-1) | dW(.,.,0) | dW(.,.,+1)].
// Same idea as conv_halo.cu's N=192 trick; here it removes 2/3 of the X loads (the bulk of the L2->SMEM traffic of
// conv_wgrad.cu: every tap re-loaded its own shifted X blocks) and amortises the X operand's shared-memory reads over
// three taps.  Everything else follows conv_wgrad.cu: MN-major operands straight from the channels-last tensors
// (SWIZZLE_128B_ATOM_32B), split-K over the SMs, TMEM chains of <= 64 MMAs drained into fp32 registers (the tensor core
// truncates when it accumulates), 3xTF32 with materialised `lo` tensors and collector reuse of X_hi.
#include <cstdlib>

#include "common.cuh"
#include "umma.cuh"

namespace pvb {
using namespace umma;

constexpr int WZ_THREADS = 512;        // w0 TMA, w1 MMA, w2 TMEM alloc, w4-15 drain (lane quarter = w % 4, dz = (w - 4) / 4)
constexpr int WZ_ROWS = 32;            // voxels per k-tile (4 MMA k-steps)
constexpr uint32_t WZ_BLK = WZ_ROWS * 128;  // bytes of one [32 rows x 32 channels] block
constexpr int WZ_MAX_STAGES = 5;
// Measured (tools/wgrad_bench.py, dense 64->64 R=32 3xTF32): 32-voxel stages 0.59 ms; streaming each k-tile as two 16-voxel
// stages (5-deep ring) 0.85 ms -- the kernel is bound by the L2->SM rate of its TMA boxes (80 KB per k-tile = 42 B/clk/SM,
// ~11.5 TB/s chip-wide), not by ring depth, so the remaining lever is bytes: the gY `lo` copies (24 of the 80 KB) are derived
// in shared memory by the drain warps instead of being TMA-loaded (GCONV), which also spares the 134 MB gY_lo tensors in HBM.

struct WgradDzParams {
  int nb, sx, sy, sz;
  int bz, by, tz, ty;
  long long num_ktiles;
  int cin, cout;
  int chunks_in, chunks_out;
  int n_ablocks;            // 9 (dx,dy) pairs x chunks_in
  int ksplit;
  int n3;                   // MMA N = 3 * 32 * chunks_out (<= 192)
  int npass, ksteps, drain_tiles, stages;
  uint32_t stage_bytes, g_bytes, box_bytes;
  float *dw;                // [cout][cin][27], accumulated with atomics (zeroed by the launcher)
  int *err;
  const int4 *ktile_list;   // optional compact list of (z0, y0, x0, b) k-tiles that can contribute
  const int *ktile_count;
};

template <bool THREE, bool GCONV>
__global__ void __launch_bounds__(WZ_THREADS, 1)
    conv_wgrad_dz_kernel(const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                         const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                         const WgradDzParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[WZ_MAX_STAGES], ready_bar[WZ_MAX_STAGES], empty_bar[WZ_MAX_STAGES], tmem_full_bar[2],
      tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = blockIdx.x / p.ksplit, split = blockIdx.x % p.ksplit;
  const int ab0 = group * 4;
  const int nab = min(4, p.n_ablocks - ab0);
  const uint32_t tmem_cols = 512;   // 2 buffers x n3 (<= 192) columns; one CTA per SM

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_x_hi);
    prefetch_tensormap(&map_g_hi);
    if (THREE) { prefetch_tensormap(&map_x_lo); if (!GCONV) prefetch_tensormap(&map_g_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&ready_bar[s], 12 * 32); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 12 * 32); }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  const long long num_ktiles = p.ktile_list ? (long long)__ldg(p.ktile_count) : p.num_ktiles;
  const long long per_cta = (num_ktiles + p.ksplit - 1) / p.ksplit;
  const long long kt_begin = (long long)split * per_cta;
  const long long my_tiles = max(0LL, min(num_ktiles, kt_begin + per_cta) - kt_begin);

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      const uint32_t passes = THREE ? 2u : 1u;
      const uint32_t tx_bytes = p.box_bytes * ((uint32_t)(3 * p.chunks_out) * ((THREE && !GCONV) ? 2u : 1u) + (uint32_t)nab * passes);
      int stage = 0;
      uint32_t phase = 0;
      int blk_c[4], blk_dx[4], blk_dy[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const int ab = min(ab0 + a, p.n_ablocks - 1);
        const int p9 = ab / p.chunks_in;
        blk_c[a] = (ab - p9 * p.chunks_in) * 32;
        blk_dx[a] = p9 / 3 - 1;
        blk_dy[a] = p9 % 3 - 1;
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
        mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 31);
        uint8_t *st = smem + (size_t)stage * p.stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        // gY line, three copies: B_dz[k] = gY[k - dz]  ->  box starts at z0 - dz = z0 + 1 - dzi
        for (int dzi = 0; dzi < 3; ++dzi)
          for (int cc = 0; cc < p.chunks_out; ++cc) {
            uint8_t *dst = st + (size_t)(dzi * p.chunks_out + cc) * WZ_BLK;
            tma_load_5d(dst, &map_g_hi, &full_bar[stage], cc * 32, z0 + 1 - dzi, y0, x0, b);
            if (THREE && !GCONV)
              tma_load_5d(dst + (size_t)3 * p.chunks_out * WZ_BLK, &map_g_lo, &full_bar[stage], cc * 32, z0 + 1 - dzi, y0, x0, b);
          }
        uint8_t *sa = st + p.g_bytes;
#pragma unroll
        for (int a = 0; a < 4; ++a)
          if (a < nab) {
            tma_load_5d(sa + (size_t)a * WZ_BLK, &map_x_hi, &full_bar[stage], blk_c[a], z0, y0 + blk_dy[a], x0 + blk_dx[a], b);
            if (THREE)
              tma_load_5d(sa + (size_t)(4 + a) * WZ_BLK, &map_x_lo, &full_bar[stage], blk_c[a], z0, y0 + blk_dy[a],
                          x0 + blk_dx[a], b);
          }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (++tzi == p.tz) { tzi = 0; if (++tyi == p.ty) { tyi = 0; if (++x0 == p.sx) { x0 = 0; ++b; } } }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(128, p.n3, /*a MN-major*/ 1, /*b MN-major*/ 1);
      // MN-major tf32 operands: 128B swizzle with 32-byte atoms; LBO = stride between 32-channel blocks (one block = 4 KB),
      // SBO = stride between 4-row K groups; +1024 bytes (64 units) per 8-voxel k-step
      constexpr uint32_t dhi = desc_hi32(512, kLayoutSW128Base32);
      int stage = 0;
      uint32_t phase = 0;
      int chain = 0;
      for (long long t = 0; t < my_tiles; ++t) {
        const int pos = (int)(t % p.drain_tiles);
        const int buf = chain & 1;
        if (pos == 0) {
          mbar_wait(&tmem_empty_bar[buf], ((chain >> 1) & 1) ^ 1, p.err, 32);
          tc_fence_after();
        }
        mbar_wait((THREE && GCONV) ? &ready_bar[stage] : &full_bar[stage], phase, p.err, 33);
        tc_fence_after();
        const uint32_t st = smem_u32(smem + (size_t)stage * p.stage_bytes);
        const uint32_t g_hi = desc_lo32(st, WZ_BLK);
        const uint32_t g_lo = g_hi + (uint32_t)(3 * p.chunks_out) * (WZ_BLK >> 4);
        const uint32_t a_hi = g_hi + (p.g_bytes >> 4);
        const uint32_t a_lo = a_hi + 4u * (WZ_BLK >> 4);
        const uint32_t d = tmem_base + (uint32_t)buf * (uint32_t)p.n3;
        auto kstep = [&](int ks) {
          const uint32_t ko = (uint32_t)ks * 64u;
          const uint32_t accum = (pos == 0 && ks == 0) ? 0u : 1u;
          if (THREE) {
            mma_tf32_lo32_c<kCollFill>(d, a_hi + ko, g_hi + ko, dhi, idesc, accum);
            mma_tf32_lo32_c<kCollLastUse>(d, a_hi + ko, g_lo + ko, dhi, idesc, 1u);
            mma_tf32_lo32(d, a_lo + ko, g_hi + ko, dhi, idesc, 1u);
          } else {
            mma_tf32_lo32(d, a_hi + ko, g_hi + ko, dhi, idesc, accum);
          }
        };
        if (p.ksteps == 4) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) kstep(ks);
        } else {
          for (int ks = 0; ks < p.ksteps; ++ks) kstep(ks);
        }
        mma_commit(&empty_bar[stage]);
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
        if (pos == p.drain_tiles - 1 || t == my_tiles - 1) {
          mma_commit(&tmem_full_bar[buf]);
          ++chain;
        }
      }
    }
  } else if (warp >= 4) {
    // ================================ drain / epilogue ================================
    const int q = (warp - 4) & 3;    // TMEM lane quarter (= warp % 4)
    const int dzi = (warp - 4) >> 2; // this warp's third of the columns = one dz tap
    const int cols = 32 * p.chunks_out;  // columns per dz (<= 64)
    float acc[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) acc[i] = 0.0f;
    const long long nchains = (my_tiles + p.drain_tiles - 1) / p.drain_tiles;
    const int tid = threadIdx.x - 4 * 32;   // 0..383
    int stage = 0;
    uint32_t phase = 0;
    for (long long ch = 0; ch < nchains; ++ch) {
      if (THREE && GCONV) {
        // derive the gY `lo` copies of this chain's stages in shared memory (lo = g - trunc_tf32(g), element-wise on the
        // swizzled bytes), then hand each stage to the MMA warp; the drain below only starts once the chain's last stage
        // has been converted, so the MMA warp never waits for a conversion that waits for it
        const long long t0 = ch * p.drain_tiles, t1 = min(my_tiles, t0 + p.drain_tiles);
        const int n16 = (int)((3u * (uint32_t)p.chunks_out * WZ_BLK) >> 4);
        for (long long t = t0; t < t1; ++t) {
          mbar_wait(&full_bar[stage], phase, p.err, 35);
          const float4 *src = reinterpret_cast<const float4 *>(smem + (size_t)stage * p.stage_bytes);
          float4 *dst = reinterpret_cast<float4 *>(smem + (size_t)stage * p.stage_bytes + (size_t)3 * p.chunks_out * WZ_BLK);
          for (int i = tid; i < n16; i += 12 * 32) {
            const float4 v = src[i];
            float4 l;
            l.x = __fsub_rn(v.x, __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
            l.y = __fsub_rn(v.y, __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
            l.z = __fsub_rn(v.z, __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
            l.w = __fsub_rn(v.w, __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
            dst[i] = l;
          }
          fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async-proxy reads
          mbar_arrive(&ready_bar[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
      const int buf = (int)(ch & 1);
      mbar_wait(&tmem_full_bar[buf], (uint32_t)((ch >> 1) & 1), p.err, 34);
      tc_fence_after();
      const uint32_t tb = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)buf * (uint32_t)p.n3 + (uint32_t)(dzi * cols);
#pragma unroll
      for (int c0 = 0; c0 < 64; c0 += 16) {
        if (c0 < cols) {  // warp-uniform
          float v[16];
          tmem_ld16(tb + (uint32_t)c0, v);
#pragma unroll
          for (int i = 0; i < 16; ++i) acc[c0 + i] += v[i];
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty_bar[buf]);
    }
    // scatter-add this CTA's partial dW:  D row m = (X block m/32, channel m%32), column = (dz, cout)
    const int m = q * 32 + lane;
    const int a = m >> 5;
    if (a < nab) {
      const int ab = ab0 + a;
      const int p9 = ab / p.chunks_in, cc = ab - p9 * p.chunks_in;
      const int ci = cc * 32 + (m & 31);
      const int tap = p9 * 3 + dzi;   // torch tap index = ((dx+1)*3 + (dy+1))*3 + (dz+1)
      if (ci < p.cin) {
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c < cols && c < p.cout) atomicAdd(p.dw + ((size_t)c * p.cin + ci) * 27 + tap, acc[c]);
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
int *device_error_flag(int slot);           // conv_igemm.cu

// cout <= 64 valid channels of g (the caller walks wider layers in blocks of 64); dw pre-zeroed by this call.
// Returns PVCNN_E_UNSUPPORTED when the 3xTF32 mode has no materialised lo tensors (caller falls back to conv_wgrad.cu).
int wgrad_dz_launch(int nb, int sx, int sy, int sz, int cin, int cout, const float *x_hi, const float *x_lo, int ldx,
                    const float *g_hi, const float *g_lo, int ldg, float *dw, int npass, cudaStream_t s,
                    const int4 *ktile_list, const int *ktile_count, int *bz_out, int *by_out) {
  if (cout > 64 || (npass > 1 && !x_lo)) return PVCNN_E_UNSUPPORTED;
  bool gconv = npass > 1 && g_lo == nullptr;     // no materialised gY_lo: derive it in shared memory
  { const char *e = getenv("PVCNN_WGRAD_GLO"); if (npass > 1 && e && e[0] == 'k') gconv = true; if (npass > 1 && e && e[0] == 't' && g_lo) gconv = false; }
  PVB_CHECK_ARG(nb > 0 && sx > 0 && sy > 0 && sz > 0 && cin > 0 && cout > 0 && x_hi && g_hi && dw && ldx % 4 == 0 && ldg % 4 == 0);
  WgradDzParams p{};
  p.err = device_error_flag(2);
  PVB_CHECK_ARG(p.err != nullptr);
  p.nb = nb; p.sx = sx; p.sy = sy; p.sz = sz;
  p.bz = ((min(sz, WZ_ROWS) + 7) / 8) * 8;      // must mirror conv_wgrad.cu (the activity lists are built in these units)
  p.by = max(1, WZ_ROWS / p.bz);
  if (p.by > sy) p.by = sy;
  const int rows = p.bz * p.by;
  p.ksteps = rows / 8;
  p.tz = ceil_div(sz, p.bz);
  p.ty = ceil_div(sy, p.by);
  p.num_ktiles = (long long)nb * sx * p.ty * p.tz;
  p.cin = cin; p.cout = cout;
  p.chunks_in = ceil_div(cin, 32);
  p.chunks_out = ceil_div(cout, 32);
  p.n_ablocks = 9 * p.chunks_in;
  p.n3 = 3 * 32 * p.chunks_out;
  const int ngroups = ceil_div(p.n_ablocks, 4);
  p.ksplit = max(1, kNumSMs / ngroups);
  if ((long long)p.ksplit > p.num_ktiles) p.ksplit = (int)p.num_ktiles;
  p.npass = npass;
  const uint32_t passes = npass > 1 ? 2 : 1;
  p.g_bytes = (uint32_t)(3 * p.chunks_out) * WZ_BLK * passes;
  p.stage_bytes = p.g_bytes + 4u * WZ_BLK * passes;
  p.box_bytes = (uint32_t)rows * 128u;
  p.drain_tiles = max(1, 64 / (p.ksteps * (npass > 1 ? 3 : 1)));   // <= 64 MMAs per TMEM chain
  p.dw = dw;
  p.ktile_list = ktile_list; p.ktile_count = ktile_count;
  if (bz_out) *bz_out = p.bz;
  if (by_out) *by_out = p.by;
  p.stages = min(WZ_MAX_STAGES, (int)((227 * 1024 - 2048) / p.stage_bytes));
  PVB_CHECK_ARG(p.stages >= 2 && p.num_ktiles < (1LL << 31));
  PVB_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)cout * cin * 27, s));

  CUtensorMap mx_hi, mx_lo, mg_hi, mg_lo;
  int rc;
  if ((rc = encode_map_5d_cl(&mx_hi, x_hi, cin, ldx, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  if ((rc = encode_map_5d_cl(&mx_lo, npass > 1 ? x_lo : x_hi, cin, ldx, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  if ((rc = encode_map_5d_cl(&mg_hi, g_hi, cout, ldg, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  if ((rc = encode_map_5d_cl(&mg_lo, (npass > 1 && !gconv) ? g_lo : g_hi, cout, ldg, nb, sx, sy, sz, p.bz, p.by, 1, true))) return rc;
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024;
  const int grid = ngroups * p.ksplit;
  if (npass > 1 && gconv) {
    PVB_CUDA((cudaFuncSetAttribute(conv_wgrad_dz_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    PVB_LAUNCH((conv_wgrad_dz_kernel<true, true>), grid, WZ_THREADS, smem, s, mx_hi, mx_lo, mg_hi, mg_lo, p);
  } else if (npass > 1) {
    PVB_CUDA((cudaFuncSetAttribute(conv_wgrad_dz_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    PVB_LAUNCH((conv_wgrad_dz_kernel<true, false>), grid, WZ_THREADS, smem, s, mx_hi, mx_lo, mg_hi, mg_lo, p);
  } else {
    PVB_CUDA((cudaFuncSetAttribute(conv_wgrad_dz_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)));
    PVB_LAUNCH((conv_wgrad_dz_kernel<false, false>), grid, WZ_THREADS, smem, s, mx_hi, mx_lo, mg_hi, mg_lo, p);
  }
  return 0;
}

}  // namespace pvb
