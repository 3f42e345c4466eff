// conv_igemm.cu -- im2col-free implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
// Replaces the reference's cuDNN calls: nn.Conv3d(k=3, pad=1) (modules/pvconv.py:21,24) and the
// 1x1 nn.Conv1d of SharedMLP (modules/shared_mlp.py:10), forward and data-gradient.
//
//   out[v, n] = bias[n] + sum_{tap, c} A[v + off(tap), c] * W[tap][n][c]
//
//   * activations are channels-last ([B, X, Y, Z, C], C contiguous); one M-tile is a box of up to 128
//     voxels fetched by ONE 5-D TMA load per (tap, 32-channel chunk).  The tap offset is added to the box
//     coordinates and TMA's out-of-bounds zero fill implements the conv's zero padding -- no im2col
//     buffer, no halo copies, no boundary branches.
//   * operands land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes
//     directly (kind::tf32, M=128, N=block_n, K=8 per instruction), accumulators live in TMEM,
//     double-buffered so the epilogue of tile i overlaps the MMAs of tile i+1.
//   * warp-specialised persistent CTAs (one per SM): warp 0 = TMA producer, warp 1 = MMA issuer,
//     warp 2 = TMEM allocator, warps 4-11 = epilogue (tcgen05.ld -> smem transpose -> bias/BN/ReLU -> global).
//   * fp32 parity: tf32 has a 10-bit mantissa, the north-star tolerance is 1e-5.  npass=3 runs the
//     error-compensated split  a*b ~= a_hi*b_hi + a_hi*b_lo + a_lo*b_hi.  Measured on B200
//     (tests/test_igemm_gpu.py::test_hw_rounding_probe): kind::tf32 TRUNCATES fp32 operands, so the
//     raw fp32 tensor serves as a_hi and only a_lo = x - trunc(x) is materialised.  The tensor core
//     also truncates when it accumulates (~2^-24 relative bias per MMA), which over a 27-tap, K=1728
//     chain reaches ~1.5e-5; therefore (1) the tiny correction terms get their own TMEM
//     accumulator and (2) the main chain is split over up to 3 accumulators of <= 72 MMAs each;
//     the epilogue adds the partial sums in round-to-nearest fp32.
//     npass=1 is plain TF32 (what cuDNN does by default in the reference).
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "umma.cuh"

namespace pvb {
using namespace umma;

constexpr int IG_BLOCK_M = 128;
constexpr int IG_KC = 32;  // channels per k-block = 128 bytes = one swizzle row
constexpr int IG_THREADS = 384;   // warps 0-3: TMA, MMA issue, TMEM allocator, idle; warps 4-11: epilogue
constexpr int IG_MAX_STAGES = 8;
constexpr uint32_t IG_EPI_WARP_BYTES = 32 * 16 * 4;   // an epilogue warp's 32 x 16 staging tile
constexpr int IG_EPI_BYTES = 8 * IG_EPI_WARP_BYTES;
constexpr int IG_MAX_CHAIN = 72;  // MMAs accumulated into one TMEM accumulator
constexpr uint32_t IG_A_TILE_BYTES = IG_BLOCK_M * IG_KC * 4;  // 16 KB

struct IgemmParams {
  int nb, sx, sy, sz;
  int bx, by, bz;  // rows per tile = bx*by*bz <= 128
  int tx, ty, tz;
  int num_m_tiles, n_tiles;
  int cin_chunks, ntaps;
  int cout, block_n;
  int npass, stages;
  int acc_split;      // main accumulators per tile (bounds the RZ-rounded tensor-core accumulation chain)
  int acc_slots;      // acc_split + (npass > 1): the correction terms get their own accumulator
  int acc_bufs;       // 2 when two tiles' accumulators fit in TMEM (512 columns), else 1
  int ldo;
  uint32_t stage_bytes, a_bytes, b_bytes, tx_bytes, tmem_cols;
  const float *bias;
  float *out;
  int *err;
  IgemmEpilogue ep;
};

template <bool THREE>
__global__ void __launch_bounds__(IG_THREADS, 1)
    igemm_conv_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                      const __grid_constant__ CUtensorMap map_w_hi, const __grid_constant__ CUtensorMap map_w_lo,
                      const IgemmParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[IG_MAX_STAGES], empty_bar[IG_MAX_STAGES], tmem_full_bar[2], tmem_empty_bar[2];
  __shared__ uint32_t tmem_base_smem;

  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = p.num_m_tiles * p.n_tiles;
  const int num_kb = p.ntaps * p.cin_chunks;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&map_a_hi);
    prefetch_tensormap(&map_w_hi);
    if (THREE) {
      prefetch_tensormap(&map_a_lo);
      prefetch_tensormap(&map_w_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full_bar[a], 1);
      mbar_init(&tmem_empty_bar[a], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_smem, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int mt = tile / p.n_tiles;
        const int z0 = (mt % p.tz) * p.bz; mt /= p.tz;
        const int y0 = (mt % p.ty) * p.by; mt /= p.ty;
        const int x0 = (mt % p.tx) * p.bx; mt /= p.tx;
        const int b = mt;
        int tap = 0, cc = 0;  // advanced incrementally: no runtime division on the producer's critical path
        for (int kb = 0; kb < num_kb; ++kb) {
          int dx = 0, dy = 0, dz = 0;
          if (p.ntaps == 27) { dx = tap / 9 - 1; dy = (tap / 3) % 3 - 1; dz = tap % 3 - 1; }
          mbar_wait(&empty_bar[stage], phase ^ 1, p.err, 1);
          uint8_t *st = smem + (size_t)stage * p.stage_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], p.tx_bytes);
          tma_load_5d(st, &map_a_hi, &full_bar[stage], cc * IG_KC, z0 + dz, y0 + dy, x0 + dx, b);
          tma_load_3d(st + p.a_bytes, &map_w_hi, &full_bar[stage], cc * IG_KC, n_tile * p.block_n, tap);
          if (THREE) {
            tma_load_5d(st + IG_A_TILE_BYTES, &map_a_lo, &full_bar[stage], cc * IG_KC, z0 + dz, y0 + dy, x0 + dx, b);
            tma_load_3d(st + p.a_bytes + p.b_bytes / 2, &map_w_lo, &full_bar[stage], cc * IG_KC, n_tile * p.block_n,
                        tap);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
          if (++cc == p.cin_chunks) { cc = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (elect_one()) {
      const uint32_t idesc = make_idesc_tf32(IG_BLOCK_M, p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const int acc = it % p.acc_bufs;
        const uint32_t acc_phase = (it / p.acc_bufs) & 1;
        mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1, p.err, 2);
        tc_fence_after();
        const uint32_t tile_tmem = tmem_base + (uint32_t)(acc * p.acc_slots * p.block_n);
        const uint32_t corr_tmem = tile_tmem + (uint32_t)(p.acc_split * p.block_n);
        // k-blocks are dealt to the main accumulators in contiguous ranges: slot = floor(kb * acc_split / num_kb), tracked
        // incrementally (this thread's instruction stream is the issue path of every MMA: no divisions in the loop)
        int slot = 0, slot_rem = 0;
        bool fresh = true;
        for (int kb = 0; kb < num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase, p.err, 3);
          tc_fence_after();
          const uint32_t d_tmem = tile_tmem + (uint32_t)(slot * p.block_n);
          constexpr uint32_t dhi = desc_hi32(1024, kLayoutSW128);
          const uint32_t a_hi = desc_lo32(smem_u32(smem + (size_t)stage * p.stage_bytes), 0);
          const uint32_t b_hi = a_hi + (p.a_bytes >> 4);
          if (THREE) {
            const uint32_t a_lo = a_hi + (IG_A_TILE_BYTES >> 4);
            const uint32_t b_lo = b_hi + (p.b_bytes >> 5);
            const uint32_t fresh_corr = kb != 0;
#pragma unroll
            for (int k = 0; k < IG_KC / 8; ++k) {
              const uint32_t ko = (uint32_t)k * 2u;  // advance 8 tf32 = 32 bytes inside the 128B swizzle row
              mma_tf32_lo32(d_tmem, a_hi + ko, b_hi + ko, dhi, idesc, (k == 0) ? (fresh ? 0u : 1u) : 1u);
              mma_tf32_lo32(corr_tmem, a_hi + ko, b_lo + ko, dhi, idesc, (k == 0) ? fresh_corr : 1u);
              mma_tf32_lo32(corr_tmem, a_lo + ko, b_hi + ko, dhi, idesc, 1u);
            }
          } else {
#pragma unroll
            for (int k = 0; k < IG_KC / 8; ++k)
              mma_tf32_lo32(d_tmem, a_hi + (uint32_t)k * 2u, b_hi + (uint32_t)k * 2u, dhi, idesc,
                            (k == 0) ? (fresh ? 0u : 1u) : 1u);
          }
          fresh = false;
          slot_rem += p.acc_split;
          if (slot_rem >= num_kb) { slot_rem -= num_kb; ++slot; fresh = true; }
          mma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        mma_commit(&tmem_full_bar[acc]);
      }
    }
  } else if (warp >= 4) {
    // ================================ epilogue ================================
    // Eight warps: warp w reads TMEM lanes 32*(w%4).. (the hardware's lane-quarter rule) and, of the tile's 32-column
    // chunks, the even ones (warps 4-7) or the odd ones (warps 8-11).
    // TMEM -> registers (thread = one output row, 32 columns at a time; partial accumulators are combined here in
    // round-to-nearest fp32: the tensor core's own accumulation truncates, so its chains are kept short, see DESIGN.md
    // "accumulation") -> a warp-private, XOR-swizzled shared-memory tile of 32 rows x 16 columns -> registers in the
    // transposed role (4 lanes x float4 = 64 contiguous bytes of one row, 8 rows per instruction) -> global.  Storing
    // straight from the row-per-thread layout (16 bytes per lane at a row stride of ldo floats) ran the wide layers of the
    // network heads at < 1 TB/s.  Bias, the optional per-cloud bias, BatchNorm + (Leaky)ReLU and the tf32 lo split are
    // applied in the transposed role, where a lane owns the same 4 columns of every 16-column group for all of its rows;
    // their coefficients are fetched before the wait for the accumulator.
    const int we = warp & 3;          // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;   // 0: even chunks, 1: odd chunks
    const int m = we * 32 + lane;
    const int lz = m % p.bz, ly = (m / p.bz) % p.by, lx = m / (p.bz * p.by);
    const uint32_t stage_tile = smem_u32(smem + (size_t)p.stages * p.stage_bytes) + (uint32_t)(warp - 4) * IG_EPI_WARP_BYTES;
    const int rsub = lane >> 2, cq = (lane & 3) * 4;
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int acc = it % p.acc_bufs;
      const uint32_t acc_phase = (it / p.acc_bufs) & 1;
      const int n_tile = tile % p.n_tiles;
      int mt = tile / p.n_tiles;
      const int z = (mt % p.tz) * p.bz + lz; mt /= p.tz;
      const int y = (mt % p.ty) * p.by + ly; mt /= p.ty;
      const int x = (mt % p.tx) * p.bx + lx; mt /= p.tx;
      const int b = mt;
      const bool valid = (lx < p.bx) && x < p.sx && y < p.sy && z < p.sz;
      const long long row_off = valid ? (long long)(((((size_t)b * p.sx + x) * p.sy + y) * p.sz + z) * p.ldo +
                                                    (size_t)n_tile * p.block_n)
                                      : -1;
      const int grp = p.ep.group_bias ? z / p.ep.group_rows : 0;   // rows are flat along z whenever an epilogue is given
      long long roff[4];
      int rgrp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        roff[i] = __shfl_sync(0xffffffffu, row_off, i * 8 + rsub);
        rgrp[i] = __shfl_sync(0xffffffffu, grp, i * 8 + rsub);
      }
      const int ncols = min(p.block_n, p.cout - n_tile * p.block_n);
      const int nchunks = (ncols + 31) >> 5;
      // coefficients of this lane's columns: chunk slot cs (this warp's cs-th chunk), 16-column group g, 4 columns
      float bs[2][2][4], sc[2][2][4], sh[2][2][4];
#pragma unroll
      for (int cs = 0; cs < 2; ++cs)
#pragma unroll
        for (int g = 0; g < 2; ++g)
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int cl = (2 * cs + half) * 32 + g * 16 + cq + k;
            const int cg = n_tile * p.block_n + cl;
            const bool ok = cl < ncols;
            bs[cs][g][k] = (ok && p.bias) ? __ldg(p.bias + cg) : 0.0f;
            sc[cs][g][k] = (ok && p.ep.scale) ? __ldg(p.ep.scale + cg) : 1.0f;
            sh[cs][g][k] = (ok && p.ep.scale) ? __ldg(p.ep.shift + cg) : 0.0f;
          }
      mbar_wait(&tmem_full_bar[acc], acc_phase, p.err, 4);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(we * 32) << 16) + (uint32_t)(acc * p.acc_slots * p.block_n);
      if (half >= nchunks) {   // nothing to read for this warp: hand the accumulator back at once
        tc_fence_before();
        mbar_arrive(&tmem_empty_bar[acc]);
      }
#pragma unroll
      for (int cs = 0; cs < 2; ++cs) {
        const int ch = 2 * cs + half;
        if (ch >= nchunks) break;
        const int c0 = ch * 32;
        float v[32];
        const bool wide = c0 + 32 <= p.block_n;   // block_n is a multiple of 16: the last chunk may hold 16 columns
        {
          uint32_t r[32];
          if (wide) {
            tmem_ld32_nowait(taddr + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          } else {
            tmem_ld16(taddr + c0, v);
#pragma unroll
            for (int i = 16; i < 32; ++i) v[i] = 0.0f;
          }
          for (int sl = 1; sl < p.acc_slots; ++sl) {
            if (wide) {
              tmem_ld32_nowait(taddr + sl * p.block_n + c0, r);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] += __uint_as_float(r[i]);
            } else {
              float t[16];
              tmem_ld16(taddr + sl * p.block_n + c0, t);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[i] += t[i];
            }
          }
        }
        if (ch + 2 >= nchunks) {   // this warp's last chunk: the accumulator can be handed back before the stores
          tc_fence_before();
          mbar_arrive(&tmem_empty_bar[acc]);
        }
#pragma unroll
        for (int g = 0; g < 2; ++g) {
          // stage 32 rows x 16 columns; the 16-byte column group j of row r sits at slot j ^ ((r >> 1) & 3)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(
                             stage_tile + (uint32_t)(lane * 16 + 4 * (j ^ ((lane >> 1) & 3))) * 4u),
                         "f"(v[g * 16 + 4 * j]), "f"(v[g * 16 + 4 * j + 1]), "f"(v[g * 16 + 4 * j + 2]),
                         "f"(v[g * 16 + 4 * j + 3])
                         : "memory");
          __syncwarp();
          const int cl = c0 + g * 16 + cq;      // this lane's first column inside the n-tile
          const int nv = ncols - cl;            // valid columns among the lane's 4 (<= 0: none)
          if (nv > 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (roff[i] < 0) continue;
              const int rr = i * 8 + rsub;
              float o[4];
              asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3])
                           : "r"(stage_tile + (uint32_t)(rr * 16 + 4 * ((lane & 3) ^ ((rr >> 1) & 3))) * 4u)
                           : "memory");
#pragma unroll
              for (int k = 0; k < 4; ++k) o[k] += bs[cs][g][k];
              if (p.ep.group_bias) {
                const float *gb = p.ep.group_bias + (size_t)rgrp[i] * p.ep.group_ld + n_tile * p.block_n + cl;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < nv) o[k] += __ldg(gb + k);
              }
              if (p.ep.scale) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                  const float t = fmaf(o[k], sc[cs][g][k], sh[cs][g][k]);
                  o[k] = t > 0.0f ? t : t * p.ep.slope;
                }
              }
              float *dst = p.out + roff[i] + cl;
              if (nv >= 4) {
                *reinterpret_cast<float4 *>(dst) = make_float4(o[0], o[1], o[2], o[3]);
              } else {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  if (k < nv) dst[k] = o[k];
              }
              if (p.ep.out_lo) {
                float *dlo = p.ep.out_lo + roff[i] + cl;
#pragma unroll
                for (int k = 0; k < 4; ++k) o[k] = __fsub_rn(o[k], __uint_as_float(__float_as_uint(o[k]) & 0xFFFFE000u));
                if (nv >= 4) {
                  *reinterpret_cast<float4 *>(dlo) = make_float4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                  for (int k = 0; k < 4; ++k)
                    if (k < nv) dlo[k] = o[k];
                }
              }
            }
          }
          __syncwarp();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, p.tmem_cols);
  }
}

// ------------------------------------------------------------------------------------------------
// Operand preparation
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // exact in tf32 (10-bit mantissa)
  lo = __uint_as_float(__float_as_uint(__fsub_rn(x, hi)) & 0xFFFFE000u);
}

__global__ void __launch_bounds__(256) split_tf32_kernel(long long n4, const float4 *__restrict__ x,
                                                         float4 *__restrict__ hi, float4 *__restrict__ lo) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    float4 h, l;
    split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
    if (hi) hi[i] = h;
    lo[i] = l;
  }
}

// w [cout][cin][ntaps] (torch Conv3d / Conv1d weight, taps flattened kd*9+kh*3+kw) ->
//   mode 0 (forward): wr[tap][cout][ld]          = w[co][ci][tap]
//   mode 1 (dgrad)  : wr[tap][cin ][ld]  (K=cout) = w[co][ci][ntaps-1-tap]
__global__ void __launch_bounds__(256) weight_prep_kernel(int cout, int cin, int ntaps, int mode, int ld,
                                                          const float *__restrict__ w, float *__restrict__ hi,
                                                          float *__restrict__ lo) {
  const int rows = mode == 0 ? cout : cin;   // GEMM N
  const int kdim = mode == 0 ? cin : cout;   // GEMM K
  const long long total = (long long)ntaps * rows * ld;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % ld);
    const int r = (int)((i / ld) % rows);
    const int tap = (int)(i / ((long long)ld * rows));
    float v = 0.0f;
    if (k < kdim) {
      if (mode == 0) v = w[((size_t)r * cin + k) * ntaps + tap];
      else v = w[((size_t)k * cin + r) * ntaps + (ntaps - 1 - tap)];
    }
    float h, l;
    split_tf32(v, h, l);
    hi[i] = h;
    lo[i] = l;
  }
}

// ------------------------------------------------------------------------------------------------
// Host side: tensor maps
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

static int encode_map(CUtensorMap *map, const void *ptr, int rank, const cuuint64_t *gdim, const cuuint64_t *gstride,
                      const cuuint32_t *box, bool atom32 = false) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return PVCNN_E_UNSUPPORTED;
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void *>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (200000 + (int)r);
}

// 5-D map over a channels-last tensor [nb,sx,sy,sz,ld] with k valid channels and a {32, bz, by, bx, 1} box.
// atom32: 128B swizzle with 32-byte atoms -- the only shared-memory layout tcgen05 accepts for MN-major
// tf32 operands (UMMA layout type SWIZZLE_128B_BASE32B).
int encode_map_5d_cl(CUtensorMap *map, const float *ptr, int k, int ld, int nb, int sx, int sy, int sz, int bz, int by,
                     int bx, bool atom32) {
  cuuint64_t gdim[5] = {(cuuint64_t)k, (cuuint64_t)sz, (cuuint64_t)sy, (cuuint64_t)sx, (cuuint64_t)nb};
  cuuint64_t gstr[4] = {(cuuint64_t)ld * 4, (cuuint64_t)sz * ld * 4, (cuuint64_t)sy * sz * ld * 4,
                        (cuuint64_t)sx * sy * sz * ld * 4};
  cuuint32_t box[5] = {(cuuint32_t)IG_KC, (cuuint32_t)bz, (cuuint32_t)by, (cuuint32_t)bx, 1};
  return encode_map(map, ptr, 5, gdim, gstr, box, atom32);
}

// generic tiled map; swizzle_kind: 128 (SWIZZLE_128B), 64 (SWIZZLE_64B), 320 (SWIZZLE_32B), 32 (SWIZZLE_128B_ATOM_32B)
int encode_map_generic(CUtensorMap *map, const void *ptr, int rank, const unsigned long long *gdim,
                       const unsigned long long *gstride_bytes, const unsigned *box, int swizzle_kind) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return PVCNN_E_UNSUPPORTED;
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5] = {1, 1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = gdim[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) gs[i] = gstride_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_kind == 320 ? CU_TENSOR_MAP_SWIZZLE_32B
                                : swizzle_kind == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_kind == 32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B
                                                     : CU_TENSOR_MAP_SWIZZLE_128B;
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<void *>(ptr), gd, gs, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : (200000 + (int)r);
}

int conv_halo_launch(int nb, int sx, int sy, int sz, int k, int cout, const float *a, int lda, const float *w_hi,
                     const float *w_lo, int ldw, const float *bias, float *out, int ldo, int npass,
                     cudaStream_t stream, const int4 *unit_list, const int *unit_count);  // conv_halo.cu

// Optional device buffer [8] of stall-cycle counters written by CTA 0 of the tensor-core kernels when
// PVCNN_STALL_PROFILE=1 (tools/stall_profile.py reads it back through pvcnn_stall_profile_read).
// Per-device scratch (several GPUs may be driven from one process, e.g. nn.DataParallel in the reference's train.py:181):
// every device gets its own error-flag ints and stall-counter buffer, created under a mutex on first use.
static std::mutex g_dev_mu;
static int *g_dev_flags[64][4];
static long long *g_dev_stall[64];
static bool g_dev_stall_init[64];

int *device_error_flag(int slot) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || slot < 0 || slot >= 4) return nullptr;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev_flags[dev][slot]) {
    int *p = nullptr;
    if (cudaMalloc((void **)&p, sizeof(int)) != cudaSuccess) return nullptr;
    cudaMemset(p, 0, sizeof(int));
    g_dev_flags[dev][slot] = p;
  }
  return g_dev_flags[dev][slot];
}

long long *stall_profile_buffer() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_dev_mu);
  if (!g_dev_stall_init[dev]) {
    g_dev_stall_init[dev] = true;
    const char *e = getenv("PVCNN_STALL_PROFILE");
    long long *buf = nullptr;
    if (e && e[0] == '1' && cudaMalloc((void **)&buf, 8 * sizeof(long long)) == cudaSuccess)
      cudaMemset(buf, 0, 8 * sizeof(long long));
    else
      buf = nullptr;
    g_dev_stall[dev] = buf;
  }
  return g_dev_stall[dev];
}

}  // namespace pvb

namespace pvb {
int igemm_launch(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                 int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                 int npass, cudaStream_t stream);
int igemm_launch_ep(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                    int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                    int npass, cudaStream_t stream, const IgemmEpilogue *ep);
}
using namespace pvb;

extern "C" {

int pvcnn_split_tf32(long long n, const float *x, float *hi, float *lo, void *stream) {
  PVB_CHECK_ARG(n > 0 && (n % 4) == 0 && x && lo);
  const long long n4 = n / 4;
  PVB_LAUNCH(split_tf32_kernel, min((long long)ceil_div(n4, 256), (long long)kNumSMs * 16), 256, 0, stream, n4,
             reinterpret_cast<const float4 *>(x), reinterpret_cast<float4 *>(hi), reinterpret_cast<float4 *>(lo));
  return 0;
}

int pvcnn_conv_weight_prep(int cout, int cin, int ntaps, int mode, int ld, const float *w, float *w_hi, float *w_lo,
                           void *stream) {
  PVB_CHECK_ARG(cout > 0 && cin > 0 && (ntaps == 1 || ntaps == 27) && (mode == 0 || mode == 1) && w && w_hi && w_lo);
  PVB_CHECK_ARG(ld % 4 == 0 && ld >= (mode == 0 ? cin : cout));
  const long long total = (long long)ntaps * (mode == 0 ? cout : cin) * ld;
  PVB_LAUNCH(weight_prep_kernel, min((long long)ceil_div(total, 256), (long long)kNumSMs * 8), 256, 0, stream, cout,
             cin, ntaps, mode, ld, w, w_hi, w_lo);
  return 0;
}

int pvcnn_igemm_conv(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi,
                     const float *a_lo, int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias,
                     float *out, int ldo, int npass, void *stream) {
  return pvb::igemm_launch(nb, sx, sy, sz, k, cout, ntaps, a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, out, ldo, npass,
                           (cudaStream_t)stream);
}
}  // extern "C"

namespace pvb {
int igemm_launch(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                 int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                 int npass, cudaStream_t stream) {
  return igemm_launch_ep(nb, sx, sy, sz, k, cout, ntaps, a_hi, a_lo, lda, w_hi, w_lo, ldw, bias, out, ldo, npass, stream,
                         nullptr);
}

int igemm_launch_ep(int nb, int sx, int sy, int sz, int k, int cout, int ntaps, const float *a_hi, const float *a_lo,
                    int lda, const float *w_hi, const float *w_lo, int ldw, const float *bias, float *out, int ldo,
                    int npass, cudaStream_t stream, const IgemmEpilogue *ep) {
  PVB_CHECK_ARG(ep == nullptr || (ntaps == 1 && nb == 1 && sx == 1 && sy == 1));   // flat rows only
  PVB_CHECK_ARG(ep == nullptr || ((ep->scale == nullptr) == (ep->shift == nullptr)));
  PVB_CHECK_ARG(ep == nullptr || ep->group_bias == nullptr || (ep->group_rows > 0 && ep->group_ld >= cout));
  PVB_CHECK_ARG(nb > 0 && sx > 0 && sy > 0 && sz > 0 && k > 0 && cout > 0 && (ntaps == 1 || ntaps == 27));
  PVB_CHECK_ARG(a_hi && w_hi && out && (npass == 1 || npass == 3) && (npass == 1 || w_lo));
  PVB_CHECK_ARG(lda % 4 == 0 && ldw % 4 == 0 && ldo % 4 == 0 && lda >= k && ldw >= k && ldo >= cout);
  if (ntaps == 27) {
    // second-generation kernel (smem halo reuse, in-kernel lo) when the shape is inside its envelope
    const char *e_conv = getenv("PVCNN_B200_CONV");
    const bool force_v1 = e_conv && e_conv[0] == 'v' && e_conv[1] == '1';
    if (!force_v1) {
      const int rc = conv_halo_launch(nb, sx, sy, sz, k, cout, a_hi, lda, w_hi, w_lo, ldw, bias, out, ldo, npass, stream,
                                      nullptr, nullptr);
      if (rc != PVCNN_E_UNSUPPORTED) return rc;
    }
  }
  PVB_CHECK_ARG(npass == 1 || a_lo);  // the v1 kernel reads a materialised lo tensor
  int *g_err_flag = device_error_flag(0);
  PVB_CHECK_ARG(g_err_flag != nullptr);
  IgemmParams p{};
  p.nb = nb; p.sx = sx; p.sy = sy; p.sz = sz;
  // tile box: up to 128 voxels, z fastest
  p.bz = min(sz, IG_BLOCK_M);
  p.by = min(sy, max(1, IG_BLOCK_M / p.bz));
  p.bx = min(sx, max(1, IG_BLOCK_M / (p.bz * p.by)));
  if (sz > IG_BLOCK_M) { p.bz = IG_BLOCK_M; p.by = 1; p.bx = 1; }
  p.tz = ceil_div(sz, p.bz); p.ty = ceil_div(sy, p.by); p.tx = ceil_div(sx, p.bx);
  p.num_m_tiles = nb * p.tx * p.ty * p.tz;
  p.cin_chunks = ceil_div(k, IG_KC);
  p.ntaps = ntaps;
  p.cout = cout;
  int bn = ((cout + 15) / 16) * 16;
  if (bn > 128) bn = 128;
  p.block_n = bn;
  p.n_tiles = ceil_div(cout, bn);
  p.npass = npass;
  p.ldo = ldo;
  p.a_bytes = IG_A_TILE_BYTES * (npass > 1 ? 2 : 1);
  p.b_bytes = (uint32_t)bn * IG_KC * 4 * (npass > 1 ? 2 : 1);
  p.stage_bytes = p.a_bytes + p.b_bytes;  // multiple of 1024 since bn % 16 == 0 -> bn*128 % 2048 == 0? ensured below
  p.stage_bytes = (p.stage_bytes + 1023) & ~1023u;
  const uint32_t rows = (uint32_t)(p.bx * p.by * p.bz);
  p.tx_bytes = (rows * IG_KC * 4 + (uint32_t)bn * IG_KC * 4) * (npass > 1 ? 2 : 1);
  const int smem_budget = 227 * 1024 - 2048 - IG_EPI_BYTES;
  p.stages = min(IG_MAX_STAGES, smem_budget / (int)p.stage_bytes);
  PVB_CHECK_ARG(p.stages >= 2);
  // the tensor core accumulates with truncation (measured: ~2^-24 relative bias per accumulate), so
  // one accumulator sees at most ~IG_MAX_CHAIN MMAs; partial sums are combined in the epilogue
  const int chain = ntaps * p.cin_chunks * (IG_KC / 8);
  int split = ceil_div(chain, IG_MAX_CHAIN);
  const int corr = npass > 1 ? 1 : 0;
  while (split > 1 && (split + corr) * bn > 512) --split;
  p.acc_split = split;
  p.acc_slots = split + corr;
  p.acc_bufs = (2 * p.acc_slots * bn <= 512) ? 2 : 1;
  uint32_t cols = 32;
  while (cols < (uint32_t)(p.acc_bufs * p.acc_slots * bn)) cols <<= 1;
  p.tmem_cols = cols;
  p.bias = bias; p.out = out; p.err = g_err_flag;
  if (ep) p.ep = *ep;
  // B tiles must start 1024-aligned inside the stage: a_bytes is a multiple of 16 KB, b_hi = bn*128 bytes;
  // b_lo starts at b_bytes/2 = bn*128 which is a multiple of 1024 only when bn % 8 == 0 (true: bn % 16 == 0).

  CUtensorMap ma_hi, ma_lo, mw_hi, mw_lo;
  {
    cuuint64_t gdim[5] = {(cuuint64_t)k, (cuuint64_t)sz, (cuuint64_t)sy, (cuuint64_t)sx, (cuuint64_t)nb};
    cuuint64_t gstr[4] = {(cuuint64_t)lda * 4, (cuuint64_t)sz * lda * 4, (cuuint64_t)sy * sz * lda * 4,
                          (cuuint64_t)sx * sy * sz * lda * 4};
    cuuint32_t box[5] = {(cuuint32_t)IG_KC, (cuuint32_t)p.bz, (cuuint32_t)p.by, (cuuint32_t)p.bx, 1};
    int rc = encode_map(&ma_hi, a_hi, 5, gdim, gstr, box);
    if (rc) return rc;
    rc = encode_map(&ma_lo, npass > 1 ? a_lo : a_hi, 5, gdim, gstr, box);
    if (rc) return rc;
  }
  {
    cuuint64_t gdim[3] = {(cuuint64_t)k, (cuuint64_t)cout, (cuuint64_t)ntaps};
    cuuint64_t gstr[2] = {(cuuint64_t)ldw * 4, (cuuint64_t)cout * ldw * 4};
    cuuint32_t box[3] = {(cuuint32_t)IG_KC, (cuuint32_t)bn, 1};
    int rc = encode_map(&mw_hi, w_hi, 3, gdim, gstr, box);
    if (rc) return rc;
    rc = encode_map(&mw_lo, npass > 1 ? w_lo : w_hi, 3, gdim, gstr, box);
    if (rc) return rc;
  }
  const size_t smem = (size_t)p.stages * p.stage_bytes + 1024 + IG_EPI_BYTES;
  const int grid = min(kNumSMs, p.num_m_tiles * p.n_tiles);
  if (npass > 1) {
    PVB_CUDA(cudaFuncSetAttribute(igemm_conv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PVB_LAUNCH(igemm_conv_kernel<true>, grid, IG_THREADS, smem, stream, ma_hi, ma_lo, mw_hi, mw_lo, p);
  } else {
    PVB_CUDA(cudaFuncSetAttribute(igemm_conv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PVB_LAUNCH(igemm_conv_kernel<false>, grid, IG_THREADS, smem, stream, ma_hi, ma_lo, mw_hi, mw_lo, p);
  }
  return 0;
}

}  // namespace pvb

extern "C" {
int pvcnn_stall_profile_read(long long *host8) {
  long long *b = pvb::stall_profile_buffer();
  if (!b) return PVCNN_E_UNSUPPORTED;
  return (int)cudaMemcpy(host8, b, 8 * sizeof(long long), cudaMemcpyDeviceToHost);
}

/* Diagnostic: code of the mbarrier wait that starved (0 = none); readable after a trapped launch only
 * through a fresh context, so mainly useful under compute-sanitizer / in bring-up tests. */
int pvcnn_igemm_last_error(int *host_code) {
  int *g_err_flag = pvb::device_error_flag(0);
  if (!g_err_flag) { *host_code = 0; return 0; }
  return (int)cudaMemcpy(host_code, g_err_flag, sizeof(int), cudaMemcpyDeviceToHost);
}

}  // extern "C"
