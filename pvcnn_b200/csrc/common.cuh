// common.cuh -- shared helpers for the sm_100a kernels of pvcnn_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/pvcnn_b200.h"

namespace pvb {

// Global launch counter (exported through pvcnn_launch_count()).
extern std::atomic<unsigned long long> g_launches;

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

#define PVB_CHECK_ARG(cond) \
  do {                      \
    if (!(cond)) return PVCNN_E_BADARG; \
  } while (0)

// Every launch goes through this macro: counts it and surfaces launch errors as return codes.
#define PVB_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
  do {                                                                          \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__);   \
    ++pvb::g_launches;                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) return (int)e__;                                    \
  } while (0)

#define PVB_CUDA(call)                          \
  do {                                          \
    cudaError_t e__ = (call);                   \
    if (e__ != cudaSuccess) return (int)e__;    \
  } while (0)

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// d^2 exactly as the reference's SASS computes it (FMUL dy, FFMA dx, FFMA dz):
//   ball_query.cu:34-37, sampling.cu:136-137, neighbor_interpolate.cu:42
__device__ __forceinline__ float sqdist(float dx, float dy, float dz) {
  return __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
}

// 128-bit streaming accesses
__device__ __forceinline__ float4 ldg_stream4(const float *p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void stg_stream4(float *p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
// vectorised reduction (sm_90+): one 16-byte red instead of four 4-byte atomics
__device__ __forceinline__ void red_add4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add2(float *p, float2 v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}

// Optional fused epilogue of igemm_conv_kernel (ntaps == 1 GEMMs).  Per output element, in this order:
//   v = acc + bias[c]  (+ group_bias[(row / group_rows) * group_ld + c])      -- a per-cloud bias: the contribution of
//                                                                                channels that are constant over a cloud
//   v = fmaf(v, scale[c], shift[c]); v = v > 0 ? v : v * slope                -- BatchNorm (given coefficients) + (Leaky)ReLU,
//                                                                                the arithmetic of bn_apply_leaky_kernel
//   out[row, c] = v;  out_lo[row, c] = v - trunc_tf32(v)                      -- lo operand of the next 3xTF32 GEMM
struct IgemmEpilogue {
  const float *scale = nullptr, *shift = nullptr;
  float slope = 0.0f;
  float *out_lo = nullptr;
  const float *group_bias = nullptr;
  int group_rows = 0, group_ld = 0;
};
}  // namespace pvb
