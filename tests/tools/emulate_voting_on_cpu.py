"""Development aid (build container has no GPU): executes the SOURCE of pvcnn_b200/csrc/eval_voting.cu on the CPU.

The voting kernels are plain integer / byte kernels (global atomics, one __syncthreads, no warp intrinsics), so a small
shim -- CUDA qualifiers defined away, blockIdx / threadIdx as thread-locals, atomics as GCC builtins, `PVB_LAUNCH`
looping over the grid (real threads + a barrier per block for the kernel that synchronises) -- lets g++ compile the
unmodified file, launchers included.  This script builds that emulation library under /tmp, points `_lib.call` at it
and runs tests/test_voting_gpu.py with CPU tensors.  It validates kernel logic, launch geometry, argument order of the
ctypes calls and the host code; it says nothing about performance and it is NOT a product path (nothing in pvcnn_b200/
refers to it; the product library still has no CPU fallback).

    python tests/tools/emulate_voting_on_cpu.py
"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SRC = os.path.join(ROOT, "pvcnn_b200", "csrc", "eval_voting.cu")
OUT = "/tmp/pvcnn_emulate_voting"

SHIM = r"""
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <functional>
#include <thread>
#include <vector>
#include <pthread.h>
#include "%(root)s/include/pvcnn_b200.h"
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(x)
#define __shared__
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
static thread_local dim3 blockIdx, threadIdx;
static dim3 blockDim, gridDim;
static pthread_barrier_t g_bar;
static inline void __syncthreads() { pthread_barrier_wait(&g_bar); }
using std::min; using std::max;
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned long long atomicMax(unsigned long long *p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
static inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
namespace pvb {
constexpr int kNumSMs = 148;
static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
unsigned int vt_hist[1 << 16];     // the dynamic shared memory of the one kernel that uses it (blocks run one at a time)
}
#define PVB_CHECK_ARG(cond) do { if (!(cond)) return PVCNN_E_BADARG; } while (0)
static void emu_launch(const char *name, dim3 grid, dim3 block, const std::function<void()> &body) {
  const bool threads = strstr(name, "vote_stats") != nullptr;   // uses __syncthreads: needs concurrent threads
  gridDim = grid; blockDim = block;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      if (!threads) {
        for (unsigned t = 0; t < block.x; ++t) { blockIdx = dim3(bx, by); threadIdx = dim3(t); body(); }
      } else {
        pthread_barrier_init(&g_bar, nullptr, block.x);
        std::vector<std::thread> ts;
        for (unsigned t = 0; t < block.x; ++t)
          ts.emplace_back([&, t] { blockIdx = dim3(bx, by); threadIdx = dim3(t); body(); });
        for (auto &t : ts) t.join();
        pthread_barrier_destroy(&g_bar);
      }
    }
}
#define PVB_LAUNCH(kernel, grid, block, smem, stream, ...) \
  emu_launch(#kernel, dim3(grid), dim3(block), [&] { kernel(__VA_ARGS__); })
"""


def build():
    os.makedirs(OUT, exist_ok=True)
    text = open(SRC).read().replace('#include "common.cuh"', "")
    cpp = os.path.join(OUT, "eval_voting_emulated.cpp")
    with open(cpp, "w") as f:
        f.write(SHIM % {"root": ROOT})
        f.write(text)
    lib = os.path.join(OUT, "libemulated.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-pthread", "-fvisibility=default",
                    "-ffp-contract=off", cpp, "-o", lib], check=True)
    return lib


def main():
    lib = ctypes.CDLL(build())
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    from pvcnn_b200 import _lib, evaluate

    def call(name, *args, device=None):
        cargs = []
        for a in args:
            if isinstance(a, torch.Tensor):
                assert a.is_contiguous(), name
                cargs.append(ctypes.c_void_p(a.data_ptr()))
            elif a is None:
                cargs.append(ctypes.c_void_p(0))
            elif isinstance(a, float):
                cargs.append(ctypes.c_float(a))
            elif isinstance(a, (bool, int)):
                cargs.append(ctypes.c_int(int(a)))
            else:
                cargs.append(a)
        rc = getattr(lib, name)(*cargs, ctypes.c_void_p(0))
        if rc != 0:
            raise _lib.PvcnnError("%s failed with code %d" % (name, rc))

    _lib.call = call
    evaluate._DEFAULT_DEVICE = "cpu"
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    torch.cuda.is_available = lambda: True
    _zeros = torch.zeros
    torch.zeros = lambda *a, **k: _zeros(*a, **{**k, "device": "cpu"} if "device" in k else k)
    import pytest
    return pytest.main([os.path.join(ROOT, "tests", "test_voting_gpu.py"), "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider"])


if __name__ == "__main__":
    sys.exit(main())
