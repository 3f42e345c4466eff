"""Build the UNMODIFIED reference CUDA extension into oracle/_ref/ (test infrastructure only).

The reference (mit-han-lab/pvcnn) ships its custom ops as a JIT torch extension
(/root/reference/modules/functional/backend.py:6-23).  It has no CPU path
(modules/functional/src/utils.hpp:7), so it can only *run* on the GPU box, but it
compiles here.  This script compiles the 13 reference source files *where they lie*
under /root/reference for sm_100a and drops the resulting shared object into
oracle/_ref/ (git-ignored, but shipped to the GPU box by gpurun) so that `-m gpu`
parity tests can compare our kernels with the reference's own kernels on the same
inputs.  No reference source is copied into the repository.

Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may load the result.
"""
import os
import shutil
import sys

REF = os.environ.get("PVCNN_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
SRCS = [
    "ball_query/ball_query.cpp", "ball_query/ball_query.cu",
    "grouping/grouping.cpp", "grouping/grouping.cu",
    "interpolate/neighbor_interpolate.cpp", "interpolate/neighbor_interpolate.cu",
    "interpolate/trilinear_devox.cpp", "interpolate/trilinear_devox.cu",
    "sampling/sampling.cpp", "sampling/sampling.cu",
    "voxelization/vox.cpp", "voxelization/vox.cu",
    "bindings.cpp",
]


def build(verbose=False):
    src_dir = os.path.join(REF, "modules", "functional", "src")
    if not os.path.isdir(src_dir):
        return None
    os.makedirs(OUT, exist_ok=True)
    target = os.path.join(OUT, "_pvcnn_backend.so")
    if os.path.exists(target):
        return target
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    from torch.utils.cpp_extension import load
    build_dir = os.path.join("/tmp", "pvcnn_ref_build")
    os.makedirs(build_dir, exist_ok=True)
    load(name="_pvcnn_backend", extra_cflags=["-O3", "-std=c++17"],
         sources=[os.path.join(src_dir, s) for s in SRCS],
         build_directory=build_dir, verbose=verbose, is_python_module=False)
    shutil.copy(os.path.join(build_dir, "_pvcnn_backend.so"), target)
    return target


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv))
