"""Build libgsb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

Used by ``__graft_entry__.build()`` and ``python -m gscodec_studio_b200._build``.  nvcc cross-compiles
without a GPU; the resulting shared object travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libgsb200.so")
SOURCES = ["abi.cu", "projection.cu", "sh.cu", "isect.cu", "isect_ordered.cu", "blend.cu", "blend_stream.cu", "quant.cu", "optim.cu", "loss.cu", "geom_ops.cu", "dynamic.cu", "indices.cu", "hashgrid.cu", "surfel.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--use_fast_math", "-std=c++17",
    "--extended-lambda", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libgsb200.so cannot be built")


def _deps_mtime() -> float:
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    files.append(os.path.join(HERE, "..", "include", "gsb200.h"))
    return max(os.path.getmtime(f) for f in files)


EXTRA_FLAGS = [f for f in os.environ.get("GSB200_NVCC_EXTRA", "").split() if f]
# quant.cu is compiled WITHOUT --use_fast_math: its fused exp / sigmoid must return torch.exp / torch.sigmoid's bits (libdevice
# expf, IEEE division), so that the activated scales / opacities -- hence conics, tile lists and every alpha >= 1/255
# decision -- equal the reference trainer's from the RAW parameters; the entropy-model kernels in the same file spell their
# fast intrinsics out (__expf) and are unaffected
NO_FAST_MATH = {"quant.cu"}


def _compile(src: str) -> str:
    obj = os.path.join(OBJ, src.replace(".cu", ".o"))
    log = obj + ".log"
    flags = [f for f in NVCC_FLAGS if not (src in NO_FAST_MATH and f == "--use_fast_math")]
    cmd = [_nvcc(), *flags, *EXTRA_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
    with open(log, "w") as fh:
        rc = subprocess.call(cmd, stdout=fh, stderr=subprocess.STDOUT)
    if rc != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n" + open(log).read()[-4000:])
    return obj


def build_library(force: bool = False, verbose: bool = False) -> str:
    global LIB, OBJ
    if os.environ.get("GSB200_LIB_OUT"):  # A/B variant build: separate output and object directory
        LIB = os.environ["GSB200_LIB_OUT"]
        OBJ = LIB + "_obj"
        force = True
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _deps_mtime():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
