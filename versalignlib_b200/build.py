"""In-tree builds (explicit nvcc / g++, no JIT cache): the shared objects land next to
the sources under versalignlib_b200/lib/ so they travel to the GPU box with the snapshot.

  libCUDAKernel.so  the plug-in: CUDA kernels + C ABI (include/versalign_cuda.h) +
                    the four dlsym entry points of the reference's plug-in boundary
  libva_host.so     driver-side loader (plugin_host.cpp), C ABI for ctypes
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIBDIR = os.path.join(PKG, "lib")
INCLUDE = os.path.join(ROOT, "include")

CUDA_PLUGIN = os.path.join(LIBDIR, "libCUDAKernel.so")
HOST_LIB = os.path.join(LIBDIR, "libva_host.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
HOST_CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-O2,-Wall,-pthread", "--use_fast_math",
    "-I", INCLUDE, "-I", CSRC,
]

CUDA_SOURCES = ["va_prep.cu", "va_kernels.cu", "va_fast.cu", "va_nw.cu", "va_intra.cu", "va_traceback.cu", "va_cabi.cu", "va_fasta.cpp", "cuda_kernel_plugin.cpp"]
CUDA_HEADERS = ["va_device.cuh", "va_fast.cuh", "va_internal.h", "va_line_streamer.h"]


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str], verbose: bool) -> None:
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if verbose and r.stderr:
        print(r.stderr, file=sys.stderr)


def build_host(force: bool = False, verbose: bool = False) -> str:
    src = os.path.join(CSRC, "plugin_host.cpp")
    deps = [src, os.path.join(INCLUDE, "versalign_plugin_abi.h")]
    if force or _newer(HOST_LIB, deps):
        os.makedirs(LIBDIR, exist_ok=True)
        _run([HOST_CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I", INCLUDE, src, "-o", HOST_LIB, "-ldl"],
             verbose)
    return HOST_LIB


def build_cuda(force: bool = False, verbose: bool = False, extra: list[str] | None = None) -> str:
    srcs = [os.path.join(CSRC, s) for s in CUDA_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in CUDA_HEADERS] + [
        os.path.join(INCLUDE, "versalign_cuda.h"), os.path.join(INCLUDE, "versalign_fasta.h"),
        os.path.join(INCLUDE, "versalign_plugin_abi.h")]
    if force or _newer(CUDA_PLUGIN, deps):
        os.makedirs(LIBDIR, exist_ok=True)
        objs = []
        for s in srcs:
            o = os.path.join(LIBDIR, os.path.basename(s) + ".o")
            if force or _newer(o, deps):
                _run([NVCC, *NVCC_FLAGS, *(extra or []), "-x", "cu", "-c", s, "-o", o], verbose)
            objs.append(o)
        _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", HOST_CXX,
              "-Xcompiler", "-pthread", "-o", CUDA_PLUGIN, *objs, "-cudart", "static", "-lpthread", "-ldl"], verbose)
    return CUDA_PLUGIN


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_host(force, verbose)
    build_cuda(force, verbose)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print(CUDA_PLUGIN)
    print(HOST_LIB)
