"""In-tree build of libcapdec.so with nvcc for sm_100a (B200).

The library is built next to this file (git-ignored, but shipped to the GPU box by gpurun).  nvcc
cross-compiles without a GPU, so `build()` works in the authoring container too.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libcapdec.so")
OBJ_DIR = os.path.join(HERE, "build")
SOURCES = ["gemm_tc.cu", "gemm_simt.cu", "pointwise.cu", "attention.cu", "loss.cu", "recurrent.cu", "adaptor.cu", "decoder.cu"]
HEADERS = ["common.cuh", "gemm.cuh", "kernels.cuh", os.path.join("..", "..", "include", "capdec.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libcapdec.so cannot be built (there is no CPU fallback)")


def _newest(paths) -> float:
    return max(os.path.getmtime(p) for p in paths)


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    srcs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    srcs.append(os.path.abspath(__file__))
    return _newest(srcs) > os.path.getmtime(LIB)


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every .cu for sm_100a and link libcapdec.so.  Returns the library path."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    os.makedirs(OBJ_DIR, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("CAPDEC_NVCC_EXTRA", "").split() + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [nvcc, "-shared", "-o", tmp] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"]
    # -lcuda is not needed: cuTensorMapEncodeTiled is resolved at run time through cudaGetDriverEntryPoint
    cmd = [c for c in cmd if c != "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
