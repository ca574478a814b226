"""Build libpromptcache_b200.so (sm_100a only) in-tree with nvcc.

`python -m promptcache_b200.build` or `build()`; nvcc cross-compiles without a GPU.  The .so is
git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libpromptcache_b200.so")
STAMP = os.path.join(HERE, ".libpromptcache_b200.stamp")
SOURCES = ["api.cu", "splice.cu", "rope_append.cu", "attn_simt.cu", "attn_combine.cu", "attn_tc_small.cu", "attn_tc2.cu", "support.cu", "linear_small.cu", "linear_chain.cu"]
HEADERS = [os.path.join(CSRC, "pc_common.cuh"), os.path.join(CSRC, "linear_common.cuh"), os.path.join(HERE, "..", "include", "promptcache_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; promptcache_b200 needs the CUDA 12.9 toolchain to build")


def _digest() -> str:
    h = hashlib.sha256()
    for f in [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    extra = os.environ.get("PC_NVCC_EXTRA", "").split()  # experiments only (e.g. -DPC_EXP_...); empty in normal builds
    digest = _digest() + "|" + " ".join(extra)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for s in SOURCES:
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = []
    for s, pr in procs:
        out, _ = pr.communicate()
        if verbose or pr.returncode != 0:
            sys.stderr.write(out)
        if pr.returncode != 0:
            failed.append(s)
    if failed:
        raise RuntimeError(f"nvcc failed for {failed}")
    link = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart_static",
            "-Xcompiler", "-fPIC"]
    subprocess.run(link, check=True)
    # a shared library links with unresolved symbols; make them a build error (dlopen resolves everything eagerly)
    import ctypes
    ctypes.CDLL(LIB, mode=ctypes.RTLD_LOCAL | os.RTLD_NOW)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
