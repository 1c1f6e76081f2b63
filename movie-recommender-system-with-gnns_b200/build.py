"""In-tree build of csrc/liblgcn_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python -m lgcn_b200.build            (or: __graft_entry__.build())

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
REPO = os.path.dirname(HERE)
INCLUDE = os.path.join(REPO, "include")
LIB = os.path.join(CSRC, "liblgcn_b200.so")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")
NVCC = os.path.join(CUDA_HOME, "bin", "nvcc")
METIS_A = os.path.join(CUDA_HOME, "targets", "x86_64-linux", "lib", "libmetis_static.a")

CU_SOURCES = ["api.cu", "graph_build.cu", "propagate.cu", "bpr.cu", "adam.cu", "cluster.cu", "score_topk.cu", "rows.cu", "sparse_step.cu", "score_topk_tc.cu", "graph_build_batched.cu", "epoch_kernel.cu", "probe.cu", "bpr_owner.cu", "partition_gpu.cu"]
C_SOURCES = ["partition_metis.c"]
HEADERS = ["common.cuh", "rowtask.cuh", "adam.cuh"]

NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-I", INCLUDE, "-I", CSRC]


def _digest() -> str:
    h = hashlib.sha256()
    for name in CU_SOURCES + C_SOURCES + HEADERS:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    with open(os.path.join(INCLUDE, "lgcn_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _run(cmd, log):
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log.append("$ " + " ".join(cmd) + "\n" + p.stdout)
    if p.returncode != 0:
        raise RuntimeError("build failed:\n" + "\n".join(log[-1:]))


def build_variant(tag: str, defines: dict) -> str:
    """Tuning aid: a differently parameterised copy of the library, csrc/build/variants/liblgcn_<tag>.so
    (select it with LGCN_LIB_PATH).  Not used by the product path."""
    vdir = os.path.join(CSRC, "build", "variants", tag)
    os.makedirs(vdir, exist_ok=True)
    lib = os.path.join(CSRC, "build", "variants", f"liblgcn_{tag}.so")
    flags = [f for f in NVCC_FLAGS if f not in ("-Xptxas", "-v")] + [f"-D{k}={v}" for k, v in defines.items()]
    log: list = []

    def cc(name):
        obj = os.path.join(vdir, name + ".o")
        _run([NVCC, *flags, "-c", os.path.join(CSRC, name), "-o", obj], log)
        return obj
    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(cc, CU_SOURCES))
    obj = os.path.join(vdir, "partition_metis.c.o")
    _run(["gcc", "-O2", "-fPIC", "-I", INCLUDE, "-c", os.path.join(CSRC, "partition_metis.c"), "-o", obj], log)
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib, *objs, obj, METIS_A, "-lcudart", "-lm"], log)
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(CSRC, "build", "digest.txt")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    os.makedirs(os.path.join(CSRC, "build"), exist_ok=True)
    log: list = []
    objs = []

    def compile_cu(name):
        obj = os.path.join(CSRC, "build", name + ".o")
        _run([NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, name), "-o", obj], log)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(CU_SOURCES))) as ex:
        objs = list(ex.map(compile_cu, CU_SOURCES))
    for name in C_SOURCES:
        obj = os.path.join(CSRC, "build", name + ".o")
        _run(["gcc", "-O2", "-fPIC", "-I", INCLUDE, "-c", os.path.join(CSRC, name), "-o", obj], log)
        objs.append(obj)
    _run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, METIS_A,
          "-lcudart", "-lm"], log)
    with open(os.path.join(CSRC, "build", "build.log"), "w") as f:
        f.write("\n".join(log))
    with open(stamp, "w") as f:
        f.write(dig)
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
