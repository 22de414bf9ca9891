"""Build driver: nvcc (sm_100a only) for libmnemo_cuda.so, gcc for the C host library and CLI.

    python -m mnemophonix_b200.build [--force]

Outputs land in-tree under mnemophonix_b200/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
OBJ = os.path.join(PKG, "_obj")
LIB_CUDA = os.path.join(PKG, "libmnemo_cuda.so")
LIB_HOST = os.path.join(PKG, "libmnemophonix.so")
CLI = os.path.join(PKG, "mnemophonix")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # exact-arithmetic contract (csrc/common.cuh): no FMA contraction, IEEE div/sqrt, denormals kept
    "-fmad=false", "-ftz=false", "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC,-O2,-fno-fast-math,-ffp-contract=off", "-Xptxas", "-v",
]


def nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libmnemo_cuda.so cannot be built (there is no CPU fallback)")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_cuda(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    cus = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs += [os.path.join(ROOT, "include", "mnemo_cuda.h"), os.path.join(PKG, "data", "permutations_table.inc")]
    cc = nvcc()
    jobs = []
    for f in cus:
        src = os.path.join(CSRC, f)
        obj = os.path.join(OBJ, f[:-3] + ".o")
        if force or _newer(obj, [src] + hdrs):
            jobs.append((src, obj))

    def one(job):
        src, obj = job
        r = subprocess.run([cc, *NVCC_FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        return src, r

    with cf.ThreadPoolExecutor(max_workers=8) as ex:
        for src, r in ex.map(one, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
            log = os.path.join(OBJ, os.path.basename(src) + ".ptxas.txt")
            with open(log, "w") as fh:
                fh.write(r.stderr)
    objs = [os.path.join(OBJ, f[:-3] + ".o") for f in cus]
    if force or jobs or _newer(LIB_CUDA, objs):
        subprocess.check_call([cc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB_CUDA, *objs,
                               "-cudart", "static"])
    return LIB_CUDA


def build_host(force: bool = False) -> tuple[str, str] | None:
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".c") and f != "main.c") \
        if os.path.isdir(HOST) else []
    if not srcs:
        return None
    inc = ["-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "include", "mnemophonix")]
    deps = srcs + [LIB_CUDA] + [os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(ROOT, "include")) for f in fs]
    if force or _newer(LIB_HOST, deps):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", *inc, *srcs, "-o", LIB_HOST,
                               "-L", PKG, "-lmnemo_cuda", "-Wl,-rpath,$ORIGIN", "-lm", "-lpthread"])
    main_c = os.path.join(HOST, "main.c")
    if os.path.exists(main_c) and (force or _newer(CLI, [main_c, LIB_HOST])):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-Wextra", *inc, main_c, "-o", CLI, "-L", PKG, "-lmnemophonix",
                               "-lmnemo_cuda", "-Wl,-rpath,$ORIGIN", "-lm"])
    return LIB_HOST, CLI


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_cuda(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print("built", LIB_CUDA)
