"""Build libwiw_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwiw_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# Development variants (never the product library): WIW_BUILD_VARIANT=phases adds -DWIW_GEMM_PHASES (per-phase epilogue cycle
# counters) and writes libwiw_b200_phases.so + build_phases/; select it at run time with WIW_B200_LIB.
VARIANT = os.environ.get("WIW_BUILD_VARIANT", "")
VARIANT_DEFS = {"phases": ["-DWIW_GEMM_PHASES"]}
if VARIANT:
    FLAGS = FLAGS + VARIANT_DEFS[VARIANT]
    LIB = os.path.join(HERE, f"libwiw_b200_{VARIANT}.so")
BUILD_DIR = os.path.join(HERE, "build" + (f"_{VARIANT}" if VARIANT else ""))


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + [os.path.join(HERE, "..", "include", "wiw.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    objs = []
    os.makedirs(BUILD_DIR, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(BUILD_DIR, os.path.basename(src).replace(".cu", ".o"))
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src)
                and all(os.path.getmtime(obj) > os.path.getmtime(h) for h in glob.glob(os.path.join(CSRC, "*.cuh")))
                and os.path.getmtime(obj) > os.path.getmtime(os.path.join(HERE, "..", "include", "wiw.h"))):
            continue
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        log = os.path.join(BUILD_DIR, os.path.basename(src) + ".ptxas.log")
        with open(log, "w") as f:
            f.write(out)
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
