"""Build ``libdvsloss.so`` (C ABI of include/dvsloss.h) for sm_100a with nvcc, in-tree.

    python deep-visual-slam_b200/csrc/build.py [--force] [--verbose]

Output: ``deep-visual-slam_b200/dvsloss/libdvsloss.so`` (git-ignored; shipped to the GPU box by gpurun).
nvcc cross-compiles without a GPU.  Objects are cached under ``csrc/_build`` and rebuilt when a source
or header is newer.
"""
from __future__ import annotations

import argparse
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)
OUT = os.path.join(PKG, "dvsloss", "libdvsloss.so")
OBJ_DIR = os.path.join(HERE, "_build")
SOURCES = ["dvs_api.cu", "dvs_fused.cu", "dvs_pair.cu", "dvs_ops.cu", "dvs_head.cu"]
HEADERS = ["dvs_fused_core.cuh", "dvs_pair_core.cuh", "dvs_host.h", os.path.join(ROOT, "include", "dvsloss.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found; the CUDA toolkit is required to build libdvsloss.so")
    return exe


def _newest_dep() -> float:
    deps = [os.path.join(HERE, h) if not os.path.isabs(h) else h for h in HEADERS] + [os.path.abspath(__file__)]
    return max(os.path.getmtime(d) for d in deps)


def build(force: bool = False, verbose: bool = False, variant: str = "", defines=()) -> str:
    """``variant`` / ``defines``: experiment builds (``libdvsloss_<variant>.so`` with extra -D macros, selected at run time
    with DVSLOSS_LIB=...); the default build has neither."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    dep_t = _newest_dep()
    extra = (["-Xptxas", "-v"] if verbose else []) + [f"-D{d}" for d in defines]
    out = OUT if not variant else OUT.replace(".so", f"_{variant}.so")
    jobs = []
    objs = []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", (f"_{variant}" if variant else "") + ".o"))
        objs.append(o)
        if force or not os.path.exists(o) or os.path.getmtime(o) < max(os.path.getmtime(s), dep_t):
            jobs.append([nvcc(), *NVCC_FLAGS, *extra, "-c", s, "-o", o])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=max(1, len(jobs))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {cmd[-3]}")
    if jobs or force or not os.path.exists(out):
        cmd = [nvcc(), "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--variant", default="")
    ap.add_argument("-D", dest="defines", action="append", default=[])
    a = ap.parse_args()
    print(build(a.force, a.verbose, a.variant, a.defines))
