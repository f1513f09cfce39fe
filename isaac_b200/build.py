"""Builds libhectorb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m isaac_b200.build [--force] [--verbose]

The env/GAE translation units are compiled with --fmad=false: the reference evaluates its
expressions as separate fp32 torch ops, and parity at 1e-5 on exp(-100*x) terms does not
survive contracted multiply-adds (SURVEY.md §7 hard part 4).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libhectorb200.so")
STAMP = os.path.join(PKG, "csrc", ".build_stamp")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", CSRC]
# translation unit -> extra flags
UNITS = {
    "api.cu": [],
    "env_kernels.cu": ["--fmad=false"],
    "gae_kernels.cu": ["--fmad=false"],
    "ppo_kernels.cu": [],
    "mlp_gemm.cu": [],
}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libhectorb200.so cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(ROOT, "include")):
        for name in sorted(os.listdir(root)):
            if name.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, name), "rb") as f:
                    h.update(name.encode() + b"\0" + f.read())
    h.update(repr(sorted(UNITS.items())).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    extra_all = os.environ.get("HB_EXTRA_NVCC_FLAGS", "").split()      # debug builds, e.g. -DHB_POST_TIMING
    digest = _digest() + "".join(extra_all)
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for unit, extra in UNITS.items():
        obj = os.path.join(CSRC, unit.replace(".cu", ".o"))
        cmd = [nvcc, *ARCH, *COMMON, *extra, *extra_all, "-c", os.path.join(CSRC, unit), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((unit, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for unit, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0 or verbose:
            print(f"--- {unit} ---\n{out}", flush=True)
        failed |= pr.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-lcudart"]
    subprocess.run(cmd, check=True)
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
