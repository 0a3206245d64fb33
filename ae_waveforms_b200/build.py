"""Build libaewave.so (hand-written CUDA for sm_100a) in-tree with nvcc.

``python -m ae_waveforms_b200.build`` or ``__graft_entry__.build()``.  nvcc
cross-compiles without a GPU.  The reductions that must match NumPy bit for bit
(k1, k3) are compiled with ``-fmad=false``; the filter (k2) keeps FMA contraction
(its exact mode uses explicit round-to-nearest intrinsics).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libaewave.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
PER_FILE = {
    "k1_zscore.cu": ["-fmad=false"],
    "k3_windows.cu": ["-fmad=false"],
    "k5_recon.cu": [],
    "k6_features.cu": ["-fmad=false"],
    "k7_delineate.cu": [],
    "k8_convert.cu": ["-fmad=false"],
    "k9_metrics.cu": ["-fmad=false"],
    "k2_filtfilt.cu": [],
    "cabi_misc.cu": [],
}


def _newer(src: str, out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(p) > t for p in [src, *deps])


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; libaewave.so cannot be built")
    obj_dir = os.path.join(CSRC, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(HERE, "..", "include", "aewave.h"))
    objs = []
    for name in sorted(os.listdir(CSRC)):
        src = os.path.join(CSRC, name)
        if name.endswith(".cu"):
            obj = os.path.join(obj_dir, name[:-3] + ".o")
            cmd = [nvcc, *ARCH, *COMMON, *PER_FILE.get(name, []), "-c", src, "-o", obj]
            if verbose:
                cmd[1:1] = ["-Xptxas", "-v"]
        elif name.endswith(".cpp"):
            obj = os.path.join(obj_dir, name[:-4] + ".o")
            cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-c", src, "-o", obj]
        else:
            continue
        objs.append(obj)
        if force or _newer(src, obj, headers):
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"compiling {name} failed")
    if force or any(_newer(o, LIB, []) for o in objs):
        r = subprocess.run([nvcc, *ARCH, "-shared", "-o", LIB, *objs], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("linking libaewave.so failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
