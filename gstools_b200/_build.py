"""Compile the sm_100a CUDA library in-tree (``gstools_b200/libgsb200.so``).

nvcc cross-compiles without a GPU.  The shared object is git-ignored but travels to
the GPU box with the repository snapshot.
"""

from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("GSB200_LIB") or os.path.join(HERE, "libgsb200.so")
SOURCES = ["gsb_api.cu"]
HEADERS = ["gsb_common.cuh", "gsb_direct.cuh", "gsb_separable.cuh", "gsb_sepk.cuh", "gsb_krige.cuh", "gsb_sampler.cuh",
           "sincos_coeffs.cuh"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off", "-shared",
    "-diag-suppress", "550",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build gstools_b200/libgsb200.so")


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    lib_m = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "gsb200.h"))
    return any(os.path.getmtime(d) > lib_m for d in deps)


def build(force: bool = False, verbose: bool = False, defines=(), out=None) -> str:
    """Build ``libgsb200.so`` if it is missing or older than its sources.

    ``defines`` / ``out`` build a tuning variant (e.g. ``defines=["GSB_SEP_KC=16"]``) next to it.
    """
    target = out or LIB
    if not force and out is None and not needs_build():
        return LIB
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-o", target + ".tmp",
           *[os.path.join(CSRC, s) for s in SOURCES]]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    env = dict(os.environ)
    # some images export CC/CXX wrappers that break nvcc's host compile; use the system g++
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, env=env)
    os.replace(target + ".tmp", target)
    return target


if __name__ == "__main__":
    print(build(force=True, verbose=True))
