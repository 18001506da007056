"""Build recipe for the native pieces (nvcc / g++), all in-tree so the .so files travel with the repo.

    librpw_b200.so        csrc/rpw_kernels.cu + csrc/rpw_capi.cu   sm_100a only, -fmad=false, -lineinfo
    synth/libscangen.so   synth/scangen.cpp                         host-side synthetic scan generators

Nothing here falls back to a CPU implementation: if nvcc is missing the build fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent
LIB = PKG / "librpw_b200.so"
SCANGEN = PKG / "synth" / "libscangen.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",  # every float op rounds once, like the strict CPU reference; FMAs are explicit fmaf()
    "-Xcompiler", "-fPIC",
]


def _newer(target: Path, sources) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(Path(s).stat().st_mtime <= t for s in sources)


def build_cuda(force: bool = False, verbose: bool = False) -> Path:
    srcs = [PKG / "csrc" / "rpw_kernels.cu", PKG / "csrc" / "rpw_capi.cu"]
    deps = srcs + [PKG / "csrc" / "rpw_kernels.h", PKG / "csrc" / "rpw_device.cuh", ROOT / "include" / "rpw_b200.h", Path(__file__)]
    if not force and _newer(LIB, deps):
        return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built (there is no CPU fallback)")
    cmd = [nvcc, *NVCC_FLAGS, f"-I{ROOT / 'include'}", f"-I{PKG / 'csrc'}", "-shared", "-o", str(LIB), *map(str, srcs)]
    if verbose:
        cmd[1:1] = ["-Xptxas", "-v"]
    subprocess.run(cmd, check=True)
    return LIB


def build_scangen(force: bool = False) -> Path:
    src = PKG / "synth" / "scangen.cpp"
    if not force and _newer(SCANGEN, [src]):
        return SCANGEN
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(SCANGEN), str(src)], check=True)
    return SCANGEN


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_scangen(force)
