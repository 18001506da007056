"""Build recipe for the native pieces (nvcc / g++), all in-tree so the .so files travel with the repo.

    librpw_b200.so        csrc/rpw_kernels.cu + rpw_fit_fast.cu + rpw_fit_replay.cu + rpw_capi.cu   sm_100a only,
                          -fmad=false, -lineinfo; the four units compile side by side, then link
    synth/libscangen.so   synth/scangen.cpp                         host-side synthetic scan generators

Nothing here falls back to a CPU implementation: if nvcc is missing the build fails loudly.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
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


CUDA_UNITS = ["rpw_kernels.cu", "rpw_fit_fast.cu", "rpw_fit_replay.cu", "rpw_capi.cu"]


def build_cuda(force: bool = False, verbose: bool = False, out: Path | None = None, extra_flags=()) -> Path:
    """nvcc -c every unit (in parallel), then nvcc -shared.  `out` / `extra_flags`: kernel-experiment variants."""
    target = Path(out) if out else LIB
    srcs = [PKG / "csrc" / u for u in CUDA_UNITS]
    deps = srcs + [PKG / "csrc" / "rpw_kernels.h", PKG / "csrc" / "rpw_device.cuh", PKG / "csrc" / "rpw_fit.cuh",
                   ROOT / "include" / "rpw_b200.h", Path(__file__)]
    if not force and not extra_flags and _newer(target, deps):
        return target
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: the sm_100a library cannot be built (there is no CPU fallback)")
    objdir = PKG / "build" / target.stem
    objdir.mkdir(parents=True, exist_ok=True)
    base = [nvcc, *NVCC_FLAGS, *extra_flags, f"-I{ROOT / 'include'}", f"-I{PKG / 'csrc'}"]
    if verbose:
        base[1:1] = ["-Xptxas", "-v"]
    procs = []
    for src in srcs:
        obj = objdir / (src.stem + ".o")
        procs.append((src, obj, subprocess.Popen([*base, "-c", "-o", str(obj), str(src)])))
    for src, obj, pr in procs:
        if pr.wait() != 0:
            raise RuntimeError(f"nvcc failed on {src}")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", str(target), *[str(o) for _, o, _ in procs]], check=True)
    return target


def build_scangen(force: bool = False) -> Path:
    src = PKG / "synth" / "scangen.cpp"
    if not force and _newer(SCANGEN, [src]):
        return SCANGEN
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-o", str(SCANGEN), str(src)], check=True)
    return SCANGEN


DROPIN_LATENCY = ROOT / "tools" / "_bin" / "dropin_latency"


def build_tools(force: bool = False) -> Path:
    """bench.py's C++ drop-in latency probe (tools/dropin_latency.cpp against host/recursive_patchwork.hpp + the library)."""
    src = ROOT / "tools" / "dropin_latency.cpp"
    if not force and _newer(DROPIN_LATENCY, [src, PKG / "host" / "recursive_patchwork.hpp", ROOT / "include" / "rpw_b200.h", LIB]):
        return DROPIN_LATENCY
    DROPIN_LATENCY.parent.mkdir(parents=True, exist_ok=True)
    cxx = shutil.which("g++") or "g++"
    subprocess.run([cxx, "-O2", "-std=c++17", f"-I{ROOT / 'include'}", f"-I{PKG / 'host'}", str(src), "-o", str(DROPIN_LATENCY),
                    f"-L{PKG}", "-lrpw_b200", "-Wl,-rpath,$ORIGIN/../../ros2-recursive-patchwork-implementation_b200"], check=True)
    return DROPIN_LATENCY


def build_all(force: bool = False) -> None:
    build_cuda(force)
    build_scangen(force)
    build_tools(force)


if __name__ == "__main__":
    # python _build.py NAME [-DFLAG=VALUE ...]  ->  _variants/NAME.so (kernel experiments, timed by tools/gpu_variants.py)
    if len(sys.argv) > 1:
        (PKG / "_variants").mkdir(exist_ok=True)
        print(build_cuda(force=True, out=PKG / "_variants" / f"{sys.argv[1]}.so", extra_flags=sys.argv[2:]))
    else:
        build_all(force=True)
