"""Seeded synthetic clouds of the shapes BASELINE.json names (ctypes over synth/libscangen.so)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

_PATH = Path(__file__).resolve().parent / "synth" / "libscangen.so"
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not _PATH.exists():
            raise RuntimeError(f"{_PATH} is missing: run __graft_entry__.build()")
        lib = C.CDLL(str(_PATH))
        fp = C.POINTER(C.c_float)
        lib.rpw_synth_testsuite.argtypes = [C.c_uint32, C.c_size_t, fp]; lib.rpw_synth_testsuite.restype = None
        lib.rpw_synth_spinning.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, fp]; lib.rpw_synth_spinning.restype = C.c_size_t
        lib.rpw_synth_solidstate.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_float, fp]; lib.rpw_synth_solidstate.restype = C.c_size_t
        _lib = lib
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def testsuite_cloud(seed: int, n: int, out=None) -> np.ndarray:
    """C1: the reference test-suite generator's distributions (RP/test/test_recursive_patchwork.cpp:12-49), seeded."""
    a = np.zeros((n, 4), np.float32) if out is None else out
    _load().rpw_synth_testsuite(seed, n, _fp(a))
    return a


def spinning_scan(seed: int, beams: int = 64, steps: int = 1875, clutter: int = 0, nan_per_million: int = 200, out=None) -> np.ndarray:
    """C2 (defaults: 64 x 1875 = 120,000 returns) / C5 (beams=128, steps=2048, clutter=1)."""
    a = np.zeros((beams * steps, 4), np.float32) if out is None else out
    n = _load().rpw_synth_spinning(seed, beams, steps, clutter, nan_per_million, _fp(a))
    return a[:n]


def dense_urban_scan(seed: int = 3000) -> np.ndarray:
    """C5: 128 beams x 2048 steps = 262,144 returns, cluttered two-layer ground (deep recursion)."""
    return spinning_scan(seed, 128, 2048, 1)


def solidstate_merged(seed: int = 2000, cols: int = 450, rows: int = 300, max_range: float = 250.0) -> np.ndarray:
    """C4: three 120-degree solid-state sensors merged, banked track, about 300k returns."""
    a = np.zeros((3 * cols * rows, 4), np.float32)
    n = _load().rpw_synth_solidstate(seed, cols, rows, max_range, _fp(a))
    return a[:n].copy()


# Config each named shape is segmented with.
def config_for(shape: str):
    from .patchwork import PatchworkConfig
    if shape in ("C2", "C3", "C5"):
        return PatchworkConfig(filtering_radius=80.0)
    return PatchworkConfig()
