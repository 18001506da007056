"""B200-native Recursive Patchwork ground segmentation (the per-scan hot path only).

Layout:
    csrc/       hand-written sm_100a kernels + the C-ABI host code  -> librpw_b200.so
    host/       C++ drop-in for the reference's recursive_patchwork.hpp, over the C-ABI
    capi.py     ctypes binding of include/rpw_b200.h
    patchwork.py  Python mirror of the reference class (RecursivePatchwork, PatchworkConfig)
    synth.py    seeded synthetic scans of the shapes BASELINE.json names
    sharding.py frame-level sharding across GPUs (no data-path collective)
"""
from . import capi, synth, sharding  # noqa: F401
from .capi import Handle, RpwError, load_library  # noqa: F401
from .patchwork import PatchworkConfig, RecursivePatchwork, clouds_from_labels  # noqa: F401
