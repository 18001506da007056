"""Host-side mirror of the reference's segmentation interface, over the C-ABI.

Mirrors `recursive_patchwork::RecursivePatchwork` (RP/include/recursive_patchwork.hpp:47-87): same
method names and argument meaning, same degenerate returns, config struct with the same fields and
defaults.  All compute goes through librpw_b200.so; there is no Python or CPU implementation of
the path here."""
from __future__ import annotations

from dataclasses import dataclass, asdict

import numpy as np

from . import capi


@dataclass
class PatchworkConfig:
    """struct PatchworkConfig, RP/include/recursive_patchwork.hpp:25-36 (same names, same defaults)."""
    sensor_height: float = 1.2
    max_range: float = 150.0
    num_sectors: int = 10
    max_iter: int = 100
    adaptive_seed_height: bool = True
    th_seeds: float = 0.15
    th_dist: float = 0.2
    th_outlier: float = 0.08
    filtering_radius: float = 150.0
    max_split_depth: int = 1000

    def to_c(self) -> capi.RpwConfig:
        d = asdict(self)
        d["adaptive_seed_height"] = 1 if self.adaptive_seed_height else 0
        return capi.RpwConfig(**d)

    @staticmethod
    def from_c(c: capi.RpwConfig) -> "PatchworkConfig":
        return PatchworkConfig(c.sensor_height, c.max_range, c.num_sectors, c.max_iter, bool(c.adaptive_seed_height),
                               c.th_seeds, c.th_dist, c.th_outlier, c.filtering_radius, c.max_split_depth)


def clouds_from_labels(points: np.ndarray, labels: np.ndarray):
    """The two clouds of filterGroundPoints, in the reference's order
    (RP/src/recursive_patchwork.cpp:402-419): ground in input order; non-ground in input order
    followed by the beyond-radius points in input order.  Host-side helper for callers that
    only asked for labels; filterGroundPoints itself gets the clouds from the device."""
    p = points[:, :3]
    ground = p[labels == capi.LABEL_GROUND]
    non_ground = np.concatenate([p[labels == capi.LABEL_NONGROUND], p[labels == capi.LABEL_BEYOND]])
    return ground, non_ground


class RecursivePatchwork:
    """RecursivePatchwork(config) — RP/include/recursive_patchwork.hpp:47-87 on a B200.

    max_points / max_batch size the device buffers (the reference allocates per call; here the
    handle owns them)."""

    def __init__(self, config: PatchworkConfig | None = None, device: int = 0, max_points: int = 1 << 19, max_batch: int = 1):
        self._config = config or PatchworkConfig()
        self._handle = capi.Handle(self._config.to_c(), device, max_points * max_batch, max_batch)

    # -- configuration (RP/include/recursive_patchwork.hpp:66-67) ---------------------------
    def setConfig(self, config: PatchworkConfig) -> None:
        self._handle.set_config(config.to_c())
        self._config = config

    def getConfig(self) -> PatchworkConfig:
        return self._config

    @property
    def handle(self) -> capi.Handle:
        return self._handle

    # -- main processing (RP/include/recursive_patchwork.hpp:53-59) --------------------------
    def filterGroundPoints(self, points):
        """points: (n, 3|4) float32.  Returns (ground_points, non_ground_points) as (k, 3) arrays.
        Empty input returns two empty clouds (RP/src/recursive_patchwork.cpp:316-318)."""
        a = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, np.shape(points)[-1] if np.ndim(points) == 2 else 3)
        if len(a) == 0:
            return np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32)
        ground, non_ground, _ = self._handle.segment_clouds(a)  # assembled on the device (K4), the host only copies
        return ground, non_ground

    def sampleGroundAndObstacles(self, points, target_height: float = 1.1, base_tol: float = 0.5, seed: int = 0):
        """RP/src/recursive_patchwork.cpp:428-465: a 2000-point random ground context sample followed by the
        non-ground points outside the 2.5 m ego radius within base_tol of target_height.  All on the device;
        seed = 0 draws like the reference (unseeded), any other value is reproducible."""
        a = np.ascontiguousarray(points, dtype=np.float32).reshape(-1, np.shape(points)[-1] if np.ndim(points) == 2 else 3)
        if len(a) == 0:
            return np.zeros((0, 3), np.float32)
        self._handle.segment(a)
        ground, obstacles = self._handle.sample_ground_and_obstacles(len(a), target_height, base_tol, 2.5, 2000, seed)
        return np.concatenate([ground, obstacles])

    def filterGroundLabels(self, points) -> np.ndarray:
        """The north-star addition: per-input-point labels (0 non-ground, 1 ground, 2 beyond
        the filtering radius, 3 dropped as non-finite)."""
        a = np.ascontiguousarray(points, dtype=np.float32)
        if len(a) == 0:
            return np.zeros(0, np.uint8)
        return self._handle.segment(a)

    def filterGroundLabelsBatch(self, clouds):
        return self._handle.segment_batch(clouds)

    def cleanPoints(self, points):
        """Drop points with a non-finite coordinate, order preserved (RP/src/recursive_patchwork.cpp:19-35).
        Host utility (numpy); the segmentation path does its own cleaning on the device."""
        a = np.asarray(points, dtype=np.float32)
        return a[np.isfinite(a[:, :3]).all(axis=1)]

    def close(self):
        self._handle.close()
