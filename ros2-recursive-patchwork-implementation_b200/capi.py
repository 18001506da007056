"""ctypes binding of the C-ABI in include/rpw_b200.h (librpw_b200.so).

This is the only way Python reaches the CUDA path.  If the shared library is missing, or no
sm_100 device is present, everything here raises — there is no CPU fallback."""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ["RPW_B200_LIB"]) if os.environ.get("RPW_B200_LIB") else PKG / "librpw_b200.so"  # (override: kernel experiments)

RPW_OK, RPW_ERR_BAD_ARG, RPW_ERR_NO_DEVICE, RPW_ERR_CUDA, RPW_ERR_CAPACITY, RPW_ERR_ALLOC = range(6)
LABEL_NONGROUND, LABEL_GROUND, LABEL_BEYOND, LABEL_DROPPED, LABEL_EGO = 0, 1, 2, 3, 4
KEY_DROPPED, KEY_BEYOND, KEY_UNBINNED, KEY_EGO = 0xFFFF, 0xFFFE, 0xFFFD, 0xFFFC
SOLVER_EIGEN_QR, SOLVER_CLOSED_FORM, SOLVER_HYBRID, SOLVER_REFERENCE = 0, 1, 2, 3
NODE_SMALL, NODE_AREA, NODE_FLAT, NODE_FIT, NODE_SPLIT = 1, 2, 3, 4, 5


class RpwConfig(C.Structure):
    """struct rpw_config == PatchworkConfig (RP/include/recursive_patchwork.hpp:25-36)."""
    _fields_ = [("sensor_height", C.c_float), ("max_range", C.c_float), ("num_sectors", C.c_int32),
                ("max_iter", C.c_int32), ("adaptive_seed_height", C.c_int32), ("th_seeds", C.c_float),
                ("th_dist", C.c_float), ("th_outlier", C.c_float), ("filtering_radius", C.c_float),
                ("max_split_depth", C.c_int32)]


class RpwStats(C.Structure):
    _fields_ = [("n_points", C.c_uint64), ("n_ground", C.c_uint64), ("n_nonground", C.c_uint64),
                ("n_beyond", C.c_uint64), ("n_dropped", C.c_uint64), ("n_levels", C.c_uint32),
                ("n_nodes", C.c_uint32), ("kernel_launches", C.c_uint64)]


class RpwSensorCloud(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("n", C.c_size_t), ("rotation_deg", C.c_float), ("ego_radius", C.c_float)]


class RpwProfile(C.Structure):
    _fields_ = [("ms", C.c_double * 4), ("launches", C.c_uint64 * 4), ("fit_grid_blocks", C.c_uint32),
                ("fit_smem_points", C.c_uint32)]


PROF_KERNELS = ("bin", "offsets", "scatter", "fit")

NODE_DTYPE = np.dtype([("scan", "<i4"), ("root", "<i4"), ("depth", "<i4"), ("start", "<i4"), ("n", "<i4"),
                       ("outcome", "<i4"), ("iters", "<i4"), ("n_inliers", "<i4"), ("split_axis", "<i4"),
                       ("centroid", "<f4", 3), ("normal", "<f4", 3), ("residual", "<f4"), ("median", "<f4"),
                       ("mean_dist", "<f4")])

# Every symbol include/rpw_b200.h declares (tests check that the library exports all of them).
EXPORTS = ["rpw_default_config", "rpw_zone_model", "rpw_create", "rpw_destroy", "rpw_set_config", "rpw_get_config", "rpw_reserve", "rpw_capacity",
           "rpw_set_plane_solver", "rpw_set_exact_replay", "rpw_set_stream", "rpw_last_error", "rpw_segment", "rpw_segment_batch", "rpw_segment_batch_async", "rpw_wait",
           "rpw_segment_pc2", "rpw_segment_fused", "rpw_segment_clouds", "rpw_segment_clouds_view", "rpw_last_clouds", "rpw_sample_ground_and_obstacles", "rpw_bev_image", "rpw_segment_device", "rpw_debug_keys", "rpw_debug_enable_nodes", "rpw_debug_nodes",
           "rpw_debug_eig3", "rpw_debug_normal", "rpw_debug_atan2", "rpw_debug_fit_timing", "rpw_debug_fit_trace", "rpw_profile_enable", "rpw_profile_read", "rpw_copy_probe", "rpw_host_alloc", "rpw_host_free", "rpw_kernel_launches", "rpw_scan_graph", "rpw_abi_version"]

_lib = None


class RpwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rpw error {code}: {msg}")
        self.code = code


def load_library() -> C.CDLL:
    """Loads librpw_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RpwError(-1, f"{LIB_PATH} is missing: build it with __graft_entry__.build() — there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    vp, sz, u8p, fp = C.c_void_p, C.c_size_t, C.POINTER(C.c_uint8), C.POINTER(C.c_float)
    cfgp = C.POINTER(RpwConfig)
    lib.rpw_default_config.argtypes = [cfgp]; lib.rpw_default_config.restype = None
    lib.rpw_zone_model.argtypes = [cfgp, fp, fp]; lib.rpw_zone_model.restype = C.c_int
    lib.rpw_create.argtypes = [cfgp, C.c_int, sz, sz, C.POINTER(vp)]; lib.rpw_create.restype = C.c_int
    lib.rpw_destroy.argtypes = [vp]; lib.rpw_destroy.restype = None
    lib.rpw_set_config.argtypes = [vp, cfgp]; lib.rpw_set_config.restype = C.c_int
    lib.rpw_get_config.argtypes = [vp, cfgp]; lib.rpw_get_config.restype = C.c_int
    lib.rpw_reserve.argtypes = [vp, sz, sz]; lib.rpw_reserve.restype = C.c_int
    lib.rpw_capacity.argtypes = [vp, C.POINTER(sz), C.POINTER(sz)]; lib.rpw_capacity.restype = C.c_int
    lib.rpw_set_plane_solver.argtypes = [vp, C.c_int]; lib.rpw_set_plane_solver.restype = C.c_int
    lib.rpw_set_exact_replay.argtypes = [vp, C.c_int]; lib.rpw_set_exact_replay.restype = C.c_int
    lib.rpw_set_stream.argtypes = [vp, vp]; lib.rpw_set_stream.restype = C.c_int
    lib.rpw_last_error.argtypes = [vp]; lib.rpw_last_error.restype = C.c_char_p
    lib.rpw_segment.argtypes = [vp, vp, sz, sz, vp, C.POINTER(RpwStats)]; lib.rpw_segment.restype = C.c_int
    lib.rpw_segment_batch.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), sz, sz, C.POINTER(vp), C.POINTER(RpwStats)]
    lib.rpw_segment_batch.restype = C.c_int
    lib.rpw_segment_batch_async.argtypes = [vp, C.POINTER(vp), C.POINTER(sz), sz, sz, C.POINTER(vp)]
    lib.rpw_segment_batch_async.restype = C.c_int
    lib.rpw_wait.argtypes = [vp, C.POINTER(RpwStats)]; lib.rpw_wait.restype = C.c_int
    lib.rpw_segment_pc2.argtypes = [vp, vp, sz, sz, sz, sz, sz, vp, C.POINTER(RpwStats)]; lib.rpw_segment_pc2.restype = C.c_int
    lib.rpw_segment_fused.argtypes = [vp, C.POINTER(RpwSensorCloud), sz, sz, C.POINTER(vp), C.POINTER(RpwStats)]
    lib.rpw_segment_fused.restype = C.c_int
    lib.rpw_segment_clouds.argtypes = [vp, vp, sz, sz, vp, vp, C.POINTER(sz), vp, C.POINTER(sz)]
    lib.rpw_segment_clouds.restype = C.c_int
    lib.rpw_segment_clouds_view.argtypes = [vp, vp, sz, sz, vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]
    lib.rpw_segment_clouds_view.restype = C.c_int
    lib.rpw_last_clouds.argtypes = [vp, vp, vp, C.c_int, C.POINTER(C.c_uint64)]
    lib.rpw_last_clouds.restype = C.c_int
    lib.rpw_sample_ground_and_obstacles.argtypes = [vp, C.c_float, C.c_float, C.c_float, sz, C.c_uint64, vp, sz, C.POINTER(sz), C.POINTER(sz)]
    lib.rpw_sample_ground_and_obstacles.restype = C.c_int
    lib.rpw_bev_image.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp]
    lib.rpw_bev_image.restype = C.c_int
    lib.rpw_segment_device.argtypes = [vp, vp, C.POINTER(C.c_uint64), sz, vp]; lib.rpw_segment_device.restype = C.c_int
    lib.rpw_debug_keys.argtypes = [vp, vp, sz]; lib.rpw_debug_keys.restype = C.c_int
    lib.rpw_debug_enable_nodes.argtypes = [vp, C.c_int]; lib.rpw_debug_enable_nodes.restype = C.c_int
    lib.rpw_debug_nodes.argtypes = [vp, vp, sz, C.POINTER(sz)]; lib.rpw_debug_nodes.restype = C.c_int
    lib.rpw_debug_eig3.argtypes = [vp, vp, sz, vp, vp]; lib.rpw_debug_eig3.restype = C.c_int
    lib.rpw_debug_normal.argtypes = [vp, vp, sz, C.c_int, vp, vp]; lib.rpw_debug_normal.restype = C.c_int
    lib.rpw_debug_atan2.argtypes = [vp, vp, vp, sz, vp]; lib.rpw_debug_atan2.restype = C.c_int
    lib.rpw_debug_fit_timing.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64)]; lib.rpw_debug_fit_timing.restype = C.c_int
    lib.rpw_debug_fit_trace.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]; lib.rpw_debug_fit_trace.restype = C.c_int
    lib.rpw_profile_enable.argtypes = [vp, C.c_int]; lib.rpw_profile_enable.restype = C.c_int
    lib.rpw_profile_read.argtypes = [vp, C.POINTER(RpwProfile)]; lib.rpw_profile_read.restype = C.c_int
    lib.rpw_copy_probe.argtypes = [C.c_int, sz, C.c_int, C.c_int, C.POINTER(C.c_double)]; lib.rpw_copy_probe.restype = C.c_int
    lib.rpw_host_alloc.argtypes = [sz]; lib.rpw_host_alloc.restype = vp
    lib.rpw_host_free.argtypes = [vp]; lib.rpw_host_free.restype = None
    lib.rpw_kernel_launches.argtypes = [vp]; lib.rpw_kernel_launches.restype = C.c_uint64
    lib.rpw_scan_graph.argtypes = [vp, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]; lib.rpw_scan_graph.restype = C.c_int
    lib.rpw_abi_version.argtypes = []; lib.rpw_abi_version.restype = C.c_int
    _lib = lib
    return lib


def default_config() -> RpwConfig:
    c = RpwConfig()
    load_library().rpw_default_config(C.byref(c))
    return c


def zone_model(cfg: RpwConfig):
    edges = (C.c_float * 9)()
    ang = C.c_float()
    rc = load_library().rpw_zone_model(C.byref(cfg), edges, C.byref(ang))
    if rc != RPW_OK:
        raise RpwError(rc, "rpw_zone_model")
    return np.array(edges[:], np.float32), np.float32(ang.value)


def copy_probe(device: int, nbytes: int, reps: int, d2h: bool = False, write_combined: bool = False, both: bool = False) -> float:
    """GB/s of `reps` pinned host<->device copies of `nbytes` on `device` (rpw_copy_probe); both: with nbytes / 12 going back
    at the same time (the rate is that of the host-to-device bytes)."""
    sec = C.c_double()
    rc = load_library().rpw_copy_probe(int(device), int(nbytes), int(reps), (1 if d2h else 0) | (2 if write_combined else 0) | (4 if both else 0), C.byref(sec))
    if rc != RPW_OK:
        raise RpwError(rc, "rpw_copy_probe")
    return nbytes * reps / sec.value / 1e9


class PinnedArray:
    """numpy view of pinned host memory from rpw_host_alloc (freed with the object)."""

    def __init__(self, shape, dtype):
        self._lib = load_library()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        self.ptr = self._lib.rpw_host_alloc(max(n, 1))
        if not self.ptr:
            raise RpwError(RPW_ERR_ALLOC, "rpw_host_alloc failed")
        buf = (C.c_uint8 * max(n, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                self.array = None
                self._lib.rpw_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


class Handle:
    """Owns one rpw_handle (one device, one stream, its buffers)."""

    def __init__(self, cfg: RpwConfig | None = None, device: int = 0, max_total_points: int = 1 << 20, max_batch: int = 1):
        self.lib = load_library()
        self._h = C.c_void_p()
        rc = self.lib.rpw_create(C.byref(cfg) if cfg is not None else None, device, max_total_points, max_batch, C.byref(self._h))
        if rc != RPW_OK:
            msg = self.lib.rpw_last_error(None).decode()
            self._h = None
            raise RpwError(rc, msg)
        self.max_total_points = max_total_points
        self.max_batch = max_batch

    def close(self):
        if getattr(self, "_h", None):
            self.lib.rpw_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != RPW_OK:
            raise RpwError(rc, self.lib.rpw_last_error(self._h).decode())

    # -- configuration ---------------------------------------------------------------------
    def set_config(self, cfg: RpwConfig):
        self._check(self.lib.rpw_set_config(self._h, C.byref(cfg)))

    def get_config(self) -> RpwConfig:
        c = RpwConfig()
        self._check(self.lib.rpw_get_config(self._h, C.byref(c)))
        return c

    def reserve(self, max_total_points: int, max_batch: int = 1):
        """Grows the handle in place (never shrinks); results of earlier calls are gone afterwards."""
        self._check(self.lib.rpw_reserve(self._h, int(max_total_points), int(max_batch)))
        self.max_total_points, self.max_batch = self.capacity()

    def capacity(self):
        n, b = C.c_size_t(), C.c_size_t()
        self._check(self.lib.rpw_capacity(self._h, C.byref(n), C.byref(b)))
        return int(n.value), int(b.value)

    def set_plane_solver(self, solver: int):
        """SOLVER_HYBRID (2, default), SOLVER_EIGEN_QR (0, the reference's own float QR sequence) or SOLVER_CLOSED_FORM (1)."""
        self._check(self.lib.rpw_set_plane_solver(self._h, int(solver)))

    def set_exact_replay(self, max_fast_iterations: int):
        """Fits of more than this many iterations are redone in the reference's arithmetic order (0: all, < 0: off)."""
        self._check(self.lib.rpw_set_exact_replay(self._h, int(max_fast_iterations)))

    def set_stream(self, cuda_stream: int | None):
        self._check(self.lib.rpw_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # -- the path --------------------------------------------------------------------------
    @staticmethod
    def _as_points(points):
        a = np.ascontiguousarray(points, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] < 3:
            raise ValueError("points must be (n, k >= 3) float32 with x, y, z first")
        return a

    def segment(self, points, want_stats=False):
        a = self._as_points(points)
        labels = np.empty(len(a), np.uint8)
        st = RpwStats()
        self._check(self.lib.rpw_segment(self._h, a.ctypes.data, len(a), a.shape[1] * 4, labels.ctypes.data, C.byref(st) if want_stats else None))
        return (labels, st) if want_stats else labels

    def segment_batch(self, clouds, want_stats=False, labels_out=None):
        arrs = [self._as_points(c) for c in clouds]
        stride = arrs[0].shape[1]
        if any(a.shape[1] != stride for a in arrs):
            raise ValueError("all scans of a batch must share one stride")
        B = len(arrs)
        labels = labels_out if labels_out is not None else [np.empty(len(a), np.uint8) for a in arrs]
        cp = (C.c_void_p * B)(*[a.ctypes.data for a in arrs])
        lp = (C.c_void_p * B)(*[l.ctypes.data for l in labels])
        ns = (C.c_size_t * B)(*[len(a) for a in arrs])
        st = RpwStats()
        self._check(self.lib.rpw_segment_batch(self._h, cp, ns, B, stride * 4, lp, C.byref(st) if want_stats else None))
        return (labels, st) if want_stats else labels

    def segment_batch_async(self, ptrs, ns, stride_bytes, label_ptrs):
        """Raw asynchronous form: arrays of host pointers (pinned), sizes; pair with wait()."""
        B = len(ns)
        cp = (C.c_void_p * B)(*ptrs)
        lp = (C.c_void_p * B)(*label_ptrs)
        na = (C.c_size_t * B)(*ns)
        self._check(self.lib.rpw_segment_batch_async(self._h, cp, na, B, stride_bytes, lp))

    def wait(self):
        self._check(self.lib.rpw_wait(self._h, None))

    def segment_pc2(self, data, n_points, point_step, off_x=0, off_y=4, off_z=8):
        """data: bytes-like PointCloud2 payload (n_points * point_step bytes)."""
        buf = np.frombuffer(data, dtype=np.uint8)
        assert buf.size >= n_points * point_step
        labels = np.empty(n_points, np.uint8)
        self._check(self.lib.rpw_segment_pc2(self._h, buf.ctypes.data, n_points, point_step, off_x, off_y, off_z, labels.ctypes.data, None))
        return labels

    def segment_fused(self, clouds, rotations_deg, ego_radii, want_stats=False):
        """One merged multi-LiDAR frame: per-sensor clouds, yaw angles (degrees) and ego radii."""
        arrs = [self._as_points(c) for c in clouds]
        stride = arrs[0].shape[1]
        k = len(arrs)
        sens = (RpwSensorCloud * k)()
        labels = [np.empty(len(a), np.uint8) for a in arrs]
        for i, a in enumerate(arrs):
            sens[i].xyz = a.ctypes.data; sens[i].n = len(a)
            sens[i].rotation_deg = float(rotations_deg[i]); sens[i].ego_radius = float(ego_radii[i])
        lp = (C.c_void_p * k)(*[l.ctypes.data for l in labels])
        st = RpwStats()
        self._check(self.lib.rpw_segment_fused(self._h, sens, k, stride * 4, lp, C.byref(st) if want_stats else None))  # (stats: a host pass over the labels)
        return (labels, st) if want_stats else labels

    def segment_clouds(self, points):
        a = self._as_points(points)
        n = len(a)
        labels = np.empty(n, np.uint8)
        g = np.empty((n, 3), np.float32)
        ng = np.empty((n, 3), np.float32)
        n_g, n_ng = C.c_size_t(), C.c_size_t()
        self._check(self.lib.rpw_segment_clouds(self._h, a.ctypes.data, n, a.shape[1] * 4, labels.ctypes.data,
                                                g.ctypes.data, C.byref(n_g), ng.ctypes.data, C.byref(n_ng)))
        return g[:n_g.value], ng[:n_ng.value], labels

    def last_clouds(self, scan_counts, total=None):
        """Ground / non-ground clouds of the last call's scans, assembled on the device (host copies).
        scan_counts: points per scan of that call.  Returns a list of (ground, nonground) arrays."""
        n = np.asarray(scan_counts, np.uint64)
        off = np.zeros(len(n) + 1, np.uint64); off[1:] = np.cumsum(n)
        total = int(off[-1]) if total is None else total
        g = np.empty((max(total, 1), 3), np.float32)
        ng = np.empty((max(total, 1), 3), np.float32)
        cnt = (C.c_uint64 * (2 * len(n)))()
        self._check(self.lib.rpw_last_clouds(self._h, g.ctypes.data, ng.ctypes.data, 0, cnt))
        return [(g[int(off[b]):int(off[b]) + int(cnt[2 * b])].copy(), ng[int(off[b]):int(off[b]) + int(cnt[2 * b + 1])].copy())
                for b in range(len(n))]

    def sample_ground_and_obstacles(self, n_points, target_height=1.1, base_tol=0.5, ego_radius=2.5, sample_size=2000, seed=0):
        """sampleGroundAndObstacles for the last single-scan call: (ground sample, obstacles) as (k, 3) arrays."""
        cap = int(n_points) + int(sample_size)
        out = np.empty((max(cap, 1), 3), np.float32)
        ng, no = C.c_size_t(), C.c_size_t()
        self._check(self.lib.rpw_sample_ground_and_obstacles(self._h, target_height, base_tol, ego_radius, int(sample_size), int(seed),
                                                             out.ctypes.data, cap, C.byref(ng), C.byref(no)))
        return out[:ng.value].copy(), out[ng.value:ng.value + no.value].copy()

    BEV_CLASSES, BEV_HEIGHT_NONGROUND, BEV_HEIGHT_ALL = 0, 1, 2

    def bev_image(self, mode, width, height, x_min, y_min, x_max, y_max):
        """BEV raster of the last single-scan call: (height, width, 3) uint8, BGR."""
        img = np.empty((int(height), int(width), 3), np.uint8)
        self._check(self.lib.rpw_bev_image(self._h, int(mode), int(width), int(height), x_min, y_min, x_max, y_max, img.ctypes.data))
        return img

    def last_clouds_device(self, d_ground_ptr: int, d_nonground_ptr: int, n_scans: int):
        """Same with caller-provided device buffers (3 floats per point of the call each); returns counts [n_scans, 2]."""
        cnt = (C.c_uint64 * (2 * n_scans))()
        self._check(self.lib.rpw_last_clouds(self._h, C.c_void_p(d_ground_ptr), C.c_void_p(d_nonground_ptr), 1, cnt))
        return np.array(cnt[:], np.uint64).reshape(n_scans, 2)

    def segment_device(self, d_points_ptr: int, scan_offsets, d_labels_ptr: int):
        off = np.ascontiguousarray(scan_offsets, dtype=np.uint64)
        self._check(self.lib.rpw_segment_device(self._h, C.c_void_p(d_points_ptr), off.ctypes.data_as(C.POINTER(C.c_uint64)),
                                                len(off) - 1, C.c_void_p(d_labels_ptr)))

    # -- parity / debug --------------------------------------------------------------------
    def debug_keys(self, n_total: int):
        k = np.empty(n_total, np.uint16)
        self._check(self.lib.rpw_debug_keys(self._h, k.ctypes.data, n_total))
        return k

    def enable_nodes(self, on=True):
        self._check(self.lib.rpw_debug_enable_nodes(self._h, 1 if on else 0))

    def debug_nodes(self):
        cnt = C.c_size_t()
        self._check(self.lib.rpw_debug_nodes(self._h, None, 0, C.byref(cnt)))
        out = np.zeros(cnt.value, NODE_DTYPE)
        if cnt.value:
            self._check(self.lib.rpw_debug_nodes(self._h, out.ctypes.data, cnt.value, C.byref(cnt)))
        return out

    def debug_eig3(self, mats):
        m = np.ascontiguousarray(mats, np.float32).reshape(-1, 9)
        ev = np.empty((len(m), 3), np.float32)
        vec = np.empty((len(m), 3, 3), np.float32)
        self._check(self.lib.rpw_debug_eig3(self._h, m.ctypes.data, len(m), ev.ctypes.data, vec.ctypes.data))
        return ev, vec

    def debug_normal(self, scatter6, mode=0):
        m = np.ascontiguousarray(scatter6, np.float32).reshape(-1, 6)
        nrm = np.empty((len(m), 3), np.float32)
        cyc = np.empty(len(m), np.uint32)
        self._check(self.lib.rpw_debug_normal(self._h, m.ctypes.data, len(m), mode, nrm.ctypes.data, cyc.ctypes.data))
        return nrm, cyc

    def debug_atan2(self, y, x):
        y = np.ascontiguousarray(y, np.float32)
        x = np.ascontiguousarray(x, np.float32)
        out = np.empty_like(y)
        self._check(self.lib.rpw_debug_atan2(self._h, y.ctypes.data, x.ctypes.data, y.size, out.ctypes.data))
        return out

    def fit_timing(self, enable=True):
        """Reads+clears the fit kernel's cycle counters, then switches the accounting on/off."""
        out = (C.c_uint64 * 16)()
        self._check(self.lib.rpw_debug_fit_timing(self._h, 1 if enable else 0, out))
        names = ["load", "seeds", "cov", "eig", "dist", "final", "label", "split", "fetch", "gridsync", "nodes", "iters", "load_loop", "dist_reduce", "qr_only", "replays"]
        return {k: int(out[i]) for i, k in enumerate(names)}

    TRACE_DTYPE = np.dtype([("t_start_ns", "<u8"), ("t_end_ns", "<u8"), ("sm", "<u4"), ("n", "<u4"), ("depth", "<u2"),
                            ("size_class", "<u2"), ("iters", "<u4")])

    def fit_trace_arm(self, cap=1 << 20):
        """Arms (cap > 0) or switches off (cap = 0) the per-node timeline of the fit kernels."""
        self._check(self.lib.rpw_debug_fit_trace(self._h, None, int(cap), None))

    def fit_trace_read(self, cap=1 << 20):
        """Records since arming (structured array, TRACE_DTYPE) and the number of nodes seen; clears the trace."""
        out = np.zeros(int(cap), self.TRACE_DTYPE)
        seen = C.c_size_t(0)
        self._check(self.lib.rpw_debug_fit_trace(self._h, out.ctypes.data, int(cap), C.byref(seen)))
        return out[:min(int(seen.value), int(cap))], int(seen.value)

    def profile_enable(self, on=True):
        self._check(self.lib.rpw_profile_enable(self._h, 1 if on else 0))

    def profile_read(self) -> dict:
        p = RpwProfile()
        self._check(self.lib.rpw_profile_read(self._h, C.byref(p)))
        out = {k: dict(ms=p.ms[i], launches=int(p.launches[i])) for i, k in enumerate(PROF_KERNELS)}
        out["fit_grid_blocks"] = int(p.fit_grid_blocks)
        out["fit_smem_points"] = int(p.fit_smem_points)
        return out

    def scan_graph(self, enable: int = -1):
        """(graph launches, graph captures) of the single-scan path; enable = 0 / 1 switches it off / on."""
        a, b = C.c_uint64(), C.c_uint64()
        self._check(self.lib.rpw_scan_graph(self._h, int(enable), C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def kernel_launches(self) -> int:
        return int(self.lib.rpw_kernel_launches(self._h))
