"""ctypes access to the CPU oracle (oracle/librpw_oracle.so) and, when it was built, the reference
itself (oracle/_ref/libref_{strict,fast}.so).  TEST INFRASTRUCTURE: importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg only."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_SO = ROOT / "oracle" / "librpw_oracle.so"
REF_DIR = ROOT / "oracle" / "_ref"


class Cfg(C.Structure):
    _fields_ = [("sensor_height", C.c_float), ("max_range", C.c_float), ("num_sectors", C.c_int32),
                ("max_iter", C.c_int32), ("adaptive_seed_height", C.c_int32), ("th_seeds", C.c_float),
                ("th_dist", C.c_float), ("th_outlier", C.c_float), ("filtering_radius", C.c_float),
                ("max_split_depth", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [(k, C.c_int64) for k in ("n_points n_clean n_zone n_binned n_ground n_root_patches n_nodes n_leaves "
                                         "n_splits n_splits_collapse n_pca_iters n_point_iters").split()] + \
               [("max_depth", C.c_int32), ("max_patch_points", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Sensor(C.Structure):
    _fields_ = [("xyz", C.c_void_p), ("n", C.c_size_t), ("rotation_deg", C.c_float), ("ego_radius", C.c_float)]


def _sensor_array(clouds, rotations_deg, ego_radii):
    arrs = [_pts(c) for c in clouds]
    sens = (Sensor * len(arrs))()
    for i, a in enumerate(arrs):
        sens[i].xyz = a.ctypes.data; sens[i].n = len(a)
        sens[i].rotation_deg = float(rotations_deg[i]); sens[i].ego_radius = float(ego_radii[i])
    return arrs, sens


NODE_DTYPE = np.dtype([("root", "<i4"), ("depth", "<i4"), ("start", "<i4"), ("n", "<i4"), ("outcome", "<i4"),
                       ("iters", "<i4"), ("n_inliers", "<i4"), ("split_axis", "<i4"), ("centroid", "<f4", 3),
                       ("normal", "<f4", 3), ("residual", "<f4"), ("median", "<f4")])

KEY_DROPPED, KEY_BEYOND, KEY_UNBINNED = 0xFFFF, 0xFFFE, 0xFFFD


def to_cfg(pc) -> Cfg:
    """PatchworkConfig (python dataclass), ctypes struct of the same layout, or None -> Cfg."""
    c = Cfg()
    if pc is None:
        Oracle().lib.rpwo_default_config(C.byref(c))
        return c
    for k, _ in Cfg._fields_:
        v = getattr(pc, k)
        setattr(c, k, int(v) if k in ("num_sectors", "max_iter", "adaptive_seed_height", "max_split_depth") else float(v))
    return c


def _pts(a):
    a = np.ascontiguousarray(a, np.float32)
    assert a.ndim == 2 and a.shape[1] in (3, 4)
    return a


class Oracle:
    _lib = None

    def __init__(self):
        if Oracle._lib is None:
            if not ORACLE_SO.exists():
                raise RuntimeError(f"{ORACLE_SO} missing: run __graft_entry__.build()")
            lib = C.CDLL(str(ORACLE_SO))
            lib.rpwo_filter_ground.restype = C.c_int
            lib.rpwo_eig3_f32.restype = C.c_int
            lib.rpwo_atan2f_restated.restype = C.c_float
            lib.rpwo_atan2f_restated.argtypes = [C.c_float, C.c_float]
            lib.rpwo_time_scan.restype = C.c_double
            Oracle._lib = lib
        self.lib = Oracle._lib

    def default_config(self) -> Cfg:
        c = Cfg()
        self.lib.rpwo_default_config(C.byref(c))
        return c

    def zone_model(self, cfg: Cfg):
        e = (C.c_float * 9)()
        a = C.c_float()
        self.lib.rpwo_zone_model(C.byref(cfg), e, C.byref(a))
        return np.array(e[:], np.float32), np.float32(a.value)

    def run(self, cfg, points, want_nodes=False):
        """Returns dict(labels, keys, dist, angle, stats[, nodes])."""
        cfg = cfg if isinstance(cfg, Cfg) else to_cfg(cfg)
        a = _pts(points)
        n = len(a)
        labels = np.zeros(n, np.uint8)
        keys = np.zeros(n, np.uint16)
        dist = np.zeros(n, np.float32)
        ang = np.zeros(n, np.float32)
        st = Stats()
        nn = C.c_size_t()
        cap = n // 2 + 1024 if want_nodes else 0
        nodes = np.zeros(cap, NODE_DTYPE)
        rc = self.lib.rpwo_filter_ground(C.byref(cfg), C.c_void_p(a.ctypes.data), C.c_size_t(n), C.c_size_t(a.shape[1]),
                                         C.c_void_p(labels.ctypes.data), C.c_void_p(keys.ctypes.data),
                                         C.c_void_p(dist.ctypes.data), C.c_void_p(ang.ctypes.data),
                                         C.c_void_p(nodes.ctypes.data if want_nodes else 0), C.c_size_t(cap), C.byref(nn), C.byref(st))
        if rc != 0:
            raise RuntimeError("rpwo_filter_ground failed")
        out = dict(labels=labels, keys=keys, dist=dist, angle=ang, stats=st.as_dict())
        if want_nodes:
            assert nn.value <= cap
            out["nodes"] = nodes[:nn.value]
        return out

    def eig3(self, mats):
        m = np.ascontiguousarray(mats, np.float32).reshape(-1, 9)
        ev = np.zeros((len(m), 3), np.float32)
        vec = np.zeros((len(m), 9), np.float32)
        for i in range(len(m)):
            self.lib.rpwo_eig3_f32(C.c_void_p(m[i].ctypes.data), C.c_void_p(ev[i].ctypes.data), C.c_void_p(vec[i].ctypes.data))
        return ev, vec.reshape(-1, 3, 3)

    def atan2_restated(self, y, x):
        f = np.frompyfunc(lambda a, b: self.lib.rpwo_atan2f_restated(float(a), float(b)), 2, 1)
        return f(np.asarray(y, np.float32), np.asarray(x, np.float32)).astype(np.float32)

    def fuse(self, clouds, rotations_deg, ego_radii):
        """Fusion front end restated: returns (fused (k,3) float32, src indices into the concatenated input)."""
        arrs, sens = _sensor_array(clouds, rotations_deg, ego_radii)
        total = sum(len(a) for a in arrs)
        fused = np.zeros((max(total, 1), 3), np.float32)
        src = np.zeros(max(total, 1), np.uint32)
        self.lib.rpwo_fuse.restype = C.c_size_t
        k = self.lib.rpwo_fuse(sens, C.c_size_t(len(arrs)), C.c_size_t(arrs[0].shape[1]), C.c_void_p(fused.ctypes.data), C.c_void_p(src.ctypes.data))
        return fused[:k], src[:k]

    def time_scan(self, cfg, points, reps=1) -> float:
        cfg = cfg if isinstance(cfg, Cfg) else to_cfg(cfg)
        a = _pts(points)
        return float(self.lib.rpwo_time_scan(C.byref(cfg), C.c_void_p(a.ctypes.data), C.c_size_t(len(a)), C.c_size_t(a.shape[1]), C.c_int(reps)))


class Reference:
    """The reference's own translation units, compiled unmodified (oracle/Makefile)."""

    def __init__(self, flavour="strict"):
        path = REF_DIR / f"libref_{flavour}.so"
        if not path.exists():
            raise FileNotFoundError(path)
        self.flavour = flavour
        self.lib = C.CDLL(str(path))
        self.lib.rpwref_filter_ground.restype = C.c_int
        self.lib.rpwref_time_scan.restype = C.c_double
        self.lib.rpwref_silence(1)

    def run(self, cfg, points):
        cfg = cfg if isinstance(cfg, Cfg) else to_cfg(cfg)
        a = _pts(points)
        n = len(a)
        g = np.zeros((n, 3), np.float32)
        ng = np.zeros((n, 3), np.float32)
        lab = np.zeros(n, np.uint8)
        n_g, n_ng, amb = C.c_size_t(), C.c_size_t(), C.c_size_t()
        rc = self.lib.rpwref_filter_ground(C.byref(cfg), C.c_void_p(a.ctypes.data), C.c_size_t(n), C.c_size_t(a.shape[1]),
                                           C.c_void_p(g.ctypes.data), C.byref(n_g), C.c_void_p(ng.ctypes.data), C.byref(n_ng),
                                           C.c_void_p(lab.ctypes.data), C.byref(amb))
        if rc != 0:
            raise RuntimeError("reference label reconstruction failed")
        return dict(ground=g[:n_g.value], non_ground=ng[:n_ng.value], labels=lab, ambiguous=amb.value)

    def fuse(self, clouds, rotations_deg, ego_radii):
        arrs, sens = _sensor_array(clouds, rotations_deg, ego_radii)
        total = sum(len(a) for a in arrs)
        fused = np.zeros((max(total, 1), 3), np.float32)
        self.lib.rpwref_fuse.restype = C.c_size_t
        k = self.lib.rpwref_fuse(sens, C.c_size_t(len(arrs)), C.c_size_t(arrs[0].shape[1]), C.c_void_p(fused.ctypes.data))
        return fused[:k]

    def sample_ground_and_obstacles(self, cfg, points, target_height=1.1, base_tol=0.5):
        """RecursivePatchwork::sampleGroundAndObstacles (RP/src/recursive_patchwork.cpp:428-465): (k, 3) float32,
        [unseeded random ground sample | obstacles in non-ground order]."""
        cfg = cfg if isinstance(cfg, Cfg) else to_cfg(cfg)
        a = _pts(points)
        cap = len(a) + 2000
        out = np.zeros((cap, 3), np.float32)
        self.lib.rpwref_sample_ground_and_obstacles.restype = C.c_size_t
        k = self.lib.rpwref_sample_ground_and_obstacles(C.byref(cfg), C.c_void_p(a.ctypes.data), C.c_size_t(len(a)), C.c_size_t(a.shape[1]),
                                                        C.c_float(target_height), C.c_float(base_tol), C.c_void_p(out.ctypes.data), C.c_size_t(cap))
        if k == C.c_size_t(-1).value:
            raise RuntimeError("reference sample larger than the buffer")
        return out[:k]

    def bev(self, mode, a, b, width, height, x_min, y_min, x_max, y_max):
        """Visualization::createGroundNonGroundImage(a, b) (mode 0) or createBEVImage(a) (mode 1),
        RP/src/visualization.cpp:18-80, through the cv::Mat stand-in: (height, width, 3) uint8 BGR."""
        a = np.ascontiguousarray(a, np.float32).reshape(-1, 3)
        b = np.ascontiguousarray(b if b is not None else np.zeros((0, 3)), np.float32).reshape(-1, 3)
        img = np.zeros((int(height), int(width), 3), np.uint8)
        self.lib.rpwref_bev.restype = C.c_int
        rc = self.lib.rpwref_bev(C.c_int(mode), C.c_void_p(a.ctypes.data), C.c_size_t(len(a)), C.c_void_p(b.ctypes.data), C.c_size_t(len(b)),
                                 C.c_int(width), C.c_int(height), C.c_float(x_min), C.c_float(y_min), C.c_float(x_max), C.c_float(y_max),
                                 C.c_void_p(img.ctypes.data))
        if rc != 0:
            raise RuntimeError("rpwref_bev failed")
        return img

    def time_scan(self, cfg, points, reps=1) -> float:
        cfg = cfg if isinstance(cfg, Cfg) else to_cfg(cfg)
        a = _pts(points)
        ng = C.c_size_t()
        return float(self.lib.rpwref_time_scan(C.byref(cfg), C.c_void_p(a.ctypes.data), C.c_size_t(len(a)), C.c_size_t(a.shape[1]),
                                               C.c_int(reps), C.byref(ng)))


def try_reference(flavour="strict"):
    try:
        return Reference(flavour)
    except (FileNotFoundError, OSError):
        return None
