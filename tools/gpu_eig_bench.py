"""Accuracy and latency of the plane-normal solvers on realistic scatter matrices."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
rng = np.random.default_rng(0)
N = 4000
# ground-like scatter: n points, sigma_z small, tilted plane
mats = []
for i in range(N):
    n = rng.integers(3, 4000)
    ext = rng.uniform(0.5, 30, 2)
    sz = rng.uniform(0.0, 0.3) if i % 10 else 0.0
    p = rng.normal(size=(n, 3)) * np.array([ext[0], ext[1], sz])
    a, b = rng.uniform(-0.3, 0.3, 2)
    p[:, 2] += a * p[:, 0] + b * p[:, 1]
    p += rng.uniform(-60, 60, 3)
    d = (p - p.mean(0)).astype(np.float32)
    S = d.T.astype(np.float32) @ d
    mats.append([S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]])
mats = np.array(mats, np.float32)
full = np.zeros((N, 3, 3)); 
full[:, 0, 0] = mats[:, 0]; full[:, 1, 0] = full[:, 0, 1] = mats[:, 1]; full[:, 1, 1] = mats[:, 2]
full[:, 2, 0] = full[:, 0, 2] = mats[:, 3]; full[:, 2, 1] = full[:, 1, 2] = mats[:, 4]; full[:, 2, 2] = mats[:, 5]
w, v = np.linalg.eigh(full)
ref = v[:, :, 0]; ref *= np.sign(ref[:, 2:3] + 1e-300)
gap = (w[:, 1] - w[:, 0]) / np.abs(w).max(1)
h = rpw.Handle(None, 0, 1 << 16, 1)
for mode, name in ((0, "closed-form fp64"), (2, "eigen QR restructured"), (1, "eigen QR f32")):
    nrm, cyc = h.debug_normal(mats, mode)
    nrm2, cyc = h.debug_normal(mats, mode)
    nd = nrm.astype(np.float64)
    ang = np.arctan2(np.linalg.norm(np.cross(nd, ref), axis=1), np.abs((nd * ref).sum(1)))
    ok = gap > 1e-3
    if mode == 1: qr1 = nrm.copy()
    if mode == 2: qr2 = nrm.copy()
    print(f"{name:18s} cycles mean {cyc.mean():7.0f} median {np.median(cyc):6.0f} p99 {np.quantile(cyc,.99):6.0f} | angle vs float64 eigh: max(gap>1e-3) {ang[ok].max():.2e} "
          f"p99 {np.quantile(ang[ok],.99):.2e} | unit-norm err {np.abs(np.linalg.norm(nrm,axis=1)-1).max():.1e}")
print("restructured QR == generic QR bitwise:", bool(np.array_equal(qr1.view(np.uint32), qr2.view(np.uint32))), "mismatching rows", int((qr1.view(np.uint32) != qr2.view(np.uint32)).any(1).sum()))
# latency with one warp per SM (no issue contention)
for mode, name in ((0, "closed-form fp64"), (2, "eigen QR restructured"), (1, "eigen QR generic")):
    _, cyc = h.debug_normal(mats[:148], mode)
    _, cyc = h.debug_normal(mats[:148], mode)
    print(f"alone-on-SM latency {name:22s} mean {cyc.mean():7.0f} median {np.median(cyc):6.0f} max {cyc.max():6.0f}")
