"""Cycles of one plane-normal solve (one warp per matrix, alone on its SM partition) for plane-like
covariances: generic Eigen sequence, restructured (IEEE ops), branch-free arithmetic, closed form."""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
rng = np.random.default_rng(5)
mats = []
for i in range(4000):
    ext = rng.uniform(2, 30, 2)
    p = rng.normal(size=(int(rng.integers(50, 3000)), 3)) * np.array([ext[0], ext[1], rng.uniform(0.01, 0.2)])
    a, b = rng.uniform(-0.1, 0.1, 2)
    p[:, 2] += a * p[:, 0] + b * p[:, 1]
    d = (p - p.mean(0)).astype(np.float32)
    S = (d.T @ d) / np.float32(len(p) - 1)
    mats.append([S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]])
mats = np.array(mats, np.float32)
if len(sys.argv) > 1 and sys.argv[1] == "real":
    # covariances of the final inlier sets of real root patches (oracle labels on spinning scans)
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_lib
    o = oracle_lib.Oracle()
    cfg = o.default_config(); cfg.filtering_radius = 80.0
    mats = []
    for seed in range(1000, 1016):
        pts = rpw.synth.spinning_scan(seed)
        r = o.run(cfg, pts)
        keys, lab = r["keys"], r["labels"]
        for k in np.unique(keys[keys < 0xFFF0]):
            p = pts[(keys == k) & (lab == 1)].astype(np.float32)
            if len(p) < 10: continue
            d = p - p.mean(0, dtype=np.float32)
            S = (d.T @ d) / np.float32(len(p) - 1)
            mats.append([S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]])
    mats = np.array(mats, np.float32)
    print("real covariances:", len(mats))
h = rpw.Handle(rpw.PatchworkConfig().to_c(), 0, 1 << 16, 1)
ref = None
for mode, name in ((1, "generic QR"), (2, "restructured, IEEE ops"), (3, "branch-free + fallback"), (4, "branch-free raw"), (0, "closed form fp64")):
    h.debug_normal(mats, mode)
    nrm, cyc = h.debug_normal(mats, mode)
    solo = np.concatenate([h.debug_normal(mats[k:k + 100], mode)[1] for k in range(0, 2000, 100)])  # <= 1 warp per SM
    if mode == 1: ref = nrm
    same = np.array_equal(ref.view(np.uint32), nrm.view(np.uint32))
    print(f"mode {mode} {name:26s} cycles median {np.median(cyc):6.0f} mean {cyc.mean():7.0f} p99 {np.quantile(cyc, .99):6.0f}  | solo median {np.median(solo):6.0f} p99 {np.quantile(solo, .99):6.0f} | bit-identical to mode 1: {same}"
          + (f"  out-of-range {np.mean(nrm[:, 2] == 2.0):.4f}" if mode == 4 else ""))
