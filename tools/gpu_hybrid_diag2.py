import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import oracle_lib
cfg = rpw.PatchworkConfig(filtering_radius=80.0)
pts = rpw.synth.spinning_scan(3000, 128, 2048, 1)
o = oracle_lib.Oracle().run(oracle_lib.to_cfg(cfg), pts, want_nodes=True)
p = pts[o["keys"] == 28][:, :3].astype(np.float32)
d = np.sqrt(p[:, 0] ** 2 + p[:, 1] ** 2, dtype=np.float32)
rel = np.float32(d.mean(dtype=np.float32) / np.float32(80.0))
zth = np.float32(1.2) + np.float32(0.2) * rel
seeds = p[p[:, 2] < zth]
print("patch points", len(p), "seeds", len(seeds), "z_th", zth, "z range", p[:, 2].min(), p[:, 2].max())
c = seeds.mean(0, dtype=np.float32)
dd = (seeds - c).astype(np.float32)
S = (dd.T @ dd).astype(np.float32)
sc = np.array([[S[0, 0], S[1, 0], S[1, 1], S[2, 0], S[2, 1], S[2, 2]]], np.float32)
print("scatter", sc)
w, v = np.linalg.eigh(S.astype(np.float64))
print("eigenvalues", w, "gap/scale", (w[1] - w[0]) / np.abs(S).max())
print("float64 normal", v[:, 0] * np.sign(v[2, 0]))
h = rpw.Handle(cfg.to_c(), 0, 1 << 19, 1)
for mode, name in ((1, "generic QR on scatter"), (0, "closed form")):
    nrm, _ = h.debug_normal(sc, mode)
    print(name, nrm)
cov = (sc / np.float32(len(seeds) - 1)).astype(np.float32)
print("QR on covariance", h.debug_normal(cov, 1)[0])
