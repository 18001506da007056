"""Where does a non-reference solver leave the reference's tree of fits on the C5 stress scene?"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import oracle_lib, parity
solver = int(sys.argv[1]) if len(sys.argv) > 1 else 2
cfg = rpw.PatchworkConfig(filtering_radius=80.0)
pts = rpw.synth.spinning_scan(3000, 128, 2048, 1)
o = oracle_lib.Oracle().run(oracle_lib.to_cfg(cfg), pts, want_nodes=True)
h = rpw.Handle(cfg.to_c(), 0, 1 << 19, 1)
h.set_plane_solver(solver); h.enable_nodes(True)
lab = h.segment(pts)
gn = h.debug_nodes()
print("agreement", (lab == o["labels"]).mean())
gk = {parity.node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in gn}
ok = {parity.node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in o["nodes"]}
print("nodes gpu", len(gk), "oracle", len(ok), "shared", len(set(gk) & set(ok)))
bad = []
for k in sorted(set(gk) & set(ok)):
    a, b = gk[k], ok[k]
    if a["outcome"] != b["outcome"] or a["n_inliers"] != b["n_inliers"] or a["iters"] != b["iters"]:
        bad.append(k)
print("shared nodes that differ:", len(bad))
for k in bad[:8]:
    a, b = gk[k], ok[k]
    ang = float(np.arctan2(np.linalg.norm(np.cross(a["normal"], b["normal"])), abs(float(np.dot(a["normal"], b["normal"])))))
    print(k, "gpu: outcome", a["outcome"], "iters", a["iters"], "inl", a["n_inliers"], "res %.4f" % a["residual"],
          "| oracle: outcome", b["outcome"], "iters", b["iters"], "inl", b["n_inliers"], "res %.4f" % b["residual"], "| angle %.2e" % ang)
