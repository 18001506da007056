"""For the scans of one shape: per-scan label agreement with the oracle and, for the worst scan, the nodes that differ."""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import oracle_lib, parity
shape = sys.argv[1] if len(sys.argv) > 1 else "C4"
solver = int(sys.argv[2]) if len(sys.argv) > 2 else 0
if shape == "C4":
    cfg, gen, seeds = rpw.PatchworkConfig(), (lambda s: rpw.synth.solidstate_merged(s)), range(2100, 2124)
elif shape == "C5":
    cfg, gen, seeds = rpw.PatchworkConfig(filtering_radius=80.0), (lambda s: rpw.synth.dense_urban_scan(s)), range(3100, 3148)
else:
    cfg, gen, seeds = rpw.PatchworkConfig(filtering_radius=80.0), (lambda s: rpw.synth.spinning_scan(s)), range(5000, 5192)
oracle = oracle_lib.Oracle(); ocfg = oracle_lib.to_cfg(cfg)
with ThreadPoolExecutor(16) as ex:
    scans = list(ex.map(gen, seeds))
    want = list(ex.map(lambda a: oracle.run(ocfg, a, want_nodes=True), scans))
h = rpw.Handle(cfg.to_c(), 0, max(len(a) for a in scans) + 4096, 1)
h.set_plane_solver(solver); h.enable_nodes(True)
res = []
for s, a, o in zip(seeds, scans, want):
    lab = h.segment(a)
    res.append((int((lab != o["labels"]).sum()), s))
res.sort(reverse=True)
print("scans", len(res), "with any difference", sum(1 for d, _ in res if d), "below 99.9 %:", sum(1 for (d, s), a in zip(res, scans) if d > 0.001 * len(a)))
print("worst:", res[:6])
d, s = res[0]
a = scans[list(seeds).index(s)]; o = want[list(seeds).index(s)]
lab = h.segment(a); gn = h.debug_nodes()
gk = {parity.node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in gn}
ok = {parity.node_key(r["root"], r["depth"], r["start"], r["n"]): r for r in o["nodes"]}
print("seed", s, "nodes gpu", len(gk), "oracle", len(ok), "shared", len(set(gk) & set(ok)))
for k in sorted(set(gk) & set(ok)):
    x, y = gk[k], ok[k]
    if x["outcome"] != y["outcome"] or x["n_inliers"] != y["n_inliers"] or x["iters"] != y["iters"]:
        ang = float(np.arctan2(np.linalg.norm(np.cross(x["normal"], y["normal"])), abs(float(np.dot(x["normal"], y["normal"])))))
        print(k, "gpu: out", x["outcome"], "it", x["iters"], "inl", x["n_inliers"], "res %.5f" % x["residual"],
              "| oracle: out", y["outcome"], "it", y["iters"], "inl", y["n_inliers"], "res %.5f" % y["residual"], "| angle %.2e" % ang)
