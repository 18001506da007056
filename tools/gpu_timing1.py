"""Cycle accounting of ONE scan's fit (thread 0 of every block): gpu_timing1.py C2|C4|C5 [seed]"""
import importlib, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
shape = sys.argv[1] if len(sys.argv) > 1 else "C2"
gen, cfg, base = {"C2": (rpw.synth.spinning_scan, rpw.PatchworkConfig(filtering_radius=80.0), 1003),
                  "C4": (rpw.synth.solidstate_merged, rpw.PatchworkConfig(), 2000),
                  "C5": (rpw.synth.dense_urban_scan, rpw.PatchworkConfig(filtering_radius=80.0), 3000)}[shape]
pts = gen(int(sys.argv[2]) if len(sys.argv) > 2 else base)
h = rpw.Handle(cfg.to_c(), 0, len(pts) + 4096, 1)
for _ in range(3): h.segment(pts)
h.fit_timing(True)
R = 20
for _ in range(R): h.segment(pts)
t = h.fit_timing(False)
iters = t["iters"] / R
print(f"{shape}: nodes {t['nodes'] / R:.0f}, plane-fit iterations {iters:.0f} per scan; cycles per ITERATION (summed over nodes / iterations):")
for k in ("eig", "dist", "dist_reduce", "cov", "seeds", "load", "load_loop", "label", "final", "split", "fetch", "gridsync"):
    print(f"  {k:12s} {t[k] / R / max(1, iters):8.0f} cycles/iteration   ({t[k] / R / 1.965e3:8.1f} us summed per scan)")
