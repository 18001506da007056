"""A few single-scan calls of one shape through rpw_segment (the command ncu wraps for the single-scan profiles):
gpu_single_scan.py C2|C4|C5 [calls]"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
shape = sys.argv[1] if len(sys.argv) > 1 else "C4"
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 4
gen, cfg, seed = {"C2": (rpw.synth.spinning_scan, rpw.PatchworkConfig(filtering_radius=80.0), 1003),
                  "C4": (rpw.synth.solidstate_merged, rpw.PatchworkConfig(), 2000),
                  "C5": (rpw.synth.dense_urban_scan, rpw.PatchworkConfig(filtering_radius=80.0), 3000)}[shape]
pts = np.ascontiguousarray(gen(seed)[:, :3])
h = rpw.Handle(cfg.to_c(), 0, len(pts) + 4096, 1)
for _ in range(calls):
    lab = h.segment(pts)
print(shape, len(pts), "points, ground", int((lab == 1).sum()), "graph (launches, captures)", h.scan_graph())
