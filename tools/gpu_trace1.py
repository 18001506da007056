"""Timeline of ONE scan's fit (single-scan call: latency class table, clusters): gpu_trace1.py C2|C4|C5 [seed]"""
import importlib, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
shape = sys.argv[1] if len(sys.argv) > 1 else "C5"
gen, cfg, base = {"C2": (rpw.synth.spinning_scan, rpw.PatchworkConfig(filtering_radius=80.0), 1000),
                  "C4": (rpw.synth.solidstate_merged, rpw.PatchworkConfig(), 2000),
                  "C5": (rpw.synth.dense_urban_scan, rpw.PatchworkConfig(filtering_radius=80.0), 3000)}[shape]
seed = int(sys.argv[2]) if len(sys.argv) > 2 else base
pts = gen(seed)
h = rpw.Handle(cfg.to_c(), 0, len(pts) + 4096, 1)
for _ in range(3): h.segment(pts)
h.fit_trace_arm(1 << 16)
h.segment(pts)
tr, seen = h.fit_trace_read(1 << 16)
t0 = tr["t_start_ns"].min()
s = (tr["t_start_ns"] - t0) / 1e3; e = (tr["t_end_ns"] - t0) / 1e3
print(f"{shape} seed {seed}: {len(pts)} points, {seen} nodes, fit makespan {e.max():.1f} us")
for d in sorted(set(tr["depth"])):
    m = tr["depth"] == d
    print(f"  depth {d}: {m.sum():4d} nodes, start {s[m].min():7.1f} .. end {e[m].max():7.1f} us, mean n {tr['n'][m].mean():7.0f}, mean iters {tr['iters'][m].mean():5.1f}, longest {(e - s)[m].max():6.1f} us")
print("  latest-ending nodes: depth class n iters start dur")
for i in np.argsort(-e)[:10]:
    print(f"   {tr['depth'][i]:3d} {tr['size_class'][i]:6d} {tr['n'][i]:6d} {tr['iters'][i]:4d} {s[i]:8.1f} {(e - s)[i]:8.1f}")
