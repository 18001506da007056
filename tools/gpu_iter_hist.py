"""Histogram of plane-fit iterations per node (by size class) over a batch of C2 scans, and how much of the
point-iterations the long runners hold.  usage: gpu_iter_hist.py [scans] [shape C2|C4|C5]"""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
shape = sys.argv[2] if len(sys.argv) > 2 else "C2"
gen = {"C2": lambda s: rpw.synth.spinning_scan(s), "C4": lambda s: rpw.synth.solidstate_merged(s), "C5": lambda s: rpw.synth.dense_urban_scan(s)}[shape]
cfg = rpw.PatchworkConfig() if shape == "C4" else rpw.PatchworkConfig(filtering_radius=80.0)
base = {"C2": 1000, "C4": 2000, "C5": 3000}[shape]
with ThreadPoolExecutor(16) as ex:
    scans = list(ex.map(gen, range(base, base + B)))
h = rpw.Handle(cfg.to_c(), 0, sum(len(s) for s in scans) + 4096, B)
h.enable_nodes(True)
h.segment_batch(scans)
nd = h.debug_nodes()
fit = nd[(nd["outcome"] == 4) | (nd["outcome"] == 5)]
print(f"{shape}: scans {B}, nodes {len(nd)}, fitted {len(fit)}, max_iter hits {(fit['iters'] >= 100).sum()}")
edges = [0, 1024, 2048, 3072, 4096, 5632, 8192, 1 << 30]
bins = [1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 99, 100, 101]
print("class   nodes  mean_it  share_of_point_iters | nodes by iterations " + " ".join(f"<{b}" for b in bins[1:]))
tot_pi = float((fit["n"].astype(np.float64) * (fit["iters"] + 1)).sum())
for c in range(7):
    m = (fit["n"] > edges[c]) & (fit["n"] <= edges[c + 1])
    if not m.any():
        continue
    f = fit[m]
    pi = (f["n"].astype(np.float64) * (f["iters"] + 1))
    hist = np.histogram(f["iters"], bins)[0]
    print(f"{c:5d} {m.sum():7d} {f['iters'].mean():8.2f} {pi.sum() / tot_pi:8.3f}              | " + " ".join(f"{x:4d}" for x in hist))
    long = f["iters"] > 8
    print(f"        iters>8: {long.sum()} nodes ({long.mean():.3f}), point-iterations share within class {pi[long].sum() / pi.sum():.3f}; "
          f"iters>=100: {(f['iters'] >= 100).sum()} nodes, share {pi[f['iters'] >= 100].sum() / pi.sum():.3f}")
long = fit["iters"] > 8
pi = fit["n"].astype(np.float64) * (fit["iters"] + 1)
print(f"all: iters>8 {long.sum()} of {len(fit)} nodes; their share of point-iterations {pi[long].sum() / pi.sum():.3f}; "
      f"iters>=100 share {pi[fit['iters'] >= 100].sum() / pi.sum():.3f}")
print("exact-order replay estimate (8 cycles per point and iteration, one warp per node): total "
      f"{(fit['n'][long].astype(np.float64) * fit['iters'][long] * 8).sum() / 1.965e9 * 1e3:.1f} warp-ms, longest node "
      f"{(fit['n'][long].astype(np.float64) * fit['iters'][long] * 8).max() / 1.965e9 * 1e3:.2f} ms")
