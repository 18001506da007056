"""Cycle accounting of the fit kernel on a batch of C2 scans: python tools/gpu_timing.py [scans]"""
import importlib, sys, time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
with ThreadPoolExecutor(8) as ex:
    scans = list(ex.map(lambda s: rpw.synth.spinning_scan(s), range(1000, 1000 + B)))
off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
total = int(off[-1])
h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, B)
if len(sys.argv) > 2: h.set_plane_solver(int(sys.argv[2]))
st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
for _ in range(3): h.segment_device(d.data_ptr(), off, lab.data_ptr())
torch.cuda.synchronize()
h.fit_timing(True)
h.profile_enable(True)
R = 5
for _ in range(R): h.segment_device(d.data_ptr(), off, lab.data_ptr())
torch.cuda.synchronize()
t = h.fit_timing(False); p = h.profile_read()
blocks = p["fit_grid_blocks"]
fit_ms = p["fit"]["ms"] / R
print(f"scans={B} fit_ms={fit_ms:.3f} bin={p['bin']['ms']/R:.3f} scatter={p['scatter']['ms']/R:.3f} blocks={blocks}")
tot_cyc = sum(v for k, v in t.items() if k not in ("nodes", "iters"))
print("nodes/launch", t["nodes"] / R, "iters/node %.2f" % (t["iters"] / max(1, t["nodes"])))
for k, v in t.items():
    if k in ("nodes", "iters"): continue
    print(f"  {k:9s} {100*v/tot_cyc:5.1f}%  {v/R/blocks/1.965e6:8.4f} ms/block-avg   per node {v/max(1,t['nodes']):9.0f} cyc")
print("  sum per block-avg ms", tot_cyc / R / blocks / 1.965e6)
