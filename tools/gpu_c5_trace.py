import importlib, sys
from concurrent.futures import ThreadPoolExecutor
import numpy as np
sys.path.insert(0, "/root/repo")
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
with ThreadPoolExecutor(16) as ex:
    scans = list(ex.map(lambda s: rpw.synth.dense_urban_scan(s), range(3100, 3228)))
B = len(scans); off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans]); total = int(off[-1])
d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, B)
for _ in range(3): h.segment_device(d.data_ptr(), off, lab.data_ptr())
torch.cuda.synchronize()
h.fit_trace_arm(1 << 18); h.segment_device(d.data_ptr(), off, lab.data_ptr()); torch.cuda.synchronize()
tr, seen = h.fit_trace_read(1 << 18)
t0 = tr["t_start_ns"].min(); s = (tr["t_start_ns"] - t0) / 1e3; e = (tr["t_end_ns"] - t0) / 1e3
print("nodes", len(tr), "makespan", e.max())
for dep in sorted(set(tr["depth"])):
    m = tr["depth"] == dep
    print(f"depth {dep}: nodes {m.sum()} mean n {tr['n'][m].mean():.0f} mean it {tr['iters'][m].mean():.1f} mean us {(e-s)[m].mean():.1f} start {s[m].min():.0f} end {e[m].max():.0f}")
