"""Device time per batch vs. batch size (QR solver): the slope is the steady-state cost per scan,
the intercept the fixed part (launch chain + fit tail).  usage: gpu_batchscale.py [solver]"""
import importlib, os, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
solver = int(sys.argv[1]) if len(sys.argv) > 1 else 0
U = 512
with ThreadPoolExecutor(16) as ex:
    uniq = list(ex.map(lambda s: rpw.synth.spinning_scan(s), range(1000, 1000 + U)))
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
rows = []
for B in (32, 64, 128, 256, 512, 1024, 2048):
    scans = [uniq[i % U] for i in range(B)]
    off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
    total = int(off[-1])
    d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
    h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, B)
    h.set_stream(st.cuda_stream); h.set_plane_solver(solver)
    for _ in range(3): h.segment_device(d.data_ptr(), off, lab.data_ptr())
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    R = 10
    e0.record(st)
    for _ in range(R): h.segment_device(d.data_ptr(), off, lab.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / R
    h.profile_enable(True) if hasattr(h, "profile_enable") else None
    rows.append((B, ms))
    print(f"B={B:5d} ms/batch={ms:.3f} scans/s={B/ms*1e3:.0f}", flush=True)
    h.close(); del d, lab
x = np.array([r[0] for r in rows], float); y = np.array([r[1] for r in rows])
A = np.vstack([x, np.ones_like(x)]).T
slope, icpt = np.linalg.lstsq(A[3:], y[3:], rcond=None)[0]
print(f"fit over B>=256: {slope*1e3:.3f} us/scan (= {1e3/slope:.0f} scans/s steady) + {icpt:.3f} ms fixed")
