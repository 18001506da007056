"""Total device time per 512-scan batch vs. number of launch groups (profile/timing off)."""
import importlib, os, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
with ThreadPoolExecutor(16) as ex:
    scans = list(ex.map(lambda s: rpw.synth.spinning_scan(s), range(1000, 1000 + B)))
off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
total = int(off[-1])
d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for solver in (2,):
    for waves in (1, 2, 3, 4, 8):
        os.environ["RPW_WAVES"] = str(waves)
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
        print(f"solver={solver} waves={waves:2d} ms/batch={ms:.3f} scans/s={B/ms*1e3:.0f}")
        h.close()
