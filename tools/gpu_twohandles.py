"""Experiment: K steps alternating over H handles/streams vs one handle (device-resident, 512 scans/step)."""
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
d = torch.from_numpy(np.concatenate(scans)).cuda()
for H in (1, 2, 3):
    hs, sts, labs = [], [], []
    for k in range(H):
        h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, B)
        st = torch.cuda.Stream(); h.set_stream(st.cuda_stream)
        hs.append(h); sts.append(st); labs.append(torch.empty(total, dtype=torch.uint8, device="cuda"))
    for r in range(6):
        hs[r % H].segment_device(d.data_ptr(), off, labs[r % H].data_ptr())
    torch.cuda.synchronize()
    K = 20
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(torch.cuda.current_stream())
    for st in sts: st.wait_event(e0)
    for r in range(K):
        hs[r % H].segment_device(d.data_ptr(), off, labs[r % H].data_ptr())
    evs = []
    for st in sts:
        ev = torch.cuda.Event(); ev.record(st); torch.cuda.current_stream().wait_event(ev)
    e1.record(torch.cuda.current_stream()); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    same = all(torch.equal(labs[0], l) for l in labs)
    print(f"handles={H} ms/step={ms:.3f} scans/s={B/ms*1e3:.0f} labels equal across handles: {same}")
    for h in hs: h.close()
