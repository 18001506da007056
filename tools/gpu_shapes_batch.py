"""Batched device-resident throughput of the other shapes (C4 merged solid-state frames, C5 dense urban scans)."""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
for name, cfg, gen, seeds in (("C4 300k", rpw.PatchworkConfig(), lambda s: rpw.synth.solidstate_merged(s), range(2100, 2164)),
                              ("C5 262k", rpw.PatchworkConfig(filtering_radius=80.0), lambda s: rpw.synth.dense_urban_scan(s), range(3100, 3228))):
    with ThreadPoolExecutor(16) as ex:
        scans = list(ex.map(gen, seeds))
    B = len(scans)
    off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
    total = int(off[-1])
    d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
    h = rpw.Handle(cfg.to_c(), 0, total, B); h.set_stream(st.cuda_stream)
    for _ in range(3): h.segment_device(d.data_ptr(), off, lab.data_ptr())
    torch.cuda.synchronize()
    h.profile_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(5): h.segment_device(d.data_ptr(), off, lab.data_ptr())
    e1.record(st); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    p = h.profile_read()
    print(f"{name}: {B} scans, {total/1e6:.1f} M points: {ms:.3f} ms/batch = {B/ms*1e3:.0f} scans/s = {total/ms/1e6:.2f} G points/s | "
          + " ".join(f"{k} {p[k]['ms']/max(1,p[k]['launches']):.3f}" for k in ("bin", "offsets", "scatter", "fit")))
    h.close(); del d, lab
