"""One resident batch of C2 scans through rpw_segment_device: three warm-up steps, then one step -- the command line
ncu wraps (skip 3 x 11 launches, capture 11).  usage: gpu_prof_step.py [scans] [solver]"""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
with ThreadPoolExecutor(16) as ex:
    scans = list(ex.map(lambda s: rpw.synth.spinning_scan(s), range(1000, 1000 + B)))
off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
total = int(off[-1])
h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, B)
if len(sys.argv) > 2: h.set_plane_solver(int(sys.argv[2]))
d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
for _ in range(4):
    h.segment_device(d.data_ptr(), off, lab.data_ptr())
    torch.cuda.synchronize()
print("ground fraction", float((lab == 1).float().mean()), "launches", h.kernel_launches())
