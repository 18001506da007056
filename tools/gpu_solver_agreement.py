"""Label agreement between the plane solvers over a batch of C2 scans (and the C5 stress scans)."""
import importlib, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
def run(scans, name):
    off = np.zeros(len(scans) + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
    total = int(off[-1])
    d = torch.from_numpy(np.concatenate(scans)).cuda()
    h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, len(scans))
    out = {}
    for sid, nm in ((0, "eigen_qr"), (1, "closed_form"), (2, "hybrid")):
        lab = torch.empty(total, dtype=torch.uint8, device="cuda")
        h.set_plane_solver(sid)
        h.segment_device(d.data_ptr(), off, lab.data_ptr())
        torch.cuda.synchronize()
        out[nm] = lab
    for nm in ("closed_form", "hybrid"):
        diff = (out[nm] != out["eigen_qr"])
        per_scan = [int(diff[int(off[b]):int(off[b + 1])].sum()) for b in range(len(scans))]
        print(f"{name}: {nm} vs eigen_qr: {int(diff.sum())} of {total} labels differ ({1 - diff.float().mean().item():.7f} agreement), "
              f"scans with any difference {sum(1 for x in per_scan if x)}/{len(scans)}, worst scan {max(per_scan)}")
    h.close()
with ThreadPoolExecutor(16) as ex:
    c2 = list(ex.map(lambda s: rpw.synth.spinning_scan(s), range(1000, 1000 + B)))
run(c2, f"C2 x{B}")
with ThreadPoolExecutor(16) as ex:
    c5 = list(ex.map(lambda s: rpw.synth.spinning_scan(s, 128, 2048, 1), range(3000, 3000 + 64)))
run(c5, "C5 x64")
