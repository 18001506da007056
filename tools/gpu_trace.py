"""Timeline of the fit kernels over one batch (rpw_debug_fit_trace): which class ran when, how many
nodes and how much shared memory were resident per SM over time, which nodes end last.
usage: gpu_trace.py [scans] [solver]   (writes gpurun_out/fit_trace.npz)"""
import importlib, os, sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
import torch
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
solver = int(sys.argv[2]) if len(sys.argv) > 2 else 0
with ThreadPoolExecutor(16) as ex:
    scans = list(ex.map(lambda s: rpw.synth.spinning_scan(s), range(1000, 1000 + B)))
off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
total = int(off[-1])
d = torch.from_numpy(np.concatenate(scans)).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream(); torch.cuda.set_stream(st)
h = rpw.Handle(rpw.PatchworkConfig(filtering_radius=80.0).to_c(), 0, total, B)
h.set_stream(st.cuda_stream); h.set_plane_solver(solver)
for _ in range(3): h.segment_device(d.data_ptr(), off, lab.data_ptr())
torch.cuda.synchronize()
h.fit_trace_arm(1 << 18)
h.segment_device(d.data_ptr(), off, lab.data_ptr())
torch.cuda.synchronize()
tr, seen = h.fit_trace_read(1 << 18)
print("nodes traced", len(tr), "seen", seen)
t0 = tr["t_start_ns"].min()
s = (tr["t_start_ns"] - t0) / 1e3; e = (tr["t_end_ns"] - t0) / 1e3  # us
dur = e - s
print(f"fit makespan {e.max():.1f} us")
caps = [1024, 2048, 3072, 4096, 5632, 8192, 256]
def kb(c): return caps[c] * 13 / 1024 if c < len(caps) else 0
print("class  nodes  mean_n  mean_it  mean_us  p99_us  first_start  last_start  last_end  KB*ms")
for c in sorted(set(tr["size_class"])):
    m = tr["size_class"] == c
    print(f"{c:5d} {m.sum():6d} {tr['n'][m].mean():7.0f} {tr['iters'][m].mean():8.2f} {dur[m].mean():8.1f} {np.quantile(dur[m], .99):7.1f} "
          f"{s[m].min():11.1f} {s[m].max():11.1f} {e[m].max():9.1f} {kb(c) * dur[m].sum() / 1e3:8.0f}")
# residency over time
step = 50.0
nb = int(e.max() / step) + 1
nsm = int(tr["sm"].max()) + 1
print(f"SMs {nsm}; time bins of {step:.0f} us: resident nodes per SM, resident KB per SM (class slots), per-class resident nodes per SM")
for b in range(nb):
    a, z = b * step, (b + 1) * step
    ov = np.clip(np.minimum(e, z) - np.maximum(s, a), 0, None) / step   # fraction of the bin each node is resident
    per_cls = [ov[tr["size_class"] == c].sum() / nsm for c in range(len(caps))]
    kbs = sum(per_cls[c] * kb(c) for c in range(len(caps)))
    print(f"{a:7.0f} us  nodes/SM {ov.sum() / nsm:5.2f}  KB/SM {kbs:6.1f}  " + " ".join(f"{x:4.2f}" for x in per_cls))
last = np.argsort(-e)[:12]
print("latest-ending nodes: class n iters start_us dur_us sm")
for i in last:
    print(f"   {tr['size_class'][i]:3d} {tr['n'][i]:6d} {tr['iters'][i]:4d} {s[i]:8.1f} {dur[i]:8.1f} {tr['sm'][i]:4d}")
os.makedirs(ROOT / "gpurun_out", exist_ok=True)
np.savez_compressed(ROOT / "gpurun_out" / f"fit_trace_{B}_{solver}.npz", trace=tr)
