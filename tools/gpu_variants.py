"""Kernel experiments: times the resident 512-scan C2 batch with every library variant given on the command line
(built into ros2-recursive-patchwork-implementation_b200/_variants/ with different -D flags) and prints a checksum
of the labels, which must not depend on the variant.

    python tools/gpu_variants.py [C2|C4|C5] [scans] name1 name2 ...   (names of _variants/*.so; 'default' = the shipped library)
"""
import hashlib, importlib, os, subprocess, sys, tempfile
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def child(path, B):
    rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
    import torch
    data = np.load(path)
    pts, off = data["pts"], data["off"]
    total = int(off[-1])
    h = rpw.Handle(rpw.synth.config_for(os.environ.get("RPW_VARIANT_SHAPE", "C2")).to_c(), 0, total, B)
    st = torch.cuda.Stream(); torch.cuda.set_stream(st); h.set_stream(st.cuda_stream)
    d = torch.from_numpy(pts).cuda(); lab = torch.empty(total, dtype=torch.uint8, device="cuda")
    for _ in range(5): h.segment_device(d.data_ptr(), off, lab.data_ptr())
    torch.cuda.synchronize()
    best = []
    for rep in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        R = 20
        e0.record(st)
        for _ in range(R): h.segment_device(d.data_ptr(), off, lab.data_ptr())
        e1.record(st); torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / R)
    h.profile_enable(True)
    R = 10
    for _ in range(R): h.segment_device(d.data_ptr(), off, lab.data_ptr())
    torch.cuda.synchronize()
    p = h.profile_read()
    labels = lab.cpu().numpy()
    digest = hashlib.sha1(labels.tobytes()).hexdigest()[:12]
    first = path + ".labels.npy"
    if os.path.exists(first):
        ndiff = int((np.load(first) != labels).sum())
    else:
        np.save(first, labels); ndiff = 0
    print(f"{os.environ.get('RPW_VARIANT_LABEL', 'default'):32s} step ms {min(best):.3f} (runs {' '.join('%.3f' % b for b in best)})  "
          f"{B / min(best):.1f} k scans/s  bin {p['bin']['ms']/R:.3f} scatter {p['scatter']['ms']/R:.3f} fit {p['fit']['ms']/R:.3f}  labels {digest} ({ndiff} differ from the first variant)", flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]))
        sys.exit(0)
    args = sys.argv[1:]
    shape = args.pop(0) if args and args[0] in ("C2", "C4", "C5") else "C2"
    os.environ["RPW_VARIANT_SHAPE"] = shape
    B = 512 if shape == "C2" else 64
    if args and args[0].isdigit(): B = int(args.pop(0))
    rpw = importlib.import_module("ros2-recursive-patchwork-implementation_b200")
    gen = {"C2": lambda s: rpw.synth.spinning_scan(1000 + s), "C4": lambda s: rpw.synth.solidstate_merged(2000 + s),
           "C5": lambda s: rpw.synth.dense_urban_scan(3000 + s)}[shape]
    with ThreadPoolExecutor(16) as ex:
        scans = list(ex.map(gen, range(B)))
    off = np.zeros(B + 1, np.uint64); off[1:] = np.cumsum([len(s) for s in scans])
    path = os.path.join(tempfile.gettempdir(), "rpw_variant_batch.npz")
    np.savez(path, pts=np.concatenate(scans), off=off)
    if os.path.exists(path + ".labels.npy"): os.remove(path + ".labels.npy")
    for spec in args:
        # "name" or "name:VAR=value,VAR=value" (environment of the child: RPW_PLANE_SOLVER, RPW_EXACT_REPLAY, ...)
        name, _, envs = spec.partition(":")
        env = dict(os.environ)
        for kv in filter(None, envs.split(",")):
            k, _, v = kv.partition("=")
            env[k] = v
        env["RPW_VARIANT_LABEL"] = spec
        if name != "default":
            env["RPW_B200_LIB"] = str(ROOT / "ros2-recursive-patchwork-implementation_b200" / "_variants" / f"{name}.so")
        subprocess.run([sys.executable, __file__, "--child", path, str(B)], env=env, check=False)
