"""Host topology of the GPU box and what it does to pinned host->device copies: prints the CPU / NUMA layout, the
GPU's CPU affinity as NVML reports it, and the H2D bandwidth of pinned memory first touched by a thread bound to
each NUMA node in turn (bench.py's e2e arm is bound by exactly this copy).

    python tools/gpu_numa.py [device]
"""
import glob, os, subprocess, sys, time
import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        out.extend(range(int(a), int(b or a) + 1))
    return out


dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
print(subprocess.run("lscpu | egrep 'Model name|Socket|NUMA|^CPU\\(s\\)'; nvidia-smi topo -m", shell=True, capture_output=True, text=True).stdout)
nodes = {}
for p in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    nodes[int(p.rsplit("node", 1)[1])] = cpulist(open(p + "/cpulist").read())
allowed = sorted(os.sched_getaffinity(0))
print("allowed cpus:", len(allowed), "numa nodes:", {k: len(v) for k, v in nodes.items()})
try:
    import pynvml
    pynvml.nvmlInit()
    hdl = pynvml.nvmlDeviceGetHandleByIndex(dev)
    words = pynvml.nvmlDeviceGetCpuAffinity(hdl, (os.cpu_count() + 63) // 64)
    aff = [i * 64 + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    print("NVML cpu affinity of GPU", dev, ":", aff[:4], "...", aff[-4:], f"({len(aff)} cpus)")
    try:
        print("NVML numa node id:", pynvml.nvmlDeviceGetNumaNodeId(hdl))
    except Exception as e:
        print("nvmlDeviceGetNumaNodeId:", e)
except Exception as e:
    print("pynvml:", e)

torch.cuda.set_device(dev)
n = 512 * 1024 * 1024
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for node, cpus in nodes.items():
    use = [c for c in cpus if c in allowed]
    if not use:
        print(f"node {node}: no allowed cpus"); continue
    os.sched_setaffinity(0, use)
    src = torch.empty(n, dtype=torch.uint8)
    src.fill_(1)  # first touch on this node
    src = src.pin_memory()
    src.fill_(2)
    for _ in range(2): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter(); R = 8
    for _ in range(R): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / R
    print(f"pinned memory touched on node {node}: H2D {n / dt / 1e9:.1f} GB/s")
    del src
os.sched_setaffinity(0, allowed)
