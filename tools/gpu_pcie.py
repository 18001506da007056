"""Host->device copy bandwidth of this box (pinned memory, large transfers): the ceiling of bench.py's e2e arm,
which moves 1.44 MB per scan in and 0.12 MB out."""
import torch, time
n = 736 * 1024 * 1024
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
back = torch.empty(n // 12, dtype=torch.uint8).pin_memory()
dsrc = torch.empty(n // 12, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
for both in (False, True):
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    R = 10
    for _ in range(R):
        with torch.cuda.stream(s1):
            dst.copy_(src, non_blocking=True)
        if both:
            with torch.cuda.stream(s2):
                back.copy_(dsrc, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / R
    print(f"H2D {n / dt / 1e9:.1f} GB/s" + (" with a concurrent D2H of 1/12 the size" if both else "") + f" -> at most {n / dt / 1.44e6:.0f} scans/s of 120k xyz points")
