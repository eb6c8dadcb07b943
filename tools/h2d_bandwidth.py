"""Developer script (gpurun): host->device copy rate of this box from pinned memory, at the bench's per-step upload size
and as one large copy; the e2e arm of bench.py is bounded by it (37.2 MB per 64-pair step)."""
import torch
dev = torch.device("cuda:0")
for mb in (0.5, 4, 37.2, 256):
    n = int(mb * 1e6)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    a.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"H2D {mb:7.1f} MB: {ms:.3f} ms  {n / ms / 1e6:.1f} GB/s")
# several copy streams at once
for ns in (2, 4):
    n = int(37.2e6 / ns)
    hs = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(ns)]
    ds = [torch.empty(n, dtype=torch.uint8, device=dev) for _ in range(ns)]
    ss = [torch.cuda.Stream() for _ in range(ns)]
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        for s, h, d in zip(ss, hs, ds):
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                d.copy_(h, non_blocking=True)
    for s in ss:
        torch.cuda.current_stream().wait_stream(s)
    b.record(); torch.cuda.synchronize()
    print(f"H2D {ns} x {n / 1e6:.1f} MB on {ns} streams: {ns * n * 20 / a.elapsed_time(b) / 1e6:.1f} GB/s")
import subprocess
print(subprocess.run("nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current --format=csv", shell=True, capture_output=True, text=True).stdout)
