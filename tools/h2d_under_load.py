"""Developer script (gpurun): time of one batch upload (64 pairs x 2 clouds x 2048 points: 2 x 16.8 MB features, 2 x 1.6 MB points)
from pinned host memory, split over 1 / 2 / 4 / 8 copy streams, on an idle GPU and while the registration kernels of another batch
run -- the e2e arm of bench.py is bounded by this when it exceeds the kernel time of a step."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
dev = torch.device("cuda:0")
B, N = 64, 2048
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=dev)
eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=True)
data = P.synthetic.make_batch(5, B, n=N)
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts")
host = {k: data[k].float().contiguous().pin_memory() for k in keys}
devt = {k: torch.empty_like(host[k], device=dev) for k in keys}
eng.register(*[data[k].to(dev) for k in keys], data["labels"].to(dev), data["gt_pose"].to(dev))
torch.cuda.synchronize()
total = sum(h.numel() * 4 for h in host.values())


def chunks(nsplit):
    """list of (dst, src) flat views: every tensor cut into nsplit pieces"""
    out = []
    for k in keys:
        h, d = host[k].view(-1), devt[k].view(-1)
        step = (h.numel() + nsplit - 1) // nsplit
        for i in range(0, h.numel(), step):
            out.append((d[i:i + step], h[i:i + step]))
    return out


def run(nstreams, nsplit, load, reps=30):
    ss = [torch.cuda.Stream() for _ in range(nstreams)]
    work = torch.cuda.Stream()
    cps = chunks(nsplit)
    cps.sort(key=lambda c: -c[1].numel())
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if load:
        with torch.cuda.stream(work):
            for _ in range(reps + 8):
                eng.run()
    cur = torch.cuda.current_stream()
    a.record()
    for _ in range(reps):
        for s in ss:
            s.wait_stream(cur)
        for i, (d, h) in enumerate(cps):
            with torch.cuda.stream(ss[i % nstreams]):
                d.copy_(h, non_blocking=True)
        for s in ss:
            cur.wait_stream(s)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    return ms


for load in (False, True):
    for nstreams, nsplit in ((1, 1), (2, 1), (4, 1), (4, 2), (8, 4), (2, 4), (1, 8)):
        ms = run(nstreams, nsplit, load)
        print(f"load={load!s:5} streams={nstreams} pieces/tensor={nsplit}: {ms:.3f} ms per batch = {total / ms / 1e6:.1f} GB/s")
