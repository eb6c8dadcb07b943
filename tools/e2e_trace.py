"""Developer script (gpurun): CUPTI timeline (torch.profiler) of the host-to-host loop PipelinedEngine.submit()/collect() -- which
copies and kernels of which lane overlap.  Writes gpurun_out/e2e_trace_l<lanes>.json (compact event list)."""
import json, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda:0")
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 2
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 16
B, N = 64, 2048
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=dev)
pipe = P.PipelinedEngine(model, batch=B, n=N, k=16, device=dev, lanes=lanes, use_graph=True)
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")
host = []
for i in range(4):
    d = P.synthetic.make_batch(100 + i, B, n=N)
    host.append({k: d[k].contiguous().pin_memory() for k in keys})


def loop(n):
    tks = [pipe.submit(*[host[i % 4][k] for k in keys]) for i in range(min(lanes, n))]
    for i in range(n):
        pipe.collect(tks[i])
        if i + lanes < n:
            tks.append(pipe.submit(*[host[(i + lanes) % 4][k] for k in keys]))


loop(8)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    loop(steps)
    torch.cuda.synchronize()
ev = []
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA:
        ev.append((e.time_range.start, e.time_range.end, e.name[:60], getattr(e, "device_index", 0)))
# kineto events carry the stream in the chrome trace only: export and re-read
path = os.path.join(ROOT, "gpurun_out", "e2e_trace_l%d_full.json" % lanes)
os.makedirs(os.path.dirname(path), exist_ok=True)
prof.export_chrome_trace(path)
tr = json.load(open(path))
out = []
for e in tr["traceEvents"]:
    if e.get("ph") == "X" and e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset"):
        out.append({"t": e["ts"], "d": e["dur"], "n": e["name"][:48], "s": e.get("args", {}).get("stream"), "c": e["cat"],
                    "b": e.get("args", {}).get("bytes")})
    elif e.get("ph") == "X" and e.get("cat") in ("cuda_runtime",) and e["name"] in ("cudaEventSynchronize", "cudaGraphLaunch"):
        out.append({"t": e["ts"], "d": e["dur"], "n": e["name"], "s": "host", "c": "rt"})
out.sort(key=lambda x: x["t"])
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "e2e_trace_l%d.json" % lanes), "w"))
os.remove(path)
print("events", len(out))
