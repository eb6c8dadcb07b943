"""Developer script: the graphed training step (16 pairs x 2048) -- replay only vs replay + input staging."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import se3_equi_graph_registration_b200 as P
dev = torch.device("cuda", 0)
model = P.build_model(bench.CKPT, device=dev, variant="train")
with torch.no_grad():
    model.egnn.embedding_out.weight.mul_(0.005); model.egnn.embedding_out.bias.mul_(0.005)
opt = torch.optim.Adam(model.parameters(), lr=1e-5, capturable=True, fused=True)
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
batches = [tuple(v.to(dev) for v in (P.synthetic.make_batch(500 + i, 16, n=2048)[k] for k in keys)) for i in range(4)]
step = P.train.GraphedTrainStep(model, opt, batches[0], k=16)
def timeit(fn, n=20):
    for _ in range(3): fn(0)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record()
    for i in range(n): fn(i)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) * 1e3 / n
print("replay only      : %.3f ms (events) %.3f ms (wall)" % timeit(lambda i: step.graph.replay()))
print("load + replay    : %.3f ms (events) %.3f ms (wall)" % timeit(lambda i: step(batches[i % 4])))
print("load + replay, next batch announced (its graph is built under this step): %.3f ms (events) %.3f ms (wall)" % timeit(lambda i: step(batches[i % 4], next_batch=batches[(i + 1) % 4])))
print("eager _step()    : %.3f ms (events) %.3f ms (wall)" % timeit(lambda i: step._step(), 5))
