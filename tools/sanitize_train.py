"""Developer script for compute-sanitizer (memcheck / initcheck / racecheck): two eager training steps through the module
API and two steps of the lean launch sequence (GraphedTrainStep._step without capture) on a small batch."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import packing, ops
DEV = "cuda:0"
N = int(sys.argv[1]) if len(sys.argv) > 1 else 256
def mk():
    m = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=DEV, variant="train")
    with torch.no_grad():
        m.egnn.embedding_out.weight.mul_(0.005); m.egnn.embedding_out.bias.mul_(0.005)
    return m
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
batches = [tuple(P.synthetic.make_batch(40 + i, 2, n=N)[k].to(DEV) for k in keys) for i in range(2)]
ones = torch.ones(2, N * 16, 1, device=DEV)
m1 = mk(); o1 = torch.optim.Adam(m1.parameters(), lr=1e-4)
for i in range(2):
    sf, sp, tf, tp, corr, labels, gt = batches[i]
    es, et = P.knn_graph_batch(sp, 16), P.knn_graph_batch(tp, 16)
    print("eager", float(P.train.train_step(m1, o1, (sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt))))
m2 = mk(); o2 = torch.optim.Adam(m2.parameters(), lr=1e-4, capturable=True)
step = P.train.GraphedTrainStep(m2, o2, batches[0], k=16, warmup=1)
for i in range(2):          # the lean launch sequence outside its CUDA graph (side-stream kernels included)
    step.load(batches[i])
    ops.build_train_graph(step.x_all, 16, out=step.sets[step._cur]["graph"])
    print("lean", step._step().tolist()[:5])
print("graphed", float(step(batches[0], next_batch=batches[1])), float(step(batches[1])))
torch.cuda.synchronize()
# the inference path as well: engine (k-NN, CSR, embed, 3 layers, eval head incl. the split pre-pass) in fp32, TF32 and bf16 modes
eng = P.RegistrationEngine(P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=DEV), batch=2, n=2048, k=16, use_graph=False)
dd = P.synthetic.make_batch(11, 2, n=2048)
for impl in (0, 4, 5):
    eng.impl = impl
    R, t = eng.register(dd["src_feat"], dd["src_pts"], dd["tgt_feat"], dd["tgt_pts"], dd["labels"], dd["gt_pose"])
    torch.cuda.synchronize()
    print("engine impl", impl, float(R.sum()))
