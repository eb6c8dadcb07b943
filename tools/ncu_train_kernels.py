"""Developer script (gpurun, under ncu): a few eager training steps (16 pairs x 2048) so that every kernel of the step can be
captured once with `ncu -k regex:... --launch-skip N`."""
import os, sys, torch
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
batch = tuple(v.to(dev) for v in (P.synthetic.make_batch(500, 16, n=2048)[k] for k in keys))
step = P.train.GraphedTrainStep(model, opt, batch, k=16, warmup=1)
for _ in range(2):
    step._step()
torch.cuda.synchronize()
