"""Small training step for compute-sanitizer (memcheck / racecheck / initcheck):
   compute-sanitizer --tool racecheck python tools/sanitize_backward.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P  # noqa: E402

dev = "cuda:0"
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=dev, variant="train")
with torch.no_grad():
    model.egnn.embedding_out.weight.mul_(0.005); model.egnn.embedding_out.bias.mul_(0.005)
d = {k: v.to(dev) for k, v in P.synthetic.make_batch(3, 2, n=300).items()}
es, et = P.knn_graph_batch(d["src_pts"], 16), P.knn_graph_batch(d["tgt_pts"], 16)
ones = torch.ones(2, 300 * 16, 1, device=dev)
opt = torch.optim.Adam(model.parameters(), lr=1e-5)
loss = P.train.train_step(model, opt, (d["src_feat"], d["src_pts"], es, ones, d["tgt_feat"], d["tgt_pts"], et, ones,
                                       d["corr"], d["labels"], d["gt_pose"]))
torch.cuda.synchronize()
print("loss", float(loss))
