"""Times the EGNN backward (3 layers + embeddings) at the training bench shape with CUDA events, and each kernel of one
layer's backward alone.  python tools/time_backward.py [pairs]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P  # noqa: E402
from se3_equi_graph_registration_b200 import ops, packing  # noqa: E402

dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=dev, variant="train")
d = {k: v.to(dev) for k, v in P.synthetic.make_batch(3, B, n=2048).items()}
x = torch.cat([d["src_pts"], d["tgt_pts"]]); f = torch.cat([d["src_feat"], d["tgt_feat"]])
nbr = ops.knn_build(x, 16)
graph = ops.with_csc(ops.csr_from_nbr(nbr))
layers, pin, pout = model.egnn.packs()
h, xo, saved = ops.egnn_forward_saved(f, x, graph, layers, pin, pout)
dh, dx = torch.randn_like(h), torch.randn_like(xo)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    for _ in range(2):
        fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


print(f"{2 * B} clouds x 2048 pts: egnn forward (saved) {timeit(lambda: ops.egnn_forward_saved(f, x, graph, layers, pin, pout)):.3f} ms, "
      f"egnn backward {timeit(lambda: ops.egnn_backward(saved, graph, layers, pin, pout, dh, dx)):.3f} ms")
