"""Golden outputs and gradients of the reference's EGNN built with num_heads in {1, 2, 8} (E_GCL(num_heads=...),
src/3dmatch_train_egnn_with_batch.py:186-207: `hidden_nf // num_heads` wide heads, concatenated at :246), produced by the
REFERENCE'S OWN classes and autograd (ast-extracted from /root/reference by oracle/ref_loader.py).  Build container only.

    python tests/golden/make_golden_heads.py        -> tests/golden/heads_{1,2,8}.pt

The shipped checkpoints have 4 heads, so these models are randomly initialised (seeded; coord_mlp's last layer scaled up
from its 1e-3 xavier gain so that the coordinate path carries signal).  Stored per head count: the EGNN state_dict
(fp32), a 96-node graph with duplicate points, outputs (h, x) and the gradients of a fixed linear functional of the
outputs w.r.t. h, x and every parameter, all from an fp64 run rounded to fp32.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.egnn_oracle import edges_from_nbr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def make(heads):
    g = torch.Generator().manual_seed(100 + heads)
    torch.manual_seed(200 + heads)
    ns, egnn, _ = ref_loader.build_reference_model("train", checkpoint=None, num_heads=heads, n_layers=2)
    with torch.no_grad():
        for i in range(2):
            getattr(egnn, "gcl_%d" % i).coord_mlp[2].weight.mul_(300.0)
    sd32 = {k: v.detach().clone().to(torch.float32) for k, v in egnn.state_dict().items()}
    egnn = egnn.double()
    n, k = 96, 8
    x = torch.rand(n, 3, generator=g) * 1.5
    x[n // 2:n // 2 + 12] = x[:12]                                    # duplicate points: zero-length edges, identity frames
    d2 = ((x[:, None] - x[None]) ** 2).sum(-1)
    nbr = d2.argsort(dim=1, stable=True)[:, :k]
    row, col = edges_from_nbr(nbr)
    h = torch.randn(n, 32, generator=g) * 0.5
    ea = torch.rand(row.numel(), 1, generator=g) + 0.5
    dh = torch.randn(n, 32, generator=g)
    dx = torch.randn(n, 3, generator=g)
    h64, x64 = h.double().requires_grad_(True), x.double().requires_grad_(True)
    torch.set_default_dtype(torch.float64)                            # the reference builds its identity frames in the default dtype
    try:
        ho, xo = egnn(h64, x64, [row, col], ea.double())              # EGNN.forward 3dm:328-340
        ((ho * dh.double()).sum() + (xo * dx.double()).sum()).backward()
    finally:
        torch.set_default_dtype(torch.float32)
    f32 = lambda t: t.detach().to(torch.float32)
    return {"heads": heads, "state_dict": sd32, "h": h, "x": x, "row": row, "col": col, "edge_attr": ea, "dh": dh, "dx": dx,
            "h_out": f32(ho), "x_out": f32(xo), "grad_h": f32(h64.grad), "grad_x": f32(x64.grad),
            "grads": {k_: f32(p.grad) for k_, p in egnn.named_parameters()}}


if __name__ == "__main__":
    for heads in (1, 2, 8):
        out = make(heads)
        path = os.path.join(HERE, "heads_%d.pt" % heads)
        torch.save(out, path)
        print(path, os.path.getsize(path), "bytes; |h_out| max", float(out["h_out"].abs().max()),
              "|x_out - x| max", float((out["x_out"] - out["x"]).abs().max()))
