import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
from oracle import egnn_oracle as O
dev = "cuda:0"
g = torch.load(os.path.join(ROOT, "tests/golden/dup_b2_n512.pt"), weights_only=False, map_location="cpu")
model = P.build_model(os.path.join(ROOT, "tests/golden/checkpoint-3dmatch.pth"), device=dev)
sd = {k: v.cpu() for k, v in model.egnn.state_dict().items()}
inp = g["inputs"]
for impl in (1, 3):
  for side in ("src", "tgt"):
    for b in range(2):
        nbr = g[f"nbr_{side}"][b:b+1]
        row, col = O.edges_from_nbr(nbr[0])
        feat, x = inp[f"{side}_feat"][b], inp[f"{side}_pts"][b]
        href, xref, lref = O.egnn_forward(sd, feat, x, row, col, torch.ones(row.shape[0], 1), return_layers=True)
        gr = ops.csr_from_nbr(nbr.to(dev))
        layers, pin, pout = model.egnn.packs()
        h, xo, lay = ops.egnn_forward(feat[None].to(dev), x[None].to(dev), gr, layers, pin, pout, impl=impl, return_layers=True)
        deg = torch.bincount(row, minlength=512)
        msg = f"impl {impl} {side}{b} maxdeg {int(deg.max())} "
        for i, (hl, xl) in enumerate(lay):
            eh = (hl[0].cpu() - lref[i][0]).abs().max(1).values
            ex = (xl[0].cpu() - lref[i][1]).abs().max(1).values
            w = int(eh.argmax())
            msg += f"| L{i} h {float(eh.max()):.2e}@{w}(deg {int(deg[w])}, |h| {float(lref[i][0][w].abs().max()):.1f}) x {float(ex.max()):.2e}@{int(ex.argmax())}(deg {int(deg[int(ex.argmax())])}) "
        eh = (h[0].cpu() - href).abs().max(1).values
        msg += f"| out h {float(eh.max()):.2e} x {float((xo[0].cpu()-xref).abs().max()):.2e}"
        print(msg)
