"""Developer script: eager training steps of a randomly initialised model with a given head count; finiteness / checksum of the
forward outputs before and after the backward pass, finiteness of every gradient."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
DEV = "cuda:0"
heads = int(sys.argv[1]) if len(sys.argv) > 1 else 2
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
batches = [tuple(P.synthetic.make_batch(70 + i, 2, n=384)[k].to(DEV) for k in keys) for i in range(2)]
ones = torch.ones(2, 384 * 16, 1, device=DEV)
torch.manual_seed(11)
model = P.build_model(None, device=DEV, variant="train", num_heads=heads)
with torch.no_grad():
    model.egnn.embedding_out.weight.mul_(0.05); model.egnn.embedding_out.bias.mul_(0.05)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)


def sig(out):
    torch.cuda.synchronize()
    return ["%s%.6g" % ("" if bool(torch.isfinite(o).all()) else "NONFINITE ", float(o.detach().double().abs().sum())) for o in out if torch.is_tensor(o)]


for i in range(6):
    sf, sp, tf, tp, corr, labels, gt = batches[i % 2]
    es, et = P.knn_graph_batch(sp, 16), P.knn_graph_batch(tp, 16)
    model.train(); opt.zero_grad(set_to_none=True)
    pfin = all(bool(torch.isfinite(p).all()) for p in model.parameters())
    out = model(sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt)
    s0 = sig(out)
    loss = P.train.training_loss(out, gt)
    l0 = float(loss.detach())
    loss.backward()
    s1 = sig(out)
    bad = [k for k, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print(i, "params finite", pfin, "loss %.7f" % l0, "bad grads", len(bad), bad[:3])
    if s0 != s1 or bad or any("NONFINITE" in s for s in s0):
        print("   before backward:", s0)
        print("   after  backward:", s1)
        print("   R", out[0].detach().cpu().numpy().round(4).tolist(), "t", out[1].detach().cpu().numpy().round(4).tolist())
        aux = model.last_aux
        print("   w finite", bool(torch.isfinite(aux["w"]).all()), "w sum", aux["w"].sum(-1).tolist(), "H", aux["H"].cpu().numpy().tolist())
        print("   labels per pair", labels.reshape(2, -1).sum(-1).tolist())
        break
    opt.step()
