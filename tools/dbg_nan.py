"""Developer script: eager training steps with every backward op wrapped: the first op whose outputs are not finite is re-run on the
same inputs several times (race or arithmetic?) and its inputs are saved to gpurun_out/."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
DEV = "cuda:0"
heads = int(sys.argv[1]) if len(sys.argv) > 1 else 4
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
batches = [tuple(P.synthetic.make_batch(70 + i, 2, n=384)[k].to(DEV) for k in keys) for i in range(2)]
ones = torch.ones(2, 384 * 16, 1, device=DEV)
torch.manual_seed(11)
model = P.build_model(None, device=DEV, variant="train", num_heads=heads)
with torch.no_grad():
    model.egnn.embedding_out.weight.mul_(0.05); model.egnn.embedding_out.bias.mul_(0.05)
opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
hit = []


def flat(o):
    if torch.is_tensor(o):
        return [o]
    if isinstance(o, (list, tuple)):
        return [t for x in o for t in flat(x)]
    if isinstance(o, dict):
        return [t for x in o.values() for t in flat(x)]
    return []


def fin(o):
    return all(bool(torch.isfinite(t).all()) for t in flat(o) if t.is_floating_point())


def wrap(name):
    orig = getattr(ops, name)

    def f(*a, **kw):
        torch.cuda.synchronize()
        ins_ok = fin(a) and fin(kw)
        a0 = [t.clone() if torch.is_tensor(t) else t for t in a]      # gp arguments are accumulated into
        out = orig(*a, **kw)
        torch.cuda.synchronize()
        if not fin(out) and not hit:
            hit.append(name)
            print("   FIRST non-finite output:", name, "inputs finite:", ins_ok, "which outputs:", [bool(torch.isfinite(t).all()) for t in flat(out) if t.is_floating_point()])
            if ins_ok:
                again = []
                for _ in range(6):
                    a1 = [t.clone() if torch.is_tensor(t) else t for t in a0]
                    o2 = orig(*a1, **kw); torch.cuda.synchronize()
                    again.append(fin(o2))
                print("   re-run on the same inputs, finite:", again)
                os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
                torch.save({"name": name, "args": [(t.cpu() if torch.is_tensor(t) else (t if not hasattr(t, "ptr") else None)) for t in a0],
                            "kw": {k: (v.cpu() if torch.is_tensor(v) else None) for k, v in kw.items()}},
                           os.path.join(ROOT, "gpurun_out", "dbg_nan_%s.pt" % name))
        return out
    setattr(ops, name, f)


for nm in ("head_train_backward", "head_train_loss_backward", "linear32_backward", "egnn_backward", "train_loss_finalize", "pose_loss"):
    wrap(nm)

for i in range(6):
    sf, sp, tf, tp, corr, labels, gt = batches[i % 2]
    es, et = P.knn_graph_batch(sp, 16), P.knn_graph_batch(tp, 16)
    model.train(); opt.zero_grad(set_to_none=True)
    out = model(sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt)
    loss = P.train.training_loss(out, gt)
    loss.backward()
    bad = [k for k, p in model.named_parameters() if p.grad is not None and not torch.isfinite(p.grad).all()]
    print(i, "loss %.7f" % float(loss.detach()), "bad grads", len(bad))
    if bad:
        break
    opt.step()
