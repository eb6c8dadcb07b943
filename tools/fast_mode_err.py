"""Developer script (gpurun): error of the reduced-precision edge modes (impl 4 = TF32, impl 5 = bf16) against the golden
fixtures: EGNN outputs (eval goldens), and -- on tempered weights, where the train-variant Kabsch is well conditioned -- the
train-variant pose / H against the fp32 path."""
import os, sys, math, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
G = os.path.join(ROOT, "tests", "golden")
def ang(Ra, Rb):
    tr = float((Ra.double().T @ Rb.double()).trace())
    return math.degrees(math.acos(max(-1.0, min(1.0, (tr - 1) / 2))))
for name in ["small_b2_n256", "dup_b2_n512", "full_b1_n2048", "kitti_b1_n1024", "noenc_b1_n512"]:
    g = torch.load(os.path.join(G, name + ".pt"), weights_only=False, map_location="cpu")
    model = P.build_model(os.path.join(G, g["meta"]["checkpoint"].split("/")[-1]), device="cuda:0")
    inp = g["inputs"]; K = g["meta"]["k"]
    layers, pin, pout = model.egnn.packs()
    for tag, f, x, nbr, href, xref in (("src", inp["src_feat"], inp["src_pts"], g["nbr_src"], g["eval_f32"]["h_src"], g["eval_f32"]["x_src"]),
                                       ("tgt", inp["tgt_feat"], inp["tgt_pts"], g["nbr_tgt"], g["eval_f32"]["h_tgt"], g["eval_f32"]["x_tgt"])):
        gr = ops.csr_from_nbr(nbr.cuda())
        for impl in (3, 4, 5):
            h, xo = ops.egnn_forward(f.cuda(), x.cuda(), gr, layers, pin, pout, impl=impl)
            eh = float((h.cpu() - href).abs().max() / href.abs().max()); ex = float((xo.cpu() - xref).abs().max())
            print(f"{name:16s} {tag} impl {impl}: h rel-to-max {eh:.2e}   x abs {ex:.2e} m  (|x| max {float(xref.abs().max()):.1f})")
    # train variant, tempered weights: pose and H of the reduced modes against the fp32 path
    tm = P.build_model(os.path.join(G, g["meta"]["checkpoint"].split("/")[-1]), device="cuda:0", variant="train")
    with torch.no_grad():
        tm.egnn.embedding_out.weight.mul_(0.005); tm.egnn.embedding_out.bias.mul_(0.005)
    d = {k: v.cuda() for k, v in inp.items()}
    es, et = P.knn_graph_batch(d["src_pts"], K), P.knn_graph_batch(d["tgt_pts"], K)
    outs = {}
    for impl in (3, 4, 5):
        tm.egnn.impl = impl
        with torch.no_grad():
            o = tm(d["src_feat"], d["src_pts"], es, None, d["tgt_feat"], d["tgt_pts"], et, None, d["corr"], d["labels"], d["gt_pose"])
        outs[impl] = (o[0].cpu(), o[1].cpu(), tm.last_aux["H"].cpu())
    for impl in (4, 5):
        R0, t0, H0 = outs[3]; R1, t1, H1 = outs[impl]
        print(f"{name:16s} train variant impl {impl}: rot {max(ang(R0[b], R1[b]) for b in range(R0.shape[0])):.3e} deg   "
              f"t {float((t0 - t1).abs().max()):.2e} m   H rel {float((H0 - H1).abs().max() / H0.abs().max()):.2e}")
