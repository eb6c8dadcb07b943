"""Developer script (gpurun): error of the reduced-precision edge mode (impl 4) against the golden fixtures."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
G = os.path.join(ROOT, "tests", "golden")
for name in ["small_b2_n256", "dup_b2_n512", "full_b1_n2048", "kitti_b1_n1024", "noenc_b1_n512"]:
    g = torch.load(os.path.join(G, name + ".pt"), weights_only=False, map_location="cpu")
    model = P.build_model(os.path.join(G, g["meta"]["checkpoint"].split("/")[-1]), device="cuda:0")
    inp = g["inputs"]; K = g["meta"]["k"]
    layers, pin, pout = model.egnn.packs()
    for tag, f, x, nbr, href, xref in (("src", inp["src_feat"], inp["src_pts"], g["nbr_src"], g["eval_f32"]["h_src"], g["eval_f32"]["x_src"]),
                                       ("tgt", inp["tgt_feat"], inp["tgt_pts"], g["nbr_tgt"], g["eval_f32"]["h_tgt"], g["eval_f32"]["x_tgt"])):
        gr = ops.csr_from_nbr(nbr.cuda())
        for impl in (3, 4):
            h, xo = ops.egnn_forward(f.cuda(), x.cuda(), gr, layers, pin, pout, impl=impl)
            eh = float((h.cpu() - href).abs().max() / href.abs().max()); ex = float((xo.cpu() - xref).abs().max())
            print(f"{name:16s} {tag} impl {impl}: h rel-to-max {eh:.2e}   x abs {ex:.2e} m  (|x| max {float(xref.abs().max()):.1f})")
