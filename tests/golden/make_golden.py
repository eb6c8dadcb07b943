"""Generates the committed golden vectors by executing the REFERENCE'S OWN classes
(ast-extracted from /root/reference by oracle/ref_loader.py) on seeded synthetic pairs.
Runs only in the build container; the resulting tests/golden/*.pt travel to the GPU box.

    python tests/golden/make_golden.py

The k-NN ids in the fixtures come from oracle/knn_oracle.c (torch_cluster is absent -> k-NN
parity is UNPINNED, see that file's header); everything downstream of the graph is the
reference's arithmetic (fp32, and an fp64 run of the same classes for tolerance calibration).
"""
import hashlib
import os
import shutil
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import knn_oracle, ref_loader  # noqa: E402
from oracle.egnn_oracle import edges_from_nbr  # noqa: E402
import se3_equi_graph_registration_b200.synthetic as synthetic  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
K = 16


def batch_graph(pts):
    """[B,N,3] -> nbr int32 [B,N,K], edges int64 [B,2,N*K]"""
    nbr = torch.from_numpy(knn_oracle.knn(pts.numpy(), K))
    edges = torch.stack([torch.stack(edges_from_nbr(nbr[b])) for b in range(nbr.shape[0])])
    return nbr, edges


def run_case(name, seed, batch, n, shape, dup_frac, variants, ckpt, with_layers=True, with_f64=True):
    data = synthetic.make_batch(seed, batch, n=n, shape=shape, dup_frac=dup_frac)
    nbr_s, edges_s = batch_graph(data["src_pts"])
    nbr_t, edges_t = batch_graph(data["tgt_pts"])
    E = edges_s.shape[-1]
    out = {"meta": {"name": name, "seed": seed, "batch": batch, "n": n, "shape": shape,
                    "dup_frac": dup_frac, "k": K, "checkpoint": ckpt},
           "inputs": dict(data), "nbr_src": nbr_s, "nbr_tgt": nbr_t}
    ea = torch.ones(batch, E, 1)
    for variant in variants:
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            if dtype == torch.float64 and not with_f64:
                continue
            prev = torch.get_default_dtype()
            torch.set_default_dtype(dtype)   # reference builds torch.eye / zeros in the default dtype
            try:
                ns, egnn, head = ref_loader.build_reference_model(variant, ckpt, dtype=dtype)
                cast = lambda v: v.to(dtype)
                with torch.no_grad():
                    res = []
                    # eval variant is B=1-only: call it per pair, exactly as evl:1122 does
                    step = 1 if variant == "eval" else batch
                    for b0 in range(0, batch, step):
                        sl = slice(b0, b0 + step)
                        r = ref_loader.run_quiet(
                            head, cast(data["src_feat"][sl]), cast(data["src_pts"][sl]), edges_s[sl], cast(ea[sl]),
                            cast(data["tgt_feat"][sl]), cast(data["tgt_pts"][sl]), edges_t[sl], cast(ea[sl]),
                            data["corr"][sl], cast(data["labels"][sl]), cast(data["gt_pose"][sl]))
                        res.append(r)
                    R = torch.cat([r[0] for r in res]); t = torch.cat([r[1] for r in res])
                    rec = {"R": R, "t": t,
                           "slot2": None if res[0][2] is None else torch.stack([r[2] for r in res]),
                           "equi_loss": torch.stack([r[3] for r in res]),
                           "h_src": torch.cat([r[4] for r in res]), "x_src": torch.cat([r[5] for r in res]),
                           "h_tgt": torch.cat([r[6] for r in res]), "x_tgt": torch.cat([r[7] for r in res])}
                    if with_layers and variant == "eval":
                        # per-layer states of pair 0 / source cloud through the reference's own gcl modules
                        h = egnn.embedding_in(cast(data["src_feat"][0])); x = cast(data["src_pts"][0])
                        edges = [edges_s[0, 0], edges_s[0, 1]]
                        layers = [(h.clone(), x.clone())]
                        for i in range(egnn.n_layers):
                            gcl = egnn._modules["gcl_%d" % i]
                            if i == 0:
                                radial, cd = gcl.coord2radial(edges, x)
                                m0 = gcl.edge_model(h[edges[0]], h[edges[1]], radial, cast(ea[0]), x, edges)
                                rec["layer0_messages"] = m0[: 64 * K].clone().to(torch.float32) if tag == "f64" else m0[: 64 * K].clone()
                            h, x, _ = gcl(h, edges, x, edge_attr=cast(ea[0]))
                            layers.append((h.clone(), x.clone()))
                        rec["layers_src0"] = layers
                if tag == "f64":   # store the fp64 answers rounded to fp32 (keeps the fixtures small)
                    for key in ("h_src", "x_src", "h_tgt", "x_tgt"):
                        rec[key] = rec[key].to(torch.float32)
                    if "layers_src0" in rec:
                        rec["layers_src0"] = [(a.to(torch.float32), b.to(torch.float32)) for a, b in rec["layers_src0"]]
                out[f"{variant}_{tag}"] = rec
            finally:
                torch.set_default_dtype(prev)
    path = os.path.join(HERE, name + ".pt")
    torch.save(out, path)
    print(name, os.path.getsize(path) // 1024, "KiB")


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    knn_oracle.build()
    # checkpoint fixtures: binary copies of the reference's shipped weights (data, not source)
    for ck in ("checkpoint-3dmatch.pth", "checkpoint-3dmatch-no-encoder.pth"):
        src = os.path.join(ref_loader.REF_ROOT, "checkpoints", ck)
        shutil.copyfile(src, os.path.join(HERE, ck))
        print(ck, hashlib.md5(open(src, "rb").read()).hexdigest())
    ck = "checkpoints/checkpoint-3dmatch.pth"
    run_case("small_b2_n256", seed=1, batch=2, n=256, shape="3dmatch", dup_frac=0.0,
             variants=("eval", "train"), ckpt=ck)
    run_case("dup_b2_n512", seed=2, batch=2, n=512, shape="3dmatch", dup_frac=0.3,
             variants=("eval", "train"), ckpt=ck)
    run_case("full_b1_n2048", seed=3, batch=1, n=2048, shape="3dmatch", dup_frac=0.0,
             variants=("eval",), ckpt=ck, with_f64=True)
    run_case("kitti_b1_n1024", seed=4, batch=1, n=1024, shape="kitti", dup_frac=0.05,
             variants=("eval",), ckpt=ck, with_layers=False)
    run_case("noenc_b1_n512", seed=5, batch=1, n=512, shape="3dmatch", dup_frac=0.1,
             variants=("eval", "train"), ckpt="checkpoints/checkpoint-3dmatch-no-encoder.pth", with_layers=False)
    # metrics known-answer vectors from tools/evaluation_metrics.py (importable: numpy/scipy only)
    sys.path.insert(0, os.path.join(ref_loader.REF_ROOT, "tools"))
    import evaluation_metrics as em
    rng = np.random.default_rng(7)
    cases = []
    for i in range(8):
        gt = np.eye(4); pr = np.eye(4)
        gt[:3, :3] = synthetic.random_rotation(rng); gt[:3, 3] = rng.random(3)
        dR = synthetic.random_rotation(rng) if i % 2 else np.eye(3)
        pr[:3, :3] = gt[:3, :3] @ (dR if i % 4 == 1 else np.eye(3)); pr[:3, 3] = gt[:3, 3] + rng.standard_normal(3) * 0.05 * (i % 3)
        src = rng.random((200, 3)) * 3
        tgt = src @ gt[:3, :3].T + gt[:3, 3] + rng.standard_normal((200, 3)) * 0.05
        re, te = em.calculate_pose_error(gt, pr)
        rec, prec = em.registration_recall(gt, pr, src, tgt)
        cases.append({"gt": gt, "pred": pr, "src": src, "tgt": tgt, "re": re, "te": te, "recall": rec, "precision": prec})
    torch.save(cases, os.path.join(HERE, "metrics_kat.pt"))
    print("metrics_kat", len(cases))


if __name__ == "__main__":
    main()
