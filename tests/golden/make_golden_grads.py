"""Golden GRADIENTS of the training step, produced by the REFERENCE'S OWN classes and autograd
(ast-extracted from /root/reference by oracle/ref_loader.py; src/3dmatch_train_egnn_with_batch.py:634-796 forward,
:896-962 pose_loss, :1094-1125 loss assembly + loss.backward()).  Build container only.

    python tests/golden/make_golden_grads.py        -> tests/golden/grads_b2_n256.pt, grads_dup_b2_n512.pt

Two scenarios on the inputs of the small_b2_n256 fixture and of the duplicate-heavy dup_b2_n512 fixture (30 % repeated
points: zero-length edges, identity frames):
  shipped   the shipped checkpoint; loss = slot 2 (corr_loss + sim_loss) + slot 3 (egnn_equi_loss).  The pose terms
            are left out because with these weights the train-variant H is ~1e-6 I and the SVD gradient
            is not defined.
  tempered  the same weights with embedding_out scaled by 0.005 (similarity logits O(1), well-conditioned H);
            loss = the training loop's total (3dm:1118): slot 2 + mean rotation loss + mean translation loss.
Stored: every parameter's gradient (fp32 run; an fp64 run rounded to fp32 where the reference's backward
supports it), the loss values, R and t.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle.egnn_oracle import edges_from_nbr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TEMPER = 0.005


def run(scenario, g, dtype):
    prev = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        ns, egnn, head = ref_loader.build_reference_model("train", g["meta"]["checkpoint"], dtype=dtype)
        if scenario == "tempered":
            with torch.no_grad():
                egnn.embedding_out.weight.mul_(TEMPER)
                egnn.embedding_out.bias.mul_(TEMPER)
        head.train()
        inp = g["inputs"]
        B = inp["labels"].shape[0]
        es = torch.stack([torch.stack(edges_from_nbr(g["nbr_src"][b])) for b in range(B)])
        et = torch.stack([torch.stack(edges_from_nbr(g["nbr_tgt"][b])) for b in range(B)])
        ea = torch.ones(B, es.shape[-1], 1, dtype=dtype)
        c = lambda v: v.to(dtype)
        out = ref_loader.run_quiet(head, c(inp["src_feat"]), c(inp["src_pts"]), es, ea, c(inp["tgt_feat"]), c(inp["tgt_pts"]),
                                   et, ea, inp["corr"], c(inp["labels"]), c(inp["gt_pose"]))
        R, t, slot2, slot3 = out[0], out[1], out[2], out[3]
        if scenario == "shipped":
            loss = slot2.mean() + slot3
        else:
            rot, trans = ns["pose_loss"](R, t, c(inp["gt_pose"]), delta=1.5)            # 3dm:1097
            loss = slot2.mean() + rot.mean() + trans.mean()                             # 3dm:1107-1118
        loss.backward()
        grads = {k: (p.grad.detach().to(torch.float32) if p.grad is not None else None) for k, p in head.named_parameters()}
        return {"loss": float(loss), "slot2": float(slot2), "slot3": float(slot3), "R": R.detach().to(torch.float32),
                "t": t.detach().to(torch.float32), "grads": grads}
    finally:
        torch.set_default_dtype(prev)


def main():
    assert ref_loader.reference_available(), "needs /root/reference"
    for case, fname in (("small_b2_n256", "grads_b2_n256.pt"), ("dup_b2_n512", "grads_dup_b2_n512.pt")):
        make(case, fname)


def make(case, fname):
    g = torch.load(os.path.join(HERE, case + ".pt"), weights_only=False, map_location="cpu")
    out = {"meta": {"case": case, "temper": TEMPER}}
    for scenario in ("shipped", "tempered"):
        for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
            try:
                out[f"{scenario}_{tag}"] = run(scenario, g, dtype)
            except RuntimeError as e:      # the reference's backward mixes dtypes when run in fp64
                print(scenario, tag, "skipped:", str(e).splitlines()[0])
                continue
            r = out[f"{scenario}_{tag}"]
            n_none = sum(v is None for v in r["grads"].values())
            gmax = max(float(v.abs().max()) for v in r["grads"].values() if v is not None)
            print(scenario, tag, "loss", r["loss"], "params without grad", n_none, "max |grad|", gmax)
    path = os.path.join(HERE, fname)
    torch.save(out, path)
    print(os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
