"""Developer script (run under gpurun): stage-by-stage parity report against the golden fixtures
and the oracle, plus a per-kernel timing table.  Not part of the product or the test suite."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P  # noqa: E402
from se3_equi_graph_registration_b200 import ops  # noqa: E402
from oracle import egnn_oracle as O, knn_oracle  # noqa: E402

dev = "cuda:0"
G = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def main():
    print(torch.cuda.get_device_name(0))
    for name in ["small_b2_n256", "dup_b2_n512", "full_b1_n2048", "kitti_b1_n1024", "noenc_b1_n512"]:
        g = torch.load(os.path.join(G, name + ".pt"), weights_only=False, map_location="cpu")
        ck = os.path.join(G, g["meta"]["checkpoint"].split("/")[-1])
        model = P.build_model(ck, device=dev)
        inp = g["inputs"]
        B, N = inp["src_feat"].shape[:2]
        K = g["meta"]["k"]
        xs_all = torch.cat([inp["src_pts"], inp["tgt_pts"]]).to(dev)
        nbr = ops.knn_build(xs_all, K)
        nbr_ref = torch.cat([g["nbr_src"], g["nbr_tgt"]])
        mism = int((nbr.cpu() != nbr_ref).sum())
        print(f"[{name}] knn mismatches: {mism} / {nbr_ref.numel()}")
        graph = ops.csr_from_nbr(nbr)
        torch.cuda.synchronize()
        # CSR check against a torch construction
        C = 2 * B
        row = (nbr_ref.view(C, -1).long() + (torch.arange(C) * N)[:, None]).view(-1)
        col = (torch.arange(N).repeat_interleave(K)[None, :] + (torch.arange(C) * N)[:, None]).view(-1)
        order = torch.sort(row, stable=True).indices
        ok_row = bool((graph.row.cpu().long() == row[order]).all())
        ok_col = bool((graph.col.cpu().long() == col[order]).all())
        ptr_ref = torch.zeros(C * N + 1, dtype=torch.long); ptr_ref[1:] = torch.bincount(row, minlength=C * N).cumsum(0)
        ok_ptr = bool((graph.ptr.cpu().long() == ptr_ref).all())
        print(f"[{name}] csr ok: row {ok_row} col {ok_col} ptr {ok_ptr} err {int(graph.err.item())}")
        # through the module API (int64 edges, like the reference scripts)
        edges_s = P.knn_graph_batch(inp["src_pts"].to(dev), K)
        edges_t = P.knn_graph_batch(inp["tgt_pts"].to(dev), K)
        E = edges_s.shape[-1]
        ea = torch.ones(B, E, 1, device=dev)
        for variant in ("eval", "train"):
            if f"{variant}_f32" not in g:
                continue
            model.variant = variant
            with torch.no_grad():
                out = model(inp["src_feat"].to(dev), inp["src_pts"].to(dev), edges_s, ea,
                            inp["tgt_feat"].to(dev), inp["tgt_pts"].to(dev), edges_t, ea,
                            inp["corr"].to(dev), inp["labels"].to(dev), inp["gt_pose"].to(dev))
            torch.cuda.synchronize()
            ref = g[f"{variant}_f32"]
            print(f"[{name}/{variant}] h_src {rel(out[4].cpu(), ref['h_src']):.2e} h_tgt {rel(out[6].cpu(), ref['h_tgt']):.2e} "
                  f"x_src {float((out[5].cpu() - ref['x_src']).abs().max()):.2e} x_tgt {float((out[7].cpu() - ref['x_tgt']).abs().max()):.2e} "
                  f"R {float((out[0].cpu() - ref['R']).abs().max()):.2e} t {float((out[1].cpu() - ref['t']).abs().max()):.2e} "
                  f"loss {float(out[3]):.6f} vs {float(ref['equi_loss'].mean()):.6f}"
                  + (f" slot2 {float(out[2]):.6f} vs {float(ref['slot2'].mean()):.6f}" if out[2] is not None else ""))
            if "eval_f64" in g and variant == "eval":
                r64 = g["eval_f64"]
                print(f"      vs f64 ref: ours h {rel(out[4].cpu(), r64['h_src']):.2e}  (ref-f32 h {rel(ref['h_src'], r64['h_src']):.2e})")
            # weights / H against the oracle's aux
            sd = torch.load(ck, map_location="cpu", weights_only=True)["cross_attention_state_dict"]
            es_c, et_c = edges_s.cpu(), edges_t.cpu()
            fn = O.forward_eval if variant == "eval" else O.forward_train
            _, aux = fn(sd, inp["src_feat"], inp["src_pts"], es_c, inp["tgt_feat"], inp["tgt_pts"], et_c,
                        inp["labels"], inp["gt_pose"], return_aux=True)
            w = model.last_aux["w"].cpu(); Hm = model.last_aux["H"].cpu()
            for b in range(B):
                if variant == "eval":
                    wref = aux["w"][b]
                else:
                    wref = torch.zeros(N); wref[aux["valid"][b]] = aux["w"][b]
                print(f"      pair {b}: w err {float((w[b] - wref).abs().max()):.2e} (max w {float(wref.max()):.2e}) "
                      f"H err {float((Hm[b] - aux['H'][b]).abs().max()):.2e} (|H| {float(aux['H'][b].abs().max()):.2e})")
        # per-layer check (engine path with layers) on src cloud 0
        if "layers_src0" in g.get("eval_f32", {}):
            layers, pin, pout = model.egnn.packs()
            gr = ops.csr_from_nbr(nbr[:1].contiguous())
            h, x, lay = ops.egnn_forward(inp["src_feat"][:1].to(dev), inp["src_pts"][:1].to(dev), gr, layers, pin, pout,
                                         return_layers=True)
            refl = g["eval_f32"]["layers_src0"]
            for i, (hl, xl) in enumerate(lay):
                print(f"      layer {i}: h {rel(hl[0].cpu(), refl[i + 1][0]):.2e} x {float((xl[0].cpu() - refl[i + 1][1]).abs().max()):.2e}")
    # ---- timing at the bench configuration
    B, N, K = 64, 2048, 16
    model = P.build_model(os.path.join(G, "checkpoint-3dmatch.pth"), device=dev)
    data = P.synthetic.make_batch(5, B, n=N, pin=True)
    for impl in (1, 3):
        eng = P.RegistrationEngine(model, batch=B, n=N, k=K, use_graph=False)
        eng.impl = impl
        eng.load(data["src_feat"], data["src_pts"], data["tgt_feat"], data["tgt_pts"], data["labels"], data["gt_pose"])
        for _ in range(3):
            eng.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            eng.run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"engine impl={impl} B={B} N={N}: {ms:.3f} ms/step -> {B / ms * 1e3:.0f} pairs/s (no graph)")
        eng.use_graph = True
        for _ in range(3):
            eng.run()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            eng.run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"engine impl={impl} graph: {ms:.3f} ms/step -> {B / ms * 1e3:.0f} pairs/s")
    # stage timings
    eng = P.RegistrationEngine(model, batch=B, n=N, k=K, use_graph=False)
    eng.load(data["src_feat"], data["src_pts"], data["tgt_feat"], data["tgt_pts"], data["labels"], data["gt_pose"])
    eng.run(); torch.cuda.synchronize()
    def timeit(fn, n=10):
        for _ in range(2): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    print("knn grid ms", timeit(lambda: ops.knn_build(eng.x, K)))
    print("knn brute ms", timeit(lambda: ops.knn_build(eng.x, K, brute_force=True)))
    print("csr      ms", timeit(lambda: ops.csr_from_nbr(eng.nbr)))
    layers, pin, pout = model.egnn.packs()
    gr = ops.csr_from_nbr(eng.nbr)
    for impl in (1, 3):
        print(f"egnn impl={impl} (embed+3 layers) ms", timeit(lambda: ops.egnn_forward(eng.feat, eng.x, gr, layers, pin, pout, impl=impl)))
    ho, xo = ops.egnn_forward(eng.feat, eng.x, gr, layers, pin, pout)
    print("head     ms", timeit(lambda: ops.head_eval(eng.feat[:B], eng.feat[B:], eng.x[:B], eng.x[B:], ho[:B], ho[B:], xo[:B], xo[B:],
                                                      eng.labels, eng.gt_pose, model._pack_head.get())))


if __name__ == "__main__":
    main()
