"""GPU: the backward kernels (egspr_egcl_backward, egspr_linear32_backward, egspr_head_train_backward) through the
C ABI and the nn.Module API, against (i) the reference's own gradients (tests/golden/grads_b2_n256.pt, produced by
tests/golden/make_golden_grads.py with the reference's classes + autograd) and (ii) torch autograd of the oracle on
the same inputs.  Tolerance: |dg| <= 1e-3 * max|g| per tensor (the reference's fp32 run is 4e-5 from its fp64 run;
our forward is 3xTF32 and the sums run in a different order)."""
import os

import numpy as np
import pytest
import torch

import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops, packing
from oracle import egnn_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
G_TOL = 1e-3


def edges_of(nbr):
    return torch.stack([torch.stack(O.edges_from_nbr(n)) for n in nbr])


def rel(a, b):
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max() / (b.detach().abs().max().double() + 1e-30))


def _model(golden_dir, temper=None):
    m = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device=DEV, variant="train")
    if temper is not None:
        with torch.no_grad():
            m.egnn.embedding_out.weight.mul_(temper)
            m.egnn.embedding_out.bias.mul_(temper)
    return m


@pytest.mark.parametrize("per_edge_attr", [False, True])
def test_egcl_layer_backward_matches_autograd(golden_dir, per_edge_attr):
    """One E_GCL layer through the module API (E_GCL.forward 3dm:280-289) on an arbitrary user graph with duplicate
    points: gradients w.r.t. h, coord and every parameter vs autograd of the oracle (fp64).  The edge part is the
    tcgen05 kernel (egnn_edge_bwd_tc.cu); its formulas are the CPU-checked ones of egnn_backward_math.cuh."""
    model = _model(golden_dir)
    gcl = model.egnn.gcl_1
    g = torch.Generator().manual_seed(3)
    n, k = 700, 16
    x = torch.rand(n, 3, generator=g) * 2
    x[n // 2: n // 2 + 100] = x[:100]
    d2 = ((x[:, None] - x[None]) ** 2).sum(-1)
    nbr = d2.argsort(dim=1, stable=True)[:, :k]
    row, col = O.edges_from_nbr(nbr)
    extra = torch.randint(0, n, (2, 1000), generator=g)
    row, col = torch.cat([row, extra[0]]), torch.cat([col, extra[1]])
    E = row.numel()
    h = torch.randn(n, 32, generator=g) * 0.5
    ea = torch.rand(E, 1, generator=g) + 0.5 if per_edge_attr else torch.ones(E, 1)
    dh = torch.randn(n, 32, generator=g); dx = torch.randn(n, 3, generator=g)
    # oracle
    sd = {"gcl_1." + k_: v.detach().cpu().double().requires_grad_(True) for k_, v in gcl.state_dict().items()}
    h64, x64 = h.double().requires_grad_(True), x.double().requires_grad_(True)
    h2r, x2r, _ = O.egcl_forward(sd, "gcl_1.", h64, x64, row, col, ea.double())
    ((h2r * dh.double()).sum() + (x2r * dx.double()).sum()).backward()
    # CUDA
    hg, xg = h.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
    for p in gcl.parameters():
        p.grad = None
    h2, x2, _ = gcl(hg, [row.to(DEV), col.to(DEV)], xg, edge_attr=ea.to(DEV))
    assert rel(h2, h2r) < 1e-4 and float((x2.detach().cpu() - x2r.detach()).abs().max()) < 1e-4
    ((h2 * dh.to(DEV)).sum() + (x2 * dx.to(DEV)).sum()).backward()
    assert rel(hg.grad, h64.grad) < G_TOL, rel(hg.grad, h64.grad)
    assert rel(xg.grad, x64.grad) < G_TOL, rel(xg.grad, x64.grad)
    for name, p in gcl.named_parameters():
        assert p.grad is not None, name
        assert rel(p.grad, sd["gcl_1." + name].grad) < G_TOL, (name, rel(p.grad, sd["gcl_1." + name].grad))


def test_head_train_backward_matches_autograd():
    """egspr_head_train_backward vs autograd of the oracle's weights + Kabsch (3dm:696-758), incl. a pair whose
    optimal alignment is a reflection (det < 0 branch, where the reference's own backward raises) and an empty pair."""
    g = torch.Generator().manual_seed(11)
    B, n = 4, 300
    hs = torch.randn(B, n, 32, generator=g) * 0.25
    ht = hs + 0.1 * torch.randn(B, n, 32, generator=g)
    xs = torch.randn(B, n, 3, generator=g)
    xt = torch.empty(B, n, 3)
    for b in range(B):
        Q = torch.linalg.qr(torch.randn(3, 3, generator=g))[0]
        if (torch.det(Q) < 0) != (b == 2):
            Q[:, 0] *= -1                                           # pair 2: a reflection
        xt[b] = xs[b] @ Q.T + 0.05 * torch.randn(n, 3, generator=g) + torch.tensor([0.3, -0.2, 0.5])
    labels = (torch.rand(B, n, generator=g) < 0.7).float()
    labels[3] = 0                                                   # pair 3: no inliers -> R = I, t = 0, no gradient
    dR = torch.randn(B, 3, 3, generator=g); dt = torch.randn(B, 3, generator=g); dsim = torch.randn(B, n, generator=g) * 0.1
    leaves = [v.double().requires_grad_(True) for v in (hs, ht, xs, xt)]
    loss = 0
    for b in range(B):
        w, valid = O.train_weights_one(leaves[0][b], leaves[1][b], labels[b])
        R, t, _ = O.kabsch(leaves[2][b][valid], leaves[3][b][valid], w)
        loss = loss + (R * dR[b].double()).sum() + (t * dt[b].double()).sum()
    loss = loss + ((leaves[0] * leaves[1]).sum(-1) * dsim.double()).sum()
    loss.backward()
    outs = ops.head_train_backward(hs.to(DEV), ht.to(DEV), xs.to(DEV), xt.to(DEV), labels.to(DEV), dR.to(DEV), dt.to(DEV), dsim.to(DEV))
    for o, l, name in zip(outs, leaves, ("dh_src", "dh_tgt", "dx_src", "dx_tgt")):
        for b in range(B):
            sc = float(l.grad[b].abs().max()) + 1e-12
            err = float((o[b].cpu().double() - l.grad[b]).abs().max())
            assert err <= 2e-4 * sc, (name, b, err, sc)


@pytest.mark.parametrize("case,gfile", [("small_b2_n256", "grads_b2_n256.pt"), ("dup_b2_n512", "grads_dup_b2_n512.pt")])
@pytest.mark.parametrize("scenario", ["shipped", "tempered"])
def test_training_step_gradients_match_reference_golden(golden_dir, scenario, case, gfile):
    """The whole training step through the drop-in module (forward train variant -> the loop's loss -> backward)
    against the gradients the reference's own classes + autograd produced for the same inputs; the second fixture is
    duplicate-heavy (30 % repeated points: zero-length edges, identity frames, twin nodes)."""
    g = torch.load(os.path.join(golden_dir, case + ".pt"), weights_only=False, map_location="cpu")
    gg = torch.load(os.path.join(golden_dir, gfile), weights_only=False, map_location="cpu")
    ref = gg[scenario + "_f32"]
    model = _model(golden_dir, gg["meta"]["temper"] if scenario == "tempered" else None)
    model.train()
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    es, et = edges_of(g["nbr_src"]).to(DEV), edges_of(g["nbr_tgt"]).to(DEV)
    ea = torch.ones(es.shape[0], es.shape[-1], 1, device=DEV)
    out = model(inp["src_feat"], inp["src_pts"], es, ea, inp["tgt_feat"], inp["tgt_pts"], et, ea, inp["corr"], inp["labels"], inp["gt_pose"])
    if scenario == "shipped":
        loss = out[2] + out[3]
    else:
        loss = P.train.training_loss(out, inp["gt_pose"])
        assert float((out[0].cpu() - ref["R"]).abs().max()) < 1e-4 and float((out[1].cpu() - ref["t"]).abs().max()) < 1e-4
    assert abs(float(loss.detach()) - ref["loss"]) <= 1e-4 * abs(ref["loss"])
    loss.backward()
    params = dict(model.named_parameters())
    worst, n = 0.0, 0
    for k, gref in ref["grads"].items():
        if gref is None:
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
            continue
        assert params[k].grad is not None, k
        e = rel(params[k].grad, gref)
        worst = max(worst, e)
        assert e < G_TOL, (k, e)
        n += 1
    assert n == 85
    print(f"{case} {scenario}: worst relative-to-max gradient error {worst:.2e}")


def test_edge_attr_none_equals_ones(golden_dir):
    """edge_attr=None (constant 1 folded into the kernels) gives the same loss and gradients as the reference's explicit
    ones tensor (get_edges_batch, 3dm:387), which goes through the per-edge gather."""
    g = torch.load(os.path.join(golden_dir, "small_b2_n256.pt"), weights_only=False, map_location="cpu")
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    es, et = edges_of(g["nbr_src"]).to(DEV), edges_of(g["nbr_tgt"]).to(DEV)
    grads = []
    for ea in (None, torch.ones(es.shape[0], es.shape[-1], 1, device=DEV)):
        model = _model(golden_dir, 0.005)
        model.train()
        out = model(inp["src_feat"], inp["src_pts"], es, ea, inp["tgt_feat"], inp["tgt_pts"], et, ea, inp["corr"], inp["labels"], inp["gt_pose"])
        P.train.training_loss(out, inp["gt_pose"]).backward()
        grads.append({k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert grads[0].keys() == grads[1].keys() and len(grads[0]) == 85
    for k in grads[0]:
        assert rel(grads[0][k], grads[1][k]) < 1e-5, k


def test_train_step_runs_and_reduces_the_loss(golden_dir):
    """train.train_step (3dm:1092-1126) with Adam on a fixed batch: finite gradients, loss goes down, packs follow
    the parameter updates."""
    model = _model(golden_dir, 0.005)
    data = P.synthetic.make_batch(21, 4, n=512)
    nbr_s = ops.knn_build(data["src_pts"].to(DEV), 16); nbr_t = ops.knn_build(data["tgt_pts"].to(DEV), 16)
    es, et = ops.nbr_to_edges(nbr_s), ops.nbr_to_edges(nbr_t)
    ea = torch.ones(4, es.shape[-1], 1, device=DEV)
    batch = (data["src_feat"].to(DEV), data["src_pts"].to(DEV), es, ea, data["tgt_feat"].to(DEV), data["tgt_pts"].to(DEV), et, ea,
             data["corr"].to(DEV), data["labels"].to(DEV), data["gt_pose"].to(DEV))
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    losses = [float(P.train.train_step(model, opt, batch)) for _ in range(8)]
    assert all(np.isfinite(losses)), losses
    assert losses[-1] < losses[0], losses


def test_graphed_train_step_equals_eager(golden_dir):
    """train.GraphedTrainStep (the fixed ~35-kernel launch sequence with the fused loss kernels, captured as one CUDA
    graph) follows the eager train_step (nn.Module API + torch.autograd): same losses and the same parameters after
    several optimizer updates.  Building the step must not train the model (warm-up is undone), and an eager forward
    after replays must see the updated weights (the replay changes the parameters behind the version counters)."""
    keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
    batches = [tuple(P.synthetic.make_batch(40 + i, 2, n=512)[k].to(DEV) for k in keys) for i in range(3)]
    ones = torch.ones(2, 512 * 16, 1, device=DEV)
    losses, finals = {}, {}
    for mode in ("eager", "graph"):
        model = _model(golden_dir, 0.005)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
        out = []
        if mode == "graph":
            w0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
            step = P.train.GraphedTrainStep(model, opt, batches[0], k=16, warmup=2)
            assert all(torch.equal(v, w0[k]) for k, v in model.state_dict().items())           # construction did not train
            for i in range(6):
                out.append(float(step(batches[i % 3])))
            assert sum(p.grad is not None for p in model.parameters()) == 85                    # dead parameters stay None
            # eager eval forward after the replays == a fresh model holding the same state_dict
            sf, sp, tf, tp, corr, labels, gt = batches[0]
            es, et = P.knn_graph_batch(sp, 16), P.knn_graph_batch(tp, 16)
            fresh = _model(golden_dir)
            fresh.load_state_dict(model.state_dict())
            with torch.no_grad():
                a_ = model(sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt)
                b_ = fresh(sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt)
            assert torch.equal(a_[4], b_[4]) and torch.equal(a_[2], b_[2])
        else:
            for i in range(6):
                sf, sp, tf, tp, corr, labels, gt = batches[i % 3]
                es, et = P.knn_graph_batch(sp, 16), P.knn_graph_batch(tp, 16)
                out.append(float(P.train.train_step(model, opt, (sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt))))
        losses[mode] = out
        finals[mode] = {k: v.detach().clone() for k, v in model.named_parameters()}
    assert np.all(np.isfinite(losses["graph"]))
    assert np.allclose(losses["eager"], losses["graph"], rtol=2e-3), losses
    # Adam divides by sqrt(v): an element whose gradient is at rounding level (the two paths sum the same terms in a
    # different order, and weight-gradient atomics land in launch order) may take a step of a different size, so a few
    # elements differ by a fraction of one step (lr = 1e-4; six steps move a weight by <= 6e-4).  Bound the worst element
    # by half a step and the typical element much tighter.
    for k in finals["eager"]:
        d = (finals["eager"][k] - finals["graph"][k]).abs()
        assert float(d.max()) <= 5e-5, (k, float(d.max()))
        assert float(d.mean()) <= 2e-6, (k, float(d.mean()))


def test_fused_train_losses_match_torch_formulas(golden_dir):
    """egspr_train_loss_forward / _finalize / egspr_head_train_loss_backward against the reference's torch
    formulas (3dm:681-694, 760-781) evaluated with autograd in fp64: loss values, the top-k SET, d loss / d (h_src_out,
    h_tgt_out, sim) and the six mlp gradients; n < top_k and heavy ties in sim included."""
    model = _model(golden_dir, 0.005)
    g = torch.Generator().manual_seed(9)
    for B, n, k in ((3, 700, 128), (2, 90, 128)):
        hs = torch.randn(B, n, 32, generator=g) * 0.3
        ht = hs + 0.2 * torch.randn(B, n, 32, generator=g)
        ht[:, n // 2:n // 2 + 20] = ht[:, :20]; hs[:, n // 2:n // 2 + 20] = hs[:, :20]        # exact ties in sim
        fs = torch.nn.functional.normalize(torch.randn(B, n, 32, generator=g), dim=-1)
        ft = torch.nn.functional.normalize(fs + 0.3 * torch.randn(B, n, 32, generator=g), dim=-1)
        labels = (torch.rand(B, n, generator=g) < 0.6).float()
        # reference formulas, fp64
        sd = {k_: v.detach().cpu().double().requires_grad_(True) for k_, v in model.mlp.state_dict().items()}
        H, T = hs.double().requires_grad_(True), ht.double().requires_grad_(True)
        sim = (H * T).sum(-1)
        sim_leaf = sim.detach().requires_grad_(True)
        kk = min(k, n)
        top = torch.topk(sim_leaf, kk, dim=-1).indices
        z = torch.cat([torch.gather(H, 1, top[..., None].expand(-1, -1, 32)), torch.gather(T, 1, top[..., None].expand(-1, -1, 32))], -1)
        a1 = torch.relu(z @ sd["0.weight"].T + sd["0.bias"]); a2 = torch.relu(a1 @ sd["2.weight"].T + sd["2.bias"])
        sc = (a2 @ sd["4.weight"].T + sd["4.bias"])[..., 0]
        corr = torch.nn.functional.binary_cross_entropy_with_logits(sc, torch.gather(labels.double(), 1, top))
        raw = (fs.double() * ft.double()).sum(-1)
        zs = (sim_leaf - sim_leaf.mean()) / (sim_leaf.std() + 1e-6); zr = (raw - raw.mean()) / (raw.std() + 1e-6)
        siml = torch.nn.functional.mse_loss(zs, zr)
        (corr + siml).backward()
        # kernels
        d = lambda v: v.float().to(DEV).contiguous()
        pack = model._pack_head.get()
        top_idx, scores, raw_k, stats, bce = ops.train_loss_forward(d(hs), d(ht), d(fs), d(ft), d(sim.detach()), d(labels), pack, k)
        loss, dsim, _, _ = ops.train_loss_finalize(d(sim.detach()), raw_k, stats, bce, k, scale=1.0)
        assert abs(float(loss[0]) - float(corr)) <= 2e-5 * max(1.0, abs(float(corr))) and abs(float(loss[1]) - float(siml)) <= 2e-5 * max(1.0, float(siml))
        for b in range(B):
            got = sorted(int(i) for i in top_idx[b].cpu() if i >= 0)
            assert got == sorted(int(i) for i in top[b]), b
        assert rel(dsim, sim_leaf.grad) < 1e-4
        gp = torch.zeros(packing.HEAD_PACK, device=DEV)
        zx = torch.zeros(B, n, 3, device=DEV)
        dhs, dht, _, _ = ops.head_train_loss_backward(d(hs), d(ht), zx, zx, d(labels), torch.zeros(B, 3, 3, device=DEV), torch.zeros(B, 3, device=DEV),
                                                      None, top_idx, pack, loss, gp, k)
        assert rel(dhs, H.grad) < 1e-4 and rel(dht, T.grad) < 1e-4, (rel(dhs, H.grad), rel(dht, T.grad))
        for got_g, (name, prm) in zip(packing.unpack_head_grad(gp, model.mlp), model.mlp.named_parameters()):
            assert rel(got_g, sd[name].grad) < 1e-4, (name, rel(got_g, sd[name].grad))


def test_backward_on_irregular_graph_hub_rows_isolated_nodes(golden_dir):
    """EGNN.forward through all 3 layers + both embeddings on user-supplied edge lists with a 700-edge hub row (spans
    several 128-edge tiles; its lists take many 8-edge steps in the gather kernel), nodes without any edge (empty
    row and col lists), multi-edges, several clouds, per-edge edge_attr: every gradient vs autograd of the oracle."""
    model = _model(golden_dir)
    g = torch.Generator().manual_seed(5)
    C, N = 2, 300
    rows, cols = [], []
    for c in range(C):
        r = torch.cat([torch.full((700,), 7 + c, dtype=torch.int64), torch.randint(0, N // 2, (900,), generator=g),
                       torch.full((130,), N // 2 - 1, dtype=torch.int64)])
        cc = torch.randint(0, N - 20, (r.numel(),), generator=g)            # nodes N-20.. have no edge at all
        perm = torch.randperm(r.numel(), generator=g)
        rows.append(r[perm]); cols.append(cc[perm])
    E = rows[0].numel()
    feat = torch.randn(C, N, 32, generator=g) * 0.5
    x = torch.rand(C, N, 3, generator=g) * 2
    ea = torch.rand(C, E, 1, generator=g) + 0.5
    dh = torch.randn(C, N, 32, generator=g); dx = torch.randn(C, N, 3, generator=g)
    sd = {k: v.detach().cpu().double().requires_grad_(True) for k, v in model.egnn.state_dict().items()}
    f64, x64 = feat.double().requires_grad_(True), x.double().requires_grad_(True)
    loss = 0
    for c in range(C):
        ho, xo = O.egnn_forward(sd, f64[c], x64[c], rows[c], cols[c], ea[c].double())
        loss = loss + (ho * dh[c].double()).sum() + (xo * dx[c].double()).sum()
    loss.backward()
    for p in model.parameters():
        p.grad = None
    edges = torch.stack([torch.stack([rows[c], cols[c]]) for c in range(C)]).to(DEV)
    graph = ops.csr_from_edges(edges, N)
    fg, xg = feat.to(DEV).requires_grad_(True), x.to(DEV).requires_grad_(True)
    ho, xo = model.egnn.forward_batch(fg, xg, graph, edge_attr=ea.to(DEV), edge_attr_const=0.0)
    ((ho * dh.to(DEV)).sum() + (xo * dx.to(DEV)).sum()).backward()
    assert rel(fg.grad, f64.grad) < G_TOL and rel(xg.grad, x64.grad) < G_TOL, (rel(fg.grad, f64.grad), rel(xg.grad, x64.grad))
    assert float(xg.grad[:, N - 20:].abs().max()) == float(dx[:, N - 20:].abs().max())      # isolated nodes: identity
    for name, p in model.egnn.named_parameters():
        assert p.grad is not None, name
        assert rel(p.grad, sd[name].grad) < G_TOL, (name, rel(p.grad, sd[name].grad))


def test_pose_loss_kernel_matches_reference_function():
    """egspr_pose_loss (values + local gradients) vs the oracle's pose_loss (3dm:896-962) and its autograd."""
    g = torch.Generator().manual_seed(4)
    B = 37
    R = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    R = R + 0.05 * torch.randn(B, 3, 3, generator=g)               # not exactly orthogonal: the loss must not assume it
    t = torch.randn(B, 3, generator=g)
    gt = torch.eye(4).repeat(B, 1, 1)
    gt[:, :3, :3] = torch.linalg.qr(torch.randn(B, 3, 3, generator=g))[0]
    gt[:, :3, 3] = torch.randn(B, 3, generator=g)
    gt[0, :3, :3] = R[0] * 1.5                                      # trace term > 1: clamped, zero gradient
    Rr, tr = R.double().requires_grad_(True), t.double().requires_grad_(True)
    rl_ref, tl_ref = O.pose_loss(Rr, tr, gt.double())
    wr, wt = torch.randn(B, generator=g).double(), torch.randn(B, generator=g).double()
    ((rl_ref * wr).sum() + (tl_ref * wt).sum()).backward()
    Rg, tg = R.to(DEV).requires_grad_(True), t.to(DEV).requires_grad_(True)
    rl, tl = P.pose_loss(Rg, tg, gt.to(DEV))
    ((rl * wr.float().to(DEV)).sum() + (tl * wt.float().to(DEV)).sum()).backward()
    assert torch.allclose(rl.detach().cpu().double(), rl_ref.detach(), atol=2e-6) and torch.allclose(tl.detach().cpu().double(), tl_ref.detach(), atol=2e-6)
    assert rel(Rg.grad, Rr.grad) < 1e-5 and rel(tg.grad, tr.grad) < 1e-5
    assert float(Rg.grad[0].abs().max()) == 0.0
    # the singular point acos'(+-1): a prediction equal to the ground truth in fp32.  torch's autograd gives -+inf there
    # (and NaN for every EGNN gradient behind it); the kernel returns a zero subgradient -- a documented deviation.
    gt1 = torch.eye(4).repeat(2, 1, 1)
    t1 = torch.tensor([[0.3, -0.2, 0.9], [-1.0, 0.5, 0.25]])
    gt1[:, :3, 3] = torch.stack([t1[0], -t1[1]])                     # cos = +1 and -1 exactly
    R1, tt = torch.eye(3).repeat(2, 1, 1).to(DEV).requires_grad_(True), t1.to(DEV).requires_grad_(True)
    rl1, tl1 = P.pose_loss(R1, tt, gt1.to(DEV))
    (rl1.sum() + tl1.sum()).backward()
    assert float(rl1.abs().max()) == 0.0 and torch.isfinite(tl1).all()
    assert torch.isfinite(R1.grad).all() and torch.isfinite(tt.grad).all() and float(R1.grad.abs().max()) == 0.0


def test_full_size_gradients_against_fp64_autograd(golden_dir):
    """BASELINE config 4's real cloud size: ONE pair of 2048 points (k = 16, 65,536 edges) through the module API, every
    one of the 85 gradients against fp64 autograd of the oracle on the same inputs (tempered weights: well-conditioned
    Kabsch), same 1e-3 * max|g| bar as the small fixtures."""
    model = _model(golden_dir, 0.005)
    model.train()
    data = P.synthetic.make_batch(77, 1, n=2048)
    nbr_s = ops.knn_build(data["src_pts"].to(DEV), 16).cpu(); nbr_t = ops.knn_build(data["tgt_pts"].to(DEV), 16).cpu()
    es, et = edges_of(nbr_s.long()), edges_of(nbr_t.long())
    sd = {k: (v.detach().cpu().double().requires_grad_(True) if v.is_floating_point() else v.cpu()) for k, v in model.state_dict().items()}
    d64 = {k: v.double() for k, v in data.items()}
    out = O.forward_train(sd, d64["src_feat"], d64["src_pts"], es, d64["tgt_feat"], d64["tgt_pts"], et, data["labels"], d64["gt_pose"])
    rot, trans = O.pose_loss(out[0], out[1], d64["gt_pose"])
    ref_loss = out[2] + rot.mean() + trans.mean()
    ref_loss.backward()
    dv = {k: v.to(DEV) for k, v in data.items()}
    o = model(dv["src_feat"], dv["src_pts"], es.to(DEV), None, dv["tgt_feat"], dv["tgt_pts"], et.to(DEV), None, dv["corr"], dv["labels"], dv["gt_pose"])
    loss = P.train.training_loss(o, dv["gt_pose"])
    assert abs(float(loss.detach()) - float(ref_loss)) <= 1e-4 * abs(float(ref_loss))
    loss.backward()
    worst, n = 0.0, 0
    for k, p in model.named_parameters():
        gref = sd[k].grad
        if gref is None:
            assert p.grad is None, k
            continue
        e = rel(p.grad, gref)
        worst = max(worst, e)
        assert e < G_TOL, (k, e)
        n += 1
    assert n == 85
    print(f"full-size gradient check: worst relative-to-max error {worst:.2e}")


@pytest.mark.parametrize("heads", [1, 2, 8])
def test_other_head_counts_forward_and_gradients_match_reference_golden(golden_dir, heads):
    """EGNN(num_heads=h), h in {1, 2, 8} (E_GCL(num_heads=...) 3dm:186-207): the tensor-core kernels read the
    heads' second Linear as one block-diagonal 32 x 32 matrix, so every head count dividing 32 runs on the same code.
    Outputs of all three tensor-core arms and the gradients w.r.t. h, x and every parameter against the reference's own
    EGNN class with that head count (tests/golden/heads_h.pt), through the module API; the CUDA-core impls refuse."""
    g = torch.load(os.path.join(golden_dir, "heads_%d.pt" % heads), weights_only=False, map_location="cpu")
    egnn = P.EGNN(32, 32, 32, in_edge_nf=1, device=DEV, n_layers=2, num_heads=heads)
    egnn.load_state_dict(g["state_dict"], strict=True)
    edges = [g["row"].to(DEV), g["col"].to(DEV)]
    ea = g["edge_attr"].to(DEV)
    hscale = float(g["h_out"].abs().max())
    with torch.no_grad():
        for impl, tol in ((0, 1e-4), (4, 3e-3), (5, 2e-2)):
            egnn.impl = impl
            ho, xo = egnn(g["h"].to(DEV), g["x"].to(DEV), edges, ea)
            assert float((ho.cpu() - g["h_out"]).abs().max()) <= tol * hscale, (impl, float((ho.cpu() - g["h_out"]).abs().max()))
            assert float((xo.cpu() - g["x_out"]).abs().max()) <= tol * max(1.0, float(g["x_out"].abs().max())), impl
        egnn.impl = 1
        with pytest.raises(NotImplementedError):
            egnn(g["h"].to(DEV), g["x"].to(DEV), edges, ea)
        egnn.impl = 0
    h, x = g["h"].to(DEV).requires_grad_(True), g["x"].to(DEV).requires_grad_(True)
    ho, xo = egnn(h, x, edges, ea)
    ((ho * g["dh"].to(DEV)).sum() + (xo * g["dx"].to(DEV)).sum()).backward()
    assert rel(h.grad, g["grad_h"]) < G_TOL and rel(x.grad, g["grad_x"]) < G_TOL
    for k, p in egnn.named_parameters():
        assert p.grad is not None and rel(p.grad, g["grads"][k]) < G_TOL, (k, rel(p.grad, g["grads"][k]))


def test_graphed_train_step_with_two_heads_equals_eager():
    """The whole training step (fused losses, FlatState gathers with the full-matrix layout of the second edge Linear) on a
    model with num_heads=2: the graphed step follows the eager nn.Module / autograd step; an engine on the same model runs
    the inference path."""
    keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
    batches = [tuple(P.synthetic.make_batch(70 + i, 2, n=384)[k].to(DEV) for k in keys) for i in range(2)]
    ones = torch.ones(2, 384 * 16, 1, device=DEV)
    losses = {}
    for mode in ("eager", "graph"):
        torch.manual_seed(11)
        model = P.build_model(None, device=DEV, variant="train", num_heads=2)
        with torch.no_grad():
            model.egnn.embedding_out.weight.mul_(0.05); model.egnn.embedding_out.bias.mul_(0.05)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
        out = []
        if mode == "graph":
            step = P.train.GraphedTrainStep(model, opt, batches[0], k=16, warmup=2)
            for i in range(4):
                out.append(float(step(batches[i % 2])))
            assert sum(p.grad is not None for p in model.parameters()) == len(list(model.egnn.parameters())) + len(list(model.mlp.parameters()))
            sf, sp, tf, tp, corr, labels, gt = batches[0]
            eng = P.RegistrationEngine(model, batch=2, n=384, k=16, use_graph=False)
            R, t = eng.register(sf, sp, tf, tp, labels, gt)
            assert torch.isfinite(R).all() and torch.isfinite(t).all()
        else:
            for i in range(4):
                sf, sp, tf, tp, corr, labels, gt = batches[i % 2]
                es, et = P.knn_graph_batch(sp, 16), P.knn_graph_batch(tp, 16)
                out.append(float(P.train.train_step(model, opt, (sf, sp, es, ones, tf, tp, et, ones, corr, labels, gt))))
        losses[mode] = out
    assert np.all(np.isfinite(losses["graph"]))
    assert np.allclose(losses["eager"], losses["graph"], rtol=2e-3), losses


def test_graphed_train_step_pipelined_graph_build_equals_plain(golden_dir):
    """step(batch, next_batch=...) builds the next batch's k-NN graph on a forked stream inside the current step's CUDA
    graph (the wide k-NN / CSR kernels run beside the narrow head kernels); the losses and the final parameters must equal
    the plain step(batch) sequence -- also when an announcement is wrong or missing (the step then builds its own graph)."""
    keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
    batches = [tuple(P.synthetic.make_batch(60 + i, 2, n=512)[k].to(DEV) for k in keys) for i in range(3)]
    order = [0, 1, 2, 0, 2, 1, 1, 0]
    res = {}
    for mode in ("plain", "pipelined"):
        model = _model(golden_dir, 0.005)
        opt = torch.optim.Adam(model.parameters(), lr=1e-4, capturable=True)
        step = P.train.GraphedTrainStep(model, opt, batches[0], k=16, warmup=2)
        out = []
        for j, bi in enumerate(order):
            if mode == "plain":
                out.append(float(step(batches[bi])))
            else:
                nxt = batches[order[j + 1]] if j + 1 < len(order) else None
                if j == 4:
                    nxt = batches[0]             # a wrong announcement: the next call gets another batch and must rebuild
                out.append(float(step(batches[bi], next_batch=nxt)))
        res[mode] = (out, torch.cat([p.detach().flatten() for p in model.parameters()]).cpu())
    assert np.allclose(res["plain"][0], res["pipelined"][0], rtol=1e-5), res
    assert float((res["plain"][1] - res["pipelined"][1]).abs().max()) < 1e-6
