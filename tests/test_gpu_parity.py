"""GPU parity tests: the CUDA path (through the C ABI / the module mirror) against the oracle, the
golden vectors produced by the reference's own classes, and size-independent properties at the
full BASELINE sizes.  Tolerances (north_star / SURVEY 8(d)):
  k-NN ids, CSR            bit-exact
  features h               max|d| <= 1e-4 * max|h_ref|           (fp32 path)
  coordinates x            max|d| <= 1e-4 * max(1, max|x_ref|) m  (1e-4 m at 3DMatch extents)
  pose (well-conditioned)  <= 0.01 deg rotation, <= 1e-4 m translation (x extent for KITTI)
"""
import math
import os

import numpy as np
import pytest
import torch

import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
from oracle import egnn_oracle as O
from oracle import knn_oracle

pytestmark = pytest.mark.gpu
# stated bounds of the reduced-precision edge modes (BASELINE config 2), set from tools/fast_mode_err.py on a B200
# (measured: bf16 h <= 1.03e-2 of max|h|, x <= 1.1e-2 of the extent; train-variant rot <= 0.21 deg, H <= 1.5e-2)
BF16_H_TOL, BF16_X_TOL = 2e-2, 2e-2
RED_ROT_DEG, RED_T_M, RED_H_REL = 0.5, 2e-2, 5e-2
DEV = "cuda:0"
H_TOL, X_TOL, ROT_TOL_DEG, T_TOL = 1e-4, 1e-4, 0.01, 1e-4
CASES = ["small_b2_n256", "dup_b2_n512", "full_b1_n2048", "kitti_b1_n1024", "noenc_b1_n512"]


def rot_angle_deg(Ra, Rb):
    """Geodesic angle between two rotations from the chord |Ra-Rb|_F = 2*sqrt(2)*sin(theta/2)
    (well-conditioned near 0, unlike acos of the trace)."""
    d = np.linalg.norm(np.asarray(Ra, np.float64) - np.asarray(Rb, np.float64))
    return math.degrees(2.0 * math.asin(min(1.0, d / (2.0 * math.sqrt(2.0)))))


def load_case(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False, map_location="cpu")
    ck = os.path.join(golden_dir, g["meta"]["checkpoint"].split("/")[-1])
    return g, ck


@pytest.fixture(scope="module")
def model(golden_dir):
    return P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device=DEV)


def edges_of(nbr):
    return torch.stack([torch.stack(O.edges_from_nbr(n)) for n in nbr])


# ---------------------------------------------------------------------------------------------
# k-NN (a1) + edge layout (a2)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n,k,kind", [(16, 16, "uniform"), (17, 16, "uniform"), (33, 8, "uniform"), (257, 16, "dup"),
                                      (2048, 16, "uniform"), (2048, 16, "dup"), (2048, 32, "uniform"), (4096, 16, "kitti"),
                                      (8192, 16, "kitti"), (5, 16, "uniform"), (3000, 16, "allsame"), (2048, 16, "planar"),
                                      (2048, 16, "clustered"), (20000, 16, "uniform"), (1, 1, "uniform")])
def test_knn_bit_exact(n, k, kind):
    rng = np.random.default_rng(n * 7 + k)
    C = 3
    if kind == "kitti":
        x = (rng.random((C, n, 3)) * np.array([100, 100, 6])).astype(np.float32)
    else:
        x = (rng.random((C, n, 3)) * 3).astype(np.float32)
    if kind == "dup":                       # sampling with replacement: >= 30 % exact duplicates
        idx = rng.integers(0, n, (C, n // 3))
        for c in range(C):
            x[c, rng.choice(n, n // 3, replace=False)] = x[c, idx[c]]
    if kind == "allsame":
        x[:] = 1.25
    if kind == "planar":
        x[:, :, 2] = 0.75                     # flat cloud: one degenerate grid axis
    if kind == "clustered":                   # a few tight clusters far apart + exact duplicates across clusters
        centres = rng.random((C, 8, 3)) * 50
        x = (centres[:, rng.integers(0, 8, n)][np.arange(C)[:, None], np.arange(n)[None] % 1, :] if False else
             np.stack([centres[c][rng.integers(0, 8, n)] + rng.standard_normal((n, 3)) * 0.01 for c in range(C)])).astype(np.float32)
    ref = knn_oracle.knn(x, k)
    for brute in (False, True):              # cell-grid search and brute-force scan: identical ids
        got = ops.knn_build(torch.from_numpy(x).to(DEV), k, brute_force=brute).cpu().numpy()
        assert np.array_equal(got, ref), ("brute" if brute else "grid")
    if n >= k and kind == "uniform":
        assert (got[:, :, 0] == np.arange(n)[None]).all()       # self is the nearest (loop=True)


def test_knn_graph_matches_torch_cluster_layout():
    x = torch.rand(300, 3, device=DEV)
    e = P.knn_graph(x, 16, loop=True)
    assert e.dtype == torch.int64 and tuple(e.shape) == (2, 300 * 16)
    ref = knn_oracle.knn(x.cpu().numpy(), 16)
    assert torch.equal(e[0].cpu(), torch.from_numpy(ref).reshape(-1).long())
    assert torch.equal(e[1].cpu(), torch.arange(300).repeat_interleave(16))
    eb = P.knn_graph_batch(torch.stack([x, x.flip(0)]), 16)
    assert torch.equal(eb[0], e) and tuple(eb.shape) == (2, 2, 4800)
    with pytest.raises(Exception):
        ops.knn_build(x[None], 64)           # k > EGSPR_MAX_K -> EGSPR_E_UNSUPPORTED


# ---------------------------------------------------------------------------------------------
# CSR transpose + unsorted_segment_sum (a9)
# ---------------------------------------------------------------------------------------------
def _csr_reference(row, col, C, N):
    E = row.shape[1]
    g_row = (row + (torch.arange(C) * N)[:, None]).reshape(-1)
    g_col = (col + (torch.arange(C) * N)[:, None]).reshape(-1)
    eid = torch.arange(E).repeat(C)
    order = torch.sort(g_row, stable=True).indices
    ptr = torch.zeros(C * N + 1, dtype=torch.long)
    ptr[1:] = torch.bincount(g_row, minlength=C * N).cumsum(0)
    return ptr, g_row[order], g_col[order], eid[order]


@pytest.mark.parametrize("C,N,E", [(1, 50, 800), (3, 257, 4112), (2, 64, 7000)])
def test_csr_from_edges_exact(C, N, E):
    g = torch.Generator().manual_seed(C * 1000 + N)
    row = torch.randint(0, N, (C, E), generator=g)
    col = torch.randint(0, N, (C, E), generator=g)
    if N == 64:
        row[:, : E // 2] = 7                  # one very high in-degree row (> 32 entries path)
    edges = torch.stack([row, col], 1).to(DEV)
    gr = ops.csr_from_edges(edges, N)
    ptr, r, c, e = _csr_reference(row, col, C, N)
    assert torch.equal(gr.ptr.cpu().long(), ptr) and torch.equal(gr.row.cpu().long(), r)
    assert torch.equal(gr.col.cpu().long(), c) and torch.equal(gr.eid.cpu().long(), e)
    gr.check()
    bad = edges.clone(); bad[0, 0, 0] = N + 5
    with pytest.raises(IndexError):
        ops.csr_from_edges(bad, N).check()


def test_unsorted_segment_sum_matches_and_is_deterministic():
    g = torch.Generator().manual_seed(0)
    data = torch.randn(5000, 35, generator=g)
    ids = torch.randint(0, 300, (5000,), generator=g)
    ref = O.segment_sum(data.double(), ids, 300)
    a = P.unsorted_segment_sum(data.to(DEV), ids.to(DEV), 300)
    b = P.unsorted_segment_sum(data.to(DEV), ids.to(DEV), 300)
    assert torch.equal(a, b)
    assert torch.allclose(a.cpu().double(), ref, atol=1e-4)
    # ascending-edge order == the order scatter_add_ visits on CPU -> bit-identical to the fp32 oracle
    assert torch.equal(a.cpu(), O.segment_sum(data, ids, 300))


# ---------------------------------------------------------------------------------------------
# EGNN / E_GCL modules (a3-a11) vs the reference golden vectors
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("impl", [0, 1, 2, 3])
def test_forward_eval_matches_reference_golden(golden_dir, name, impl):
    g, ck = load_case(golden_dir, name)
    model = P.build_model(ck, device=DEV, variant="eval")
    model.egnn.impl = impl
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    B, N = inp["labels"].shape
    es, et = P.knn_graph_batch(inp["src_pts"], 16), P.knn_graph_batch(inp["tgt_pts"], 16)
    assert torch.equal(es.cpu(), edges_of(g["nbr_src"])) and torch.equal(et.cpu(), edges_of(g["nbr_tgt"]))
    ea = torch.ones(B, es.shape[-1], 1, device=DEV)                       # get_edges_batch 3dm:387
    with torch.no_grad():
        out = model(inp["src_feat"], inp["src_pts"], es, ea, inp["tgt_feat"], inp["tgt_pts"], et, ea,
                    inp["corr"], inp["labels"], inp["gt_pose"])
    ref = g["eval_f32"]
    assert out[2] is None and len(out) == 9 and out[8] is inp["labels"]
    for i, key in ((4, "h_src"), (6, "h_tgt")):
        assert float((out[i].cpu() - ref[key]).abs().max()) <= H_TOL * float(ref[key].abs().max()), key
    for i, key in ((5, "x_src"), (7, "x_tgt")):
        assert float((out[i].cpu() - ref[key]).abs().max()) <= X_TOL * max(1.0, float(ref[key].abs().max())), key
    scale = max(1.0, float(g["inputs"]["tgt_pts"].abs().max()))
    for b in range(B):
        assert rot_angle_deg(out[0][b].cpu().numpy(), ref["R"][b].numpy()) <= ROT_TOL_DEG
        assert float((out[1][b].cpu() - ref["t"][b]).abs().max()) <= T_TOL * scale
    assert abs(float(out[3]) - float(ref["equi_loss"].mean())) <= 1e-4 * abs(float(ref["equi_loss"].mean()))
    # metrics of the eval loop (tools/evaluation_metrics.py) agree with the reference's pose
    m_ours = P.metrics.evaluate_batch(out[0].cpu().numpy(), out[1].cpu().numpy(), g["inputs"]["gt_pose"].numpy(),
                                      g["inputs"]["src_pts"].numpy(), g["inputs"]["tgt_pts"].numpy())
    m_ref = P.metrics.evaluate_batch(ref["R"].numpy(), ref["t"].numpy(), g["inputs"]["gt_pose"].numpy(),
                                     g["inputs"]["src_pts"].numpy(), g["inputs"]["tgt_pts"].numpy())
    assert np.allclose(m_ours["recall"], m_ref["recall"], atol=0.02) and np.allclose(m_ours["trans_err"], m_ref["trans_err"], atol=0.05 * scale)


def test_per_layer_states_match_reference(golden_dir, model):
    g, _ = load_case(golden_dir, "full_b1_n2048")
    layers, pin, pout = model.egnn.packs()
    inp = g["inputs"]
    gr = ops.csr_from_nbr(g["nbr_src"][:1].to(DEV))
    h, x, lay = ops.egnn_forward(inp["src_feat"][:1].to(DEV), inp["src_pts"][:1].to(DEV), gr, layers, pin, pout, return_layers=True)
    ref = g["eval_f32"]["layers_src0"]
    for i, (hl, xl) in enumerate(lay):
        assert float((hl[0].cpu() - ref[i + 1][0]).abs().max()) <= H_TOL * float(ref[i + 1][0].abs().max())
        assert float((xl[0].cpu() - ref[i + 1][1]).abs().max()) <= X_TOL
    # distance to the fp64 reference: the 3xTF32 tensor-core path (dropped lo*lo term, tensor-core accumulation)
    # measures ~4e-6 of max|h| on a B200 against ~4e-7 for the reference's own fp32 run; both are far inside
    # the 1e-4 parity bar.  Bound it at 1e-5 so a precision regression (e.g. a lost lo term) is caught.
    r64 = g["eval_f64"]["h_src"][0]
    ours = float((h[0].cpu() - r64).abs().max())
    assert ours <= 1e-5 * float(r64.abs().max())


@torch.no_grad()
def test_egnn_and_egcl_module_signatures(golden_dir, model):
    g, _ = load_case(golden_dir, "small_b2_n256")
    inp = g["inputs"]
    row, col = O.edges_from_nbr(g["nbr_src"][0])
    E = row.shape[0]
    sd = {k: v.cpu() for k, v in model.egnn.state_dict().items()}
    h_in, x_in = inp["src_feat"][0], inp["src_pts"][0]
    # EGNN.forward(h, x, [row, col], edge_attr) with list edges, as 3dm:662
    h, x = model.egnn(h_in.to(DEV), x_in.to(DEV), [row.to(DEV), col.to(DEV)], torch.ones(E, 1, device=DEV))
    href, xref = O.egnn_forward(sd, h_in, x_in, row, col, torch.ones(E, 1))
    assert float((h.cpu() - href).abs().max()) <= H_TOL * float(href.abs().max()) and float((x.cpu() - xref).abs().max()) <= X_TOL
    # a non-constant edge_attr goes through the csr_eid gather
    ea = torch.rand(E, 1)
    h, x = model.egnn(h_in.to(DEV), x_in.to(DEV), torch.stack([row, col]).to(DEV), ea.to(DEV))
    href, xref = O.egnn_forward(sd, h_in, x_in, row, col, ea)
    assert float((h.cpu() - href).abs().max()) <= H_TOL * float(href.abs().max()) and float((x.cpu() - xref).abs().max()) <= X_TOL
    # E_GCL.forward(h, edge_index, coord, edge_attr) -> (h, coord, edge_attr), shuffled edge order, ragged degrees
    perm = torch.randperm(E)[: E - 37]
    r2, c2, ea2 = row[perm], col[perm], ea[perm]
    gcl = model.egnn.gcl_1
    hh = torch.randn(256, 32)
    h1, x1, ea_out = gcl(hh.to(DEV), [r2.to(DEV), c2.to(DEV)], x_in.to(DEV), edge_attr=ea2.to(DEV))
    h1r, x1r, _ = O.egcl_forward(sd, "gcl_1.", hh, x_in, r2, c2, ea2)
    assert float((h1.cpu() - h1r).abs().max()) <= H_TOL * float(h1r.abs().max()) and float((x1.cpu() - x1r).abs().max()) <= X_TOL
    assert ea_out.shape == ea2.shape


@pytest.mark.parametrize("impl", [1, 2, 3, 4])
def test_twin_points_stay_bit_identical(model, impl):
    """Exact duplicate correspondences ("twins": same coordinates AND features, normal in the datasets,
    datasets/ThreeDMatch.py:319,329) whose incoming-edge sets coincide must stay bit-identical through
    every layer: the degenerate-frame rule (3dm:152-163) is discontinuous at x_i == x_j, so a
    rounding-level split would change the messages by O(1).  Guaranteed by summing every node strictly
    in ascending edge order (as scatter_add_ does) and by position-independent per-edge arithmetic."""
    rng = np.random.default_rng(9)
    n = 1000
    x = (rng.random((n, 3)) * 3).astype(np.float32)
    f = rng.standard_normal((n, 32)).astype(np.float32)
    src = rng.integers(0, n // 2, 150); dst = rng.choice(np.arange(n // 2, n), 150, replace=False)
    x[dst] = x[src]; f[dst] = f[src]
    xt, ft = torch.from_numpy(x)[None].to(DEV), torch.from_numpy(f)[None].to(DEV)
    nbr = ops.knn_build(xt, 16)
    gr = ops.csr_from_nbr(nbr)
    layers, pin, pout = model.egnn.packs()
    h, xo = ops.egnn_forward(ft, xt, gr, layers, pin, pout, impl=impl)
    row = nbr[0].cpu().long().reshape(-1); col = torch.arange(n).repeat_interleave(16)
    insets = [set() for _ in range(n)]
    for r, c in zip(row.tolist(), col.tolist()):
        insets[r].add(c)
    checked = 0
    for s_, d_ in zip(src.tolist(), dst.tolist()):
        if insets[s_] - {s_, d_} == insets[d_] - {s_, d_} and (s_ in insets[s_]) == (s_ in insets[d_]) and (d_ in insets[s_]) == (d_ in insets[d_]):
            assert torch.equal(h[0, s_], h[0, d_]) and torch.equal(xo[0, s_], xo[0, d_])
            checked += 1
    assert checked > 20
    sd = {k: v.cpu() for k, v in model.egnn.state_dict().items()}
    href, xref = O.egnn_forward(sd, torch.from_numpy(f), torch.from_numpy(x), row, col, torch.ones(n * 16, 1))
    loose = 30.0 if impl == 4 else 1.0          # impl 4 = reduced-precision edge mode, stated bound 3e-3
    assert float((h[0].cpu() - href).abs().max()) <= loose * H_TOL * float(href.abs().max())
    assert float((xo[0].cpu() - xref).abs().max()) <= loose * X_TOL * max(1.0, float(xref.abs().max()))


def test_forward_train_variant(golden_dir):
    for name in ("small_b2_n256", "noenc_b1_n512"):
        g, ck = load_case(golden_dir, name)
        model = P.build_model(ck, device=DEV, variant="train")
        inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
        es, et = P.knn_graph_batch(inp["src_pts"], 16), P.knn_graph_batch(inp["tgt_pts"], 16)
        with torch.no_grad():
            out = model(inp["src_feat"], inp["src_pts"], es, None, inp["tgt_feat"], inp["tgt_pts"], et, None,
                        inp["corr"], inp["labels"], inp["gt_pose"])
        ref = g["train_f32"]
        assert float((out[4].cpu() - ref["h_src"]).abs().max()) <= H_TOL * float(ref["h_src"].abs().max())
        assert abs(float(out[2]) - float(ref["slot2"])) <= 2e-4 * abs(float(ref["slot2"]))          # corr_loss + sim_loss
        assert abs(float(out[3]) - float(ref["equi_loss"])) <= 1e-4 * abs(float(ref["equi_loss"]))
        # with the shipped weights the train-variant H is ~1e-6*I (one-hot softmax) and R is
        # LAPACK's arbitrary basis -> compare weights and H against the oracle, not R
        sd = torch.load(ck, map_location="cpu", weights_only=True)["cross_attention_state_dict"]
        _, aux = O.forward_train(sd, g["inputs"]["src_feat"], g["inputs"]["src_pts"], es.cpu(), g["inputs"]["tgt_feat"],
                                 g["inputs"]["tgt_pts"], et.cpu(), g["inputs"]["labels"], g["inputs"]["gt_pose"], return_aux=True)
        for b in range(out[0].shape[0]):
            wref = torch.zeros(inp["labels"].shape[1]); wref[aux["valid"][b]] = aux["w"][b]
            assert float((model.last_aux["w"][b].cpu() - wref).abs().max()) <= 1e-3 * float(wref.max())
            assert float((model.last_aux["H"][b].cpu() - aux["H"][b]).abs().max()) <= 1e-3 * float(aux["H"][b].abs().max())
            assert abs(float(torch.det(out[0][b].cpu().double())) - 1.0) < 1e-4


def test_forward_train_variant_kitti_top_k(golden_dir):
    """kit:616-766 is the train-variant body with top_k = 2048 (= every point at the KITTI loader's size): here
    top_k = N on the 256-point fixture -- slot 2 (BCE over ALL points + similarity loss) against the oracle."""
    g, ck = load_case(golden_dir, "small_b2_n256")
    model = P.build_model(ck, device=DEV, variant="train")
    model.top_k = 256
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    es, et = P.knn_graph_batch(inp["src_pts"], 16), P.knn_graph_batch(inp["tgt_pts"], 16)
    with torch.no_grad():
        out = model(inp["src_feat"], inp["src_pts"], es, None, inp["tgt_feat"], inp["tgt_pts"], et, None,
                    inp["corr"], inp["labels"], inp["gt_pose"])
    sd = torch.load(ck, map_location="cpu", weights_only=True)["cross_attention_state_dict"]
    ref = O.forward_train(sd, g["inputs"]["src_feat"], g["inputs"]["src_pts"], es.cpu(), g["inputs"]["tgt_feat"],
                          g["inputs"]["tgt_pts"], et.cpu(), g["inputs"]["labels"], g["inputs"]["gt_pose"], top_k=256)
    assert abs(float(out[2]) - float(ref[2])) <= 2e-4 * abs(float(ref[2]))
    assert abs(float(out[3]) - float(ref[3])) <= 1e-4 * abs(float(ref[3]))


# ---------------------------------------------------------------------------------------------
# Kabsch (a15)
# ---------------------------------------------------------------------------------------------
def test_kabsch_against_oracle_and_edge_cases():
    rng = np.random.default_rng(5)
    B, n = 6, 500
    p = torch.tensor(rng.random((B, n, 3)) * 3, dtype=torch.float32)
    q = torch.empty_like(p)
    for b in range(B):
        R = torch.tensor(P.synthetic.random_rotation(rng), dtype=torch.float32)
        q[b] = p[b] @ R.T + torch.tensor(rng.random(3), dtype=torch.float32) + 0.01 * torch.randn(n, 3)
    w = torch.rand(B, n); w = w / w.sum(1, keepdim=True)
    mask = (torch.rand(B, n) < 0.7).float()
    mask[1] = 0                                   # empty set -> identity (3dm:708-711)
    q[2] = p[2] * torch.tensor([-1.0, 1.0, 1.0])  # pure reflection -> det fix path
    p[3, :, 2] = 0.5; q[3] = p[3] @ torch.tensor(P.synthetic.random_rotation(rng), dtype=torch.float32).T   # planar (rank 2)
    R, t, Hm = ops.kabsch(p.to(DEV), q.to(DEV), w.to(DEV), mask.to(DEV))
    for b in range(B):
        sel = mask[b].bool()
        Rr, tr, Hr = O.kabsch(p[b][sel].double(), q[b][sel].double(), w[b][sel].double())
        assert abs(float(torch.det(R[b].cpu().double())) - 1.0) < 1e-5
        if b == 1:
            assert torch.equal(R[b].cpu(), torch.eye(3)) and float(t[b].abs().sum()) == 0
            continue
        assert float((Hm[b].cpu().double() - Hr).abs().max()) <= 1e-5 * float(Hr.abs().max())
        if b == 2:
            continue                              # reflection: the optimal rotation is not unique across SVD bases
        assert rot_angle_deg(R[b].cpu().numpy(), Rr.numpy()) <= ROT_TOL_DEG
        assert float((t[b].cpu().double() - tr).abs().max()) <= T_TOL
    # well-separated singular values after a reflection fix: compare with the oracle too
    pr = torch.tensor(rng.random((1, 200, 3)), dtype=torch.float32) * torch.tensor([3.0, 2.0, 1.0])
    qr = pr * torch.tensor([1.0, 1.0, -1.0])
    wr = torch.full((1, 200), 1 / 200)
    R, t, _ = ops.kabsch(pr.to(DEV), qr.to(DEV), wr.to(DEV))
    Rr, tr, _ = O.kabsch(pr[0].double(), qr[0].double(), wr[0].double())
    assert rot_angle_deg(R[0].cpu().numpy(), Rr.numpy()) <= ROT_TOL_DEG


# ---------------------------------------------------------------------------------------------
# equivariance: (i) ours == reference on transformed inputs is covered by the golden
# tests; (ii) strict E(3) equivariance once the 10 non-invariant input columns are zeroed
# ---------------------------------------------------------------------------------------------
def test_equivariance_with_invariant_columns_only(golden_dir):
    model = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device=DEV)
    with torch.no_grad():
        for i in range(3):
            for m in model.egnn._modules["gcl_%d" % i].edge_mlps:
                m[0].weight[:, 66:76] = 0          # dot product (66) and the raw frame components (67-75)
    rng = np.random.default_rng(2)
    x = torch.tensor(rng.random((1, 1024, 3)) * 3, dtype=torch.float32)
    h = torch.nn.functional.normalize(torch.randn(1, 1024, 32), dim=-1)
    R = torch.tensor(P.synthetic.random_rotation(rng), dtype=torch.float32)
    t = torch.tensor([0.3, -1.2, 0.7])
    gr = ops.csr_from_nbr(ops.knn_build(x.to(DEV), 16))     # graph held fixed
    h1, x1 = model.egnn.forward_batch(h.to(DEV), x.to(DEV), gr)
    h2, x2 = model.egnn.forward_batch(h.to(DEV), (x @ R.T + t).to(DEV), gr)
    assert float((h1 - h2).abs().max()) <= 1e-4 * float(h1.abs().max())
    assert float((x1.cpu() @ R.T + t - x2.cpu()).abs().max()) <= 2e-4
    # and the unmodified model is NOT equivariant (documented reference behaviour, F5)
    ref_model = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device=DEV)
    a, _ = ref_model.egnn.forward_batch(h.to(DEV), x.to(DEV), gr)
    b, _ = ref_model.egnn.forward_batch(h.to(DEV), (x @ R.T + t).to(DEV), gr)
    assert float((a - b).abs().max()) > 1e-2 * float(a.abs().max())


@pytest.mark.parametrize("name", CASES)
def test_reduced_precision_edge_mode_within_looser_bound(golden_dir, name):
    """impl 4 (BASELINE config 2's reduced-precision edge MLP: single-pass TF32 operands + MUFU.TANH SiLU).
    Stated looser bound: features 3e-3 of max|h|, coordinates 3e-3 * max(1, max|x|) m (measured on B200:
    h <= 1.1e-3, x <= 1.2e-2 m at 3DMatch extents); the eval-variant pose keeps the fp32 bars because it
    is solved on the ORIGINAL coordinates with near-uniform weights (evl:717-718)."""
    g, ck = load_case(golden_dir, name)
    model = P.build_model(ck, device=DEV, variant="eval")
    model.egnn.impl = 4
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    es, et = P.knn_graph_batch(inp["src_pts"], 16), P.knn_graph_batch(inp["tgt_pts"], 16)
    with torch.no_grad():
        out = model(inp["src_feat"], inp["src_pts"], es, None, inp["tgt_feat"], inp["tgt_pts"], et, None,
                    inp["corr"], inp["labels"], inp["gt_pose"])
    ref = g["eval_f32"]
    for i, key in ((4, "h_src"), (6, "h_tgt")):
        err = float((out[i].cpu() - ref[key]).abs().max())
        assert 1e-6 * float(ref[key].abs().max()) < err <= 3e-3 * float(ref[key].abs().max()), key   # looser, and really the other path
    for i, key in ((5, "x_src"), (7, "x_tgt")):
        assert float((out[i].cpu() - ref[key]).abs().max()) <= 3e-3 * max(1.0, float(ref[key].abs().max())), key
    scale = max(1.0, float(g["inputs"]["tgt_pts"].abs().max()))
    for b in range(inp["labels"].shape[0]):
        assert rot_angle_deg(out[0][b].cpu().numpy(), ref["R"][b].numpy()) <= ROT_TOL_DEG
        assert float((out[1][b].cpu() - ref["t"][b]).abs().max()) <= T_TOL * scale


@pytest.mark.parametrize("name", CASES)
def test_bf16_edge_mode_within_stated_bound(golden_dir, name):
    """impl 5 (BASELINE config 2's bf16 edge MLP: tcgen05.mma.kind::f16, bf16 activations and weights, fp32 accumulation,
    geometric inputs as two bf16 terms, tanh SiLU).  Stated bound: features BF16_H_TOL of max|h|, coordinates
    BF16_X_TOL * max(1, max|x|) m; the eval-variant pose keeps the fp32 bars (original coordinates, near-uniform weights).
    On tempered weights (well-conditioned train-variant Kabsch) the TRAIN-variant pose and H of both reduced modes stay
    within RED_ROT_DEG / RED_T_M / RED_H_REL of the fp32 path."""
    g, ck = load_case(golden_dir, name)
    model = P.build_model(ck, device=DEV, variant="eval")
    model.egnn.impl = 5
    inp = {k: v.to(DEV) for k, v in g["inputs"].items()}
    es, et = P.knn_graph_batch(inp["src_pts"], 16), P.knn_graph_batch(inp["tgt_pts"], 16)
    with torch.no_grad():
        out = model(inp["src_feat"], inp["src_pts"], es, None, inp["tgt_feat"], inp["tgt_pts"], et, None,
                    inp["corr"], inp["labels"], inp["gt_pose"])
    ref = g["eval_f32"]
    for i, key in ((4, "h_src"), (6, "h_tgt")):
        err = float((out[i].cpu() - ref[key]).abs().max())
        assert 1e-5 * float(ref[key].abs().max()) < err <= BF16_H_TOL * float(ref[key].abs().max()), (key, err / float(ref[key].abs().max()))
    for i, key in ((5, "x_src"), (7, "x_tgt")):
        assert float((out[i].cpu() - ref[key]).abs().max()) <= BF16_X_TOL * max(1.0, float(ref[key].abs().max())), key
    scale = max(1.0, float(g["inputs"]["tgt_pts"].abs().max()))
    for b in range(inp["labels"].shape[0]):
        assert rot_angle_deg(out[0][b].cpu().numpy(), ref["R"][b].numpy()) <= ROT_TOL_DEG
        assert float((out[1][b].cpu() - ref["t"][b]).abs().max()) <= T_TOL * scale
    # train variant on tempered weights: pose / H of the reduced modes against the fp32 path
    tm = P.build_model(ck, device=DEV, variant="train")
    with torch.no_grad():
        tm.egnn.embedding_out.weight.mul_(0.005); tm.egnn.embedding_out.bias.mul_(0.005)
    res = {}
    for impl in (3, 4, 5):
        tm.egnn.impl = impl
        with torch.no_grad():
            o = tm(inp["src_feat"], inp["src_pts"], es, None, inp["tgt_feat"], inp["tgt_pts"], et, None, inp["corr"], inp["labels"], inp["gt_pose"])
        res[impl] = (o[0].cpu(), o[1].cpu(), tm.last_aux["H"].cpu())
    for impl in (4, 5):
        for b in range(inp["labels"].shape[0]):
            assert rot_angle_deg(res[impl][0][b].numpy(), res[3][0][b].numpy()) <= RED_ROT_DEG, (impl, b)
        assert float((res[impl][1] - res[3][1]).abs().max()) <= RED_T_M * scale, impl
        assert float((res[impl][2] - res[3][2]).abs().max()) <= RED_H_REL * float(res[3][2].abs().max()), impl


# ---------------------------------------------------------------------------------------------
# engine (fused launch sequence) == module API; full-size properties
# ---------------------------------------------------------------------------------------------
def test_engine_equals_module_api_and_oracle_full_size(golden_dir, model):
    B, N = 4, 2048
    data = P.synthetic.make_batch(21, B, n=N, dup_frac=0.1)
    eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=True)
    R, t = eng.register(*[data[k] for k in ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")])
    o = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in eng.outputs().items()}
    d = {k: v.to(DEV) for k, v in data.items()}
    es, et = P.knn_graph_batch(d["src_pts"], 16), P.knn_graph_batch(d["tgt_pts"], 16)
    model.variant = "eval"
    with torch.no_grad():
        out = model(d["src_feat"], d["src_pts"], es, None, d["tgt_feat"], d["tgt_pts"], et, None, d["corr"], d["labels"], d["gt_pose"])
    assert torch.equal(out[0], o["R"]) and torch.equal(out[1], o["t"]) and torch.equal(out[4], o["h_src"])   # same kernels, same order
    # oracle on pair 0
    sd = {k: v.cpu() for k, v in model.state_dict().items()}
    ref = O.forward_eval(sd, data["src_feat"][:1], data["src_pts"][:1], es[:1].cpu(), data["tgt_feat"][:1], data["tgt_pts"][:1],
                         et[:1].cpu(), data["labels"][:1], data["gt_pose"][:1])
    assert rot_angle_deg(o["R"][0].cpu().numpy(), ref[0][0].numpy()) <= ROT_TOL_DEG
    assert float((o["t"][0].cpu() - ref[1][0]).abs().max()) <= T_TOL
    assert float((o["h_tgt"][0].cpu() - ref[6][0]).abs().max()) <= H_TOL * float(ref[6][0].abs().max())
    # determinism: replaying the graph is bit-identical (no atomics in the float path)
    eng.run(); torch.cuda.synchronize()
    assert torch.equal(eng.R, o["R"]) and torch.equal(eng.h_out[:B], o["h_src"]) and torch.equal(eng.x_out[B:], o["x_tgt"])


def test_full_batch_properties(model):
    """BASELINE size (64 pairs x 2048): pair independence / permutation, and poses are rigid."""
    B, N = 64, 2048
    data = P.synthetic.make_batch(31, B, n=N)
    keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")
    eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=False)
    eng.register(*[data[k] for k in keys])
    R1, t1, h1 = eng.R.clone(), eng.t.clone(), eng.h_out.clone()
    perm = torch.randperm(B)
    eng.register(*[data[k][perm] for k in keys])
    assert torch.equal(eng.R, R1[perm]) and torch.equal(eng.t, t1[perm])        # a pair's result ignores its neighbours
    assert torch.equal(eng.h_out[:B], h1[:B][perm])
    Rd = R1.double()
    assert float((Rd @ Rd.transpose(1, 2) - torch.eye(3, dtype=torch.float64, device=DEV)).abs().max()) < 1e-5
    assert float((torch.det(Rd) - 1).abs().max()) < 1e-5
    assert torch.isfinite(h1).all() and torch.isfinite(t1).all()
    # a single pair run alone gives the same answer as inside the batch
    e1 = P.RegistrationEngine(model, batch=1, n=N, k=16, use_graph=False)
    e1.register(*[data[k][7:8] for k in keys])
    assert torch.equal(e1.R[0], R1[7]) and torch.equal(e1.t[0], t1[7])


def test_engine_submit_collect_pipeline_matches_register(model):
    """Pipelined host-to-host path (two input sets, upload on a copy stream) == the plain register() call,
    batch after batch, in order."""
    B, N = 3, 1024
    keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")
    batches = [P.synthetic.make_batch(40 + i, B, n=N, pin=True) for i in range(5)]
    ref_eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=False)
    want = []
    for b in batches:
        R, t = ref_eng.register(*[b[k] for k in keys])
        want.append((R.cpu().clone(), t.cpu().clone()))
    eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=True)
    got, tk = [], eng.submit(*[batches[0][k] for k in keys])
    for b in batches[1:]:
        nxt = eng.submit(*[b[k] for k in keys])
        R, t = eng.collect(tk)
        got.append((R.clone(), t.clone()))
        tk = nxt
    R, t = eng.collect(tk)
    got.append((R.clone(), t.clone()))
    for (Rw, tw), (Rg, tg) in zip(want, got):
        assert torch.equal(Rw, Rg) and torch.equal(tw, tg)


def test_device_pose_metrics_match_reference_metrics(golden_dir, model):
    """egspr_pose_metrics (fp64 on the device) == tools/evaluation_metrics.py + evl:1277 (numpy port pinned by the
    known-answer fixture in test_oracle.py): RE/TE to 1e-9 relative, inlier counts exact."""
    B, N = 6, 2048
    data = P.synthetic.make_batch(77, B, n=N)
    keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")
    eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=False)
    R, t = eng.register(*[data[k] for k in keys])
    got = eng.metrics().cpu().numpy()
    Rn, tn = R.cpu().numpy(), t.cpu().numpy()
    ref = {k: [] for k in ("rot_err", "trans_err", "recall", "precision", "f1")}
    for b in range(B):                                        # the oracle's numpy restatement, pair by pair like evl:1249-1281
        T = np.eye(4); T[:3, :3] = Rn[b].astype(np.float64); T[:3, 3] = tn[b].astype(np.float64)
        gtp = data["gt_pose"][b].numpy().astype(np.float64)
        re, te = O.calculate_pose_error(gtp, T)
        rec, prec = O.registration_recall(gtp, T, data["src_pts"][b].numpy().astype(np.float64), data["tgt_pts"][b].numpy().astype(np.float64))
        for k, v in zip(ref, (re, te, rec, prec, 2 * prec * rec / (prec + rec + 1e-6))):
            ref[k].append(v)
    for j, key in enumerate(("rot_err", "trans_err", "recall", "precision", "f1")):
        assert np.allclose(got[:, j], np.asarray(ref[key], dtype=np.float64), rtol=1e-9, atol=1e-9), key
    # the module-level mirrors of the reference's two functions run the same kernel: known answers of the reference file
    for c in torch.load(os.path.join(golden_dir, "metrics_kat.pt"), weights_only=False):
        re, te = P.metrics.calculate_pose_error(c["gt"], c["pred"])
        rec, prec = P.metrics.registration_recall(c["gt"], c["pred"], c["src"], c["tgt"])
        assert np.isclose(re, c["re"], atol=2e-2) and np.isclose(te, c["te"], rtol=1e-5, atol=1e-4)       # fp32 pose in, fp64 math
        assert np.isclose(rec, c["recall"], atol=1e-6) and np.isclose(prec, c["precision"], atol=1e-6)
    # a pose that is exactly the ground truth: zero errors, recall = sqrt(fraction of inliers within tau)
    gt = data["gt_pose"].to(DEV)
    m = ops.pose_metrics(gt[:, :3, :3].contiguous(), gt[:, :3, 3].contiguous(), gt, data["src_pts"].to(DEV), data["tgt_pts"].to(DEV)).cpu().numpy()
    assert np.all(m[:, 0] < 0.05) and np.all(m[:, 1] < 1e-6) and np.all(m[:, 3] > 0.5)     # acos near 1: sqrt(fp32 eps) ~ 0.01 deg


@pytest.mark.parametrize("impl", [1, 3, 4])
def test_irregular_graph_hub_rows_and_isolated_nodes(model, impl):
    """User-supplied edge lists (the module API accepts any [row, col]): an aggregation row with hundreds of
    incoming edges (spans several 128-edge tiles of the streaming reduction -> the carry path), rows with
    no incoming edge at all (zero aggregate, unchanged coordinates), multi-edges, several clouds."""
    g = torch.Generator().manual_seed(5)
    C, N = 3, 300
    rows, cols = [], []
    for c in range(C):
        r = [torch.full((700,), 7 + c, dtype=torch.int64), torch.randint(0, N // 2, (900,), generator=g),   # hub row; rows N/2.. get nothing
             torch.full((130,), N // 2 - 1, dtype=torch.int64)]                                            # a second long row at the end of the used range
        r = torch.cat(r)
        cc = torch.randint(0, N, (r.numel(),), generator=g)
        perm = torch.randperm(r.numel(), generator=g)
        rows.append(r[perm]); cols.append(cc[perm])
    edges = torch.stack([torch.stack([rows[c], cols[c]]) for c in range(C)])          # [C,2,E]
    E = edges.shape[-1]
    feat = torch.randn(C, N, 32, generator=g); x = torch.rand(C, N, 3, generator=g) * 2
    ea = torch.rand(C, E, generator=g)
    gr = ops.csr_from_edges(edges.to(DEV), N)
    layers, pin, pout = model.egnn.packs()
    h, xo = ops.egnn_forward(feat.to(DEV), x.to(DEV), gr, layers, pin, pout, edge_attr=ea.to(DEV), impl=impl)
    sd = {k: v.cpu() for k, v in model.egnn.state_dict().items()}
    loose = 30.0 if impl == 4 else 1.0
    for c in range(C):
        href, xref = O.egnn_forward(sd, feat[c], x[c], rows[c], cols[c], ea[c][:, None])
        assert float((h[c].cpu() - href).abs().max()) <= loose * H_TOL * float(href.abs().max()), c
        assert float((xo[c].cpu() - xref).abs().max()) <= loose * X_TOL * max(1.0, float(xref.abs().max())), c


# ---------------------------------------------------------------------------------------------
# feature-space correspondence search (tcgen05 GEMM + fused argmin)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ns,nt", [(300, 257), (3000, 2500), (128, 4096)])
def test_feature_nn_matches_reference_correspondence_search(ns, nt):
    from oracle import feature_match_oracle as FO
    rng = np.random.default_rng(ns + nt)
    fs = rng.standard_normal((ns, 32)).astype(np.float32); fs /= np.linalg.norm(fs, axis=1, keepdims=True)
    ft = rng.standard_normal((nt, 32)).astype(np.float32); ft /= np.linalg.norm(ft, axis=1, keepdims=True)
    m = min(ns, nt) // 3
    ft[rng.permutation(nt)[:m]] = fs[rng.permutation(ns)[:m]]        # exact matches: distance sqrt(1e-6)
    ft[5] = ft[11]                                                   # duplicated target: the FIRST index must win
    corr_ref, idx_ref, dis_ref, D = FO.correspondences(fs, ft, use_mutual=False)
    idx, dis = ops.feature_nn(torch.from_numpy(fs).to(DEV), torch.from_numpy(ft).to(DEV))
    idx, dis = idx.cpu().numpy().astype(np.int64), dis.cpu().numpy()
    # distances: the two GEMMs round differently (BLAS fp32 vs 3xTF32 with the tensor core's truncating fp32
    # accumulation: measured bias ~1e-6 on <a,a> = 1), i.e. 2 - 2s moves by a few 1e-6; compare in that domain
    # (near an exact match sqrt amplifies it: the reference's own d scatters between 8.7e-4 and 1.06e-3 there)
    assert np.abs(dis.astype(np.float64) ** 2 - dis_ref.astype(np.float64) ** 2).max() <= 8e-6
    same = idx == idx_ref
    assert same.mean() >= 0.999
    # where the argmin differs the two candidates are within rounding of each other in the reference's own matrix
    bad = np.where(~same)[0]
    assert np.all(np.abs(D[bad, idx[bad]].astype(np.float64) ** 2 - D[bad, idx_ref[bad]].astype(np.float64) ** 2) <= 8e-6)
    exact = np.where(dis_ref < 3e-3)[0]                                # rows with an exact copy in ft
    assert len(exact) >= m - 2 and np.array_equal(idx[exact], idx_ref[exact])
    # mutual variant through the public helper
    corr_m_ref, *_ = FO.correspondences(fs, ft, use_mutual=True)
    corr_m, _ = ops.feature_correspondences(torch.from_numpy(fs).to(DEV), torch.from_numpy(ft).to(DEV), use_mutual=True)
    got = set(map(tuple, corr_m.cpu().numpy().tolist())); want = set(map(tuple, corr_m_ref.tolist()))
    assert len(got ^ want) <= max(2, int(0.002 * len(want)))


@pytest.mark.parametrize("n,k,extent", [(2048, 16, 3.0), (8192, 16, 100.0), (20000, 32, 3.0)])
def test_knn_kernel_against_independent_kdtree(n, k, extent):
    """The GPU k-NN (cell-grid search) against scipy cKDTree on tie-free clouds: identical ids in identical order on every
    row whose top-(k+1) distances are separated beyond fp32 resolution -- an oracle-independent check of a1."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(77 + n)
    x = (rng.random((n, 3)) * extent).astype(np.float32)
    d, idx = cKDTree(x.astype(np.float64)).query(x.astype(np.float64), k=k + 1)
    d2 = d ** 2
    ok = (np.diff(d2, axis=1) > 1e-5 * np.maximum(d2[:, 1:], 1e-12)).all(axis=1)
    assert ok.mean() > 0.9
    got = ops.knn_build(torch.from_numpy(x)[None].to(DEV), k)[0].cpu().numpy()
    assert np.array_equal(got[ok], idx[ok, :k].astype(np.int32))


def test_module_api_rejects_bad_edge_indices_and_loop_false(model):
    """Advice items: out-of-range edge ids raise (the reference's index ops would) instead of being clamped silently;
    knn_graph(loop=False) = torch_cluster's default (the centre itself is not a neighbour); unsorted_segment_sum refuses
    inputs that require grad."""
    n = 300
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, 3, generator=g).to(DEV)
    h = torch.randn(n, 32, generator=g).to(DEV)
    e = P.knn_graph(x, 16, loop=True)
    bad = e.clone(); bad[0, 5] = n + 7
    with pytest.raises(IndexError):
        model.egnn(h, x, [bad[0], bad[1]], torch.ones(bad.shape[1], 1, device=DEV))
    with pytest.raises(IndexError):
        P.unsorted_segment_sum(torch.ones(4, 2, device=DEV), torch.tensor([0, 1, 9, 2], device=DEV), 3)
    e0 = P.knn_graph(x, 8, loop=False)
    assert e0.shape == (2, n * 8) and int((e0[0] == e0[1]).sum()) == 0
    ref = knn_oracle.knn(x.cpu().numpy(), 9)
    want = np.stack([[j for j in ref[i] if j != i][:8] for i in range(n)])
    assert np.array_equal(e0[0].view(n, 8).cpu().numpy(), want)
    with pytest.raises(NotImplementedError):
        P.unsorted_segment_sum(torch.ones(4, 2, device=DEV, requires_grad=True), torch.tensor([0, 1, 1, 2], device=DEV), 3)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_knn_grid_randomized_shapes_equal_the_scan(seed):
    """The cell-grid query (shell pruning by the k-th best distance, x-range trimming, 64-bit (d2, index) keys) against the
    brute-force scan kernel and -- on a subset -- the C oracle: clustered / planar / collinear / duplicate-heavy /
    lattice (exact ties everywhere) / anisotropic clouds, scales 1e-3 .. 1e4 with offsets, n = 3 .. 9000, k = 1 .. 32.
    Pruning may only skip candidates that lose anyway: ids must stay bit-identical."""
    rng = np.random.default_rng(100 + seed)
    for it in range(24):
        n = int(rng.choice([3, 17, 100, 333, 1024, 2048, 4096, 9000]))
        C = int(rng.choice([1, 2, 5]))
        k = int(rng.choice([1, 4, 8, 16, 16, 24, 32]))
        kind = rng.choice(["uniform", "cluster", "plane", "line", "dups", "lattice", "shell", "aniso"])
        pts = rng.random((C, n, 3))
        if kind == "cluster":
            centres = rng.random((C, 6, 3))
            pts = centres[np.arange(C)[:, None], rng.integers(0, 6, (C, n))] + 0.01 * rng.standard_normal((C, n, 3))
        elif kind == "plane":
            pts[..., 2] = 0.5
        elif kind == "line":
            pts[..., 1:] = 0.25
        elif kind == "dups":
            src = rng.integers(0, max(1, n // 7), (C, n))
            pts = np.take_along_axis(pts, src[..., None].repeat(3, -1), axis=1)
        elif kind == "lattice":
            m = int(np.ceil(n ** (1 / 3)))
            g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
            pts = np.broadcast_to(g, (C, n, 3)).astype(np.float64) / m
        elif kind == "shell":
            v = rng.standard_normal((C, n, 3)); pts = v / np.linalg.norm(v, axis=-1, keepdims=True)
        elif kind == "aniso":
            pts = pts * np.array([100.0, 100.0, 6.0])
        scale = float(rng.choice([1e-3, 1.0, 3.0, 100.0, 1e4]))
        off = float(rng.choice([0.0, -5.0, 1000.0])) * scale
        xh = (pts * scale + off).astype(np.float32)
        x = torch.from_numpy(xh).to(DEV).contiguous()
        a = ops.knn_build(x, k)
        b = ops.knn_build(x, k, brute_force=True)
        assert torch.equal(a, b), (kind, n, C, k, scale, off)
        if n <= 1024:
            assert np.array_equal(a.cpu().numpy(), knn_oracle.knn(xh, k)), (kind, n, C, k, scale, off)


@pytest.mark.parametrize("C,N,k,kind", [(64, 2048, 16, "uniform"), (16, 4096, 16, "uniform"), (16, 4096, 16, "identical"),
                                        (4, 9000, 16, "dups"), (4, 9000, 8, "uniform"), (2, 16384, 16, "uniform"),
                                        (16, 16384, 16, "dups"), (2, 20000, 16, "uniform"), (16, 4096, 32, "identical")])
def test_csr_from_nbr_every_build_path_exact(C, N, k, kind):
    """The reverse-k-NN lists from a neighbour table on every build path -- one CTA per cloud (<= 2048-point clouds), the
    split build (`split` CTAs per cloud owning 2048 rows each; all-identical / duplicate-heavy clouds send every edge to a
    few low-index rows and force its multi-round path), the generic kernels (few or very large clouds) -- against a stable
    sort of the edge list by row: ptr / row / col / eid bit-exact."""
    g = torch.Generator().manual_seed(N + k)
    x = torch.rand(C, N, 3, generator=g)
    if kind == "identical":
        x[:] = 0.25
    elif kind == "dups":
        src = torch.randint(0, max(1, N // 50), (C, N), generator=g)
        x = torch.gather(x, 1, src[..., None].expand(-1, -1, 3))
    nbr = ops.knn_build(x.to(DEV).contiguous(), k)
    gr = ops.csr_from_nbr(nbr)
    row = nbr.cpu().long().reshape(C, N * k)
    col = torch.arange(N).repeat_interleave(k)[None].expand(C, -1)
    ptr, r, c, e = _csr_reference(row, col, C, N)
    assert torch.equal(gr.ptr.cpu().long(), ptr) and torch.equal(gr.row.cpu().long(), r)
    assert torch.equal(gr.col.cpu().long(), c) and torch.equal(gr.eid.cpu().long(), e)
