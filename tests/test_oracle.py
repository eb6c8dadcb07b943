"""CPU: the oracle (oracle/) pinned against the golden vectors produced by the reference's own
classes (tests/golden/make_golden.py), and against the live reference when /root/reference exists."""
import os

import numpy as np
import pytest
import torch

from oracle import egnn_oracle as O
from oracle import knn_oracle, ref_loader

CASES = ["small_b2_n256", "dup_b2_n512", "full_b1_n2048", "kitti_b1_n1024", "noenc_b1_n512"]


def load_case(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False, map_location="cpu")
    ck = torch.load(os.path.join(golden_dir, g["meta"]["checkpoint"].split("/")[-1]), map_location="cpu", weights_only=True)
    return g, ck["cross_attention_state_dict"]


def edges_of(nbr):
    return torch.stack([torch.stack(O.edges_from_nbr(n)) for n in nbr])


@pytest.mark.parametrize("name", CASES)
def test_oracle_eval_matches_reference_golden(golden_dir, name):
    g, sd = load_case(golden_dir, name)
    inp = g["inputs"]
    out = O.forward_eval(sd, inp["src_feat"], inp["src_pts"], edges_of(g["nbr_src"]), inp["tgt_feat"], inp["tgt_pts"],
                         edges_of(g["nbr_tgt"]), inp["labels"], inp["gt_pose"])
    ref = g["eval_f32"]
    # same torch ops in the same order as the reference -> essentially bit-identical
    for i, key in ((4, "h_src"), (5, "x_src"), (6, "h_tgt"), (7, "x_tgt"), (0, "R"), (1, "t")):
        assert torch.allclose(out[i], ref[key], rtol=0, atol=1e-6 * float(ref[key].abs().max())), key
    assert out[2] is None
    # the reference was called per pair (B=1); its loss is the per-pair loss
    per_pair = torch.stack([O.forward_eval(sd, inp["src_feat"][b:b + 1], inp["src_pts"][b:b + 1], edges_of(g["nbr_src"][b:b + 1]),
                                           inp["tgt_feat"][b:b + 1], inp["tgt_pts"][b:b + 1], edges_of(g["nbr_tgt"][b:b + 1]),
                                           inp["labels"][b:b + 1], inp["gt_pose"][b:b + 1])[3] for b in range(inp["labels"].shape[0])])
    assert torch.allclose(per_pair, ref["equi_loss"], rtol=1e-6)


@pytest.mark.parametrize("name", ["small_b2_n256", "dup_b2_n512", "noenc_b1_n512"])
def test_oracle_train_matches_reference_golden(golden_dir, name):
    g, sd = load_case(golden_dir, name)
    inp = g["inputs"]
    out = O.forward_train(sd, inp["src_feat"], inp["src_pts"], edges_of(g["nbr_src"]), inp["tgt_feat"], inp["tgt_pts"],
                          edges_of(g["nbr_tgt"]), inp["labels"], inp["gt_pose"])
    ref = g["train_f32"]
    assert torch.allclose(out[4], ref["h_src"], rtol=0, atol=1e-6 * float(ref["h_src"].abs().max()))
    assert torch.allclose(out[2], ref["slot2"].reshape(()), rtol=1e-6)
    assert torch.allclose(out[3], ref["equi_loss"].reshape(()), rtol=1e-6)
    assert torch.allclose(out[0], ref["R"], atol=1e-5) and torch.allclose(out[1], ref["t"], atol=1e-5)


def test_oracle_per_layer_states(golden_dir):
    g, sd = load_case(golden_dir, "small_b2_n256")
    esd = {k[5:]: v for k, v in sd.items() if k.startswith("egnn.")}
    inp = g["inputs"]
    row, col = O.edges_from_nbr(g["nbr_src"][0])
    ea = torch.ones(row.shape[0], 1)
    h, x, layers = O.egnn_forward(esd, inp["src_feat"][0], inp["src_pts"][0], row, col, ea, return_layers=True)
    ref = g["eval_f32"]["layers_src0"]
    for i, (hl, xl) in enumerate(layers):
        assert torch.allclose(hl, ref[i + 1][0], atol=1e-6 * float(ref[i + 1][0].abs().max()))
        assert torch.allclose(xl, ref[i + 1][1], atol=1e-5)
    m, _ = O.egcl_edge_messages(esd, "gcl_0.", torch.nn.functional.linear(inp["src_feat"][0], esd["embedding_in.weight"], esd["embedding_in.bias"]),
                                inp["src_pts"][0], row, col, ea)
    assert torch.allclose(m[: 64 * 16], g["eval_f32"]["layer0_messages"], atol=1e-5)


def test_knn_c_oracle_matches_numpy_spec():
    rng = np.random.default_rng(0)
    for n, dup in ((64, 0), (300, 100), (1000, 0)):
        x = (rng.random((n, 3)) * 3).astype(np.float32)
        if dup:
            x[-dup:] = x[:dup]
        assert np.array_equal(knn_oracle.knn(x, 16), knn_oracle.knn_numpy(x, 16))
    x = (rng.random((10, 3))).astype(np.float32)          # n < k: unfilled slots are -1
    out = knn_oracle.knn(x, 16)
    assert np.array_equal(out, knn_oracle.knn_numpy(x, 16)) and (out[:, 10:] == -1).all()
    # nearest-first, self first on distinct points, ties -> lower index on exact duplicates
    x = np.zeros((40, 3), np.float32)
    assert np.array_equal(knn_oracle.knn(x, 16), np.tile(np.arange(16, dtype=np.int32), (40, 1)))


def test_knn_golden_fixture_is_oracle_output(golden_dir):
    g, _ = load_case(golden_dir, "dup_b2_n512")
    assert np.array_equal(knn_oracle.knn(g["inputs"]["src_pts"].numpy(), 16), g["nbr_src"].numpy())


def test_metrics_known_answers(golden_dir):
    """The oracle's numpy restatement of tools/evaluation_metrics.py against the known answers that file itself produced
    (the product's metrics are the device kernel: checked against the same fixture in tests/test_gpu_parity.py)."""
    cases = torch.load(os.path.join(golden_dir, "metrics_kat.pt"), weights_only=False)
    for c in cases:
        re, te = O.calculate_pose_error(c["gt"], c["pred"])
        rec, prec = O.registration_recall(c["gt"], c["pred"], c["src"], c["tgt"])
        assert np.isclose(re, c["re"], atol=1e-9) and np.isclose(te, c["te"], atol=1e-9)
        assert np.isclose(rec, c["recall"]) and np.isclose(prec, c["precision"])


def test_kabsch_oracle_recovers_known_pose_and_reflection_fix():
    rng = np.random.default_rng(3)
    from se3_equi_graph_registration_b200.synthetic import random_rotation
    R = torch.tensor(random_rotation(rng), dtype=torch.float64)
    t = torch.tensor(rng.random(3), dtype=torch.float64)
    p = torch.tensor(rng.random((50, 3)), dtype=torch.float64)
    q = p @ R.T + t
    w = torch.full((50,), 1.0 / 50, dtype=torch.float64)
    Rk, tk, _ = O.kabsch(p, q, w)
    assert torch.allclose(Rk, R, atol=1e-5) and torch.allclose(tk, t, atol=1e-5)
    # planar + mirrored target: det fix must still return a proper rotation
    p[:, 2] = 0
    q = p.clone(); q[:, 0] = -q[:, 0]
    Rk, _, _ = O.kabsch(p, q, w)
    assert abs(float(torch.det(Rk)) - 1.0) < 1e-6
    Re, te, _ = O.kabsch(p[:0], q[:0], w[:0])
    assert torch.equal(Re, torch.eye(3, dtype=torch.float64)) and float(te.abs().sum()) == 0.0


@pytest.mark.skipif(not ref_loader.reference_available(), reason="live reference only exists in the build container")
def test_oracle_against_live_reference():
    """Extra pin in the build container: run the reference classes right now on fresh inputs."""
    from se3_equi_graph_registration_b200 import synthetic
    data = synthetic.make_batch(99, 1, n=200)
    nbr = torch.from_numpy(knn_oracle.knn(data["src_pts"].numpy(), 16))
    nbt = torch.from_numpy(knn_oracle.knn(data["tgt_pts"].numpy(), 16))
    es, et = edges_of(nbr), edges_of(nbt)
    ns, egnn, head = ref_loader.build_reference_model("eval")
    ea = torch.ones(1, es.shape[-1], 1)
    with torch.no_grad():
        ref = ref_loader.run_quiet(head, data["src_feat"], data["src_pts"], es, ea, data["tgt_feat"], data["tgt_pts"], et, ea,
                                   data["corr"], data["labels"], data["gt_pose"])
    out = O.forward_eval(head.state_dict(), data["src_feat"], data["src_pts"], es, data["tgt_feat"], data["tgt_pts"], et,
                         data["labels"], data["gt_pose"])
    for i in (0, 1, 4, 5, 6, 7):
        assert torch.allclose(out[i], ref[i], rtol=0, atol=1e-6 * float(ref[i].abs().max()))


def test_feature_match_oracle_known_answers():
    """oracle/feature_match_oracle.py restates data_preprocess/3DMatch_Feature.py:158-166 (three numpy lines inside a
    file-processing loop that cannot be imported); pin it on hand-computable cases."""
    from oracle import feature_match_oracle as FO
    e = np.eye(4, 32, dtype=np.float32)                          # 4 orthonormal descriptors
    src = e[[0, 1, 2, 3]]
    tgt = np.stack([e[2], e[0], e[0], (e[1] + e[3]) / np.sqrt(2)]).astype(np.float32)
    corr, sidx, sdis, D = FO.correspondences(src, tgt, use_mutual=False)
    assert sidx.tolist() == [1, 3, 0, 3]                         # duplicate target e0 at 1 and 2: the first wins (np.argmin)
    assert np.allclose(sdis[[0, 2]], np.sqrt(1e-6), rtol=1e-3)   # exact matches: sqrt(2 - 2 + 1e-6)
    assert np.allclose(sdis[[1, 3]], np.sqrt(2 - np.sqrt(2) + 1e-6), rtol=1e-5)
    assert corr.tolist() == [[0, 1], [1, 3], [2, 0], [3, 3]]
    corr_m, *_ = FO.correspondences(src, tgt, use_mutual=True)
    # target 3 is equally close to sources 1 and 3 -> argmin over axis 0 picks source 1: (3,3) is not mutual
    assert corr_m.tolist() == [[0, 1], [1, 3], [2, 0]]
    assert D.dtype == np.float32


# ---------------------------------------------------------------------------------------------
# training step: autograd of the oracle restatement == the reference's own backward (golden gradients)
# ---------------------------------------------------------------------------------------------
def _train_loss(out, scenario, gt_pose):
    if scenario == "shipped":
        return out[2] + out[3]
    rot, trans = O.pose_loss(out[0], out[1], gt_pose)
    return out[2] + rot.mean() + trans.mean()                      # 3dm:1107-1118


GRAD_CASES = [("small_b2_n256", "grads_b2_n256.pt"), ("dup_b2_n512", "grads_dup_b2_n512.pt")]


@pytest.mark.parametrize("case,gfile", GRAD_CASES)
@pytest.mark.parametrize("scenario", ["shipped", "tempered"])
def test_oracle_gradients_match_reference_golden(golden_dir, scenario, case, gfile):
    g, sd = load_case(golden_dir, case)
    gg = torch.load(os.path.join(golden_dir, gfile), weights_only=False, map_location="cpu")
    ref = gg[scenario + "_f32"]
    sd = {k: v.clone() for k, v in sd.items()}
    if scenario == "tempered":
        sd["egnn.embedding_out.weight"] *= gg["meta"]["temper"]
        sd["egnn.embedding_out.bias"] *= gg["meta"]["temper"]
    sd = {k: (v.requires_grad_(True) if v.is_floating_point() else v) for k, v in sd.items()}
    inp = g["inputs"]
    out = O.forward_train(sd, inp["src_feat"], inp["src_pts"], edges_of(g["nbr_src"]), inp["tgt_feat"], inp["tgt_pts"],
                          edges_of(g["nbr_tgt"]), inp["labels"], inp["gt_pose"])
    loss = _train_loss(out, scenario, inp["gt_pose"])
    assert abs(float(loss.detach()) - ref["loss"]) <= 1e-5 * abs(ref["loss"])
    loss.backward()
    n = 0
    for k, gref in ref["grads"].items():
        if gref is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
            continue
        err = float((sd[k].grad - gref).abs().max())
        assert err <= 2e-3 * float(gref.abs().max()) + 1e-7, (k, err, float(gref.abs().max()))
        n += 1
    assert n == 85                                                  # 85 of 97 tensors get gradients
    if scenario + "_f64" in gg:     # how far the reference's own fp32 gradients are from its fp64 run (calibrates the GPU bar)
        r64 = gg[scenario + "_f64"]["grads"]
        worst = max(float((ref["grads"][k] - r64[k]).abs().max() / r64[k].abs().max()) for k in r64 if r64[k] is not None)
        assert worst < 1e-3


def _tie_free_cloud(rng, n, extent):
    """Points whose pairwise squared distances (fp32, the spec's fma order) are all distinct per query row."""
    return (rng.random((n, 3)) * extent).astype(np.float32)


@pytest.mark.parametrize("n,k,extent", [(600, 16, 3.0), (2048, 16, 3.0), (1500, 32, 100.0)])
def test_knn_oracle_against_independent_kdtree(n, k, extent):
    """INDEPENDENT pin of oracle/knn_oracle.c (torch_cluster is absent): on clouds without distance ties
    any correct k-NN returns the same ids in the same order, so scipy.spatial.cKDTree (different algorithm, different
    author, fp64) must agree exactly.  Rows where cKDTree's own top-(k+1) distances are closer than fp32 resolution are
    excluded (there the fp32 spec may legitimately order differently)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(123 + n)
    x = _tie_free_cloud(rng, n, extent)
    d, idx = cKDTree(x.astype(np.float64)).query(x.astype(np.float64), k=k + 1)
    d2 = d ** 2
    gap = np.diff(d2, axis=1)
    ok = (gap > 1e-5 * np.maximum(d2[:, 1:], 1e-12)).all(axis=1)           # well-separated ranks only
    assert ok.mean() > 0.9
    got = knn_oracle.knn(x, k)
    assert np.array_equal(got[ok], idx[ok, :k].astype(np.int32))


def test_knn_oracle_duplicate_heavy_properties():
    """On duplicate-heavy clouds (many exact ties) the spec's order is (d2, index): ids sorted by that key, every
    returned distance <= every omitted distance, ties broken towards the lower index."""
    rng = np.random.default_rng(5)
    n, k = 700, 16
    x = (rng.random((n, 3)) * 3).astype(np.float32)
    x[200:500] = x[rng.integers(0, 200, 300)]                               # > 40 % repeated points
    got = knn_oracle.knn(x, k)
    d2 = ((x[:, None, :].astype(np.float64) - x[None].astype(np.float64)) ** 2).sum(-1)
    for i in range(0, n, 7):
        row = got[i]
        key = [(np.float32(d2[i, j]), j) for j in row]
        assert key == sorted(key), i
        kth = key[-1]
        others = np.setdiff1d(np.arange(n), row)
        assert all((np.float32(d2[i, j]), j) > kth for j in others), i


@pytest.mark.parametrize("heads", [1, 2, 8])
def test_oracle_other_head_counts_match_reference_golden(golden_dir, heads):
    """E_GCL(num_heads=h) for h in {1, 2, 8} (3dm:186-207, heads concatenated at :246): the oracle against outputs AND
    gradients of the reference's own EGNN class built with that head count (tests/golden/make_golden_heads.py;
    randomly initialised -- the shipped checkpoints have 4 heads)."""
    g = torch.load(os.path.join(golden_dir, "heads_%d.pt" % heads), weights_only=False, map_location="cpu")
    assert O.num_heads_of(g["state_dict"]) == heads
    sd = {k: v.double().requires_grad_(True) for k, v in g["state_dict"].items()}
    h, x = g["h"].double().requires_grad_(True), g["x"].double().requires_grad_(True)
    ho, xo = O.egnn_forward(sd, h, x, g["row"], g["col"], g["edge_attr"].double())
    assert float((ho - g["h_out"].double()).abs().max()) <= 1e-5 * float(g["h_out"].abs().max())
    assert float((xo - g["x_out"].double()).abs().max()) <= 1e-5
    ((ho * g["dh"].double()).sum() + (xo * g["dx"].double()).sum()).backward()
    rel = lambda a, b: float((a - b.double()).abs().max() / (b.double().abs().max() + 1e-30))
    assert rel(h.grad, g["grad_h"]) < 1e-5 and rel(x.grad, g["grad_x"]) < 1e-5
    for k, gr in g["grads"].items():
        assert rel(sd[k].grad, gr) < 1e-5, k
