"""CPU: host-side mirror (module / parameter layout, checkpoint loading, packs), the C-ABI
library's exported symbols, loud failure without a GPU, and the pair-sharding logic (gloo, 2 ranks)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import _lib, packing
from oracle import egnn_oracle as O
import pack_emulator as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "egspr_b200.h")).read()
    declared = set(re.findall(r"\b(egspr_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert getattr(handle, name) is not None
    lib = _lib.lib()
    assert lib.egspr_version() >= 100
    assert lib.egspr_error_string(-2).decode().startswith("unsupported")
    assert lib.egspr_csr_workspace_bytes(2048, 32768) >= 4 * (2048 + 32768)
    # pack sizes in the header agree with the host packer
    for macro, val in (("EGSPR_LAYER_PACK_FLOATS", packing.LAYER_PACK), ("EGSPR_EMBED_PACK_FLOATS", packing.EMBED_PACK),
                       ("EGSPR_HEAD_PACK_FLOATS", packing.HEAD_PACK)):
        assert int(re.search(macro + r"\s+(\d+)", hdr).group(1)) == val


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out


@pytest.mark.parametrize("ck", ["checkpoint-3dmatch.pth", "checkpoint-3dmatch-no-encoder.pth"])
def test_checkpoint_loads_unchanged(golden_dir, ck):
    model = P.build_model(None, device="cpu")
    ckpt, epoch = P.load_checkpoint(os.path.join(golden_dir, ck), None, model.egnn, model, device="cpu")   # strict
    assert epoch == ckpt["epoch"] and epoch in (21, 58)
    assert set(model.egnn.state_dict()) == set(ckpt["egnn_state_dict"])
    assert set(model.state_dict()) == set(ckpt["cross_attention_state_dict"])
    for k, v in ckpt["cross_attention_state_dict"].items():
        assert tuple(model.state_dict()[k].shape) == tuple(v.shape), k
    assert sum(p.numel() for p in model.parameters()) == 45742
    # Adam state of the checkpoint maps onto the parameter order (97 params in one group)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4)
    opt.load_state_dict(ckpt["optimizer_state_dict"])


def test_checkpoint_round_trip(tmp_path, golden_dir):
    model = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device="cpu")
    path = P.save_checkpoint(3, None, model.egnn, model, torch.optim.Adam(model.parameters()), save_dir=str(tmp_path), is_best=True)   # 3dm:1310 signature
    assert os.path.exists(os.path.join(str(tmp_path), "best_checkpoint"))
    m2 = P.build_model(None, device="cpu")
    _, ep = P.load_checkpoint(path, None, m2.egnn, m2, device="cpu")
    assert ep == 3
    for (k, a), (_, b) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    with pytest.raises(FileNotFoundError):
        P.load_checkpoint(str(tmp_path / "nope.pth"), None, m2.egnn, m2, device="cpu")


def test_reference_default_num_heads_would_not_load(golden_dir):
    """with the reference's num_heads=1 default the shipped checkpoint does not load."""
    egnn = P.EGNN(32, 32, 32, in_edge_nf=1, device="cpu", n_layers=3, num_heads=1)
    ck = torch.load(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), map_location="cpu", weights_only=True)
    with pytest.raises(RuntimeError):
        egnn.load_state_dict(ck["egnn_state_dict"])


@pytest.mark.parametrize("name", ["small_b2_n256", "kitti_b1_n1024"])
def test_weight_packs_reproduce_the_reference(golden_dir, name):
    """The packs the kernels read, pushed through a plain-torch emulation of the kernel algebra,
    give the reference's outputs -> layout / transposes / P-Q split / bias folding are right."""
    g = torch.load(os.path.join(golden_dir, name + ".pt"), weights_only=False, map_location="cpu")
    model = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device="cpu")
    layers, pin, pout = model.egnn.packs()
    inp = g["inputs"]
    row, col = O.edges_from_nbr(g["nbr_src"][0])
    h, x = E.egnn(layers, pin, pout, inp["src_feat"][0], inp["src_pts"][0], row, col)
    ref = g["eval_f32"]
    assert float((h - ref["h_src"][0]).abs().max()) <= 2e-6 * float(ref["h_src"][0].abs().max())
    assert float((x - ref["x_src"][0]).abs().max()) <= 1e-5 * max(1.0, float(ref["x_src"][0].abs().max()))
    z = torch.randn(5, 64)
    assert torch.allclose(E.head_mlp(model._pack_head.get(), z), model.mlp(z).squeeze(-1), atol=1e-5)


def test_pack_cache_tracks_parameter_updates():
    model = P.build_model(None, device="cpu")
    gcl = model.egnn.gcl_0
    a = gcl.layer_pack()
    assert gcl.layer_pack() is a
    with torch.no_grad():
        gcl.layer_norm.weight.add_(1.0)
    b = gcl.layer_pack()
    assert b is not a and not torch.equal(a, b)
    assert torch.allclose(b[packing.OFF["LNG"]:packing.OFF["LNG"] + 32], gcl.layer_norm.weight)


def test_unsupported_configurations_fail_loudly():
    with pytest.raises(NotImplementedError):
        P.EGNN(32, 32, 32, in_edge_nf=1, device="cpu", n_layers=1, tanh=True).packs()
    with pytest.raises(NotImplementedError):
        P.EGNN(32, 32, 32, in_edge_nf=1, device="cpu", n_layers=1, num_heads=3).packs()        # 32 % 3 != 0
    with pytest.raises(NotImplementedError):
        P.EGNN(33, 64, 33, in_edge_nf=1, device="cpu", n_layers=1).packs()
    # head counts other than 4 run on the tensor-core kernels only: the CUDA-core impls refuse, and their region is poisoned
    e2 = P.EGNN(32, 32, 32, in_edge_nf=1, device="cpu", n_layers=1, num_heads=2)
    with pytest.raises(NotImplementedError):
        e2.check_impl(1)
    e2.check_impl(0); e2.check_impl(5)
    pk = e2.packs()[0][0]
    assert torch.isnan(pk[packing.OFF["W2P"]:packing.OFF["W2P"] + 256]).all() and not torch.isnan(pk[packing.OFF["W2F"]:]).any()


@pytest.mark.parametrize("heads", [1, 2, 4, 8, 16])
def test_layer_pack_any_head_count(heads):
    """pack_layer / unpack_layer_grad for every head count dividing 32 (E_GCL(num_heads=...), 3dm:186-207): the second
    edge Linear travels as one block-diagonal [out][in] matrix; the cached (index-gather) rebuild equals the direct one;
    the gradient unpack is the inverse permutation on the heads' blocks."""
    torch.manual_seed(heads)
    egnn = P.EGNN(32, 32, 32, in_edge_nf=1, device="cpu", n_layers=1, num_heads=heads)
    gcl = egnn.gcl_0
    d = 32 // heads
    pk = packing.pack_layer(gcl)
    w2f = pk[packing.OFF["W2F"]:packing.OFF["W2F"] + 1024].reshape(32, 32)
    for g, m in enumerate(gcl.edge_mlps):
        assert torch.equal(w2f[d * g:d * g + d, d * g:d * g + d], m[2].weight.detach())
    mask = torch.block_diag(*[torch.ones(d, d)] * heads).bool()
    assert float(w2f[~mask].abs().max() if (~mask).any() else 0.0) == 0.0
    cached = gcl.layer_pack()                       # first get(): builds the index map, then gathers
    with torch.no_grad():
        gcl.edge_mlps[0][2].weight.mul_(1.5)
    cached = gcl.layer_pack()                       # second get(): pure gather
    direct = packing.pack_layer(gcl)
    assert torch.equal(torch.nan_to_num(cached, nan=-7.0), torch.nan_to_num(direct, nan=-7.0))
    # gradient direction: a pack-shaped gradient holding its own positions
    gp = torch.arange(packing.LAYER_PACK, dtype=torch.float32)
    grads = packing.unpack_layer_grad(gp, gcl)
    for (name, prm), gk in zip(gcl.named_parameters(), grads):
        assert gk.shape == prm.shape, name
    for g, m in enumerate(gcl.edge_mlps):
        gk = grads[[id(q) for q in gcl.parameters()].index(id(m[2].weight))]
        want = gp[packing.OFF["W2F"]:packing.OFF["W2F"] + 1024].reshape(32, 32)[d * g:d * g + d, d * g:d * g + d]
        assert torch.equal(gk, want)


def test_no_cpu_fallback():
    model = P.build_model(None, device="cpu")
    x = torch.rand(32, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        P.knn_graph(x, 16, loop=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.egnn(torch.rand(32, 32), x, [torch.zeros(4, dtype=torch.long)] * 2, torch.ones(4, 1))
    with pytest.raises(RuntimeError):
        P.RegistrationEngine(model, batch=1, n=32, device="cpu")
    # the product never imports the oracle
    pkg = os.path.join(ROOT, "se3-equi-graph-registration_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read(), fn


def test_get_edges_batch_layout():
    gi = torch.tensor([[1, 2, 0], [0, 0, 1]])
    (row, col), ea = P.get_edges_batch(gi, 3, 1)
    assert torch.equal(row, gi[0]) and torch.equal(col, gi[1]) and tuple(ea.shape) == (3, 1) and bool((ea == 1).all())
    (row, col), ea = P.get_edges_batch(gi, 3, 2)
    assert row.tolist() == [1, 2, 0, 4, 5, 3] and col.tolist() == [0, 0, 1, 3, 3, 4] and ea.shape[0] == 6


def test_synthetic_pairs_follow_the_dataset_contract():
    d = P.synthetic.make_pair(0, n=2048, dup_frac=0.3)
    assert d["src_pts"].shape == (2048, 3) and d["src_feat"].shape == (2048, 32) and d["gt_pose"].shape == (4, 4)
    assert np.allclose(np.linalg.norm(d["src_feat"], axis=1), 1, atol=1e-5)
    R, t = d["gt_pose"][:3, :3], d["gt_pose"][:3, 3]
    res = np.linalg.norm(d["src_pts"] @ R.T + t - d["tgt_pts"], axis=1)
    assert (res[d["labels"] > 0] < 0.2).all() and abs(np.linalg.det(R) - 1) < 1e-5
    assert len(np.unique(d["src_pts"], axis=0)) < 2048 * 0.8          # duplicate-heavy variant
    assert np.array_equal(P.synthetic.make_pair(0, n=2048, dup_frac=0.3)["tgt_pts"], d["tgt_pts"])   # seeded


def test_pair_sharding_two_ranks_gloo(tmp_path):
    """bench.py's partition: contiguous slices of the pair batch per rank, no data-path collective;
    throughput is aggregated with a MAX over ranks of the elapsed time."""
    script = tmp_path / "w.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import bench\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "lo, hi = bench.shard_range(10, r, w)\n"
        "t = torch.tensor([float(hi - lo)]); dist.all_reduce(t)\n"
        "m = bench.max_over_ranks(1.0 + r, 'cpu')\n"
        "assert t.item() == 10 and m == 2.0, (t, m)\n"
        f"open(os.path.join({str(tmp_path)!r}, 'r%d.txt' % r), 'w').write('%d %d' % (lo, hi))\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert (tmp_path / "r0.txt").read_text() == "0 5" and (tmp_path / "r1.txt").read_text() == "5 10"


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    """train.allreduce_gradients / packing.FlatState: ONE flat fp32 bucket (BASELINE config 4's only collective);
    after it every rank holds the mean gradient, parameters without a gradient (dead modules) included
    as zeros, so identical optimizer steps keep the replicas identical."""
    script = tmp_path / "g.py"
    script.write_text(
        "import os, sys, torch, torch.distributed as dist\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import se3_equi_graph_registration_b200 as P\n"
        "dist.init_process_group('gloo')\n"
        "r, w = dist.get_rank(), dist.get_world_size()\n"
        "torch.manual_seed(0)\n"
        "egnn = P.EGNN(32, 32, 32, in_edge_nf=1, device='cpu', n_layers=3)\n"
        "model = P.CrossAttentionPoseRegression(egnn, hidden_nf=32, device='cpu')\n"
        "params = list(model.parameters())\n"
        "for i, p in enumerate(params):\n"
        "    if i % 7 != 3: p.grad = torch.full_like(p, float(r + 1) * (i + 1))\n"     # some parameters get no gradient
        "P.train.allreduce_gradients(params)\n"
        "for i, p in enumerate(params):\n"
        "    if i % 7 == 3:\n"
        "        assert p.grad is None, i\n"                      # no gradient on any rank: stays None (Adam skips it, as in 1-rank training)
        "    else:\n"
        "        assert torch.all(p.grad == 1.5 * (i + 1)), (i, p.grad.flatten()[0])\n"
        "# FlatState: the parameters' gradients as views of ONE persistent buffer -> the collective is a single all_reduce of it\n"
        "from se3_equi_graph_registration_b200 import packing\n"
        "sd0 = {k: v.clone() for k, v in model.state_dict().items()}\n"
        "fs = packing.FlatState(model)\n"
        "assert all(torch.equal(v, sd0[k]) for k, v in model.state_dict().items())\n"
        "assert fs.flat_grad.numel() == 25953\n"
        "fs.flat_grad.copy_(torch.arange(fs.n_live, dtype=torch.float32) * (r + 1))\n"
        "dist.all_reduce(fs.flat_grad)\n"
        "live = list(model.egnn.parameters()) + list(model.mlp.parameters())\n"
        "off = 0\n"
        "for p in live:\n"
        "    assert p.grad.data_ptr() == fs.flat_grad[off:].data_ptr() and torch.all(p.grad.flatten() == 3.0 * torch.arange(off, off + p.numel()))\n"
        "    off += p.numel()\n"
        "params = list(model.parameters())\n"
        "opt = torch.optim.Adam(params, lr=1e-3); opt.step()\n"
        "chk = torch.cat([p.detach().flatten() for p in params]).double().sum().reshape(1)\n"
        "both = [torch.zeros(1, dtype=torch.float64) for _ in range(w)]; dist.all_gather(both, chk)\n"
        "assert both[0].item() == both[1].item()\n"
        f"open(os.path.join({str(tmp_path)!r}, 'ok%d' % r), 'w').write('ok')\n")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()


def test_compute_losses_matches_reference():
    """modules.compute_losses == the reference's compute_losses (3dm:799-858) when /root/reference is present,
    else against a direct restatement."""
    from oracle import ref_loader
    g = torch.Generator().manual_seed(0)
    R = torch.linalg.qr(torch.randn(3, 3, 3, generator=g))[0]; t = torch.randn(3, 3, generator=g)
    hs, ht = torch.randn(3, 50, 32, generator=g), torch.randn(3, 50, 32, generator=g)
    xs, xt = torch.randn(3, 50, 3, generator=g), torch.randn(3, 50, 3, generator=g)
    lab = (torch.rand(3, 50, generator=g) < 0.6).float(); lab[2] = 0
    got = P.compute_losses(R, t, hs, xs, ht, xt, lab)
    if ref_loader.reference_available():
        want = ref_loader.load("train")["compute_losses"](R, t, hs, xs, ht, xt, lab)
    else:
        d = (torch.einsum("bij,bnj->bni", R, xs) + t[:, None] - xt).norm(dim=-1) * lab
        want = ((d.sum(1) / lab.sum(1).clamp(min=1)).mean(), (hs[lab == 1] - ht[lab == 1]).norm(dim=-1).mean())
    assert torch.allclose(got[0], want[0]) and torch.allclose(got[1], want[1])


def test_pack_cache_and_grad_unpacker_gather_paths(golden_dir):
    """PackCache rebuilds packs as `cat(params)[index]` and GradUnpacker maps gradient packs back with one gather: both
    must equal the direct re-layout (pack_layer / unpack_layer_grad), also after parameter updates and for a layer
    without an edge_attr column (76-wide edge input)."""
    model = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device="cpu")
    for i in range(3):
        gcl = model.egnn._modules["gcl_%d" % i]
        assert torch.equal(gcl.layer_pack(), packing.pack_layer(gcl))
        with torch.no_grad():
            gcl.coord_mlp[0].weight.add_(0.5); gcl.edge_mlps[2][0].bias.mul_(3.0)
        assert torch.equal(gcl.layer_pack(), packing.pack_layer(gcl))            # second build = the gather path
        gp = torch.randn(packing.LAYER_PACK)
        fast = packing.GradUnpacker(packing.unpack_layer_grad, packing.LAYER_PACK, gcl)(gp)
        for a, b, prm in zip(fast, packing.unpack_layer_grad(gp, gcl), gcl.parameters()):
            assert torch.equal(a, b) and a.shape == prm.shape
    layers, pin, pout = model.egnn.packs()
    assert torch.equal(pin, packing.pack_linear32(model.egnn.embedding_in)) and torch.equal(pout, packing.pack_linear32(model.egnn.embedding_out))
    assert torch.equal(model._pack_head.get(), packing.pack_head(model.mlp))
    g0 = P.E_GCL(32, 32, 32, edges_in_d=0, num_heads=4, device="cpu")
    with torch.no_grad():
        g0.layer_norm.weight.mul_(2.0)
    assert torch.equal(g0.layer_pack(), packing.pack_layer(g0))
    with torch.no_grad():
        g0.layer_norm.bias.add_(1.0)
    assert torch.equal(g0.layer_pack(), packing.pack_layer(g0))
    got = packing.GradUnpacker(packing.unpack_layer_grad, packing.LAYER_PACK, g0)(torch.arange(packing.LAYER_PACK, dtype=torch.float32))
    assert [tuple(t.shape) for t in got] == [tuple(p.shape) for p in g0.parameters()]


def test_flat_state_packs_and_gradient_maps(golden_dir):
    """packing.FlatState (the training step's one-gather weight packs and one-gather gradient unpacking): the packs
    equal the per-module builders, follow in-place parameter updates, the gradient map is the inverse permutation, the
    dead parameters keep grad = None, state_dict is unchanged."""
    model = P.build_model(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), device="cpu", variant="train")
    sd0 = {k: v.clone() for k, v in model.state_dict().items()}
    fs = packing.FlatState(model)
    assert all(torch.equal(v, sd0[k]) for k, v in model.state_dict().items())
    with torch.no_grad():
        model.egnn.gcl_1.coord_mlp[0].weight.add_(0.25); model.mlp[2].bias.sub_(1.0)
    fs.refresh_packs()
    for i in range(3):
        assert torch.equal(fs.layer_packs[i], packing.pack_layer(model.egnn._modules["gcl_%d" % i]))
    assert torch.equal(fs.pack_in, packing.pack_linear32(model.egnn.embedding_in))
    assert torch.equal(fs.pack_out, packing.pack_linear32(model.egnn.embedding_out))
    assert torch.equal(fs.pack_head, packing.pack_head(model.mlp))
    fs.gpack_buf.copy_(torch.randn(fs.gpack_buf.numel()))
    fs.gather_gradients()
    for i in range(3):
        gcl = model.egnn._modules["gcl_%d" % i]
        for a, prm in zip(packing.unpack_layer_grad(fs.layer_gpacks[i], gcl), gcl.parameters()):
            assert torch.equal(a, prm.grad)
    for a, prm in zip(packing.unpack_head_grad(fs.gpack_head, model.mlp), model.mlp.parameters()):
        assert torch.equal(a, prm.grad)
    for a, prm in zip(packing.unpack_linear32_grad(fs.gpack_out, model.egnn.embedding_out), model.egnn.embedding_out.parameters()):
        assert torch.equal(a, prm.grad)
    n_grad = sum(p.grad is not None for p in model.parameters())
    assert n_grad == 85 and len(list(model.parameters())) == 97
