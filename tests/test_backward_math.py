"""CPU: the backward arithmetic the CUDA kernels use (csrc/egnn_backward_math.cuh), compiled with g++ through
tests/bwd_host_harness.cpp, against torch autograd of the oracle's E_GCL layer (= the reference's autograd graph,
3dm:1125 `loss.backward()`), and the pack <-> parameter gradient mapping."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import egnn_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "build", "libbwd_host_harness.so")


@pytest.fixture(scope="module")
def harness():
    src = os.path.join(ROOT, "tests", "bwd_host_harness.cpp")
    hdr = os.path.join(ROOT, "se3-equi-graph-registration_b200", "csrc", "egnn_backward_math.cuh")
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", src, "-o", SO])
    lib = ctypes.CDLL(SO)
    lib.egcl_backward_host.restype = None
    return lib


def _layer_module(sd, i):
    from se3_equi_graph_registration_b200.modules import E_GCL
    gcl = E_GCL(32, 32, 32, edges_in_d=1, num_heads=4, device="cpu")
    pre = f"egnn.gcl_{i}."
    gcl.load_state_dict({k[len(pre):]: v for k, v in sd.items() if k.startswith(pre)})
    return gcl


def _graph(n, k, seed, dup=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, generator=g) * 1.5
    if dup:
        x[n // 2:] = x[:n - n // 2]                       # exact duplicates: zero-length edges / identity frames
    d2 = ((x[:, None] - x[None]) ** 2).sum(-1)
    nbr = d2.argsort(dim=1, stable=True)[:, :k]
    row, col = O.edges_from_nbr(nbr)
    # plus some arbitrary extra edges (user graphs are not restricted to k-NN)
    extra = torch.randint(0, n, (2, 3 * n), generator=g)
    return x, torch.cat([row, extra[0]]), torch.cat([col, extra[1]])


@pytest.mark.parametrize("layer,dup,per_edge_attr", [(0, False, False), (1, True, False), (2, False, True)])
def test_layer_backward_matches_autograd(golden_dir, harness, layer, dup, per_edge_attr):
    from se3_equi_graph_registration_b200 import packing
    sd = torch.load(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), map_location="cpu", weights_only=True)["cross_attention_state_dict"]
    gcl = _layer_module(sd, layer)
    n, k = 96, 8
    x, row, col = _graph(n, k, 5 + layer, dup)
    E = row.numel()
    g = torch.Generator().manual_seed(77)
    h = torch.randn(n, 32, generator=g) * 0.5
    ea = (torch.rand(E, 1, generator=g) + 0.5) if per_edge_attr else torch.ones(E, 1)
    dh_out = torch.randn(n, 32, generator=g)
    dx_out = torch.randn(n, 3, generator=g)

    # autograd of the oracle restatement (fp64 for a clean comparison)
    p = f"gcl_{layer}."
    lsd = {p + k_: v.detach().double().requires_grad_(True) for k_, v in gcl.state_dict().items()}
    h64 = h.double().requires_grad_(True)
    x64 = x.double().requires_grad_(True)
    h2, x2, _ = O.egcl_forward(lsd, p, h64, x64, row, col, ea.double())
    loss = (h2 * dh_out.double()).sum() + (x2 * dx_out.double()).sum()
    loss.backward()

    pack = packing.pack_layer(gcl).contiguous()
    gpack = np.zeros(packing.LAYER_PACK, dtype=np.float32)
    dh_in = np.zeros((n, 32), dtype=np.float32)
    dx_in = np.zeros((n, 3), dtype=np.float32)
    fp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    tp = lambda t: ctypes.c_void_p(t.data_ptr())
    row32, col32 = row.to(torch.int32).contiguous(), col.to(torch.int32).contiguous()
    eac = ea.reshape(-1).contiguous()
    harness.egcl_backward_host(tp(pack), ctypes.c_int(n), ctypes.c_int(E), tp(row32), tp(col32),
                               tp(eac) if per_edge_attr else None, ctypes.c_float(1.0), tp(h.contiguous()), tp(x.contiguous()),
                               tp(dh_out.contiguous()), tp(dx_out.contiguous()), fp(dh_in), fp(dx_in), fp(gpack))

    def close(a, b, name, tol=2e-4):
        a = torch.as_tensor(a).double()
        err = float((a - b).abs().max()); scale = float(b.abs().max()) + 1e-30
        assert err <= tol * scale, f"{name}: err {err:.3e} vs scale {scale:.3e}"

    close(dh_in, h64.grad, "dh")
    close(dx_in, x64.grad, "dx")
    # the host harness (CUDA-core arithmetic, 4 heads) leaves d W2 in the per-head [head][in][out] region; the kernels of
    # the product write the full [out][in] matrix, which is what unpack_layer_grad reads
    gp = torch.from_numpy(gpack)
    w2p = gp[packing.OFF["W2P"]:packing.OFF["W2P"] + 256].reshape(4, 8, 8)
    w2f = torch.zeros(32, 32)
    for hd in range(4):
        w2f[8 * hd:8 * hd + 8, 8 * hd:8 * hd + 8] = w2p[hd].t()
    gp[packing.OFF["W2F"]:packing.OFF["W2F"] + 1024] = w2f.reshape(-1)
    grads = packing.unpack_layer_grad(gp, gcl)
    for (name, _), gk in zip(gcl.named_parameters(), grads):
        close(gk, lsd[p + name].grad, name)


def test_pack_unpack_is_a_permutation(golden_dir):
    from se3_equi_graph_registration_b200 import packing
    sd = torch.load(os.path.join(golden_dir, "checkpoint-3dmatch.pth"), map_location="cpu", weights_only=True)["cross_attention_state_dict"]
    gcl = _layer_module(sd, 1)
    back = packing.unpack_layer_grad(packing.pack_layer(gcl), gcl)
    for prm, b in zip(gcl.parameters(), back):
        assert torch.equal(prm.detach(), b)
    lin = torch.nn.Linear(32, 32)
    w, b = packing.unpack_linear32_grad(packing.pack_linear32(lin), lin)
    assert torch.equal(w, lin.weight.detach()) and torch.equal(b, lin.bias.detach())
