"""Test helper: a plain torch (CPU) emulation of the kernels' algebra driven by the SAME weight
packs the CUDA kernels read (packing.py).  It lets the CPU suite prove the pack layout (offsets,
transposes, P/Q split, bias / edge_attr folding) against the oracle without a GPU.  Test code only."""
import torch
import torch.nn.functional as F

from se3_equi_graph_registration_b200.packing import OFF, H


def _m(pack, name, *shape):
    n = 1
    for s in shape:
        n *= s
    return pack[OFF[name]:OFF[name] + n].reshape(*shape)


def pq(pack, h):
    return h @ _m(pack, "WPT", 32, 32), h @ _m(pack, "WQT", 32, 32) + _m(pack, "BQ", 32)


def layer(pack, h, x, P, Q, row, col, ea):
    from oracle.egnn_oracle import edge_geometry, segment_sum
    n = h.shape[0]
    d, radial, dist, dot, so3 = edge_geometry(x, row, col)
    geo = torch.cat([radial, dist, dot, so3], 1)                       # [E,12]
    pre = P[row] + Q[col] + ea * _m(pack, "WEA", 32)[None] + geo @ _m(pack, "WG", 12, 32)
    act = F.silu(pre).view(-1, 4, 8)
    u = torch.einsum("ehi,hio->eho", act, _m(pack, "W2P", 4, 8, 8)).reshape(-1, 32) + _m(pack, "B2", 32)
    m = F.layer_norm(u, (32,), _m(pack, "LNG", 32), _m(pack, "LNB", 32), 1e-5)
    t = F.silu(m @ _m(pack, "WC1", 32, 32).t() + _m(pack, "BC1", 32))
    s = t @ _m(pack, "WC2", 32)
    x2 = x + segment_sum(d * s[:, None], row, n)
    agg = segment_sum(m, row, n)
    hid = F.silu(torch.cat([h, agg], 1) @ _m(pack, "WN1T", 64, 32) + _m(pack, "BN1", 32))
    h2 = h + hid @ _m(pack, "WN2T", 32, 32) + _m(pack, "BN2", 32)
    return h2, x2


def egnn(layer_packs, pin, pout, feat, x, row, col, ea_const=1.0):
    h = feat @ pin[:1024].view(32, 32) + pin[1024:1056]
    E = row.shape[0]
    ea = torch.full((E, 1), ea_const)
    for lp in layer_packs:
        P, Q = pq(lp, h)
        h, x = layer(lp, h, x, P, Q, row, col, ea)
    return h @ pout[:1024].view(32, 32) + pout[1024:1056], x


def head_mlp(hp, z):
    h0 = F.relu(z @ hp[0:2048].view(64, 32) + hp[2048:2080])
    h1 = F.relu(h0 @ hp[2080:2592].view(32, 16) + hp[2592:2608])
    return h1 @ hp[2608:2624] + hp[2624]
