"""Thin functional layer over the C ABI (include/egspr_b200.h): argument checking, output
allocation with torch (device memory + streams are PyTorch's job), ctypes calls on the current
CUDA stream.  Every function requires CUDA tensors; there is no CPU path."""
import ctypes
from dataclasses import dataclass

import torch

from . import _lib

H = 32


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _req(t, name, dtype, ndim=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the egspr_b200 hot path has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name} must have {ndim} dims, got shape {tuple(t.shape)}")
    return t.contiguous()


def knn_build(x, k, brute_force=False):
    """x [C,N,3] f32 -> nbr [C,N,k] i32 (nearest first, ties -> lower index, self included).
    brute_force=True runs the O(N^2) scan instead of the cell-grid search (identical output)."""
    x = _req(x, "x", torch.float32, 3)
    C, N, D = x.shape
    if D != 3:
        raise ValueError("k-NN is built over 3-d points")
    nbr = torch.empty((C, N, k), dtype=torch.int32, device=x.device)
    if C * N == 0:
        return nbr
    lib = _lib.lib()
    ws, ws_bytes = None, 0
    if not brute_force:
        ws_bytes = lib.egspr_knn_workspace_bytes(C, N)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.egspr_knn_build(_ptr(x), C, N, k, _ptr(nbr), _ptr(ws), ws_bytes, _stream()), "egspr_knn_build")
    return nbr


def build_train_graph(x, k, out=None):
    """k-NN graph of x [C,N,3] as the CSR + column positions the training step needs (knn_build -> csr_from_nbr ->
    with_csc), written into the buffer set `out` of an earlier call when given: same addresses on every call, so a
    captured CUDA graph can consume a graph that another stream rebuilds between its replays.  -> BatchGraph (its
    private `_bufs` holds the two workspaces)."""
    x = _req(x, "x", torch.float32, 3)
    C, N, D = x.shape
    if D != 3:
        raise ValueError("k-NN is built over 3-d points")
    if out is None:
        return with_csc(csr_from_nbr(knn_build(x, k)))
    lib = _lib.lib()
    g = out
    if (g.clouds, g.n, g.edges_per_cloud) != (C, N, N * k) or g.nbr is None:
        raise ValueError("the buffer set was built for another shape")
    dev = x.device
    bufs = g.__dict__.get("_bufs")
    if bufs is None:
        bufs = g.__dict__["_bufs"] = (torch.empty(lib.egspr_knn_workspace_bytes(C, N), dtype=torch.uint8, device=dev),
                                      torch.empty(lib.egspr_csr_workspace_bytes(C * N, C * N * k), dtype=torch.uint8, device=dev))
    kws, cws = bufs
    with torch.cuda.device(dev):
        st = _stream()
        _lib.check(lib.egspr_knn_build(_ptr(x), C, N, k, _ptr(g.nbr), _ptr(kws), kws.numel(), st), "egspr_knn_build")
        _lib.check(lib.egspr_csr_from_nbr(_ptr(g.nbr), C, N, k, _ptr(g.ptr), _ptr(g.row), _ptr(g.col), _ptr(g.eid),
                                          _ptr(cws), cws.numel(), _ptr(g.err), st), "egspr_csr_from_nbr")
        _lib.check(lib.egspr_csr_edge_positions(_ptr(g.row), _ptr(g.eid), None, N, N * k, C * N * k, _ptr(g.cpos), None, st),
                   "egspr_csr_edge_positions")
    return g


def nbr_to_edges(nbr):
    """nbr [C,N,k] i32 -> edges [C,2,N*k] i64 in torch_cluster.knn_graph layout."""
    nbr = _req(nbr, "nbr", torch.int32, 3)
    C, N, k = nbr.shape
    edges = torch.empty((C, 2, N * k), dtype=torch.int64, device=nbr.device)
    with torch.cuda.device(nbr.device):
        _lib.check(_lib.lib().egspr_nbr_to_edges(_ptr(nbr), C, N, k, _ptr(edges), _stream()), "egspr_nbr_to_edges")
    return edges


@dataclass
class BatchGraph:
    """Row-major CSR of a batch of per-cloud graphs (global node ids)."""
    ptr: torch.Tensor     # [G+1] i32
    row: torch.Tensor     # [E] i32
    col: torch.Tensor     # [E] i32
    eid: torch.Tensor     # [E] i32  original edge id inside the cloud
    clouds: int
    n: int
    edges_per_cloud: int
    err: torch.Tensor     # [1] i32, 1 if an index was out of range
    cptr: torch.Tensor = None   # [G+1] i32  the same edges grouped by col (backward pass only, see with_csc)
    cpos: torch.Tensor = None   # [E] i32    row-CSR position of every entry of the col-grouped lists
    edges: torch.Tensor = None  # the [C,2,E] i64 edge tensor the graph was built from (kept for with_csc)
    nbr: torch.Tensor = None    # ... or the k-NN ids it was built from

    def check(self):
        if int(self.err.item()) != 0:
            raise IndexError("edge index out of range for the given number of nodes")


def _csr_alloc(C, N, epc, device):
    G, E = C * N, C * epc
    lib = _lib.lib()
    ws_bytes = lib.egspr_csr_workspace_bytes(G, E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=device)
    g = BatchGraph(torch.empty(G + 1, dtype=torch.int32, device=device),
                   torch.empty(E, dtype=torch.int32, device=device),
                   torch.empty(E, dtype=torch.int32, device=device),
                   torch.empty(E, dtype=torch.int32, device=device),
                   C, N, epc, torch.zeros(1, dtype=torch.int32, device=device))
    return g, ws, ws_bytes


def csr_from_nbr(nbr):
    nbr = _req(nbr, "nbr", torch.int32, 3)
    C, N, k = nbr.shape
    g, ws, ws_bytes = _csr_alloc(C, N, N * k, nbr.device)
    with torch.cuda.device(nbr.device):
        _lib.check(_lib.lib().egspr_csr_from_nbr(_ptr(nbr), C, N, k, _ptr(g.ptr), _ptr(g.row), _ptr(g.col), _ptr(g.eid),
                                                 _ptr(ws), ws_bytes, _ptr(g.err), _stream()), "egspr_csr_from_nbr")
    g.nbr = nbr
    return g


def csr_from_edges(edges, n):
    """edges [C,2,E] i64 (row = edge_index[0], col = edge_index[1], cloud-local ids)."""
    edges = _req(edges, "edges", torch.int64, 3)
    C, two, E = edges.shape
    if two != 2:
        raise ValueError("edges must be [C,2,E]")
    if E == 0:
        raise ValueError("empty edge list")
    g, ws, ws_bytes = _csr_alloc(C, n, E, edges.device)
    with torch.cuda.device(edges.device):
        _lib.check(_lib.lib().egspr_csr_from_edges(_ptr(edges), C, n, E, _ptr(g.ptr), _ptr(g.row), _ptr(g.col), _ptr(g.eid),
                                                   _ptr(ws), ws_bytes, _ptr(g.err), _stream()), "egspr_csr_from_edges")
    g.edges = edges
    return g


_IMPLICIT_CSC = {}


def with_csc(graph, nbr=None):
    """Adds the col-grouped edge lists the backward pass needs: graph.cptr [G+1] and graph.cpos [E] = the row-CSR position
    of every entry (the per-edge gradients are stored in row-CSR order).  k-NN graphs: edge e = i * k + s has col = i, so
    the col-grouped order is the original order -- cptr[g] = g * k is a constant of the shape and cpos is one small
    kernel; general graphs: the CSR of the edge tensor with its two rows swapped gives the lists, then the positions."""
    if graph.cptr is not None:
        return graph
    nbr = nbr if nbr is not None else graph.nbr
    lib = _lib.lib()
    C, N, epc = graph.clouds, graph.n, graph.edges_per_cloud
    E = C * epc
    dev = graph.ptr.device
    pos = torch.empty(E, dtype=torch.int32, device=dev)
    if nbr is not None and graph.edges is None:
        k = nbr.shape[2]
        key = (C, N, k, str(dev))
        cptr = _IMPLICIT_CSC.get(key)
        if cptr is None:
            cptr = _IMPLICIT_CSC[key] = (torch.arange(C * N + 1, dtype=torch.int64, device=dev) * k).to(torch.int32)
        with torch.cuda.device(dev):
            _lib.check(lib.egspr_csr_edge_positions(_ptr(graph.row), _ptr(graph.eid), None, N, epc, E, _ptr(pos), None, _stream()),
                       "egspr_csr_edge_positions")
        graph.cptr, graph.cpos = cptr, pos
        return graph
    if graph.edges is None:
        raise ValueError("with_csc needs the edge tensor (or the k-NN ids) the graph was built from")
    t = csr_from_edges(graph.edges.flip(1).contiguous(), N)
    cpos = torch.empty(E, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.egspr_csr_edge_positions(_ptr(graph.row), _ptr(graph.eid), _ptr(t.eid), N, epc, E, _ptr(pos), _ptr(cpos), _stream()),
                   "egspr_csr_edge_positions")
    graph.cptr, graph.cpos = t.ptr, cpos
    return graph


def egnn_forward(feat, x, graph, layer_packs, embed_in_pack, embed_out_pack, edge_attr=None,
                 edge_attr_const=1.0, impl=0, return_layers=False):
    """EGNN.forward (3dm:328-340) for all clouds of a batch.
    feat [C,N,32], x [C,N,3] -> h_out [C,N,32], x_out [C,N,3].
    embed_in_pack / embed_out_pack may be None (stand-alone E_GCL stack)."""
    feat = _req(feat, "h", torch.float32, 3)
    x = _req(x, "x", torch.float32, 3)
    C, N, F = feat.shape
    if F != H:
        raise NotImplementedError(f"feature width must be {H}, got {F}")
    if (C, N) != (graph.clouds, graph.n) or tuple(x.shape) != (C, N, 3):
        raise ValueError("feature / coordinate / graph shapes disagree")
    if edge_attr is not None:
        edge_attr = _req(edge_attr, "edge_attr", torch.float32).reshape(-1)
        if edge_attr.numel() != C * graph.edges_per_cloud:
            raise ValueError("edge_attr must have one scalar per edge")
    dev = feat.device
    G = C * N
    lib = _lib.lib()
    hbuf = [torch.empty((G, H), dtype=torch.float32, device=dev) for _ in range(2)]
    xbuf = [torch.empty((G, 4), dtype=torch.float32, device=dev) for _ in range(2)]
    pbuf = [torch.empty((G, H), dtype=torch.float32, device=dev) for _ in range(2)]
    qbuf = [torch.empty((G, H), dtype=torch.float32, device=dev) for _ in range(2)]
    x_out = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
    agg_ws = torch.empty((G, H), dtype=torch.float32, device=dev) if impl in (0, 3, 4, 5) else None
    layers = []
    with torch.cuda.device(dev):
        st = _stream()
        _lib.check(lib.egspr_node_embed(_ptr(feat), _ptr(x), G, _ptr(embed_in_pack), _ptr(layer_packs[0]),
                                        _ptr(hbuf[0]), _ptr(xbuf[0]), _ptr(pbuf[0]), _ptr(qbuf[0]), st), "egspr_node_embed")
        cur = 0
        L = len(layer_packs)
        for i in range(L):
            last = i == L - 1
            nxt = 1 - cur
            want_x3 = last or return_layers
            x3 = x_out if last else (torch.empty((C, N, 3), dtype=torch.float32, device=dev) if return_layers else None)
            _lib.check(lib.egspr_egcl_forward(
                _ptr(hbuf[cur]), _ptr(xbuf[cur]), _ptr(pbuf[cur]), _ptr(qbuf[cur]),
                _ptr(graph.ptr), _ptr(graph.row), _ptr(graph.col), _ptr(graph.eid),
                _ptr(edge_attr), float(edge_attr_const), G, graph.edges_per_cloud, N,
                _ptr(layer_packs[i]), None if last else _ptr(layer_packs[i + 1]),
                _ptr(embed_out_pack) if last else None,
                _ptr(hbuf[nxt]), _ptr(xbuf[nxt]), _ptr(x3) if want_x3 else None,
                None if last else _ptr(pbuf[nxt]), None if last else _ptr(qbuf[nxt]), _ptr(agg_ws), int(impl), st),
                "egspr_egcl_forward")
            cur = nxt
            if return_layers and not last:
                layers.append((hbuf[cur].view(C, N, H).clone(), x3))
    h_out = hbuf[cur].view(C, N, H)
    if return_layers:
        return h_out, x_out, layers
    return h_out, x_out


def kabsch(p, q, w, mask=None):
    """Batched weighted Kabsch (3dm:726-758).  p,q [B,n,3], w [B,n] -> R [B,3,3], t [B,3], H [B,3,3]."""
    p = _req(p, "p", torch.float32, 3)
    q = _req(q, "q", torch.float32, 3)
    w = _req(w, "w", torch.float32, 2)
    B, n, _ = p.shape
    if mask is not None:
        mask = _req(mask.to(torch.float32), "mask", torch.float32, 2)
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=p.device)
    t = torch.empty((B, 3), dtype=torch.float32, device=p.device)
    Hm = torch.empty((B, 3, 3), dtype=torch.float32, device=p.device)
    with torch.cuda.device(p.device):
        _lib.check(_lib.lib().egspr_kabsch(_ptr(p), _ptr(q), _ptr(w), _ptr(mask), B, n, _ptr(R), _ptr(t), _ptr(Hm), _stream()),
                   "egspr_kabsch")
    return R, t, Hm


def head_eval(feat_src, feat_tgt, x_src, x_tgt, h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, gt_pose,
              head_pack, top_k=128):
    """Eval-variant weights + Kabsch (evl:691-818), every pair treated as its own B=1 call.
    Returns R [B,3,3], t [B,3], w [B,n], H [B,3,3], loss_parts [B,2]."""
    ts = [_req(v, nm, torch.float32, 3) for v, nm in
          ((feat_src, "feat_src"), (feat_tgt, "feat_tgt"), (x_src, "x_src"), (x_tgt, "x_tgt"),
           (h_out_src, "h_out_src"), (h_out_tgt, "h_out_tgt"), (x_out_src, "x_out_src"), (x_out_tgt, "x_out_tgt"))]
    B, n, _ = ts[0].shape
    labels = _req(labels.to(torch.float32), "labels", torch.float32, 2)
    gt_pose = _req(gt_pose.to(torch.float32), "gt_pose", torch.float32, 3)
    dev = ts[0].device
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    Hm = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    w = torch.empty((B, n), dtype=torch.float32, device=dev)
    lp = torch.empty((B, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().egspr_head_eval(*[_ptr(v) for v in ts], _ptr(labels), _ptr(gt_pose), _ptr(head_pack),
                                              B, n, int(top_k), _ptr(w), _ptr(R), _ptr(t), _ptr(Hm), _ptr(lp), _stream()),
                   "egspr_head_eval")
    return R, t, w, Hm, lp


def head_train(h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, gt_pose):
    """Train-variant weights + Kabsch (3dm:696-758).  Returns R, t, w, sim, H, loss_parts."""
    ts = [_req(v, nm, torch.float32, 3) for v, nm in
          ((h_out_src, "h_out_src"), (h_out_tgt, "h_out_tgt"), (x_out_src, "x_out_src"), (x_out_tgt, "x_out_tgt"))]
    B, n, _ = ts[0].shape
    labels = _req(labels.to(torch.float32), "labels", torch.float32, 2)
    gt_pose = _req(gt_pose.to(torch.float32), "gt_pose", torch.float32, 3)
    dev = ts[0].device
    R = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    t = torch.empty((B, 3), dtype=torch.float32, device=dev)
    Hm = torch.empty((B, 3, 3), dtype=torch.float32, device=dev)
    w = torch.empty((B, n), dtype=torch.float32, device=dev)
    sim = torch.empty((B, n), dtype=torch.float32, device=dev)
    lp = torch.empty((B, 2), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().egspr_head_train(*[_ptr(v) for v in ts], _ptr(labels), _ptr(gt_pose), B, n,
                                               _ptr(w), _ptr(sim), _ptr(R), _ptr(t), _ptr(Hm), _ptr(lp), _stream()),
                   "egspr_head_train")
    return R, t, w, sim, Hm, lp


def pose_metrics(R, t, gt_pose, src_pts, tgt_pts, tau=0.09):
    """tools/evaluation_metrics.py:14-43 + the F1 of evl:1277 for a batch, on the device (fp64).
    R [B,3,3], t [B,3], gt_pose [B,4,4], src_pts / tgt_pts [B,n,3]  ->  float64 [B,5] =
    (rotation error deg, translation error cm, recall, precision, F1)."""
    R = _req(R, "R", torch.float32, 3); t = _req(t, "t", torch.float32, 2)
    gt_pose = _req(gt_pose.to(torch.float32), "gt_pose", torch.float32, 3)
    src_pts = _req(src_pts, "src_pts", torch.float32, 3); tgt_pts = _req(tgt_pts, "tgt_pts", torch.float32, 3)
    B, n, _ = src_pts.shape
    out = torch.empty((B, 5), dtype=torch.float64, device=R.device)
    with torch.cuda.device(R.device):
        _lib.check(_lib.lib().egspr_pose_metrics(_ptr(R), _ptr(t), _ptr(gt_pose), _ptr(src_pts), _ptr(tgt_pts), B, n,
                                                 float(tau), _ptr(out), _stream()), "egspr_pose_metrics")
    return out


def feature_nn(a, b):
    """Nearest descriptor of b for every descriptor of a under the reference's feature distance
    sqrt(2 - 2 <a,b> + 1e-6) (data_preprocess/3DMatch_Feature.py:158-160).
    a [Na,32], b [Nb,32] f32 (unit norm) -> idx [Na] i32 (np.argmin, first index on ties), dist [Na] f32."""
    a = _req(a, "a", torch.float32, 2); b = _req(b, "b", torch.float32, 2)
    if a.shape[1] != H or b.shape[1] != H:
        raise NotImplementedError(f"descriptor width must be {H}")
    na, nb = a.shape[0], b.shape[0]
    idx = torch.empty(na, dtype=torch.int32, device=a.device)
    dist = torch.empty(na, dtype=torch.float32, device=a.device)
    ws = torch.empty(na, dtype=torch.int64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().egspr_feature_nn(_ptr(a), na, _ptr(b), nb, _ptr(ws), ws.numel() * 8, _ptr(idx), _ptr(dist), _stream()),
                   "egspr_feature_nn")
    return idx, dist


def feature_correspondences(src_desc, tgt_desc, use_mutual=False):
    """Correspondence set of data_preprocess/3DMatch_Feature.py:158-166: nearest target descriptor of every
    source descriptor; with use_mutual only the pairs that are each other's nearest neighbour.
    Returns corr [M,2] i64 (source index, target index) and the per-source distances [Ns]."""
    source_idx, source_dis = feature_nn(src_desc, tgt_desc)
    ar = torch.arange(source_idx.numel(), device=source_idx.device)
    if use_mutual:
        target_idx, _ = feature_nn(tgt_desc, src_desc)
        keep = target_idx.long()[source_idx.long()] == ar
        corr = torch.stack([ar[keep], source_idx.long()[keep]], dim=-1)
    else:
        corr = torch.stack([ar, source_idx.long()], dim=-1)
    return corr, source_dis


# ---------------------------------------------------------------------------------------------------------------
# training path: forward that keeps the per-layer state, and the backward kernels
# ---------------------------------------------------------------------------------------------------------------
def linear32_forward(x, pack):
    x = _req(x, "x", torch.float32)
    y = torch.empty_like(x)
    rows = x.numel() // H
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().egspr_linear32_forward(_ptr(x), rows, _ptr(pack), _ptr(y), _stream()), "egspr_linear32_forward")
    return y


def linear32_backward(x, dy, pack, need_dx=True, gp=None):
    """-> (dx | None, grad_pack [EMBED_PACK]); gp: an already-zeroed gradient pack to accumulate into (else a new one)"""
    x = _req(x, "x", torch.float32); dy = _req(dy, "dy", torch.float32)
    rows = x.numel() // H
    dx = torch.empty_like(x) if need_dx else None
    if gp is None:
        gp = torch.zeros(pack.numel(), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().egspr_linear32_backward(_ptr(x), _ptr(dy), rows, _ptr(pack), _ptr(dx), _ptr(gp), _stream()),
                   "egspr_linear32_backward")
    return dx, gp


def egnn_forward_saved(feat, x, graph, layer_packs, embed_in_pack, embed_out_pack, edge_attr=None, edge_attr_const=1.0):
    """EGNN.forward (3dm:328-340) keeping what the backward pass needs: per layer the input state (h, x4, P, Q) and
    the message sums agg.  Tensor-core forward kernels (impl 3); embedding_out runs as its own Linear so that the
    last layer's output survives.  Returns h_out [C,N,32], x_out [C,N,3], saved (dict)."""
    feat = _req(feat, "h", torch.float32, 3)
    x = _req(x, "x", torch.float32, 3)
    C, N, F = feat.shape
    if F != H:
        raise NotImplementedError(f"feature width must be {H}, got {F}")
    if (C, N) != (graph.clouds, graph.n) or tuple(x.shape) != (C, N, 3):
        raise ValueError("feature / coordinate / graph shapes disagree")
    if edge_attr is not None:
        edge_attr = _req(edge_attr, "edge_attr", torch.float32).reshape(-1)
        if edge_attr.numel() != C * graph.edges_per_cloud:
            raise ValueError("edge_attr must have one scalar per edge")
    dev = feat.device
    G, L = C * N, len(layer_packs)
    lib = _lib.lib()
    new = lambda w: torch.empty((G, w), dtype=torch.float32, device=dev)
    hs = [new(H) for _ in range(L + 1)]
    x4 = [new(4) for _ in range(L + 1)]
    Ps = [new(H) for _ in range(L)]
    Qs = [new(H) for _ in range(L)]
    aggs = [new(H) for _ in range(L)]
    x_out = torch.empty((C, N, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        st = _stream()
        _lib.check(lib.egspr_node_embed(_ptr(feat), _ptr(x), G, _ptr(embed_in_pack), _ptr(layer_packs[0]),
                                        _ptr(hs[0]), _ptr(x4[0]), _ptr(Ps[0]), _ptr(Qs[0]), st), "egspr_node_embed")
        for i in range(L):
            last = i == L - 1
            _lib.check(lib.egspr_egcl_forward(
                _ptr(hs[i]), _ptr(x4[i]), _ptr(Ps[i]), _ptr(Qs[i]),
                _ptr(graph.ptr), _ptr(graph.row), _ptr(graph.col), _ptr(graph.eid),
                _ptr(edge_attr), float(edge_attr_const), G, graph.edges_per_cloud, N,
                _ptr(layer_packs[i]), None if last else _ptr(layer_packs[i + 1]), None,
                _ptr(hs[i + 1]), _ptr(x4[i + 1]), _ptr(x_out) if last else None,
                None if last else _ptr(Ps[i + 1]), None if last else _ptr(Qs[i + 1]), _ptr(aggs[i]), 3, st),
                "egspr_egcl_forward")
    h_out = linear32_forward(hs[L], embed_out_pack) if embed_out_pack is not None else hs[L].clone()
    saved = dict(feat=feat, hs=hs, x4=x4, Ps=Ps, Qs=Qs, aggs=aggs, edge_attr=edge_attr, edge_attr_const=float(edge_attr_const))
    return h_out.view(C, N, H), x_out, saved


def egnn_backward(saved, graph, layer_packs, embed_in_pack, embed_out_pack, dh_out, dx_out, need_dfeat=True, gpacks_out=None):
    """Backward of egnn_forward_saved.  dh_out [C,N,32], dx_out [C,N,3] (either may be None = zero).
    gpacks_out: optional (layer grad packs, embed_in grad pack, embed_out grad pack), ALREADY ZEROED, to accumulate into
    (views of one persistent gradient buffer in the training step); otherwise new zero-filled packs are allocated.
    Returns dfeat [C,N,32] | None, dx [C,N,3], layer grad packs (list), embed_in grad pack | None, embed_out grad pack | None."""
    if graph.cptr is None:
        raise ValueError("the backward pass needs graph.cptr / graph.cpos: call ops.with_csc(graph) first")
    C, N = graph.clouds, graph.n
    G, L = C * N, len(layer_packs)
    dev = saved["feat"].device
    lib = _lib.lib()
    dh = torch.zeros((G, H), dtype=torch.float32, device=dev) if dh_out is None else _req(dh_out, "dh_out", torch.float32).reshape(G, H)
    dx = torch.zeros((G, 3), dtype=torch.float32, device=dev) if dx_out is None else _req(dx_out, "dx_out", torch.float32).reshape(G, 3)
    g_out = None
    gl_out, gi_out, go_out = gpacks_out if gpacks_out is not None else (None, None, None)
    if embed_out_pack is not None:
        dh, g_out = linear32_backward(saved["hs"][L], dh, embed_out_pack, need_dx=True, gp=go_out)
    E = C * graph.edges_per_cloud
    ws_bytes = lib.egspr_egcl_backward_workspace_bytes(G, E)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    gpacks = [None] * L
    with torch.cuda.device(dev):
        st = _stream()
        for i in range(L - 1, -1, -1):
            gp = gl_out[i] if gl_out is not None else torch.zeros(layer_packs[i].numel(), dtype=torch.float32, device=dev)
            dh_in = torch.empty((G, H), dtype=torch.float32, device=dev)
            dx_in = torch.empty((G, 3), dtype=torch.float32, device=dev)
            _lib.check(lib.egspr_egcl_backward(
                _ptr(saved["hs"][i]), _ptr(saved["x4"][i]), _ptr(saved["Ps"][i]), _ptr(saved["Qs"][i]), _ptr(saved["aggs"][i]),
                _ptr(graph.ptr), _ptr(graph.row), _ptr(graph.col), _ptr(graph.eid), _ptr(graph.cptr), _ptr(graph.cpos),
                _ptr(saved["edge_attr"]), saved["edge_attr_const"], G, graph.edges_per_cloud, N, _ptr(layer_packs[i]),
                _ptr(dh), _ptr(dx), _ptr(dh_in), _ptr(dx_in), _ptr(gp), _ptr(ws), ws_bytes, st), "egspr_egcl_backward")
            gpacks[i] = gp
            dh, dx = dh_in, dx_in
    g_in, dfeat = None, dh
    if embed_in_pack is not None:
        dfeat, g_in = linear32_backward(saved["feat"].reshape(G, H), dh, embed_in_pack, need_dx=need_dfeat, gp=gi_out)
    return (dfeat.view(C, N, H) if dfeat is not None else None), dx.view(C, N, 3), gpacks, g_in, g_out


def head_train_backward(h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, dR, dt, dsim=None):
    """Backward of head_train (3dm:696-758) -> dh_src, dh_tgt [B,n,32], dx_src, dx_tgt [B,n,3]."""
    ts = [_req(v, nm, torch.float32, 3) for v, nm in
          ((h_out_src, "h_out_src"), (h_out_tgt, "h_out_tgt"), (x_out_src, "x_out_src"), (x_out_tgt, "x_out_tgt"))]
    B, n, _ = ts[0].shape
    labels = _req(labels.to(torch.float32), "labels", torch.float32, 2)
    dR = _req(dR.to(torch.float32), "dR", torch.float32, 3)
    dt = _req(dt.to(torch.float32), "dt", torch.float32, 2)
    if dsim is not None:
        dsim = _req(dsim.to(torch.float32), "dsim", torch.float32, 2)
    outs = [torch.empty_like(ts[0]), torch.empty_like(ts[1]), torch.empty_like(ts[2]), torch.empty_like(ts[3])]
    with torch.cuda.device(ts[0].device):
        _lib.check(_lib.lib().egspr_head_train_backward(*[_ptr(v) for v in ts], _ptr(labels), _ptr(dR), _ptr(dt), _ptr(dsim),
                                                        B, n, *[_ptr(o) for o in outs], _stream()), "egspr_head_train_backward")
    return outs


def pose_loss(R, t, gt_pose, need_grad=True):
    """pose_loss (3dm:896-962) for a batch -> rot_loss [B], trans_loss [B], d rot_loss / d R [B,3,3] | None,
    d trans_loss / d t [B,3] | None."""
    R = _req(R, "R", torch.float32, 3); t = _req(t, "t", torch.float32, 2)
    gt_pose = _req(gt_pose.to(torch.float32), "gt_pose", torch.float32, 3)
    B = R.shape[0]
    rl = torch.empty(B, dtype=torch.float32, device=R.device)
    tl = torch.empty(B, dtype=torch.float32, device=R.device)
    gR = torch.empty((B, 3, 3), dtype=torch.float32, device=R.device) if need_grad else None
    gt = torch.empty((B, 3), dtype=torch.float32, device=R.device) if need_grad else None
    with torch.cuda.device(R.device):
        _lib.check(_lib.lib().egspr_pose_loss(_ptr(R), _ptr(t), _ptr(gt_pose), B, _ptr(rl), _ptr(tl), _ptr(gR), _ptr(gt), _stream()),
                   "egspr_pose_loss")
    return rl, tl, gR, gt


# ---------------------------------------------------------------------------------------------------------------
# training losses on the device
# ---------------------------------------------------------------------------------------------------------------
def train_loss_outputs(B, n, top_k, dev):
    """Output buffers of train_loss_forward (allocate them on the consumer's stream when the kernel runs on another one)."""
    return (torch.empty((B, top_k), dtype=torch.int32, device=dev), torch.empty((B, top_k), dtype=torch.float32, device=dev),
            torch.empty((B, n), dtype=torch.float32, device=dev), torch.empty((B, 4), dtype=torch.float64, device=dev),
            torch.empty(B, dtype=torch.float32, device=dev))


def train_loss_forward(h_out_src, h_out_tgt, feat_src, feat_tgt, sim, labels, head_pack, top_k=128, outs=None):
    """3dm:681-694, 760-773 for a batch -> top_idx [B,k] i32 (the top-k set of sim, in ascending point order), scores [B,k] (mlp logits),
    raw [B,n] (input-feature similarity), stats [B,4] f64, bce [B] (per-pair BCE sums).  sim = head_train's similarity
    output, or None: the kernel recomputes it (same bits) and is then independent of head_train."""
    ts = [_req(v, nm, torch.float32, 3) for v, nm in ((h_out_src, "h_out_src"), (h_out_tgt, "h_out_tgt"),
                                                      (feat_src, "feat_src"), (feat_tgt, "feat_tgt"))]
    if sim is not None:
        sim = _req(sim, "sim", torch.float32, 2)
    labels = _req(labels.to(torch.float32), "labels", torch.float32, 2)
    B, n, _ = ts[0].shape
    dev = ts[0].device
    top_idx, scores, raw, stats, bce = outs if outs is not None else train_loss_outputs(B, n, top_k, dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().egspr_train_loss_forward(*[_ptr(v) for v in ts], _ptr(sim), _ptr(labels), _ptr(head_pack), B, n,
                                                       int(top_k), _ptr(top_idx), _ptr(scores), _ptr(raw), _ptr(stats), _ptr(bce),
                                                       _stream()), "egspr_train_loss_forward")
    return top_idx, scores, raw, stats, bce


def train_loss_finalize(sim, raw, stats, bce, top_k, R=None, t=None, gt_pose=None, scale=1.0, need_grad=True):
    """-> loss [8] = (corr, sim, mean rot, mean trans, total, 0, 0, scale), dsim [B,n] | None, dR [B,3,3] | None,
    dt [B,3] | None (the seeds are multiplied by `scale`); the pose terms need R, t, gt_pose."""
    sim = _req(sim, "sim", torch.float32, 2)
    B, n = sim.shape
    dev = sim.device
    loss = torch.empty(8, dtype=torch.float32, device=dev)
    dsim = torch.empty((B, n), dtype=torch.float32, device=dev) if need_grad else None
    pose = R is not None
    if pose:
        R = _req(R, "R", torch.float32, 3); t = _req(t, "t", torch.float32, 2)
        gt_pose = _req(gt_pose.to(torch.float32), "gt_pose", torch.float32, 3)
    dR = torch.empty((B, 3, 3), dtype=torch.float32, device=dev) if (pose and need_grad) else None
    dt = torch.empty((B, 3), dtype=torch.float32, device=dev) if (pose and need_grad) else None
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().egspr_train_loss_finalize(_ptr(sim), _ptr(raw), _ptr(stats), _ptr(bce), B, n, int(top_k), _ptr(R),
                                                        _ptr(t), _ptr(gt_pose), float(scale), _ptr(loss), _ptr(dsim), _ptr(dR),
                                                        _ptr(dt), _stream()), "egspr_train_loss_finalize")
    return loss, dsim, dR, dt


def head_train_loss_backward(h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, dR, dt, dsim, top_idx, head_pack, loss,
                             head_gpack, top_k=128, outs=None):
    """egspr_head_train_loss_backward: backward of head_train (dR, dt, dsim) plus the mean-BCE backward through mlp.
    head_gpack [HEAD_PACK] is accumulated into (zero it first).  -> dh_src, dh_tgt, dx_src, dx_tgt."""
    ts = [_req(v, nm, torch.float32, 3) for v, nm in
          ((h_out_src, "h_out_src"), (h_out_tgt, "h_out_tgt"), (x_out_src, "x_out_src"), (x_out_tgt, "x_out_tgt"))]
    B, n, _ = ts[0].shape
    labels = _req(labels.to(torch.float32), "labels", torch.float32, 2)
    dR = _req(dR, "dR", torch.float32, 3); dt = _req(dt, "dt", torch.float32, 2)
    if dsim is not None:
        dsim = _req(dsim, "dsim", torch.float32, 2)
    if outs is None:
        outs = [torch.empty_like(ts[0]), torch.empty_like(ts[1]), torch.empty_like(ts[2]), torch.empty_like(ts[3])]
    with torch.cuda.device(ts[0].device):
        _lib.check(_lib.lib().egspr_head_train_loss_backward(*[_ptr(v) for v in ts], _ptr(labels), _ptr(dR), _ptr(dt), _ptr(dsim),
                                                             _ptr(top_idx), _ptr(head_pack), _ptr(loss), B, n, int(top_k),
                                                             *[_ptr(o) for o in outs], _ptr(head_gpack), _stream()),
                   "egspr_head_train_loss_backward")
    return outs
