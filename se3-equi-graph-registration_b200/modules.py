"""Host-side mirror of the reference's Python interface for the registration hot path.

Same class / function names, constructor + forward signatures, sub-module and parameter names as
the reference (paths under its tree; 3dm = src/3dmatch_train_egnn_with_batch.py,
evl = src/eval_egnn_metrics.py), so `checkpoints/checkpoint-3dmatch.pth` loads unchanged and the
classes drop into the train / eval scripts:

  E_GCL                          3dm:185-289
  EGNN                           3dm:293-340
  CrossAttentionPoseRegression   3dm:585-796 (train variant) / evl:594-827 (eval variant)
  unsorted_segment_sum           3dm:343-348
  knn_graph                      torch_cluster.knn_graph call sites 3dm:1005-1006, evl:1156-1157
  get_edges_batch                3dm:380-403
  save_checkpoint / load_checkpoint  3dm:1310-1395

The forward passes run on the hand-written sm_100a kernels through the C ABI (ops.py); there is no
CPU or eager-PyTorch fallback -- CPU tensors raise.
"""
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import autograd as _ag
from . import ops, packing


# ---------------------------------------------------------------------------------------------
# functions
# ---------------------------------------------------------------------------------------------
def knn_graph(x, k, batch=None, loop=False, flow="source_to_target", cosine=False, num_workers=1):
    """torch_cluster.knn_graph replacement (same signature).  x [N,3] CUDA f32 -> LongTensor [2,N*k]
    with edge_index[0] = neighbour ids (k per centre, nearest first), edge_index[1] = centre ids."""
    if batch is not None or cosine:
        raise NotImplementedError("egspr_b200.knn_graph: use knn_graph_batch for batches; cosine is unsupported")
    if x.dim() != 2 or x.shape[1] != 3:
        raise ValueError("x must be [N,3]")
    if loop:                                     # the reference's call sites (3dm:1005): self included
        nbr = ops.knn_build(x.unsqueeze(0).to(torch.float32), k)
    else:
        # torch_cluster's default: the query point itself (by INDEX) is not a neighbour.  k + 1 candidates, drop the
        # entry equal to the centre id if it is among them, else the farthest one (> k exact duplicates of a point)
        n = x.shape[0]
        if k + 1 > 32:
            raise NotImplementedError("loop=False supports k <= 31")
        cand = ops.knn_build(x.unsqueeze(0).to(torch.float32), k + 1)[0].long()          # [N, k+1]
        centre = torch.arange(n, device=x.device)[:, None]
        is_self = cand == centre
        drop = torch.where(is_self.any(1), is_self.float().argmax(1), torch.full((n,), k, device=x.device))
        keep = torch.arange(k + 1, device=x.device)[None, :] != drop[:, None]
        nbr = cand[keep].view(1, n, k).to(torch.int32).contiguous()
    edges = ops.nbr_to_edges(nbr)[0]
    if flow == "target_to_source":
        edges = edges.flip(0)
    return edges


def knn_graph_batch(x, k):
    """x [B,N,3] -> edges LongTensor [B,2,N*k]: the stacked per-item knn_graph(x[i], k, loop=True)
    the reference assembles in a Python loop (3dm:1003-1013), in one launch."""
    return ops.nbr_to_edges(ops.knn_build(x.to(torch.float32), k))


def get_edges_batch(graph_idx, n_nodes, batch_size):
    """3dm:380-403 -> ([row, col], edge_attr = ones[E*batch_size, 1]) on the graph's device."""
    row, col = graph_idx[0], graph_idx[1]
    edge_attr = torch.ones(row.numel() * batch_size, 1, device=row.device)
    if batch_size == 1:
        return [row, col], edge_attr
    rows = [row + n_nodes * i for i in range(batch_size)]
    cols = [col + n_nodes * i for i in range(batch_size)]
    return [torch.cat(rows), torch.cat(cols)], edge_attr


def get_edges_from_idx(graph_idx):
    """3dm:372-378: [2,E] graph index -> [src, dst]."""
    return [graph_idx[0], graph_idx[1]]


def unsorted_segment_mean(data, segment_ids, num_segments):
    """3dm:351-358 (the commented-out mean variant of coord_model, 3dm:266): segment sums / max(count, 1)."""
    s = unsorted_segment_sum(data, segment_ids, num_segments)
    cnt = torch.bincount(segment_ids, minlength=num_segments).clamp(min=1).to(s.dtype)
    return s / cnt[:, None]


# user-supplied edge indices are range-checked on the device while the CSR is built (bad ids are clamped and flagged);
# the module entry points read the flag back -- one host sync per call, skipped during CUDA-graph capture -- and raise
# like the reference's index_select / scatter_add_ would.  The engine / training step build their graphs from k-NN ids.
CHECK_EDGE_INDICES = True


def _checked(graph):
    if CHECK_EDGE_INDICES and not torch.cuda.is_current_stream_capturing():
        graph.check()
    return graph


def unsorted_segment_sum(data, segment_ids, num_segments):
    """3dm:343-348, deterministic (ascending edge order) instead of atomics.  Forward only: the model's own segment sums
    are fused into the layer kernels (and differentiated there); a tensor that requires grad is refused instead of
    silently dropping its gradient."""
    from . import _lib
    if torch.is_grad_enabled() and data.requires_grad:
        raise NotImplementedError("unsorted_segment_sum is not differentiable here: use E_GCL / EGNN (their backward kernels "
                                  "differentiate the fused segment sums) or detach the input")
    data = ops._req(data, "data", torch.float32, 2)
    ids = ops._req(segment_ids, "segment_ids", torch.int64, 1)
    E, C = data.shape
    edges = torch.stack([ids, torch.zeros_like(ids)]).unsqueeze(0)
    g = _checked(ops.csr_from_edges(edges, num_segments))
    out = torch.empty((num_segments, C), dtype=torch.float32, device=data.device)
    with torch.cuda.device(data.device):
        _lib.check(_lib.lib().egspr_segment_sum(ops._ptr(data), C, ops._ptr(g.ptr), ops._ptr(g.eid), num_segments,
                                                ops._ptr(out), ops._stream()), "egspr_segment_sum")
    return out


def _needs_grad(module, *tensors):
    """True when autograd is recording and a parameter or an input wants a gradient: the training path."""
    return torch.is_grad_enabled() and (any(p.requires_grad for p in module.parameters()) or
                                        any(torch.is_tensor(t) and t.requires_grad for t in tensors))


def _edges_to_tensor(edges):
    """[row, col] list / tuple / [2,E] tensor -> [1,2,E] int64 contiguous."""
    if isinstance(edges, (list, tuple)):
        edges = torch.stack([edges[0], edges[1]])
    if edges.dim() == 2:
        edges = edges.unsqueeze(0)
    return edges.to(torch.int64).contiguous()


# ---------------------------------------------------------------------------------------------
# E_GCL
# ---------------------------------------------------------------------------------------------
class E_GCL(nn.Module):
    """Multi-head E(n) graph conv layer with the 77-wide geometric edge input (3dm:185-289).
    Parameter layout identical to the reference; forward runs the fused CUDA layer kernel."""

    def __init__(self, input_nf, output_nf, hidden_nf, edges_in_d=0, num_heads=1,
                 act_fn=nn.SiLU(), residual=True, attention=False, normalize=False, tanh=False, device='cuda:0'):
        super().__init__()
        self.residual = residual
        self.attention = attention
        self.normalize = normalize
        self.tanh = tanh
        self.num_heads = num_heads
        self.device = device
        input_edge = input_nf * 2
        edge_coords_nf = 1
        so3_feat_dim = 9
        feature_dim = input_edge + edges_in_d + edge_coords_nf + so3_feat_dim + 2          # 3dm:199
        self.edge_mlps = nn.ModuleList([
            nn.Sequential(nn.Linear(feature_dim, hidden_nf // num_heads), act_fn,
                          nn.Linear(hidden_nf // num_heads, hidden_nf // num_heads))
            for _ in range(num_heads)])
        self.layer_norm = nn.LayerNorm(hidden_nf)
        self.node_mlp = nn.Sequential(nn.Linear(hidden_nf + input_nf, hidden_nf), act_fn,
                                      nn.Linear(hidden_nf, output_nf))
        layer = nn.Linear(hidden_nf, edge_coords_nf, bias=False)
        nn.init.xavier_uniform_(layer.weight, gain=1e-3)                                   # 3dm:220
        coord_mlp = [nn.Linear(hidden_nf, hidden_nf), act_fn, layer]
        if self.tanh:
            coord_mlp.append(nn.Tanh())
        self.coord_mlp = nn.Sequential(*coord_mlp)
        self._act_is_silu = isinstance(act_fn, nn.SiLU)
        self._pack = packing.PackCache(lambda: list(self.parameters()), lambda: packing.pack_layer(self))

    def _check_supported(self):
        if not self._act_is_silu or self.tanh or self.normalize or not self.residual:
            raise NotImplementedError(
                "egspr_b200 E_GCL kernels implement the shipped configuration only: act_fn=SiLU, residual=True, "
                "normalize=False, tanh=False (3dm:1600-1603)")

    def layer_pack(self):
        self._check_supported()
        return self._pack.get()

    def forward(self, h, edge_index, coord, edge_attr=None):
        """(h [N,32], [row,col], coord [N,3], edge_attr [E,1]|None) -> (h', coord', edge_attr)  3dm:280-289"""
        graph = _checked(ops.csr_from_edges(_edges_to_tensor(edge_index), h.shape[0]))
        ea = None if edge_attr is None else edge_attr.to(torch.float32)
        if _needs_grad(self, h, coord):
            self._check_supported()
            spec = ([self], None, None, ops.with_csc(graph), ea, 0.0)
            h2, x2 = _ag.EGNNFunction.apply(spec, h.unsqueeze(0).to(torch.float32), coord.unsqueeze(0).to(torch.float32),
                                            *_ag.egnn_param_list([self]))
            return h2[0], x2[0], edge_attr
        h2, x2 = ops.egnn_forward(h.unsqueeze(0), coord.unsqueeze(0), graph, [self.layer_pack()], None, None,
                                  edge_attr=ea, edge_attr_const=0.0)
        return h2[0], x2[0], edge_attr


# ---------------------------------------------------------------------------------------------
# EGNN
# ---------------------------------------------------------------------------------------------
class EGNN(nn.Module):
    """3dm:293-340.  `num_heads` is an extra keyword (default 4; any divisor of 32): the shipped checkpoints carry 4
    edge-MLP heads per layer although the reference constructor never forwards num_heads
    (SURVEY F2), so 4 is what makes `load_state_dict(strict=True)` succeed."""

    def __init__(self, in_node_nf, hidden_nf, out_node_nf, in_edge_nf=0, device='cuda:0', act_fn=nn.SiLU(),
                 n_layers=5, residual=True, attention=True, normalize=False, tanh=False, num_heads=4):
        super().__init__()
        self.hidden_nf = hidden_nf
        self.device = device
        self.n_layers = n_layers
        self.embedding_in = nn.Linear(in_node_nf, self.hidden_nf)
        self.embedding_out = nn.Linear(self.hidden_nf, out_node_nf)
        for i in range(0, n_layers):
            self.add_module("gcl_%d" % i, E_GCL(self.hidden_nf, self.hidden_nf, self.hidden_nf, edges_in_d=in_edge_nf,
                                                num_heads=num_heads, act_fn=act_fn, residual=residual,
                                                attention=attention, normalize=normalize, tanh=tanh, device=device))
        self._pack_in = packing.PackCache(lambda: list(self.embedding_in.parameters()),
                                          lambda: packing.pack_linear32(self.embedding_in))
        self._pack_out = packing.PackCache(lambda: list(self.embedding_out.parameters()),
                                           lambda: packing.pack_linear32(self.embedding_out))
        self.impl = 0
        self.num_heads = num_heads
        self.to(device)                                                                     # 3dm:326

    def check_impl(self, impl):
        """The CUDA-core kernels (impl 1 / 2) read the per-head [4][8][8] layout of the second edge Linear: 4 heads only.
        The tensor-core kernels (impl 0 / 3 / 4 / 5 and the backward pass) take any head count dividing 32."""
        if int(impl) in (1, 2) and self.num_heads != 4:
            raise NotImplementedError(f"impl {int(impl)} (CUDA-core E_GCL kernels) supports num_heads=4 only; this EGNN has "
                                      f"{self.num_heads} -- use the tensor-core path (impl 0)")

    def packs(self):
        layers = [self._modules["gcl_%d" % i].layer_pack() for i in range(self.n_layers)]
        return layers, self._pack_in.get(), self._pack_out.get()

    def forward_batch(self, h, x, graph, edge_attr=None, edge_attr_const=1.0):
        """Batched form used by the head and the engine: h [C,N,32], x [C,N,3], graph = ops.BatchGraph."""
        if _needs_grad(self, h, x):
            # training path (3dm:1092-1126): forward keeps the per-layer state, backward = the gradient kernels
            gcls = [self._modules["gcl_%d" % i] for i in range(self.n_layers)]
            for g in gcls:
                g._check_supported()
            spec = (gcls, self.embedding_in, self.embedding_out, ops.with_csc(graph), edge_attr, float(edge_attr_const))
            return _ag.EGNNFunction.apply(spec, h, x, *_ag.egnn_param_list(gcls, self.embedding_in, self.embedding_out))
        self.check_impl(self.impl)
        layers, pin, pout = self.packs()
        return ops.egnn_forward(h, x, graph, layers, pin, pout, edge_attr=edge_attr,
                                edge_attr_const=edge_attr_const, impl=self.impl)

    def forward(self, h, x, edges, edge_attr):
        """(h [N,32], x [N,3], [row,col], edge_attr [E,1]) -> (h [N,32], x [N,3])  3dm:328-340"""
        graph = _checked(ops.csr_from_edges(_edges_to_tensor(edges), h.shape[0]))
        ea = None if edge_attr is None else edge_attr.to(torch.float32)
        ho, xo = self.forward_batch(h.unsqueeze(0).to(torch.float32), x.unsqueeze(0).to(torch.float32), graph,
                                    edge_attr=ea, edge_attr_const=0.0)
        return ho[0], xo[0]


# ---------------------------------------------------------------------------------------------
# losses (small reductions on device tensors)
# ---------------------------------------------------------------------------------------------
def egnn_equi_loss(h_src, x_src, h_tgt, x_tgt, R_gt, t_gt, labels):
    """3dm:860-893."""
    xs = torch.einsum("bij,bnj->bni", R_gt, x_src) + t_gt[:, None, :]
    rot = (((xs - x_tgt) ** 2).sum(-1) * labels).mean()
    cs = F.cosine_similarity(h_src, h_tgt, dim=-1)
    return rot + F.mse_loss(cs, labels.float())


def pose_loss(pred_rot, pred_translation, gt_pose, delta=1.5):
    """3dm:896-962 -> (rotation_loss [B], translation_loss [B]); one kernel (egspr_pose_loss) on CUDA tensors,
    differentiable w.r.t. pred_rot and pred_translation."""
    return _ag.PoseLossFunction.apply(pred_rot.to(torch.float32), pred_translation.to(torch.float32), gt_pose.to(torch.float32))


def compute_losses(rot, translation, h_src_norm, x_src, h_tgt_norm, x_tgt, gt_labels):
    """3dm:799-858 (logged by the training loop, 3dm:1094): mean inlier point error under the predicted pose and the
    mean feature distance over the inliers -> (point_error, feature_loss)."""
    xs = torch.matmul(rot, x_src.transpose(1, 2)).transpose(1, 2) + translation.unsqueeze(1)
    dist = torch.norm(xs - x_tgt, dim=-1) * gt_labels
    per_pair = dist.sum(dim=1) / torch.clamp(gt_labels.sum(dim=1), min=1)
    sel = gt_labels == 1
    return per_pair.mean(), torch.norm(h_src_norm[sel] - h_tgt_norm[sel], dim=-1).mean()


# ---------------------------------------------------------------------------------------------
# CrossAttentionPoseRegression
# ---------------------------------------------------------------------------------------------
class CrossAttentionPoseRegression(nn.Module):
    """EGNN on both clouds + correspondence-weight head + weighted Kabsch pose.

    `variant`: 'train' (default) = the class of the training scripts, 3dm:634-796 (weights = softmax of
    output-feature similarity over the GT inliers, Kabsch on EGNN coords; used unchanged under
    model.eval() by the reference's validate(), 3dm:1141, 1272-1290), 'eval' = the class of
    src/eval_egnn_metrics.py, evl:643-827 (weights from the input-feature similarity / top-128 /
    mlp chain, Kabsch on the original coords, all points; the reference body only works for B=1 --
    here every pair of the batch is treated as its own B=1 call).  The variant never follows
    self.training: the two reference files define two different classes, not two modes.
    All parameters of the reference exist (incl. the dead shared_mlp_decoder / shallow_mlp_pose /
    bn1 / bn2) so strict checkpoint loading works."""

    def __init__(self, egnn, num_nodes=2048, hidden_nf=33, device='cuda:0', variant='train'):
        super().__init__()
        self.egnn = egnn
        self.hidden_nf = hidden_nf
        self.num_nodes = num_nodes
        self.device = device
        if variant not in ("train", "eval"):
            raise ValueError("variant must be 'train' (3dm / kit class) or 'eval' (evl class)")
        self.variant = variant
        self.mlp = nn.Sequential(nn.Linear(2 * hidden_nf, hidden_nf), nn.ReLU(),
                                 nn.Linear(hidden_nf, hidden_nf // 2), nn.ReLU(),
                                 nn.Linear(hidden_nf // 2, 1))
        self.shared_mlp_decoder = nn.Sequential(nn.Linear((32 + 3) * 2, 128), nn.ReLU(), nn.Linear(128, 64), nn.ReLU())
        self.shallow_mlp_pose = nn.Sequential(nn.Linear(64, 32), nn.ReLU(), nn.Linear(32, 7))
        self.global_pooling = nn.AdaptiveMaxPool1d(1)
        self.mean_pooling = nn.AdaptiveAvgPool1d(1)
        self.bn1 = nn.BatchNorm1d(self.hidden_nf)
        self.bn2 = nn.BatchNorm1d(self.hidden_nf + 3)
        self.initialize_weights()
        self._pack_head = packing.PackCache(lambda: list(self.mlp.parameters()), lambda: packing.pack_head(self.mlp))
        self.top_k = 128

    def initialize_weights(self):
        for layer in self.mlp:
            if isinstance(layer, nn.Linear):
                nn.init.xavier_uniform_(layer.weight)
                nn.init.zeros_(layer.bias)

    def _variant(self):
        return self.variant

    def _egnn_both(self, h_src, x_src, edges_src, edge_attr_src, h_tgt, x_tgt, edges_tgt, edge_attr_tgt):
        B, N, _ = h_src.shape
        tensors = torch.is_tensor(edges_src) and torch.is_tensor(edges_tgt) and edges_src.shape == edges_tgt.shape
        if tensors and (edge_attr_src is None) == (edge_attr_tgt is None):
            # both cloud sets as ONE graph of 2B clouds: one launch sequence (sources = clouds 0..B-1, targets B..2B-1)
            g = _checked(ops.csr_from_edges(torch.cat([edges_src, edges_tgt]).to(torch.int64), N))
            ea = None if edge_attr_src is None else torch.cat([edge_attr_src, edge_attr_tgt])
            h, x = self.egnn.forward_batch(torch.cat([h_src, h_tgt]).to(torch.float32), torch.cat([x_src, x_tgt]).to(torch.float32),
                                           g, edge_attr=ea, edge_attr_const=0.0 if ea is not None else 1.0)
            return h[:B], x[:B], h[B:], x[B:]
        g_src = edges_src if isinstance(edges_src, ops.BatchGraph) else _checked(ops.csr_from_edges(edges_src.to(torch.int64), N))
        g_tgt = edges_tgt if isinstance(edges_tgt, ops.BatchGraph) else _checked(ops.csr_from_edges(edges_tgt.to(torch.int64), N))
        hs, xs = self.egnn.forward_batch(h_src.to(torch.float32), x_src.to(torch.float32), g_src,
                                         edge_attr=edge_attr_src, edge_attr_const=0.0 if edge_attr_src is not None else 1.0)
        ht, xt = self.egnn.forward_batch(h_tgt.to(torch.float32), x_tgt.to(torch.float32), g_tgt,
                                         edge_attr=edge_attr_tgt, edge_attr_const=0.0 if edge_attr_tgt is not None else 1.0)
        return hs, xs, ht, xt

    def forward(self, h_src, x_src, edges_src, edge_attr_src, h_tgt, x_tgt, edges_tgt, edge_attr_tgt, corr, labels, gt_pose):
        """Same 11 inputs / 9 outputs as the reference (3dm:634, 796; evl:643, 827):
        (R [B,3,3], t [B,3], corr_loss+sim_loss | None, egnn_equi_loss, h_src, x_src, h_tgt, x_tgt, labels).
        edges_* : [B,2,E] int64 (or a prebuilt ops.BatchGraph); edge_attr_* : [B,E,1] or None (= ones)."""
        B, N, _ = h_src.shape
        labels_f = labels.to(torch.float32).reshape(B, N)
        hs, xs, ht, xt = self._egnn_both(h_src, x_src, edges_src, edge_attr_src, h_tgt, x_tgt, edges_tgt, edge_attr_tgt)
        if self._variant() == "eval":
            R, t, w, Hm, lp = ops.head_eval(h_src.to(torch.float32), h_tgt.to(torch.float32), x_src.to(torch.float32),
                                            x_tgt.to(torch.float32), hs, ht, xs, xt, labels_f, gt_pose,
                                            self._pack_head.get(), top_k=self.top_k)
            total_loss = lp.sum(0).sum() / (B * N)                                           # evl:687
            self.last_aux = {"w": w, "H": Hm}
            return R, t, None, total_loss, hs, xs, ht, xt, labels
        grad = _needs_grad(self, h_src, h_tgt, x_src, x_tgt)
        if grad:
            R, t, sim, w, Hm, lp = _ag.HeadTrainFunction.apply(hs, ht, xs, xt, labels_f, gt_pose.to(torch.float32))
            total_loss = _ag.EquiLossFunction.apply(egnn_equi_loss, hs, xs, ht, xt, gt_pose[:, :3, :3].to(torch.float32),
                                                    gt_pose[:, :3, -1].to(torch.float32), labels_f, lp)   # 3dm:677, differentiable
        else:
            R, t, w, sim, Hm, lp = ops.head_train(hs, ht, xs, xt, labels_f, gt_pose)
            total_loss = lp.sum(0).sum() / (B * N)                                           # 3dm:677
        self.last_aux = {"w": w, "H": Hm}
        # correspondence BCE on the top-128 + similarity-consistency loss (3dm:681-694, 760-781): two kernels
        # (egspr_train_loss_forward / _finalize), differentiable w.r.t. the EGNN outputs, sim and mlp
        fs, ft = h_src.to(torch.float32).contiguous(), h_tgt.to(torch.float32).contiguous()
        spec = (self._pack_head.get(), int(self.top_k), self.mlp)
        if grad:
            corr_sim = _ag.CorrSimLossFunction.apply(spec, hs, ht, fs, ft, sim, labels_f, *self.mlp.parameters())
        else:
            top_idx, scores, raw, stats, bce = ops.train_loss_forward(hs, ht, fs, ft, sim, labels_f, spec[0], spec[1])
            lv, _, _, _ = ops.train_loss_finalize(sim, raw, stats, bce, spec[1], need_grad=False)
            corr_sim = lv[0] + lv[1]
        return R, t, corr_sim, total_loss, hs, xs, ht, xt, labels


# ---------------------------------------------------------------------------------------------
# checkpoints (3dm:1310-1395)
# ---------------------------------------------------------------------------------------------
def save_checkpoint(epoch, pointnet, egnn, cross_attention, optimizer, save_dir="./checkpoints2", is_best=False, use_pointnet=False):
    """Same signature, dict layout and file names as 3dm:1310-1349: `model_epoch_{epoch}.pth` (+ `best_checkpoint` when
    is_best) holding epoch / egnn_state_dict / cross_attention_state_dict / optimizer_state_dict (/ pointnet_state_dict).
    Returns the path written (the reference returns None)."""
    os.makedirs(save_dir, exist_ok=True)
    ck = {"epoch": epoch, "egnn_state_dict": egnn.state_dict(),
          "cross_attention_state_dict": cross_attention.state_dict()}
    if optimizer is not None:
        ck["optimizer_state_dict"] = optimizer.state_dict()
    if use_pointnet and pointnet is not None:
        ck["pointnet_state_dict"] = pointnet.state_dict()
    path = os.path.join(save_dir, f"model_epoch_{epoch}.pth")
    torch.save(ck, path)
    if is_best:
        torch.save(ck, os.path.join(save_dir, "best_checkpoint"))
    return path


def load_checkpoint(checkpoint_path, pointnet=None, egnn=None, cross_attention=None, optimizer=None, use_pointnet=False, device='cuda:0'):
    """Same contract as 3dm:1351-1395: strict load of 'egnn_state_dict' and 'cross_attention_state_dict' (+ optional
    pointnet / optimizer) for the modules that are given; returns (checkpoint, epoch)."""
    if not os.path.exists(checkpoint_path):
        raise FileNotFoundError(f"Checkpoint file {checkpoint_path} not found.")
    ck = torch.load(checkpoint_path, map_location=device, weights_only=True)
    if use_pointnet and pointnet is not None and "pointnet_state_dict" in ck:
        pointnet.load_state_dict(ck["pointnet_state_dict"])
    if egnn is not None and "egnn_state_dict" in ck:
        egnn.load_state_dict(ck["egnn_state_dict"])
    if cross_attention is not None and "cross_attention_state_dict" in ck:
        cross_attention.load_state_dict(ck["cross_attention_state_dict"])
    if optimizer is not None and "optimizer_state_dict" in ck:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return ck, ck.get("epoch", 0)
