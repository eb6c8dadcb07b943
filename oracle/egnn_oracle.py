"""TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the Equi-GSPR registration hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import this module, and only as the checker / CPU baseline.  The product path
(`se3-equi-graph-registration_b200/`) never imports it and has no CPU fallback.

Parity pins: `tests/golden/*.pt` were produced by executing the reference's OWN classes
(ast-extracted by oracle/ref_loader.py from /root/reference) on seeded synthetic inputs;
tests/test_oracle.py checks this restatement against those fixtures (and against the live
reference when /root/reference exists).  k-NN is the exception: its arithmetic lives in the
un-vendored dependency torch-cluster==1.6.3 (environment.yml:158) which is absent, and no
reference test pins its output -> k-NN PARITY UNPINNED; the algorithm restated in
oracle/knn_oracle.c is torch_cluster's published brute-force CUDA kernel (strict '>' insertion,
ties -> lower index) at the reference's call sites (3dm:1005-1006, evl:1156-1157).

All citations are path:line under /root/reference:
  3dm = src/3dmatch_train_egnn_with_batch.py, evl = src/eval_egnn_metrics.py,
  met = tools/evaluation_metrics.py

Written functionally over a flat state_dict (the checkpoint layout) with
torch CPU ops, generic in dtype (fp32 = the reference's arithmetic, fp64 = tie-breaker).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------
# graph helpers
# ----------------------------------------------------------------------------------------
def edges_from_nbr(nbr):
    """nbr [N,k] int -> (row, col) int64 in torch_cluster.knn_graph(flow='source_to_target')
    layout: edge e = i*k+s has row[e] = nbr[i,s] (neighbour), col[e] = i (centre).
    Call sites: 3dm:1005-1006, evl:1156-1157."""
    n, k = nbr.shape
    row = nbr.reshape(-1).to(torch.int64)
    col = torch.arange(n, dtype=torch.int64).repeat_interleave(k)
    return row, col


def segment_sum(data, ids, num_segments):
    """3dm:343-348 unsorted_segment_sum: zeros[N,C].scatter_add_(0, ids, data)."""
    out = data.new_zeros((num_segments, data.shape[1]))
    out.scatter_add_(0, ids.unsqueeze(-1).expand(-1, data.shape[1]), data)
    return out


# ----------------------------------------------------------------------------------------
# per-edge geometry (3dm:128-181, 271-278)
# ----------------------------------------------------------------------------------------
def edge_geometry(x, row, col):
    """Returns coord_diff [E,3], radial [E,1], dist [E,1], dot [E,1], so3 [E,9]."""
    xr, xc = x[row], x[col]
    d = xr - xc                                             # 3dm:273
    radial = (d * d).sum(-1, keepdim=True)                  # 3dm:274
    dist = torch.norm(d, dim=1, keepdim=True)               # 3dm:179
    dot = (xr * xc).sum(dim=1, keepdim=True)                # 3dm:180
    a = d / (d.norm(dim=1, keepdim=True) + 1e-8)            # 3dm:139-140
    cr = torch.cross(xr, xc, dim=1)                         # 3dm:143
    b = cr / (cr.norm(dim=1, keepdim=True) + 1e-8)          # 3dm:144
    c = torch.cross(a, b, dim=1)                            # 3dm:149
    bad = (a.norm(dim=1) < 1e-6) | (b.norm(dim=1) < 1e-6) | (c.norm(dim=1) < 1e-6)  # 3dm:152-156
    so3 = torch.stack([a, b, c], dim=2)                     # 3dm:159  columns are a,b,c
    eye = torch.eye(3, dtype=x.dtype).expand_as(so3)
    so3 = torch.where(bad[:, None, None], eye, so3)         # 3dm:160-163
    return d, radial, dist, dot, so3.reshape(-1, 9)         # 3dm:165


# ----------------------------------------------------------------------------------------
# E_GCL / EGNN (3dm:185-340)
# ----------------------------------------------------------------------------------------
def num_heads_of(sd, prefix="gcl_0."):
    n = 0
    while f"{prefix}edge_mlps.{n}.0.weight" in sd:
        n += 1
    return n


def egcl_edge_messages(sd, p, h, x, row, col, edge_attr):
    """edge_model (3dm:231-250) -> m [E,H]; also returns coord_diff."""
    d, radial, dist, dot, so3 = edge_geometry(x, row, col)
    feats = [h[row], h[col], radial, dist, dot, so3]        # 3dm:238
    if edge_attr is not None:
        feats.append(edge_attr)                             # 3dm:240-241
    f = torch.cat(feats, dim=1)                             # 3dm:242  [E,77]
    heads = []
    for g in range(num_heads_of(sd, p)):                    # 3dm:245
        u = F.linear(f, sd[f"{p}edge_mlps.{g}.0.weight"], sd[f"{p}edge_mlps.{g}.0.bias"])
        u = F.silu(u)
        u = F.linear(u, sd[f"{p}edge_mlps.{g}.2.weight"], sd[f"{p}edge_mlps.{g}.2.bias"])
        heads.append(u)
    comb = torch.cat(heads, dim=1)                          # 3dm:246
    m = F.layer_norm(comb, (comb.shape[1],), sd[f"{p}layer_norm.weight"],
                     sd[f"{p}layer_norm.bias"], 1e-5)       # 3dm:249
    return m, d


def egcl_forward(sd, p, h, x, row, col, edge_attr):
    """E_GCL.forward (3dm:280-289).  Everything from the PRE-update h and x."""
    n = h.shape[0]
    m, d = egcl_edge_messages(sd, p, h, x, row, col, edge_attr)
    # coord_model 3dm:262-268 (SUM aggregation over row)
    s = F.linear(F.silu(F.linear(m, sd[f"{p}coord_mlp.0.weight"], sd[f"{p}coord_mlp.0.bias"])),
                 sd[f"{p}coord_mlp.2.weight"])
    x_new = x + segment_sum(d * s, row, n)
    # node_model 3dm:252-260
    agg = segment_sum(m, row, n)
    o = torch.cat([h, agg], dim=1)
    o = F.linear(F.silu(F.linear(o, sd[f"{p}node_mlp.0.weight"], sd[f"{p}node_mlp.0.bias"])),
                 sd[f"{p}node_mlp.2.weight"], sd[f"{p}node_mlp.2.bias"])
    return h + o, x_new, m


def n_layers_of(sd):
    n = 0
    while f"gcl_{n}.node_mlp.0.weight" in sd:
        n += 1
    return n


def egnn_forward(sd, h, x, row, col, edge_attr, return_layers=False):
    """EGNN.forward (3dm:328-340)."""
    h = F.linear(h, sd["embedding_in.weight"], sd["embedding_in.bias"])
    layers = []
    for i in range(n_layers_of(sd)):
        h, x, _ = egcl_forward(sd, f"gcl_{i}.", h, x, row, col, edge_attr)
        if return_layers:
            layers.append((h.clone(), x.clone()))
    h = F.linear(h, sd["embedding_out.weight"], sd["embedding_out.bias"])
    if return_layers:
        return h, x, layers
    return h, x


# ----------------------------------------------------------------------------------------
# losses (3dm:860-962)
# ----------------------------------------------------------------------------------------
def egnn_equi_loss(h_src, x_src, h_tgt, x_tgt, R_gt, t_gt, labels):
    """3dm:860-893."""
    xs = torch.einsum("bij,bnj->bni", R_gt, x_src) + t_gt[:, None, :]
    ch = ((xs - x_tgt) ** 2).sum(-1)
    rot = (ch * labels).mean()
    cs = F.cosine_similarity(h_src, h_tgt, dim=-1)
    return rot + F.mse_loss(cs, labels.to(cs.dtype))


def pose_loss(pred_rot, pred_t, gt_pose):
    """3dm:896-962 (the two returned losses only)."""
    gt_t, gt_R = gt_pose[:, :3, 3], gt_pose[:, :3, :3]
    Rd = torch.matmul(pred_rot.transpose(-1, -2), gt_R)
    tr = Rd.diagonal(dim1=-2, dim2=-1).sum(-1)
    rl = torch.arccos(torch.clamp((tr - 1) / 2, -1, 1))
    cos = (pred_t * gt_t).sum(-1) / (pred_t.norm(dim=-1) * gt_t.norm(dim=-1))
    return rl, torch.arccos(torch.clamp(cos, -1, 1))


# ----------------------------------------------------------------------------------------
# Kabsch (3dm:726-758 / evl:786-818)
# ----------------------------------------------------------------------------------------
def kabsch(p, q, w):
    """Weighted Procrustes.  p,q [n,3], w [n] (already normalised by the caller).
    Returns R[3,3], t[3], H[3,3] (H includes the +1e-6*I regulariser)."""
    if p.shape[0] == 0:                                     # 3dm:708-711
        return torch.eye(3, dtype=p.dtype), torch.zeros(3, dtype=p.dtype), torch.zeros(3, 3, dtype=p.dtype)
    cs = (w[:, None] * p).sum(0)
    ct = (w[:, None] * q).sum(0)
    pc, qc = p - cs, q - ct
    H = (w[:, None, None] * pc[:, :, None] * qc[:, None, :]).sum(0)
    H = H + 1e-6 * torch.eye(3, dtype=p.dtype)
    U, S, Vt = torch.linalg.svd(H)
    R = Vt.T @ U.T
    if torch.det(R) < 0:                                    # 3dm:749-751
        Vt = Vt.clone()
        Vt[-1, :] *= -1
        R = Vt.T @ U.T
    t = ct - R @ cs
    return R, t, H


def mlp_head(sd, z):
    """CrossAttentionPoseRegression.mlp 64->32->16->1 with ReLU (3dm:594-600)."""
    z = F.relu(F.linear(z, sd["mlp.0.weight"], sd["mlp.0.bias"]))
    z = F.relu(F.linear(z, sd["mlp.2.weight"], sd["mlp.2.bias"]))
    return F.linear(z, sd["mlp.4.weight"], sd["mlp.4.bias"])


def eval_weights_one(sd, h_in_s, h_in_t, h_out_s, h_out_t, top_k=128):
    """evl:691-783 for ONE pair (the reference forward is B=1-only).
    Reproduces the broadcasting quirk literally with the same torch semantics:
    pred[128] vs sims_topk[1,128,1] -> [1,128,128]; scatter_ reads src[0,i,0]."""
    n = h_in_s.shape[0]
    sim0 = (h_in_s * h_in_t).sum(-1, keepdim=True)[None]            # [1,N,1]  evl:691
    _, top_idx = torch.topk(sim0.squeeze(-1), k=top_k, dim=-1)       # [1,128]  evl:694
    ch = torch.cat([h_out_s[top_idx[0]], h_out_t[top_idx[0]]], dim=-1)  # evl:698-699,736
    pred = mlp_head(sd, ch).squeeze(-1)                              # [128]    evl:742
    sims_topk = torch.gather(sim0, 1, top_idx.unsqueeze(-1))         # [1,128,1] evl:758
    c1 = (pred > 0.5) & (torch.abs(pred - 1) < sims_topk)            # evl:761
    c2 = (pred > 0.5) & (pred < sims_topk)                           # evl:762
    fw_topk = torch.where(c1 | c2, pred, sims_topk)                  # [1,128,128] evl:764
    fw = sim0.clone()
    fw.scatter_(dim=1, index=top_idx.unsqueeze(-1), src=fw_topk)     # evl:768
    fw = fw / (fw.sum(dim=1, keepdim=True) + 1e-6)                   # evl:771
    w = F.softmax(fw.squeeze(0).squeeze(-1), dim=-1)                 # evl:773-774
    w = w / (w.sum() + 1e-6)                                         # evl:783
    return w, top_idx[0], pred


def train_weights_one(h_out_s, h_out_t, labels_b):
    """3dm:696-724 for one batch item: valid set = labels != 0; w = softmax(sim)/(sum+1e-6)."""
    valid = labels_b.bool()
    ws = (h_out_s[valid] * h_out_t[valid]).sum(-1)
    w = F.softmax(ws, dim=-1)
    w = w / (w.sum() + 1e-6)
    return w, valid


# ----------------------------------------------------------------------------------------
# full forwards
# ----------------------------------------------------------------------------------------
def _strip(sd_cross):
    egnn_sd = {k[len("egnn."):]: v for k, v in sd_cross.items() if k.startswith("egnn.")}
    return egnn_sd


def forward_eval(sd_cross, h_src, x_src, edges_src, h_tgt, x_tgt, edges_tgt, labels, gt_pose,
                 return_aux=False):
    """Eval-variant CrossAttentionPoseRegression.forward (evl:643-827), applied per pair.
    edges_* : [B,2,E] int64.  Returns the reference's 9-tuple."""
    esd = _strip(sd_cross)
    B, N, _ = h_src.shape
    hs, xs, ht, xt = [], [], [], []
    for b in range(B):
        ea = torch.ones(edges_src.shape[-1], 1, dtype=h_src.dtype)   # get_edges_batch 3dm:387
        a, c = egnn_forward(esd, h_src[b], x_src[b], edges_src[b, 0], edges_src[b, 1], ea)
        hs.append(a); xs.append(c)
        ea = torch.ones(edges_tgt.shape[-1], 1, dtype=h_src.dtype)
        a, c = egnn_forward(esd, h_tgt[b], x_tgt[b], edges_tgt[b, 0], edges_tgt[b, 1], ea)
        ht.append(a); xt.append(c)
    hs, xs, ht, xt = map(torch.stack, (hs, xs, ht, xt))
    loss = egnn_equi_loss(hs, xs, ht, xt, gt_pose[:, :3, :3], gt_pose[:, :3, -1], labels)  # evl:687
    R = torch.zeros(B, 3, 3, dtype=h_src.dtype)
    t = torch.zeros(B, 3, dtype=h_src.dtype)
    aux = {"w": [], "H": [], "top_idx": [], "pred": []}
    for b in range(B):
        w, top_idx, pred = eval_weights_one(sd_cross, h_src[b], h_tgt[b], hs[b], ht[b])
        R[b], t[b], H = kabsch(x_src[b], x_tgt[b], w)                # ORIGINAL coords, all N (evl:717-718)
        aux["w"].append(w); aux["H"].append(H); aux["top_idx"].append(top_idx); aux["pred"].append(pred)
    out = (R, t, None, loss, hs, xs, ht, xt, labels)
    return (out, aux) if return_aux else out


def forward_train(sd_cross, h_src, x_src, edges_src, h_tgt, x_tgt, edges_tgt, labels, gt_pose,
                  top_k=128, return_aux=False):
    """Train-variant forward (3dm:634-796)."""
    esd = _strip(sd_cross)
    B, N, _ = h_src.shape
    hs, xs, ht, xt = [], [], [], []
    for b in range(B):
        ea = torch.ones(edges_src.shape[-1], 1, dtype=h_src.dtype)
        a, c = egnn_forward(esd, h_src[b], x_src[b], edges_src[b, 0], edges_src[b, 1], ea)
        hs.append(a); xs.append(c)
        a, c = egnn_forward(esd, h_tgt[b], x_tgt[b], edges_tgt[b, 0], edges_tgt[b, 1], ea)
        ht.append(a); xt.append(c)
    hs, xs, ht, xt = map(torch.stack, (hs, xs, ht, xt))
    equi = egnn_equi_loss(hs, xs, ht, xt, gt_pose[:, :3, :3], gt_pose[:, :3, -1], labels)  # 3dm:677
    sim = (hs * ht).sum(-1, keepdim=True)                            # 3dm:681
    raw = (h_src * h_tgt).sum(-1, keepdim=True)                      # 3dm:682
    _, top_idx = torch.topk(sim.squeeze(-1), k=top_k, dim=-1)        # 3dm:684
    F_ = h_src.shape[-1]
    chs = torch.gather(hs, 1, top_idx.unsqueeze(-1).expand(-1, -1, F_))
    cht = torch.gather(ht, 1, top_idx.unsqueeze(-1).expand(-1, -1, F_))
    cl = torch.gather(labels, 1, top_idx)                            # 3dm:694
    R = torch.zeros(B, 3, 3, dtype=h_src.dtype)
    t = torch.zeros(B, 3, dtype=h_src.dtype)
    aux = {"w": [], "H": [], "valid": []}
    for b in range(B):
        w, valid = train_weights_one(hs[b], ht[b], labels[b])
        R[b], t[b], H = kabsch(xs[b][valid], xt[b][valid], w)        # EGNN coords, GT inliers (3dm:703-704)
        aux["w"].append(w); aux["H"].append(H); aux["valid"].append(valid)
    scores = mlp_head(sd_cross, torch.cat([chs, cht], -1).view(-1, 2 * F_)).view(B, top_k)  # 3dm:760-770
    corr_loss = F.binary_cross_entropy_with_logits(scores, cl)       # 3dm:772-773
    simn = (sim - sim.mean()) / (sim.std() + 1e-6)                   # 3dm:777
    rawn = (raw - raw.mean()) / (raw.std() + 1e-6)                   # 3dm:778
    sim_loss = F.mse_loss(simn, rawn)                                # 3dm:781
    out = (R, t, corr_loss + sim_loss, equi, hs, xs, ht, xt, labels)
    return (out, aux) if return_aux else out


# ----------------------------------------------------------------------------------------
# metrics (tools/evaluation_metrics.py:14-43) -- numpy float64 like the reference
# ----------------------------------------------------------------------------------------
def calculate_pose_error(gt_pose, pred_pose):
    """met:14-24.  Returns (rotation error in degrees, translation error in CENTIMETRES)."""
    te = np.linalg.norm(gt_pose[:3, 3] - pred_pose[:3, 3]) * 100.0
    rd = gt_pose[:3, :3].T @ pred_pose[:3, :3]
    re = np.degrees(np.arccos(np.clip((np.trace(rd) - 1.0) / 2.0, -1.0, 1.0)))
    return re, te


def registration_recall(gt_pose, pred_pose, src_pts, tgt_pts, tau=0.09):
    """met:26-43.  recall = sqrt(TP/N), precision = TP/N."""
    st = (pred_pose[:3, :3] @ src_pts.T).T + pred_pose[:3, 3]
    dist = np.linalg.norm(st - tgt_pts, axis=1)
    tp = np.sum(dist < tau)
    recall = np.sqrt(tp / len(src_pts))
    precision = tp / len(st) if len(st) > 0 else 0.0
    return recall, precision


def rotation_angle_deg(Ra, Rb):
    tr = float(np.trace(np.asarray(Ra, dtype=np.float64).T @ np.asarray(Rb, dtype=np.float64)))
    return math.degrees(math.acos(max(-1.0, min(1.0, (tr - 1.0) / 2.0))))
