"""torch.autograd glue for the training step (3dm:1092-1126): two Functions whose backward passes are the
hand-written kernels (ops.egnn_backward, ops.head_train_backward).  PyTorch only routes the gradients between them
and onto the nn.Parameters' .grad -- the parameters stay the single source of truth, the weight packs and gradient
packs are derived re-layouts (packing.py)."""
import torch

from . import ops, packing


def _unpacker(module, fn, floats):
    u = module.__dict__.get("_egspr_grad_unpacker")
    if u is None:
        u = module.__dict__["_egspr_grad_unpacker"] = packing.GradUnpacker(fn, floats, module)
    return u


def _linear_pack(lin):
    c = lin.__dict__.get("_egspr_pack_cache")
    if c is None:
        c = lin.__dict__["_egspr_pack_cache"] = packing.PackCache(lambda: list(lin.parameters()), lambda: packing.pack_linear32(lin))
    return c.get()


def egnn_param_list(egnn_or_layers, embedding_in=None, embedding_out=None):
    """Flat parameter list in the order EGNNFunction returns gradients:
    [embedding_in.weight, .bias]? + [embedding_out.weight, .bias]? + every layer's parameters()."""
    params = []
    if embedding_in is not None:
        params += [embedding_in.weight, embedding_in.bias]
    if embedding_out is not None:
        params += [embedding_out.weight, embedding_out.bias]
    for gcl in egnn_or_layers:
        params += list(gcl.parameters())
    return params


class EGNNFunction(torch.autograd.Function):
    """(feat [C,N,32], x [C,N,3], *params) -> (h_out, x_out); `spec` carries the non-tensor context."""

    @staticmethod
    def forward(ctx, spec, feat, x, *params):
        layers, emb_in, emb_out, graph, edge_attr, edge_attr_const = spec
        layer_packs = [g.layer_pack() for g in layers]
        pin = _linear_pack(emb_in) if emb_in is not None else None
        pout = _linear_pack(emb_out) if emb_out is not None else None
        h_out, x_out, saved = ops.egnn_forward_saved(feat, x, graph, layer_packs, pin, pout, edge_attr=edge_attr,
                                                     edge_attr_const=edge_attr_const)
        ctx.spec = spec
        ctx.saved = saved
        ctx.packs = (layer_packs, pin, pout)
        return h_out, x_out

    @staticmethod
    def backward(ctx, dh_out, dx_out):
        layers, emb_in, emb_out, graph, _, _ = ctx.spec
        layer_packs, pin, pout = ctx.packs
        need_dfeat = ctx.needs_input_grad[1]
        dfeat, dx, gpacks, g_in, g_out = ops.egnn_backward(ctx.saved, graph, layer_packs, pin, pout, dh_out, dx_out,
                                                           need_dfeat=need_dfeat or emb_in is None)
        grads = []
        if emb_in is not None:
            grads += _unpacker(emb_in, packing.unpack_linear32_grad, packing.EMBED_PACK)(g_in)
        if emb_out is not None:
            grads += _unpacker(emb_out, packing.unpack_linear32_grad, packing.EMBED_PACK)(g_out)
        for gcl, gp in zip(layers, gpacks):
            grads += _unpacker(gcl, packing.unpack_layer_grad, packing.LAYER_PACK)(gp)
        return (None, dfeat if need_dfeat else None, dx, *grads)


class HeadTrainFunction(torch.autograd.Function):
    """(h_src_out, h_tgt_out, x_src_out, x_tgt_out, labels, gt_pose) -> (R, t, sim, w, H, loss_parts); gradients flow
    through R, t and sim (3dm:681, 696-758)."""

    @staticmethod
    def forward(ctx, hs, ht, xs, xt, labels_f, gt_pose):
        R, t, w, sim, Hm, lp = ops.head_train(hs, ht, xs, xt, labels_f, gt_pose)
        ctx.save_for_backward(hs, ht, xs, xt, labels_f)
        ctx.mark_non_differentiable(w, Hm, lp)
        return R, t, sim, w, Hm, lp

    @staticmethod
    def backward(ctx, dR, dt, dsim, *_):
        hs, ht, xs, xt, labels_f = ctx.saved_tensors
        B = hs.shape[0]
        if dR is None:
            dR = torch.zeros((B, 3, 3), dtype=torch.float32, device=hs.device)
        if dt is None:
            dt = torch.zeros((B, 3), dtype=torch.float32, device=hs.device)
        dhs, dht, dxs, dxt = ops.head_train_backward(hs, ht, xs, xt, labels_f, dR, dt, dsim)
        return dhs, dht, dxs, dxt, None, None


class PoseLossFunction(torch.autograd.Function):
    """pose_loss (3dm:896-962) as one kernel: (R, t, gt_pose) -> (rot_loss [B], trans_loss [B]); the kernel also
    returns the two local gradients, so backward is a broadcast multiply."""

    @staticmethod
    def forward(ctx, R, t, gt_pose):
        rl, tl, gR, gt = ops.pose_loss(R, t, gt_pose, need_grad=True)
        ctx.save_for_backward(gR, gt)
        return rl, tl

    @staticmethod
    def backward(ctx, d_rl, d_tl):
        gR, gt = ctx.saved_tensors
        dR = gR * d_rl.view(-1, 1, 1) if d_rl is not None else None
        dt = gt * d_tl.view(-1, 1) if d_tl is not None else None
        return dR, dt, None


class EquiLossFunction(torch.autograd.Function):
    """egnn_equi_loss (3dm:860-893, slot 3 of the forward's 9-tuple) without its forward cost: the value comes from the
    partial sums the head kernel already produced; the gradient -- only needed when the caller adds this slot to the
    training loss, which the reference loop does not (3dm:1118) -- is computed on demand."""

    @staticmethod
    def forward(ctx, loss_fn, hs, xs, ht, xt, R_gt, t_gt, labels_f, loss_parts):
        ctx.loss_fn = loss_fn
        ctx.save_for_backward(hs, xs, ht, xt, R_gt, t_gt, labels_f)
        B, N = labels_f.shape
        return loss_parts.sum(0).sum() / (B * N)

    @staticmethod
    def backward(ctx, g):
        hs, xs, ht, xt, R_gt, t_gt, labels_f = ctx.saved_tensors
        with torch.enable_grad():
            leaves = [v.detach().requires_grad_(True) for v in (hs, xs, ht, xt)]
            loss = ctx.loss_fn(leaves[0], leaves[1], leaves[2], leaves[3], R_gt, t_gt, labels_f)
            grads = torch.autograd.grad(loss, leaves, g)
        return (None, *grads, None, None, None, None)


class CorrSimLossFunction(torch.autograd.Function):
    """corr_loss + sim_loss (slot 2 of the train-variant forward, 3dm:681-694, 760-781) as two kernels
    (egspr_train_loss_forward / _finalize); backward = the mean-BCE backward through mlp (egspr_head_train_loss_backward
    with zero pose seeds) and the closed-form gradient of the z-scored similarity MSE.
    (hs, ht, feat_src, feat_tgt, sim, labels_f, *mlp parameters) -> scalar; gradients flow to hs, ht, sim and mlp."""

    @staticmethod
    def forward(ctx, spec, hs, ht, fs, ft, sim, labels_f, *mlp_params):
        head_pack, top_k, mlp = spec
        top_idx, scores, raw, stats, bce = ops.train_loss_forward(hs, ht, fs, ft, sim, labels_f, head_pack, top_k)
        loss, dsim, _, _ = ops.train_loss_finalize(sim, raw, stats, bce, top_k, scale=1.0, need_grad=True)
        ctx.save_for_backward(hs, ht, labels_f, top_idx, dsim, loss)
        ctx.spec = spec
        ctx.aux = {"top_idx": top_idx, "scores": scores, "loss": loss}
        return loss[0] + loss[1]

    @staticmethod
    def backward(ctx, g):
        hs, ht, labels_f, top_idx, dsim, loss = ctx.saved_tensors
        head_pack, top_k, mlp = ctx.spec
        B, n, _ = hs.shape
        lvec = loss.clone()
        lvec[7] = g                                           # upstream gradient, read by the kernel from device memory
        zR = torch.zeros((B, 3, 3), dtype=torch.float32, device=hs.device)
        zt = torch.zeros((B, 3), dtype=torch.float32, device=hs.device)
        zx = torch.zeros((B, n, 3), dtype=torch.float32, device=hs.device)
        gp = torch.zeros(packing.HEAD_PACK, dtype=torch.float32, device=hs.device)
        # zero pose seeds: the Kabsch part of the kernel contributes nothing, only the BCE-through-mlp part remains
        dhs, dht, _, _ = ops.head_train_loss_backward(hs, ht, zx, zx, labels_f, zR, zt, None, top_idx, head_pack, lvec, gp, top_k)
        grads = _unpacker(mlp, packing.unpack_head_grad, packing.HEAD_PACK)(gp)
        return (None, dhs, dht, None, None, dsim * g, None, *grads)
