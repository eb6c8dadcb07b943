"""The training step of the reference's loop (src/3dmatch_train_egnn_with_batch.py:1092-1126) on the CUDA path,
plus the data-parallel gradient exchange of BASELINE config 4.

One process per GPU, every rank holds a replica and its own slice of the pair batch (a registration pair is an
independent unit); the only collective is ONE all-reduce of the flat fp32 gradient of the live parameters (25,953
elements: one contiguous, persistent buffer -- the parameters' .grad are views of it) per step, after which every rank
applies the same optimizer update.  The reference itself has no distributed code.

Two forms of the step:
  train_step        the loop body as written in the reference: model(...) -> training_loss -> loss.backward() ->
                    optimizer.step(), through the nn.Module API and torch.autograd (whose backward passes are the gradient
                    kernels).  Drop-in, eager.
  GraphedTrainStep  the same arithmetic as a fixed launch sequence (36 launches, no autograd, no per-parameter tensor
                    ops), captured once as a CUDA graph: k-NN -> CSR -> weight packs (one gather) -> EGNN forward ->
                    head + losses -> backward kernels -> gradients (one gather) -> all-reduce -> optimizer.
"""
import torch
import torch.distributed as dist

from . import ops, packing
from .modules import pose_loss


def training_loss(outputs, gt_pose):
    """3dm:1094-1118: total = corr_loss.mean() + rot_loss.mean() + trans_loss.mean() (the similarity loss is inside
    slot 2; egnn_equi_loss, slot 3, is returned by the model but not part of the total)."""
    R, t, corr_loss = outputs[0], outputs[1], outputs[2]
    rot, trans = pose_loss(R, t, gt_pose.to(R.dtype), delta=1.5)
    return corr_loss.mean() + rot.mean() + trans.mean()


def _world(group=None):
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


def allreduce_gradients(params, group=None, average=True):
    """Sum (or mean) over the ranks of the gradients of the parameters that HAVE one, in a single bucket; parameters
    whose grad is None on every rank (the dead modules of SURVEY F8) stay None, so the optimizer skips them exactly as
    in single-process training.  No-op without an initialised process group or with world size 1."""
    world = _world(group)
    if world == 1:
        return
    grads = [p.grad for p in params if p.requires_grad and p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat.div_(world)
    torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])


def train_step(model, optimizer, batch, group=None):
    """One iteration of train_one_epoch (3dm:1092-1126): forward (train variant), loss, backward through the
    gradient kernels, gradient all-reduce, optimizer step.  batch: the 11 forward inputs as a tuple / list.
    Returns the detached loss."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    outputs = model(*batch)
    loss = training_loss(outputs, batch[10])
    loss.backward()
    allreduce_gradients(model.parameters(), group=group)
    optimizer.step()
    return loss.detach()


class GraphedTrainStep:
    """The training step as ONE CUDA graph of 36 launches.

    batch = (src_feat, src_pts, tgt_feat, tgt_pts, corr, labels, gt_pose), as the reference's loader yields them
    (3dm:975-979); the k-NN graphs of 3dm:1003-1089 are built inside the step, directly as CSR (edge_attr = the reference's
    all-ones, get_edges_batch 3dm:387, as a constant).  Construction re-homes the model's parameters into one flat buffer
    (packing.FlatState: same Parameter objects, same state_dict) and runs `warmup` eager steps to create the optimizer
    state, then RESTORES parameters and optimizer state, so building the step does not train the model.
    The optimizer must be capturable (torch.optim.Adam(..., capturable=True)); its zero_grad() must not be called with
    set_to_none=True afterwards (the gradients are persistent views and are overwritten every step)."""

    def __init__(self, model, optimizer, example_batch, k=16, group=None, warmup=2):
        self.model, self.opt, self.k, self.group = model, optimizer, int(k), group
        self.world = _world(group)
        sf, sp, tf, tp, corr, labels, gt = example_batch
        self.B, self.N = int(sp.shape[0]), int(sp.shape[1])
        dev = sp.device
        B = self.B
        self.state = packing.FlatState(model)
        # static input buffers of the graph (own storage: .to() / .contiguous() alias the example batch when nothing changes)
        self.feat_all = torch.cat([sf, tf]).to(torch.float32).contiguous()
        self.x_all = torch.cat([sp, tp]).to(torch.float32).contiguous()
        self.labels_f = labels.to(torch.float32).reshape(B, self.N).clone()
        self.gt_pose = gt.to(torch.float32).clone()
        self.top_k = int(model.top_k)
        self._side = None
        # optimizer state before the warm-up (None = not created yet)
        saved_state = {id(p): {n: (v.clone() if torch.is_tensor(v) else v) for n, v in st.items()} for p, st in self.opt.state.items()}
        saved_flat = self.state.flat.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the capture: lazy inits, caches, optimizer state
            for _ in range(max(1, warmup)):
                self._step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            self.state.flat.copy_(saved_flat)
            for p, st in self.opt.state.items():
                old = saved_state.get(id(p))
                for n, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(old[n]) if old is not None else v.zero_()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss_vec = self._step()
        self.loss = self.loss_vec[4]

    def _step(self):
        st, B, dev = self.state, self.B, self.x_all.device
        st.refresh_packs()
        graph = ops.with_csc(ops.csr_from_nbr(ops.knn_build(self.x_all, self.k)))
        h, x, saved = ops.egnn_forward_saved(self.feat_all, self.x_all, graph, st.layer_packs, st.pack_in, st.pack_out)
        hs, ht, xs, xt = h[:B], h[B:], x[:B], x[B:]
        # the two one-CTA-per-pair kernels side by side (16 + 16 CTAs on 148 SMs): the loss kernel recomputes the
        # similarity instead of reading head_train's, so it only depends on the EGNN outputs
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side = torch.cuda.Stream(device=dev)
        louts = ops.train_loss_outputs(B, self.N, self.top_k, dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            top_idx, _, raw, stats, bce = ops.train_loss_forward(hs, ht, self.feat_all[:B], self.feat_all[B:], None, self.labels_f,
                                                                 st.pack_head, self.top_k, outs=louts)
        R, t, _, sim, _, _ = ops.head_train(hs, ht, xs, xt, self.labels_f, self.gt_pose)
        cur.wait_stream(self._side)
        # total = corr + sim + mean rot + mean trans (3dm:1118); the seeds carry 1 / world so that the SUM all-reduce
        # below yields the mean gradient
        loss, dsim, dR, dt = ops.train_loss_finalize(sim, raw, stats, bce, self.top_k, R, t, self.gt_pose, scale=1.0 / self.world)
        st.gpack_buf.zero_()
        dh, dx = torch.empty_like(h), torch.empty_like(x)
        ops.head_train_loss_backward(hs, ht, xs, xt, self.labels_f, dR, dt, dsim, top_idx, st.pack_head, loss, st.gpack_head,
                                     self.top_k, outs=[dh[:B], dh[B:], dx[:B], dx[B:]])
        ops.egnn_backward(saved, graph, st.layer_packs, st.pack_in, st.pack_out, dh, dx, need_dfeat=False,
                          gpacks_out=(st.layer_gpacks, st.gpack_in, st.gpack_out))
        st.gather_gradients()
        if self.world > 1:
            dist.all_reduce(st.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.opt.step()
        return loss

    def load(self, batch):
        sf, sp, tf, tp, corr, labels, gt = batch
        B = self.B
        self.feat_all[:B].copy_(sf, non_blocking=True); self.feat_all[B:].copy_(tf, non_blocking=True)
        self.x_all[:B].copy_(sp, non_blocking=True); self.x_all[B:].copy_(tp, non_blocking=True)
        self.labels_f.copy_(labels.reshape(B, self.N), non_blocking=True)
        self.gt_pose.copy_(gt, non_blocking=True)

    def __call__(self, batch):
        self.load(batch)
        self.graph.replay()
        packing.invalidate_packs()      # the replay rewrote the parameters behind the version counters
        return self.loss
