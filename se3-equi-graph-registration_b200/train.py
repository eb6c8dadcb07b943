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
    (3dm:975-979); the k-NN graphs of 3dm:1003-1089 are built on the device, directly as CSR (edge_attr = the reference's
    all-ones, get_edges_batch 3dm:387, as a constant).  Construction re-homes the model's parameters into one flat buffer
    (packing.FlatState: same Parameter objects, same state_dict) and runs `warmup` eager steps to create the optimizer
    state, then RESTORES parameters and optimizer state, so building the step does not train the model.
    The optimizer must be capturable (torch.optim.Adam(..., capturable=True)); its zero_grad() must not be called with
    set_to_none=True afterwards (the gradients are persistent views and are overwritten every step).

    Pipelined graph build: `step(batch, next_batch=...)` also uploads the NEXT batch into the second input set and builds
    ITS k-NN graph on a forked stream inside the same CUDA graph -- the wide k-NN / CSR kernels run beside the
    one-CTA-per-pair head and loss kernels of the current step (16 CTAs on 148 SMs), and the next call, given that batch,
    starts with its graph ready (a DataLoader iterator knows the next batch one step ahead).  Without `next_batch` every
    call builds its own graph first, as before."""

    def __init__(self, model, optimizer, example_batch, k=16, group=None, warmup=2):
        self.model, self.opt, self.k, self.group = model, optimizer, int(k), group
        self.world = _world(group)
        sf, sp, tf, tp, corr, labels, gt = example_batch
        self.B, self.N = int(sp.shape[0]), int(sp.shape[1])
        dev = sp.device
        B = self.B
        self.state = packing.FlatState(model)
        # two static input sets (own storage: .to() / .contiguous() alias the example batch when nothing changes) and two
        # static graph buffer sets: the captured step `s` trains on set s and rebuilds the graph of set 1 - s
        self.sets = []
        for _ in range(2):
            x_all = torch.cat([sp, tp]).to(torch.float32).contiguous()
            self.sets.append(dict(feat_all=torch.cat([sf, tf]).to(torch.float32).contiguous(), x_all=x_all,
                                  labels_f=labels.to(torch.float32).reshape(B, self.N).clone(), gt_pose=gt.to(torch.float32).clone(),
                                  graph=ops.build_train_graph(x_all, self.k)))
            ops.build_train_graph(x_all, self.k, out=self.sets[-1]["graph"])      # allocates the set's workspaces (outside any capture)
        self.top_k = int(model.top_k)
        self._side = self._side2 = None
        self._cur, self._ready, self._ready_for = 0, [False, False], None
        # optimizer state before the warm-up (None = not created yet)
        saved_state = {id(p): {n: (v.clone() if torch.is_tensor(v) else v) for n, v in st.items()} for p, st in self.opt.state.items()}
        saved_flat = self.state.flat.clone()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                      # warm-up off the capture: lazy inits, caches, optimizer state
            for _ in range(max(1, warmup)):
                self._step(0)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            self.state.flat.copy_(saved_flat)
            for p, st in self.opt.state.items():
                old = saved_state.get(id(p))
                for n, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(old[n]) if old is not None else v.zero_()
        self.graphs, self.loss_vecs = [], []
        for s in range(2):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                lv = self._step(s)
            self.graphs.append(g); self.loss_vecs.append(lv)
        self.graph, self.loss_vec, self.loss = self.graphs[0], self.loss_vecs[0], self.loss_vecs[0][4]

    # the set the next call trains on (tests / tools that drive _step by hand)
    @property
    def feat_all(self): return self.sets[self._cur]["feat_all"]
    @property
    def x_all(self): return self.sets[self._cur]["x_all"]
    @property
    def labels_f(self): return self.sets[self._cur]["labels_f"]
    @property
    def gt_pose(self): return self.sets[self._cur]["gt_pose"]

    def _step(self, s=None):
        s = self._cur if s is None else s
        cur_set, nxt_set = self.sets[s], self.sets[1 - s]
        feat_all, x_all, labels_f, gt_pose, graph = (cur_set[k] for k in ("feat_all", "x_all", "labels_f", "gt_pose", "graph"))
        st, B, dev = self.state, self.B, x_all.device
        cur = torch.cuda.current_stream()
        if self._side is None:
            self._side, self._side2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
        st.refresh_packs()
        h, x, saved = ops.egnn_forward_saved(feat_all, x_all, graph, st.layer_packs, st.pack_in, st.pack_out)
        hs, ht, xs, xt = h[:B], h[B:], x[:B], x[B:]
        # the OTHER set's k-NN graph on a forked stream, starting where the narrow phase of this step starts: wide kernels
        # that fill the SMs the one-CTA-per-pair head / loss kernels leave idle (forked at the top of the step they only
        # competed with the forward edge kernels); joined at the end of the step
        self._side2.wait_stream(cur)
        with torch.cuda.stream(self._side2):
            ops.build_train_graph(nxt_set["x_all"], self.k, out=nxt_set["graph"])
        # the two one-CTA-per-pair kernels side by side (16 + 16 CTAs on 148 SMs): the loss kernel recomputes the
        # similarity instead of reading head_train's, so it only depends on the EGNN outputs
        louts = ops.train_loss_outputs(B, self.N, self.top_k, dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            top_idx, _, raw, stats, bce = ops.train_loss_forward(hs, ht, feat_all[:B], feat_all[B:], None, labels_f,
                                                                 st.pack_head, self.top_k, outs=louts)
        R, t, _, sim, _, _ = ops.head_train(hs, ht, xs, xt, labels_f, gt_pose)
        cur.wait_stream(self._side)
        # total = corr + sim + mean rot + mean trans (3dm:1118); the seeds carry 1 / world so that the SUM all-reduce
        # below yields the mean gradient
        loss, dsim, dR, dt = ops.train_loss_finalize(sim, raw, stats, bce, self.top_k, R, t, gt_pose, scale=1.0 / self.world)
        st.gpack_buf.zero_()
        dh, dx = torch.empty_like(h), torch.empty_like(x)
        ops.head_train_loss_backward(hs, ht, xs, xt, labels_f, dR, dt, dsim, top_idx, st.pack_head, loss, st.gpack_head,
                                     self.top_k, outs=[dh[:B], dh[B:], dx[:B], dx[B:]])
        ops.egnn_backward(saved, graph, st.layer_packs, st.pack_in, st.pack_out, dh, dx, need_dfeat=False,
                          gpacks_out=(st.layer_gpacks, st.gpack_in, st.gpack_out))
        st.gather_gradients()
        if self.world > 1:
            dist.all_reduce(st.flat_grad, op=dist.ReduceOp.SUM, group=self.group)
        self.opt.step()
        cur.wait_stream(self._side2)
        return loss

    def load(self, batch, s=None):
        sf, sp, tf, tp, corr, labels, gt = batch
        B = self.B
        d = self.sets[self._cur if s is None else s]
        d["feat_all"][:B].copy_(sf, non_blocking=True); d["feat_all"][B:].copy_(tf, non_blocking=True)
        d["x_all"][:B].copy_(sp, non_blocking=True); d["x_all"][B:].copy_(tp, non_blocking=True)
        d["labels_f"].copy_(labels.reshape(B, self.N), non_blocking=True)
        d["gt_pose"].copy_(gt, non_blocking=True)

    def __call__(self, batch, next_batch=None):
        s = self._cur
        if not (self._ready[s] and self._ready_for is batch):
            self.load(batch, s)                                  # not announced by the previous call: upload + build now
            ops.build_train_graph(self.sets[s]["x_all"], self.k, out=self.sets[s]["graph"])
        if next_batch is not None:
            self.load(next_batch, 1 - s)                         # its graph is built inside this replay
        self.graphs[s].replay()
        packing.invalidate_packs()      # the replay rewrote the parameters behind the version counters
        self._ready = [False, False]
        self._ready[1 - s] = next_batch is not None
        self._ready_for = next_batch
        self._cur = 1 - s
        self.loss_vec, self.loss = self.loss_vecs[s], self.loss_vecs[s][4]
        return self.loss
