"""RegistrationEngine: the whole inference hot path for a batch of pairs as one launch sequence.

What the reference's eval loop does per batch (src/eval_egnn_metrics.py:1122-1250):
  .to(device) x5  ->  2*B knn_graph calls  ->  2*B get_edges_batch  ->  model(...)  ->  .cpu()
here: one H2D copy per input into persistent device buffers laid out clouds-major
([2B,N,*]: clouds 0..B-1 = sources, B..2B-1 = targets), then
  k-NN (1 launch, all 2B clouds) -> CSR transpose (memset + 4 launches) -> embed (1) ->
  3 fused E_GCL layers (3) -> eval head + Kabsch (1)
on one stream, optionally replayed as a CUDA graph; the int64 [B,2,E] edge tensors of the
reference API are never materialised on this path.
"""
import ctypes

import torch

from . import _lib, ops

H = 32


class RegistrationEngine:
    def __init__(self, model, batch, n=2048, k=16, device=None, use_graph=True):
        """model: modules.CrossAttentionPoseRegression (its parameters stay the source of truth)."""
        self.model = model
        self.B, self.N, self.k = int(batch), int(n), int(k)
        self.device = torch.device(device if device is not None else next(model.parameters()).device)
        if self.device.type != "cuda":
            raise RuntimeError("RegistrationEngine needs a CUDA device: the egspr_b200 hot path has no CPU fallback")
        if self.N < self.k:
            raise ValueError("need at least k points per cloud")
        self.use_graph = use_graph
        B, N, dev = self.B, self.N, self.device
        C = 2 * B
        G, E = C * N, C * N * self.k
        f32, i32 = torch.float32, torch.int32
        z = lambda *s, dt=f32: torch.empty(s, dtype=dt, device=dev)
        # inputs (clouds-major), two sets: set 0 is the one register()/load()/run() use; submit() alternates
        # between them so the next batch's host->device copy overlaps the current batch's kernels
        self._in = [dict(feat=z(C, N, H), x=z(C, N, 3), labels=torch.zeros(B, N, dtype=f32, device=dev),
                         gt_pose=torch.eye(4, dtype=f32, device=dev).repeat(B, 1, 1).contiguous()) for _ in range(2)]
        self._set = 0
        self._bind_inputs(0)
        self._copy_stream = None
        self._set_free = [None, None]        # event: the kernels that read input set s have finished
        self._host_out = None
        self._tickets = 0
        # graph
        self.nbr = z(C, N, self.k, dt=i32)
        self.knn_ws_bytes = _lib.lib().egspr_knn_workspace_bytes(C, N)
        self.knn_ws = torch.empty(self.knn_ws_bytes, dtype=torch.uint8, device=dev)
        self.knn_brute_force = False
        self.csr_ptr = z(G + 1, dt=i32); self.csr_row = z(E, dt=i32); self.csr_col = z(E, dt=i32); self.csr_eid = z(E, dt=i32)
        self.ws_bytes = _lib.lib().egspr_csr_workspace_bytes(G, E)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=dev)
        self.err = torch.zeros(1, dtype=i32, device=dev)
        # layer state (ping-pong)
        self.h = [z(G, H), z(G, H)]; self.x4 = [z(G, 4), z(G, 4)]
        self.P = [z(G, H), z(G, H)]; self.Q = [z(G, H), z(G, H)]
        self.x_out = z(C, N, 3)
        self.agg_ws = z(G, H)
        # outputs
        self.R = z(B, 3, 3); self.t = z(B, 3); self.Hm = z(B, 3, 3); self.w = z(B, N); self.loss_parts = z(B, 2)
        self.head_ws_bytes = _lib.lib().egspr_head_eval_workspace_bytes(B)
        self.head_ws = torch.empty(max(self.head_ws_bytes, 16), dtype=torch.uint8, device=dev)
        self.h_out = None
        self._graph = [None, None]
        self._graph_key = [None, None]
        self.impl = 0
        self.launches_per_step = 0
        self.stage_events = None     # set to [] before an un-graphed run(): receives (stage name, CUDA event after the stage)

    def _bind_inputs(self, s):
        self._set = s
        d = self._in[s]
        self.feat, self.x, self.labels, self.gt_pose = d["feat"], d["x"], d["labels"], d["gt_pose"]

    # ---- input staging -----------------------------------------------------------------------
    def load(self, src_feat, src_pts, tgt_feat, tgt_pts, labels=None, gt_pose=None):
        """Copy one batch (host pinned or device tensors, [B,N,*]) into the persistent buffers."""
        B = self.B
        self.feat[:B].copy_(src_feat, non_blocking=True); self.feat[B:].copy_(tgt_feat, non_blocking=True)
        self.x[:B].copy_(src_pts, non_blocking=True); self.x[B:].copy_(tgt_pts, non_blocking=True)
        if labels is not None:
            self.labels.copy_(labels, non_blocking=True)
        if gt_pose is not None:
            self.gt_pose.copy_(gt_pose, non_blocking=True)

    # ---- launch sequence ---------------------------------------------------------------------
    def _enqueue(self):
        lib = _lib.lib()
        p = ops._ptr
        B, N, k = self.B, self.N, self.k
        C = 2 * B
        G = C * N
        self.model.egnn.check_impl(self.impl)
        layers, pin, pout = self.model.egnn.packs()
        head = self.model._pack_head.get()
        st = ops._stream()
        n_launch = 0

        def mark(name):
            if self.stage_events is not None:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                self.stage_events.append((name, ev))

        mark("start")
        if self.knn_brute_force:
            _lib.check(lib.egspr_knn_build(p(self.x), C, N, k, p(self.nbr), None, 0, st), "egspr_knn_build"); n_launch += 1
        else:
            _lib.check(lib.egspr_knn_build(p(self.x), C, N, k, p(self.nbr), p(self.knn_ws), self.knn_ws_bytes, st),
                       "egspr_knn_build")
            n_launch += 5 if (N >= 8192 and C <= 8) else 2      # split grid build (bbox, count, scan, scatter) + query
        mark("knn")
        _lib.check(lib.egspr_csr_from_nbr(p(self.nbr), C, N, k, p(self.csr_ptr), p(self.csr_row), p(self.csr_col),
                                          p(self.csr_eid), p(self.ws), self.ws_bytes, p(self.err), st), "egspr_csr_from_nbr")
        # csr: one fused launch for small clouds (csr.cu: shared-memory build), else memset + count/scan/fill/emit
        n_launch += 1 if (4 * (2 * N + 1 + N * k) <= 200 * 1024 and C >= 16) else 4
        mark("csr")
        _lib.check(lib.egspr_node_embed(p(self.feat), p(self.x), G, p(pin), p(layers[0]), p(self.h[0]), p(self.x4[0]),
                                        p(self.P[0]), p(self.Q[0]), st), "egspr_node_embed"); n_launch += 1
        mark("embed")
        cur = 0
        L = len(layers)
        for i in range(L):
            last = i == L - 1
            nxt = 1 - cur
            _lib.check(lib.egspr_egcl_forward(
                p(self.h[cur]), p(self.x4[cur]), p(self.P[cur]), p(self.Q[cur]),
                p(self.csr_ptr), p(self.csr_row), p(self.csr_col), p(self.csr_eid), None, 1.0,
                G, N * k, N, p(layers[i]), None if last else p(layers[i + 1]), p(pout) if last else None,
                p(self.h[nxt]), p(self.x4[nxt]), p(self.x_out) if last else None,
                None if last else p(self.P[nxt]), None if last else p(self.Q[nxt]), p(self.agg_ws), int(self.impl), st),
                "egspr_egcl_forward")
            n_launch += 2 if self.impl in (0, 3, 4, 5) else 1
            mark("layer%d" % i)
            cur = nxt
        self.h_out = self.h[cur].view(C, N, H)
        ho, xo = self.h_out, self.x_out
        _lib.check(lib.egspr_head_eval_ws(p(self.feat[:B]), p(self.feat[B:]), p(self.x[:B]), p(self.x[B:]),
                                          p(ho[:B]), p(ho[B:]), p(xo[:B]), p(xo[B:]), p(self.labels), p(self.gt_pose),
                                          p(head), B, N, int(self.model.top_k), p(self.w), p(self.R), p(self.t), p(self.Hm),
                                          p(self.loss_parts), p(self.head_ws), self.head_ws_bytes, st), "egspr_head_eval_ws")
        n_launch += 2 if (B < 2 * 148 and N >= 2048) else 1
        mark("head")
        self.launches_per_step = n_launch

    def run(self):
        """Enqueue the hot path on the current stream (CUDA-graph replay when enabled)."""
        with torch.cuda.device(self.device):
            if not self.use_graph:
                self._enqueue()
                return
            packs = self.model.egnn.packs()
            key = tuple(t.data_ptr() for t in packs[0]) + (packs[1].data_ptr(), packs[2].data_ptr(),
                                                           self.model._pack_head.get().data_ptr(), self.impl)
            s = self._set
            if self._graph[s] is None or key != self._graph_key[s]:
                self._enqueue()                      # warm-up outside capture (function attributes, lazy init)
                torch.cuda.current_stream().synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._enqueue()
                self._graph[s], self._graph_key[s] = g, key
            self._graph[s].replay()

    def register(self, src_feat, src_pts, tgt_feat, tgt_pts, labels=None, gt_pose=None):
        """Full step: stage inputs, run, return (R [B,3,3], t [B,3]) device tensors."""
        self._bind_inputs(0)
        self.load(src_feat, src_pts, tgt_feat, tgt_pts, labels, gt_pose)
        self.run()
        return self.R, self.t

    # ---- pipelined host-to-host path ------------------------------------------------------------
    def submit(self, src_feat, src_pts, tgt_feat, tgt_pts, labels=None, gt_pose=None):
        """Enqueue one batch given as HOST (pinned) tensors: H2D on a copy stream into the free input set,
        the hot path on the current stream, (R, t) back to pinned host buffers.  Returns a ticket for
        collect().  Up to two batches are in flight, so batch i+1's upload overlaps batch i's kernels
        (the reference's loop uploads, computes and downloads strictly in turn, evl:1126-1250)."""
        with torch.cuda.device(self.device):
            if self._copy_stream is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
                self._copy_stream2 = torch.cuda.Stream(device=self.device)
                self._host_out = [(torch.empty((self.B, 3, 3), dtype=torch.float32).pin_memory(),
                                   torch.empty((self.B, 3), dtype=torch.float32).pin_memory()) for _ in range(2)]
            s = self._tickets & 1
            self._tickets += 1
            cur = torch.cuda.current_stream()
            self._bind_inputs(s)
            # source and target halves on two copy streams: one stream moves a 17 MB tensor at ~25-30 GB/s on this
            # platform, two concurrent copies reach ~37 GB/s of the PCIe 5 x16 link (tools/h2d_bandwidth.py)
            B = self.B
            halves = ((self._copy_stream, (self.feat[:B], src_feat), (self.x[:B], src_pts), (self.labels, labels)),
                      (self._copy_stream2, (self.feat[B:], tgt_feat), (self.x[B:], tgt_pts), (self.gt_pose, gt_pose)))
            for cs, *copies in halves:
                if self._set_free[s] is not None:
                    cs.wait_event(self._set_free[s])      # the kernels that read this input set two batches ago are done
                with torch.cuda.stream(cs):
                    for dst, src in copies:
                        if src is not None:
                            dst.copy_(src, non_blocking=True)
                    uploaded = torch.cuda.Event()
                    uploaded.record(cs)
                cur.wait_event(uploaded)
            self.run()
            Rh, th = self._host_out[s]
            Rh.copy_(self.R, non_blocking=True)
            th.copy_(self.t, non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            self._set_free[s] = done
            return (s, done)

    def collect(self, ticket):
        """Wait for a submitted batch; returns its (R [B,3,3], t [B,3]) pinned HOST tensors (valid until the
        second-next submit())."""
        s, done = ticket
        done.synchronize()
        return self._host_out[s]

    def metrics(self, tau=0.09):
        """Per-pair (rotation error deg, translation error cm, recall, precision, F1) of the last batch against
        its gt_pose, on the device: float64 [B,5] (tools/evaluation_metrics.py:14-43, evl:1277)."""
        B = self.B
        return ops.pose_metrics(self.R, self.t, self.gt_pose, self.x[:B], self.x[B:], tau)

    def outputs(self):
        B = self.B
        return {"R": self.R, "t": self.t, "w": self.w, "H": self.Hm,
                "h_src": self.h_out[:B], "h_tgt": self.h_out[B:], "x_src": self.x_out[:B], "x_tgt": self.x_out[B:],
                "equi_loss": self.loss_parts.sum() / (B * self.N), "nbr": self.nbr}


class PipelinedEngine:
    """`lanes` RegistrationEngines, each with its own buffers, CUDA graph and stream: batch i runs on lane i % lanes.

    One batch's launch sequence has wide kernels (k-NN query, the edge / node kernels: >= one CTA per SM) and narrow ones
    (eval head: ONE CTA per pair = 64 of 148 SMs for 117 us; CSR build; k-NN grid build: one CTA per cloud).  With two
    batches in flight the narrow kernels of one run beside the wide kernels of the other, and uploads overlap compute
    as in RegistrationEngine.submit().  The reference's loop processes its batches strictly one after another
    (src/eval_egnn_metrics.py:1122-1250).  Results are identical to a single engine's (same kernels, same order per batch)."""

    def __init__(self, model, batch, n=2048, k=16, device=None, lanes=2, use_graph=True):
        self.engines = [RegistrationEngine(model, batch, n=n, k=k, device=device, use_graph=use_graph) for _ in range(int(lanes))]
        self.device = self.engines[0].device
        self.streams = [torch.cuda.Stream(device=self.device) for _ in self.engines]
        self._count = 0
        self.active_lanes = len(self.engines)      # batches rotate over the first `active_lanes` lanes

    @property
    def impl(self):
        return self.engines[0].impl

    @impl.setter
    def impl(self, v):
        for e in self.engines:
            e.impl = v

    @property
    def launches_per_step(self):
        return self.engines[0].launches_per_step

    def _next(self):
        lane = self._count % max(1, min(int(self.active_lanes), len(self.engines)))
        self._count += 1
        return lane

    def register(self, src_feat, src_pts, tgt_feat, tgt_pts, labels=None, gt_pose=None):
        """Enqueue one batch (device or pinned-host tensors) on the next lane -> ticket; result(ticket) waits for it."""
        lane = self._next()
        eng, st = self.engines[lane], self.streams[lane]
        st.wait_stream(torch.cuda.current_stream(self.device))          # the inputs may have been produced there
        with torch.cuda.stream(st):
            eng.register(src_feat, src_pts, tgt_feat, tgt_pts, labels, gt_pose)
            done = torch.cuda.Event()
            done.record(st)
        return (lane, done)

    def result(self, ticket):
        """(R [B,3,3], t [B,3]) device tensors of a registered batch (valid until the lane is used again)."""
        lane, done = ticket
        done.synchronize()
        return self.engines[lane].R, self.engines[lane].t

    def submit(self, src_feat, src_pts, tgt_feat, tgt_pts, labels=None, gt_pose=None):
        """Host-to-host form (pinned host tensors in, pinned host poses out), see RegistrationEngine.submit()."""
        lane = self._next()
        with torch.cuda.stream(self.streams[lane]):
            return (lane, self.engines[lane].submit(src_feat, src_pts, tgt_feat, tgt_pts, labels, gt_pose))

    def collect(self, ticket):
        lane, inner = ticket
        return self.engines[lane].collect(inner)

    def join(self):
        """The current stream waits for every lane (use before recording an end-of-region event)."""
        cur = torch.cuda.current_stream(self.device)
        for st in self.streams:
            cur.wait_stream(st)

    def synchronize(self):
        for st in self.streams:
            st.synchronize()
