"""Developer / report script (gpurun): throughput of the hot path away from the headline shape --
BASELINE configs[2] (KITTI-shaped, 2048-8192 pts) and configs[4] (16k-128k-point clouds, k = 16 / 32),
with the k-NN ids of every configuration checked bit-exact against the brute-force scan kernel.
Prints one JSON line per configuration."""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
import bench

CONFIGS = [  # (label, shape, points, k, pairs)
    ("3dmatch", "3dmatch", 2048, 16, 64),
    ("kitti", "kitti", 2048, 16, 64), ("kitti", "kitti", 4096, 16, 32), ("kitti", "kitti", 8192, 16, 16),
    ("sweep", "cube", 16384, 16, 8), ("sweep", "cube", 32768, 16, 4), ("sweep", "cube", 65536, 16, 2), ("sweep", "cube", 131072, 16, 1),
    ("sweep", "cube", 16384, 32, 8), ("sweep", "cube", 131072, 32, 1),
]
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")
model = P.build_model(bench.CKPT, device="cuda:0")


def make(shape, n, B, seed):
    if shape != "cube":
        return P.synthetic.make_batch(seed, B, n=n, shape=shape)
    d = P.synthetic.make_batch(seed, B, n=2048)          # features / labels recipe, then constant-density clouds
    g = torch.Generator().manual_seed(seed)
    out = {}
    pts = torch.stack([torch.from_numpy(P.synthetic.make_cloud(seed * 977 + i, n)) for i in range(B)])
    R = d["gt_pose"][:, :3, :3]; t = d["gt_pose"][:, :3, 3]
    out["src_pts"] = pts
    out["tgt_pts"] = pts @ R.transpose(1, 2) + t[:, None, :] + 0.01 * torch.randn(B, n, 3, generator=g)
    f = torch.nn.functional.normalize(torch.randn(B, n, 32, generator=g), dim=-1)
    out["src_feat"] = f
    out["tgt_feat"] = torch.nn.functional.normalize(f + 0.2 * torch.randn(B, n, 32, generator=g), dim=-1)
    out["labels"] = torch.ones(B, n)
    out["gt_pose"] = d["gt_pose"]
    return out


for label, shape, n, k, B in CONFIGS:
    data = make(shape, n, B, 7)
    eng = P.RegistrationEngine(model, batch=B, n=n, k=k, use_graph=True)
    eng.register(*[data[kk] for kk in keys])
    torch.cuda.synchronize()
    nbr_brute = ops.knn_build(eng.x, k, brute_force=True)
    exact = bool(torch.equal(nbr_brute, eng.nbr))
    finite = bool(torch.isfinite(eng.R).all() and torch.isfinite(eng.h_out).all())
    dev = {kk: data[kk].cuda() for kk in keys}
    for _ in range(3):
        eng.register(*[dev[kk] for kk in keys])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        eng.register(*[dev[kk] for kk in keys])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    # stage times
    def tm(fn, r=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(r): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / r
    knn_ms = tm(lambda: ops.knn_build(eng.x, k))
    eng.use_graph = False
    edge_ms = bench.eng_layer_time(eng, reps=5)
    # training step (BASELINE configs[3]) at this shape: a quarter of the pairs, whole step as one CUDA graph
    train_ms = None
    if shape != "cube" and n <= 8192:
        Bt = max(1, B // 4)
        tm_model = P.build_model(bench.CKPT, device="cuda:0", variant="train")
        with torch.no_grad():
            tm_model.egnn.embedding_out.weight.mul_(bench.TRAIN_TEMPER); tm_model.egnn.embedding_out.bias.mul_(bench.TRAIN_TEMPER)
        tkeys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "corr", "labels", "gt_pose")
        tb = tuple(data[kk][:Bt].cuda() for kk in tkeys)
        step = P.train.GraphedTrainStep(tm_model, torch.optim.Adam(tm_model.parameters(), lr=1e-5, capturable=True, fused=True), tb, k=k)
        train_ms = tm(lambda: step(tb), r=10)
        train_pairs = Bt
        del step, tm_model
    E = 2 * B * n * k
    alg = E * (2 * 32 * 4 + 2 * 12 + 4) + 2 * B * n * (32 * 4 + 12)
    print(json.dumps({"config": label, "shape": shape, "points": n, "k": k, "pairs": B, "ms_per_step": round(ms, 3),
                      "pairs_per_s": round(B / ms * 1e3, 1), "points_per_s": round(2 * B * n / ms * 1e3),
                      "knn_ms": round(knn_ms, 3), "edge_kernel_ms": round(edge_ms, 3),
                      "edge_alg_GBs": round(alg / edge_ms / 1e6, 1), "knn_ids_equal_brute_force": exact, "finite": finite,
                      "train_pairs": None if train_ms is None else train_pairs,
                      "train_ms_per_step": None if train_ms is None else round(train_ms, 3),
                      "train_pairs_per_s": None if train_ms is None else round(train_pairs / train_ms * 1e3, 1)}), flush=True)
    del eng
    torch.cuda.empty_cache()
