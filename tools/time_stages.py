"""Developer script (gpurun): per-stage device times of one un-graphed inference step (CUDA events between the stages), as
bench.py's stages_ms, for the product library or an A/B build (tools/build_ab.sh <name>): python tools/time_stages.py [name] [B] [N]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_equi_graph_registration_b200 import _lib
if len(sys.argv) > 1 and sys.argv[1]:
    _lib.LIB_PATH = os.path.join(ROOT, "build", sys.argv[1], "libegspr_b200.so")
import se3_equi_graph_registration_b200 as P
import bench
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N = int(sys.argv[3]) if len(sys.argv) > 3 else 2048
model = P.build_model(bench.CKPT, device="cuda:0")
data = P.synthetic.make_batch(5, B, n=N)
eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=False)
eng.load(*[data[k] for k in ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")])
acc, reps = {}, 12
for rep in range(reps + 2):
    eng.stage_events = []
    eng._bind_inputs(0)
    eng.run()
    torch.cuda.synchronize()
    if rep >= 2:
        ev = eng.stage_events
        for (_, e_prev), (name, e_cur) in zip(ev[:-1], ev[1:]):
            acc[name] = acc.get(name, 0.0) + e_prev.elapsed_time(e_cur) / reps
print(sys.argv[1] if len(sys.argv) > 1 else "product", {k: round(v * 1e3, 1) for k, v in acc.items()}, "total %.1f us" % (sum(acc.values()) * 1e3))
