"""Developer script (gpurun): CUDA-event time of one E_GCL layer launch sequence per impl at the bench shape."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_equi_graph_registration_b200 import _lib
if len(sys.argv) > 1:        # A/B runs: a developer build of the library (build/<name>/libegspr_b200.so)
    _lib.LIB_PATH = os.path.join(ROOT, "build", sys.argv[1], "libegspr_b200.so")
import se3_equi_graph_registration_b200 as P
import bench
impls = [int(x) for x in os.environ.get("EGSPR_IMPLS", "1,3").split(",")]
B = int(os.environ.get("EGSPR_B", "64")); N = int(os.environ.get("EGSPR_N", "2048"))
model = P.build_model(bench.CKPT, device="cuda:0")
data = P.synthetic.make_batch(5, B, n=N)
eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=False)
eng.load(*[data[k] for k in ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")])
for impl in impls:
    eng.impl = impl
    ms = bench.eng_layer_time(eng, reps=20)
    eng.use_graph = False
    print(f"impl {impl}: layer (edge+node kernels) {ms*1e3:.1f} us")
