"""Developer script: run the engine a few times at the bench configuration (for ncu captures)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
B = int(os.environ.get("EGSPR_B", "64")); N = int(os.environ.get("EGSPR_N", "2048"))
impl = int(os.environ.get("EGSPR_IMPL", "3")); iters = int(os.environ.get("EGSPR_ITERS", "2"))
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device="cuda:0")
data = P.synthetic.make_batch(5, B, n=N)
eng = P.RegistrationEngine(model, batch=B, n=N, k=16, use_graph=False)
eng.impl = impl
for i in range(iters):
    eng.register(*[data[k] for k in ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")])
torch.cuda.synchronize()
print("done", float(eng.R.sum()))
