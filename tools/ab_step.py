"""Developer script (gpurun): resident-input step time of the engine (CUDA graph), for A/B runs with EGSPR_LIB_PATH."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
import bench
keys = ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")
model = P.build_model(bench.CKPT, device="cuda:0")
data = {k: v.cuda() for k, v in P.synthetic.make_batch(5, 64, n=2048).items()}
eng = P.RegistrationEngine(model, batch=64, n=2048, k=16, use_graph=True)
eng.impl = int(os.environ.get("EGSPR_IMPL", "0"))
for _ in range(5):
    eng.register(*[data[k] for k in keys])
torch.cuda.synchronize()
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50):
        eng.register(*[data[k] for k in keys])
    e1.record(); torch.cuda.synchronize()
    print(f"{os.environ.get('EGSPR_LIB_PATH', 'tree')}: {e0.elapsed_time(e1) / 50:.4f} ms/step  R checksum {float(eng.R.double().sum()):.9f}")
