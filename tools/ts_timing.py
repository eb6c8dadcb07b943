"""Developer script (gpurun): per-phase clock64() timeline of one group of the impl-5 edge kernel.
Needs build/timing/libegspr_b200.so (tools/build_timing.sh, -DEGSPR_TS_TIMING)."""
import ctypes, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_equi_graph_registration_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "timing", "libegspr_b200.so")
import se3_equi_graph_registration_b200 as P
import bench
model = P.build_model(bench.CKPT, device="cuda:0")
data = P.synthetic.make_batch(5, 64, n=2048)
eng = P.RegistrationEngine(model, batch=64, n=2048, k=16, use_graph=False)
eng.impl = 3
for _ in range(3):
    eng.register(*[data[k] for k in ("src_feat", "src_pts", "tgt_feat", "tgt_pts", "labels", "gt_pose")])
torch.cuda.synchronize()
buf = np.zeros(4 * 64 * 12, dtype=np.int64)
fn = ctypes.CDLL(_lib.LIB_PATH).egspr_debug_read_ts
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
t = buf.reshape(4, 64, 12)
names = ["top->bar1 arrive", "bar1 wait", "bar1->mbar1", "mbar1 wait", "mbar1->bar2 arrive", "bar2 wait", "bar2->mbar2 wait done",
         "mbar2->bar3 arrive", "bar3 wait", "bar3->segsum done", "mbar3 wait", "epilogue->next top"]
for w in range(4):
    d = []
    for i in range(5, 50):
        row = t[w, i]; nxt = t[w, i + 1, 0]
        seg = [row[j + 1] - row[j] for j in range(11)] + [nxt - row[11]]
        d.append(seg)
    d = np.array(d, dtype=np.float64)
    print(f"warp {w}: tile total {d.sum(1).mean():.0f} cycles")
    for n, m, s in zip(names, d.mean(0), d.std(0)):
        print(f"    {n:28s} {m:8.0f} +- {s:6.0f}")
