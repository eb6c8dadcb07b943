"""Developer script (gpurun): per-phase clock64() timeline of one group of the tcgen05 edge-backward kernel.
Builds nothing: needs build/timing/libegspr_b200.so compiled with -DEGSPR_XB_TIMING (tools/build_timing.sh)."""
import ctypes, os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_equi_graph_registration_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", sys.argv[1] if len(sys.argv) > 1 else "timing", "libegspr_b200.so")
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
dev = "cuda:0"
B = 16
model = P.build_model(os.path.join(ROOT, "tests", "golden", "checkpoint-3dmatch.pth"), device=dev, variant="train")
d = {k: v.to(dev) for k, v in P.synthetic.make_batch(3, B, n=2048).items()}
x = torch.cat([d["src_pts"], d["tgt_pts"]]); f = torch.cat([d["src_feat"], d["tgt_feat"]])
graph = ops.with_csc(ops.csr_from_nbr(ops.knn_build(x, 16)))
layers, pin, pout = model.egnn.packs()
h, xo, saved = ops.egnn_forward_saved(f, x, graph, layers, pin, pout)
dh, dx = torch.randn_like(h), torch.randn_like(xo)
for _ in range(3):
    ops.egnn_backward(saved, graph, layers, pin, pout, dh, dx)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    ops.egnn_backward(saved, graph, layers, pin, pout, dh, dx)
b.record(); torch.cuda.synchronize()
print(f"egnn backward (3 layers): {a.elapsed_time(b) / 5:.3f} ms")
buf = np.zeros(4 * 64 * 32, dtype=np.int64)
fn = ctypes.CDLL(_lib.LIB_PATH).egspr_debug_read_xb
fn.argtypes = [ctypes.c_void_p]
assert fn(buf.ctypes.data) == 0
t = buf.reshape(4, 64, 32)
names = ["top: cp.async issue, idx loads, geometry", "wait WG_c (prev tile) + geo tile write", "bar1 wait", "S1 issue + cp.async wait",
         "mbar S1", "S1 epilogue (P+Q, SiLU, stash)", "a1 tile write", "bar2 wait", "coords prefetch + mbar S2", "LayerNorm",
         "m tile write", "bar3 wait", "mbar S3", "S3 epilogue (SiLU, dc1, priv)", "dc1 tile write", "bar4 wait",
         "dagg/uh loads + mbar S4", "LN backward + priv", "wait WG_a", "du tile write", "bar5 wait", "stash loads + mbar S5",
         "dpre", "dpre tile write", "bar6 wait", "dpre store + mbar S6", "geometry backward + dxe store", "loop tail"]
for w in range(4):
    rows = []
    for i in range(3, 16):
        r = t[w, i]
        if r[27] == 0:
            break
        rows.append([r[1] - r[0]] + [r[j + 1] - r[j] for j in range(1, 27)] + [t[w, i + 1, 0] - r[27] if t[w, i + 1, 0] else 0])
    dd = np.array(rows, dtype=np.float64)
    print(f"warp {w}: {len(rows)} tiles, tile total {dd.sum(1).mean():.0f} cycles")
    for n, m, sdev in zip(names, dd.mean(0), dd.std(0)):
        print(f"    {n:44s} {m:8.0f} +- {sdev:6.0f}")
print("arrival skew at the barriers (cycles after the first warp), tiles 5..8:")
for i in range(5, 9):
    for name, slot in (("bar1", 2), ("bar2", 7), ("bar3", 11), ("bar4", 15), ("bar5", 20), ("bar6", 24)):
        arr = t[:, i, slot]; ex = t[:, i, slot + 1]
        print(f"  tile {i} {name}: arrive {[int(a - arr.min()) for a in arr]}  leave {[int(e - arr.min()) for e in ex]}")
