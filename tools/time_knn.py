"""Developer script (gpurun): k-NN graph build (grid build + query) and CSR build times at the bench shape, L2 flushed between
launches; ids checked against the brute-force scan kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import se3_equi_graph_registration_b200 as P
from se3_equi_graph_registration_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
d = P.synthetic.make_batch(5, B, n=N)
x = torch.cat([d["src_pts"], d["tgt_pts"]]).cuda().contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def med(fn, reps=15):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2] * 1e3


nbr = ops.knn_build(x, 16)
ref = ops.knn_build(x, 16, brute_force=True)
print(f"{2 * B} clouds x {N}: k-NN {med(lambda: ops.knn_build(x, 16)):.1f} us, CSR {med(lambda: ops.csr_from_nbr(nbr)):.1f} us, "
      f"ids == brute force: {bool((nbr == ref).all())}")
