"""Developer script (gpurun): randomized stress of the cell-grid k-NN (shell pruning, 64-bit keys) against the brute-force scan
kernel (same total order, no grid): clustered / planar / collinear / duplicate-heavy / lattice clouds, offsets and scales,
n from 3 to 20000, k from 1 to 32."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from se3_equi_graph_registration_b200 import ops
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = 0
cases = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 160):
    n = int(rng.choice([3, 17, 100, 333, 1024, 2048, 4096, 9000, 20000]))
    C = int(rng.choice([1, 2, 5]))
    k = int(rng.choice([1, 4, 8, 16, 16, 16, 24, 32]))
    kind = rng.choice(["uniform", "cluster", "plane", "line", "dups", "lattice", "shell", "aniso"])
    pts = rng.random((C, n, 3))
    if kind == "cluster":
        centres = rng.random((C, 6, 3))
        pts = centres[np.arange(C)[:, None], rng.integers(0, 6, (C, n))] + 0.01 * rng.standard_normal((C, n, 3))
    elif kind == "plane":
        pts[..., 2] = 0.5 + (1e-6 * rng.standard_normal((C, n)) if rng.random() < 0.5 else 0.0)
    elif kind == "line":
        pts[..., 1:] = 0.25
    elif kind == "dups":
        src = rng.integers(0, max(1, n // 7), (C, n))
        pts = np.take_along_axis(pts, src[..., None].repeat(3, -1), axis=1)
    elif kind == "lattice":
        m = int(np.ceil(n ** (1 / 3)))
        g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3)[:n]
        pts = np.broadcast_to(g, (C, n, 3)).astype(np.float64) / m          # exact ties everywhere
    elif kind == "shell":
        v = rng.standard_normal((C, n, 3)); pts = v / np.linalg.norm(v, axis=-1, keepdims=True)
    elif kind == "aniso":
        pts = pts * np.array([100.0, 100.0, 6.0])
    scale = float(rng.choice([1e-3, 1.0, 3.0, 100.0, 1e4]))
    off = rng.choice([0.0, -5.0, 1000.0]) * scale
    x = torch.from_numpy((pts * scale + off).astype(np.float32)).cuda().contiguous()
    a = ops.knn_build(x, k)
    b = ops.knn_build(x, k, brute_force=True)
    cases += 1
    if not bool((a == b).all()):
        bad += 1
        print("MISMATCH", kind, "n", n, "C", C, "k", k, "scale", scale, "off", off, "rows differing", int((a != b).any(-1).sum()))
print(f"{cases} cases, {bad} mismatches")
