"""Seeded synthetic registration pairs shaped like the reference's dataset samples.

No dataset is available offline, so bench.py, smoke() and the tests all consume these.  The
shapes and distributions follow the reference's loaders (paths under /root/reference):
  datasets/ThreeDMatch.py:225-409  -> fixed 2048 correspondence-paired points per sample,
      7-tuple (corr, labels, src_pts, tgt_pts, src_feat, tgt_feat, gt_pose); points are sampled
      WITH replacement (:319,:329) so exact duplicate points are normal.
  data_preprocess/3DMatch_Feature.py:63,204 -> inlier label threshold 0.10 m
  data_preprocess/process_kitti.py:175      -> KITTI label threshold 0.60 m
  README.md:40 -> FCGF descriptors are unit-norm, 32-d.
"""
import numpy as np
import torch


def random_rotation(rng):
    q, r = np.linalg.qr(rng.standard_normal((3, 3)))
    q = q * np.sign(np.diag(r))[None, :]
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    return q


_SHAPES = {
    # extent (m), noise sigma range (m)
    "3dmatch": ((3.0, 3.0, 3.0), (0.005, 0.03)),
    "kitti": ((100.0, 100.0, 6.0), (0.05, 0.3)),
}


def make_pair(seed, n=2048, feat_dim=32, shape="3dmatch", inlier_ratio=None, dup_frac=0.0):
    """One synthetic pair.  Returns dict of float32 numpy arrays:
    src_pts [n,3], tgt_pts [n,3], src_feat [n,F], tgt_feat [n,F], labels [n], gt_pose [4,4], corr [n,2]."""
    rng = np.random.default_rng(seed)
    extent, (s_lo, s_hi) = _SHAPES[shape]
    extent = np.asarray(extent)
    src = rng.random((n, 3)) * extent
    R = random_rotation(rng)
    t = rng.random(3) - 0.5
    ratio = inlier_ratio if inlier_ratio is not None else rng.uniform(0.6, 0.9)
    labels = (rng.random(n) < ratio)
    sigma = rng.uniform(s_lo, s_hi)
    tgt = src @ R.T + t + rng.standard_normal((n, 3)) * sigma
    centre = (extent / 2) @ R.T + t
    outl = (rng.random((n, 3)) - 0.5) * extent + centre
    tgt = np.where(labels[:, None], tgt, outl)
    fs = rng.standard_normal((n, feat_dim))
    fs /= np.linalg.norm(fs, axis=1, keepdims=True)
    ft_in = fs + rng.uniform(0.1, 0.3) * rng.standard_normal((n, feat_dim))
    ft_out = rng.standard_normal((n, feat_dim))
    ft = np.where(labels[:, None], ft_in, ft_out)
    ft /= np.linalg.norm(ft, axis=1, keepdims=True)
    if dup_frac > 0:
        # sampling with replacement / many-to-one matches: whole correspondences repeat
        ndup = int(n * dup_frac)
        dst = rng.choice(n, ndup, replace=False)
        srcidx = rng.integers(0, n, ndup)
        for a in (src, tgt, fs, ft):
            a[dst] = a[srcidx]
        labels[dst] = labels[srcidx]
    pose = np.eye(4)
    pose[:3, :3] = R
    pose[:3, 3] = t
    f32 = np.float32
    return {
        "src_pts": src.astype(f32), "tgt_pts": tgt.astype(f32),
        "src_feat": fs.astype(f32), "tgt_feat": ft.astype(f32),
        "labels": labels.astype(f32), "gt_pose": pose.astype(f32),
        "corr": np.stack([np.arange(n), np.arange(n)], 1).astype(f32),
    }


def make_batch(seed, batch, n=2048, feat_dim=32, shape="3dmatch", dup_frac=0.0, pin=False):
    """A default-collate style batch (datasets/dataloader.py) of `batch` pairs as CPU tensors:
    corr [B,n,2], labels [B,n], src_pts [B,n,3], tgt_pts [B,n,3], src_feat, tgt_feat [B,n,F], gt_pose [B,4,4]."""
    items = [make_pair(seed * 100003 + i, n, feat_dim, shape, dup_frac=dup_frac) for i in range(batch)]
    out = {}
    for key in items[0]:
        ten = torch.from_numpy(np.stack([it[key] for it in items]))
        out[key] = ten.pin_memory() if pin else ten
    return out


def make_cloud(seed, n, density=75.85):
    """Scaling-sweep cloud (BASELINE config 5): uniform cube at constant density
    (default = 2048 points in a 3 m cube)."""
    rng = np.random.default_rng(seed)
    side = (n / density) ** (1.0 / 3.0)
    return (rng.random((n, 3)) * side).astype(np.float32)


def make_sweep_batch(seed, batch, n, feat_dim=32, pin=False):
    """Scaling-sweep batch (BASELINE config 5): constant-density uniform cubes of n points (make_cloud), targets = the
    rigidly moved sources + 1 cm noise, all points inliers, unit-norm features."""
    base = make_batch(seed, batch, n=16, feat_dim=feat_dim)                 # poses only
    g = torch.Generator().manual_seed(seed)
    pts = torch.stack([torch.from_numpy(make_cloud(seed * 977 + i, n)) for i in range(batch)])
    R, t = base["gt_pose"][:, :3, :3], base["gt_pose"][:, :3, 3]
    f = torch.nn.functional.normalize(torch.randn(batch, n, feat_dim, generator=g), dim=-1)
    out = {"src_pts": pts, "tgt_pts": pts @ R.transpose(1, 2) + t[:, None, :] + 0.01 * torch.randn(batch, n, 3, generator=g),
           "src_feat": f, "tgt_feat": torch.nn.functional.normalize(f + 0.2 * torch.randn(batch, n, feat_dim, generator=g), dim=-1),
           "labels": torch.ones(batch, n), "gt_pose": base["gt_pose"],
           "corr": torch.arange(n, dtype=torch.float32)[None, :, None].expand(batch, n, 2).contiguous()}
    return {k: (v.pin_memory() if pin else v) for k, v in out.items()}
