"""Evaluation metrics of the reference's eval loop (tools/evaluation_metrics.py:14-43, F1 at
src/eval_egnn_metrics.py:1277) -- a thin host layer over the device kernel `egspr_pose_metrics` (fp64, one CTA per pair).
The reference's two functions are mirrored by name and return convention; the arithmetic lives in csrc/head.cu
(pose_metrics_kernel) and is checked against known answers produced by the reference file itself
(tests/golden/metrics_kat.pt).  CUDA only: there is no host fallback."""
import numpy as np
import torch

from . import ops


def _dev(device):
    return torch.device(device if device is not None else "cuda:0")


def evaluate_batch(R, t, gt_pose, src_pts, tgt_pts, tau=0.09, device=None):
    """Per-pair metrics for a batch -> dict of lists (rot_err deg, trans_err cm, recall, precision, f1), the quantities
    evl:1262-1281 collects.  Inputs: tensors or arrays, R [B,3,3], t [B,3], gt_pose [B,4,4], src_pts / tgt_pts [B,n,3]."""
    as_t = lambda v: (v if torch.is_tensor(v) else torch.as_tensor(np.asarray(v))).to(torch.float32)
    dev = R.device if torch.is_tensor(R) and R.is_cuda else _dev(device)
    m = ops.pose_metrics(as_t(R).to(dev), as_t(t).to(dev), as_t(gt_pose).to(dev), as_t(src_pts).to(dev), as_t(tgt_pts).to(dev), tau).cpu()
    return {"rot_err": m[:, 0].tolist(), "trans_err": m[:, 1].tolist(), "recall": m[:, 2].tolist(),
            "precision": m[:, 3].tolist(), "f1": m[:, 4].tolist()}


def _one(gt_pose, pred_pose, src_pts, tgt_pts, tau, device):
    pred = torch.as_tensor(np.asarray(pred_pose), dtype=torch.float32)
    n = 1 if src_pts is None else None
    src = torch.zeros(1, 1, 3) if src_pts is None else torch.as_tensor(np.asarray(src_pts), dtype=torch.float32)[None]
    tgt = torch.zeros(1, 1, 3) if tgt_pts is None else torch.as_tensor(np.asarray(tgt_pts), dtype=torch.float32)[None]
    out = evaluate_batch(pred[None, :3, :3], pred[None, :3, 3], torch.as_tensor(np.asarray(gt_pose), dtype=torch.float32)[None],
                         src, tgt, tau, device)
    return {k: v[0] for k, v in out.items()}


def calculate_pose_error(gt_pose, pred_pose, device=None):
    """-> (rotation error in degrees, translation error in CENTIMETRES), as tools/evaluation_metrics.py:14-24."""
    m = _one(gt_pose, pred_pose, None, None, 0.09, device)
    return m["rot_err"], m["trans_err"]


def registration_recall(gt_pose, pred_pose, src_pts, tgt_pts, tau=0.09, device=None):
    """-> (recall = sqrt(TP / N), precision = TP / N), as tools/evaluation_metrics.py:26-43."""
    m = _one(gt_pose, pred_pose, src_pts, tgt_pts, tau, device)
    return m["recall"], m["precision"]


def pose_matrix(R, t):
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, dtype=np.float64)
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T
