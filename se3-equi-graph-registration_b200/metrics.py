"""Evaluation metrics of the reference's eval loop (tools/evaluation_metrics.py:14-43), numpy
float64 like the reference (host side, negligible cost), plus batched helpers."""
import numpy as np


def calculate_pose_error(gt_pose, pred_pose):
    """(rotation error in degrees, translation error in CENTIMETRES)  tools/evaluation_metrics.py:14-24"""
    translation_error = np.linalg.norm(gt_pose[:3, 3] - pred_pose[:3, 3]) * 100
    rotation_diff = gt_pose[:3, :3].T @ pred_pose[:3, :3]
    rot_error = np.arccos(np.clip((np.trace(rotation_diff) - 1) / 2, -1.0, 1.0))
    return np.degrees(rot_error), translation_error


def registration_recall(gt_pose, pred_pose, src_pts, tgt_pts, tau=0.09):
    """(recall = sqrt(TP/N), precision = TP/N)  tools/evaluation_metrics.py:26-43"""
    src_transformed = (pred_pose[:3, :3] @ src_pts.T).T + pred_pose[:3, 3]
    distances = np.linalg.norm(src_transformed - tgt_pts, axis=1)
    true_positives = np.sum(distances < tau)
    recall = np.sqrt(true_positives / len(src_pts))
    precision = true_positives / len(src_transformed) if len(src_transformed) > 0 else 0.0
    return recall, precision


def f1_score(precision, recall):
    """src/eval_egnn_metrics.py:1277"""
    return 2 * (precision * recall) / (precision + recall + 1e-6)


def pose_matrix(R, t):
    T = np.eye(4)
    T[:3, :3] = np.asarray(R, dtype=np.float64)
    T[:3, 3] = np.asarray(t, dtype=np.float64)
    return T


def evaluate_batch(R, t, gt_pose, src_pts, tgt_pts):
    """Per-pair metrics for a batch (arrays on host).  Returns dict of lists, as evl:1262-1281 collects."""
    out = {"rot_err": [], "trans_err": [], "recall": [], "precision": [], "f1": []}
    for b in range(len(R)):
        T = pose_matrix(R[b], t[b])
        g = np.asarray(gt_pose[b], dtype=np.float64)
        re, te = calculate_pose_error(g, T)
        rec, prec = registration_recall(g, T, np.asarray(src_pts[b], np.float64), np.asarray(tgt_pts[b], np.float64))
        out["rot_err"].append(re); out["trans_err"].append(te); out["recall"].append(rec)
        out["precision"].append(prec); out["f1"].append(f1_score(prec, rec))
    return out
