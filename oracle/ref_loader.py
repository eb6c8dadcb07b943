"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Loads the reference's own model classes from /root/reference WITHOUT copying them:
the three runnable scripts cannot be imported (PyG / torch_cluster missing, hard
`.cuda()` calls, argparse at import), but the model definitions are pure torch.
We `ast`-parse the script, keep only the top-level class / function definitions we
name, and `exec` them in a namespace that provides torch / nn / F.

Used only in the build container (where /root/reference exists) by
`tests/golden/make_golden.py` to generate the committed golden vectors and by the
CPU tests that pin `oracle/egnn_oracle.py` against the real reference when it is
present.  Nothing on the GPU box reads /root/reference.

Reference symbols pulled (path:line relative to /root/reference):
  src/eval_egnn_metrics.py   : compute_so3_matrix, compute_edge_features, E_GCL, EGNN,
                               unsorted_segment_sum, egnn_equi_loss,
                               CrossAttentionPoseRegression (eval variant, :594-827)
  src/3dmatch_train_egnn_with_batch.py : same + train-variant head (:585-796),
                               pose_loss (:896-962), rotation_matrix_to_quaternion_batch
"""
import ast
import contextlib
import io
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("EGSPR_REFERENCE_ROOT", "/root/reference")

_WANTED = {
    "compute_so3_matrix", "compute_edge_features", "E_GCL", "EGNN",
    "unsorted_segment_sum", "unsorted_segment_mean", "egnn_equi_loss", "pose_loss",
    "CrossAttentionPoseRegression", "rotation_matrix_to_quaternion_batch",
    "compute_losses", "get_edges_from_idx",
}


def reference_available():
    return os.path.isfile(os.path.join(REF_ROOT, "src", "eval_egnn_metrics.py"))


def _load_defs(relpath):
    path = os.path.join(REF_ROOT, relpath)
    with open(path, "r") as f:
        tree = ast.parse(f.read(), filename=path)
    body = [n for n in tree.body
            if isinstance(n, (ast.ClassDef, ast.FunctionDef)) and n.name in _WANTED]
    mod = ast.Module(body=body, type_ignores=[])
    ns = {"torch": torch, "nn": nn, "F": F, "__name__": "reference_extract:" + relpath}
    exec(compile(mod, path, "exec"), ns)
    return ns


_cache = {}


def load(variant="eval"):
    """variant: 'eval' -> src/eval_egnn_metrics.py, 'train' -> src/3dmatch_train_egnn_with_batch.py"""
    rel = {"eval": "src/eval_egnn_metrics.py",
           "train": "src/3dmatch_train_egnn_with_batch.py",
           "kitti": "src/kitti_train_egnn_with_batch.py"}[variant]
    if rel not in _cache:
        _cache[rel] = _load_defs(rel)
    return _cache[rel]


def build_reference_model(variant="eval", checkpoint="checkpoints/checkpoint-3dmatch.pth",
                          num_heads=4, n_layers=3, dtype=torch.float32):
    """Instantiate the reference EGNN + CrossAttentionPoseRegression on CPU with the shipped
    checkpoint.  The only patch is num_heads=4 (EGNN never forwards num_heads to
    E_GCL, but the checkpoint was trained with 4 heads) and device='cpu'."""
    ns = load(variant)
    E_GCL = ns["E_GCL"]
    orig_init = E_GCL.__init__

    def patched_init(self, *a, **kw):
        kw["num_heads"] = num_heads
        kw["device"] = "cpu"
        orig_init(self, *a, **kw)

    E_GCL.__init__ = patched_init
    try:
        egnn = ns["EGNN"](32, 32, 32, in_edge_nf=1, device="cpu", n_layers=n_layers)
        head = ns["CrossAttentionPoseRegression"](egnn, num_nodes=2048, hidden_nf=32, device="cpu")
    finally:
        E_GCL.__init__ = orig_init
    if checkpoint is not None:
        ck = torch.load(os.path.join(REF_ROOT, checkpoint), map_location="cpu", weights_only=True)
        egnn.load_state_dict(ck["egnn_state_dict"], strict=True)
        head.load_state_dict(ck["cross_attention_state_dict"], strict=True)
    if dtype != torch.float32:
        head = head.to(dtype)
    head.eval()
    return ns, egnn, head


def run_quiet(fn, *a, **kw):
    """The eval-variant forward prints debug spam (evl:723-725, 780-781); swallow it."""
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **kw)
