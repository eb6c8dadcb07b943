"""equi-gspr-b200: B200-native (sm_100a) registration hot path of Equi-GSPR.

k-NN graph -> stacked multi-head E_GCL/EGNN layers -> correspondence-weight head -> weighted
Kabsch / 3x3 SVD pose, as hand-written CUDA kernels behind a C ABI (include/egspr_b200.h),
exposed through nn.Modules that keep the reference's signatures and checkpoint layout.
See DESIGN.md.  No CPU fallback: ops raise on CPU tensors or when the library is missing.
"""
from . import synthetic, metrics  # noqa: F401
from . import _lib, packing, ops  # noqa: F401
from .modules import (E_GCL, EGNN, CrossAttentionPoseRegression, knn_graph, knn_graph_batch,  # noqa: F401
                      get_edges_batch, get_edges_from_idx, unsorted_segment_sum, unsorted_segment_mean, egnn_equi_loss, pose_loss, compute_losses,
                      save_checkpoint, load_checkpoint)
from .engine import RegistrationEngine, PipelinedEngine  # noqa: F401
from . import train  # noqa: F401


def build_model(checkpoint=None, device="cuda:0", n_layers=3, variant="eval", num_heads=4):
    """EGNN(32,32,32,in_edge_nf=1,n_layers=3) + CrossAttentionPoseRegression(hidden_nf=32), the
    configuration of the reference scripts (src/eval_egnn_metrics.py:1371-1375).  variant: 'eval' = the class of the
    evaluation script (default here: the inference engine), 'train' = the class of the training scripts.  num_heads: 4 for
    the shipped checkpoints; any divisor of 32 for models trained with another head count."""
    egnn = EGNN(32, 32, 32, in_edge_nf=1, device=device, n_layers=n_layers, num_heads=num_heads)
    model = CrossAttentionPoseRegression(egnn, num_nodes=2048, hidden_nf=32, device=device, variant=variant).to(device)
    if checkpoint is not None:
        load_checkpoint(checkpoint, None, egnn, model, device=device)
    model.eval()
    return model
