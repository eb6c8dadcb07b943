/* egspr_b200.h -- C ABI of the B200-native Equi-GSPR registration hot path.
 *
 * The reference (alexandor91/se3-equi-graph-registration) has NO native/FFI layer: its boundary
 * for this path is the Python nn.Module / function API of the three runnable scripts plus the
 * torch_cluster.knn_graph call.  Each entry point below names the reference interface it replaces
 * (path:line under the reference tree; 3dm = src/3dmatch_train_egnn_with_batch.py,
 * evl = src/eval_egnn_metrics.py).  The Python mirror of that API
 * (se3-equi-graph-registration_b200/modules.py) binds these with ctypes; INTEGRATION.md shows the
 * stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers + sizes only; every pointer is a DEVICE pointer unless the name ends in _host
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no entry point
 *     synchronises, allocates device memory, or throws
 *   - return value: 0 = ok, <0 = EGSPR_E_* (invalid argument / unsupported shape / launch failure)
 *   - outputs and workspaces are caller-allocated; *_workspace_bytes() tells how much
 *   - node / edge ids: the clouds of a batch are addressed as ONE graph with global node id
 *     g = cloud * n + i; edges never cross clouds
 */
#ifndef EGSPR_B200_H
#define EGSPR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGSPR_OK 0
#define EGSPR_E_INVALID (-1)      /* null pointer / non-positive size / k out of range        */
#define EGSPR_E_UNSUPPORTED (-2)  /* shape outside what the kernels are compiled for           */
#define EGSPR_E_WORKSPACE (-3)    /* workspace too small                                       */
#define EGSPR_E_LAUNCH (-4)       /* cudaGetLastError() != cudaSuccess after a launch          */

#define EGSPR_HIDDEN 32           /* hidden width the kernels are specialised for (3dm:1600)   */
#define EGSPR_MAX_K 32            /* neighbours per node supported by the warp-select k-NN     */

/* layer-pack layout (floats) produced by the host mirror from the live nn.Parameters; see
 * se3-equi-graph-registration_b200/packing.py and DESIGN.md "weight packs" */
#define EGSPR_LAYER_PACK_FLOATS 8128
#define EGSPR_EMBED_PACK_FLOATS 1056
#define EGSPR_HEAD_PACK_FLOATS 2640

int egspr_version(void);
const char *egspr_error_string(int code);

/* ---- a1: torch_cluster.knn_graph(x, k, loop=True) call sites 3dm:1005-1006, evl:1156-1157 ------
 * x [clouds][n][3] f32 -> nbr [clouds][n][k] i32: the k nearest points of the same cloud (self
 * included), ascending by (d2, index); d2 = fma(dz,dz,fma(dy,dy,dx*dx)) in fp32.  Slots that
 * cannot be filled (n < k) hold -1.  All clouds are processed by one launch sequence.
 * workspace (egspr_knn_workspace_bytes) selects the exact cell-grid search; workspace == NULL runs
 * the brute-force scan.  Both return identical ids (the order (d2, index) is total). */
size_t egspr_knn_workspace_bytes(int clouds, int n);
int egspr_knn_build(const float *x, int clouds, int n, int k, int32_t *nbr, void *workspace,
                    size_t workspace_bytes, void *stream);

/* ---- a2: get_edges_from_idx / get_edges_batch 3dm:372-403 -------------------------------------
 * nbr -> the reference's edge tensor edges[clouds][2][n*k] i64 (row = neighbour, col = centre,
 * cloud-local ids) for callers that want the torch_cluster layout. */
int egspr_nbr_to_edges(const int32_t *nbr, int clouds, int n, int k, int64_t *edges, void *stream);

/* ---- a9: unsorted_segment_sum 3dm:343-348, re-cast as a one-time graph transpose ----------------
 * Aggregation in E_GCL is over row = edge_index[0] (the NEIGHBOUR id), so every node
 * sums over its reverse-kNN set.  These build, once per graph, the row-major CSR the layer kernels
 * pull from: csr_ptr[G+1], and per sorted position p: csr_row[p] (global row id), csr_col[p] (global
 * col id), csr_eid[p] (original edge id inside its cloud); within a row, positions are in ascending
 * original edge order (= the order scatter_add_ visits them on CPU).  G = clouds*n.
 * err_flag (optional, device int) is set to 1 if any id is out of range (such edges are dropped). */
size_t egspr_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges);
int egspr_csr_from_nbr(const int32_t *nbr, int clouds, int n, int k, int32_t *csr_ptr,
                       int32_t *csr_row, int32_t *csr_col, int32_t *csr_eid, void *workspace,
                       size_t workspace_bytes, int32_t *err_flag, void *stream);
int egspr_csr_from_edges(const int64_t *edges, int clouds, int n, int64_t edges_per_cloud,
                         int32_t *csr_ptr, int32_t *csr_row, int32_t *csr_col, int32_t *csr_eid,
                         void *workspace, size_t workspace_bytes, int32_t *err_flag, void *stream);

/* stand-alone unsorted_segment_sum(data, segment_ids, num_segments) 3dm:343-348: build the CSR with
 * egspr_csr_from_edges (row = segment_ids, one "cloud"), then out[g][c] = sum over the segment in
 * ascending original order.  data [E][channels], out [num_segments][channels]. */
int egspr_segment_sum(const float *data, int channels, const int32_t *csr_ptr, const int32_t *csr_eid,
                      int64_t num_segments, float *out, void *stream);

/* ---- a11 (head of EGNN.forward 3dm:332): embedding_in, plus the per-node halves P,Q of layer 0's
 * first edge-MLP Linear (W1 [h_row|h_col|geo] = P[row] + Q[col] + Wgeo geo + b, bias folded in Q).
 * feat [G][32] -> h [G][32], P [G][32], Q [G][32]; x3 [G][3] -> x4 [G][4] (16-byte padded copy used
 * by the layer kernels).  embed_pack == NULL: h = feat (stand-alone E_GCL call). x4 == NULL: skip. */
int egspr_node_embed(const float *feat, const float *x3, int64_t num_nodes, const float *embed_pack,
                     const float *layer0_pack, float *h, float *x4, float *P, float *Q, void *stream);

/* ---- a3-a10: one E_GCL.forward (3dm:280-289) for every cloud of the batch in one launch ---------
 * Reads h,x4,P,Q of the layer input, writes h_out,x4_out (and x3_out [G][3] if not NULL) and either
 * the next layer's P_out,Q_out (next_pack != NULL) or, after the last layer, embedding_out(h_out)
 * into h_out (out_pack != NULL, 3dm:337).  Outputs must not alias inputs (other CTAs still gather
 * the layer input).  edge_attr: optional [clouds*edges_per_cloud] per-edge scalar indexed through
 * csr_eid (NULL = constant edge_attr_const, the reference's ones, 3dm:387; a layer built with
 * edges_in_d=0 has a zero edge_attr column in its pack).
 * impl: 0 = auto (the tensor-core path if agg_ws != NULL, else the fused CUDA-core kernel);
 *       1 / 2 = fused fp32 CUDA-core kernel with 64 / 256 nodes per block (edge, reduce and node phases
 *           in one launch).  Reads the per-head [4][8][8] layout of the second edge Linear: layers with 4
 *           heads only (packing.py fills that region with NaN for other head counts);
 *       3 = tensor-core path: edge kernel (first edge Linear's geometry block, the heads' second Linear
 *           and coord_mlp.0 on tcgen05 with the A operand handed over through tensor memory, 3xTF32 =
 *           fp32-level accuracy; streaming in-order segment sums) + node kernel (node MLP, residual, next
 *           layer's P/Q or embedding_out, also tcgen05); needs agg_ws [num_nodes][32] floats of scratch.
 *           The heads' second Linear is read as ONE block-diagonal 32 x 32 matrix (pack offset 7104), so
 *           this path and egspr_egcl_backward take any num_heads dividing 32 (E_GCL(num_heads=...), 3dm:186-207).
 *       4 = impl 3 with the edge kernel in reduced precision (BASELINE config 2's "looser bound" edge MLP):
 *           single-pass TF32 operands (10-bit mantissa, round to nearest) and SiLU through MUFU.TANH;
 *           segment sums, node kernel and everything else as in impl 3.
 *       5 = impl 3 with the edge kernel's MLPs in bf16 (BASELINE config 2 "bf16 edge MLP"): tcgen05.mma.kind::f16 with
 *           bf16 activations (A operand in tensor memory, two K elements per column) and bf16 weights, fp32
 *           accumulation; the 13 geometric inputs as two bf16 terms; SiLU through MUFU.TANH.
 *       3 | EGSPR_IMPL_EDGE_ONLY = the edge kernel of impl 3 alone (writes agg_ws, x4_out, x3_out; no node
 *           update) -- for benchmarks and profiling of that kernel. */
#define EGSPR_IMPL_EDGE_ONLY 0x100
int egspr_egcl_forward(const float *h, const float *x4, const float *P, const float *Q,
                       const int32_t *csr_ptr, const int32_t *csr_row, const int32_t *csr_col,
                       const int32_t *csr_eid, const float *edge_attr, float edge_attr_const,
                       int64_t num_nodes, int64_t edges_per_cloud, int n_per_cloud,
                       const float *layer_pack, const float *next_pack, const float *out_pack,
                       float *h_out, float *x4_out, float *x3_out, float *P_out, float *Q_out,
                       float *agg_ws, int impl, void *stream);

/* ---- a15: weighted Kabsch / Procrustes block 3dm:726-758 (evl:786-818) --------------------------
 * One CTA per pair: centroids, H = sum w (p-cs)(q-ct)^T + 1e-6 I, 3x3 SVD (fp64 Jacobi), R = V U^T
 * with the det<0 fix on the smallest-sigma row of Vt, t = ct - R cs.  mask (optional, [pairs][n])
 * restricts the point set (train variant: GT inliers); an empty set gives R=I, t=0 (3dm:708-711).
 * w is used as given (already normalised by the caller).  Outputs R [pairs][9], t [pairs][3],
 * Hout [pairs][9] (optional). */
int egspr_kabsch(const float *p, const float *q, const float *w, const float *mask, int pairs, int n,
                 float *R, float *t, float *Hout, void *stream);

/* ---- a14: eval-variant weights evl:691-783 + Kabsch on the ORIGINAL coords, all n points --------
 * Per pair: sim0 = <feat_src,feat_tgt>, top-128 set (ties -> lower index), p0 = mlp([h_out_src|
 * h_out_tgt][argmax]), the scatter/renormalise/softmax chain of SURVEY A.4, then Kabsch.
 * Also the egnn_equi_loss partial sums (3dm:860-893): loss_parts [pairs][2] =
 * (sum_n label*|R_gt x_src_out + t_gt - x_tgt_out|^2, sum_n (cos(h_src_out,h_tgt_out)-label)^2). */
int egspr_head_eval(const float *feat_src, const float *feat_tgt, const float *x_src,
                    const float *x_tgt, const float *h_out_src, const float *h_out_tgt,
                    const float *x_out_src, const float *x_out_tgt, const float *labels,
                    const float *gt_pose, const float *head_pack, int pairs, int n, int top_k,
                    float *w_out, float *R, float *t, float *Hout, float *loss_parts, void *stream);

/* The same with a scratch buffer (egspr_head_eval_workspace_bytes(pairs) bytes): for few pairs of large clouds (pairs <
 * 2 x SMs and n >= 8192) the passes over the 32-wide rows (input-feature similarity, egnn_equi_loss) are split over up to
 * EGSPR_HEAD_MAX_SPLIT CTAs per pair (needs w_out, which doubles as scratch); identical results. */
#define EGSPR_HEAD_MAX_SPLIT 128
size_t egspr_head_eval_workspace_bytes(int pairs);
int egspr_head_eval_ws(const float *feat_src, const float *feat_tgt, const float *x_src,
                       const float *x_tgt, const float *h_out_src, const float *h_out_tgt,
                       const float *x_out_src, const float *x_out_tgt, const float *labels,
                       const float *gt_pose, const float *head_pack, int pairs, int n, int top_k,
                       float *w_out, float *R, float *t, float *Hout, float *loss_parts, void *workspace,
                       size_t workspace_bytes, void *stream);

/* ---- a13: train-variant weights 3dm:696-724 + Kabsch on the EGNN coords of the GT inliers -------
 * w = softmax over {i: labels!=0} of <h_out_src,h_out_tgt>, /(sum+1e-6).  Same outputs as above;
 * sim_out [pairs][n] receives the similarity scores (used by the host for top-k / BCE / sim loss). */
int egspr_head_train(const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                     const float *x_out_tgt, const float *labels, const float *gt_pose, int pairs,
                     int n, float *w_out, float *sim_out, float *R, float *t, float *Hout,
                     float *loss_parts, void *stream);

/* ---- a17: calculate_pose_error / registration_recall tools/evaluation_metrics.py:14-43 and the F1 of
 * evl:1277, batched on the device (the reference pulls every pose to the host: evl:1249-1270).  fp64 like the
 * reference's numpy code.  src_pts, tgt_pts [pairs][n][3]; gt_pose [pairs][16]; out [pairs][5] doubles =
 * (rotation error deg, translation error cm, recall = sqrt(TP/n), precision = TP/n, F1). tau: 0.09 m. */
int egspr_pose_metrics(const float *R, const float *t, const float *gt_pose, const float *src_pts,
                       const float *tgt_pts, int pairs, int n, double tau, double *out, void *stream);

/* ---- 8(f).3: feature-space correspondence search, data_preprocess/3DMatch_Feature.py:158-166 -------------
 * For every descriptor a[i] (32 floats, unit norm) the nearest b[j] under
 *     distance = sqrt(2 - 2 * <a[i], b[j]> + 1e-6)        (float32, operation by operation like the numpy code)
 * idx[i] = np.argmin(distance[i, :]) (first index on ties), dist[i] = np.min(distance[i, :]).  The [na, nb]
 * similarity matrix is a tcgen05 GEMM (3xTF32) that is never written to memory.  The reference's target_idx
 * (argmin over axis 0, mutual check) is the same call with a and b swapped.
 * workspace: na * 8 bytes. */
int egspr_feature_nn(const float *a, int na, const float *b, int nb, void *workspace, size_t workspace_bytes,
                     int32_t *idx, float *dist, void *stream);

/* =====================================================================================================================
 * Training step (BASELINE config 4): the backward pass `loss.backward()` runs through the same path at 3dm:1125.
 * All gradient buffers are fp32; weight gradients are ACCUMULATED (+=) into grad packs that use the SAME layouts as
 * the weight packs (the host mirror zero-fills them and maps them back onto the nn.Parameters' .grad).
 * ===================================================================================================================== */

/* ---- E_GCL.forward (3dm:280-289) backward for every cloud of the batch.
 * Inputs saved by the forward pass: the layer INPUT state h, x4, P, Q and the per-node message sums agg (what
 * egspr_egcl_forward left in agg_ws).  Nothing per-edge is saved: the edge kernel recomputes each edge's forward.
 * csr_*: the row-major CSR of the forward pass; csc_ptr [G+1] / csc_pos [E]: the same edges grouped by col =
 * edge_index[1], each entry = the edge's POSITION in the row-major CSR (egspr_csr_edge_positions; the per-edge
 * gradients live in row-CSR order so that a node's row list is one contiguous run of rows).
 * dh_out [G][32], dx_out [G][3]: gradient w.r.t. the layer outputs (h', coord').  Writes dh_in [G][32], dx_in [G][3]
 * (gradient w.r.t. the layer inputs; must not alias the *_out buffers) and adds to grad_pack [EGSPR_LAYER_PACK_FLOATS].
 * workspace: egspr_egcl_backward_workspace_bytes(G, E).  Data path is deterministic (segment sums in CSR order);
 * weight gradients are reduced per CTA and combined with one atomicAdd per entry and CTA. */
size_t egspr_egcl_backward_workspace_bytes(int64_t num_nodes, int64_t num_edges);
/* pos_of_edge [E]: row-CSR position of edge (cloud, original id) at [cloud * edges_per_cloud + id].  With csc_eid (the
 * original ids of the col-grouped lists, i.e. the csr_eid output of egspr_csr_from_edges on the swapped edge tensor) also
 * csc_pos [E] = the positions in col-grouped order.  For k-NN graphs (edge id = centre * k + slot) the col-grouped order IS
 * the original order: csc_ptr[g] = g * k and csc_pos = pos_of_edge (pass csc_eid = NULL). */
int egspr_csr_edge_positions(const int32_t *csr_row, const int32_t *csr_eid, const int32_t *csc_eid, int n_per_cloud,
                             int64_t edges_per_cloud, int64_t num_edges, int32_t *pos_of_edge, int32_t *csc_pos, void *stream);
int egspr_egcl_backward(const float *h, const float *x4, const float *P, const float *Q, const float *agg,
                        const int32_t *csr_ptr, const int32_t *csr_row, const int32_t *csr_col,
                        const int32_t *csr_eid, const int32_t *csc_ptr, const int32_t *csc_pos,
                        const float *edge_attr, float edge_attr_const, int64_t num_nodes,
                        int64_t edges_per_cloud, int n_per_cloud, const float *layer_pack,
                        const float *dh_out, const float *dx_out, float *dh_in, float *dx_in,
                        float *grad_pack, void *workspace, size_t workspace_bytes, void *stream);

/* ---- embedding_in / embedding_out (nn.Linear(32,32), 3dm:320-321, 332, 337) as stand-alone ops for the training
 * path (the inference path fuses them into the node kernels).  embed_pack: WT [in][out] + bias.
 * backward: dx (optional) = dy W, grad_pack [EGSPR_EMBED_PACK_FLOATS] += (x^T dy, sum dy). */
int egspr_linear32_forward(const float *x, int64_t rows, const float *embed_pack, float *y, void *stream);
int egspr_linear32_backward(const float *x, const float *dy, int64_t rows, const float *embed_pack, float *dx,
                            float *grad_pack, void *stream);

/* ---- backward of egspr_head_train (3dm:696-758): given dR [pairs][9], dt [pairs][3] and (optionally) dsim
 * [pairs][n] = gradient w.r.t. the similarity scores sim_out, writes the gradients of the four EGNN outputs:
 * dh_src, dh_tgt [pairs][n][32], dx_src, dx_tgt [pairs][n][3].  The SVD is differentiated in closed form
 * (R = V D U^T; for sigma pairs of equal sign only 1/(s_i + s_j) appears), fp64 per pair, fp32 per point.
 * 2*n floats of dynamic shared memory: n <= 24576. */
int egspr_head_train_backward(const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                              const float *x_out_tgt, const float *labels, const float *dR, const float *dt,
                              const float *dsim, int pairs, int n, float *dh_src, float *dh_tgt,
                              float *dx_src, float *dx_tgt, void *stream);

/* ---- 8(f).2: the training losses of one batch on the device.  After egspr_head_train (which leaves sim [pairs][n]):
 * egspr_train_loss_forward, one CTA per pair (src/3dmatch_train_egnn_with_batch.py:681-694, 760-773):
 *   top_idx [pairs][top_k]  members of torch.topk(sim, top_k) (the k largest, ties -> lower index; UNORDERED -- every
 *                           consumer is a mean over the set; -1 past min(top_k, n)),
 *   scores  [pairs][top_k]  mlp([h_out_src | h_out_tgt][top_idx]) logits (3dm:760-768),
 *   bce     [pairs]         sum over the pair's rows of BCEWithLogits(score, labels[top_idx]) (3dm:772),
 *   raw     [pairs][n]      <feat_src, feat_tgt> (3dm:773),
 *   stats   [pairs][4]      fp64 sums of sim, sim^2, raw, raw^2 (the z-scores of 3dm:776-777 use batch-wide statistics).
 *   sim may be NULL: the kernel then recomputes <h_out_src, h_out_tgt> itself (bit-identical to egspr_head_train's sim_out)
 *   and does not depend on egspr_head_train, so the two one-CTA-per-pair launches can run on two streams.
 * egspr_train_loss_finalize, one CTA: loss[0..4] = corr_loss (mean BCE), sim_loss = MSE(zscore(sim), zscore(raw))
 * (unbiased std, +1e-6), mean rot_loss, mean trans_loss (pose_loss 3dm:896-962; 0 when R is null), and their sum =
 * the loop's total (3dm:1118); loss[7] = scale.  Optional seeds of the backward pass, all multiplied by `scale` (the
 * upstream gradient, e.g. 1 / world size): dsim [pairs][n] = d sim_loss / d sim, dR [pairs][9] / dt [pairs][3] =
 * d (mean rot + mean trans) / d (R, t).
 * egspr_head_train_loss_backward = egspr_head_train_backward plus the backward of the mean BCE through mlp: adds
 * d corr_loss / d h_out rows into dh_src / dh_tgt at top_idx and accumulates d corr_loss / d mlp into
 * head_grad_pack [EGSPR_HEAD_PACK_FLOATS] (atomics; zero it first).  `loss` = the finalize kernel's output (scale). */
int egspr_train_loss_forward(const float *h_out_src, const float *h_out_tgt, const float *feat_src,
                             const float *feat_tgt, const float *sim, const float *labels, const float *head_pack,
                             int pairs, int n, int top_k, int32_t *top_idx, float *scores, float *raw,
                             double *stats, float *bce, void *stream);
int egspr_train_loss_finalize(const float *sim, const float *raw, const double *stats, const float *bce, int pairs,
                              int n, int top_k, const float *R, const float *t, const float *gt_pose, float scale,
                              float *loss, float *dsim, float *dR, float *dt, void *stream);
int egspr_head_train_loss_backward(const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                                   const float *x_out_tgt, const float *labels, const float *dR, const float *dt,
                                   const float *dsim, const int32_t *top_idx, const float *head_pack,
                                   const float *loss, int pairs, int n, int top_k, float *dh_src, float *dh_tgt,
                                   float *dx_src, float *dx_tgt, float *head_grad_pack, void *stream);

/* ---- a16: pose_loss(pred_rot, pred_translation, gt_pose) 3dm:896-962 (the two returned losses): per pair
 * rot_loss = acos(clamp((trace(R^T R_gt) - 1) / 2, -1, 1)), trans_loss = acos(clamp(cos(t, t_gt), -1, 1)), and
 * (optional) their gradients grad_R [pairs][9] = d rot_loss / d R, grad_t [pairs][3] = d trans_loss / d t, so that
 * the backward pass is one multiply by the upstream gradient.  R [pairs][9], t [pairs][3], gt_pose [pairs][16]. */
int egspr_pose_loss(const float *R, const float *t, const float *gt_pose, int pairs, float *rot_loss,
                    float *trans_loss, float *grad_R, float *grad_t, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EGSPR_B200_H */
