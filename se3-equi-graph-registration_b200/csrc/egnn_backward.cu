// Backward pass of the E_GCL / EGNN stack (the gradient `loss.backward()` produces in the reference's training
// step, src/3dmatch_train_egnn_with_batch.py:1125, through E_GCL.forward 3dm:280-289 and EGNN.forward 3dm:328-340).
//
// Three kernels per layer, deterministic data path (no atomics on activations):
//   node_mlp_backward_ts_kernel (egnn_node_ts.cu, tcgen05) thread = node: node_model (3dm:252-260) backward -> dh (direct
//                              part), dagg; its weight-gradient outer products run beside the MMAs
//   edge_backward_tc_kernel    (egnn_edge_bwd_tc.cu, tcgen05) thread = edge (row-CSR order): recomputes the edge's forward
//                              from the layer input (nothing per-edge is kept from the forward pass) and pushes
//                              (dagg[row], dx_out[row]) back to dpre (= dP[row] = dQ[col]) and the two endpoints'
//                              coordinate gradients, written in row-CSR order (one row per edge position)
//   node_gather_backward_kernel thread = node: sums dpre / dx over the node's row list (row-CSR) and col list
//                              (col-CSR), then the P/Q halves of the first edge Linear back to dh
// Weight gradients: every kernel stages the rows of its outer products in shared memory per 128-row tile, each
// thread owns 8 (or 2 / 4) entries of a weight matrix in registers across all tiles of its CTA, column sums
// (biases, LayerNorm, wc2) are transposed-reduced with warp shuffles; one atomicAdd per entry and CTA at the end.
// The arithmetic itself lives in egnn_backward_math.cuh, which the CPU test-suite compiles with g++.
#include "egnn_backward.cuh"
#include "egnn_backward_math.cuh"

namespace egspr {
using namespace bwd;

constexpr int BT = 128;    // threads per CTA = rows (edges / nodes) per tile
constexpr int RS = 33;     // stash row stride in floats: odd -> own-row writes and column reads are conflict-free

__device__ __forceinline__ void load_row32g(float (&v)[32], const float *__restrict__ p) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = ldg4(p + 4 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row32g(float *__restrict__ p, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4 *>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
constexpr int NT = 1024;            // threads of the 8-lanes-per-row kernels
constexpr int NOS = BT + 4;         // feature-major row stride of their "out" tiles

// ---------------------------------------------------------------------------------------------------------------
// node gather + P/Q backward
// ---------------------------------------------------------------------------------------------------------------
struct GatherArgs {
    const float *h;
    const int32_t *csr_ptr, *csc_ptr, *csc_pos;
    int64_t num_nodes, edges_per_cloud;
    int n_per_cloud;
    const float *pack;
    const float *dpre, *dxe, *dx_out;
    float *dh_in, *dx_in;      // dh_in holds the node-MLP part on entry
    float *gpack;
};
// One CTA = 1024 threads = a tile of 128 nodes, 8 lanes per node (4 features each): every node's two edge lists are
// walked concurrently (the kernel is a chain of dependent loads per node -- id, then row -- so its time is the chain
// length times the number of waves, not the bytes), rows move as coalesced 128-byte lines, sums in list order.
constexpr int GT = 1024;
constexpr int GOS = BT + 4;          // feature-major row stride of the dP / dQ tiles (as OS in the edge kernel)
constexpr size_t GB_SMEM = sizeof(float) * (2048 + BT * RS + 2 * 32 * GOS);

__global__ void __launch_bounds__(GT, 1) node_gather_backward_kernel(const GatherArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *swT = smem;                                   // W_P and W_Q as [out][in] (transposed packs): float4 over "in"
    float *sH = smem + 2048;                             // h rows, node-major [BT][RS]
    float *sDp = sH + BT * RS, *sDq = sDp + 32 * GOS;    // dP, dQ feature-major [32][GOS]: four nodes per 128-bit load
    for (int i = threadIdx.x; i < 2048; i += GT) {
        const int m = i >> 10, o = (i >> 5) & 31, in = i & 31;
        swT[i] = __ldg(a.pack + B_WPT + 1024 * m + 32 * in + o);
    }
    const int sub = threadIdx.x & 7, ln = threadIdx.x >> 3;          // node of the tile, 4-feature slice
    const int ai = threadIdx.x & 31, ao = threadIdx.x >> 5;          // weight-gradient entry [in = ai][out = ao]
    float accP0 = 0.f, accP1 = 0.f, accQ0 = 0.f, accQ1 = 0.f, colQ = 0.f;
    const int64_t G = a.num_nodes;
    const int64_t tiles = (G + BT - 1) / BT;
    const unsigned gmask = 0xffu << (threadIdx.x & 24);
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t n = tile * BT + ln;
        float4 dp = make_float4(0.f, 0.f, 0.f, 0.f), dq = dp, hv = dp;
        float dxa = 0.f;                                               // lanes 0..2 carry x, y, z
        if (n < G) {
            // edges with row == n: dpre / dxe are stored in row-CSR order, so this list is one contiguous run of rows
            for (int p = __ldg(a.csr_ptr + n), pe = __ldg(a.csr_ptr + n + 1); p < pe; p += 8) {
                float4 t[8];
                float tx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool in = p + j < pe;
                    const int64_t ge = in ? p + j : p;
                    t[j] = in ? ldg4(a.dpre + ge * H + 4 * sub) : make_float4(0.f, 0.f, 0.f, 0.f);
                    tx[j] = (!in || sub >= 3) ? 0.f : __ldg(a.dxe + ge * 8 + sub);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { dp.x += t[j].x; dp.y += t[j].y; dp.z += t[j].z; dp.w += t[j].w; dxa += tx[j]; }
            }
            for (int p = __ldg(a.csc_ptr + n), pe = __ldg(a.csc_ptr + n + 1); p < pe; p += 8) {   // edges with col == n
                const int mine = p + sub < pe ? __ldg(a.csc_pos + p + sub) : -1;      // row-CSR position of the edge
                float4 t[8];
                float tx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ej = __shfl_sync(gmask, mine, j, 8);
                    const int64_t ge = ej < 0 ? 0 : ej;
                    t[j] = ej < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4(a.dpre + ge * H + 4 * sub);
                    tx[j] = (ej < 0 || sub >= 3) ? 0.f : __ldg(a.dxe + ge * 8 + 4 + sub);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { dq.x += t[j].x; dq.y += t[j].y; dq.z += t[j].z; dq.w += t[j].w; dxa += tx[j]; }
            }
            if (sub < 3) a.dx_in[n * 3 + sub] = __ldg(a.dx_out + n * 3 + sub) + dxa;
            hv = ldg4(a.h + n * H + 4 * sub);
        }
        {
            float *rh = sH + ln * RS + 4 * sub, *rp = sDp + (4 * sub) * GOS + ln, *rq = sDq + (4 * sub) * GOS + ln;
            rh[0] = hv.x; rh[1] = hv.y; rh[2] = hv.z; rh[3] = hv.w;
            rp[0] = dp.x; rp[GOS] = dp.y; rp[2 * GOS] = dp.z; rp[3 * GOS] = dp.w;
            rq[0] = dq.x; rq[GOS] = dq.y; rq[2 * GOS] = dq.z; rq[3 * GOS] = dq.w;
        }
        __syncthreads();
        // P/Q halves of the first edge Linear back to dh: this lane's 4 inputs i = 4 sub .. 4 sub + 3
        if (n < G) {
            float4 acc = *reinterpret_cast<const float4 *>(a.dh_in + n * H + 4 * sub);
            const float *rp = sDp + ln, *rq = sDq + ln;
#pragma unroll 4
            for (int o = 0; o < 32; ++o) {
                const float pv = rp[o * GOS], qv = rq[o * GOS];
                const float4 wp = *reinterpret_cast<const float4 *>(swT + 32 * o + 4 * sub);
                const float4 wq = *reinterpret_cast<const float4 *>(swT + 1024 + 32 * o + 4 * sub);
                acc.x = fmaf(wp.x, pv, acc.x); acc.y = fmaf(wp.y, pv, acc.y); acc.z = fmaf(wp.z, pv, acc.z); acc.w = fmaf(wp.w, pv, acc.w);
                acc.x = fmaf(wq.x, qv, acc.x); acc.y = fmaf(wq.y, qv, acc.y); acc.z = fmaf(wq.z, qv, acc.z); acc.w = fmaf(wq.w, qv, acc.w);
            }
            *reinterpret_cast<float4 *>(a.dh_in + n * H + 4 * sub) = acc;
        }
        // weight gradients: dWPT[i][o] += h[i] dP[o], dWQT[i][o] += h[i] dQ[o], dbq[o] += dQ[o]; one entry per thread,
        // four nodes per 128-bit load of the dP / dQ rows (broadcast: `ao` is the warp index)
        {
            const float *in = sH + ai, *op = sDp + ao * GOS, *oq = sDq + ao * GOS;
#pragma unroll 2
            for (int e = 0; e < BT; e += 4) {
                const float h0 = in[e * RS], h1 = in[(e + 1) * RS], h2 = in[(e + 2) * RS], h3 = in[(e + 3) * RS];
                const float4 pv = *reinterpret_cast<const float4 *>(op + e), qv = *reinterpret_cast<const float4 *>(oq + e);
                fma2(accP0, accP1, h0, h1, pv.x, pv.y); fma2(accP0, accP1, h2, h3, pv.z, pv.w);
                fma2(accQ0, accQ1, h0, h1, qv.x, qv.y); fma2(accQ0, accQ1, h2, h3, qv.z, qv.w);
            }
        }
        if (threadIdx.x < 32) {
            const float *oq = sDq + threadIdx.x * GOS;
#pragma unroll 4
            for (int e = 0; e < BT; e += 4) {
                const float4 qv = *reinterpret_cast<const float4 *>(oq + e);
                colQ += (qv.x + qv.y) + (qv.z + qv.w);
            }
        }
        __syncthreads();
    }
    atomicAdd(a.gpack + B_WPT + 32 * ai + ao, accP0 + accP1);
    atomicAdd(a.gpack + B_WQT + 32 * ai + ao, accQ0 + accQ1);
    if (threadIdx.x < 32) atomicAdd(a.gpack + B_BQ + threadIdx.x, colQ);
}

// ---------------------------------------------------------------------------------------------------------------
// Linear(32,32) forward / backward (embedding_in 3dm:332, embedding_out 3dm:337) -- embed pack: WT [in][out], b
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) linear32_forward_kernel(const float *__restrict__ x, int64_t rows,
                                                              const float *__restrict__ pack, float *__restrict__ y) {
    __shared__ __align__(16) float sw[1056];
    for (int i = threadIdx.x; i < 1056; i += BT) sw[i] = __ldg(pack + i);
    __syncthreads();
    for (int64_t n = (int64_t)blockIdx.x * BT + threadIdx.x; n < rows; n += (int64_t)gridDim.x * BT) {
        float xv[32], yv[32];
        load_row32g(xv, x + n * H);
#pragma unroll
        for (int o = 0; o < 32; ++o) yv[o] = sw[1024 + o];
#pragma unroll
        for (int i = 0; i < 32; ++i)
#pragma unroll
            for (int o = 0; o < 32; ++o) yv[o] = fmaf(sw[32 * i + o], xv[i], yv[o]);
        store_row32g(y + n * H, yv);
    }
}

// Linear(32,32) backward, 8 lanes per row (the default): CTA = 1024 threads = 128 rows, every lane owns 4 of the 32
// channels, rows move as coalesced 128-byte lines, dy tile feature-major so the reduction reads 4 rows per 128-bit load
// (34 us per launch vs 64 us for the thread-per-row kernel; the same restructuring of the node-MLP backward measured
// SLOWER than its thread-per-node kernel, 106 vs 91 us, and was dropped)
constexpr size_t LW_SMEM = sizeof(float) * (1024 + BT * RS + 32 * NOS);

__global__ void __launch_bounds__(NT, 1) linear32_backward_wide_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                                        int64_t rows, const float *__restrict__ pack,
                                                                        float *__restrict__ dx, float *__restrict__ gpack) {
    extern __shared__ __align__(16) float smem[];
    float *sWo = smem, *sX = smem + 1024, *sDy = sX + BT * RS;       // W as [out][in]; x node-major; dy feature-major
    for (int i = threadIdx.x; i < 1024; i += NT) sWo[(i & 31) * 32 + (i >> 5)] = __ldg(pack + i);
    const int sub = threadIdx.x & 7, ln = threadIdx.x >> 3, c0 = 4 * sub;
    const int ai = threadIdx.x & 31, ao = threadIdx.x >> 5;
    float acc[2] = {0.f, 0.f}, colDy = 0.f;
    const int64_t tiles = (rows + BT - 1) / BT;
    __syncthreads();
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t n = tile * BT + ln;
        const bool valid = n < rows;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f), dv = xv;
        if (valid) { xv = ldg4(x + n * H + c0); dv = ldg4(dy + n * H + c0); }
        float *rX = sX + ln * RS, *rD = sDy + ln;
        rX[c0] = xv.x; rX[c0 + 1] = xv.y; rX[c0 + 2] = xv.z; rX[c0 + 3] = xv.w;
        rD[c0 * NOS] = dv.x; rD[(c0 + 1) * NOS] = dv.y; rD[(c0 + 2) * NOS] = dv.z; rD[(c0 + 3) * NOS] = dv.w;
        __syncwarp();
        if (dx && valid) {
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
            for (int o = 0; o < 32; ++o) {
                const float v = rD[o * NOS];
                const float4 w = *reinterpret_cast<const float4 *>(sWo + 32 * o + c0);
                fma2(g.x, g.y, w.x, w.y, v, v); fma2(g.z, g.w, w.z, w.w, v, v);
            }
            *reinterpret_cast<float4 *>(dx + n * H + c0) = g;
        }
        __syncthreads();
        {
            const float *in = sX + ai, *od = sDy + ao * NOS;
#pragma unroll 2
            for (int e = 0; e < BT; e += 4) {
                const float4 dd = *reinterpret_cast<const float4 *>(od + e);
                fma2(acc[0], acc[1], in[e * RS], in[(e + 1) * RS], dd.x, dd.y);
                fma2(acc[0], acc[1], in[(e + 2) * RS], in[(e + 3) * RS], dd.z, dd.w);
            }
        }
        if (threadIdx.x < 32) {
            const float *o = sDy + threadIdx.x * NOS;
#pragma unroll 4
            for (int e = 0; e < BT; e += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(o + e);
                colDy += (v.x + v.y) + (v.z + v.w);
            }
        }
        __syncthreads();
    }
    atomicAdd(gpack + 32 * ai + ao, acc[0] + acc[1]);
    if (threadIdx.x < 32) atomicAdd(gpack + 1024 + threadIdx.x, colDy);
}

template <class K>
static int prep_kernel(K kernel, size_t smem, int &ctas_per_sm, int threads = BT) {
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return EGSPR_E_LAUNCH;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kernel, threads, smem) != cudaSuccess || ctas_per_sm < 1)
        ctas_per_sm = 1;
    return EGSPR_OK;
}
static unsigned grid_for(int64_t tiles, int ctas_per_sm) {
    int64_t g = (int64_t)sm_count() * ctas_per_sm;
    if (g > tiles) g = tiles;
    return (unsigned)(g < 1 ? 1 : g);
}

}  // namespace egspr

namespace egspr {
// pos[cloud * epc + eid] = p for every row-CSR position p;  cpos[q] = pos[cloud(q) * epc + ceid[q]] for a col-grouped list
__global__ void csr_positions_kernel(const int32_t *__restrict__ csr_row, const int32_t *__restrict__ csr_eid, int n, int64_t epc,
                                     int64_t E, int32_t *__restrict__ pos) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < E; p += (int64_t)gridDim.x * blockDim.x)
        pos[(int64_t)(__ldg(csr_row + p) / n) * epc + __ldg(csr_eid + p)] = (int32_t)p;
}
__global__ void csc_positions_kernel(const int32_t *__restrict__ pos, const int32_t *__restrict__ ceid, int64_t epc, int64_t E,
                                     int32_t *__restrict__ cpos) {
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < E; q += (int64_t)gridDim.x * blockDim.x)
        cpos[q] = __ldg(pos + (q / epc) * epc + __ldg(ceid + q));
}
}  // namespace egspr

extern "C" int egspr_csr_edge_positions(const int32_t *csr_row, const int32_t *csr_eid, const int32_t *csc_eid, int n_per_cloud,
                                        int64_t edges_per_cloud, int64_t num_edges, int32_t *pos_of_edge, int32_t *csc_pos,
                                        void *stream) {
    using namespace egspr;
    if (!csr_row || !csr_eid || !pos_of_edge || n_per_cloud <= 0 || edges_per_cloud <= 0 || num_edges <= 0) return EGSPR_E_INVALID;
    if (csc_eid && !csc_pos) return EGSPR_E_INVALID;
    int64_t g = (num_edges + 255) / 256;
    if (g > 8 * sm_count()) g = 8 * sm_count();
    csr_positions_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(csr_row, csr_eid, n_per_cloud, edges_per_cloud, num_edges, pos_of_edge);
    if (csc_eid) csc_positions_kernel<<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>(pos_of_edge, csc_eid, edges_per_cloud, num_edges, csc_pos);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" size_t egspr_egcl_backward_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
    if (num_nodes <= 0 || num_edges < 0) return 0;
    // dagg [G][32] | dpre [E][32] | dxe [E][8] | the tensor-core edge kernel's per-thread scratch
    return sizeof(float) * ((size_t)num_nodes * 32 + (size_t)num_edges * 40) + egspr::edge_backward_tc_stash_bytes() + 512;
}

extern "C" int egspr_egcl_backward(const float *h, const float *x4, const float *P, const float *Q, const float *agg,
                                   const int32_t *csr_ptr, const int32_t *csr_row, const int32_t *csr_col,
                                   const int32_t *csr_eid, const int32_t *csc_ptr, const int32_t *csc_pos,
                                   const float *edge_attr, float edge_attr_const, int64_t num_nodes,
                                   int64_t edges_per_cloud, int n_per_cloud, const float *layer_pack,
                                   const float *dh_out, const float *dx_out, float *dh_in, float *dx_in,
                                   float *grad_pack, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace egspr;
    if (!h || !x4 || !P || !Q || !agg || !csr_ptr || !csr_row || !csr_col || !csr_eid || !csc_ptr || !csc_pos ||
        !layer_pack || !dh_out || !dx_out || !dh_in || !dx_in || !grad_pack || !workspace || num_nodes <= 0 ||
        n_per_cloud <= 0 || edges_per_cloud <= 0 || num_nodes % n_per_cloud != 0)
        return EGSPR_E_INVALID;
    if (dh_in == dh_out || dx_in == dx_out) return EGSPR_E_INVALID;
    const int64_t E = (num_nodes / n_per_cloud) * edges_per_cloud;
    if (workspace_bytes < egspr_egcl_backward_workspace_bytes(num_nodes, E)) return EGSPR_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    float *dagg = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~(uintptr_t)255);
    float *dpre = dagg + (size_t)num_nodes * 32;
    float *dxe = dpre + (size_t)E * 32;
    float *stash = reinterpret_cast<float *>((reinterpret_cast<uintptr_t>(dxe + (size_t)E * 8) + 255) & ~(uintptr_t)255);
    int occ_gather = 1;
    if (int e = prep_kernel(node_gather_backward_kernel, GB_SMEM, occ_gather, GT)) return e;
    const int64_t ntiles = (num_nodes + BT - 1) / BT;
    if (int e = launch_node_mlp_backward_ts(h, agg, dh_out, num_nodes, csr_ptr, layer_pack, dh_in, dagg, grad_pack, st)) return e;
    EdgeBwdArgs ea{x4, P, Q, csr_ptr, csr_row, csr_col, csr_eid, edge_attr, edge_attr_const, num_nodes, edges_per_cloud,
                   n_per_cloud, layer_pack, dagg, dx_out, dpre, dxe, grad_pack, stash};
    if (int e = launch_edge_backward_tc(ea, st)) return e;
    GatherArgs ga{h, csr_ptr, csc_ptr, csc_pos, num_nodes, edges_per_cloud, n_per_cloud, layer_pack, dpre, dxe,
                  dx_out, dh_in, dx_in, grad_pack};
    node_gather_backward_kernel<<<grid_for(ntiles, occ_gather), GT, GB_SMEM, st>>>(ga);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_linear32_forward(const float *x, int64_t rows, const float *embed_pack, float *y, void *stream) {
    using namespace egspr;
    if (!x || !embed_pack || !y || rows <= 0) return EGSPR_E_INVALID;
    const int64_t tiles = (rows + BT - 1) / BT;
    linear32_forward_kernel<<<grid_for(tiles, 8), BT, 0, (cudaStream_t)stream>>>(x, rows, embed_pack, y);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_linear32_backward(const float *x, const float *dy, int64_t rows, const float *embed_pack, float *dx,
                                       float *grad_pack, void *stream) {
    using namespace egspr;
    if (!x || !dy || !embed_pack || !grad_pack || rows <= 0) return EGSPR_E_INVALID;
    int occ_wide = 1;
    if (int e = prep_kernel(linear32_backward_wide_kernel, LW_SMEM, occ_wide, NT)) return e;
    const int64_t tiles = (rows + BT - 1) / BT;
    linear32_backward_wide_kernel<<<grid_for(tiles, occ_wide), NT, LW_SMEM, (cudaStream_t)stream>>>(x, dy, rows, embed_pack, dx, grad_pack);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}
