// Fused E_GCL layer (src/3dmatch_train_egnn_with_batch.py:185-289) for all clouds of a batch.
//
// One persistent launch per layer.  A work item is a block of L_NB consecutive global nodes; the
// CTA walks the block's incoming-edge list (row-major CSR, csr.cu) in tiles of L_TILE edges,
// one edge per thread:
//   edge phase   : geometry (coord2radial :271-278, compute_edge_features :176-181,
//                  compute_so3_matrix :128-173) -> first edge Linear as P[row] + Q[col] + Wgeo*geo
//                  (algebraic split of the 77-wide concat :238-242) -> SiLU -> per-head 8x8 Linear
//                  (:202-208, :245-246) -> LayerNorm (:249) -> coord MLP (:219-229, :264)
//   reduce phase : the tile's messages m_e [32] and d_e*s_e [3] are summed per row in shared memory
//                  and registers, in ascending edge order -- no atomics (:254, :265; SUM not mean)
//   node phase   : node MLP + residual (:252-260), coordinate update (:267), then the NEXT layer's
//                  P/Q halves (or embedding_out :337 after the last layer) while h' is in registers
// so per layer the only HBM/L2 traffic is the gathers and one write of h', x', P', Q'.
#include <cstdlib>

#include "egnn_layer.cuh"

namespace egspr {

template <int NB>
__global__ void __launch_bounds__(L_THREADS, 2) egcl_layer_kernel(const LayerArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;
    float *swea = smem + L_WFLOATS;                  // this layer's edge_attr column (32 floats)
    float *tile = swea + 32;                         // [L_TILE][L_ROW] per-edge messages of the current tile
    float *sacc = tile + L_TILE * L_ROW;             // [NB][L_ROW] per-node accumulators (32 m-sums + 3 coord sums)
    int *sptr = reinterpret_cast<int *>(sacc + NB * L_ROW);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // weights -> shared memory: this layer's edge + node parts, the NEXT layer's P/Q part (or the
    // embedding_out pack) for the node phase, and this layer's own edge_attr column for the edge phase
    for (int i = tid; i < (EDGE_PART + NODE_PART) / 4; i += L_THREADS)
        reinterpret_cast<float4 *>(sw)[i] = ldg4(a.layer_pack + 4 * i);
    if (tid < 32) swea[tid] = __ldg(a.layer_pack + OFF_WEA + tid);
    if (a.next_pack) {
        for (int i = tid; i < PQ_PART / 4; i += L_THREADS)
            reinterpret_cast<float4 *>(sw + OFF_WPT)[i] = ldg4(a.next_pack + OFF_WPT + 4 * i);
    } else if (a.out_pack) {
        for (int i = tid; i < EMBED_PACK / 4; i += L_THREADS)
            reinterpret_cast<float4 *>(sw + OFF_WPT)[i] = ldg4(a.out_pack + 4 * i);
    }

    const int64_t G = a.num_nodes;
    const int64_t items = (G + NB - 1) / NB;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t n0 = item * NB;
        const int nb = (G - n0 < NB) ? (int)(G - n0) : NB;
        __syncthreads();   // previous item's node phase is done with tile / sacc / sptr (also covers the weight load)
        for (int i = tid; i <= nb; i += L_THREADS) sptr[i] = __ldg(a.csr_ptr + n0 + i);
        for (int i = tid; i < NB * L_ROW; i += L_THREADS) sacc[i] = 0.f;
        __syncthreads();
        const int pbeg = sptr[0], pend = sptr[nb];
        int ncur = 0;      // first node of the block whose segment may reach into the current tile (uniform)
        for (int p0 = pbeg; p0 < pend; p0 += L_TILE) {
            {
                int p = p0 + tid;
                if (p >= pend) p = pend - 1;   // idle slot: recompute the last edge, never reduced
                edge_phase(a, sw, swea, tile + tid * L_ROW, p);
            }
            __syncthreads();
            // segmented reduction of the tile into the per-node accumulators: rows of one node are
            // contiguous; the nodes touching this tile are dealt round-robin to the warps; lane = channel
            while (sptr[ncur + 1] <= p0) ++ncur;
            const int tend = min(p0 + L_TILE, pend);
            for (int nl = ncur + warp; nl < nb && sptr[nl] < tend; nl += L_THREADS / 32) {
                const int lo = max(sptr[nl], p0), hi = min(sptr[nl + 1], tend);
                // Continue the node's RUNNING sum (not partial-sum-then-add): every node is summed
                // strictly sequentially in ascending edge order, exactly like scatter_add_ on CPU.
                // This matters beyond rounding noise: exact duplicate points ("twins", normal in the
                // datasets) stay bit-identical layer after layer only if their sums round identically,
                // and the frame rule (:152-163) is discontinuous at x_i == x_j.
                float s0 = sacc[nl * L_ROW + lane];
                float s1 = (lane < 3) ? sacc[nl * L_ROW + 32 + lane] : 0.f;
                for (int p = lo; p < hi; ++p) {
                    const float *rw = tile + (p - p0) * L_ROW;
                    s0 += rw[lane];
                    if (lane < 3) s1 += rw[32 + lane];
                }
                sacc[nl * L_ROW + lane] = s0;
                if (lane < 3) sacc[nl * L_ROW + 32 + lane] = s1;
            }
            __syncthreads();
        }
        // node phase: thread t < nb updates node n0+t (private rows: tile row t for h, sacc row t for agg)
        for (int nl = tid; nl < nb; nl += L_THREADS) {
            const int64_t g = n0 + nl;
            float *arow = sacc + nl * L_ROW;
            const float4 xo = ldg4(a.x4 + g * 4);
            const float nx = xo.x + arow[32], ny = xo.y + arow[33], nz = xo.z + arow[34];   // coord + agg  :267
            *reinterpret_cast<float4 *>(a.x4_out + g * 4) = make_float4(nx, ny, nz, 0.f);
            if (a.x3_out) { a.x3_out[g * 3] = nx; a.x3_out[g * 3 + 1] = ny; a.x3_out[g * 3 + 2] = nz; }
            node_update(a, g, tile + (nl % L_TILE) * L_ROW, arow, sw);
        }
    }
}

// ---- embedding_in + layer-0 P/Q (EGNN.forward :332), one thread per node -----------------------
__device__ __forceinline__ void matvec32_reg(float (&acc)[32], const float *__restrict__ wt, const float (&vec)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) {
#pragma unroll
        for (int o4 = 0; o4 < 8; ++o4) {
            const float4 w = *reinterpret_cast<const float4 *>(wt + 32 * i + 4 * o4);
            acc[4 * o4 + 0] = fmaf(w.x, vec[i], acc[4 * o4 + 0]);
            acc[4 * o4 + 1] = fmaf(w.y, vec[i], acc[4 * o4 + 1]);
            acc[4 * o4 + 2] = fmaf(w.z, vec[i], acc[4 * o4 + 2]);
            acc[4 * o4 + 3] = fmaf(w.w, vec[i], acc[4 * o4 + 3]);
        }
    }
}

__global__ void __launch_bounds__(128) node_embed_kernel(const float *__restrict__ feat,
                                                         const float *__restrict__ x3, int64_t G,
                                                         const float *__restrict__ embed_pack,
                                                         const float *__restrict__ layer0_pack,
                                                         float *__restrict__ h, float *__restrict__ x4,
                                                         float *__restrict__ P, float *__restrict__ Q) {
    __shared__ __align__(16) float sw_raw[EMBED_PACK + PQ_PART];
    float *se = sw_raw;                           // embed: WT [32][32], b [32]
    float *spq = sw_raw + EMBED_PACK - OFF_WPT;   // so that spq[OFF_WPT] is the first PQ float
    if (embed_pack)
        for (int i = threadIdx.x; i < EMBED_PACK; i += blockDim.x) se[i] = __ldg(embed_pack + i);
    for (int i = threadIdx.x; i < PQ_PART; i += blockDim.x) sw_raw[EMBED_PACK + i] = __ldg(layer0_pack + OFF_WPT + i);
    __syncthreads();
    for (int64_t g = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; g < G; g += (int64_t)gridDim.x * blockDim.x) {
        float f[32], hv[32];
        load_row32(f, feat + g * H);
        if (embed_pack) {
            load_bias32(hv, se + 1024);
            matvec32_reg(hv, se, f);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) hv[i] = f[i];
        }
        store_row32(h + g * H, hv);
#pragma unroll
        for (int o = 0; o < 32; ++o) f[o] = 0.f;
        matvec32_reg(f, spq + OFF_WPT, hv);
        store_row32(P + g * H, f);
        load_bias32(f, spq + OFF_BQ);
        matvec32_reg(f, spq + OFF_WQT, hv);
        store_row32(Q + g * H, f);
        if (x4) {
            *reinterpret_cast<float4 *>(x4 + g * 4) =
                make_float4(__ldg(x3 + g * 3), __ldg(x3 + g * 3 + 1), __ldg(x3 + g * 3 + 2), 0.f);
        }
    }
}

template <int NB>
static int launch_layer(const LayerArgs &a, cudaStream_t st) {
    static int ctas_per_sm = 0;        // occupancy of this kernel (the same on every B200 of the box)
    constexpr size_t smem = layer_smem_bytes<NB>();
    if (!opt_in_smem(egcl_layer_kernel<NB>, smem)) return EGSPR_E_LAUNCH;
    if (ctas_per_sm == 0) {
        int occ = 1;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, egcl_layer_kernel<NB>, L_THREADS, smem) != cudaSuccess || occ < 1)
            occ = 1;
        ctas_per_sm = occ;
    }
    const int64_t items = (a.num_nodes + NB - 1) / NB;
    int64_t grid = (int64_t)sm_count() * ctas_per_sm;
    if (grid > items) grid = items;
    egcl_layer_kernel<NB><<<(unsigned)grid, L_THREADS, smem, st>>>(a);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

int launch_layer_ts(const LayerArgs &a, float *agg_ws, bool edge_only, int mode, cudaStream_t st);   // egnn_edge_ts.cu
int launch_node_embed_ts(const float *feat, const float *x3, int64_t G, const float *embed_pack, const float *layer0_pack,
                         float *h, float *x4, float *P, float *Q, cudaStream_t st);   // egnn_node_ts.cu

}  // namespace egspr

extern "C" int egspr_node_embed(const float *feat, const float *x3, int64_t num_nodes, const float *embed_pack,
                                const float *layer0_pack, float *h, float *x4, float *P, float *Q, void *stream) {
    using namespace egspr;
    if (!feat || !layer0_pack || !h || !P || !Q || num_nodes <= 0) return EGSPR_E_INVALID;
    if (x4 && !x3) return EGSPR_E_INVALID;
    return launch_node_embed_ts(feat, x3, num_nodes, embed_pack, layer0_pack, h, x4, P, Q, (cudaStream_t)stream);
}

extern "C" int egspr_egcl_forward(const float *h, const float *x4, const float *P, const float *Q,
                                  const int32_t *csr_ptr, const int32_t *csr_row, const int32_t *csr_col,
                                  const int32_t *csr_eid, const float *edge_attr, float edge_attr_const,
                                  int64_t num_nodes, int64_t edges_per_cloud, int n_per_cloud,
                                  const float *layer_pack, const float *next_pack, const float *out_pack,
                                  float *h_out, float *x4_out, float *x3_out, float *P_out, float *Q_out,
                                  float *agg_ws, int impl, void *stream) {
    using namespace egspr;
    if (!h || !x4 || !P || !Q || !csr_ptr || !csr_row || !csr_col || !layer_pack || !h_out || !x4_out ||
        num_nodes <= 0 || n_per_cloud <= 0)
        return EGSPR_E_INVALID;
    if (edge_attr && !csr_eid) return EGSPR_E_INVALID;
    if (next_pack && (!P_out || !Q_out)) return EGSPR_E_INVALID;
    if (h_out == h || x4_out == x4) return EGSPR_E_INVALID;   // other CTAs still gather the layer input
    LayerArgs a{h, x4, P, Q, csr_ptr, csr_row, csr_col, csr_eid, edge_attr, edge_attr_const, num_nodes,
                edges_per_cloud, n_per_cloud, layer_pack, next_pack, out_pack, h_out, x4_out, x3_out, P_out, Q_out};
    // impl 0 (auto): tensor-core path when scratch is available, else the fused CUDA-core kernel
    const bool big = num_nodes >= (int64_t)256 * 2 * sm_count();
    const bool edge_only = (impl & EGSPR_IMPL_EDGE_ONLY) != 0;
    if (edge_only && (impl & 0xff) != 3 && (impl & 0xff) != 4 && (impl & 0xff) != 5) return EGSPR_E_UNSUPPORTED;
    switch (impl & 0xff) {
        case 0:
            if (agg_ws) return launch_layer_ts(a, agg_ws, false, 0, (cudaStream_t)stream);
            return big ? launch_layer<256>(a, (cudaStream_t)stream) : launch_layer<64>(a, (cudaStream_t)stream);
        case 1: return launch_layer<64>(a, (cudaStream_t)stream);
        case 2: return launch_layer<256>(a, (cudaStream_t)stream);
        case 3:
            if (!agg_ws) return EGSPR_E_WORKSPACE;
            return launch_layer_ts(a, agg_ws, edge_only, 0, (cudaStream_t)stream);
        case 4:
            if (!agg_ws) return EGSPR_E_WORKSPACE;
            return launch_layer_ts(a, agg_ws, edge_only, 1, (cudaStream_t)stream);
        case 5:
            if (!agg_ws) return EGSPR_E_WORKSPACE;
            return launch_layer_ts(a, agg_ws, edge_only, 2, (cudaStream_t)stream);
        default: return EGSPR_E_UNSUPPORTED;
    }
}
