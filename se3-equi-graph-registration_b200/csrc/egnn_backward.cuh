// Shared declarations of the E_GCL backward kernels (egnn_backward.cu: node / gather / Linear kernels on CUDA cores,
// egnn_edge_bwd_tc.cu: the edge kernel on tcgen05).
#pragma once
#include "egspr_common.cuh"

namespace egspr {

struct EdgeBwdArgs {
    const float *x4, *P, *Q;
    const int32_t *csr_ptr, *csr_row, *csr_col, *csr_eid;
    const float *edge_attr;
    float edge_attr_const;
    int64_t num_nodes, edges_per_cloud;
    int n_per_cloud;
    const float *pack;
    const float *dagg, *dx_out;     // [G][32], [G][3]
    float *dpre, *dxe;              // [E][32], [E][8] indexed by cloud * edges_per_cloud + original edge id
    float *gpack;
    float *stash;                   // [CTAs][threads][32] scratch of the tensor-core edge kernel (edge_backward_tc_stash_bytes)
};

int launch_edge_backward_tc(const EdgeBwdArgs &a, cudaStream_t st);      // egnn_edge_bwd_tc.cu
size_t edge_backward_tc_stash_bytes();
// node_model backward on tcgen05 (egnn_node_ts.cu)
int launch_node_mlp_backward_ts(const float *h, const float *agg, const float *dh_out, int64_t G, const int32_t *csr_ptr,
                                const float *pack, float *dh_in, float *dagg, float *gpack, cudaStream_t st);

// sum over the 32 lanes of v[lane'] for every column: returns, in lane L, sum over lanes of v[L]
__device__ __forceinline__ float warp_colsum32(const float (&v)[32]) {
    constexpr unsigned FULLM = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float a[16], b[8], c[4], d[2];
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const float keep = up ? v[j + 16] : v[j], send = up ? v[j] : v[j + 16];
            a[j] = keep + __shfl_xor_sync(FULLM, send, 16);
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float keep = up ? a[j + 8] : a[j], send = up ? a[j] : a[j + 8];
            b[j] = keep + __shfl_xor_sync(FULLM, send, 8);
        }
    }
    {
        const bool up = lane & 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float keep = up ? b[j + 4] : b[j], send = up ? b[j] : b[j + 4];
            c[j] = keep + __shfl_xor_sync(FULLM, send, 4);
        }
    }
    {
        const bool up = lane & 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float keep = up ? c[j + 2] : c[j], send = up ? c[j] : c[j + 2];
            d[j] = keep + __shfl_xor_sync(FULLM, send, 2);
        }
    }
    const bool up = lane & 1;
    const float keep = up ? d[1] : d[0], send = up ? d[0] : d[1];
    return keep + __shfl_xor_sync(FULLM, send, 1);
}

}  // namespace egspr
