// k-NN graph build: replaces torch_cluster.knn_graph(x, k, loop=True) at the reference call sites
// src/3dmatch_train_egnn_with_batch.py:1005-1006 / src/eval_egnn_metrics.py:1156-1157, for ALL
// clouds of a batch in one launch (the reference loops over 2*B clouds in Python).
//
// Spec (== oracle/knn_oracle.c): k smallest candidates under the total order (d2, index),
// d2 = fma(dz,dz,fma(dy,dy,dx*dx)) in fp32, self included, nearest first, never-filled slots -1,
// candidates with d2 >= 1e10 are never selected (torch_cluster's best_dist init).
//
// Mapping: one warp per query.  Candidate tiles of the query's cloud are staged in shared memory
// (SoA, coalesced loads); each lane scores one candidate per step; the running top-k lives one
// entry per lane (k <= 32), kept sorted; candidates beating the current k-th distance are
// inserted with a ballot + shuffle-up (no local memory, no divergence across queries).
#include "egspr_common.cuh"

namespace egspr {

constexpr int KNN_WARPS = 8;          // queries per CTA
constexpr int KNN_TILE = 2048;        // candidates staged per shared-memory tile

__global__ void __launch_bounds__(KNN_WARPS * 32) knn_warp_select_kernel(
    const float *__restrict__ x, int n, int k, int32_t *__restrict__ nbr) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    const int cloud = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * KNN_WARPS + warp;
    const float *xc = x + (size_t)cloud * n * 3;
    const bool active = qi < n;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = __ldg(xc + 3 * qi); qy = __ldg(xc + 3 * qi + 1); qz = __ldg(xc + 3 * qi + 2); }
    float bd = 1e10f;   // lane s < k: s-th best distance so far
    int bi = -1;
    for (int base = 0; base < n; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += KNN_WARPS * 32) {
            const float *p = xc + 3 * (size_t)(base + i);
            sx[i] = __ldg(p); sy[i] = __ldg(p + 1); sz[i] = __ldg(p + 2);
        }
        __syncthreads();
        if (!active) continue;
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            float d = 3e38f;
            if (j < cnt) {
                const float dx = sx[j] - qx, dy = sy[j] - qy, dz = sz[j] - qz;
                d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            }
            float thr = __shfl_sync(0xffffffffu, bd, k - 1);
            unsigned m = __ballot_sync(0xffffffffu, d < thr);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float dd = __shfl_sync(0xffffffffu, d, src);
                if (dd < thr) {   // warp-uniform: thr may have dropped since the ballot
                    // position = number of kept entries <= dd (the newcomer has the highest index so far)
                    const int pos = __popc(__ballot_sync(0xffffffffu, lane < k && bd <= dd));
                    const float ubd = __shfl_up_sync(0xffffffffu, bd, 1);
                    const int ubi = __shfl_up_sync(0xffffffffu, bi, 1);
                    if (lane > pos) { bd = ubd; bi = ubi; }
                    if (lane == pos) { bd = dd; bi = base + j0 + src; }
                    thr = __shfl_sync(0xffffffffu, bd, k - 1);
                }
            }
        }
    }
    if (active && lane < k) nbr[((size_t)cloud * n + qi) * k + lane] = bi;
}

__global__ void nbr_to_edges_kernel(const int32_t *__restrict__ nbr, int n, int k,
                                    int64_t *__restrict__ edges, int64_t per_cloud) {
    // edges[cloud][0][e] = nbr (row, neighbour); edges[cloud][1][e] = e / k (col, centre)
    const int cloud = blockIdx.y;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < per_cloud;
         e += (int64_t)gridDim.x * blockDim.x) {
        edges[(cloud * 2 + 0) * per_cloud + e] = nbr[cloud * per_cloud + e];
        edges[(cloud * 2 + 1) * per_cloud + e] = e / k;
    }
}

}  // namespace egspr

extern "C" int egspr_knn_build(const float *x, int clouds, int n, int k, int32_t *nbr, void *stream) {
    using namespace egspr;
    if (!x || !nbr || clouds <= 0 || n <= 0) return EGSPR_E_INVALID;
    if (k <= 0 || k > EGSPR_MAX_K) return EGSPR_E_UNSUPPORTED;
    if (clouds > 65535) return EGSPR_E_UNSUPPORTED;
    dim3 grid((n + KNN_WARPS - 1) / KNN_WARPS, clouds);
    knn_warp_select_kernel<<<grid, KNN_WARPS * 32, 0, (cudaStream_t)stream>>>(x, n, k, nbr);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_nbr_to_edges(const int32_t *nbr, int clouds, int n, int k, int64_t *edges, void *stream) {
    using namespace egspr;
    if (!nbr || !edges || clouds <= 0 || n <= 0 || k <= 0) return EGSPR_E_INVALID;
    if (clouds > 65535) return EGSPR_E_UNSUPPORTED;
    const int64_t per_cloud = (int64_t)n * k;
    int64_t gx = (per_cloud + 255) / 256; if (gx > 1024) gx = 1024;
    dim3 grid((unsigned)gx, clouds);
    nbr_to_edges_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nbr, n, k, edges, per_cloud);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}
