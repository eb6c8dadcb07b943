// Shared device/host helpers for the egspr_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/egspr_b200.h"

#define EGSPR_CHECK_LAUNCH()                                   \
    do {                                                       \
        if (cudaGetLastError() != cudaSuccess) return EGSPR_E_LAUNCH; \
    } while (0)

namespace egspr {

constexpr int H = EGSPR_HIDDEN;  // hidden width

// ---- layer pack offsets (floats); mirrored by packing.py -------------------------------------
constexpr int OFF_WG = 0;        // [12][32]   geo rows of the first edge Linear (cols 64..75), out-contiguous
constexpr int OFF_W2P = 384;     // [4][8][8]  second edge Linear per head, [head][in][out]
constexpr int OFF_B2 = 640;      // [32]
constexpr int OFF_LNG = 672;     // [32]
constexpr int OFF_LNB = 704;     // [32]
constexpr int OFF_WC1 = 736;     // [32][32]   coord_mlp.0.weight as stored ([out][in])
constexpr int OFF_BC1 = 1760;    // [32]
constexpr int OFF_WC2 = 1792;    // [32]
constexpr int EDGE_PART = 1824;
constexpr int OFF_WN1T = 1824;   // [64][32]   node_mlp.0.weight transposed ([in][out])
constexpr int OFF_BN1 = 3872;    // [32]
constexpr int OFF_WN2T = 3904;   // [32][32]
constexpr int OFF_BN2 = 4928;    // [32]
constexpr int NODE_PART = 4960 - 1824;
constexpr int OFF_WPT = 4960;    // [32][32]   first edge Linear, h[row] block, [in][out]
constexpr int OFF_WQT = 5984;    // [32][32]   first edge Linear, h[col] block, [in][out]
constexpr int OFF_BQ = 7008;     // [32]       first edge Linear bias (heads concatenated)
constexpr int OFF_WEA = 7040;    // [32]       first edge Linear edge_attr column (zeros if edges_in_d=0)
constexpr int PQ_PART = 7072 - 4960;
constexpr int LAYER_PACK = EGSPR_LAYER_PACK_FLOATS;
static_assert(LAYER_PACK >= 7072, "pack size");
// embed pack: WT [32][32] ([in][out]) + bias [32]
constexpr int EMBED_PACK = EGSPR_EMBED_PACK_FLOATS;
// head pack: W0T [64][32], b0 [32], W1T [32][16], b1 [16], w2 [16], b2 [1]
constexpr int HOFF_W0T = 0, HOFF_B0 = 2048, HOFF_W1T = 2080, HOFF_B1 = 2592, HOFF_W2 = 2608, HOFF_B2 = 2624;
constexpr int HEAD_PACK = EGSPR_HEAD_PACK_FLOATS;

__device__ __forceinline__ float silu(float v) {
    // v * sigmoid(v) with MUFU.EX2 / MUFU.RCP (each within ~2 ulp, far inside the 1e-4 parity budget)
    return __fdividef(v, 1.0f + __expf(-v));
}
__device__ __forceinline__ float fast_sqrt(float v) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ float fast_rcp(float v) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace egspr
