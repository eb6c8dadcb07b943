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
constexpr int OFF_W2P = 384;     // [4][8][8]  second edge Linear per head, [head][in][out]: num_heads = 4 only (CUDA-core
                                 //            kernels, impl 1 / 2); NaN for other head counts
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
constexpr int OFF_W2F = 7104;    // [32][32]   second edge Linear of ALL heads as one block-diagonal matrix, [out][in] (any
                                 //            num_heads dividing 32; the tensor-core kernels read this, never OFF_W2P)
constexpr int LAYER_PACK = EGSPR_LAYER_PACK_FLOATS;
static_assert(LAYER_PACK >= OFF_W2F + 1024, "pack size");
// embed pack: WT [32][32] ([in][out]) + bias [32]
constexpr int EMBED_PACK = EGSPR_EMBED_PACK_FLOATS;
// head pack: W0T [64][32], b0 [32], W1T [32][16], b1 [16], w2 [16], b2 [1]
constexpr int HOFF_W0T = 0, HOFF_B0 = 2048, HOFF_W1T = 2080, HOFF_B1 = 2592, HOFF_W2 = 2608, HOFF_B2 = 2624;
constexpr int HEAD_PACK = EGSPR_HEAD_PACK_FLOATS;

__device__ __forceinline__ float silu(float v) {
    // v * sigmoid(v) = v * rcp(1 + 2^(-v log2 e)): MUFU.EX2 + MUFU.RCP (each within ~2 ulp, far inside the
    // 1e-4 parity budget), no range fix-ups: v -> -inf gives v * rcp(inf) = -0
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
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

// packed fp32 pair FMA (sm_100 FFMA2): (d0, d1) = (a0, a1) * (b0, b1) + (d0, d1), each lane rounded like fmaf --
// one issue slot for two FMAs
__device__ __forceinline__ void ffma2(float &d0, float &d1, float a0, float a1, float b0, float b1) {
    asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rd, {%0, %1};\n"
        "fma.rn.f32x2 rd, ra, rb, rd;\nmov.b64 {%0, %1}, rd;\n}"
        : "+f"(d0), "+f"(d1)
        : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}

__device__ __forceinline__ void fadd2(float &d0, float &d1, float b0, float b1) {        // (d0,d1) += (b0,b1), FADD2
    asm("{\n.reg .b64 ra, rb;\nmov.b64 ra, {%0, %1};\nmov.b64 rb, {%2, %3};\nadd.rn.f32x2 ra, ra, rb;\nmov.b64 {%0, %1}, ra;\n}"
        : "+f"(d0), "+f"(d1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fmul2(float &d0, float &d1, float b0, float b1) {        // (d0,d1) *= (b0,b1), FMUL2
    asm("{\n.reg .b64 ra, rb;\nmov.b64 ra, {%0, %1};\nmov.b64 rb, {%2, %3};\nmul.rn.f32x2 ra, ra, rb;\nmov.b64 {%0, %1}, ra;\n}"
        : "+f"(d0), "+f"(d1) : "f"(b0), "f"(b1));
}
// SiLU on a pair: the multiplies / adds run packed (FMUL2, FADD2); 3 MUFU per pair (two ex2, ONE shared rcp)
__device__ __forceinline__ void silu2(float &x0, float &x1) {
    float t0 = x0, t1 = x1;
    fmul2(t0, t1, -1.4426950408889634f, -1.4426950408889634f);
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(t0));
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(t1));
    // one reciprocal for the pair: 1/a = b / (a b), 1/b = a / (a b); e is clamped so a*b cannot overflow
    // (x < -41: silu(x) = x / (1 + 1e18) instead of x e^x, both below 1e-16 in magnitude)
    t0 = fminf(t0, 1e18f); t1 = fminf(t1, 1e18f);
    fadd2(t0, t1, 1.0f, 1.0f);
    float r = t0 * t1;
    asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(r));
    float r0 = t1, r1 = t0;
    fmul2(r0, r1, r, r);
    fmul2(x0, x1, r0, r1);
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

// ---- programmatic dependent launch (PDL): a kernel launched with launch_pdl may start its CTAs as soon as
// every CTA of the kernel before it in the stream has called pdl_trigger() (or exited); it must call
// pdl_wait() -- which returns once that kernel has COMPLETED and flushed -- before it touches any global
// memory another kernel of the chain produces or consumes.  What runs before pdl_wait() (weight tiles,
// TMEM allocation, barrier init) overlaps the tail of the previous kernel, SM by SM.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// SM count of the CURRENT device (cached per device: one process may drive several GPUs)
inline int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) dev = 0;
    int n = cache[dev];
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
        cache[dev] = n;
    }
    return n;
}

// opt a kernel in to more than 48 KB of dynamic shared memory.  The attribute belongs to the (kernel, device) pair, so
// it is set on every launch path for the current device (a host-side call of ~1 us; launches replayed from a CUDA
// graph do not pay it) instead of being latched in a process-wide flag.
template <class K>
inline bool opt_in_smem(K kernel, size_t bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) == cudaSuccess;
}

}  // namespace egspr
