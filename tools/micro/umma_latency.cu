// Developer micro-benchmark: round-trip time of small tcgen05.mma batches (issue -> commit -> mbarrier wait), one CTA,
// straight-line issue on the uniform datapath (as the product kernels issue them).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_latency umma_latency.cu
#include <cstdio>
#include <cstdlib>

#include "../../se3-equi-graph-registration_b200/csrc/tcgen05.cuh"

using namespace egspr::tc;

template <int MODE, int NM, int INDEP>
__device__ __forceinline__ long long run(uint32_t tm, uint32_t a, uint32_t b, uint32_t mbar, uint32_t &phase) {
    long long best = 1ll << 60;
    const uint64_t da = make_desc_sw128(a), db = make_desc_sw128(b);
    for (int rep = 0; rep < 10; ++rep) {
        const long long t0 = clock64();
#pragma unroll
        for (int j = 0; j < NM; ++j) {
            if constexpr (MODE == 0) umma_bf16_ss(tm + (INDEP ? 32u * (j & 7) : 0u), da + 2 * (j & 3), db + 2 * (j & 3), idesc_bf16(128, 32, 0, 0), 1);
            if constexpr (MODE == 1) umma_bf16_ss(tm + (INDEP ? 32u * (j & 7) : 0u), da + 128 * (j & 7), db + 128 * (j & 7), idesc_bf16(64, 32, 1, 1), 1);
            if constexpr (MODE == 2) umma_bf16_ss(tm + (INDEP ? 128u * (j & 1) : 0u), da + 2 * (j & 3), db + 2 * (j & 3), idesc_bf16(128, 128, 0, 0), 1);
            if constexpr (MODE == 3) umma_bf16_ss(tm + (INDEP ? 64u * (j & 3) : 0u), da + 2 * (j & 3), db + 2 * (j & 3), idesc_bf16(128, 64, 0, 0), 1);
            if constexpr (MODE == 4) umma_tf32_ts(tm + (INDEP ? 32u * (j & 7) : 0u), tm + 256 + 8 * (j & 3), db + 2 * (j & 3), IDESC_TF32_M128_N32, 1);
            if constexpr (MODE == 5) umma_bf16_ss(tm + (INDEP ? 32u * (j & 7) : 0u), da + 128 * (j & 7), db + 2 * (j & 3), idesc_bf16(64, 8, 1, 0), 1);
            if constexpr (MODE == 6) umma_bf16_ss(tm + (INDEP ? 64u * (j & 3) : 0u), da + 128 * (j & 7), db + 128 * (j & 7), idesc_bf16(64, 64, 1, 1), 1);
        }
        umma_commit(mbar);
        mbar_wait(mbar, phase); phase ^= 1;
        const long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
    }
    return best;
}

template <int MODE>
__global__ void __launch_bounds__(128) lat_kernel(long long *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint32_t *holder = reinterpret_cast<uint32_t *>(base + 65536);
    const uint32_t mbar = smem_u32(base + 65536 + 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;
    if (tid < 32) tmem_alloc(smem_u32(holder), 512);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = __shfl_sync(0xffffffffu, *holder, 0);
    const uint32_t a = __shfl_sync(0xffffffffu, smem_u32(base), 0), b = a + 16384, mb = __shfl_sync(0xffffffffu, mbar, 0);
    uint32_t phase = 0;
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    if (warp_u == 0 && elect_one()) {
        out[0] = run<MODE, 1, 0>(tm, a, b, mb, phase);
        out[1] = run<MODE, 2, 0>(tm, a, b, mb, phase);
        out[2] = run<MODE, 4, 0>(tm, a, b, mb, phase);
        out[3] = run<MODE, 8, 0>(tm, a, b, mb, phase);
        out[4] = run<MODE, 16, 0>(tm, a, b, mb, phase);
        out[5] = run<MODE, 32, 0>(tm, a, b, mb, phase);
        out[6] = run<MODE, 8, 1>(tm, a, b, mb, phase);
        out[7] = run<MODE, 32, 1>(tm, a, b, mb, phase);
    }
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 512);
}

template <int MODE>
static void go(const char *name, long long *d) {
    const int smem = 65536 + 64 + 1024;
    cudaFuncSetAttribute(lat_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaMemset(d, 0, 64 * 8);
    lat_kernel<MODE><<<1, 128, smem>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("%s: failed\n", name); exit(1); }
    long long h[8];
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    printf("%-36s n=1 %5lld  2 %5lld  4 %5lld  8 %5lld  16 %5lld  32 %5lld | different D: 8 %5lld  32 %5lld\n", name, h[0], h[1], h[2], h[3], h[4],
           h[5], h[6], h[7]);
}

int main() {
    long long *d;
    cudaMalloc(&d, 64 * 8);
    go<0>("M128 N32 K16 bf16 SS K-major", d);
    go<1>("M64 N32 K16 bf16 SS MN-major", d);
    go<2>("M128 N128 K16 bf16 SS", d);
    go<3>("M128 N64 K16 bf16 SS", d);
    go<4>("M128 N32 K8 tf32 TS", d);
    go<5>("M64 N8 K16 bf16 (A MN, B K)", d);
    go<6>("M64 N64 K16 bf16 SS MN-major", d);
    return 0;
}
