// Probe of the tcgen05 operand-layout assumptions the tensor-core edge-backward kernel relies on (developer tool,
// not part of the product library).  One CTA, exact small-integer bf16 data, so every check is bit-exact.
//   T1  K-major A (128 x 64 bf16, SW128) x K-major B with a sub-row K offset per K-step
//   T2  the same weight tile read as an MN-major B (transposed product without a transposed copy)
//   T3  MN-major A (M=64) x MN-major B (N=32 slices at +0 / +64 B of a 128-byte row), accumulate, 8 K-steps
//   T4  as T3 at TMEM lane offset 16 (interleaved M=64 accumulators)
//   T5  MN-major B slice of N=16 at +32 B
//   T6  column sums through a constant all-ones B (N=8)
//   T7  A operand from TENSOR MEMORY as packed bf16 pairs (TS mode, kind::f16)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o umma_probe umma_probe.cu ; run on a B200.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../se3-equi-graph-registration_b200/csrc/tcgen05.cuh"

using namespace egspr::tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t addr) { return make_desc_sw128(addr); }
__device__ __forceinline__ uint64_t desc_none(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr & 0x3ffff) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

// tiles: [rows][64] bf16 logical; stored 128-byte rows with 16-byte chunk c of row r at c ^ (r & 7)
__device__ __forceinline__ int sw_off(int row, int col_bf16) { return row * 128 + ((((col_bf16 >> 3) ^ (row & 7)) << 4) | ((col_bf16 & 7) << 1)); }

__global__ void __launch_bounds__(128) probe_kernel(const uint16_t *gA, const uint16_t *gY, const uint16_t *gW, float *out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *TA = base, *TY = base + 16384, *TW = base + 32768, *ONES = base + 32768 + 8192;
    uint32_t *holder = reinterpret_cast<uint32_t *>(base + 32768 + 8192 + 1024);
    const uint32_t mbar = smem_u32(base + 32768 + 8192 + 1024 + 16);
    const int tid = threadIdx.x;
    for (int i = tid; i < 128 * 64; i += 128) {
        const int r = i >> 6, c = i & 63;
        *reinterpret_cast<uint16_t *>(TA + sw_off(r, c)) = gA[i];
        *reinterpret_cast<uint16_t *>(TY + sw_off(r, c)) = gY[i];
        if (r < 64) *reinterpret_cast<uint16_t *>(TW + sw_off(r, c)) = gW[i];
    }
    for (int i = tid; i < 512; i += 128) reinterpret_cast<uint16_t *>(ONES)[i] = 0x3F80;
    if (tid < 32) tmem_alloc(smem_u32(holder), 512);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tm = *holder;
    // zero the accumulator region this probe reads (all 128 columns of this lane)
    {
        float z[16];
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
        for (int c = 0; c < 128; c += 16) tmem_st16(tm + ((uint32_t)((tid >> 5) * 32) << 16) + c, z);
        tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        fence_after_sync();
        const uint32_t a = smem_u32(TA), y = smem_u32(TY), w = smem_u32(TW), o = smem_u32(ONES);
        // T1: cols 0..31
        for (int j = 0; j < 4; ++j)
            umma_bf16_ss(tm + 0, desc_sw128(a + 32 * j), desc_sw128(w + 32 * (j & 1)), idesc_bf16(128, 32, 0, 0), j > 0);
        // T2: cols 32..63   D[e][i] = sum_{o<32} A[e][o] W[o][i]
        for (int j = 0; j < 2; ++j)
            umma_bf16_ss(tm + 32, desc_sw128(a + 32 * j), desc_sw128(w + 2048 * j), idesc_bf16(128, 32, 0, 1), j > 0);
        // T3: cols 64..95, lanes +0
        for (int s = 0; s < 2; ++s)
            for (int j = 0; j < 8; ++j)
                umma_bf16_ss(tm + 64, desc_sw128(a + 2048 * j), desc_sw128(y + 64 * s + 2048 * j), idesc_bf16(64, 32, 1, 1), (s | j) > 0);
        // T4: cols 64..95, lanes +16: only the second slice
        for (int j = 0; j < 8; ++j)
            umma_bf16_ss(tm + 64 + (16u << 16), desc_sw128(a + 2048 * j), desc_sw128(y + 64 + 2048 * j), idesc_bf16(64, 32, 1, 1), j > 0);
        // T5: cols 96..111: N = 16 slice at +32 B
        for (int j = 0; j < 8; ++j)
            umma_bf16_ss(tm + 96, desc_sw128(a + 2048 * j), desc_sw128(y + 32 + 2048 * j), idesc_bf16(64, 16, 1, 1), j > 0);
        // T6: cols 112..119: ones
        for (int j = 0; j < 8; ++j)
            umma_bf16_ss(tm + 112, desc_sw128(a + 2048 * j), desc_none(o), idesc_bf16(64, 8, 1, 0), j > 0);
        umma_commit(mbar);
    }
    // T7: A from TENSOR MEMORY as packed bf16 pairs (column c of lane e = {A[e][2c] low half, A[e][2c+1] high half}), K = 32:
    // two K-steps of 8 columns; B = first 64 bytes of the weight tile rows (K-major).  D -> cols 120..151 would collide,
    // so it reuses cols 0..31 AFTER T1 was read back (second phase below).
    mbar_wait(mbar, 0);
    fence_after_sync();
    for (int c = 0; c < 128; c += 32) {
        float v[32];
        tmem_ld32(tm + ((uint32_t)((tid >> 5) * 32) << 16) + c, v);
        for (int i = 0; i < 32; ++i) out[tid * 128 + c + i] = v[i];
    }
    {   // second phase: T7
        float pk[16];
        for (int c = 0; c < 16; ++c) {
            const uint32_t lo = gA[tid * 64 + 2 * c], hi = gA[tid * 64 + 2 * c + 1];
            pk[c] = __uint_as_float(lo | (hi << 16));
        }
        tmem_st16(tm + ((uint32_t)((tid >> 5) * 32) << 16) + 256, pk);
        tmem_wait_st();
        fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            fence_after_sync();
            const uint32_t w = smem_u32(TW);
            for (int j = 0; j < 2; ++j) {
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tm + 160), "r"(tm + 256 + 8 * j),
                    "l"(desc_sw128(w + 32 * j)), "r"(idesc_bf16(128, 32, 0, 0)), "r"(j)
                    : "memory");
            }
            umma_commit(mbar);
        }
        mbar_wait(mbar, 1);
        fence_after_sync();
        float v[32];
        tmem_ld32(tm + ((uint32_t)((tid >> 5) * 32) << 16) + 160, v);
        for (int i = 0; i < 32; ++i) out[128 * 128 + tid * 32 + i] = v[i];
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tm, 512);
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }
static float bf2f(uint16_t b) { uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f; }

int main(int argc, char **argv) {
    std::vector<uint16_t> A(128 * 64), Y(128 * 64), W(64 * 64);
    srand(7);
    auto rv = [] { return f2bf((float)((rand() % 9) - 4) * 0.25f); };
    for (auto &v : A) v = rv();
    for (auto &v : Y) v = rv();
    for (auto &v : W) v = rv();
    uint16_t *dA, *dY, *dW; float *dO;
    cudaMalloc(&dA, A.size() * 2); cudaMalloc(&dY, Y.size() * 2); cudaMalloc(&dW, W.size() * 2); cudaMalloc(&dO, 128 * 160 * 4);
    cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dY, Y.data(), Y.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dO, 0, 128 * 160 * 4);
    const int smem = 32768 + 8192 + 1024 + 64 + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(dA, dY, dW, dO);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 2; }
    std::vector<float> O(128 * 160);
    cudaMemcpy(O.data(), dO, O.size() * 4, cudaMemcpyDeviceToHost);
    auto a = [&](int r, int c) { return (double)bf2f(A[r * 64 + c]); };
    auto y = [&](int r, int c) { return (double)bf2f(Y[r * 64 + c]); };
    auto w = [&](int r, int c) { return (double)bf2f(W[r * 64 + c]); };
    auto lane64 = [](int m) { return (m & 15) + 32 * (m >> 4); };
    int bad[8] = {0};
    for (int e2 = 0; e2 < 128; ++e2)
        for (int n = 0; n < 32; ++n) {
            double r1 = 0, r2 = 0;
            for (int k = 0; k < 32; ++k) r1 += (a(e2, k) + a(e2, 32 + k)) * w(n, k);
            for (int o = 0; o < 32; ++o) r2 += a(e2, o) * w(o, n);
            if (O[e2 * 128 + n] != (float)r1) ++bad[1];
            if (O[e2 * 128 + 32 + n] != (float)r2) ++bad[2];
        }
    for (int m = 0; m < 64; ++m) {
        for (int n = 0; n < 32; ++n) {
            double r3 = 0, r4 = 0;
            for (int e2 = 0; e2 < 128; ++e2) { r3 += a(e2, m) * (y(e2, n) + y(e2, 32 + n)); r4 += a(e2, m) * y(e2, 32 + n); }
            if (O[lane64(m) * 128 + 64 + n] != (float)r3) ++bad[3];
            if (O[(lane64(m) + 16) * 128 + 64 + n] != (float)r4) ++bad[4];
        }
        for (int n = 0; n < 16; ++n) {
            double r5 = 0;
            for (int e2 = 0; e2 < 128; ++e2) r5 += a(e2, m) * y(e2, 16 + n);
            if (O[lane64(m) * 128 + 96 + n] != (float)r5) ++bad[5];
        }
        double r6 = 0;
        for (int e2 = 0; e2 < 128; ++e2) r6 += a(e2, m);
        for (int n = 0; n < 8; ++n)
            if (O[lane64(m) * 128 + 112 + n] != (float)r6) ++bad[6];
    }
    for (int e2 = 0; e2 < 128; ++e2)
        for (int n = 0; n < 32; ++n) {
            double r7 = 0;
            for (int k = 0; k < 32; ++k) r7 += a(e2, k) * w(n, k);
            if (O[128 * 128 + e2 * 32 + n] != (float)r7) ++bad[7];
        }
    for (int t = 1; t <= 7; ++t) printf("T%d %s (%d mismatches)\n", t, bad[t] ? "FAIL" : "PASS", bad[t]);
    if (argc > 1) {
        FILE *f = fopen(argv[1], "wb");
        if (f) {
            fwrite(A.data(), 2, A.size(), f); fwrite(Y.data(), 2, Y.size(), f); fwrite(W.data(), 2, W.size(), f);
            fwrite(O.data(), 4, O.size(), f);
            fclose(f);
        }
    }
    int tot = 0;
    for (int t = 1; t <= 7; ++t) tot += bad[t];
    return tot ? 1 : 0;
}
