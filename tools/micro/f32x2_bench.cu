// Micro-benchmark (developer tool): issue/throughput of FFMA vs FFMA2, FADD vs FADD2 on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void add2(float &a0, float &a1, float b0, float b1) {
    asm volatile("{\n.reg .b64 ra, rb;\nmov.b64 ra, {%0, %1};\nmov.b64 rb, {%2, %3};\nadd.rn.f32x2 ra, ra, rb;\nmov.b64 {%0, %1}, ra;\n}" : "+f"(a0), "+f"(a1) : "f"(b0), "f"(b1));
}
__device__ __forceinline__ void fma2(float &d0, float &d1, float a0, float a1, float b0, float b1) {
    asm volatile("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2, %3};\nmov.b64 rb, {%4, %5};\nmov.b64 rd, {%0, %1};\nfma.rn.f32x2 rd, ra, rb, rd;\nmov.b64 {%0, %1}, rd;\n}" : "+f"(d0), "+f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
template <int MODE>
__global__ void k(float *o, int iters, float s) {
    float v[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
            if (MODE == 0) { v[i] = fmaf(v[i], s, 0.5f); v[i + 1] = fmaf(v[i + 1], s, 0.5f); }
            if (MODE == 1) { fma2(v[i], v[i + 1], v[i], v[i + 1], s, s); }
            if (MODE == 2) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i]) : "f"(s)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v[i + 1]) : "f"(s)); }
            if (MODE == 3) { add2(v[i], v[i + 1], s, s); }
            if (MODE == 4) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i])); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i + 1])); }
            if (MODE == 5) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i])); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i + 1])); }
        }
    }
    float t = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) t += v[i];
    o[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> void run(const char *name, int threads) {
    float *o; cudaMalloc(&o, 148 * 1024 * 4 * 4);
    const int iters = 4096;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    k<MODE><<<148, threads>>>(o, 16, 1.0001f);
    cudaEventRecord(a);
    k<MODE><<<148, threads>>>(o, iters, 1.0001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double ops = 148.0 * threads * iters * 16;   // scalar element-ops
    printf("%-8s threads/SM %4d: %.3f ms  %.1f elem-ops/clk/SM (at 1.9 GHz)\n", name, threads, ms, ops / (ms * 1e-3) / 148 / 1.9e9);
    cudaFree(o);
}
int main() {
    for (int t : {128, 512, 1024}) {
        run<0>("FFMA", t); run<1>("FFMA2", t); run<2>("FADD", t); run<3>("FADD2", t); run<4>("EX2", t); run<5>("TANH", t);
    }
    return 0;
}
