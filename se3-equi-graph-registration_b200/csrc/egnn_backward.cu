// Backward pass of the E_GCL / EGNN stack (the gradient `loss.backward()` produces in the reference's training
// step, src/3dmatch_train_egnn_with_batch.py:1125, through E_GCL.forward 3dm:280-289 and EGNN.forward 3dm:328-340).
//
// Three kernels per layer, all fp32 on CUDA cores, deterministic data path (no atomics on activations):
//   node_mlp_backward_kernel   thread = node: node_model (3dm:252-260) backward -> dh (direct part), dagg
//   edge_backward_kernel       thread = edge (row-CSR order): recomputes the edge's forward from the layer input
//                              (nothing per-edge is kept from the forward pass) and pushes (dagg[row], dx_out[row])
//                              back to dpre (= dP[row] = dQ[col]) and the two endpoints' coordinate gradients,
//                              written per ORIGINAL edge id
//   node_gather_backward_kernel thread = node: sums dpre / dx over the node's row list (row-CSR) and col list
//                              (col-CSR), then the P/Q halves of the first edge Linear back to dh
// Weight gradients: every kernel stages the rows of its outer products in shared memory per 128-row tile, each
// thread owns 8 (or 2 / 4) entries of a weight matrix in registers across all tiles of its CTA, column sums
// (biases, LayerNorm, wc2) are transposed-reduced with warp shuffles; one atomicAdd per entry and CTA at the end.
// The arithmetic itself lives in egnn_backward_math.cuh, which the CPU test-suite compiles with g++.
#include <cstdlib>

#include "egnn_backward_math.cuh"
#include "egspr_common.cuh"

namespace egspr {
using namespace bwd;

constexpr int BT = 128;    // threads per CTA = rows (edges / nodes) per tile
constexpr int RS = 33;     // stash row stride in floats: odd -> own-row writes and column reads are conflict-free
constexpr unsigned FULLM = 0xffffffffu;

// sum over the 32 lanes of v[lane'] for every column: returns, in lane L, sum over lanes of v[L]
__device__ __forceinline__ float warp_colsum32(const float (&v)[32]) {
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

// acc[q] += sum over the tile's rows of In[e][i] * Out[e][8 ob + q],  i = tid & 31, ob = tid >> 5
__device__ __forceinline__ void outer8(float (&acc)[8], const float *__restrict__ sIn, const float *__restrict__ sOut) {
    const int i = threadIdx.x & 31, ob = (threadIdx.x >> 5) * 8;
#pragma unroll 4
    for (int e = 0; e < BT; ++e) {
        const float a = sIn[e * RS + i];
        const float *o = sOut + e * RS + ob;
#pragma unroll
        for (int q = 0; q < 8; q += 2) fma2(acc[q], acc[q + 1], a, a, o[q], o[q + 1]);
    }
}
// in_major: gradient matrix stored [in][out] (transposed packs) else [out][in]
__device__ __forceinline__ void flush8(const float (&acc)[8], float *__restrict__ gp, bool in_major) {
    const int i = threadIdx.x & 31, ob = (threadIdx.x >> 5) * 8;
#pragma unroll
    for (int q = 0; q < 8; ++q) atomicAdd(gp + (in_major ? 32 * i + ob + q : 32 * (ob + q) + i), acc[q]);
}

__device__ __forceinline__ void stash_row(float *__restrict__ dst, const float (&v)[32]) {
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[j] = v[j];
}
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
__device__ __forceinline__ void add_row32g(float (&v)[32], const float *__restrict__ p) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = ldg4(p + 4 * i);
        v[4 * i] += t.x; v[4 * i + 1] += t.y; v[4 * i + 2] += t.z; v[4 * i + 3] += t.w;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// node MLP backward
// ---------------------------------------------------------------------------------------------------------------
constexpr int NB_W_LO = B_WN1T, NB_W_N = B_BN2 + 32 - B_WN1T;       // pack slice this kernel reads
constexpr size_t NB_SMEM = sizeof(float) * (NB_W_N + 5 * BT * RS);

__global__ void __launch_bounds__(BT) node_mlp_backward_kernel(const float *__restrict__ h, const float *__restrict__ agg,
                                                               const float *__restrict__ dh_out, int64_t G,
                                                               const float *__restrict__ pack, float *__restrict__ dh_in,
                                                               float *__restrict__ dagg, float *__restrict__ gpack) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                                   // pack[NB_W_LO, +NB_W_N)
    float *sA = smem + NB_W_N, *sDout = sA + BT * RS, *sH = sDout + BT * RS, *sAgg = sH + BT * RS, *sDz = sAgg + BT * RS;
    for (int i = threadIdx.x; i < NB_W_N; i += BT) sw[i] = __ldg(pack + NB_W_LO + i);
    __syncthreads();
    const float *w = sw - NB_W_LO;
    float accW2[8] = {0}, accW1h[8] = {0}, accW1a[8] = {0};
    float colDout = 0.f, colDz = 0.f;
    const int64_t tiles = (G + BT - 1) / BT;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t n = tile * BT + threadIdx.x;
        float *rA = sA + threadIdx.x * RS, *rD = sDout + threadIdx.x * RS, *rH = sH + threadIdx.x * RS,
              *rG = sAgg + threadIdx.x * RS, *rZ = sDz + threadIdx.x * RS;
        float dout[32], dz1[32];
        if (n < G) {
            {
                float t[32];
                load_row32g(t, h + n * H);
                stash_row(rH, t);
                load_row32g(t, agg + n * H);
                stash_row(rG, t);
            }
            load_row32g(dout, dh_out + n * H);
            stash_row(rD, dout);
            node_backward(w, rH, rG, dout, rD, rA, rZ, dh_in + n * H, dagg + n * H);
#pragma unroll
            for (int j = 0; j < 32; ++j) dz1[j] = rZ[j];
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { dout[j] = 0.f; dz1[j] = 0.f; rA[j] = 0.f; rH[j] = 0.f; rG[j] = 0.f; rZ[j] = 0.f; rD[j] = 0.f; }
        }
        colDout += warp_colsum32(dout);
        colDz += warp_colsum32(dz1);
        __syncthreads();
        outer8(accW2, sA, sDout);        // dWn2T[i][o] += a[i] dout[o]
        outer8(accW1h, sH, sDz);         // dWn1T[i][o] += h[i] dz1[o]
        outer8(accW1a, sAgg, sDz);       // dWn1T[32+i][o] += agg[i] dz1[o]
        __syncthreads();
    }
    flush8(accW2, gpack + B_WN2T, true);
    flush8(accW1h, gpack + B_WN1T, true);
    flush8(accW1a, gpack + B_WN1T + 1024, true);
    const int lane = threadIdx.x & 31;
    atomicAdd(gpack + B_BN2 + lane, colDout);
    atomicAdd(gpack + B_BN1 + lane, colDz);
}

constexpr int NT = 1024;            // threads of the 8-lanes-per-row kernels
constexpr int NOS = BT + 4;         // feature-major row stride of their "out" tiles

// ---------------------------------------------------------------------------------------------------------------
// edge backward
// ---------------------------------------------------------------------------------------------------------------
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
};

constexpr int EB_W_N = EDGE_PART + 32;     // edge part of the pack + the edge_attr column
// shared memory: weights | M, A1 edge-major [BT][RS] | DC1, DU, DPRE feature-major [32][OS] | geo [BT][GS (13 used)]
constexpr int GS = 17;                     // geo row stride (odd: conflict-free own-row writes)
constexpr int OS = BT + 4;                 // feature-major row stride: 16-byte aligned rows, and rows o, o+2, o+4, o+6 (or o, o+4) that one
                                           // 128-bit load instruction touches start 8 (16) banks apart
constexpr size_t EB_SMEM = sizeof(float) * (EB_W_N + 2 * BT * RS + 3 * 32 * OS + BT * GS);

struct DevSink {
    float colacc[C_COUNT];
    template <int ID>
    __host__ __device__ __forceinline__ void col(const float (&v)[32]) {
#ifdef __CUDA_ARCH__
        colacc[ID] += warp_colsum32(v);
#endif
    }
};

__global__ void __launch_bounds__(BT, 2) edge_backward_kernel(const EdgeBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                          // [0,1824) edge part, [1824,1856) edge_attr column
    float *sM = smem + EB_W_N, *sA1 = sM + BT * RS;                              // "in" rows of the outer products
    float *sDC1 = sA1 + BT * RS, *sDU = sDC1 + 32 * OS, *sDPRE = sDU + 32 * OS;  // "out" rows, feature-major
    float *sGeo = sDPRE + 32 * OS;
    for (int i = threadIdx.x; i < EDGE_PART; i += BT) sw[i] = __ldg(a.pack + i);
    if (threadIdx.x < 32) sw[EDGE_PART + threadIdx.x] = __ldg(a.pack + B_WEA + threadIdx.x);
    __syncthreads();
    const int tid = threadIdx.x;
    DevSink sink;
    float *rM = sM + tid * RS, *rA1 = sA1 + tid * RS, *rDC1 = sDC1 + tid, *rDU = sDU + tid, *rDPRE = sDPRE + tid,
          *rGeo = sGeo + tid * GS;
#pragma unroll
    for (int c = 0; c < C_COUNT; ++c) sink.colacc[c] = 0.f;
    float accWc1[16], accW2[4], accWg[8];
#pragma unroll
    for (int q = 0; q < 16; ++q) accWc1[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) accW2[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) accWg[q] = 0.f;
    const int64_t E = __ldg(a.csr_ptr + a.num_nodes);
    const int64_t tiles = (E + BT - 1) / BT;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t p0 = tile * BT + tid;
        const bool valid = p0 < E;
        const int64_t p = valid ? p0 : E - 1;          // idle threads redo the last edge with zero upstream gradient
        const int r = __ldg(a.csr_row + p), c = __ldg(a.csr_col + p);
        const int64_t cloud = r / a.n_per_cloud;
        const int64_t ge = cloud * a.edges_per_cloud + __ldg(a.csr_eid + p);
        const float ea = a.edge_attr ? __ldg(a.edge_attr + ge) : a.edge_attr_const;
        const float4 xr4 = ldg4(a.x4 + (int64_t)r * 4), xc4 = ldg4(a.x4 + (int64_t)c * 4);
        const float xr[3] = {xr4.x, xr4.y, xr4.z}, xc[3] = {xc4.x, xc4.y, xc4.z};
        {
            float pq[32];
            load_row32g(pq, a.P + (int64_t)r * H);
            add_row32g(pq, a.Q + (int64_t)c * H);
#pragma unroll
            for (int j = 0; j < 32; ++j) rDPRE[j * OS] = pq[j];
        }
        float dagg[32], dxo[3];
        if (valid) {
            load_row32g(dagg, a.dagg + (int64_t)r * H);
            dxo[0] = __ldg(a.dx_out + (int64_t)r * 3); dxo[1] = __ldg(a.dx_out + (int64_t)r * 3 + 1); dxo[2] = __ldg(a.dx_out + (int64_t)r * 3 + 2);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) dagg[j] = 0.f;
            dxo[0] = dxo[1] = dxo[2] = 0.f;
        }
        float dxr[3], dxc[3];
        edge_backward<OS>(sw, sw + EDGE_PART, xr, xc, ea, dagg, dxo, rM, rDC1, rA1, rDU, rDPRE, rGeo, sink, dxr, dxc);
        if (valid) {
            float dpre[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) dpre[j] = rDPRE[j * OS];
            store_row32g(a.dpre + ge * H, dpre);
            *reinterpret_cast<float4 *>(a.dxe + ge * 8) = make_float4(dxr[0], dxr[1], dxr[2], 0.f);
            *reinterpret_cast<float4 *>(a.dxe + ge * 8 + 4) = make_float4(dxc[0], dxc[1], dxc[2], 0.f);
        }
        __syncthreads();
        // ---- weight gradients of this tile: "in" rows read per edge, "out" rows four edges per 128-bit load ----
        {
            const int i = tid & 31, ob = (tid >> 5) * 8;                      // dWc1[o][i] += dc1[o] m[i], o = ob + q
            const float *in = sM + i;
#pragma unroll 2
            for (int e = 0; e < BT; e += 4) {
                const float m0 = in[e * RS], m1 = in[(e + 1) * RS], m2 = in[(e + 2) * RS], m3 = in[(e + 3) * RS];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 ov = *reinterpret_cast<const float4 *>(sDC1 + (ob + q) * OS + e);
                    fma2(accWc1[2 * q], accWc1[2 * q + 1], m0, m1, ov.x, ov.y);
                    fma2(accWc1[2 * q], accWc1[2 * q + 1], m2, m3, ov.z, ov.w);
                }
            }
        }
        {
            const int idx = 2 * tid, hb = (idx >> 6) * 8, ii = (idx >> 3) & 7, oo = idx & 7;   // dW2P[hd][i][o] += a1 du
            const float *in = sA1 + hb + ii;
            const float *o0 = sDU + (hb + oo) * OS, *o1 = o0 + OS;
#pragma unroll 2
            for (int e = 0; e < BT; e += 4) {
                const float a0 = in[e * RS], a1 = in[(e + 1) * RS], a2 = in[(e + 2) * RS], a3 = in[(e + 3) * RS];
                const float4 u = *reinterpret_cast<const float4 *>(o0 + e), v = *reinterpret_cast<const float4 *>(o1 + e);
                fma2(accW2[0], accW2[1], a0, a1, u.x, u.y); fma2(accW2[0], accW2[1], a2, a3, u.z, u.w);
                fma2(accW2[2], accW2[3], a0, a1, v.x, v.y); fma2(accW2[2], accW2[3], a2, a3, v.z, v.w);
            }
        }
        {
            const int k = tid & 15, ob = (tid >> 4) * 4;                      // dWg[k][o] += geo[k] dpre[o], o = ob + q
            if (k < 13) {
                const float *in = sGeo + k;
#pragma unroll 2
                for (int e = 0; e < BT; e += 4) {
                    const float g0 = in[e * GS], g1 = in[(e + 1) * GS], g2 = in[(e + 2) * GS], g3 = in[(e + 3) * GS];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ov = *reinterpret_cast<const float4 *>(sDPRE + (ob + q) * OS + e);
                        fma2(accWg[2 * q], accWg[2 * q + 1], g0, g1, ov.x, ov.y);
                        fma2(accWg[2 * q], accWg[2 * q + 1], g2, g3, ov.z, ov.w);
                    }
                }
            }
        }
        __syncthreads();
    }
    {
        const int i = tid & 31, ob = (tid >> 5) * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(a.gpack + B_WC1 + 32 * (ob + q) + i, accWc1[2 * q] + accWc1[2 * q + 1]);
    }
    atomicAdd(a.gpack + B_W2P + 2 * tid, accW2[0] + accW2[1]);
    atomicAdd(a.gpack + B_W2P + 2 * tid + 1, accW2[2] + accW2[3]);
    {
        const int k = tid & 15, ob = (tid >> 4) * 4;
        if (k < 13) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                atomicAdd(a.gpack + (k < 12 ? B_WG + 32 * k : B_WEA) + ob + q, accWg[2 * q] + accWg[2 * q + 1]);
        }
    }
    const int lane = tid & 31;
    atomicAdd(a.gpack + B_WC2 + lane, sink.colacc[C_DWC2]);
    atomicAdd(a.gpack + B_BC1 + lane, sink.colacc[C_DBC1]);
    atomicAdd(a.gpack + B_LNB + lane, sink.colacc[C_DLNB]);
    atomicAdd(a.gpack + B_LNG + lane, sink.colacc[C_DLNG]);
    atomicAdd(a.gpack + B_B2 + lane, sink.colacc[C_DB2]);
}

// ---------------------------------------------------------------------------------------------------------------
// edge backward, two threads per edge ("half-edge" kernel, the default)
// ---------------------------------------------------------------------------------------------------------------
// The one-thread-per-edge kernel above is capped at 8 warps per SM (253 registers, 692 bytes of tile rows per edge) and
// is latency-bound there.  Here lanes l and l+16 of a warp share an edge: each computes 16 of the 32 channels of every
// stage (2 of the 4 heads -- the second edge Linear is block-diagonal, so each half needs only its own heads), the
// LayerNorm statistics, the coord-MLP scalar and the 12 geometric gradients are completed with one xor-16 shuffle, and
// the two halves exchange m / dc1 through the edge's shared-memory rows (__syncwarp).  Same arithmetic as
// edge_backward() in egnn_backward_math.cuh (which the CPU suite checks), half the registers and twice the warps per
// byte of shared memory: 12 warps per SM.  Tile = 64 edges per 128 threads; the tile reduction is unchanged.
constexpr int HE = BT / 2;                 // edges per tile
constexpr int HOS = HE + 4;                // feature-major row stride
constexpr size_t EH_SMEM = sizeof(float) * (EB_W_N + 2 * HE * RS + 3 * 32 * HOS + HE * GS);

// sum over the 16 lanes of a half-warp of v[lane'] for every column: returns, in half-lane j, sum of v[j]
__device__ __forceinline__ float halfwarp_colsum16(const float (&v)[16]) {
    const int lane = threadIdx.x & 31;
    float b[8], c[4], d[2];
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float keep = up ? v[j + 8] : v[j], send = up ? v[j] : v[j + 8];
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
__device__ __forceinline__ float pair_sum(float v) { return v + __shfl_xor_sync(FULLM, v, 16); }

__global__ void __launch_bounds__(BT, 3) edge_backward_half_kernel(const EdgeBwdArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;                          // [0,1824) edge part, [1824,1856) edge_attr column
    float *sM = smem + EB_W_N, *sA1 = sM + HE * RS;
    float *sDC1 = sA1 + HE * RS, *sDU = sDC1 + 32 * HOS, *sDPRE = sDU + 32 * HOS;
    float *sGeo = sDPRE + 32 * HOS;
    for (int i = threadIdx.x; i < EDGE_PART; i += BT) sw[i] = __ldg(a.pack + i);
    if (threadIdx.x < 32) sw[EDGE_PART + threadIdx.x] = __ldg(a.pack + B_WEA + threadIdx.x);
    __syncthreads();
    const float *w = sw, *wea = sw + EDGE_PART;
    const int tid = threadIdx.x, lane = tid & 31;
    const int hf = (lane >> 4) & 1, ob0 = 16 * hf;                 // this thread's channels: ob0 .. ob0 + 15
    const int el = (tid >> 5) * 16 + (lane & 15);                  // edge of the tile
    float *rM = sM + el * RS, *rA1 = sA1 + el * RS, *rDC1 = sDC1 + el, *rDU = sDU + el, *rDPRE = sDPRE + el,
          *rGeo = sGeo + el * GS;
    float colacc[C_COUNT];
#pragma unroll
    for (int c = 0; c < C_COUNT; ++c) colacc[c] = 0.f;
    float accWc1[16], accW2[4], accWg[8];
#pragma unroll
    for (int q = 0; q < 16; ++q) accWc1[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) accW2[q] = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) accWg[q] = 0.f;
    const int64_t E = __ldg(a.csr_ptr + a.num_nodes);
    const int64_t tiles = (E + HE - 1) / HE;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t p0 = tile * HE + el;
        const bool valid = p0 < E;
        const int64_t p = valid ? p0 : E - 1;          // idle pairs redo the last edge with zero upstream gradient
        const int r = __ldg(a.csr_row + p), c = __ldg(a.csr_col + p);
        const int64_t cloud = r / a.n_per_cloud;
        const int64_t ge = cloud * a.edges_per_cloud + __ldg(a.csr_eid + p);
        const float ea = a.edge_attr ? __ldg(a.edge_attr + ge) : a.edge_attr_const;
        const float4 xr4 = ldg4(a.x4 + (int64_t)r * 4), xc4 = ldg4(a.x4 + (int64_t)c * 4);
        const float xr[3] = {xr4.x, xr4.y, xr4.z}, xc[3] = {xc4.x, xc4.y, xc4.z};
        float dxo[3] = {0.f, 0.f, 0.f};
        if (valid) {
            dxo[0] = __ldg(a.dx_out + (int64_t)r * 3); dxo[1] = __ldg(a.dx_out + (int64_t)r * 3 + 1); dxo[2] = __ldg(a.dx_out + (int64_t)r * 3 + 2);
        }
        // geometry (both halves, identical) ------------------------------------------------------------------------
        EdgeGeo g;
        float gk[13];
        edge_geometry(xr, xc, g, gk);
        gk[12] = ea;
        if (hf == 0) {
#pragma unroll
            for (int k = 0; k < 13; ++k) rGeo[k] = gk[k];
        }
        // first edge Linear + SiLU, own 16 channels.  rDPRE: d silu / d pre ---------------------------------------------
#pragma unroll 2
        for (int o4 = ob0; o4 < ob0 + 16; o4 += 4) {
            const float4 pv = ldg4(a.P + (int64_t)r * H + o4), qv = ldg4(a.Q + (int64_t)c * H + o4);
            const F4 we = ld4(wea + o4);
            float p0v = fmaf(we.x, ea, pv.x + qv.x), p1v = fmaf(we.y, ea, pv.y + qv.y), p2v = fmaf(we.z, ea, pv.z + qv.z),
                  p3v = fmaf(we.w, ea, pv.w + qv.w);
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const F4 wv = ld4(w + B_WG + 32 * k + o4);
                fma2(p0v, p1v, wv.x, wv.y, gk[k], gk[k]); fma2(p2v, p3v, wv.z, wv.w, gk[k], gk[k]);
            }
            const float pr[4] = {p0v, p1v, p2v, p3v};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float sg = sigmoidf_(pr[q]);
                rA1[o4 + q] = pr[q] * sg;
                rDPRE[(o4 + q) * HOS] = sg * (1.0f + pr[q] * (1.0f - sg));
            }
        }
        // second edge Linear (own two heads) + LayerNorm ------------------------------------------------------------
        float uh[16];
#pragma unroll
        for (int o4 = 0; o4 < 16; o4 += 4) {
            const F4 b = ld4(w + B_B2 + ob0 + o4);
            uh[o4] = b.x; uh[o4 + 1] = b.y; uh[o4 + 2] = b.z; uh[o4 + 3] = b.w;
        }
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
                const float av = rA1[ob0 + 8 * hd + i];
                const float *wp = w + B_W2P + 64 * (2 * hf + hd) + 8 * i;
                const F4 w0 = ld4(wp), w1 = ld4(wp + 4);
                float *u = uh + 8 * hd;
                fma2(u[0], u[1], w0.x, w0.y, av, av); fma2(u[2], u[3], w0.z, w0.w, av, av);
                fma2(u[4], u[5], w1.x, w1.y, av, av); fma2(u[6], u[7], w1.z, w1.w, av, av);
            }
        }
        float mean = 0.f;
#pragma unroll
        for (int o = 0; o < 16; ++o) mean += uh[o];
        mean = pair_sum(mean) * (1.0f / 32.0f);
        float var = 0.f;
#pragma unroll
        for (int o = 0; o < 16; ++o) { const float t = uh[o] - mean; var = fmaf(t, t, var); }
        var = pair_sum(var);
        const float rstd = 1.0f / sqrtf(var * (1.0f / 32.0f) + 1e-5f);
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            uh[o] = (uh[o] - mean) * rstd;
            rM[ob0 + o] = fmaf(uh[o], w[B_LNG + ob0 + o], w[B_LNB + ob0 + o]);
        }
        __syncwarp();
        // coord MLP, own 16 outputs over the full message ---------------------------------------------------------------
        const float dsc = g.d[0] * dxo[0] + g.d[1] * dxo[1] + g.d[2] * dxo[2];   // d loss / d s
        float s = 0.f;
        {
            float m[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) m[i] = rM[i];
#pragma unroll 2
            for (int o = ob0; o < ob0 + 16; ++o) {
                float c0 = w[B_BC1 + o], c1 = 0.f, c2 = 0.f, c3 = 0.f, c4 = 0.f, c5 = 0.f, c6 = 0.f, c7 = 0.f;
#pragma unroll
                for (int i = 0; i < 32; i += 8) {
                    const F4 w0 = ld4(w + B_WC1 + 32 * o + i), w1 = ld4(w + B_WC1 + 32 * o + i + 4);
                    fma2(c0, c1, w0.x, w0.y, m[i], m[i + 1]); fma2(c2, c3, w0.z, w0.w, m[i + 2], m[i + 3]);
                    fma2(c4, c5, w1.x, w1.y, m[i + 4], m[i + 5]); fma2(c6, c7, w1.z, w1.w, m[i + 6], m[i + 7]);
                }
                c0 += c2; c1 += c3; c4 += c6; c5 += c7; c0 += c4; c1 += c5;
                const float cc = c0 + c1;
                const float sg = sigmoidf_(cc);
                const float a2 = cc * sg;
                const float wc = w[B_WC2 + o];
                s = fmaf(wc, a2, s);
                rDU[o * HOS] = a2 * dsc;                                         // scratch: rows of the wc2 gradient
                rDC1[o * HOS] = wc * dsc * (sg * (1.0f + cc * (1.0f - sg)));
            }
        }
        s = pair_sum(s);
        {
            float v[16];
#pragma unroll
            for (int o = 0; o < 16; ++o) v[o] = rDU[(ob0 + o) * HOS];
            colacc[C_DWC2] += halfwarp_colsum16(v);
#pragma unroll
            for (int o = 0; o < 16; ++o) v[o] = rDC1[(ob0 + o) * HOS];
            colacc[C_DBC1] += halfwarp_colsum16(v);
        }
        __syncwarp();
        // message gradient, own 16 channels, from all 32 coord-MLP rows -----------------------------------------------
        float dm[16];
        if (valid) {
#pragma unroll
            for (int i4 = 0; i4 < 16; i4 += 4) {
                const float4 t = ldg4(a.dagg + (int64_t)r * H + ob0 + i4);
                dm[i4] = t.x; dm[i4 + 1] = t.y; dm[i4 + 2] = t.z; dm[i4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) dm[i] = 0.f;
        }
#pragma unroll 2
        for (int o = 0; o < 32; ++o) {
            const float dc = rDC1[o * HOS];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
                const F4 wv = ld4(w + B_WC1 + 32 * o + ob0 + i);
                fma2(dm[i], dm[i + 1], wv.x, wv.y, dc, dc); fma2(dm[i + 2], dm[i + 3], wv.z, wv.w, dc, dc);
            }
        }
        colacc[C_DLNB] += halfwarp_colsum16(dm);
        // LayerNorm backward ------------------------------------------------------------------------------------------
        {
            float dg[16];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                dg[i] = dm[i] * uh[i];
                dm[i] *= w[B_LNG + ob0 + i];
                s1 += dm[i];
                s2 = fmaf(dm[i], uh[i], s2);
            }
            colacc[C_DLNG] += halfwarp_colsum16(dg);
            s1 = pair_sum(s1) * (1.0f / 32.0f);
            s2 = pair_sum(s2) * (1.0f / 32.0f);
#pragma unroll
            for (int i = 0; i < 16; ++i) { dm[i] = rstd * (dm[i] - s1 - uh[i] * s2); rDU[(ob0 + i) * HOS] = dm[i]; }   // dm is now du
        }
        colacc[C_DB2] += halfwarp_colsum16(dm);
        // second Linear backward (own heads) + SiLU backward.  rDPRE: d silu / d pre -> dpre ---------------------------
#pragma unroll
        for (int hd = 0; hd < 2; ++hd) {
#pragma unroll 2
            for (int i = 0; i < 8; ++i) {
                const float *wp = w + B_W2P + 64 * (2 * hf + hd) + 8 * i;
                const F4 w0 = ld4(wp), w1 = ld4(wp + 4);
                const float *du = dm + 8 * hd;
                float t0 = 0.f, t1 = 0.f;
                fma2(t0, t1, w0.x, w0.y, du[0], du[1]); fma2(t0, t1, w0.z, w0.w, du[2], du[3]);
                fma2(t0, t1, w1.x, w1.y, du[4], du[5]); fma2(t0, t1, w1.z, w1.w, du[6], du[7]);
                rDPRE[(ob0 + 8 * hd + i) * HOS] *= (t0 + t1);
            }
        }
        // geometry backward: partial over own channels, completed across the pair ---------------------------------------
        float gg[12], gh[12];
#pragma unroll
        for (int k = 0; k < 12; ++k) { gg[k] = 0.f; gh[k] = 0.f; }
        float dpre[16];
#pragma unroll
        for (int o = 0; o < 16; ++o) dpre[o] = rDPRE[(ob0 + o) * HOS];
#pragma unroll
        for (int o4 = 0; o4 < 16; o4 += 4) {
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                const F4 wv = ld4(w + B_WG + 32 * k + ob0 + o4);
                fma2(gg[k], gh[k], wv.x, wv.y, dpre[o4], dpre[o4 + 1]); fma2(gg[k], gh[k], wv.z, wv.w, dpre[o4 + 2], dpre[o4 + 3]);
            }
        }
#pragma unroll
        for (int k = 0; k < 12; ++k) gg[k] = pair_sum(gg[k] + gh[k]);
        const float gde[3] = {s * dxo[0], s * dxo[1], s * dxo[2]};
        float dxr[3], dxc[3];
        edge_geometry_backward(xr, xc, g, gg, gde, dxr, dxc);
        if (valid) {
#pragma unroll
            for (int o4 = 0; o4 < 16; o4 += 4)
                *reinterpret_cast<float4 *>(a.dpre + ge * H + ob0 + o4) = make_float4(dpre[o4], dpre[o4 + 1], dpre[o4 + 2], dpre[o4 + 3]);
            if (hf == 0) *reinterpret_cast<float4 *>(a.dxe + ge * 8) = make_float4(dxr[0], dxr[1], dxr[2], 0.f);
            else *reinterpret_cast<float4 *>(a.dxe + ge * 8 + 4) = make_float4(dxc[0], dxc[1], dxc[2], 0.f);
        }
        __syncthreads();
        // ---- weight gradients of this tile (as in edge_backward_kernel, 64 edges) ----
        {
            const int i = tid & 31, ob = (tid >> 5) * 8;                      // dWc1[o][i] += dc1[o] m[i], o = ob + q
            const float *in = sM + i;
#pragma unroll 2
            for (int e = 0; e < HE; e += 4) {
                const float m0 = in[e * RS], m1 = in[(e + 1) * RS], m2 = in[(e + 2) * RS], m3 = in[(e + 3) * RS];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const float4 ov = *reinterpret_cast<const float4 *>(sDC1 + (ob + q) * HOS + e);
                    fma2(accWc1[2 * q], accWc1[2 * q + 1], m0, m1, ov.x, ov.y);
                    fma2(accWc1[2 * q], accWc1[2 * q + 1], m2, m3, ov.z, ov.w);
                }
            }
        }
        {
            const int idx = 2 * tid, hb = (idx >> 6) * 8, ii = (idx >> 3) & 7, oo = idx & 7;   // dW2P[hd][i][o] += a1 du
            const float *in = sA1 + hb + ii;
            const float *o0 = sDU + (hb + oo) * HOS, *o1 = o0 + HOS;
#pragma unroll 2
            for (int e = 0; e < HE; e += 4) {
                const float a0 = in[e * RS], a1 = in[(e + 1) * RS], a2 = in[(e + 2) * RS], a3 = in[(e + 3) * RS];
                const float4 u = *reinterpret_cast<const float4 *>(o0 + e), v = *reinterpret_cast<const float4 *>(o1 + e);
                fma2(accW2[0], accW2[1], a0, a1, u.x, u.y); fma2(accW2[0], accW2[1], a2, a3, u.z, u.w);
                fma2(accW2[2], accW2[3], a0, a1, v.x, v.y); fma2(accW2[2], accW2[3], a2, a3, v.z, v.w);
            }
        }
        {
            const int k = tid & 15, ob = (tid >> 4) * 4;                      // dWg[k][o] += geo[k] dpre[o], o = ob + q
            if (k < 13) {
                const float *in = sGeo + k;
#pragma unroll 2
                for (int e = 0; e < HE; e += 4) {
                    const float g0 = in[e * GS], g1 = in[(e + 1) * GS], g2 = in[(e + 2) * GS], g3 = in[(e + 3) * GS];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 ov = *reinterpret_cast<const float4 *>(sDPRE + (ob + q) * HOS + e);
                        fma2(accWg[2 * q], accWg[2 * q + 1], g0, g1, ov.x, ov.y);
                        fma2(accWg[2 * q], accWg[2 * q + 1], g2, g3, ov.z, ov.w);
                    }
                }
            }
        }
        __syncthreads();
    }
    {
        const int i = tid & 31, ob = (tid >> 5) * 8;
#pragma unroll
        for (int q = 0; q < 8; ++q) atomicAdd(a.gpack + B_WC1 + 32 * (ob + q) + i, accWc1[2 * q] + accWc1[2 * q + 1]);
    }
    atomicAdd(a.gpack + B_W2P + 2 * tid, accW2[0] + accW2[1]);
    atomicAdd(a.gpack + B_W2P + 2 * tid + 1, accW2[2] + accW2[3]);
    {
        const int k = tid & 15, ob = (tid >> 4) * 4;
        if (k < 13) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                atomicAdd(a.gpack + (k < 12 ? B_WG + 32 * k : B_WEA) + ob + q, accWg[2 * q] + accWg[2 * q + 1]);
        }
    }
    // column sums: half-lane j of half hf holds column 16 hf + j = lane
    atomicAdd(a.gpack + B_WC2 + lane, colacc[C_DWC2]);
    atomicAdd(a.gpack + B_BC1 + lane, colacc[C_DBC1]);
    atomicAdd(a.gpack + B_LNB + lane, colacc[C_DLNB]);
    atomicAdd(a.gpack + B_LNG + lane, colacc[C_DLNG]);
    atomicAdd(a.gpack + B_B2 + lane, colacc[C_DB2]);
}

// ---------------------------------------------------------------------------------------------------------------
// node gather + P/Q backward
// ---------------------------------------------------------------------------------------------------------------
struct GatherArgs {
    const float *h;
    const int32_t *csr_ptr, *csr_eid, *csc_ptr, *csc_eid;
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
            const int64_t ebase = (n / a.n_per_cloud) * a.edges_per_cloud;
            for (int p = __ldg(a.csr_ptr + n), pe = __ldg(a.csr_ptr + n + 1); p < pe; p += 8) {   // edges with row == n
                const int mine = p + sub < pe ? __ldg(a.csr_eid + p + sub) : -1;
                float4 t[8];
                float tx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ej = __shfl_sync(gmask, mine, j, 8);
                    const int64_t ge = ebase + (ej < 0 ? 0 : ej);
                    t[j] = ej < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4(a.dpre + ge * H + 4 * sub);
                    tx[j] = (ej < 0 || sub >= 3) ? 0.f : __ldg(a.dxe + ge * 8 + sub);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) { dp.x += t[j].x; dp.y += t[j].y; dp.z += t[j].z; dp.w += t[j].w; dxa += tx[j]; }
            }
            for (int p = __ldg(a.csc_ptr + n), pe = __ldg(a.csc_ptr + n + 1); p < pe; p += 8) {   // edges with col == n
                const int mine = p + sub < pe ? __ldg(a.csc_eid + p + sub) : -1;
                float4 t[8];
                float tx[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int ej = __shfl_sync(gmask, mine, j, 8);
                    const int64_t ge = ebase + (ej < 0 ? 0 : ej);
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

constexpr size_t LB_SMEM = sizeof(float) * (1024 + 2 * BT * RS);

__global__ void __launch_bounds__(BT) linear32_backward_kernel(const float *__restrict__ x, const float *__restrict__ dy,
                                                               int64_t rows, const float *__restrict__ pack,
                                                               float *__restrict__ dx, float *__restrict__ gpack) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem, *sX = smem + 1024, *sDy = sX + BT * RS;
    for (int i = threadIdx.x; i < 1024; i += BT) sw[i] = __ldg(pack + i);
    __syncthreads();
    float acc[8] = {0};
    float colDy = 0.f;
    const int64_t tiles = (rows + BT - 1) / BT;
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t n = tile * BT + threadIdx.x;
        float *rX = sX + threadIdx.x * RS, *rD = sDy + threadIdx.x * RS;
        float dyv[32];
        if (n < rows) {
            float xv[32];
            load_row32g(xv, x + n * H);
            load_row32g(dyv, dy + n * H);
            stash_row(rX, xv);
            if (dx) linear32_backward_input(sw, dyv, dx + n * H, false);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) { dyv[j] = 0.f; rX[j] = 0.f; }
        }
        stash_row(rD, dyv);
        colDy += warp_colsum32(dyv);
        __syncthreads();
        outer8(acc, sX, sDy);       // dWT[i][o] += x[i] dy[o]
        __syncthreads();
    }
    flush8(acc, gpack, true);
    atomicAdd(gpack + 1024 + (threadIdx.x & 31), colDy);
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

extern "C" size_t egspr_egcl_backward_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
    if (num_nodes <= 0 || num_edges < 0) return 0;
    return sizeof(float) * ((size_t)num_nodes * 32 + (size_t)num_edges * 40) + 256;
}

extern "C" int egspr_egcl_backward(const float *h, const float *x4, const float *P, const float *Q, const float *agg,
                                   const int32_t *csr_ptr, const int32_t *csr_row, const int32_t *csr_col,
                                   const int32_t *csr_eid, const int32_t *csc_ptr, const int32_t *csc_eid,
                                   const float *edge_attr, float edge_attr_const, int64_t num_nodes,
                                   int64_t edges_per_cloud, int n_per_cloud, const float *layer_pack,
                                   const float *dh_out, const float *dx_out, float *dh_in, float *dx_in,
                                   float *grad_pack, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace egspr;
    if (!h || !x4 || !P || !Q || !agg || !csr_ptr || !csr_row || !csr_col || !csr_eid || !csc_ptr || !csc_eid ||
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
    static int occ_node = 0, occ_edge = 0, occ_gather = 0, occ_half = 0;
    const bool full_thread = getenv("EGSPR_EDGE_BWD_FULL") != nullptr;     // switch (read per call): the one-thread-per-edge / per-node kernels
    if (!occ_node) {
        if (int e = prep_kernel(node_mlp_backward_kernel, NB_SMEM, occ_node)) return e;
        if (int e = prep_kernel(edge_backward_kernel, EB_SMEM, occ_edge)) return e;
        if (int e = prep_kernel(edge_backward_half_kernel, EH_SMEM, occ_half)) return e;
        if (int e = prep_kernel(node_gather_backward_kernel, GB_SMEM, occ_gather, GT)) return e;
    }
    const int64_t ntiles = (num_nodes + BT - 1) / BT, etiles = (E + BT - 1) / BT;
    node_mlp_backward_kernel<<<grid_for(ntiles, occ_node), BT, NB_SMEM, st>>>(h, agg, dh_out, num_nodes, layer_pack, dh_in, dagg, grad_pack);
    EGSPR_CHECK_LAUNCH();
    EdgeBwdArgs ea{x4, P, Q, csr_ptr, csr_row, csr_col, csr_eid, edge_attr, edge_attr_const, num_nodes, edges_per_cloud,
                   n_per_cloud, layer_pack, dagg, dx_out, dpre, dxe, grad_pack};
    if (full_thread) edge_backward_kernel<<<grid_for(etiles, occ_edge), BT, EB_SMEM, st>>>(ea);
    else edge_backward_half_kernel<<<grid_for((E + HE - 1) / HE, occ_half), BT, EH_SMEM, st>>>(ea);
    EGSPR_CHECK_LAUNCH();
    GatherArgs ga{h, csr_ptr, csr_eid, csc_ptr, csc_eid, num_nodes, edges_per_cloud, n_per_cloud, layer_pack, dpre, dxe,
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
    static int occ = 0, occ_wide = 0;
    if (!occ) {
        if (int e = prep_kernel(linear32_backward_kernel, LB_SMEM, occ)) return e;
        if (int e = prep_kernel(linear32_backward_wide_kernel, LW_SMEM, occ_wide, NT)) return e;
    }
    const int64_t tiles = (rows + BT - 1) / BT;
    if (getenv("EGSPR_EDGE_BWD_FULL") != nullptr)
        linear32_backward_kernel<<<grid_for(tiles, occ), BT, LB_SMEM, (cudaStream_t)stream>>>(x, dy, rows, embed_pack, dx, grad_pack);
    else
        linear32_backward_wide_kernel<<<grid_for(tiles, occ_wide), NT, LW_SMEM, (cudaStream_t)stream>>>(x, dy, rows, embed_pack, dx, grad_pack);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}
