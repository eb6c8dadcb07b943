// Shared device code of the E_GCL layer kernels (egnn_layer.cu: fp32 CUDA-core path,
// egnn_layer_tc.cu: tcgen05 tensor-core path).  Reference: src/3dmatch_train_egnn_with_batch.py:128-289.
#pragma once
#include "egspr_common.cuh"

namespace egspr {

constexpr int L_THREADS = 256;
constexpr int L_TILE = L_THREADS;   // edges per tile: one edge per thread
constexpr int L_ROW = 37;           // tile / accumulator row stride in floats (35 used); odd => conflict-free
constexpr int L_WFLOATS = EDGE_PART + NODE_PART + PQ_PART;   // 7072 floats of weights in smem

struct LayerArgs {
    const float *h, *x4, *P, *Q;
    const int32_t *csr_ptr, *csr_row, *csr_col, *csr_eid;
    const float *edge_attr;
    float edge_attr_const;
    int64_t num_nodes, edges_per_cloud;
    int n_per_cloud;
    const float *layer_pack, *next_pack, *out_pack;
    float *h_out, *x4_out, *x3_out, *P_out, *Q_out;
};

template <int NB>
constexpr size_t layer_smem_bytes() {
    return sizeof(float) * (L_WFLOATS + 32 + L_TILE * L_ROW + NB * L_ROW) + sizeof(int) * (NB + 4);
}

__device__ __forceinline__ void load_row32(float (&v)[32], const float *__restrict__ p) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = ldg4(p + 4 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
__device__ __forceinline__ void store_row32(float *__restrict__ p, const float (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4 *>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ void load_bias32(float (&v)[32], const float *__restrict__ sp) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = *reinterpret_cast<const float4 *>(sp + 4 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}

// acc[o] += sum_i wt[i][o] * vec[i], i < 32: wt [32][32] in shared memory (warp-uniform addresses ->
// broadcast LDS.128), vec = this thread's private row in shared memory.  Rolled x4 to keep the
// node phase small in the instruction cache.
__device__ __forceinline__ void matvec32_smem(float (&acc)[32], const float *__restrict__ wt,
                                              const float *__restrict__ vec) {
#pragma unroll 4
    for (int i = 0; i < 32; ++i) {
        const float v = vec[i];
#pragma unroll
        for (int o4 = 0; o4 < 8; ++o4) {
            const float4 w = *reinterpret_cast<const float4 *>(wt + 32 * i + 4 * o4);
            ffma2(acc[4 * o4 + 0], acc[4 * o4 + 1], w.x, w.y, v, v);
            ffma2(acc[4 * o4 + 2], acc[4 * o4 + 3], w.z, w.w, v, v);
        }
    }
}

// ---- node phase: one thread per node; `hrow` / `arow` are the thread's private 32-float rows in
// shared memory holding h[g] and agg[g] (both are overwritten) ---------------------------------------
__device__ __forceinline__ void node_update(const LayerArgs &a, int64_t g, float *__restrict__ hrow,
                                            float *__restrict__ arow, const float *__restrict__ sw) {
    float acc[32];
    load_row32(acc, a.h + g * H);
#pragma unroll
    for (int i = 0; i < 32; ++i) hrow[i] = acc[i];
    load_bias32(acc, sw + OFF_BN1);
    matvec32_smem(acc, sw + OFF_WN1T, hrow);                    // cat[h | agg]  :256
    matvec32_smem(acc, sw + OFF_WN1T + 32 * 32, arow);
#pragma unroll
    for (int o = 0; o < 32; ++o) arow[o] = silu(acc[o]);
    load_bias32(acc, sw + OFF_BN2);
    matvec32_smem(acc, sw + OFF_WN2T, arow);
#pragma unroll
    for (int o = 0; o < 32; ++o) { acc[o] += hrow[o]; hrow[o] = acc[o]; }   // residual :258-259
    if (a.next_pack) {
        store_row32(a.h_out + g * H, acc);
#pragma unroll
        for (int o = 0; o < 32; ++o) acc[o] = 0.f;
        matvec32_smem(acc, sw + OFF_WPT, hrow);
        store_row32(a.P_out + g * H, acc);
        load_bias32(acc, sw + OFF_BQ);
        matvec32_smem(acc, sw + OFF_WQT, hrow);
        store_row32(a.Q_out + g * H, acc);
    } else if (a.out_pack) {                                      // embedding_out :337
        load_bias32(acc, sw + OFF_WPT + 1024);
        matvec32_smem(acc, sw + OFF_WPT, hrow);
        store_row32(a.h_out + g * H, acc);
    } else {
        store_row32(a.h_out + g * H, acc);
    }
}

// ---- edge phase: one edge per thread, weights broadcast from shared memory -------------------
// geometry + first edge Linear (P[row] + Q[col] + Wgeo geo) + SiLU + per-head 8x8 Linear + LayerNorm
// for the edge at sorted position p: returns the message m in um[32] and coord_diff in (dx,dy,dz)
__device__ __forceinline__ void edge_message(const LayerArgs &a, const float *__restrict__ sw,
                                             const float *__restrict__ swea, int p, float (&um)[32],
                                             float &dx, float &dy, float &dz) {
    float geo[12];
    const int r = __ldg(a.csr_row + p), c = __ldg(a.csr_col + p);
    const float *Pr = a.P + (int64_t)r * H, *Qc = a.Q + (int64_t)c * H;
    float ea = a.edge_attr_const;
    if (a.edge_attr) {
        const int64_t cloud = r / a.n_per_cloud;
        ea = __ldg(a.edge_attr + cloud * a.edges_per_cloud + __ldg(a.csr_eid + p));
    }
    const float4 xr = ldg4(a.x4 + (int64_t)r * 4), xc = ldg4(a.x4 + (int64_t)c * 4);
    dx = xr.x - xc.x; dy = xr.y - xc.y; dz = xr.z - xc.z;                        // :273
    {
        const float radial = dx * dx + dy * dy + dz * dz;                        // :274
        const float dist = fast_sqrt(radial);                                    // :179
        const float ia = fast_rcp(dist + 1e-8f);                                 // :140
        float ax = dx * ia, ay = dy * ia, az = dz * ia;
        const float cx = xr.y * xc.z - xr.z * xc.y, cy = xr.z * xc.x - xr.x * xc.z,  // :143
                    cz = xr.x * xc.y - xr.y * xc.x;
        const float ib = fast_rcp(fast_sqrt(cx * cx + cy * cy + cz * cz) + 1e-8f);   // :144
        float bx = cx * ib, by = cy * ib, bz = cz * ib;
        float ex = ay * bz - az * by, ey = az * bx - ax * bz, ez = ax * by - ay * bx;  // :149
        const float na2 = ax * ax + ay * ay + az * az, nb2 = bx * bx + by * by + bz * bz,
                    nc2 = ex * ex + ey * ey + ez * ez;
        if (na2 < 1e-12f || nb2 < 1e-12f || nc2 < 1e-12f) {                      // norms < 1e-6  :152-163
            ax = 1.f; ay = 0.f; az = 0.f; bx = 0.f; by = 1.f; bz = 0.f; ex = 0.f; ey = 0.f; ez = 1.f;
        }
        geo[0] = radial; geo[1] = dist; geo[2] = xr.x * xc.x + xr.y * xc.y + xr.z * xc.z;   // :180
        // so3 flattened row-major with columns (a,b,c): [a0,b0,c0,a1,b1,c1,a2,b2,c2]  :159,:165
        geo[3] = ax; geo[4] = bx; geo[5] = ex;
        geo[6] = ay; geo[7] = by; geo[8] = ey;
        geo[9] = az; geo[10] = bz; geo[11] = ez;
    }
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) {
        float pre[8];
        {
            const float4 p0v = ldg4(Pr + 8 * hd), p1v = ldg4(Pr + 8 * hd + 4);
            const float4 q0v = ldg4(Qc + 8 * hd), q1v = ldg4(Qc + 8 * hd + 4);
            const float4 e0 = *reinterpret_cast<const float4 *>(swea + 8 * hd);
            const float4 e1 = *reinterpret_cast<const float4 *>(swea + 8 * hd + 4);
            pre[0] = p0v.x + fmaf(e0.x, ea, q0v.x); pre[1] = p0v.y + fmaf(e0.y, ea, q0v.y);
            pre[2] = p0v.z + fmaf(e0.z, ea, q0v.z); pre[3] = p0v.w + fmaf(e0.w, ea, q0v.w);
            pre[4] = p1v.x + fmaf(e1.x, ea, q1v.x); pre[5] = p1v.y + fmaf(e1.y, ea, q1v.y);
            pre[6] = p1v.z + fmaf(e1.z, ea, q1v.z); pre[7] = p1v.w + fmaf(e1.w, ea, q1v.w);
        }
#pragma unroll
        for (int g = 0; g < 12; ++g) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sw + OFF_WG + 32 * g + 8 * hd);
            const float4 w1 = *reinterpret_cast<const float4 *>(sw + OFF_WG + 32 * g + 8 * hd + 4);
            const float v = geo[g];
            pre[0] = fmaf(w0.x, v, pre[0]); pre[1] = fmaf(w0.y, v, pre[1]);
            pre[2] = fmaf(w0.z, v, pre[2]); pre[3] = fmaf(w0.w, v, pre[3]);
            pre[4] = fmaf(w1.x, v, pre[4]); pre[5] = fmaf(w1.y, v, pre[5]);
            pre[6] = fmaf(w1.z, v, pre[6]); pre[7] = fmaf(w1.w, v, pre[7]);
        }
        {
            const float4 b0 = *reinterpret_cast<const float4 *>(sw + OFF_B2 + 8 * hd);
            const float4 b1 = *reinterpret_cast<const float4 *>(sw + OFF_B2 + 8 * hd + 4);
#pragma unroll
            for (int i = 0; i < 8; ++i) pre[i] = silu(pre[i]);
            um[8 * hd + 0] = b0.x; um[8 * hd + 1] = b0.y; um[8 * hd + 2] = b0.z; um[8 * hd + 3] = b0.w;
            um[8 * hd + 4] = b1.x; um[8 * hd + 5] = b1.y; um[8 * hd + 6] = b1.z; um[8 * hd + 7] = b1.w;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sw + OFF_W2P + 64 * hd + 8 * i);
            const float4 w1 = *reinterpret_cast<const float4 *>(sw + OFF_W2P + 64 * hd + 8 * i + 4);
            const float v = pre[i];
            float *o = &um[8 * hd];
            o[0] = fmaf(w0.x, v, o[0]); o[1] = fmaf(w0.y, v, o[1]); o[2] = fmaf(w0.z, v, o[2]); o[3] = fmaf(w0.w, v, o[3]);
            o[4] = fmaf(w1.x, v, o[4]); o[5] = fmaf(w1.y, v, o[5]); o[6] = fmaf(w1.z, v, o[6]); o[7] = fmaf(w1.w, v, o[7]);
        }
    }
    // LayerNorm(32), eps 1e-5, biased variance (:209,:249)
    {
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) mean += um[j];
        mean *= (1.0f / 32.0f);
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float t = um[j] - mean; var = fmaf(t, t, var); }
        const float rstd = rsqrtf(var * (1.0f / 32.0f) + 1e-5f);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 gm = *reinterpret_cast<const float4 *>(sw + OFF_LNG + 4 * j4);
            const float4 bt = *reinterpret_cast<const float4 *>(sw + OFF_LNB + 4 * j4);
            um[4 * j4 + 0] = fmaf((um[4 * j4 + 0] - mean) * rstd, gm.x, bt.x);
            um[4 * j4 + 1] = fmaf((um[4 * j4 + 1] - mean) * rstd, gm.y, bt.y);
            um[4 * j4 + 2] = fmaf((um[4 * j4 + 2] - mean) * rstd, gm.z, bt.z);
            um[4 * j4 + 3] = fmaf((um[4 * j4 + 3] - mean) * rstd, gm.w, bt.w);
        }
    }
}

// SIMT edge phase: message -> tile row, then the coord MLP on CUDA cores
__device__ __forceinline__ void edge_phase(const LayerArgs &a, const float *__restrict__ sw,
                                           const float *__restrict__ swea, float *__restrict__ row,
                                           int p) {
    float um[32], dx, dy, dz;
    edge_message(a, sw, swea, p, um, dx, dy, dz);
#pragma unroll
    for (int j = 0; j < 32; ++j) row[j] = um[j];

    // coord MLP: s = wc2 . SiLU(Wc1 m + bc1)   (:219-229, no bias on the last Linear, tanh=False)
    float s = 0.f;
#pragma unroll 2
    for (int o = 0; o < 32; o += 4) {
        float t[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = sw[OFF_BC1 + o + q];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = *reinterpret_cast<const float4 *>(sw + OFF_WC1 + 32 * (o + q) + i);
                t[q] = fmaf(w.x, um[i], t[q]); t[q] = fmaf(w.y, um[i + 1], t[q]);
                t[q] = fmaf(w.z, um[i + 2], t[q]); t[q] = fmaf(w.w, um[i + 3], t[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) s = fmaf(sw[OFF_WC2 + o + q], silu(t[q]), s);
    }
    row[32] = dx * s; row[33] = dy * s; row[34] = dz * s;                               // trans = coord_diff * s  :264
}


}  // namespace egspr
