// Fused E_GCL layer (src/3dmatch_train_egnn_with_batch.py:185-289) for all clouds of a batch.
//
// One persistent launch per layer.  A work item is a block of L_NB consecutive global nodes; the
// CTA walks the block's incoming-edge list (row-major CSR, csr.cu) in tiles of L_THREADS*EPT
// edges, one edge per thread-slot:
//   edge phase   : geometry (coord2radial :271-278, compute_edge_features :176-181,
//                  compute_so3_matrix :128-173) -> first edge Linear as P[row] + Q[col] + Wgeo*geo
//                  (algebraic split of the 77-wide concat :238-242) -> SiLU -> per-head 8x8 Linear
//                  (:202-208, :245-246) -> LayerNorm (:249) -> coord MLP (:219-229, :264)
//   reduce phase : the tile's messages m_e [32] and d_e*s_e [3] are summed per row in shared memory
//                  and registers, in ascending edge order -- no atomics (:254, :265; SUM not mean)
//   node phase   : node MLP + residual (:252-260), coordinate update (:267), then the NEXT layer's
//                  P/Q halves (or embedding_out :337 after the last layer) while h' is in registers
// so per layer the only HBM/L2 traffic is the gathers and one write of h', x', P', Q'.
#include "egspr_common.cuh"

namespace egspr {

constexpr int L_THREADS = 256;
constexpr int L_NB = 64;            // nodes per work item (must be a multiple of L_THREADS/32)
constexpr int L_ROW = 37;           // tile row stride (35 floats used); odd => conflict-free rows
constexpr int L_AGG = 33;           // stride of the per-node aggregate rows handed to the node phase
constexpr int L_WFLOATS = EDGE_PART + NODE_PART + PQ_PART;   // 7072 floats of weights in smem
constexpr int L_QPW = L_NB / (L_THREADS / 32);               // nodes reduced per warp (8)

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

template <int EPT>
constexpr size_t layer_smem_bytes() {
    return sizeof(float) * (L_WFLOATS + L_THREADS * EPT * L_ROW + 32) + sizeof(int) * (L_NB + 4);
}

// ---- node phase: one thread per node -----------------------------------------------------------
__device__ __forceinline__ void matvec32_acc(float (&acc)[32], const float *__restrict__ wt, float v) {
    // acc[o] += wt[o] * v for o = 0..31, wt contiguous in shared memory (warp-uniform address)
#pragma unroll
    for (int o4 = 0; o4 < 8; ++o4) {
        const float4 w = *reinterpret_cast<const float4 *>(wt + 4 * o4);
        acc[4 * o4 + 0] = fmaf(w.x, v, acc[4 * o4 + 0]);
        acc[4 * o4 + 1] = fmaf(w.y, v, acc[4 * o4 + 1]);
        acc[4 * o4 + 2] = fmaf(w.z, v, acc[4 * o4 + 2]);
        acc[4 * o4 + 3] = fmaf(w.w, v, acc[4 * o4 + 3]);
    }
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

// writes P,Q (first edge Linear halves of the layer whose PQ part sits at sw+OFF_WPT) for one node
__device__ __forceinline__ void emit_pq(const float (&hn)[32], const float *__restrict__ sw,
                                        float *__restrict__ Pg, float *__restrict__ Qg) {
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) matvec32_acc(acc, sw + OFF_WPT + 32 * i, hn[i]);
    store_row32(Pg, acc);
    load_bias32(acc, sw + OFF_BQ);
#pragma unroll
    for (int i = 0; i < 32; ++i) matvec32_acc(acc, sw + OFF_WQT + 32 * i, hn[i]);
    store_row32(Qg, acc);
}

__device__ __forceinline__ void node_update(const LayerArgs &a, int64_t g, const float *__restrict__ agg_s,
                                            const float *__restrict__ sw) {
    float hv[32], hid[32];
    load_row32(hv, a.h + g * H);
    load_bias32(hid, sw + OFF_BN1);
#pragma unroll
    for (int i = 0; i < 32; ++i) matvec32_acc(hid, sw + OFF_WN1T + 32 * i, hv[i]);           // cat[h | agg] :256
#pragma unroll
    for (int i = 0; i < 32; ++i) matvec32_acc(hid, sw + OFF_WN1T + 32 * (32 + i), agg_s[i]);
#pragma unroll
    for (int o = 0; o < 32; ++o) hid[o] = silu(hid[o]);
    float out[32];
    load_bias32(out, sw + OFF_BN2);
#pragma unroll
    for (int i = 0; i < 32; ++i) matvec32_acc(out, sw + OFF_WN2T + 32 * i, hid[i]);
#pragma unroll
    for (int o = 0; o < 32; ++o) out[o] += hv[o];                                              // residual :258-259
    if (a.next_pack) {
        store_row32(a.h_out + g * H, out);
        emit_pq(out, sw, a.P_out + g * H, a.Q_out + g * H);
    } else if (a.out_pack) {                                                                   // embedding_out :337
        float ho[32];
        load_bias32(ho, sw + OFF_WPT + 1024);
#pragma unroll
        for (int i = 0; i < 32; ++i) matvec32_acc(ho, sw + OFF_WPT + 32 * i, out[i]);
        store_row32(a.h_out + g * H, ho);
    } else {
        store_row32(a.h_out + g * H, out);
    }
}

// ---- edge phase: EPT edges per thread, weights broadcast from shared memory -------------------
template <int EPT>
__device__ __forceinline__ void edge_phase(const LayerArgs &a, const float *__restrict__ sw,
                                           const float *__restrict__ swea, float *__restrict__ tile,
                                           int p0, int pend, int tid) {
    float geo[EPT][12], dv[EPT][3], um[EPT][32], ea[EPT];
    const float *Pr[EPT], *Qc[EPT];
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        int p = p0 + u * L_THREADS + tid;
        if (p >= pend) p = pend - 1;               // idle slot: recompute the last edge, never reduced
        const int r = __ldg(a.csr_row + p), c = __ldg(a.csr_col + p);
        Pr[u] = a.P + (int64_t)r * H;
        Qc[u] = a.Q + (int64_t)c * H;
        ea[u] = a.edge_attr_const;
        if (a.edge_attr) {
            const int64_t cloud = r / a.n_per_cloud;
            ea[u] = __ldg(a.edge_attr + cloud * a.edges_per_cloud + __ldg(a.csr_eid + p));
        }
        const float4 xr = ldg4(a.x4 + (int64_t)r * 4), xc = ldg4(a.x4 + (int64_t)c * 4);
        const float dx = xr.x - xc.x, dy = xr.y - xc.y, dz = xr.z - xc.z;            // :273
        const float radial = dx * dx + dy * dy + dz * dz;                            // :274
        const float dist = sqrtf(radial);                                            // :179
        const float dot = xr.x * xc.x + xr.y * xc.y + xr.z * xc.z;                   // :180
        const float ia = 1.0f / (dist + 1e-8f);                                      // :140
        float ax = dx * ia, ay = dy * ia, az = dz * ia;
        const float cx = xr.y * xc.z - xr.z * xc.y, cy = xr.z * xc.x - xr.x * xc.z,  // :143
                    cz = xr.x * xc.y - xr.y * xc.x;
        const float ib = 1.0f / (sqrtf(cx * cx + cy * cy + cz * cz) + 1e-8f);        // :144
        float bx = cx * ib, by = cy * ib, bz = cz * ib;
        float ex = ay * bz - az * by, ey = az * bx - ax * bz, ez = ax * by - ay * bx;  // :149
        const float na = sqrtf(ax * ax + ay * ay + az * az), nb = sqrtf(bx * bx + by * by + bz * bz),
                    nc = sqrtf(ex * ex + ey * ey + ez * ez);
        if (na < 1e-6f || nb < 1e-6f || nc < 1e-6f) {                                // :152-163
            ax = 1.f; ay = 0.f; az = 0.f; bx = 0.f; by = 1.f; bz = 0.f; ex = 0.f; ey = 0.f; ez = 1.f;
        }
        dv[u][0] = dx; dv[u][1] = dy; dv[u][2] = dz;
        geo[u][0] = radial; geo[u][1] = dist; geo[u][2] = dot;
        // so3 flattened row-major with columns (a,b,c): [a0,b0,c0,a1,b1,c1,a2,b2,c2]  :159,:165
        geo[u][3] = ax; geo[u][4] = bx; geo[u][5] = ex;
        geo[u][6] = ay; geo[u][7] = by; geo[u][8] = ey;
        geo[u][9] = az; geo[u][10] = bz; geo[u][11] = ez;
    }
#pragma unroll
    for (int hd = 0; hd < 4; ++hd) {
        float pre[EPT][8];
#pragma unroll
        for (int u = 0; u < EPT; ++u) {
            const float4 p0v = ldg4(Pr[u] + 8 * hd), p1v = ldg4(Pr[u] + 8 * hd + 4);
            const float4 q0v = ldg4(Qc[u] + 8 * hd), q1v = ldg4(Qc[u] + 8 * hd + 4);
            const float4 e0 = *reinterpret_cast<const float4 *>(swea + 8 * hd);
            const float4 e1 = *reinterpret_cast<const float4 *>(swea + 8 * hd + 4);
            pre[u][0] = p0v.x + fmaf(e0.x, ea[u], q0v.x); pre[u][1] = p0v.y + fmaf(e0.y, ea[u], q0v.y);
            pre[u][2] = p0v.z + fmaf(e0.z, ea[u], q0v.z); pre[u][3] = p0v.w + fmaf(e0.w, ea[u], q0v.w);
            pre[u][4] = p1v.x + fmaf(e1.x, ea[u], q1v.x); pre[u][5] = p1v.y + fmaf(e1.y, ea[u], q1v.y);
            pre[u][6] = p1v.z + fmaf(e1.z, ea[u], q1v.z); pre[u][7] = p1v.w + fmaf(e1.w, ea[u], q1v.w);
        }
#pragma unroll
        for (int g = 0; g < 12; ++g) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sw + OFF_WG + 32 * g + 8 * hd);
            const float4 w1 = *reinterpret_cast<const float4 *>(sw + OFF_WG + 32 * g + 8 * hd + 4);
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const float v = geo[u][g];
                pre[u][0] = fmaf(w0.x, v, pre[u][0]); pre[u][1] = fmaf(w0.y, v, pre[u][1]);
                pre[u][2] = fmaf(w0.z, v, pre[u][2]); pre[u][3] = fmaf(w0.w, v, pre[u][3]);
                pre[u][4] = fmaf(w1.x, v, pre[u][4]); pre[u][5] = fmaf(w1.y, v, pre[u][5]);
                pre[u][6] = fmaf(w1.z, v, pre[u][6]); pre[u][7] = fmaf(w1.w, v, pre[u][7]);
            }
        }
        {
            const float4 b0 = *reinterpret_cast<const float4 *>(sw + OFF_B2 + 8 * hd);
            const float4 b1 = *reinterpret_cast<const float4 *>(sw + OFF_B2 + 8 * hd + 4);
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
#pragma unroll
                for (int i = 0; i < 8; ++i) pre[u][i] = silu(pre[u][i]);
                um[u][8 * hd + 0] = b0.x; um[u][8 * hd + 1] = b0.y; um[u][8 * hd + 2] = b0.z; um[u][8 * hd + 3] = b0.w;
                um[u][8 * hd + 4] = b1.x; um[u][8 * hd + 5] = b1.y; um[u][8 * hd + 6] = b1.z; um[u][8 * hd + 7] = b1.w;
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 w0 = *reinterpret_cast<const float4 *>(sw + OFF_W2P + 64 * hd + 8 * i);
            const float4 w1 = *reinterpret_cast<const float4 *>(sw + OFF_W2P + 64 * hd + 8 * i + 4);
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                const float v = pre[u][i];
                float *o = &um[u][8 * hd];
                o[0] = fmaf(w0.x, v, o[0]); o[1] = fmaf(w0.y, v, o[1]); o[2] = fmaf(w0.z, v, o[2]); o[3] = fmaf(w0.w, v, o[3]);
                o[4] = fmaf(w1.x, v, o[4]); o[5] = fmaf(w1.y, v, o[5]); o[6] = fmaf(w1.z, v, o[6]); o[7] = fmaf(w1.w, v, o[7]);
            }
        }
    }
    // LayerNorm(32), eps 1e-5, biased variance (:209,:249)
#pragma unroll
    for (int u = 0; u < EPT; ++u) {
        float mean = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) mean += um[u][j];
        mean *= (1.0f / 32.0f);
        float var = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j) { const float t = um[u][j] - mean; var = fmaf(t, t, var); }
        const float rstd = rsqrtf(var * (1.0f / 32.0f) + 1e-5f);
        float *row = tile + (u * L_THREADS + tid) * L_ROW;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const float m = fmaf((um[u][j] - mean) * rstd, sw[OFF_LNG + j], sw[OFF_LNB + j]);
            um[u][j] = m;
            row[j] = m;
        }
    }
    // coord MLP: s = wc2 . SiLU(Wc1 m + bc1)   (:219-229, no bias on the last Linear, tanh=False)
    float s[EPT];
#pragma unroll
    for (int u = 0; u < EPT; ++u) s[u] = 0.f;
#pragma unroll 4
    for (int o = 0; o < 32; o += 2) {
        float t0[EPT], t1[EPT];
#pragma unroll
        for (int u = 0; u < EPT; ++u) { t0[u] = sw[OFF_BC1 + o]; t1[u] = sw[OFF_BC1 + o + 1]; }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            const float4 wa = *reinterpret_cast<const float4 *>(sw + OFF_WC1 + 32 * o + i);
            const float4 wb = *reinterpret_cast<const float4 *>(sw + OFF_WC1 + 32 * (o + 1) + i);
#pragma unroll
            for (int u = 0; u < EPT; ++u) {
                t0[u] = fmaf(wa.x, um[u][i], t0[u]); t1[u] = fmaf(wb.x, um[u][i], t1[u]);
                t0[u] = fmaf(wa.y, um[u][i + 1], t0[u]); t1[u] = fmaf(wb.y, um[u][i + 1], t1[u]);
                t0[u] = fmaf(wa.z, um[u][i + 2], t0[u]); t1[u] = fmaf(wb.z, um[u][i + 2], t1[u]);
                t0[u] = fmaf(wa.w, um[u][i + 3], t0[u]); t1[u] = fmaf(wb.w, um[u][i + 3], t1[u]);
            }
        }
        const float c0 = sw[OFF_WC2 + o], c1 = sw[OFF_WC2 + o + 1];
#pragma unroll
        for (int u = 0; u < EPT; ++u) s[u] = fmaf(c1, silu(t1[u]), fmaf(c0, silu(t0[u]), s[u]));
    }
#pragma unroll
    for (int u = 0; u < EPT; ++u) {                                                   // trans = coord_diff * s  :264
        float *row = tile + (u * L_THREADS + tid) * L_ROW;
        row[32] = dv[u][0] * s[u]; row[33] = dv[u][1] * s[u]; row[34] = dv[u][2] * s[u];
    }
}

template <int EPT>
__global__ void __launch_bounds__(L_THREADS) egcl_layer_kernel(const LayerArgs a) {
    extern __shared__ __align__(16) float smem[];
    float *sw = smem;
    float *tile = smem + L_WFLOATS;
    int *sptr = reinterpret_cast<int *>(tile + L_THREADS * EPT * L_ROW + 32);
    constexpr int TILE = L_THREADS * EPT;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // weights -> shared memory: this layer's edge + node parts, the NEXT layer's P/Q part (or the
    // embedding_out pack) for the node phase, and this layer's own edge_attr column for the edge phase
    float *swea = tile + L_THREADS * EPT * L_ROW;   // 32 floats
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
    __syncthreads();

    const int64_t G = a.num_nodes;
    const int64_t items = (G + L_NB - 1) / L_NB;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t n0 = item * L_NB;
        const int nb = (G - n0 < L_NB) ? (int)(G - n0) : L_NB;
        __syncthreads();   // previous item's node phase has finished reading the aggregate rows
        if (tid <= nb) sptr[tid] = __ldg(a.csr_ptr + n0 + tid);
        __syncthreads();
        const int pbeg = sptr[0], pend = sptr[nb];
        float acc[L_QPW], accx[L_QPW];
#pragma unroll
        for (int q = 0; q < L_QPW; ++q) { acc[q] = 0.f; accx[q] = 0.f; }
        for (int p0 = pbeg; p0 < pend; p0 += TILE) {
            edge_phase<EPT>(a, sw, swea, tile, p0, pend, tid);
            __syncthreads();
#pragma unroll
            for (int q = 0; q < L_QPW; ++q) {
                const int nl = warp + (L_THREADS / 32) * q;
                if (nl < nb) {
                    const int lo = max(sptr[nl], p0), hi = min(sptr[nl + 1], p0 + TILE);
                    for (int p = lo; p < hi; ++p) {
                        const float *row = tile + (p - p0) * L_ROW;
                        acc[q] += row[lane];
                        if (lane < 3) accx[q] += row[32 + lane];
                    }
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int q = 0; q < L_QPW; ++q) {
            const int nl = warp + (L_THREADS / 32) * q;
            if (nl < nb) {
                tile[nl * L_AGG + lane] = acc[q];                                      // agg = sum m  :254
                const int64_t g = n0 + nl;
                float xv = 0.f;
                if (lane < 3) xv = __ldg(a.x4 + g * 4 + lane) + accx[q];               // coord + agg  :267
                if (lane < 4) a.x4_out[g * 4 + lane] = xv;
                if (a.x3_out && lane < 3) a.x3_out[g * 3 + lane] = xv;
            }
        }
        __syncthreads();
        if (tid < nb) node_update(a, n0 + tid, tile + tid * L_AGG, sw);
    }
}

// ---- embedding_in + layer-0 P/Q (EGNN.forward :332), one thread per node -----------------------
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
#pragma unroll
            for (int i = 0; i < 32; ++i) matvec32_acc(hv, se + 32 * i, f[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) hv[i] = f[i];
        }
        store_row32(h + g * H, hv);
        emit_pq(hv, spq, P + g * H, Q + g * H);
        if (x4) {
            *reinterpret_cast<float4 *>(x4 + g * 4) =
                make_float4(__ldg(x3 + g * 3), __ldg(x3 + g * 3 + 1), __ldg(x3 + g * 3 + 2), 0.f);
        }
    }
}

template <int EPT>
static int launch_layer(const LayerArgs &a, cudaStream_t st) {
    static bool configured = false;
    static int ctas_per_sm = 1;
    constexpr size_t smem = layer_smem_bytes<EPT>();
    if (!configured) {
        if (cudaFuncSetAttribute(egcl_layer_kernel<EPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return EGSPR_E_LAUNCH;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, egcl_layer_kernel<EPT>, L_THREADS, smem) != cudaSuccess ||
            ctas_per_sm < 1)
            ctas_per_sm = 1;
        configured = true;
    }
    const int64_t items = (a.num_nodes + L_NB - 1) / L_NB;
    int64_t grid = (int64_t)sm_count() * ctas_per_sm;
    if (grid > items) grid = items;
    egcl_layer_kernel<EPT><<<(unsigned)grid, L_THREADS, smem, st>>>(a);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

}  // namespace egspr

extern "C" int egspr_node_embed(const float *feat, const float *x3, int64_t num_nodes, const float *embed_pack,
                                const float *layer0_pack, float *h, float *x4, float *P, float *Q, void *stream) {
    using namespace egspr;
    if (!feat || !layer0_pack || !h || !P || !Q || num_nodes <= 0) return EGSPR_E_INVALID;
    if (x4 && !x3) return EGSPR_E_INVALID;
    int64_t grid = (num_nodes + 127) / 128;
    if (grid > (int64_t)sm_count() * 8) grid = (int64_t)sm_count() * 8;
    node_embed_kernel<<<(unsigned)grid, 128, 0, (cudaStream_t)stream>>>(feat, x3, num_nodes, embed_pack, layer0_pack, h, x4, P, Q);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_egcl_forward(const float *h, const float *x4, const float *P, const float *Q,
                                  const int32_t *csr_ptr, const int32_t *csr_row, const int32_t *csr_col,
                                  const int32_t *csr_eid, const float *edge_attr, float edge_attr_const,
                                  int64_t num_nodes, int64_t edges_per_cloud, int n_per_cloud,
                                  const float *layer_pack, const float *next_pack, const float *out_pack,
                                  float *h_out, float *x4_out, float *x3_out, float *P_out, float *Q_out, int impl,
                                  void *stream) {
    using namespace egspr;
    if (!h || !x4 || !P || !Q || !csr_ptr || !csr_row || !csr_col || !layer_pack || !h_out || !x4_out ||
        num_nodes <= 0 || n_per_cloud <= 0)
        return EGSPR_E_INVALID;
    if (edge_attr && !csr_eid) return EGSPR_E_INVALID;
    if (next_pack && (!P_out || !Q_out)) return EGSPR_E_INVALID;
    if (h_out == h || x4_out == x4) return EGSPR_E_INVALID;   // other CTAs still gather the layer input
    LayerArgs a{h, x4, P, Q, csr_ptr, csr_row, csr_col, csr_eid, edge_attr, edge_attr_const, num_nodes,
                edges_per_cloud, n_per_cloud, layer_pack, next_pack, out_pack, h_out, x4_out, x3_out, P_out, Q_out};
    switch (impl) {
        case 0: return launch_layer<2>(a, (cudaStream_t)stream);
        case 1: return launch_layer<1>(a, (cudaStream_t)stream);
        default: return EGSPR_E_UNSUPPORTED;
    }
}
