// E_GCL edge kernel with ALL THREE per-edge contractions on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM) -- impl 4 of egspr_egcl_forward.
//
// Reference arithmetic: src/3dmatch_train_egnn_with_batch.py:231-250 (edge_model), :262-268
// (coord_model), :252-254 (segment sums of node_model); SURVEY Appendix A.2.
//
// Per tile of 128 edges (one edge per thread, edges in CSR order = grouped by aggregation row):
//   stage 1  pre = P[row] + Q[col] + [geo(12) | edge_attr] Wg^T       M128 N32 K16   (first edge Linear;
//            the h[row]/h[col] blocks were folded per node into P, Q -- egspr_node_embed / node kernel)
//   stage 2  u   = SiLU(pre) W2^T + b2, W2 = block-diag of the 4 heads' (8x8) second Linear   K32
//            m   = LayerNorm_32(u)
//   stage 3  s   = wc2 . SiLU(m Wc1^T + bc1)                          K32   (coord_mlp)
// Each stage: threads write their row of the A operand into a 128B-swizzled K-major shared-memory
// tile, one elected thread issues the tcgen05.mma's, the accumulator comes back with tcgen05.ld
// (32x32b: thread t <- row t).  The three stages are serially dependent, so they share ONE pair of
// A tiles and ONE 32-column TMEM accumulator.
//
// fp32 parity: 3xTF32 split (x = hi + lo, hi = x with the low 13 mantissa bits cleared, lo exact):
//   A W^T ~= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T, fp32 accumulation in TMEM.
// Stage 1 (K = 13 <= 16) packs [hi | lo] into one K=32 row: 6 MMAs; stages 2, 3: 12 MMAs each.
//
// CTA = 2 independent halves of 128 threads (named barriers) that share the weight tiles; each half
// walks its own work items (M_NB consecutive aggregation rows).  Per-node sums are continued in
// ascending edge order across tiles (bit-reproducible; duplicate points stay bit-identical).
#include <cstdio>
#include <cstdlib>

#include "egnn_layer.cuh"
#include "tcgen05.cuh"

namespace egspr {
using namespace tc;

constexpr int M_HALVES = 2;
constexpr int M_THREADS = 128 * M_HALVES;
constexpr int M_NB = 64;             // aggregation rows (nodes) per work item
constexpr int M_ROW = 35;            // accumulator row: 32 features + 3 coordinates

// shared-memory carve-up (bytes from a 1024-aligned base)
constexpr int MS_A = 0;                              // per half: hi tile (16 KB) + lo tile (16 KB)
constexpr int MS_W = MS_A + M_HALVES * 32768;        // X1 | W2hi | W2lo | W3hi | W3lo, 4 KB each
constexpr int MS_PAR = MS_W + 5 * 4096;              // b2, ln gamma, ln beta, bc1, wc2 (32 floats each)
constexpr int MS_HALF = MS_PAR + 5 * 128;
constexpr int MH_DXS = 0;                            // float4[128]: coord_diff * s of the tile
constexpr int MH_SACC = MH_DXS + 128 * 16;
constexpr int MH_SPTR = MH_SACC + M_NB * M_ROW * 4;
constexpr int MH_MBAR = ((MH_SPTR + (M_NB + 1) * 4 + 7) / 8) * 8;
constexpr int MH_SIZE = ((MH_MBAR + 8 + 15) / 16) * 16;
constexpr int MS_TMEM = MS_HALF + M_HALVES * MH_SIZE;
constexpr int MS_END = MS_TMEM + 16;
constexpr size_t M_SMEM_BYTES = MS_END + 1024;       // + slack for the manual 1024-byte alignment

// write 32 values as the hi / lo rows of this thread in the two A tiles
__device__ __forceinline__ void store_hilo(uint8_t *Ahi, uint8_t *Alo, int row, const float (&v)[32]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        float4 hi, lo;
        hi.x = tf32_hi(v[4 * c]); hi.y = tf32_hi(v[4 * c + 1]); hi.z = tf32_hi(v[4 * c + 2]); hi.w = tf32_hi(v[4 * c + 3]);
        lo.x = v[4 * c] - hi.x; lo.y = v[4 * c + 1] - hi.y; lo.z = v[4 * c + 2] - hi.z; lo.w = v[4 * c + 3] - hi.w;
        const int off = row * 128 + ((c ^ (row & 7)) << 4);
        *reinterpret_cast<float4 *>(Ahi + off) = hi;
        *reinterpret_cast<float4 *>(Alo + off) = lo;
    }
}

// D = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T over K = 32 (4 K-blocks of 8)
__device__ __forceinline__ void issue_3xtf32_k32(uint32_t d, uint64_t ahi, uint64_t alo, uint64_t whi, uint64_t wlo) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32(d, ahi + 2 * k, whi + 2 * k, IDESC_TF32_M128_N32, k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32(d, alo + 2 * k, whi + 2 * k, IDESC_TF32_M128_N32, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32(d, ahi + 2 * k, wlo + 2 * k, IDESC_TF32_M128_N32, 1);
}

__global__ void __launch_bounds__(M_THREADS, 2) egcl_edge_mma_kernel(const LayerArgs a, float *__restrict__ agg_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, half = tid >> 7, ht = tid & 127, lane = tid & 31, hw = ht >> 5;
    uint8_t *Ahi = base + MS_A + half * 32768, *Alo = Ahi + 16384;
    const float *sb2 = reinterpret_cast<const float *>(base + MS_PAR);
    const float *slng = sb2 + 32, *slnb = sb2 + 64, *sbc1 = sb2 + 96, *swc2 = sb2 + 128;
    uint8_t *hb = base + MS_HALF + half * MH_SIZE;
    float4 *dxs = reinterpret_cast<float4 *>(hb + MH_DXS);
    float *sacc = reinterpret_cast<float *>(hb + MH_SACC);
    int *sptr = reinterpret_cast<int *>(hb + MH_SPTR);
    const uint32_t mbar = smem_u32(hb + MH_MBAR);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + MS_TMEM);
    const int bar_id = 1 + half;

    // ---- one-time setup: swizzled hi/lo weight tiles (B operands: row = output o, K = input) ----
    for (int i = tid; i < 1024; i += M_THREADS) {
        const int o = i >> 5, k = i & 31;
        {   // stage 1: K 0..15 = hi of [Wg(12) | w_edge_attr | 0 0 0], K 16..31 = lo of the same
            const int kk = k & 15;
            float w = 0.f;
            if (kk < 12) w = __ldg(a.layer_pack + OFF_WG + 32 * kk + o);
            else if (kk == 12) w = __ldg(a.layer_pack + OFF_WEA + o);
            const float hi = tf32_hi(w);
            *reinterpret_cast<float *>(base + MS_W + sw128_off(o, k)) = (k < 16) ? hi : (w - hi);
        }
        {   // stage 2: block-diagonal of the heads' second Linear, pack layout [head][in][out]
            const float w = ((o >> 3) == (k >> 3)) ? __ldg(a.layer_pack + OFF_W2P + 64 * (o >> 3) + 8 * (k & 7) + (o & 7)) : 0.f;
            const float hi = tf32_hi(w);
            *reinterpret_cast<float *>(base + MS_W + 4096 + sw128_off(o, k)) = hi;
            *reinterpret_cast<float *>(base + MS_W + 8192 + sw128_off(o, k)) = w - hi;
        }
        {   // stage 3: coord_mlp.0.weight [out][in]
            const float w = __ldg(a.layer_pack + OFF_WC1 + i);
            const float hi = tf32_hi(w);
            *reinterpret_cast<float *>(base + MS_W + 12288 + sw128_off(o, k)) = hi;
            *reinterpret_cast<float *>(base + MS_W + 16384 + sw128_off(o, k)) = w - hi;
        }
    }
    if (tid < 32) {
        float *par = reinterpret_cast<float *>(base + MS_PAR);
        par[tid] = __ldg(a.layer_pack + OFF_B2 + tid);
        par[32 + tid] = __ldg(a.layer_pack + OFF_LNG + tid);
        par[64 + tid] = __ldg(a.layer_pack + OFF_LNB + tid);
        par[96 + tid] = __ldg(a.layer_pack + OFF_BC1 + tid);
        par[128 + tid] = __ldg(a.layer_pack + OFF_WC2 + tid);
    }
    if (tid < 32) tmem_alloc(smem_u32(tmem_holder), 32 * M_HALVES);
    if (ht == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();            // the weight tiles were written through the generic proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_d = *tmem_holder + 32u * half;
    const uint32_t tmem_row = tmem_d + ((uint32_t)(hw * 32) << 16);     // this warp's 32 TMEM lanes
    const uint64_t dAhi = make_desc_sw128(smem_u32(Ahi)), dAlo = make_desc_sw128(smem_u32(Alo));
    const uint64_t dX1 = make_desc_sw128(smem_u32(base + MS_W));
    const uint64_t dW2hi = make_desc_sw128(smem_u32(base + MS_W + 4096)), dW2lo = make_desc_sw128(smem_u32(base + MS_W + 8192));
    const uint64_t dW3hi = make_desc_sw128(smem_u32(base + MS_W + 12288)), dW3lo = make_desc_sw128(smem_u32(base + MS_W + 16384));
    uint32_t phase = 0;

    const int64_t G = a.num_nodes;
    const int64_t items = (G + M_NB - 1) / M_NB;
    for (int64_t item = (int64_t)blockIdx.x * M_HALVES + half; item < items; item += (int64_t)gridDim.x * M_HALVES) {
        const int64_t n0 = item * M_NB;
        const int nb = (G - n0 < M_NB) ? (int)(G - n0) : M_NB;
        bar_sync(bar_id, 128);       // the previous item's accumulators have been written out
        for (int i = ht; i <= nb; i += 128) sptr[i] = __ldg(a.csr_ptr + n0 + i);
        for (int i = ht; i < M_NB * M_ROW; i += 128) sacc[i] = 0.f;
        bar_sync(bar_id, 128);
        const int pbeg = sptr[0], pend = sptr[nb];
        int ncur = 0;
        for (int p0 = pbeg; p0 < pend; p0 += 128) {
            int p = p0 + ht;
            if (p >= pend) p = pend - 1;               // idle slot: recompute the last edge, never reduced
            const int r = __ldg(a.csr_row + p), c = __ldg(a.csr_col + p);
            float dx, dy, dz;
            // ---------------- stage 1 operand: geometry (:271-278, :128-181) ----------------
            {
                float geo[16];
                const float4 xr = ldg4(a.x4 + (int64_t)r * 4), xc = ldg4(a.x4 + (int64_t)c * 4);
                float ea = a.edge_attr_const;
                if (a.edge_attr) {
                    const int64_t cloud = r / a.n_per_cloud;
                    ea = __ldg(a.edge_attr + cloud * a.edges_per_cloud + __ldg(a.csr_eid + p));
                }
                dx = xr.x - xc.x; dy = xr.y - xc.y; dz = xr.z - xc.z;                        // :273
                const float radial = dx * dx + dy * dy + dz * dz;                            // :274
                const float dist = fast_sqrt(radial);                                        // :179
                const float ia = fast_rcp(dist + 1e-8f);                                     // :140
                float ax = dx * ia, ay = dy * ia, az = dz * ia;
                const float cx = xr.y * xc.z - xr.z * xc.y, cy = xr.z * xc.x - xr.x * xc.z,  // :143
                            cz = xr.x * xc.y - xr.y * xc.x;
                const float ib = fast_rcp(fast_sqrt(cx * cx + cy * cy + cz * cz) + 1e-8f);   // :144
                float bx = cx * ib, by = cy * ib, bz = cz * ib;
                float ex = ay * bz - az * by, ey = az * bx - ax * bz, ez = ax * by - ay * bx;  // :149
                const float na2 = ax * ax + ay * ay + az * az, nb2 = bx * bx + by * by + bz * bz,
                            nc2 = ex * ex + ey * ey + ez * ez;
                if (na2 < 1e-12f || nb2 < 1e-12f || nc2 < 1e-12f) {                          // norms < 1e-6  :152-163
                    ax = 1.f; ay = 0.f; az = 0.f; bx = 0.f; by = 1.f; bz = 0.f; ex = 0.f; ey = 0.f; ez = 1.f;
                }
                geo[0] = radial; geo[1] = dist; geo[2] = xr.x * xc.x + xr.y * xc.y + xr.z * xc.z;   // :180
                geo[3] = ax; geo[4] = bx; geo[5] = ex;      // so3 row-major, columns (a,b,c)  :159,:165
                geo[6] = ay; geo[7] = by; geo[8] = ey;
                geo[9] = az; geo[10] = bz; geo[11] = ez;
                geo[12] = ea; geo[13] = 0.f; geo[14] = 0.f; geo[15] = 0.f;
#pragma unroll
                for (int cch = 0; cch < 4; ++cch) {
                    float4 hi, lo;
                    hi.x = tf32_hi(geo[4 * cch]); hi.y = tf32_hi(geo[4 * cch + 1]); hi.z = tf32_hi(geo[4 * cch + 2]); hi.w = tf32_hi(geo[4 * cch + 3]);
                    lo.x = geo[4 * cch] - hi.x; lo.y = geo[4 * cch + 1] - hi.y; lo.z = geo[4 * cch + 2] - hi.z; lo.w = geo[4 * cch + 3] - hi.w;
                    *reinterpret_cast<float4 *>(Ahi + ht * 128 + ((cch ^ (ht & 7)) << 4)) = hi;
                    *reinterpret_cast<float4 *>(Ahi + ht * 128 + (((cch + 4) ^ (ht & 7)) << 4)) = lo;
                }
            }
            fence_proxy_async();
            fence_before_sync();
            bar_sync(bar_id, 128);
            if (ht == 0) {
                fence_after_sync();
                umma_tf32(tmem_d, dAhi + 0, dX1 + 0, IDESC_TF32_M128_N32, 0);     // hi x Whi
                umma_tf32(tmem_d, dAhi + 2, dX1 + 2, IDESC_TF32_M128_N32, 1);
                umma_tf32(tmem_d, dAhi + 4, dX1 + 0, IDESC_TF32_M128_N32, 1);     // lo x Whi
                umma_tf32(tmem_d, dAhi + 6, dX1 + 2, IDESC_TF32_M128_N32, 1);
                umma_tf32(tmem_d, dAhi + 0, dX1 + 4, IDESC_TF32_M128_N32, 1);     // hi x Wlo
                umma_tf32(tmem_d, dAhi + 2, dX1 + 6, IDESC_TF32_M128_N32, 1);
                umma_commit(mbar);
            }
            float v[32];
            {   // P[row] + Q[col] while the tensor core works  (first edge Linear, node halves; bias in Q)
                const float *Pr = a.P + (int64_t)r * H, *Qc = a.Q + (int64_t)c * H;
                float pq[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float4 pv = ldg4(Pr + 4 * i), qv = ldg4(Qc + 4 * i);
                    pq[4 * i] = pv.x + qv.x; pq[4 * i + 1] = pv.y + qv.y; pq[4 * i + 2] = pv.z + qv.z; pq[4 * i + 3] = pv.w + qv.w;
                }
                mbar_wait(mbar, phase); phase ^= 1;
                fence_after_sync();
                tmem_ld32(tmem_row, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = silu(v[i] + pq[i]);                      // :203-206 act
            }
            // ---------------- stage 2: per-head second Linear (block-diagonal) ----------------
            store_hilo(Ahi, Alo, ht, v);
            fence_proxy_async();
            fence_before_sync();
            bar_sync(bar_id, 128);
            if (ht == 0) {
                fence_after_sync();
                issue_3xtf32_k32(tmem_d, dAhi, dAlo, dW2hi, dW2lo);
                umma_commit(mbar);
            }
            mbar_wait(mbar, phase); phase ^= 1;
            fence_after_sync();
            tmem_ld32(tmem_row, v);
            {   // + b2, LayerNorm(32), eps 1e-5, biased variance (:209,:249)
                float mean = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) { v[j] += sb2[j]; mean += v[j]; }
                mean *= (1.0f / 32.0f);
                float var = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) { const float t = v[j] - mean; var = fmaf(t, t, var); }
                const float rstd = rsqrtf(var * (1.0f / 32.0f) + 1e-5f);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = fmaf((v[j] - mean) * rstd, slng[j], slnb[j]);
            }
            // ---------------- stage 3: coord_mlp.0 ----------------
            store_hilo(Ahi, Alo, ht, v);
            fence_proxy_async();
            fence_before_sync();
            bar_sync(bar_id, 128);
            if (ht == 0) {
                fence_after_sync();
                issue_3xtf32_k32(tmem_d, dAhi, dAlo, dW3hi, dW3lo);
                umma_commit(mbar);
            }
            // ---- feature segment sums while the tensor core works (both only READ the A tiles) ----
            while (sptr[ncur + 1] <= p0) ++ncur;
            const int tend = min(p0 + 128, pend);
            for (int nl = ncur + hw; nl < nb && sptr[nl] < tend; nl += 4) {
                const int lo = max(sptr[nl], p0) - p0, hi = min(sptr[nl + 1], tend) - p0;
                float s0 = sacc[nl * M_ROW + lane];     // running sum: strictly sequential edge order (twin stability)
                int q = lo;
                for (; q + 4 <= hi; q += 4) {
                    float m4[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int off = sw128_off(q + u, lane);
                        m4[u] = *reinterpret_cast<const float *>(Ahi + off) + *reinterpret_cast<const float *>(Alo + off);
                    }
                    s0 += m4[0]; s0 += m4[1]; s0 += m4[2]; s0 += m4[3];
                }
                for (; q < hi; ++q) {
                    const int off = sw128_off(q, lane);
                    s0 += *reinterpret_cast<const float *>(Ahi + off) + *reinterpret_cast<const float *>(Alo + off);
                }
                sacc[nl * M_ROW + lane] = s0;
            }
            // ---- accumulator -> registers, SiLU + wc2 epilogue (:219-229, :264) ----
            mbar_wait(mbar, phase); phase ^= 1;
            fence_after_sync();
            tmem_ld32(tmem_row, v);
            float s = 0.f;
#pragma unroll
            for (int o = 0; o < 32; ++o) s = fmaf(swc2[o], silu(v[o] + sbc1[o]), s);
            dxs[ht] = make_float4(dx * s, dy * s, dz * s, 0.f);                               // trans = coord_diff * s
            fence_before_sync();
            bar_sync(bar_id, 128);      // dxs complete; every thread is done with the A tiles and the accumulator
            {   // coordinate segment sums: thread (slot = ht/4, component = ht%4)
                const int comp = ht & 3;
                for (int nl = ncur + (ht >> 2); nl < nb && sptr[nl] < tend; nl += 32) {
                    if (comp < 3) {
                        const int lo = max(sptr[nl], p0) - p0, hi = min(sptr[nl + 1], tend) - p0;
                        float s1 = sacc[nl * M_ROW + 32 + comp];
                        for (int q = lo; q < hi; ++q) s1 += reinterpret_cast<const float *>(dxs + q)[comp];
                        sacc[nl * M_ROW + 32 + comp] = s1;
                    }
                }
            }
        }
        bar_sync(bar_id, 128);
        // ---- write the block's aggregates and updated coordinates ----
        for (int nl = hw; nl < nb; nl += 4) {
            const int64_t g = n0 + nl;
            agg_out[g * H + lane] = sacc[nl * M_ROW + lane];
            float xv = 0.f;
            if (lane < 3) xv = __ldg(a.x4 + g * 4 + lane) + sacc[nl * M_ROW + 32 + lane];      // coord + agg  :267
            if (lane < 4) a.x4_out[g * 4 + lane] = xv;
            if (a.x3_out && lane < 3) a.x3_out[g * 3 + lane] = xv;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tmem_holder, 32 * M_HALVES);
}

void launch_node_kernel(const LayerArgs &a, const float *agg, cudaStream_t st);   // egnn_layer_tc.cu

int launch_layer_mma(const LayerArgs &a, float *agg_ws, cudaStream_t st) {
    static bool configured = false;
    static int ctas = 2;
    if (!configured) {
        if (cudaFuncSetAttribute(egcl_edge_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M_SMEM_BYTES) != cudaSuccess)
            return EGSPR_E_LAUNCH;
        if (const char *ov = getenv("EGSPR_MMA_CTAS")) ctas = atoi(ov) > 0 ? atoi(ov) : 2;
        configured = true;
    }
    const int64_t items = (a.num_nodes + M_NB - 1) / M_NB;
    int64_t grid = (int64_t)sm_count() * ctas;
    const int64_t need = (items + M_HALVES - 1) / M_HALVES;
    if (grid > need) grid = need;
    egcl_edge_mma_kernel<<<(unsigned)grid, M_THREADS, M_SMEM_BYTES, st>>>(a, agg_ws);
    if (cudaGetLastError() != cudaSuccess) return EGSPR_E_LAUNCH;
    launch_node_kernel(a, agg_ws, st);
    if (cudaGetLastError() != cudaSuccess) return EGSPR_E_LAUNCH;
    return EGSPR_OK;
}

}  // namespace egspr
