// E_GCL edge kernel, impl 5: all three per-edge contractions on tcgen05 with the A operand handed
// to the tensor core THROUGH TENSOR MEMORY (tcgen05.st -> tcgen05.mma [d], [a_tmem], b_desc), and a
// streaming segment reduction that writes finished aggregation rows straight to global memory.
//
// Reference arithmetic: src/3dmatch_train_egnn_with_batch.py:231-250 (edge_model), :262-268
// (coord_model), :252-254 (segment sums of node_model); SURVEY Appendix A.2.
//
// A group = 128 threads = 128 TMEM lanes = one tile of 128 edges (one edge per thread, edges in CSR
// order = grouped by aggregation row).  A CTA has 4 independent groups (named barriers) sharing the
// weight tiles; each group streams over its own contiguous range of aggregation rows.
//   stage 1  pre = P[row] + Q[col] + [geo(12) | edge_attr] Wg^T       M128 N32 K16
//   stage 2  u   = SiLU(pre) W2^T + b2 (W2 = block-diag of the heads' 8x8),  m = LayerNorm_32(u)
//   stage 3  s   = wc2 . SiLU(m Wc1^T + bc1)
// Each thread writes its row of the A operand into TMEM columns (32x32b: thread t <-> lane t), one
// elected thread issues the MMAs, the accumulator comes back with tcgen05.ld.  No shared-memory
// A tiles, no async-proxy fences.  TMEM columns per group: D 0..31 | A_hi 32..63 | A_lo 64..95.
//
// fp32 parity: 3xTF32 split (x = hi + lo, hi = x with the low 13 mantissa bits cleared, lo exact):
//   A W^T ~= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T, fp32 accumulation in TMEM.
//
// Segment sums: the messages of the tile are also written (full fp32) to a shared-memory tile; a
// warp per aggregation row continues the row's running sum in ascending edge order (bit-reproducible;
// duplicate points stay bit-identical), writes the row to global memory when its last edge is in the
// tile, or parks it in a double-buffered carry when the row continues in the next tile.
#include <cstdio>
#include <cstdlib>

#include "egnn_layer.cuh"
#include "tcgen05.cuh"

namespace egspr {
using namespace tc;

constexpr int V_GROUPS = 4;
constexpr int V_THREADS = 128 * V_GROUPS;
constexpr int V_MROW = 36;            // floats per row of the message tile (144 B: conflict-free STS.128)

// shared-memory carve-up (bytes from a 1024-aligned base)
constexpr int VS_W = 0;                               // X1 | W2hi | W2lo | W3hi | W3lo, 4 KB each (SW128 K-major)
constexpr int VS_PAR = VS_W + 5 * 4096;               // b2, ln gamma, ln beta, bc1, wc2 (32 floats each)
constexpr int VS_GRP = VS_PAR + 5 * 128;
constexpr int VG_MT = 0;                              // float[128][36]  messages of the tile
constexpr int VG_DXS = VG_MT + 128 * V_MROW * 4;      // float4[128]     coord_diff * s of the tile
constexpr int VG_CARRY = VG_DXS + 128 * 16;           // float[2][36]    running sums of a row that spans tiles
constexpr int VG_RLAST = VG_CARRY + 2 * 36 * 4;       // int             aggregation row of the tile's last edge
constexpr int VG_MBAR = VG_RLAST + 8;
constexpr int VG_SIZE = ((VG_MBAR + 8 + 127) / 128) * 128;
constexpr int VS_TMEM = VS_GRP + V_GROUPS * VG_SIZE;
constexpr int VS_END = VS_TMEM + 16;
constexpr size_t V_SMEM_BYTES = VS_END + 1024;        // + slack for the manual 1024-byte alignment

__device__ __forceinline__ float silu_fast(float v) {
    // v * sigmoid(v) = v * rcp(1 + 2^(-v log2 e)); MUFU.EX2 + MUFU.RCP, no range fix-ups
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(v * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return v * r;
}

// this thread's 32 values -> hi / lo halves of the A operand in TMEM (columns [0,32) of each)
__device__ __forceinline__ void store_hilo_tmem(uint32_t t_hi, uint32_t t_lo, const float (&v)[32]) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { hi[i] = tf32_hi(v[16 * b + i]); lo[i] = v[16 * b + i] - hi[i]; }
        tmem_st16(t_hi + 16 * b, hi);
        tmem_st16(t_lo + 16 * b, lo);
    }
}

// D = A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T over K = 32 (4 K-blocks of 8 columns)
__device__ __forceinline__ void issue_3xtf32_ts(uint32_t d, uint32_t ahi, uint32_t alo, uint64_t whi, uint64_t wlo) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, ahi + 8 * k, whi + 2 * k, IDESC_TF32_M128_N32, k > 0);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, alo + 8 * k, whi + 2 * k, IDESC_TF32_M128_N32, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, ahi + 8 * k, wlo + 2 * k, IDESC_TF32_M128_N32, 1);
}

__global__ void __launch_bounds__(V_THREADS, 1) egcl_edge_ts_kernel(const LayerArgs a, float *__restrict__ agg_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, grp = tid >> 7, ht = tid & 127, lane = tid & 31, hw = ht >> 5;
    const float *sb2 = reinterpret_cast<const float *>(base + VS_PAR);
    const float *slng = sb2 + 32, *slnb = sb2 + 64, *sbc1 = sb2 + 96, *swc2 = sb2 + 128;
    uint8_t *gb = base + VS_GRP + grp * VG_SIZE;
    float *mt = reinterpret_cast<float *>(gb + VG_MT);
    float4 *dxs = reinterpret_cast<float4 *>(gb + VG_DXS);
    float *carry = reinterpret_cast<float *>(gb + VG_CARRY);
    int *s_rlast = reinterpret_cast<int *>(gb + VG_RLAST);
    const uint32_t mbar = smem_u32(gb + VG_MBAR);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + VS_TMEM);
    const int bar_id = 1 + grp;

    // ---- one-time setup: swizzled hi/lo weight tiles (B operands: row = output o, K = input) ----
    for (int i = tid; i < 1024; i += V_THREADS) {
        const int o = i >> 5, k = i & 31;
        {   // stage 1: K 0..15 = hi of [Wg(12) | w_edge_attr | 0 0 0], K 16..31 = lo of the same
            const int kk = k & 15;
            float w = 0.f;
            if (kk < 12) w = __ldg(a.layer_pack + OFF_WG + 32 * kk + o);
            else if (kk == 12) w = __ldg(a.layer_pack + OFF_WEA + o);
            const float hi = tf32_hi(w);
            *reinterpret_cast<float *>(base + VS_W + sw128_off(o, k)) = (k < 16) ? hi : (w - hi);
        }
        {   // stage 2: block-diagonal of the heads' second Linear, pack layout [head][in][out]
            const float w = ((o >> 3) == (k >> 3)) ? __ldg(a.layer_pack + OFF_W2P + 64 * (o >> 3) + 8 * (k & 7) + (o & 7)) : 0.f;
            const float hi = tf32_hi(w);
            *reinterpret_cast<float *>(base + VS_W + 4096 + sw128_off(o, k)) = hi;
            *reinterpret_cast<float *>(base + VS_W + 8192 + sw128_off(o, k)) = w - hi;
        }
        {   // stage 3: coord_mlp.0.weight [out][in]
            const float w = __ldg(a.layer_pack + OFF_WC1 + i);
            const float hi = tf32_hi(w);
            *reinterpret_cast<float *>(base + VS_W + 12288 + sw128_off(o, k)) = hi;
            *reinterpret_cast<float *>(base + VS_W + 16384 + sw128_off(o, k)) = w - hi;
        }
    }
    if (tid < 32) {
        float *par = reinterpret_cast<float *>(base + VS_PAR);
        par[tid] = __ldg(a.layer_pack + OFF_B2 + tid);
        par[32 + tid] = __ldg(a.layer_pack + OFF_LNG + tid);
        par[64 + tid] = __ldg(a.layer_pack + OFF_LNB + tid);
        par[96 + tid] = __ldg(a.layer_pack + OFF_BC1 + tid);
        par[128 + tid] = __ldg(a.layer_pack + OFF_WC2 + tid);
    }
    if (tid < 32) tmem_alloc(smem_u32(tmem_holder), 512);
    if (ht == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();            // the weight tiles were written through the generic proxy
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_g = *tmem_holder + 128u * grp;                  // this group's 128 columns
    const uint32_t tmem_w = tmem_g + ((uint32_t)(hw * 32) << 16);       // ... at this warp's 32 lanes
    const uint32_t tD = tmem_g, tAhi = tmem_g + 32, tAlo = tmem_g + 64;
    const uint64_t dX1 = make_desc_sw128(smem_u32(base + VS_W));
    const uint64_t dW2hi = make_desc_sw128(smem_u32(base + VS_W + 4096)), dW2lo = make_desc_sw128(smem_u32(base + VS_W + 8192));
    const uint64_t dW3hi = make_desc_sw128(smem_u32(base + VS_W + 12288)), dW3lo = make_desc_sw128(smem_u32(base + VS_W + 16384));
    uint32_t phase = 0;

    // ---- this group's contiguous range of aggregation rows ----
    const int64_t G = a.num_nodes;
    const int64_t NG = (int64_t)gridDim.x * V_GROUPS, gi = (int64_t)blockIdx.x * V_GROUPS + grp;
    const int nA = (int)(G * gi / NG), nB = (int)(G * (gi + 1) / NG);
    const int pbeg = __ldg(a.csr_ptr + nA), pend = __ldg(a.csr_ptr + nB);
    int nstart = nA;                 // first row whose sums have not been written out yet
    int par = 0;                     // carry buffer read by this tile (the other one is written)

    for (int p0 = pbeg; p0 < pend; p0 += 128, par ^= 1) {
        const int tend = min(p0 + 128, pend);
        int p = p0 + ht;
        if (p >= pend) p = pend - 1;               // idle slot: recompute the last edge, never reduced
        const int r = __ldg(a.csr_row + p), c = __ldg(a.csr_col + p);
        if (ht == tend - 1 - p0) *s_rlast = r;
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
            float hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { hi[i] = tf32_hi(geo[i]); lo[i] = geo[i] - hi[i]; }
            tmem_st16(tmem_w + 32, hi);          // A_hi columns 0..15  = hi(geo)
            tmem_st16(tmem_w + 48, lo);          // A_hi columns 16..31 = lo(geo)
        }
        tmem_wait_st();
        fence_before_sync();
        bar_sync(bar_id, 128);
        if (ht == 0) {
            fence_after_sync();
            umma_tf32_ts(tD, tAhi + 0, dX1 + 0, IDESC_TF32_M128_N32, 0);      // hi x Whi
            umma_tf32_ts(tD, tAhi + 8, dX1 + 2, IDESC_TF32_M128_N32, 1);
            umma_tf32_ts(tD, tAhi + 16, dX1 + 0, IDESC_TF32_M128_N32, 1);     // lo x Whi
            umma_tf32_ts(tD, tAhi + 24, dX1 + 2, IDESC_TF32_M128_N32, 1);
            umma_tf32_ts(tD, tAhi + 0, dX1 + 4, IDESC_TF32_M128_N32, 1);      // hi x Wlo
            umma_tf32_ts(tD, tAhi + 8, dX1 + 6, IDESC_TF32_M128_N32, 1);
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
            tmem_ld32(tmem_w, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = silu_fast(v[i] + pq[i]);                 // :203-206 act
        }
        // ---------------- stage 2: per-head second Linear (block-diagonal) ----------------
        store_hilo_tmem(tmem_w + 32, tmem_w + 64, v);
        tmem_wait_st();
        fence_before_sync();
        bar_sync(bar_id, 128);
        if (ht == 0) {
            fence_after_sync();
            issue_3xtf32_ts(tD, tAhi, tAlo, dW2hi, dW2lo);
            umma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        tmem_ld32(tmem_w, v);
        {   // + b2, LayerNorm(32), eps 1e-5, biased variance (:209,:249)
            float mean = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] += sb2[j]; mean += v[j]; }
            mean *= (1.0f / 32.0f);
            float var = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) { v[j] -= mean; var = fmaf(v[j], v[j], var); }
            const float rstd = rsqrtf(var * (1.0f / 32.0f) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaf(v[j] * rstd, slng[j], slnb[j]);
        }
        // ---------------- stage 3: coord_mlp.0; messages also to shared memory ----------------
        store_hilo_tmem(tmem_w + 32, tmem_w + 64, v);
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4 *>(mt + ht * V_MROW + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        tmem_wait_st();
        fence_before_sync();
        bar_sync(bar_id, 128);
        if (ht == 0) {
            fence_after_sync();
            issue_3xtf32_ts(tD, tAhi, tAlo, dW3hi, dW3lo);
            umma_commit(mbar);
        }
        // ---- feature segment sums while the tensor core works ----
        const int rlast = *s_rlast;
        for (int n = nstart + hw; n <= rlast; n += 4) {
            const int b0 = __ldg(a.csr_ptr + n), b1 = __ldg(a.csr_ptr + n + 1);
            const int lo = max(b0, p0) - p0, hi = min(b1, tend) - p0;
            float s0 = (b0 < p0) ? carry[par * 36 + lane] : 0.f;     // row continues from the previous tile
            int q = lo;
            for (; q + 4 <= hi; q += 4) {                             // strictly sequential edge order (twin stability)
                const float m0 = mt[q * V_MROW + lane], m1 = mt[(q + 1) * V_MROW + lane];
                const float m2 = mt[(q + 2) * V_MROW + lane], m3 = mt[(q + 3) * V_MROW + lane];
                s0 += m0; s0 += m1; s0 += m2; s0 += m3;
            }
            for (; q < hi; ++q) s0 += mt[q * V_MROW + lane];
            if (b1 <= tend) agg_out[(int64_t)n * H + lane] = s0;      // row complete: out it goes
            else carry[(par ^ 1) * 36 + lane] = s0;
        }
        // ---- accumulator -> registers, SiLU + wc2 epilogue (:219-229, :264) ----
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        tmem_ld32(tmem_w, v);
        float s = 0.f;
#pragma unroll
        for (int o = 0; o < 32; ++o) s = fmaf(swc2[o], silu_fast(v[o] + sbc1[o]), s);
        dxs[ht] = make_float4(dx * s, dy * s, dz * s, 0.f);                               // trans = coord_diff * s
        fence_before_sync();
        bar_sync(bar_id, 128);      // dxs complete; every thread is done with the message tile and the accumulator
        {   // coordinate segment sums: thread (slot = ht/4, component = ht%4)
            const int comp = ht & 3;
            for (int n = nstart + (ht >> 2); n <= rlast; n += 32) {
                const int b0 = __ldg(a.csr_ptr + n), b1 = __ldg(a.csr_ptr + n + 1);
                const int lo = max(b0, p0) - p0, hi = min(b1, tend) - p0;
                float s1 = (b0 < p0) ? carry[par * 36 + 32 + comp] : 0.f;
                for (int q = lo; q < hi; ++q) s1 += reinterpret_cast<const float *>(dxs + q)[comp];
                if (b1 <= tend) {
                    const float xv = (comp < 3) ? __ldg(a.x4 + (int64_t)n * 4 + comp) + s1 : 0.f;   // coord + agg  :267
                    a.x4_out[(int64_t)n * 4 + comp] = xv;
                    if (a.x3_out && comp < 3) a.x3_out[(int64_t)n * 3 + comp] = xv;
                } else {
                    carry[(par ^ 1) * 36 + 32 + comp] = s1;
                }
            }
        }
        nstart = (__ldg(a.csr_ptr + rlast + 1) <= tend) ? rlast + 1 : rlast;
    }
    // rows after the last edge of the range have no edges at all: zero aggregate, unchanged coordinates
    for (int n = nstart + hw; n < nB; n += 4) {
        agg_out[(int64_t)n * H + lane] = 0.f;
        if (lane < 4) {
            const float xv = (lane < 3) ? __ldg(a.x4 + (int64_t)n * 4 + lane) : 0.f;
            a.x4_out[(int64_t)n * 4 + lane] = xv;
            if (a.x3_out && lane < 3) a.x3_out[(int64_t)n * 3 + lane] = xv;
        }
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tmem_holder, 512);
}

void launch_node_kernel(const LayerArgs &a, const float *agg, cudaStream_t st);   // egnn_layer_tc.cu

int launch_layer_ts(const LayerArgs &a, float *agg_ws, cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(egcl_edge_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)V_SMEM_BYTES) != cudaSuccess)
            return EGSPR_E_LAUNCH;
        configured = true;
    }
    // one persistent CTA per SM; small graphs: at least ~2 tiles of edges per group
    int64_t grid = sm_count();
    const int64_t need = (a.num_nodes + 63) / 64;
    if (grid > need) grid = need;
    egcl_edge_ts_kernel<<<(unsigned)grid, V_THREADS, V_SMEM_BYTES, st>>>(a, agg_ws);
    if (cudaGetLastError() != cudaSuccess) return EGSPR_E_LAUNCH;
    launch_node_kernel(a, agg_ws, st);
    if (cudaGetLastError() != cudaSuccess) return EGSPR_E_LAUNCH;
    return EGSPR_OK;
}

}  // namespace egspr
