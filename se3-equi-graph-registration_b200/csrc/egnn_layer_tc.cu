// Tensor-core (tcgen05 / TMEM) path of the E_GCL layer for sm_100a.
//
// The per-edge dense contraction that dominates the layer -- coord_mlp.0, a [E,32] x [32,32] GEMM
// (src/3dmatch_train_egnn_with_batch.py:219-229, 1024 of the ~1700 MACs per edge) -- runs on the
// 5th-gen tensor cores as 128-edge x 32 x 32 tiles: every thread writes its edge's LayerNorm'd
// message row into a 128B-swizzled K-major shared-memory tile, one thread issues tcgen05.mma
// (kind::tf32, M=128, N=32, K=8 x4), the accumulator lives in TMEM and is read back with
// tcgen05.ld (32x32b: thread t gets row t) for the SiLU / wc2 epilogue.
//
// fp32 parity: the tensor cores read TF32 (10-bit mantissa), so the product is formed as the 3xTF32
// split  A W^T ~= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T  (hi = fp32 with the low 13 mantissa bits
// cleared, lo = x - hi, exact), accumulated in fp32 in TMEM: relative error ~2^-20 per term instead
// of 2^-11.  The same two tiles (hi + lo == m exactly) feed the per-node segment sums, so the
// messages never exist anywhere else.
//
// Per layer: egcl_edge_tc_kernel (gather, geometry, first edge Linear + heads + LayerNorm on CUDA
// cores; coord MLP on tcgen05; sequential-order segment sums in shared memory; writes agg, x')
// then egcl_node_kernel (node MLP + residual, next layer's P/Q or embedding_out).
#include <cstdio>
#include <cstdlib>

#include "egnn_layer.cuh"
#include "tcgen05.cuh"

namespace egspr {

constexpr int T_THREADS = 128;
constexpr int T_TILE = 128;          // edges per tile == UMMA M
constexpr int T_NB = 128;            // nodes per work item
constexpr int T_SW = 736;            // floats of the layer pack the CUDA-core stages need (WG, W2P, B2, LN)
constexpr uint32_t T_TMEM_COLS = 32;

// shared-memory carve-up (bytes from a 1024-aligned base)
constexpr int TS_AHI = 0, TS_ALO = 16384, TS_WHI = 32768, TS_WLO = 36864, TS_SW = 40960;
constexpr int TS_SWEA = TS_SW + T_SW * 4, TS_BC1 = TS_SWEA + 128, TS_WC2 = TS_BC1 + 128, TS_DXS = TS_WC2 + 128;
constexpr int TS_SACC = TS_DXS + T_TILE * 16, TS_SPTR = TS_SACC + T_NB * L_ROW * 4;
constexpr int TS_MBAR = ((TS_SPTR + (T_NB + 4) * 4 + 7) / 8) * 8, TS_TMEM = TS_MBAR + 8, TS_END = TS_TMEM + 8;
constexpr size_t T_SMEM_BYTES = TS_END + 1024;   // + slack for the manual 1024-byte alignment

using namespace tc;
constexpr uint32_t T_IDESC = IDESC_TF32_M128_N32;
__device__ __forceinline__ void tc_fence_before() { fence_before_sync(); }
__device__ __forceinline__ void tc_fence_after() { fence_after_sync(); }

__global__ void __launch_bounds__(T_THREADS, 3) egcl_edge_tc_kernel(const LayerArgs a, float *__restrict__ agg_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *Ahi = base + TS_AHI, *Alo = base + TS_ALO;
    float *sw = reinterpret_cast<float *>(base + TS_SW);
    float *swea = reinterpret_cast<float *>(base + TS_SWEA);
    float *sbc1 = reinterpret_cast<float *>(base + TS_BC1);
    float *swc2 = reinterpret_cast<float *>(base + TS_WC2);
    float4 *dxs = reinterpret_cast<float4 *>(base + TS_DXS);
    float *sacc = reinterpret_cast<float *>(base + TS_SACC);
    int *sptr = reinterpret_cast<int *>(base + TS_SPTR);
    const uint32_t mbar = smem_u32(base + TS_MBAR);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + TS_TMEM);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- one-time setup: weights, swizzled hi/lo copies of coord_mlp.0.weight, TMEM, mbarrier ----
    for (int i = tid; i < T_SW; i += T_THREADS) sw[i] = __ldg(a.layer_pack + i);
    if (tid < 32) {
        swea[tid] = __ldg(a.layer_pack + OFF_WEA + tid);
        sbc1[tid] = __ldg(a.layer_pack + OFF_BC1 + tid);
        swc2[tid] = __ldg(a.layer_pack + OFF_WC2 + tid);
    }
    for (int i = tid; i < 32 * 32; i += T_THREADS) {          // B operand: row = output o, K = input i (K-major)
        const int o = i >> 5, k = i & 31;
        const float w = __ldg(a.layer_pack + OFF_WC1 + i);
        const float hi = tf32_hi(w);
        *reinterpret_cast<float *>(base + TS_WHI + sw128_off(o, k)) = hi;
        *reinterpret_cast<float *>(base + TS_WLO + sw128_off(o, k)) = w - hi;
    }
    if (warp == 0) tmem_alloc(smem_u32(tmem_holder), T_TMEM_COLS);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();            // the weight tiles were written through the generic proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_d = *tmem_holder;
    const uint64_t dAhi = make_desc_sw128(smem_u32(Ahi)), dAlo = make_desc_sw128(smem_u32(Alo));
    const uint64_t dWhi = make_desc_sw128(smem_u32(base + TS_WHI)), dWlo = make_desc_sw128(smem_u32(base + TS_WLO));
    uint32_t phase = 0;

    const int64_t G = a.num_nodes;
    const int64_t items = (G + T_NB - 1) / T_NB;
    for (int64_t item = blockIdx.x; item < items; item += gridDim.x) {
        const int64_t n0 = item * T_NB;
        const int nb = (G - n0 < T_NB) ? (int)(G - n0) : T_NB;
        __syncthreads();
        for (int i = tid; i <= nb; i += T_THREADS) sptr[i] = __ldg(a.csr_ptr + n0 + i);
        for (int i = tid; i < T_NB * L_ROW; i += T_THREADS) sacc[i] = 0.f;
        __syncthreads();
        const int pbeg = sptr[0], pend = sptr[nb];
        int ncur = 0;
        for (int p0 = pbeg; p0 < pend; p0 += T_TILE) {
            float dx, dy, dz;
            {
                float um[32];
                int p = p0 + tid;
                if (p >= pend) p = pend - 1;           // idle slot: recompute the last edge, never reduced
                edge_message(a, sw, swea, p, um, dx, dy, dz);
                // message row -> hi / lo tiles (128B-swizzled: 16-byte chunk c of row t sits at chunk c ^ (t & 7))
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    float4 hi, lo;
                    hi.x = tf32_hi(um[4 * c]); hi.y = tf32_hi(um[4 * c + 1]); hi.z = tf32_hi(um[4 * c + 2]); hi.w = tf32_hi(um[4 * c + 3]);
                    lo.x = um[4 * c] - hi.x; lo.y = um[4 * c + 1] - hi.y; lo.z = um[4 * c + 2] - hi.z; lo.w = um[4 * c + 3] - hi.w;
                    const int off = tid * 128 + ((c ^ (tid & 7)) << 4);
                    *reinterpret_cast<float4 *>(Ahi + off) = hi;
                    *reinterpret_cast<float4 *>(Alo + off) = lo;
                }
            }
            fence_proxy_async();        // generic-proxy smem writes -> visible to the tensor core (async proxy)
            tc_fence_before();          // this thread's earlier tcgen05.ld is ordered before the next MMA
            __syncthreads();
            if (tid == 0) {
                tc_fence_after();
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, dAhi + 2 * k, dWhi + 2 * k, T_IDESC, k > 0);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, dAlo + 2 * k, dWhi + 2 * k, T_IDESC, 1);
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_tf32(tmem_d, dAhi + 2 * k, dWlo + 2 * k, T_IDESC, 1);
                umma_commit(mbar);      // arrives on the mbarrier when all 12 MMAs have completed
            }
            // ---- feature segment sums while the tensor core works (both only READ the A tiles) ----
            while (sptr[ncur + 1] <= p0) ++ncur;
            const int tend = min(p0 + T_TILE, pend);
            for (int nl = ncur + warp; nl < nb && sptr[nl] < tend; nl += T_THREADS / 32) {
                const int lo = max(sptr[nl], p0), hi = min(sptr[nl + 1], tend);
                float s0 = sacc[nl * L_ROW + lane];     // running sum: strictly sequential edge order (twin stability)
                for (int p = lo; p < hi; ++p) {
                    const int off = sw128_off(p - p0, lane);
                    s0 += *reinterpret_cast<const float *>(Ahi + off) + *reinterpret_cast<const float *>(Alo + off);
                }
                sacc[nl * L_ROW + lane] = s0;
            }
            // ---- accumulator -> registers, SiLU + wc2 epilogue (:219-229, :264) ----
            mbar_wait(mbar, phase);
            phase ^= 1;
            tc_fence_after();
            float s = 0.f;
            {
                float t[32];
                tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16), t);
#pragma unroll
                for (int o4 = 0; o4 < 8; ++o4) {
                    const float4 b = *reinterpret_cast<const float4 *>(sbc1 + 4 * o4);
                    const float4 w = *reinterpret_cast<const float4 *>(swc2 + 4 * o4);
                    s = fmaf(w.x, silu(t[4 * o4] + b.x), s); s = fmaf(w.y, silu(t[4 * o4 + 1] + b.y), s);
                    s = fmaf(w.z, silu(t[4 * o4 + 2] + b.z), s); s = fmaf(w.w, silu(t[4 * o4 + 3] + b.w), s);
                }
            }
            dxs[tid] = make_float4(dx * s, dy * s, dz * s, 0.f);
            tc_fence_before();
            __syncthreads();            // dxs complete; every thread is done with the A tiles and the accumulator
            for (int nl = ncur + warp; nl < nb && sptr[nl] < tend; nl += T_THREADS / 32) {
                if (lane < 3) {
                    const int lo = max(sptr[nl], p0), hi = min(sptr[nl + 1], tend);
                    float s1 = sacc[nl * L_ROW + 32 + lane];
                    for (int p = lo; p < hi; ++p) s1 += reinterpret_cast<const float *>(dxs + (p - p0))[lane];
                    sacc[nl * L_ROW + 32 + lane] = s1;
                }
            }
        }
        __syncthreads();
        // ---- write the block's aggregates and updated coordinates ----
        for (int nl = warp; nl < nb; nl += T_THREADS / 32) {
            const int64_t g = n0 + nl;
            agg_out[g * H + lane] = sacc[nl * L_ROW + lane];
            float xv = 0.f;
            if (lane < 3) xv = __ldg(a.x4 + g * 4 + lane) + sacc[nl * L_ROW + 32 + lane];      // coord + agg  :267
            if (lane < 4) a.x4_out[g * 4 + lane] = xv;
            if (a.x3_out && lane < 3) a.x3_out[g * 3 + lane] = xv;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, T_TMEM_COLS);
}

// ---- node MLP + residual + next P/Q (or embedding_out), one thread per node ---------------------
constexpr int N_THREADS = 128;
constexpr int N_ROW = 33;
constexpr size_t N_SMEM_BYTES = sizeof(float) * (L_WFLOATS + 2 * N_THREADS * N_ROW);

__global__ void __launch_bounds__(N_THREADS) egcl_node_kernel(const LayerArgs a, const float *__restrict__ agg) {
    extern __shared__ __align__(16) float nsm[];
    float *sw = nsm;
    float *rows = nsm + L_WFLOATS;
    const int tid = threadIdx.x;
    for (int i = tid; i < NODE_PART / 4; i += N_THREADS)
        reinterpret_cast<float4 *>(sw + OFF_WN1T)[i] = ldg4(a.layer_pack + OFF_WN1T + 4 * i);
    if (a.next_pack) {
        for (int i = tid; i < PQ_PART / 4; i += N_THREADS)
            reinterpret_cast<float4 *>(sw + OFF_WPT)[i] = ldg4(a.next_pack + OFF_WPT + 4 * i);
    } else if (a.out_pack) {
        for (int i = tid; i < EMBED_PACK / 4; i += N_THREADS)
            reinterpret_cast<float4 *>(sw + OFF_WPT)[i] = ldg4(a.out_pack + 4 * i);
    }
    __syncthreads();
    float *hrow = rows + tid * N_ROW, *arow = rows + (N_THREADS + tid) * N_ROW;
    for (int64_t g = blockIdx.x * (int64_t)N_THREADS + tid; g < a.num_nodes; g += (int64_t)gridDim.x * N_THREADS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = ldg4(agg + g * H + 4 * i);
            arow[4 * i] = t.x; arow[4 * i + 1] = t.y; arow[4 * i + 2] = t.z; arow[4 * i + 3] = t.w;
        }
        node_update(a, g, hrow, arow, sw);
    }
}

static int g_node_ctas = 0;
void launch_node_kernel(const LayerArgs &a, const float *agg, cudaStream_t st) {
    if (g_node_ctas == 0) {
        int n = 1;
        cudaFuncSetAttribute(egcl_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)N_SMEM_BYTES);
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, egcl_node_kernel, N_THREADS, N_SMEM_BYTES) != cudaSuccess || n < 1) n = 1;
        g_node_ctas = n;
    }
    int64_t ngrid = (a.num_nodes + N_THREADS - 1) / N_THREADS;
    if (ngrid > (int64_t)sm_count() * g_node_ctas) ngrid = (int64_t)sm_count() * g_node_ctas;
    egcl_node_kernel<<<(unsigned)ngrid, N_THREADS, N_SMEM_BYTES, st>>>(a, agg);
}

int launch_layer_tc(const LayerArgs &a, float *agg_ws, cudaStream_t st) {
    static bool configured = false;
    static int edge_ctas = 1, node_ctas = 1;
    if (!configured) {
        if (cudaFuncSetAttribute(egcl_edge_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)T_SMEM_BYTES) != cudaSuccess ||
            cudaFuncSetAttribute(egcl_node_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)N_SMEM_BYTES) != cudaSuccess)
            return EGSPR_E_LAUNCH;
        cudaError_t e1 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&edge_ctas, egcl_edge_tc_kernel, T_THREADS, T_SMEM_BYTES);
        cudaError_t e2 = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&node_ctas, egcl_node_kernel, N_THREADS, N_SMEM_BYTES);
        if (getenv("EGSPR_DEBUG"))
            fprintf(stderr, "[egspr] tc occupancy: edge %d (%s) node %d (%s) smem %zu %zu\n", edge_ctas, cudaGetErrorString(e1),
                    node_ctas, cudaGetErrorString(e2), T_SMEM_BYTES, N_SMEM_BYTES);
        // The occupancy API reports 1 block/SM for the TMEM-allocating kernel although 3 are resident
        // (ncu: sm__maximum_warps_avg_per_active_cycle = 12); the launch bound guarantees the registers
        // and 3 x 67 KB of shared memory fit, so size the persistent grid from the launch bound.
        edge_ctas = 3;
        if (const char *ov = getenv("EGSPR_TC_CTAS")) edge_ctas = atoi(ov) > 0 ? atoi(ov) : 3;
        if (e2 != cudaSuccess || node_ctas < 1) node_ctas = 1;
        cudaGetLastError();
        configured = true;
    }
    const int64_t items = (a.num_nodes + T_NB - 1) / T_NB;
    int64_t grid = (int64_t)sm_count() * edge_ctas;
    if (grid > items) grid = items;
    egcl_edge_tc_kernel<<<(unsigned)grid, T_THREADS, T_SMEM_BYTES, st>>>(a, agg_ws);
    int64_t ngrid = (a.num_nodes + N_THREADS - 1) / N_THREADS;
    if (ngrid > (int64_t)sm_count() * node_ctas) ngrid = (int64_t)sm_count() * node_ctas;
    egcl_node_kernel<<<(unsigned)ngrid, N_THREADS, N_SMEM_BYTES, st>>>(a, agg_ws);
    if (cudaGetLastError() != cudaSuccess) return EGSPR_E_LAUNCH;
    return EGSPR_OK;
}

}  // namespace egspr
