// E_GCL edge backward on tcgen05 (the dominant kernel of the training step): the gradient of edge_model + coord_model
// (src/3dmatch_train_egnn_with_batch.py:231-250, 262-268, through :128-181 and :271-278) that `loss.backward()` at
// 3dm:1125 produces, for all edges of a batch.
//
// Thread = edge = TMEM lane, a group of 128 threads owns a tile of 128 edges (row-CSR order), three independent groups
// per CTA (named barriers), one CTA per SM.  Nothing per-edge is kept from the forward pass: the tile's forward is
// recomputed from the layer input (P[row] + Q[col], coordinates) and the gradient pushed back, as six per-edge
// contractions [128 x 32] . [32 x 32] (three forward, three transposed) -- and the three weight-gradient products
// [32 x 128] . [128 x 32] per tile -- all on the tensor core:
//
//   every 32-vector of an edge (geo, a1 = SiLU(pre), m, dc1, du, dpre) is written ONCE, as its thread's 128-byte row of
//   a SWIZZLE_128B shared-memory tile, split into bf16 terms [b0 | b1] (x = b0 + b1 to 16 mantissa bits; geo: three
//   terms).  The same bytes are read three ways by `tcgen05.mma.kind::f16` (fp32 accumulation in tensor memory):
//     * K-major A operand of the per-edge product           D[e][o]  = sum_k  x[e][k] W[o][k]
//     * MN-major A operand of the weight-gradient product   dW[o][i] += sum_e dx[e][o] y[e][i]   (K = the 128 edges)
//     * MN-major B operand of the same product (the y side)
//   and the weight tiles (bf16 terms [b0 | b1] per row) are read K-major for W^T and MN-major for the transposed
//   products, so no transposed copies exist.  Bias gradients are the same MMAs against a constant all-ones B.
//   (operand layouts checked bit-exactly on a B200 by tools/micro/umma_probe.cu)
//
// Column sums that are not sums of a tile (LayerNorm gamma, wc2) accumulate per thread in tensor memory and are reduced
// once per CTA; d LayerNorm beta = sum_e dm_e = sum_n deg(n) dagg[n] + (sum_e dc1_e) Wc1 needs no per-edge work at all
// (first term: node_mlp_backward_kernel, second: this kernel's epilogue).
//
// TMEM (512 columns): per group D 32 | private sums 64; weight-gradient accumulators 96 columns (M = 64 layout: 16 lanes
// per warp quarter), groups 0 and 1 interleaved in the same columns at lane offsets 0 / 16, group 2 in its own.
#include "egnn_backward.cuh"
#include "egnn_backward_math.cuh"
#include "tcgen05.cuh"

namespace egspr {
using namespace tc;
using namespace bwd;

constexpr int XG = 3;                         // groups per CTA
constexpr int X_THREADS = 128 * XG;
constexpr int XS_W = 0;                       // weight tiles: Wg | W2 | Wc1, 32 rows x 128 B each
constexpr int XS_ONES = XS_W + 3 * 4096;      // 1 KB of bf16 1.0
constexpr int XS_PAR = XS_ONES + 1024;        // b2, ln gamma, ln beta, bc1, wc2 (32 floats each) | Wc1 fp32 [32][32] for the epilogue
constexpr int XS_GRP = ((XS_PAR + 5 * 128 + 4096 + 1023) / 1024) * 1024;
constexpr int XG_TILE = 16384;                // one 128 x 128-byte tile
constexpr int XG_SIZE = 4 * XG_TILE;          // geo | a1 | m -> du | dc1 -> dpre
constexpr int XS_MISC = XS_GRP + XG * XG_SIZE;
constexpr size_t X_SMEM_BYTES = XS_MISC + 16 + 32 * XG + 1024;

// TMEM columns
constexpr uint32_t XT_D = 0;                  // + 32 g
constexpr uint32_t XT_PRIV = 96;              // + 64 g: [dlng 32 | dwc2 32]
constexpr uint32_t XT_WG = 288;               // set 0 (groups 0, 1 interleaved), set 1 at + 96 (group 2)
constexpr uint32_t XW_WC1 = 0, XW_W2 = 32, XW_WGEO = 64, XW_BC1 = 80, XW_B2 = 88;

constexpr uint32_t ID_KK_N32 = idesc_bf16(128, 32, 0, 0);
constexpr uint32_t ID_KM_N32 = idesc_bf16(128, 32, 0, 1);
constexpr uint32_t ID_KM_N16 = idesc_bf16(128, 16, 0, 1);
constexpr uint32_t ID_MM_N32 = idesc_bf16(64, 32, 1, 1);
constexpr uint32_t ID_MM_N16 = idesc_bf16(64, 16, 1, 1);
constexpr uint32_t ID_MK_N8 = idesc_bf16(64, 8, 1, 0);

#ifdef EGSPR_XB_TIMING       // developer build: per-phase clock64() timeline of one group (tools/xb_timing.py)
__device__ long long g_xb_dbg[4 * 64 * 32];
#define XB_MARK(slot)                                                                                              \
    do {                                                                                                           \
        if (blockIdx.x == 7 && grp == 1 && lane == 0 && tile_no < 64) g_xb_dbg[(hw * 64 + tile_no) * 32 + (slot)] = clock64(); \
    } while (0)
#else
#define XB_MARK(slot)
#endif

__device__ __forceinline__ void cp_async16_cg(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// loads that must be ISSUED where they are written (look-ahead loads: a plain __ldg is "invariant" and gets sunk to its
// first use, which turns the prefetch into a stall)
__device__ __forceinline__ int ldg_now(const int32_t *p) {
    int v;
    asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float ldg_now(const float *p) {
    float v;
    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldg4_now(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ldcg4_now(const float *p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// low half = bf16(a), high half = bf16(b), both truncated
__device__ __forceinline__ uint32_t pack_hi16(float a, float b) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
    return r;
}
__device__ __forceinline__ float trunc16(float v) { return __uint_as_float(__float_as_uint(v) & 0xffff0000u); }

// this thread's 32 values -> its row of a tile: chunks 0..3 = b0 (truncated bf16), chunks 4..7 = b1 (bf16 of the exact residual)
__device__ __forceinline__ void write_row_split2(uint8_t *tile, int row, const float (&v)[32]) {
    uint8_t *rp = tile + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint32_t b0[4], b1[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x0 = v[8 * c + 2 * q], x1 = v[8 * c + 2 * q + 1];
            b0[q] = pack_hi16(x0, x1);
            float r0 = x0, r1 = x1;
            fadd2(r0, r1, -trunc16(x0), -trunc16(x1));
            b1[q] = pack_hi16(r0, r1);
        }
        *reinterpret_cast<uint4 *>(rp + ((c ^ sw) << 4)) = make_uint4(b0[0], b0[1], b0[2], b0[3]);
        *reinterpret_cast<uint4 *>(rp + (((c + 4) ^ sw) << 4)) = make_uint4(b1[0], b1[1], b1[2], b1[3]);
    }
}
// 16 values -> [b0 (16) | b1 (16) | b2 (16) | unused]: three bf16 terms (24 mantissa bits) for the geometric inputs, whose
// magnitudes (x_i . x_j, |d|^2) follow the coordinate scale
__device__ __forceinline__ void write_row_split3(uint8_t *tile, int row, const float (&v)[16]) {
    uint8_t *rp = tile + row * 128;
    const int sw = row & 7;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint32_t b0[4], b1[4], b2[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float x0 = v[8 * c + 2 * q], x1 = v[8 * c + 2 * q + 1];
            b0[q] = pack_hi16(x0, x1);
            const float r0 = x0 - trunc16(x0), r1 = x1 - trunc16(x1);
            b1[q] = pack_hi16(r0, r1);
            b2[q] = pack_hi16(r0 - trunc16(r0), r1 - trunc16(r1));
        }
        *reinterpret_cast<uint4 *>(rp + ((c ^ sw) << 4)) = make_uint4(b0[0], b0[1], b0[2], b0[3]);
        *reinterpret_cast<uint4 *>(rp + (((c + 2) ^ sw) << 4)) = make_uint4(b1[0], b1[1], b1[2], b1[3]);
        *reinterpret_cast<uint4 *>(rp + (((c + 4) ^ sw) << 4)) = make_uint4(b2[0], b2[1], b2[2], b2[3]);
    }
}

// sigmoid of a pair with one shared reciprocal (as silu2 in egspr_common.cuh)
__device__ __forceinline__ void sigmoid2(float x0, float x1, float &s0, float &s1) {
    float t0 = x0, t1 = x1;
    fmul2(t0, t1, -1.4426950408889634f, -1.4426950408889634f);
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(t0));
    asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(t1));
    t0 = fminf(t0, 1e18f); t1 = fminf(t1, 1e18f);
    fadd2(t0, t1, 1.0f, 1.0f);
    float r = t0 * t1;
    asm("rcp.approx.ftz.f32 %0, %0;" : "+f"(r));
    s0 = t1 * r; s1 = t0 * r;
}

// private per-thread sums in tensor memory: acc[0..15] += t
__device__ __forceinline__ void tmem_add16(uint32_t taddr, const float (&t)[16]) {
    float acc[16];
    tmem_ld16(taddr, acc);
#pragma unroll
    for (int i = 0; i < 16; i += 2) fadd2(acc[i], acc[i + 1], t[i], t[i + 1]);
    tmem_st16(taddr, acc);
}

__global__ void __launch_bounds__(X_THREADS, 1) edge_backward_tc_kernel(const EdgeBwdArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, grp = tid >> 7, ht = tid & 127, lane = tid & 31, hw = ht >> 5;
    float *spar = reinterpret_cast<float *>(base + XS_PAR);
    const float *sb2 = spar, *slng = spar + 32, *slnb = spar + 64, *sbc1 = spar + 96, *swc2 = spar + 128;
    float *swc1f = spar + 160;                                      // Wc1 fp32 [out][in] (epilogue: d ln beta)
    uint8_t *gb = base + XS_GRP + grp * XG_SIZE;
    uint8_t *bufA = gb, *bufB = gb + XG_TILE, *bufC = gb + 2 * XG_TILE, *bufD = gb + 3 * XG_TILE;
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + XS_MISC);
    const uint32_t mbar = smem_u32(base + XS_MISC + 16 + 8 * grp);
    const int bar_id = 1 + grp;
    (void)slnb;

    // ---- one-time setup: weight tiles as bf16 terms, row = output o: [b0 (32) | b1 (32)]; Wg: [b0 | b1 | b2] of 16 ----
    for (int i = tid; i < 1024; i += X_THREADS) {
        const int o = i >> 5, k = i & 31;
        {
            const float w = __ldg(a.pack + OFF_W2F + i);
            const float h0 = trunc16(w);
            *reinterpret_cast<uint16_t *>(base + XS_W + 4096 + sw128_off_bf16(o, k)) = (uint16_t)(__float_as_uint(h0) >> 16);
            *reinterpret_cast<uint16_t *>(base + XS_W + 4096 + sw128_off_bf16(o, 32 + k)) = (uint16_t)(__float_as_uint(w - h0) >> 16);
        }
        {
            const float w = __ldg(a.pack + OFF_WC1 + i);
            swc1f[i] = w;
            const float h0 = trunc16(w);
            *reinterpret_cast<uint16_t *>(base + XS_W + 8192 + sw128_off_bf16(o, k)) = (uint16_t)(__float_as_uint(h0) >> 16);
            *reinterpret_cast<uint16_t *>(base + XS_W + 8192 + sw128_off_bf16(o, 32 + k)) = (uint16_t)(__float_as_uint(w - h0) >> 16);
        }
        if (k < 16) {
            float w = 0.f;
            if (k < 12) w = __ldg(a.pack + OFF_WG + 32 * k + o);
            else if (k == 12) w = __ldg(a.pack + OFF_WEA + o);
            const float h0 = trunc16(w), r0 = w - h0, h1 = trunc16(r0);
            *reinterpret_cast<uint16_t *>(base + XS_W + sw128_off_bf16(o, k)) = (uint16_t)(__float_as_uint(h0) >> 16);
            *reinterpret_cast<uint16_t *>(base + XS_W + sw128_off_bf16(o, 16 + k)) = (uint16_t)(__float_as_uint(h1) >> 16);
            *reinterpret_cast<uint16_t *>(base + XS_W + sw128_off_bf16(o, 32 + k)) = (uint16_t)(__float_as_uint(r0 - h1) >> 16);
            *reinterpret_cast<uint16_t *>(base + XS_W + sw128_off_bf16(o, 48 + k)) = 0;
        }
    }
    for (int i = tid; i < 512; i += X_THREADS) reinterpret_cast<uint16_t *>(base + XS_ONES)[i] = 0x3F80;
    if (tid < 32) {
        spar[tid] = __ldg(a.pack + OFF_B2 + tid);
        spar[32 + tid] = __ldg(a.pack + OFF_LNG + tid);
        spar[64 + tid] = __ldg(a.pack + OFF_LNB + tid);
        spar[96 + tid] = __ldg(a.pack + OFF_BC1 + tid);
        spar[128 + tid] = __ldg(a.pack + OFF_WC2 + tid);
    }
    if (tid < 32) tmem_alloc(smem_u32(tmem_holder), 512);
    if (ht == 0) {
        mbar_init(mbar, 1); mbar_init(mbar + 8 * XG, 1); mbar_init(mbar + 16 * XG, 1); mbar_init(mbar + 24 * XG, 1);
        fence_mbar_init();
    }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int hw_u = __shfl_sync(0xffffffffu, hw, 0), grp_u = __shfl_sync(0xffffffffu, grp, 0);
    const uint32_t tmem0 = __shfl_sync(0xffffffffu, *tmem_holder, 0);
    const uint32_t lane_base = (uint32_t)(hw * 32) << 16;
    const uint32_t tD = tmem0 + XT_D + 32u * grp_u;                                      // MMA destination (lane 0)
    const uint32_t tDw = tD + lane_base;                                                 // this warp's lanes of it
    const uint32_t tPriv = tmem0 + XT_PRIV + 64u * grp_u + lane_base;
    const uint32_t tWG = tmem0 + XT_WG + (grp_u == 2 ? 96u : 0u) + (grp_u == 1 ? (16u << 16) : 0u);
    {   // zero the private sums and the weight-gradient accumulators: columns 96 .. 479, 128 per group
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
#pragma unroll
        for (int cc = 0; cc < 128; cc += 16) tmem_st16(tmem0 + 96u + 128u * grp_u + cc + lane_base, z);
        tmem_wait_st();
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // Shared-memory descriptors: only two 32-bit words stay live across the tile loop (the low words of the weight-tile
    // and group-tile descriptors; the high word is a constant).  Every MMA rebuilds its operands from them inside the
    // issuing block -- as loop invariants the ~230 distinct 64-bit descriptors of a tile were hoisted and spilled.
    const uint32_t w_lo = (uint32_t)make_desc_sw128(__shfl_sync(0xffffffffu, smem_u32(base + XS_W), 0));
    const uint32_t g_lo = (uint32_t)make_desc_sw128(__shfl_sync(0xffffffffu, smem_u32(gb), 0));
    constexpr uint32_t DESC_HI = (uint32_t)(((uint64_t)(1024 >> 4) << 32 | (uint64_t)1 << 46 | (uint64_t)2 << 61) >> 32);
    constexpr uint32_t ONES_HI = (uint32_t)(((uint64_t)(128 >> 4) << 32 | (uint64_t)1 << 46) >> 32);    // K-major, no swizzle
    constexpr uint32_t O_WG = 0, O_W2 = 4096 >> 4, O_WC1 = 8192 >> 4, O_ONES = (XS_ONES - XS_W) >> 4;
    constexpr uint32_t O_A = 0, O_B = XG_TILE >> 4, O_C = 2 * (XG_TILE >> 4), O_D = 3 * (XG_TILE >> 4);
    auto opaque = [](uint32_t x) { asm volatile("" : "+r"(x)); return x; };
    auto dsc64 = [](uint32_t lo) { return ((uint64_t)DESC_HI << 32) | lo; };
    const uint32_t mbar_u = __shfl_sync(0xffffffffu, mbar, 0);
    // mbarriers of the group: [0] per-edge products (six phases per tile), [1..3] the weight-gradient batches of stages
    // 4 / 5 / 6 (one phase per tile each), waited only where their operand tiles are about to be overwritten
    const uint32_t mbarA = mbar + 8 * XG, mbarB = mbar + 16 * XG, mbarC = mbar + 24 * XG;
    const uint32_t mbarA_u = mbar_u + 8 * XG, mbarB_u = mbar_u + 16 * XG, mbarC_u = mbar_u + 24 * XG;
    uint32_t phase = 0, phase_w = 0;
    // per-thread scratch in L2 (dSiLU/dpre: chunks 0..7, LayerNorm uh: chunks 8..15), CHUNK-major: chunk i of thread t at
    // (i * X_THREADS + t) * 16 bytes, so one warp-wide 128-bit access covers 512 contiguous bytes (4 L1 wavefronts; with a
    // row per thread every access touched 32 lines and the stash alone was a third of the kernel's L1 wavefronts)
    float *stash = a.stash + (size_t)blockIdx.x * X_THREADS * 64 + (size_t)tid * 4;
    constexpr int SCH = X_THREADS * 4;          // floats between consecutive chunks of a thread
    // row of this thread / rows this lane fills in the coalesced gather, inside a 128 x 128-byte tile
    const uint32_t own_row = (uint32_t)(ht * 128), own_sw = (uint32_t)(ht & 7);
    const uint32_t bufB_s = smem_u32(bufB), bufC_s = smem_u32(bufC);

    const int64_t E = __ldg(a.csr_ptr + a.num_nodes);
    const int64_t T = (E + 127) / 128;
    const int64_t NG = (int64_t)gridDim.x * XG, gi = (int64_t)blockIdx.x * XG + grp;
#ifdef EGSPR_XB_SOLO      // developer experiment: only group 1 of every CTA works (uncontended per-stage latencies)
    const int64_t tile0 = T * gi / NG, tile1 = grp == 1 ? T * (gi + 1) / NG : tile0;
#else
    const int64_t tile0 = T * gi / NG, tile1 = T * (gi + 1) / NG;
#endif

    // per-tile edge state, fetched ONE TILE AHEAD (indices at the top of the previous tile, coordinates in its middle)
    struct EdgeIn { int r, c; float ea; bool valid; float xr[3], xc[3], dxo[3]; };
    auto load_indices = [&](int64_t tile, int &r_, int &c_, int &eid_, bool &valid_) {
        const int64_t p0 = tile * 128 + ht;
        valid_ = p0 < E;
        const int64_t p = valid_ ? p0 : E - 1;          // idle slots redo the last edge with zero upstream gradient
        r_ = ldg_now(a.csr_row + p); c_ = ldg_now(a.csr_col + p);
        eid_ = a.edge_attr ? ldg_now(a.csr_eid + p) : 0;              // the original edge id only addresses a per-edge edge_attr
    };
    auto load_coords = [&](EdgeIn &e, int eid_) {
        e.ea = a.edge_attr ? ldg_now(a.edge_attr + (int64_t)(e.r / a.n_per_cloud) * a.edges_per_cloud + eid_) : a.edge_attr_const;
        const float4 t0 = ldg4_now(a.x4 + (int64_t)e.r * 4), t1 = ldg4_now(a.x4 + (int64_t)e.c * 4);
        e.xr[0] = t0.x; e.xr[1] = t0.y; e.xr[2] = t0.z; e.xc[0] = t1.x; e.xc[1] = t1.y; e.xc[2] = t1.z;
        e.dxo[0] = e.dxo[1] = e.dxo[2] = 0.f;
        if (e.valid) {
            e.dxo[0] = ldg_now(a.dx_out + (int64_t)e.r * 3); e.dxo[1] = ldg_now(a.dx_out + (int64_t)e.r * 3 + 1);
            e.dxo[2] = ldg_now(a.dx_out + (int64_t)e.r * 3 + 2);
        }
    };
    EdgeIn nx;
    nx.r = nx.c = 0; nx.valid = false;
    if (tile0 < tile1) {
        int eid0;
        load_indices(tile0, nx.r, nx.c, eid0, nx.valid);
        load_coords(nx, eid0);
    }
#ifdef EGSPR_XB_TIMING
    int tile_no = -1;
#endif
    for (int64_t tile = tile0; tile < tile1; ++tile) {
#ifdef EGSPR_XB_TIMING
        ++tile_no;
#endif
        const EdgeIn cur = nx;
        XB_MARK(0);
        const int r = cur.r, c = cur.c;
        const bool valid = cur.valid;
        // ---- this tile's P[row] / Q[col] rows -> shared memory with cp.async (no registers, a tile's worth of latency
        // hidden behind the geometry), both gathered COALESCED: 8 lanes per 128-byte row, 4 rows per warp-wide instruction
        // (4 L1 wavefronts).  With every thread copying its own P row, an instruction wrote 32 different shared-memory rows
        // = 32 wavefronts, and those 8 instructions were 36 % of the kernel's shared-memory wavefronts (ncu, source page).
        // Q -> tile C, P -> tile B; both tiles were last read by the previous tile's stage-5 weight-gradient MMAs.
        if (tile > tile0) { mbar_wait(mbarB, phase_w ^ 1); fence_after_sync(); }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c_src = __shfl_sync(0xffffffffu, c, 4 * i + (lane >> 3));
            const int r_src = __shfl_sync(0xffffffffu, r, 4 * i + (lane >> 3));
            const uint32_t row = (uint32_t)(hw * 32 + 4 * i + (lane >> 3));
            const uint32_t off = row * 128 + ((((uint32_t)lane & 7u) ^ (row & 7u)) << 4);
            cp_async16_cg(bufC_s + off, a.Q + (int64_t)c_src * H + 4 * (lane & 7));
            cp_async16_cg(bufB_s + off, a.P + (int64_t)r_src * H + 4 * (lane & 7));
        }
        cp_async_commit();
        int rn = 0, cn = 0, eidn = 0;
        const bool more = tile + 1 < tile1;
        if (more) load_indices(tile + 1, rn, cn, eidn, nx.valid);
        float dsc;                                      // d loss / d s  (trans = coord_diff * s, 3dm:264)
        // ---------------- stage 1: pre = P[row] + Q[col] + [geo | edge_attr] Wg^T ----------------
        {
            EdgeGeo g;
            float geo[16];
            edge_geometry(cur.xr, cur.xc, g, geo);
            geo[12] = cur.ea; geo[13] = 0.f; geo[14] = 0.f; geo[15] = 0.f;
            dsc = g.d[0] * cur.dxo[0] + g.d[1] * cur.dxo[1] + g.d[2] * cur.dxo[2];
            XB_MARK(1);
            if (tile > tile0) { mbar_wait(mbarC, phase_w ^ 1); fence_after_sync(); }    // stage-6 batch of the previous tile read tile A
            write_row_split3(bufA, ht, geo);
        }
        fence_proxy_async();
        fence_before_sync();
        XB_MARK(2);
        bar_sync(bar_id, 128);
        XB_MARK(3);
        if (hw_u == 0 && elect_one()) {
            fence_after_sync();
            // geo terms (K-steps 0..2) x weight terms (offsets 0 / 32 / 64 B): all products down to 2^-24
            const uint32_t ga = opaque(g_lo) + O_A, wg = opaque(w_lo) + O_WG;
            umma_bf16_ss(tD, dsc64(ga + 0), dsc64(wg + 0), ID_KK_N32, 0);
            umma_bf16_ss(tD, dsc64(ga + 0), dsc64(wg + 2), ID_KK_N32, 1);
            umma_bf16_ss(tD, dsc64(ga + 2), dsc64(wg + 0), ID_KK_N32, 1);
            umma_bf16_ss(tD, dsc64(ga + 0), dsc64(wg + 4), ID_KK_N32, 1);
            umma_bf16_ss(tD, dsc64(ga + 2), dsc64(wg + 2), ID_KK_N32, 1);
            umma_bf16_ss(tD, dsc64(ga + 4), dsc64(wg + 0), ID_KK_N32, 1);
            umma_commit(mbar_u);
        }
        float v[32];
        cp_async_wait_all();
        __syncwarp();                                   // the P / Q rows of this warp's edges were written by its other lanes
        XB_MARK(4);
        mbar_wait(mbar, phase); phase ^= 1;
        XB_MARK(5);
        fence_after_sync();
        tmem_ld32(tDw, v);
        {
            float ds1[32];
#pragma unroll
            for (int i = 0; i < 8; ++i) {               // a1 = SiLU(pre);  ds1 = d SiLU / d pre
                const float4 pv = *reinterpret_cast<const float4 *>(bufB + own_row + (((uint32_t)i ^ own_sw) << 4));
                const float4 qv = *reinterpret_cast<const float4 *>(bufC + own_row + (((uint32_t)i ^ own_sw) << 4));
                fadd2(v[4 * i], v[4 * i + 1], pv.x, pv.y); fadd2(v[4 * i + 2], v[4 * i + 3], pv.z, pv.w);
                fadd2(v[4 * i], v[4 * i + 1], qv.x, qv.y); fadd2(v[4 * i + 2], v[4 * i + 3], qv.z, qv.w);
#pragma unroll
                for (int j = 4 * i; j < 4 * i + 4; j += 2) {
                    float s0, s1;
                    sigmoid2(v[j], v[j + 1], s0, s1);
                    ds1[j] = s0 * (1.0f + v[j] * (1.0f - s0)); ds1[j + 1] = s1 * (1.0f + v[j + 1] * (1.0f - s1));
                    fmul2(v[j], v[j + 1], s0, s1);
                }
            }
            // dSiLU/dpre is needed again only in stage 5: parked in this thread's 128-byte slot of an L2-resident scratch
            // (8 vector stores + 8 vector loads; keeping it in registers next to uh and the working row spilled ~180 words)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                __stcg(reinterpret_cast<float4 *>(stash + i * SCH), make_float4(ds1[4 * i], ds1[4 * i + 1], ds1[4 * i + 2], ds1[4 * i + 3]));
        }
        // ---------------- stage 2: u = a1 W2^T + b2 (block diagonal), m = LayerNorm(u) ----------------
        XB_MARK(6);
        write_row_split2(bufB, ht, v);
        fence_proxy_async();
        fence_before_sync();
        XB_MARK(7);
        bar_sync(bar_id, 128);
        XB_MARK(8);
        if (hw_u == 1 && elect_one()) {
            fence_after_sync();
            const uint32_t gB = opaque(g_lo) + O_B, w2 = opaque(w_lo) + O_W2;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4 - 2 * q; ++j) umma_bf16_ss(tD, dsc64(gB + 2 * j), dsc64(w2 + 4 * q + 2 * (j & 1)), ID_KK_N32, (q | j) > 0);
            umma_commit(mbar_u);
        }
        if (more) { nx.r = rn; nx.c = cn; load_coords(nx, eidn); }      // the next tile's coordinates (its indices have arrived)
        mbar_wait(mbar, phase); phase ^= 1;
        XB_MARK(9);
        fence_after_sync();
        tmem_ld32(tDw, v);
        float rstd;
        {
            float m4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 b = *reinterpret_cast<const float4 *>(sb2 + j);
                fadd2(v[j], v[j + 1], b.x, b.y); fadd2(v[j + 2], v[j + 3], b.z, b.w);
                fadd2(m4[0], m4[1], v[j], v[j + 1]); fadd2(m4[2], m4[3], v[j + 2], v[j + 3]);
            }
            const float nmean = ((m4[0] + m4[1]) + (m4[2] + m4[3])) * (-1.0f / 32.0f);
            float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                fadd2(v[j], v[j + 1], nmean, nmean); fadd2(v[j + 2], v[j + 3], nmean, nmean);
                ffma2(q4[0], q4[1], v[j], v[j + 1], v[j], v[j + 1]); ffma2(q4[2], q4[3], v[j + 2], v[j + 3], v[j + 2], v[j + 3]);
            }
            rstd = rsqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / 32.0f) + 1e-5f);
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                const float4 gm = *reinterpret_cast<const float4 *>(slng + j), bt = *reinterpret_cast<const float4 *>(spar + 64 + j);
                fmul2(v[j], v[j + 1], rstd, rstd); fmul2(v[j + 2], v[j + 3], rstd, rstd);
                __stcg(reinterpret_cast<float4 *>(stash + (8 + j / 4) * SCH), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));   // uh: needed again in stage 4
                float o0 = bt.x, o1 = bt.y, o2 = bt.z, o3 = bt.w;
                ffma2(o0, o1, v[j], v[j + 1], gm.x, gm.y); ffma2(o2, o3, v[j + 2], v[j + 3], gm.z, gm.w);
                v[j] = o0; v[j + 1] = o1; v[j + 2] = o2; v[j + 3] = o3;
            }
        }
        // ---------------- stage 3: c = m Wc1^T + bc1, s = wc2 . SiLU(c) ----------------
        XB_MARK(10);
        write_row_split2(bufC, ht, v);
        fence_proxy_async();
        fence_before_sync();
        XB_MARK(11);
        bar_sync(bar_id, 128);
        XB_MARK(12);
        if (hw_u == 2 && elect_one()) {
            fence_after_sync();
            const uint32_t gC = opaque(g_lo) + O_C, wc = opaque(w_lo) + O_WC1;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4 - 2 * q; ++j) umma_bf16_ss(tD, dsc64(gC + 2 * j), dsc64(wc + 4 * q + 2 * (j & 1)), ID_KK_N32, (q | j) > 0);
            umma_commit(mbar_u);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        XB_MARK(13);
        fence_after_sync();
        tmem_ld32(tDw, v);
        float s = 0.f;
#pragma unroll
        for (int hb = 0; hb < 32; hb += 16) {           // dc1 = wc2 dsc dSiLU(c);  d wc2 rows a2 * dsc -> private sums
            float t[16];
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                const int o = hb + i;
                const float2 bc = *reinterpret_cast<const float2 *>(sbc1 + o), wc2v = *reinterpret_cast<const float2 *>(swc2 + o);
                fadd2(v[o], v[o + 1], bc.x, bc.y);
                float s0, s1;
                sigmoid2(v[o], v[o + 1], s0, s1);
                const float a20 = v[o] * s0, a21 = v[o + 1] * s1;
                const float w0 = wc2v.x, w1 = wc2v.y;
                s = fmaf(w0, a20, s); s = fmaf(w1, a21, s);
                t[i] = a20 * dsc; t[i + 1] = a21 * dsc;
                v[o] = w0 * dsc * (s0 * (1.0f + v[o] * (1.0f - s0)));
                v[o + 1] = w1 * dsc * (s1 * (1.0f + v[o + 1] * (1.0f - s1)));
            }
            tmem_add16(tPriv + 32 + hb, t);
        }
        // ---------------- stage 4: dm = dagg[row] + dc1 Wc1;  dWc1 += dc1^T m, dbc1 += dc1^T 1 ----------------
        XB_MARK(14);
        write_row_split2(bufD, ht, v);
        tmem_wait_st();
        fence_proxy_async();
        fence_before_sync();
        XB_MARK(15);
        bar_sync(bar_id, 128);
        XB_MARK(16);
        if (hw_u == 3 && elect_one()) {
            fence_after_sync();
            const uint32_t gD = opaque(g_lo) + O_D, gC = opaque(g_lo) + O_C, wc = opaque(w_lo) + O_WC1;
            const uint64_t ones = ((uint64_t)ONES_HI << 32) | ((opaque(w_lo) + O_ONES) | (uint32_t)(128 >> 4) << 16);
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4 - 2 * q; ++j) umma_bf16_ss(tD, dsc64(gD + 2 * j), dsc64(wc + 4 * q + 128 * (j & 1)), ID_KM_N32, (q | j) > 0);
            umma_commit(mbar_u);
#ifndef EGSPR_XB_NO_WG
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                umma_bf16_ss(tWG + XW_WC1, dsc64(gD + 128 * j), dsc64(gC + 128 * j), ID_MM_N32, 1);
                umma_bf16_ss(tWG + XW_WC1, dsc64(gD + 128 * j), dsc64(gC + 4 + 128 * j), ID_MM_N32, 1);
                umma_bf16_ss(tWG + XW_BC1, dsc64(gD + 128 * j), ones, ID_MK_N8, 1);
            }
#endif
            umma_commit(mbarA_u);
        }
        {   // upstream message gradient while the tensor core works (zero for idle slots)
            const float *dg = a.dagg + (int64_t)r * H;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 t = valid ? ldg4_now(dg + 4 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
                v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
            }
        }
        float uh[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = ldcg4_now(stash + (8 + i) * SCH);
            uh[4 * i] = t.x; uh[4 * i + 1] = t.y; uh[4 * i + 2] = t.z; uh[4 * i + 3] = t.w;
        }
        mbar_wait(mbar, phase); phase ^= 1;
        XB_MARK(17);
        fence_after_sync();
#pragma unroll
        for (int hb = 0; hb < 32; hb += 16) {
            float dm[16];
            tmem_ld16(tDw + hb, dm);
#pragma unroll
            for (int i = 0; i < 16; i += 2) fadd2(v[hb + i], v[hb + i + 1], dm[i], dm[i + 1]);
        }
        {   // d ln gamma rows dm * uh -> private sums; LayerNorm backward -> du
#pragma unroll
            for (int hb = 0; hb < 32; hb += 16) {
                float t[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) t[i] = v[hb + i] * uh[hb + i];
                tmem_add16(tPriv + hb, t);
            }
            float s1a[2] = {0.f, 0.f}, s2a[2] = {0.f, 0.f};
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float2 gmv = *reinterpret_cast<const float2 *>(slng + i);
                fmul2(v[i], v[i + 1], gmv.x, gmv.y);
                fadd2(s1a[0], s1a[1], v[i], v[i + 1]);
                ffma2(s2a[0], s2a[1], v[i], v[i + 1], uh[i], uh[i + 1]);
            }
            const float s1 = (s1a[0] + s1a[1]) * (1.0f / 32.0f), s2 = (s2a[0] + s2a[1]) * (1.0f / 32.0f);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = rstd * (v[i] - s1 - uh[i] * s2);
        }
        // ---------------- stage 5: da1 = du W2;  dW2 += du^T a1, db2 += du^T 1 ----------------
        XB_MARK(18);
        mbar_wait(mbarA, phase_w); fence_after_sync();          // the stage-4 batch read m in tile C
        XB_MARK(19);
        write_row_split2(bufC, ht, v);
        tmem_wait_st();
        fence_proxy_async();
        fence_before_sync();
        XB_MARK(20);
        bar_sync(bar_id, 128);
        XB_MARK(21);
        if (hw_u == 0 && elect_one()) {
            fence_after_sync();
            const uint32_t gC = opaque(g_lo) + O_C, gB = opaque(g_lo) + O_B, w2 = opaque(w_lo) + O_W2;
            const uint64_t ones = ((uint64_t)ONES_HI << 32) | ((opaque(w_lo) + O_ONES) | (uint32_t)(128 >> 4) << 16);
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4 - 2 * q; ++j) umma_bf16_ss(tD, dsc64(gC + 2 * j), dsc64(w2 + 4 * q + 128 * (j & 1)), ID_KM_N32, (q | j) > 0);
            umma_commit(mbar_u);
#ifndef EGSPR_XB_NO_WG
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                umma_bf16_ss(tWG + XW_W2, dsc64(gC + 128 * j), dsc64(gB + 128 * j), ID_MM_N32, 1);
                umma_bf16_ss(tWG + XW_W2, dsc64(gC + 128 * j), dsc64(gB + 4 + 128 * j), ID_MM_N32, 1);
                umma_bf16_ss(tWG + XW_B2, dsc64(gC + 128 * j), ones, ID_MK_N8, 1);
            }
#endif
            umma_commit(mbarB_u);
        }
        float4 dsv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dsv[i] = ldcg4_now(stash + i * SCH);
        mbar_wait(mbar, phase); phase ^= 1;
        XB_MARK(22);
        fence_after_sync();
        tmem_ld32(tDw, v);
#pragma unroll
        for (int i = 0; i < 8; ++i) {                                                     // dpre = dSiLU(pre) * da1
            fmul2(v[4 * i], v[4 * i + 1], dsv[i].x, dsv[i].y); fmul2(v[4 * i + 2], v[4 * i + 3], dsv[i].z, dsv[i].w);
        }
        // ---------------- stage 6: d geo = dpre Wg;  dWg += dpre^T [geo | edge_attr] ----------------
        // (tile D: its stage-4 readers finished before the stage-5 products that were just waited for)
        XB_MARK(23);
        write_row_split2(bufD, ht, v);
        fence_proxy_async();
        fence_before_sync();
        XB_MARK(24);
        bar_sync(bar_id, 128);
        XB_MARK(25);
        if (hw_u == 1 && elect_one()) {
            fence_after_sync();
            const uint32_t gD = opaque(g_lo) + O_D, ga = opaque(g_lo) + O_A, wg = opaque(w_lo) + O_WG;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int j = 0; j < 4 - 2 * q; ++j) umma_bf16_ss(tD, dsc64(gD + 2 * j), dsc64(wg + 2 * q + 128 * (j & 1)), ID_KM_N16, (q | j) > 0);
            umma_commit(mbar_u);
#ifndef EGSPR_XB_NO_WG
#pragma unroll
            for (int j = 0; j < 8; ++j)
#pragma unroll
                for (int q = 0; q < 3; ++q) umma_bf16_ss(tWG + XW_WGEO, dsc64(gD + 128 * j), dsc64(ga + 2 * q + 128 * j), ID_MM_N16, 1);
#endif
            umma_commit(mbarC_u);
        }
        {   // dpre = the gradient of P[row] and of Q[col], one row per edge in ROW-CSR ORDER (the gather kernel reads a node's
            // row list as one contiguous run and its col list through csc_pos).  Written COALESCED: 8 lanes per 128-byte row
            // (4 full rows per warp-wide store) out of the b0 + b1 terms this warp just wrote to tile D (16 mantissa bits)
            // -- a row per thread touched 32 lines per store instruction.
            const int j = lane & 7;
            const int64_t pw = tile * 128 + hw * 32;                       // first edge position of this warp
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                // rows of one instruction: within each HALF-warp (a 64-bit access is served per half-warp) one row with
                // (row & 4) == 0 and one with (row & 4) != 0, so that their swizzled 64-byte halves cover all 32 banks
                const int rl = 8 * (it >> 1) + 2 * (it & 1) + 4 * ((lane >> 3) & 1) + (lane >> 4);
                const int row = hw * 32 + rl;
                const uint8_t *rp = bufD + row * 128 + (j & 1) * 8;
                const uint2 t0 = *reinterpret_cast<const uint2 *>(rp + (((j >> 1) ^ (row & 7)) << 4));
                const uint2 t1 = *reinterpret_cast<const uint2 *>(rp + ((((j >> 1) + 4) ^ (row & 7)) << 4));
                float4 o;
                o.x = __uint_as_float(t0.x << 16) + __uint_as_float(t1.x << 16);
                o.y = __uint_as_float(t0.x & 0xffff0000u) + __uint_as_float(t1.x & 0xffff0000u);
                o.z = __uint_as_float(t0.y << 16) + __uint_as_float(t1.y << 16);
                o.w = __uint_as_float(t0.y & 0xffff0000u) + __uint_as_float(t1.y & 0xffff0000u);
                if (pw + rl < E) *reinterpret_cast<float4 *>(a.dpre + (pw + rl) * H + 4 * j) = o;
            }
        }
        mbar_wait(mbar, phase); phase ^= 1;
        XB_MARK(26);
        fence_after_sync();
        {
            float gg[16];
            tmem_ld16(tDw, gg);
            EdgeGeo g;
            float geo[12];
            edge_geometry(cur.xr, cur.xc, g, geo);
            const float gde[3] = {s * cur.dxo[0], s * cur.dxo[1], s * cur.dxo[2]};
            float dxr[3], dxc[3];
            edge_geometry_backward(cur.xr, cur.xc, g, gg, gde, dxr, dxc);
            if (valid) {
                const int64_t pe = tile * 128 + ht;                         // row-CSR position of this thread's edge
                *reinterpret_cast<float4 *>(a.dxe + pe * 8) = make_float4(dxr[0], dxr[1], dxr[2], 0.f);
                *reinterpret_cast<float4 *>(a.dxe + pe * 8 + 4) = make_float4(dxc[0], dxc[1], dxc[2], 0.f);
            }
        }
        XB_MARK(27);
        phase_w ^= 1;
        fence_before_sync();        // the tcgen05.ld above is ordered before the next tile's first MMA by its barrier
    }
    if (tile0 < tile1) {            // the last tile's weight-gradient batches
        mbar_wait(mbarB, phase_w ^ 1);
        mbar_wait(mbarC, phase_w ^ 1);
        fence_after_sync();
    }

    // ---- epilogue: private sums and the weight-gradient accumulators -> gradient pack ----
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    // every (group, b0 / b1 term) writes its accumulator rows to its own slot, the slots are summed on the way out
    float *red = reinterpret_cast<float *>(base + XS_GRP);        // tiles are free now
    constexpr int R_WC1 = 0, R_W2 = 1024, R_WG = 2048, R_BC1 = 2560, R_B2 = 2592, R_N = 2624, R_SLOTS = 2 * XG;
    float *colp = red + R_SLOTS * R_N;                            // [warps][64]: d ln gamma | d wc2 column sums per warp
    {
        float t[32];
        tmem_ld32(tPriv, t);
        const float cs0 = warp_colsum32(t);
        tmem_ld32(tPriv + 32, t);
        const float cs1 = warp_colsum32(t);
        colp[(tid >> 5) * 64 + lane] = cs0;
        colp[(tid >> 5) * 64 + 32 + lane] = cs1;
    }
    {
        // M = 64 accumulator layout: row m = 16 * warp + (lane & 15), at lanes 0..15 of every warp quarter (group 1: 16..31);
        // rows 0..31 = b0 term of feature m, rows 32..63 = b1 term of feature m - 32
        const bool mine = ((lane >> 4) == (grp == 1 ? 1 : 0));
        const int m = 16 * hw + (lane & 15), feat = m & 31;
        float *slot = red + (2 * grp + (m >> 5)) * R_N;
        const uint32_t tw = tmem0 + XT_WG + (grp == 2 ? 96u : 0u) + lane_base;
        float t[32];
        tmem_ld32(tw + XW_WC1, t);
        if (mine) {
#pragma unroll
            for (int i = 0; i < 32; i += 4)                                                   // dWc1[o = feat][i]
                *reinterpret_cast<float4 *>(slot + R_WC1 + 32 * feat + i) = make_float4(t[i], t[i + 1], t[i + 2], t[i + 3]);
        }
        tmem_ld32(tw + XW_W2, t);
        if (mine) {
#pragma unroll
            for (int i = 0; i < 32; ++i) slot[R_W2 + 32 * feat + i] = t[i];                  // full 32 x 32, block diagonal picked below
        }
        tmem_ld32(tw + XW_WGEO, t);          // 16 geo columns | dbc1 (8 equal columns) | db2 (8 equal columns)
        if (mine) {
#pragma unroll
            for (int k = 0; k < 16; ++k) slot[R_WG + 32 * k + feat] = t[k];                  // rows of Wg, row 12 = edge_attr
            slot[R_BC1 + feat] = t[16];
            slot[R_B2 + feat] = t[24];
        }
    }
    fence_before_sync();
    __syncthreads();
    // every CTA starts at a different entry: all CTAs reach this loop together, and 148 atomics onto the same address in
    // the same order serialise in L2
    const int rot = (int)((blockIdx.x * 211u) % (unsigned)R_N);
    for (int i0 = tid; i0 < R_N; i0 += X_THREADS) {
        const int i = (i0 + rot) % R_N;
        float val = 0.f;
#pragma unroll
        for (int sl = 0; sl < R_SLOTS; ++sl) val += red[sl * R_N + i];
        int dst = -1;
        if (i < R_W2) dst = OFF_WC1 + i;
        else if (i < R_WG) dst = OFF_W2F + (i - R_W2);       // dW2[o][i2], the full matrix (the host keeps the heads' diagonal blocks)
        else if (i < R_WG + 384) dst = OFF_WG + (i - R_WG);
        else if (i < R_WG + 416) dst = OFF_WEA + (i - R_WG - 384);
        else if (i < R_BC1) dst = -1;                          // geo rows 13..15 are padding
        else if (i < R_B2) { dst = OFF_BC1 + (i - R_BC1); red[i] = val; }      // kept for the d ln beta product below
        else dst = OFF_B2 + (i - R_B2);
        if (dst >= 0) atomicAdd(a.gpack + dst, val);
    }
    if (tid < 64) {
        float val = 0.f;
#pragma unroll
        for (int w = 0; w < X_THREADS / 32; ++w) val += colp[w * 64 + tid];
        atomicAdd(a.gpack + (tid < 32 ? OFF_LNG + tid : OFF_WC2 + tid - 32), val);
    }
    __syncthreads();
    if (tid < 32) {      // d ln beta, coord-MLP part: (sum_e dc1_e) Wc1  (the dagg part comes from node_mlp_backward_kernel)
        float acc = 0.f;
#pragma unroll 8
        for (int o = 0; o < 32; ++o) acc = fmaf(red[R_BC1 + o], swc1f[32 * o + tid], acc);     // slot 0 now holds the CTA total
        atomicAdd(a.gpack + OFF_LNB + tid, acc);
    }
    __syncthreads();
    if (tid < 32) tmem_dealloc(tmem0, 512);
}

size_t edge_backward_tc_stash_bytes() { return (size_t)sm_count() * X_THREADS * 64 * sizeof(float); }

#ifdef EGSPR_XB_TIMING
extern "C" int egspr_debug_read_xb(long long *host_dst) {
    return cudaMemcpyFromSymbol(host_dst, g_xb_dbg, sizeof(long long) * 4 * 64 * 32) == cudaSuccess ? 0 : -4;
}
#endif

int launch_edge_backward_tc(const EdgeBwdArgs &a, cudaStream_t st) {
    if (!opt_in_smem(edge_backward_tc_kernel, X_SMEM_BYTES)) return EGSPR_E_LAUNCH;
    const int64_t E = (a.num_nodes / a.n_per_cloud) * a.edges_per_cloud;
    const int64_t tiles = (E + 127) / 128;
    int64_t grid = sm_count();
    const int64_t need = (tiles + XG - 1) / XG;        // at least one tile per group
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    edge_backward_tc_kernel<<<(unsigned)grid, X_THREADS, X_SMEM_BYTES, st>>>(a);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

}  // namespace egspr
