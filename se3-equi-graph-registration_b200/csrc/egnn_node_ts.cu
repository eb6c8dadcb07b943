// Per-node dense stages of the EGNN on tcgen05 (A operand through TMEM, 3xTF32 = fp32-level accuracy):
//   node kernel  (E_GCL.node_model, src/3dmatch_train_egnn_with_batch.py:252-260, + what follows it)
//       t      = SiLU([h | agg] Wn1^T + bn1)                 M128 N32 K64
//       h'     = h + t Wn2^T + bn2                           M128 N32 K32      (residual :258-259)
//       P', Q' = h' Wp^T, h' Wq^T + b1                       M128 N64 K32      (next layer's first edge Linear,
//                                                                               node halves; egnn_layer.cuh)
//       or  out = h' Wout^T + bout  after the last layer (embedding_out, :337)
//   embed kernel (EGNN.forward :332)   h = feat Win^T + bin, then P, Q of layer 0; x [G,3] -> x4 [G,4]
// One thread = one node = one TMEM lane; a group of 128 threads owns 160 TMEM columns
// (D 0..63 | A_hi 32..95 or 64..95 | A_lo 96..159, see the stage comments); 3 groups per CTA share the
// weight tiles.  Replaces the one-thread-per-node CUDA-core kernels (6400 issue slots per node -> ~700).
#include "egnn_layer.cuh"
#include "egnn_backward.cuh"
#include "tcgen05.cuh"

namespace egspr {
using namespace tc;

constexpr int NT_GROUPS = 3;
constexpr int NT_THREADS = 128 * NT_GROUPS;
constexpr int NT_COLS = 160;          // TMEM columns per group

struct NodeTsArgs {
    const float *h_in;       // [G][32] layer input features (embed mode: raw input features)
    const float *agg;        // [G][32] aggregated messages (NULL: no node MLP -- embed mode)
    const float *x3;         // embed mode: [G][3] coordinates to pad into x4_out (or NULL)
    float *x4_out;
    const float *w1t, *b1;   // node_mlp.0: [64][32] in-major, [32]
    const float *w2t, *b2;   // node_mlp.2 (or embedding_in): [32][32] in-major, [32]; NULL: h' = h_in
    const float *w3pt, *w3qt, *b3;   // stage 3: [32][32] in-major each (w3qt NULL -> N = 32), bias of the LAST 32 outputs
    float *h_out, *P_out, *Q_out;
    int64_t G;
    int residual;            // h' = h + ...
    int out_to_h;            // stage 3 (N = 32) result goes to h_out (embedding_out); h' itself is not stored
};

// shared memory (bytes from a 1024-aligned base): SW128 K-major weight tiles
constexpr int NS_W1A = 0;            // Wn1[:, 0:32]  hi, lo   (2 x 4 KB)
constexpr int NS_W1B = 8192;         // Wn1[:, 32:64] hi, lo
constexpr int NS_W2 = 16384;         // Wn2 / Win     hi, lo
constexpr int NS_W3 = 24576;         // [Wp ; Wq] 64 rows hi (8 KB), lo (8 KB)
constexpr int NS_PAR = 40960;        // b1[32], b2[32], b3[64]
constexpr int NS_MBAR = NS_PAR + 512;
constexpr int NS_TMEM = NS_MBAR + 8 * NT_GROUPS;
constexpr int NT_SROW = 36;           // floats per staged row (144 B: conflict-free 16-byte accesses)
constexpr int NS_STAGE = ((NS_TMEM + 16 + 127) / 128) * 128;     // per warp: 3 x [32][36] floats (h rows, agg rows, store staging)
constexpr int NT_WSTAGE = 3 * 32 * NT_SROW * 4;
constexpr int NS_END = NS_STAGE + (NT_THREADS / 32) * NT_WSTAGE;
constexpr size_t NT_SMEM_BYTES = NS_END + 1024;

constexpr uint32_t IDESC_TF32_M128_N64 = (1u << 4) | (2u << 7) | (2u << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);

__device__ __forceinline__ void nt_store_hilo(uint32_t t_hi, uint32_t t_lo, const float (&v)[32]) {
#pragma unroll
    for (int b = 0; b < 2; ++b) {
        float hi[16], lo[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { hi[i] = tf32_hi(v[16 * b + i]); lo[i] = v[16 * b + i] - hi[i]; }
        tmem_st16(t_hi + 16 * b, hi);
        tmem_st16(t_lo + 16 * b, lo);
    }
}
// D (+)= A_hi W_hi^T + A_lo W_hi^T + A_hi W_lo^T over one K = 32 block
__device__ __forceinline__ void nt_issue_k32(uint32_t d, uint32_t ahi, uint32_t alo, uint64_t whi, uint64_t wlo, uint32_t idesc, bool first) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, ahi + 8 * k, whi + 2 * k, idesc, !(first && k == 0));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, alo + 8 * k, whi + 2 * k, idesc, 1);
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_tf32_ts(d, ahi + 8 * k, wlo + 2 * k, idesc, 1);
}

__device__ __forceinline__ void nt_cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// The 32 rows of a warp are CONTIGUOUS in global memory (4 KB).  Moving them one row per thread (8 x LDG.128,
// every instruction touching 32 different lines) costs 32 L1 wavefronts per instruction; moved as 8 lanes per
// row through a shared-memory staging tile it costs 4.
__device__ __forceinline__ void nt_rows_async(uint32_t stage_s, const float *src, int64_t g0, int64_t G, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + (lane >> 3);
        const int64_t g = min(g0 + row, G - 1);
        nt_cp_async16(stage_s + (uint32_t)(row * (NT_SROW * 4) + (lane & 7) * 16), src + g * H + 4 * (lane & 7));
    }
}
__device__ __forceinline__ void nt_row_from_stage(float (&v)[32], const float *stage, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 t = *reinterpret_cast<const float4 *>(stage + lane * NT_SROW + 4 * i);
        v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
}
// registers (one row per thread) -> staging -> coalesced global store of the warp's rows [g0, g0 + 32) below G
__device__ __forceinline__ void nt_rows_store(const float (&v)[32], float *stage, float *dst, int64_t g0, int64_t G, int lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
        *reinterpret_cast<float4 *>(stage + lane * NT_SROW + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = 4 * i + (lane >> 3);
        if (g0 + row < G)
            *reinterpret_cast<float4 *>(dst + (g0 + row) * H + 4 * (lane & 7)) =
                *reinterpret_cast<const float4 *>(stage + row * NT_SROW + 4 * (lane & 7));
    }
    __syncwarp();
}

__device__ __forceinline__ void nt_fill_tile(uint8_t *hi_t, uint8_t *lo_t, const float *wt_in_major, int rows, int tid) {
    // B operand: row = output o, K = input k; source is [in][out] with 32 outputs per input row
    for (int i = tid; i < rows * 32; i += NT_THREADS) {
        const int o = i >> 5, k = i & 31;
        const float w = __ldg(wt_in_major + 32 * k + (o & 31));
        const float hi = tf32_hi(w);
        *reinterpret_cast<float *>(hi_t + sw128_off(o, k)) = hi;
        *reinterpret_cast<float *>(lo_t + sw128_off(o, k)) = w - hi;
    }
}

__global__ void __launch_bounds__(NT_THREADS, 1) egnn_node_ts_kernel(const NodeTsArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, grp = tid >> 7, ht = tid & 127, hw = ht >> 5;
    float *par = reinterpret_cast<float *>(base + NS_PAR);
    const uint32_t mbar = smem_u32(base + NS_MBAR + 8 * grp);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + NS_TMEM);
    const int lane = tid & 31;
    float *stA = reinterpret_cast<float *>(base + NS_STAGE + (tid >> 5) * NT_WSTAGE), *stB = stA + 32 * NT_SROW, *stC = stB + 32 * NT_SROW;
    const uint32_t stA_s = smem_u32(stA), stB_s = smem_u32(stB);
    const bool do_s1 = a.agg != nullptr, do_s2 = a.w2t != nullptr;
    const int n3 = a.w3pt ? (a.w3qt ? 64 : 32) : 0;
    pdl_trigger();                  // the next kernel's prologue may overlap this kernel's tail

    // ---- weight tiles (hi / lo, swizzled) ----
    if (do_s1) {
        nt_fill_tile(base + NS_W1A, base + NS_W1A + 4096, a.w1t, 32, tid);
        nt_fill_tile(base + NS_W1B, base + NS_W1B + 4096, a.w1t + 32 * 32, 32, tid);
    }
    if (do_s2) nt_fill_tile(base + NS_W2, base + NS_W2 + 4096, a.w2t, 32, tid);
    if (n3) {
        nt_fill_tile(base + NS_W3, base + NS_W3 + 8192, a.w3pt, 32, tid);
        if (n3 == 64) nt_fill_tile(base + NS_W3 + 4096, base + NS_W3 + 8192 + 4096, a.w3qt, 32, tid);
    }
    if (tid < 32) {
        par[tid] = do_s1 ? __ldg(a.b1 + tid) : 0.f;
        par[32 + tid] = do_s2 ? __ldg(a.b2 + tid) : 0.f;
        // stage-3 bias: P half has none, Q half carries the first edge Linear's bias; N = 32: embedding_out bias
        par[64 + tid] = (n3 == 32 && a.b3) ? __ldg(a.b3 + tid) : 0.f;
        par[96 + tid] = (n3 == 64 && a.b3) ? __ldg(a.b3 + tid) : 0.f;
    }
    if (tid < 32) tmem_alloc(smem_u32(tmem_holder), 512);
    if (ht == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    pdl_wait();                     // everything above touched only static weights / on-chip state
    // warp-uniform copies (shuffle from lane 0) so the MMA issue code runs on the uniform datapath
    const int hw_u = __shfl_sync(0xffffffffu, hw, 0), grp_u = __shfl_sync(0xffffffffu, grp, 0);
    const uint32_t tg = __shfl_sync(0xffffffffu, *tmem_holder, 0) + NT_COLS * grp_u;
    const uint32_t w_s = __shfl_sync(0xffffffffu, smem_u32(base), 0);
    const uint32_t mbar_u = __shfl_sync(0xffffffffu, mbar, 0);
    const uint32_t tw = tg + ((uint32_t)(hw * 32) << 16);
    const uint64_t dW1Ahi = make_desc_sw128(w_s + NS_W1A), dW1Alo = make_desc_sw128(w_s + NS_W1A + 4096);
    const uint64_t dW1Bhi = make_desc_sw128(w_s + NS_W1B), dW1Blo = make_desc_sw128(w_s + NS_W1B + 4096);
    const uint64_t dW2hi = make_desc_sw128(w_s + NS_W2), dW2lo = make_desc_sw128(w_s + NS_W2 + 4096);
    const uint64_t dW3hi = make_desc_sw128(w_s + NS_W3), dW3lo = make_desc_sw128(w_s + NS_W3 + 8192);
    uint32_t phase = 0;
    const int bar_id = 1 + grp;

    const int64_t tiles = (a.G + 127) / 128;
    const int64_t tstep = (int64_t)gridDim.x * NT_GROUPS;
    // the next tile's input rows travel one tile ahead (cp.async into the warp's staging tiles)
    {
        const int64_t t0 = (int64_t)blockIdx.x * NT_GROUPS + grp;
        if (t0 < tiles) {
            nt_rows_async(stA_s, a.h_in, t0 * 128 + hw * 32, a.G, lane);
            if (do_s1) nt_rows_async(stB_s, a.agg, t0 * 128 + hw * 32, a.G, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int64_t tile = (int64_t)blockIdx.x * NT_GROUPS + grp; tile < tiles; tile += tstep) {
        const int64_t g = tile * 128 + ht;
        const int64_t gw0 = tile * 128 + hw * 32;          // first row of this warp
        const bool live = g < a.G;
        float h[32], v[32];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        nt_row_from_stage(h, stA, lane);
        if (do_s1) nt_row_from_stage(v, stB, lane);
        __syncwarp();
        if (tile + tstep < tiles) {              // next tile's rows: in flight during this tile's three stages
            nt_rows_async(stA_s, a.h_in, (tile + tstep) * 128 + hw * 32, a.G, lane);
            if (do_s1) nt_rows_async(stB_s, a.agg, (tile + tstep) * 128 + hw * 32, a.G, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        if (a.x3 && live) {
            *reinterpret_cast<float4 *>(a.x4_out + g * 4) =
                make_float4(__ldg(a.x3 + g * 3), __ldg(a.x3 + g * 3 + 1), __ldg(a.x3 + g * 3 + 2), 0.f);
        }
        if (do_s1) {
            // stage 1: A = [h | agg]: A_hi columns 32..95, A_lo columns 96..159, D columns 0..31
            nt_store_hilo(tw + 32, tw + 96, h);
            nt_store_hilo(tw + 64, tw + 128, v);
            tmem_wait_st();
            fence_before_sync();
            bar_sync(bar_id, 128);
            if (hw_u == 0 && elect_one()) {
                fence_after_sync();
                nt_issue_k32(tg, tg + 32, tg + 96, dW1Ahi, dW1Alo, IDESC_TF32_M128_N32, true);
                nt_issue_k32(tg, tg + 64, tg + 128, dW1Bhi, dW1Blo, IDESC_TF32_M128_N32, false);
                umma_commit(mbar_u);
            }
            mbar_wait(mbar, phase); phase ^= 1;
            fence_after_sync();
            tmem_ld32(tw, v);
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = silu(v[i] + par[i]);                      // :213-215
        } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = h[i];
        }
        if (do_s2) {
            // stage 2: A_hi columns 32..63, A_lo columns 96..127, D columns 0..31
            nt_store_hilo(tw + 32, tw + 96, v);
            tmem_wait_st();
            fence_before_sync();
            bar_sync(bar_id, 128);
            if (hw_u == 1 && elect_one()) {
                fence_after_sync();
                nt_issue_k32(tg, tg + 32, tg + 96, dW2hi, dW2lo, IDESC_TF32_M128_N32, true);
                umma_commit(mbar_u);
            }
            mbar_wait(mbar, phase); phase ^= 1;
            fence_after_sync();
            tmem_ld32(tw, v);
            if (a.residual) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = h[i] + (v[i] + par[32 + i]);          // :256-259
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += par[32 + i];
            }
        }
        if (!a.out_to_h) nt_rows_store(v, stC, a.h_out, gw0, a.G, lane);
        if (n3) {
            // stage 3: A_hi columns 64..95, A_lo columns 96..127, D columns 0..n3-1
            nt_store_hilo(tw + 64, tw + 96, v);
            tmem_wait_st();
            fence_before_sync();
            bar_sync(bar_id, 128);
            if (hw_u == 2 && elect_one()) {
                fence_after_sync();
                nt_issue_k32(tg, tg + 64, tg + 96, dW3hi, dW3lo, n3 == 64 ? IDESC_TF32_M128_N64 : IDESC_TF32_M128_N32, true);
                umma_commit(mbar_u);
            }
            mbar_wait(mbar, phase); phase ^= 1;
            fence_after_sync();
            tmem_ld32(tw, v);
            if (n3 == 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += par[64 + i];
            }
            nt_rows_store(v, stC, a.out_to_h ? a.h_out : a.P_out, gw0, a.G, lane);
            if (n3 == 64) {
                tmem_ld32(tw + 32, v);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += par[96 + i];
                nt_rows_store(v, stC, a.Q_out, gw0, a.G, lane);
            }
        }
        fence_before_sync();     // this tile's tcgen05.ld are ordered before the next tile's MMAs (after its barrier)
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tmem_holder, 512);
}

static int launch_node_ts(const NodeTsArgs &a, cudaStream_t st) {
    if (!opt_in_smem(egnn_node_ts_kernel, NT_SMEM_BYTES)) return EGSPR_E_LAUNCH;
    const int64_t tiles = (a.G + 127) / 128;
    int64_t grid = (tiles + NT_GROUPS - 1) / NT_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    const cudaError_t le = launch_pdl(egnn_node_ts_kernel, dim3((unsigned)grid), dim3(NT_THREADS), NT_SMEM_BYTES, st, a);
    return (le == cudaSuccess && cudaGetLastError() == cudaSuccess) ? EGSPR_OK : EGSPR_E_LAUNCH;
}

// ================================================================================================================
// node_model BACKWARD on tcgen05 (training step; src/3dmatch_train_egnn_with_batch.py:252-260 under loss.backward()):
//   recompute   z1 = [h | agg] Wn1^T + bn1,  a = SiLU(z1)                      M128 N32 K64   (as the forward stage 1)
//   da  = dout Wn2,   dz1 = da * SiLU'(z1)                                      M128 N32 K32   (B rows = Wn2 columns)
//   [d h | dagg] = dz1 Wn1,   dh_in = dout + d h                                M128 N64 K32
// with the A operands handed over through tensor memory (3xTF32) exactly like the forward node kernel; the three weight
// gradients  dWn2 += a^T dout,  dWn1 += [h | agg]^T dz1  (outer products over the 128 rows of a tile) stay on the CUDA
// cores but run while the stage-2 / stage-3 MMAs are in flight, out of the rows every warp keeps in its staging tiles.
// Thread = node = TMEM lane, a group of 128 threads per tile, two groups per CTA (four staged row sets per warp).
// Replaces the thread-per-node CUDA-core kernel (6144 FMAs per node from shared-memory weights): 100 -> see DESIGN 4.2.
// ================================================================================================================
constexpr int NB_GROUPS = 2;
constexpr int NB_THREADS = 128 * NB_GROUPS;
constexpr int NBS_W1A = 0;            // Wn1[:, 0:32]  hi, lo (rows = outputs)           stage 1
constexpr int NBS_W1B = 8192;         // Wn1[:, 32:64] hi, lo
constexpr int NBS_W2R = 16384;        // rows = INPUTS of node_mlp.2 (Wn2 columns) hi, lo  stage 2
constexpr int NBS_W3R = 24576;        // rows = the 64 inputs of node_mlp.0 (Wn1 columns): hi 8 KB, lo 8 KB   stage 3
constexpr int NBS_PAR = 40960;        // bn1[32]
constexpr int NBS_MBAR = NBS_PAR + 128;
constexpr int NBS_TMEM = NBS_MBAR + 8 * NB_GROUPS;
constexpr int NBS_STAGE = ((NBS_TMEM + 16 + 127) / 128) * 128;     // per warp: 4 x [32][36] floats: h | agg | dout | a -> dz1
constexpr int NB_TILE_F = 32 * NT_SROW;
constexpr int NB_WSTAGE = 4 * NB_TILE_F * 4;
constexpr int NBS_END = NBS_STAGE + (NB_THREADS / 32) * NB_WSTAGE;
constexpr size_t NB_TS_SMEM = NBS_END + 1024;

__global__ void __launch_bounds__(NB_THREADS, 1) node_mlp_backward_ts_kernel(const float *__restrict__ h_g, const float *__restrict__ agg_g,
                                                                             const float *__restrict__ dout_g, int64_t G,
                                                                             const int32_t *__restrict__ csr_ptr, const float *__restrict__ pack,
                                                                             float *__restrict__ dh_in, float *__restrict__ dagg_g,
                                                                             float *__restrict__ gpack) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, grp = tid >> 7, ht = tid & 127, hw = ht >> 5, lane = tid & 31;
    float *par = reinterpret_cast<float *>(base + NBS_PAR);
    const uint32_t mbar = smem_u32(base + NBS_MBAR + 8 * grp);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + NBS_TMEM);
    float *gstage = reinterpret_cast<float *>(base + NBS_STAGE + grp * 4 * NB_WSTAGE);      // this group's four warps
    float *stH = gstage + hw * (NB_WSTAGE / 4), *stG = stH + NB_TILE_F, *stD = stG + NB_TILE_F, *stX = stD + NB_TILE_F;
    const uint32_t stH_s = smem_u32(stH), stG_s = smem_u32(stG), stD_s = smem_u32(stD);
    // row e (0..127) of the group's tile, set t (0 h, 1 agg, 2 dout, 3 a / dz1)
    auto row_of = [&](int t, int e) -> const float * { return gstage + (e >> 5) * (NB_WSTAGE / 4) + t * NB_TILE_F + (e & 31) * NT_SROW; };

    // ---- weight tiles (hi / lo, swizzled).  pack: WN1T [64 in][32 out], WN2T [32 in][32 out] ----
    {   // 20 elements per thread, ALL loads issued before the first store (a load / store loop pays one L2 round trip
        // per iteration, and this kernel has no predecessor whose tail would hide its prologue)
        float wv[20];
#pragma unroll
        for (int j = 0; j < 20; ++j) {
            const int i = (tid + j * NB_THREADS) & 1023, r = i >> 5, k = i & 31;
            const float *src;
            if (j < 4) src = pack + OFF_WN1T + 32 * k + r;                                 // W1A  row o = r, K i = k: WN1T[i][o]
            else if (j < 8) src = pack + OFF_WN1T + 1024 + 32 * k + r;                     // W1B
            else if (j < 12) src = pack + OFF_WN2T + 32 * r + k;                           // W2R  row i = r, K o = k: WN2T[i][o]
            else src = pack + OFF_WN1T + 32 * (r + (j >= 16 ? 32 : 0)) + k;               // W3R  row j, K o: WN1T[j][o], 64 rows
            wv[j] = __ldg(src);
        }
#pragma unroll
        for (int j = 0; j < 20; ++j) {
            const int i = (tid + j * NB_THREADS) & 1023, r = i >> 5, k = i & 31;
            uint8_t *hi_t, *lo_t;
            int row = r;
            if (j < 4) { hi_t = base + NBS_W1A; lo_t = hi_t + 4096; }
            else if (j < 8) { hi_t = base + NBS_W1B; lo_t = hi_t + 4096; }
            else if (j < 12) { hi_t = base + NBS_W2R; lo_t = hi_t + 4096; }
            else { hi_t = base + NBS_W3R; lo_t = hi_t + 8192; row = r + (j >= 16 ? 32 : 0); }
            const float hi = tf32_hi(wv[j]);
            *reinterpret_cast<float *>(hi_t + sw128_off(row, k)) = hi;
            *reinterpret_cast<float *>(lo_t + sw128_off(row, k)) = wv[j] - hi;
        }
    }
    if (tid < 32) par[tid] = __ldg(pack + OFF_BN1 + tid);
    if (tid < 32) tmem_alloc(smem_u32(tmem_holder), 512);
    if (ht == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int hw_u = __shfl_sync(0xffffffffu, hw, 0), grp_u = __shfl_sync(0xffffffffu, grp, 0);
    const uint32_t tg = __shfl_sync(0xffffffffu, *tmem_holder, 0) + NT_COLS * grp_u;
    const uint32_t w_s = __shfl_sync(0xffffffffu, smem_u32(base), 0);
    const uint32_t mbar_u = __shfl_sync(0xffffffffu, mbar, 0);
    const uint32_t tw = tg + ((uint32_t)(hw * 32) << 16);
    const uint64_t dW1Ahi = make_desc_sw128(w_s + NBS_W1A), dW1Alo = make_desc_sw128(w_s + NBS_W1A + 4096);
    const uint64_t dW1Bhi = make_desc_sw128(w_s + NBS_W1B), dW1Blo = make_desc_sw128(w_s + NBS_W1B + 4096);
    const uint64_t dW2hi = make_desc_sw128(w_s + NBS_W2R), dW2lo = make_desc_sw128(w_s + NBS_W2R + 4096);
    const uint64_t dW3hi = make_desc_sw128(w_s + NBS_W3R), dW3lo = make_desc_sw128(w_s + NBS_W3R + 8192);
    uint32_t phase = 0;
    const int bar_id = 1 + grp;

    // weight-gradient entries of this thread: in-index i = lane, out-indices ob .. ob + 7 (ob = 8 * warp of the group)
    const int wi = lane, ob = hw * 8;
    float accW2[8], accW1h[8], accW1a[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { accW2[q] = 0.f; accW1h[q] = 0.f; accW1a[q] = 0.f; }
    float colDout = 0.f, colDz = 0.f, colDagg = 0.f;
    // acc[q] += sum over the tile's rows of In[e][wi] * Out[e][ob + q]; rows of warp w at constant offsets from its tile
    auto outer = [&](float (&acc)[8], int tin, int tout) {
        const float *pin = gstage + tin * NB_TILE_F + wi, *pout = gstage + tout * NB_TILE_F + ob;
#pragma unroll 1
        for (int w = 0; w < 4; ++w, pin += NB_WSTAGE / 4, pout += NB_WSTAGE / 4) {
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float x = pin[r * NT_SROW];
                const float4 o0 = *reinterpret_cast<const float4 *>(pout + r * NT_SROW), o1 = *reinterpret_cast<const float4 *>(pout + r * NT_SROW + 4);
                ffma2(acc[0], acc[1], x, x, o0.x, o0.y); ffma2(acc[2], acc[3], x, x, o0.z, o0.w);
                ffma2(acc[4], acc[5], x, x, o1.x, o1.y); ffma2(acc[6], acc[7], x, x, o1.z, o1.w);
            }
        }
    };

    const int64_t tiles = (G + 127) / 128;
    const int64_t tstep = (int64_t)gridDim.x * NB_GROUPS;
    for (int64_t tile = (int64_t)blockIdx.x * NB_GROUPS + grp; tile < tiles; tile += tstep) {
        const int64_t g = tile * 128 + ht, gw0 = tile * 128 + hw * 32;
        const bool live = g < G;
        // rows of this warp: coalesced cp.async into the warp's staging tiles (first tile here, later ones at the end of
        // the previous tile, as soon as its staging tiles are free)
        if (tile == (int64_t)blockIdx.x * NB_GROUPS + grp) {
            nt_rows_async(stH_s, h_g, gw0, G, lane);
            nt_rows_async(stG_s, agg_g, gw0, G, lane);
            nt_rows_async(stD_s, dout_g, gw0, G, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
        const float deg = live ? (float)(__ldg(csr_ptr + g + 1) - __ldg(csr_ptr + g)) : 0.f;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        float v[32], sp[32];
        // ---- stage 1: z1 = [h | agg] Wn1^T ----
        nt_row_from_stage(v, stH, lane);
        nt_store_hilo(tw + 32, tw + 96, v);
        nt_row_from_stage(v, stG, lane);
        nt_store_hilo(tw + 64, tw + 128, v);
        tmem_wait_st();
        fence_before_sync();
        bar_sync(bar_id, 128);
        if (hw_u == 0 && elect_one()) {
            fence_after_sync();
            nt_issue_k32(tg, tg + 32, tg + 96, dW1Ahi, dW1Alo, IDESC_TF32_M128_N32, true);
            nt_issue_k32(tg, tg + 64, tg + 128, dW1Bhi, dW1Blo, IDESC_TF32_M128_N32, false);
            umma_commit(mbar_u);
        }
        // dout row: zeroed for rows past the end (their staged copy repeats the last node), column sums for d bn2
        nt_row_from_stage(v, stD, lane);
        if (!live) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) *reinterpret_cast<float4 *>(stD + lane * NT_SROW + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        colDout += warp_colsum32(v);
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        {
            float z[32];
            tmem_ld32(tw, z);
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                const float zz = z[i] + par[i];
                float e, r;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(zz * -1.4426950408889634f));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
                z[i] = live ? zz * r : 0.f;                                   // a = SiLU(z1)
                sp[i] = live ? r * (1.0f + zz * (1.0f - r)) : 0.f;            // SiLU'(z1)
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
                *reinterpret_cast<float4 *>(stX + lane * NT_SROW + 4 * i) = make_float4(z[4 * i], z[4 * i + 1], z[4 * i + 2], z[4 * i + 3]);
        }
        // ---- stage 2: da = dout Wn2 ----
        nt_store_hilo(tw + 32, tw + 96, v);
        tmem_wait_st();
        fence_before_sync();
        bar_sync(bar_id, 128);                 // also: every a row of the group's tile is in shared memory
        if (hw_u == 1 && elect_one()) {
            fence_after_sync();
            nt_issue_k32(tg, tg + 32, tg + 96, dW2hi, dW2lo, IDESC_TF32_M128_N32, true);
            umma_commit(mbar_u);
        }
        outer(accW2, 3, 2);                    // dWn2T[i][o] += a[i] dout[o], while the MMAs run
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        tmem_ld32(tw, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= sp[i];                                             // dz1
        colDz += warp_colsum32(v);
        bar_sync(bar_id, 128);                 // every thread of the group is done reading the a rows
#pragma unroll
        for (int i = 0; i < 8; ++i)
            *reinterpret_cast<float4 *>(stX + lane * NT_SROW + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        // ---- stage 3: [d h | dagg] = dz1 Wn1 ----
        nt_store_hilo(tw + 64, tw + 96, v);
        tmem_wait_st();
        fence_before_sync();
        bar_sync(bar_id, 128);                 // also: every dz1 row is in shared memory
        if (hw_u == 2 && elect_one()) {
            fence_after_sync();
            nt_issue_k32(tg, tg + 64, tg + 96, dW3hi, dW3lo, IDESC_TF32_M128_N64, true);
            umma_commit(mbar_u);
        }
        outer(accW1h, 0, 3);                   // dWn1T[i][o]      += h[i]   dz1[o]
        outer(accW1a, 1, 3);                   // dWn1T[32 + i][o] += agg[i] dz1[o]
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        tmem_ld32(tw, v);                      // d h
        {
            float d[32];
            nt_row_from_stage(d, stD, lane);   // dout (own row)
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += d[i];
        }
        bar_sync(bar_id, 128);                 // the group is done with the outer products: the staging tiles may be reused
        if (tile + tstep < tiles) {            // the next tile's rows travel while this tile's results are stored
            const int64_t gn = (tile + tstep) * 128 + hw * 32;
            nt_rows_async(stH_s, h_g, gn, G, lane);
            nt_rows_async(stG_s, agg_g, gn, G, lane);
            nt_rows_async(stD_s, dout_g, gn, G, lane);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        nt_rows_store(v, stX, dh_in, gw0, G, lane);
        tmem_ld32(tw + 32, v);                 // dagg
        nt_rows_store(v, stX, dagg_g, gw0, G, lane);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= deg;          // d LayerNorm beta, aggregate part: sum_n deg(n) dagg[n]
        colDagg += warp_colsum32(v);
        fence_before_sync();     // this tile's tcgen05.ld are ordered before the next tile's MMAs (after its barrier)
    }
    // weight gradients: the groups and warps of the CTA meet in shared memory (the staging tiles are free now), then ONE
    // atomicAdd per entry and CTA, every CTA starting at a different entry -- all CTAs finish together, and 296 groups
    // x 27 atomics per thread onto the same 3168 addresses in the same order were 30 % of the kernel's stall samples
    __syncthreads();
    {
        float *red = reinterpret_cast<float *>(base + NBS_STAGE);            // [NB_GROUPS][3072] + [warps][96]
        float *rg = red + grp * 3072;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            rg[32 * wi + ob + q] = accW2[q];                                   // dWn2T [32][32]
            rg[1024 + 32 * wi + ob + q] = accW1h[q];                           // dWn1T [64][32], rows 0..31
            rg[2048 + 32 * wi + ob + q] = accW1a[q];                           //                 rows 32..63
        }
        float *rc = red + NB_GROUPS * 3072 + (tid >> 5) * 96;
        rc[lane] = colDout; rc[32 + lane] = colDz; rc[64 + lane] = colDagg;
        __syncthreads();
        const int rot = (int)((blockIdx.x * 211u) % 3072u);
        for (int i0 = tid; i0 < 3072; i0 += NB_THREADS) {
            const int i = (i0 + rot) % 3072;
            float t = 0.f;
#pragma unroll
            for (int gI = 0; gI < NB_GROUPS; ++gI) t += red[gI * 3072 + i];
            atomicAdd(gpack + (i < 1024 ? OFF_WN2T + i : OFF_WN1T + (i - 1024)), t);
        }
        if (tid < 96) {
            float t = 0.f;
#pragma unroll
            for (int w = 0; w < NB_THREADS / 32; ++w) t += red[NB_GROUPS * 3072 + w * 96 + tid];
            atomicAdd(gpack + (tid < 32 ? OFF_BN2 + tid : (tid < 64 ? OFF_BN1 + (tid - 32) : OFF_LNB + (tid - 64))), t);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (tid < 32) tmem_dealloc(*tmem_holder, 512);
}

int launch_node_mlp_backward_ts(const float *h, const float *agg, const float *dh_out, int64_t G, const int32_t *csr_ptr,
                                const float *pack, float *dh_in, float *dagg, float *gpack, cudaStream_t st) {
    if (!opt_in_smem(node_mlp_backward_ts_kernel, NB_TS_SMEM)) return EGSPR_E_LAUNCH;
    const int64_t tiles = (G + 127) / 128;
    int64_t grid = (tiles + NB_GROUPS - 1) / NB_GROUPS;
    if (grid > sm_count()) grid = sm_count();
    node_mlp_backward_ts_kernel<<<(unsigned)grid, NB_THREADS, NB_TS_SMEM, st>>>(h, agg, dh_out, G, csr_ptr, pack, dh_in, dagg, gpack);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

// node MLP + residual + next P/Q (or embedding_out) after the edge kernel of a layer
int launch_node_update_ts(const LayerArgs &a, const float *agg, cudaStream_t st) {
    NodeTsArgs n{};
    n.h_in = a.h; n.agg = agg; n.x3 = nullptr; n.x4_out = nullptr;
    n.w1t = a.layer_pack + OFF_WN1T; n.b1 = a.layer_pack + OFF_BN1;
    n.w2t = a.layer_pack + OFF_WN2T; n.b2 = a.layer_pack + OFF_BN2;
    n.h_out = a.h_out; n.P_out = a.P_out; n.Q_out = a.Q_out; n.G = a.num_nodes; n.residual = 1; n.out_to_h = 0;
    if (a.next_pack) {
        n.w3pt = a.next_pack + OFF_WPT; n.w3qt = a.next_pack + OFF_WQT; n.b3 = a.next_pack + OFF_BQ;
    } else if (a.out_pack) {
        n.w3pt = a.out_pack; n.w3qt = nullptr; n.b3 = a.out_pack + 1024; n.out_to_h = 1;
    }
    return launch_node_ts(n, st);
}

// embedding_in (optional) + layer-0 P/Q + x3 -> x4
int launch_node_embed_ts(const float *feat, const float *x3, int64_t G, const float *embed_pack, const float *layer0_pack,
                         float *h, float *x4, float *P, float *Q, cudaStream_t st) {
    NodeTsArgs n{};
    n.h_in = feat; n.agg = nullptr; n.x3 = x4 ? x3 : nullptr; n.x4_out = x4;
    n.w2t = embed_pack; n.b2 = embed_pack ? embed_pack + 1024 : nullptr;
    n.w3pt = layer0_pack + OFF_WPT; n.w3qt = layer0_pack + OFF_WQT; n.b3 = layer0_pack + OFF_BQ;
    n.h_out = h; n.P_out = P; n.Q_out = Q; n.G = G; n.residual = 0; n.out_to_h = 0;
    return launch_node_ts(n, st);
}

}  // namespace egspr
