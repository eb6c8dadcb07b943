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
