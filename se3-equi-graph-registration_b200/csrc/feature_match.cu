// Feature-space nearest neighbour of every descriptor of A among the descriptors of B -- the correspondence
// construction of data_preprocess/3DMatch_Feature.py:158-166:
//     distance   = np.sqrt(2 - 2 * (src_desc @ tgt_desc.T) + 1e-6)
//     source_idx = np.argmin(distance, axis=1);  source_dis = np.min(distance, axis=1)
//     target_idx = np.argmin(distance, axis=0)   (mutual check: target_idx[source_idx] == arange)
// The [Na, Nb] similarity matrix is a dense GEMM with K = 32: it runs on tcgen05 (3xTF32 split, fp32
// accumulation in TMEM) in 128 x 128 tiles and is never written to memory -- the epilogue turns each
// accumulator row into distances exactly as the reference does (fp32: 2*s, 2 - ., + 1e-6, IEEE sqrt) and
// keeps the running (distance, index) minimum, first index on ties like np.argmin.  The axis-0 argmin is the
// same kernel with A and B swapped.
//
// CTA = 128 threads = 128 rows of A (thread = row = TMEM lane; A_hi | A_lo written once with tcgen05.st);
// blockIdx.y splits the B range; B tiles (128 rows x 32) are produced coalesced by the threads as hi / lo
// 128B-swizzled K-major shared-memory tiles; 12 MMAs (M128 N128 K8) per tile; results of the splits meet in
// a packed (distance bits << 32 | index) atomicMin.
#include "egspr_common.cuh"
#include "tcgen05.cuh"

namespace egspr {
using namespace tc;

constexpr int FM_THREADS = 128;
constexpr int FM_BN = 128;                   // rows of B per tile == MMA N
constexpr int FS_BHI = 0, FS_BLO = 16384, FS_MBAR = 32768, FS_TMEM = 32776, FS_END = 32784;
constexpr size_t FM_SMEM_BYTES = FS_END + 1024;
constexpr uint32_t IDESC_TF32_M128_N128 = (1u << 4) | (2u << 7) | (2u << 10) | ((128u >> 3) << 17) | ((128u >> 4) << 24);

__global__ void __launch_bounds__(FM_THREADS) feature_nn_kernel(const float *__restrict__ A, int na, const float *__restrict__ B, int nb,
                                                               unsigned long long *__restrict__ best_packed) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t mbar = smem_u32(base + FS_MBAR);
    uint32_t *tmem_holder = reinterpret_cast<uint32_t *>(base + FS_TMEM);
    if (warp == 0) tmem_alloc(smem_u32(tmem_holder), 256);
    if (tid == 0) { mbar_init(mbar, 1); fence_mbar_init(); }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const int warp_u = __shfl_sync(0xffffffffu, warp, 0);
    const uint32_t tm = __shfl_sync(0xffffffffu, *tmem_holder, 0);       // D: columns 0..127, A_hi 128..159, A_lo 160..191
    const uint32_t tw = tm + ((uint32_t)(warp * 32) << 16);
    const uint32_t b_s = __shfl_sync(0xffffffffu, smem_u32(base), 0);
    const uint64_t dBhi = make_desc_sw128(b_s + FS_BHI), dBlo = make_desc_sw128(b_s + FS_BLO);

    // ---- this thread's row of A -> TMEM (hi | lo) ----
    const int row = blockIdx.x * 128 + tid;
    {
        float v[32];
        const float *ar = A + (size_t)min(row, na - 1) * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float4 t = ldg4(ar + 4 * i); v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w; }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            float hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) { hi[i] = tf32_hi(v[16 * b + i]); lo[i] = v[16 * b + i] - hi[i]; }
            tmem_st16(tw + 128 + 16 * b, hi);
            tmem_st16(tw + 160 + 16 * b, lo);
        }
        tmem_wait_st();
    }
    float best_d = 3.0e38f, best_t = 3.0e38f;      // best distance and its radicand (sqrt is evaluated only on improvements)
    int best_j = 0x7fffffff;
    uint32_t phase = 0;
    // B range of this split
    const int tiles = (nb + FM_BN - 1) / FM_BN;
    const int t_lo = (int)((int64_t)tiles * blockIdx.y / gridDim.y), t_hi = (int)((int64_t)tiles * (blockIdx.y + 1) / gridDim.y);
    for (int t = t_lo; t < t_hi; ++t) {
        const int n0 = t * FM_BN;
        // ---- B tile -> hi / lo swizzled K-major tiles, coalesced (8 lanes per 128-byte row) ----
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int r = 16 * i + (tid >> 3), ch = tid & 7;
            float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n0 + r < nb) w = ldg4(B + (size_t)(n0 + r) * 32 + 4 * ch);
            float4 hi, lo;
            hi.x = tf32_hi(w.x); hi.y = tf32_hi(w.y); hi.z = tf32_hi(w.z); hi.w = tf32_hi(w.w);
            lo.x = w.x - hi.x; lo.y = w.y - hi.y; lo.z = w.z - hi.z; lo.w = w.w - hi.w;
            const int off = r * 128 + ((ch ^ (r & 7)) << 4);
            *reinterpret_cast<float4 *>(base + FS_BHI + off) = hi;
            *reinterpret_cast<float4 *>(base + FS_BLO + off) = lo;
        }
        fence_proxy_async();            // generic-proxy writes -> visible to the tensor core
        fence_before_sync();            // (also orders the previous tile's tcgen05.ld before the next MMAs)
        __syncthreads();
        if (warp_u == 0 && elect_one()) {
            fence_after_sync();
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32_ts(tm, tm + 128 + 8 * k, dBhi + 2 * k, IDESC_TF32_M128_N128, k > 0);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32_ts(tm, tm + 160 + 8 * k, dBhi + 2 * k, IDESC_TF32_M128_N128, 1);
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_tf32_ts(tm, tm + 128 + 8 * k, dBlo + 2 * k, IDESC_TF32_M128_N128, 1);
            umma_commit(mbar);
        }
        mbar_wait(mbar, phase); phase ^= 1;
        fence_after_sync();
        // ---- epilogue: distances of this row against the tile's 128 columns, running argmin ----
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float s[32];
            tmem_ld32(tw + 32 * q, s);
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const int j = n0 + 32 * q + c;
                // np.sqrt(2 - 2 * s + 1e-6) in float32, operation by operation (3DMatch_Feature.py:158).  sqrt is
                // monotone: a radicand that is not smaller cannot give a smaller distance, so the IEEE sqrt runs
                // only on improvements; equal distances from different radicands keep the FIRST index (np.argmin)
                const float tr = __fadd_rn(__fsub_rn(2.0f, __fmul_rn(2.0f, s[c])), 1e-6f);
                if (j < nb && tr < best_t) {
                    const float d = __fsqrt_rn(tr);
                    best_t = tr;
                    if (d < best_d) { best_d = d; best_j = j; }
                }
            }
        }
        // the next iteration's __syncthreads orders these tcgen05.ld before the next tile's MMAs and B writes
    }
    if (row < na && best_j != 0x7fffffff) {
        // NaN distances (2 - 2s + 1e-6 < 0 cannot happen for unit descriptors) never win; d >= 0 -> bits are ordered
        const unsigned long long packed = ((unsigned long long)__float_as_uint(best_d) << 32) | (unsigned)best_j;
        atomicMin(best_packed + row, packed);
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 256);
}

__global__ void feature_nn_unpack_kernel(const unsigned long long *__restrict__ packed, int n, int32_t *__restrict__ idx,
                                         float *__restrict__ dist) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const unsigned long long p = packed[i];
        idx[i] = (int32_t)(p & 0xffffffffull);
        dist[i] = __uint_as_float((unsigned)(p >> 32));
    }
}

}  // namespace egspr

extern "C" int egspr_feature_nn(const float *a, int na, const float *b, int nb, void *workspace, size_t workspace_bytes,
                                int32_t *idx, float *dist, void *stream) {
    using namespace egspr;
    if (!a || !b || !workspace || !idx || !dist || na <= 0 || nb <= 0) return EGSPR_E_INVALID;
    if (workspace_bytes < sizeof(unsigned long long) * (size_t)na) return EGSPR_E_WORKSPACE;
    cudaStream_t st = (cudaStream_t)stream;
    if (!opt_in_smem(feature_nn_kernel, FM_SMEM_BYTES)) return EGSPR_E_LAUNCH;
    unsigned long long *packed = (unsigned long long *)workspace;
    if (cudaMemsetAsync(packed, 0xff, sizeof(unsigned long long) * (size_t)na, st) != cudaSuccess) return EGSPR_E_LAUNCH;
    const int mt = (na + 127) / 128, nt = (nb + FM_BN - 1) / FM_BN;
    int splits = (2 * sm_count() + mt - 1) / mt;          // ~2 CTAs per SM overall
    if (splits > nt) splits = nt;
    if (splits < 1) splits = 1;
    feature_nn_kernel<<<dim3(mt, splits), FM_THREADS, FM_SMEM_BYTES, st>>>(a, na, b, nb, packed);
    feature_nn_unpack_kernel<<<(na + 255) / 256, 256, 0, st>>>(packed, na, idx, dist);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}
