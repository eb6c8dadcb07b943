// Correspondence-weight head + weighted Kabsch / 3x3 SVD pose solve, one CTA per registration pair.
//
// Replaces the per-pair Python loops of CrossAttentionPoseRegression.forward:
//   eval variant  src/eval_egnn_metrics.py:691-818   (weights from the input-feature similarity,
//                 top-128 / mlp / scatter / renormalise / softmax chain, Kabsch on ORIGINAL coords)
//   train variant src/3dmatch_train_egnn_with_batch.py:696-758 (softmax of output-feature similarity
//                 over the GT inliers, Kabsch on the EGNN coords)
// and the cuSOLVER launch + `if det<0` host sync per pair (:741-751) with an in-kernel fp64 Jacobi SVD.
#include "egspr_common.cuh"

namespace egspr {

constexpr int HD_THREADS = 256;       // CTA size for clouds up to HD_BIG_N points
constexpr int HD_THREADS_BIG = 1024;  // ... and beyond (one CTA still owns a pair; every loop strides by blockDim.x)
constexpr int HD_BIG_N = 8192;
constexpr int HD_MAX_N = 48 * 1024;   // n floats of dynamic shared memory (<= 192 KB); larger clouds stage in w_out
constexpr int HD_WARPS = HD_THREADS_BIG / 32;     // scratch is sized for the largest CTA

struct BlockScratch {
    float red[HD_WARPS][16];
    float bc[16];
    unsigned hist[256];
    unsigned long long red64[HD_WARPS];
    unsigned u[8];
};

template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], BlockScratch &sc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) sc.red[warp][i] = v[i];
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sc.red[w][threadIdx.x];
        sc.bc[threadIdx.x] = s;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = sc.bc[i];
}

__device__ __forceinline__ float block_max(float v, BlockScratch &sc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) sc.red[warp][0] = v;
    __syncthreads();
    float m = sc.red[0][0];
#pragma unroll
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, sc.red[w][0]);
    return m;
}

// ---- one sweep of one-sided Jacobi (Hestenes) on the columns of A, accumulating the rotations in V; returns whether any pair
// was rotated.  t = tan(theta) from the column norms alpha, beta and the inner product gamma without forming zeta:
//   t = sgn(zeta) / (|zeta| + sqrt(1 + zeta^2)),  zeta = (beta - alpha) / (2 gamma)   ==   2 gamma sgn(delta) / (|delta| + sqrt(delta^2 + 4 gamma^2))
// -- one sqrt, one division and one rsqrt per rotation (fp64 sqrt / div are ~100-cycle software sequences, and this loop
// runs on ONE thread of the CTA: it was most of the head kernels' serial tail).
template <typename T>
__device__ __forceinline__ bool jacobi_sweep(T (&A)[3][3], T (&V)[3][3], T tol2) {
    bool rotated = false;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
        const int p = (pq == 2) ? 1 : 0, q = (pq == 0) ? 1 : 2;
        T alpha = 0, beta = 0, gamma = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { alpha += A[i][p] * A[i][p]; beta += A[i][q] * A[i][q]; gamma += A[i][p] * A[i][q]; }
        if (gamma * gamma > tol2 * (alpha * beta) && gamma != T(0)) {        // |gamma| > tol sqrt(alpha beta)
            rotated = true;
            const T delta = beta - alpha;
            const T t = (T(2) * gamma * copysign(T(1), delta)) / (fabs(delta) + sqrt(delta * delta + T(4) * gamma * gamma));
            const T c = rsqrt(T(1) + t * t), s = c * t;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const T ap = A[i][p], aq = A[i][q];
                A[i][p] = c * ap - s * aq; A[i][q] = s * ap + c * aq;
                const T vp = V[i][p], vq = V[i][q];
                V[i][p] = c * vp - s * vq; V[i][q] = s * vp + c * vq;
            }
        }
    }
    return rotated;
}

// ---- 3x3 SVD by one-sided Jacobi in fp64, R = V diag(1,1,det) U^T  (3dm:741-751) ---------------
// Hm = U diag(sg) W^T with sg descending; d = -1 if det(W U^T) < 0 (the reference then flips the smallest-sigma
// row of Vt).  Returns false for Hm == 0 (any basis; R = I).
__device__ bool kabsch_svd(const double (&Hm)[3][3], double (&U)[3][3], double (&W)[3][3], double (&sg)[3], double &d) {
    double A[3][3], V[3][3];
    {   // fp32 sweeps first (MUFU-speed rotations) bring V to ~1e-7 of the answer; V is re-orthonormalised in fp64
        // (Gram-Schmidt) and the fp64 sweeps below then converge quadratically from there: 1-2 sweeps instead of 4-6.
        // Scaled by the largest entry so that squares of tiny H (~1e-6 I) stay normal fp32 numbers.
        double hmax = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) hmax = fmax(hmax, fabs(Hm[i][j]));
        float Af[3][3], Vf[3][3];
        const double sc = hmax > 0.0 ? 1.0 / hmax : 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) { Af[i][j] = (float)(Hm[i][j] * sc); Vf[i][j] = (i == j) ? 1.f : 0.f; }
        for (int sweep = 0; sweep < 6; ++sweep)
            if (!jacobi_sweep<float>(Af, Vf, 1e-12f)) break;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) V[i][j] = (double)Vf[i][j];
        // Gram-Schmidt on the columns of V (they are orthonormal to ~1e-7 already)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
            for (int l = 0; l < j; ++l) {
                const double dot = V[0][j] * V[0][l] + V[1][j] * V[1][l] + V[2][j] * V[2][l];
#pragma unroll
                for (int i = 0; i < 3; ++i) V[i][j] -= dot * V[i][l];
            }
            const double inv = rsqrt(V[0][j] * V[0][j] + V[1][j] * V[1][j] + V[2][j] * V[2][j]);
#pragma unroll
            for (int i = 0; i < 3; ++i) V[i][j] *= inv;
        }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) A[i][j] = Hm[i][0] * V[0][j] + Hm[i][1] * V[1][j] + Hm[i][2] * V[2][j];      // A = H V
    }
    for (int sweep = 0; sweep < 40; ++sweep)
        if (!jacobi_sweep<double>(A, V, 1e-34)) break;
    double sig[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) sig[j] = sqrt(A[0][j] * A[0][j] + A[1][j] * A[1][j] + A[2][j] * A[2][j]);
    // order columns by descending sigma (LAPACK order: the det fix flips the SMALLEST-sigma row of Vt)
    int o0 = 0, o1 = 1, o2 = 2;
    if (sig[o0] < sig[o1]) { int t = o0; o0 = o1; o1 = t; }
    if (sig[o1] < sig[o2]) { int t = o1; o1 = o2; o2 = t; }
    if (sig[o0] < sig[o1]) { int t = o0; o0 = o1; o1 = t; }
    const int ord[3] = {o0, o1, o2};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const int s = ord[j];
        const double inv = sig[s] > 1e-300 ? 1.0 / sig[s] : 0.0;
        sg[j] = sig[s];
#pragma unroll
        for (int i = 0; i < 3; ++i) { U[i][j] = A[i][s] * inv; W[i][j] = V[i][s]; }
    }
    const double smax = sig[o0];
    d = 1.0;
    if (!(smax > 0.0)) return false;   // H == 0: any basis; LAPACK returns identity factors
    if (sig[o1] <= 1e-14 * smax) {   // rank 1: complete U with any unit vector orthogonal to U0
        const double ax = fabs(U[0][0]), ay = fabs(U[1][0]), az = fabs(U[2][0]);
        double e[3] = {0, 0, 0};
        e[(ax <= ay && ax <= az) ? 0 : (ay <= az ? 1 : 2)] = 1.0;
        const double d = e[0] * U[0][0] + e[1] * U[1][0] + e[2] * U[2][0];
        double v[3] = {e[0] - d * U[0][0], e[1] - d * U[1][0], e[2] - d * U[2][0]};
        const double nv = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        U[0][1] = v[0] / nv; U[1][1] = v[1] / nv; U[2][1] = v[2] / nv;
    }
    if (sig[o2] <= 1e-14 * smax) {   // rank <= 2: third left vector = U0 x U1 (sign is fixed below)
        U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
        U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
        U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    }
    auto det3 = [](const double (&M)[3][3]) {
        return M[0][0] * (M[1][1] * M[2][2] - M[1][2] * M[2][1]) - M[0][1] * (M[1][0] * M[2][2] - M[1][2] * M[2][0]) +
               M[0][2] * (M[1][0] * M[2][1] - M[1][1] * M[2][0]);
    };
    d = det3(W) * det3(U) < 0.0 ? -1.0 : 1.0;   // det(V U^T) < 0 -> Vt[-1,:] *= -1
    return true;
}

__device__ void kabsch_solve(const double (&Hm)[3][3], double (&R)[3][3]) {
    double U[3][3], W[3][3], sg[3], d;
    if (!kabsch_svd(Hm, U, W, sg, d)) {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) R[i][j] = (i == j) ? 1.0 : 0.0;
        return;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) R[i][j] = W[i][0] * U[j][0] + W[i][1] * U[j][1] + d * W[i][2] * U[j][2];
}

// p,q: point rows with `stride` floats; w: weights in shared or global memory (0 for excluded points).
// Writes R[9], t[3], Hout[9] for this CTA's pair.  `count` = number of included points (0 -> I, 0).
__device__ void block_kabsch(const float *__restrict__ p, const float *__restrict__ q, int stride,
                             const float *w, int n, int count, float *__restrict__ R, float *__restrict__ t,
                             float *__restrict__ Hout, BlockScratch &sc) {
    float c[6] = {0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float wi = w[i];
        if (wi != 0.f) {
            c[0] = fmaf(wi, p[i * stride], c[0]); c[1] = fmaf(wi, p[i * stride + 1], c[1]); c[2] = fmaf(wi, p[i * stride + 2], c[2]);
            c[3] = fmaf(wi, q[i * stride], c[3]); c[4] = fmaf(wi, q[i * stride + 1], c[4]); c[5] = fmaf(wi, q[i * stride + 2], c[5]);
        }
    }
    block_sum<6>(c, sc);
    float hm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float wi = w[i];
        if (wi != 0.f) {
            const float a0 = wi * (p[i * stride] - c[0]), a1 = wi * (p[i * stride + 1] - c[1]), a2 = wi * (p[i * stride + 2] - c[2]);
            const float b0 = q[i * stride] - c[3], b1 = q[i * stride + 1] - c[4], b2 = q[i * stride + 2] - c[5];
            hm[0] = fmaf(a0, b0, hm[0]); hm[1] = fmaf(a0, b1, hm[1]); hm[2] = fmaf(a0, b2, hm[2]);
            hm[3] = fmaf(a1, b0, hm[3]); hm[4] = fmaf(a1, b1, hm[4]); hm[5] = fmaf(a1, b2, hm[5]);
            hm[6] = fmaf(a2, b0, hm[6]); hm[7] = fmaf(a2, b1, hm[7]); hm[8] = fmaf(a2, b2, hm[8]);
        }
    }
    block_sum<9>(hm, sc);
    if (threadIdx.x == 0) {
        if (count == 0) {
            for (int i = 0; i < 9; ++i) { R[i] = (i % 4 == 0) ? 1.f : 0.f; if (Hout) Hout[i] = 0.f; }
            t[0] = t[1] = t[2] = 0.f;
        } else {
            hm[0] += 1e-6f; hm[4] += 1e-6f; hm[8] += 1e-6f;                       // regulariser :738
            double Hd[3][3], Rd[3][3];
            for (int i = 0; i < 9; ++i) { Hd[i / 3][i % 3] = (double)hm[i]; if (Hout) Hout[i] = hm[i]; }
            kabsch_solve(Hd, Rd);
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) R[i * 3 + j] = (float)Rd[i][j];
                t[i] = (float)((double)c[3 + i] - (Rd[i][0] * c[0] + Rd[i][1] * c[1] + Rd[i][2] * c[2]));  // :754
            }
        }
    }
}

__device__ __forceinline__ float dot32(const float *__restrict__ a, const float *__restrict__ b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 u = ldg4(a + 4 * i), v = ldg4(b + 4 * i);
        s = fmaf(u.x, v.x, s); s = fmaf(u.y, v.y, s); s = fmaf(u.z, v.z, s); s = fmaf(u.w, v.w, s);
    }
    return s;
}

// egnn_equi_loss partial sums (3dm:860-893) for this pair
__device__ void block_equi_loss(const float *__restrict__ hs, const float *__restrict__ ht,
                                const float *__restrict__ xs, const float *__restrict__ xt,
                                const float *__restrict__ labels, const float *__restrict__ gt, int n,
                                float *__restrict__ loss_parts, BlockScratch &sc) {
    float g[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g[3 * i] = __ldg(gt + 4 * i); g[3 * i + 1] = __ldg(gt + 4 * i + 1); g[3 * i + 2] = __ldg(gt + 4 * i + 2);
        g[9 + i] = __ldg(gt + 4 * i + 3);
    }
    float acc[2] = {0.f, 0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float lab = __ldg(labels + i);
        const float x0 = xs[3 * i], x1 = xs[3 * i + 1], x2 = xs[3 * i + 2];
        float ch = 0.f;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float d = (g[3 * r] * x0 + g[3 * r + 1] * x1 + g[3 * r + 2] * x2 + g[9 + r]) - xt[3 * i + r];
            ch = fmaf(d, d, ch);
        }
        acc[0] = fmaf(ch, lab, acc[0]);
        const float ab = dot32(hs + (size_t)i * H, ht + (size_t)i * H);
        const float aa = dot32(hs + (size_t)i * H, hs + (size_t)i * H), bb = dot32(ht + (size_t)i * H, ht + (size_t)i * H);
        // F.cosine_similarity: x.y / (max(|x|,eps) * max(|y|,eps)), eps = 1e-8
        const float cs = ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
        acc[1] = fmaf(cs - lab, cs - lab, acc[1]);
    }
    block_sum<2>(acc, sc);
    if (threadIdx.x == 0) { loss_parts[0] = acc[0]; loss_parts[1] = acc[1]; }
}

__device__ __forceinline__ unsigned order_key(float f) {   // ascending float order == ascending uint order
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct HeadEvalArgs {
    const float *feat_src, *feat_tgt, *x_src, *x_tgt, *h_out_src, *h_out_tgt, *x_out_src, *x_out_tgt, *labels,
        *gt_pose, *head_pack;
    int n, top_k;
    float *w_out, *R, *t, *Hout, *loss_parts;
    // few pairs x large clouds: the per-point passes over the 32-wide rows (input-feature similarity, egnn_equi_loss) are
    // done by head_eval_pre_kernel on `split` CTAs per pair; this kernel then starts from their partial results
    int split;                                   // 0 = everything in this kernel
    const unsigned long long *pre_best;          // [pairs][split] packed (similarity key, index) maxima
    const float *pre_lp;                         // [pairs][split][2] egnn_equi_loss partial sums
};

// grid (split, pairs): chunk of the pair's points -> sim0 into the pair's row of w_out (scratch), chunk argmax, chunk loss sums
__global__ void __launch_bounds__(256) head_eval_pre_kernel(const HeadEvalArgs a, unsigned long long *__restrict__ best_out,
                                                            float *__restrict__ lp_out) {
    __shared__ BlockScratch sc;
    const int b = blockIdx.y, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t nb = (size_t)b * n;
    const int per = (n + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    float g[12];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        g[3 * i] = __ldg(a.gt_pose + b * 16 + 4 * i); g[3 * i + 1] = __ldg(a.gt_pose + b * 16 + 4 * i + 1);
        g[3 * i + 2] = __ldg(a.gt_pose + b * 16 + 4 * i + 2); g[9 + i] = __ldg(a.gt_pose + b * 16 + 4 * i + 3);
    }
    unsigned long long best = 0ull;
    float acc[2] = {0.f, 0.f};
    for (int i = i0 + tid; i < i1; i += blockDim.x) {
        const float s = dot32(a.feat_src + (nb + i) * H, a.feat_tgt + (nb + i) * H);
        a.w_out[nb + i] = s;
        const unsigned long long cand = ((unsigned long long)order_key(s) << 32) | (unsigned)(0xffffffffu - (unsigned)i);
        best = cand > best ? cand : best;
        if (a.loss_parts) {                       // as block_equi_loss (3dm:860-893)
            const float *hs = a.h_out_src + (nb + i) * H, *ht = a.h_out_tgt + (nb + i) * H;
            const float *xs = a.x_out_src + (nb + i) * 3, *xt = a.x_out_tgt + (nb + i) * 3;
            const float lab = __ldg(a.labels + nb + i);
            const float x0 = xs[0], x1 = xs[1], x2 = xs[2];
            float ch = 0.f;
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float d = (g[3 * r] * x0 + g[3 * r + 1] * x1 + g[3 * r + 2] * x2 + g[9 + r]) - xt[r];
                ch = fmaf(d, d, ch);
            }
            acc[0] = fmaf(ch, lab, acc[0]);
            const float ab = dot32(hs, ht), aa = dot32(hs, hs), bb = dot32(ht, ht);
            const float cs = ab / (fmaxf(sqrtf(aa), 1e-8f) * fmaxf(sqrtf(bb), 1e-8f));
            acc[1] = fmaf(cs - lab, cs - lab, acc[1]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if (lane == 0) sc.red64[warp] = best;
    block_sum<2>(acc, sc);               // (its barriers also publish red64)
    if (tid == 0) {
        unsigned long long m = sc.red64[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = sc.red64[w] > m ? sc.red64[w] : m;
        best_out[(size_t)b * gridDim.x + blockIdx.x] = m;
        lp_out[((size_t)b * gridDim.x + blockIdx.x) * 2] = acc[0];
        lp_out[((size_t)b * gridDim.x + blockIdx.x) * 2 + 1] = acc[1];
    }
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT) head_eval_kernel(const HeadEvalArgs a) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ BlockScratch sc;
    const int b = blockIdx.x, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t nb = (size_t)b * n;
    // [n] sim0, later the weights: shared memory, or (n > HD_MAX_N) this pair's row of w_out in global memory
    float *ssim = (n > HD_MAX_N) ? a.w_out + nb : dyn;
    // 1. input-feature similarity (evl:691)
    unsigned long long best = 0ull;
    if (a.split > 0) {                   // done by head_eval_pre_kernel: sim0 sits in this pair's row of w_out
        if (n <= HD_MAX_N)
            for (int i = tid; i < n; i += blockDim.x) ssim[i] = a.w_out[nb + i];
        for (int i = tid; i < a.split; i += blockDim.x) {
            const unsigned long long cand = a.pre_best[(size_t)b * a.split + i];
            best = cand > best ? cand : best;
        }
    } else
    for (int i = tid; i < n; i += blockDim.x) {
        const float s = dot32(a.feat_src + (nb + i) * H, a.feat_tgt + (nb + i) * H);
        ssim[i] = s;
        const unsigned long long cand = ((unsigned long long)order_key(s) << 32) | (unsigned)(0xffffffffu - (unsigned)i);
        best = cand > best ? cand : best;
    }
    // 2. argmax (top_indices[:,0]; ties -> lowest index)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
        best = other > best ? other : best;
    }
    if (lane == 0) sc.red64[warp] = best;
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = sc.red64[0];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = sc.red64[w] > m ? sc.red64[w] : m;
        sc.u[0] = 0xffffffffu - (unsigned)(m & 0xffffffffull);   // i0
    }
    // 3. k-th largest key by 4-pass radix select (evl:694 topk, k=128)
    const int kk = a.top_k < n ? a.top_k : n;
    unsigned prefix = 0, pmask = 0;
    int want = kk;                 // rank (1-based, from the top) still to locate inside the prefix class
    for (int shift = 24; shift >= 0; shift -= 8) {
        __syncthreads();
        if (tid < 256) sc.hist[tid] = 0;
        __syncthreads();
        // four independent loads in flight per thread (large clouds keep sim0 in global memory) and warp-aggregated
        // histogram updates (similarities crowd into a few bins: one atomic per distinct bin and warp, not per lane)
        for (int base = 0; base < n; base += 4 * blockDim.x) {        // block-uniform trip count: __match_any_sync needs whole warps
            const int i0 = base + tid;
            float v4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int i = i0 + q * blockDim.x; v4[q] = i < n ? ssim[i] : 0.f; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned key = order_key(v4[q]);
                const bool in = (i0 + q * (int)blockDim.x < n) && (key & pmask) == prefix;
                const unsigned bin = in ? ((key >> shift) & 0xffu) : 0x100u;
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (in && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sc.hist[bin], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, bin = 255;
            for (; bin > 0; --bin) {
                if (cum + (int)sc.hist[bin] >= want) break;
                cum += (int)sc.hist[bin];
            }
            sc.u[1] = (unsigned)bin; sc.u[2] = (unsigned)(want - cum);
        }
        __syncthreads();
        prefix |= sc.u[1] << shift; pmask |= 0xffu << shift; want = (int)sc.u[2];
    }
    const unsigned kth_key = prefix;          // `want` = how many elements equal to kth_key belong to the top-k
    const int i0 = (int)sc.u[0];
    // 4. p0 = mlp([h_out_src | h_out_tgt][i0])  (evl:736-742; only pred[0] survives the scatter)
    if (warp == 0) {
        const float zs = __ldg(a.h_out_src + (nb + i0) * H + lane), zt = __ldg(a.h_out_tgt + (nb + i0) * H + lane);
        const float *hp = a.head_pack;
        float h0 = __ldg(hp + HOFF_B0 + lane);
        for (int i = 0; i < 32; ++i) h0 = fmaf(__ldg(hp + HOFF_W0T + 32 * i + lane), __shfl_sync(0xffffffffu, zs, i), h0);
        for (int i = 0; i < 32; ++i) h0 = fmaf(__ldg(hp + HOFF_W0T + 32 * (32 + i) + lane), __shfl_sync(0xffffffffu, zt, i), h0);
        h0 = fmaxf(h0, 0.f);
        float h1 = lane < 16 ? __ldg(hp + HOFF_B1 + lane) : 0.f;
        for (int i = 0; i < 32; ++i) {
            const float v = __shfl_sync(0xffffffffu, h0, i);
            if (lane < 16) h1 = fmaf(__ldg(hp + HOFF_W1T + 16 * i + lane), v, h1);
        }
        h1 = fmaxf(h1, 0.f);
        float pr = lane < 16 ? h1 * __ldg(hp + HOFF_W2 + lane) : 0.f;
        pr = warp_sum(pr);
        if (lane == 0) sc.bc[15] = pr + __ldg(hp + HOFF_B2);
    }
    __syncthreads();
    const float p0 = sc.bc[15];
    // 5. final weights: members of the top-k set take p0 when the (quirky) conditions hold (evl:761-768)
    float part[1] = {0.f};
    unsigned eq_carry = 0;       // equals seen in earlier index chunks
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        float s = 0.f;
        unsigned key = 0;
        if (i < n) { s = ssim[i]; key = order_key(s); }
        const bool is_eq = (i < n) && key == kth_key;
        const unsigned bal = __ballot_sync(0xffffffffu, is_eq);
        __syncthreads();
        if (lane == 0) sc.hist[warp] = __popc(bal);
        __syncthreads();
        unsigned before = eq_carry, total = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { const unsigned c = sc.hist[w]; if (w < warp) before += c; total += c; }
        before += __popc(bal & ((1u << lane) - 1u));
        eq_carry += total;
        const bool member = (i < n) && (key > kth_key || (is_eq && (int)before < want));
        float f = s;
        if (member && p0 > 0.5f && (fabsf(p0 - 1.0f) < s || p0 < s)) f = p0;
        if (i < n) { ssim[i] = f; part[0] += f; }
    }
    block_sum<1>(part, sc);
    const float inv_s = 1.0f / (part[0] + 1e-6f);                                  // evl:771
    float mx = -3.4e38f;
    // (the four passes below keep four independent loads in flight per thread; per-thread partial sums are taken in
    // index order i, i + T, i + 2T, ... exactly as a plain strided loop would)
#define EGSPR_PASS4(BODY)                                                                              \
    for (int i0 = tid; i0 < n; i0 += 4 * blockDim.x) {                                                 \
        float v4[4];                                                                                   \
        _Pragma("unroll") for (int q = 0; q < 4; ++q) { const int i = i0 + q * blockDim.x; v4[q] = i < n ? ssim[i] : 0.f; } \
        _Pragma("unroll") for (int q = 0; q < 4; ++q) { const int i = i0 + q * blockDim.x; if (i < n) { const float v = v4[q]; BODY } } \
    }
    EGSPR_PASS4({ const float f = v * inv_s; ssim[i] = f; mx = fmaxf(mx, f); })
    mx = block_max(mx, sc);
    float z[1] = {0.f};
    EGSPR_PASS4({ const float e = expf(v - mx); ssim[i] = e; z[0] += e; })                     // softmax evl:774
    block_sum<1>(z, sc);
    const float inv_z = 1.0f / z[0];
    float sw[1] = {0.f};
    EGSPR_PASS4({ const float w = v * inv_z; ssim[i] = w; sw[0] += w; })
    block_sum<1>(sw, sc);
    const float inv_w = 1.0f / (sw[0] + 1e-6f);                                    // evl:783
    EGSPR_PASS4({ const float w = v * inv_w; ssim[i] = w; if (a.w_out) a.w_out[nb + i] = w; })
#undef EGSPR_PASS4
    __syncthreads();
    // 6. Kabsch on the original coordinates, all n points (evl:717-718, 786-818)
    block_kabsch(a.x_src + nb * 3, a.x_tgt + nb * 3, 3, ssim, n, n, a.R + b * 9, a.t + b * 3,
                 a.Hout ? a.Hout + b * 9 : nullptr, sc);
    // 7. egnn_equi_loss partial sums on the EGNN outputs (evl:687)
    if (a.loss_parts && a.split > 0) {
        if (tid < 2) {
            float t2 = 0.f;
            for (int i = 0; i < a.split; ++i) t2 += a.pre_lp[((size_t)b * a.split + i) * 2 + tid];
            a.loss_parts[b * 2 + tid] = t2;
        }
    } else if (a.loss_parts)
        block_equi_loss(a.h_out_src + nb * H, a.h_out_tgt + nb * H, a.x_out_src + nb * 3, a.x_out_tgt + nb * 3,
                        a.labels + nb, a.gt_pose + b * 16, n, a.loss_parts + b * 2, sc);
}

struct HeadTrainArgs {
    const float *h_out_src, *h_out_tgt, *x_out_src, *x_out_tgt, *labels, *gt_pose;
    int n;
    float *w_out, *sim_out, *R, *t, *Hout, *loss_parts;
};

__global__ void __launch_bounds__(HD_THREADS_BIG) head_train_kernel(const HeadTrainArgs a) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ BlockScratch sc;
    const int b = blockIdx.x, n = a.n, tid = threadIdx.x;
    const size_t nb = (size_t)b * n;
    float *ssim = (n > HD_MAX_N) ? a.w_out + nb : dyn;
    float mx = -3.4e38f;
    float cnt[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) {
        const float s = dot32(a.h_out_src + (nb + i) * H, a.h_out_tgt + (nb + i) * H);   // 3dm:681, 717
        if (a.sim_out) a.sim_out[nb + i] = s;
        const bool valid = __ldg(a.labels + nb + i) != 0.f;                              // 3dm:696
        ssim[i] = valid ? s : -3.4e38f;
        if (valid) { mx = fmaxf(mx, s); cnt[0] += 1.f; }
    }
    mx = block_max(mx, sc);
    block_sum<1>(cnt, sc);
    float z[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) {
        const float v = ssim[i];
        const float e = v > -3.0e38f ? expf(v - mx) : 0.f;                               // softmax over the valid set 3dm:718
        ssim[i] = e; z[0] += e;
    }
    block_sum<1>(z, sc);
    const float inv_z = z[0] > 0.f ? 1.0f / z[0] : 0.f;
    float sw[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) { const float w = ssim[i] * inv_z; ssim[i] = w; sw[0] += w; }
    block_sum<1>(sw, sc);
    const float inv_w = 1.0f / (sw[0] + 1e-6f);                                          // 3dm:724
    for (int i = tid; i < n; i += blockDim.x) {
        const float w = ssim[i] * inv_w;
        ssim[i] = w;
        if (a.w_out) a.w_out[nb + i] = w;
    }
    __syncthreads();
    block_kabsch(a.x_out_src + nb * 3, a.x_out_tgt + nb * 3, 3, ssim, n, (int)(cnt[0] + 0.5f), a.R + b * 9,
                 a.t + b * 3, a.Hout ? a.Hout + b * 9 : nullptr, sc);
    if (a.loss_parts)
        block_equi_loss(a.h_out_src + nb * H, a.h_out_tgt + nb * H, a.x_out_src + nb * 3, a.x_out_tgt + nb * 3,
                        a.labels + nb, a.gt_pose + b * 16, n, a.loss_parts + b * 2, sc);
}

template <int MAXT>
__global__ void __launch_bounds__(MAXT) kabsch_kernel(const float *__restrict__ p, const float *__restrict__ q,
                                                            const float *__restrict__ w, const float *__restrict__ mask,
                                                            int n, float *R, float *t, float *Hout) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ BlockScratch sc;
    const int b = blockIdx.x;
    const size_t nb = (size_t)b * n;
    float cnt[1] = {0.f};
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const bool inc = !mask || __ldg(mask + nb + i) != 0.f;
        dyn[i] = inc ? __ldg(w + nb + i) : 0.f;
        if (inc) cnt[0] += 1.f;
    }
    block_sum<1>(cnt, sc);
    block_kabsch(p + nb * 3, q + nb * 3, 3, dyn, n, (int)(cnt[0] + 0.5f), R + b * 9, t + b * 3,
                 Hout ? Hout + b * 9 : nullptr, sc);
}

// ---- a17: tools/evaluation_metrics.py:14-43 (+ the F1 of src/eval_egnn_metrics.py:1277) on the device ----
// One CTA per pair, fp64 like the reference's numpy code (float32 inputs promoted by the float64 pose):
//   TE = 100 |t_gt - t|  (cm),  RE = deg(acos(clip((tr(R_gt^T R) - 1) / 2, -1, 1))),
//   TP = #{ |R p_i + t - q_i| < tau },  recall = sqrt(TP / n),  precision = TP / n,
//   F1 = 2 P R / (P + R + 1e-6).      out[pair] = (RE, TE, recall, precision, F1)
__global__ void __launch_bounds__(256) pose_metrics_kernel(const float *__restrict__ R, const float *__restrict__ t,
                                                           const float *__restrict__ gt_pose, const float *__restrict__ src,
                                                           const float *__restrict__ tgt, int n, double tau,
                                                           double *__restrict__ out) {
    __shared__ int warp_cnt[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double Rm[9], tv[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) Rm[i] = (double)__ldg(R + b * 9 + i);
#pragma unroll
    for (int i = 0; i < 3; ++i) tv[i] = (double)__ldg(t + b * 3 + i);
    const float *p = src + (size_t)b * n * 3, *q = tgt + (size_t)b * n * 3;
    int cnt = 0;
    for (int i = tid; i < n; i += 256) {
        const double x = (double)__ldg(p + 3 * i), y = (double)__ldg(p + 3 * i + 1), z = (double)__ldg(p + 3 * i + 2);
        const double dx = (Rm[0] * x + Rm[1] * y + Rm[2] * z) + tv[0] - (double)__ldg(q + 3 * i);
        const double dy = (Rm[3] * x + Rm[4] * y + Rm[5] * z) + tv[1] - (double)__ldg(q + 3 * i + 1);
        const double dz = (Rm[6] * x + Rm[7] * y + Rm[8] * z) + tv[2] - (double)__ldg(q + 3 * i + 2);
        cnt += (sqrt(dx * dx + dy * dy + dz * dz) < tau) ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if (lane == 0) warp_cnt[warp] = cnt;
    __syncthreads();
    if (tid == 0) {
        int tp = 0;
        for (int w = 0; w < 8; ++w) tp += warp_cnt[w];
        const float *g = gt_pose + b * 16;
        double te = 0.0, tr = 0.0;
#pragma unroll
        for (int i = 0; i < 3; ++i) { const double d = (double)g[4 * i + 3] - tv[i]; te += d * d; }
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) tr += (double)g[4 * k + i] * Rm[3 * k + i];        // trace(R_gt^T R)
        double c = (tr - 1.0) / 2.0;
        c = c < -1.0 ? -1.0 : (c > 1.0 ? 1.0 : c);
        const double prec = n > 0 ? (double)tp / (double)n : 0.0;
        const double rec = n > 0 ? sqrt((double)tp / (double)n) : 0.0;
        out[b * 5 + 0] = acos(c) * (180.0 / 3.14159265358979323846);
        out[b * 5 + 1] = sqrt(te) * 100.0;
        out[b * 5 + 2] = rec;
        out[b * 5 + 3] = prec;
        out[b * 5 + 4] = 2.0 * (prec * rec) / (prec + rec + 1e-6);
    }
}

// ---- backward of the train-variant head (3dm:696-758): (dR, dt, dsim) -> gradients of the EGNN outputs ----------
// R = V D U^T of H = U S V^T is differentiated in closed form (no 1/(s_i^2 - s_j^2) terms for equal-sign pairs):
//   A = V^T (dR - dt cs^T) U,   Mbar_ij = alpha_ij A_ij + beta_ji A_ji (i != j),   dH = U Mbar V^T
//   same-sign pair: alpha_ij = -d/(s_i+s_j), beta_ij = d/(s_i+s_j);  mixed pair: alpha_ij = beta_ij = d_i/(s_i-s_j)
// then through H, the centroids, w = softmax(sim)/(sum+1e-6) to sim, x_src_out, x_tgt_out, and sim = <h_s,h_t> to h.
struct HeadTrainBwdArgs {
    const float *h_out_src, *h_out_tgt, *x_out_src, *x_out_tgt, *labels, *dR, *dt, *dsim;
    int n;
    float *dh_src, *dh_tgt, *dx_src, *dx_tgt;
    // optional: backward of the correspondence BCE (3dm:760-772) through mlp into the top-k rows and the head pack
    const int32_t *top_idx;     // [pairs][k] or null
    const float *head_pack;
    float *head_gpack;
    int k, pairs;
    const float *corr_scale;    // device scalar: upstream gradient of the mean BCE (loss[7] of egspr_train_loss_finalize)
};

// mlp 64 -> 32 -> 16 -> 1 (3dm:594-600) on one row [zs | zt], one warp: lane = output unit.  w = head pack (shared memory).
struct MlpRow { float h0, h1, score; };      // post-ReLU activations of this lane's unit (h1: lanes < 16)
__device__ __forceinline__ MlpRow mlp_row(const float *__restrict__ w, float zs, float zt, int lane) {
    MlpRow r;
    float h0 = w[HOFF_B0 + lane];
#pragma unroll 8
    for (int i = 0; i < 32; ++i) h0 = fmaf(w[HOFF_W0T + 32 * i + lane], __shfl_sync(0xffffffffu, zs, i), h0);
#pragma unroll 8
    for (int i = 0; i < 32; ++i) h0 = fmaf(w[HOFF_W0T + 32 * (32 + i) + lane], __shfl_sync(0xffffffffu, zt, i), h0);
    r.h0 = fmaxf(h0, 0.f);
    float h1 = lane < 16 ? w[HOFF_B1 + lane] : 0.f;
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {
        const float v = __shfl_sync(0xffffffffu, r.h0, i);
        if (lane < 16) h1 = fmaf(w[HOFF_W1T + 16 * i + lane], v, h1);
    }
    r.h1 = fmaxf(h1, 0.f);
    float pr = lane < 16 ? r.h1 * w[HOFF_W2 + lane] : 0.f;
    r.score = warp_sum(pr) + w[HOFF_B2];
    return r;
}

__global__ void __launch_bounds__(HD_THREADS_BIG) head_train_backward_kernel(const HeadTrainBwdArgs a) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ BlockScratch sc;
    __shared__ float gsh[24];      // G_H [9], dcs' [3], dct' [3], cs [3], ct [3], ok
    const int b = blockIdx.x, n = a.n, tid = threadIdx.x;
    const size_t nb = (size_t)b * n;
    float *sw = dyn, *sdw = dyn + n;
    const float *p = a.x_out_src + nb * 3, *q = a.x_out_tgt + nb * 3;
    // forward recompute: weights
    float mx = -3.4e38f;
    float cnt[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) {
        const float s = dot32(a.h_out_src + (nb + i) * H, a.h_out_tgt + (nb + i) * H);
        const bool valid = __ldg(a.labels + nb + i) != 0.f;
        sw[i] = valid ? s : -3.4e38f;
        if (valid) { mx = fmaxf(mx, s); cnt[0] += 1.f; }
    }
    mx = block_max(mx, sc);
    block_sum<1>(cnt, sc);
    float z[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) {
        const float v = sw[i];
        const float e = v > -3.0e38f ? expf(v - mx) : 0.f;
        sw[i] = e; z[0] += e;
    }
    block_sum<1>(z, sc);
    const float inv_z = z[0] > 0.f ? 1.0f / z[0] : 0.f;
    float swt[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) { const float w = sw[i] * inv_z; sw[i] = w; swt[0] += w; }
    block_sum<1>(swt, sc);
    const float inv_w = 1.0f / (swt[0] + 1e-6f);
    for (int i = tid; i < n; i += blockDim.x) sw[i] *= inv_w;
    __syncthreads();
    // centroids, H, sums of w*pc, w*qc
    float c[7] = {0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += blockDim.x) {
        const float wi = sw[i];
        if (wi != 0.f) {
            c[0] = fmaf(wi, p[i * 3], c[0]); c[1] = fmaf(wi, p[i * 3 + 1], c[1]); c[2] = fmaf(wi, p[i * 3 + 2], c[2]);
            c[3] = fmaf(wi, q[i * 3], c[3]); c[4] = fmaf(wi, q[i * 3 + 1], c[4]); c[5] = fmaf(wi, q[i * 3 + 2], c[5]);
            c[6] += wi;
        }
    }
    block_sum<7>(c, sc);
    float hm[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = tid; i < n; i += blockDim.x) {
        const float wi = sw[i];
        if (wi != 0.f) {
            const float a0 = wi * (p[i * 3] - c[0]), a1 = wi * (p[i * 3 + 1] - c[1]), a2 = wi * (p[i * 3 + 2] - c[2]);
            const float b0 = q[i * 3] - c[3], b1 = q[i * 3 + 1] - c[4], b2 = q[i * 3 + 2] - c[5];
            hm[0] = fmaf(a0, b0, hm[0]); hm[1] = fmaf(a0, b1, hm[1]); hm[2] = fmaf(a0, b2, hm[2]);
            hm[3] = fmaf(a1, b0, hm[3]); hm[4] = fmaf(a1, b1, hm[4]); hm[5] = fmaf(a1, b2, hm[5]);
            hm[6] = fmaf(a2, b0, hm[6]); hm[7] = fmaf(a2, b1, hm[7]); hm[8] = fmaf(a2, b2, hm[8]);
        }
    }
    block_sum<9>(hm, sc);
    if (tid == 0) {
        bool ok = cnt[0] > 0.5f;
        double GH[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, dcs[3] = {0, 0, 0}, dct[3] = {0, 0, 0};
        if (ok) {
            hm[0] += 1e-6f; hm[4] += 1e-6f; hm[8] += 1e-6f;
            double Hd[3][3], U[3][3], W[3][3], sg[3], d3;
            for (int i = 0; i < 9; ++i) Hd[i / 3][i % 3] = (double)hm[i];
            ok = kabsch_svd(Hd, U, W, sg, d3);
            if (ok) {
                const double dd[3] = {1.0, 1.0, d3};
                double R[3][3], GR[3][3], A[3][3], Mb[3][3];
                double dtv[3] = {(double)a.dt[b * 3], (double)a.dt[b * 3 + 1], (double)a.dt[b * 3 + 2]};
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        R[i][j] = W[i][0] * U[j][0] + W[i][1] * U[j][1] + d3 * W[i][2] * U[j][2];
                        GR[i][j] = (double)a.dR[b * 9 + i * 3 + j] - dtv[i] * (double)c[j];     // t = ct - R cs
                    }
                for (int i = 0; i < 3; ++i) {
                    dcs[i] = -(R[0][i] * dtv[0] + R[1][i] * dtv[1] + R[2][i] * dtv[2]);
                    dct[i] = dtv[i];
                }
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        double t = 0;
                        for (int k = 0; k < 3; ++k)
                            for (int l = 0; l < 3; ++l) t += W[k][i] * GR[k][l] * U[l][j];
                        A[i][j] = t;
                    }
                auto coef = [&](int i, int j, bool beta) {   // alpha_ij / beta_ij
                    if (dd[i] == dd[j]) {
                        const double den = fmax(sg[i] + sg[j], 1e-300);
                        return (beta ? dd[i] : -dd[i]) / den;
                    }
                    double den = sg[i] - sg[j];
                    if (fabs(den) < 1e-300) den = den < 0 ? -1e-300 : 1e-300;
                    return dd[i] / den;
                };
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) Mb[i][j] = (i == j) ? 0.0 : coef(i, j, false) * A[i][j] + coef(j, i, true) * A[j][i];
                for (int i = 0; i < 3; ++i)
                    for (int j = 0; j < 3; ++j) {
                        double t = 0;
                        for (int k = 0; k < 3; ++k)
                            for (int l = 0; l < 3; ++l) t += U[i][k] * Mb[k][l] * W[j][l];
                        GH[i][j] = t;
                    }
                // centroids enter pc, qc as well:  dcs' = dcs - G_H sum_j w_j qc_j,  dct' = dct - G_H^T sum_j w_j pc_j
                const double om = 1.0 - (double)c[6];       // sum_j w_j qc_j = ct (1 - sum w), likewise for pc
                for (int i = 0; i < 3; ++i) {
                    double u = 0, v = 0;
                    for (int j = 0; j < 3; ++j) { u += GH[i][j] * (double)c[3 + j] * om; v += GH[j][i] * (double)c[j] * om; }
                    dcs[i] -= u; dct[i] -= v;
                }
            }
        }
        for (int i = 0; i < 9; ++i) gsh[i] = (float)GH[i / 3][i % 3];
        for (int i = 0; i < 3; ++i) { gsh[9 + i] = (float)dcs[i]; gsh[12 + i] = (float)dct[i]; gsh[15 + i] = c[i]; gsh[18 + i] = c[3 + i]; }
        gsh[21] = ok ? 1.f : 0.f;
    }
    __syncthreads();
    const bool ok = gsh[21] != 0.f;
    float g[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) g[i] = gsh[i];
    const float dcs0 = gsh[9], dcs1 = gsh[10], dcs2 = gsh[11], dct0 = gsh[12], dct1 = gsh[13], dct2 = gsh[14];
    const float cs0 = gsh[15], cs1 = gsh[16], cs2 = gsh[17], ct0 = gsh[18], ct1 = gsh[19], ct2 = gsh[20];
    // pass A: dw_i, coordinate gradients, D1 = sum_i w_i dw_i
    float d1[1] = {0.f};
    for (int i = tid; i < n; i += blockDim.x) {
        const float wi = sw[i];
        float dw = 0.f, dp0 = 0.f, dp1 = 0.f, dp2 = 0.f, dq0 = 0.f, dq1 = 0.f, dq2 = 0.f;
        if (ok && wi != 0.f) {
            const float p0 = p[i * 3], p1 = p[i * 3 + 1], p2 = p[i * 3 + 2], q0 = q[i * 3], q1 = q[i * 3 + 1], q2 = q[i * 3 + 2];
            const float pc0 = p0 - cs0, pc1 = p1 - cs1, pc2 = p2 - cs2, qc0 = q0 - ct0, qc1 = q1 - ct1, qc2 = q2 - ct2;
            const float gq0 = g[0] * qc0 + g[1] * qc1 + g[2] * qc2, gq1 = g[3] * qc0 + g[4] * qc1 + g[5] * qc2,
                        gq2 = g[6] * qc0 + g[7] * qc1 + g[8] * qc2;                       // G_H qc
            const float gp0 = g[0] * pc0 + g[3] * pc1 + g[6] * pc2, gp1 = g[1] * pc0 + g[4] * pc1 + g[7] * pc2,
                        gp2 = g[2] * pc0 + g[5] * pc1 + g[8] * pc2;                       // G_H^T pc
            dw = pc0 * gq0 + pc1 * gq1 + pc2 * gq2 + p0 * dcs0 + p1 * dcs1 + p2 * dcs2 + q0 * dct0 + q1 * dct1 + q2 * dct2;
            dp0 = wi * (gq0 + dcs0); dp1 = wi * (gq1 + dcs1); dp2 = wi * (gq2 + dcs2);
            dq0 = wi * (gp0 + dct0); dq1 = wi * (gp1 + dct1); dq2 = wi * (gp2 + dct2);
            d1[0] = fmaf(wi, dw, d1[0]);
        }
        sdw[i] = dw;
        float *o = a.dx_src + (nb + i) * 3, *o2 = a.dx_tgt + (nb + i) * 3;
        o[0] = dp0; o[1] = dp1; o[2] = dp2; o2[0] = dq0; o2[1] = dq1; o2[2] = dq2;
    }
    block_sum<1>(d1, sc);
    const float D1 = d1[0] * (1.0f + 1e-6f);
    // pass B: d sim_i = w_i (dw_i - D1 (1 + eps)) + external dsim_i;  sim = <h_s, h_t>
    const int lane8 = tid & 7;
    for (int i0 = (tid >> 3); i0 < n; i0 += (blockDim.x >> 3)) {      // 8 lanes per 128-byte row
        const int i = i0;
        float ds = sw[i] * (sdw[i] - D1);
        if (a.dsim) ds += __ldg(a.dsim + nb + i);
        const float4 hs = ldg4(a.h_out_src + (nb + i) * H + 4 * lane8), ht = ldg4(a.h_out_tgt + (nb + i) * H + 4 * lane8);
        *reinterpret_cast<float4 *>(a.dh_src + (nb + i) * H + 4 * lane8) = make_float4(ds * ht.x, ds * ht.y, ds * ht.z, ds * ht.w);
        *reinterpret_cast<float4 *>(a.dh_tgt + (nb + i) * H + 4 * lane8) = make_float4(ds * hs.x, ds * hs.y, ds * hs.z, ds * hs.w);
    }
}

// ---- correspondence loss backward (3dm:760-772): BCEWithLogits(mlp([h_s | h_t][top-k]), labels[top-k]), mean over
// pairs * k.  grid (pairs, CB_SPLIT), one warp per selected row: recompute the mlp, push d score back through it into the
// row of dh_src / dh_tgt (already written by head_train_backward_kernel; the rows of a pair's top-k set are distinct) and
// into the head-pack gradient (registers per warp -> shared memory per CTA -> one global atomicAdd per entry and CTA). ----
constexpr int CB_SPLIT = 4, CB_THREADS = 256;
__global__ void __launch_bounds__(CB_THREADS) corr_loss_backward_kernel(const HeadTrainBwdArgs a) {
    __shared__ __align__(16) float swp[HEAD_PACK];
    __shared__ float sgp[HEAD_PACK];
    const int b = blockIdx.x, n = a.n, tid = threadIdx.x;
    const size_t nb = (size_t)b * n;
    for (int i = tid; i < HEAD_PACK; i += blockDim.x) { swp[i] = __ldg(a.head_pack + i); sgp[i] = 0.f; }
    __syncthreads();
    const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const float gscale = __ldg(a.corr_scale) / (float)(a.pairs * (a.k < n ? a.k : n));
    // weight gradients: every lane owns the entries of ITS unit in registers across the warp's rows (lane = output unit of
    // the first / second Linear, lane < 16: of the last two), one shared-memory atomicAdd per entry and WARP at the end
    float gW0[64], gW1[32], gb0 = 0.f, gb1 = 0.f, gw2 = 0.f, gb2 = 0.f;      // dW0T[in][lane], dW1T[in][lane < 16]
#pragma unroll
    for (int i = 0; i < 64; ++i) gW0[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) gW1[i] = 0.f;
    for (int r = blockIdx.y * nwarps + warp; r < a.k; r += gridDim.y * nwarps) {
        const int i = __ldg(a.top_idx + (size_t)b * a.k + r);
        if (i < 0) continue;                                       // fewer than k points
        const float zs = __ldg(a.h_out_src + (nb + i) * H + lane), zt = __ldg(a.h_out_tgt + (nb + i) * H + lane);
        const MlpRow m = mlp_row(swp, zs, zt, lane);
        const float y = __ldg(a.labels + nb + i);
        const float dsc = gscale * (1.0f / (1.0f + expf(-m.score)) - y);          // d BCEWithLogits / d score
        // last Linear (16 -> 1)
        float dh1 = 0.f;
        if (lane < 16) {
            gw2 = fmaf(dsc, m.h1, gw2);
            dh1 = m.h1 > 0.f ? dsc * swp[HOFF_W2 + lane] : 0.f;
            gb1 += dh1;
        }
        gb2 += dsc;
        // middle Linear (32 -> 16): dW1T[i][j] += h0[i] dh1[j] (lane j < 16 owns column j);  dh0[i] = sum_j W1T[i][j] dh1[j]
        float dh0 = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) dh0 = fmaf(swp[HOFF_W1T + 16 * lane + j], __shfl_sync(0xffffffffu, dh1, j), dh0);
#pragma unroll
        for (int in = 0; in < 32; ++in) gW1[in] = fmaf(__shfl_sync(0xffffffffu, m.h0, in), dh1, gW1[in]);
        dh0 = m.h0 > 0.f ? dh0 : 0.f;
        gb0 += dh0;
        // first Linear (64 -> 32): dW0T[in][out] += z[in] dh0[out] (lane owns column out = lane);  dz[in] = sum_out W0T[in][out] dh0[out]
        float dzs = 0.f, dzt = 0.f;
#pragma unroll 8
        for (int o = 0; o < 32; ++o) {
            const float d = __shfl_sync(0xffffffffu, dh0, o);
            dzs = fmaf(swp[HOFF_W0T + 32 * lane + o], d, dzs);
            dzt = fmaf(swp[HOFF_W0T + 32 * (32 + lane) + o], d, dzt);
        }
#pragma unroll
        for (int in = 0; in < 32; ++in) {
            gW0[in] = fmaf(__shfl_sync(0xffffffffu, zs, in), dh0, gW0[in]);
            gW0[32 + in] = fmaf(__shfl_sync(0xffffffffu, zt, in), dh0, gW0[32 + in]);
        }
        a.dh_src[(nb + i) * H + lane] += dzs;
        a.dh_tgt[(nb + i) * H + lane] += dzt;
    }
#pragma unroll
    for (int in = 0; in < 64; ++in) atomicAdd(sgp + HOFF_W0T + 32 * in + lane, gW0[in]);
    atomicAdd(sgp + HOFF_B0 + lane, gb0);
    if (lane < 16) {
#pragma unroll
        for (int in = 0; in < 32; ++in) atomicAdd(sgp + HOFF_W1T + 16 * in + lane, gW1[in]);
        atomicAdd(sgp + HOFF_B1 + lane, gb1);
        atomicAdd(sgp + HOFF_W2 + lane, gw2);
    }
    if (lane == 0) atomicAdd(sgp + HOFF_B2, gb2);
    __syncthreads();
    for (int i = tid; i < HEAD_PACK; i += blockDim.x) {
        const float g = sgp[i];
        if (g != 0.f) atomicAdd(a.head_gpack + i, g);
    }
}

// d acos(clamp(c, -1, 1)) / dc.  Outside [-1, 1] the clamp blocks the gradient (0, as torch.clamp).  AT c == +-1 exactly
// torch's autograd returns -+inf (clamp passes the boundary, acos' is singular there): a prediction that equals the ground
// truth to fp32 rounding -- reachable on noise-free synthetic pairs -- then turns every EGNN gradient into NaN and the next
// Adam step destroys the model.  Deliberate deviation from the reference at that single point: a zero subgradient.
__device__ __forceinline__ float acos_slope(float c) {
    const float s = 1.0f - c * c;
    return s > 0.f ? -1.0f / sqrtf(s) : 0.f;
}

// ---- training losses on the device: 3dm:681-694 (top-k of the output-feature similarity), 3dm:760-781
// (BCE of mlp([h_s | h_t][top-k]) against the labels; MSE between the z-scored similarity and the z-scored input-feature
// similarity, mean / unbiased std over the whole batch), plus pose_loss 3dm:896-962 and the loop's total 3dm:1118. ----
struct TrainLossArgs {
    const float *h_out_src, *h_out_tgt, *feat_src, *feat_tgt, *sim, *labels, *head_pack;
    int n, k;
    int32_t *top_idx;       // [pairs][k]: members of the top-k set in ascending point order (-1 past min(k, n))
    float *scores;          // [pairs][k]: mlp logits of those rows
    float *raw;             // [pairs][n]: input-feature similarity
    double *stats;          // [pairs][4]: sum sim, sum sim^2, sum raw, sum raw^2
    float *bce;             // [pairs]: sum over the pair's rows of BCEWithLogits
};

__global__ void __launch_bounds__(HD_THREADS_BIG) train_loss_forward_kernel(const TrainLossArgs a) {
    extern __shared__ __align__(16) float dyn[];
    __shared__ BlockScratch sc;
    __shared__ double dred[HD_WARPS][4];
    const int b = blockIdx.x, n = a.n, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
    const size_t nb = (size_t)b * n;
    float *ssim = dyn;                       // [n]
    int *slist = reinterpret_cast<int *>(dyn + n);      // [k]
    float *swp = dyn + n + a.k;              // [HEAD_PACK]
    for (int i = tid; i < HEAD_PACK; i += blockDim.x) swp[i] = __ldg(a.head_pack + i);
    double acc[4] = {0, 0, 0, 0};
    for (int i = tid; i < n; i += blockDim.x) {
        // sim = <h_out_src, h_out_tgt> (3dm:681): read from egspr_head_train's output, or recomputed (same dot32, same bits)
        // so that this kernel does not depend on that one and the two narrow launches can run side by side
        const float s = a.sim ? __ldg(a.sim + nb + i) : dot32(a.h_out_src + (nb + i) * H, a.h_out_tgt + (nb + i) * H);
        ssim[i] = s;
        const float r = dot32(a.feat_src + (nb + i) * H, a.feat_tgt + (nb + i) * H);     // 3dm:773
        a.raw[nb + i] = r;
        acc[0] += (double)s; acc[1] += (double)s * (double)s; acc[2] += (double)r; acc[3] += (double)r * (double)r;
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
        if (lane == 0) dred[warp][q] = acc[q];
    }
    __syncthreads();
    if (tid < 4) {
        double t = 0;
        for (int w = 0; w < nwarps; ++w) t += dred[w][tid];
        a.stats[(size_t)b * 4 + tid] = t;
    }
    // k-th largest similarity by 4-pass radix select (torch.topk; ties -> lower index), as in head_eval_kernel
    const int kk = a.k < n ? a.k : n;
    unsigned prefix = 0, pmask = 0;
    int want = kk;
    for (int shift = 24; shift >= 0; shift -= 8) {
        __syncthreads();
        if (tid < 256) sc.hist[tid] = 0;
        __syncthreads();
        // four independent loads in flight per thread (large clouds keep sim0 in global memory) and warp-aggregated
        // histogram updates (similarities crowd into a few bins: one atomic per distinct bin and warp, not per lane)
        for (int base = 0; base < n; base += 4 * blockDim.x) {        // block-uniform trip count: __match_any_sync needs whole warps
            const int i0 = base + tid;
            float v4[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) { const int i = i0 + q * blockDim.x; v4[q] = i < n ? ssim[i] : 0.f; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const unsigned key = order_key(v4[q]);
                const bool in = (i0 + q * (int)blockDim.x < n) && (key & pmask) == prefix;
                const unsigned bin = in ? ((key >> shift) & 0xffu) : 0x100u;
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (in && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sc.hist[bin], (unsigned)__popc(peers));
            }
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, bin = 255;
            for (; bin > 0; --bin) {
                if (cum + (int)sc.hist[bin] >= want) break;
                cum += (int)sc.hist[bin];
            }
            sc.u[1] = (unsigned)bin; sc.u[2] = (unsigned)(want - cum);
        }
        __syncthreads();
        prefix |= sc.u[1] << shift; pmask |= 0xffu << shift; want = (int)sc.u[2];
    }
    const unsigned kth_key = prefix;
    unsigned eq_carry = 0, sel_carry = 0;
    for (int base = 0; base < n; base += blockDim.x) {
        const int i = base + tid;
        unsigned key = 0;
        if (i < n) key = order_key(ssim[i]);
        const bool is_eq = (i < n) && key == kth_key;
        const unsigned bal = __ballot_sync(0xffffffffu, is_eq);
        __syncthreads();
        if (lane == 0) sc.hist[warp] = __popc(bal);
        __syncthreads();
        unsigned before = eq_carry, total = 0;
        for (int w = 0; w < nwarps; ++w) { const unsigned c = sc.hist[w]; if (w < warp) before += c; total += c; }
        before += __popc(bal & ((1u << lane) - 1u));
        eq_carry += total;
        // members of the set go to the list in INDEX order (block-wide prefix of the selection flags, not an atomic
        // cursor): the BCE sum below then adds the same rows in the same order on every run -- bit-reproducible loss
        const bool sel = (i < n) && (key > kth_key || (is_eq && (int)before < want));
        const unsigned sbal = __ballot_sync(0xffffffffu, sel);
        if (lane == 0) sc.hist[32 + warp] = __popc(sbal);
        __syncthreads();
        unsigned sbefore = sel_carry, stotal = 0;
        for (int w = 0; w < nwarps; ++w) { const unsigned c = sc.hist[32 + w]; if (w < warp) sbefore += c; stotal += c; }
        sbefore += __popc(sbal & ((1u << lane) - 1u));
        sel_carry += stotal;
        if (sel) slist[sbefore] = i;
    }
    __syncthreads();
    float bsum[1] = {0.f};
    for (int r = warp; r < a.k; r += nwarps) {
        if (r >= kk) {
            if (lane == 0) { a.top_idx[(size_t)b * a.k + r] = -1; a.scores[(size_t)b * a.k + r] = 0.f; }
            continue;
        }
        const int i = slist[r];
        const float zs = __ldg(a.h_out_src + (nb + i) * H + lane), zt = __ldg(a.h_out_tgt + (nb + i) * H + lane);
        const MlpRow m = mlp_row(swp, zs, zt, lane);
        if (lane == 0) {
            const float x = m.score, y = __ldg(a.labels + nb + i);
            bsum[0] += fmaxf(x, 0.f) - x * y + log1pf(expf(-fabsf(x)));              // BCEWithLogits
            a.top_idx[(size_t)b * a.k + r] = i;
            a.scores[(size_t)b * a.k + r] = x;
        }
    }
    block_sum<1>(bsum, sc);
    if (tid == 0) a.bce[b] = bsum[0];
}

// one CTA: batch statistics, losses, and the seeds of the backward pass
//   loss[0..7] = corr, sim, mean rot, mean trans, total (3dm:1118), 0, 0, upstream scale of the mean BCE (= scale)
__global__ void __launch_bounds__(1024) train_loss_finalize_kernel(const float *__restrict__ sim, const float *__restrict__ raw,
                                                                   const double *__restrict__ stats, const float *__restrict__ bce,
                                                                   int pairs, int n, int k, const float *__restrict__ R,
                                                                   const float *__restrict__ t, const float *__restrict__ gt_pose,
                                                                   float scale, float *__restrict__ loss, float *__restrict__ dsim,
                                                                   float *__restrict__ dR, float *__restrict__ dt) {
    __shared__ double dred[32][4];
    __shared__ double bc[8];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long M = (long long)pairs * n;
    auto block_sum_d = [&](double (&v)[4]) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        }
        __syncthreads();
        if (lane == 0) { dred[warp][0] = v[0]; dred[warp][1] = v[1]; dred[warp][2] = v[2]; dred[warp][3] = v[3]; }
        __syncthreads();
        if (tid < 4) {
            double s = 0;
            for (int w = 0; w < 32; ++w) s += dred[w][tid];
            bc[tid] = s;
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = bc[q];
    };
    double st[4] = {0, 0, 0, 0};
    for (int b = tid; b < pairs; b += blockDim.x) {
#pragma unroll
        for (int q = 0; q < 4; ++q) st[q] += stats[(size_t)b * 4 + q];
    }
    block_sum_d(st);
    const double mu_s = st[0] / (double)M, mu_r = st[2] / (double)M;
    const double var_s = fmax((st[1] - (double)M * mu_s * mu_s) / (double)(M - 1), 0.0);        // torch.std: unbiased
    const double var_r = fmax((st[3] - (double)M * mu_r * mu_r) / (double)(M - 1), 0.0);
    const double sd_s = sqrt(var_s), sd_r = sqrt(var_r);
    const double is = 1.0 / (sd_s + 1e-6), ir = 1.0 / (sd_r + 1e-6);                             // 3dm:776-777
    const double two_over_M = 2.0 / (double)M;      // (an fp64 division per element in each pass was most of this kernel's time)
    // pass 1: sim loss and the two sums its gradient needs
    double a1[4] = {0, 0, 0, 0};       // sum d^2, sum g, sum g (sim - mu)
    for (long long i = tid; i < M; i += blockDim.x) {
        const double s = (double)sim[i] - mu_s;
        const double d = s * is - ((double)raw[i] - mu_r) * ir;
        const double g = d * two_over_M;
        a1[0] += d * d; a1[1] += g; a1[2] += g * s;
    }
    block_sum_d(a1);
    const double sim_loss = a1[0] / (double)M;                                                   // F.mse_loss 3dm:779
    const double gmean = a1[1] / (double)M;
    const double kcoef = sd_s > 0.0 ? a1[2] * is * is / ((double)(M - 1) * sd_s) : 0.0;
    if (dsim) {
        for (long long i = tid; i < M; i += blockDim.x) {
            const double s = (double)sim[i] - mu_s;
            const double d = s * is - ((double)raw[i] - mu_r) * ir;
            const double g = d * two_over_M;
            dsim[i] = (float)((double)scale * ((g - gmean) * is - s * kcoef));
        }
    }
    // pose_loss (3dm:896-962) per pair + means; correspondence BCE mean
    double pl[4] = {0, 0, 0, 0};       // sum rot, sum trans, sum bce
    for (int b = tid; b < pairs; b += blockDim.x) {
        pl[2] += (double)bce[b];
        if (!R) continue;
        const float *Rb = R + b * 9, *G = gt_pose + b * 16;
        float tr = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) tr = fmaf(Rb[i * 3 + j], G[i * 4 + j], tr);
        const float c = (tr - 1.0f) * 0.5f, cc = fminf(fmaxf(c, -1.0f), 1.0f);
        pl[0] += (double)acosf(cc);
        const float kr = 0.5f * acos_slope(c);
        const float t0 = t[b * 3], t1 = t[b * 3 + 1], t2 = t[b * 3 + 2], g0 = G[3], g1 = G[7], g2 = G[11];
        const float dot = t0 * g0 + t1 * g1 + t2 * g2;
        const float nt = sqrtf(t0 * t0 + t1 * t1 + t2 * t2), ng = sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
        const float cs = dot / (nt * ng), cl = fminf(fmaxf(cs, -1.0f), 1.0f);
        pl[1] += (double)acosf(cl);
        const float kt = acos_slope(cs);
        const float ia = 1.0f / (nt * ng), bb = cs / (nt * nt), sc = scale / (float)pairs;       // .mean() over the pairs 3dm:1118
        if (dR) {
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) dR[b * 9 + i * 3 + j] = sc * kr * G[i * 4 + j];
            dt[b * 3] = sc * kt * (g0 * ia - bb * t0); dt[b * 3 + 1] = sc * kt * (g1 * ia - bb * t1); dt[b * 3 + 2] = sc * kt * (g2 * ia - bb * t2);
        }
    }
    block_sum_d(pl);
    if (tid == 0) {
        const int kk = k < n ? k : n;
        const double corr = pl[2] / ((double)pairs * kk);
        const double rot = R ? pl[0] / pairs : 0.0, tra = R ? pl[1] / pairs : 0.0;
        loss[0] = (float)corr; loss[1] = (float)sim_loss; loss[2] = (float)rot; loss[3] = (float)tra;
        loss[4] = (float)(corr + sim_loss + rot + tra);
        loss[5] = 0.f; loss[6] = 0.f; loss[7] = scale;
    }
}

// ---- pose_loss (3dm:896-962): rotation loss = acos(clamp((trace(R^T R_gt) - 1) / 2, -1, 1)), translation loss =
// acos(clamp(cos(t, t_gt), -1, 1)) per pair, together with their gradients w.r.t. R and t (what autograd would
// return for a unit upstream gradient; clamp passes the gradient on [-1, 1] inclusive like torch.clamp). ----
__global__ void pose_loss_kernel(const float *__restrict__ R, const float *__restrict__ t, const float *__restrict__ gt_pose,
                                 int pairs, float *__restrict__ rot_loss, float *__restrict__ trans_loss,
                                 float *__restrict__ gR, float *__restrict__ gt_) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= pairs) return;
    const float *Rb = R + b * 9, *G = gt_pose + b * 16;
    float tr = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) tr = fmaf(Rb[i * 3 + j], G[i * 4 + j], tr);        // trace(R^T R_gt) = sum_ij R_ij Rgt_ij
    const float c = (tr - 1.0f) * 0.5f;
    const float cc = fminf(fmaxf(c, -1.0f), 1.0f);
    rot_loss[b] = acosf(cc);
    if (gR) {
        const float k = 0.5f * acos_slope(c);   // d acos(c)/dc * dc/dtr
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) gR[b * 9 + i * 3 + j] = k * G[i * 4 + j];
    }
    const float t0 = t[b * 3], t1 = t[b * 3 + 1], t2 = t[b * 3 + 2], g0 = G[3], g1 = G[7], g2 = G[11];
    const float dot = t0 * g0 + t1 * g1 + t2 * g2;
    const float nt = sqrtf(t0 * t0 + t1 * t1 + t2 * t2), ng = sqrtf(g0 * g0 + g1 * g1 + g2 * g2);
    const float cs = dot / (nt * ng);
    const float cl = fminf(fmaxf(cs, -1.0f), 1.0f);
    trans_loss[b] = acosf(cl);
    if (gt_) {
        const float k = acos_slope(cs);
        // d cos / d t = g / (|t||g|) - cos * t / |t|^2
        const float a = 1.0f / (nt * ng), bb = cs / (nt * nt);
        gt_[b * 3] = k * (g0 * a - bb * t0); gt_[b * 3 + 1] = k * (g1 * a - bb * t1); gt_[b * 3 + 2] = k * (g2 * a - bb * t2);
    }
}

static int head_threads(int n) { return n > HD_BIG_N ? HD_THREADS_BIG : (n > 1024 ? 512 : HD_THREADS); }

template <class K>
static int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess)
            return EGSPR_E_LAUNCH;
    }
    return EGSPR_OK;
}

}  // namespace egspr

extern "C" int egspr_kabsch(const float *p, const float *q, const float *w, const float *mask, int pairs, int n,
                            float *R, float *t, float *Hout, void *stream) {
    using namespace egspr;
    if (!p || !q || !w || !R || !t || pairs <= 0 || n < 0) return EGSPR_E_INVALID;
    if (n > HD_MAX_N) return EGSPR_E_UNSUPPORTED;
    const size_t smem = sizeof(float) * (size_t)(n > 0 ? n : 1);
    const bool big_cta = head_threads(n) > 512;        // <= 512 threads: 128 registers per thread, the fp64 SVD stays in registers
    if (int e = big_cta ? ensure_smem(kabsch_kernel<1024>, smem) : ensure_smem(kabsch_kernel<512>, smem)) return e;
    if (big_cta) kabsch_kernel<1024><<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(p, q, w, mask, n, R, t, Hout);
    else kabsch_kernel<512><<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(p, q, w, mask, n, R, t, Hout);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_head_eval(const float *feat_src, const float *feat_tgt, const float *x_src, const float *x_tgt,
                               const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                               const float *x_out_tgt, const float *labels, const float *gt_pose,
                               const float *head_pack, int pairs, int n, int top_k, float *w_out, float *R, float *t,
                               float *Hout, float *loss_parts, void *stream) {
    return egspr_head_eval_ws(feat_src, feat_tgt, x_src, x_tgt, h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, gt_pose,
                              head_pack, pairs, n, top_k, w_out, R, t, Hout, loss_parts, nullptr, 0, stream);
}

extern "C" size_t egspr_head_eval_workspace_bytes(int pairs) {
    return pairs > 0 ? (size_t)pairs * EGSPR_HEAD_MAX_SPLIT * 16 : 0;
}

extern "C" int egspr_head_eval_ws(const float *feat_src, const float *feat_tgt, const float *x_src, const float *x_tgt,
                                  const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                                  const float *x_out_tgt, const float *labels, const float *gt_pose,
                                  const float *head_pack, int pairs, int n, int top_k, float *w_out, float *R, float *t,
                                  float *Hout, float *loss_parts, void *workspace, size_t workspace_bytes, void *stream) {
    using namespace egspr;
    void *pre_scratch = (workspace && workspace_bytes >= egspr_head_eval_workspace_bytes(pairs)) ? workspace : nullptr;
    if (!feat_src || !feat_tgt || !x_src || !x_tgt || !h_out_src || !h_out_tgt || !head_pack || !R || !t || pairs <= 0 ||
        n <= 0 || top_k <= 0)
        return EGSPR_E_INVALID;
    if (loss_parts && (!x_out_src || !x_out_tgt || !labels || !gt_pose)) return EGSPR_E_INVALID;
    if (n > HD_MAX_N && !w_out) return EGSPR_E_WORKSPACE;      // large clouds: the weights row doubles as scratch
    const size_t smem = n > HD_MAX_N ? 0 : sizeof(float) * (size_t)n;
    const bool big_cta = head_threads(n) > 512;        // <= 512 threads: 128 registers per thread, the fp64 SVD stays in registers
    if (int e = big_cta ? ensure_smem(head_eval_kernel<1024>, smem) : ensure_smem(head_eval_kernel<512>, smem)) return e;
    HeadEvalArgs a{feat_src, feat_tgt, x_src, x_tgt, h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, gt_pose,
                   head_pack, n, top_k, w_out, R, t, Hout, loss_parts, 0, nullptr, nullptr};
    // few pairs x large clouds (one CTA per pair would leave most SMs idle while it streams 2 x 2 x n x 128 bytes):
    // the row-wide passes go to `split` CTAs per pair; needs the w_out row as scratch and the caller's partial buffers
    if (pre_scratch && w_out && pairs < 2 * sm_count() && (size_t)n >= 2048) {
        int split = (2 * sm_count() + pairs - 1) / pairs;
        if (split > EGSPR_HEAD_MAX_SPLIT) split = EGSPR_HEAD_MAX_SPLIT;
        if (split > n / 512) split = n / 512;
        if (split >= 2) {
            unsigned long long *pb = reinterpret_cast<unsigned long long *>(pre_scratch);
            float *pl = reinterpret_cast<float *>(pb + (size_t)pairs * split);
            a.split = split; a.pre_best = pb; a.pre_lp = pl;
            head_eval_pre_kernel<<<dim3((unsigned)split, (unsigned)pairs), 256, 0, (cudaStream_t)stream>>>(a, pb, pl);
            EGSPR_CHECK_LAUNCH();
        }
    }
    if (big_cta) head_eval_kernel<1024><<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(a);
    else head_eval_kernel<512><<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(a);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_head_train(const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                                const float *x_out_tgt, const float *labels, const float *gt_pose, int pairs, int n,
                                float *w_out, float *sim_out, float *R, float *t, float *Hout, float *loss_parts,
                                void *stream) {
    using namespace egspr;
    if (!h_out_src || !h_out_tgt || !x_out_src || !x_out_tgt || !labels || !R || !t || pairs <= 0 || n <= 0)
        return EGSPR_E_INVALID;
    if (loss_parts && !gt_pose) return EGSPR_E_INVALID;
    if (n > HD_MAX_N && !w_out) return EGSPR_E_WORKSPACE;
    const size_t smem = n > HD_MAX_N ? 0 : sizeof(float) * (size_t)n;
    if (int e = ensure_smem(head_train_kernel, smem)) return e;
    HeadTrainArgs a{h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, gt_pose, n, w_out, sim_out, R, t, Hout, loss_parts};
    head_train_kernel<<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(a);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

static int head_train_backward_impl(const egspr::HeadTrainBwdArgs &a, int pairs, int n, void *stream) {
    using namespace egspr;
    if (!a.h_out_src || !a.h_out_tgt || !a.x_out_src || !a.x_out_tgt || !a.labels || !a.dR || !a.dt || !a.dh_src || !a.dh_tgt ||
        !a.dx_src || !a.dx_tgt || pairs <= 0 || n <= 0)
        return EGSPR_E_INVALID;
    if (2 * n > HD_MAX_N) return EGSPR_E_UNSUPPORTED;
    const size_t smem = sizeof(float) * 2 * (size_t)n;
    if (int e = ensure_smem(head_train_backward_kernel, smem)) return e;
    head_train_backward_kernel<<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(a);
    EGSPR_CHECK_LAUNCH();
    if (a.top_idx) {
        corr_loss_backward_kernel<<<dim3((unsigned)pairs, CB_SPLIT), CB_THREADS, 0, (cudaStream_t)stream>>>(a);
        EGSPR_CHECK_LAUNCH();
    }
    return EGSPR_OK;
}

extern "C" int egspr_head_train_backward(const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                                         const float *x_out_tgt, const float *labels, const float *dR, const float *dt,
                                         const float *dsim, int pairs, int n, float *dh_src, float *dh_tgt,
                                         float *dx_src, float *dx_tgt, void *stream) {
    egspr::HeadTrainBwdArgs a{h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, dR, dt, dsim, n, dh_src, dh_tgt, dx_src, dx_tgt,
                              nullptr, nullptr, nullptr, 0, pairs, nullptr};
    return head_train_backward_impl(a, pairs, n, stream);
}

extern "C" int egspr_head_train_loss_backward(const float *h_out_src, const float *h_out_tgt, const float *x_out_src,
                                              const float *x_out_tgt, const float *labels, const float *dR, const float *dt,
                                              const float *dsim, const int32_t *top_idx, const float *head_pack,
                                              const float *loss, int pairs, int n, int top_k, float *dh_src, float *dh_tgt,
                                              float *dx_src, float *dx_tgt, float *head_grad_pack, void *stream) {
    if (!top_idx || !head_pack || !loss || !head_grad_pack || top_k <= 0) return EGSPR_E_INVALID;
    egspr::HeadTrainBwdArgs a{h_out_src, h_out_tgt, x_out_src, x_out_tgt, labels, dR, dt, dsim, n, dh_src, dh_tgt, dx_src, dx_tgt,
                              top_idx, head_pack, head_grad_pack, top_k, pairs, loss + 7};
    return head_train_backward_impl(a, pairs, n, stream);
}

extern "C" int egspr_train_loss_forward(const float *h_out_src, const float *h_out_tgt, const float *feat_src,
                                        const float *feat_tgt, const float *sim, const float *labels, const float *head_pack,
                                        int pairs, int n, int top_k, int32_t *top_idx, float *scores, float *raw,
                                        double *stats, float *bce, void *stream) {
    using namespace egspr;
    if (!h_out_src || !h_out_tgt || !feat_src || !feat_tgt || !labels || !head_pack || !top_idx || !scores || !raw ||
        !stats || !bce || pairs <= 0 || n <= 0 || top_k <= 0)
        return EGSPR_E_INVALID;
    const size_t smem = sizeof(float) * ((size_t)n + top_k + HEAD_PACK);
    if (smem > 200 * 1024) return EGSPR_E_UNSUPPORTED;
    if (int e = ensure_smem(train_loss_forward_kernel, smem)) return e;
    TrainLossArgs a{h_out_src, h_out_tgt, feat_src, feat_tgt, sim, labels, head_pack, n, top_k, top_idx, scores, raw, stats, bce};
    train_loss_forward_kernel<<<pairs, head_threads(n), smem, (cudaStream_t)stream>>>(a);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_train_loss_finalize(const float *sim, const float *raw, const double *stats, const float *bce, int pairs,
                                         int n, int top_k, const float *R, const float *t, const float *gt_pose, float scale,
                                         float *loss, float *dsim, float *dR, float *dt, void *stream) {
    using namespace egspr;
    if (!sim || !raw || !stats || !bce || !loss || pairs <= 0 || n <= 0 || top_k <= 0) return EGSPR_E_INVALID;
    if (R && (!t || !gt_pose)) return EGSPR_E_INVALID;
    if (dR && (!dt || !R)) return EGSPR_E_INVALID;
    train_loss_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(sim, raw, stats, bce, pairs, n, top_k, R, t, gt_pose, scale, loss,
                                                                     dsim, dR, dt);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_pose_loss(const float *R, const float *t, const float *gt_pose, int pairs, float *rot_loss,
                               float *trans_loss, float *grad_R, float *grad_t, void *stream) {
    using namespace egspr;
    if (!R || !t || !gt_pose || !rot_loss || !trans_loss || pairs <= 0) return EGSPR_E_INVALID;
    pose_loss_kernel<<<(pairs + 127) / 128, 128, 0, (cudaStream_t)stream>>>(R, t, gt_pose, pairs, rot_loss, trans_loss, grad_R, grad_t);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_version(void) { return 102; }

extern "C" const char *egspr_error_string(int code) {
    switch (code) {
        case EGSPR_OK: return "ok";
        case EGSPR_E_INVALID: return "invalid argument (null pointer, non-positive size, aliasing outputs)";
        case EGSPR_E_UNSUPPORTED: return "unsupported shape for the compiled kernels";
        case EGSPR_E_WORKSPACE: return "workspace too small";
        case EGSPR_E_LAUNCH: return "CUDA launch failed";
        default: return "unknown egspr error";
    }
}

extern "C" int egspr_pose_metrics(const float *R, const float *t, const float *gt_pose, const float *src_pts,
                                  const float *tgt_pts, int pairs, int n, double tau, double *out, void *stream) {
    using namespace egspr;
    if (!R || !t || !gt_pose || !src_pts || !tgt_pts || !out || pairs <= 0 || n < 0) return EGSPR_E_INVALID;
    pose_metrics_kernel<<<pairs, 256, 0, (cudaStream_t)stream>>>(R, t, gt_pose, src_pts, tgt_pts, n, tau, out);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}
