// k-NN graph build: replaces torch_cluster.knn_graph(x, k, loop=True) at the reference call sites
// src/3dmatch_train_egnn_with_batch.py:1005-1006 / src/eval_egnn_metrics.py:1156-1157, for ALL
// clouds of a batch in one launch (the reference loops over 2*B clouds in Python).
//
// Spec (== oracle/knn_oracle.c): k smallest candidates under the total order (d2, index),
// d2 = fma(dz,dz,fma(dy,dy,dx*dx)) in fp32, self included, nearest first, never-filled slots -1,
// candidates with d2 >= 1e10 are never selected (torch_cluster's best_dist init).
//
// Mapping: one warp per query.  Candidate tiles of the query's cloud are staged in shared memory
// (SoA, coalesced loads); each lane scores one candidate per step; the running top-k lives one
// entry per lane (k <= 32), kept sorted; candidates beating the current k-th distance are
// inserted with a ballot + shuffle-up (no local memory, no divergence across queries).
#include <cstdlib>

#include "egspr_common.cuh"

namespace egspr {

constexpr int KNN_WARPS = 8;          // queries per CTA
constexpr int KNN_TILE = 2048;        // candidates staged per shared-memory tile

__global__ void __launch_bounds__(KNN_WARPS * 32) knn_warp_select_kernel(
    const float *__restrict__ x, int n, int k, int32_t *__restrict__ nbr) {
    __shared__ float sx[KNN_TILE], sy[KNN_TILE], sz[KNN_TILE];
    const int cloud = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int qi = blockIdx.x * KNN_WARPS + warp;
    const float *xc = x + (size_t)cloud * n * 3;
    const bool active = qi < n;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (active) { qx = __ldg(xc + 3 * qi); qy = __ldg(xc + 3 * qi + 1); qz = __ldg(xc + 3 * qi + 2); }
    float bd = 1e10f;   // lane s < k: s-th best distance so far
    int bi = -1;
    for (int base = 0; base < n; base += KNN_TILE) {
        const int cnt = min(KNN_TILE, n - base);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += KNN_WARPS * 32) {
            const float *p = xc + 3 * (size_t)(base + i);
            sx[i] = __ldg(p); sy[i] = __ldg(p + 1); sz[i] = __ldg(p + 2);
        }
        __syncthreads();
        if (!active) continue;
        for (int j0 = 0; j0 < cnt; j0 += 32) {
            const int j = j0 + lane;
            float d = 3e38f;
            if (j < cnt) {
                const float dx = sx[j] - qx, dy = sy[j] - qy, dz = sz[j] - qz;
                d = __fmaf_rn(dz, dz, __fmaf_rn(dy, dy, __fmul_rn(dx, dx)));
            }
            float thr = __shfl_sync(0xffffffffu, bd, k - 1);
            unsigned m = __ballot_sync(0xffffffffu, d < thr);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float dd = __shfl_sync(0xffffffffu, d, src);
                if (dd < thr) {   // warp-uniform: thr may have dropped since the ballot
                    // position = number of kept entries <= dd (the newcomer has the highest index so far)
                    const int pos = __popc(__ballot_sync(0xffffffffu, lane < k && bd <= dd));
                    const float ubd = __shfl_up_sync(0xffffffffu, bd, 1);
                    const int ubi = __shfl_up_sync(0xffffffffu, bi, 1);
                    if (lane > pos) { bd = ubd; bi = ubi; }
                    if (lane == pos) { bd = dd; bi = base + j0 + src; }
                    thr = __shfl_sync(0xffffffffu, bd, k - 1);
                }
            }
        }
    }
    if (active && lane < k) nbr[((size_t)cloud * n + qi) * k + lane] = bi;
}

// ================================================================================================
// Cell-grid k-NN (exact): same result as the brute-force scan because the selection is defined by
// the TOTAL order (d2, index) -- visiting order does not matter.
//   build : one CTA per cloud: bounding box -> anisotropic cell grid (~4 points per cell, <= 4096
//           cells) -> counting sort of the points by cell (x,y,z,index packed in a float4)
//   query : one warp per query, in cell-sorted order (neighbouring warps touch neighbouring cells):
//           scan the cubic shell of cells at Chebyshev radius L = 0,1,2,... around the query's cell
//           (contiguous x-runs of cells = contiguous candidate ranges, one candidate per lane, the
//           same ballot/shuffle insertion as above but with the lexicographic (d2, index) compare);
//           stop when the k-th best distance is strictly inside the scanned block (margin test with
//           1e-4 relative slack) or the block covers the whole grid.
// ================================================================================================
constexpr int GRID_MAXC = 32768;      // upper bound of cells per cloud (shared-memory counters of the build CTA)
// cell capacity reserved per cloud: the grid aims at ~2 points per cell (n/2 cells; the per-axis rounding can
// overshoot, the build shrinks the grid until it fits)
static inline int grid_capacity(int n) { return n < 1024 ? 1024 : (n > GRID_MAXC ? GRID_MAXC : n); }
constexpr int GRID_BUILD_THREADS = 512;

struct GridParams {        // per cloud, 16 floats
    float mn[3];
    float inv[3];          // 1 / cell size
    float cs[3];           // cell size
    int dims[3];
    float slack;           // absolute slack of the stopping rule: rounding of the point -> cell assignment
    int pad[3];
};

__device__ GridParams grid_params_from_bbox(const float (&mn)[3], const float (&ext)[3], int n, int maxc, float pts_per_cell) {
    GridParams gp;
    const float emax = fmaxf(fmaxf(ext[0], ext[1]), fmaxf(ext[2], 1e-30f));
    // cells wanted: n / pts_per_cell; degenerate (flat) axes get a single cell
    float target = fmaxf(1.0f, (float)n / pts_per_cell);
    if (target > (float)maxc) target = (float)maxc;
    float e[3]; int live = 0; float vol = 1.0f;
    for (int a = 0; a < 3; ++a) { e[a] = ext[a]; if (e[a] > 1e-4f * emax) { ++live; vol *= e[a]; } }
    float c = live ? powf(vol / target, 1.0f / (float)live) : 1.0f;
    int dims[3]; long long total = 1;
    for (int a = 0; a < 3; ++a) {
        int d = (e[a] > 1e-4f * emax) ? (int)(e[a] / c) + 1 : 1;
        if (d > 64) d = 64;
        if (d < 1) d = 1;
        dims[a] = d; total *= d;
    }
    while (total > maxc) {           // shrink the largest dimension until the grid fits
        int am = 0;
        for (int a = 1; a < 3; ++a) if (dims[a] > dims[am]) am = a;
        total /= dims[am]; dims[am] -= 1; total *= dims[am];
    }
    for (int a = 0; a < 3; ++a) {
        gp.mn[a] = mn[a]; gp.dims[a] = dims[a];
        const float cs = (ext[a] > 0.f) ? ext[a] / (float)dims[a] : 1.0f;
        gp.cs[a] = cs; gp.inv[a] = 1.0f / cs;
    }
    // a point may be assigned to the cell next to its geometric one when (p - mn) * inv rounds across an
    // integer: the error is a few ulp of the coordinate magnitude
    gp.slack = 8.0f * 1.1920929e-7f * (fmaxf(fmaxf(fabsf(mn[0]), fabsf(mn[1])), fabsf(mn[2])) + emax);
    gp.pad[0] = gp.pad[1] = gp.pad[2] = 0;
    return gp;
}

__global__ void __launch_bounds__(GRID_BUILD_THREADS) knn_grid_build_kernel(
    const float *__restrict__ x, int n, float4 *__restrict__ sorted, int *__restrict__ cell_start,
    GridParams *__restrict__ params, int maxc, float pts_per_cell) {
    extern __shared__ int cnt[];          // [maxc + 1]
    __shared__ float red[6][GRID_BUILD_THREADS / 32];
    __shared__ GridParams gp;
    __shared__ int wtot[GRID_BUILD_THREADS / 32];
    __shared__ int carry_s;
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *xc = x + (size_t)cloud * n * 3;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = tid; i < n; i += GRID_BUILD_THREADS) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = __ldg(xc + 3 * i + a); lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { red[a][warp] = lo[a]; red[3 + a][warp] = hi[a]; }
    }
    for (int i = tid; i <= maxc; i += GRID_BUILD_THREADS) cnt[i] = 0;
    __syncthreads();
    if (tid == 0) {
        float mn[3], ext[3];
        for (int a = 0; a < 3; ++a) {
            float l = red[a][0], h = red[3 + a][0];
            for (int w = 1; w < GRID_BUILD_THREADS / 32; ++w) { l = fminf(l, red[a][w]); h = fmaxf(h, red[3 + a][w]); }
            mn[a] = l; ext[a] = h - l;
        }
        gp = grid_params_from_bbox(mn, ext, n, maxc, pts_per_cell);
        params[cloud] = gp;
    }
    __syncthreads();
    const int nx = gp.dims[0], ny = gp.dims[1], nz = gp.dims[2];
    const int ncell = nx * ny * nz;
    auto cell_of = [&](float px, float py, float pz) {
        int cx = min(nx - 1, max(0, (int)((px - gp.mn[0]) * gp.inv[0])));
        int cy = min(ny - 1, max(0, (int)((py - gp.mn[1]) * gp.inv[1])));
        int cz = min(nz - 1, max(0, (int)((pz - gp.mn[2]) * gp.inv[2])));
        return (cz * ny + cy) * nx + cx;
    };
    for (int i = tid; i < n; i += GRID_BUILD_THREADS)
        atomicAdd(&cnt[cell_of(__ldg(xc + 3 * i), __ldg(xc + 3 * i + 1), __ldg(xc + 3 * i + 2))], 1);
    __syncthreads();
    // exclusive scan of cnt[0..ncell) -> cell_start (global) and cursor (cnt reused)
    if (tid == 0) carry_s = 0;
    __syncthreads();
    int *cs_out = cell_start + (size_t)cloud * (maxc + 1);
    for (int base = 0; base < ncell; base += GRID_BUILD_THREADS) {
        const int i = base + tid;
        const int v = i < ncell ? cnt[i] : 0;
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
        if (lane == 31) wtot[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            int t = lane < GRID_BUILD_THREADS / 32 ? wtot[lane] : 0, u = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(0xffffffffu, u, o); if (lane >= o) u += w; }
            if (lane < GRID_BUILD_THREADS / 32) wtot[lane] = u - t;
        }
        __syncthreads();
        const int excl = carry_s + wtot[warp] + sc - v;
        if (i < ncell) { cs_out[i] = excl; cnt[i] = excl; }
        __syncthreads();
        if (tid == GRID_BUILD_THREADS - 1) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) cs_out[ncell] = n;
    float4 *so = sorted + (size_t)cloud * n;
    for (int i = tid; i < n; i += GRID_BUILD_THREADS) {
        const float px = __ldg(xc + 3 * i), py = __ldg(xc + 3 * i + 1), pz = __ldg(xc + 3 * i + 2);
        const int pos = atomicAdd(&cnt[cell_of(px, py, pz)], 1);
        so[pos] = make_float4(px, py, pz, __int_as_float(i));
    }
}

// ---- split build for few, large clouds (one CTA per cloud would serialise a 131k-point counting sort on one SM):
// bounding box (ordered-int atomics) -> per-cell counts (global atomics) -> scan (one CTA per cloud) -> scatter.  The order of
// the points inside a cell differs from run to run; the query's selection is a total order, so its result does not. ----
__device__ __forceinline__ unsigned f2ord(float f) { const unsigned u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
__device__ __forceinline__ float ord2f(unsigned u) { return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u); }

__global__ void __launch_bounds__(256) knn_grid_bbox_kernel(const float *__restrict__ x, int n, unsigned *__restrict__ bbox) {
    const int cloud = blockIdx.y, lane = threadIdx.x & 31;
    const float *xc = x + (size_t)cloud * n * 3;
    float lo[3] = {3e38f, 3e38f, 3e38f}, hi[3] = {-3e38f, -3e38f, -3e38f};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { const float v = __ldg(xc + 3 * i + a); lo[a] = fminf(lo[a], v); hi[a] = fmaxf(hi[a], v); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
            hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
        }
        if (lane == 0) { atomicMin(bbox + cloud * 8 + a, f2ord(lo[a])); atomicMax(bbox + cloud * 8 + 4 + a, f2ord(hi[a])); }
    }
}

__device__ __forceinline__ GridParams grid_params_of_cloud(const unsigned *__restrict__ bbox, int cloud, int n, int maxc, float ppc) {
    float mn[3], ext[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { mn[a] = ord2f(bbox[cloud * 8 + a]); ext[a] = ord2f(bbox[cloud * 8 + 4 + a]) - mn[a]; }
    return grid_params_from_bbox(mn, ext, n, maxc, ppc);
}
__device__ __forceinline__ int grid_cell_of(const GridParams &gp, float px, float py, float pz) {
    const int nx = gp.dims[0], ny = gp.dims[1], nz = gp.dims[2];
    const int cx = min(nx - 1, max(0, (int)((px - gp.mn[0]) * gp.inv[0])));
    const int cy = min(ny - 1, max(0, (int)((py - gp.mn[1]) * gp.inv[1])));
    const int cz = min(nz - 1, max(0, (int)((pz - gp.mn[2]) * gp.inv[2])));
    return (cz * ny + cy) * nx + cx;
}

// pass 0: counts into cell_start (zeroed);  pass 1: scatter through `cursor` (= exclusive offsets)
template <int PASS>
__global__ void __launch_bounds__(256) knn_grid_count_scatter_kernel(const float *__restrict__ x, int n, const unsigned *__restrict__ bbox,
                                                                     int *__restrict__ cell_start, int *__restrict__ cursor,
                                                                     float4 *__restrict__ sorted, GridParams *__restrict__ params,
                                                                     int maxc, float ppc) {
    __shared__ GridParams gp;
    const int cloud = blockIdx.y;
    if (threadIdx.x == 0) {
        gp = grid_params_of_cloud(bbox, cloud, n, maxc, ppc);
        if (PASS == 0 && blockIdx.x == 0) params[cloud] = gp;
    }
    __syncthreads();
    const float *xc = x + (size_t)cloud * n * 3;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float px = __ldg(xc + 3 * i), py = __ldg(xc + 3 * i + 1), pz = __ldg(xc + 3 * i + 2);
        const int c = grid_cell_of(gp, px, py, pz);
        if (PASS == 0) atomicAdd(cell_start + (size_t)cloud * (maxc + 1) + c, 1);
        else sorted[(size_t)cloud * n + atomicAdd(cursor + (size_t)cloud * (maxc + 1) + c, 1)] = make_float4(px, py, pz, __int_as_float(i));
    }
}

// exclusive scan of the cell counts of one cloud (<= GRID_MAXC cells) -> cell_start and the scatter cursors
__global__ void __launch_bounds__(1024) knn_grid_scan_kernel(const GridParams *__restrict__ params, int n, int *__restrict__ cell_start,
                                                             int *__restrict__ cursor, int maxc) {
    __shared__ int wtot[32];
    __shared__ int carry_s;
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const GridParams gp = params[cloud];
    const int ncell = gp.dims[0] * gp.dims[1] * gp.dims[2];
    int *cs = cell_start + (size_t)cloud * (maxc + 1), *cu = cursor + (size_t)cloud * (maxc + 1);
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < ncell; base += 1024) {
        const int i = base + tid;
        const int v = i < ncell ? cs[i] : 0;
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
        if (lane == 31) wtot[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            int t = wtot[lane], u = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(0xffffffffu, u, o); if (lane >= o) u += w; }
            wtot[lane] = u - t;
        }
        __syncthreads();
        const int excl = carry_s + wtot[warp] + sc - v;
        if (i < ncell) { cs[i] = excl; cu[i] = excl; }
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) cs[ncell] = n;
}

constexpr int GQ_WARPS = 8;

// ---- sub-warp variant (k <= 16): 32/W queries per warp, W lanes each ---------------------------------
// The per-iteration instruction cost of the flat kernel (segment search, candidate scoring, the
// serial insertion loop) is paid once per warp, so sharing a warp between 32/W neighbouring queries
// (cell-sorted order: same or adjacent cells, similar trip counts) divides the cost per query.  The
// sorted top-k list of a query lives E = 16/W entries per lane (position = lane_in_group * E + j).
// Identical selection rule: total order (d2, index), same stopping test.
template <int W, int E>
__global__ void __launch_bounds__(GQ_WARPS * 32) knn_grid_query_sub_kernel(
    const float4 *__restrict__ sorted, const int *__restrict__ cell_start, const GridParams *__restrict__ params,
    int n, int k, int32_t *__restrict__ nbr, int maxc) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int l0 = 1;            // radius of the first block of cells          } re-swept on the final kernel: ppc 2,
    constexpr int merge_min = 5;     // batches with >= 5 passing candidates are bitonic-merged } l0 1, merge_min 3..5 are the optimum
    // E = list entries per lane: the list holds W * E >= k entries
    constexpr int QPW = 32 / W;                    // queries per warp
    constexpr unsigned GMASK = (W == 32) ? 0xffffffffu : ((1u << W) - 1u);
    const int cloud = blockIdx.y;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / W, sl = lane % W, sh = W * sub;
    const int qs = (blockIdx.x * GQ_WARPS + warp) * QPW + sub;      // query = qs-th point in cell-sorted order
    const bool valid = qs < n;
    const float4 *so = sorted + (size_t)cloud * n;
    const int *cs = cell_start + (size_t)cloud * (maxc + 1);
    const GridParams gp = params[cloud];
    const float4 q = __ldg(so + (valid ? qs : n - 1));
    const int qi = __float_as_int(q.w);
    const int nx = gp.dims[0], ny = gp.dims[1], nz = gp.dims[2];
    const int cx = min(nx - 1, max(0, (int)((q.x - gp.mn[0]) * gp.inv[0])));
    const int cy = min(ny - 1, max(0, (int)((q.y - gp.mn[1]) * gp.inv[1])));
    const int cz = min(nz - 1, max(0, (int)((q.z - gp.mn[2]) * gp.inv[2])));
    // list entries and candidates as ONE 64-bit key (bits of d2 << 32 | index): d2 >= +0, so the unsigned order of the keys
    // is the total order (d2, index) and every compare / select of the networks below is a 64-bit integer one.  An empty
    // slot is (1e10, 0): no candidate with d2 >= 1e10 sorts before it (torch_cluster's best_dist), every other one does.
    typedef unsigned long long u64;
    constexpr u64 KEY_EMPTY = (u64)0x501502f9u << 32, KEY_NONE = ~0ull;
    u64 bk[E];
#pragma unroll
    for (int j = 0; j < E; ++j) bk[j] = KEY_EMPTY;
    u64 thr = KEY_EMPTY;                                   // key of the k-th best; < KEY_EMPTY once the list is full
    const int kl = (k - 1) / E, kj = (k - 1) % E;          // lane / slot of the k-th entry
    const int Lmax = max(max(max(max(cx, nx - 1 - cx), max(cy, ny - 1 - cy)), max(cz, nz - 1 - cz)), l0);
    bool done = !valid;
    const int w0 = 2 * l0 + 1;         // the first step scans the whole (2 l0 + 1)^3 block (shells 0 .. l0 together)
    for (int L = l0; !__all_sync(FULL, done); ++L) {
        const int w = 2 * L - 1;
        const unsigned winv = 65536u / (unsigned)w + 1u;      // rr / w == (rr * winv) >> 16 while rr * w < 65536 (L <= 20)
        const int nseg = done ? 0 : ((L == l0) ? w0 * w0 : 8 * L + 2 * w * w);
        for (int s0 = 0; __any_sync(FULL, s0 < nseg); s0 += W) {
            const int s = s0 + sl;
            int start = 0, len = 0;
            if (s < nseg) {
                int dz, dy, x0, x1;
                if (L == l0) {
                    // l0 = 1: the query's own cell row first, then the four face rows, then the corner rows -- the k-th best
                    // falls fastest, so later batches bring fewer candidates that still pass it
                    const int so = (l0 == 1) ? (int)((0x862075314ull >> (4 * s)) & 15ull) : s;
                    dz = so / w0 - l0; dy = so % w0 - l0; x0 = cx - l0; x1 = cx + l0;
                } else if (s < 8 * L) {
                    x0 = cx - L; x1 = cx + L;
                    if (s < 2 * L + 1) { dz = -L; dy = s - L; }
                    else if (s < 2 * (2 * L + 1)) { dz = L; dy = s - (2 * L + 1) - L; }
                    else { const int t = s - 2 * (2 * L + 1); dy = (t >= w) ? L : -L; dz = (t >= w ? t - w : t) - (L - 1); }
                } else {
                    const int t = s - 8 * L, rr = t >> 1;
                    const int qd = (L <= 20) ? (int)(((unsigned)rr * winv) >> 16) : rr / w;
                    dz = qd - (L - 1); dy = (rr - qd * w) - (L - 1);
                    x0 = x1 = (t & 1) ? cx + L : cx - L;
                }
                const int z = cz + dz, y = cy + dy;
                x0 = max(x0, 0); x1 = min(x1, nx - 1);
                if (L > l0 && thr < KEY_EMPTY) {
                    const float thr_d = __uint_as_float((unsigned)(thr >> 32));
                    // shells: with the list full, a cell row whose nearest point is farther than the k-th best cannot
                    // contribute (the selection is a total order, so skipping candidates that lose anyway changes
                    // nothing); the x range shrinks to the cells the ball of radius sqrt(thr) reaches.  Conservative by
                    // the grid's rounding slack and 1e-4 relative, so ties at the threshold are still visited.
                    float gy = dy > 0 ? (gp.mn[1] + (float)y * gp.cs[1]) - q.y : (dy < 0 ? q.y - (gp.mn[1] + (float)(y + 1) * gp.cs[1]) : 0.f);
                    float gz = dz > 0 ? (gp.mn[2] + (float)z * gp.cs[2]) - q.z : (dz < 0 ? q.z - (gp.mn[2] + (float)(z + 1) * gp.cs[2]) : 0.f);
                    gy = fmaxf(gy - gp.slack, 0.f) * 0.9999f; gz = fmaxf(gz - gp.slack, 0.f) * 0.9999f;
                    const float rem = thr_d * 1.0001f - (gy * gy + gz * gz);
                    if (rem < 0.f) x1 = x0 - 1;
                    else {
                        const float rx = sqrtf(rem) * 1.0001f + gp.slack;
                        x0 = max(x0, (int)fmaxf((q.x - rx - gp.mn[0]) * gp.inv[0], 0.f));
                        x1 = min(x1, (int)fminf(fmaxf((q.x + rx - gp.mn[0]) * gp.inv[0], -1.f), 1e6f));
                    }
                }
                if (z >= 0 && z < nz && y >= 0 && y < ny && x0 <= x1) {
                    const int rowbase = (z * ny + y) * nx;
                    start = __ldg(cs + rowbase + x0);
                    len = __ldg(cs + rowbase + x1 + 1) - start;
                }
            }
            int incl = len;
#pragma unroll
            for (int o = 1; o < W; o <<= 1) { const int t = __shfl_up_sync(FULL, incl, o, W); if (sl >= o) incl += t; }
            const int excl = incl - len;
            const int T = __shfl_sync(FULL, incl, W - 1, W);
            for (int t0 = 0; __any_sync(FULL, t0 < T); t0 += W) {
                const int t = t0 + sl;
                int lo = 0;
#pragma unroll
                for (int h = W / 2; h > 0; h >>= 1) {
                    const int v = __shfl_sync(FULL, excl, lo + h, W);
                    if (v <= t) lo += h;
                }
                const int sstart = __shfl_sync(FULL, start, lo, W), sbase = __shfl_sync(FULL, excl, lo, W);
                u64 ck = KEY_NONE;
                if (t < T) {
                    const float4 c = __ldg(so + sstart + (t - sbase));
                    const float dx = c.x - q.x, dy2 = c.y - q.y, dz2 = c.z - q.z;
                    const float d = __fmaf_rn(dz2, dz2, __fmaf_rn(dy2, dy2, __fmul_rn(dx, dx)));
                    ck = ((u64)__float_as_uint(d) << 32) | (unsigned)__float_as_int(c.w);
                }
                unsigned m = (__ballot_sync(FULL, ck < thr) >> sh) & GMASK;
                if constexpr (E == 1) {
                    // many candidates beat the k-th best (the first batches of a query): sort the batch with a
                    // bitonic network and merge it into the list in one go instead of inserting one by one
                    if (__any_sync(FULL, __popc(m) >= merge_min)) {
                        u64 sk = ((m >> sl) & 1u) ? ck : KEY_NONE;        // non-passers cannot enter the list
#pragma unroll
                        for (int size = 2; size <= W; size <<= 1) {
#pragma unroll
                            for (int stride = size >> 1; stride > 0; stride >>= 1) {
                                const u64 pk = __shfl_xor_sync(FULL, sk, stride, W);
                                const bool plt = pk < sk;                                     // partner sorts first
                                const bool up = ((sl & size) == 0);                           // ascending block
                                const bool lower = ((sl & stride) == 0);
                                if (plt == (up == lower)) sk = pk;                            // keep min in the lower lane of an ascending pair
                            }
                        }
                        // lowest W of (list, batch): elementwise min against the reversed batch is bitonic -> log2(W) merge stages
                        {
                            const u64 rk = __shfl_sync(FULL, sk, W - 1 - sl, W);
                            if (rk < bk[0]) bk[0] = rk;
                        }
#pragma unroll
                        for (int stride = W / 2; stride > 0; stride >>= 1) {
                            const u64 pk = __shfl_xor_sync(FULL, bk[0], stride, W);
                            const bool plt = pk < bk[0];
                            const bool lower = ((sl & stride) == 0);
                            if (plt == lower) bk[0] = pk;
                        }
                        thr = __shfl_sync(FULL, bk[0], k - 1, W);
                        m = 0;
                    }
                }
                // serial insertion of the candidates that beat the current k-th best, each group on its own list
                while (__any_sync(FULL, m != 0)) {
                    const bool act = m != 0;
                    const int src = act ? __ffs(m) - 1 : 0;
                    m &= m - 1;
                    const u64 cand = __shfl_sync(FULL, ck, src, W);
                    const bool ins = act && cand < thr;                  // thr may have dropped since the ballot
                    // position = number of kept entries that sort before the candidate (empty slots never do)
                    int pos = 0;
#pragma unroll
                    for (int j = 0; j < E; ++j) {
                        const bool before = (sl * E + j) < k && bk[j] < cand;
                        pos += __popc((__ballot_sync(FULL, before) >> sh) & GMASK);
                    }
                    const u64 ubk = __shfl_up_sync(FULL, bk[E - 1], 1, W);
                    if (ins) {
#pragma unroll
                        for (int j = E - 1; j >= 0; --j) {
                            const int p = sl * E + j;
                            const u64 pk = (j == 0) ? ubk : bk[j - 1];
                            if (p > pos) bk[j] = pk;
                            else if (p == pos) bk[j] = cand;
                        }
                    }
                    u64 tk = bk[0];
#pragma unroll
                    for (int j = 1; j < E; ++j) if (kj == j) tk = bk[j];
                    thr = __shfl_sync(FULL, tk, kl, W);
                }
            }
        }
        if (!done) {
            float margin = 3e38f;
            if (cx - L > 0) margin = fminf(margin, q.x - (gp.mn[0] + (float)(cx - L) * gp.cs[0]));
            if (cx + L < nx - 1) margin = fminf(margin, (gp.mn[0] + (float)(cx + L + 1) * gp.cs[0]) - q.x);
            if (cy - L > 0) margin = fminf(margin, q.y - (gp.mn[1] + (float)(cy - L) * gp.cs[1]));
            if (cy + L < ny - 1) margin = fminf(margin, (gp.mn[1] + (float)(cy + L + 1) * gp.cs[1]) - q.y);
            if (cz - L > 0) margin = fminf(margin, q.z - (gp.mn[2] + (float)(cz - L) * gp.cs[2]));
            if (cz + L < nz - 1) margin = fminf(margin, (gp.mn[2] + (float)(cz + L + 1) * gp.cs[2]) - q.z);
            const float ms = margin * 0.9999f - gp.slack;
            if (margin > 1e37f || L >= Lmax) done = true;                     // block covers the whole grid
            else if (thr < KEY_EMPTY && ms > 0.f && __uint_as_float((unsigned)(thr >> 32)) < ms * ms) done = true;  // k-th best strictly inside the block
        }
    }
    if (valid) {
#pragma unroll
        for (int j = 0; j < E; ++j)
            if (sl * E + j < k) nbr[((size_t)cloud * n + qi) * k + sl * E + j] = bk[j] < KEY_EMPTY ? (int)(unsigned)bk[j] : -1;
    }
}

template <int W, int E>
static void launch_knn_sub(const float4 *sorted, const int *cell_start, const GridParams *params, int clouds, int n, int k,
                           int32_t *nbr, int maxc, cudaStream_t st) {
    constexpr int QPB = GQ_WARPS * (32 / W);
    dim3 grid((n + QPB - 1) / QPB, clouds);
    knn_grid_query_sub_kernel<W, E><<<grid, GQ_WARPS * 32, 0, st>>>(sorted, cell_start, params, n, k, nbr, maxc);
}

__global__ void nbr_to_edges_kernel(const int32_t *__restrict__ nbr, int n, int k,
                                    int64_t *__restrict__ edges, int64_t per_cloud) {
    // edges[cloud][0][e] = nbr (row, neighbour); edges[cloud][1][e] = e / k (col, centre)
    const int cloud = blockIdx.y;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < per_cloud;
         e += (int64_t)gridDim.x * blockDim.x) {
        edges[(cloud * 2 + 0) * per_cloud + e] = nbr[cloud * per_cloud + e];
        edges[(cloud * 2 + 1) * per_cloud + e] = e / k;
    }
}

}  // namespace egspr

extern "C" size_t egspr_knn_workspace_bytes(int clouds, int n) {
    using namespace egspr;
    // per cloud: cell-sorted points | cell offsets | scatter cursors (split build) | grid parameters | bounding box (8 uints)
    return (size_t)clouds * ((size_t)n * sizeof(float4) + 2 * (size_t)(grid_capacity(n) + 1) * sizeof(int) + sizeof(GridParams) + 32) + 256;
}

extern "C" int egspr_knn_build(const float *x, int clouds, int n, int k, int32_t *nbr, void *workspace,
                               size_t workspace_bytes, void *stream) {
    using namespace egspr;
    if (!x || !nbr || clouds <= 0 || n <= 0) return EGSPR_E_INVALID;
    if (k <= 0 || k > EGSPR_MAX_K) return EGSPR_E_UNSUPPORTED;
    if (clouds > 65535) return EGSPR_E_UNSUPPORTED;
    cudaStream_t st = (cudaStream_t)stream;
    if (!workspace) {                       // no scratch: brute-force scan
        dim3 grid((n + KNN_WARPS - 1) / KNN_WARPS, clouds);
        knn_warp_select_kernel<<<grid, KNN_WARPS * 32, 0, st>>>(x, n, k, nbr);
        EGSPR_CHECK_LAUNCH();
        return EGSPR_OK;
    }
    if (workspace_bytes < egspr_knn_workspace_bytes(clouds, n)) return EGSPR_E_WORKSPACE;
    uint8_t *w = (uint8_t *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    float4 *sorted = (float4 *)w;
    int *cell_start = (int *)(w + (size_t)clouds * n * sizeof(float4));
    const int maxc = grid_capacity(n);
    int *cursor = cell_start + (size_t)clouds * (maxc + 1);
    GridParams *params = (GridParams *)(cursor + (size_t)clouds * (maxc + 1));
    unsigned *bbox = (unsigned *)(params + clouds);
    constexpr float ppc = 2.0f;      // target points per cell (measured optimum at 2048-point clouds, k = 16)
    if (n >= 8192 && clouds <= 8) {
        // few large clouds: the build split over many CTAs per cloud (measured on B200: 2 x 131k points 0.59 -> 0.46 ms for
        // build + query; from 16 clouds up the one-CTA-per-cloud kernel is faster than the 4 launches + memsets)
        int split = (4 * sm_count() + clouds - 1) / clouds;
        if (split > (n + 1023) / 1024) split = (n + 1023) / 1024;
        if (cudaMemsetAsync(bbox, 0xff, (size_t)clouds * 32, st) != cudaSuccess) return EGSPR_E_LAUNCH;          // min slots: 0xffffffff
        for (int c = 0; c < clouds; ++c)
            if (cudaMemsetAsync(bbox + c * 8 + 4, 0, 16, st) != cudaSuccess) return EGSPR_E_LAUNCH;              // max slots: 0
        if (cudaMemsetAsync(cell_start, 0, (size_t)clouds * (maxc + 1) * sizeof(int), st) != cudaSuccess) return EGSPR_E_LAUNCH;
        dim3 g((unsigned)split, (unsigned)clouds);
        knn_grid_bbox_kernel<<<g, 256, 0, st>>>(x, n, bbox);
        knn_grid_count_scatter_kernel<0><<<g, 256, 0, st>>>(x, n, bbox, cell_start, cursor, sorted, params, maxc, ppc);
        knn_grid_scan_kernel<<<clouds, 1024, 0, st>>>(params, n, cell_start, cursor, maxc);
        knn_grid_count_scatter_kernel<1><<<g, 256, 0, st>>>(x, n, bbox, cell_start, cursor, sorted, params, maxc, ppc);
    } else {
        const size_t build_smem = (size_t)(maxc + 1) * sizeof(int);
        if (build_smem > 48 * 1024 && !opt_in_smem(knn_grid_build_kernel, build_smem)) return EGSPR_E_LAUNCH;
        knn_grid_build_kernel<<<clouds, GRID_BUILD_THREADS, build_smem, st>>>(x, n, sorted, cell_start, params, maxc, ppc);
    }
    if (k <= 16) launch_knn_sub<16, 1>(sorted, cell_start, params, clouds, n, k, nbr, maxc, st);       // 16 lanes per query
    else launch_knn_sub<32, 1>(sorted, cell_start, params, clouds, n, k, nbr, maxc, st);               // 16 < k <= 32
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" int egspr_nbr_to_edges(const int32_t *nbr, int clouds, int n, int k, int64_t *edges, void *stream) {
    using namespace egspr;
    if (!nbr || !edges || clouds <= 0 || n <= 0 || k <= 0) return EGSPR_E_INVALID;
    if (clouds > 65535) return EGSPR_E_UNSUPPORTED;
    const int64_t per_cloud = (int64_t)n * k;
    int64_t gx = (per_cloud + 255) / 256; if (gx > 1024) gx = 1024;
    dim3 grid((unsigned)gx, clouds);
    nbr_to_edges_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(nbr, n, k, edges, per_cloud);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}
