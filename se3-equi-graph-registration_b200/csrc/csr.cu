// Row-major CSR ("reverse k-NN lists") of the batch graph, built once per graph.
//
// E_GCL aggregates with unsorted_segment_sum(..., row) where row = edge_index[0]
// (src/3dmatch_train_egnn_with_batch.py:253-254, 263-265, 343-348); for knn_graph output row is
// the NEIGHBOUR id, so node g sums over {e : row[e] = g}, a variable-length set.  Instead of
// scatter-add atomics per layer, the edge list is transposed once: a counting sort by row
// (integer atomics only decide slots; a per-row rank pass then restores ascending edge order, so
// the result is deterministic and sums run in the same order as scatter_add_ on CPU).
#include <cstdlib>

#include "egspr_common.cuh"

namespace egspr {

struct NbrSource {   // edges implied by nbr[cloud][i][s]: row = nbr, col = i, e = i*k+s
    const int32_t *nbr; int n, k;
    unsigned kmul;       // ceil(2^32 / k) (0 for k = 1): e / k == umulhi(e, kmul) while e * k < 2^32 (the fused kernel: e < 2^16, k <= 64)
    static constexpr bool kColImplied = true;        // col = e / k is in range by construction
    __device__ __forceinline__ void get(int cloud, int64_t e, int64_t epc, int &r, int &c) const {
        r = nbr[cloud * epc + e]; c = (int)((unsigned)e / (unsigned)k);      // e < n*k < 2^31: 32-bit division
    }
    __device__ __forceinline__ int row(int cloud, int e, int64_t epc) const { return nbr[cloud * epc + e]; }
    __device__ __forceinline__ int col_small(int cloud, int e, int64_t epc) const { return kmul ? (int)__umulhi((unsigned)e, kmul) : e; }
    bool kmul_ok(int64_t epc) const { return epc * (int64_t)k < ((int64_t)1 << 32); }      // exactness range of col_small
};
struct EdgeSource {  // edges[cloud][2][E] int64 (torch_cluster / reference layout)
    const int64_t *edges; int n;
    static constexpr bool kColImplied = false;
    __device__ __forceinline__ void get(int cloud, int64_t e, int64_t epc, int &r, int &c) const {
        r = (int)edges[(cloud * 2 + 0) * epc + e]; c = (int)edges[(cloud * 2 + 1) * epc + e];
    }
    __device__ __forceinline__ int row(int cloud, int e, int64_t epc) const { return (int)edges[(cloud * 2 + 0) * epc + e]; }
    __device__ __forceinline__ int col_small(int cloud, int e, int64_t epc) const { return (int)edges[(cloud * 2 + 1) * epc + e]; }
};

__device__ __forceinline__ bool fix_range(int &v, int n) {
    if ((unsigned)v < (unsigned)n) return false;
    v = v < 0 ? 0 : n - 1;
    return true;
}

template <class Src>
__global__ void csr_count_kernel(Src src, int n, int64_t epc, int32_t *__restrict__ deg,
                                 int32_t *__restrict__ err_flag) {
    const int cloud = blockIdx.y;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < epc; e += (int64_t)gridDim.x * blockDim.x) {
        int r, c;
        src.get(cloud, e, epc, r, c);
        bool bad = fix_range(r, n);
        bad |= fix_range(c, n);
        if (bad && err_flag) *err_flag = 1;
        atomicAdd(&deg[(int64_t)cloud * n + r], 1);
    }
}

// ptr = exclusive scan of the degrees over ALL nodes of the batch (every cloud's degrees sum to epc, so this equals
// cloud*epc + the scan inside the cloud); deg is reset to 0 (it becomes the fill cursor).  Two launches over chunks of
// 4096 nodes, any number of CTAs: a CTA per cloud serialised a 131072-node cloud on one SM (118 us of the 237 us build).
constexpr int SC_CHUNK = 4096;
__global__ void __launch_bounds__(1024) csr_chunk_sum_kernel(const int32_t *__restrict__ deg, int64_t G, int32_t *__restrict__ chunk_sums) {
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t base = (int64_t)blockIdx.x * SC_CHUNK;
    int v = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int64_t i = base + u * 1024 + threadIdx.x; if (i < G) v += deg[i]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) warp_tot[warp] = v;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) chunk_sums[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(1024) csr_scan_apply_kernel(int32_t *__restrict__ deg, int64_t G, const int32_t *__restrict__ chunk_sums,
                                                              int32_t *__restrict__ ptr, int32_t total_edges) {
    __shared__ int warp_tot[32];
    __shared__ int prefix_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int pre = 0;                                             // sum of the chunks before this one
    for (int c = threadIdx.x; c < (int)blockIdx.x; c += 1024) pre += chunk_sums[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) pre += __shfl_xor_sync(0xffffffffu, pre, o);
    if (lane == 0) warp_tot[warp] = pre;
    __syncthreads();
    if (warp == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) prefix_s = t;
    }
    __syncthreads();
    // this thread's four CONSECUTIVE nodes
    const int64_t i0 = (int64_t)blockIdx.x * SC_CHUNK + 4 * threadIdx.x;
    int d[4], s = 0;
#pragma unroll
    for (int u = 0; u < 4; ++u) { d[u] = (i0 + u < G) ? deg[i0 + u] : 0; s += d[u]; }
    int sc = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
    __syncthreads();                                         // warp_tot is reused
    if (lane == 31) warp_tot[warp] = sc;
    __syncthreads();
    if (warp == 0) {
        const int t = warp_tot[lane];
        int u = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int w = __shfl_up_sync(0xffffffffu, u, o); if (lane >= o) u += w; }
        warp_tot[lane] = u - t;
    }
    __syncthreads();
    int excl = prefix_s + warp_tot[warp] + sc - s;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if (i0 + u < G) { ptr[i0 + u] = excl; deg[i0 + u] = 0; }
        excl += d[u];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) ptr[G] = total_edges;
}

template <class Src>
__global__ void csr_fill_kernel(Src src, int n, int64_t epc, int32_t *__restrict__ cursor,
                                const int32_t *__restrict__ ptr, int32_t *__restrict__ tmp) {
    const int cloud = blockIdx.y;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < epc; e += (int64_t)gridDim.x * blockDim.x) {
        int r, c;
        src.get(cloud, e, epc, r, c);
        fix_range(r, n);
        const int64_t g = (int64_t)cloud * n + r;
        const int slot = atomicAdd(&cursor[g], 1);
        tmp[ptr[g] + slot] = (int32_t)e;
    }
}

// thread per list entry: rank = entries of the same row's (unsorted) list with a smaller edge id -- restores ascending
// edge order -- then emit.  (A warp per row with a shuffle loop spent most of its instructions on half-empty warps.)
template <class Src>
__global__ void __launch_bounds__(256) csr_emit_kernel(Src src, int n, int64_t epc, int64_t num_edges,
                                                       const int32_t *__restrict__ ptr,
                                                       const int32_t *__restrict__ tmp,
                                                       int32_t *__restrict__ csr_row,
                                                       int32_t *__restrict__ csr_col,
                                                       int32_t *__restrict__ csr_eid) {
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < num_edges; p += (int64_t)gridDim.x * blockDim.x) {
        const int cloud = (int)((unsigned)p / (unsigned)epc);                 // num_edges < 2^31 (checked by the caller)
        const int e = tmp[p];
        int r, c;
        src.get(cloud, e, epc, r, c);
        fix_range(r, n); fix_range(c, n);
        const int32_t g = cloud * n + r;
        const int base = ptr[g], deg = ptr[g + 1] - base;
        int rank = 0;
        for (int j = 0; j < deg; ++j) rank += (__ldg(tmp + base + j) < e);
        csr_eid[base + rank] = e;
        csr_row[base + rank] = g;
        csr_col[base + rank] = cloud * n + c;
    }
}

// ---- fused build for small clouds: ONE CTA per cloud does count -> scan -> fill -> rank/emit with the
// degree counters, row offsets and the unsorted edge list all in shared memory (one launch instead of a
// memset + four kernels, no global atomics).  Same output as the generic path.
// The unsorted list holds (row << 16 | edge) -- both fit 16 bits at shared-memory sizes -- so the rank pass runs one
// THREAD per entry (rank = entries of its row's list that sort before it, read from shared memory; equal rows make
// the packed order the edge order) instead of one warp per row with a shuffle loop, which was 54 % of this kernel's
// instructions (ncu, profiles/r02r_csr_knn_ncu.txt).
constexpr int CF_THREADS = 1024;
constexpr size_t CF_MAX_SMEM = 200 * 1024;

template <class Src>
__global__ void __launch_bounds__(CF_THREADS) csr_fused_kernel(Src src, int n, int64_t epc, int clouds,
                                                               int32_t *__restrict__ csr_ptr, int32_t *__restrict__ csr_row,
                                                               int32_t *__restrict__ csr_col, int32_t *__restrict__ csr_eid,
                                                               int32_t *__restrict__ err_flag) {
    extern __shared__ int32_t cf_smem[];
    int32_t *deg = cf_smem;                 // [n]      degree, later the fill cursor
    int32_t *off = deg + n;                 // [n + 1]  exclusive offsets inside the cloud
    int32_t *tmp = off + n + 1;             // [epc]    (row << 16 | edge id) grouped by row, unsorted inside a row
    __shared__ int warp_tot[32];
    __shared__ int carry_s;
    const int cloud = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int E = (int)epc;
    for (int i = tid; i < n; i += CF_THREADS) deg[i] = 0;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int e0 = tid; e0 < E; e0 += 8 * CF_THREADS) {        // 8 edges per thread in flight (the loads are independent)
        int r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * CF_THREADS;
            r[u] = 0;
            if (e < E) r[u] = src.row(cloud, e, epc);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (e0 + u * CF_THREADS < E) {
                if (fix_range(r[u], n) && err_flag) *err_flag = 1;
                atomicAdd(&deg[r[u]], 1);
            }
        }
    }
    __syncthreads();
    for (int base = 0; base < n; base += CF_THREADS) {        // exclusive scan of deg -> off; deg becomes the cursor
        const int i = base + tid;
        const int v = i < n ? deg[i] : 0;
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
        if (lane == 31) warp_tot[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane], u = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(0xffffffffu, u, o); if (lane >= o) u += w; }
            warp_tot[lane] = u - t;
        }
        __syncthreads();
        const int excl = carry_s + warp_tot[warp] + sc - v;
        if (i < n) {
            off[i] = excl; deg[i] = excl;
            csr_ptr[(int64_t)cloud * n + i] = (int32_t)(cloud * epc) + excl;
        }
        __syncthreads();
        if (tid == CF_THREADS - 1) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        off[n] = E;
        if (cloud == clouds - 1) csr_ptr[(int64_t)clouds * n] = (int32_t)(clouds * epc);
    }
    for (int e0 = tid; e0 < E; e0 += 8 * CF_THREADS) {
        int r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * CF_THREADS;
            r[u] = 0;
            if (e < E) r[u] = src.row(cloud, e, epc);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * CF_THREADS;
            if (e < E) { fix_range(r[u], n); tmp[atomicAdd(&deg[r[u]], 1)] = (r[u] << 16) | e; }
        }
    }
    __syncthreads();
    const int64_t gbase = (int64_t)cloud * epc;
    for (int p0 = tid; p0 < E; p0 += 4 * CF_THREADS) {        // thread per entry; 4 entries (their column loads) in flight
        int v[4], c[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int p = p0 + u * CF_THREADS;
            v[u] = p < E ? tmp[p] : 0;
            c[u] = src.col_small(cloud, v[u] & 0xffff, epc);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (p0 + u * CF_THREADS >= E) continue;
            const int r = v[u] >> 16, e = v[u] & 0xffff;
            const int base = off[r], d = off[r + 1] - base;
            int rank = 0;
            for (int j = 0; j < d; ++j) rank += (tmp[base + j] < v[u]);       // same row: packed order == edge order
            if (!Src::kColImplied) { if (fix_range(c[u], n) && err_flag) *err_flag = 1; }
            const int64_t o = gbase + base + rank;
            csr_eid[o] = e;
            csr_row[o] = cloud * n + r;
            csr_col[o] = cloud * n + c[u];
        }
    }
}

// ---- mid-size k-NN clouds (their degree counters + edge list do not fit one CTA's shared memory): the same fused build with
// `split` CTAs per cloud, CTA j owning the rows [j R, (j + 1) R) of the cloud.  Every CTA scans ALL edges of its cloud
// (the neighbour table is L2-resident; `split` <= 8) and keeps those whose row it owns; the number of edges with a
// smaller row -- the range's offset in the cloud's lists -- falls out of the same scan, so the CTAs of a cloud never talk
// to each other.  The unsorted list holds (local row << 18 | edge).  A range whose edges exceed the list's capacity
// (duplicate-heavy clouds send every edge to a few low-index rows) is done in several rounds of as many whole rows as
// fit; a k-NN row has at most n <= 16384 entries, which always fits.
constexpr int CS_ROWS = 2048;                       // rows per CTA
constexpr int CS_CAP = (int)(CF_MAX_SMEM / 4) - (2 * CS_ROWS + 1);      // list entries per round

__global__ void __launch_bounds__(CF_THREADS) csr_split_kernel(NbrSource src, int n, int64_t epc, int clouds,
                                                               int32_t *__restrict__ csr_ptr, int32_t *__restrict__ csr_row,
                                                               int32_t *__restrict__ csr_col, int32_t *__restrict__ csr_eid,
                                                               int32_t *__restrict__ err_flag) {
    extern __shared__ int32_t cf_smem[];
    int32_t *deg = cf_smem;                 // [CS_ROWS]      degree, later the fill cursor
    int32_t *off = deg + CS_ROWS;           // [CS_ROWS + 1]  exclusive offsets inside the range
    int32_t *tmp = off + CS_ROWS + 1;       // [CS_CAP]       (local row << 18 | edge id), unsorted inside a row
    __shared__ int warp_tot[32];
    __shared__ int carry_s, below_s, round_end_s;
    const int cloud = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.x * CS_ROWS, R = min(CS_ROWS, n - r0);
    const int E = (int)epc;
    for (int i = tid; i < CS_ROWS; i += CF_THREADS) deg[i] = 0;
    if (tid == 0) { carry_s = 0; below_s = 0; }
    __syncthreads();
    int below = 0;                          // edges of the cloud with row < r0
    for (int e0 = tid; e0 < E; e0 += 8 * CF_THREADS) {
        int r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = e0 + u * CF_THREADS;
            r[u] = e < E ? src.row(cloud, e, epc) : 0x7fffffff;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (e0 + u * CF_THREADS < E) {
                if (fix_range(r[u], n) && err_flag) *err_flag = 1;
                below += r[u] < r0;
                const int l = r[u] - r0;
                if ((unsigned)l < (unsigned)R) atomicAdd(&deg[l], 1);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) below += __shfl_xor_sync(0xffffffffu, below, o);
    if (lane == 0 && below) atomicAdd(&below_s, below);
    __syncthreads();
    for (int base = 0; base < CS_ROWS; base += CF_THREADS) {       // exclusive scan of deg -> off; deg becomes the cursor
        const int i = base + tid;
        const int v = i < R ? deg[i] : 0;
        int sc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
        if (lane == 31) warp_tot[warp] = sc;
        __syncthreads();
        if (warp == 0) {
            int t = warp_tot[lane], u = t;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int w = __shfl_up_sync(0xffffffffu, u, o); if (lane >= o) u += w; }
            warp_tot[lane] = u - t;
        }
        __syncthreads();
        const int excl = carry_s + warp_tot[warp] + sc - v;
        if (i < R) {
            off[i] = excl; deg[i] = excl;
            csr_ptr[(int64_t)cloud * n + r0 + i] = (int32_t)(cloud * epc) + below_s + excl;
        }
        __syncthreads();
        if (tid == CF_THREADS - 1) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) {
        off[R] = carry_s;
        if (cloud == clouds - 1 && r0 + R == n) csr_ptr[(int64_t)clouds * n] = (int32_t)(clouds * epc);
    }
    __syncthreads();
    const int64_t gbase = (int64_t)cloud * epc + below_s;
    for (int a = 0; a < R;) {               // rounds of whole rows [a, b) whose lists fit the shared-memory list
        if (tid == 0) {
            int lo = a + 1, hi = R;         // largest b with off[b] - off[a] <= CS_CAP (a single row always fits)
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (off[mid] - off[a] <= CS_CAP) lo = mid; else hi = mid - 1; }
            round_end_s = lo;
        }
        __syncthreads();
        const int b = round_end_s, oa = off[a], cnt = off[b] - oa;
        for (int e0 = tid; e0 < E; e0 += 8 * CF_THREADS) {
            int r[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * CF_THREADS;
                r[u] = e < E ? src.row(cloud, e, epc) : 0x7fffffff;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int e = e0 + u * CF_THREADS;
                if (e < E) {
                    fix_range(r[u], n);
                    const int l = r[u] - r0;
                    if (l >= a && l < b) tmp[atomicAdd(&deg[l], 1) - oa] = (l << 18) | e;
                }
            }
        }
        __syncthreads();
        for (int p = tid; p < cnt; p += CF_THREADS) {          // thread per entry: rank inside its row's list, emit
            const int v = tmp[p];
            const int l = v >> 18, e = v & 0x3ffff;
            const int base = off[l] - oa, d = off[l + 1] - off[l];
            int rank = 0;
            for (int j = 0; j < d; ++j) rank += (tmp[base + j] < v);
            const int64_t o = gbase + off[l] + rank;
            csr_eid[o] = e;
            csr_row[o] = cloud * n + r0 + l;
            csr_col[o] = cloud * n + src.col_small(cloud, e, epc);
        }
        __syncthreads();
        a = b;
    }
}

template <class Src>
static int csr_build(Src src, int clouds, int n, int64_t epc, int32_t *csr_ptr, int32_t *csr_row,
                     int32_t *csr_col, int32_t *csr_eid, void *ws, size_t ws_bytes, int32_t *err_flag,
                     cudaStream_t st) {
    const int64_t G = (int64_t)clouds * n, E = (int64_t)clouds * epc;
    if (clouds > 65535 || E >= (int64_t)0x7fffffff || G >= (int64_t)0x7fffffff) return EGSPR_E_UNSUPPORTED;
    if (ws_bytes < egspr_csr_workspace_bytes(G, E)) return EGSPR_E_WORKSPACE;
    {   // small clouds, enough of them to fill the GPU: fused single-launch build in shared memory
        const size_t smem = sizeof(int32_t) * (size_t)(2 * (int64_t)n + 1 + epc);
        if (smem <= CF_MAX_SMEM && clouds >= 16 && epc <= 65536 && n <= 32768) {      // (row << 16 | edge) must fit an int32
            if (!opt_in_smem(csr_fused_kernel<Src>, CF_MAX_SMEM)) return EGSPR_E_LAUNCH;
            csr_fused_kernel<Src><<<clouds, CF_THREADS, smem, st>>>(src, n, epc, clouds, csr_ptr, csr_row, csr_col, csr_eid, err_flag);
            EGSPR_CHECK_LAUNCH();
            return EGSPR_OK;
        }
    }
    if constexpr (Src::kColImplied) {   // k-NN tables of mid-size clouds: split build, `split` CTAs per cloud
        const int split = (n + CS_ROWS - 1) / CS_ROWS;
        if (split <= 8 && epc < (1 << 18) && (int64_t)split * clouds >= 16 && src.kmul_ok(epc)) {
            if (!opt_in_smem(csr_split_kernel, CF_MAX_SMEM)) return EGSPR_E_LAUNCH;
            csr_split_kernel<<<dim3((unsigned)split, (unsigned)clouds), CF_THREADS, CF_MAX_SMEM, st>>>(src, n, epc, clouds, csr_ptr, csr_row,
                                                                                                      csr_col, csr_eid, err_flag);
            EGSPR_CHECK_LAUNCH();
            return EGSPR_OK;
        }
    }
    int32_t *deg = (int32_t *)ws;
    int32_t *tmp = deg + ((G + 31) / 32) * 32;
    if (cudaMemsetAsync(deg, 0, sizeof(int32_t) * G, st) != cudaSuccess) return EGSPR_E_LAUNCH;
    int64_t gx = (epc + 255) / 256;
    if (gx > 2048) gx = 2048;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, clouds);
    csr_count_kernel<<<grid, 256, 0, st>>>(src, n, epc, deg, err_flag);
    {
        const unsigned chunks = (unsigned)((G + SC_CHUNK - 1) / SC_CHUNK);
        int32_t *chunk_sums = tmp;                 // tmp is written by the fill kernel, after the scan
        if ((int64_t)chunks > E + 32) return EGSPR_E_WORKSPACE;
        csr_chunk_sum_kernel<<<chunks, 1024, 0, st>>>(deg, G, chunk_sums);
        csr_scan_apply_kernel<<<chunks, 1024, 0, st>>>(deg, G, chunk_sums, csr_ptr, (int32_t)E);
    }
    csr_fill_kernel<<<grid, 256, 0, st>>>(src, n, epc, deg, csr_ptr, tmp);
    int64_t gw = (E + 255) / 256;
    if (gw > 148 * 32) gw = 148 * 32;
    csr_emit_kernel<<<(unsigned)gw, 256, 0, st>>>(src, n, epc, E, csr_ptr, tmp, csr_row, csr_col, csr_eid);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

// unsorted_segment_sum (3dm:343-348) as a deterministic pull over the CSR: warp per segment,
// lanes over channels, entries in ascending original order
__global__ void __launch_bounds__(256) segment_sum_kernel(const float *__restrict__ data, int channels,
                                                          const int32_t *__restrict__ ptr,
                                                          const int32_t *__restrict__ eid, int64_t segments,
                                                          float *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t g = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); g < segments; g += warps) {
        const int lo = ptr[g], hi = ptr[g + 1];
        for (int c = lane; c < channels; c += 32) {
            float acc = 0.f;
            for (int p = lo; p < hi; ++p) acc += __ldg(data + (int64_t)eid[p] * channels + c);
            out[g * channels + c] = acc;
        }
    }
}

}  // namespace egspr

extern "C" int egspr_segment_sum(const float *data, int channels, const int32_t *csr_ptr, const int32_t *csr_eid,
                                 int64_t num_segments, float *out, void *stream) {
    using namespace egspr;
    if (!data || !csr_ptr || !csr_eid || !out || channels <= 0 || num_segments <= 0) return EGSPR_E_INVALID;
    int64_t gw = (num_segments + 7) / 8;
    if (gw > 148 * 32) gw = 148 * 32;
    segment_sum_kernel<<<(unsigned)gw, 256, 0, (cudaStream_t)stream>>>(data, channels, csr_ptr, csr_eid, num_segments, out);
    EGSPR_CHECK_LAUNCH();
    return EGSPR_OK;
}

extern "C" size_t egspr_csr_workspace_bytes(int64_t num_nodes, int64_t num_edges) {
    return sizeof(int32_t) * (size_t)(((num_nodes + 31) / 32) * 32 + num_edges + 32);
}

extern "C" int egspr_csr_from_nbr(const int32_t *nbr, int clouds, int n, int k, int32_t *csr_ptr,
                                  int32_t *csr_row, int32_t *csr_col, int32_t *csr_eid, void *workspace,
                                  size_t workspace_bytes, int32_t *err_flag, void *stream) {
    using namespace egspr;
    if (!nbr || !csr_ptr || !csr_row || !csr_col || !csr_eid || !workspace || clouds <= 0 || n <= 0 || k <= 0)
        return EGSPR_E_INVALID;
    NbrSource src{nbr, n, k, k > 1 ? (unsigned)((0x100000000ull + (unsigned)k - 1) / (unsigned)k) : 0u};
    return csr_build(src, clouds, n, (int64_t)n * k, csr_ptr, csr_row, csr_col, csr_eid, workspace,
                     workspace_bytes, err_flag, (cudaStream_t)stream);
}

extern "C" int egspr_csr_from_edges(const int64_t *edges, int clouds, int n, int64_t edges_per_cloud,
                                    int32_t *csr_ptr, int32_t *csr_row, int32_t *csr_col, int32_t *csr_eid,
                                    void *workspace, size_t workspace_bytes, int32_t *err_flag, void *stream) {
    using namespace egspr;
    if (!edges || !csr_ptr || !csr_row || !csr_col || !csr_eid || !workspace || clouds <= 0 || n <= 0 ||
        edges_per_cloud <= 0)
        return EGSPR_E_INVALID;
    EdgeSource src{edges, n};
    return csr_build(src, clouds, n, edges_per_cloud, csr_ptr, csr_row, csr_col, csr_eid, workspace,
                     workspace_bytes, err_flag, (cudaStream_t)stream);
}
