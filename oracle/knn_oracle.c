/* TEST INFRASTRUCTURE ONLY -- CPU oracle for the k-NN graph build.
 *
 * PARITY UNPINNED: the reference calls torch_cluster.knn_graph(x, k=16, loop=True)
 * (src/3dmatch_train_egnn_with_batch.py:1005-1006, src/eval_egnn_metrics.py:1156-1157); the
 * arithmetic lives in torch-cluster==1.6.3 (environment.yml:158), which is not vendored under
 * /root/reference and is not installed, and no reference test pins its output.  This file
 * restates torch_cluster's published brute-force CUDA kernel (csrc/cuda/knn_cuda.cu):
 *   for every query i, scan candidates j = 0..N-1 in ascending order,
 *   d2 = sum_d (x[j][d]-x[i][d])^2 accumulated left to right in fp32 with the FMA contraction
 *   nvcc applies by default ( d2 = fma(dz,dz, fma(dy,dy, dx*dx)) ),
 *   insert into a sorted best-k list at the first slot whose distance is STRICTLY greater
 *   (ties keep the earlier = lower index), self included (loop=True).
 * Output: nbr[i*k+s] = s-th nearest candidate of i, nearest first; slots never filled (N<k) = -1.
 * knn_graph then returns edge_index[0] = nbr (flattened), edge_index[1] = i repeated k times.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>

static void knn_range(const float *x, int n, int k, int32_t *nbr, int i0, int i1)
{
    for (int i = i0; i < i1; ++i) {
        float bd[64];
        int32_t bi[64];
        for (int s = 0; s < k; ++s) { bd[s] = 1e10f; bi[s] = -1; }
        const float qx = x[3 * i], qy = x[3 * i + 1], qz = x[3 * i + 2];
        for (int j = 0; j < n; ++j) {
            const float dx = x[3 * j] - qx, dy = x[3 * j + 1] - qy, dz = x[3 * j + 2] - qz;
            const float d2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            for (int e1 = 0; e1 < k; ++e1) {
                if (bd[e1] > d2) {
                    for (int e2 = k - 1; e2 > e1; --e2) { bd[e2] = bd[e2 - 1]; bi[e2] = bi[e2 - 1]; }
                    bd[e1] = d2; bi[e1] = j;
                    break;
                }
            }
        }
        for (int s = 0; s < k; ++s) nbr[(int64_t)i * k + s] = bi[s];
    }
}

void egspr_oracle_knn(const float *x, int n, int k, int32_t *nbr)
{
    knn_range(x, n, k, nbr, 0, n);
}

/* clouds: x [c][n][3] -> nbr [c][n][k]; `threads` host threads split the c*n queries
 * (libgomp is not in this image, so plain pthreads). */
typedef struct { const float *x; int c, n, k; int32_t *nbr; int tid, nthreads; } job_t;

static void *worker(void *arg)
{
    job_t *j = (job_t *)arg;
    const int64_t total = (int64_t)j->c * j->n;
    const int64_t q0 = total * j->tid / j->nthreads, q1 = total * (j->tid + 1) / j->nthreads;
    for (int64_t q = q0; q < q1;) {
        const int b = (int)(q / j->n), i0 = (int)(q % j->n);
        int64_t left = q1 - q;
        int i1 = (left < j->n - i0) ? i0 + (int)left : j->n;
        knn_range(j->x + (int64_t)b * j->n * 3, j->n, j->k, j->nbr + (int64_t)b * j->n * j->k, i0, i1);
        q += i1 - i0;
    }
    return 0;
}

void egspr_oracle_knn_batch(const float *x, int c, int n, int k, int32_t *nbr, int threads)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    job_t jobs[256];
    for (int t = 0; t < threads; ++t) {
        jobs[t] = (job_t){x, c, n, k, nbr, t, threads};
        pthread_create(&th[t], 0, worker, &jobs[t]);
    }
    for (int t = 0; t < threads; ++t) pthread_join(th[t], 0);
}
