// K7: brute-force k-NN routing.  One thread per query, training points streamed through shared memory
// in index order.  Distances are the exact fp64 reduced distances of scikit-learn's KDTree
// (SK/metrics/_dist_metrics.pxd.tp:44-49: tmp = x1[j]-x2[j]; d += tmp*tmp, un-fused), so the k-neighbour
// set -- and with it the label -- is bit-exact.  A candidate replaces a kept neighbour only when strictly
// closer (SK/utils/_heap.pyx:46-47), i.e. equal distances keep the lowest training index.
#pragma once
#include "common.cuh"

#define KNN_TILE 1024
#define KNN_THREADS 256

template <int D>
__global__ void __launch_bounds__(KNN_THREADS)
k_knn(const double* __restrict__ X, const long long* __restrict__ lab, int n, int k, const double* __restrict__ Q, long long M,
      long long* __restrict__ out, int only_flagged, const long long* __restrict__ qlist, const int* __restrict__ qcount,
      int qcap) {
    extern __shared__ __align__(16) double sm[];
    double* xs = sm;                                          // [KNN_TILE][D]
    int* ls = reinterpret_cast<int*>(sm + KNN_TILE * D);      // [KNN_TILE]
    // Passes behind the tensor-core prefilter (knn_tc.cuh), which marks the queries it could not certify with
    // out[q] = -1 and appends them to qlist (up to qcap entries):
    //   qlist != NULL   dense pass over the listed queries (thread i handles qlist[i]);
    //   only_flagged    sweep over ALL queries for marks that did not fit the list; blocks without one exit at once.
    long long q = (long long)blockIdx.x * KNN_THREADS + threadIdx.x;
    bool active = q < M;
    if (qlist) {
        const int cnt = min(*qcount, qcap);
        if ((long long)blockIdx.x * KNN_THREADS >= cnt) return;
        active = q < cnt;
        q = active ? qlist[q] : 0;
    } else if (only_flagged) {
        active = active && out[q] == -1;
        if (!__syncthreads_or(active)) return;
    }
    double qv[D];
#pragma unroll
    for (int p = 0; p < D; ++p) qv[p] = active ? Q[q * D + p] : 0.0;
    double bd[CGP_MAX_K];
    int bl[CGP_MAX_K];
#pragma unroll
    for (int s = 0; s < CGP_MAX_K; ++s) {
        bd[s] = INFINITY;
        bl[s] = -1;
    }
    double worst = INFINITY;
    for (int t0 = 0; t0 < n; t0 += KNN_TILE) {
        const int tn = min(KNN_TILE, n - t0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < tn * D; idx += KNN_THREADS) xs[idx] = X[(long long)t0 * D + idx];
        for (int idx = threadIdx.x; idx < tn; idx += KNN_THREADS) ls[idx] = (int)lab[t0 + idx];
        __syncthreads();
        for (int j = 0; j < tn; ++j) {
            double dist = 0.0;
#pragma unroll
            for (int p = 0; p < D; ++p) {
                double t = __dsub_rn(qv[p], xs[j * D + p]);
                dist = __dadd_rn(dist, __dmul_rn(t, t));
            }
            if (dist < worst) {
                double cd = dist;
                int cl = ls[j];
                bool ins = false;
#pragma unroll
                for (int s = 0; s < CGP_MAX_K; ++s) {
                    if (s < k && (ins || cd < bd[s])) {
                        double td = bd[s];
                        int tl = bl[s];
                        bd[s] = cd;
                        bl[s] = cl;
                        cd = td;
                        cl = tl;
                        ins = true;
                    }
                }
#pragma unroll
                for (int s = 0; s < CGP_MAX_K; ++s)
                    if (s == k - 1) worst = bd[s];
            }
        }
    }
    if (!active) return;
    // uniform vote, smallest label on ties (scipy.stats.mode, SK/utils/fixes.py:58-59)
    int best_l = -1, best_c = 0;
#pragma unroll
    for (int s = 0; s < CGP_MAX_K; ++s) {
        if (s < k && bl[s] >= 0) {
            int cnt = 0;
#pragma unroll
            for (int u = 0; u < CGP_MAX_K; ++u) cnt += (u < k && bl[u] == bl[s]) ? 1 : 0;
            if (cnt > best_c || (cnt == best_c && bl[s] < best_l)) {
                best_c = cnt;
                best_l = bl[s];
            }
        }
    }
    out[q] = best_l;
}

