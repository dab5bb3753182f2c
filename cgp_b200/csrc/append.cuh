// Incremental refit: append ONE sample to a factorised cluster with theta unchanged.
// The reference re-factorises every cluster from scratch in every sequential iteration
// (REF/cGP/cGP_constrained.py:276,315-327 -> SK/_gpr.py:349-367); with K' = [[K, k], [k^T, kappa]] the new factor is
//   L' = [[L, 0], [l^T, dd]],  l = L^-1 k = W k,  dd = sqrt(kappa - l.l),
//   W' = L'^-1 = [[W, 0], [-(l^T W)/dd, 1/dd]],
// i.e. two triangular mat-vecs against the stored W = L^-1 (2 n^2 flops, 16 n^2 bytes: HBM-bound) instead of n^3/3.
// alpha' = K'^-1 y' is recomputed in full (y is re-normalised when a sample arrives, SK/_gpr.py:275-280):
//   z = W' y',  alpha' = W'^T z,  lml = -1/2 z.z - sum log L'_ii - n/2 log 2pi.
// All reductions have a fixed order (no atomics).
#pragma once
#include "common.cuh"

#define APP_CH 128  // rows per partial chunk of the transposed mat-vec

// out_i = sum_{j <= i} W[i,j] x[j]   (i < n).  warp per row; grid ceil(n/8), block 256
__global__ void __launch_bounds__(256)
k_lower_matvec(const double* __restrict__ W, long long ld, int n, const double* __restrict__ x, double* __restrict__ out) {
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= n) return;
    const double* w = W + (long long)row * ld;
    double s = 0.0;
    for (int j = lane; j <= row; j += 32) s += w[j] * x[j];
    s = warp_sum(s);
    if (lane == 0) out[row] = s;
}

// part[rc][j] = sum_{i in chunk rc, i >= j, i < n} W[i,j] x[i].   grid (ceil(n/128) col blocks, ceil(n/APP_CH) row chunks),
// block 128 (thread = column: coalesced row reads).  Chunks above the diagonal block are skipped (their part is unused).
__global__ void __launch_bounds__(128)
k_lower_tmatvec_part(const double* __restrict__ W, long long ld, int n, const double* __restrict__ x,
                     double* __restrict__ part, int pstride) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    const int i0 = blockIdx.y * APP_CH;
    if (i0 + APP_CH <= blockIdx.x * 128) return;  // chunk entirely above the diagonal
    __shared__ double xs[APP_CH];
    if (threadIdx.x < APP_CH) xs[threadIdx.x] = (i0 + threadIdx.x < n) ? x[i0 + threadIdx.x] : 0.0;
    __syncthreads();
    if (j >= n) return;
    const int iend = min(n, i0 + APP_CH);
    double s = 0.0;
    for (int i = max(i0, j); i < iend; ++i) s += W[(long long)i * ld + j] * xs[i - i0];
    part[(long long)blockIdx.y * pstride + j] = s;
}

// out_j = scale * sum_{rc >= rc0(j)} part[rc][j] in ascending rc.   grid ceil(n/256), block 256
// scale_ptr == nullptr -> 1;  otherwise scale = -1 / *scale_ptr  (the new W row of the append).
__global__ void __launch_bounds__(256)
k_lower_tmatvec_fin(const double* __restrict__ part, int pstride, int n, const double* __restrict__ scale_ptr,
                    double* __restrict__ out, long long ostride) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= n) return;
    const int nch = (n + APP_CH - 1) / APP_CH;
    double s = 0.0;
    for (int rc = (j / 128) * 128 / APP_CH; rc < nch; ++rc) s += part[(long long)rc * pstride + j];
    if (scale_ptr) s = -s / *scale_ptr;
    out[(long long)j * ostride] = s;
}

// One block: dd = sqrt(kappa - l.l); L[n, 0:n] = l, L[n,n] = dd, W[n,n] = 1/dd, dd_out, info (n+1 when not PD).
// n = OLD sample count = index of the new row.
__global__ void __launch_bounds__(256)
k_append_finish(double* __restrict__ L, double* __restrict__ W, long long ld, int n, const double* __restrict__ l,
                const double* __restrict__ theta, int d, double alpha, double* __restrict__ dd_out,
                int* __restrict__ info) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const double v = l[i];
        s += v * v;
        L[(long long)n * ld + i] = v;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double q = 0.0;
        for (int w = 0; w < 8; ++w) q += red[w];
        const double kappa = (1.0 + exp(theta[d])) + alpha;  // same diagonal as k_assemble
        const double piv = kappa - q;
        const bool ok = piv > 0.0;
        const double dd = ok ? sqrt(piv) : 1.0;
        if (!ok) *info = n + 1;
        L[(long long)n * ld + n] = dd;
        W[(long long)n * ld + n] = 1.0 / dd;
        *dd_out = dd;
    }
}

// lml = -1/2 z.z - sum_i log L_ii - n/2 log 2pi  (SK/_gpr.py:613-617).  one block of 256
__global__ void __launch_bounds__(256)
k_lml_from_factor(const double* __restrict__ L, long long ld, int n, const double* __restrict__ z, double* __restrict__ lml) {
    __shared__ double r1[8], r2[8];
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) {
        a += z[i] * z[i];
        b += log(L[(long long)i * ld + i]);
    }
    a = warp_sum(a);
    b = warp_sum(b);
    if ((threadIdx.x & 31) == 0) r1[threadIdx.x >> 5] = a, r2[threadIdx.x >> 5] = b;
    __syncthreads();
    if (threadIdx.x == 0) {
        double q = 0.0, ldet = 0.0;
        for (int w = 0; w < 8; ++w) q += r1[w], ldet += r2[w];
        *lml = -0.5 * q - ldet - 0.5 * (double)n * 1.8378770664093453;
    }
}
