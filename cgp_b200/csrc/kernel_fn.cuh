// Covariance-function kernels: batched training-matrix assembly (K1), dense / cross kernel matrices,
// and the hyper-parameter gradient contraction (K2).  HBM-write-bound elementwise work: one CTA per
// 128x128 tile, the 2x128 input rows staged in shared memory, coalesced row-major stores.
#pragma once
#include "common.cuh"

#define SQRT3 1.7320508075688772

template <int KIND>
__device__ __forceinline__ double stationary_from_sq(double s) {
    if (KIND == 0) {  // Matern nu=1.5 (SK kernels.py:1725-1726)
        double k = sqrt(s) * SQRT3;
        return (1.0 + k) * exp(-k);
    } else {  // RBF (SK kernels.py:1561-1562)
        return exp(-0.5 * s);
    }
}

// i-major enumeration of lower tiles: idx -> (i, j), i >= j
__device__ __forceinline__ void lower_tile(int idx, int& i, int& j) {
    i = 0;
    while (idx > i) {
        idx -= i + 1;
        ++i;
    }
    j = idx;
}

// K1 batched: K tile (i >= j) of matrix `mat` = k(X_c/l, X_c/l) + (noise + alpha) I, pad block = I.
// grid (Tmax(Tmax+1)/2, n_mats), block 256.  smem: 2 * 128 * (d|1) doubles.
template <int KIND>
__global__ void __launch_bounds__(256)
k_assemble(double* __restrict__ Abuf, const double* __restrict__ X, const double* __restrict__ theta,
           const MatMeta* __restrict__ meta, int d, double alpha) {
    extern __shared__ __align__(16) double sm[];
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    int ti, tj;
    lower_tile(blockIdx.x, ti, tj);
    if (ti >= T) return;
    const int sd = d | 1;
    double* xi = sm;
    double* xj = sm + CGP_TILE * sd;
    const double* th = theta + (long long)m.theta_idx * (d + 1);
    for (int idx = threadIdx.x; idx < CGP_TILE * d; idx += 256) {
        int r = idx / d, p = idx - r * d;
        double l = exp(th[p]);
        int gi = ti * CGP_TILE + r, gj = tj * CGP_TILE + r;
        xi[r * sd + p] = (gi < m.n) ? X[(long long)(m.xoff + gi) * d + p] / l : 0.0;
        xj[r * sd + p] = (gj < m.n) ? X[(long long)(m.xoff + gj) * d + p] / l : 0.0;
    }
    __syncthreads();
    const double noise = exp(th[d]);
    const int c = threadIdx.x & 127;
    const int gj = tj * CGP_TILE + c;
    double* A = Abuf + m.a_off;
    const long long ld = m.npad;
    for (int r = threadIdx.x >> 7; r < CGP_TILE; r += 2) {
        const int gi = ti * CGP_TILE + r;
        double v;
        if (gi >= m.n || gj >= m.n) {
            v = (gi == gj) ? 1.0 : 0.0;
        } else if (gi == gj) {
            v = (1.0 + noise) + alpha;  // fill_diagonal(1) + White (Sum) + alpha, SK kernels.py:1743,871; _gpr.py:589
        } else {
            double s = 0.0;
            for (int p = 0; p < d; ++p) {
                double t = xi[r * sd + p] - xj[c * sd + p];
                s += t * t;
            }
            v = stationary_from_sq<KIND>(s);
        }
        A[(long long)gi * ld + gj] = v;
    }
}

// Dense k(Q, X) tile kernel.  out[a, b] for a < mrows_out, b < ncols_out (zero beyond m / n).
// self != 0: Q == X and the diagonal gets (1 + noise) + alpha.   grid (ceil(ncols_out/128), ceil(mrows_out/128))
template <int KIND>
__global__ void __launch_bounds__(256)
k_cross(const double* __restrict__ Q, long long m, const double* __restrict__ X, long long n, int d,
        const double* __restrict__ th, double* __restrict__ out, long long ldo, long long mrows_out,
        long long ncols_out, int self, double alpha) {
    extern __shared__ __align__(16) double sm[];
    const int sd = d | 1;
    double* xi = sm;
    double* xj = sm + CGP_TILE * sd;
    const long long r0 = (long long)blockIdx.y * CGP_TILE, c0 = (long long)blockIdx.x * CGP_TILE;
    for (int idx = threadIdx.x; idx < CGP_TILE * d; idx += 256) {
        int r = idx / d, p = idx - r * d;
        double l = exp(th[p]);
        xi[r * sd + p] = (r0 + r < m) ? Q[(r0 + r) * d + p] / l : 0.0;
        xj[r * sd + p] = (c0 + r < n) ? X[(c0 + r) * d + p] / l : 0.0;
    }
    __syncthreads();
    const double noise = exp(th[d]);
    const int c = threadIdx.x & 127;
    const long long gj = c0 + c;
    if (gj >= ncols_out) return;
    for (int r = threadIdx.x >> 7; r < CGP_TILE; r += 2) {
        const long long gi = r0 + r;
        if (gi >= mrows_out) break;
        double v = 0.0;
        if (gi < m && gj < n) {
            if (self && gi == gj) {
                v = (1.0 + noise) + alpha;
            } else {
                double s = 0.0;
                for (int p = 0; p < d; ++p) {
                    double t = xi[r * sd + p] - xj[c * sd + p];
                    s += t * t;
                }
                v = stationary_from_sq<KIND>(s);
            }
        }
        out[gi * ldo + gj] = v;
    }
}

// K2: per-tile partial of  G_p = sum_ij (a_i a_j - Kinv_ij) dK_ij/dtheta_p  over the lower tiles
// (off-diagonal tiles counted twice), dK recomputed on the fly from X (SK kernels.py:1753,1768 / :1581-1584)
// instead of materialising sklearn's [n,n,P] tensor.  gpart [n_mats][ntiles_max][DMAX].
// grid (Tmax(Tmax+1)/2, n_mats), block 256.  smem: 2*128*(d|1) + 2*128 + 8*DMAX doubles.
template <int KIND, int DMAX>
__global__ void __launch_bounds__(256)
k_grad_tiles(const double* __restrict__ Kinv, const double* __restrict__ X, const double* __restrict__ abuf,
             const double* __restrict__ theta, double* __restrict__ gpart, const MatMeta* __restrict__ meta, int d,
             int zstride, int ntiles_max) {
    extern __shared__ __align__(16) double sm[];
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    int ti, tj;
    lower_tile(blockIdx.x, ti, tj);
    if (ti >= T) return;
    const int sd = d | 1;
    double* xi = sm;
    double* xj = xi + CGP_TILE * sd;
    double* ai = xj + CGP_TILE * sd;
    double* aj = ai + CGP_TILE;
    double* red = aj + CGP_TILE;  // [8][DMAX]
    const double* th = theta + (long long)m.theta_idx * (d + 1);
    const double* a = abuf + (long long)blockIdx.y * zstride;
    for (int idx = threadIdx.x; idx < CGP_TILE * d; idx += 256) {
        int r = idx / d, p = idx - r * d;
        int gi = ti * CGP_TILE + r, gj = tj * CGP_TILE + r;
        xi[r * sd + p] = (gi < m.n) ? X[(long long)(m.xoff + gi) * d + p] : 0.0;
        xj[r * sd + p] = (gj < m.n) ? X[(long long)(m.xoff + gj) * d + p] : 0.0;
    }
    if (threadIdx.x < CGP_TILE) {
        ai[threadIdx.x] = a[ti * CGP_TILE + threadIdx.x];
        aj[threadIdx.x] = a[tj * CGP_TILE + threadIdx.x];
    }
    __syncthreads();
    double il2[DMAX], G[DMAX], xc[DMAX];
    const int c = threadIdx.x & 127;
    const int gj = tj * CGP_TILE + c;
#pragma unroll
    for (int p = 0; p < DMAX; ++p) {
        G[p] = 0.0;
        double l = (p < d) ? exp(th[p]) : 1.0;
        il2[p] = 1.0 / (l * l);
        xc[p] = (p < d) ? xj[c * sd + p] : 0.0;
    }
    const double* Kt = Kinv + m.a_off;
    const long long ld = m.npad;
    if (gj < m.n) {
        const double ajc = aj[c];
        for (int r = threadIdx.x >> 7; r < CGP_TILE; r += 2) {
            const int gi = ti * CGP_TILE + r;
            if (gi >= m.n) break;
            if (gi == gj) continue;
            const double w = ai[r] * ajc - Kt[(long long)gi * ld + gj];
            double Dp[DMAX], s = 0.0;
#pragma unroll
            for (int p = 0; p < DMAX; ++p) {
                double t = (p < d) ? xi[r * sd + p] - xc[p] : 0.0;
                Dp[p] = (t * t) * il2[p];
                s += Dp[p];
            }
            double e;
            if (KIND == 0)
                e = 3.0 * exp(-sqrt(3.0 * s));  // SK kernels.py:1768
            else
                e = exp(-0.5 * s);  // SK kernels.py:1581-1584 (K * D)
            const double we = w * e;
#pragma unroll
            for (int p = 0; p < DMAX; ++p) G[p] += we * Dp[p];
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int p = 0; p < DMAX; ++p) {
        double s = warp_sum(G[p]);
        if (lane == 0) red[warp * DMAX + p] = s;
    }
    __syncthreads();
    if (threadIdx.x < DMAX) {
        double s = 0.0;
        for (int w = 0; w < 8; ++w) s += red[w * DMAX + threadIdx.x];
        if (ti != tj) s *= 2.0;
        gpart[((long long)blockIdx.y * ntiles_max + blockIdx.x) * DMAX + threadIdx.x] = s;
    }
}

// grad_p = 1/2 sum_tiles gpart ; grad_noise = 1/2 noise (a^T a - tr Kinv)  (SK kernels.py:1412, _gpr.py:648-651).
// grid (n_mats), block 256
template <int DMAX>
__global__ void __launch_bounds__(256)
k_grad_finish(const double* __restrict__ Kinv, const double* __restrict__ abuf, const double* __restrict__ gpart,
              const double* __restrict__ theta, const int* __restrict__ info, double* __restrict__ grad,
              const MatMeta* __restrict__ meta, int d, int zstride, int ntiles_max) {
    const MatMeta m = meta[blockIdx.x];
    const int P = d + 1;
    __shared__ double red[8];
    const double* a = abuf + (long long)blockIdx.x * zstride;
    const double* Kt = Kinv + m.a_off;
    double s = 0.0;
    for (int i = threadIdx.x; i < m.n; i += 256) s += a[i] * a[i] - Kt[(long long)i * m.npad + i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    const bool bad = info[m.out_idx] != 0;
    if (threadIdx.x == 0) {
        double q = 0.0;
        for (int w = 0; w < 8; ++w) q += red[w];
        const double noise = exp(theta[(long long)m.theta_idx * P + d]);
        grad[(long long)m.out_idx * P + d] = bad ? 0.0 : 0.5 * noise * q;
    }
    if (threadIdx.x < d) {
        const int T = m.npad / CGP_TILE;
        const int nt = T * (T + 1) / 2;
        double g = 0.0;
        for (int t = 0; t < nt; ++t) g += gpart[((long long)blockIdx.x * ntiles_max + t) * DMAX + threadIdx.x];
        grad[(long long)m.out_idx * P + threadIdx.x] = bad ? 0.0 : 0.5 * g;
    }
}
