// Diagonal-block kernel of the blocked Cholesky: one CTA factorises a 128x128 block held in shared
// memory (unblocked right-looking dpotf2), writes L_jj, then inverts it in place (dtrti2) and writes
// D_j = inv(L_jj) (row-major, for the panel GEMMs) and U_jj = D_j^T (the diagonal block of U = L^-T).
// Also the whole factorisation for clusters with n <= 128 ("one CTA per matrix").
#pragma once
#include "common.cuh"

#define DIAG_LD 129
#define DIAG_SMEM_BYTES ((CGP_TILE * DIAG_LD + CGP_TILE) * 8)

extern __shared__ __align__(16) double diag_smem[];

// grid (n_mats), block 256.  logdet_part [n_mats][Tmax], info [n_out]
__global__ void __launch_bounds__(256, 1)
k_potrf_diag(double* __restrict__ Abuf, double* __restrict__ Ubuf, double* __restrict__ Dbuf,
             double* __restrict__ logdet_part, int* __restrict__ info, const MatMeta* __restrict__ meta, int j,
             int Tmax) {
    const MatMeta m = meta[blockIdx.x];
    const int T = m.npad / CGP_TILE;
    if (j >= T) return;
    const int tid = threadIdx.x;
    const long long ld = m.npad;
    double* S = diag_smem;
    double* tmp = diag_smem + CGP_TILE * DIAG_LD;
    double* tile = Abuf + m.a_off + (long long)j * CGP_TILE * ld + (long long)j * CGP_TILE;
    __shared__ int fail;
    if (tid == 0) fail = 0;
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {
        int r = idx >> 7, c = idx & 127;
        S[r * DIAG_LD + c] = tile[(long long)r * ld + c];
    }
    __syncthreads();

    // ---- dpotf2 (lower), right-looking
    const int tx = tid & 15, ty = tid >> 4;
    for (int c = 0; c < CGP_TILE; ++c) {
        if (tid == 0) {
            double p = S[c * DIAG_LD + c];
            if (!(p > 0.0) && fail == 0) fail = c + 1;
            S[c * DIAG_LD + c] = sqrt(p);
        }
        __syncthreads();
        const double piv = S[c * DIAG_LD + c];
        for (int r = c + 1 + tid; r < CGP_TILE; r += 256) S[r * DIAG_LD + c] /= piv;
        __syncthreads();
        for (int r = c + 1 + ty; r < CGP_TILE; r += 16) {
            const double lr = S[r * DIAG_LD + c];
            for (int q = c + 1 + tx; q <= r; q += 16) S[r * DIAG_LD + q] -= lr * S[q * DIAG_LD + c];
        }
        __syncthreads();
    }

    // ---- write L_jj (upper part zeroed), log-determinant partial, status
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {
        int r = idx >> 7, c = idx & 127;
        tile[(long long)r * ld + c] = (c <= r) ? S[r * DIAG_LD + c] : 0.0;
    }
    if (tid < 32) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) s += log(S[(tid + 32 * q) * DIAG_LD + tid + 32 * q]);
        s = warp_sum(s);
        if (tid == 0) {
            logdet_part[(long long)blockIdx.x * Tmax + j] = s;
            if (fail != 0 && info[m.out_idx] == 0) info[m.out_idx] = j * CGP_TILE + fail;
        }
    }

    // ---- dtrti2 (lower, non-unit) in place: columns from last to first
    for (int jj = CGP_TILE - 1; jj >= 0; --jj) {
        __syncthreads();
        if (tid > jj && tid < CGP_TILE) tmp[tid] = S[tid * DIAG_LD + jj];
        const double djj = 1.0 / S[jj * DIAG_LD + jj];
        __syncthreads();
        const int i = jj + 1 + (tid >> 1), h = tid & 1;
        double s = 0.0;
        if (i < CGP_TILE)
            for (int k = jj + 1 + h; k <= i; k += 2) s += S[i * DIAG_LD + k] * tmp[k];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        if (i < CGP_TILE && h == 0) S[i * DIAG_LD + jj] = -s * djj;
        if (tid == 0) S[jj * DIAG_LD + jj] = djj;
    }
    __syncthreads();

    double* D = Dbuf + m.d_off + (long long)j * CGP_TILE * CGP_TILE;
    double* Ut = Ubuf + m.u_off + (long long)j * CGP_TILE * ld + (long long)j * CGP_TILE;
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {
        int r = idx >> 7, c = idx & 127;
        D[r * CGP_TILE + c] = (c <= r) ? S[r * DIAG_LD + c] : 0.0;
        Ut[(long long)r * ld + c] = (c >= r) ? S[c * DIAG_LD + r] : 0.0;
    }
}

// ---- y solves and the LML value ----------------------------------------------------------------
// z = L^-1 y = U^T y :  z_i = sum_{k<=i} U[k,i] y_k.   grid (Tmax, n_mats), block 128 (thread = column)
__global__ void __launch_bounds__(128)
k_solve_z(const double* __restrict__ Ubuf, const double* __restrict__ y, double* __restrict__ zbuf,
          const MatMeta* __restrict__ meta, int zstride) {
    const MatMeta m = meta[blockIdx.y];
    const int i = blockIdx.x * CGP_TILE + threadIdx.x;
    if (i >= m.npad) return;
    const double* U = Ubuf + m.u_off;
    const long long ld = m.npad;
    const double* yy = y + m.xoff;
    double s = 0.0;
    const int kend = min(i, m.n - 1);
    for (int k = 0; k <= kend; ++k) s += U[(long long)k * ld + i] * yy[k];
    zbuf[(long long)blockIdx.y * zstride + i] = (i < m.n) ? s : 0.0;
}

// a = L^-T z = U z :  a_j = sum_{k>=j} U[j,k] z_k.   warp per row; grid (npad_max/8, n_mats), block 256
__global__ void __launch_bounds__(256)
k_solve_a(const double* __restrict__ Ubuf, const double* __restrict__ zbuf, double* __restrict__ abuf,
          const MatMeta* __restrict__ meta, int zstride) {
    const MatMeta m = meta[blockIdx.y];
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= m.npad) return;
    const double* U = Ubuf + m.u_off + (long long)row * m.npad;
    const double* z = zbuf + (long long)blockIdx.y * zstride;
    double s = 0.0;
    for (int k = (row & ~31) + lane; k < m.n; k += 32)
        if (k >= row) s += U[k] * z[k];
    s = warp_sum(s);
    if (lane == 0) abuf[(long long)blockIdx.y * zstride + row] = (row < m.n) ? s : 0.0;
}

// lml = -1/2 z^T z - sum log diag L - n/2 log 2pi   (SK/_gpr.py:613-617).  grid (n_mats), block 256
__global__ void __launch_bounds__(256)
k_lml_finish(const double* __restrict__ zbuf, const double* __restrict__ logdet_part, const int* __restrict__ info,
             double* __restrict__ lml, const MatMeta* __restrict__ meta, int zstride, int Tmax) {
    const MatMeta m = meta[blockIdx.x];
    __shared__ double red[8];
    const double* z = zbuf + (long long)blockIdx.x * zstride;
    double s = 0.0;
    for (int i = threadIdx.x; i < m.n; i += 256) s += z[i] * z[i];
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double q = 0.0;
        for (int w = 0; w < 8; ++w) q += red[w];
        double ld = 0.0;
        const int T = m.npad / CGP_TILE;
        for (int j = 0; j < T; ++j) ld += logdet_part[(long long)blockIdx.x * Tmax + j];
        double v = -0.5 * q - ld - 0.5 * (double)m.n * 1.8378770664093453;  // log(2 pi)
        if (info[m.out_idx] != 0) v = -INFINITY;
        lml[m.out_idx] = v;
    }
}

// out-of-place transpose of the padded square factor: W = U^T.  grid (Tmax*4, Tmax*4, n_mats), block (32, 8)
__global__ void k_transpose(const double* __restrict__ Ubuf, double* __restrict__ Wbuf, const MatMeta* __restrict__ meta) {
    const MatMeta m = meta[blockIdx.z];
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    if (bx >= m.npad || by >= m.npad) return;
    const double* U = Ubuf + m.u_off;
    double* W = Wbuf + m.a_off;
    const long long ld = m.npad;
    for (int r = threadIdx.y; r < 32; r += 8) t[r][threadIdx.x] = U[(long long)(by + r) * ld + bx + threadIdx.x];
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += 8) W[(long long)(bx + r) * ld + by + threadIdx.x] = t[threadIdx.x][r];
}
