// Diagonal-block kernel of the blocked Cholesky: one CTA factorises a 128x128 block held in shared memory,
// writes L_jj, inverts it in place and writes D_j = inv(L_jj) (row-major, the B operand of the panel GEMMs)
// and U_jj = D_j^T (the diagonal block of U = L^-T).  For clusters with n <= 128 this is the whole
// factorisation ("one CTA per matrix").
//
// Both steps are blocked by 32 so that almost all flops run as register-tiled updates between few barriers:
//   potrf: per 32-column sub-block  (1) one warp factors the 32x32 diagonal piece in registers with
//          warp-shuffle column broadcasts, (2) one thread per row solves the sub-panel below it,
//          (3) all threads apply the rank-32 update to the trailing part in 4x4 register tiles;
//   trtri: the four 32x32 diagonal pieces are inverted by four warps (one lane per column), the off-diagonal
//          32x32 pieces follow column-block by column-block as small register-tiled products.
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
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long ld = m.npad;
    double* S = diag_smem;
    double* rinv = diag_smem + CGP_TILE * DIAG_LD;
    double* tile = Abuf + m.a_off + (long long)j * CGP_TILE * ld + (long long)j * CGP_TILE;
    __shared__ int fail;
    if (tid == 0) fail = 0;
    for (int idx = tid; idx < CGP_TILE * CGP_TILE / 2; idx += 256) {
        int r = idx >> 6, c = (idx & 63) * 2;
        double2 v = *reinterpret_cast<const double2*>(tile + (long long)r * ld + c);
        S[r * DIAG_LD + c] = v.x;
        S[r * DIAG_LD + c + 1] = v.y;
    }
    __syncthreads();

    // ---------------- potrf, blocked by 32
    for (int k0 = 0; k0 < CGP_TILE; k0 += 32) {
        if (warp == 0) {
            double row[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) row[q] = (q <= lane) ? S[(k0 + lane) * DIAG_LD + k0 + q] : 0.0;
            int failcol = -1;
            double dinv = 0.0;
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                const double dcc = __shfl_sync(0xffffffffu, row[c], c);
                if (!(dcc > 0.0) && failcol < 0) failcol = c;
                const double ip = rsqrt(dcc);
                const double lc = (lane == c) ? dcc * ip : row[c] * ip;
                row[c] = lc;
                if (lane == c) dinv = ip;
#pragma unroll
                for (int q = c + 1; q < 32; ++q) {
                    const double lq = __shfl_sync(0xffffffffu, lc, q);
                    if (lane >= q) row[q] -= lc * lq;
                }
            }
#pragma unroll
            for (int q = 0; q < 32; ++q)
                if (q <= lane) S[(k0 + lane) * DIAG_LD + k0 + q] = row[q];
            rinv[k0 + lane] = dinv;
            if (lane == 0 && failcol >= 0 && fail == 0) fail = k0 + failcol + 1;
        }
        __syncthreads();
        const int mrem = CGP_TILE - k0 - 32;
        if (tid < mrem) {  // sub-panel: rows below the diagonal piece, forward substitution
            const int r = k0 + 32 + tid;
            double x[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) {
                double sacc = S[r * DIAG_LD + k0 + c];
#pragma unroll
                for (int q = 0; q < c; ++q) sacc -= x[q] * S[(k0 + c) * DIAG_LD + k0 + q];
                x[c] = sacc * rinv[k0 + c];
            }
#pragma unroll
            for (int c = 0; c < 32; ++c) S[r * DIAG_LD + k0 + c] = x[c];
        }
        __syncthreads();
        const int nsub = mrem / 4, cnt = nsub * (nsub + 1) / 2;
        for (int st = tid; st < cnt; st += 256) {  // rank-32 update of the trailing lower part, 4x4 register tiles
            int a = (int)((sqrtf(8.0f * (float)st + 1.0f) - 1.0f) * 0.5f);
            while (a * (a + 1) / 2 > st) --a;
            while ((a + 1) * (a + 2) / 2 <= st) ++a;
            const int b = st - a * (a + 1) / 2;
            const int r0 = k0 + 32 + 4 * a, q0 = k0 + 32 + 4 * b;
            double acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
#pragma unroll 4
            for (int c = 0; c < 32; ++c) {
                double ra[4], rb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ra[i] = S[(r0 + i) * DIAG_LD + k0 + c];
                    rb[i] = S[(q0 + i) * DIAG_LD + k0 + c];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] += ra[i] * rb[jj];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) S[(r0 + i) * DIAG_LD + q0 + jj] -= acc[i][jj];
        }
        __syncthreads();
    }

    // ---------------- zero the strict upper part, write L_jj, log-determinant partial, status
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {
        int r = idx >> 7, c = idx & 127;
        if (c > r) S[r * DIAG_LD + c] = 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < CGP_TILE * CGP_TILE / 2; idx += 256) {
        int r = idx >> 6, c = (idx & 63) * 2;
        double2 v;
        v.x = S[r * DIAG_LD + c];
        v.y = S[r * DIAG_LD + c + 1];
        *reinterpret_cast<double2*>(tile + (long long)r * ld + c) = v;
    }
    if (tid < 32) {
        double sl = 0.0;
#pragma unroll
        for (int q = 0; q < 4; ++q) sl += log(S[(tid + 32 * q) * DIAG_LD + tid + 32 * q]);
        sl = warp_sum(sl);
        if (tid == 0) {
            logdet_part[(long long)blockIdx.x * Tmax + j] = sl;
            if (fail != 0 && info[m.out_idx] == 0) info[m.out_idx] = j * CGP_TILE + fail;
        }
    }
    __syncthreads();

    // ---------------- trtri (lower, in place), blocked by 32
    if (warp < 4) {  // the four diagonal pieces: lane = column, forward substitution against e_lane
        const int k0 = warp * 32;
        double x[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            double sacc = (i == lane) ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < i; ++k) sacc -= S[(k0 + i) * DIAG_LD + k0 + k] * x[k];
            x[i] = sacc * rinv[k0 + i];
        }
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i >= lane) S[(k0 + i) * DIAG_LD + k0 + lane] = x[i];
    }
    __syncthreads();
    for (int b = 2; b >= 0; --b) {
        // phase 1: T[a,b] = sum_{a'=b+1}^{a} W[a,a'] L[a',b]   for a = b+1..3  (W trailing part already inverted)
        const int blk = tid >> 6, sub = tid & 63;
        const int a = b + 1 + blk;
        const int ti = sub >> 3, tj = sub & 7;
        const bool act = a <= 3;
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
        if (act) {
            for (int ap = b + 1; ap <= a; ++ap) {
#pragma unroll 4
                for (int k = 0; k < 32; ++k) {
                    double wa[4], lb[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        wa[i] = S[(32 * a + 4 * ti + i) * DIAG_LD + 32 * ap + k];
                        lb[i] = S[(32 * ap + k) * DIAG_LD + 32 * b + 4 * tj + i];
                    }
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) acc[i][jj] += wa[i] * lb[jj];
                }
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) S[(32 * a + 4 * ti + i) * DIAG_LD + 32 * b + 4 * tj + jj] = acc[i][jj];
        }
        __syncthreads();
        // phase 2: W[a,b] = -T[a,b] W[b,b]
        if (act) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) acc[i][jj] = 0.0;
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                double ta[4], wb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    ta[i] = S[(32 * a + 4 * ti + i) * DIAG_LD + 32 * b + k];
                    wb[i] = S[(32 * b + k) * DIAG_LD + 32 * b + 4 * tj + i];
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) acc[i][jj] += ta[i] * wb[jj];
            }
        }
        __syncthreads();
        if (act) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) S[(32 * a + 4 * ti + i) * DIAG_LD + 32 * b + 4 * tj + jj] = -acc[i][jj];
        }
        __syncthreads();
    }

    double* D = Dbuf + m.d_off + (long long)j * CGP_TILE * CGP_TILE;
    double* Ut = Ubuf + m.u_off + (long long)j * CGP_TILE * ld + (long long)j * CGP_TILE;
    for (int idx = tid; idx < CGP_TILE * CGP_TILE / 2; idx += 256) {
        int r = idx >> 6, c = (idx & 63) * 2;
        double2 v;
        v.x = S[r * DIAG_LD + c];
        v.y = S[r * DIAG_LD + c + 1];
        *reinterpret_cast<double2*>(D + r * CGP_TILE + c) = v;
        double2 u;
        u.x = S[c * DIAG_LD + r];
        u.y = S[(c + 1) * DIAG_LD + r];
        *reinterpret_cast<double2*>(Ut + (long long)r * ld + c) = u;
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
