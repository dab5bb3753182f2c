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

#ifdef DIAG_TIMING  // development probe (tools/dev/diag_timing.cu): clock stamps of thread 0 at phase boundaries
#define DIAG_STAMP(i)                                          \
    do {                                                       \
        if (threadIdx.x == 0 && blockIdx.x == 0) g_diag_t[i] = clock64(); \
    } while (0)
#else
#define DIAG_STAMP(i)
#endif

// The three serial 32-wide routines below are written as template recursions instead of runtime loops so that
// every index of the per-thread arrays is a compile-time constant: the arrays stay in registers and the
// shared-memory operands (broadcast reads at static offsets) are hoisted out of the dependent FMA chains.

// (1) one column step of the in-register 32x32 Cholesky (lane = row).  Software-pipelined: as soon as the first
// trailing column has been updated, the pivot broadcast + rsqrt of the NEXT column are issued, so their latency hides
// behind the remaining shuffle/FMA updates of this column.  (dcc, ip) = pivot of column C and its inverse square root.
template <int C>
struct FactorCol {
    static __device__ __forceinline__ void run(double (&row)[32], int lane, int& failcol, double& dinv, double dcc, double ip) {
        if (!(dcc > 0.0) && failcol < 0) failcol = C;
        const double lc = (lane == C) ? dcc * ip : row[C] * ip;
        row[C] = lc;
        if (lane == C) dinv = ip;
        double dnext = 0.0, ipnext = 0.0;
        if (C + 1 < 32) {
            const double lq = __shfl_sync(0xffffffffu, lc, C + 1);
            if (lane >= C + 1) row[C + 1 < 32 ? C + 1 : 31] -= lc * lq;
            dnext = __shfl_sync(0xffffffffu, row[C + 1 < 32 ? C + 1 : 31], C + 1 < 32 ? C + 1 : 31);
            ipnext = rsqrt(dnext);
        }
#pragma unroll
        for (int q = C + 2; q < 32; ++q) {
            const double lq = __shfl_sync(0xffffffffu, lc, q);
            if (lane >= q) row[q] -= lc * lq;
        }
        FactorCol<C + 1>::run(row, lane, failcol, dinv, dnext, ipnext);
    }
};
template <>
struct FactorCol<32> {
    static __device__ __forceinline__ void run(double (&)[32], int, int&, double&, double, double) {}
};

// (2) forward substitution of one row against the 32x32 factor Lp (row stride DIAG_LD): x <- x Lp^-T
template <int C>
struct PanelCol {
    static __device__ __forceinline__ void run(double (&x)[32], const double* __restrict__ Lp, const double* __restrict__ rinv) {
        double s0 = x[C], s1 = 0.0;
#pragma unroll
        for (int q = 0; q < C; ++q) {
            if (q & 1)
                s1 -= x[q] * Lp[C * DIAG_LD + q];
            else
                s0 -= x[q] * Lp[C * DIAG_LD + q];
        }
        x[C] = (s0 + s1) * rinv[C];
        PanelCol<C + 1>::run(x, Lp, rinv);
    }
};
template <>
struct PanelCol<32> {
    static __device__ __forceinline__ void run(double (&)[32], const double*, const double*) {}
};

// (3) column `lane` of the inverse of the 32x32 factor Lp: x_i = (delta_{i,lane} - sum_{k<i} Lp[i][k] x_k) / Lp[i][i]
template <int I>
struct InvRow {
    static __device__ __forceinline__ void run(double (&x)[32], const double* __restrict__ Lp, const double* __restrict__ rinv, int lane) {
        double s0 = (I == lane) ? 1.0 : 0.0, s1 = 0.0;
#pragma unroll
        for (int k = 0; k < I; ++k) {
            if (k & 1)
                s1 -= Lp[I * DIAG_LD + k] * x[k];
            else
                s0 -= Lp[I * DIAG_LD + k] * x[k];
        }
        x[I] = (s0 + s1) * rinv[I];
        InvRow<I + 1>::run(x, Lp, rinv, lane);
    }
};
template <>
struct InvRow<32> {
    static __device__ __forceinline__ void run(double (&)[32], const double*, const double*, int) {}
};

// rank-32 update of the trailing block of order 16*MB that starts at row/col t0, from the panel columns [k0, k0+32).
// 16x16 thread grid, thread (ty, tx) owns rows ty + 16 i, cols tx + 16 j: row reads are 2-address broadcasts, column
// reads walk 16 consecutive rows (stride 129 doubles) -> no bank conflicts.  Only 16x16 blocks with i >= j are touched.
template <int MB>
__device__ __forceinline__ void trail_update(double* __restrict__ S, int t0, int k0, int tid) {
    const int ty = tid >> 4, tx = tid & 15;
    double acc[MB][MB];
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int jj = 0; jj < MB; ++jj) acc[i][jj] = 0.0;
#pragma unroll 2
    for (int c = 0; c < 32; ++c) {
        double ra[MB], rb[MB];
#pragma unroll
        for (int i = 0; i < MB; ++i) {
            ra[i] = S[(t0 + ty + 16 * i) * DIAG_LD + k0 + c];
            rb[i] = S[(t0 + tx + 16 * i) * DIAG_LD + k0 + c];
        }
#pragma unroll
        for (int i = 0; i < MB; ++i)
#pragma unroll
            for (int jj = 0; jj <= i; ++jj) acc[i][jj] += ra[i] * rb[jj];
    }
#pragma unroll
    for (int i = 0; i < MB; ++i)
#pragma unroll
        for (int jj = 0; jj <= i; ++jj) S[(t0 + ty + 16 * i) * DIAG_LD + t0 + tx + 16 * jj] -= acc[i][jj];
}

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
    DIAG_STAMP(0);
#pragma unroll 8
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {  // 64 asynchronous 8-byte copies in flight per thread
        int r = idx >> 7, c = idx & 127;
        cp_async8(S + r * DIAG_LD + c, tile + (long long)r * ld + c);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    DIAG_STAMP(1);

    // ---------------- potrf, blocked by 32
#pragma unroll 1
    for (int k0 = 0; k0 < CGP_TILE; k0 += 32) {
        if (warp == 0) {
            double row[32];
#pragma unroll
            for (int q = 0; q < 32; ++q) row[q] = (q <= lane) ? S[(k0 + lane) * DIAG_LD + k0 + q] : 0.0;
            int failcol = -1;
            double dinv = 0.0;
            const double d0 = __shfl_sync(0xffffffffu, row[0], 0);
            FactorCol<0>::run(row, lane, failcol, dinv, d0, rsqrt(d0));
#pragma unroll
            for (int q = 0; q < 32; ++q) S[(k0 + lane) * DIAG_LD + k0 + q] = (q <= lane) ? row[q] : 0.0;
            rinv[k0 + lane] = dinv;
            if (lane == 0 && failcol >= 0 && fail == 0) fail = k0 + failcol + 1;
        }
        __syncthreads();
        DIAG_STAMP(2 + 3 * (k0 / 32));
        const int mrem = CGP_TILE - k0 - 32;
        if (tid < mrem) {  // sub-panel below the diagonal piece: one thread per row
            double* xr = S + (k0 + 32 + tid) * DIAG_LD + k0;
            double x[32];
#pragma unroll
            for (int c = 0; c < 32; ++c) x[c] = xr[c];
            PanelCol<0>::run(x, S + k0 * DIAG_LD + k0, rinv + k0);
#pragma unroll
            for (int c = 0; c < 32; ++c) xr[c] = x[c];
        }
        __syncthreads();
        DIAG_STAMP(3 + 3 * (k0 / 32));
        if (mrem == 96)
            trail_update<6>(S, k0 + 32, k0, tid);
        else if (mrem == 64)
            trail_update<4>(S, k0 + 32, k0, tid);
        else if (mrem == 32)
            trail_update<2>(S, k0 + 32, k0, tid);
        __syncthreads();
        DIAG_STAMP(4 + 3 * (k0 / 32));
    }

    // ---------------- write L_jj (strict upper zeroed), log-determinant partial, status
#pragma unroll 8
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {
        int r = idx >> 7, c = idx & 127;
        tile[(long long)r * ld + c] = (c <= r) ? S[r * DIAG_LD + c] : 0.0;
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
    DIAG_STAMP(14);

    // ---------------- trtri (lower, in place), blocked by 32
    if (warp < 4) {  // the four diagonal pieces (their strict upper parts were zeroed above): lane = column
        const int k0 = warp * 32;
        double x[32];
        InvRow<0>::run(x, S + k0 * DIAG_LD + k0, rinv + k0, lane);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if (i >= lane) S[(k0 + i) * DIAG_LD + k0 + lane] = x[i];
    }
    __syncthreads();
    DIAG_STAMP(15);
#pragma unroll 1
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
                        wa[i] = S[(32 * a + ti + 8 * i) * DIAG_LD + 32 * ap + k];
                        lb[i] = S[(32 * ap + k) * DIAG_LD + 32 * b + tj + 8 * i];
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
                for (int jj = 0; jj < 4; ++jj) S[(32 * a + ti + 8 * i) * DIAG_LD + 32 * b + tj + 8 * jj] = acc[i][jj];
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
                    ta[i] = S[(32 * a + ti + 8 * i) * DIAG_LD + 32 * b + k];
                    wb[i] = S[(32 * b + k) * DIAG_LD + 32 * b + tj + 8 * i];
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
                for (int jj = 0; jj < 4; ++jj) S[(32 * a + ti + 8 * i) * DIAG_LD + 32 * b + tj + 8 * jj] = -acc[i][jj];
        }
        __syncthreads();
    }

    // D_j = inverse (row-major lower), U_jj = D_j^T.  Only the lower part of S is meaningful here.
    double* D = Dbuf + m.d_off + (long long)j * CGP_TILE * CGP_TILE;
    double* Ut = Ubuf + m.u_off + (long long)j * CGP_TILE * ld + (long long)j * CGP_TILE;
#pragma unroll 8
    for (int idx = tid; idx < CGP_TILE * CGP_TILE; idx += 256) {
        int r = idx >> 7, c = idx & 127;
        D[r * CGP_TILE + c] = (c <= r) ? S[r * DIAG_LD + c] : 0.0;
        Ut[(long long)r * ld + c] = (c >= r) ? S[c * DIAG_LD + r] : 0.0;
    }
    DIAG_STAMP(19);
}

// ---- y solves and the LML value ----------------------------------------------------------------
// z = L^-1 y = U^T y :  z_i = sum_{k<=i} U[k,i] y_k.   grid (Tmax, n_mats), block 1024:
// thread (g = tid>>7, c = tid&127) sums rows k = g, g+8, ... of column c; the 8 partials are added in fixed order.
__global__ void __launch_bounds__(1024)
k_solve_z(const double* __restrict__ Ubuf, const double* __restrict__ y, double* __restrict__ zbuf,
          const MatMeta* __restrict__ meta, int zstride) {
    const MatMeta m = meta[blockIdx.y];
    __shared__ double part[8][CGP_TILE];
    const int c = threadIdx.x & 127, g = threadIdx.x >> 7;
    const int i = blockIdx.x * CGP_TILE + c;
    if (blockIdx.x * CGP_TILE >= m.npad) return;
    const double* U = Ubuf + m.u_off;
    const long long ld = m.npad;
    const double* yy = y + m.xoff;
    double s = 0.0;
    const int kend = min(i, m.n - 1);
#pragma unroll 4
    for (int k = g; k <= kend; k += 8) s += U[(long long)k * ld + i] * yy[k];
    part[g][c] = s;
    __syncthreads();
    if (g == 0) {
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += part[q][c];
        zbuf[(long long)blockIdx.y * zstride + i] = (i < m.n) ? t : 0.0;
    }
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
