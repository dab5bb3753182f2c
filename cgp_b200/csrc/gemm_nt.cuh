// fp64 tensor-core (DMMA) tile GEMM, "NT" form:  acc[128x128] = sum_k A[m,k] * B[n,k]
// Both operands are row-major with k contiguous.  Every blocked step of the factorisation
// (left-looking Cholesky update, panel scaling by an inverted diagonal block, triangular inverse,
// U U^T, and the predictive-variance contraction W k*) is phrased in this one form, so there is a
// single mainloop: cp.async 3-stage ring -> swizzled shared tiles -> LDS.128 fragment loads ->
// mma.sync.m8n8k4.f64.
#pragma once
#include "common.cuh"

// One pipeline stage: A tile [BM][BK] + B tile [BN][BK] doubles (BK*8 bytes per row in 16-byte chunks).
// Chunk c of row r lives at physical chunk c ^ ((r & 1) << 2): the 8 lanes of a quarter-warp
// (rows r, r+1; 4 chunks each) then cover 128 distinct bytes -> conflict-free LDS.128.
//
// The mainloop is templated on the CTA tile BM x BN.  128x128 (8 warps of 64x32, 1 CTA/SM) is the throughput shape.
// 64x64 (8 warps of 32x16, 2 CTAs/SM) and 64x128 compute a 128x128 tile as four / two independent pieces: every
// output element still accumulates exactly the same k sequence, so the result is BIT-IDENTICAL to the 128x128 path.
// The small shapes are chosen when a launch has too few tiles to fill the SMs (the tail of the L-BFGS lock-step,
// small batches per GPU), where they cut the latency of each dependent step of the blocked factorisation.
// In-place steps (tile <- tile * D^T) use 64x128: a piece then owns whole rows of the tile it overwrites.
#define GEMM_SMEM_BYTES_T(BM, BN) (CGP_STAGES * ((BM) + (BN)) * CGP_BK * 8)
#define GEMM_SMEM_BYTES GEMM_SMEM_BYTES_T(CGP_TILE, CGP_TILE)

__device__ __forceinline__ int gemm_swz(int row, int chunk) { return chunk ^ ((row & 1) << 2); }

template <int ROWS>
__device__ __forceinline__ void gemm_load_rows(double* sm, const double* __restrict__ g, long long ld, int k0, int tid) {
    constexpr int CPR = CGP_BK / 2;  // 16-byte chunks per tile row
#pragma unroll
    for (int i = 0; i < ROWS * CPR / CGP_GEMM_THREADS; ++i) {
        int idx = tid + i * CGP_GEMM_THREADS;
        int row = idx / CPR, c = idx % CPR;
        cp_async16(sm + row * CGP_BK + gemm_swz(row, c) * 2, g + (long long)row * ld + k0 + c * 2);
    }
}

// Same, but tile row r is row ridx[r] of g (ridx in shared memory): the candidate tiles of the pruned EI scan are
// gathered from the surviving rows of the cross-kernel block without compacting it.
template <int ROWS>
__device__ __forceinline__ void gemm_load_rows_gather(double* sm, const double* __restrict__ g, long long ld,
                                                      const int* __restrict__ ridx, int k0, int tid) {
    constexpr int CPR = CGP_BK / 2;
#pragma unroll
    for (int i = 0; i < ROWS * CPR / CGP_GEMM_THREADS; ++i) {
        int idx = tid + i * CGP_GEMM_THREADS;
        int row = idx / CPR, c = idx % CPR;
        cp_async16(sm + row * CGP_BK + gemm_swz(row, c) * 2, g + (long long)ridx[row] * ld + k0 + c * 2);
    }
}

// acc[mt][nt][e]: row = wm*(BM/2) + mt*8 + g, col = wn*(BN/4) + nt*8 + 2*t + e   (g = lane>>2, t = lane&3)
template <int BM, int BN>
struct GemmAccT {
    static constexpr int MT = BM / 16, NT = BN / 32;
    double v[MT][NT][2];
};
typedef GemmAccT<CGP_TILE, CGP_TILE> GemmAcc;

template <int BM, int BN>
__device__ __forceinline__ void gemm_zero(GemmAccT<BM, BN>& acc) {
#pragma unroll
    for (int mt = 0; mt < GemmAccT<BM, BN>::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < GemmAccT<BM, BN>::NT; ++nt) acc.v[mt][nt][0] = acc.v[mt][nt][1] = 0.0;
}

// K must be a positive multiple of CGP_BK.  smem: GEMM_SMEM_BYTES_T(BM, BN), 16-byte aligned.
template <int BM, int BN, bool GATHER_B = false>
__device__ __forceinline__ void gemm_mainloop(GemmAccT<BM, BN>& acc, double* smem, const double* __restrict__ gA,
                                              long long lda, const double* __restrict__ gB,
                                              long long ldb, int K, const int* __restrict__ bidx = nullptr) {
    constexpr int MT = GemmAccT<BM, BN>::MT, NT = GemmAccT<BM, BN>::NT;
    constexpr int A_ELEMS = BM * CGP_BK, B_ELEMS = BN * CGP_BK;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 1, wn = warp >> 1;
    const int g = lane >> 2, t = lane & 3;
    double* sA = smem;
    double* sB = smem + CGP_STAGES * A_ELEMS;
    const int nk = K / CGP_BK;

#pragma unroll
    for (int s = 0; s < CGP_STAGES - 1; ++s) {
        if (s < nk) {
            gemm_load_rows<BM>(sA + s * A_ELEMS, gA, lda, s * CGP_BK, tid);
            if (GATHER_B)
                gemm_load_rows_gather<BN>(sB + s * B_ELEMS, gB, ldb, bidx, s * CGP_BK, tid);
            else
                gemm_load_rows<BN>(sB + s * B_ELEMS, gB, ldb, s * CGP_BK, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; ++kt) {
        cp_async_wait<CGP_STAGES - 2>();
        __syncthreads();
        {
            int nx = kt + CGP_STAGES - 1;
            if (nx < nk) {
                int slot = nx % CGP_STAGES;
                gemm_load_rows<BM>(sA + slot * A_ELEMS, gA, lda, nx * CGP_BK, tid);
                if (GATHER_B)
                    gemm_load_rows_gather<BN>(sB + slot * B_ELEMS, gB, ldb, bidx, nx * CGP_BK, tid);
                else
                    gemm_load_rows<BN>(sB + slot * B_ELEMS, gB, ldb, nx * CGP_BK, tid);
            }
            cp_async_commit();
        }
        const double* a_s = sA + (kt % CGP_STAGES) * A_ELEMS;
        const double* b_s = sB + (kt % CGP_STAGES) * B_ELEMS;
#pragma unroll
        for (int kk = 0; kk < CGP_BK / 8; ++kk) {
            double2 af[MT], bf[NT];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                int row = wm * (BM / 2) + mt * 8 + g;
                af[mt] = *reinterpret_cast<const double2*>(a_s + row * CGP_BK + gemm_swz(row, kk * 4 + t) * 2);
            }
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) {
                int row = wn * (BN / 4) + nt * 8 + g;
                bf[nt] = *reinterpret_cast<const double2*>(b_s + row * CGP_BK + gemm_swz(row, kk * 4 + t) * 2);
            }
            // k-slot s of the first DMMA is k = 2s, of the second k = 2s+1 (same permutation in A and B)
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    dmma884(acc.v[mt][nt][0], acc.v[mt][nt][1], af[mt].x, bf[nt].x);
                    dmma884(acc.v[mt][nt][0], acc.v[mt][nt][1], af[mt].y, bf[nt].y);
                }
        }
    }
    cp_async_wait<0>();
    __syncthreads();
}

// ---- epilogues -------------------------------------------------------------------------------
// C[row, col] (row-major, ldc) op= acc.   MODE 0: C = acc   1: C = -acc   2: C -= acc   3: C += acc
template <int MODE, int BM, int BN>
__device__ __forceinline__ void gemm_store(const GemmAccT<BM, BN>& acc, double* __restrict__ C, long long ldc) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 1, wn = warp >> 1, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int mt = 0; mt < GemmAccT<BM, BN>::MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < GemmAccT<BM, BN>::NT; ++nt) {
            int row = wm * (BM / 2) + mt * 8 + g, col = wn * (BN / 4) + nt * 8 + 2 * t;
            double2* p = reinterpret_cast<double2*>(C + (long long)row * ldc + col);
            double2 o;
            if (MODE == 0) {
                o.x = acc.v[mt][nt][0];
                o.y = acc.v[mt][nt][1];
            } else if (MODE == 1) {
                o.x = -acc.v[mt][nt][0];
                o.y = -acc.v[mt][nt][1];
            } else if (MODE == 2) {
                o = *p;
                o.x -= acc.v[mt][nt][0];
                o.y -= acc.v[mt][nt][1];
            } else {
                o = *p;
                o.x += acc.v[mt][nt][0];
                o.y += acc.v[mt][nt][1];
            }
            *p = o;
        }
}

// out[col] = sum over the 128 rows of acc^2 (fixed summation order -> bit-reproducible).
// red: shared scratch of 2*128 doubles.
__device__ __forceinline__ void gemm_colsumsq(const GemmAcc& acc, double* red, double* __restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = warp & 1, wn = warp >> 1, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            double s = 0.0;
#pragma unroll
            for (int mt = 0; mt < 8; ++mt) s += acc.v[mt][nt][e] * acc.v[mt][nt][e];
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            s += __shfl_xor_sync(0xffffffffu, s, 8);
            s += __shfl_xor_sync(0xffffffffu, s, 16);
            if (g == 0) red[wm * CGP_TILE + wn * 32 + nt * 8 + 2 * t + e] = s;
        }
    __syncthreads();
    if (threadIdx.x < CGP_TILE) out[threadIdx.x] = red[threadIdx.x] + red[CGP_TILE + threadIdx.x];
}

// ---- kernels ---------------------------------------------------------------------------------
extern __shared__ __align__(16) double gemm_smem[];

#define GEMM_BOUNDS(BM, BN) __launch_bounds__(CGP_GEMM_THREADS, ((BM) + (BN) <= 128) ? 2 : 1)
#define GEMM_PIECES(BM, BN) ((CGP_TILE / (BM)) * (CGP_TILE / (BN)))  // CTAs per 128x128 tile

// One 128x128 tile product of the blocked kernels, or one BM x BN piece of it: blockIdx.x = tile * PIECES + piece.
//   acc = A[piece rows, K] * B[piece rows, K]^T ; C[piece] op= acc     (pointers are those of the whole tile)
template <int MODE, int BM, int BN>
__device__ __forceinline__ void gemm_tile(const double* __restrict__ gA, long long lda, const double* __restrict__ gB,
                                          long long ldb, int K, double* __restrict__ C, long long ldc, int piece) {
    const int qi = piece / (CGP_TILE / BN), qj = piece % (CGP_TILE / BN);
    GemmAccT<BM, BN> acc;
    gemm_zero(acc);
    gemm_mainloop<BM, BN>(acc, gemm_smem, gA + (long long)qi * BM * lda, lda, gB + (long long)qj * BN * ldb, ldb, K);
    gemm_store<MODE, BM, BN>(acc, C + (long long)qi * BM * ldc + (long long)qj * BN, ldc);
}

// Rows >= n_c of the last tile row are the identity pad: L and A are zero there off the diagonal, so a Cholesky update
// of those rows subtracts exact zeros.  When the lower 64 rows of tile row i are all pad (n_c <= 128 i + 64), only the
// upper 64 rows of the tile are computed: the 128x128 CTA runs the 64x128 piece 0, the 64x64 pieces with qi = 1 exit.
// Every computed element accumulates the same k sequence -> same bits.  (Config 3: six of eight clusters pad 4099..4173
// rows to 4224, i.e. their 33rd tile row is 40-98 % pad; it takes part in 9 % of the trailing-update tiles.)
template <int MODE, int BM, int BN>
__device__ __forceinline__ void gemm_tile_rows(const MatMeta& m, int i, const double* __restrict__ gA, long long lda,
                                               const double* __restrict__ gB, long long ldb, int K, double* __restrict__ C,
                                               long long ldc, int piece) {
    if (m.n <= i * CGP_TILE + 64) {
        if (BM == CGP_TILE) {
            static_assert(BM != CGP_TILE || BN == CGP_TILE, "the 128-row shape is 128x128");
            gemm_tile<MODE, 64, CGP_TILE>(gA, lda, gB, ldb, K, C, ldc, 0);
            return;
        }
        if (piece / (CGP_TILE / BN) == 1) return;
    }
    gemm_tile<MODE, BM, BN>(gA, lda, gB, ldb, K, C, ldc, piece);
}

// Plain C = A B^T (measurement + tests).  grid (N/128, M/128).
__global__ void __launch_bounds__(CGP_GEMM_THREADS, 1)
k_dgemm_nt(const double* __restrict__ A, const double* __restrict__ B, double* __restrict__ C, long long M,
           long long N, int K) {
    gemm_tile<0, 128, 128>(A + (long long)blockIdx.y * CGP_TILE * K, K, B + (long long)blockIdx.x * CGP_TILE * K, K, K,
                           C + (long long)blockIdx.y * CGP_TILE * N + (long long)blockIdx.x * CGP_TILE, N, 0);
}

// ---- blocked Cholesky, hybrid schedule --------------------------------------------------------------
// Block columns are grouped in panels of CGP_PANEL tiles.  Inside a panel the factorisation is left-looking
// (k_potrf_update with a short K), after a panel every trailing tile receives ONE rank-(128*CGP_PANEL) update
// (k_potrf_trail).  The long left-looking K loops of a pure left-looking schedule put up to 4096-deep tile
// products on the critical path and leave whole waves almost empty; this schedule keeps the critical path at
// <= 3 short tile products per column while the bulk of the flops runs in large, uniform launches.
// The summation order of every tile is fixed by (n, CGP_PANEL) alone -> results do not depend on the batch.
// All kernels: gridDim.x = (number of tiles) * GEMM_PIECES(BM, BN), gridDim.y = n_mats.

// Column update inside a panel:  A[i,j] -= sum_{k=kb}^{j-1} L[i,k] L[j,k]^T  for i >= j.   tiles: Tmax - j
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_potrf_update(double* __restrict__ Abuf, const MatMeta* __restrict__ meta, int j, int kb) {
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    const int i = j + blockIdx.x / GEMM_PIECES(BM, BN);
    if (i >= T) return;
    double* A = Abuf + m.a_off;
    const long long ld = m.npad;
    gemm_tile_rows<2, BM, BN>(m, i, A + (long long)i * CGP_TILE * ld + (long long)kb * CGP_TILE, ld,
                              A + (long long)j * CGP_TILE * ld + (long long)kb * CGP_TILE, ld, (j - kb) * CGP_TILE,
                              A + (long long)i * CGP_TILE * ld + (long long)j * CGP_TILE, ld, blockIdx.x % GEMM_PIECES(BM, BN));
}

// Trailing update after panel [c0, c1):  A[i,k] -= sum_{c=c0}^{c1-1} L[i,c] L[k,c]^T  for t0 <= k <= i < T  (t0 >= c1).
// tiles: Tt(Tt+1)/2 with Tt = Tmax - t0.  The update is issued in two launches (look-ahead): the block columns of the
// NEXT panel first (k_potrf_head, on the critical path of the factorisation), the rest (this kernel, t0 = start of
// the panel after next) on a bulk stream where it overlaps the latency-bound diag/panel chain of the next panel.
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_potrf_trail(double* __restrict__ Abuf, const MatMeta* __restrict__ meta, int c0, int c1, int t0) {
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    int idx = blockIdx.x / GEMM_PIECES(BM, BN), a = 0;
    while (idx > a) {
        idx -= a + 1;
        ++a;
    }
    const int i = t0 + a, k = t0 + idx;
    if (i >= T) return;
    double* A = Abuf + m.a_off;
    const long long ld = m.npad;
    gemm_tile_rows<2, BM, BN>(m, i, A + (long long)i * CGP_TILE * ld + (long long)c0 * CGP_TILE, ld,
                              A + (long long)k * CGP_TILE * ld + (long long)c0 * CGP_TILE, ld, (c1 - c0) * CGP_TILE,
                              A + (long long)i * CGP_TILE * ld + (long long)k * CGP_TILE, ld, blockIdx.x % GEMM_PIECES(BM, BN));
}

// Same update restricted to the block columns k in [c1, c1 + H) (the next panel), rows i >= k.
// tiles: H * (Tmax - c1), tile = q * (Tmax - c1) + r  ->  k = c1 + q, i = k + r
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_potrf_head(double* __restrict__ Abuf, const MatMeta* __restrict__ meta, int c0, int c1, int Tmax) {
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    const int w = Tmax - c1;
    const int tile = blockIdx.x / GEMM_PIECES(BM, BN);
    const int k = c1 + tile / w, i = k + tile % w;
    if (i >= T) return;
    double* A = Abuf + m.a_off;
    const long long ld = m.npad;
    gemm_tile_rows<2, BM, BN>(m, i, A + (long long)i * CGP_TILE * ld + (long long)c0 * CGP_TILE, ld,
                              A + (long long)k * CGP_TILE * ld + (long long)c0 * CGP_TILE, ld, (c1 - c0) * CGP_TILE,
                              A + (long long)i * CGP_TILE * ld + (long long)k * CGP_TILE, ld, blockIdx.x % GEMM_PIECES(BM, BN));
}

// Panel: L[i,j] = A[i,j] * inv(L[j,j])^T for i > j  (triangular solve recast as a GEMM with the explicitly inverted
// 128x128 diagonal block D_j, row-major in Dbuf).  In place -> BN must be 128 (a piece owns whole rows).  tiles: Tmax-j-1
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_potrf_panel(double* __restrict__ Abuf, const double* __restrict__ Dbuf, const MatMeta* __restrict__ meta, int j) {
    static_assert(BN == CGP_TILE, "in-place tile product: pieces must own whole rows");
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    const int i = j + 1 + blockIdx.x / GEMM_PIECES(BM, BN);
    if (i >= T) return;
    const long long ld = m.npad;
    double* tile = Abuf + m.a_off + (long long)i * CGP_TILE * ld + (long long)j * CGP_TILE;
    const double* D = Dbuf + m.d_off + (long long)j * CGP_TILE * CGP_TILE;
    gemm_tile<0, BM, BN>(tile, ld, D, CGP_TILE, CGP_TILE, tile, ld, blockIdx.x % GEMM_PIECES(BM, BN));
}

// ---- triangular inverse, stored transposed: U = L^-T (upper, row-major), same hybrid schedule ------------
//   U[j,i] = -( sum_{k=j}^{i-1} U[j,k] L[i,k]^T ) D_i^T   for j < i,   U[i,i] = D_i^T.
// The sum over k is accumulated in place in the U[j,i] tile: contributions of a finished panel [c0, c1) are
// pushed to every later column (k_trtri_trail), the columns of the current panel add their short inner part
// (k_trtri_acc), then the tile is scaled by -D_i^T (k_trtri_scale).  The first contribution of a tile (the one
// containing k = j) stores, later ones accumulate.

// inner part of column i of panel [c0, ..):  rows j < i, k from max(j, c0) to i-1.   tiles: i
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_trtri_acc(const double* __restrict__ Abuf, double* __restrict__ Ubuf, const MatMeta* __restrict__ meta, int i, int c0) {
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    const int j = blockIdx.x / GEMM_PIECES(BM, BN), piece = blockIdx.x % GEMM_PIECES(BM, BN);
    const int kb = max(j, c0);
    if (i >= T || j >= i || kb >= i) return;
    const long long ld = m.npad;
    const double* L = Abuf + m.a_off;
    double* U = Ubuf + m.u_off;
    const double* gA = U + (long long)j * CGP_TILE * ld + (long long)kb * CGP_TILE;
    const double* gB = L + (long long)i * CGP_TILE * ld + (long long)kb * CGP_TILE;
    double* out = U + (long long)j * CGP_TILE * ld + (long long)i * CGP_TILE;
    if (j >= c0)
        gemm_tile<0, BM, BN>(gA, ld, gB, ld, (i - kb) * CGP_TILE, out, ld, piece);
    else
        gemm_tile<3, BM, BN>(gA, ld, gB, ld, (i - kb) * CGP_TILE, out, ld, piece);
}

// push of the finished panel [c0, c1) into every later column i >= c1, rows j < c1.
// tiles: c1 * (Tmax - c1), tile = j * (Tmax - c1) + (i - c1)
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_trtri_trail(const double* __restrict__ Abuf, double* __restrict__ Ubuf, const MatMeta* __restrict__ meta, int c0, int c1,
              int Tmax) {
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    const int w = Tmax - c1;
    const int tile = blockIdx.x / GEMM_PIECES(BM, BN), piece = blockIdx.x % GEMM_PIECES(BM, BN);
    const int j = tile / w, i = c1 + tile % w;
    if (i >= T) return;
    const int kb = max(j, c0);
    const long long ld = m.npad;
    const double* L = Abuf + m.a_off;
    double* U = Ubuf + m.u_off;
    const double* gA = U + (long long)j * CGP_TILE * ld + (long long)kb * CGP_TILE;
    const double* gB = L + (long long)i * CGP_TILE * ld + (long long)kb * CGP_TILE;
    double* out = U + (long long)j * CGP_TILE * ld + (long long)i * CGP_TILE;
    if (j >= c0)
        gemm_tile<0, BM, BN>(gA, ld, gB, ld, (c1 - kb) * CGP_TILE, out, ld, piece);
    else
        gemm_tile<3, BM, BN>(gA, ld, gB, ld, (c1 - kb) * CGP_TILE, out, ld, piece);
}

// U[j,i] <- -U[j,i] D_i^T, in place -> BN must be 128.   tiles: i
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_trtri_scale(double* __restrict__ Ubuf, const double* __restrict__ Dbuf, const MatMeta* __restrict__ meta, int i) {
    static_assert(BN == CGP_TILE, "in-place tile product: pieces must own whole rows");
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    const int j = blockIdx.x / GEMM_PIECES(BM, BN);
    if (i >= T || j >= i) return;
    const long long ld = m.npad;
    double* tile = Ubuf + m.u_off + (long long)j * CGP_TILE * ld + (long long)i * CGP_TILE;
    const double* D = Dbuf + m.d_off + (long long)i * CGP_TILE * CGP_TILE;
    gemm_tile<1, BM, BN>(tile, ld, D, CGP_TILE, CGP_TILE, tile, ld, blockIdx.x % GEMM_PIECES(BM, BN));
}

// K^-1 = U U^T, lower tiles (i >= j):  Kinv[i,j] = sum_{k>=i} U[i,k] U[j,k]^T, written over A.
// tiles: Tmax(Tmax+1)/2, enumerated i-major so the longest K (i = 0) starts first.
template <int BM, int BN>
__global__ void GEMM_BOUNDS(BM, BN)
k_lauum(double* __restrict__ Abuf, const double* __restrict__ Ubuf, const MatMeta* __restrict__ meta) {
    const MatMeta m = meta[blockIdx.y];
    const int T = m.npad / CGP_TILE;
    int idx = blockIdx.x / GEMM_PIECES(BM, BN), i = 0;
    while (idx > i) {
        idx -= i + 1;
        ++i;
    }
    const int j = idx;
    if (i >= T) return;
    const long long ld = m.npad;
    const double* U = Ubuf + m.u_off;
    // columns k >= n of U are the identity pad: zero in every real row, so whole k-chunks of pad are skipped (adds of
    // exact zeros: same bits in the n x n part, which is all the gradient kernels read)
    const int trim = (m.npad - m.n) / CGP_BK * CGP_BK;
    gemm_tile<0, BM, BN>(U + (long long)i * CGP_TILE * ld + (long long)i * CGP_TILE, ld,
                         U + (long long)j * CGP_TILE * ld + (long long)i * CGP_TILE, ld, (T - i) * CGP_TILE - trim,
                         Abuf + m.a_off + (long long)i * CGP_TILE * ld + (long long)j * CGP_TILE, ld,
                         blockIdx.x % GEMM_PIECES(BM, BN));
}

// Predictive variance contraction for one cluster and one candidate chunk:
//   V[ib-block rows, cand tile] = sum_{k <= ib} W[ib, k] Kstar[cand, k]^T ;  vpart[ib][cand] = sum_rows V^2
// grid (n_cand_tiles * T); row blocks with the longest K first inside each candidate tile so that
// consecutive CTAs reuse the same Kstar tile from L2.
__global__ void __launch_bounds__(CGP_GEMM_THREADS, 1)
k_predict_vnorm(const double* __restrict__ W, const double* __restrict__ Kstar, double* __restrict__ vpart, int npad,
                long long cand_stride /* rows in vpart */) {
    const int T = npad / CGP_TILE;
    const int ct = blockIdx.x / T;
    const int ib = T - 1 - (blockIdx.x % T);
    GemmAcc acc;
    gemm_zero(acc);
    gemm_mainloop<CGP_TILE, CGP_TILE>(acc, gemm_smem, W + (long long)ib * CGP_TILE * npad, npad,
                                      Kstar + (long long)ct * CGP_TILE * npad, npad, (ib + 1) * CGP_TILE);
    gemm_colsumsq(acc, gemm_smem, vpart + (long long)ib * cand_stride + (long long)ct * CGP_TILE);
}

// Pruned scan: the same contraction for the SURVIVING candidates of a chunk only.  sel[0..*count) lists their rows in
// Kstar (ascending); candidate tile ct holds survivors ct*128..ct*128+127 (the tail of the last tile repeats row
// sel[0]; its results are never read).  The grid is sized for the whole chunk, tiles past the survivor count exit:
// no host synchronisation is needed to size the launch.  Every surviving candidate accumulates exactly the k sequence
// of k_predict_vnorm, so its variance (and score) has the same bits as in the exhaustive scan.
__global__ void __launch_bounds__(CGP_GEMM_THREADS, 1)
k_predict_vnorm_sel(const double* __restrict__ W, const double* __restrict__ Kstar, double* __restrict__ vpart, int npad,
                    long long cand_stride, const int* __restrict__ sel, const int* __restrict__ count) {
    const int T = npad / CGP_TILE;
    const int ct = blockIdx.x / T;
    const int cnt = *count;
    if (ct * CGP_TILE >= cnt) return;
    const int ib = T - 1 - (blockIdx.x % T);
    __shared__ int ridx[CGP_TILE];
    if (threadIdx.x < CGP_TILE) {
        const int q = ct * CGP_TILE + threadIdx.x;
        ridx[threadIdx.x] = sel[q < cnt ? q : 0];
    }
    __syncthreads();
    GemmAcc acc;
    gemm_zero(acc);
    gemm_mainloop<CGP_TILE, CGP_TILE, true>(acc, gemm_smem, W + (long long)ib * CGP_TILE * npad, npad, Kstar, npad,
                                            (ib + 1) * CGP_TILE, ridx);
    gemm_colsumsq(acc, gemm_smem, vpart + (long long)ib * cand_stride + (long long)ct * CGP_TILE);
}
