// K7, large form: k-NN routing with a tensor-core PREFILTER and an exact fp64 finish (bit-exact labels).
//
// Brute-force routing is M x n distance evaluations; in fp64 (knn.cuh) it is bound by the fp64 ALU rate (3 d un-fused
// operations per pair).  Here the ranking quantity
//        s(q, x) = |x|^2 - 2 q.x          ( = |q - x|^2 - |q|^2 : same order of the x for a fixed q )
// is evaluated on the tensor cores (mma.sync.m16n8k8 TF32, fp32 accumulate) for every pair, with both operands split in
// two TF32 parts (v = v_hi + v_lo, 3 MMAs: hi.hi + hi.lo + lo.hi), so that the absolute error of the computed s~ is
// bounded by  eps_q = 2^-17 (max_j |x_j|^2 + 2 sum_i |q_i| max_j |x_ji|)   (analysis in DESIGN.md §3: 27 fp32
// accumulation roundings + operand representation <= 2^-19.2 of that scale; the factor 4.5 is safety margin).
// Per query the k+1 smallest s~ of each of the four lanes that share a query row are kept WITH their indices.  Then
//   tau = k-th smallest s~ over all points          (exact: every lane kept its k+1 smallest)
//   survivors = kept points with s~ <= tau + 2 eps  (every true k-nearest neighbour is among them: its exact s is
//               <= the exact k-th smallest <= tau + eps, hence its s~ <= tau + 2 eps)
// and the survivors are re-ranked by the EXACT un-fused fp64 distance of knn.cuh, ties to the lowest training index,
// i.e. exactly scikit-learn's neighbour set (SK/metrics/_dist_metrics.pxd.tp:44-49, SK/utils/_heap.pyx:46-47).
// A query is handed to the exact brute-force kernel instead (out[q] = -1) when the prefilter cannot certify the set:
// a lane whose list is full and whose largest kept value is itself a survivor may have dropped survivors (many points
// within 2 eps of the k-th distance: ties on integer grids), or a survivor's s~ is further than eps_q from its exact
// value (self-check of the error bound on every survivor).
#pragma once
#include "common.cuh"

#define KTC_KP 4          // kept per lane = k + 1 for k <= 3
#define KTC_TILE 256      // training points per shared-memory stage
#define KTC_QPB 128       // queries per CTA: 8 warps x 16 rows
#define KTC_EPS_SCALE 7.62939453125e-06f  // 2^-17

__device__ __forceinline__ unsigned f2tf32(float v) {
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;\n" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// ---- preparation: split training set + norms ----------------------------------------------------------------------
// Xp[j][p] = (hi, lo) TF32 split of (float) x_jp  (0 for p >= d and for the pad rows j >= n), xn[j] = (float)|x_j|^2
// (pad rows: 3e38 -> never kept).  grid ceil(npad/256), block 256.
template <int DP>
__global__ void __launch_bounds__(256)
k_knn_prep(const double* __restrict__ X, int n, int npad, int d, float2* __restrict__ Xp, float* __restrict__ xn) {
    const int j = blockIdx.x * 256 + threadIdx.x;
    if (j >= npad) return;
    double nn = 0.0;
#pragma unroll
    for (int p = 0; p < DP; ++p) {
        float v = 0.f;
        if (j < n && p < d) {
            const double x = X[(long long)j * d + p];
            nn += x * x;
            v = (float)x;
        }
        const float hi = __uint_as_float(f2tf32(v));
        const float lo = __uint_as_float(f2tf32(v - hi));
        Xp[(long long)j * DP + p] = make_float2(hi, lo);
    }
    xn[j] = j < n ? (float)nn : 3.0e38f;
}

// stats[p] = max_j |x_jp| (p < d, else 0), stats[DP] = max_j |x_j|^2.  One block of 1024 threads.
template <int DP>
__global__ void __launch_bounds__(1024)
k_knn_stats(const double* __restrict__ X, int n, int d, float* __restrict__ stats) {
    __shared__ float red[32][DP + 1];
    float mx[DP + 1];
#pragma unroll
    for (int p = 0; p <= DP; ++p) mx[p] = 0.f;
    for (int j = threadIdx.x; j < n; j += 1024) {
        double nn = 0.0;
#pragma unroll
        for (int p = 0; p < DP; ++p)
            if (p < d) {
                const double x = X[(long long)j * d + p];
                nn += x * x;
                mx[p] = fmaxf(mx[p], (float)fabs(x) * 1.0000002f);
            }
        mx[DP] = fmaxf(mx[DP], (float)nn * 1.0000002f);
    }
#pragma unroll
    for (int p = 0; p <= DP; ++p) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx[p] = fmaxf(mx[p], __shfl_xor_sync(0xffffffffu, mx[p], o));
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][p] = mx[p];
    }
    __syncthreads();
    if (threadIdx.x <= DP) {
        float m = 0.f;
        for (int w = 0; w < 32; ++w) m = fmaxf(m, red[w][threadIdx.x]);
        stats[threadIdx.x] = m;
    }
}

// sorted (ascending) insert into a 4-entry list; the caller has checked s < v[3]
__device__ __forceinline__ void ktc_insert(float (&v)[KTC_KP], int (&ix)[KTC_KP], float s, int j) {
    if (s < v[2]) {
        v[3] = v[2];
        ix[3] = ix[2];
        if (s < v[1]) {
            v[2] = v[1];
            ix[2] = ix[1];
            if (s < v[0]) {
                v[1] = v[0];
                ix[1] = ix[0];
                v[0] = s;
                ix[0] = j;
            } else {
                v[1] = s;
                ix[1] = j;
            }
        } else {
            v[2] = s;
            ix[2] = j;
        }
    } else {
        v[3] = s;
        ix[3] = j;
    }
}

// k-th smallest (k <= 3) of the values the four lanes of a query row have kept: every lane contributes its three
// smallest (sorted); two xor-shuffle rounds inside the 4-lane group merge them (branch-free insertion of the partner's
// triple), after which every lane of the group holds the row's three smallest.  All 32 lanes must call it.
__device__ __forceinline__ float ktc_row_kth(const float (&v)[KTC_KP], int k) {
    float a0 = v[0], a1 = v[1], a2 = v[2];
#pragma unroll
    for (int m = 1; m <= 2; m <<= 1) {
        const float b0 = __shfl_xor_sync(0xffffffffu, a0, m);
        const float b1 = __shfl_xor_sync(0xffffffffu, a1, m);
        const float b2 = __shfl_xor_sync(0xffffffffu, a2, m);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const float x = q == 0 ? b0 : (q == 1 ? b1 : b2);
            a2 = fminf(a2, x);
            float lo = fminf(a1, a2);
            a2 = fmaxf(a1, a2);
            a1 = lo;
            lo = fminf(a0, a1);
            a1 = fmaxf(a0, a1);
            a0 = lo;
        }
    }
    return k == 1 ? a0 : (k == 2 ? a1 : a2);
}

// ---- main kernel --------------------------------------------------------------------------------------------------
// grid ceil(M / 128), block 256.  Dynamic smem: 2 stages x (KTC_TILE x DP float2 + KTC_TILE float) + merge area
// 8 warps x 16 rows x 16 entries x (float + int).
// Row stride of a training point in shared memory: DP + 4 float2.  The B fragment load of lane (g, t) reads the float2
// at row g, slot t (and t + 4): with stride DP + 4 the 16 lanes of a half-warp hit (4 g + t) mod 16 = 16 different
// 8-byte bank pairs (stride DP: 8 g + t -> 2-way conflicts, ncu: 4.4e9 conflicts for the 4M x 32k routing).
#define KTC_LD(DP) ((DP) + 4)
#define KTC_STAGE_BYTES(DP) (KTC_TILE * KTC_LD(DP) * 8 + KTC_TILE * 4)
#define KTC_SMEM_BYTES(DP) (2 * KTC_STAGE_BYTES(DP) + 8 * 16 * 16 * 8)

template <int DP>
__device__ __forceinline__ void ktc_load_stage(char* stage, const float2* __restrict__ Xp, const float* __restrict__ xn, int t0,
                                               int tid) {
    // KTC_TILE rows of DP float2 (DP * 8 bytes = CPR 16-byte chunks each) at row stride KTC_LD(DP), then the norms
    constexpr int CPR = DP * 8 / 16, XCH = KTC_TILE * CPR;
    const char* gx = reinterpret_cast<const char*>(Xp + (long long)t0 * DP);
    for (int c = tid; c < XCH; c += 256) cp_async16(stage + (c / CPR) * (KTC_LD(DP) * 8) + (c % CPR) * 16, gx + c * 16);
    const char* gn = reinterpret_cast<const char*>(xn + t0);
    char* sn = stage + KTC_TILE * KTC_LD(DP) * 8;
    for (int c = tid; c < KTC_TILE * 4 / 16; c += 256) cp_async16(sn + c * 16, gn + c * 16);
}

template <int DP>
__global__ void __launch_bounds__(256)
k_knn_tc(const float2* __restrict__ Xp, const float* __restrict__ xn, const float* __restrict__ stats,
         const double* __restrict__ X, const long long* __restrict__ lab, int n, int npad, int d, int k,
         const double* __restrict__ Q, long long M, long long* __restrict__ out, long long* __restrict__ qlist,
         int* __restrict__ qcount, int qcap) {
    extern __shared__ __align__(16) char ktc_smem[];
    constexpr int KS = DP / 8;  // k-steps per tile
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const long long qbase = (long long)blockIdx.x * KTC_QPB + warp * 16;

    // A fragments of -2 q (rows g and g+8), TF32 hi / lo
    unsigned ahi[KS][4], alo[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const long long q = qbase + g + ((e & 1) ? 8 : 0);
            const int p = ks * 8 + t + ((e & 2) ? 4 : 0);
            float v = 0.f;
            if (q < M && p < d) v = (float)(-2.0 * Q[q * d + p]);
            const unsigned h = f2tf32(v);
            ahi[ks][e] = h;
            alo[ks][e] = f2tf32(v - __uint_as_float(h));
        }
    float v0[KTC_KP], v1[KTC_KP];
    int i0[KTC_KP], i1[KTC_KP];
#pragma unroll
    for (int s = 0; s < KTC_KP; ++s) {
        v0[s] = v1[s] = INFINITY;
        i0[s] = i1[s] = -1;
    }
    // Row-wide cut.  A lane only has to keep points that can still be survivors, i.e. s~ <= tau_final + 2 eps; tau_final is
    // at most the k-th smallest s~ the whole ROW (its four lanes) has kept so far, so everything above
    // rc = (k-th smallest of the row so far) + 2 eps' (eps' >= eps_q) is dropped at once.  The row sees four times the
    // points of a lane and uses its k-th instead of the lane's (k+1)-th value, which cuts the number of list insertions
    // (the slow path of this loop) ~4.5x.  rc is refreshed after every insertion (two xor-shuffle rounds in the 4-lane group).
    float e2r0, e2r1;  // 2 eps' of rows g and g + 8
    {
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int p = ks * 8 + t + ((e & 2) ? 4 : 0);
                const float a = fabsf(__uint_as_float(ahi[ks][e])) + fabsf(__uint_as_float(alo[ks][e]));  // >= |2 q_p|
                if (e & 1)
                    s1 += a * stats[p];
                else
                    s0 += a * stats[p];
            }
        s0 += __shfl_xor_sync(0xffffffffu, s0, 1);
        s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
        s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
        e2r0 = 2.f * 1.001f * KTC_EPS_SCALE * (stats[DP] + s0 * 1.0000005f);
        e2r1 = 2.f * 1.001f * KTC_EPS_SCALE * (stats[DP] + s1 * 1.0000005f);
    }
    float rc0 = INFINITY, rc1 = INFINITY;  // row cuts
    float h0 = INFINITY, h1 = INFINITY;    // hit thresholds = min(lane's largest kept value, row cut)

    const int ntiles = npad / KTC_TILE;
    ktc_load_stage<DP>(ktc_smem, Xp, xn, 0, tid);
    cp_async_commit();
    for (int tl = 0; tl < ntiles; ++tl) {
        if (tl + 1 < ntiles) ktc_load_stage<DP>(ktc_smem + ((tl + 1) & 1) * KTC_STAGE_BYTES(DP), Xp, xn, (tl + 1) * KTC_TILE, tid);
        cp_async_commit();
        cp_async_wait<1>();
        __syncthreads();
        const char* stage = ktc_smem + (tl & 1) * KTC_STAGE_BYTES(DP);
        const float2* xs = reinterpret_cast<const float2*>(stage);
        const float* ns = reinterpret_cast<const float*>(stage + KTC_TILE * KTC_LD(DP) * 8);
#pragma unroll 4
        for (int j8 = 0; j8 < KTC_TILE / 8; ++j8) {
            const float2 nn = *reinterpret_cast<const float2*>(ns + j8 * 8 + 2 * t);
            float c[4] = {nn.x, nn.y, nn.x, nn.y};
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const float2 b0 = xs[(j8 * 8 + g) * KTC_LD(DP) + ks * 8 + t];
                const float2 b1 = xs[(j8 * 8 + g) * KTC_LD(DP) + ks * 8 + t + 4];
                const unsigned bh0 = __float_as_uint(b0.x), bh1 = __float_as_uint(b1.x);
                const unsigned bl0 = __float_as_uint(b0.y), bl1 = __float_as_uint(b1.y);
                mma_tf32_16x8x8(c, alo[ks], bh0, bh1);  // small terms first
                mma_tf32_16x8x8(c, ahi[ks], bl0, bl1);
                mma_tf32_16x8x8(c, ahi[ks], bh0, bh1);
            }
            // almost never taken after the first tiles: one warp-uniform branch keeps the (long, predicated) insertion
            // code out of the steady-state instruction stream
            const bool hit = fminf(c[0], c[1]) < h0 || fminf(c[2], c[3]) < h1;
            if (__any_sync(0xffffffffu, hit)) {
                const int j = tl * KTC_TILE + j8 * 8 + 2 * t;
                if (c[0] < v0[KTC_KP - 1] && c[0] < rc0) ktc_insert(v0, i0, c[0], j);
                if (c[1] < v0[KTC_KP - 1] && c[1] < rc0) ktc_insert(v0, i0, c[1], j + 1);
                if (c[2] < v1[KTC_KP - 1] && c[2] < rc1) ktc_insert(v1, i1, c[2], j);
                if (c[3] < v1[KTC_KP - 1] && c[3] < rc1) ktc_insert(v1, i1, c[3], j + 1);
                rc0 = ktc_row_kth(v0, k) + e2r0;
                rc1 = ktc_row_kth(v1, k) + e2r1;
                h0 = fminf(v0[KTC_KP - 1], rc0);
                h1 = fminf(v1[KTC_KP - 1], rc1);
            }
        }
        __syncthreads();
    }
    cp_async_wait<0>();

    // ---- merge the four lane lists of every query row, certify, exact finish (one thread per query)
    float* mv = reinterpret_cast<float*>(ktc_smem + 2 * KTC_STAGE_BYTES(DP)) + warp * 16 * 16;
    int* mi = reinterpret_cast<int*>(ktc_smem + 2 * KTC_STAGE_BYTES(DP) + 8 * 16 * 16 * 4) + warp * 16 * 16;
#pragma unroll
    for (int s = 0; s < KTC_KP; ++s) {
        mv[g * 16 + t * KTC_KP + s] = v0[s];
        mi[g * 16 + t * KTC_KP + s] = i0[s];
        mv[(g + 8) * 16 + t * KTC_KP + s] = v1[s];
        mi[(g + 8) * 16 + t * KTC_KP + s] = i1[s];
    }
    __syncwarp();
    if (lane >= 16) return;
    const long long q = qbase + lane;
    if (q >= M) return;
    const float* rv = mv + lane * 16;
    const int* ri = mi + lane * 16;
    // tau = k-th smallest of the 16 kept values
    float tau;
    {
        float best[3] = {INFINITY, INFINITY, INFINITY};
        for (int e = 0; e < 16; ++e) {
            const float s = rv[e];
            if (s < best[2]) {
                if (s < best[1]) {
                    best[2] = best[1];
                    if (s < best[0]) {
                        best[1] = best[0];
                        best[0] = s;
                    } else {
                        best[1] = s;
                    }
                } else {
                    best[2] = s;
                }
            }
        }
        tau = k == 1 ? best[0] : (k == 2 ? best[1] : best[2]);
    }
    double qv[DP];
    float sq = stats[DP];
    double qn = 0.0;
#pragma unroll
    for (int p = 0; p < DP; ++p) {
        qv[p] = p < d ? Q[q * d + p] : 0.0;
        qn += qv[p] * qv[p];
        sq += 2.f * (float)fabs(qv[p]) * stats[p] * 1.0000002f;
    }
    const float eps = KTC_EPS_SCALE * sq;
    const float cut = tau + 2.f * eps;
    bool certain = isfinite(tau);
    // a full lane list whose largest entry is a survivor may have dropped survivors
#pragma unroll
    for (int l = 0; l < 4; ++l)
        if (ri[l * KTC_KP + KTC_KP - 1] >= 0 && rv[l * KTC_KP + KTC_KP - 1] <= cut) certain = false;
    double bd[3] = {INFINITY, INFINITY, INFINITY};
    int bi[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int e = 0; e < 16 && certain; ++e) {
        const int j = ri[e];
        if (j < 0 || !(rv[e] <= cut)) continue;
        double dist = 0.0;
        for (int p = 0; p < d; ++p) {  // exactly the arithmetic of knn.cuh
            const double tt = __dsub_rn(qv[p], X[(long long)j * d + p]);
            dist = __dadd_rn(dist, __dmul_rn(tt, tt));
        }
        // self-check of the error bound: s~ against the exact s = dist - |q|^2
        if (fabs((double)rv[e] - (dist - qn)) > (double)eps) certain = false;
        double cd = dist;
        int ci = j;
#pragma unroll
        for (int s = 0; s < 3; ++s)
            if (cd < bd[s] || (cd == bd[s] && ci < bi[s])) {
                const double td = bd[s];
                const int ti = bi[s];
                bd[s] = cd;
                bi[s] = ci;
                cd = td;
                ci = ti;
            }
    }
    if (!certain || bi[k - 1] == 0x7fffffff) {
        out[q] = -1;  // handed to the exact brute-force kernel: listed for its dense pass (the mark alone if the list is full)
        const int pos = atomicAdd(qcount, 1);
        if (pos < qcap) qlist[pos] = q;
        return;
    }
    // uniform vote, smallest label on ties (same rule as knn.cuh)
    long long l3[3];
    for (int s = 0; s < 3; ++s) l3[s] = s < k ? lab[bi[s]] : -1;
    long long best_l = -1;
    int best_c = 0;
    for (int s = 0; s < k; ++s) {
        int cnt = 0;
        for (int u = 0; u < k; ++u) cnt += l3[u] == l3[s] ? 1 : 0;
        if (cnt > best_c || (cnt == best_c && l3[s] < best_l)) {
            best_c = cnt;
            best_l = l3[s];
        }
    }
    out[q] = best_l;
}
