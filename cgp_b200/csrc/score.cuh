// K8 plumbing around the DMMA variance contraction: stable bucketing of candidates by routed cluster,
// predictive mean, variance / std / EI finishing, and the deterministic arg-max.
#pragma once
#include "common.cuh"

#define BUCKET_CH 2048  // candidates per bucketing block

// ---- stable counting sort of candidates by cluster label --------------------------------------------
// pass 1: per-block histogram.  blkhist [C][nblk].   grid (nblk), block 256, smem C ints
__global__ void __launch_bounds__(256)
k_bucket_hist(const long long* __restrict__ qlabel, long long M, int C, int nblk, int* __restrict__ blkhist) {
    extern __shared__ int hist[];
    for (int c = threadIdx.x; c < C; c += 256) hist[c] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * BUCKET_CH;
    for (int i = threadIdx.x; i < BUCKET_CH; i += 256) {
        long long q = base + i;
        if (q < M) {
            int l = (int)qlabel[q];
            l = max(0, min(C - 1, l));
            atomicAdd(&hist[l], 1);
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) blkhist[(long long)c * nblk + blockIdx.x] = hist[c];
}

// pass 2: exclusive scan over (cluster-major, block-minor).  One block of 1024 threads.
// blkoff [C][nblk] start position of each (cluster, block) run; coff [C+1] cluster offsets.
__global__ void __launch_bounds__(1024)
k_bucket_scan(const int* __restrict__ blkhist, long long* __restrict__ blkoff, long long* __restrict__ coff, int C,
              int nblk) {
    __shared__ long long part[1024];
    const long long total = (long long)C * nblk;
    const long long per = (total + 1023) / 1024;
    const long long s = threadIdx.x * per, e = min(total, s + per);
    long long acc = 0;
    for (long long i = s; i < e; ++i) acc += blkhist[i];
    part[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int t = 0; t < 1024; ++t) {
            long long v = part[t];
            part[t] = run;
            run += v;
        }
        coff[C] = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (long long i = s; i < e; ++i) {
        blkoff[i] = run;
        if (i % nblk == 0) coff[i / nblk] = run;
        run += blkhist[i];
    }
}

// pass 3: one warp per block walks its candidates in index order; __match_any groups equal labels so
// ranks inside a cluster keep ascending candidate index (stable).  perm[pos] = candidate, Qs[pos] = Q[cand].
// grid (nblk), block 32, smem C long longs
__global__ void __launch_bounds__(32)
k_bucket_scatter(const long long* __restrict__ qlabel, const double* __restrict__ Q, long long M, int C, int nblk, int d,
                 const long long* __restrict__ blkoff, long long* __restrict__ perm, double* __restrict__ Qs) {
    extern __shared__ long long cnt[];
    const int lane = threadIdx.x;
    for (int c = lane; c < C; c += 32) cnt[c] = blkoff[(long long)c * nblk + blockIdx.x];
    __syncwarp();
    const long long base = (long long)blockIdx.x * BUCKET_CH;
    for (int r = 0; r < BUCKET_CH; r += 32) {
        const long long q = base + r + lane;
        const bool ok = q < M;
        int l = ok ? (int)qlabel[q] : -1;
        if (ok) l = max(0, min(C - 1, l));
        const unsigned mask = __match_any_sync(0xffffffffu, l);
        const int rank = __popc(mask & ((1u << lane) - 1u));
        const int leader = __ffs(mask) - 1;
        long long pos = 0;
        if (ok) pos = cnt[l] + rank;
        __syncwarp();
        if (ok && lane == leader) cnt[l] += __popc(mask);
        __syncwarp();
        if (ok) {
            perm[pos] = q;
            for (int p = 0; p < d; ++p) Qs[pos * d + p] = Q[q * d + p];
        }
    }
}

// ---- predictive mean: mraw[r] = sum_{k<n} Kstar[r,k] alpha[k].  warp per row; grid (ceil(rows/8)), block 256
// kmax (optional): kmax[r] = max_k |Kstar[r,k]|, the cross-kernel value of the nearest training point (input of the
// single-neighbour variance bound of the pruned scan).
__global__ void __launch_bounds__(256)
k_mean(const double* __restrict__ Kstar, const double* __restrict__ alpha, int n, int npad, long long rows,
       double* __restrict__ mraw, double* __restrict__ kmax) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* kr = Kstar + r * npad;
    double s = 0.0, mx = 0.0;
    for (int k = lane; k < n; k += 32) {
        const double v = kr[k];
        s += v * alpha[k];
        mx = fmax(mx, fabs(v));
    }
    s = warp_sum(s);
    if (kmax) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (lane == 0) {
        mraw[r] = s;
        if (kmax) kmax[r] = mx;
    }
}

__device__ __forceinline__ double ei_value(double mean, double sd, double ymax) {
    // REF/cGP/acquisition.py:23-27 with xi = 0
    if (!(sd > 0.0)) return 0.0;
    const double imp = mean - ymax;
    const double Z = imp / sd;
    const double cdf = 0.5 * erfc(-Z * 0.70710678118654752440);
    const double pdf = exp(-0.5 * Z * Z) * 0.39894228040143267794;
    return imp * cdf + sd * pdf;
}

// ---- finishing: variance from the per-row-block partials, std, mean un-normalisation, EI / n_c.
// Outputs are scattered to the caller's candidate order through perm; score_sorted keeps bucket order for
// the arg-max.  grid (ceil(rows/256)), block 256
__global__ void __launch_bounds__(256)
k_score_finish(const double* __restrict__ vpart, long long vstride, int T, const double* __restrict__ mraw,
               long long rows, long long pos0, const long long* __restrict__ perm, const double* __restrict__ th, int d,
               const double* __restrict__ ymean, const double* __restrict__ ystd, const double* __restrict__ ymax,
               int cluster, double n_c, double* __restrict__ mean_out, double* __restrict__ std_out,
               double* __restrict__ score_out, double* __restrict__ score_sorted, const double* __restrict__ penalty) {
    const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    double vs = 0.0;
    for (int ib = 0; ib < T; ++ib) vs += vpart[(long long)ib * vstride + r];
    const double noise = exp(th[d]);
    double var = (1.0 + noise) - vs;  // kernel_.diag = Matern 1 + White (SK kernels.py:890,1438); _gpr.py:482-483
    if (var < 0.0) var = 0.0;         // _gpr.py:485-491
    const double ys = ystd[cluster];
    const double sd = sqrt(var * (ys * ys));        // _gpr.py:494-500
    const double mu = ys * mraw[r] + ymean[cluster];  // _gpr.py:451
    const long long orig = perm[pos0 + r];
    if (mean_out) mean_out[orig] = mu;
    if (std_out) std_out[orig] = sd;
    if (ymax) {
        // acquisition.py:29-32: (EI + boundary_penalty) / n_c ; penalty is indexed by the caller's candidate order
        const double sc = (penalty ? ei_value(mu, sd, ymax[cluster]) + penalty[orig] : ei_value(mu, sd, ymax[cluster])) / n_c;
        if (score_out) score_out[orig] = sc;
        score_sorted[pos0 + r] = sc;
    }
}

// ---- exact bound-pruned scan --------------------------------------------------------------------------------
// EI(mu, sigma) grows with sigma, and the posterior variance given ALL samples of the cluster never exceeds the
// posterior variance given only ONE of them (conditioning on more data cannot increase a Gaussian's variance):
//     var(x) <= (1 + noise) - k(x, x_j)^2 / (1 + noise + alpha)      for every training point j,
// tightest for the nearest one (largest cross-kernel value kmax, a by-product of the mean pass).  So
//     ub = EI(mu, sigma_ub) / n_c,  sigma_ub^2 = ((1 + noise) - kmax^2 / (1 + noise + alpha)) y_std^2  (inflated by 1e-6
// relative + 1e-12 of the prior against rounding of the exact variance) bounds the score of a candidate from its
// predictive MEAN pass alone (n_c flops instead of n_c^2).  Far from every sample kmax -> 0 and the bound is the prior
// variance; next to a sample it drops to ~noise.  A candidate whose bound is below the
// best exact score found so far (`lb`, device scalar) can neither win nor tie; it skips the variance contraction.
// Survivors are scored exactly as in the exhaustive scan (same bits), so the selected sample is identical.
// The bound is inflated by 1e-9 relative: the computed EI is monotone in sigma only up to rounding (the two terms of
// REF/cGP/acquisition.py:26 cancel by a factor ~Z^2 in the lower tail, i.e. relative errors up to ~1e-13).
// One block of 1024 threads per chunk: thread t owns rows [t*per, (t+1)*per) -> survivors are listed in ascending order.
__global__ void __launch_bounds__(1024)
k_prune_select(const double* __restrict__ mraw, long long rows, const double* __restrict__ th, int d,
               const double* __restrict__ ymean, const double* __restrict__ ystd, const double* __restrict__ ymax,
               int cluster, double n_c, const double* __restrict__ lb, int* __restrict__ sel, int* __restrict__ count,
               double* __restrict__ score_chunk, const double* __restrict__ penalty, const long long* __restrict__ perm_chunk,
               const double* __restrict__ kmax, double alpha) {
    __shared__ int part[1024];
    const int per = (int)((rows + 1023) / 1024);
    const long long s = (long long)threadIdx.x * per, e = min(rows, s + per);
    const double ys = ystd[cluster], ym = ymean[cluster], yx = ymax[cluster];
    const double prior = 1.0 + exp(th[d]), kjj = prior + alpha;
    const double bound = *lb;
    unsigned long long keep = 0ull;  // per <= 64 (chunks hold at most 65536 candidates)
    int c = 0;
    for (long long r = s; r < e; ++r) {
        const double km = kmax[r];
        const double var_ub = fmin(prior, (prior - km * km / kjj) * (1.0 + 1e-6) + 1e-12 * prior);
        double ub = ei_value(ys * mraw[r] + ym, sqrt(fmax(var_ub, 0.0) * (ys * ys)), yx);
        // the additive penalty shifts bound and score alike; the 1e-9 inflation is applied to the EI part only
        const double pen = penalty ? penalty[perm_chunk[r]] : 0.0;
        ub = (ub * (1.0 + 1e-9) + 1e-300 + pen) / n_c;
        const bool k = !(ub + 1e-15 * fabs(ub) < bound);
        if (k) {
            keep |= 1ull << (int)(r - s);
            ++c;
        } else {
            score_chunk[r] = -INFINITY;
        }
    }
    part[threadIdx.x] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int t = 0; t < 1024; ++t) {
            int v = part[t];
            part[t] = run;
            run += v;
        }
        *count = run;
    }
    __syncthreads();
    int o = part[threadIdx.x];
    for (long long r = s; r < e; ++r)
        if (keep >> (int)(r - s) & 1ull) sel[o++] = (int)r;
}

// exact score of survivor q (row sel[q] of the chunk); same arithmetic as k_score_finish.  grid ceil(rows/256)
__global__ void __launch_bounds__(256)
k_score_finish_sel(const double* __restrict__ vpart, long long vstride, int T, const double* __restrict__ mraw,
                   const int* __restrict__ sel, const int* __restrict__ count, const double* __restrict__ th, int d,
                   const double* __restrict__ ymean, const double* __restrict__ ystd, const double* __restrict__ ymax,
                   int cluster, double n_c, double* __restrict__ score_chunk, const double* __restrict__ penalty,
                   const long long* __restrict__ perm_chunk) {
    const int q = blockIdx.x * 256 + threadIdx.x;
    if (q >= *count) return;
    const int r = sel[q];
    double vs = 0.0;
    for (int ib = 0; ib < T; ++ib) vs += vpart[(long long)ib * vstride + q];
    const double noise = exp(th[d]);
    double var = (1.0 + noise) - vs;
    if (var < 0.0) var = 0.0;
    const double ys = ystd[cluster];
    const double sd = sqrt(var * (ys * ys));
    const double mu = ys * mraw[r] + ymean[cluster];
    score_chunk[r] = (penalty ? ei_value(mu, sd, ymax[cluster]) + penalty[perm_chunk[r]] : ei_value(mu, sd, ymax[cluster])) / n_c;
}

// lb = max(lb, max over the chunk's exact scores).  One block of 1024.
__global__ void __launch_bounds__(1024)
k_lb_update(const double* __restrict__ score_chunk, long long rows, double* __restrict__ lb) {
    __shared__ double sv[32];
    double v = -INFINITY;
    for (long long i = threadIdx.x; i < rows; i += 1024) {
        const double s = score_chunk[i];
        if (s > v) v = s;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) sv[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 32; ++w) v = fmax(v, sv[w]);
        if (v > *lb) *lb = v;
    }
}

__global__ void k_set_double(double* p, double v) { *p = v; }

// ---- arg-max over bucket order (cluster-major, candidate-index-minor): the FIRST maximum wins, which is
// "lowest cluster id, then lowest candidate index" (REF/cGP/cGP_constrained.py:305,376-378).
__device__ __forceinline__ void argmax_combine(double& v, long long& p, double ov, long long op) {
    if (ov > v || (ov == v && op < p)) {
        v = ov;
        p = op;
    }
}

// grid (nb), block 256: blkv/blkp [nb]
__global__ void __launch_bounds__(256)
k_argmax_blocks(const double* __restrict__ score_sorted, long long M, double* __restrict__ blkv, long long* __restrict__ blkp) {
    __shared__ double sv[8];
    __shared__ long long sp[8];
    double v = -INFINITY;
    long long p = 0x7fffffffffffffffLL;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < M; i += (long long)gridDim.x * 256) {
        double s = score_sorted[i];
        if (s == s) argmax_combine(v, p, s, i);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        long long op = __shfl_xor_sync(0xffffffffu, p, o);
        argmax_combine(v, p, ov, op);
    }
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = v;
        sp[threadIdx.x >> 5] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) argmax_combine(v, p, sv[w], sp[w]);
        blkv[blockIdx.x] = v;
        blkp[blockIdx.x] = p;
    }
}

// one block: final winner -> best_score, best_idx (original index + idx_base), best_cluster
__global__ void __launch_bounds__(256)
k_argmax_final(const double* __restrict__ blkv, const long long* __restrict__ blkp, int nb, const long long* __restrict__ perm,
               const long long* __restrict__ coff, int C, long long idx_base, double* __restrict__ best_score,
               long long* __restrict__ best_idx, long long* __restrict__ best_cluster) {
    __shared__ double sv[8];
    __shared__ long long sp[8];
    double v = -INFINITY;
    long long p = 0x7fffffffffffffffLL;
    for (int i = threadIdx.x; i < nb; i += 256) argmax_combine(v, p, blkv[i], blkp[i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double ov = __shfl_xor_sync(0xffffffffu, v, o);
        long long op = __shfl_xor_sync(0xffffffffu, p, o);
        argmax_combine(v, p, ov, op);
    }
    if ((threadIdx.x & 31) == 0) {
        sv[threadIdx.x >> 5] = v;
        sp[threadIdx.x >> 5] = p;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) argmax_combine(v, p, sv[w], sp[w]);
        best_score[0] = v;
        if (p == 0x7fffffffffffffffLL) {
            best_idx[0] = -1;
            best_cluster[0] = -1;
        } else {
            best_idx[0] = perm[p] + idx_base;
            int c = 0;
            while (c + 1 < C && coff[c + 1] <= p) ++c;
            best_cluster[0] = c;
        }
    }
}

// out[r] = sum_{k<n} A[r,k] * B[r,k]  (row-wise dot of two row-major [rows, ld] matrices).  warp per row, fixed order.
// grid ceil(rows/8), block 256.  Used by the dense MSPE scan (quadratic form s_xo S_oo s_xo^T, REF/cGP/acquisition.py:68).
__global__ void __launch_bounds__(256)
k_rowdot(const double* __restrict__ A, const double* __restrict__ B, long long rows, int n, long long ld, double* __restrict__ out) {
    const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const double* a = A + r * ld;
    const double* b = B + r * ld;
    double s = 0.0;
    for (int k = lane; k < n; k += 32) s += a[k] * b[k];
    s = warp_sum(s);
    if (lane == 0) out[r] = s;
}

// ---- dense quadratic-form scan (MSPE): value = (c0 - q + penalty) / n_c, the caller MINIMISES it, so the bucket-ordered
// array for the arg-max holds -value (first maximum = lowest cluster id, then lowest candidate index).
// grid ceil(rows/256), block 256
__global__ void __launch_bounds__(256)
k_quadform_finish(const double* __restrict__ q, long long rows, long long pos0, const long long* __restrict__ perm, double c0,
                  double n_c, const double* __restrict__ penalty, double* __restrict__ value_out,
                  double* __restrict__ neg_sorted) {
    const long long r = (long long)blockIdx.x * 256 + threadIdx.x;
    if (r >= rows) return;
    const long long orig = perm[pos0 + r];
    const double v = ((c0 - q[r]) + (penalty ? penalty[orig] : 0.0)) / n_c;
    if (value_out) value_out[orig] = v;
    neg_sorted[pos0 + r] = -v;
}
