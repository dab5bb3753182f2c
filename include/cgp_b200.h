/*
 * cgp_b200 — C-ABI of the B200-native per-cluster GP hot path of gptune/cGP.
 *
 * The reference (pure Python) has no FFI: the boundary it crosses for this path is the
 * scikit-learn object API.  Each entry point below replaces the reference call site cited
 * beside it (REF = /root/reference, SK = scikit-learn 1.9.0, SP = SciPy 1.18.1); the Python
 * binding a maintainer would add is cgp_b200/_lib.py (ctypes) — see INTEGRATION.md.
 *
 * Conventions
 *   - all bulk pointers are DEVICE pointers to contiguous fp64 / int64 arrays owned by the caller;
 *     pointers named *_host are small HOST arrays (launch geometry);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it;
 *   - the library never allocates device memory: callers pass a workspace sized by the
 *     matching *_workspace_bytes() query;
 *   - return value: 0 OK; <0 argument -i invalid (LAPACK style); >0 cudaError_t.
 *     Numerical status is per item in `info[]`: 0 OK, j>0 leading minor j not positive definite
 *     (LAPACK dpotrf convention; SK/gaussian_process/_gpr.py:591-593 maps it to lml=-inf, grad=0).
 *   - training rows are stored CLUSTER-SORTED: cluster c owns rows offs[c]..offs[c+1]-1 of X / y
 *     (REF/cGP/cGP_constrained.py:307-312 slices them per cluster).
 *   - theta = [log l_1..log l_d, log sigma_n^2]  (SK/gaussian_process/kernels.py:752,763-765),
 *     P = d + 1.
 *   - factor storage: every cluster matrix is padded to npad = cgp_padded_n(n_c) (multiple of the
 *     128-wide tile) and stored row-major [npad, npad]; cluster c starts at element
 *     sum_{c'<c} npad_{c'}^2.  The pad block is the identity.
 */
#ifndef CGP_B200_H
#define CGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cgp_handle cgp_handle;

enum { CGP_KIND_MATERN32 = 0, CGP_KIND_RBF = 1 };

/* Library version (major*10000 + minor*100 + patch). */
int cgp_version(void);

/* Handle: device id, SM count and a small pinned staging buffer for launch metadata. */
int cgp_create(cgp_handle** out, int device);
int cgp_destroy(cgp_handle* h);

/* Runtime switches of a handle (defaults come from the environment variables CGP_KNN_MODE = exact|tc,
 * CGP_FORK_MAX, CGP_TILE_MODE = big|small at cgp_create).  None of them changes a result bit:
 *   CGP_OPT_KNN_MODE   0 choose by problem size, 1 exact fp64 brute-force kernel only, 2 tensor-core prefilter + exact
 *                      finish whenever k <= 3 (cgp_knn_labels)
 *   CGP_OPT_FORK_MAX   batches of more matrices than this run the factorisation on ONE stream instead of the
 *                      three-stream look-ahead schedule (0: never fork; measurement passes use it for exclusive timing)
 *   CGP_OPT_TILE_MODE  0 CTA tile shape by launch size, 1 always 128x128, 2 always the small shapes */
enum { CGP_OPT_KNN_MODE = 1, CGP_OPT_FORK_MAX = 2, CGP_OPT_TILE_MODE = 3 };
int cgp_set_option(cgp_handle* h, int option, int value);
int cgp_get_option(cgp_handle* h, int option, int* value);

/* npad for a cluster of n samples. */
int64_t cgp_padded_n(int64_t n);
/* sum_c npad_c^2 — element count of the L / W factor buffers. */
int64_t cgp_model_elems(const int64_t* offs_host, int n_clusters);

/* K1: dense kernel matrix K = k(X,X) + (exp(theta[d]) + alpha) I, row-major [n,n].
 * Replaces `kernel_(X_train_)` + `K[diag] += alpha`, SK/gaussian_process/_gpr.py:584-589,349-350
 * (Matern nu=1.5: SK kernels.py:1713-1743; RBF :1558-1565; White :1406-1407; Sum :866-871). */
int cgp_kernel_matrix(cgp_handle* h, int kind, const double* X, int64_t n, int d,
                      const double* theta, double alpha, double* K, void* stream);

/* Cross kernel k(Q, X) [m,n] (White contributes 0, SK kernels.py:1418-1419);
 * replaces `kernel_(X, X_train_)`, SK/_gpr.py:447. */
int cgp_kernel_cross(cgp_handle* h, int kind, const double* Q, int64_t m, const double* X, int64_t n,
                     int d, const double* theta, double* Kqx, void* stream);

/* K2..K6: log-marginal likelihood and gradient for B (cluster, theta) pairs in one call.
 * Replaces `GaussianProcessRegressor.log_marginal_likelihood(theta, eval_gradient=True)`,
 * SK/_gpr.py:541-656, called once per L-BFGS-B evaluation per restart per cluster
 * (SK/_gpr.py:302-333; REF/cGP/cGP_constrained.py:327).
 *   y          normalised targets (SK/_gpr.py:275-280), cluster-sorted
 *   pair_cluster_host[B]  cluster of each pair;  theta [B,P] device
 *   lml[B], grad[B,P] (grad may be NULL), info[B] device outputs. */
size_t cgp_lml_workspace_bytes(const int64_t* offs_host, int n_clusters,
                               const int32_t* pair_cluster_host, int n_pairs, int d);
int cgp_lml_grad_batched(cgp_handle* h, int kind, const double* X, const double* y,
                         const int64_t* offs_host, int n_clusters, int d,
                         const int32_t* pair_cluster_host, int n_pairs, const double* theta,
                         double alpha, double* lml, double* grad, int32_t* info,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Final factorisation at the selected theta of every cluster.
 * Replaces SK/_gpr.py:349-367 (`L_ = cholesky(K)`, `alpha_ = cho_solve(L_, y)`).
 *   L [model_elems]   lower Cholesky factors (padded layout)
 *   W [model_elems]   W = L^-1 (lower) — what the predict kernel multiplies by instead of
 *                     LAPACK dtrtrs (SK/_gpr.py:460)
 *   alpha_vec [n]     K^-1 y, cluster-sorted;  lml[C];  info[C]. */
size_t cgp_factorize_workspace_bytes(const int64_t* offs_host, int n_clusters, int d);
int cgp_factorize(cgp_handle* h, int kind, const double* X, const double* y,
                  const int64_t* offs_host, int n_clusters, int d, const double* theta,
                  double alpha, double* L, double* W, double* alpha_vec, double* lml,
                  int32_t* info, void* workspace, size_t workspace_bytes, void* stream);

/* K7: k-NN routing.  Replaces `classify.predict(X)`, REF/cGP/acquisition.py:7 ->
 * SK/neighbors/_classification.py:245-312 (k nearest by exact fp64 squared Euclidean distance,
 * un-fused mul/add as SK/metrics/_dist_metrics.pxd.tp:44-49; distance ties -> lowest training
 * index; vote ties -> smallest label).  X here is in the ORIGINAL sample order with labels[n]. */
int cgp_knn_labels(cgp_handle* h, const double* X, const int64_t* labels, int64_t n, int d, int k,
                   int n_classes, const double* Q, int64_t m, int64_t* out_labels, void* stream);

/* K8: routed posterior mean / std for m query points.
 * Replaces `GPR_list[label].predict(x, return_std=True)`, REF/cGP/f_truth_dill.py:20-30 and
 * SK/_gpr.py:446-500.  X / alpha_vec cluster-sorted, theta [C,P], y_mean[C], y_std[C] device.
 * mean[m], std[m] in the caller's query order (std may be NULL). */
size_t cgp_score_workspace_bytes(const int64_t* offs_host, int n_clusters, int d, int64_t m);
int cgp_predict(cgp_handle* h, int kind, const double* X, const int64_t* offs_host, int n_clusters,
                int d, const double* theta, const double* W, const double* alpha_vec,
                const double* y_mean, const double* y_std, const double* Q,
                const int64_t* qlabel, int64_t m, double* mean, double* std,
                void* workspace, size_t workspace_bytes, void* stream);

/* K8 + EI + argmax.  Replaces the per-point `expected_improvement` objective,
 * REF/cGP/acquisition.py:4-32, and the cross-cluster strict-max, REF/cGP/cGP_constrained.py:376-378,
 * by a dense scan of m candidates:
 *   score_i = EI(mean_i, std_i; y_max[c]) / n_c   with c = qlabel[i]
 *   best = argmax score; ties -> lowest cluster id, then lowest candidate index.
 * penalty [m] (device, candidate order, NULL = none): the additive soft-constraint term of REF/cGP/acquisition.py:29,
 *   score_i = (EI_i + penalty_i) / n_c  (the caller evaluates its boundary_penalty(x_i, X_c) hook per candidate).
 * Outputs (device): best_score[1], best_idx[1] (index into Q, plus idx_base), best_cluster[1];
 * score/mean/std [m] optional (NULL to skip). */
int cgp_ei_argmax(cgp_handle* h, int kind, const double* X, const int64_t* offs_host,
                  int n_clusters, int d, const double* theta, const double* W,
                  const double* alpha_vec, const double* y_mean, const double* y_std,
                  const double* y_max, const double* Q, const int64_t* qlabel, int64_t m,
                  int64_t idx_base, const double* penalty, double* score, double* mean, double* std,
                  double* best_score, int64_t* best_idx, int64_t* best_cluster, void* workspace,
                  size_t workspace_bytes, void* stream);

/* Same selection as cgp_ei_argmax (identical best_score bits, best_idx, best_cluster) with exact bound pruning:
 * EI grows with sigma and the posterior variance given all samples never exceeds the one given a single sample,
 *   var_i <= (1 + s2) - kmax_i^2 / (1 + s2 + alpha),  kmax_i = largest cross-kernel value of candidate i,
 * so EI(mean_i, sigma_ub_i)/n_c bounds score_i from the predictive-mean pass alone; candidates whose bound is below the
 * best exact score found so far skip the n_c^2 variance contraction.  alpha = the jitter the model was factorised with
 * (cgp_factorize).  No per-candidate outputs (pruned candidates have no exact score). */
int cgp_ei_argmax_pruned(cgp_handle* h, int kind, const double* X, const int64_t* offs_host,
                         int n_clusters, int d, const double* theta, double alpha, const double* W,
                         const double* alpha_vec, const double* y_mean, const double* y_std,
                         const double* y_max, const double* Q, const int64_t* qlabel, int64_t m,
                         int64_t idx_base, const double* penalty, double* best_score, int64_t* best_idx,
                         int64_t* best_cluster, void* workspace, size_t workspace_bytes, void* stream);

/* Incremental refit.  The reference re-clusters and re-factorises everything after
 * each new sample (REF/cGP/cGP_constrained.py:276,315-327).  When the labels of the old samples did not change, the
 * factor of the one cluster that received the sample is extended by one row at unchanged theta:
 *   L' = [[L,0],[l^T,dd]],  l = W k,  dd = sqrt(kappa - l.l),  W' = [[W,0],[-(l^T W)/dd, 1/dd]]     (2 n^2 flops).
 *   Xc [n_new, d]     rows of the cluster, the new sample LAST
 *   theta [P]         the cluster's hyper-parameters (device)
 *   L, W [npad,npad]  the cluster's padded factor blocks, npad = cgp_padded_n(n_new); the leading (n_new-1)^2 block
 *                     holds the old factor, the rest the identity pad.  Row n_new-1 is written.
 *   info[1]           0, or n_new when the extended matrix is not positive definite. */
size_t cgp_append_workspace_bytes(int64_t n_new);
int cgp_append_sample(cgp_handle* h, int kind, const double* Xc, int64_t n_new, int d,
                      const double* theta, double alpha, double* L, double* W, int64_t npad,
                      int32_t* info, void* workspace, size_t workspace_bytes, void* stream);

/* alpha = K^-1 y = W^T (W y) and the log-marginal likelihood value of one factorised cluster
 * (SK/_gpr.py:363-367, 601-617) - needed after an append because y is re-normalised when a sample arrives.
 * Workspace: cgp_append_workspace_bytes(n). */
int cgp_refresh_alpha(cgp_handle* h, const double* L, const double* W, int64_t npad, int64_t n,
                      const double* y, double* alpha_out, double* lml, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Dense MSPE scan, the reference's second acquisition (REF/cGP/acquisition.py:34-75, minimised :75) over m candidates:
 *   mspe(x) = s_xx - s_xo S_oo s_xo^T  from the joint posterior covariance of [x; X_c]  =  c0[c] - k*^T B_c k*
 * with the per-cluster symmetric matrix B_c = ys^2 K^-1 + ys^6 G S~ G^T (G = I - K^-1 K_o, S~ = K_o + s2 I - K_o K^-1 K_o;
 * built once per cluster by the caller with cgp_dgemm_nt) and c0[c] = ys^2 (1 + s2); value_i = (mspe_i + penalty_i) / n_c.
 * One [rows, n_c] x [n_c, n_c] DMMA product + a row dot per candidate chunk.  best_value holds MINUS the minimum
 * (the arg-max machinery of cgp_ei_argmax runs on -value); ties -> lowest cluster id, then lowest candidate index.
 *   Bmat  per-cluster [npad, npad] blocks, packed like W;  c0 [n_clusters] HOST array;  value [m] optional. */
size_t cgp_quadform_workspace_bytes(const int64_t* offs_host, int n_clusters, int d, int64_t m);
int cgp_quadform_scan(cgp_handle* h, int kind, const double* X, const int64_t* offs_host, int n_clusters, int d,
                      const double* theta, const double* Bmat, const double* c0_host, const double* Q,
                      const int64_t* qlabel, int64_t m, int64_t idx_base, const double* penalty, double* value,
                      double* best_value, int64_t* best_idx, int64_t* best_cluster, void* workspace,
                      size_t workspace_bytes, void* stream);

/* Building blocks of the dense MSPE scan (the reference's second acquisition, REF/cGP/acquisition.py:34-75: joint
 * posterior covariance of [x; X_sample] from predict(return_cov=True), mspe = s_xx - s_xo S_oo s_xo^T).  As a scan over
 * m candidates the two products are [m, n_c] x [n_c, n_c] GEMMs through cgp_dgemm_nt (cgp_b200/engine.py mspe_scan);
 * cgp_rowdot finishes the quadratic form, cgp_argmax_sorted applies the winner rule of cgp_ei_argmax to scores in
 * bucket order (cluster-major, candidate-index-minor).  scratch: 16 KiB. */
int cgp_rowdot(cgp_handle* h, const double* A, const double* B, int64_t rows, int64_t n, int64_t ld,
               double* out, void* stream);
int cgp_argmax_sorted(cgp_handle* h, const double* score_sorted, int64_t m, const int64_t* perm,
                      const int64_t* coff, int n_clusters, int64_t idx_base, double* best_score,
                      int64_t* best_idx, int64_t* best_cluster, void* scratch, size_t scratch_bytes,
                      void* stream);

/* Measurement helper (bench / tests only): C[M,N] = A[M,K] * B[N,K]^T through the same fp64
 * tensor-core (DMMA) tile mainloop every factorisation step uses.  M,N multiples of 128, K of 32. */
int cgp_dgemm_nt(cgp_handle* h, const double* A, const double* B, double* C, int64_t M, int64_t N,
                 int64_t K, void* stream);

/* Measurement: every kernel launch is counted per kernel class; with profiling enabled launches are also bracketed by
 * CUDA events on the stream they are launched on: on = 1 every launch, on = 2 only the large kernels (k_lauum,
 * k_predict_vnorm, the trailing updates, k-NN, cross kernel, gradient, assembly) so that the event pairs do not sit on
 * the latency-bound chain of small kernels.  cgp_profile_read() folds the finished events into per-class totals
 * (milliseconds) and returns them with the launch counts (arrays of cgp_profile_classes() entries). */
int cgp_profile_classes(void);
const char* cgp_profile_class_name(int cls);
int cgp_profile_enable(cgp_handle* h, int on);
int cgp_profile_read(cgp_handle* h, double* ms, int64_t* launches, int n_cls, int reset);
/* Algorithmic flops (real rows / columns only) and count of the TIMED launches per class since the last
 * reset; call before the resetting cgp_profile_read().  Batches of at most CGP_FORK_MAX (default 24) matrices run the
 * three-stream look-ahead schedule: their launches are counted but never timed (a bracket there would include work of
 * the other two streams); larger batches run on one stream, so ms / flops describe exclusive kernel time. */
int cgp_profile_flops(cgp_handle* h, double* flops, int64_t* timed, int n_cls);
/* on != 0: count but do not time the following launches of this handle (the caller is running a batch of ANOTHER handle
 * concurrently on the same device, so event brackets would not be exclusive). */
int cgp_profile_mute(cgp_handle* h, int on);

#ifdef __cplusplus
}
#endif
#endif /* CGP_B200_H */
