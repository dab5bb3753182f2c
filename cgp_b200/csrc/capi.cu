// C-ABI of cgp_b200 (see include/cgp_b200.h).  Host-side orchestration only: every entry point is a
// fixed sequence of asynchronous kernel launches on the caller's stream.
#include <cstdio>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/cgp_b200.h"
#include "append.cuh"
#include "chol_diag.cuh"
#include "common.cuh"
#include "gemm_nt.cuh"
#include "kernel_fn.cuh"
#include "knn.cuh"
#include "knn_tc.cuh"
#include "score.cuh"

#define CGP_VERSION 100
#define PINNED_BYTES (8u << 20)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
static inline int64_t pad_n(int64_t n) { return (n + CGP_TILE - 1) / CGP_TILE * CGP_TILE; }

enum {
    CLS_ASSEMBLE, CLS_POTRF_UPDATE, CLS_POTRF_TRAIL, CLS_POTRF_DIAG, CLS_POTRF_PANEL, CLS_TRTRI_ACC, CLS_TRTRI_TRAIL,
    CLS_TRTRI_SCALE, CLS_SOLVE,
    CLS_LAUUM, CLS_GRAD, CLS_TRANSPOSE, CLS_KNN, CLS_BUCKET, CLS_CROSS, CLS_MEAN, CLS_PREDICT_VNORM,
    CLS_SCORE_FINISH, CLS_ARGMAX, CLS_DGEMM, CLS_APPEND, CLS_KNN_PREP, CLS_KNN_EXACT, CLS_COUNT
};
static const char* const kClassNames[CLS_COUNT] = {
    "assemble", "potrf_update", "potrf_trail", "potrf_diag", "potrf_panel", "trtri_acc", "trtri_trail",
    "trtri_scale", "solve_lml",
    "lauum", "grad", "transpose", "knn", "bucket", "cross", "mean", "predict_vnorm",
    "score_finish", "argmax", "dgemm_nt", "append", "knn_prep", "knn_exact"};

struct ProfRec {
    cudaEvent_t a, b;
    int cls;
    double flops;  // algorithmic flops of the bracketed launch (0 when not tallied)
};

struct HandleImpl : cgp_handle {
    cudaEvent_t ring_event;
    bool ring_event_pending;
    // per-kernel-class launch counters (always) and CUDA-event timing (when enabled)
    int prof_on;      // 0 off, 1 every launch bracketed by events, 2 only the large kernels the roofline is quoted for
    bool prof_skip;   // the launch in flight is counted but not timed
    bool user_mute;   // cgp_profile_mute: another lane (handle) runs concurrently on this device -> count only
    bool prof_mute;   // inside a forked (three-stream) batch: brackets would overlap other streams -> count only
    std::vector<ProfRec> recs;
    size_t recs_used;
    double ms[CLS_COUNT];
    double flops[CLS_COUNT];       // algorithmic flops of the TIMED launches of each class (roofline numerator)
    double pending_flops;          // set by FL() for the next KL launch
    long long launches[CLS_COUNT];
    long long timed[CLS_COUNT];    // launches that were bracketed by events
    void* knn_buf;     // scratch of the tensor-core k-NN prefilter (grow-only)
    size_t knn_bytes;
    long long* knn_qlist;  // list / count of the queries the last prefilter launch could not certify (inside knn_buf)
    int* knn_qcount;
    int knn_qcap;
    int knn_mode;      // 0 = by problem size, 1 = always the exact fp64 kernel, 2 = prefilter whenever k <= 3 (CGP_KNN_MODE)
    int fork_max;  // batches with more matrices than this run the one-stream schedule (CGP_FORK_MAX, default 24)
    // internal streams of run_pipeline: sc (high priority) = the dependent chain update -> diag -> panel + look-ahead part
    // of the trailing update; s2 (low priority) = triangular-inverse work of panel p beside the Cholesky of panel p+1.
    // The caller's stream carries the bulk trailing updates and everything else; all three are joined before return.
    cudaStream_t s2, sc;
    int tile_mode;  // 0 = by launch size, 1 = always 128x128 CTAs, 2 = always the small shapes
    std::vector<cudaEvent_t> panel_ev;  // [2p] = panel p factored (on sc), [2p+1] = bulk update of panel p done (on s)
    cudaEvent_t join_ev, start_ev, chain_ev;
};

static inline void prof_begin(HandleImpl* h, int cls, cudaStream_t s) {
    h->launches[cls]++;
    if (!h->prof_on || h->prof_mute || h->user_mute) {
        h->prof_skip = true;
        return;
    }
    // level 2: the event pair costs ~2 x 1.5 us of stream latency per launch, which is visible on the latency-bound
    // chain of small kernels (diag / panel / update ...) when few matrices are in flight; time only the big ones.
    h->prof_skip = h->prof_on == 2 && !(cls == CLS_LAUUM || cls == CLS_PREDICT_VNORM || cls == CLS_POTRF_TRAIL ||
                                        cls == CLS_TRTRI_TRAIL || cls == CLS_KNN || cls == CLS_KNN_EXACT || cls == CLS_CROSS || cls == CLS_GRAD ||
                                        cls == CLS_ASSEMBLE);
    if (h->prof_skip) return;
    if (h->recs_used == h->recs.size()) {
        ProfRec r;
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        h->recs.push_back(r);
    }
    ProfRec& r = h->recs[h->recs_used];
    r.cls = cls;
    r.flops = h->pending_flops;
    cudaEventRecord(r.a, s);
}
static inline void prof_end(HandleImpl* h, cudaStream_t s) {
    h->pending_flops = 0.0;
    if (!h->prof_on || h->prof_skip) return;
    cudaEventRecord(h->recs[h->recs_used].b, s);
    h->recs_used++;
}
// KL(handle, class, stream, kernel<<<...>>>(...)) : count (and optionally time) one launch
#define KL(h, cls, s, ...)    \
    do {                      \
        prof_begin(h, cls, s); \
        __VA_ARGS__;          \
        prof_end(h, s);       \
    } while (0)

// FL(h, flops): algorithmic flops of the NEXT KL launch (tallied only if that launch is timed)
#define FL(h, f) ((h)->pending_flops = (double)(f))

// Every entry point that launches work runs on the handle's device, whatever the caller's current device is, and
// restores the caller's device on return.
struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};
#define CGP_ENTER(hh)                                  \
    HandleImpl* h = static_cast<HandleImpl*>(hh);      \
    if (!h) return -1;                                 \
    DeviceGuard guard_(h->device)

template <typename K>
static cudaError_t set_smem(K kernel, int bytes) {
    return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}

extern "C" int cgp_version(void) { return CGP_VERSION; }
extern "C" int64_t cgp_padded_n(int64_t n) { return pad_n(n); }
extern "C" int64_t cgp_model_elems(const int64_t* offs, int C) {
    int64_t s = 0;
    for (int c = 0; c < C; ++c) {
        int64_t np = pad_n(offs[c + 1] - offs[c]);
        s += np * np;
    }
    return s;
}

// Releases whatever a (possibly partially constructed) handle owns.
static void handle_free(HandleImpl* h) {
    for (auto& r : h->recs) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    for (auto& e : h->panel_ev) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->join_ev, h->start_ev, h->chain_ev, h->ring_event})
        if (e) cudaEventDestroy(e);
    if (h->s2) cudaStreamDestroy(h->s2);
    if (h->sc) cudaStreamDestroy(h->sc);
    if (h->pinned) cudaFreeHost(h->pinned);
    if (h->knn_buf) cudaFree(h->knn_buf);
    delete h;
}

static int handle_init(HandleImpl* h, int device) {
    h->device = device;
    CGP_CHECK(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device));
    CGP_CHECK(cudaMallocHost(&h->pinned, PINNED_BYTES));
    h->pinned_bytes = PINNED_BYTES;
    h->pinned_used = 0;
    CGP_CHECK(cudaEventCreateWithFlags(&h->ring_event, cudaEventDisableTiming));
    {
        int lo = 0, hi = 0;
        CGP_CHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CGP_CHECK(cudaStreamCreateWithPriority(&h->s2, cudaStreamNonBlocking, lo));  // lowest priority: fills gaps
        const char* pe = getenv("CGP_CHAIN_PRIO");  // "0": chain stream at default priority (measurement switch)
        CGP_CHECK(cudaStreamCreateWithPriority(&h->sc, cudaStreamNonBlocking, (pe && pe[0] == '0') ? lo : hi));  // highest: the dependent chain
        CGP_CHECK(cudaEventCreateWithFlags(&h->join_ev, cudaEventDisableTiming));
        CGP_CHECK(cudaEventCreateWithFlags(&h->start_ev, cudaEventDisableTiming));
        CGP_CHECK(cudaEventCreateWithFlags(&h->chain_ev, cudaEventDisableTiming));
    }
    CGP_CHECK(set_smem(k_dgemm_nt, GEMM_SMEM_BYTES));
#define SET3(K)                                                                       \
    CGP_CHECK(set_smem(K<128, 128>, GEMM_SMEM_BYTES_T(128, 128)));                    \
    CGP_CHECK(set_smem(K<64, 64>, GEMM_SMEM_BYTES_T(64, 64)))
#define SET2(K)                                                                       \
    CGP_CHECK(set_smem(K<128, 128>, GEMM_SMEM_BYTES_T(128, 128)));                    \
    CGP_CHECK(set_smem(K<64, 128>, GEMM_SMEM_BYTES_T(64, 128)))
    SET3(k_potrf_update);
    SET3(k_potrf_trail);
    SET3(k_potrf_head);
    SET3(k_trtri_acc);
    SET3(k_trtri_trail);
    SET3(k_lauum);
    SET2(k_potrf_panel);
    SET2(k_trtri_scale);
#undef SET3
#undef SET2
    CGP_CHECK(set_smem(k_predict_vnorm, GEMM_SMEM_BYTES));
    CGP_CHECK(set_smem(k_predict_vnorm_sel, GEMM_SMEM_BYTES));
    {
        const char* e = getenv("CGP_TILE_MODE");  // "big" | "small" | unset = by launch size (results are bit-identical)
        h->tile_mode = (e && !strcmp(e, "big")) ? 1 : (e && !strcmp(e, "small")) ? 2 : 0;
    }
    CGP_CHECK(set_smem(k_potrf_diag, DIAG_SMEM_BYTES));
    return 0;
}

extern "C" int cgp_create(cgp_handle** out, int device) {
    if (!out) return -1;
    *out = nullptr;
    DeviceGuard guard(device);  // the caller's current device is restored on return
    int cur = -1;
    CGP_CHECK(cudaGetDevice(&cur));
    if (cur != device) return (int)cudaErrorInvalidDevice;
    HandleImpl* h = new (std::nothrow) HandleImpl();
    if (!h) return (int)cudaErrorMemoryAllocation;
    h->pinned = nullptr;
    h->ring_event = h->join_ev = h->start_ev = h->chain_ev = nullptr;
    h->s2 = h->sc = nullptr;
    h->ring_event_pending = false;
    h->prof_on = 0;
    h->prof_skip = false;
    h->recs_used = 0;
    for (int i = 0; i < CLS_COUNT; ++i) h->ms[i] = 0.0, h->launches[i] = 0, h->flops[i] = 0.0, h->timed[i] = 0;
    h->pending_flops = 0.0;
    h->prof_mute = false;
    h->user_mute = false;
    h->knn_buf = nullptr;
    h->knn_bytes = 0;
    {
        const char* e = getenv("CGP_KNN_MODE");
        h->knn_mode = (e && !strcmp(e, "exact")) ? 1 : (e && !strcmp(e, "tc")) ? 2 : 0;
    }
    {
        const char* e = getenv("CGP_FORK_MAX");
        h->fork_max = e ? atoi(e) : 24;
    }
    int rc = handle_init(h, device);
    if (rc) {  // nothing leaks when a step of the construction fails
        handle_free(h);
        return rc;
    }
    *out = h;
    return 0;
}

extern "C" int cgp_destroy(cgp_handle* hh) {
    if (!hh) return -1;
    HandleImpl* h = static_cast<HandleImpl*>(hh);
    DeviceGuard guard(h->device);
    handle_free(h);
    return 0;
}

// Runtime switches (the environment variables of the same name set the defaults at cgp_create):
//   CGP_OPT_KNN_MODE 0 by size | 1 exact fp64 kernel only | 2 tensor-core prefilter whenever k <= 3
//   CGP_OPT_FORK_MAX batches with more matrices than this use the one-stream schedule (0 = never fork)
//   CGP_OPT_TILE_MODE 0 by launch size | 1 always 128x128 CTAs | 2 always the small shapes
extern "C" int cgp_set_option(cgp_handle* hh, int option, int value) {
    if (!hh) return -1;
    HandleImpl* h = static_cast<HandleImpl*>(hh);
    switch (option) {
        case CGP_OPT_KNN_MODE:
            if (value < 0 || value > 2) return -3;
            h->knn_mode = value;
            return 0;
        case CGP_OPT_FORK_MAX:
            if (value < 0) return -3;
            h->fork_max = value;
            return 0;
        case CGP_OPT_TILE_MODE:
            if (value < 0 || value > 2) return -3;
            h->tile_mode = value;
            return 0;
    }
    return -2;
}
extern "C" int cgp_get_option(cgp_handle* hh, int option, int* value) {
    if (!hh) return -1;
    if (!value) return -3;
    HandleImpl* h = static_cast<HandleImpl*>(hh);
    switch (option) {
        case CGP_OPT_KNN_MODE: *value = h->knn_mode; return 0;
        case CGP_OPT_FORK_MAX: *value = h->fork_max; return 0;
        case CGP_OPT_TILE_MODE: *value = h->tile_mode; return 0;
    }
    return -2;
}

extern "C" int cgp_profile_classes(void) { return CLS_COUNT; }
extern "C" const char* cgp_profile_class_name(int cls) { return (cls >= 0 && cls < CLS_COUNT) ? kClassNames[cls] : ""; }
extern "C" int cgp_profile_enable(cgp_handle* hh, int on) {
    if (!hh) return -1;
    static_cast<HandleImpl*>(hh)->prof_on = on < 0 ? 0 : (on > 2 ? 2 : on);
    return 0;
}
// Count but do not time the following launches (the caller runs another handle's batch concurrently on this device).
extern "C" int cgp_profile_mute(cgp_handle* hh, int on) {
    if (!hh) return -1;
    static_cast<HandleImpl*>(hh)->user_mute = on != 0;
    return 0;
}
// Folds the finished event pairs into ms[] (synchronises on them), copies the totals out and optionally resets.
extern "C" int cgp_profile_read(cgp_handle* hh, double* ms, int64_t* launches, int n_cls, int reset) {
    CGP_ENTER(hh);
    for (size_t i = 0; i < h->recs_used; ++i) {
        float t = 0.f;
        CGP_CHECK(cudaEventSynchronize(h->recs[i].b));
        CGP_CHECK(cudaEventElapsedTime(&t, h->recs[i].a, h->recs[i].b));
        h->ms[h->recs[i].cls] += t;
        h->flops[h->recs[i].cls] += h->recs[i].flops;
        h->timed[h->recs[i].cls]++;
    }
    h->recs_used = 0;
    for (int i = 0; i < n_cls && i < CLS_COUNT; ++i) {
        if (ms) ms[i] = h->ms[i];
        if (launches) launches[i] = h->launches[i];
    }
    if (reset)
        for (int i = 0; i < CLS_COUNT; ++i) h->ms[i] = 0.0, h->launches[i] = 0, h->flops[i] = 0.0, h->timed[i] = 0;
    return 0;
}

// Algorithmic flops and number of the TIMED launches per class since the last reset (call before the resetting
// cgp_profile_read).  With profiling on, batches that run the three-stream schedule are counted but not timed (their
// event brackets would include work of the other two streams), so ms / flops / timed describe exclusive kernel time.
extern "C" int cgp_profile_flops(cgp_handle* hh, double* flops, int64_t* timed, int n_cls) {
    CGP_ENTER(hh);
    int rc = cgp_profile_read(hh, nullptr, nullptr, 0, 0);  // fold finished event pairs
    if (rc) return rc;
    for (int i = 0; i < n_cls && i < CLS_COUNT; ++i) {
        if (flops) flops[i] = h->flops[i];
        if (timed) timed[i] = h->timed[i];
    }
    return 0;
}

// Stage `bytes` of launch metadata through the pinned ring and copy it to `dst` on `s`.
static int upload(HandleImpl* h, void* dst, const void* src, size_t bytes, cudaStream_t s) {
    size_t need = align_up(bytes, 256);
    if (need > h->pinned_bytes) return -100;
    if (h->pinned_used + need > h->pinned_bytes) {
        if (h->ring_event_pending) CGP_CHECK(cudaEventSynchronize(h->ring_event));
        h->pinned_used = 0;
    }
    char* p = static_cast<char*>(h->pinned) + h->pinned_used;
    memcpy(p, src, bytes);
    h->pinned_used += need;
    CGP_CHECK(cudaMemcpyAsync(dst, p, bytes, cudaMemcpyHostToDevice, s));
    CGP_CHECK(cudaEventRecord(h->ring_event, s));
    h->ring_event_pending = true;
    return 0;
}

static inline int dmax_for(int d) { return d <= 4 ? 4 : (d <= 8 ? 8 : 16); }
static inline size_t xtile_smem(int d) { return (size_t)2 * CGP_TILE * (d | 1) * 8; }

// ------------------------------------------------------------------------------------------------
// dense kernel matrices (K1 single-matrix form, cross kernel)
// ------------------------------------------------------------------------------------------------
static int launch_cross(HandleImpl* h, int kind, const double* Q, int64_t m, const double* X, int64_t n, int d, const double* theta,
                        double* out, int64_t ldo, int64_t rows_out, int64_t cols_out, int self, double alpha,
                        cudaStream_t s) {
    dim3 grid((unsigned)((cols_out + CGP_TILE - 1) / CGP_TILE), (unsigned)((rows_out + CGP_TILE - 1) / CGP_TILE));
    if (grid.x == 0 || grid.y == 0) return 0;
    size_t sm = xtile_smem(d);
    if (kind == CGP_KIND_MATERN32)
        KL(h, CLS_CROSS, s, k_cross<0><<<grid, 256, sm, s>>>(Q, m, X, n, d, theta, out, ldo, rows_out, cols_out, self, alpha));
    else
        KL(h, CLS_CROSS, s, k_cross<1><<<grid, 256, sm, s>>>(Q, m, X, n, d, theta, out, ldo, rows_out, cols_out, self, alpha));
    return (int)cudaGetLastError();
}

extern "C" int cgp_kernel_matrix(cgp_handle* hh, int kind, const double* X, int64_t n, int d, const double* theta,
                                 double alpha, double* K, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X || n < 0) return -3;
    if (d < 1 || d > CGP_MAX_D) return -5;
    if (!theta) return -6;
    if (!K) return -8;
    return launch_cross(h, kind, X, n, X, n, d, theta, K, n, n, n, 1, alpha, (cudaStream_t)stream);
}

extern "C" int cgp_kernel_cross(cgp_handle* hh, int kind, const double* Q, int64_t m, const double* X, int64_t n, int d,
                                const double* theta, double* Kqx, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!Q || m < 0) return -3;
    if (!X || n < 0) return -5;
    if (d < 1 || d > CGP_MAX_D) return -7;
    if (!theta) return -8;
    if (!Kqx) return -9;
    return launch_cross(h, kind, Q, m, X, n, d, theta, Kqx, n, m, n, 0, 0.0, (cudaStream_t)stream);
}

extern "C" int cgp_dgemm_nt(cgp_handle* hh, const double* A, const double* B, double* C, int64_t M, int64_t N, int64_t K,
                            void* stream) {
    CGP_ENTER(hh);
    if (!A || !B || !C) return -2;
    if (M <= 0 || M % CGP_TILE || N <= 0 || N % CGP_TILE || K <= 0 || K % CGP_BK) return -5;
    dim3 grid((unsigned)(N / CGP_TILE), (unsigned)(M / CGP_TILE));
    cudaStream_t s = (cudaStream_t)stream;
    KL(h, CLS_DGEMM, s, k_dgemm_nt<<<grid, CGP_GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(A, B, C, M, N, (int)K));
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// batched factorisation pipeline
// ------------------------------------------------------------------------------------------------
// A launch with fewer 128x128 tiles than ~4 waves of SMs is split into 64x64 (or, for in-place steps, 64x128) pieces:
// same bits, 2-4x more CTAs, shorter dependent steps (see gemm_nt.cuh).
static inline bool small_tiles(const HandleImpl* h, long long tiles) {
    if (h->tile_mode == 1) return false;
    if (h->tile_mode == 2) return true;
    return tiles < 4LL * h->sm_count;
}
// GL(class, stream, kernel, tiles_x, n_mats, args...): out-of-place tile kernels;  GLI: in-place (tile <- tile * D^T)
#define GL(cls, st, K, tx, nm, ...)                                                                                   \
    do {                                                                                                              \
        if (small_tiles(h, (long long)(tx) * (nm)))                                                                    \
            KL(h, cls, st, (K<64, 64><<<dim3((tx)*4, nm), CGP_GEMM_THREADS, GEMM_SMEM_BYTES_T(64, 64), st>>>(__VA_ARGS__))); \
        else                                                                                                          \
            KL(h, cls, st, (K<128, 128><<<dim3(tx, nm), CGP_GEMM_THREADS, GEMM_SMEM_BYTES_T(128, 128), st>>>(__VA_ARGS__))); \
    } while (0)
#define GLI(cls, st, K, tx, nm, ...)                                                                                  \
    do {                                                                                                              \
        if (small_tiles(h, (long long)(tx) * (nm)))                                                                    \
            KL(h, cls, st, (K<64, 128><<<dim3((tx)*2, nm), CGP_GEMM_THREADS, GEMM_SMEM_BYTES_T(64, 128), st>>>(__VA_ARGS__))); \
        else                                                                                                          \
            KL(h, cls, st, (K<128, 128><<<dim3(tx, nm), CGP_GEMM_THREADS, GEMM_SMEM_BYTES_T(128, 128), st>>>(__VA_ARGS__))); \
    } while (0)

struct Batch {
    int nm = 0, Tmax = 0, npad_max = 0, ntl_max = 0;
    MatMeta* meta = nullptr;  // device
    double *A = nullptr, *U = nullptr, *D = nullptr, *logdet = nullptr, *z = nullptr, *a = nullptr, *gpart = nullptr;
};

// ---- algorithmic flop counts of the large launches (roofline numerators; real rows / columns only) ----
// number of elements (r, c) with a <= c < b and c <= r < n  (lower-triangular part of the columns [a, b))
static inline double tri_cols(double n, double a, double b) {
    b = std::min(b, n);
    return b > a ? (b - a) * (2.0 * n - a - b + 1.0) * 0.5 : 0.0;
}
// rank-K update of the lower triangle restricted to the element columns [128 ta, 128 tb): 2 K per element
static double flops_syrk(const MatMeta* hm, int nm, int c0, int c1, int ta, int tb) {
    double f = 0.0;
    for (int i = 0; i < nm; ++i)
        f += 2.0 * CGP_TILE * (c1 - c0) * tri_cols(hm[i].n, (double)CGP_TILE * ta, (double)CGP_TILE * tb);
    return f;
}
// push of the inverse panel [c0, c1) into the columns >= c1: rows j < c1 (128 each), real columns n - 128 c1,
// K_j = 128 (c1 - max(j, c0)), the diagonal block U[j,j] (j >= c0) is triangular (half of its 128 k)
static double flops_trtri_trail(const MatMeta* hm, int nm, int c0, int c1) {
    double ksum = 0.0;
    for (int j = 0; j < c1; ++j) ksum += CGP_TILE * (c1 - std::max(j, c0)) - (j >= c0 ? CGP_TILE / 2 : 0);
    double f = 0.0;
    for (int i = 0; i < nm; ++i) f += 2.0 * CGP_TILE * std::max(0.0, (double)hm[i].n - (double)CGP_TILE * c1) * ksum;
    return f;
}
static double flops_cube_third(const MatMeta* hm, int nm) {
    double f = 0.0;
    for (int i = 0; i < nm; ++i) f += (double)hm[i].n * hm[i].n * hm[i].n / 3.0;
    return f;
}

static int run_pipeline_body(HandleImpl* h, int kind, const double* X, const double* y, const double* theta, double alpha,
                             int d, const Batch& b, const MatMeta* hm, double* lml, double* grad, int* info, cudaStream_t s);

// assemble -> blocked Cholesky -> U = L^-T -> z, a, lml [-> K^-1 -> gradient].   hm: host copy of b.meta
// A CUDA error in the middle of the three-stream section must not leave the internal streams running behind the
// caller's back (or the handle muted for profiling): drain them before the error code is returned.
static int run_pipeline(HandleImpl* h, int kind, const double* X, const double* y, const double* theta, double alpha, int d,
                        const Batch& b, const MatMeta* hm, double* lml, double* grad, int* info, cudaStream_t s) {
    const int rc = run_pipeline_body(h, kind, X, y, theta, alpha, d, b, hm, lml, grad, info, s);
    if (rc != 0) {
        cudaStreamSynchronize(h->sc);
        cudaStreamSynchronize(h->s2);
        h->prof_mute = false;
        h->pending_flops = 0.0;
    }
    return rc;
}

static int run_pipeline_body(HandleImpl* h, int kind, const double* X, const double* y, const double* theta, double alpha,
                             int d, const Batch& b, const MatMeta* hm, double* lml, double* grad, int* info, cudaStream_t s) {
    const int nm = b.nm, T = b.Tmax;
    const size_t xs = xtile_smem(d);
    dim3 gl(b.ntl_max, nm);
    if (kind == 0)
        KL(h, CLS_ASSEMBLE, s, k_assemble<0><<<gl, 256, xs, s>>>(b.A, X, theta, b.meta, d, alpha));
    else
        KL(h, CLS_ASSEMBLE, s, k_assemble<1><<<gl, 256, xs, s>>>(b.A, X, theta, b.meta, d, alpha));
    // Three streams (when there is more than one panel):
    //   sc  (high priority)  the latency-bound chain  update -> diag -> panel  of each block column, plus the part of the
    //                        trailing update that the NEXT panel needs (k_potrf_head: look-ahead);
    //   s   (caller)         the bulk of each trailing update (k_potrf_trail on the columns after the next panel);
    //   s2  (low priority)   the triangular-inverse work of panel p, which only needs the Cholesky panel p.
    // With few matrices in the batch (tail of the L-BFGS lock-step, small share per GPU) the chain leaves most SMs idle;
    // the bulk and inverse launches of the same matrices fill them.  Every tile sees the same sequence of updates as in
    // a single-stream schedule, so results are bit-identical.
    // Large batches fill the SMs from one stream and run the plain one-stream schedule (fork_max, CGP_FORK_MAX): same
    // bits, and every event bracket of the profiling levels is then exclusive kernel time.
    const int n_panels = (T + CGP_PANEL - 1) / CGP_PANEL;
    const bool fork = n_panels > 1 && nm <= h->fork_max;
    h->prof_mute = fork;
    cudaStream_t st = fork ? h->s2 : s;
    cudaStream_t sc = fork ? h->sc : s;
    while ((int)h->panel_ev.size() < 2 * n_panels) {
        cudaEvent_t e;
        CGP_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        h->panel_ev.push_back(e);
    }
    if (fork) {
        CGP_CHECK(cudaEventRecord(h->start_ev, s));
        CGP_CHECK(cudaStreamWaitEvent(sc, h->start_ev, 0));
    }
    bool rest_pending = false;  // a bulk update whose result the next k_potrf_head must see
    for (int c0 = 0, p = 0; c0 < T; c0 += CGP_PANEL, ++p) {
        const int c1 = std::min(T, c0 + CGP_PANEL);
        for (int j = c0; j < c1; ++j) {
            if (j > c0)
                GL(CLS_POTRF_UPDATE, sc, k_potrf_update, T - j, nm, b.A, b.meta, j, c0);
            KL(h, CLS_POTRF_DIAG, sc, k_potrf_diag<<<nm, 256, DIAG_SMEM_BYTES, sc>>>(b.A, b.U, b.D, b.logdet, info, b.meta, j, T));
            if (j + 1 < T)
                GLI(CLS_POTRF_PANEL, sc, k_potrf_panel, T - j - 1, nm, b.A, b.D, b.meta, j);
        }
        if (fork) {
            CGP_CHECK(cudaEventRecord(h->panel_ev[2 * p], sc));
            CGP_CHECK(cudaStreamWaitEvent(st, h->panel_ev[2 * p], 0));
        }
        if (c1 < T) {
            const int H = std::min(CGP_PANEL, T - c1);
            if (fork && rest_pending) CGP_CHECK(cudaStreamWaitEvent(sc, h->panel_ev[2 * (p - 1) + 1], 0));
            rest_pending = false;
            FL(h, flops_syrk(hm, nm, c0, c1, c1, c1 + H));
            GL(CLS_POTRF_TRAIL, sc, k_potrf_head, H * (T - c1), nm, b.A, b.meta, c0, c1, T);
            const int Tr = T - c1 - H;
            if (Tr > 0) {
                if (fork) CGP_CHECK(cudaStreamWaitEvent(s, h->panel_ev[2 * p], 0));
                FL(h, flops_syrk(hm, nm, c0, c1, c1 + H, T));
                GL(CLS_POTRF_TRAIL, s, k_potrf_trail, Tr * (Tr + 1) / 2, nm, b.A, b.meta, c0, c1, c1 + H);
                if (fork) {
                    CGP_CHECK(cudaEventRecord(h->panel_ev[2 * p + 1], s));
                    rest_pending = true;
                }
            }
        }
        for (int i = std::max(c0, 1); i < c1; ++i) {
            if (i > c0)
                GL(CLS_TRTRI_ACC, st, k_trtri_acc, i, nm, b.A, b.U, b.meta, i, c0);
            GLI(CLS_TRTRI_SCALE, st, k_trtri_scale, i, nm, b.U, b.D, b.meta, i);
        }
        if (c1 < T) {
            FL(h, flops_trtri_trail(hm, nm, c0, c1));
            GL(CLS_TRTRI_TRAIL, st, k_trtri_trail, c1 * (T - c1), nm, b.A, b.U, b.meta, c0, c1, T);
        }
    }
    if (fork) {
        CGP_CHECK(cudaEventRecord(h->chain_ev, sc));
        CGP_CHECK(cudaStreamWaitEvent(s, h->chain_ev, 0));
        CGP_CHECK(cudaEventRecord(h->join_ev, st));
        CGP_CHECK(cudaStreamWaitEvent(s, h->join_ev, 0));
    }
    h->prof_mute = false;
    KL(h, CLS_SOLVE, s, k_solve_z<<<dim3(T, nm), 1024, 0, s>>>(b.U, y, b.z, b.meta, b.npad_max));
    KL(h, CLS_SOLVE, s, k_solve_a<<<dim3(b.npad_max / 8, nm), 256, 0, s>>>(b.U, b.z, b.a, b.meta, b.npad_max));
    KL(h, CLS_SOLVE, s, k_lml_finish<<<nm, 256, 0, s>>>(b.z, b.logdet, info, lml, b.meta, b.npad_max, T));
    if (grad) {
        FL(h, flops_cube_third(hm, nm));
        GL(CLS_LAUUM, s, k_lauum, b.ntl_max, nm, b.A, b.U, b.meta);
        const int dm = dmax_for(d);
        const size_t gs = xs + (2 * CGP_TILE + 8 * dm) * 8;
#define LAUNCH_GRAD(KIND, DM)                                                                                        \
    KL(h, CLS_GRAD, s, (k_grad_tiles<KIND, DM><<<gl, 256, gs, s>>>(b.A, X, b.a, theta, b.gpart, b.meta, d, b.npad_max, b.ntl_max))); \
    KL(h, CLS_GRAD, s, (k_grad_finish<DM><<<nm, 256, 0, s>>>(b.A, b.a, b.gpart, theta, info, grad, b.meta, d, b.npad_max, b.ntl_max)))
        if (kind == 0) {
            if (dm == 4) { LAUNCH_GRAD(0, 4); } else if (dm == 8) { LAUNCH_GRAD(0, 8); } else { LAUNCH_GRAD(0, 16); }
        } else {
            if (dm == 4) { LAUNCH_GRAD(1, 4); } else if (dm == 8) { LAUNCH_GRAD(1, 8); } else { LAUNCH_GRAD(1, 16); }
        }
#undef LAUNCH_GRAD
    }
    return (int)cudaGetLastError();
}

// Carve the batch buffers for mats [m0, m1) out of ws.  Returns bytes used, or 0 if it does not fit.
static size_t carve(char* ws, size_t ws_bytes, const std::vector<MatMeta>& metas, int m0, int m1, int d, bool with_A,
                    Batch& b) {
    int nm = m1 - m0, npmax = 0;
    size_t a_el = 0, d_el = 0;
    for (int i = m0; i < m1; ++i) {
        npmax = std::max(npmax, metas[i].npad);
        a_el += (size_t)metas[i].npad * metas[i].npad;
        d_el += (size_t)metas[i].npad * CGP_TILE;
    }
    int T = npmax / CGP_TILE, ntl = T * (T + 1) / 2;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    size_t o_meta = take((size_t)nm * sizeof(MatMeta));
    size_t o_A = with_A ? take(a_el * 8) : 0;
    size_t o_U = take(a_el * 8);
    size_t o_D = take(d_el * 8);
    size_t o_ld = take((size_t)nm * T * 8);
    size_t o_z = take((size_t)nm * npmax * 8);
    size_t o_a = take((size_t)nm * npmax * 8);
    size_t o_g = take((size_t)nm * ntl * dmax_for(d) * 8);
    if (off > ws_bytes) return 0;
    b.nm = nm;
    b.Tmax = T;
    b.npad_max = npmax;
    b.ntl_max = ntl;
    b.meta = reinterpret_cast<MatMeta*>(ws + o_meta);
    b.A = with_A ? reinterpret_cast<double*>(ws + o_A) : nullptr;
    b.U = reinterpret_cast<double*>(ws + o_U);
    b.D = reinterpret_cast<double*>(ws + o_D);
    b.logdet = reinterpret_cast<double*>(ws + o_ld);
    b.z = reinterpret_cast<double*>(ws + o_z);
    b.a = reinterpret_cast<double*>(ws + o_a);
    b.gpart = reinterpret_cast<double*>(ws + o_g);
    return off;
}

static size_t batch_bytes(const std::vector<MatMeta>& metas, int m0, int m1, int d, bool with_A) {
    Batch b;
    return carve(nullptr, (size_t)-1, metas, m0, m1, d, with_A, b);
}

extern "C" size_t cgp_lml_workspace_bytes(const int64_t* offs, int C, const int32_t* pair_cluster, int B, int d) {
    if (!offs || !pair_cluster || B <= 0) return 0;
    std::vector<MatMeta> metas(B);
    for (int i = 0; i < B; ++i) {
        int c = pair_cluster[i];
        if (c < 0 || c >= C) return 0;
        metas[i].n = (int)(offs[c + 1] - offs[c]);
        metas[i].npad = (int)pad_n(metas[i].n);
    }
    return batch_bytes(metas, 0, B, d, true);
}

extern "C" int cgp_lml_grad_batched(cgp_handle* hh, int kind, const double* X, const double* y, const int64_t* offs,
                                    int C, int d, const int32_t* pair_cluster, int B, const double* theta, double alpha,
                                    double* lml, double* grad, int32_t* info, void* workspace, size_t ws_bytes,
                                    void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X) return -3;
    if (!y) return -4;
    if (!offs || C <= 0) return -5;
    if (d < 1 || d > CGP_MAX_D) return -7;
    if (!pair_cluster || B < 0) return -8;
    if (!theta) return -10;
    if (!lml) return -12;
    if (!info) return -14;
    if (!workspace) return -15;
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<MatMeta> metas(B);
    for (int i = 0; i < B; ++i) {
        int c = pair_cluster[i];
        if (c < 0 || c >= C) return -8;
        MatMeta& m = metas[i];
        m.n = (int)(offs[c + 1] - offs[c]);
        if (m.n <= 0) return -5;
        m.npad = (int)pad_n(m.n);
        m.xoff = (int)offs[c];
        m.theta_idx = i;
        m.out_idx = i;
        m.pad_ = 0;
    }
    CGP_CHECK(cudaMemsetAsync(info, 0, (size_t)B * sizeof(int32_t), s));
    // process the pairs in as few chunks as the workspace allows
    int m0 = 0;
    while (m0 < B) {
        int m1 = m0 + 1;
        if (batch_bytes(metas, m0, m1, d, true) > ws_bytes) return -16;
        // largest chunk that fits the workspace (bisection); gridDim.y carries the matrix index -> at most 65535
        int lo = m1, hi = std::min(B, m0 + 65535);
        while (lo < hi) {
            int mid = (lo + hi + 1) / 2;
            if (batch_bytes(metas, m0, mid, d, true) <= ws_bytes) lo = mid; else hi = mid - 1;
        }
        m1 = lo;
        Batch b;
        carve(static_cast<char*>(workspace), ws_bytes, metas, m0, m1, d, true, b);
        long long ao = 0, dof = 0;
        for (int i = m0; i < m1; ++i) {
            metas[i].a_off = ao;
            metas[i].u_off = ao;
            metas[i].d_off = dof;
            ao += (long long)metas[i].npad * metas[i].npad;
            dof += (long long)metas[i].npad * CGP_TILE;
        }
        int rc = upload(h, b.meta, &metas[m0], (size_t)(m1 - m0) * sizeof(MatMeta), s);
        if (rc) return rc;
        rc = run_pipeline(h, kind, X, y, theta, alpha, d, b, &metas[m0], lml, grad, info, s);
        if (rc) return rc;
        m0 = m1;
    }
    return 0;
}

// copy a (per-matrix strided scratch) into the cluster-sorted alpha vector
__global__ void k_gather_alpha(const double* __restrict__ abuf, double* __restrict__ alpha_vec, const MatMeta* __restrict__ meta,
                               int zstride) {
    const MatMeta m = meta[blockIdx.y];
    int i = blockIdx.x * 256 + threadIdx.x;
    if (i < m.n) alpha_vec[m.xoff + i] = abuf[(long long)blockIdx.y * zstride + i];
}

static void cluster_metas(const int64_t* offs, int C, std::vector<MatMeta>& metas) {
    metas.resize(C);
    long long ao = 0, dof = 0;
    for (int c = 0; c < C; ++c) {
        MatMeta& m = metas[c];
        m.n = (int)(offs[c + 1] - offs[c]);
        m.npad = (int)pad_n(m.n);
        m.xoff = (int)offs[c];
        m.theta_idx = c;
        m.out_idx = c;
        m.pad_ = 0;
        m.a_off = ao;
        m.u_off = ao;
        m.d_off = dof;
        ao += (long long)m.npad * m.npad;
        dof += (long long)m.npad * CGP_TILE;
    }
}

extern "C" size_t cgp_factorize_workspace_bytes(const int64_t* offs, int C, int d) {
    if (!offs || C <= 0) return 0;
    std::vector<MatMeta> metas;
    cluster_metas(offs, C, metas);
    return batch_bytes(metas, 0, C, d, false);
}

extern "C" int cgp_factorize(cgp_handle* hh, int kind, const double* X, const double* y, const int64_t* offs, int C, int d,
                             const double* theta, double alpha, double* L, double* W, double* alpha_vec, double* lml,
                             int32_t* info, void* workspace, size_t ws_bytes, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X) return -3;
    if (!y) return -4;
    if (!offs || C <= 0) return -5;
    if (d < 1 || d > CGP_MAX_D) return -7;
    if (!theta) return -8;
    if (!L) return -10;
    if (!W) return -11;
    if (!alpha_vec) return -12;
    if (!lml) return -13;
    if (!info) return -14;
    if (!workspace) return -15;
    cudaStream_t s = (cudaStream_t)stream;
    std::vector<MatMeta> metas;
    if (C > 65535) return -5;
    cluster_metas(offs, C, metas);
    for (int c = 0; c < C; ++c)
        if (metas[c].n <= 0) return -5;
    Batch b;
    if (carve(static_cast<char*>(workspace), ws_bytes, metas, 0, C, d, false, b) == 0) return -16;
    b.A = L;
    CGP_CHECK(cudaMemsetAsync(info, 0, (size_t)C * sizeof(int32_t), s));
    int64_t elems = cgp_model_elems(offs, C);
    CGP_CHECK(cudaMemsetAsync(b.U, 0, (size_t)elems * 8, s));
    CGP_CHECK(cudaMemsetAsync(L, 0, (size_t)elems * 8, s));  // strictly-upper tiles are never written by the pipeline
    int rc = upload(h, b.meta, metas.data(), (size_t)C * sizeof(MatMeta), s);
    if (rc) return rc;
    rc = run_pipeline(h, kind, X, y, theta, alpha, d, b, metas.data(), lml, nullptr, info, s);
    if (rc) return rc;
    KL(h, CLS_TRANSPOSE, s, k_transpose<<<dim3(b.npad_max / 32, b.npad_max / 32, C), dim3(32, 8), 0, s>>>(b.U, W, b.meta));
    KL(h, CLS_SOLVE, s, k_gather_alpha<<<dim3((b.npad_max + 255) / 256, C), 256, 0, s>>>(b.a, alpha_vec, b.meta, b.npad_max));
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// incremental refit: rank-1 append to one cluster (append.cuh)
// ------------------------------------------------------------------------------------------------
struct AppendLayout {
    int64_t npad;
    int nch;
    size_t o_k, o_l, o_dd, o_part, total;
};
static AppendLayout append_layout(int64_t n) {
    AppendLayout A;
    A.npad = pad_n(std::max<int64_t>(n, 1));
    A.nch = (int)((A.npad + APP_CH - 1) / APP_CH);
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    A.o_k = take((size_t)A.npad * 8);
    A.o_l = take((size_t)A.npad * 8);
    A.o_dd = take(8);
    A.o_part = take((size_t)A.nch * A.npad * 8);
    A.total = off;
    return A;
}

extern "C" size_t cgp_append_workspace_bytes(int64_t n_new) {
    if (n_new <= 0) return 0;
    return append_layout(n_new).total;
}

// out = W^T x over the leading n x n lower block (two launches, fixed summation order)
static void launch_tmatvec(HandleImpl* h, const double* W, int64_t ld, int n, const double* x, double* part, int64_t pstride,
                           const double* scale_ptr, double* out, int64_t ostride, cudaStream_t s) {
    dim3 g((unsigned)((n + 127) / 128), (unsigned)((n + APP_CH - 1) / APP_CH));
    KL(h, CLS_APPEND, s, k_lower_tmatvec_part<<<g, 128, 0, s>>>(W, ld, n, x, part, (int)pstride));
    KL(h, CLS_APPEND, s, k_lower_tmatvec_fin<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(part, (int)pstride, n, scale_ptr, out, ostride));
}

extern "C" int cgp_append_sample(cgp_handle* hh, int kind, const double* Xc, int64_t n_new, int d, const double* theta,
                                 double alpha, double* L, double* W, int64_t npad, int32_t* info, void* workspace,
                                 size_t ws_bytes, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!Xc) return -3;
    if (n_new < 1 || n_new > 0x7fffffff) return -4;
    if (d < 1 || d > CGP_MAX_D) return -5;
    if (!theta) return -6;
    if (!L) return -8;
    if (!W) return -9;
    if (npad != pad_n(n_new)) return -10;
    if (!info) return -11;
    if (!workspace) return -12;
    AppendLayout A = append_layout(n_new);
    if (A.total > ws_bytes) return -13;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    double* kv = reinterpret_cast<double*>(ws + A.o_k);
    double* lv = reinterpret_cast<double*>(ws + A.o_l);
    double* dd = reinterpret_cast<double*>(ws + A.o_dd);
    double* part = reinterpret_cast<double*>(ws + A.o_part);
    const int n = (int)(n_new - 1);  // old sample count = index of the new row
    CGP_CHECK(cudaMemsetAsync(info, 0, sizeof(int32_t), s));
    if (n > 0) {
        // k = k(x_new, X_old): one row of the cross kernel (no White term off the diagonal)
        int rc = launch_cross(h, kind, Xc + (long long)n * d, 1, Xc, n, d, theta, kv, n, 1, n, 0, 0.0, s);
        if (rc) return rc;
        KL(h, CLS_APPEND, s, k_lower_matvec<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(W, npad, n, kv, lv));
    }
    KL(h, CLS_APPEND, s, k_append_finish<<<1, 256, 0, s>>>(L, W, npad, n, lv, theta, d, alpha, dd, info));
    if (n > 0) launch_tmatvec(h, W, npad, n, lv, part, A.npad, dd, W + (long long)n * npad, 1, s);
    return (int)cudaGetLastError();
}

extern "C" int cgp_refresh_alpha(cgp_handle* hh, const double* L, const double* W, int64_t npad, int64_t n, const double* y,
                                 double* alpha_out, double* lml, void* workspace, size_t ws_bytes, void* stream) {
    CGP_ENTER(hh);
    if (!L) return -2;
    if (!W) return -3;
    if (n < 1 || n > 0x7fffffff || npad != pad_n(n)) return -5;
    if (!y) return -6;
    if (!alpha_out) return -7;
    if (!lml) return -8;
    if (!workspace) return -9;
    AppendLayout A = append_layout(n);
    if (A.total > ws_bytes) return -10;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    double* z = reinterpret_cast<double*>(ws + A.o_l);
    double* part = reinterpret_cast<double*>(ws + A.o_part);
    KL(h, CLS_APPEND, s, k_lower_matvec<<<(unsigned)((n + 7) / 8), 256, 0, s>>>(W, npad, (int)n, y, z));
    launch_tmatvec(h, W, npad, (int)n, z, part, A.npad, nullptr, alpha_out, 1, s);
    KL(h, CLS_APPEND, s, k_lml_from_factor<<<1, 256, 0, s>>>(L, npad, (int)n, z, lml));
    return (int)cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// k-NN routing
// ------------------------------------------------------------------------------------------------
// Tensor-core prefilter (knn_tc.cuh) in front of the exact kernel: worth it for large training sets and many queries.
// CGP_KNN_MODE = "exact" | "tc" forces one path (tests run both).
template <int DP>
static int launch_knn_tc(HandleImpl* h, const double* X, const long long* lab, int n, int d, int k, const double* Q, long long M,
                         long long* out, cudaStream_t s) {
    const int npad = (n + KTC_TILE - 1) / KTC_TILE * KTC_TILE;
    const size_t o_xn = align_up((size_t)npad * DP * 8, 256), o_st = align_up(o_xn + (size_t)npad * 4, 256);
    const int qcap = (int)std::min<long long>(M, 1 << 20);  // listed uncertified queries (more only carry the -1 mark)
    const size_t o_cnt = align_up(o_st + (DP + 1) * 4, 256), o_list = o_cnt + 256;
    const size_t need = o_list + (size_t)qcap * 8;
    if (need > h->knn_bytes) {  // grow-only scratch of the handle (split training set, norms, statistics)
        CGP_CHECK(cudaStreamSynchronize(s));
        if (h->knn_buf) CGP_CHECK(cudaFree(h->knn_buf));
        h->knn_buf = nullptr;
        h->knn_bytes = 0;
        CGP_CHECK(cudaMalloc(&h->knn_buf, need));
        h->knn_bytes = need;
    }
    char* buf = static_cast<char*>(h->knn_buf);
    float2* Xp = reinterpret_cast<float2*>(buf);
    float* xn = reinterpret_cast<float*>(buf + o_xn);
    float* stats = reinterpret_cast<float*>(buf + o_st);
    int* qcount = reinterpret_cast<int*>(buf + o_cnt);
    long long* qlist = reinterpret_cast<long long*>(buf + o_list);
    CGP_CHECK(cudaMemsetAsync(qcount, 0, sizeof(int), s));
    KL(h, CLS_KNN_PREP, s, k_knn_prep<DP><<<(unsigned)((npad + 255) / 256), 256, 0, s>>>(X, n, npad, d, Xp, xn));
    KL(h, CLS_KNN_PREP, s, k_knn_stats<DP><<<1, 1024, 0, s>>>(X, n, d, stats));
    const int sm = KTC_SMEM_BYTES(DP);
    CGP_CHECK(cudaFuncSetAttribute(k_knn_tc<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
    KL(h, CLS_KNN, s, k_knn_tc<DP><<<(unsigned)((M + KTC_QPB - 1) / KTC_QPB), 256, sm, s>>>(Xp, xn, stats, X, lab, n, npad, d, k, Q,
                                                                                           M, out, qlist, qcount, qcap));
    h->knn_qlist = qlist;
    h->knn_qcount = qcount;
    h->knn_qcap = qcap;
    return (int)cudaGetLastError();
}

template <int D>
static int launch_knn(HandleImpl* h, const double* X, const long long* lab, int n, int k, const double* Q, long long M, long long* out,
                      cudaStream_t s) {
    size_t sm = (size_t)KNN_TILE * D * 8 + KNN_TILE * 4;
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(k_knn<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return (int)e;
    }
    bool tc = k <= 3 && n >= 2048 && M >= 4096;
    if (h->knn_mode == 1) tc = false;
    if (h->knn_mode == 2) tc = k <= 3;
    unsigned grid = (unsigned)((M + KNN_THREADS - 1) / KNN_THREADS);
    if (tc) {
        int rc = D <= 8 ? launch_knn_tc<8>(h, X, lab, n, D, k, Q, M, out, s) : launch_knn_tc<16>(h, X, lab, n, D, k, Q, M, out, s);
        if (rc) return rc;
        // queries the prefilter could not certify: exact brute force, dense over the list (blocks past the count exit),
        // then a sweep for marks that did not fit the list (blocks without one exit at once)
        const unsigned lgrid = (unsigned)((h->knn_qcap + KNN_THREADS - 1) / KNN_THREADS);
        KL(h, CLS_KNN_EXACT, s, k_knn<D><<<lgrid, KNN_THREADS, sm, s>>>(X, lab, n, k, Q, M, out, 0, h->knn_qlist, h->knn_qcount,
                                                                        h->knn_qcap));
        KL(h, CLS_KNN_EXACT, s, k_knn<D><<<grid, KNN_THREADS, sm, s>>>(X, lab, n, k, Q, M, out, 1, nullptr, nullptr, 0));
        return (int)cudaGetLastError();
    }
    KL(h, CLS_KNN_EXACT, s, k_knn<D><<<grid, KNN_THREADS, sm, s>>>(X, lab, n, k, Q, M, out, 0, nullptr, nullptr, 0));
    return (int)cudaGetLastError();
}

extern "C" int cgp_knn_labels(cgp_handle* hh, const double* X, const int64_t* labels, int64_t n, int d, int k, int n_classes,
                              const double* Q, int64_t m, int64_t* out, void* stream) {
    CGP_ENTER(hh);
    if (!X) return -2;
    if (!labels) return -3;
    if (n <= 0 || n > 0x7fffffff) return -4;
    if (d < 1 || d > CGP_MAX_D) return -5;
    if (k < 1 || k > CGP_MAX_K || k > n) return -6;
    if (n_classes < 1) return -7;
    if (m < 0) return -9;
    if (m == 0) return 0;
    if (!Q) return -8;
    if (!out) return -10;
    cudaStream_t s = (cudaStream_t)stream;
    const long long* lab = reinterpret_cast<const long long*>(labels);
    long long* o = reinterpret_cast<long long*>(out);
    switch (d) {
#define KNN_CASE(D) \
    case D:         \
        return launch_knn<D>(h, X, lab, (int)n, k, Q, m, o, s);
        KNN_CASE(1) KNN_CASE(2) KNN_CASE(3) KNN_CASE(4) KNN_CASE(5) KNN_CASE(6) KNN_CASE(7) KNN_CASE(8)
        KNN_CASE(9) KNN_CASE(10) KNN_CASE(11) KNN_CASE(12) KNN_CASE(13) KNN_CASE(14) KNN_CASE(15) KNN_CASE(16)
#undef KNN_CASE
    }
    return -5;
}

// ------------------------------------------------------------------------------------------------
// routed prediction / EI scan
// ------------------------------------------------------------------------------------------------
struct ScoreLayout {
    int64_t mch;  // candidates per chunk (multiple of 128)
    int nblk, Tmax;
    int64_t npad_max;
    size_t o_perm, o_qs, o_sorted, o_hist, o_blkoff, o_coff, o_blkv, o_blkp, o_kstar, o_vpart, o_mraw, o_kmax, o_sel, o_cnt, o_lb, total;
};

static int64_t score_chunk_min() {
    static const int64_t v = [] {
        const char* e = getenv("CGP_SCORE_CHUNK_MIN");
        return e ? (int64_t)atoll(e) : (int64_t)16384;
    }();
    return v;
}
#define CGP_FIRST_PRUNE_CHUNK 2048  // the first chunk of a pruned scan is scored exhaustively (no bound yet): keep it small

static ScoreLayout score_layout(const int64_t* offs, int C, int d, int64_t m) {
    ScoreLayout L;
    L.npad_max = 0;
    for (int c = 0; c < C; ++c) L.npad_max = std::max<int64_t>(L.npad_max, pad_n(offs[c + 1] - offs[c]));
    L.Tmax = (int)(L.npad_max / CGP_TILE);
    // candidates per chunk: a 1 GB cross-kernel block, but at least CGP_SCORE_CHUNK_MIN (16384) candidates as long as the
    // block stays <= 8 GB: with giant clusters (configs[3]: n = 65536 -> 2048 candidates per GB) the survivors of the
    // pruned scan are too few per chunk to fill the 128-candidate tiles of the variance contraction
    int64_t mch = ((int64_t)1 << 27) / std::max<int64_t>(L.npad_max, 1);
    mch = std::max<int64_t>(mch, std::min<int64_t>(score_chunk_min(), ((int64_t)1 << 30) / std::max<int64_t>(L.npad_max, 1)));
    mch = std::max<int64_t>(CGP_TILE, std::min<int64_t>(mch, 65536));
    mch = mch / CGP_TILE * CGP_TILE;
    mch = std::min<int64_t>(mch, pad_n(std::max<int64_t>(m, 1)));
    L.mch = mch;
    L.nblk = (int)((m + BUCKET_CH - 1) / BUCKET_CH);
    if (L.nblk < 1) L.nblk = 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    L.o_perm = take((size_t)m * 8);
    L.o_qs = take((size_t)m * d * 8);
    L.o_sorted = take((size_t)m * 8);
    L.o_hist = take((size_t)C * L.nblk * 4);
    L.o_blkoff = take((size_t)C * L.nblk * 8);
    L.o_coff = take((size_t)(C + 1) * 8);
    L.o_blkv = take(1024 * 8);
    L.o_blkp = take(1024 * 8);
    L.o_kstar = take((size_t)mch * L.npad_max * 8);
    L.o_vpart = take((size_t)L.Tmax * mch * 8);
    L.o_mraw = take((size_t)mch * 8);
    L.o_kmax = take((size_t)mch * 8);
    L.o_sel = take((size_t)mch * 4);
    L.o_cnt = take(8);
    L.o_lb = take(8);
    L.total = off;
    return L;
}

extern "C" size_t cgp_score_workspace_bytes(const int64_t* offs, int C, int d, int64_t m) {
    if (!offs || C <= 0 || m < 0) return 0;
    return score_layout(offs, C, d, m).total;
}

static int run_score(HandleImpl* h, int kind, const double* X, const int64_t* offs, int C, int d, const double* theta,
                     const double* W, const double* alpha_vec, const double* y_mean, const double* y_std,
                     const double* y_max, const double* Q, const int64_t* qlabel, int64_t m, int64_t idx_base,
                     double* score, double* mean, double* std_, double* best_score, int64_t* best_idx,
                     int64_t* best_cluster, void* workspace, size_t ws_bytes, cudaStream_t s, bool prune = false,
                     const double* penalty = nullptr, double alpha = 0.0) {
    if (C > 4096) return -5;
    ScoreLayout L = score_layout(offs, C, d, m);
    if (L.total > ws_bytes) return -100;
    char* ws = static_cast<char*>(workspace);
    long long* perm = reinterpret_cast<long long*>(ws + L.o_perm);
    double* Qs = reinterpret_cast<double*>(ws + L.o_qs);
    double* sorted = reinterpret_cast<double*>(ws + L.o_sorted);
    int* hist = reinterpret_cast<int*>(ws + L.o_hist);
    long long* blkoff = reinterpret_cast<long long*>(ws + L.o_blkoff);
    long long* coff = reinterpret_cast<long long*>(ws + L.o_coff);
    double* blkv = reinterpret_cast<double*>(ws + L.o_blkv);
    long long* blkp = reinterpret_cast<long long*>(ws + L.o_blkp);
    double* kstar = reinterpret_cast<double*>(ws + L.o_kstar);
    double* vpart = reinterpret_cast<double*>(ws + L.o_vpart);
    double* mraw = reinterpret_cast<double*>(ws + L.o_mraw);
    double* kmaxv = reinterpret_cast<double*>(ws + L.o_kmax);
    int* sel = reinterpret_cast<int*>(ws + L.o_sel);
    int* cnt = reinterpret_cast<int*>(ws + L.o_cnt);
    double* lb = reinterpret_cast<double*>(ws + L.o_lb);
    const long long* ql = reinterpret_cast<const long long*>(qlabel);

    KL(h, CLS_BUCKET, s, k_bucket_hist<<<L.nblk, 256, (size_t)C * 4, s>>>(ql, m, C, L.nblk, hist));
    KL(h, CLS_BUCKET, s, k_bucket_scan<<<1, 1024, 0, s>>>(hist, blkoff, coff, C, L.nblk));
    KL(h, CLS_BUCKET, s, k_bucket_scatter<<<L.nblk, 32, (size_t)C * 8, s>>>(ql, Q, m, C, L.nblk, d, blkoff, perm, Qs));
    CGP_CHECK(cudaGetLastError());
    std::vector<long long> hoff(C + 1);
    CGP_CHECK(cudaMemcpyAsync(hoff.data(), coff, (size_t)(C + 1) * 8, cudaMemcpyDeviceToHost, s));
    CGP_CHECK(cudaStreamSynchronize(s));  // launch geometry of the per-cluster chunks depends on the counts

    const int P = d + 1;
    long long woff = 0;
    if (prune) KL(h, CLS_SCORE_FINISH, s, k_set_double<<<1, 1, 0, s>>>(lb, -INFINITY));
    bool first_chunk = prune;
    for (int c = 0; c < C; ++c) {
        const int64_t n = offs[c + 1] - offs[c];
        const int64_t npad = pad_n(n);
        const int T = (int)(npad / CGP_TILE);
        const double* Wc = W + woff;
        woff += npad * npad;
        const double* th = theta + (long long)c * P;
        const double* Xc = X + offs[c] * d;
        for (long long p0 = hoff[c], rows = 0; p0 < hoff[c + 1]; p0 += rows) {
            rows = std::min<long long>(first_chunk ? std::min<long long>(L.mch, CGP_FIRST_PRUNE_CHUNK) : L.mch, hoff[c + 1] - p0);
            first_chunk = false;
            const long long rows_pad = pad_n(rows);
            int rc = launch_cross(h, kind, Qs + p0 * d, rows, Xc, n, d, th, kstar, npad, rows_pad, npad, 0, 0.0, s);
            if (rc) return rc;
            KL(h, CLS_MEAN, s, k_mean<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(kstar, alpha_vec + offs[c], (int)n, (int)npad, rows, mraw,
                                                                              prune ? kmaxv : nullptr));
            if (prune) {  // mean-only bound -> survivors -> exact variance / score of the survivors -> tighter bound
                KL(h, CLS_SCORE_FINISH, s, k_prune_select<<<1, 1024, 0, s>>>(mraw, rows, th, d, y_mean, y_std, y_max, c, (double)n, lb,
                                                                              sel, cnt, sorted + p0, penalty, perm + p0, kmaxv,
                                                                              alpha));
                KL(h, CLS_PREDICT_VNORM, s, k_predict_vnorm_sel<<<(unsigned)((rows_pad / CGP_TILE) * T), CGP_GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(
                    Wc, kstar, vpart, (int)npad, L.mch, sel, cnt));
                KL(h, CLS_SCORE_FINISH, s, k_score_finish_sel<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(
                    vpart, L.mch, T, mraw, sel, cnt, th, d, y_mean, y_std, y_max, c, (double)n, sorted + p0, penalty, perm + p0));
                KL(h, CLS_SCORE_FINISH, s, k_lb_update<<<1, 1024, 0, s>>>(sorted + p0, rows, lb));
                continue;
            }
            // mean-only callers (predict(return_std=False)) skip the n_c^2-per-candidate variance contraction
            const bool need_var = std_ != nullptr || y_max != nullptr;
            if (need_var) FL(h, (double)rows * (double)n * (double)n);
            if (need_var)
                KL(h, CLS_PREDICT_VNORM, s, k_predict_vnorm<<<(unsigned)((rows_pad / CGP_TILE) * T), CGP_GEMM_THREADS, GEMM_SMEM_BYTES, s>>>(
                    Wc, kstar, vpart, (int)npad, L.mch));
            KL(h, CLS_SCORE_FINISH, s, k_score_finish<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(
                vpart, L.mch, need_var ? T : 0, mraw, rows, p0, perm, th, d, y_mean, y_std, y_max, c, (double)n, mean, std_, score, sorted, penalty));
        }
    }
    if (y_max && best_score && best_idx && best_cluster) {
        int nb = (int)std::min<long long>(1024, (m + 255) / 256);
        if (nb < 1) nb = 1;
        KL(h, CLS_ARGMAX, s, k_argmax_blocks<<<nb, 256, 0, s>>>(sorted, m, blkv, blkp));
        KL(h, CLS_ARGMAX, s, k_argmax_final<<<1, 256, 0, s>>>(blkv, blkp, nb, perm, coff, C, idx_base, best_score, reinterpret_cast<long long*>(best_idx),
                                         reinterpret_cast<long long*>(best_cluster)));
    }
    return (int)cudaGetLastError();
}

extern "C" int cgp_predict(cgp_handle* hh, int kind, const double* X, const int64_t* offs, int C, int d, const double* theta,
                           const double* W, const double* alpha_vec, const double* y_mean, const double* y_std,
                           const double* Q, const int64_t* qlabel, int64_t m, double* mean, double* std_, void* workspace,
                           size_t ws_bytes, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X) return -3;
    if (!offs || C <= 0) return -4;
    if (d < 1 || d > CGP_MAX_D) return -6;
    if (!theta) return -7;
    if (!W) return -8;
    if (!alpha_vec) return -9;
    if (!y_mean) return -10;
    if (!y_std) return -11;
    if (m < 0) return -14;
    if (m == 0) return 0;
    if (!Q) return -12;
    if (!qlabel) return -13;
    if (!mean) return -15;
    if (!workspace) return -17;
    return run_score(h, kind, X, offs, C, d, theta, W, alpha_vec, y_mean, y_std, nullptr, Q, qlabel, m, 0, nullptr, mean,
                     std_, nullptr, nullptr, nullptr, workspace, ws_bytes, (cudaStream_t)stream);
}

extern "C" int cgp_ei_argmax(cgp_handle* hh, int kind, const double* X, const int64_t* offs, int C, int d,
                             const double* theta, const double* W, const double* alpha_vec, const double* y_mean,
                             const double* y_std, const double* y_max, const double* Q, const int64_t* qlabel, int64_t m,
                             int64_t idx_base, const double* penalty, double* score, double* mean, double* std_,
                             double* best_score, int64_t* best_idx, int64_t* best_cluster, void* workspace, size_t ws_bytes,
                             void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X) return -3;
    if (!offs || C <= 0) return -4;
    if (d < 1 || d > CGP_MAX_D) return -6;
    if (!theta) return -7;
    if (!W) return -8;
    if (!alpha_vec) return -9;
    if (!y_mean) return -10;
    if (!y_std) return -11;
    if (!y_max) return -12;
    if (m <= 0) return -15;
    if (!Q) return -13;
    if (!qlabel) return -14;
    if (!best_score) return -20;
    if (!best_idx) return -21;
    if (!best_cluster) return -22;
    if (!workspace) return -23;
    return run_score(h, kind, X, offs, C, d, theta, W, alpha_vec, y_mean, y_std, y_max, Q, qlabel, m, idx_base, score, mean,
                     std_, best_score, best_idx, best_cluster, workspace, ws_bytes, (cudaStream_t)stream, false, penalty);
}

extern "C" int cgp_ei_argmax_pruned(cgp_handle* hh, int kind, const double* X, const int64_t* offs, int C, int d,
                                    const double* theta, double alpha, const double* W, const double* alpha_vec, const double* y_mean,
                                    const double* y_std, const double* y_max, const double* Q, const int64_t* qlabel,
                                    int64_t m, int64_t idx_base, const double* penalty, double* best_score, int64_t* best_idx,
                                    int64_t* best_cluster, void* workspace, size_t ws_bytes, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X) return -3;
    if (!offs || C <= 0) return -4;
    if (d < 1 || d > CGP_MAX_D) return -6;
    if (!theta) return -7;
    if (!(alpha >= 0.0)) return -8;
    if (!W) return -9;
    if (!alpha_vec) return -10;
    if (!y_mean) return -11;
    if (!y_std) return -12;
    if (!y_max) return -13;
    if (m <= 0) return -16;
    if (!Q) return -14;
    if (!qlabel) return -15;
    if (!best_score) return -17;
    if (!best_idx) return -18;
    if (!best_cluster) return -19;
    if (!workspace) return -20;
    return run_score(h, kind, X, offs, C, d, theta, W, alpha_vec, y_mean, y_std, y_max, Q, qlabel, m, idx_base, nullptr,
                     nullptr, nullptr, best_score, best_idx, best_cluster, workspace, ws_bytes, (cudaStream_t)stream, true,
                     penalty, alpha);
}

// ------------------------------------------------------------------------------------------------
// dense MSPE scan: value_i = (c0[c] - k*_i^T B_c k*_i + penalty_i) / n_c, minimum wins
// ------------------------------------------------------------------------------------------------
extern "C" size_t cgp_quadform_workspace_bytes(const int64_t* offs, int C, int d, int64_t m) {
    if (!offs || C <= 0 || m < 0) return 0;
    ScoreLayout L = score_layout(offs, C, d, m);
    return L.total + align_up((size_t)L.mch * L.npad_max * 8, 256);
}

extern "C" int cgp_quadform_scan(cgp_handle* hh, int kind, const double* X, const int64_t* offs, int C, int d,
                                 const double* theta, const double* Bmat, const double* c0_host, const double* Q,
                                 const int64_t* qlabel, int64_t m, int64_t idx_base, const double* penalty, double* value,
                                 double* best_value, int64_t* best_idx, int64_t* best_cluster, void* workspace,
                                 size_t ws_bytes, void* stream) {
    CGP_ENTER(hh);
    if (kind != 0 && kind != 1) return -2;
    if (!X) return -3;
    if (!offs || C <= 0 || C > 4096) return -4;
    if (d < 1 || d > CGP_MAX_D) return -6;
    if (!theta) return -7;
    if (!Bmat) return -8;
    if (!c0_host) return -9;
    if (m <= 0) return -12;
    if (!Q) return -10;
    if (!qlabel) return -11;
    if (!best_value) return -16;
    if (!best_idx) return -17;
    if (!best_cluster) return -18;
    if (!workspace) return -19;
    ScoreLayout L = score_layout(offs, C, d, m);
    const size_t o_y = L.total;
    if (o_y + align_up((size_t)L.mch * L.npad_max * 8, 256) > ws_bytes) return -100;
    cudaStream_t s = (cudaStream_t)stream;
    char* ws = static_cast<char*>(workspace);
    long long* perm = reinterpret_cast<long long*>(ws + L.o_perm);
    double* Qs = reinterpret_cast<double*>(ws + L.o_qs);
    double* sorted = reinterpret_cast<double*>(ws + L.o_sorted);
    int* hist = reinterpret_cast<int*>(ws + L.o_hist);
    long long* blkoff = reinterpret_cast<long long*>(ws + L.o_blkoff);
    long long* coff = reinterpret_cast<long long*>(ws + L.o_coff);
    double* blkv = reinterpret_cast<double*>(ws + L.o_blkv);
    long long* blkp = reinterpret_cast<long long*>(ws + L.o_blkp);
    double* kstar = reinterpret_cast<double*>(ws + L.o_kstar);
    double* qf = reinterpret_cast<double*>(ws + L.o_mraw);
    double* Y = reinterpret_cast<double*>(ws + o_y);
    const long long* ql = reinterpret_cast<const long long*>(qlabel);
    KL(h, CLS_BUCKET, s, k_bucket_hist<<<L.nblk, 256, (size_t)C * 4, s>>>(ql, m, C, L.nblk, hist));
    KL(h, CLS_BUCKET, s, k_bucket_scan<<<1, 1024, 0, s>>>(hist, blkoff, coff, C, L.nblk));
    KL(h, CLS_BUCKET, s, k_bucket_scatter<<<L.nblk, 32, (size_t)C * 8, s>>>(ql, Q, m, C, L.nblk, d, blkoff, perm, Qs));
    CGP_CHECK(cudaGetLastError());
    std::vector<long long> hoff(C + 1);
    CGP_CHECK(cudaMemcpyAsync(hoff.data(), coff, (size_t)(C + 1) * 8, cudaMemcpyDeviceToHost, s));
    CGP_CHECK(cudaStreamSynchronize(s));
    const int P = d + 1;
    long long boff = 0;
    for (int c = 0; c < C; ++c) {
        const int64_t n = offs[c + 1] - offs[c];
        const int64_t npad = pad_n(n);
        const double* Bc = Bmat + boff;
        boff += npad * npad;
        const double* th = theta + (long long)c * P;
        const double* Xc = X + offs[c] * d;
        for (long long p0 = hoff[c]; p0 < hoff[c + 1]; p0 += L.mch) {
            const long long rows = std::min<long long>(L.mch, hoff[c + 1] - p0);
            const long long rows_pad = pad_n(rows);
            int rc = launch_cross(h, kind, Qs + p0 * d, rows, Xc, n, d, th, kstar, npad, rows_pad, npad, 0, 0.0, s);
            if (rc) return rc;
            // Y = K* B  (B symmetric: K* B^T through the NT tile GEMM), q = rowdot(Y, K*)
            FL(h, 2.0 * (double)rows * (double)n * (double)n);
            KL(h, CLS_DGEMM, s, k_dgemm_nt<<<dim3((unsigned)(npad / CGP_TILE), (unsigned)(rows_pad / CGP_TILE)), CGP_GEMM_THREADS,
                                             GEMM_SMEM_BYTES, s>>>(kstar, Bc, Y, rows_pad, npad, (int)npad));
            KL(h, CLS_SCORE_FINISH, s, k_rowdot<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(Y, kstar, rows, (int)n, npad, qf));
            KL(h, CLS_SCORE_FINISH, s, k_quadform_finish<<<(unsigned)((rows + 255) / 256), 256, 0, s>>>(
                                           qf, rows, p0, perm, c0_host[c], (double)n, penalty, value, sorted));
        }
    }
    int nb = (int)std::min<long long>(1024, (m + 255) / 256);
    KL(h, CLS_ARGMAX, s, k_argmax_blocks<<<nb, 256, 0, s>>>(sorted, m, blkv, blkp));
    KL(h, CLS_ARGMAX, s, k_argmax_final<<<1, 256, 0, s>>>(blkv, blkp, nb, perm, coff, C, idx_base, best_value,
                                                          reinterpret_cast<long long*>(best_idx),
                                                          reinterpret_cast<long long*>(best_cluster)));
    return (int)cudaGetLastError();
}

extern "C" int cgp_rowdot(cgp_handle* hh, const double* A, const double* B, int64_t rows, int64_t n, int64_t ld, double* out,
                          void* stream) {
    CGP_ENTER(hh);
    if (!A) return -2;
    if (!B) return -3;
    if (rows < 0) return -4;
    if (n < 0 || n > 0x7fffffff || ld < n) return -5;
    if (!out) return -7;
    if (rows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    KL(h, CLS_SCORE_FINISH, s, k_rowdot<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(A, B, rows, (int)n, ld, out));
    return (int)cudaGetLastError();
}

// First maximum of score_sorted[0..m) in bucket order (cluster-major, candidate-index-minor): the winner rule of
// cgp_ei_argmax for scores the caller computed itself (dense MSPE scan).  perm[m] bucket position -> candidate index,
// coff[C+1] bucket offsets (device).  scratch: 2 * 1024 * 8 bytes.
extern "C" int cgp_argmax_sorted(cgp_handle* hh, const double* score_sorted, int64_t m, const int64_t* perm, const int64_t* coff,
                                 int C, int64_t idx_base, double* best_score, int64_t* best_idx, int64_t* best_cluster,
                                 void* scratch, size_t scratch_bytes, void* stream) {
    CGP_ENTER(hh);
    if (!score_sorted) return -2;
    if (m <= 0) return -3;
    if (!perm) return -4;
    if (!coff || C <= 0) return -5;
    if (!best_score) return -8;
    if (!best_idx) return -9;
    if (!best_cluster) return -10;
    if (!scratch || scratch_bytes < 2 * 1024 * 8) return -11;
    cudaStream_t s = (cudaStream_t)stream;
    double* blkv = static_cast<double*>(scratch);
    long long* blkp = reinterpret_cast<long long*>(static_cast<char*>(scratch) + 1024 * 8);
    int nb = (int)std::min<long long>(1024, (m + 255) / 256);
    KL(h, CLS_ARGMAX, s, k_argmax_blocks<<<nb, 256, 0, s>>>(score_sorted, m, blkv, blkp));
    KL(h, CLS_ARGMAX, s, k_argmax_final<<<1, 256, 0, s>>>(blkv, blkp, nb, reinterpret_cast<const long long*>(perm),
                                                          reinterpret_cast<const long long*>(coff), C, idx_base, best_score,
                                                          reinterpret_cast<long long*>(best_idx),
                                                          reinterpret_cast<long long*>(best_cluster)));
    return (int)cudaGetLastError();
}
