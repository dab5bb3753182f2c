// Development probe (not part of the product): phase timestamps of k_potrf_diag on one 128x128 SPD block.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DDIAG_TIMING -o tools/dev/diag_timing tools/dev/diag_timing.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
__device__ unsigned long long g_diag_t[64];
#include "../../cgp_b200/csrc/chol_diag.cuh"

int main(int argc, char** argv) {
    (void)argc; (void)argv;
    const int n = 128;
    std::vector<double> A(n * n);
    srand(1);
    std::vector<double> X(n * 3);
    for (auto& v : X) v = rand() / (double)RAND_MAX;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int p = 0; p < 3; ++p) s += (X[i * 3 + p] - X[j * 3 + p]) * (X[i * 3 + p] - X[j * 3 + p]) / 0.25;
            double k = sqrt(3 * s);
            A[i * n + j] = (1 + k) * exp(-k) + (i == j ? 0.02 : 0);
        }
    double *dA, *dU, *dD, *dld;
    int* dinfo;
    MatMeta m{};
    m.n = n; m.npad = n;
    MatMeta* dm;
    cudaMalloc(&dA, n * n * 8); cudaMalloc(&dU, n * n * 8); cudaMalloc(&dD, n * n * 8); cudaMalloc(&dld, 64);
    cudaMalloc(&dinfo, 4); cudaMalloc(&dm, sizeof(MatMeta));
    cudaMemcpy(dm, &m, sizeof(MatMeta), cudaMemcpyHostToDevice);
    cudaMemset(dinfo, 0, 4);
    cudaFuncSetAttribute(k_potrf_diag, cudaFuncAttributeMaxDynamicSharedMemorySize, DIAG_SMEM_BYTES);

    for (int rep = 0; rep < 3; ++rep) {
        cudaMemcpy(dA, A.data(), n * n * 8, cudaMemcpyHostToDevice);
        cudaEvent_t a, b;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a);
        cudaMemsetAsync(dU, 0xFF, n * n * 8);
        cudaMemsetAsync(dD, 0xFF, n * n * 8);
        unsigned long long z[64] = {0};
        cudaMemcpyToSymbol(g_diag_t, z, sizeof(z));
        k_potrf_diag<<<1, 256, DIAG_SMEM_BYTES>>>(dA, dU, dD, dld, dinfo, dm, 0, 1);
        cudaEventRecord(b);
        cudaDeviceSynchronize();
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        unsigned long long t[64];
        cudaMemcpyFromSymbol(t, g_diag_t, sizeof(t));
        printf("rep %d: %.1f us, err=%s\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()));
        for (int i = 1; i < 20; ++i) if (t[i]) printf("  phase %2d: +%6llu cycles (total %llu)\n", i, t[i] - t[i - 1], t[i] - t[0]);
    }
    std::vector<double> L(n * n);
    cudaMemcpy(L.data(), dA, n * n * 8, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double s = 0;
            for (int k = 0; k <= j; ++k) s += L[i * n + k] * L[j * n + k];
            err = fmax(err, fabs(s - A[i * n + j]));
        }
    printf("max |L L^T - A| = %.3e\n", err);
    // D L = I, U = D^T, log-determinant
    std::vector<double> Dm(n * n), Um(n * n);
    cudaMemcpy(Dm.data(), dD, n * n * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Um.data(), dU, n * n * 8, cudaMemcpyDeviceToHost);
    double e2 = 0, e3 = 0, ldet = 0, ldg;
    for (int i = 0; i < n; ++i) {
        ldet += log(L[i * n + i]);
        for (int j = 0; j < n; ++j) {
            double s = 0;
            for (int k = 0; k < n; ++k) s += Dm[i * n + k] * L[k * n + j];
            e2 = fmax(e2, fabs(s - (i == j)));
            e3 = fmax(e3, fabs(Um[i * n + j] - Dm[j * n + i]));
        }
    }
    cudaMemcpy(&ldg, dld, 8, cudaMemcpyDeviceToHost);
    int inf;
    cudaMemcpy(&inf, dinfo, 4, cudaMemcpyDeviceToHost);
    printf("max |D L - I| = %.3e, max |U - D^T| = %.3e, logdet err = %.3e, info = %d\n", e2, e3, fabs(ldet - ldg), inf);
    return 0;
}
