// Microbenchmark of the 64x64 diagonal factor+inverse leaf (k_potrf64_inv): the kernel that paces the
// critical path of the blocked Cholesky.  Includes the translation unit to reach the kernel.
#define FEBA_POTRF_PROF 1
#include "../../fish-eye_bundle_adjustment_b200/csrc/feba_chol.cu"

#include <cmath>
#include <cstdio>
#include <vector>

int main() {
    using namespace feba;
    const int n = 64, reps = 2000;
    std::vector<double> A(n * n), L(n * n), Li(n * n);
    // SPD test matrix: M = B B' + n I
    std::vector<double> B(n * n);
    unsigned s = 12345;
    for (auto& v : B) { s = s * 1664525u + 1013904223u; v = ((s >> 8) & 0xffff) / 65536.0 - 0.5; }
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double a = (i == j) ? n : 0.0;
            for (int k = 0; k < n; ++k) a += B[i + n * k] * B[j + n * k];
            A[i + n * j] = a;
        }
    double *dA, *dW, *dLi;
    int* dinfo;
    cudaMalloc(&dA, sizeof(double) * n * n);
    cudaMalloc(&dW, sizeof(double) * n * n * reps);
    cudaMalloc(&dLi, sizeof(double) * n * n);
    cudaMalloc(&dinfo, sizeof(int));
    cudaMemset(dinfo, 0, sizeof(int));
    cudaMemcpy(dA, A.data(), sizeof(double) * n * n, cudaMemcpyHostToDevice);
    for (int r = 0; r < reps; ++r) cudaMemcpy(dW + (size_t)r * n * n, dA, sizeof(double) * n * n, cudaMemcpyDeviceToDevice);
    constexpr size_t psmem = (2 * kBlk * (kBlk + 1) + 3 * 16 * 17 + kBlk) * sizeof(double);
    cudaFuncSetAttribute(k_potrf64_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_potrf64_inv<<<1, 256, psmem>>>(dW, n, dLi, dinfo);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 1; r < reps; ++r) k_potrf64_inv<<<1, 256, psmem>>>(dW + (size_t)r * n * n, n, dLi, dinfo);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(L.data(), dW + (size_t)(reps - 1) * n * n, sizeof(double) * n * n, cudaMemcpyDeviceToHost);
    cudaMemcpy(Li.data(), dLi, sizeof(double) * n * n, cudaMemcpyDeviceToHost);
    int info = 0;
    cudaMemcpy(&info, dinfo, sizeof(int), cudaMemcpyDeviceToHost);
    // checks: L L' = A (lower), Linv L = I
    double e1m = 0, e2m = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j <= i; ++j) {
            double a = 0, b = 0;
            for (int k = 0; k <= j; ++k) a += L[i + n * k] * L[j + n * k];
            for (int k = j; k <= i; ++k) b += Li[i + n * k] * L[k + n * j];
            e1m = fmax(e1m, fabs(a - A[i + n * j]) / n);
            e2m = fmax(e2m, fabs(b - (i == j ? 1.0 : 0.0)));
        }
    printf("k_potrf64_inv: %.2f us per launch (back to back, %d launches), info %d, |LL'-A| %.2e, |Linv L - I| %.2e (%s)\n",
           1e3 * ms / (reps - 1), reps - 1, info, e1m, e2m, cudaGetErrorString(cudaGetLastError()));
    long long prof[32];
    cudaMemcpyFromSymbol(prof, g_potrf_prof, sizeof(prof));
    printf("  load %lld\n", prof[1] - prof[0]);
    long long prev = prof[1];
    for (int kb = 0; kb < 4; ++kb) {
        printf("  kb %d: diag16 %lld  panel %lld  trailing %lld\n", kb, prof[10 + 3 * kb] - prev,
               prof[11 + 3 * kb] - prof[10 + 3 * kb], prof[12 + 3 * kb] - prof[11 + 3 * kb]);
        prev = prof[12 + 3 * kb];
    }
    printf("  store L %lld  diag inverses %lld  off-diagonal inverse %lld  store Linv %lld  total %lld\n", prof[6] - prof[5],
           prof[7] - prof[6], prof[8] - prof[7], prof[9] - prof[8], prof[9] - prof[0]);
    return 0;
}
