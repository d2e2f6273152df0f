// Microbenchmark: FP64 tensor (DMMA) issue rate per mma.sync shape, and plain DFMA rate, on sm_100a.
// Used to pick the shape for the Cholesky trailing-update kernel and to state an FP64 roofline.
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE>
__global__ void k_dmma(double* out, int iters) {
    double c[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[8], b[4];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    for (int i = 0; i < 4; ++i) b[i] = threadIdx.x * 1e-4 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            if (SHAPE == 0)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a[0]), "d"(b[0]));
            else if (SHAPE == 1)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            else if (SHAPE == 2)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            else if (SHAPE == 3)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j) c[t][j] = fma(a[j], b[j], c[t][j]);
            }
        }
    }
    double s = 0.0;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE>
void run(const char* name, double flop_per_inst_per_warp, int warps_per_sm) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int threads = 128, blocks = sms * warps_per_sm / 4, iters = 20000;
    double* out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dmma<SHAPE><<<blocks, threads>>>(out, 100);
    cudaEventRecord(e0);
    k_dmma<SHAPE><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flop = (double)blocks * (threads / 32) * iters * 8.0 * flop_per_inst_per_warp;
    printf("%-12s warps/SM %2d : %8.3f ms  %7.2f TFLOP/s  (%s)\n", name, warps_per_sm, ms, flop / ms / 1e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}

int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("m8n8k4", 2.0 * 8 * 8 * 4, w);
        run<1>("m16n8k4", 2.0 * 16 * 8 * 4, w);
        run<2>("m16n8k8", 2.0 * 16 * 8 * 8, w);
        run<3>("m16n8k16", 2.0 * 16 * 8 * 16, w);
        run<4>("dfma x4", 2.0 * 32 * 4, w);
    }
    return 0;
}
