// microbenchmark: FP64 mma.sync throughput on sm_100a (m8n8k4, m16n8k8, m16n8k16) vs plain DFMA
#include <cstdio>
#include <cuda_runtime.h>
template <int SHAPE>
__global__ void __launch_bounds__(512) k(double *out, int iters, double a0, double b0) {
    const int lane = threadIdx.x & 31;
    double a[8], b[4];
    for (int i = 0; i < 8; i++) a[i] = a0 + lane * 1e-3 + i;
    for (int i = 0; i < 4; i++) b[i] = b0 + lane * 1e-3 + i;
    double c[8][4];
    for (int t = 0; t < 8; t++) for (int i = 0; i < 4; i++) c[t][i] = 0.0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int t = 0; t < 8; t++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[t][0]), "+d"(c[t][1]) : "d"(a[0]), "d"(b[0]));
            } else if (SHAPE == 1) {
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else if (SHAPE == 2) {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            } else if (SHAPE == 3) {
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[t][0]), "+d"(c[t][1]), "+d"(c[t][2]), "+d"(c[t][3]) : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else {
#pragma unroll
                for (int i = 0; i < 4; i++) c[t][i] = fma(a[i], b[i], c[t][i]);
            }
        }
    }
    double s = 0; for (int t = 0; t < 8; t++) for (int i = 0; i < 4; i++) s += c[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int SHAPE> void run(const char *name, double fma_per_warp_inst, int warps) {
    double *out; cudaMalloc(&out, 148 * 4 * 512 * 8);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = 148 * (warps > 16 ? 2 : 1), block = (warps > 16 ? warps / 2 : warps) * 32;
    k<SHAPE><<<grid, block>>>(out, 100, 1.0, 2.0);
    cudaEventRecord(e0);
    k<SHAPE><<<grid, block>>>(out, iters, 1.0, 2.0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)grid * (block / 32) * iters * 8 * fma_per_warp_inst;
    printf("%-12s warps/SM %2d: %8.3f ms  %7.2f TFLOP/s  (%6.1f FMA/clk/SM at 1.965 GHz)  err=%s\n", name, warps, ms, 2 * fmas / ms * 1e-9,
           fmas / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out);
}
int main() {
    for (int w : {4, 8, 16, 32}) {
        run<0>("m8n8k4", 256, w);
        run<3>("m16n8k4", 512, w);
        run<1>("m16n8k8", 1024, w);
        run<2>("m16n8k16", 2048, w);
        run<4>("dfma x4", 128, w);
    }
    return 0;
}
