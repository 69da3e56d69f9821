// Microbenchmark: fp64 vector (DFMA) and fp64 tensor (DMMA m8n8k4 / m16n8k8) throughput on one B200.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_kernel(double* out, int iters) {
    double a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3 + i;
    const double b = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b, c);
    }
    double s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void ffma_kernel(float* out, int iters) {
    float a[8];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-3f + i;
    const float b = 1.0000001f, c = 1e-9f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b, c);
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma884_kernel(double* out, int iters) {
    double c[4][2];
    for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
    double a = threadIdx.x * 1e-3, b = 1.0 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void dmma1688_kernel(double* out, int iters) {
    double c[2][4];
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.0;
    double a[4], b[2];
    for (int j = 0; j < 4; ++j) a[j] = threadIdx.x * 1e-3 + j;
    b[0] = 1.0 + threadIdx.x * 1e-6; b[1] = 0.5;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 2; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                         : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
    }
    double s = 0;
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float time_it(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 8 * 1024 * 8);
    const int blocks = 148 * 2, threads = 1024, iters = 20000;
    float ms = time_it([&] { dfma_kernel<<<blocks, threads>>>(out, iters); });
    double fl = 2.0 * blocks * threads * 8.0 * iters;
    printf("DFMA   %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
    ms = time_it([&] { ffma_kernel<<<blocks, threads>>>((float*)out, iters); });
    printf("FFMA   %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
    ms = time_it([&] { dmma884_kernel<<<blocks, threads>>>(out, iters); });
    fl = 2.0 * blocks * (threads / 32) * 4.0 * iters * (8 * 8 * 4);
    printf("DMMA m8n8k4   %.3f ms  %.2f TFLOP/s (%.1f FMA/clk/SM)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
    ms = time_it([&] { dmma1688_kernel<<<blocks, threads>>>(out, iters); });
    fl = 2.0 * blocks * (threads / 32) * 2.0 * iters * (16 * 8 * 8);
    printf("DMMA m16n8k8  %.3f ms  %.2f TFLOP/s (%.1f FMA/clk/SM)\n", ms, fl / ms / 1e9, fl / 2 / (ms * 1e-3) / 148 / 1.965e9);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
