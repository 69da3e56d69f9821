// Latency microbenchmarks (one warp unless noted): dependent DFMA, fp64 reciprocal, shuffle, LDS, barriers, DMMA.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* t, double seed) {
    __shared__ double sm[1024];
    const int tid = threadIdx.x;
    sm[tid] = seed + tid;
    __syncthreads();
    double a = seed + tid * 1e-3, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    const int IT = 512;
    if (tid < 32) {
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < IT; ++i) a = fma(a, b, c);
        t1 = clock64(); if (tid == 0) t[0] = (t1 - t0) / IT;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < IT; ++i) a = 1.0 / a + 1.5;
        t1 = clock64(); if (tid == 0) t[1] = (t1 - t0) / IT;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < IT; ++i) a = __shfl_sync(0xffffffffu, a, (tid + 1) & 31);
        t1 = clock64(); if (tid == 0) t[2] = (t1 - t0) / IT;
        int idx = tid;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < IT; ++i) idx = (int)sm[idx & 1023] & 1023;
        t1 = clock64(); if (tid == 0) t[3] = (t1 - t0) / IT;
        a += idx;
        float f = (float)a;
        t0 = clock64();
#pragma unroll 16
        for (int i = 0; i < IT; ++i) f = fmaf(f, 1.0000001f, 1e-9f);
        t1 = clock64(); if (tid == 0) t[6] = (t1 - t0) / IT;
        a += f;
        double cc[4] = {a, a, a, a}, aa[4] = {a, b, c, a}, bb[2] = {b, c};
        t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < IT; ++i)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+d"(cc[0]), "+d"(cc[1]), "+d"(cc[2]), "+d"(cc[3]) : "d"(aa[0]), "d"(aa[1]), "d"(aa[2]), "d"(aa[3]), "d"(bb[0]), "d"(bb[1]));
        t1 = clock64(); if (tid == 0) t[7] = (t1 - t0) / IT;
        a += cc[0] + cc[1] + cc[2] + cc[3];
    }
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < 256; ++i) __syncthreads();
    t1 = clock64(); if (tid == 0) t[4] = (t1 - t0) / 256;
    if (tid < 384) {
        t0 = clock64();
        for (int i = 0; i < 256; ++i) asm volatile("bar.sync 1, 384;" ::: "memory");
        t1 = clock64(); if (tid == 0) t[5] = (t1 - t0) / 256;
    }
    out[tid] = a;
}
int main() {
    double* out; long long* t; cudaMalloc(&out, 8192); cudaMallocManaged(&t, 128);
    k<<<1, 1024>>>(out, t, 1.0); cudaDeviceSynchronize();
    k<<<1, 1024>>>(out, t, 1.0); cudaDeviceSynchronize();
    printf("DFMA dep %lld  rcp64+add %lld  shfl(double) %lld  LDS.64 chase(+cvt) %lld  syncthreads(1024) %lld  bar(384) %lld  FFMA dep %lld  DMMA16x8x8 dep %lld\n",
           t[0], t[1], t[2], t[3], t[4], t[5], t[6], t[7]);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
