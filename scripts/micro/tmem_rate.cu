// Tensor-memory read throughput of one SM: W warps (W/4 per lane quarter) loop over tcgen05.ld of width x4 / x32 / x64.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_rate tmem_rate.cu && ./tmem_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int X> struct Ld;
template <> struct Ld<4> { static __device__ __forceinline__ uint32_t go(uint32_t a) { uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]) : "r"(a));
    return r[0] ^ r[1] ^ r[2] ^ r[3]; } };
template <> struct Ld<16> { static __device__ __forceinline__ uint32_t go(uint32_t a) { uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]) : "r"(a));
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= r[i];
    return x; } };
template <> struct Ld<32> { static __device__ __forceinline__ uint32_t go(uint32_t a) { uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
          "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]) : "r"(a));
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= r[i];
    return x; } };

template <int X, int BATCH>
__global__ void k(int iters, unsigned long long* out, uint32_t* sink) {
    __shared__ uint32_t base_s;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t t0 = base_s + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    __syncthreads();
    const long long c0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int b = 0; b < BATCH; ++b) acc ^= Ld<X>::go(t0 + ((b * X) & 511 & ~(X - 1)) % (512 - X + 1));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    }
    const long long c1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0) out[0] = (unsigned long long)(c1 - c0);
    if (acc == 0x12345) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base_s) : "memory");
}

template <int X, int BATCH>
void run(int warps) {
    unsigned long long* d; uint32_t* s;
    cudaMalloc(&d, 8); cudaMalloc(&s, 4);
    const int iters = 2000;
    k<X, BATCH><<<1, warps * 32>>>(iters, d, s);
    unsigned long long h = 0;
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    const double bytes = (double)iters * BATCH * X * 4 * 32 * warps;
    printf("x%-3d batch %2d warps %2d: %8.1f clk/iter  %7.1f B/clk/SM  (%s)\n", X, BATCH, warps, (double)h / iters, bytes / (double)h,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(s);
}

int main() {
    for (int w : {4, 8, 16}) { run<4, 18>(w); run<16, 8>(w); run<32, 4>(w); }
    return 0;
}
