// PTX wrappers shared by the tcgen05 convolution kernels (unet_tc.cu, unet_tcx.cu): mbarriers, TMA / bulk copies,
// tensor-memory allocation, UMMA issue / commit, tcgen05.ld and the shared-memory matrix descriptor.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>

namespace ct {

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error must surface as a launch failure, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One lane of a converged warp.  The compiler recognises elect.sync and issues the uniform-datapath instructions
// (UTCHMMA, UTMALDG) once; a `lane == 0` test instead wraps each of them in an ELECT/branch loop (~60 clk per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.b32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (fp16 inputs, fp32 accumulate), issued by one thread for the CTA
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once every MMA issued so far by this thread has retired
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Packed fp32 pairs (FFMA2 / FADD2 / FMUL2): one issue slot for two IEEE operations, bit-identical to the scalar forms.
// The drain warps are issue bound (tensor-memory loads + accumulate + epilogue on 2 warps per scheduler), so every
// element-wise step of theirs works on channel pairs.
__device__ __forceinline__ float2 f2_fma(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n"
        "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
__device__ __forceinline__ float2 f2_add(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n"
        "add.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_mul(float2 a, float2 b) {
    float2 d;
    asm("{ .reg .b64 ra, rb, rd; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n"
        "mul.rn.f32x2 rd, ra, rb; mov.b64 {%0, %1}, rd; }"
        : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return d;
}
__device__ __forceinline__ float2 f2_splat(float v) { return make_float2(v, v); }

// bias -> LeakyReLU / ReLU -> BatchNorm(eval) of one channel pair (unet3d.py:117-119): acc * inv_scale + bias, activation,
// * scale + shift.  ep = shared-memory table [3][n] (bias | scale | shift), ch even.  alpha in [0, 1], so
// t > 0 ? t : alpha t == max(t, alpha t).  Updates the running max|output|.
__device__ __forceinline__ float2 block_epilogue(float2 acc, float inv_scale, float alpha, const float* ep, int n, int ch, float& amax) {
    float2 t = f2_fma(acc, f2_splat(inv_scale), *reinterpret_cast<const float2*>(ep + ch));
    const float2 ta = f2_mul(t, f2_splat(alpha));
    t = make_float2(fmaxf(t.x, ta.x), fmaxf(t.y, ta.y));
    const float2 o = f2_fma(t, *reinterpret_cast<const float2*>(ep + n + ch), *reinterpret_cast<const float2*>(ep + 2 * n + ch));
    amax = fmaxf(amax, fmaxf(fabsf(o.x), fabsf(o.y)));
    return o;
}
// x s -> fp16 pair images hi = RN16(x s), lo' = RN16((x s - hi) 2^11); the difference and its scaling are exact
__device__ __forceinline__ void split_pair2(float2 x, float s, uint32_t& hi, uint32_t& lo) {
    const float2 os = f2_mul(x, f2_splat(s));
    const __half2 h = __floats2half2_rn(os.x, os.y);
    const float2 hf = __half22float2(h);
    const float2 d = f2_fma(hf, f2_splat(-2048.f), f2_mul(os, f2_splat(2048.f)));
    const __half2 l = __floats2half2_rn(d.x, d.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// shared-memory matrix descriptor, no swizzle, K-major: ((8, m), 2) : ((16 B, SBO), LBO)   [units of 16 B]
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr_bytes, uint32_t lbo16, uint32_t sbo16) {
    const uint32_t lo = ((addr_bytes >> 4) & 0x3FFFu) | ((lbo16 & 0x3FFFu) << 16);
    const uint32_t hi = (sbo16 & 0x3FFFu) | (1u << 14);            // descriptor version 1 (sm_100)
    return ((uint64_t)hi << 32) | lo;
}

}  // namespace ct
