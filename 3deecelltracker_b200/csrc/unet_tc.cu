// tcgen05 implicit-GEMM 3x3x3 'same' convolution with fused bias -> LeakyReLU/ReLU -> BatchNorm(eval).
//
// Replaces one Conv3D + LeakyReLU + BatchNormalization block of unet3d.py:117-119 (or the ReLU variant
// :139-140) for a batch of independent tiles, on the 5th-generation tensor cores.
//
// GEMM view.  M = output voxels, N = Cout, K = 27 taps x Cin.  One CTA owns an output block of
// BX (x) x 16 (y) x 8 (z) voxels = BX UMMA tiles of M = 128 (row r of a tile is voxel y = r / 8, z = r % 8),
// all Cout channels, accumulated in tensor memory.
//
// A operand = im2col WITHOUT materialisation.  Per 4-channel input chunk, TMA loads the haloed input block
// (BX+2) x 18 x 10 voxels x 16 B straight from the c4-blocked activation tensor into shared memory
// (5-D tensor map; out-of-bounds coordinates are zero filled, which IS the Keras 'same' padding of the tile).
// A voxel-chunk is 16 bytes = one row of a no-swizzle K-major UMMA core matrix; 8 consecutive z voxels are
// one core matrix, the 16 y rows of the block sit at a uniform stride of 160 B (SBO).  The A tile of tap
// (dx,dy,dz) is therefore the SAME shared-memory block addressed from a shifted start address; the two K
// halves of one K = 8 MMA are two different taps (LBO = address distance between the taps).
//
// Precision.  The reference computes in fp32 (TF-CPU); parity target 1e-4 on probabilities after 15 stacked
// convolutions.  One pass of any 16/19-bit tensor-core format cannot meet it, so both operands are split into two
// fp16 images x * s = hi + lo' * 2^-11  (hi = RN_fp16(x s), lo' = RN_fp16((x s - hi) * 2^11); 22 significant bits),
// with s a power of two chosen per tile from the running max|x| bound of the source buffer (slab header, see
// unet_common.cuh) so that max|x s| lies in [2^13, 2^14): no overflow, and fp16's subnormal threshold sits 2^27 below
// the largest element.  Weights are split the same way on the host with one power-of-two scale per layer.  The cross
// terms are kept in separate accumulator column groups because they carry different weights:
//     MMA1 = A_hi  x [B_hi | B_lo']  -> columns [0,N) (x1) and [N,2N) (x 2^-11)
//     MMA2 = A_lo' x [B_hi (| B_lo')] -> columns [N,2N) (x 2^-11)  (and [2N,3N) (x 2^-22) when Cout = 8)
// K = 16 per MMA: one K step covers two taps x EIGHT input channels, so the shared-memory operand traffic -- the
// measured bound of this kernel, see DESIGN.md -- is half that of a TF32 split (K = 8).  Activations stay single fp32
// tensors in HBM: worker warps convert the TMA-landed fp32 block in place (two 4-channel fp32 planes become the
// 8-channel fp16 hi plane and lo plane).
//
// The tensor core adds with truncation, so accumulation chains in tensor memory are kept short: every stage (28 MMAs
// per accumulator) goes into a FRESH accumulator set (ping-pong) that the worker warps drain into fp32 registers
// (round-to-nearest) while the next stage's MMAs run.
//
// Warp roles (320 threads, persistent CTA): warp 0 = TMA producer + TMEM allocator, warp 1 = MMA issuer (one elected
// lane), warps 2-9 = operand conversion, accumulator drain, epilogue (scale back, bias/activation/BN, 16-byte
// channel-chunk stores, max|value| bound of the output).  Stage ring: full (TMA landed) -> conv (converted) -> empty.
#include "unet_common.cuh"
#include "tc_ptx.cuh"
#include <cmath>
#include <cstring>
#include <mutex>

namespace ct {

constexpr int TC_SYH = 18, TC_SZH = 10;          // haloed block extent in y and z (16 + 2, 8 + 2)
constexpr int TC_PAIRS = 14;                     // 27 taps -> 14 K=16 steps (tap 0 is paired with a zero column)

// voxel offset of tap t = (dx*3 + dy)*3 + dz inside the haloed block (x stride 180, y stride 10)
__host__ __device__ constexpr int tap_off(int t) { return ((t / 9) * TC_SYH + (t / 3) % 3) * TC_SZH + t % 3; }
// K step p covers taps (first, second); step 0 is (tap 0, zero weights read at tap 1's address)
__host__ __device__ constexpr int pair_first(int p) { return p == 0 ? 0 : 2 * p - 1; }
__host__ __device__ constexpr int pair_second(int p) { return p == 0 ? 1 : 2 * p; }

template <int N, int BX, int STAGES>
struct TcCfg {
    static constexpr bool N8 = (N == 8);
    // B rows per K half: [hi | lo'] -- plus a second copy of hi when Cout = 8, so that the 16-wide MMA2 can read
    // [lo' | hi] from row 8 and land in columns of its own (one accumulate flag covers a whole MMA)
    static constexpr int NPR = N8 ? 24 : 2 * N;
    static constexpr int N1 = 2 * N;                                   // width of MMA1 = A_hi  x [B_hi | B_lo']
    static constexpr int N2 = N8 ? 16 : N;                             // width of MMA2 = A_lo' x [B_lo' | B_hi] or B_hi
    static constexpr int NPD = N8 ? 32 : 2 * N;                        // accumulator columns per M tile
    static constexpr int SXH = BX + 2;
    static constexpr int PLANE = SXH * TC_SYH * TC_SZH * 16;           // bytes of one plane (fp32 x4 in, fp16 x8 out)
    static constexpr int B_BYTES = TC_PAIRS * NPR * 32;                // [pair][k half][NPR rows][8 fp16]
    static constexpr int STAGE = 2 * PLANE + B_BYTES;
    static constexpr int SET_COLS = BX * NPD;                          // one accumulator set
    static constexpr int COLS_NEEDED = 2 * SET_COLS;                   // ping-pong sets
    static constexpr int TMEM_COLS = COLS_NEEDED <= 32 ? 32 : COLS_NEEDED <= 64 ? 64 : COLS_NEEDED <= 128 ? 128
                                     : COLS_NEEDED <= 256 ? 256 : 512;
    static constexpr int SMEM = STAGES * STAGE + 1024;                 // + slack to align the ring to 1 KiB
    static constexpr int TILES_PER_HALF = BX / 2;                      // M tiles drained by one worker thread
    static_assert(BX % 2 == 0, "BX must be even (two worker halves)");
    static_assert(COLS_NEEDED <= 512, "accumulators exceed tensor memory");
    static_assert(PLANE % 128 == 0 && B_BYTES % 128 == 0, "stage parts must stay 128-byte aligned");
    static_assert(N1 % 16 == 0 && N1 <= 256 && N2 % 16 == 0, "UMMA N out of range for M = 128");
    static_assert(SMEM <= 232448, "shared memory ring too large");
};

// ---------------------------------------------------------------------------------------------
// kernel
// ---------------------------------------------------------------------------------------------
constexpr int TC_WORKERS = 256;                  // 8 worker warps: two per TMEM lane quarter
constexpr int TC_THREADS = 64 + TC_WORKERS;

struct TcGeom {
    int cin8, X, Y, Z, nbx, nby, nbz, units;     // cin8 = K stages (8 input channels each); units = blocks * tiles
    int dst_c4off;
    size_t dst_tile_stride4;
    const float* amax_src;                       // slab header slot of the source buffer (tile stride = slab stride)
    float* amax_dst;                             // ... of the destination buffer
    size_t slab_stride;                          // floats
    float w_inv_scale;                           // 1 / weight scale of the layer
    // split-fp16 buffers (unet_common.cuh): header scale slots and the a-priori output bound of the block
    const float* scale_src;
    float* scale_dst;
    float bound_p, bound_q;
};

struct TcUnit { int x0, y0, z0, tile; };
__device__ __forceinline__ TcUnit tc_unit(int u, const TcGeom& g, int bx) {
    TcUnit r;
    r.x0 = (u % g.nbx) * bx; u /= g.nbx;
    r.y0 = (u % g.nby) * 16; u /= g.nby;
    r.z0 = (u % g.nbz) * 8;
    r.tile = u / g.nbz;
    return r;
}

// Persistent CTA (one per SM).  Work unit = BX x 16 x 8 output voxels x all Cout of one tile; units are taken
// round-robin.  `g` counts K stages (8-channel input chunks) across all units of the CTA: ring slot g % STAGES,
// accumulator set g & 1.  Every stage accumulates into a FRESH tensor-memory set (28 MMAs per accumulator) that
// the worker warps drain into fp32 registers while the next stage's MMAs run: the tensor core adds with
// truncation, so long in-TMEM accumulation chains would bias the result by ~(number of MMAs) * 2^-25 (measured 2e-5
// of scale at K = 3456 with a TF32 split); register accumulation rounds to nearest.
// SRC_SPLIT: the source buffer already holds the fp16 hi / lo' operand images (the workers skip the conversion and the
// MMA issuer waits for the TMA directly).  DST_SPLIT: the epilogue writes the destination in that form.
template <int N, int BX, int STAGES, bool SRC_SPLIT, bool DST_SPLIT>
__global__ void __launch_bounds__(TC_THREADS, 1)
conv3_tc_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wpack,
                const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                float alpha, float4* __restrict__ dst, const TcGeom geo) {
    using Cfg = TcCfg<N, BX, STAGES>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[STAGES], bar_conv[STAGES], bar_empty[STAGES], bar_acc_full[2], bar_acc_empty[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float ep_s[3][N];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int cin8 = geo.cin8;
    const int n_units = ((int)blockIdx.x < geo.units) ? (geo.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_stages = n_units * cin8;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_conv[s], TC_WORKERS / 32);
            mbar_init(&bar_empty[s], 1);
        }
#pragma unroll
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_acc_full[a], 1);
            mbar_init(&bar_acc_empty[a], TC_WORKERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, Cfg::TMEM_COLS);
    if (threadIdx.x >= 64) {
        for (int i = threadIdx.x - 64; i < 3 * N; i += TC_WORKERS)
            ep_s[i / N][i % N] = (i < N) ? bias[i] : (i < 2 * N ? scale[i - N] : shift[i - 2 * N]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        // ---------------- TMA producer
        if (elect_one()) {
            int g = 0;
            for (int k = 0; k < n_units; ++k) {
                const TcUnit un = tc_unit((int)blockIdx.x + k * (int)gridDim.x, geo, BX);
                for (int c = 0; c < cin8; ++c, ++g) {
                    const int s = g % STAGES, use = g / STAGES;
                    if (use > 0) mbar_wait(&bar_empty[s], (use - 1) & 1);
                    uint8_t* st = ring + (size_t)s * Cfg::STAGE;
                    mbar_expect_tx(&bar_full[s], 2 * Cfg::PLANE + Cfg::B_BYTES);
                    // two 4-channel fp32 planes (a missing second plane of a 4-channel input is zero filled)
                    tma_load_5d(st, &tmap, &bar_full[s], (un.z0 - 1) * 4, un.y0 - 1, un.x0 - 1, 2 * c, un.tile);
                    bulk_load(st + 2 * Cfg::PLANE, wpack + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (elect_one()) {
            // instruction descriptors: D = f32, A = B = f16 (format 0), both K-major, M = 128, N = 2N / N2
            constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(Cfg::N1 >> 3) << 17) | (8u << 24);
            constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(Cfg::N2 >> 3) << 17) | (8u << 24);
            // descriptor high words: SBO | version 1.  A: 16 y rows 160 B apart; B: 8-row groups 128 B apart
            constexpr uint64_t a_hi_word = (uint64_t)((uint32_t)TC_SZH | (1u << 14)) << 32;
            constexpr uint64_t b_hi_word = (uint64_t)(8u | (1u << 14)) << 32;
            const uint32_t ring16 = smem_u32(ring) >> 4;           // everything below in 16-byte units (< 2^14)
            for (int g = 0; g < n_stages; ++g) {
                const int s = g % STAGES, use = g / STAGES, set = g & 1, use_a = g >> 1;
                if (use_a > 0) mbar_wait(&bar_acc_empty[set], (use_a - 1) & 1);
                mbar_wait(SRC_SPLIT ? &bar_full[s] : &bar_conv[s], use & 1);
                tc_fence_after();
                const uint32_t a_hi = ring16 + (uint32_t)s * (Cfg::STAGE / 16), a_lo = a_hi + Cfg::PLANE / 16;
                const uint32_t b_base = a_hi + 2 * (Cfg::PLANE / 16);
                const uint32_t d_set = tmem_base + (uint32_t)set * Cfg::SET_COLS;
#pragma unroll
                for (int p = 0; p < TC_PAIRS; ++p) {
                    const uint32_t lbo = (uint32_t)(tap_off(pair_second(p)) - tap_off(pair_first(p))) << 16;
                    const uint32_t ah = (a_hi + (uint32_t)tap_off(pair_first(p))) | lbo;
                    const uint32_t al = (a_lo + (uint32_t)tap_off(pair_first(p))) | lbo;
                    // B: rows 16 B apart, 8-row groups 128 B apart, K halves NPR rows apart
                    const uint32_t b_lo32 = (b_base + (uint32_t)p * (Cfg::NPR * 2)) | ((uint32_t)Cfg::NPR << 16);
                    const uint64_t bdesc1 = b_hi_word | (uint64_t)b_lo32;
                    const uint64_t bdesc2 = b_hi_word | (uint64_t)(b_lo32 + (Cfg::N8 ? 8u : 0u));
#pragma unroll
                    for (int i = 0; i < BX; ++i) {
                        const uint32_t xo = (uint32_t)i * (TC_SYH * TC_SZH);
                        const uint32_t d = d_set + (uint32_t)i * Cfg::NPD;
                        // columns [0, 2N) = A_hi x [B_hi | B_lo'].  A_lo' x B_hi joins columns [N, 2N) (same weight
                        // 2^-11); for Cout = 8 MMA2 is A_lo' x [B_lo' | B_hi] into columns [16, 32) of its own
                        umma_f16(d, a_hi_word | (uint64_t)(ah + xo), bdesc1, idesc1, p != 0);
                        umma_f16(d + (Cfg::N8 ? 16 : N), a_hi_word | (uint64_t)(al + xo), bdesc2, idesc2,
                                 Cfg::N8 ? (uint32_t)(p != 0) : 1u);
                    }
                }
                umma_commit(&bar_empty[s]);
                umma_commit(&bar_acc_full[set]);
            }
        }
        __syncwarp();
    } else {
        // ---------------- workers: fp16 hi/lo images of the landed block, accumulator drain, epilogue
        const int wt = threadIdx.x - 64;
        const int q = warp & 3;                        // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;              // which BX/2 M tiles this thread drains
        const int row = q * 32 + lane;
        const size_t vol = (size_t)geo.X * geo.Y * geo.Z;
        float2 acc[Cfg::TILES_PER_HALF][N / 2];               // channel pairs: packed fp32 arithmetic (tc_ptx.cuh)
        constexpr float W2 = 1.f / 2048.f, W3 = W2 * W2;      // weights of the lo' cross terms

        auto drain = [&](int g, bool first) {
            const int set = g & 1, use_a = g >> 1;
            mbar_wait(&bar_acc_full[set], use_a & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)set * Cfg::SET_COLS +
                                (uint32_t)(half * Cfg::TILES_PER_HALF) * Cfg::NPD;
#pragma unroll
            for (int i = 0; i < Cfg::TILES_PER_HALF; ++i) {
#pragma unroll
                for (int n8 = 0; n8 < N / 8; ++n8) {
                    float a[8], b2[8];
                    tmem_ld8(t0 + i * Cfg::NPD + n8 * 8, a);
                    tmem_ld8(t0 + i * Cfg::NPD + N + n8 * 8, b2);
                    if (Cfg::N8) {
                        float c3[8], c4[8];
                        tmem_ld8(t0 + i * Cfg::NPD + 16, c3);
                        tmem_ld8(t0 + i * Cfg::NPD + 24, c4);
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 v = make_float2(fmaf(c3[2 * k], W3, fmaf(b2[2 * k] + c4[2 * k], W2, a[2 * k])),
                                                         fmaf(c3[2 * k + 1], W3, fmaf(b2[2 * k + 1] + c4[2 * k + 1], W2, a[2 * k + 1])));
                            acc[i][k] = first ? v : f2_add(acc[i][k], v);
                        }
                    } else {
                        tmem_ld_wait();
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float2 v = f2_fma(make_float2(b2[2 * k], b2[2 * k + 1]), f2_splat(W2), make_float2(a[2 * k], a[2 * k + 1]));
                            acc[i][n8 * 4 + k] = first ? v : f2_add(acc[i][n8 * 4 + k], v);
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_acc_empty[set]);
        };
        auto store_unit = [&](const TcUnit& un, float inv_scale, float s_out) {
            const int y = un.y0 + (row >> 3), z = un.z0 + (row & 7);
            float amax = 0.f;
            if (DST_SPLIT && wt == 0) geo.scale_dst[(size_t)un.tile * geo.slab_stride] = s_out;
            if (y < geo.Y) {
                float4* d_tile = dst + (size_t)un.tile * geo.dst_tile_stride4 + (size_t)geo.dst_c4off * vol;
#pragma unroll
                for (int i = 0; i < Cfg::TILES_PER_HALF; ++i) {
                    const int x = un.x0 + half * Cfg::TILES_PER_HALF + i;
                    if (x >= geo.X) break;
                    const size_t vox = ((size_t)x * geo.Y + y) * geo.Z + z;
                    if constexpr (!DST_SPLIT) {
#pragma unroll
                        for (int c4 = 0; c4 < N / 4; ++c4) {
                            float2 o[2];
#pragma unroll
                            for (int k = 0; k < 2; ++k) o[k] = block_epilogue(acc[i][c4 * 2 + k], inv_scale, alpha, &ep_s[0][0], N, c4 * 4 + 2 * k, amax);
                            d_tile[(size_t)c4 * vol + vox] = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
                        }
                    } else {
                        // plane 2 c8 = fp16 hi image of channels [8 c8, 8 c8 + 8), plane 2 c8 + 1 = lo' image
                        uint4* d4 = reinterpret_cast<uint4*>(d_tile);
#pragma unroll
                        for (int c8 = 0; c8 < N / 8; ++c8) {
                            uint32_t hi[4], lo[4];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const float2 o = block_epilogue(acc[i][c8 * 4 + k], inv_scale, alpha, &ep_s[0][0], N, c8 * 8 + 2 * k, amax);
                                split_pair2(o, s_out, hi[k], lo[k]);
                            }
                            d4[(size_t)(2 * c8) * vol + vox] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                            d4[(size_t)(2 * c8 + 1) * vol + vox] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                        }
                    }
                }
            }
            amax = warp_max(amax);
            if (lane == 0) amax_update(geo.amax_dst + (size_t)un.tile * geo.slab_stride, amax);
        };

        // flattened stage loop: stage g is converted, then stage g - 1 is drained while the MMAs of stage g run
        int c = 0, k = 0;                              // chunk / unit ordinal of stage g
        int pc = 0;                                    // chunk of stage g - 1
        TcUnit prev{}, cur{};
        float s_cur = 1.f, s_prev = 1.f, so_cur = 1.f, so_prev = 1.f;
        // max|x| of a unit's tile is fetched one unit ahead, so the load's latency is never in front of a conversion
        TcUnit nxt = tc_unit((int)blockIdx.x, geo, BX);
        float am_nxt = n_units > 0 ? geo.amax_src[(size_t)nxt.tile * geo.slab_stride] : 0.f;
        float sc_nxt = (SRC_SPLIT && n_units > 0) ? geo.scale_src[(size_t)nxt.tile * geo.slab_stride] : 1.f;
        for (int g = 0; g <= n_stages; ++g) {
            if (g < n_stages) {
                if (c == 0) {
                    cur = nxt;
                    s_cur = SRC_SPLIT ? sc_nxt : tc_operand_scale(am_nxt);
                    if constexpr (DST_SPLIT) so_cur = split_out_scale(am_nxt, geo.bound_p, geo.bound_q);
                    if (k + 1 < n_units) {
                        nxt = tc_unit((int)blockIdx.x + (k + 1) * (int)gridDim.x, geo, BX);
                        am_nxt = geo.amax_src[(size_t)nxt.tile * geo.slab_stride];
                        if constexpr (SRC_SPLIT) sc_nxt = geo.scale_src[(size_t)nxt.tile * geo.slab_stride];
                    }
                }
                const int s = g % STAGES, use = g / STAGES;
                if constexpr (!SRC_SPLIT) {
                mbar_wait(&bar_full[s], use & 1);
                uint4* p0 = reinterpret_cast<uint4*>(ring + (size_t)s * Cfg::STAGE);
                uint4* p1 = reinterpret_cast<uint4*>(ring + (size_t)s * Cfg::STAGE + Cfg::PLANE);
#pragma unroll 2
                for (int i = wt; i < Cfg::PLANE / 16; i += TC_WORKERS) {
                    const float4 v0 = *reinterpret_cast<const float4*>(p0 + i);      // channels 0-3 of the voxel
                    const float4 v1 = *reinterpret_cast<const float4*>(p1 + i);      // channels 4-7
                    const float x[8] = {v0.x * s_cur, v0.y * s_cur, v0.z * s_cur, v0.w * s_cur,
                                        v1.x * s_cur, v1.y * s_cur, v1.z * s_cur, v1.w * s_cur};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const __half2 h = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn((x[2 * j] - hf.x) * 2048.f, (x[2 * j + 1] - hf.y) * 2048.f);
                        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    p0[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    p1[i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_conv[s]);
                }
            }
            if (g > 0) {
                drain(g - 1, pc == 0);
                if (pc == cin8 - 1) store_unit(prev, geo.w_inv_scale / s_prev, so_prev);
            }
            if (c == 0) { prev = cur; s_prev = s_cur; so_prev = so_cur; }
            pc = c;
            if (++c == cin8) { c = 0; ++k; }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool tc_shape_ok(int cout) { return cout == 8 || cout == 16 || cout == 32 || cout == 64; }
static int tc_rows(int cout) { return cout == 8 ? 24 : 2 * cout; }     // TcCfg::NPR

size_t tc_weight_floats(int cin_pad, int cout) {
    if (!tc_shape_ok(cout)) return 0;
    return (size_t)((cin_pad + 7) / 8) * TC_PAIRS * tc_rows(cout) * 8;  // 32 bytes per row per pair
}

// keras kernel (kx,ky,kz,ci,co) -> fp16 image [ci/8][pair][k half][row][ci % 8], rows = co (hi) | cout + co (lo')
// (| 2 cout + co (hi again) when cout = 8); values scaled by a power of two so that max|w| lands in [2^13, 2^14).
float tc_pack_weights(const float* w, int cin, int cin_pad, int cout, float* dst) {
    const int npr = tc_rows(cout), c8n = (cin_pad + 7) / 8;
    std::memset(dst, 0, tc_weight_floats(cin_pad, cout) * sizeof(float));
    float wmax = 0.f;
    for (size_t i = 0; i < (size_t)27 * cin * cout; ++i) wmax = std::fmax(wmax, std::fabs(w[i]));
    int e = 0;
    if (wmax > 0.f) std::frexp(wmax, &e);                 // wmax = m * 2^e, m in [0.5, 1)
    const float scale = std::ldexp(1.f, 14 - e);          // wmax * scale in [2^13, 2^14)
    __half* img = reinterpret_cast<__half*>(dst);
    for (int c = 0; c < c8n; ++c)
        for (int p = 0; p < TC_PAIRS; ++p)
            for (int j = 0; j < 2; ++j) {
                if (p == 0 && j == 1) continue;          // zero column paired with tap 0
                const int tap = j == 0 ? pair_first(p) : pair_second(p);
                __half* blk = img + ((((size_t)c * TC_PAIRS + p) * 2 + j) * npr) * 8;
                for (int co = 0; co < cout; ++co)
                    for (int qd = 0; qd < 8; ++qd) {
                        const int ci = c * 8 + qd;
                        if (ci >= cin) continue;
                        const float v = w[((size_t)tap * cin + ci) * cout + co] * scale;
                        const __half h = __float2half_rn(v);
                        const __half l = __float2half_rn((v - __half2float(h)) * 2048.f);
                        blk[(size_t)co * 8 + qd] = h;
                        blk[(size_t)(cout + co) * 8 + qd] = l;
                        if (cout == 8) blk[(size_t)(2 * cout + co) * 8 + qd] = h;
                    }
            }
    return 1.f / scale;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

static int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

template <int N, int BX, int STAGES, bool SRC_SPLIT, bool DST_SPLIT>
static int launch_tc(const CUtensorMap& map, const ConvLayer& L, float alpha, float4* dst, int X, int Y, int Z,
                     size_t stride4, int dst_c4off, int tiles, const float* amax_src, float* amax_dst, cudaStream_t s) {
    using Cfg = TcCfg<N, BX, STAGES>;
    // per device / context attribute: set on every launch (cheap) so several GPUs in one process are correct
    CT_CUDA(cudaFuncSetAttribute(conv3_tc_kernel<N, BX, STAGES, SRC_SPLIT, DST_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    TcGeom g;
    g.cin8 = (L.cin_pad + 7) / 8; g.X = X; g.Y = Y; g.Z = Z;
    g.amax_src = amax_src; g.amax_dst = amax_dst; g.slab_stride = stride4 * 4; g.w_inv_scale = L.w_tc_inv_scale;
    g.nbx = cdiv(X, BX); g.nby = cdiv(Y, 16); g.nbz = Z / 8;
    g.units = g.nbx * g.nby * g.nbz * tiles;
    g.dst_c4off = dst_c4off; g.dst_tile_stride4 = stride4;
    g.scale_src = amax_src + SCALE_SLOT0; g.scale_dst = amax_dst + SCALE_SLOT0; g.bound_p = L.bound_p; g.bound_q = L.bound_q;
    const int sms = sm_count() - g_reserved_sms;
    const int grid = g.units < sms ? g.units : sms;
    conv3_tc_kernel<N, BX, STAGES, SRC_SPLIT, DST_SPLIT><<<grid, TC_THREADS, Cfg::SMEM, s>>>(map, L.w_tc, L.bias, L.scale, L.shift, alpha, dst, g);
    return 0;
}

int tc_sm_count() { return sm_count(); }

int tc_make_map(CUtensorMap* map, float* base, int X, int Y, int Z, int c4, int tiles, size_t slab_stride, int bx) {
    EncodeTiledFn enc = encode_fn();
    CT_REQUIRE(enc, "unet: cuTensorMapEncodeTiled is unavailable in this driver");
    const cuuint64_t dims[5] = {(cuuint64_t)Z * 4, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)c4, (cuuint64_t)tiles};
    const cuuint64_t strides[4] = {(cuuint64_t)Z * 16, (cuuint64_t)Y * Z * 16, (cuuint64_t)X * Y * Z * 16,
                                   (cuuint64_t)slab_stride * 4};
    const cuuint32_t box[5] = {TC_SZH * 4, TC_SYH, (cuuint32_t)bx + 2, 2, 1};      // two 4-channel planes per stage
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CT_REQUIRE(r == CUDA_SUCCESS, "unet: cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

// same tensor (dims z*4 | y | x | 4-channel plane | tile) with a caller-chosen box
int tc_make_map_box(CUtensorMap* map, float* base, int X, int Y, int Z, int c4, int tiles, size_t slab_stride, const unsigned box_in[5]) {
    EncodeTiledFn enc = encode_fn();
    CT_REQUIRE(enc, "unet: cuTensorMapEncodeTiled is unavailable in this driver");
    const cuuint64_t dims[5] = {(cuuint64_t)Z * 4, (cuuint64_t)Y, (cuuint64_t)X, (cuuint64_t)c4, (cuuint64_t)tiles};
    const cuuint64_t strides[4] = {(cuuint64_t)Z * 16, (cuuint64_t)Y * Z * 16, (cuuint64_t)X * Y * Z * 16,
                                   (cuuint64_t)slab_stride * 4};
    const cuuint32_t box[5] = {box_in[0], box_in[1], box_in[2], box_in[3], box_in[4]};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, base, dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CT_REQUIRE(r == CUDA_SUCCESS, "unet: cuTensorMapEncodeTiled failed with code %d", (int)r);
    return 0;
}

int launch_conv_tc(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s, int fmt) {
    const ConvLayer& L = net->layers[op.layer];
    const int X = op.sx, Y = op.sy, Z = op.sz;
    if (!L.w_tc || !tc_shape_ok(L.cout) || Z % 8 != 0) return 2;
    CT_REQUIRE(op.src_c == L.cin_pad, "conv: source buffer has %d channels, layer expects %d", op.src_c, L.cin_pad);
    CT_REQUIRE(slab_stride % 4 == 0 && op.src_off % 4 == 0 && op.dst_off % 4 == 0, "conv: misaligned slab");
    float4* dst = reinterpret_cast<float4*>(slab0 + op.dst_off);
    CUtensorMap map;
    ProfScope prof(PROF_CONV, s);
    int rc;
    const size_t st4 = slab_stride / 4;
    const int co4 = op.dst_coff / 4, c4 = L.cin_pad / 4;
    float* src = slab0 + op.src_off;
    const float* am_s = slab0 + op.src_slot;
    float* am_d = slab0 + op.dst_slot;
    if (fmt != 0) {
        // split-fp16 buffers: only the Cout = 64 blocks run on this kernel inside the network (unet_tcx.cu takes the rest)
        if (L.cout != 64 || ((fmt & FMT_SRC_SPLIT) && L.cin_pad % 8 != 0)) return 2;
        CT_REQUIRE(!(fmt & FMT_DST_SPLIT) || op.dst_coff % 8 == 0, "conv: split destination at channel offset %d", op.dst_coff);
        if (tc_make_map(&map, src, X, Y, Z, c4, tiles, slab_stride, 2)) return 1;
        if (fmt == FMT_SRC_SPLIT) rc = launch_tc<64, 2, 2, true, false>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
        else if (fmt == FMT_DST_SPLIT) rc = launch_tc<64, 2, 2, false, true>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
        else rc = launch_tc<64, 2, 2, true, true>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
    } else if (L.cout == 8) {
        if (tc_make_map(&map, src, X, Y, Z, c4, tiles, slab_stride, 8)) return 1;
        rc = launch_tc<8, 8, 3, false, false>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
    } else if (L.cout == 16) {
        if (tc_make_map(&map, src, X, Y, Z, c4, tiles, slab_stride, 8)) return 1;
        rc = launch_tc<16, 8, 3, false, false>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
    } else if (L.cout == 32) {
        if (tc_make_map(&map, src, X, Y, Z, c4, tiles, slab_stride, 4)) return 1;
        rc = launch_tc<32, 4, 2, false, false>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
    } else {
        if (tc_make_map(&map, src, X, Y, Z, c4, tiles, slab_stride, 2)) return 1;
        rc = launch_tc<64, 2, 2, false, false>(map, L, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, s);
    }
    if (rc) return 1;
    CT_LAUNCHED("conv3_tc_kernel");
    return 0;
}

}  // namespace ct
