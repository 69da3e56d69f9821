// tcgen05 convolution over a NEAREST-UPSAMPLED source without materialising the up-sampling ("phase" kernel).
//
// Reference: the decoder of unet3d.py:84-98 -- `UpSampling3D((2,2,1))` (:96) -> `concatenate([up, skip])` (:97) ->
// `_conv3d_leakyrelu_bn` (:101-120).  The 3x3x3 convolution of the up-sampled half of the concatenation reads every
// low-resolution voxel several times: for an output voxel (2X+px, 2Y+py, z) the three x-taps fall on only TWO
// low-resolution planes,
//     px = 0:  taps -1, 0, +1  ->  planes X-1, X, X        weights  W[0] | W[1] + W[2]
//     px = 1:  taps -1, 0, +1  ->  planes X, X, X+1        weights  W[0] + W[1] | W[2]
// (the same in y).  So the up-sampled half is a convolution on the LOW-resolution grid with 2 x 2 x 3 effective taps
// per output phase (px, py) instead of 27 per full-resolution voxel: 12/27 of the multiply-adds, a quarter of the
// operand voxels, and no up-sampled tensor in HBM at all.  Zero padding is unchanged: a tap that leaves the tile reads
// up-sampled coordinate -1 or 2X, which is low-resolution plane -1 or X -- out of range for the TMA box, zero filled.
//
// GEMM layout (same operand format as unet_tcx.cu: fp16 hi / lo' images, M = 16 y x 8 z low-resolution voxels of one
// plane, K = 16 = two taps x 8 channels).  One accumulator set = (input plane j, output y-phase py); its 6 (ay, dz)
// taps are 3 K steps.  The x direction is stacked in N like in the x-stacked kernel, with the four (ax, px) pairs
// (-1,0) (0,0) (0,1) (+1,1) taking the place of the three x-taps:
//     B rows per K half:  [hi k0..k3 | lo' k0..k3]  (k = (ax,px) pair, Cout rows each)      8 Cout rows
//     MMA1 = A_hi  x rows [0, 8 Cout)  -> columns [0, 4 Cout) = hi.hi, [4 Cout, 8 Cout) = hi.lo'
//     MMA2 = A_lo' x rows [0, 4 Cout)  -> accumulated onto columns [4 Cout, 8 Cout)
// Input plane j feeds output plane j - 1 - ax, phase px: the drain threads (one low-resolution voxel row each) add the
// pairs into acc[plane][px][py][channel] -- all four full-resolution outputs of a low-resolution voxel live in the
// same thread.  The kernel writes the PARTIAL pre-activation sums (true units) of the block into the destination
// buffer; the x-stacked kernel then convolves the skip half of the concatenation, adds the partial sums in its
// epilogue (TxGeom::add_partial) and applies bias -> activation -> BatchNorm.  N = 8 Cout per MMA1 (64 / 128 / 256
// columns) is what makes this pass math-bound instead of operand-read-bound: DESIGN.md section 3.2.
#include "unet_common.cuh"
#include "tc_ptx.cuh"
#include <cmath>
#include <cstring>

namespace ct {

int tc_sm_count();
int tc_make_map(CUtensorMap* map, float* base, int X, int Y, int Z, int c4, int tiles, size_t slab_stride, int bx);

constexpr int TU_SYH = 18, TU_SZH = 10;
constexpr int TU_PLANE_VOX = TU_SYH * TU_SZH;
constexpr int TU_KSTEPS = 3;                     // 6 (ay, dz) taps of one y-phase
constexpr int TU_CONV_WARPS = 2;

// tap t = 0..5 of phase py: ay = (py == 0 ? -1 : 0) + t / 3, dz = t % 3 - 1; offset inside the haloed plane
__host__ __device__ constexpr int tu_off(int py, int t) { return (py + t / 3) * TU_SZH + t % 3; }

template <int N, int BX, int STAGES>
struct TuCfg {
    static constexpr int NPR = 8 * N;                                  // B rows per K half
    static constexpr int N1 = 8 * N, N2 = 4 * N;
    static constexpr int NPD = 8 * N;                                  // accumulator columns of one set
    static constexpr int NSETS = (4 * NPD <= 512) ? 4 : 2;
    static constexpr int SXH = BX + 2;
    static constexpr int PLANE = SXH * TU_PLANE_VOX * 16;
    static constexpr int B_BYTES = 2 * TU_KSTEPS * 2 * NPR * 16;       // [py][K step][K half][row][8 fp16]
    static constexpr int STAGE = 2 * PLANE + B_BYTES;
    static constexpr int TMEM_COLS = NSETS * NPD <= 256 ? 256 : 512;
    static constexpr int SMEM = STAGES * STAGE + 1024;
    static constexpr int DRAIN_WARPS = 8;
    static constexpr int THREADS = 64 + 32 * (TU_CONV_WARPS + DRAIN_WARPS);
    static constexpr int CH = N / (DRAIN_WARPS / 4);                   // 4 / 8 / 16 output channels per drain thread
    // acc[BX][2][2][CH] = 128 registers per drain thread in every configuration (BX = 8 / 4 / 2): the drain warps take
    // registers from warps 0-3 (setmaxnreg inside the launch-time pool of 384 x 168: 128 x 56 + 256 x 224 = 64512)
#ifdef CT_NO_REBALANCE
    static constexpr bool REBALANCE = false;                           // debug build: no setmaxnreg (spills instead)
#else
    static constexpr bool REBALANCE = true;
#endif
    static_assert(N1 % 16 == 0 && N1 <= 256 && N2 % 16 == 0, "UMMA N out of range for M = 128");
    static_assert(NSETS * NPD <= 512, "accumulators exceed tensor memory");
    static_assert(PLANE % 128 == 0 && B_BYTES % 128 == 0, "stage parts must stay 128-byte aligned");
    static_assert(SMEM <= 232448, "shared memory ring too large");
};

struct TuGeom {
    int cin8, X, Y, Z, nbx, nby, nbz, units;     // low-resolution source extents
    int dst_c4off;
    size_t dst_tile_stride4;
    const float* amax_src;
    size_t slab_stride;
    float w_inv_scale;
    const float* scale_src;                      // split-fp16 source: header scale slot (unet_common.cuh)
    // The partial sums are written in the ACCUMULATOR units of the kernel that adds the skip half (unet_tcx.cu): true sum
    // x skip operand scale (per tile: its header slot, an operand scale or -- fp32 buffers -- a max|x| bound) x
    // post_w_scale (the skip weights' power-of-two scale).  That kernel then loads them straight into its accumulator
    // registers: no arithmetic, so nothing waits for the loads at the head of a unit.
    const float* post_slot;
    int post_slot_is_scale;
    float post_w_scale;
    // 1: "P8" layout for the plane-walk kernel (unet_tcz.cu): of a thread's channel quadruple 4 k .. 4 k + 3 the first
    // pair goes to plane 2 (k / 2), the second pair to plane 2 (k / 2) + 1, both at byte 8 (k % 2) of the voxel's 16 --
    // exactly the bytes the consumer's thread that owns those channels overwrites with their fp16 hi / lo' halves
    int post_p8;
};

struct TuUnit { int x0, y0, z0, tile; };
__device__ __forceinline__ TuUnit tu_unit(int u, const TuGeom& g, int bx) {
    TuUnit r;
    r.x0 = (u % g.nbx) * bx; u /= g.nbx;
    r.y0 = (u % g.nby) * 16; u /= g.nby;
    r.z0 = (u % g.nbz) * 8;
    r.tile = u / g.nbz;
    return r;
}

template <int CH>
__device__ __forceinline__ void tu_ld_issue(uint32_t taddr, uint32_t (&r)[CH]) {
    if constexpr (CH == 4) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                     : "r"(taddr));
    } else if constexpr (CH == 16) {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                       "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                     : "r"(taddr));
    } else {
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
    }
}

// Persistent CTA.  Work unit = BX x 16 x 8 LOW-resolution voxels (= 2 BX x 32 x 8 output voxels) of one tile.
// Stage g = one 8-channel chunk of one unit; accumulator step a = (g * (BX+2) + j) * 2 + py.
// SRC_SPLIT: the low-resolution source already holds the fp16 hi / lo' operand images (no conversion warps).
template <int N, int BX, int STAGES, bool SRC_SPLIT>
__global__ void __launch_bounds__(TuCfg<N, BX, STAGES>::THREADS, 1)
conv3_tcu_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wpack, float4* __restrict__ dst,
                 const TuGeom geo) {
    using Cfg = TuCfg<N, BX, STAGES>;
    constexpr int SXH = Cfg::SXH, CH = Cfg::CH, NSETS = Cfg::NSETS;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[STAGES], bar_conv[STAGES], bar_empty[STAGES], bar_acc_full[NSETS], bar_acc_empty[NSETS];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* ring = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int cin8 = geo.cin8;
    const int n_units = ((int)blockIdx.x < geo.units) ? (geo.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_stages = n_units * cin8;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_conv[s], TU_CONV_WARPS);
            mbar_init(&bar_empty[s], 1);
        }
#pragma unroll
        for (int a = 0; a < NSETS; ++a) {
            mbar_init(&bar_acc_full[a], 1);
            mbar_init(&bar_acc_empty[a], Cfg::DRAIN_WARPS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 2 + TU_CONV_WARPS) {
    if constexpr (Cfg::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ---------------- TMA producer
        if (elect_one()) {
            int g = 0;
            for (int k = 0; k < n_units; ++k) {
                const TuUnit un = tu_unit((int)blockIdx.x + k * (int)gridDim.x, geo, BX);
                for (int c = 0; c < cin8; ++c, ++g) {
                    const int s = g % STAGES, use = g / STAGES;
                    if (use > 0) mbar_wait(&bar_empty[s], (use - 1) & 1);
                    uint8_t* st = ring + (size_t)s * Cfg::STAGE;
                    mbar_expect_tx(&bar_full[s], 2 * Cfg::PLANE + Cfg::B_BYTES);
                    tma_load_5d(st, &tmap, &bar_full[s], (un.z0 - 1) * 4, un.y0 - 1, un.x0 - 1, 2 * c, un.tile);
                    bulk_load(st + 2 * Cfg::PLANE, wpack + (size_t)c * (Cfg::B_BYTES / 4), Cfg::B_BYTES, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ---------------- MMA issuer
        if (elect_one()) {
            constexpr uint32_t idesc1 = (1u << 4) | ((uint32_t)(Cfg::N1 >> 3) << 17) | (8u << 24);
            constexpr uint32_t idesc2 = (1u << 4) | ((uint32_t)(Cfg::N2 >> 3) << 17) | (8u << 24);
            constexpr uint64_t a_hi_word = (uint64_t)((uint32_t)TU_SZH | (1u << 14)) << 32;   // SBO: y rows 160 B apart
            constexpr uint64_t b_hi_word = (uint64_t)(8u | (1u << 14)) << 32;                 // SBO: 8-row groups 128 B apart
            const uint32_t ring16 = smem_u32(ring) >> 4;
            int a = 0;
            for (int g = 0; g < n_stages; ++g) {
                const int s = g % STAGES, use = g / STAGES;
                mbar_wait(SRC_SPLIT ? &bar_full[s] : &bar_conv[s], use & 1);
                if constexpr (SRC_SPLIT) tc_fence_after();
                const uint32_t a_hi = ring16 + (uint32_t)s * (Cfg::STAGE / 16), a_lo = a_hi + Cfg::PLANE / 16;
                const uint32_t b_base = a_hi + 2 * (Cfg::PLANE / 16);
#pragma unroll 1
                for (int j = 0; j < SXH; ++j) {
                    const uint32_t pl = (uint32_t)j * TU_PLANE_VOX;
#pragma unroll
                    for (int py = 0; py < 2; ++py, ++a) {
                        const int set = a % NSETS, use_a = a / NSETS;
                        if (use_a > 0) mbar_wait(&bar_acc_empty[set], (use_a - 1) & 1);
                        tc_fence_after();
                        const uint32_t d = tmem_base + (uint32_t)set * Cfg::NPD;
#pragma unroll
                        for (int p = 0; p < TU_KSTEPS; ++p) {
                            const uint32_t first = (uint32_t)tu_off(py, 2 * p);
                            const uint32_t lbo = (uint32_t)(tu_off(py, 2 * p + 1) - tu_off(py, 2 * p)) << 16;
                            const uint32_t ah = (a_hi + pl + first) | lbo;
                            const uint32_t al = (a_lo + pl + first) | lbo;
                            const uint32_t b32 = (b_base + (uint32_t)(py * TU_KSTEPS + p) * (Cfg::NPR * 2)) | ((uint32_t)Cfg::NPR << 16);
                            umma_f16(d, a_hi_word | (uint64_t)ah, b_hi_word | (uint64_t)b32, idesc1, p != 0);
                            umma_f16(d + Cfg::N2, a_hi_word | (uint64_t)al, b_hi_word | (uint64_t)b32, idesc2, 1u);
                        }
                        umma_commit(&bar_acc_full[set]);
                    }
                }
                umma_commit(&bar_empty[s]);
            }
        }
        __syncwarp();
    } else if constexpr (!SRC_SPLIT) {
        // ---------------- converters: fp32 -> fp16 hi / lo' images of every landed stage, in place
        const int ct = threadIdx.x - 64;
        int g = 0;
        for (int k = 0; k < n_units; ++k) {
            const TuUnit un = tu_unit((int)blockIdx.x + k * (int)gridDim.x, geo, BX);
            const float sc = tc_operand_scale(geo.amax_src[(size_t)un.tile * geo.slab_stride]);
            for (int c = 0; c < cin8; ++c, ++g) {
                const int s = g % STAGES, use = g / STAGES;
                mbar_wait(&bar_full[s], use & 1);
                uint4* p0 = reinterpret_cast<uint4*>(ring + (size_t)s * Cfg::STAGE);
                uint4* p1 = p0 + Cfg::PLANE / 16;
#pragma unroll 2
                for (int i = ct; i < Cfg::PLANE / 16; i += 32 * TU_CONV_WARPS) {
                    const float4 v0 = *reinterpret_cast<const float4*>(p0 + i);
                    const float4 v1 = *reinterpret_cast<const float4*>(p1 + i);
                    const float x[8] = {v0.x * sc, v0.y * sc, v0.z * sc, v0.w * sc, v1.x * sc, v1.y * sc, v1.z * sc, v1.w * sc};
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int q2 = 0; q2 < 4; ++q2) {
                        const __half2 h = __floats2half2_rn(x[2 * q2], x[2 * q2 + 1]);
                        const float2 hf = __half22float2(h);
                        const __half2 l = __floats2half2_rn((x[2 * q2] - hf.x) * 2048.f, (x[2 * q2 + 1] - hf.y) * 2048.f);
                        hi[q2] = *reinterpret_cast<const uint32_t*>(&h);
                        lo[q2] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    p0[i] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    p1[i] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
                fence_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_conv[s]);
            }
        }
    }
    } else {
        // ---------------- drain warps: tensor memory -> registers (phase add), partial sums -> destination
        if constexpr (Cfg::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int q = warp & 3;
        const int part = (warp - 2 - TU_CONV_WARPS) >> 2;
        const int ch0 = part * CH;
        const int row = q * 32 + lane;
        const int DX = 2 * geo.X, DY = 2 * geo.Y;
        const size_t vol = (size_t)DX * DY * geo.Z;
        float2 acc[BX][2][2][CH / 2];                              // [plane][px][py][channel pair] (packed fp32, tc_ptx.cuh)
        constexpr float W2 = 1.f / 2048.f;
        int a = 0;
        TuUnit un_next = tu_unit((int)blockIdx.x, geo, BX);
        const float* s_slot = SRC_SPLIT ? geo.scale_src : geo.amax_src;      // split source: the scale itself
        float am_next = n_units > 0 ? s_slot[(size_t)un_next.tile * geo.slab_stride] : 0.f;
        for (int k = 0; k < n_units; ++k) {
            const TuUnit un = un_next;
            const float post = geo.post_slot[(size_t)un.tile * geo.slab_stride];
            const float inv_scale = geo.w_inv_scale / (SRC_SPLIT ? am_next : tc_operand_scale(am_next)) *
                                    (geo.post_slot_is_scale ? post : tc_operand_scale(post)) * geo.post_w_scale;
            if (k + 1 < n_units) {
                un_next = tu_unit((int)blockIdx.x + (k + 1) * (int)gridDim.x, geo, BX);
                am_next = s_slot[(size_t)un_next.tile * geo.slab_stride];
            }
            const int yl = un.y0 + (row >> 3), z = un.z0 + (row & 7);
            float4* d_base = dst + (size_t)un.tile * geo.dst_tile_stride4 + (size_t)geo.dst_c4off * vol;
            float4* d_tile = d_base + (size_t)(ch0 / 4) * vol;
            auto store_plane = [&](int i) {
                const int xl = un.x0 + i;
                if (yl >= geo.Y || xl >= geo.X) return;
#pragma unroll
                for (int px = 0; px < 2; ++px)
#pragma unroll
                    for (int py = 0; py < 2; ++py) {
                        const size_t vox = ((size_t)(2 * xl + px) * DY + (2 * yl + py)) * geo.Z + z;
#pragma unroll
                        for (int c4 = 0; c4 < CH / 4; ++c4) {
                            const float2 lo2 = f2_mul(acc[i][px][py][c4 * 2], f2_splat(inv_scale));
                            const float2 hi2 = f2_mul(acc[i][px][py][c4 * 2 + 1], f2_splat(inv_scale));
                            if (geo.post_p8) {                      // uniform over the grid
                                const int k4 = ch0 / 4 + c4;
                                float2* b2 = reinterpret_cast<float2*>(d_base + ((size_t)(k4 & ~1) * vol + vox)) + (k4 & 1);
                                b2[0] = lo2;
                                b2[vol * 2] = hi2;
                            } else {
                                d_tile[(size_t)c4 * vol + vox] = make_float4(lo2.x, lo2.y, hi2.x, hi2.y);
                            }
                        }
                    }
            };
#pragma unroll
            for (int i = 0; i < BX; ++i)
#pragma unroll
                for (int p = 0; p < 4; ++p)
#pragma unroll
                    for (int ch = 0; ch < CH / 2; ++ch) acc[i][p >> 1][p & 1][ch] = make_float2(0.f, 0.f);
#pragma unroll 1
            for (int c = 0; c < cin8; ++c) {
                const bool last = (c == cin8 - 1);
#pragma unroll
                for (int j = 0; j < SXH; ++j) {
#pragma unroll
                    for (int py = 0; py < 2; ++py, ++a) {
                        const int set = a % NSETS, use_a = a / NSETS;
                        mbar_wait(&bar_acc_full[set], use_a & 1);
                        tc_fence_after();
                        const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)set * Cfg::NPD + (uint32_t)ch0;
                        // pair kk = (ax, px): (-1,0) (0,0) (0,1) (+1,1) -> output plane j - 1 - ax.  The loads of LDG
                        // pairs are issued in front of one wait (a wait per pair left the tensor pipe idle behind the
                        // drain: 80 tensor-memory round trips per stage); LDG = 4 while 8 CH registers per pair fit.
                        constexpr int LDG = (CH <= 8) ? 4 : 2;
#pragma unroll
                        for (int k0 = 0; k0 < 4; k0 += LDG) {
                            uint32_t v[LDG][2][CH];
#pragma unroll
                            for (int kq = 0; kq < LDG; ++kq) {
                                const int kk = k0 + kq;
                                const int ax = (kk == 0) ? -1 : (kk == 3 ? 1 : 0);
                                const int i = j - 1 - ax;
                                if (i < 0 || i >= BX) continue;
                                tu_ld_issue<CH>(t0 + kk * N, v[kq][0]);
                                tu_ld_issue<CH>(t0 + 4 * N + kk * N, v[kq][1]);
                            }
                            tmem_ld_wait();
#pragma unroll
                            for (int kq = 0; kq < LDG; ++kq) {
                                const int kk = k0 + kq;
                                const int ax = (kk == 0) ? -1 : (kk == 3 ? 1 : 0), px = kk >> 1;
                                const int i = j - 1 - ax;
                                if (i < 0 || i >= BX) continue;
#pragma unroll
                                for (int ch = 0; ch < CH / 2; ++ch)
                                    acc[i][px][py][ch] = f2_add(acc[i][px][py][ch],
                                        f2_fma(make_float2(__uint_as_float(v[kq][1][2 * ch]), __uint_as_float(v[kq][1][2 * ch + 1])), f2_splat(W2),
                                               make_float2(__uint_as_float(v[kq][0][2 * ch]), __uint_as_float(v[kq][0][2 * ch + 1]))));
                            }
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_acc_empty[set]);
                    }
                    if (last && j >= 2) store_plane(j - 2);
                }
            }
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
size_t tcu_weight_floats(int c_up, int cout) {
    if ((cout != 8 && cout != 16 && cout != 32) || c_up % 8 != 0 || c_up <= 0) return 0;
    return (size_t)(c_up / 8) * 2 * TU_KSTEPS * 2 * (8 * cout) * 4;      // 16 bytes per row per K half
}

// keras kernel (kx,ky,kz,ci,co), input channels [0, c_up) = the up-sampled half of the concatenation ->
// fp16 image [ci/8][py][K step][K half][row][ci % 8], rows k*cout + co (hi) | 4 cout + k*cout + co (lo'),
// k = (ax,px) pair.  Effective low-resolution weights: sum of the full-resolution taps that fall on the same plane.
// Returns 1 / (power-of-two scale).
float tcu_pack_weights(const float* w, int cin, int c_up, int cout, float* dst) {
    const int npr = 8 * cout, c8n = c_up / 8;
    std::memset(dst, 0, tcu_weight_floats(c_up, cout) * sizeof(float));
    // x-taps of pair k (kx indices into the keras kernel), y-taps of (py, ay index 0/1)
    static const int kx_of[4][2] = {{0, -1}, {1, 2}, {0, 1}, {2, -1}};
    static const int ky_of[2][2][2] = {{{0, -1}, {1, 2}}, {{0, 1}, {2, -1}}};
    auto eff = [&](int k, int py, int t, int ci, int co) {
        const int ayi = t / 3, kz = t % 3;
        double s = 0.0;
        for (int a = 0; a < 2; ++a) {
            const int kx = kx_of[k][a];
            if (kx < 0) continue;
            for (int b = 0; b < 2; ++b) {
                const int ky = ky_of[py][ayi][b];
                if (ky < 0) continue;
                s += (double)w[((size_t)((kx * 3 + ky) * 3 + kz) * cin + ci) * cout + co];
            }
        }
        return (float)s;
    };
    float wmax = 0.f;
    for (int k = 0; k < 4; ++k)
        for (int py = 0; py < 2; ++py)
            for (int t = 0; t < 6; ++t)
                for (int ci = 0; ci < c_up; ++ci)
                    for (int co = 0; co < cout; ++co) wmax = std::fmax(wmax, std::fabs(eff(k, py, t, ci, co)));
    int e = 0;
    if (wmax > 0.f) std::frexp(wmax, &e);
    const float scale = std::ldexp(1.f, 14 - e);
    __half* img = reinterpret_cast<__half*>(dst);
    for (int c = 0; c < c8n; ++c)
        for (int py = 0; py < 2; ++py)
            for (int p = 0; p < TU_KSTEPS; ++p)
                for (int j = 0; j < 2; ++j) {
                    const int t = 2 * p + j;
                    __half* blk = img + (((((size_t)c * 2 + py) * TU_KSTEPS + p) * 2 + j) * npr) * 8;
                    for (int k = 0; k < 4; ++k)
                        for (int co = 0; co < cout; ++co)
                            for (int qd = 0; qd < 8; ++qd) {
                                const float v = eff(k, py, t, c * 8 + qd, co) * scale;
                                const __half h = __float2half_rn(v);
                                const __half l = __float2half_rn((v - __half2float(h)) * 2048.f);
                                blk[(size_t)(k * cout + co) * 8 + qd] = h;
                                blk[(size_t)(4 * cout + k * cout + co) * 8 + qd] = l;
                            }
                }
    return 1.f / scale;
}

template <int N, int BX, int STAGES, bool SRC_SPLIT>
static int launch_tcu(const CUtensorMap& map, const float* wpack, float inv_scale, float4* dst, int X, int Y, int Z,
                      int cin8, size_t stride4, int dst_c4off, int tiles, const float* amax_src, const float* post_slot,
                      int post_is_scale, float post_w_scale, int p8, cudaStream_t s) {
    using Cfg = TuCfg<N, BX, STAGES>;
    CT_CUDA(cudaFuncSetAttribute(conv3_tcu_kernel<N, BX, STAGES, SRC_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    TuGeom g;
    g.cin8 = cin8; g.X = X; g.Y = Y; g.Z = Z;
    g.amax_src = amax_src; g.slab_stride = stride4 * 4; g.w_inv_scale = inv_scale;
    g.scale_src = amax_src + SCALE_SLOT0;
    g.post_slot = post_slot; g.post_slot_is_scale = post_is_scale; g.post_w_scale = post_w_scale; g.post_p8 = p8;
    g.nbx = cdiv(X, BX); g.nby = cdiv(Y, 16); g.nbz = Z / 8;
    g.units = g.nbx * g.nby * g.nbz * tiles;
    g.dst_c4off = dst_c4off; g.dst_tile_stride4 = stride4;
    const int sms = tc_sm_count() - g_reserved_sms;
    const int grid = g.units < sms ? g.units : sms;
    conv3_tcu_kernel<N, BX, STAGES, SRC_SPLIT><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(map, wpack, dst, g);
    return 0;
}

// Partial sums of conv block `L` over its up-sampled input half: low-resolution source `up` (c_up channels at
// X x Y x Z, header slot up_slot; fp32 or, with src_split, split-fp16) -> fp32 pre-activation sums in the block's
// destination buffer (2X x 2Y x Z).  Returns 2 when the shape is not handled.
int launch_conv_tcu(const CtUNet* net, const ConvLayer& L, float* slab0, size_t slab_stride, int tiles, size_t up_off,
                    int up_slot, int X, int Y, int Z, size_t dst_off, int dst_coff, cudaStream_t s, bool src_split,
                    int skip_slot, bool p8) {
    if (!L.w_tcu || Z % 8 != 0) return 2;
    CT_REQUIRE(slab_stride % 4 == 0 && up_off % 4 == 0 && dst_off % 4 == 0, "conv: misaligned slab");
    CUtensorMap map;
    ProfScope prof(PROF_CONV, s);
    if (tc_make_map(&map, slab0 + up_off, X, Y, Z, L.c_up / 4, tiles, slab_stride, L.cout == 8 ? 8 : (L.cout == 16 ? 4 : 2))) return 1;
    float4* dst = reinterpret_cast<float4*>(slab0 + dst_off);
    const float* am = slab0 + up_slot;
    const size_t st4 = slab_stride / 4;
    const int cin8 = L.c_up / 8, co4 = dst_coff / 4;
    // the skip half's operand scale: its scale slot (split buffers) or max|x| slot (fp32 buffers) + the weights' scale
    const float* post = slab0 + skip_slot + (src_split ? SCALE_SLOT0 : 0);
    const int pis = src_split ? 1 : 0;
    // the partial sums are left in the accumulator units (and, p8, the layout) of the kernel that adds the skip half
    const float pw = 1.f / (p8 ? L.w_tcz_skip_inv_scale : L.w_tcx_skip_inv_scale);
    const int p8i = p8 ? 1 : 0;
    int rc;
    if (src_split) {
        if (L.cout == 8) rc = launch_tcu<8, 8, 3, true>(map, L.w_tcu, L.w_tcu_inv_scale, dst, X, Y, Z, cin8, st4, co4, tiles, am, post, pis, pw, p8i, s);
        else if (L.cout == 16) rc = launch_tcu<16, 4, 3, true>(map, L.w_tcu, L.w_tcu_inv_scale, dst, X, Y, Z, cin8, st4, co4, tiles, am, post, pis, pw, p8i, s);
        else rc = launch_tcu<32, 2, 3, true>(map, L.w_tcu, L.w_tcu_inv_scale, dst, X, Y, Z, cin8, st4, co4, tiles, am, post, pis, pw, p8i, s);
    } else {
        if (L.cout == 8) rc = launch_tcu<8, 8, 3, false>(map, L.w_tcu, L.w_tcu_inv_scale, dst, X, Y, Z, cin8, st4, co4, tiles, am, post, pis, pw, p8i, s);
        else if (L.cout == 16) rc = launch_tcu<16, 4, 3, false>(map, L.w_tcu, L.w_tcu_inv_scale, dst, X, Y, Z, cin8, st4, co4, tiles, am, post, pis, pw, p8i, s);
        else rc = launch_tcu<32, 2, 3, false>(map, L.w_tcu, L.w_tcu_inv_scale, dst, X, Y, Z, cin8, st4, co4, tiles, am, post, pis, pw, p8i, s);
    }
    if (rc) return 1;
    CT_LAUNCHED("conv3_tcu_kernel");
    (void)net;
    return 0;
}

}  // namespace ct
