// tcgen05 implicit-GEMM 3x3x3 convolution, "x-stacked" variant for layers with few output channels (Cout 8 / 16).
//
// Same operator, operand precision and memory layout as unet_tc.cu (Conv3D 3x3x3 'same' + bias -> LeakyReLU/ReLU ->
// BatchNorm(eval), unet3d.py:117-119 / :139-140; fp16 hi/lo operand images, fp32 accumulation drained into registers),
// but a different GEMM decomposition.  unet_tc.cu issues one MMA per (tap, output plane) with N = 2 Cout: with
// Cout <= 32 every MMA is bound by the 4 KB shared-memory read of its A tile (128 voxels x 16 K), not by math.  Here
// the three x-taps of the kernel are stacked in the N dimension instead:
//
//     D_j[voxel (y,z), (dx, co)] = sum over (dy, dz, ci) of  in[plane j][y+dy, z+dz, ci] * W[dx, dy, dz, ci, co]
//     out[plane i] = D_i[dx = 0] + D_(i+1)[dx = 1] + D_(i+2)[dx = 2]          (input planes j are haloed by one)
//
// so one A tile (one of 9 (dy,dz) taps of one INPUT plane) feeds N = 3 x 2 Cout accumulator columns: a third of the
// shared-memory operand reads per output voxel, at the price of (BX+2)/BX more MMAs (halo planes) and of the shift-add
// over dx, which the worker warps do in registers while they drain tensor memory (each thread owns one voxel row of
// every plane, so the three contributions of an output voxel meet in the same thread).
//
// B rows per K half: [hi dx0 | hi dx1 | hi dx2 | lo' dx0 | lo' dx1 | lo' dx2]:
//     MMA1 = A_hi  x rows [0, 6N)            -> columns [0, 3N) = hi.hi, [3N, 6N) = hi.lo'
//     MMA2 = A_lo' x rows [0, 3N)            -> accumulated onto columns [3N, 6N)          (same weight 2^-11)
//     Cout = 8 (3N = 24 is not a legal MMA width): MMA2 = A_lo' x rows [0, 6N) at column 3N -- lo'.hi still lands on
//     columns [3N, 6N); columns [6N, 9N) receive lo'.lo' (weight 2^-22), which is never read, like in the wider layers.
// K = 16 per MMA = two (dy,dz) taps x 8 input channels; 9 taps -> 5 K steps (tap 7 appears twice, once with zero
// weights).  One accumulator set = one input plane x one 8-channel chunk = 5 MMAs per column group, drained into fp32
// registers (round-to-nearest) while the next plane's MMAs run into the other set.
#include "unet_common.cuh"
#include "tc_ptx.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>

namespace ct {

int tc_sm_count();
int tc_make_map(CUtensorMap* map, float* base, int X, int Y, int Z, int c4, int tiles, size_t slab_stride, int bx);

constexpr int TX_SYH = 18, TX_SZH = 10;          // haloed block extent in y and z (16 + 2, 8 + 2)
constexpr int TX_PLANE_VOX = TX_SYH * TX_SZH;    // voxels (16-byte units) per haloed x plane
constexpr int TX_PAIRS = 5;

// (dy,dz) tap t2 = dy*3 + dz sits t2/3 rows and t2%3 voxels into the haloed plane
__host__ __device__ constexpr int tx_off(int t2) { return (t2 / 3) * TX_SZH + t2 % 3; }
// K step p covers taps (first, second); step 4 is (tap 7 with zero weights, tap 8)
__host__ __device__ constexpr int tx_first(int p) { return p < 4 ? 2 * p : 7; }
__host__ __device__ constexpr int tx_second(int p) { return p < 4 ? 2 * p + 1 : 8; }

#ifdef TX_TIMING
// debug build only: cycles block 0's role warps spend waiting (see scripts/tcx_timing.py)
__device__ unsigned long long g_tx_timers[16];
#define TX_T0() const long long _t0 = clock64()
#define TX_ACC(var) var += clock64() - _t0
#else
#define TX_T0()
#define TX_ACC(var)
#endif

constexpr int TX_CONV_WARPS = 2;                // operand conversion (fp32 -> fp16 hi / lo' images, in place)
// + 8 warps (two per tensor-memory lane quarter, half the channels each) for accumulator drain, x shift-add, epilogue;
// Cout = 32: two accumulator sets of 192 columns, and the drain warps take registers from warps 0-3 (setmaxnreg)

template <int N, int BX, int STAGES, int DW = 8>
struct TxCfg {
    static constexpr bool N8 = (N == 8);
    static constexpr int NPR = 6 * N;                                  // B rows per K half
    static constexpr int N1 = 6 * N;
    static constexpr int N2 = N8 ? 6 * N : 3 * N;
    static constexpr int D2_COL = 3 * N;                               // first accumulator column of MMA2
    static constexpr int NPD = N8 ? 9 * N : 6 * N;                     // accumulator columns of one set
    static constexpr int NSETS = (4 * NPD <= 512) ? 4 : 2;             // accumulator sets in flight
    static constexpr int SXH = BX + 2;
    static constexpr int PLANE = SXH * TX_PLANE_VOX * 16;              // bytes of one operand image of the block
    static constexpr int B_BYTES = TX_PAIRS * 2 * NPR * 16;
    static constexpr int STAGE = 2 * PLANE + B_BYTES;
    static constexpr int TMEM_COLS = NSETS * NPD <= 256 ? 256 : 512;
    static constexpr int SMEM = STAGES * STAGE + 1024;
    // 8 drain warps: two per tensor-memory lane quarter, half the channels each.  16: additionally each warp owns only
    // half of the unit's output planes (PH = 2) -- twice the warps to hide tensor-memory / shared-memory latencies
    // behind, half the accumulator registers per thread.  Measured on B200 (38 tiles, all 13 blocks): 16 warps are
    // 14 % SLOWER than 8 (6.51 vs 5.68 ms) -- the extra warps spin on the same accumulator barriers and take issue
    // slots from the ones that work -- so only DW = 8 is instantiated.
    static constexpr int DRAIN_WARPS = DW;
    static constexpr int PH = DW / 8;                                  // plane halves
    static constexpr int PB = BX / PH;                                 // output planes per drain thread
    static constexpr int THREADS = 64 + 32 * (TX_CONV_WARPS + DRAIN_WARPS);
    static constexpr int CH = N / 2;                                   // output channels per drain thread
    static_assert(DW == 8 || DW == 16, "8 or 16 drain warps");
    // Cout = 32 (acc[8][16] = 128 registers per drain thread): the three x-taps of a set are loaded one at a time (32
    // live registers instead of 96), and the drain warps take registers from the producer / MMA / converter warpgroup.
    // setmaxnreg moves registers inside the CTA's launch-time pool (384 threads x 168): 128 x 56 + 256 x 224 = 64512.
    static constexpr bool SPLIT_LD = (N == 32);
#ifdef CT_NO_REBALANCE
    static constexpr bool REBALANCE = false;                           // debug build: no setmaxnreg (spills instead)
#else
    static constexpr bool REBALANCE = (N == 32) || DW == 16;
#endif
    // setmaxnreg targets: 384 threads x 168 -> 128 x 56 + 256 x 224; 640 threads x 96 -> 128 x 32 + 512 x 112
    static constexpr int REG_DEC = DW == 16 ? 32 : 56, REG_INC = DW == 16 ? 112 : 224;
    static constexpr bool POOL = (N == 16);                            // fused (2,2,1) max-pool epilogue (d0b only)
    static_assert(NSETS * NPD <= 512, "accumulators exceed tensor memory");
    static_assert(N1 % 16 == 0 && N1 <= 256 && N2 % 16 == 0, "UMMA N out of range for M = 128");
    static_assert(PLANE % 128 == 0 && B_BYTES % 128 == 0, "stage parts must stay 128-byte aligned");
    static_assert(SMEM <= 232448, "shared memory ring too large");
};

struct TxGeom {
    int cin8, X, Y, Z, nbx, nby, nbz, units;
    int dst_c4off;
    size_t dst_tile_stride4;
    const float* amax_src;
    float* amax_dst;
    size_t slab_stride;
    float w_inv_scale;
    // optional fused MaxPooling3D((2,2,1)) of the output (unet3d.py:168): pooled copy written by the epilogue
    float4* pool_dst;                            // null = no pooling; c4-blocked [Cout/4][X/2][Y/2][Z][4] per tile
    float* amax_pool;                            // slab header slot of the pooled buffer
    // 1: dst already holds partial pre-activation sums of this block (the up-sampled half of a decoder block's
    // concatenated input, convolved on the low-resolution grid by unet_tcu.cu); the epilogue adds them in
    int add_partial;
    // split-fp16 buffers (unet_common.cuh): header scale slots of the source / destination, the a-priori output bound
    // of the block, and (decoder blocks) the max|x| slot of the up-sampled half that fed the partial sums
    const float* scale_src;
    float* scale_dst;
    const float* amax_src2;
    float bound_p, bound_q;
};

struct TxUnit { int x0, y0, z0, tile; };
__device__ __forceinline__ TxUnit tx_unit(int u, const TxGeom& g, int bx) {
    TxUnit r;
    r.x0 = (u % g.nbx) * bx; u /= g.nbx;
    r.y0 = (u % g.nby) * 16; u /= g.nby;
    r.z0 = (u % g.nbz) * 8;
    r.tile = u / g.nbz;
    return r;
}

// issue only: the caller batches several loads in front of one tcgen05.wait::ld
template <int CH>
__device__ __forceinline__ void tmem_ld_issue(uint32_t taddr, uint32_t (&r)[CH]) {
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

// Persistent CTA.  Work unit = BX x 16 x 8 output voxels x all Cout of one tile.  Stage g = one 8-channel chunk of one
// unit (ring slot g % STAGES); accumulator step a = g * (BX+2) + j = input plane j of stage g (tensor-memory set
// a % NSETS).  Warp roles: 0 TMA producer (+ tensor-memory allocator), 1 MMA issuer, 2-3 operand conversion,
// 4-11 accumulator drain / x shift-add / epilogue.
// SRC_SPLIT: the source buffer already holds the fp16 hi / lo' operand images (no conversion warps, the MMA issuer waits
// for the TMA directly).  DST_SPLIT: the epilogue writes the destination in that form.
template <int N, int BX, int STAGES, int DW, bool SRC_SPLIT, bool DST_SPLIT>
__global__ void __launch_bounds__(TxCfg<N, BX, STAGES, DW>::THREADS, 1)
conv3_tcx_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wpack,
                 const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                 float alpha, float4* __restrict__ dst, const TxGeom geo) {
    using Cfg = TxCfg<N, BX, STAGES, DW>;
    constexpr int SXH = Cfg::SXH, CH = Cfg::CH, NSETS = Cfg::NSETS, PB = Cfg::PB;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[STAGES], bar_conv[STAGES], bar_empty[STAGES], bar_acc_full[NSETS], bar_acc_empty[NSETS];
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
            mbar_init(&bar_conv[s], TX_CONV_WARPS);
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
    if (threadIdx.x >= 64) {
        for (int i = threadIdx.x - 64; i < 3 * N; i += Cfg::THREADS - 64)
            ep_s[i / N][i % N] = (i < N) ? bias[i] : (i < 2 * N ? scale[i - N] : shift[i - 2 * N]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 2 + TX_CONV_WARPS) {
    // warpgroup 0 (producer, MMA issuer, converters) gives registers to the drain warpgroups
    if constexpr (Cfg::REBALANCE) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REG_DEC));
    if (warp == 0) {
        // ---------------- TMA producer
        if (elect_one()) {
            int g = 0;
            for (int k = 0; k < n_units; ++k) {
                const TxUnit un = tx_unit((int)blockIdx.x + k * (int)gridDim.x, geo, BX);
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
            constexpr uint64_t a_hi_word = (uint64_t)((uint32_t)TX_SZH | (1u << 14)) << 32;   // SBO: y rows 160 B apart
            constexpr uint64_t b_hi_word = (uint64_t)(8u | (1u << 14)) << 32;                 // SBO: 8-row groups 128 B apart
            const uint32_t ring16 = smem_u32(ring) >> 4;
            int a = 0;
            long long w_conv = 0, w_acc = 0, t_begin = clock64();
            for (int g = 0; g < n_stages; ++g) {
                const int s = g % STAGES, use = g / STAGES;
                { TX_T0(); mbar_wait(SRC_SPLIT ? &bar_full[s] : &bar_conv[s], use & 1); TX_ACC(w_conv); }
                if constexpr (SRC_SPLIT) tc_fence_after();
                const uint32_t a_hi = ring16 + (uint32_t)s * (Cfg::STAGE / 16), a_lo = a_hi + Cfg::PLANE / 16;
                const uint32_t b_base = a_hi + 2 * (Cfg::PLANE / 16);
#pragma unroll 1
                for (int j = 0; j < SXH; ++j, ++a) {
                    const int set = a % NSETS, use_a = a / NSETS;
                    if (use_a > 0) { TX_T0(); mbar_wait(&bar_acc_empty[set], (use_a - 1) & 1); TX_ACC(w_acc); }
                    tc_fence_after();
                    const uint32_t d = tmem_base + (uint32_t)set * Cfg::NPD;
                    const uint32_t pl = (uint32_t)j * TX_PLANE_VOX;
#pragma unroll
                    for (int p = 0; p < TX_PAIRS; ++p) {
                        const uint32_t lbo = (uint32_t)(tx_off(tx_second(p)) - tx_off(tx_first(p))) << 16;
                        const uint32_t ah = (a_hi + pl + (uint32_t)tx_off(tx_first(p))) | lbo;
                        const uint32_t al = (a_lo + pl + (uint32_t)tx_off(tx_first(p))) | lbo;
                        const uint32_t b32 = (b_base + (uint32_t)p * (Cfg::NPR * 2)) | ((uint32_t)Cfg::NPR << 16);
                        umma_f16(d, a_hi_word | (uint64_t)ah, b_hi_word | (uint64_t)b32, idesc1, p != 0);
                        umma_f16(d + Cfg::D2_COL, a_hi_word | (uint64_t)al, b_hi_word | (uint64_t)b32, idesc2, 1u);
                    }
                    umma_commit(&bar_acc_full[set]);
                }
                umma_commit(&bar_empty[s]);
            }
#ifdef TX_TIMING
            if (blockIdx.x == 0) { g_tx_timers[0] = w_conv; g_tx_timers[1] = w_acc; g_tx_timers[2] = clock64() - t_begin; }
#else
            (void)w_conv; (void)w_acc; (void)t_begin;
#endif
        }
        __syncwarp();
    } else if constexpr (!SRC_SPLIT) {
        // ---------------- converters: fp32 -> fp16 hi / lo' images of every landed stage, in place
        const int ct = threadIdx.x - 64;
        int g = 0;
        long long w_full = 0, t_begin = clock64();
        for (int k = 0; k < n_units; ++k) {
            const TxUnit un = tx_unit((int)blockIdx.x + k * (int)gridDim.x, geo, BX);
            const float sc = tc_operand_scale(geo.amax_src[(size_t)un.tile * geo.slab_stride]);
            for (int c = 0; c < cin8; ++c, ++g) {
                const int s = g % STAGES, use = g / STAGES;
                { TX_T0(); mbar_wait(&bar_full[s], use & 1); TX_ACC(w_full); }
                uint4* p0 = reinterpret_cast<uint4*>(ring + (size_t)s * Cfg::STAGE);
                uint4* p1 = p0 + Cfg::PLANE / 16;
#pragma unroll 2
                for (int i = ct; i < Cfg::PLANE / 16; i += 32 * TX_CONV_WARPS) {
                    const float4 v0 = *reinterpret_cast<const float4*>(p0 + i);      // channels 0-3 of the voxel
                    const float4 v1 = *reinterpret_cast<const float4*>(p1 + i);      // channels 4-7
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
#ifdef TX_TIMING
        if (blockIdx.x == 0 && threadIdx.x == 64) { g_tx_timers[3] = w_full; g_tx_timers[4] = clock64() - t_begin; }
#else
        (void)w_full; (void)t_begin;
#endif
    }
    } else {
        // ---------------- drain warps: tensor memory -> registers with the x shift-add, epilogue
        if constexpr (Cfg::REBALANCE) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REG_INC));
        const int q = warp & 3;                                    // tensor-memory lane quarter this warp may read
        const int dwi = warp - 2 - TX_CONV_WARPS;                  // drain warp index
        const int part = (dwi >> 2) & 1;                           // which CH channels this thread owns
        const int ch0 = part * CH;
        const int row = q * 32 + lane;
        const size_t vol = (size_t)geo.X * geo.Y * geo.Z;
        constexpr float W2 = 1.f / 2048.f;                         // weight of the hi.lo' + lo'.hi column group
        constexpr int TERMS = 2;                                   // column groups per x-tap

        // the body is instantiated per plane half so that every accumulator index stays a compile-time constant
        auto run = [&](auto ph_c) {
        constexpr int I0 = decltype(ph_c)::value * PB;             // first output plane this warp owns
        float2 acc[PB][CH / 2];                                    // channel pairs: packed fp32 arithmetic (tc_ptx.cuh)
        int a = 0;
        long long w_accf = 0, t_epi = 0, t_begin = clock64();
        // max|x| (and, for split sources, the operand scale) of the unit's tile is fetched one unit ahead: the load's
        // latency never sits in front of the epilogue
        TxUnit un_next = tx_unit((int)blockIdx.x, geo, BX);
        auto tile_amax = [&](int tile, float& am, float& am2) {
            am = geo.amax_src[(size_t)tile * geo.slab_stride];
            am2 = (DST_SPLIT && geo.amax_src2) ? geo.amax_src2[(size_t)tile * geo.slab_stride] : 0.f;
        };
        float am_next = 0.f, am2_next = 0.f;
        if (n_units > 0) tile_amax(un_next.tile, am_next, am2_next);
        float sc_next = (SRC_SPLIT && n_units > 0) ? geo.scale_src[(size_t)un_next.tile * geo.slab_stride] : 1.f;
        for (int k = 0; k < n_units; ++k) {
            const TxUnit un = un_next;
            const float s_in = SRC_SPLIT ? sc_next : tc_operand_scale(am_next);
            const float inv_scale = geo.w_inv_scale / s_in;
            // destination operand scale (power of two) from the a-priori bound of this block's output on this tile
            const float s_out = DST_SPLIT ? split_out_scale(fmaxf(am_next, am2_next), geo.bound_p, geo.bound_q) : 1.f;
            if (DST_SPLIT && warp == 2 + TX_CONV_WARPS && lane == 0) {
                geo.scale_dst[(size_t)un.tile * geo.slab_stride] = s_out;
                if (Cfg::POOL && geo.pool_dst != nullptr) geo.amax_pool[SCALE_SLOT0 + (size_t)un.tile * geo.slab_stride] = s_out;
            }
            if (k + 1 < n_units) {
                un_next = tx_unit((int)blockIdx.x + (k + 1) * (int)gridDim.x, geo, BX);
                tile_amax(un_next.tile, am_next, am2_next);
                if constexpr (SRC_SPLIT) sc_next = geo.scale_src[(size_t)un_next.tile * geo.slab_stride];
            }
            const int y = un.y0 + (row >> 3), z = un.z0 + (row & 7);
            float4* d_base = dst + (size_t)un.tile * geo.dst_tile_stride4 + (size_t)geo.dst_c4off * vol;
            float4* d_tile = d_base + (size_t)(ch0 / 4) * vol;
            float amax = 0.f;
            // epilogue of ONE output plane: scale back, bias -> activation -> BatchNorm, 16-byte channel-chunk stores.
            // Plane i is complete once input plane i + 2 of the last chunk is drained, so its stores are issued there
            // and trickle out under the remaining drains instead of bursting at the end of the unit.
            // With pooling fused, the pair of x planes (2p, 2p+1) meets in this thread and the pair of y rows sits 8
            // lanes apart (row = 8 y + z), so a 2 x 2 x 1 window is one register max and one shuffle.
            // Split destination: plane 2 c8 of the buffer takes the fp16 hi image of channels [8 c8, 8 c8 + 8), plane
            // 2 c8 + 1 the lo' image; a thread that owns 4 channels (Cout = 8) writes its 8-byte half of both.
            float2 keep[CH / 2];
            auto put = [&](float4* base, size_t pvol, size_t pvox, const float2 (&v)[CH / 2]) {
                if constexpr (!DST_SPLIT) {
#pragma unroll
                    for (int c4 = 0; c4 < CH / 4; ++c4)
                        base[(size_t)(ch0 / 4 + c4) * pvol + pvox] = make_float4(v[c4 * 2].x, v[c4 * 2].y, v[c4 * 2 + 1].x, v[c4 * 2 + 1].y);
                } else {
                    uint32_t hi[CH / 2], lo[CH / 2];
#pragma unroll
                    for (int h2 = 0; h2 < CH / 2; ++h2) {
                        // x s -> hi = RN16(x s), lo' = RN16((x s - hi) 2^11): the difference and its scaling are exact
                        const float2 os = f2_mul(v[h2], f2_splat(s_out));
                        const __half2 h = __floats2half2_rn(os.x, os.y);
                        const float2 hf = __half22float2(h);
                        const float2 d = f2_fma(hf, f2_splat(-2048.f), f2_mul(os, f2_splat(2048.f)));
                        const __half2 l = __floats2half2_rn(d.x, d.y);
                        hi[h2] = *reinterpret_cast<const uint32_t*>(&h);
                        lo[h2] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    if constexpr (CH >= 8) {
                        uint4* b4 = reinterpret_cast<uint4*>(base);
#pragma unroll
                        for (int c8 = 0; c8 < CH / 8; ++c8) {
                            b4[(size_t)(ch0 / 4 + 2 * c8) * pvol + pvox] = make_uint4(hi[4 * c8], hi[4 * c8 + 1], hi[4 * c8 + 2], hi[4 * c8 + 3]);
                            b4[(size_t)(ch0 / 4 + 2 * c8 + 1) * pvol + pvox] = make_uint4(lo[4 * c8], lo[4 * c8 + 1], lo[4 * c8 + 2], lo[4 * c8 + 3]);
                        }
                    } else {
                        uint2* b2 = reinterpret_cast<uint2*>(base);
                        b2[pvox * 2 + (ch0 / 4)] = make_uint2(hi[0], hi[1]);
                        b2[(pvol + pvox) * 2 + (ch0 / 4)] = make_uint2(lo[0], lo[1]);
                    }
                }
            };
            auto store_plane = [&](int i) {
                const int x = un.x0 + i;
                const bool ok = (y < geo.Y && x < geo.X);
                const size_t vox = ((size_t)x * geo.Y + y) * geo.Z + z;
                float2 o[CH / 2];
#pragma unroll
                for (int kk = 0; kk < CH / 2; ++kk) {
                    const int ch = ch0 + 2 * kk;
                    float2 t = f2_fma(acc[i - I0][kk], f2_splat(inv_scale), *reinterpret_cast<const float2*>(&ep_s[0][ch]));
                    const float2 ta = f2_mul(t, f2_splat(alpha));      // alpha in [0, 1]: t > 0 ? t : alpha t == max(t, alpha t)
                    t = make_float2(fmaxf(t.x, ta.x), fmaxf(t.y, ta.y));
                    o[kk] = f2_fma(t, *reinterpret_cast<const float2*>(&ep_s[1][ch]), *reinterpret_cast<const float2*>(&ep_s[2][ch]));
                    if (ok) amax = fmaxf(amax, fmaxf(fabsf(o[kk].x), fabsf(o[kk].y)));
                }
                if (ok) put(d_base, vol, vox, o);
                if (Cfg::POOL && geo.pool_dst != nullptr) {            // uniform over the CTA
                    if ((i & 1) == 0) {
#pragma unroll
                        for (int kk = 0; kk < CH / 2; ++kk) keep[kk] = o[kk];
                    } else {
                        float2 m[CH / 2];
#pragma unroll
                        for (int kk = 0; kk < CH / 2; ++kk) {
                            m[kk] = make_float2(fmaxf(keep[kk].x, o[kk].x), fmaxf(keep[kk].y, o[kk].y));
                            m[kk].x = fmaxf(m[kk].x, __shfl_xor_sync(0xffffffffu, m[kk].x, 8));
                            m[kk].y = fmaxf(m[kk].y, __shfl_xor_sync(0xffffffffu, m[kk].y, 8));
                        }
                        if (ok && ((row >> 3) & 1) == 0) {
                            const int PX = geo.X >> 1, PY = geo.Y >> 1;
                            put(geo.pool_dst + (size_t)un.tile * geo.dst_tile_stride4, (size_t)PX * PY * geo.Z,
                                ((size_t)(x >> 1) * PY + (y >> 1)) * geo.Z + z, m);
                        }
                    }
                }
            };
#pragma unroll
            for (int i = 0; i < PB; ++i)
#pragma unroll
                for (int ch = 0; ch < CH / 2; ++ch) acc[i][ch] = make_float2(0.f, 0.f);
            if (geo.add_partial) {                                  // uniform over the CTA
                // The accumulators START from the partial sums the phase kernel left in dst (decoder blocks), which it
                // wrote in this kernel's accumulator units (TuGeom::post_*): the loads land straight in the accumulator
                // registers and are first needed by the first drain's add, so their latency hides behind the unit's
                // first MMAs.  (Loading them in the per-plane epilogue put a dependent global load in front of every
                // plane's store: ~40 % of the block's time; rescaling them here stalled every unit's head: ~15 %.)
#pragma unroll
                for (int i = 0; i < PB; ++i) {
                    const int x = un.x0 + I0 + i;
                    if (y < geo.Y && x < geo.X) {
                        const size_t vox = ((size_t)x * geo.Y + y) * geo.Z + z;
#pragma unroll
                        for (int c4 = 0; c4 < CH / 4; ++c4) {
                            const float4 pv = d_tile[(size_t)c4 * vol + vox];
                            acc[i][c4 * 2 + 0] = make_float2(pv.x, pv.y);
                            acc[i][c4 * 2 + 1] = make_float2(pv.z, pv.w);
                        }
                    }
                }
                // Cout = 8 with a split destination: the two threads of a voxel read one fp32 partial plane each but
                // write an 8-byte half of BOTH planes -- every partial must be read before any output is stored
                if constexpr (DST_SPLIT && CH < 8) asm volatile("bar.sync 1, %0;" ::"n"(32 * Cfg::DRAIN_WARPS) : "memory");
            }
#pragma unroll 1
            for (int c = 0; c < cin8; ++c) {
                const bool last = (c == cin8 - 1);
#pragma unroll
                for (int j = 0; j < SXH; ++j, ++a) {
                    const int set = a % NSETS, use_a = a / NSETS;
                    { TX_T0(); mbar_wait(&bar_acc_full[set], use_a & 1); TX_ACC(w_accf); }
                    tc_fence_after();
                    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)set * Cfg::NPD + (uint32_t)ch0;
                    if constexpr (Cfg::SPLIT_LD) {
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const int i = j - dx;
                            if (i < I0 || i >= I0 + PB) continue;  // output plane j - dx is fed through x-tap dx
                            uint32_t v[TERMS][CH];
#pragma unroll
                            for (int t = 0; t < TERMS; ++t) tmem_ld_issue<CH>(t0 + t * 3 * N + dx * N, v[t]);
                            tmem_ld_wait();
#pragma unroll
                            for (int ch = 0; ch < CH / 2; ++ch)
                                acc[i - I0][ch] = f2_add(acc[i - I0][ch],
                                    f2_fma(make_float2(__uint_as_float(v[1][2 * ch]), __uint_as_float(v[1][2 * ch + 1])), f2_splat(W2),
                                           make_float2(__uint_as_float(v[0][2 * ch]), __uint_as_float(v[0][2 * ch + 1]))));
                        }
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_acc_empty[set]);      // values are in registers: set is free
                    } else {
                        uint32_t v[3][TERMS][CH];
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            if (j - dx < I0 || j - dx >= I0 + PB) continue;  // output plane j - dx is fed through x-tap dx
#pragma unroll
                            for (int t = 0; t < TERMS; ++t) tmem_ld_issue<CH>(t0 + t * 3 * N + dx * N, v[dx][t]);
                        }
                        tmem_ld_wait();
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&bar_acc_empty[set]);      // values are in registers: set is free
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx) {
                            const int i = j - dx;
                            if (i < I0 || i >= I0 + PB) continue;
#pragma unroll
                            for (int ch = 0; ch < CH / 2; ++ch)
                                acc[i - I0][ch] = f2_add(acc[i - I0][ch],
                                    f2_fma(make_float2(__uint_as_float(v[dx][1][2 * ch]), __uint_as_float(v[dx][1][2 * ch + 1])), f2_splat(W2),
                                           make_float2(__uint_as_float(v[dx][0][2 * ch]), __uint_as_float(v[dx][0][2 * ch + 1]))));
                        }
                    }
                    if (last && j - 2 >= I0 && j - 2 < I0 + PB) { TX_T0(); store_plane(j - 2); TX_ACC(t_epi); }
                }
            }
            amax = warp_max(amax);
            if (lane == 0) {
                amax_update(geo.amax_dst + (size_t)un.tile * geo.slab_stride, amax);
                if (Cfg::POOL && geo.pool_dst != nullptr) amax_update(geo.amax_pool + (size_t)un.tile * geo.slab_stride, amax);
            }
        }
#ifdef TX_TIMING
        if (blockIdx.x == 0 && warp == 2 + TX_CONV_WARPS && lane == 0) {
            g_tx_timers[5] = w_accf; g_tx_timers[6] = t_epi; g_tx_timers[7] = clock64() - t_begin;
        }
#else
        (void)w_accf; (void)t_epi; (void)t_begin;
#endif
        };
        if constexpr (Cfg::PH == 1) run(std::integral_constant<int, 0>{});
        else if (dwi < 8) run(std::integral_constant<int, 0>{});
        else run(std::integral_constant<int, 1>{});
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
static int tx_rows(int cout) { return 6 * cout; }

size_t tcx_weight_floats(int cin_pad, int cout) {
    if (cout != 8 && cout != 16 && cout != 32) return 0;
    return (size_t)((cin_pad + 7) / 8) * TX_PAIRS * 2 * tx_rows(cout) * 4;      // 16 bytes per row per K half
}

// keras kernel (kx,ky,kz,ci,co) -> fp16 image [ci/8][K step][K half][row][ci % 8] with rows
// dx*cout + co (hi) | 3 cout + dx*cout + co (lo'); same power-of-two
// scale as the classic packing (max|w| in [2^13, 2^14)).  Returns 1 / scale.
float tcx_pack_weights(const float* w, int cin, int cin_pad, int cout, float* dst) {
    (void)cin_pad;
    return tcx_pack_weights_range(w, cin, 0, cin, cout, dst);
}

float tcx_pack_weights_range(const float* w, int cin_total, int c_begin, int cin, int cout, float* dst) {
    const int cin_pad = (cin + 3) / 4 * 4;
    const int npr = tx_rows(cout), c8n = (cin_pad + 7) / 8;
    std::memset(dst, 0, tcx_weight_floats(cin_pad, cout) * sizeof(float));
    float wmax = 0.f;
    for (int tap = 0; tap < 27; ++tap)
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                wmax = std::fmax(wmax, std::fabs(w[((size_t)tap * cin_total + c_begin + ci) * cout + co]));
    int e = 0;
    if (wmax > 0.f) std::frexp(wmax, &e);
    const float scale = std::ldexp(1.f, 14 - e);
    __half* img = reinterpret_cast<__half*>(dst);
    for (int c = 0; c < c8n; ++c)
        for (int p = 0; p < TX_PAIRS; ++p)
            for (int j = 0; j < 2; ++j) {
                if (p == 4 && j == 0) continue;          // tap 7's second appearance carries zero weights
                const int t2 = j == 0 ? tx_first(p) : tx_second(p);
                __half* blk = img + ((((size_t)c * TX_PAIRS + p) * 2 + j) * npr) * 8;
                for (int dx = 0; dx < 3; ++dx)
                    for (int co = 0; co < cout; ++co)
                        for (int qd = 0; qd < 8; ++qd) {
                            const int ci = c * 8 + qd;
                            if (ci >= cin) continue;
                            const int tap = dx * 9 + t2;
                            const float v = w[((size_t)tap * cin_total + c_begin + ci) * cout + co] * scale;
                            const __half h = __float2half_rn(v);
                            const __half l = __float2half_rn((v - __half2float(h)) * 2048.f);
                            blk[(size_t)(dx * cout + co) * 8 + qd] = h;
                            blk[(size_t)(3 * cout + dx * cout + co) * 8 + qd] = l;
                        }
            }
    return 1.f / scale;
}

struct TxSource { const float* w; float inv_scale; int cin8; int add_partial; const float* amax2; };

template <int N, int BX, int STAGES, int DW, bool SRC_SPLIT, bool DST_SPLIT>
static int launch_tcx(const CUtensorMap& map, const ConvLayer& L, const TxSource& src, float alpha, float4* dst, int X, int Y, int Z,
                      size_t stride4, int dst_c4off, int tiles, const float* amax_src, float* amax_dst,
                      float4* pool_dst, float* amax_pool, cudaStream_t s) {
    using Cfg = TxCfg<N, BX, STAGES, DW>;
    // per device / context attribute: set on every launch (cheap) so several GPUs in one process are correct
    CT_CUDA(cudaFuncSetAttribute(conv3_tcx_kernel<N, BX, STAGES, DW, SRC_SPLIT, DST_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
    TxGeom g;
    g.cin8 = src.cin8; g.X = X; g.Y = Y; g.Z = Z;
    g.amax_src = amax_src; g.amax_dst = amax_dst; g.slab_stride = stride4 * 4; g.w_inv_scale = src.inv_scale;
    g.nbx = cdiv(X, BX); g.nby = cdiv(Y, 16); g.nbz = Z / 8;
    g.units = g.nbx * g.nby * g.nbz * tiles;
    g.dst_c4off = dst_c4off; g.dst_tile_stride4 = stride4;
    g.pool_dst = pool_dst; g.amax_pool = amax_pool; g.add_partial = src.add_partial;
    g.scale_src = amax_src + SCALE_SLOT0; g.scale_dst = amax_dst + SCALE_SLOT0; g.amax_src2 = src.amax2;
    g.bound_p = L.bound_p; g.bound_q = L.bound_q;
    const int sms = tc_sm_count() - g_reserved_sms;
    const int grid = g.units < sms ? g.units : sms;
    conv3_tcx_kernel<N, BX, STAGES, DW, SRC_SPLIT, DST_SPLIT><<<grid, Cfg::THREADS, Cfg::SMEM, s>>>(map, src.w, L.bias, L.scale, L.shift, alpha, dst, g);
    return 0;
}

template <bool SRC_SPLIT, bool DST_SPLIT>
static int launch_tcx_n(int cout, const CUtensorMap& map, const ConvLayer& L, const TxSource& src, float alpha, float4* dst, int X,
                        int Y, int Z, size_t stride4, int dst_c4off, int tiles, const float* amax_src, float* amax_dst,
                        float4* pool_dst, float* amax_pool, cudaStream_t s) {
    if (cout == 8) return launch_tcx<8, 8, 3, 8, SRC_SPLIT, DST_SPLIT>(map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
    if (cout == 16) return launch_tcx<16, 8, 3, 8, SRC_SPLIT, DST_SPLIT>(map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
    return launch_tcx<32, 8, 2, 8, SRC_SPLIT, DST_SPLIT>(map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
}

static int launch_tcx_fmt(int fmt, int cout, const CUtensorMap& map, const ConvLayer& L, const TxSource& src, float alpha, float4* dst,
                          int X, int Y, int Z, size_t stride4, int dst_c4off, int tiles, const float* amax_src, float* amax_dst,
                          float4* pool_dst, float* amax_pool, cudaStream_t s) {
    switch (fmt & 3) {
    case 0: return launch_tcx_n<false, false>(cout, map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
    case FMT_SRC_SPLIT: return launch_tcx_n<true, false>(cout, map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
    case FMT_DST_SPLIT: return launch_tcx_n<false, true>(cout, map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
    default: return launch_tcx_n<true, true>(cout, map, L, src, alpha, dst, X, Y, Z, stride4, dst_c4off, tiles, amax_src, amax_dst, pool_dst, amax_pool, s);
    }
}

// returns 2 when the layer is not handled by the stacked kernel (the caller falls back to unet_tc.cu's kernel)
int launch_conv_tcx(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                    const Op* pool, bool* pool_fused, int fmt) {
    const ConvLayer& L = net->layers[op.layer];
    const int X = op.sx, Y = op.sy, Z = op.sz;
    if (!L.w_tcx || Z % 8 != 0) return 2;
    if ((fmt & FMT_SRC_SPLIT) && L.cin_pad % 8 != 0) return 2;
    CT_REQUIRE(op.src_c == L.cin_pad, "conv: source buffer has %d channels, layer expects %d", op.src_c, L.cin_pad);
    CT_REQUIRE(slab_stride % 4 == 0 && op.src_off % 4 == 0 && op.dst_off % 4 == 0, "conv: misaligned slab");
    CT_REQUIRE(!(fmt & FMT_DST_SPLIT) || op.dst_coff % 8 == 0, "conv: split destination at channel offset %d", op.dst_coff);
    float4* dst = reinterpret_cast<float4*>(slab0 + op.dst_off);
    CUtensorMap map;
    ProfScope prof(PROF_CONV, s);
    const size_t st4 = slab_stride / 4;
    const int co4 = op.dst_coff / 4, c4 = L.cin_pad / 4;
    float* src = slab0 + op.src_off;
    const float* am_s = slab0 + op.src_slot;
    float* am_d = slab0 + op.dst_slot;
    if (tc_make_map(&map, src, X, Y, Z, c4, tiles, slab_stride, 8)) return 1;
    // the MaxPooling3D that follows this block in the plan, when it is the (2,2,1) pool of exactly this output
    float4* pool_dst = nullptr;
    float* am_p = nullptr;
    if (pool_fused) *pool_fused = false;
    if (pool && pool_fused && pool->kind == OP_POOL && pool->src_off == op.dst_off && pool->src_coff == op.dst_coff &&
        pool->c == L.cout && pool->dst_coff == 0 && pool->dst_c == L.cout && net->spec.pool_x == 2 && net->spec.pool_y == 2 &&
        net->spec.pool_z == 1 && X % 2 == 0 && Y % 2 == 0 && pool->dx == X / 2 && pool->dy == Y / 2 && pool->dz == Z &&
        pool->dst_off % 4 == 0 && L.cout == 16) {
        pool_dst = reinterpret_cast<float4*>(slab0 + pool->dst_off);
        am_p = slab0 + pool->dst_slot;
        *pool_fused = true;
    }
    const TxSource whole{L.w_tcx, L.w_tc_inv_scale, (L.cin_pad + 7) / 8, 0, nullptr};
    if (launch_tcx_fmt(fmt, L.cout, map, L, whole, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, pool_dst, am_p, s)) return 1;
    CT_LAUNCHED("conv3_tcx_kernel");
    return 0;
}

// The skip half of a decoder block: input channels [c_up, cin) of the concatenation buffer, partial sums of the
// up-sampled half (unet_tcu.cu) already in the destination.  up_slot = header slot of the low-resolution source of
// that half (its max|x| enters the output bound of a split destination).
int launch_conv_tcx_skip(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                         int fmt, int up_slot) {
    const ConvLayer& L = net->layers[op.layer];
    const int X = op.sx, Y = op.sy, Z = op.sz;
    if (!L.w_tcx_skip || L.c_up <= 0 || Z % 8 != 0) return 2;
    const int c_skip = L.cin - L.c_up;
    CT_REQUIRE(c_skip % 8 == 0 && L.c_up % 4 == 0, "conv: skip half of layer %d has %d channels", op.layer, c_skip);
    float4* dst = reinterpret_cast<float4*>(slab0 + op.dst_off);
    CUtensorMap map;
    ProfScope prof(PROF_CONV, s);
    const size_t st4 = slab_stride / 4, vol = (size_t)X * Y * Z;
    float* src = slab0 + op.src_off + (size_t)L.c_up * vol;           // c4-blocked: channel chunk c starts at c * vol * 4
    if (tc_make_map(&map, src, X, Y, Z, c_skip / 4, tiles, slab_stride, 8)) return 1;
    const TxSource skip{L.w_tcx_skip, L.w_tcx_skip_inv_scale, c_skip / 8, 1, up_slot >= 0 ? slab0 + up_slot : nullptr};
    const float* am_s = slab0 + op.src_slot;
    float* am_d = slab0 + op.dst_slot;
    const int co4 = op.dst_coff / 4;
    if (launch_tcx_fmt(fmt, L.cout, map, L, skip, net->alpha, dst, X, Y, Z, st4, co4, tiles, am_s, am_d, nullptr, nullptr, s)) return 1;
    CT_LAUNCHED("conv3_tcx_kernel");
    return 0;
}

}  // namespace ct

#ifdef TX_TIMING
extern "C" int ct_debug_tcx_timers(unsigned long long* out16) {
    return cudaMemcpyFromSymbol(out16, ct::g_tx_timers, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1;
}
#endif
