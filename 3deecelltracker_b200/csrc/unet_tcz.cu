// tcgen05 implicit-GEMM 3x3x3 convolution, "plane-walk" variant: the x- and z-taps are stacked in N, the y-taps are K, and
// the sum over the x-taps is left in the tensor core's accumulators.
//
// Same operator, operand precision and buffers as unet_tcx.cu (Conv3D 3x3x3 'same' + bias -> LeakyReLU/ReLU ->
// BatchNorm(eval), unet3d.py:117-119 / :139-140; split-fp16 hi / lo' operand images, fp32 accumulation), but a GEMM
// decomposition in which every MMA is wide enough to be bound by tensor math instead of by the shared-memory fetch of
// its A tile (an M = 128, K = 16 MMA costs max(N/2, ~47 + N/6) clocks):
//
//     D_j[voxel (y,z), (dx, dz, co)] = sum over (dy, ci) of  in[plane j][y+dy-1, z, ci] * W[dx, dy, dz, ci, co]
//     out[x][y, z] = sum over (dx, dz) of D_(x+dx-1)[(y, z+dz-1), (dx, dz, co)]
//
//   * M tile = 8 y-rows x ALL 16 z of one x-plane (TMEM lane = 16 y + z).  The z-halo of a tile is Keras' zero padding,
//     so the z shift-add is one lane up / down inside a 16-lane group (two shuffles) with zeros at z = 0 / 15: no halo
//     rows are computed in y or z, and Y = 40 / 20 fit the 8-row tile (the 16-row tile of unet_tcx.cu wastes 17 / 37 %).
//   * N = 3 x-taps x 48 columns = 144 per group of 8 output channels.  A block of 48 columns is ONE OUTPUT PLANE's
//     accumulator: (hi.hi | hi.lo' + lo'.hi) x 3 z-taps x 8 channels.  Output planes live in a ring of such blocks in
//     tensor memory, and the MMA of input plane j accumulates onto the three consecutive blocks of output planes j-1, j,
//     j+1 (B rows ordered dx = 2, 1, 0): the sum over the x-taps happens IN the accumulators, and a plane is drained
//     ONCE, when its third contribution has landed (48 columns per thread and plane; the first version of this kernel
//     added the x-taps in registers, 144 columns per input plane: its drain took 900-1100 clocks per plane and group
//     against 260-400 clocks of MMAs).
//         MMA1 = A_hi  x image 0 (hi | lo' rows)      MMA2 = A_lo' x image 1 (0 | hi rows), both N = 144, same columns
//   * K = 16 per MMA = two (ci-chunk, dy) taps x 8 input channels; 3 cin/8 taps -> ceil(3 cin / 16) K steps; a plane's
//     accumulator takes 3 x 2 x steps MMAs (<= 36; the 27-tap kernel chains 28).
//   * The CTA walks a segment of x-planes of one (tile, y-block): one shared-memory stage = ONE x-plane of all input
//     channels (haloed rows x 16 z), loaded once and used by the three output planes it feeds: there is no x-halo
//     recomputation inside a segment, (S + 2) / S input planes per S output planes.
//   * The packed weights of the whole block stay resident in shared memory (<= 110 KB), loaded once per CTA.
//
// Source buffers are split-fp16 only (unet_common.cuh); the destination is split-fp16 or fp32 c4 planes.  Decoder
// blocks: the phase kernel (unet_tcu.cu) leaves the partial sums of the up-sampled half in the destination in the
// "P8" layout -- a thread's 2 x 16 bytes (8 channels of a voxel) sit exactly where that thread later writes its hi / lo'
// images -- so a thread only ever reads bytes it overwrites itself.
#include "unet_common.cuh"
#include "tc_ptx.cuh"
#include <cmath>
#include <cstdlib>
#include <cstring>

namespace ct {

int tc_sm_count();
int tc_make_map_box(CUtensorMap* map, float* base, int X, int Y, int Z, int c4, int tiles, size_t slab_stride, const unsigned box[5]);

constexpr int TZ_Z = 16;                          // the kernel needs the whole z extent of a tile in one M tile
// The CTA always runs TWO lanes, each with its own ring of output planes, MMA issuer and drain warps: at Cout = 16 the two
// groups of 8 output channels of one 8-row M tile (NY = 1), at Cout = 8 two 8-row M tiles of the same plane (NY = 2).  Two
// issuers overlap the ~365 clocks of latency a thread pays per plane around its MMAs; each stays in plane order.
constexpr int TZ_ROW16 = TZ_Z;                    // 16-byte units per y row of one operand image
__host__ __device__ constexpr int tz_by(int ny) { return 8 * ny; }                 // y rows per unit
__host__ __device__ constexpr int tz_yh(int ny) { return 8 * ny + 2; }             // ... haloed
__host__ __device__ constexpr int tz_img16(int ny) { return tz_yh(ny) * TZ_ROW16; } // one operand image (8 channels) of one plane
__host__ __device__ constexpr int tz_chunk16(int ny) { return 2 * tz_img16(ny); }   // hi + lo'
constexpr int TZ_BLK = 48;                        // accumulator columns of one output plane and group: (hi.hi | cross) x 3 dz x 8
constexpr int TZ_NMMA = 3 * TZ_BLK;               // one MMA feeds three consecutive output planes (dx = 2, 1, 0)
constexpr int TZ_NROW = TZ_NMMA;                  // B rows per K half
constexpr int TZ_WIMG16 = 2 * TZ_NROW;            // 16-byte units of one B image (two K halves); two images per (K step, group)
constexpr int TZ_LANES = 2;
constexpr int TZ_R = 5;                           // ring slots of output planes per lane
constexpr int TZ_LCOLS = TZ_R * TZ_BLK;
static_assert(TZ_LANES * TZ_LCOLS <= 512, "plane rings exceed tensor memory");
constexpr int TZ_DRAIN_WARPS = 8;
constexpr int TZ_THREADS = 128 + 32 * TZ_DRAIN_WARPS;      // warpgroup 0: producer + three MMA issuers
constexpr int TZ_MAX_STAGES = 8;
constexpr int TZ_SMEM_MAX = 232448;

#ifdef TZ_TIMING
// debug build only: cycles block 0's role warps spend waiting (scripts/tcz_timing.py)
__device__ unsigned long long g_tz_timers[16];
#define TZ_T0() const long long _t0 = clock64()
#define TZ_ACC(var) var += clock64() - _t0
#else
#define TZ_T0()
#define TZ_ACC(var)
#endif

struct TzGeom {
    int cin8, nsteps, X, Y, by, nby, nseg, sxseg, units, stages;
    uint32_t wbytes;
    int dst_c4off;
    size_t dst_tile_stride4, slab_stride;
    const float* amax_src;
    float* amax_dst;
    const float* scale_src;
    float* scale_dst;
    const float* amax_src2;                       // decoder blocks: max|x| slot of the up-sampled half's source
    float w_inv_scale, bound_p, bound_q;
    int add_partial;                              // dst holds P8 partial sums of the up-sampled half (split destination only)
};

struct TzUnit { int x0, nout, y0, tile; };
__device__ __forceinline__ TzUnit tz_unit(int u, const TzGeom& g) {
    TzUnit r;
    r.x0 = (u % g.nseg) * g.sxseg; u /= g.nseg;
    r.nout = min(g.sxseg, g.X - r.x0);
    r.y0 = (u % g.nby) * g.by;
    r.tile = u / g.nby;
    return r;
}

// K-half taps of step p: tap t = (ci chunk t / 3, dy = t % 3) at t/3 chunks + t%3 rows into the plane's stage
__host__ __device__ constexpr uint32_t tz_tap_off16(int t, int chunk16) { return (uint32_t)(t / 3) * chunk16 + (uint32_t)(t % 3) * TZ_ROW16; }

// the 48 accumulator columns of one output plane and group, of the thread's lane
__device__ __forceinline__ void tz_ld48(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr));
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]),
                   "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47])
                 : "r"(taddr + 32));
}

// Persistent CTA.  Work unit = a segment of x-planes x 8 or 16 y-rows x 16 z x all Cout of one tile.  Warp roles: 0 TMA producer
// (+ tensor-memory allocator, resident weights), 1-2 MMA issuers (one per lane), 4-11 drain / epilogue (one warp per lane
// and tensor-memory lane quarter).
//
// Output planes live in a RING of tensor-memory blocks (48 columns per plane and lane).  Every plane of the CTA's
// whole sequence of units has an id G (two pseudo ids separate consecutive units: they take the contributions that fall
// outside a segment and are discarded); the MMA of the input plane with newest id G accumulates onto ids G-2, G-1, G
// = three consecutive blocks starting at block (G-2) mod R -- issued in two pieces where the ring wraps.  The FIRST
// contribution to a block (its id as the newest of the three) overwrites it (accumulate = 0), so nothing is ever
// zeroed: at K step 0 the MMA is split into the 96 columns of the two older planes (accumulate) and the 48 columns of
// the newest one (overwrite).  One thread issues all MMAs of a group in plane order, so a block always receives
// newest -> middle -> oldest, K step by K step: the rounding of a voxel does not depend on the segment, the batch or
// the grid (tile-sharded and spatially decomposed runs stay bit-identical to the single-GPU run).
//   drain -> issuer: blk_free[g][slot]  the id that last used the slot has been read out of tensor memory
//   issuer -> drain: blk_done[g][slot]  (tcgen05.commit after the MMAs with newest id G) id G-2 is complete
// Drain: each warp takes WHOLE blocks (all 8 channels of a plane, 48 columns per thread) of one lane -- warps 4-7 lane 0,
// warps 8-11 lane 1 -- so two items are in flight per tensor-memory lane quarter and the load latency of one (hundreds
// of clocks while MMAs read-modify-write their accumulators) overlaps the epilogue of the other.  The block is handed
// back right after the load, before the epilogue.
template <int CIN8, int NGRP, bool DST_SPLIT>
__global__ void __launch_bounds__(TZ_THREADS, 1)
conv3_tcz_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ wpack,
                 const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                 float alpha, float4* __restrict__ dst, const TzGeom geo) {
    constexpr int R = TZ_R, GCOLS = TZ_LCOLS, NG = TZ_LANES;
    constexpr int NY = 3 - NGRP;                                   // one channel group: two 8-row M tiles per plane
    constexpr int IMG16 = tz_img16(NY), CHUNK16 = tz_chunk16(NY);
    static_assert(NGRP == 1 || NGRP == 2, "one or two groups of 8 output channels");
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t bar_full[TZ_MAX_STAGES], bar_empty[TZ_MAX_STAGES], blk_free[NG][R], blk_done[NG][R], bar_w;
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(16) float ep_s[3][8 * NGRP];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* wsm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* ring = wsm + geo.wbytes;
    const int stages = geo.stages;
    constexpr uint32_t stage_bytes = (uint32_t)CIN8 * (CHUNK16 * 16);
    constexpr int NTAPS = 3 * CIN8, NSTEPS = (NTAPS + 1) / 2;
    const int n_units = ((int)blockIdx.x < geo.units) ? (geo.units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], NG);                          // one commit per issuer
        }
#pragma unroll
        for (int g = 0; g < NG; ++g)
#pragma unroll
            for (int r = 0; r < R; ++r) {
                mbar_init(&blk_free[g][r], TZ_DRAIN_WARPS / 2);
                mbar_init(&blk_done[g][r], 1);
            }
        mbar_init(&bar_w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (threadIdx.x >= 128) {
        for (int i = threadIdx.x - 128; i < 3 * 8 * NGRP; i += TZ_THREADS - 128)
            ep_s[i / (8 * NGRP)][i % (8 * NGRP)] = (i < 8 * NGRP) ? bias[i] : (i < 16 * NGRP ? scale[i - 8 * NGRP] : shift[i - 16 * NGRP]);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ---------------- TMA producer: resident weights once, then one x-plane of all input channels per stage
        if (elect_one()) {
            mbar_expect_tx(&bar_w, geo.wbytes);
            for (uint32_t o = 0; o < geo.wbytes; o += 32768u) {
                const uint32_t n = geo.wbytes - o < 32768u ? geo.wbytes - o : 32768u;
                bulk_load(wsm + o, reinterpret_cast<const uint8_t*>(wpack) + o, n, &bar_w);
            }
            int gp = 0;
            long long w_empty = 0;
            for (int k = 0; k < n_units; ++k) {
                const TzUnit un = tz_unit((int)blockIdx.x + k * (int)gridDim.x, geo);
                const int nh = un.nout + 2;
                for (int h = 0; h < nh; ++h, ++gp) {
                    const int s = gp % stages, use = gp / stages;
                    if (use > 0) { TZ_T0(); mbar_wait(&bar_empty[s], (use - 1) & 1); TZ_ACC(w_empty); }
                    mbar_expect_tx(&bar_full[s], stage_bytes);
                    tma_load_5d(ring + (size_t)s * stage_bytes, &tmap, &bar_full[s], 0, un.y0 - 1, un.x0 - 1 + h, 0, un.tile);
                }
            }
#ifdef TZ_TIMING
            if (blockIdx.x == 0) g_tz_timers[3] = w_empty;
#else
            (void)w_empty;
#endif
        }
        __syncwarp();
    } else if (warp <= TZ_LANES) {
        // ---------------- MMA issuer of group g = warp - 1 (one thread issues a group's MMAs, in plane order: the order in
        // which a block receives its contributions, hence its rounding, is fixed)
        if (elect_one()) {
            constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(TZ_NMMA >> 3) << 17) | (8u << 24);
            constexpr uint32_t idesc96 = (1u << 4) | ((uint32_t)((2 * TZ_BLK) >> 3) << 17) | (8u << 24);
            constexpr uint32_t idesc48 = (1u << 4) | ((uint32_t)(TZ_BLK >> 3) << 17) | (8u << 24);
            constexpr uint64_t sbo8_word = (uint64_t)(8u | (1u << 14)) << 32;     // 8-row groups 128 B apart (A and B)
            const uint32_t ring16 = smem_u32(ring) >> 4, w16 = smem_u32(wsm) >> 4;
            const int g = warp - 1;                                 // lane: channel group (NGRP = 2) or y half (NY = 2)
            const int gch = NGRP == 2 ? g : 0;
            const uint32_t yoff16 = NY == 2 ? (uint32_t)(g * 8 * TZ_ROW16) : 0u;
            const uint32_t gcol = tmem_base + (uint32_t)(g * GCOLS);
            mbar_wait(&bar_w, 0);
            int gp = 0;
            int slot = 2 % R, use = 2 / R;                          // ring slot / use count of the newest id (ids start at 2)
            long long w_full = 0, w_acc = 0, t_begin = clock64();
            for (int k = 0; k < n_units; ++k) {
                const TzUnit un = tz_unit((int)blockIdx.x + k * (int)gridDim.x, geo);
                const int nh = un.nout + 2;
                for (int h = 0; h < nh; ++h, ++gp) {
                    const int s = gp % stages, us = gp / stages;
                    { TZ_T0(); mbar_wait(&bar_full[s], us & 1); TZ_ACC(w_full); }
                    if (use > 0) { TZ_T0(); mbar_wait(&blk_free[g][slot], (use - 1) & 1); TZ_ACC(w_acc); }
                    tc_fence_after();
                    const uint32_t a_hi = ring16 + (uint32_t)s * (stage_bytes >> 4) + yoff16;
                    const int b = slot >= 2 ? slot - 2 : slot + R - 2;              // block of id G - 2
                    const uint32_t d = gcol + (uint32_t)(b * TZ_BLK);
#pragma unroll
                    for (int p = 0; p < NSTEPS; ++p) {
                        // the last step of an odd tap count repeats the tap before it against zero weights;
                        // every descriptor is the stage / weight base plus a compile-time constant
                        const int t1 = (2 * p + 1 < NTAPS) ? 2 * p + 1 : NTAPS - 1;
                        const int t0 = t1 - 1;
                        const uint32_t o0 = tz_tap_off16(t0, CHUNK16), lbo = (tz_tap_off16(t1, CHUNK16) - o0) << 16;
                        const uint64_t ah = sbo8_word | (uint64_t)((a_hi + o0) | lbo);
                        const uint64_t al = sbo8_word | (uint64_t)((a_hi + o0 + IMG16) | lbo);
                        const uint32_t w1 = (w16 + (uint32_t)((p * NGRP + gch) * 2) * TZ_WIMG16) | ((uint32_t)TZ_NROW << 16);
                        const uint32_t w2 = (w16 + (uint32_t)((p * NGRP + gch) * 2 + 1) * TZ_WIMG16) | ((uint32_t)TZ_NROW << 16);
                        // B rows [0, 48) feed the oldest plane (id G - 2), [48, 96) the middle one, [96, 144) the newest,
                        // which this plane's first MMA overwrites.  Where the ring wraps the MMA is issued in two pieces.
                        if (b <= R - 3) {
                            if (p == 0) {
                                umma_f16(d, ah, sbo8_word | (uint64_t)w1, idesc96, 1u);
                                umma_f16(d + 2 * TZ_BLK, ah, sbo8_word | (uint64_t)(w1 + 2 * TZ_BLK), idesc48, 0u);
                            } else {
                                umma_f16(d, ah, sbo8_word | (uint64_t)w1, idesc, 1u);
                            }
                            umma_f16(d, al, sbo8_word | (uint64_t)w2, idesc, 1u);
                        } else if (b == R - 2) {                    // the newest plane's block is block 0
                            umma_f16(d, ah, sbo8_word | (uint64_t)w1, idesc96, 1u);
                            umma_f16(gcol, ah, sbo8_word | (uint64_t)(w1 + 2 * TZ_BLK), idesc48, p == 0 ? 0u : 1u);
                            umma_f16(d, al, sbo8_word | (uint64_t)w2, idesc96, 1u);
                            umma_f16(gcol, al, sbo8_word | (uint64_t)(w2 + 2 * TZ_BLK), idesc48, 1u);
                        } else {                                    // b == R - 1: the middle and newest planes are blocks 0, 1
                            umma_f16(d, ah, sbo8_word | (uint64_t)w1, idesc48, 1u);
                            if (p == 0) {
                                umma_f16(gcol, ah, sbo8_word | (uint64_t)(w1 + TZ_BLK), idesc48, 1u);
                                umma_f16(gcol + TZ_BLK, ah, sbo8_word | (uint64_t)(w1 + 2 * TZ_BLK), idesc48, 0u);
                            } else {
                                umma_f16(gcol, ah, sbo8_word | (uint64_t)(w1 + TZ_BLK), idesc96, 1u);
                            }
                            umma_f16(d, al, sbo8_word | (uint64_t)w2, idesc48, 1u);
                            umma_f16(gcol, al, sbo8_word | (uint64_t)(w2 + TZ_BLK), idesc96, 1u);
                        }
                    }
                    umma_commit(&blk_done[g][b]);                   // id G - 2 has all three contributions
                    umma_commit(&bar_empty[s]);
                    if (++slot == R) { slot = 0; ++use; }
                }
            }
#ifdef TZ_TIMING
            if (blockIdx.x == 0 && g == 0) { g_tz_timers[0] = w_full; g_tz_timers[1] = w_acc; g_tz_timers[2] = clock64() - t_begin; }
#else
            (void)w_full; (void)w_acc; (void)t_begin;
#endif
        }
        __syncwarp();
    }
    } else {
        // ---------------- drain warps: finished planes tensor memory -> registers, z shift-add, epilogue
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int q = warp & 3;                                    // tensor-memory lane quarter this warp may read
        const int set = (warp - 4) >> 2;                           // which half of the (plane, group) items
        const int row = q * 32 + lane;
        const int yl = row >> 4, z = row & 15;
        const size_t vol = (size_t)geo.X * geo.Y * TZ_Z;
        constexpr float W2 = 1.f / 2048.f;                         // weight of the hi.lo' + lo'.hi columns
        const uint32_t t_lane = tmem_base + ((uint32_t)(q * 32) << 16);
        const int g_mine = set;                                    // lane of this warp set
        const int gch = NGRP == 2 ? set : 0, yoff = NY == 2 ? 8 * set : 0;
        const float m_up = (z == 0) ? 0.f : 1.f, m_dn = (z == TZ_Z - 1) ? 0.f : 1.f;      // zero padding at the z ends of a tile
        long long w_accf = 0, t_ld = 0, t_fin = 0, t_begin = clock64();
        int slot = 0, use = 0;                                     // ring slot / use count of the id being drained
        // the id sequence: two pseudo ids, then per unit its nout planes and two pseudo ids; the last two never complete
        TzUnit un_next = tz_unit((int)blockIdx.x, geo);
        float am_next = 0.f, am2_next = 0.f, sc_next = 1.f;
        auto tile_hdr = [&](int tile) {
            am_next = geo.amax_src[(size_t)tile * geo.slab_stride];
            am2_next = (DST_SPLIT && geo.amax_src2) ? geo.amax_src2[(size_t)tile * geo.slab_stride] : 0.f;
            sc_next = geo.scale_src[(size_t)tile * geo.slab_stride];
        };
        if (n_units > 0) tile_hdr(un_next.tile);
        // takes the finished id in (slot, use) out of tensor memory (real planes only) and hands its blocks back
        auto take = [&](bool real, uint32_t (&v)[48]) {
            { TZ_T0(); mbar_wait(&blk_done[g_mine][slot], use & 1); TZ_ACC(w_accf); }
            tc_fence_after();
#ifdef TZ_TIMING
            const long long _tl = clock64();
#endif
            if (real) {
                tz_ld48(t_lane + (uint32_t)(g_mine * GCOLS + slot * TZ_BLK), v);
                tmem_ld_wait();
            }
#ifdef TZ_TIMING
            t_ld += clock64() - _tl;
#endif
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&blk_free[g_mine][slot]);
        };
        auto next_id = [&]() { if (++slot == R) { slot = 0; ++use; } };
        if (n_units > 0) {
            uint32_t dummy[48];
            for (int e = 0; e < 2; ++e) {
                take(false, dummy);
                next_id();
            }
        }
        for (int k = 0; k < n_units; ++k) {
            const TzUnit un = un_next;
            const float inv_scale = geo.w_inv_scale / sc_next;
            const float s_out = DST_SPLIT ? split_out_scale(fmaxf(am_next, am2_next), geo.bound_p, geo.bound_q) : 1.f;
            if (DST_SPLIT && warp == 4 && lane == 0) {
                geo.scale_dst[(size_t)un.tile * geo.slab_stride] = s_out;
            }
            if (k + 1 < n_units) {
                un_next = tz_unit((int)blockIdx.x + (k + 1) * (int)gridDim.x, geo);
                tile_hdr(un_next.tile);
            }
            const int y = un.y0 + yoff + yl;
            const bool ok_y = y < geo.Y;
            uint8_t* d_tile = reinterpret_cast<uint8_t*>(dst + (size_t)un.tile * geo.dst_tile_stride4);
            float amax = 0.f;
            // the voxel's 16 bytes in the group's two channel planes: split buffers hold the fp16 hi image of the 8 channels
            // in plane 2 g and the lo' image in 2 g + 1; fp32 buffers channels 0-3 / 4-7; P8 partial sums channel pairs
            // (0,1 | 4,5) / (2,3 | 6,7)
            const size_t plane_b = (size_t)geo.Y * TZ_Z * 16, vol_b = vol * 16;
            uint8_t* const d_grp = d_tile + (size_t)(geo.dst_c4off + 2 * gch) * vol_b + (((size_t)un.x0 * geo.Y + y) * TZ_Z + z) * 16;
            auto put = [&](uint8_t* pa, size_t pvol_b, const float2 (&v)[4], float so) {
#ifdef TZ_EXP_NOSTORE
                if (so != 12345.678f) { amax = fmaxf(amax, v[0].x + v[1].y + v[2].x + v[3].y); return; }      // experiment: no stores
#endif
                if constexpr (!DST_SPLIT) {
                    *reinterpret_cast<float4*>(pa) = make_float4(v[0].x, v[0].y, v[1].x, v[1].y);
                    *reinterpret_cast<float4*>(pa + pvol_b) = make_float4(v[2].x, v[2].y, v[3].x, v[3].y);
                } else {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) split_pair2(v[kk], so, hi[kk], lo[kk]);
                    *reinterpret_cast<uint4*>(pa) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                    *reinterpret_cast<uint4*>(pa + pvol_b) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                }
            };
            // epilogue of output plane x0 + i from its 48 accumulator columns [term][dz][channel]: hi.hi + 2^-11 (hi.lo' +
            // lo'.hi), z shift-add of the three z-taps, partial sums, scale back, bias -> activation -> BatchNorm, store
            auto finish = [&](int i, const uint32_t (&v)[48], const float2 (&ps)[4]) {
                float2 o[4];
                float am = 0.f;
                // stage by stage over the four channel pairs, so that their dependency chains interleave
                float2 c[4][3], up[4], dn[4];
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int dz = 0; dz < 3; ++dz)
                        c[kk][dz] = f2_fma(make_float2(__uint_as_float(v[24 + dz * 8 + 2 * kk]), __uint_as_float(v[24 + dz * 8 + 2 * kk + 1])),
                                           f2_splat(W2),
                                           make_float2(__uint_as_float(v[dz * 8 + 2 * kk]), __uint_as_float(v[dz * 8 + 2 * kk + 1])));
                // out[z] takes tap dz = 0 from input row z - 1 and tap dz = 2 from input row z + 1 (zero outside the tile)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    up[kk] = make_float2(__shfl_up_sync(0xffffffffu, c[kk][0].x, 1), __shfl_up_sync(0xffffffffu, c[kk][0].y, 1));
                    dn[kk] = make_float2(__shfl_down_sync(0xffffffffu, c[kk][2].x, 1), __shfl_down_sync(0xffffffffu, c[kk][2].y, 1));
                }
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    // the zeros at the z ends: one packed multiply-add per neighbour (x 1 / x 0 is exact) -- a select on z
                    // compiles to a divergent branch per pair
                    const float2 av = f2_fma(up[kk], f2_splat(m_up), f2_fma(dn[kk], f2_splat(m_dn), f2_add(c[kk][1], ps[kk])));
                    o[kk] = block_epilogue(av, inv_scale, alpha, &ep_s[0][0], 8 * NGRP, 8 * gch + 2 * kk, am);
                }
                if (ok_y) {
                    amax = fmaxf(amax, am);
                    put(d_grp + (size_t)i * plane_b, vol_b, o, s_out);
                }
            };
            const bool last_unit = (k == n_units - 1);
#pragma unroll 1
            for (int i = 0; i < un.nout + 2; ++i) {
                const bool real = i < un.nout;
                if (last_unit && !real) break;                      // the sequence's last two pseudo ids never complete
                {
                    // partial sums (P8) of this plane: the loads fly while the warp waits for the plane to complete
                    float2 ps[4];
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) ps[kk] = make_float2(0.f, 0.f);
                    if (geo.add_partial && ok_y && real) {         // add_partial is uniform over the CTA
                        const uint8_t* pp = d_grp + (size_t)i * plane_b;
                        const float4 p0 = *reinterpret_cast<const float4*>(pp), p1 = *reinterpret_cast<const float4*>(pp + vol_b);
                        ps[0] = make_float2(p0.x, p0.y); ps[1] = make_float2(p1.x, p1.y);
                        ps[2] = make_float2(p0.z, p0.w); ps[3] = make_float2(p1.z, p1.w);
                    }
                    uint32_t v[48];
                    take(real, v);
                    if (real) { TZ_T0(); finish(i, v, ps); TZ_ACC(t_fin); }
                }
                next_id();
            }
            amax = warp_max(amax);
            if (lane == 0) {
                amax_update(geo.amax_dst + (size_t)un.tile * geo.slab_stride, amax);
            }
        }
#ifdef TZ_TIMING
        if (blockIdx.x == 0 && warp == 4 && lane == 0) {
            g_tz_timers[5] = w_accf; g_tz_timers[6] = t_ld; g_tz_timers[7] = t_fin; g_tz_timers[8] = clock64() - t_begin;
        }
#else
        (void)w_accf; (void)t_ld; (void)t_fin; (void)t_begin;
#endif
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int tz_steps(int cin) { return (3 * (cin / 8) + 1) / 2; }

size_t tcz_weight_floats(int cin, int cout) {
    if ((cout != 8 && cout != 16) || (cin != 8 && cin != 16 && cin != 32)) return 0;      // instantiated shapes
    const size_t bytes = (size_t)tz_steps(cin) * (cout / 8) * 2 * TZ_WIMG16 * 16;       // two B images per (K step, group)
    if (bytes + 3 * (size_t)(cin / 8) * tz_chunk16(cout == 8 ? 2 : 1) * 16 + 1024 > (size_t)TZ_SMEM_MAX) return 0;     // resident weights + 3 stages
    return bytes / 4;
}

// keras kernel (kx,ky,kz,ci,co), input channels [c_begin, c_begin + cin) -> fp16 images
// [K step][group co/8][image][K half][row][ci % 8].  Row = 48 blk + 24 term + 8 dz + co%8, blk 0 / 1 / 2 =
// x-tap dx 2 / 1 / 0 (the oldest of the three output planes an input plane feeds comes first).  Image 0 (times A_hi):
// term 0 = hi(w), term 1 = lo'(w); image 1 (times A_lo'): term 0 = 0, term 1 = hi(w) -- so hi.lo' and lo'.hi meet in
// the same accumulator columns.  K half j of step p is tap t = 2p + j -> (ci chunk t/3, dy = t%3); an odd tap count
// ends with (zero weights, last tap).  Same power-of-two scale as the x-stacked image of the same channel range (max|w|
// in [2^13, 2^14)).  Returns 1 / scale.
float tcz_pack_weights_range(const float* w, int cin_total, int c_begin, int cin, int cout, float* dst) {
    const int ng = cout / 8, ntaps = 3 * (cin / 8), nsteps = tz_steps(cin);
    std::memset(dst, 0, tcz_weight_floats(cin, cout) * sizeof(float));
    float wmax = 0.f;
    for (int tap = 0; tap < 27; ++tap)
        for (int ci = 0; ci < cin; ++ci)
            for (int co = 0; co < cout; ++co)
                wmax = std::fmax(wmax, std::fabs(w[((size_t)tap * cin_total + c_begin + ci) * cout + co]));
    int e = 0;
    if (wmax > 0.f) std::frexp(wmax, &e);
    const float scale = std::ldexp(1.f, 14 - e);
    __half* img = reinterpret_cast<__half*>(dst);
    for (int p = 0; p < nsteps; ++p)
        for (int g = 0; g < ng; ++g)
            for (int j = 0; j < 2; ++j) {
                int t = 2 * p + j;
                if (2 * p + 1 >= ntaps) { if (j == 0) continue; t = ntaps - 1; }      // odd tap count: (zero weights, last tap)
                const int c = t / 3, dy = t % 3;
                __half* img1 = img + (((((size_t)p * ng + g) * 2 + 0) * 2 + j) * TZ_NROW) * 8;
                __half* img2 = img + (((((size_t)p * ng + g) * 2 + 1) * 2 + j) * TZ_NROW) * 8;
                for (int blk = 0; blk < 3; ++blk)
                    for (int dz = 0; dz < 3; ++dz)
                        for (int col = 0; col < 8; ++col)
                            for (int qd = 0; qd < 8; ++qd) {
                                const int dx = 2 - blk, ci = c * 8 + qd, tap = (dx * 3 + dy) * 3 + dz;
                                const float v = w[((size_t)tap * cin_total + c_begin + ci) * cout + 8 * g + col] * scale;
                                const __half hh = __float2half_rn(v);
                                const __half ll = __float2half_rn((v - __half2float(hh)) * 2048.f);
                                const int r = blk * TZ_BLK + dz * 8 + col;
                                img1[(size_t)r * 8 + qd] = hh;
                                img1[(size_t)(r + 24) * 8 + qd] = ll;
                                img2[(size_t)(r + 24) * 8 + qd] = hh;
                            }
            }
    return 1.f / scale;
}

struct TzSource { const float* w; float inv_scale; int cin; int add_partial; const float* amax2; };

// fewest (rounds of units over the grid) x (planes per unit): segments of the x-walk trade halo planes against balance
static void tz_segments(int X, int per_x, int grid, bool even, int* sxseg, int* nseg) {
    long best = -1;
    for (int n = 1; n <= (X + 3) / 4; ++n) {
        int sx = cdiv(X, n);
        if (even && (sx & 1)) ++sx;
        const int ns = cdiv(X, sx);
        const long units = (long)per_x * ns, rounds = (units + grid - 1) / grid;
        const long cost = rounds * (sx + 3);                          // + 2 halo planes + ~1 plane of pipeline hand-over
        if (best < 0 || cost < best) { best = cost; *sxseg = sx; *nseg = ns; }
    }
}

template <int CIN8, int NGRP, bool DST_SPLIT>
static int launch_tcz(const CUtensorMap& map, const ConvLayer& L, const TzSource& src, float alpha, float4* dst, int X, int Y,
                      size_t stride4, int dst_c4off, int tiles, const float* amax_src, float* amax_dst, cudaStream_t s) {
    TzGeom g;
    g.cin8 = src.cin / 8; g.nsteps = tz_steps(src.cin); g.X = X; g.Y = Y;
    constexpr int NY = 3 - NGRP;
    g.by = tz_by(NY); g.nby = cdiv(Y, g.by);
    g.wbytes = (uint32_t)(tcz_weight_floats(src.cin, 8 * NGRP) * 4);
    const size_t stage_bytes = (size_t)g.cin8 * tz_chunk16(NY) * 16;
    int stages = (int)(((size_t)TZ_SMEM_MAX - 1024 - g.wbytes) / stage_bytes);
    if (stages > TZ_MAX_STAGES) stages = TZ_MAX_STAGES;
    CT_REQUIRE(stages >= 3, "conv: plane-walk kernel has no room for 3 stages (cin %d, cout %d)", src.cin, 8 * NGRP);
    g.stages = stages;
    const size_t smem = 1024 + g.wbytes + (size_t)stages * stage_bytes;
    const int sms = tc_sm_count() - g_reserved_sms;
    tz_segments(X, tiles * g.nby, sms, false, &g.sxseg, &g.nseg);
    g.units = tiles * g.nby * g.nseg;
    g.dst_c4off = dst_c4off; g.dst_tile_stride4 = stride4; g.slab_stride = stride4 * 4;
    g.amax_src = amax_src; g.amax_dst = amax_dst;
    g.scale_src = amax_src + SCALE_SLOT0; g.scale_dst = amax_dst + SCALE_SLOT0; g.amax_src2 = src.amax2;
    g.w_inv_scale = src.inv_scale; g.bound_p = L.bound_p; g.bound_q = L.bound_q;
    g.add_partial = src.add_partial;
    // per device / context attribute: set on every launch (cheap) so several GPUs in one process are correct
    CT_CUDA(cudaFuncSetAttribute(conv3_tcz_kernel<CIN8, NGRP, DST_SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int grid = g.units < sms ? g.units : sms;
    conv3_tcz_kernel<CIN8, NGRP, DST_SPLIT><<<grid, TZ_THREADS, smem, s>>>(map, src.w, L.bias, L.scale, L.shift, alpha, dst, g);
    return 0;
}

static int launch_tcz_any(int cout, bool dst_split, const CUtensorMap& map, const ConvLayer& L, const TzSource& src, float alpha,
                          float4* dst, int X, int Y, size_t stride4, int dst_c4off, int tiles, const float* amax_src,
                          float* amax_dst, cudaStream_t s) {
#define TZ_GO(C8, NG, DS, PL) return launch_tcz<C8, NG, DS>(map, L, src, alpha, dst, X, Y, stride4, dst_c4off, tiles, amax_src, amax_dst, s)
#define TZ_C8(NG, DS, PL) do { if (src.cin == 8) TZ_GO(1, NG, DS, PL); if (src.cin == 16) TZ_GO(2, NG, DS, PL); if (src.cin == 32) TZ_GO(4, NG, DS, PL); return 2; } while (0)
    if (cout == 8) { if (dst_split) TZ_C8(1, true, false); TZ_C8(1, false, false); }
    if (cout == 16) { if (dst_split) TZ_C8(2, true, false); TZ_C8(2, false, false); }
#undef TZ_C8
#undef TZ_GO
    return 2;
}

// Which blocks the plane-walk kernel takes in the `auto` mix [measured on B200, 38 tiles, ms against unet_tcx.cu]: every
// block it is instantiated for -- 16 -> 16: 0.20 vs 0.24; 8 -> 8: 0.31 vs 0.41; 32 -> 8: 0.75 vs 1.46; skip halves
// 16 -> 8: 0.48 vs 0.83, 32 -> 16: 0.39 vs 0.48 -- except 8 -> 16 (0.60 vs 0.55: the drain, ~350 instructions per warp
// and plane, is the bound, and that block also carries the fused pool).  CT3D_TCZ_ALL=1 routes it there as well.
static bool tz_all() {
    static int v = -1;
    if (v < 0) {
        const char* e = std::getenv("CT3D_TCZ_ALL");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
static bool tz_takes(int cin, int cout, bool skip_half) {
    if (cout > 16) return false;
    return tz_all() || skip_half || !(cin == 8 && cout == 16);
}

static int tz_map(CUtensorMap* map, float* src, int X, int Y, int cin, int cout, int tiles, size_t slab_stride) {
    const unsigned box[5] = {TZ_Z * 4, (unsigned)tz_yh(cout == 8 ? 2 : 1), 1, (unsigned)(cin / 4), 1};
    return tc_make_map_box(map, src, X, Y, TZ_Z, cin / 4, tiles, slab_stride, box);
}

bool tcz_takes_skip(const ConvLayer& L) { return L.w_tcz_skip != nullptr && L.c_up > 0 && tz_takes(L.cin - L.c_up, L.cout, true); }

// returns 2 when the block is not the plane-walk kernel's (the caller falls back to unet_tcx.cu / unet_tc.cu)
int launch_conv_tcz(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                    const Op* pool, bool* pool_fused, int fmt) {
    const ConvLayer& L = net->layers[op.layer];
    const int X = op.sx, Y = op.sy, Z = op.sz;
    (void)pool;                                  // the (2,2,1) pool after 8 -> 16 stays with unet_tcx.cu's fused epilogue
    if (pool_fused) *pool_fused = false;
    if (!L.w_tcz || Z != TZ_Z || !(fmt & FMT_SRC_SPLIT) || L.cin_pad != L.cin || (net->engine != 5 && !tz_takes(L.cin, L.cout, false)) || L.cout > 16) return 2;
    CT_REQUIRE(op.src_c == L.cin_pad, "conv: source buffer has %d channels, layer expects %d", op.src_c, L.cin_pad);
    CT_REQUIRE(slab_stride % 4 == 0 && op.src_off % 4 == 0 && op.dst_off % 4 == 0, "conv: misaligned slab");
    CT_REQUIRE(!(fmt & FMT_DST_SPLIT) || op.dst_coff % 8 == 0, "conv: split destination at channel offset %d", op.dst_coff);
    float4* dst = reinterpret_cast<float4*>(slab0 + op.dst_off);
    CUtensorMap map;
    ProfScope prof(PROF_CONV, s);
    if (tz_map(&map, slab0 + op.src_off, X, Y, L.cin, L.cout, tiles, slab_stride)) return 1;
    const TzSource whole{L.w_tcz, L.w_tcz_inv_scale, L.cin, 0, nullptr};
    const int rc = launch_tcz_any(L.cout, (fmt & FMT_DST_SPLIT) != 0, map, L, whole, net->alpha, dst, X, Y, slab_stride / 4,
                                  op.dst_coff / 4, tiles, slab0 + op.src_slot, slab0 + op.dst_slot, s);
    if (rc) return rc;
    CT_LAUNCHED("conv3_tcz_kernel");
    return 0;
}

// The skip half of a decoder block: input channels [c_up, cin) of the concatenation buffer; the destination already holds
// the partial sums of the up-sampled half in the P8 layout (unet_tcu.cu, post_layout 1).  Split buffers only.
int launch_conv_tcz_skip(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                         int fmt, int up_slot) {
    const ConvLayer& L = net->layers[op.layer];
    const int X = op.sx, Y = op.sy, Z = op.sz;
    if (!tcz_takes_skip(L) || Z != TZ_Z || fmt != (FMT_SRC_SPLIT | FMT_DST_SPLIT)) return 2;
    const int c_skip = L.cin - L.c_up;
    CT_REQUIRE(c_skip % 8 == 0 && L.c_up % 4 == 0 && op.dst_coff % 8 == 0, "conv: skip half of layer %d has %d channels", op.layer, c_skip);
    float4* dst = reinterpret_cast<float4*>(slab0 + op.dst_off);
    CUtensorMap map;
    ProfScope prof(PROF_CONV, s);
    const size_t vol = (size_t)X * Y * Z;
    if (tz_map(&map, slab0 + op.src_off + (size_t)L.c_up * vol, X, Y, c_skip, L.cout, tiles, slab_stride)) return 1;
    const TzSource skip{L.w_tcz_skip, L.w_tcz_skip_inv_scale, c_skip, 1, up_slot >= 0 ? slab0 + up_slot : nullptr};
    const int rc = launch_tcz_any(L.cout, true, map, L, skip, net->alpha, dst, X, Y, slab_stride / 4, op.dst_coff / 4, tiles,
                                  slab0 + op.src_slot, slab0 + op.dst_slot, s);
    if (rc) return rc;
    CT_LAUNCHED("conv3_tcz_kernel");
    return 0;
}

}  // namespace ct

#ifdef TZ_TIMING
extern "C" int ct_debug_tcz_timers(unsigned long long* out16) {
    return cudaMemcpyFromSymbol(out16, ct::g_tz_timers, sizeof(unsigned long long) * 16) == cudaSuccess ? 0 : 1;
}
#endif
