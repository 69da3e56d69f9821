// Internal structures of the U-Net engine (shared by the CUDA-core and tcgen05 convolution paths).
//
// Activation layout in HBM ("c4 blocked"): [tile][C/4][X][Y][Z][4] float32, i.e. channel chunks of 4 are
// the slowest per-tile dimension and the 4 channels of a chunk are interleaved per voxel (16 B).  z is the
// fastest spatial axis, as in the reference's (x, y, z) arrays.  One voxel-chunk is one 16-byte vector:
// it is the unit of coalesced global access, of shared-memory staging, and -- for the tensor-core path --
// exactly one row of a no-swizzle K-major UMMA core matrix (8 rows x 16 B).
#pragma once
#include "common.cuh"
#include <cuda_fp16.h>
#include <vector>

namespace ct {

struct ConvLayer {
    int cin, cout;            // logical channels (cin of the first layer is 1, stored padded to 4)
    int cin_pad;              // multiple of 4
    float* w_direct;          // device: [cin_pad/4][27][cout][4]   (CUDA-core kernel)
    float* w_tc;              // device: tcgen05 layout (see unet_tc.cu), may be null
    float* w_tcx;             // device: x-stacked tcgen05 layout (see unet_tcx.cu), may be null
    float w_tc_inv_scale;     // 1 / (power-of-two scale applied to the fp16 weight images)
    // decoder blocks that read concatenate([UpSampling3D(x), skip]) (unet3d.py:96-97): c_up = channels of the up-sampled
    // half; w_tcu = phase weights of that half on the low-resolution grid (unet_tcu.cu), w_tcx_skip = x-stacked image of
    // the skip half alone.  c_up = 0 / null pointers for every other block.
    int c_up;
    float* w_tcu;
    float w_tcu_inv_scale;
    float* w_tcx_skip;
    float w_tcx_skip_inv_scale;
    // plane-walk kernel (unet_tcz.cu): image of the whole block and of the skip half alone; null when not applicable
    float* w_tcz;
    float w_tcz_inv_scale;
    float* w_tcz_skip;
    float w_tcz_skip_inv_scale;
    // a-priori bound of the block's output, |out| <= bound_p * max|in| + bound_q  (sum of |w| per output channel, bias,
    // |alpha| <= 1, BatchNorm affine): the operand scale of a destination written in split-fp16 form is derived from it
    float bound_p, bound_q;
    float* bias;              // device [cout]
    float* scale;             // device [cout]  gamma / sqrt(var + eps)
    float* shift;             // device [cout]  beta - mean * scale
};

enum OpKind { OP_CONV = 0, OP_POOL = 1, OP_UPSAMPLE = 2 };

struct Op {
    OpKind kind;
    int layer;                // OP_CONV: index into layers
    size_t src_off, dst_off;  // float offsets inside one tile's workspace slab
    int src_c, dst_c;         // total channels of the src / dst buffers
    int src_coff, dst_coff;   // channel offset inside the buffer (multiple of 4)
    int c;                    // channels moved (pool / upsample) or cout (conv)
    int sx, sy, sz;           // spatial size of src
    int dx, dy, dz;           // spatial size of dst
    int src_slot, dst_slot;   // index of the buffers' max|value| slots in the per-tile slab header
};

// Every tile slab starts with a header of AMAX_SLOTS floats: slot b holds an upper bound of max|value| of buffer b
// of that tile (zeroed per batch; producers atomicMax the bit pattern, which orders like the value for floats >= 0).
// The tensor-core convolution scales its fp16 operand images by a power of two derived from it.
constexpr int AMAX_SLOTS = 64;
// Split-fp16 activation buffers.  Between two tensor-core blocks an activation tensor is stored as the operand images the
// consumer's MMAs read -- per 8 channels one 16-byte plane of fp16 `hi` and one of fp16 `lo'` with
// x * s = hi + lo' * 2^-11 -- instead of two 4-channel fp32 planes: the same bytes at the same addresses (plane
// 2 c8 = hi, plane 2 c8 + 1 = lo'), so TMA boxes and channel offsets are unchanged, but the consumer needs no
// conversion pass through shared memory.  s is a power of two per tile, chosen by the PRODUCER before it has seen its
// outputs, from the a-priori bound of ConvLayer::bound_p/q; it is published in the header's scale slot of the buffer
// (slots [SCALE_SLOT0, SCALE_SLOT0 + AMAX_SLOTS)).  A loose bound costs nothing: hi and lo' are floating point, so
// their relative precision holds until an element is 2^-13 / s small.
constexpr int FMT_SRC_SPLIT = 1, FMT_DST_SPLIT = 2;
constexpr int SCALE_SLOT0 = AMAX_SLOTS;
constexpr int HDR_FLOATS = 2 * AMAX_SLOTS;

}  // namespace ct

struct CtUNet {
    CtUNetSpec spec;
    std::vector<ct::ConvLayer> layers;
    std::vector<ct::Op> ops;
    size_t slab_floats;       // workspace floats per tile
    size_t in_off, last_off;  // offsets of the padded input buffer and of the last conv output
    int in_slot;              // header slot of the input buffer
    int last_c;
    float* head_w;            // device [last_c]
    float head_b;
    float alpha;              // 0.3 (LeakyReLU) or 0 (ReLU)
    bool split_ok;            // activation buffers between the blocks can be split-fp16 (unet.cu: use_split)
    int engine;               // 0 auto, 1 direct, 2 tcgen05 (stacked for Cout 8/16, else 27-tap), 3 27-tap only,
                              // 4 stacked wherever the shape allows (= 2 today); 5 (internal, ct_unet_conv_block only):
                              // plane-walk kernel wherever it can run
    double flops_per_tile;
    float* all_dev;           // one allocation holding every device array
};

namespace ct {
struct TileGeom;
// implemented in unet_direct.cu
// First block of the network (Cin = 1, Cout = 8) fused with the tile gather: reads the normalised volume (mode 0,
// reflect padding) or explicit tiles (mode 1) and writes the block's output.  Returns 2 when the layer shape is not
// the fused kernel's (the caller then runs gather_tiles + a generic convolution).
int launch_first_conv_fused(const CtUNet* net, const Op& op, const float* src, int mode, int tile_first,
                            const TileGeom& geo, float* slab0, size_t slab_stride, int tiles, cudaStream_t s);
int launch_conv_direct(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles,
                       cudaStream_t s);
// implemented in unet_tc.cu (returns 2 when the layer shape is not supported by the tensor-core path)
int launch_conv_tc(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles,
                   cudaStream_t s, int fmt = 0);
// implemented in unet_tcx.cu: x-stacked variant for Cout 8/16/32 (returns 2 when it does not take the layer)
// `pool` = the op that follows in the plan; when it is the (2,2,1) max-pool of this block's output the kernel writes
// the pooled copy itself and sets *pool_fused (the caller then skips that op).
// `fmt` = FMT_SRC_SPLIT | FMT_DST_SPLIT: which of the two buffers are in split-fp16 form.
int launch_conv_tcx(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles,
                    cudaStream_t s, const Op* pool = nullptr, bool* pool_fused = nullptr, int fmt = 0);
size_t tcx_weight_floats(int cin_pad, int cout);
float tcx_pack_weights(const float* keras_kernel, int cin, int cin_pad, int cout, float* dst);
// input channels [c_begin, c_begin + c_count) of the kernel only (the skip half of a decoder block)
float tcx_pack_weights_range(const float* keras_kernel, int cin, int c_begin, int c_count, int cout, float* dst);
// x-stacked block over the skip half of a concatenation, adding the partial sums the phase kernel left in dst
int launch_conv_tcx_skip(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                         int fmt = 0, int up_slot = -1);
// implemented in unet_tcz.cu: plane-walk variant ((dx,dz) taps stacked in N, dy taps in K) for Cout 8/16/32 between
// split-fp16 buffers of tiles with z = 16 (returns 2 when it does not take the block); same contract as launch_conv_tcx
int launch_conv_tcz(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                    const Op* pool = nullptr, bool* pool_fused = nullptr, int fmt = 0);
int launch_conv_tcz_skip(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles, cudaStream_t s,
                         int fmt = 0, int up_slot = -1);
size_t tcz_weight_floats(int cin, int cout);
bool tcz_takes_skip(const ConvLayer& L);       // does launch_conv_tcz_skip take this decoder block?
float tcz_pack_weights_range(const float* keras_kernel, int cin_total, int c_begin, int c_count, int cout, float* dst);
// implemented in unet_tcu.cu: convolution over the up-sampled half, on the low-resolution grid
size_t tcu_weight_floats(int c_up, int cout);
float tcu_pack_weights(const float* keras_kernel, int cin, int c_up, int cout, float* dst);
int launch_conv_tcu(const CtUNet* net, const ConvLayer& L, float* slab0, size_t slab_stride, int tiles, size_t up_off,
                    int up_slot, int X, int Y, int Z, size_t dst_off, int dst_coff, cudaStream_t s, bool src_split,
                    int skip_slot, bool p8 = false);
size_t tc_weight_floats(int cin_pad, int cout);
// returns 1 / scale
float tc_pack_weights(const float* keras_kernel, int cin, int cin_pad, int cout, float* dst);

__device__ __forceinline__ int reflect_index(int j, int n) {
    // numpy.pad(mode='reflect') for arbitrarily wide pads: triangle wave of period 2(n-1)
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    j %= period;
    if (j < 0) j += period;
    return j < n ? j : period - j;
}

// Geometry of a tiled prediction.  The tiles visited are the sub-grid tlo <= (i,j,k) < tlo + tn of the volume's tile
// grid, enumerated row-major (k fastest); source voxels live in a box `in_lo .. in_lo + in_dim` of the volume
// (the whole volume on one GPU, the rank's haloed block under spatial decomposition) and results go to a box
// `out_lo .. out_lo + out_dim`.
struct TileGeom {
    int X, Y, Z;                  // whole volume (reflect padding and the final crop refer to it)
    int TX, TY, TZ;               // model input tile
    int tlo[3], tn[3];            // tile sub-grid: origin and extent
    int c[3], b[3];               // centre window size and shrink (= offset of the window inside the tile)
    int in_lo[3], in_dim[3];
    int out_lo[3], out_dim[3];
};
__device__ __forceinline__ void tile_ijk(const TileGeom& g, int ordinal, int& i, int& j, int& k) {
    k = g.tlo[2] + ordinal % g.tn[2]; ordinal /= g.tn[2];
    j = g.tlo[1] + ordinal % g.tn[1]; ordinal /= g.tn[1];
    i = g.tlo[0] + ordinal;
}

// Power-of-two operand scale of a tile for the fp16 hi/lo split of the tensor-core kernels: max|x s| lands in
// [2^13, 2^14), so nothing overflows fp16 and its subnormal threshold sits 2^27 below the largest element.
// `am` = the tile's max|x| bound from the slab header (0 or subnormal -> scale 1).
__device__ __forceinline__ float tc_operand_scale(float am) {
    const int e = (int)((__float_as_uint(am) >> 23) & 0xffu);          // biased exponent, 0 for zero / subnormal
    const int se = (267 - e > 254) ? 254 : 267 - e;                    // 2^(13 - floor(log2 am)), clamped finite
    return (e == 0) ? 1.f : __uint_as_float((uint32_t)se << 23);
}
// operand scale of a block's split-fp16 output from the a-priori bound |out| <= p * max|in| + q
__device__ __forceinline__ float split_out_scale(float am_in, float p, float q) {
    return tc_operand_scale(fmaf(p, am_in, q) * 1.001f);
}
// x * s -> (hi, lo') fp16 pairs of two consecutive channels
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn((a - hf.x) * 2048.f, (b - hf.y) * 2048.f);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float2 unsplit_pair(uint32_t hi, uint32_t lo, float inv_s) {
    const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
    const float2 lf = __half22float2(*reinterpret_cast<const __half2*>(&lo));
    return make_float2(fmaf(lf.x, 1.f / 2048.f, hf.x) * inv_s, fmaf(lf.y, 1.f / 2048.f, hf.y) * inv_s);
}
__device__ __forceinline__ void amax_update(float* slot, float v) {
    atomicMax(reinterpret_cast<unsigned int*>(slot), __float_as_uint(v));
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
}  // namespace ct
