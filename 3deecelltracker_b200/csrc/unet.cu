// U-Net engine: weight packing, the per-tile op plan, tile gather / scatter and the element-wise
// kernels between convolutions.
//
// Reference semantics reproduced here:
//   unet3d.py:84-98   graph of _unet3_depth3 (and :40-67 unet3_b): two convs per level, MaxPooling3D,
//                     two convs on the LOWER level, UpSampling3D, concatenate([up, skip]), ..., 1x1x1 head
//   unet3d.py:203-256 unet3_prediction: reflect pre-pad, tiles of the model input size at a stride of the
//                     centre size, centre crop, scatter, final crop to the original size
// Every tile is convolved with its own zero 'same' padding (tiles are batch entries, never merged), so
// results depend on the tile grid exactly as in the reference.
#include "unet_common.cuh"
#include <cmath>
#include <cstring>

namespace ct {

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
// Tile gather.  mode 0: tile t of the volume's tile grid, reflect-padded (unet3d.py:235,247-249).
//               mode 1: tiles are given explicitly as (B, x, y, z) (Keras model.predict).
// Output: [tile][1 chunk][TX][TY][TZ][4] with channels 1..3 = 0.
__global__ void __launch_bounds__(256)
gather_tiles(const float* __restrict__ src, float4* __restrict__ slab, size_t slab_stride4, size_t in_off4,
             int mode, int tile_first, const TileGeom g, int in_slot) {
    // grid (TX, tiles): one x plane of one tile per block, threads over the (y, z) plane -- no 64-bit div / mod
    const int t = blockIdx.y, a = blockIdx.x;
    float amax = 0.f;
    const int TY = g.TY, TZ = g.TZ, plane = TY * TZ;
    const size_t tile_vox = (size_t)g.TX * plane;
    float4* out = slab + (size_t)t * slab_stride4 + in_off4 + (size_t)a * plane;
    int i, j, k;
    tile_ijk(g, tile_first + t, i, j, k);
    if (mode == 0) {
        const int sx = reflect_index(i * g.c[0] + a - g.b[0], g.X) - g.in_lo[0];
        const float* row0 = src + (size_t)sx * g.in_dim[1] * g.in_dim[2];
        for (int f = threadIdx.x; f < plane; f += blockDim.x) {
            const int b = f / TZ, c = f - b * TZ;
            const int sy = reflect_index(j * g.c[1] + b - g.b[1], g.Y) - g.in_lo[1];
            const int sz = reflect_index(k * g.c[2] + c - g.b[2], g.Z) - g.in_lo[2];
            const float val = row0[(size_t)sy * g.in_dim[2] + sz];
            out[f] = make_float4(val, 0.f, 0.f, 0.f);
            amax = fmaxf(amax, fabsf(val));
        }
    } else {
        const float* row0 = src + (size_t)(tile_first + t) * tile_vox + (size_t)a * plane;
        for (int f = threadIdx.x; f < plane; f += blockDim.x) {
            const float val = row0[f];
            out[f] = make_float4(val, 0.f, 0.f, 0.f);
            amax = fmaxf(amax, fabsf(val));
        }
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0) amax_update(reinterpret_cast<float*>(slab + (size_t)t * slab_stride4) + in_slot, amax);
}

// MaxPooling3D(pool) on c4-blocked buffers (unet3d.py:168).  grid (c4 * DX, tiles): one destination x plane of one
// channel chunk per block, threads over the destination (y, z) plane.
__global__ void __launch_bounds__(256)
pool_kernel(const float4* __restrict__ slab_src, float4* __restrict__ slab_dst, size_t slab_stride4,
            size_t src_off4, size_t dst_off4, int src_c4off, int c4, int SXs, int SYs, int SZs,
            int DX, int DY, int DZ, int px, int py, int pz, int src_slot, int dst_slot) {
    const int t = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) {          // max over a subset <= max of the source buffer
        float* hdr = reinterpret_cast<float*>(slab_dst + (size_t)t * slab_stride4);
        amax_update(hdr + dst_slot, hdr[src_slot]);
    }
    const int ck = blockIdx.x / DX, x = blockIdx.x - ck * DX;
    const size_t dvol = (size_t)DX * DY * DZ, svol = (size_t)SXs * SYs * SZs;
    const float4* src = slab_src + (size_t)t * slab_stride4 + src_off4 + (size_t)(src_c4off + ck) * svol;
    float4* dst = slab_dst + (size_t)t * slab_stride4 + dst_off4 + (size_t)ck * dvol + (size_t)x * DY * DZ;
    for (int f = threadIdx.x; f < DY * DZ; f += blockDim.x) {
        const int y = f / DZ, z = f - y * DZ;
        float4 m = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        for (int a = 0; a < px; ++a)
            for (int b = 0; b < py; ++b)
                for (int c = 0; c < pz; ++c) {
                    const float4 q = src[((size_t)(x * px + a) * SYs + (y * py + b)) * SZs + (z * pz + c)];
                    m.x = fmaxf(m.x, q.x); m.y = fmaxf(m.y, q.y); m.z = fmaxf(m.z, q.z); m.w = fmaxf(m.w, q.w);
                }
        dst[f] = m;
    }
}

// MaxPooling3D(pool) between split-fp16 buffers (unet_common.cuh): planes (2 c8, 2 c8 + 1) hold the fp16 hi / lo' images of
// 8 channels; all voxels of a tile share one scale, so the window maximum is the element with the largest
// hi + lo' 2^-11 (exact in fp32) and its two halves are copied unchanged.  grid (c8 * DX, tiles).
__global__ void __launch_bounds__(256)
pool_split_kernel(const uint4* __restrict__ slab_src, uint4* __restrict__ slab_dst, size_t slab_stride4,
                  size_t src_off4, size_t dst_off4, int src_c4off, int c8n, int SXs, int SYs, int SZs,
                  int DX, int DY, int DZ, int px, int py, int pz, int src_slot, int dst_slot) {
    const int t = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float* hdr = reinterpret_cast<float*>(slab_dst + (size_t)t * slab_stride4);
        amax_update(hdr + dst_slot, hdr[src_slot]);
        hdr[SCALE_SLOT0 + dst_slot] = hdr[SCALE_SLOT0 + src_slot];
    }
    const int ck = blockIdx.x / DX, x = blockIdx.x - ck * DX;
    const size_t dvol = (size_t)DX * DY * DZ, svol = (size_t)SXs * SYs * SZs;
    const uint4* src = slab_src + (size_t)t * slab_stride4 + src_off4 + (size_t)(src_c4off + 2 * ck) * svol;
    uint4* dst = slab_dst + (size_t)t * slab_stride4 + dst_off4 + (size_t)(2 * ck) * dvol + (size_t)x * DY * DZ;
    for (int f = threadIdx.x; f < DY * DZ; f += blockDim.x) {
        const int y = f / DZ, z = f - y * DZ;
        float best[8];
        uint32_t bh[4], bl[4];                           // fp16 pairs: channel 2k in the low half
#pragma unroll
        for (int k = 0; k < 8; ++k) best[k] = -INFINITY;
        for (int a = 0; a < px; ++a)
            for (int b = 0; b < py; ++b)
                for (int c = 0; c < pz; ++c) {
                    const size_t v = ((size_t)(x * px + a) * SYs + (y * py + b)) * SZs + (z * pz + c);
                    const uint4 h = src[v], l = src[svol + v];
                    const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float2 val = unsplit_pair(hh[k], ll[k], 1.f);
                        if ((a | b | c) == 0) {
                            best[2 * k] = val.x; best[2 * k + 1] = val.y; bh[k] = hh[k]; bl[k] = ll[k];
                        } else {
                            if (val.x > best[2 * k]) {
                                best[2 * k] = val.x;
                                bh[k] = (bh[k] & 0xffff0000u) | (hh[k] & 0xffffu);
                                bl[k] = (bl[k] & 0xffff0000u) | (ll[k] & 0xffffu);
                            }
                            if (val.y > best[2 * k + 1]) {
                                best[2 * k + 1] = val.y;
                                bh[k] = (bh[k] & 0xffffu) | (hh[k] & 0xffff0000u);
                                bl[k] = (bl[k] & 0xffffu) | (ll[k] & 0xffff0000u);
                            }
                        }
                    }
                }
        dst[f] = make_uint4(bh[0], bh[1], bh[2], bh[3]);
        dst[dvol + f] = make_uint4(bl[0], bl[1], bl[2], bl[3]);
    }
}

// UpSampling3D(size) nearest, written into the first channels of the concat buffer (unet3d.py:199).  grid (c4 * SX,
// tiles): one SOURCE x plane of one channel chunk per block; every source voxel is read once and stored px*py*pz times.
__global__ void __launch_bounds__(256)
upsample_kernel(const float4* __restrict__ slab_src, float4* __restrict__ slab_dst, size_t slab_stride4,
                size_t src_off4, size_t dst_off4, int dst_c4off, int c4, int SXs, int SYs, int SZs,
                int DX, int DY, int DZ, int px, int py, int pz, int src_slot, int dst_slot) {
    const int t = blockIdx.y;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float* hdr = reinterpret_cast<float*>(slab_dst + (size_t)t * slab_stride4);
        amax_update(hdr + dst_slot, hdr[src_slot]);
    }
    const int ck = blockIdx.x / SXs, x = blockIdx.x - ck * SXs;
    const size_t dvol = (size_t)DX * DY * DZ, svol = (size_t)SXs * SYs * SZs;
    const float4* src = slab_src + (size_t)t * slab_stride4 + src_off4 + (size_t)ck * svol + (size_t)x * SYs * SZs;
    float4* dst = slab_dst + (size_t)t * slab_stride4 + dst_off4 + (size_t)(dst_c4off + ck) * dvol;
    for (int f = threadIdx.x; f < SYs * SZs; f += blockDim.x) {
        const int y = f / SZs, z = f - y * SZs;
        const float4 q = src[f];
        for (int a = 0; a < px; ++a)
            for (int b = 0; b < py; ++b)
                for (int c = 0; c < pz; ++c)
                    dst[((size_t)(x * px + a) * DY + (y * py + b)) * DZ + (z * pz + c)] = q;
    }
}

// Head: Conv3D(1, 1, activation='sigmoid') (unet3d.py:96) + centre crop + scatter (unet3d.py:250-255).
// mode 0: write the centre window of tile (i,j,k) into prob (X,Y,Z), clipped to the volume.
// mode 1: write the whole tile to prob (B, TX, TY, TZ).
__global__ void __launch_bounds__(256)
head_scatter(const float4* __restrict__ slab, size_t slab_stride4, size_t last_off4, int c4,
             const float* __restrict__ head_w, float head_b, float* __restrict__ prob,
             int mode, int tile_first, const TileGeom g) {
    // grid (window x extent, tiles): one x plane of the written window per block, threads over its (y, z) plane
    const int t = blockIdx.y, a = blockIdx.x;
    const int TX = g.TX, TY = g.TY, TZ = g.TZ;
    const size_t tile_vox = (size_t)TX * TY * TZ;
    const float4* in = slab + (size_t)t * slab_stride4 + last_off4;
    int i, j, k;
    tile_ijk(g, tile_first + t, i, j, k);
    const int wy = mode == 0 ? g.c[1] : TY, wz = mode == 0 ? g.c[2] : TZ;
    const int ox = mode == 0 ? g.b[0] : 0, oy = mode == 0 ? g.b[1] : 0, oz = mode == 0 ? g.b[2] : 0;
    const int gx = i * g.c[0] + a, lx = gx - g.out_lo[0];
    if (mode == 0 && (gx >= g.X || lx < 0 || lx >= g.out_dim[0])) return;
    for (int f = threadIdx.x; f < wy * wz; f += blockDim.x) {
        const int b = f / wz, c = f - b * wz;
        const size_t tv = ((size_t)(a + ox) * TY + (b + oy)) * TZ + (c + oz);
        float acc = head_b;
        for (int ck = 0; ck < c4; ++ck) {
            const float4 q = in[(size_t)ck * tile_vox + tv];
            acc = fmaf(q.x, head_w[ck * 4 + 0], acc);
            acc = fmaf(q.y, head_w[ck * 4 + 1], acc);
            acc = fmaf(q.z, head_w[ck * 4 + 2], acc);
            acc = fmaf(q.w, head_w[ck * 4 + 3], acc);
        }
        const float p = 1.f / (1.f + expf(-acc));
        if (mode == 0) {
            const int gy = j * g.c[1] + b, gz = k * g.c[2] + c;
            const int ly = gy - g.out_lo[1], lz = gz - g.out_lo[2];
            if (gy < g.Y && gz < g.Z && ly >= 0 && lz >= 0 && ly < g.out_dim[1] && lz < g.out_dim[2])
                prob[((size_t)lx * g.out_dim[1] + ly) * g.out_dim[2] + lz] = p;
        } else {
            prob[(size_t)(tile_first + t) * tile_vox + tv] = p;
        }
    }
}

// Keras channels-last (B, X, Y, Z, C) <-> c4-blocked [B][C/4][X][Y][Z][4] (channel pad = 0); used by the
// single-block operator ct_unet_conv_block.
__global__ void __launch_bounds__(256)
ndhwc_to_c4(const float* __restrict__ src, float4* __restrict__ dst, size_t vol, int c, int c4, size_t total,
            float* __restrict__ amax_slot) {
    float amax = 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t v = idx % vol;
        const size_t r = idx / vol;
        const int ck = (int)(r % c4);
        const size_t b = r / c4;
        const float* s = src + (b * vol + v) * c + ck * 4;
        float q[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) q[j] = (ck * 4 + j < c) ? s[j] : 0.f;
        dst[idx] = make_float4(q[0], q[1], q[2], q[3]);
        amax = fmaxf(fmaxf(amax, fmaxf(fabsf(q[0]), fabsf(q[1]))), fmaxf(fabsf(q[2]), fabsf(q[3])));
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0) amax_update(amax_slot, amax);
}
__global__ void __launch_bounds__(256)
c4_to_ndhwc(const float4* __restrict__ src, float* __restrict__ dst, size_t vol, int c, int c4, size_t total) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t v = idx % vol;
        const size_t r = idx / vol;
        const int ck = (int)(r % c4);
        const size_t b = r / c4;
        const float4 q = src[idx];
        float* d = dst + (b * vol + v) * c + ck * 4;
        const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ck * 4 + j < c) d[j] = qq[j];
    }
}

// fp32 c4-blocked <-> split-fp16 (unet_common.cuh), in place: planes (2 c8, 2 c8 + 1) of a buffer hold channels
// [8 c8, 8 c8 + 8) either as two 4-channel fp32 planes or as the fp16 hi / lo' images.  grid (blocks, tiles).
__global__ void __launch_bounds__(256)
c4_to_split(float4* __restrict__ slab, size_t slab_stride4, size_t off4, int c8n, size_t vol, int slot) {
    float* hdr = reinterpret_cast<float*>(slab + (size_t)blockIdx.y * slab_stride4);
    const float sc = tc_operand_scale(hdr[slot]);
    if (blockIdx.x == 0 && threadIdx.x == 0) hdr[SCALE_SLOT0 + slot] = sc;
    float4* buf = slab + (size_t)blockIdx.y * slab_stride4 + off4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < vol * c8n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t c8 = idx / vol, v = idx - c8 * vol;
        float4* p0 = buf + (2 * c8) * vol + v;
        float4* p1 = p0 + vol;
        const float4 a = *p0, b = *p1;
        uint4 hi, lo;
        split_pair(a.x * sc, a.y * sc, hi.x, lo.x); split_pair(a.z * sc, a.w * sc, hi.y, lo.y);
        split_pair(b.x * sc, b.y * sc, hi.z, lo.z); split_pair(b.z * sc, b.w * sc, hi.w, lo.w);
        *reinterpret_cast<uint4*>(p0) = hi;
        *reinterpret_cast<uint4*>(p1) = lo;
    }
}
__global__ void __launch_bounds__(256)
split_to_c4(float4* __restrict__ slab, size_t slab_stride4, size_t off4, int c8n, size_t vol, int slot) {
    const float* hdr = reinterpret_cast<const float*>(slab + (size_t)blockIdx.y * slab_stride4);
    const float inv = 1.f / hdr[SCALE_SLOT0 + slot];
    float4* buf = slab + (size_t)blockIdx.y * slab_stride4 + off4;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < vol * c8n; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t c8 = idx / vol, v = idx - c8 * vol;
        float4* p0 = buf + (2 * c8) * vol + v;
        float4* p1 = p0 + vol;
        const uint4 hi = *reinterpret_cast<const uint4*>(p0), lo = *reinterpret_cast<const uint4*>(p1);
        const float2 a = unsplit_pair(hi.x, lo.x, inv), b = unsplit_pair(hi.y, lo.y, inv);
        const float2 c = unsplit_pair(hi.z, lo.z, inv), d = unsplit_pair(hi.w, lo.w, inv);
        *p0 = make_float4(a.x, a.y, b.x, b.y);
        *p1 = make_float4(c.x, c.y, d.x, d.y);
    }
}

static int grid_for(size_t work) {
    size_t b = (work + 255) / 256;
    if (b > 148 * 16) b = 148 * 16;
    if (b < 1) b = 1;
    return (int)b;
}

}  // namespace ct

using namespace ct;

// ---------------------------------------------------------------------------------------------
// plan + weights
// ---------------------------------------------------------------------------------------------
static void conv_list(const CtUNetSpec* sp, std::vector<std::pair<int, int>>& out) {
    int c = 1;
    std::vector<int> skips;
    for (int l = 0; l < sp->levels; ++l) {
        out.push_back({c, sp->down[l][0]});
        out.push_back({sp->down[l][0], sp->down[l][1]});
        skips.push_back(sp->down[l][1]);
        c = sp->down[l][1];
    }
    for (int u = 0; u < sp->levels; ++u) {
        out.push_back({c, sp->up[u][0]});
        out.push_back({sp->up[u][0], sp->up[u][1]});
        c = sp->up[u][1] + skips[sp->levels - 1 - u];
    }
    out.push_back({c, sp->out[0]});
    out.push_back({sp->out[0], sp->out[1]});
}

extern "C" size_t ct_unet_weight_count(const CtUNetSpec* sp) {
    std::vector<std::pair<int, int>> convs;
    conv_list(sp, convs);
    size_t n = 0;
    for (auto& c : convs) n += (size_t)27 * c.first * c.second + 5 * (size_t)c.second;
    return n + sp->out[1] + 1;
}

static int validate_spec(const CtUNetSpec* sp) {
    CT_REQUIRE(sp->levels >= 1 && sp->levels <= CT_UNET_MAX_LEVELS, "unet: levels %d out of range", sp->levels);
    int dx = sp->in_x, dy = sp->in_y, dz = sp->in_z;
    for (int l = 0; l < sp->levels; ++l) {
        CT_REQUIRE(dx % sp->pool_x == 0 && dy % sp->pool_y == 0 && dz % sp->pool_z == 0,
                   "unet: input %dx%dx%d not divisible by the pool size at level %d", sp->in_x, sp->in_y, sp->in_z, l);
        dx /= sp->pool_x; dy /= sp->pool_y; dz /= sp->pool_z;
        for (int q = 0; q < 2; ++q) {
            CT_REQUIRE(sp->down[l][q] % 8 == 0 && sp->up[l][q] % 8 == 0, "unet: filter counts must be multiples of 8");
        }
    }
    CT_REQUIRE(sp->out[0] % 8 == 0 && sp->out[1] % 4 == 0, "unet: output filter counts must be multiples of 8 / 4");
    return 0;
}

extern "C" int ct_unet_create(const CtUNetSpec* sp, const float* w, size_t n_floats, CtUNet** out) {
    CT_REQUIRE(sp && w && out, "ct_unet_create: null argument");
    if (validate_spec(sp)) return 1;
    CT_REQUIRE(n_floats == ct_unet_weight_count(sp), "ct_unet_create: expected %zu weights, got %zu",
               ct_unet_weight_count(sp), n_floats);
    CtUNet* net = new CtUNet();
    net->spec = *sp;
    net->alpha = sp->act_relu ? 0.f : 0.3f;     // keras LeakyReLU() default alpha
    net->engine = 0;
    std::vector<std::pair<int, int>> convs;
    conv_list(sp, convs);

    // ---- pack weights on the host into one staging vector, then one device allocation
    std::vector<float> host;
    struct Offs { size_t wd, wt, wx, b, sc, sh, wu, ws, wz, wzs; int c_up; float inv_u, inv_s, inv_z, inv_zs, bp, bq; };
    std::vector<Offs> offs;
    std::vector<float> inv_scales;
    // decoder blocks that read concatenate([up, skip]): layer index -> channels of the up-sampled half
    std::vector<int> c_up_of(convs.size(), 0);
    if (sp->pool_x == 2 && sp->pool_y == 2 && sp->pool_z == 1)
        for (int u = 0; u < sp->levels; ++u)
            c_up_of[2 * sp->levels + 2 * (u + 1)] = sp->up[u][1];          // next up block's first conv, or the output conv
    const float* p = w;
    const float eps = 1e-3f;                    // keras BatchNormalization default epsilon
    for (auto& c : convs) {
        const int cin = c.first, cout = c.second, cin_pad = (cin + 3) / 4 * 4;
        Offs o;
        auto take = [&](size_t n) { size_t at = (host.size() + 63) / 64 * 64; host.resize(at + n, 0.f); return at; };
        o.wd = take((size_t)cin_pad * 27 * cout);
        // keras kernel (kx,ky,kz,ci,co) -> [ci/4][tap][co][ci%4]
        for (int tap = 0; tap < 27; ++tap)
            for (int ci = 0; ci < cin; ++ci)
                for (int co = 0; co < cout; ++co)
                    host[o.wd + (((size_t)(ci / 4) * 27 + tap) * cout + co) * 4 + (ci % 4)] =
                        p[((size_t)tap * cin + ci) * cout + co];
        const size_t tcn = tc_weight_floats(cin_pad, cout);
        o.wt = take(tcn);
        inv_scales.push_back(tcn ? tc_pack_weights(p, cin, cin_pad, cout, &host[o.wt]) : 1.f);
        const size_t txn = tcx_weight_floats(cin_pad, cout);
        o.wx = take(txn);
        if (txn) tcx_pack_weights(p, cin, cin_pad, cout, &host[o.wx]);       // same scale as the classic image
        const size_t tzn = tcz_weight_floats(cin, cout);                      // plane-walk image (cin % 8 == 0, Cout 8/16/32)
        o.wz = take(tzn); o.wzs = 0; o.inv_z = o.inv_zs = 1.f;
        if (tzn) o.inv_z = tcz_pack_weights_range(p, cin, 0, cin, cout, &host[o.wz]);
        o.c_up = 0; o.wu = o.ws = 0; o.inv_u = o.inv_s = 1.f;
        const int cu = c_up_of[offs.size()];
        if (cu > 0 && cu % 8 == 0 && (cin - cu) % 8 == 0 && tcu_weight_floats(cu, cout) && tcx_weight_floats(cin - cu, cout)) {
            o.c_up = cu;
            o.wu = take(tcu_weight_floats(cu, cout));
            o.inv_u = tcu_pack_weights(p, cin, cu, cout, &host[o.wu]);
            o.ws = take(tcx_weight_floats(cin - cu, cout));
            o.inv_s = tcx_pack_weights_range(p, cin, cu, cin - cu, cout, &host[o.ws]);
            if (tcz_weight_floats(cin - cu, cout)) {
                o.wzs = take(tcz_weight_floats(cin - cu, cout));
                o.inv_zs = tcz_pack_weights_range(p, cin, cu, cin - cu, cout, &host[o.wzs]);
            }
        }
        std::vector<double> w1(cout, 0.0);                                   // sum of |w| per output channel
        for (size_t i = 0; i < (size_t)27 * cin; ++i)
            for (int co = 0; co < cout; ++co) w1[co] += std::fabs((double)p[i * cout + co]);
        p += (size_t)27 * cin * cout;
        o.b = take(cout); o.sc = take(cout); o.sh = take(cout);
        const float *bias = p, *gamma = p + cout, *beta = p + 2 * cout, *mean = p + 3 * cout, *var = p + 4 * cout;
        double bp = 0.0, bq = 0.0;
        for (int co = 0; co < cout; ++co) {
            const float s = gamma[co] / std::sqrt(var[co] + eps);
            host[o.b + co] = bias[co];
            host[o.sc + co] = s;
            host[o.sh + co] = beta[co] - mean[co] * s;
            bp = std::fmax(bp, std::fabs((double)s) * w1[co]);
            bq = std::fmax(bq, std::fabs((double)s) * std::fabs((double)bias[co]) + std::fabs((double)host[o.sh + co]));
        }
        o.bp = (float)bp; o.bq = (float)bq;
        p += 5 * (size_t)cout;
        offs.push_back(o);
    }
    net->last_c = sp->out[1];
    size_t head_at = (host.size() + 63) / 64 * 64;
    host.resize(head_at + net->last_c, 0.f);
    for (int c = 0; c < net->last_c; ++c) host[head_at + c] = p[c];
    net->head_b = p[net->last_c];

    if (check_cuda(cudaMalloc(&net->all_dev, host.size() * sizeof(float)), "cudaMalloc(weights)") ||
        check_cuda(cudaMemcpy(net->all_dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice),
                   "cudaMemcpy(weights)")) {
        delete net;
        return 1;
    }
    for (size_t i = 0; i < convs.size(); ++i) {
        ConvLayer L;
        L.cin = convs[i].first; L.cout = convs[i].second; L.cin_pad = (L.cin + 3) / 4 * 4;
        L.w_direct = net->all_dev + offs[i].wd;
        L.w_tc = tc_weight_floats(L.cin_pad, L.cout) ? net->all_dev + offs[i].wt : nullptr;
        L.w_tcx = tcx_weight_floats(L.cin_pad, L.cout) ? net->all_dev + offs[i].wx : nullptr;
        L.w_tc_inv_scale = inv_scales[i];
        L.c_up = offs[i].c_up;
        L.w_tcu = offs[i].c_up ? net->all_dev + offs[i].wu : nullptr;
        L.w_tcx_skip = offs[i].c_up ? net->all_dev + offs[i].ws : nullptr;
        L.w_tcu_inv_scale = offs[i].inv_u; L.w_tcx_skip_inv_scale = offs[i].inv_s;
        L.w_tcz = tcz_weight_floats(L.cin, L.cout) ? net->all_dev + offs[i].wz : nullptr;
        L.w_tcz_inv_scale = offs[i].inv_z;
        L.w_tcz_skip = offs[i].wzs ? net->all_dev + offs[i].wzs : nullptr;
        L.w_tcz_skip_inv_scale = offs[i].inv_zs;
        L.bound_p = offs[i].bp; L.bound_q = offs[i].bq;
        L.bias = net->all_dev + offs[i].b; L.scale = net->all_dev + offs[i].sc; L.shift = net->all_dev + offs[i].sh;
        net->layers.push_back(L);
    }
    net->head_w = net->all_dev + head_at;
    // can the activation buffers between the blocks be split-fp16 (see use_split / run_plan)?
    net->split_ok = sp->pool_x == 2 && sp->pool_y == 2 && sp->pool_z == 1 && sp->in_z % 8 == 0 &&
                    net->layers[0].cin == 1 && net->layers[0].cout == 8;
    for (size_t i = 1; i < net->layers.size() && net->split_ok; ++i) {
        const ConvLayer& L = net->layers[i];
        const bool decoder_first = c_up_of[i] > 0;
        net->split_ok = L.cin % 8 == 0 && ((L.cout <= 32 && L.w_tcx) || (L.cout == 64 && L.w_tc)) &&
                        (!decoder_first || (L.c_up > 0 && L.w_tcu && L.w_tcx_skip && L.cout <= 32));
    }
    for (int u = 0; u < sp->levels && net->split_ok; ++u)        // every up-sampling must be absorbed by a phase kernel
        net->split_ok = c_up_of[2 * sp->levels + 2 * (u + 1)] > 0;

    // ---- op plan over one tile's slab
    int lx[CT_UNET_MAX_LEVELS + 1], ly[CT_UNET_MAX_LEVELS + 1], lz[CT_UNET_MAX_LEVELS + 1];
    lx[0] = sp->in_x; ly[0] = sp->in_y; lz[0] = sp->in_z;
    for (int l = 1; l <= sp->levels; ++l) { lx[l] = lx[l - 1] / sp->pool_x; ly[l] = ly[l - 1] / sp->pool_y; lz[l] = lz[l - 1] / sp->pool_z; }
    size_t off = HDR_FLOATS;                       // slab header: one max|value| slot and one operand-scale slot per buffer
    int n_slots = 0;
    std::vector<std::pair<size_t, int>> slot_of;  // buffer offset -> slot
    auto buf = [&](int c, int l) {
        size_t at = off; off += (size_t)c * lx[l] * ly[l] * lz[l]; off = (off + 63) / 64 * 64;
        slot_of.push_back({at, n_slots++});
        return at;
    };
    auto slot = [&](size_t at) { for (auto& q : slot_of) if (q.first == at) return q.second; return -1; };
    double flops = 0;
    int li = 0;
    auto conv = [&](size_t s_off, int s_c, size_t d_off, int d_c, int d_coff, int l) {
        Op o{}; o.kind = OP_CONV; o.layer = li; o.src_off = s_off; o.dst_off = d_off; o.src_c = s_c; o.dst_c = d_c;
        o.src_coff = 0; o.dst_coff = d_coff; o.c = net->layers[li].cout;
        o.src_slot = slot(s_off); o.dst_slot = slot(d_off);
        o.sx = o.dx = lx[l]; o.sy = o.dy = ly[l]; o.sz = o.dz = lz[l];
        flops += 2.0 * 27 * net->layers[li].cin * net->layers[li].cout * (double)lx[l] * ly[l] * lz[l];
        net->ops.push_back(o); ++li;
    };
    net->in_off = buf(4, 0);
    size_t x_off = net->in_off; int x_c = 4;
    size_t cat_off[CT_UNET_MAX_LEVELS]; int cat_c[CT_UNET_MAX_LEVELS];
    for (int l = 0; l < sp->levels; ++l) {
        const int f1 = sp->down[l][0], f2 = sp->down[l][1];
        const int upc = sp->up[sp->levels - 1 - l][1];
        size_t t = buf(f1, l);
        conv(x_off, x_c, t, f1, 0, l);
        cat_c[l] = upc + f2; cat_off[l] = buf(cat_c[l], l);
        conv(t, f1, cat_off[l], cat_c[l], upc, l);
        size_t pl = buf(f2, l + 1);
        Op o{}; o.kind = OP_POOL; o.src_off = cat_off[l]; o.dst_off = pl; o.src_c = cat_c[l]; o.dst_c = f2;
        o.src_coff = upc; o.dst_coff = 0; o.c = f2;
        o.src_slot = slot(cat_off[l]); o.dst_slot = slot(pl);
        o.sx = lx[l]; o.sy = ly[l]; o.sz = lz[l]; o.dx = lx[l + 1]; o.dy = ly[l + 1]; o.dz = lz[l + 1];
        net->ops.push_back(o);
        x_off = pl; x_c = f2;
    }
    for (int u = 0; u < sp->levels; ++u) {
        const int l = sp->levels - 1 - u;      // level the result is upsampled INTO; convs run at l+1
        const int f1 = sp->up[u][0], f2 = sp->up[u][1];
        size_t t1 = buf(f1, l + 1), t2 = buf(f2, l + 1);
        conv(x_off, x_c, t1, f1, 0, l + 1);
        conv(t1, f1, t2, f2, 0, l + 1);
        Op o{}; o.kind = OP_UPSAMPLE; o.src_off = t2; o.dst_off = cat_off[l]; o.src_c = f2; o.dst_c = cat_c[l];
        o.src_coff = 0; o.dst_coff = 0; o.c = f2;
        o.src_slot = slot(t2); o.dst_slot = slot(cat_off[l]);
        o.sx = lx[l + 1]; o.sy = ly[l + 1]; o.sz = lz[l + 1]; o.dx = lx[l]; o.dy = ly[l]; o.dz = lz[l];
        net->ops.push_back(o);
        x_off = cat_off[l]; x_c = cat_c[l];
    }
    size_t o1 = buf(sp->out[0], 0), o2 = buf(sp->out[1], 0);
    conv(x_off, x_c, o1, sp->out[0], 0, 0);
    conv(o1, sp->out[0], o2, sp->out[1], 0, 0);
    flops += 2.0 * sp->out[1] * (double)lx[0] * ly[0] * lz[0];
    net->last_off = o2;
    net->slab_floats = off;
    net->flops_per_tile = flops;
    net->in_slot = slot(net->in_off);
    if (n_slots > AMAX_SLOTS) { set_error("unet: %d buffers exceed the %d header slots", n_slots, AMAX_SLOTS); delete net; return 1; }
    *out = net;
    return 0;
}

extern "C" void ct_unet_destroy(CtUNet* net) {
    if (!net) return;
    cudaFree(net->all_dev);
    delete net;
}

extern "C" int ct_unet_set_engine(CtUNet* net, int engine) {
    CT_REQUIRE(net && engine >= 0 && engine <= 4, "ct_unet_set_engine: bad argument");
    net->engine = engine;
    return 0;
}

extern "C" double ct_unet_flops_per_tile(const CtUNet* net) { return net ? net->flops_per_tile : 0.0; }

extern "C" size_t ct_unet_workspace_bytes(const CtUNet* net, int tiles_per_batch) {
    if (!net || tiles_per_batch < 1) return 0;
    return net->slab_floats * sizeof(float) * (size_t)tiles_per_batch + 256;
}

static int tile_geometry(const CtUNet* net, int x, int y, int z, const int shrink[3], int centre[3], int num[3]) {
    const int in[3] = {net->spec.in_x, net->spec.in_y, net->spec.in_z};
    const int sz[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        centre[a] = in[a] - 2 * shrink[a];
        CT_REQUIRE(shrink[a] >= 0 && centre[a] > 0, "unet3_prediction: shrink %d too large for input size %d", shrink[a], in[a]);
        CT_REQUIRE(sz[a] > 0, "unet3_prediction: empty volume");
        num[a] = (sz[a] + centre[a] - 1) / centre[a];
    }
    return 0;
}

extern "C" int ct_unet_tile_count(const CtUNet* net, int x, int y, int z, const int shrink[3], int counts_out[3]) {
    int centre[3], num[3];
    if (!net || tile_geometry(net, x, y, z, shrink, centre, num)) return -1;
    if (counts_out) { counts_out[0] = num[0]; counts_out[1] = num[1]; counts_out[2] = num[2]; }
    return num[0] * num[1] * num[2];
}

static TileGeom make_geom(const CtUNet* net, int x, int y, int z, const int centre[3], const int shrink[3],
                          const int tlo[3], const int tn[3], const int in_lo[3], const int in_dim[3],
                          const int out_lo[3], const int out_dim[3]) {
    TileGeom g{};
    g.X = x; g.Y = y; g.Z = z;
    g.TX = net->spec.in_x; g.TY = net->spec.in_y; g.TZ = net->spec.in_z;
    for (int a = 0; a < 3; ++a) {
        g.tlo[a] = tlo[a]; g.tn[a] = tn[a]; g.c[a] = centre[a]; g.b[a] = shrink[a];
        g.in_lo[a] = in_lo[a]; g.in_dim[a] = in_dim[a]; g.out_lo[a] = out_lo[a]; g.out_dim[a] = out_dim[a];
    }
    return g;
}

// One convolution block on the engine the network is set to (see CtUNet::engine).
static int launch_conv(const CtUNet* net, const Op& op, float* slab0, size_t stride, int tiles, cudaStream_t s,
                       const Op* next = nullptr, bool* next_fused = nullptr, int fmt = 0) {
    int rc = 2;
    if (next_fused) *next_fused = false;
    // auto: the plane-walk kernel wherever it applies (split-fp16 source, z = 16, Cout <= 32), then the x-stacked one
    if (net->engine == 0 || net->engine == 5) rc = launch_conv_tcz(net, op, slab0, stride, tiles, s, next, next_fused, fmt);
    if (rc == 2 && net->engine != 1 && net->engine != 3) rc = launch_conv_tcx(net, op, slab0, stride, tiles, s, next, next_fused, fmt);
    if (rc == 2 && net->engine != 1) rc = launch_conv_tc(net, op, slab0, stride, tiles, s, fmt);
    if (rc == 1) return 1;
    CT_REQUIRE(rc == 0 || fmt == 0, "unet: layer %d has no tensor-core kernel for split-fp16 buffers", op.layer);
    if (rc == 2) {
        CT_REQUIRE(net->engine == 0 || net->engine == 1 || net->engine == 5,
                   "unet: tcgen05 engine forced but layer %d (cin %d, cout %d, z %d) is unsupported",
                   op.layer, net->layers[op.layer].cin, net->layers[op.layer].cout, op.sz);
        if (launch_conv_direct(net, op, slab0, stride, tiles, s)) return 1;
    }
    return 0;
}

// Runs the op plan on `tiles` slabs that already hold their padded input (or, with skip_first, the output of the
// first convolution block).
// Split-fp16 activation buffers (unet_common.cuh) are used when every block after the first runs on a tensor-core
// kernel that reads and writes them and the engine is one of the tensor-core mixes.
static bool use_split(const CtUNet* net) { return net->split_ok && (net->engine == 0 || net->engine == 2 || net->engine == 4); }

static int run_plan(const CtUNet* net, float* slab0, int tiles, cudaStream_t s, bool skip_first = false) {
    const size_t stride = net->slab_floats;
    // the fused first block writes fp32; every later activation buffer is split-fp16 except the last block's output,
    // which the 1x1x1 head reads as fp32
    const bool split = skip_first && use_split(net);
    const int last_layer = (int)net->layers.size() - 1;
    for (size_t idx = skip_first ? 1 : 0; idx < net->ops.size(); ++idx) {
        const Op& op = net->ops[idx];
        if (op.kind == OP_CONV) {
            bool fused = false;
            const Op* next = idx + 1 < net->ops.size() ? &net->ops[idx + 1] : nullptr;
            const int fmt = !split ? 0 : (op.layer == 1 ? 0 : FMT_SRC_SPLIT) | (op.layer == last_layer ? 0 : FMT_DST_SPLIT);
            if (launch_conv(net, op, slab0, stride, tiles, s, next, &fused, fmt)) return 1;
            if (fused) ++idx;                      // the pooling that followed was written by the block's epilogue
        } else {
            const float4* src = reinterpret_cast<const float4*>(slab0);
            float4* dst = reinterpret_cast<float4*>(slab0);
            if (op.kind == OP_POOL && split) {
                dim3 grid((op.c / 8) * op.dx, tiles);
                pool_split_kernel<<<grid, 256, 0, s>>>(reinterpret_cast<const uint4*>(slab0), reinterpret_cast<uint4*>(slab0), stride / 4,
                                                       op.src_off / 4, op.dst_off / 4, op.src_coff / 4, op.c / 8, op.sx, op.sy,
                                                       op.sz, op.dx, op.dy, op.dz, net->spec.pool_x, net->spec.pool_y,
                                                       net->spec.pool_z, op.src_slot, op.dst_slot);
                CT_LAUNCHED("pool_split_kernel");
            } else if (op.kind == OP_POOL) {
                dim3 grid((op.c / 4) * op.dx, tiles);
                pool_kernel<<<grid, 256, 0, s>>>(src, dst, stride / 4, op.src_off / 4, op.dst_off / 4, op.src_coff / 4,
                                                 op.c / 4, op.sx, op.sy, op.sz, op.dx, op.dy, op.dz,
                                                 net->spec.pool_x, net->spec.pool_y, net->spec.pool_z, op.src_slot, op.dst_slot);
                CT_LAUNCHED("pool_kernel");
            } else {
                // UpSampling3D + concatenate + conv block (unet3d.py:96-98) without the up-sampled tensor: the phase
                // kernel convolves the low-resolution source into partial sums, the x-stacked kernel adds the skip half
                const Op* nx = idx + 1 < net->ops.size() ? &net->ops[idx + 1] : nullptr;
                if (nx && nx->kind == OP_CONV && nx->src_off == op.dst_off && (net->engine == 0 || net->engine == 2 || net->engine == 4) &&
                    net->layers[nx->layer].c_up == op.c && op.dst_coff == 0 && op.dx == 2 * op.sx && op.dy == 2 * op.sy && op.dz == op.sz) {
                    const ConvLayer& NL = net->layers[nx->layer];
                    // auto: the skip half runs on the plane-walk kernel, which wants the partial sums in its P8 layout
                    const bool zskip = net->engine == 0 && split && tcz_takes_skip(NL) && nx->sz == 16 && nx->layer != last_layer &&
                                       nx->dst_coff % 8 == 0;
                    const int rc = launch_conv_tcu(net, NL, slab0, stride, tiles, op.src_off, op.src_slot,
                                                   op.sx, op.sy, op.sz, nx->dst_off, nx->dst_coff, s, split, nx->src_slot, zskip);
                    if (rc == 1) return 1;
                    CT_REQUIRE(rc == 0 || !split, "unet: phase kernel refused layer %d between split-fp16 buffers", nx->layer);
                    if (rc == 0) {
                        const int fmt = !split ? 0 : FMT_SRC_SPLIT | (nx->layer == last_layer ? 0 : FMT_DST_SPLIT);
                        const int rc2 = zskip ? launch_conv_tcz_skip(net, *nx, slab0, stride, tiles, s, fmt, op.src_slot)
                                              : launch_conv_tcx_skip(net, *nx, slab0, stride, tiles, s, fmt, split ? op.src_slot : -1);
                        CT_REQUIRE(rc2 == 0, "unet: skip-half convolution of layer %d failed", nx->layer);
                        ++idx;
                        continue;
                    }
                }
                CT_REQUIRE(!split, "unet: up-sampling op %zu has no phase kernel between split-fp16 buffers", idx);
                dim3 grid((op.c / 4) * op.sx, tiles);
                upsample_kernel<<<grid, 256, 0, s>>>(src, dst, stride / 4, op.src_off / 4, op.dst_off / 4, op.dst_coff / 4,
                                                     op.c / 4, op.sx, op.sy, op.sz, op.dx, op.dy, op.dz,
                                                     net->spec.pool_x, net->spec.pool_y, net->spec.pool_z, op.src_slot, op.dst_slot);
                CT_LAUNCHED("upsample_kernel");
            }
        }
    }
    return 0;
}

static int run_tiles(const CtUNet* net, const float* src, float* prob, int mode, int first, int last,
                     const TileGeom& geo, void* ws, size_t ws_bytes, int tiles_per_batch, cudaStream_t s) {
    CT_REQUIRE(tiles_per_batch >= 1, "unet: tiles_per_batch must be >= 1");
    CT_REQUIRE(ws_bytes >= ct_unet_workspace_bytes(net, tiles_per_batch), "unet: workspace too small (%zu < %zu)",
               ws_bytes, ct_unet_workspace_bytes(net, tiles_per_batch));
    CT_REQUIRE(((uintptr_t)ws & 255) == 0, "unet: workspace must be 256-byte aligned");
    float* slab0 = static_cast<float*>(ws);
    const int TX = net->spec.in_x, TY = net->spec.in_y, TZ = net->spec.in_z;
    for (int t0 = first; t0 < last; t0 += tiles_per_batch) {
        const int nt = (last - t0 < tiles_per_batch) ? last - t0 : tiles_per_batch;
        dim3 g(TX, nt), gh(mode == 0 ? geo.c[0] : TX, nt);
        CT_CUDA(cudaMemset2DAsync(slab0, net->slab_floats * sizeof(float), 0, HDR_FLOATS * sizeof(float), nt, s));
        // engines 0 / 2 / 4: the first block (Cin = 1) reads the volume itself; otherwise gather, then the generic engines
        int fused = 2;
        if (net->engine != 1 && net->engine != 3)
            fused = launch_first_conv_fused(net, net->ops[0], src, mode, t0, geo, slab0, net->slab_floats, nt, s);
        if (fused == 1) return 1;
        if (fused == 2) {
            gather_tiles<<<g, 256, 0, s>>>(src, reinterpret_cast<float4*>(slab0), net->slab_floats / 4, net->in_off / 4, mode, t0,
                                           geo, net->in_slot);
            CT_LAUNCHED("gather_tiles");
        }
        if (run_plan(net, slab0, nt, s, fused == 0)) return 1;
        head_scatter<<<gh, 256, 0, s>>>(reinterpret_cast<const float4*>(slab0), net->slab_floats / 4, net->last_off / 4,
                                       net->last_c / 4, net->head_w, net->head_b, prob, mode, t0, geo);
        CT_LAUNCHED("head_scatter");
    }
    return 0;
}

extern "C" size_t ct_unet_conv_block_workspace_bytes(const CtUNet* net, int layer, int batch, int x, int y, int z) {
    if (!net || layer < 0 || layer >= (int)net->layers.size() || batch < 1) return 0;
    const ConvLayer& L = net->layers[layer];
    const size_t vol = (size_t)x * y * z;
    return (((size_t)L.cin_pad + (size_t)L.cout) * vol + HDR_FLOATS) * sizeof(float) * (size_t)batch + 512;
}

extern "C" int ct_unet_conv_block(const CtUNet* net, int layer, int engine, const float* in, float* out, int batch,
                                  int x, int y, int z, void* ws, size_t ws_bytes, void* stream) {
    CT_REQUIRE(net && in && out && ws, "ct_unet_conv_block: null argument");
    CT_REQUIRE(layer >= 0 && layer < (int)net->layers.size(), "ct_unet_conv_block: layer %d out of range", layer);
    CT_REQUIRE(batch >= 1 && x > 0 && y > 0 && z > 0, "ct_unet_conv_block: bad shape");
    CT_REQUIRE(engine >= 1 && engine <= 9, "ct_unet_conv_block: engine must be 1 (direct), 2 (tcgen05), 3 (tcgen05 classic), 4 (tcgen05 "
               "x-stacked), 5-7 (tcgen05 on split-fp16 buffers: both / destination only / source only) or 8-9 (the plane-walk "
               "kernel wherever it can run, on split-fp16 buffers: both / source only)");
    // engines 5-9: the block as it runs inside the network, between split-fp16 activation buffers (unet_common.cuh)
    const int fmt = (engine == 5 || engine == 8) ? (FMT_SRC_SPLIT | FMT_DST_SPLIT) : engine == 6 ? FMT_DST_SPLIT
                    : (engine == 7 || engine == 9) ? FMT_SRC_SPLIT : 0;
    if (engine >= 8) engine = 5;            // internal: the plane-walk kernel wherever it can run, else the tcgen05 mix
    else if (fmt) engine = 2;
    CT_REQUIRE(ws_bytes >= ct_unet_conv_block_workspace_bytes(net, layer, batch, x, y, z), "ct_unet_conv_block: workspace too small");
    CT_REQUIRE(((uintptr_t)ws & 255) == 0, "ct_unet_conv_block: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    const ConvLayer& L = net->layers[layer];
    const size_t vol = (size_t)x * y * z;
    // one "slab" per batch entry: [cin_pad planes | cout planes]
    const size_t stride = ((size_t)L.cin_pad + L.cout) * vol + HDR_FLOATS;
    float* slab0 = static_cast<float*>(ws);
    CT_CUDA(cudaMemset2DAsync(slab0, stride * sizeof(float), 0, HDR_FLOATS * sizeof(float), batch, s));
    Op op{};
    op.kind = OP_CONV; op.layer = layer; op.src_off = HDR_FLOATS; op.dst_off = HDR_FLOATS + (size_t)L.cin_pad * vol;
    op.src_slot = 0; op.dst_slot = 1;
    op.src_c = L.cin_pad; op.dst_c = L.cout; op.src_coff = 0; op.dst_coff = 0; op.c = L.cout;
    op.sx = op.dx = x; op.sy = op.dy = y; op.sz = op.dz = z;
    const int cin4 = L.cin_pad / 4, cout4 = L.cout / 4;
    for (int b = 0; b < batch; ++b) {
        const size_t total = vol * cin4;
        ndhwc_to_c4<<<grid_for(total), 256, 0, s>>>(in + (size_t)b * vol * L.cin,
                                                     reinterpret_cast<float4*>(slab0 + b * stride + op.src_off),
                                                     vol, L.cin, cin4, total, slab0 + b * stride + op.src_slot);
        CT_LAUNCHED("ndhwc_to_c4");
    }
    {
        CtUNet view = *net;                     // shallow copy: same device arrays, engine chosen for this call
        view.engine = engine;
        float4* slab4 = reinterpret_cast<float4*>(slab0);
        if (fmt & FMT_SRC_SPLIT) {
            CT_REQUIRE(L.cin_pad % 8 == 0, "ct_unet_conv_block: a split-fp16 source needs a multiple of 8 input channels");
            c4_to_split<<<dim3(grid_for(vol * (L.cin_pad / 8)), batch), 256, 0, s>>>(slab4, stride / 4, op.src_off / 4, L.cin_pad / 8, vol, op.src_slot);
            CT_LAUNCHED("c4_to_split");
        }
        if (launch_conv(&view, op, slab0, stride, batch, s, nullptr, nullptr, fmt)) return 1;
        if (fmt & FMT_DST_SPLIT) {
            split_to_c4<<<dim3(grid_for(vol * (L.cout / 8)), batch), 256, 0, s>>>(slab4, stride / 4, op.dst_off / 4, L.cout / 8, vol, op.dst_slot);
            CT_LAUNCHED("split_to_c4");
        }
    }
    for (int b = 0; b < batch; ++b) {
        const size_t total = vol * cout4;
        c4_to_ndhwc<<<grid_for(total), 256, 0, s>>>(reinterpret_cast<const float4*>(slab0 + b * stride + op.dst_off),
                                                     out + (size_t)b * vol * L.cout, vol, L.cout, cout4, total);
        CT_LAUNCHED("c4_to_ndhwc");
    }
    return 0;
}

extern "C" int ct_unet_predict_tiles(const CtUNet* net, const float* tiles, float* prob, int batch,
                                     void* ws, size_t ws_bytes, int tiles_per_batch, void* stream) {
    CT_REQUIRE(net && tiles && prob && batch >= 0, "ct_unet_predict_tiles: bad argument");
    TileGeom g{};
    g.TX = net->spec.in_x; g.TY = net->spec.in_y; g.TZ = net->spec.in_z;
    g.tn[0] = batch; g.tn[1] = g.tn[2] = 1;
    for (int a = 0; a < 3; ++a) g.c[a] = 1;
    return run_tiles(net, tiles, prob, 1, 0, batch, g, ws, ws_bytes, tiles_per_batch, (cudaStream_t)stream);
}

extern "C" int ct_unet3_prediction(const CtUNet* net, const float* vol_norm, float* prob, int x, int y, int z,
                                   const int shrink[3], int tile_begin, int tile_end,
                                   void* ws, size_t ws_bytes, int tiles_per_batch, void* stream) {
    CT_REQUIRE(net && vol_norm && prob && shrink, "ct_unet3_prediction: null argument");
    int centre[3], num[3];
    if (tile_geometry(net, x, y, z, shrink, centre, num)) return 1;
    const int total = num[0] * num[1] * num[2];
    CT_REQUIRE(tile_begin >= 0 && tile_begin <= tile_end && tile_end <= total,
               "ct_unet3_prediction: tile range [%d,%d) outside [0,%d)", tile_begin, tile_end, total);
    const int zero[3] = {0, 0, 0}, dim[3] = {x, y, z};
    const TileGeom g = make_geom(net, x, y, z, centre, shrink, zero, num, zero, dim, zero, dim);
    return run_tiles(net, vol_norm, prob, 0, tile_begin, tile_end, g, ws, ws_bytes, tiles_per_batch, (cudaStream_t)stream);
}

// numpy.pad(mode='reflect') index on the host (same triangle wave as the device function)
static int reflect_host(int j, int n) {
    if (n == 1) return 0;
    const int period = 2 * (n - 1);
    j %= period;
    if (j < 0) j += period;
    return j < n ? j : period - j;
}

extern "C" int ct_unet3_prediction_block(const CtUNet* net, const float* vol_block, const int in_lo[3], const int in_dim[3],
                                         float* prob_block, const int out_lo[3], const int out_dim[3], int x, int y, int z,
                                         const int shrink[3], const int tile_lo[3], const int tile_hi[3],
                                         void* ws, size_t ws_bytes, int tiles_per_batch, void* stream) {
    CT_REQUIRE(net && vol_block && prob_block && in_lo && in_dim && out_lo && out_dim && shrink && tile_lo && tile_hi,
               "ct_unet3_prediction_block: null argument");
    int centre[3], num[3], tn[3];
    if (tile_geometry(net, x, y, z, shrink, centre, num)) return 1;
    const int in[3] = {net->spec.in_x, net->spec.in_y, net->spec.in_z};
    const int vol[3] = {x, y, z};
    for (int a = 0; a < 3; ++a) {
        CT_REQUIRE(tile_lo[a] >= 0 && tile_lo[a] <= tile_hi[a] && tile_hi[a] <= num[a],
                   "ct_unet3_prediction_block: tile range [%d,%d) outside [0,%d) on axis %d", tile_lo[a], tile_hi[a], num[a], a);
        tn[a] = tile_hi[a] - tile_lo[a];
        CT_REQUIRE(in_dim[a] > 0 && out_dim[a] > 0 && in_lo[a] >= 0 && in_lo[a] + in_dim[a] <= vol[a],
                   "ct_unet3_prediction_block: input box [%d,%d) outside the volume on axis %d", in_lo[a], in_lo[a] + in_dim[a], a);
        // every coordinate the tiles read on this axis, after reflection, must be inside the input box
        for (int j = tile_lo[a] * centre[a] - shrink[a]; j < (tile_hi[a] - 1) * centre[a] - shrink[a] + in[a] && tn[a] > 0; ++j) {
            const int r = reflect_host(j, vol[a]);
            CT_REQUIRE(r >= in_lo[a] && r < in_lo[a] + in_dim[a],
                       "ct_unet3_prediction_block: tiles read voxel %d on axis %d, outside the input box [%d,%d)", r, a,
                       in_lo[a], in_lo[a] + in_dim[a]);
        }
    }
    const int total = tn[0] * tn[1] * tn[2];
    if (total == 0) return 0;
    const TileGeom g = make_geom(net, x, y, z, centre, shrink, tile_lo, tn, in_lo, in_dim, out_lo, out_dim);
    return run_tiles(net, vol_block, prob_block, 0, 0, total, g, ws, ws_bytes, tiles_per_batch, (cudaStream_t)stream);
}
