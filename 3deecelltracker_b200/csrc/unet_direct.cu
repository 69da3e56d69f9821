// CUDA-core fp32 3x3x3 'same' convolution with fused bias -> LeakyReLU/ReLU -> BatchNorm(eval).
//
// Replaces one Conv3D + LeakyReLU + BatchNormalization block of unet3d.py:117-119 (or the ReLU variant
// :139-140) for a batch of independent tiles.  This is the exact-fp32 engine: it serves the layers the
// tensor-core path does not take (Cin = 1) and is the numerical cross-check for the 3xTF32 path.
//
// CTA = 256 threads = 2 (x groups of 4) x 8 (y) x 16 (z) -> output block 8 x 8 x 16 voxels, CO output
// channels.  Each thread owns 4 consecutive-x voxels x CO channels (register tile).  Per 4-channel input
// chunk the 10 x 10 x 18 halo block (zero filled outside the tile = Keras 'same' padding) and the
// 27 x CO x 4 weights are staged in shared memory; lanes run along z so every activation LDS.128 of a
// quarter warp touches 128 contiguous bytes, and weight reads are warp-uniform broadcasts.
#include "unet_common.cuh"
#include "tc_ptx.cuh"

namespace ct {

constexpr int BX = 8, BY = 8, BZ = 16;
constexpr int SX = BX + 2, SY = BY + 2, SZ = BZ + 2;

template <int CO>
__global__ void __launch_bounds__(256, 2)
conv3_direct_kernel(const float4* __restrict__ src, float4* __restrict__ dst,
                    const float4* __restrict__ wts,      // [cin4][27][cout][1] float4 (4 cin)
                    const float* __restrict__ bias, const float* __restrict__ scale,
                    const float* __restrict__ shift, float alpha,
                    int cin4, int cout, int X, int Y, int Z, int nbx, int nby,
                    size_t src_tile_stride4, size_t dst_tile_stride4, int dst_c4off, float* __restrict__ amax_hdr,
                    int dst_slot) {
    __shared__ float4 in_s[SX][SY][SZ];
    __shared__ float4 w_s[27][CO];

    const int tid = threadIdx.x;
    const int tz = tid & 15, ty = (tid >> 4) & 7, tx = (tid >> 7) * 4;
    int b = blockIdx.x;
    const int bxi = b % nbx; b /= nbx;
    const int byi = b % nby; b /= nby;
    const int bzi = b;
    const int x0 = bxi * BX, y0 = byi * BY, z0 = bzi * BZ;
    const int co0 = blockIdx.y * CO;
    const int tile = blockIdx.z;
    const size_t vol = (size_t)X * Y * Z;
    const float4* s_tile = src + (size_t)tile * src_tile_stride4;

    float acc[4][CO];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int c = 0; c < CO; ++c) acc[v][c] = 0.f;

    for (int ck = 0; ck < cin4; ++ck) {
        __syncthreads();
        const float4* s_ck = s_tile + (size_t)ck * vol;
        for (int i = tid; i < SX * SY * SZ; i += 256) {
            int sz = i % SZ, r = i / SZ;
            int sy = r % SY, sx = r / SY;
            int gx = x0 + sx - 1, gy = y0 + sy - 1, gz = z0 + sz - 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gx >= 0 && gx < X && gy >= 0 && gy < Y && gz >= 0 && gz < Z)
                v = s_ck[((size_t)gx * Y + gy) * Z + gz];
            in_s[sx][sy][sz] = v;
        }
        const float4* w_ck = wts + ((size_t)ck * 27) * cout;
        for (int i = tid; i < 27 * CO; i += 256) {
            int t = i / CO, c = i % CO;
            w_s[t][c] = w_ck[(size_t)t * cout + co0 + c];
        }
        __syncthreads();

#pragma unroll 1
        for (int dy = 0; dy < 3; ++dy) {
#pragma unroll 1
            for (int dz = 0; dz < 3; ++dz) {
                float4 a[6];
#pragma unroll
                for (int i = 0; i < 6; ++i) a[i] = in_s[tx + i][ty + dy][tz + dz];
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int tap = (dx * 3 + dy) * 3 + dz;
#pragma unroll
                    for (int c = 0; c < CO; ++c) {
                        const float4 w = w_s[tap][c];
#pragma unroll
                        for (int v = 0; v < 4; ++v) {
                            acc[v][c] = fmaf(a[v + dx].x, w.x, acc[v][c]);
                            acc[v][c] = fmaf(a[v + dx].y, w.y, acc[v][c]);
                            acc[v][c] = fmaf(a[v + dx].z, w.z, acc[v][c]);
                            acc[v][c] = fmaf(a[v + dx].w, w.w, acc[v][c]);
                        }
                    }
                }
            }
        }
    }

    // epilogue: bias -> activation -> BN affine, written as 16-byte channel chunks
    const int gy = y0 + ty, gz = z0 + tz;
    float amax = 0.f;
    if (gy < Y && gz < Z) {
        float4* d_tile = dst + (size_t)tile * dst_tile_stride4;
#pragma unroll
        for (int c4 = 0; c4 < CO / 4; ++c4) {
            float bs[4], sc[4], sh[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int co = co0 + c4 * 4 + j;
                bs[j] = bias[co]; sc[j] = scale[co]; sh[j] = shift[co];
            }
            float4* d_ck = d_tile + (size_t)(dst_c4off + (co0 >> 2) + c4) * vol;
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                int gx = x0 + tx + v;
                if (gx < X) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = acc[v][c4 * 4 + j] + bs[j];
                        t = t > 0.f ? t : alpha * t;
                        o[j] = fmaf(t, sc[j], sh[j]);
                    }
                    d_ck[((size_t)gx * Y + gy) * Z + gz] = make_float4(o[0], o[1], o[2], o[3]);
                    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o[0]), fabsf(o[1]))), fmaxf(fabsf(o[2]), fabsf(o[3])));
                }
            }
        }
    }
    // keep the per-tile max|value| bound of the destination buffer current (read by the tensor-core engine)
    amax = warp_max(amax);
    if ((tid & 31) == 0) amax_update(amax_hdr + (size_t)tile * dst_tile_stride4 * 4 + dst_slot, amax);
}

int launch_conv_direct(const CtUNet* net, const Op& op, float* slab0, size_t slab_stride, int tiles,
                       cudaStream_t s) {
    const ConvLayer& L = net->layers[op.layer];
    const int X = op.sx, Y = op.sy, Z = op.sz;
    const int nbx = cdiv(X, BX), nby = cdiv(Y, BY), nbz = cdiv(Z, BZ);
    const float4* src = reinterpret_cast<const float4*>(slab0 + op.src_off);
    float4* dst = reinterpret_cast<float4*>(slab0 + op.dst_off);
    CT_REQUIRE(slab_stride % 4 == 0 && op.src_off % 4 == 0 && op.dst_off % 4 == 0, "conv: misaligned slab");
    CT_REQUIRE(op.src_c == L.cin_pad, "conv: source buffer has %d channels, layer expects %d", op.src_c, L.cin_pad);
    dim3 grid(nbx * nby * nbz, 1, tiles);
    ProfScope prof(PROF_CONV, s);
    if (L.cout % 16 == 0) {
        grid.y = L.cout / 16;
        conv3_direct_kernel<16><<<grid, 256, 0, s>>>(src, dst, reinterpret_cast<const float4*>(L.w_direct), L.bias,
                                                    L.scale, L.shift, net->alpha, L.cin_pad / 4, L.cout, X, Y, Z,
                                                    nbx, nby, slab_stride / 4, slab_stride / 4, op.dst_coff / 4, slab0, op.dst_slot);
    } else {
        CT_REQUIRE(L.cout % 8 == 0, "conv: cout %d must be a multiple of 8", L.cout);
        grid.y = L.cout / 8;
        conv3_direct_kernel<8><<<grid, 256, 0, s>>>(src, dst, reinterpret_cast<const float4*>(L.w_direct), L.bias,
                                                   L.scale, L.shift, net->alpha, L.cin_pad / 4, L.cout, X, Y, Z,
                                                   nbx, nby, slab_stride / 4, slab_stride / 4, op.dst_coff / 4, slab0, op.dst_slot);
    }
    CT_LAUNCHED("conv3_direct_kernel");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// First block of the network: Cin = 1 -> Cout = 8, fused with the tile gather.
//
// With one input channel the tensor-core path wastes 7/8 of its K (channels are chunked by 8) and the generic
// CUDA-core kernel 3/4 of its FMAs (chunks of 4); and the padded 4-channel tile copy that gather_tiles writes exists
// only to feed them.  Here a CTA stages the haloed 10 x 18 x 18 single-channel block straight from the normalised
// volume (reflect padding of unet3d.py:235 resolved per voxel; zero outside the TILE = the Keras 'same' padding) and
// each thread computes 8 consecutive-x voxels x 8 output channels for its (y, z): one shared-memory read of an input
// value feeds 24 FMAs.  Exact fp32.
// ---------------------------------------------------------------------------------------------
constexpr int FX = 8, FY = 16, FZ = 16;

__global__ void __launch_bounds__(256, 2)
first_conv_kernel(const float* __restrict__ src, int mode, int tile_first, const TileGeom g,
                  float4* __restrict__ dst, const float4* __restrict__ wts,      // [27][8] float4, .x = the one input channel
                  const float* __restrict__ bias, const float* __restrict__ scale, const float* __restrict__ shift,
                  float alpha, int nbx, int nby, size_t dst_tile_stride4, int dst_c4off, float* __restrict__ amax_hdr,
                  int dst_slot) {
    __shared__ float in_s[FX + 2][FY + 2][FZ + 2];
    __shared__ float4 w_s[27][2];

    const int tid = threadIdx.x;
    const int tz = tid & 15, ty = tid >> 4;
    int b = blockIdx.x;
    const int bxi = b % nbx; b /= nbx;
    const int byi = b % nby; b /= nby;
    const int x0 = bxi * FX, y0 = byi * FY, z0 = b * FZ;
    const int tile = blockIdx.z;
    const int TX = g.TX, TY = g.TY, TZ = g.TZ;
    const size_t vol = (size_t)TX * TY * TZ;
    int ti, tj, tk;
    tile_ijk(g, tile_first + tile, ti, tj, tk);

    // source index of every haloed x / y / z coordinate of the block (reflect padding is separable), -1 = outside
    // the tile: 46 reflections per CTA instead of 3 per voxel
    __shared__ int ix_s[FX + 2], iy_s[FY + 2], iz_s[FZ + 2];
    if (tid < FX + 2) {
        const int l = x0 + tid - 1;
        ix_s[tid] = (l < 0 || l >= TX) ? -1 : (mode == 0 ? reflect_index(ti * g.c[0] + l - g.b[0], g.X) - g.in_lo[0] : l);
    } else if (tid >= 32 && tid < 32 + FY + 2) {
        const int l = y0 + tid - 32 - 1;
        iy_s[tid - 32] = (l < 0 || l >= TY) ? -1 : (mode == 0 ? reflect_index(tj * g.c[1] + l - g.b[1], g.Y) - g.in_lo[1] : l);
    } else if (tid >= 64 && tid < 64 + FZ + 2) {
        const int l = z0 + tid - 64 - 1;
        iz_s[tid - 64] = (l < 0 || l >= TZ) ? -1 : (mode == 0 ? reflect_index(tk * g.c[2] + l - g.b[2], g.Z) - g.in_lo[2] : l);
    }
    __syncthreads();
    const int d1 = mode == 0 ? g.in_dim[1] : TY, d2 = mode == 0 ? g.in_dim[2] : TZ;
    const float* base = mode == 0 ? src : src + (size_t)(tile_first + tile) * vol;
    // all of a thread's 13 gathers are issued before the first is stored: with a rolled loop every trip waited for its own
    // load (~13 L2 round trips per CTA, longer than the 1728 multiply-adds per thread that follow)
    constexpr int N_IN = (FX + 2) * (FY + 2) * (FZ + 2), N_LD = (N_IN + 255) / 256;
    float stage[N_LD];
#pragma unroll
    for (int k = 0; k < N_LD; ++k) {
        const int i = tid + k * 256;
        float v = 0.f;
        if (i < N_IN) {
            const int sz = i % (FZ + 2), r = i / (FZ + 2);
            const int sy = r % (FY + 2), sx = r / (FY + 2);
            const int gx = ix_s[sx], gy = iy_s[sy], gz = iz_s[sz];
            if ((gx | gy | gz) >= 0) v = __ldg(base + ((size_t)gx * d1 + gy) * d2 + gz);
        }
        stage[k] = v;
    }
#pragma unroll
    for (int k = 0; k < N_LD; ++k) {
        const int i = tid + k * 256;
        if (i < N_IN) (&in_s[0][0][0])[i] = stage[k];
    }
    for (int i = tid; i < 27 * 8; i += 256) reinterpret_cast<float*>(&w_s[0][0])[i] = wts[i].x;
    __syncthreads();

    // channel pairs: packed fp32 multiply-adds (FFMA2: one issue slot for two IEEE operations, bit-identical to FFMA)
    float2 acc[FX][4];
#pragma unroll
    for (int v = 0; v < FX; ++v)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[v][c] = make_float2(0.f, 0.f);
#pragma unroll 1
    for (int dy = 0; dy < 3; ++dy) {
#pragma unroll 1
        for (int dz = 0; dz < 3; ++dz) {
            float2 w[3][4];
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                const float4 w0 = w_s[(dx * 3 + dy) * 3 + dz][0], w1 = w_s[(dx * 3 + dy) * 3 + dz][1];
                w[dx][0] = make_float2(w0.x, w0.y); w[dx][1] = make_float2(w0.z, w0.w);
                w[dx][2] = make_float2(w1.x, w1.y); w[dx][3] = make_float2(w1.z, w1.w);
            }
#pragma unroll
            for (int xx = 0; xx < FX + 2; ++xx) {
                const float2 a = f2_splat(in_s[xx][ty + dy][tz + dz]);
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const int v = xx - dx;
                    if (v < 0 || v >= FX) continue;
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[v][c] = f2_fma(a, w[dx][c], acc[v][c]);
                }
            }
        }
    }

    const int ly = y0 + ty, lz = z0 + tz;
    float amax = 0.f;
    if (ly < TY && lz < TZ) {
        float4* d_tile = dst + (size_t)tile * dst_tile_stride4;
#pragma unroll
        for (int c4 = 0; c4 < 2; ++c4) {
            float bs[4], sc[4], sh[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { bs[j] = bias[c4 * 4 + j]; sc[j] = scale[c4 * 4 + j]; sh[j] = shift[c4 * 4 + j]; }
            float4* d_ck = d_tile + (size_t)(dst_c4off + c4) * vol;
#pragma unroll
            for (int v = 0; v < FX; ++v) {
                const int lx = x0 + v;
                if (lx < TX) {
                    float o[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 ap = acc[v][c4 * 2 + (j >> 1)];
                        float t = ((j & 1) ? ap.y : ap.x) + bs[j];
                        t = t > 0.f ? t : alpha * t;
                        o[j] = fmaf(t, sc[j], sh[j]);
                    }
                    d_ck[((size_t)lx * TY + ly) * TZ + lz] = make_float4(o[0], o[1], o[2], o[3]);
                    amax = fmaxf(fmaxf(amax, fmaxf(fabsf(o[0]), fabsf(o[1]))), fmaxf(fabsf(o[2]), fabsf(o[3])));
                }
            }
        }
    }
    amax = warp_max(amax);
    if ((tid & 31) == 0) amax_update(amax_hdr + (size_t)tile * dst_tile_stride4 * 4 + dst_slot, amax);
}

int launch_first_conv_fused(const CtUNet* net, const Op& op, const float* src, int mode, int tile_first,
                            const TileGeom& geo, float* slab0, size_t slab_stride, int tiles, cudaStream_t s) {
    const ConvLayer& L = net->layers[op.layer];
    if (L.cin != 1 || L.cout != 8 || L.cin_pad != 4) return 2;
    CT_REQUIRE(slab_stride % 4 == 0 && op.dst_off % 4 == 0, "conv: misaligned slab");
    const int nbx = cdiv(op.sx, FX), nby = cdiv(op.sy, FY), nbz = cdiv(op.sz, FZ);
    dim3 grid(nbx * nby * nbz, 1, tiles);
    ProfScope prof(PROF_CONV, s);
    first_conv_kernel<<<grid, 256, 0, s>>>(src, mode, tile_first, geo, reinterpret_cast<float4*>(slab0 + op.dst_off),
                                           reinterpret_cast<const float4*>(L.w_direct), L.bias, L.scale, L.shift, net->alpha,
                                           nbx, nby, slab_stride / 4, op.dst_coff / 4, slab0, op.dst_slot);
    CT_LAUNCHED("first_conv_kernel");
    return 0;
}

}  // namespace ct
