// LCN normalisation for the U-Net input (sm_100a, memory-bound CUDA-core kernels).
//
// Replaces preprocess.py:170-188 (_normalize_image), :136-167 (lcn_gpu) and :117-133 (conv3d_keras):
//   v   = max(raw - median(raw), 0)
//   avg = box27x27x1(v) / 729            (zero padded, float32 like the Keras Conv3D output)
//   std = sqrt(box27x27x1((v-avg)^2)/729)
//   out = (v - avg) / (std + noise_level)
//
// The exact median (np.median: mean of the two middle order statistics for even counts) is found by
// an 8-bit-digit radix select over order-preserving integer keys: 2 passes for uint16/uint8 input,
// 4 for float32.  Histograms are privatised per warp in shared memory; lanes that hit the same bin are
// merged with __match_any_sync before the shared-memory atomic, because microscopy stacks put most
// voxels into a handful of background bins.
#include "common.cuh"
#include <cstddef>
#include <type_traits>

namespace ct {

// ---------------------------------------------------------------------------------------------
// keys
// ---------------------------------------------------------------------------------------------
template <typename T> struct KeyOf;
template <> struct KeyOf<uint16_t> {
    static constexpr int bits = 16;
    __device__ static uint32_t key(uint16_t v) { return v; }
    __device__ static double value(uint32_t k) { return (double)k; }
};
template <> struct KeyOf<uint8_t> {
    static constexpr int bits = 8;
    __device__ static uint32_t key(uint8_t v) { return v; }
    __device__ static double value(uint32_t k) { return (double)k; }
};
template <> struct KeyOf<float> {
    static constexpr int bits = 32;
    __device__ static uint32_t key(float f) {
        uint32_t u = __float_as_uint(f);
        return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    }
    __device__ static double value(uint32_t k) {
        uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
        return (double)__uint_as_float(u);
    }
};

struct SelectState {
    unsigned long long rank[2];    // remaining rank inside the current prefix bucket
    uint32_t prefix[2];            // key bits fixed so far (high bits)
    uint32_t hist[2][256];
};

__global__ void select_init(SelectState* st, long long count) {
    int t = threadIdx.x;
    if (t == 0) {
        st->rank[0] = (unsigned long long)((count - 1) / 2);
        st->rank[1] = (unsigned long long)(count / 2);
        st->prefix[0] = st->prefix[1] = 0;
    }
    for (int i = t; i < 512; i += blockDim.x) (&st->hist[0][0])[i] = 0;
}

// One digit pass: histogram of digit `shift` for keys whose higher bits equal prefix[r].
template <typename T>
__global__ void __launch_bounds__(256) select_hist(const T* __restrict__ data, long long count,
                                                   SelectState* st, int shift, int key_bits) {
    __shared__ uint32_t sh[8][2][256];
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 8 * 2 * 256; i += 256) (&sh[0][0][0])[i] = 0;
    __syncthreads();
    const uint32_t p0 = st->prefix[0], p1 = st->prefix[1];
    const bool same = (p0 == p1);
    const int hi_shift = shift + 8;                       // bits above the current digit
    const bool top = hi_shift >= key_bits;                // first pass: no prefix to match
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long n_iter = (count + stride - 1) / stride;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long it = 0; it < n_iter; ++it, i += stride) {
        bool valid = i < count;
        uint32_t k = valid ? KeyOf<T>::key(data[i]) : 0u;
        uint32_t digit = (k >> shift) & 0xffu;
        uint32_t hi = top ? 0u : (k >> hi_shift);
        bool m0 = valid && (top || hi == (p0 >> hi_shift));
        bool m1 = valid && !same && (top || hi == (p1 >> hi_shift));
        // merge lanes with the same (bucket, digit)
        uint32_t tag = m0 ? digit : (m1 ? 256u + digit : 0xffffu);
        uint32_t peers = __match_any_sync(0xffffffffu, tag);
        if (tag != 0xffffu && (__ffs(peers) - 1) == (threadIdx.x & 31)) {
            atomicAdd(&sh[warp][tag >> 8][tag & 0xffu], __popc(peers));
        }
    }
    __syncthreads();
    for (int b = threadIdx.x; b < 512; b += 256) {
        uint32_t s = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += (&sh[w][0][0])[b];
        if (s) atomicAdd(&(&st->hist[0][0])[b], s);
    }
}

// Pick the digit holding the wanted rank, update prefix / rank, clear histograms for the next pass.
__global__ void select_scan(SelectState* st, int shift) {
    if (threadIdx.x < 2) {
        int r = threadIdx.x;
        bool same = st->prefix[0] == st->prefix[1];
        const uint32_t* h = st->hist[(r == 1 && same) ? 0 : r];
        unsigned long long rank = st->rank[r];
        uint32_t d = 0;
        for (; d < 255; ++d) {
            if (rank < h[d]) break;
            rank -= h[d];
        }
        st->rank[r] = rank;
        // defer the prefix write until both lanes have read `same`
        __syncwarp(0x3);
        st->prefix[r] |= d << shift;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 512; i += blockDim.x) (&st->hist[0][0])[i] = 0;
}

template <typename T>
__global__ void select_finish(const SelectState* st, double* median_out) {
    *median_out = 0.5 * (KeyOf<T>::value(st->prefix[0]) + KeyOf<T>::value(st->prefix[1]));
}

template <typename T>
static int median_impl(const T* data, long long count, double* median_out, SelectState* st, cudaStream_t s) {
    const int bits = KeyOf<T>::bits;
    select_init<<<1, 256, 0, s>>>(st, count);
    CT_LAUNCHED("select_init");
    int blocks = (int)((count + 256 * 16 - 1) / (256 * 16));
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    for (int shift = bits - 8; shift >= 0; shift -= 8) {
        select_hist<T><<<blocks, 256, 0, s>>>(data, count, st, shift, bits);
        CT_LAUNCHED("select_hist");
        select_scan<<<1, 256, 0, s>>>(st, shift);
        CT_LAUNCHED("select_scan");
    }
    select_finish<T><<<1, 1, 0, s>>>(st, median_out);
    CT_LAUNCHED("select_finish");
    return 0;
}

// ---------------------------------------------------------------------------------------------
// box filters.  Layout (x, y, z), z fastest; the filter spans x and y only.
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float clamped(const T* __restrict__ raw, long long idx, double med) {
    // med is NaN when the caller asked for plain lcn_gpu (no median subtraction, no clamp)
    if (med != med) return (float)raw[idx];
    double v = (double)raw[idx] - med;
    return v > 0.0 ? (float)v : 0.0f;       // keras casts the float64 array to float32 (exact here)
}

// Every box-filter thread produces LCN_R consecutive outputs along the filtered axis from ONE pass over the
// LCN_R + 2r inputs they share (register window), instead of 2r + 1 loads per output.  Each output still sums ITS OWN
// window in ascending order with its own accumulator, so a value does not depend on where the block / volume starts:
// this is what keeps the spatially decomposed LCN (spatial.py) bit-identical to the single-GPU one.
constexpr int LCN_R = 8;

// pass 1: T1 = sum over dy of v   (float32; exact for integer-valued input)
template <typename T>
__global__ void __launch_bounds__(256) box_y_v(const T* __restrict__ raw, const double* __restrict__ med_p,
                                               float* __restrict__ t1, int X, int Y, int Z, int ry) {
    // thread = (y group of LCN_R, z); grid.y = x
    const int groups = (Y + LCN_R - 1) / LCN_R;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= groups * Z) return;
    const int z = f % Z, yb = (f / Z) * LCN_R;
    const long long plane = (long long)Y * Z;
    const T* base = raw + (long long)blockIdx.y * plane + z;
    float* out = t1 + (long long)blockIdx.y * plane + z;
    const double med = *med_p;
    float acc[LCN_R];
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) acc[k] = 0.f;
    const int lo = max(yb - ry, 0), hi = min(yb + LCN_R - 1 + ry, Y - 1);
    for (int yy = lo; yy <= hi; ++yy) {
        const float v = clamped(base, (long long)yy * Z, med);
#pragma unroll
        for (int k = 0; k < LCN_R; ++k)
            if (yy >= yb + k - ry && yy <= yb + k + ry) acc[k] += v;
    }
#pragma unroll
    for (int k = 0; k < LCN_R; ++k)
        if (yb + k < Y) out[(long long)(yb + k) * Z] = acc[k];
}

// pass 2: avg = float32(sum over dx of T1) / volume
__global__ void __launch_bounds__(256) box_x_avg(const float* __restrict__ t1, float* __restrict__ avg,
                                                 int X, int Y, int Z, int rx, float volume) {
    // thread = one (y, z) column position; grid.y = x group of LCN_R
    const long long plane = (long long)Y * Z;
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= plane) return;
    const int xb = blockIdx.y * LCN_R;
    double acc[LCN_R];
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) acc[k] = 0.0;
    const int lo = max(xb - rx, 0), hi = min(xb + LCN_R - 1 + rx, X - 1);
    for (int xx = lo; xx <= hi; ++xx) {
        const double v = (double)t1[(long long)xx * plane + f];
#pragma unroll
        for (int k = 0; k < LCN_R; ++k)
            if (xx >= xb + k - rx && xx <= xb + k + rx) acc[k] += v;
    }
#pragma unroll
    for (int k = 0; k < LCN_R; ++k)
        if (xb + k < X) avg[(long long)(xb + k) * plane + f] = (float)acc[k] / volume;
}

// pass 3: T2 = sum over dy of float32((v-avg)^2)
template <typename T>
__global__ void __launch_bounds__(256) box_y_sq(const T* __restrict__ raw, const double* __restrict__ med_p,
                                                const float* __restrict__ avg, float* __restrict__ t2,
                                                int X, int Y, int Z, int ry) {
    const int groups = (Y + LCN_R - 1) / LCN_R;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= groups * Z) return;
    const int z = f % Z, yb = (f / Z) * LCN_R;
    const long long plane = (long long)Y * Z;
    const long long off = (long long)blockIdx.y * plane + z;
    const double med = *med_p;
    double acc[LCN_R];
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) acc[k] = 0.0;
    const int lo = max(yb - ry, 0), hi = min(yb + LCN_R - 1 + ry, Y - 1);
    for (int yy = lo; yy <= hi; ++yy) {
        const long long i = off + (long long)yy * Z;
        const double d = (double)clamped(raw, i, med) - (double)avg[i];
        const double sq = (double)(float)(d * d);
#pragma unroll
        for (int k = 0; k < LCN_R; ++k)
            if (yy >= yb + k - ry && yy <= yb + k + ry) acc[k] += sq;
    }
#pragma unroll
    for (int k = 0; k < LCN_R; ++k)
        if (yb + k < Y) t2[off + (long long)(yb + k) * Z] = (float)acc[k];
}

// pass 4: std, normalise
template <typename T>
__global__ void __launch_bounds__(256) box_x_norm(const T* __restrict__ raw, const double* __restrict__ med_p,
                                                  const float* __restrict__ avg, const float* __restrict__ t2,
                                                  float* __restrict__ out, int X, int Y, int Z, int rx,
                                                  float volume, float noise) {
    const long long plane = (long long)Y * Z;
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= plane) return;
    const int xb = blockIdx.y * LCN_R;
    double acc[LCN_R];
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) acc[k] = 0.0;
    const int lo = max(xb - rx, 0), hi = min(xb + LCN_R - 1 + rx, X - 1);
    for (int xx = lo; xx <= hi; ++xx) {
        const double v = (double)t2[(long long)xx * plane + f];
#pragma unroll
        for (int k = 0; k < LCN_R; ++k)
            if (xx >= xb + k - rx && xx <= xb + k + rx) acc[k] += v;
    }
    const double med = *med_p;
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) {
        if (xb + k >= X) break;
        const long long i = (long long)(xb + k) * plane + f;
        const float sd = sqrtf((float)acc[k] / volume);
        const float den = sd + noise;
        const double d = (double)clamped(raw, i, med) - (double)avg[i];
        out[i] = (float)(d / (double)den);
    }
}

// ---- the reference's window (27 x 27 x 1, preprocess.py:185): radius known at compile time.  Each thread first loads
// the LCN_R + 2 R inputs its LCN_R outputs share (independent loads: the memory system sees them all at once, the generic
// kernels above issue one dependent load per loop trip), then every output sums ITS OWN window from the registers in
// ascending order.  Out-of-range inputs are 0, and x + 0 is exact, so the results are bit-identical to the generic kernels.
template <typename T, int R>
__global__ void __launch_bounds__(256) box_y_v_r(const T* __restrict__ raw, const double* __restrict__ med_p,
                                                 float* __restrict__ t1, int X, int Y, int Z) {
    const int groups = (Y + LCN_R - 1) / LCN_R;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= groups * Z) return;
    const int z = f % Z, yb = (f / Z) * LCN_R;
    const long long plane = (long long)Y * Z;
    const T* base = raw + (long long)blockIdx.y * plane + z;
    float* out = t1 + (long long)blockIdx.y * plane + z;
    const double med = *med_p;
    float w[LCN_R + 2 * R];
#pragma unroll
    for (int j = 0; j < LCN_R + 2 * R; ++j) {
        const int yy = yb - R + j;
        w[j] = (yy >= 0 && yy < Y) ? clamped(base, (long long)yy * Z, med) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j <= 2 * R; ++j) acc += w[k + j];
        if (yb + k < Y) out[(long long)(yb + k) * Z] = acc;
    }
}
// INT: T1 is a multiple of 0.5 below 2^22 (integer raw data minus a median that is k or k + 0.5, summed over 27 rows), so
// the window sum is done on 2 T1 in int32 -- exact, like the fp64 sum it replaces, and float(isum) * 0.5f rounds exactly
// like float(double sum): the result is bit-identical, without the fp64 pipe (2 DADD lanes per scheduler on this part:
// the four filters were bound by it, 73-79 % busy at 0.09-0.25 ms each).
template <int R, bool INT>
__global__ void __launch_bounds__(256) box_x_avg_r(const float* __restrict__ t1, float* __restrict__ avg,
                                                   int X, int Y, int Z, float volume) {
    const long long plane = (long long)Y * Z;
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= plane) return;
    const int xb = blockIdx.y * LCN_R;
    if constexpr (INT) {
        int w[LCN_R + 2 * R];
#pragma unroll
        for (int j = 0; j < LCN_R + 2 * R; ++j) {
            const int xx = xb - R + j;
            w[j] = (xx >= 0 && xx < X) ? __float2int_rn(2.f * t1[(long long)xx * plane + f]) : 0;
        }
#pragma unroll
        for (int k = 0; k < LCN_R; ++k) {
            int acc = 0;
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) acc += w[k + j];
            if (xb + k < X) avg[(long long)(xb + k) * plane + f] = (__int2float_rn(acc) * 0.5f) / volume;
        }
    } else {
        float w[LCN_R + 2 * R];
#pragma unroll
        for (int j = 0; j < LCN_R + 2 * R; ++j) {
            const int xx = xb - R + j;
            w[j] = (xx >= 0 && xx < X) ? t1[(long long)xx * plane + f] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < LCN_R; ++k) {
            double acc = 0.0;
#pragma unroll
            for (int j = 0; j <= 2 * R; ++j) acc += (double)w[k + j];
            if (xb + k < X) avg[(long long)(xb + k) * plane + f] = (float)acc / volume;
        }
    }
}
template <typename T, int R>
__global__ void __launch_bounds__(256) box_y_sq_r(const T* __restrict__ raw, const double* __restrict__ med_p,
                                                  const float* __restrict__ avg, float* __restrict__ t2,
                                                  int X, int Y, int Z) {
    const int groups = (Y + LCN_R - 1) / LCN_R;
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= groups * Z) return;
    const int z = f % Z, yb = (f / Z) * LCN_R;
    const long long plane = (long long)Y * Z;
    const long long off = (long long)blockIdx.y * plane + z;
    const double med = *med_p;
    float w[LCN_R + 2 * R];
#pragma unroll
    for (int j = 0; j < LCN_R + 2 * R; ++j) {
        const int yy = yb - R + j;
        float sq = 0.f;
        if (yy >= 0 && yy < Y) {
            const long long i = off + (long long)yy * Z;
            const float d = __fsub_rn(clamped(raw, i, med), avg[i]);      // correctly rounded difference, then its square
            sq = __fmul_rn(d, d);
        }
        w[j] = sq;
    }
    // float32 window sums in ascending order (the reference's Conv3D sums in float32 as well; every output still sums its
    // OWN window, so the value does not depend on where a block starts)
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j <= 2 * R; ++j) acc += w[k + j];
        if (yb + k < Y) t2[off + (long long)(yb + k) * Z] = acc;
    }
}
template <typename T, int R>
__global__ void __launch_bounds__(256) box_x_norm_r(const T* __restrict__ raw, const double* __restrict__ med_p,
                                                    const float* __restrict__ avg, const float* __restrict__ t2,
                                                    float* __restrict__ out, int X, int Y, int Z, float volume, float noise) {
    const long long plane = (long long)Y * Z;
    const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= plane) return;
    const int xb = blockIdx.y * LCN_R;
    float w[LCN_R + 2 * R];
#pragma unroll
    for (int j = 0; j < LCN_R + 2 * R; ++j) {
        const int xx = xb - R + j;
        w[j] = (xx >= 0 && xx < X) ? t2[(long long)xx * plane + f] : 0.f;
    }
    const double med = *med_p;
#pragma unroll
    for (int k = 0; k < LCN_R; ++k) {
        if (xb + k >= X) break;
        float acc = 0.f;
#pragma unroll
        for (int j = 0; j <= 2 * R; ++j) acc += w[k + j];
        const long long i = (long long)(xb + k) * plane + f;
        const float sd = sqrtf(acc / volume);
        const float den = sd + noise;
        // float32 subtraction is correctly rounded, i.e. equal to float(double(v) - double(avg)); one float32 division
        const float d = __fsub_rn(clamped(raw, i, med), avg[i]);
        out[i] = __fdiv_rn(d, den);
    }
}

__global__ void set_nan(double* p) { *p = __longlong_as_double(0x7ff8000000000000LL); }

template <typename T>
static int normalize_impl(const T* raw, float* out, int X, int Y, int Z, float noise, int fx, int fy,
                          int subtract_median, const double* median_given, void* ws, size_t ws_bytes, cudaStream_t s) {
    const long long n = (long long)X * Y * Z;
    ProfScope prof(PROF_LCN, s);
    Arena a(ws, ws_bytes);
    SelectState* st = a.take<SelectState>(1);
    double* med = a.take<double>(1);
    float* t1 = a.take<float>(n);
    float* avg = a.take<float>(n);
    float* t2 = t1;     // t1 is dead after pass 2
    CT_REQUIRE(a.ok(), "ct_normalize_image: workspace too small (%zu < %zu)", ws_bytes, a.off);
    if (median_given) {
        med = const_cast<double*>(median_given);
    } else if (subtract_median) {
        if (median_impl<T>(raw, n, med, st, s)) return 1;
    } else {
        set_nan<<<1, 1, 0, s>>>(med);
        CT_LAUNCHED("set_nan");
    }
    const long long plane = (long long)Y * Z;
    const int ygroups = (Y + LCN_R - 1) / LCN_R, xgroups = (X + LCN_R - 1) / LCN_R;
    dim3 grid_y((unsigned)(((long long)ygroups * Z + 255) / 256), X);       // y filters: thread = (y group, z), per x
    dim3 grid_x((unsigned)((plane + 255) / 256), xgroups);                  // x filters: thread = (y, z), per x group
    const float volume = (float)(fx * fy);
    if (fx == 27 && fy == 27) {
        box_y_v_r<T, 13><<<grid_y, 256, 0, s>>>(raw, med, t1, X, Y, Z);
        CT_LAUNCHED("box_y_v");
        // integer raw data: 2 T1 is an integer (the median of integers is k or k + 0.5)
        box_x_avg_r<13, !std::is_floating_point<T>::value><<<grid_x, 256, 0, s>>>(t1, avg, X, Y, Z, volume);
        CT_LAUNCHED("box_x_avg");
        box_y_sq_r<T, 13><<<grid_y, 256, 0, s>>>(raw, med, avg, t2, X, Y, Z);
        CT_LAUNCHED("box_y_sq");
        box_x_norm_r<T, 13><<<grid_x, 256, 0, s>>>(raw, med, avg, t2, out, X, Y, Z, volume, noise);
        CT_LAUNCHED("box_x_norm");
        return 0;
    }
    box_y_v<T><<<grid_y, 256, 0, s>>>(raw, med, t1, X, Y, Z, fy / 2);
    CT_LAUNCHED("box_y_v");
    box_x_avg<<<grid_x, 256, 0, s>>>(t1, avg, X, Y, Z, fx / 2, volume);
    CT_LAUNCHED("box_x_avg");
    box_y_sq<T><<<grid_y, 256, 0, s>>>(raw, med, avg, t2, X, Y, Z, fy / 2);
    CT_LAUNCHED("box_y_sq");
    box_x_norm<T><<<grid_x, 256, 0, s>>>(raw, med, avg, t2, out, X, Y, Z, fx / 2, volume, noise);
    CT_LAUNCHED("box_x_norm");
    return 0;
}

}  // namespace ct

extern "C" {

size_t ct_normalize_workspace_bytes(int x, int y, int z) {
    size_t n = (size_t)x * y * z;
    return 2 * ct::align_up(n * sizeof(float), 256) + ct::align_up(sizeof(ct::SelectState), 256) + 1024;
}

int ct_median(const void* raw, int dtype, long long count, double* median_out, void* ws, size_t ws_bytes,
              void* stream) {
    CT_REQUIRE(count > 0, "ct_median: empty input");
    CT_REQUIRE(ws_bytes >= sizeof(ct::SelectState) + 256, "ct_median: workspace too small");
    ct::Arena a(ws, ws_bytes);
    ct::SelectState* st = a.take<ct::SelectState>(1);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case 0: return ct::median_impl<uint16_t>((const uint16_t*)raw, count, median_out, st, s);
        case 1: return ct::median_impl<float>((const float*)raw, count, median_out, st, s);
        case 2: return ct::median_impl<uint8_t>((const uint8_t*)raw, count, median_out, st, s);
    }
    ct::set_error("ct_median: unsupported dtype %d", dtype);
    return 1;
}

int ct_normalize_image(const void* raw, int dtype, float* out, int x, int y, int z, float noise_level,
                       int filter_x, int filter_y, int subtract_median, void* ws, size_t ws_bytes, void* stream) {
    CT_REQUIRE(x > 0 && y > 0 && z > 0, "ct_normalize_image: empty volume");
    CT_REQUIRE((filter_x & 1) && (filter_y & 1), "ct_normalize_image: filter sizes must be odd");
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case 0: return ct::normalize_impl<uint16_t>((const uint16_t*)raw, out, x, y, z, noise_level, filter_x, filter_y, subtract_median, nullptr, ws, ws_bytes, s);
        case 1: return ct::normalize_impl<float>((const float*)raw, out, x, y, z, noise_level, filter_x, filter_y, subtract_median, nullptr, ws, ws_bytes, s);
        case 2: return ct::normalize_impl<uint8_t>((const uint8_t*)raw, out, x, y, z, noise_level, filter_x, filter_y, subtract_median, nullptr, ws, ws_bytes, s);
    }
    ct::set_error("ct_normalize_image: unsupported dtype %d", dtype);
    return 1;
}

int ct_normalize_image_with_median(const void* raw, int dtype, float* out, int x, int y, int z, float noise_level,
                                   int filter_x, int filter_y, const double* median, void* ws, size_t ws_bytes,
                                   void* stream) {
    CT_REQUIRE(x > 0 && y > 0 && z > 0, "ct_normalize_image_with_median: empty block");
    CT_REQUIRE((filter_x & 1) && (filter_y & 1), "ct_normalize_image_with_median: filter sizes must be odd");
    CT_REQUIRE(median, "ct_normalize_image_with_median: null median");
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype) {
        case 0: return ct::normalize_impl<uint16_t>((const uint16_t*)raw, out, x, y, z, noise_level, filter_x, filter_y, 1, median, ws, ws_bytes, s);
        case 1: return ct::normalize_impl<float>((const float*)raw, out, x, y, z, noise_level, filter_x, filter_y, 1, median, ws, ws_bytes, s);
        case 2: return ct::normalize_impl<uint8_t>((const uint8_t*)raw, out, x, y, z, noise_level, filter_x, filter_y, 1, median, ws, ws_bytes, s);
    }
    ct::set_error("ct_normalize_image_with_median: unsupported dtype %d", dtype);
    return 1;
}

// ---- lock-step radix select for a volume spread over several GPUs (see ct3d.h)
size_t ct_select_state_bytes(void) { return sizeof(ct::SelectState); }
size_t ct_select_hist_offset(void) { return offsetof(ct::SelectState, hist); }
int ct_select_passes(int dtype) { return dtype == 0 ? 2 : dtype == 1 ? 4 : dtype == 2 ? 1 : -1; }

int ct_select_begin(void* state, long long total_count, void* stream) {
    CT_REQUIRE(state && total_count > 0, "ct_select_begin: bad argument");
    ct::select_init<<<1, 256, 0, (cudaStream_t)stream>>>((ct::SelectState*)state, total_count);
    CT_LAUNCHED("select_init");
    return 0;
}

int ct_select_hist(const void* raw, int dtype, long long local_count, void* state, int pass, void* stream) {
    const int passes = ct_select_passes(dtype);
    CT_REQUIRE(state && passes > 0 && pass >= 0 && pass < passes && local_count >= 0, "ct_select_hist: bad argument");
    if (local_count == 0) return 0;
    CT_REQUIRE(raw, "ct_select_hist: null data");
    const int bits = passes * 8, shift = bits - 8 * (pass + 1);
    int blocks = (int)((local_count + 256 * 16 - 1) / (256 * 16));
    if (blocks > 148 * 8) blocks = 148 * 8;
    cudaStream_t s = (cudaStream_t)stream;
    ct::SelectState* st = (ct::SelectState*)state;
    switch (dtype) {
        case 0: ct::select_hist<uint16_t><<<blocks, 256, 0, s>>>((const uint16_t*)raw, local_count, st, shift, bits); break;
        case 1: ct::select_hist<float><<<blocks, 256, 0, s>>>((const float*)raw, local_count, st, shift, bits); break;
        default: ct::select_hist<uint8_t><<<blocks, 256, 0, s>>>((const uint8_t*)raw, local_count, st, shift, bits); break;
    }
    CT_LAUNCHED("select_hist");
    return 0;
}

int ct_select_scan(void* state, int dtype, int pass, void* stream) {
    const int passes = ct_select_passes(dtype);
    CT_REQUIRE(state && passes > 0 && pass >= 0 && pass < passes, "ct_select_scan: bad argument");
    ct::select_scan<<<1, 256, 0, (cudaStream_t)stream>>>((ct::SelectState*)state, passes * 8 - 8 * (pass + 1));
    CT_LAUNCHED("select_scan");
    return 0;
}

int ct_select_finish(const void* state, int dtype, double* median_out, void* stream) {
    CT_REQUIRE(state && median_out, "ct_select_finish: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    const ct::SelectState* st = (const ct::SelectState*)state;
    switch (dtype) {
        case 0: ct::select_finish<uint16_t><<<1, 1, 0, s>>>(st, median_out); break;
        case 1: ct::select_finish<float><<<1, 1, 0, s>>>(st, median_out); break;
        case 2: ct::select_finish<uint8_t><<<1, 1, 0, s>>>(st, median_out); break;
        default: ct::set_error("ct_select_finish: unsupported dtype %d", dtype); return 1;
    }
    CT_LAUNCHED("select_finish");
    return 0;
}

}  // extern "C"
