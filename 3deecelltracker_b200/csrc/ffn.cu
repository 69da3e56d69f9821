// FFN match: k-NN features, the factored FFN forward and the pairwise correspondence kernel.
//
// Reference: ffn.py:225-265 (FFN.call), ffn.py:268-327 (initial_matching_ffn), track.py:117-178
// (initial_matching_quick).  The reference materialises an (M*N, 122) pair grid and pushes it through
//   h = leaky(bn2([f(ref n) ; f(tgt m)] . W2)),  f(v) = leaky(bn1(v . W1)),  corr = sigmoid(h . w3 + b3).
// Because the concatenation is followed by a bias-free Dense, [f_r ; f_t] . W2 = f_r . W2[:512] + f_t . W2[512:]
// exactly, so the work factors into two small GEMMs over N + M rows and one pairwise kernel over M x N
// pairs that never materialises the grid.  All arithmetic is fp32 like the reference.
#include "common.cuh"
#include <vector>

namespace ct {

constexpr int FEAT = 61;
constexpr int HID = 512;
constexpr float LEAKY = 0.3f;      // keras LeakyReLU default
constexpr float BN_EPS = 1e-3f;    // keras BatchNormalization default

// ---------------------------------------------------------------------------------------------
// k-NN features: one warp per point, k+1 rounds of "smallest (d^2, index) greater than the last pick"
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) knn_features_kernel(const double* __restrict__ pts, int n, int k,
                                                           float* __restrict__ feat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n) return;
    const double px = pts[3 * warp], py = pts[3 * warp + 1], pz = pts[3 * warp + 2];
    double last_d = -1.0;
    int last_i = -1;
    double dist[32];           // k + 1 <= 32 picks, identical in every lane after the butterfly
    int nbr[32];
    for (int r = 0; r <= k; ++r) {
        double best_d = INFINITY;
        int best_i = 0x7fffffff;
        for (int j = lane; j < n; j += 32) {
            const double dx = pts[3 * j] - px, dy = pts[3 * j + 1] - py, dz = pts[3 * j + 2] - pz;
            const double d = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
            const bool after = (d > last_d) || (d == last_d && j > last_i);
            if (after && (d < best_d || (d == best_d && j < best_i))) { best_d = d; best_i = j; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double od = __shfl_xor_sync(0xffffffffu, best_d, o);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
            if (od < best_d || (od == best_d && oi < best_i)) { best_d = od; best_i = oi; }
        }
        last_d = best_d; last_i = best_i;
#pragma unroll
        for (int q = 0; q < 32; ++q) if (q == r) { dist[q] = sqrt(best_d); nbr[q] = best_i; }
    }
    // np.mean over k+1 values: numpy's pairwise_sum keeps 8 accumulators for 8 <= n < 128
    double sum;
    const int cnt = k + 1;
    if (cnt < 8) {
        sum = 0.0;
#pragma unroll
        for (int i = 0; i < 32; ++i) if (i < cnt) sum += dist[i];
    } else {
        double r8[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r8[q] = dist[q];
        const int full = cnt - (cnt % 8);
#pragma unroll
        for (int i = 8; i < 32; ++i) if (i < full) r8[i & 7] += dist[i];
        sum = ((r8[0] + r8[1]) + (r8[2] + r8[3])) + ((r8[4] + r8[5]) + (r8[6] + r8[7]));
#pragma unroll
        for (int i = 8; i < 32; ++i) if (i >= full && i < cnt) sum += dist[i];
    }
    const double mean_dist = sum / (double)cnt;
    float* out = feat + (size_t)warp * (3 * k + 1);
    const int centre = nbr[0];
#pragma unroll
    for (int r = 1; r < 32; ++r) {
        if (r <= k && lane < 3)
            out[(r - 1) * 3 + lane] = (float)((pts[3 * nbr[r] + lane] - pts[3 * centre + lane]) / mean_dist);
    }
    if (lane == 0) out[3 * k] = (float)mean_dist;
}

// ---------------------------------------------------------------------------------------------
// small fp32 GEMM  C[M,N] (+)= A[M,K] . B[K,N]  with optional BN + LeakyReLU epilogue
// 64 x 64 tile per CTA, 256 threads, 4 x 4 register tile, K staged in chunks of 16.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
sgemm_bn_kernel(const float* __restrict__ A, int lda, const float* __restrict__ B, int ldb,
                float* __restrict__ C, int ldc, int M, int N, int K, int accumulate,
                const float* __restrict__ scale, const float* __restrict__ shift, int act) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
    float acc[4][4] = {};
    for (int k0 = 0; k0 < K; k0 += 16) {
        for (int i = tid; i < 64 * 16; i += 256) {
            const int kk = i & 15, mm = i >> 4;
            const int gm = m0 + mm, gk = k0 + kk;
            As[kk][mm] = (gm < M && gk < K) ? A[(size_t)gm * lda + gk] : 0.f;
        }
        for (int i = tid; i < 64 * 16; i += 256) {
            const int nn = i & 63, kk = i >> 6;
            const int gn = n0 + nn, gk = k0 + kk;
            Bs[kk][nn] = (gn < N && gk < K) ? B[(size_t)gk * ldb + gn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (accumulate) v += C[(size_t)gm * ldc + gn];
            if (scale) v = fmaf(v, scale[gn], shift[gn]);
            if (act) v = v > 0.f ? v : LEAKY * v;
            C[(size_t)gm * ldc + gn] = v;
        }
    }
}

// corr[m, n] = sigmoid(b3 + sum_c w3[c] * leaky(s2[c] * (A[n,c] + B[m,c]) + t2[c]))
// CTA = 16 (n) x 16 (m) pairs; hidden dimension streamed through shared memory in chunks of 64.
__global__ void __launch_bounds__(256)
ffn_pair_kernel(const float* __restrict__ A, const float* __restrict__ B, int N, int M,
                const float* __restrict__ s2, const float* __restrict__ t2, const float* __restrict__ w3,
                float b3, float* __restrict__ corr) {
    __shared__ float As[64][17];
    __shared__ float Bs[64][17];
    __shared__ float Ss[64], Ts[64], Ws[64];
    const int tid = threadIdx.x;
    const int tn = tid & 15, tm = tid >> 4;
    const int n0 = blockIdx.x * 16, m0 = blockIdx.y * 16;
    float acc = 0.f;
    for (int c0 = 0; c0 < HID; c0 += 64) {
        for (int i = tid; i < 16 * 64; i += 256) {
            const int c = i & 63, r = i >> 6;
            As[c][r] = (n0 + r < N) ? A[(size_t)(n0 + r) * HID + c0 + c] : 0.f;
            Bs[c][r] = (m0 + r < M) ? B[(size_t)(m0 + r) * HID + c0 + c] : 0.f;
        }
        if (tid < 64) { Ss[tid] = s2[c0 + tid]; Ts[tid] = t2[c0 + tid]; Ws[tid] = w3[c0 + tid]; }
        __syncthreads();
#pragma unroll 16
        for (int c = 0; c < 64; ++c) {
            float h = fmaf(As[c][tn] + Bs[c][tm], Ss[c], Ts[c]);
            h = h > 0.f ? h : LEAKY * h;
            acc = fmaf(h, Ws[c], acc);
        }
        __syncthreads();
    }
    const int n = n0 + tn, m = m0 + tm;
    if (n < N && m < M) corr[(size_t)m * N + n] = 1.f / (1.f + expf(-(acc + b3)));
}

// out[r] = sigmoid(H[r,:] . w3 + b3)   (one warp per row)
__global__ void __launch_bounds__(256) ffn_head_kernel(const float* __restrict__ H, int rows,
                                                       const float* __restrict__ w3, float b3,
                                                       float* __restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= rows) return;
    float acc = 0.f;
    for (int c = lane; c < HID; c += 32) acc = fmaf(H[(size_t)warp * HID + c], w3[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[warp] = 1.f / (1.f + expf(-(acc + b3)));
}

static int sgemm(const float* A, int lda, const float* B, int ldb, float* C, int ldc, int M, int N, int K,
                 int accumulate, const float* scale, const float* shift, int act, cudaStream_t s) {
    if (M == 0) return 0;
    dim3 grid(cdiv(N, 64), cdiv(M, 64));
    sgemm_bn_kernel<<<grid, 256, 0, s>>>(A, lda, B, ldb, C, ldc, M, N, K, accumulate, scale, shift, act);
    CT_LAUNCHED("sgemm_bn_kernel");
    return 0;
}

}  // namespace ct

struct CtFFN {
    float* dev;                 // single allocation
    float *W1, *s1, *t1, *W2, *s2, *t2, *w3;
    float b3;
};

using namespace ct;

extern "C" size_t ct_ffn_weight_count(void) {
    return (size_t)FEAT * HID + 4 * HID + (size_t)2 * HID * HID + 4 * HID + HID + 1;
}

extern "C" int ct_ffn_create(const float* w, size_t n_floats, CtFFN** out) {
    CT_REQUIRE(w && out, "ct_ffn_create: null argument");
    CT_REQUIRE(n_floats == ct_ffn_weight_count(), "ct_ffn_create: expected %zu weights, got %zu",
               ct_ffn_weight_count(), n_floats);
    const float* W1 = w;
    const float* bn1 = W1 + (size_t)FEAT * HID;
    const float* W2 = bn1 + 4 * HID;
    const float* bn2 = W2 + (size_t)2 * HID * HID;
    const float* W3 = bn2 + 4 * HID;
    std::vector<float> host((size_t)FEAT * HID + 2 * HID + (size_t)2 * HID * HID + 2 * HID + HID);
    size_t o = 0;
    auto put = [&](const float* p, size_t n) { size_t at = o; for (size_t i = 0; i < n; ++i) host[o++] = p[i]; return at; };
    auto put_bn = [&](const float* bn, size_t& s_at, size_t& t_at) {
        s_at = o;
        for (int c = 0; c < HID; ++c) host[o++] = bn[c] / std::sqrt(bn[3 * HID + c] + BN_EPS);
        t_at = o;
        for (int c = 0; c < HID; ++c) host[o++] = bn[HID + c] - bn[2 * HID + c] * host[s_at + c];
    };
    size_t w1_at = put(W1, (size_t)FEAT * HID), s1_at, t1_at, s2_at, t2_at;
    put_bn(bn1, s1_at, t1_at);
    size_t w2_at = put(W2, (size_t)2 * HID * HID);
    put_bn(bn2, s2_at, t2_at);
    size_t w3_at = put(W3, HID);
    CtFFN* f = new CtFFN();
    if (check_cuda(cudaMalloc(&f->dev, host.size() * sizeof(float)), "cudaMalloc(ffn)") ||
        check_cuda(cudaMemcpy(f->dev, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(ffn)")) {
        delete f;
        return 1;
    }
    f->W1 = f->dev + w1_at; f->s1 = f->dev + s1_at; f->t1 = f->dev + t1_at;
    f->W2 = f->dev + w2_at; f->s2 = f->dev + s2_at; f->t2 = f->dev + t2_at;
    f->w3 = f->dev + w3_at; f->b3 = W3[HID];
    *out = f;
    return 0;
}

extern "C" void ct_ffn_destroy(CtFFN* f) {
    if (!f) return;
    cudaFree(f->dev);
    delete f;
}

extern "C" int ct_knn_features(const double* pts, int n, int k, float* feat, void* stream) {
    CT_REQUIRE(pts && feat, "ct_knn_features: null argument");
    CT_REQUIRE(k >= 1 && k <= 31, "ct_knn_features: k = %d out of range [1,31]", k);
    CT_REQUIRE(n >= k + 1, "Expected n_neighbors <= n_samples, but n_samples = %d, n_neighbors = %d", n, k + 1);
    knn_features_kernel<<<cdiv(n, 4), 128, 0, (cudaStream_t)stream>>>(pts, n, k, feat);
    CT_LAUNCHED("knn_features_kernel");
    return 0;
}

extern "C" size_t ct_ffn_match_workspace_bytes(int n_ref, int n_tgt) {
    const size_t rows = (size_t)n_ref + n_tgt;
    return align_up(rows * FEAT * sizeof(float), 256) + 2 * align_up(rows * HID * sizeof(float), 256) + 512;
}

extern "C" int ct_ffn_match(const CtFFN* f, const double* ref, int N, const double* tgt, int M, int k,
                            float* corr, void* ws, size_t ws_bytes, void* stream) {
    CT_REQUIRE(f && ref && tgt && corr, "ct_ffn_match: null argument");
    CT_REQUIRE(3 * k + 1 == FEAT, "ct_ffn_match: the FFN takes %d features per point (k = 20), got k = %d", FEAT, k);
    CT_REQUIRE(ws_bytes >= ct_ffn_match_workspace_bytes(N, M), "ct_ffn_match: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    const int rows = N + M;
    ProfScope prof(PROF_FFN, s);
    float* feat = a.take<float>((size_t)rows * FEAT);
    float* F = a.take<float>((size_t)rows * HID);
    float* AB = a.take<float>((size_t)rows * HID);
    if (ct_knn_features(ref, N, k, feat, stream)) return 1;
    if (ct_knn_features(tgt, M, k, feat + (size_t)N * FEAT, stream)) return 1;
    // f(v) = leaky(bn1(v . W1)) for ref rows then tgt rows (ffn.py:261-262, shared layer)
    if (sgemm(feat, FEAT, f->W1, HID, F, HID, rows, HID, FEAT, 0, f->s1, f->t1, 1, s)) return 1;
    // A = f(ref) . W2[:512],  B = f(tgt) . W2[512:]   (ffn.py:263-264 with the concat distributed)
    if (sgemm(F, HID, f->W2, HID, AB, HID, N, HID, HID, 0, nullptr, nullptr, 0, s)) return 1;
    if (sgemm(F + (size_t)N * HID, HID, f->W2 + (size_t)HID * HID, HID, AB + (size_t)N * HID, HID, M, HID, HID, 0,
              nullptr, nullptr, 0, s)) return 1;
    dim3 grid(cdiv(N, 16), cdiv(M, 16));
    ffn_pair_kernel<<<grid, 256, 0, s>>>(AB, AB + (size_t)N * HID, N, M, f->s2, f->t2, f->w3, f->b3, corr);
    CT_LAUNCHED("ffn_pair_kernel");
    return 0;
}

extern "C" size_t ct_ffn_predict_workspace_bytes(int rows) {
    return 3 * align_up((size_t)rows * HID * sizeof(float), 256) + 512;
}

extern "C" int ct_ffn_predict(const CtFFN* f, const float* x, int rows, float* out, void* ws, size_t ws_bytes,
                              void* stream) {
    CT_REQUIRE(f && x && out, "ct_ffn_predict: null argument");
    CT_REQUIRE(ws_bytes >= ct_ffn_predict_workspace_bytes(rows), "ct_ffn_predict: workspace too small");
    if (rows == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    Arena a(ws, ws_bytes);
    float* F1 = a.take<float>((size_t)rows * HID);
    float* F2 = a.take<float>((size_t)rows * HID);
    float* H = a.take<float>((size_t)rows * HID);
    if (sgemm(x, 2 * FEAT, f->W1, HID, F1, HID, rows, HID, FEAT, 0, f->s1, f->t1, 1, s)) return 1;
    if (sgemm(x + FEAT, 2 * FEAT, f->W1, HID, F2, HID, rows, HID, FEAT, 0, f->s1, f->t1, 1, s)) return 1;
    if (sgemm(F1, HID, f->W2, HID, H, HID, rows, HID, HID, 0, nullptr, nullptr, 0, s)) return 1;
    if (sgemm(F2, HID, f->W2 + (size_t)HID * HID, HID, H, HID, rows, HID, HID, 1, f->s2, f->t2, 1, s)) return 1;
    ffn_head_kernel<<<cdiv(rows, 8), 256, 0, s>>>(H, rows, f->w3, f->b3, out);
    CT_LAUNCHED("ffn_head_kernel");
    return 0;
}
