// PR-GLS / CPD EM for LARGE point sets: the same iteration as prgls.cu (track.py:11-114, trackerlite.py:262-417), spread
// over the whole GPU as a sequence of kernels instead of one persistent CTA.
//
// prgls.cu keeps a problem inside one CTA, which is the right shape for N ~ 100-200 cells (latency bound, batched over
// ensemble members) but walks off a cliff once the N x N system leaves shared memory: one SM then runs an O(N^3) fp64
// elimination out of L2 (2.8 ms per iteration at N = 512, ~0.7 s at N = 2048 -- slower than LAPACK on the host).
// Config 3 of BASELINE.json has ~2048 cells per volume.  Here every phase of an iteration is a grid-wide kernel:
//   E-step            one warp per target row (posterior row + its normaliser)           HBM bound: M x N fp64 read + write
//   column moments    one thread per reference point, rows in a fixed order              HBM bound: M x N read
//   assembly          S = diag(p) G + lambda sigma^2 I, right-hand sides appended        HBM bound: N x N read + write
//   solve             blocked right-looking elimination WITHOUT row exchanges (see prgls.cu for why that is stable):
//                     per 32-column panel: diagonal block (one CTA), L21 = S21 U11^-1 (one thread per row -- no pivoting
//                     means the rows are independent), U12 = L11^-1 S12 (one thread per column), trailing update
//                     S22 -= L21 U12 as a register-tiled FP64 GEMM over all SMs
//                     (2/3 N^3 flops, FP64-pipe bound); blocked back substitution for the 3 right-hand sides
//   apply / scalars   move = G W, gamma, sigma^2, convergence                            HBM bound: N x N + M x N read
// Scalars (sigma^2, gamma, convergence flag) live in device memory; the host only enqueues.  Every reduction has a fixed
// shape, so results do not depend on timing.  Results agree with the single-CTA kernel / the reference to rounding
// (tests/test_gpu_ffn_prgls.py: 1e-8 on T_X at N = 300 .. 2048).
#include "common.cuh"
#include <vector>

namespace ct {

constexpr int GE_NB = 32;                         // panel width of the elimination
constexpr int GE_TILE = 128;                      // output tile of the trailing update (128 x 128, K = GE_NB)

struct GridProblem {
    const double* X; const double* Y; const double* tracked;      // (N,3) (M,3) (L,3)
    const double* prior;                                          // (M,N) fp64 (greedy prior or converted corr)
    double* P;                                                    // (M,N) posterior (output)
    double* gram;                                                 // (N,N)
    double* gram_nl;                                              // (N,L)
    double* S;                                                    // (N, ld) system + 3 right-hand sides
    double* colsum; double* ytp; double* cur; double* cur_l; double* W;      // (N) (N,3) (N,3) (L,3) (N,3)
    double* rowq;                                                 // (M) per-row partial of sum P d^2
    double* part;                                                 // (1024) block partials
    double* ref_out; double* coef; double* tracked_out; int* iterations;
    int N, M, L, ld;
    int lite, prior_f32;
    double beta, lambda, vol;
};
struct GridState {                                // device-resident scalars of one problem
    double sigma2, gamma, sumP, move2;
    int it, done, pad0, pad1;
};

__device__ __forceinline__ double g_dist2(const double* a, const double* b) {
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ double g_block_sum_256(double v, double* sh) {      // fixed-shape tree, 256 threads
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    double t = (threadIdx.x < 8) ? sh[threadIdx.x] : 0.0;
    if (w == 0) {
        t = warp_sum(t);
        if (lane == 0) sh[8] = t;
    }
    __syncthreads();
    return sh[8];
}

// ---- set-up: Gram matrices, start positions, sigma^2 (track.py:44-56 / trackerlite.py:319-323)
__global__ void __launch_bounds__(256) ge_gram(GridProblem p) {
    const double two_b2 = 2.0 * p.beta * p.beta;
    const size_t nn = (size_t)p.N * p.N, nl = (size_t)p.N * p.L;
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < nn + nl; e += (size_t)gridDim.x * 256) {
        if (e < nn) {
            const int i = (int)(e / p.N), j = (int)(e % p.N);
            p.gram[e] = exp(-g_dist2(p.X + 3 * j, p.X + 3 * i) / two_b2);
        } else {
            const size_t f = e - nn;
            const int n = (int)(f / p.L), l = (int)(f % p.L);
            p.gram_nl[f] = exp(-g_dist2(p.tracked + 3 * l, p.X + 3 * n) / two_b2);
        }
    }
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < (size_t)3 * p.N; e += (size_t)gridDim.x * 256) p.cur[e] = p.X[e];
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < (size_t)3 * p.L; e += (size_t)gridDim.x * 256) p.cur_l[e] = p.tracked[e];
}
// sum over all pairs of |x_n - y_m|^2: block b sums rows m = b, b + 1024, ... ; partials combined by ge_init_state
__global__ void __launch_bounds__(256) ge_sigma0(GridProblem p) {
    __shared__ double sh[9];
    double acc = 0.0;
    for (int m = blockIdx.x; m < p.M; m += gridDim.x)
        for (int n = threadIdx.x; n < p.N; n += 256) acc += g_dist2(p.X + 3 * n, p.Y + 3 * m);
    const double t = g_block_sum_256(acc, sh);
    if (threadIdx.x == 0) p.part[blockIdx.x] = t;
}
__global__ void __launch_bounds__(256) ge_init_state(GridProblem p, GridState* st, int nparts) {
    __shared__ double sh[9];
    double acc = 0.0;
    for (int i = threadIdx.x; i < nparts; i += 256) acc += p.part[i];
    const double t = g_block_sum_256(acc, sh);
    if (threadIdx.x == 0) {
        st->sigma2 = p.lite ? (t / ((double)p.M * (double)p.N)) / 3.0 : t / (3.0 * (double)p.N * (double)p.M);
        st->gamma = p.lite ? 0.05 : 0.1;
        st->it = 0; st->done = 0; st->move2 = 0.0; st->sumP = 0.0;
    }
}

// ---- E-step (track.py:81-88 / trackerlite.py:375-382): one warp per target row
__global__ void __launch_bounds__(256) ge_estep(GridProblem p, const GridState* st) {
    if (st->done) return;
    const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= p.M) return;
    const double PI = 3.141592653589793;
    const double sigma2 = st->sigma2, gamma = st->gamma;
    const double neg_inv_two_s2 = -1.0 / (2.0 * sigma2);
    const double norm15 = pow(2.0 * PI * sigma2, 1.5);
    const double outlier = p.lite ? gamma / p.vol : gamma * norm15 / ((1.0 - gamma) * p.vol);
    const double one_m_g = 1.0 - gamma;
    const double y0 = p.Y[3 * m], y1 = p.Y[3 * m + 1], y2 = p.Y[3 * m + 2];
    double* row = p.P + (size_t)m * p.N;
    const double* pri = p.prior + (size_t)m * p.N;
    double rs = 0.0;
    for (int n = lane; n < p.N; n += 32) {
        const double dx = p.cur[3 * n] - y0, dy = p.cur[3 * n + 1] - y1, dz = p.cur[3 * n + 2] - y2;
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        const double like = exp(d2 * neg_inv_two_s2);
        const double pr = pri[n];
        const double w1 = p.prior_f32 ? (double)((float)one_m_g * (float)pr) : one_m_g * pr;
        const double v = p.lite ? w1 * like / norm15 : pr * like;
        row[n] = v;
        rs += v;
    }
    rs = warp_sum(rs);
    const double rden = 1.0 / (rs + outlier);
    for (int n = lane; n < p.N; n += 32) row[n] *= rden;
}

// ---- column moments p_n = sum_m P[m,n], ytp_n = sum_m P[m,n] Y[m]: one thread per n, rows in ascending order
__global__ void __launch_bounds__(128) ge_moments(GridProblem p, const GridState* st) {
    if (st->done) return;
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= p.N) return;
    double c0 = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int m = 0; m < p.M; ++m) {
        const double pv = p.P[(size_t)m * p.N + n];
        c0 += pv;
        a0 = fma(pv, p.Y[3 * m], a0); a1 = fma(pv, p.Y[3 * m + 1], a1); a2 = fma(pv, p.Y[3 * m + 2], a2);
    }
    p.colsum[n] = c0;
    p.ytp[3 * n] = a0; p.ytp[3 * n + 1] = a1; p.ytp[3 * n + 2] = a2;
}

// ---- assembly (track.py:91-96 / trackerlite.py:411-415): S[i][j] = p_i G[i][j] + lambda sigma^2 [i == j], rhs appended
__global__ void __launch_bounds__(256) ge_assemble(GridProblem p, const GridState* st) {
    if (st->done) return;
    const double reg = p.lambda * st->sigma2;
    const int i = blockIdx.x;
    const double pi_ = p.colsum[i];
    const double* g = p.gram + (size_t)i * p.N;
    double* srow = p.S + (size_t)i * p.ld;
    for (int j = threadIdx.x; j < p.N; j += 256) {
        double v = g[j] * pi_;
        if (i == j) v += reg;
        srow[j] = v;
    }
    if (threadIdx.x < 3) {
        const int e = 3 * i + threadIdx.x;
        srow[p.N + threadIdx.x] = p.ytp[e] - (p.lite ? p.cur[e] : p.X[e]) * pi_;
    }
}

// ---- elimination, panel [k0, k0 + nb).  Without row exchanges the rows below the diagonal block do not interact
// inside a panel: factor the nb x nb diagonal block (one warp-sized job), then every row below is an independent
// triangular solve L21[r,:] = S21[r,:] U11^-1 (one thread per row), and U12 = L11^-1 S12 one thread per column.
__global__ void __launch_bounds__(GE_NB * GE_NB) ge_diag(GridProblem p, const GridState* st, int k0, int nb) {
    if (st->done) return;
    __shared__ double A[GE_NB][GE_NB + 1];
    const int r = threadIdx.x / GE_NB, c = threadIdx.x % GE_NB;
    if (r < nb && c < nb) A[r][c] = p.S[(size_t)(k0 + r) * p.ld + k0 + c];
    __syncthreads();
    for (int kk = 0; kk < nb; ++kk) {
        if (r > kk && r < nb && c == kk) A[r][kk] = A[r][kk] / A[kk][kk];
        __syncthreads();
        if (r > kk && r < nb && c > kk && c < nb) A[r][c] = fma(-A[r][kk], A[kk][c], A[r][c]);
        __syncthreads();
    }
    if (r < nb && c < nb) p.S[(size_t)(k0 + r) * p.ld + k0 + c] = A[r][c];
}
__global__ void __launch_bounds__(128) ge_colblock(GridProblem p, const GridState* st, int k0, int nb) {
    if (st->done) return;
    __shared__ double U11[GE_NB][GE_NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += 128) U11[e / nb][e % nb] = p.S[(size_t)(k0 + e / nb) * p.ld + k0 + e % nb];
    __syncthreads();
    const int r = k0 + nb + blockIdx.x * 128 + threadIdx.x;
    if (r >= p.N) return;
    double* row = p.S + (size_t)r * p.ld + k0;
    double l[GE_NB];
#pragma unroll
    for (int c = 0; c < GE_NB; ++c) l[c] = c < nb ? row[c] : 0.0;
#pragma unroll
    for (int c = 0; c < GE_NB; ++c) {
        if (c < nb) {
            double v = l[c];
#pragma unroll
            for (int k = 0; k < GE_NB; ++k) if (k < c) v = fma(-l[k], U11[k][c], v);
            l[c] = v / U11[c][c];
        }
    }
#pragma unroll
    for (int c = 0; c < GE_NB; ++c) if (c < nb) row[c] = l[c];
}
// U12 = L11^-1 S[k0 : k0+nb, k0+nb : cols): one thread per column
__global__ void __launch_bounds__(128) ge_rowblock(GridProblem p, const GridState* st, int k0, int nb, int cols) {
    if (st->done) return;
    __shared__ double L11[GE_NB][GE_NB + 1];
    for (int e = threadIdx.x; e < nb * nb; e += 128) L11[e / nb][e % nb] = p.S[(size_t)(k0 + e / nb) * p.ld + k0 + e % nb];
    __syncthreads();
    const int c = k0 + nb + blockIdx.x * 128 + threadIdx.x;
    if (c >= cols) return;
    double u[GE_NB];
#pragma unroll
    for (int r = 0; r < GE_NB; ++r) u[r] = r < nb ? p.S[(size_t)(k0 + r) * p.ld + c] : 0.0;
#pragma unroll
    for (int kk = 0; kk < GE_NB; ++kk)
#pragma unroll
        for (int r = kk + 1; r < GE_NB; ++r)
            if (r < nb) u[r] = fma(-L11[r][kk], u[kk], u[r]);
#pragma unroll
    for (int r = 0; r < GE_NB; ++r) if (r < nb) p.S[(size_t)(k0 + r) * p.ld + c] = u[r];
}
// S22 -= L21 U12: 128 x 128 tile per CTA (256 threads, 8 x 8 outputs each), K = nb in two chunks of 16.
// 64 DFMA per 16 shared-memory doubles read: the FP64 pipe, not shared memory, is the limit (a 4 x 4 register tile was
// shared-memory bound at 0.7 TFLOP/s).  Thread (ty, tx) owns rows ty + 16 i and columns tx + 16 j: bank-conflict free.
__global__ void __launch_bounds__(256) ge_update(GridProblem p, const GridState* st, int k0, int nb, int cols) {
    if (st->done) return;
    constexpr int KC = 16;
    __shared__ double Ls[GE_TILE][KC + 1];
    __shared__ double Us[KC][GE_TILE + 1];
    const int r0 = k0 + nb + blockIdx.y * GE_TILE, c0 = k0 + nb + blockIdx.x * GE_TILE;
    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    double acc[8][8] = {};
    for (int kc = 0; kc < nb; kc += KC) {
        for (int e = threadIdx.x; e < GE_TILE * KC; e += 256) {
            const int r = e / KC, k = e % KC;
            Ls[r][k] = (r0 + r < p.N && kc + k < nb) ? p.S[(size_t)(r0 + r) * p.ld + k0 + kc + k] : 0.0;
        }
        for (int e = threadIdx.x; e < KC * GE_TILE; e += 256) {
            const int k = e / GE_TILE, c = e % GE_TILE;
            Us[k][c] = (c0 + c < cols && kc + k < nb) ? p.S[(size_t)(k0 + kc + k) * p.ld + c0 + c] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < KC; ++k) {
            double l[8], u[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { l[i] = Ls[ty + 16 * i][k]; u[i] = Us[k][tx + 16 * i]; }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fma(l[i], u[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int r = r0 + ty + 16 * i, c = c0 + tx + 16 * j;
            if (r < p.N && c < cols) p.S[(size_t)r * p.ld + c] -= acc[i][j];
        }
}
// back substitution, block [k0, k0 + nb): solve the upper-triangular diagonal block for the 3 right-hand sides ...
__global__ void __launch_bounds__(32) ge_back_diag(GridProblem p, const GridState* st, int k0, int nb) {
    if (st->done) return;
    const int d = threadIdx.x;                                    // lanes 0..2: one right-hand side each
    if (d >= 3) return;
    for (int r = nb - 1; r >= 0; --r) {
        const double* row = p.S + (size_t)(k0 + r) * p.ld;
        double v = row[p.N + d];
        for (int c = r + 1; c < nb; ++c) v = fma(-row[k0 + c], p.W[3 * (k0 + c) + d], v);
        p.W[3 * (k0 + r) + d] = v / row[k0 + r];
    }
}
// ... and remove its contribution from the right-hand sides of the rows above
__global__ void __launch_bounds__(128) ge_back_update(GridProblem p, const GridState* st, int k0, int nb) {
    if (st->done) return;
    const int r = blockIdx.x * 128 + threadIdx.x;
    if (r >= k0) return;
    double* row = p.S + (size_t)r * p.ld;
    double v0 = row[p.N], v1 = row[p.N + 1], v2 = row[p.N + 2];
    for (int c = 0; c < nb; ++c) {
        const double u = row[k0 + c];
        v0 = fma(-u, p.W[3 * (k0 + c)], v0); v1 = fma(-u, p.W[3 * (k0 + c) + 1], v1); v2 = fma(-u, p.W[3 * (k0 + c) + 2], v2);
    }
    row[p.N] = v0; row[p.N + 1] = v1; row[p.N + 2] = v2;
}

// ---- apply: move = G W (track.py:100 / trackerlite.py:337-341); one warp per reference point
__global__ void __launch_bounds__(256) ge_apply(GridProblem p, const GridState* st) {
    if (st->done) return;
    const int lane = threadIdx.x & 31, i = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (i >= p.N) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    const double* g = p.gram + (size_t)i * p.N;
    for (int j = lane; j < p.N; j += 32) {
        const double gv = g[j];
        a0 = fma(gv, p.W[3 * j], a0); a1 = fma(gv, p.W[3 * j + 1], a1); a2 = fma(gv, p.W[3 * j + 2], a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) {
        if (p.lite) {
            p.rowq[p.M + i] = a0 * a0 + a1 * a1 + a2 * a2;                  // per-point |move|^2, summed by ge_scalars
            if (st->it + 1 > 1) { p.cur[3 * i] += a0; p.cur[3 * i + 1] += a1; p.cur[3 * i + 2] += a2; }
        } else {
            p.cur[3 * i] = p.X[3 * i] + a0; p.cur[3 * i + 1] = p.X[3 * i + 1] + a1; p.cur[3 * i + 2] = p.X[3 * i + 2] + a2;
        }
    }
}
__global__ void __launch_bounds__(128) ge_apply_tracked(GridProblem p, const GridState* st) {
    if (st->done || st->it + 1 <= 1) return;
    const int e = blockIdx.x * 128 + threadIdx.x;
    if (e >= 3 * p.L) return;
    const int l = e / 3, dim = e % 3;
    double a = 0.0;
    for (int n = 0; n < p.N; ++n) a = fma(p.gram_nl[(size_t)n * p.L + l], p.W[3 * n + dim], a);
    p.cur_l[e] += a;
}
// per-row partial of sum_mn P[m,n] |cur_n - y_m|^2 (track.py:106-110 / trackerlite.py:348-350)
__global__ void __launch_bounds__(256) ge_rowq(GridProblem p, const GridState* st) {
    if (st->done) return;
    const int lane = threadIdx.x & 31, m = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (m >= p.M) return;
    const double y0 = p.Y[3 * m], y1 = p.Y[3 * m + 1], y2 = p.Y[3 * m + 2];
    const double* row = p.P + (size_t)m * p.N;
    double q = 0.0;
    for (int n = lane; n < p.N; n += 32) {
        const double dx = p.cur[3 * n] - y0, dy = p.cur[3 * n + 1] - y1, dz = p.cur[3 * n + 2] - y2;
        const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
        q = fma(row[n], d2, q);
    }
    q = warp_sum(q);
    if (lane == 0) p.rowq[m] = q;
}
// gamma, sigma^2, iteration count, convergence (track.py:103-112 / trackerlite.py:342-356): one CTA, fixed order
__global__ void __launch_bounds__(256) ge_scalars(GridProblem p, GridState* st) {
    if (st->done) return;
    __shared__ double sh[9];
    double cs = 0.0, q = 0.0, mv = 0.0;
    for (int n = threadIdx.x; n < p.N; n += 256) cs += p.colsum[n];
    for (int m = threadIdx.x; m < p.M; m += 256) q += p.rowq[m];
    if (p.lite) for (int n = threadIdx.x; n < p.N; n += 256) mv += p.rowq[p.M + n];
    const double sumP = g_block_sum_256(cs, sh);
    const double qs = g_block_sum_256(q, sh);
    const double move2 = g_block_sum_256(mv, sh);
    if (threadIdx.x == 0) {
        double gamma = 1.0 - sumP / (double)p.M;
        if (p.lite && gamma < 1e-4) gamma = 1e-4;
        double s2 = qs / (3.0 * sumP);
        if (!p.lite && s2 < 1.0) s2 = 1.0;
        st->gamma = gamma; st->sigma2 = s2; st->sumP = sumP; st->move2 = move2;
        st->it += 1;
        if (p.lite && sqrt(move2) < 1e-3) st->done = 1;
    }
}
__global__ void __launch_bounds__(256) ge_outputs(GridProblem p, const GridState* st) {
    for (int e = blockIdx.x * 256 + threadIdx.x; e < 3 * p.N; e += gridDim.x * 256) {
        if (p.ref_out) p.ref_out[e] = p.cur[e];
        if (p.coef) p.coef[(size_t)(e % 3) * p.N + e / 3] = st->it > 0 ? p.W[e] : 0.0;
    }
    if (p.lite && p.tracked_out)
        for (int e = blockIdx.x * 256 + threadIdx.x; e < 3 * p.L; e += gridDim.x * 256) p.tracked_out[e] = p.cur_l[e];
    if (blockIdx.x == 0 && threadIdx.x == 0 && p.iterations) *p.iterations = st->it;
}
__global__ void __launch_bounds__(256) ge_convert_prior(const void* corr, int is_f64, double* prior, size_t n) {
    for (size_t e = (size_t)blockIdx.x * 256 + threadIdx.x; e < n; e += (size_t)gridDim.x * 256)
        prior[e] = is_f64 ? static_cast<const double*>(corr)[e] : (double)static_cast<const float*>(corr)[e];
}

// ---------------------------------------------------------------------------------------------
// host side: workspace layout and the launch sequence of one problem
// ---------------------------------------------------------------------------------------------
static inline int ge_ld(int N) { return (N + 3 + 7) / 8 * 8; }

size_t grid_em_workspace_bytes(int N, int M, int L) {
    size_t t = 0;
    auto take = [&](size_t b) { t += align_up(b, 256); };
    take((size_t)N * N * 8); take((size_t)N * (L > 0 ? L : 1) * 8); take((size_t)N * ge_ld(N) * 8);
    take((size_t)M * N * 8);                                        // prior
    take((size_t)N * 8); take((size_t)N * 24); take((size_t)N * 24); take((size_t)(L > 0 ? L : 1) * 24); take((size_t)N * 24);
    take((size_t)(M + N) * 8); take(1024 * 8); take(sizeof(GridState)); take(256);
    return t + 256;
}

// Runs one problem on the whole GPU.  `prior_ready`: the (M,N) fp64 prior has already been written to the workspace's
// prior slot by the caller (greedy kernel) -- returns the slot through *prior_slot when called with run = false.
int grid_em_run(const CtPrglsParams& prm, const CtPrglsProblem& q, void* ws, size_t ws_bytes, double** prior_slot, bool run,
                cudaStream_t s, unsigned long long* launches) {
    const int N = q.n_ref, M = q.n_tgt, L = prm.mode == CT_PRGLS_LITE ? q.n_tracked : 0;
    CT_REQUIRE(ws_bytes >= grid_em_workspace_bytes(N, M, L), "ct_prgls: workspace too small for the grid path");
    CT_REQUIRE(N <= 4096, "ct_prgls: at most 4096 reference points are supported, got %d", N);
    char* base = reinterpret_cast<char*>(((uintptr_t)ws + 255) / 256 * 256);
    auto take = [&](size_t b) { char* r = base; base += align_up(b, 256); return r; };
    GridProblem p{};
    p.X = q.ref; p.Y = q.tgt; p.tracked = q.tracked; p.P = q.post;
    p.gram = (double*)take((size_t)N * N * 8);
    p.gram_nl = (double*)take((size_t)N * (L > 0 ? L : 1) * 8);
    p.ld = ge_ld(N);
    p.S = (double*)take((size_t)N * p.ld * 8);
    double* prior = (double*)take((size_t)M * N * 8);
    p.prior = prior;
    p.colsum = (double*)take((size_t)N * 8); p.ytp = (double*)take((size_t)N * 24); p.cur = (double*)take((size_t)N * 24);
    p.cur_l = (double*)take((size_t)(L > 0 ? L : 1) * 24); p.W = (double*)take((size_t)N * 24);
    p.rowq = (double*)take((size_t)(M + N) * 8); p.part = (double*)take(1024 * 8);
    GridState* st = (GridState*)take(sizeof(GridState));
    p.ref_out = q.ref_out; p.coef = q.coef; p.tracked_out = q.tracked_out; p.iterations = q.iterations;
    p.N = N; p.M = M; p.L = L; p.lite = prm.mode == CT_PRGLS_LITE; p.prior_f32 = p.lite && !q.corr_is_f64;
    p.beta = prm.beta; p.lambda = prm.lambda; p.vol = prm.vol;
    if (prior_slot) *prior_slot = prior;
    if (!run) return 0;
    unsigned long long n = 0;
#define GE_LAUNCH(k, grid, block, ...) do { k<<<grid, block, 0, s>>>(__VA_ARGS__); ++n; } while (0)
    if (q.prior_given) GE_LAUNCH(ge_convert_prior, 296, 256, q.corr, q.corr_is_f64, prior, (size_t)M * N);
    GE_LAUNCH(ge_gram, 592, 256, p);
    const int nparts = M < 1024 ? M : 1024;
    GE_LAUNCH(ge_sigma0, nparts, 256, p);
    GE_LAUNCH(ge_init_state, 1, 256, p, st, nparts);
    const int cols = N + 3;
    int* done_host = nullptr;
    if (p.lite) CT_CUDA(cudaMallocHost(&done_host, sizeof(int)));
    for (int it = 1; it < prm.max_iteration; ++it) {
        GE_LAUNCH(ge_estep, cdiv(M, 8), 256, p, st);
        GE_LAUNCH(ge_moments, cdiv(N, 128), 128, p, st);
        GE_LAUNCH(ge_assemble, N, 256, p, st);
        for (int k0 = 0; k0 < N; k0 += GE_NB) {
            const int nb = N - k0 < GE_NB ? N - k0 : GE_NB;
            GE_LAUNCH(ge_diag, 1, GE_NB * GE_NB, p, st, k0, nb);
            if (N - (k0 + nb) > 0) GE_LAUNCH(ge_colblock, cdiv(N - (k0 + nb), 128), 128, p, st, k0, nb);
            const int rest = cols - (k0 + nb);
            if (rest > 0) GE_LAUNCH(ge_rowblock, cdiv(rest, 128), 128, p, st, k0, nb, cols);
            const int rrows = N - (k0 + nb);
            if (rrows > 0) GE_LAUNCH(ge_update, dim3(cdiv(rest, GE_TILE), cdiv(rrows, GE_TILE)), 256, p, st, k0, nb, cols);
        }
        for (int k0 = (N - 1) / GE_NB * GE_NB; k0 >= 0; k0 -= GE_NB) {
            const int nb = N - k0 < GE_NB ? N - k0 : GE_NB;
            GE_LAUNCH(ge_back_diag, 1, 32, p, st, k0, nb);
            if (k0 > 0) GE_LAUNCH(ge_back_update, cdiv(k0, 128), 128, p, st, k0, nb);
        }
        GE_LAUNCH(ge_apply, cdiv(N, 8), 256, p, st);
        if (p.lite && L > 0) GE_LAUNCH(ge_apply_tracked, cdiv(3 * L, 128), 128, p, st);
        GE_LAUNCH(ge_rowq, cdiv(M, 8), 256, p, st);
        GE_LAUNCH(ge_scalars, 1, 256, p, st);
        if (p.lite && (it % 4 == 0)) {                           // the LITE loop stops on convergence (trackerlite.py:353-356)
            CT_CUDA(cudaMemcpyAsync(done_host, &st->done, sizeof(int), cudaMemcpyDeviceToHost, s));
            CT_CUDA(cudaStreamSynchronize(s));
            if (*done_host) break;
        }
    }
    GE_LAUNCH(ge_outputs, 8, 256, p, st);
#undef GE_LAUNCH
    if (done_host) cudaFreeHost(done_host);
    if (launches) *launches += n;
    CT_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace ct
