// Greedy prior + PR-GLS / coherent-point-drift EM as ONE persistent CTA per problem (fp64).
//
// Reference: track.py:11-114 (pr_gls_quick), trackerlite.py:242-259 (simple_match), :262-358
// (prgls_quick / prgls_with_two_ref), :361-382 (dist_squares, gaussian_kernel, estimate_posterior),
// :409-417 (solve_movements_ref), tracker.py:1269-1289 (_predict_one_rep).
//
// The reference drives every EM iteration from Python (about ten NumPy dispatches building (M,N,3)
// temporaries plus one LAPACK gesv).  Here a whole problem -- greedy prior, Gram matrix, all EM
// iterations with the convergence test -- runs inside one 1024-thread CTA with no host round trip;
// a batch (ensemble members x repetitions) is one launch with one CTA per problem.  The N x N system
//      (G diag(p) + lambda sigma^2 I)^T C^T = (Y^T P - X^T diag(p))^T
// is non-symmetric; the reference solves it with LAPACK gesv (track.py:97).  Here it is eliminated WITHOUT row
// exchanges (blocked right-looking LU, see lu_solve below: S^T is a positive row scaling of an SPD matrix, so the
// unpivoted elimination is backward stable while lambda sigma^2 > 0 -- ct_prgls rejects lambda <= 0 and the kernel
// raises an error flag on a non-finite solution).  The system matrix and the five per-point vectors
// live in shared memory when they fit (N <= 164: 8 N (N|1) + 88 N bytes <= 226 KiB), otherwise the matrix
// moves to the L2-resident workspace (vectors stay in shared memory up to N ~ 2600).
#include "common.cuh"
#include <vector>

namespace ct {

constexpr int EM_THREADS = 1024;
constexpr int EM_WARPS = EM_THREADS / 32;

struct DevProblem {
    CtPrglsProblem p;
    double* gram;       // (N,N)
    double* gram_nl;    // (N,L)   LITE
    double* sys;        // (N,N)   only when N > SMEM_N_MAX
    double* prior;      // (M,N)
    double* colsum;     // (N)     p_n
    double* ytp;        // (N,3)   sum_m P[m,n] Y[m]
    double* colpart;    // (R*N,4) partial column moments, R*N <= max(N, 1024)
    double* rhs;        // (N,3)   right-hand side, overwritten by the solution W = C^T
    double* cur;        // (N,3)   T_X / predicted ref
    double* cur_l;      // (L,3)   predicted tracked
    double* rowmax;     // (M)     greedy scratch
    int* rowarg;        // (M)
    int* match;         // (M)     matched ref index per target row or -1
    unsigned char* col_dead;  // (N)
    int* pairs;         // (min(M,N), 2)
    int* n_pairs;       // (1)
};

struct Scratch {
    double red[EM_WARPS];
    int redi[EM_WARPS];
    double bc[4];
    int bci[4];
};

__device__ __forceinline__ double block_sum(double v, Scratch& s) {
    v = warp_sum(v);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s.red[w] = v;
    __syncthreads();
    double t = (threadIdx.x < EM_WARPS) ? s.red[threadIdx.x] : 0.0;
    if (w == 0) {
        t = warp_sum(t);
        if (lane == 0) s.bc[0] = t;
    }
    __syncthreads();
    return s.bc[0];
}

// argmax with "first occurrence" tie-break (smallest index among equal values)
__device__ __forceinline__ void better(double& v, int& i, double ov, int oi) {
    if (ov > v || (ov == v && oi < i)) { v = ov; i = oi; }
}
__device__ __forceinline__ void warp_argmax(double& v, int& i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        better(v, i, ov, oi);
    }
}
__device__ __forceinline__ void block_argmax(double& v, int& i, Scratch& s) {
    warp_argmax(v, i);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { s.red[w] = v; s.redi[w] = i; }
    __syncthreads();
    if (w == 0) {
        double tv = s.red[lane];
        int ti = s.redi[lane];
        warp_argmax(tv, ti);
        if (lane == 0) { s.bc[0] = tv; s.bci[0] = ti; }
    }
    __syncthreads();
    v = s.bc[0]; i = s.bci[0];
}

__device__ __forceinline__ double corr_at(const DevProblem& d, int m, int n) {
    const size_t idx = (size_t)m * d.p.n_ref + n;
    return d.p.corr_is_f64 ? static_cast<const double*>(d.p.corr)[idx]
                           : (double)static_cast<const float*>(d.p.corr)[idx];
}

__device__ __forceinline__ double dist2(const double* a, const double* b) {
    const double dx = a[0] - b[0], dy = a[1] - b[1], dz = a[2] - b[2];
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------------------------
// greedy one-to-one assignment -> prior.   track.py:58-70 / trackerlite.py:242-259
// Row maxima are cached; a round is: argmax over the M cached maxima, retire row m* and column n*,
// rescan only the rows whose cached argmax was n* (zeroed entries count as 0, like the reference).
// ---------------------------------------------------------------------------------------------
__device__ void scan_row(const DevProblem& d, int m, bool row_dead, int lane) {
    const int N = d.p.n_ref;
    double v = -INFINITY;
    int i = 0x7fffffff;
    for (int n = lane; n < N; n += 32) {
        const double c = (row_dead || d.col_dead[n]) ? 0.0 : corr_at(d, m, n);
        better(v, i, c, n);
    }
    warp_argmax(v, i);
    if (lane == 0) { d.rowmax[m] = v; d.rowarg[m] = i; }
}

__device__ void greedy_prior(const DevProblem& d, int mode, double threshold, Scratch& s) {
    const int N = d.p.n_ref, M = d.p.n_tgt;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    for (int n = tid; n < N; n += EM_THREADS) d.col_dead[n] = 0;
    for (int m = tid; m < M; m += EM_THREADS) d.match[m] = -1;
    __syncthreads();
    for (int m = w; m < M; m += EM_WARPS) scan_row(d, m, false, lane);
    __syncthreads();
    int n_pairs = 0;
    for (int round = 0; round < N; ++round) {
        double v = -INFINITY;
        int i = 0x7fffffff;
        for (int m = tid; m < M; m += EM_THREADS) better(v, i, d.rowmax[m], m);
        // ties between rows: smallest m wins, which is the row-major first maximum of np.argmax
        block_argmax(v, i, s);
        if (v < threshold) break;
        const int ms = i, ns = d.rowarg[ms];
        __syncthreads();
        if (tid == 0) {
            d.match[ms] = ns;
            d.col_dead[ns] = 1;
            if (d.pairs) { d.pairs[2 * n_pairs] = ms; d.pairs[2 * n_pairs + 1] = ns; }
        }
        ++n_pairs;
        __syncthreads();
        // retired row: all zeros.  Rows that pointed at column ns (or whose maximum is <= 0, for which a
        // newly zeroed entry at a smaller index could now be the first maximum) are rescanned.
        for (int m = w; m < M; m += EM_WARPS) {
            const bool dead = d.match[m] >= 0;
            if (m == ms || (!dead && (d.rowarg[m] == ns || d.rowmax[m] <= 0.0))) scan_row(d, m, dead, lane);
        }
        __syncthreads();
    }
    if (tid == 0 && d.n_pairs) *d.n_pairs = n_pairs;
    // fill the prior
    const double lo64 = 0.1 / (double)(N - 1);
    double lo, hi, unmatched;
    if (mode == CT_PRGLS_LITE) {
        // np.full_like(match_matrix, 0.1/(N-1)) keeps corr's dtype (trackerlite.py:256)
        lo = d.p.corr_is_f64 ? lo64 : (double)(float)lo64;
        hi = d.p.corr_is_f64 ? 0.9 : (double)0.9f;
        unmatched = lo;
    } else {
        lo = lo64; hi = 0.9; unmatched = 1.0 / (double)N;
    }
    for (size_t e = tid; e < (size_t)M * N; e += EM_THREADS) {
        const int m = (int)(e / N), n = (int)(e % N);
        const int mt = d.match[m];
        d.prior[e] = (mt < 0) ? unmatched : (n == mt ? hi : lo);
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Solve of the M-step system, S (N x N) with three right-hand sides stored as columns N..N+2 of the same
// row-major array (leading dimension ld, odd in shared memory so row walks are bank-conflict free).
//
//      S = a^T = diag(p) G + lambda sigma^2 I = diag(p) (G + lambda sigma^2 diag(p)^-1)
//
// is a positive row scaling of a symmetric positive definite matrix (G is a Gaussian Gram matrix), so Gaussian
// elimination needs no row exchanges to be backward stable (the pivots are the row scale times the pivots of the
// SPD factor).  The reference calls LAPACK gesv (partial pivoting, track.py:97 / trackerlite.py:416); both are
// backward-stable solves of the same system and agree to rounding (measured <= 4e-13 relative on C at
// cond(S) = 1e4, worm4 parameters), which is what the parity tests pin.  Dropping the pivot search removes the
// serial argmax + row swap from all N elimination steps.
//
// Right-looking blocked elimination, panel width 8, "row-scaled" form S = L U' with U' unit upper triangular
// (every pivot row is divided by its pivot as soon as it is final), so only the 8 x 8 diagonal blocks see
// divisions, and the right-hand sides -- being columns of S -- come out forward-substituted AND scaled.
// Per panel [k0, k0+8):
//   diag8   warp 0 eliminates the 8 x 8 diagonal block in registers (lane = (row, column pair), shuffles
//           broadcast the pivot row).  This chain of 8 reciprocals per panel is the only serial part; it runs
//           one panel AHEAD: in phase (3) warp 0 updates the tile holding the next diagonal block first and
//           then factors it while the other warps finish the phase.
//   (2)     one thread per row below the panel computes its 8 entries of L (forward substitution against U11'),
//           one thread per column right of the panel its 8 entries of U' (forward substitution against L11,
//           times the pivot reciprocals) -- no divisions, no communication, stores after the last load.
//   (3)     the trailing block gets its rank-8 update on the FP64 tensor cores (mma.sync m16n8k8, K = panel
//           width): same FMA rate as DFMA on B200 (measured 63.8 vs 62.7 FMA/clk/SM, scripts/micro/fp64_rate.cu)
//           but a fraction of the shared-memory wavefronts of a register-tiled DFMA loop; rows and columns are
//           permuted inside a tile so every fragment access is bank-conflict free for any odd ld.
// Substitution (not explicit block inverses) keeps the result within ~1e-10 of LAPACK at cond(S) ~ 1e7.
// Back substitution is blocked the same way (8 x 8 unit-upper solves in one warp + parallel updates).
// The whole solve is bound by the FP64 pipe of ONE SM (2 N^3 / 3 flops at 64 FMA/clk) -- see DESIGN.md.
// ---------------------------------------------------------------------------------------------
#ifdef EM_TIMING
// phase timers: thread 0 of block 0 accumulates clock64 deltas in SHARED memory (a global read-modify-write per tick
// costs ~700 clk on the critical warp and swamps what is being measured)
__device__ long long g_em_t[16];
__shared__ long long em_t_s[16];
#define EM_TICK(k) do { if (threadIdx.x == 0 && blockIdx.x == 0) { long long now_ = clock64(); em_t_s[k] += now_ - em_t0_; em_t0_ = clock64(); } } while (0)
#define EM_TICK_DECL long long em_t0_ = clock64()
#else
#define EM_TICK(k) do { } while (0)
#define EM_TICK_DECL do { } while (0)
#endif
constexpr int LU_NB = 8;
constexpr int LU_SIDE_THREADS = 384;      // phase (2), SM path: warps 0-5 one thread per row, warps 6-11 one per column

struct LuScratch {
    double blk[LU_NB][LU_NB];              // factored diagonal block of the current panel (fixed stride: immediate offsets)
    double rinv[LU_NB];                    // its pivot reciprocals
    double x[3][LU_NB];                    // solved block of the back substitution
};

__device__ __forceinline__ void dmma_16x8x8(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
        : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}

// Warp-wide elimination of the nb x nb diagonal block at (k0, k0), padded to 8 x 8 with the identity.
// Lane = 8 q + i holds columns 2q, 2q+1 of row i.  Result in place: strictly lower = multipliers, diagonal =
// pivots, strictly upper = U11' (scaled); rinv[j] = 1 / pivot j.
__device__ __forceinline__ void diag8(double* __restrict__ S, const int ld, const int k0, const int nb,
                                      double* __restrict__ rinv, double (*__restrict__ blk)[LU_NB]) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, i = lane & 7, q = lane >> 3;
    double x[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const int ci = 2 * q + sl;
        x[sl] = (i < nb && ci < nb) ? S[(k0 + i) * ld + k0 + ci] : (ci == i ? 1.0 : 0.0);
    }
#pragma unroll 1
    for (int j = 0; j < LU_NB; ++j) {                 // rolled on purpose: the kernel must stay I-cache resident
        const int pl = 8 * (j >> 1);
        const double xs = (j & 1) ? x[1] : x[0];
        const double r = 1.0 / __shfl_sync(FULL, xs, pl + j);
        const double mult = __shfl_sync(FULL, xs, pl + i);
        if (lane == j) rinv[j] = r;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const double u = __shfl_sync(FULL, x[sl], 8 * q + j) * r;
            if (2 * q + sl > j) {
                if (i == j) x[sl] = u;
                else if (i > j) x[sl] = fma(-mult, u, x[sl]);
            }
        }
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const int ci = 2 * q + sl;
        if (i < nb && ci < nb) S[(k0 + i) * ld + k0 + ci] = x[sl];
        blk[i][ci] = x[sl];
    }
}

// One 16 x 16 tile of the rank-8 update of panel [k0, k0+8): C -= L21 U12'.
__device__ __forceinline__ void update_tile(double* __restrict__ S, const int ld, const int N, const int NC, const int k0,
                                            const int i0, const int c0) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    const int rg = 4 * (g & 3) + (g >> 2);                   // row permutation: conflict-free for any odd ld
    const int ia = i0 + rg, ib = ia + 2;
    const int ra = ((ia < N) ? ia : N - 1) * ld, rb = ((ib < N) ? ib : N - 1) * ld;
    double a[4];
    a[0] = -S[ra + k0 + t]; a[1] = -S[rb + k0 + t]; a[2] = -S[ra + k0 + t + 4]; a[3] = -S[rb + k0 + t + 4];
    const double* __restrict__ u0 = S + (k0 + t) * ld;
    const double* __restrict__ u1 = u0 + 4 * ld;
    const int nofs = (g >> 1) + 4 * (g & 1);                 // column permutation inside an 8-column block
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int cb = c0 + 8 * h;
        if (cb >= NC) break;
        const int cn = (cb + nofs < NC) ? cb + nofs : NC - 1;
        const int ca = cb + t, cbb = cb + t + 4;
        const int la = (ca < NC) ? ca : NC - 1, lb = (cbb < NC) ? cbb : NC - 1;
        double b[2], c[4];
        b[0] = u0[cn]; b[1] = u1[cn];
        c[0] = S[ra + la]; c[1] = S[ra + lb]; c[2] = S[rb + la]; c[3] = S[rb + lb];
        dmma_16x8x8(c, a, b);
        if (ia < N) {
            if (ca < NC) S[ra + ca] = c[0];
            if (cbb < NC) S[ra + cbb] = c[1];
        }
        if (ib < N) {
            if (ca < NC) S[rb + ca] = c[2];
            if (cbb < NC) S[rb + cbb] = c[3];
        }
    }
}

template <bool SM>
__device__ __forceinline__ void lu_solve(double* __restrict__ S, const int N, const int ld, double* __restrict__ sol,
                                         LuScratch& ls) {
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int NC = N + 3;
    const unsigned FULL = 0xffffffffu;
    const int P = (N + LU_NB - 1) / LU_NB;
    EM_TICK_DECL;
    for (int p = -1; p < P; ++p) {                     // p = -1: only the look-ahead factorisation of block 0
        const int k0 = p * LU_NB;
        const int nb = (N - k0 < LU_NB) ? N - k0 : LU_NB;
        const int kend = k0 + nb;
        if (p >= 0) {
            // ---------------- (2) L below the panel (thread per row), U' right of it (thread per column)
            if (!SM || tid < LU_SIDE_THREADS) {
                const int TA = SM ? LU_SIDE_THREADS / 2 : EM_THREADS;
                const bool row_role = !SM || tid < TA, col_role = !SM || tid >= TA;
                const int t = SM ? (tid < TA ? tid : tid - TA) : tid;
                // the factored diagonal block is read from ls.blk (fixed stride -> immediate offsets, no index math);
                // rows / columns of a partial last panel are padded with the identity there, so no guards are needed
                // on the arithmetic, only on the loads / stores of S
                if (row_role) {
                    for (int i = kend + t; i < N; i += TA) {
                        double* row = S + i * ld + k0;
                        double a[LU_NB];
#pragma unroll
                        for (int c = 0; c < LU_NB; ++c) a[c] = row[c];          // nb == 8 whenever rows exist below
#pragma unroll
                        for (int j = 1; j < LU_NB; ++j) {
#pragma unroll
                            for (int m = 0; m < j; ++m) a[j] = fma(-a[m], ls.blk[m][j], a[j]);
                            asm volatile("" ::: "memory");      // 64-register budget: do not hoist all 28 block loads
                        }
#pragma unroll
                        for (int j = 1; j < LU_NB; ++j) row[j] = a[j];
                    }
                }
                if (col_role) {
                    for (int c = kend + t; c < NC; c += TA) {
                        double* col = S + k0 * ld + c;
                        double a[LU_NB];
                        if (nb == LU_NB) {
#pragma unroll
                            for (int j = 0; j < LU_NB; ++j) a[j] = col[j * ld];
                        } else {
#pragma unroll
                            for (int j = 0; j < LU_NB; ++j) a[j] = (j < nb) ? col[j * ld] : 0.0;
                        }
#pragma unroll
                        for (int j = 0; j < LU_NB; ++j) {
#pragma unroll
                            for (int m = 0; m < j; ++m) a[j] = fma(-ls.blk[j][m], a[m], a[j]);
                            a[j] *= ls.rinv[j];
                            asm volatile("" ::: "memory");
                        }
                        if (nb == LU_NB) {
#pragma unroll
                            for (int j = 0; j < LU_NB; ++j) col[j * ld] = a[j];
                        } else {
#pragma unroll
                            for (int j = 0; j < LU_NB; ++j)
                                if (j < nb) col[j * ld] = a[j];
                        }
                    }
                }
            }
            __syncthreads();
            EM_TICK(5);
            if (kend >= N) break;
        }
        // ---------------- (3) trailing block -= L21 U12' (16 x 16 tiles); warp 0 looks ahead
        {
            const int RT = (N - kend + 15) >> 4, CG = (NC - kend + 15) >> 4;
            const int items = (p >= 0) ? RT * CG : 0;
            // Warp 0 takes the tile of the next diagonal block and then factors it.  The other warps of ITS scheduler
            // (w % 4 == 0) stay out of the tensor-core work: diag8 is a chain of dependent FP64 operations, and behind a
            // queue of DMMAs on the same FP64 pipe every link costs ~60 clk instead of 8 (measured 3600 vs ~1000 clk).
            constexpr int WORKERS = EM_WARPS - EM_WARPS / 4;
            if (w == 0 || (w & 3)) {
                const int wi = (w == 0) ? 0 : w - (w >> 2);           // warp 0 -> item 0 only; workers 1..WORKERS
                const int step = (w == 0) ? items : WORKERS;
                int rt = wi / CG, cg = wi - rt * CG;
                for (int item = wi; item < items; item += step) {
                    update_tile(S, ld, N, NC, k0, kend + 16 * rt, kend + 16 * cg);
                    cg += step;
                    while (cg >= CG) { cg -= CG; ++rt; }
                }
            }
            if (w == 0) {
                __syncwarp();
                EM_TICK(11);
                const int nk = (p >= 0) ? kend : 0;
                diag8(S, ld, nk, (N - nk < LU_NB) ? N - nk : LU_NB, ls.rinv, ls.blk);
                EM_TICK(12);
            }
        }
        __syncthreads();
        EM_TICK(10);
    }
    // ---------------- back substitution with the unit upper factor, blocked like the panels.
    // warp d (< 3) solves the 8 x 8 block of right-hand side d in registers; threads (i, d) then update rows above.
    for (int k0 = (P - 1) * LU_NB; k0 >= 0; k0 -= LU_NB) {
        const int nb = (N - k0 < LU_NB) ? N - k0 : LU_NB;
        if (w < 3) {
            const int li = lane & 7;
            const double* __restrict__ urow = S + (k0 + (li < nb ? li : 0)) * ld + k0;
            double y = (li < nb) ? urow[N - k0 + w] : 0.0;
#pragma unroll 1
            for (int j = nb - 1; j > 0; --j) {
                const double uj = urow[j];                       // off the critical chain
                const double xj = __shfl_sync(FULL, y, j);
                if (li < j) y = fma(-uj, xj, y);
            }
            if (lane < nb) {
                ls.x[w][lane] = y;
                sol[3 * (k0 + lane) + w] = y;
            }
        }
        if (k0 == 0) break;
        __syncthreads();
        for (int e = tid; e < 3 * k0; e += EM_THREADS) {
            const int i = e / 3, dd = e - 3 * i;
            const double* __restrict__ urow = S + i * ld + k0;
            double y = S[i * ld + N + dd];
#pragma unroll
            for (int c = 0; c < LU_NB; ++c)
                if (c < nb) y = fma(-urow[c], ls.x[dd][c], y);
            S[i * ld + N + dd] = y;
        }
        __syncthreads();
    }
    __syncthreads();
    EM_TICK(6);
}

// Shared-memory plan of one CTA (dynamic): the per-point vectors (7 N doubles) first, then the N x ld augmented
// system when it fits.  EM_SMEM_BUDGET leaves room for the static Scratch.
constexpr size_t EM_STATIC_SMEM = 3584;                    // allowance for the static Scratch + LuScratch
constexpr size_t EM_SMEM_BUDGET = 232448 - EM_STATIC_SMEM;  // 227 KiB opt-in limit minus the static part
static_assert(sizeof(Scratch) + sizeof(LuScratch) + 64 <= EM_STATIC_SMEM, "static shared memory allowance too small");
__host__ __device__ inline int sys_ld(int N) { return (N + 3) | 1; }
__host__ __device__ inline bool vec_fits(int N) { return (size_t)56 * N <= EM_SMEM_BUDGET; }
__host__ __device__ inline bool sys_fits(int N) { return (size_t)56 * N + (size_t)8 * N * sys_ld(N) <= EM_SMEM_BUDGET; }

// ---------------------------------------------------------------------------------------------
// the EM kernel
// ---------------------------------------------------------------------------------------------
// SM = true: every problem of the launch keeps its vectors and augmented system in shared memory (N <= 164).
template <bool SM>
__global__ void __launch_bounds__(EM_THREADS, 1)
prgls_kernel(const DevProblem* __restrict__ probs, CtPrglsParams prm) {
    extern __shared__ double dyn_smem[];
    __shared__ Scratch s;
    __shared__ LuScratch lus;
    const DevProblem d = probs[blockIdx.x];
    const int N = d.p.n_ref, M = d.p.n_tgt, L = d.p.n_tracked;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const bool lite = prm.mode == CT_PRGLS_LITE;
    const double* X = d.p.ref;
    const double* Y = d.p.tgt;
    double* P = d.p.post;
    // per-point vectors in shared memory when they fit (N <= ~4100), else in the workspace
    const bool vsm = SM || vec_fits(N);
    double* rhs = SM ? dyn_smem : (vsm ? dyn_smem : d.rhs);                         // (N,3) solution W = C^T
    double* cur = SM ? dyn_smem + 3 * N : (vsm ? dyn_smem + 3 * (size_t)N : d.cur);
    double* colsum = SM ? dyn_smem + 6 * N : (vsm ? dyn_smem + 6 * (size_t)N : d.colsum);
    double* ytp = d.ytp;
    double* S = SM ? dyn_smem + 7 * N : d.sys;
    const int ld = sys_ld(N);
    const double two_b2 = 2.0 * prm.beta * prm.beta;
    const bool prior_f32 = lite && !d.p.corr_is_f64;

    // ---- prior
    if (d.p.prior_given) {
        for (size_t e = tid; e < (size_t)M * N; e += EM_THREADS)
            d.prior[e] = d.p.corr_is_f64 ? static_cast<const double*>(d.p.corr)[e]
                                         : (double)static_cast<const float*>(d.p.corr)[e];
        __syncthreads();
    } else {
        greedy_prior(d, prm.mode, prm.threshold, s);
    }

    // ---- Gram matrices (track.py:44-46, trackerlite.py:319-320), sigma^2 (track.py:53-56, trackerlite.py:321)
    for (size_t e = tid; e < (size_t)N * N; e += EM_THREADS) {
        const int i = (int)(e / N), j = (int)(e % N);
        d.gram[e] = exp(-dist2(X + 3 * j, X + 3 * i) / two_b2);
    }
    if (lite) {
        for (size_t e = tid; e < (size_t)N * L; e += EM_THREADS) {
            const int n = (int)(e / L), l = (int)(e % L);
            d.gram_nl[e] = exp(-dist2(d.p.tracked + 3 * l, X + 3 * n) / two_b2);
        }
        for (int e = tid; e < 3 * L; e += EM_THREADS) d.cur_l[e] = d.p.tracked[e];
    }
    for (int e = tid; e < 3 * N; e += EM_THREADS) cur[e] = X[e];
    double acc = 0.0;
    for (size_t e = tid; e < (size_t)M * N; e += EM_THREADS) {
        const int m = (int)(e / N), n = (int)(e % N);
        acc += dist2(X + 3 * n, Y + 3 * m);
    }
    double sigma2 = block_sum(acc, s);
    sigma2 = lite ? (sigma2 / ((double)M * (double)N)) / 3.0 : sigma2 / (3.0 * (double)N * (double)M);
    double gamma = lite ? 0.05 : 0.1;
    const double PI = 3.141592653589793;
    int iterations = 0;

#ifdef EM_TIMING
    if (tid < 16) em_t_s[tid] = 0;
    __syncthreads();
#endif
    EM_TICK_DECL;
    EM_TICK(0);
    for (int it = 1; it < prm.max_iteration; ++it) {
        iterations = it;
        EM_TICK(9);
        // ---------------- E-step (track.py:81-88 / trackerlite.py:375-382): one warp per target row
        const double neg_inv_two_s2 = -1.0 / (2.0 * sigma2);
        const double norm15 = pow(2.0 * PI * sigma2, 1.5);
        const double outlier = lite ? gamma / prm.vol : gamma * norm15 / ((1.0 - gamma) * prm.vol);
        const double one_m_g = 1.0 - gamma;
        for (int m = w; m < M; m += EM_WARPS) {
            const double y0 = Y[3 * m], y1 = Y[3 * m + 1], y2 = Y[3 * m + 2];
            double rs = 0.0;
            for (int n = lane; n < N; n += 32) {
                const double dx = cur[3 * n] - y0, dy = cur[3 * n + 1] - y1, dz = cur[3 * n + 2] - y2;
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                const double like = exp(d2 * neg_inv_two_s2);
                const double pr = d.prior[(size_t)m * N + n];
                // NumPy dtype rule kept: a float32 prior times the scalar (1 - gamma) is a float32 product
                // (trackerlite.py:377-378 with prior from simple_match on the float32 FFN output)
                const double w1 = prior_f32 ? (double)((float)one_m_g * (float)pr) : one_m_g * pr;
                const double v = lite ? w1 * like / norm15 : pr * like;
                P[(size_t)m * N + n] = v;
                rs += v;
            }
            rs = warp_sum(rs);
            const double rden = 1.0 / (rs + outlier);
            for (int n = lane; n < N; n += 32) P[(size_t)m * N + n] *= rden;
        }
        __syncthreads();
        EM_TICK(1);
        // ---------------- column moments: p_n = sum_m P[m,n], ytp_n = sum_m P[m,n] Y[m]
        // R = 1024/N row groups work in parallel; partials are combined in a fixed order (deterministic).
        {
            const int R = (N >= EM_THREADS) ? 1 : EM_THREADS / N;
            for (int t = tid; t < R * N; t += EM_THREADS) {
                const int g = t / N, n = t % N;
                double c0 = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll 4
                for (int m = g; m < M; m += R) {
                    const double pv = P[(size_t)m * N + n];
                    c0 += pv;
                    a0 = fma(pv, Y[3 * m], a0); a1 = fma(pv, Y[3 * m + 1], a1); a2 = fma(pv, Y[3 * m + 2], a2);
                }
                double* o = d.colpart + 4 * (size_t)t;
                o[0] = c0; o[1] = a0; o[2] = a1; o[3] = a2;
            }
            __syncthreads();
            for (int n = tid; n < N; n += EM_THREADS) {
                double c0 = 0.0, a0 = 0.0, a1 = 0.0, a2 = 0.0;
                for (int g = 0; g < R; ++g) {
                    const double* o = d.colpart + 4 * ((size_t)g * N + n);
                    c0 += o[0]; a0 += o[1]; a1 += o[2]; a2 += o[3];
                }
                colsum[n] = c0;
                ytp[3 * n] = a0; ytp[3 * n + 1] = a1; ytp[3 * n + 2] = a2;
            }
        }
        __syncthreads();
        EM_TICK(2);
        // ---------------- M-step assembly (track.py:91-96 / trackerlite.py:411-415)
        //   S = a^T:  S[i][j] = p_i G[i][j] + lambda sigma^2 [i==j]     (G symmetric)
        //   rhs[i]  = ytp_i - p_i * base_i,  base = X (TRACK) or the current prediction (LITE)
        const double reg = prm.lambda * sigma2;
        for (int i = w; i < N; i += EM_WARPS) {
            const double pi_ = colsum[i];
            const double* g = d.gram + (size_t)i * N;
            double* srow = S + (size_t)i * ld;
            for (int j = lane; j < N; j += 32) {
                double v = g[j] * pi_;
                if (i == j) v += reg;
                srow[j] = v;
            }
        }
        for (int e = tid; e < 3 * N; e += EM_THREADS) {
            const int i = e / 3;
            S[(size_t)i * ld + N + (e - 3 * i)] = ytp[e] - (lite ? cur[e] : X[e]) * colsum[i];
        }
        __syncthreads();
        EM_TICK(3);
        lu_solve<SM>(S, N, ld, rhs, lus);   // rhs now holds W = C^T (N,3)
        EM_TICK(9);
        // ---------------- apply: move = G W   (track.py:100 / trackerlite.py:337-341)
        double move2 = 0.0;
        for (int i = w; i < N; i += EM_WARPS) {
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            const double* g = d.gram + (size_t)i * N;
            for (int j = lane; j < N; j += 32) {
                const double gv = g[j];
                a0 = fma(gv, rhs[3 * j], a0); a1 = fma(gv, rhs[3 * j + 1], a1); a2 = fma(gv, rhs[3 * j + 2], a2);
            }
            a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
            if (lane == 0) {
                if (lite) {
                    move2 += a0 * a0 + a1 * a1 + a2 * a2;
                    if (it > 1) { cur[3 * i] += a0; cur[3 * i + 1] += a1; cur[3 * i + 2] += a2; }
                } else {
                    cur[3 * i] = X[3 * i] + a0; cur[3 * i + 1] = X[3 * i + 1] + a1; cur[3 * i + 2] = X[3 * i + 2] + a2;
                }
            }
        }
        if (lite && it > 1) {
            // tracked points: move_l = (C G_nl)^T, thread per (l, dim) running down the N rows (coalesced in l)
            for (int e = tid; e < 3 * L; e += EM_THREADS) {
                const int l = e / 3, dim = e % 3;
                double a = 0.0;
                for (int n = 0; n < N; ++n) a = fma(d.gram_nl[(size_t)n * L + l], rhs[3 * n + dim], a);
                d.cur_l[e] += a;
            }
        }
        if (lite) move2 = block_sum(move2, s); else __syncthreads();
        EM_TICK(7);
        // ---------------- gamma, sigma^2 (track.py:103-112 / trackerlite.py:342-350)
        double cs = 0.0;
        for (int n = tid; n < N; n += EM_THREADS) cs += colsum[n];
        const double sumP = block_sum(cs, s);
        gamma = 1.0 - sumP / (double)M;
        if (lite && gamma < 1e-4) gamma = 1e-4;
        double q = 0.0;
        for (int m = w; m < M; m += EM_WARPS) {
            const double y0 = Y[3 * m], y1 = Y[3 * m + 1], y2 = Y[3 * m + 2];
            for (int n = lane; n < N; n += 32) {
                const double dx = cur[3 * n] - y0, dy = cur[3 * n + 1] - y1, dz = cur[3 * n + 2] - y2;
                const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                q = fma(P[(size_t)m * N + n], d2, q);
            }
        }
        sigma2 = block_sum(q, s) / (3.0 * sumP);
        if (!lite && sigma2 < 1.0) sigma2 = 1.0;
        EM_TICK(8);
        if (lite && sqrt(move2) < 1e-3) break;
    }

    // ---- outputs
    for (int e = tid; e < 3 * N; e += EM_THREADS) {
        if (d.p.ref_out) d.p.ref_out[e] = cur[e];
        if (d.p.coef) d.p.coef[(size_t)(e % 3) * N + e / 3] = (iterations > 0) ? rhs[e] : 0.0;
    }
    if (lite && d.p.tracked_out)
        for (int e = tid; e < 3 * L; e += EM_THREADS) d.p.tracked_out[e] = d.cur_l[e];
    if (tid == 0 && d.p.iterations) *d.p.iterations = iterations;
#ifdef EM_TIMING
    __syncthreads();
    if (tid < 16 && blockIdx.x == 0) g_em_t[tid] = em_t_s[tid];
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0)
        printf("EMT setup %lld estep %lld moments %lld assemble %lld luPre %lld lu2 %lld lu3wait %lld back %lld apply %lld sigma %lld other %lld tile0 %lld diag %lld\n",
               g_em_t[0], g_em_t[1], g_em_t[2], g_em_t[3], g_em_t[4], g_em_t[5], g_em_t[10], g_em_t[6], g_em_t[7], g_em_t[8], g_em_t[9], g_em_t[11], g_em_t[12]);
#endif
}

__global__ void __launch_bounds__(EM_THREADS, 1)
greedy_kernel(const DevProblem* __restrict__ probs, int mode, double threshold) {
    __shared__ Scratch s;
    greedy_prior(probs[blockIdx.x], mode, threshold, s);
}

// tracker.py:1269-1289: post = pre + (C G)^T,  G[n,l] = exp(-|pre_l - inter_n|^2 / (2 beta^2))
__global__ void __launch_bounds__(128)
predict_one_rep_kernel(const double* __restrict__ pre, int L, const double* __restrict__ inter, int N,
                       double two_b2, const double* __restrict__ coef, double* __restrict__ post) {
    const int l = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (l >= L) return;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
    for (int n = lane; n < N; n += 32) {
        const double g = exp(-dist2(pre + 3 * l, inter + 3 * n) / two_b2);
        a0 = fma(coef[n], g, a0); a1 = fma(coef[N + n], g, a1); a2 = fma(coef[2 * N + n], g, a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane == 0) { post[3 * l] = pre[3 * l] + a0; post[3 * l + 1] = pre[3 * l + 1] + a1; post[3 * l + 2] = pre[3 * l + 2] + a2; }
}

// scipy.stats.trim_mean(axis=0): per output element sort the E samples, drop int(p*E) at both ends.
__global__ void __launch_bounds__(128)
trim_mean_kernel(const double* __restrict__ stack, int E, int count, int cut, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    // rank-based selection (E is small, <= a few dozen): an element is kept when its rank is in [cut, E-cut)
    double sum = 0.0;
    for (int a = 0; a < E; ++a) {
        const double va = stack[(size_t)a * count + i];
        int rank = 0;
        for (int b = 0; b < E; ++b) {
            const double vb = stack[(size_t)b * count + i];
            rank += (vb < va) || (vb == va && b < a);
        }
        if (rank >= cut && rank < E - cut) sum += va;
    }
    out[i] = sum / (double)(E - 2 * cut);
}

struct Layout {
    size_t gram, gram_nl, sys, prior, colsum, ytp, colpart, rhs, cur, cur_l, rowmax, rowarg, match, col_dead, pairs, n_pairs, total;
};

static Layout layout_for(int N, int M, int L) {
    Layout o{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t at = off; off = align_up(off + bytes, 256); return at; };
    o.gram = take((size_t)N * N * 8);
    o.gram_nl = take((size_t)N * (L > 0 ? L : 1) * 8);
    o.sys = take((size_t)N * sys_ld(N) * 8);
    o.prior = take((size_t)M * N * 8);
    o.colsum = take((size_t)N * 8);
    o.ytp = take((size_t)N * 24);
    o.colpart = take((size_t)(N > EM_THREADS ? N : EM_THREADS) * 32);
    o.rhs = take((size_t)N * 24);
    o.cur = take((size_t)N * 24);
    o.cur_l = take((size_t)(L > 0 ? L : 1) * 24);
    o.rowmax = take((size_t)M * 8);
    o.rowarg = take((size_t)M * 4);
    o.match = take((size_t)M * 4);
    o.col_dead = take((size_t)N);
    o.pairs = take((size_t)(M < N ? M : N) * 8 + 8);
    o.n_pairs = take(8);
    o.total = off;
    return o;
}

static void bind(DevProblem& d, char* base, const Layout& o) {
    d.gram = (double*)(base + o.gram); d.gram_nl = (double*)(base + o.gram_nl); d.sys = (double*)(base + o.sys);
    d.prior = (double*)(base + o.prior); d.colsum = (double*)(base + o.colsum); d.ytp = (double*)(base + o.ytp); d.colpart = (double*)(base + o.colpart);
    d.rhs = (double*)(base + o.rhs); d.cur = (double*)(base + o.cur);
    d.cur_l = (double*)(base + o.cur_l); d.rowmax = (double*)(base + o.rowmax); d.rowarg = (int*)(base + o.rowarg);
    d.match = (int*)(base + o.match); d.col_dead = (unsigned char*)(base + o.col_dead);
    d.pairs = (int*)(base + o.pairs); d.n_pairs = (int*)(base + o.n_pairs);
}

}  // namespace ct

namespace ct {
// prgls_grid.cu: the same EM spread over the whole GPU, for point sets too large for one CTA
size_t grid_em_workspace_bytes(int N, int M, int L);
int grid_em_run(const CtPrglsParams& prm, const CtPrglsProblem& q, void* ws, size_t ws_bytes, double** prior_slot, bool run,
                cudaStream_t s, unsigned long long* launches);
constexpr int EM_GRID_MIN_N = 512;               // problems with at least this many reference points take the grid path
                                                 // (one CTA: 2.85 ms per iteration at N = 512; grid: see profiles/r2_em_timings.md)
}  // namespace ct

using namespace ct;

extern "C" size_t ct_prgls_workspace_bytes(int n_ref, int n_tgt, int n_tracked) {
    size_t need = layout_for(n_ref, n_tgt, n_tracked).total + align_up(sizeof(DevProblem), 256) + 256;
    if (n_ref >= EM_GRID_MIN_N) need += grid_em_workspace_bytes(n_ref, n_tgt, n_tracked) + 256;
    return need;
}

extern "C" size_t ct_greedy_workspace_bytes(int n_ref, int n_tgt) { return ct_prgls_workspace_bytes(n_ref, n_tgt, 0); }

extern "C" int ct_prgls(const CtPrglsParams* prm, const CtPrglsProblem* problems, int batch,
                        void* ws, size_t ws_bytes, void* stream) {
    CT_REQUIRE(prm && problems && ws, "ct_prgls: null argument");
    CT_REQUIRE(prm->mode == CT_PRGLS_TRACK || prm->mode == CT_PRGLS_LITE, "ct_prgls: unknown mode %d", prm->mode);
    // the unpivoted elimination needs the regularised system (lambda sigma^2 > 0 keeps every pivot positive); with
    // lambda = 0 and duplicate points G is singular and LAPACK gesv (track.py:97) would raise LinAlgError
    CT_REQUIRE(prm->lambda > 0.0, "ct_prgls: lambda must be > 0 (singular M-step system), got %g", prm->lambda);
    if (batch == 0) return 0;
    CT_REQUIRE(((uintptr_t)ws & 255) == 0, "ct_prgls: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    // large problems: one at a time on the whole GPU (prgls_grid.cu); the rest as one batched launch, one CTA each
    {
        std::vector<CtPrglsProblem> small;
        size_t used = 0;
        bool any_large = false;
        for (int b = 0; b < batch; ++b) any_large = any_large || problems[b].n_ref >= EM_GRID_MIN_N;
        if (any_large) {
            for (int b = 0; b < batch; ++b) {
                const CtPrglsProblem& p = problems[b];
                char* wsb = static_cast<char*>(ws) + used;
                const size_t avail = ws_bytes - used;
                if (p.n_ref < EM_GRID_MIN_N) {
                    const size_t need = ct_prgls_workspace_bytes(p.n_ref, p.n_tgt, p.n_tracked);
                    CT_REQUIRE(need <= avail, "ct_prgls: workspace too small");
                    if (ct_prgls(prm, &p, 1, wsb, need, stream)) return 1;
                    used += align_up(need, 256);
                    continue;
                }
                CT_REQUIRE(p.n_ref >= 2 && p.n_tgt >= 1 && p.ref && p.tgt && p.corr && p.post, "ct_prgls: problem %d is malformed", b);
                const int L = prm->mode == CT_PRGLS_LITE ? p.n_tracked : 0;
                const size_t greedy_bytes = layout_for(p.n_ref, p.n_tgt, 0).total + align_up(sizeof(DevProblem), 256) + 256;
                const size_t grid_bytes = grid_em_workspace_bytes(p.n_ref, p.n_tgt, L);
                CT_REQUIRE(greedy_bytes + grid_bytes <= avail, "ct_prgls: workspace too small (%zu < %zu)", avail, greedy_bytes + grid_bytes);
                double* prior = nullptr;
                unsigned long long launched = 0;
                if (grid_em_run(*prm, p, wsb + greedy_bytes, grid_bytes, &prior, false, s, nullptr)) return 1;
                if (!p.prior_given) {
                    DevProblem d{};
                    d.p = p;
                    Layout o = layout_for(p.n_ref, p.n_tgt, 0);
                    bind(d, wsb + align_up(sizeof(DevProblem), 256), o);
                    d.prior = prior; d.pairs = nullptr; d.n_pairs = nullptr;
                    if (stage_h2d(wsb, &d, sizeof(d), s)) return 1;
                    greedy_kernel<<<1, EM_THREADS, 0, s>>>(reinterpret_cast<const DevProblem*>(wsb), prm->mode, prm->threshold);
                    CT_LAUNCHED("greedy_kernel");
                }
                ProfScope prof(PROF_EM, s);
                if (grid_em_run(*prm, p, wsb + greedy_bytes, grid_bytes, nullptr, true, s, &launched)) return 1;
                g_launches.fetch_add(launched, std::memory_order_relaxed);
                used += align_up(greedy_bytes + grid_bytes, 256);
            }
            return 0;
        }
    }
    std::vector<DevProblem> host(batch);
    char* base = static_cast<char*>(ws);
    size_t off = align_up((size_t)batch * sizeof(DevProblem), 256);
    int max_n = 0;
    for (int b = 0; b < batch; ++b) {
        const CtPrglsProblem& p = problems[b];
        CT_REQUIRE(p.n_ref >= 2 && p.n_tgt >= 1 && p.n_tracked >= 0, "ct_prgls: problem %d has bad sizes", b);
        CT_REQUIRE(p.ref && p.tgt && p.corr && p.post, "ct_prgls: problem %d has a null required pointer", b);
        CT_REQUIRE(prm->mode != CT_PRGLS_LITE || p.n_tracked == 0 || (p.tracked && p.tracked_out),
                   "ct_prgls: problem %d needs tracked / tracked_out", b);
        Layout o = layout_for(p.n_ref, p.n_tgt, p.n_tracked);
        host[b].p = p;
        bind(host[b], base + off, o);
        if (!(prm->mode == CT_PRGLS_LITE)) host[b].p.n_tracked = 0;
        off += o.total;
        if (p.n_ref > max_n) max_n = p.n_ref;
    }
    CT_REQUIRE(off <= ws_bytes, "ct_prgls: workspace too small (%zu < %zu)", ws_bytes, off);
    if (stage_h2d(ws, host.data(), (size_t)batch * sizeof(DevProblem), s)) return 1;
    // every CTA carves its own plan out of the same allocation: size it for the most demanding problem
    size_t smem = 0;
    for (int b = 0; b < batch; ++b) {
        const int n = problems[b].n_ref;
        const size_t need = sys_fits(n) ? (size_t)56 * n + (size_t)8 * n * sys_ld(n) : (vec_fits(n) ? (size_t)56 * n : 0);
        if (need > smem) smem = need;
    }
    bool all_sm = true;
    for (int b = 0; b < batch; ++b) all_sm = all_sm && sys_fits(problems[b].n_ref);
    // per device / context attribute: set on every call (cheap) so a process that drives several GPUs is correct
    CT_CUDA(cudaFuncSetAttribute(prgls_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EM_SMEM_BUDGET));
    CT_CUDA(cudaFuncSetAttribute(prgls_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EM_SMEM_BUDGET));
    ProfScope prof(PROF_EM, s);
    if (all_sm) {
        prgls_kernel<true><<<batch, EM_THREADS, smem, s>>>(static_cast<const DevProblem*>(ws), *prm);
    } else {
        // mixed / large problems: vectors in shared memory when they fit, system in the workspace
        size_t vsmem = 0;
        for (int b = 0; b < batch; ++b) {
            const size_t need = vec_fits(problems[b].n_ref) ? (size_t)56 * problems[b].n_ref : 0;
            if (need > vsmem) vsmem = need;
        }
        prgls_kernel<false><<<batch, EM_THREADS, vsmem, s>>>(static_cast<const DevProblem*>(ws), *prm);
    }
    CT_LAUNCHED("prgls_kernel");
    return 0;
}

extern "C" int ct_greedy_prior(const void* corr, int corr_is_f64, int n_tgt, int n_ref, int mode, double threshold,
                               double* prior, int* pairs, int* n_pairs, void* ws, size_t ws_bytes, void* stream) {
    CT_REQUIRE(corr && prior && ws, "ct_greedy_prior: null argument");
    CT_REQUIRE(n_ref >= 2 && n_tgt >= 1, "ct_greedy_prior: bad sizes");
    CT_REQUIRE(ws_bytes >= ct_greedy_workspace_bytes(n_ref, n_tgt), "ct_greedy_prior: workspace too small");
    CT_REQUIRE(((uintptr_t)ws & 255) == 0, "ct_greedy_prior: workspace must be 256-byte aligned");
    cudaStream_t s = (cudaStream_t)stream;
    DevProblem d{};
    d.p.corr = corr; d.p.corr_is_f64 = corr_is_f64; d.p.n_ref = n_ref; d.p.n_tgt = n_tgt;
    Layout o = layout_for(n_ref, n_tgt, 0);
    char* base = static_cast<char*>(ws) + align_up(sizeof(DevProblem), 256);
    bind(d, base, o);
    d.prior = prior;
    d.pairs = pairs;
    d.n_pairs = n_pairs;
    if (stage_h2d(ws, &d, sizeof(d), s)) return 1;
    greedy_kernel<<<1, EM_THREADS, 0, s>>>(static_cast<const DevProblem*>(ws), mode, threshold);
    CT_LAUNCHED("greedy_kernel");
    return 0;
}

extern "C" int ct_predict_one_rep(const double* pre, int L, const double* inter, int N, double beta,
                                  const double* coef, double* post, void* stream) {
    CT_REQUIRE(pre && inter && coef && post, "ct_predict_one_rep: null argument");
    if (L == 0) return 0;
    predict_one_rep_kernel<<<cdiv(L, 4), 128, 0, (cudaStream_t)stream>>>(pre, L, inter, N, 2.0 * beta * beta, coef, post);
    CT_LAUNCHED("predict_one_rep_kernel");
    return 0;
}

extern "C" int ct_trim_mean(const double* stack, int e, int count, double proportion, double* out, void* stream) {
    CT_REQUIRE(stack && out && e >= 1, "ct_trim_mean: bad argument");
    const int cut = (int)(proportion * e);
    CT_REQUIRE(e - 2 * cut > 0, "ct_trim_mean: proportion too big");
    if (count == 0) return 0;
    trim_mean_kernel<<<cdiv(count, 128), 128, 0, (cudaStream_t)stream>>>(stack, e, count, cut, out);
    CT_LAUNCHED("trim_mean_kernel");
    return 0;
}

// One volume of the replay chain in one call (tracker.py:1269-1289 applied n_rep times, then the single-mode trimmed mean
// of tracker.py:1503-1507 over a stack of one): the time-lapse driver replays hundreds of fitted transforms in volume
// order on one rank, and six separate calls per volume were host-bound.  scratch: 2 x L x 3 doubles.
extern "C" int ct_replay_fit(const double* pre, int L, int n_rep, const double* const* inter, const int* n_ref,
                             const double* beta, const double* const* coef, double proportion, double* scratch,
                             double* out, void* stream) {
    CT_REQUIRE(pre && inter && n_ref && beta && coef && scratch && out && n_rep >= 1, "ct_replay_fit: bad argument");
    if (L == 0) return 0;
    const double* cur = pre;
    for (int i = 0; i < n_rep; ++i) {
        double* nxt = scratch + (size_t)(i & 1) * L * 3;
        if (ct_predict_one_rep(cur, L, inter[i], n_ref[i], beta[i], coef[i], nxt, stream)) return 1;
        cur = nxt;
    }
    return ct_trim_mean(cur, 1, L * 3, proportion, out, stream);
}
