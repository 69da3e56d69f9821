// Watershed + centroid stage on the GPU: CUDA instantiation of the pass pipeline in watershed_core.cuh.
//
// Reference: Tracker._watershed (tracker.py:671-684) = watershed_2d (watershed.py:16-52) + watershed_3d (:55-108) +
// relabel_sequential, then ndimage.center_of_mass (tracker.py:646-648).  The reference runs it on the host between
// the U-Net and the matcher, one Python loop iteration per z slice; here the probability map never leaves HBM.
//
// Data-parallel passes (threshold, separable exact EDT, SciPy-exact separable Gaussian, separable maximum filter,
// peak test, union-find connected components, boundaries, sizes, relabel, centre-of-mass sums) are one thread per
// voxel; every pass is HBM-bound elementwise / short-stencil work (algorithmic bytes per voxel in DESIGN.md).  The
// priority flood is inherently sequential PER BASIN GROUP, but floods of different connected components of the
// foreground never interact: one thread floods one component with a private binary heap (L1-resident for cell-sized
// components), all components of all slices in parallel.  The order (value, age, index) restricted to a component is
// the order of the global queue restricted to that component, so the labels equal the single-queue result.
#include "common.cuh"
#include "watershed_core.cuh"

namespace ct {

template <class F>
__global__ void __launch_bounds__(256) ws_pass_kernel(F f, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}
// Same pass as a resident grid-stride loop.  A pass over 9 M voxels as 36 000 one-shot blocks is bound by block turnover
// (every block is one dependent load: ~24 us per pass on B200 however little the functor does); with a resident grid
// the element-wise passes take 18 us and the Gaussian / min / relabel passes 20-60 % less.  The passes whose threads
// chase pointers or walk long windows (union-find link, maximum filter, boundary test) measured 20-70 % SLOWER this way
// and keep the one-shot form.
constexpr int WS_PASS_BLOCKS = 148 * 8;
template <class F>
__global__ void __launch_bounds__(256) ws_pass_loop_kernel(F f, long long n) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) f(i);
}
template <class F> struct ws_one_shot { static constexpr bool value = false; };
template <> struct ws_one_shot<ws::UfLink> { static constexpr bool value = true; };
template <> struct ws_one_shot<ws::Boundary2D> { static constexpr bool value = true; };
template <int A> struct ws_one_shot<ws::Max1D<A>> { static constexpr bool value = true; };

// Roots are sparse and their floods long: a warp that holds one root keeps its 31 other lanes idle anyway, so the flood
// pass runs with small blocks to spread the long-running threads over all SMs' schedulers.
template <class F>
__global__ void __launch_bounds__(64) ws_sparse_kernel(F f, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}

// ---- priority flood, one WARP per connected component of the foreground (roots list written by HeapAlloc).
// The pop order is sequential, but everything around it is not: lane 0 keeps the binary heap in SHARED memory
// (components of up to FLOOD_CAP voxels; larger ones use their slice of the global heap), and after every pop lanes 0-5
// fetch the six neighbours (mask, label, value) at once, so a pop costs one round trip to L2 instead of eighteen.
// Labels are read with ld.global.cg: the label a lane wrote for an earlier pop must be seen by the other lanes.
// Same order (value, age, index) and same neighbour order as ws::Flood, so the labels are identical to it.
constexpr int FLOOD_CAP = 2048;

__global__ void __launch_bounds__(32) ws_flood_warp(ws::Flood f, const int* __restrict__ roots, const int* __restrict__ n_roots,
                                                    const int* __restrict__ csize) {
    extern __shared__ ws::HeapE sheap[];
    const int lane = threadIdx.x;
    const int total = *n_roots;
    const ws::i64 sx = (ws::i64)f.d.Y * f.d.Z, sy = f.d.Z;
    for (int k = blockIdx.x; k < total; k += gridDim.x) {
        const int root = roots[k];
        int n = f.hcnt[root];
        if (n == 0) continue;                                        // uniform over the warp
        ws::HeapE* gh = f.heap + f.hoff[root];
        const bool in_smem = csize[root] <= FLOOD_CAP;
        ws::HeapE* h = in_smem ? sheap : gh;
        if (in_smem) {
            for (int e = lane; e < n; e += 32) sheap[e] = gh[e];
            __syncwarp();
        }
        if (lane == 0)
            for (int e = n / 2 - 1; e >= 0; --e) ws::heap_down(h, n, e);
        __syncwarp();
        int age = 0;
        while (true) {
            n = __shfl_sync(0xffffffffu, n, 0);
            if (n == 0) break;
            int top_idx = 0, l = 0;
            if (lane == 0) {
                top_idx = h[0].idx;
                --n;
                if (n > 0) { h[0] = h[n]; ws::heap_down(h, n, 0); }
                l = __ldcg(f.lab + top_idx);
            }
            top_idx = __shfl_sync(0xffffffffu, top_idx, 0);
            l = __shfl_sync(0xffffffffu, l, 0);
            int x, y, z; f.d.split(top_idx, x, y, z);
            // C order of the neighbour offsets: x-1, y-1, z-1, z+1, y+1, x+1 (lane = position in that order)
            ws::i64 j = -1;
            bool ok = false;
            switch (lane) {
                case 0: ok = x > 0; j = top_idx - sx; break;
                case 1: ok = y > 0; j = top_idx - sy; break;
                case 2: ok = !f.planar && z > 0; j = (ws::i64)top_idx - 1; break;
                case 3: ok = !f.planar && z + 1 < f.d.Z; j = (ws::i64)top_idx + 1; break;
                case 4: ok = y + 1 < f.d.Y; j = top_idx + sy; break;
                case 5: ok = x + 1 < f.d.X; j = top_idx + sx; break;
                default: break;
            }
            // the three loads of a neighbour are independent of each other: one round trip, not three
            double v = 0.0;
            bool cand = false;
            if (ok) {
                const uint8_t mk = __ldg(f.mask + j);
                const int lb = __ldcg(f.lab + j);
                const double im = __ldg(f.img + j);
                cand = mk != 0 && lb == 0;
                v = f.sign * im;
            }
            const unsigned bits = __ballot_sync(0xffffffffu, cand);
#pragma unroll
            for (int q = 0; q < 6; ++q) {
                const long long jq = __shfl_sync(0xffffffffu, (long long)j, q);
                const double vq = __shfl_sync(0xffffffffu, v, q);
                if (lane == 0 && ((bits >> q) & 1u)) {
                    ++age;
                    f.lab[jq] = l;
                    ws::HeapE e; e.v = vq; e.age = age; e.idx = (int)jq;
                    h[n] = e;
                    ws::heap_up(h, n);
                    ++n;
                }
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

__global__ void ws_fill_u64(unsigned long long* p, unsigned long long v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ---- exclusive scan of 0/1 flags over the volume (rank of every surviving label root in raster order)
constexpr int SCAN_BLOCK = 1024, SCAN_PER_THREAD = 4, SCAN_TILE = SCAN_BLOCK * SCAN_PER_THREAD;

__device__ __forceinline__ int block_exclusive_scan(int v, int* total) {
    __shared__ int warp_tot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, t, o);
            if (lane >= o) t += u;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    const int base = w > 0 ? warp_tot[w - 1] : 0;
    *total = warp_tot[31];
    __syncthreads();
    return base + inc - v;
}
__global__ void __launch_bounds__(SCAN_BLOCK) ws_scan_tiles(const int* flag, int* tile_sum, long long n) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
    int s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) if (base + k < n) s += flag[base + k];
    int tot;
    block_exclusive_scan(s, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(SCAN_BLOCK) ws_scan_tile_sums(int* tile_sum, int tiles, int* total_out) {
    int carry = 0;
    for (int t0 = 0; t0 < tiles; t0 += SCAN_BLOCK) {
        const int t = t0 + threadIdx.x;
        const int v = t < tiles ? tile_sum[t] : 0;
        int tot;
        const int ex = block_exclusive_scan(v, &tot);
        if (t < tiles) tile_sum[t] = carry + ex;
        carry += tot;
    }
    if (threadIdx.x == 0) *total_out = carry;
}
__global__ void __launch_bounds__(SCAN_BLOCK) ws_scan_apply(const int* flag, const int* tile_sum, int* rank, long long n) {
    const long long base = (long long)blockIdx.x * SCAN_TILE + (long long)threadIdx.x * SCAN_PER_THREAD;
    int f[SCAN_PER_THREAD], s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) { f[k] = base + k < n ? flag[base + k] : 0; s += f[k]; }
    int tot;
    int ex = block_exclusive_scan(s, &tot) + tile_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_PER_THREAD; ++k) {
        if (base + k < n) rank[base + k] = ex;
        ex += f[k];
    }
}

struct CudaPolicy {
    cudaStream_t s;
    int err = 0;
    unsigned long long launches = 0;
    int* tile_sum;                      // scratch for the scan (carved from the workspace by the caller)

    template <class F> void run(const F& f, long long n) {
        if (n <= 0 || err) return;
        const long long blocks = (n + 255) / 256;
        if (ws_one_shot<F>::value || blocks <= WS_PASS_BLOCKS) ws_pass_kernel<F><<<(unsigned)blocks, 256, 0, s>>>(f, n);
        else ws_pass_loop_kernel<F><<<WS_PASS_BLOCKS, 256, 0, s>>>(f, n);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    template <class F> void run_sparse(const F& f, long long n) {
        if (n <= 0 || err) return;
        ws_sparse_kernel<F><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(f, n);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    void run_flood(const ws::Flood& f, const int* roots, const int* n_roots, const int* csize, long long) {
        if (err) return;
        static const int smem = FLOOD_CAP * (int)sizeof(ws::HeapE);
        ws_flood_warp<<<148 * 6, 32, smem, s>>>(f, roots, n_roots, csize);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    void zero(void* p, size_t bytes) { if (cudaMemsetAsync(p, 0, bytes, s) != cudaSuccess) err = 1; }
    void fill_u64(unsigned long long* p, unsigned long long v, int n) {
        ws_fill_u64<<<(n + 255) / 256, 256, 0, s>>>(p, v, n);
        ++launches;
    }
    void exclusive_scan(const int* flag, int* rank, long long n, int* total) {
        const int tiles = (int)((n + SCAN_TILE - 1) / SCAN_TILE);
        ws_scan_tiles<<<tiles, SCAN_BLOCK, 0, s>>>(flag, tile_sum, n);
        ws_scan_tile_sums<<<1, SCAN_BLOCK, 0, s>>>(tile_sum, tiles, total);
        ws_scan_apply<<<tiles, SCAN_BLOCK, 0, s>>>(flag, tile_sum, rank, n);
        launches += 3;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
};

}  // namespace ct

extern "C" size_t ct_watershed_workspace_bytes(int x, int y, int z, int max_cells) {
    const long long n = (long long)x * y * z;
    const size_t tiles = (size_t)((n + ct::SCAN_TILE - 1) / ct::SCAN_TILE);
    return ws::workspace_bytes(n, z, max_cells) + ws::a256(tiles * 4) + 256;
}

extern "C" int ct_watershed_segment(const float* prob, int x, int y, int z, double z_xy_ratio, int method, int min_size,
                                    int cell_num, const double* gauss_w_xy9, const double* gauss_w_z2, int32_t* labels,
                                    double* centres, int max_cells, int32_t* scalars_out, void* wsp, size_t ws_bytes,
                                    void* stream) {
    CT_REQUIRE(prob && labels && centres && scalars_out && wsp && gauss_w_xy9 && gauss_w_z2, "ct_watershed_segment: null argument");
    CT_REQUIRE(x >= 1 && y >= 1 && z >= 1 && (long long)x * y * z < 0x7fffffffLL, "ct_watershed_segment: bad shape %d x %d x %d", x, y, z);
    CT_REQUIRE(x < ws::COL_INF && y < ws::COL_INF, "ct_watershed_segment: plane too large");
    CT_REQUIRE(method == 0 || method == 1, "ct_watershed_segment: method must be 0 (min_size) or 1 (cell_num)");
    CT_REQUIRE(max_cells >= 1, "ct_watershed_segment: max_cells must be >= 1");
    CT_REQUIRE(ws_bytes >= ct_watershed_workspace_bytes(x, y, z, max_cells), "ct_watershed_segment: workspace too small");
    const long long n = (long long)x * y * z;
    ws::Buffers b;
    ws::carve(b, wsp, n, z, max_cells);
    ws::Params prm;
    prm.X = x; prm.Y = y; prm.Z = z; prm.z_xy_ratio = z_xy_ratio; prm.method = method; prm.min_size = min_size;
    prm.cell_num = cell_num; prm.max_cells = max_cells;
    for (int j = 0; j < 9; ++j) prm.w_xy[j] = gauss_w_xy9[j];
    for (int j = 0; j < 2; ++j) prm.w_z[j] = gauss_w_z2[j];
    ct::CudaPolicy pol;
    pol.s = (cudaStream_t)stream;
    pol.tile_sum = reinterpret_cast<int*>(reinterpret_cast<char*>(b.thr) + 256);
    ct::ProfScope prof(ct::PROF_WATERSHED, pol.s);
    ws::segment(pol, prm, prob, labels, centres, b);
    CT_REQUIRE(!pol.err, "ct_watershed_segment: a launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    // scalars_out (device, 4 x int32): n_cells, min_size, cell_num, background voxel count
    CT_CUDA(cudaMemcpyAsync(scalars_out, &b.sc->n_cells, 4, cudaMemcpyDeviceToDevice, pol.s));
    CT_CUDA(cudaMemcpyAsync(scalars_out + 1, &b.sc->min_size, 4, cudaMemcpyDeviceToDevice, pol.s));
    CT_CUDA(cudaMemcpyAsync(scalars_out + 2, &b.sc->cell_num, 4, cudaMemcpyDeviceToDevice, pol.s));
    CT_CUDA(cudaMemcpyAsync(scalars_out + 3, &b.sc->bg_count, 4, cudaMemcpyDeviceToDevice, pol.s));
    ct::g_launches.fetch_add(pol.launches, std::memory_order_relaxed);
    return 0;
}

extern "C" size_t ct_label_components_workspace_bytes(int x, int y, int z) { return ct_watershed_workspace_bytes(x, y, z, 1); }

extern "C" int ct_label_components(const int32_t* image, int x, int y, int z, int32_t* labels, int32_t* n_out, void* wsp,
                                   size_t ws_bytes, void* stream) {
    CT_REQUIRE(image && labels && n_out && wsp, "ct_label_components: null argument");
    CT_REQUIRE(x >= 1 && y >= 1 && z >= 1 && (long long)x * y * z < 0x7fffffffLL, "ct_label_components: bad shape");
    CT_REQUIRE(ws_bytes >= ct_label_components_workspace_bytes(x, y, z), "ct_label_components: workspace too small");
    const long long n = (long long)x * y * z;
    ws::Buffers b;
    ws::carve(b, wsp, n, z, 1);
    ct::CudaPolicy pol;
    pol.s = (cudaStream_t)stream;
    pol.tile_sum = reinterpret_cast<int*>(reinterpret_cast<char*>(b.thr) + 256);
    ws::label_equal_values(pol, ws::Dims{x, y, z}, image, labels, b, &b.sc->n_cells);
    CT_REQUIRE(!pol.err, "ct_label_components: a launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CT_CUDA(cudaMemcpyAsync(n_out, &b.sc->n_cells, 4, cudaMemcpyDeviceToDevice, pol.s));
    ct::g_launches.fetch_add(pol.launches, std::memory_order_relaxed);
    return 0;
}
