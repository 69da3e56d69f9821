// Accurate correction + tracked label image on the GPU: CUDA instantiation of correction_core.cuh.
// Reference: tracker.py:1177-1191, 1310-1413 and watershed.py:111-151 (see correction_core.cuh).
#include "common.cuh"
#include "correction_core.cuh"

namespace ct {

constexpr int FLOOD_CAP = 2048;                  // watershed.cu
__global__ void ws_flood_warp(ws::Flood f, const int* roots, const int* n_roots, const int* csize);

template <class F>
__global__ void __launch_bounds__(256) corr_pass_kernel(F f, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}
template <class F>
__global__ void __launch_bounds__(64) corr_sparse_kernel(F f, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f(i);
}
// one warp per cell: lanes sum strided voxels, a fixed-shape shuffle tree adds the 32 partial sums
template <class F>
__global__ void __launch_bounds__(128) corr_cells_kernel(F f, int n_cells) {
    const int cell = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (cell >= n_cells) return;
    double s[4];
    f.partial(cell, lane, 32, s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
        for (int k = 0; k < 4; ++k) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
    if (lane == 0) f.finish(cell, s);
}

struct CorrPolicy {
    cudaStream_t s;
    int err = 0;
    unsigned long long launches = 0;
    template <class F> void run(const F& f, long long n) {
        if (n <= 0 || err) return;
        corr_pass_kernel<F><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(f, n);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    template <class F> void run_sparse(const F& f, long long n) {
        if (n <= 0 || err) return;
        corr_sparse_kernel<F><<<(unsigned)((n + 63) / 64), 64, 0, s>>>(f, n);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    template <class F> void run_cells(const F& f, int n_cells) {
        if (n_cells <= 0 || err) return;
        corr_cells_kernel<F><<<(n_cells + 3) / 4, 128, 0, s>>>(f, n_cells);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    void run_flood(const ws::Flood& f, const int* roots, const int* n_roots, const int* csize, long long) {
        if (err) return;
        ws_flood_warp<<<148 * 6, 32, FLOOD_CAP * (int)sizeof(ws::HeapE), s>>>(f, roots, n_roots, csize);
        ++launches;
        if (cudaGetLastError() != cudaSuccess) err = 1;
    }
    void zero(void* p, size_t bytes) { if (cudaMemsetAsync(p, 0, bytes, s) != cudaSuccess) err = 1; }
    void copy_i32(int* dst, const int* src, long long n) {
        if (cudaMemcpyAsync(dst, src, (size_t)n * 4, cudaMemcpyDeviceToDevice, s) != cudaSuccess) err = 1;
    }
};

static size_t corr_small_bytes(long long n, int n_cells, int n_vox) {
    return ws::a256((size_t)n * 4) * 2 + ws::a256((size_t)n_vox * 4) + ws::a256((size_t)n_cells * 12) * 2 + ws::a256(sizeof(corr::State)) + 512;
}
struct CorrBuffers { int *cover, *label, *cell_of, *i_disp, *i_next; corr::State* st; char* rest; };
static CorrBuffers corr_carve(void* wsp, long long n, int n_cells, int n_vox) {
    char* p = reinterpret_cast<char*>(((uintptr_t)wsp + 255) / 256 * 256);
    auto take = [&](size_t bytes) { char* r = p; p += ws::a256(bytes); return r; };
    CorrBuffers b;
    b.cover = (int*)take((size_t)n * 4); b.label = (int*)take((size_t)n * 4); b.cell_of = (int*)take((size_t)n_vox * 4);
    b.i_disp = (int*)take((size_t)n_cells * 12); b.i_next = (int*)take((size_t)n_cells * 12);
    b.st = (corr::State*)take(sizeof(corr::State));
    b.rest = p;
    return b;
}

}  // namespace ct

static int fill_cells(corr::Cells& c, const int16_t* vox4, const int32_t* start, const int32_t* region_min,
                      const int32_t* region_width, int n_cells, int n_vox, const int32_t pad[3], int xi, int yi, int zi,
                      int z_scaling) {
    CT_REQUIRE(vox4 && start && region_min && region_width && pad, "correction: null cell description");
    CT_REQUIRE(n_cells >= 1 && n_vox >= 1 && z_scaling >= 1 && xi >= 1 && yi >= 1 && zi >= 1, "correction: bad sizes");
    c.vox = vox4; c.start = start; c.region_min = region_min; c.region_width = region_width;
    c.n_cells = n_cells; c.n_vox = n_vox;
    for (int a = 0; a < 3; ++a) c.pad[a] = pad[a];
    c.xi = xi; c.yi = yi; c.zi = zi; c.z_scaling = z_scaling;
    return 0;
}

extern "C" size_t ct_correction_workspace_bytes(int x, int y, int z, int n_cells, int n_vox) {
    return ct::corr_small_bytes((long long)x * y * z, n_cells, n_vox);
}

extern "C" int ct_accurate_correction(const int16_t* vox4, const int32_t* start, const int32_t* region_min,
                                      const int32_t* region_width, int n_cells, int n_vox, const int32_t* pad_host,
                                      int xi, int yi, int zi, int z_scaling, const float* prob, const void* raw,
                                      int raw_dtype, int x, int y, int z, double z_xy_ratio, const double* r_tracked_t0,
                                      const double* r_disp_prev, const double* r_tracked_prev, const double* r_pred,
                                      const int32_t* on_boundary, int max_rep, double* r_disp_out, int32_t* i_disp_out,
                                      int32_t* reps_out, void* wsp, size_t ws_bytes, void* stream) {
    corr::Cells c;
    if (fill_cells(c, vox4, start, region_min, region_width, n_cells, n_vox, pad_host, xi, yi, zi, z_scaling)) return 1;
    CT_REQUIRE(prob && raw && r_tracked_t0 && r_disp_prev && r_tracked_prev && r_pred && on_boundary && r_disp_out && i_disp_out && wsp,
               "ct_accurate_correction: null argument");
    CT_REQUIRE(raw_dtype >= 0 && raw_dtype <= 2 && max_rep >= 1 && max_rep <= 64 && z_xy_ratio > 0, "ct_accurate_correction: bad argument");
    CT_REQUIRE(x == xi && y == yi && z * z_scaling == zi, "ct_accurate_correction: interpolated grid %dx%dx%d does not match %dx%dx%d x%d",
               xi, yi, zi, x, y, z, z_scaling);
    CT_REQUIRE(ws_bytes >= ct_correction_workspace_bytes(x, y, z, n_cells, n_vox), "ct_accurate_correction: workspace too small");
    const ws::Dims d{x, y, z};
    ct::CorrBuffers b = ct::corr_carve(wsp, d.n(), n_cells, n_vox);
    ct::CorrPolicy pol;
    pol.s = (cudaStream_t)stream;
    corr::accurate_correction(pol, c, d, prob, raw, raw_dtype, z_xy_ratio, r_tracked_t0, r_disp_prev, r_tracked_prev, r_pred,
                              on_boundary, max_rep, r_disp_out, b.i_disp, b.i_next, b.cell_of, b.cover, b.st);
    CT_REQUIRE(!pol.err, "ct_accurate_correction: a launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    CT_CUDA(cudaMemcpyAsync(i_disp_out, b.i_disp, (size_t)n_cells * 12, cudaMemcpyDeviceToDevice, pol.s));
    if (reps_out) {
        CT_CUDA(cudaMemcpyAsync(reps_out, &b.st->reps, 4, cudaMemcpyDeviceToDevice, pol.s));
        CT_CUDA(cudaMemcpyAsync(reps_out + 1, &b.st->done, 4, cudaMemcpyDeviceToDevice, pol.s));
    }
    ct::g_launches.fetch_add(pol.launches, std::memory_order_relaxed);
    return 0;
}

extern "C" size_t ct_tracked_labels_workspace_bytes(int x, int y, int z, int n_cells, int n_vox) {
    const long long n = (long long)x * y * z;
    return ct::corr_small_bytes(n, n_cells, n_vox) + ws::workspace_bytes(n, z, 1);
}

extern "C" int ct_tracked_labels(const int16_t* vox4, const int32_t* start, const int32_t* region_min,
                                 const int32_t* region_width, int n_cells, int n_vox, const int32_t* pad_host, int xi, int yi,
                                 int zi, int z_scaling, const int32_t* i_disp, const int32_t* on_boundary, int x, int y, int z,
                                 int32_t* labels_out, void* wsp, size_t ws_bytes, void* stream) {
    corr::Cells c;
    if (fill_cells(c, vox4, start, region_min, region_width, n_cells, n_vox, pad_host, xi, yi, zi, z_scaling)) return 1;
    CT_REQUIRE(i_disp && on_boundary && labels_out && wsp, "ct_tracked_labels: null argument");
    CT_REQUIRE(x == xi && y == yi && z * z_scaling == zi, "ct_tracked_labels: grids do not match");
    CT_REQUIRE(ws_bytes >= ct_tracked_labels_workspace_bytes(x, y, z, n_cells, n_vox), "ct_tracked_labels: workspace too small");
    const ws::Dims d{x, y, z};
    ct::CorrBuffers b = ct::corr_carve(wsp, d.n(), n_cells, n_vox);
    ws::Buffers wb;
    ws::carve(wb, b.rest, d.n(), z, 1);
    ct::CorrPolicy pol;
    pol.s = (cudaStream_t)stream;
    pol.run(corr::CellOf{c, b.cell_of}, n_cells);
    corr::motion_to_image(pol, c, d, i_disp, on_boundary, b.cell_of, b.cover, b.label, labels_out, wb);
    CT_REQUIRE(!pol.err, "ct_tracked_labels: a launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ct::g_launches.fetch_add(pol.launches, std::memory_order_relaxed);
    return 0;
}

extern "C" size_t ct_recalculate_cell_boundaries_workspace_bytes(int x, int y, int z) {
    return ws::workspace_bytes((long long)x * y * z, z, 1) + 256;
}

extern "C" int ct_recalculate_cell_boundaries(int32_t* segmentation, const int32_t* overlaps, int x, int y, int z,
                                              int32_t* labels_out, void* wsp, size_t ws_bytes, void* stream) {
    CT_REQUIRE(segmentation && overlaps && labels_out && wsp, "ct_recalculate_cell_boundaries: null argument");
    CT_REQUIRE(x >= 1 && y >= 1 && z >= 1 && (long long)x * y * z < 0x7fffffffLL, "ct_recalculate_cell_boundaries: bad shape");
    CT_REQUIRE(ws_bytes >= ct_recalculate_cell_boundaries_workspace_bytes(x, y, z), "ct_recalculate_cell_boundaries: workspace too small");
    const ws::Dims d{x, y, z};
    ws::Buffers wb;
    ws::carve(wb, wsp, d.n(), z, 1);
    ct::CorrPolicy pol;
    pol.s = (cudaStream_t)stream;
    corr::recalculate_cell_boundaries(pol, d, segmentation, overlaps, nullptr, labels_out, wb);
    CT_REQUIRE(!pol.err, "ct_recalculate_cell_boundaries: a launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    ct::g_launches.fetch_add(pol.launches, std::memory_order_relaxed);
    return 0;
}
