// TEST INFRASTRUCTURE ONLY: sequential instantiation of the watershed pass pipeline (csrc/watershed_core.cuh) so the
// pass logic can be checked against oracle/watershed.py on a machine without a GPU (tests/test_watershed_emul.py
// builds this file with g++ -ffp-contract=off).  The product library contains only the CUDA instantiation
// (csrc/watershed.cu); nothing in the package loads this harness.
#include <cstdlib>
#include <cstring>
#include "../../3deecelltracker_b200/csrc/watershed_core.cuh"
#include "../../3deecelltracker_b200/csrc/correction_core.cuh"

struct HostPolicy {
    template <class F> void run(const F& f, long long n) { for (long long i = 0; i < n; ++i) f(i); }
    template <class F> void run_sparse(const F& f, long long n) { run(f, n); }
    template <class F> void run_flood(const F& f, const int*, const int*, const int*, long long n) { run(f, n); }
    template <class F> void run_cells(const F& f, int n_cells) {
        for (int c = 0; c < n_cells; ++c) { double s[4]; f.partial(c, 0, 1, s); f.finish(c, s); }
    }
    void copy_i32(int* dst, const int* src, long long n) { std::memcpy(dst, src, (size_t)n * 4); }
    void zero(void* p, size_t bytes) { std::memset(p, 0, bytes); }
    void fill_u64(unsigned long long* p, unsigned long long v, int n) { for (int i = 0; i < n; ++i) p[i] = v; }
    void exclusive_scan(const int* flag, int* rank, long long n, int* total) {
        int acc = 0;
        for (long long i = 0; i < n; ++i) { rank[i] = acc; acc += flag[i]; }
        *total = acc;
    }
};

extern "C" int ws_emul_segment(const float* prob, int x, int y, int z, double z_xy_ratio, int method, int min_size,
                               int cell_num, const double* w_xy9, const double* w_z2, int* labels, double* centres,
                               int max_cells, int* scalars_out) {   // centres: (2, max_cells, 3)
    const long long n = (long long)x * y * z;
    void* wsp = std::malloc(ws::workspace_bytes(n, z, max_cells));
    if (!wsp) return 1;
    ws::Buffers b;
    ws::carve(b, wsp, n, z, max_cells);
    ws::Params prm;
    prm.X = x; prm.Y = y; prm.Z = z; prm.z_xy_ratio = z_xy_ratio; prm.method = method; prm.min_size = min_size;
    prm.cell_num = cell_num; prm.max_cells = max_cells;
    for (int j = 0; j < 9; ++j) prm.w_xy[j] = w_xy9[j];
    for (int j = 0; j < 2; ++j) prm.w_z[j] = w_z2[j];
    HostPolicy pol;
    ws::segment(pol, prm, prob, labels, centres, b);
    scalars_out[0] = b.sc->n_cells; scalars_out[1] = b.sc->min_size; scalars_out[2] = b.sc->cell_num;
    scalars_out[3] = b.sc->bg_count;
    std::free(wsp);
    return 0;
}

// Stage dump for debugging: the 2-D stage only (labels of the per-slice flood, bn_output, smoothed distance, peaks).
extern "C" int ws_emul_stage2d(const float* prob, int x, int y, int z, const double* w_xy9, int* lab2d, unsigned char* mask2,
                               double* smooth, unsigned char* peak) {
    const long long n = (long long)x * y * z;
    void* wsp = std::malloc(ws::workspace_bytes(n, z, 16));
    if (!wsp) return 1;
    ws::Buffers b;
    ws::carve(b, wsp, n, z, 16);
    const ws::Dims d{x, y, z};
    HostPolicy pol;
    ws::Gauss1D<0> gx{d, nullptr, nullptr, 8, {}, ws::ActGrid{nullptr, 0}};
    ws::Gauss1D<1> gy{d, nullptr, nullptr, 8, {}, ws::ActGrid{nullptr, 0}};
    for (int j = 0; j < 9; ++j) { gx.w[j] = w_xy9[j]; gy.w[j] = w_xy9[j]; }
    pol.run(ws::Threshold{prob, b.mask}, n);
    pol.run(ws::ColDist{d, b.mask, b.g}, n);
    pol.run(ws::RowDist{d, b.mask, b.g, b.d2}, n);
    pol.run(ws::SqrtPlane{b.d2, b.fa}, n);
    gx.in = b.fa; gx.out = b.fb; pol.run(gx, n);
    gy.in = b.fb; gy.out = b.fa; pol.run(gy, n);
    pol.fill_u64(b.minslot, 0x7ff0000000000000ull, d.Z + 1);
    pol.run(ws::MinReduce{d, b.fa, b.minslot, 1}, n);
    pol.run(ws::Max1D<0>{d, b.fa, b.fb, 7, ws::ActGrid{nullptr, 0}}, n);
    pol.run(ws::Max1D<1>{d, b.fb, b.fc, 7, ws::ActGrid{nullptr, 0}}, n);
    pol.run(ws::Peaks{d, b.fa, b.fc, b.minslot, 1, 7, b.peak, ws::ActGrid{nullptr, 0}}, n);
    ws::flood_stage(pol, d, b, b.mask, b.fa, 1);
    pol.run(ws::Boundary2D{d, b.mask, b.lab, b.mask2}, n);
    std::memcpy(lab2d, b.lab, n * 4); std::memcpy(mask2, b.mask2, n); std::memcpy(smooth, b.fa, n * 8);
    std::memcpy(peak, b.peak, n);
    std::free(wsp);
    return 0;
}


// ---- accurate correction / tracked label image (correction_core.cuh), sequential instantiation
static corr::Cells make_cells(const short* vox4, const int* start, const int* rmin, const int* rwidth, int n_cells, int n_vox,
                              const int* pad, int xi, int yi, int zi, int zs) {
    corr::Cells c;
    c.vox = vox4; c.start = start; c.region_min = rmin; c.region_width = rwidth; c.n_cells = n_cells; c.n_vox = n_vox;
    for (int a = 0; a < 3; ++a) c.pad[a] = pad[a];
    c.xi = xi; c.yi = yi; c.zi = zi; c.z_scaling = zs;
    return c;
}

extern "C" int corr_emul_accurate_correction(const short* vox4, const int* start, const int* rmin, const int* rwidth, int n_cells,
                                             int n_vox, const int* pad, int xi, int yi, int zi, int zs, const float* prob,
                                             const void* raw, int raw_dtype, int x, int y, int z, double ratio,
                                             const double* t0, const double* disp_prev, const double* tracked_prev,
                                             const double* r_pred, const int* on_boundary, int max_rep, double* r_disp,
                                             int* i_disp, int* reps_out) {
    const corr::Cells c = make_cells(vox4, start, rmin, rwidth, n_cells, n_vox, pad, xi, yi, zi, zs);
    const ws::Dims d{x, y, z};
    int* cover = (int*)std::malloc((size_t)d.n() * 4);
    int* cell_of = (int*)std::malloc((size_t)n_vox * 4);
    int* i_next = (int*)std::malloc((size_t)n_cells * 12);
    corr::State st;
    HostPolicy pol;
    corr::accurate_correction(pol, c, d, prob, raw, raw_dtype, ratio, t0, disp_prev, tracked_prev, r_pred, on_boundary, max_rep,
                              r_disp, i_disp, i_next, cell_of, cover, &st);
    reps_out[0] = st.reps; reps_out[1] = st.done;
    std::free(cover); std::free(cell_of); std::free(i_next);
    return 0;
}

extern "C" int corr_emul_tracked_labels(const short* vox4, const int* start, const int* rmin, const int* rwidth, int n_cells,
                                        int n_vox, const int* pad, int xi, int yi, int zi, int zs, const int* i_disp,
                                        const int* on_boundary, int x, int y, int z, int* labels_out) {
    const corr::Cells c = make_cells(vox4, start, rmin, rwidth, n_cells, n_vox, pad, xi, yi, zi, zs);
    const ws::Dims d{x, y, z};
    const long long n = d.n();
    void* wsp = std::malloc(ws::workspace_bytes(n, z, 1));
    ws::Buffers b;
    ws::carve(b, wsp, n, z, 1);
    int* cover = (int*)std::malloc((size_t)n * 4);
    int* label = (int*)std::malloc((size_t)n * 4);
    int* cell_of = (int*)std::malloc((size_t)n_vox * 4);
    HostPolicy pol;
    pol.run(corr::CellOf{c, cell_of}, n_cells);
    corr::motion_to_image(pol, c, d, i_disp, on_boundary, cell_of, cover, label, labels_out, b);
    std::free(cover); std::free(label); std::free(cell_of); std::free(wsp);
    return 0;
}
