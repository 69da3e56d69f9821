// Accurate correction of the tracked positions and the tracked label image (Tracker._accurate_correction,
// tracker.py:1177-1191; _correction_once_interp :1310-1350; _transform_cells_quick :1352-1389;
// _evaluate_correction :1401-1413; _transform_motion_to_image :1391-1399; recalculate_cell_boundaries,
// watershed.py:111-151) as data-parallel passes, launched through the same Policy as watershed_core.cuh (CUDA in the
// product, sequential in the CPU test harness).
//
// The reference stamps every cell's cropped sub-image (region_list) into a padded label volume at its integer
// displacement, counts overlaps, and takes the intensity-weighted centre of mass of every label on the z planes that
// exist in the raw stack.  Here a cell is a list of voxel coordinates (interpolated grid, volume 1); one pass counts how
// many cells cover each voxel of the raw grid (scatter), one pass per cell sums weight and weight x coordinate over its
// voxels that are covered exactly once (segmented reduction, fixed order inside a cell), and a third applies the
// reference's displacement update.  Up to REP_NUM_CORRECTION = 20 repetitions are enqueued back to back; a device flag
// turns the remaining ones into no-ops once the correction has converged, so the host never waits inside the loop.
#pragma once
#include "watershed_core.cuh"

namespace corr {

using ws::i64;

struct Cells {                       // cells of volume 1 on the interpolated grid (cal_subregions, tracker.py:1093-1110)
    const short* vox;                // (n_vox, 4): x, y, z (interpolated), 0
    const int* start;                // (L + 1) offsets into vox
    const int* region_min;           // (L, 3)
    const int* region_width;         // (L, 3)
    int n_cells, n_vox;
    int pad[3];                      // max region width per axis (label_padding)
    int xi, yi, zi;                  // interpolated volume: x_siz, y_siz, z_siz * z_scaling
    int z_scaling;
};

struct State {                       // device-resident loop state
    unsigned long long maxabs;       // bit pattern of max |correction test| of the repetition (>= 0: bit order = value order)
    int done, reps, pad0, pad1;
};

WS_HD double nan_value() { return ws::bits_dbl(0x7ff8000000000000ull); }

// a cell is stamped only when its whole sub-image fits the padded volume (tracker.py:1378-1379)
WS_HD bool cell_fits(const Cells& c, int cell, const int* i_disp) {
    const int dims[3] = {c.xi, c.yi, c.zi};
    for (int a = 0; a < 3; ++a) {
        const int lo = c.region_min[cell * 3 + a] + i_disp[cell * 3 + a] + c.pad[a];
        if (lo < 0 || lo + c.region_width[cell * 3 + a] > dims[a] + 2 * c.pad[a]) return false;
    }
    return true;
}
// voxel v of a cell moved by its displacement -> raw-grid index, or -1 when it is cropped away or lies between planes
WS_HD i64 moved_index(const Cells& c, const ws::Dims& d, const short* v, const int* disp) {
    const int x = v[0] + disp[0], y = v[1] + disp[1], z = v[2] + disp[2];
    if (x < 0 || x >= c.xi || y < 0 || y >= c.yi || z < 0 || z >= c.zi) return -1;
    const int zo = z - c.z_scaling / 2;
    if (zo < 0 || zo % c.z_scaling != 0) return -1;
    const int k = zo / c.z_scaling;
    if (k >= d.Z) return -1;
    return ((i64)x * d.Y + y) * d.Z + k;
}

// ---- r_disp = r_disp_prev + (r_pred - r_tracked_prev); i_disp = rint(r_disp * [1, 1, z_scaling / z_xy_ratio])
struct Begin {
    int n; const double* r_disp_prev; const double* r_pred; const double* r_tracked_prev; double zs_over_ratio;
    double* r_disp; int* i_disp; State* st;
    WS_HD void operator()(i64 i) const {
        if (i == 0) { st->done = 0; st->reps = 0; st->maxabs = 0ull; }
        const double v = ws::add_rn(r_disp_prev[i], ws::add_rn(r_pred[i], -r_tracked_prev[i]));
        r_disp[i] = v;
        i_disp[i] = (int)rint(i % 3 == 2 ? ws::mul_rn(v, zs_over_ratio) : v);
    }
};

// ---- overlap counts on the raw grid: mask of _transform_cells_quick restricted to the planes that exist
struct CountCover {
    Cells c; ws::Dims d; const int* i_disp; const int* cell_of; int* cover; const State* st;
    WS_HD void operator()(i64 v) const {
        if (st->done) return;
        const int cell = cell_of[v];
        if (!cell_fits(c, cell, i_disp)) return;
        const i64 idx = moved_index(c, d, c.vox + v * 4, i_disp + cell * 3);
        if (idx >= 0) ws::atomic_add_i(cover + idx, 1);
    }
};
struct CellOf {                                  // voxel -> its cell (filled once)
    Cells c; int* cell_of;
    WS_HD void operator()(i64 cell) const {
        for (int v = c.start[cell]; v < c.start[cell + 1]; ++v) cell_of[v] = (int)cell;
    }
};

// ---- per cell: centre of mass of (prob + raw / 65536) over its exclusively covered voxels, then the update
// raw_dtype: 0 = uint16, 1 = float32, 2 = uint8
WS_HD double weight_at(const float* prob, const void* raw, int raw_dtype, i64 i) {
    double r;
    if (raw_dtype == 0) r = (double)reinterpret_cast<const unsigned short*>(raw)[i];
    else if (raw_dtype == 2) r = (double)reinterpret_cast<const unsigned char*>(raw)[i];
    else r = (double)reinterpret_cast<const float*>(raw)[i];
    return ws::add_rn((double)prob[i], r / 65536.0);
}
struct CellUpdate {
    Cells c; ws::Dims d; const float* prob; const void* raw; int raw_dtype; const int* cover; const int* on_boundary;
    const double* r_tracked_t0; double ratio, inv_ratio, inv_zs, ratio_over_zs, zs_over_ratio;
    double* r_disp; const int* i_disp; int* i_disp_next; State* st;
    // lane `lane` of `lanes` sums every lanes-th voxel of the cell (fixed order: the result does not depend on timing)
    WS_HD void partial(i64 cell, int lane, int lanes, double (&s)[4]) const {
        s[0] = s[1] = s[2] = s[3] = 0.0;
        if (st->done || on_boundary[cell] || !cell_fits(c, (int)cell, i_disp)) return;
        const int* disp = i_disp + cell * 3;
        for (int v = c.start[cell] + lane; v < c.start[cell + 1]; v += lanes) {
            const i64 idx = moved_index(c, d, c.vox + (i64)v * 4, disp);
            if (idx < 0 || cover[idx] != 1) continue;
            int x, y, z; d.split(idx, x, y, z);
            const double w = weight_at(prob, raw, raw_dtype, idx);
            s[0] += w; s[1] += w * x; s[2] += w * y; s[3] += w * z;
        }
    }
    WS_HD void finish(i64 cell, const double (&s)[4]) const {
        if (st->done) return;
        const int* disp = i_disp + cell * 3;
        const double sw = s[0];
        const bool lost = !(sw > 0.0) && !(sw < 0.0);                      // 0 / 0 -> NaN in the reference: cell lost
        const double cen[3] = {s[1] / sw, s[2] / sw, s[3] / sw};
        double corr3[3];
        for (int a = 0; a < 3; ++a) {
            // l_coordinates_prgls_int_move = r_tracked_t0 * [1, 1, 1/ratio] + i_disp * [1, 1, 1/z_scaling]
            const double t0 = a == 2 ? ws::mul_rn(r_tracked_t0[cell * 3 + a], inv_ratio) : r_tracked_t0[cell * 3 + a];
            const double mv = a == 2 ? ws::mul_rn((double)disp[a], inv_zs) : (double)disp[a];
            double cr = lost ? 0.0 : ws::add_rn(cen[a], -ws::add_rn(t0, mv));
            if (a == 2) cr = ws::mul_rn(cr, ratio);
            corr3[a] = cr;
        }
        double worst = 0.0;
        for (int a = 0; a < 3; ++a) {
            // r_displacement_from_vol1 = i_disp * [1, 1, ratio / z_scaling] + correction;  i_disp_new = rint(r * [1, 1, zs / ratio])
            const double base = a == 2 ? ws::mul_rn((double)disp[a], ratio_over_zs) : (double)disp[a];
            const double r = ws::add_rn(base, corr3[a]);
            r_disp[cell * 3 + a] = r;
            i_disp_next[cell * 3 + a] = (int)rint(a == 2 ? ws::mul_rn(r, zs_over_ratio) : r);
            const double t = fabs(a == 2 ? ws::mul_rn(corr3[a], zs_over_ratio) : corr3[a]);   // _evaluate_correction
            if (t > worst) worst = t;
        }
        ws::atomic_max_u64(&st->maxabs, ws::dbl_bits(worst));
    }
};
struct Decide {                                  // end of one repetition: adopt the new displacements, test convergence
    int* i_disp; const int* i_disp_next; const State* st;
    WS_HD void operator()(i64 i) const {
        if (st->done) return;
        i_disp[i] = i_disp_next[i];
    }
};
struct DecideFlag {
    State* st;
    WS_HD void operator()(i64) const {
        if (st->done) return;
        st->reps += 1;
        if (ws::bits_dbl(st->maxabs) < 0.5) st->done = 1;
        st->maxabs = 0ull;
    }
};

// ---- tracked label image (_transform_motion_to_image): labels stamped at the final displacement, overlaps and
// boundary cells removed, then cell boundaries re-drawn inside the overlaps by a per-slice watershed
struct StampLabel {
    Cells c; ws::Dims d; const int* i_disp; const int* cell_of; int* label;
    WS_HD void operator()(i64 v) const {
        const int cell = cell_of[v];
        if (!cell_fits(c, cell, i_disp)) return;
        const i64 idx = moved_index(c, d, c.vox + v * 4, i_disp + cell * 3);
        if (idx >= 0) label[idx] = cell + 1;                             // unique wherever cover == 1; zeroed elsewhere
    }
};
struct CountCoverAlways {
    Cells c; ws::Dims d; const int* i_disp; const int* cell_of; int* cover;
    WS_HD void operator()(i64 v) const {
        const int cell = cell_of[v];
        if (!cell_fits(c, cell, i_disp)) return;
        const i64 idx = moved_index(c, d, c.vox + v * 4, i_disp + cell * 3);
        if (idx >= 0) ws::atomic_add_i(cover + idx, 1);
    }
};
struct MarkersAndMasks {
    const int* cover; const int* on_boundary; int* label; uint8_t* mask_image; uint8_t* overlap;   // on_boundary may be null
    WS_HD void operator()(i64 i) const {
        int l = label[i];
        const bool ov = cover[i] > 1;
        if (ov || (l > 0 && on_boundary && on_boundary[l - 1])) l = 0;     // tracker.py:1395-1397
        label[i] = l;
        overlap[i] = ov ? 1 : 0;
        mask_image[i] = (l > 0 || ov) ? 1 : 0;                             // watershed.py:137
    }
};

// watershed.py:111-151 for every z slice at once.  label: segmentation (modified: overlaps / boundary cells zeroed),
// cover: cell_overlaps_mask (> 1 = overlap).  out = watershed(EDT of the overlaps, markers = label, mask = label > 0 | overlap).
template <class P>
void recalculate_cell_boundaries(P& pol, const ws::Dims& d, int* label, const int* cover, const int* on_boundary, int* out,
                                 const ws::Buffers& b) {
    const i64 n = d.n();
    pol.run(MarkersAndMasks{cover, on_boundary, label, b.mask, b.mask2}, n);          // mask = mask_image, mask2 = overlap
    pol.run(ws::ColDist{d, b.mask2, b.g}, n);                            // distance_transform_edt(overlap, (1, 1))
    pol.run(ws::RowDist{d, b.mask2, b.g, b.d2}, n);
    pol.run(ws::SqrtPlane{b.d2, b.fa}, n);
    ws::flood_from_labels(pol, d, b, b.mask, label, b.fa);
    pol.copy_i32(out, b.lab, n);
}

template <class P>
void accurate_correction(P& pol, const Cells& c, const ws::Dims& d, const float* prob, const void* raw, int raw_dtype,
                         double z_xy_ratio, const double* r_tracked_t0, const double* r_disp_prev,
                         const double* r_tracked_prev, const double* r_pred, const int* on_boundary, int max_rep,
                         double* r_disp, int* i_disp, int* i_disp_next, int* cell_of, int* cover, State* st) {
    const double zs = (double)c.z_scaling;
    const double zs_over_ratio = zs / z_xy_ratio, ratio_over_zs = z_xy_ratio / zs, inv_ratio = 1.0 / z_xy_ratio, inv_zs = 1.0 / zs;
    pol.run(CellOf{c, cell_of}, c.n_cells);
    pol.run(Begin{c.n_cells * 3, r_disp_prev, r_pred, r_tracked_prev, zs_over_ratio, r_disp, i_disp, st}, (i64)c.n_cells * 3);
    for (int rep = 0; rep < max_rep; ++rep) {
        pol.zero(cover, (size_t)d.n() * 4);
        pol.run(CountCover{c, d, i_disp, cell_of, cover, st}, c.n_vox);
        pol.run_cells(CellUpdate{c, d, prob, raw, raw_dtype, cover, on_boundary, r_tracked_t0, z_xy_ratio, inv_ratio, inv_zs,
                                 ratio_over_zs, zs_over_ratio, r_disp, i_disp, i_disp_next, st}, c.n_cells);
        pol.run(Decide{i_disp, i_disp_next, st}, (i64)c.n_cells * 3);
        pol.run(DecideFlag{st}, 1);
    }
}

// tracked_labels = recalculate_cell_boundaries(stamped labels, overlaps) on the raw grid.  `b`: watershed buffers.
template <class P>
void motion_to_image(P& pol, const Cells& c, const ws::Dims& d, const int* i_disp, const int* on_boundary, const int* cell_of,
                     int* cover, int* label, int* tracked_labels, const ws::Buffers& b) {
    const i64 n = d.n();
    pol.zero(cover, (size_t)n * 4);
    pol.zero(label, (size_t)n * 4);
    pol.run(CountCoverAlways{c, d, i_disp, cell_of, cover}, c.n_vox);
    pol.run(StampLabel{c, d, i_disp, cell_of, label}, c.n_vox);
    recalculate_cell_boundaries(pol, d, label, cover, on_boundary, tracked_labels, b);
}

}  // namespace corr
