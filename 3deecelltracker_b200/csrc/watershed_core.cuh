// Watershed + centroid stage (Tracker._watershed, tracker.py:671-684; watershed.py:16-108; centre of mass
// tracker.py:646-648) as data-parallel passes over the (x, y, z) volume plus one priority flood per connected
// component of the foreground.
//
// Every pass is a functor `void operator()(long long i)` over a flat index range, launched through a Policy:
// watershed.cu instantiates the pipeline with a CUDA policy (one thread per index, global atomics); the CPU test
// harness tests/emul/ws_host.cpp instantiates THE SAME pipeline with a sequential policy so that the pass logic can be
// checked against the oracle on a machine without a GPU.  (The harness is test infrastructure: the product library
// only contains the CUDA instantiation.)
//
// Exactness.  The label map must be bit-identical to the CPU path, so every floating-point value that feeds a
// comparison is computed with the reference's own arithmetic:
//   * distance_transform_edt: exact integer squared distances in the plane (two separable passes), the z term added
//     as (double)(dx^2 + dy^2) + fl(fl(dz r) fl(dz r)) -- SciPy's `dt *= sampling; dt *= dt; add.reduce(axis 0)` -- and
//     a correctly rounded sqrt.
//   * gaussian_filter: SciPy's correlate1d for symmetric kernels, centre term first, then
//     `acc += (in[-j] + in[+j]) * w[j]` for j = radius .. 1, with separately rounded multiply and add (no FMA), zero
//     padding, axes in order x, y, z.  The weights come from the caller (NumPy's exp, as SciPy computes them).
//   * maximum filter / peak test / flood order only compare those values.
// Flood order (value, age, index) and the other choices scikit-image leaves open are stated in oracle/watershed.py.
#pragma once
#include <cstdint>
#include <cmath>

#if defined(__CUDACC__)
#define WS_HD __host__ __device__ __forceinline__
#else
#define WS_HD inline
#endif

namespace ws {

typedef long long i64;
constexpr int COL_INF = 30000;              // "no background in this column": 30000^2 + 30000^2 < 2^31

WS_HD int atomic_min_i(int* p, int v) {
#ifdef __CUDA_ARCH__
    return atomicMin(p, v);
#else
    const int o = *p; if (v < o) *p = v; return o;
#endif
}
WS_HD int atomic_add_i(int* p, int v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    const int o = *p; *p = o + v; return o;
#endif
}
WS_HD void atomic_add_u64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}
WS_HD void atomic_min_u64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
WS_HD void atomic_max_u64(unsigned long long* p, unsigned long long v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
WS_HD double mul_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
WS_HD double add_rn(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
WS_HD unsigned long long dbl_bits(double v) {
#ifdef __CUDA_ARCH__
    return (unsigned long long)__double_as_longlong(v);
#else
    unsigned long long u; __builtin_memcpy(&u, &v, 8); return u;
#endif
}
WS_HD double bits_dbl(unsigned long long u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double v; __builtin_memcpy(&v, &u, 8); return v;
#endif
}
WS_HD int vload(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

struct Dims {
    int X, Y, Z;
    WS_HD i64 n() const { return (i64)X * Y * Z; }
    WS_HD void split(i64 i, int& x, int& y, int& z) const {     // volumes have < 2^31 voxels: 32-bit division (the 64-bit
        unsigned u = (unsigned)i;                                // one costs ~100 instructions per thread on the GPU and was
        const unsigned q = u / (unsigned)Z;                      // the whole cost of the element-wise passes)
        z = (int)(u - q * (unsigned)Z);
        const unsigned r = q / (unsigned)Y;
        y = (int)(q - r * (unsigned)Y);
        x = (int)r;
    }
};

// ---------------------------------------------------------------------------------------------- threshold
struct Threshold {                               // watershed.py:38,49: image_pred > 0.5
    const float* prob; uint8_t* mask;
    WS_HD void operator()(i64 i) const { mask[i] = prob[i] > 0.5f ? 1 : 0; }
};

// ---------------------------------------------------------------------------------------------- in-plane EDT
// pass 1, one index per voxel: distance along y to the nearest background voxel of its column (COL_INF when the column
// has none).  Foreground voxels walk outwards until they meet background -- a handful of steps inside cell-sized
// regions -- instead of one thread scanning each whole column twice (17 920 threads for 512 x 512 x 35: half a warp per SM).
struct ColDist {
    Dims d; const uint8_t* mask; int* g;
    WS_HD void operator()(i64 i) const {
        if (!mask[i]) { g[i] = 0; return; }
        const int y = (int)(((unsigned)i / (unsigned)d.Z) % (unsigned)d.Y);
        int best = COL_INF;
        for (int o = 1; o < d.Y; ++o) {
            const bool lo = y - o >= 0, hi = y + o < d.Y;
            if (!lo && !hi) break;
            if ((lo && !mask[i - (i64)o * d.Z]) || (hi && !mask[i + (i64)o * d.Z])) { best = o; break; }
        }
        g[i] = best;
    }
};
// pass 2, one index per voxel: exact squared in-plane distance min over x' of (x - x')^2 + g(x', y)^2
struct RowDist {
    Dims d; const uint8_t* mask; const int* g; int* d2;
    WS_HD void operator()(i64 i) const {
        if (!mask[i]) { d2[i] = 0; return; }
        int x, y, z; d.split(i, x, y, z);
        const i64 sx = (i64)d.Y * d.Z;
        int best = g[i] * g[i];
        for (int o = 1; o * o < best; ++o) {
            if (x - o >= 0) { const int v = g[i - o * sx]; const int c = o * o + v * v; if (c < best) best = c; }
            if (x + o < d.X) { const int v = g[i + o * sx]; const int c = o * o + v * v; if (c < best) best = c; }
            if (x - o < 0 && x + o >= d.X) break;
        }
        d2[i] = best;
    }
};
struct SqrtPlane {                               // distance_transform_edt(bn_image, sampling=[1, 1])
    const int* d2; double* dist;
    WS_HD void operator()(i64 i) const { dist[i] = sqrt((double)d2[i]); }
};
// distance_transform_edt(volume, sampling=[1, 1, r]) from the in-plane squared distances of every slice
struct DistZ {
    Dims d; const uint8_t* mask; const int* d2; double r; double* dist;
    WS_HD void operator()(i64 i) const {
        if (!mask[i]) { dist[i] = 0.0; return; }
        int x, y, z; d.split(i, x, y, z);
        const i64 col = i - z;
        double best = (double)d2[i];
        for (int zz = 0; zz < d.Z; ++zz) {
            if (zz == z) continue;
            const double dz = mul_rn((double)(zz - z), r);
            const double c = add_rn((double)d2[col + zz], mul_rn(dz, dz));
            if (c < best) best = c;
        }
        dist[i] = sqrt(best);
    }
};

// ---------------------------------------------------------------------------------------------- activity grid
// Foreground is a few per cent of a volume, and everything the smoothing / peak passes compute is exactly 0 farther than
// 15 voxels (Gaussian radius 8 + maximum-filter radius 7) from it.  The (x, y) plane is cut into 16 x 16 cells per z; a
// cell is active when a foreground voxel lies in its 3 x 3 cell neighbourhood (3-D stage: also within 4 slices).  The
// stencil passes write 0 for voxels of inactive cells without reading anything -- same values, a fraction of the work.
constexpr int ACT_CELL = 16;
struct ActGrid {
    const uint8_t* act; int cy;                  // cells along y
    WS_HD bool on(int x, int y, int z, int Z) const {
        return act == nullptr || act[((i64)(x / ACT_CELL) * cy + (y / ACT_CELL)) * Z + z] != 0;
    }
};
struct ActMark {
    Dims d; const uint8_t* mask; uint8_t* act0; int cy;
    WS_HD void operator()(i64 i) const {
        if (!mask[i]) return;
        int x, y, z; d.split(i, x, y, z);
        act0[((i64)(x / ACT_CELL) * cy + (y / ACT_CELL)) * d.Z + z] = 1;
    }
};
struct ActDilate {                               // one index per cell-voxel (cx, cy, z)
    int cx, cy, Z; const uint8_t* in; uint8_t* out; int rz;      // OR over the 3 x 3 cell neighbourhood and z +- rz
    WS_HD void operator()(i64 c) const {
        const int z = (int)(c % Z); const int b = (int)((c / Z) % cy), a = (int)(c / ((i64)Z * cy));
        uint8_t v = 0;
        for (int da = -1; da <= 1; ++da)
            for (int db = -1; db <= 1; ++db)
                for (int dz = -rz; dz <= rz; ++dz) {
                    const int aa = a + da, bb = b + db, zz = z + dz;
                    if (aa < 0 || aa >= cx || bb < 0 || bb >= cy || zz < 0 || zz >= Z) continue;
                    v |= in[((i64)aa * cy + bb) * Z + zz];
                }
        out[c] = v;
    }
};

// ---------------------------------------------------------------------------------------------- gaussian_filter
template <int AXIS>
struct Gauss1D {                                 // scipy.ndimage.correlate1d, symmetric kernel, mode='constant'
    Dims d; const double* in; double* out; int radius; double w[9]; ActGrid ag;
    WS_HD void operator()(i64 i) const {
        int c[3]; d.split(i, c[0], c[1], c[2]);
        if (!ag.on(c[0], c[1], c[2], d.Z)) { out[i] = 0.0; return; }
        const int len = AXIS == 0 ? d.X : (AXIS == 1 ? d.Y : d.Z);
        const i64 st = AXIS == 0 ? (i64)d.Y * d.Z : (AXIS == 1 ? (i64)d.Z : 1);
        const int p = c[AXIS];
        double acc = mul_rn(in[i], w[0]);
        for (int j = radius; j >= 1; --j) {
            const double l = p - j >= 0 ? in[i - j * st] : 0.0;
            const double h = p + j < len ? in[i + j * st] : 0.0;
            acc = add_rn(acc, mul_rn(add_rn(l, h), w[j]));
        }
        out[i] = acc;
    }
};

// ---------------------------------------------------------------------------------------------- peak_local_max
template <int AXIS>
struct Max1D {                                   // maximum_filter(size = 2 r + 1, mode='constant'): values are >= 0
    Dims d; const double* in; double* out; int radius; ActGrid ag;
    WS_HD void operator()(i64 i) const {
        int c[3]; d.split(i, c[0], c[1], c[2]);
        if (!ag.on(c[0], c[1], c[2], d.Z)) { out[i] = 0.0; return; }
        const int len = AXIS == 0 ? d.X : (AXIS == 1 ? d.Y : d.Z);
        const i64 st = AXIS == 0 ? (i64)d.Y * d.Z : (AXIS == 1 ? (i64)d.Z : 1);
        const int p = c[AXIS];
        const int lo = p - radius < 0 ? 0 : p - radius, hi = p + radius >= len ? len - 1 : p + radius;
        double m = (p - radius < 0 || p + radius >= len) ? 0.0 : in[i];          // zero padding takes part
        for (int q = lo; q <= hi; ++q) { const double v = in[i + (q - p) * st]; if (v > m) m = v; }
        out[i] = m;
    }
};
struct MinReduce {                               // image.min(): per slice (2-D stage) or global (3-D stage)
    Dims d; const double* in; unsigned long long* slot; int per_slice;
    WS_HD void operator()(i64 i) const {
        unsigned long long* s = slot + (per_slice ? (int)((unsigned)i % (unsigned)d.Z) : 0);
        const unsigned long long b = dbl_bits(in[i]);                               // values >= +0: bit order = value order
        // almost every voxel is background (value 0 = the minimum): only values below the slot's current content go to
        // the atomic unit (a stale read can only cause a redundant atomic, never a missed one)
        if (b < *s) atomic_min_u64(s, b);
    }
};
struct Peaks {
    Dims d; const double* img; const double* imgmax; const unsigned long long* minslot; int per_slice; int border;
    uint8_t* peak; ActGrid ag;
    WS_HD void operator()(i64 i) const {
        int x, y, z; d.split(i, x, y, z);
        if (!ag.on(x, y, z, d.Z)) { peak[i] = 0; return; }
        const double thr = bits_dbl(minslot[per_slice ? z : 0]);
        bool p = img[i] == imgmax[i] && img[i] > thr;
        if (border > 0 && (x < border || x >= d.X - border || y < border || y >= d.Y - border)) p = false;
        peak[i] = p ? 1 : 0;
    }
};

// ---------------------------------------------------------------------------------------------- connected components
// Lock-free union-find on voxel indices (root = smallest index of the set = first voxel in raster order).
WS_HD int uf_find(const int* L, int a) {
    int p = vload(L + a);
    while (p != a) { a = p; p = vload(L + a); }
    return a;
}
WS_HD void uf_union(int* L, int a, int b) {
    bool done;
    do {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a < b) { const int old = atomic_min_i(L + b, a); done = (old == b); b = old; }
        else if (b < a) { const int old = atomic_min_i(L + a, b); done = (old == a); a = old; }
        else done = true;
    } while (!done);
}
struct UfInit {
    const uint8_t* on; int* L;
    WS_HD void operator()(i64 i) const { L[i] = on[i] ? (int)i : -1; }
};
// FULL = 0: connectivity 1 (4 in the plane / 6 in the volume); FULL = 1: full connectivity (8 / 26).  planar = no z links.
struct UfLink {
    Dims d; const uint8_t* on; int* L; int full; int planar;
    WS_HD void operator()(i64 i) const {
        if (!on[i]) return;
        int x, y, z; d.split(i, x, y, z);
        // the 13 (or 4 in the plane) neighbours that precede voxel i in raster order
        for (int dx = -1; dx <= 0; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    if (dx == 0 && (dy > 0 || (dy == 0 && dz >= 0))) continue;
                    if (planar && dz != 0) continue;
                    if (!full && (dx != 0) + (dy != 0) + (dz != 0) != 1) continue;
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || yy < 0 || yy >= d.Y || zz < 0 || zz >= d.Z) continue;
                    const i64 j = ((i64)xx * d.Y + yy) * d.Z + zz;
                    if (on[j]) uf_union(L, (int)i, (int)j);
                }
    }
};
// components of EQUAL non-zero value, full connectivity: skimage.measure.label(label image, connectivity = ndim)
struct UfInitValue {
    const int* val; int* L;
    WS_HD void operator()(i64 i) const { L[i] = val[i] != 0 ? (int)i : -1; }
};
struct UfLinkEqual {
    Dims d; const int* val; int* L;
    WS_HD void operator()(i64 i) const {
        const int v = val[i];
        if (v == 0) return;
        int x, y, z; d.split(i, x, y, z);
        for (int dx = -1; dx <= 0; ++dx)
            for (int dy = -1; dy <= 1; ++dy)
                for (int dz = -1; dz <= 1; ++dz) {
                    if (dx == 0 && (dy > 0 || (dy == 0 && dz >= 0))) continue;
                    const int xx = x + dx, yy = y + dy, zz = z + dz;
                    if (xx < 0 || yy < 0 || yy >= d.Y || zz < 0 || zz >= d.Z) continue;
                    const i64 j = ((i64)xx * d.Y + yy) * d.Z + zz;
                    if (val[j] == v) uf_union(L, (int)i, (int)j);
                }
    }
};
struct UfFlattenValue {
    const int* val; int* L; int* rootflag;
    WS_HD void operator()(i64 i) const {
        if (val[i] != 0) { const int r = uf_find(L, (int)i); L[i] = r; rootflag[i] = (r == (int)i) ? 1 : 0; }
        else rootflag[i] = 0;
    }
};
struct RankToLabel {                             // component -> 1 + number of components that start earlier in raster order
    const int* val; const int* L; const int* rank; int* out;
    WS_HD void operator()(i64 i) const { out[i] = val[i] != 0 ? rank[L[i]] + 1 : 0; }
};
struct UfFlatten {
    const uint8_t* on; int* L;
    WS_HD void operator()(i64 i) const { if (on[i]) L[i] = uf_find(L, (int)i); }
};

// ---------------------------------------------------------------------------------------------- flood
struct HeapE { double v; int age; int idx; };
WS_HD bool heap_less(const HeapE& a, const HeapE& b) {
    if (a.v != b.v) return a.v < b.v;
    if (a.age != b.age) return a.age < b.age;
    return a.idx < b.idx;
}
WS_HD void heap_down(HeapE* h, int n, int k) {
    const HeapE e = h[k];
    for (;;) {
        int c = 2 * k + 1;
        if (c >= n) break;
        if (c + 1 < n && heap_less(h[c + 1], h[c])) ++c;
        if (!heap_less(h[c], e)) break;
        h[k] = h[c];
        k = c;
    }
    h[k] = e;
}
WS_HD void heap_up(HeapE* h, int k) {
    const HeapE e = h[k];
    while (k > 0) {
        const int p = (k - 1) >> 1;
        if (!heap_less(e, h[p])) break;
        h[k] = h[p];
        k = p;
    }
    h[k] = e;
}
struct CompSize {                                // voxels per foreground component, accumulated at the root
    const uint8_t* mask; const int* comp; int* csize;
    WS_HD void operator()(i64 i) const { if (mask[i]) atomic_add_i(csize + comp[i], 1); }
};
struct HeapAlloc {                               // every root reserves heap room for its whole component and joins the list
    const uint8_t* mask; const int* comp; const int* csize; int* hoff; int* hcnt; int* counter; int* roots; int* n_roots;
    WS_HD void operator()(i64 i) const {
        if (mask[i] && comp[i] == (int)i) {
            hoff[i] = atomic_add_i(counter, csize[i]); hcnt[i] = 0;
            roots[atomic_add_i(n_roots, 1)] = (int)i;
        }
    }
};
struct SeedMarkers {                             // markers = label(local_maxi) * mask; seeds enter with age 0
    const uint8_t* mask; const uint8_t* peak; const int* mk; const int* comp; const double* img;
    const int* hoff; int* hcnt; HeapE* heap; int* lab;
    WS_HD void operator()(i64 i) const {
        if (mask[i] && peak[i]) {
            const int r = comp[i];
            const int slot = atomic_add_i(hcnt + r, 1);
            HeapE e; e.v = -img[i]; e.age = 0; e.idx = (int)i;      // the stage floods -dist_smooth (watershed.py:44,94)
            heap[hoff[r] + slot] = e;
            lab[i] = mk[i] + 1;                  // label id = 1 + first voxel of the marker's plateau
        } else {
            lab[i] = 0;
        }
    }
};
struct SeedLabels {                              // markers given as a label image (recalculate_cell_boundaries,
    const uint8_t* mask; const int* markers; const int* comp; const double* img;   // watershed.py:111-151): floods +img
    const int* hoff; int* hcnt; HeapE* heap; int* lab;
    WS_HD void operator()(i64 i) const {
        const int m = mask[i] ? markers[i] : 0;
        if (m != 0) {
            const int r = comp[i];
            const int slot = atomic_add_i(hcnt + r, 1);
            HeapE e; e.v = img[i]; e.age = 0; e.idx = (int)i;
            heap[hoff[r] + slot] = e;
        }
        lab[i] = m;
    }
};
struct Flood {                                   // skimage.segmentation.watershed(sign * img, markers, mask), connectivity 1
    Dims d; const uint8_t* mask; const int* comp; const double* img; const int* hoff; const int* hcnt;
    HeapE* heap; int* lab; int planar; double sign;
    WS_HD void operator()(i64 i) const {
        if (!mask[i] || comp[i] != (int)i) return;
        int n = hcnt[i];
        if (n == 0) return;
        HeapE* h = heap + hoff[i];
        for (int k = n / 2 - 1; k >= 0; --k) heap_down(h, n, k);
        const i64 sx = (i64)d.Y * d.Z, sy = d.Z;
        int age = 0;
        while (n > 0) {
            const HeapE top = h[0];
            --n;
            if (n > 0) { h[0] = h[n]; heap_down(h, n, 0); }
            int x, y, z; d.split(top.idx, x, y, z);
            const int l = lab[top.idx];
            // C order of the neighbour offsets: x-1, y-1, z-1, z+1, y+1, x+1
            const i64 nb[6] = {top.idx - sx, top.idx - sy, (i64)top.idx - 1, (i64)top.idx + 1, top.idx + sy, top.idx + sx};
            const bool ok[6] = {x > 0, y > 0, !planar && z > 0, !planar && z + 1 < d.Z, y + 1 < d.Y, x + 1 < d.X};
            for (int k = 0; k < 6; ++k) {
                if (!ok[k]) continue;
                const i64 j = nb[k];
                if (!mask[j] || lab[j] != 0) continue;
                ++age;
                lab[j] = l;
                HeapE e; e.v = sign * img[j]; e.age = age; e.idx = (int)j;
                h[n] = e;
                heap_up(h, n);
                ++n;
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------- find_boundaries
struct Boundary2D {                              // find_boundaries(labels, connectivity=2, mode='outer', background=0)
    Dims d; const uint8_t* mask; const int* lab; uint8_t* out;      // out = mask & ~boundary  (watershed.py:49-50)
    WS_HD void operator()(i64 i) const {
        int x, y, z; d.split(i, x, y, z);
        const int l = lab[i];
        int mx = l, mn_nz = l ? l : 0x7fffffff;
        for (int dx = -1; dx <= 1; ++dx)
            for (int dy = -1; dy <= 1; ++dy) {
                const int xx = x + dx, yy = y + dy;
                if (xx < 0 || xx >= d.X || yy < 0 || yy >= d.Y) continue;
                const int v = lab[((i64)xx * d.Y + yy) * d.Z + z];
                if (v > mx) mx = v;
                if (v != 0 && v < mn_nz) mn_nz = v;
            }
        // background voxel: boundary iff a label is adjacent; labelled voxel: iff two different labels are adjacent
        const bool boundary = l == 0 ? mx != 0 : mx != mn_nz;
        out[i] = (mask[i] && !boundary) ? 1 : 0;
    }
};

// ---------------------------------------------------------------------------------------------- sizes, relabel, centres
struct Scalars {                                 // device-resident scalars of one call
    int min_size, cell_num, n_labels, n_cells, bg_count, heap_counter, count_ge, pad;
};
struct LabelSize {                               // np.bincount(labels_ws.ravel()): voxels per label, kept at the label's root
    const int* lab; int* lsize;
    WS_HD void operator()(i64 i) const {
        const int l = lab[i];
        if (l) atomic_add_i(lsize + (l - 1), 1);
    }
};
struct BackgroundCount {                         // bincount[0] = all voxels - labelled voxels (one atomic per label, not per voxel)
    const uint8_t* peak; const int* mk; const int* lsize; int* labelled;
    WS_HD void operator()(i64 i) const {
        if (peak[i] && mk[i] == (int)i && lsize[i] > 0) atomic_add_i(labelled, lsize[i]);
    }
};
struct BackgroundFinish {
    Scalars* sc; i64 n;
    WS_HD void operator()(i64) const { sc->bg_count = (int)(n - (i64)sc->count_ge); sc->count_ge = 0; }
};
struct CountGE {                                 // number of marker labels with at least `thr` voxels (roots only)
    const uint8_t* peak; const int* mk; const int* lsize; const int* thr; int* out;
    WS_HD void operator()(i64 i) const {
        if (peak[i] && mk[i] == (int)i && lsize[i] >= *thr) atomic_add_i(out, 1);
    }
};
// min_size / cell_num (watershed.py:95-98).  One-index passes drive the logic on the device, the counting itself is
// the data-parallel CountGE pass.  thr[0] = probe threshold, thr[1] = #(marker labels with >= thr[0] voxels),
// thr[2..3] = binary-search bounds.
struct MinSizeBegin {
    Scalars* sc; int* thr; int method, min_size, cell_num;
    WS_HD void operator()(i64) const {
        sc->min_size = min_size; sc->cell_num = cell_num;
        thr[0] = method == 0 ? min_size : 0; thr[1] = 0; thr[2] = 0; thr[3] = 0x7ffffffe;
        if (method != 0) thr[0] = thr[2] + (thr[3] - thr[2] + 1) / 2;
    }
};
struct MinSizeFromCount {                        // "min_size": cell_num = #(bincount >= min_size) - 1, background bin included
    Scalars* sc; const int* thr;
    WS_HD void operator()(i64) const { sc->cell_num = thr[1] + (sc->bg_count >= sc->min_size ? 1 : 0) - 1; }
};
struct MinSizeSearchStep {                       // "cell_num": min_size = sorted(bincount)[-cell_num - 1]
    Scalars* sc; int* thr;                       //           = the largest t with #(bincount >= t) >= cell_num + 1
    WS_HD void operator()(i64) const {
        int lo = thr[2], hi = thr[3];
        if (lo < hi) {
            const int cnt = thr[1] + (sc->bg_count >= thr[0] ? 1 : 0);
            if (cnt >= sc->cell_num + 1) lo = thr[0]; else hi = thr[0] - 1;
        }
        thr[2] = lo; thr[3] = hi;
        thr[0] = lo < hi ? lo + (hi - lo + 1) / 2 : lo;
        thr[1] = 0;
        sc->min_size = lo;
    }
};
struct KeepFlag {                                // remove_small_objects: labels with fewer than min_size voxels go
    const uint8_t* peak; const int* mk; const int* lsize; const Scalars* sc; int* flag;
    WS_HD void operator()(i64 i) const {
        flag[i] = (peak[i] && mk[i] == (int)i && lsize[i] > 0 && lsize[i] >= sc->min_size) ? 1 : 0;
    }
};
struct Relabel {                                 // relabel_sequential + centre-of-mass sums (tracker.py:646-648, :682)
    Dims d; const int* lab; const int* flag; const int* rank; int* out; unsigned long long* sums; int max_cells;
    WS_HD void operator()(i64 i) const {
        const int l = lab[i];
        int id = 0;
        if (l && flag[l - 1]) id = rank[l - 1] + 1;
        out[i] = id;
        if (id && id <= max_cells) {
            int x, y, z; d.split(i, x, y, z);
            unsigned long long* s = sums + (i64)(id - 1) * 4;
            atomic_add_u64(s + 0, (unsigned long long)x);
            atomic_add_u64(s + 1, (unsigned long long)y);
            atomic_add_u64(s + 2, (unsigned long long)z);
            atomic_add_u64(s + 3, 1ull);
        }
    }
};
struct Centres {                                 // centres[0] = voxel units (l_center_coordinates), centres[1] = real units
    const unsigned long long* sums; double* centres; const Scalars* sc; int max_cells; double z_xy_ratio;
    WS_HD void operator()(i64 k) const {         // (r_coordinates_segment = _transform_layer_to_real, tracker.py:648)
        if (k >= sc->n_cells || k >= max_cells) return;
        const double n = (double)sums[k * 4 + 3];
        const double cx = (double)sums[k * 4 + 0] / n, cy = (double)sums[k * 4 + 1] / n, cz = (double)sums[k * 4 + 2] / n;
        double* real = centres + (i64)max_cells * 3;
        centres[k * 3 + 0] = cx; centres[k * 3 + 1] = cy; centres[k * 3 + 2] = cz;
        real[k * 3 + 0] = cx; real[k * 3 + 1] = cy; real[k * 3 + 2] = mul_rn(cz, z_xy_ratio);
    }
};

// ---------------------------------------------------------------------------------------------- workspace
struct Buffers {
    uint8_t *mask, *mask2, *peak, *act0, *act2, *act3;
    int *g, *d2, *mk, *comp, *lab, *csize, *hoff, *hcnt, *flag, *rank;
    double *fa, *fb, *fc;
    HeapE* heap;
    unsigned long long* minslot;     // Z + 1 slots
    unsigned long long* sums;        // max_cells x 4
    Scalars* sc;
    int* thr;                        // 2 ints: binary-search threshold, count
};
inline size_t a256(size_t v) { return (v + 255) / 256 * 256; }
inline size_t act_bytes(i64 n, int Z) { return (size_t)(n / Z / (ACT_CELL * ACT_CELL) + 2 * 4096 + 64) * Z; }   // generous bound
inline size_t workspace_bytes(i64 n, int Z, int max_cells) {
    size_t t = 3 * a256(act_bytes(n, Z));
    t += 3 * a256((size_t)n);                       // mask, mask2, peak
    t += 10 * a256((size_t)n * 4);                  // g, d2, mk, comp, lab, csize, hoff, hcnt, flag, rank
    t += 3 * a256((size_t)n * 8);                   // fa, fb, fc
    t += a256((size_t)n * sizeof(HeapE));
    t += a256((size_t)(Z + 1) * 8) + a256((size_t)max_cells * 32) + a256(sizeof(Scalars)) + a256(64);
    return t + 256;
}
inline void carve(Buffers& b, void* ws, i64 n, int Z, int max_cells) {
    char* p = reinterpret_cast<char*>(((uintptr_t)ws + 255) / 256 * 256);
    auto take = [&](size_t bytes) { char* r = p; p += a256(bytes); return r; };
    b.mask = (uint8_t*)take(n); b.mask2 = (uint8_t*)take(n); b.peak = (uint8_t*)take(n);
    b.act0 = (uint8_t*)take(act_bytes(n, Z)); b.act2 = (uint8_t*)take(act_bytes(n, Z)); b.act3 = (uint8_t*)take(act_bytes(n, Z));
    b.g = (int*)take(n * 4); b.d2 = (int*)take(n * 4); b.mk = (int*)take(n * 4); b.comp = (int*)take(n * 4);
    b.lab = (int*)take(n * 4); b.csize = (int*)take(n * 4); b.hoff = (int*)take(n * 4); b.hcnt = (int*)take(n * 4);
    b.flag = (int*)take(n * 4); b.rank = (int*)take(n * 4);
    b.fa = (double*)take(n * 8); b.fb = (double*)take(n * 8); b.fc = (double*)take(n * 8);
    b.heap = (HeapE*)take(n * sizeof(HeapE));
    b.minslot = (unsigned long long*)take((size_t)(Z + 1) * 8);
    b.sums = (unsigned long long*)take((size_t)max_cells * 32);
    b.sc = (Scalars*)take(sizeof(Scalars));
    b.thr = (int*)take(64);
}

struct Params {
    int X, Y, Z;
    double z_xy_ratio;
    int method;                       // 0 = "min_size", 1 = "cell_num" (watershed.py:95-98)
    int min_size, cell_num;
    int max_cells;
    double w_xy[9];                   // gaussian weights sigma = 2, radius 8: w[j] = weight at offset +-j
    double w_z[2];                    // sigma = 0.3, radius 1
};

// One flood stage: markers -> labels.  `img` = smoothed distance, `fg` = foreground, planar = per-slice (2-D stage).
template <class P>
void flood_stage(P& pol, const Dims& d, const Buffers& b, const uint8_t* fg, const double* img, int planar) {
    const i64 n = d.n();
    pol.run(UfInit{b.peak, b.mk}, n);
    pol.run(UfLink{d, b.peak, b.mk, 1, planar}, n);
    pol.run(UfFlatten{b.peak, b.mk}, n);
    pol.run(UfInit{fg, b.comp}, n);
    pol.run(UfLink{d, fg, b.comp, 0, planar}, n);
    pol.run(UfFlatten{fg, b.comp}, n);
    pol.zero(b.csize, (size_t)n * 4);
    pol.zero(&b.sc->heap_counter, 4);
    pol.zero(&b.sc->n_labels, 4);
    pol.run(CompSize{fg, b.comp, b.csize}, n);
    pol.run(HeapAlloc{fg, b.comp, b.csize, b.hoff, b.hcnt, &b.sc->heap_counter, b.rank, &b.sc->n_labels}, n);
    pol.run(SeedMarkers{fg, b.peak, b.mk, b.comp, img, b.hoff, b.hcnt, b.heap, b.lab}, n);
    pol.run_flood(Flood{d, fg, b.comp, img, b.hoff, b.hcnt, b.heap, b.lab, planar, -1.0}, b.rank, &b.sc->n_labels, b.csize, n);
}

// watershed(img, markers, mask) per z slice with a given marker label image (recalculate_cell_boundaries).
template <class P>
void flood_from_labels(P& pol, const Dims& d, const Buffers& b, const uint8_t* fg, const int* markers, const double* img) {
    const i64 n = d.n();
    pol.run(UfInit{fg, b.comp}, n);
    pol.run(UfLink{d, fg, b.comp, 0, 1}, n);
    pol.run(UfFlatten{fg, b.comp}, n);
    pol.zero(b.csize, (size_t)n * 4);
    pol.zero(&b.sc->heap_counter, 4);
    pol.zero(&b.sc->n_labels, 4);
    pol.run(CompSize{fg, b.comp, b.csize}, n);
    pol.run(HeapAlloc{fg, b.comp, b.csize, b.hoff, b.hcnt, &b.sc->heap_counter, b.rank, &b.sc->n_labels}, n);
    pol.run(SeedLabels{fg, markers, b.comp, img, b.hoff, b.hcnt, b.heap, b.lab}, n);
    pol.run_flood(Flood{d, fg, b.comp, img, b.hoff, b.hcnt, b.heap, b.lab, 1, 1.0}, b.rank, &b.sc->n_labels, b.csize, n);
}

// skimage.measure.label(int image, connectivity = 3): out = component ids 1..n in raster order of their first voxel
template <class P>
void label_equal_values(P& pol, const Dims& d, const int* val, int* out, const Buffers& b, int* n_out) {
    const i64 n = d.n();
    pol.run(UfInitValue{val, b.comp}, n);
    pol.run(UfLinkEqual{d, val, b.comp}, n);
    pol.run(UfFlattenValue{val, b.comp, b.flag}, n);
    pol.exclusive_scan(b.flag, b.rank, n, n_out);
    pol.run(RankToLabel{val, b.comp, b.rank, out}, n);
}

// The whole stage.  prob (x,y,z) float32 -> labels (x,y,z) int32, centres (2,max_cells,3) float64, scalars.
template <class P>
void segment(P& pol, const Params& prm, const float* prob, int* labels, double* centres, const Buffers& b) {
    const Dims d{prm.X, prm.Y, prm.Z};
    const i64 n = d.n();
    const int cxn = (d.X + ACT_CELL - 1) / ACT_CELL, cyn = (d.Y + ACT_CELL - 1) / ACT_CELL;
    const i64 ncell = (i64)cxn * cyn * d.Z;
    const ActGrid a2{b.act2, cyn}, a3{b.act3, cyn};
    Gauss1D<0> gx{d, nullptr, nullptr, 8, {}, a2};
    Gauss1D<1> gy{d, nullptr, nullptr, 8, {}, a2};
    Gauss1D<2> gz{d, nullptr, nullptr, 1, {}, a3};
    for (int j = 0; j < 9; ++j) { gx.w[j] = prm.w_xy[j]; gy.w[j] = prm.w_xy[j]; gz.w[j] = j < 2 ? prm.w_z[j] : 0.0; }

    // ---- watershed_2d (watershed.py:16-52), all slices at once
    pol.run(Threshold{prob, b.mask}, n);
    pol.zero(b.act0, (size_t)ncell);
    pol.run(ActMark{d, b.mask, b.act0, cyn}, n);
    pol.run(ActDilate{cxn, cyn, d.Z, b.act0, b.act2, 0}, ncell);
    pol.run(ActDilate{cxn, cyn, d.Z, b.act0, b.act3, 4}, ncell);
    pol.run(ColDist{d, b.mask, b.g}, n);
    pol.run(RowDist{d, b.mask, b.g, b.d2}, n);
    pol.run(SqrtPlane{b.d2, b.fa}, n);
    gx.in = b.fa; gx.out = b.fb; pol.run(gx, n);
    gy.in = b.fb; gy.out = b.fa; pol.run(gy, n);                              // fa = dist_smooth
    pol.fill_u64(b.minslot, 0x7ff0000000000000ull, d.Z + 1);
    pol.run(MinReduce{d, b.fa, b.minslot, 1}, n);
    pol.run(Max1D<0>{d, b.fa, b.fb, 7, a2}, n);
    pol.run(Max1D<1>{d, b.fb, b.fc, 7, a2}, n);
    pol.run(Peaks{d, b.fa, b.fc, b.minslot, 1, 7, b.peak, a2}, n);
    flood_stage(pol, d, b, b.mask, b.fa, 1);
    pol.run(Boundary2D{d, b.mask, b.lab, b.mask2}, n);                         // mask2 = bn_output

    // ---- watershed_3d (watershed.py:55-101)
    pol.run(ColDist{d, b.mask2, b.g}, n);
    pol.run(RowDist{d, b.mask2, b.g, b.d2}, n);
    pol.run(DistZ{d, b.mask2, b.d2, prm.z_xy_ratio, b.fa}, n);
    gx.ag = a3; gy.ag = a3;
    gx.in = b.fa; gx.out = b.fb; pol.run(gx, n);
    gy.in = b.fb; gy.out = b.fa; pol.run(gy, n);
    gz.in = b.fa; gz.out = b.fb; pol.run(gz, n);                              // fb = dist_smooth
    pol.fill_u64(b.minslot, 0x7ff0000000000000ull, d.Z + 1);
    pol.run(MinReduce{d, b.fb, b.minslot, 0}, n);
    pol.run(Max1D<0>{d, b.fb, b.fa, 3, a3}, n);
    pol.run(Max1D<1>{d, b.fa, b.fc, 3, a3}, n);
    pol.run(Max1D<2>{d, b.fc, b.fa, 3, a3}, n);
    pol.run(Peaks{d, b.fb, b.fa, b.minslot, 0, 0, b.peak, a3}, n);
    flood_stage(pol, d, b, b.mask2, b.fb, 0);

    // ---- sizes, min_size / cell_num, remove_small_objects, relabel_sequential, centres
    int* lsize = b.csize;                                                      // reuse: voxels per marker label (at its root)
    pol.zero(lsize, (size_t)n * 4);
    pol.zero(b.sc, sizeof(Scalars));
    pol.run(LabelSize{b.lab, lsize}, n);
    pol.run(BackgroundCount{b.peak, b.mk, lsize, &b.sc->count_ge}, n);
    pol.run(BackgroundFinish{b.sc, n}, 1);
    pol.run(MinSizeBegin{b.sc, b.thr, prm.method, prm.min_size, prm.cell_num}, 1);
    if (prm.method == 0) {
        pol.run(CountGE{b.peak, b.mk, lsize, b.thr, b.thr + 1}, n);
        pol.run(MinSizeFromCount{b.sc, b.thr}, 1);
    } else {
        for (int it = 0; it < 32; ++it) {
            pol.run(CountGE{b.peak, b.mk, lsize, b.thr, b.thr + 1}, n);
            pol.run(MinSizeSearchStep{b.sc, b.thr}, 1);
        }
    }
    pol.run(KeepFlag{b.peak, b.mk, lsize, b.sc, b.flag}, n);
    pol.exclusive_scan(b.flag, b.rank, n, &b.sc->n_cells);
    pol.zero(b.sums, (size_t)prm.max_cells * 32);
    pol.run(Relabel{d, b.lab, b.flag, b.rank, labels, b.sums, prm.max_cells}, n);
    pol.run(Centres{b.sums, centres, b.sc, prm.max_cells, prm.z_xy_ratio}, prm.max_cells);
}

}  // namespace ws
