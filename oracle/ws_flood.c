/* TEST INFRASTRUCTURE ONLY: C restatement of oracle/watershed.py::watershed (the priority flood of
 * skimage.segmentation.watershed, connectivity 1, no compactness, no watershed line) so that the CPU arm of bench.py is
 * not dominated by a pure-Python heap loop (scikit-image's own flood is compiled Cython).  Same order as the Python
 * restatement: (value, age, index) lexicographic, markers enter with age 0, one global age counter, neighbours in C
 * order of their offsets, a voxel is labelled when it is pushed.  tests/test_watershed_emul.py checks C == Python.
 * Built by __graft_entry__.build():  gcc -O2 -shared -fPIC -o oracle/_build/libws_flood.so oracle/ws_flood.c */
#include <stdint.h>
#include <stdlib.h>

typedef struct { double v; int64_t age; int64_t idx; } Elem;

static int less(const Elem* a, const Elem* b) {
    if (a->v != b->v) return a->v < b->v;
    if (a->age != b->age) return a->age < b->age;
    return a->idx < b->idx;
}
static void sift_down(Elem* h, int64_t n, int64_t k) {
    Elem e = h[k];
    for (;;) {
        int64_t c = 2 * k + 1;
        if (c >= n) break;
        if (c + 1 < n && less(&h[c + 1], &h[c])) ++c;
        if (!less(&h[c], &e)) break;
        h[k] = h[c];
        k = c;
    }
    h[k] = e;
}
static void sift_up(Elem* h, int64_t k) {
    Elem e = h[k];
    while (k > 0) {
        int64_t p = (k - 1) / 2;
        if (!less(&e, &h[p])) break;
        h[k] = h[p];
        k = p;
    }
    h[k] = e;
}

/* image (double), out (int64: markers * mask on entry, labels on exit), mask (uint8); ndim 2 or 3; shape[ndim]. */
int ws_flood(const double* image, int64_t* out, const uint8_t* mask, int ndim, const int64_t* shape) {
    int64_t dims[3] = {1, 1, 1}, n = 1;
    for (int a = 0; a < ndim; ++a) { dims[3 - ndim + a] = shape[a]; n *= shape[a]; }
    const int64_t sx = dims[1] * dims[2], sy = dims[2];
    Elem* h = (Elem*)malloc((size_t)(n > 0 ? n : 1) * sizeof(Elem));
    if (!h) return 1;
    int64_t cnt = 0, age = 0;
    for (int64_t i = 0; i < n; ++i)
        if (out[i]) { h[cnt].v = image[i]; h[cnt].age = 0; h[cnt].idx = i; ++cnt; }
    for (int64_t k = cnt / 2 - 1; k >= 0; --k) sift_down(h, cnt, k);
    while (cnt > 0) {
        const Elem top = h[0];
        --cnt;
        if (cnt > 0) { h[0] = h[cnt]; sift_down(h, cnt, 0); }
        const int64_t i = top.idx, x = i / sx, y = (i / sy) % dims[1], z = i % dims[2];
        const int64_t nb[6] = {i - sx, i - sy, i - 1, i + 1, i + sy, i + sx};
        const int ok[6] = {x > 0, y > 0, z > 0, z + 1 < dims[2], y + 1 < dims[1], x + 1 < dims[0]};
        for (int k = 0; k < 6; ++k) {
            if (!ok[k]) continue;
            const int64_t j = nb[k];
            if (!mask[j] || out[j]) continue;
            ++age;
            out[j] = out[i];
            h[cnt].v = image[j]; h[cnt].age = age; h[cnt].idx = j;
            sift_up(h, cnt);
            ++cnt;
        }
    }
    free(h);
    return 0;
}
