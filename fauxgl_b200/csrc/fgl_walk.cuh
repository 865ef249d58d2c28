// fgl_walk.cuh -- the scanline walker shared by the two front ends (fgl_span.cu:
// grid-wide span stage of small draws; fgl_geom.cu: k_front, the fused geometry +
// span kernel of large draws).
//
// One call replays the reference's inner loop for one row of one triangle
// (context.go:184-221): the per-row adds from y0, the skip-ahead `d`, then the
// per-pixel forward-differencing adds, exactly as written, once, left to right.
// The covered run is cut at strip boundaries (64 pixels) into segments that carry
// the edge values at their first pixel, so the back end continues the same chain of
// adds and reproduces coverage, barycentrics and depth bit for bit.
#pragma once
#include "fgl_internal.h"
#include "fgl_math.cuh"

namespace fgl {

// The per-record fields the back end needs besides the walk itself.
struct RecTail { double r0, r1, r2; uint32_t src, flags; };

__device__ __forceinline__ SegV make_segv(double w0, double w1, double w2, double ra, double z0, double z1, double z2,
                                          double a12, double a20, double a01, const RecTail &t, uint16_t x, uint8_t cnt,
                                          uint32_t wrap = 0) {
    SegV v;
    v.w0 = w0; v.w1 = w1; v.w2 = w2; v.ra = ra; v.z0 = z0; v.z1 = z1; v.z2 = z2;
    v.a12 = a12; v.a20 = a20; v.a01 = a01;
    v.r0 = t.r0; v.r1 = t.r1; v.r2 = t.r2; v.src = t.src; v.flags = t.flags | wrap;
    v.x = x; v.yt = 0; v.cnt = cnt; v._pad0 = 0; v._pad[0] = v._pad[1] = 0;
    return v;
}

// First segment of a row, kept in registers between the counting walk and the write.
struct ParkedSeg {
    double w0, w1, w2;
    int32_t x;      // first covered pixel: column in the row it lands in
    uint32_t cnt;   // covered pixels
    uint32_t key;   // strip id
    uint32_t wrap;  // REC_WRAP if the segment lies outside [0, width) of its own row (it aliases into a neighbour row)
};

// Which pixels of row y the reference keeps (context.go:223-228): i = y*W + x must satisfy 0 <= i < W*H; x itself
// is never range-checked, so a covered pixel left or right of the framebuffer aliases into a neighbouring row --
// it lands on pixel (i mod W, i div W): depth is tested and written there, a blended colour too (PixOffset,
// :259), an opaque colour is dropped by SetNRGBA's bounds check (:269).  With state.x_guard such pixels are
// dropped instead.  [xlo, xhi] is the range of x that is kept in row y (empty when xlo > xhi).
__device__ __forceinline__ void row_x_range(const DrawParams &p, int y, long long &xlo, long long &xhi) {
    if (p.state.x_guard) {
        const bool on = y >= 0 && y < p.height;
        xlo = 0; xhi = on ? (long long)p.width - 1 : -1;
    } else {
        xlo = -(long long)y * p.width;
        xhi = (long long)(p.height - y) * p.width - 1;
    }
}
__device__ __forceinline__ long long floor_div(long long a, long long b) {  // b > 0
    const long long q = a / b;
    return (a % b < 0) ? q - 1 : q;
}
// First row of the bounding box [x0, x1] x [y0, ..] that can keep a pixel: the (record, scanline) items of the
// walk are counted from it.  On-screen boxes (0 <= x1 < width, the common case) start at max(y0, 0).
__device__ __forceinline__ int row_base(const DrawParams &p, int x1, int y0) {
    int ylo = 0;
    if (!p.state.x_guard && (unsigned)x1 >= (unsigned)p.width)  // smallest y with -y*W <= x1
        ylo = (int)-floor_div((long long)x1, (long long)p.width);
    return max(y0, ylo);
}

// Per-triangle constants of the row walk (context.go:167-181).
struct EdgeSetup {
    double a01, b01, a12, b12, a20, b20;
    double ra, ra12, ra20, ra01;
    int32_t x0, x1;
};
template <class R>
__device__ __forceinline__ EdgeSetup edge_setup(const R &r) {
    EdgeSetup e;
    e.a01 = r.s1y - r.s0y; e.b01 = r.s0x - r.s1x;  // context.go:167-172
    e.a12 = r.s2y - r.s1y; e.b12 = r.s1x - r.s2x;
    e.a20 = r.s0y - r.s2y; e.b20 = r.s2x - r.s0x;
    e.ra = r.ra; e.ra12 = r.ra12; e.ra20 = r.ra20; e.ra01 = r.ra01;
    e.x0 = r.x0; e.x1 = r.x1;
    return e;
}

// One row, given the edge values (w00, w01, w02) at its first pixel centre (x0 + .5, y + .5):
// context.go:185-221.  emit(w0, w1, w2, x, cnt, strip) is called for every segment, left to right.
// Returns the number of segments; adds the covered pixels to *covered.
template <class F>
__device__ __forceinline__ uint32_t walk_row_core(const DrawParams &p, const EdgeSetup &e, int y, double w00, double w01,
                                                  double w02, F &&emit, unsigned long long *covered) {
    // skip-ahead, context.go:185-205
    double d = 0;
    const double d0 = -w00 * e.ra12, d1 = -w01 * e.ra20, d2 = -w02 * e.ra01;
    if (w00 < 0 && d0 > d) d = d0;
    if (w01 < 0 && d1 > d) d = d1;
    if (w02 < 0 && d2 > d) d = d2;
    d = (double)go_int(d);
    if (d < 0) d = 0;
    double w0 = w00 + e.a12 * d, w1 = w01 + e.a20 * d, w2 = w02 + e.a01 * d;
    long long x = (long long)e.x0 + go_int(d);
    long long xlo, xhi;
    row_x_range(p, y, xlo, xhi);
    const long long xend = min((long long)e.x1, xhi);
    if (x > xend) return 0;
    // Pixels the reference tests but cannot keep (i < 0, or x outside the framebuffer under the x-guard rule): the
    // chain of adds runs over them, and so does wasInside (context.go:216-219 come before the index test).
    bool was_inside = false;
    for (; x < xlo; x++) {
        const double b0 = w0 * e.ra, b1 = w1 * e.ra, b2 = w2 * e.ra;
        if (b0 < 0 || b1 < 0 || b2 < 0) { if (was_inside) return 0; }
        else was_inside = true;
        w0 += e.a12; w1 += e.a20; w2 += e.a01;
    }
    // where pixel x of this row lands: (tx, ty) = (i mod W, i div W), advanced incrementally
    const long long q = floor_div(x, (long long)p.width);
    int tx = (int)(x - q * p.width), ty = y + (int)q;
    uint32_t nseg = 0, cnt = 0, skey = 0, swrap = 0;
    int sx = 0;
    double sw0 = 0, sw1 = 0, sw2 = 0;
    // Exact early exit.  fl(w + a) is monotone in w, so an edge value whose per-pixel increment is <= 0 never
    // grows along the row, and neither does b = fl(w * ra) for ra > 0: once such a b is negative, every later
    // pixel of the row is outside as well -- the reference would test them all and find nothing.
    const bool pos = e.ra > 0;
    const bool n12 = pos && e.a12 <= 0, n20 = pos && e.a20 <= 0, n01 = pos && e.a01 <= 0;
    for (; x <= xend; x++) {
        const double b0 = w0 * e.ra, b1 = w1 * e.ra, b2 = w2 * e.ra;  // context.go:208-210
        if (b0 < 0 || b1 < 0 || b2 < 0) {
            if (was_inside) break;  // context.go:216-218
            if ((n12 && b0 < 0) || (n20 && b1 < 0) || (n01 && b2 < 0)) break;
        } else {
            was_inside = true;
            const uint32_t key = (uint32_t)ty * (uint32_t)p.tiles_x + (uint32_t)(tx >> p.tile_shift);
            if (cnt == 0 || key != skey) {  // a new strip (also: the next row, after a wrap)
                if (cnt > 0) { emit(sw0, sw1, sw2, sx, cnt, skey, swrap); *covered += cnt; nseg++; }
                skey = key; sx = tx; sw0 = w0; sw1 = w1; sw2 = w2; cnt = 0;
                swrap = ty != y ? REC_WRAP : 0u;
            }
            cnt++;
        }
        w0 += e.a12; w1 += e.a20; w2 += e.a01;  // context.go:211-213
        if (++tx == p.width) { tx = 0; ty++; }
    }
    if (cnt > 0) { emit(sw0, sw1, sw2, sx, cnt, skey, swrap); *covered += cnt; nseg++; }
    return nseg;
}

// b0 < 0, b1 < 0, b2 < 0 (context.go:215) as a bit mask / as one predicate.  Written in PTX so that the three
// IEEE compares stay three DSETP: the optimiser otherwise rewrites `a < 0 || b < 0` into a NaN-aware minimum
// (DSETP.MIN + selects + moves, 30-40 instructions per pixel of the loops below).
__device__ __forceinline__ uint32_t neg_mask3(double b0, double b1, double b2) {
    uint32_t m;
    asm("{\n\t.reg .pred p0, p1, p2;\n\t.reg .u32 t1, t2;\n\t"
        "setp.lt.f64 p0, %1, 0d0000000000000000;\n\t"
        "setp.lt.f64 p1, %2, 0d0000000000000000;\n\t"
        "setp.lt.f64 p2, %3, 0d0000000000000000;\n\t"
        "selp.u32 %0, 1, 0, p0;\n\t"
        "selp.u32 t1, 2, 0, p1;\n\t"
        "selp.u32 t2, 4, 0, p2;\n\t"
        "or.b32 %0, %0, t1;\n\t"
        "or.b32 %0, %0, t2;\n\t}"
        : "=r"(m) : "d"(b0), "d"(b1), "d"(b2));
    return m;
}
__device__ __forceinline__ bool any_neg3(double b0, double b1, double b2) {
    uint32_t m;
    asm("{\n\t.reg .pred p;\n\t"
        "setp.lt.f64 p, %1, 0d0000000000000000;\n\t"
        "setp.lt.or.f64 p, %2, 0d0000000000000000, p;\n\t"
        "setp.lt.or.f64 p, %3, 0d0000000000000000, p;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(m) : "d"(b0), "d"(b1), "d"(b2));
    return m != 0;
}

// The general walker as an out-of-line call that only reports the first segment: rows of boxes that leave the
// framebuffer sideways (they may alias into neighbouring rows).  Replays the per-row adds from y0 itself.
static __device__ __noinline__ uint32_t walk_row_general(const DrawParams &p, const EdgeSetup e, int y0, int y, double w00,
                                                         double w01, double w02, ParkedSeg *first,
                                                         unsigned long long *covered) {
    for (int yy = y0; yy < y; yy++) { w00 += e.b12; w01 += e.b20; w02 += e.b01; }  // context.go:275-277
    ParkedSeg f;
    f.w0 = f.w1 = f.w2 = 0; f.x = 0; f.cnt = 0; f.key = 0; f.wrap = 0;
    uint32_t k = 0;
    const uint32_t n = walk_row_core(p, e, y, w00, w01, w02, [&](double sw0, double sw1, double sw2, int sx, uint32_t cnt, uint32_t key, uint32_t wrap) {
        if (k == 0) { f.w0 = sw0; f.w1 = sw1; f.w2 = sw2; f.x = sx; f.cnt = cnt; f.key = key; f.wrap = wrap; }
        k++;
    }, covered);
    *first = f;
    return n;
}

// Count pass of the fused front end: the covered run of row y, found with two lean loops (skip the pixels left of
// the run, then count the run) instead of the per-pixel strip bookkeeping of walk_row_core -- the same chain of adds
// and the same tests, so the run, and the edge values at its first pixel, are those of walk_row_core bit for bit;
// the strip cuts follow from the run's end points.  Returns the number of segments, the first one in `first`.
template <class R>
__device__ __forceinline__ uint32_t walk_row_count(const DrawParams &p, const R &r, int y, ParkedSeg &first,
                                                   unsigned long long *covered) {
    // A box that leaves the framebuffer sideways: its rows may start left of the pixels the reference keeps, or
    // alias into neighbouring rows (row_x_range) -- the general walker handles both, out of line (rare: fat lines
    // at the screen border), so that its state stays off this path's registers.  Every other box lies in
    // 0 <= x0 <= x1 < width, its walked rows in [0, height): all its pixels are kept where they are.
    if ((unsigned)r.x0 >= (unsigned)p.width || (unsigned)r.x1 >= (unsigned)p.width) {
        ParkedSeg tmp;
        unsigned long long cov = 0;
        const uint32_t ns = walk_row_general(p, edge_setup(r), r.y0, y, r.w00, r.w01, r.w02, &tmp, &cov);
        first = tmp;
        *covered += cov;
        return ns;
    }
    const double a01 = r.s1y - r.s0y, b01 = r.s0x - r.s1x;  // context.go:167-172
    const double a12 = r.s2y - r.s1y, b12 = r.s1x - r.s2x;
    const double a20 = r.s0y - r.s2y, b20 = r.s2x - r.s0x;
    double w00 = r.w00, w01 = r.w01, w02 = r.w02;
    for (int yy = r.y0; yy < y; yy++) { w00 += b12; w01 += b20; w02 += b01; }  // context.go:275-277
    // skip-ahead, context.go:185-205
    double d = 0;
    const double d0 = -w00 * r.ra12, d1 = -w01 * r.ra20, d2 = -w02 * r.ra01;
    if (w00 < 0 && d0 > d) d = d0;
    if (w01 < 0 && d1 > d) d = d1;
    if (w02 < 0 && d2 > d) d = d2;
    const long long di = go_int(d);
    d = (double)di;
    if (d < 0) d = 0;
    double w0 = w00 + a12 * d, w1 = w01 + a20 * d, w2 = w02 + a01 * d;
    // x0 + int(d) of the clamped d (context.go:207); beyond 2^40 the row starts right of any bounding box
    const long long xl = (long long)r.x0 + (di < 0 ? 0ll : (di > (1ll << 40) ? (1ll << 40) : di));
    const long long xe64 = (long long)r.x1;
    if (xl > xe64) return 0;
    const int xe = (int)xe64;
    int x = (int)xl;
    const double ra = r.ra;
    // Exact early exit, see walk_row_core: an edge whose per-pixel increment is <= 0 never recovers (ra > 0).
    const bool pos = ra > 0;
    const bool n12 = pos && a12 <= 0, n20 = pos && a20 <= 0, n01 = pos && a01 <= 0;
    const uint32_t dead = (n12 ? 1u : 0u) | (n20 ? 2u : 0u) | (n01 ? 4u : 0u);
    for (;;) {  // pixels left of the run (context.go:208-219 with wasInside == false)
        const uint32_t out = neg_mask3(w0 * ra, w1 * ra, w2 * ra);
        if (out == 0) break;
        if (out & dead) return 0;
        w0 += a12; w1 += a20; w2 += a01;  // context.go:211-213
        if (++x > xe) return 0;
    }
    first.w0 = w0; first.w1 = w1; first.w2 = w2; first.x = x; first.wrap = 0;
    const int xs = x;
    do {  // the run: up to the first outside pixel (context.go:216-218) or the end of the row
        w0 += a12; w1 += a20; w2 += a01;
        if (++x > xe) break;
    } while (!any_neg3(w0 * ra, w1 * ra, w2 * ra));
    const uint32_t n = (uint32_t)(x - xs);
    *covered += n;
    const int c0 = xs >> p.tile_shift, c1 = (x - 1) >> p.tile_shift;
    first.cnt = min(n, (uint32_t)(((c0 + 1) << p.tile_shift) - xs));
    first.key = (uint32_t)y * (uint32_t)p.tiles_x + (uint32_t)c0;
    return (uint32_t)(c1 - c0 + 1);
}

// The count walk in two halves, for the compacting front end (fgl_geom.cu, FGL_FRONT_COMPACT): walk_row_find runs the
// chain up to the first covered pixel (row replay, skip-ahead, the pixels left of the run) and says whether there is
// one; walk_row_run continues from there over the run.  Together they execute exactly the adds and tests of
// walk_row_count.  A box that leaves the framebuffer sideways is reported as found with x = WALK_GENERAL: the second
// half then calls the general walker for the whole row.
constexpr int WALK_GENERAL = (int)0x80000000;
template <class R>
__device__ __forceinline__ bool walk_row_find(const DrawParams &p, const R &r, int y, double &w0, double &w1, double &w2, int &x) {
    if ((unsigned)r.x0 >= (unsigned)p.width || (unsigned)r.x1 >= (unsigned)p.width) { x = WALK_GENERAL; w0 = w1 = w2 = 0; return true; }
    const double a01 = r.s1y - r.s0y, b01 = r.s0x - r.s1x;  // context.go:167-172
    const double a12 = r.s2y - r.s1y, b12 = r.s1x - r.s2x;
    const double a20 = r.s0y - r.s2y, b20 = r.s2x - r.s0x;
    double w00 = r.w00, w01 = r.w01, w02 = r.w02;
    for (int yy = r.y0; yy < y; yy++) { w00 += b12; w01 += b20; w02 += b01; }  // context.go:275-277
    double d = 0;  // skip-ahead, context.go:185-205
    const double d0 = -w00 * r.ra12, d1 = -w01 * r.ra20, d2 = -w02 * r.ra01;
    if (w00 < 0 && d0 > d) d = d0;
    if (w01 < 0 && d1 > d) d = d1;
    if (w02 < 0 && d2 > d) d = d2;
    const long long di = go_int(d);
    d = (double)di;
    if (d < 0) d = 0;
    w0 = w00 + a12 * d; w1 = w01 + a20 * d; w2 = w02 + a01 * d;
    const long long xl = (long long)r.x0 + (di < 0 ? 0ll : (di > (1ll << 40) ? (1ll << 40) : di));
    if (xl > (long long)r.x1) return false;
    const int xe = r.x1;
    x = (int)xl;
    const double ra = r.ra;
    const bool pos = ra > 0;  // exact early exit, see walk_row_core
    const uint32_t dead = ((pos && a12 <= 0) ? 1u : 0u) | ((pos && a20 <= 0) ? 2u : 0u) | ((pos && a01 <= 0) ? 4u : 0u);
    for (;;) {  // pixels left of the run (context.go:208-219 with wasInside == false)
        const uint32_t out = neg_mask3(w0 * ra, w1 * ra, w2 * ra);
        if (out == 0) return true;
        if (out & dead) return false;
        w0 += a12; w1 += a20; w2 += a01;  // context.go:211-213
        if (++x > xe) return false;
    }
}
// (w0, w1, w2, x): the first covered pixel found by walk_row_find.  Returns the segments of the row, the first in `first`.
template <class R>
__device__ __forceinline__ uint32_t walk_row_run(const DrawParams &p, const R &r, int y, double w0, double w1, double w2, int x,
                                                 ParkedSeg &first, unsigned long long *covered) {
    if (x == WALK_GENERAL) {
        ParkedSeg tmp;
        unsigned long long cov = 0;
        const uint32_t ns = walk_row_general(p, edge_setup(r), r.y0, y, r.w00, r.w01, r.w02, &tmp, &cov);
        first = tmp;
        *covered += cov;
        return ns;
    }
    const double a01 = r.s1y - r.s0y, a12 = r.s2y - r.s1y, a20 = r.s0y - r.s2y, ra = r.ra;
    const int xe = r.x1;
    first.w0 = w0; first.w1 = w1; first.w2 = w2; first.x = x; first.wrap = 0;
    const int xs = x;
    do {  // the run: up to the first outside pixel (context.go:216-218) or the end of the row
        w0 += a12; w1 += a20; w2 += a01;
        if (++x > xe) break;
    } while (!any_neg3(w0 * ra, w1 * ra, w2 * ra));
    const uint32_t n = (uint32_t)(x - xs);
    *covered += n;
    const int c0 = xs >> p.tile_shift, c1 = (x - 1) >> p.tile_shift;
    first.cnt = min(n, (uint32_t)(((c0 + 1) << p.tile_shift) - xs));
    first.key = (uint32_t)y * (uint32_t)p.tiles_x + (uint32_t)c0;
    return (uint32_t)(c1 - c0 + 1);
}

// Walk row y of the triangle described by r (any struct with the fields s0x..s2y, w00..w02, ra, ra12, ra20,
// ra01, z0..z2, x0, x1, y0): the per-row adds are replayed from y0 (context.go:275-277).
//   WRITE == false: the first segment is returned in `first`, nothing is stored.
//   WRITE == true : every segment k is stored as segv[base + k] / keys[base + k] (if base + k < cap).
template <bool WRITE, class R>
__device__ __forceinline__ uint32_t walk_row_segments(const DrawParams &p, const R &r, int y, ParkedSeg &first,
                                                      const RecTail *tail, SegV *__restrict__ segv,
                                                      uint32_t *__restrict__ keys, uint32_t base, uint32_t cap,
                                                      unsigned long long *covered) {
    const EdgeSetup e = edge_setup(r);
    double w00 = r.w00, w01 = r.w01, w02 = r.w02;
    for (int yy = r.y0; yy < y; yy++) { w00 += e.b12; w01 += e.b20; w02 += e.b01; }  // context.go:275-277
    uint32_t k = 0;
    return walk_row_core(p, e, y, w00, w01, w02, [&](double sw0, double sw1, double sw2, int sx, uint32_t cnt, uint32_t key, uint32_t wrap) {
        if (WRITE) {
            const uint32_t slot = base + k;
            if (slot < cap) {
                segv[slot] = make_segv(sw0, sw1, sw2, r.ra, r.z0, r.z1, r.z2, e.a12, e.a20, e.a01, *tail, (uint16_t)sx, (uint8_t)cnt, wrap);
                keys[slot] = key;
            }
        } else if (k == 0) {
            first.w0 = sw0; first.w1 = sw1; first.w2 = sw2; first.x = sx; first.cnt = cnt; first.key = key; first.wrap = wrap;
        }
        k++;
    }, covered);
}

}  // namespace fgl
