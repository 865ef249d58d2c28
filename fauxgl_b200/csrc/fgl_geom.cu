// fgl_geom.cu -- primitive-parallel front end: mesh ingest (AoS -> planar SoA),
// Mesh.Transform, and the geometry stage of a draw (Shader.Vertex, outcodes,
// ClipTriangle/ClipLine, NDC divide, cull, screen transform, wireframe/fat-line
// expansion, integer bounding box) with order-preserving compaction.
//
// Replaces Context.DrawTriangle / DrawLine / drawClippedTriangle /
// drawClippedLine / line / wireframe (context.go:283-389) and clipping.go.
// One thread per input primitive; coalesced f64 loads from the position planes.
// Compaction keeps primitive order: a count pass, an exclusive
// scan, and an emit pass that re-runs the same arithmetic (deterministic, so
// both passes agree) and writes each record at its ordered slot.
#include "fgl_internal.h"
#include "fgl_block.cuh"
#include "fgl_math.cuh"
#include "fgl_walk.cuh"

namespace fgl {

// ---- mesh ingest / export / transform ---------------------------------------------------

// aos: [n][nverts][ncomp_in] -> planes[(v*ncomp_out + c)*n + i], c < ncomp_out <= ncomp_in.
__global__ void k_mesh_ingest(const double *__restrict__ aos, double *__restrict__ planes, uint32_t n, int nverts,
                              int ncomp_in, int ncomp_out) {
    const size_t total = (size_t)n * nverts * ncomp_out;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(o % n);
        const uint32_t vc = (uint32_t)(o / n);
        const uint32_t v = vc / ncomp_out, c = vc % ncomp_out;
        planes[o] = aos[((size_t)i * nverts + v) * ncomp_in + c];
    }
}
__global__ void k_mesh_export(const double *__restrict__ planes, double *__restrict__ aos, uint32_t n, int nverts,
                              int ncomp) {
    const size_t total = (size_t)n * nverts * ncomp;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(o % n);
        const uint32_t vc = (uint32_t)(o / n);
        const uint32_t v = vc / ncomp, c = vc % ncomp;
        aos[((size_t)i * nverts + v) * ncomp + c] = planes[o];
    }
}
struct Mat16 { double m[16]; };
// Mesh.Transform: mesh.go:167-175, triangle.go:66-73, line.go:23-28.
__global__ void k_mesh_transform(double *__restrict__ pos, double *__restrict__ nrm, uint32_t n, int nverts,
                                 const Mat16 M) {
    const size_t total = (size_t)n * nverts;
    for (size_t o = blockIdx.x * (size_t)blockDim.x + threadIdx.x; o < total; o += (size_t)gridDim.x * blockDim.x) {
        const uint32_t i = (uint32_t)(o % n), v = (uint32_t)(o / n);
        double *px = pos + (size_t)(v * 3 + 0) * n + i, *py = pos + (size_t)(v * 3 + 1) * n + i,
               *pz = pos + (size_t)(v * 3 + 2) * n + i;
        V3 p = m_mul_position(M.m, v3(*px, *py, *pz));
        *px = p.x; *py = p.y; *pz = p.z;
        double *nx = nrm + (size_t)(v * 3 + 0) * n + i, *ny = nrm + (size_t)(v * 3 + 1) * n + i,
               *nz = nrm + (size_t)(v * 3 + 2) * n + i;
        V3 d = m_mul_direction(M.m, v3(*nx, *ny, *nz));
        *nx = d.x; *ny = d.y; *nz = d.z;
    }
}

static int grid_for(size_t total, int threads) {
    size_t b = (total + threads - 1) / threads;
    size_t cap = GRID_WAVE * 2;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
int launch_mesh_ingest(const double *aos, double *planes, uint32_t n, int nverts, int ncomp_in, int ncomp_out,
                       cudaStream_t st) {
    if (n == 0) return 0;
    k_mesh_ingest<<<grid_for((size_t)n * nverts * ncomp_out, 256), 256, 0, st>>>(aos, planes, n, nverts, ncomp_in,
                                                                                ncomp_out);
    return 1;
}
int launch_mesh_export(const double *planes, double *aos, uint32_t n, int nverts, int ncomp, cudaStream_t st) {
    if (n == 0) return 0;
    k_mesh_export<<<grid_for((size_t)n * nverts * ncomp, 256), 256, 0, st>>>(planes, aos, n, nverts, ncomp);
    return 1;
}
int launch_mesh_transform(double *pos, double *nrm, uint32_t n, int nverts, const double m[16], cudaStream_t st) {
    if (n == 0) return 0;
    Mat16 M;
    for (int i = 0; i < 16; i++) M.m[i] = m[i];
    k_mesh_transform<<<grid_for((size_t)n * nverts, 256), 256, 0, st>>>(pos, nrm, n, nverts, M);
    return 1;
}

// ---- geometry stage -----------------------------------------------------------------------

struct FullVertex {  // vertex.go:3-12 without Texture.Z
    V3 pos, nrm;
    double tu, tv;
    C4 col;
    V4 out;
};

FGL_DI int32_t sat_i32(long long v) {
    const long long L = 1LL << 30;
    return (int32_t)(v < -L ? -L : (v > L ? L : v));
}

// Does the box stay inside the framebuffer sideways?  Then none of its pixels aliases into another row under the
// reference's index rule, and the x-guard rule and the reference's agree on it.
FGL_DI bool box_on_screen_x(const DrawParams &p, int32_t x0, int32_t x1) {
    return p.state.x_guard || ((unsigned)x0 < (unsigned)p.width && (unsigned)x1 < (unsigned)p.width);
}
// Upper bound of the strips one scanline of the box [x0, x1] can touch (the segment regions are sized with it).
FGL_DI uint32_t box_cols(const DrawParams &p, int32_t x0, int32_t x1) {
    if (box_on_screen_x(p, x0, x1))
        return (uint32_t)((min(x1, p.width - 1) >> p.tile_shift) - (max(x0, 0) >> p.tile_shift) + 1);
    // its span in the linear index space, plus one strip per row boundary it crosses
    const long long span = (long long)x1 - x0;
    const long long cols = (span >> p.tile_shift) + 3 + span / p.width;
    return (uint32_t)min(cols, (long long)p.tiles_x * ((long long)p.height + 2));
}

// Integer bounding box of a screen triangle, context.go:155-160, and its on-screen scanlines.
struct BBox {
    int32_t x0, x1, y0, y1; bool visible; uint32_t rows, cols; bool origin;  // cols: strips a row can touch
    double mnx, mny, mxx, mxy, fy0, fx1, fy1;  // the extremes and their floor / ceil as float64 (tighten_box)
};
FGL_DI BBox compute_bbox(const DrawParams &p, V3 s0, V3 s1, V3 s2) {
    BBox b;
    const double mnx = go_min(s0.x, go_min(s1.x, s2.x)), mny = go_min(s0.y, go_min(s1.y, s2.y));
    const double mxx = go_max(s0.x, go_max(s1.x, s2.x)), mxy = go_max(s0.y, go_max(s1.y, s2.y));
    b.mnx = mnx; b.mny = mny; b.mxx = mxx; b.mxy = mxy;
    b.fy0 = floor(mny); b.fx1 = ceil(mxx); b.fy1 = ceil(mxy);
    const long long ix0 = go_int(floor(mnx)), ix1 = go_int(b.fx1), iy0 = go_int(b.fy0), iy1 = go_int(b.fy1);
    b.x0 = sat_i32(ix0);
    b.x1 = sat_i32(ix1);
    b.y0 = sat_i32(iy0);
    b.y1 = sat_i32(iy1);
    b.origin = false;
    {
        // All four bounds the "integer indefinite" value int(NaN) = -2^63: every screen coordinate is NaN -- what a
        // degenerate triangle (two equal vertices) becomes in ClipTriangle, whose Barycentric divides 0 by 0; the cull
        // test `a <= 0` is false for NaN, so it is drawn under every Cull mode.  The reference's loops then visit
        // exactly one "pixel", (x, y) = (-2^63, -2^63), and Go's wrapping arithmetic gives it the index
        // -2^63 * W - 2^63 = 0 when W is odd (-2^63, skipped, when W is even): pixel (0, 0) is counted in TotalPixels
        // and, with ReadDepth off, gets a NaN depth.  Reproduced as a one-pixel box at the origin whose edge values
        // are set up at (-2^63 + .5, -2^63 + .5) and whose segment is marked as aliased (its own x is off screen).
        const long long IND = (long long)0x8000000000000000ull;
        if (ix0 == IND && ix1 == IND && iy0 == IND && iy1 == IND) {
            b.x0 = b.x1 = b.y0 = b.y1 = 0;
            b.visible = (p.width & 1) && !p.state.x_guard;
            b.origin = b.visible;
            b.rows = b.visible ? 1u : 0u;
            b.cols = b.rows;
            return b;
        }
    }
    // The forward-differencing chains are replayed from (x0, y0); a box that starts
    // millions of pixels off screen (infinite/overflowing coordinates -- the reference
    // itself would spin for 2^63 iterations on those) is dropped instead of walked.
    constexpr int32_t FAR = 1 << 22;
    const bool absurd = b.x0 < -FAR || b.y0 < -FAR || b.x1 > FAR || b.y1 > FAR;
    if (box_on_screen_x(p, b.x0, b.x1)) {
        // only the on-screen part of the box can keep pixels
        const int32_t cx0 = max(b.x0, 0), cx1 = min(b.x1, p.width - 1);
        const int32_t cy0 = max(b.y0, 0), cy1 = min(b.y1, p.height - 1);
        b.visible = !absurd && cx0 <= cx1 && cy0 <= cy1;
        b.rows = b.visible ? (uint32_t)(cy1 - cy0 + 1) : 0u;
    } else {
        // The reference's index rule (context.go:223-228, fgl_walk.cuh row_x_range): row y keeps the pixels with
        // -y*W <= x <= (H-y)*W - 1.  Rows of the box for which that range meets [x0, x1]:
        const long long W = p.width, H = p.height;
        const int cy0 = row_base(p, b.x1, b.y0);
        const long long yhi = floor_div(H * W - 1 - (long long)b.x0, W);
        const int cy1 = (int)min((long long)b.y1, yhi);
        b.visible = !absurd && cy0 <= cy1 && b.x0 <= b.x1;
        b.rows = b.visible ? (uint32_t)(cy1 - cy0 + 1) : 0u;
    }
    b.cols = b.visible ? box_cols(p, b.x0, b.x1) : 0u;
    return b;
}

FGL_DI double edge_fn(V3 a, V3 b, V3 c) {  // context.go:147-149
    return (b.x - c.x) * (a.y - c.y) - (b.y - c.y) * (a.x - c.x);
}

// Rows and columns of an on-screen box that provably keep no pixel (fused front end, fast path).  The reference
// walks every row y of [floor(min y), ceil(max y)] and every column of [floor(min x), ceil(max x)]
// (context.go:155-160, 184-221), although a pixel centre above, below or right of the triangle is never covered: the
// last row and the last column never are, the first row half of the time -- for the one-to-two-pixel triangles of the
// benchmark mesh 43 % of the (record, scanline) items.  Skipping them changes nothing only if the reference's
// FLOATING-POINT test fails there as well, so a row is dropped when that is certain:
//   exact arithmetic: T the top vertex, d_a, d_b the edges from T, q = p - T with q.y = delta > 0.  The edge
//   functions u_a = s cross(d_a, q), u_b = s cross(q, d_b) of those edges (s = the sign that makes them positive
//   inside) satisfy |d_b.y| u_a + |d_a.y| u_b = -|cross(d_a, d_b)| delta = -|A| delta, A = the area term
//   edge_fn(s0, s1, s2) (twice the triangle's area), hence
//   min(u_a, u_b) <= -|A| delta / (|d_a.y| + |d_b.y|) <= -|A| delta / (2 H),  H = max y - min y:
//   SOME edge function is that negative at EVERY pixel of the row (mirrored: bottom rows, right columns).
//   rounding: the reference's values come from edge_fn at (x0 + .5, y0 + .5) and a chain of at most rows + cols + 2
//   additions of values bounded by 2 B^2, B = the box size + 2: they differ from the exact ones by less than
//   2^-50 B^2 (B + 8).  E below is 64 times that.
// A row is dropped when delta >= 4 E H / |A| + 1e-6 (the constant covers the rounding of the comparison itself for
// coordinates below 2^22), i.e. when that edge value is below -2 E, twice the rounding bound, itself 64-fold;
// sign(w * ra) = sign(w) sign(ra) then holds without underflow (|w| > E, |ra| >= 2^-48).  Slivers (|A| <= 16 E) and anything non-finite are left alone.
// Written without branches or float <-> integer conversions (the floor / ceil values of compute_bbox are reused), so
// that it schedules between the seven divisions of the set-up.  Returns the first row to walk; may clear b.visible.
FGL_DI int tighten_box(const DrawParams &p, BBox &b, double ra) {
    if (!b.visible || b.origin || (unsigned)b.x0 >= (unsigned)p.width || (unsigned)b.x1 >= (unsigned)p.width) return b.y0;
    const double hx = b.mxx - b.mnx, hy = b.mxy - b.mny;
    const double B = fmax(hx, hy) + 4.0;   // >= max(x1 - x0, y1 - y0) + 2
    const double E = 0x1p-44 * B * B * (B + 8.0);
    const double k = 4.0 * E * fabs(ra);   // 4 E / |A|
    const double my = k * hy + 1e-6, mx = k * hx + 1e-6;
    const bool sure = k < 0.25 && my < 0.25 && mx < 0.25;  // (false for NaN)
    // first row y0 = floor(min y): dropped when y0 + .5 <= min y - my
    const int drop_first = (sure && b.mny - b.fy0 >= 0.5 + my) ? 1 : 0;
    // last row y1 = ceil(max y): y1 + .5 >= max y + my always (my < .25); the row before it when y1 - .5 >= max y + my
    const int drop_last = sure ? ((b.fy1 - b.mxy >= 0.5 + my) ? 2 : 1) : 0;
    const int drop_right = sure ? ((b.fx1 - b.mxx >= 0.5 + mx) ? 2 : 1) : 0;
    const int cy0 = max(max(b.y0, 0), b.y0 + drop_first), cy1 = min(min(b.y1, p.height - 1), b.y1 - drop_last);
    const int cx1 = b.x1 - drop_right;  // (b.x1 < width)
    const bool keep = cy0 <= cy1 && cx1 >= b.x0;
    b.visible = keep;
    b.x1 = keep ? cx1 : b.x1;
    b.rows = keep ? (uint32_t)(cy1 - cy0 + 1) : 0u;
    b.cols = keep ? (uint32_t)((cx1 >> p.tile_shift) - (b.x0 >> p.tile_shift) + 1) : 0u;
    return cy0;
}

// Write one record at slot r; its scanlines start at row_off within the block's share of the
// (record, scanline) item space of the span stage.
FGL_DI void write_record(const WorkBuffers *wb, uint32_t r, uint32_t row_off, const BBox &b, V3 s0, V3 s1, V3 s2,
                         double w0, double w1, double w2, uint32_t src, uint32_t flags) {
    if (r >= wb->cap_records) return;
    Rec rec;
    rec.s[0] = s0.x; rec.s[1] = s0.y; rec.s[2] = s0.z;
    rec.s[3] = s1.x; rec.s[4] = s1.y; rec.s[5] = s1.z;
    rec.s[6] = s2.x; rec.s[7] = s2.y; rec.s[8] = s2.z;
    // per-triangle setup of Context.rasterize, context.go:163-181
    const V3 pc = b.origin ? v3(-9223372036854775808.0, -9223372036854775808.0, 0) : v3((double)b.x0 + 0.5, (double)b.y0 + 0.5, 0);
    rec.w00 = edge_fn(s1, s2, pc);
    rec.w01 = edge_fn(s2, s0, pc);
    rec.w02 = edge_fn(s0, s1, pc);
    const double a01 = s1.y - s0.y, a12 = s2.y - s1.y, a20 = s0.y - s2.y;
    rec.ra = 1 / edge_fn(s0, s1, s2);
    rec.r0 = 1 / w0; rec.r1 = 1 / w1; rec.r2 = 1 / w2;
    rec.ra12 = 1 / a12; rec.ra20 = 1 / a20; rec.ra01 = 1 / a01;
    rec.src = src; rec.flags = flags | (b.origin ? REC_WRAP : 0u);
    rec.x0 = b.x0; rec.x1 = b.x1; rec.y0 = b.y0; rec.y1 = b.y1;
    wb->recs[r] = rec;
    wb->rec_local_row[r] = row_off;
}

// First pass over a primitive: counts its records and scanlines and keeps the first record in
// registers (the common case is exactly one), so that it can be written without recomputation
// once the ordered offset is known.
struct CountEmit {
    uint32_t n, rows;
    unsigned long long cells;  // sum of rows x strip columns: an upper bound of the segments
    static constexpr bool kWrite = false;
    FGL_DI void record(const DrawParams &p, V3 a0, V3 a1, V3 a2, double, double, double, uint32_t, uint32_t) {
        const BBox bb = compute_bbox(p, a0, a1, a2);
        if (!bb.visible) return;
        n++;
        rows += bb.rows;
        cells += (unsigned long long)bb.rows * bb.cols;
    }
    FGL_DI uint32_t pool_alloc(const FullVertex *, uint32_t) { return 0; }
};
struct WriteEmit {
    uint32_t next, row_next;
    const WorkBuffers *wb;
    static constexpr bool kWrite = true;
    FGL_DI bool wants(uint32_t) const { return true; }
    FGL_DI void skip(uint32_t) {}
    FGL_DI void record(const DrawParams &p, V3 s0, V3 s1, V3 s2, double w0, double w1, double w2, uint32_t src,
                       uint32_t flags) {
        const BBox b = compute_bbox(p, s0, s1, s2);
        if (!b.visible) return;
        write_record(wb, next, row_next, b, s0, s1, s2, w0, w1, w2, src, flags);
        next++;
        row_next += b.rows;
    }
    FGL_DI uint32_t pool_alloc(const FullVertex *v, uint32_t prim) {
        const uint32_t slot = atomicAdd(&wb->counters->n_clip, 1u);
        if (slot >= wb->cap_clip) { atomicOr(&wb->counters->overflow, OVF_CLIP); return 0; }
        ClipTri t;
#pragma unroll
        for (int k = 0; k < 3; k++) {
            t.v[k].pos[0] = v[k].pos.x; t.v[k].pos[1] = v[k].pos.y; t.v[k].pos[2] = v[k].pos.z;
            t.v[k].nrm[0] = v[k].nrm.x; t.v[k].nrm[1] = v[k].nrm.y; t.v[k].nrm[2] = v[k].nrm.z;
            t.v[k].tex[0] = v[k].tu; t.v[k].tex[1] = v[k].tv;
            t.v[k].col[0] = v[k].col.r; t.v[k].col[1] = v[k].col.g; t.v[k].col[2] = v[k].col.b; t.v[k].col[3] = v[k].col.a;
        }
        t.prim = prim; t._pad = 0;
        wb->clip_pool[slot] = t;
        return slot;
    }
};

FGL_DI uint32_t vmap3(uint32_t a, uint32_t b, uint32_t c) { return a | (b << 2) | (c << 4); }

// Context.line, context.go:283-294: fat line -> two triangles (v1,v0,v0; s11,s01,s00), (v1,v1,v0; s10,s11,s00).
template <class Emit>
FGL_DI void emit_line(const DrawParams &p, Emit &e, V3 s0, V3 s1, uint32_t i0, uint32_t i1, double w0, double w1,
                      uint32_t src, uint32_t srcflags) {
    const double half = p.state.line_width / 2;
    const V3 n = v_muls(v_perpendicular(v_sub(s1, s0)), half);
    s0 = v_add(s0, v_muls(v_normalize(v_sub(s0, s1)), half));
    s1 = v_add(s1, v_muls(v_normalize(v_sub(s1, s0)), half));
    const V3 s00 = v_add(s0, n), s01 = v_sub(s0, n), s10 = v_add(s1, n), s11 = v_sub(s1, n);
    e.record(p, s11, s01, s00, w1, w0, w0, src, srcflags | vmap3(i1, i0, i0));
    e.record(p, s10, s11, s00, w1, w1, w0, src, srcflags | vmap3(i1, i1, i0));
}

// drawClippedTriangle, context.go:316-349.  o[] = Output of the three vertices.
template <class Emit>
FGL_DI void emit_clipped_triangle(const DrawParams &p, Emit &e, const V4 o[3], uint32_t src, uint32_t srcflags) {
    V3 ndc0 = v3(o[0].x / o[0].w, o[0].y / o[0].w, o[0].z / o[0].w);
    V3 ndc1 = v3(o[1].x / o[1].w, o[1].y / o[1].w, o[1].z / o[1].w);
    V3 ndc2 = v3(o[2].x / o[2].w, o[2].y / o[2].w, o[2].z / o[2].w);
    double a = (ndc1.x - ndc0.x) * (ndc2.y - ndc0.y) - (ndc2.x - ndc0.x) * (ndc1.y - ndc0.y);
    uint32_t i0 = 0, i2 = 2;
    if (a < 0) {
        V3 t = ndc0; ndc0 = ndc2; ndc2 = t;
        i0 = 2; i2 = 0;
    }
    if (p.state.cull == FGL_CULL_FRONT) a = -a;
    if (p.state.front_face == FGL_FACE_CW) a = -a;
    if (p.state.cull != FGL_CULL_NONE && a <= 0) return;
    const V3 s0 = m_mul_position(p.screen, ndc0), s1 = m_mul_position(p.screen, ndc1),
             s2 = m_mul_position(p.screen, ndc2);
    const double w0 = o[i0].w, w1 = o[1].w, w2 = o[i2].w;
    if (p.state.wireframe) {  // context.go:296-301
        emit_line(p, e, s0, s1, i0, 1, w0, w1, src, srcflags);
        emit_line(p, e, s1, s2, 1, i2, w1, w2, src, srcflags);
        emit_line(p, e, s2, s0, i2, i0, w2, w0, src, srcflags);
    } else {
        e.record(p, s0, s1, s2, w0, w1, w2, src, srcflags | vmap3(i0, 1, i2));
    }
}

// ---- clipping.go ------------------------------------------------------------------------------
struct ClipPlane { double px, py, pz, pw, nx, ny, nz, nw; };
__constant__ ClipPlane c_clip_planes[6] = {  // clipping.go:3-10
    {1, 0, 0, 1, -1, 0, 0, 1}, {-1, 0, 0, 1, 1, 0, 0, 1}, {0, 1, 0, 1, 0, -1, 0, 1},
    {0, -1, 0, 1, 0, 1, 0, 1}, {0, 0, 1, 1, 0, 0, -1, 1}, {0, 0, -1, 1, 0, 0, 1, 1},
};
FGL_DI bool point_in_front(const ClipPlane &pl, V4 v) {  // clipping.go:16-18
    return w_dot(w_sub(v, v4(pl.px, pl.py, pl.pz, pl.pw)), v4(pl.nx, pl.ny, pl.nz, pl.nw)) > 0;
}
// Does sutherlandHodgman (clipping.go:28-52) return an empty polygon for this triangle?  Decided exactly as it would:
// the planes in its order, the same point_in_front arithmetic; as long as all three vertices are in front of a plane
// the polygon passes it unchanged, and the first plane that has all three behind it empties it.  A plane that cuts
// the triangle ends the test (false: the real clipper decides).  Lets the front ends drop geometry outside the view
// volume -- everything beyond the frame of a close-up -- without entering the general path.
FGL_DI bool clips_to_nothing(const V4 *o) {
#pragma unroll
    for (int pi = 0; pi < 6; pi++) {
        const ClipPlane pl = c_clip_planes[pi];
        const bool f0 = point_in_front(pl, o[0]), f1 = point_in_front(pl, o[1]), f2 = point_in_front(pl, o[2]);
        if (!f0 && !f1 && !f2) return true;
        if (!(f0 && f1 && f2)) return false;
    }
    return false;
}
FGL_DI V4 intersect_segment(const ClipPlane &pl, V4 v0, V4 v1) {  // clipping.go:20-26
    const V4 N = v4(pl.nx, pl.ny, pl.nz, pl.nw);
    const V4 u = w_sub(v1, v0);
    const V4 w = w_sub(v0, v4(pl.px, pl.py, pl.pz, pl.pw));
    const double d = w_dot(N, u);
    const double n = -w_dot(N, w);
    return w_add(v0, w_muls(u, n / d));
}
constexpr int MAX_POLY = 12;
// sutherlandHodgman, clipping.go:28-52.  pts holds the input and receives the output.
__device__ __noinline__ int sutherland_hodgman(V4 *pts, int n) {
    V4 tmp[MAX_POLY];
    V4 *in = tmp, *out = pts;
    for (int pi = 0; pi < 6; pi++) {
        const ClipPlane pl = c_clip_planes[pi];
        V4 *t = in; in = out; out = t;
        const int nin = n;
        n = 0;
        if (nin == 0) return 0;
        V4 s = in[nin - 1];
        for (int k = 0; k < nin; k++) {
            const V4 e = in[k];
            if (point_in_front(pl, e)) {
                if (!point_in_front(pl, s)) { if (n < MAX_POLY) out[n++] = intersect_segment(pl, s, e); }
                if (n < MAX_POLY) out[n++] = e;
            } else if (point_in_front(pl, s)) {
                if (n < MAX_POLY) out[n++] = intersect_segment(pl, s, e);
            }
            s = e;
        }
    }
    if (out != pts)
        for (int k = 0; k < n; k++) pts[k] = out[k];
    return n;
}
FGL_DI V4 barycentric(V3 p1, V3 p2, V3 p3, V3 p) {  // vertex.go:81-95
    const V3 v0 = v_sub(p2, p1), v1 = v_sub(p3, p1), v2 = v_sub(p, p1);
    const double d00 = v_dot(v0, v0), d01 = v_dot(v0, v1), d11 = v_dot(v1, v1);
    const double d20 = v_dot(v2, v0), d21 = v_dot(v2, v1);
    const double d = d00 * d11 - d01 * d01;
    const double v = (d11 * d20 - d01 * d21) / d;
    const double w = (d00 * d21 - d01 * d20) / d;
    const double u = 1 - v - w;
    return v4(u, v, w, 1);
}
FGL_DI double interp1(double a, double b, double c, V4 B) {  // vertex.go:49-79 per component
    double n = 0;
    n = n + a * B.x;
    n = n + b * B.y;
    n = n + c * B.z;
    return n * B.w;
}
__device__ __noinline__ FullVertex interpolate_vertexes(const FullVertex *t, V4 B) {  // vertex.go:18-47
    FullVertex v;
    v.pos = v3(interp1(t[0].pos.x, t[1].pos.x, t[2].pos.x, B), interp1(t[0].pos.y, t[1].pos.y, t[2].pos.y, B),
               interp1(t[0].pos.z, t[1].pos.z, t[2].pos.z, B));
    v.nrm = v_normalize(v3(interp1(t[0].nrm.x, t[1].nrm.x, t[2].nrm.x, B),
                           interp1(t[0].nrm.y, t[1].nrm.y, t[2].nrm.y, B),
                           interp1(t[0].nrm.z, t[1].nrm.z, t[2].nrm.z, B)));
    v.tu = interp1(t[0].tu, t[1].tu, t[2].tu, B);
    v.tv = interp1(t[0].tv, t[1].tv, t[2].tv, B);
    v.col = c4(interp1(t[0].col.r, t[1].col.r, t[2].col.r, B), interp1(t[0].col.g, t[1].col.g, t[2].col.g, B),
               interp1(t[0].col.b, t[1].col.b, t[2].col.b, B), interp1(t[0].col.a, t[1].col.a, t[2].col.a, B));
    v.out = v4(interp1(t[0].out.x, t[1].out.x, t[2].out.x, B), interp1(t[0].out.y, t[1].out.y, t[2].out.y, B),
               interp1(t[0].out.z, t[1].out.z, t[2].out.z, B), interp1(t[0].out.w, t[1].out.w, t[2].out.w, B));
    return v;
}
FGL_DI void fix_normals(FullVertex *t) {  // triangle.go:33-58
    const V3 e1 = v_sub(t[1].pos, t[0].pos), e2 = v_sub(t[2].pos, t[0].pos);
    const V3 n = v_normalize(v_cross(e1, e2));
#pragma unroll
    for (int k = 0; k < 3; k++)
        if (v_is_zero(t[k].nrm)) t[k].nrm = n;
}

FGL_DI double plane_at(const double *base, uint32_t n, uint32_t v, uint32_t ncomp, uint32_t c, uint32_t i) {
    return __ldg(base + (size_t)(v * ncomp + c) * n + i);
}

// DrawTriangle's clip branch, context.go:376-384 + ClipTriangle, clipping.go:54-74.
template <class Emit>
__device__ __noinline__ void clip_and_emit(const DrawParams &p, Emit &e, uint32_t prim, const V4 o[3]) {
    const MeshPlanes &m = p.mesh;
    // The polygon first: a triangle that lies outside the view volume altogether (everything beyond the frame of a
    // close-up, the caps of the 10 M-triangle sphere that the 30-degree field of view cuts off) clips to nothing, and
    // the 36 attribute loads, FixNormals and the fan below are only needed for what survives.  (They used to come
    // first: a fully outside triangle cost three times a visible one -- the sort-last ranks that hold such ranges
    // spent 0.19 ms on 1.25 M triangles that draw nothing.)
    V4 np[MAX_POLY];
    np[0] = o[0]; np[1] = o[1]; np[2] = o[2];
    const int n = sutherland_hodgman(np, 3);
    if (n < 3) return;
    FullVertex t[3];
#pragma unroll
    for (uint32_t v = 0; v < 3; v++) {
        t[v].pos = v3(plane_at(m.pos, m.n, v, 3, 0, prim), plane_at(m.pos, m.n, v, 3, 1, prim),
                      plane_at(m.pos, m.n, v, 3, 2, prim));
        t[v].nrm = v3(plane_at(m.nrm, m.n, v, 3, 0, prim), plane_at(m.nrm, m.n, v, 3, 1, prim),
                      plane_at(m.nrm, m.n, v, 3, 2, prim));
        t[v].tu = plane_at(m.tex, m.n, v, 2, 0, prim);
        t[v].tv = plane_at(m.tex, m.n, v, 2, 1, prim);
        t[v].col = c4(plane_at(m.col, m.n, v, 4, 0, prim), plane_at(m.col, m.n, v, 4, 1, prim),
                      plane_at(m.col, m.n, v, 4, 2, prim), plane_at(m.col, m.n, v, 4, 3, prim));
        t[v].out = o[v];
    }
    fix_normals(t);  // NewTriangle(v1, v2, v3), context.go:378
    const V3 p1 = w_xyz(o[0]), p2 = w_xyz(o[1]), p3 = w_xyz(o[2]);
    for (int i = 2; i < n; i++) {
        FullVertex nv[3];
        nv[0] = interpolate_vertexes(t, barycentric(p1, p2, p3, w_xyz(np[0])));
        nv[1] = interpolate_vertexes(t, barycentric(p1, p2, p3, w_xyz(np[i - 1])));
        nv[2] = interpolate_vertexes(t, barycentric(p1, p2, p3, w_xyz(np[i])));
        fix_normals(nv);  // NewTriangle, clipping.go:71
        const V4 oo[3] = {nv[0].out, nv[1].out, nv[2].out};
        if (Emit::kWrite) {
            // only allocate a pool slot if the fan triangle survives culling: probe with a counter
            CountEmit probe{0, 0, 0};
            emit_clipped_triangle(p, probe, oo, 0, 0);
            if (probe.n == 0) continue;
            if constexpr (Emit::kWrite) {
                if (!e.wants(probe.n)) { e.skip(probe.n); continue; }  // none of its records is wanted in this pass
            }
            const uint32_t slot = e.pool_alloc(nv, prim);
            emit_clipped_triangle(p, e, oo, slot, REC_SRC_POOL);
        } else {
            emit_clipped_triangle(p, e, oo, 0, REC_SRC_POOL);
        }
    }
}

template <class Emit>
FGL_DI void process_triangle(const DrawParams &p, Emit &e, uint32_t prim) {
    const MeshPlanes &m = p.mesh;
    V4 o[3];
    bool outside = false;
#pragma unroll
    for (uint32_t v = 0; v < 3; v++) {
        const V3 pos = v3(plane_at(m.pos, m.n, v, 3, 0, prim), plane_at(m.pos, m.n, v, 3, 1, prim),
                          plane_at(m.pos, m.n, v, 3, 2, prim));
        o[v] = m_mul_position_w(p.matrix, pos);  // Shader.Vertex, shader.go:20,39,70
        outside = outside || w_outside(o[v]);
    }
    if (outside) clip_and_emit(p, e, prim, o);
    else emit_clipped_triangle(p, e, o, prim, 0);
}

// DrawLine, context.go:351-368 + ClipLine, clipping.go:76-98 + drawClippedLine, context.go:303-314.
template <class Emit>
FGL_DI void process_line(const DrawParams &p, Emit &e, uint32_t prim) {
    const MeshPlanes &m = p.mesh;
    V4 w1, w2;
    {
        const V3 a = v3(plane_at(m.pos, m.n, 0, 3, 0, prim), plane_at(m.pos, m.n, 0, 3, 1, prim),
                        plane_at(m.pos, m.n, 0, 3, 2, prim));
        const V3 b = v3(plane_at(m.pos, m.n, 1, 3, 0, prim), plane_at(m.pos, m.n, 1, 3, 1, prim),
                        plane_at(m.pos, m.n, 1, 3, 2, prim));
        w1 = m_mul_position_w(p.matrix, a);
        w2 = m_mul_position_w(p.matrix, b);
    }
    if (w_outside(w1) || w_outside(w2)) {
        for (int pi = 0; pi < 6; pi++) {
            const ClipPlane pl = c_clip_planes[pi];
            const bool f1 = point_in_front(pl, w1), f2 = point_in_front(pl, w2);
            if (f1 && f2) continue;
            else if (f1) w2 = intersect_segment(pl, w1, w2);
            else if (f2) w1 = intersect_segment(pl, w2, w1);
            else return;
        }
    }
    const V3 ndc0 = v3(w1.x / w1.w, w1.y / w1.w, w1.z / w1.w);
    const V3 ndc1 = v3(w2.x / w2.w, w2.y / w2.w, w2.z / w2.w);
    const V3 s0 = m_mul_position(p.screen, ndc0), s1 = m_mul_position(p.screen, ndc1);
    emit_line(p, e, s0, s1, 0, 1, w1.w, w2.w, prim, 0);
}

// ---- single-pass compaction without inter-block waits ------------------------------------------------
// Every block scans its own (scanlines << 30 | records) counts, takes a region of the record array
// with one atomic and writes its records there -- in primitive order inside the region, regions in
// arrival order.  The block aggregates are scanned by whichever block finishes last, and k_rec_index
// lists the record slots in primitive order (rec_slot, rec_row_off) for the span stage.  An ordered
// look-back was measured first: blocks behind a slow block (many visible triangles) idled at the barrier
// for their prefix (44 % of warp time, profiles/README.md).
constexpr int GT = 256;
constexpr unsigned long long REC_MASK = (1ull << 30) - 1ull;

// General path (lines, wireframe, clipped triangles), out of line so that its registers and stack
// do not weigh on the common case.  Counting and writing run the same deterministic arithmetic.
__device__ __noinline__ void count_general(const DrawParams &p, uint32_t prim, uint32_t *n, uint32_t *rows,
                                           unsigned long long *cells = nullptr) {
    CountEmit e{0, 0, 0};
    if (p.is_lines) process_line(p, e, prim);
    else process_triangle(p, e, prim);
    *n = e.n;
    *rows = e.rows;
    if (cells) *cells = e.cells;
}
__device__ __noinline__ void write_general(const DrawParams &p, const WorkBuffers &wb, uint32_t prim, uint32_t rec0,
                                           uint32_t row0) {
    WriteEmit e{rec0, row0, &wb};
    if (p.is_lines) process_line(p, e, prim);
    else process_triangle(p, e, prim);
}

__global__ void __launch_bounds__(GT, 4)
k_geometry(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb) {
    __shared__ uint32_t s_region;
    __shared__ bool s_last;
    __shared__ unsigned long long s_scan[GT / 32 + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t vb = blockIdx.x;
    const uint32_t nblocks = (p.count + GT - 1) / GT;
    const uint32_t i = vb * GT + tid;

    // Fast path, kept lean for occupancy: a triangle entirely inside the view volume, not in wireframe
    // mode, yields at most one record, which stays in registers until its ordered slot is known.
    // Everything else (lines, wireframe, clipped triangles) goes through the out-of-line general path.
    uint32_t n = 0, rows = 0;
    bool slow = false;
    V3 s0, s1, s2;
    double w0 = 0, w1 = 0, w2 = 0;
    uint32_t vflags = 0;
    BBox bb;
    bb.visible = false;
    if (i < p.count) {
        const uint32_t prim = p.first + i;
        if (p.is_lines || p.state.wireframe) {
            slow = true;
        } else {
            const MeshPlanes &m = p.mesh;
            V4 o[3];
            bool outside = false;
#pragma unroll
            for (uint32_t v = 0; v < 3; v++) {
                const V3 pos = v3(plane_at(m.pos, m.n, v, 3, 0, prim), plane_at(m.pos, m.n, v, 3, 1, prim),
                                  plane_at(m.pos, m.n, v, 3, 2, prim));
                o[v] = m_mul_position_w(p.matrix, pos);  // Shader.Vertex, shader.go:20,39,70
                outside = outside || w_outside(o[v]);
            }
            if (outside) {
                slow = !clips_to_nothing(o);  // (nothing survives ClipTriangle: n stays 0)
            } else {  // drawClippedTriangle, context.go:316-341
                V3 ndc0 = v3(o[0].x / o[0].w, o[0].y / o[0].w, o[0].z / o[0].w);
                V3 ndc1 = v3(o[1].x / o[1].w, o[1].y / o[1].w, o[1].z / o[1].w);
                V3 ndc2 = v3(o[2].x / o[2].w, o[2].y / o[2].w, o[2].z / o[2].w);
                double a = (ndc1.x - ndc0.x) * (ndc2.y - ndc0.y) - (ndc2.x - ndc0.x) * (ndc1.y - ndc0.y);
                uint32_t i0 = 0, i2 = 2;
                if (a < 0) {
                    V3 t = ndc0; ndc0 = ndc2; ndc2 = t;
                    i0 = 2; i2 = 0;
                }
                if (p.state.cull == FGL_CULL_FRONT) a = -a;
                if (p.state.front_face == FGL_FACE_CW) a = -a;
                if (!(p.state.cull != FGL_CULL_NONE && a <= 0)) {
                    s0 = m_mul_position(p.screen, ndc0);
                    s1 = m_mul_position(p.screen, ndc1);
                    s2 = m_mul_position(p.screen, ndc2);
                    bb = compute_bbox(p, s0, s1, s2);
                    if (bb.visible) {
                        n = 1; rows = bb.rows;
                        w0 = i0 == 0 ? o[0].w : o[2].w; w1 = o[1].w; w2 = i2 == 2 ? o[2].w : o[0].w;
                        vflags = vmap3(i0, 1, i2);
                    }
                }
            }
        }
        if (slow) count_general(p, prim, &n, &rows);
    }
    // block-wide exclusive scan of (rows << 30 | records)
    const unsigned long long mine = ((unsigned long long)rows << 30) | n;
    unsigned long long incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = lane < GT / 32 ? s_scan[lane] : 0, vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < GT / 32) s_scan[lane] = vi - v;
        if (lane == GT / 32 - 1) s_scan[GT / 32] = vi;
    }
    __syncthreads();
    const unsigned long long local_excl = s_scan[warp] + incl - mine;
    const unsigned long long block_total = s_scan[GT / 32];

    if (tid == 0) {
        const uint32_t brec = (uint32_t)(block_total & REC_MASK);
        s_region = brec ? atomicAdd(&wb.counters->rec_cursor, brec) : 0u;
        wb.blk_agg[vb] = block_total;
        wb.blk_region[vb] = s_region;
        __threadfence();
        s_last = atomicAdd(&wb.counters->blocks_done, 1u) == nblocks - 1u;
    }
    __syncthreads();
    const uint32_t rec0 = s_region + (uint32_t)(local_excl & REC_MASK), row0 = (uint32_t)(local_excl >> 30);
    if (slow) {
        if (n > 0) write_general(p, wb, p.first + i, rec0, row0);  // run again, writing at the block's slots
    } else if (n == 1) {
        write_record(&wb, rec0, row0, bb, s0, s1, s2, w0, w1, w2, p.first + i, vflags);
    }
    if (!s_last) return;

    // Last block: exclusive scan of the block aggregates (a few thousand entries), grand totals.
    __threadfence();
    const uint32_t per = (nblocks + GT - 1) / GT;
    const uint32_t b0 = min((uint32_t)tid * per, nblocks), b1 = min(b0 + per, nblocks);
    unsigned long long sum = 0;
    for (uint32_t b = b0; b < b1; b++) sum += __ldcg(&wb.blk_agg[b]);
    unsigned long long incl2 = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl2, o);
        if (lane >= o) incl2 += t;
    }
    __syncthreads();  // s_scan is reused
    if (lane == 31) s_scan[warp] = incl2;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = lane < GT / 32 ? s_scan[lane] : 0, vi = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < GT / 32) s_scan[lane] = vi - v;
        if (lane == GT / 32 - 1) s_scan[GT / 32] = vi;
    }
    __syncthreads();
    unsigned long long run = s_scan[warp] + incl2 - sum;
    for (uint32_t b = b0; b < b1; b++) {
        wb.blk_base[b] = run;
        run += __ldcg(&wb.blk_agg[b]);
    }
    if (tid == 0) {
        const unsigned long long tot = s_scan[GT / 32];
        const uint32_t nrec = (uint32_t)(tot & REC_MASK), nrows = (uint32_t)(tot >> 30);
        DrawCounters *c = wb.counters;
        c->n_records = nrec; c->need_records = nrec;
        c->n_rows = nrows; c->need_rows = nrows;
        unsigned ovf = 0;
        if (nrec > wb.cap_records) ovf |= OVF_RECORDS;
        if (nrows > wb.cap_rows) ovf |= OVF_ROWS;
        if (ovf) atomicOr(&c->overflow, ovf);
        if (nrec <= wb.cap_records) wb.rec_row_off[nrec] = nrows;  // sentinel for the span stage's search
        wb.tile_ctl->nheavy = 0; wb.tile_ctl->nlight = 0; wb.tile_ctl->head = 0;  // the busy-strip list of this draw
    }
}

// Record slots in primitive order: rec_slot[c], and the first span-stage item of record c, rec_row_off[c].
__global__ void __launch_bounds__(GT)
k_rec_index(const __grid_constant__ WorkBuffers wb, uint32_t nblocks) {
    if (wb.counters->overflow) return;
    for (uint32_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const unsigned long long agg = wb.blk_agg[b], base = wb.blk_base[b];
        const uint32_t cnt = (uint32_t)(agg & REC_MASK), region = wb.blk_region[b];
        const uint32_t rec_base = (uint32_t)(base & REC_MASK), row_base = (uint32_t)(base >> 30);
        for (uint32_t k = threadIdx.x; k < cnt; k += GT) {
            const uint32_t slot = region + k, c = rec_base + k;
            wb.rec_slot[c] = slot;
            wb.rec_row_off[c] = row_base + wb.rec_local_row[slot];
        }
    }
}


// ================================================================================================
// Fused front end of large draws: geometry + span stage in ONE kernel.
//
// A block takes FT consecutive primitives.  Its raster records never leave the SM: they are set up in
// shared memory, and the same threads then walk their scanlines -- FT (record, scanline) items at a time,
// the item -> record map being a binary search in a shared-memory offset table -- count the segments of
// each row (the first one stays in registers), compact them in order with a block scan and store them as
// SegV.  Compared with k_geometry -> k_rec_index -> k_span_walk -> scan x3 -> k_span_place this removes
// the 176-byte record round trip through HBM, the per-row record search in global memory, the global row
// arrays and six launches (profiles/README.md).
//
// Order without waiting for other blocks: a block reserves, with one atomic, a region of the
// segment array large enough for an upper bound of its segments (sum over records of rows x strip columns,
// known after the geometry phase) and fills it from the front in (primitive, scanline, column) order; regions
// are in block-arrival order.  Every block publishes its exact segment count, the last block to finish scans
// the counts, and k_seg_index lists the segments in primitive order for the stable sort by strip.
constexpr int FT = FRONT_FT;  // (fgl_internal.h, with FRONT_GROUP: front-end blocks per group sum)
constexpr unsigned long long CELL_SHIFT = 24, NREC_MASK = (1ull << CELL_SHIFT) - 1ull;
static_assert(FT * 64 < (1 << CELL_SHIFT), "records of one block fit the low bits of its aggregate");
static_assert((FT & (FT - 1)) == 0, "the item -> record search halves a power of two");

struct __align__(16) SRec {  // a raster record in shared memory: RowSetup + the back end's tail
    double s0x, s0y, s1x, s1y, s2x, s2y;
    double w00, w01, w02, ra, ra12, ra20, ra01;
    double z0, z1, z2;
    double r0, r1, r2;
    int32_t x0, x1, y0;
    uint32_t rows;
    uint32_t src, flags;
};
static_assert(sizeof(SRec) == 176, "SRec layout");

// Per-triangle setup of Context.rasterize, context.go:155-181 (the same arithmetic as write_record).
// ystart > b.y0: the per-row adds of the rows in front of ystart (context.go:275-277) are executed here, once, and
// the record's chain then starts at ystart -- the same additions in the same order as replaying them in every row.
struct SetupRcp { double ra, r0, r1, r2, ra12, ra20, ra01; };  // the seven divisions of context.go:163-181
FGL_DI SetupRcp setup_rcp(V3 s0, V3 s1, V3 s2, double w0, double w1, double w2) {
    SetupRcp q;
    const double a01 = s1.y - s0.y, a12 = s2.y - s1.y, a20 = s0.y - s2.y;
    q.ra = 1 / edge_fn(s0, s1, s2);
    q.r0 = 1 / w0; q.r1 = 1 / w1; q.r2 = 1 / w2;
    q.ra12 = 1 / a12; q.ra20 = 1 / a20; q.ra01 = 1 / a01;
    return q;
}
// The same seven reciprocals side by side (fgl_math.cuh: the fast path of the compiler's own expansion as
// straight-line code, one range test for all; the operator itself when it fails -- a zero edge slope, say).
#ifndef FGL_FRONT_FASTDIV
#define FGL_FRONT_FASTDIV 1
#endif
FGL_DI SetupRcp setup_rcp_fast(V3 s0, V3 s1, V3 s2, double w0, double w1, double w2) {
#if FGL_FRONT_FASTDIV
    SetupRcp q;
    const double a01 = s1.y - s0.y, a12 = s2.y - s1.y, a20 = s0.y - s2.y;
    const double area = edge_fn(s0, s1, s2);
    bool ok = true;
    q.ra = rcp_fast(area, ok);
    q.r0 = rcp_fast(w0, ok); q.r1 = rcp_fast(w1, ok); q.r2 = rcp_fast(w2, ok);
    q.ra12 = rcp_fast(a12, ok); q.ra20 = rcp_fast(a20, ok); q.ra01 = rcp_fast(a01, ok);
    if (!ok) {
        q.ra = 1 / area;
        q.r0 = 1 / w0; q.r1 = 1 / w1; q.r2 = 1 / w2;
        q.ra12 = 1 / a12; q.ra20 = 1 / a20; q.ra01 = 1 / a01;
    }
    return q;
#else
    return setup_rcp(s0, s1, s2, w0, w1, w2);
#endif
}
// Vector{X / W, Y / W, Z / W} of the three vertices (context.go:318-320): one reciprocal refinement per W.
FGL_DI void ndc_divide(const V4 *o, V3 &n0, V3 &n1, V3 &n2) {
#if FGL_FRONT_FASTDIV
    bool ok = true;
    const double y0 = div_refine(o[0].w), y1 = div_refine(o[1].w), y2 = div_refine(o[2].w);
    n0 = v3(div_tail(o[0].x, o[0].w, y0, ok), div_tail(o[0].y, o[0].w, y0, ok), div_tail(o[0].z, o[0].w, y0, ok));
    n1 = v3(div_tail(o[1].x, o[1].w, y1, ok), div_tail(o[1].y, o[1].w, y1, ok), div_tail(o[1].z, o[1].w, y1, ok));
    n2 = v3(div_tail(o[2].x, o[2].w, y2, ok), div_tail(o[2].y, o[2].w, y2, ok), div_tail(o[2].z, o[2].w, y2, ok));
    if (ok) return;
#endif
    n0 = v3(o[0].x / o[0].w, o[0].y / o[0].w, o[0].z / o[0].w);
    n1 = v3(o[1].x / o[1].w, o[1].y / o[1].w, o[1].z / o[1].w);
    n2 = v3(o[2].x / o[2].w, o[2].y / o[2].w, o[2].z / o[2].w);
}
FGL_DI void fill_srec(SRec &r, const BBox &b, V3 s0, V3 s1, V3 s2, uint32_t src, uint32_t flags, const SetupRcp &q,
                      int ystart) {
    r.s0x = s0.x; r.s0y = s0.y; r.s1x = s1.x; r.s1y = s1.y; r.s2x = s2.x; r.s2y = s2.y;
    r.z0 = s0.z; r.z1 = s1.z; r.z2 = s2.z;
    const V3 pc = b.origin ? v3(-9223372036854775808.0, -9223372036854775808.0, 0) : v3((double)b.x0 + 0.5, (double)b.y0 + 0.5, 0);
    double w00 = edge_fn(s1, s2, pc), w01 = edge_fn(s2, s0, pc), w02 = edge_fn(s0, s1, pc);
    if (ystart > b.y0) {
        const double b01 = s0.x - s1.x, b12 = s1.x - s2.x, b20 = s2.x - s0.x;  // context.go:168-172
        for (int yy = b.y0; yy < ystart; yy++) { w00 += b12; w01 += b20; w02 += b01; }
    }
    r.w00 = w00; r.w01 = w01; r.w02 = w02;
    r.ra = q.ra;
    r.r0 = q.r0; r.r1 = q.r1; r.r2 = q.r2;
    r.ra12 = q.ra12; r.ra20 = q.ra20; r.ra01 = q.ra01;
    r.x0 = b.x0; r.x1 = b.x1; r.y0 = max(b.y0, ystart); r.rows = b.rows;
    r.src = src; r.flags = flags | (b.origin ? REC_WRAP : 0u);
}

// Emits the records of one primitive whose block-local index falls into the window [win0, win1).
struct SmemEmit {
    uint32_t next, win0, win1;
    SRec *s_rec;
    const WorkBuffers *wb;
    static constexpr bool kWrite = true;
    FGL_DI bool wants(uint32_t n) const { return next < win1 && next + n > win0; }
    FGL_DI void skip(uint32_t n) { next += n; }
    FGL_DI void record(const DrawParams &p, V3 s0, V3 s1, V3 s2, double w0, double w1, double w2, uint32_t src,
                       uint32_t flags) {
        const BBox b = compute_bbox(p, s0, s1, s2);
        if (!b.visible) return;
        if (next >= win0 && next < win1)
            fill_srec(s_rec[next - win0], b, s0, s1, s2, src, flags, setup_rcp(s0, s1, s2, w0, w1, w2), b.y0);
        next++;
    }
    FGL_DI uint32_t pool_alloc(const FullVertex *v, uint32_t prim) {
        WriteEmit w{0, 0, wb};
        return w.pool_alloc(v, prim);
    }
};
__device__ __noinline__ void emit_general_smem(const DrawParams &p, const WorkBuffers &wb, uint32_t prim, uint32_t rec0,
                                               uint32_t win0, uint32_t win1, SRec *s_rec) {
    SmemEmit e{rec0, win0, win1, s_rec, &wb};
    if (p.is_lines) process_line(p, e, prim);
    else process_triangle(p, e, prim);
}

// Conservative back-face filter for triangles inside the view volume (all w > 0).  The reference culls on the sign
// of a = (ndc1.X-ndc0.X)*(ndc2.Y-ndc0.Y) - (ndc2.X-ndc0.X)*(ndc1.Y-ndc0.Y) computed in float64 from the divided
// coordinates (context.go:318-336).  In real arithmetic a = D / (w0 w1 w2) with D = det[(x0 y0 w0)(x1 y1 w1)(x2 y2 w2)];
// |ndc| <= 1 bounds the rounding error of the computed a by ~50 ulp(1) = 5.6e-15, and the error of D evaluated
// below by ~8 ulp of M, the same sum with absolute values.  So when s*D < -(1e-12 w0 w1 w2 + 1e-14 M), s = +-1 the
// orientation the state culls, the reference's a has the culled sign for certain and the nine divisions, the
// screen transform and the set-up can be skipped; every other triangle (and any NaN/Inf) takes the full path, whose
// arithmetic decides.  Whole warps of back-facing triangles (they come in runs) then cost a few multiplications.
FGL_DI bool surely_culled(const DrawParams &p, const V4 *o) {
    if (p.state.cull == FGL_CULL_NONE) return false;
    const double p1 = o[1].y * o[2].w, p2 = o[2].y * o[1].w, p3 = o[1].x * o[2].w, p4 = o[2].x * o[1].w;
    const double p5 = o[1].x * o[2].y, p6 = o[2].x * o[1].y;
    double d = (o[0].x * (p1 - p2) - o[0].y * (p3 - p4)) + o[0].w * (p5 - p6);
    const double m = (fabs(o[0].x) * (fabs(p1) + fabs(p2)) + fabs(o[0].y) * (fabs(p3) + fabs(p4))) +
                     fabs(o[0].w) * (fabs(p5) + fabs(p6));
    const double w = o[0].w * o[1].w * o[2].w;
    if ((p.state.cull == FGL_CULL_FRONT) != (p.state.front_face == FGL_FACE_CW)) d = -d;
    return w > 0 && d < -(1e-12 * w + 1e-14 * m);
}

#ifndef FGL_FRONT_MINB
#define FGL_FRONT_MINB 6
#endif
#ifndef FGL_FRONT_CLOCK
#define FGL_FRONT_CLOCK 0  // tuning aid (variant build, FGL_TILE_CLOCK=1): cycles of warp 0 per phase of a fast block, summed
                           // into the debug counters behind wb.tile_clock (tools/front_cycles.py)
#endif
#if FGL_FRONT_CLOCK
#define FRONT_TICK(k) do { if (wb.tile_clock && tid == 0) { const long long t_ = clock64(); fc_dt[k] = t_ - fc_t; fc_t = t_; } } while (0)
#else
#define FRONT_TICK(k) do { } while (0)
#endif
#ifndef FGL_FRONT_TIGHT
#define FGL_FRONT_TIGHT 1  // drop the rows / columns of a box that provably keep no pixel (tighten_box)
#endif
#ifndef FGL_FRONT_PREFETCH
#define FGL_FRONT_PREFETCH 0
#endif
#ifndef FGL_FRONT_COMPACT
#define FGL_FRONT_COMPACT 0  // 1: queue the scanlines that cover something and walk their runs with full warps.  Measured
                             // SLOWER (k_front 100.4 -> 112.7 us at 1080p, 301 -> 330 us at 8K): the ring traffic and the
                             // 10 KB of shared memory cost more than the idle lanes it removes.  Kept as a tuning variant.
#endif
#if FGL_FRONT_COMPACT
struct __align__(8) QEntry {  // a scanline whose run has been found: edge values at its first covered pixel
    double w0, w1, w2;
    int32_t x, y;
    uint32_t ridx, _pad;
};
constexpr uint32_t QCAP = 64;  // ring entries per warp: up to 31 left over + 32 new
#endif
__global__ void __launch_bounds__(FT, FGL_FRONT_MINB)
k_front(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb) {
    __shared__ SRec s_rec[FT];
    __shared__ uint32_t s_rowoff[FT + 1];
    __shared__ uint16_t s_order[FT];
    __shared__ unsigned long long s_scan[FT / 32];
    __shared__ uint32_t s_scanr[FT / 32];
    __shared__ uint32_t s_scan32[2][FT / 32];
    uint32_t scan_parity = 0;
    __shared__ unsigned long long s_region;
    // fast blocks (no clipping / lines / wireframe): every warp compacts into a sub-region of its own
    __shared__ uint32_t s_celloff[FT];           // cells (rows x strip columns) of the records before this one
    __shared__ volatile uint32_t s_region_ready;
#if FGL_FRONT_COMPACT
    __shared__ QEntry s_ring[FT / 32][QCAP];
#endif
    if (threadIdx.x == 0) s_region_ready = 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#if FGL_FRONT_CLOCK
    long long fc_t = clock64(), fc_dt[6] = {0, 0, 0, 0, 0, 0};
    const long long fc_t0 = fc_t;
#endif
    pdl_trigger();  // k_seg_index may be scheduled while the last wave of this grid drains
    const uint32_t vb = blockIdx.x;
    const uint32_t i = vb * FT + tid;
    const uint32_t prim = p.first + i;
#if FGL_FRONT_PREFETCH
    // Tuning aid (off): every block pulls the nine 1-KB position-plane pieces of the block FGL_FRONT_PREFETCH further on
    // into L2.  The first use of the position planes is the kernel's largest single stall item (11 % of the samples),
    // but 0 / 512 / 1024 / 2048 blocks ahead all measured the same (k_front 102.6 / 101.2 / 102.6 / 102.4 us): the
    // other resident blocks already cover that latency.
    if (!p.is_lines && tid < 72) {
        const uint32_t ahead = (vb + (uint32_t)FGL_FRONT_PREFETCH) * FT + (uint32_t)(tid & 7) * 16u;
        if (ahead < p.count)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p.mesh.pos + (size_t)(tid >> 3) * p.mesh.n + p.first + ahead));
    }
#endif

    // ---- geometry: one thread per primitive -----------------------------------------------------
    // Fast path: a triangle entirely inside the view volume, not in wireframe mode, yields at most one
    // record, set up straight into this thread's shared-memory slot.  Lines, wireframe and triangles
    // that need clipping are only counted here (out-of-line general path).
    uint32_t n = 0, frows = 0;  // frows: scanlines of this thread's fast-path record
    unsigned long long cells = 0;
    bool slow = false;
    if (i < p.count) {
        if (p.is_lines || p.state.wireframe) {
            slow = true;
        } else {
            const MeshPlanes &m = p.mesh;
            V4 o[3];
            bool outside = false;
#pragma unroll
            for (uint32_t v = 0; v < 3; v++) {
                const V3 pos = v3(plane_at(m.pos, m.n, v, 3, 0, prim), plane_at(m.pos, m.n, v, 3, 1, prim),
                                  plane_at(m.pos, m.n, v, 3, 2, prim));
                o[v] = m_mul_position_w(p.matrix, pos);  // Shader.Vertex, shader.go:20,39,70
                outside = outside || w_outside(o[v]);
            }
            if (outside) {
                slow = !clips_to_nothing(o);  // (nothing survives ClipTriangle: n stays 0)
            } else if (surely_culled(p, o)) {
                // (n stays 0: the reference computes a signed area that is provably on the culled side)
            } else {  // drawClippedTriangle, context.go:316-341
                V3 ndc0, ndc1, ndc2;
                ndc_divide(o, ndc0, ndc1, ndc2);
                double a = (ndc1.x - ndc0.x) * (ndc2.y - ndc0.y) - (ndc2.x - ndc0.x) * (ndc1.y - ndc0.y);
                uint32_t i0 = 0, i2 = 2;
                if (a < 0) {
                    V3 t = ndc0; ndc0 = ndc2; ndc2 = t;
                    i0 = 2; i2 = 0;
                }
                if (p.state.cull == FGL_CULL_FRONT) a = -a;
                if (p.state.front_face == FGL_FACE_CW) a = -a;
                if (!(p.state.cull != FGL_CULL_NONE && a <= 0)) {
                    const V3 s0 = m_mul_position(p.screen, ndc0), s1 = m_mul_position(p.screen, ndc1),
                             s2 = m_mul_position(p.screen, ndc2);
                    BBox bb = compute_bbox(p, s0, s1, s2);
                    if (bb.visible) {
                        // the seven divisions of the set-up first: the tightening below schedules between them
                        const SetupRcp q = setup_rcp_fast(s0, s1, s2, i0 == 0 ? o[0].w : o[2].w, o[1].w, i2 == 2 ? o[2].w : o[0].w);
#if FGL_FRONT_TIGHT
                        const int ystart = tighten_box(p, bb, q.ra);
#else
                        const int ystart = bb.y0;
#endif
                        if (bb.visible) {
                            n = 1;
                            frows = bb.rows;
                            cells = (unsigned long long)bb.rows * bb.cols;
                            fill_srec(s_rec[tid], bb, s0, s1, s2, prim, vmap3(i0, 1, i2), q, ystart);
                        }
                    }
                }
            }
        }
        if (slow) {
            uint32_t rows_unused;
            count_general(p, prim, &n, &rows_unused, &cells);
        }
    }
    FRONT_TICK(0);  // phase 1 of this thread
    // Block-wide exclusive scans of (cells << CELL_SHIFT | records) and of the rows of the fast-path records, with ONE
    // barrier: the one that tells whether the block needs the general path and orders the s_rec writes.  (Six barriers
    // -- or-reduction, a two-level scan, the order table, a second scan for the rows -- used to separate the geometry
    // from the walk: 4 000 of a block's 21 700 cycles, tools/front_cycles.py.)
    const unsigned long long mine = (cells << CELL_SHIFT) | n;
    unsigned long long incl = mine;
    uint32_t rincl = frows;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        const uint32_t tr = __shfl_up_sync(0xffffffffu, rincl, o);
        if (lane >= o) { incl += t; rincl += tr; }
    }
    if (lane == 31) { s_scan[warp] = incl; s_scanr[warp] = rincl; }
    const bool general = __syncthreads_or(slow && n > 0) != 0;
    FRONT_TICK(1);  // waiting for the block's slowest thread
    unsigned long long excl = incl - mine, block_total = 0;
    uint32_t rexcl = rincl - frows, items_all = 0;
#pragma unroll
    for (int w = 0; w < FT / 32; w++) {
        const unsigned long long t = s_scan[w];
        const uint32_t tr = s_scanr[w];
        if (w < warp) { excl += t; rexcl += tr; }
        block_total += t; items_all += tr;
    }
    const uint32_t rec_off = (uint32_t)(excl & NREC_MASK);
    const uint32_t nrec_blk = (uint32_t)(block_total & NREC_MASK);
    // The reservation's round trip is not waited for here: thread 0 keeps the old cursor in a register and hands it
    // to the block just before the first compaction barrier below, a scan and a row walk later.
    unsigned long long my_region = 0;
    if (tid == 0 && nrec_blk) my_region = atomicAdd(&wb.counters->seg_cursor, block_total >> CELL_SHIFT);
    if (!general) {
        if (n == 1) {  // compacted record rec_off: its thread, the cells and the (record, scanline) items before it
            s_order[rec_off] = (uint16_t)tid;
            s_celloff[rec_off] = (uint32_t)(excl >> CELL_SHIFT);
            s_rowoff[rec_off] = rexcl;
        }
        if (tid >= (int)nrec_blk) s_rowoff[tid] = items_all;
        if (tid == 0) s_rowoff[FT] = items_all;
    }
    __syncthreads();
    unsigned long long region = 0;

    // ---- spans: FT records per window, FT (record, scanline) items per pass --------------------------------
    uint32_t seg_run = 0;  // segments this block has stored so far (block-uniform)
    unsigned long long covered = 0;
    ParkedSeg first;
    first.w0 = first.w1 = first.w2 = 0; first.x = 0; first.cnt = 0; first.key = 0; first.wrap = 0;
    if (!general) {
        // ---- fast blocks: no block barrier inside the walk.  The (record, scanline) items are split into FT/32
        // contiguous ranges (multiples of 32), one per warp; a warp's segments go to the block's region at the
        // offset given by the CELL upper bound of the items before its range -- the bound the region was sized
        // with -- so the warps never wait for each other's counts (the block scan per 128 items and its barrier
        // were 9 % of the kernel's stall samples).  Inside the region the warps' runs are in primitive order but
        // not contiguous; k_seg_index closes the gaps (blk_wcnt / blk_woff).
        const uint32_t items = items_all;
        // Hand the reservation to the block: a release store of the flag after the value, acquire loads in the
        // readers (message passing without a barrier -- compute-sanitizer's racecheck reports exactly this pair).
        // Warp 0 absorbs what is left of the atomic's round trip; the other warps look only after their first walk.
        const uint32_t flag_addr = (uint32_t)__cvta_generic_to_shared(const_cast<uint32_t *>(&s_region_ready));
        if (tid == 0) {
            s_region = my_region;
            asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(flag_addr), "r"(1u) : "memory");
        }
        // the last entry with s_rowoff[lo] <= it (it < items, warp-uniform; entries >= nrec_blk hold `items`): the table
        // is sorted and starts at 0, so lo = (entries <= it) - 1 -- four independent loads per lane and one reduction
        auto locate = [&](uint32_t it) {
            uint32_t cnt = 0;
#pragma unroll
            for (int j = 0; j < FT / 32; j++) cnt += s_rowoff[lane + 32 * j] <= it ? 1u : 0u;
            return __reduce_add_sync(0xffffffffu, cnt) - 1u;
        };
        const uint32_t ipw = (((items + FT / 32 - 1) / (FT / 32)) + 31u) & ~31u;
        const uint32_t my0 = min((uint32_t)warp * ipw, items), my1 = min(my0 + ipw, items);
        uint32_t wbase = 0, wrun = 0;
        // Item -> record without a search per item: the record of the warp's first item is found once (lo_base);
        // inside a chunk of 32 consecutive items, lane j looks up where record lo_base + 1 + j begins, the lanes OR
        // those boundaries into one mask, and an item's record is lo_base + the boundaries at or before it.
        uint32_t lo_base = 0;
        if (my0 < my1) {
            lo_base = locate(my0);
            const SRec &r = s_rec[s_order[lo_base]];
            wbase = s_celloff[lo_base] + (my0 - s_rowoff[lo_base]) * box_cols(p, r.x0, r.x1);
        }
        bool have_region = false;
        FRONT_TICK(2);  // scans, item ranges, the first record lookup
#if FGL_FRONT_COMPACT
        // Two thirds of the (record, scanline) items of sub-pixel triangles cover nothing, and the run loop and the
        // 128-byte segment store behind them ran with a dozen of 32 lanes.  So the walk is split: every lane first
        // FINDS the run of its item (replay, skip-ahead, the pixels left of the run), the items that have one are
        // queued -- in item order -- in a small per-warp ring in shared memory, and whenever 32 are waiting a full
        // warp walks their runs and stores their segments.  Slots are still handed out in item order (the ring is a
        // FIFO), so the segment order is unchanged.
        QEntry *const ring = s_ring[warp];
        uint32_t qhead = 0, qn = 0;  // warp-uniform
        auto drain = [&](uint32_t take) {  // the first `take` (<= 32) queued items
            uint32_t nseg = 0, ridx = 0;
            int y = 0;
            if ((uint32_t)lane < take) {
                const QEntry e = ring[(qhead + (uint32_t)lane) % QCAP];
                ridx = e.ridx; y = e.y;
                const SRec &r = s_rec[ridx];
                const unsigned long long before = covered;
                nseg = walk_row_run(p, r, y, e.w0, e.w1, e.w2, e.x, first, &covered);
                if (p.prim_info && covered != before)  // per-primitive TotalPixels (fgl_draw_*_each)
                    atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, r.src, r.flags)], covered - before);
            }
            const uint32_t incl_w = warp_incl_scan(nseg);
            const uint32_t ex = incl_w - nseg, batch_total = __shfl_sync(0xffffffffu, incl_w, 31);
            if (!have_region) {
                uint32_t ready;
                do {
                    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ready) : "r"(flag_addr) : "memory");
                } while (ready == 0u);
                region = *(volatile unsigned long long *)&s_region;
                have_region = true;
            }
            if (nseg) {
                const unsigned long long slot64 = region + wbase + wrun + ex;
                const SRec &r = s_rec[ridx];
                RecTail tail;
                tail.r0 = r.r0; tail.r1 = r.r1; tail.r2 = r.r2; tail.src = r.src; tail.flags = r.flags;
                if (slot64 + nseg <= (unsigned long long)wb.cap_segs) {
                    const uint32_t slot = (uint32_t)slot64;
                    if (nseg == 1) {
                        wb.segv[slot] = make_segv(first.w0, first.w1, first.w2, r.ra, r.z0, r.z1, r.z2, r.s2y - r.s1y,
                                                  r.s0y - r.s2y, r.s1y - r.s0y, tail, (uint16_t)first.x, (uint8_t)first.cnt, first.wrap);
                        wb.seg_key[1][slot] = first.key;
                    } else {  // the scanline crosses strip boundaries: walk it again, storing every segment
                        unsigned long long dummy = 0;
                        walk_row_segments<true>(p, r, y, first, &tail, wb.segv, wb.seg_key[1], slot, wb.cap_segs, &dummy);
                    }
                }
            }
            wrun += batch_total;
            qhead = (qhead + take) % QCAP;
            qn -= take;
            __syncwarp();
        };
        for (uint32_t c0 = my0; c0 < my1; c0 += 32) {
            const uint32_t it = c0 + (uint32_t)lane;
            const uint32_t rel = s_rowoff[min(lo_base + 1u + (uint32_t)lane, (uint32_t)FT)] - c0;  // >= 1
            const uint32_t bmask = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
            const uint32_t lo = lo_base + (uint32_t)__popc(bmask & (0xffffffffu >> (31 - lane)));
            lo_base += (uint32_t)__popc(bmask) + (__any_sync(0xffffffffu, rel == 32u) ? 1u : 0u);  // record of item c0 + 32
            bool found = false;
            QEntry e;
            e.w0 = e.w1 = e.w2 = 0; e.x = 0; e.y = 0; e.ridx = 0;
            if (it < my1) {
                e.ridx = s_order[lo];
                const SRec &r = s_rec[e.ridx];
                e.y = row_base(p, r.x1, r.y0) + (int)(it - s_rowoff[lo]);
                found = walk_row_find(p, r, e.y, e.w0, e.w1, e.w2, e.x);
            }
            const uint32_t fm = __ballot_sync(0xffffffffu, found);
            if (found) ring[(qhead + qn + (uint32_t)__popc(fm & ((1u << lane) - 1u))) % QCAP] = e;
            qn += (uint32_t)__popc(fm);
            __syncwarp();
            if (qn >= 32u) drain(32u);
        }
        if (qn) drain(qn);
#else
        for (uint32_t c0 = my0; c0 < my1; c0 += 32) {
            const uint32_t it = c0 + (uint32_t)lane;
            const uint32_t rel = s_rowoff[min(lo_base + 1u + (uint32_t)lane, (uint32_t)FT)] - c0;  // >= 1
            const uint32_t bmask = __reduce_or_sync(0xffffffffu, rel < 32u ? (1u << rel) : 0u);
            const uint32_t lo = lo_base + (uint32_t)__popc(bmask & (0xffffffffu >> (31 - lane)));
            lo_base += (uint32_t)__popc(bmask) + (__any_sync(0xffffffffu, rel == 32u) ? 1u : 0u);  // record of item c0 + 32
            uint32_t nseg = 0, ridx = 0;
            int y = 0;
            if (it < my1) {
                ridx = s_order[lo];
                const SRec &r = s_rec[ridx];
                y = row_base(p, r.x1, r.y0) + (int)(it - s_rowoff[lo]);
                const unsigned long long before = covered;
                nseg = walk_row_count(p, r, y, first, &covered);
                if (p.prim_info && covered != before)  // per-primitive TotalPixels (fgl_draw_*_each)
                    atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, r.src, r.flags)], covered - before);
            }
            const uint32_t incl_w = warp_incl_scan(nseg);
            const uint32_t ex = incl_w - nseg, chunk_total = __shfl_sync(0xffffffffu, incl_w, 31);
            if (!have_region) {
                uint32_t ready;
                do {
                    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(ready) : "r"(flag_addr) : "memory");
                } while (ready == 0u);
                region = *(volatile unsigned long long *)&s_region;
                have_region = true;
            }
            if (nseg) {
                const unsigned long long slot64 = region + wbase + wrun + ex;
                const SRec &r = s_rec[ridx];
                RecTail tail;
                tail.r0 = r.r0; tail.r1 = r.r1; tail.r2 = r.r2; tail.src = r.src; tail.flags = r.flags;
                if (slot64 + nseg <= (unsigned long long)wb.cap_segs) {
                    const uint32_t slot = (uint32_t)slot64;
                    if (nseg == 1) {
                        wb.segv[slot] = make_segv(first.w0, first.w1, first.w2, r.ra, r.z0, r.z1, r.z2, r.s2y - r.s1y,
                                                  r.s0y - r.s2y, r.s1y - r.s0y, tail, (uint16_t)first.x, (uint8_t)first.cnt, first.wrap);
                        wb.seg_key[1][slot] = first.key;
                    } else {  // the scanline crosses strip boundaries: walk it again, storing every segment
                        unsigned long long dummy = 0;
                        walk_row_segments<true>(p, r, y, first, &tail, wb.segv, wb.seg_key[1], slot, wb.cap_segs, &dummy);
                    }
                }
            }
            wrun += chunk_total;
        }
#endif
        FRONT_TICK(3);  // the walk of warp 0's items
        // Every warp publishes its own run and leaves: no barrier behind the walk (the block's warps used to wait
        // here for the slowest of the four, then for thread 0's round trip to the block counter -- 15 % of the
        // kernel's stall samples).  k_seg_index adds the runs up itself (group sums + the blocks of its group).
        if (lane == 0) {
            reinterpret_cast<uint32_t *>(wb.blk_wcnt + vb)[warp] = wrun;
            reinterpret_cast<uint32_t *>(wb.blk_woff + vb)[warp] = wbase;
            if (wrun) atomicAdd(&wb.blk_base[vb / FRONT_GROUP], (unsigned long long)wrun);
            if (warp == 0) {
                wb.blk_region[vb] = (uint32_t)min(my_region, 0xffffffffull);
                if (nrec_blk) atomicAdd(&wb.counters->n_records, nrec_blk);
            }
        }
        if (!__any_sync(0xffffffffu, (covered >> 27) != 0ull)) {  // (always, short of absurd rows) one 32-bit reduction
            covered = (unsigned long long)__reduce_add_sync(0xffffffffu, (uint32_t)covered);
        } else {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) covered += __shfl_down_sync(0xffffffffu, covered, o);
        }
        if (lane == 0 && covered) atomicAdd(&wb.counters->total_pixels, covered);  // TotalPixels, context.go:229
#if FGL_FRONT_CLOCK
        if (wb.tile_clock && lane == 0) {
            unsigned long long *dbg = wb.tile_clock + 2 * (size_t)wb.ntiles;
            if (tid == 0) {
                FRONT_TICK(4);
                for (int k = 0; k < 5; k++) atomicAdd(&dbg[k], (unsigned long long)fc_dt[k]);
                atomicAdd(&dbg[5], 1ull);
            }
            atomicAdd(&dbg[6], (unsigned long long)(clock64() - fc_t0));  // lifetime of every warp
            atomicAdd(&dbg[7], 1ull);
        }
#endif
        return;
    } else {
    for (uint32_t win0 = 0; win0 < nrec_blk; win0 += FT) {
        const uint32_t nw = min((uint32_t)FT, nrec_blk - win0);
        if (general) {  // re-run the deterministic geometry, keeping the records of this window
            __syncthreads();
            if (n > 0 && rec_off < win0 + nw && rec_off + n > win0) emit_general_smem(p, wb, prim, rec_off, win0, win0 + nw, s_rec);
            if (tid < (int)nw) s_order[tid] = (uint16_t)tid;
            __syncthreads();
        }
        uint32_t items;
        {
            const uint32_t rows = tid < (int)nw ? s_rec[s_order[tid]].rows : 0u;
            const uint32_t ex = block_excl_scan1<FT>(rows, s_scan32, scan_parity, &items);
            s_rowoff[tid] = ex;
            if (tid == 0) s_rowoff[FT] = items;
            __syncthreads();
        }
        for (uint32_t c0 = 0; c0 < items; c0 += FT) {
            const uint32_t it = c0 + tid;
            uint32_t nseg = 0, ridx = 0;
            int y = 0;
            if (it < items) {
                uint32_t lo = 0;  // the last entry with s_rowoff[lo] <= it (entries >= nw hold `items`): branch-free
#pragma unroll
                for (uint32_t step = FT / 2; step > 0; step >>= 1)
                    if (s_rowoff[lo + step] <= it) lo += step;
                ridx = s_order[lo];
                const SRec &r = s_rec[ridx];
                y = row_base(p, r.x1, r.y0) + (int)(it - s_rowoff[lo]);
                const unsigned long long before = covered;
                nseg = walk_row_count(p, r, y, first, &covered);
                if (p.prim_info && covered != before)  // per-primitive TotalPixels (fgl_draw_*_each)
                    atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, r.src, r.flags)], covered - before);
            }
            uint32_t chunk_total;
            if (tid == 0) s_region = my_region;
            const uint32_t ex = block_excl_scan1<FT>(nseg, s_scan32, scan_parity, &chunk_total);
            region = s_region;
            if (nseg) {
                const unsigned long long slot64 = region + seg_run + ex;
                const SRec &r = s_rec[ridx];
                RecTail tail;
                tail.r0 = r.r0; tail.r1 = r.r1; tail.r2 = r.r2; tail.src = r.src; tail.flags = r.flags;
                if (slot64 + nseg <= (unsigned long long)wb.cap_segs) {
                    const uint32_t slot = (uint32_t)slot64;
                    if (nseg == 1) {
                        wb.segv[slot] = make_segv(first.w0, first.w1, first.w2, r.ra, r.z0, r.z1, r.z2, r.s2y - r.s1y,
                                                  r.s0y - r.s2y, r.s1y - r.s0y, tail, (uint16_t)first.x, (uint8_t)first.cnt, first.wrap);
                        wb.seg_key[1][slot] = first.key;
                    } else {  // the scanline crosses strip boundaries: walk it again, storing every segment
                        unsigned long long dummy = 0;
                        walk_row_segments<true>(p, r, y, first, &tail, wb.segv, wb.seg_key[1], slot, wb.cap_segs, &dummy);
                    }
                }
            }
            seg_run += chunk_total;
        }
    }
    }

    // TotalPixels, context.go:229: every covered in-range pixel, before any depth test
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) covered += __shfl_down_sync(0xffffffffu, covered, o);
    if (lane == 0 && covered) atomicAdd(&wb.counters->total_pixels, covered);

    // ---- publish: one contiguous run from the front of the region ---------------------------------------------
    if (tid == 0) {
        static_assert(FT / 32 == 4, "blk_wcnt / blk_woff hold one entry per warp of a 128-thread block");
        wb.blk_wcnt[vb] = make_uint4(seg_run, 0u, 0u, 0u);
        wb.blk_woff[vb] = make_uint4(0u, 0u, 0u, 0u);
        wb.blk_region[vb] = (uint32_t)min(my_region, 0xffffffffull);
        if (seg_run) atomicAdd(&wb.blk_base[vb / FRONT_GROUP], (unsigned long long)seg_run);
        if (nrec_blk) atomicAdd(&wb.counters->n_records, nrec_blk);
    }
}

// Segments in primitive order for the stable sort: position (segments of all earlier blocks) + (segments of the
// block's earlier warps) + k  <-  slot blk_region[b] + blk_woff[b][w] + k.  A CTA takes a contiguous range of
// front-end blocks.  It first adds up what lies before its range -- the group sums k_front accumulated (one per
// FRONT_GROUP blocks) plus the per-warp counts of the blocks between the group boundary and its range -- so k_front
// needs neither a last-block scan nor a block counter; then one warp per (block, warp run) copies the run with
// coalesced reads and writes, four independent loads in flight per lane.  CTA 0 also publishes the draw's totals
// (segments, buffer needs, overflow) for the kernels behind it.
constexpr int SI_THREADS = 256;
__global__ void __launch_bounds__(SI_THREADS)
k_seg_index(const __grid_constant__ WorkBuffers wb, uint32_t nent) {
    __shared__ unsigned long long s_part[SI_THREADS / 32];
    __shared__ uint32_t s_wtot[SI_THREADS / 32];
    __shared__ uint32_t s_first[SI_THREADS];  // ordered position of the first segment of each block of the range
    pdl_wait();
    pdl_trigger();
    DrawCounters *ctr = wb.counters;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t ngroups = (nent + FRONT_GROUP - 1) / FRONT_GROUP;
    // every block has made its reservation: the cursor is the sum of the upper bounds
    const unsigned long long cells_total = ctr->seg_cursor;
    const unsigned nclip = ctr->n_clip;
    unsigned ovf = ctr->overflow;  // (k_front sets OVF_CLIP itself when the pool runs out)
    if (cells_total > (unsigned long long)wb.cap_segs) ovf |= OVF_SEGS;
    if (nclip > wb.cap_clip) ovf |= OVF_CLIP;
    if (blockIdx.x == 0) {
        // totals: all group sums
        unsigned long long tot = 0;
        for (uint32_t g = tid; g < ngroups; g += SI_THREADS) tot += wb.blk_base[g];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
        if (lane == 0) s_part[warp] = tot;
        __syncthreads();
        if (tid == 0) {
            tot = 0;
            for (int w = 0; w < SI_THREADS / 32; w++) tot += s_part[w];
            wb.tile_ctl->nheavy = 0; wb.tile_ctl->nlight = 0; wb.tile_ctl->head = 0;  // the busy-strip list of this draw
            ctr->need_records = 0; ctr->n_rows = 0; ctr->need_rows = 0;
            ctr->n_segs = (unsigned int)min(tot, 0xfffffff0ull);
            ctr->need_segs = (unsigned int)min(cells_total, 0xfffffff0ull);
            ctr->need_clip = nclip;
            if (ovf) atomicOr(&ctr->overflow, ovf);
        }
        __syncthreads();
    }
    if (ovf) return;
    const uint32_t per = (nent + gridDim.x - 1) / gridDim.x;  // <= SI_THREADS (launch_seg_index)
    const uint32_t b0 = min(blockIdx.x * per, nent), b1 = min(b0 + per, nent);
    if (b0 >= b1) return;
    // segments before block b0
    const uint32_t g0 = b0 / FRONT_GROUP;
    unsigned long long before = 0;
    for (uint32_t g = tid; g < g0; g += SI_THREADS) before += wb.blk_base[g];
    for (uint32_t b = g0 * FRONT_GROUP + tid; b < b0; b += SI_THREADS) {
        const uint4 wc = wb.blk_wcnt[b];
        before += (unsigned long long)wc.x + wc.y + wc.z + wc.w;
    }
    // ... and of each block of the range (per <= SI_THREADS: one block per thread, a block scan)
    uint32_t mine = 0;
    if (b0 + tid < b1) {
        const uint4 wc = wb.blk_wcnt[b0 + tid];
        mine = wc.x + wc.y + wc.z + wc.w;
    }
    const uint32_t incl = warp_incl_scan(mine);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) before += __shfl_down_sync(0xffffffffu, before, o);
    if (lane == 0) s_part[warp] = before;
    if (lane == 31) s_wtot[warp] = incl;
    __syncthreads();
    unsigned long long base = 0;
    uint32_t wprev = 0;
#pragma unroll
    for (int w = 0; w < SI_THREADS / 32; w++) {
        base += s_part[w];
        if (w < (int)warp) wprev += s_wtot[w];
    }
    if (tid < per) s_first[tid] = (uint32_t)min(base + wprev + incl - mine, 0xffffffffull);
    __syncthreads();
    const uint32_t n = wb.cap_segs;  // (without overflow every position is below the capacity: segments <= cells <= cap_segs)
    const uint32_t ntasks = (b1 - b0) * 4u;
    for (uint32_t t = warp; t < ntasks; t += SI_THREADS / 32) {
        const uint32_t k_blk = t >> 2, b = b0 + k_blk, w = t & 3u;
        const uint4 wc = wb.blk_wcnt[b];
        const uint32_t cnt = w == 0 ? wc.x : (w == 1 ? wc.y : (w == 2 ? wc.z : wc.w));
        if (cnt == 0) continue;
        const uint4 wo = wb.blk_woff[b];
        const uint32_t before_w = w == 0 ? 0u : (w == 1 ? wc.x : (w == 2 ? wc.x + wc.y : wc.x + wc.y + wc.z));
        const uint32_t pos0 = s_first[k_blk] + before_w;
        const uint32_t slot0 = wb.blk_region[b] + (w == 0 ? wo.x : (w == 1 ? wo.y : (w == 2 ? wo.z : wo.w)));
        for (uint32_t k0 = 0; k0 < cnt; k0 += 128u) {
            uint32_t key[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + 32u * u + lane;
                key[u] = (k < cnt && pos0 + k < n) ? wb.seg_key[1][slot0 + k] : 0u;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const uint32_t k = k0 + 32u * u + lane;
                if (k < cnt && pos0 + k < n) {
                    wb.seg_key[0][pos0 + k] = key[u];
                    wb.seg_val[0][pos0 + k] = slot0 + k;
                }
            }
        }
    }
}

int launch_front(const DrawParams &p, const WorkBuffers &wb, bool counters_clean, cudaStream_t st) {
    const uint32_t blocks = (p.count + FT - 1) / FT;
    if (!counters_clean) cudaMemsetAsync(wb.counters, 0, sizeof(DrawCounters), st);
    k_front<<<blocks, FT, 0, st>>>(p, wb);
    return 1;
}
int launch_seg_index(const DrawParams &p, const WorkBuffers &wb, cudaStream_t st) {
    const uint32_t blocks = (p.count + FT - 1) / FT;
    // a CTA scans its range of front-end blocks with one block per thread: at least blocks / SI_THREADS CTAs
    const uint32_t grid = max(min(blocks, wb.nsm * 8u), (blocks + SI_THREADS - 1) / SI_THREADS);
    launch_pdl(k_seg_index, grid, SI_THREADS, 0, st, wb, blocks);
    return 1;
}

int launch_geometry(const DrawParams &p, const WorkBuffers &wb, bool counters_clean, cudaStream_t st) {
    const uint32_t blocks = (p.count + GT - 1) / GT;
    // counters + look-back state of this draw (stream-ordered before the kernel)
    if (!counters_clean) cudaMemsetAsync(wb.counters, 0, sizeof(DrawCounters), st);
    k_geometry<<<blocks, GT, 0, st>>>(p, wb);
    k_rec_index<<<blocks < GRID_WAVE ? blocks : GRID_WAVE, GT, 0, st>>>(wb, blocks);
    return 2;
}

}  // namespace fgl
