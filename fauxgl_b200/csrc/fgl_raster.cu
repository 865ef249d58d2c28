// fgl_raster.cu -- tile-parallel back end: one CTA per 64x16 screen tile keeps
// the tile's depth (f64) and colour (NRGBA8) in shared memory, consumes its bin
// in primitive order and writes the tile back once, coalesced.
//
// Replaces Context.rasterize (context.go:151-281), InterpolateVertexes
// (vertex.go:18-47), the three built-in Fragment shaders (shader.go:25,44,75),
// ImageTexture.BilinearSample (texture.go:41-63), Color.NRGBA (color.go:56) and
// the mutex-guarded depth retest / write / blend (context.go:245-273).
//
// Arithmetic parity.  The reference evaluates the edge functions by forward
// differencing: per row `w00 += b12`, a skip-ahead `d`, then `w0 += a12` per
// pixel (context.go:184-213, 275).  Those chains of float64 adds are
// reproduced literally: work item = one (triangle, scanline) span, walked left
// to right by one thread with the same adds in the same order, so coverage,
// barycentrics, depth and colour come out bit-identical to a sequential run of
// the reference in triangle-index order.
//
// Ordering.  Spans of a batch are processed in parallel; pixels touched by
// several spans of one batch are resolved in rounds: every pending fragment
// bids its span index with a shared-memory atomicMin on a per-pixel ticket, the
// lowest index wins, runs the reference's depth test / shade / retest / write
// on the tile copy, and retires.  That is exactly index order per pixel, which
// makes `<=` ties, DepthBias, blending and UpdatedPixels well defined.
#include "fgl_internal.h"
#include "fgl_block.cuh"
#include "fgl_math.cuh"

namespace fgl {

constexpr int RT = 256;      // threads per CTA
constexpr int CHUNK = 128;   // triangles set up per chunk
constexpr uint32_t NO_TICKET = 0xffffffffu;

struct TriSetup {
    double a12, a20, a01, b12, b20, b01;  // context.go:167-172
    double w00, w01, w02;                 // edge values at (x0+.5, yfirst+.5)
    double ra, ra12, ra20, ra01;          // context.go:175,179-181
    double r0, r1, r2;                    // 1/Output.W, context.go:176-178
    double z0, z1, z2;                    // s0.z s1.z s2.z
    int32_t x0, x1;                       // bbox columns
    int32_t yfirst, nrows;                // rows of the bbox inside this tile
    uint32_t src, flags;
};

FGL_DI double edge_fn(double ax, double ay, double bx, double by, double cx, double cy) {  // context.go:147
    return (bx - cx) * (ay - cy) - (by - cy) * (ax - cx);
}

FGL_DI double interp1(double a, double b, double c, double bx, double by, double bz, double bw) {  // vertex.go:49-79
    double n = 0;
    n = n + a * bx;
    n = n + b * by;
    n = n + c * bz;
    return n * bw;
}

// ---- attribute fetch: mesh planes or clip pool ---------------------------------------------
struct AttrSrc {
    const DrawParams *p;
    const ClipTri *pool;
    uint32_t src, flags;
    FGL_DI uint32_t vsrc(int k) const { return (flags >> (2 * k)) & 3u; }
    FGL_DI double pos(int k, int c) const {
        if (flags & REC_SRC_POOL) return pool[src].v[vsrc(k)].pos[c];
        return __ldg(p->mesh.pos + (size_t)(vsrc(k) * 3 + c) * p->mesh.n + src);
    }
    FGL_DI double nrm(int k, int c) const {
        if (flags & REC_SRC_POOL) return pool[src].v[vsrc(k)].nrm[c];
        return __ldg(p->mesh.nrm + (size_t)(vsrc(k) * 3 + c) * p->mesh.n + src);
    }
    FGL_DI double tex(int k, int c) const {
        if (flags & REC_SRC_POOL) return pool[src].v[vsrc(k)].tex[c];
        return __ldg(p->mesh.tex + (size_t)(vsrc(k) * 2 + c) * p->mesh.n + src);
    }
    FGL_DI double col(int k, int c) const {
        if (flags & REC_SRC_POOL) return pool[src].v[vsrc(k)].col[c];
        return __ldg(p->mesh.col + (size_t)(vsrc(k) * 4 + c) * p->mesh.n + src);
    }
};

// ---- texture.go -----------------------------------------------------------------------------
FGL_DI C4 tex_at(const DrawParams &p, long long x, long long y) {  // image.At + MakeColor (color.go:25-29)
    if (x < 0 || y < 0 || x >= p.tex_w || y >= p.tex_h) return c4(0, 0, 0, 0);
    const uint32_t t = __ldg(reinterpret_cast<const uint32_t *>(p.tex) + (size_t)y * p.tex_w + x);
    uint32_t r = t & 0xff, g = (t >> 8) & 0xff, b = (t >> 16) & 0xff, a = t >> 24;
    if (p.tex_format == FGL_TEX_NRGBA) {  // color.NRGBA.RGBA(): premultiply
        r |= r << 8; r *= a; r /= 0xff;
        g |= g << 8; g *= a; g /= 0xff;
        b |= b << 8; b *= a; b /= 0xff;
        a |= a << 8;
    } else {  // color.RGBA.RGBA()
        r |= r << 8; g |= g << 8; b |= b << 8; a |= a << 8;
    }
    const double d = 65535.0;
    return c4((double)r / d, (double)g / d, (double)b / d, (double)a / d);
}
__device__ __noinline__ C4 bilinear_sample(const DrawParams &p, double u, double v) {  // texture.go:41-63
    v = 1 - v;
    u -= floor(u);
    v -= floor(v);
    double x = u * (double)(p.tex_w - 1);
    double y = v * (double)(p.tex_h - 1);
    const long long x0 = go_int(x), y0 = go_int(y);
    const long long x1 = x0 + 1, y1 = y0 + 1;
    x -= (double)x0;
    y -= (double)y0;
    const C4 c00 = tex_at(p, x0, y0), c01 = tex_at(p, x0, y1), c10 = tex_at(p, x1, y0), c11 = tex_at(p, x1, y1);
    C4 c = c4(0, 0, 0, 0);
    c = c_add(c, c_muls(c00, (1 - x) * (1 - y)));
    c = c_add(c, c_muls(c10, x * (1 - y)));
    c = c_add(c, c_muls(c01, (1 - x) * y));
    c = c_add(c, c_muls(c11, x * y));
    return c;
}

// ---- fragment: interpolation (vertex.go:18-47) + Shader.Fragment (shader.go) -----------------------
// (bx,by,bz,bw) are the perspective-corrected weights of context.go:236-237.
__device__ __noinline__ C4 shade_fragment(const DrawParams &p, const AttrSrc &a, double bx, double by, double bz,
                                          double bw) {
    if (p.kind == FGL_SHADER_SOLID) return c4(p.color[0], p.color[1], p.color[2], p.color[3]);
    if (p.kind == FGL_SHADER_TEXTURE) {
        const double tu = interp1(a.tex(0, 0), a.tex(1, 0), a.tex(2, 0), bx, by, bz, bw);
        const double tv = interp1(a.tex(0, 1), a.tex(1, 1), a.tex(2, 1), bx, by, bz, bw);
        return bilinear_sample(p, tu, tv);
    }
    // PhongShader.Fragment, shader.go:75-96
    C4 light = c4(p.ambient[0], p.ambient[1], p.ambient[2], p.ambient[3]);
    C4 color;
    if (p.has_texture) {
        const double tu = interp1(a.tex(0, 0), a.tex(1, 0), a.tex(2, 0), bx, by, bz, bw);
        const double tv = interp1(a.tex(0, 1), a.tex(1, 1), a.tex(2, 1), bx, by, bz, bw);
        color = bilinear_sample(p, tu, tv);
    } else if (!p.object_is_discard) {
        color = c4(p.object[0], p.object[1], p.object[2], p.object[3]);
    } else {
        color = c4(interp1(a.col(0, 0), a.col(1, 0), a.col(2, 0), bx, by, bz, bw),
                   interp1(a.col(0, 1), a.col(1, 1), a.col(2, 1), bx, by, bz, bw),
                   interp1(a.col(0, 2), a.col(1, 2), a.col(2, 2), bx, by, bz, bw),
                   interp1(a.col(0, 3), a.col(1, 3), a.col(2, 3), bx, by, bz, bw));
    }
    const V3 normal = v_normalize(v3(interp1(a.nrm(0, 0), a.nrm(1, 0), a.nrm(2, 0), bx, by, bz, bw),
                                     interp1(a.nrm(0, 1), a.nrm(1, 1), a.nrm(2, 1), bx, by, bz, bw),
                                     interp1(a.nrm(0, 2), a.nrm(1, 2), a.nrm(2, 2), bx, by, bz, bw)));
    const V3 ld = v3(p.light[0], p.light[1], p.light[2]);
    const double diffuse = go_max(v_dot(normal, ld), 0);
    light = c_add(light, c_muls(c4(p.diffuse[0], p.diffuse[1], p.diffuse[2], p.diffuse[3]), diffuse));
    if (diffuse > 0 && p.specular_power > 0) {
        const V3 position = v3(interp1(a.pos(0, 0), a.pos(1, 0), a.pos(2, 0), bx, by, bz, bw),
                               interp1(a.pos(0, 1), a.pos(1, 1), a.pos(2, 1), bx, by, bz, bw),
                               interp1(a.pos(0, 2), a.pos(1, 2), a.pos(2, 2), bx, by, bz, bw));
        const V3 camera = v_normalize(v_sub(v3(p.camera[0], p.camera[1], p.camera[2]), position));
        const V3 reflected = v_reflect(v_negate(ld), normal);
        double specular = go_max(v_dot(camera, reflected), 0);
        if (specular > 0) {
            specular = go_pow(specular, p.specular_power);
            light = c_add(light, c_muls(c4(p.specular[0], p.specular[1], p.specular[2], p.specular[3]), specular));
        }
    }
    C4 r = c_mul(color, light);
    r = c4(go_min(r.r, 1), go_min(r.g, 1), go_min(r.b, 1), go_min(r.a, 1));
    r.a = color.a;
    return r;
}

// Alpha blend, context.go:256-267 (Go stdlib color.NRGBA.RGBA(), u32 arithmetic).
FGL_DI uint32_t blend_over(uint32_t dst, uint32_t c8) {
    const uint32_t A8 = c8 >> 24;
    uint32_t sa = A8; sa |= sa << 8;
    uint32_t sr = c8 & 0xff; sr |= sr << 8; sr *= A8; sr /= 0xff;
    uint32_t sg = (c8 >> 8) & 0xff; sg |= sg << 8; sg *= A8; sg /= 0xff;
    uint32_t sb = (c8 >> 16) & 0xff; sb |= sb << 8; sb *= A8; sb /= 0xff;
    const uint32_t a = (0xffffu - sa) * 0x101u;
    const uint32_t dr = (((dst & 0xff) * a / 0xffffu + sr) >> 8) & 0xff;
    const uint32_t dg = ((((dst >> 8) & 0xff) * a / 0xffffu + sg) >> 8) & 0xff;
    const uint32_t db = ((((dst >> 16) & 0xff) * a / 0xffffu + sb) >> 8) & 0xff;
    const uint32_t da = (((dst >> 24) * a / 0xffffu + sa) >> 8) & 0xff;
    return dr | (dg << 8) | (db << 16) | (da << 24);
}

__global__ void __launch_bounds__(RT)
k_raster(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb,
         const uint32_t *__restrict__ pair_val, uint32_t *__restrict__ gcolor, double *__restrict__ gdepth) {
    const uint32_t tile = blockIdx.x;
    const uint32_t bin_beg = wb.tile_start[tile], bin_end = wb.tile_end[tile];
    if (bin_beg >= bin_end) return;
    if (wb.counters->overflow) return;  // work buffers too small: the host regrows and re-issues the draw

    __shared__ double s_depth[TILE_PIX];
    __shared__ uint32_t s_color[TILE_PIX];
    __shared__ uint32_t s_ticket[TILE_PIX];
    __shared__ TriSetup s_tri[CHUNK];
    __shared__ uint32_t s_span_off[CHUNK + 1];
    __shared__ uint32_t s_scan[RT / 32 + 1];
    __shared__ unsigned long long s_info[2];

    const int tid = threadIdx.x;
    const int tile_x0 = (int)(tile % (uint32_t)p.tiles_x) * TILE_W;
    const int tile_y0 = (int)(tile / (uint32_t)p.tiles_x) * TILE_H;
    const int tw = min(TILE_W, p.width - tile_x0);   // valid columns of this tile
    const int th = min(TILE_H, p.height - tile_y0);  // valid rows
    const int tile_x1 = tile_x0 + tw - 1, tile_y1 = tile_y0 + th - 1;

    // ---- load the tile ------------------------------------------------------------------------
    for (int i = tid; i < TILE_PIX; i += RT) {
        const int lx = i % TILE_W, ly = i / TILE_W;
        s_ticket[i] = NO_TICKET;
        if (lx < tw && ly < th) {
            const size_t g = (size_t)(tile_y0 + ly) * p.width + (tile_x0 + lx);
            s_depth[i] = gdepth[g];
            s_color[i] = gcolor[g];
        }
    }
    if (tid < 2) s_info[tid] = 0;
    __syncthreads();

    const fgl_state st = p.state;
    unsigned long long my_total = 0, my_updated = 0;

    for (uint32_t chunk = bin_beg; chunk < bin_end; chunk += CHUNK) {
        const uint32_t ntri = min((uint32_t)CHUNK, bin_end - chunk);

        // ---- per-triangle setup: context.go:163-181, rows clipped to the tile ------------------
        uint32_t my_rows = 0;
        if (tid < (int)ntri) {
            const Rec *rp = wb.recs + pair_val[chunk + tid];
            const double s0x = rp->s[0], s0y = rp->s[1], s0z = rp->s[2];
            const double s1x = rp->s[3], s1y = rp->s[4], s1z = rp->s[5];
            const double s2x = rp->s[6], s2y = rp->s[7], s2z = rp->s[8];
            const int x0 = rp->x0, x1 = rp->x1, y0 = rp->y0, y1 = rp->y1;
            TriSetup t;
            const double px = (double)x0 + 0.5, py = (double)y0 + 0.5;
            double w00 = edge_fn(s1x, s1y, s2x, s2y, px, py);
            double w01 = edge_fn(s2x, s2y, s0x, s0y, px, py);
            double w02 = edge_fn(s0x, s0y, s1x, s1y, px, py);
            t.a01 = s1y - s0y; t.b01 = s0x - s1x;
            t.a12 = s2y - s1y; t.b12 = s1x - s2x;
            t.a20 = s0y - s2y; t.b20 = s2x - s0x;
            t.ra = 1 / edge_fn(s0x, s0y, s1x, s1y, s2x, s2y);
            t.r0 = 1 / rp->w[0]; t.r1 = 1 / rp->w[1]; t.r2 = 1 / rp->w[2];
            t.ra12 = 1 / t.a12; t.ra20 = 1 / t.a20; t.ra01 = 1 / t.a01;
            t.z0 = s0z; t.z1 = s1z; t.z2 = s2z;
            const int ylo = max(y0, tile_y0), yhi = min(y1, tile_y1);
            t.nrows = max(0, yhi - ylo + 1);
            t.yfirst = ylo;
            // the reference adds b12/b20/b01 once per row from y0 (context.go:275): replay it
            if (t.nrows > 0)
                for (int y = y0; y < ylo; y++) { w00 += t.b12; w01 += t.b20; w02 += t.b01; }
            t.w00 = w00; t.w01 = w01; t.w02 = w02;
            t.x0 = x0; t.x1 = x1;
            t.src = rp->src; t.flags = rp->flags;
            s_tri[tid] = t;
            my_rows = (uint32_t)t.nrows;
        }
        uint32_t nspans;
        const uint32_t off = block_excl_scan<RT>(my_rows, s_scan, &nspans);
        if (tid <= (int)ntri) s_span_off[tid] = (tid < (int)ntri) ? off : nspans;
        __syncthreads();

        // ---- spans, RT at a time, in (triangle, row) order -----------------------------------------
        for (uint32_t sbase = 0; sbase < nspans; sbase += RT) {
            const uint32_t sid = sbase + tid;
            const bool have = sid < nspans;
            // span state
            int xa = 0, cnt = 0, y = 0;
            double wa0 = 0, wa1 = 0, wa2 = 0;
            uint32_t tj = 0;
            if (have) {
                uint32_t lo = 0, hi = ntri;  // s_span_off[lo] <= sid < s_span_off[hi]
                while (hi - lo > 1) {
                    const uint32_t mid = (lo + hi) >> 1;
                    if (s_span_off[mid] <= sid) lo = mid; else hi = mid;
                }
                tj = lo;
                const TriSetup &t = s_tri[tj];
                const int row = (int)(sid - s_span_off[tj]);
                y = t.yfirst + row;
                double w00 = t.w00, w01 = t.w01, w02 = t.w02;
                for (int k = 0; k < row; k++) { w00 += t.b12; w01 += t.b20; w02 += t.b01; }
                // skip-ahead, context.go:185-205
                double d = 0;
                const double d0 = -w00 * t.ra12, d1 = -w01 * t.ra20, d2 = -w02 * t.ra01;
                if (w00 < 0 && d0 > d) d = d0;
                if (w01 < 0 && d1 > d) d = d1;
                if (w02 < 0 && d2 > d) d = d2;
                d = (double)go_int(d);
                if (d < 0) d = 0;
                double w0 = w00 + t.a12 * d, w1 = w01 + t.a20 * d, w2 = w02 + t.a01 * d;
                long long x = (long long)t.x0 + go_int(d);
                const long long xend = min((long long)t.x1, (long long)tile_x1);
                if (x <= xend) {
                    // replay the per-pixel adds up to the tile's first column (context.go:211-213)
                    for (; x < tile_x0; x++) { w0 += t.a12; w1 += t.a20; w2 += t.a01; }
                    for (; x <= xend; x++) {
                        const double b0 = w0 * t.ra, b1 = w1 * t.ra, b2 = w2 * t.ra;
                        if (b0 < 0 || b1 < 0 || b2 < 0) {
                            if (cnt > 0) break;  // wasInside, context.go:216-218
                        } else {
                            if (cnt == 0) { xa = (int)x; wa0 = w0; wa1 = w1; wa2 = w2; }
                            cnt++;
                        }
                        w0 += t.a12; w1 += t.a20; w2 += t.a01;
                    }
                }
                my_total += (unsigned long long)cnt;  // context.go:229
            }
            unsigned long long pend = cnt > 0 ? ((cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull)) << (xa - tile_x0)) : 0ull;
            const int rowbase = (y - tile_y0) * TILE_W - tile_x0;

            // ---- ordered resolution in rounds ---------------------------------------------------
            while (true) {
                if (pend) {
                    unsigned long long m = pend;
                    while (m) {
                        const int bit = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        atomicMin(&s_ticket[(y - tile_y0) * TILE_W + bit], (uint32_t)tid);
                    }
                }
                __syncthreads();
                if (pend) {
                    const TriSetup &t = s_tri[tj];
                    double w0 = wa0, w1 = wa1, w2 = wa2;
                    for (int x = xa; x < xa + cnt; x++) {
                        const unsigned long long bitm = 1ull << (x - tile_x0);
                        const int pi = rowbase + x;
                        if ((pend & bitm) && s_ticket[pi] == (uint32_t)tid) {
                            pend &= ~bitm;
                            s_ticket[pi] = NO_TICKET;
                            const double b0 = w0 * t.ra, b1 = w1 * t.ra, b2 = w2 * t.ra;
                            const double z = b0 * t.z0 + b1 * t.z1 + b2 * t.z2;     // context.go:230
                            const double bz = z + st.depth_bias;
                            const double dcur = s_depth[pi];
                            if (!(st.read_depth && bz > dcur)) {                    // context.go:232
                                const double bx = b0 * t.r0, by = b1 * t.r1, bzz = b2 * t.r2;  // :236
                                const double bw = 1 / (bx + by + bzz);
                                AttrSrc a{&p, wb.clip_pool, t.src, t.flags};
                                const C4 color = shade_fragment(p, a, bx, by, bzz, bw);
                                if (!c_is_discard(color)) {                         // context.go:241
                                    if (bz <= dcur || !st.read_depth) {             // context.go:248
                                        my_updated++;
                                        if (st.write_depth) s_depth[pi] = z;
                                        if (st.write_color) {
                                            const uint32_t c8 = c_nrgba(color);
                                            if (st.alpha_blend && color.a < 1) s_color[pi] = blend_over(s_color[pi], c8);
                                            else s_color[pi] = c8;
                                        }
                                    }
                                }
                            }
                        }
                        w0 += t.a12; w1 += t.a20; w2 += t.a01;
                    }
                }
                if (!__syncthreads_or(pend != 0)) break;
            }
        }
        __syncthreads();  // s_tri / s_span_off are rewritten by the next chunk
    }

    // ---- write the tile back, RasterizeInfo -------------------------------------------------------
    for (int i = tid; i < TILE_PIX; i += RT) {
        const int lx = i % TILE_W, ly = i / TILE_W;
        if (lx < tw && ly < th) {
            const size_t g = (size_t)(tile_y0 + ly) * p.width + (tile_x0 + lx);
            gdepth[g] = s_depth[i];
            gcolor[g] = s_color[i];
        }
    }
    // warp-reduce the counters, then one atomic per warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        my_total += __shfl_down_sync(0xffffffffu, my_total, o);
        my_updated += __shfl_down_sync(0xffffffffu, my_updated, o);
    }
    if ((tid & 31) == 0) {
        atomicAdd(&s_info[0], my_total);
        atomicAdd(&s_info[1], my_updated);
    }
    __syncthreads();
    if (tid == 0) {
        if (s_info[0]) atomicAdd(&wb.counters->total_pixels, s_info[0]);
        if (s_info[1]) atomicAdd(&wb.counters->updated_pixels, s_info[1]);
    }
}

int launch_raster(const DrawParams &p, const WorkBuffers &wb, int sorted_buf, uint32_t *color, double *depth,
                  cudaStream_t st) {
    k_raster<<<wb.ntiles, RT, 0, st>>>(p, wb, wb.pair_val[sorted_buf], color, depth);
    return 1;
}

}  // namespace fgl
