// fgl_raster.cu -- tile-parallel, ordered back end: one CTA per 64x16 screen tile
// keeps the tile's depth (f64) and colour (NRGBA8) in shared memory, consumes its
// bin of span segments in primitive order and writes the tile back once,
// coalesced -- replacing the reference's per-pixel mutex array.
//
// Replaces the per-pixel body of Context.rasterize (context.go:207-273),
// InterpolateVertexes (vertex.go:18-47), the three built-in Fragment shaders
// (shader.go:25,44,75), ImageTexture.BilinearSample (texture.go:41-63),
// Color.NRGBA (color.go:56) and the depth retest / write / blend (context.go:245-273).
//
// Arithmetic parity.  Each segment carries the reference's forward-differenced
// edge values at its first pixel (fgl_span.cu); this kernel continues the same
// `w += a` chain (context.go:211-213), so barycentrics, depth and colour are
// bit-identical to a sequential run of the reference in triangle-index order.
//
// Ordering.  RT segments are resolved in parallel; pixels touched by several
// segments of a batch are settled in rounds: every pending fragment bids its
// batch index with a shared-memory atomicMin on a per-pixel ticket, the lowest
// wins, applies the reference's depth test / retest / write to the tile copy
// and retires.  That is index order per pixel, which makes `<=` ties,
// DepthBias, blending and UpdatedPixels well defined.
//
// Deferred shading.  When the draw's shader can neither discard nor blend
// (SolidColor, or Phong with an ObjectColor and no texture, alpha != 0 and no
// effective blending) the fragment colour cannot influence any depth decision,
// so the ordered phase resolves depth only and records, per pixel, the winning
// record and its edge values; the colour of the final winners is computed once
// per pixel at the end of the tile, with every thread busy.  Otherwise
// fragments are shaded inline, in order, exactly like the reference.
#include "fgl_internal.h"
#include "fgl_block.cuh"
#include "fgl_math.cuh"

namespace fgl {

constexpr int RT = 256;      // threads per CTA == segments per batch
constexpr uint32_t NO_TICKET = 0xffffffffu;
constexpr uint32_t NO_WINNER = 0xffffffffu;

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

struct SegVis {   // what the ordered phase needs from the record
    double ra, z0, z1, z2, a12, a20, a01;
};

// One fragment, inline mode: context.go:229-273 for a pixel whose ticket this thread holds.
FGL_DI void fragment_inline(const DrawParams &p, const fgl_state &st, const WorkBuffers &wb, const Rec &r,
                            double w0, double w1, double w2, int pi, double *s_depth, uint32_t *s_color,
                            unsigned long long &updated) {
    const double b0 = w0 * r.ra, b1 = w1 * r.ra, b2 = w2 * r.ra;
    const double z = b0 * r.s[2] + b1 * r.s[5] + b2 * r.s[8];  // context.go:230
    const double bz = z + st.depth_bias;
    const double dcur = s_depth[pi];
    if (st.read_depth && bz > dcur) return;                     // context.go:232
    const double bx = b0 * r.r0, by = b1 * r.r1, bzz = b2 * r.r2;  // context.go:236
    const double bw = 1 / (bx + by + bzz);
    AttrSrc a{&p, wb.clip_pool, r.src, r.flags};
    const C4 color = shade_fragment(p, a, bx, by, bzz, bw);
    if (c_is_discard(color)) return;                            // context.go:241
    if (bz <= dcur || !st.read_depth) {                         // context.go:248
        updated++;
        if (st.write_depth) s_depth[pi] = z;
        if (st.write_color) {
            const uint32_t c8 = c_nrgba(color);
            if (st.alpha_blend && color.a < 1) s_color[pi] = blend_over(s_color[pi], c8);
            else s_color[pi] = c8;
        }
    }
}

template <bool DEFERRED>
__global__ void __launch_bounds__(RT)
k_tile(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb,
       const uint32_t *__restrict__ seg_order, uint32_t *__restrict__ gcolor, double *__restrict__ gdepth) {
    const uint32_t tile = blockIdx.x;
    const uint32_t bin_beg = wb.tile_start[tile], bin_end = wb.tile_end[tile];
    if (bin_beg >= bin_end) return;
    if (wb.counters->overflow) return;  // work buffers too small: the host regrows and re-issues the draw

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_depth = reinterpret_cast<double *>(smem_raw);                 // [TILE_PIX]
    double *s_w = s_depth + TILE_PIX;                                        // [3][TILE_PIX]   (deferred only)
    uint32_t *s_color = reinterpret_cast<uint32_t *>(s_w + (DEFERRED ? 3 * TILE_PIX : 0));  // [TILE_PIX]
    uint32_t *s_ticket = s_color + TILE_PIX;                                 // [TILE_PIX]
    uint32_t *s_winner = s_ticket + TILE_PIX;                                // [TILE_PIX]      (deferred only)
    __shared__ unsigned long long s_updated;

    const int tid = threadIdx.x;
    const int tile_x0 = (int)(tile % (uint32_t)p.tiles_x) * TILE_W;
    const int tile_y0 = (int)(tile / (uint32_t)p.tiles_x) * TILE_H;
    const int tw = min(TILE_W, p.width - tile_x0);   // valid columns of this tile
    const int th = min(TILE_H, p.height - tile_y0);  // valid rows

    // ---- load the tile ------------------------------------------------------------------------
    for (int i = tid; i < TILE_PIX; i += RT) {
        const int lx = i % TILE_W, ly = i / TILE_W;
        s_ticket[i] = NO_TICKET;
        if (DEFERRED) s_winner[i] = NO_WINNER;
        if (lx < tw && ly < th) {
            const size_t g = (size_t)(tile_y0 + ly) * p.width + (tile_x0 + lx);
            s_depth[i] = gdepth[g];
            s_color[i] = gcolor[g];
        }
    }
    if (tid == 0) s_updated = 0;
    __syncthreads();

    const fgl_state st = p.state;
    unsigned long long my_updated = 0;

    for (uint32_t batch = bin_beg; batch < bin_end; batch += RT) {
        const bool have = batch + tid < bin_end;
        Seg sg;
        sg.cnt = 0; sg.x = 0; sg.yt = 0; sg.rec = 0; sg.w0 = sg.w1 = sg.w2 = 0;
        SegVis v;
        v.ra = v.z0 = v.z1 = v.z2 = v.a12 = v.a20 = v.a01 = 0;
        const Rec *rp = nullptr;
        if (have) {
            sg = wb.segs[seg_order[batch + tid]];
            rp = wb.recs + sg.rec;
            v.ra = rp->ra; v.z0 = rp->s[2]; v.z1 = rp->s[5]; v.z2 = rp->s[8];
            v.a01 = rp->s[4] - rp->s[1]; v.a12 = rp->s[7] - rp->s[4]; v.a20 = rp->s[1] - rp->s[7];
        }
        const int xa = (int)sg.x, cnt = (int)sg.cnt;
        const int rowbase = (int)sg.yt * TILE_W - tile_x0;
        unsigned long long pend = cnt > 0 ? ((cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull)) << (xa - tile_x0)) : 0ull;

        // ---- ordered resolution in rounds ---------------------------------------------------
        while (true) {
            if (pend) {
                unsigned long long m = pend;
                while (m) {
                    const int bit = __ffsll((long long)m) - 1;
                    m &= m - 1;
                    atomicMin(&s_ticket[(int)sg.yt * TILE_W + bit], (uint32_t)tid);
                }
            }
            __syncthreads();
            if (pend) {
                double w0 = sg.w0, w1 = sg.w1, w2 = sg.w2;
                for (int x = xa; x < xa + cnt; x++) {
                    const unsigned long long bitm = 1ull << (x - tile_x0);
                    const int pi = rowbase + x;
                    if ((pend & bitm) && s_ticket[pi] == (uint32_t)tid) {
                        pend &= ~bitm;
                        s_ticket[pi] = NO_TICKET;
                        if (DEFERRED) {
                            const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
                            const double z = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
                            const double bz = z + st.depth_bias;
                            const double dcur = s_depth[pi];
                            // context.go:232 early-out, then (no discard possible) the retest at :248
                            if (!(st.read_depth && bz > dcur) && (bz <= dcur || !st.read_depth)) {
                                my_updated++;
                                if (st.write_depth) s_depth[pi] = z;
                                if (st.write_color) {
                                    s_winner[pi] = sg.rec;
                                    s_w[pi] = w0; s_w[TILE_PIX + pi] = w1; s_w[2 * TILE_PIX + pi] = w2;
                                }
                            }
                        } else {
                            fragment_inline(p, st, wb, *rp, w0, w1, w2, pi, s_depth, s_color, my_updated);
                        }
                    }
                    w0 += v.a12; w1 += v.a20; w2 += v.a01;  // context.go:211-213
                }
            }
            if (!__syncthreads_or(pend != 0)) break;
        }
    }

    // ---- deferred shading of the tile's final winners ------------------------------------------
    if (DEFERRED) {
        __syncthreads();
        for (int pi = tid; pi < TILE_PIX; pi += RT) {
            const uint32_t rid = s_winner[pi];
            if (rid == NO_WINNER) continue;
            const Rec *rp = wb.recs + rid;
            const double ra = rp->ra;
            const double b0 = s_w[pi] * ra, b1 = s_w[TILE_PIX + pi] * ra, b2 = s_w[2 * TILE_PIX + pi] * ra;
            const double bx = b0 * rp->r0, by = b1 * rp->r1, bzz = b2 * rp->r2;  // context.go:236
            const double bw = 1 / (bx + by + bzz);
            AttrSrc a{&p, wb.clip_pool, rp->src, rp->flags};
            const C4 color = shade_fragment(p, a, bx, by, bzz, bw);
            s_color[pi] = c_nrgba(color);  // SetNRGBA, context.go:269 (blending excluded by the mode)
        }
        __syncthreads();
    }

    // ---- write the tile back, UpdatedPixels ---------------------------------------------------------
    for (int i = tid; i < TILE_PIX; i += RT) {
        const int lx = i % TILE_W, ly = i / TILE_W;
        if (lx < tw && ly < th) {
            const size_t g = (size_t)(tile_y0 + ly) * p.width + (tile_x0 + lx);
            gdepth[g] = s_depth[i];
            gcolor[g] = s_color[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_updated += __shfl_down_sync(0xffffffffu, my_updated, o);
    if ((tid & 31) == 0 && my_updated) atomicAdd(&s_updated, my_updated);
    __syncthreads();
    if (tid == 0 && s_updated) atomicAdd(&wb.counters->updated_pixels, s_updated);
}

static size_t tile_smem(bool deferred) {
    return sizeof(double) * TILE_PIX * (deferred ? 4 : 1) + sizeof(uint32_t) * TILE_PIX * (deferred ? 3 : 2);
}

int launch_raster(const DrawParams &p, const WorkBuffers &wb, int sorted_buf, uint32_t *color, double *depth,
                  cudaStream_t st) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(k_tile<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(true));
        cudaFuncSetAttribute(k_tile<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(false));
        configured = true;
    }
    if (p.deferred)
        k_tile<true><<<wb.ntiles, RT, tile_smem(true), st>>>(p, wb, wb.seg_val[sorted_buf], color, depth);
    else
        k_tile<false><<<wb.ntiles, RT, tile_smem(false), st>>>(p, wb, wb.seg_val[sorted_buf], color, depth);
    return 1;
}

}  // namespace fgl
