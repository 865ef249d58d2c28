// fgl_shade.cuh -- fragment shading in the reference's arithmetic: attribute
// interpolation (vertex.go:18-47), the three built-in Fragment shaders
// (shader.go:25,44,75), ImageTexture.BilinearSample (texture.go:41-63) and the
// alpha blend of context.go:256-267.  Included by fgl_raster.cu only.
#pragma once
#include "fgl_internal.h"
#include "fgl_math.cuh"

namespace fgl {

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
    uint32_t r, g, b, a;
    if (p.tex_format == FGL_TEX_RGBA64) {  // the 16-bit values of Color.RGBA() themselves
        const uint2 t = __ldg(reinterpret_cast<const uint2 *>(p.tex) + (size_t)y * p.tex_w + x);
        r = t.x & 0xffff; g = t.x >> 16; b = t.y & 0xffff; a = t.y >> 16;
        const double d = 65535.0;
        return c4((double)r / d, (double)g / d, (double)b / d, (double)a / d);
    }
    const uint32_t t = __ldg(reinterpret_cast<const uint32_t *>(p.tex) + (size_t)y * p.tex_w + x);
    r = t & 0xff; g = (t >> 8) & 0xff; b = (t >> 16) & 0xff; a = t >> 24;
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

}  // namespace fgl
