/*
 * fauxgl_oracle.c -- CPU restatement of fogleman/fauxgl's DrawMesh hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see fauxgl_oracle.h).  PARITY UNPINNED: the
 * reference has no tests/golden vectors and cannot be built here.
 *
 * Every function cites the reference file:line it restates.  All arithmetic is
 * IEEE float64, unfused (build with -ffp-contract=off), evaluated left to right
 * exactly as the Go source is written (Go/amd64 does not contract a*b+c).
 * Go stdlib pieces that the path depends on (math.Max/Min/Pow, float->int
 * conversion on amd64, image/color RGBA() conversions) are restated below and
 * flagged "Go stdlib".
 */
#include "fauxgl_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double x, y, z; } V3;
typedef struct { double x, y, z, w; } V4;
typedef struct { double r, g, b, a; } C4;
/* vertex.go:3-12; identical layout to overtex (17 doubles). */
typedef struct { V3 position, normal, texture; C4 color; V4 output; } Vertex;

/* ---- Go stdlib semantics ------------------------------------------------ */

/* Go stdlib math.Max (math/dim.go). */
static double go_max(double x, double y) {
    if ((isinf(x) && x > 0) || (isinf(y) && y > 0)) return INFINITY;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 && x == y) { if (signbit(x)) return y; return x; }
    if (x > y) return x;
    return y;
}
/* Go stdlib math.Min (math/dim.go). */
static double go_min(double x, double y) {
    if ((isinf(x) && x < 0) || (isinf(y) && y < 0)) return -INFINITY;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0 && x == y) { if (signbit(x)) return x; return y; }
    if (x < y) return x;
    return y;
}
/* Go float64 -> int on amd64 (CVTTSD2SQ): truncation; NaN and out-of-range
 * values give the "integer indefinite" 0x8000000000000000. */
static int64_t go_int(double x) {
    if (!(x > -9223372036854775808.0 && x < 9223372036854775808.0)) return INT64_MIN;
    return (int64_t)x;
}
static int go_is_odd_int(double x) { /* math/pow.go isOddInt */
    if (fabs(x) >= 9007199254740992.0) return 0;
    double xi, xf = modf(x, &xi);
    return xf == 0 && (((int64_t)xi) & 1) == 1;
}
/* Go stdlib math.Pow (math/pow.go, pure-Go path used on amd64). */
double oracle_pow(double x, double y) {
    if (y == 0 || x == 1) return 1;
    if (y == 1) return x;
    if (isnan(x) || isnan(y)) return NAN;
    if (x == 0) {
        if (y < 0) {
            if (signbit(x) && go_is_odd_int(y)) return copysign(INFINITY, x);
            return INFINITY;
        } else if (y > 0) {
            if (signbit(x) && go_is_odd_int(y)) return x;
            return 0;
        }
    }
    if (isinf(y)) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    if (isinf(x)) {
        if (x < 0) return oracle_pow(1 / x, -y);
        if (y < 0) return 0;
        if (y > 0) return INFINITY;
    }
    if (y == 0.5) return sqrt(x);
    if (y == -0.5) return 1 / sqrt(x);

    double yi, yf = modf(fabs(y), &yi);
    if (yf != 0 && x < 0) return NAN;
    if (yi >= 9223372036854775808.0) {
        if (x == -1) return 1;
        if ((fabs(x) < 1) == (y > 0)) return 0;
        return INFINITY;
    }
    double a1 = 1.0;
    int64_t ae = 0;
    if (yf != 0) {
        if (yf > 0.5) { yf -= 1; yi += 1; }
        /* Exp/Log: not bit-pinned against Go (libm here). Only reached for a
         * non-integer SpecularPower, which no reference example uses. */
        a1 = exp(yf * log(x));
    }
    int xe_i;
    double x1 = frexp(x, &xe_i);
    int64_t xe = xe_i;
    for (int64_t i = (int64_t)yi; i != 0; i >>= 1) {
        if (xe < -(1 << 12) || (1 << 12) < xe) { ae += xe; break; }
        if (i & 1) { a1 *= x1; ae += xe; }
        x1 *= x1;
        xe <<= 1;
        if (x1 < .5) { x1 += x1; xe--; }
    }
    if (y < 0) { a1 = 1 / a1; ae = -ae; }
    if (ae > 100000) ae = 100000;
    if (ae < -100000) ae = -100000;
    return ldexp(a1, (int)ae);
}

/* ---- vector.go ----------------------------------------------------------- */

static V3 v3(double x, double y, double z) { V3 r = {x, y, z}; return r; }
static V3 v_add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }        /* :96  */
static V3 v_sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }        /* :100 */
static V3 v_muls(V3 a, double b) { return v3(a.x * b, a.y * b, a.z * b); }         /* :128 */
static double v_dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }      /* :72  */
static V3 v_cross(V3 a, V3 b) {                                                    /* :76  */
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static V3 v_normalize(V3 a) {                                                      /* :83  */
    double r = 1 / sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return v3(a.x * r, a.y * r, a.z * r);
}
static V3 v_negate(V3 a) { return v3(-a.x, -a.y, -a.z); }                          /* :88  */
static V3 v_min(V3 a, V3 b) { return v3(go_min(a.x, b.x), go_min(a.y, b.y), go_min(a.z, b.z)); } /* :140 */
static V3 v_max(V3 a, V3 b) { return v3(go_max(a.x, b.x), go_max(a.y, b.y), go_max(a.z, b.z)); } /* :144 */
static V3 v_reflect(V3 i, V3 n) { return v_sub(i, v_muls(n, 2 * v_dot(n, i))); }   /* :175 */
static V3 v_perpendicular(V3 a) {                                                  /* :179 */
    if (a.x == 0 && a.y == 0) {
        if (a.z == 0) return v3(0, 0, 0);
        return v3(0, 1, 0);
    }
    return v_normalize(v3(-a.y, a.x, 0));
}
static int v_is_zero(V3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }

static V4 v4(double x, double y, double z, double w) { V4 r = {x, y, z, w}; return r; }
static int w_outside(V4 a) {                                                       /* :212 */
    double x = a.x, y = a.y, z = a.z, w = a.w;
    return x < -w || x > w || y < -w || y > w || z < -w || z > w;
}
static double w_dot(V4 a, V4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; } /* :217 */
static V4 w_add(V4 a, V4 b) { return v4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }    /* :221 */
static V4 w_sub(V4 a, V4 b) { return v4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }    /* :225 */
static V4 w_muls(V4 a, double b) { return v4(a.x * b, a.y * b, a.z * b, a.w * b); }       /* :229 */
static V4 w_divs(V4 a, double b) { return v4(a.x / b, a.y / b, a.z / b, a.w / b); }       /* :233 */
static V3 w_vector(V4 a) { return v3(a.x, a.y, a.z); }                                    /* :208 */

/* ---- color.go ------------------------------------------------------------- */

static C4 c4(double r, double g, double b, double a) { C4 c = {r, g, b, a}; return c; }
static C4 c_add(C4 a, C4 b) { return c4(a.r + b.r, a.g + b.g, a.b + b.b, a.a + b.a); }  /* :77 */
static C4 c_mul(C4 a, C4 b) { return c4(a.r * b.r, a.g * b.g, a.b * b.b, a.a * b.a); }  /* :85 */
static C4 c_muls(C4 a, double b) { return c4(a.r * b, a.g * b, a.b * b, a.a * b); }     /* :101 */
static C4 c_min(C4 a, C4 b) {                                                            /* :113 */
    return c4(go_min(a.r, b.r), go_min(a.g, b.g), go_min(a.b, b.b), go_min(a.a, b.a));
}
static int c_is_discard(C4 c) { return c.r == 0 && c.g == 0 && c.b == 0 && c.a == 0; }   /* :11 */
static double clampd(double x, double lo, double hi) {                                   /* util.go:74 */
    if (x < lo) return lo;
    if (x > hi) return hi;
    return x;
}
/* color.go:56-63: clamp, then truncating uint8(c*255). */
static void c_nrgba(C4 c, uint8_t out[4]) {
    const double d = 0xff;
    out[0] = (uint8_t)(go_int(clampd(c.r, 0, 1) * d) & 0xff);
    out[1] = (uint8_t)(go_int(clampd(c.g, 0, 1) * d) & 0xff);
    out[2] = (uint8_t)(go_int(clampd(c.b, 0, 1) * d) & 0xff);
    out[3] = (uint8_t)(go_int(clampd(c.a, 0, 1) * d) & 0xff);
}

/* ---- matrix.go ------------------------------------------------------------ */

static V4 m_mul_position_w(const double *a, V3 b) {                                /* :216 */
    double x = a[0] * b.x + a[1] * b.y + a[2] * b.z + a[3];
    double y = a[4] * b.x + a[5] * b.y + a[6] * b.z + a[7];
    double z = a[8] * b.x + a[9] * b.y + a[10] * b.z + a[11];
    double w = a[12] * b.x + a[13] * b.y + a[14] * b.z + a[15];
    return v4(x, y, z, w);
}
static V3 m_mul_position(const double *a, V3 b) {                                  /* :209 */
    double x = a[0] * b.x + a[1] * b.y + a[2] * b.z + a[3];
    double y = a[4] * b.x + a[5] * b.y + a[6] * b.z + a[7];
    double z = a[8] * b.x + a[9] * b.y + a[10] * b.z + a[11];
    return v3(x, y, z);
}
static void m_screen(int w, int h, double *m) {                                    /* :119 */
    double w2 = (double)w / 2, h2 = (double)h / 2;
    double s[16] = {w2, 0, 0, w2, 0, -h2, 0, h2, 0, 0, 0.5, 0.5, 0, 0, 0, 1};
    memcpy(m, s, sizeof s);
}

/* ---- vertex.go ------------------------------------------------------------ */

static V3 interpolate_vectors(V3 v1, V3 v2, V3 v3_, V4 b) {                         /* :65 */
    V3 n = {0, 0, 0};
    n = v_add(n, v_muls(v1, b.x));
    n = v_add(n, v_muls(v2, b.y));
    n = v_add(n, v_muls(v3_, b.z));
    return v_muls(n, b.w);
}
static C4 interpolate_colors(C4 v1, C4 v2, C4 v3_, V4 b) {                          /* :57 */
    C4 n = {0, 0, 0, 0};
    n = c_add(n, c_muls(v1, b.x));
    n = c_add(n, c_muls(v2, b.y));
    n = c_add(n, c_muls(v3_, b.z));
    return c_muls(n, b.w);
}
static V4 interpolate_vectorws(V4 v1, V4 v2, V4 v3_, V4 b) {                        /* :73 */
    V4 n = {0, 0, 0, 0};
    n = w_add(n, w_muls(v1, b.x));
    n = w_add(n, w_muls(v2, b.y));
    n = w_add(n, w_muls(v3_, b.z));
    return w_muls(n, b.w);
}
static Vertex interpolate_vertexes(const Vertex *v1, const Vertex *v2, const Vertex *v3_, V4 b) { /* :18 */
    Vertex v;
    v.position = interpolate_vectors(v1->position, v2->position, v3_->position, b);
    v.normal = v_normalize(interpolate_vectors(v1->normal, v2->normal, v3_->normal, b));
    v.texture = interpolate_vectors(v1->texture, v2->texture, v3_->texture, b);
    v.color = interpolate_colors(v1->color, v2->color, v3_->color, b);
    v.output = interpolate_vectorws(v1->output, v2->output, v3_->output, b);
    return v;
}
static V4 barycentric(V3 p1, V3 p2, V3 p3, V3 p) {                                  /* :81 */
    V3 v0 = v_sub(p2, p1), v1 = v_sub(p3, p1), v2 = v_sub(p, p1);
    double d00 = v_dot(v0, v0), d01 = v_dot(v0, v1), d11 = v_dot(v1, v1);
    double d20 = v_dot(v2, v0), d21 = v_dot(v2, v1);
    double d = d00 * d11 - d01 * d01;
    double v = (d11 * d20 - d01 * d21) / d;
    double w = (d00 * d21 - d01 * d20) / d;
    double u = 1 - v - w;
    return v4(u, v, w, 1);
}

/* ---- triangle.go ---------------------------------------------------------- */

static void fix_normals(Vertex t[3]) {                                              /* :33-58 */
    V3 e1 = v_sub(t[1].position, t[0].position);
    V3 e2 = v_sub(t[2].position, t[0].position);
    V3 n = v_normalize(v_cross(e1, e2));
    for (int i = 0; i < 3; i++)
        if (v_is_zero(t[i].normal)) t[i].normal = n;
}

/* ---- clipping.go ---------------------------------------------------------- */

typedef struct { V4 P, N; } ClipPlane;
static const ClipPlane clip_planes[6] = {                                           /* :3-10 */
    {{1, 0, 0, 1}, {-1, 0, 0, 1}}, {{-1, 0, 0, 1}, {1, 0, 0, 1}},
    {{0, 1, 0, 1}, {0, -1, 0, 1}}, {{0, -1, 0, 1}, {0, 1, 0, 1}},
    {{0, 0, 1, 1}, {0, 0, -1, 1}}, {{0, 0, -1, 1}, {0, 0, 1, 1}},
};
static int point_in_front(const ClipPlane *p, V4 v) {                               /* :16 */
    return w_dot(w_sub(v, p->P), p->N) > 0;
}
static V4 intersect_segment(const ClipPlane *p, V4 v0, V4 v1) {                     /* :20 */
    V4 u = w_sub(v1, v0);
    V4 w = w_sub(v0, p->P);
    double d = w_dot(p->N, u);
    double n = -w_dot(p->N, w);
    return w_add(v0, w_muls(u, n / d));
}
#define MAX_POLY 16
static int sutherland_hodgman(const V4 *points, int npoints, V4 *out) {             /* :28 */
    V4 bufa[MAX_POLY], bufb[MAX_POLY];
    V4 *output = bufa, *input = bufb;
    int nout = npoints;
    memcpy(output, points, sizeof(V4) * (size_t)npoints);
    for (int pi = 0; pi < 6; pi++) {
        const ClipPlane *plane = &clip_planes[pi];
        V4 *t = input; input = output; output = t;
        int nin = nout;
        nout = 0;
        if (nin == 0) return 0;
        V4 s = input[nin - 1];
        for (int k = 0; k < nin; k++) {
            V4 e = input[k];
            if (point_in_front(plane, e)) {
                if (!point_in_front(plane, s)) output[nout++] = intersect_segment(plane, s, e);
                output[nout++] = e;
            } else if (point_in_front(plane, s)) {
                output[nout++] = intersect_segment(plane, s, e);
            }
            s = e;
        }
    }
    memcpy(out, output, sizeof(V4) * (size_t)nout);
    return nout;
}
/* clipping.go:54-74; out holds up to 7 triangles (3 vertices each). */
static int clip_triangle(const Vertex t[3], Vertex *out) {
    V4 w1 = t[0].output, w2 = t[1].output, w3 = t[2].output;
    V3 p1 = w_vector(w1), p2 = w_vector(w2), p3 = w_vector(w3);
    V4 points[3] = {w1, w2, w3};
    V4 np[MAX_POLY];
    int n = sutherland_hodgman(points, 3, np);
    int nt = 0;
    for (int i = 2; i < n; i++) {
        V4 b1 = barycentric(p1, p2, p3, w_vector(np[0]));
        V4 b2 = barycentric(p1, p2, p3, w_vector(np[i - 1]));
        V4 b3 = barycentric(p1, p2, p3, w_vector(np[i]));
        Vertex *o = out + 3 * nt;
        o[0] = interpolate_vertexes(&t[0], &t[1], &t[2], b1);
        o[1] = interpolate_vertexes(&t[0], &t[1], &t[2], b2);
        o[2] = interpolate_vertexes(&t[0], &t[1], &t[2], b3);
        fix_normals(o); /* NewTriangle, triangle.go:7-11 */
        nt++;
    }
    return nt;
}
/* clipping.go:76-98; returns 0 if the line is entirely outside. */
static int clip_line(Vertex l[2]) {
    V4 w1 = l[0].output, w2 = l[1].output;
    for (int pi = 0; pi < 6; pi++) {
        const ClipPlane *plane = &clip_planes[pi];
        int f1 = point_in_front(plane, w1), f2 = point_in_front(plane, w2);
        if (f1 && f2) continue;
        else if (f1) w2 = intersect_segment(plane, w1, w2);
        else if (f2) w1 = intersect_segment(plane, w2, w1);
        else return 0;
    }
    l[0].output = w1;
    l[1].output = w2;
    return 1;
}

/* ---- texture.go ----------------------------------------------------------- */

/* image.Image.At + MakeColor (color.go:25-29); Go stdlib image/color RGBA(). */
static C4 tex_at(const oshader *s, int64_t x, int64_t y) {
    if (x < 0 || y < 0 || x >= s->tex_w || y >= s->tex_h) return c4(0, 0, 0, 0);
    const size_t t = (size_t)y * (size_t)s->tex_w + (size_t)x;
    const uint8_t *p = s->tex + t * 4;
    uint32_t r, g, b, a;
    if (s->tex_format == O_TEX_RGBA64) { /* the 16-bit values of Color.RGBA() themselves (YCbCr, Gray, Paletted, *64 ...) */
        uint16_t q[4];
        memcpy(q, s->tex + t * 8, 8);
        r = q[0]; g = q[1]; b = q[2]; a = q[3];
    } else if (s->tex_format == O_TEX_NRGBA) { /* color.NRGBA.RGBA() */
        r = p[0]; r |= r << 8; r *= p[3]; r /= 0xff;
        g = p[1]; g |= g << 8; g *= p[3]; g /= 0xff;
        b = p[2]; b |= b << 8; b *= p[3]; b /= 0xff;
        a = p[3]; a |= a << 8;
    } else {                            /* color.RGBA.RGBA() */
        r = p[0]; r |= r << 8;
        g = p[1]; g |= g << 8;
        b = p[2]; b |= b << 8;
        a = p[3]; a |= a << 8;
    }
    const double d = 0xffff;
    return c4((double)r / d, (double)g / d, (double)b / d, (double)a / d);
}
static C4 bilinear_sample(const oshader *s, double u, double v) {                   /* :41-63 */
    v = 1 - v;
    u -= floor(u);
    v -= floor(v);
    double x = u * (double)(s->tex_w - 1);
    double y = v * (double)(s->tex_h - 1);
    int64_t x0 = go_int(x), y0 = go_int(y);
    int64_t x1 = x0 + 1, y1 = y0 + 1;
    x -= (double)x0;
    y -= (double)y0;
    C4 c00 = tex_at(s, x0, y0), c01 = tex_at(s, x0, y1);
    C4 c10 = tex_at(s, x1, y0), c11 = tex_at(s, x1, y1);
    C4 c = {0, 0, 0, 0};
    c = c_add(c, c_muls(c00, (1 - x) * (1 - y)));
    c = c_add(c, c_muls(c10, x * (1 - y)));
    c = c_add(c, c_muls(c01, (1 - x) * y));
    c = c_add(c, c_muls(c11, x * y));
    return c;
}

/* ---- shader.go ------------------------------------------------------------ */

static C4 c_from(const double *p) { return c4(p[0], p[1], p[2], p[3]); }
static V3 v_from(const double *p) { return v3(p[0], p[1], p[2]); }

static C4 shader_fragment(const oshader *s, const Vertex *v) {
    if (s->kind == O_SHADER_SOLID) return c_from(s->color);                         /* :25 */
    if (s->kind == O_SHADER_TEXTURE) return bilinear_sample(s, v->texture.x, v->texture.y); /* :44 */
    /* PhongShader.Fragment, shader.go:75-96 */
    C4 light = c_from(s->ambient);
    C4 color = v->color;
    C4 object = c_from(s->object);
    if (!c_is_discard(object)) color = object;
    if (s->has_texture) color = bilinear_sample(s, v->texture.x, v->texture.y);
    V3 ld = v_from(s->light);
    double diffuse = go_max(v_dot(v->normal, ld), 0);
    light = c_add(light, c_muls(c_from(s->diffuse), diffuse));
    if (diffuse > 0 && s->specular_power > 0) {
        V3 camera = v_normalize(v_sub(v_from(s->camera), v->position));
        V3 reflected = v_reflect(v_negate(ld), v->normal);
        double specular = go_max(v_dot(camera, reflected), 0);
        if (specular > 0) {
            specular = oracle_pow(specular, s->specular_power);
            light = c_add(light, c_muls(c_from(s->specular), specular));
        }
    }
    C4 r = c_min(c_mul(color, light), c4(1, 1, 1, 1));
    r.a = color.a;
    return r;
}
void oracle_fragment(const oshader *s, const overtex *v, double rgba[4]) {
    C4 c = shader_fragment(s, (const Vertex *)v);
    rgba[0] = c.r; rgba[1] = c.g; rgba[2] = c.b; rgba[3] = c.a;
}
static Vertex shader_vertex(const oshader *s, const Vertex *v) {                    /* :20,39,70 */
    Vertex r = *v;
    r.output = m_mul_position_w(s->matrix, v->position);
    return r;
}

/* ---- context.go ----------------------------------------------------------- */

typedef struct {
    octx *c;
    const oshader *s;
    double screen[16];
    pthread_mutex_t *locks; /* NULL when sequential */
} Draw;

void oracle_clear_color(octx *c, const double rgba[4]) {                            /* :119-131 */
    uint8_t p[4];
    c_nrgba(c_from(rgba), p);
    size_t n = (size_t)c->width * (size_t)c->height;
    for (size_t i = 0; i < n; i++) memcpy(c->color + 4 * i, p, 4);
}
void oracle_clear_depth(octx *c, double value) {                                    /* :137-141 */
    size_t n = (size_t)c->width * (size_t)c->height;
    for (size_t i = 0; i < n; i++) c->depth[i] = value;
}

static double edge(V3 a, V3 b, V3 c) {                                              /* :147 */
    return (b.x - c.x) * (a.y - c.y) - (b.y - c.y) * (a.x - c.x);
}

static double depth_load(const Draw *d, int64_t i) {
    if (!d->locks) return d->c->depth[i];
    uint64_t bits = __atomic_load_n((const uint64_t *)&d->c->depth[i], __ATOMIC_RELAXED);
    double v;
    memcpy(&v, &bits, 8);
    return v;
}
static void depth_store(const Draw *d, int64_t i, double v) {
    if (!d->locks) { d->c->depth[i] = v; return; }
    uint64_t bits;
    memcpy(&bits, &v, 8);
    __atomic_store_n((uint64_t *)&d->c->depth[i], bits, __ATOMIC_RELAXED);
}

/* context.go:151-281 */
static oinfo rasterize(const Draw *dr, const Vertex *v0, const Vertex *v1, const Vertex *v2,
                       V3 s0, V3 s1, V3 s2) {
    octx *dc = dr->c;
    oinfo info = {0, 0};
    const int64_t W = dc->width, H = dc->height, len = W * H;

    /* integer bounding box (:155-160) */
    V3 mn = v_min(s0, v_min(s1, s2));
    V3 mx = v_max(s0, v_max(s1, s2));
    int64_t x0 = go_int(floor(mn.x)), x1 = go_int(ceil(mx.x));
    int64_t y0 = go_int(floor(mn.y)), y1 = go_int(ceil(mx.y));

    /* forward differencing variables (:163-172) */
    V3 p = v3((double)x0 + 0.5, (double)y0 + 0.5, 0);
    double w00 = edge(s1, s2, p), w01 = edge(s2, s0, p), w02 = edge(s0, s1, p);
    double a01 = s1.y - s0.y, b01 = s0.x - s1.x;
    double a12 = s2.y - s1.y, b12 = s1.x - s2.x;
    double a20 = s0.y - s2.y, b20 = s2.x - s0.x;

    /* reciprocals (:175-181) */
    double ra = 1 / edge(s0, s1, s2);
    double r0 = 1 / v0->output.w, r1 = 1 / v1->output.w, r2 = 1 / v2->output.w;
    double ra12 = 1 / a12, ra20 = 1 / a20, ra01 = 1 / a01;

    for (int64_t y = y0; y <= y1; y++) {                                            /* :184 */
        double d = 0;
        double d0 = -w00 * ra12, d1 = -w01 * ra20, d2 = -w02 * ra01;
        if (w00 < 0 && d0 > d) d = d0;
        if (w01 < 0 && d1 > d) d = d1;
        if (w02 < 0 && d2 > d) d = d2;
        d = (double)go_int(d);
        if (d < 0) d = 0;
        double w0 = w00 + a12 * d, w1 = w01 + a20 * d, w2 = w02 + a01 * d;
        int was_inside = 0;
        for (int64_t x = x0 + go_int(d); x <= x1; x++) {                            /* :207 */
            double b0 = w0 * ra, b1 = w1 * ra, b2 = w2 * ra;
            w0 += a12; w1 += a20; w2 += a01;
            if (b0 < 0 || b1 < 0 || b2 < 0) {
                if (was_inside) break;
                continue;
            }
            was_inside = 1;
            /* Go integer arithmetic wraps: a triangle whose screen coordinates are all NaN (a degenerate triangle
             * that went through ClipTriangle: Barycentric divides 0 by 0) has x0 = x1 = y0 = y1 = int(NaN) =
             * -2^63, and its one "pixel" gets the index -2^63 * W - 2^63, which is 0 when W is odd. */
            int64_t i = (int64_t)((uint64_t)y * (uint64_t)W + (uint64_t)x);
            if (i < 0 || i >= len) continue;                                        /* :224 */
            if (dc->x_guard && (x < 0 || x >= W)) continue;    /* adopted rule */
            info.total_pixels++;
            double z = b0 * s0.z + b1 * s1.z + b2 * s2.z;
            double bz = z + dc->depth_bias;
            if (dc->read_depth && bz > depth_load(dr, i)) continue;                 /* :232 */
            /* perspective-correct interpolation (:236-238) */
            V4 b = v4(b0 * r0, b1 * r1, b2 * r2, 0);
            b.w = 1 / (b.x + b.y + b.z);
            Vertex v = interpolate_vertexes(v0, v1, v2, b);
            C4 color = shader_fragment(dr->s, &v);                                  /* :240 */
            if (c_is_discard(color)) continue;
            pthread_mutex_t *lock = dr->locks ? &dr->locks[(x + y) & 255] : NULL;   /* :245 */
            if (lock) pthread_mutex_lock(lock);
            if (bz <= depth_load(dr, i) || !dc->read_depth) {                       /* :248 */
                info.updated_pixels++;
                if (dc->write_depth) depth_store(dr, i, z);
                if (dc->write_color) {
                    uint8_t c8[4];
                    c_nrgba(color, c8);
                    if (dc->alpha_blend && color.a < 1) {                           /* :256-267 */
                        /* Go stdlib color.NRGBA.RGBA() */
                        uint32_t sa = c8[3]; sa |= sa << 8;
                        uint32_t sr = c8[0]; sr |= sr << 8; sr *= c8[3]; sr /= 0xff;
                        uint32_t sg = c8[1]; sg |= sg << 8; sg *= c8[3]; sg /= 0xff;
                        uint32_t sb = c8[2]; sb |= sb << 8; sb *= c8[3]; sb /= 0xff;
                        uint32_t a = (0xffff - sa) * 0x101;
                        uint8_t *px = dc->color + 4 * i; /* PixOffset(x,y) == 4*i */
                        px[0] = (uint8_t)(((uint32_t)px[0] * a / 0xffff + sr) >> 8);
                        px[1] = (uint8_t)(((uint32_t)px[1] * a / 0xffff + sg) >> 8);
                        px[2] = (uint8_t)(((uint32_t)px[2] * a / 0xffff + sb) >> 8);
                        px[3] = (uint8_t)(((uint32_t)px[3] * a / 0xffff + sa) >> 8);
                    } else if (x >= 0 && x < W && y >= 0 && y < H) {                /* SetNRGBA bounds check */
                        memcpy(dc->color + 4 * (y * W + x), c8, 4);
                    }
                }
            }
            if (lock) pthread_mutex_unlock(lock);
        }
        w00 += b12; w01 += b20; w02 += b01;                                         /* :275 */
    }
    return info;
}

static oinfo info_add(oinfo a, oinfo b) {
    oinfo r = {a.total_pixels + b.total_pixels, a.updated_pixels + b.updated_pixels};
    return r;
}

/* Screen-space corners of the fat line, context.go:283-290. */
static void line_corners(double lw, V3 s0, V3 s1, V3 *s00, V3 *s01, V3 *s10, V3 *s11) {
    V3 n = v_muls(v_perpendicular(v_sub(s1, s0)), lw / 2);
    s0 = v_add(s0, v_muls(v_normalize(v_sub(s0, s1)), lw / 2));
    s1 = v_add(s1, v_muls(v_normalize(v_sub(s1, s0)), lw / 2));
    *s00 = v_add(s0, n); *s01 = v_sub(s0, n);
    *s10 = v_add(s1, n); *s11 = v_sub(s1, n);
}
static oinfo draw_line_quad(const Draw *dr, const Vertex *v0, const Vertex *v1, V3 s0, V3 s1) { /* :283-294 */
    V3 s00, s01, s10, s11;
    line_corners(dr->c->line_width, s0, s1, &s00, &s01, &s10, &s11);
    oinfo i1 = rasterize(dr, v1, v0, v0, s11, s01, s00);
    oinfo i2 = rasterize(dr, v1, v1, v0, s10, s11, s00);
    return info_add(i1, i2);
}
static oinfo draw_wireframe(const Draw *dr, const Vertex *v0, const Vertex *v1, const Vertex *v2,
                            V3 s0, V3 s1, V3 s2) {                                  /* :296-301 */
    oinfo i1 = draw_line_quad(dr, v0, v1, s0, s1);
    oinfo i2 = draw_line_quad(dr, v1, v2, s1, s2);
    oinfo i3 = draw_line_quad(dr, v2, v0, s2, s0);
    return info_add(info_add(i1, i2), i3);
}
static oinfo draw_clipped_line(const Draw *dr, const Vertex *v0, const Vertex *v1) { /* :303-314 */
    V3 ndc0 = w_vector(w_divs(v0->output, v0->output.w));
    V3 ndc1 = w_vector(w_divs(v1->output, v1->output.w));
    V3 s0 = m_mul_position(dr->screen, ndc0);
    V3 s1 = m_mul_position(dr->screen, ndc1);
    return draw_line_quad(dr, v0, v1, s0, s1);
}

/* context.go:316-341: NDC, swap, cull, screen.  Returns 0 if culled; on
 * success v[] is in its final (possibly swapped) order and s[] holds the
 * three screen-space vertices. */
static int setup_clipped_triangle(const octx *dc, const double *screen, const Vertex **v, V3 s[3]) {
    V3 ndc0 = w_vector(w_divs(v[0]->output, v[0]->output.w));
    V3 ndc1 = w_vector(w_divs(v[1]->output, v[1]->output.w));
    V3 ndc2 = w_vector(w_divs(v[2]->output, v[2]->output.w));
    double a = (ndc1.x - ndc0.x) * (ndc2.y - ndc0.y) - (ndc2.x - ndc0.x) * (ndc1.y - ndc0.y);
    if (a < 0) {
        const Vertex *tv = v[0]; v[0] = v[2]; v[2] = tv;
        V3 tn = ndc0; ndc0 = ndc2; ndc2 = tn;
    }
    if (dc->cull == O_CULL_FRONT) a = -a;
    if (dc->front_face == O_FACE_CW) a = -a;
    if (dc->cull != O_CULL_NONE && a <= 0) return 0;
    s[0] = m_mul_position(screen, ndc0);
    s[1] = m_mul_position(screen, ndc1);
    s[2] = m_mul_position(screen, ndc2);
    return 1;
}
static oinfo draw_clipped_triangle(const Draw *dr, const Vertex *v0, const Vertex *v1, const Vertex *v2) {
    const Vertex *v[3] = {v0, v1, v2};
    V3 s[3];
    oinfo none = {0, 0};
    if (!setup_clipped_triangle(dr->c, dr->screen, v, s)) return none;
    if (dr->c->wireframe) return draw_wireframe(dr, v[0], v[1], v[2], s[0], s[1], s[2]); /* :344 */
    return rasterize(dr, v[0], v[1], v[2], s[0], s[1], s[2]);
}

static oinfo draw_line(const Draw *dr, const Vertex *l) {                            /* :351-368 */
    Vertex v[2] = {shader_vertex(dr->s, &l[0]), shader_vertex(dr->s, &l[1])};
    oinfo none = {0, 0};
    if (w_outside(v[0].output) || w_outside(v[1].output)) {
        if (!clip_line(v)) return none;
        return draw_clipped_line(dr, &v[0], &v[1]);
    }
    return draw_clipped_line(dr, &v[0], &v[1]);
}

static oinfo draw_triangle(const Draw *dr, const Vertex *t) {                        /* :370-389 */
    Vertex v[3] = {shader_vertex(dr->s, &t[0]), shader_vertex(dr->s, &t[1]), shader_vertex(dr->s, &t[2])};
    if (w_outside(v[0].output) || w_outside(v[1].output) || w_outside(v[2].output)) {
        fix_normals(v); /* NewTriangle(v1,v2,v3) at :378 */
        Vertex clipped[3 * 8];
        int n = clip_triangle(v, clipped);
        oinfo result = {0, 0};
        for (int k = 0; k < n; k++)
            result = info_add(result, draw_clipped_triangle(dr, &clipped[3 * k], &clipped[3 * k + 1], &clipped[3 * k + 2]));
        return result;
    }
    return draw_clipped_triangle(dr, &v[0], &v[1], &v[2]);
}

size_t oracle_setup_triangle(const octx *c, const oshader *s, const overtex tri[3],
                             overtex *out_v, double *out_s, size_t max_out) {
    const Vertex *t = (const Vertex *)tri;
    double screen[16];
    m_screen(c->width, c->height, screen);
    Vertex v[3] = {shader_vertex(s, &t[0]), shader_vertex(s, &t[1]), shader_vertex(s, &t[2])};
    Vertex clipped[3 * 8];
    const Vertex *src = v;
    int n = 1;
    if (w_outside(v[0].output) || w_outside(v[1].output) || w_outside(v[2].output)) {
        fix_normals(v);
        n = clip_triangle(v, clipped);
        src = clipped;
    }
    size_t emitted = 0;
    for (int k = 0; k < n && emitted < max_out; k++) {
        const Vertex *pv[3] = {&src[3 * k], &src[3 * k + 1], &src[3 * k + 2]};
        V3 sc[3];
        if (!setup_clipped_triangle(c, screen, pv, sc)) continue;
        for (int j = 0; j < 3; j++) {
            memcpy(&out_v[3 * emitted + j], pv[j], sizeof(Vertex));
            out_s[9 * emitted + 3 * j + 0] = sc[j].x;
            out_s[9 * emitted + 3 * j + 1] = sc[j].y;
            out_s[9 * emitted + 3 * j + 2] = sc[j].z;
        }
        emitted++;
    }
    return emitted;
}

typedef struct {
    Draw dr;
    const Vertex *prims;
    size_t n;
    int wi, wn, is_lines;
    oinfo result;
} Worker;

static void *worker_main(void *arg) {                                                /* :416-426 */
    Worker *w = (Worker *)arg;
    oinfo result = {0, 0};
    for (size_t i = 0; i < w->n; i++) {
        if ((int)(i % (size_t)w->wn) != w->wi) continue;
        oinfo info = w->is_lines ? draw_line(&w->dr, w->prims + 2 * i) : draw_triangle(&w->dr, w->prims + 3 * i);
        result = info_add(result, info);
    }
    w->result = result;
    return NULL;
}

static void draw_prims(octx *c, const oshader *s, const Vertex *prims, size_t n, int nthreads,
                       int is_lines, oinfo *info) {
    Draw dr;
    dr.c = c; dr.s = s; dr.locks = NULL;
    m_screen(c->width, c->height, dr.screen);
    oinfo result = {0, 0};
    if (nthreads <= 1) {
        for (size_t i = 0; i < n; i++)
            result = info_add(result, is_lines ? draw_line(&dr, prims + 2 * i) : draw_triangle(&dr, prims + 3 * i));
    } else {
        pthread_mutex_t locks[256];                                                  /* :78 */
        for (int i = 0; i < 256; i++) pthread_mutex_init(&locks[i], NULL);
        dr.locks = locks;
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
        Worker *ws = (Worker *)malloc(sizeof(Worker) * (size_t)nthreads);
        for (int wi = 0; wi < nthreads; wi++) {
            ws[wi].dr = dr; ws[wi].prims = prims; ws[wi].n = n;
            ws[wi].wi = wi; ws[wi].wn = nthreads; ws[wi].is_lines = is_lines;
            pthread_create(&th[wi], NULL, worker_main, &ws[wi]);
        }
        for (int wi = 0; wi < nthreads; wi++) {
            pthread_join(th[wi], NULL);
            result = info_add(result, ws[wi].result);
        }
        free(th); free(ws);
        for (int i = 0; i < 256; i++) pthread_mutex_destroy(&locks[i]);
    }
    if (info) *info = result;
}

void oracle_draw_triangles(octx *c, const oshader *s, const overtex *tris, size_t ntris,
                           int nthreads, oinfo *info) {
    draw_prims(c, s, (const Vertex *)tris, ntris, nthreads, 0, info);
}
void oracle_draw_lines(octx *c, const oshader *s, const overtex *lines, size_t nlines,
                       int nthreads, oinfo *info) {
    draw_prims(c, s, (const Vertex *)lines, nlines, nthreads, 1, info);
}

/* One RasterizeInfo per primitive, drawn one by one in index order: what a caller
 * looping over Context.DrawTriangle / DrawLine sees (context.go:351-389;
 * examples/silhouette.go:163-166 uses the per-line UpdatedPixels/TotalPixels). */
void oracle_draw_each(octx *c, const oshader *s, const overtex *prims, size_t n, int is_lines, oinfo *infos) {
    Draw dr;
    dr.c = c; dr.s = s; dr.locks = NULL;
    m_screen(c->width, c->height, dr.screen);
    const Vertex *v = (const Vertex *)prims;
    for (size_t i = 0; i < n; i++) infos[i] = is_lines ? draw_line(&dr, v + 2 * i) : draw_triangle(&dr, v + 3 * i);
}

/* Context.DepthImage, context.go:87-117: Gray16 of the depth buffer. */
void oracle_depth_image(const double *depth, int width, int height, uint16_t *out) {
    const double MAXF = 1.7976931348623157e308;
    double lo = MAXF, hi = -MAXF;                                                    /* :88-89 */
    size_t n = (size_t)width * (size_t)height;
    for (size_t i = 0; i < n; i++) {
        double d = depth[i];
        if (d == MAXF) continue;                                                     /* :91-93 */
        if (d < lo) lo = d;
        if (d > hi) hi = d;
    }
    for (size_t i = 0; i < n; i++) {
        double d = depth[i];
        double t = (d - lo) / (hi - lo);                                             /* :107 */
        if (d == MAXF) t = 1;                                                        /* :108-110 */
        out[i] = (uint16_t)(go_int(t * 65535.0) & 0xffff);                           /* :111 uint16(t * 0xffff) */
    }
}

/* loadSTLB, stl.go:86-154: `records` = n 50-byte binary STL triangle records.  out = 3n
 * vertices: positions widened from float32 (makeFloat, :82-84), normals = Triangle.Normal()
 * (triangle.go:33-37) copied to the three vertices; everything else zero. */
void oracle_stl_triangles(const uint8_t *records, size_t n, overtex *out) {
    memset(out, 0, sizeof(overtex) * 3 * n);
    Vertex *v = (Vertex *)out;
    for (size_t i = 0; i < n; i++) {
        const uint8_t *b = records + 50 * i;
        V3 p[3];
        for (int k = 0; k < 3; k++) {
            float f[3];
            for (int c = 0; c < 3; c++) {
                const uint8_t *q = b + 12 + 12 * k + 4 * c;
                uint32_t bits = (uint32_t)q[0] | ((uint32_t)q[1] << 8) | ((uint32_t)q[2] << 16) | ((uint32_t)q[3] << 24);
                memcpy(&f[c], &bits, 4);
            }
            p[k] = v3((double)f[0], (double)f[1], (double)f[2]);
        }
        V3 nn = v_normalize(v_cross(v_sub(p[1], p[0]), v_sub(p[2], p[0])));
        for (int k = 0; k < 3; k++) { v[3 * i + k].position = p[k]; v[3 * i + k].normal = nn; }
    }
}

/* ---- SSAA resolve: nfnt/resize Bilinear on *image.NRGBA -------------------
 * External dependency, absent from /root/reference and unpinned (no go.mod);
 * restated from the published algorithm of github.com/nfnt/resize
 * (resize.go Resize + createWeights8, converter.go resizeNRGBA/resizeRGBA,
 * filters.go linear).  Call sites: examples/teapot.go:60, dragon.go:82, ... */

static uint8_t clamp_u8(int32_t v) { /* converter.go clampUint8 */
    if ((uint32_t)v < 256) return (uint8_t)v;
    if (v > 255) return 255;
    return 0;
}
typedef struct { int16_t *coeffs; int *start; int flen; } Weights;
static Weights create_weights8(int dy, int taps, double scale) {
    const double blur = 1.0;
    Weights w;
    w.flen = taps * (int)go_max(ceil(blur * scale), 1);
    double ff = go_min(1. / (blur * scale), 1);
    w.coeffs = (int16_t *)malloc(sizeof(int16_t) * (size_t)dy * (size_t)w.flen);
    w.start = (int *)malloc(sizeof(int) * (size_t)dy);
    for (int y = 0; y < dy; y++) {
        double ix = scale * ((double)y + 0.5) - 0.5;
        w.start[y] = (int)go_int(ix) - w.flen / 2 + 1;
        ix -= (double)w.start[y];
        for (int i = 0; i < w.flen; i++) {
            double in = fabs((ix - (double)i) * ff);
            double k = in <= 1 ? 1 - in : 0;
            w.coeffs[y * w.flen + i] = (int16_t)go_int(k * 256);
        }
    }
    return w;
}
/* One separable pass.  in: iw x ih, stride 4*iw.  out is transposed:
 * width ih, height ow.  premul!=0: forward alpha premultiplication of NRGBA. */
static void resize_pass(const uint8_t *in, int iw, int ih, uint8_t *out, int ow,
                        const Weights *w, int premul) {
    int max_x = iw - 1;
    for (int x = 0; x < ih; x++) {
        const uint8_t *row = in + (size_t)x * (size_t)iw * 4;
        for (int y = 0; y < ow; y++) {
            int32_t rgba[4] = {0, 0, 0, 0}, sum = 0;
            int start = w->start[y];
            for (int i = 0; i < w->flen; i++) {
                int16_t coeff = w->coeffs[y * w->flen + i];
                if (coeff == 0) continue;
                int xi = start + i;
                if (xi >= 0 && xi < max_x) xi *= 4;
                else if (xi >= max_x) xi = 4 * max_x;
                else xi = 0;
                int32_t a = row[xi + 3], r = row[xi], g = row[xi + 1], b = row[xi + 2];
                if (premul) { r = r * a / 0xff; g = g * a / 0xff; b = b * a / 0xff; }
                rgba[0] += coeff * r; rgba[1] += coeff * g; rgba[2] += coeff * b; rgba[3] += coeff * a;
                sum += coeff;
            }
            uint8_t *o = out + ((size_t)y * (size_t)ih + (size_t)x) * 4;
            o[0] = clamp_u8(rgba[0] / sum); o[1] = clamp_u8(rgba[1] / sum);
            o[2] = clamp_u8(rgba[2] / sum); o[3] = clamp_u8(rgba[3] / sum);
        }
    }
}
void oracle_resolve(const uint8_t *src, int sw, int sh, int dw, int dh, uint8_t *dst) {
    if (sw == dw && sh == dh) { memcpy(dst, src, (size_t)sw * (size_t)sh * 4); return; }
    double scale_x = (double)sw / (double)dw, scale_y = (double)sh / (double)dh;
    uint8_t *temp = (uint8_t *)malloc((size_t)sh * (size_t)dw * 4); /* width sh, height dw */
    Weights wx = create_weights8(dw, 2, scale_x);
    resize_pass(src, sw, sh, temp, dw, &wx, 1);
    Weights wy = create_weights8(dh, 2, scale_y);
    resize_pass(temp, sh, dw, dst, dh, &wy, 0);
    free(wx.coeffs); free(wx.start); free(wy.coeffs); free(wy.start); free(temp);
}

/* ---- sort-last composite key (not in the reference) ------------ */
uint64_t oracle_pack_key(double depth, const uint8_t rgba[4]) {
    uint32_t d32;
    if (!(depth >= 0)) d32 = 0;                       /* NaN / negative roundoff */
    else if (depth > 1) d32 = 0xFFFFFFFFu;            /* cleared (MaxFloat64) */
    else if (depth == 1) d32 = 0xFFFFFFFEu;
    else d32 = (uint32_t)(depth * 4294967295.0);
    uint32_t c = ((uint32_t)rgba[0] << 24) | ((uint32_t)rgba[1] << 16) | ((uint32_t)rgba[2] << 8) | rgba[3];
    /* biased by 2^63 so that a signed 64-bit min (ncclInt64) orders by depth, then colour */
    return (((uint64_t)d32 << 32) | c) ^ 0x8000000000000000ull;
}
