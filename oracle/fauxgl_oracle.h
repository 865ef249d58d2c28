/*
 * fauxgl_oracle.h -- CPU restatement of fogleman/fauxgl's DrawMesh path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or the
 * reported CPU baseline.  The product (libfauxgl_b200.so) never links it.
 *
 * PARITY UNPINNED: the reference ships no tests, no golden images and cannot
 * be compiled here (no Go toolchain; mesh.go:6 imports an un-vendored module).
 * This file follows the reference source line by line (citations below are
 * into /root/reference) in float64 with unfused arithmetic
 * (-ffp-contract=off), evaluation order as written in the Go source.
 */
#ifndef FAUXGL_ORACLE_H
#define FAUXGL_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* vertex.go:3-12 -- 17 float64, the memory layout of the Go struct. */
typedef struct {
    double position[3];
    double normal[3];
    double texture[3];
    double color[4];
    double output[4];
} overtex;

enum { O_SHADER_SOLID = 1, O_SHADER_TEXTURE = 2, O_SHADER_PHONG = 3 };
enum { O_FACE_CW = 1, O_FACE_CCW = 2 };               /* context.go:13-17 */
enum { O_CULL_NONE = 1, O_CULL_FRONT = 2, O_CULL_BACK = 3 }; /* context.go:21-26 */
enum { O_TEX_RGBA = 0, O_TEX_NRGBA = 1,    /* Go *image.RGBA vs *image.NRGBA: 4 bytes per texel */
       O_TEX_RGBA64 = 2 };                 /* any other image type: what At(x,y).RGBA() returns, 4 x uint16 per texel */

/* shader.go:11-14, 30-33, 49-59 */
typedef struct {
    int32_t kind;
    int32_t has_texture;      /* Texture != nil */
    double matrix[16];        /* row-major X00..X33, matrix.go:5-10 */
    double light[3];
    double camera[3];
    double object[4];         /* ObjectColor; all-zero == Discard (shader.go:78) */
    double ambient[4];
    double diffuse[4];
    double specular[4];
    double specular_power;
    double color[4];          /* SolidColorShader.Color */
    const uint8_t *tex;       /* tex_w*tex_h*4 bytes, row-major, stride 4*tex_w */
    int32_t tex_w, tex_h;
    int32_t tex_format;       /* O_TEX_RGBA (png RGB8 -> *image.RGBA) or O_TEX_NRGBA */
    int32_t _pad;
} oshader;

/* context.go:40-58 */
typedef struct {
    int32_t width, height;
    uint8_t *color;           /* NRGBA8, stride 4*width  (image.NRGBA.Pix) */
    double *depth;            /* width*height            (DepthBuffer)     */
    int32_t read_depth, write_depth, write_color, alpha_blend, wireframe;
    int32_t front_face, cull;
    int32_t x_guard;          /* 0: faithful context.go:223-228 (x never range
                                 checked, wraps into the next row); 1: drop
                                 fragments with x outside [0,width) -- the rule
                                 the GPU back end adopts. */
    double line_width, depth_bias;
} octx;

typedef struct { uint64_t total_pixels, updated_pixels; } oinfo; /* context.go:28-31 */

void oracle_clear_color(octx *c, const double rgba[4]);          /* context.go:119-131 */
void oracle_clear_depth(octx *c, double value);                  /* context.go:137-141 */

/* context.go:413-433.  nthreads<=1: sequential, triangle-index order (the
 * canonical schedule).  nthreads>1: the reference's own schedule:
 * worker wi takes i%wn==wi, 256 mutexes hashed by (x+y)&255, unlocked early-Z. */
void oracle_draw_triangles(octx *c, const oshader *s, const overtex *tris,
                           size_t ntris, int nthreads, oinfo *info);
/* context.go:391-411 */
void oracle_draw_lines(octx *c, const oshader *s, const overtex *lines,
                       size_t nlines, int nthreads, oinfo *info);

/* Primitives drawn one by one in index order, one RasterizeInfo each (Context.DrawTriangle /
 * DrawLine return values, context.go:351-389). */
void oracle_draw_each(octx *c, const oshader *s, const overtex *prims, size_t n, int is_lines, oinfo *infos);
/* Context.DepthImage, context.go:87-117 -> width*height Gray16 values. */
void oracle_depth_image(const double *depth, int width, int height, uint16_t *out);
/* loadSTLB, stl.go:86-154: n 50-byte records -> 3n vertices (position, face normal). */
void oracle_stl_triangles(const uint8_t *records, size_t n, overtex *out);

/* Per-stage probes used by unit tests. */
/* DrawTriangle up to (not including) rasterize: emits the post-clip,
 * post-cull, post-swap triangles as 3 overtex + 3 screen vectors each.
 * Returns the number emitted (<= max_out). context.go:370-389,316-341. */
size_t oracle_setup_triangle(const octx *c, const oshader *s, const overtex tri[3],
                             overtex *out_v, double *out_s, size_t max_out);
void oracle_fragment(const oshader *s, const overtex *v, double rgba[4]); /* shader.go */
double oracle_pow(double x, double y);                                   /* Go math.Pow */

/* SSAA resolve == nfnt/resize Resize(w/f,h/f,img,Bilinear) on *image.NRGBA
 * (external, un-vendored, unpinned: restated from the published algorithm). */
void oracle_resolve(const uint8_t *src, int sw, int sh, int dw, int dh, uint8_t *dst);

/* packed composite key used by the sort-last multi-GPU path (not in the
 * reference).  Restated here so tests can check the kernel. */
uint64_t oracle_pack_key(double depth, const uint8_t rgba[4]);

#ifdef __cplusplus
}
#endif
#endif
