/*
 * fauxgl_b200.h -- C ABI of the B200-native rasterisation back end for
 * fogleman/fauxgl's DrawMesh path.
 *
 * The reference is pure Go and has no FFI seam; the drop-in boundary is its
 * exported Context method set.  Each entry point below replaces
 * the reference function cited beside it (file:line into the reference tree)
 * and is what the cgo shim in go/fauxgl binds.
 *
 * Conventions: plain C, no exceptions cross the boundary.  Every function
 * returns 0 on success or a negative fgl_status; fgl_last_error() gives the
 * message.  Handles are opaque.  The caller owns host buffers, the library
 * owns device buffers.  All entry points are thread-safe per context (calls on
 * one context are serialised by a mutex, matching the reference's contract
 * that concurrent DrawTriangle calls on one Context are legal) and may be
 * called from any OS thread (goroutines migrate): each call selects the
 * context's device itself.  There is no CPU fallback anywhere behind this
 * interface: without a CUDA device fgl_context_create fails.
 */
#ifndef FAUXGL_B200_H
#define FAUXGL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FGL_ABI_VERSION 2

typedef enum {
    FGL_OK = 0,
    FGL_E_INVALID = -1,      /* bad argument */
    FGL_E_CUDA = -2,         /* CUDA runtime error (message has the detail) */
    FGL_E_NO_DEVICE = -3,    /* no usable CUDA device: there is no CPU fallback */
    FGL_E_UNSUPPORTED = -4,  /* e.g. a Shader kind with no device implementation */
    FGL_E_OOM = -5,
    FGL_E_OVERFLOW = -6      /* async draw outgrew its work buffers; re-issue the frame */
} fgl_status;

typedef struct fgl_ctx fgl_ctx;    /* context.go:40-58  Context            */
typedef struct fgl_mesh fgl_mesh;  /* mesh.go:9-13      Mesh (device copy) */
typedef struct fgl_tex fgl_tex;    /* texture.go:21-25  ImageTexture       */

/* context.go:11-26 */
enum { FGL_FACE_CW = 1, FGL_FACE_CCW = 2 };
enum { FGL_CULL_NONE = 1, FGL_CULL_FRONT = 2, FGL_CULL_BACK = 3 };
/* shader.go:11,30,49 -- the closed set of built-in shaders */
enum { FGL_SHADER_SOLID = 1, FGL_SHADER_TEXTURE = 2, FGL_SHADER_PHONG = 3 };
/* Which Go image type the texture decoded to; decides the RGBA() conversion
 * MakeColor applies (color.go:25-29): *image.RGBA or *image.NRGBA, 4 bytes per texel.  Every other image type
 * (*image.YCbCr from a JPEG -- examples/capsule.go:35 --, Gray, Paletted, RGBA64, NRGBA64 ...) is uploaded as
 * FGL_TEX_RGBA64: per texel the four 16-bit values At(x, y).RGBA() returns (r, g, b, a as uint16, 8 bytes),
 * which MakeColor divides by 0xffff -- exact for any image.Image, the conversion itself stays in the host
 * language's image package. */
enum { FGL_TEX_RGBA = 0, FGL_TEX_NRGBA = 1, FGL_TEX_RGBA64 = 2 };

/* The render state fields of Context, context.go:47-55. */
typedef struct {
    int32_t read_depth, write_depth, write_color, alpha_blend, wireframe;
    int32_t front_face;   /* FGL_FACE_*  */
    int32_t cull;         /* FGL_CULL_*  */
    /* 0 (default): the reference's index rule, context.go:223-228 -- a covered pixel is kept iff 0 <= y*W + x < W*H;
     * x itself is never range-checked, so a fat line leaving the screen sideways aliases into the neighbouring
     * row: depth is tested and written there and a blended colour too (PixOffset, :259), an opaque colour is
     * dropped by SetNRGBA's bounds check (:269); all of them count in RasterizeInfo.  1: drop every fragment with
     * x outside [0, width) instead (what a caller who considers the aliasing a bug wants). */
    int32_t x_guard;
    double line_width;
    double depth_bias;
} fgl_state;

/* SolidColorShader / TextureShader / PhongShader, shader.go:11-14,30-33,49-59. */
typedef struct {
    int32_t kind;         /* FGL_SHADER_* */
    int32_t _pad;
    double matrix[16];    /* row-major X00..X33, matrix.go:5-10 */
    double light[3];      /* PhongShader.LightDirection (used as given) */
    double camera[3];     /* PhongShader.CameraPosition */
    double object[4];     /* ObjectColor; all zero == Discard => use vertex colour (shader.go:78) */
    double ambient[4];
    double diffuse[4];
    double specular[4];
    double specular_power;
    double color[4];      /* SolidColorShader.Color */
    const fgl_tex *texture; /* NULL == nil */
} fgl_shader;

/* Host-side mesh description.  Attribute arrays are float64,
 * [nprims][verts_per_prim][k] with k = 3 (position, normal, texture) or 4
 * (color) -- i.e. the fields of the reference's Vertex (vertex.go:3-12)
 * gathered per attribute.  NULL means all-zero.  position is required when the
 * matching count is non-zero. */
typedef struct {
    uint64_t ntriangles;
    const double *position, *normal, *texture, *color;     /* [T][3][k] */
    uint64_t nlines;
    const double *lposition, *lnormal, *ltexture, *lcolor; /* [L][2][k] */
} fgl_mesh_desc;

/* context.go:28-31 */
typedef struct { uint64_t total_pixels, updated_pixels; } fgl_raster_info;

/* Device-side statistics of the most recent draw call on the context. */
typedef struct {
    uint64_t prims_in;       /* primitives submitted                       */
    uint64_t records;        /* raster triangles after clip/cull/expansion */
    uint64_t pairs;          /* covered span segments binned into tiles    */
    uint64_t clip_triangles; /* triangles produced by the clipper          */
    uint32_t tiles_x, tiles_y, tile_w, tile_h;
    uint32_t kernel_launches; /* kernels launched by that draw             */
    uint32_t retries;         /* work-buffer regrows (sync draws only)     */
} fgl_draw_stats;

int fgl_abi_version(void);
/* Thread-local message of the last failing call (ctx may be NULL). */
const char *fgl_last_error(const fgl_ctx *ctx);
int fgl_device_count(void);

/* Page-locked host memory for the streaming entry points (fgl_mesh_update_async, fgl_frame_end): copies
 * from/to it are asynchronous and run at full PCIe rate.  A Go caller keeps flattened meshes here (C
 * memory, no Go pointers inside: cgo may pass it freely). */
int fgl_host_alloc(size_t bytes, void **out);
int fgl_host_free(void *ptr);

/* NewContext, context.go:60-81: colour buffer zeroed (transparent), depth
 * buffer cleared to math.MaxFloat64. */
int fgl_context_create(int width, int height, int device, fgl_ctx **out);
int fgl_context_destroy(fgl_ctx *ctx);
int fgl_context_size(const fgl_ctx *ctx, int *width, int *height);

/* ClearColorBufferWith, context.go:119-131 (rgba = Color.NRGBA(), color.go:56). */
int fgl_clear_color(fgl_ctx *ctx, const uint8_t rgba[4]);
/* ClearDepthBufferWith, context.go:137-141. */
int fgl_clear_depth(fgl_ctx *ctx, double value);

/* Upload a mesh once; draws reference it by handle (the reference walks
 * []*Triangle on every DrawTriangles call, context.go:413-433). */
int fgl_mesh_create(fgl_ctx *ctx, const fgl_mesh_desc *desc, fgl_mesh **out);
/* Re-upload attributes of an existing mesh (same primitive counts) into its
 * device buffers; NULL pointers leave that attribute untouched.  This is what
 * the shim calls when a cached Mesh was mutated on the host between frames
 * (examples/animate.go:66 transforms the mesh on the CPU every frame). */
int fgl_mesh_update(fgl_ctx *ctx, fgl_mesh *mesh, const fgl_mesh_desc *desc);
/* Streaming variant for frame pipelines (examples/animate.go re-poses the mesh on the host every frame):
 * the copies and the transposing kernels run on the context's copy stream and the call returns at
 * once; draws that use the mesh afterwards wait for the upload on the device, and the upload itself
 * waits for earlier draws that still read the buffers.  With two meshes used alternately the upload of
 * frame i+1 overlaps the draw of frame i.  The host arrays (pinned memory for a real overlap) must stay
 * unchanged until fgl_mesh_upload_wait returns. */
int fgl_mesh_update_async(fgl_ctx *ctx, fgl_mesh *mesh, const fgl_mesh_desc *desc);
int fgl_mesh_upload_wait(fgl_ctx *ctx, fgl_mesh *mesh);
int fgl_mesh_destroy(fgl_mesh *mesh);
int fgl_mesh_counts(const fgl_mesh *mesh, uint64_t *ntriangles, uint64_t *nlines);
/* Mesh.Transform, mesh.go:167-175 (+ triangle.go:66-73, line.go:23-28): positions
 * by MulPosition, normals by MulDirection (normalised), on the device. */
int fgl_mesh_transform(fgl_ctx *ctx, fgl_mesh *mesh, const double matrix[16]);
/* An indexed triangle mesh in the shape LoadOBJ parses (obj.go:19-79): the v / vt / vn tables (entry 0 = the zero
 * vector of the reference's 1-based tables) and, per triangle corner, one index into each.  The expansion to the
 * planar device layout -- the gather of obj.go:61-72 and Triangle.FixNormals (triangle.go:46-58: a zero normal
 * becomes the face normal) -- runs on the device, so 36 B of indices per triangle cross PCIe instead of 216 B of
 * expanded float64 attributes.  Indices are validated on the host (FGL_E_INVALID; the reference panics). */
typedef struct fgl_indexed_desc {
    const double *v;        /* [nv][3]  */
    const double *vt;       /* [nvt][3] (Z unused by the built-in shaders) */
    const double *vn;       /* [nvn][3] */
    uint64_t nv, nvt, nvn;
    const int32_t *corners; /* [ntriangles][3][3]: (v, vt, vn) index of V1, V2, V3 */
    uint64_t ntriangles;
} fgl_indexed_desc;
int fgl_mesh_create_indexed(fgl_ctx *ctx, const fgl_indexed_desc *desc, fgl_mesh **out);
/* Re-pose an indexed mesh (a host loop like examples/animate.go:66, which transforms the mesh every frame): the
 * corner indices stayed on the device, only the v / vt / vn tables given (NULL = unchanged; sizes as at creation;
 * desc->corners is ignored) are copied again on the context's copy stream, which carries nothing else, and expanded
 * into the planes on the draw stream, in order behind the draws that still read them.  For the 871 306-triangle benchmark mesh
 * (436 k shared vertices) that is 21 MB per frame instead of 125 MB of expanded position + normal soup.  The host
 * arrays (pinned memory for a real overlap) must stay unchanged until fgl_mesh_upload_wait returns. */
int fgl_mesh_update_indexed_async(fgl_ctx *ctx, fgl_mesh *mesh, const fgl_indexed_desc *desc);
/* Mesh.SmoothNormals, mesh.go:105-120, on the device: every triangle corner receives the normalised sum of the
 * normals of all corners at exactly the same position.  The sums are taken in the reference's order (triangle
 * index, then V1, V2, V3, starting from the zero vector), so the result is bit-identical: corners are grouped by a
 * stable sort on a hash of the position, collisions are separated by comparing the positions themselves (+0 == -0;
 * a corner whose position has a NaN component gets the zero vector, as a Go map never finds a NaN key again).  Lines are left alone, as in the reference. */
int fgl_mesh_smooth_normals(fgl_ctx *ctx, fgl_mesh *mesh);
/* Mesh.SmoothNormalsThreshold, mesh.go:80-103: a corner's new normal is the normalised sum, in corner order, of
 * those normals x at its position with x.Dot(normal) >= cos_threshold.  The caller passes the cosine itself
 * (math.Cos(radians) evaluated by the host language), so no libm difference can move the comparison.  A corner
 * whose position has a NaN component gets NaN (empty list: Vector{}.Normalize()). */
int fgl_mesh_smooth_normals_threshold(fgl_ctx *ctx, fgl_mesh *mesh, double cos_threshold);
/* Read the device copy back (same layout as fgl_mesh_desc; any pointer may be NULL). */
int fgl_mesh_read(fgl_ctx *ctx, const fgl_mesh *mesh, double *position, double *normal,
                  double *lposition, double *lnormal);

/* loadSTLB, stl.go:86-154, straight to the device layout: `records` are the `count` 50-byte
 * triangle records of a binary STL file (everything after the 84-byte header).  Only these bytes
 * cross PCIe; a kernel widens the float32 positions (makeFloat, stl.go:82-84) and gives every
 * vertex the face normal Triangle.Normal() (triangle.go:33-37), as the loader does.  Texture
 * coordinates and colours are zero. */
int fgl_mesh_create_stl(fgl_ctx *ctx, const uint8_t *records, uint64_t count, fgl_mesh **out);
/* Mesh.BoundingBox, mesh.go:153-165, of the device copy (triangles and lines).  An empty mesh gives
 * the zero box.  NaN coordinates are ignored (math.Min/Max would propagate them). */
int fgl_mesh_bounds(fgl_ctx *ctx, const fgl_mesh *mesh, double min_xyz[3], double max_xyz[3]);

/* NewImageTexture, texture.go:27-30.  texels: h rows of w texels, 4 bytes each (FGL_TEX_RGBA / FGL_TEX_NRGBA)
 * or 4 x uint16 each (FGL_TEX_RGBA64), tightly packed. */
int fgl_texture_create(fgl_ctx *ctx, const uint8_t *texels, int width, int height, int format, fgl_tex **out);
int fgl_texture_destroy(fgl_tex *tex);

/* DrawTriangles, context.go:413-433, over triangles [first, first+count) of
 * the mesh, in triangle-index order.  Blocks until the
 * RasterizeInfo is known (the reference returns it by value). */
int fgl_draw_triangles(fgl_ctx *ctx, const fgl_state *state, const fgl_shader *shader,
                       const fgl_mesh *mesh, uint64_t first, uint64_t count, fgl_raster_info *info);
/* DrawLines, context.go:391-411. */
int fgl_draw_lines(fgl_ctx *ctx, const fgl_state *state, const fgl_shader *shader,
                   const fgl_mesh *mesh, uint64_t first, uint64_t count, fgl_raster_info *info);
/* The same draws with one RasterizeInfo PER PRIMITIVE: infos[i] is what Context.DrawTriangle /
 * Context.DrawLine (context.go:370-389, 351-368) would have returned for primitive first+i had the
 * caller drawn the primitives one by one in index order (examples/silhouette.go:163-166 decides
 * line visibility from UpdatedPixels/TotalPixels of each DrawLine).  One launch sequence for the
 * whole range; `info` (may be NULL) receives the sum. */
int fgl_draw_triangles_each(fgl_ctx *ctx, const fgl_state *state, const fgl_shader *shader,
                            const fgl_mesh *mesh, uint64_t first, uint64_t count, fgl_raster_info *infos,
                            fgl_raster_info *info);
int fgl_draw_lines_each(fgl_ctx *ctx, const fgl_state *state, const fgl_shader *shader,
                        const fgl_mesh *mesh, uint64_t first, uint64_t count, fgl_raster_info *infos,
                        fgl_raster_info *info);
/* Same draws, enqueued without waiting; the RasterizeInfo of all draws since
 * the last fgl_sync is accumulated and returned by fgl_sync. */
int fgl_draw_triangles_async(fgl_ctx *ctx, const fgl_state *state, const fgl_shader *shader,
                             const fgl_mesh *mesh, uint64_t first, uint64_t count);
int fgl_draw_lines_async(fgl_ctx *ctx, const fgl_state *state, const fgl_shader *shader,
                         const fgl_mesh *mesh, uint64_t first, uint64_t count);
/* Wait for everything enqueued on the context; info (may be NULL) receives the
 * sum over the async draws since the previous fgl_sync. */
int fgl_sync(fgl_ctx *ctx, fgl_raster_info *info);
/* End of a pipelined frame: enqueues the read-back of the colour buffer into color_dst (may be NULL;
 * pinned memory for a real overlap; stride 0 = width*4) and of the RasterizeInfo accumulated by the async
 * draws since the previous fgl_sync / fgl_frame_end, restarts that accumulation and records a fence.  *fence
 * is created when NULL and reused otherwise.  fgl_fence_wait blocks until that frame's results are in host
 * memory -- later frames keep running -- and returns its RasterizeInfo, or FGL_E_OVERFLOW like fgl_sync. */
typedef struct fgl_fence fgl_fence;
int fgl_frame_end(fgl_ctx *ctx, uint8_t *color_dst, size_t stride_bytes, fgl_fence **fence);
int fgl_fence_wait(fgl_ctx *ctx, fgl_fence *fence, fgl_raster_info *info);
int fgl_fence_destroy(fgl_fence *fence);
int fgl_get_draw_stats(const fgl_ctx *ctx, fgl_draw_stats *out);

/* Recorded frames.  A frame of an animation loop (examples/animate.go:43-67) issues the same calls every time --
 * clears, DrawMesh, resolve -- and each costs ten-odd kernel launches from the host.  Between fgl_graph_begin and
 * fgl_graph_end the context RECORDS the calls that only enqueue work (fgl_clear_color / fgl_clear_depth,
 * fgl_draw_*_async, fgl_mesh_transform, fgl_resolve_device, the composites) into a CUDA graph instead of running
 * them; fgl_graph_launch replays the whole frame with one call, after which fgl_sync / fgl_frame_end / the
 * read-backs work as after the calls themselves (same kernels, same order: identical results).  The shader, the
 * render state and the primitive range are part of the recording.  Calls that wait (synchronous draws,
 * read-backs, fgl_sync) are errors while recording; the work buffers must already be large enough (draw the mesh
 * once synchronously first).  A graph belongs to its context. */
typedef struct fgl_graph fgl_graph;
int fgl_graph_begin(fgl_ctx *ctx);
int fgl_graph_end(fgl_ctx *ctx, fgl_graph **out);
int fgl_graph_launch(fgl_ctx *ctx, fgl_graph *graph);
int fgl_graph_destroy(fgl_graph *graph);

/* Per-stage device timing (CUDA events on the context's stream).  While enabled,
 * every draw records events between its stages; fgl_get_stage_times waits for
 * the stream and returns the milliseconds accumulated since the last call and
 * the number of draws they cover. */
typedef struct {
    float geometry_ms;   /* large draws: k_front alone (vertex transform, clip, cull, setup AND the exact scanline
                            walk, fused); small draws: the geometry kernel                              */
    float spans_ms;      /* large draws: k_seg_index (segments listed in primitive order); small draws: the
                            grid-wide span kernels                                                      */
    float sort_ms;       /* stable sort of segments by strip, busy-strip list                           */
    float raster_ms;     /* k_strip (ordered depth resolve) + k_shade (deferred shading of the winners) */
    uint32_t draws;
} fgl_stage_times;
int fgl_set_profiling(fgl_ctx *ctx, int enabled);
int fgl_get_stage_times(fgl_ctx *ctx, fgl_stage_times *out);

/* Image(), context.go:83-85: ColorBuffer.Pix (NRGBA8, non-premultiplied). */
int fgl_read_color(fgl_ctx *ctx, uint8_t *dst, size_t stride_bytes);
/* DepthBuffer, context.go:44. */
int fgl_read_depth(fgl_ctx *ctx, double *dst);
/* DepthImage(), context.go:87-117: the depth buffer as Gray16, normalised on the device between
 * the smallest and largest depth that is not math.MaxFloat64; cleared pixels give 0xffff.  dst
 * receives width*height uint16 (host). */
int fgl_depth_image(fgl_ctx *ctx, uint16_t *dst);
/* The reference exposes both buffers as writable fields; these upload them. */
int fgl_write_color(fgl_ctx *ctx, const uint8_t *src, size_t stride_bytes);
int fgl_write_depth(fgl_ctx *ctx, const double *src);

/* SSAA resolve: resize.Resize(w/factor, h/factor, Image(), resize.Bilinear) as
 * called by every example (examples/teapot.go:60, dragon.go:82, ...).  dst
 * receives (h/factor) rows of (w/factor) RGBA8 (premultiplied, as
 * *image.RGBA).  fgl_resolve_device leaves the result on the device only. */
int fgl_resolve(fgl_ctx *ctx, int factor, uint8_t *dst_rgba8);
int fgl_resolve_device(fgl_ctx *ctx, int factor);
int fgl_read_resolved(fgl_ctx *ctx, uint8_t *dst_rgba8);

/* Sort-last multi-GPU composite (not in the reference).
 * fgl_composite_pack writes one uint64 key per pixel,
 * (depth32 << 32 | R<<24 | G<<16 | B<<8 | A) with depth32 a monotone map of the
 * float64 depth, into device memory `keys_dev` (width*height uint64).  The
 * caller min-reduces the key buffers across ranks (NCCL ncclMin on
 * ncclUint64 / torch.distributed) and hands the result to
 * fgl_composite_unpack, which rewrites this context's colour buffer (and, as
 * float32-quantised values, its depth buffer).  fgl_composite_min merges a
 * second key buffer on the same device (used by the single-GPU test of the
 * composite and by the peer-memory path). */
int fgl_composite_pack(fgl_ctx *ctx, void *keys_dev);
int fgl_composite_unpack(fgl_ctx *ctx, const void *keys_dev);
int fgl_composite_min(fgl_ctx *ctx, void *keys_dev_inout, const void *keys_dev_other, uint64_t count);

/* Peer-memory composite: the exact, fused alternative to pack / all-reduce / unpack.  Every rank
 * exports CUDA IPC handles of its colour and depth buffers (fgl_ipc_export, 64 bytes each), the host
 * exchanges them (torch.distributed / MPI / a Go channel), and every rank opens its peers' handles
 * (fgl_ipc_open).  fgl_composite_peer then runs ONE kernel on this rank that, for this rank's stripe of
 * the frame, loads every rank's float64 depth over NVLink (P2P), selects the nearest -- on a tie the
 * higher rank, i.e. the later triangle range, the reference's `<=` rule (context.go:248) -- and stores
 * the winning depth and colour into ALL ranks' buffers.  color[r]/depth[r] are the device pointers of
 * rank r (this rank's own entries are fgl_color/depth_device_ptr).  The caller must make sure all ranks
 * have finished drawing before any rank calls it, and that all have returned from fgl_sync before
 * anyone draws again (a host barrier each; see fauxgl_b200/multigpu.py).  The result equals a single-GPU
 * render of the whole mesh bit for bit when ReadDepth, WriteDepth are on, DepthBias is 0 and the output is
 * opaque. */
#define FGL_MAX_PEERS 8
#define FGL_IPC_HANDLE_BYTES 64
int fgl_ipc_export(fgl_ctx *ctx, void *color_handle, void *depth_handle);
int fgl_ipc_open(fgl_ctx *ctx, const void *color_handle, const void *depth_handle, void **color_ptr, void **depth_ptr);
int fgl_ipc_close(fgl_ctx *ctx, void *color_ptr, void *depth_ptr);
int fgl_composite_peer(fgl_ctx *ctx, int rank, int nranks, void *const *color, void *const *depth);

/* ---- Sort-last composite inside the library ------------------------------------------------------
 * The reference reduces the RasterizeInfo of its goroutines over a channel (context.go:393-410, 413-433) because
 * they share one address space; ranks on different GPUs have to exchange pixels instead.  Everything below is
 * enqueued on the context's stream and returns without waiting; read-backs, clears and draws issued afterwards are
 * ordered behind it.  Collective: every rank of the group calls it once per frame, in the same order.
 *
 * (1) NCCL.  fgl_comm_unique_id: rank 0 obtains an id (FGL_COMM_ID_BYTES) and hands it to the other ranks by any
 * means (torch.distributed, MPI, a file, a Go channel).  fgl_comm_init: every rank, with its context (collective;
 * blocks until all ranks have joined).  fgl_composite: packs this rank's buffers into keys
 * (depth32 << 32 | R<<24 | G<<16 | B<<8 | A), min-reduces them with ncclReduceScatter(ncclMin, ncclUint64) by screen
 * stripe, and brings the stripes to rank `root` (grouped ncclSend/ncclRecv; only root's buffers then hold the frame)
 * or, with root < 0, to every rank (ncclAllGather).  NCCL is loaded with dlopen (libnccl.so.2, or $FGL_NCCL_LIB) on
 * first use; FGL_E_UNSUPPORTED if it is not installed.  Differences against a single-GPU render: depth ties (the
 * smaller colour wins) and the 32-bit depth key; see fgl_peer_composite for the exact alternative. */
#define FGL_COMM_ID_BYTES 128
typedef struct fgl_comm fgl_comm;
int fgl_comm_unique_id(void *id_out);
int fgl_comm_init(fgl_ctx *ctx, int nranks, int rank, const void *id, fgl_comm **out);
int fgl_comm_destroy(fgl_comm *comm);
int fgl_composite(fgl_ctx *ctx, fgl_comm *comm, int root);
/* Device time per stage of the composites issued while fgl_set_profiling was on, summed since the last call:
 * ms[0] pack, ms[1] reduce-scatter, ms[2] gather / all-gather, ms[3] unpack; *composites = how many. */
int fgl_comm_stage_times(fgl_ctx *ctx, fgl_comm *comm, float ms[4], uint32_t *composites);

/* (2) Peer memory: exact (float64 depth, ties to the higher rank -- the later triangle range, the reference's `<=`
 * rule -- so the result equals a single-GPU render bit for bit), sparse (only strips a rank has drawn into since its
 * last depth clear are read from it) and synchronised on the device (flags in peer memory; no host barrier).
 * fgl_peer_export fills a record of FGL_PEER_EXPORT_BYTES for this context (CUDA IPC handles of its colour, depth,
 * dirty-strip and flag buffers); the host gathers the records of all ranks, in rank order, and passes them to
 * fgl_peer_group_create on every rank.  Ranks may be processes (one per GPU) or contexts of one process (one host
 * thread per GPU).  fgl_peer_composite(ctx, group, root, flags): rank r composites the scanlines y = r (mod nranks) with ONE
 * kernel of P2P loads and stores over NVLink and leaves the result in root's buffers (root >= 0) or in every rank's
 * (root < 0).  fgl_peer_status waits for the stream and reports a rank that never arrived (10 s) as FGL_E_CUDA. */
#define FGL_PEER_EXPORT_BYTES 512
typedef struct fgl_peer_group fgl_peer_group;
int fgl_peer_export(fgl_ctx *ctx, void *record);
int fgl_peer_group_create(fgl_ctx *ctx, int rank, int nranks, const void *records, fgl_peer_group **out);
int fgl_peer_group_destroy(fgl_peer_group *group);
/* flags: FGL_COMPOSITE_COLOR_ONLY -- the target(s) receive the winning colours but not the depths: enough to PRESENT the
 * frame (a third of the bytes into the presenting rank, whose NVLink ingress is what bounds the composite at 8 ranks),
 * not to keep drawing into it (its depth buffer stays as the rank drew it). */
#define FGL_COMPOSITE_COLOR_ONLY 1
int fgl_peer_composite(fgl_ctx *ctx, fgl_peer_group *group, int root, int flags);
/* The same in three steps, for ranks that share a DEVICE (several contexts of one process on one GPU, as in the
 * single-GPU tests): a kernel that waits for a flag can occupy the hardware queue the kernel that sets the flag is
 * submitted to, so the host has to submit phase 1 ("I have drawn") on every rank, then phase 2 (wait for all,
 * composite, "I am done") on every rank, then phase 3 (wait for all) on every rank.  Ranks on different GPUs --
 * processes or threads -- call fgl_peer_composite (= phase 0: all three at once). */
int fgl_peer_composite_phase(fgl_ctx *ctx, fgl_peer_group *group, int root, int flags, int phase);
int fgl_peer_status(fgl_ctx *ctx, fgl_peer_group *group);
/* As fgl_comm_stage_times: ms[0] signal + wait until every rank has drawn (load imbalance shows up here), ms[1]
 * bitmap gather + the composite kernel, ms[2] signal + wait until every rank has finished, ms[3] unused (0). */
int fgl_peer_stage_times(fgl_ctx *ctx, fgl_peer_group *group, float ms[4], uint32_t *composites);

/* Tuning aid: with FGL_TILE_CLOCK=1 in the environment when the context is created, the
 * tile kernel records, per screen tile, its SM cycles and (smid << 32 | segments in its bin);
 * dst receives 2*ntiles uint64 (ntiles = tiles_x * tiles_y of fgl_draw_stats). */
int fgl_debug_tile_cycles(fgl_ctx *ctx, uint64_t *dst, uint64_t ntiles);

/* Self-check of the device's branch-free float64 division (the fused front end divides with the fast path of the
 * compiler's own expansion written as straight-line code, the refinement of a reciprocal shared by the quotients of one
 * denominator; csrc/fgl_math.cuh): `pairs` generated operand pairs -- raw bit patterns, NaN, infinities, subnormals,
 * zeros, rasteriser-sized values -- are divided both ways on the device; *mismatches receives the results that took
 * the helpers' fast path and differ from `a / b` or `1 / b` in any bit (must be 0), *fast_path how many took it. */
int fgl_debug_div_check(fgl_ctx *ctx, uint64_t seed, uint64_t pairs, uint64_t *mismatches, uint64_t *fast_path);

/* Fragment-rate bound of the roofline: the reference resolves every fragment with a locked
 * read-modify-write of DepthBuffer[i] (context.go:245-273); the device primitive that could replace it one
 * fragment at a time is a 64-bit atomicMin on a packed (depth, colour) key.  This measures that primitive on the
 * context's device: `ops` atomicMin on uniformly random words of a buffer of width*height uint64 (the depth buffer's
 * footprint), timed with CUDA events; *ops_per_second receives the rate.  The path itself does not use global
 * atomics per fragment (strips resolve in shared memory); the figure is the denominator of bound (b). */
int fgl_probe_atomic_rate(fgl_ctx *ctx, uint64_t ops, double *ops_per_second);

/* Interop for the host harness (timing with CUDA events on the launching
 * stream; zero-copy views of the buffers).  Clears run on a stream of their own (they overlap the front end of
 * the next draw); every library call that touches the framebuffer orders itself after a pending clear, but work
 * the caller enqueues on fgl_stream() against the raw pointers does not: call fgl_sync() (or any read-back /
 * draw) after the last clear first. */
void *fgl_stream(const fgl_ctx *ctx);            /* cudaStream_t */
void *fgl_color_device_ptr(const fgl_ctx *ctx);  /* width*height*4 bytes */
void *fgl_depth_device_ptr(const fgl_ctx *ctx);  /* width*height doubles */

#ifdef __cplusplus
}
#endif
#endif /* FAUXGL_B200_H */
