// fgl_internal.h -- device-side data layout and kernel launch interfaces shared
// by the translation units of libfauxgl_b200.so.  Not part of the C ABI.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/fauxgl_b200.h"

namespace fgl {

// The library is written for B200 (sm_100a, 148 SMs): grid-stride kernels are launched with one wave of 8 CTAs per SM.
// (Correct on any device -- the loops cover the data whatever the grid -- but sized for this one.)
constexpr unsigned B200_SMS = 148u;
constexpr unsigned GRID_WAVE = B200_SMS * 8u;

// ---- screen strips ------------------------------------------------------------
// The framebuffer is binned into strips of 64 x 1 pixels.  The reference walks each scanline
// left to right with forward differencing (context.go:207-213); a covered run is cut into one
// segment per strip it crosses, so wide strips mean few segments, and a segment never spans
// rows, so one-row strips add none.  One warp resolves a strip (fgl_raster.cu): a 64-bit mask
// per lane describes the pixels a segment still has to write.  TILE_W is the widest strip; a context
// uses 32-pixel strips up to 4 Mpixel (the heaviest strips, which bound the strip kernel's duration,
// halve; measured 84 -> 77 us at 1920x1080) and 64-pixel strips above (fewer segments and strips;
// measured 408 vs 442 us at 7680x4320).
#ifndef FGL_TILE_W
#define FGL_TILE_W 64
#endif
constexpr int TILE_W = FGL_TILE_W;

// ---- device mesh: planar SoA ----------------------------------------------------
// plane(attr, v, c)[i] = attr_base[(v * ncomp + c) * n + i]; a warp reading one
// component of 32 consecutive primitives touches 256 contiguous bytes.
struct MeshPlanes {
    const double *pos;  // 3 comps
    const double *nrm;  // 3 comps
    const double *tex;  // 2 comps (Texture.X, Texture.Y; .Z is never read by a built-in shader)
    const double *col;  // 4 comps
    uint32_t n;         // primitives in the mesh (plane stride)
    uint32_t nverts;    // 3 for triangles, 2 for lines
};

// ---- raster record: one screen-space triangle with the per-triangle setup of
// Context.rasterize (context.go:155-181) done once by the geometry stage.
struct __align__(16) Rec {
    double s[9];             // s0.xyz s1.xyz s2.xyz (screen space)
    double w00, w01, w02;    // edge functions at the first pixel centre (x0+.5, y0+.5), context.go:163-166
    double ra;               // 1 / edge(s0,s1,s2), context.go:175
    double ra12, ra20, ra01; // context.go:179-181
    double r0, r1, r2;       // 1 / Output.W, context.go:176-178
    uint32_t src;            // primitive index in the mesh planes, or clip-pool triangle index
    uint32_t flags;          // REC_* below
    int32_t x0, x1;          // integer bounding box, context.go:155-160 (saturated to int32)
    int32_t y0, y1;
};
static_assert(sizeof(Rec) == 176, "Rec layout");

constexpr uint32_t REC_VMAP_MASK = 0x3f;    // 3 x 2 bits: source vertex of v0,v1,v2
constexpr uint32_t REC_SRC_POOL = 1u << 6;  // src indexes the clip pool
constexpr uint32_t REC_WRAP = 1u << 7;      // (segments only) the pixels lie outside [0, width) of the row they were covered in and
                                            // alias into a neighbouring row (context.go:223-228): depth and blended colour are
                                            // written there, an opaque colour is dropped (SetNRGBA's bounds check, :269)

// Clip pool: vertices produced by ClipTriangle (clipping.go:54-74), AoS.
struct ClipVertex {
    double pos[3], nrm[3], tex[2], col[4];
};
struct ClipTri {
    ClipVertex v[3];
    uint32_t prim, _pad;  // mesh primitive the triangle was clipped from (per-primitive RasterizeInfo)
};

// ---- span segment: the covered pixels of one scanline of one triangle inside one
// tile, with the forward-differenced edge values at its first pixel.
struct __align__(16) Seg {
    double w0, w1, w2;  // w0,w1,w2 of context.go:208-213 at pixel x (before that pixel's increment)
    uint32_t rec;       // record index
    uint16_t x;         // first covered pixel (absolute column)
    uint8_t yt;         // 1: the segment aliases into a neighbouring row (REC_WRAP of its SegV)
    uint8_t cnt;        // covered pixels (1..TILE_W)
};
static_assert(sizeof(Seg) == 32, "Seg layout");

// ---- binned segment: a Seg plus every field of its record the back end needs (depth phase, perspective
// weights, attribute source), so that neither the strip kernel nor the shading kernel chases the record.
// 128 bytes: one cache line per segment.  The first 96 bytes (SegHead: three sectors) are all the depth phase of
// the strip kernel reads; the perspective weights behind them are for shading only.
struct __align__(16) SegHead {
    double w0, w1, w2;
    double ra, z0, z1, z2;   // 1/area and the three screen depths: z = (b0*z0 + b1*z1) + b2*z2
    double a12, a20, a01;    // per-pixel increments of w0, w1, w2 (context.go:167-172)
    uint16_t x;              // first covered pixel (absolute column)
    uint8_t yt, cnt;         // (unused), covered pixels (1..TILE_W)
    uint32_t src;            // primitive index in the mesh planes, or clip-pool triangle index
    uint32_t flags;          // REC_* (vertex map, pool bit)
    uint32_t _pad0;
};
static_assert(sizeof(SegHead) == 96, "SegHead layout");
struct __align__(16) SegV : SegHead {
    double r0, r1, r2;       // 1 / Output.W of the three vertices, context.go:176-178
    uint32_t _pad[2];
};
static_assert(sizeof(SegV) == 128, "SegV layout");

// Strip work queue (device): k_tile_ranges lists the busy strips, warps of the strip kernel pull them.
struct TileCtl {
    uint32_t nheavy, nlight;  // busy_list[0 .. nheavy) = heavy strips, busy_list[ntiles-1 .. ntiles-nlight] = the others
    uint32_t head, _pad;
    uint32_t bar_count, bar_gen;  // grid-wide barrier of k_order (fgl_order.cu): arrivals of the round, rounds completed
};
// Fused front end (fgl_geom.cu): primitives per k_front block, and blocks per group sum (wb.blk_base[g], added up by
// k_seg_index / k_order).
#ifndef FGL_FRONT_FT
#define FGL_FRONT_FT 128
#endif
constexpr int FRONT_FT = FGL_FRONT_FT;
constexpr uint32_t FRONT_GROUP = 64;
#ifndef FGL_HEAVY_SEGS
#define FGL_HEAVY_SEGS 96
#endif
constexpr uint32_t HEAVY_SEGS = FGL_HEAVY_SEGS;  // a strip with at least this many segments is queued first

// ---- per-draw device constants ------------------------------------------------------
struct DrawParams {
    fgl_state state;
    // shader (fgl_shader with the texture resolved)
    int32_t kind, has_texture;
    double matrix[16], light[3], camera[3], object[4], ambient[4], diffuse[4], specular[4];
    double specular_power, color[4];
    const uint8_t *tex;
    int32_t tex_w, tex_h, tex_format, object_is_discard;
    // framebuffer
    int32_t width, height, tiles_x, tiles_y;  // strips per row (ceil(width / tile_w)), strip rows (= height)
    int32_t tile_w, tile_shift;               // strip width of this context: 32 or 64 pixels (<= TILE_W), and its log2
    double screen[16];  // Screen(w,h), matrix.go:119-128
    // input
    MeshPlanes mesh;
    uint32_t first, count;  // primitive range
    int32_t is_lines;
    int32_t deferred;       // 1: shading cannot discard or blend -> resolve depth in order, shade final winners
    // per-primitive RasterizeInfo (fgl_draw_*_each): [count][2] = (TotalPixels, UpdatedPixels); null otherwise
    unsigned long long *prim_info;
};

// Device-side counters/results of one draw (also copied to pinned host memory).
struct DrawCounters {
    unsigned long long total_pixels, updated_pixels;
    unsigned int n_records, n_rows, n_segs, n_clip;
    unsigned int overflow;  // bit0 records, bit1 rows, bit2 clip pool, bit3 segments
    unsigned int need_records, need_rows, need_segs, need_clip;
    unsigned int shade_done;   // CTAs of k_shade that have finished (the last one accumulates the frame's counters)
    unsigned int rec_cursor;   // geometry stage: next free record slot (regions are handed out per block)
    unsigned int blocks_done;  // geometry stage: blocks that have published their aggregate
    unsigned long long seg_cursor;  // fused front end: next free segment slot (regions are handed out per block)
};
constexpr unsigned OVF_RECORDS = 1u, OVF_ROWS = 2u, OVF_CLIP = 4u, OVF_SEGS = 8u;

struct WorkBuffers {
    // geometry
    // The geometry kernel never waits for other blocks: a block takes a region of `recs` with one atomic
    // (slots are in block-arrival order), publishes (scanlines << 30 | records), and the last block to
    // finish scans the aggregates; k_rec_index then lists the slots in primitive order.
    // (k_front uses the same three arrays with one entry per WARP, 32 primitives: segments << 32 | records,
    // the ordered position of the warp's first segment, the first SEGMENT slot of the warp.)
    unsigned long long *blk_agg;    // [cap_prims/32+8] scanlines << 30 | records of each geometry block
    unsigned long long *blk_base;   // [cap_prims/32+8] exclusive prefix of blk_agg
    uint32_t *blk_region;           // [cap_prims/32+8] first record slot of the block
    uint4 *blk_wcnt, *blk_woff;     // [cap_prims/32+8] k_front: segments stored by each of the block's four warps and the
                                    //                  offset of each warp's run inside the block's region
    Rec *recs;                // [cap_records]   indexed by slot
    uint32_t *rec_local_row;  // [cap_records]   by slot: scanline offset of the record inside its block
    uint32_t *rec_slot;       // [cap_records+1] by primitive order: slot of the record
    uint32_t *rec_row_off;    // [cap_records+1] by primitive order: first (record, scanline) item; [n] = total
    ClipTri *clip_pool;       // [cap_clip]
    // spans
    uint32_t *row_nseg;       // [cap_rows]   segments of each (record, scanline)
    uint32_t *row_seg_off;    // [cap_rows+1] exclusive scan
    Seg *row_first;           // [cap_rows]   first segment of the scanline, kept by the count pass
    uint32_t *row_key;        // [cap_rows]   its tile id
    // binning: stable sort of segment indices by tile
    uint32_t *seg_key[2];     // [cap_segs] tile id (ping-pong for the radix sort)
    uint32_t *seg_val[2];     // [cap_segs] segment index
    SegV *segv;               // [cap_segs]  segments in (record, scanline, column) order; bins index into it
    uint2 *busy_list;         // [ntiles]    (strip, first sorted position of its bin) of every strip with a non-empty bin:
                              //             heavy ones from the front, the rest from the back
    uint32_t nsm;             // SMs of the device
    TileCtl *tile_ctl;        // device
    uint32_t *vis_seg;        // [ntiles*TILE_W] deferred shading, indexed by busy-LIST position: per pixel of a busy strip, the segment (index into
                              //                 segv) whose fragment won, or 0xffffffff (allocated on the first deferred draw)
    uint8_t *dirty;           // [ntiles] 1: some pixel of the strip had its depth written since the last depth clear (the
                              //          sparse sort-last composite only exchanges such strips, fgl_comm.cu)
    unsigned long long *tile_clock;  // [ntiles][2] (cycles, smid<<32|segments) of the last k_strip launch; null unless FGL_TILE_CLOCK=1
    uint32_t *scan_tmp;       // block sums for scans / radix histograms
    DrawCounters *counters;   // device
    uint32_t cap_prims, cap_records, cap_rows, cap_segs, cap_clip, ntiles, scan_tmp_words;  // ntiles = strips = tiles_x * height
};

#ifdef __CUDACC__
// ---- programmatic dependent launch ------------------------------------------------------------------
// The kernels of one draw are short (3-130 us) and strictly ordered; launched with the programmatic-serialisation
// attribute, the next kernel's CTAs are scheduled while the last wave of its predecessor drains and block in
// pdl_wait() until that grid has completed and flushed -- the launch latency and ramp of ~10 kernels per frame
// overlap instead of adding up.  Rules: a kernel launched through launch_pdl() calls pdl_wait() before it reads
// or writes anything and before any early return (completion must stay transitive along the chain); pdl_trigger()
// at the top lets the successor be scheduled as soon as every CTA of this grid is resident.  Both are no-ops
// in a kernel that was launched without the attribute.  FGL_PDL=0 turns the attribute off (tuning aid).
extern bool g_pdl;
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#ifndef FGL_PDL_TRIGGER
#define FGL_PDL_TRIGGER 0
#endif
__device__ __forceinline__ void pdl_trigger() {
#if FGL_PDL_TRIGGER
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <class... P, class... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, A &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr; cfg.numAttrs = g_pdl ? 1u : 0u;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<P>(args)...);
}

// Index (relative to the draw's first primitive) of the mesh primitive a record / segment came from.
__device__ __forceinline__ uint32_t src_primitive(const WorkBuffers &wb, const DrawParams &p, uint32_t src, uint32_t flags) {
    return ((flags & REC_SRC_POOL) ? wb.clip_pool[src].prim : src) - p.first;
}
__device__ __forceinline__ uint32_t rec_primitive(const WorkBuffers &wb, const DrawParams &p, uint32_t rec) {
    const Rec *rp = wb.recs + rec;
    return src_primitive(wb, p, rp->src, rp->flags);
}
#endif

// ---- kernel launchers (each returns the number of kernels it launched) --------------
int launch_mesh_ingest(const double *aos, double *planes, uint32_t n, int nverts, int ncomp_in, int ncomp_out,
                       cudaStream_t st);
int launch_mesh_export(const double *planes, double *aos, uint32_t n, int nverts, int ncomp, cudaStream_t st);
int launch_mesh_transform(double *pos, double *nrm, uint32_t n, int nverts, const double m[16], cudaStream_t st);

int launch_clear_color(uint32_t *color, size_t npix, uint32_t rgba, cudaStream_t st);
int launch_clear_depth(double *depth, size_t npix, double v, cudaStream_t st);

// exclusive scan of in[0..n) into out[0..n], out[n] = total; n read from *n_dev if n_dev != nullptr
// (bounded by n_max).  tmp needs >= scan_tmp_words(n_max) words.
size_t scan_tmp_words(uint32_t n_max);
// Where the scan publishes its total (all pointers device, any may be null): count/need = total,
// *overflow |= bit when total > cap.  clip_*: the clip-pool counters are finalised on the way.
struct ScanSink {
    unsigned int *count, *need, *overflow;
    unsigned int cap, bit;
    const unsigned int *clip_n;
    unsigned int *clip_need;
    unsigned int clip_cap;
};
int launch_exclusive_scan(const uint32_t *in, uint32_t *out, uint32_t n_max, const unsigned int *n_dev,
                          uint32_t *tmp, const ScanSink &sink, cudaStream_t st);
// stable LSD radix sort of (key,val) pairs on `bits` key bits; returns launches, *sorted_buf = buffer index
// skip_if (nullable): the draw's overflow word -- the keys of an overflowed draw were never written, sort nothing
int launch_sort_pairs(uint32_t *const key[2], uint32_t *const val[2], const unsigned int *n_dev, uint32_t n_max,
                      int bits, uint32_t *tmp, int *sorted_buf, cudaStream_t st, const unsigned int *skip_if = nullptr);

// Binning by strip for 8 < bits <= 20: one global radix pass on the low eight bits, then one CTA per bucket groups its
// pairs by the remaining bits and lists its busy strips (fgl_scan_sort.cu).  The bins come out grouped, not in
// ascending strip order -- the back end only needs them contiguous.  g_bin_buckets: FGL_BIN=lsd turns it off.
extern bool g_bin_buckets;
bool bin_buckets_ok(int bits);
int launch_bin_buckets(uint32_t *const key[2], uint32_t *const val[2], DrawCounters *ctr, uint32_t n_max, int bits,
                       uint32_t *tmp, int *sorted_buf, uint2 *busy_list, uint32_t ntiles, TileCtl *ctl,
                       unsigned long long *group_sums, uint32_t ngroups, cudaStream_t st);

// counters_clean: the previous draw's last kernel left the draw counters zeroed (no memset node needed)
int launch_geometry(const DrawParams &p, const WorkBuffers &wb, bool counters_clean, cudaStream_t st);
// fused geometry + span stage of large draws: k_front, then k_seg_index leaves (seg_key[0], seg_val[0]) in
// primitive order, like launch_geometry + launch_spans
int launch_front(const DrawParams &p, const WorkBuffers &wb, bool counters_clean, cudaStream_t st);
int launch_seg_index(const DrawParams &p, const WorkBuffers &wb, cudaStream_t st);
int launch_spans(const DrawParams &p, const WorkBuffers &wb, int *sorted_buf, cudaStream_t st);
int launch_bin(const DrawParams &p, const WorkBuffers &wb, int *sorted_buf, cudaStream_t st);
// launch_seg_index (fused front end) + launch_bin as ONE cooperative kernel with grid-wide barriers (fgl_order.cu);
// order_supported: the device can run it (cooperative launch, one 1024-thread CTA per SM)
bool order_supported(int device, uint32_t nsm);
int launch_order(const DrawParams &p, const WorkBuffers &wb, bool fused, int *sorted_buf, cudaStream_t st);
// acc (nullable): counters of the frame's async draws; *accumulated tells whether the draw's counters were added
// to it by the last kernel (k_shade's last CTA) -- otherwise the caller launches k_accumulate
int launch_raster(const DrawParams &p, const WorkBuffers &wb, int sorted_buf, uint32_t *color, double *depth,
                  DrawCounters *acc, bool *accumulated, cudaStream_t st);
// acc += cur (total/updated pixels, overflow, needs): one thread.  (k_shade's last CTA then zeroes *cur for the next
// draw, which saves that draw its memset node.)
__device__ __forceinline__ void accumulate_counters(const DrawCounters *cur, DrawCounters *acc) {
    acc->total_pixels += cur->total_pixels;
    acc->updated_pixels += cur->updated_pixels;
    acc->overflow |= cur->overflow;
    acc->need_records = max(acc->need_records, cur->need_records);
    acc->need_rows = max(acc->need_rows, cur->need_rows);
    if (!(cur->overflow & (OVF_RECORDS | OVF_ROWS))) acc->need_segs = max(acc->need_segs, cur->need_segs);  // (else: not computed)
    acc->need_clip = max(acc->need_clip, cur->need_clip);
}

// Mesh.SmoothNormals, mesh.go:105-120: (hash, corner) pairs of the 3n corners, then -- on the pairs sorted by hash --
// the group sums in corner order, written to every member
int launch_corner_hash(const double *pos, uint32_t n, uint32_t *keys, uint32_t *vals, cudaStream_t st);
int launch_smooth_groups(const uint32_t *keys, const uint32_t *vals, const double *pos, double *nrm, uint32_t n,
                         cudaStream_t st);
// Mesh.SmoothNormalsThreshold: nrm_out (planes, separate from nrm_in: every corner reads all normals of its group)
int launch_smooth_threshold(const uint32_t *keys, const uint32_t *vals, const double *pos, const double *nrm_in,
                            double *nrm_out, uint32_t n, double threshold, cudaStream_t st);
// indexed (OBJ-shaped) mesh -> position / normal / texture planes, obj.go:58-74 + triangle.go:46-58
int launch_indexed_ingest(const double *v, const double *vt, const double *vn, const int32_t *corners, double *pos,
                          double *nrm, double *tex, uint32_t n, cudaStream_t st);
// binary STL records (50 B each) -> position / normal planes, stl.go:86-154
int launch_stl_ingest(const uint8_t *records, double *pos, double *nrm, uint32_t n, cudaStream_t st);
// bounding box of position planes: bounds[0..2] = ordered-u64 min, [3..5] = max (see fgl_post.cu)
int launch_mesh_bounds(const double *pos, uint32_t n, int nverts, unsigned long long *bounds, cudaStream_t st);
// DepthImage, context.go:87-117; scratch = 2 x u64
int launch_depth_image(const double *depth, size_t npix, uint16_t *out, unsigned long long *scratch, cudaStream_t st);
int launch_resolve(const uint32_t *src, int sw, int sh, uint32_t *dst, int factor, cudaStream_t st);
int launch_composite_pack(const uint32_t *color, const double *depth, unsigned long long *keys, size_t npix,
                          cudaStream_t st, bool bias = true);
int launch_composite_unpack(uint32_t *color, double *depth, const unsigned long long *keys, size_t npix,
                            cudaStream_t st, bool bias = true);
int launch_composite_min(unsigned long long *inout, const unsigned long long *other, size_t n, cudaStream_t st);
// `ops` 64-bit atomicMin on pseudo-random words of buf[0..words) (fgl_probe_atomic_rate)
int launch_atomic_probe(unsigned long long *buf, size_t words, unsigned long long ops, cudaStream_t st);
// the division helpers of fgl_math.cuh against the operator over `pairs` generated operand pairs: out[0] += mismatches,
// out[1] += results that took the helpers' fast path (fgl_debug_div_check)
int launch_div_check(unsigned long long seed, unsigned long long pairs, unsigned long long *out, cudaStream_t st);
int launch_composite_peer(uint32_t *const *color, double *const *depth, int nranks, size_t px0, size_t px1,
                          cudaStream_t st);

}  // namespace fgl
