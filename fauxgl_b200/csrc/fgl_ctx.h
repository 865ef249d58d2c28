// fgl_ctx.h -- the private definitions behind the opaque handles of include/fauxgl_b200.h, shared by the translation
// units that implement the C ABI (fgl_api.cu: contexts, meshes, draws; fgl_comm.cu: multi-GPU composite).
#pragma once
#include <mutex>
#include <string>

#include "fgl_internal.h"

using fgl::DrawCounters;
using fgl::WorkBuffers;

struct fgl_tex {
    int device;
    uint8_t *pixels;
    int w, h, format;
};

struct fgl_mesh {
    int device;
    uint64_t nt, nl;
    double *tpos, *tnrm, *ttex, *tcol;  // planar, triangles
    double *lpos, *lnrm, *ltex, *lcol;  // planar, lines
    double *staging;                    // AoS landing buffer for H2D copies (kept for fgl_mesh_update)
    size_t staging_elems;
    // streaming uploads (fgl_mesh_update_async): the copy stream and the draw stream hand the
    // buffers back and forth through these two events
    cudaEvent_t ev_uploaded, ev_drawn;
    bool has_events, upload_pending, drawn_recorded;
    // Indexed meshes (fgl_mesh_create_indexed) keep their corner indices and table buffers on the device, so that a
    // re-posed mesh only sends its v / vt / vn tables again (fgl_mesh_update_indexed_async).
    int32_t *corners;                   // [nt][3][3]
    double *tab_v, *tab_vt, *tab_vn;    // [nv][3], [nvt][3], [nvn][3]
    uint64_t nv, nvt, nvn;
    // The context whose streams order this mesh's uploads and in-place edits.  Any context of the same device may
    // DRAW the mesh (read-only); update / transform / smooth-normals must go through the owner, and a mesh that is
    // re-uploaded through the streaming entry point (events above) can only be drawn by its owner.
    const fgl_ctx *owner;
};

struct fgl_fence {
    int device;
    cudaEvent_t done;
    DrawCounters *counters;  // pinned: the frame's accumulated RasterizeInfo / overflow flags
    bool recorded;
};

constexpr int PROF_RING = 32;
constexpr int PROF_EVENTS = 5;  // start, after geometry, after spans, after sort, after tile kernel
struct ProfSlot { cudaEvent_t e[PROF_EVENTS]; };

struct fgl_ctx {
    int device;
    int w, h;
    int tile_w;                        // strip width of this context: 32 or 64 (fgl_internal.h)
    int front_mode;                    // 0 auto, 1 fused front end always, 2 split stages always (FGL_FRONT)
    bool order_coop;                   // the order plumbing as one cooperative kernel (fgl_order.cu); FGL_ORDER=split: separate kernels
    cudaStream_t stream;
    cudaStream_t copy_stream;          // H2D of streaming mesh uploads, overlapping the draw stream
    // Clears run on a stream of their own: the front end of the next draw touches no framebuffer, so the clear
    // of a frame overlaps its k_front instead of preceding it (fb_clear_begin / fb_clear_end / fb_join below).
    cudaStream_t fb_stream;
    // Frame read-back (fgl_frame_end): the colour buffer and the frame's counters are copied into one of two staging
    // slots on the draw stream (device to device, a few microseconds) and go to the host on a stream of their own, so
    // that the PCIe transfer of frame k overlaps the draw of frame k + 1 instead of holding the framebuffer.
    cudaStream_t read_stream;
    uint32_t *rb_color[2];
    DrawCounters *rb_counters[2];
    cudaEvent_t ev_rb_staged[2], ev_rb_done[2];
    bool rb_done_recorded[2], rb_overlap;
    int rb_next;
    cudaEvent_t ev_fb_free, ev_cleared;
    bool clear_overlap, clear_pending;
    std::mutex mu;
    mutable std::mutex err_mu;         // guards err / err_seq only (fail() runs both inside and outside `mu`)
    std::string err;
    unsigned long long err_seq;
    uint32_t *color;
    double *depth;
    uint32_t *resolved;
    int rw, rh;
    WorkBuffers wb;
    DrawCounters *host_counters;       // pinned
    DrawCounters *acc_dev;             // async accumulation (total/updated/overflow)
    bool async_pending;
    bool capturing;                    // between fgl_graph_begin and fgl_graph_end: the stream records instead of running
    unsigned long long buffer_epoch;   // bumped whenever a device buffer recorded launches may point at is reallocated
    bool counters_clean;               // the last draw's k_shade zeroed the device-side draw counters
    unsigned long long *prim_info;     // per-primitive RasterizeInfo of fgl_draw_*_each, [prim_info_cap][2]
    uint64_t prim_info_cap;
    unsigned long long *scratch;       // 8 words: reductions of fgl_mesh_bounds / fgl_depth_image
    uint16_t *gray16;                  // DepthImage staging, allocated on first use
    double clear_depth_value;          // what the last fgl_clear_depth wrote (NaN once the buffer was uploaded by the caller)
    uint32_t clear_color_value;        // what the last fgl_clear_color wrote ...
    bool clear_color_known;            // ... if nothing but draws has touched the colour buffer since
    unsigned long long *peer_flags;    // fgl_comm.cu: ready / done epochs written by the peers of a peer group (device)
    unsigned long long peer_epoch;     // last composite epoch this context took part in
    fgl_draw_stats stats;
    // per-stage profiling (fgl_set_profiling)
    bool profiling;
    ProfSlot prof[PROF_RING];
    int prof_used;          // slots recorded since the last drain
    bool prof_created;
    fgl_stage_times prof_acc;
};

struct fgl_graph {
    int device;
    fgl_ctx *ctx;
    cudaGraph_t graph;
    cudaGraphExec_t exec;
    unsigned long long buffer_epoch;  // the context's buffer_epoch when it was recorded: a replay after a reallocation is refused
    bool has_draws;        // launching it leaves RasterizeInfo to collect (fgl_sync / fgl_frame_end)
    bool counters_clean;   // state of the draw counters after the recorded frame
};

namespace fgl {
// Record an error message (thread-locally and in the context, see fgl_api.cu) and return `code`.
int api_fail(fgl_ctx *ctx, int code, const char *fmt, ...);
// cudaSetDevice(ctx->device); FGL_OK or an error code.
int api_check_ctx(fgl_ctx *ctx);
// Order the draw stream after a pending clear (clears run on a stream of their own).
void api_fb_join(fgl_ctx *ctx);
}  // namespace fgl
