// fgl_api.cu -- the C ABI of include/fauxgl_b200.h: contexts, meshes, textures,
// draw calls, read-back.  Host-side only; the kernels live in the other
// translation units.  There is no CPU rendering path in this library: every
// entry point either runs CUDA kernels or fails with an fgl_status.
#include <algorithm>
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>

#include "fgl_ctx.h"

using namespace fgl;

namespace fgl { bool g_pdl = true; bool g_bin_buckets = true; }

namespace {

// Error messages.  A failing call stores its message twice: thread-locally (a C/C++/ctypes caller asks right away
// on the same thread) and, when a context is involved, in the context under a small mutex of its own (goroutines
// migrate between OS threads from one cgo call to the next, so a Go caller can only rely on the context's copy).
// Both carry a global sequence number; fgl_last_error(ctx) returns the NEWER of the two, copied into a
// thread-local buffer, so the returned pointer is never invalidated by another thread.
thread_local std::string t_last_error;
thread_local unsigned long long t_last_seq = 0;
thread_local std::string t_error_out;
std::atomic<unsigned long long> g_error_seq{0};

int fail(fgl_ctx *ctx, int code, const char *fmt, ...);

#define CK(ctx, expr)                                                                              \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            cudaGetLastError();                                                                    \
            return fail(ctx, _e == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "%s: %s", #expr, \
                        cudaGetErrorString(_e));                                                   \
        }                                                                                          \
    } while (0)

}  // namespace

namespace {

int fail(fgl_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    const unsigned long long seq = ++g_error_seq;
    t_last_error = buf;
    t_last_seq = seq;
    if (ctx) {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        ctx->err = buf;
        ctx->err_seq = seq;
    }
    return code;
}

// A clear waits for everything queued on the draw stream so far (it may read or write the framebuffer) ...
cudaStream_t fb_clear_begin(fgl_ctx *c) {
    if (!c->clear_overlap) return c->stream;
    if (!c->clear_pending) {  // (still pending: nothing that touches the framebuffer was queued since the last clear)
        cudaEventRecord(c->ev_fb_free, c->stream);
        cudaStreamWaitEvent(c->fb_stream, c->ev_fb_free, 0);
    }
    return c->fb_stream;
}
void fb_clear_end(fgl_ctx *c) {
    if (!c->clear_overlap) return;
    cudaEventRecord(c->ev_cleared, c->fb_stream);
    c->clear_pending = true;
}
// ... and whatever touches the framebuffer on the draw stream next waits for the clear.
void fb_join(fgl_ctx *c) {
    if (c->clear_pending) {
        cudaStreamWaitEvent(c->stream, c->ev_cleared, 0);
        c->clear_pending = false;
    }
}

template <class T>
cudaError_t dev_alloc(T **p, size_t n) {
    *p = nullptr;
    if (n == 0) n = 1;
    return cudaMalloc(reinterpret_cast<void **>(p), n * sizeof(T));
}
template <class T>
void dev_free(T *&p) {
    if (p) cudaFree(p);
    p = nullptr;
}

void free_work(WorkBuffers &wb) {
    dev_free(wb.blk_agg); dev_free(wb.blk_base); dev_free(wb.blk_region); dev_free(wb.blk_wcnt); dev_free(wb.blk_woff);
    dev_free(wb.recs);
    dev_free(wb.rec_local_row); dev_free(wb.rec_slot);
    dev_free(wb.rec_row_off); dev_free(wb.clip_pool); dev_free(wb.row_nseg); dev_free(wb.row_seg_off);
    dev_free(wb.row_first); dev_free(wb.row_key); dev_free(wb.segv);
    dev_free(wb.seg_key[0]); dev_free(wb.seg_key[1]); dev_free(wb.seg_val[0]); dev_free(wb.seg_val[1]);
    dev_free(wb.scan_tmp);
    wb.cap_prims = wb.cap_records = wb.cap_rows = wb.cap_segs = wb.cap_clip = 0;
    wb.scan_tmp_words = 0;
}

struct Caps { uint64_t prims, records, rows, segs, clip; };

int ensure_work_impl(fgl_ctx *c, const Caps &want);
// (Re)allocate work buffers so that they hold at least the given capacities; a reallocation invalidates recorded graphs.
int ensure_work(fgl_ctx *c, const Caps &want) {
    const WorkBuffers before = c->wb;
    const int rc = ensure_work_impl(c, want);
    const WorkBuffers &wb = c->wb;
    if (before.segv != wb.segv || before.recs != wb.recs || before.row_first != wb.row_first || before.clip_pool != wb.clip_pool ||
        before.blk_agg != wb.blk_agg || before.scan_tmp != wb.scan_tmp || before.seg_key[0] != wb.seg_key[0])
        c->buffer_epoch++;
    return rc;
}
int ensure_work_impl(fgl_ctx *c, const Caps &want) {
    WorkBuffers &wb = c->wb;
    const uint64_t LIM = 0xfffffff0ull;
    if (want.records >= (1ull << 30)) return fail(c, FGL_E_INVALID, "draw too large: more than 2^30 raster records");
    if (want.prims > LIM || want.records > LIM || want.rows > LIM || want.segs > LIM || want.clip > LIM)
        return fail(c, FGL_E_INVALID, "draw too large for 32-bit work indices");
    if (want.prims > wb.cap_prims) {
        dev_free(wb.blk_agg); dev_free(wb.blk_base); dev_free(wb.blk_region); dev_free(wb.blk_wcnt); dev_free(wb.blk_woff);
        CK(c, dev_alloc(&wb.blk_wcnt, want.prims / 32 + 8));
        CK(c, dev_alloc(&wb.blk_woff, want.prims / 32 + 8));
        CK(c, dev_alloc(&wb.blk_agg, want.prims / 32 + 8));
        CK(c, dev_alloc(&wb.blk_base, want.prims / 32 + 8));
        // (its first entries double as the group sums of the fused front end, which every draw leaves zeroed)
        CK(c, cudaMemsetAsync(wb.blk_base, 0, sizeof(unsigned long long) * (want.prims / 32 + 8), c->stream));
        CK(c, dev_alloc(&wb.blk_region, want.prims / 32 + 8));
        wb.cap_prims = (uint32_t)want.prims;
    }
    if (want.records > wb.cap_records) {
        dev_free(wb.recs); dev_free(wb.rec_row_off); dev_free(wb.rec_local_row); dev_free(wb.rec_slot);
        CK(c, dev_alloc(&wb.rec_local_row, want.records));
        CK(c, dev_alloc(&wb.rec_slot, want.records + 1));
        CK(c, dev_alloc(&wb.recs, want.records));
        CK(c, dev_alloc(&wb.rec_row_off, want.records + 1));
        wb.cap_records = (uint32_t)want.records;
    }
    if (want.rows > wb.cap_rows) {
        dev_free(wb.row_nseg); dev_free(wb.row_seg_off); dev_free(wb.row_first); dev_free(wb.row_key);
        CK(c, dev_alloc(&wb.row_first, want.rows));
        CK(c, dev_alloc(&wb.row_key, want.rows));
        CK(c, dev_alloc(&wb.row_nseg, want.rows));
        CK(c, dev_alloc(&wb.row_seg_off, want.rows + 1));
        wb.cap_rows = (uint32_t)want.rows;
    }
    if (want.segs > wb.cap_segs) {
        dev_free(wb.segv);
        CK(c, dev_alloc(&wb.segv, want.segs));
        for (int k = 0; k < 2; k++) {
            dev_free(wb.seg_key[k]); dev_free(wb.seg_val[k]);
            CK(c, dev_alloc(&wb.seg_key[k], want.segs));
            CK(c, dev_alloc(&wb.seg_val[k], want.segs));
        }
        wb.cap_segs = (uint32_t)want.segs;
    }
    if (want.clip > wb.cap_clip) {
        dev_free(wb.clip_pool);
        CK(c, dev_alloc(&wb.clip_pool, want.clip));
        wb.cap_clip = (uint32_t)want.clip;
    }
    const size_t words = std::max(std::max(scan_tmp_words(wb.cap_prims), scan_tmp_words(wb.cap_records)),
                                  scan_tmp_words(wb.cap_rows));
    if (words > wb.scan_tmp_words) {
        dev_free(wb.scan_tmp);
        CK(c, dev_alloc(&wb.scan_tmp, words));
        wb.scan_tmp_words = (uint32_t)words;
    }
    return FGL_OK;
}

Caps grown(const fgl_ctx *c, const DrawCounters &hc) {
    auto g = [](uint32_t cap, uint32_t need) { return std::max<uint64_t>(cap, (uint64_t)need + need / 4 + 1024); };
    // The stages run in sequence and stop at the first buffer that is too small: when the records or the scanlines did
    // not fit, the span stage never ran and the segment count behind it is a sum over unwritten memory (found with
    // compute-sanitizer, whose fill pattern turned that sum into a 500 GB request) -- the segment need is only
    // trusted when everything in front of it fitted; the next attempt reports it.
    const bool segs_known = !(hc.overflow & (OVF_RECORDS | OVF_ROWS));
    return Caps{c->wb.cap_prims, g(c->wb.cap_records, hc.need_records), g(c->wb.cap_rows, hc.need_rows),
                segs_known ? g(c->wb.cap_segs, hc.need_segs) : c->wb.cap_segs, g(c->wb.cap_clip, hc.need_clip)};
}

// Entry points that wait for the stream (or allocate / free buffers recorded launches may use) are errors while a
// graph is being recorded: a synchronisation would invalidate the capture.
#define NOT_WHILE_RECORDING(c, what)                                                                              \
    do {                                                                                                          \
        if ((c)->capturing) return fail(c, FGL_E_INVALID, what " waits for the device and cannot be recorded into a graph"); \
    } while (0)

int check_ctx(fgl_ctx *c) {
    if (!c) return fail(nullptr, FGL_E_INVALID, "null context");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(c, FGL_E_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    return FGL_OK;
}

// A mesh whose last upload went through the copy stream is handed to the draw stream here; after the
// draw stream used it, `release` lets the next streaming upload know when it may overwrite the buffers.
void mesh_acquire(fgl_ctx *c, const fgl_mesh *cm) {
    fgl_mesh *m = const_cast<fgl_mesh *>(cm);
    if (m->upload_pending && !c->capturing) {
        cudaStreamWaitEvent(c->stream, m->ev_uploaded, 0);
        m->upload_pending = false;
    }
}
void mesh_release(fgl_ctx *c, const fgl_mesh *cm) {
    fgl_mesh *m = const_cast<fgl_mesh *>(cm);
    if (m->has_events && !c->capturing) {
        cudaEventRecord(m->ev_drawn, c->stream);
        m->drawn_recorded = true;
    }
}

__global__ void k_accumulate(const DrawCounters *cur, DrawCounters *acc) {
    pdl_wait();
    accumulate_counters(cur, acc);
}

int build_params(fgl_ctx *c, const fgl_state *state, const fgl_shader *sh, const fgl_mesh *mesh, uint64_t first,
                 uint64_t count, bool lines, DrawParams *p) {
    if (!state || !sh || !mesh) return fail(c, FGL_E_INVALID, "null state/shader/mesh");
    if (mesh->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on device %d, context on %d", mesh->device, c->device);
    if (mesh->owner != c && mesh->has_events)
        return fail(c, FGL_E_INVALID, "a mesh that is re-uploaded with fgl_mesh_update_async can only be drawn by the context that owns it");
    const uint64_t n = lines ? mesh->nl : mesh->nt;
    if (first > n || count > n - first) return fail(c, FGL_E_INVALID, "primitive range [%llu,+%llu) outside mesh of %llu",
                                                    (unsigned long long)first, (unsigned long long)count, (unsigned long long)n);
    if (sh->kind != FGL_SHADER_SOLID && sh->kind != FGL_SHADER_TEXTURE && sh->kind != FGL_SHADER_PHONG)
        return fail(c, FGL_E_UNSUPPORTED, "shader kind %d has no device implementation (no CPU fallback)", sh->kind);
    if (sh->kind == FGL_SHADER_TEXTURE && !sh->texture)
        return fail(c, FGL_E_INVALID, "TextureShader without a texture");
    if (sh->texture && sh->texture->device != c->device) return fail(c, FGL_E_INVALID, "texture lives on another device");
    if (state->cull < FGL_CULL_NONE || state->cull > FGL_CULL_BACK || state->front_face < FGL_FACE_CW ||
        state->front_face > FGL_FACE_CCW)
        return fail(c, FGL_E_INVALID, "bad cull/front_face");
    memset(p, 0, sizeof *p);
    p->state = *state;
    p->kind = sh->kind;
    memcpy(p->matrix, sh->matrix, sizeof p->matrix);
    memcpy(p->light, sh->light, sizeof p->light);
    memcpy(p->camera, sh->camera, sizeof p->camera);
    memcpy(p->object, sh->object, sizeof p->object);
    memcpy(p->ambient, sh->ambient, sizeof p->ambient);
    memcpy(p->diffuse, sh->diffuse, sizeof p->diffuse);
    memcpy(p->specular, sh->specular, sizeof p->specular);
    p->specular_power = sh->specular_power;
    memcpy(p->color, sh->color, sizeof p->color);
    p->object_is_discard = (sh->object[0] == 0 && sh->object[1] == 0 && sh->object[2] == 0 && sh->object[3] == 0);
    const bool uses_tex = sh->kind == FGL_SHADER_TEXTURE || (sh->kind == FGL_SHADER_PHONG && sh->texture);
    if (uses_tex) {
        p->has_texture = 1;
        p->tex = sh->texture->pixels;
        p->tex_w = sh->texture->w; p->tex_h = sh->texture->h; p->tex_format = sh->texture->format;
    }
    p->width = c->w; p->height = c->h;
    p->tile_w = c->tile_w; p->tile_shift = c->tile_w == 64 ? 6 : 5;
    p->tiles_x = (c->w + c->tile_w - 1) / c->tile_w;
    p->tiles_y = c->h;
    // Screen(w, h), matrix.go:119-128
    const double w2 = (double)c->w / 2, h2 = (double)c->h / 2;
    const double scr[16] = {w2, 0, 0, w2, 0, -h2, 0, h2, 0, 0, 0.5, 0.5, 0, 0, 0, 1};
    memcpy(p->screen, scr, sizeof scr);
    if (lines) {
        p->mesh.pos = mesh->lpos; p->mesh.nrm = mesh->lnrm; p->mesh.tex = mesh->ltex; p->mesh.col = mesh->lcol;
        p->mesh.n = (uint32_t)mesh->nl; p->mesh.nverts = 2;
    } else {
        p->mesh.pos = mesh->tpos; p->mesh.nrm = mesh->tnrm; p->mesh.tex = mesh->ttex; p->mesh.col = mesh->tcol;
        p->mesh.n = (uint32_t)mesh->nt; p->mesh.nverts = 3;
    }
    p->first = (uint32_t)first; p->count = (uint32_t)count;
    p->is_lines = lines ? 1 : 0;
    // Deferred shading is legal when the fragment colour can neither be Discard (context.go:241)
    // nor take the blend path (context.go:256): the alpha is a per-draw constant then.
    {
        bool known = false;
        double alpha = 0;
        if (sh->kind == FGL_SHADER_SOLID) { known = true; alpha = sh->color[3]; }
        else if (sh->kind == FGL_SHADER_PHONG && !uses_tex && !p->object_is_discard) { known = true; alpha = sh->object[3]; }
        const bool blends = state->alpha_blend && alpha < 1;
        p->deferred = (known && alpha != 0 && !blends) ? 1 : 0;
    }
    return FGL_OK;
}

void prof_drain(fgl_ctx *c) {
    if (!c->prof_used) return;
    cudaStreamSynchronize(c->stream);
    for (int i = 0; i < c->prof_used; i++) {
        float g = 0, s = 0, b = 0, r = 0;
        cudaEventElapsedTime(&g, c->prof[i].e[0], c->prof[i].e[1]);
        cudaEventElapsedTime(&s, c->prof[i].e[1], c->prof[i].e[2]);
        cudaEventElapsedTime(&b, c->prof[i].e[2], c->prof[i].e[3]);
        cudaEventElapsedTime(&r, c->prof[i].e[3], c->prof[i].e[4]);
        c->prof_acc.geometry_ms += g; c->prof_acc.spans_ms += s; c->prof_acc.sort_ms += b; c->prof_acc.raster_ms += r;
    }
    c->prof_acc.draws += (uint32_t)c->prof_used;
    c->prof_used = 0;
}

// Front end of a draw.  Large draws take the fused kernel (k_front): every block of 128 primitives walks its
// own scanlines, which balances well when there are thousands of blocks.  Small draws keep the split
// stages, whose span kernels spread the scanlines of a few large triangles over the whole GPU.
// FGL_FRONT=fused|split (read when the context is created) forces one of them, for tests and tuning.
constexpr uint64_t FUSED_MIN_PRIMS = 16384;
bool use_fused_front(const fgl_ctx *c, const DrawParams &p) {
    if (c->front_mode == 1) return true;
    if (c->front_mode == 2) return false;
    return p.count >= FUSED_MIN_PRIMS;
}

int enqueue_draw(fgl_ctx *c, const DrawParams &p, bool async) {
    int launches = 0;
    ProfSlot *ps = nullptr;
    if (c->profiling) {
        if (c->prof_used == PROF_RING) prof_drain(c);
        ps = &c->prof[c->prof_used++];
        cudaEventRecord(ps->e[0], c->stream);
    }
    if (p.prim_info) cudaMemsetAsync(p.prim_info, 0, sizeof(unsigned long long) * 2 * p.count, c->stream);
    int sorted = 0;
    const bool fused = use_fused_front(c, p);
    if (fused) {
        // large draws: geometry and spans in one kernel (the stage timers then read: geometry = k_front alone,
        // spans = k_seg_index -- nothing when k_order does the listing, whose time is then all under `sort`)
        launches += launch_front(p, c->wb, c->counters_clean, c->stream);
        if (ps) cudaEventRecord(ps->e[1], c->stream);
        if (!c->order_coop) launches += launch_seg_index(p, c->wb, c->stream);
        if (ps) cudaEventRecord(ps->e[2], c->stream);
    } else {
        launches += launch_geometry(p, c->wb, c->counters_clean, c->stream);
        if (ps) cudaEventRecord(ps->e[1], c->stream);
        launches += launch_spans(p, c->wb, &sorted, c->stream);
        if (ps) cudaEventRecord(ps->e[2], c->stream);
    }
    if (c->order_coop) launches += launch_order(p, c->wb, fused, &sorted, c->stream);
    else launches += launch_bin(p, c->wb, &sorted, c->stream);
    if (ps) cudaEventRecord(ps->e[3], c->stream);
    fb_join(c);  // the front end and the binning ran beside a pending clear; the strips need the framebuffer
    bool accumulated = false;
    launches += launch_raster(p, c->wb, sorted, c->color, c->depth, async ? c->acc_dev : nullptr, &accumulated, c->stream);
    c->counters_clean = async && accumulated;  // k_shade's last CTA zeroes them after accumulating
    if (async && !accumulated) {  // (the deferred-shading kernel does it on the way, other paths need the extra launch)
        launch_pdl(k_accumulate, 1, 1, 0, c->stream, (const DrawCounters *)c->wb.counters, c->acc_dev);
        launches++;
    }
    if (ps) cudaEventRecord(ps->e[4], c->stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(c, FGL_E_CUDA, "kernel launch: %s", cudaGetErrorString(e));
    c->stats.kernel_launches = (uint32_t)launches;
    return FGL_OK;
}

int initial_capacity(fgl_ctx *c, const DrawParams &p) {
    const uint64_t n = p.count;
    // a primitive normally yields <= 1 record (2 for a line, 6 in wireframe); clipping may add a few.
    // Scanlines and segments depend on the projected size: start from a guess, the draw reports what
    // it needed and the buffers are regrown on overflow.
    const uint64_t per = p.is_lines ? 2 : (p.state.wireframe ? 6 : 1);
    const uint64_t rec = n * per + n / 8 + 1024;
    Caps want{n, std::max<uint64_t>(rec, c->wb.cap_records), std::max<uint64_t>(rec * 4, c->wb.cap_rows),
              std::max<uint64_t>(rec * 4, c->wb.cap_segs), std::max<uint64_t>(std::max<uint64_t>(n / 16, 4096), c->wb.cap_clip)};
    want.rows = std::max<uint64_t>(want.rows, 1u << 16);
    want.segs = std::max<uint64_t>(want.segs, 1u << 16);
    return ensure_work(c, want);
}

int draw_common(fgl_ctx *c, const fgl_state *state, const fgl_shader *sh, const fgl_mesh *mesh, uint64_t first,
                uint64_t count, bool lines, bool async, fgl_raster_info *info, fgl_raster_info *each = nullptr,
                bool want_each = false) {
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    if (info) { info->total_pixels = 0; info->updated_pixels = 0; }
    DrawParams p;
    rc = build_params(c, state, sh, mesh, first, count, lines, &p);
    if (rc) return rc;
    c->stats.prims_in = count;
    c->stats.retries = 0;
    if (want_each && !each && count) return fail(c, FGL_E_INVALID, "null per-primitive info array");
    if (count == 0) return FGL_OK;
    if (want_each) {
        static_assert(sizeof(fgl_raster_info) == 2 * sizeof(unsigned long long), "fgl_raster_info layout");
        if (count > c->prim_info_cap) {
            dev_free(c->prim_info);
            c->prim_info_cap = 0;
            CK(c, dev_alloc(&c->prim_info, (size_t)count * 2));
            c->prim_info_cap = count;
        }
        p.prim_info = c->prim_info;
    }
    if (c->capturing) {
        // a recorded frame replays fixed launches over fixed buffers: nothing may block, allocate or wait for an upload
        if (!async) return fail(c, FGL_E_INVALID, "only the *_async draws can be recorded into a graph (a synchronous draw waits for its RasterizeInfo)");
        if (mesh->upload_pending) return fail(c, FGL_E_INVALID, "fgl_mesh_upload_wait the mesh before recording a graph that draws it");
        if (c->profiling) return fail(c, FGL_E_INVALID, "switch the stage timers off while recording a graph");
        const WorkBuffers before = c->wb;
        rc = initial_capacity(c, p);
        if (rc) return rc;
        if ((p.deferred && !c->wb.vis_seg) || before.segv != c->wb.segv || before.recs != c->wb.recs ||
            before.row_first != c->wb.row_first || before.clip_pool != c->wb.clip_pool || before.blk_agg != c->wb.blk_agg)
            return fail(c, FGL_E_INVALID, "the work buffers had to grow while recording: issue one synchronous draw of this mesh "
                                          "first (it sizes them), then record again");
    }
    if (p.deferred && !c->wb.vis_seg)  // winners of the deferred-shading path, strip-major
        CK(c, dev_alloc(&c->wb.vis_seg, (size_t)c->wb.ntiles * c->tile_w));
    rc = initial_capacity(c, p);
    if (rc) return rc;
    mesh_acquire(c, mesh);
    if (async) {
        rc = enqueue_draw(c, p, true);
        if (rc) return rc;
        mesh_release(c, mesh);
        c->async_pending = true;
        return FGL_OK;
    }
    for (int attempt = 0; attempt < 6; attempt++) {
        rc = enqueue_draw(c, p, false);
        if (rc) return rc;
        CK(c, cudaMemcpyAsync(c->host_counters, c->wb.counters, sizeof(DrawCounters), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaStreamSynchronize(c->stream));
        const DrawCounters &hc = *c->host_counters;
        if (!hc.overflow) {
            mesh_release(c, mesh);
            if (p.prim_info) {
                CK(c, cudaMemcpyAsync(each, p.prim_info, sizeof(fgl_raster_info) * count, cudaMemcpyDeviceToHost, c->stream));
                CK(c, cudaStreamSynchronize(c->stream));
            }
            if (info) { info->total_pixels = hc.total_pixels; info->updated_pixels = hc.updated_pixels; }
            c->stats.records = hc.n_records; c->stats.pairs = hc.n_segs; c->stats.clip_triangles = hc.n_clip;
            return FGL_OK;
        }
        // the raster kernel did not run (it checks the flag): grow and re-issue
        c->stats.retries++;
        rc = ensure_work(c, grown(c, hc));
        if (rc) return rc;
    }
    return fail(c, FGL_E_OVERFLOW, "work buffers still too small after regrowing");
}

}  // namespace

namespace fgl {
int api_fail(fgl_ctx *ctx, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    return fail(ctx, code, "%s", buf);
}
int api_check_ctx(fgl_ctx *ctx) { return check_ctx(ctx); }
void api_fb_join(fgl_ctx *ctx) { fb_join(ctx); }
}  // namespace fgl

extern "C" {

// ---- recorded frames (CUDA graphs) -----------------------------------------------------------------------------

int fgl_graph_begin(fgl_ctx *c) {
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->capturing) return fail(c, FGL_E_INVALID, "already recording a graph on this context");
    fb_join(c);  // (a clear issued before the recording belongs to the work before it)
    // relaxed: the recorded calls may use cudaMalloc-free helper calls of other threads' contexts meanwhile
    CK(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
    c->capturing = true;
    c->counters_clean = false;  // the recorded frame zeroes the draw counters itself, whatever ran before a replay
    return FGL_OK;
}

int fgl_graph_end(fgl_ctx *c, fgl_graph **out) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!out) return fail(c, FGL_E_INVALID, "null out pointer");
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    if (!c->capturing) return fail(c, FGL_E_INVALID, "fgl_graph_end without fgl_graph_begin");
    fb_join(c);  // the clear stream joins the recording again
    c->capturing = false;
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(c->stream, &g);
    if (e != cudaSuccess || !g) {
        cudaGetLastError();
        c->counters_clean = false;
        return fail(c, FGL_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
    }
    cudaGraphExec_t ex = nullptr;
    e = cudaGraphInstantiate(&ex, g, 0);
    if (e != cudaSuccess) {
        cudaGetLastError();
        cudaGraphDestroy(g);
        c->counters_clean = false;
        return fail(c, FGL_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
    }
    fgl_graph *gr = new (std::nothrow) fgl_graph();
    if (!gr) { cudaGraphExecDestroy(ex); cudaGraphDestroy(g); return fail(c, FGL_E_OOM, "host allocation failed"); }
    gr->device = c->device; gr->ctx = c; gr->graph = g; gr->exec = ex; gr->buffer_epoch = c->buffer_epoch;
    gr->has_draws = c->async_pending;   // (set by the recorded draws; nothing ran yet)
    gr->counters_clean = c->counters_clean;
    c->async_pending = false;           // the recording itself drew nothing
    c->counters_clean = false;
    *out = gr;
    return FGL_OK;
}

int fgl_graph_launch(fgl_ctx *c, fgl_graph *g) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!g || g->ctx != c) return fail(c, FGL_E_INVALID, "graph belongs to another context");
    std::lock_guard<std::mutex> lock(c->mu);
    if (c->capturing) return fail(c, FGL_E_INVALID, "cannot launch a graph while recording one");
    if (g->buffer_epoch != c->buffer_epoch)
        return fail(c, FGL_E_INVALID, "the context's work or resolve buffers were reallocated after this graph was recorded "
                                      "(a larger draw, another resolve size): record it again");
    fb_join(c);
    CK(c, cudaGraphLaunch(g->exec, c->stream));
    if (g->has_draws) c->async_pending = true;
    c->counters_clean = g->counters_clean;
    return FGL_OK;
}

int fgl_graph_destroy(fgl_graph *g) {
    if (!g) return FGL_OK;
    cudaSetDevice(g->device);
    if (g->ctx && g->ctx->stream) cudaStreamSynchronize(g->ctx->stream);
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return FGL_OK;
}

int fgl_abi_version(void) { return FGL_ABI_VERSION; }

const char *fgl_last_error(const fgl_ctx *ctx) {
    t_error_out = t_last_error;
    if (ctx) {
        std::lock_guard<std::mutex> g(ctx->err_mu);
        if (ctx->err_seq > t_last_seq) t_error_out = ctx->err;
    }
    return t_error_out.c_str();
}

int fgl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int fgl_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(nullptr, FGL_E_INVALID, "null out pointer");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, e == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "cudaMallocHost(%zu): %s", bytes,
                    cudaGetErrorString(e));
    }
    return FGL_OK;
}
int fgl_host_free(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, FGL_E_CUDA, "cudaFreeHost failed");
    }
    return FGL_OK;
}

int fgl_context_create(int width, int height, int device, fgl_ctx **out) {
    if (!out) return fail(nullptr, FGL_E_INVALID, "null out pointer");
    *out = nullptr;
    if (width <= 0 || height <= 0 || width > 65535 || height > 65535 ||
        (uint64_t)width * (uint64_t)height > (1ull << 31))
        return fail(nullptr, FGL_E_INVALID, "bad framebuffer size %dx%d", width, height);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(nullptr, FGL_E_NO_DEVICE, "no CUDA device available (%s); fauxgl_b200 has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (device < 0 || device >= ndev) return fail(nullptr, FGL_E_INVALID, "device %d out of range [0,%d)", device, ndev);
    CK(nullptr, cudaSetDevice(device));
    fgl_ctx *c = new (std::nothrow) fgl_ctx();
    if (!c) return fail(nullptr, FGL_E_OOM, "host allocation failed");
    c->device = device; c->w = width; c->h = height; c->err_seq = 0;
    {
        const char *fm = getenv("FGL_FRONT");
        c->front_mode = !fm ? 0 : (strcmp(fm, "fused") == 0 ? 1 : (strcmp(fm, "split") == 0 ? 2 : 0));
    }
    c->color = nullptr; c->depth = nullptr; c->resolved = nullptr; c->rw = c->rh = 0;
    memset(&c->wb, 0, sizeof c->wb);
    memset(&c->stats, 0, sizeof c->stats);
    c->host_counters = nullptr; c->acc_dev = nullptr; c->async_pending = false; c->counters_clean = false;
    c->capturing = false; c->buffer_epoch = 0;
    c->prim_info = nullptr; c->prim_info_cap = 0; c->scratch = nullptr; c->gray16 = nullptr;
    c->peer_flags = nullptr; c->peer_epoch = 0; c->clear_depth_value = 1.7976931348623157e308;
    c->clear_color_value = 0u; c->clear_color_known = true;  // NewContext: transparent black
    c->profiling = false; c->prof_used = 0; c->prof_created = false;
    memset(&c->prof_acc, 0, sizeof c->prof_acc);
    const size_t npix = (size_t)width * height;
    c->stream = nullptr; c->copy_stream = nullptr; c->fb_stream = nullptr;
    c->ev_fb_free = nullptr; c->ev_cleared = nullptr; c->clear_pending = false;
    c->read_stream = nullptr; c->rb_next = 0;
    for (int k = 0; k < 2; k++) {
        c->rb_color[k] = nullptr; c->rb_counters[k] = nullptr; c->ev_rb_staged[k] = nullptr; c->ev_rb_done[k] = nullptr;
        c->rb_done_recorded[k] = false;
    }
    {
        const char *ro = getenv("FGL_READBACK_OVERLAP");  // tuning aid: 0 reads back on the draw stream, straight from the framebuffer
        c->rb_overlap = !(ro && atoi(ro) == 0);
    }
    {
        const char *pd = getenv("FGL_PDL");  // tuning aid: 0 launches every kernel fully serialised
        fgl::g_pdl = !(pd && atoi(pd) == 0);
        const char *bn = getenv("FGL_BIN");  // tuning aid: lsd = every radix pass global + k_tile_ranges (fgl_scan_sort.cu)
        fgl::g_bin_buckets = !(bn && strcmp(bn, "lsd") == 0);
    }
    {
        const char *co = getenv("FGL_CLEAR_OVERLAP");  // tuning aid: 0 keeps the clears on the draw stream
        c->clear_overlap = !(co && atoi(co) == 0);
    }
    cudaError_t err = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&c->fb_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_fb_free, cudaEventDisableTiming);
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&c->ev_cleared, cudaEventDisableTiming);
    if (err == cudaSuccess) err = dev_alloc(&c->color, npix);
    if (err == cudaSuccess) err = dev_alloc(&c->depth, npix);
    if (err == cudaSuccess) err = dev_alloc(&c->wb.counters, 1);
    if (err == cudaSuccess) err = dev_alloc(&c->acc_dev, 1);
    if (err == cudaSuccess) err = dev_alloc(&c->scratch, 8);
    if (err == cudaSuccess) err = cudaMallocHost(reinterpret_cast<void **>(&c->host_counters), sizeof(DrawCounters));
    {   // strips of 64 x 1 pixels
        const char *tw = getenv("FGL_STRIP_W");  // tuning aid: force 32 or 64
        c->tile_w = tw && atoi(tw) == 32 ? 32 : (tw && atoi(tw) == 64 ? 64 : ((uint64_t)width * height <= (4ull << 20) ? 32 : 64));
        const int tx = (width + c->tile_w - 1) / c->tile_w;
        c->wb.ntiles = (uint32_t)tx * (uint32_t)height;
        cudaDeviceProp prop;
        c->wb.nsm = cudaGetDeviceProperties(&prop, device) == cudaSuccess ? (uint32_t)prop.multiProcessorCount : 148u;
    }
    if (err == cudaSuccess) err = dev_alloc(&c->wb.busy_list, c->wb.ntiles);
    if (err == cudaSuccess) err = dev_alloc(&c->wb.dirty, (size_t)c->wb.ntiles + 16);
    if (err == cudaSuccess) err = cudaMemsetAsync(c->wb.dirty, 0, (size_t)c->wb.ntiles + 16, c->stream);
    if (err == cudaSuccess) err = dev_alloc(&c->wb.tile_ctl, 1);
    if (err == cudaSuccess) err = cudaMemsetAsync(c->wb.tile_ctl, 0, sizeof(fgl::TileCtl), c->stream);  // (the barrier words of k_order)
    {
        // FGL_ORDER=coop: the order plumbing (k_seg_index + radix passes + k_tile_ranges) as ONE cooperative kernel with
        // grid-wide barriers (fgl_order.cu).  Bit-identical, but measured SLOWER than the six separate launches (1080p:
        // 59 vs 53 us, 8K: 203 vs 197 us; four contexts in flight: 0.217 vs 0.175 ms per frame, a gang-scheduled grid
        // with one 1024-thread CTA per SM cannot overlap other streams): off by default, kept as a tuning variant.
        const char *om = getenv("FGL_ORDER");
        c->order_coop = om && strcmp(om, "coop") == 0 && fgl::order_supported(device, c->wb.nsm);
    }
    if (err == cudaSuccess && getenv("FGL_TILE_CLOCK")) {  // tuning aid: per-tile cycle counts of k_tile
        err = dev_alloc(&c->wb.tile_clock, (size_t)c->wb.ntiles * 2 + 16);  // + 16 path counters of k_strip
        if (err == cudaSuccess) err = cudaMemset(c->wb.tile_clock, 0, sizeof(unsigned long long) * ((size_t)c->wb.ntiles * 2 + 16));
    }
    if (err == cudaSuccess) err = cudaMemsetAsync(c->acc_dev, 0, sizeof(DrawCounters), c->stream);
    if (err != cudaSuccess) {
        cudaGetLastError();
        int rc = fail(nullptr, err == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "context allocation: %s",
                      cudaGetErrorString(err));
        fgl_context_destroy(c);
        return rc;
    }
    c->stats.tiles_x = (uint32_t)((width + c->tile_w - 1) / c->tile_w);
    c->stats.tiles_y = (uint32_t)height;
    c->stats.tile_w = (uint32_t)c->tile_w; c->stats.tile_h = 1;
    // NewContext: image.NewNRGBA is zeroed; ClearDepthBuffer() -> math.MaxFloat64 (context.go:64,79)
    launch_clear_color(c->color, npix, 0u, c->stream);
    launch_clear_depth(c->depth, npix, 1.7976931348623157e308, c->stream);
    err = cudaStreamSynchronize(c->stream);
    if (err != cudaSuccess) {
        int rc = fail(nullptr, FGL_E_CUDA, "context init: %s", cudaGetErrorString(err));
        fgl_context_destroy(c);
        return rc;
    }
    *out = c;
    return FGL_OK;
}

int fgl_context_destroy(fgl_ctx *c) {
    if (!c) return FGL_OK;
    cudaSetDevice(c->device);
    if (c->copy_stream) { cudaStreamSynchronize(c->copy_stream); cudaStreamDestroy(c->copy_stream); }
    if (c->fb_stream) { cudaStreamSynchronize(c->fb_stream); cudaStreamDestroy(c->fb_stream); }
    if (c->read_stream) { cudaStreamSynchronize(c->read_stream); cudaStreamDestroy(c->read_stream); }
    for (int k = 0; k < 2; k++) {
        if (c->ev_rb_staged[k]) cudaEventDestroy(c->ev_rb_staged[k]);
        if (c->ev_rb_done[k]) cudaEventDestroy(c->ev_rb_done[k]);
        dev_free(c->rb_color[k]); dev_free(c->rb_counters[k]);
    }
    if (c->ev_fb_free) cudaEventDestroy(c->ev_fb_free);
    if (c->ev_cleared) cudaEventDestroy(c->ev_cleared);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_work(c->wb);
    dev_free(c->wb.counters); dev_free(c->wb.tile_clock);
    dev_free(c->wb.tile_ctl); dev_free(c->wb.busy_list); dev_free(c->wb.vis_seg); dev_free(c->wb.dirty);
    dev_free(c->prim_info); dev_free(c->scratch); dev_free(c->gray16); dev_free(c->peer_flags);
    dev_free(c->acc_dev); dev_free(c->color); dev_free(c->depth); dev_free(c->resolved);
    if (c->host_counters) cudaFreeHost(c->host_counters);
    if (c->prof_created)
        for (int i = 0; i < PROF_RING; i++)
            for (int k = 0; k < PROF_EVENTS; k++) cudaEventDestroy(c->prof[i].e[k]);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return FGL_OK;
}

int fgl_context_size(const fgl_ctx *c, int *width, int *height) {
    if (!c) return fail(nullptr, FGL_E_INVALID, "null context");
    if (width) *width = c->w;
    if (height) *height = c->h;
    return FGL_OK;
}

int fgl_clear_color(fgl_ctx *c, const uint8_t rgba[4]) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!rgba) return fail(c, FGL_E_INVALID, "null colour");
    std::lock_guard<std::mutex> lock(c->mu);
    const uint32_t v = (uint32_t)rgba[0] | ((uint32_t)rgba[1] << 8) | ((uint32_t)rgba[2] << 16) | ((uint32_t)rgba[3] << 24);
    launch_clear_color(c->color, (size_t)c->w * c->h, v, fb_clear_begin(c));
    fb_clear_end(c);
    c->clear_color_value = v; c->clear_color_known = true;
    CK(c, cudaGetLastError());
    return FGL_OK;
}

int fgl_clear_depth(fgl_ctx *c, double value) {
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    cudaStream_t cs = fb_clear_begin(c);
    cudaMemsetAsync(c->wb.dirty, 0, c->wb.ntiles, cs);  // no strip has drawn depth any more
    c->clear_depth_value = value;
    launch_clear_depth(c->depth, (size_t)c->w * c->h, value, cs);
    fb_clear_end(c);
    CK(c, cudaGetLastError());
    return FGL_OK;
}

// ---- meshes ---------------------------------------------------------------------------------

// Each attribute lands in its own region of the staging buffer, so all H2D copies of
// one upload are queued back to back and the transposing ingest kernels trail them.
static int upload_attr(fgl_ctx *c, const double *host, double **planes, uint64_t n, int nverts, int ncomp_in,
                       int ncomp_out, double *staging, bool create, cudaStream_t st) {
    if (create) CK(c, dev_alloc(planes, (size_t)n * nverts * ncomp_out));
    if (n == 0) return FGL_OK;
    if (!host) {
        if (create) CK(c, cudaMemsetAsync(*planes, 0, sizeof(double) * n * nverts * ncomp_out, st));
        return FGL_OK;
    }
    CK(c, cudaMemcpyAsync(staging, host, sizeof(double) * n * nverts * ncomp_in, cudaMemcpyHostToDevice, st));
    launch_mesh_ingest(staging, *planes, (uint32_t)n, nverts, ncomp_in, ncomp_out, st);
    CK(c, cudaGetLastError());
    return FGL_OK;
}

static int upload_all(fgl_ctx *c, fgl_mesh *m, const fgl_mesh_desc *d, bool create, cudaStream_t st = nullptr,
                      bool wait = true) {
    if (!st) st = c->stream;
    // staging regions: [pos | nrm | tex | col] for triangles, then the same for lines
    const size_t t3 = (size_t)m->nt * 9, t4 = (size_t)m->nt * 12, l3 = (size_t)m->nl * 6, l4 = (size_t)m->nl * 8;
    double *s = m->staging;
    int rc = upload_attr(c, d->position, &m->tpos, m->nt, 3, 3, 3, s, create, st);
    if (!rc) rc = upload_attr(c, d->normal, &m->tnrm, m->nt, 3, 3, 3, s + t3, create, st);
    if (!rc) rc = upload_attr(c, d->texture, &m->ttex, m->nt, 3, 3, 2, s + 2 * t3, create, st);
    if (!rc) rc = upload_attr(c, d->color, &m->tcol, m->nt, 3, 4, 4, s + 3 * t3, create, st);
    s += 3 * t3 + t4;
    if (!rc) rc = upload_attr(c, d->lposition, &m->lpos, m->nl, 2, 3, 3, s, create, st);
    if (!rc) rc = upload_attr(c, d->lnormal, &m->lnrm, m->nl, 2, 3, 3, s + l3, create, st);
    if (!rc) rc = upload_attr(c, d->ltexture, &m->ltex, m->nl, 2, 3, 2, s + 2 * l3, create, st);
    if (!rc) rc = upload_attr(c, d->lcolor, &m->lcol, m->nl, 2, 4, 4, s + 3 * l3, create, st);
    (void)l4;
    if (!wait) return rc;
    cudaError_t se = cudaStreamSynchronize(st);  // the caller may reuse its host arrays
    if (!rc && se != cudaSuccess) rc = fail(c, FGL_E_CUDA, "mesh upload: %s", cudaGetErrorString(se));
    return rc;
}

int fgl_mesh_create(fgl_ctx *c, const fgl_mesh_desc *d, fgl_mesh **out) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!d || !out) return fail(c, FGL_E_INVALID, "null mesh description/out pointer");
    *out = nullptr;
    if (d->ntriangles > 0xfffffff0ull || d->nlines > 0xfffffff0ull) return fail(c, FGL_E_INVALID, "mesh too large");
    if ((d->ntriangles && !d->position) || (d->nlines && !d->lposition))
        return fail(c, FGL_E_INVALID, "position array missing");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_create");
    fgl_mesh *m = new (std::nothrow) fgl_mesh();
    if (!m) return fail(c, FGL_E_OOM, "host allocation failed");
    memset(m, 0, sizeof *m);
    m->device = c->device; m->owner = c; m->nt = d->ntriangles; m->nl = d->nlines;
    m->staging_elems = (size_t)m->nt * (9 * 3 + 12) + (size_t)m->nl * (6 * 3 + 8);
    cudaError_t e = dev_alloc(&m->staging, m->staging_elems);
    if (e != cudaSuccess) { delete m; cudaGetLastError(); return fail(c, FGL_E_OOM, "staging: %s", cudaGetErrorString(e)); }
    rc = upload_all(c, m, d, true);
    if (rc) { fgl_mesh_destroy(m); return rc; }
    *out = m;
    return FGL_OK;
}

int fgl_mesh_update(fgl_ctx *c, fgl_mesh *m, const fgl_mesh_desc *d) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m || !d) return fail(c, FGL_E_INVALID, "null mesh/description");
    if (m->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on another device");
    if (m->owner != c) return fail(c, FGL_E_INVALID, "a mesh is modified through the context it was created with (any context of the device may draw it)");
    if (d->ntriangles != m->nt || d->nlines != m->nl)
        return fail(c, FGL_E_INVALID, "fgl_mesh_update needs the same primitive counts (%llu/%llu vs %llu/%llu)",
                    (unsigned long long)d->ntriangles, (unsigned long long)d->nlines, (unsigned long long)m->nt,
                    (unsigned long long)m->nl);
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_update");
    mesh_acquire(c, m);
    return upload_all(c, m, d, false);
}

int fgl_mesh_update_async(fgl_ctx *c, fgl_mesh *m, const fgl_mesh_desc *d) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m || !d) return fail(c, FGL_E_INVALID, "null mesh/description");
    if (m->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on another device");
    if (m->owner != c) return fail(c, FGL_E_INVALID, "a mesh is modified through the context it was created with (any context of the device may draw it)");
    if (d->ntriangles != m->nt || d->nlines != m->nl)
        return fail(c, FGL_E_INVALID, "fgl_mesh_update_async needs the same primitive counts");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_update_async");
    if (!m->has_events) {
        CK(c, cudaEventCreateWithFlags(&m->ev_uploaded, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&m->ev_drawn, cudaEventDisableTiming));
        m->has_events = true;
        // everything enqueued on the draw stream so far may still read the buffers
        cudaEventRecord(m->ev_drawn, c->stream);
        m->drawn_recorded = true;
    }
    if (m->drawn_recorded) CK(c, cudaStreamWaitEvent(c->copy_stream, m->ev_drawn, 0));
    rc = upload_all(c, m, d, false, c->copy_stream, false);
    CK(c, cudaEventRecord(m->ev_uploaded, c->copy_stream));
    m->upload_pending = true;
    return rc;
}

int fgl_mesh_update_indexed_async(fgl_ctx *c, fgl_mesh *m, const fgl_indexed_desc *d) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m || !d) return fail(c, FGL_E_INVALID, "null mesh/description");
    if (m->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on another device");
    if (m->owner != c) return fail(c, FGL_E_INVALID, "a mesh is modified through the context it was created with (any context of the device may draw it)");
    if (!m->corners) return fail(c, FGL_E_INVALID, "not an indexed mesh (fgl_mesh_create_indexed)");
    if ((d->v && d->nv != m->nv) || (d->vt && d->nvt != m->nvt) || (d->vn && d->nvn != m->nvn))
        return fail(c, FGL_E_INVALID, "fgl_mesh_update_indexed_async needs tables of the sizes the mesh was created with");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_update_indexed_async");
    if (!m->has_events) {
        CK(c, cudaEventCreateWithFlags(&m->ev_uploaded, cudaEventDisableTiming));
        CK(c, cudaEventCreateWithFlags(&m->ev_drawn, cudaEventDisableTiming));
        m->has_events = true;
        cudaEventRecord(m->ev_drawn, c->stream);  // everything enqueued on the draw stream so far may still read the planes
        m->drawn_recorded = true;
    }
    cudaStream_t st = c->copy_stream;
    if (m->drawn_recorded) CK(c, cudaStreamWaitEvent(st, m->ev_drawn, 0));
    if (d->v) CK(c, cudaMemcpyAsync(m->tab_v, d->v, sizeof(double) * 3 * m->nv, cudaMemcpyHostToDevice, st));
    if (d->vt) CK(c, cudaMemcpyAsync(m->tab_vt, d->vt, sizeof(double) * 3 * m->nvt, cudaMemcpyHostToDevice, st));
    if (d->vn) CK(c, cudaMemcpyAsync(m->tab_vn, d->vn, sizeof(double) * 3 * m->nvn, cudaMemcpyHostToDevice, st));
    CK(c, cudaEventRecord(m->ev_uploaded, st));  // (fgl_mesh_upload_wait: the host tables may be reused)
    // The expansion runs on the DRAW stream, behind the upload and, the stream being in order, behind every draw that
    // still reads the planes: the copy stream carries nothing but PCIe transfers, so the tables of the next frame (another
    // mesh) follow this frame's without a gap.  The texture planes are rewritten only when vt was sent.
    CK(c, cudaStreamWaitEvent(c->stream, m->ev_uploaded, 0));
    launch_indexed_ingest(m->tab_v, m->tab_vt, m->tab_vn, m->corners, m->tpos, m->tnrm, d->vt ? m->ttex : nullptr, (uint32_t)m->nt,
                          c->stream);
    CK(c, cudaGetLastError());
    CK(c, cudaEventRecord(m->ev_drawn, c->stream));  // the tables and the planes are in use up to here
    m->drawn_recorded = true;
    m->upload_pending = false;
    return FGL_OK;
}

int fgl_mesh_upload_wait(fgl_ctx *c, fgl_mesh *m) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m) return fail(c, FGL_E_INVALID, "null mesh");
    if (m->has_events) CK(c, cudaEventSynchronize(m->ev_uploaded));
    return FGL_OK;
}

int fgl_mesh_create_stl(fgl_ctx *c, const uint8_t *records, uint64_t count, fgl_mesh **out) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!out || (count && !records)) return fail(c, FGL_E_INVALID, "null STL records/out pointer");
    *out = nullptr;
    if (count > 0xfffffff0ull) return fail(c, FGL_E_INVALID, "mesh too large");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_create_stl");
    fgl_mesh *m = new (std::nothrow) fgl_mesh();
    if (!m) return fail(c, FGL_E_OOM, "host allocation failed");
    memset(m, 0, sizeof *m);
    m->device = c->device; m->owner = c; m->nt = count; m->nl = 0;
    m->staging_elems = (size_t)m->nt * (9 * 3 + 12);  // 312 B per triangle: also holds the 50-byte records (+ padding)
    cudaError_t e = dev_alloc(&m->staging, m->staging_elems);
    if (e == cudaSuccess) e = dev_alloc(&m->tpos, (size_t)count * 9);
    if (e == cudaSuccess) e = dev_alloc(&m->tnrm, (size_t)count * 9);
    if (e == cudaSuccess) e = dev_alloc(&m->ttex, (size_t)count * 6);
    if (e == cudaSuccess) e = dev_alloc(&m->tcol, (size_t)count * 12);
    if (e == cudaSuccess) e = dev_alloc(&m->lpos, 1);
    if (e == cudaSuccess) e = dev_alloc(&m->lnrm, 1);
    if (e == cudaSuccess) e = dev_alloc(&m->ltex, 1);
    if (e == cudaSuccess) e = dev_alloc(&m->lcol, 1);
    if (e == cudaSuccess && count) {
        e = cudaMemcpyAsync(m->staging, records, (size_t)count * 50, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(m->ttex, 0, sizeof(double) * count * 6, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(m->tcol, 0, sizeof(double) * count * 12, c->stream);
        if (e == cudaSuccess) {
            launch_stl_ingest(reinterpret_cast<const uint8_t *>(m->staging), m->tpos, m->tnrm, (uint32_t)count, c->stream);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // the caller may free `records`
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fgl_mesh_destroy(m);
        return fail(c, e == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "STL upload: %s", cudaGetErrorString(e));
    }
    *out = m;
    return FGL_OK;
}

int fgl_mesh_create_indexed(fgl_ctx *c, const fgl_indexed_desc *d, fgl_mesh **out) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!out || !d) return fail(c, FGL_E_INVALID, "null descriptor/out pointer");
    *out = nullptr;
    const uint64_t count = d->ntriangles;
    if (count > 0xfffffff0ull) return fail(c, FGL_E_INVALID, "mesh too large");
    if (count && (!d->corners || !d->v || !d->vt || !d->vn || !d->nv || !d->nvt || !d->nvn))
        return fail(c, FGL_E_INVALID, "null / empty table in an indexed mesh");
    for (uint64_t k = 0; k < count * 3; k++) {  // the reference panics on an index outside its table
        const int32_t *x = d->corners + k * 3;
        if (x[0] < 0 || (uint64_t)x[0] >= d->nv || x[1] < 0 || (uint64_t)x[1] >= d->nvt || x[2] < 0 || (uint64_t)x[2] >= d->nvn)
            return fail(c, FGL_E_INVALID, "corner %llu of the indexed mesh points outside its table", (unsigned long long)k);
    }
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_create_indexed");
    fgl_mesh *m = new (std::nothrow) fgl_mesh();
    if (!m) return fail(c, FGL_E_OOM, "host allocation failed");
    memset(m, 0, sizeof *m);
    m->device = c->device; m->owner = c; m->nt = count; m->nl = 0;
    m->staging_elems = (size_t)m->nt * (9 * 3 + 12);  // as fgl_mesh_create: later fgl_mesh_update calls land here
    double *&tv = m->tab_v, *&tvt = m->tab_vt, *&tvn = m->tab_vn;   // kept: fgl_mesh_update_indexed_async reuses them
    int32_t *&tc = m->corners;
    m->nv = d->nv; m->nvt = d->nvt; m->nvn = d->nvn;
    cudaError_t e = dev_alloc(&m->staging, m->staging_elems);
    if (e == cudaSuccess) e = dev_alloc(&m->tpos, (size_t)count * 9);
    if (e == cudaSuccess) e = dev_alloc(&m->tnrm, (size_t)count * 9);
    if (e == cudaSuccess) e = dev_alloc(&m->ttex, (size_t)count * 6);
    if (e == cudaSuccess) e = dev_alloc(&m->tcol, (size_t)count * 12);
    if (e == cudaSuccess) e = dev_alloc(&m->lpos, 1);
    if (e == cudaSuccess) e = dev_alloc(&m->lnrm, 1);
    if (e == cudaSuccess) e = dev_alloc(&m->ltex, 1);
    if (e == cudaSuccess) e = dev_alloc(&m->lcol, 1);
    if (e == cudaSuccess && count) {
        e = dev_alloc(&tv, (size_t)d->nv * 3);
        if (e == cudaSuccess) e = dev_alloc(&tvt, (size_t)d->nvt * 3);
        if (e == cudaSuccess) e = dev_alloc(&tvn, (size_t)d->nvn * 3);
        if (e == cudaSuccess) e = dev_alloc(&tc, (size_t)count * 9);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tv, d->v, sizeof(double) * 3 * d->nv, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tvt, d->vt, sizeof(double) * 3 * d->nvt, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tvn, d->vn, sizeof(double) * 3 * d->nvn, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(tc, d->corners, sizeof(int32_t) * 9 * count, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) e = cudaMemsetAsync(m->tcol, 0, sizeof(double) * count * 12, c->stream);
        if (e == cudaSuccess) {
            launch_indexed_ingest(tv, tvt, tvn, tc, m->tpos, m->tnrm, m->ttex, (uint32_t)count, c->stream);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // the caller may free its arrays
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        fgl_mesh_destroy(m);
        return fail(c, e == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "indexed mesh upload: %s", cudaGetErrorString(e));
    }
    *out = m;
    return FGL_OK;
}

int fgl_mesh_bounds(fgl_ctx *c, const fgl_mesh *m, double mn[3], double mx[3]) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m || !mn || !mx) return fail(c, FGL_E_INVALID, "null mesh/out pointer");
    if (m->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on another device");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_bounds");
    unsigned long long h[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
    mesh_acquire(c, m);
    CK(c, cudaMemcpyAsync(c->scratch, h, sizeof h, cudaMemcpyHostToDevice, c->stream));
    launch_mesh_bounds(m->tpos, (uint32_t)m->nt, 3, c->scratch, c->stream);
    launch_mesh_bounds(m->lpos, (uint32_t)m->nl, 2, c->scratch, c->stream);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(h, c->scratch, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < 3; k++) {
        if (h[k] == ~0ull || h[3 + k] == 0ull) { mn[k] = 0; mx[k] = 0; continue; }  // EmptyBox
        for (int e = 0; e < 2; e++) {
            const unsigned long long key = h[3 * e + k];
            const unsigned long long bits = (key >> 63) ? (key & 0x7fffffffffffffffull) : ~key;
            memcpy(e ? &mx[k] : &mn[k], &bits, sizeof bits);
        }
    }
    return FGL_OK;
}

int fgl_mesh_destroy(fgl_mesh *m) {
    if (!m) return FGL_OK;
    cudaSetDevice(m->device);
    dev_free(m->tpos); dev_free(m->tnrm); dev_free(m->ttex); dev_free(m->tcol);
    dev_free(m->lpos); dev_free(m->lnrm); dev_free(m->ltex); dev_free(m->lcol);
    dev_free(m->staging);
    dev_free(m->corners); dev_free(m->tab_v); dev_free(m->tab_vt); dev_free(m->tab_vn);
    if (m->has_events) { cudaEventDestroy(m->ev_uploaded); cudaEventDestroy(m->ev_drawn); }
    delete m;
    return FGL_OK;
}

int fgl_mesh_counts(const fgl_mesh *m, uint64_t *nt, uint64_t *nl) {
    if (!m) return fail(nullptr, FGL_E_INVALID, "null mesh");
    if (nt) *nt = m->nt;
    if (nl) *nl = m->nl;
    return FGL_OK;
}

int fgl_mesh_transform(fgl_ctx *c, fgl_mesh *m, const double matrix[16]) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m || !matrix) return fail(c, FGL_E_INVALID, "null mesh/matrix");
    if (m->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on another device");
    if (m->owner != c) return fail(c, FGL_E_INVALID, "a mesh is modified through the context it was created with (any context of the device may draw it)");
    std::lock_guard<std::mutex> lock(c->mu);
    mesh_acquire(c, m);
    launch_mesh_transform(m->tpos, m->tnrm, (uint32_t)m->nt, 3, matrix, c->stream);
    launch_mesh_transform(m->lpos, m->lnrm, (uint32_t)m->nl, 2, matrix, c->stream);
    mesh_release(c, m);
    CK(c, cudaGetLastError());
    return FGL_OK;
}

static int smooth_normals_common(fgl_ctx *c, fgl_mesh *m, bool with_threshold, double threshold) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m) return fail(c, FGL_E_INVALID, "null mesh");
    if (m->device != c->device) return fail(c, FGL_E_INVALID, "mesh lives on another device");
    if (m->owner != c) return fail(c, FGL_E_INVALID, "a mesh is modified through the context it was created with (any context of the device may draw it)");
    if (m->nt == 0) return FGL_OK;
    if (m->nt > 0x55555555ull) return fail(c, FGL_E_INVALID, "mesh too large for 32-bit corner indices");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_smooth_normals");
    const uint32_t n = (uint32_t)m->nt, nc = 3u * n;
    uint32_t *key[2] = {nullptr, nullptr}, *val[2] = {nullptr, nullptr}, *tmp = nullptr;
    unsigned int *n_dev = nullptr;
    double *nout = nullptr;
    cudaError_t e = dev_alloc(&key[0], nc);
    if (e == cudaSuccess) e = dev_alloc(&key[1], nc);
    if (e == cudaSuccess) e = dev_alloc(&val[0], nc);
    if (e == cudaSuccess) e = dev_alloc(&val[1], nc);
    if (e == cudaSuccess) e = dev_alloc(&tmp, scan_tmp_words(nc));
    if (e == cudaSuccess) e = dev_alloc(&n_dev, 1);
    if (e == cudaSuccess && with_threshold) e = dev_alloc(&nout, (size_t)n * 9);
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_dev, &nc, sizeof nc, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) {
        mesh_acquire(c, m);
        launch_corner_hash(m->tpos, n, key[0], val[0], c->stream);
        int sorted = 0;
        launch_sort_pairs(key, val, n_dev, nc, 32, tmp, &sorted, c->stream);
        if (with_threshold) {
            launch_smooth_threshold(key[sorted], val[sorted], m->tpos, m->tnrm, nout, n, threshold, c->stream);
            cudaMemcpyAsync(m->tnrm, nout, sizeof(double) * 9 * n, cudaMemcpyDeviceToDevice, c->stream);
        } else {
            launch_smooth_groups(key[sorted], val[sorted], m->tpos, m->tnrm, n, c->stream);
        }
        mesh_release(c, m);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);  // the temporaries are freed below
    }
    dev_free(key[0]); dev_free(key[1]); dev_free(val[0]); dev_free(val[1]); dev_free(tmp); dev_free(n_dev); dev_free(nout);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(c, e == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "smooth normals: %s", cudaGetErrorString(e));
    }
    return FGL_OK;
}
int fgl_mesh_smooth_normals(fgl_ctx *c, fgl_mesh *m) { return smooth_normals_common(c, m, false, 0); }
int fgl_mesh_smooth_normals_threshold(fgl_ctx *c, fgl_mesh *m, double cos_threshold) {
    return smooth_normals_common(c, m, true, cos_threshold);
}

static int read_attr(fgl_ctx *c, const double *planes, double *host, uint64_t n, int nverts, int ncomp) {
    if (!host || n == 0) return FGL_OK;
    double *staging = nullptr;
    CK(c, dev_alloc(&staging, (size_t)n * nverts * ncomp));
    launch_mesh_export(planes, staging, (uint32_t)n, nverts, ncomp, c->stream);
    cudaError_t e = cudaMemcpyAsync(host, staging, sizeof(double) * n * nverts * ncomp, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(staging);
    if (e != cudaSuccess) return fail(c, FGL_E_CUDA, "mesh read: %s", cudaGetErrorString(e));
    return FGL_OK;
}

int fgl_mesh_read(fgl_ctx *c, const fgl_mesh *m, double *position, double *normal, double *lposition, double *lnormal) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!m) return fail(c, FGL_E_INVALID, "null mesh");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_mesh_read");
    mesh_acquire(c, m);
    rc = read_attr(c, m->tpos, position, m->nt, 3, 3);
    if (!rc) rc = read_attr(c, m->tnrm, normal, m->nt, 3, 3);
    if (!rc) rc = read_attr(c, m->lpos, lposition, m->nl, 2, 3);
    if (!rc) rc = read_attr(c, m->lnrm, lnormal, m->nl, 2, 3);
    return rc;
}

// ---- textures ---------------------------------------------------------------------------------

int fgl_texture_create(fgl_ctx *c, const uint8_t *rgba8, int width, int height, int format, fgl_tex **out) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!rgba8 || !out || width <= 0 || height <= 0) return fail(c, FGL_E_INVALID, "bad texture arguments");
    if (format != FGL_TEX_RGBA && format != FGL_TEX_NRGBA && format != FGL_TEX_RGBA64)
        return fail(c, FGL_E_INVALID, "bad texture format %d", format);
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_texture_create");
    fgl_tex *t = new (std::nothrow) fgl_tex();
    if (!t) return fail(c, FGL_E_OOM, "host allocation failed");
    t->device = c->device; t->w = width; t->h = height; t->format = format; t->pixels = nullptr;
    const size_t bytes = (size_t)width * height * (format == FGL_TEX_RGBA64 ? 8 : 4);
    cudaError_t e = dev_alloc(&t->pixels, bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(t->pixels, rgba8, bytes, cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dev_free(t->pixels);
        delete t;
        return fail(c, FGL_E_CUDA, "texture upload: %s", cudaGetErrorString(e));
    }
    *out = t;
    return FGL_OK;
}

int fgl_texture_destroy(fgl_tex *t) {
    if (!t) return FGL_OK;
    cudaSetDevice(t->device);
    dev_free(t->pixels);
    delete t;
    return FGL_OK;
}

// ---- draws ----------------------------------------------------------------------------------------

int fgl_draw_triangles(fgl_ctx *c, const fgl_state *s, const fgl_shader *sh, const fgl_mesh *m, uint64_t first,
                       uint64_t count, fgl_raster_info *info) {
    return draw_common(c, s, sh, m, first, count, false, false, info);
}
int fgl_draw_lines(fgl_ctx *c, const fgl_state *s, const fgl_shader *sh, const fgl_mesh *m, uint64_t first,
                   uint64_t count, fgl_raster_info *info) {
    return draw_common(c, s, sh, m, first, count, true, false, info);
}
int fgl_draw_triangles_each(fgl_ctx *c, const fgl_state *s, const fgl_shader *sh, const fgl_mesh *m, uint64_t first,
                            uint64_t count, fgl_raster_info *infos, fgl_raster_info *info) {
    return draw_common(c, s, sh, m, first, count, false, false, info, infos, true);
}
int fgl_draw_lines_each(fgl_ctx *c, const fgl_state *s, const fgl_shader *sh, const fgl_mesh *m, uint64_t first,
                        uint64_t count, fgl_raster_info *infos, fgl_raster_info *info) {
    return draw_common(c, s, sh, m, first, count, true, false, info, infos, true);
}
int fgl_draw_triangles_async(fgl_ctx *c, const fgl_state *s, const fgl_shader *sh, const fgl_mesh *m, uint64_t first,
                             uint64_t count) {
    return draw_common(c, s, sh, m, first, count, false, true, nullptr);
}
int fgl_draw_lines_async(fgl_ctx *c, const fgl_state *s, const fgl_shader *sh, const fgl_mesh *m, uint64_t first,
                         uint64_t count) {
    return draw_common(c, s, sh, m, first, count, true, true, nullptr);
}

int fgl_sync(fgl_ctx *c, fgl_raster_info *info) {
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_sync");
    if (info) { info->total_pixels = 0; info->updated_pixels = 0; }
    fb_join(c);
    if (c->async_pending) {
        CK(c, cudaMemcpyAsync(c->host_counters, c->acc_dev, sizeof(DrawCounters), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemsetAsync(c->acc_dev, 0, sizeof(DrawCounters), c->stream));
    }
    CK(c, cudaStreamSynchronize(c->stream));
    if (c->read_stream) CK(c, cudaStreamSynchronize(c->read_stream));  // read-backs of earlier fgl_frame_end calls
    if (c->async_pending) {
        c->async_pending = false;
        const DrawCounters hc = *c->host_counters;
        if (hc.overflow) {
            // grow so that re-issuing the frame succeeds
            rc = ensure_work(c, grown(c, hc));
            if (rc) return rc;  // (out of memory while regrowing: re-issuing would not help)
            return fail(c, FGL_E_OVERFLOW, "an async draw outgrew its work buffers (now regrown): re-issue the frame");
        }
        if (info) { info->total_pixels = hc.total_pixels; info->updated_pixels = hc.updated_pixels; }
    }
    return FGL_OK;
}

int fgl_frame_end(fgl_ctx *c, uint8_t *color_dst, size_t stride, fgl_fence **fence) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!fence) return fail(c, FGL_E_INVALID, "null fence pointer");
    if (stride == 0) stride = (size_t)c->w * 4;
    if (stride < (size_t)c->w * 4) return fail(c, FGL_E_INVALID, "stride too small");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_frame_end");
    fgl_fence *f = *fence;
    if (!f) {
        f = new (std::nothrow) fgl_fence();
        if (!f) return fail(c, FGL_E_OOM, "host allocation failed");
        f->device = c->device; f->counters = nullptr; f->recorded = false;
        cudaError_t e = cudaEventCreateWithFlags(&f->done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void **>(&f->counters), sizeof(DrawCounters));
        if (e != cudaSuccess) {
            cudaGetLastError();
            delete f;
            return fail(c, FGL_E_CUDA, "fence: %s", cudaGetErrorString(e));
        }
        *fence = f;
    } else if (f->device != c->device) {
        return fail(c, FGL_E_INVALID, "fence belongs to another device");
    }
    fb_join(c);
    // Overlapped read-back: stage the frame on the device, send it to the host from the read stream.  The first use
    // creates the stream, two staging slots and their events; if that fails the frame is read back on the draw stream.
    const size_t row_bytes = (size_t)c->w * 4, fb_bytes = row_bytes * c->h;
    int slot = -1;
    if (c->rb_overlap) {
        slot = c->rb_next;
        cudaError_t e = cudaSuccess;
        if (!c->read_stream) e = cudaStreamCreateWithFlags(&c->read_stream, cudaStreamNonBlocking);
        if (e == cudaSuccess && !c->ev_rb_staged[slot]) e = cudaEventCreateWithFlags(&c->ev_rb_staged[slot], cudaEventDisableTiming);
        if (e == cudaSuccess && !c->ev_rb_done[slot]) e = cudaEventCreateWithFlags(&c->ev_rb_done[slot], cudaEventDisableTiming);
        if (e == cudaSuccess && !c->rb_counters[slot]) e = dev_alloc(&c->rb_counters[slot], 1);
        if (e == cudaSuccess && color_dst && !c->rb_color[slot]) e = dev_alloc(&c->rb_color[slot], (size_t)c->w * c->h);
        if (e != cudaSuccess) { cudaGetLastError(); slot = -1; c->rb_overlap = false; }
    }
    if (slot >= 0) {
        c->rb_next = slot ^ 1;
        if (c->rb_done_recorded[slot]) CK(c, cudaStreamWaitEvent(c->stream, c->ev_rb_done[slot], 0));  // the slot's last transfer
        if (color_dst) CK(c, cudaMemcpyAsync(c->rb_color[slot], c->color, fb_bytes, cudaMemcpyDeviceToDevice, c->stream));
        CK(c, cudaMemcpyAsync(c->rb_counters[slot], c->acc_dev, sizeof(DrawCounters), cudaMemcpyDeviceToDevice, c->stream));
        CK(c, cudaMemsetAsync(c->acc_dev, 0, sizeof(DrawCounters), c->stream));
        CK(c, cudaEventRecord(c->ev_rb_staged[slot], c->stream));
        CK(c, cudaStreamWaitEvent(c->read_stream, c->ev_rb_staged[slot], 0));
        if (color_dst) {
            if (stride == row_bytes)  // tightly packed: one linear copy (the 2-D form goes row by row)
                CK(c, cudaMemcpyAsync(color_dst, c->rb_color[slot], fb_bytes, cudaMemcpyDeviceToHost, c->read_stream));
            else
                CK(c, cudaMemcpy2DAsync(color_dst, stride, c->rb_color[slot], row_bytes, row_bytes, c->h, cudaMemcpyDeviceToHost,
                                        c->read_stream));
        }
        CK(c, cudaMemcpyAsync(f->counters, c->rb_counters[slot], sizeof(DrawCounters), cudaMemcpyDeviceToHost, c->read_stream));
        CK(c, cudaEventRecord(c->ev_rb_done[slot], c->read_stream));
        c->rb_done_recorded[slot] = true;
        CK(c, cudaEventRecord(f->done, c->read_stream));
    } else {
        if (color_dst) {
            if (stride == row_bytes)
                CK(c, cudaMemcpyAsync(color_dst, c->color, fb_bytes, cudaMemcpyDeviceToHost, c->stream));
            else
                CK(c, cudaMemcpy2DAsync(color_dst, stride, c->color, row_bytes, row_bytes, c->h, cudaMemcpyDeviceToHost, c->stream));
        }
        CK(c, cudaMemcpyAsync(f->counters, c->acc_dev, sizeof(DrawCounters), cudaMemcpyDeviceToHost, c->stream));
        CK(c, cudaMemsetAsync(c->acc_dev, 0, sizeof(DrawCounters), c->stream));
        CK(c, cudaEventRecord(f->done, c->stream));
    }
    f->recorded = true;
    c->async_pending = false;
    return FGL_OK;
}

int fgl_fence_wait(fgl_ctx *c, fgl_fence *f, fgl_raster_info *info) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!f) return fail(c, FGL_E_INVALID, "null fence");
    if (info) { info->total_pixels = 0; info->updated_pixels = 0; }
    if (!f->recorded) return FGL_OK;
    CK(c, cudaEventSynchronize(f->done));
    f->recorded = false;
    const DrawCounters hc = *f->counters;
    if (hc.overflow) {
        std::lock_guard<std::mutex> lock(c->mu);
        rc = ensure_work(c, grown(c, hc));
        if (rc) return rc;
        return fail(c, FGL_E_OVERFLOW, "an async draw of this frame outgrew its work buffers (now regrown): re-issue the frame");
    }
    if (info) { info->total_pixels = hc.total_pixels; info->updated_pixels = hc.updated_pixels; }
    return FGL_OK;
}

int fgl_fence_destroy(fgl_fence *f) {
    if (!f) return FGL_OK;
    cudaSetDevice(f->device);
    cudaEventSynchronize(f->done);
    cudaEventDestroy(f->done);
    if (f->counters) cudaFreeHost(f->counters);
    delete f;
    return FGL_OK;
}

int fgl_get_draw_stats(const fgl_ctx *c, fgl_draw_stats *out) {
    if (!c || !out) return fail(nullptr, FGL_E_INVALID, "null argument");
    *out = c->stats;
    return FGL_OK;
}

int fgl_set_profiling(fgl_ctx *c, int enabled) {
    int rc = check_ctx(c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_set_profiling");
    if (enabled && !c->prof_created) {
        for (int i = 0; i < PROF_RING; i++)
            for (int k = 0; k < PROF_EVENTS; k++) CK(c, cudaEventCreate(&c->prof[i].e[k]));
        c->prof_created = true;
    }
    if (!enabled) prof_drain(c);
    c->profiling = enabled != 0;
    return FGL_OK;
}

int fgl_get_stage_times(fgl_ctx *c, fgl_stage_times *out) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!out) return fail(c, FGL_E_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    prof_drain(c);
    *out = c->prof_acc;
    memset(&c->prof_acc, 0, sizeof c->prof_acc);
    return FGL_OK;
}

// ---- read-back / upload ---------------------------------------------------------------------------

int fgl_read_color(fgl_ctx *c, uint8_t *dst, size_t stride) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!dst) return fail(c, FGL_E_INVALID, "null destination");
    if (stride == 0) stride = (size_t)c->w * 4;
    if (stride < (size_t)c->w * 4) return fail(c, FGL_E_INVALID, "stride too small");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_read_color");
    fb_join(c);
    if (stride == (size_t)c->w * 4)
        CK(c, cudaMemcpyAsync(dst, c->color, (size_t)c->w * 4 * c->h, cudaMemcpyDeviceToHost, c->stream));
    else
        CK(c, cudaMemcpy2DAsync(dst, stride, c->color, (size_t)c->w * 4, (size_t)c->w * 4, c->h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}
int fgl_read_depth(fgl_ctx *c, double *dst) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!dst) return fail(c, FGL_E_INVALID, "null destination");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_read_depth");
    fb_join(c);
    CK(c, cudaMemcpyAsync(dst, c->depth, sizeof(double) * c->w * c->h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}
int fgl_depth_image(fgl_ctx *c, uint16_t *dst) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!dst) return fail(c, FGL_E_INVALID, "null destination");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_depth_image");
    const size_t npix = (size_t)c->w * c->h;
    if (!c->gray16) CK(c, dev_alloc(&c->gray16, npix));
    fb_join(c);
    launch_depth_image(c->depth, npix, c->gray16, c->scratch, c->stream);
    CK(c, cudaGetLastError());
    CK(c, cudaMemcpyAsync(dst, c->gray16, npix * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}

int fgl_write_color(fgl_ctx *c, const uint8_t *src, size_t stride) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!src) return fail(c, FGL_E_INVALID, "null source");
    if (stride == 0) stride = (size_t)c->w * 4;
    if (stride < (size_t)c->w * 4) return fail(c, FGL_E_INVALID, "stride too small");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_write_color");
    fb_join(c);
    CK(c, cudaMemcpy2DAsync(c->color, (size_t)c->w * 4, src, stride, (size_t)c->w * 4, c->h, cudaMemcpyHostToDevice, c->stream));
    c->clear_color_known = false;
    CK(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}
int fgl_write_depth(fgl_ctx *c, const double *src) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!src) return fail(c, FGL_E_INVALID, "null source");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_write_depth");
    fb_join(c);
    CK(c, cudaMemcpyAsync(c->depth, src, sizeof(double) * c->w * c->h, cudaMemcpyHostToDevice, c->stream));
    CK(c, cudaMemsetAsync(c->wb.dirty, 1, c->wb.ntiles, c->stream));  // any strip may hold drawn depth now
    c->clear_depth_value = __builtin_nan("");
    CK(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}

// ---- SSAA resolve --------------------------------------------------------------------------------

int fgl_resolve_device(fgl_ctx *c, int factor) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (factor < 1 || factor > 16 || c->w % factor || c->h % factor)
        return fail(c, FGL_E_INVALID, "resolve factor %d must be in [1,16] and divide %dx%d", factor, c->w, c->h);
    std::lock_guard<std::mutex> lock(c->mu);
    const int dw = c->w / factor, dh = c->h / factor;
    fb_join(c);
    if (dw != c->rw || dh != c->rh) {
        if (c->capturing) return fail(c, FGL_E_INVALID, "resolve once with this factor before recording (its buffer is allocated on first use)");
        c->buffer_epoch++;
        dev_free(c->resolved);
        CK(c, dev_alloc(&c->resolved, (size_t)dw * dh));
        c->rw = dw; c->rh = dh;
    }
    if (factor == 1) {
        // resize.Resize returns its input unchanged when the size already matches
        CK(c, cudaMemcpyAsync(c->resolved, c->color, (size_t)dw * dh * 4, cudaMemcpyDeviceToDevice, c->stream));
    } else {
        launch_resolve(c->color, c->w, c->h, c->resolved, factor, c->stream);
        CK(c, cudaGetLastError());
    }
    return FGL_OK;
}
int fgl_read_resolved(fgl_ctx *c, uint8_t *dst) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!dst) return fail(c, FGL_E_INVALID, "null destination");
    if (!c->resolved) return fail(c, FGL_E_INVALID, "nothing resolved yet");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_read_resolved");
    CK(c, cudaMemcpyAsync(dst, c->resolved, (size_t)c->rw * c->rh * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    return FGL_OK;
}
int fgl_resolve(fgl_ctx *c, int factor, uint8_t *dst) {
    int rc = fgl_resolve_device(c, factor);
    if (rc) return rc;
    return fgl_read_resolved(c, dst);
}

// ---- composite -----------------------------------------------------------------------------------------

int fgl_composite_pack(fgl_ctx *c, void *keys_dev) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!keys_dev) return fail(c, FGL_E_INVALID, "null key buffer");
    std::lock_guard<std::mutex> lock(c->mu);
    fb_join(c);
    launch_composite_pack(c->color, c->depth, static_cast<unsigned long long *>(keys_dev), (size_t)c->w * c->h, c->stream);
    CK(c, cudaGetLastError());
    return FGL_OK;
}
int fgl_composite_unpack(fgl_ctx *c, const void *keys_dev) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!keys_dev) return fail(c, FGL_E_INVALID, "null key buffer");
    std::lock_guard<std::mutex> lock(c->mu);
    fb_join(c);
    launch_composite_unpack(c->color, c->depth, static_cast<const unsigned long long *>(keys_dev), (size_t)c->w * c->h, c->stream);
    cudaMemsetAsync(c->wb.dirty, 1, c->wb.ntiles, c->stream);
    CK(c, cudaGetLastError());
    return FGL_OK;
}
int fgl_composite_min(fgl_ctx *c, void *inout, const void *other, uint64_t count) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!inout || !other) return fail(c, FGL_E_INVALID, "null key buffer");
    std::lock_guard<std::mutex> lock(c->mu);
    launch_composite_min(static_cast<unsigned long long *>(inout), static_cast<const unsigned long long *>(other), count, c->stream);
    CK(c, cudaGetLastError());
    return FGL_OK;
}

int fgl_ipc_export(fgl_ctx *c, void *color_handle, void *depth_handle) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!color_handle || !depth_handle) return fail(c, FGL_E_INVALID, "null handle buffer");
    static_assert(sizeof(cudaIpcMemHandle_t) == FGL_IPC_HANDLE_BYTES, "IPC handle size");
    cudaIpcMemHandle_t hc, hd;
    CK(c, cudaIpcGetMemHandle(&hc, c->color));
    CK(c, cudaIpcGetMemHandle(&hd, c->depth));
    memcpy(color_handle, &hc, sizeof hc);
    memcpy(depth_handle, &hd, sizeof hd);
    return FGL_OK;
}

int fgl_ipc_open(fgl_ctx *c, const void *color_handle, const void *depth_handle, void **color_ptr, void **depth_ptr) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!color_handle || !depth_handle || !color_ptr || !depth_ptr) return fail(c, FGL_E_INVALID, "null argument");
    cudaIpcMemHandle_t hc, hd;
    memcpy(&hc, color_handle, sizeof hc);
    memcpy(&hd, depth_handle, sizeof hd);
    CK(c, cudaIpcOpenMemHandle(color_ptr, hc, cudaIpcMemLazyEnablePeerAccess));
    CK(c, cudaIpcOpenMemHandle(depth_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return FGL_OK;
}

int fgl_ipc_close(fgl_ctx *c, void *color_ptr, void *depth_ptr) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (color_ptr) CK(c, cudaIpcCloseMemHandle(color_ptr));
    if (depth_ptr) CK(c, cudaIpcCloseMemHandle(depth_ptr));
    return FGL_OK;
}

int fgl_composite_peer(fgl_ctx *c, int rank, int nranks, void *const *color, void *const *depth) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (nranks < 1 || nranks > FGL_MAX_PEERS || rank < 0 || rank >= nranks || !color || !depth)
        return fail(c, FGL_E_INVALID, "bad rank %d / nranks %d (at most %d peers)", rank, nranks, FGL_MAX_PEERS);
    for (int r = 0; r < nranks; r++)
        if (!color[r] || !depth[r]) return fail(c, FGL_E_INVALID, "null buffer pointer for rank %d", r);
    std::lock_guard<std::mutex> lock(c->mu);
    const size_t npix = (size_t)c->w * c->h;
    const size_t px0 = npix * (size_t)rank / nranks, px1 = npix * (size_t)(rank + 1) / nranks;
    fb_join(c);
    launch_composite_peer(reinterpret_cast<uint32_t *const *>(color), reinterpret_cast<double *const *>(depth), nranks,
                          px0, px1, c->stream);
    cudaMemsetAsync(c->wb.dirty, 1, c->wb.ntiles, c->stream);  // (the peers write into this rank's buffers as well)
    CK(c, cudaGetLastError());
    return FGL_OK;
}

int fgl_debug_tile_cycles(fgl_ctx *c, uint64_t *dst, uint64_t ntiles) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!c->wb.tile_clock) return fail(c, FGL_E_INVALID, "set FGL_TILE_CLOCK=1 before creating the context");
    if (!dst || (ntiles != c->wb.ntiles && ntiles != c->wb.ntiles + 8))  // + 8: also the 16 path counters
        return fail(c, FGL_E_INVALID, "need a buffer of 2*%u uint64", c->wb.ntiles);
    std::lock_guard<std::mutex> lock(c->mu);
    CK(c, cudaStreamSynchronize(c->stream));
    CK(c, cudaMemcpy(dst, c->wb.tile_clock, sizeof(uint64_t) * 2 * ntiles, cudaMemcpyDeviceToHost));
    return FGL_OK;
}

int fgl_debug_div_check(fgl_ctx *c, uint64_t seed, uint64_t pairs, uint64_t *mismatches, uint64_t *fast_path) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!mismatches || !fast_path) return fail(c, FGL_E_INVALID, "null result pointer");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_debug_div_check");
    CK(c, cudaMemsetAsync(c->scratch, 0, 2 * sizeof(unsigned long long), c->stream));
    launch_div_check(seed, pairs, c->scratch, c->stream);
    CK(c, cudaGetLastError());
    unsigned long long h[2] = {0, 0};
    CK(c, cudaMemcpyAsync(h, c->scratch, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(c, cudaStreamSynchronize(c->stream));
    *mismatches = h[0];
    *fast_path = h[1];
    return FGL_OK;
}

int fgl_probe_atomic_rate(fgl_ctx *c, uint64_t ops, double *ops_per_second) {
    int rc = check_ctx(c);
    if (rc) return rc;
    if (!ops_per_second || ops == 0) return fail(c, FGL_E_INVALID, "null result pointer / zero operations");
    std::lock_guard<std::mutex> lock(c->mu);
    NOT_WHILE_RECORDING(c, "fgl_probe_atomic_rate");
    const size_t words = (size_t)c->w * c->h;
    unsigned long long *buf = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = dev_alloc(&buf, words);
    if (e == cudaSuccess) e = cudaMemsetAsync(buf, 0xff, words * sizeof(unsigned long long), c->stream);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    float ms = 0;
    if (e == cudaSuccess) {
        launch_atomic_probe(buf, words, ops / 8 + 1, c->stream);  // warm-up
        cudaEventRecord(e0, c->stream);
        launch_atomic_probe(buf, words, ops, c->stream);
        cudaEventRecord(e1, c->stream);
        e = cudaEventSynchronize(e1);
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    dev_free(buf);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(c, FGL_E_CUDA, "atomic probe: %s", cudaGetErrorString(e)); }
    const unsigned long long threads = (unsigned long long)GRID_WAVE * 256ull;
    const unsigned long long done = (ops + threads - 1) / threads * threads;
    *ops_per_second = ms > 0 ? (double)done / ((double)ms * 1e-3) : 0.0;
    return FGL_OK;
}

void *fgl_stream(const fgl_ctx *c) { return c ? (void *)c->stream : nullptr; }
void *fgl_color_device_ptr(const fgl_ctx *c) { return c ? (void *)c->color : nullptr; }
void *fgl_depth_device_ptr(const fgl_ctx *c) { return c ? (void *)c->depth : nullptr; }

}  // extern "C"
