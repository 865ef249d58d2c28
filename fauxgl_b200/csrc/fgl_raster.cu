// fgl_raster.cu -- strip-parallel, ordered back end.  The framebuffer is cut into
// strips of 64 x 1 pixels; the bin of a strip is the range of the stably sorted
// segment array that carries its id.  ONE WARP owns a strip: it keeps the strip's
// depth (f64) -- and colour, when shading inline -- in shared memory, consumes the
// bin 32 segments at a time in primitive order, and writes the strip back once,
// coalesced.  This replaces the reference's per-pixel mutex array.
//
// Replaces the per-pixel body of Context.rasterize (context.go:207-273),
// InterpolateVertexes (vertex.go:18-47), the three built-in Fragment shaders
// (shader.go:25,44,75), ImageTexture.BilinearSample (texture.go:41-63),
// Color.NRGBA (color.go:56) and the depth retest / write / blend (context.go:245-273).
//
// Arithmetic parity.  Each segment carries the reference's forward-differenced
// edge values at its first pixel (fgl_span.cu); the lanes here continue the same
// `w += a` chain (context.go:211-213), so barycentrics, depth and colour are
// bit-identical to a sequential run of the reference in triangle-index order.
//
// Ordering without barriers.  Lane i of a chunk holds the i-th segment in
// primitive order.  If no two segments of the chunk share a pixel (the common
// case, one popcount test) every lane just walks its segment.  Otherwise:
//   * deferred mode: every segment lane stages the depths of its fragments and
//     registers itself in a per-pixel cover mask; then every PIXEL lane replays
//     the fragments of its pixel in lane order == primitive order against a
//     running depth in registers -- the serial part of the reference's per-pixel
//     order shrinks to one shared-memory load and a compare per fragment;
//   * inline mode: a pixel is `ready` for lane i when no earlier lane still wants
//     it (exclusive prefix-OR of the pending masks); ready pixels are resolved,
//     the masks shrink, until all are empty.
// Either way every pixel sees its fragments in primitive order, which makes `<=`
// ties, DepthBias, blending and UpdatedPixels well defined, and the only
// synchronisation is __syncwarp.  (An earlier version resolved 256 segments per
// CTA in rounds separated by __syncthreads with shared-memory tickets: 35 cycles
// per segment, and the heaviest 64x4 tile of the benchmark frame WAS the kernel's
// duration, profiles/README.md.)
//
// Deferred shading.  When the draw's shader can neither discard nor blend
// (SolidColor, or Phong with an ObjectColor and no texture, alpha != 0 and no
// effective blending) the fragment colour cannot influence any depth decision:
// the strip kernel settles depth only and records, per pixel, WHICH segment won;
// k_shade then colours each pixel's FINAL winner once, pixel-parallel at full
// occupancy: it re-walks the winner's chain of adds from the segment start to the
// pixel (a few DADDs) and evaluates Phong.  Otherwise fragments are shaded inline,
// in order, exactly like the reference.
#include "fgl_internal.h"
#include "fgl_block.cuh"
#include "fgl_math.cuh"
#include "fgl_shade.cuh"

#include <type_traits>

namespace fgl {

// CTAs per SM of the deferred strip kernel.  At 4 (64 registers) the compiler spills values that were just loaded,
// and a spill store is a consumer -- it waits for the load the pipeline wanted to leave in flight (heaviest strip of
// the benchmark frame 69.5 k cycles at 4 CTAs/SM, 42.4 k at 3; k_strip 36.7 us at 3, 28.7 us at 2 with no spill at all;
// at 7680x4320, where the kernel is throughput-bound, 2 and 3 measure the same).
#ifndef FGL_STRIP_MINB
#define FGL_STRIP_MINB 2
#endif
#ifndef FGL_HEAVY_SKIP
#define FGL_HEAVY_SKIP 4  // rounds a warp sits out after a heavy strip (strip stage at 1080p: 74 / 68 / 64 us for 0 / 2 / 3 in round 1;
                          // raster stage 62.6 / 61.4 / 61.6 / 63.4 / 62.6 us for 3 / 4 / 5 / 6 / 8 on the final build of round 2)
#endif
constexpr uint32_t HEAVY_SKIP = FGL_HEAVY_SKIP;
#ifndef FGL_COOP_FACTOR
#define FGL_COOP_FACTOR 1  // heavy strips go to whole CTAs while there are at most this many per CTA
#endif
#ifndef FGL_STRIP_PHASES
#define FGL_STRIP_PHASES 0  // tuning aid (variant build): cycles per phase of the chunks of heavy strips -> tile_clock[2 * ntiles + 10 ..]
#endif
#if FGL_STRIP_PHASES
#define PHASE_CLOCK(var, dep) long long var; { uint32_t dep_ = (uint32_t)(dep); asm volatile("mov.u32 %0, %0;" : "+r"(dep_)); __syncwarp(); var = clock64(); }
#define PHASE_SET(var, dep) { uint32_t dep_ = (uint32_t)(dep); asm volatile("mov.u32 %0, %0;" : "+r"(dep_)); __syncwarp(); var = clock64(); }
#else
#define PHASE_CLOCK(var, dep)
#define PHASE_SET(var, dep)
#endif
constexpr int SWARPS = 8;               // warps (= strips in flight) per CTA
constexpr int STHREADS = SWARPS * 32;
constexpr uint32_t NO_WINNER = 0xffffffffu;
constexpr uint32_t NO_SEG = 0xffffffffu;
static_assert(TILE_W == 64, "a strip is one 64-bit pending mask wide at most");
constexpr int STRIP_H = TILE_W / 32;  // pixels per lane at most (p.tile_w / 32 are used)

constexpr int ZCAP = 256;  // fragments of a chunk the pixel-parallel resolution can stage
template <bool DEFERRED> struct StripMem;
// Gather of a chunk's 32 segment heads (deferred mode).  Eight lanes per record, six of them copy one 16-byte
// piece each with cp.async straight into shared memory: one warp instruction touches 4 cache lines instead of 32,
// no register holds data in flight, and the copy of the NEXT chunk of the warp's stream (the next 32 segments of
// this strip, or the first 32 of its next strip) runs while the current one is resolved -- the segment records
// are scattered over the whole array and come from DRAM (ncu: 99 % L2 misses, profiles/README.md), and a dependent
// gather per chunk was the strip kernel's critical path.  Every lane then reads its own record back.
constexpr int REC_STRIDE16 = 6;  // record stride in the staging buffer, in 16-byte units (2-way bank conflict on the read)
template <> struct StripMem<true> {
    double zbuf[ZCAP];                   // depth of every fragment of the chunk, segment-major
    uint4 recbuf[32 * REC_STRIDE16];     // staging of the gather: refilled as soon as the lanes hold their records
    double depth[TILE_W];
    uint32_t winseg[TILE_W]; // segment (index into segv) whose fragment currently owns the pixel
    // pixel-parallel resolution of a chunk whose segments overlap
    uint32_t cover[TILE_W];  // lanes (= segments, in primitive order) that cover the pixel
    uint32_t segidx[32];     // segv index of the lane's segment
    uint32_t upd[32];        // UpdatedPixels per segment (per-primitive RasterizeInfo only)
    int16_t segoff[32];      // first zbuf slot of the segment - its first pixel: fragment of pixel i = zbuf[segoff + i]
};
template <> struct StripMem<false> {
    double depth[TILE_W];
    uint32_t color[TILE_W];
};

// Exclusive prefix-OR over the lanes of a warp.
FGL_DI unsigned long long warp_excl_or(unsigned long long v, int lane) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t tl = __shfl_up_sync(0xffffffffu, lo, o), th = __shfl_up_sync(0xffffffffu, hi, o);
        if (lane >= o) { lo |= tl; hi |= th; }
    }
    lo = __shfl_up_sync(0xffffffffu, lo, 1);
    hi = __shfl_up_sync(0xffffffffu, hi, 1);
    if (lane == 0) lo = hi = 0;
    return ((unsigned long long)hi << 32) | lo;
}

// One fragment, inline mode: context.go:229-273 for strip pixel pi.
FGL_DI void fragment_inline(const DrawParams &p, const fgl_state &st, const WorkBuffers &wb, const SegV &v,
                            double w0, double w1, double w2, int pi, StripMem<false> &sm, unsigned long long &updated) {
    const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
    const double z = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
    const double bz = z + st.depth_bias;
    const double dcur = sm.depth[pi];
    if (st.read_depth && bz > dcur) return;               // context.go:232
    const double bx = b0 * v.r0, by = b1 * v.r1, bzz = b2 * v.r2;  // context.go:236
    const double bw = 1 / (bx + by + bzz);
    AttrSrc a{&p, wb.clip_pool, v.src, v.flags};
    const C4 color = shade_fragment(p, a, bx, by, bzz, bw);
    if (c_is_discard(color)) return;                      // context.go:241
    if (bz <= dcur || !st.read_depth) {                   // context.go:248
        updated++;
        if (st.write_depth) sm.depth[pi] = z;
        if (st.write_color) {
            const uint32_t c8 = c_nrgba(color);
            if (st.alpha_blend && color.a < 1) sm.color[pi] = blend_over(sm.color[pi], c8);  // PixOffset: lands where the index does
            else if (!(v.flags & REC_WRAP)) sm.color[pi] = c8;  // SetNRGBA drops a pixel whose own x is outside the image, context.go:269
        }
    }
}

// Deferred mode, one fragment whose depth z is known: context.go:232 early-out, then (no discard possible)
// the retest at :248; the winning segment is remembered for the shading kernel.
FGL_DI void resolve_deferred(const fgl_state &st, StripMem<true> &sm, int pi, double z, uint32_t seg, bool wrapped,
                             unsigned long long &updated) {
    const double bz = z + st.depth_bias;
    const double dcur = sm.depth[pi];
    if (!(st.read_depth && bz > dcur) && (bz <= dcur || !st.read_depth)) {
        updated++;
        if (st.write_depth) sm.depth[pi] = z;
        // (a fragment that aliased in from a neighbouring row wins the depth but its opaque colour is dropped by
        // SetNRGBA, context.go:269: the pixel keeps the colour of the last winner that was drawn in its own row)
        if (!wrapped) sm.winseg[pi] = seg;
    }
}

FGL_DI uint2 busy_at(const WorkBuffers &wb, uint32_t nheavy, uint32_t q) {
    return q < nheavy ? wb.busy_list[q] : wb.busy_list[wb.ntiles - 1u - (q - nheavy)];
}
FGL_DI void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }
#ifndef FGL_SEG_PREFETCH
#define FGL_SEG_PREFETCH 1  // sectors of a segment record prefetched into L2 ahead of its chunk (tuning aid: 0, 1, 4)
#endif
#ifndef FGL_SEG_LOAD
#define FGL_SEG_LOAD 0      // 0: plain loads, 1: ld.global.cg (L2 only), 2: ld.global.nc (tuning aid)
#endif
FGL_DI void prefetch_seg(const SegV *sp) {
#if FGL_SEG_PREFETCH >= 1
    prefetch_l2(sp);
#endif
#if FGL_SEG_PREFETCH >= 4
    prefetch_l2(reinterpret_cast<const char *>(sp) + 32);
    prefetch_l2(reinterpret_cast<const char *>(sp) + 64);
    prefetch_l2(reinterpret_cast<const char *>(sp) + 96);
#endif
}
FGL_DI void gather_issue(uint4 *buf, const SegV *__restrict__ segv, uint32_t idx, int lane) {
    const int piece = lane & 7, sub = lane >> 3;
#pragma unroll
    for (int r = 0; r < 8; r++) {
        const uint32_t ridx = __shfl_sync(0xffffffffu, idx, r * 4 + sub);
        if (ridx != NO_SEG && piece < 6) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(buf + (r * 4 + sub) * REC_STRIDE16 + piece);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(reinterpret_cast<const uint4 *>(segv + ridx) + piece)
                         : "memory");
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
}
FGL_DI void gather_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
FGL_DI SegHead gather_take(const uint4 *buf, uint32_t idx, int lane, int x0) {
    SegHead h;
    uint4 *hp = reinterpret_cast<uint4 *>(&h);
#pragma unroll
    for (int k = 0; k < 6; k++) hp[k] = buf[lane * REC_STRIDE16 + k];
    if (idx == NO_SEG) { h.cnt = 0; h.x = (uint16_t)x0; }
    return h;
}
FGL_DI SegHead load_head(const SegV *sp) {  // lane-per-record load (rare paths)
    SegHead h;
    uint4 *hp = reinterpret_cast<uint4 *>(&h);
#pragma unroll
    for (int k = 0; k < 6; k++) hp[k] = reinterpret_cast<const uint4 *>(sp)[k];
    return h;
}
FGL_DI SegV load_seg(const SegV *sp) {
#if FGL_SEG_LOAD == 0
    return *sp;
#else
    SegV v;
    const uint4 *s = reinterpret_cast<const uint4 *>(sp);
    uint4 *d = reinterpret_cast<uint4 *>(&v);
#pragma unroll
    for (int k = 0; k < 8; k++) d[k] = FGL_SEG_LOAD == 1 ? __ldcg(s + k) : __ldg(s + k);
    return v;
#endif
}

// EACH: also attribute UpdatedPixels to the primitive of every segment (fgl_draw_*_each).
//
// Heavy strips (deferred mode).  A bin of hundreds of segments -- the poles of a parametric surface put 400+
// sub-pixel slivers into one 32-pixel strip -- used to be one warp's serial chain of 32-segment steps and WAS the
// kernel's duration (profiles/README.md).  The expensive part of a step, gathering the segments and walking their
// depths into shared memory, does not depend on order; only the per-pixel replay does.  So the strips at the head
// of the busy list (>= HEAVY_SEGS segments) are taken by whole CTAs: the eight warps stage eight consecutive
// 32-segment chunks at once, then warp 0 replays the eight chunks in order, one lane per pixel.  Same fragments,
// same order, same tests -- for every render state.
template <bool DEFERRED, bool EACH>
__global__ void __launch_bounds__(STHREADS, DEFERRED ? FGL_STRIP_MINB : 3)
k_strip(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb,
        const uint32_t *__restrict__ seg_order, const uint32_t *__restrict__ seg_keys, uint32_t *__restrict__ gcolor,
        double *__restrict__ gdepth) {
    pdl_wait();
    pdl_trigger();
    if (wb.counters->overflow) return;  // work buffers too small: the host regrows and re-issues the draw

    extern __shared__ __align__(16) unsigned char s_raw[];  // SWARPS x StripMem (over the 48 KB static limit when deferred)
    StripMem<DEFERRED> *const s_all = reinterpret_cast<StripMem<DEFERRED> *>(s_raw);
    __shared__ uint32_t s_nseg;          // heavy strips: segments of the strip (tuning aid)
    __shared__ uint32_t s_flag[SWARPS];  // heavy strips: 0 = the warp's chunk is empty, 1 = staged, 2 = too many fragments to stage
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    StripMem<DEFERRED> &sm = s_all[warp];
    const fgl_state st = p.state;
    unsigned long long my_updated = 0;
    const int tile_w = p.tile_w, strip_h = tile_w >> 5;  // strip width of this context (32 or 64), pixels per lane
    if constexpr (DEFERRED) {  // scratch of the pixel-parallel resolution: every use leaves it clean
        for (int h = 0; h < STRIP_H; h++) sm.cover[lane + 32 * h] = 0;
        sm.upd[lane] = 0;
        __syncwarp();
    }
    const uint32_t nsegs = min(wb.counters->n_segs, wb.cap_segs);
    const uint32_t nheavy = wb.tile_ctl->nheavy, nbusy = nheavy + wb.tile_ctl->nlight;
    // (key, segment index) of sorted position pos, raw: nothing here consumes the loaded values, so the loads stay
    // in flight until pair_idx() resolves them an iteration (or a strip) later
    auto load_raw = [&](uint32_t pos) {
        uint2 r = make_uint2(0xffffffffu, 0u);
        if (pos < nsegs) { r.x = seg_keys[pos]; r.y = seg_order[pos]; }
        return r;
    };
    auto pair_idx = [&](uint32_t strip, uint2 raw) { return (raw.x == strip && strip != 0xffffffffu) ? raw.y : NO_SEG; };
    // (segment index, valid) of sorted position pos for a bin of `strip`
    auto load_pair = [&](uint32_t strip, uint32_t pos) {
        uint32_t idx = NO_SEG;
        if (strip != 0xffffffffu && pos < nsegs) {  // two independent loads, then the select
            const uint32_t k = seg_keys[pos], o = seg_order[pos];
            if (k == strip) idx = o;
        }
        return idx;
    };

    // ---- building blocks of a 32-segment chunk ---------------------------------------------------------------
    // (1) every segment lane stages the depths of its fragments in d.zbuf and registers itself in the cover mask
    // of its pixels
    auto stage_chunk = [&](auto &d, const auto &v, uint32_t idx, int xa, int cnt) {
        const uint32_t fbase = warp_incl_scan((uint32_t)cnt) - (uint32_t)cnt;
        d.segoff[lane] = (int16_t)((int)fbase - xa);
        d.segidx[lane] = idx;
        double w0 = v.w0, w1 = v.w1, w2 = v.w2;
        for (int k = 0; k < cnt; k++) {
            const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
            d.zbuf[fbase + k] = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
            atomicOr(&d.cover[xa + k], 1u << lane);
            w0 += v.a12; w1 += v.a20; w2 += v.a01;  // context.go:211-213
        }
    };
    // (2) every PIXEL lane replays the staged fragments of its pixel in lane order == primitive order against a
    // running depth in registers; the strip's depth and winners live in t
    auto replay_chunk = [&](auto &d, auto &t) {
#pragma unroll
        for (int h = 0; h < strip_h; h++) {
            const int pix = lane + 32 * h;
            uint32_t m = d.cover[pix];
            if (m) {
                d.cover[pix] = 0;
                double dcur = t.depth[pix];
                int win = -1;
                // Four fragments at a time: their depths are fetched first (independent loads, two shared-memory
                // levels each), then tested in order -- the serial chain per fragment is the compare alone.  At
                // the poles of a parametric surface hundreds of fragments pile onto one pixel.
                while (m) {
                    int jj[4];
                    double zz[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        jj[u] = m ? __ffs(m) - 1 : -1;
                        m &= m - 1;
                    }
                    int off[4];  // (missing fragments read segment 0's slot: any staged value, never tested)
#pragma unroll
                    for (int u = 0; u < 4; u++) off[u] = (int)d.segoff[jj[u] & 31];
#pragma unroll
                    for (int u = 0; u < 4; u++) zz[u] = d.zbuf[jj[u] >= 0 ? off[u] + pix : 0];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const double z = zz[u];
                        const double bz = z + st.depth_bias;
                        // context.go:232 early-out, then (no discard possible) the retest at :248
                        if (jj[u] >= 0 && !(st.read_depth && bz > dcur) && (bz <= dcur || !st.read_depth)) {
                            my_updated++;
                            if (st.write_depth) dcur = z;
                            win = jj[u];
                            if constexpr (EACH) atomicAdd(&d.upd[jj[u]], 1u);
                        }
                    }
                }
                if (win >= 0) {
                    if (st.write_depth) t.depth[pix] = dcur;
                    t.winseg[pix] = d.segidx[win];
                }
            }
        }
    };
    // per-primitive UpdatedPixels of a staged chunk, by the lanes that hold its segments
    auto flush_each = [&](auto &d, const auto &v) {
        const uint32_t u = d.upd[lane];
        d.upd[lane] = 0;
        if (u) atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, v.src, v.flags) + 1], (unsigned long long)u);
    };
    // (3) without staging: lanes walk their segments straight against the strip t; overlapping segments take turns
    // (a pixel is ready for lane i when no earlier lane still wants it)
    auto direct_chunk = [&](auto &t, const auto &v, uint32_t idx, int xa, unsigned long long pend, bool disjoint) {
        const unsigned long long updated_before = my_updated;
        while (true) {
            const unsigned long long ready = disjoint ? pend : (pend & ~warp_excl_or(pend, lane));
            if (ready) {
                double w0 = v.w0, w1 = v.w1, w2 = v.w2;
                const int last = 63 - __clzll((long long)ready);
                for (int pi = xa; pi <= last; pi++) {
                    if ((ready >> pi) & 1ull) {
                        if constexpr (DEFERRED) {
                            const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
                            const double z = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
                            resolve_deferred(st, t, pi, z, idx, (v.flags & REC_WRAP) != 0, my_updated);
                        } else {
                            fragment_inline(p, st, wb, v, w0, w1, w2, pi, t, my_updated);
                        }
                    }
                    w0 += v.a12; w1 += v.a20; w2 += v.a01;  // context.go:211-213
                }
                pend &= ~ready;
            }
            __syncwarp();
            if (FGL_STRIP_PHASES && wb.tile_clock && lane == 0 && !disjoint) atomicAdd(&wb.tile_clock[2 * (size_t)wb.ntiles + 8], 1ull);
            if (disjoint || !__any_sync(0xffffffffu, pend != 0)) break;
        }
        if (EACH && my_updated != updated_before)
            atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, v.src, v.flags) + 1], my_updated - updated_before);
    };
    auto pend_mask = [](int xa, int cnt) {
        return cnt > 0 ? ((cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull)) << xa) : 0ull;
    };

    if (wb.tile_clock && threadIdx.x == 0 && blockIdx.x == 0) wb.tile_clock[2 * (size_t)wb.ntiles + 9] = nheavy;  // tuning aid
    // ---- heavy strips: one CTA per strip ------------------------------------------------------------------------
    uint32_t light0 = 0;  // first entry of the per-warp part of the busy list
    // Only when every CTA gets at most one: with thousands of long bins (large triangles at 8K) there is enough
    // strip-level parallelism and eight independent warps beat eight warps in lock step.
    if constexpr (DEFERRED) {
    if (nheavy <= (uint32_t)FGL_COOP_FACTOR * gridDim.x) {
        light0 = nheavy;
        auto &t = s_all[0];  // the strip: depth and winners
        for (uint32_t hq = blockIdx.x; hq < nheavy; hq += gridDim.x) {
            const uint2 e = wb.busy_list[hq];
            const uint32_t strip = e.x, bin_beg = e.y;
            const long long t_begin = wb.tile_clock ? clock64() : 0;
            const int x0 = (int)(strip % (uint32_t)p.tiles_x) * tile_w;
            const int y = (int)(strip / (uint32_t)p.tiles_x);
            const int tw = min(tile_w, p.width - x0);
            const size_t grow = (size_t)y * p.width + x0;
            __syncthreads();  // the previous strip of this CTA is written back
            if (threadIdx.x == 0) s_nseg = 0;
            if (warp == 0) {
#pragma unroll
                for (int h = 0; h < strip_h; h++) {
                    const int i = lane + 32 * h;
                    t.winseg[i] = NO_WINNER;
                    if (i < tw) t.depth[i] = gdepth[grow + i];
                }
            }
            const unsigned long long updated_at_start = my_updated;
            uint32_t idx = load_pair(strip, bin_beg + warp * 32 + lane);
            gather_issue(sm.recbuf, wb.segv, idx, lane);
            for (uint32_t round = bin_beg;; round += SWARPS * 32) {
                const uint32_t idx_next = load_pair(strip, round + SWARPS * 32 + warp * 32 + lane);
                gather_wait();
                __syncwarp();
                const SegHead v = gather_take(sm.recbuf, idx, lane, x0);
                __syncwarp();
                gather_issue(sm.recbuf, wb.segv, idx_next, lane);  // the warp's chunk of the next round, while this one is resolved
                const int xa = (int)v.x - x0, cnt = (int)v.cnt;
                const uint32_t have = __ballot_sync(0xffffffffu, idx != NO_SEG);
                const uint32_t nfrag = __reduce_add_sync(0xffffffffu, (uint32_t)cnt);
                // (staging replays depths only; aliased segments need the per-fragment colour rule of resolve_deferred)
                const bool can_stage = nfrag <= (uint32_t)ZCAP && !__any_sync(0xffffffffu, cnt > 0 && (v.flags & REC_WRAP));
                if (have && can_stage) stage_chunk(sm, v, idx, xa, cnt);
                if (lane == 0) s_flag[warp] = have ? (can_stage ? 1u : 2u) : 0u;
                if (have && !can_stage) sm.segidx[lane] = idx;
                __syncthreads();
                if (warp == 0) {
                    for (int c = 0; c < SWARPS; c++) {
                        const uint32_t f = s_flag[c];
                        if (f == 0) break;
                        if (f == 1) {
                            replay_chunk(s_all[c], t);
                        } else {  // rare (long segments): warp 0 walks the chunk itself, in order
                            const uint32_t ic = s_all[c].segidx[lane];
                            SegHead vc;
                            vc.cnt = 0; vc.x = (uint16_t)x0;
                            if (ic != NO_SEG) vc = load_head(wb.segv + ic);
                            const int xc = (int)vc.x - x0;
                            direct_chunk(t, vc, ic, xc, pend_mask(xc, (int)vc.cnt), false);
                        }
                        __syncwarp();
                    }
                }
                __syncthreads();
                if constexpr (EACH) {
                    if (have && can_stage) flush_each(sm, v);
                }
                if (wb.tile_clock && lane == 0 && have) atomicAdd(&s_nseg, (uint32_t)__popc(have));
                const bool more = s_flag[SWARPS - 1] != 0;
                idx = idx_next;
                __syncthreads();  // s_flag is rewritten by the next round
                if (!more) break;
            }
            if (warp == 0) {
                const bool touched = __any_sync(0xffffffffu, my_updated != updated_at_start);
                if (st.write_color)
                    for (int h = 0; h < strip_h; h++) wb.vis_seg[(size_t)hq * tile_w + lane + 32 * h] = t.winseg[lane + 32 * h];
                if (touched && st.write_depth) {
#pragma unroll
                    for (int h = 0; h < strip_h; h++) {
                        const int i = lane + 32 * h;
                        if (i < tw) gdepth[grow + i] = t.depth[i];
                    }
                    if (lane == 0) wb.dirty[strip] = 1;
                }
            }
            if (wb.tile_clock && threadIdx.x == 0) {  // (s_nseg is complete: the last round ended with a barrier)
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                wb.tile_clock[2 * strip] = (unsigned long long)(clock64() - t_begin);
                wb.tile_clock[2 * strip + 1] = ((unsigned long long)smid << 32) | s_nseg;
            }
        }
        __syncthreads();  // s_all[0] goes back to warp 0's own strips
    }
    }

    // ---- the other strips: one warp per strip --------------------------------------------------------------------
    // Dealt out round-robin over all warps of the grid.  The strip loop is software-pipelined so that no dependent
    // load is waited for: while strip t is processed, the list entry of t+3, the first 32 (segment, key) pairs of
    // t+2 and an L2 prefetch of the first segments and the depth row of t+1 are in flight.
    const uint32_t nwarps = gridDim.x * SWARPS, gw = blockIdx.x * SWARPS + (uint32_t)warp;
    // The t-th strip of this warp.  Plain round-robin over the list -- except that a warp whose first strip is a
    // heavy one (list positions < nheavy) sits out the next HEAVY_SKIP rounds: a heavy bin costs several average
    // ones, and with equal counts those warps were the kernel's tail (heaviest strip 21 us + three more strips).
    const bool weighted = light0 == 0 && nheavy > 0 && nheavy < nwarps;
    auto qpos = [&](uint32_t t) -> uint32_t {
        if (!weighted) return light0 + gw + t * nwarps;
        const uint32_t wl = nwarps - nheavy;  // warps without a heavy strip
        if (gw < nheavy) return t == 0 ? gw : nheavy + (HEAVY_SKIP + 1u) * wl + (t - 1u) * nwarps + gw;
        return t <= HEAVY_SKIP ? nheavy + t * wl + (gw - nheavy)
                               : nheavy + (HEAVY_SKIP + 1u) * wl + (t - HEAVY_SKIP - 1u) * nwarps + gw;
    };
    uint32_t t_strip = 0, q = qpos(0);
    // list entries are loaded raw as well (index clamped, validity from the position)
    auto load_entry = [&](uint32_t qq) { return busy_at(wb, nheavy, min(qq, nbusy ? nbusy - 1u : 0u)); };
    uint2 e0 = load_entry(q), e1 = load_entry(qpos(1)), e2 = load_entry(qpos(2));
    // the first 64 pairs of strips t and t+1 (a = positions 0..31, b = 32..63)
    uint2 r0a = load_raw(e0.y + lane), r0b = load_raw(e0.y + 32 + lane);
    uint2 r1a = load_raw(e1.y + lane), r1b = load_raw(e1.y + 32 + lane);
    if constexpr (DEFERRED) gather_issue(sm.recbuf, wb.segv, q < nbusy ? pair_idx(e0.x, r0a) : NO_SEG, lane);
    // the depth row of a strip is loaded one strip ahead as well, straight into registers (a strip belongs to one warp
    // of this launch, so nobody else writes the row in between)
    auto load_depth_row = [&](uint32_t strip_id, double d[STRIP_H]) {
#pragma unroll
        for (int h = 0; h < STRIP_H; h++) d[h] = 0;
        if (strip_id == 0xffffffffu) return;
        const int sx0 = (int)(strip_id % (uint32_t)p.tiles_x) * tile_w;
        const size_t srow = (size_t)(strip_id / (uint32_t)p.tiles_x) * p.width + sx0;
        const int stw = min(tile_w, p.width - sx0);
#pragma unroll
        for (int h = 0; h < STRIP_H; h++)
            if (h < strip_h && lane + 32 * h < stw) d[h] = gdepth[srow + lane + 32 * h];
    };
    double dnext[STRIP_H];
    load_depth_row(q < nbusy ? e0.x : 0xffffffffu, dnext);
    while (q < nbusy) {
        const uint32_t strip = e0.x, bin_beg = e0.y;
        const uint32_t strip1 = qpos(t_strip + 1) < nbusy ? e1.x : 0xffffffffu;  // the warp's next strip
        // pipeline: entry of t+3, pairs of t+2, prefetch of t+1
        const uint2 e3 = load_entry(qpos(t_strip + 3));
        const uint2 r2a = load_raw(e2.y + lane), r2b = load_raw(e2.y + 32 + lane);
        if constexpr (!DEFERRED) {
            const uint32_t ik1 = pair_idx(strip1, r1a);
            if (ik1 != NO_SEG) prefetch_seg(wb.segv + ik1);
        }
        double dcur[STRIP_H];
#pragma unroll
        for (int h = 0; h < STRIP_H; h++) dcur[h] = dnext[h];
        load_depth_row(strip1, dnext);
        const long long t_begin = wb.tile_clock ? clock64() : 0;
        uint32_t nseg = 0;
        const int x0 = (int)(strip % (uint32_t)p.tiles_x) * tile_w;
        const int y = (int)(strip / (uint32_t)p.tiles_x);
        const int tw = min(tile_w, p.width - x0);  // valid columns of this strip
        const size_t grow = (size_t)y * p.width + x0;

        // ---- load the strip --------------------------------------------------------------
        __syncwarp();
#pragma unroll
        for (int h = 0; h < strip_h; h++) {
            const int i = lane + 32 * h;
            if constexpr (DEFERRED) sm.winseg[i] = NO_WINNER;
            if (i < tw) {
                sm.depth[i] = dcur[h];
                if constexpr (!DEFERRED) sm.color[i] = gcolor[grow + i];
            }
        }
        __syncwarp();
        const unsigned long long updated_at_start = my_updated;

        // ---- the bin, 32 segments at a time; the pairs of chunk c+2 are loaded (raw) and the records of chunk c+1
        // copied while chunk c is resolved.  The bin ends where the sorted key changes. -------------
        uint32_t idx = pair_idx(strip, r0a);
        uint2 rb = r0b;
        for (uint32_t chunk = bin_beg; __any_sync(0xffffffffu, idx != NO_SEG); chunk += 32) {
            PHASE_CLOCK(pc0, idx)
            const uint2 rc = load_raw(chunk + 64 + lane);
            const uint32_t idx_b = pair_idx(strip, rb);
            nseg += (uint32_t)__popc(__ballot_sync(0xffffffffu, idx != NO_SEG));
            PHASE_CLOCK(pc0a, idx_b)
#if FGL_STRIP_PHASES
            long long pc0b = 0;
#endif
            std::conditional_t<DEFERRED, SegHead, SegV> v;
            if constexpr (DEFERRED) {
                // the next chunk of this warp's stream: the next 32 segments of the strip, or the first 32 of its
                // next strip; its copy runs while this chunk is resolved
                gather_wait();
                __syncwarp();
                v = gather_take(sm.recbuf, idx, lane, x0);
                __syncwarp();
                PHASE_SET(pc0b, v.cnt)
                const uint32_t nidx = __any_sync(0xffffffffu, idx_b != NO_SEG) ? idx_b : pair_idx(strip1, r1a);
                gather_issue(sm.recbuf, wb.segv, nidx, lane);
            } else {
                v.cnt = 0; v.x = (uint16_t)x0;
                if (idx != NO_SEG) v = load_seg(wb.segv + idx);
                if (idx_b != NO_SEG) prefetch_seg(wb.segv + idx_b);
            }
            const int xa = (int)v.x - x0, cnt = (int)v.cnt;
            PHASE_CLOCK(pc1, __reduce_add_sync(0xffffffffu, (uint32_t)cnt + (uint32_t)__double2loint(v.a01)))
            const unsigned long long pend = pend_mask(xa, cnt);

            // no two segments of the chunk overlap <=> popcounts add up
            const uint32_t orl = __reduce_or_sync(0xffffffffu, (uint32_t)pend);
            const uint32_t orh = __reduce_or_sync(0xffffffffu, (uint32_t)(pend >> 32));
            const uint32_t nfrag = __reduce_add_sync(0xffffffffu, (uint32_t)cnt);
            const bool disjoint = nfrag == (uint32_t)(__popc(orl) + __popc(orh));
            bool staged = false;
            if (FGL_STRIP_PHASES && wb.tile_clock && lane == 0) {  // tuning aid (variant build: same-address atomics distort the timing): chunks / fragments by resolution path
                unsigned long long *dbg = wb.tile_clock + 2 * (size_t)wb.ntiles;
                const int path = disjoint ? 0 : ((DEFERRED && nfrag <= (uint32_t)ZCAP) ? 1 : 2);
                atomicAdd(&dbg[path], 1ull);
                atomicAdd(&dbg[4 + path], (unsigned long long)nfrag);
            }
            if constexpr (DEFERRED) {
                if (!disjoint && nfrag <= (uint32_t)ZCAP && !__any_sync(0xffffffffu, cnt > 0 && (v.flags & REC_WRAP))) {
                    staged = true;
                    PHASE_CLOCK(pc2, nfrag)
                    stage_chunk(sm, v, idx, xa, cnt);
                    __syncwarp();
                    PHASE_CLOCK(pc3, sm.cover[lane])
                    replay_chunk(sm, sm);
                    if constexpr (EACH) {
                        __syncwarp();
                        flush_each(sm, v);
                    }
                    __syncwarp();
                    PHASE_CLOCK(pc4, sm.winseg[lane])
#if FGL_STRIP_PHASES
                    if (wb.tile_clock && lane == 0 && (FGL_STRIP_PHASES == 3 || q - light0 < nheavy)) {
                        unsigned long long *ph = wb.tile_clock + 2 * (size_t)wb.ntiles + 10;
                        atomicAdd(&ph[0], 1ull);
                        atomicAdd(&ph[1], (unsigned long long)(pc0a - pc0));  // (key, index) pairs of chunk + 2
                        atomicAdd(&ph[2], (unsigned long long)(pc0b - pc0a)); // wait for the records, take them
                        atomicAdd(&ph[3], (unsigned long long)(pc1 - pc0b));  // issue the next gather
                        atomicAdd(&ph[4], (unsigned long long)(pc3 - pc1));   // overlap test + staging
                        atomicAdd(&ph[5], (unsigned long long)(pc4 - pc3));   // replay
                    }
#endif
                }
            }
            if (!staged) direct_chunk(sm, v, idx, xa, pend, disjoint);
            idx = idx_b; rb = rc;
        }

        // ---- write the strip back ---------------------------------------------------------------
        const bool touched = __any_sync(0xffffffffu, my_updated != updated_at_start);
        if constexpr (DEFERRED) {
            if (st.write_color) {  // k_shade reads the winners of every busy strip, by LIST position (it does not wait for the entry)
                for (int h = 0; h < strip_h; h++) wb.vis_seg[(size_t)q * tile_w + lane + 32 * h] = sm.winseg[lane + 32 * h];
            }
        } else {
            if (touched && st.write_color) {
#pragma unroll
                for (int h = 0; h < strip_h; h++) {
                    const int i = lane + 32 * h;
                    if (i < tw) gcolor[grow + i] = sm.color[i];
                }
            }
        }
        if (touched && st.write_depth) {
#pragma unroll
            for (int h = 0; h < strip_h; h++) {
                const int i = lane + 32 * h;
                if (i < tw) gdepth[grow + i] = sm.depth[i];
            }
            if (lane == 0) wb.dirty[strip] = 1;
        }
        if (wb.tile_clock && lane == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            wb.tile_clock[2 * strip] = (unsigned long long)(clock64() - t_begin);
            wb.tile_clock[2 * strip + 1] = ((unsigned long long)smid << 32) | nseg;
        }
        q = qpos(++t_strip);
        e0 = e1; e1 = e2; e2 = e3;
        r0a = r1a; r0b = r1b; r1a = r2a; r1b = r2b;
    }

    // UpdatedPixels: one atomic per warp per kernel
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_updated += __shfl_down_sync(0xffffffffu, my_updated, o);
    if (lane == 0 && my_updated) atomicAdd(&wb.counters->updated_pixels, my_updated);
}

// ---- deferred shading of the final winners -------------------------------------------------------------
// One thread per pixel of every busy strip.  The deferred mode only admits SolidColor, or Phong with an
// ObjectColor and no texture (fgl_api.cu), so the attribute set is fixed: 9 normal + 9 position components.
constexpr int SHT = 256;
__global__ void __launch_bounds__(SHT, 4)
k_shade(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb, uint32_t *__restrict__ gcolor,
        DrawCounters *acc) {
    pdl_wait();
    pdl_trigger();
    const uint32_t nheavy = wb.tile_ctl->nheavy;
    // (work buffers too small: nothing to shade, but the overflow flag still has to reach the frame's counters)
    const uint32_t nbusy = wb.counters->overflow ? 0u : nheavy + wb.tile_ctl->nlight;
    const uint32_t tile_w = (uint32_t)p.tile_w, SPB = SHT / tile_w;  // strips per CTA pass
    const int pix = (int)(threadIdx.x % tile_w);
    // list entry -> winner -> segment record -> 18 attributes were four dependent memory round trips per pixel
    // (long-scoreboard stall 9.7 per issue, profiles/README.md).  The winners are stored by LIST position (k_strip
    // knows it), so the entry and the winner are loaded side by side, and both are in flight one strip ahead.
    const uint32_t stride = gridDim.x * SPB;
    uint32_t q = blockIdx.x * SPB + threadIdx.x / tile_w;
    auto strip_at = [&](uint32_t qq) { return qq < nbusy ? busy_at(wb, nheavy, qq).x : 0xffffffffu; };
    auto winner_at = [&](uint32_t qq) { return qq < nbusy ? wb.vis_seg[(size_t)qq * tile_w + pix] : NO_WINNER; };
    uint32_t strip = strip_at(q), sidx = winner_at(q);
    for (; q < nbusy; q += stride) {
        const uint32_t strip1 = strip_at(q + stride), sidx1 = winner_at(q + stride);  // in flight during this iteration
        const uint32_t cur_strip = strip, cur_sidx = sidx;
        strip = strip1; sidx = sidx1;
        if (cur_sidx == NO_WINNER) continue;
        const int x = (int)((cur_strip % (uint32_t)p.tiles_x) * tile_w) + pix;
        const int y = (int)(cur_strip / (uint32_t)p.tiles_x);
        uint32_t *out = gcolor + (size_t)y * p.width + x;
        if (p.kind == FGL_SHADER_SOLID) {  // SolidColorShader.Fragment, shader.go:25-27: nothing to interpolate
            *out = c_nrgba(c4(p.color[0], p.color[1], p.color[2], p.color[3]));
            continue;
        }
        const SegV *sp = wb.segv + cur_sidx;
        const uint32_t src = sp->src, flags = sp->flags;
        double n[3][3], pos[3][3];  // all attribute loads are issued up front
        if (flags & REC_SRC_POOL) {
            const ClipTri *ct = wb.clip_pool + src;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const ClipVertex *cv = &ct->v[(flags >> (2 * k)) & 3u];
#pragma unroll
                for (int c = 0; c < 3; c++) { n[k][c] = cv->nrm[c]; pos[k][c] = cv->pos[c]; }
            }
        } else {
            const size_t N = p.mesh.n;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                const size_t vs = (flags >> (2 * k)) & 3u;
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    n[k][c] = __ldg(p.mesh.nrm + (vs * 3 + c) * N + src);
                    pos[k][c] = __ldg(p.mesh.pos + (vs * 3 + c) * N + src);
                }
            }
        }
        // the winner's edge values: the segment's chain of adds from its first pixel (context.go:211-213)
        double w0 = sp->w0, w1 = sp->w1, w2 = sp->w2;
        {
            const double a12 = sp->a12, a20 = sp->a20, a01 = sp->a01;
            for (int k = x - (int)sp->x; k > 0; k--) { w0 += a12; w1 += a20; w2 += a01; }
        }
        const double ra = sp->ra;
        const double b0 = w0 * ra, b1 = w1 * ra, b2 = w2 * ra;
        const double bx = b0 * sp->r0, by = b1 * sp->r1, bzz = b2 * sp->r2;  // context.go:236
        const double bw = 1 / (bx + by + bzz);
        // PhongShader.Fragment, shader.go:75-96
        const V3 normal = v_normalize(v3(interp1(n[0][0], n[1][0], n[2][0], bx, by, bzz, bw),
                                         interp1(n[0][1], n[1][1], n[2][1], bx, by, bzz, bw),
                                         interp1(n[0][2], n[1][2], n[2][2], bx, by, bzz, bw)));
        const V3 ld = v3(p.light[0], p.light[1], p.light[2]);
        C4 light = c4(p.ambient[0], p.ambient[1], p.ambient[2], p.ambient[3]);
        const double diffuse = go_max(v_dot(normal, ld), 0);
        light = c_add(light, c_muls(c4(p.diffuse[0], p.diffuse[1], p.diffuse[2], p.diffuse[3]), diffuse));
        if (diffuse > 0 && p.specular_power > 0) {
            const V3 position = v3(interp1(pos[0][0], pos[1][0], pos[2][0], bx, by, bzz, bw),
                                   interp1(pos[0][1], pos[1][1], pos[2][1], bx, by, bzz, bw),
                                   interp1(pos[0][2], pos[1][2], pos[2][2], bx, by, bzz, bw));
            const V3 camera = v_normalize(v_sub(v3(p.camera[0], p.camera[1], p.camera[2]), position));
            const V3 reflected = v_reflect(v_negate(ld), normal);
            double specular = go_max(v_dot(camera, reflected), 0);
            if (specular > 0) {
                specular = go_pow(specular, p.specular_power);
                light = c_add(light, c_muls(c4(p.specular[0], p.specular[1], p.specular[2], p.specular[3]), specular));
            }
        }
        const C4 color = c4(p.object[0], p.object[1], p.object[2], p.object[3]);
        C4 r = c_mul(color, light);
        r = c4(go_min(r.r, 1), go_min(r.g, 1), go_min(r.b, 1), color.a);
        *out = c_nrgba(r);  // SetNRGBA, context.go:269 (blending excluded by the mode)
    }
    // Async draws: the last CTA to finish adds the draw's counters to the frame's (this was a kernel of its own).
    if (acc) {
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&wb.counters->shade_done, 1u) == gridDim.x - 1u) {
                accumulate_counters(wb.counters, acc);
                // every CTA is past its last read of the counters: leave them zeroed for the next draw
                unsigned int *w = reinterpret_cast<unsigned int *>(wb.counters);
                for (unsigned k = 0; k < sizeof(DrawCounters) / sizeof(unsigned int); k++) w[k] = 0u;
            }
        }
    }
}

int launch_raster(const DrawParams &p, const WorkBuffers &wb, int sorted_buf, uint32_t *color, double *depth,
                  DrawCounters *acc, bool *accumulated, cudaStream_t st) {
    *accumulated = false;
    auto strip_kernel = p.deferred ? (p.prim_info ? k_strip<true, true> : k_strip<true, false>)
                                   : (p.prim_info ? k_strip<false, true> : k_strip<false, false>);
    const uint32_t per_sm = p.deferred ? (uint32_t)FGL_STRIP_MINB : 3u;
    const uint32_t want = (wb.ntiles + SWARPS - 1u) / SWARPS;  // never more warps than strips
    const uint32_t grid = want < wb.nsm * per_sm ? (want ? want : 1u) : wb.nsm * per_sm;
    const size_t smem = (p.deferred ? sizeof(StripMem<true>) : sizeof(StripMem<false>)) * SWARPS;
    cudaFuncSetAttribute(strip_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device, cheap
    launch_pdl(strip_kernel, grid, STHREADS, smem, st, p, wb, (const uint32_t *)wb.seg_val[sorted_buf],
               (const uint32_t *)wb.seg_key[sorted_buf], color, depth);
    int launches = 1;
    if (p.deferred && p.state.write_color) {
        launch_pdl(k_shade, wb.nsm * 8u, SHT, 0, st, p, wb, color, acc);
        *accumulated = acc != nullptr;
        launches++;
    }
    return launches;
}

}  // namespace fgl
