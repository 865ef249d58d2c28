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

namespace fgl {

constexpr int SWARPS = 8;               // warps (= strips in flight) per CTA
constexpr int STHREADS = SWARPS * 32;
constexpr uint32_t NO_WINNER = 0xffffffffu;
constexpr uint32_t NO_SEG = 0xffffffffu;
static_assert(TILE_W == 64, "a strip is one 64-bit pending mask wide at most");
constexpr int STRIP_H = TILE_W / 32;  // pixels per lane at most (p.tile_w / 32 are used)

constexpr int ZCAP = 256;  // fragments of a chunk the pixel-parallel resolution can stage
template <bool DEFERRED> struct StripMem;
template <> struct StripMem<true> {
    double zbuf[ZCAP];       // depth of every fragment of the chunk, segment-major
    double depth[TILE_W];
    uint32_t winseg[TILE_W]; // segment (index into segv) whose fragment currently owns the pixel
    // pixel-parallel resolution of a chunk whose segments overlap
    uint32_t cover[TILE_W];  // lanes (= segments, in primitive order) that cover the pixel
    uint32_t segidx[32];     // segv index of the lane's segment
    uint32_t upd[32];        // UpdatedPixels per segment (per-primitive RasterizeInfo only)
    uint16_t segbase[32];    // first zbuf slot of the segment
    uint8_t segxa[32];       // its first pixel
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
            if (st.alpha_blend && color.a < 1) sm.color[pi] = blend_over(sm.color[pi], c8);
            else sm.color[pi] = c8;
        }
    }
}

// Deferred mode, one fragment whose depth z is known: context.go:232 early-out, then (no discard possible)
// the retest at :248; the winning segment is remembered for the shading kernel.
FGL_DI void resolve_deferred(const fgl_state &st, StripMem<true> &sm, int pi, double z, uint32_t seg,
                             unsigned long long &updated) {
    const double bz = z + st.depth_bias;
    const double dcur = sm.depth[pi];
    if (!(st.read_depth && bz > dcur) && (bz <= dcur || !st.read_depth)) {
        updated++;
        if (st.write_depth) sm.depth[pi] = z;
        sm.winseg[pi] = seg;
    }
}

FGL_DI uint2 busy_at(const WorkBuffers &wb, uint32_t nheavy, uint32_t q) {
    return q < nheavy ? wb.busy_list[q] : wb.busy_list[wb.ntiles - 1u - (q - nheavy)];
}
FGL_DI void prefetch_l2(const void *ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

// EACH: also attribute UpdatedPixels to the primitive of every segment (fgl_draw_*_each).
template <bool DEFERRED, bool EACH>
__global__ void __launch_bounds__(STHREADS, DEFERRED ? 4 : 3)
k_strip(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb,
        const uint32_t *__restrict__ seg_order, const uint32_t *__restrict__ seg_keys, uint32_t *__restrict__ gcolor,
        double *__restrict__ gdepth) {
    if (wb.counters->overflow) return;  // work buffers too small: the host regrows and re-issues the draw

    __shared__ StripMem<DEFERRED> s_all[SWARPS];
    const int lane = threadIdx.x & 31;
    StripMem<DEFERRED> &sm = s_all[threadIdx.x >> 5];
    const fgl_state st = p.state;
    unsigned long long my_updated = 0;
    const int tile_w = p.tile_w, strip_h = tile_w >> 5;  // strip width of this context (32 or 64), pixels per lane
    if constexpr (DEFERRED) {  // scratch of the pixel-parallel resolution: every use leaves it clean
        for (int h = 0; h < STRIP_H; h++) sm.cover[lane + 32 * h] = 0;
        sm.upd[lane] = 0;
        __syncwarp();
    }

    // Busy strips (heavy ones first in the list, k_tile_ranges) are dealt out round-robin over all warps of the
    // grid.  The strip loop is software-pipelined so that no dependent load is waited for: while strip t is
    // processed, the list entry of t+3, the first 32 (segment, key) pairs of t+2 and an L2 prefetch of the first
    // segments and the depth row of t+1 are in flight.
    const uint32_t nsegs = min(wb.counters->n_segs, wb.cap_segs);
    const uint32_t nheavy = wb.tile_ctl->nheavy, nbusy = nheavy + wb.tile_ctl->nlight;
    const uint32_t nwarps = gridDim.x * SWARPS;
    uint32_t q = blockIdx.x * SWARPS + (threadIdx.x >> 5);
    auto load_entry = [&](uint32_t qq) { return qq < nbusy ? busy_at(wb, nheavy, qq) : make_uint2(0xffffffffu, 0u); };
    // (segment index, valid) of sorted position pos for a bin of `strip`
    auto load_pair = [&](uint32_t strip, uint32_t pos) {
        uint32_t idx = NO_SEG;
        if (strip != 0xffffffffu && pos < nsegs) {  // two independent loads, then the select
            const uint32_t k = seg_keys[pos], o = seg_order[pos];
            if (k == strip) idx = o;
        }
        return idx;
    };
    uint2 e0 = load_entry(q), e1 = load_entry(q + nwarps), e2 = load_entry(q + 2 * nwarps);
    uint32_t ik0 = load_pair(e0.x, e0.y + lane), ik1 = load_pair(e1.x, e1.y + lane);
    while (e0.x != 0xffffffffu) {
        const uint32_t strip = e0.x, bin_beg = e0.y;
        // pipeline: entry of t+3, pairs of t+2, prefetch of t+1
        const uint2 e3 = load_entry(q + 3 * nwarps);
        const uint32_t ik2 = load_pair(e2.x, e2.y + lane);
        if (ik1 != NO_SEG) prefetch_l2(wb.segv + ik1);
        if (e1.x != 0xffffffffu && lane < tile_w / 16)
            prefetch_l2(gdepth + (size_t)(e1.x / (uint32_t)p.tiles_x) * p.width + (e1.x % (uint32_t)p.tiles_x) * tile_w + lane * 16);
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
                sm.depth[i] = gdepth[grow + i];
                if constexpr (!DEFERRED) sm.color[i] = gcolor[grow + i];
            }
        }
        __syncwarp();
        const unsigned long long updated_at_start = my_updated;

        // ---- the bin, 32 segments at a time; the pairs of chunk c+2 and an L2 prefetch of the segments of chunk
        // c+1 are in flight while chunk c is resolved.  The bin ends where the sorted key changes. -------------
        uint32_t idx = ik0, idx_b = load_pair(strip, bin_beg + 32 + lane);
        for (uint32_t chunk = bin_beg; __any_sync(0xffffffffu, idx != NO_SEG); chunk += 32) {
            const uint32_t idx_c = load_pair(strip, chunk + 64 + lane);
            nseg += (uint32_t)__popc(__ballot_sync(0xffffffffu, idx != NO_SEG));
            SegV v;
            v.cnt = 0; v.x = (uint16_t)x0;
            if (idx != NO_SEG) v = wb.segv[idx];
            if (idx_b != NO_SEG) prefetch_l2(wb.segv + idx_b);
            const int xa = (int)v.x - x0, cnt = (int)v.cnt;
            unsigned long long pend = cnt > 0 ? ((cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull)) << xa) : 0ull;
            const unsigned long long updated_before = my_updated;

            // no two segments of the chunk overlap <=> popcounts add up
            const uint32_t orl = __reduce_or_sync(0xffffffffu, (uint32_t)pend);
            const uint32_t orh = __reduce_or_sync(0xffffffffu, (uint32_t)(pend >> 32));
            const uint32_t nfrag = __reduce_add_sync(0xffffffffu, (uint32_t)cnt);
            const bool disjoint = nfrag == (uint32_t)(__popc(orl) + __popc(orh));
            bool staged = false;
            if (wb.tile_clock && lane == 0) {  // tuning aid: chunks / fragments by resolution path
                unsigned long long *dbg = wb.tile_clock + 2 * (size_t)wb.ntiles;
                const int path = disjoint ? 0 : ((DEFERRED && nfrag <= (uint32_t)ZCAP) ? 1 : 2);
                atomicAdd(&dbg[path], 1ull);
                atomicAdd(&dbg[4 + path], (unsigned long long)nfrag);
            }
            if constexpr (DEFERRED) {
                if (!disjoint && nfrag <= (uint32_t)ZCAP) {
                    // (1) every segment lane stages the depths of its fragments and registers itself in the
                    // cover mask of its pixels; (2) every PIXEL lane replays the fragments of its pixel in lane
                    // order == primitive order against a running depth in registers.
                    staged = true;
                    const uint32_t fbase = warp_incl_scan((uint32_t)cnt) - (uint32_t)cnt;
                    sm.segbase[lane] = (uint16_t)fbase;
                    sm.segxa[lane] = (uint8_t)xa;
                    sm.segidx[lane] = idx;
                    {
                        double w0 = v.w0, w1 = v.w1, w2 = v.w2;
                        for (int k = 0; k < cnt; k++) {
                            const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
                            sm.zbuf[fbase + k] = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
                            atomicOr(&sm.cover[xa + k], 1u << lane);
                            w0 += v.a12; w1 += v.a20; w2 += v.a01;  // context.go:211-213
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int h = 0; h < strip_h; h++) {
                        const int pix = lane + 32 * h;
                        uint32_t m = sm.cover[pix];
                        if (m) {
                            sm.cover[pix] = 0;
                            double d = sm.depth[pix];
                            int win = -1;
                            while (m) {
                                const int j = __ffs(m) - 1;
                                m &= m - 1;
                                const double z = sm.zbuf[(int)sm.segbase[j] + pix - (int)sm.segxa[j]];
                                const double bz = z + st.depth_bias;
                                // context.go:232 early-out, then (no discard possible) the retest at :248
                                if (!(st.read_depth && bz > d) && (bz <= d || !st.read_depth)) {
                                    my_updated++;
                                    if (st.write_depth) d = z;
                                    win = j;
                                    if constexpr (EACH) atomicAdd(&sm.upd[j], 1u);
                                }
                            }
                            if (win >= 0) {
                                if (st.write_depth) sm.depth[pix] = d;
                                sm.winseg[pix] = sm.segidx[win];
                            }
                        }
                    }
                    if constexpr (EACH) {
                        __syncwarp();
                        const uint32_t u = sm.upd[lane];
                        sm.upd[lane] = 0;
                        if (u) atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, v.src, v.flags) + 1], (unsigned long long)u);
                    }
                    __syncwarp();
                }
            }
            if (!staged) {
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
                                    resolve_deferred(st, sm, pi, z, idx, my_updated);
                                } else {
                                    fragment_inline(p, st, wb, v, w0, w1, w2, pi, sm, my_updated);
                                }
                            }
                            w0 += v.a12; w1 += v.a20; w2 += v.a01;  // context.go:211-213
                        }
                        pend &= ~ready;
                    }
                    __syncwarp();
                    if (wb.tile_clock && lane == 0 && !disjoint) atomicAdd(&wb.tile_clock[2 * (size_t)wb.ntiles + 8], 1ull);
                    if (disjoint || !__any_sync(0xffffffffu, pend != 0)) break;
                }
                if (EACH && my_updated != updated_before)
                    atomicAdd(&p.prim_info[2 * (size_t)src_primitive(wb, p, v.src, v.flags) + 1], my_updated - updated_before);
            }
            idx = idx_b; idx_b = idx_c;
        }

        // ---- write the strip back ---------------------------------------------------------------
        const bool touched = __any_sync(0xffffffffu, my_updated != updated_at_start);
        if constexpr (DEFERRED) {
            if (st.write_color) {  // k_shade reads the winners of every busy strip
                for (int h = 0; h < strip_h; h++) wb.vis_seg[(size_t)strip * tile_w + lane + 32 * h] = sm.winseg[lane + 32 * h];
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
        }
        if (wb.tile_clock && lane == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            wb.tile_clock[2 * strip] = (unsigned long long)(clock64() - t_begin);
            wb.tile_clock[2 * strip + 1] = ((unsigned long long)smid << 32) | nseg;
        }
        q += nwarps;
        e0 = e1; e1 = e2; e2 = e3;
        ik0 = ik1; ik1 = ik2;
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
k_shade(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb, uint32_t *__restrict__ gcolor) {
    if (wb.counters->overflow) return;
    const uint32_t nheavy = wb.tile_ctl->nheavy, nbusy = nheavy + wb.tile_ctl->nlight;
    const uint32_t tile_w = (uint32_t)p.tile_w, SPB = SHT / tile_w;  // strips per CTA pass
    const int pix = (int)(threadIdx.x % tile_w);
    for (uint32_t q = blockIdx.x * SPB + threadIdx.x / tile_w; q < nbusy; q += gridDim.x * SPB) {
        const uint32_t strip = busy_at(wb, nheavy, q).x;
        const uint32_t sidx = wb.vis_seg[(size_t)strip * tile_w + pix];
        if (sidx == NO_WINNER) continue;
        const int x = (int)((strip % (uint32_t)p.tiles_x) * tile_w) + pix;
        const int y = (int)(strip / (uint32_t)p.tiles_x);
        uint32_t *out = gcolor + (size_t)y * p.width + x;
        if (p.kind == FGL_SHADER_SOLID) {  // SolidColorShader.Fragment, shader.go:25-27: nothing to interpolate
            *out = c_nrgba(c4(p.color[0], p.color[1], p.color[2], p.color[3]));
            continue;
        }
        const SegV *sp = wb.segv + sidx;
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
}

int launch_raster(const DrawParams &p, const WorkBuffers &wb, int sorted_buf, uint32_t *color, double *depth,
                  cudaStream_t st) {
    auto strip_kernel = p.deferred ? (p.prim_info ? k_strip<true, true> : k_strip<true, false>)
                                   : (p.prim_info ? k_strip<false, true> : k_strip<false, false>);
    const uint32_t per_sm = p.deferred ? 4u : 3u;
    const uint32_t want = (wb.ntiles + SWARPS - 1u) / SWARPS;  // never more warps than strips
    const uint32_t grid = want < wb.nsm * per_sm ? (want ? want : 1u) : wb.nsm * per_sm;
    strip_kernel<<<grid, STHREADS, 0, st>>>(p, wb, wb.seg_val[sorted_buf], wb.seg_key[sorted_buf], color, depth);
    int launches = 1;
    if (p.deferred && p.state.write_color) {
        k_shade<<<wb.nsm * 8u, SHT, 0, st>>>(p, wb, color);
        launches++;
    }
    return launches;
}

}  // namespace fgl
