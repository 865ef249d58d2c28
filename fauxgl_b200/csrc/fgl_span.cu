// fgl_span.cu -- span stage: one thread per (triangle, scanline) replays the
// reference's inner loop for that row (context.go:184-221): the per-row adds from
// y0, the skip-ahead `d`, then the per-pixel forward-differencing adds, exactly as
// written, once, left to right.  The covered run is cut at tile-column
// boundaries into segments that carry the edge values at their first pixel, so
// the tile kernel continues the same chain of adds and reproduces coverage,
// barycentrics and depth bit for bit -- without ever restarting a row.
//
// This stage is unordered and load-balanced (grid-wide, no ordering barriers);
// only covered segments reach the ordered, tile-serial back end.  Segments are
// written in (triangle, scanline, column) order and then stably sorted by tile,
// which keeps primitive order inside every bin.
//
// Two passes around one scan (ordered compaction): the walk pass counts the
// segments of every scanline and parks the first one (almost every scanline of a
// small triangle has exactly one); the place pass copies parked segments to their
// ordered slots and re-walks only the scanlines that cross a tile boundary.
#include "fgl_internal.h"
#include "fgl_block.cuh"
#include "fgl_math.cuh"
#include "fgl_walk.cuh"

namespace fgl {

constexpr int ST = 256;            // threads per CTA == (record, scanline) items per virtual block
constexpr int SPAN_GRID = (int)GRID_WAVE;  // persistent grid: virtual blocks are strided over it

// The fields of Rec the row walker needs.
struct RowSetup {
    double s0x, s0y, s1x, s1y, s2x, s2y;
    double w00, w01, w02, ra, ra12, ra20, ra01;
    double z0, z1, z2;
    int32_t x0, x1, y0;
};
__device__ __forceinline__ RowSetup load_setup(const Rec *rp) {
    RowSetup r;
    r.s0x = rp->s[0]; r.s0y = rp->s[1]; r.s1x = rp->s[3]; r.s1y = rp->s[4]; r.s2x = rp->s[6]; r.s2y = rp->s[7];
    r.w00 = rp->w00; r.w01 = rp->w01; r.w02 = rp->w02;
    r.ra = rp->ra; r.ra12 = rp->ra12; r.ra20 = rp->ra20; r.ra01 = rp->ra01;
    r.z0 = rp->s[2]; r.z1 = rp->s[5]; r.z2 = rp->s[8];
    r.x0 = rp->x0; r.x1 = rp->x1; r.y0 = rp->y0;
    return r;
}

__device__ __forceinline__ RecTail load_tail(const Rec *rp) {
    RecTail t;
    t.r0 = rp->r0; t.r1 = rp->r1; t.r2 = rp->r2; t.src = rp->src; t.flags = rp->flags;
    return t;
}

// Largest r in [0,n) with off[r] <= i, found by one warp with a 32-ary search (5 rounds for 2^24).
__device__ __forceinline__ uint32_t warp_find(const uint32_t *__restrict__ off, uint32_t n, uint32_t i) {
    const uint32_t lane = threadIdx.x & 31;
    uint32_t lo = 0, hi = n;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const uint32_t step = (hi - lo + 31) / 32;
        const uint32_t idx = lo + (lane + 1) * step;
        const bool le = idx < hi && off[idx] <= i;
        const uint32_t cnt = __popc(__ballot_sync(0xffffffffu, le));  // monotone: the first cnt lanes are true
        const uint32_t nlo = lo + cnt * step;
        hi = min(hi, nlo + step);
        lo = nlo;
    }
    return lo;
}

// Pass 1: walk every (record, scanline); row_nseg, row_first/row_key; TotalPixels.
__global__ void __launch_bounds__(ST)
k_span_walk(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb) {
    const DrawCounters *ctr = wb.counters;
    if (ctr->overflow) return;
    const uint32_t nrec = min(ctr->n_records, wb.cap_records);
    const uint32_t nrows = min(ctr->n_rows, wb.cap_rows);
    __shared__ uint32_t s_first;
    __shared__ uint32_t s_off[ST + 1];
    __shared__ uint32_t s_slot[ST + 1];
    unsigned long long covered = 0;
    for (uint32_t i0 = blockIdx.x * ST; i0 < nrows; i0 += gridDim.x * ST) {
        // The ST items of this virtual block belong to at most ST consecutive records (every record has
        // >= 1 scanline): find the first with one warp, stage that window of the offset array in shared
        // memory and let every thread finish its search there.
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t r0 = warp_find(wb.rec_row_off, nrec, i0);
            if (threadIdx.x == 0) s_first = r0;
        }
        __syncthreads();
        const uint32_t r0 = s_first;
        for (uint32_t k = threadIdx.x; k <= ST; k += ST) {
            const uint32_t r = r0 + k;
            s_off[k] = r <= nrec ? wb.rec_row_off[r] : 0xffffffffu;
            s_slot[k] = r < nrec ? wb.rec_slot[r] : 0u;
        }
        __syncthreads();
        const uint32_t i = i0 + threadIdx.x;
        if (i < nrows) {
            uint32_t lo = 0, hi = ST;  // s_off[lo] <= i < s_off[hi]  (s_off[ST] >= i0 + ST > i)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (s_off[mid] <= i) lo = mid; else hi = mid;
            }
            const uint32_t rid = s_slot[lo];  // records live at their block's slots, not in primitive order
            const RowSetup r = load_setup(wb.recs + rid);
            const int y = row_base(p, r.x1, r.y0) + (int)(i - s_off[lo]);
            const unsigned long long before = covered;
            ParkedSeg f;
            const uint32_t nseg = walk_row_segments<false>(p, r, y, f, nullptr, nullptr, nullptr, 0, 0, &covered);
            wb.row_nseg[i] = nseg;
            if (nseg) {  // park the first segment (almost every scanline of a small triangle has exactly one)
                Seg sg;
                sg.w0 = f.w0; sg.w1 = f.w1; sg.w2 = f.w2; sg.rec = rid; sg.x = (uint16_t)f.x; sg.cnt = (uint8_t)f.cnt;
                sg.yt = f.wrap ? 1 : 0;
                wb.row_first[i] = sg;
                // one segment: its strip; several: the scanline itself, which the place pass walks again (a row
                // that aliases into its neighbours does not tell its own y through its first strip)
                wb.row_key[i] = nseg == 1 ? f.key : (uint32_t)y;
            }
            if (p.prim_info && covered != before)  // per-primitive TotalPixels (fgl_draw_*_each)
                atomicAdd(&p.prim_info[2 * (size_t)rec_primitive(wb, p, rid)], covered - before);
        }
    }
    // TotalPixels, context.go:229: every covered in-range pixel, before any depth test
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) covered += __shfl_down_sync(0xffffffffu, covered, o);
    if ((threadIdx.x & 31) == 0 && covered) atomicAdd(&wb.counters->total_pixels, covered);
}

// Pass 2: place the segments at their ordered slots.
__global__ void __launch_bounds__(ST)
k_span_place(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb) {
    const DrawCounters *ctr = wb.counters;
    if (ctr->overflow) return;
    const uint32_t nrows = min(ctr->n_rows, wb.cap_rows);
    unsigned long long dummy = 0;
    for (uint32_t i = blockIdx.x * ST + threadIdx.x; i < nrows; i += gridDim.x * ST) {
        const uint32_t n = wb.row_nseg[i];
        if (n == 0) continue;
        const uint32_t base = wb.row_seg_off[i];
        const Seg s = wb.row_first[i];
        const uint32_t key = wb.row_key[i];
        if (n == 1) {
            if (base < wb.cap_segs) {  // records are visited in order here: the Rec reads are near-sequential
                const Rec *rp = wb.recs + s.rec;
                wb.segv[base] = make_segv(s.w0, s.w1, s.w2, rp->ra, rp->s[2], rp->s[5], rp->s[8], rp->s[7] - rp->s[4],
                                          rp->s[1] - rp->s[7], rp->s[4] - rp->s[1], load_tail(rp), s.x, s.cnt,
                                          s.yt ? REC_WRAP : 0u);
                wb.seg_key[0][base] = key;
                wb.seg_val[0][base] = base;
            }
        } else {  // the scanline crosses tile columns: walk it again, writing every segment
            const RowSetup r = load_setup(wb.recs + s.rec);
            const RecTail tail = load_tail(wb.recs + s.rec);
            const int y = (int)key;
            ParkedSeg f;
            walk_row_segments<true>(p, r, y, f, &tail, wb.segv, wb.seg_key[0], base, wb.cap_segs, &dummy);
            for (uint32_t k = 0; k < n && base + k < wb.cap_segs; k++) wb.seg_val[0][base + k] = base + k;
        }
    }
}

// The list of busy strips: (strip, first position of its bin in the sorted array); a bin ends where the sorted
// key changes.  Heavy strips (>= HEAVY_SEGS segments) are appended from the front, the others from the back,
// so that the strip kernel starts the long ones first.
constexpr int TR_THREADS = 1024;
// Each CTA owns a contiguous part of the sorted array and goes over it twice: first it only counts the strips
// that start there (heavy / light), reserves its share of the busy list with ONE global atomic per list, then
// it fills in the entries (ranks inside the CTA from shared-memory counters).  A thread takes four consecutive
// keys per step with one 16-byte load (plus the key before them), so a bin starts where a key differs from its
// predecessor; only those few positions look 95 keys ahead for the heavy test.
__global__ void __launch_bounds__(TR_THREADS)
k_tile_ranges(const uint32_t *__restrict__ keys, DrawCounters *__restrict__ ctr, uint32_t n_max,
              uint2 *__restrict__ busy_list, uint32_t ntiles, TileCtl *ctl, unsigned long long *__restrict__ group_sums,
              uint32_t ngroups) {
    __shared__ uint32_t s_cnt[2], s_base[2], s_fill[2];
    pdl_wait();
    pdl_trigger();
    // the group sums of the fused front end (k_front adds to them, k_seg_index has read them) are left zeroed for
    // the next draw -- also when this one overflowed and is going to be re-issued
    if (blockIdx.x == 0)
        for (uint32_t g = threadIdx.x; g < ngroups; g += TR_THREADS) group_sums[g] = 0ull;
    if (ctr->overflow) return;  // the draw is going to be re-issued: its keys are incomplete
    const uint32_t n = min(ctr->n_segs, n_max);
    const uint32_t lane = threadIdx.x & 31;
    constexpr uint32_t STEP = TR_THREADS * 4u;
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + STEP - 1) / STEP * STEP;
    const uint64_t b64 = (uint64_t)blockIdx.x * per;
    const uint32_t beg = b64 < n ? (uint32_t)b64 : n, end = b64 + per < n ? (uint32_t)(b64 + per) : n;
    if (threadIdx.x < 2) { s_cnt[threadIdx.x] = 0; s_fill[threadIdx.x] = 0; }
    __syncthreads();
    // the four keys at i .. i+3 (i a multiple of 4) and the one before them; returns the mask of positions where a
    // bin starts and, of those, the heavy ones
    auto classify4 = [&](uint32_t i, uint32_t k[4], uint32_t &heavy) {
        uint32_t starts = 0;
        heavy = 0;
        if (i >= end) return starts;
        if (i + 3u < n) {
            const uint4 v = *reinterpret_cast<const uint4 *>(keys + i);
            k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
        } else {
#pragma unroll
            for (int u = 0; u < 4; u++) k[u] = i + u < n ? keys[i + u] : 0xffffffffu;
        }
        uint32_t prev = i ? keys[i - 1u] : ~k[0];
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            if (i + u < end && k[u] != prev) {
                starts |= 1u << u;
                if (i + u + HEAVY_SEGS - 1u < n && keys[i + u + HEAVY_SEGS - 1u] == k[u]) heavy |= 1u << u;
            }
            prev = k[u];
        }
        return starts;
    };
    uint32_t nh = 0, nl = 0;
    for (uint32_t i = beg + threadIdx.x * 4u; i < end; i += STEP) {
        uint32_t k[4], heavy;
        const uint32_t starts = classify4(i, k, heavy);
        nh += (uint32_t)__popc(heavy);
        nl += (uint32_t)__popc(starts & ~heavy);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { nh += __shfl_down_sync(0xffffffffu, nh, o); nl += __shfl_down_sync(0xffffffffu, nl, o); }
    if (lane == 0) { if (nh) atomicAdd(&s_cnt[0], nh); if (nl) atomicAdd(&s_cnt[1], nl); }
    __syncthreads();
    if (threadIdx.x == 0) {
        s_base[0] = s_cnt[0] ? atomicAdd(&ctl->nheavy, s_cnt[0]) : 0u;
        s_base[1] = s_cnt[1] ? atomicAdd(&ctl->nlight, s_cnt[1]) : 0u;
    }
    __syncthreads();
    for (uint32_t i0 = beg; i0 < end; i0 += STEP) {  // (all threads of a warp iterate together: shuffles below)
        const uint32_t i = i0 + threadIdx.x * 4u;
        uint32_t k[4], heavy;
        const uint32_t starts = classify4(i, k, heavy);
        const uint32_t ch = (uint32_t)__popc(heavy), cl = (uint32_t)__popc(starts & ~heavy);
        // ranks inside the warp (inclusive scans), the warp's base from the CTA counters
        uint32_t ih = ch, il = cl;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t th = __shfl_up_sync(0xffffffffu, ih, o), tl = __shfl_up_sync(0xffffffffu, il, o);
            if ((int)lane >= o) { ih += th; il += tl; }
        }
        const uint32_t toth = __shfl_sync(0xffffffffu, ih, 31), totl = __shfl_sync(0xffffffffu, il, 31);
        uint32_t wh = 0, wl = 0;
        if (lane == 0) {
            if (toth) wh = atomicAdd(&s_fill[0], toth);
            if (totl) wl = atomicAdd(&s_fill[1], totl);
        }
        wh = __shfl_sync(0xffffffffu, wh, 0);
        wl = __shfl_sync(0xffffffffu, wl, 0);
        uint32_t ph = s_base[0] + wh + ih - ch, pl = s_base[1] + wl + il - cl;
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            if (!((starts >> u) & 1u)) continue;
            // (a sorted key array has at most ntiles bins; the bound keeps a corrupted one from writing outside the list)
            if ((heavy >> u) & 1u) { if (ph < ntiles) busy_list[ph] = make_uint2(k[u], i + u); ph++; }
            else { if (pl < ntiles) busy_list[ntiles - 1u - pl] = make_uint2(k[u], i + u); pl++; }
        }
    }
}

int launch_spans(const DrawParams &p, const WorkBuffers &wb, int *sorted_buf, cudaStream_t st) {
    (void)sorted_buf;
    int launches = 0;
    DrawCounters *c = wb.counters;
    const uint32_t vblocks = (wb.cap_rows + ST - 1) / ST;
    const uint32_t grid = vblocks < (uint32_t)SPAN_GRID ? (vblocks ? vblocks : 1) : (uint32_t)SPAN_GRID;
    k_span_walk<<<grid, ST, 0, st>>>(p, wb);
    launches++;
    // (this scan's spine also finalises the clip-pool counters)
    ScanSink segs{&c->n_segs, &c->need_segs, &c->overflow, wb.cap_segs, OVF_SEGS, &c->n_clip, &c->need_clip, wb.cap_clip};
    launches += launch_exclusive_scan(wb.row_nseg, wb.row_seg_off, wb.cap_rows, &c->n_rows, wb.scan_tmp, segs, st);
    k_span_place<<<grid, ST, 0, st>>>(p, wb);
    launches++;
    return launches;
}

int launch_bin(const DrawParams &p, const WorkBuffers &wb, int *sorted_buf, cudaStream_t st) {
    (void)p;
    int launches = 0;
    DrawCounters *c = wb.counters;
    // stable sort of the segment indices by tile id
    int bits = 1;
    while ((1u << bits) < wb.ntiles) bits++;
    if (g_bin_buckets && bin_buckets_ok(bits))
        return launch_bin_buckets(wb.seg_key, wb.seg_val, c, wb.cap_segs, bits, wb.scan_tmp, sorted_buf, wb.busy_list, wb.ntiles,
                                  wb.tile_ctl, wb.blk_base, wb.cap_prims / (128u * 64u) + 2u, st);
    launches += launch_sort_pairs(wb.seg_key, wb.seg_val, &c->n_segs, wb.cap_segs, bits, wb.scan_tmp, sorted_buf, st, &c->overflow);
    launch_pdl(k_tile_ranges, B200_SMS, TR_THREADS, 0, st, (const uint32_t *)wb.seg_key[*sorted_buf], c, wb.cap_segs, wb.busy_list,
               wb.ntiles, wb.tile_ctl, wb.blk_base, wb.cap_prims / (128u * 64u) + 2u);
    launches++;
    return launches;
}

}  // namespace fgl
