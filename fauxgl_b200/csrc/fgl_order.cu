// fgl_order.cu -- the order plumbing between the front end and the strip kernel as ONE cooperative kernel.
//
// Between k_front (segments in block regions) and k_strip (bins of segments per strip, in primitive order) a draw
// used to launch six small kernels: k_seg_index, 2-3 x (k_radix_hist, k_radix_scatter), k_tile_ranges -- 53 us for the
// 458 k segments of the 1080p benchmark frame, i.e. 3.7 MB of pairs: launch ramps, drains and one dependent L2 round
// trip after the other, not work.  k_order runs the same steps as phases of one launch with one CTA per SM, separated
// by grid-wide barriers (a self-resetting arrive / generation pair in TileCtl; the kernel is launched with the
// cooperative attribute, which guarantees that all its CTAs are resident -- also next to the kernels of other
// streams, tools/coop_probe.cu):
//   A  (fused front end only) ordered position of every (block, warp run) of k_front and the primitive-ordered
//      (strip, segment) list -- k_seg_index;
//   B  per 8-bit digit: per-CTA digit histogram of its contiguous input range | barrier | bucket bases from the
//      [CTA][digit] table, stable ranking and scatter (the single tile of a small sort stays in registers between the
//      two halves) | barrier -- k_radix_hist / k_radix_scatter;
//   C  the busy-strip list -- k_tile_ranges.
// The arithmetic and the results are those of the separate kernels (fgl_geom.cu, fgl_scan_sort.cu, fgl_span.cu),
// which remain the path of FGL_ORDER=split and of devices without cooperative launch.
//
// Data written by one CTA and read by another inside the launch goes through L2: plain stores, __ldcg loads (L1 is
// not coherent across SMs), a __threadfence on both sides of every barrier.
#include "fgl_internal.h"
#include "fgl_block.cuh"

namespace fgl {

constexpr int OT = 1024;  // threads per CTA, one CTA per SM
constexpr int OWARPS = OT / 32;
constexpr int OBITS = 8, OBINS = 1 << OBITS;
constexpr int OITEMS = 4;
constexpr int OTILE = OT * OITEMS;

// Grid-wide barrier.  `count` collects the arrivals of one round, the last arriver resets it and bumps `gen`, the
// others wait for the bump.  Reading gen before arriving is safe: it cannot advance until this CTA has arrived.
__device__ __forceinline__ void grid_barrier(TileCtl *ctl) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&ctl->bar_gen) : "memory");
        __threadfence();
        if (atomicAdd(&ctl->bar_count, 1u) == gridDim.x - 1u) {
            ctl->bar_count = 0u;
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(&ctl->bar_gen), "r"(gen + 1u) : "memory");
        } else {
            unsigned int now;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(now) : "l"(&ctl->bar_gen) : "memory");
            } while (now == gen);
        }
        __threadfence();
    }
    __syncthreads();
}

// contiguous input range of this CTA for the radix passes: a multiple of the tile, like radix_segment
__device__ __forceinline__ void order_range(uint32_t n, uint32_t &beg, uint32_t &end) {
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + OTILE - 1) / OTILE * OTILE;
    const uint64_t b = (uint64_t)blockIdx.x * per;
    beg = b < n ? (uint32_t)b : n;
    const uint64_t e = b + per;
    end = e < n ? (uint32_t)e : n;
}

__global__ void __launch_bounds__(OT, 1)
k_order(const __grid_constant__ WorkBuffers wb, uint32_t nent, int bits, uint32_t ngroups_alloc) {
    extern __shared__ uint32_t s_dyn[];
    uint32_t *cursor = s_dyn;                                                       // [OBINS]
    uint32_t(*warp_cnt)[OBINS] = reinterpret_cast<uint32_t(*)[OBINS]>(s_dyn + OBINS);  // [OWARPS][OBINS]
    __shared__ uint32_t s_scan[OWARPS + 1];
    __shared__ unsigned long long s_part[OWARPS];
    __shared__ uint32_t s_cnt[2], s_base[2], s_fill[2];
    pdl_wait();
    DrawCounters *ctr = wb.counters;
    TileCtl *ctl = wb.tile_ctl;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t G = gridDim.x;

    // ---------------- A: primitive-ordered (strip, segment) list of the fused front end ----------------
    if (nent) {
        const uint32_t ngroups = (nent + FRONT_GROUP - 1) / FRONT_GROUP;
        const unsigned long long cells_total = ctr->seg_cursor;  // every block has made its reservation
        const unsigned nclip = ctr->n_clip;
        unsigned ovf = ctr->overflow;  // (k_front sets OVF_CLIP itself when the pool runs out)
        if (cells_total > (unsigned long long)wb.cap_segs) ovf |= OVF_SEGS;
        if (nclip > wb.cap_clip) ovf |= OVF_CLIP;
        if (blockIdx.x == 0) {
            unsigned long long tot = 0;
            for (uint32_t g = tid; g < ngroups; g += OT) tot += wb.blk_base[g];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) tot += __shfl_down_sync(0xffffffffu, tot, o);
            if (lane == 0) s_part[warp] = tot;
            __syncthreads();
            if (tid == 0) {
                tot = 0;
                for (int w = 0; w < OWARPS; w++) tot += s_part[w];
                ctl->nheavy = 0; ctl->nlight = 0; ctl->head = 0;  // the busy-strip list of this draw
                ctr->need_records = 0; ctr->n_rows = 0; ctr->need_rows = 0;
                ctr->n_segs = (unsigned int)min(tot, 0xfffffff0ull);
                ctr->need_segs = (unsigned int)min(cells_total, 0xfffffff0ull);
                ctr->need_clip = nclip;
                if (ovf) atomicOr(&ctr->overflow, ovf);
            }
            __syncthreads();
            // an overflowed draw is re-issued: leave the group sums zeroed for it (nobody else reads them now)
            if (ovf)
                for (uint32_t g = tid; g < ngroups_alloc; g += OT) wb.blk_base[g] = 0ull;
        }
        if (ovf) return;  // (the same value in every CTA: no barrier is left waiting)
        // this CTA's contiguous range of front-end blocks, OT blocks per step
        const uint32_t per = (nent + G - 1) / G;
        const uint32_t b0 = min(blockIdx.x * per, nent), b1 = min(b0 + per, nent);
        if (b0 < b1) {
            const uint32_t g0 = b0 / FRONT_GROUP;
            unsigned long long before = 0;  // segments before block b0: group sums + the blocks of its own group
            for (uint32_t g = tid; g < g0; g += OT) before += wb.blk_base[g];
            for (uint32_t b = g0 * FRONT_GROUP + tid; b < b0; b += OT) {
                const uint4 wc = wb.blk_wcnt[b];
                before += (unsigned long long)wc.x + wc.y + wc.z + wc.w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) before += __shfl_down_sync(0xffffffffu, before, o);
            if (lane == 0) s_part[warp] = before;
            __syncthreads();
            unsigned long long carry = 0;
            for (int w = 0; w < OWARPS; w++) carry += s_part[w];
            __syncthreads();
            uint32_t *s_first = s_dyn;  // [OT] ordered position of the first segment of each block of the step
            const uint32_t n = wb.cap_segs;  // (without overflow every position is below the capacity)
            for (uint32_t c0 = b0; c0 < b1; c0 += OT) {
                const uint32_t nb = min((uint32_t)OT, b1 - c0);
                uint32_t mine = 0;
                if (tid < nb) {
                    const uint4 wc = wb.blk_wcnt[c0 + tid];
                    mine = wc.x + wc.y + wc.z + wc.w;
                }
                uint32_t total;
                const uint32_t ex = block_excl_scan<OT>(mine, s_scan, &total);
                if (tid < nb) s_first[tid] = (uint32_t)min(carry + ex, 0xffffffffull);
                __syncthreads();
                const uint32_t ntasks = nb * 4u;
                for (uint32_t t = warp; t < ntasks; t += OWARPS) {  // one warp per (block, warp run), four loads in flight
                    const uint32_t k_blk = t >> 2, b = c0 + k_blk, w = t & 3u;
                    const uint4 wc = wb.blk_wcnt[b];
                    const uint32_t cnt = w == 0 ? wc.x : (w == 1 ? wc.y : (w == 2 ? wc.z : wc.w));
                    if (cnt == 0) continue;
                    const uint4 wo = wb.blk_woff[b];
                    const uint32_t before_w = w == 0 ? 0u : (w == 1 ? wc.x : (w == 2 ? wc.x + wc.y : wc.x + wc.y + wc.z));
                    const uint32_t pos0 = s_first[k_blk] + before_w;
                    const uint32_t slot0 = wb.blk_region[b] + (w == 0 ? wo.x : (w == 1 ? wo.y : (w == 2 ? wo.z : wo.w)));
                    for (uint32_t k0 = 0; k0 < cnt; k0 += 128u) {
                        uint32_t key[4];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t k = k0 + 32u * u + lane;
                            key[u] = (k < cnt && pos0 + k < n) ? wb.seg_key[1][slot0 + k] : 0u;
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const uint32_t k = k0 + 32u * u + lane;
                            if (k < cnt && pos0 + k < n) {
                                wb.seg_key[0][pos0 + k] = key[u];
                                wb.seg_val[0][pos0 + k] = slot0 + k;
                            }
                        }
                    }
                }
                carry += total;
                __syncthreads();
            }
        }
        grid_barrier(ctl);
        // the group sums have been read by everybody: left zeroed for the next draw
        if (blockIdx.x == 0)
            for (uint32_t g = tid; g < ngroups_alloc; g += OT) wb.blk_base[g] = 0ull;
    } else {
        // split front end: its own kernels have listed the segments; an overflowed draw has no complete list
        if (blockIdx.x == 0)
            for (uint32_t g = tid; g < ngroups_alloc; g += OT) wb.blk_base[g] = 0ull;
        if (ctr->overflow) return;
    }

    // ---------------- B: stable LSD radix sort of the pairs by strip id ----------------
    const uint32_t n = min(__ldcg(&ctr->n_segs), wb.cap_segs);
    uint32_t beg, end;
    order_range(n, beg, end);
    uint32_t *hist = wb.scan_tmp;  // [G][OBINS]
    const uint32_t ltmask = (1u << lane) - 1u;
    int cur = 0;
    for (int shift = 0; shift < bits; shift += OBITS, cur ^= 1) {
        const uint32_t *keys_in = wb.seg_key[cur], *vals_in = wb.seg_val[cur];
        uint32_t *keys_out = wb.seg_key[cur ^ 1], *vals_out = wb.seg_val[cur ^ 1];
        // -- histogram of this CTA's range; its first tile stays in registers
        for (int k = tid; k < OBINS; k += OT) cursor[k] = 0;
        __syncthreads();
        uint32_t key0[OITEMS], val0[OITEMS];
        {
            const uint32_t wbase = beg + warp * (32 * OITEMS);
#pragma unroll
            for (int r = 0; r < OITEMS; r++) {
                const uint32_t i = wbase + r * 32 + lane;
                key0[r] = i < end ? __ldcg(&keys_in[i]) : 0u;
                val0[r] = i < end ? __ldcg(&vals_in[i]) : 0u;
            }
#pragma unroll
            for (int r = 0; r < OITEMS; r++) {
                const uint32_t i = wbase + r * 32 + lane;
                if (i < end) atomicAdd(&cursor[(key0[r] >> shift) & (OBINS - 1)], 1u);
            }
        }
        for (uint32_t i = beg + OTILE + tid; i < end; i += OT) atomicAdd(&cursor[(__ldcg(&keys_in[i]) >> shift) & (OBINS - 1)], 1u);
        __syncthreads();
        for (int k = tid; k < OBINS; k += OT) hist[blockIdx.x * OBINS + k] = cursor[k];
        grid_barrier(ctl);
        // -- bucket bases: digit totals over all CTAs, and this digit's count in the CTAs before this one
        {
            constexpr uint32_t Q = OT / OBINS;  // 4 threads per digit
            const uint32_t d = tid & (OBINS - 1), q = tid / OBINS;
            uint32_t row = 0, before = 0;
            constexpr uint32_t BATCH = 10;
            for (uint32_t c0 = q; c0 < G; c0 += Q * BATCH) {
                uint32_t v[BATCH];
#pragma unroll
                for (uint32_t k = 0; k < BATCH; k++) {
                    const uint32_t c = c0 + k * Q;
                    v[k] = c < G ? __ldcg(&hist[c * OBINS + d]) : 0u;
                }
#pragma unroll
                for (uint32_t k = 0; k < BATCH; k++) {
                    row += v[k];
                    if (c0 + k * Q < blockIdx.x) before += v[k];
                }
            }
            warp_cnt[q][d] = row;
            warp_cnt[Q + q][d] = before;
            __syncthreads();
            row = before = 0;
            if (tid < OBINS)
                for (uint32_t k = 0; k < Q; k++) { row += warp_cnt[k][d]; before += warp_cnt[Q + k][d]; }
            uint32_t total;
            const uint32_t digit_base = block_excl_scan<OT>(tid < OBINS ? row : 0u, s_scan, &total);
            if (tid < OBINS) cursor[d] = digit_base + before;
        }
        __syncthreads();
        // -- ranking and scatter, tile by tile (k_radix_scatter)
        for (uint32_t base = beg; base < end; base += OTILE) {
            for (int k = tid; k < OWARPS * OBINS; k += OT) (&warp_cnt[0][0])[k] = 0;
            __syncthreads();
            uint32_t key[OITEMS], val[OITEMS], rank[OITEMS];
            bool valid[OITEMS];
            const uint32_t wbase = base + warp * (32 * OITEMS);
#pragma unroll
            for (int r = 0; r < OITEMS; r++) {
                const uint32_t i = wbase + r * 32 + lane;
                valid[r] = i < end;
                if (base == beg) { key[r] = key0[r]; val[r] = val0[r]; }
                else {
                    key[r] = valid[r] ? __ldcg(&keys_in[i]) : 0u;
                    val[r] = valid[r] ? __ldcg(&vals_in[i]) : 0u;
                }
            }
#pragma unroll
            for (int r = 0; r < OITEMS; r++) {
                const uint32_t d = (key[r] >> shift) & (OBINS - 1);
                const uint32_t active = __ballot_sync(0xffffffffu, valid[r]);
                rank[r] = 0;
                if (valid[r]) {
                    uint32_t peers = active;  // lanes with the same digit, from one ballot per digit bit
#pragma unroll
                    for (int b = 0; b < OBITS; b++) {
                        const uint32_t m = __ballot_sync(active, (d >> b) & 1u);
                        peers &= ((d >> b) & 1u) ? m : ~m;
                    }
                    const int leader = __ffs(peers) - 1;
                    uint32_t old = 0;
                    if ((int)lane == leader) { old = warp_cnt[warp][d]; warp_cnt[warp][d] = old + __popc(peers); }
                    old = __shfl_sync(peers, old, leader);
                    rank[r] = old + __popc(peers & ltmask);
                }
                __syncwarp();
            }
            __syncthreads();
            uint32_t run = 0;
            if (tid < OBINS) {  // the per-warp counts of digit tid -> exclusive offsets
#pragma unroll 8
                for (int w = 0; w < OWARPS; w++) {
                    const uint32_t c = warp_cnt[w][tid];
                    warp_cnt[w][tid] = run;
                    run += c;
                }
            }
            __syncthreads();
#pragma unroll
            for (int r = 0; r < OITEMS; r++) {
                if (valid[r]) {
                    const uint32_t d = (key[r] >> shift) & (OBINS - 1);
                    const uint32_t pos = cursor[d] + warp_cnt[warp][d] + rank[r];
                    keys_out[pos] = key[r];
                    vals_out[pos] = val[r];
                }
            }
            __syncthreads();
            if (tid < OBINS) cursor[tid] += run;
            __syncthreads();
        }
        grid_barrier(ctl);
    }

    // ---------------- C: the busy-strip list (k_tile_ranges) ----------------
    {
        const uint32_t *keys = wb.seg_key[cur];
        const uint32_t ntiles = wb.ntiles;
        uint2 *busy_list = wb.busy_list;
        constexpr uint32_t STEP = OT * 4u;
        uint32_t per = (n + G - 1) / G;
        per = (per + STEP - 1) / STEP * STEP;
        const uint64_t b64 = (uint64_t)blockIdx.x * per;
        const uint32_t rbeg = b64 < n ? (uint32_t)b64 : n, rend = b64 + per < n ? (uint32_t)(b64 + per) : n;
        if (tid < 2) { s_cnt[tid] = 0; s_fill[tid] = 0; }
        __syncthreads();
        // the four keys at i .. i+3 (i a multiple of 4) and the one before them; returns the mask of positions where a
        // bin starts and, of those, the heavy ones
        auto classify4 = [&](uint32_t i, uint32_t k[4], uint32_t &heavy) {
            uint32_t starts = 0;
            heavy = 0;
            if (i >= rend) return starts;
            if (i + 3u < n) {
                const uint4 v = __ldcg(reinterpret_cast<const uint4 *>(keys + i));
                k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; u++) k[u] = i + u < n ? __ldcg(&keys[i + u]) : 0xffffffffu;
            }
            uint32_t prev = i ? __ldcg(&keys[i - 1u]) : ~k[0];
#pragma unroll
            for (uint32_t u = 0; u < 4; u++) {
                if (i + u < rend && k[u] != prev) {
                    starts |= 1u << u;
                    if (i + u + HEAVY_SEGS - 1u < n && __ldcg(&keys[i + u + HEAVY_SEGS - 1u]) == k[u]) heavy |= 1u << u;
                }
                prev = k[u];
            }
            return starts;
        };
        uint32_t nh = 0, nl = 0;
        for (uint32_t i = rbeg + tid * 4u; i < rend; i += STEP) {
            uint32_t k[4], heavy;
            const uint32_t starts = classify4(i, k, heavy);
            nh += (uint32_t)__popc(heavy);
            nl += (uint32_t)__popc(starts & ~heavy);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nh += __shfl_down_sync(0xffffffffu, nh, o); nl += __shfl_down_sync(0xffffffffu, nl, o); }
        if (lane == 0) { if (nh) atomicAdd(&s_cnt[0], nh); if (nl) atomicAdd(&s_cnt[1], nl); }
        __syncthreads();
        if (tid == 0) {
            s_base[0] = s_cnt[0] ? atomicAdd(&ctl->nheavy, s_cnt[0]) : 0u;
            s_base[1] = s_cnt[1] ? atomicAdd(&ctl->nlight, s_cnt[1]) : 0u;
        }
        __syncthreads();
        for (uint32_t i0 = rbeg; i0 < rend; i0 += STEP) {  // (all threads of a warp iterate together: shuffles below)
            const uint32_t i = i0 + tid * 4u;
            uint32_t k[4], heavy;
            const uint32_t starts = classify4(i, k, heavy);
            const uint32_t ch = (uint32_t)__popc(heavy), cl = (uint32_t)__popc(starts & ~heavy);
            uint32_t ih = ch, il = cl;  // ranks inside the warp (inclusive scans), the warp's base from the CTA counters
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
}

// Can the device run k_order?  (cooperative launch, one CTA of OT threads per SM with the sort's shared memory)
static int g_order_ok[64];  // per device: 0 unknown, 1 yes, -1 no
bool order_supported(int device, uint32_t nsm) {
    if (device < 0 || device >= 64) return false;
    if (g_order_ok[device]) return g_order_ok[device] > 0;
    const size_t smem = sizeof(uint32_t) * (size_t)OBINS * (OWARPS + 1);
    int coop = 0, per_sm = 0;
    bool ok = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device) == cudaSuccess && coop;
    ok = ok && cudaFuncSetAttribute(k_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) == cudaSuccess;
    ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_order, OT, smem) == cudaSuccess && per_sm >= 1;
    ok = ok && nsm >= 1 && (size_t)OBINS * nsm <= (size_t)1024 * B200_SMS;  // the [CTA][digit] table fits scan_tmp
    if (!ok) cudaGetLastError();
    g_order_ok[device] = ok ? 1 : -1;
    return ok;
}

// fused: the front end was k_front (nblocks front-end blocks); otherwise (seg_key[0], seg_val[0]) are listed already.
int launch_order(const DrawParams &p, const WorkBuffers &wb, bool fused, int *sorted_buf, cudaStream_t st) {
    int bits = 1;
    while ((1u << bits) < wb.ntiles) bits++;
    const int passes = (bits + OBITS - 1) / OBITS;
    const uint32_t nent = fused ? (p.count + FRONT_FT - 1) / FRONT_FT : 0u;
    const size_t smem = sizeof(uint32_t) * (size_t)OBINS * (OWARPS + 1);
    cudaFuncSetAttribute(k_order, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  // per device and cheap
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = wb.nsm; cfg.blockDim = OT; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = g_pdl ? 2u : 1u;
    cudaLaunchKernelEx(&cfg, k_order, wb, nent, bits, wb.cap_prims / (128u * 64u) + 2u);
    *sorted_buf = passes & 1;
    return 1;
}

}  // namespace fgl
