// fgl_scan_sort.cu -- order-preserving plumbing between the primitive-parallel
// front end and the tile-parallel back end: exclusive scan, a stable LSD radix
// sort of (tile, record) pairs and the per-tile bin ranges.
//
// The reference distributes triangles to goroutines by index (context.go:413-
// 433) and serialises pixels with mutexes; here bins must keep primitive order
// (SURVEY A.12), hence a *stable* sort on the tile id only.
//
// All element counts live in device memory (they are produced by earlier
// kernels of the same draw); grids are sized for the buffer capacity and
// surplus blocks exit.
#include "fgl_internal.h"
#include "fgl_block.cuh"

namespace fgl {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr int RADIX_BINS_HOST = 1024;
constexpr int RADIX_GRID = (int)B200_SMS;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const uint32_t *__restrict__ in, uint32_t n_max, const unsigned int *__restrict__ n_dev,
              uint32_t *__restrict__ tmp, const unsigned int *__restrict__ skip_if) {
    // skip_if (the draw's overflow word): an earlier stage ran out of buffer and the producer of `in` never ran --
    // the input is unwritten memory; scan nothing (the host regrows and re-issues the draw)
    const uint32_t n = (skip_if && *skip_if) ? 0u : (n_dev ? min(*n_dev, n_max) : n_max);
    const uint32_t base = blockIdx.x * SCAN_TILE;
    uint32_t sum = 0;
    if (base < n) {
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) {
            uint32_t i = base + k * SCAN_THREADS + threadIdx.x;
            if (i < n) sum += in[i];
        }
    }
    __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
    uint32_t total;
    block_excl_scan<SCAN_THREADS>(sum, sm, &total);
    if (threadIdx.x == 0) tmp[blockIdx.x] = total;
}

// One block: exclusive scan of the block sums in place; total -> out[n].
__global__ void __launch_bounds__(1024)
k_scan_spine(uint32_t *__restrict__ tmp, uint32_t nblocks, uint32_t *__restrict__ out, uint32_t n_max,
             const unsigned int *__restrict__ n_dev, const ScanSink sink, const unsigned int *__restrict__ skip_if) {
    __shared__ uint32_t sm[1024 / 32 + 1];
    uint32_t carry = 0;
    const bool skipped = skip_if && *skip_if;
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? tmp[i] : 0;
        uint32_t total;
        uint32_t ex = block_excl_scan<1024>(v, sm, &total);
        if (i < nblocks) tmp[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) {
        out[skipped ? 0u : (n_dev ? min(*n_dev, n_max) : n_max)] = carry;
        if (sink.count) *sink.count = carry;
        if (sink.need) *sink.need = carry;
        if (sink.overflow && carry > sink.cap) *sink.overflow |= sink.bit;
        if (sink.clip_n) {
            const unsigned int nc = *sink.clip_n;
            *sink.clip_need = nc;
            if (nc > sink.clip_cap) *sink.overflow |= OVF_CLIP;
        }
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n_max,
            const unsigned int *__restrict__ n_dev, const uint32_t *__restrict__ tmp, const unsigned int *__restrict__ skip_if) {
    const uint32_t n = (skip_if && *skip_if) ? 0u : (n_dev ? min(*n_dev, n_max) : n_max);
    const uint32_t base = blockIdx.x * SCAN_TILE;
    if (base >= n) return;
    // thread t owns SCAN_ITEMS consecutive items
    uint32_t v[SCAN_ITEMS], sum = 0;
    const uint32_t i0 = base + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        v[k] = (i0 + k < n) ? in[i0 + k] : 0;
        sum += v[k];
    }
    __shared__ uint32_t sm[SCAN_THREADS / 32 + 1];
    uint32_t total;
    uint32_t ex = block_excl_scan<SCAN_THREADS>(sum, sm, &total) + tmp[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++) {
        if (i0 + k < n) out[i0 + k] = ex;
        ex += v[k];
    }
}

size_t scan_tmp_words(uint32_t n_max) {
    size_t w = (size_t)(n_max + SCAN_TILE - 1) / SCAN_TILE + 2;
    size_t r = (size_t)RADIX_BINS_HOST * RADIX_GRID;
    return w > r ? w : r;
}

// out[0..n) = exclusive scan, out[n] = total (also for n == 0).
int launch_exclusive_scan(const uint32_t *in, uint32_t *out, uint32_t n_max, const unsigned int *n_dev,
                          uint32_t *tmp, const ScanSink &sink, cudaStream_t st) {
    uint32_t nblocks = (n_max + SCAN_TILE - 1) / SCAN_TILE;
    if (nblocks == 0) nblocks = 1;
    // (the overflow word is sampled by all three kernels before the spine may set this scan's own bit in it: the
    // reduce and the spine read it first thing, the down-sweep skips harmlessly when the spine has just set it)
    k_scan_reduce<<<nblocks, SCAN_THREADS, 0, st>>>(in, n_max, n_dev, tmp, sink.overflow);
    k_scan_spine<<<1, 1024, 0, st>>>(tmp, nblocks, out, n_max, n_dev, sink, sink.overflow);
    k_scan_down<<<nblocks, SCAN_THREADS, 0, st>>>(in, out, n_max, n_dev, tmp, sink.overflow);
    return 3;
}

// ---- stable LSD radix sort, 8- or 10-bit digits ---------------------------------------------
// One CTA of 1024 threads per SM; every CTA owns a contiguous segment of the input, cut into tiles of
// 4096 pairs.  Inside a tile warp w owns the 128 consecutive pairs [128 w, 128 w + 128) and takes them in
// four coalesced rounds of 32, so (warp, round, lane) order IS input order and the ranks below are stable:
//   rank = pairs with the same digit in earlier warps of the tile      (prefix over warp_cnt[.][digit])
//        + pairs with the same digit in earlier rounds of this warp    (running warp_cnt[warp][digit])
//        + lanes with the same digit and a lower lane id               (peer mask: one ballot per digit bit)
// Four pairs per thread amortise the per-tile clearing and the prefix over the warps, which is what makes
// 1024 bins affordable: strip ids of an 8K framebuffer (19 bits) sort in two passes instead of three.

#ifndef FGL_RADIX_BALLOT
#define FGL_RADIX_BALLOT 1  // peer mask from one ballot per digit bit (measured slightly faster than MATCH.ANY: sort stage 40.8 -> 38.8 us at 1080p, 165.5 -> 158.1 us at 8K)
#endif
#ifndef FGL_RADIX_BATCH
#define FGL_RADIX_BATCH 1
#endif
constexpr int RADIX_THREADS = 1024;
constexpr int RADIX_WARPS = RADIX_THREADS / 32;
constexpr int RADIX_ITEMS = 4;
constexpr int RADIX_TILE = RADIX_THREADS * RADIX_ITEMS;

#ifndef FGL_RADIX_CGRID
#define FGL_RADIX_CGRID 1  // CTAs per SM of the chunked 8-bit pass (k_radix_hist + k_radix_scatter_chunked); 2 measured slower
                           // (sort stage 34.9 vs 28.6 us at 1080p, 110.9 vs 108.2 us at 8K: twice the [CTA][digit] rows to add up)
#endif
constexpr int RADIX_GRID_CHUNKED = RADIX_GRID * FGL_RADIX_CGRID;
constexpr uint32_t RADIX_GRAN_CHUNKED = 1024;  // the chunked scatter takes any contiguous split: 32 warps x 32 pairs
__device__ __forceinline__ void radix_segment(uint32_t n, uint32_t &beg, uint32_t &end, uint32_t gran = RADIX_TILE) {
    // contiguous segment of block b; a multiple of the tile so that tiles stay aligned
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + gran - 1) / gran * gran;
    uint64_t b = (uint64_t)blockIdx.x * per;
    beg = b < n ? (uint32_t)b : n;
    uint64_t e = b + per;
    end = e < n ? (uint32_t)e : n;
}

template <int BITS>
__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_hist(const uint32_t *__restrict__ keys, const unsigned int *__restrict__ n_dev, uint32_t n_max, int shift,
             uint32_t *__restrict__ hist /*[grid][BINS]*/, const unsigned int *__restrict__ skip_if, uint32_t gran) {
    constexpr int BINS = 1 << BITS;
    __shared__ uint32_t h[BINS];
    pdl_wait();
    pdl_trigger();
    for (int k = threadIdx.x; k < BINS; k += RADIX_THREADS) h[k] = 0;
    __syncthreads();
    const uint32_t n = (skip_if && *skip_if) ? 0u : min(*n_dev, n_max);  // (overflowed draw: the keys were never written)
    uint32_t beg, end;
    radix_segment(n, beg, end, gran);
    for (uint32_t i = beg + threadIdx.x; i < end; i += RADIX_THREADS)
        atomicAdd(&h[(keys[i] >> shift) & (BINS - 1)], 1u);
    __syncthreads();
    for (int k = threadIdx.x; k < BINS; k += RADIX_THREADS)
        hist[blockIdx.x * BINS + k] = h[k];  // [block][digit]: coalesced here and in the scatter
}

template <int BITS>
__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                const unsigned int *__restrict__ n_dev, uint32_t n_max, int shift,
                const uint32_t *__restrict__ hist /*[grid][BINS], raw counts*/, const unsigned int *__restrict__ skip_if,
                uint32_t *__restrict__ bucket_start /*[BINS + 1] or null: where each digit's run begins in the output*/) {
    constexpr int BINS = 1 << BITS;
    constexpr int DPT = (BINS + RADIX_THREADS - 1) / RADIX_THREADS;  // digits per thread in the bucket-base scan (1)
    pdl_wait();
    pdl_trigger();
    static_assert(DPT == 1, "one digit per thread in the bucket-base scan");
    extern __shared__ uint32_t s_dyn[];
    uint32_t *cursor = s_dyn;                                   // [BINS]
    uint32_t(*warp_cnt)[BINS] = reinterpret_cast<uint32_t(*)[BINS]>(s_dyn + BINS);  // [RADIX_WARPS][BINS]
    __shared__ uint32_t s_scan[RADIX_WARPS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // Every block derives its own bucket bases from the raw histogram (L2 resident): Q threads per digit
        // sum it over all blocks (coalesced across threads) -- no separate scan kernel.
        constexpr uint32_t Q = RADIX_THREADS / BINS;  // 4 (8 bits) or 1 (10 bits)
        const uint32_t d = threadIdx.x & (BINS - 1), q = threadIdx.x / BINS;
        uint32_t row = 0, before = 0;
#if FGL_RADIX_BATCH
        // the histogram rows in batches of independent loads (one L2 round trip per batch instead of one per row)
        constexpr uint32_t BATCH = 10;
        for (uint32_t b0 = q; b0 < gridDim.x; b0 += Q * BATCH) {
            uint32_t v[BATCH];
#pragma unroll
            for (uint32_t k = 0; k < BATCH; k++) {
                const uint32_t b = b0 + k * Q;
                v[k] = b < gridDim.x ? __ldcg(&hist[b * BINS + d]) : 0u;
            }
#pragma unroll
            for (uint32_t k = 0; k < BATCH; k++) {
                row += v[k];
                if (b0 + k * Q < blockIdx.x) before += v[k];
            }
        }
#else
        for (uint32_t b = q; b < gridDim.x; b += Q) {
            const uint32_t v = hist[b * BINS + d];
            row += v;
            if (b < blockIdx.x) before += v;
        }
#endif
        if (Q > 1) {
            warp_cnt[q][d] = row;
            warp_cnt[Q + q][d] = before;
            __syncthreads();
            row = before = 0;
            if (threadIdx.x < BINS)
                for (uint32_t k = 0; k < Q; k++) { row += warp_cnt[k][d]; before += warp_cnt[Q + k][d]; }
        }
        uint32_t total;
        const uint32_t digit_base = block_excl_scan<RADIX_THREADS>(threadIdx.x < BINS ? row : 0u, s_scan, &total);
        if (threadIdx.x < BINS) cursor[d] = digit_base + before;
        if (bucket_start && blockIdx.x == 0 && threadIdx.x < BINS) {
            bucket_start[d] = digit_base;
            if (d == BINS - 1) bucket_start[BINS] = total;
        }
    }
    __syncthreads();
    const uint32_t n = (skip_if && *skip_if) ? 0u : min(*n_dev, n_max);
    uint32_t beg, end;
    radix_segment(n, beg, end);
    const uint32_t ltmask = (1u << lane) - 1u;
    for (uint32_t base = beg; base < end; base += RADIX_TILE) {
        for (int k = threadIdx.x; k < RADIX_WARPS * BINS; k += RADIX_THREADS) (&warp_cnt[0][0])[k] = 0;
        __syncthreads();
        uint32_t key[RADIX_ITEMS], val[RADIX_ITEMS], rank[RADIX_ITEMS];
        bool valid[RADIX_ITEMS];
        const uint32_t wbase = base + warp * (32 * RADIX_ITEMS);
#pragma unroll
        for (int r = 0; r < RADIX_ITEMS; r++) {  // loads first: independent of the ranking
            const uint32_t i = wbase + r * 32 + lane;
            valid[r] = i < end;
            key[r] = valid[r] ? keys_in[i] : 0u;
            val[r] = valid[r] ? vals_in[i] : 0u;
        }
#pragma unroll
        for (int r = 0; r < RADIX_ITEMS; r++) {
            const uint32_t d = (key[r] >> shift) & (BINS - 1);
            const uint32_t active = __ballot_sync(0xffffffffu, valid[r]);
            rank[r] = 0;
            if (valid[r]) {
#if FGL_RADIX_BALLOT
                uint32_t peers = active;  // lanes with the same digit, from one ballot per digit bit
#pragma unroll
                for (int b = 0; b < BITS; b++) {
                    const uint32_t m = __ballot_sync(active, (d >> b) & 1u);
                    peers &= ((d >> b) & 1u) ? m : ~m;
                }
#else
                const uint32_t peers = __match_any_sync(active, d);
#endif
                const int leader = __ffs(peers) - 1;
                uint32_t old = 0;
                if (lane == leader) { old = warp_cnt[warp][d]; warp_cnt[warp][d] = old + __popc(peers); }
                old = __shfl_sync(peers, old, leader);
                rank[r] = old + __popc(peers & ltmask);
            }
            __syncwarp();
        }
        __syncthreads();
        // thread t < BINS turns the per-warp counts of digit t into exclusive offsets
        uint32_t run = 0;
        if (threadIdx.x < BINS) {
#pragma unroll 8
            for (int w = 0; w < RADIX_WARPS; w++) {
                const uint32_t c = warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = run;
                run += c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RADIX_ITEMS; r++) {
            if (valid[r]) {
                const uint32_t d = (key[r] >> shift) & (BINS - 1);
                const uint32_t pos = cursor[d] + warp_cnt[warp][d] + rank[r];
                keys_out[pos] = key[r];
                vals_out[pos] = val[r];
            }
        }
        __syncthreads();
        if (threadIdx.x < BINS) cursor[threadIdx.x] += run;
        __syncthreads();
    }
}

// The 8-bit scatter in the shape that paid off in k_bucket_sort (below): the CTA's contiguous range is cut into 32
// contiguous chunks, one per warp; a warp counts its chunk's digits into a row of its own (one shared-memory atomic per
// pair), one sweep turns the rows into the absolute position of every (warp, digit) run -- digit base over all CTAs +
// this digit in the CTAs before this one (the [CTA][digit] table of k_radix_hist) + this digit in the warps before --
// and the warp places its pairs in steps of 32, ranks from MATCH.ANY: two block barriers per CTA whatever the range
// (the tile loop above has six and a 32 x 256 counter sweep per 4096 pairs), and two CTAs per SM.  Stable for the same
// reason: chunks are contiguous and in order, steps go through a chunk in order, lanes in order.
#ifndef FGL_RADIX_CHUNKED
#define FGL_RADIX_CHUNKED 1
#endif
__global__ void __launch_bounds__(RADIX_THREADS, 2)
k_radix_scatter_chunked(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                        uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                        const unsigned int *__restrict__ n_dev, uint32_t n_max, int shift,
                        const uint32_t *__restrict__ hist /*[grid][256], raw counts*/, const unsigned int *__restrict__ skip_if,
                        uint32_t *__restrict__ bucket_start /*[257] or null*/) {
    constexpr int BINS = 256;
    extern __shared__ uint32_t s_dyn[];
    uint32_t(*part)[BINS] = reinterpret_cast<uint32_t(*)[BINS]>(s_dyn);                       // [8][BINS] partial sums
    uint32_t(*warp_cnt)[BINS] = reinterpret_cast<uint32_t(*)[BINS]>(s_dyn + 8 * BINS);        // [RADIX_WARPS][BINS]
    __shared__ uint32_t s_scan[RADIX_WARPS + 1];
    pdl_wait();
    pdl_trigger();
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t n = (skip_if && *skip_if) ? 0u : min(*n_dev, n_max);
    uint32_t beg, end;
    radix_segment(n, beg, end, RADIX_GRAN_CHUNKED);
    const uint32_t ltmask = (1u << lane) - 1u;
    const uint32_t chunk = (((end - beg + RADIX_WARPS - 1) / RADIX_WARPS) + 31u) & ~31u;
    const uint32_t cbeg = min(beg + warp * chunk, end), cend = min(cbeg + chunk, end);
    // ---- count (this warp's chunk) ...
#pragma unroll
    for (int k = 0; k < BINS / 32; k++) warp_cnt[warp][lane + 32 * k] = 0;
    __syncwarp();
    {
        uint32_t nk = cbeg + lane < cend ? keys_in[cbeg + lane] : 0u;
        for (uint32_t i = cbeg; i < cend; i += 32) {
            const uint32_t k = nk;
            const bool valid = i + lane < cend;
            if (i + 32 < cend) nk = i + 32 + lane < cend ? keys_in[i + 32 + lane] : 0u;  // next step's keys are in flight
            if (valid) atomicAdd(&warp_cnt[warp][(k >> shift) & (BINS - 1)], 1u);
        }
    }
    // ---- ... and the [CTA][digit] table: four threads per digit, batches of independent loads
    {
        constexpr uint32_t Q = RADIX_THREADS / BINS, BATCH = 10;
        const uint32_t d = tid & (BINS - 1), q = tid / BINS;
        uint32_t row = 0, before = 0;
        for (uint32_t b0 = q; b0 < gridDim.x; b0 += Q * BATCH) {
            uint32_t v[BATCH];
#pragma unroll
            for (uint32_t k = 0; k < BATCH; k++) {
                const uint32_t b = b0 + k * Q;
                v[k] = b < gridDim.x ? __ldcg(&hist[b * BINS + d]) : 0u;
            }
#pragma unroll
            for (uint32_t k = 0; k < BATCH; k++) {
                row += v[k];
                if (b0 + k * Q < blockIdx.x) before += v[k];
            }
        }
        part[q][d] = row;
        part[Q + q][d] = before;
    }
    __syncthreads();
    {
        uint32_t row = 0, before = 0;
        if (tid < BINS) {
#pragma unroll
            for (uint32_t k = 0; k < 4; k++) { row += part[k][tid]; before += part[4 + k][tid]; }
        }
        uint32_t total;
        const uint32_t digit_base = block_excl_scan<RADIX_THREADS>(tid < BINS ? row : 0u, s_scan, &total);
        if (tid < BINS) {
            uint32_t c = digit_base + before;
#pragma unroll 8
            for (int w = 0; w < RADIX_WARPS; w++) {
                const uint32_t v = warp_cnt[w][tid];
                warp_cnt[w][tid] = c;
                c += v;
            }
            if (bucket_start && blockIdx.x == 0) {
                bucket_start[tid] = digit_base;
                if (tid == BINS - 1) bucket_start[BINS] = total;
            }
        }
    }
    __syncthreads();
    // ---- place
    {
        uint32_t nk = cbeg + lane < cend ? keys_in[cbeg + lane] : 0u, nv = cbeg + lane < cend ? vals_in[cbeg + lane] : 0u;
        for (uint32_t i = cbeg; i < cend; i += 32) {
            const uint32_t k = nk, v = nv;
            const bool valid = i + lane < cend;
            if (i + 32 < cend) {
                nk = i + 32 + lane < cend ? keys_in[i + 32 + lane] : 0u;
                nv = i + 32 + lane < cend ? vals_in[i + 32 + lane] : 0u;
            }
            const uint32_t d = (k >> shift) & (BINS - 1);
            const uint32_t active = __ballot_sync(0xffffffffu, valid);
            uint32_t peers = 0;
            if (valid) {
                peers = __match_any_sync(active, d);
                const uint32_t pos = warp_cnt[warp][d] + (uint32_t)__popc(peers & ltmask);
                keys_out[pos] = k;
                vals_out[pos] = v;
            }
            __syncwarp();
            if (valid && (peers & ltmask) == 0u) warp_cnt[warp][d] += (uint32_t)__popc(peers);
            __syncwarp();
        }
    }
}

template <int BITS>
static int radix_pass(uint32_t *const key[2], uint32_t *const val[2], int cur, const unsigned int *n_dev, uint32_t n_max,
                      int shift, uint32_t *tmp, cudaStream_t st, const unsigned int *skip_if, uint32_t *bucket_start = nullptr) {
    constexpr int BINS = 1 << BITS;
    if (BITS == 8 && FGL_RADIX_CHUNKED) {
        const size_t smem = sizeof(uint32_t) * (size_t)256 * (RADIX_WARPS + 8);
        cudaFuncSetAttribute(k_radix_scatter_chunked, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_pdl(k_radix_hist<BITS>, RADIX_GRID_CHUNKED, RADIX_THREADS, 0, st, key[cur], n_dev, n_max, shift, tmp, skip_if,
                   RADIX_GRAN_CHUNKED);
        launch_pdl(k_radix_scatter_chunked, RADIX_GRID_CHUNKED, RADIX_THREADS, smem, st, key[cur], val[cur], key[cur ^ 1],
                   val[cur ^ 1], n_dev, n_max, shift, (const uint32_t *)tmp, skip_if, bucket_start);
        return 2;
    }
    launch_pdl(k_radix_hist<BITS>, RADIX_GRID, RADIX_THREADS, 0, st, key[cur], n_dev, n_max, shift, tmp, skip_if, (uint32_t)RADIX_TILE);
    const size_t smem = sizeof(uint32_t) * (size_t)BINS * (RADIX_WARPS + 1);
    // per device and cheap: set on every call (a process may drive several GPUs)
    cudaFuncSetAttribute(k_radix_scatter<BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    launch_pdl(k_radix_scatter<BITS>, RADIX_GRID, RADIX_THREADS, smem, st, key[cur], val[cur], key[cur ^ 1], val[cur ^ 1],
               n_dev, n_max, shift, (const uint32_t *)tmp, skip_if, bucket_start);
    return 2;
}

// ---- binning by strip: one global pass, the rest inside the buckets ---------------------------------------------
// The strip id has 16 bits at 1080p and 19 at 8K: two or three LSD passes (k_radix_hist + k_radix_scatter each) and
// k_tile_ranges -- five or seven short, latency-bound launches over a few megabytes.  Only ONE pass has to be global,
// and the back end does not need the strips in ascending order at all: it needs every strip's segments contiguous,
// in primitive order, and a list of where each such bin begins.  So the global pass is LSD's first -- on the LOW
// eight bits -- which cuts the pairs into 256 buckets, each a contiguous run of the output in primitive order and,
// because neighbouring strips fall into different buckets, all about the same size (the pile-up at a pole of the
// benchmark mesh spreads over all of them; bucketing by the TOP bits was measured first: the pole's bucket alone
// took 29 k cycles).  What is left -- grouping a bucket by the remaining high bits, stable -- is local: ONE CTA per
// bucket, eight bits per local pass, ping-ponging through the two global buffers (L2-resident; a CTA reads only what
// it wrote itself).  Inside the CTA the bucket is cut into 32 contiguous chunks, one per warp: a warp counts the digits
// of its chunk into a row of its own, ONE sweep over the rows turns them into the absolute position of each
// (warp, digit) run, and the warp goes over its chunk again placing the pairs -- two block barriers per pass whatever
// the bucket's length; steps of 32 pairs in chunk order, ranks from one ballot per digit bit: stable.  The histogram
// of the high bits IS the bucket's bin table (bin of high value h: strip h << 8 | bucket, starting at the exclusive
// prefix, as long as its count), so the busy-strip list -- the job of k_tile_ranges -- is written from it without
// another look at the keys.  1080p: 5 launches -> 3; 8K: 7 -> 3.  Up to 20 key bits (12 high bits: 16 KB of bins).
constexpr int BUCKET_MAX_HI_BITS = 12;
#ifndef FGL_BUCKET_MATCH
#define FGL_BUCKET_MATCH 1  // place phase: peers by MATCH.ANY (1) or by one ballot per digit bit (0)
#endif
__global__ void __launch_bounds__(RADIX_THREADS, 2)
k_bucket_sort(uint32_t *key0, uint32_t *val0, uint32_t *key1, uint32_t *val1, int hi_bits,
              const uint32_t *__restrict__ bucket_start, DrawCounters *ctr, uint2 *__restrict__ busy_list,
              uint32_t ntiles, TileCtl *ctl, unsigned long long *__restrict__ group_sums, uint32_t ngroups) {
    constexpr int BINS = 256;
    extern __shared__ uint32_t s_dyn[];
    uint32_t *s_list = s_dyn;                                                     // [2][8] busy strips per warp of digits
    uint32_t(*warp_cnt)[BINS] = reinterpret_cast<uint32_t(*)[BINS]>(s_dyn + BINS);  // [RADIX_WARPS][BINS]
    uint32_t *hfull = s_dyn + BINS * (RADIX_WARPS + 1);                           // [1 << hi_bits], two local passes only
    __shared__ uint32_t s_scan[RADIX_WARPS + 1];
    __shared__ uint32_t s_base[2];
    pdl_wait();
    pdl_trigger();
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    // the group sums of the fused front end are left zeroed for the next draw (also when this one overflowed)
    if (blockIdx.x == 0)
        for (uint32_t g = tid; g < ngroups; g += RADIX_THREADS) group_sums[g] = 0ull;
    if (ctr->overflow) return;  // the draw is going to be re-issued: its keys are incomplete
    const uint32_t beg = bucket_start[blockIdx.x], end = bucket_start[blockIdx.x + 1];
    if (beg >= end) return;
    const uint32_t ltmask = (1u << lane) - 1u;
    const uint32_t chunk = (((end - beg + RADIX_WARPS - 1) / RADIX_WARPS) + 31u) & ~31u;
    const uint32_t cbeg = min(beg + warp * chunk, end), cend = min(cbeg + chunk, end);
    const int passes = (hi_bits + 7) / 8;
    const uint32_t nfull = 1u << hi_bits;
#if !FGL_BUCKET_MATCH
    // lanes with the same digit as this one, among the valid lanes of the step (ballot variant, FGL_BUCKET_MATCH=0)
    auto same_digit_of = [&](uint32_t d, uint32_t active) {
        uint32_t peers = active;
#pragma unroll
        for (int b = 0; b < 8; b++) {
            const uint32_t m = __ballot_sync(active, (d >> b) & 1u);
            peers &= ((d >> b) & 1u) ? m : ~m;
        }
        return peers;
    };
#endif
    if (passes > 1) {
        for (uint32_t k = tid; k < nfull; k += RADIX_THREADS) hfull[k] = 0;
        __syncthreads();
    }
    for (int pass = 0; pass < passes; pass++) {
        const int shift = 8 + 8 * pass;
        const uint32_t mask = (hi_bits - 8 * pass >= 8) ? 255u : ((1u << (hi_bits - 8 * pass)) - 1u);
        const uint32_t *keys_in = (pass & 1) ? key0 : key1, *vals_in = (pass & 1) ? val0 : val1;  // (the global pass wrote buffer 1)
        uint32_t *keys_out = (pass & 1) ? key1 : key0, *vals_out = (pass & 1) ? val1 : val0;
        // ---- count: warp_cnt[warp][d] = pairs of this warp's chunk with digit d
#pragma unroll
        for (int k = 0; k < BINS / 32; k++) warp_cnt[warp][lane + 32 * k] = 0;
        __syncwarp();
        {
            uint32_t nk = cbeg + lane < cend ? keys_in[cbeg + lane] : 0u;
            for (uint32_t i = cbeg; i < cend; i += 32) {
                const uint32_t k = nk;
                const bool valid = i + lane < cend;
                if (i + 32 < cend) nk = i + 32 + lane < cend ? keys_in[i + 32 + lane] : 0u;  // next step's keys are in flight
                const uint32_t d = (k >> shift) & mask;
                // (one shared-memory atomic per pair: this kernel is issue-bound -- 64 resident warps per SM -- and the
                // ballot ranking of the place phase costs 40 issue slots per step; the count does not need the ranks)
                if (valid) atomicAdd(&warp_cnt[warp][d], 1u);
                if (passes > 1 && pass == 0 && valid) atomicAdd(&hfull[(k >> 8) & (nfull - 1u)], 1u);
            }
        }
        __syncthreads();
        // ---- bases: digit totals, digit bases, the absolute position of every (warp, digit) run
        {
            uint32_t c = 0;
            if (tid < BINS) {
#pragma unroll 8
                for (int w = 0; w < RADIX_WARPS; w++) {
                    const uint32_t v = warp_cnt[w][tid];
                    warp_cnt[w][tid] = c;
                    c += v;
                }
            }
            uint32_t total;
            const uint32_t ex = block_excl_scan<RADIX_THREADS>(c, s_scan, &total);
            uint32_t mh = 0, ml = 0;
            if (warp < BINS / 32) {  // (the digits live in the first eight warps)
                const uint32_t base = beg + ex;
#pragma unroll 8
                for (int w = 0; w < RADIX_WARPS; w++) warp_cnt[w][tid] += base;
                if (passes == 1) {  // the digit totals are the bucket's bins
                    const bool heavy = c >= HEAVY_SEGS, light = c > 0u && !heavy;
                    mh = __ballot_sync(0xffffffffu, heavy);
                    ml = __ballot_sync(0xffffffffu, light);
                    if (lane == 0) { s_list[warp] = (uint32_t)__popc(mh); s_list[8 + warp] = (uint32_t)__popc(ml); }
                }
            }
            __syncthreads();
            if (passes == 1) {
                if (tid == 0) {
                    uint32_t nh = 0, nl = 0;
                    for (int w = 0; w < BINS / 32; w++) { nh += s_list[w]; nl += s_list[8 + w]; }
                    s_base[0] = nh ? atomicAdd(&ctl->nheavy, nh) : 0u;
                    s_base[1] = nl ? atomicAdd(&ctl->nlight, nl) : 0u;
                }
                __syncthreads();
                if (warp < BINS / 32 && c > 0u) {
                    const bool heavy = c >= HEAVY_SEGS;
                    uint32_t before = 0;
                    for (uint32_t w = 0; w < warp; w++) before += s_list[(heavy ? 0 : 8) + w];
                    const uint32_t strip = (tid << 8) | blockIdx.x;
                    // (at most ntiles bins; the bound keeps a corrupted key array from writing outside the list)
                    if (heavy) {
                        const uint32_t ph = s_base[0] + before + (uint32_t)__popc(mh & ltmask);
                        if (ph < ntiles) busy_list[ph] = make_uint2(strip, beg + ex);
                    } else {
                        const uint32_t pl = s_base[1] + before + (uint32_t)__popc(ml & ltmask);
                        if (pl < ntiles) busy_list[ntiles - 1u - pl] = make_uint2(strip, beg + ex);
                    }
                }
            }
        }
        // ---- place: the same steps again, every pair to the running position of its (warp, digit) run
        {
            uint32_t nk = cbeg + lane < cend ? keys_in[cbeg + lane] : 0u, nv = cbeg + lane < cend ? vals_in[cbeg + lane] : 0u;
            for (uint32_t i = cbeg; i < cend; i += 32) {
                const uint32_t k = nk, v = nv;
                const bool valid = i + lane < cend;
                if (i + 32 < cend) {
                    nk = i + 32 + lane < cend ? keys_in[i + 32 + lane] : 0u;
                    nv = i + 32 + lane < cend ? vals_in[i + 32 + lane] : 0u;
                }
                const uint32_t d = (k >> shift) & mask;
                const uint32_t active = __ballot_sync(0xffffffffu, valid);
                uint32_t peers = 0;
                if (valid) {
#if FGL_BUCKET_MATCH
                    peers = __match_any_sync(active, d);  // lanes of the step with the same digit: one issue slot
#else
                    peers = same_digit_of(d, active);
#endif
                    const uint32_t pos = warp_cnt[warp][d] + (uint32_t)__popc(peers & ltmask);
                    keys_out[pos] = k;
                    vals_out[pos] = v;
                }
                __syncwarp();
                if (valid && (peers & ltmask) == 0u) warp_cnt[warp][d] += (uint32_t)__popc(peers);
                __syncwarp();
            }
        }
        __syncthreads();  // the next pass reads what other warps of this CTA have just stored
    }
    if (passes > 1) {
        // ---- the bucket's bins from the histogram of all high bits: thread t owns nfull / 1024 consecutive bins
        const uint32_t per = nfull / RADIX_THREADS;  // 1, 2 or 4 (9 <= hi_bits <= 12: nfull >= 512; below 1024 threads idle)
        const uint32_t per1 = per ? per : 1u;
        const uint32_t j0 = tid * per1;
        uint32_t cnt[4] = {0, 0, 0, 0}, sum = 0, nh = 0, nl = 0;
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            if (u < per1 && j0 + u < nfull) {
                cnt[u] = hfull[j0 + u];
                sum += cnt[u];
                if (cnt[u] >= HEAVY_SEGS) nh++; else if (cnt[u]) nl++;
            }
        }
        uint32_t total, tot_list;
        uint32_t pos = beg + block_excl_scan<RADIX_THREADS>(sum, s_scan, &total);
        const uint32_t lex = block_excl_scan<RADIX_THREADS>((nh << 16) | nl, s_scan, &tot_list);  // (<= 4096 bins: no carry)
        if (tid == 0) {
            s_base[0] = (tot_list >> 16) ? atomicAdd(&ctl->nheavy, tot_list >> 16) : 0u;
            s_base[1] = (tot_list & 0xffffu) ? atomicAdd(&ctl->nlight, tot_list & 0xffffu) : 0u;
        }
        __syncthreads();
        uint32_t ph = s_base[0] + (lex >> 16), pl = s_base[1] + (lex & 0xffffu);
#pragma unroll
        for (uint32_t u = 0; u < 4; u++) {
            if (u < per1 && j0 + u < nfull && cnt[u]) {
                const uint32_t strip = ((j0 + u) << 8) | blockIdx.x;
                if (cnt[u] >= HEAVY_SEGS) { if (ph < ntiles) busy_list[ph] = make_uint2(strip, pos); ph++; }
                else { if (pl < ntiles) busy_list[ntiles - 1u - pl] = make_uint2(strip, pos); pl++; }
                pos += cnt[u];
            }
        }
    }
}

// Binning of a draw's segments by strip (8 < bits <= 20): the global pass on the low eight bits, then k_bucket_sort.
// *sorted_buf = the buffer that holds the binned pairs.  tmp: [RADIX_GRID_CHUNKED][256] histogram + 257 bucket starts.
int launch_bin_buckets(uint32_t *const key[2], uint32_t *const val[2], DrawCounters *ctr, uint32_t n_max, int bits,
                       uint32_t *tmp, int *sorted_buf, uint2 *busy_list, uint32_t ntiles, TileCtl *ctl,
                       unsigned long long *group_sums, uint32_t ngroups, cudaStream_t st) {
    const int hi_bits = bits - 8;
    uint32_t *bucket_start = tmp + (size_t)RADIX_GRID_CHUNKED * 256;
    int launches = radix_pass<8>(key, val, 0, &ctr->n_segs, n_max, 0, tmp, st, &ctr->overflow, bucket_start);
    const int passes = (hi_bits + 7) / 8;
    const size_t smem = sizeof(uint32_t) * ((size_t)256 * (RADIX_WARPS + 1) + (passes > 1 ? ((size_t)1 << hi_bits) : 0));
    cudaFuncSetAttribute(k_bucket_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) * (256 * (RADIX_WARPS + 1) + (1 << BUCKET_MAX_HI_BITS))));
    launch_pdl(k_bucket_sort, 256, RADIX_THREADS, smem, st, key[0], val[0], key[1], val[1], hi_bits,
               (const uint32_t *)bucket_start, ctr, busy_list, ntiles, ctl, group_sums, ngroups);
    *sorted_buf = (1 + passes) & 1;
    return launches + 1;
}
bool bin_buckets_ok(int bits) { return bits > 8 && bits - 8 <= BUCKET_MAX_HI_BITS; }

// Stable LSD radix sort of (key, val) on `bits` key bits.  8-bit digits: measured on the 8K benchmark frame
// (3.3 M pairs, 19 key bits) three 8-bit passes take 3 x 43 us, two 10-bit passes 2 x 76 us -- clearing and
// scanning 32 x 1024 counters per tile costs more than the third pass.  10-bit digits are used only when
// they save a pass AND the sort is small enough to be latency-bound (a launch saved > work added).
int launch_sort_pairs(uint32_t *const key[2], uint32_t *const val[2], const unsigned int *n_dev, uint32_t n_max,
                      int bits, uint32_t *tmp, int *sorted_buf, cudaStream_t st, const unsigned int *skip_if) {
    int launches = 0, cur = 0;
    const int passes = (bits + 9) / 10;
    const bool narrow = passes * 8 >= bits || n_max > (1u << 18);
    for (int shift = 0; shift < bits; shift += narrow ? 8 : 10) {
        launches += narrow ? radix_pass<8>(key, val, cur, n_dev, n_max, shift, tmp, st, skip_if)
                           : radix_pass<10>(key, val, cur, n_dev, n_max, shift, tmp, st, skip_if);
        cur ^= 1;
    }
    *sorted_buf = cur;
    return launches;
}

}  // namespace fgl
