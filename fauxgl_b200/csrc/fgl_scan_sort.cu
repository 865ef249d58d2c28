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
constexpr int RADIX_BINS_HOST = 256;
constexpr int RADIX_GRID = 148 * 2;

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_reduce(const uint32_t *__restrict__ in, uint32_t n_max, const unsigned int *__restrict__ n_dev,
              uint32_t *__restrict__ tmp) {
    const uint32_t n = n_dev ? min(*n_dev, n_max) : n_max;
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
             const unsigned int *__restrict__ n_dev) {
    __shared__ uint32_t sm[1024 / 32 + 1];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? tmp[i] : 0;
        uint32_t total;
        uint32_t ex = block_excl_scan<1024>(v, sm, &total);
        if (i < nblocks) tmp[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) out[n_dev ? min(*n_dev, n_max) : n_max] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_down(const uint32_t *__restrict__ in, uint32_t *__restrict__ out, uint32_t n_max,
            const unsigned int *__restrict__ n_dev, const uint32_t *__restrict__ tmp) {
    const uint32_t n = n_dev ? min(*n_dev, n_max) : n_max;
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
                          uint32_t *tmp, cudaStream_t st) {
    uint32_t nblocks = (n_max + SCAN_TILE - 1) / SCAN_TILE;
    if (nblocks == 0) nblocks = 1;
    k_scan_reduce<<<nblocks, SCAN_THREADS, 0, st>>>(in, n_max, n_dev, tmp);
    k_scan_spine<<<1, 1024, 0, st>>>(tmp, nblocks, out, n_max, n_dev);
    k_scan_down<<<nblocks, SCAN_THREADS, 0, st>>>(in, out, n_max, n_dev, tmp);
    return 3;
}

// ---- stable LSD radix sort on 8-bit digits -------------------------------------------

constexpr int RADIX_THREADS = 256;
constexpr int RADIX_BINS = 256;

__device__ __forceinline__ void radix_segment(uint32_t n, uint32_t &beg, uint32_t &end) {
    // contiguous segment of block b; multiple of RADIX_THREADS so sub-tiles stay aligned
    uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    per = (per + RADIX_THREADS - 1) / RADIX_THREADS * RADIX_THREADS;
    uint64_t b = (uint64_t)blockIdx.x * per;
    beg = b < n ? (uint32_t)b : n;
    uint64_t e = b + per;
    end = e < n ? (uint32_t)e : n;
}

__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_hist(const uint32_t *__restrict__ keys, const unsigned int *__restrict__ n_dev, uint32_t n_max, int shift,
             uint32_t *__restrict__ hist /*[256][grid]*/) {
    __shared__ uint32_t h[RADIX_BINS];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = min(*n_dev, n_max);
    uint32_t beg, end;
    radix_segment(n, beg, end);
    for (uint32_t i = beg + threadIdx.x; i < end; i += RADIX_THREADS)
        atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
    __syncthreads();
    hist[threadIdx.x * gridDim.x + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                const unsigned int *__restrict__ n_dev, uint32_t n_max, int shift,
                const uint32_t *__restrict__ hist_scanned /*[256][grid]*/) {
    __shared__ uint32_t cursor[RADIX_BINS];
    __shared__ uint32_t warp_cnt[RADIX_THREADS / 32][RADIX_BINS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cursor[threadIdx.x] = hist_scanned[threadIdx.x * gridDim.x + blockIdx.x];
    const uint32_t n = min(*n_dev, n_max);
    uint32_t beg, end;
    radix_segment(n, beg, end);
    for (uint32_t base = beg; base < end; base += RADIX_THREADS) {
#pragma unroll
        for (int w = 0; w < RADIX_THREADS / 32; w++) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const uint32_t i = base + threadIdx.x;
        const bool valid = i < end;
        uint32_t key = 0, val = 0, d = 0;
        if (valid) { key = keys_in[i]; val = vals_in[i]; d = (key >> shift) & 0xff; }
        // stable rank inside the warp: lanes with the same digit and a lower lane id
        const uint32_t active = __ballot_sync(0xffffffffu, valid);
        uint32_t rank = 0;
        if (valid) {
            uint32_t peers = __match_any_sync(active, d);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0) warp_cnt[warp][d] = __popc(peers);
        }
        __syncthreads();
        // thread t turns the per-warp counts of digit t into exclusive offsets
        {
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < RADIX_THREADS / 32; w++) {
                uint32_t c = warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = run;
                run += c;
            }
            // stash the digit total in the high half via a second array-free trick:
            // cursor is advanced after the scatter below, so keep `run` in a register.
            __syncthreads();
            if (valid) {
                uint32_t pos = cursor[d] + warp_cnt[warp][d] + rank;
                keys_out[pos] = key;
                vals_out[pos] = val;
            }
            __syncthreads();
            cursor[threadIdx.x] += run;
        }
        __syncthreads();
    }
}

// hist[256][G] -> exclusive scan in (digit-major, block-minor) order == global base of
// each (digit, block) bucket.  Single block.
__global__ void __launch_bounds__(1024)
k_radix_scan(uint32_t *__restrict__ hist, uint32_t n) {
    __shared__ uint32_t sm[1024 / 32 + 1];
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < n ? hist[i] : 0;
        uint32_t total;
        uint32_t ex = block_excl_scan<1024>(v, sm, &total);
        if (i < n) hist[i] = carry + ex;
        carry += total;
    }
}

__global__ void k_tile_ranges(const uint32_t *__restrict__ keys, const unsigned int *__restrict__ n_dev, uint32_t n_max,
                              uint32_t *__restrict__ tile_start, uint32_t *__restrict__ tile_end) {
    const uint32_t n = min(*n_dev, n_max);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t k = keys[i];
        if (i == 0 || keys[i - 1] != k) tile_start[k] = i;
        if (i == n - 1 || keys[i + 1] != k) tile_end[k] = i + 1;
    }
}

// Pair generation: thread per pair; the owning record is found by binary search
// in the scanned per-record pair counts (load-balanced whatever the bbox sizes).
__global__ void __launch_bounds__(256)
k_emit_pairs(const uint32_t *__restrict__ rec_pair_off, const RecTiles *__restrict__ rec_tiles,
             const DrawCounters *__restrict__ ctr, uint32_t cap_records, uint32_t cap_pairs, uint32_t tiles_x,
             uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t nrec = min(ctr->n_records, cap_records);
    const uint32_t npairs = min(ctr->n_pairs, cap_pairs);
    if (ctr->overflow) return;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < npairs; p += gridDim.x * blockDim.x) {
        // largest r with rec_pair_off[r] <= p
        uint32_t lo = 0, hi = nrec;  // invariant: off[lo] <= p < off[hi]
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (rec_pair_off[mid] <= p) lo = mid; else hi = mid;
        }
        const RecTiles t = rec_tiles[lo];
        const uint32_t q = p - rec_pair_off[lo];
        const uint32_t tw = (uint32_t)t.tx1 - t.tx0 + 1u;
        const uint32_t ty = t.ty0 + q / tw, tx = t.tx0 + q % tw;
        keys[p] = ty * tiles_x + tx;
        vals[p] = lo;
    }
}

__global__ void k_set_npairs(DrawCounters *ctr, const uint32_t *__restrict__ rec_pair_off, uint32_t cap_records,
                             uint32_t cap_pairs) {
    const uint32_t nrec = min(ctr->n_records, cap_records);
    const uint32_t total = rec_pair_off[nrec];
    ctr->n_pairs = total;
    ctr->need_pairs = total;
    if (total > cap_pairs) ctr->overflow |= 2u;
}

int launch_binning(const DrawParams &p, const WorkBuffers &wb, int *sorted_buf, cudaStream_t st) {
    int launches = 0;
    // 1. per-record pair counts -> offsets; total -> counters
    launches += launch_exclusive_scan(wb.rec_npairs, wb.rec_pair_off, wb.cap_records, &wb.counters->n_records,
                                      wb.scan_tmp, st);
    k_set_npairs<<<1, 1, 0, st>>>(wb.counters, wb.rec_pair_off, wb.cap_records, wb.cap_pairs);
    launches++;
    // 2. pairs in record order
    const int G = RADIX_GRID;
    k_emit_pairs<<<148 * 8, 256, 0, st>>>(wb.rec_pair_off, wb.rec_tiles, wb.counters, wb.cap_records, wb.cap_pairs,
                                         (uint32_t)p.tiles_x, wb.pair_key[0], wb.pair_val[0]);
    launches++;
    // 3. stable sort by tile id
    int bits = 1;
    while ((1u << bits) < wb.ntiles) bits++;
    int cur = 0;
    for (int shift = 0; shift < bits; shift += 8) {
        k_radix_hist<<<G, RADIX_THREADS, 0, st>>>(wb.pair_key[cur], &wb.counters->n_pairs, wb.cap_pairs, shift,
                                                  wb.scan_tmp);
        k_radix_scan<<<1, 1024, 0, st>>>(wb.scan_tmp, RADIX_BINS * G);
        k_radix_scatter<<<G, RADIX_THREADS, 0, st>>>(wb.pair_key[cur], wb.pair_val[cur], wb.pair_key[cur ^ 1],
                                                     wb.pair_val[cur ^ 1], &wb.counters->n_pairs, wb.cap_pairs, shift,
                                                     wb.scan_tmp);
        cur ^= 1;
        launches += 3;
    }
    // 4. bin ranges
    cudaMemsetAsync(wb.tile_start, 0, sizeof(uint32_t) * wb.ntiles, st);
    cudaMemsetAsync(wb.tile_end, 0, sizeof(uint32_t) * wb.ntiles, st);
    k_tile_ranges<<<148 * 4, 256, 0, st>>>(wb.pair_key[cur], &wb.counters->n_pairs, wb.cap_pairs, wb.tile_start,
                                          wb.tile_end);
    launches++;
    *sorted_buf = cur;
    return launches;
}

}  // namespace fgl
