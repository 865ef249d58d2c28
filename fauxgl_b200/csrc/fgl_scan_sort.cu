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
constexpr int RADIX_GRID = 148;

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
             const unsigned int *__restrict__ n_dev, const ScanSink sink) {
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
    if (threadIdx.x == 0) {
        out[n_dev ? min(*n_dev, n_max) : n_max] = carry;
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
                          uint32_t *tmp, const ScanSink &sink, cudaStream_t st) {
    uint32_t nblocks = (n_max + SCAN_TILE - 1) / SCAN_TILE;
    if (nblocks == 0) nblocks = 1;
    k_scan_reduce<<<nblocks, SCAN_THREADS, 0, st>>>(in, n_max, n_dev, tmp);
    k_scan_spine<<<1, 1024, 0, st>>>(tmp, nblocks, out, n_max, n_dev, sink);
    k_scan_down<<<nblocks, SCAN_THREADS, 0, st>>>(in, out, n_max, n_dev, tmp);
    return 3;
}

// ---- stable LSD radix sort on 8-bit digits -------------------------------------------

constexpr int RADIX_THREADS = 1024;  // one CTA per SM: many warps hide the load latency of each chunk
constexpr int RADIX_BINS = 256;
constexpr int RADIX_WARPS = RADIX_THREADS / 32;

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
    if (threadIdx.x < RADIX_BINS) h[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t n = min(*n_dev, n_max);
    uint32_t beg, end;
    radix_segment(n, beg, end);
    for (uint32_t i = beg + threadIdx.x; i < end; i += RADIX_THREADS)
        atomicAdd(&h[(keys[i] >> shift) & 0xff], 1u);
    __syncthreads();
    if (threadIdx.x < RADIX_BINS)
        hist[blockIdx.x * RADIX_BINS + threadIdx.x] = h[threadIdx.x];  // [block][digit]: coalesced here and in the scatter
}

__global__ void __launch_bounds__(RADIX_THREADS)
k_radix_scatter(const uint32_t *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out,
                const unsigned int *__restrict__ n_dev, uint32_t n_max, int shift,
                const uint32_t *__restrict__ hist /*[grid][256], raw counts*/) {
    __shared__ uint32_t cursor[RADIX_BINS];
    __shared__ uint32_t warp_cnt[RADIX_WARPS][RADIX_BINS];
    __shared__ uint32_t s_scan[RADIX_WARPS + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // Every block derives its own bucket bases from the raw histogram (151 KB, L2 resident): four threads
        // per digit sum it over all blocks (coalesced across threads) -- no separate scan kernel.
        const uint32_t d = threadIdx.x & (RADIX_BINS - 1), q = threadIdx.x / RADIX_BINS;
        constexpr uint32_t Q = RADIX_THREADS / RADIX_BINS;
        uint32_t row = 0, before = 0;
        for (uint32_t b = q; b < gridDim.x; b += Q) {
            const uint32_t v = hist[b * RADIX_BINS + d];
            row += v;
            if (b < blockIdx.x) before += v;
        }
        warp_cnt[q][d] = row;
        warp_cnt[Q + q][d] = before;
        __syncthreads();
        row = before = 0;
        if (threadIdx.x < RADIX_BINS)
            for (uint32_t k = 0; k < Q; k++) { row += warp_cnt[k][d]; before += warp_cnt[Q + k][d]; }
        uint32_t total;
        const uint32_t digit_base = block_excl_scan<RADIX_THREADS>(row, s_scan, &total);
        if (threadIdx.x < RADIX_BINS) cursor[d] = digit_base + before;
    }
    __syncthreads();
    const uint32_t n = min(*n_dev, n_max);
    uint32_t beg, end;
    radix_segment(n, beg, end);
    for (uint32_t base = beg; base < end; base += RADIX_THREADS) {
        for (int k = threadIdx.x; k < RADIX_WARPS * RADIX_BINS; k += RADIX_THREADS) (&warp_cnt[0][0])[k] = 0;
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
        // thread t < 256 turns the per-warp counts of digit t into exclusive offsets
        uint32_t run = 0;
        if (threadIdx.x < RADIX_BINS) {
#pragma unroll 8
            for (int w = 0; w < RADIX_WARPS; w++) {
                uint32_t c = warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = run;
                run += c;
            }
        }
        __syncthreads();
        if (valid) {
            uint32_t pos = cursor[d] + warp_cnt[warp][d] + rank;
            keys_out[pos] = key;
            vals_out[pos] = val;
        }
        __syncthreads();
        if (threadIdx.x < RADIX_BINS) cursor[threadIdx.x] += run;
        __syncthreads();
    }
}


// hist[256][G] -> exclusive scan in (digit-major, block-minor) order == global base of
// each (digit, block) bucket.  One CTA: the whole histogram (148 KB) is staged in shared
// memory with coalesced loads, each thread scans a contiguous slice there (stride 37 words:
// conflict-free), one block scan joins the slices, and the result streams back coalesced.
__global__ void __launch_bounds__(1024)
k_radix_scan(uint32_t *__restrict__ hist, uint32_t n) {
    extern __shared__ uint32_t s_hist[];
    __shared__ uint32_t sm[1024 / 32 + 1];
    for (uint32_t i = threadIdx.x; i < n; i += 1024) s_hist[i] = hist[i];
    __syncthreads();
    const uint32_t per = (n + 1023) / 1024;
    const uint32_t beg = min(threadIdx.x * per, n), end = min(beg + per, n);
    uint32_t sum = 0;
    for (uint32_t i = beg; i < end; i++) sum += s_hist[i];
    uint32_t total;
    uint32_t run = block_excl_scan<1024>(sum, sm, &total);
    for (uint32_t i = beg; i < end; i++) {
        const uint32_t v = s_hist[i];
        s_hist[i] = run;
        run += v;
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += 1024) hist[i] = s_hist[i];
}

// Stable LSD radix sort of (key, val) on `bits` key bits, 8 bits per pass.
int launch_sort_pairs(uint32_t *const key[2], uint32_t *const val[2], const unsigned int *n_dev, uint32_t n_max,
                      int bits, uint32_t *tmp, int *sorted_buf, cudaStream_t st) {
    int launches = 0, cur = 0;
    const int G = RADIX_GRID;
    const size_t scan_smem = sizeof(uint32_t) * RADIX_BINS * G;
    // per device and cheap: set on every call (a process may drive several GPUs)
    cudaFuncSetAttribute(k_radix_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scan_smem);
    for (int shift = 0; shift < bits; shift += 8) {
        k_radix_hist<<<G, RADIX_THREADS, 0, st>>>(key[cur], n_dev, n_max, shift, tmp);
        k_radix_scatter<<<G, RADIX_THREADS, 0, st>>>(key[cur], val[cur], key[cur ^ 1], val[cur ^ 1], n_dev, n_max, shift,
                                                     tmp);
        cur ^= 1;
        launches += 2;
    }
    *sorted_buf = cur;
    return launches;
}

}  // namespace fgl
