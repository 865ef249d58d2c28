// fgl_post.cu -- framebuffer kernels around the draw: clears, the SSAA resolve
// and the pack/unpack/min kernels of the sort-last depth composite.
#include "fgl_internal.h"
#include "fgl_math.cuh"

namespace fgl {

// ---- clears: ClearColorBufferWith / ClearDepthBufferWith, context.go:119-141 --------------
__global__ void k_clear_color(uint4 *__restrict__ color4, uint32_t *__restrict__ color, size_t npix, uint32_t rgba) {
    const size_t n4 = npix / 4;
    const uint4 v = make_uint4(rgba, rgba, rgba, rgba);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        color4[i] = v;
    if (blockIdx.x == 0 && threadIdx.x < (npix & 3)) color[n4 * 4 + threadIdx.x] = rgba;
}
__global__ void k_clear_depth(double2 *__restrict__ depth2, double *__restrict__ depth, size_t npix, double value) {
    const size_t n2 = npix / 2;
    const double2 v = make_double2(value, value);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x)
        depth2[i] = v;
    if (blockIdx.x == 0 && threadIdx.x == 0 && (npix & 1)) depth[npix - 1] = value;
}
int launch_clear_color(uint32_t *color, size_t npix, uint32_t rgba, cudaStream_t st) {
    k_clear_color<<<GRID_WAVE, 256, 0, st>>>(reinterpret_cast<uint4 *>(color), color, npix, rgba);
    return 1;
}
int launch_clear_depth(double *depth, size_t npix, double v, cudaStream_t st) {
    k_clear_depth<<<GRID_WAVE, 256, 0, st>>>(reinterpret_cast<double2 *>(depth), depth, npix, v);
    return 1;
}

// ---- SSAA resolve -------------------------------------------------------------------------------
// resize.Resize(w/f, h/f, img, resize.Bilinear) of github.com/nfnt/resize on an
// *image.NRGBA (the call every reference example makes, e.g. examples/teapot.go:60;
// the library is external and unpinned -- restated from its published algorithm:
// createWeights8 + resizeNRGBA + resizeRGBA).  Two separable passes of a 2f-tap tent
// with int16 weights round(tent*256), int32 accumulation, truncating division by the
// weight sum and clamping after each pass; the first pass premultiplies by alpha.
// For an integer factor the weights are the same for every output pixel.
constexpr int RES_OX = 32, RES_OY = 8;  // output tile per CTA
constexpr int RES_MAX_TAPS = 32;        // factor <= 16

struct ResolveWeights {
    int16_t coeff[RES_MAX_TAPS];
    int flen;    // 2 * factor
    int start0;  // start(y) = factor * y + start0
    int sum;
};

__device__ __forceinline__ uint32_t clamp_u8(int v) { return v < 0 ? 0u : (v > 255 ? 255u : (uint32_t)v); }

__global__ void __launch_bounds__(RES_OX *RES_OY)
k_resolve(const uint32_t *__restrict__ src, int sw, int sh, uint32_t *__restrict__ dst, int dw, int dh, int factor,
          const ResolveWeights W) {
    extern __shared__ uint32_t s_temp[];  // [rows][RES_OX] horizontally filtered, premultiplied
    const int ox0 = blockIdx.x * RES_OX, oy0 = blockIdx.y * RES_OY;
    const int rows = (RES_OY - 1) * factor + W.flen;  // source rows this tile needs
    const int row_start = factor * oy0 + W.start0;
    const int tid = threadIdx.y * RES_OX + threadIdx.x;
    // Fast path of the usual 4x4 SSAA (8 taps, weights <= 255, sum a power of two) for opaque pixels away from
    // the left/right border: three aligned 16-byte loads bring in the 8 taps, the bytes are transposed into one
    // word per channel (PRMT) and each channel is two 4-way byte dot products (DP4A) -- the same integer
    // arithmetic as the generic loop below (with alpha == 255 the premultiplication r * a / 255 is the identity).
    const bool fast4 = factor == 4 && W.flen == 8 && W.sum == 1024 && (sw & 3) == 0;
    uint32_t cA = 0, cB = 0;
    if (fast4) {
        cA = (uint32_t)W.coeff[0] | ((uint32_t)W.coeff[1] << 8) | ((uint32_t)W.coeff[2] << 16) | ((uint32_t)W.coeff[3] << 24);
        cB = (uint32_t)W.coeff[4] | ((uint32_t)W.coeff[5] << 8) | ((uint32_t)W.coeff[6] << 16) | ((uint32_t)W.coeff[7] << 24);
    }
    // pass 1: horizontal (resizeNRGBA)
    for (int idx = tid; idx < rows * RES_OX; idx += RES_OX * RES_OY) {
        const int r = idx / RES_OX, c = idx % RES_OX;
        const int ox = ox0 + c;
        int sy = row_start + r;
        uint32_t out = 0;
        if (ox < dw) {
            sy = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);  // pass 2 clamps its (row) index the same way
            const uint32_t *row = src + (size_t)sy * sw;
            const int start = factor * ox + W.start0;
            if (fast4 && start >= 2 && start + 10 <= sw) {  // taps start .. start+7 = pixels 2..9 of the 12 loaded
                const uint4 *q = reinterpret_cast<const uint4 *>(row + (start - 2));
                const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
                const uint32_t p0 = v0.z, p1 = v0.w, p2 = v1.x, p3 = v1.y, p4 = v1.z, p5 = v1.w, p6 = v2.x, p7 = v2.y;
                if (((p0 & p1 & p2 & p3 & p4 & p5 & p6 & p7) >> 24) == 0xffu) {
                    const uint32_t lo = __byte_perm(p0, p1, 0x5140), hi = __byte_perm(p2, p3, 0x5140);      // r r g g
                    const uint32_t lo2 = __byte_perm(p0, p1, 0x7362), hi2 = __byte_perm(p2, p3, 0x7362);    // b b a a
                    const uint32_t mo = __byte_perm(p4, p5, 0x5140), mi = __byte_perm(p6, p7, 0x5140);
                    const uint32_t mo2 = __byte_perm(p4, p5, 0x7362), mi2 = __byte_perm(p6, p7, 0x7362);
                    const uint32_t rr = __dp4a(__byte_perm(mo, mi, 0x5410), cB, __dp4a(__byte_perm(lo, hi, 0x5410), cA, 0u));
                    const uint32_t gg = __dp4a(__byte_perm(mo, mi, 0x7632), cB, __dp4a(__byte_perm(lo, hi, 0x7632), cA, 0u));
                    const uint32_t bb = __dp4a(__byte_perm(mo2, mi2, 0x5410), cB, __dp4a(__byte_perm(lo2, hi2, 0x5410), cA, 0u));
                    s_temp[idx] = (rr >> 10) | ((gg >> 10) << 8) | ((bb >> 10) << 16) | 0xff000000u;  // alpha: 1024 * 255 >> 10
                    continue;
                }
            }
            int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
            for (int i = 0; i < W.flen; i++) {
                const int coeff = W.coeff[i];
                if (coeff == 0) continue;
                int xi = start + i;
                xi = xi < 0 ? 0 : (xi >= sw - 1 ? sw - 1 : xi);
                const uint32_t px = __ldg(row + xi);
                const int a = (int)(px >> 24);
                const int r8 = (int)(px & 0xff) * a / 0xff;
                const int g8 = (int)((px >> 8) & 0xff) * a / 0xff;
                const int b8 = (int)((px >> 16) & 0xff) * a / 0xff;
                acc0 += coeff * r8; acc1 += coeff * g8; acc2 += coeff * b8; acc3 += coeff * a;
            }
            out = clamp_u8(acc0 / W.sum) | (clamp_u8(acc1 / W.sum) << 8) | (clamp_u8(acc2 / W.sum) << 16) |
                  (clamp_u8(acc3 / W.sum) << 24);
        }
        s_temp[idx] = out;
    }
    __syncthreads();
    // pass 2: vertical (resizeRGBA on the transposed temporary)
    const int ox = ox0 + threadIdx.x, oy = oy0 + threadIdx.y;
    if (ox < dw && oy < dh && fast4) {
        // the same fast path for the vertical pass: eight taps from shared memory, bytes transposed per channel,
        // two DP4A per channel; sum == 1024 makes the truncating division a shift and the clamp a no-op
        // (the weights add up to 1024 and every byte is <= 255).  No alpha condition here: pass 2 does not premultiply.
        const uint32_t *col = s_temp + (threadIdx.y * 4) * RES_OX + threadIdx.x;
        const uint32_t p0 = col[0], p1 = col[RES_OX], p2 = col[2 * RES_OX], p3 = col[3 * RES_OX];
        const uint32_t p4 = col[4 * RES_OX], p5 = col[5 * RES_OX], p6 = col[6 * RES_OX], p7 = col[7 * RES_OX];
        const uint32_t lo = __byte_perm(p0, p1, 0x5140), hi = __byte_perm(p2, p3, 0x5140);      // r r g g
        const uint32_t lo2 = __byte_perm(p0, p1, 0x7362), hi2 = __byte_perm(p2, p3, 0x7362);    // b b a a
        const uint32_t mo = __byte_perm(p4, p5, 0x5140), mi = __byte_perm(p6, p7, 0x5140);
        const uint32_t mo2 = __byte_perm(p4, p5, 0x7362), mi2 = __byte_perm(p6, p7, 0x7362);
        const uint32_t rr = __dp4a(__byte_perm(mo, mi, 0x5410), cB, __dp4a(__byte_perm(lo, hi, 0x5410), cA, 0u));
        const uint32_t gg = __dp4a(__byte_perm(mo, mi, 0x7632), cB, __dp4a(__byte_perm(lo, hi, 0x7632), cA, 0u));
        const uint32_t bb = __dp4a(__byte_perm(mo2, mi2, 0x5410), cB, __dp4a(__byte_perm(lo2, hi2, 0x5410), cA, 0u));
        const uint32_t aa = __dp4a(__byte_perm(mo2, mi2, 0x7632), cB, __dp4a(__byte_perm(lo2, hi2, 0x7632), cA, 0u));
        dst[(size_t)oy * dw + ox] = (rr >> 10) | ((gg >> 10) << 8) | ((bb >> 10) << 16) | ((aa >> 10) << 24);
    } else if (ox < dw && oy < dh) {
        int acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
        for (int i = 0; i < W.flen; i++) {
            const int coeff = W.coeff[i];
            if (coeff == 0) continue;
            const uint32_t px = s_temp[(threadIdx.y * factor + i) * RES_OX + threadIdx.x];
            acc0 += coeff * (int)(px & 0xff); acc1 += coeff * (int)((px >> 8) & 0xff);
            acc2 += coeff * (int)((px >> 16) & 0xff); acc3 += coeff * (int)(px >> 24);
        }
        dst[(size_t)oy * dw + ox] = clamp_u8(acc0 / W.sum) | (clamp_u8(acc1 / W.sum) << 8) |
                                    (clamp_u8(acc2 / W.sum) << 16) | (clamp_u8(acc3 / W.sum) << 24);
    }
}

int launch_resolve(const uint32_t *src, int sw, int sh, uint32_t *dst, int factor, cudaStream_t st) {
    const int dw = sw / factor, dh = sh / factor;
    // createWeights8 with scale == factor, blur == 1, taps == 2 (Bilinear)
    ResolveWeights W;
    const double scale = (double)factor;
    W.flen = 2 * factor;
    const double ff = 1.0 / scale;
    double ix = scale * 0.5 - 0.5;  // y == 0
    const int start = (int)ix - W.flen / 2 + 1;
    W.start0 = start;
    ix -= (double)start;
    W.sum = 0;
    for (int i = 0; i < W.flen; i++) {
        double in = (ix - (double)i) * ff;
        if (in < 0) in = -in;
        const double k = in <= 1 ? 1 - in : 0;
        W.coeff[i] = (int16_t)(k * 256);
        W.sum += W.coeff[i];
    }
    dim3 grid((dw + RES_OX - 1) / RES_OX, (dh + RES_OY - 1) / RES_OY), block(RES_OX, RES_OY);
    const size_t smem = sizeof(uint32_t) * (size_t)((RES_OY - 1) * factor + W.flen) * RES_OX;
    k_resolve<<<grid, block, smem, st>>>(src, sw, sh, dst, dw, dh, factor, W);
    return 1;
}

// ---- sort-last composite keys -----------------------------------------------------------
// key = (depth32 << 32 | R<<24 | G<<16 | B<<8 | A) ^ 2^63, stored as int64: the bias makes
// *signed* 64-bit min (ncclInt64 / torch int64, which every collective library has)
// order keys by depth first, then colour.
__device__ __forceinline__ uint32_t depth32(double z) {
    if (!(z >= 0)) return 0u;          // NaN / negative roundoff
    if (z > 1) return 0xFFFFFFFFu;     // cleared (MaxFloat64) or beyond the far plane
    if (z == 1) return 0xFFFFFFFEu;
    return (uint32_t)(z * 4294967295.0);
}
__global__ void k_composite_pack(const uint32_t *__restrict__ color, const double *__restrict__ depth,
                                 unsigned long long *__restrict__ keys, size_t npix, unsigned long long bias) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t c = color[i];
        const uint32_t be = ((c & 0xff) << 24) | (((c >> 8) & 0xff) << 16) | (((c >> 16) & 0xff) << 8) | (c >> 24);
        keys[i] = (((unsigned long long)depth32(depth[i]) << 32) | be) ^ bias;
    }
}
__global__ void k_composite_unpack(uint32_t *__restrict__ color, double *__restrict__ depth,
                                   const unsigned long long *__restrict__ keys, size_t npix, unsigned long long bias) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const unsigned long long k = keys[i] ^ bias;
        const uint32_t be = (uint32_t)k, d32 = (uint32_t)(k >> 32);
        color[i] = (be >> 24) | (((be >> 16) & 0xff) << 8) | (((be >> 8) & 0xff) << 16) | ((be & 0xff) << 24);
        depth[i] = d32 == 0xFFFFFFFFu ? 1.7976931348623157e308 : (double)d32 / 4294967295.0;
    }
}
__global__ void k_composite_min(long long *__restrict__ inout, const long long *__restrict__ other, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const long long a = inout[i], b = other[i];
        inout[i] = b < a ? b : a;
    }
}
// bias: keys ^ 2^63, so that a SIGNED 64-bit min (torch int64, gloo) orders them; unsigned collectives (ncclUint64) take them as they are
int launch_composite_pack(const uint32_t *color, const double *depth, unsigned long long *keys, size_t npix,
                          cudaStream_t st, bool bias) {
    k_composite_pack<<<GRID_WAVE, 256, 0, st>>>(color, depth, keys, npix, bias ? 0x8000000000000000ull : 0ull);
    return 1;
}
int launch_composite_unpack(uint32_t *color, double *depth, const unsigned long long *keys, size_t npix,
                            cudaStream_t st, bool bias) {
    k_composite_unpack<<<GRID_WAVE, 256, 0, st>>>(color, depth, keys, npix, bias ? 0x8000000000000000ull : 0ull);
    return 1;
}
// ---- peer-memory composite: compute + exchange in ONE kernel over NVLink ---------------------------------
// Rank r owns the pixel stripe [px0, px1).  For each of its pixels it loads the float64 depth of every
// rank straight from that rank's depth buffer (P2P loads through NVLink/NVSwitch), picks the smallest --
// on a tie the HIGHER rank, i.e. the later triangle range, which is what the reference's `<=` retest
// (context.go:248) gives in index order -- fetches only the winner's colour, and stores depth + colour into
// every rank's buffers (P2P stores).  Unlike the 32-bit packed key this keeps the float64 depth, so the
// composite equals a single-GPU render bit for bit (order-independent state only: see multigpu.py).
struct PeerBuffers {
    uint32_t *color[FGL_MAX_PEERS];
    double *depth[FGL_MAX_PEERS];
};
__global__ void __launch_bounds__(256)
k_composite_peer(const PeerBuffers P, int nranks, size_t px0, size_t px1) {
    for (size_t i = px0 + blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < px1; i += (size_t)gridDim.x * blockDim.x) {
        double d[FGL_MAX_PEERS];
#pragma unroll
        for (int r = 0; r < FGL_MAX_PEERS; r++)
            if (r < nranks) d[r] = __ldcv(P.depth[r] + i);  // all loads in flight before the first compare
        double best = d[0];
        int win = 0;
#pragma unroll
        for (int r = 1; r < FGL_MAX_PEERS; r++)
            if (r < nranks && d[r] <= best) { best = d[r]; win = r; }
        const uint32_t c = __ldcv(P.color[win] + i);
#pragma unroll
        for (int r = 0; r < FGL_MAX_PEERS; r++)
            if (r < nranks) {
                if (d[r] != best) P.depth[r][i] = best;
                P.color[r][i] = c;
            }
    }
}
int launch_composite_peer(uint32_t *const *color, double *const *depth, int nranks, size_t px0, size_t px1,
                          cudaStream_t st) {
    PeerBuffers P;
    for (int r = 0; r < FGL_MAX_PEERS; r++) {
        P.color[r] = r < nranks ? color[r] : nullptr;
        P.depth[r] = r < nranks ? depth[r] : nullptr;
    }
    if (px1 > px0) k_composite_peer<<<GRID_WAVE, 256, 0, st>>>(P, nranks, px0, px1);
    return 1;
}

int launch_composite_min(unsigned long long *inout, const unsigned long long *other, size_t n, cudaStream_t st) {
    k_composite_min<<<GRID_WAVE, 256, 0, st>>>(reinterpret_cast<long long *>(inout),
                                            reinterpret_cast<const long long *>(other), n);
    return 1;
}

// ---- fragment-rate probe ---------------------------------------------------------------------------------
// Every thread issues atomicMin(u64) on pseudo-random words (a 64-bit LCG per thread, so neighbouring lanes hit
// unrelated addresses, like fragments of unrelated triangles), with the depth-like key decreasing slowly so that a
// share of the operations actually writes.
__global__ void __launch_bounds__(256)
k_atomic_probe(unsigned long long *__restrict__ buf, size_t words, unsigned long long ops_per_thread) {
    unsigned long long x = 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1ull);
    for (unsigned long long k = 0; k < ops_per_thread; k++) {
        x = x * 6364136223846793005ull + 1442695040888963407ull;
        const size_t w = (size_t)((x >> 20) % words);
        atomicMin(&buf[w], (x >> 1) | 1ull);
    }
}
// Self-check of the branch-free division helpers (fgl_math.cuh) against the operator: every thread draws operand pairs
// from a counter-based generator -- raw 64-bit patterns (NaN, infinities, subnormals, zeros included), values of the
// magnitudes the rasteriser sees, operands with tiny and huge exponents, exact zeros -- and compares, wherever the
// helper's range test says "fast path", its bits with those of `a / b` and `1 / b`.  out[0] += mismatches,
// out[1] += pairs that passed the test (so that a test that never passes does not go unnoticed).
__global__ void __launch_bounds__(256)
k_div_check(unsigned long long seed, unsigned long long per_thread, unsigned long long *out) {
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * (blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1ull);
    auto next = [&]() {
        x += 0x9E3779B97F4A7C15ull;
        unsigned long long z = x;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    };
    auto draw = [&](unsigned kind) {
        const unsigned long long r = next();
        const double u = (double)(r >> 11) * 0x1p-53;  // [0, 1)
        const double sgn = (r & 1ull) ? -1.0 : 1.0;
        switch (kind & 7u) {
            case 0: return __longlong_as_double((long long)r);                       // any bit pattern
            case 1: return sgn * (u * 4096.0);                                      // screen coordinates
            case 2: return sgn * (u * 2.0);                                         // clip / NDC magnitudes
            case 3: return sgn * (1.0 + u * 9.0);                                   // w between the planes
            case 4: return sgn * u * 0x1p-1000;                                     // tiny
            case 5: return sgn * (1.0 + u) * 0x1p+1000;                             // huge
            case 6: return (r & 6ull) ? sgn * (double)(long long)(r >> 44) : sgn * 0.0;  // integers and zeros
            default: return sgn * u * 0x1p-1060;                                    // subnormal
        }
    };
    unsigned long long bad = 0, fast = 0;
    for (unsigned long long k = 0; k < per_thread; k++) {
        const unsigned long long sel = next();
        const double a = draw((unsigned)sel), b = draw((unsigned)(sel >> 3));
        bool ok = true;
        const double q = div_tail(a, b, div_refine(b), ok);
        if (ok) {
            const double ref = a / b;
            fast++;
            if (__double_as_longlong(q) != __double_as_longlong(ref) && !(q != q && ref != ref)) bad++;
        }
        ok = true;
        const double r1 = rcp_fast(b, ok);
        if (ok) {
            const double ref = 1 / b;
            fast++;
            if (__double_as_longlong(r1) != __double_as_longlong(ref) && !(r1 != r1 && ref != ref)) bad++;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { bad += __shfl_down_sync(0xffffffffu, bad, o); fast += __shfl_down_sync(0xffffffffu, fast, o); }
    if ((threadIdx.x & 31) == 0) { if (bad) atomicAdd(&out[0], bad); atomicAdd(&out[1], fast); }
}
int launch_div_check(unsigned long long seed, unsigned long long pairs, unsigned long long *out, cudaStream_t st) {
    const unsigned threads = GRID_WAVE * 256u;
    k_div_check<<<GRID_WAVE, 256, 0, st>>>(seed, (pairs + threads - 1) / threads, out);
    return 1;
}

int launch_atomic_probe(unsigned long long *buf, size_t words, unsigned long long ops, cudaStream_t st) {
    const unsigned threads = GRID_WAVE * 256u;
    k_atomic_probe<<<GRID_WAVE, 256, 0, st>>>(buf, words, (ops + threads - 1) / threads);
    return 1;
}

}  // namespace fgl
