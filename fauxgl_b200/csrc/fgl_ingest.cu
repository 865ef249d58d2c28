// fgl_ingest.cu -- data formats either side of the draw, handled on the device so that only
// the compact form crosses PCIe: binary STL records -> planar f64 mesh (stl.go:86-154),
// Mesh.BoundingBox (mesh.go:153-165) and Context.DepthImage (context.go:87-117).
#include "fgl_internal.h"
#include "fgl_math.cuh"

namespace fgl {

// Order-preserving map of an IEEE double onto uint64 (all non-NaN values, -0 < +0), so that
// minima / maxima can be taken with integer atomics.
FGL_DI unsigned long long ord_key(double d) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(d);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
FGL_DI double ord_value(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}
FGL_DI unsigned long long warp_min_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_down_sync(0xffffffffu, v, o);
        v = t < v ? t : v;
    }
    return v;
}
FGL_DI unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long t = __shfl_down_sync(0xffffffffu, v, o);
        v = t > v ? t : v;
    }
    return v;
}

// ---- binary STL -> planes ---------------------------------------------------------------------
// A record is 50 bytes: normal (3 x f32, ignored by the loader), 3 vertices (9 x f32), 2 attribute
// bytes.  A CTA stages 256 records (12 800 contiguous bytes) in shared memory with word loads,
// then one thread per triangle widens the positions (makeFloat, stl.go:82-84) and computes the
// face normal Triangle.Normal(), triangle.go:33-37, which the loader stores in all three vertices.
constexpr int STL_T = 256;
__global__ void __launch_bounds__(STL_T)
k_stl_ingest(const uint8_t *__restrict__ rec, double *__restrict__ pos, double *__restrict__ nrm, uint32_t n) {
    __shared__ uint32_t s_words[STL_T * 50 / 4];
    const uint32_t tid = threadIdx.x;
    for (uint32_t b0 = blockIdx.x * STL_T; b0 < n; b0 += gridDim.x * STL_T) {
        const uint32_t cnt = min((uint32_t)STL_T, n - b0);
        const uint32_t nwords = (cnt * 50 + 3) / 4;  // the staging buffer is padded to a word
        const uint32_t *src = reinterpret_cast<const uint32_t *>(rec + (size_t)b0 * 50);
        __syncthreads();
        for (uint32_t w = tid; w < nwords; w += STL_T) s_words[w] = src[w];
        __syncthreads();
        if (tid < cnt) {
            const uint16_t *h = reinterpret_cast<const uint16_t *>(s_words) + tid * 25 + 6;  // byte 50*tid + 12
            double f[9];
#pragma unroll
            for (int k = 0; k < 9; k++)
                f[k] = (double)__uint_as_float((uint32_t)h[2 * k] | ((uint32_t)h[2 * k + 1] << 16));
            const V3 v1 = v3(f[0], f[1], f[2]), v2 = v3(f[3], f[4], f[5]), v3_ = v3(f[6], f[7], f[8]);
            const V3 nn = v_normalize(v_cross(v_sub(v2, v1), v_sub(v3_, v1)));
            const size_t i = b0 + tid;
#pragma unroll
            for (int v = 0; v < 3; v++) {
                pos[(size_t)(v * 3 + 0) * n + i] = f[v * 3 + 0];
                pos[(size_t)(v * 3 + 1) * n + i] = f[v * 3 + 1];
                pos[(size_t)(v * 3 + 2) * n + i] = f[v * 3 + 2];
                nrm[(size_t)(v * 3 + 0) * n + i] = nn.x;
                nrm[(size_t)(v * 3 + 1) * n + i] = nn.y;
                nrm[(size_t)(v * 3 + 2) * n + i] = nn.z;
            }
        }
    }
}
int launch_stl_ingest(const uint8_t *records, double *pos, double *nrm, uint32_t n, cudaStream_t st) {
    if (n == 0) return 0;
    const uint32_t blocks = (n + STL_T - 1) / STL_T;
    k_stl_ingest<<<blocks < GRID_WAVE ? blocks : GRID_WAVE, STL_T, 0, st>>>(records, pos, nrm, n);
    return 1;
}

// ---- bounding box -----------------------------------------------------------------------------
// bounds[c] = min key of component c, bounds[3 + c] = max key; the caller initialises them to
// ~0 / 0 and decodes with ord_value on the host side of the ABI (same bit trick).
__global__ void k_mesh_bounds(const double *__restrict__ pos, uint32_t n, int nverts, unsigned long long *bounds) {
    unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0ull, 0ull, 0ull};
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        for (int v = 0; v < nverts; v++)
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const double d = pos[(size_t)(v * 3 + c) * n + i];
                if (d != d) continue;
                const unsigned long long k = ord_key(d);
                lo[c] = k < lo[c] ? k : lo[c];
                hi[c] = k > hi[c] ? k : hi[c];
            }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        const unsigned long long a = warp_min_u64(lo[c]), b = warp_max_u64(hi[c]);
        if ((threadIdx.x & 31) == 0) {
            if (a != ~0ull) atomicMin(&bounds[c], a);
            if (b != 0ull) atomicMax(&bounds[3 + c], b);
        }
    }
}
int launch_mesh_bounds(const double *pos, uint32_t n, int nverts, unsigned long long *bounds, cudaStream_t st) {
    if (n == 0) return 0;
    const uint32_t blocks = (n + 255) / 256;
    k_mesh_bounds<<<blocks < GRID_WAVE ? blocks : GRID_WAVE, 256, 0, st>>>(pos, n, nverts, bounds);
    return 1;
}

// ---- DepthImage, context.go:87-117 --------------------------------------------------------------
constexpr double GO_MAX_FLOAT64 = 1.7976931348623157e308;
__global__ void k_depth_range(const double *__restrict__ depth, size_t npix, unsigned long long *range) {
    unsigned long long lo = ~0ull, hi = 0ull;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const double d = depth[i];
        if (d == GO_MAX_FLOAT64 || d != d) continue;  // :91-93; a NaN never passes `d < lo` / `d > hi`
        const unsigned long long k = ord_key(d);
        lo = k < lo ? k : lo;
        hi = k > hi ? k : hi;
    }
    lo = warp_min_u64(lo);
    hi = warp_max_u64(hi);
    if ((threadIdx.x & 31) == 0) {
        if (lo != ~0ull) atomicMin(&range[0], lo);
        if (hi != 0ull) atomicMax(&range[1], hi);
    }
}
__global__ void k_depth_gray16(const double *__restrict__ depth, size_t npix, const unsigned long long *range,
                               uint16_t *__restrict__ out) {
    const double lo = range[0] == ~0ull ? GO_MAX_FLOAT64 : ord_value(range[0]);   // :88
    const double hi = range[1] == 0ull ? -GO_MAX_FLOAT64 : ord_value(range[1]);   // :89
    const double span = hi - lo;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < npix; i += (size_t)gridDim.x * blockDim.x) {
        const double d = depth[i];
        double t = (d - lo) / span;               // :107
        if (d == GO_MAX_FLOAT64) t = 1;           // :108-110
        out[i] = (uint16_t)(go_int(t * 65535.0) & 0xffff);  // uint16(t * 0xffff), :111
    }
}
int launch_depth_image(const double *depth, size_t npix, uint16_t *out, unsigned long long *scratch, cudaStream_t st) {
    static const unsigned long long init[2] = {~0ull, 0ull};
    cudaMemcpyAsync(scratch, init, sizeof init, cudaMemcpyHostToDevice, st);
    k_depth_range<<<GRID_WAVE, 256, 0, st>>>(depth, npix, scratch);
    k_depth_gray16<<<GRID_WAVE, 256, 0, st>>>(depth, npix, scratch, out);
    return 2;
}

// ---- indexed (OBJ) mesh -> planes, obj.go:58-74 --------------------------------------------------
// One thread per triangle: gather the three corners from the tables, then Triangle.FixNormals (triangle.go:46-58):
// a corner whose normal == Vector{} (so -0 counts) takes the face normal Triangle.Normal() (triangle.go:33-37).
__global__ void __launch_bounds__(256)
k_indexed_ingest(const double *__restrict__ tv, const double *__restrict__ tvt, const double *__restrict__ tvn,
                 const int32_t *__restrict__ corners, double *__restrict__ pos, double *__restrict__ nrm,
                 double *__restrict__ tex, uint32_t n) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        V3 p[3], nn[3];
        double uv[3][2];
#pragma unroll
        for (int v = 0; v < 3; v++) {
            const int32_t *c = corners + ((size_t)i * 3 + v) * 3;
            const double *a = tv + (size_t)c[0] * 3, *b = tvt + (size_t)c[1] * 3, *d = tvn + (size_t)c[2] * 3;
            p[v] = v3(a[0], a[1], a[2]);
            uv[v][0] = tex ? b[0] : 0.0; uv[v][1] = tex ? b[1] : 0.0;
            nn[v] = v3(d[0], d[1], d[2]);
        }
        const V3 face = v_normalize(v_cross(v_sub(p[1], p[0]), v_sub(p[2], p[0])));
#pragma unroll
        for (int v = 0; v < 3; v++) {
            if (nn[v].x == 0 && nn[v].y == 0 && nn[v].z == 0) nn[v] = face;
            pos[(size_t)(v * 3 + 0) * n + i] = p[v].x;
            pos[(size_t)(v * 3 + 1) * n + i] = p[v].y;
            pos[(size_t)(v * 3 + 2) * n + i] = p[v].z;
            nrm[(size_t)(v * 3 + 0) * n + i] = nn[v].x;
            nrm[(size_t)(v * 3 + 1) * n + i] = nn[v].y;
            nrm[(size_t)(v * 3 + 2) * n + i] = nn[v].z;
            if (tex) {  // (null: the texture planes are up to date)
                tex[(size_t)(v * 2 + 0) * n + i] = uv[v][0];
                tex[(size_t)(v * 2 + 1) * n + i] = uv[v][1];
            }
        }
    }
}
int launch_indexed_ingest(const double *v, const double *vt, const double *vn, const int32_t *corners, double *pos,
                          double *nrm, double *tex, uint32_t n, cudaStream_t st) {
    if (n == 0) return 0;
    const uint32_t blocks = (n + 255) / 256;
    k_indexed_ingest<<<blocks < GRID_WAVE ? blocks : GRID_WAVE, 256, 0, st>>>(v, vt, vn, corners, pos, nrm, tex, n);
    return 1;
}

// ---- Mesh.SmoothNormals, mesh.go:105-120 ----------------------------------------------------------
// Corner c = 3 t + v (the order the reference's loops visit them: t0.V1, t0.V2, t0.V3, t1.V1, ...).
FGL_DI V3 corner_position(const double *__restrict__ pos, uint32_t n, uint32_t c) {
    const uint32_t t = c / 3u, v = c % 3u;
    // (+ 0.0: -0 -> +0, the two are the same map key in Go)
    return v3(pos[(size_t)(v * 3 + 0) * n + t] + 0.0, pos[(size_t)(v * 3 + 1) * n + t] + 0.0, pos[(size_t)(v * 3 + 2) * n + t] + 0.0);
}
FGL_DI bool same_position(V3 a, V3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }  // NaN: never
__global__ void __launch_bounds__(256)
k_corner_hash(const double *__restrict__ pos, uint32_t n, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals) {
    const uint32_t nc = 3u * n;
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        const V3 p = corner_position(pos, n, c);
        unsigned long long h = 0x9E3779B97F4A7C15ull;
        const unsigned long long w[3] = {(unsigned long long)__double_as_longlong(p.x), (unsigned long long)__double_as_longlong(p.y),
                                         (unsigned long long)__double_as_longlong(p.z)};
#pragma unroll
        for (int k = 0; k < 3; k++) {  // splitmix64-style mixing of the three bit patterns
            h ^= w[k] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
            h = (h ^ (h >> 30)) * 0xBF58476D1CE4E5B9ull;
            h = (h ^ (h >> 27)) * 0x94D049BB133111EBull;
            h ^= h >> 31;
        }
        keys[c] = (uint32_t)(h >> 32);
        vals[c] = c;
    }
}
// One thread per sorted pair.  The pairs of one position are contiguous up to hash collisions and, the sort being
// stable, in corner order.  (Go map semantics for the key: +0 == -0; a NaN component makes the key unfindable.)
// The first corner of a position (no equal position earlier in its hash run) is the
// group's leader: it adds the members' normals in order, starting from zero like `lookup[k].Add(n)` on an empty map
// entry, normalises (vector.go:83-86) and stores the result to every member.  A corner's normal is read and written
// by its own leader only, so the update is in place.
__global__ void __launch_bounds__(256)
k_smooth_groups(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, const double *__restrict__ pos,
                double *__restrict__ nrm, uint32_t n) {
    const uint32_t nc = 3u * n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
        const uint32_t key = keys[i];
        const V3 p = corner_position(pos, n, vals[i]);
        if (p.x != p.x || p.y != p.y || p.z != p.z) {
            // a NaN position is a map key that is never found again: lookup[t.V1.Position] yields the zero Vector
            const uint32_t c = vals[i], t = c / 3u, v = c % 3u;
            nrm[(size_t)(v * 3 + 0) * n + t] = 0; nrm[(size_t)(v * 3 + 1) * n + t] = 0; nrm[(size_t)(v * 3 + 2) * n + t] = 0;
            continue;
        }
        bool leader = true;
        for (uint32_t j = i; j > 0 && keys[j - 1] == key; j--)
            if (same_position(corner_position(pos, n, vals[j - 1]), p)) { leader = false; break; }
        if (!leader) continue;
        V3 acc = v3(0, 0, 0);
        for (uint32_t j = i; j < nc && keys[j] == key; j++) {
            const uint32_t c = vals[j];
            if (j != i && !same_position(corner_position(pos, n, c), p)) continue;
            const uint32_t t = c / 3u, v = c % 3u;
            acc = v_add(acc, v3(nrm[(size_t)(v * 3 + 0) * n + t], nrm[(size_t)(v * 3 + 1) * n + t], nrm[(size_t)(v * 3 + 2) * n + t]));
        }
        const V3 r = v_normalize(acc);
        for (uint32_t j = i; j < nc && keys[j] == key; j++) {
            const uint32_t c = vals[j];
            if (j != i && !same_position(corner_position(pos, n, c), p)) continue;
            const uint32_t t = c / 3u, v = c % 3u;
            nrm[(size_t)(v * 3 + 0) * n + t] = r.x;
            nrm[(size_t)(v * 3 + 1) * n + t] = r.y;
            nrm[(size_t)(v * 3 + 2) * n + t] = r.z;
        }
    }
}
// Mesh.SmoothNormalsThreshold, mesh.go:80-103: one thread per sorted pair = per corner; it walks its hash run from
// the start (corner order) and adds the normals at its own position that pass the filter against ITS normal.
__global__ void __launch_bounds__(256)
k_smooth_threshold(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, const double *__restrict__ pos,
                   const double *__restrict__ nin, double *__restrict__ nout, uint32_t n, double threshold) {
    const uint32_t nc = 3u * n;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) {
        const uint32_t key = keys[i], ci = vals[i], ti = ci / 3u, vi = ci % 3u;
        const V3 p = corner_position(pos, n, ci);
        const V3 mine = v3(nin[(size_t)(vi * 3 + 0) * n + ti], nin[(size_t)(vi * 3 + 1) * n + ti], nin[(size_t)(vi * 3 + 2) * n + ti]);
        V3 acc = v3(0, 0, 0);
        if (p.x == p.x && p.y == p.y && p.z == p.z) {  // (a NaN key is never found again: the list is empty)
            uint32_t j = i;
            while (j > 0 && keys[j - 1] == key) j--;
            for (; j < nc && keys[j] == key; j++) {
                const uint32_t c = vals[j], t = c / 3u, v = c % 3u;
                if (j != i && !same_position(corner_position(pos, n, c), p)) continue;
                const V3 x = v3(nin[(size_t)(v * 3 + 0) * n + t], nin[(size_t)(v * 3 + 1) * n + t], nin[(size_t)(v * 3 + 2) * n + t]);
                if (v_dot(x, mine) >= threshold) acc = v_add(acc, x);  // mesh.go:82-86
            }
        }
        const V3 r = v_normalize(acc);
        nout[(size_t)(vi * 3 + 0) * n + ti] = r.x;
        nout[(size_t)(vi * 3 + 1) * n + ti] = r.y;
        nout[(size_t)(vi * 3 + 2) * n + ti] = r.z;
    }
}
int launch_smooth_threshold(const uint32_t *keys, const uint32_t *vals, const double *pos, const double *nrm_in,
                            double *nrm_out, uint32_t n, double threshold, cudaStream_t st) {
    if (n == 0) return 0;
    k_smooth_threshold<<<GRID_WAVE, 256, 0, st>>>(keys, vals, pos, nrm_in, nrm_out, n, threshold);
    return 1;
}
int launch_corner_hash(const double *pos, uint32_t n, uint32_t *keys, uint32_t *vals, cudaStream_t st) {
    if (n == 0) return 0;
    k_corner_hash<<<GRID_WAVE, 256, 0, st>>>(pos, n, keys, vals);
    return 1;
}
int launch_smooth_groups(const uint32_t *keys, const uint32_t *vals, const double *pos, double *nrm, uint32_t n,
                         cudaStream_t st) {
    if (n == 0) return 0;
    k_smooth_groups<<<GRID_WAVE, 256, 0, st>>>(keys, vals, pos, nrm, n);
    return 1;
}

}  // namespace fgl
