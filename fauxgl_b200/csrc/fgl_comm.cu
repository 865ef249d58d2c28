// fgl_comm.cu -- sort-last depth composite across the GPUs of one box, inside the library (the
// reference's only analogue is the channel reduce of DrawTriangles, context.go:393-410,413-433).  Rank r has drawn
// its triangle range into its own full-frame buffers; afterwards the presenting rank (or every rank) holds, per
// pixel, the fragment with the smallest depth.  Valid for the order-independent state only (ReadDepth, WriteDepth,
// DepthBias 0, opaque output).  Two implementations behind the C ABI, both stream-ordered and host-sync free:
//
//  1. fgl_comm_* / fgl_composite -- the north star's scheme: packed keys (depth32 << 32 | rgba8), min-reduced with
//     NCCL over NVLink.  Not an all-reduce: a REDUCE-SCATTER by screen stripe (ncclMin on ncclUint64: every rank
//     ends up with the composited keys of 1/N of the frame, (N-1)/N * W*H*8 bytes in and out per GPU), then either
//     a gather of the stripes to the presenting rank (grouped ncclSend/ncclRecv) or, if every rank needs the frame,
//     an all-gather.  NCCL is loaded with dlopen when the first communicator is created, so the library itself has
//     no link-time dependency on it (a single-GPU host needs no NCCL); without it fgl_comm_init fails loudly.
//
//  2. fgl_peer_* -- peer memory: ONE kernel per rank does compare + exchange with P2P loads/stores through
//     NVLink/NVSwitch, on float64 depth (no quantisation; ties go to the higher rank = the later triangle range, the
//     reference's `<=` rule), so the result equals the single-GPU frame bit for bit.  It is SPARSE: every context
//     keeps a bitmap of the strips whose depth was written since the last depth clear (k_strip sets it), rank r
//     composites the scanlines y = r (mod N) and only reads, from each peer, the strips that peer has drawn into --
//     a 10 M-triangle sphere that covers 59 % of an 8K frame moves ~12 B per COVERED pixel instead of 8 B per
//     pixel of the frame per rank.  Ranks synchronise through flags in peer memory (release stores / acquire loads at
//     system scope), not through host barriers: signal(ready) -> wait(all ready) -> composite -> signal(done) ->
//     wait(all done), all enqueued on the context's stream.
#include <dlfcn.h>
#include <unistd.h>

#include <cstring>
#include <new>

#include "fgl_ctx.h"

using namespace fgl;

// ---------------------------------------------------------------------------------------------------------
// NCCL through dlopen: the few entry points used, with the types of nccl.h (2.x ABI)
// ---------------------------------------------------------------------------------------------------------
namespace {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
static_assert(sizeof(ncclUniqueId) == FGL_COMM_ID_BYTES, "NCCL unique id size");
enum { ncclSuccess = 0 };
enum { ncclMin = 3 };      // ncclRedOp_t
enum { ncclUint64 = 5 };   // ncclDataType_t

struct NcclApi {
    void *lib;
    int (*GetUniqueId)(ncclUniqueId *);
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
    int (*CommDestroy)(ncclComm_t);
    int (*ReduceScatter)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
    int (*GroupStart)();
    int (*GroupEnd)();
    const char *(*GetErrorString)(int);
    int (*GetVersion)(int *);
};
NcclApi g_nccl{};
std::mutex g_nccl_mu;

// Returns nullptr and a message when NCCL cannot be loaded.
const NcclApi *nccl_api(const char **why) {
    std::lock_guard<std::mutex> g(g_nccl_mu);
    if (g_nccl.lib) return &g_nccl;
    const char *names[] = {getenv("FGL_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    void *lib = nullptr;
    for (const char *n : names) {
        if (!n || !*n) continue;
        lib = dlopen(n, RTLD_NOW | RTLD_LOCAL);
        if (lib) break;
    }
    if (!lib) { *why = "libnccl.so.2 not found (set FGL_NCCL_LIB)"; return nullptr; }
    NcclApi a{};
    a.lib = lib;
#define SYM(field, name) a.field = reinterpret_cast<decltype(a.field)>(dlsym(lib, name)); if (!a.field) { *why = "NCCL symbol missing: " name; dlclose(lib); return nullptr; }
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(ReduceScatter, "ncclReduceScatter")
    SYM(AllGather, "ncclAllGather")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(GetErrorString, "ncclGetErrorString")
    SYM(GetVersion, "ncclGetVersion")
#undef SYM
    g_nccl = a;
    return &g_nccl;
}

}  // namespace

constexpr int STAGE_EVENTS = 5;
struct StageTimer {  // per-stage device times of the composites issued while the context's profiling is on
    cudaEvent_t ev[STAGE_EVENTS];
    bool created, pending;
    float acc[STAGE_EVENTS - 1];
    unsigned n;
    void create() {
        if (created) return;
        for (auto &e : ev) cudaEventCreate(&e);
        created = true;
    }
    void destroy() {
        if (created) for (auto &e : ev) cudaEventDestroy(e);
        created = false;
    }
    void drain() {  // (the events of the last composite, once its stream has passed them)
        if (!pending) return;
        cudaEventSynchronize(ev[STAGE_EVENTS - 1]);
        for (int k = 0; k + 1 < STAGE_EVENTS; k++) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[k], ev[k + 1]);
            acc[k] += ms;
        }
        n++;
        pending = false;
    }
};

struct fgl_comm {
    int device, nranks, rank;
    ncclComm_t comm;
    unsigned long long *keys;  // [chunk * nranks] packed keys of the whole frame (padded to a multiple of nranks)
    size_t chunk;              // keys per stripe
    size_t npix;
    StageTimer timer;          // pack | reduce-scatter | gather / all-gather | unpack
};

#define NCK(c, expr)                                                                                      \
    do {                                                                                                  \
        int _r = (expr);                                                                                  \
        if (_r != ncclSuccess) return api_fail(c, FGL_E_CUDA, "%s: %s", #expr, api->GetErrorString(_r)); \
    } while (0)
#define CCK(c, expr)                                                                                       \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) {                                                                           \
            cudaGetLastError();                                                                            \
            return api_fail(c, _e == cudaErrorMemoryAllocation ? FGL_E_OOM : FGL_E_CUDA, "%s: %s", #expr, \
                            cudaGetErrorString(_e));                                                       \
        }                                                                                                  \
    } while (0)

// ---------------------------------------------------------------------------------------------------------
// peer groups
// ---------------------------------------------------------------------------------------------------------
namespace fgl {

constexpr int FLAG_READY = 0, FLAG_DONE = FGL_MAX_PEERS, FLAG_ERROR = 2 * FGL_MAX_PEERS, FLAG_WORDS = 2 * FGL_MAX_PEERS + 8;

struct PeerSet {
    uint32_t *color[FGL_MAX_PEERS];
    double *depth[FGL_MAX_PEERS];
    uint8_t *dirty[FGL_MAX_PEERS];
    unsigned long long *flags[FGL_MAX_PEERS];
};

__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// Thread p tells rank p that this rank has reached `epoch` (everything this stream did before -- draws, or the
// composite's loads and stores -- is ordered before the flag by the fence + release).
__global__ void k_peer_signal(const PeerSet P, int rank, int nranks, int slot, unsigned long long epoch) {
    const int p = threadIdx.x;
    if (p >= nranks) return;
    __threadfence_system();
    st_release_sys(P.flags[p] + slot + rank, epoch);
}
// Thread p waits until rank p has reached `epoch` (flags live in THIS rank's memory: local polling).  Bounded: a
// peer that never arrives sets the error word instead of hanging the GPU.
__global__ void k_peer_wait(unsigned long long *myflags, int nranks, int slot, unsigned long long epoch,
                            unsigned long long timeout_ns) {
    const int p = threadIdx.x;
    if (p >= nranks) return;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(myflags + slot + p) < epoch) {
        if (global_ns() - t0 > timeout_ns) {
            atomicOr(myflags + FLAG_ERROR, 1ull);
            break;
        }
        __nanosleep(256);
    }
}

// Copy every peer's dirty-strip bitmap into local memory (16-byte P2P loads; the bitmaps are padded by 16 bytes).
__global__ void k_peer_gather_dirty(const PeerSet P, int nranks, uint32_t ntiles, uint8_t *__restrict__ local_all) {
    const uint32_t n16 = (ntiles + 15u) / 16u;
    for (int p = blockIdx.y; p < nranks; p += gridDim.y) {
        const uint4 *src = reinterpret_cast<const uint4 *>(P.dirty[p]);
        uint4 *dst = reinterpret_cast<uint4 *>(local_all + (size_t)p * n16 * 16u);
        for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += gridDim.x * blockDim.x) dst[i] = __ldcv(src + i);
    }
}

// The sparse composite.  One warp per strip of the scanlines this rank owns (y = rank mod nranks).  Candidates: with
// a presenting rank `root`, the ranks whose bitmap has the strip -- a strip only root has drawn needs nothing, and an
// undrawn strip of root's holds its clear value everywhere, which no drawn fragment of a peer can lose against
// (it passed `z <= clear` on its own rank; all ranks clear with the same value), so root's buffer is not even read
// there, and a pixel that still holds the clear depth AND the clear colour (`clear_z`, `clear_c`; NaN / unknown after
// an upload: then everything is stored) is not stored -- root's own pixel is identical; with root < 0 all
// ranks are candidates (every rank receives the result, so every rank's own pixel takes part).  Per pixel: the
// candidates' float64 depths are loaded (all loads issued before the first compare), the smallest wins -- on a tie
// the higher rank --, only the winner's colour is fetched, and depth + colour go to the target(s) that do not
// already hold them.
__global__ void __launch_bounds__(256)
k_peer_composite(const PeerSet P, int rank, int nranks, int root, int width, int height, int tile_w, int tiles_x,
                 const uint8_t *__restrict__ dirty_all, uint32_t dirty_stride, double clear_z, uint32_t clear_c,
                 int clear_known, int color_only) {
    const int lane = threadIdx.x & 31;
    const uint32_t warps = gridDim.x * (blockDim.x >> 5), gw = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const uint32_t my_rows = height > rank ? (uint32_t)((height - rank + nranks - 1) / nranks) : 0u;
    const uint32_t ntask = my_rows * (uint32_t)tiles_x;
    const int ppl = tile_w >> 5;  // pixels per lane: 1 or 2
    for (uint32_t t = gw; t < ntask; t += warps) {
        const int y = rank + (int)(t / (uint32_t)tiles_x) * nranks, col = (int)(t % (uint32_t)tiles_x);
        const uint32_t strip = (uint32_t)y * (uint32_t)tiles_x + (uint32_t)col;
        // who has drawn into this strip
        const uint32_t drawn = __ballot_sync(0xffffffffu, lane < nranks && dirty_all[(size_t)lane * dirty_stride + strip] != 0);
        if (drawn == 0) continue;  // nobody: every rank keeps its own (cleared) pixels
        if (root >= 0 && drawn == (1u << root)) continue;  // only the presenting rank: it already holds the result
        const uint32_t all = nranks >= 32 ? 0xffffffffu : ((1u << nranks) - 1u);
        const uint32_t cand = root >= 0 ? drawn : all;
        const uint32_t targets = root >= 0 ? (1u << root) : all;
        const bool root_undrawn = root >= 0 && !((drawn >> root) & 1u);
        const int x0 = col * tile_w;
        // Both pixels of the lane at once, and every load issued before its first use: a strip costs two dependent
        // round trips over NVLink (depths, then the winner's colour) -- one when a single rank drew it, because the
        // winner is then known without comparing.
        size_t idx[2];
        bool live[2];
        double d[2][FGL_MAX_PEERS];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int x = x0 + lane + 32 * h;
            live[h] = h < ppl && x < width;
            idx[h] = (size_t)y * width + (live[h] ? x : x0);
        }
        const bool single = (cand & (cand - 1u)) == 0u;
        const int only = __ffs(cand) - 1;
        uint32_t cpre[2] = {0u, 0u};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (!live[h]) continue;
#pragma unroll
            for (int r = 0; r < FGL_MAX_PEERS; r++)
                if ((cand >> r) & 1u) d[h][r] = __ldcv(P.depth[r] + idx[h]);
            if (single) cpre[h] = __ldcv(P.color[only] + idx[h]);
        }
        double best[2] = {0, 0};
        int win[2] = {-1, -1};
        bool need[2] = {false, false};
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (!live[h]) continue;
#pragma unroll
            for (int r = 0; r < FGL_MAX_PEERS; r++)
                if (((cand >> r) & 1u) && (win[h] < 0 || d[h][r] <= best[h])) { best[h] = d[h][r]; win[h] = r; }
            // (a target that is the winner, or that ties with a LOWER-ranked winner, already holds the result)
#pragma unroll
            for (int r = 0; r < FGL_MAX_PEERS; r++)
                if (((targets >> r) & 1u) && r != win[h]) need[h] = true;
        }
        uint32_t cwin[2];
#pragma unroll
        for (int h = 0; h < 2; h++) cwin[h] = (need[h] && !single) ? __ldcv(P.color[win[h]] + idx[h]) : cpre[h];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            if (!need[h]) continue;
            const uint32_t c = cwin[h];
            if (root_undrawn && clear_known && best[h] == clear_z && c == clear_c) continue;  // root's own cleared pixel is identical
#pragma unroll
            for (int r = 0; r < FGL_MAX_PEERS; r++)
                if (((targets >> r) & 1u) && r != win[h]) {
                    if (!color_only && (!((cand >> r) & 1u) || d[h][r] != best[h])) P.depth[r][idx[h]] = best[h];
                    P.color[r][idx[h]] = c;
                }
        }
        // the targets' buffers now hold drawn depth in this strip
        if (!color_only && lane < nranks && ((targets >> lane) & 1u) && !((drawn >> lane) & 1u)) P.dirty[lane][strip] = 1;
    }
}

}  // namespace fgl

struct fgl_peer_group {
    int device, rank, nranks;
    bool ipc;  // peers' pointers were opened with cudaIpcOpenMemHandle (closed on destroy)
    fgl::PeerSet set;
    unsigned long long *flags;  // this rank's flag block (FLAG_WORDS words)
    uint8_t *dirty_all;         // [nranks][dirty_stride] gathered copies of the bitmaps
    uint32_t dirty_stride;
    unsigned long long epoch;
    unsigned long long *host_err;  // pinned: the error word read back by fgl_peer_status
    fgl_ctx *ctx;
    StageTimer timer;              // signal + wait for all ranks' draws | bitmaps + composite kernel | signal + wait done | -
};

namespace {

struct PeerExport {  // what fgl_peer_export writes: FGL_PEER_EXPORT_BYTES
    cudaIpcMemHandle_t color, depth, dirty, flags;
    int width, height, tile_w, device;
    unsigned long long local_ptrs[4];  // the raw device pointers, for groups inside one process
    int pid, _pad;
    unsigned long long epoch;  // the context's last composite epoch: a new group starts above every member's
};
static_assert(sizeof(PeerExport) <= FGL_PEER_EXPORT_BYTES, "peer export record");

int ensure_flags(fgl_ctx *c) {
    if (c->peer_flags) return FGL_OK;
    CCK(c, cudaMalloc(reinterpret_cast<void **>(&c->peer_flags), sizeof(unsigned long long) * fgl::FLAG_WORDS));
    CCK(c, cudaMemset(c->peer_flags, 0, sizeof(unsigned long long) * fgl::FLAG_WORDS));
    return FGL_OK;
}

}  // namespace

extern "C" {

// ---- NCCL communicator ----------------------------------------------------------------------------------------

int fgl_comm_unique_id(void *id_out) {
    if (!id_out) return api_fail(nullptr, FGL_E_INVALID, "null id buffer");
    const char *why = "";
    const NcclApi *api = nccl_api(&why);
    if (!api) return api_fail(nullptr, FGL_E_UNSUPPORTED, "NCCL unavailable: %s", why);
    ncclUniqueId id;
    NCK(nullptr, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof id);
    return FGL_OK;
}

int fgl_comm_init(fgl_ctx *c, int nranks, int rank, const void *id, fgl_comm **out) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!out || !id || nranks < 1 || rank < 0 || rank >= nranks) return api_fail(c, FGL_E_INVALID, "bad communicator arguments");
    *out = nullptr;
    const char *why = "";
    const NcclApi *api = nccl_api(&why);
    if (!api) return api_fail(c, FGL_E_UNSUPPORTED, "NCCL unavailable: %s", why);
    fgl_comm *m = new (std::nothrow) fgl_comm();
    if (!m) return api_fail(c, FGL_E_OOM, "host allocation failed");
    m->device = c->device; m->nranks = nranks; m->rank = rank; m->comm = nullptr; m->keys = nullptr;
    memset(&m->timer, 0, sizeof m->timer);
    m->npix = (size_t)c->w * c->h;
    m->chunk = (m->npix + nranks - 1) / nranks;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&m->keys), sizeof(unsigned long long) * m->chunk * nranks);
    if (e != cudaSuccess) { cudaGetLastError(); delete m; return api_fail(c, FGL_E_OOM, "key buffer: %s", cudaGetErrorString(e)); }
    // the padding behind the last pixel never wins a min: all ones
    cudaMemset(m->keys, 0xff, sizeof(unsigned long long) * m->chunk * nranks);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof uid);
    int r = api->CommInitRank(&m->comm, nranks, uid, rank);
    if (r != ncclSuccess) {
        cudaFree(m->keys);
        delete m;
        return api_fail(c, FGL_E_CUDA, "ncclCommInitRank: %s", api->GetErrorString(r));
    }
    *out = m;
    return FGL_OK;
}

int fgl_comm_destroy(fgl_comm *m) {
    if (!m) return FGL_OK;
    cudaSetDevice(m->device);
    const char *why = "";
    const NcclApi *api = nccl_api(&why);
    if (api && m->comm) api->CommDestroy(m->comm);
    if (m->keys) cudaFree(m->keys);
    m->timer.destroy();
    delete m;
    return FGL_OK;
}

int fgl_composite(fgl_ctx *c, fgl_comm *m, int root) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!m) return api_fail(c, FGL_E_INVALID, "null communicator");
    if (m->device != c->device || m->npix != (size_t)c->w * c->h) return api_fail(c, FGL_E_INVALID, "communicator belongs to another context shape/device");
    if (root >= m->nranks) return api_fail(c, FGL_E_INVALID, "root %d outside the %d ranks", root, m->nranks);
    const char *why = "";
    const NcclApi *api = nccl_api(&why);
    if (!api) return api_fail(c, FGL_E_UNSUPPORTED, "NCCL unavailable: %s", why);
    std::lock_guard<std::mutex> lock(c->mu);
    api_fb_join(c);
    const size_t npix = m->npix, chunk = m->chunk;
    StageTimer *tm = c->profiling ? &m->timer : nullptr;
    if (tm) { tm->drain(); tm->create(); cudaEventRecord(tm->ev[0], c->stream); }
    launch_composite_pack(c->color, c->depth, m->keys, npix, c->stream, /*bias=*/false);
    if (tm) cudaEventRecord(tm->ev[1], c->stream);
    if (m->nranks > 1) {
        // stripe s of the frame is reduced onto rank s (in place: recvbuff = sendbuff + rank * chunk)
        NCK(c, api->ReduceScatter(m->keys, m->keys + (size_t)m->rank * chunk, chunk, ncclUint64, ncclMin, m->comm, c->stream));
        if (tm) cudaEventRecord(tm->ev[2], c->stream);
        if (root < 0) {
            NCK(c, api->AllGather(m->keys + (size_t)m->rank * chunk, m->keys, chunk, ncclUint64, m->comm, c->stream));
        } else {
            NCK(c, api->GroupStart());
            if (m->rank == root) {
                for (int r = 0; r < m->nranks; r++)
                    if (r != root) NCK(c, api->Recv(m->keys + (size_t)r * chunk, chunk, ncclUint64, r, m->comm, c->stream));
            } else {
                NCK(c, api->Send(m->keys + (size_t)m->rank * chunk, chunk, ncclUint64, root, m->comm, c->stream));
            }
            NCK(c, api->GroupEnd());
        }
    } else if (tm) {
        cudaEventRecord(tm->ev[2], c->stream);
    }
    if (tm) cudaEventRecord(tm->ev[3], c->stream);
    if (root < 0 || m->rank == root) {
        launch_composite_unpack(c->color, c->depth, m->keys, npix, c->stream, /*bias=*/false);
        cudaMemsetAsync(c->wb.dirty, 1, c->wb.ntiles, c->stream);
    }
    if (tm) { cudaEventRecord(tm->ev[4], c->stream); tm->pending = true; }
    CCK(c, cudaGetLastError());
    return FGL_OK;
}

int fgl_comm_stage_times(fgl_ctx *c, fgl_comm *m, float ms[4], uint32_t *composites) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!m || !ms) return api_fail(c, FGL_E_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    m->timer.drain();
    for (int k = 0; k < 4; k++) { ms[k] = m->timer.acc[k]; m->timer.acc[k] = 0; }
    if (composites) *composites = m->timer.n;
    m->timer.n = 0;
    return FGL_OK;
}

// ---- peer groups ------------------------------------------------------------------------------------------------

int fgl_peer_export(fgl_ctx *c, void *record) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!record) return api_fail(c, FGL_E_INVALID, "null export record");
    std::lock_guard<std::mutex> lock(c->mu);
    rc = ensure_flags(c);
    if (rc) return rc;
    PeerExport e;
    memset(&e, 0, sizeof e);
    CCK(c, cudaIpcGetMemHandle(&e.color, c->color));
    CCK(c, cudaIpcGetMemHandle(&e.depth, c->depth));
    CCK(c, cudaIpcGetMemHandle(&e.dirty, c->wb.dirty));
    CCK(c, cudaIpcGetMemHandle(&e.flags, c->peer_flags));
    e.width = c->w; e.height = c->h; e.tile_w = c->tile_w; e.device = c->device;
    e.local_ptrs[0] = (unsigned long long)c->color; e.local_ptrs[1] = (unsigned long long)c->depth;
    e.local_ptrs[2] = (unsigned long long)c->wb.dirty; e.local_ptrs[3] = (unsigned long long)c->peer_flags;
    e.pid = (int)getpid();
    e.epoch = c->peer_epoch;
    memset(record, 0, FGL_PEER_EXPORT_BYTES);
    memcpy(record, &e, sizeof e);
    return FGL_OK;
}

int fgl_peer_group_create(fgl_ctx *c, int rank, int nranks, const void *records, fgl_peer_group **out) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!out || !records || nranks < 1 || nranks > FGL_MAX_PEERS || rank < 0 || rank >= nranks)
        return api_fail(c, FGL_E_INVALID, "bad rank %d / nranks %d (at most %d peers)", rank, nranks, FGL_MAX_PEERS);
    *out = nullptr;
    std::lock_guard<std::mutex> lock(c->mu);
    rc = ensure_flags(c);
    if (rc) return rc;
    fgl_peer_group *g = new (std::nothrow) fgl_peer_group();
    if (!g) return api_fail(c, FGL_E_OOM, "host allocation failed");
    memset(g, 0, sizeof *g);
    g->device = c->device; g->rank = rank; g->nranks = nranks; g->ctx = c; g->flags = c->peer_flags;
    g->epoch = c->peer_epoch;  // flags only grow: the group starts above the last epoch of every member (below)
    const int mypid = (int)getpid();
    for (int r = 0; r < nranks; r++) {
        PeerExport e;
        memcpy(&e, static_cast<const char *>(records) + (size_t)r * FGL_PEER_EXPORT_BYTES, sizeof e);
        if (e.epoch > g->epoch) g->epoch = e.epoch;
        if (e.width != c->w || e.height != c->h || e.tile_w != c->tile_w) {
            fgl_peer_group_destroy(g);
            return api_fail(c, FGL_E_INVALID, "rank %d renders %dx%d (strips of %d), this rank %dx%d (%d)", r, e.width, e.height,
                            e.tile_w, c->w, c->h, c->tile_w);
        }
        if (r == rank) {
            g->set.color[r] = c->color; g->set.depth[r] = c->depth; g->set.dirty[r] = c->wb.dirty; g->set.flags[r] = c->peer_flags;
            continue;
        }
        if (e.pid == mypid) {
            // same process (one host thread per GPU, or several contexts on one GPU in the tests): plain pointers;
            // across devices peer access has to be switched on once
            if (e.device != c->device) {
                cudaError_t pe = cudaDeviceEnablePeerAccess(e.device, 0);
                if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) {
                    cudaGetLastError();
                    fgl_peer_group_destroy(g);
                    return api_fail(c, FGL_E_CUDA, "no peer access from device %d to %d: %s", c->device, e.device, cudaGetErrorString(pe));
                }
                cudaGetLastError();
            }
            g->set.color[r] = (uint32_t *)e.local_ptrs[0]; g->set.depth[r] = (double *)e.local_ptrs[1];
            g->set.dirty[r] = (uint8_t *)e.local_ptrs[2]; g->set.flags[r] = (unsigned long long *)e.local_ptrs[3];
        } else {
            g->ipc = true;
            void *p[4] = {nullptr, nullptr, nullptr, nullptr};
            const cudaIpcMemHandle_t *h[4] = {&e.color, &e.depth, &e.dirty, &e.flags};
            for (int k = 0; k < 4; k++) {
                cudaError_t oe = cudaIpcOpenMemHandle(&p[k], *h[k], cudaIpcMemLazyEnablePeerAccess);
                if (oe != cudaSuccess) {
                    cudaGetLastError();
                    g->set.color[r] = (uint32_t *)p[0]; g->set.depth[r] = (double *)p[1]; g->set.dirty[r] = (uint8_t *)p[2];
                    fgl_peer_group_destroy(g);
                    return api_fail(c, FGL_E_CUDA, "cudaIpcOpenMemHandle (rank %d): %s", r, cudaGetErrorString(oe));
                }
            }
            g->set.color[r] = (uint32_t *)p[0]; g->set.depth[r] = (double *)p[1]; g->set.dirty[r] = (uint8_t *)p[2];
            g->set.flags[r] = (unsigned long long *)p[3];
        }
    }
    g->dirty_stride = (c->wb.ntiles + 15u) / 16u * 16u;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&g->dirty_all), (size_t)g->dirty_stride * nranks);
    if (e == cudaSuccess) e = cudaMallocHost(reinterpret_cast<void **>(&g->host_err), sizeof(unsigned long long));
    if (e != cudaSuccess) {
        cudaGetLastError();
        fgl_peer_group_destroy(g);
        return api_fail(c, FGL_E_OOM, "peer group buffers: %s", cudaGetErrorString(e));
    }
    *g->host_err = 0;
    *out = g;
    return FGL_OK;
}

int fgl_peer_group_destroy(fgl_peer_group *g) {
    if (!g) return FGL_OK;
    cudaSetDevice(g->device);
    if (g->ctx && g->ctx->stream) cudaStreamSynchronize(g->ctx->stream);
    if (g->ipc)
        for (int r = 0; r < g->nranks; r++) {
            if (r == g->rank) continue;
            if (g->set.color[r]) cudaIpcCloseMemHandle(g->set.color[r]);
            if (g->set.depth[r]) cudaIpcCloseMemHandle(g->set.depth[r]);
            if (g->set.dirty[r]) cudaIpcCloseMemHandle(g->set.dirty[r]);
            if (g->set.flags[r]) cudaIpcCloseMemHandle(g->set.flags[r]);
        }
    cudaGetLastError();
    if (g->dirty_all) cudaFree(g->dirty_all);
    if (g->host_err) cudaFreeHost(g->host_err);
    g->timer.destroy();
    delete g;
    return FGL_OK;
}

// phase 0: everything (ranks on different devices); 1: signal "drawn"; 2: wait for all, composite, signal "done";
// 3: wait for all "done".
int fgl_peer_composite_phase(fgl_ctx *c, fgl_peer_group *g, int root, int flags, int phase) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!g || g->ctx != c) return api_fail(c, FGL_E_INVALID, "peer group belongs to another context");
    if (root >= g->nranks) return api_fail(c, FGL_E_INVALID, "root %d outside the %d ranks", root, g->nranks);
    if (phase < 0 || phase > 3) return api_fail(c, FGL_E_INVALID, "phase %d outside 0..3", phase);
    std::lock_guard<std::mutex> lock(c->mu);
    api_fb_join(c);
    if (phase <= 1) c->peer_epoch = ++g->epoch;
    const unsigned long long epoch = g->epoch;
    const unsigned long long timeout_ns = 10ull * 1000ull * 1000ull * 1000ull;
    cudaStream_t st = c->stream;
    StageTimer *tm = c->profiling ? &g->timer : nullptr;
    if (phase <= 1) {
        if (tm) { tm->drain(); tm->create(); cudaEventRecord(tm->ev[0], st); }
        // my draws are complete -> tell everybody
        fgl::k_peer_signal<<<1, 32, 0, st>>>(g->set, g->rank, g->nranks, fgl::FLAG_READY, epoch);
    }
    if (phase == 0 || phase == 2) {
        // wait until everybody's are
        fgl::k_peer_wait<<<1, 32, 0, st>>>(g->flags, g->nranks, fgl::FLAG_READY, epoch, timeout_ns);
        if (tm) cudaEventRecord(tm->ev[1], st);
        const uint32_t ntiles = c->wb.ntiles;
        {
            const uint32_t n16 = (ntiles + 15u) / 16u;
            dim3 grid(min((n16 + 255u) / 256u, 64u), (unsigned)g->nranks);
            fgl::k_peer_gather_dirty<<<grid, 256, 0, st>>>(g->set, g->nranks, ntiles, g->dirty_all);
        }
        const int tiles_x = (c->w + c->tile_w - 1) / c->tile_w;
        fgl::k_peer_composite<<<c->wb.nsm * 8u, 256, 0, st>>>(g->set, g->rank, g->nranks, root, c->w, c->h, c->tile_w, tiles_x,
                                                            g->dirty_all, g->dirty_stride, c->clear_depth_value, c->clear_color_value,
                                                            c->clear_color_known ? 1 : 0, (flags & FGL_COMPOSITE_COLOR_ONLY) ? 1 : 0);
        if (tm) cudaEventRecord(tm->ev[2], st);
        fgl::k_peer_signal<<<1, 32, 0, st>>>(g->set, g->rank, g->nranks, fgl::FLAG_DONE, epoch);
    }
    if (phase == 0 || phase == 3) {
        // nobody may touch a framebuffer again (next frame's clear, a read-back) before every rank has finished
        // reading and writing it
        fgl::k_peer_wait<<<1, 32, 0, st>>>(g->flags, g->nranks, fgl::FLAG_DONE, epoch, timeout_ns);
        if (tm) { cudaEventRecord(tm->ev[3], st); cudaEventRecord(tm->ev[4], st); tm->pending = true; }
        CCK(c, cudaMemcpyAsync(g->host_err, g->flags + fgl::FLAG_ERROR, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
    CCK(c, cudaGetLastError());
    return FGL_OK;
}

int fgl_peer_composite(fgl_ctx *c, fgl_peer_group *g, int root, int flags) { return fgl_peer_composite_phase(c, g, root, flags, 0); }

int fgl_peer_stage_times(fgl_ctx *c, fgl_peer_group *g, float ms[4], uint32_t *composites) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!g || g->ctx != c || !ms) return api_fail(c, FGL_E_INVALID, "bad peer group / null argument");
    std::lock_guard<std::mutex> lock(c->mu);
    g->timer.drain();
    for (int k = 0; k < 4; k++) { ms[k] = g->timer.acc[k]; g->timer.acc[k] = 0; }
    if (composites) *composites = g->timer.n;
    g->timer.n = 0;
    return FGL_OK;
}

int fgl_peer_status(fgl_ctx *c, fgl_peer_group *g) {
    int rc = api_check_ctx(c);
    if (rc) return rc;
    if (!g || g->ctx != c) return api_fail(c, FGL_E_INVALID, "peer group belongs to another context");
    CCK(c, cudaStreamSynchronize(c->stream));
    if (*g->host_err) return api_fail(c, FGL_E_CUDA, "peer composite: a rank did not arrive within 10 s (its flags never reached this epoch)");
    return FGL_OK;
}

}  // extern "C"
