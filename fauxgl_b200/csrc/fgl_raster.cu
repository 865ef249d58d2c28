// fgl_raster.cu -- tile-parallel, ordered back end.  Persistent CTAs pull busy 64x8
// screen tiles heaviest-first from a device queue, keep the tile's depth (f64) (and
// colour, when shading inline) in shared memory, consume the tile's bin of span
// segments in primitive order and write the tile back once, coalesced -- replacing
// the reference's per-pixel mutex array.
//
// Replaces the per-pixel body of Context.rasterize (context.go:207-273),
// InterpolateVertexes (vertex.go:18-47), the three built-in Fragment shaders
// (shader.go:25,44,75), ImageTexture.BilinearSample (texture.go:41-63),
// Color.NRGBA (color.go:56) and the depth retest / write / blend (context.go:245-273).
//
// Arithmetic parity.  Each segment carries the reference's forward-differenced
// edge values at its first pixel (fgl_span.cu); the kernels here continue the same
// `w += a` chain (context.go:211-213), so barycentrics, depth and colour are
// bit-identical to a sequential run of the reference in triangle-index order.
//
// Ordering.  RT segments are resolved in parallel; pixels touched by several
// segments of a batch are settled in rounds: every pending fragment bids its
// batch index with a shared-memory atomicMin on a per-pixel ticket, the lowest
// wins, applies the reference's depth test / retest / write to the tile copy
// and retires.  That is index order per pixel, which makes `<=` ties,
// DepthBias, blending and UpdatedPixels well defined.
//
// Deferred shading.  When the draw's shader can neither discard nor blend
// (SolidColor, or Phong with an ObjectColor and no texture, alpha != 0 and no
// effective blending) the fragment colour cannot influence any depth decision:
// k_tile_resolve settles depth only and records, per pixel, the winning record and
// its edge values; k_shade then colours each pixel's FINAL winner once, as a
// separate full-occupancy kernel (no per-tile serial tail).  Otherwise
// k_tile_inline shades fragments in order, exactly like the reference.
#include "fgl_internal.h"
#include "fgl_block.cuh"
#include "fgl_math.cuh"
#include "fgl_shade.cuh"

namespace fgl {

constexpr int RT = 256;  // threads per CTA == segments per batch
constexpr uint32_t NO_TICKET = 0xffffffffu;
constexpr uint32_t NO_WINNER = 0xffffffffu;
static_assert(TILE_W <= 64, "the pending mask is one 64-bit word per segment");
static_assert((TILE_W * 4) % RT == 0, "k_shade splits a tile into chunks of RT pixels");

// ---- busy-tile queue, heaviest first ------------------------------------------------------------------
__global__ void k_tile_bucket(const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ tile_end,
                              uint32_t ntiles, TileCtl *ctl) {
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        const uint32_t cnt = tile_end[t] - tile_start[t];
        if (cnt) atomicAdd(&ctl->bucket_cnt[31 - __clz(cnt)], 1u);
    }
}
__global__ void k_tile_enqueue(const uint32_t *__restrict__ tile_start, const uint32_t *__restrict__ tile_end,
                               uint32_t ntiles, TileCtl *ctl, uint32_t *__restrict__ busy_list) {
    __shared__ uint32_t s_base[32];
    if (threadIdx.x < 32) {  // bucket b starts after all heavier buckets
        uint32_t base = 0;
        for (int b = 31; b > (int)threadIdx.x; b--) base += ctl->bucket_cnt[b];
        s_base[threadIdx.x] = base;
        if (blockIdx.x == 0 && threadIdx.x == 0) ctl->nbusy = base + ctl->bucket_cnt[0];
    }
    __syncthreads();
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < ntiles; t += gridDim.x * blockDim.x) {
        const uint32_t cnt = tile_end[t] - tile_start[t];
        if (cnt) {
            const int b = 31 - __clz(cnt);
            busy_list[s_base[b] + atomicAdd(&ctl->bucket_fill[b], 1u)] = t;
        }
    }
}

// ---- the ordered tile kernel ------------------------------------------------------------------------
// One fragment, inline mode: context.go:229-273 for a pixel whose ticket this thread holds.
FGL_DI void fragment_inline(const DrawParams &p, const fgl_state &st, const WorkBuffers &wb, const SegV &v,
                            double w0, double w1, double w2, int pi, double *s_depth, uint32_t *s_color,
                            unsigned long long &updated) {
    const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
    const double z = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
    const double bz = z + st.depth_bias;
    const double dcur = s_depth[pi];
    if (st.read_depth && bz > dcur) return;               // context.go:232
    const Rec *rp = wb.recs + v.rec;
    const double bx = b0 * rp->r0, by = b1 * rp->r1, bzz = b2 * rp->r2;  // context.go:236
    const double bw = 1 / (bx + by + bzz);
    AttrSrc a{&p, wb.clip_pool, rp->src, rp->flags};
    const C4 color = shade_fragment(p, a, bx, by, bzz, bw);
    if (c_is_discard(color)) return;                      // context.go:241
    if (bz <= dcur || !st.read_depth) {                   // context.go:248
        updated++;
        if (st.write_depth) s_depth[pi] = z;
        if (st.write_color) {
            const uint32_t c8 = c_nrgba(color);
            if (st.alpha_blend && color.a < 1) s_color[pi] = blend_over(s_color[pi], c8);
            else s_color[pi] = c8;
        }
    }
}

// EACH: also attribute UpdatedPixels to the primitive of every segment (fgl_draw_*_each).
template <bool DEFERRED, bool EACH>
__global__ void __launch_bounds__(RT, 3)
k_tile(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb,
       const uint32_t *__restrict__ seg_order, uint32_t *__restrict__ gcolor, double *__restrict__ gdepth) {
    if (wb.counters->overflow) return;  // work buffers too small: the host regrows and re-issues the draw

    extern __shared__ __align__(16) unsigned char smem_raw[];
    double *s_depth = reinterpret_cast<double *>(smem_raw);                                 // [TILE_PIX]
    double *s_w = s_depth + TILE_PIX;                                                        // [3][TILE_PIX] (deferred)
    uint32_t *s_color = reinterpret_cast<uint32_t *>(s_w + (DEFERRED ? 3 * TILE_PIX : 0));  // [TILE_PIX] (inline)
    uint32_t *s_ticket = s_color + (DEFERRED ? 0 : TILE_PIX);                                // [TILE_PIX]
    uint32_t *s_winner = s_ticket + TILE_PIX;                                                // [TILE_PIX] (deferred)
    __shared__ uint32_t s_q;

    const int tid = threadIdx.x;
    const fgl_state st = p.state;
    unsigned long long my_updated = 0;
    TileCtl *ctl = wb.tile_ctl;

    const int tpix = TILE_W * p.tile_h;
    bool first = true;
    while (true) {
        __syncthreads();
        if (tid == 0) {
            // Queue entries are ordered heaviest-first.  A CTA's FIRST tile is assigned by (SM, arrival
            // order on that SM) so that the heaviest tiles start on different SMs instead of on the few
            // SMs whose CTAs happen to start first; afterwards tiles are pulled dynamically.  Every entry
            // is claimed with an atomic exchange, so each tile is processed exactly once.
            const uint32_t nbusy = ctl->nbusy;
            uint32_t q = 0xffffffffu;
            if (first) {
                unsigned smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                if (smid < wb.nsm && smid < 256u) {
                    const uint32_t slot = atomicAdd(&ctl->sm_arrivals[smid], 1u);
                    const unsigned long long cand = (unsigned long long)slot * wb.nsm + smid;
                    if (cand < nbusy && atomicExch(&wb.tile_claimed[cand], 1u) == 0u) q = (uint32_t)cand;
                }
            }
            while (q == 0xffffffffu) {
                const uint32_t qq = atomicAdd(&ctl->head_resolve, 1u);
                if (qq >= nbusy) break;
                if (atomicExch(&wb.tile_claimed[qq], 1u) == 0u) q = qq;
            }
            s_q = q;
        }
        first = false;
        __syncthreads();
        const uint32_t q = s_q;
        if (q == 0xffffffffu) break;
        const uint32_t tile = wb.busy_list[q];
        const uint32_t bin_beg = wb.tile_start[tile], bin_end = wb.tile_end[tile];
        const long long t_begin = wb.tile_clock ? clock64() : 0;
        const int tile_x0 = (int)(tile % (uint32_t)p.tiles_x) * TILE_W;
        const int tile_y0 = (int)(tile / (uint32_t)p.tiles_x) * p.tile_h;
        const int tw = min(TILE_W, p.width - tile_x0);   // valid columns of this tile
        const int th = min(p.tile_h, p.height - tile_y0);  // valid rows

        // ---- load the tile ----------------------------------------------------------------------
        for (int i = tid; i < tpix; i += RT) {
            const int lx = i % TILE_W, ly = i / TILE_W;
            s_ticket[i] = NO_TICKET;
            if (DEFERRED) s_winner[i] = NO_WINNER;
            if (lx < tw && ly < th) {
                const size_t g = (size_t)(tile_y0 + ly) * p.width + (tile_x0 + lx);
                s_depth[i] = gdepth[g];
                if (!DEFERRED) s_color[i] = gcolor[g];
            }
        }
        __syncthreads();

        // The bin is a range of the sorted index array; segments are fetched through it two batches
        // ahead (index) / one batch ahead (segment), so neither load sits on the critical path.
        auto load_idx = [&](uint32_t i) { return i < bin_end ? seg_order[i] : 0xffffffffu; };
        auto load_seg = [&](uint32_t idx) {
            SegV s;
            s.cnt = 0; s.x = 0; s.yt = 0;
            if (idx != 0xffffffffu) s = wb.segv[idx];
            return s;
        };
        SegV v_next = load_seg(load_idx(bin_beg + tid));
        uint32_t idx_next2 = load_idx(bin_beg + RT + tid);
        for (uint32_t batch = bin_beg; batch < bin_end; batch += RT) {
            const SegV v = v_next;
            v_next = load_seg(idx_next2);
            idx_next2 = load_idx(batch + 2 * RT + tid);
            const int xa = (int)v.x, cnt = (int)v.cnt;
            const int rowbase = (int)v.yt * TILE_W - tile_x0;
            unsigned long long pend = cnt > 0 ? ((cnt >= 64 ? ~0ull : ((1ull << cnt) - 1ull)) << (xa - tile_x0)) : 0ull;

            // ---- ordered resolution in rounds ---------------------------------------------------
            while (true) {
                if (pend) {
                    unsigned long long m = pend;
                    while (m) {
                        const int bit = __ffsll((long long)m) - 1;
                        m &= m - 1;
                        atomicMin(&s_ticket[(int)v.yt * TILE_W + bit], (uint32_t)tid);
                    }
                }
                __syncthreads();
                if (pend) {
                    double w0 = v.w0, w1 = v.w1, w2 = v.w2;
                    const unsigned long long updated_before = my_updated;
                    for (int x = xa; x < xa + cnt; x++) {
                        const unsigned long long bitm = 1ull << (x - tile_x0);
                        const int pi = rowbase + x;
                        if ((pend & bitm) && s_ticket[pi] == (uint32_t)tid) {
                            pend &= ~bitm;
                            s_ticket[pi] = NO_TICKET;
                            if (DEFERRED) {
                                const double b0 = w0 * v.ra, b1 = w1 * v.ra, b2 = w2 * v.ra;
                                const double z = b0 * v.z0 + b1 * v.z1 + b2 * v.z2;  // context.go:230
                                const double bz = z + st.depth_bias;
                                const double dcur = s_depth[pi];
                                // context.go:232 early-out, then (no discard possible) the retest at :248
                                if (!(st.read_depth && bz > dcur) && (bz <= dcur || !st.read_depth)) {
                                    my_updated++;
                                    if (st.write_depth) s_depth[pi] = z;
                                    if (st.write_color) {
                                        s_winner[pi] = v.rec;
                                        s_w[pi] = w0; s_w[TILE_PIX + pi] = w1; s_w[2 * TILE_PIX + pi] = w2;
                                    }
                                }
                            } else {
                                fragment_inline(p, st, wb, v, w0, w1, w2, pi, s_depth, s_color, my_updated);
                            }
                        }
                        w0 += v.a12; w1 += v.a20; w2 += v.a01;  // context.go:211-213
                    }
                    if (EACH && my_updated != updated_before)
                        atomicAdd(&p.prim_info[2 * (size_t)rec_primitive(wb, p, v.rec) + 1], my_updated - updated_before);
                }
                if (!__syncthreads_or(pend != 0)) break;
            }
        }

        // ---- write the tile back -----------------------------------------------------------------
        const size_t vbase = (size_t)tile * tpix;
        const size_t vplane = (size_t)wb.ntiles * tpix;
        for (int i = tid; i < tpix; i += RT) {
            const int lx = i % TILE_W, ly = i / TILE_W;
            if (lx < tw && ly < th) {
                const size_t g = (size_t)(tile_y0 + ly) * p.width + (tile_x0 + lx);
                if (st.write_depth) gdepth[g] = s_depth[i];
                if (!DEFERRED) gcolor[g] = s_color[i];
            }
            if (DEFERRED && st.write_color) {
                const uint32_t win = s_winner[i];
                wb.vis_winner[vbase + i] = win;
                if (win != NO_WINNER) {
                    wb.vis_w[vbase + i] = s_w[i];
                    wb.vis_w[vplane + vbase + i] = s_w[TILE_PIX + i];
                    wb.vis_w[2 * vplane + vbase + i] = s_w[2 * TILE_PIX + i];
                }
            }
        }
        if (wb.tile_clock && tid == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            wb.tile_clock[2 * tile] = (unsigned long long)(clock64() - t_begin);
            wb.tile_clock[2 * tile + 1] = ((unsigned long long)smid << 32) | (bin_end - bin_beg);
        }
    }

    // UpdatedPixels: one atomic per warp per CTA lifetime
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my_updated += __shfl_down_sync(0xffffffffu, my_updated, o);
    if ((tid & 31) == 0 && my_updated) atomicAdd(&wb.counters->updated_pixels, my_updated);
}

// ---- deferred shading of the final winners ---------------------------------------------------------------
__global__ void __launch_bounds__(RT, 4)
k_shade(const __grid_constant__ DrawParams p, const __grid_constant__ WorkBuffers wb, uint32_t *__restrict__ gcolor) {
    if (wb.counters->overflow) return;
    const int tpix = TILE_W * p.tile_h;
    const uint32_t CHUNKS = (uint32_t)tpix / RT;
    __shared__ uint32_t s_q;
    TileCtl *ctl = wb.tile_ctl;
    const size_t vplane = (size_t)wb.ntiles * tpix;
    while (true) {
        __syncthreads();
        if (threadIdx.x == 0) s_q = atomicAdd(&ctl->head_shade, 1u);
        __syncthreads();
        const uint32_t q = s_q;
        if (q >= ctl->nbusy * CHUNKS) break;
        const uint32_t tile = wb.busy_list[q / CHUNKS];
        const int pi = (int)(q % CHUNKS) * RT + (int)threadIdx.x;
        const size_t vi = (size_t)tile * tpix + pi;
        const uint32_t rid = wb.vis_winner[vi];
        if (rid == NO_WINNER) continue;
        const int x = (int)(tile % (uint32_t)p.tiles_x) * TILE_W + pi % TILE_W;
        const int y = (int)(tile / (uint32_t)p.tiles_x) * p.tile_h + pi / TILE_W;
        uint32_t *out = gcolor + (size_t)y * p.width + x;
        if (p.kind == FGL_SHADER_SOLID) {  // SolidColorShader.Fragment, shader.go:25-27: nothing to interpolate
            *out = c_nrgba(c4(p.color[0], p.color[1], p.color[2], p.color[3]));
            continue;
        }
        // The deferred mode only admits Phong with an ObjectColor and no texture (fgl_api.cu), so the
        // attribute set is fixed: 9 normal + 9 position components.  All loads are issued up front.
        const Rec *rp = wb.recs + rid;
        const double ra = rp->ra, r0 = rp->r0, r1 = rp->r1, r2 = rp->r2;
        const uint32_t src = rp->src, flags = rp->flags;
        const double w0 = wb.vis_w[vi], w1 = wb.vis_w[vplane + vi], w2 = wb.vis_w[2 * vplane + vi];
        double n[3][3], pos[3][3];
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
        const double b0 = w0 * ra, b1 = w1 * ra, b2 = w2 * ra;
        const double bx = b0 * r0, by = b1 * r1, bzz = b2 * r2;  // context.go:236
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

static size_t tile_smem(bool deferred) {
    return deferred ? sizeof(double) * TILE_PIX * 4 + sizeof(uint32_t) * TILE_PIX * 2
                    : sizeof(double) * TILE_PIX + sizeof(uint32_t) * TILE_PIX * 2;
}

int launch_raster(const DrawParams &p, const WorkBuffers &wb, int sorted_buf, uint32_t *color, double *depth,
                  cudaStream_t st) {
    auto tile_kernel = p.deferred ? (p.prim_info ? k_tile<true, true> : k_tile<true, false>)
                                  : (p.prim_info ? k_tile<false, true> : k_tile<false, false>);
    cudaFuncSetAttribute(tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tile_smem(p.deferred));
    int launches = 0;
    cudaMemsetAsync(wb.tile_ctl, 0, sizeof(TileCtl), st);
    cudaMemsetAsync(wb.tile_claimed, 0, sizeof(uint32_t) * wb.ntiles, st);
    k_tile_bucket<<<148, 256, 0, st>>>(wb.tile_start, wb.tile_end, wb.ntiles, wb.tile_ctl);
    k_tile_enqueue<<<148, 256, 0, st>>>(wb.tile_start, wb.tile_end, wb.ntiles, wb.tile_ctl, wb.busy_list);
    launches += 2;
    const uint32_t grid = wb.ntiles < 148u * 4u ? wb.ntiles : 148u * 4u;
    if (p.deferred) {
        tile_kernel<<<grid, RT, tile_smem(true), st>>>(p, wb, wb.seg_val[sorted_buf], color, depth);
        launches++;
        if (p.state.write_color) {
            k_shade<<<148 * 4, RT, 0, st>>>(p, wb, color);
            launches++;
        }
    } else {
        tile_kernel<<<grid, RT, tile_smem(false), st>>>(p, wb, wb.seg_val[sorted_buf], color, depth);
        launches++;
    }
    return launches;
}

}  // namespace fgl
