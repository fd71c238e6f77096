// msda_bwd_tile.cu -- tiled multi-scale deformable attention backward for the encoder (Nq == S), sm_100a.
//
// The flat backward (msda_bwd.cu) sends every one of the 4 x L*P corner rows of a (query, head) to the L2
// atomic unit as its own 128-byte vector reduction and runs at that unit's request rate (profiles/r01d):
// 91.4 M reductions per encoder call, half of them onto the two coarsest levels (6 % of the rows).
// This kernel cuts the reductions instead of tuning them (north star: "a scatter that cuts atomic contention"):
//
//   * work decomposition of msda_tile.cuh: a persistent CTA walks (image, 8x8 query tile, head) items; for each
//     item it measures the window of every level that the tile's samples touch and gives the windows that fit
//     shared-memory rows (coarsest level first);
//   * grad_value of a resident level is accumulated in those rows with PLAIN ld.shared / fma / st.shared -- no
//     shared-memory atomics (fp32 atomicAdd on shared memory is an ATOMS.CAST.SPIN loop on sm_100a: 15 clk per
//     row, slower than L2) -- by one "scatter warp" whose four 8-lane quarters (8 lanes x 4 channels = one
//     128-byte row) each own a different level: a quarter handles one sample at a time, the four corners of a
//     sample are four different pixels, and nobody else touches that level's window.  Program order within a
//     thread is the only ordering needed;
//   * the window is flushed once per item: one red.global.add.v4.f32 per touched (pixel, head) row instead of
//     one per corner -- measured 5.9x fewer L2 reductions at configs[1] (profiles/r02b_ncu_summary_tile_v1.md);
//   * the other eight "gather warps" do what the flat kernel does (8 lanes x 4 channels per (query, head),
//     value corners gathered with 16-byte read-only loads, grad_loc / grad_attn by xor-shuffles) at the same
//     time, and issue direct reductions only for levels that are not resident.  Their per-sample results stay
//     in registers (lane j of a group keeps samples j*K .. j*K+K-1) and leave as dense 16-byte stores.
//
// Correct for any input: a level without locality (window too large) simply takes the direct path.
//
// STATUS (round 2, measured on B200, profiles/r02a..r02c): parity-green but SLOWER than the flat kernel --
// 3.3 ms vs 1.8 ms at configs[1].  The flat kernel is bound by the L2 reduction rate (50 G rows/s) at 67 % issue
// utilisation; this kernel removes 83 % of the reductions but needs 1 300-1 700 warp instructions per (query, head)
// against 966 (records, windows, a second pass over every sample by the scatter warp, barriers between the
// phases: 30 % of the stall samples) and becomes issue / latency bound at 2-3 CTAs per SM.  It is therefore only
// used when rdetr_msda_set_tile_mode(2) / RDETR_MSDA_TILE=2 asks for it.  DESIGN.md section 7 has the numbers.
#include <atomic>
#include <cstdlib>

#include "msda_tile.cuh"

namespace rdetr {

constexpr int kBwdTileGatherWarps = 8;
constexpr int kBwdTileScatterWarps = 1;
constexpr int kBwdTileThreads = 32 * (kBwdTileGatherWarps + kBwdTileScatterWarps);

template <typename VT, int L, int P, typename IO, int MINB>
__global__ void __launch_bounds__(kBwdTileThreads, MINB)
msda_bwd_tile_kernel(const VT *__restrict__ value, const int64_t *__restrict__ spatial_shapes,
                     const int64_t *__restrict__ level_start_index, const IO io, const VT *__restrict__ grad_out,
                     float *__restrict__ grad_value_f32, int B, int S, int M, int Nq, int cap_rows)
{
    constexpr int D = 32, CH = 4, kLanes = D / CH;
    constexpr int LP = L * P;
    constexpr int G = LP <= 16 ? 16 : 32;
    constexpr int K = (LP + kLanes - 1) / kLanes;  // samples whose results one lane of a gather group keeps
    constexpr int kRecStride = LP + 1;
    using SL = Slice<VT, CH>;
    static_assert(LP <= 32, "tiled backward: at most 32 samples per (query, head)");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);                          // [kTileQ][LP+1]
    float *s_g = reinterpret_cast<float *>(s_rec + kTileQ * kRecStride);           // [kTileQ][32] grad_out rows of this head
    float *s_acc = s_g + kTileQ * D;                                               // [cap_rows][32]
    __shared__ TileGeom geo;
    __shared__ int s_q[kTileQ];
    __shared__ int s_bb[kMaxLevels][4];
    __shared__ TilePlace s_place[kMaxLevels];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) tile_geom_init(geo, spatial_shapes, level_start_index, L, Nq);
    if (tid < L) { s_bb[tid][0] = kBBoxEmptyMin; s_bb[tid][1] = kBBoxEmptyMax; s_bb[tid][2] = kBBoxEmptyMin; s_bb[tid][3] = kBBoxEmptyMax; }
    for (int i = tid; i < cap_rows * (D / 4); i += kBwdTileThreads) reinterpret_cast<float4 *>(s_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    const long long items = (long long)B * geo.ntiles * M;
    const int pix_stride = M * D;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int m = (int)(item % M);
        const long long bt = item / M;
        const int tile = (int)(bt % geo.ntiles);
        const int b = (int)(bt / geo.ntiles);

        // ---- phase 1: records, queries, windows; grad_out rows of the tile ---------------------------
        tile_phase1<L, P, G, kBwdTileThreads>(io, geo, b, m, tile, S, M, Nq, s_rec, s_q, s_bb);
        __syncthreads();
        if (tid == kBwdTileThreads - 1) tile_place_levels<L>(s_bb, s_place, cap_rows);
        for (int i = tid; i < kTileQ * kLanes; i += kBwdTileThreads) {
            const int slot = i / kLanes, l8 = i % kLanes;
            const int q = s_q[slot];
            float g4[CH] = {0.f, 0.f, 0.f, 0.f};
            if (q >= 0) SL::load_stream(grad_out + (((long long)b * Nq + q) * M + m) * D + l8 * CH, g4);
            *reinterpret_cast<float4 *>(s_g + slot * D + l8 * CH) = make_float4(g4[0], g4[1], g4[2], g4[3]);
        }
        __syncthreads();

        if (warp < kBwdTileGatherWarps) {
            // ---- gather warps: grad_loc / grad_attn, direct reductions for non-resident levels ----------
            const int grp = warp * 4 + (lane >> 3);  // 0..31
            const int l8 = lane & 7;
            const VT *vbase = value + ((long long)b * S * M + m) * D + l8 * CH;
            float *gvbase = grad_value_f32 + ((long long)b * S * M + m) * D + l8 * CH;
            int lvW[L], lvStart[L];
            float lvWf[L], lvHf[L];
            unsigned direct = 0;  // bit l: level l is not resident -> direct reductions
#pragma unroll
            for (int l = 0; l < L; ++l) {
                lvW[l] = geo.W[l];
                lvStart[l] = geo.start[l];
                lvWf[l] = (float)geo.W[l];
                lvHf[l] = (float)geo.H[l];
                if (!s_place[l].resident) direct |= 1u << l;
            }
#pragma unroll 1
            for (int pass = 0; pass < kTileQ / 32; ++pass) {
                const int slot = grp + 32 * pass;
                const int q = s_q[slot];
                const float4 g4 = *reinterpret_cast<const float4 *>(s_g + slot * D + l8 * CH);
                const float g[CH] = {g4.x, g4.y, g4.z, g4.w};
                const float4 *rec = s_rec + slot * kRecStride;
                float kgx[K], kgy[K], kga[K], ka[K];
#pragma unroll
                for (int k = 0; k < K; ++k) kgx[k] = kgy[k] = kga[k] = ka[k] = 0.f;
#pragma unroll
                for (int lp = 0; lp < LP; ++lp) {
                    const int l = lp / P;
                    const float4 r = rec[lp];
                    const RecView rv = unpack_rec(r.x);
                    const float lw = r.y, lh = r.z, a = r.w;
                    const float hw = 1.f - lw, hh = 1.f - lh;
                    const int Wl = lvW[l];
                    const long long base = (long long)(lvStart[l] + rv.h0 * Wl + rv.w0) * pix_stride;
                    const long long o0 = base, o1 = base + pix_stride, o2 = base + (long long)Wl * pix_stride, o3 = o2 + pix_stride;
                    float v0[CH], v1[CH], v2[CH], v3[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) v0[c] = v1[c] = v2[c] = v3[c] = 0.f;
                    if (rv.vm & 1u) SL::load(vbase + o0, v0);
                    if (rv.vm & 2u) SL::load(vbase + o1, v1);
                    if (rv.vm & 4u) SL::load(vbase + o2, v2);
                    if (rv.vm & 8u) SL::load(vbase + o3, v3);
                    const float w0 = hh * hw, w1 = hh * lw, w2 = lh * hw, w3 = lh * lw;
                    float p_attn = 0.f, p_gw = 0.f, p_gh = 0.f;
                    float tg[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        tg[c] = g[c] * a;
                        const float bil = fmaf(w3, v3[c], fmaf(w2, v2[c], fmaf(w1, v1[c], w0 * v0[c])));
                        const float dw = fmaf(hh, v1[c] - v0[c], lh * (v3[c] - v2[c]));  // cuh:110-141 of the reference
                        const float dh = fmaf(hw, v2[c] - v0[c], lw * (v3[c] - v1[c]));
                        p_attn = fmaf(g[c], bil, p_attn);
                        p_gw = fmaf(tg[c], dw, p_gw);
                        p_gh = fmaf(tg[c], dh, p_gh);
                    }
                    if (direct & (1u << l)) {  // direct path: one vector reduction per corner row, as the flat kernel
                        if (rv.vm & 1u) red_add_f32x4(gvbase + o0, w0 * tg[0], w0 * tg[1], w0 * tg[2], w0 * tg[3]);
                        if (rv.vm & 2u) red_add_f32x4(gvbase + o1, w1 * tg[0], w1 * tg[1], w1 * tg[2], w1 * tg[3]);
                        if (rv.vm & 4u) red_add_f32x4(gvbase + o2, w2 * tg[0], w2 * tg[1], w2 * tg[2], w2 * tg[3]);
                        if (rv.vm & 8u) red_add_f32x4(gvbase + o3, w3 * tg[0], w3 * tg[1], w3 * tg[2], w3 * tg[3]);
                    }
#pragma unroll
                    for (int off = kLanes / 2; off > 0; off >>= 1) {
                        p_attn += __shfl_xor_sync(0xffffffffu, p_attn, off);
                        p_gw += __shfl_xor_sync(0xffffffffu, p_gw, off);
                        p_gh += __shfl_xor_sync(0xffffffffu, p_gh, off);
                    }
                    if (l8 == lp / K) {  // compile-time owner and slot: stays in registers
                        kgx[lp % K] = p_gw * lvWf[l];
                        kgy[lp % K] = p_gh * lvHf[l];
                        kga[lp % K] = p_attn;
                        ka[lp % K] = a;
                    }
                }
                // softmax backward (fused prologue) needs sum_j a_j * dL/da_j over the samples of the (query, head);
                // every lane of the warp takes part in the shuffles, also those of empty slots
                float dot = 0.f;
                if constexpr (IO::kFused) {
#pragma unroll
                    for (int k = 0; k < K; ++k) dot = fmaf(ka[k], kga[k], dot);
#pragma unroll
                    for (int off = kLanes / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
                }
                if (q < 0) continue;
                const long long bq = (long long)b * Nq + q;
                const long long gs0 = (bq * M + m) * LP + l8 * K;  // first sample this lane reports
                if constexpr (IO::kFused) {
#pragma unroll
                    for (int k = 0; k < K; ++k) {
                        const int lp = l8 * K + k;
                        if (lp >= LP) break;
                        const int l = lp / P;
                        float gox, goy;
                        if (io.ref_dim == 2) {
                            gox = kgx[k] / (float)geo.W[l];  // runtime l: read the level table
                            goy = kgy[k] / (float)geo.H[l];
                        } else {
                            const float *rp = io.ref + (bq * L + l) * 4;
                            gox = ((kgx[k] * 0.5f) * rp[2]) / (float)P;
                            goy = ((kgy[k] * 0.5f) * rp[3]) / (float)P;
                        }
                        from_f32(io.grad_offsets[2 * (gs0 + k)], gox);
                        from_f32(io.grad_offsets[2 * (gs0 + k) + 1], goy);
                        from_f32(io.grad_logits[gs0 + k], ka[k] * (kga[k] - dot));
                    }
                } else {
                    if constexpr (K == 2) {
                        *reinterpret_cast<float4 *>(io.grad_loc + 2 * gs0) = make_float4(kgx[0], kgy[0], kgx[1], kgy[1]);
                        *reinterpret_cast<float2 *>(io.grad_attn + gs0) = make_float2(kga[0], kga[1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < K; ++k) {
                            if (l8 * K + k >= LP) break;
                            *reinterpret_cast<float2 *>(io.grad_loc + 2 * (gs0 + k)) = make_float2(kgx[k], kgy[k]);
                            io.grad_attn[gs0 + k] = kga[k];
                        }
                    }
                }
            }
        } else {
            // ---- scatter warp: exclusive shared-memory accumulation of the resident levels ------------------
            // quarter k of the warp (8 lanes x 4 channels = one 128-byte row per access) owns level 4*sweep + k:
            // the four quarters never share a window, a quarter handles one sample at a time, and the four
            // corners of a sample are four different pixels -> plain ld.shared / fma / st.shared is race-free.
            const int qk = lane >> 3, l8 = lane & 7;
#pragma unroll 1
            for (int sweep = 0; sweep * 4 < L; ++sweep) {
                const int lv = sweep * 4 + qk;
                const bool mine = lv < L && s_place[lv < L ? lv : 0].resident;
                if (!__any_sync(0xffffffffu, mine)) continue;
                const int l = mine ? lv : 0;  // lanes without a level walk level 0's records with every corner masked off
                const TilePlace pl = s_place[l];
                float *accl = s_acc + (long long)pl.off * D + l8 * CH;
                const int rowpitch = pl.bw * D;
                // records and the grad_out slice of slot+1 are fetched before slot is accumulated: the compiler
                // cannot move shared-memory loads above the accumulator stores on its own (they may alias)
                float4 rn[P];
                float4 gn = *reinterpret_cast<const float4 *>(s_g + l8 * CH);
                int qn = s_q[0];
#pragma unroll
                for (int p = 0; p < P; ++p) rn[p] = s_rec[l * P + p];
#pragma unroll 1
                for (int slot = 0; slot < kTileQ; ++slot) {
                    float4 r[P];
#pragma unroll
                    for (int p = 0; p < P; ++p) r[p] = rn[p];
                    const float4 gq = gn;
                    const int q = qn;
                    if (slot + 1 < kTileQ) {
                        gn = *reinterpret_cast<const float4 *>(s_g + (slot + 1) * D + l8 * CH);
                        qn = s_q[slot + 1];
#pragma unroll
                        for (int p = 0; p < P; ++p) rn[p] = s_rec[(slot + 1) * kRecStride + l * P + p];
                    }
                    if (q < 0) continue;
#pragma unroll
                    for (int p = 0; p < P; ++p) {
                        const RecView rv = unpack_rec(r[p].x);
                        const unsigned vm = mine ? rv.vm : 0u;
                        if (!__any_sync(0xffffffffu, vm != 0u)) continue;
                        const float lw = r[p].y, lh = r[p].z, a = r[p].w;
                        const float hw = 1.f - lw, hh = 1.f - lh;
                        const float w0 = (hh * hw) * a, w1 = (hh * lw) * a, w2 = (lh * hw) * a, w3 = (lh * lw) * a;
                        float4 *p0 = reinterpret_cast<float4 *>(accl + ((rv.h0 - pl.y0) * pl.bw + (rv.w0 - pl.x0)) * D);
                        float4 *p1 = p0 + D / 4, *p2 = reinterpret_cast<float4 *>(reinterpret_cast<float *>(p0) + rowpitch), *p3 = p2 + D / 4;
                        // the four corners are four different pixels: load all, then store all
                        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, a2 = a0, a3 = a0;
                        if (vm & 1u) a0 = *p0;
                        if (vm & 2u) a1 = *p1;
                        if (vm & 4u) a2 = *p2;
                        if (vm & 8u) a3 = *p3;
                        if (vm & 1u) *p0 = make_float4(fmaf(w0, gq.x, a0.x), fmaf(w0, gq.y, a0.y), fmaf(w0, gq.z, a0.z), fmaf(w0, gq.w, a0.w));
                        if (vm & 2u) *p1 = make_float4(fmaf(w1, gq.x, a1.x), fmaf(w1, gq.y, a1.y), fmaf(w1, gq.z, a1.z), fmaf(w1, gq.w, a1.w));
                        if (vm & 4u) *p2 = make_float4(fmaf(w2, gq.x, a2.x), fmaf(w2, gq.y, a2.y), fmaf(w2, gq.z, a2.z), fmaf(w2, gq.w, a2.w));
                        if (vm & 8u) *p3 = make_float4(fmaf(w3, gq.x, a3.x), fmaf(w3, gq.y, a3.y), fmaf(w3, gq.z, a3.z), fmaf(w3, gq.w, a3.w));
                    }
                }
            }
        }
        __syncthreads();

        // ---- flush: one vector reduction per touched (pixel, head) row; rows are left zeroed ---------------
        {
            const int l8 = lane & 7;
            float *gvb = grad_value_f32 + ((long long)b * S * M + m) * D + l8 * CH;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const TilePlace pl = s_place[l];
                if (!pl.resident) continue;
                const int Wl = geo.W[l], st = geo.start[l];
                for (int ry = warp; ry < pl.bh; ry += kBwdTileThreads / 32) {
                    const long long pixrow = st + (long long)(pl.y0 + ry) * Wl + pl.x0;
                    for (int rx = lane >> 3; rx < pl.bw; rx += 4) {
                        float4 *cell = reinterpret_cast<float4 *>(s_acc + (long long)(pl.off + ry * pl.bw + rx) * D + l8 * CH);
                        const float4 v = *cell;
                        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
                            red_add_f32x4(gvb + (pixrow + rx) * pix_stride, v.x, v.y, v.z, v.w);
                            *cell = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                }
            }
        }
        // no barrier needed here: the next phase 1 writes s_rec / s_g / s_q / s_bb, which the flush does not
        // read, and s_place is rewritten only after the barrier that follows phase 1
    }
}

static std::atomic<int> g_tile_mode{-1};  // -1 unset (environment decides), 0 auto, 1 flat only, 2 tiled whenever the shape allows

int msda_tile_mode()
{
    int mode = g_tile_mode.load(std::memory_order_relaxed);
    if (mode < 0) {
        const char *e = getenv("RDETR_MSDA_TILE");
        mode = e ? atoi(e) : 0;
        if (mode < 0 || mode > 2) mode = 0;
        g_tile_mode.store(mode, std::memory_order_relaxed);
    }
    return mode;
}

static std::atomic<int> g_tile_cap_rows{0};  // 0: default

int msda_tile_rows() { return g_tile_cap_rows.load(std::memory_order_relaxed); }

template <typename VT, int L, int P, typename IO>
static int launch_bwd_tile_lp(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out,
                              float *gv_f32, int B, int S, int M, int Nq, cudaStream_t stream)
{
    constexpr int MINB = 3;
    auto kern = msda_bwd_tile_kernel<VT, L, P, IO, MINB>;
    int cap_rows = g_tile_cap_rows.load(std::memory_order_relaxed);
    if (cap_rows <= 0) cap_rows = 384;
    const size_t smem = (size_t)kTileQ * (L * P + 1) * sizeof(float4) + (size_t)kTileQ * 32 * sizeof(float) + (size_t)cap_rows * 128;
    if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(msda_bwd_tile)"))
        return rc;
    int dev = 0, sms = 0, occ = 0;
    if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    if (int rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
    if (int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kBwdTileThreads, smem), "occupancy(msda_bwd_tile)"))
        return rc;
    if (occ < 1) return fail(RDETR_ERR_UNSUPPORTED, "msda_backward (tiled): %zu bytes of shared memory do not fit an SM", smem);
    const long long grid = (long long)sms * occ;  // persistent CTAs; the item count is only known on the device (shapes live there)
    kern<<<(unsigned)grid, kBwdTileThreads, smem, stream>>>(static_cast<const VT *>(value), shapes, lsi, io,
                                                             static_cast<const VT *>(grad_out), gv_f32, B, S, M, Nq, cap_rows);
    return check_cuda(cudaGetLastError(), "msda_bwd_tile_kernel launch");
}

// Returns -1 when the shape is outside what the tiled kernel is built for (caller uses the flat kernel).
template <typename VT, typename IO>
int launch_bwd_tile(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out, float *gv_f32,
                    int B, int S, int M, int L, int Nq, int P, cudaStream_t stream)
{
    if (P != 4) return -1;
    if (L == 4) return launch_bwd_tile_lp<VT, 4, 4, IO>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, Nq, stream);
    if (L == 5) return launch_bwd_tile_lp<VT, 5, 4, IO>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, Nq, stream);
    return -1;
}

template int launch_bwd_tile<float, PlainIO>(const void *, const int64_t *, const int64_t *, const PlainIO &, const void *, float *, int,
                                             int, int, int, int, int, cudaStream_t);
template int launch_bwd_tile<__nv_bfloat16, PlainIO>(const void *, const int64_t *, const int64_t *, const PlainIO &, const void *,
                                                     float *, int, int, int, int, int, int, cudaStream_t);
template int launch_bwd_tile<float, FusedIO<float>>(const void *, const int64_t *, const int64_t *, const FusedIO<float> &, const void *,
                                                    float *, int, int, int, int, int, int, cudaStream_t);
template int launch_bwd_tile<__nv_bfloat16, FusedIO<__nv_bfloat16>>(const void *, const int64_t *, const int64_t *,
                                                                    const FusedIO<__nv_bfloat16> &, const void *, float *, int, int, int,
                                                                    int, int, int, cudaStream_t);

}  // namespace rdetr

extern "C" int rdetr_msda_set_tile_mode(int mode)
{
    if (mode < 0 || mode > 2) return rdetr::fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_set_tile_mode: mode %d not in {0,1,2}", mode);
    rdetr::g_tile_mode.store(mode, std::memory_order_relaxed);
    return RDETR_OK;
}

extern "C" int rdetr_msda_set_tile_rows(int rows)
{
    if (rows < 0 || rows > 1600) return rdetr::fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_set_tile_rows: %d not in [0, 1600]", rows);
    rdetr::g_tile_cap_rows.store(rows, std::memory_order_relaxed);
    return RDETR_OK;
}
