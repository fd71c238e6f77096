// msda_tile.cuh -- work decomposition shared by the tiled MSDA kernels (encoder self-attention, Nq == S).
//
// In the encoder every query IS a pixel of the value pyramid (reference points = the pixel grid,
// upstream models/bricks/base_transformer.py:57-70) and the learned offsets are a few pixels long
// (init: ms_deform_attn.py:266-278), so the samples of spatially adjacent queries of ONE head land in a small
// window of every level.  The flat kernels (msda_fwd.cu / msda_bwd.cu) give a CTA 8 heads of a few
// queries -- neighbouring queries run on different SMs and nothing is shared.  Here a CTA owns
//     (image b, head m, an 8x8 tile of the query grid of one level)
// and computes, per sampled level, the bounding box of every corner pixel its 64 x L*P samples touch.
// Levels whose box fits the CTA's shared-memory budget are served from shared memory (backward: exclusive
// fp32 grad_value accumulators flushed once per tile; forward: the staged value window); the others -- and
// any input without locality, e.g. uniform-random locations -- fall back to the global path of the flat
// kernels inside the same CTA.  Results never depend on which path a level took.
//
// Nothing here assumes the offsets are small: the boxes are measured from the data, per tile.
#pragma once

#include "common.cuh"

namespace rdetr {

constexpr int kTileH = 8, kTileW = 8, kTileQ = kTileH * kTileW;

// Levels of the pyramid plus the tiling of the query grid; one copy per CTA in shared memory.
struct TileGeom {
    int H[kMaxLevels], W[kMaxLevels], start[kMaxLevels];
    float invW[kMaxLevels], invH[kMaxLevels];
    int qstart[kMaxLevels];         // first query of level l when queries are the pixel grid
    int tiles_x[kMaxLevels];        // tiles per row of level l
    int tile_base[kMaxLevels + 1];  // prefix sum of tile counts
    int ntiles;
    int grid_mode;  // 1: sum_l H_l*W_l == Nq (queries = pixels); 0: 1-D chunks of 64 consecutive queries
};

// Filled by thread 0; callers __syncthreads() afterwards.
__device__ __forceinline__ void tile_geom_init(TileGeom &g, const int64_t *spatial_shapes, const int64_t *level_start_index, int L,
                                               int Nq)
{
    int total = 0, tiles = 0;
    for (int l = 0; l < L; ++l) {
        const int H = (int)spatial_shapes[2 * l], W = (int)spatial_shapes[2 * l + 1];
        g.H[l] = H;
        g.W[l] = W;
        g.start[l] = (int)level_start_index[l];
        g.invW[l] = 1.0f / (float)W;
        g.invH[l] = 1.0f / (float)H;
        g.qstart[l] = total;
        total += H * W;
        g.tiles_x[l] = (W + kTileW - 1) / kTileW;
        g.tile_base[l] = tiles;
        tiles += g.tiles_x[l] * ((H + kTileH - 1) / kTileH);
        // records pack h0+1 and w0+1 into 14 bits each (see pack_rec): refuse loudly rather than wrap
        if (H > 16382 || W > 16382) asm volatile("trap;");
    }
    g.tile_base[L] = tiles;
    g.grid_mode = (total == Nq);
    g.ntiles = g.grid_mode ? tiles : (Nq + kTileQ - 1) / kTileQ;
}

// Query of slot (0..63) of tile `tile`; -1 if the slot lies outside the level / past Nq.
__device__ __forceinline__ int tile_query(const TileGeom &g, int L, int Nq, int tile, int slot)
{
    if (!g.grid_mode) {
        const int q = tile * kTileQ + slot;
        return q < Nq ? q : -1;
    }
    int l = 0;
    while (l + 1 < L && tile >= g.tile_base[l + 1]) ++l;
    const int t = tile - g.tile_base[l];
    const int ty = t / g.tiles_x[l], tx = t - ty * g.tiles_x[l];
    const int qy = ty * kTileH + (slot >> 3), qx = tx * kTileW + (slot & 7);
    return (qy < g.H[l] && qx < g.W[l]) ? g.qstart[l] + qy * g.W[l] + qx : -1;
}

// One sample as the tiled kernels keep it in shared memory (16 bytes):
//   .x  bits: [0,14) h0+1   [14,28) w0+1   [28,32) corner validity (bit0 (h0,w0), bit1 (h0,w1), bit2 (h1,w0), bit3 (h1,w1))
//   .y  lw   .z  lh   .w  attention weight
// h0, w0 >= -1 for every sample inside the validity window (cuh:261-285 of the reference), so the +1 keeps
// the fields unsigned.  A record of all zeros is "contributes nothing".
struct TapHW {
    int h0, w0;
    unsigned vm;
    float lw, lh;
};

// Same arithmetic as make_tap (common.cuh), keeping (h0, w0) and a validity mask instead of four pixel indices.
__device__ __forceinline__ TapHW make_tap_hw(float lx, float ly, int H, int W)
{
    TapHW t;
    const float w_im = fmaf(lx, (float)W, -0.5f);
    const float h_im = fmaf(ly, (float)H, -0.5f);
    const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    const float h0f = floorf(h_im), w0f = floorf(w_im);
    t.lh = h_im - h0f;
    t.lw = w_im - w0f;
    t.h0 = -1; t.w0 = -1; t.vm = 0;
    if (inside) {
        t.h0 = (int)h0f;
        t.w0 = (int)w0f;
        const bool h0ok = t.h0 >= 0, h1ok = t.h0 + 1 <= H - 1, w0ok = t.w0 >= 0, w1ok = t.w0 + 1 <= W - 1;
        t.vm = ((h0ok && w0ok) ? 1u : 0u) | ((h0ok && w1ok) ? 2u : 0u) | ((h1ok && w0ok) ? 4u : 0u) | ((h1ok && w1ok) ? 8u : 0u);
    } else {
        t.lh = 0.f; t.lw = 0.f;  // NaN / inf locations must not leak into weights
    }
    return t;
}

__device__ __forceinline__ float4 pack_rec(const TapHW &t, float a)
{
    if (t.vm == 0) return make_float4(0.f, 0.f, 0.f, a);  // the weight survives: the fused softmax backward needs it
    const unsigned bits = (unsigned)(t.h0 + 1) | ((unsigned)(t.w0 + 1) << 14) | (t.vm << 28);
    return make_float4(__uint_as_float(bits), t.lw, t.lh, a);
}

struct RecView {
    unsigned vm;
    int h0, w0;
};
__device__ __forceinline__ RecView unpack_rec(float x)
{
    const unsigned bits = __float_as_uint(x);
    RecView r;
    r.vm = bits >> 28;
    r.h0 = (int)(bits & 0x3fffu) - 1;
    r.w0 = (int)((bits >> 14) & 0x3fffu) - 1;
    return r;
}

// Per-level window of one tile.  resident != 0: the level's corner pixels [x0, x0+bw) x [y0, y0+bh) own rows
// [off, off + bw*bh) of the CTA's shared-memory row buffer (128 bytes = one (pixel, head) row each).
struct TilePlace {
    int x0, y0, bw, bh, off, resident;
};

constexpr int kBBoxEmptyMin = 0x7fffffff, kBBoxEmptyMax = -0x7fffffff;

// Phase 1 of the tiled kernels, all threads of the CTA: taps of the tile's kTileQ x LP samples -> s_rec
// (slot-major, stride LP+1 records so that consecutive slots start 4 banks apart), the queries of the slots ->
// s_q, and the per-level bounding boxes of the valid corners -> s_bb[l] = (xmin, xmax, ymin, ymax) by shared-
// memory atomics (s_bb must hold (kBBoxEmptyMin, kBBoxEmptyMax, kBBoxEmptyMin, kBBoxEmptyMax) on entry).
// G = threads per query slot (power of two >= L*P, <= 32); thread j < L*P of a group owns sample j.
template <int L, int P, int G, int THREADS, typename IO>
__device__ __forceinline__ void tile_phase1(const IO &io, const TileGeom &geo, int b, int m, int tile, int S, int M, int Nq,
                                            float4 *s_rec, int *s_q, int (*s_bb)[4])
{
    constexpr int LP = L * P;
    constexpr int kGroups = THREADS / G;
    static_assert(G >= LP && G <= 32 && (G & (G - 1)) == 0, "group must cover the samples of a query");
    static_assert(THREADS % 32 == 0 && 32 % G == 0, "groups must not straddle warps");
    const int grp = threadIdx.x / G, j = threadIdx.x % G;
    const int lane = threadIdx.x & 31;
    const float inv_P = 1.0f / (float)P;
    for (int base = 0; base < kTileQ; base += kGroups) {  // same trip count for every thread: shuffles below are warp-wide
        const int slot = base + grp;
        int q = -1;
        if (j == 0 && slot < kTileQ) q = tile_query(geo, L, Nq, tile, slot);
        q = __shfl_sync(0xffffffffu, q, lane & ~(G - 1));  // one evaluation per query slot
        const bool live = q >= 0 && j < LP;
        const int l = j < LP ? j / P : L - 1;
        const long long bq = (long long)b * Nq + (q >= 0 ? q : 0);
        const long long gs = (bq * M + m) * LP + (j < LP ? j : 0);  // global sample index
        float a = 0.f;
        float2 xy = make_float2(-2.f, -2.f);
        if constexpr (IO::kFused) {
            float z = live ? ld_stream_scalar(io.logits + gs) : -INFINITY;
            float mx = z;
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float e = live ? __expf(z - mx) : 0.f;
            float sum = e;
#pragma unroll
            for (int off = G / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            if (live) {
                a = e / sum;
                const float2 off2 = ld_stream_pair(io.offsets + 2 * gs);
                xy = fused_location(io.ref + (bq * L + l) * io.ref_dim, io.ref_dim, off2.x, off2.y, geo.invW[l], geo.invH[l], inv_P);
            }
        } else {
            if (live) {
                xy = ld_stream_f2(reinterpret_cast<const float2 *>(io.loc) + gs);
                a = ld_stream_f1(io.attn + gs);
            }
        }
        float4 rec = make_float4(0.f, 0.f, 0.f, 0.f);
        int xmin = kBBoxEmptyMin, xmax = kBBoxEmptyMax, ymin = kBBoxEmptyMin, ymax = kBBoxEmptyMax;
        if (live) {
            const int H = geo.H[l], W = geo.W[l];
            TapHW t = make_tap_hw(xy.x, xy.y, H, W);
            if (t.vm && (geo.start[l] < 0 || geo.start[l] + (t.h0 + 1) * W + t.w0 + 1 >= S)) {  // inconsistent shape tensors only
                const int p00 = geo.start[l] + t.h0 * W + t.w0;
                if (geo.start[l] < 0 || p00 >= S) t.vm &= ~1u;
                if (geo.start[l] < 0 || p00 + 1 >= S) t.vm &= ~2u;
                if (geo.start[l] < 0 || p00 + W >= S) t.vm &= ~4u;
                if (geo.start[l] < 0 || p00 + W + 1 >= S) t.vm &= ~8u;
            }
            if constexpr (IO::kFused) {
                if (io.mask != nullptr && t.vm) {
                    const uint8_t *mrow = io.mask + (long long)b * S + geo.start[l] + t.h0 * W + t.w0;
                    if ((t.vm & 1u) && mrow[0]) t.vm &= ~1u;
                    if ((t.vm & 2u) && mrow[1]) t.vm &= ~2u;
                    if ((t.vm & 4u) && mrow[W]) t.vm &= ~4u;
                    if ((t.vm & 8u) && mrow[W + 1]) t.vm &= ~8u;
                }
            }
            rec = pack_rec(t, a);
            if (t.vm) {
                xmin = (t.vm & 5u) ? t.w0 : t.w0 + 1;
                xmax = (t.vm & 10u) ? t.w0 + 1 : t.w0;
                ymin = (t.vm & 3u) ? t.h0 : t.h0 + 1;
                ymax = (t.vm & 12u) ? t.h0 + 1 : t.h0;
            }
        }
        if (slot < kTileQ) {
            if (j < LP) s_rec[slot * (LP + 1) + j] = rec;
            if (j == 0) s_q[slot] = q;
        }
        // bounding boxes: combine the lanes of the warp that work on the same level, then one atomic per level
        const unsigned peers = __match_any_sync(0xffffffffu, l);
        xmin = __reduce_min_sync(peers, xmin);
        xmax = __reduce_max_sync(peers, xmax);
        ymin = __reduce_min_sync(peers, ymin);
        ymax = __reduce_max_sync(peers, ymax);
        if ((int)(__ffs(peers) - 1) == lane && xmin <= xmax) {
            atomicMin(&s_bb[l][0], xmin);
            atomicMax(&s_bb[l][1], xmax);
            atomicMin(&s_bb[l][2], ymin);
            atomicMax(&s_bb[l][3], ymax);
        }
    }
}

// Thread 0 after phase 1: give the levels shared-memory rows, coarsest level first (a coarse pixel receives
// the most updates per row of budget), and reset the boxes for the next tile.
template <int L>
__device__ __forceinline__ int tile_place_levels(int (*s_bb)[4], TilePlace *place, int cap_rows)
{
    int used = 0;
    for (int l = L - 1; l >= 0; --l) {
        const int xmin = s_bb[l][0], xmax = s_bb[l][1], ymin = s_bb[l][2], ymax = s_bb[l][3];
        s_bb[l][0] = kBBoxEmptyMin; s_bb[l][1] = kBBoxEmptyMax; s_bb[l][2] = kBBoxEmptyMin; s_bb[l][3] = kBBoxEmptyMax;
        TilePlace p{0, 0, 0, 0, 0, 0};
        if (xmin <= xmax) {
            p.x0 = xmin; p.y0 = ymin; p.bw = xmax - xmin + 1; p.bh = ymax - ymin + 1;
            const long long rows = (long long)p.bw * p.bh;
            if (used + rows <= cap_rows) { p.off = used; p.resident = 1; used += (int)rows; }
        }
        place[l] = p;
    }
    return used;
}

}  // namespace rdetr
