// rel_attn.cu -- decoder self-attention with the position-relation bias generated on the fly.
//
// Upstream (models/bricks/relation_transformer.py:369-374, 453-459) materialises bias = PositionRelationEmbedding(
// src_boxes, tgt_boxes) as a [B, H, N, N] fp32 tensor (207-310 MB per decoder layer at B = 8), fills -inf where the
// denoising mask blocks, and hands it to nn.MultiheadAttention as attn_mask: it is written once, rewritten by the
// masked_fill_, read by the attention forward, kept for the backward, read again there, and its equally large
// gradient is written by the attention backward and read by the embedding's backward -- five times per step.
//
// Here the bias never exists in HBM.  The operator is the attention core between MultiheadAttention's input and
// output projections,
//     out[b,h,i,:] = sum_j softmax_j( q[b,h,i,:].k[b,h,j,:] / sqrt(D) + relu(W[h,:].f(src_i, tgt_j) + c[h]) [-inf if masked] ) v[b,h,j,:]
// and a CTA owns (image b, 32 query rows, all 8 heads):
//   forward   per 32-key tile: (A) all 8 warps build the bias tile [8 heads][32 keys][32 rows] in shared memory with the
//             FAST arithmetic of rel_fwd_fast_kernel (per-box tables + angle-difference identities, MUFU for the centre
//             features, packed fma.rn.f32x2 projection -- the 64 geometry features of a pair are shared by the heads);
//             (B) thread (row = lane, head = warp) runs an fp32 online softmax over the tile against K / V rows that all
//             lanes of the warp read from shared memory as broadcasts.  Saves only the log-sum-exp per row.
//   backward  recomputes the bias tile the same way, then (B) the row thread forms p, dp, ds and accumulates dq;
//             p and ds go to shared memory; (C) thread (key = lane, head = warp) accumulates dk, dv of the tile from
//             them, which leave as coalesced red.global.add.v4.f32 through a swizzled staging buffer; the score
//             gradient, gated by the ReLU of the recomputed bias, is handed to rel_bwd_kernel (rel.cu) through the
//             workspace for grad_weight / grad_bias.  (That hand-over still writes dS once; the forward bias, its
//             masked copy and SDPA's reads of both are gone.)
// CUDA cores, fp32 throughout: the 1e-4 gradient tolerance of the north star rules out tf32 / bf16 products here.
#include <stdlib.h>

#include "rel_common.cuh"

namespace rdetr {

constexpr int kAD = 32;       // head dimension
constexpr int kATI = 32;      // query rows per CTA
constexpr int kATJ = 32;      // keys per tile
constexpr int kAPad = 33;     // row pitch of the [head][key][row] tiles (conflict-free for both access directions)
constexpr int kAThreads = 32 * kRelHeads;

struct AttnSmem {
    RelConsts rc;                     // projection weights (transposed, duplicated), bias, angle constants (rel_common.cuh)
    float2 row2[kATI / 2][kTab];      // FAST table rows of the CTA's query boxes (src), as row pairs
    float tgt[kTab][kATJ];            // FAST table rows of the tile's key boxes (tgt), [field][key]
};

// Per-CTA constants (all threads; caller syncs afterwards).
__device__ __forceinline__ void attn_setup(AttnSmem &sm, const float *weight, const float *bias, const float *dim_t,
                                           const float *src_tab, int b, int i0, int N, int tid)
{
    rel_consts_setup(sm.rc, weight, bias, dim_t, tid, kAThreads);
    stage_row_pairs(&sm.row2[0][0], src_tab + ((size_t)b * N + i0) * kTab, min(kATI, N - i0), kATI, tid, kAThreads);
}

// K / V rows and the key boxes' table rows of tile j0 -> shared memory (rows past N are zero / clamped).
__device__ __forceinline__ void attn_load_tile(AttnSmem &sm, float4 *s_k4, float4 *s_v4, const float *k, const float *v,
                                               const float *tgt_tab, int b, int j0, int N, int tid, int lane, int warp)
{
    for (int e = tid; e < kRelHeads * kATJ * (kAD / 4); e += kAThreads) {
        const int h = e / (kATJ * 8), rem = e - h * (kATJ * 8);
        const int j = rem >> 3, c = rem & 7;
        const int gj = j0 + j;
        float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
        if (gj < N) {
            const size_t off = (((size_t)b * kRelHeads + h) * N + gj) * (kAD / 4) + c;
            kv = __ldg(reinterpret_cast<const float4 *>(k) + off);
            vv = __ldg(reinterpret_cast<const float4 *>(v) + off);
        }
        s_k4[e] = kv;
        s_v4[e] = vv;
    }
    const float *trow = tgt_tab + ((size_t)b * N + min(j0 + lane, N - 1)) * kTab;
    for (int f = warp; f < kTab; f += kRelHeads) sm.tgt[f][lane] = __ldg(trow + f);
}

// Phase A: the bias tile of (rows of this CTA) x (keys of tile j0), ReLU applied, -inf where masked / past N.
__device__ __forceinline__ void attn_bias_tile(const AttnSmem &sm, float *s_pb, float scale, const uint8_t *mask, int i0, int j0,
                                               int N, int lane, int warp)
{
    f32x2 acc[2][kRelHeads];
    fast_bias_rows<2>(sm.rc, &sm.row2[warp * 2][0], sm.tgt, scale, lane, acc);   // rows warp * 4 .. + 3 of the CTA tile
    const int gj = j0 + lane;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int i = warp * 4 + r, gi = i0 + i;
        const bool blocked = gj >= N || (mask != nullptr && gi < N && mask[(size_t)gi * N + gj] != 0);
#pragma unroll
        for (int h = 0; h < kRelHeads; ++h) {
            float lo, hi;
            unpack2(acc[r >> 1][h], lo, hi);
            const float a = (r & 1) ? hi : lo;
            s_pb[(h * kATJ + lane) * kAPad + i] = blocked ? -INFINITY : fmaxf(a, 0.f);
        }
    }
}

__global__ void __launch_bounds__(kAThreads, 2)
relattn_fwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                   const float *__restrict__ src_tab, const float *__restrict__ tgt_tab, const float *__restrict__ weight,
                   const float *__restrict__ bias, const float *__restrict__ dim_t, float scale,
                   const uint8_t *__restrict__ mask, float *__restrict__ out, float *__restrict__ lse, int N, float sm_scale,
                   int tiles_per_split, float *__restrict__ part_o, float *__restrict__ part_ml)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnSmem &sm = *reinterpret_cast<AttnSmem *>(smem_raw);
    float4 *s_k4 = reinterpret_cast<float4 *>(smem_raw + ((sizeof(AttnSmem) + 15) & ~size_t(15)));  // [8][32][8] float4
    float4 *s_v4 = s_k4 + kRelHeads * kATJ * (kAD / 4);
    float *s_pb = reinterpret_cast<float *>(s_v4 + kRelHeads * kATJ * (kAD / 4));                     // [8][32][33]

    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    const int b = blockIdx.y, i0 = blockIdx.x * kATI;
    attn_setup(sm, weight, bias, dim_t, src_tab, b, i0, N, tid);

    // thread (row = lane, head = warp)
    const int gi = i0 + lane;
    const bool rowok = gi < N;
    float qv[kAD], o[kAD];
    {
        const float4 *q4 = reinterpret_cast<const float4 *>(q) + (((size_t)b * kRelHeads + warp) * N + (rowok ? gi : 0)) * (kAD / 4);
#pragma unroll
        for (int c = 0; c < kAD / 4; ++c) {
            const float4 t = rowok ? __ldg(q4 + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            qv[4 * c] = t.x * sm_scale; qv[4 * c + 1] = t.y * sm_scale; qv[4 * c + 2] = t.z * sm_scale; qv[4 * c + 3] = t.w * sm_scale;
        }
#pragma unroll
        for (int d = 0; d < kAD; ++d) o[d] = 0.f;
    }
    float m = -INFINITY, l = 0.f;
    __syncthreads();

    // key split (small grids, e.g. the training shape B = 2): blockIdx.z owns the key tiles [jbeg, jend)
    const int jbeg = blockIdx.z * tiles_per_split * kATJ, jend = min(N, jbeg + tiles_per_split * kATJ);
    for (int j0 = jbeg; j0 < jend; j0 += kATJ) {
        attn_load_tile(sm, s_k4, s_v4, k, v, tgt_tab, b, j0, N, tid, lane, warp);
        __syncthreads();
        attn_bias_tile(sm, s_pb, scale, mask, i0, j0, N, lane, warp);
        __syncthreads();

        const float4 *kh = s_k4 + warp * kATJ * (kAD / 4);
        const float4 *vh = s_v4 + warp * kATJ * (kAD / 4);
        const float *pb = s_pb + (warp * kATJ) * kAPad + lane;
#pragma unroll 1
        for (int half = 0; half < kATJ; half += 16) {
            float s[16];
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float4 *kr = kh + (half + jj) * (kAD / 4);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
                for (int c = 0; c < kAD / 4; ++c) {
                    const float4 t = kr[c];  // same address for every lane of the warp: broadcast
                    a0 = fmaf(qv[4 * c], t.x, a0); a1 = fmaf(qv[4 * c + 1], t.y, a1);
                    a2 = fmaf(qv[4 * c + 2], t.z, a2); a3 = fmaf(qv[4 * c + 3], t.w, a3);
                }
                s[jj] = (a0 + a1) + (a2 + a3) + pb[(half + jj) * kAPad];
            }
            float mt = s[0];
#pragma unroll
            for (int jj = 1; jj < 16; ++jj) mt = fmaxf(mt, s[jj]);
            const float mn = fmaxf(m, mt);
            if (mn == -INFINITY) continue;  // every key so far is blocked for this row
            const float corr = __expf(m - mn);  // m = -inf -> 0
            l *= corr;
#pragma unroll
            for (int d = 0; d < kAD; ++d) o[d] *= corr;
#pragma unroll
            for (int jj = 0; jj < 16; ++jj) {
                const float p = __expf(s[jj] - mn);
                l += p;
                const float4 *vr = vh + (half + jj) * (kAD / 4);
#pragma unroll
                for (int c = 0; c < kAD / 4; ++c) {
                    const float4 t = vr[c];
                    o[4 * c] = fmaf(p, t.x, o[4 * c]); o[4 * c + 1] = fmaf(p, t.y, o[4 * c + 1]);
                    o[4 * c + 2] = fmaf(p, t.z, o[4 * c + 2]); o[4 * c + 3] = fmaf(p, t.w, o[4 * c + 3]);
                }
            }
            m = mn;
        }
        __syncthreads();  // the next tile overwrites K / V / bias
    }
    if (rowok && part_o != nullptr) {  // one of several key splits: unnormalised partial, merged by relattn_combine_kernel
        const size_t prow = (((size_t)blockIdx.z * gridDim.y + b) * kRelHeads + warp) * N + gi;
        float4 *o4 = reinterpret_cast<float4 *>(part_o) + prow * (kAD / 4);
#pragma unroll
        for (int c = 0; c < kAD / 4; ++c) o4[c] = make_float4(o[4 * c], o[4 * c + 1], o[4 * c + 2], o[4 * c + 3]);
        reinterpret_cast<float2 *>(part_ml)[prow] = make_float2(m, l);
    } else if (rowok) {
        const float inv = 1.0f / l;  // l == 0 (every key blocked) -> NaN row, as torch's softmax of an all -inf row
        float4 *o4 = reinterpret_cast<float4 *>(out) + (((size_t)b * kRelHeads + warp) * N + gi) * (kAD / 4);
#pragma unroll
        for (int c = 0; c < kAD / 4; ++c) o4[c] = make_float4(o[4 * c] * inv, o[4 * c + 1] * inv, o[4 * c + 2] * inv, o[4 * c + 3] * inv);
        lse[((size_t)b * kRelHeads + warp) * N + gi] = m + logf(l);
    }
}

// Merges the key splits of the forward: out = sum_s o_s e^(m_s - m) / sum_s l_s e^(m_s - m), lse = m + log(sum ...), m = max_s m_s.
// One thread per (image, head, row, 4 channels); `rows` = B * H * N.
__global__ void __launch_bounds__(256)
relattn_combine_kernel(const float *__restrict__ part_o, const float *__restrict__ part_ml, float *__restrict__ out,
                       float *__restrict__ lse, long long rows, int splits)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = idx >> 3;
    const int c = (int)(idx & 7);
    if (row >= rows) return;
    float m = -INFINITY;
    for (int s = 0; s < splits; ++s) m = fmaxf(m, reinterpret_cast<const float2 *>(part_ml)[(size_t)s * rows + row].x);
    float l = 0.f;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int s = 0; s < splits; ++s) {
        const float2 ml = reinterpret_cast<const float2 *>(part_ml)[(size_t)s * rows + row];
        const float w = ml.x == -INFINITY ? 0.f : __expf(ml.x - m);   // a split whose keys are all blocked contributes nothing
        const float4 o = reinterpret_cast<const float4 *>(part_o)[((size_t)s * rows + row) * (kAD / 4) + c];
        l = fmaf(ml.y, w, l);
        acc.x = fmaf(o.x, w, acc.x); acc.y = fmaf(o.y, w, acc.y); acc.z = fmaf(o.z, w, acc.z); acc.w = fmaf(o.w, w, acc.w);
    }
    const float inv = 1.0f / l;  // every key of the row blocked: 0 * inf = NaN, as the unsplit kernel and torch's softmax
    reinterpret_cast<float4 *>(out)[row * (kAD / 4) + c] = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
    if (c == 0) lse[row] = m + logf(l);
}

__global__ void __launch_bounds__(kAThreads, 1)
relattn_bwd_kernel(const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                   const float *__restrict__ src_tab, const float *__restrict__ tgt_tab, const float *__restrict__ weight,
                   const float *__restrict__ bias, const float *__restrict__ dim_t, float scale,
                   const uint8_t *__restrict__ mask, const float *__restrict__ out, const float *__restrict__ lse,
                   const float *__restrict__ gout, float *__restrict__ dq, float *__restrict__ dk, float *__restrict__ dv,
                   float *__restrict__ dS, int N, float sm_scale, int tiles_per_split)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AttnSmem &sm = *reinterpret_cast<AttnSmem *>(smem_raw);
    constexpr int kTile4 = kRelHeads * kATJ * (kAD / 4);
    float4 *s_k4 = reinterpret_cast<float4 *>(smem_raw + ((sizeof(AttnSmem) + 15) & ~size_t(15)));
    float4 *s_v4 = s_k4 + kTile4;
    float4 *s_q4 = s_v4 + kTile4;   // [8][32 rows][8]: sm_scale * q of the CTA's rows
    float4 *s_do4 = s_q4 + kTile4;  // [8][32 rows][8]: grad_out of the CTA's rows
    float *s_pb = reinterpret_cast<float *>(s_do4 + kTile4);  // [8][32 keys][33]: bias, then +-p (sign = ReLU gate)
    float *s_ds = s_pb + kRelHeads * kATJ * kAPad;            // [8][32 keys][33]: score gradient

    const int lane = threadIdx.x, warp = threadIdx.y, tid = warp * 32 + lane;
    const int b = blockIdx.y, i0 = blockIdx.x * kATI;
    attn_setup(sm, weight, bias, dim_t, src_tab, b, i0, N, tid);

    const int gi = i0 + lane;
    const bool rowok = gi < N;
    const size_t rowoff = ((size_t)b * kRelHeads + warp) * N + (rowok ? gi : 0);
    float qv[kAD], gov[kAD], dqv[kAD];
    float delta = 0.f;
    {
        const float4 *q4 = reinterpret_cast<const float4 *>(q) + rowoff * (kAD / 4);
        const float4 *g4 = reinterpret_cast<const float4 *>(gout) + rowoff * (kAD / 4);
        const float4 *o4 = reinterpret_cast<const float4 *>(out) + rowoff * (kAD / 4);
#pragma unroll
        for (int c = 0; c < kAD / 4; ++c) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 t = rowok ? __ldg(q4 + c) : z, g = rowok ? __ldg(g4 + c) : z, ov = rowok ? __ldg(o4 + c) : z;
            qv[4 * c] = t.x * sm_scale; qv[4 * c + 1] = t.y * sm_scale; qv[4 * c + 2] = t.z * sm_scale; qv[4 * c + 3] = t.w * sm_scale;
            gov[4 * c] = g.x; gov[4 * c + 1] = g.y; gov[4 * c + 2] = g.z; gov[4 * c + 3] = g.w;
            delta = fmaf(g.x, ov.x, fmaf(g.y, ov.y, fmaf(g.z, ov.z, fmaf(g.w, ov.w, delta))));
            s_q4[(warp * kATI + lane) * (kAD / 4) + c] = make_float4(qv[4 * c], qv[4 * c + 1], qv[4 * c + 2], qv[4 * c + 3]);
            s_do4[(warp * kATI + lane) * (kAD / 4) + c] = g;
        }
#pragma unroll
        for (int d = 0; d < kAD; ++d) dqv[d] = 0.f;
    }
    // a row past N must produce p = 0 everywhere: exp(s - inf) = 0
    const float row_lse = rowok ? __ldg(lse + rowoff) : INFINITY;
    __syncthreads();

    const int jbeg = blockIdx.z * tiles_per_split * kATJ, jend = min(N, jbeg + tiles_per_split * kATJ);   // key split, as the forward
    for (int j0 = jbeg; j0 < jend; j0 += kATJ) {
        attn_load_tile(sm, s_k4, s_v4, k, v, tgt_tab, b, j0, N, tid, lane, warp);
        __syncthreads();
        attn_bias_tile(sm, s_pb, scale, mask, i0, j0, N, lane, warp);
        __syncthreads();

        // ---- phase B: thread (row = lane, head = warp) ----------------------------------------------------
        {
            const float4 *kh = s_k4 + warp * kATJ * (kAD / 4);
            const float4 *vh = s_v4 + warp * kATJ * (kAD / 4);
            float *pb = s_pb + (warp * kATJ) * kAPad + lane;
            float *dsp = s_ds + (warp * kATJ) * kAPad + lane;
#pragma unroll 2
            for (int j = 0; j < kATJ; ++j) {
                const float4 *kr = kh + j * (kAD / 4), *vr = vh + j * (kAD / 4);
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, e0 = 0.f, e1 = 0.f, e2 = 0.f, e3 = 0.f;
                float4 kk[kAD / 4];
#pragma unroll
                for (int c = 0; c < kAD / 4; ++c) {
                    kk[c] = kr[c];
                    const float4 t = vr[c];
                    a0 = fmaf(qv[4 * c], kk[c].x, a0); a1 = fmaf(qv[4 * c + 1], kk[c].y, a1);
                    a2 = fmaf(qv[4 * c + 2], kk[c].z, a2); a3 = fmaf(qv[4 * c + 3], kk[c].w, a3);
                    e0 = fmaf(gov[4 * c], t.x, e0); e1 = fmaf(gov[4 * c + 1], t.y, e1);
                    e2 = fmaf(gov[4 * c + 2], t.z, e2); e3 = fmaf(gov[4 * c + 3], t.w, e3);
                }
                const float bij = pb[j * kAPad];
                const float sij = (a0 + a1) + (a2 + a3) + bij;
                const float p = __expf(sij - row_lse);  // bias = -inf (masked / past N) -> 0
                const float ds = p * (((e0 + e1) + (e2 + e3)) - delta);
#pragma unroll
                for (int c = 0; c < kAD / 4; ++c) {
                    dqv[4 * c] = fmaf(ds, kk[c].x, dqv[4 * c]); dqv[4 * c + 1] = fmaf(ds, kk[c].y, dqv[4 * c + 1]);
                    dqv[4 * c + 2] = fmaf(ds, kk[c].z, dqv[4 * c + 2]); dqv[4 * c + 3] = fmaf(ds, kk[c].w, dqv[4 * c + 3]);
                }
                pb[j * kAPad] = bij > 0.f ? p : -p;  // p >= 0: the sign carries the ReLU gate of the relation bias
                dsp[j * kAPad] = ds;
            }
        }
        __syncthreads();

        // ---- phase C: thread (key = lane, head = warp): dk, dv of the tile; the gated score gradient leaves -----
        {
            float dkv[kAD], dvv[kAD];
#pragma unroll
            for (int d = 0; d < kAD; ++d) dkv[d] = dvv[d] = 0.f;
            const float *pb = s_pb + (warp * kATJ + lane) * kAPad;
            const float *dsp = s_ds + (warp * kATJ + lane) * kAPad;
            const float4 *qh = s_q4 + warp * kATI * (kAD / 4);
            const float4 *gh = s_do4 + warp * kATI * (kAD / 4);
            const int gj = j0 + lane;
            float *dsrow = dS + (((size_t)b * kRelHeads + warp) * N + i0) * N + gj;
#pragma unroll 2
            for (int i = 0; i < kATI; ++i) {
                const float pv = pb[i], ds = dsp[i];
                const float p = fabsf(pv);
                if (gj < N && i0 + i < N) dsrow[(size_t)i * N] = (__float_as_uint(pv) >> 31) ? 0.f : ds;
#pragma unroll
                for (int c = 0; c < kAD / 4; ++c) {
                    const float4 qq = qh[i * (kAD / 4) + c], gg = gh[i * (kAD / 4) + c];  // broadcasts
                    dkv[4 * c] = fmaf(ds, qq.x, dkv[4 * c]); dkv[4 * c + 1] = fmaf(ds, qq.y, dkv[4 * c + 1]);
                    dkv[4 * c + 2] = fmaf(ds, qq.z, dkv[4 * c + 2]); dkv[4 * c + 3] = fmaf(ds, qq.w, dkv[4 * c + 3]);
                    dvv[4 * c] = fmaf(p, gg.x, dvv[4 * c]); dvv[4 * c + 1] = fmaf(p, gg.y, dvv[4 * c + 1]);
                    dvv[4 * c + 2] = fmaf(p, gg.z, dvv[4 * c + 2]); dvv[4 * c + 3] = fmaf(p, gg.w, dvv[4 * c + 3]);
                }
            }
            // K / V of this tile are dead (phase B is behind the barrier above): stage dk / dv there, XOR-swizzled so that
            // both this row-per-lane write and the 8-lanes-per-row read below are bank-conflict free
#pragma unroll
            for (int c = 0; c < kAD / 4; ++c) {
                const int slot = (warp * kATJ + lane) * (kAD / 4) + (c ^ (lane & 7));
                s_k4[slot] = make_float4(dkv[4 * c], dkv[4 * c + 1], dkv[4 * c + 2], dkv[4 * c + 3]);
                s_v4[slot] = make_float4(dvv[4 * c], dvv[4 * c + 1], dvv[4 * c + 2], dvv[4 * c + 3]);
            }
        }
        __syncthreads();
        {
            const int c = lane & 7;
#pragma unroll
            for (int rr = 0; rr < kATJ / 4; ++rr) {
                const int j = rr * 4 + (lane >> 3), gj = j0 + j;
                if (gj < N) {
                    const int slot = (warp * kATJ + j) * (kAD / 4) + (c ^ (j & 7));
                    const float4 a = s_k4[slot], bb = s_v4[slot];
                    const size_t off = (((size_t)b * kRelHeads + warp) * N + gj) * kAD + c * 4;
                    red_add_f32x4(dk + off, a.x, a.y, a.z, a.w);
                    red_add_f32x4(dv + off, bb.x, bb.y, bb.z, bb.w);
                }
            }
        }
        __syncthreads();  // the next tile overwrites K / V / bias
    }
    if (rowok && gridDim.z > 1) {  // key splits add their share of dq (zero-filled by the caller)
#pragma unroll
        for (int c = 0; c < kAD / 4; ++c)
            red_add_f32x4(dq + rowoff * kAD + 4 * c, dqv[4 * c] * sm_scale, dqv[4 * c + 1] * sm_scale, dqv[4 * c + 2] * sm_scale,
                          dqv[4 * c + 3] * sm_scale);
    } else if (rowok) {
        float4 *d4 = reinterpret_cast<float4 *>(dq) + rowoff * (kAD / 4);
#pragma unroll
        for (int c = 0; c < kAD / 4; ++c)
            d4[c] = make_float4(dqv[4 * c] * sm_scale, dqv[4 * c + 1] * sm_scale, dqv[4 * c + 2] * sm_scale, dqv[4 * c + 3] * sm_scale);
    }
}

static size_t fwd_smem_bytes() { return ((sizeof(AttnSmem) + 15) & ~size_t(15)) + 2 * kRelHeads * kATJ * kAD * sizeof(float) + kRelHeads * kATJ * kAPad * sizeof(float); }
static size_t bwd_smem_bytes() { return ((sizeof(AttnSmem) + 15) & ~size_t(15)) + 4 * kRelHeads * kATJ * kAD * sizeof(float) + 2 * kRelHeads * kATJ * kAPad * sizeof(float); }

// Key splits for small grids (B200: 148 SMs; the forward keeps 2 CTAs per SM resident, the backward 1).  The grid is
// B x ceil(N / 32) CTAs: 232 at B = 8, N = 900 (no split) but 70 at the training shape B = 2, N = 1100 and 91 at B = 1, N = 2900.
constexpr int kSmCount = 148;
constexpr int kMaxSplits = 8;
static int attn_splits(int B, int N, bool backward)
{
    const int tiles = (N + kATJ - 1) / kATJ;
    const long long ctas = (long long)B * ((N + kATI - 1) / kATI);
    if (const char *e = getenv("RDETR_RELATTN_SPLITS")) {   // tuning / tests: force a split count
        const int v = atoi(e);
        if (v >= 1) return v < tiles ? (v < kMaxSplits ? v : kMaxSplits) : tiles;
    }
    // time ~ waves x (key tiles per CTA + ~2 tiles' worth of per-CTA setup) + a little per split for the merge / the dq
    // reductions; fewest splits among the best
    const long long slots = backward ? kSmCount : 2 * kSmCount;
    int best = 1;
    double best_cost = 0.0;
    for (int ks = 1; ks <= kMaxSplits && ks <= tiles; ++ks) {
        const long long waves = (ctas * ks + slots - 1) / slots;
        const double cost = (double)waves * ((tiles + ks - 1) / ks + 2) + 0.5 * (ks - 1);
        if (ks == 1 || cost < best_cost - 1e-9) { best = ks; best_cost = cost; }
    }
    return best;
}
static size_t attn_partial_bytes(int B, int N, int H, int splits)
{
    return splits > 1 ? (size_t)splits * B * H * N * (kAD + 2) * sizeof(float) : 0;
}

static int validate_attn(const char *who, int B, int N, int H, int D)
{
    if (B < 0 || N < 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: negative size (B=%d N=%d)", who, B, N);
    if (H != kRelHeads || D != kAD) return fail(RDETR_ERR_UNSUPPORTED, "%s: H=%d D=%d unsupported (built for %d heads of %d)", who, H, D, kRelHeads, kAD);
    if (B > 65535) return fail(RDETR_ERR_UNSUPPORTED, "%s: B=%d exceeds gridDim.y", who, B);
    if ((long long)N * N >= (1LL << 31)) return fail(RDETR_ERR_UNSUPPORTED, "%s: N=%d too large", who, N);
    return RDETR_OK;
}

}  // namespace rdetr

extern "C" size_t rdetr_relation_attention_workspace_bytes(int B, int N, int H, int backward)
{
    if (B <= 0 || N <= 0) return 0;
    size_t bytes = rdetr_relation_workspace_bytes(B, N, N, RDETR_REL_FAST);
    bytes = (bytes + 255) & ~size_t(255);
    if (backward) bytes += (size_t)B * H * N * N * sizeof(float);  // the gated score gradient handed to rel_bwd_kernel
    else bytes += rdetr::attn_partial_bytes(B, N, H, rdetr::attn_splits(B, N, false));   // partial outputs of the forward's key splits
    return bytes;
}

extern "C" int rdetr_relation_attention_forward(const float *q, const float *k, const float *v, const float *src_boxes,
                                                const float *tgt_boxes, const float *weight, const float *bias,
                                                const float *dim_t, float scale, float eps, const uint8_t *attn_mask, float *out,
                                                float *lse, int B, int N, int H, int D, void *workspace, size_t workspace_bytes,
                                                rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_attn("rdetr_relation_attention_forward", B, N, H, D)) return rc;
    if (B == 0 || N == 0) return RDETR_OK;
    if (!q || !k || !v || !src_boxes || !tgt_boxes || !weight || !bias || !dim_t || !out || !lse)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_attention_forward: null pointer argument");
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)src_boxes | (uintptr_t)tgt_boxes) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_attention_forward: q/k/v/out/boxes must be 16-byte aligned");
    const DeviceGuard guard(out);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float *ts = nullptr, *tt = nullptr;
    if (int rc = prepare_tables("rdetr_relation_attention_forward", src_boxes, tgt_boxes, dim_t, scale, eps, B, N, N, workspace,
                                workspace_bytes, st, &ts, &tt))
        return rc;
    const size_t smem = fwd_smem_bytes();
    if (int rc = check_cuda(cudaFuncSetAttribute(relattn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(relattn_fwd)"))
        return rc;
    const int splits = attn_splits(B, N, false);
    const int tiles = (N + kATJ - 1) / kATJ, tiles_per_split = (tiles + splits - 1) / splits;
    float *part_o = nullptr, *part_ml = nullptr;
    if (splits > 1) {
        const size_t tab_bytes = (rdetr_relation_workspace_bytes(B, N, N, RDETR_REL_FAST) + 255) & ~size_t(255);
        if (workspace_bytes < tab_bytes + attn_partial_bytes(B, N, H, splits))
            return fail(RDETR_ERR_WORKSPACE, "rdetr_relation_attention_forward: workspace of %zu bytes required, got %zu",
                        rdetr_relation_attention_workspace_bytes(B, N, H, 0), workspace_bytes);
        part_o = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + tab_bytes);
        part_ml = part_o + (size_t)splits * B * H * N * kAD;
    }
    const dim3 block(32, kRelHeads), grid((N + kATI - 1) / kATI, B, splits);
    relattn_fwd_kernel<<<grid, block, smem, st>>>(q, k, v, ts, tt, weight, bias, dim_t, scale, attn_mask, out, lse, N,
                                                  1.0f / sqrtf((float)D), tiles_per_split, part_o, part_ml);
    if (int rc = check_cuda(cudaGetLastError(), "relattn_fwd_kernel launch")) return rc;
    if (splits > 1) {
        const long long rows = (long long)B * H * N;
        relattn_combine_kernel<<<(unsigned)((rows * 8 + 255) / 256), 256, 0, st>>>(part_o, part_ml, out, lse, rows, splits);
        return check_cuda(cudaGetLastError(), "relattn_combine_kernel launch");
    }
    return RDETR_OK;
}

extern "C" int rdetr_relation_attention_backward(const float *q, const float *k, const float *v, const float *src_boxes,
                                                 const float *tgt_boxes, const float *weight, const float *bias,
                                                 const float *dim_t, float scale, float eps, const uint8_t *attn_mask,
                                                 const float *out, const float *lse, const float *grad_out, float *grad_q,
                                                 float *grad_k, float *grad_v, float *grad_weight, float *grad_bias, int B, int N,
                                                 int H, int D, void *workspace, size_t workspace_bytes, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_attn("rdetr_relation_attention_backward", B, N, H, D)) return rc;
    if (!grad_weight || !grad_bias) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_attention_backward: null gradient buffer");
    const DeviceGuard guard(grad_weight);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = check_cuda(cudaMemsetAsync(grad_weight, 0, sizeof(float) * kRelHeads * kRelFeat, st), "cudaMemsetAsync(grad_weight)")) return rc;
    if (int rc = check_cuda(cudaMemsetAsync(grad_bias, 0, sizeof(float) * kRelHeads, st), "cudaMemsetAsync(grad_bias)")) return rc;
    if (B == 0 || N == 0) return RDETR_OK;
    if (!q || !k || !v || !src_boxes || !tgt_boxes || !weight || !bias || !dim_t || !out || !lse || !grad_out || !grad_q || !grad_k || !grad_v)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_attention_backward: null pointer argument");
    if (((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out | (uintptr_t)grad_out | (uintptr_t)grad_q | (uintptr_t)grad_k |
         (uintptr_t)grad_v | (uintptr_t)src_boxes | (uintptr_t)tgt_boxes) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_attention_backward: buffers must be 16-byte aligned");
    const size_t need = rdetr_relation_attention_workspace_bytes(B, N, H, 1);
    if (!workspace || workspace_bytes < need)
        return fail(RDETR_ERR_WORKSPACE, "rdetr_relation_attention_backward: workspace of %zu bytes required, got %zu", need,
                    workspace ? workspace_bytes : (size_t)0);
    const float *ts = nullptr, *tt = nullptr;
    if (int rc = prepare_tables("rdetr_relation_attention_backward", src_boxes, tgt_boxes, dim_t, scale, eps, B, N, N, workspace,
                                workspace_bytes, st, &ts, &tt))
        return rc;
    const size_t tab_bytes = (rdetr_relation_workspace_bytes(B, N, N, RDETR_REL_FAST) + 255) & ~size_t(255);
    float *dS = reinterpret_cast<float *>(static_cast<unsigned char *>(workspace) + tab_bytes);
    const size_t kv_bytes = (size_t)B * H * N * D * sizeof(float);
    if (int rc = check_cuda(cudaMemsetAsync(grad_k, 0, kv_bytes, st), "cudaMemsetAsync(grad_k)")) return rc;
    if (int rc = check_cuda(cudaMemsetAsync(grad_v, 0, kv_bytes, st), "cudaMemsetAsync(grad_v)")) return rc;
    const size_t smem = bwd_smem_bytes();
    if (int rc = check_cuda(cudaFuncSetAttribute(relattn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(relattn_bwd)"))
        return rc;
    const int splits = attn_splits(B, N, true);
    const int tiles = (N + kATJ - 1) / kATJ, tiles_per_split = (tiles + splits - 1) / splits;
    if (splits > 1)
        if (int rc = check_cuda(cudaMemsetAsync(grad_q, 0, kv_bytes, st), "cudaMemsetAsync(grad_q)")) return rc;
    const dim3 block(32, kRelHeads), grid((N + kATI - 1) / kATI, B, splits);
    relattn_bwd_kernel<<<grid, block, smem, st>>>(q, k, v, ts, tt, weight, bias, dim_t, scale, attn_mask, out, lse, grad_out, grad_q,
                                                  grad_k, grad_v, dS, N, 1.0f / sqrtf((float)D), tiles_per_split);
    if (int rc = check_cuda(cudaGetLastError(), "relattn_bwd_kernel launch")) return rc;
    // grad_weight / grad_bias from the gated score gradient (relu_bits = nullptr: already gated)
    return launch_rel_bwd_fast(src_boxes, tgt_boxes, ts, tt, dim_t, scale, eps, dS, nullptr, grad_weight, grad_bias, B, N, N, st);
}
