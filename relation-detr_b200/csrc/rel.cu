// rel.cu -- fused PositionRelationEmbedding forward / backward for sm_100a.
//
// The reference (upstream models/bricks/relation_transformer.py:481-532) runs ~15 eager ATen ops
// that materialise [B,N1,N2,4] -> [B,N1,N2,4,8] x (sin,cos) -> [B,N1,N2,64] -> permute -> 1x1 conv
// -> ReLU -> clone: about 5.5 GB of HBM traffic for a 207 MB result at B=8, N=900.  Here the box
// pair geometry, the sin/cos encoding and the 64 -> H projection live in registers; the only HBM
// traffic is the [B,H,N1,N2] result (plus 1 bit per element for the ReLU mask) in the forward and
// its gradient in the backward.  No tensor cores: 64 -> 8 per pair on CUDA cores by design.
//
// Arithmetic, per pair (i = src row, j = tgt column), following the reference's evaluation order:
//   e0 = log(|x1-x2| / (w1+eps) + 1)   e1 = log(|y1-y2| / (h1+eps) + 1)
//   e2 = log((w1+eps) / (w2+eps))      e3 = log((h1+eps) / (h2+eps))
//   theta[c,k] = (e_c * scale) / dim_t[k];  f[c*16+2k] = sin(theta), f[c*16+2k+1] = cos(theta)
//   out[h] = relu(bias[h] + sum_n W[h,n] f[n])
//
// EXACT mode keeps every operation of that chain as torch evaluates it on CUDA (IEEE divisions,
// logf, (e*scale)/dim_t, sinf/cosf), so the 64 features are bit-identical to the eager path on the
// same device and only the order of the 64-term sum differs.
//
// FAST mode restructures the arithmetic (the op is FP32-pipe bound, not HBM bound: 512 FMA + 64
// transcendentals per pair for 32 bytes of output):
//   * the two size features are separable, e2 = log(w1+eps) - log(w2+eps), so sin/cos(theta) for
//     c = 2,3 come from per-box tables sin/cos(A_i), sin/cos(B_j) (computed once per box in fp64 by
//     a tiny pre-kernel) through the angle-difference identities: 4 FMA-pipe ops per angle and no
//     per-pair transcendental for half of the features;
//   * for the two centre features the per-row reciprocal 1/(w1+eps) replaces the division, the
//     division by dim_t becomes a multiply with one FMA correction step, and sin/cos use a
//     two-term Cody-Waite reduction followed by MUFU.SIN / MUFU.COS.
// Both modes walk kRowsPerIter src rows per thread so that each shared-memory weight fetch feeds
// kRowsPerIter * 4 FMAs.
#include <stdlib.h>

#include "rel_common.cuh"

namespace rdetr {

__global__ void __launch_bounds__(128)
rel_tables_kernel(const float *__restrict__ boxes, const float *__restrict__ dim_t, float scale, float eps,
                  float *__restrict__ table, int nboxes)
{
    // one thread per (box, k): 16 double-precision sin/cos pairs per box, nboxes ~ 1e4 => negligible
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int box = idx >> 3, k = idx & 7;
    if (box >= nboxes) return;
    const float4 bx = __ldg(reinterpret_cast<const float4 *>(boxes) + box);
    const float we = bx.z + eps, he = bx.w + eps;  // rounded to fp32 exactly as the reference does
    float *row = table + (size_t)box * kTab;
    if (k == 0) {
        row[0] = bx.x;
        row[1] = bx.y;
        row[2] = 1.0f / we;
        row[3] = 1.0f / he;
    }
    const double f = (double)scale / (double)__ldg(dim_t + k);
    double s, c;
    sincos(log((double)we) * f, &s, &c);
    row[4 + k] = (float)s;
    row[20 + k] = (float)c;
    sincos(log((double)he) * f, &s, &c);
    row[12 + k] = (float)s;
    row[28 + k] = (float)c;
}


// ------------------------------------------------------------------------------------------------
// forward.  CTA = 4 warps, tile = 32 tgt columns (lanes) x kFwdRowsPerCta src rows; each warp walks
// its rows kRowsPerIter at a time.  src rows are staged once per CTA in shared memory.
// relu_bits layout: [B][N1][ceil(N2/32)][H] words (the 8 heads of one (row, column word) contiguous).
// ------------------------------------------------------------------------------------------------
constexpr int kRelFwdWarps = 4;
constexpr int kRowsPerIter = 4;       // FAST: rows per thread and iteration (2: 0.326 ms, 4: 0.300, 8: 0.303 at N=900)
constexpr int kRowsPerIterExact = 2;  // EXACT: the sincosf-heavy body prefers fewer registers (2: 0.58 ms, 4: 0.62, 8: 0.83)
constexpr int kFwdRowsPerWarp = 16;
constexpr int kFwdRowsPerCta = kRelFwdWarps * kFwdRowsPerWarp;

template <bool FAST>
__global__ void __launch_bounds__(32 * kRelFwdWarps)
rel_fwd_kernel(const float *__restrict__ src, const float *__restrict__ tgt, const float *__restrict__ src_tab,
               const float *__restrict__ tgt_tab, const float *__restrict__ weight, const float *__restrict__ bias,
               const float *__restrict__ dim_t, float scale, float eps, const uint8_t *__restrict__ mask,
               float *__restrict__ out, uint32_t *__restrict__ relu_bits, int N1, int N2)
{
    constexpr int R = FAST ? kRowsPerIter : kRowsPerIterExact;
    constexpr int kRowF = FAST ? kTab : 4;  // floats staged per src row
    __shared__ __align__(16) float2 s_wt[kRelFeat][kRelHeads];  // transposed and duplicated: [n][h] = {w, w}
    __shared__ __align__(16) float s_row[kFwdRowsPerCta][kRowF];
    __shared__ float s_bias[kRelHeads];
    __shared__ float s_d[kRelK], s_invd[kRelK];
    __shared__ float s_tgt[FAST ? kTab : 1][32];  // FAST: tgt table of the CTA's columns, [field][lane]

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane;
    const int b = blockIdx.z;
    const int i_cta = blockIdx.y * kFwdRowsPerCta;
    for (int idx = tid; idx < kRelFeat * kRelHeads; idx += 32 * kRelFwdWarps) {
        const int h = idx / kRelFeat, n = idx - h * kRelFeat;
        const float wv = weight[idx];
        s_wt[n][h] = make_float2(wv, wv);
    }
    if (tid < kRelHeads) s_bias[tid] = bias[tid];
    if (tid < kRelK) {
        if constexpr (FAST) rev_constants(dim_t[tid], s_d[tid], s_invd[tid]);  // (chi, clo) of 1/(2 pi d_k)
        else { s_d[tid] = dim_t[tid]; s_invd[tid] = 1.0f / dim_t[tid]; }
    }
    {
        const float *rows = (FAST ? src_tab : src) + ((size_t)b * N1 + i_cta) * kRowF;
        const int nvalid = min(kFwdRowsPerCta, N1 - i_cta) * kRowF;
        float *flat = &s_row[0][0];
        for (int idx = tid; idx < kFwdRowsPerCta * kRowF; idx += 32 * kRelFwdWarps) flat[idx] = idx < nvalid ? rows[idx] : 1.0f;
    }
    __syncthreads();

    const int j = blockIdx.x * 32 + lane;
    const bool jok = j < N2;
    const int nwords = (N2 + 31) >> 5;

    // per-column constants.  FAST: the tgt table rows of the CTA's 32 columns live in shared memory
    // transposed ([field][lane], conflict-free) so the (c, k) loops below can stay rolled: fully
    // unrolled, this kernel is ~55 KB of SASS and stalls on instruction fetch (ncu: "no instruction").
    float x2 = 0.f, y2 = 0.f, w2e = 1.f, h2e = 1.f;
    if constexpr (FAST) {
        const float *trow = tgt_tab + ((size_t)b * N2 + (jok ? j : 0)) * kTab;
        for (int f = warp; f < kTab; f += kRelFwdWarps) s_tgt[f][lane] = __ldg(trow + f);
        __syncthreads();
        x2 = s_tgt[0][lane];
        y2 = s_tgt[1][lane];
    } else {
        const float4 tb = jok ? __ldg(reinterpret_cast<const float4 *>(tgt) + (size_t)b * N2 + j) : make_float4(0.f, 0.f, 1.f, 1.f);
        x2 = tb.x; y2 = tb.y; w2e = tb.z + eps; h2e = tb.w + eps;
    }

    for (int r0 = 0; r0 < kFwdRowsPerWarp; r0 += R) {
        const int lrow = warp * kFwdRowsPerWarp + r0;  // row inside the CTA tile
        const int i0 = i_cta + lrow;
        if (i0 >= N1) break;  // warp-uniform

#ifdef RDETR_REL_SCALAR_FMA
        float acc[R][kRelHeads];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int h = 0; h < kRelHeads; ++h) acc[r][h] = s_bias[h];
#define RDETR_PROJECT project_scalar<R>
#else
        f32x2 acc[R / 2][kRelHeads];
#pragma unroll
        for (int p = 0; p < R / 2; ++p)
#pragma unroll
            for (int h = 0; h < kRelHeads; ++h) acc[p][h] = pack2(s_bias[h], s_bias[h]);
#define RDETR_PROJECT project<R>
#endif

        // ---- centre features c = 0, 1 ----
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            const float t_xy = c == 0 ? x2 : y2;
            float es[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float *row = s_row[lrow + r];
                float e;
                if constexpr (FAST) e = logf(fmaf(fabsf(row[c] - t_xy), row[2 + c], 1.0f));
                else e = logf(fabsf(row[c] - t_xy) / (row[2 + c] + eps) + 1.0f);
                es[r] = e * scale;  // (x * scale) first, position_encoding.py:133
            }
#pragma unroll 2
            for (int k = 0; k < kRelK; ++k) {
                float sn[R], cs[R];
#pragma unroll
                for (int r = 0; r < R; ++r) angle_sincos<FAST>(es[r], s_d[k], s_invd[k], sn[r], cs[r]);
                RDETR_PROJECT(s_wt, c * 2 * kRelK + 2 * k, sn, cs, acc);
            }
        }
        // ---- size features c = 2, 3 ----
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            if constexpr (FAST) {
#pragma unroll 1
                for (int k4 = 0; k4 < kRelK; k4 += 4) {
                    float4 sA4[R], cA4[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        sA4[r] = *reinterpret_cast<const float4 *>(&s_row[lrow + r][4 + c * 8 + k4]);
                        cA4[r] = *reinterpret_cast<const float4 *>(&s_row[lrow + r][20 + c * 8 + k4]);
                    }
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const int k = k4 + kk;
                        const float sBk = s_tgt[4 + c * 8 + k][lane], cBk = s_tgt[20 + c * 8 + k][lane];
                        float sn[R], cs[R];
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            const float sA = kk == 0 ? sA4[r].x : kk == 1 ? sA4[r].y : kk == 2 ? sA4[r].z : sA4[r].w;
                            const float cA = kk == 0 ? cA4[r].x : kk == 1 ? cA4[r].y : kk == 2 ? cA4[r].z : cA4[r].w;
                            sn[r] = fmaf(sA, cBk, -(cA * sBk));  // sin(A - B)
                            cs[r] = fmaf(cA, cBk, sA * sBk);     // cos(A - B)
                        }
                        RDETR_PROJECT(s_wt, (2 + c) * 2 * kRelK + 2 * k, sn, cs, acc);
                    }
                }
            } else {
                const float t_den = c == 0 ? w2e : h2e;
                float es[R];
#pragma unroll
                for (int r = 0; r < R; ++r) es[r] = logf((s_row[lrow + r][2 + c] + eps) / t_den) * scale;
#pragma unroll 2
                for (int k = 0; k < kRelK; ++k) {
                    float sn[R], cs[R];
#pragma unroll
                    for (int r = 0; r < R; ++r) angle_sincos<false>(es[r], s_d[k], s_invd[k], sn[r], cs[r]);
                    RDETR_PROJECT(s_wt, (2 + c) * 2 * kRelK + 2 * k, sn, cs, acc);
                }
            }
        }

        // ---- epilogue: ReLU, optional -inf mask, 1-bit sign record ----
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = i0 + r;
            if (i >= N1) break;  // warp-uniform
            const bool blocked = mask != nullptr && jok && mask[(size_t)i * N2 + j] != 0;
            uint32_t words[kRelHeads];
#pragma unroll
            for (int h = 0; h < kRelHeads; ++h) {
#ifdef RDETR_REL_SCALAR_FMA
                const float a = acc[r][h];
#else
                float lo, hi;
                unpack2(acc[r >> 1][h], lo, hi);
                const float a = (r & 1) ? hi : lo;
#endif
                // a blocked (masked_fill -inf) position gets no gradient, whatever the caller feeds back there:
                // the reference's masked_fill_ cuts the graph at those elements (relation_transformer.py:372-374)
                const bool pos = jok && !blocked && a > 0.f;
                words[h] = __ballot_sync(0xffffffffu, pos);
                if (jok) out[(((size_t)b * kRelHeads + h) * N1 + i) * N2 + j] = blocked ? -INFINITY : (a > 0.f ? a : 0.f);
            }
            if (relu_bits != nullptr && lane == 0) {
                uint4 *dst = reinterpret_cast<uint4 *>(relu_bits + (((size_t)b * N1 + i) * nwords + blockIdx.x) * kRelHeads);
                dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
                dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// backward: grad_weight[h,n] = sum_pairs G[h] f[n], grad_bias[h] = sum_pairs G[h], G = grad_out * bit.
// CTA = 8 warps, tile = 32 columns (lanes) x kBwdRowsPerCta rows.  All warps walk the same pairs; warp
// w owns the feature chunk (box feature c = w/2, frequencies k = 4*(w%2) .. +3) => 8 features x 8
// heads = 64 accumulators per lane.  The gradient tile (8 heads x kBwdTileRows x 32) and its ReLU words are
// staged once per CTA in shared memory with cp.async (each element read from HBM exactly once,
// coalesced, double buffered), so the 8 warps read it with conflict-free LDS instead of eight
// redundant global loads.
// One shuffle tree + 64 atomics per warp at the end.
// ------------------------------------------------------------------------------------------------
constexpr int kRelBwdWarps = 8;
constexpr int kBwdRowsPerCta = 64;
constexpr int kBwdTileRows = 16;

template <bool FAST>
__global__ void __launch_bounds__(32 * kRelBwdWarps, 2)
rel_bwd_kernel(const float *__restrict__ src, const float *__restrict__ tgt, const float *__restrict__ src_tab,
               const float *__restrict__ tgt_tab, const float *__restrict__ dim_t, float scale, float eps,
               const float *__restrict__ grad_out, const uint32_t *__restrict__ relu_bits, float *__restrict__ grad_weight,
               float *__restrict__ grad_bias, int N1, int N2)
{
    constexpr int kRowF = FAST ? kTab : 4;
    __shared__ __align__(16) float s_g[2][kBwdTileRows][kRelHeads][32];
    __shared__ uint32_t s_bits[2][kBwdTileRows][kRelHeads];
    __shared__ __align__(16) float s_row[kBwdRowsPerCta][kRowF];  // rows >= nrows hold 1.0 (harmless geometry)

    const int lane = threadIdx.x;
    const int w = threadIdx.y;
    const int tid = w * 32 + lane;
    const int c = w >> 1;  // box feature of this warp
    const int k0 = (w & 1) * 4;
    const int b = blockIdx.z;
    const int j = blockIdx.x * 32 + lane;
    const bool jok = j < N2;
    const int nwords = (N2 + 31) >> 5;
    const int i_cta = blockIdx.y * kBwdRowsPerCta;
    const int nrows = min(kBwdRowsPerCta, N1 - i_cta);

    {
        const float *rows = (FAST ? src_tab : src) + ((size_t)b * N1 + i_cta) * kRowF;
        float *flat = &s_row[0][0];
        for (int idx = tid; idx < kBwdRowsPerCta * kRowF; idx += 32 * kRelBwdWarps) flat[idx] = idx < nrows * kRowF ? rows[idx] : 1.0f;
    }

    float d[4], invd[4];  // EXACT: (d_k, 1/d_k); FAST: (chi, clo) of 1/(2 pi d_k)
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if constexpr (FAST) rev_constants(__ldg(dim_t + k0 + k), d[k], invd[k]);
        else { d[k] = __ldg(dim_t + k0 + k); invd[k] = 1.0f / d[k]; }
    }
    // per-column constants of this warp's feature
    float t_xy = 0.f, t_den = 1.f;  // centre coordinate / (size + eps) of the tgt box for feature c
    float sB[4], cB[4];
    if constexpr (FAST) {
        const float *trow = tgt_tab + ((size_t)b * N2 + (jok ? j : 0)) * kTab;
        t_xy = __ldg(trow + (c & 1));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            sB[k] = __ldg(trow + 4 + (c & 1) * 8 + k0 + k);
            cB[k] = __ldg(trow + 20 + (c & 1) * 8 + k0 + k);
        }
    } else {
        const float4 tb = jok ? __ldg(reinterpret_cast<const float4 *>(tgt) + (size_t)b * N2 + j) : make_float4(0.f, 0.f, 1.f, 1.f);
        t_xy = (c & 1) ? tb.y : tb.x;
        t_den = ((c & 1) ? tb.w : tb.z) + eps;
    }

    float acc[kRelHeads][8];
    float accb[kRelHeads];
#pragma unroll
    for (int h = 0; h < kRelHeads; ++h) {
        accb[h] = 0.f;
#pragma unroll
        for (int n = 0; n < 8; ++n) acc[h][n] = 0.f;
    }

    // Tiles of kBwdTileRows rows of grad_out (8 heads x 32 columns) and their ReLU words are copied
    // global -> shared with cp.async, double buffered: the copy of tile t+1 is in flight while tile t
    // is consumed (ncu on the single-buffered version: long-scoreboard + barrier stalls dominate).
    auto issue_tile = [&](int t0, int buf) {
        const int trows = min(kBwdTileRows, nrows - t0);
        for (int r = 0; r < kBwdTileRows; ++r) {
            float *dst = &s_g[buf][r][w][lane];
            if (r < trows && jok) {
                const float *gsrc = grad_out + (((size_t)b * kRelHeads + w) * N1 + (i_cta + t0 + r)) * N2 + j;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(gsrc) : "memory");
            } else {
                *dst = 0.f;
            }
        }
        if (tid < kBwdTileRows * kRelHeads) {
            const int r = tid >> 3, h = tid & 7;
            uint32_t *dst = &s_bits[buf][r][h];
            if (r < trows && relu_bits == nullptr) {
                *dst = 0xffffffffu;  // gradient already gated by its producer (rel_attn.cu)
            } else if (r < trows) {
                const uint32_t *bsrc = relu_bits + (((size_t)b * N1 + (i_cta + t0 + r)) * nwords + blockIdx.x) * kRelHeads + h;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(bsrc) : "memory");
            } else {
                *dst = 0u;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    issue_tile(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < nrows; t0 += kBwdTileRows, buf ^= 1) {
        const int trows = min(kBwdTileRows, nrows - t0);
        if (t0 + kBwdTileRows < nrows) {
            issue_tile(t0 + kBwdTileRows, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // tile `buf` (and, on the first pass, s_row) visible to every warp
        // apply the ReLU mask once per CTA (warp w owns head w of the tile) instead of once per consumer warp
#pragma unroll 4
        for (int r = 0; r < kBwdTileRows; ++r) {
            const float raw = s_g[buf][r][w][lane];
            s_g[buf][r][w][lane] = ((s_bits[buf][r][w] >> lane) & 1u) ? raw : 0.f;
        }
        __syncthreads();

        // FAST: two rows per iteration (two independent feature chains in flight per lane; the kernel is latency
        // bound at 2 CTAs/SM otherwise).  EXACT: one row -- two inlined sincosf bodies per iteration are ~120 KB of
        // SASS and stall on instruction fetch.  Rows past `trows` read zeroed s_g / padded s_row.
        constexpr int U = FAST ? 2 : 1;
        for (int r = 0; r < trows; r += U) {
            float f[U][8];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const float *row = s_row[t0 + r + u];
                if constexpr (FAST) {
                    if (c < 2) {
                        const float es = logf(fmaf(fabsf(row[c] - t_xy), row[2 + c], 1.0f)) * scale;
#pragma unroll
                        for (int k = 0; k < 4; ++k) angle_sincos<true>(es, d[k], invd[k], f[u][2 * k], f[u][2 * k + 1]);
                    } else {
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const float sA = row[4 + (c & 1) * 8 + k0 + k], cA = row[20 + (c & 1) * 8 + k0 + k];
                            f[u][2 * k] = fmaf(sA, cB[k], -(cA * sB[k]));
                            f[u][2 * k + 1] = fmaf(cA, cB[k], sA * sB[k]);
                        }
                    }
                } else {
                    float e;
                    if (c < 2) e = logf(fabsf(row[c] - t_xy) / (row[2 + c] + eps) + 1.0f);
                    else e = logf((row[c] + eps) / t_den);
                    const float es = e * scale;
#pragma unroll
                    for (int k = 0; k < 4; ++k) angle_sincos<false>(es, d[k], invd[k], f[u][2 * k], f[u][2 * k + 1]);
                }
            }
#pragma unroll
            for (int h = 0; h < kRelHeads; ++h) {
                if constexpr (U == 2) {
                    const float g0 = s_g[buf][r][h][lane], g1 = s_g[buf][r + 1][h][lane];
#pragma unroll
                    for (int n = 0; n < 8; ++n) acc[h][n] = fmaf(g1, f[1][n], fmaf(g0, f[0][n], acc[h][n]));
                    if (w == 0) accb[h] += g0 + g1;  // grad_bias is accumulated by one warp only (warp-uniform branch)
                } else {
                    const float g0 = s_g[buf][r][h][lane];
#pragma unroll
                    for (int n = 0; n < 8; ++n) acc[h][n] = fmaf(g0, f[0][n], acc[h][n]);
                    if (w == 0) accb[h] += g0;
                }
            }
        }
        __syncthreads();  // every warp is done with tile `buf` before the copy after next overwrites it
    }

    // reduce over the 32 columns, then one atomic per (h, n) and warp
#pragma unroll
    for (int h = 0; h < kRelHeads; ++h) {
#pragma unroll
        for (int n = 0; n < 8; ++n) {
            float v = acc[h][n];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) atomicAdd(grad_weight + h * kRelFeat + c * 2 * kRelK + 2 * k0 + n, v);
        }
        if (w == 0) {
            float v = accb[h];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) atomicAdd(grad_bias + h, v);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FAST forward, round 2: the arithmetic of two src rows is packed into fp32x2 instructions (rel_common.cuh), the
// epilogue addresses with one pointer per row and a constant head stride and stores under a predicate (the round-1
// epilogue spent ~180 of ~850 instructions per pair on 64-bit index arithmetic and divergence bookkeeping).
// CTA = 4 warps, tile = 32 tgt columns (lanes) x 64 src rows; a warp walks its 16 rows 2 * RP at a time.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void st_global_pred(float *p, float v, bool ok)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)ok) : "memory");
}

template <int RP, bool MASKED, int MINB>
__global__ void __launch_bounds__(32 * kRelFwdWarps, MINB)
rel_fwd_fast_kernel(const float *__restrict__ src_tab, const float *__restrict__ tgt_tab, const float *__restrict__ weight,
                    const float *__restrict__ bias, const float *__restrict__ dim_t, float scale, const uint8_t *__restrict__ mask,
                    float *__restrict__ out, uint32_t *__restrict__ relu_bits, int N1, int N2)
{
    __shared__ __align__(16) RelConsts rc;
    __shared__ __align__(16) float2 s_row2[kFwdRowsPerCta / 2][kTab];
    __shared__ float s_tgt[kTab][32];

    const int lane = threadIdx.x, warp = threadIdx.y;
    const int tid = warp * 32 + lane;
    const int b = blockIdx.z;
    const int i_cta = blockIdx.y * kFwdRowsPerCta;
    const int j = blockIdx.x * 32 + lane;
    const bool jok = j < N2;
    rel_consts_setup(rc, weight, bias, dim_t, tid, 32 * kRelFwdWarps);
    stage_row_pairs(&s_row2[0][0], src_tab + ((size_t)b * N1 + i_cta) * kTab, min(kFwdRowsPerCta, N1 - i_cta), kFwdRowsPerCta, tid,
                    32 * kRelFwdWarps);
    {
        const float *trow = tgt_tab + ((size_t)b * N2 + (jok ? j : 0)) * kTab;
        for (int f = warp; f < kTab; f += kRelFwdWarps) s_tgt[f][lane] = __ldg(trow + f);
    }
    __syncthreads();

    const int nwords = (N2 + 31) >> 5;
    const size_t head_stride = (size_t)N1 * N2;
    for (int r0 = 0; r0 < kFwdRowsPerWarp; r0 += 2 * RP) {
        const int lrow = warp * kFwdRowsPerWarp + r0;  // row inside the CTA tile (even)
        const int i0 = i_cta + lrow;
        if (i0 >= N1) break;  // warp-uniform

        f32x2 acc[RP][kRelHeads];
        fast_bias_rows<RP>(rc, &s_row2[lrow >> 1][0], s_tgt, scale, lane, acc);

        // ---- epilogue: ReLU, optional -inf mask, 1-bit sign record ----
        float *prow = out + ((size_t)b * kRelHeads * N1 + i0) * N2 + j;
#pragma unroll
        for (int r = 0; r < 2 * RP; ++r) {
            const int i = i0 + r;
            const bool ok = jok && i < N1;
            bool blocked = false;
            if constexpr (MASKED) blocked = ok && mask[(size_t)i * N2 + j] != 0;
            float *p = prow + (size_t)r * N2;
            uint32_t words[kRelHeads];
#pragma unroll
            for (int h = 0; h < kRelHeads; ++h) {
                float lo, hi;
                unpack2(acc[r >> 1][h], lo, hi);
                const float a = (r & 1) ? hi : lo;
                // a blocked (masked_fill -inf) position gets no gradient, whatever the caller feeds back there:
                // the reference's masked_fill_ cuts the graph at those elements (relation_transformer.py:372-374)
                words[h] = __ballot_sync(0xffffffffu, ok && !blocked && a > 0.f);
                st_global_pred(p, blocked ? -INFINITY : fmaxf(a, 0.f), ok);
                p += head_stride;
            }
            if (relu_bits != nullptr && lane == 0 && i < N1) {
                uint4 *dst = reinterpret_cast<uint4 *>(relu_bits + (((size_t)b * N1 + i) * nwords + blockIdx.x) * kRelHeads);
                dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
                dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// FAST backward, round 2.  Same decomposition as rel_bwd_kernel (8 warps walk the same pairs, warp w owns box feature
// c = w / 2 and frequencies 4 * (w % 2) .. + 3: 8 features x 8 heads per lane; gradient tiles double buffered with
// cp.async), but the 64 accumulators are 32 fp32x2 pairs {sin, cos} of one frequency:
//   centre warps: the angle reduction runs packed over two FREQUENCIES ({chi_k, chi_k+1} are natural register pairs);
//   size warps:   {sin(A-B), cos(A-B)} = {sA, cA} * {cB, cB} + {-cA, sA} * {sB, sB}: one packed mul + one packed fma,
//                 the src side staged as float4 {sA, cA, -cA, sA} per (row, size feature, frequency);
//   grad_weight:  acc2[h][k] += {g, g} * {sin, cos}: 32 packed FMAs per pair instead of 64 scalar ones.
// Per-element operations and their order are those of the round-1 kernel.
// Dynamic shared memory (kRelBwdFastSmem bytes): gradient tiles, ReLU words, staged src rows.
// ------------------------------------------------------------------------------------------------
constexpr int kBwdFastTileRows = 32;   // half as many CTA barriers per pair as the 16-row tiles of rel_bwd_kernel
struct RelBwdFastSmem {
    float g[2][kBwdFastTileRows][kRelHeads][32];
    float4 rowS[kBwdRowsPerCta][2][kRelK];   // {sA, cA, -cA, sA} per (row, w / h, k)
    float4 geo[kBwdRowsPerCta];              // {cx, cy, 1 / (w + eps), 1 / (h + eps)}
    uint32_t bits[2][kBwdFastTileRows][kRelHeads];
};

__global__ void __launch_bounds__(32 * kRelBwdWarps, 2)
rel_bwd_fast_kernel(const float *__restrict__ src_tab, const float *__restrict__ tgt_tab, const float *__restrict__ dim_t, float scale,
                    const float *__restrict__ grad_out, const uint32_t *__restrict__ relu_bits, float *__restrict__ grad_weight,
                    float *__restrict__ grad_bias, int N1, int N2)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RelBwdFastSmem &sm = *reinterpret_cast<RelBwdFastSmem *>(smem_raw);

    const int lane = threadIdx.x;
    const int w = threadIdx.y;
    const int tid = w * 32 + lane;
    const int c = w >> 1;  // box feature of this warp
    const int k0 = (w & 1) * 4;
    const int b = blockIdx.z;
    const int j = blockIdx.x * 32 + lane;
    const bool jok = j < N2;
    const int nwords = (N2 + 31) >> 5;
    const int i_cta = blockIdx.y * kBwdRowsPerCta;
    const int nrows = min(kBwdRowsPerCta, N1 - i_cta);

    {   // src rows: rows >= nrows hold 1.0 (harmless geometry; their gradient tile entries are zero)
        const float *rows = src_tab + ((size_t)b * N1 + i_cta) * kTab;
        for (int idx = tid; idx < kBwdRowsPerCta * 17; idx += 32 * kRelBwdWarps) {
            const int r = idx / 17, e = idx - r * 17;
            const bool valid = r < nrows;
            const float *row = rows + (size_t)r * kTab;
            if (e == 16) {
                sm.geo[r] = valid ? make_float4(row[0], row[1], row[2], row[3]) : make_float4(1.f, 1.f, 1.f, 1.f);
            } else {
                const float sA = valid ? row[4 + e] : 1.f, cA = valid ? row[20 + e] : 1.f;   // e = (w / h) * 8 + k
                sm.rowS[r][e >> 3][e & 7] = make_float4(sA, cA, -cA, sA);
            }
        }
    }

    // per-warp constants: centre warps the two frequency PAIRS (k0, k0+1), (k0+2, k0+3); size warps the tgt side
    // (one register array for both roles: kc[0..1] = chi, kc[2..3] = -chi, kc[4..5] = clo  |  kc[0..3] = {cB, cB}, kc[4..7] = {sB, sB})
    f32x2 kc[8];
    float t_xy = 0.f;
    {
        const float *trow = tgt_tab + ((size_t)b * N2 + (jok ? j : 0)) * kTab;
        if (c < 2) {
            t_xy = __ldg(trow + c);
#pragma unroll
            for (int kp = 0; kp < 2; ++kp) {
                float h0, l0, h1, l1;
                rev_constants(__ldg(dim_t + k0 + 2 * kp), h0, l0);
                rev_constants(__ldg(dim_t + k0 + 2 * kp + 1), h1, l1);
                kc[kp] = pack2(h0, h1);
                kc[2 + kp] = pack2(-h0, -h1);
                kc[4 + kp] = pack2(l0, l1);
            }
            kc[6] = kc[7] = 0ull;
        } else {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float sB = __ldg(trow + 4 + (c & 1) * 8 + k0 + k), cB = __ldg(trow + 20 + (c & 1) * 8 + k0 + k);
                kc[k] = pack2(cB, cB);
                kc[4 + k] = pack2(sB, sB);
            }
        }
    }
    const f32x2 twopi2 = pack2(6.283185307179586f, 6.283185307179586f);
    const f32x2 magic2 = pack2(12582912.f, 12582912.f), nmagic2 = pack2(-12582912.f, -12582912.f);

    f32x2 acc[kRelHeads][4];   // {sum g sin, sum g cos} of frequency k0 + k
    float accb[kRelHeads];
#pragma unroll
    for (int h = 0; h < kRelHeads; ++h) {
        accb[h] = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[h][k] = 0ull;
    }

    auto issue_tile = [&](int t0, int buf) {
        const int trows = min(kBwdFastTileRows, nrows - t0);
        for (int r = 0; r < kBwdFastTileRows; ++r) {
            float *dst = &sm.g[buf][r][w][lane];
            if (r < trows && jok) {
                const float *gsrc = grad_out + (((size_t)b * kRelHeads + w) * N1 + (i_cta + t0 + r)) * N2 + j;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(gsrc) : "memory");
            } else {
                *dst = 0.f;
            }
        }
        if (tid < kBwdFastTileRows * kRelHeads) {
            const int r = tid >> 3, h = tid & 7;
            uint32_t *dst = &sm.bits[buf][r][h];
            if (r < trows && relu_bits == nullptr) {
                *dst = 0xffffffffu;  // gradient already gated by its producer (rel_attn.cu)
            } else if (r < trows) {
                const uint32_t *bsrc = relu_bits + (((size_t)b * N1 + (i_cta + t0 + r)) * nwords + blockIdx.x) * kRelHeads + h;
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(bsrc) : "memory");
            } else {
                *dst = 0u;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    issue_tile(0, 0);
    int buf = 0;
    for (int t0 = 0; t0 < nrows; t0 += kBwdFastTileRows, buf ^= 1) {
        const int trows = min(kBwdFastTileRows, nrows - t0);
        if (t0 + kBwdFastTileRows < nrows) {
            issue_tile(t0 + kBwdFastTileRows, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
        } else {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();  // tile `buf` (and, on the first pass, the staged rows) visible to every warp
        // apply the ReLU mask once per CTA (warp w owns head w of the tile) instead of once per consumer warp
#pragma unroll 4
        for (int r = 0; r < kBwdFastTileRows; ++r) {
            const float raw = sm.g[buf][r][w][lane];
            sm.g[buf][r][w][lane] = ((sm.bits[buf][r][w] >> lane) & 1u) ? raw : 0.f;
        }
        __syncthreads();

        // two rows per iteration (two independent feature chains in flight); rows past `trows` read zeroed gradients
        for (int r = 0; r < trows; r += 2) {
            f32x2 f2[2][4];
            if (c < 2) {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const float4 geo = sm.geo[t0 + r + u];
                    const float xy = c == 0 ? geo.x : geo.y, inv = c == 0 ? geo.z : geo.w;
                    const float es = logf(fmaf(fabsf(xy - t_xy), inv, 1.0f)) * scale;
                    const f32x2 es2 = pack2(es, es);
#pragma unroll
                    for (int kp = 0; kp < 2; ++kp) {
                        const f32x2 nn2 = fadd2(fadd2(fmul2(es2, kc[2 + kp]), magic2), nmagic2);   // -rint(es * chi), see angle_sincos2
                        f32x2 f = ffma2r(es2, kc[kp], nn2);
                        f = ffma2r(es2, kc[4 + kp], f);
                        float a0, a1;
                        unpack2(fmul2(f, twopi2), a0, a1);
                        f2[u][2 * kp] = pack2(__sinf(a0), __cosf(a0));
                        f2[u][2 * kp + 1] = pack2(__sinf(a1), __cosf(a1));
                    }
                }
            } else {
#pragma unroll
                for (int u = 0; u < 2; ++u)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const ulonglong2 A = *reinterpret_cast<const ulonglong2 *>(&sm.rowS[t0 + r + u][c & 1][k0 + k]);
                        f2[u][k] = ffma2r(A.x, kc[k], fmul2(A.y, kc[4 + k]));   // {sA cB - cA sB, cA cB + sA sB}
                    }
            }
#pragma unroll
            for (int h = 0; h < kRelHeads; ++h) {
                const float g0 = sm.g[buf][r][h][lane], g1 = sm.g[buf][r + 1][h][lane];
                const f32x2 g02 = pack2(g0, g0), g12 = pack2(g1, g1);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    ffma2(acc[h][k], g02, f2[0][k]);
                    ffma2(acc[h][k], g12, f2[1][k]);
                }
            }
        }
        if (w == 7) {  // grad_bias: one warp only, in a loop of its own (inside the loop above the compiler if-converts it for all 8 warps)
            for (int r = 0; r < trows; r += 2)
#pragma unroll
                for (int h = 0; h < kRelHeads; ++h) accb[h] += sm.g[buf][r][h][lane] + sm.g[buf][r + 1][h][lane];
        }
        __syncthreads();  // every warp is done with tile `buf` before the copy after next overwrites it
    }

    // reduce over the 32 columns, then one atomic per (h, n) and warp
#pragma unroll
    for (int h = 0; h < kRelHeads; ++h) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float vs, vc;
            unpack2(acc[h][k], vs, vc);
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                vs += __shfl_xor_sync(0xffffffffu, vs, off);
                vc += __shfl_xor_sync(0xffffffffu, vc, off);
            }
            if (lane == 0) {
                atomicAdd(grad_weight + h * kRelFeat + c * 2 * kRelK + 2 * (k0 + k), vs);
                atomicAdd(grad_weight + h * kRelFeat + c * 2 * kRelK + 2 * (k0 + k) + 1, vc);
            }
        }
        if (w == 7) {
            float v = accb[h];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) atomicAdd(grad_bias + h, v);
        }
    }
}

// rows per thread and iteration of the FAST forward: 4 (default) or 8 (RDETR_REL_ROWS=8, tuning only)
static int rel_fast_rows_per_iter()
{
    const char *e = getenv("RDETR_REL_ROWS");
    return (e && e[0] == '8') ? 8 : 4;
}

template <int RP>
static int launch_rel_fwd_fast(const float *ts, const float *tt, const float *weight, const float *bias, const float *dim_t, float scale,
                               const uint8_t *mask, float *out, uint32_t *relu_bits, int B, int N1, int N2, cudaStream_t st)
{
    const dim3 block(32, kRelFwdWarps);
    const dim3 grid((N2 + 31) / 32, (N1 + kFwdRowsPerCta - 1) / kFwdRowsPerCta, B);
    // no register cap: ptxas takes 120 registers at RP = 2 (4 CTAs / SM) and that is the fastest variant measured -- 0.248 ms at
    // B = 8, N = 900 against 0.256 (88 registers, 5 CTAs) and 0.261 (80 registers, 6 CTAs): the kernel wants ILP, not warps
    if (mask) rel_fwd_fast_kernel<RP, true, 1><<<grid, block, 0, st>>>(ts, tt, weight, bias, dim_t, scale, mask, out, relu_bits, N1, N2);
    else rel_fwd_fast_kernel<RP, false, 1><<<grid, block, 0, st>>>(ts, tt, weight, bias, dim_t, scale, nullptr, out, relu_bits, N1, N2);
    return check_cuda(cudaGetLastError(), "rel_fwd_fast_kernel launch");
}

static int validate_rel(const char *who, int B, int N1, int N2, int H, int flags)
{
    if (B < 0 || N1 < 0 || N2 < 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: negative size (B=%d N1=%d N2=%d)", who, B, N1, N2);
    if (H != kRelHeads) return fail(RDETR_ERR_UNSUPPORTED, "%s: H=%d unsupported (kernels are built for %d heads)", who, H, kRelHeads);
    if (flags != RDETR_REL_EXACT && flags != RDETR_REL_FAST) return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: unknown flags %d", who, flags);
    if (B > 65535) return fail(RDETR_ERR_UNSUPPORTED, "%s: B=%d exceeds gridDim.z", who, B);
    return RDETR_OK;
}

// FAST mode: fills the per-box tables in `workspace` ([B*N1 + B*N2] rows of kTab floats)
int prepare_tables(const char *who, const float *src, const float *tgt, const float *dim_t, float scale, float eps,
                          int B, int N1, int N2, void *workspace, size_t workspace_bytes, cudaStream_t st,
                          const float **src_tab, const float **tgt_tab)
{
    const size_t need = rdetr_relation_workspace_bytes(B, N1, N2, RDETR_REL_FAST);
    if (!workspace || workspace_bytes < need)
        return fail(RDETR_ERR_WORKSPACE, "%s: FAST mode needs a workspace of %zu bytes, got %zu", who, need, workspace ? workspace_bytes : (size_t)0);
    if ((uintptr_t)workspace & 15) return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: workspace must be 16-byte aligned", who);
    float *ts = static_cast<float *>(workspace);
    float *tt = ts + (size_t)B * N1 * kTab;
    const int ns = B * N1, nt = B * N2;
    rel_tables_kernel<<<(ns * 8 + 127) / 128, 128, 0, st>>>(src, dim_t, scale, eps, ts, ns);
    rel_tables_kernel<<<(nt * 8 + 127) / 128, 128, 0, st>>>(tgt, dim_t, scale, eps, tt, nt);
    *src_tab = ts;
    *tgt_tab = tt;
    return check_cuda(cudaGetLastError(), "rel_tables_kernel launch");
}

// grad_weight / grad_bias (accumulated: the caller zeroes them) from a gradient tile stream; relu_bits == nullptr means
// the gradient is already gated.  Used by rdetr_relation_backward and by the fused relation attention (rel_attn.cu).
int launch_rel_bwd_fast(const float *src, const float *tgt, const float *src_tab, const float *tgt_tab, const float *dim_t,
                        float scale, float eps, const float *grad, const uint32_t *relu_bits, float *grad_weight,
                        float *grad_bias, int B, int N1, int N2, cudaStream_t st)
{
    const dim3 block(32, kRelBwdWarps);
    const dim3 grid((N2 + 31) / 32, (N1 + kBwdRowsPerCta - 1) / kBwdRowsPerCta, B);
    if (grid.y > 65535) return fail(RDETR_ERR_UNSUPPORTED, "relation backward: N1=%d too large", N1);
    if (int rc = check_cuda(cudaFuncSetAttribute(rel_bwd_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RelBwdFastSmem)),
                            "cudaFuncSetAttribute(rel_bwd_fast)"))
        return rc;
    rel_bwd_fast_kernel<<<grid, block, sizeof(RelBwdFastSmem), st>>>(src_tab, tgt_tab, dim_t, scale, grad, relu_bits, grad_weight, grad_bias, N1, N2);
    return check_cuda(cudaGetLastError(), "rel_bwd_kernel launch");
}

}  // namespace rdetr

extern "C" size_t rdetr_relation_workspace_bytes(int B, int N1, int N2, int flags)
{
    if (flags != RDETR_REL_FAST || B <= 0) return 0;
    return ((size_t)B * (size_t)(N1 > 0 ? N1 : 0) + (size_t)B * (size_t)(N2 > 0 ? N2 : 0)) * rdetr::kTab * sizeof(float);
}

extern "C" int rdetr_relation_forward(const float *src_boxes, const float *tgt_boxes, const float *weight, const float *bias,
                                      const float *dim_t, float scale, float eps, const uint8_t *attn_mask, float *out,
                                      uint32_t *relu_bits, int B, int N1, int N2, int H, int flags, void *workspace,
                                      size_t workspace_bytes, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_rel("rdetr_relation_forward", B, N1, N2, H, flags)) return rc;
    if (B == 0 || N1 == 0 || N2 == 0) return RDETR_OK;
    if (!src_boxes || !tgt_boxes || !weight || !bias || !dim_t || !out)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_forward: null pointer argument");
    if (((uintptr_t)src_boxes | (uintptr_t)tgt_boxes | (uintptr_t)relu_bits) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_forward: boxes / relu_bits must be 16-byte aligned");
    const DeviceGuard guard(out);
    if (guard.status()) return guard.status();
    const dim3 block(32, kRelFwdWarps);
    const dim3 grid((N2 + 31) / 32, (N1 + kFwdRowsPerCta - 1) / kFwdRowsPerCta, B);
    if (grid.y > 65535) return fail(RDETR_ERR_UNSUPPORTED, "rdetr_relation_forward: N1=%d too large", N1);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (flags == RDETR_REL_FAST) {
        const float *ts = nullptr, *tt = nullptr;
        if (int rc = prepare_tables("rdetr_relation_forward", src_boxes, tgt_boxes, dim_t, scale, eps, B, N1, N2, workspace,
                                    workspace_bytes, st, &ts, &tt))
            return rc;
        return rel_fast_rows_per_iter() == 8 ? launch_rel_fwd_fast<4>(ts, tt, weight, bias, dim_t, scale, attn_mask, out, relu_bits, B, N1, N2, st)
                                             : launch_rel_fwd_fast<2>(ts, tt, weight, bias, dim_t, scale, attn_mask, out, relu_bits, B, N1, N2, st);
    } else {
        rel_fwd_kernel<false><<<grid, block, 0, st>>>(src_boxes, tgt_boxes, nullptr, nullptr, weight, bias, dim_t, scale, eps,
                                                      attn_mask, out, relu_bits, N1, N2);
    }
    return check_cuda(cudaGetLastError(), "rel_fwd_kernel launch");
}

extern "C" int rdetr_relation_backward(const float *src_boxes, const float *tgt_boxes, const float *dim_t, float scale,
                                       float eps, const float *grad_out, const uint32_t *relu_bits, float *grad_weight,
                                       float *grad_bias, int B, int N1, int N2, int H, int flags, void *workspace,
                                       size_t workspace_bytes, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_rel("rdetr_relation_backward", B, N1, N2, H, flags)) return rc;
    if (!grad_weight || !grad_bias) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_backward: null gradient buffer");
    const DeviceGuard guard(grad_weight);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = check_cuda(cudaMemsetAsync(grad_weight, 0, sizeof(float) * kRelHeads * kRelFeat, st), "cudaMemsetAsync(grad_weight)")) return rc;
    if (int rc = check_cuda(cudaMemsetAsync(grad_bias, 0, sizeof(float) * kRelHeads, st), "cudaMemsetAsync(grad_bias)")) return rc;
    if (B == 0 || N1 == 0 || N2 == 0) return RDETR_OK;
    if (!src_boxes || !tgt_boxes || !dim_t || !grad_out || !relu_bits)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_backward: null pointer argument");
    if (((uintptr_t)src_boxes | (uintptr_t)tgt_boxes) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_relation_backward: boxes must be 16-byte aligned");
    const dim3 block(32, kRelBwdWarps);
    const dim3 grid((N2 + 31) / 32, (N1 + kBwdRowsPerCta - 1) / kBwdRowsPerCta, B);
    if (grid.y > 65535) return fail(RDETR_ERR_UNSUPPORTED, "rdetr_relation_backward: N1=%d too large", N1);
    if (flags == RDETR_REL_FAST) {
        const float *ts = nullptr, *tt = nullptr;
        if (int rc = prepare_tables("rdetr_relation_backward", src_boxes, tgt_boxes, dim_t, scale, eps, B, N1, N2, workspace,
                                    workspace_bytes, st, &ts, &tt))
            return rc;
        return launch_rel_bwd_fast(src_boxes, tgt_boxes, ts, tt, dim_t, scale, eps, grad_out, relu_bits, grad_weight, grad_bias, B, N1, N2, st);
    } else {
        rel_bwd_kernel<false><<<grid, block, 0, st>>>(src_boxes, tgt_boxes, nullptr, nullptr, dim_t, scale, eps, grad_out,
                                                      relu_bits, grad_weight, grad_bias, N1, N2);
    }
    return check_cuda(cudaGetLastError(), "rel_bwd_kernel launch");
}
