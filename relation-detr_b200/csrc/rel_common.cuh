// rel_common.cuh -- pieces shared by the relation-bias kernels (rel.cu) and the fused relation attention (rel_attn.cu).
#pragma once

#include "common.cuh"

namespace rdetr {

constexpr int kRelHeads = 8;
constexpr int kRelK = 8;             // frequencies per box feature
constexpr int kRelFeat = 8 * kRelK;  // 4 features x K x (sin, cos) = 64
constexpr int kTab = 36;             // floats per box in the FAST tables (144 B, 16-byte aligned)
// table row: [0]=cx [1]=cy [2]=1/(w+eps) [3]=1/(h+eps) [4..11]=sin(Aw_k) [12..19]=sin(Ah_k)
//            [20..27]=cos(Aw_k) [28..35]=cos(Ah_k),  A*_k = log(size+eps) * scale / dim_t[k]

__global__ void rel_tables_kernel(const float *__restrict__ boxes, const float *__restrict__ dim_t, float scale, float eps,
                                  float *__restrict__ table, int nboxes);  // rel.cu

// sin/cos of theta = es / d for the centre features.
// EXACT: IEEE division + accurate sincosf, as torch evaluates it.
// FAST: the angle is never materialised in fp32 (theta reaches ~1e3 rad, ulp 6e-5).  With c = 1/(2 pi d)
// split into chi + clo, t = es*chi is the angle in revolutions; n = rint(t); the fractional part
// f = fma(es, chi, -n) + es*clo is exact to one rounding at |f| <= 0.5, and sin/cos(2 pi f) go to MUFU.
template <bool FAST>
__device__ __forceinline__ void angle_sincos(float es, float d_or_chi, float invd_or_clo, float &sn, float &cs)
{
    if constexpr (!FAST) {
        const float th = es / d_or_chi;  // IEEE division: nvcc emits div.rn.f32 without -use_fast_math
        sincosf(th, &sn, &cs);
    } else {
        const float n = rintf(es * d_or_chi);
        float f = fmaf(es, d_or_chi, -n);
        f = fmaf(es, invd_or_clo, f);
        const float x = f * 6.283185307179586f;
        sn = __sinf(x);
        cs = __cosf(x);
    }
}

// hi/lo split of 1 / (2 pi d) for the FAST angle evaluation
__device__ __forceinline__ void rev_constants(float d, float &chi, float &clo)
{
    const double c = 1.0 / (6.283185307179586476925 * (double)d);
    chi = (float)c;
    clo = (float)(c - (double)chi);
}

// (A MUFU.LG2-based log was measured too: 3 % faster, but it adds ~5e-6 of error at the k = 0 frequency and
// costs FAST its accuracy advantage over the reference's own fp32 evaluation, so logf stays accurate.)

// Packed fp32 pairs (Blackwell fma.rn.f32x2): one issue slot performs two IEEE FMAs.  The forward is
// issue bound (ncu: 73 % issue-active, 50 % FMA pipe), so halving the instruction count of the
// 64 -> 8 projection is worth more than anything else; per-element results are unchanged.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi)
{
    asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ void ffma2(f32x2 &acc, f32x2 a, f32x2 b)
{
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}

__device__ __forceinline__ f32x2 fmul2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 ffma2r(f32x2 a, f32x2 b, f32x2 c)   // a * b + c
{
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ f32x2 fadd2(f32x2 a, f32x2 b)
{
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ f32x2 ld2(const float2 *p) { return *reinterpret_cast<const f32x2 *>(p); }

// ---- FAST arithmetic on ROW PAIRS (round 2) ----------------------------------------------------------------------------
// The forward kernels are issue bound, so everything that is the same operation on two src rows runs as ONE packed
// fp32x2 instruction (mul / fma.rn.f32x2: two IEEE operations, per-element results unchanged): the Cody-Waite angle
// reduction of the centre features, the angle-difference identities of the size features, and the projection.
// Shared-memory layouts:
//   row2[p][f]  = { table[2p][f], table[2p + 1][f] }   the FAST table rows (kTab fields) of two consecutive src boxes
//   tgt[f][col] =   table[col][f]                      the table rows of the CTA's 32 tgt boxes, transposed
struct RelConsts {                       // per-CTA constants
    float2 wt[kRelFeat][kRelHeads];      // projection weights transposed and duplicated: [n][h] = {w, w}
    float2 chi2[kRelK], nchi2[kRelK], clo2[kRelK];   // {c, c}, {-c, -c}: hi part of 1 / (2 pi dim_t[k]); {lo, lo}
    float bias[kRelHeads];
};

__device__ __forceinline__ void rel_consts_setup(RelConsts &rc, const float *__restrict__ weight, const float *__restrict__ bias,
                                                 const float *__restrict__ dim_t, int tid, int nthreads)
{
    for (int idx = tid; idx < kRelFeat * kRelHeads; idx += nthreads) {
        const int h = idx / kRelFeat, n = idx - h * kRelFeat;
        const float wv = weight[idx];
        rc.wt[n][h] = make_float2(wv, wv);
    }
    if (tid < kRelHeads) rc.bias[tid] = bias[tid];
    if (tid < kRelK) {
        float chi, clo;
        rev_constants(dim_t[tid], chi, clo);
        rc.chi2[tid] = make_float2(chi, chi);
        rc.nchi2[tid] = make_float2(-chi, -chi);
        rc.clo2[tid] = make_float2(clo, clo);
    }
}

// rows [0, nvalid) of `rows` (kTab floats each) -> row2[0 .. nrows / 2); rows past nvalid hold 1.0 (harmless geometry)
__device__ __forceinline__ void stage_row_pairs(float2 *row2, const float *__restrict__ rows, int nvalid, int nrows, int tid, int nthreads)
{
    float *flat = reinterpret_cast<float *>(row2);
    for (int idx = tid; idx < nrows * kTab; idx += nthreads) {
        const int r = idx / kTab, f = idx - r * kTab;
        flat[((r >> 1) * kTab + f) * 2 + (r & 1)] = r < nvalid ? rows[idx] : 1.0f;
    }
}

// {sin, sin} and {cos, cos} of the angles es / d_k of two rows: angle_sincos<true> on both halves, bit for bit.
// rint() is two packed additions of 1.5 * 2^23 (round-to-nearest-even at integer granularity, exact for |t| < 2^22; here
// |t| = |es * chi| < 200) instead of two FRNDs on the 4-lane XU pipe, and it is applied to -t: rint(-t) = -rint(t).
__device__ __forceinline__ void angle_sincos2(f32x2 es2, f32x2 chi2, f32x2 nchi2, f32x2 clo2, f32x2 twopi2, f32x2 magic2, f32x2 nmagic2,
                                              f32x2 &sn2, f32x2 &cs2)
{
    const f32x2 nn2 = fadd2(fadd2(fmul2(es2, nchi2), magic2), nmagic2);   // -rint(es * chi)
    f32x2 f = ffma2r(es2, chi2, nn2);
    f = ffma2r(es2, clo2, f);
    float a, b;
    unpack2(fmul2(f, twopi2), a, b);
    sn2 = pack2(__sinf(a), __sinf(b));
    cs2 = pack2(__cosf(a), __cosf(b));
}

// acc[p][h] += W[h][n] * fs[p] + W[h][n + 1] * fc[p] (sin term first), fs / fc = the feature of rows (2p, 2p + 1)
template <int RP>
__device__ __forceinline__ void project2(const float2 (*wt)[kRelHeads], int n, const f32x2 (&fs)[RP], const f32x2 (&fc)[RP],
                                         f32x2 (&acc)[RP][kRelHeads])
{
    const ulonglong2 *ws = reinterpret_cast<const ulonglong2 *>(&wt[n][0]);
    const ulonglong2 *wc = reinterpret_cast<const ulonglong2 *>(&wt[n + 1][0]);
#pragma unroll
    for (int q = 0; q < kRelHeads / 2; ++q) {
        const ulonglong2 w = ws[q];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            ffma2(acc[p][2 * q], w.x, fs[p]);
            ffma2(acc[p][2 * q + 1], w.y, fs[p]);
        }
    }
#pragma unroll
    for (int q = 0; q < kRelHeads / 2; ++q) {
        const ulonglong2 w = wc[q];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            ffma2(acc[p][2 * q], w.x, fc[p]);
            ffma2(acc[p][2 * q + 1], w.y, fc[p]);
        }
    }
}

// Pre-activation of the relation bias for the 2 * RP src rows whose pairs start at `row2` and tgt column `lane`, all heads:
// acc[p][h] = (row 2p, row 2p + 1) of head h.  Same per-element operations as the scalar FAST path of round 1; the 64 terms
// are summed in the order (x, w, y, h) x k x (sin, cos).
template <int RP>
__device__ __forceinline__ void fast_bias_rows(const RelConsts &rc, const float2 *row2, const float (*tgt)[32], float scale, int lane,
                                               f32x2 (&acc)[RP][kRelHeads])
{
#pragma unroll
    for (int p = 0; p < RP; ++p)
#pragma unroll
        for (int h = 0; h < kRelHeads; ++h) acc[p][h] = pack2(rc.bias[h], rc.bias[h]);
    const f32x2 twopi2 = pack2(6.283185307179586f, 6.283185307179586f);
    const f32x2 magic2 = pack2(12582912.f, 12582912.f), nmagic2 = pack2(-12582912.f, -12582912.f);
    // One loop over (x / w, k) and (y / h, k): its body evaluates one CENTRE frequency (XU pipe: 4 MUFU per row pair) and one
    // SIZE frequency (FMA pipe only) and projects both, so that every warp always has work for both pipes in flight.
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
        const float t_xy = tgt[c][lane];
        f32x2 es2[RP];
#pragma unroll
        for (int p = 0; p < RP; ++p) {
            const float2 xy = row2[p * kTab + c], inv = row2[p * kTab + 2 + c];
            const float e0 = logf(fmaf(fabsf(xy.x - t_xy), inv.x, 1.0f)) * scale;  // (x * scale) first, position_encoding.py:133
            const float e1 = logf(fmaf(fabsf(xy.y - t_xy), inv.y, 1.0f)) * scale;
            es2[p] = pack2(e0, e1);
        }
#pragma unroll 2
        for (int k = 0; k < kRelK; ++k) {
            f32x2 fs[RP], fc[RP];
            {   // centre feature c, frequency k
                const f32x2 chi2 = ld2(&rc.chi2[k]), nchi2 = ld2(&rc.nchi2[k]), clo2 = ld2(&rc.clo2[k]);
#pragma unroll
                for (int p = 0; p < RP; ++p) angle_sincos2(es2[p], chi2, nchi2, clo2, twopi2, magic2, nmagic2, fs[p], fc[p]);
                project2<RP>(rc.wt, c * 2 * kRelK + 2 * k, fs, fc, acc);
            }
            {   // size feature 2 + c, frequency k: per-box tables + angle-difference identities
                const float sB = tgt[4 + c * 8 + k][lane], cB = tgt[20 + c * 8 + k][lane];
                const f32x2 sB2 = pack2(sB, sB), nsB2 = pack2(-sB, -sB), cB2 = pack2(cB, cB);
#pragma unroll
                for (int p = 0; p < RP; ++p) {
                    const f32x2 sA2 = ld2(&row2[p * kTab + 4 + c * 8 + k]), cA2 = ld2(&row2[p * kTab + 20 + c * 8 + k]);
                    fs[p] = ffma2r(sA2, cB2, fmul2(cA2, nsB2));  // sin(A - B) = sA cB - cA sB
                    fc[p] = ffma2r(cA2, cB2, fmul2(sA2, sB2));   // cos(A - B) = cA cB + sA sB
                }
                project2<RP>(rc.wt, (2 + c) * 2 * kRelK + 2 * k, fs, fc, acc);
            }
        }
    }
}

// acc[p][h] = (row 2p, row 2p+1) of head h;  acc += W[h][n] * sn + W[h][n+1] * cs, sin term first.
// Weights sit in shared memory transposed and duplicated: s_wt2[n][h] = {w, w}.
template <int R>
__device__ __forceinline__ void project(const float2 (*s_wt2)[kRelHeads], int n, const float (&sn)[R], const float (&cs)[R],
                                        f32x2 (&acc)[R / 2][kRelHeads])
{
    f32x2 fs[R / 2], fc[R / 2];
#pragma unroll
    for (int p = 0; p < R / 2; ++p) {
        fs[p] = pack2(sn[2 * p], sn[2 * p + 1]);
        fc[p] = pack2(cs[2 * p], cs[2 * p + 1]);
    }
    const ulonglong2 *ws = reinterpret_cast<const ulonglong2 *>(&s_wt2[n][0]);
    const ulonglong2 *wc = reinterpret_cast<const ulonglong2 *>(&s_wt2[n + 1][0]);
#pragma unroll
    for (int q = 0; q < kRelHeads / 2; ++q) {
        const ulonglong2 w = ws[q];
#pragma unroll
        for (int p = 0; p < R / 2; ++p) {
            ffma2(acc[p][2 * q], w.x, fs[p]);
            ffma2(acc[p][2 * q + 1], w.y, fs[p]);
        }
    }
#pragma unroll
    for (int q = 0; q < kRelHeads / 2; ++q) {
        const ulonglong2 w = wc[q];
#pragma unroll
        for (int p = 0; p < R / 2; ++p) {
            ffma2(acc[p][2 * q], w.x, fc[p]);
            ffma2(acc[p][2 * q + 1], w.y, fc[p]);
        }
    }
}

// Scalar twin of project() (plain FFMA, accumulators as float): selected with -DRDETR_REL_SCALAR_FMA for
// A/B measurements of the packed path.
template <int R>
__device__ __forceinline__ void project_scalar(const float2 (*s_wt2)[kRelHeads], int n, const float (&sn)[R], const float (&cs)[R],
                                               float (&acc)[R][kRelHeads])
{
#pragma unroll
    for (int q = 0; q < kRelHeads / 2; ++q) {
        const float4 ws = *reinterpret_cast<const float4 *>(&s_wt2[n][2 * q]);      // {w_h, w_h, w_h+1, w_h+1}
        const float4 wc = *reinterpret_cast<const float4 *>(&s_wt2[n + 1][2 * q]);
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r][2 * q] = fmaf(ws.x, sn[r], acc[r][2 * q]);
            acc[r][2 * q + 1] = fmaf(ws.z, sn[r], acc[r][2 * q + 1]);
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            acc[r][2 * q] = fmaf(wc.x, cs[r], acc[r][2 * q]);
            acc[r][2 * q + 1] = fmaf(wc.z, cs[r], acc[r][2 * q + 1]);
        }
    }
}


// FAST mode: fills the per-box tables in `workspace` ([B*N1 + B*N2] rows of kTab floats); rel.cu
int prepare_tables(const char *who, const float *src, const float *tgt, const float *dim_t, float scale, float eps, int B, int N1,
                   int N2, void *workspace, size_t workspace_bytes, cudaStream_t st, const float **src_tab, const float **tgt_tab);

int launch_rel_bwd_fast(const float *src, const float *tgt, const float *src_tab, const float *tgt_tab, const float *dim_t,
                        float scale, float eps, const float *grad, const uint32_t *relu_bits, float *grad_weight,
                        float *grad_bias, int B, int N1, int N2, cudaStream_t st);  // rel.cu

}  // namespace rdetr
