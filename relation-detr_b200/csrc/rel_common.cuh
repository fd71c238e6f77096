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
