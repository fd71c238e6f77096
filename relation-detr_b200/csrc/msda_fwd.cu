// msda_fwd.cu -- multi-scale deformable attention forward for sm_100a.
//
// Replaces ms_deformable_im2col_gpu_kernel (upstream models/bricks/ops/cuda/ms_deform_im2col_cuda.cuh:226-288),
// which uses one thread per output scalar and makes all 32 channel threads of a (b,q,m) re-read the
// same sampling location / weight and redo the same coordinate arithmetic.
//
// Layout of the work here:
//   * a "pair" is one (b, q, m): its L*P samples are contiguous in sampling_locations and
//     attention_weights, its D outputs are contiguous in `out`, and consecutive pairs are
//     contiguous too, so a CTA that owns kPairs consecutive pairs (8 for fp32, 16 for bf16: 64 threads)
//     streams three dense blocks;
//   * phase 1 (one thread per sample): coalesced streaming loads of loc/attn, coordinate
//     arithmetic done ONCE per sample, corner pixel indices + attention-scaled bilinear weights
//     staged in shared memory (32 B per sample);
//   * phase 2 (16 B of channels per lane): kLanes = D*sizeof(VT)/16 lanes own one pair (8 lanes for
//     fp32, 4 for bf16); every corner is one 16-byte read-only load per lane, i.e. a full 128-byte
//     (fp32) or 64-byte (bf16) contiguous row per pair, 4 corners x 4 points in flight per lane;
//     fp32 accumulation in registers, one 16-byte store per lane.
// No tensor cores: the op is a gather, not a contraction.
#include <cstdlib>

#include "common.cuh"

namespace rdetr {

// THREADS: CTA size.  Small CTAs win: the kernel is latency bound (ncu: long-scoreboard stalls dominate, the
// L1 data pipe is 61 % busy), and with 64 threads the prologue / barrier of one CTA overlaps the gathers
// of the others resident on the SM (tools/tune_fwd.py: 0.68 -> 0.61 ms at configs[1]; capping fp32 at 32
// registers for full occupancy gives 0.57 ms).  MINB = 0 leaves the register budget to ptxas.
// LEAN (default): phase 1 stages ELEMENT OFFSETS (pixel index * M*D) instead of pixel indices, so that a corner's
// address is one IMAD.WIDE (the 64-bit multiply-add per corner was 5 instructions: 79 of the 225 of four samples),
// and a warp whose four corners are all inside the level -- decided by one vote -- gathers without predicates
// and without zero-filling the 16 landing registers (31 CS2R + 16 ISETP per four samples).  LEAN = false is the
// round-1 loop, kept for tuning builds (tools/tune_fwd.py).
template <typename VT, int CH, int D, typename IO, int THREADS, int MINB = 0, bool LEAN = true>
__global__ void __launch_bounds__(THREADS, MINB)
msda_fwd_kernel(const VT *__restrict__ value, const int64_t *__restrict__ spatial_shapes,
                const int64_t *__restrict__ level_start_index, const IO io, VT *__restrict__ out, int S, int M, int L, int Nq,
                int P, long long total_pairs)
{
    using SL = Slice<VT, CH>;
    constexpr int kCh = CH;
    constexpr int kLanes = D / kCh;
    constexpr int kFwdThreads = THREADS;
    constexpr int kPairs = kFwdThreads / kLanes;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_H[kMaxLevels], s_W[kMaxLevels], s_start[kMaxLevels];
    __shared__ float s_invW[kMaxLevels], s_invH[kMaxLevels];
    __shared__ unsigned char s_lvl[kMaxLevels * kMaxPoints];  // level of sample slot lp (= lp / P)
    __shared__ long long s_bq[THREADS / (D / CH)];  // FusedIO: b*Nq + q of every pair (64-bit division once per pair, not per sample)
    __shared__ int s_b[THREADS / (D / CH)];
    const float inv_P = 1.0f / (float)P;

    const int LP = L * P;
    const int stride = LP + 1;  // one slot of padding de-phases consecutive pairs across banks
    int4 *s_pix = reinterpret_cast<int4 *>(smem_raw);              // [kPairs][stride]
    float4 *s_wgt = reinterpret_cast<float4 *>(s_pix + kPairs * stride);  // [kPairs][stride]

    if (threadIdx.x < L) {
        s_H[threadIdx.x] = (int)spatial_shapes[2 * threadIdx.x];
        s_W[threadIdx.x] = (int)spatial_shapes[2 * threadIdx.x + 1];
        s_start[threadIdx.x] = (int)level_start_index[threadIdx.x];
        s_invW[threadIdx.x] = 1.0f / (float)s_W[threadIdx.x];
        s_invH[threadIdx.x] = 1.0f / (float)s_H[threadIdx.x];
    }
    for (int i = threadIdx.x; i < L * P; i += THREADS) s_lvl[i] = (unsigned char)(i / P);
    __syncthreads();

    const long long pair0 = (long long)blockIdx.x * kPairs;
    const long long left = total_pairs - pair0;
    const int npairs = left < kPairs ? (int)left : kPairs;

    // ---- phase 1: one thread per sample ----------------------------------------------------------
    const int nsamples = npairs * LP;
    float2 *s_stat = reinterpret_cast<float2 *>(s_wgt + kPairs * stride);  // FusedIO: per-pair (max, sum) of the softmax
    if constexpr (IO::kFused) fused_softmax_stats<kLanes>(io, pair0, npairs, LP, M, Nq, s_wgt, stride, s_stat, s_bq, s_b);
    for (SampleWalk sw(threadIdx.x, kFwdThreads, LP); sw.s < nsamples; sw.next(kFwdThreads)) {
        const int pair = sw.pair, lp = sw.lp, l = s_lvl[lp];
        float a;
        const Tap t = sample_tap(io, pair0, sw.s, pair, lp, l, LP, L, S, s_wgt[pair * stride + lp].x, s_stat, s_bq, s_b, s_H, s_W,
                                 s_start, s_invW, s_invH, inv_P, a);
        const float hh = 1.f - t.lh, hw = 1.f - t.lw;
        if constexpr (LEAN) {
            const int ps = M * D;  // S*M*D < 2^31 (validate_msda): the offsets fit an int
            s_pix[pair * stride + lp] = make_int4(t.pix[0] >= 0 ? t.pix[0] * ps : -1, t.pix[1] >= 0 ? t.pix[1] * ps : -1,
                                                  t.pix[2] >= 0 ? t.pix[2] * ps : -1, t.pix[3] >= 0 ? t.pix[3] * ps : -1);
        } else {
            s_pix[pair * stride + lp] = make_int4(t.pix[0], t.pix[1], t.pix[2], t.pix[3]);
        }
        s_wgt[pair * stride + lp] = make_float4(a * (hh * hw), a * (hh * t.lw), a * (t.lh * hw), a * (t.lh * t.lw));
    }
    if constexpr (LEAN) {  // pairs past the end (last CTA only): all corners invalid, so phase 2 needs no per-lane guard
        for (int s = nsamples + threadIdx.x; s < kPairs * LP; s += kFwdThreads) {
            s_pix[(s / LP) * stride + s % LP] = make_int4(-1, -1, -1, -1);
            s_wgt[(s / LP) * stride + s % LP] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    __syncthreads();

    // ---- phase 2: gather ---------------------------------------------------------------------------
    const int pair = threadIdx.x / kLanes;
    const int lane = threadIdx.x - pair * kLanes;
    if constexpr (!LEAN) {
        if (pair >= npairs) return;
    }
    const bool active = pair < npairs;
    const long long gp = pair0 + (active ? pair : 0);
    const int m = (int)(gp % M);
    const long long b = (gp / M) / Nq;
    const VT *vbase = value + (b * S * M + m) * (long long)D + lane * kCh;
    const int pix_stride = M * D;

    float acc[kCh];
#pragma unroll
    for (int c = 0; c < kCh; ++c) acc[c] = 0.f;

    const int4 *my_pix = s_pix + pair * stride;
    const float4 *my_wgt = s_wgt + pair * stride;
    if constexpr (LEAN) {
#pragma unroll 4
        for (int lp = 0; lp < LP; ++lp) {
            const int4 px = my_pix[lp];
            const float4 w = my_wgt[lp];
            float v0[kCh], v1[kCh], v2[kCh], v3[kCh];
            if (__all_sync(0xffffffffu, (px.x | px.y | px.z | px.w) >= 0)) {
                SL::load(elem_ptr(vbase, px.x), v0);
                SL::load(elem_ptr(vbase, px.y), v1);
                SL::load(elem_ptr(vbase, px.z), v2);
                SL::load(elem_ptr(vbase, px.w), v3);
            } else {
#pragma unroll
                for (int c = 0; c < kCh; ++c) v0[c] = v1[c] = v2[c] = v3[c] = 0.f;
                if (px.x >= 0) SL::load(elem_ptr(vbase, px.x), v0);
                if (px.y >= 0) SL::load(elem_ptr(vbase, px.y), v1);
                if (px.z >= 0) SL::load(elem_ptr(vbase, px.z), v2);
                if (px.w >= 0) SL::load(elem_ptr(vbase, px.w), v3);
            }
#pragma unroll
            for (int c = 0; c < kCh; ++c) {
                acc[c] = fmaf(w.x, v0[c], acc[c]);
                acc[c] = fmaf(w.y, v1[c], acc[c]);
                acc[c] = fmaf(w.z, v2[c], acc[c]);
                acc[c] = fmaf(w.w, v3[c], acc[c]);
            }
        }
        if (!active) return;
    } else {
#pragma unroll 4
    for (int lp = 0; lp < LP; ++lp) {
        const int4 px = my_pix[lp];
        const float4 w = my_wgt[lp];
        float v0[kCh], v1[kCh], v2[kCh], v3[kCh];
#pragma unroll
        for (int c = 0; c < kCh; ++c) v0[c] = v1[c] = v2[c] = v3[c] = 0.f;
        if (px.x >= 0) SL::load(vbase + (long long)px.x * pix_stride, v0);
        if (px.y >= 0) SL::load(vbase + (long long)px.y * pix_stride, v1);
        if (px.z >= 0) SL::load(vbase + (long long)px.z * pix_stride, v2);
        if (px.w >= 0) SL::load(vbase + (long long)px.w * pix_stride, v3);
#pragma unroll
        for (int c = 0; c < kCh; ++c) {
            acc[c] = fmaf(w.x, v0[c], acc[c]);
            acc[c] = fmaf(w.y, v1[c], acc[c]);
            acc[c] = fmaf(w.z, v2[c], acc[c]);
            acc[c] = fmaf(w.w, v3[c], acc[c]);
        }
    }
    }
    SL::store(out + gp * D + lane * kCh, acc);
}

// msda_fwd_tile.cu: tiled forward for encoder self-attention (queries = pixels of the pyramid).  Returns -1 when
// (L, P) is outside what it is built for.
int msda_tile_mode();
template <typename VT, int CH, typename IO>
int launch_fwd_tile(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, void *out, int B, int S, int M, int L,
                    int Nq, int P, cudaStream_t stream);

template <typename VT, int CH, typename IO, int THREADS, int MINB = 0, bool LEAN = true>
static int launch_fwd_variant(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, void *out, int B,
                              int S, int M, int L, int Nq, int P, cudaStream_t stream)
{
    constexpr int D = 32;
    constexpr int kLanes = D / CH;
    constexpr int kPairs = THREADS / kLanes;
    const long long total_pairs = (long long)B * Nq * M;
    const size_t smem = (size_t)kPairs * (L * P + 1) * 32 + (IO::kFused ? kPairs * sizeof(float2) : 0);
    auto kern = msda_fwd_kernel<VT, CH, D, IO, THREADS, MINB, LEAN>;
    if (smem > 48 * 1024) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "cudaFuncSetAttribute(msda_fwd)"))
            return rc;
    }
    const long long grid = (total_pairs + kPairs - 1) / kPairs;
    if (grid > 0x7fffffffLL) return fail(RDETR_ERR_UNSUPPORTED, "msda_forward: B*Nq*M too large (%lld pairs)", total_pairs);
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(static_cast<const VT *>(value), shapes, lsi, io, static_cast<VT *>(out), S, M,
                                                    L, Nq, P, total_pairs);
    return check_cuda(cudaGetLastError(), "msda_fwd_kernel launch");
}

template <typename VT, int CH, typename IO>
static int launch_fwd(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, void *out, int B, int S,
                      int M, int L, int Nq, int P, cudaStream_t stream)
{
    // encoder self-attention: tiled kernel with shared-memory value windows (msda_fwd_tile.cu);
    // rdetr_msda_set_tile_mode(1) / RDETR_MSDA_TILE=1 keeps the flat kernel
    if (Nq == S && msda_tile_mode() == 2) {
        const int rc = launch_fwd_tile<VT, CH, IO>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (rc >= 0) return rc;
    }
#ifdef RDETR_TUNE_FWD
    // tuning builds only (tools/tune_fwd.py): CTA size from the environment
    const char *e = getenv("RDETR_MSDA_FWD_VARIANT");
    const int v = e ? atoi(e) : 0;
    if (v == 1) return launch_fwd_variant<VT, CH, IO, 256>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 2) return launch_fwd_variant<VT, CH, IO, 128>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 3) return launch_fwd_variant<VT, CH, IO, 32>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 4) return launch_fwd_variant<VT, CH, IO, 64, 0>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 5) return launch_fwd_variant<VT, CH, IO, 64, 24>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 6) return launch_fwd_variant<VT, CH, IO, 128, 16>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 7) return launch_fwd_variant<VT, CH, IO, 32, 32>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 8) return launch_fwd_variant<VT, CH, IO, 64, sizeof(VT) == 4 ? 32 : 0, false>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 9) return launch_fwd_variant<VT, CH, IO, 64, 0, false>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if constexpr (sizeof(VT) == 4) {
        if (v == 12) return launch_fwd_variant<VT, 2, IO, 64, 32, false>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 13) return launch_fwd_variant<VT, 2, IO, 64, 32, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 14) return launch_fwd_variant<VT, 2, IO, 128, 16, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 15) return launch_fwd_variant<VT, 2, IO, 64, 0, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 16) return launch_fwd_variant<VT, 8, IO, 64, 0, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 17) return launch_fwd_variant<VT, 8, IO, 64, 16, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 18) return launch_fwd_variant<VT, 8, IO, 64, 0, false>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 19) return launch_fwd_variant<VT, 8, IO, 128, 8, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
        if (v == 20) return launch_fwd_variant<VT, 8, IO, 32, 0, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    }
    if (v == 10) return launch_fwd_variant<VT, CH, IO, 64, 24>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    if (v == 11) return launch_fwd_variant<VT, CH, IO, 128, 16>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
#endif
    // fp32: the round-1 loop capped at 32 registers (32 CTAs x 2 warps = full occupancy; 0.62 -> 0.57 ms at configs[1]) as
    // long as one image's value tensor fits L2 comfortably; for the 1200x2000 pyramid (209 MB per image) the extra
    // CTAs in flight only widen the window of lines competing for L2 (0.90 ms uncapped vs 1.15 ms capped).  The lean
    // loop has a third fewer instructions per sample but does not fit 32 registers (spills: 0.99 ms) and at 42-48
    // registers runs at 0.59-0.65 ms: the fp32 forward is bound by the L1 wavefront rate and by the warps in flight,
    // not by issue slots (profiles/r02aa_exp_lean.txt).  bf16 lanes hold 8 channels (71 registers either way): there
    // the lean loop wins (0.520 -> 0.493 ms; 1200x2000: 0.737 -> 0.684 ms).
    const bool fits_l2 = (size_t)S * M * 32 * sizeof(VT) <= (size_t)96 << 20;
    if (sizeof(VT) == 4 && fits_l2)
        return launch_fwd_variant<VT, CH, IO, 64, sizeof(VT) == 4 ? 32 : 0, false>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
    return launch_fwd_variant<VT, CH, IO, 64, 0, true>(value, shapes, lsi, io, out, B, S, M, L, Nq, P, stream);
}

int validate_msda(const char *who, int B, int S, int M, int D, int L, int Nq, int P, int value_dtype)
{
    if (B < 0 || S <= 0 || M <= 0 || D <= 0 || L <= 0 || Nq < 0 || P <= 0)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: non-positive size (B=%d S=%d M=%d D=%d L=%d Nq=%d P=%d)", who, B, S, M,
                    D, L, Nq, P);
    if (D != 32) return fail(RDETR_ERR_UNSUPPORTED, "%s: head dim D=%d unsupported (kernels are built for D=32)", who, D);
    if (L > kMaxLevels || P > kMaxPoints)
        return fail(RDETR_ERR_UNSUPPORTED, "%s: L=%d P=%d unsupported (max %d levels, %d points)", who, L, P, kMaxLevels,
                    kMaxPoints);
    if ((long long)S * M * D >= (1LL << 31))
        return fail(RDETR_ERR_UNSUPPORTED, "%s: S*M*D = %lld does not fit 31 bits", who, (long long)S * M * D);
    if (value_dtype != RDETR_DTYPE_F32 && value_dtype != RDETR_DTYPE_BF16)
        return fail(RDETR_ERR_UNSUPPORTED, "%s: value_dtype %d unsupported (0 = f32, 1 = bf16)", who, value_dtype);
    return RDETR_OK;
}

}  // namespace rdetr

extern "C" int rdetr_msda_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                  const float *sampling_locations, const float *attention_weights, void *out, int B,
                                  int S, int M, int D, int L, int Nq, int P, int value_dtype, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_msda("rdetr_msda_forward", B, S, M, D, L, Nq, P, value_dtype)) return rc;
    if (B == 0 || Nq == 0) return RDETR_OK;
    if (!value || !spatial_shapes || !level_start_index || !sampling_locations || !attention_weights || !out)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_forward: null pointer argument");
    if (((uintptr_t)value | (uintptr_t)out | (uintptr_t)sampling_locations | (uintptr_t)attention_weights) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_forward: value/out/loc/attn must be 16-byte aligned");
    const DeviceGuard guard(value);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PlainIO io{sampling_locations, attention_weights, nullptr, nullptr};
    if (value_dtype == RDETR_DTYPE_F32)
        return launch_fwd<float, 4>(value, spatial_shapes, level_start_index, io, out, B, S, M, L, Nq, P, st);
    return launch_fwd<__nv_bfloat16, 8>(value, spatial_shapes, level_start_index, io, out, B, S, M, L, Nq, P, st);
}

extern "C" int rdetr_msda_fused_forward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                        const float *reference_points, const void *sampling_offsets,
                                        const void *attention_logits, const uint8_t *key_padding_mask, void *out, int B,
                                        int S, int M, int D, int L, int Nq, int P, int ref_dim, int dtype,
                                        rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_msda("rdetr_msda_fused_forward", B, S, M, D, L, Nq, P, dtype)) return rc;
    if (ref_dim != 2 && ref_dim != 4)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_forward: last dim of reference_points must be 2 or 4, got %d", ref_dim);
    if (B == 0 || Nq == 0) return RDETR_OK;
    if (!value || !spatial_shapes || !level_start_index || !reference_points || !sampling_offsets || !attention_logits || !out)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_forward: null pointer argument");
    if (((uintptr_t)value | (uintptr_t)out) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_forward: value/out must be 16-byte aligned");
    const DeviceGuard guard(value);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (dtype == RDETR_DTYPE_F32) {
        const FusedIO<float> io{reference_points, static_cast<const float *>(sampling_offsets),
                                static_cast<const float *>(attention_logits), key_padding_mask, nullptr, nullptr, ref_dim};
        return launch_fwd<float, 4>(value, spatial_shapes, level_start_index, io, out, B, S, M, L, Nq, P, st);
    }
    const FusedIO<__nv_bfloat16> io{reference_points, static_cast<const __nv_bfloat16 *>(sampling_offsets),
                                    static_cast<const __nv_bfloat16 *>(attention_logits), key_padding_mask, nullptr, nullptr, ref_dim};
    return launch_fwd<__nv_bfloat16, 8>(value, spatial_shapes, level_start_index, io, out, B, S, M, L, Nq, P, st);
}
