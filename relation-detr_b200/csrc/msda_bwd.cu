// msda_bwd.cu -- multi-scale deformable attention backward for sm_100a.
//
// Replaces ms_deformable_col2im_gpu_kernel_shm_blocksize_aware_reduce_v1<float,32> and its five
// siblings (upstream models/bricks/ops/cuda/ms_deform_im2col_cuda.cuh:290-909): one-warp CTAs, 64
// scalar global atomics per thread and a serial thread-0 reduction per sample.
//
// Same work layout as the forward (see msda_fwd.cu):
//   * phase 1, one thread per sample: corner pixel indices, fractional offsets and the attention
//     weight staged in shared memory (32 B per sample);
//   * phase 2, 4 channels per lane, 8 lanes per (b,q,m) for fp32 AND bf16 (every corner is one request to the L2: the
//     L2 charges per request and per 32-byte sector, profiles/r02al_microbench_red.txt): re-gather the four corners,
//       - grad_value: one vector reduction per lane and corner instead of 32 scalar atomics -- red.global.add.v4.f32
//         (a whole 128-byte fp32 row), or, for the fine levels of a bf16 call, red.global.add.noftz.v2.bf16x2 straight
//         into the bf16 output (a 64-byte row: 1.69x the row rate; rdetr_msda_set_bf16_scatter);
//       - grad_attn / grad_loc: per corner the lane forms <grad_out, value_i> over its channels; the four partials are
//         reduced over the pair's lanes by a transposing butterfly (4 shuffles) and parked in the sample's shared-memory
//         slot; the bilinear algebra on top of them is linear, so it runs once per sample in phase 3;
//   * phase 3: grad_loc / grad_attn leave the CTA as dense, coalesced stores (every element is
//     written, so the caller does not have to zero them; the reference zero-fills and then
//     overwrites, ms_deform_attn_cuda.cu:113-115).
// grad_value accumulates in fp32.  For bf16 value the accumulation target of the coarse levels is an fp32 workspace that
// a second kernel converts (bf16 accumulation would lose the small addends where one address receives ~1e3 updates).
#include <atomic>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace rdetr {

int validate_msda(const char *who, int B, int S, int M, int D, int L, int Nq, int P, int value_dtype);

// msda_bwd_tile.cu: tiled backward for encoder self-attention (queries = pixels of the pyramid).  Returns -1 when
// (L, P) is outside what it is built for.
int msda_tile_mode();
template <typename VT, typename IO>
int launch_bwd_tile(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out, float *gv_f32,
                    int B, int S, int M, int L, int Nq, int P, cudaStream_t stream);

// msda_bwd_coarse.cu: grad_value of the levels that fit shared memory, on a side stream next to the scatter kernel
int msda_coarse_cap_rows(int Nq, int P);
template <typename VT, typename IO>
int fork_coarse(const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out, float *gv_f32, int B, int S, int M,
                int L, int Nq, int P, cudaStream_t main, cudaStream_t *side_out, cudaEvent_t *done_out);
int join_coarse(cudaStream_t main, cudaEvent_t done);

// bf16 backward: average updates per grad_value row up to which a level is scattered with packed bf16x2 reductions
// straight into grad_value (0: never; rdetr_msda_set_bf16_scatter / RDETR_MSDA_BF16_SCATTER)
static std::atomic<int> g_bf16_scatter{-1};
static int msda_bf16_scatter_max_updates()
{
    int v = g_bf16_scatter.load(std::memory_order_relaxed);
    if (v < 0) {
        const char *e = getenv("RDETR_MSDA_BF16_SCATTER");
        v = e ? atoi(e) : kBf16ScatterDefault;
        if (v < 0) v = 0;
        g_bf16_scatter.store(v, std::memory_order_relaxed);
    }
    return v;
}

template <typename VT, int CH, int D, typename IO, int THREADS, int MINB = 0>
__global__ void __launch_bounds__(THREADS, MINB)
msda_bwd_kernel(const VT *__restrict__ value, const int64_t *__restrict__ spatial_shapes,
                const int64_t *__restrict__ level_start_index, const IO io, const VT *__restrict__ grad_out,
                float *__restrict__ grad_value_f32, int S, int M, int L, int Nq, int P, long long total_pairs,
                int coarse_cap_rows, VT *__restrict__ grad_value_direct, int direct_max_updates)
{
    using SL = Slice<VT, CH>;
    constexpr int kCh = CH;
    constexpr int kLanes = D / kCh;
    constexpr int kBwdThreads = THREADS;
    constexpr int kPairs = kBwdThreads / kLanes;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_H[kMaxLevels], s_W[kMaxLevels], s_start[kMaxLevels];
    __shared__ float s_invW[kMaxLevels], s_invH[kMaxLevels];
    __shared__ unsigned char s_lvl[kMaxLevels * kMaxPoints];  // level of sample slot lp (= lp / P)
    __shared__ unsigned char s_scatter[kMaxLevels];           // 0: the level's grad_value is accumulated by msda_bwd_coarse_kernel,
                                                              // 1: fp32 reductions, 2: bf16x2 reductions straight into grad_value
    __shared__ long long s_bq[THREADS / (D / CH)];  // FusedIO: b*Nq + q of every pair (64-bit division once per pair, not per sample)
    __shared__ int s_b[THREADS / (D / CH)];
    const float inv_P = 1.0f / (float)P;

    const int LP = L * P;
    const int stride = LP + 1;
    int4 *s_pix = reinterpret_cast<int4 *>(smem_raw);                      // [kPairs][stride]
    float4 *s_meta = reinterpret_cast<float4 *>(s_pix + kPairs * stride);  // in: (lw, lh, attn, -)  out: (gx, gy, ga, -)

    if (threadIdx.x < L) {
        s_H[threadIdx.x] = (int)spatial_shapes[2 * threadIdx.x];
        s_W[threadIdx.x] = (int)spatial_shapes[2 * threadIdx.x + 1];
        s_start[threadIdx.x] = (int)level_start_index[threadIdx.x];
        s_invW[threadIdx.x] = 1.0f / (float)s_W[threadIdx.x];
        s_invH[threadIdx.x] = 1.0f / (float)s_H[threadIdx.x];
        int mode = !coarse_level(s_H[threadIdx.x], s_W[threadIdx.x], s_start[threadIdx.x], S, coarse_cap_rows);
        if (sizeof(VT) == 2 && mode && grad_value_direct != nullptr &&
            direct_bf16_level(s_H[threadIdx.x], s_W[threadIdx.x], Nq, P, direct_max_updates, S))
            mode = 2;
        s_scatter[threadIdx.x] = (unsigned char)mode;
    }
    for (int i = threadIdx.x; i < L * P; i += THREADS) s_lvl[i] = (unsigned char)(i / P);
    __syncthreads();

    const long long pair0 = (long long)blockIdx.x * kPairs;
    const long long left = total_pairs - pair0;
    const int npairs = left < kPairs ? (int)left : kPairs;
    const int nsamples = npairs * LP;

    // ---- phase 1 -----------------------------------------------------------------------------------
    float2 *s_stat = reinterpret_cast<float2 *>(s_meta + kPairs * stride);  // FusedIO: per-pair softmax statistics
    if constexpr (IO::kFused) fused_softmax_stats<kLanes>(io, pair0, npairs, LP, M, Nq, s_meta, stride, s_stat, s_bq, s_b);
    for (SampleWalk sw(threadIdx.x, kBwdThreads, LP); sw.s < nsamples; sw.next(kBwdThreads)) {
        const int pair = sw.pair, lp = sw.lp, l = s_lvl[lp];
        float a;
        const Tap t = sample_tap(io, pair0, sw.s, pair, lp, l, LP, L, S, s_meta[pair * stride + lp].x, s_stat, s_bq, s_b, s_H, s_W,
                                 s_start, s_invW, s_invH, inv_P, a);
        const int ps = M * D;  // element offsets (S*M*D < 2^31, validate_msda): a corner's address is one multiply-add
        s_pix[pair * stride + lp] = make_int4(t.pix[0] >= 0 ? t.pix[0] * ps : -1, t.pix[1] >= 0 ? t.pix[1] * ps : -1,
                                              t.pix[2] >= 0 ? t.pix[2] * ps : -1, t.pix[3] >= 0 ? t.pix[3] * ps : -1);
        s_meta[pair * stride + lp] = make_float4(t.lw, t.lh, a, 0.f);
    }
    // pairs past the end (last CTA only): all corners invalid, so phase 2 needs no per-lane guard
    for (int s = nsamples + threadIdx.x; s < kPairs * LP; s += kBwdThreads) {
        s_pix[(s / LP) * stride + s % LP] = make_int4(-1, -1, -1, -1);
        s_meta[(s / LP) * stride + s % LP] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    // ---- phase 2 -----------------------------------------------------------------------------------
    // Per corner i the lane forms d_i = <grad_out, value_i> over its channels; the bilinear algebra on top of the four
    // d_i (cuh:110-141 of the reference: d/d(w_im), d/d(h_im), d/d(attn)) is linear in them, so it moves behind the
    // cross-lane reduction and into phase 3, where one THREAD does it per sample instead of one WARP per four pairs.
    // The four partials are reduced over the pair's 8 lanes by a transposing butterfly (4 shuffles instead of 9).
    const int pair = threadIdx.x / kLanes;
    const int lane = threadIdx.x - pair * kLanes;
    const bool active = pair < npairs;
    const long long gp = pair0 + (active ? pair : 0);
    const int m = (int)(gp % M);
    const long long b = (gp / M) / Nq;
    const long long batch_off = (b * S * M + m) * (long long)D + lane * kCh;
    const VT *vbase = value + batch_off;
    float *gvbase = grad_value_f32 + batch_off;
    VT *gdbase = grad_value_direct + batch_off;  // only dereferenced for levels in mode 2

    float g[kCh];
#pragma unroll
    for (int c = 0; c < kCh; ++c) g[c] = 0.f;
    if (active) SL::load_stream(grad_out + gp * D + lane * kCh, g);

    const int4 *my_pix = s_pix + pair * stride;
    float4 *my_meta = s_meta + pair * stride;
    const bool odd1 = lane & 1, odd2 = lane & 2;
    const int dot_slot = ((lane & 1) << 1) | ((lane >> 1) & 1);  // lanes 0..3 of the pair end up with d0, d2, d1, d3
    // one level's P samples; MODE (compile time) = where the level's grad_value goes: 0 nowhere (msda_bwd_coarse_kernel
    // accumulates it), 1 fp32 reductions into grad_value_f32, 2 packed bf16x2 reductions straight into grad_value
    auto level_loop = [&](auto mode_c, int lp) {
        constexpr int scatter = decltype(mode_c)::value;
#pragma unroll 2
        for (int p = 0; p < P; ++p, ++lp) {
            const int4 px = my_pix[lp];
            const float4 mt = my_meta[lp];
            const float lw = mt.x, lh = mt.y, a = mt.z;
            const float hw = 1.f - lw, hh = 1.f - lh;
            const float w0 = (hh * hw) * a, w1 = (hh * lw) * a, w2 = (lh * hw) * a, w3 = (lh * lw) * a;
            float v0[kCh], v1[kCh], v2[kCh], v3[kCh];
            if (__all_sync(0xffffffffu, (px.x | px.y | px.z | px.w) >= 0)) {
                SL::load(elem_ptr(vbase, px.x), v0);
                SL::load(elem_ptr(vbase, px.y), v1);
                SL::load(elem_ptr(vbase, px.z), v2);
                SL::load(elem_ptr(vbase, px.w), v3);
                if constexpr (scatter == 1) {
                    red_row<kCh>(elem_ptr(gvbase, px.x), w0, g);
                    red_row<kCh>(elem_ptr(gvbase, px.y), w1, g);
                    red_row<kCh>(elem_ptr(gvbase, px.z), w2, g);
                    red_row<kCh>(elem_ptr(gvbase, px.w), w3, g);
                } else if constexpr (scatter == 2) {
                    if constexpr (sizeof(VT) == 2 && kCh == 4) {
                        red_add_bf16x4(elem_ptr(gdbase, px.x), w0 * g[0], w0 * g[1], w0 * g[2], w0 * g[3]);
                        red_add_bf16x4(elem_ptr(gdbase, px.y), w1 * g[0], w1 * g[1], w1 * g[2], w1 * g[3]);
                        red_add_bf16x4(elem_ptr(gdbase, px.z), w2 * g[0], w2 * g[1], w2 * g[2], w2 * g[3]);
                        red_add_bf16x4(elem_ptr(gdbase, px.w), w3 * g[0], w3 * g[1], w3 * g[2], w3 * g[3]);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < kCh; ++c) v0[c] = v1[c] = v2[c] = v3[c] = 0.f;
                if (px.x >= 0) SL::load(elem_ptr(vbase, px.x), v0);
                if (px.y >= 0) SL::load(elem_ptr(vbase, px.y), v1);
                if (px.z >= 0) SL::load(elem_ptr(vbase, px.z), v2);
                if (px.w >= 0) SL::load(elem_ptr(vbase, px.w), v3);
                if constexpr (scatter == 1) {
                    if (px.x >= 0) red_row<kCh>(elem_ptr(gvbase, px.x), w0, g);
                    if (px.y >= 0) red_row<kCh>(elem_ptr(gvbase, px.y), w1, g);
                    if (px.z >= 0) red_row<kCh>(elem_ptr(gvbase, px.z), w2, g);
                    if (px.w >= 0) red_row<kCh>(elem_ptr(gvbase, px.w), w3, g);
                } else if constexpr (scatter == 2) {
                    if constexpr (sizeof(VT) == 2 && kCh == 4) {
                        if (px.x >= 0) red_add_bf16x4(elem_ptr(gdbase, px.x), w0 * g[0], w0 * g[1], w0 * g[2], w0 * g[3]);
                        if (px.y >= 0) red_add_bf16x4(elem_ptr(gdbase, px.y), w1 * g[0], w1 * g[1], w1 * g[2], w1 * g[3]);
                        if (px.z >= 0) red_add_bf16x4(elem_ptr(gdbase, px.z), w2 * g[0], w2 * g[1], w2 * g[2], w2 * g[3]);
                        if (px.w >= 0) red_add_bf16x4(elem_ptr(gdbase, px.w), w3 * g[0], w3 * g[1], w3 * g[2], w3 * g[3]);
                    }
                }
            }
            float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
            for (int c = 0; c < kCh; ++c) {
                d0 = fmaf(g[c], v0[c], d0);
                d1 = fmaf(g[c], v1[c], d1);
                d2 = fmaf(g[c], v2[c], d2);
                d3 = fmaf(g[c], v3[c], d3);
            }
            // transposing butterfly over the pair's lanes: 4 values x 8 lanes -> every value summed, one per lane
            static_assert(kLanes == 8 || kLanes == 16, "the butterfly below is written for 8 or 16 lanes per pair");
            float x = odd1 ? d2 : d0, y = odd1 ? d3 : d1;
            x += __shfl_xor_sync(0xffffffffu, odd1 ? d0 : d2, 1);
            y += __shfl_xor_sync(0xffffffffu, odd1 ? d1 : d3, 1);
            float z = odd2 ? y : x;
            z += __shfl_xor_sync(0xffffffffu, odd2 ? x : y, 2);
            z += __shfl_xor_sync(0xffffffffu, z, 4);
            if constexpr (kLanes == 16) z += __shfl_xor_sync(0xffffffffu, z, 8);
            // the pixel slot is dead after this iteration: it receives (d0, d1, d2, d3)
            if (lane < 4) reinterpret_cast<float *>(const_cast<int4 *>(my_pix + lp))[dot_slot] = z;
        }
    };
    for (int l = 0, lp = 0; l < L; ++l, lp += P) {
        const int scatter = s_scatter[l];
        if (scatter == 1) level_loop(std::integral_constant<int, 1>{}, lp);
        else if (sizeof(VT) == 2 && scatter == 2) level_loop(std::integral_constant<int, 2>{}, lp);
        else level_loop(std::integral_constant<int, 0>{}, lp);
    }
    __syncthreads();

    // (gw, gh, ga, a) of every sample from its four reduced dot products -- one thread per sample
    for (SampleWalk sw(threadIdx.x, kBwdThreads, LP); sw.s < nsamples; sw.next(kBwdThreads)) {
        const int idx = sw.pair * stride + sw.lp;
        const float4 d = *reinterpret_cast<const float4 *>(s_pix + idx);
        const float4 mt = s_meta[idx];
        const float lw = mt.x, lh = mt.y, a = mt.z;
        const float hw = 1.f - lw, hh = 1.f - lh;
        const float ga = fmaf(lh * lw, d.w, fmaf(lh * hw, d.z, fmaf(hh * lw, d.y, (hh * hw) * d.x)));
        const float gw = a * fmaf(hh, d.y - d.x, lh * (d.w - d.z));
        const float gh = a * fmaf(hw, d.z - d.x, lw * (d.w - d.y));
        s_meta[idx] = make_float4(gw, gh, ga, a);
    }
    if constexpr (IO::kFused) __syncthreads();  // the softmax backward below reads other threads' slots

    // ---- phase 3: dense stores of the per-sample gradients -------------------------------------------
    if constexpr (IO::kFused) {
        // softmax backward needs sum_j a_j * dL/da_j of the pair; then chain through the location formula
        {
            const int spair = threadIdx.x / kLanes, slane = threadIdx.x - spair * kLanes;
            const float4 *row = s_meta + (spair < npairs ? spair : 0) * stride;
            float dot = 0.f;
            if (spair < npairs)
                for (int lp = slane; lp < LP; lp += kLanes) dot = fmaf(row[lp].w, row[lp].z, dot);
#pragma unroll
            for (int off = kLanes / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
            if (slane == 0 && spair < npairs) s_stat[spair].x = dot;
        }
        __syncthreads();
        for (SampleWalk sw(threadIdx.x, kBwdThreads, LP); sw.s < nsamples; sw.next(kBwdThreads)) {
            const int s = sw.s, pr = sw.pair, lp = sw.lp;
            const int l = s_lvl[lp];
            const float4 r = s_meta[pr * stride + lp];
            const long long gs = pair0 * LP + s;
            const float glx = (float)s_W[l] * r.x, gly = (float)s_H[l] * r.y;  // d/d(loc), as the plain path returns it
            float gox, goy;
            if (io.ref_dim == 2) {
                gox = glx / (float)s_W[l];
                goy = gly / (float)s_H[l];
            } else {
                const float *rp = io.ref + (s_bq[pr] * L + l) * 4;
                gox = ((glx * 0.5f) * rp[2]) / (float)P;
                goy = ((gly * 0.5f) * rp[3]) / (float)P;
            }
            from_f32(io.grad_offsets[2 * gs], gox);
            from_f32(io.grad_offsets[2 * gs + 1], goy);
            from_f32(io.grad_logits[gs], r.w * (r.z - s_stat[pr].x));
        }
    } else {
        float2 *gl2 = reinterpret_cast<float2 *>(io.grad_loc) + pair0 * LP;
        float *ga0 = io.grad_attn + pair0 * LP;
        for (SampleWalk sw(threadIdx.x, kBwdThreads, LP); sw.s < nsamples; sw.next(kBwdThreads)) {
            const int s = sw.s, pr = sw.pair, lp = sw.lp;
            const int l = s_lvl[lp];
            const float4 r = s_meta[pr * stride + lp];
            gl2[s] = make_float2((float)s_W[l] * r.x, (float)s_H[l] * r.y);
            ga0[s] = r.z;
        }
    }
}

#ifdef RDETR_TUNE_FWD
__global__ void spin_kernel(long long ns)
{
    long long t0;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (;;) {
        long long t;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
        if (t - t0 >= ns) break;
        __nanosleep(200);
    }
}
#endif

// fp32 accumulation buffer -> bf16 grad_value (8 elements per thread).  Rows of levels that were scattered straight into
// grad_value (mode 2 of msda_bwd_kernel: direct_max_updates > 0) are left alone; everything else is overwritten.
__global__ void __launch_bounds__(256) f32_to_bf16_kernel(const float *__restrict__ src, __nv_bfloat16 *__restrict__ dst,
                                                          long long n8, const int64_t *__restrict__ spatial_shapes,
                                                          const int64_t *__restrict__ level_start_index, int L, int S, int M, int Nq,
                                                          int P, int coarse_cap_rows, int direct_max_updates)
{
    __shared__ int s_lo[kMaxLevels], s_hi[kMaxLevels];  // row ranges of the directly scattered levels (empty: lo = hi = 0)
    if (threadIdx.x < kMaxLevels) {
        int lo = 0, hi = 0;
        if (direct_max_updates > 0 && (int)threadIdx.x < L) {
            const int H = (int)spatial_shapes[2 * threadIdx.x], W = (int)spatial_shapes[2 * threadIdx.x + 1];
            const int start = (int)level_start_index[threadIdx.x];
            if (!coarse_level(H, W, start, S, coarse_cap_rows) && direct_bf16_level(H, W, Nq, P, direct_max_updates, S)) {
                lo = start;
                hi = start + H * W;
            }
        }
        s_lo[threadIdx.x] = lo;
        s_hi[threadIdx.x] = hi;
    }
    __syncthreads();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    if (direct_max_updates > 0) {
        const int row = (int)((i * 8 / ((long long)M * 32)) % S);
#pragma unroll
        for (int l = 0; l < kMaxLevels; ++l)
            if (row >= s_lo[l] && row < s_hi[l]) return;
    }
    const float4 a = ld_stream_f4(reinterpret_cast<const float4 *>(src) + 2 * i);
    const float4 b = ld_stream_f4(reinterpret_cast<const float4 *>(src) + 2 * i + 1);
    uint4 t;
    t.x = pack_bf16(a.x, a.y); t.y = pack_bf16(a.z, a.w);
    t.z = pack_bf16(b.x, b.y); t.w = pack_bf16(b.z, b.w);
    reinterpret_cast<uint4 *>(dst)[i] = t;
}

template <typename VT, int CH, typename IO, int THREADS, int MINB = 0>
static int launch_bwd_variant(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out,
                              float *gv_f32, int B, int S, int M, int L, int Nq, int P, int coarse_cap, void *gv_direct,
                              int direct_max, cudaStream_t stream)
{
    constexpr int D = 32;
    constexpr int kLanes = D / CH;
    constexpr int kPairs = THREADS / kLanes;
    const long long total_pairs = (long long)B * Nq * M;
    const size_t smem = (size_t)kPairs * (L * P + 1) * 32 + (IO::kFused ? kPairs * sizeof(float2) : 0);
    auto kern = msda_bwd_kernel<VT, CH, D, IO, THREADS, MINB>;
    if (smem > 48 * 1024) {
        if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                "cudaFuncSetAttribute(msda_bwd)"))
            return rc;
    }
#ifdef RDETR_TUNE_FWD
    if (const char *e = getenv("RDETR_MSDA_BWD_SKIP_ROWS")) coarse_cap = atoi(e);  // timing experiments only (levels of at most that many pixels get NO grad_value)
#endif
    // next to the coarse kernel (which needs the largest shared-memory carve-out) the scatter kernel asks for the same
    // L1 / shared split: an SM does not hold CTAs of two kernels that want different splits.  Set only when it changes.
    {
#ifdef RDETR_TUNE_FWD
        static const int exp_carve = getenv("RDETR_COARSE_CARVEOUT") ? atoi(getenv("RDETR_COARSE_CARVEOUT")) : 1;
#else
        constexpr int exp_carve = 1;
#endif
        static std::atomic<int> current{(int)cudaSharedmemCarveoutDefault};  // per instantiation of this template
        const int want = (coarse_cap > 0 && exp_carve) ? (int)cudaSharedmemCarveoutMaxShared : (int)cudaSharedmemCarveoutDefault;
        if (current.load(std::memory_order_relaxed) != want) {
            if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, want),
                                    "cudaFuncSetAttribute(msda_bwd carveout)"))
                return rc;
            current.store(want, std::memory_order_relaxed);
        }
    }
    const long long grid = (total_pairs + kPairs - 1) / kPairs;
    if (grid > 0x7fffffffLL) return fail(RDETR_ERR_UNSUPPORTED, "msda_backward: B*Nq*M too large (%lld pairs)", total_pairs);
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(static_cast<const VT *>(value), shapes, lsi, io,
                                                    static_cast<const VT *>(grad_out), gv_f32, S, M, L, Nq, P, total_pairs, coarse_cap,
                                                    static_cast<VT *>(gv_direct), direct_max);
    return check_cuda(cudaGetLastError(), "msda_bwd_kernel launch");
}

template <typename VT, int CH, typename IO>
static int launch_bwd(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out,
                      float *gv_f32, int B, int S, int M, int L, int Nq, int P, int coarse_cap, void *gv_direct, int direct_max,
                      cudaStream_t stream)
{
    // encoder self-attention (every query is a pixel of the pyramid): tiled kernel with shared-memory grad_value
    // accumulators (msda_bwd_tile.cu); rdetr_msda_set_tile_mode(1) / RDETR_MSDA_TILE=1 keeps the flat kernel
    if (Nq == S && msda_tile_mode() == 2) {
        const int rc = launch_bwd_tile<VT, IO>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, stream);
        if (rc >= 0) return rc;
    }
#ifdef RDETR_TUNE_FWD
    const char *e = getenv("RDETR_MSDA_BWD_VARIANT");
    const int v = e ? atoi(e) : 0;
    if (v == 1) return launch_bwd_variant<VT, CH, IO, 256>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
    if (v == 2) return launch_bwd_variant<VT, CH, IO, 64>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
    if (v == 3) return launch_bwd_variant<VT, CH, IO, 128, 10>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
    if constexpr (sizeof(VT) == 4) {
        if (v == 5) return launch_bwd_variant<VT, 2, IO, 128>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
        if (v == 6) return launch_bwd_variant<VT, 2, IO, 256>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
        if (v == 7) return launch_bwd_variant<VT, 2, IO, 128, 12>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
    }
    if (v == 4) return launch_bwd_variant<VT, CH, IO, 128, 12>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
#endif
    // 128 threads: within noise of 256 / 64 (the kernel is L2-atomic bound), fewest barrier stalls; register
    // caps for more occupancy spill and are 30-70 % slower (variant 3)
    return launch_bwd_variant<VT, CH, IO, 128>(value, shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, coarse_cap, gv_direct, direct_max, stream);
}

// zero-fill of the fp32 accumulation target, the scatter kernel, and (bf16) the final conversion
template <typename IOF32, typename IOBF16>
static int run_backward(const void *value, const int64_t *shapes, const int64_t *lsi, const IOF32 &io32, const IOBF16 &io16,
                        const void *grad_out, void *grad_value, int B, int S, int M, int D, int L, int Nq, int P, int dtype,
                        void *workspace, cudaStream_t st)
{
    const size_t nvalue = (size_t)B * S * M * D;
    float *acc = dtype == RDETR_DTYPE_F32 ? static_cast<float *>(grad_value) : static_cast<float *>(workspace);
    if (int rc = check_cuda(cudaMemsetAsync(acc, 0, nvalue * sizeof(float), st), "cudaMemsetAsync(grad_value)")) return rc;
    const bool tiled = Nq == S && msda_tile_mode() == 2;
    // bf16: levels whose rows receive few updates are scattered straight into the (zeroed) bf16 grad_value
    const int direct_max = (dtype == RDETR_DTYPE_BF16 && !tiled && Nq > 0) ? msda_bf16_scatter_max_updates() : 0;
    void *gv_direct = direct_max > 0 ? grad_value : nullptr;
    if (direct_max > 0)
        if (int rc = check_cuda(cudaMemsetAsync(grad_value, 0, nvalue * sizeof(__nv_bfloat16), st), "cudaMemsetAsync(grad_value bf16)")) return rc;
    int cap = 0;
    if (Nq > 0) {
        // coarse levels: shared-memory accumulation on a side stream, concurrent with the scatter kernel (msda_bwd_coarse.cu)
        cap = tiled ? 0 : msda_coarse_cap_rows(Nq, P);
        cudaStream_t side = nullptr;
        cudaEvent_t done = nullptr;
        if (cap > 0) {
            const int rc = dtype == RDETR_DTYPE_F32
                               ? fork_coarse<float>(shapes, lsi, io32, grad_out, acc, B, S, M, L, Nq, P, st, &side, &done)
                               : fork_coarse<__nv_bfloat16>(shapes, lsi, io16, grad_out, acc, B, S, M, L, Nq, P, st, &side, &done);
            if (rc) return rc;
        }
        int rc = RDETR_OK;
        bool skip_scatter = false;
#ifdef RDETR_TUNE_FWD
        // tuning builds only (tools/exp_coarse2.py): hold the scatter kernel back so that the coarse CTAs are resident first,
        // or leave it out altogether (wrong gradients: timing of the coarse kernel alone)
        static const int exp_delay_us = getenv("RDETR_COARSE_DELAY_US") ? atoi(getenv("RDETR_COARSE_DELAY_US")) : 0;
        static const int exp_skip_scatter = getenv("RDETR_COARSE_ONLY") ? atoi(getenv("RDETR_COARSE_ONLY")) : 0;
        if (cap > 0 && exp_delay_us > 0) spin_kernel<<<1, 32, 0, st>>>(exp_delay_us * 1000LL);
        skip_scatter = cap > 0 && exp_skip_scatter;
#endif
        if (!skip_scatter)
            rc = dtype == RDETR_DTYPE_F32
                     ? launch_bwd<float, 4>(value, shapes, lsi, io32, grad_out, acc, B, S, M, L, Nq, P, cap, nullptr, 0, st)
                     : launch_bwd<__nv_bfloat16, 4>(value, shapes, lsi, io16, grad_out, acc, B, S, M, L, Nq, P, cap, gv_direct, direct_max, st);
        if (cap > 0) {
            const int rj = join_coarse(st, done);
            if (rc == RDETR_OK && rj) return rj;
        }
        if (rc) return rc;
    }
    if (dtype == RDETR_DTYPE_BF16) {
        const long long n8 = (long long)(nvalue / 8);  // D == 32 => divisible
        f32_to_bf16_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(acc, static_cast<__nv_bfloat16 *>(grad_value), n8, shapes, lsi, L, S,
                                                                         M, Nq, P, cap, direct_max);
        return check_cuda(cudaGetLastError(), "f32_to_bf16_kernel launch");
    }
    return RDETR_OK;
}

}  // namespace rdetr

extern "C" size_t rdetr_msda_backward_workspace_bytes(int B, int S, int M, int D, int L, int Nq, int P, int value_dtype)
{
    (void)L; (void)Nq; (void)P;
    if (value_dtype == RDETR_DTYPE_BF16) return (size_t)B * S * M * D * sizeof(float);
    return 0;
}

extern "C" int rdetr_msda_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                   const float *sampling_locations, const float *attention_weights, const void *grad_out,
                                   void *grad_value, float *grad_loc, float *grad_attn, int B, int S, int M, int D, int L,
                                   int Nq, int P, int value_dtype, void *workspace, size_t workspace_bytes,
                                   rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_msda("rdetr_msda_backward", B, S, M, D, L, Nq, P, value_dtype)) return rc;
    if (B == 0) return RDETR_OK;
    if (!value || !spatial_shapes || !level_start_index || !grad_value)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_backward: null pointer argument");
    if (Nq > 0 && (!sampling_locations || !attention_weights || !grad_out || !grad_loc || !grad_attn))
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_backward: null pointer argument");
    if (((uintptr_t)value | (uintptr_t)grad_value | (uintptr_t)grad_out | (uintptr_t)sampling_locations |
         (uintptr_t)attention_weights | (uintptr_t)grad_loc | (uintptr_t)grad_attn | (uintptr_t)workspace) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_backward: buffers must be 16-byte aligned");
    const size_t need = rdetr_msda_backward_workspace_bytes(B, S, M, D, L, Nq, P, value_dtype);
    if (need && (!workspace || workspace_bytes < need))
        return fail(RDETR_ERR_WORKSPACE, "rdetr_msda_backward: workspace of %zu bytes required, got %zu", need,
                    workspace ? workspace_bytes : (size_t)0);
    const DeviceGuard guard(value);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const PlainIO io{sampling_locations, attention_weights, grad_loc, grad_attn};
    return run_backward(value, spatial_shapes, level_start_index, io, io, grad_out, grad_value, B, S, M, D, L, Nq, P,
                        value_dtype, workspace, st);
}

extern "C" int rdetr_msda_fused_backward(const void *value, const int64_t *spatial_shapes, const int64_t *level_start_index,
                                         const float *reference_points, const void *sampling_offsets,
                                         const void *attention_logits, const uint8_t *key_padding_mask, const void *grad_out,
                                         void *grad_value, void *grad_offsets, void *grad_logits, int B, int S, int M, int D,
                                         int L, int Nq, int P, int ref_dim, int dtype, void *workspace, size_t workspace_bytes,
                                         rdetr_stream_t stream)
{
    using namespace rdetr;
    if (int rc = validate_msda("rdetr_msda_fused_backward", B, S, M, D, L, Nq, P, dtype)) return rc;
    if (ref_dim != 2 && ref_dim != 4)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_backward: last dim of reference_points must be 2 or 4, got %d", ref_dim);
    if (B == 0) return RDETR_OK;
    if (!value || !spatial_shapes || !level_start_index || !grad_value)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_backward: null pointer argument");
    if (Nq > 0 && (!reference_points || !sampling_offsets || !attention_logits || !grad_out || !grad_offsets || !grad_logits))
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_backward: null pointer argument");
    if (((uintptr_t)value | (uintptr_t)grad_value | (uintptr_t)grad_out | (uintptr_t)workspace) & 15)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_fused_backward: value/grad buffers must be 16-byte aligned");
    const size_t need = rdetr_msda_backward_workspace_bytes(B, S, M, D, L, Nq, P, dtype);
    if (need && (!workspace || workspace_bytes < need))
        return fail(RDETR_ERR_WORKSPACE, "rdetr_msda_fused_backward: workspace of %zu bytes required, got %zu", need,
                    workspace ? workspace_bytes : (size_t)0);
    const DeviceGuard guard(value);
    if (guard.status()) return guard.status();
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const FusedIO<float> io32{reference_points, static_cast<const float *>(sampling_offsets),
                              static_cast<const float *>(attention_logits), key_padding_mask,
                              static_cast<float *>(grad_offsets), static_cast<float *>(grad_logits), ref_dim};
    const FusedIO<__nv_bfloat16> io16{reference_points, static_cast<const __nv_bfloat16 *>(sampling_offsets),
                                      static_cast<const __nv_bfloat16 *>(attention_logits), key_padding_mask,
                                      static_cast<__nv_bfloat16 *>(grad_offsets), static_cast<__nv_bfloat16 *>(grad_logits), ref_dim};
    return run_backward(value, spatial_shapes, level_start_index, io32, io16, grad_out, grad_value, B, S, M, D, L, Nq, P, dtype,
                        workspace, st);
}

extern "C" int rdetr_msda_set_bf16_scatter(int max_updates_per_row)
{
    if (max_updates_per_row < 0)
        return rdetr::fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_set_bf16_scatter: max_updates_per_row must be >= 0, got %d", max_updates_per_row);
    rdetr::g_bf16_scatter.store(max_updates_per_row, std::memory_order_relaxed);
    return RDETR_OK;
}
