// common.cuh -- shared device helpers and host-side error plumbing for librdetr_ops.so (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "rdetr_ops.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "librdetr_ops is written for sm_100a (B200) only"
#endif

namespace rdetr {

// ---- host side ---------------------------------------------------------------------------------
int fail(int code, const char *fmt, ...);  // records the thread-local message, returns `code`
int check_cuda(cudaError_t e, const char *what);
// Makes the device that owns `ptr` current for this thread for the lifetime of the guard and restores
// the previous one afterwards (the reference has no device guard, ms_deform_attn_cuda.cu:57; we derive
// the device from the data so multi-GPU callers are safe whatever their current device is).
class DeviceGuard {
public:
    explicit DeviceGuard(const void *ptr);
    ~DeviceGuard();
    int status() const { return rc_; }  // RDETR_OK or an error code (message recorded)
private:
    int prev_ = -1;
    bool switched_ = false;
    int rc_ = 0;
};

constexpr int kMaxLevels = 8;
constexpr int kMaxPoints = 8;
// msda_bwd_coarse.cu: a level with at most kCoarseCapRows pixels has its grad_value plane accumulated in shared
// memory (128 B per row; 25x42 = 1 050 rows of the 800x1333 pyramid is the design point), one CTA per SM; planes of
// at most kCoarseSmallRows (13x21 = 273) run several CTAs per SM.
constexpr int kBf16ScatterDefault = 100;  // see rdetr_msda_set_bf16_scatter (include/rdetr_ops.h)
constexpr int kCoarseCapRows = 1056;
constexpr int kCoarseSmallRows = 280;

// ---- device side -------------------------------------------------------------------------------
#ifdef __CUDACC__

// streaming (read-once) loads: keep them out of L1 so gathered value lines stay resident
__device__ __forceinline__ float2 ld_stream_f2(const float2 *p)
{
    float2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_stream_f1(const float *p)
{
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ld_stream_f4(const float4 *p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// vector reduction to global memory without a return value: one 16-byte request per lane
__device__ __forceinline__ void red_add_f32x4(float *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

__device__ __forceinline__ void red_add_f32x2(float *p, float a, float b)
{
    asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
// w * g[0..CH) added to row `p`: one vector reduction per lane
template <int CH>
__device__ __forceinline__ void red_row(float *p, float w, const float (&g)[CH])
{
    if constexpr (CH == 2) {
        red_add_f32x2(p, w * g[0], w * g[1]);
    } else {
#pragma unroll
        for (int c0 = 0; c0 < CH; c0 += 4) red_add_f32x4(p + c0, w * g[c0], w * g[c0 + 1], w * g[c0 + 2], w * g[c0 + 3]);
    }
}

// base + off elements (off >= 0) as ONE instruction (IMAD.WIDE.U32); left to itself ptxas folds the lane's constant
// offset into a 32-bit add and rebuilds the 64-bit address with four LEA / IADD3 per corner, and a signed offset
// costs an IADD3 + LEA.HI.X.SX32 pair
template <typename T>
__device__ __forceinline__ T *elem_ptr(T *base, int off)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"((unsigned)off), "r"((unsigned)sizeof(T)), "l"((unsigned long long)base));
    return reinterpret_cast<T *>(r);
}

// bf16 pair packed in a 32-bit word -> two floats (exact: bf16 is the top half of an fp32)
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi)
{
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // .x = lo (low half), .y = hi
    return *reinterpret_cast<uint32_t *>(&v);
}

__device__ __forceinline__ uint2 ld_stream_u2(const uint2 *p)
{
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// Per-lane slice of one (pixel, head) row of `value`: CH channels held by one lane.
//   <float, 4>  16-byte loads, 8 lanes per D=32 row (one 128-byte line per row)
//   <bf16, 8>   16-byte loads, 4 lanes per row (64 bytes per row)
//   <bf16, 4>    8-byte loads, 8 lanes per row; used where each lane must still issue full
//                16-byte fp32 vector reductions (backward)
template <typename VT, int CH>
struct Slice;

template <>
struct Slice<float, 4> {
    static constexpr int kCh = 4;
    __device__ __forceinline__ static void load(const float *p, float (&v)[4])
    {
        const float4 t = __ldg(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ static void load_stream(const float *p, float (&v)[4])
    {
        const float4 t = ld_stream_f4(reinterpret_cast<const float4 *>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ static void load_shared(const float *p, float (&v)[4])  // plain load (shared-memory window)
    {
        const float4 t = *reinterpret_cast<const float4 *>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ static void store(float *p, const float (&v)[4])
    {
        *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

// <float, 2>: 8-byte loads, 16 lanes per row -- a warp-wide load touches TWO 128-byte lines instead of four (tuning
// variant of the forward: the L1 charges the lines of one instruction more than lines of separate instructions)
template <>
struct Slice<float, 2> {
    static constexpr int kCh = 2;
    __device__ __forceinline__ static void load(const float *p, float (&v)[2])
    {
        const float2 t = __ldg(reinterpret_cast<const float2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ static void load_stream(const float *p, float (&v)[2])
    {
        const float2 t = ld_stream_f2(reinterpret_cast<const float2 *>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ static void store(float *p, const float (&v)[2])
    {
        *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]);
    }
};

// <float, 8>: 32-byte loads (LDG.E.256, sm_100), 4 lanes per row -- a warp-wide load gathers EIGHT 128-byte rows.  The LSU
// accepts about one warp-level load per 7 cycles and SM whatever its width (profiles/r01_microbench.txt: 39 / 79 / 170 G
// rows/s at 1 / 2 / 4 rows per instruction), so rows per instruction is the lever the fp32 forward had left.
template <>
struct Slice<float, 8> {
    static constexpr int kCh = 8;
    __device__ __forceinline__ static void load(const float *p, float (&v)[8])
    {
        asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                     : "l"(p));
    }
    __device__ __forceinline__ static void store(float *p, const float (&v)[8])
    {
        reinterpret_cast<float4 *>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4 *>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <>
struct Slice<__nv_bfloat16, 8> {
    static constexpr int kCh = 8;
    __device__ __forceinline__ static void unpack(const uint4 t, float (&v)[8])
    {
        v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
        v[4] = bf16lo(t.z); v[5] = bf16hi(t.z); v[6] = bf16lo(t.w); v[7] = bf16hi(t.w);
    }
    __device__ __forceinline__ static void load(const __nv_bfloat16 *p, float (&v)[8])
    {
        unpack(__ldg(reinterpret_cast<const uint4 *>(p)), v);
    }
    __device__ __forceinline__ static void load_stream(const __nv_bfloat16 *p, float (&v)[8])
    {
        unpack(ld_stream_u4(reinterpret_cast<const uint4 *>(p)), v);
    }
    __device__ __forceinline__ static void load_shared(const __nv_bfloat16 *p, float (&v)[8])
    {
        unpack(*reinterpret_cast<const uint4 *>(p), v);
    }
    __device__ __forceinline__ static void store(__nv_bfloat16 *p, const float (&v)[8])
    {
        uint4 t;
        t.x = pack_bf16(v[0], v[1]); t.y = pack_bf16(v[2], v[3]);
        t.z = pack_bf16(v[4], v[5]); t.w = pack_bf16(v[6], v[7]);
        *reinterpret_cast<uint4 *>(p) = t;
    }
};

template <>
struct Slice<__nv_bfloat16, 4> {
    static constexpr int kCh = 4;
    __device__ __forceinline__ static void unpack(const uint2 t, float (&v)[4])
    {
        v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
    }
    __device__ __forceinline__ static void load(const __nv_bfloat16 *p, float (&v)[4])
    {
        unpack(__ldg(reinterpret_cast<const uint2 *>(p)), v);
    }
    __device__ __forceinline__ static void load_stream(const __nv_bfloat16 *p, float (&v)[4])
    {
        unpack(ld_stream_u2(reinterpret_cast<const uint2 *>(p)), v);
    }
    __device__ __forceinline__ static void store(__nv_bfloat16 *p, const float (&v)[4])
    {
        uint2 t;
        t.x = pack_bf16(v[0], v[1]); t.y = pack_bf16(v[2], v[3]);
        *reinterpret_cast<uint2 *>(p) = t;
    }
};

// One bilinear tap, prepared once per sample and shared by the lanes that own its channels.
//   pix[i] : index of the corner's pixel inside the batch element (level_start + h*W + w), -1 if the
//            corner is outside the level (zero padding) or the sample is outside the validity window
//   lw, lh : fractional parts; hw = 1-lw, hh = 1-lh
// Arithmetic follows ms_deform_im2col_cuda.cuh:261-285 and :22-73 of the reference:
//   w_im = loc_x*W - 0.5, h_im = loc_y*H - 0.5; contributes iff h_im>-1 && w_im>-1 && h_im<H && w_im<W
struct Tap {
    int pix[4];
    float lw, lh;
    int base;  // start + h0*W + w0 whether or not that corner exists (distance between the 2x2 footprints of two samples)
};

// `S` = rows of one image's value tensor: a corner whose pixel index reaches S (spatial_shapes / level_start_index
// inconsistent with the value tensor the caller passed) is treated as padding instead of being dereferenced -- the
// reference asserts the sizes on the host with a device sync per call (ms_deform_attn.py:313), which the drop-in
// module does not do.
__device__ __forceinline__ Tap make_tap(float lx, float ly, int H, int W, int start, int S = 0x7fffffff)
{
    Tap t;
    const float w_im = fmaf(lx, (float)W, -0.5f);
    const float h_im = fmaf(ly, (float)H, -0.5f);
    const bool inside = (h_im > -1.f) && (w_im > -1.f) && (h_im < (float)H) && (w_im < (float)W);
    const float h0f = floorf(h_im), w0f = floorf(w_im);
    const int h0 = (int)h0f, w0 = (int)w0f;
    const int h1 = h0 + 1, w1 = w0 + 1;
    t.lh = h_im - h0f;
    t.lw = w_im - w0f;
    const bool h0ok = inside && h0 >= 0, h1ok = inside && h1 <= H - 1;
    const bool w0ok = w0 >= 0, w1ok = w1 <= W - 1;
    const int base = start + h0 * W + w0;
    t.base = base;
    t.pix[0] = (h0ok && w0ok) ? base : -1;
    t.pix[1] = (h0ok && w1ok) ? base + 1 : -1;
    t.pix[2] = (h1ok && w0ok) ? base + W : -1;
    t.pix[3] = (h1ok && w1ok) ? base + W + 1 : -1;
    if (base + W + 1 >= S || start < 0) {  // only with inconsistent shape tensors
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (t.pix[i] >= S || start < 0) t.pix[i] = -1;
    }
    if (!inside) { t.lh = 0.f; t.lw = 0.f; }  // NaN / inf locations must not leak into weights
    return t;
}

// Levels whose whole (image, head) plane of grad_value fits a CTA's shared memory are accumulated there by
// msda_bwd_coarse_kernel instead of going to L2 one vector reduction per corner; the scatter kernel and the coarse
// kernel both decide with this predicate, from the shape tensors on the device.  `cap_rows` = 0 disables it.
__device__ __forceinline__ bool coarse_level(int H, int W, int start, int S, int cap_rows)
{
    const long long rows = (long long)H * W;
    return cap_rows > 0 && H > 0 && W > 0 && rows <= cap_rows && start >= 0 && start + rows <= S;
}

// bf16 backward: a level whose rows receive few updates (Nq*P*4 corners spread over H*W rows <= max_updates each, on
// average) is scattered straight into the bf16 grad_value with packed bf16x2 reductions -- 64 bytes per row instead of
// 128, 81 vs 48 G rows/s at the L2 (profiles/r02al_microbench_red.txt) -- instead of into the fp32 workspace.
// The average is representative when every pixel of the pyramid is a query (Nq == S: encoder self-attention); the queries
// of a decoder call cluster on objects, so a quarter of the limit is applied there.  Levels of fewer than
// kBf16ScatterMinRows pixels never qualify: on a small level the in-range samples pile up on a few rows (the average says
// little) and there is nothing to gain.
constexpr int kBf16ScatterMinRows = 1024;
__device__ __forceinline__ bool direct_bf16_level(int H, int W, int Nq, int P, int max_updates, int S)
{
    const long long lim = Nq == S ? max_updates : max_updates / 4;
    const long long rows = (long long)H * W;
    return lim > 0 && H > 0 && W > 0 && rows >= kBf16ScatterMinRows && (long long)Nq * P * 4 <= lim * rows;
}
// {a, b, c, d} added to four consecutive bf16 (8 bytes per lane, 8 lanes = one 64-byte row)
__device__ __forceinline__ void red_add_bf16x4(__nv_bfloat16 *p, float a, float b, float c, float d)
{
    asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1,%2};" ::"l"(p), "r"(pack_bf16(a, b)), "r"(pack_bf16(c, d)) : "memory");
}

// Walks the sample ids s = tid, tid + T, tid + 2T, ... of a CTA and keeps (pair, lp) = (s / LP, s % LP) up to
// date with an add and a conditional carry: the two runtime integer divisions happen once per thread instead
// of once per sample (they were ~40 instructions of every phase-1 / phase-3 iteration).
struct SampleWalk {
    int s, pair, lp, dpair, dlp, LP;
    __device__ __forceinline__ SampleWalk(int tid, int T, int LP_) : s(tid), pair(tid / LP_), lp(tid % LP_), dpair(T / LP_), dlp(T % LP_), LP(LP_) {}
    __device__ __forceinline__ void next(int T)
    {
        s += T;
        pair += dpair;
        lp += dlp;
        if (lp >= LP) { lp -= LP; ++pair; }
    }
};

// ---- where phase 1 of the MSDA kernels gets a sample's (x, y, attention weight) from -------------
// PlainIO: the reference's operator signature -- sampling_locations / attention_weights are tensors.
// FusedIO: the module prologue folded into the kernel: softmax over the L*P logits
// of a (b,q,m), loc = ref + off/(W,H) (2-d reference points) or ref_xy + off/P * ref_wh * 0.5 (4-d
// boxes), and the key-padding mask, as models/bricks/ms_deform_attn.py:318-349 of the reference
// computes them with separate elementwise kernels.  QT = dtype of offsets / logits and their grads.
struct PlainIO {
    static constexpr bool kFused = false;
    const float *loc;
    const float *attn;
    float *grad_loc;
    float *grad_attn;
};

template <typename QT>
struct FusedIO {
    static constexpr bool kFused = true;
    const float *ref;        // [B, Nq, L, ref_dim] fp32
    const QT *offsets;       // [B, Nq, M, L, P, 2]
    const QT *logits;        // [B, Nq, M, L*P]
    const uint8_t *mask;     // [B, S] or nullptr; non-zero = padded pixel (value treated as 0)
    QT *grad_offsets;
    QT *grad_logits;
    int ref_dim;             // 2 or 4
};

// streaming loads of the fused prologue's inputs (read once; keep them out of L1 like loc / attn)
__device__ __forceinline__ float ld_stream_scalar(const float *p) { return ld_stream_f1(p); }
__device__ __forceinline__ float ld_stream_scalar(const __nv_bfloat16 *p)
{
    unsigned short r;
    asm volatile("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return __uint_as_float((uint32_t)r << 16);
}
__device__ __forceinline__ float2 ld_stream_pair(const float *p) { return ld_stream_f2(reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ float2 ld_stream_pair(const __nv_bfloat16 *p)
{
    uint32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
    return make_float2(bf16lo(r), bf16hi(r));
}

__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void from_f32(float &d, float v) { d = v; }
__device__ __forceinline__ void from_f32(__nv_bfloat16 &d, float v) { d = __float2bfloat16_rn(v); }

// sampling location of one sample from reference point + raw offset.  torch divides (off / W, off / P); the
// kernel multiplies by the reciprocal (exact for the power-of-two P of every shipped config, <= 1 ulp of the
// offset term otherwise -- the same class of difference as the summation order of the softmax).
__device__ __forceinline__ float2 fused_location(const float *rp, int ref_dim, float offx, float offy, float invW, float invH,
                                                 float invP)
{
    if (ref_dim == 2) return make_float2(fmaf(offx, invW, rp[0]), fmaf(offy, invH, rp[1]));
    return make_float2(fmaf((offx * invP) * rp[2], 0.5f, rp[0]), fmaf((offy * invP) * rp[3], 0.5f, rp[1]));
}

// ---- pieces of phase 1 shared by the forward and the backward kernel ------------------------------

// FusedIO only.  Softmax statistics (max, sum of exp) of every pair of the CTA: the pair's kLanes lanes split
// its L*P logits -- streamed once and parked in `.x` of the sample's shared-memory slot for the main loop --
// and combine with xor-shuffles.  Also records b*Nq+q and b per pair so that the 64-bit divisions happen once
// per pair, not once per sample.  Every thread of the CTA must call it; ends with a barrier.
template <int kLanes, typename IO>
__device__ __forceinline__ void fused_softmax_stats(const IO &io, long long pair0, int npairs, int LP, int M, int Nq,
                                                    float4 *slots, int stride, float2 *s_stat, long long *s_bq, int *s_b)
{
    const int spair = threadIdx.x / kLanes, slane = threadIdx.x - spair * kLanes;
    const bool live = spair < npairs;
    const auto *zrow = io.logits + (pair0 + (live ? spair : 0)) * LP;
    float4 *zslot = slots + (live ? spair : 0) * stride;
    float mx = -INFINITY;
    if (live)
        for (int lp = slane; lp < LP; lp += kLanes) {
            const float z = ld_stream_scalar(zrow + lp);
            zslot[lp].x = z;
            mx = fmaxf(mx, z);
        }
#pragma unroll
    for (int off = kLanes / 2; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
    float sum = 0.f;
    if (live)
        for (int lp = slane; lp < LP; lp += kLanes) sum += __expf(zslot[lp].x - mx);
#pragma unroll
    for (int off = kLanes / 2; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
    if (slane == 0 && live) {
        s_stat[spair] = make_float2(mx, sum);
        const long long bq = (pair0 + spair) / M;
        s_bq[spair] = bq;
        s_b[spair] = (int)(bq / Nq);
    }
    __syncthreads();
}

// Attention weight and bilinear tap of sample (pair, lp) of the CTA, from either source of inputs.
//   PlainIO: streamed from sampling_locations / attention_weights.
//   FusedIO: softmax from the parked logit and the pair's statistics, location from reference point + offset,
//            corners on padded pixels invalidated (value.masked_fill of the reference, ms_deform_attn.py:318-319).
template <typename IO>
__device__ __forceinline__ Tap sample_tap(const IO &io, long long pair0, int s, int pair, int lp, int l, int LP, int L, int S,
                                          float logit_x, const float2 *s_stat, const long long *s_bq, const int *s_b,
                                          const int *s_H, const int *s_W, const int *s_start, const float *s_invW,
                                          const float *s_invH, float inv_P, float &a)
{
    float2 xy;
    if constexpr (IO::kFused) {
        const float2 st = s_stat[pair];
        a = __expf(logit_x - st.x) / st.y;  // exp(z - max) / sum with the MUFU exponential (2 ulp)
        const float2 off = ld_stream_pair(io.offsets + 2 * (pair0 * LP + s));
        xy = fused_location(io.ref + (s_bq[pair] * L + l) * io.ref_dim, io.ref_dim, off.x, off.y, s_invW[l], s_invH[l], inv_P);
    } else {
        xy = ld_stream_f2(reinterpret_cast<const float2 *>(io.loc) + pair0 * LP + s);
        a = ld_stream_f1(io.attn + pair0 * LP + s);
    }
    Tap t = make_tap(xy.x, xy.y, s_H[l], s_W[l], s_start[l], S);
    if constexpr (IO::kFused) {
        if (io.mask != nullptr) {
            const uint8_t *mrow = io.mask + (long long)s_b[pair] * S;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t.pix[i] >= 0 && mrow[t.pix[i]]) t.pix[i] = -1;
        }
    }
    return t;
}

#endif  // __CUDACC__

}  // namespace rdetr
