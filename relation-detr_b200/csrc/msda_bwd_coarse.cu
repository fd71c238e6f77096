// msda_bwd_coarse.cu -- grad_value of the COARSE pyramid levels, accumulated in shared memory (sm_100a).
//
// Why: the scatter of msda_bwd.cu sends every bilinear corner to L2 as its own 128-byte vector reduction and runs at
// the L2 reduction rate (51 G rows/s).  On the 800x1333 pyramid half of those reductions land on the two coarsest
// levels -- 6 % of the rows, 1 308 / 340 updates per row (upstream ms_deform_im2col_cuda.cuh:290-392
// issues them as 32x more scalar atomics).  A level whose whole (image, head) plane of grad_value fits a CTA's
// shared memory (rows * 128 B: 35 KB for 13x21, 134 KB for 25x42) needs no atomics at all:
//
//   * a CTA owns one private fp32 copy of the plane for one (image, head, query chunk) and is the only writer;
//     the updates are plain ld.shared / fma / st.shared by ONE consumer warp (lane = corner x 4 channels, so one
//     128-bit access per lane covers the four 128-byte rows of a sample), flushed once with red.global.add.v4.f32
//     (rows * chunks requests per plane instead of 4 * Nq * P);
//   * the consumer is a dependent chain (load row, add, store row; the next sample may hit the same row), so it
//     works on FOUR query streams a quarter of the chunk apart: their footprints almost never overlap, the four
//     loads are issued together, and a per-group flag computed by the producers falls back to one-at-a-time
//     updates when two footprints could overlap -- correct for any sampling locations, no locality assumed;
//   * three producer warps decode samples (location -> corner rows and attention-scaled bilinear weights, from
//     either the plain tensors or the fused prologue's raw offsets / logits) one thread per sample and stage the
//     grad_output rows, each into its own shared-memory slot, handed over with named barriers; their global-load
//     latency hides behind the consumer.
//
// The kernel runs on a side stream next to the scatter kernel (which skips these levels): it needs issue slots and
// shared-memory bandwidth, the scatter kernel needs L2 reduction throughput.  Which levels it takes is decided on the
// device from spatial_shapes (coarse_level() in common.cuh) -- the shape tensors never visit the host.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace rdetr {

constexpr int kCoarseThreads = 128;   // warp 0 consumes, warps 1..3 produce
constexpr int kProducers = 3;
constexpr int kBatch = 32;            // samples per batch: 4 streams x 8 (query step, point) slots

// Named barriers with IMMEDIATE ids: with an id in a register ptxas reserves all 16 hardware barriers for the CTA, and a
// CTA that holds 16 barriers shares its SM with nothing else (measured: no overlap at all with the scatter kernel).
template <int ID>
__device__ __forceinline__ void bar_sync_c() { asm volatile("bar.sync %0, 64;" ::"n"(ID) : "memory"); }
template <int ID>
__device__ __forceinline__ void bar_arrive_c() { asm volatile("bar.arrive %0, 64;" ::"n"(ID) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int)
{
    switch (id) {
    case 1: bar_sync_c<1>(); break;
    case 2: bar_sync_c<2>(); break;
    case 3: bar_sync_c<3>(); break;
    case 4: bar_sync_c<4>(); break;
    case 5: bar_sync_c<5>(); break;
    default: bar_sync_c<6>(); break;
    }
}
__device__ __forceinline__ void named_bar_arrive(int id, int)
{
    switch (id) {
    case 1: bar_arrive_c<1>(); break;
    case 2: bar_arrive_c<2>(); break;
    case 3: bar_arrive_c<3>(); break;
    case 4: bar_arrive_c<4>(); break;
    case 5: bar_arrive_c<5>(); break;
    default: bar_arrive_c<6>(); break;
    }
}

// Attention weight and tap of sample (pair gp, level l, point p); rows are LOCAL to the level (start = 0).
template <typename IO>
__device__ __forceinline__ Tap coarse_decode(const IO &io, long long gp, long long bq, int b, int l, int p, int L, int P, int H,
                                             int W, int start, int S, float &a)
{
    const int LP = L * P;
    const long long gs = gp * LP + l * P + p;
    float2 xy;
    if constexpr (IO::kFused) {
        const auto *z = io.logits + gp * LP;
        float mx = -INFINITY;
        for (int j = 0; j < LP; ++j) mx = fmaxf(mx, ld_stream_scalar(z + j));
        float sum = 0.f;
        for (int j = 0; j < LP; ++j) sum += __expf(ld_stream_scalar(z + j) - mx);
        a = __expf(ld_stream_scalar(z + l * P + p) - mx) / sum;
        const float2 off = ld_stream_pair(io.offsets + 2 * gs);
        xy = fused_location(io.ref + (bq * L + l) * io.ref_dim, io.ref_dim, off.x, off.y, 1.0f / (float)W, 1.0f / (float)H,
                            1.0f / (float)P);
    } else {
        xy = ld_stream_f2(reinterpret_cast<const float2 *>(io.loc) + gs);
        a = ld_stream_f1(io.attn + gs);
    }
    Tap t = make_tap(xy.x, xy.y, H, W, 0);
    if constexpr (IO::kFused) {
        if (io.mask != nullptr) {
            const uint8_t *mrow = io.mask + (long long)b * S + start;
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (t.pix[i] >= 0 && mrow[t.pix[i]]) t.pix[i] = -1;
        }
    }
    return t;
}

// grid = (chunks, B*M, L); a CTA whose level is outside (rows_lo, rows_hi] exits at once.
// shared memory: acc[rows_hi][8] float4 | kProducers x { rec[32][4] float2, g[32/P][8] float4 } | conf[kProducers]
template <typename VT, typename IO>
__global__ void __launch_bounds__(kCoarseThreads, 1)
msda_bwd_coarse_kernel(const int64_t *__restrict__ spatial_shapes, const int64_t *__restrict__ level_start_index, const IO io,
                       const VT *__restrict__ grad_out, float *__restrict__ grad_value_f32, int S, int M, int L, int Nq, int P,
                       int rows_lo, int rows_hi)
{
    using SL = Slice<VT, 4>;
    const int l = blockIdx.z;
    const int H = (int)spatial_shapes[2 * l], W = (int)spatial_shapes[2 * l + 1], start = (int)level_start_index[l];
    if (!coarse_level(H, W, start, S, rows_hi) || H * W <= rows_lo) return;
    const int rows = H * W;
    const int b = blockIdx.y / M, m = blockIdx.y - b * M;
    const int c = blockIdx.x, C = gridDim.x;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *acc = reinterpret_cast<float4 *>(smem_raw);                                     // [rows_hi][8]
    const int gpairs = kBatch / P;                                                         // distinct (stream, query step) per batch
    const int slot_f4 = kBatch * 2 + gpairs * 8;                                           // float4 per producer slot
    float4 *slots = acc + (size_t)rows_hi * 8;
    unsigned *conf = reinterpret_cast<unsigned *>(slots + kProducers * slot_f4);           // [kProducers] x {conflict mask, live mask}

    for (int i = threadIdx.x; i < rows * 8; i += kCoarseThreads) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    const int n_c = c < Nq ? (Nq - c + C - 1) / C : 0;   // queries q = c + C*i of this chunk
    const int n4 = (n_c + 3) / 4;                        // four streams of n4 consecutive chunk entries
    const int tsteps = 8 / P;                            // query steps per batch (P divides 8)
    const int nbatch = (n4 + tsteps - 1) / tsteps;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp > 0) {
        // ---- producers: batch n goes to slot (n % kProducers), decoded by warp 1 + n % kProducers --------------
        const int w = warp - 1;
        float4 *slot = slots + w * slot_f4;
        float2 *rec = reinterpret_cast<float2 *>(slot);    // [32 samples][4 corners] (row as int bits, weight)
        float4 *gbuf = slot + kBatch * 2;                   // [gpairs][8]
        const int s = lane & 3, k = lane >> 2;              // stream, (step, point) slot of this lane's sample
        const int dt = k / P, p = k - dt * P;
        int round = 0;
        for (int n = w; n < nbatch; n += kProducers, ++round) {
            const int t = n * tsteps + dt;
            const int idx = s * n4 + t;
            const bool live = t < n4 && idx < n_c;
            int pix[4] = {-1, -1, -1, -1};
            float wgt[4] = {0.f, 0.f, 0.f, 0.f};
            int base = 0;
            if (live) {
                const int q = c + C * idx;
                const long long bq = (long long)b * Nq + q;
                float a;
                const Tap tp = coarse_decode(io, bq * M + m, bq, b, l, p, L, P, H, W, start, S, a);
                const float hh = 1.f - tp.lh, hw = 1.f - tp.lw;
                pix[0] = tp.pix[0]; pix[1] = tp.pix[1]; pix[2] = tp.pix[2]; pix[3] = tp.pix[3];
                wgt[0] = (hh * hw) * a; wgt[1] = (hh * tp.lw) * a; wgt[2] = (tp.lh * hw) * a; wgt[3] = (tp.lh * tp.lw) * a;
                base = tp.base;
            }
            const bool any = (pix[0] & pix[1] & pix[2] & pix[3]) >= 0;  // at least one corner exists
            // two samples of one group (lanes 4k..4k+3) whose 2x2 footprints could share a row
            bool clash = false;
#pragma unroll
            for (int x = 1; x < 4; ++x) {
                const int ob = __shfl_xor_sync(0xffffffffu, base, x);
                const bool oany = __shfl_xor_sync(0xffffffffu, (int)any, x) != 0;
                const int d = ob > base ? ob - base : base - ob;
                clash = clash || (any && oany && d <= W + 1);
            }
            const unsigned clash_mask = __ballot_sync(0xffffffffu, clash);
            const unsigned live_mask = __ballot_sync(0xffffffffu, any);
            // grad_output rows of the batch's pairs: 8 lanes x 4 channels per row, 4 rows per instruction
            float g[8][4];  // gpairs / 4 <= 8 instructions
            const int prow = lane >> 3, cg = lane & 7;
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                g[r][0] = g[r][1] = g[r][2] = g[r][3] = 0.f;
                const int pj = r * 4 + prow;  // pair slot: stream pj / tsteps, step pj % tsteps
                if (pj < gpairs) {
                    const int ps = pj / tsteps, pt = n * tsteps + (pj - ps * tsteps);
                    const int pidx = ps * n4 + pt;
                    if (pt < n4 && pidx < n_c) {
                        const long long gp = ((long long)b * Nq + c + (long long)C * pidx) * M + m;
                        SL::load_stream(grad_out + gp * 32 + cg * 4, g[r]);
                    }
                }
            }
            if (round > 0) named_bar_sync(1 + 2 * w + 1, 64);  // slot drained by the consumer
            reinterpret_cast<float4 *>(rec)[lane * 2] = make_float4(__int_as_float(pix[0]), wgt[0], __int_as_float(pix[1]), wgt[1]);
            reinterpret_cast<float4 *>(rec)[lane * 2 + 1] = make_float4(__int_as_float(pix[2]), wgt[2], __int_as_float(pix[3]), wgt[3]);
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const int pj = r * 4 + prow;
                if (pj < gpairs) gbuf[pj * 8 + cg] = make_float4(g[r][0], g[r][1], g[r][2], g[r][3]);
            }
            if (lane == 0) {
                conf[2 * w] = clash_mask;
                conf[2 * w + 1] = live_mask;
            }
            __syncwarp();
            named_bar_arrive(1 + 2 * w, 64);  // slot full
        }
    } else {
        // ---- consumer ---------------------------------------------------------------------------------------------
        const int i = lane >> 3, cg = lane & 7;  // corner, channel group
        for (int n = 0; n < nbatch; ++n) {
            const int w = n % kProducers;
            const float4 *slot = slots + w * slot_f4;
            const float2 *rec = reinterpret_cast<const float2 *>(slot);
            const float4 *gbuf = slot + kBatch * 2;
            named_bar_sync(1 + 2 * w, 64);
            const unsigned clash_mask = conf[2 * w], live_mask = conf[2 * w + 1];
#pragma unroll 2
            for (int k = 0; k < 8; ++k) {
                if (((live_mask >> (4 * k)) & 0xfu) == 0) continue;
                const int dt = k / P;
                float2 rw[4];
                float4 g4[4];
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    rw[s] = rec[(4 * k + s) * 4 + i];
                    g4[s] = gbuf[(s * tsteps + dt) * 8 + cg];
                }
                if (((clash_mask >> (4 * k)) & 0xfu) == 0) {
                    float4 a4[4];
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int row = __float_as_int(rw[s].x);
                        a4[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row >= 0) a4[s] = acc[row * 8 + cg];
                    }
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const float wv = rw[s].y;
                        a4[s].x = fmaf(wv, g4[s].x, a4[s].x);
                        a4[s].y = fmaf(wv, g4[s].y, a4[s].y);
                        a4[s].z = fmaf(wv, g4[s].z, a4[s].z);
                        a4[s].w = fmaf(wv, g4[s].w, a4[s].w);
                    }
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int row = __float_as_int(rw[s].x);
                        if (row >= 0) acc[row * 8 + cg] = a4[s];
                    }
                    __syncwarp();  // the next group may touch these rows from other lanes
                } else {
#pragma unroll
                    for (int s = 0; s < 4; ++s) {
                        const int row = __float_as_int(rw[s].x);
                        if (row >= 0) {
                            float4 a4 = acc[row * 8 + cg];
                            const float wv = rw[s].y;
                            a4.x = fmaf(wv, g4[s].x, a4.x);
                            a4.y = fmaf(wv, g4[s].y, a4.y);
                            a4.z = fmaf(wv, g4[s].z, a4.z);
                            a4.w = fmaf(wv, g4[s].w, a4.w);
                            acc[row * 8 + cg] = a4;
                        }
                        __syncwarp();  // the four corners of a sample are distinct rows; the next sample may reuse them
                    }
                }
            }
            __syncwarp();
            named_bar_arrive(1 + 2 * w + 1, 64);  // slot drained
        }
    }
    __syncthreads();

    // ---- flush: one vector reduction per touched row ----------------------------------------------------------------
    float *gv = grad_value_f32 + (((long long)b * S + start) * M + m) * 32;
    for (int r = threadIdx.x >> 3; r < rows; r += kCoarseThreads / 8) {
        const int cg = threadIdx.x & 7;
        const float4 v = acc[r * 8 + cg];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f)
            red_add_f32x4(gv + (long long)r * M * 32 + cg * 4, v.x, v.y, v.z, v.w);
    }
}

// ---- host side -----------------------------------------------------------------------------------------------------

static std::atomic<int> g_coarse_mode{-1};  // -1: not initialised (RDETR_MSDA_COARSE); 0 / 1 off, 2 on whenever P divides 8

int msda_coarse_mode()
{
    int mode = g_coarse_mode.load(std::memory_order_relaxed);
    if (mode < 0) {
        const char *e = getenv("RDETR_MSDA_COARSE");
        mode = e ? atoi(e) : 0;
        if (mode < 0 || mode > 2) mode = 0;
        g_coarse_mode.store(mode, std::memory_order_relaxed);
    }
    return mode;
}

// rows a level may have to be taken by the coarse kernel (0: none), for this call
int msda_coarse_cap_rows(int Nq, int P)
{
    // Off unless asked for (mode 2): measured at configs[1] the kernel alone takes 2.55 ms (latency-bound chains of
    // ld.shared / fma / st.shared at 4 warps per SM), the scatter kernel relieved of the two coarse levels 1.39 ms, both
    // together 3.5-3.8 ms against 1.77 ms for the scatter alone -- the read-modify-write in shared memory costs the
    // SM's load/store pipe two wavefronts per row where the fire-and-forget reduction costs one, and that pipe is what
    // the scatter kernel saturates (profiles/r02ac-r02ae_*).
    (void)Nq;
    const int mode = msda_coarse_mode();
    if (mode != 2 || (8 % P) != 0) return 0;
    return kCoarseCapRows;
}

// a lazily created high-priority side stream per device (never destroyed; the library has no unload hook)
static int side_stream(cudaStream_t *out)
{
    static std::mutex mu;
    static cudaStream_t streams[64] = {};
    int dev = 0;
    if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    if (dev < 0 || dev >= 64) return fail(RDETR_ERR_UNSUPPORTED, "msda_backward: device ordinal %d out of range", dev);
    std::lock_guard<std::mutex> lock(mu);
    if (!streams[dev]) {
        int lo = 0, hi = 0;
        if (int rc = check_cuda(cudaDeviceGetStreamPriorityRange(&lo, &hi), "cudaDeviceGetStreamPriorityRange")) return rc;
        if (int rc = check_cuda(cudaStreamCreateWithPriority(&streams[dev], cudaStreamNonBlocking, hi), "cudaStreamCreateWithPriority"))
            return rc;
    }
    *out = streams[dev];
    return RDETR_OK;
}

static int pick_chunks(int items, int slots)
{
    // time ~ ceil(items * C / slots) / C; smallest C within 3 % of the best of 1..24
    double best = 1e30;
    for (int C = 1; C <= 24; ++C) best = std::min(best, (double)((items * C + slots - 1) / slots) / C);
    for (int C = 1; C <= 24; ++C)
        if ((double)((items * C + slots - 1) / slots) / C <= best * 1.03) return C;
    return 1;
}

template <typename VT, typename IO>
static int launch_coarse_class(const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out, float *gv_f32, int B,
                               int S, int M, int L, int Nq, int P, int rows_lo, int rows_hi, int per_sm, int sms, cudaStream_t st)
{
    auto kern = msda_bwd_coarse_kernel<VT, IO>;
    const size_t smem = (size_t)rows_hi * 128 + (size_t)kProducers * ((kBatch * 2 + (kBatch / P) * 8) * 16) + kProducers * 2 * sizeof(unsigned);
    if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(msda_bwd_coarse)"))
        return rc;
    const int C = pick_chunks(B * M, sms * per_sm);
    const dim3 grid((unsigned)C, (unsigned)(B * M), (unsigned)L);
    kern<<<grid, kCoarseThreads, smem, st>>>(shapes, lsi, io, static_cast<const VT *>(grad_out), gv_f32, S, M, L, Nq, P, rows_lo, rows_hi);
    return check_cuda(cudaGetLastError(), "msda_bwd_coarse_kernel launch");
}

// Enqueues the coarse-level accumulation on the side stream: it starts after everything already in `main` (the
// zero-fill of grad_value) and `main` waits for it at the end (join_coarse).  Works under stream capture too (the side
// stream joins the capture through the event).
template <typename VT, typename IO>
int fork_coarse(const int64_t *shapes, const int64_t *lsi, const IO &io, const void *grad_out, float *gv_f32, int B, int S, int M,
                int L, int Nq, int P, cudaStream_t main, cudaStream_t *side_out, cudaEvent_t *done_out)
{
    cudaStream_t side;
    if (int rc = side_stream(&side)) return rc;
    int dev = 0, sms = 0;
    if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    if (int rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
    cudaEvent_t start, done;
    if (int rc = check_cuda(cudaEventCreateWithFlags(&start, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    if (int rc = check_cuda(cudaEventCreateWithFlags(&done, cudaEventDisableTiming), "cudaEventCreate")) return rc;
    int rc = check_cuda(cudaEventRecord(start, main), "cudaEventRecord");
    if (!rc) rc = check_cuda(cudaStreamWaitEvent(side, start, 0), "cudaStreamWaitEvent");
    // big planes first (one CTA per SM, the long pole), then the small ones (up to three per SM next to them)
    int exp_skip = 0;
#ifdef RDETR_TUNE_FWD
    // tuning builds only (profiles/r02ac-r02ae_exp_coarse2.txt): 1 skips the big class, 2 the small one, 3 both
    static const int exp_skip_env = getenv("RDETR_COARSE_SKIP") ? atoi(getenv("RDETR_COARSE_SKIP")) : 0;
    exp_skip = exp_skip_env;
#endif
    if (!rc && !(exp_skip & 1)) rc = launch_coarse_class<VT, IO>(shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, kCoarseSmallRows, kCoarseCapRows, 1, sms, side);
    if (!rc && !(exp_skip & 2)) rc = launch_coarse_class<VT, IO>(shapes, lsi, io, grad_out, gv_f32, B, S, M, L, Nq, P, 0, kCoarseSmallRows, 2, sms, side);
    if (!rc) rc = check_cuda(cudaEventRecord(done, side), "cudaEventRecord");
    cudaEventDestroy(start);  // released when it completes
    if (rc) {
        cudaEventDestroy(done);
        return rc;
    }
    *side_out = side;
    *done_out = done;
    return RDETR_OK;
}

int join_coarse(cudaStream_t main, cudaEvent_t done)
{
    const int rc = check_cuda(cudaStreamWaitEvent(main, done, 0), "cudaStreamWaitEvent");
    cudaEventDestroy(done);
    return rc;
}

template int fork_coarse<float, PlainIO>(const int64_t *, const int64_t *, const PlainIO &, const void *, float *, int, int, int, int,
                                         int, int, cudaStream_t, cudaStream_t *, cudaEvent_t *);
template int fork_coarse<__nv_bfloat16, PlainIO>(const int64_t *, const int64_t *, const PlainIO &, const void *, float *, int, int,
                                                 int, int, int, int, cudaStream_t, cudaStream_t *, cudaEvent_t *);
template int fork_coarse<float, FusedIO<float>>(const int64_t *, const int64_t *, const FusedIO<float> &, const void *, float *, int,
                                                int, int, int, int, int, cudaStream_t, cudaStream_t *, cudaEvent_t *);
template int fork_coarse<__nv_bfloat16, FusedIO<__nv_bfloat16>>(const int64_t *, const int64_t *, const FusedIO<__nv_bfloat16> &,
                                                                const void *, float *, int, int, int, int, int, int, cudaStream_t,
                                                                cudaStream_t *, cudaEvent_t *);

}  // namespace rdetr

extern "C" int rdetr_msda_set_coarse_mode(int mode)
{
    if (mode < 0 || mode > 2) return rdetr::fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_msda_set_coarse_mode: mode must be 0, 1 or 2, got %d", mode);
    rdetr::g_coarse_mode.store(mode, std::memory_order_relaxed);
    return RDETR_OK;
}
