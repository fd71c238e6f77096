// topk.cu -- two-stage query selection (second half).
//
// Replaces, in RelationTransformer.forward (upstream models/bricks/relation_transformer.py:92-96 and :107-111),
//   topk_index = torch.topk(enc_outputs_class.max(-1)[0], topk, dim=1)[1].unsqueeze(-1)
//   enc_outputs_class = enc_outputs_class.gather(1, topk_index.expand(-1, -1, num_classes))
//   enc_outputs_coord = enc_outputs_coord.gather(1, topk_index.expand(-1, -1, 4))        (after .sigmoid(), :90)
// Index work: the selected indices are torch.topk's (descending scores; NaN ranks above everything, as torch does);
// equal scores are ordered by ascending index, a deterministic refinement of torch's unspecified tie order.
//
//   rowmax_kernel        warp per (image, token): max over the C class logits            (HBM: reads B*S*C floats once)
//   topk_select_kernel   one cluster (<= 8 CTAs) per image: 4-pass radix select of the K-th largest score over the L2-resident
//                        row maxima, collection of the winners, bitonic sort of (score, index) in shared memory
//   gather_rows_kernel   warp per selected row: class row copy, box row copy (+ sigmoid)
//   scatter_rows_kernel  backward of the two gathers (+ sigmoid') into zero-filled gradients
#include <cooperative_groups.h>

#include "common.cuh"

namespace rdetr {

namespace cg = cooperative_groups;

namespace {

constexpr int kSelThreads = 1024;
constexpr int kMaxK = 4096;

// order-preserving image of a float: a > b  <=>  key(a) > key(b), a == b <=> key(a) == key(b); NaN is the largest key
// (torch.topk's convention)
__device__ __forceinline__ uint32_t order_key(float x)
{
    if (x != x) return 0xffffffffu;
    const uint32_t b = x == 0.f ? 0u : __float_as_uint(x);   // -0 and +0 compare equal: one key
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void __launch_bounds__(256) rowmax_kernel(const float *__restrict__ logits, float *__restrict__ row_max, long long rows, int C)
{
    const int lane = threadIdx.x & 31;
    const long long warp = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long r = warp; r < rows; r += nwarps) {
        const float *p = logits + r * C;
        float m = -INFINITY;
        bool nan = false;
        for (int c = lane; c < C; c += 32) {
            const float v = ld_stream_f1(p + c);
            nan |= (v != v);
            m = fmaxf(m, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        nan = __any_sync(0xffffffffu, nan);
        if (lane == 0) row_max[r] = nan ? __int_as_float(0x7fc00000) : m;   // torch's max propagates NaN
    }
}

// scores [B, S] -> indices [B, K] (int64), optional values [B, K].
// One thread-block CLUSTER per image (1, 2, 4 or 8 CTAs; the host picks the size so that a CTA's share of the row is a few
// thousand scores).  Every CTA converts its share to order keys ONCE (kept in shared memory when STAGED), the four radix
// passes run over shared memory, the per-CTA digit histograms are summed by every CTA through distributed shared memory
// (so all of them take the same decision without a broadcast), the winners are written straight into the sort buffer of
// the cluster's CTA 0, which sorts and stores them.
// Dynamic shared memory: Kpad * 8 bytes (sort buffer, used by CTA 0) + share * 4 bytes (keys, when STAGED).
template <bool STAGED>
__global__ void __launch_bounds__(kSelThreads) topk_select_kernel(const float *__restrict__ scores, int S, int K, int Kpad, int share,
                                                                  int64_t *__restrict__ indices, float *__restrict__ values)
{
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned CL = cluster.num_blocks(), rank = cluster.block_rank();
    const int image = blockIdx.x / CL;

    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long *sbuf = reinterpret_cast<unsigned long long *>(smem_raw);   // (key << 32 | ~index) of the winners
    uint32_t *skeys = reinterpret_cast<uint32_t *>(sbuf + Kpad);
    __shared__ uint32_t hist[4][256];          // one per pass: a peer may still be summing pass p while this CTA runs p + 1
    __shared__ uint32_t ghist[256];
    __shared__ uint32_t s_prefix, s_remaining, s_cnt_gt, s_ties;
    __shared__ uint32_t warp_tot[kSelThreads / 32];

    const int tid = threadIdx.x, lane = tid & 31;
    const float *row = scores + (size_t)image * S;
    const int lo = min(S, (int)rank * share), n = min(S, lo + share) - lo;          // this CTA's slice [lo, lo + n)
    const int iters = (n + kSelThreads - 1) / kSelThreads;
    auto key_at = [&](int j) -> uint32_t { return STAGED ? skeys[j] : order_key(row[lo + j]); };

    for (int i = tid; i < 4 * 256; i += kSelThreads) (&hist[0][0])[i] = 0;
    if (tid == 0) { s_prefix = 0; s_remaining = (uint32_t)K; s_cnt_gt = 0; s_ties = 0; }
    if (rank == 0)
        for (int i = K + tid; i < Kpad; i += kSelThreads) sbuf[i] = 0ull;          // padding sorts to the end
    if (STAGED)
        for (int j = tid; j < n; j += kSelThreads) skeys[j] = order_key(row[lo + j]);
    __syncthreads();

    uint32_t mask = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        const uint32_t prefix = s_prefix;
        uint32_t *h = hist[pass];
        for (int it = 0; it < iters; ++it) {
            const int j = it * kSelThreads + tid;
            uint32_t tag = 0xffff0000u | lane;            // lanes without a candidate: a group of their own
            if (j < n) {
                const uint32_t key = key_at(j);
                if ((key & mask) == prefix) tag = (key >> shift) & 255u;
            }
            // scores of one image share their leading bytes: aggregate equal digits inside the warp before the atomic
            const uint32_t peers = __match_any_sync(0xffffffffu, tag);
            if (tag < 256u && lane == __ffs(peers) - 1) atomicAdd(&h[tag], (uint32_t)__popc(peers));
        }
        cluster.sync();
        if (tid < 256) {
            uint32_t c = 0;
            for (unsigned r = 0; r < CL; ++r) c += *cluster.map_shared_rank(&h[tid], r);
            ghist[tid] = c;
        }
        __syncthreads();
        // digit d with  count(> d) < remaining <= count(>= d): suffix sums over the 256 bins by warp 0
        if (tid < 32) {
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int u = 0; u < 8; ++u) { c[u] = ghist[tid * 8 + u]; tot += c[u]; }
            uint32_t run = tot;   // becomes: elements in the bins of this lane and of all higher lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_down_sync(0xffffffffu, run, o);
                if (lane + o < 32) run += t;
            }
            const uint32_t above = run - tot;
            const uint32_t rem = s_remaining;
            __syncwarp();   // every lane has read s_remaining before the owner of the digit rewrites it
            if (above < rem && rem <= above + tot) {
                uint32_t acc = above;
#pragma unroll
                for (int u = 7; u >= 0; --u) {
                    if (acc < rem && rem <= acc + c[u]) {
                        s_prefix = prefix | ((uint32_t)(tid * 8 + u) << shift);
                        s_remaining = rem - acc;
                    }
                    acc += c[u];
                }
            }
        }
        mask |= 255u << shift;
        __syncthreads();
    }
    const uint32_t T = s_prefix;               // key of the K-th largest score
    const uint32_t need_eq = s_remaining;      // how many of the scores equal to it belong to the result (>= 1)
    const uint32_t n_gt = (uint32_t)K - need_eq;
    unsigned long long *out0 = cluster.map_shared_rank(sbuf, 0);
    uint32_t *cnt0 = cluster.map_shared_rank(&s_cnt_gt, 0);

    // scores above the threshold: one slot counter for the whole cluster (in CTA 0), one atomic per warp and iteration
    for (int it = 0; it < iters; ++it) {
        const int j = it * kSelThreads + tid;
        const uint32_t key = j < n ? key_at(j) : 0u;
        const bool win = j < n && key > T;
        const uint32_t votes = __ballot_sync(0xffffffffu, win);
        if (votes) {
            uint32_t base = 0;
            if (lane == __ffs(votes) - 1) base = atomicAdd(cnt0, (uint32_t)__popc(votes));
            base = __shfl_sync(0xffffffffu, base, __ffs(votes) - 1);
            if (win) out0[base + __popc(votes & ((1u << lane) - 1u))] = ((unsigned long long)key << 32) | (0xffffffffu - (uint32_t)(lo + j));
        }
    }
    // scores equal to the threshold, by ascending index: thread t owns the contiguous range [t * chunk, (t + 1) * chunk) of the
    // slice, slices are contiguous ranges of the row -> rank of a tie = cluster prefix + block prefix + position in the thread
    {
        const int chunk = (n + kSelThreads - 1) / kSelThreads;
        const int a = min(n, tid * chunk), b = min(n, a + chunk);
        uint32_t mine = 0;
        for (int j = a; j < b; ++j) mine += (key_at(j) == T);
        uint32_t incl = mine;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_tot[tid >> 5] = incl;
        __syncthreads();
        if (tid < 32) {
            const uint32_t w = warp_tot[tid];
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_tot[tid] = wi - w;
            if (tid == 31) s_ties = wi;
        }
        cluster.sync();
        uint32_t before = 0;
        for (unsigned r = 0; r < rank; ++r) before += *cluster.map_shared_rank(&s_ties, r);
        uint32_t slot = before + warp_tot[tid >> 5] + incl - mine;
        for (int j = a; j < b && slot < need_eq; ++j)
            if (key_at(j) == T) {
                out0[n_gt + slot] = ((unsigned long long)T << 32) | (0xffffffffu - (uint32_t)(lo + j));
                ++slot;
            }
    }
    cluster.sync();          // all winners are in CTA 0's buffer; nobody reads a peer's shared memory after this point
    if (rank != 0) return;

    // bitonic sort, descending on (key, ~index): scores descending, equal scores by ascending index
    for (int k = 2; k <= Kpad; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Kpad; i += kSelThreads) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned long long a = sbuf[i], b = sbuf[p];
                    if ((a < b) == ((i & k) == 0)) { sbuf[i] = b; sbuf[p] = a; }
                }
            }
            __syncthreads();
        }
    for (int r = tid; r < K; r += kSelThreads) {
        const uint32_t i = 0xffffffffu - (uint32_t)(sbuf[r] & 0xffffffffull);
        indices[(size_t)image * K + r] = (int64_t)i;
        if (values) values[(size_t)image * K + r] = row[i];
    }
}

// torch's sigmoid and its derivative, operation for operation (aten UnarySignKernels.cu / BinaryMiscBackwardOpsKernels.cu)
__device__ __forceinline__ float sigmoid_like_torch(float x) { return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x))); }

__global__ void __launch_bounds__(256) gather_rows_kernel(const float *__restrict__ cls, const float *__restrict__ coord,
                                                          const int64_t *__restrict__ indices, float *__restrict__ out_cls,
                                                          float *__restrict__ out_coord, int S, int C, int K, long long rows, int sigmoid)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const long long b = r / K;
    const long long src = b * S + indices[r];
    for (int c = lane; c < C; c += 32) out_cls[r * C + c] = cls[src * C + c];
    if (coord != nullptr && lane < 4) {
        const float x = coord[src * 4 + lane];
        out_coord[r * 4 + lane] = sigmoid ? sigmoid_like_torch(x) : x;
    }
}

__global__ void __launch_bounds__(256) scatter_rows_kernel(const float *__restrict__ g_cls, const float *__restrict__ g_coord,
                                                           const float *__restrict__ out_coord, const int64_t *__restrict__ indices,
                                                           float *__restrict__ grad_cls, float *__restrict__ grad_coord, int S, int C,
                                                           int K, long long rows, int sigmoid)
{
    const int lane = threadIdx.x & 31;
    const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const long long b = r / K;
    const long long dst = b * S + indices[r];
    if (grad_cls != nullptr)
        for (int c = lane; c < C; c += 32) grad_cls[dst * C + c] = g_cls[r * C + c];
    if (grad_coord != nullptr && lane < 4) {
        float g = g_coord[r * 4 + lane];
        if (sigmoid) {
            const float y = out_coord[r * 4 + lane];
            g = __fmul_rn(__fmul_rn(g, __fsub_rn(1.0f, y)), y);   // (grad * (1 - y)) * y, aten's order (sigmoid_backward)
        }
        grad_coord[dst * 4 + lane] = g;
    }
}

int next_pow2(int v)
{
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

constexpr int kMaxCluster = 8;           // portable cluster size
constexpr int kShareTarget = 4096;       // scores per CTA the host aims for
constexpr int kMaxStagedShare = 40960;   // 160 KB of keys + 32 KB sort buffer fit one SM's shared memory

template <bool STAGED>
int launch_select_as(const float *scores, int B, int S, int K, int Kpad, int CL, int share, int64_t *indices, float *values, cudaStream_t st)
{
    const size_t smem = (size_t)Kpad * sizeof(unsigned long long) + (STAGED ? (size_t)share * sizeof(uint32_t) : 0);
    if (int rc = check_cuda(cudaFuncSetAttribute(topk_select_kernel<STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(topk_select)"))
        return rc;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CL));
    cfg.blockDim = dim3(kSelThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return check_cuda(cudaLaunchKernelEx(&cfg, topk_select_kernel<STAGED>, scores, S, K, Kpad, share, indices, values), "topk_select_kernel launch");
}

int launch_select(const float *scores, int B, int S, int K, int64_t *indices, float *values, cudaStream_t st)
{
    const int Kpad = next_pow2(K < 2 ? 2 : K);
    int CL = 1;
    while (CL < kMaxCluster && (S + CL - 1) / CL > kShareTarget) CL <<= 1;
    const int share = (S + CL - 1) / CL;
    if (share <= kMaxStagedShare) return launch_select_as<true>(scores, B, S, K, Kpad, CL, share, indices, values, st);
    return launch_select_as<false>(scores, B, S, K, Kpad, CL, share, indices, values, st);
}

int check_topk_shape(const char *fn, int B, int S, int K)
{
    if (B <= 0 || S <= 0 || K <= 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: B, S and K must be positive (got %d, %d, %d)", fn, B, S, K);
    if (K > S) return fail(RDETR_ERR_INVALID_ARGUMENT, "%s: selected index k out of range (k = %d > %d elements per row)", fn, K, S);
    if (K > kMaxK) return fail(RDETR_ERR_UNSUPPORTED, "%s: k = %d above the supported maximum %d", fn, K, kMaxK);
    return RDETR_OK;
}

}  // namespace

}  // namespace rdetr

using namespace rdetr;

extern "C" size_t rdetr_two_stage_workspace_bytes(int B, int S)
{
    if (B <= 0 || S <= 0) return 0;
    return (size_t)B * (size_t)S * sizeof(float);
}

extern "C" int rdetr_topk_rows(const float *scores, int B, int S, int K, int64_t *indices, float *values, rdetr_stream_t stream)
{
    if (!scores || !indices) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_topk_rows: null pointer");
    if (int rc = check_topk_shape("rdetr_topk_rows", B, S, K)) return rc;
    DeviceGuard guard(scores);
    if (guard.status()) return guard.status();
    return launch_select(scores, B, S, K, indices, values, (cudaStream_t)stream);
}

extern "C" int rdetr_two_stage_select(const float *class_logits, const float *coord, int B, int S, int C, int K, int apply_sigmoid,
                                      float *topk_class, float *topk_coord, int64_t *topk_index, void *workspace,
                                      size_t workspace_bytes, rdetr_stream_t stream)
{
    if (!class_logits || !topk_class || !topk_index || (coord != nullptr) != (topk_coord != nullptr))
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_two_stage_select: null pointer (coord and topk_coord go together)");
    if (int rc = check_topk_shape("rdetr_two_stage_select", B, S, K)) return rc;
    if (C <= 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_two_stage_select: C must be positive");
    if (!workspace || workspace_bytes < rdetr_two_stage_workspace_bytes(B, S))
        return fail(RDETR_ERR_WORKSPACE, "rdetr_two_stage_select: workspace of %zu bytes needed, %zu given",
                    rdetr_two_stage_workspace_bytes(B, S), workspace_bytes);
    DeviceGuard guard(class_logits);
    if (guard.status()) return guard.status();
    cudaStream_t st = (cudaStream_t)stream;
    float *row_max = (float *)workspace;
    const long long rows = (long long)B * S;
    const int grid = (int)((rows + 7) / 8 < 148 * 8 ? (rows + 7) / 8 : 148 * 8);
    rowmax_kernel<<<grid, 256, 0, st>>>(class_logits, row_max, rows, C);
    if (int rc = check_cuda(cudaGetLastError(), "rowmax_kernel launch")) return rc;
    if (int rc = launch_select(row_max, B, S, K, topk_index, nullptr, st)) return rc;
    const long long out_rows = (long long)B * K;
    gather_rows_kernel<<<(unsigned)((out_rows + 7) / 8), 256, 0, st>>>(class_logits, coord, topk_index, topk_class, topk_coord, S, C, K,
                                                                     out_rows, apply_sigmoid);
    return check_cuda(cudaGetLastError(), "gather_rows_kernel launch");
}

extern "C" int rdetr_two_stage_select_backward(const float *grad_topk_class, const float *grad_topk_coord, const float *topk_coord,
                                               const int64_t *topk_index, int B, int S, int C, int K, int apply_sigmoid,
                                               float *grad_class, float *grad_coord, rdetr_stream_t stream)
{
    if (!topk_index || (!grad_class && !grad_coord)) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_two_stage_select_backward: null pointer");
    if ((grad_class && !grad_topk_class) || (grad_coord && (!grad_topk_coord || (apply_sigmoid && !topk_coord))))
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_two_stage_select_backward: a requested gradient lacks its input");
    if (int rc = check_topk_shape("rdetr_two_stage_select_backward", B, S, K)) return rc;
    DeviceGuard guard(topk_index);
    if (guard.status()) return guard.status();
    cudaStream_t st = (cudaStream_t)stream;
    if (grad_class)
        if (int rc = check_cuda(cudaMemsetAsync(grad_class, 0, (size_t)B * S * C * sizeof(float), st), "zero grad_class")) return rc;
    if (grad_coord)
        if (int rc = check_cuda(cudaMemsetAsync(grad_coord, 0, (size_t)B * S * 4 * sizeof(float), st), "zero grad_coord")) return rc;
    const long long out_rows = (long long)B * K;
    scatter_rows_kernel<<<(unsigned)((out_rows + 7) / 8), 256, 0, st>>>(grad_topk_class, grad_topk_coord, topk_coord, topk_index, grad_class,
                                                                      grad_coord, S, C, K, out_rows, apply_sigmoid);
    return check_cuda(cudaGetLastError(), "scatter_rows_kernel launch");
}
