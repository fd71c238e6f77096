// lsap.cu -- batched rectangular linear-sum-assignment on the device.
//
// Replaces `linear_sum_assignment(c.cpu())` in models/matcher/hungarian_matcher.py:80,87 of the
// reference: one device->host copy plus a SciPy call per image and per decoder layer, i.e. ~14*B host
// synchronisations per training step.  Here all problems of a step are solved by ONE launch, one CTA per
// problem, and the index tensors stay on the device.
//
// Algorithm: the shortest-augmenting-path solver SciPy implements (Crouse 2016), in double precision,
// with SciPy's conventions restated so that the SAME optimum is returned when several exist (transpose
// when rows > columns, reverse-filled unscanned-column vector with swap removal, "prefer an unassigned
// column, the last one in vector order, else the first minimal one").  The sequential column scan of one
// search step becomes a CTA-wide reduction whose comparator reproduces that scan's outcome exactly:
//   key  = order-preserving integer image of the reduced cost (minimum wins)
//   code = unassigned ? 2^31 | position : 2^31-1 - position    (maximum wins among equal keys)
// Every reduced cost is produced by the same three double-precision additions in the same order as the
// CPU code, so results are bit-identical, ties included (tests/test_lsap_gpu.py).
//
// State per problem lives in shared memory (29 bytes per column + 13 per row); the cost matrix is read
// from global memory through L1 (row-major in the orientation rows <= columns; a transposed fp32 copy is
// written to the caller's workspace when the input has more rows than columns, which is the matcher's
// case: queries x ground-truth boxes).
#include <cmath>

#include "common.cuh"

namespace rdetr {

// 256 threads, 8 columns per thread and batch.  A 1024-thread x 2-column shape was measured and dropped: 0.636 vs
// 0.630 ms on one step's problems, 5.46 vs 5.43 ms with crowded images (profiles/r01_tune_lsap.txt) -- a search step costs
// one cost-row round trip plus the reduction whatever the thread count.
constexpr int kLsapThreads = 256;
constexpr int kLsapWarps = kLsapThreads / 32;
constexpr int kLsapBatch = 8;
constexpr int kLsapTasksPerLaunch = 64;          // 48-byte descriptors: 3 KB of the 4 KB kernel-parameter space
constexpr size_t kLsapMaxSmem = 200 * 1024;

struct LsapTask {
    const float *cost;      // [n_rows, n_cols] row-major
    float *scratch;         // [n_cols, n_rows] when n_rows > n_cols, else unused
    long long *row_ind;     // [min(n_rows, n_cols)]
    long long *col_ind;
    int *status;            // one int: 0 ok, 1 infeasible, 2 NaN / -inf entry
    int n_rows, n_cols;
};
struct LsapLaunch { LsapTask t[kLsapTasksPerLaunch]; };

__host__ __device__ inline size_t lsap_smem_bytes(int nr, int nc)
{
    return (size_t)nc * (8 + 8 + 4 + 4 + 4 + 1) + (size_t)nr * (8 + 4 + 1) + 16;
}

// doubles -> unsigned integers with the same ordering (-0.0 folded into +0.0 first, as == does)
__device__ __forceinline__ unsigned long long order_key(double x)
{
    const long long b = __double_as_longlong(x + 0.0);
    return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ULL));
}
__device__ __forceinline__ double key_value(unsigned long long k)
{
    const long long b = (k >> 63) ? (long long)(k ^ 0x8000000000000000ULL) : (long long)~k;
    return __longlong_as_double(b);
}

// (key, code) of the winner over the warp: smallest key, then largest code
__device__ __forceinline__ void warp_pick(unsigned long long &key, unsigned &code)
{
    const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    const bool mine = hi == mhi && lo == mlo;
    code = __reduce_max_sync(0xffffffffu, mine ? code : 0u);
    key = ((unsigned long long)mhi << 32) | mlo;
}

__global__ void __launch_bounds__(kLsapThreads)
lsap_kernel(const __grid_constant__ LsapLaunch launch)
{
    const LsapTask &task = launch.t[blockIdx.x];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool transposed = task.n_cols < task.n_rows;
    const int nr = transposed ? task.n_cols : task.n_rows;
    const int nc = transposed ? task.n_rows : task.n_cols;
    if (nr <= 0) {
        if (tid == 0) *task.status = 0;
        return;
    }

    extern __shared__ __align__(16) unsigned char smem[];
    double *dist = reinterpret_cast<double *>(smem);     // reduced cost of the best path into each column
    double *v = dist + nc;                               // column duals
    double *u = v + nc;                                  // row duals
    int *path = reinterpret_cast<int *>(u + nr);         // predecessor row of each column
    int *row4col = path + nc;
    int *todo = row4col + nc;                            // unscanned columns, SciPy's vector order
    int *col4row = todo + nc;
    unsigned char *SC = reinterpret_cast<unsigned char *>(col4row + nr);
    unsigned char *SR = SC + nc;

    __shared__ unsigned long long s_key[2][kLsapWarps];
    __shared__ unsigned s_code[2][kLsapWarps];
    __shared__ int s_col[2][kLsapWarps];
    __shared__ int s_bad;
    __shared__ int s_scan[kLsapWarps];

    if (tid == 0) s_bad = 0;
    __syncthreads();

    // ---- pass 0: validate, and bring the matrix into the rows <= columns orientation -------------------
    const unsigned total = (unsigned)nr * (unsigned)nc;
    bool bad = false;
    if (transposed) {
#pragma unroll 8
        for (unsigned e = tid; e < total; e += kLsapThreads) {
            const float x = __ldg(task.cost + e);
            bad |= (x != x) || (x == -INFINITY);
            const unsigned q = e / (unsigned)nr, g = e - q * (unsigned)nr;   // input row q (column here), input column g
            task.scratch[(size_t)g * nc + q] = x;
        }
    } else {
#pragma unroll 8
        for (unsigned e = tid; e < total; e += kLsapThreads) {
            const float x = __ldg(task.cost + e);
            bad |= (x != x) || (x == -INFINITY);
        }
    }
    if (bad) atomicOr(&s_bad, 1);
    for (int j = tid; j < nc; j += kLsapThreads) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }
    for (int i = tid; i < nr; i += kLsapThreads) { u[i] = 0.0; col4row[i] = -1; }
    __syncthreads();
    const float *c = transposed ? task.scratch : task.cost;

    int status = s_bad ? 2 : 0;

    // ---- one augmenting path per row --------------------------------------------------------------------
    for (int cur = 0; cur < nr && status == 0; ++cur) {
        for (int j = tid; j < nc; j += kLsapThreads) { dist[j] = INFINITY; SC[j] = 0; todo[j] = nc - 1 - j; }
        for (int i = tid; i < nr; i += kLsapThreads) SR[i] = 0;
        __syncthreads();

        int i = cur, n_todo = nc, sink = -1;
        double min_val = 0.0;
        for (int it = 0; sink < 0; ++it) {
            if (tid == 0) SR[i] = 1;
            const double u_i = u[i];
            const float *crow = c + (size_t)i * nc;
            double best = INFINITY;                  // with code 0: loses against every real column (+inf included)
            unsigned code = 0;
            int col = -1;
            // kLsapBatch columns per thread at a time: all loads of a batch are issued before its first store, so
            // the cost-row reads (L1/L2 latency) of a thread overlap instead of queueing behind `dist[j] = r`
            // (the compiler cannot prove that two slots name different columns; the algorithm guarantees it)
            for (int base = tid; base < n_todo; base += kLsapThreads * kLsapBatch) {
                int jj[kLsapBatch], owner[kLsapBatch];
                float cc[kLsapBatch];
                double vv[kLsapBatch], dd[kLsapBatch];
#pragma unroll
                for (int k = 0; k < kLsapBatch; ++k) {
                    const int t = base + k * kLsapThreads;
                    jj[k] = t < n_todo ? todo[t] : -1;
                }
#pragma unroll
                for (int k = 0; k < kLsapBatch; ++k) {
                    const int j = jj[k] < 0 ? 0 : jj[k];
                    cc[k] = crow[j]; vv[k] = v[j]; dd[k] = dist[j]; owner[k] = row4col[j];
                }
#pragma unroll
                for (int k = 0; k < kLsapBatch; ++k) {
                    if (jj[k] < 0) continue;
                    const int t = base + k * kLsapThreads, j = jj[k];
                    const double r = ((min_val + (double)cc[k]) - u_i) - vv[k];
                    double d = dd[k];
                    if (r < d) { path[j] = i; dist[j] = r; d = r; }
                    const unsigned cd = owner[k] < 0 ? (0x80000000u | (unsigned)t) : (0x7fffffffu - (unsigned)t);
                    if (d < best || (d == best && cd > code)) { best = d; code = cd; col = j; }
                }
            }
            unsigned long long key = col >= 0 ? order_key(best) : ~0ULL;   // integer image once per thread
            const unsigned my_code = code;
            const unsigned long long my_key = key;
            warp_pick(key, code);
            const int buf = it & 1;
            if (my_key == key && my_code == code && col >= 0) s_col[buf][warp] = col;   // codes are unique: one writer
            if (lane == 0) { s_key[buf][warp] = key; s_code[buf][warp] = code; }
            __syncthreads();
            // every warp finishes the reduction on its own: no second barrier, partials are double-buffered
            unsigned long long gkey = lane < kLsapWarps ? s_key[buf][lane] : ~0ULL;
            unsigned gcode = lane < kLsapWarps ? s_code[buf][lane] : 0u;
            const unsigned long long wkey = gkey;
            const unsigned wcode = gcode;
            warp_pick(gkey, gcode);
            const unsigned winners = __ballot_sync(0xffffffffu, lane < kLsapWarps && wkey == gkey && wcode == gcode);
            const double lowest = key_value(gkey);
            if (!(lowest < INFINITY)) { status = 1; break; }      // uniform over the CTA: no finite candidate left
            const int j = s_col[buf][__ffs(winners) - 1];
            const int pick = (gcode & 0x80000000u) ? (int)(gcode & 0x7fffffffu) : (int)(0x7fffffffu - gcode);
            min_val = lowest;
            --n_todo;
            // the owner of slot `pick` is its only reader in the next scan; slot n_todo is not written this step
            if (pick % kLsapThreads == tid) todo[pick] = todo[n_todo];
            if (tid == 0) SC[j] = 1;
            const int owner = row4col[j];
            if (owner < 0) sink = j; else i = owner;
        }
        __syncthreads();
        if (status != 0) break;

        // dual update (before the path is flipped: col4row still describes the old matching)
        for (int r = tid; r < nr; r += kLsapThreads)
            if (SR[r]) u[r] += (r == cur) ? min_val : (min_val - dist[col4row[r]]);
        for (int j = tid; j < nc; j += kLsapThreads)
            if (SC[j]) v[j] -= min_val - dist[j];
        __syncthreads();
        if (tid == 0) {
            int j = sink;
            for (int hop = 0; hop <= nr; ++hop) {                 // an alternating path visits each row at most once
                const int r = path[j];
                row4col[j] = r;
                const int prev = col4row[r];
                col4row[r] = j;
                j = prev;
                if (r == cur) break;
            }
        }
        // the reset at the top of the next round touches other arrays; its barrier also publishes the flip
    }
    __syncthreads();

    // ---- output: pairs ordered by input row -------------------------------------------------------------
    if (tid == 0) *task.status = status;
    if (status != 0) {
        for (int k = tid; k < nr; k += kLsapThreads) { task.row_ind[k] = -1; task.col_ind[k] = -1; }
        return;
    }
    if (!transposed) {
        for (int k = tid; k < nr; k += kLsapThreads) { task.row_ind[k] = k; task.col_ind[k] = col4row[k]; }
        return;
    }
    // input rows are this kernel's columns: compact the assigned ones in ascending order
    const int per = (nc + kLsapThreads - 1) / kLsapThreads;
    const int lo = min(tid * per, nc), hi = min(lo + per, nc);
    int mine = 0;
    for (int q = lo; q < hi; ++q) mine += row4col[q] >= 0;
    int incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int up = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += up;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    int base = incl - mine;
    for (int w = 0; w < warp; ++w) base += s_scan[w];
    for (int q = lo; q < hi; ++q)
        if (row4col[q] >= 0) { task.row_ind[base] = q; task.col_ind[base] = row4col[q]; ++base; }
}

}  // namespace rdetr

extern "C" size_t rdetr_lsap_workspace_bytes(const int64_t *n_rows, const int64_t *n_cols, int n_problems)
{
    size_t total = 0;
    for (int p = 0; p < n_problems; ++p)
        if (n_rows[p] > n_cols[p] && n_cols[p] > 0) total += (((size_t)n_rows[p] * (size_t)n_cols[p] * sizeof(float)) + 255) & ~(size_t)255;
    return total;
}

extern "C" int rdetr_lsap_solve(const float *const *cost, const int64_t *n_rows, const int64_t *n_cols,
                                int64_t *const *row_ind, int64_t *const *col_ind, int32_t *status, int n_problems,
                                void *workspace, size_t workspace_bytes, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (n_problems < 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "lsap: n_problems=%d", n_problems);
    if (n_problems == 0) return RDETR_OK;
    if (!cost || !n_rows || !n_cols || !row_ind || !col_ind || !status)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "lsap: null argument");

    size_t smem_max = 0;
    for (int p = 0; p < n_problems; ++p) {
        if (n_rows[p] < 0 || n_cols[p] < 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "lsap: problem %d is %lld x %lld", p, (long long)n_rows[p], (long long)n_cols[p]);
        const int64_t nr = n_rows[p] < n_cols[p] ? n_rows[p] : n_cols[p], nc = n_rows[p] < n_cols[p] ? n_cols[p] : n_rows[p];
        if (nr == 0) continue;
        if (nr * nc >= (int64_t)1 << 31 || lsap_smem_bytes((int)nr, (int)nc) > kLsapMaxSmem)
            return fail(RDETR_ERR_UNSUPPORTED, "lsap: problem %d is %lld x %lld; the solver keeps its state in shared memory "
                        "(29 B per column + 13 B per row of the wide orientation, %zu B available)", p, (long long)n_rows[p],
                        (long long)n_cols[p], kLsapMaxSmem);
        if (!cost[p] || !row_ind[p] || !col_ind[p]) return fail(RDETR_ERR_INVALID_ARGUMENT, "lsap: null buffer in problem %d", p);
        const size_t s = lsap_smem_bytes((int)nr, (int)nc);
        if (s > smem_max) smem_max = s;
    }
    if (workspace_bytes < rdetr_lsap_workspace_bytes(n_rows, n_cols, n_problems))
        return fail(RDETR_ERR_WORKSPACE, "lsap: workspace of %zu bytes, %zu needed (rdetr_lsap_workspace_bytes)", workspace_bytes,
                    rdetr_lsap_workspace_bytes(n_rows, n_cols, n_problems));
    if (workspace_bytes > 0 && !workspace) return fail(RDETR_ERR_INVALID_ARGUMENT, "lsap: null workspace");
    const DeviceGuard guard(status);
    if (guard.status() != RDETR_OK) return guard.status();
    static thread_local int configured_device = -1;
    int dev = 0;
    int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice");
    if (rc != RDETR_OK) return rc;
    if (configured_device != dev) {
        rc = check_cuda(cudaFuncSetAttribute(lsap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kLsapMaxSmem),
                        "cudaFuncSetAttribute(lsap_kernel)");
        if (rc != RDETR_OK) return rc;
        configured_device = dev;
    }

    unsigned char *ws = static_cast<unsigned char *>(workspace);
    for (int first = 0; first < n_problems; first += kLsapTasksPerLaunch) {
        LsapLaunch launch;
        const int count = n_problems - first < kLsapTasksPerLaunch ? n_problems - first : kLsapTasksPerLaunch;
        for (int k = 0; k < count; ++k) {
            const int p = first + k;
            LsapTask &t = launch.t[k];
            t.cost = cost[p];
            t.row_ind = reinterpret_cast<long long *>(row_ind[p]);
            t.col_ind = reinterpret_cast<long long *>(col_ind[p]);
            t.status = status + p;
            t.n_rows = (int)n_rows[p];
            t.n_cols = (int)n_cols[p];
            t.scratch = nullptr;
            if (n_rows[p] > n_cols[p] && n_cols[p] > 0) {
                t.scratch = reinterpret_cast<float *>(ws);
                ws += (((size_t)n_rows[p] * (size_t)n_cols[p] * sizeof(float)) + 255) & ~(size_t)255;
            }
        }
        lsap_kernel<<<count, kLsapThreads, smem_max, static_cast<cudaStream_t>(stream)>>>(launch);
        rc = check_cuda(cudaGetLastError(), "lsap_kernel launch");
        if (rc != RDETR_OK) return rc;
    }
    return RDETR_OK;
}
