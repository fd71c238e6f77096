// diag.cu -- two micro-kernels that measure, on the box the benchmark runs on, the hardware rates that
// bound the MSDA kernels: (1) gathers of random 128-byte rows through L1/L2 with
// 8 lanes x LDG.128 per row -- the forward's access pattern; (2) red.global.add.v4.f32 on random
// 128-byte rows -- the backward's scatter.  bench.py reports the kernels' achieved rows/s against
// these ceilings next to the HBM roofline.  Diagnostic only: nothing on the product path calls them.
#include "common.cuh"

namespace rdetr {

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// Index arithmetic is kept to an LCG step + mask (rows rounded down to a power of two) so that the load
// path, not the ALU, is what saturates: with a hash + modulo per row the same kernel tops out 25 % lower.
__global__ void __launch_bounds__(256)
diag_gather_rows_kernel(const uint4 *__restrict__ table, uint32_t row_mask, int iters, float *__restrict__ sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid >> 3, lane = tid & 7;
    float acc = 0.f;
    uint32_t seed = group * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = seed * 1664525u + 1013904223u;
            v[u] = __ldg(table + (size_t)((seed >> 9) & row_mask) * 8 + lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += __uint_as_float(v[u].x);
    }
    if (acc == 123.456f) sink[0] = acc;  // never true for a zeroed table; keeps the loads alive
}

__global__ void __launch_bounds__(256)
diag_red_rows_kernel(float *__restrict__ table, uint32_t row_mask, int iters)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid >> 3, lane = tid & 7;
    uint32_t seed = group * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = seed * 1664525u + 1013904223u;
            red_add_f32x4(table + (size_t)((seed >> 9) & row_mask) * 32 + lane * 4, 1.f, 1.f, 1.f, 1.f);
        }
    }
}

}  // namespace rdetr

// Both: `table` = device buffer of nrows*128 bytes (16-byte aligned), 148*16 CTAs of 256 threads, every
// 8-lane group touches iters*8 random rows.  *rows_out (host) receives the number of rows touched.
extern "C" int rdetr_diag_gather_rows(const void *table, long long nrows, int iters, float *sink, long long *rows_out,
                                      rdetr_stream_t stream)
{
    using namespace rdetr;
    if (!table || !sink || nrows <= 0 || nrows > 0x7fffffffLL || iters <= 0 || ((uintptr_t)table & 15))
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_diag_gather_rows: bad argument");
    const DeviceGuard guard(table);
    if (guard.status()) return guard.status();
    const int blocks = 148 * 16, threads = 256;
    uint32_t pow2 = 1;
    while ((long long)pow2 * 2 <= nrows) pow2 *= 2;  // rows actually touched: the largest power of two <= nrows
    diag_gather_rows_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const uint4 *>(table),
                                                                                      pow2 - 1, iters, sink);
    if (rows_out) *rows_out = (long long)blocks * threads / 8 * iters * 8;
    return check_cuda(cudaGetLastError(), "diag_gather_rows_kernel launch");
}

extern "C" int rdetr_diag_red_rows(void *table, long long nrows, int iters, long long *rows_out, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (!table || nrows <= 0 || nrows > 0x7fffffffLL || iters <= 0 || ((uintptr_t)table & 15))
        return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_diag_red_rows: bad argument");
    const DeviceGuard guard(table);
    if (guard.status()) return guard.status();
    const int blocks = 148 * 16, threads = 256;
    uint32_t pow2 = 1;
    while ((long long)pow2 * 2 <= nrows) pow2 *= 2;
    diag_red_rows_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<float *>(table), pow2 - 1, iters);
    if (rows_out) *rows_out = (long long)blocks * threads / 8 * iters * 8;
    return check_cuda(cudaGetLastError(), "diag_red_rows_kernel launch");
}
