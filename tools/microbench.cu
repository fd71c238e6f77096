// microbench.cu -- B200 primitives that bound the two kernels (gather rows, vector REDs, smem atomics,
// FP32 FMA, MUFU).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench tools/microbench.cu
// Prints one line per experiment: name, ms, derived rate.  Design input only -- never a bench value.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each group of LANES lanes reads ROWS_PER_THREAD random rows of LANES*16 bytes; row r lives at r*row_stride bytes
template <int LANES>
__global__ void gather_rows(const uint4* __restrict__ table, uint32_t nrows, uint32_t row_stride16, int iters, float* sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid / LANES, lane = tid % LANES;
    float acc = 0.f;
    uint32_t seed = group * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const uint32_t row = seed % nrows;
            v[u] = __ldg(table + (size_t)row * row_stride16 + lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].w);
    }
    if (acc == 123.456f) sink[0] = acc;
}

// rows gathered with narrower per-lane loads: W words (4 bytes each) per lane, 32/W lanes per 128-byte row
template <int W>
__global__ void gather_rows_narrow(const float* __restrict__ table, uint32_t nrows, int iters, float* sink)
{
    constexpr int LANES = 32 / W;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid / LANES, lane = tid % LANES;
    float acc = 0.f;
    uint32_t seed = group * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        float v[8][W];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const float* p = table + (size_t)(seed % nrows) * 32 + lane * W;
            if (W == 1) v[u][0] = __ldg(p);
            else { const float2 t = __ldg(reinterpret_cast<const float2*>(p)); v[u][0] = t.x; v[u][W - 1] = t.y; }
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u][0] + v[u][W - 1];
    }
    if (acc == 123.456f) sink[0] = acc;
}

// cheap index arithmetic (LCG + mask, nrows a power of two) so that the load path, not the ALU, is the limit;
// INFLIGHT independent 16-byte loads per lane before any use
template <int INFLIGHT>
__global__ void gather_rows_lean(const uint4* __restrict__ table, uint32_t row_mask, int iters, float* sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid >> 3, lane = tid & 7;
    float acc = 0.f;
    uint32_t seed = group * 2654435761u + 12345u;
    for (int it = 0; it < iters; ++it) {
        uint4 v[INFLIGHT];
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) {
            seed = seed * 1664525u + 1013904223u;
            v[u] = __ldg(table + (size_t)((seed >> 9) & row_mask) * 8 + lane);
        }
#pragma unroll
        for (int u = 0; u < INFLIGHT; ++u) acc += __uint_as_float(v[u].x);
    }
    if (acc == 123.456f) sink[0] = acc;
}

// same access pattern but "spatially coherent": consecutive groups read rows near group index (L1 reuse)
template <int LANES>
__global__ void gather_rows_local(const uint4* __restrict__ table, uint32_t nrows, uint32_t row_stride16, int iters, float* sink)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid / LANES, lane = tid % LANES;
    float acc = 0.f;
    uint32_t seed = group * 2654435761u + 12345u;
    const uint32_t centre = (group / 8) % nrows;  // 8 heads share a pixel neighbourhood
    for (int it = 0; it < iters; ++it) {
        uint4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const uint32_t row = (centre + (seed & 15u) + ((seed >> 4) & 7u) * 168u) % nrows;
            v[u] = __ldg(table + (size_t)row * row_stride16 + lane);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].w);
    }
    if (acc == 123.456f) sink[0] = acc;
}

__device__ __forceinline__ void red4(float* p, float a) { asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(a) : "memory"); }
__device__ __forceinline__ void red2(float* p, float a) { asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(a) : "memory"); }

// MODE 0: v4 reds, 8 lanes per 128-byte row;  MODE 1: scalar reds, 32 lanes per row
template <int MODE>
__global__ void red_rows(float* table, uint32_t nrows, int iters)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    constexpr int LANES = MODE == 0 ? 8 : 32;
    const uint32_t group = tid / LANES, lane = tid % LANES;
    uint32_t seed = group * 2654435761u + 777u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const uint32_t row = seed % nrows;
            if (MODE == 0) red4(table + (size_t)row * 32 + lane * 4, 1.0f);
            else atomicAdd(table + (size_t)row * 32 + lane, 1.0f);
        }
    }
}

// shared-memory accumulation: each warp adds 128-byte rows into a CTA-private smem table of nrows rows
// MODE 0: atomicAdd (ATOMS / RED.shared), MODE 1: plain load-add-store (only valid with 1 warp per table; measures the bandwidth bound)
template <int MODE>
__global__ void smem_rows(int nrows, int iters, float* sink)
{
    extern __shared__ float tab[];
    for (int i = threadIdx.x; i < nrows * 32; i += blockDim.x) tab[i] = 0.f;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t seed = warp * 2654435761u + 99u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const uint32_t row = seed % nrows;
            if (MODE == 0) atomicAdd(&tab[row * 32 + lane], 1.0f);
            else tab[row * 32 + lane] += 1.0f;
        }
    }
    __syncthreads();
    if (tab[threadIdx.x] == -1.f) sink[0] = 1.f;
}

// TMA reduce-add: each 8-lane group stages one 128-byte row in shared memory and one lane issues
// cp.reduce.async.bulk (.add.f32) of 128 bytes to a random global row.  Moves the scatter off the LSU/L1 path.
template <int DEPTH>
__global__ void tma_red_rows(float* table, uint32_t nrows, int iters)
{
    extern __shared__ __align__(128) float stage[];  // [warps][DEPTH][4 groups][32 floats]
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    const uint32_t group = tid / 8, g = lane >> 3, l8 = lane & 7;
    float* mystage = stage + (size_t)warp_in_cta * DEPTH * 4 * 32;
    uint32_t seed = group * 2654435761u + 777u;
    int slot = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const uint32_t row = seed % nrows;
            if (l8 == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
            __syncwarp();
            float* dst = mystage + (slot * 4 + g) * 32;
            *reinterpret_cast<float4*>(dst + l8 * 4) = make_float4(1.f, 1.f, 1.f, 1.f);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (l8 == 0) {
                const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(dst);
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(table + (size_t)row * 32), "r"(saddr) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            slot = (slot + 1) % DEPTH;
        }
    }
    if (l8 == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void ffma_chain(int iters, float* sink, float a, float b)
{
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) sink[0] = s;
}

__global__ void ffma2_chain(int iters, float* sink, float a, float b)
{
    unsigned long long x[8], aa, bb;
    asm("mov.b64 %0, {%1,%1};" : "=l"(aa) : "f"(a));
    asm("mov.b64 %0, {%1,%1};" : "=l"(bb) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float t = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1,%1};" : "=l"(x[i]) : "f"(t)); }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(aa), "l"(bb));
    }
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    if (s == 123456ull) sink[0] = 1.f;
}

__global__ void mufu_chain(int iters, float* sink)
{
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 0.001f + i * 0.1f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) x[i] = __sinf(x[i]) + __cosf(x[i]);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    if (s == 123.456f) sink[0] = s;
}

__global__ void sincosf_chain(int iters, float* sink)
{
    float x[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) x[i] = threadIdx.x * 0.37f + i * 100.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { float s, c; sincosf(x[i], &s, &c); x[i] = x[i] * 0.999f + s + c + 300.f * (it & 1); }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += x[i];
    if (s == 123.456f) sink[0] = s;
}


// Mixed scatter: of every 4 consecutive 8-lane groups, TMAQ use the TMA reduce path (stage the row in shared memory,
// cp.reduce.async.bulk.add.f32 of 128 bytes) and the rest the LSU path (red.global.add.v4.f32).  If the two paths had
// independent ceilings the combined row rate would exceed either alone.
template <int TMAQ, int DEPTH>
__global__ void mixed_red_rows(float* table, uint32_t nrows, int iters)
{
    extern __shared__ __align__(128) float stage[];  // [warps][DEPTH][4 groups][32 floats]
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5;
    const uint32_t group = tid / 8, g = lane >> 3, l8 = lane & 7;
    const bool use_tma = (int)g < TMAQ;
    float* mystage = stage + (size_t)warp_in_cta * DEPTH * 4 * 32;
    uint32_t seed = group * 2654435761u + 777u;
    int slot = 0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll 1
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            const uint32_t row = seed % nrows;
            if (use_tma) {
                if (l8 == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
                __syncwarp(TMAQ == 4 ? 0xffffffffu : ((1u << (8 * TMAQ)) - 1u));
                float* dst = mystage + (slot * 4 + g) * 32;
                *reinterpret_cast<float4*>(dst + l8 * 4) = make_float4(1.f, 1.f, 1.f, 1.f);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp(TMAQ == 4 ? 0xffffffffu : ((1u << (8 * TMAQ)) - 1u));
                if (l8 == 0) {
                    const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(dst);
                    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(table + (size_t)row * 32), "r"(saddr) : "memory");
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                slot = (slot + 1) % DEPTH;
            } else {
                red4(table + (size_t)row * 32 + l8 * 4, 1.0f);
            }
        }
    }
    if (use_tma && l8 == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <typename F>
static float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);
    const int SMS = prop.multiProcessorCount;
    float* sink; CK(cudaMalloc(&sink, 64));
    const size_t big_rows = 22323ull * 8 * 8;  // one batch-8 value tensor: rows of 128 B
    uint4* table; CK(cudaMalloc(&table, big_rows * 128)); CK(cudaMemset(table, 0, big_rows * 128));
    const int threads = 256, blocks = SMS * 16, iters = 64;
    const double nthreads = (double)threads * blocks;

    struct { const char* name; uint32_t rows; uint32_t stride16; } cfgs[] = {
        {"183MB-table", (uint32_t)big_rows, 8}, {"2.9MB-table", 22323u, 8}, {"35KB-table", 273u, 8}};
    for (auto& c : cfgs) {
        float ms = time_ms([&] { gather_rows<8><<<blocks, threads>>>(table, c.rows, c.stride16, iters, sink); });
        double rows = nthreads / 8 * iters * 8;
        printf("gather 128B rows (8 lanes x 16B) random %-12s : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
        ms = time_ms([&] { gather_rows<4><<<blocks, threads>>>(table, c.rows, c.stride16, iters, sink); });
        rows = nthreads / 4 * iters * 8;
        printf("gather  64B rows (4 lanes x 16B) random %-12s : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 64 / ms / 1e9);
    }
    for (uint32_t lg : {15u, 20u}) {
        const uint32_t mask = (1u << lg) - 1;
        for (int ctas_per_sm : {2, 4, 8}) {
            const int blk = SMS * ctas_per_sm;
            float ms = time_ms([&] { gather_rows_lean<8><<<blk, 256>>>(table, mask, 256 / ctas_per_sm, sink); });
            double rows = (double)blk * 256 / 8 * (256 / ctas_per_sm) * 8;
            printf("lean gather 128B rows, 2^%u rows, %d CTAs/SM x 256 thr,  8 in flight : %8.3f ms  %7.2f Grows/s\n", lg, ctas_per_sm, ms, rows / ms / 1e6);
            ms = time_ms([&] { gather_rows_lean<16><<<blk, 256>>>(table, mask, 128 / ctas_per_sm, sink); });
            rows = (double)blk * 256 / 8 * (128 / ctas_per_sm) * 16;
            printf("lean gather 128B rows, 2^%u rows, %d CTAs/SM x 256 thr, 16 in flight : %8.3f ms  %7.2f Grows/s\n", lg, ctas_per_sm, ms, rows / ms / 1e6);
        }
    }
    for (auto& c : cfgs) {
        const float* ft = reinterpret_cast<const float*>(table);
        float ms = time_ms([&] { gather_rows_narrow<1><<<blocks, threads>>>(ft, c.rows, iters, sink); });
        double rows = nthreads / 32 * iters * 8;
        printf("gather 128B rows (32 lanes x 4B, 1 row/instr) %-12s : %8.3f ms  %7.2f Grows/s\n", c.name, ms, rows / ms / 1e6);
        ms = time_ms([&] { gather_rows_narrow<2><<<blocks, threads>>>(ft, c.rows, iters, sink); });
        rows = nthreads / 16 * iters * 8;
        printf("gather 128B rows (16 lanes x 8B, 2 rows/instr) %-12s : %8.3f ms  %7.2f Grows/s\n", c.name, ms, rows / ms / 1e6);
    }
    {
        float ms = time_ms([&] { gather_rows_local<8><<<blocks, threads>>>(table, 22323u * 8, 8, iters, sink); });
        double rows = nthreads / 8 * iters * 8;
        printf("gather 128B rows, spatially coherent (L1 reuse)        : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
        ms = time_ms([&] { gather_rows_local<4><<<blocks, threads>>>(table, 22323u * 8, 8, iters, sink); });
        rows = nthreads / 4 * iters * 8;
        printf("gather  64B rows, spatially coherent (L1 reuse)        : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", ms, rows / ms / 1e6, rows * 64 / ms / 1e9);
    }
    float* ftab = reinterpret_cast<float*>(table);
    struct { const char* name; uint32_t rows; } rcfgs[] = {{"183MB (no contention)", (uint32_t)big_rows}, {"17k rows (level-2-like, 64 tables)", 1050u * 64}, {"17k rows/4 (level-3-like)", 273u * 64}, {"273 rows (one hot table)", 273u}};
    for (auto& c : rcfgs) {
        const int it2 = 16;
        float ms = time_ms([&] { red_rows<0><<<blocks, threads>>>(ftab, c.rows, it2); });
        double rows = nthreads / 8 * it2 * 8;
        printf("red.v4.f32 128B rows, %-36s : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
        ms = time_ms([&] { red_rows<1><<<blocks, threads>>>(ftab, c.rows, it2); });
        rows = nthreads / 32 * it2 * 8;
        printf("atomicAdd f32 128B rows, %-33s : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
    }
    for (auto& c : rcfgs) {
        const int it2 = 16;
        constexpr int DEPTH = 4;
        const size_t smem = (size_t)(threads / 32) * DEPTH * 4 * 32 * sizeof(float);
        CK(cudaFuncSetAttribute(tma_red_rows<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float ms = time_ms([&] { tma_red_rows<DEPTH><<<blocks, threads, smem>>>(ftab, c.rows, it2); });
        double rows = nthreads / 8 * it2 * 8;
        printf("TMA cp.reduce.async.bulk 128B rows, %-22s : %8.3f ms  %7.2f Grows/s  %7.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
    }
    {   // LSU red.v4 and TMA bulk reduce at the same time: independent ceilings?
        const int it2 = 16;
        constexpr int DEPTH = 4;
        const size_t smem = (size_t)(threads / 32) * DEPTH * 4 * 32 * sizeof(float);
        struct { const char* name; uint32_t rows; } mc[] = {{"17k rows (level-2-like, 64 tables)", 1050u * 64}, {"183MB (no contention)", (uint32_t)big_rows}};
        for (auto& c : mc) {
            float ms; double rows = nthreads / 8 * it2 * 8;
#define MIXED(Q) CK(cudaFuncSetAttribute(mixed_red_rows<Q, DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            ms = time_ms([&] { mixed_red_rows<Q, DEPTH><<<blocks, threads, smem>>>(ftab, c.rows, it2); }); \
            printf("mixed scatter, %d of 4 groups via TMA reduce, %-36s : %8.3f ms  %7.2f Grows/s\n", Q, c.name, ms, rows / ms / 1e6);
            MIXED(0) MIXED(1) MIXED(2) MIXED(3) MIXED(4)
#undef MIXED
        }
    }
    {
        const int nrows = 1024;  // 128 KB table per CTA
        CK(cudaFuncSetAttribute(smem_rows<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrows * 128));
        CK(cudaFuncSetAttribute(smem_rows<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, nrows * 128));
        const int it2 = 64, th = 512;
        float ms = time_ms([&] { smem_rows<0><<<SMS, th, nrows * 128>>>(nrows, it2, sink); });
        double rows = (double)SMS * th / 32 * it2 * 8;
        printf("smem atomicAdd 128B rows (16 warps/SM)                  : %8.3f ms  %7.2f Grows/s  (%.2f clk/row/SM at %d kHz)\n", ms, rows / ms / 1e6, ms * 1e-3 * prop.clockRate * 1e3 / (rows / SMS), prop.clockRate);
        ms = time_ms([&] { smem_rows<1><<<SMS, th, nrows * 128>>>(nrows, it2, sink); });
        printf("smem ld+add+st 128B rows (racy, bandwidth bound only)   : %8.3f ms  %7.2f Grows/s  (%.2f clk/row/SM)\n", ms, rows / ms / 1e6, ms * 1e-3 * prop.clockRate * 1e3 / (rows / SMS));
    }
    {
        const int it2 = 4096;
        float ms = time_ms([&] { ffma_chain<<<SMS * 8, 256>>>(it2, sink, 1.0001f, 0.5f); });
        double fma = (double)SMS * 8 * 256 * it2 * 16;
        printf("FFMA  (16 indep chains)  : %8.3f ms  %7.2f TFMA/s\n", ms, fma / ms / 1e9);
        ms = time_ms([&] { ffma2_chain<<<SMS * 8, 256>>>(it2, sink, 1.0001f, 0.5f); });
        printf("FFMA2 (fma.rn.f32x2)     : %8.3f ms  %7.2f TFMA/s\n", ms, fma / ms / 1e9);
        ms = time_ms([&] { mufu_chain<<<SMS * 8, 256>>>(it2, sink); });
        double mufu = (double)SMS * 8 * 256 * it2 * 16;
        printf("MUFU sin+cos             : %8.3f ms  %7.2f T mufu/s\n", ms, mufu / ms / 1e9);
        ms = time_ms([&] { sincosf_chain<<<SMS * 8, 256>>>(512, sink); });
        double sc = (double)SMS * 8 * 256 * 512 * 4;
        printf("sincosf accurate (|x|~1e2..1e3) : %8.3f ms  %7.3f T sincos/s\n", ms, sc / ms / 1e9);
    }
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    printf("done\n");
    return 0;
}
