// microbench_red.cu -- what a vector reduction to L2 costs as a function of its width (B200).  Not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench_red tools/microbench_red.cu
// Question: would bf16 reductions (64 bytes per 32-channel row) run faster than fp32 ones (128 bytes)?
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// MODE 0: red.v4.f32, 8 lanes x 16 B = one 128-byte row, 4 rows per warp instruction
// MODE 1: red.v2.f32, 16 lanes x 8 B = one 128-byte row, 2 rows per instruction
// MODE 2: red.v4.bf16x2, 4 lanes x 16 B = one 64-byte row (32 bf16 channels), 8 rows per instruction
// MODE 3: red.v2.bf16x2, 8 lanes x 8 B = one 64-byte row, 4 rows per instruction
// MODE 4: red.v4.f32 on HALF rows: 4 lanes x 16 B = 64 bytes, 8 half-rows per instruction (16 channels of fp32)
template <int MODE>
__global__ void red_rows(unsigned char* table, uint32_t nrows, int iters)
{
    constexpr int LANES = MODE == 0 ? 8 : MODE == 1 ? 16 : MODE == 2 ? 4 : MODE == 3 ? 8 : 4;
    constexpr int ROWB = (MODE == 0 || MODE == 1) ? 128 : 64;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t group = tid / LANES, lane = tid % LANES;
    uint32_t seed = group * 2654435761u + 777u;
    const uint32_t one2 = 0x3f803f80u;  // bf16x2 {1.0, 1.0}
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            seed = hash32(seed + u);
            unsigned char* p = table + (size_t)(seed % nrows) * ROWB + lane * (ROWB / LANES);
            if (MODE == 0 || MODE == 4) asm volatile("red.global.add.v4.f32 [%0], {%1,%1,%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
            if (MODE == 1) asm volatile("red.global.add.v2.f32 [%0], {%1,%1};" ::"l"(p), "f"(1.0f) : "memory");
            if (MODE == 2) asm volatile("red.global.add.noftz.v4.bf16x2 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(one2) : "memory");
            if (MODE == 3) asm volatile("red.global.add.noftz.v2.bf16x2 [%0], {%1,%1};" ::"l"(p), "r"(one2) : "memory");
        }
    }
}

template <typename F>
static float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(a);
    for (int i = 0; i < reps; ++i) f();
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / reps;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs\n", prop.name, prop.multiProcessorCount);
    const size_t bytes = (size_t)183 << 20;
    unsigned char* table; CK(cudaMalloc(&table, bytes)); CK(cudaMemset(table, 0, bytes));
    const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 16;
    const double nthreads = (double)threads * blocks;
    struct { const char* name; uint32_t rows; } cfgs[] = {{"1.43M rows (no contention)", 1430000u}, {"67k rows (level-2-like, 64 planes)", 1050u * 64}, {"17k rows (level-3-like)", 273u * 64}};
    for (auto& c : cfgs) {
        float ms; double rows;
        ms = time_ms([&] { red_rows<0><<<blocks, threads>>>(table, c.rows, iters); }); rows = nthreads / 8 * iters * 8;
        printf("%-36s red.v4.f32     128B row, 4 rows/instr : %7.3f ms %7.2f Grows/s %6.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
        ms = time_ms([&] { red_rows<1><<<blocks, threads>>>(table, c.rows, iters); }); rows = nthreads / 16 * iters * 8;
        printf("%-36s red.v2.f32     128B row, 2 rows/instr : %7.3f ms %7.2f Grows/s %6.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 128 / ms / 1e9);
        ms = time_ms([&] { red_rows<2><<<blocks, threads>>>(table, c.rows, iters); }); rows = nthreads / 4 * iters * 8;
        printf("%-36s red.v4.bf16x2   64B row, 8 rows/instr : %7.3f ms %7.2f Grows/s %6.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 64 / ms / 1e9);
        ms = time_ms([&] { red_rows<3><<<blocks, threads>>>(table, c.rows, iters); }); rows = nthreads / 8 * iters * 8;
        printf("%-36s red.v2.bf16x2   64B row, 4 rows/instr : %7.3f ms %7.2f Grows/s %6.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 64 / ms / 1e9);
        ms = time_ms([&] { red_rows<4><<<blocks, threads>>>(table, c.rows, iters); }); rows = nthreads / 4 * iters * 8;
        printf("%-36s red.v4.f32 64B half row, 8 per instr  : %7.3f ms %7.2f Ghalf-rows/s %6.2f TB/s\n", c.name, ms, rows / ms / 1e6, rows * 64 / ms / 1e9);
    }
    printf("done\n");
    return 0;
}
