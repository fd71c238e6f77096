// memfuse.cu -- the encoder's memory_fusion input Linear as a K-split GEMM on the 5th-generation tensor cores
// (SURVEY.md section 8f, row N4).
//
// Upstream (models/bricks/relation_transformer.py:168-173, 203-204):
//     query = torch.cat(queries, -1)            # 7 x [B, S, 256] -> [B, S, 1792]: 1.28 GB written, read straight back
//     query = memory_fusion(query)              # Linear(1792, 256) -> ReLU -> Linear(256, 256) -> LayerNorm
// and autograd keeps the concatenated copy for the weight gradient.  Here the first Linear reads the seven encoder
// states IN PLACE: out = relu(sum_t X_t W_t^T + b) with W_t = W[:, 256 t : 256 (t+1)].  The contraction is dense
// (164 GFLOP at B = 8, K = 1792), so unlike the rest of this library it runs on tcgen05:
//
//   * one CTA per 128-row tile of the M = B*S rows, full N = 256 columns: the accumulator is 128 lanes x 256 columns of
//     TMEM (tcgen05.alloc of 256 columns), written by tcgen05.mma.cta_group::1.kind::tf32 (M 128, N 256, K 8 per issue);
//   * warp 0: one elected thread issues TMA loads (cp.async.bulk.tensor.2d, 128-byte swizzle) of the A tile [128 rows x 32
//     floats] of the current source and of the W tile [256 rows x 32 floats] into a 4-stage ring (48 KB per stage),
//     completion by mbarrier transaction count; one tensor map per source, so the K loop simply walks the sources;
//   * warp 1: one elected thread issues 4 MMAs per stage from shared-memory matrix descriptors (K-major, SWIZZLE_128B,
//     8-row groups 1024 bytes apart) and releases the stage with tcgen05.commit;
//   * warps 2-5: epilogue -- tcgen05.ld.32x32b of the lane quarter each warp may address, + bias, ReLU, 128-byte row
//     stores.
// TF32 products with fp32 accumulation: closer to the fp32 truth than the bf16 GEMM autocast runs upstream, but not
// fp32-exact, so the drop-in only uses it under autocast / allow_tf32 (modules.py).  Every mbarrier wait is bounded:
// a protocol error traps instead of hanging the device.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace rdetr {

constexpr int kMfBM = 128;        // rows per CTA
constexpr int kMfBN = 256;        // output features (the whole width)
constexpr int kMfBK = 32;         // floats per stage along K = one 128-byte swizzle row
constexpr int kMfUmmaK = 8;       // tf32: 32 bytes per MMA along K
constexpr int kMfMaxSrc = 8;
constexpr int kMfThreads = 192;
constexpr uint32_t kMfBytesA = kMfBM * kMfBK * 4;  // 16 KB
constexpr uint32_t kMfBytesB = kMfBN * kMfBK * 4;  // 32 KB
constexpr uint32_t kMfStageBytes = kMfBytesA + kMfBytesB;

struct MfMaps {
    CUtensorMap a[kMfMaxSrc];
    CUtensorMap w;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait (about a second of polling): a protocol error must trap, not hang the box.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        uint32_t done;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(addr), "r"(parity) : "memory");
        if (done) return;
    }
    asm volatile("trap;");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// K-major operand tile with the 128-byte swizzle: rows are 128 bytes, 8-row groups 1024 bytes apart
// (cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::K>: LBO = 1, SBO = 64, version 1, layout SWIZZLE_128B = 2).
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fffu);   // start address, bits [0,14)
    d |= (uint64_t)1u << 16;                       // leading byte offset (unused for swizzled K-major), bits [16,30)
    d |= (uint64_t)(1024u >> 4) << 32;             // stride byte offset, bits [32,46)
    d |= (uint64_t)1u << 46;                       // descriptor version (Blackwell), bits [46,48)
    d |= (uint64_t)2u << 61;                       // SWIZZLE_128B, bits [61,64)
    return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 256 (cute/arch/mma_sm100_desc.hpp InstrDescriptor)
constexpr uint32_t kMfIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kMfBN >> 3) << 17) | ((uint32_t)(kMfBM >> 4) << 24);

// STAGES = 4: one CTA per SM (193 KB of shared memory); STAGES = 2: two CTAs per SM (97 KB each, 2 x 256 TMEM columns), so
// that one CTA's epilogue overlaps the other's main loop.
template <int kMfStages>
__global__ void __launch_bounds__(kMfThreads, kMfStages <= 2 ? 2 : 1)
memfuse_kernel(const __grid_constant__ MfMaps maps, const float *__restrict__ bias, float *__restrict__ out, int M, int nsrc,
               int kb_per_src, int src_cols, int relu)
{
    extern __shared__ unsigned char mf_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)mf_raw + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024-byte aligned
    __shared__ __align__(8) uint64_t full_bar[kMfStages], empty_bar[kMfStages], tmem_full_bar;
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kMfBM;
    const int nkb = nsrc * kb_per_src;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kMfStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the allocating warp also deallocates
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)kMfBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ------------------------------------------------------------------
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kMfStages;
                const uint32_t ph = (kb / kMfStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);  // passes at once on a fresh barrier
                mbar_expect_tx(&full_bar[s], kMfStageBytes);
                const int src = kb / kb_per_src, kc = kb - src * kb_per_src;
                unsigned char *st = tiles + (size_t)s * kMfStageBytes;
                tma_load_2d(st, &maps.a[src], kc * kMfBK, m0, &full_bar[s]);
                tma_load_2d(st + kMfBytesA, &maps.w, src * src_cols + kc * kMfBK, 0, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer --------------------------------------------------------------------
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % kMfStages;
                const uint32_t ph = (kb / kMfStages) & 1;
                mbar_wait(&full_bar[s], ph);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a_addr = smem_u32(tiles + (size_t)s * kMfStageBytes);
                const uint32_t b_addr = a_addr + kMfBytesA;
#pragma unroll
                for (int k = 0; k < kMfBK / kMfUmmaK; ++k) {
                    const uint64_t adesc = umma_desc_k_sw128(a_addr + k * kMfUmmaK * 4);
                    const uint64_t bdesc = umma_desc_k_sw128(b_addr + k * kMfUmmaK * 4);
                    const uint32_t accumulate = (kb | k) != 0 ? 1u : 0u;
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(tmem_base), "l"(adesc), "l"(bdesc), "r"(kMfIdesc), "r"(accumulate) : "memory");
                }
                // frees the stage once the MMAs that read it have completed (implies fence::before_thread_sync)
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tmem_full_bar)) : "memory");
        }
    } else {
        // ---- epilogue: warp w may address TMEM lanes 32 (w % 4) .. + 31 -------------------------------------------
        const int quarter = warp & 3;
        mbar_wait(&tmem_full_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int row = m0 + quarter * 32 + lane;
        float *orow = out + (size_t)row * kMfBN;
#pragma unroll 1
        for (int cc = 0; cc < kMfBN / 32; ++cc) {
            uint32_t r[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cc * 32);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                         "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                           "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                           "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                           "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                         : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (row < M) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + cc * 32 + j));
                    float4 v = make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                                           __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
                    if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    *reinterpret_cast<float4 *>(orow + cc * 32 + j) = v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)kMfBN) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... with a 4-stage operand ring that runs
// on ACROSS tiles and TWO 256-column accumulators in TMEM (all 512 columns): while warps 2-5 drain accumulator (i & 1) of tile i
// (tcgen05.ld -> bias -> ReLU -> stores), warp 1 is already issuing the MMAs of tile i + 1 into the other one.  The 2-CTA-per-SM
// kernel above gets its overlap from the second resident CTA but leaves each CTA only 2 stages (96 KB) of loads in flight.
// Barriers: full / empty per stage (TMA <-> MMA), tmem_full / tmem_empty per accumulator (MMA <-> epilogue; tmem_empty counts
// the four epilogue warps).
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kMfPStages = 4;

__global__ void __launch_bounds__(kMfThreads, 1)
memfuse_persistent_kernel(const __grid_constant__ MfMaps maps, const float *__restrict__ bias, float *__restrict__ out, int M, int nsrc,
                          int kb_per_src, int src_cols, int relu, int ntiles)
{
    extern __shared__ unsigned char mf_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)mf_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kMfPStages], empty_bar[kMfPStages], tmem_full_bar[2], tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = nsrc * kb_per_src;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kMfPStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tmem_full_bar[a], 1); mbar_init(&tmem_empty_bar[a], 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // the allocating warp also deallocates; 512 columns = the whole TMEM of the SM (one CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2u * kMfBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer: one continuous stream of k-blocks over all tiles of this CTA -------
            uint32_t g = 0;  // running k-block counter
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = tile * kMfBM;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % kMfPStages;
                    const uint32_t ph = (g / kMfPStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);  // passes at once on a fresh barrier
                    mbar_expect_tx(&full_bar[s], kMfStageBytes);
                    const int src = kb / kb_per_src, kc = kb - src * kb_per_src;
                    unsigned char *st = tiles + (size_t)s * kMfStageBytes;
                    tma_load_2d(st, &maps.a[src], kc * kMfBK, m0, &full_bar[s]);
                    tma_load_2d(st + kMfBytesA, &maps.w, src * src_cols + kc * kMfBK, 0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer --------------------------------------------------------------------
            uint32_t g = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tmem_empty_bar[acc], ((it >> 1) & 1) ^ 1);  // the epilogue has drained this accumulator (fresh: passes)
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d_addr = tmem_base + (uint32_t)(acc * kMfBN);
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % kMfPStages;
                    const uint32_t ph = (g / kMfPStages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * kMfStageBytes);
                    const uint32_t b_addr = a_addr + kMfBytesA;
#pragma unroll
                    for (int k = 0; k < kMfBK / kMfUmmaK; ++k) {
                        const uint64_t adesc = umma_desc_k_sw128(a_addr + k * kMfUmmaK * 4);
                        const uint64_t bdesc = umma_desc_k_sw128(b_addr + k * kMfUmmaK * 4);
                        const uint32_t accumulate = (kb | k) != 0 ? 1u : 0u;
                        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                     "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                     ::"r"(d_addr), "l"(adesc), "l"(bdesc), "r"(kMfIdesc), "r"(accumulate) : "memory");
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tmem_full_bar[acc])) : "memory");
            }
        }
    } else {
        // ---- epilogue: warp w may address TMEM lanes 32 (w % 4) .. + 31 -------------------------------------------
        const int quarter = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            mbar_wait(&tmem_full_bar[acc], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int row = tile * kMfBM + quarter * 32 + lane;
            float *orow = out + (size_t)row * kMfBN;
#pragma unroll 1
            for (int cc = 0; cc < kMfBN / 32; ++cc) {
                uint32_t r[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kMfBN + cc * 32);
                asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                             "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                             "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                             : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                               "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                               "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                               "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                             : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == kMfBN / 32 - 1) {
                    // every column of this accumulator is in registers: hand it back to the MMA warp before the stores
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[acc])) : "memory");
                }
                if (row < M) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + cc * 32 + j));
                        float4 v = make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                                               __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
                        if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        *reinterpret_cast<float4 *>(orow + cc * 32 + j) = v;
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kMfBN) : "memory");
}

// ------------------------------------------------------------------------------------------------------------------------
// Persistent variant with 256-row super-tiles: the two TMEM accumulators hold the two 128-row halves of ONE tile, so a stage
// carries A [256 x 32] + W [256 x 32] = 64 KB for 8 MMAs (8 KB per MMA instead of 12 KB: every W byte that crosses the L2 -> SM
// fabric is used by twice as many rows).  The epilogue drains accumulator 0, hands it back, then accumulator 1; the MMA warp
// of the next tile needs accumulator 1 only after its first four MMAs, so what is lost of the overlap is about one drain per tile.
// ------------------------------------------------------------------------------------------------------------------------
constexpr int kMfWStages = 3;
constexpr uint32_t kMfWideStageBytes = 2 * kMfBytesA + kMfBytesB;   // 64 KB

__global__ void __launch_bounds__(kMfThreads, 1)
memfuse_wide_kernel(const __grid_constant__ MfMaps maps, const float *__restrict__ bias, float *__restrict__ out, int M, int nsrc,
                    int kb_per_src, int src_cols, int relu, int ntiles)
{
    extern __shared__ unsigned char mf_raw[];
    unsigned char *tiles = reinterpret_cast<unsigned char *>(((uintptr_t)mf_raw + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t full_bar[kMfWStages], empty_bar[kMfWStages], tmem_full_bar, tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_s;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nkb = nsrc * kb_per_src;

    if (warp == 0 && lane == 0) {
        for (int s = 0; s < kMfWStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&tmem_full_bar, 1);
        for (int a = 0; a < 2; ++a) mbar_init(&tmem_empty_bar[a], 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(2u * kMfBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer ------------------------------------------------------------------
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int m0 = tile * 2 * kMfBM;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % kMfWStages;
                    const uint32_t ph = (g / kMfWStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], kMfWideStageBytes);
                    const int src = kb / kb_per_src, kc = kb - src * kb_per_src;
                    unsigned char *st = tiles + (size_t)s * kMfWideStageBytes;
                    tma_load_2d(st, &maps.a[src], kc * kMfBK, m0, &full_bar[s]);   // box of 256 rows: both halves, 16 KB apart
                    tma_load_2d(st + 2 * kMfBytesA, &maps.w, src * src_cols + kc * kMfBK, 0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer --------------------------------------------------------------------
            uint32_t g = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
                const uint32_t drained = (uint32_t)(it & 1) ^ 1u;   // parity of "the epilogue of the previous tile has drained it"
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % kMfWStages;
                    const uint32_t ph = (g / kMfWStages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_addr = smem_u32(tiles + (size_t)s * kMfWideStageBytes);
                    const uint32_t b_addr = a_addr + 2 * kMfBytesA;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        if (kb == 0) {
                            mbar_wait(&tmem_empty_bar[half], drained);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        }
                        const uint32_t d_addr = tmem_base + (uint32_t)(half * kMfBN);
#pragma unroll
                        for (int k = 0; k < kMfBK / kMfUmmaK; ++k) {
                            const uint64_t adesc = umma_desc_k_sw128(a_addr + half * kMfBytesA + k * kMfUmmaK * 4);
                            const uint64_t bdesc = umma_desc_k_sw128(b_addr + k * kMfUmmaK * 4);
                            const uint32_t accumulate = (kb | k) != 0 ? 1u : 0u;
                            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                                         "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                                         ::"r"(d_addr), "l"(adesc), "l"(bdesc), "r"(kMfIdesc), "r"(accumulate) : "memory");
                        }
                    }
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&empty_bar[s])) : "memory");
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&tmem_full_bar)) : "memory");
            }
        }
    } else {
        // ---- epilogue: warp w may address TMEM lanes 32 (w % 4) .. + 31 of both accumulators ---------------------
        const int quarter = warp & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            mbar_wait(&tmem_full_bar, it & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int half = 0; half < 2; ++half) {
                const int row = tile * 2 * kMfBM + half * kMfBM + quarter * 32 + lane;
                float *orow = out + (size_t)row * kMfBN;
#pragma unroll 1
                for (int cc = 0; cc < kMfBN / 32; ++cc) {
                    uint32_t r[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * kMfBN + cc * 32);
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                                 : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (cc == kMfBN / 32 - 1) {  // this accumulator is in registers: hand it back before the stores
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&tmem_empty_bar[half])) : "memory");
                    }
                    if (row < M) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 bv = __ldg(reinterpret_cast<const float4 *>(bias + cc * 32 + j));
                            float4 v = make_float4(__uint_as_float(r[j]) + bv.x, __uint_as_float(r[j + 1]) + bv.y,
                                                   __uint_as_float(r[j + 2]) + bv.z, __uint_as_float(r[j + 3]) + bv.w);
                            if (relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            *reinterpret_cast<float4 *>(orow + cc * 32 + j) = v;
                        }
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2u * kMfBN) : "memory");
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = nullptr;  // resolved once; the driver library is already loaded by the runtime
    if (fn == nullptr) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// [rows, cols] fp32 row-major (row pitch `pitch_elems`), boxes of [box_rows x 32 floats], 128-byte swizzle
static int make_map(CUtensorMap *map, const float *base, uint64_t rows, uint64_t cols, uint64_t pitch_elems, uint32_t box_rows)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(RDETR_ERR_CUDA, "memory_fusion: cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {pitch_elems * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kMfBK, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    // L2 promotion: a 128-byte box row is one k-block of a source row; promoting the fill to 256 bytes would bring the NEXT
    // k-block of the same row into L2 with the same DRAM access -- measured: no difference (RDETR_MEMFUSE_L2PROMO = 64 / 128 / 256)
    static const CUtensorMapL2promotion promo = [] {
        const char *e = getenv("RDETR_MEMFUSE_L2PROMO");
        const int v = e ? atoi(e) : 128;
        return v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
    }();
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(RDETR_ERR_CUDA, "memory_fusion: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r);
    return RDETR_OK;
}

}  // namespace rdetr

extern "C" int rdetr_memory_fusion_forward(const float *const *sources, int nsrc, const float *weight, const float *bias, float *out,
                                           long long M, int C, int N, int relu, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (M < 0 || nsrc <= 0 || C <= 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_memory_fusion_forward: bad size (M=%lld nsrc=%d C=%d)", M, nsrc, C);
    if (N != kMfBN || nsrc > kMfMaxSrc || C % kMfBK != 0)
        return fail(RDETR_ERR_UNSUPPORTED, "rdetr_memory_fusion_forward: built for N = %d outputs, <= %d sources, C %% %d == 0 (got N=%d nsrc=%d C=%d)",
                    kMfBN, kMfMaxSrc, kMfBK, N, nsrc, C);
    if (M >= (1LL << 31) - kMfBM) return fail(RDETR_ERR_UNSUPPORTED, "rdetr_memory_fusion_forward: M=%lld too large", M);
    if (M == 0) return RDETR_OK;
    if (!sources || !weight || !bias || !out) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_memory_fusion_forward: null pointer argument");
    uintptr_t bits = (uintptr_t)weight | (uintptr_t)bias | (uintptr_t)out;
    for (int t = 0; t < nsrc; ++t) {
        if (!sources[t]) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_memory_fusion_forward: source %d is null", t);
        bits |= (uintptr_t)sources[t];
    }
    if (bits & 15) return fail(RDETR_ERR_INVALID_ARGUMENT, "rdetr_memory_fusion_forward: buffers must be 16-byte aligned");
    const DeviceGuard guard(out);
    if (guard.status()) return guard.status();
    // RDETR_MEMFUSE_STAGES: 0 (default) = persistent kernel (1 CTA / SM, 4 stages, two TMEM accumulators = two tiles in flight);
    // 8 = persistent kernel with 256-row super-tiles (two accumulators = two halves of one tile, W traffic halved per row);
    // 2 = one tile per CTA, 2 CTAs / SM, 2 stages; 4 = one tile per CTA, 1 CTA / SM, 4 stages (A/B measurements)
    static const int stages = [] { const char *e = getenv("RDETR_MEMFUSE_STAGES"); const int v = e ? atoi(e) : 0; return (v == 4 || v == 2 || v == 8) ? v : 0; }();
    MfMaps maps;
    memset(&maps, 0, sizeof(maps));
    for (int t = 0; t < nsrc; ++t)
        if (int rc = make_map(&maps.a[t], sources[t], (uint64_t)M, (uint64_t)C, (uint64_t)C, stages == 8 ? 2 * kMfBM : kMfBM)) return rc;
    if (int rc = make_map(&maps.w, weight, (uint64_t)N, (uint64_t)nsrc * C, (uint64_t)nsrc * C, kMfBN)) return rc;
    const unsigned grid = (unsigned)((M + kMfBM - 1) / kMfBM);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (stages == 0) {
        const size_t smem = (size_t)kMfPStages * kMfStageBytes + 1024;
        if (int rc = check_cuda(cudaFuncSetAttribute(memfuse_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(memfuse)")) return rc;
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned ctas = grid < (unsigned)sms ? grid : (unsigned)sms;
        memfuse_persistent_kernel<<<ctas, kMfThreads, smem, st>>>(maps, bias, out, (int)M, nsrc, C / kMfBK, C, relu, (int)grid);
        return check_cuda(cudaGetLastError(), "memfuse_persistent_kernel launch");
    }
    if (stages == 8) {
        const size_t smem = (size_t)kMfWStages * kMfWideStageBytes + 1024;
        if (int rc = check_cuda(cudaFuncSetAttribute(memfuse_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(memfuse)")) return rc;
        int dev = 0, sms = 148;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const unsigned wide_tiles = (unsigned)((M + 2 * kMfBM - 1) / (2 * kMfBM));
        const unsigned ctas = wide_tiles < (unsigned)sms ? wide_tiles : (unsigned)sms;
        memfuse_wide_kernel<<<ctas, kMfThreads, smem, st>>>(maps, bias, out, (int)M, nsrc, C / kMfBK, C, relu, (int)wide_tiles);
        return check_cuda(cudaGetLastError(), "memfuse_wide_kernel launch");
    }
    const size_t smem = (size_t)stages * kMfStageBytes + 1024;
    if (stages == 4) {
        if (int rc = check_cuda(cudaFuncSetAttribute(memfuse_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(memfuse)")) return rc;
        memfuse_kernel<4><<<grid, kMfThreads, smem, st>>>(maps, bias, out, (int)M, nsrc, C / kMfBK, C, relu);
    } else {
        if (int rc = check_cuda(cudaFuncSetAttribute(memfuse_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute(memfuse)")) return rc;
        memfuse_kernel<2><<<grid, kMfThreads, smem, st>>>(maps, bias, out, (int)M, nsrc, C / kMfBK, C, relu);
    }
    return check_cuda(cudaGetLastError(), "memfuse_kernel launch");
}
