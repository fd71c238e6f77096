// msda_fwd_tile.cu -- tiled multi-scale deformable attention forward for the encoder (Nq == S), sm_100a.
//
// The flat forward (msda_fwd.cu) gathers every corner row through L1: one 128-byte line per wavefront, 4 lines per
// LDG.128 instruction, 0.55-0.59 rows per clock and SM whether the line hits or not (profiles/r01_microbench.txt),
// i.e. 91.4 M corner rows = 0.53 ms at configs[1] however good the locality is.  Shared memory delivers
// 128 B/clk/SM with no tag stage.  With the decomposition of msda_tile.cuh (a CTA = one head x an 8x8 query tile) the
// value rows a tile needs form a small window per level; the CTA copies the windows that fit its budget into
// shared memory once (each row is then read ~5-17 times) and gathers from there; levels whose window does not fit
// (no locality) are gathered from global memory exactly as the flat kernel does.
//
// STATUS (round 2, measured on B200, profiles/r02a..r02b): parity-green but SLOWER than the flat kernel -- 1.44 ms vs
// 0.58-0.61 ms at configs[1].  The flat kernel already issues at 76 % of the schedulers' rate; this one needs 518 warp
// instructions per (query, head) against 344 (records, windows, staging, four barriers per item) at 25 % occupancy and
// its time scales with 1 / (CTAs per SM): latency bound.  Staging through the LSU also costs the L1 wavefronts it was
// meant to save (a TMA box copy would not).  Only used when rdetr_msda_set_tile_mode(2) / RDETR_MSDA_TILE=2 asks for it.
#include "msda_tile.cuh"

namespace rdetr {

int msda_tile_mode();
int msda_tile_rows();

constexpr int kFwdTileThreads = 256;

template <typename VT, int CH, int L, int P, typename IO, int MINB>
__global__ void __launch_bounds__(kFwdTileThreads, MINB)
msda_fwd_tile_kernel(const VT *__restrict__ value, const int64_t *__restrict__ spatial_shapes,
                     const int64_t *__restrict__ level_start_index, const IO io, VT *__restrict__ out, int B, int S, int M, int Nq,
                     int cap_rows)
{
    constexpr int D = 32, kLanes = D / CH;
    constexpr int LP = L * P;
    constexpr int G = LP <= 16 ? 16 : 32;
    constexpr int kRecStride = LP + 1;
    constexpr int kGroups = kFwdTileThreads / kLanes;  // (query, head) pairs gathered at a time
    using SL = Slice<VT, CH>;
    static_assert(kTileQ % kGroups == 0 || kGroups % kTileQ == 0, "tile / group mismatch");

    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4 *s_rec = reinterpret_cast<float4 *>(smem_raw);              // [kTileQ][LP+1]
    VT *s_val = reinterpret_cast<VT *>(s_rec + kTileQ * kRecStride);   // [cap_rows][32]
    __shared__ TileGeom geo;
    __shared__ int s_q[kTileQ];
    __shared__ int s_bb[kMaxLevels][4];
    __shared__ TilePlace s_place[kMaxLevels];

    const int tid = threadIdx.x;
    if (tid == 0) tile_geom_init(geo, spatial_shapes, level_start_index, L, Nq);
    if (tid < L) { s_bb[tid][0] = kBBoxEmptyMin; s_bb[tid][1] = kBBoxEmptyMax; s_bb[tid][2] = kBBoxEmptyMin; s_bb[tid][3] = kBBoxEmptyMax; }
    __syncthreads();

    const long long items = (long long)B * geo.ntiles * M;
    const int pix_stride = M * D;
    for (long long item = blockIdx.x; item < items; item += gridDim.x) {
        const int m = (int)(item % M);
        const long long bt = item / M;
        const int tile = (int)(bt % geo.ntiles);
        const int b = (int)(bt / geo.ntiles);
        const VT *vhead = value + ((long long)b * S * M + m) * D;

        tile_phase1<L, P, G, kFwdTileThreads>(io, geo, b, m, tile, S, M, Nq, s_rec, s_q, s_bb);
        __syncthreads();
        if (tid == 0) tile_place_levels<L>(s_bb, s_place, cap_rows);
        __syncthreads();

        // ---- stage the resident windows: 16 bytes per thread, kLanes threads per (pixel, head) row -------
        {
            const int sl = tid % kLanes;
#pragma unroll 1
            for (int l = 0; l < L; ++l) {
                const TilePlace pl = s_place[l];
                if (!pl.resident) continue;
                const int rows = pl.bw * pl.bh;
                const int Wl = geo.W[l], st = geo.start[l];
#pragma unroll 4
                for (int r = tid / kLanes; r < rows; r += kFwdTileThreads / kLanes) {
                    const int ry = r / pl.bw, rx = r - ry * pl.bw;
                    const long long pix = st + (long long)(pl.y0 + ry) * Wl + (pl.x0 + rx);
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(vhead + pix * pix_stride + sl * CH));
                    *reinterpret_cast<uint4 *>(s_val + (long long)(pl.off + r) * D + sl * CH) = v;
                }
            }
        }
        __syncthreads();

        // ---- gather: kLanes lanes x CH channels per (query, head) ----------------------------------------
        {
            const int grp = tid / kLanes, sl = tid % kLanes;
            int lvW[L], lvStart[L], lvRow0[L], lvBw[L];
            unsigned resident = 0;
#pragma unroll
            for (int l = 0; l < L; ++l) {
                const TilePlace pl = s_place[l];
                lvW[l] = geo.W[l];
                lvStart[l] = geo.start[l];
                lvBw[l] = pl.bw;
                lvRow0[l] = pl.off - pl.y0 * pl.bw - pl.x0;  // window row of pixel (h, w) = lvRow0 + h*bw + w
                if (pl.resident) resident |= 1u << l;
            }
            const VT *vbase = vhead + sl * CH;
            const VT *sbase = s_val + sl * CH;
#pragma unroll 1
            for (int slot = grp; slot < kTileQ; slot += kGroups) {
                const int q = s_q[slot];
                const float4 *rec = s_rec + slot * kRecStride;
                float acc[CH];
#pragma unroll
                for (int c = 0; c < CH; ++c) acc[c] = 0.f;
#pragma unroll
                for (int lp = 0; lp < LP; ++lp) {
                    const int l = lp / P;
                    const float4 r = rec[lp];
                    const RecView rv = unpack_rec(r.x);
                    const float lw = r.y, lh = r.z, a = r.w;
                    const float hw = 1.f - lw, hh = 1.f - lh;
                    float v0[CH], v1[CH], v2[CH], v3[CH];
#pragma unroll
                    for (int c = 0; c < CH; ++c) v0[c] = v1[c] = v2[c] = v3[c] = 0.f;
                    if (resident & (1u << l)) {
                        const int row = lvRow0[l] + rv.h0 * lvBw[l] + rv.w0;
                        const VT *p0 = sbase + (long long)row * D;
                        if (rv.vm & 1u) SL::load_shared(p0, v0);
                        if (rv.vm & 2u) SL::load_shared(p0 + D, v1);
                        if (rv.vm & 4u) SL::load_shared(p0 + lvBw[l] * D, v2);
                        if (rv.vm & 8u) SL::load_shared(p0 + lvBw[l] * D + D, v3);
                    } else {
                        const long long base = (long long)(lvStart[l] + rv.h0 * lvW[l] + rv.w0) * pix_stride;
                        if (rv.vm & 1u) SL::load(vbase + base, v0);
                        if (rv.vm & 2u) SL::load(vbase + base + pix_stride, v1);
                        if (rv.vm & 4u) SL::load(vbase + base + (long long)lvW[l] * pix_stride, v2);
                        if (rv.vm & 8u) SL::load(vbase + base + (long long)lvW[l] * pix_stride + pix_stride, v3);
                    }
                    const float w0 = a * (hh * hw), w1 = a * (hh * lw), w2 = a * (lh * hw), w3 = a * (lh * lw);
#pragma unroll
                    for (int c = 0; c < CH; ++c) {
                        acc[c] = fmaf(w0, v0[c], acc[c]);
                        acc[c] = fmaf(w1, v1[c], acc[c]);
                        acc[c] = fmaf(w2, v2[c], acc[c]);
                        acc[c] = fmaf(w3, v3[c], acc[c]);
                    }
                }
                if (q >= 0) SL::store(out + (((long long)b * Nq + q) * M + m) * D + sl * CH, acc);
            }
        }
        __syncthreads();  // the next item's phase 1 / staging overwrite s_rec and s_val
    }
}

template <typename VT, int CH, int L, int P, typename IO>
static int launch_fwd_tile_lp(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, void *out, int B, int S,
                              int M, int Nq, cudaStream_t stream)
{
    constexpr int MINB = 2;
    auto kern = msda_fwd_tile_kernel<VT, CH, L, P, IO, MINB>;
    int cap_rows = msda_tile_rows();
    if (cap_rows <= 0) cap_rows = sizeof(VT) == 4 ? 640 : 1024;
    const size_t smem = (size_t)kTileQ * (L * P + 1) * sizeof(float4) + (size_t)cap_rows * 32 * sizeof(VT);
    if (int rc = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                            "cudaFuncSetAttribute(msda_fwd_tile)"))
        return rc;
    int dev = 0, sms = 0, occ = 0;
    if (int rc = check_cuda(cudaGetDevice(&dev), "cudaGetDevice")) return rc;
    if (int rc = check_cuda(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev), "cudaDeviceGetAttribute")) return rc;
    if (int rc = check_cuda(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kFwdTileThreads, smem), "occupancy(msda_fwd_tile)"))
        return rc;
    if (occ < 1) return fail(RDETR_ERR_UNSUPPORTED, "msda_forward (tiled): %zu bytes of shared memory do not fit an SM", smem);
    const long long grid = (long long)sms * occ;  // persistent CTAs; the item count is only known on the device
    kern<<<(unsigned)grid, kFwdTileThreads, smem, stream>>>(static_cast<const VT *>(value), shapes, lsi, io, static_cast<VT *>(out), B, S,
                                                             M, Nq, cap_rows);
    return check_cuda(cudaGetLastError(), "msda_fwd_tile_kernel launch");
}

// Returns -1 when the shape is outside what the tiled kernel is built for (caller uses the flat kernel).
template <typename VT, int CH, typename IO>
int launch_fwd_tile(const void *value, const int64_t *shapes, const int64_t *lsi, const IO &io, void *out, int B, int S, int M, int L,
                    int Nq, int P, cudaStream_t stream)
{
    if (P != 4) return -1;
    if (L == 4) return launch_fwd_tile_lp<VT, CH, 4, 4, IO>(value, shapes, lsi, io, out, B, S, M, Nq, stream);
    if (L == 5) return launch_fwd_tile_lp<VT, CH, 5, 4, IO>(value, shapes, lsi, io, out, B, S, M, Nq, stream);
    return -1;
}

template int launch_fwd_tile<float, 4, PlainIO>(const void *, const int64_t *, const int64_t *, const PlainIO &, void *, int, int, int, int,
                                                int, int, cudaStream_t);
template int launch_fwd_tile<__nv_bfloat16, 8, PlainIO>(const void *, const int64_t *, const int64_t *, const PlainIO &, void *, int, int,
                                                        int, int, int, int, cudaStream_t);
template int launch_fwd_tile<float, 4, FusedIO<float>>(const void *, const int64_t *, const int64_t *, const FusedIO<float> &, void *, int,
                                                       int, int, int, int, int, cudaStream_t);
template int launch_fwd_tile<__nv_bfloat16, 8, FusedIO<__nv_bfloat16>>(const void *, const int64_t *, const int64_t *,
                                                                       const FusedIO<__nv_bfloat16> &, void *, int, int, int, int, int,
                                                                       int, cudaStream_t);

}  // namespace rdetr
