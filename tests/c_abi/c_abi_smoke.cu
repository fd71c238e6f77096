// c_abi_smoke.cu -- uses librdetr_ops.so the way a non-Python host would: cudaMalloc'ed buffers, a user
// stream, plain C calls; results are checked against the C oracle (librdetr_oracle.so).  TEST ONLY.
// Build + run: see tests/test_c_abi_gpu.py.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "rdetr_ops.h"

extern "C" {
void rdetr_oracle_msda_forward_f64(const double *, const int64_t *, const int64_t *, const double *, const double *, double *, int,
                                   int, int, int, int, int, int);
void rdetr_oracle_msda_backward_f64(const double *, const int64_t *, const int64_t *, const double *, const double *,
                                    const double *, double *, double *, double *, int, int, int, int, int, int, int);
int rdetr_oracle_lsap(const double *, int64_t, int64_t, int64_t *, int64_t *);
void rdetr_oracle_rel_forward_f64(const double *, const double *, const double *, const double *, const double *, double, double,
                                  const uint8_t *, double *, int, int, int, int, int);
}

#define CK(x)                                                                               \
    do {                                                                                    \
        cudaError_t e_ = (x);                                                               \
        if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } \
    } while (0)
#define RK(x)                                                                     \
    do {                                                                          \
        int rc_ = (x);                                                            \
        if (rc_) { printf("rdetr error %d: %s (line %d)\n", rc_, rdetr_last_error(), __LINE__); return 3; } \
    } while (0)

static double frand() { return rand() / (double)RAND_MAX; }

template <typename T>
static T *to_dev(const std::vector<T> &h)
{
    T *d = nullptr;
    if (cudaMalloc(&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
    cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
    return d;
}

int main()
{
    srand(7);
    if (rdetr_abi_version() != RDETR_ABI_VERSION) { printf("ABI version mismatch\n"); return 1; }
    const int B = 2, M = 8, D = 32, L = 3, P = 4, Nq = 21;
    const int64_t hw[L][2] = {{9, 13}, {5, 7}, {2, 3}};
    std::vector<int64_t> shapes, lsi;
    int S = 0;
    for (int l = 0; l < L; ++l) { shapes.push_back(hw[l][0]); shapes.push_back(hw[l][1]); lsi.push_back(S); S += (int)(hw[l][0] * hw[l][1]); }
    const size_t nv = (size_t)B * S * M * D, ns = (size_t)B * Nq * M * L * P, no = (size_t)B * Nq * M * D;
    std::vector<float> value(nv), loc(ns * 2), attn(ns), go(no);
    for (auto &v : value) v = (float)(frand() * 2 - 1);
    for (auto &v : loc) v = (float)(frand() * 1.3 - 0.15);
    for (auto &v : attn) v = (float)frand();
    for (auto &v : go) v = (float)(frand() * 2 - 1);

    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float *d_value = to_dev(value), *d_loc = to_dev(loc), *d_attn = to_dev(attn), *d_go = to_dev(go);
    int64_t *d_shapes = to_dev(shapes), *d_lsi = to_dev(lsi);
    float *d_out, *d_gv, *d_gl, *d_ga;
    CK(cudaMalloc(&d_out, no * 4)); CK(cudaMalloc(&d_gv, nv * 4)); CK(cudaMalloc(&d_gl, ns * 8)); CK(cudaMalloc(&d_ga, ns * 4));
    RK(rdetr_msda_forward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_out, B, S, M, D, L, Nq, P, RDETR_DTYPE_F32, st));
    if (rdetr_msda_backward_workspace_bytes(B, S, M, D, L, Nq, P, RDETR_DTYPE_F32) != 0) { printf("unexpected workspace\n"); return 1; }
    RK(rdetr_msda_backward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_go, d_gv, d_gl, d_ga, B, S, M, D, L, Nq, P, RDETR_DTYPE_F32,
                           nullptr, 0, st));
    CK(cudaStreamSynchronize(st));
    std::vector<float> out(no), gv(nv), ga(ns);
    CK(cudaMemcpy(out.data(), d_out, no * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(gv.data(), d_gv, nv * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(ga.data(), d_ga, ns * 4, cudaMemcpyDeviceToHost));

    std::vector<double> v64(value.begin(), value.end()), l64(loc.begin(), loc.end()), a64(attn.begin(), attn.end()), g64(go.begin(), go.end());
    std::vector<double> o64(no), gv64(nv), gl64(ns * 2), ga64(ns);
    rdetr_oracle_msda_forward_f64(v64.data(), shapes.data(), lsi.data(), l64.data(), a64.data(), o64.data(), B, S, M, D, L, Nq, P);
    rdetr_oracle_msda_backward_f64(v64.data(), shapes.data(), lsi.data(), l64.data(), a64.data(), g64.data(), gv64.data(), gl64.data(),
                                   ga64.data(), B, S, M, D, L, Nq, P);
    double e_out = 0, e_gv = 0, e_ga = 0;
    for (size_t i = 0; i < no; ++i) e_out = fmax(e_out, fabs(out[i] - o64[i]));
    for (size_t i = 0; i < nv; ++i) e_gv = fmax(e_gv, fabs(gv[i] - gv64[i]));
    for (size_t i = 0; i < ns; ++i) e_ga = fmax(e_ga, fabs(ga[i] - ga64[i]));
    printf("msda: out %.2e grad_value %.2e grad_attn %.2e\n", e_out, e_gv, e_ga);
    if (!(e_out <= 1e-5 && e_gv <= 1e-4 && e_ga <= 1e-4)) return 4;

    // relation bias, EXACT and FAST (FAST needs the table workspace)
    const int N1 = 37, N2 = 29, H = 8;
    std::vector<float> src((size_t)B * N1 * 4), tgt((size_t)B * N2 * 4), w((size_t)H * 64), bias(H), dim_t(8);
    for (size_t i = 0; i < src.size(); ++i) src[i] = (float)((i % 4 < 2) ? frand() : 1e-3 + frand() * 0.499);
    for (size_t i = 0; i < tgt.size(); ++i) tgt[i] = (float)((i % 4 < 2) ? frand() : 1e-3 + frand() * 0.499);
    for (auto &v : w) v = (float)((frand() * 2 - 1) * 0.125);
    for (auto &v : bias) v = (float)((frand() * 2 - 1) * 0.125);
    for (int k = 0; k < 8; ++k) dim_t[k] = powf(10000.f, (float)(2 * k) / 16.f);
    float *d_src = to_dev(src), *d_tgt = to_dev(tgt), *d_w = to_dev(w), *d_b = to_dev(bias), *d_dim = to_dev(dim_t);
    const size_t nr = (size_t)B * H * N1 * N2;
    float *d_rel;
    uint32_t *d_bits;
    CK(cudaMalloc(&d_rel, nr * 4));
    CK(cudaMalloc(&d_bits, (size_t)B * N1 * ((N2 + 31) / 32) * H * 4));
    std::vector<double> s64(src.begin(), src.end()), t64(tgt.begin(), tgt.end()), w64(w.begin(), w.end()), b64(bias.begin(), bias.end()),
        d64(dim_t.begin(), dim_t.end()), r64(nr);
    rdetr_oracle_rel_forward_f64(s64.data(), t64.data(), w64.data(), b64.data(), d64.data(), 100.0, 1e-5, nullptr, r64.data(), B, N1, N2, H, 8);
    for (int flags = 0; flags < 2; ++flags) {
        const size_t wsb = rdetr_relation_workspace_bytes(B, N1, N2, flags);
        void *ws = nullptr;
        if (wsb) CK(cudaMalloc(&ws, wsb));
        RK(rdetr_relation_forward(d_src, d_tgt, d_w, d_b, d_dim, 100.f, 1e-5f, nullptr, d_rel, d_bits, B, N1, N2, H, flags, ws, wsb, st));
        CK(cudaStreamSynchronize(st));
        std::vector<float> rel(nr);
        CK(cudaMemcpy(rel.data(), d_rel, nr * 4, cudaMemcpyDeviceToHost));
        double e = 0;
        for (size_t i = 0; i < nr; ++i) e = fmax(e, fabs(rel[i] - r64[i]));
        printf("relation (%s): out %.2e\n", flags ? "FAST" : "EXACT", e);
        if (!(e <= 1e-4)) return 5;
        if (ws) cudaFree(ws);
    }
    // ---- assignment: two problems (one per orientation), duplicated columns so that ties must be broken SciPy's way
    {
        const int64_t rows[2] = {40, 6}, cols[2] = {9, 15};
        std::vector<float> c0(40 * 9), c1(6 * 15);
        for (int i = 0; i < 40; ++i) for (int j = 0; j < 9; ++j) c0[i * 9 + j] = (float)(((i * 7 + (j % 3) * 13) % 11) * 0.25);
        for (size_t i = 0; i < c1.size(); ++i) c1[i] = (float)((i * 2654435761u >> 7) % 5);
        float *d_c0 = to_dev(c0), *d_c1 = to_dev(c1);
        int64_t *d_idx;
        int32_t *d_status;
        CK(cudaMalloc(&d_idx, (9 + 9 + 6 + 6) * sizeof(int64_t)));
        CK(cudaMalloc(&d_status, 2 * sizeof(int32_t)));
        const float *cost[2] = {d_c0, d_c1};
        int64_t *ri[2] = {d_idx, d_idx + 18}, *ci[2] = {d_idx + 9, d_idx + 24};
        const size_t wsb = rdetr_lsap_workspace_bytes(rows, cols, 2);
        void *ws = nullptr;
        if (wsb) CK(cudaMalloc(&ws, wsb));
        RK(rdetr_lsap_solve(cost, rows, cols, ri, ci, d_status, 2, ws, wsb, st));
        CK(cudaStreamSynchronize(st));
        std::vector<int64_t> idx(30);
        int32_t status[2];
        CK(cudaMemcpy(idx.data(), d_idx, idx.size() * sizeof(int64_t), cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(status, d_status, sizeof(status), cudaMemcpyDeviceToHost));
        std::vector<double> c0d(c0.begin(), c0.end()), c1d(c1.begin(), c1.end());
        int64_t r0[9], k0[9], r1[6], k1[6];
        if (rdetr_oracle_lsap(c0d.data(), 40, 9, r0, k0) != 0 || rdetr_oracle_lsap(c1d.data(), 6, 15, r1, k1) != 0) return 8;
        if (status[0] != 0 || status[1] != 0) return 8;
        for (int k = 0; k < 9; ++k) if (idx[k] != r0[k] || idx[9 + k] != k0[k]) { printf("lsap problem 0 differs at %d\n", k); return 8; }
        for (int k = 0; k < 6; ++k) if (idx[18 + k] != r1[k] || idx[24 + k] != k1[k]) { printf("lsap problem 1 differs at %d\n", k); return 8; }
        printf("lsap: 2 problems identical to the oracle\n");
        if (rdetr_lsap_solve(cost, rows, cols, ri, ci, d_status, 2, nullptr, 0, st) != RDETR_ERR_WORKSPACE) return 9;
    }
    // error behaviour: unsupported head dim, host pointer
    if (rdetr_msda_forward(d_value, d_shapes, d_lsi, d_loc, d_attn, d_out, B, S, M, 16, L, Nq, P, RDETR_DTYPE_F32, st) != RDETR_ERR_UNSUPPORTED) return 6;
    if (rdetr_msda_forward(value.data(), d_shapes, d_lsi, d_loc, d_attn, d_out, B, S, M, D, L, Nq, P, RDETR_DTYPE_F32, st) == RDETR_OK) return 7;
    printf("C ABI OK\n");
    return 0;
}
