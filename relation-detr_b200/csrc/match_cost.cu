// match_cost.cu -- the matcher's cost matrix for a batch of (image, prediction set) problems in one launch
// (SURVEY.md §8 row N3).
//
// Replaces HungarianMatcher.calculate_cost (models/matcher/hungarian_matcher.py:40-72 of the reference):
// ~30 small eager kernels per problem (sigmoid, pow, log, two gathers, cdist, box conversion, GIoU, the
// weighted sum), 28 problems per training step at batch 2.  The solver consumes float32 costs and the
// assignment must be the reference's, so this kernel is written to return the SAME BITS as that chain of
// eager kernels, not merely close values: every operation below is the IEEE single-precision operation the
// corresponding eager kernel performs, in the same order, with contraction into FMAs ruled out by the
// __f*_rn intrinsics.  Specifically
//   sigmoid      1 / (1 + exp(-x))                          (aten sigmoid kernel, opmath float)
//   focal terms  ((p*p) * -(1-alpha)) * log((1-p) + 1e-6)   and the mirrored positive term; gamma == 2 is a
//                square as in aten's pow, any other gamma goes through powf
//   L1 cdist     (|d0| + |d2|) + (|d1| + |d3|)               -- the order of aten's shuffle-down reduction
//                over the 4 coordinates held by lanes 0..3
//   GIoU         torchvision.ops.generalized_box_iou on cxcywh->xyxy boxes, operation for operation
//   total        (w_bbox*bbox + w_class*class) + w_giou*giou (hungarian_matcher.py:71)
// tests/test_lsap_gpu.py checks torch.equal against the eager chain on the GPU box.
#include "common.cuh"

namespace rdetr {

constexpr int kCostTasksPerLaunch = 64;
constexpr int kCostThreads = 256;

struct CostTask {
    const float *pred_boxes;     // [Nq, 4] cxcywh
    const float *pred_logits;    // [Nq, C]
    const float *gt_boxes;       // [G, 4] cxcywh
    const long long *gt_labels;  // [G]
    float *cost;                 // [Nq, G]
    int n_queries, n_gt;
};
struct CostLaunch { CostTask t[kCostTasksPerLaunch]; };

struct CostParams {
    int num_classes;
    float w_class, w_bbox, w_giou;
    float neg_scale, pos_scale;   // -(1 - alpha), -alpha  (rounded to float from the double, as a Python scalar is)
    float gamma;
};

// aten's pow(tensor, scalar) special-cases a few exponents; the reference's gamma is 2
__device__ __forceinline__ float focal_pow(float x, float gamma)
{
    if (gamma == 2.0f) return __fmul_rn(x, x);
    if (gamma == 1.0f) return x;
    if (gamma == 3.0f) return __fmul_rn(__fmul_rn(x, x), x);
    if (gamma == 0.5f) return sqrtf(x);
    return powf(x, gamma);
}

__device__ __forceinline__ void to_xyxy(const float4 b, float &x0, float &y0, float &x1, float &y1)
{
    const float hw = __fmul_rn(0.5f, b.z), hh = __fmul_rn(0.5f, b.w);
    x0 = __fsub_rn(b.x, hw); y0 = __fsub_rn(b.y, hh);
    x1 = __fadd_rn(b.x, hw); y1 = __fadd_rn(b.y, hh);
}

__global__ void __launch_bounds__(kCostThreads)
match_cost_kernel(const __grid_constant__ CostLaunch launch, const CostParams prm)
{
    const CostTask &task = launch.t[blockIdx.y];
    const int G = task.n_gt;
    const unsigned total = (unsigned)task.n_queries * (unsigned)G;
    for (unsigned e = blockIdx.x * kCostThreads + threadIdx.x; e < total; e += gridDim.x * kCostThreads) {
        const unsigned q = e / (unsigned)G, g = e - q * (unsigned)G;
        const float4 pb = __ldg(reinterpret_cast<const float4 *>(task.pred_boxes) + q);
        const float4 gb = __ldg(reinterpret_cast<const float4 *>(task.gt_boxes) + g);
        const long long label = __ldg(task.gt_labels + g);

        // classification: focal-style cost of the target's class
        // a label outside [0, num_classes) would be an out-of-bounds read (the reference's advanced indexing raises a
        // device-side assert): poison the cost instead, so that rdetr_lsap_solve reports status 2 for the problem
        if (label < 0 || label >= prm.num_classes) {
            task.cost[e] = __int_as_float(0x7fc00000);
            continue;
        }
        const float x = __ldg(task.pred_logits + (size_t)q * prm.num_classes + label);
        const float p = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
        const float omp = __fsub_rn(1.0f, p);
        const float neg = __fmul_rn(__fmul_rn(prm.neg_scale, focal_pow(p, prm.gamma)), logf(__fadd_rn(omp, 1e-6f)));
        const float pos = __fmul_rn(__fmul_rn(prm.pos_scale, focal_pow(omp, prm.gamma)), logf(__fadd_rn(p, 1e-6f)));
        const float c_class = __fsub_rn(pos, neg);

        // L1 distance between the cxcywh boxes
        const float d0 = fabsf(__fsub_rn(pb.x, gb.x)), d1 = fabsf(__fsub_rn(pb.y, gb.y));
        const float d2 = fabsf(__fsub_rn(pb.z, gb.z)), d3 = fabsf(__fsub_rn(pb.w, gb.w));
        const float c_bbox = __fadd_rn(__fadd_rn(d0, d2), __fadd_rn(d1, d3));

        // generalized IoU
        float ax0, ay0, ax1, ay1, bx0, by0, bx1, by1;
        to_xyxy(pb, ax0, ay0, ax1, ay1);
        to_xyxy(gb, bx0, by0, bx1, by1);
        const float area_a = __fmul_rn(__fsub_rn(ax1, ax0), __fsub_rn(ay1, ay0));
        const float area_b = __fmul_rn(__fsub_rn(bx1, bx0), __fsub_rn(by1, by0));
        const float iw = fmaxf(__fsub_rn(fminf(ax1, bx1), fmaxf(ax0, bx0)), 0.0f);
        const float ih = fmaxf(__fsub_rn(fminf(ay1, by1), fmaxf(ay0, by0)), 0.0f);
        const float inter = __fmul_rn(iw, ih);
        const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
        const float iou = __fdiv_rn(inter, uni);
        const float ew = fmaxf(__fsub_rn(fmaxf(ax1, bx1), fminf(ax0, bx0)), 0.0f);
        const float eh = fmaxf(__fsub_rn(fmaxf(ay1, by1), fminf(ay0, by0)), 0.0f);
        const float enclose = __fmul_rn(ew, eh);
        const float giou = __fsub_rn(iou, __fdiv_rn(__fsub_rn(enclose, uni), enclose));
        const float c_giou = -giou;

        task.cost[e] = __fadd_rn(__fadd_rn(__fmul_rn(prm.w_bbox, c_bbox), __fmul_rn(prm.w_class, c_class)),
                                 __fmul_rn(prm.w_giou, c_giou));
    }
}

}  // namespace rdetr

extern "C" int rdetr_match_cost(const float *const *pred_boxes, const float *const *pred_logits, const float *const *gt_boxes,
                                const int64_t *const *gt_labels, float *const *cost, const int64_t *n_queries,
                                const int64_t *n_gt, int num_classes, float w_class, float w_bbox, float w_giou,
                                double focal_alpha, double focal_gamma, int n_problems, rdetr_stream_t stream)
{
    using namespace rdetr;
    if (n_problems < 0 || num_classes <= 0) return fail(RDETR_ERR_INVALID_ARGUMENT, "match_cost: n_problems=%d num_classes=%d", n_problems, num_classes);
    if (n_problems == 0) return RDETR_OK;
    if (!pred_boxes || !pred_logits || !gt_boxes || !gt_labels || !cost || !n_queries || !n_gt)
        return fail(RDETR_ERR_INVALID_ARGUMENT, "match_cost: null argument");
    const void *anchor = nullptr;
    int64_t max_elems = 0;
    for (int p = 0; p < n_problems; ++p) {
        if (n_queries[p] < 0 || n_gt[p] < 0 || n_queries[p] * n_gt[p] >= (int64_t)1 << 31)
            return fail(RDETR_ERR_INVALID_ARGUMENT, "match_cost: problem %d is %lld x %lld", p, (long long)n_queries[p], (long long)n_gt[p]);
        if (n_queries[p] * n_gt[p] == 0) continue;
        if (!pred_boxes[p] || !pred_logits[p] || !gt_boxes[p] || !gt_labels[p] || !cost[p])
            return fail(RDETR_ERR_INVALID_ARGUMENT, "match_cost: null buffer in problem %d", p);
        if ((reinterpret_cast<uintptr_t>(pred_boxes[p]) | reinterpret_cast<uintptr_t>(gt_boxes[p])) & 15)
            return fail(RDETR_ERR_INVALID_ARGUMENT, "match_cost: box buffers of problem %d are not 16-byte aligned", p);
        if (!anchor) anchor = cost[p];
        if (n_queries[p] * n_gt[p] > max_elems) max_elems = n_queries[p] * n_gt[p];
    }
    if (!anchor) return RDETR_OK;   // nothing to compute (no ground truth anywhere)
    const DeviceGuard guard(anchor);
    if (guard.status() != RDETR_OK) return guard.status();

    CostParams prm;
    prm.num_classes = num_classes;
    prm.w_class = w_class; prm.w_bbox = w_bbox; prm.w_giou = w_giou;
    prm.neg_scale = (float)(-(1.0 - focal_alpha));
    prm.pos_scale = (float)(-focal_alpha);
    prm.gamma = (float)focal_gamma;
    const unsigned blocks_x = (unsigned)((max_elems + kCostThreads - 1) / kCostThreads);
    for (int first = 0; first < n_problems; first += kCostTasksPerLaunch) {
        CostLaunch launch;
        const int count = n_problems - first < kCostTasksPerLaunch ? n_problems - first : kCostTasksPerLaunch;
        for (int k = 0; k < count; ++k) {
            const int p = first + k;
            CostTask &t = launch.t[k];
            t.pred_boxes = pred_boxes[p]; t.pred_logits = pred_logits[p]; t.gt_boxes = gt_boxes[p];
            t.gt_labels = reinterpret_cast<const long long *>(gt_labels[p]); t.cost = cost[p];
            t.n_queries = (int)n_queries[p]; t.n_gt = (int)n_gt[p];
        }
        match_cost_kernel<<<dim3(blocks_x < 1 ? 1 : blocks_x, count), kCostThreads, 0, static_cast<cudaStream_t>(stream)>>>(launch, prm);
        const int rc = check_cuda(cudaGetLastError(), "match_cost_kernel launch");
        if (rc != RDETR_OK) return rc;
    }
    return RDETR_OK;
}
