/*
 * rdetr_oracle.c -- CPU restatement of the two Relation-DETR hot-path operators.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (relation-detr_b200/) may link,
 * import or call this file; it is the checker that tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py compare the CUDA kernels against.
 *
 * Parity status: PINNED.  The reference ships no golden vectors of its own (no tests at all), so
 * this restatement is pinned against outputs of the reference itself, generated in the build
 * container by oracle/make_golden.py (imports /root/reference) and committed under tests/golden/.
 *
 * What is restated (reference file:line, relative to the upstream repository root):
 *   MSDA forward    models/bricks/ops/cuda/ms_deform_im2col_cuda.cuh:22-73 (bilinear corner read),
 *                   :226-288 (loop order l outer / p inner, the -0.5 shift, validity window);
 *                   layout conventions from models/bricks/ms_deform_attn.py:159-212.
 *   MSDA backward   models/bricks/ops/cuda/ms_deform_im2col_cuda.cuh:76-148 (corner gradients),
 *                   :290-392 (per-(b,q,m) accumulation of grad_loc / grad_attn).
 *   REL forward     models/bricks/relation_transformer.py:481-490 (box_rel_encoding), :512-532
 *                   (PositionRelationEmbedding.forward); models/bricks/position_encoding.py:101-105
 *                   (dim_t) and :131-138 (sin/cos interleave, (x*scale)/dim_t evaluation order).
 *   REL backward    autograd of the 1x1 Conv2d + ReLU above; geometry is under no_grad (:527-529).
 *
 * Every routine exists in an f32 and an f64 flavour (same source, REAL switched by macro) so tests
 * can report new-vs-f32-oracle, new-vs-f64-oracle and the f32-vs-f64 noise floor side by side.
 * OpenMP is used only to let bench.py time the port on all host cores; results do not depend on
 * the thread count (each output element is produced by exactly one thread, in a fixed order).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int rdetr_oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

static inline float  floor_f32(float x)  { return floorf(x); }
static inline double floor_f64(double x) { return floor(x); }
static inline float  log_f32(float x)  { return logf(x); }
static inline double log_f64(double x) { return log(x); }
static inline float  sin_f32(float x)  { return sinf(x); }
static inline double sin_f64(double x) { return sin(x); }
static inline float  cos_f32(float x)  { return cosf(x); }
static inline double cos_f64(double x) { return cos(x); }
static inline float  abs_f32(float x)  { return fabsf(x); }
static inline double abs_f64(double x) { return fabs(x); }

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define SFX f32
#include "rdetr_oracle_impl.inc"
#undef REAL
#undef SFX

#define REAL double
#define SFX f64
#include "rdetr_oracle_impl.inc"
#undef REAL
#undef SFX
