// flat.cuh -- kernels over NHWC bf16 ("staged") activations: the bf16-interior path of the plan (plan.cu, pass "residency").
//
// Inside a plan compiled with DOPT_B200_PLAN_BF16_INTERIOR the activations that only tensor-core convolutions, batch norms
// and residual adds touch never exist as NCHW fp32: the convolution epilogue writes [N][H][W][Cp] bf16 (Cp = C rounded up to
// 8), batchNormTrain / batchNormGrad / add read and write that layout directly.  Per element these passes move 2 bytes per
// operand instead of 4 + the staging copy.  All arithmetic is fp32 in registers; only the stored
// tensors are bf16.
#pragma once
#include "common.cuh"

namespace db {

struct FlatGeom {
    int64_t P;   // pixels = N * H * W
    int C, Cp;   // channels, channels rounded up to 8
};

bool flat_supported(int64_t N, int64_t C, int64_t HW);
// workspace of one batch-norm kernel object: [acc: 2*C doubles | counter (16 B) | coef: 4*C floats]
size_t flat_bn_workspace_bytes(int C);

struct FlatBnTrain {
    const void* x;             // NHWC bf16
    void* y;                   // NHWC bf16: relu?(scale * (x - mean) * istd + bias)
    const float* scale; const float* bias;
    const float* rmean; const float* rvar;   // running statistics (read)
    float* new_mean; float* new_var;         // packed tail of the op's result
    float* mean2; float* var2;               // optional second copy (the caller's return buffers), may alias rmean / rvar
    double factor;                           // 1 - momentum (cudnn7.d:592)
    bool relu;
    int stats_source;                        // 0: own statistics kernel; the producer of x accumulated them already, pivoted by
                                             // pixel 0 (1: flat_add_stats) or unpivoted (2: the convolution epilogue)
    void* workspace;                         // flat_bn_workspace_bytes(C), zeroed once by the caller
};
// returns the coefficient block [mean | a | b | istd] (4*C floats inside the workspace) the backward pass reads
const float* flat_bn_train(const FlatBnTrain& a, const FlatGeom& g, cudaStream_t s);

struct FlatBnGrad {
    const void* dy; const void* x;   // NHWC bf16
    const void* addend;              // optional NHWC bf16: the result is dx + addend
    void* dx;                        // NHWC bf16
    const float* scale;
    const float* fcoef;              // [mean | a | b | istd] of the forward pass (flat_bn_train)
    bool gate;                       // dy is the gradient w.r.t. relu(y): gate it by [y > 0], y recomputed from x and fcoef
    float* dscale; float* dbias;     // packed tail of the op's result
    void* workspace;
    bool stats_from_producer;        // the feature-gradient convolution that wrote dy accumulated sum(g), sum(g * (x - mean))
                                     // in its epilogue (tc_kernel<.., EPI = 3>): no statistics kernel here
};
void flat_bn_grad(const FlatBnGrad& a, const FlatGeom& g, cudaStream_t s);

// where a producer of x accumulates the statistics of a flat batchNormTrain (its workspace): see FlatWs in flat.cu
void flat_stats_sink(void* bn_workspace, int C, unsigned** epoch, float** sums, int* copies);
// out = a + b plus the batch-norm statistics of out into the workspace of the batchNormTrain that reads it
void flat_add_stats(const void* a, const void* b, void* out, const FlatGeom& g, void* bn_workspace, cudaStream_t s);
// out = a + b, all NHWC bf16 of n_elems elements (multiple of 8)
void flat_add(const void* a, const void* b, void* out, int64_t n_elems, cudaStream_t s);
// [N][HW][Cp] bf16 -> NCHW fp32 (the inverse of stage_nchw_to_nhwc_bf16), for the few fp32 readers of a bf16-resident value
void unstage_nhwc_bf16_to_nchw(const void* in, float* out, int N, int C, int64_t HW, cudaStream_t s);

}  // namespace db
