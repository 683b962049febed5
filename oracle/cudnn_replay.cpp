/*
 * cudnn_replay.cpp -- TEST INFRASTRUCTURE ONLY (oracle).  Never linked into, loaded by or called from the product
 * (libdopt_b200.so / libdopt_host.so); only tests/ and tools/ load it.
 *
 * dopt's CUDA backend has no kernels of its own for the nnet ops: it is a sequence of cuDNN / cuBLAS calls
 * (cuda/source/dopt/cuda/nnet/cudnn7.d, cuda/source/dopt/cuda/math.d:214-247, cuda/source/dopt/cuda/basic.d:219-247).
 * The reference itself is D and cannot be built in this image, so this file issues THE SAME descriptors and THE SAME
 * library calls, argument for argument, against the cuDNN 9 / cuBLAS 12 installed here (cuDNN 9 still exports every
 * legacy entry point the reference uses).  What comes out is the arithmetic of "the reference CUDA backend" on this
 * machine, which the north star names as the parity target for convolution / batch-norm / pooling.
 *
 * Each function cites the reference lines it replays.  Device pointers in, device pointers out, NULL stream followed by
 * a device synchronise (the reference does cuCtxSynchronize() after most calls, e.g. cudnn7.d:157).
 *
 * `math`: 0 = leave the convolution descriptor's math type at its default, exactly like the reference (which never
 *             calls cudnnSetConvolutionMathType) -- on sm_80+ cuDNN may then run fp32 convolutions on TF32 tensor cores;
 *         1 = CUDNN_FMA_MATH: strict fp32 FMA arithmetic (used to check the product's MATH_FP32 path tightly).
 */
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cudnn.h>

#include <cstdio>
#include <cstring>
#include <string>

namespace {

thread_local std::string g_err;
cudnnHandle_t g_dnn = nullptr;
cublasHandle_t g_blas = nullptr;

int fail(const char* what, const char* detail, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (cudnn_replay.cpp:%d)", what, detail, line);
    g_err = buf;
    return 1;
}

#define DNN(call)                                                                           \
    do {                                                                                    \
        cudnnStatus_t s_ = (call);                                                          \
        if (s_ != CUDNN_STATUS_SUCCESS) return fail(#call, cudnnGetErrorString(s_), __LINE__); \
    } while (0)
#define RT(call)                                                                            \
    do {                                                                                    \
        cudaError_t s_ = (call);                                                            \
        if (s_ != cudaSuccess) return fail(#call, cudaGetErrorString(s_), __LINE__);        \
    } while (0)
#define BLAS(call)                                                                          \
    do {                                                                                    \
        cublasStatus_t s_ = (call);                                                         \
        if (s_ != CUBLAS_STATUS_SUCCESS) return fail(#call, "cublas error", __LINE__);      \
    } while (0)

int ensure() {
    if (!g_dnn) DNN(cudnnCreate(&g_dnn));           /* cudnn7.d:24-30 (initialize) */
    if (!g_blas) BLAS(cublasCreate(&g_blas));       /* math.d:69-77 */
    return 0;
}

/* ConvolutionBase, cudnn7.d:55-109 */
struct ConvDescs {
    cudnnTensorDescriptor_t x = nullptr, y = nullptr;
    cudnnFilterDescriptor_t w = nullptr;
    cudnnConvolutionDescriptor_t conv = nullptr;
    ~ConvDescs() {
        if (w) cudnnDestroyFilterDescriptor(w);
        if (y) cudnnDestroyTensorDescriptor(y);
        if (conv) cudnnDestroyConvolutionDescriptor(conv);
        if (x) cudnnDestroyTensorDescriptor(x);
    }
    int init(const int* in, const int* filt, const int* out, const int* pad, const int* stride, int math) {
        DNN(cudnnCreateTensorDescriptor(&x));
        DNN(cudnnCreateFilterDescriptor(&w));
        DNN(cudnnCreateConvolutionDescriptor(&conv));
        DNN(cudnnCreateTensorDescriptor(&y));
        DNN(cudnnSetTensor4dDescriptor(x, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, in[0], in[1], in[2], in[3]));
        DNN(cudnnSetFilter4dDescriptor(w, CUDNN_DATA_FLOAT, CUDNN_TENSOR_NCHW, filt[0], filt[1], filt[2], filt[3]));
        /* dilation fixed at 1, CUDNN_CONVOLUTION = true convolution (flipped filters), cudnn7.d:75-77,87 */
        DNN(cudnnSetConvolution2dDescriptor(conv, pad[0], pad[1], stride[0], stride[1], 1, 1, CUDNN_CONVOLUTION,
                                            CUDNN_DATA_FLOAT));
        DNN(cudnnSetTensor4dDescriptor(y, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, out[0], out[1], out[2], out[3]));
        if (math == 1) DNN(cudnnSetConvolutionMathType(conv, CUDNN_FMA_MATH));
        return 0;
    }
};

struct Workspace {
    void* p = nullptr;
    size_t n = 0;
    ~Workspace() { if (p) cudaFree(p); }
    int reserve(size_t bytes) {
        n = bytes;
        if (bytes) RT(cudaMalloc(&p, bytes));
        return 0;
    }
};

/* 4-d NCHW descriptor over [s0, s1, prod(rest), 1] as Softmax / ReLU / AddBias build it (cudnn7.d:339-349,381-383) */
int flat_desc(cudnnTensorDescriptor_t* d, int n, int c, int vol) {
    DNN(cudnnCreateTensorDescriptor(d));
    DNN(cudnnSetTensor4dDescriptor(*d, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, n, c, vol, 1));
    return 0;
}

}  // namespace

extern "C" {

const char* cudnn_replay_last_error() { return g_err.c_str(); }

/* returns cudnnGetVersion(); *cudart = runtime version */
long cudnn_replay_versions(int* cudart) {
    if (cudart) cudaRuntimeGetVersion(cudart);
    return (long)cudnnGetVersion();
}

/* kind 0: ConvolutionForward          cudnn7.d:113-159   y  = conv(x, w)
 * kind 1: ConvolutionFeaturesGrad     cudnn7.d:161-204   dx = f(dy, w)       (x is the output, y holds dy)
 * kind 2: ConvolutionFiltersGrad      cudnn7.d:206-249   dw = f(x, dy)       (w is the output, y holds dy)
 * The algorithm is the fastest one cudnnFind*Algorithm reports, as in the reference; *algo_out receives its number. */
int cudnn_replay_conv(int kind, const int* in_shape, const int* filt_shape, const int* out_shape, const int* pad,
                      const int* stride, float* x, float* w, float* y, int math, int* algo_out) {
    if (ensure()) return 1;
    ConvDescs d;
    if (d.init(in_shape, filt_shape, out_shape, pad, stride, math)) return 1;
    Workspace ws;
    const float alpha = 1.f, beta = 0.f;
    int n = 0;
    if (kind == 0) {
        cudnnConvolutionFwdAlgoPerf_t perf[9];
        DNN(cudnnFindConvolutionForwardAlgorithm(g_dnn, d.x, d.w, d.conv, d.y, 9, &n, perf));
        if (n < 1 || perf[0].status != CUDNN_STATUS_SUCCESS) return fail("find fwd", "no algorithm", __LINE__);
        if (ws.reserve(perf[0].memory)) return 1;
        if (algo_out) *algo_out = (int)perf[0].algo;
        DNN(cudnnConvolutionForward(g_dnn, &alpha, d.x, x, d.w, w, d.conv, perf[0].algo, ws.p, ws.n, &beta, d.y, y));
    } else if (kind == 1) {
        cudnnConvolutionBwdDataAlgoPerf_t perf[9];
        DNN(cudnnFindConvolutionBackwardDataAlgorithm(g_dnn, d.w, d.y, d.conv, d.x, 9, &n, perf));
        if (n < 1 || perf[0].status != CUDNN_STATUS_SUCCESS) return fail("find dgrad", "no algorithm", __LINE__);
        if (ws.reserve(perf[0].memory)) return 1;
        if (algo_out) *algo_out = (int)perf[0].algo;
        DNN(cudnnConvolutionBackwardData(g_dnn, &alpha, d.w, w, d.y, y, d.conv, perf[0].algo, ws.p, ws.n, &beta, d.x, x));
    } else if (kind == 2) {
        cudnnConvolutionBwdFilterAlgoPerf_t perf[9];
        DNN(cudnnFindConvolutionBackwardFilterAlgorithm(g_dnn, d.x, d.y, d.conv, d.w, 9, &n, perf));
        if (n < 1 || perf[0].status != CUDNN_STATUS_SUCCESS) return fail("find wgrad", "no algorithm", __LINE__);
        if (ws.reserve(perf[0].memory)) return 1;
        if (algo_out) *algo_out = (int)perf[0].algo;
        DNN(cudnnConvolutionBackwardFilter(g_dnn, &alpha, d.x, x, d.y, y, d.conv, perf[0].algo, ws.p, ws.n, &beta, d.w, w));
    } else {
        return fail("cudnn_replay_conv", "bad kind", __LINE__);
    }
    RT(cudaDeviceSynchronize());
    return 0;
}

/* MaxpoolForward / MaxpoolGrad, cudnn7.d:251-333: CUDNN_POOLING_MAX, nanOpt 1 (= CUDNN_PROPAGATE_NAN), window = stride =
 * dims, no padding.  backward != 0: dx = PoolingBackward(y, dy, x). */
int cudnn_replay_maxpool(int backward, const int* in_shape, const int* out_shape, const int* dims, float* x, float* y,
                         float* dy, float* dx) {
    if (ensure()) return 1;
    cudnnPoolingDescriptor_t pd;
    cudnnTensorDescriptor_t xd, yd;
    DNN(cudnnCreatePoolingDescriptor(&pd));
    DNN(cudnnSetPooling2dDescriptor(pd, CUDNN_POOLING_MAX, (cudnnNanPropagation_t)1, dims[0], dims[1], 0, 0, dims[0],
                                    dims[1]));
    DNN(cudnnCreateTensorDescriptor(&xd));
    DNN(cudnnCreateTensorDescriptor(&yd));
    DNN(cudnnSetTensor4dDescriptor(xd, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, in_shape[0], in_shape[1], in_shape[2],
                                   in_shape[3]));
    DNN(cudnnSetTensor4dDescriptor(yd, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, out_shape[0], out_shape[1], out_shape[2],
                                   out_shape[3]));
    const float alpha = 1.f, beta = 0.f;
    if (!backward)
        DNN(cudnnPoolingForward(g_dnn, pd, &alpha, xd, x, &beta, yd, y));
    else
        DNN(cudnnPoolingBackward(g_dnn, pd, &alpha, yd, y, yd, dy, xd, x, &beta, xd, dx));
    RT(cudaDeviceSynchronize());
    cudnnDestroyPoolingDescriptor(pd);
    cudnnDestroyTensorDescriptor(xd);
    cudnnDestroyTensorDescriptor(yd);
    return 0;
}

/* Softmax / SoftmaxGrad, cudnn7.d:335-404: ACCURATE, MODE_CHANNEL over [n, c, vol, 1].
 * backward: dx = SoftmaxBackward(y, dy) */
int cudnn_replay_softmax(int backward, int n, int c, int vol, float* a, float* dy, float* out) {
    if (ensure()) return 1;
    cudnnTensorDescriptor_t d;
    if (flat_desc(&d, n, c, vol)) return 1;
    const float alpha = 1.f, beta = 0.f;
    if (!backward)
        DNN(cudnnSoftmaxForward(g_dnn, CUDNN_SOFTMAX_ACCURATE, CUDNN_SOFTMAX_MODE_CHANNEL, &alpha, d, a, &beta, d, out));
    else
        DNN(cudnnSoftmaxBackward(g_dnn, CUDNN_SOFTMAX_ACCURATE, CUDNN_SOFTMAX_MODE_CHANNEL, &alpha, d, a, d, dy, &beta, d,
                                 out));
    RT(cudaDeviceSynchronize());
    cudnnDestroyTensorDescriptor(d);
    return 0;
}

/* ReLU / ReLUGrad, cudnn7.d:406-478: CUDNN_ACTIVATION_RELU, CUDNN_PROPAGATE_NAN, coef 0.
 * forward: out = relu(x).  backward: out = ActivationBackward(y, dy, x). */
int cudnn_replay_relu(int backward, int n, int c, int vol, float* x, float* y, float* dy, float* out) {
    if (ensure()) return 1;
    cudnnTensorDescriptor_t d;
    cudnnActivationDescriptor_t act;
    if (flat_desc(&d, n, c, vol)) return 1;
    DNN(cudnnCreateActivationDescriptor(&act));
    DNN(cudnnSetActivationDescriptor(act, CUDNN_ACTIVATION_RELU, CUDNN_PROPAGATE_NAN, 0.0));
    const float alpha = 1.f, beta = 0.f;
    if (!backward)
        DNN(cudnnActivationForward(g_dnn, act, &alpha, d, x, &beta, d, out));
    else
        DNN(cudnnActivationBackward(g_dnn, act, &alpha, d, y, d, dy, d, x, &beta, d, out));
    RT(cudaDeviceSynchronize());
    cudnnDestroyActivationDescriptor(act);
    cudnnDestroyTensorDescriptor(d);
    return 0;
}

/* AddBias, cudnn7.d:480-512: copy x to the output, then AddTensor(alpha 1, [1,c,1,1] bias, beta 1, output) */
int cudnn_replay_add_bias(int n, int c, int vol, const float* x, const float* bias, float* out) {
    if (ensure()) return 1;
    cudnnTensorDescriptor_t cd, ad;
    if (flat_desc(&cd, n, c, vol)) return 1;
    if (flat_desc(&ad, 1, c, 1)) return 1;
    RT(cudaMemcpy(out, x, sizeof(float) * (size_t)n * c * vol, cudaMemcpyDeviceToDevice));
    const float alpha = 1.f, beta = 1.f;
    DNN(cudnnAddTensor(g_dnn, &alpha, ad, bias, &beta, cd, out));
    RT(cudaDeviceSynchronize());
    cudnnDestroyTensorDescriptor(cd);
    cudnnDestroyTensorDescriptor(ad);
    return 0;
}

/* AddBiasGrad, cudnn7.d:514-545: ConvolutionBackwardBias with alpha 1 and BETA 1 into the op's buffer -- `out` must
 * hold what the plan buffer holds (zeros on the first execution, cuda/source/dopt/cuda/package.d:152). */
int cudnn_replay_add_bias_grad(int n, int c, int vol, const float* dy, float* out) {
    if (ensure()) return 1;
    cudnnTensorDescriptor_t dyd, dbd;
    if (flat_desc(&dyd, n, c, vol)) return 1;
    if (flat_desc(&dbd, 1, c, 1)) return 1;
    const float alpha = 1.f, beta = 1.f;
    DNN(cudnnConvolutionBackwardBias(g_dnn, &alpha, dyd, dy, &beta, dbd, out));
    RT(cudaDeviceSynchronize());
    cudnnDestroyTensorDescriptor(dyd);
    cudnnDestroyTensorDescriptor(dbd);
    return 0;
}

/* BatchNormBase, cudnn7.d:547-585.  `op_rank` is the rank the reference tests (`op.rank == 2` -> PER_ACTIVATION): for
 * batchNormTrain / batchNormGrad that is the rank of the PACKED rank-1 result, so those always run SPATIAL; for
 * batchNormInference it is the rank of x.  shape4 = x's shape padded with ones to 4 dims. */
static int bn_descs(int op_rank, const int* shape4, cudnnBatchNormMode_t* mode, cudnnTensorDescriptor_t* xd,
                    cudnnTensorDescriptor_t* bnd) {
    *mode = op_rank == 2 ? CUDNN_BATCHNORM_PER_ACTIVATION : CUDNN_BATCHNORM_SPATIAL;
    DNN(cudnnCreateTensorDescriptor(xd));
    DNN(cudnnCreateTensorDescriptor(bnd));
    DNN(cudnnSetTensor4dDescriptor(*xd, CUDNN_TENSOR_NCHW, CUDNN_DATA_FLOAT, shape4[0], shape4[1], shape4[2], shape4[3]));
    DNN(cudnnDeriveBNTensorDescriptor(*bnd, *xd, *mode));
    return 0;
}

/* BatchNormTrain, cudnn7.d:587-614: running mean / var are first copied behind y in the packed output, then
 * ForwardTraining updates them in place with factor 1 - momentum, eps 1e-5f, no saved statistics.
 * packed = [y (vol) | mean (c) | var (c)]. */
int cudnn_replay_bn_train(const int* shape4, double momentum, const float* x, const float* scale, const float* bias,
                          const float* mean, const float* var, float* packed) {
    if (ensure()) return 1;
    cudnnBatchNormMode_t mode;
    cudnnTensorDescriptor_t xd, bnd;
    if (bn_descs(1, shape4, &mode, &xd, &bnd)) return 1;
    const size_t vol = (size_t)shape4[0] * shape4[1] * shape4[2] * shape4[3];
    const size_t c = shape4[1];
    float* pm = packed + vol;
    float* pv = pm + c;
    RT(cudaMemcpy(pm, mean, c * sizeof(float), cudaMemcpyDeviceToDevice));
    RT(cudaMemcpy(pv, var, c * sizeof(float), cudaMemcpyDeviceToDevice));
    const float alpha = 1.f, beta = 0.f;
    const double factor = 1.0 - momentum;           /* cudnn7.d:592 */
    DNN(cudnnBatchNormalizationForwardTraining(g_dnn, mode, &alpha, &beta, xd, x, xd, packed, bnd, scale, bias, factor, pm,
                                               pv, 1e-5f, nullptr, nullptr));
    RT(cudaDeviceSynchronize());
    cudnnDestroyTensorDescriptor(xd);
    cudnnDestroyTensorDescriptor(bnd);
    return 0;
}

/* BatchNormGrad, cudnn7.d:616-636: null saved statistics (recomputed from x); packed = [dx (vol) | dscale (c) | dbias (c)] */
int cudnn_replay_bn_grad(const int* shape4, const float* dy, const float* x, const float* scale, float* packed) {
    if (ensure()) return 1;
    cudnnBatchNormMode_t mode;
    cudnnTensorDescriptor_t xd, bnd;
    if (bn_descs(1, shape4, &mode, &xd, &bnd)) return 1;
    const size_t vol = (size_t)shape4[0] * shape4[1] * shape4[2] * shape4[3];
    const size_t c = shape4[1];
    const float alpha = 1.f, beta = 0.f;
    DNN(cudnnBatchNormalizationBackward(g_dnn, mode, &alpha, &beta, &alpha, &beta, xd, x, xd, dy, xd, packed, bnd, scale,
                                        packed + vol, packed + vol + c, 1e-5f, nullptr, nullptr));
    RT(cudaDeviceSynchronize());
    cudnnDestroyTensorDescriptor(xd);
    cudnnDestroyTensorDescriptor(bnd);
    return 0;
}

/* BatchNormInference, cudnn7.d:638-654 (x_rank = rank of x: 2 selects PER_ACTIVATION) */
int cudnn_replay_bn_inference(int x_rank, const int* shape4, const float* x, const float* scale, const float* bias,
                              const float* mean, const float* var, float* y) {
    if (ensure()) return 1;
    cudnnBatchNormMode_t mode;
    cudnnTensorDescriptor_t xd, bnd;
    if (bn_descs(x_rank, shape4, &mode, &xd, &bnd)) return 1;
    const float alpha = 1.f, beta = 0.f;
    DNN(cudnnBatchNormalizationForwardInference(g_dnn, mode, &alpha, &beta, xd, x, xd, y, bnd, scale, bias, mean, var,
                                                1e-5));
    RT(cudaDeviceSynchronize());
    cudnnDestroyTensorDescriptor(xd);
    cudnnDestroyTensorDescriptor(bnd);
    return 0;
}

/* MatmulKernel, math.d:214-247: row-major C[M,N] = A[M,K] B[K,N] as column-major sgemm with swapped operands */
int cudnn_replay_matmul(int M, int K, int N, const float* a, const float* b, float* c) {
    if (ensure()) return 1;
    const float alpha = 1.f, beta = 0.f;
    BLAS(cublasSgemm(g_blas, CUBLAS_OP_N, CUBLAS_OP_N, N, M, K, &alpha, b, N, a, K, &beta, c, N));
    RT(cudaDeviceSynchronize());
    return 0;
}

/* Transpose, basic.d:219-247: out[rows_out, cols_out] from a[cols_out, rows_out] through cublasSgeam(T, T) */
int cudnn_replay_transpose(int rows_out, int cols_out, const float* a, float* c) {
    if (ensure()) return 1;
    const float alpha = 1.f, beta = 0.f;
    BLAS(cublasSgeam(g_blas, CUBLAS_OP_T, CUBLAS_OP_T, cols_out, rows_out, &alpha, a, rows_out, &beta, a, rows_out, c,
                     cols_out));
    RT(cudaDeviceSynchronize());
    return 0;
}

}  // extern "C"
