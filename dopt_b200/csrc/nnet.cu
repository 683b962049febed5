// nnet.cu -- relu / reluGrad / addBias / addBiasGrad / maxpool / maxpoolGrad / softmax / softmaxGrad.
//
// Reference: cuDNN calls in cuda/source/dopt/cuda/nnet/cudnn7.d, each followed by cuCtxSynchronize:
//   relu        cudnnActivationForward(RELU, PROPAGATE_NAN)          cudnn7.d:406-437   y = max(x,0), NaN stays NaN
//   reluGrad    cudnnActivationBackward(y, dy, x)                    cudnn7.d:439-478   dx = dy * [x > 0]; deps [dy, y, x]
//   addBias     cuMemcpy + cudnnAddTensor([1,C,1,1] -> [N,C,HW,1])    cudnn7.d:480-512   y = x + b[c]  (two passes there, one here)
//   addBiasGrad cudnnConvolutionBackwardBias                          cudnn7.d:514-545   db[c] = sum_{n,hw} dy
//               (the reference passes beta = 1 into a buffer zeroed only at plan creation, so its result accumulates across
//                executions -- survey F12; this kernel implements the first-execution value, beta = 0)
//   maxpool     cudnnPoolingForward(MAX, PROPAGATE_NAN, window = stride = dims, pad 0)   cudnn7.d:251-307
//   maxpoolGrad cudnnPoolingBackward(y, dy, x)                        cudnn7.d:309-333   deps [dy, y, x]
//   softmax     cudnnSoftmaxForward(ACCURATE, MODE_CHANNEL) on [N,C,vol,1]               cudnn7.d:335-371
//   softmaxGrad cudnnSoftmaxBackward(y, dy): dx = y * (dy - sum_c dy*y)                  cudnn7.d:373-404   deps [dy, y]
// All are HBM-bound; bytes per launch are listed at each kernel.
#include "common.cuh"
#include <cfloat>
#include <cstdlib>

namespace db {

// ---- relu: 2V*4 B -------------------------------------------------------------------------------------------------
__device__ __forceinline__ float relu1(float x) { return (x > 0.f || x != x) ? x : 0.f; }

__global__ void __launch_bounds__(256) relu_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n,
                                                   int64_t nv) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* xv = (const float4*)x;
    float4* yv = (float4*)y;
    for (; i + 3 * stride < nv; i += 4 * stride) {
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = dbk::ld_stream(xv + i + u * stride);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            v[u].x = relu1(v[u].x); v[u].y = relu1(v[u].y); v[u].z = relu1(v[u].z); v[u].w = relu1(v[u].w);
            dbk::st_stream(yv + i + u * stride, v[u]);
        }
    }
    for (; i < nv; i += stride) {
        float4 v = dbk::ld_stream(xv + i);
        v.x = relu1(v.x); v.y = relu1(v.y); v.z = relu1(v.z); v.w = relu1(v.w);
        dbk::st_stream(yv + i, v);
    }
    for (int64_t t = (nv << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride) y[t] = relu1(x[t]);
}

// ---- reluGrad: 3V*4 B (reads dy and x, writes dx; y is not needed) ---------------------------------------------------
__global__ void __launch_bounds__(256) relu_grad_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        float* __restrict__ dx, int64_t n, int64_t nv) {
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const float4* gv = (const float4*)dy;
    const float4* xv = (const float4*)x;
    float4* ov = (float4*)dx;
    for (; i + stride < nv; i += 2 * stride) {
        float4 g0 = dbk::ld_stream(gv + i), g1 = dbk::ld_stream(gv + i + stride);
        float4 a0 = dbk::ld_stream(xv + i), a1 = dbk::ld_stream(xv + i + stride);
        float4 r0, r1;
        r0.x = a0.x > 0.f ? g0.x : 0.f; r0.y = a0.y > 0.f ? g0.y : 0.f; r0.z = a0.z > 0.f ? g0.z : 0.f; r0.w = a0.w > 0.f ? g0.w : 0.f;
        r1.x = a1.x > 0.f ? g1.x : 0.f; r1.y = a1.y > 0.f ? g1.y : 0.f; r1.z = a1.z > 0.f ? g1.z : 0.f; r1.w = a1.w > 0.f ? g1.w : 0.f;
        dbk::st_stream(ov + i, r0);
        dbk::st_stream(ov + i + stride, r1);
    }
    for (; i < nv; i += stride) {
        float4 g = dbk::ld_stream(gv + i), a = dbk::ld_stream(xv + i), r;
        r.x = a.x > 0.f ? g.x : 0.f; r.y = a.y > 0.f ? g.y : 0.f; r.z = a.z > 0.f ? g.z : 0.f; r.w = a.w > 0.f ? g.w : 0.f;
        dbk::st_stream(ov + i, r);
    }
    for (int64_t t = (nv << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
        dx[t] = x[t] > 0.f ? dy[t] : 0.f;
}

// ---- addBias: 2V*4 B.  x viewed as [N*C][HW]; one CTA row-chunk at a time so b[c] is a register ---------------------
__global__ void __launch_bounds__(256) add_bias_kernel(const float* __restrict__ x, const float* __restrict__ b,
                                                       float* __restrict__ y, int64_t rows, int64_t hw, int C) {
    int64_t n = rows * hw;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int c = (int)((i / hw) % C);
        y[i] = __fadd_rn(x[i], b[c]);
    }
}

// ---- addBiasGrad: V*4 B.  db[c] = sum over n, hw.  grid (C, splits); second pass when split ---------------------------
__global__ void __launch_bounds__(256) bias_grad_kernel(const float* __restrict__ dy, float* __restrict__ out, int N,
                                                        int C, int64_t hw, int splits) {
    __shared__ float sm[32];
    int c = blockIdx.x, sp = blockIdx.y;
    int64_t total = (int64_t)N * hw;
    int64_t per = (total + splits - 1) / splits, lo = sp * per, hi = lo + per < total ? lo + per : total;
    float acc = 0.f;
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) {
        int64_t n = i / hw, j = i - n * hw;
        acc += dy[(n * C + c) * hw + j];
    }
    acc = dbk::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0.f;
        v = dbk::warp_sum(v);
        if (threadIdx.x == 0) out[(int64_t)sp * C + c] = v;
    }
}
__global__ void bias_grad_final(const float* __restrict__ part, float* __restrict__ out, int C, int splits) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * C + c];
    out[c] = acc;
}

// ---- maxpool: (V + V/d^2)*4 B.  thread per output element ------------------------------------------------------------
__global__ void __launch_bounds__(256) maxpool_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t maps,
                                                      int H, int W, int OH, int OW, int dh, int dw) {
    int64_t n = maps * OH * OW;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int ow = (int)(i % OW);
        int64_t t = i / OW;
        int oh = (int)(t % OH);
        int64_t m = t / OH;
        const float* p = x + (m * H + (int64_t)oh * dh) * W + (int64_t)ow * dw;
        float best = -FLT_MAX;
        bool nan = false;
        for (int a = 0; a < dh; ++a)
            for (int b = 0; b < dw; ++b) {
                float v = p[a * W + b];
                nan |= (v != v);
                best = v > best ? v : best;
            }
        y[i] = nan ? __int_as_float(0x7fc00000) : best;
    }
}

// ---- maxpoolGrad: (V/d^2 + V + V)*4 B.  thread per INPUT element: dx = dy[window] where x equals the window max --------
// tie_mode 1 (default): only the first (row-major) element equal to the window maximum receives dy -- what cuDNN's
// PoolingBackward does for CUDNN_POOLING_MAX, measured on the GPU box against the replayed reference call
// (tests/test_cudnn_replay_gpu.py::test_maxpool_grad_tie_rule_is_cudnns, profiles/r01p_cudnn_replay_report.jsonl);
// tie_mode 0: every tied element does.  The scan over the window only runs for elements that equal the maximum.
__global__ void __launch_bounds__(256) maxpool_grad_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           const float* __restrict__ x, float* __restrict__ dx,
                                                           int64_t maps, int H, int W, int OH, int OW, int dh, int dw,
                                                           int tie_mode) {
    int64_t n = maps * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int w = (int)(i % W);
        int64_t t = i / W;
        int h = (int)(t % H);
        int64_t m = t / H;
        int oh = h / dh, ow = w / dw;
        float r = 0.f;
        if (oh < OH && ow < OW) {
            int64_t o = (m * OH + oh) * OW + ow;
            float xv = x[i], yv = y[o];
            if (xv == yv) {
                bool take = true;
                if (tie_mode == 1) {
                    const float* p = x + (m * H + (int64_t)oh * dh) * W + (int64_t)ow * dw;
                    int my = (h - oh * dh) * dw + (w - ow * dw);
                    for (int a = 0; a < dh && take; ++a)
                        for (int b = 0; b < dw; ++b) {
                            if (a * dw + b >= my) break;
                            if (p[a * W + b] == yv) { take = false; break; }
                        }
                }
                if (take) r = dy[o];
            }
        }
        dx[i] = r;
    }
}

// ---- softmax over dim 1 of [N, C, vol]: 2V*4 B.  one warp per (n, i) ---------------------------------------------------
__global__ void __launch_bounds__(256) softmax_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t N,
                                                      int C, int64_t vol) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < N * vol; r += nw) {
        int64_t n = r / vol, i = r - n * vol;
        const float* p = x + n * C * vol + i;
        float* q = y + n * C * vol + i;
        float m = -FLT_MAX;
        for (int c = lane; c < C; c += 32) m = fmaxf(m, p[(int64_t)c * vol]);
        m = dbk::warp_max(m);
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += expf(p[(int64_t)c * vol] - m);
        s = dbk::warp_sum(s);
        for (int c = lane; c < C; c += 32) q[(int64_t)c * vol] = __fdiv_rn(expf(p[(int64_t)c * vol] - m), s);
    }
}
__global__ void __launch_bounds__(256) softmax_grad_kernel(const float* __restrict__ dy, const float* __restrict__ y,
                                                           float* __restrict__ dx, int64_t N, int C, int64_t vol) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = warp; r < N * vol; r += nw) {
        int64_t n = r / vol, i = r - n * vol;
        int64_t base = n * C * vol + i;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s = fmaf(dy[base + (int64_t)c * vol], y[base + (int64_t)c * vol], s);
        s = dbk::warp_sum(s);
        for (int c = lane; c < C; c += 32) {
            int64_t o = base + (int64_t)c * vol;
            dx[o] = __fmul_rn(y[o], __fsub_rn(dy[o], s));
        }
    }
}

void relu_launch(const float* x, float* y, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    bool al = (uintptr_t)x % 16 == 0 && (uintptr_t)y % 16 == 0;
    relu_kernel<<<stream_grid(ceil_div(n, al ? 16 : 1), 256, 8), 256, 0, s>>>(x, y, n, al ? (n >> 2) : 0);
    DB_LAUNCH_CHECK();
}
void relu_grad_launch(const float* dy, const float* x, float* dx, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    bool al = (uintptr_t)x % 16 == 0 && (uintptr_t)dy % 16 == 0 && (uintptr_t)dx % 16 == 0;
    relu_grad_kernel<<<stream_grid(ceil_div(n, al ? 8 : 1), 256, 8), 256, 0, s>>>(dy, x, dx, n, al ? (n >> 2) : 0);
    DB_LAUNCH_CHECK();
}

static int g_pool_tie_mode = 1;

namespace {

static void split_ncv(const dopt_b200_tensor& t, int64_t& N, int64_t& C, int64_t& vol) {
    // the reference describes every tensor to cuDNN as [shape0, shape1, prod(shape[2..]), 1] (cudnn7.d:341-349,412-414)
    DB_REQUIRE(t.rank >= 2, "nnet op needs rank >= 2");
    N = t.shape[0];
    C = t.shape[1];
    vol = 1;
    for (int i = 2; i < t.rank; ++i) vol *= t.shape[i];
}

struct ReluKernel : Kernel {
    int64_t n;
    ReluKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 1 && d.output.dtype == DOPT_B200_FLOAT32, "relu: one float32 operand");
        n = volume(d.output);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "relu: one input");
        relu_launch((const float*)in[0], (float*)out, n, s);
    }
};
struct ReluGradKernel : Kernel {
    int64_t n;
    ReluGradKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 3, "reluGrad: deps are [parentGrad, y, x]");   // core/ops/nnet.d:190-193,461-464
        n = volume(d.output);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 3, "reluGrad: three inputs");
        relu_grad_launch((const float*)in[0], (const float*)in[2], (float*)out, n, s);
    }
};
struct AddBiasKernel : Kernel {
    int64_t N, C, vol;
    AddBiasKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 2, "addBias: deps are [input, bias]");
        split_ncv(d.inputs[0], N, C, vol);
        DB_REQUIRE(volume(d.inputs[1]) == C, "addBias: bias length must equal channel count");
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 2, "addBias: two inputs");
        int64_t n = N * C * vol;
        if (n == 0) return;
        add_bias_kernel<<<stream_grid(n, 256, 16), 256, 0, s>>>((const float*)in[0], (const float*)in[1], (float*)out,
                                                                N * C, vol, (int)C);
        DB_LAUNCH_CHECK();
    }
};
struct AddBiasGradKernel : Kernel {
    int64_t N, C, vol;
    Scratch ws;
    AddBiasGradKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 1, "addBiasGrad: one operand");
        split_ncv(d.inputs[0], N, C, vol);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "addBiasGrad: one input");
        int64_t total = N * vol;
        int splits = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(4 * sm_count(), C), ceil_div(total, 4096)));
        if (splits == 1) {
            bias_grad_kernel<<<dim3((unsigned)C, 1), 256, 0, s>>>((const float*)in[0], (float*)out, (int)N, (int)C, vol, 1);
            DB_LAUNCH_CHECK();
        } else {
            float* part = (float*)ws.get((size_t)splits * C * sizeof(float));
            bias_grad_kernel<<<dim3((unsigned)C, (unsigned)splits), 256, 0, s>>>((const float*)in[0], part, (int)N,
                                                                                 (int)C, vol, splits);
            DB_LAUNCH_CHECK();
            bias_grad_final<<<(unsigned)ceil_div(C, 128), 128, 0, s>>>(part, (float*)out, (int)C, splits);
            DB_LAUNCH_CHECK();
        }
    }
};
struct MaxpoolKernel : Kernel {
    int64_t maps;
    int H, W, OH, OW, dh, dw;
    MaxpoolKernel(const dopt_b200_op& d) {
        const auto& x = d.inputs[0];
        DB_REQUIRE(d.n_inputs == 1 && x.rank == 4, "maxpool: one rank-4 operand");   // core/ops/nnet.d:89-95
        dh = (int)d.pool_dims[0];
        dw = (int)d.pool_dims[1];
        DB_REQUIRE(dh > 0 && dw > 0, "maxpool: bad dims");
        maps = x.shape[0] * x.shape[1];
        H = (int)x.shape[2];
        W = (int)x.shape[3];
        OH = H / dh;
        OW = W / dw;   // judgeMaxpool, core/ops/nnet.d:97-107
        DB_REQUIRE(d.output.shape[2] == OH && d.output.shape[3] == OW, "maxpool: bad output shape");
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "maxpool: one input");
        int64_t n = maps * OH * OW;
        if (n == 0) return;
        maxpool_kernel<<<stream_grid(n, 256, 16), 256, 0, s>>>((const float*)in[0], (float*)out, maps, H, W, OH, OW, dh, dw);
        DB_LAUNCH_CHECK();
    }
};
struct MaxpoolGradKernel : Kernel {
    int64_t maps;
    int H, W, OH, OW, dh, dw, tie;
    MaxpoolGradKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 3 && d.inputs[2].rank == 4, "maxpoolGrad: deps are [parentGrad, y, x]");
        // DOPT_B200_POOL_TIES=all|first overrides the rule at kernel construction (diagnostics; the default, first, is
        // the one tests/test_cudnn_replay_gpu.py measures cuDNN's own PoolingBackward to follow)
        tie = g_pool_tie_mode;
        if (const char* e = getenv("DOPT_B200_POOL_TIES")) tie = (e[0] == 'a' || e[0] == '0') ? 0 : 1;
        const auto& x = d.inputs[2];
        dh = (int)d.pool_dims[0];
        dw = (int)d.pool_dims[1];
        DB_REQUIRE(dh > 0 && dw > 0, "maxpoolGrad: bad dims");
        maps = x.shape[0] * x.shape[1];
        H = (int)x.shape[2];
        W = (int)x.shape[3];
        OH = H / dh;
        OW = W / dw;
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 3, "maxpoolGrad: three inputs");
        int64_t n = maps * H * W;
        if (n == 0) return;
        maxpool_grad_kernel<<<stream_grid(n, 256, 16), 256, 0, s>>>((const float*)in[0], (const float*)in[1],
                                                                    (const float*)in[2], (float*)out, maps, H, W, OH,
                                                                    OW, dh, dw, tie);
        DB_LAUNCH_CHECK();
    }
};
struct SoftmaxKernel : Kernel {
    int64_t N, C, vol;
    bool grad;
    SoftmaxKernel(const dopt_b200_op& d, bool g) : grad(g) {
        DB_REQUIRE(d.n_inputs == (g ? 2 : 1), "softmax: wrong operand count");
        split_ncv(d.output, N, C, vol);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == (grad ? 2 : 1), "softmax: wrong input count");
        int64_t rows = N * vol;
        if (rows == 0) return;
        int grid = stream_grid(rows * 32, 256, 8);
        if (grad)
            softmax_grad_kernel<<<grid, 256, 0, s>>>((const float*)in[0], (const float*)in[1], (float*)out, N, (int)C, vol);
        else
            softmax_kernel<<<grid, 256, 0, s>>>((const float*)in[0], (float*)out, N, (int)C, vol);
        DB_LAUNCH_CHECK();
    }
};

template <class K> Kernel* make(const dopt_b200_op& d) { return new K(d); }
Kernel* make_softmax(const dopt_b200_op& d) { return new SoftmaxKernel(d, false); }
Kernel* make_softmax_grad(const dopt_b200_op& d) { return new SoftmaxKernel(d, true); }
}  // namespace

void set_pool_tie_mode(int m) { g_pool_tie_mode = m; }

void register_nnet() {
    register_kernel("relu", make<ReluKernel>);
    register_kernel("reluGrad", make<ReluGradKernel>);
    register_kernel("addBias", make<AddBiasKernel>);
    register_kernel("addBiasGrad", make<AddBiasGradKernel>);
    register_kernel("maxpool", make<MaxpoolKernel>);
    register_kernel("maxpoolGrad", make<MaxpoolGradKernel>);
    register_kernel("softmax", make_softmax);
    register_kernel("softmaxGrad", make_softmax_grad);
}

}  // namespace db
