// conv_simt.cu -- fp32 direct convolution kernels (MATH_FP32 path and the fallback for shapes the tcgen05 path does not
// take: C < 16, odd spatial sizes, 5x5 MNIST filters ...).
//
// Semantics are the reference's: cuDNN CUDNN_CONVOLUTION == TRUE convolution, filters flipped in both spatial axes
// (cuda/source/dopt/cuda/nnet/cudnn7.d:87; pinned by the known-answer test core/source/dopt/core/ops/nnet.d:270-293),
// NCHW fp32 activations, KCRS fp32 filters, dilation 1:
//   fwd    y[n,k,p,q]  = sum_{c,r,s} x[n,c,p*u-ph+r, q*v-pw+s] * w[k,c,R-1-r,S-1-s]              cudnn7.d:113-159
//   dgrad  dx[n,c,h,w] = sum_{k,r,s : h=p*u-ph+r, w=q*v-pw+s} dy[n,k,p,q] * w[k,c,R-1-r,S-1-s]      cudnn7.d:161-204
//   wgrad  dw[k,c,R-1-r,S-1-s] = sum_{n,p,q} dy[n,k,p,q] * x[n,c,p*u-ph+r, q*v-pw+s]               cudnn7.d:206-249
// fp32 multiply-accumulate throughout.  These are correctness-first kernels; the tensor-core path is conv_tc.cu.
#include "common.cuh"
#include "conv.cuh"

namespace db {

// thread = one output pixel x KT output channels
template <int KT>
__global__ void __launch_bounds__(128) conv_fwd_simt(const float* __restrict__ x, const float* __restrict__ w,
                                                     float* __restrict__ y, ConvGeom g) {
    int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t npix = (int64_t)g.N * g.P * g.Q;
    if (pix >= npix) return;
    int k0 = blockIdx.y * KT;
    int q = (int)(pix % g.Q);
    int64_t t = pix / g.Q;
    int p = (int)(t % g.P);
    int n = (int)(t / g.P);
    float acc[KT];
#pragma unroll
    for (int i = 0; i < KT; ++i) acc[i] = 0.f;
    const int h0 = p * g.u - g.ph, w0 = q * g.v - g.pw;
    for (int c = 0; c < g.C; ++c) {
        const float* xc = x + ((int64_t)n * g.C + c) * g.H * g.W;
        for (int r = 0; r < g.R; ++r) {
            int h = h0 + r;
            if (h < 0 || h >= g.H) continue;
            for (int s = 0; s < g.S; ++s) {
                int ww = w0 + s;
                if (ww < 0 || ww >= g.W) continue;
                float xv = xc[h * g.W + ww];
                int widx = (g.R - 1 - r) * g.S + (g.S - 1 - s);
#pragma unroll
                for (int i = 0; i < KT; ++i) {
                    int k = k0 + i;
                    if (k < g.K) acc[i] = fmaf(xv, __ldg(w + ((int64_t)k * g.C + c) * g.R * g.S + widx), acc[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < KT; ++i) {
        int k = k0 + i;
        if (k < g.K) y[(((int64_t)n * g.K + k) * g.P + p) * g.Q + q] = acc[i];
    }
}

// thread = one input pixel x CT input channels
template <int CT>
__global__ void __launch_bounds__(128) conv_dgrad_simt(const float* __restrict__ dy, const float* __restrict__ w,
                                                       float* __restrict__ dx, ConvGeom g) {
    int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t npix = (int64_t)g.N * g.H * g.W;
    if (pix >= npix) return;
    int c0 = blockIdx.y * CT;
    int wq = (int)(pix % g.W);
    int64_t t = pix / g.W;
    int h = (int)(t % g.H);
    int n = (int)(t / g.H);
    float acc[CT];
#pragma unroll
    for (int i = 0; i < CT; ++i) acc[i] = 0.f;
    for (int r = 0; r < g.R; ++r) {
        int ph_ = h + g.ph - r;
        if (ph_ < 0 || ph_ % g.u) continue;
        int p = ph_ / g.u;
        if (p >= g.P) continue;
        for (int s = 0; s < g.S; ++s) {
            int pw_ = wq + g.pw - s;
            if (pw_ < 0 || pw_ % g.v) continue;
            int q = pw_ / g.v;
            if (q >= g.Q) continue;
            int widx = (g.R - 1 - r) * g.S + (g.S - 1 - s);
            for (int k = 0; k < g.K; ++k) {
                float gy = dy[(((int64_t)n * g.K + k) * g.P + p) * g.Q + q];
#pragma unroll
                for (int i = 0; i < CT; ++i) {
                    int c = c0 + i;
                    if (c < g.C) acc[i] = fmaf(gy, __ldg(w + ((int64_t)k * g.C + c) * g.R * g.S + widx), acc[i]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CT; ++i) {
        int c = c0 + i;
        if (c < g.C) dx[(((int64_t)n * g.C + c) * g.H + h) * g.W + wq] = acc[i];
    }
}

// CTA = one (k, c) filter plane and one slice of the batch; every thread keeps all R*S taps; atomics merge the slices.
template <int MAXRS>
__global__ void __launch_bounds__(256) conv_wgrad_simt(const float* __restrict__ dy, const float* __restrict__ x,
                                                       float* __restrict__ dw, ConvGeom g, int nsplit) {
    __shared__ float sm[8][MAXRS];
    int kc = blockIdx.x;
    int k = kc / g.C, c = kc % g.C;
    int sp = blockIdx.y;
    int n_per = (g.N + nsplit - 1) / nsplit;
    int n0 = sp * n_per, n1 = min(n0 + n_per, g.N);
    float acc[MAXRS];
#pragma unroll
    for (int i = 0; i < MAXRS; ++i) acc[i] = 0.f;
    const int PQ = g.P * g.Q;
    const int RS = g.R * g.S;
    int64_t total = (int64_t)(n1 - n0) * PQ;
    for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
        int n = n0 + (int)(i / PQ);
        int pq = (int)(i % PQ);
        int p = pq / g.Q, q = pq % g.Q;
        float gy = dy[((int64_t)n * g.K + k) * PQ + pq];
        const float* xc = x + ((int64_t)n * g.C + c) * g.H * g.W;
        int h0 = p * g.u - g.ph, w0 = q * g.v - g.pw;
#pragma unroll
        for (int t = 0; t < MAXRS; ++t) {
            if (t < RS) {
                int r = t / g.S, s = t % g.S;
                int h = h0 + r, ww = w0 + s;
                if (h >= 0 && h < g.H && ww >= 0 && ww < g.W) acc[t] = fmaf(gy, xc[h * g.W + ww], acc[t]);
            }
        }
    }
    int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
    for (int t = 0; t < MAXRS; ++t) {
        float v = dbk::warp_sum(acc[t]);
        if (lane == 0) sm[wp][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < RS) {
        int t = threadIdx.x;
        float v = 0.f;
        for (int i = 0; i < 8; ++i) v += sm[i][t];
        int r = t / g.S, s = t % g.S;
        float* dst = dw + ((int64_t)k * g.C + c) * RS + (g.R - 1 - r) * g.S + (g.S - 1 - s);
        if (nsplit == 1) *dst = v;
        else atomicAdd(dst, v);
    }
}

void conv_fwd_simt_launch(const float* x, const float* w, float* y, const ConvGeom& g, cudaStream_t s) {
    int64_t npix = (int64_t)g.N * g.P * g.Q;
    if (npix == 0 || g.K == 0) return;
    dim3 grid((unsigned)ceil_div(npix, 128), (unsigned)ceil_div(g.K, 8));
    conv_fwd_simt<8><<<grid, 128, 0, s>>>(x, w, y, g);
    DB_LAUNCH_CHECK();
}
void conv_dgrad_simt_launch(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t s) {
    int64_t npix = (int64_t)g.N * g.H * g.W;
    if (npix == 0 || g.C == 0) return;
    dim3 grid((unsigned)ceil_div(npix, 128), (unsigned)ceil_div(g.C, 8));
    conv_dgrad_simt<8><<<grid, 128, 0, s>>>(dy, w, dx, g);
    DB_LAUNCH_CHECK();
}
void conv_wgrad_simt_launch(const float* dy, const float* x, float* dw, const ConvGeom& g, cudaStream_t s) {
    int RS = g.R * g.S;
    DB_REQUIRE(RS <= 49, "convolutionFiltersGrad (fp32 path): filters larger than 7x7 are not supported");
    int64_t planes = (int64_t)g.K * g.C;
    if (planes == 0) return;
    int nsplit = 1;
    if (planes < 4 * sm_count()) nsplit = (int)std::min<int64_t>(g.N, ceil_div(4 * sm_count(), planes));
    if (nsplit > 1) {
        DB_CUDA(cudaMemsetAsync(dw, 0, (size_t)planes * RS * sizeof(float), s));
        count_launch();
    }
    dim3 grid((unsigned)planes, (unsigned)nsplit);
    if (RS <= 9) conv_wgrad_simt<9><<<grid, 256, 0, s>>>(dy, x, dw, g, nsplit);
    else if (RS <= 25) conv_wgrad_simt<25><<<grid, 256, 0, s>>>(dy, x, dw, g, nsplit);
    else conv_wgrad_simt<49><<<grid, 256, 0, s>>>(dy, x, dw, g, nsplit);
    DB_LAUNCH_CHECK();
}

}  // namespace db
