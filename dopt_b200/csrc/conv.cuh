// conv.cuh -- geometry shared by the convolution kernels.
#pragma once
#include "common.cuh"

namespace db {

// x: [N,C,H,W]   w: [K,C,R,S]   y: [N,K,P,Q]   stride (u,v)   padding (ph,pw)
struct ConvGeom {
    int N, C, H, W, K, R, S, P, Q, u, v, ph, pw;
};

void conv_fwd_simt_launch(const float* x, const float* w, float* y, const ConvGeom& g, cudaStream_t s);
void conv_dgrad_simt_launch(const float* dy, const float* w, float* dx, const ConvGeom& g, cudaStream_t s);
void conv_wgrad_simt_launch(const float* dy, const float* x, float* dw, const ConvGeom& g, cudaStream_t s);

}  // namespace db
