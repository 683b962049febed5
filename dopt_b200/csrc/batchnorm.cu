// batchnorm.cu -- batchNormTrain / batchNormGrad / batchNormInference on NCHW fp32 tensors.
//
// Reference (cuda/source/dopt/cuda/nnet/cudnn7.d:547-654): cuDNN CUDNN_BATCHNORM_SPATIAL (the [N,C] case is described to
// cuDNN as [N,C,1,1], cudnn7.d:562-566), eps = 1e-5, exponentialAverageFactor = 1 - momentum computed in double
// (cudnn7.d:592), no saved mean / inv-variance, so backward recomputes the batch statistics from x (cudnn7.d:628-634).
//   train     deps [x, scale(1,C,1,1), bias(C), mean(C), var(C)] -> packed rank-1 [ y (V) | newMean (C) | newVar (C) ]
//             (core/source/dopt/core/ops/nnet.d:222-225,476-489).  y uses the biased batch variance, the running variance
//             the unbiased one: new = old*(1-f) + batch*f.
//   grad      deps [dy, x, scale] -> packed [ dx (V) | dscale (C) | dbias (C) ] (core/ops/nnet.d:232-235)
//             dbias = sum dy, dscale = sum dy*xhat, dx = scale*istd/M * (M*dy - dbias - xhat*dscale)
//   inference deps [x, scale, bias, mean, var] -> y = scale*(x-mean)/sqrt(var+eps) + bias
//
// Kernels: a per-channel statistics reduction (grid = C x splits, partial sums in fp32 around a per-channel pivot so the
// E[x^2]-E[x]^2 cancellation stays harmless, final combine in double) followed by a fully vectorised apply pass.
// HBM-bound.  Algorithmic bytes: train 2V*4 (+ the statistics re-read, which the 126 MB L2 absorbs when the tensor was
// just produced), grad 3V*4, inference 2V*4.
#include "common.cuh"
#include "flat.cuh"
#include <cstdlib>

namespace db {

static constexpr double kBnEps = 1e-5;

struct BnGeom {
    int64_t N, C, HW;
};

// ---- finalize (per channel), run by the LAST statistics block of a channel ----------------------------------------------
// Separate finalize launches were 50 tiny kernels per WRN step.  Each statistics block publishes its partial sums, then
// bumps a per-channel counter; the block that sees splits-1 combines the partials in split order (deterministic) and
// resets the counter for the next execution.
struct BnFin {
    unsigned* counters;                       // [C], zero between launches
    const float* scale; const float* bias;    // train
    const float* rmean; const float* rvar;    // train; may alias mean2 / var2
    float* coef;                              // train [mean | a | b], grad [mean | A | B | C]
    float* out0; float* out1;                 // train: new mean / var (packed tail); grad: dscale / dbias
    float* mean2; float* var2;                // train: optional second copy (the caller's return buffers)
    double factor;
};

__device__ __forceinline__ void bn_train_finalize_channel(const float* part, const float* x, const BnFin& f, const BnGeom& g,
                                                          int splits, int c) {
    double s1 = 0, s2 = 0;
    for (int s = 0; s < splits; ++s) {
        s1 += __ldcg(part + ((int64_t)s * g.C + c) * 2 + 0);
        s2 += __ldcg(part + ((int64_t)s * g.C + c) * 2 + 1);
    }
    double M = (double)g.N * (double)g.HW;
    double K = x[(int64_t)c * g.HW];
    double d = s1 / M;
    double mean = K + d;
    double var = s2 / M - d * d;
    if (var < 0) var = 0;
    double istd = 1.0 / sqrt(var + kBnEps);
    f.coef[c] = (float)mean;
    f.coef[g.C + c] = (float)((double)f.scale[c] * istd);
    f.coef[2 * g.C + c] = f.bias[c];
    double unbiased = M > 1 ? var * M / (M - 1) : var;
    const float nm = (float)((double)f.rmean[c] * (1.0 - f.factor) + mean * f.factor);
    const float nv = (float)((double)f.rvar[c] * (1.0 - f.factor) + unbiased * f.factor);
    f.out0[c] = nm;
    f.out1[c] = nv;
    // second copy for the caller's buffers (may be the very rmean / rvar just read: read-before-write per channel)
    if (f.mean2) f.mean2[c] = nm;
    if (f.var2) f.var2[c] = nv;
}

// grad: dx = dy*A + (x-mean)*B + Cc
__device__ __forceinline__ void bn_grad_finalize_channel(const float* part, const float* x, const BnFin& f, const BnGeom& g,
                                                         int splits, int c) {
    double s1 = 0, s2 = 0, sd = 0, sdx = 0;
    for (int s = 0; s < splits; ++s) {
        const float* p = part + ((int64_t)s * g.C + c) * 4;
        s1 += __ldcg(p + 0); s2 += __ldcg(p + 1); sd += __ldcg(p + 2); sdx += __ldcg(p + 3);
    }
    double M = (double)g.N * (double)g.HW;
    double K = x[(int64_t)c * g.HW];
    double d = s1 / M;
    double mean = K + d;
    double var = s2 / M - d * d;
    if (var < 0) var = 0;
    double istd = 1.0 / sqrt(var + kBnEps);
    double dbeta = sd;
    double dgamma = (sdx - d * sd) * istd;   // sum dy*(x-mean)*istd
    f.out0[c] = (float)dgamma;
    f.out1[c] = (float)dbeta;
    // dx = scale*istd * (dy - dbeta/M - xhat*dgamma/M),  xhat = (x-mean)*istd
    double A = (double)f.scale[c] * istd;
    double B = -A * istd * dgamma / M;
    double Cc = -A * dbeta / M;
    f.coef[c] = (float)mean;
    f.coef[g.C + c] = (float)A;
    f.coef[2 * g.C + c] = (float)B;
    f.coef[3 * g.C + c] = (float)Cc;
}

// ---- statistics ----------------------------------------------------------------------------------------------------
// part layout: [split][C][NS] floats.  NS = 2 (train: sum(x-K), sum((x-K)^2)) or 4 (grad: + sum dy, sum dy*(x-K)).
// MASK (grad only): dy is the gradient w.r.t. relu(y); it is gated by [y > 0] with y recomputed from x and the FORWARD
// pass's per-channel coefficients (fcoef = [mean | a | b], the very values and operation the forward apply used, so the
// gate is bit-identical to testing the stored relu output)
template <bool GRAD, bool MASK = false>
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       float* __restrict__ part, BnGeom g, int splits,
                                                       const __grid_constant__ BnFin fin,
                                                       const float* __restrict__ fcoef = nullptr) {
    constexpr int NS = GRAD ? 4 : 2;
    float fm = 0.f, fa = 0.f, fb = 0.f;
    if (MASK) {
        fm = fcoef[blockIdx.x];
        fa = fcoef[g.C + blockIdx.x];
        fb = fcoef[2 * g.C + blockIdx.x];
    }
    __shared__ float sm[NS][8];
    const int c = blockIdx.x, sp = blockIdx.y;
    const float K = x[(int64_t)c * g.HW];   // pivot: first sample of the channel
    int64_t n_per = (g.N + splits - 1) / splits;
    int64_t n0 = (int64_t)sp * n_per, n1 = n0 + n_per < g.N ? n0 + n_per : g.N;
    float s1 = 0.f, s2 = 0.f, sd = 0.f, sdx = 0.f;
    if ((g.HW & 3) == 0 && (((uintptr_t)x | (GRAD ? (uintptr_t)dy : (uintptr_t)0)) & 15) == 0) {
        const int64_t hw4 = g.HW >> 2;
        const int64_t total = (n1 - n0) * hw4;
        // four independent 128-bit loads (per operand) in flight per thread: the one-load-per-trip version ran at half the
        // HBM bandwidth (profiles/r01_final_timeline.txt)
        constexpr int U = 4;
        auto accumulate = [&](const float4& v, const float4& q0) {
            float4 q = q0;
            float a = v.x - K, b = v.y - K, cc = v.z - K, d = v.w - K;
            s1 += (a + b) + (cc + d);
            s2 += (a * a + b * b) + (cc * cc + d * d);
            if (GRAD) {
                if (MASK) {
                    q.x = fmaf(v.x - fm, fa, fb) > 0.f ? q.x : 0.f;
                    q.y = fmaf(v.y - fm, fa, fb) > 0.f ? q.y : 0.f;
                    q.z = fmaf(v.z - fm, fa, fb) > 0.f ? q.z : 0.f;
                    q.w = fmaf(v.w - fm, fa, fb) > 0.f ? q.w : 0.f;
                }
                sd += (q.x + q.y) + (q.z + q.w);
                sdx += (q.x * a + q.y * b) + (q.z * cc + q.w * d);
            }
        };
        // index arithmetic in 32 bits (a 64-bit division per load made this memory-bound kernel ALU-heavy); the host keeps
        // (images per split) * HW/4 below 2^31
        const uint32_t hw4u = (uint32_t)hw4, totalu = total > 0 ? (uint32_t)total : 0u, bd = blockDim.x;   // empty tail splits
        const int64_t img_stride = g.C * g.HW;
        const float* xb = x + ((int64_t)n0 * g.C + c) * g.HW;
        const float* db = GRAD ? dy + ((int64_t)n0 * g.C + c) * g.HW : nullptr;
        uint32_t i = threadIdx.x;
        for (; i + (U - 1) * bd < totalu; i += U * bd) {
            float4 v[U], q[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const uint32_t iu = i + u * bd;
                const uint32_t n = iu / hw4u, j = iu - n * hw4u;
                const int64_t off = (int64_t)n * img_stride + (j << 2);
                v[u] = *(const float4*)(xb + off);
                q[u] = GRAD ? *(const float4*)(db + off) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) accumulate(v[u], q[u]);
        }
        for (; i < totalu; i += bd) {
            const uint32_t n = i / hw4u, j = i - n * hw4u;
            const int64_t off = ((((int64_t)n0 + n) * g.C + c) * g.HW) + (j << 2);
            const float4 v = *(const float4*)(x + off);
            const float4 q = GRAD ? *(const float4*)(dy + off) : make_float4(0.f, 0.f, 0.f, 0.f);
            accumulate(v, q);
        }
    } else {
        const int64_t total = (n1 - n0) * g.HW;
        for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
            int64_t n = n0 + i / g.HW, j = i % g.HW;
            int64_t off = (n * g.C + c) * g.HW + j;
            float a = x[off] - K;
            s1 += a;
            s2 += a * a;
            if (GRAD) {
                float q = dy[off];
                if (MASK) q = fmaf(x[off] - fm, fa, fb) > 0.f ? q : 0.f;
                sd += q;
                sdx += q * a;
            }
        }
    }
    float vals[4] = {s1, s2, sd, sdx};
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NS; ++k) {
        float v = dbk::warp_sum(vals[k]);
        if (lane == 0) sm[k][w] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 0; k < NS; ++k) {
            float v = 0.f;
            for (int i = 0; i < 8; ++i) v += sm[k][i];
            part[((int64_t)sp * g.C + c) * NS + k] = v;
        }
        __threadfence();   // the partials are visible before the counter moves
        const unsigned seen = atomicAdd(&fin.counters[c], 1u);
        if (seen == (unsigned)splits - 1u) {
            fin.counters[c] = 0u;
            __threadfence();
            if (GRAD) bn_grad_finalize_channel(part, x, fin, g, splits, c);
            else bn_train_finalize_channel(part, x, fin, g, splits, c);
        }
    }
}

__global__ void bn_infer_coef(const float* __restrict__ scale, const float* __restrict__ bias,
                              const float* __restrict__ mean, const float* __restrict__ var, float* __restrict__ coef,
                              int C) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double istd = 1.0 / sqrt((double)var[c] + kBnEps);
    coef[c] = mean[c];
    coef[C + c] = (float)((double)scale[c] * istd);
    coef[2 * C + c] = bias[c];
}

// ---- apply: y = (x-mean[c])*a[c] + b[c]  (MODE 0)   /   dx = dy*A[c] + (x-mean[c])*B[c] + Cc[c]  (MODE 1) ---------------------------------
// RELU: MODE 0 -> relu on the result; MODE 1 -> dy is gated by the forward relu, recomputed from x and fcoef (see bn_stats_kernel)
template <int MODE, bool RELU>
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                       const float* __restrict__ coef, float* __restrict__ out,
                                                       BnGeom g, const float* __restrict__ fcoef = nullptr,
                                                       const float* __restrict__ addend = nullptr) {
    const int64_t V = g.N * g.C * g.HW;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    if ((g.HW & 3) == 0 && (((uintptr_t)out | (uintptr_t)x | (uintptr_t)addend | (MODE == 1 ? (uintptr_t)dy : (uintptr_t)0)) & 15) == 0) {
        const int64_t nv = V >> 2, hw4 = g.HW >> 2;
        const bool small = nv < (1ll << 31);   // 32-bit index arithmetic: a 64-bit division per vector is most of this loop's ALU work
        const uint32_t hw4u = (uint32_t)hw4, Cu = (uint32_t)g.C;
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
            int c;
            if (small) {
                const uint32_t row = (uint32_t)i / hw4u;
                c = (int)(row % Cu);
            } else {
                c = (int)((i / hw4) % g.C);
            }
            float4 v = dbk::ld_stream((const float4*)x + i), r;
            const float mu = coef[c];
            bool m0 = true, m1 = true, m2 = true, m3 = true;
            if (MODE == 1 && RELU) {
                const float fm = fcoef[c], fa = fcoef[g.C + c], fb = fcoef[2 * g.C + c];
                m0 = fmaf(v.x - fm, fa, fb) > 0.f; m1 = fmaf(v.y - fm, fa, fb) > 0.f;
                m2 = fmaf(v.z - fm, fa, fb) > 0.f; m3 = fmaf(v.w - fm, fa, fb) > 0.f;
            }
            v.x -= mu; v.y -= mu; v.z -= mu; v.w -= mu;
            if (MODE == 0) {
                float a = coef[g.C + c], b = coef[2 * g.C + c];
                r.x = fmaf(v.x, a, b); r.y = fmaf(v.y, a, b); r.z = fmaf(v.z, a, b); r.w = fmaf(v.w, a, b);
                if (RELU) { r.x = fmaxf(r.x, 0.f); r.y = fmaxf(r.y, 0.f); r.z = fmaxf(r.z, 0.f); r.w = fmaxf(r.w, 0.f); }
            } else {
                float A = coef[g.C + c], B = coef[2 * g.C + c], Cc = coef[3 * g.C + c];
                float4 q = dbk::ld_stream((const float4*)dy + i);
                if (RELU) { q.x = m0 ? q.x : 0.f; q.y = m1 ? q.y : 0.f; q.z = m2 ? q.z : 0.f; q.w = m3 ? q.w : 0.f; }
                r.x = fmaf(q.x, A, fmaf(v.x, B, Cc)); r.y = fmaf(q.y, A, fmaf(v.y, B, Cc));
                r.z = fmaf(q.z, A, fmaf(v.z, B, Cc)); r.w = fmaf(q.w, A, fmaf(v.w, B, Cc));
                if (addend) {   // the `add` node that follows, folded in (same rounding as the stand-alone add)
                    const float4 e = dbk::ld_stream((const float4*)addend + i);
                    r.x = __fadd_rn(r.x, e.x); r.y = __fadd_rn(r.y, e.y); r.z = __fadd_rn(r.z, e.z); r.w = __fadd_rn(r.w, e.w);
                }
            }
            dbk::st_stream((float4*)out + i, r);
        }
    } else {
        for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < V; i += stride) {
            int c = (int)((i / g.HW) % g.C);
            float v = x[i] - coef[c], r;
            if (MODE == 0) {
                r = fmaf(v, coef[g.C + c], coef[2 * g.C + c]);
                if (RELU) r = fmaxf(r, 0.f);
            } else {
                float q = dy[i];
                if (RELU) q = fmaf(x[i] - fcoef[c], fcoef[g.C + c], fcoef[2 * g.C + c]) > 0.f ? q : 0.f;
                r = fmaf(q, coef[g.C + c], fmaf(v, coef[2 * g.C + c], coef[3 * g.C + c]));
                if (addend) r = __fadd_rn(r, addend[i]);
            }
            out[i] = r;
        }
    }
}

// ---- apply + relu + NHWC bf16 staging in one pass -------------------------------------------------------------------------
// The plan compiler folds relu(batchNormTrain(x).y) and the NHWC bf16 staging of the result (the operand format of the
// tensor-core convolutions) into the apply pass: x is read once, and per element 4 B (fp32 result, when anybody reads it)
// + 2 B (bf16 copy) are written, instead of 22 B for apply + relu + staging as three kernels.  The arithmetic is the
// same as bn_apply_kernel / relu_kernel / nchw_to_nhwc_bf16_kernel, so results are bit-identical.
// Tile: 16 channels x PX pixels of one image (PX = 256, or 64 for small maps).  Reads are PX*4-byte contiguous runs per channel
// (128-bit loads; the earlier 64-channel x 32-pixel tile read 128-byte pieces 4 KB apart and ran at half the HBM
// bandwidth), fp32 results go back the same way, the bf16 copy is written as one full 32-byte sector per pixel
// (16 channels) -- neighbouring channel groups fill the rest of the line in L2.
// grid (ceil(HW/PX), ceil(Cp/16), N), 256 threads.
template <int MODE, bool RELU, int PX>
__global__ void __launch_bounds__(256) bn_apply_stage_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                             const float* __restrict__ coef, float* __restrict__ out,
                                                             __nv_bfloat16* __restrict__ staged, BnGeom g, int Cp,
                                                             const float* __restrict__ fcoef = nullptr,
                                                             const float* __restrict__ addend = nullptr) {
    constexpr int PITCH = PX + 4;            // floats; keeps 128-bit shared stores aligned
    constexpr int V4 = PX / 4;               // float4 per channel row
    constexpr int PER_CH = V4 / 32 > 0 ? V4 / 32 : 1;   // float4 per lane per channel (PX=256: 2, PX=64: 1 for 16 lanes)
    __shared__ __align__(16) float tile[16][PITCH];
    const int n = blockIdx.z;
    const int hw0 = blockIdx.x * PX, c0 = blockIdx.y * 16;
    const int HW = (int)g.HW, C = (int)g.C;
    const int64_t img = (int64_t)n * C * HW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool vec = (HW & 3) == 0 && ((((uintptr_t)x | (uintptr_t)out | (uintptr_t)addend | (MODE == 1 ? (uintptr_t)dy : (uintptr_t)0)) & 15) == 0);
    // warp w handles channels c0 + w and c0 + w + 8
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cl = w + half * 8, c = c0 + cl;
        float cm = 0.f, ca = 0.f, cb = 0.f, cc = 0.f, fm = 0.f, fa = 0.f, fb = 0.f;
        if (c < C) {
            cm = coef[c]; ca = coef[C + c]; cb = coef[2 * C + c];
            if (MODE == 1) cc = coef[3 * C + c];
            if (MODE == 1 && RELU) { fm = fcoef[c]; fa = fcoef[C + c]; fb = fcoef[2 * C + c]; }
        }
        auto apply = [&](float xv, float q) {
            const float v = xv - cm;
            if (MODE == 0) {
                float r = fmaf(v, ca, cb);
                if (RELU) r = (r > 0.f || r != r) ? r : 0.f;
                return r;
            }
            if (RELU) q = fmaf(xv - fm, fa, fb) > 0.f ? q : 0.f;
            return fmaf(q, ca, fmaf(v, cb, cc));
        };
        if (vec) {
            float4 xv[PER_CH], qv[PER_CH], ev[PER_CH];
#pragma unroll
            for (int u = 0; u < PER_CH; ++u) {
                const int p4 = lane + u * 32;                  // float4 index inside the tile row
                const int hw = hw0 + p4 * 4;
                const bool ok = c < C && p4 < V4 && hw < HW;   // HW % 4 == 0: a float4 is inside or outside as a whole
                const int64_t i = img + (int64_t)c * HW + hw;
                xv[u] = ok ? *(const float4*)(x + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                qv[u] = (MODE == 1 && ok) ? *(const float4*)(dy + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                ev[u] = (MODE == 1 && ok && addend) ? *(const float4*)(addend + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < PER_CH; ++u) {
                const int p4 = lane + u * 32;
                const int hw = hw0 + p4 * 4;
                const bool ok = c < C && p4 < V4 && hw < HW;
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (ok) {
                    r.x = apply(xv[u].x, qv[u].x); r.y = apply(xv[u].y, qv[u].y);
                    r.z = apply(xv[u].z, qv[u].z); r.w = apply(xv[u].w, qv[u].w);
                    if (MODE == 1 && addend) {
                        r.x = __fadd_rn(r.x, ev[u].x); r.y = __fadd_rn(r.y, ev[u].y);
                        r.z = __fadd_rn(r.z, ev[u].z); r.w = __fadd_rn(r.w, ev[u].w);
                    }
                    if (out) *(float4*)(out + img + (int64_t)c * HW + hw) = r;
                }
                if (p4 < V4) *(float4*)&tile[cl][p4 * 4] = r;
            }
        } else {
            for (int p = lane; p < PX; p += 32) {
                const int hw = hw0 + p;
                float r = 0.f;
                if (c < C && hw < HW) {
                    const int64_t i = img + (int64_t)c * HW + hw;
                    r = apply(x[i], MODE == 1 ? dy[i] : 0.f);
                    if (MODE == 1 && addend) r = __fadd_rn(r, addend[i]);
                    if (out) out[i] = r;
                }
                tile[cl][p] = r;
            }
        }
    }
    if (!staged) return;
    __syncthreads();
    // one 16-byte store (8 channels) per thread trip: lanes 2p, 2p+1 write the two halves of pixel p's 32-byte sector
    __nv_bfloat16* dst = staged + (int64_t)n * HW * Cp;
    for (int idx = threadIdx.x; idx < PX * 2; idx += 256) {
        const int p = idx >> 1, h = idx & 1;
        const int hw = hw0 + p, c = c0 + h * 8;
        if (hw < HW && c < Cp) {   // Cp is a multiple of 8
            __nv_bfloat162 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __floats2bfloat162_rn(tile[h * 8 + 2 * k][p], tile[h * 8 + 2 * k + 1][p]);
            *(uint4*)(dst + (int64_t)hw * Cp + c) = *(uint4*)v;
        }
    }
}

namespace {

// relu: forward -> relu on the result; backward -> gate dy by the forward relu (fcoef).  fp32 == nullptr: no fp32 result.
template <int MODE>
static void bn_apply_tiled(const float* x, const float* dy, const float* coef, float* fp32, void* staged, bool relu,
                           const float* fcoef, const BnGeom& g, cudaStream_t s, const float* addend = nullptr) {
    const int Cp = (int)((g.C + 7) / 8 * 8);
    if (g.HW > 64) {
        dim3 grid((unsigned)ceil_div(g.HW, 256), (unsigned)ceil_div(Cp, 16), (unsigned)g.N);
        if (relu) bn_apply_stage_kernel<MODE, true, 256><<<grid, 256, 0, s>>>(x, dy, coef, fp32, (__nv_bfloat16*)staged, g, Cp, fcoef, addend);
        else bn_apply_stage_kernel<MODE, false, 256><<<grid, 256, 0, s>>>(x, dy, coef, fp32, (__nv_bfloat16*)staged, g, Cp, fcoef, addend);
    } else {
        dim3 grid((unsigned)ceil_div(g.HW, 64), (unsigned)ceil_div(Cp, 16), (unsigned)g.N);
        if (relu) bn_apply_stage_kernel<MODE, true, 64><<<grid, 256, 0, s>>>(x, dy, coef, fp32, (__nv_bfloat16*)staged, g, Cp, fcoef, addend);
        else bn_apply_stage_kernel<MODE, false, 64><<<grid, 256, 0, s>>>(x, dy, coef, fp32, (__nv_bfloat16*)staged, g, Cp, fcoef, addend);
    }
    DB_LAUNCH_CHECK();
}

static BnGeom geom_of(const dopt_b200_tensor& x) {
    // shape padded with ones to rank 4 (cudnn7.d:562-566)
    DB_REQUIRE(x.rank >= 2 && x.rank <= 4, "batchNorm: input rank must be 2..4");
    BnGeom g;
    g.N = x.shape[0];
    g.C = x.shape[1];
    g.HW = 1;
    for (int i = 2; i < x.rank; ++i) g.HW *= x.shape[i];
    return g;
}

// grid = C x splits CTAs of 256 threads (40 registers: up to 6 resident per SM).  Experiment knobs, read per kernel
// construction: DOPT_B200_BN_CTAS_PER_SM (default 4) sets the target number of CTAs per SM, DOPT_B200_BN_SPLITS forces the
// split count (tools/bn_sweep.sh sweeps them; 640 CTAs on 148 SMs leave 48 SMs with one CTA more than the rest).
static int pick_splits(const BnGeom& g) {
    int per_sm = 4;
    if (const char* e = getenv("DOPT_B200_BN_CTAS_PER_SM")) per_sm = std::max(1, std::min(16, atoi(e)));
    if (const char* e = getenv("DOPT_B200_BN_SPLITS")) return (int)std::max<int64_t>(1, std::min<int64_t>(atoi(e), g.N));
    int64_t want = ceil_div(per_sm * (int64_t)sm_count(), g.C);
    int64_t s = std::min<int64_t>(want, g.N);
    // keep at least ~2048 elements per CTA
    int64_t per = g.N * g.HW;
    while (s > 1 && per / s < 2048) --s;
    return (int)std::max<int64_t>(1, s);
}

struct BnTrainKernel : Kernel {
    BnGeom g;
    double factor;
    int splits;
    Scratch ws;
    Absorb ab;
    float* coef_dev = nullptr;   // [mean | a | b] of the last run: the backward pass recomputes the relu gate from it
    bool can_absorb() const override { return g.HW < (1ll << 30) && g.C < (1 << 24) && g.N < 65536; }
    void set_absorbed(const Absorb& a) override { ab = a; }
    const void* aux_ptr() const override { return coef_dev; }
    float* mean_out2 = nullptr;
    float* var_out2 = nullptr;
    const void* counters_for = nullptr;
    // bf16-interior mode (flat.cu): x and the result are NHWC bf16
    bool flat = false;
    const void* staged_x = nullptr;
    Scratch fws;
    const void* fws_for = nullptr;
    int stats_source = 0;   // 1 / 2: the producer of x accumulates the statistics (flat.cuh)
    bool can_flat() const override { return flat_supported(g.N, g.C, g.HW); }
    void set_flat(bool on) override { flat = on; }
    void set_staged_input(int input, const void* p) override {
        if (input == 0) staged_x = p;
    }
    void* flat_workspace(cudaStream_t s, bool sync) {
        const size_t wb = flat_bn_workspace_bytes((int)g.C);
        void* w = fws.get(wb);
        if (w != fws_for) {
            if (sync) DB_CUDA(cudaMemset(w, 0, wb));
            else DB_CUDA(cudaMemsetAsync(w, 0, wb, s));
            fws_for = w;
        }
        return w;
    }
    void* stats_workspace(int mode) override {
        if (!flat || (mode != 1 && mode != 2)) return nullptr;
        stats_source = mode;
        return flat_workspace(nullptr, true);
    }
    bool set_stat_outputs(float* m, float* v) override {
        mean_out2 = m;
        var_out2 = v;
        return true;
    }
    BnTrainKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 5, "batchNormTrain: deps are [x, scale, bias, mean, var]");
        g = geom_of(d.inputs[0]);
        for (int i = 1; i < 5; ++i) DB_REQUIRE(volume(d.inputs[i]) == g.C, "batchNormTrain: per-channel operand size");
        DB_REQUIRE(volume(d.output) == g.N * g.C * g.HW + 2 * g.C, "batchNormTrain: packed output size");
        factor = 1.0 - d.momentum;   // cudnn7.d:592
        splits = pick_splits(g);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 5, "batchNormTrain: five inputs");
        const float* x = (const float*)in[0];
        int64_t V = g.N * g.C * g.HW;
        float* y = (float*)out;
        if (flat) {
            DB_REQUIRE(staged_x && ab.staged && ab.skip_fp32, "batchNormTrain: flat mode needs staged input and output");
            void* w = flat_workspace(s, false);
            FlatBnTrain a{};
            a.stats_source = stats_source;
            a.x = staged_x; a.y = ab.staged;
            a.scale = (const float*)in[1]; a.bias = (const float*)in[2];
            a.rmean = (const float*)in[3]; a.rvar = (const float*)in[4];
            a.new_mean = y + V; a.new_var = y + V + g.C;
            a.mean2 = mean_out2; a.var2 = var_out2;
            a.factor = factor;
            a.relu = ab.relu;
            a.workspace = w;
            FlatGeom fg{g.N * g.HW, (int)g.C, (int)((g.C + 7) / 8 * 8)};
            coef_dev = const_cast<float*>(flat_bn_train(a, fg, s));
            return;
        }
        float* part = (float*)ws.get(((size_t)splits * g.C * 2 + 4 * g.C) * sizeof(float));
        float* coef = part + (size_t)splits * g.C * 2;
        unsigned* counters = (unsigned*)(coef + 3 * g.C);
        if (ws.ptr != counters_for) {   // new workspace: counters start at zero and return to zero after every launch
            DB_CUDA(cudaMemsetAsync(counters, 0, (size_t)g.C * sizeof(unsigned), s));
            counters_for = ws.ptr;
        }
        BnFin fin{};
        fin.counters = counters;
        fin.scale = (const float*)in[1]; fin.bias = (const float*)in[2];
        fin.rmean = (const float*)in[3]; fin.rvar = (const float*)in[4];
        fin.coef = coef;
        fin.out0 = y + V; fin.out1 = y + V + g.C;
        fin.mean2 = mean_out2; fin.var2 = var_out2;
        fin.factor = factor;
        bn_stats_kernel<false><<<dim3((unsigned)g.C, (unsigned)splits), 256, 0, s>>>(x, nullptr, part, g, splits, fin);
        DB_LAUNCH_CHECK();
        coef_dev = coef;
        if (ab.relu || ab.staged || ab.skip_fp32 || ab.redirect) {
            float* fp32 = ab.skip_fp32 ? nullptr : (ab.redirect ? ab.redirect : y);
            bn_apply_tiled<0>(x, nullptr, coef, fp32, ab.staged, ab.relu, nullptr, g, s);
            return;
        }
        bn_apply_kernel<0, false><<<stream_grid(ceil_div(V, 4), 256, 8), 256, 0, s>>>(x, nullptr, coef, y, g);
        DB_LAUNCH_CHECK();
    }
};

struct BnGradKernel : Kernel {
    BnGeom g;
    int splits;
    Scratch ws;
    Absorb ab;
    const Kernel* fwd = nullptr;   // batchNormTrain kernel whose relu gates dy (plan pass "absorb")
    const void* counters_for = nullptr;
    // bf16-interior mode (flat.cu): dy, x, the optional addend and dx are NHWC bf16; mean / istd come from the forward pass
    bool flat = false;
    const void* staged_in[4] = {nullptr, nullptr, nullptr, nullptr};   // 0 = dy, 1 = x, 3 = addend
    Scratch fws;
    const void* fws_for = nullptr;
    bool producer_stats = false;   // the convolution that writes dy accumulates the statistics in its epilogue
    bool can_flat() const override { return flat_supported(g.N, g.C, g.HW); }
    void set_flat(bool on) override { flat = on; }
    void set_staged_input(int input, const void* p) override {
        if (input >= 0 && input < 4) staged_in[input] = p;
    }
    void* flat_workspace(cudaStream_t s, bool sync) {
        const size_t wb = flat_bn_workspace_bytes((int)g.C);
        void* w = fws.get(wb);
        if (w != fws_for) {
            if (sync) DB_CUDA(cudaMemset(w, 0, wb));
            else DB_CUDA(cudaMemsetAsync(w, 0, wb, s));
            fws_for = w;
        }
        return w;
    }
    void* stats_workspace(int mode) override {
        if (!flat || mode != 3 || !fwd) return nullptr;
        producer_stats = true;
        return flat_workspace(nullptr, true);
    }
    bool can_absorb() const override { return g.HW < (1ll << 30) && g.C < (1 << 24) && g.N < 65536; }
    void set_absorbed(const Absorb& a) override {
        DB_REQUIRE(!a.relu, "batchNormGrad cannot absorb a relu");
        DB_REQUIRE(flat || (a.redirect != nullptr) == (a.addend != nullptr), "batchNormGrad: redirect and addend come together");
        ab = a;
    }
    void set_gate_source(const Kernel* forward) override { fwd = forward; }
    BnGradKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 3, "batchNormGrad: deps are [parentGrad, x, scale]");
        g = geom_of(d.inputs[1]);
        DB_REQUIRE(volume(d.inputs[0]) == g.N * g.C * g.HW, "batchNormGrad: parentGrad volume");
        DB_REQUIRE(volume(d.inputs[2]) == g.C, "batchNormGrad: scale size");
        // judge: volume = vol(dy) + vol(x) + vol(scale), of which only V + 2C is written (core/ops/nnet.d:232-235, F4)
        DB_REQUIRE(volume(d.output) >= g.N * g.C * g.HW + 2 * g.C, "batchNormGrad: packed output size");
        splits = pick_splits(g);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 3, "batchNormGrad: three inputs");
        const float* dy = (const float*)in[0];
        const float* x = (const float*)in[1];
        int64_t V = g.N * g.C * g.HW;
        float* dx = (float*)out;
        if (flat) {
            const float* fc = fwd ? (const float*)fwd->aux_ptr() : nullptr;
            DB_REQUIRE(fc, "batchNormGrad: flat mode needs the forward pass's coefficients");
            DB_REQUIRE(staged_in[0] && staged_in[1] && ab.staged && ab.skip_fp32, "batchNormGrad: flat mode needs staged operands and output");
            void* w = flat_workspace(s, false);
            FlatBnGrad a{};
            a.stats_from_producer = producer_stats;
            a.dy = staged_in[0]; a.x = staged_in[1]; a.addend = staged_in[3];
            a.dx = ab.staged;
            a.scale = (const float*)in[2];
            a.fcoef = fc;
            a.gate = true;
            a.dscale = dx + V; a.dbias = dx + V + g.C;
            a.workspace = w;
            FlatGeom fg{g.N * g.HW, (int)g.C, (int)((g.C + 7) / 8 * 8)};
            flat_bn_grad(a, fg, s);
            return;
        }
        float* part = (float*)ws.get(((size_t)splits * g.C * 4 + 5 * g.C) * sizeof(float));
        float* coef = part + (size_t)splits * g.C * 4;
        unsigned* counters = (unsigned*)(coef + 4 * g.C);
        if (ws.ptr != counters_for) {
            DB_CUDA(cudaMemsetAsync(counters, 0, (size_t)g.C * sizeof(unsigned), s));
            counters_for = ws.ptr;
        }
        BnFin fin{};
        fin.counters = counters;
        fin.scale = (const float*)in[2];
        fin.coef = coef;
        fin.out0 = dx + V; fin.out1 = dx + V + g.C;
        const float* fcoef = fwd ? (const float*)fwd->aux_ptr() : nullptr;
        DB_REQUIRE(!fwd || fcoef, "batchNormGrad: the forward pass it takes its relu gate from has not run");
        if (fcoef) bn_stats_kernel<true, true><<<dim3((unsigned)g.C, (unsigned)splits), 256, 0, s>>>(x, dy, part, g, splits, fin, fcoef);
        else bn_stats_kernel<true><<<dim3((unsigned)g.C, (unsigned)splits), 256, 0, s>>>(x, dy, part, g, splits, fin);
        DB_LAUNCH_CHECK();
        float* dst = ab.redirect ? ab.redirect : dx;   // with an absorbed add: the add node's buffer receives dx + addend
        if (ab.staged || ab.skip_fp32) {
            bn_apply_tiled<1>(x, dy, coef, ab.skip_fp32 ? nullptr : dst, ab.staged, fcoef != nullptr, fcoef, g, s, ab.addend);
            return;
        }
        if (fcoef) bn_apply_kernel<1, true><<<stream_grid(ceil_div(V, 4), 256, 8), 256, 0, s>>>(x, dy, coef, dst, g, fcoef, ab.addend);
        else bn_apply_kernel<1, false><<<stream_grid(ceil_div(V, 4), 256, 8), 256, 0, s>>>(x, dy, coef, dst, g, nullptr, ab.addend);
        DB_LAUNCH_CHECK();
    }
};

struct BnInferKernel : Kernel {
    BnGeom g;
    Scratch ws;
    Absorb ab;   // test-time plans: the relu that follows and the NHWC bf16 copy the next convolution reads (plan pass "absorb")
    bool can_absorb() const override { return g.HW < (1ll << 30) && g.C < (1 << 24) && g.N < 65536; }
    void set_absorbed(const Absorb& a) override { ab = a; }
    BnInferKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 5, "batchNormInference: deps are [x, scale, bias, mean, var]");
        g = geom_of(d.inputs[0]);
        for (int i = 1; i < 5; ++i) DB_REQUIRE(volume(d.inputs[i]) == g.C, "batchNormInference: per-channel operand size");
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 5, "batchNormInference: five inputs");
        float* coef = (float*)ws.get((size_t)3 * g.C * sizeof(float));
        bn_infer_coef<<<(unsigned)ceil_div(g.C, 128), 128, 0, s>>>((const float*)in[1], (const float*)in[2],
                                                                   (const float*)in[3], (const float*)in[4], coef, (int)g.C);
        DB_LAUNCH_CHECK();
        int64_t V = g.N * g.C * g.HW;
        if (ab.relu || ab.staged || ab.skip_fp32 || ab.redirect) {
            // same apply pass as the training kernel's: relu(y) into the relu node's buffer (or nowhere), staged copy beside it
            float* fp32 = ab.skip_fp32 ? nullptr : (ab.redirect ? ab.redirect : (float*)out);
            bn_apply_tiled<0>((const float*)in[0], nullptr, coef, fp32, ab.staged, ab.relu, nullptr, g, s);
            return;
        }
        bn_apply_kernel<0, false><<<stream_grid(ceil_div(V, 4), 256, 8), 256, 0, s>>>((const float*)in[0], nullptr, coef,
                                                                                      (float*)out, g);
        DB_LAUNCH_CHECK();
    }
};

template <class K> Kernel* make(const dopt_b200_op& d) { return new K(d); }
}  // namespace

void register_batchnorm() {
    register_kernel("batchNormTrain", make<BnTrainKernel>);
    register_kernel("batchNormGrad", make<BnGradKernel>);
    register_kernel("batchNormInference", make<BnInferKernel>);
}

}  // namespace db
