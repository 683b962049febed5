// flat.cu -- batch norm, residual add and layout conversion over NHWC bf16 activations (see flat.cuh).
//
// Reference arithmetic: cuDNN CUDNN_BATCHNORM_SPATIAL as dopt calls it (cuda/source/dopt/cuda/nnet/cudnn7.d:547-636), relu /
// reluGrad (cudnn7.d:406-478) folded in, `add` (cuda/source/dopt/cuda/math.d:129-207).  Same formulas as batchnorm.cu; what
// differs is the storage type of the activations (bf16 instead of fp32) and therefore the stated tolerance.
//
// Layout: the tensor is a [P][G] array of 16-byte vectors (P pixels, G = Cp/8 channel groups of 8 bf16).  A CTA has R*G
// threads (R = 256/G pixel rows); thread t owns channel group t % G for the whole kernel, so its per-channel coefficients and
// accumulators live in registers and every trip of the CTA reads R*G*16 contiguous bytes.  HBM-bound:
//   train  statistics 2 B/elem (mostly L2 hits right after the convolution that wrote x) + apply 2 + 2 B/elem
//   grad   statistics 4 B/elem + apply 4 (+2 with an addend) + 2 B/elem
//   add    6 B/elem
#include "flat.cuh"
#include <cstdlib>

namespace db {

static constexpr double kFlatEps = 1e-5;   // cudnn7.d:587-636 pass CUDNN_BN_MIN_EPSILON
static constexpr int kFlatU = 4;           // 16-byte loads in flight per operand per thread

namespace {

struct FlatFin {
    double* acc;          // [C][2], zero between launches
    unsigned* counter;    // zero between launches
    float* coef;          // train: [mean | a | b | istd]; grad: [A | B | Cc]
    const float* scale; const float* bias; const float* rmean; const float* rvar;
    float* out0; float* out1;       // train: new mean / new var; grad: dscale / dbias
    float* mean2; float* var2;
    const float* fcoef;             // grad: forward coefficients
    double factor;
};

__device__ __forceinline__ void unpack8(const uint4& v, float* f) {
    f[0] = __uint_as_float(v.x << 16); f[1] = __uint_as_float(v.x & 0xffff0000u);
    f[2] = __uint_as_float(v.y << 16); f[3] = __uint_as_float(v.y & 0xffff0000u);
    f[4] = __uint_as_float(v.z << 16); f[5] = __uint_as_float(v.z & 0xffff0000u);
    f[6] = __uint_as_float(v.w << 16); f[7] = __uint_as_float(v.w & 0xffff0000u);
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *(uint32_t*)&h;
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    uint4 v;
    v.x = pack2(f[0], f[1]); v.y = pack2(f[2], f[3]); v.z = pack2(f[4], f[5]); v.w = pack2(f[6], f[7]);
    return v;
}
__device__ __forceinline__ uint4 ld16(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st16(uint4* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- statistics ------------------------------------------------------------------------------------------------------
// train: per channel sum(x - K), sum((x - K)^2) around the pivot K = x[pixel 0] (keeps E[x^2] - E[x]^2 harmless in fp32)
// grad:  per channel sum(g), sum(g * (x - mean)) with g = dy gated by the forward relu, mean from the forward pass
// Per-thread fp32 partials -> shared-memory reduction over the CTA's pixel rows -> one double atomic per (channel, sum) per CTA;
// the CTA that arrives last turns the sums into the per-channel coefficients and clears the accumulators for the next launch.
template <bool GRAD, bool GATE>
__global__ void __launch_bounds__(256) flat_bn_stats_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                            int64_t P, int G, int C, int R,
                                                            const __grid_constant__ FlatFin fin) {
    __shared__ float red[16][256];
    __shared__ int s_last;
    const int t = threadIdx.x;
    const int cg = t % G;
    const bool active = t < R * G;
    float piv[8], fa[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) piv[j] = fa[j] = fb[j] = 0.f;
    if (!GRAD) {
        unpack8(x[cg], piv);   // pixel 0
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg * 8 + j;
            if (c < C) {
                piv[j] = fin.fcoef[c];   // the batch mean
                if (GATE) { fa[j] = fin.fcoef[C + c]; fb[j] = fin.fcoef[2 * C + c]; }
            }
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int64_t trip = (int64_t)R * G, total = P * G;
    for (int64_t v0 = (int64_t)blockIdx.x * kFlatU * trip; v0 < total; v0 += (int64_t)gridDim.x * kFlatU * trip) {
        uint4 xv[kFlatU], qv[kFlatU];
        bool ok[kFlatU];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            const int64_t i = v0 + u * trip + t;
            ok[u] = active && i < total;
            xv[u] = ok[u] ? ld16(x + i) : make_uint4(0, 0, 0, 0);
            if (GRAD) qv[u] = ok[u] ? ld16(dy + i) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            if (!ok[u]) continue;
            float xf[8], qf[8];
            unpack8(xv[u], xf);
            if (GRAD) unpack8(qv[u], qf);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = xf[j] - piv[j];
                if (!GRAD) {
                    s1[j] += d;
                    s2[j] = fmaf(d, d, s2[j]);
                } else {
                    float q = qf[j];
                    if (GATE) q = fmaf(d, fa[j], fb[j]) > 0.f ? q : 0.f;
                    s1[j] += q;
                    s2[j] = fmaf(q, d, s2[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[j][t] = active ? s1[j] : 0.f;
        red[8 + j][t] = active ? s2[j] : 0.f;
    }
    __syncthreads();
    for (int idx = t; idx < G * 16; idx += blockDim.x) {
        const int g2 = idx % G, k = idx / G;
        float v = 0.f;
        for (int r = 0; r < R; ++r) v += red[k][r * G + g2];
        const int c = g2 * 8 + (k & 7);
        if (c < C) atomicAdd(&fin.acc[2 * c + (k >> 3)], (double)v);
    }
    __threadfence();
    __syncthreads();
    if (t == 0) s_last = atomicAdd(fin.counter, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const double M = (double)P;
    for (int c = t; c < C; c += blockDim.x) {
        const double a0 = __ldcg(&fin.acc[2 * c]), a1 = __ldcg(&fin.acc[2 * c + 1]);
        fin.acc[2 * c] = 0.0;
        fin.acc[2 * c + 1] = 0.0;
        if (!GRAD) {
            // same combine as bn_train_finalize_channel (batchnorm.cu)
            const uint4* x0 = x + c / 8;
            const uint32_t w = ((const uint32_t*)x0)[(c & 7) >> 1];
            const double K = (double)__uint_as_float((c & 1) ? (w & 0xffff0000u) : (w << 16));
            const double d = a0 / M;
            const double mean = K + d;
            double var = a1 / M - d * d;
            if (var < 0) var = 0;
            const double istd = 1.0 / sqrt(var + kFlatEps);
            fin.coef[c] = (float)mean;
            fin.coef[C + c] = (float)((double)fin.scale[c] * istd);
            fin.coef[2 * C + c] = fin.bias[c];
            fin.coef[3 * C + c] = (float)istd;
            const double unbiased = M > 1 ? var * M / (M - 1) : var;
            const float nm = (float)((double)fin.rmean[c] * (1.0 - fin.factor) + mean * fin.factor);
            const float nv = (float)((double)fin.rvar[c] * (1.0 - fin.factor) + unbiased * fin.factor);
            fin.out0[c] = nm;
            fin.out1[c] = nv;
            if (fin.mean2) fin.mean2[c] = nm;
            if (fin.var2) fin.var2[c] = nv;
        } else {
            // same algebra as bn_grad_finalize_channel: dx = dy*A + (x - mean)*B + Cc
            const double istd = (double)fin.fcoef[3 * C + c];
            const double dbeta = a0, dgamma = a1 * istd;
            fin.out0[c] = (float)dgamma;
            fin.out1[c] = (float)dbeta;
            const double A = (double)fin.scale[c] * istd;
            fin.coef[c] = (float)A;
            fin.coef[C + c] = (float)(-A * istd * dgamma / M);
            fin.coef[2 * C + c] = (float)(-A * dbeta / M);
        }
    }
    if (t == 0) *fin.counter = 0u;
}

// ---- apply -------------------------------------------------------------------------------------------------------------
// MODE 0: y = relu?((x - mean) * a + b)                                  coef = [mean | a | b | istd]
// MODE 1: dx = g * A + (x - mean) * B + Cc (+ addend), g = gated dy      coef = [A | B | Cc], fcoef = forward [mean | a | b | istd]
template <int MODE, bool RELU, bool ADDEND>
__global__ void __launch_bounds__(256) flat_bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                            const uint4* __restrict__ addend, uint4* __restrict__ out,
                                                            int64_t P, int G, int C, int R, const float* __restrict__ coef,
                                                            const float* __restrict__ fcoef) {
    const int t = threadIdx.x;
    const int cg = t % G;
    if (t >= R * G) return;
    float mu[8], ca[8], cb[8], gA[8], gB[8], gC[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        mu[j] = ca[j] = cb[j] = gA[j] = gB[j] = gC[j] = 0.f;
        if (c < C) {
            if (MODE == 0) {
                mu[j] = coef[c]; ca[j] = coef[C + c]; cb[j] = coef[2 * C + c];
            } else {
                mu[j] = fcoef[c];
                if (RELU) { ca[j] = fcoef[C + c]; cb[j] = fcoef[2 * C + c]; }
                gA[j] = coef[c]; gB[j] = coef[C + c]; gC[j] = coef[2 * C + c];
            }
        }
    }
    const int64_t trip = (int64_t)R * G, total = P * G;
    for (int64_t v0 = (int64_t)blockIdx.x * kFlatU * trip; v0 < total; v0 += (int64_t)gridDim.x * kFlatU * trip) {
        uint4 xv[kFlatU], qv[kFlatU], ev[kFlatU];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            const int64_t i = v0 + u * trip + t;
            if (i < total) {
                xv[u] = ld16(x + i);
                if (MODE == 1) qv[u] = ld16(dy + i);
                if (ADDEND) ev[u] = ld16(addend + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            const int64_t i = v0 + u * trip + t;
            if (i >= total) continue;
            float xf[8], qf[8], ef[8], r[8];
            unpack8(xv[u], xf);
            if (MODE == 1) unpack8(qv[u], qf);
            if (ADDEND) unpack8(ev[u], ef);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = xf[j] - mu[j];
                if (MODE == 0) {
                    float y = fmaf(d, ca[j], cb[j]);
                    if (RELU) y = (y > 0.f || y != y) ? y : 0.f;
                    r[j] = y;
                } else {
                    float q = qf[j];
                    if (RELU) q = fmaf(d, ca[j], cb[j]) > 0.f ? q : 0.f;
                    float v = fmaf(q, gA[j], fmaf(d, gB[j], gC[j]));
                    if (ADDEND) v = __fadd_rn(v, ef[j]);
                    r[j] = v;
                }
            }
            st16(out + i, pack8(r));
        }
    }
}

__global__ void __launch_bounds__(256) flat_add_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                       uint4* __restrict__ out, int64_t nvec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (kFlatU - 1) * stride < nvec; i += kFlatU * stride) {
        uint4 av[kFlatU], bv[kFlatU];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            av[u] = ld16(a + i + u * stride);
            bv[u] = ld16(b + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            float af[8], bf[8], r[8];
            unpack8(av[u], af);
            unpack8(bv[u], bf);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(af[j], bf[j]);
            st16(out + i + u * stride, pack8(r));
        }
    }
    for (; i < nvec; i += stride) {
        float af[8], bf[8], r[8];
        unpack8(ld16(a + i), af);
        unpack8(ld16(b + i), bf);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(af[j], bf[j]);
        st16(out + i, pack8(r));
    }
}

// [N][HW][Cp] bf16 -> [N][C][HW] fp32 through a 16-channel x PX-pixel shared-memory tile: reads are one 32-byte sector per
// pixel (neighbouring channel groups complete the line in L2), writes PX*4-byte contiguous runs per channel.
template <int PX>
__global__ void __launch_bounds__(256) unstage_kernel(const __nv_bfloat16* __restrict__ in, float* __restrict__ out, int C,
                                                      int HW, int Cp) {
    constexpr int PITCH = PX + 1;
    __shared__ float tile[16][PITCH];
    const int n = blockIdx.z;
    const int hw0 = blockIdx.x * PX, c0 = blockIdx.y * 16;
    const __nv_bfloat16* src = in + (int64_t)n * HW * Cp;
    for (int idx = threadIdx.x; idx < PX * 2; idx += 256) {
        const int p = idx >> 1, h = idx & 1;
        const int hw = hw0 + p, c = c0 + h * 8;
        float f[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
        if (hw < HW && c < Cp) unpack8(*(const uint4*)(src + (int64_t)hw * Cp + c), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) tile[h * 8 + j][p] = f[j];
    }
    __syncthreads();
    float* dst = out + (int64_t)n * C * HW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cl = w + half * 8, c = c0 + cl;
        if (c >= C) continue;
        for (int p = lane; p < PX; p += 32) {
            const int hw = hw0 + p;
            if (hw < HW) dst[(int64_t)c * HW + hw] = tile[cl][p];
        }
    }
}

struct FlatLaunch {
    int G, R, threads, blocks;
};
static FlatLaunch flat_launch(const FlatGeom& g, int ctas_per_sm) {
    FlatLaunch L;
    L.G = g.Cp / 8;
    L.R = std::max(1, 256 / L.G);
    L.threads = L.R * L.G;
    const int64_t per_block = (int64_t)kFlatU * L.R;   // pixels per CTA trip
    static int env_ctas = -1;
    if (env_ctas < 0) {
        const char* e = getenv("DOPT_B200_FLAT_CTAS");
        env_ctas = e ? std::max(1, std::min(16, atoi(e))) : 0;
    }
    if (env_ctas > 0) ctas_per_sm = env_ctas;
    L.blocks = (int)std::min<int64_t>(ceil_div(g.P, per_block), (int64_t)sm_count() * ctas_per_sm);
    if (L.blocks < 1) L.blocks = 1;
    return L;
}

}  // namespace

bool flat_supported(int64_t N, int64_t C, int64_t HW) {
    return C >= 1 && (C + 7) / 8 <= 256 && N >= 1 && HW >= 1 && N * HW < (1ll << 40);
}
size_t flat_bn_workspace_bytes(int C) { return (size_t)2 * C * sizeof(double) + 16 + (size_t)4 * C * sizeof(float); }

const float* flat_bn_train(const FlatBnTrain& a, const FlatGeom& g, cudaStream_t s) {
    FlatFin fin{};
    fin.acc = (double*)a.workspace;
    fin.counter = (unsigned*)(fin.acc + 2 * g.C);
    fin.coef = (float*)((char*)fin.counter + 16);
    fin.scale = a.scale; fin.bias = a.bias; fin.rmean = a.rmean; fin.rvar = a.rvar;
    fin.out0 = a.new_mean; fin.out1 = a.new_var;
    fin.mean2 = a.mean2; fin.var2 = a.var2;
    fin.factor = a.factor;
    const FlatLaunch Ls = flat_launch(g, 4);
    flat_bn_stats_kernel<false, false><<<Ls.blocks, Ls.threads, 0, s>>>((const uint4*)a.x, nullptr, g.P, Ls.G, g.C, Ls.R, fin);
    DB_LAUNCH_CHECK();
    const FlatLaunch La = flat_launch(g, 4);
    if (a.relu)
        flat_bn_apply_kernel<0, true, false><<<La.blocks, La.threads, 0, s>>>((const uint4*)a.x, nullptr, nullptr, (uint4*)a.y, g.P,
                                                                             La.G, g.C, La.R, fin.coef, nullptr);
    else
        flat_bn_apply_kernel<0, false, false><<<La.blocks, La.threads, 0, s>>>((const uint4*)a.x, nullptr, nullptr, (uint4*)a.y,
                                                                              g.P, La.G, g.C, La.R, fin.coef, nullptr);
    DB_LAUNCH_CHECK();
    return fin.coef;
}

void flat_bn_grad(const FlatBnGrad& a, const FlatGeom& g, cudaStream_t s) {
    FlatFin fin{};
    fin.acc = (double*)a.workspace;
    fin.counter = (unsigned*)(fin.acc + 2 * g.C);
    fin.coef = (float*)((char*)fin.counter + 16);
    fin.scale = a.scale;
    fin.out0 = a.dscale; fin.out1 = a.dbias;
    fin.fcoef = a.fcoef;
    const FlatLaunch Ls = flat_launch(g, 3);
    if (a.gate) flat_bn_stats_kernel<true, true><<<Ls.blocks, Ls.threads, 0, s>>>((const uint4*)a.x, (const uint4*)a.dy, g.P, Ls.G, g.C, Ls.R, fin);
    else flat_bn_stats_kernel<true, false><<<Ls.blocks, Ls.threads, 0, s>>>((const uint4*)a.x, (const uint4*)a.dy, g.P, Ls.G, g.C, Ls.R, fin);
    DB_LAUNCH_CHECK();
    const FlatLaunch La = flat_launch(g, 2);
    const uint4 *x = (const uint4*)a.x, *dy = (const uint4*)a.dy, *ad = (const uint4*)a.addend;
    uint4* out = (uint4*)a.dx;
#define FLAT_GRAD_APPLY(RELU, ADD) \
    flat_bn_apply_kernel<1, RELU, ADD><<<La.blocks, La.threads, 0, s>>>(x, dy, ad, out, g.P, La.G, g.C, La.R, fin.coef, a.fcoef)
    if (a.gate && ad) FLAT_GRAD_APPLY(true, true);
    else if (a.gate) FLAT_GRAD_APPLY(true, false);
    else if (ad) FLAT_GRAD_APPLY(false, true);
    else FLAT_GRAD_APPLY(false, false);
#undef FLAT_GRAD_APPLY
    DB_LAUNCH_CHECK();
}

void flat_add(const void* a, const void* b, void* out, int64_t n_elems, cudaStream_t s) {
    DB_REQUIRE(n_elems % 8 == 0, "flat_add: element count must be a multiple of 8");
    const int64_t nvec = n_elems / 8;
    if (nvec == 0) return;
    flat_add_kernel<<<stream_grid(ceil_div(nvec, kFlatU), 256, 6), 256, 0, s>>>((const uint4*)a, (const uint4*)b, (uint4*)out, nvec);
    DB_LAUNCH_CHECK();
}

void unstage_nhwc_bf16_to_nchw(const void* in, float* out, int N, int C, int64_t HW, cudaStream_t s) {
    const int Cp = (C + 7) / 8 * 8;
    if (HW > 64) {
        dim3 grid((unsigned)ceil_div(HW, (int64_t)256), (unsigned)ceil_div(Cp, 16), (unsigned)N);
        unstage_kernel<256><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, out, C, (int)HW, Cp);
    } else {
        dim3 grid((unsigned)ceil_div(HW, (int64_t)64), (unsigned)ceil_div(Cp, 16), (unsigned)N);
        unstage_kernel<64><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, out, C, (int)HW, Cp);
    }
    DB_LAUNCH_CHECK();
}

}  // namespace db
