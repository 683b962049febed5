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
static constexpr int kFlatU = 4;           // 16-byte loads in flight per operand per thread (apply / add kernels)

namespace {

// per batch-norm kernel object: [epochs (16 B) | sums: 2 buffers x kFlatCopies x 2 * Cp floats | coef: 4 * C floats]
//   sums   train: sum(x - K), sum((x - K)^2) per channel; grad: sum(g), sum(g * (x - mean)).  Accumulated with fire-and-forget
//          fp32 atomics by the statistics kernel -- CTA b adds into copy b % kFlatCopies, because atomics on ONE address
//          serialise at ~25 ns each (profiles/r02_flat_bn.md) -- and summed by the apply kernel's prologue.
//          Two buffers alternate between launches: the statistics kernel of launch e accumulates into buffer e & 1 and its
//          first CTA clears the other one (last read by the apply kernel of launch e - 1, long finished).  The launch number
//          lives in device memory (a CUDA graph replays the same arguments every step): epoch[0] is read by the statistics
//          kernel and written by the apply kernel, epoch[1] the other way round, so no kernel reads a word it writes.
//   coef   forward [mean | a | b | istd], written by the first CTA of the forward apply kernel, read by the backward kernels
static constexpr int kFlatCopies = 8;
struct FlatWs {
    unsigned* epoch;
    float* sums;
    float* coef;
};
struct FlatApplyArgs {
    FlatWs ws;
    const float* scale; const float* bias; const float* rmean; const float* rvar;
    float* out0; float* out1;       // train: new mean / new var; grad: dscale / dbias
    float* mean2; float* var2;
    const float* fcoef;             // grad: forward coefficients [mean | a | b | istd]
    double factor;
    int pivot_zero;                 // train: the sums are plain sum(x), sum(x^2) (convolution epilogue) instead of pivoted by pixel 0
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
// Per-thread fp32 partials -> shared-memory reduction over the CTA's pixel rows -> one fire-and-forget fp32 atomic per CTA,
// channel and sum.  Nothing waits for the result inside this kernel: the apply kernel turns the sums into coefficients in
// its prologue (profiles/r02_flat_bn.md: a "last CTA finalizes" tail cost more than the streaming of these 10-40 MB tensors).
// U = independent 16-byte loads in flight per operand per thread.
// epoch protocol of a statistics producer (see FlatWs): returns this CTA's accumulator copy
__device__ __forceinline__ float* flat_stats_begin(const FlatWs& ws, int Cp) {
    const unsigned e = ws.epoch[0];
    if (blockIdx.x == 0) {
        float* other = ws.sums + (size_t)((e + 1u) & 1u) * kFlatCopies * 2 * Cp;
        for (int i = threadIdx.x; i < kFlatCopies * 2 * Cp; i += blockDim.x) other[i] = 0.f;
        if (threadIdx.x == 0) ws.epoch[1] = e;
    }
    return ws.sums + ((size_t)(e & 1u) * kFlatCopies + blockIdx.x % kFlatCopies) * 2 * Cp;
}
// per-thread partials (8 channels x 2 sums) -> CTA sums -> one atomic per channel and sum
__device__ __forceinline__ void flat_stats_end(const float (&s1)[8], const float (&s2)[8], float (&red)[16][256], float* sums, int G,
                                               int R) {
    const int t = threadIdx.x;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        red[j][t] = s1[j];
        red[8 + j][t] = s2[j];
    }
    __syncthreads();
    const int Cp = G * 8;
    for (int idx = t; idx < G * 16; idx += blockDim.x) {
        const int g2 = idx % G, k = idx / G;
        float v = 0.f;
        for (int r = 0; r < R; ++r) v += red[k][r * G + g2];
        atomicAdd(sums + (k >> 3) * Cp + g2 * 8 + (k & 7), v);
    }
}

template <bool GRAD, bool GATE, int U>
__global__ void __launch_bounds__(256) flat_bn_stats_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                            int64_t P, int G, int C, int R, const __grid_constant__ FlatWs ws,
                                                            const float* __restrict__ fcoef) {
    __shared__ float red[16][256];
    const int t = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    float* sums = flat_stats_begin(ws, G * 8);
    const int cg_ = t % G;
    float piv[8], fa[8], fb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) piv[j] = fa[j] = fb[j] = 0.f;
    if (!GRAD) {
        unpack8(x[cg_], piv);   // pixel 0
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cg_ * 8 + j;
            if (c < C) {
                piv[j] = fcoef[c];   // the batch mean
                if (GATE) { fa[j] = fcoef[C + c]; fb[j] = fcoef[2 * C + c]; }
            }
        }
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int64_t trip = (int64_t)R * G, total = P * G;
    for (int64_t v0 = (int64_t)blockIdx.x * U * trip; v0 < total; v0 += (int64_t)gridDim.x * U * trip) {
        uint4 xv[U], qv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = v0 + u * trip + t;
            if (i < total) {
                xv[u] = ld16(x + i);
                if (GRAD) qv[u] = ld16(dy + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (v0 + u * trip + t >= total) continue;
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
    flat_stats_end(s1, s2, red, sums, G, R);
}

// ---- apply -------------------------------------------------------------------------------------------------------------
// Prologue (every CTA, for the 8 channels of each thread): sums -> per-channel coefficients, same algebra as
// bn_train_finalize_channel / bn_grad_finalize_channel (batchnorm.cu) in fp32.  The first CTA also writes the forward
// coefficient block, the running statistics (train) or dscale / dbias (grad); the CTA that finishes its prologue last clears
// the sums for the next launch.
// MODE 0: y = relu?((x - mean) * a + b)
// MODE 1: dx = g * A + (x - mean) * B + Cc (+ addend), g = dy gated by the forward relu
template <int MODE, bool RELU, bool ADDEND>
__global__ void __launch_bounds__(256) flat_bn_apply_kernel(const uint4* __restrict__ x, const uint4* __restrict__ dy,
                                                            const uint4* __restrict__ addend, uint4* __restrict__ out,
                                                            int64_t P, int G, int C, int R,
                                                            const __grid_constant__ FlatApplyArgs a) {
    __shared__ __align__(16) float ssum[2 * 2048];   // [2][Cp]: the launch's sums, copies added in a fixed order
    const int t = threadIdx.x;
    const int cg_ = t % G;
    const int Cp = G * 8;
    const float invM = 1.0f / (float)P;
    pdl_trigger();
    pdl_wait();
    // All loads of the prologue are issued back to back (one L2 round trip instead of three): the epoch, BOTH sum buffers
    // (the current one is selected afterwards), and further down the per-channel parameters.  Thread f < 2*Cp/4 owns float4
    // column f of the sums and adds its kFlatCopies copies in a fixed order.
    const unsigned e = __ldcg(a.ws.epoch + 1);
    const int nf4 = 2 * Cp / 4;
    for (int f = t; f < nf4; f += blockDim.x) {
        const float4* b0 = (const float4*)a.ws.sums + f;
        const float4* b1 = b0 + (size_t)kFlatCopies * nf4;
        float4 v0[kFlatCopies], v1[kFlatCopies];
#pragma unroll
        for (int k = 0; k < kFlatCopies; ++k) {
            v0[k] = __ldcg(b0 + (size_t)k * nf4);
            v1[k] = __ldcg(b1 + (size_t)k * nf4);
        }
        const bool odd = e & 1u;
        float4 r = odd ? v1[0] : v0[0];
#pragma unroll
        for (int k = 1; k < kFlatCopies; ++k) {
            const float4 w = odd ? v1[k] : v0[k];
            r.x += w.x; r.y += w.y; r.z += w.z; r.w += w.w;
        }
        *(float4*)&ssum[f * 4] = r;
    }
    __syncthreads();
    if (blockIdx.x == 0 && t == 0) a.ws.epoch[0] = e + 1u;   // read by the NEXT statistics kernel only
    float mu[8], ca[8], cb[8], gA[8], gB[8], gC[8];
    float piv[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) piv[j] = 0.f;
    if (MODE == 0 && !a.pivot_zero) unpack8(x[cg_], piv);   // the statistics producer's pivot: pixel 0
    const bool writer = blockIdx.x == 0 && t < G;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg_ * 8 + j;
        mu[j] = ca[j] = cb[j] = gA[j] = gB[j] = gC[j] = 0.f;
        if (c >= C) continue;
        const float q0 = ssum[c], q1 = ssum[Cp + c];
        if (MODE == 0) {
            const float d = q0 * invM;
            const float mean = piv[j] + d;
            const float var = fmaxf(fmaf(-d, d, q1 * invM), 0.f);
            const float istd = 1.0f / sqrtf(var + (float)kFlatEps);
            mu[j] = mean;
            ca[j] = a.scale[c] * istd;
            cb[j] = a.bias[c];
            if (writer) {
                a.ws.coef[c] = mean;
                a.ws.coef[C + c] = ca[j];
                a.ws.coef[2 * C + c] = cb[j];
                a.ws.coef[3 * C + c] = istd;
                const double M = (double)P;
                const double unbiased = M > 1 ? (double)var * M / (M - 1) : (double)var;
                const float nm = (float)((double)a.rmean[c] * (1.0 - a.factor) + (double)mean * a.factor);
                const float nv = (float)((double)a.rvar[c] * (1.0 - a.factor) + unbiased * a.factor);
                a.out0[c] = nm;
                a.out1[c] = nv;
                if (a.mean2) a.mean2[c] = nm;
                if (a.var2) a.var2[c] = nv;
            }
        } else {
            mu[j] = a.fcoef[c];
            if (RELU) { ca[j] = a.fcoef[C + c]; cb[j] = a.fcoef[2 * C + c]; }
            const float istd = a.fcoef[3 * C + c];
            const float dbeta = q0, dgamma = q1 * istd;
            const float A = a.scale[c] * istd;
            gA[j] = A;
            gB[j] = -A * istd * dgamma * invM;
            gC[j] = -A * dbeta * invM;
            if (writer) {
                a.out0[c] = dgamma;
                a.out1[c] = dbeta;
            }
        }
    }
    const int64_t trip = (int64_t)R * G, total = P * G;
    // (L2 residency hints -- evict_last in the statistics pass, evict_first here -- and walking the tensor from its end were
    // measured: no net gain, profiles/r02_summary.md)
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
    pdl_trigger();
    pdl_wait();
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

// out = a + b with the batch-norm statistics of `out` (the residual sum feeding the next block's first batch norm) accumulated
// on the way: same thread <-> channel-group mapping and the same tail as flat_bn_stats_kernel, pivot = out[pixel 0].
__global__ void __launch_bounds__(256) flat_add_stats_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                             uint4* __restrict__ out, int64_t P, int G, int R,
                                                             const __grid_constant__ FlatWs ws) {
    __shared__ float red[16][256];
    const int t = threadIdx.x;
    pdl_trigger();
    pdl_wait();
    float* sums = flat_stats_begin(ws, G * 8);
    const int cg_ = t % G;
    float piv[8];
    {
        float af[8], bf[8], r[8];
        unpack8(a[cg_], af);
        unpack8(b[cg_], bf);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(af[j], bf[j]);
        unpack8(pack8(r), piv);   // the stored (bf16-rounded) value of pixel 0
    }
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    const int64_t trip = (int64_t)R * G, total = P * G;
    for (int64_t v0 = (int64_t)blockIdx.x * kFlatU * trip; v0 < total; v0 += (int64_t)gridDim.x * kFlatU * trip) {
        uint4 av[kFlatU], bv[kFlatU];
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            const int64_t i = v0 + u * trip + t;
            if (i < total) {
                av[u] = ld16(a + i);
                bv[u] = ld16(b + i);
            }
        }
#pragma unroll
        for (int u = 0; u < kFlatU; ++u) {
            const int64_t i = v0 + u * trip + t;
            if (i >= total) continue;
            float af[8], bf[8], r[8];
            unpack8(av[u], af);
            unpack8(bv[u], bf);
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(af[j], bf[j]);
            const uint4 o = pack8(r);
            st16(out + i, o);
            unpack8(o, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = r[j] - piv[j];
                s1[j] += d;
                s2[j] = fmaf(d, d, s2[j]);
            }
        }
    }
    flat_stats_end(s1, s2, red, sums, G, R);
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
static FlatLaunch flat_launch(const FlatGeom& g, int ctas_per_sm, int unroll = kFlatU) {
    FlatLaunch L;
    L.G = g.Cp / 8;
    L.R = std::max(1, 256 / L.G);
    L.threads = L.R * L.G;
    const int64_t per_block = (int64_t)unroll * L.R;   // pixels per CTA trip
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

static constexpr int kStatsCtasPerSm = 2;

}  // namespace

bool flat_supported(int64_t N, int64_t C, int64_t HW) {
    return C >= 1 && (C + 7) / 8 <= 256 && N >= 1 && HW >= 1 && N * HW < (1ll << 31);
}
size_t flat_bn_workspace_bytes(int C) {
    const int Cp = (C + 7) / 8 * 8;
    return 16 + (size_t)2 * kFlatCopies * 2 * Cp * sizeof(float) + (size_t)4 * C * sizeof(float);
}
static FlatWs flat_ws(void* workspace, int C) {
    const int Cp = (C + 7) / 8 * 8;
    FlatWs w;
    w.epoch = (unsigned*)workspace;
    w.sums = (float*)((char*)workspace + 16);
    w.coef = w.sums + (size_t)2 * kFlatCopies * 2 * Cp;
    return w;
}

const float* flat_bn_train(const FlatBnTrain& a, const FlatGeom& g, cudaStream_t s) {
    FlatApplyArgs ap{};
    ap.ws = flat_ws(a.workspace, g.C);
    ap.scale = a.scale; ap.bias = a.bias; ap.rmean = a.rmean; ap.rvar = a.rvar;
    ap.out0 = a.new_mean; ap.out1 = a.new_var;
    ap.mean2 = a.mean2; ap.var2 = a.var2;
    ap.factor = a.factor;
    ap.pivot_zero = a.stats_source == 2 ? 1 : 0;
    if (a.stats_source == 0) {
        const FlatLaunch Ls = flat_launch(g, kStatsCtasPerSm, 8);
        DB_CUDA(launch_pdl(flat_bn_stats_kernel<false, false, 8>, dim3(Ls.blocks), dim3(Ls.threads), 0, s, (const uint4*)a.x,
                           (const uint4*)nullptr, g.P, Ls.G, g.C, Ls.R, ap.ws, (const float*)nullptr));
        count_launch();
    }
    const FlatLaunch La = flat_launch(g, 4);
    const uint4* none = nullptr;
    if (a.relu)
        DB_CUDA(launch_pdl(flat_bn_apply_kernel<0, true, false>, dim3(La.blocks), dim3(La.threads), 0, s, (const uint4*)a.x, none, none,
                           (uint4*)a.y, g.P, La.G, g.C, La.R, ap));
    else
        DB_CUDA(launch_pdl(flat_bn_apply_kernel<0, false, false>, dim3(La.blocks), dim3(La.threads), 0, s, (const uint4*)a.x, none, none,
                           (uint4*)a.y, g.P, La.G, g.C, La.R, ap));
    count_launch();
    return ap.ws.coef;
}

void flat_bn_grad(const FlatBnGrad& a, const FlatGeom& g, cudaStream_t s) {
    FlatApplyArgs ap{};
    ap.ws = flat_ws(a.workspace, g.C);
    ap.scale = a.scale;
    ap.out0 = a.dscale; ap.out1 = a.dbias;
    ap.fcoef = a.fcoef;
    const FlatLaunch Ls = flat_launch(g, kStatsCtasPerSm, 4);
    const uint4 *x = (const uint4*)a.x, *dy = (const uint4*)a.dy, *ad = (const uint4*)a.addend;
    if (!a.stats_from_producer) {
        if (a.gate) DB_CUDA(launch_pdl(flat_bn_stats_kernel<true, true, 4>, dim3(Ls.blocks), dim3(Ls.threads), 0, s, x, dy, g.P, Ls.G, g.C, Ls.R, ap.ws, a.fcoef));
        else DB_CUDA(launch_pdl(flat_bn_stats_kernel<true, false, 4>, dim3(Ls.blocks), dim3(Ls.threads), 0, s, x, dy, g.P, Ls.G, g.C, Ls.R, ap.ws, a.fcoef));
        count_launch();
    }
    const FlatLaunch La = flat_launch(g, 2);
    uint4* out = (uint4*)a.dx;
#define FLAT_GRAD_APPLY(RELU, ADD) \
    DB_CUDA(launch_pdl(flat_bn_apply_kernel<1, RELU, ADD>, dim3(La.blocks), dim3(La.threads), 0, s, x, dy, ad, out, g.P, La.G, g.C, La.R, ap))
    if (a.gate && ad) FLAT_GRAD_APPLY(true, true);
    else if (a.gate) FLAT_GRAD_APPLY(true, false);
    else if (ad) FLAT_GRAD_APPLY(false, true);
    else FLAT_GRAD_APPLY(false, false);
#undef FLAT_GRAD_APPLY
    count_launch();
}

void flat_stats_sink(void* workspace, int C, unsigned** epoch, float** sums, int* copies) {
    const FlatWs w = flat_ws(workspace, C);
    *epoch = w.epoch;
    *sums = w.sums;
    *copies = kFlatCopies;
}

void flat_add_stats(const void* a, const void* b, void* out, const FlatGeom& g, void* bn_workspace, cudaStream_t s) {
    const FlatLaunch L = flat_launch(g, kStatsCtasPerSm);
    DB_CUDA(launch_pdl(flat_add_stats_kernel, dim3(L.blocks), dim3(L.threads), 0, s, (const uint4*)a, (const uint4*)b, (uint4*)out,
                       g.P, L.G, L.R, flat_ws(bn_workspace, g.C)));
    count_launch();
}

void flat_add(const void* a, const void* b, void* out, int64_t n_elems, cudaStream_t s) {
    DB_REQUIRE(n_elems % 8 == 0, "flat_add: element count must be a multiple of 8");
    const int64_t nvec = n_elems / 8;
    if (nvec == 0) return;
    DB_CUDA(launch_pdl(flat_add_kernel, dim3(stream_grid(ceil_div(nvec, kFlatU), 256, 6)), dim3(256), 0, s, (const uint4*)a,
                       (const uint4*)b, (uint4*)out, nvec));
    count_launch();
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
