// pointwise.cu -- the 19 elementwise ops of dopt.cuda.math (cuda/source/dopt/cuda/math.d:79-207) and the 12 unary functions
// (sin ... atanh) the reference evaluates on the host for the CUDA backend (cuda/source/dopt/cuda/package.d:81-119).
//
// Reference: one NVRTC template `out[i] = a[i] OP b[i]` / `out[i] = f(a[i])`, T in {float,int}, 512 threads, scalar 4-byte
// accesses, cuCtxSynchronize after every launch (math.d:129-170,199-201; nvrtc.d:111).
// Here: one grid-stride kernel per (op, T), 128-bit loads/stores, 4 independent vectors in flight per thread, no sync.
// Semantics are those of the CUDA C expressions the reference compiles:
//   comparisons yield T(0/1); sgn = (0<a)-(a<0); max/min/pow/abs/exp/log/sqrt resolve to the CUDA math overloads
//   (float: fmaxf/fminf/powf/fabsf/expf/logf/sqrtf, full precision; int: integer max/min/abs, and pow/exp/log/sqrt
//   through double with a truncating conversion back to int).
// HBM-bound: 2V*4 B (unary) or 3V*4 B (binary) per launch.
#include "common.cuh"
#include "flat.cuh"
#include "pointwise.cuh"

namespace db {

template <int OP, typename T, int BMODE>
__global__ void __launch_bounds__(256) pw_kernel(const T* __restrict__ a, const T* __restrict__ b, T* __restrict__ o,
                                                 int64_t n) {
    using V = typename dbk::Vec4<T>::type;
    const int64_t nvec = n >> 2;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    T sa = T(0), sb = T(0);
    if (BMODE == dbk::B_SCALAR_A) sa = a[0];
    if (BMODE == dbk::B_SCALAR_B) sb = b[0];
    const V* av = reinterpret_cast<const V*>(a);
    const V* bv = reinterpret_cast<const V*>(b);
    V* ov = reinterpret_cast<V*>(o);
    constexpr int U = 4;
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        V x[U], y[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (BMODE != dbk::B_SCALAR_A) x[u] = dbk::ldv(av + i + u * stride);
            if (BMODE != dbk::B_SCALAR_B) y[u] = dbk::ldv(bv + i + u * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            V r;
            r.x = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x[u].x, BMODE != dbk::B_SCALAR_B ? y[u].x : sb);
            r.y = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x[u].y, BMODE != dbk::B_SCALAR_B ? y[u].y : sb);
            r.z = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x[u].z, BMODE != dbk::B_SCALAR_B ? y[u].z : sb);
            r.w = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x[u].w, BMODE != dbk::B_SCALAR_B ? y[u].w : sb);
            dbk::stv(ov + i + u * stride, r);
        }
    }
    for (; i < nvec; i += stride) {
        V x, y, r;
        if (BMODE != dbk::B_SCALAR_A) x = dbk::ldv(av + i);
        if (BMODE != dbk::B_SCALAR_B) y = dbk::ldv(bv + i);
        r.x = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x.x, BMODE != dbk::B_SCALAR_B ? y.x : sb);
        r.y = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x.y, BMODE != dbk::B_SCALAR_B ? y.y : sb);
        r.z = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x.z, BMODE != dbk::B_SCALAR_B ? y.z : sb);
        r.w = dbk::apply<OP, T>(BMODE == dbk::B_SCALAR_A ? sa : x.w, BMODE != dbk::B_SCALAR_B ? y.w : sb);
        dbk::stv(ov + i, r);
    }
    // scalar tail (n % 4 elements)
    int64_t t = (nvec << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n) {
        T x = BMODE == dbk::B_SCALAR_A ? sa : a[t];
        T y = BMODE != dbk::B_SCALAR_B ? b[t] : sb;
        o[t] = dbk::apply<OP, T>(x, y);
    }
}

// unaligned fallback (sub-buffers at odd offsets): scalar accesses
template <int OP, typename T, int BMODE>
__global__ void __launch_bounds__(256) pw_kernel_scalar(const T* __restrict__ a, const T* __restrict__ b,
                                                        T* __restrict__ o, int64_t n) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    T sa = T(0), sb = T(0);
    if (BMODE == dbk::B_SCALAR_A) sa = a[0];
    if (BMODE == dbk::B_SCALAR_B) sb = b[0];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        T x = BMODE == dbk::B_SCALAR_A ? sa : a[i];
        T y = BMODE != dbk::B_SCALAR_B ? b[i] : sb;
        o[i] = dbk::apply<OP, T>(x, y);
    }
}

template <int OP, typename T, int BMODE>
static void launch_pw(const void* a, const void* b, void* o, int64_t n, cudaStream_t s) {
    if (n <= 0) return;
    bool aligned = ((uintptr_t)o % 16 == 0) && (BMODE == dbk::B_SCALAR_A || (uintptr_t)a % 16 == 0) &&
                   (BMODE == dbk::B_SCALAR_B || (uintptr_t)b % 16 == 0);
    if (aligned) {
        int grid = stream_grid(ceil_div(n, 16), 256, 8);   // 4 vectors of 4 per thread per trip
        pw_kernel<OP, T, BMODE><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (T*)o, n);
    } else {
        int grid = stream_grid(n, 256, 16);
        pw_kernel_scalar<OP, T, BMODE><<<grid, 256, 0, s>>>((const T*)a, (const T*)b, (T*)o, n);
    }
    DB_LAUNCH_CHECK();
}

template <int OP, typename T>
static void launch_pw_mode(int bmode, const void* a, const void* b, void* o, int64_t n, cudaStream_t s) {
    switch (bmode) {
        case dbk::B_TENSOR: launch_pw<OP, T, dbk::B_TENSOR>(a, b, o, n, s); break;
        case dbk::B_SCALAR_B: launch_pw<OP, T, dbk::B_SCALAR_B>(a, b, o, n, s); break;
        case dbk::B_SCALAR_A: launch_pw<OP, T, dbk::B_SCALAR_A>(a, b, o, n, s); break;
        default: throw Error("pointwise: bad broadcast mode");
    }
}

template <typename T>
static void launch_pw_op(int op, int bmode, const void* a, const void* b, void* o, int64_t n, cudaStream_t s) {
    switch (op) {
#define C(OP) case dbk::OP: launch_pw_mode<dbk::OP, T>(bmode, a, b, o, n, s); break;
        C(OP_ADD) C(OP_SUB) C(OP_MUL) C(OP_DIV) C(OP_LT) C(OP_LTE) C(OP_GT) C(OP_GTE) C(OP_EQ) C(OP_NEQ) C(OP_MAX)
        C(OP_MIN) C(OP_POW)
#undef C
#define C(OP) case dbk::OP: launch_pw<dbk::OP, T, dbk::B_SCALAR_B>(a, a, o, n, s); break;   /* unary: b unused */
        C(OP_NEG) C(OP_ABS) C(OP_SGN) C(OP_EXP) C(OP_LOG) C(OP_SQRT)
        C(OP_SIN) C(OP_COS) C(OP_TAN) C(OP_ASIN) C(OP_ACOS) C(OP_ATAN) C(OP_SINH) C(OP_COSH) C(OP_TANH) C(OP_ASINH) C(OP_ACOSH)
        C(OP_ATANH)
#undef C
        default: throw Error("pointwise: unknown op");
    }
}

void pointwise_launch(int op, int dtype, int bmode, const void* a, const void* b, void* o, int64_t n, cudaStream_t s) {
    if (dtype == DOPT_B200_FLOAT32) launch_pw_op<float>(op, bmode, a, b, o, n, s);
    else if (dtype == DOPT_B200_INT32) launch_pw_op<int>(op, bmode, a, b, o, n, s);
    else throw Error("pointwise: unsupported dtype");
}

static const char* kNames[] = {"add", "sub", "mul", "div", "lt",  "lte", "gt",  "gte", "eq",  "neq",
                               "max", "min", "pow", "neg", "abs", "sgn", "exp", "log", "sqrt",
                               "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh"};

int pointwise_op_id(const char* name) {
    for (int i = 0; i < dbk::OP_COUNT; ++i)
        if (strcmp(name, kNames[i]) == 0) return i;
    return -1;
}
bool pointwise_is_unary(int op) { return op >= dbk::OP_NEG; }

// a + b on NCHW fp32 tensors that ALSO emits the NHWC bf16 copy the tensor-core convolutions read (the plan folds the staging
// of a gradient sum into the add that produces it).  Same tile structure as bn_apply_stage_kernel: 16 channels x PX pixels,
// PX*4-byte contiguous reads, one full 32-byte sector of bf16 per pixel.  grid (ceil(HW/PX), ceil(Cp/16), N), 256 threads.
template <int PX>
__global__ void __launch_bounds__(256) add_stage_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        float* __restrict__ out, __nv_bfloat16* __restrict__ staged, int C,
                                                        int HW, int Cp) {
    constexpr int PITCH = PX + 4;
    constexpr int V4 = PX / 4;
    constexpr int PER_CH = V4 / 32 > 0 ? V4 / 32 : 1;
    __shared__ __align__(16) float tile[16][PITCH];
    const int n = blockIdx.z;
    const int hw0 = blockIdx.x * PX, c0 = blockIdx.y * 16;
    const int64_t img = (int64_t)n * C * HW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool vec = (HW & 3) == 0 && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)out) & 15) == 0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cl = w + half * 8, c = c0 + cl;
        if (vec) {
            float4 av[PER_CH], bv[PER_CH];
#pragma unroll
            for (int u = 0; u < PER_CH; ++u) {
                const int p4 = lane + u * 32;
                const int hw = hw0 + p4 * 4;
                const bool ok = c < C && p4 < V4 && hw < HW;
                const int64_t i = img + (int64_t)c * HW + hw;
                av[u] = ok ? *(const float4*)(a + i) : make_float4(0.f, 0.f, 0.f, 0.f);
                bv[u] = ok ? *(const float4*)(b + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < PER_CH; ++u) {
                const int p4 = lane + u * 32;
                const int hw = hw0 + p4 * 4;
                float4 r;
                r.x = dbk::apply<dbk::OP_ADD, float>(av[u].x, bv[u].x); r.y = dbk::apply<dbk::OP_ADD, float>(av[u].y, bv[u].y);
                r.z = dbk::apply<dbk::OP_ADD, float>(av[u].z, bv[u].z); r.w = dbk::apply<dbk::OP_ADD, float>(av[u].w, bv[u].w);
                if (c < C && p4 < V4 && hw < HW) *(float4*)(out + img + (int64_t)c * HW + hw) = r;
                if (p4 < V4) *(float4*)&tile[cl][p4 * 4] = r;
            }
        } else {
            for (int p = lane; p < PX; p += 32) {
                const int hw = hw0 + p;
                float r = 0.f;
                if (c < C && hw < HW) {
                    const int64_t i = img + (int64_t)c * HW + hw;
                    r = dbk::apply<dbk::OP_ADD, float>(a[i], b[i]);
                    out[i] = r;
                }
                tile[cl][p] = r;
            }
        }
    }
    __syncthreads();
    __nv_bfloat16* dst = staged + (int64_t)n * HW * Cp;
    for (int idx = threadIdx.x; idx < PX * 2; idx += 256) {
        const int p = idx >> 1, h = idx & 1;
        const int hw = hw0 + p, c = c0 + h * 8;
        if (hw < HW && c < Cp) {
            __nv_bfloat162 v[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) v[k] = __floats2bfloat162_rn(tile[h * 8 + 2 * k][p], tile[h * 8 + 2 * k + 1][p]);
            *(uint4*)(dst + (int64_t)hw * Cp + c) = *(uint4*)v;
        }
    }
}

namespace {
struct PointwiseKernel : Kernel {
    int op, dtype;
    int64_t n;
    bool unary;
    int64_t aN = 0, aC = 0, aHW = 0;   // NCHW view of the result (rank 4), for the add + staging variant
    Absorb ab;
    bool can_absorb() const override {
        return op == dbk::OP_ADD && dtype == DOPT_B200_FLOAT32 && aN > 0 && aN < 65536 && aC < (1 << 24) && aHW < (1ll << 30);
    }
    // bf16-interior mode (flat.cu): both operands and the result are NHWC bf16
    bool flat = false;
    const void* staged_in[2] = {nullptr, nullptr};
    bool can_flat() const override { return can_absorb() && flat_supported(aN, aC, aHW); }
    void set_flat(bool on) override { flat = on; }
    void set_staged_input(int input, const void* p) override {
        if (input >= 0 && input < 2) staged_in[input] = p;
    }
    void* bn_ws = nullptr;   // statistics workspace of the batchNormTrain reading the sum (flat mode)
    int can_produce_stats() const override { return flat ? 1 : 0; }
    void set_stats_workspace(void* w, int channels) override { bn_ws = channels == (int)aC ? w : nullptr; }
    void set_absorbed(const Absorb& a) override {
        DB_REQUIRE(!a.relu && !a.redirect && (flat || !a.skip_fp32), "pointwise add can only absorb the NHWC staging");
        ab = a;
    }
    PointwiseKernel(const dopt_b200_op& d) {
        op = pointwise_op_id(d.op_type);
        DB_REQUIRE(op >= 0, "unknown pointwise op");
        unary = pointwise_is_unary(op);
        DB_REQUIRE(d.n_inputs == (unary ? 1 : 2), "pointwise: wrong number of operands");
        dtype = d.output.dtype;
        n = volume(d.output);
        if (d.output.rank == 4) {
            aN = d.output.shape[0];
            aC = d.output.shape[1];
            aHW = d.output.shape[2] * d.output.shape[3];
        }
        // verifier of the reference: operand types must be identical (core/source/dopt/core/ops/math.d:22-25)
        for (int i = 0; i < d.n_inputs; ++i) {
            DB_REQUIRE(d.inputs[i].dtype == dtype && volume(d.inputs[i]) == n, "pointwise: operand type mismatch");
        }
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == (unary ? 1 : 2), "pointwise: wrong number of inputs");
        if (flat) {
            DB_REQUIRE(staged_in[0] && staged_in[1] && ab.staged, "add: flat mode needs staged operands and output");
            if (bn_ws) {
                FlatGeom fg{aN * aHW, (int)aC, (int)((aC + 7) / 8 * 8)};
                flat_add_stats(staged_in[0], staged_in[1], ab.staged, fg, bn_ws, s);
                return;
            }
            flat_add(staged_in[0], staged_in[1], ab.staged, aN * aHW * ((aC + 7) / 8 * 8), s);
            return;
        }
        if (ab.staged) {
            const int Cp = (int)((aC + 7) / 8 * 8);
            if (aHW > 64) {
                dim3 grid((unsigned)ceil_div(aHW, (int64_t)256), (unsigned)ceil_div(Cp, 16), (unsigned)aN);
                add_stage_kernel<256><<<grid, 256, 0, s>>>((const float*)in[0], (const float*)in[1], (float*)out,
                                                           (__nv_bfloat16*)ab.staged, (int)aC, (int)aHW, Cp);
            } else {
                dim3 grid((unsigned)ceil_div(aHW, (int64_t)64), (unsigned)ceil_div(Cp, 16), (unsigned)aN);
                add_stage_kernel<64><<<grid, 256, 0, s>>>((const float*)in[0], (const float*)in[1], (float*)out,
                                                          (__nv_bfloat16*)ab.staged, (int)aC, (int)aHW, Cp);
            }
            DB_LAUNCH_CHECK();
            return;
        }
        pointwise_launch(op, dtype, dbk::B_TENSOR, in[0], unary ? in[0] : in[1], out, n, s);
    }
};
Kernel* make_pointwise(const dopt_b200_op& d) { return new PointwiseKernel(d); }
}  // namespace

void register_pointwise() {
    for (int i = 0; i < dbk::OP_COUNT; ++i) register_kernel(kNames[i], make_pointwise);
}

}  // namespace db
