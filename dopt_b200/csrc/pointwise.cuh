// pointwise.cuh -- op table and per-element semantics shared by pointwise.cu and the plan's fused-region kernel.
#pragma once
#include "common.cuh"

namespace dbk {

enum PwOp {
    OP_ADD = 0, OP_SUB, OP_MUL, OP_DIV, OP_LT, OP_LTE, OP_GT, OP_GTE, OP_EQ, OP_NEQ, OP_MAX, OP_MIN, OP_POW,
    OP_NEG, OP_ABS, OP_SGN, OP_EXP, OP_LOG, OP_SQRT,
    // unary functions the reference only has CPU kernels for (cpu/source/dopt/cpu/math.d:323-324: `op(cast(float)x)`) and
    // reaches from the CUDA backend through the CUDACPUKernel round trip (cuda/source/dopt/cuda/package.d:81-119)
    OP_SIN, OP_COS, OP_TAN, OP_ASIN, OP_ACOS, OP_ATAN, OP_SINH, OP_COSH, OP_TANH, OP_ASINH, OP_ACOSH, OP_ATANH,
    OP_COUNT
};
// ops below this id can be instructions of a fused pointwise region (fused.cu); the transcendental tail runs as plain launches
static constexpr int OP_FUSABLE_END = OP_SIN;
enum PwBroadcast { B_TENSOR = 0, B_SCALAR_B = 1, B_SCALAR_A = 2 };

template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<int> { using type = int4; };

__device__ __forceinline__ float4 ldv(const float4* p) { return ld_stream(p); }
__device__ __forceinline__ void stv(float4* p, const float4& v) { st_stream(p, v); }
__device__ __forceinline__ int4 ldv(const int4* p) {
    int4 r;
    asm volatile("ld.global.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stv(int4* p, const int4& v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
                 "r"(v.w)
                 : "memory");
}

// float semantics: the CUDA C overloads the reference's NVRTC template resolves to (math.d:129-170).
// Explicit _rn intrinsics keep the compiler from contracting neighbouring ops into FMAs when these are inlined
// into fused kernels, so fused and unfused results are bit-identical.
template <int OP, typename T> struct Apply;
template <int OP> struct Apply<OP, float> {
    static __device__ __forceinline__ float f(float a, float b) {
        switch (OP) {
            case OP_ADD: return __fadd_rn(a, b);
            case OP_SUB: return __fsub_rn(a, b);
            case OP_MUL: return __fmul_rn(a, b);
            case OP_DIV: return __fdiv_rn(a, b);
            case OP_LT: return a < b ? 1.0f : 0.0f;
            case OP_LTE: return a <= b ? 1.0f : 0.0f;
            case OP_GT: return a > b ? 1.0f : 0.0f;
            case OP_GTE: return a >= b ? 1.0f : 0.0f;
            case OP_EQ: return a == b ? 1.0f : 0.0f;
            case OP_NEQ: return a != b ? 1.0f : 0.0f;
            case OP_MAX: return fmaxf(a, b);
            case OP_MIN: return fminf(a, b);
            case OP_POW: return powf(a, b);
            case OP_NEG: return -a;
            case OP_ABS: return fabsf(a);
            case OP_SGN: return (float)((0.0f < a) - (a < 0.0f));
            case OP_EXP: return expf(a);
            case OP_LOG: return logf(a);
            case OP_SQRT: return __fsqrt_rn(a);
            case OP_SIN: return sinf(a);
            case OP_COS: return cosf(a);
            case OP_TAN: return tanf(a);
            case OP_ASIN: return asinf(a);
            case OP_ACOS: return acosf(a);
            case OP_ATAN: return atanf(a);
            case OP_SINH: return sinhf(a);
            case OP_COSH: return coshf(a);
            case OP_TANH: return tanhf(a);
            case OP_ASINH: return asinhf(a);
            case OP_ACOSH: return acoshf(a);
            case OP_ATANH: return atanhf(a);
        }
        return 0.0f;
    }
};
// integer power as D's std.math.pow(int, int) computes it (exact, by squaring); negative exponents give 1/x^n truncated
__device__ __forceinline__ int ipow(int a, int b) {
    if (b < 0) return a == 1 ? 1 : (a == -1 ? ((b & 1) ? -1 : 1) : 0);
    int r = 1;
    while (b) {
        if (b & 1) r *= a;
        a *= a;
        b >>= 1;
    }
    return r;
}
template <int OP> struct Apply<OP, int> {
    static __device__ __forceinline__ int f(int a, int b) {
        switch (OP) {
            case OP_ADD: return a + b;
            case OP_SUB: return a - b;
            case OP_MUL: return a * b;
            case OP_DIV: return b == 0 ? 0 : a / b;   // C leaves x/0 undefined; pick 0 rather than trapping
            case OP_LT: return a < b;
            case OP_LTE: return a <= b;
            case OP_GT: return a > b;
            case OP_GTE: return a >= b;
            case OP_EQ: return a == b;
            case OP_NEQ: return a != b;
            case OP_MAX: return max(a, b);
            case OP_MIN: return min(a, b);
            case OP_POW: return ipow(a, b);
            case OP_NEG: return -a;
            case OP_ABS: return abs(a);
            case OP_SGN: return (0 < a) - (a < 0);
            case OP_EXP: return (int)exp((double)a);
            case OP_LOG: return (int)log((double)a);
            case OP_SQRT: return (int)sqrt((double)a);
            // cast(int)(f(cast(float)a)), cpu/source/dopt/cpu/math.d:416-423
            case OP_SIN: return (int)sinf((float)a);
            case OP_COS: return (int)cosf((float)a);
            case OP_TAN: return (int)tanf((float)a);
            case OP_ASIN: return (int)asinf((float)a);
            case OP_ACOS: return (int)acosf((float)a);
            case OP_ATAN: return (int)atanf((float)a);
            case OP_SINH: return (int)sinhf((float)a);
            case OP_COSH: return (int)coshf((float)a);
            case OP_TANH: return (int)tanhf((float)a);
            case OP_ASINH: return (int)asinhf((float)a);
            case OP_ACOSH: return (int)acoshf((float)a);
            case OP_ATANH: return (int)atanhf((float)a);
        }
        return 0;
    }
};
template <int OP, typename T> __device__ __forceinline__ T apply(T a, T b) { return Apply<OP, T>::f(a, b); }

}  // namespace dbk

namespace db {
// op: dbk::PwOp; bmode: dbk::PwBroadcast (scalar operands are DEVICE pointers to one element)
void pointwise_launch(int op, int dtype, int bmode, const void* a, const void* b, void* o, int64_t n, cudaStream_t s);
int pointwise_op_id(const char* name);
bool pointwise_is_unary(int op);
inline bool pointwise_fusable(int op) { return op >= 0 && op < dbk::OP_FUSABLE_END; }
}  // namespace db
