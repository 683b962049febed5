// reduce.cu -- sum / maxElement / argmin on the GPU.
//
// The reference registers NO CUDA kernel for these (cuda/source/dopt/cuda/math.d:89-92); its CUDAPlan silently wraps the
// CPU kernel in D2H -> evaluateCPU -> H2D (cuda/source/dopt/cuda/package.d:81-119,284).  That fallback is hit by the
// loss and by every weight-decay term on every step.  The arithmetic restated here is the CPU kernel's
// (cpu/source/dopt/cpu/math.d:90-310): axes are reduced one after another, each as [outer, A, inner] -> [outer, inner];
// `sum` starts from 0, `maxElement` from -T.max, `argmin` keeps the FIRST minimum (strict `<`, math.d:281).
// fp32 summation order differs from the serial CPU loop (tree within a block), so float sums agree to rounding, not bit
// for bit; int32 results are exact.  HBM-bound: volume(in) * 4 B.
#include "common.cuh"
#include <cfloat>
#include <climits>

namespace db {

template <typename T> struct RedSum {
    static __device__ __forceinline__ T init() { return T(0); }
    static __device__ __forceinline__ T op(T a, T b) { return a + b; }
};
template <typename T> struct RedMax;
template <> struct RedMax<float> {
    static __device__ __forceinline__ float init() { return -FLT_MAX; }
    static __device__ __forceinline__ float op(float a, float b) { return a > b ? a : b; }   // std.algorithm.max
};
template <> struct RedMax<int> {
    static __device__ __forceinline__ int init() { return -INT_MAX; }
    static __device__ __forceinline__ int op(int a, int b) { return a > b ? a : b; }
};

template <typename T, class R>
__device__ __forceinline__ T block_reduce(T v, T* smem) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = R::op(v, __shfl_xor_sync(0xffffffffu, v, o));
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) smem[w] = v;
    __syncthreads();
    int nw = blockDim.x >> 5;
    v = (threadIdx.x < nw) ? smem[threadIdx.x] : R::init();
    if (w == 0) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = R::op(v, __shfl_xor_sync(0xffffffffu, v, o));
    }
    return v;   // valid in thread 0
}

// rows: out[r] = reduce_a in[r*A + a]; one CTA per (row, chunk) when A is large, partials go to `part`
template <typename T, class R>
__global__ void __launch_bounds__(256) reduce_rows(const T* __restrict__ in, T* __restrict__ out, int64_t A,
                                                   int chunks) {
    __shared__ T smem[32];
    int64_t row = blockIdx.x / chunks;
    int chunk = blockIdx.x % chunks;
    int64_t per = (A + chunks - 1) / chunks;
    int64_t lo = (int64_t)chunk * per, hi = lo + per < A ? lo + per : A;
    const T* p = in + row * A;
    T acc = R::init();
    for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = R::op(acc, p[i]);
    acc = block_reduce<T, R>(acc, smem);
    if (threadIdx.x == 0) out[(int64_t)row * chunks + chunk] = acc;
}

// columns: out[o, i] = reduce_a in[o, a, i]; thread per (o, i), coalesced over i
template <typename T, class R>
__global__ void __launch_bounds__(256) reduce_cols(const T* __restrict__ in, T* __restrict__ out, int64_t outer,
                                                   int64_t A, int64_t inner) {
    int64_t n = outer * inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = idx / inner, i = idx % inner;
        const T* p = in + o * A * inner + i;
        T acc = R::init();
        for (int64_t a = 0; a < A; ++a) acc = R::op(acc, p[a * inner]);
        out[idx] = acc;
    }
}

template <typename T>
__global__ void __launch_bounds__(256) argmin_kernel(const T* __restrict__ in, int* __restrict__ out, int64_t outer,
                                                     int64_t A, int64_t inner, T tmax) {
    int64_t n = outer * inner;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * blockDim.x) {
        int64_t o = idx / inner, i = idx % inner;
        const T* p = in + o * A * inner + i;
        T best = tmax;
        int arg = 0;
        for (int64_t a = 0; a < A; ++a) {
            T v = p[a * inner];
            if (v < best) {
                best = v;
                arg = (int)a;
            }
        }
        out[idx] = arg;
    }
}

template <typename T, class R>
static void reduce_axis(const T* in, T* out, int64_t outer, int64_t A, int64_t inner, Scratch& ws, cudaStream_t s) {
    if (outer * inner == 0) return;   // empty result: nothing to launch (a zero-sized grid is an invalid configuration)
    if (inner == 1) {
        // split long rows over several CTAs so that a full reduction (outer == 1) still fills the chip
        int chunks = 1;
        if (outer < 2 * sm_count() && A >= 4096) {
            int64_t want = ceil_div(4 * sm_count(), outer);
            int64_t maxc = ceil_div(A, 2048);
            chunks = (int)(want < maxc ? want : maxc);
            if (chunks < 1) chunks = 1;
        }
        if (chunks == 1) {
            reduce_rows<T, R><<<(unsigned)outer, 256, 0, s>>>(in, out, A, 1);
            DB_LAUNCH_CHECK();
        } else {
            T* part = (T*)ws.get((size_t)outer * chunks * sizeof(T));
            reduce_rows<T, R><<<(unsigned)(outer * chunks), 256, 0, s>>>(in, part, A, chunks);
            DB_LAUNCH_CHECK();
            reduce_rows<T, R><<<(unsigned)outer, 256, 0, s>>>(part, out, chunks, 1);
            DB_LAUNCH_CHECK();
        }
    } else {
        reduce_cols<T, R><<<stream_grid(outer * inner, 256, 8), 256, 0, s>>>(in, out, outer, A, inner);
        DB_LAUNCH_CHECK();
    }
}

namespace {

struct Step {
    int64_t outer, A, inner;
};

struct ReduceKernel : Kernel {
    bool is_max;
    int dtype;
    std::vector<Step> steps;
    int64_t in_vol, out_vol;
    Scratch tmp[2], ws;
    ReduceKernel(const dopt_b200_op& d, bool mx) : is_max(mx) {
        const auto& in = d.inputs[0];
        DB_REQUIRE(d.n_inputs == 1, "reduction: one operand");
        dtype = in.dtype;
        in_vol = volume(in);
        out_vol = volume(d.output);
        std::vector<int64_t> shape(in.shape, in.shape + in.rank);
        // verifier: axes in range and unique (core/source/dopt/core/ops/math.d:144-160)
        for (int i = 0; i < d.n_axes; ++i) {
            DB_REQUIRE(d.axes[i] >= 0 && d.axes[i] < in.rank, "reduction: axis out of range");
            for (int j = 0; j < i; ++j) DB_REQUIRE(d.axes[j] != d.axes[i], "reduction: duplicate axis");
        }
        // sequential per-axis passes exactly like cpu/math.d:124-150; adjacent axes are merged into one pass
        int i = 0;
        std::vector<int64_t> ax(d.axes, d.axes + d.n_axes);
        while (i < (int)ax.size()) {
            int64_t a0 = ax[i];
            int64_t A = shape[a0];
            int64_t a_hi = a0;
            shape[a0] = 1;
            int j = i + 1;
            while (j < (int)ax.size() && ax[j] == a_hi + 1) {   // merge runs of increasing adjacent axes
                A *= shape[ax[j]];
                shape[ax[j]] = 1;
                a_hi = ax[j];
                ++j;
            }
            int64_t outer = 1, inner = 1;
            for (int64_t k = 0; k < a0; ++k) outer *= shape[k];
            for (int64_t k = a_hi + 1; k < (int64_t)shape.size(); ++k) inner *= shape[k];
            steps.push_back({outer, A, inner});
            i = j;
        }
        int64_t v = 1;
        for (auto x : shape) v *= x;
        DB_REQUIRE(v == out_vol, "reduction: output volume mismatch");
    }
    template <typename T> void go(const void* in, void* out, cudaStream_t s) {
        if (steps.empty()) {
            DB_CUDA(cudaMemcpyAsync(out, in, (size_t)out_vol * sizeof(T), cudaMemcpyDeviceToDevice, s));
            count_launch();
            return;
        }
        const T* cur = (const T*)in;
        for (size_t i = 0; i < steps.size(); ++i) {
            bool last = (i + 1 == steps.size());
            T* dst = last ? (T*)out : (T*)tmp[i & 1].get((size_t)steps[i].outer * steps[i].inner * sizeof(T));
            if (is_max) reduce_axis<T, RedMax<T>>(cur, dst, steps[i].outer, steps[i].A, steps[i].inner, ws, s);
            else reduce_axis<T, RedSum<T>>(cur, dst, steps[i].outer, steps[i].A, steps[i].inner, ws, s);
            cur = dst;
        }
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "reduction: one input");
        if (dtype == DOPT_B200_FLOAT32) go<float>(in[0], out, s);
        else go<int>(in[0], out, s);
    }
};

struct ArgminKernel : Kernel {
    int dtype;
    int64_t outer = 1, A = 1, inner = 1;
    ArgminKernel(const dopt_b200_op& d) {
        const auto& in = d.inputs[0];
        DB_REQUIRE(d.n_inputs == 1 && d.axis >= 0 && d.axis < in.rank, "argmin: axis out of range");
        dtype = in.dtype;
        for (int i = 0; i < in.rank; ++i) {
            if (i < d.axis) outer *= in.shape[i];
            else if (i > d.axis) inner *= in.shape[i];
            else A = in.shape[i];
        }
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "argmin: one input");
        int grid = stream_grid(outer * inner, 256, 8);
        if (dtype == DOPT_B200_FLOAT32)
            argmin_kernel<float><<<grid, 256, 0, s>>>((const float*)in[0], (int*)out, outer, A, inner, FLT_MAX);
        else
            argmin_kernel<int><<<grid, 256, 0, s>>>((const int*)in[0], (int*)out, outer, A, inner, INT_MAX);
        DB_LAUNCH_CHECK();
    }
};

Kernel* make_sum(const dopt_b200_op& d) { return new ReduceKernel(d, false); }
Kernel* make_max(const dopt_b200_op& d) { return new ReduceKernel(d, true); }
Kernel* make_argmin(const dopt_b200_op& d) { return new ArgminKernel(d); }
}  // namespace

// flat sum of n floats into out[0] (used by the plan for loss / weight-decay terms)
void sum_flat_launch(const float* in, float* out, int64_t n, Scratch& ws, cudaStream_t s) {
    reduce_axis<float, RedSum<float>>(in, out, 1, n, 1, ws, s);
}

void register_reduce() {
    register_kernel("sum", make_sum);
    register_kernel("maxElement", make_max);
    register_kernel("argmin", make_argmin);
}

}  // namespace db
