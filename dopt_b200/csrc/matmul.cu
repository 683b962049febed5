// matmul.cu -- rank-2 matmul (cuda/source/dopt/cuda/math.d:214-247: cublasSgemm_v2, row-major via swapped operands,
// alpha = 1, beta = 0; op definition core/source/dopt/core/ops/math.d:129-142).
//
// In dopt's graphs most matmuls are NOT dense contractions: scalar broadcasts are lowered to [V,1]x[1,1], bias
// broadcasts to [N,1]x[1,out], row / column sums to a product with a ones vector (core/source/dopt/core/ops/package.d:96-103,
// core/source/dopt/core/ops/basic.d:370-381, core/source/dopt/core/ops/math.d:257-279).  Those shapes are HBM-bound and get their own
// kernels here:
//   K == 1   outer product   C[m,n] = A[m]*B[n]              bytes: (M + N + M*N) * 4   (bit-exact: one multiply)
//   N == 1   row dots        C[m]   = sum_k A[m,k]*B[k]      bytes: (M*K + K + M) * 4
//   M == 1   column sums     C[n]   = sum_k A[k]*B[k,n]      bytes: (K + K*N + N) * 4
// Everything else goes to a 64x64x16 register-tiled fp32 SIMT kernel (MATH_FP32) or, for shapes that fill a
// 128-row tile, to the tcgen05 bf16 GEMM in tc_gemm.cu (MATH_BF16).
#include "common.cuh"
#include "tc.cuh"

namespace db {

__global__ void __launch_bounds__(256) outer_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                    float* __restrict__ c, int64_t M, int64_t N) {
    int64_t n = M * N;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t m = i / N, j = i - m * N;
        c[i] = __fmul_rn(a[m], b[j]);
    }
}
// N == 1 && K == 1 fast path (scalar broadcast [V,1]x[1,1]): vectorised
__global__ void __launch_bounds__(256) scale_bcast_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                          float* __restrict__ c, int64_t M) {
    float s = b[0];
    int64_t nv = M >> 2;
    const float4* av = (const float4*)a;
    float4* cv = (float4*)c;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
        float4 v = av[i];
        v.x = __fmul_rn(v.x, s); v.y = __fmul_rn(v.y, s); v.z = __fmul_rn(v.z, s); v.w = __fmul_rn(v.w, s);
        dbk::st_stream(cv + i, v);
    }
    int64_t t = (nv << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < M) c[t] = __fmul_rn(a[t], s);
}

// one warp per row
__global__ void __launch_bounds__(256) rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ c, int64_t M, int64_t K) {
    int lane = threadIdx.x & 31;
    int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t m = warp; m < M; m += nwarps) {
        const float* row = a + m * K;
        float acc = 0.f;
        for (int64_t k = lane; k < K; k += 32) acc = fmaf(row[k], b[k], acc);
        acc = dbk::warp_sum(acc);
        if (lane == 0) c[m] = acc;
    }
}

// thread per output column, K split over blockIdx.y with a second pass when K is long
__global__ void __launch_bounds__(256) colsum_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                     float* __restrict__ c, int64_t K, int64_t N, int64_t kchunk) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    int64_t k0 = (int64_t)blockIdx.y * kchunk, k1 = k0 + kchunk < K ? k0 + kchunk : K;
    float acc = 0.f;
    for (int64_t k = k0; k < k1; ++k) acc = fmaf(a[k], b[k * N + n], acc);
    c[(int64_t)blockIdx.y * N + n] = acc;
}
__global__ void __launch_bounds__(256) colsum_final(const float* __restrict__ part, float* __restrict__ c, int64_t N,
                                                    int parts) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float acc = 0.f;
    for (int p = 0; p < parts; ++p) acc += part[(int64_t)p * N + n];
    c[n] = acc;
}

// general fp32 GEMM, C = A(MxK) * B(KxN), all row-major.  64x64 CTA tile, 16-deep k slab, 4x4 outputs per thread.
// blockIdx.z selects a k range of `kchunk` (split-K): slice z writes its partial product to C + z * M * N
__global__ void __launch_bounds__(256) sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                    float* __restrict__ C, int M, int N, int K, int kchunk) {
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    int tid = threadIdx.x;
    int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64;   // row tiles on gridDim.x: tall operands (M in the millions) are common
    int tr = (tid >> 4) * 4, tc = (tid & 15) * 4;
    float acc[4][4] = {};
    const int kbeg = blockIdx.z * kchunk;
    const int kend = min(K, kbeg + kchunk);
    C += (int64_t)blockIdx.z * M * N;
    for (int k0 = kbeg; k0 < kend; k0 += 16) {
        // A tile: 64 rows x 16 k -> As[k][m]; 1024 elements, 4 per thread
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int r = e >> 4, kk = e & 15;
            int gm = m0 + r, gk = k0 + kk;
            As[kk][r] = (gm < M && gk < kend) ? A[(int64_t)gm * K + gk] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int e = tid + i * 256;
            int kk = e >> 6, c = e & 63;
            int gk = k0 + kk, gn = n0 + c;
            Bs[kk][c] = (gk < kend && gn < N) ? B[(int64_t)gk * N + gn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][tr + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tc + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int gm = m0 + tr + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tc + j;
            if (gn < N) C[(int64_t)gm * N + gn] = acc[i][j];
        }
    }
}

// sum of `parts` partial products (split-K), in slice order: deterministic
__global__ void __launch_bounds__(256) sgemm_reduce_kernel(const float* __restrict__ part, float* __restrict__ c, int64_t n,
                                                           int parts) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < parts; ++p) acc += part[(int64_t)p * n + i];
        c[i] = acc;
    }
}

// how many k slices a small-output GEMM is split into (1 = no split): few 64x64 output tiles and a long reduction would leave
// most SMs idle behind a serial k loop (the dense layer of the WRN configs: 4 tiles x 40 k-slabs took 22 us)
int sgemm_splits(int64_t M, int64_t N, int64_t K) {
    const int64_t tiles = ceil_div(N, 64) * ceil_div(M, 64);
    if (tiles * 2 > sm_count() || K < 64) return 1;
    int64_t s = std::min<int64_t>(ceil_div(K, 32), std::max<int64_t>(1, sm_count() / tiles));
    return (int)std::max<int64_t>(1, s);
}

void sgemm_launch(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, cudaStream_t s, float* ws = nullptr,
                  int splits = 1) {
    if (splits > 1 && ws) {
        int kchunk = (int)(ceil_div(ceil_div(K, splits), 16) * 16);
        splits = (int)ceil_div(K, kchunk);
        dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(N, 64), (unsigned)splits);
        sgemm_kernel<<<grid, 256, 0, s>>>(A, B, ws, (int)M, (int)N, (int)K, kchunk);
        DB_LAUNCH_CHECK();
        sgemm_reduce_kernel<<<stream_grid(M * N, 256, 4), 256, 0, s>>>(ws, C, M * N, splits);
        DB_LAUNCH_CHECK();
        return;
    }
    DB_REQUIRE(ceil_div(N, 64) <= 65535, "matmul: more than 4.19 M result columns are not supported by the fp32 kernel");
    dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(N, 64));
    sgemm_kernel<<<grid, 256, 0, s>>>(A, B, C, (int)M, (int)N, (int)K, (int)K);
    DB_LAUNCH_CHECK();
}

namespace {
struct MatmulKernel : Kernel {
    int64_t M, K, N;
    int math;
    Scratch ws;
    TcGemm* tc = nullptr;
    MatmulKernel(const dopt_b200_op& d) {
        // verifier: core/source/dopt/core/ops/math.d:129-136
        DB_REQUIRE(d.n_inputs == 2 && d.inputs[0].rank == 2 && d.inputs[1].rank == 2, "matmul: rank-2 operands required");
        DB_REQUIRE(d.inputs[0].dtype == DOPT_B200_FLOAT32 && d.inputs[1].dtype == DOPT_B200_FLOAT32,
                   "Element type not supported.");   // math.d:243
        M = d.inputs[0].shape[0];
        K = d.inputs[0].shape[1];
        N = d.inputs[1].shape[1];
        DB_REQUIRE(d.inputs[1].shape[0] == K, "matmul: inner dimensions differ");
        DB_REQUIRE(d.output.rank == 2 && d.output.shape[0] == M && d.output.shape[1] == N, "matmul: bad output shape");
        math = resolve_math(d.math);
        if (math == DOPT_B200_MATH_BF16 && tc_gemm_supported(M, N, K)) tc = tc_gemm_create(M, N, K);
    }
    ~MatmulKernel() { tc_gemm_destroy(tc); }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 2, "matmul: two inputs");
        const float* a = (const float*)in[0];
        const float* b = (const float*)in[1];
        float* c = (float*)out;
        if (M * N == 0) return;
        if (K == 1 && N == 1 && ((uintptr_t)a % 16 == 0) && ((uintptr_t)c % 16 == 0)) {
            scale_bcast_kernel<<<stream_grid(ceil_div(M, 4), 256, 8), 256, 0, s>>>(a, b, c, M);
            DB_LAUNCH_CHECK();
        } else if (K == 1) {
            outer_kernel<<<stream_grid(M * N, 256, 8), 256, 0, s>>>(a, b, c, M, N);
            DB_LAUNCH_CHECK();
        } else if (N == 1) {
            rowdot_kernel<<<stream_grid(M * 32, 256, 8), 256, 0, s>>>(a, b, c, M, K);
            DB_LAUNCH_CHECK();
        } else if (M == 1) {
            int parts = (int)std::min<int64_t>(ceil_div(K, 256), 64);
            if (ceil_div(N, 256) * parts < sm_count() / 2 && K < 4096) parts = (int)std::min<int64_t>(parts, 8);
            int64_t kchunk = ceil_div(K, parts);
            parts = (int)ceil_div(K, kchunk);
            dim3 grid((unsigned)ceil_div(N, 256), (unsigned)parts);
            if (parts == 1) {
                colsum_kernel<<<grid, 256, 0, s>>>(a, b, c, K, N, kchunk);
                DB_LAUNCH_CHECK();
            } else {
                float* part = (float*)ws.get((size_t)parts * N * sizeof(float));
                colsum_kernel<<<grid, 256, 0, s>>>(a, b, part, K, N, kchunk);
                DB_LAUNCH_CHECK();
                colsum_final<<<(unsigned)ceil_div(N, 256), 256, 0, s>>>(part, c, N, parts);
                DB_LAUNCH_CHECK();
            }
        } else if (tc) {
            tc_gemm_run(tc, a, b, c, s);
        } else {
            const int splits = sgemm_splits(M, N, K);
            float* part = splits > 1 ? (float*)ws.get((size_t)splits * M * N * sizeof(float)) : nullptr;
            sgemm_launch(a, b, c, M, N, K, s, part, splits);
        }
    }
};
Kernel* make_matmul(const dopt_b200_op& d) { return new MatmulKernel(d); }
}  // namespace

void register_matmul() { register_kernel("matmul", make_matmul); }

}  // namespace db
