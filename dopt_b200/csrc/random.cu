// random.cu -- `uniform` (cuda/source/dopt/cuda/random.d:56-83: curandGenerateUniform, default generator, seeded from
// unpredictableSeed).  cuRAND's uniform is (0, 1]; so is this one.  The reference is unseeded, so parity is at the level of
// the distribution, not the stream.  Counter-based Philox-4x32-10: thread i produces elements 4i..4i+3.  Write-only:
// V*4 B per launch.  The per-kernel call counter (so that every execution draws fresh numbers) lives in DEVICE memory and
// is advanced by a one-thread launch that follows the generator on the same stream: a plan that contains `uniform`
// (dropout) can therefore be captured in a CUDA graph and still produces a new mask on every replay.
#include "common.cuh"
#include <random>

namespace db {

__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

__global__ void __launch_bounds__(256) uniform_kernel(float* __restrict__ out, int64_t n, uint64_t seed,
                                                      const uint64_t* __restrict__ calls) {
    const uint64_t call = *calls;
    int64_t nq = (n + 3) >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
        uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            philox_round(c, k0, k1);
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int64_t e = (i << 2) + j;
            // 24 random bits -> (0, 1]
            if (e < n) out[e] = (float)((c[j] >> 8) + 1u) * (1.0f / 16777216.0f);
        }
    }
}

__global__ void uniform_advance_kernel(uint64_t* calls) { *calls += 1; }

namespace {
struct UniformKernel : Kernel {
    int64_t n;
    uint64_t seed;
    uint64_t* calls = nullptr;   // device-resident execution counter
    ~UniformKernel() override { if (calls) cudaFree(calls); }
    UniformKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 0 && d.output.dtype == DOPT_B200_FLOAT32, "uniform: no operands, float32 result");
        n = volume(d.output);
        seed = d.seed;
        if (seed == 0) {   // std.random.unpredictableSeed in the reference (random.d:66)
            std::random_device rd;
            seed = ((uint64_t)rd() << 32) | rd();
        }
        DB_CUDA(cudaMalloc(&calls, sizeof(uint64_t)));
        DB_CUDA(cudaMemset(calls, 0, sizeof(uint64_t)));
    }
    void run(const void* const*, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 0, "uniform: no inputs");
        if (n == 0) return;
        uniform_kernel<<<stream_grid(ceil_div(n, 4), 256, 8), 256, 0, s>>>((float*)out, n, seed, calls);
        DB_LAUNCH_CHECK();
        uniform_advance_kernel<<<1, 1, 0, s>>>(calls);
        DB_LAUNCH_CHECK();
    }
};
Kernel* make_uniform(const dopt_b200_op& d) { return new UniformKernel(d); }
}  // namespace

void register_random() { register_kernel("uniform", make_uniform); }

}  // namespace db
