// msum.cu -- many full reductions in two launches.
//
// dopt's weight decay is `wd * sum(W * W)` per parameter tensor (nnet/layers/conv.d:100-108, dense.d:96-104): for a WRN that
// is 29 products and 29 `sum` nodes, each a two-pass reduction -- 56 launches moving 4 words per parameter.  The plan batches
// every full `sum` (and the product feeding it, when nothing else reads the product) that is ready at the same time into
// ONE pair of launches: row r computes sum_i a_r[i] * b_r[i]  (b_r == nullptr: sum_i a_r[i]).  One read per operand.
// Deterministic: fixed chunking, per-chunk partials combined in chunk order.  Same products (`__fmul_rn`) as the
// stand-alone `mul`; the summation order differs from reduce_rows', within the 1e-4 tolerance stated for fp32 reductions.
#include "common.cuh"

namespace db {

static constexpr int kMsChunk = 16384;   // elements per CTA

__global__ void __launch_bounds__(256) msum_partial_kernel(const MsumRow* __restrict__ rows, int n_rows,
                                                           float* __restrict__ partial) {
    int lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (rows[mid].chunk0 <= (int64_t)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const MsumRow r = rows[lo];
    const int64_t base = ((int64_t)blockIdx.x - r.chunk0) * kMsChunk;
    const int64_t end = base + kMsChunk < r.n ? base + kMsChunk : r.n;
    float acc = 0.f;
    const bool vec = (((uintptr_t)r.a | (uintptr_t)r.b) & 15) == 0;
    int64_t i = base + (int64_t)threadIdx.x * 4;
    if (vec) {
        for (; i + 3 < end; i += 256 * 4) {
            float4 a = dbk::ld_stream((const float4*)(r.a + i));
            if (r.b) {
                float4 b = dbk::ld_stream((const float4*)(r.b + i));
                acc += (__fmul_rn(a.x, b.x) + __fmul_rn(a.y, b.y)) + (__fmul_rn(a.z, b.z) + __fmul_rn(a.w, b.w));
            } else {
                acc += (a.x + a.y) + (a.z + a.w);
            }
        }
        // tail of the chunk (fewer than 4 elements left for this thread)
        for (int64_t t = i; t < end && t < i + 4; ++t) acc += r.b ? __fmul_rn(r.a[t], r.b[t]) : r.a[t];
    } else {
        for (int64_t t = base + threadIdx.x; t < end; t += 256) acc += r.b ? __fmul_rn(r.a[t], r.b[t]) : r.a[t];
    }
    __shared__ float sm[8];
    acc = dbk::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += sm[w];
        partial[blockIdx.x] = v;
    }
}

__global__ void __launch_bounds__(32) msum_finish_kernel(const MsumRow* __restrict__ rows, const float* __restrict__ partial) {
    const MsumRow r = rows[blockIdx.x];
    const int64_t chunks = (r.n + kMsChunk - 1) / kMsChunk;
    // lane l adds chunks l, l+32, ... in order; lanes are combined by a fixed butterfly
    float acc = 0.f;
    for (int64_t c = threadIdx.x; c < chunks; c += 32) acc += partial[r.chunk0 + c];
    acc = dbk::warp_sum(acc);
    if (threadIdx.x == 0) *r.out = acc;
}

int64_t msum_layout(MsumRow* rows, int n) {
    int64_t chunks = 0;
    for (int i = 0; i < n; ++i) {
        rows[i].chunk0 = chunks;
        chunks += std::max<int64_t>(1, ceil_div(rows[i].n, (int64_t)kMsChunk));
    }
    return chunks;
}

void msum_launch(const MsumRow* dev_rows, int n, int64_t chunks, float* partial, cudaStream_t s) {
    if (n <= 0) return;
    msum_partial_kernel<<<(unsigned)chunks, 256, 0, s>>>(dev_rows, n, partial);
    DB_LAUNCH_CHECK();
    msum_finish_kernel<<<(unsigned)n, 32, 0, s>>>(dev_rows, partial);
    DB_LAUNCH_CHECK();
}

}  // namespace db
