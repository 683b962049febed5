// basic.cu -- slice / pad / repeat / transpose (cuda/source/dopt/cuda/basic.d).
//
// Reference: slice and pad recurse down to one cuMemcpy per innermost row (basic.d:31-141); repeat runs one
// byte-granular NVRTC kernel per repeated axis through freshly allocated temporaries (basic.d:143-217); transpose is
// cublasSgeam (basic.d:219-247).  All element types are 4 bytes.
// Here: one index-mapping gather kernel for slice/pad/repeat (output-stationary, coalesced writes), a straight
// cudaMemcpyAsync when the region is contiguous (the batch-norm pack/unpack slices are), and a 32x32 shared-memory
// tile transpose.  All HBM-bound: 2 * volume(out) * 4 B.
#include "common.cuh"

namespace db {

struct MapParams {
    int rank;
    int64_t out_shape[DOPT_B200_MAX_RANK];
    int64_t in_shape[DOPT_B200_MAX_RANK];
    int64_t in_stride[DOPT_B200_MAX_RANK];
    int64_t delta[DOPT_B200_MAX_RANK];   // src coord = out coord + delta (slice: +start, pad: -before)
};

// MODE 0: shifted window with zero fill outside the source (slice and pad); MODE 1: modulo (repeat)
template <int MODE>
__global__ void __launch_bounds__(256) map_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                  int64_t n, MapParams p) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += stride) {
        int64_t rem = idx, src = 0;
        bool ok = true;
#pragma unroll 1
        for (int d = p.rank - 1; d >= 0; --d) {
            int64_t c = rem % p.out_shape[d];
            rem /= p.out_shape[d];
            if (MODE == 0) {
                c += p.delta[d];
                ok = ok && c >= 0 && c < p.in_shape[d];
            } else {
                c %= p.in_shape[d];
            }
            src += c * p.in_stride[d];
        }
        out[idx] = ok ? in[src] : 0u;
    }
}

__global__ void __launch_bounds__(256) transpose_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                        int rows_in, int cols_in) {
    // in: [rows_in][cols_in] -> out: [cols_in][rows_in]
    // the row tiles are on gridDim.x (2^31 - 1 tiles), the column tiles on gridDim.y and looped over when there are more than 65535
    __shared__ uint32_t tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int by = blockIdx.x * 32;
    for (int bx = blockIdx.y * 32; bx < cols_in; bx += gridDim.y * 32) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            int r = by + ty + j, c = bx + tx;
            if (r < rows_in && c < cols_in) tile[ty + j][tx] = in[(int64_t)r * cols_in + c];
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
            int r = bx + ty + j, c = by + tx;   // out row = in col
            if (r < cols_in && c < rows_in) out[(int64_t)r * rows_in + c] = tile[tx][ty + j];
        }
        __syncthreads();
    }
}

void transpose2d_launch(const void* in, void* out, int rows_in, int cols_in, cudaStream_t s) {
    if (rows_in <= 0 || cols_in <= 0) return;
    dim3 grid((unsigned)ceil_div(rows_in, 32), (unsigned)std::min<int64_t>(ceil_div(cols_in, 32), 65535));
    transpose_kernel<<<grid, 256, 0, s>>>((const uint32_t*)in, (uint32_t*)out, rows_in, cols_in);
    DB_LAUNCH_CHECK();
}

namespace {

static void fill_strides(const dopt_b200_tensor& t, int64_t* st) {
    int64_t s = 1;
    for (int d = t.rank - 1; d >= 0; --d) {
        st[d] = s;
        s *= t.shape[d];
    }
}

struct SliceKernel : Kernel {
    MapParams p{};
    int64_t n, contiguous_offset = -1;
    SliceKernel(const dopt_b200_op& d) {
        const auto& in = d.inputs[0];
        DB_REQUIRE(d.n_inputs == 1 && in.rank == d.output.rank && in.rank <= DOPT_B200_MAX_RANK, "slice: bad operands");
        p.rank = in.rank;
        n = volume(d.output);
        fill_strides(in, p.in_stride);
        for (int i = 0; i < in.rank; ++i) {
            // verifier of the reference (core/source/dopt/core/ops/basic.d:35-58)
            DB_REQUIRE(d.start[i] >= 0 && d.start[i] < d.stop[i] && d.stop[i] <= in.shape[i], "slice: bad range");
            DB_REQUIRE(d.output.shape[i] == d.stop[i] - d.start[i], "slice: output shape mismatch");
            p.out_shape[i] = d.output.shape[i];
            p.in_shape[i] = in.shape[i];
            p.delta[i] = d.start[i];
        }
        // contiguous when only the outermost non-unit dimension is cut
        bool contig = true;
        for (int i = in.rank - 1; i >= 0; --i) {
            bool full = (d.output.shape[i] == in.shape[i]);
            if (!full) {
                // every dimension outside (to the left of) this one must have extent 1 in the output
                for (int j = 0; j < i; ++j)
                    if (d.output.shape[j] != 1) contig = false;
                break;
            }
        }
        if (contig) {
            int64_t off = 0;
            for (int i = 0; i < in.rank; ++i) off += d.start[i] * p.in_stride[i];
            contiguous_offset = off;
        }
        if (in.rank == 0) contiguous_offset = 0;
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "slice: one input");
        if (n == 0) return;
        if (contiguous_offset >= 0) {
            DB_CUDA(cudaMemcpyAsync(out, (const uint32_t*)in[0] + contiguous_offset, (size_t)n * 4,
                                    cudaMemcpyDeviceToDevice, s));
            count_launch();
            return;
        }
        map_kernel<0><<<stream_grid(n, 256, 16), 256, 0, s>>>((const uint32_t*)in[0], (uint32_t*)out, n, p);
        DB_LAUNCH_CHECK();
    }
};

struct PadKernel : Kernel {
    MapParams p{};
    int64_t n;
    PadKernel(const dopt_b200_op& d) {
        const auto& in = d.inputs[0];
        DB_REQUIRE(d.n_inputs == 1 && in.rank == d.output.rank && in.rank <= DOPT_B200_MAX_RANK, "pad: bad operands");
        p.rank = in.rank;
        n = volume(d.output);
        fill_strides(in, p.in_stride);
        for (int i = 0; i < in.rank; ++i) {
            DB_REQUIRE(d.before[i] >= 0 && d.after[i] >= 0, "pad: negative padding");
            DB_REQUIRE(d.output.shape[i] == in.shape[i] + d.before[i] + d.after[i], "pad: output shape mismatch");
            p.out_shape[i] = d.output.shape[i];
            p.in_shape[i] = in.shape[i];
            p.delta[i] = -d.before[i];
        }
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "pad: one input");
        if (n == 0) return;
        map_kernel<0><<<stream_grid(n, 256, 16), 256, 0, s>>>((const uint32_t*)in[0], (uint32_t*)out, n, p);
        DB_LAUNCH_CHECK();
    }
};

struct RepeatKernel : Kernel {
    MapParams p{};
    int64_t n;
    RepeatKernel(const dopt_b200_op& d) {
        const auto& in = d.inputs[0];
        DB_REQUIRE(d.n_inputs == 1 && in.rank == d.output.rank && in.rank <= DOPT_B200_MAX_RANK, "repeat: bad operands");
        p.rank = in.rank;
        n = volume(d.output);
        fill_strides(in, p.in_stride);
        for (int i = 0; i < in.rank; ++i) {
            DB_REQUIRE(d.repetitions[i] > 0, "repeat: repetitions must be positive");
            DB_REQUIRE(d.output.shape[i] == in.shape[i] * d.repetitions[i], "repeat: output shape mismatch");
            p.out_shape[i] = d.output.shape[i];
            p.in_shape[i] = in.shape[i];
        }
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "repeat: one input");
        if (n == 0) return;
        map_kernel<1><<<stream_grid(n, 256, 16), 256, 0, s>>>((const uint32_t*)in[0], (uint32_t*)out, n, p);
        DB_LAUNCH_CHECK();
    }
};

struct TransposeKernel : Kernel {
    int rank;
    int64_t n, rows_in = 1, cols_in = 1;
    bool swap;
    TransposeKernel(const dopt_b200_op& d) {
        const auto& in = d.inputs[0];
        rank = in.rank;
        // "Currently only implemented for rank 2 tensors" (core/source/dopt/core/ops/basic.d:296)
        DB_REQUIRE(d.n_inputs == 1 && rank <= 2, "transpose is only implemented for rank <= 2");
        n = volume(d.output);
        swap = (rank == 2 && d.order[0] == 1 && d.order[1] == 0);
        if (rank == 2) {
            rows_in = in.shape[0];
            cols_in = in.shape[1];
            DB_REQUIRE((d.order[0] == 0 && d.order[1] == 1) || swap, "transpose: order must be a permutation of [0,1]");
        }
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 1, "transpose: one input");
        if (n == 0) return;
        if (!swap) {
            DB_CUDA(cudaMemcpyAsync(out, in[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
            count_launch();
            return;
        }
        transpose2d_launch(in[0], out, (int)rows_in, (int)cols_in, s);
    }
};

template <class K> Kernel* make(const dopt_b200_op& d) { return new K(d); }
}  // namespace

void register_basic() {
    register_kernel("slice", make<SliceKernel>);
    register_kernel("pad", make<PadKernel>);
    register_kernel("repeat", make<RepeatKernel>);
    register_kernel("transpose", make<TransposeKernel>);
}

}  // namespace db
