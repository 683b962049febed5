// input.cu -- on-device input pipeline (the callers' side of the hot path).
//
// The reference prepares every batch on the host, single-threaded, in three passes over the pixels:
//   * the loaders turn the dataset's bytes into floats, `x / 128.0f - 1.0f` (nnet/source/dopt/nnet/data/cifar.d:50) and the
//     label byte into a one-hot row (cifar.d:52-55);
//   * ImageTransformer.getBatch (nnet/source/dopt/nnet/data/imagetransformer.d:45-138) pads each image with its own
//     reflection (jitterX / jitterY pixels), crops a window at a random offset, then mirrors it horizontally and / or
//     vertically with probability 1/2 each;
//   * CUDAPlan copies the float batch to the device (cuda/source/dopt/cuda/package.d:373-381).
// Here the batch crosses PCIe as BYTES (4x less) and ONE kernel does normalise + reflect-pad + crop + flip straight into the
// NCHW fp32 tensor the plan reads: the padded intermediate never exists, every output pixel is a closed-form gather
//     out[c, y, x] = src[c, R_H(fy(y) + yOff - jy), R_W(fx(x) + xOff - jx)]
// with R_n(i) = i < 0 ? -1 - i : (i >= n ? 2n - 1 - i : i) (the edge pixel is repeated, as the reference's slice-reverse
// does, imagetransformer.d:77-99), fx(x) = flipX ? W-1-x : x, fy likewise (the reference flips AFTER cropping, :118-137).
// Bit-exact with the host loops: the only arithmetic is the exact u8 -> float conversion.
//
// Memory-bound, write-dominated: 1 B read + 4 B written per element (5 B/elem, u8 source), 8 B/elem for a float source.
// One thread produces a 4 x 4 block of output pixels, four 128-bit stores; the byte gathers hit L1/L2 (a 32x32x3 image is 3 KB).
//
// The per-image decisions (xOff, yOff, flipX, flipY) are an int4 table in device memory, so a test can supply exactly the
// draws of a host run; dopt_b200_jitter_sample fills it on the device (Philox-4x32-10, one counter block per image) with the
// reference's distributions: offsets uniform on [0, 2*jitter), flips Bernoulli(1/2) (imagetransformer.d:101-102,118,126).
#include "common.cuh"

namespace db {

__device__ __forceinline__ int reflect_index(int i, int n) { return i < 0 ? -1 - i : (i >= n ? 2 * n - 1 - i : i); }

// x / 128.0f - 1.0f: dividing by a power of two is an exact scaling, so the multiplication gives the same bits as the division
__device__ __forceinline__ float load_pixel(const uint8_t* p) { return __fsub_rn(__fmul_rn((float)*p, 0.0078125f), 1.0f); }
__device__ __forceinline__ float load_pixel(const float* p) { return *p; }

template <typename T>
__global__ void __launch_bounds__(256) image_transform_kernel(const T* __restrict__ src, float* __restrict__ dst, int64_t n_img,
                                                              int C, int H, int W, int jx, int jy,
                                                              const dopt_b200_jitter* __restrict__ jit) {
    // Work item = a 4 x 4 block of output pixels (four 128-bit stores): the index arithmetic, the per-image draws and the four
    // reflected source columns are computed once per block.  All of it in 32 bits -- with one item per 128-bit store and three
    // 64-bit divisions each the kernel was bound by the integer pipe (1.7 TB/s over a CIFAR-sized training set, then 2.3 TB/s
    // with 32-bit divisions: bench.py adjacent_rows) instead of by the 5 B it moves per element.
    const int wq = (W + 3) >> 2, hq = (H + 3) >> 2;
    const unsigned per_img = (unsigned)(C * hq * wq);   // blocks per image (the host checks that it fits)
    const int64_t items = n_img * (int64_t)per_img;
    const bool vec = (W & 3) == 0 && ((uintptr_t)dst & 15) == 0;
    const bool small = items < ((int64_t)1 << 32);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < items; q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t img = small ? (int64_t)((unsigned)q / per_img) : q / per_img;
        unsigned r = (unsigned)(q - img * per_img);
        const int xq = (int)(r % (unsigned)wq);
        r /= (unsigned)wq;
        const int yq = (int)(r % (unsigned)hq);
        const int c = (int)(r / (unsigned)hq);
        int xo = jx, yo = jy, fx = 0, fy = 0;          // no table: centre crop, no flip == the identity
        if (jit) {
            const dopt_b200_jitter j = jit[img];
            xo = j.x_off; yo = j.y_off; fx = j.flip_x; fy = j.flip_y;
        }
        int sx[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int x = xq * 4 + k;
            sx[k] = x < W ? reflect_index((fx ? W - 1 - x : x) + xo - jx, W) : 0;
        }
        const int64_t plane = (img * C + c) * (int64_t)H * W;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int y = yq * 4 + j;
            if (y < H) {
                const int sy = reflect_index((fy ? H - 1 - y : y) + yo - jy, H);
                const T* row = src + plane + sy * W;
                float v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) v[k] = (xq * 4 + k < W) ? load_pixel(row + sx[k]) : 0.f;
                float* out = dst + plane + y * W + xq * 4;
                if (vec) {
                    dbk::st_stream((float4*)out, make_float4(v[0], v[1], v[2], v[3]));
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        if (xq * 4 + k < W) out[k] = v[k];
                }
            }
        }
    }
}

// one-hot rows from label bytes: `ls[] = 0; ls[tmp[labelIdx]] = 1.0f` (cifar.d:52-55).  4 B written per element.
__global__ void __launch_bounds__(256) one_hot_kernel(const uint8_t* __restrict__ labels, float* __restrict__ dst, int64_t n,
                                                      int classes) {
    const int64_t total = n * classes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = (int)(i % classes) == (int)labels[i / classes] ? 1.0f : 0.0f;
}

__device__ __forceinline__ void philox4(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
        uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
        uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
        uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

__global__ void __launch_bounds__(256) jitter_sample_kernel(dopt_b200_jitter* __restrict__ out, int64_t n, int jx, int jy,
                                                            int flip_x, int flip_y, uint64_t seed, uint64_t call) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t c[4] = {(uint32_t)i, (uint32_t)(i >> 32), (uint32_t)call, (uint32_t)(call >> 32)};
        philox4(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        dopt_b200_jitter j;
        // uniform integer on [0, 2*jitter): multiply-high maps 32 random bits onto the range without a division
        j.x_off = jx > 0 ? (int)__umulhi(c[0], (uint32_t)(2 * jx)) : 0;
        j.y_off = jy > 0 ? (int)__umulhi(c[1], (uint32_t)(2 * jy)) : 0;
        j.flip_x = flip_x ? (int)(c[2] >> 31) : 0;
        j.flip_y = flip_y ? (int)(c[3] >> 31) : 0;
        out[i] = j;
    }
}

template <typename T>
static void image_transform(const T* src, float* dst, int64_t n, int c, int h, int w, int jx, int jy,
                            const dopt_b200_jitter* jit, cudaStream_t s) {
    DB_REQUIRE(src && dst, "image_transform: null buffer");
    DB_REQUIRE(n >= 0 && c > 0 && h > 0 && w > 0, "image_transform: bad shape");
    // the reference's slice arithmetic needs the reflected border to fit inside the image (imagetransformer.d:77-99)
    DB_REQUIRE(jx >= 0 && jy >= 0 && jx <= w && jy <= h, "image_transform: jitter larger than the image");
    if (n == 0) return;
    DB_REQUIRE((int64_t)c * h * w < ((int64_t)1 << 31), "image_transform: image too large");
    const int64_t items = n * c * (int64_t)((h + 3) / 4) * ((w + 3) / 4);
    image_transform_kernel<T><<<stream_grid(items, 256, 8), 256, 0, s>>>(src, dst, n, c, h, w, jx, jy, jit);
    DB_LAUNCH_CHECK();
}

}  // namespace db

extern "C" {

#define DB_TRY try { db::require_device();
#define DB_END                                  \
    }                                           \
    catch (const std::exception& e) {           \
        db::set_last_error(e.what());           \
        return 1;                               \
    }                                           \
    catch (...) {                               \
        db::set_last_error("unknown error");    \
        return 1;                               \
    }                                           \
    return 0;

int dopt_b200_image_transform_u8(const uint8_t* src, float* dst, int64_t n, int c, int h, int w, int jitter_x, int jitter_y,
                                 const dopt_b200_jitter* per_image, void* stream) {
    DB_TRY
    db::image_transform<uint8_t>(src, dst, n, c, h, w, jitter_x, jitter_y, per_image, (cudaStream_t)stream);
    DB_END
}

int dopt_b200_image_transform_f32(const float* src, float* dst, int64_t n, int c, int h, int w, int jitter_x, int jitter_y,
                                  const dopt_b200_jitter* per_image, void* stream) {
    DB_TRY
    DB_REQUIRE(src != dst, "image_transform_f32: in-place is not supported (the gather reads pixels other threads write)");
    db::image_transform<float>(src, dst, n, c, h, w, jitter_x, jitter_y, per_image, (cudaStream_t)stream);
    DB_END
}

int dopt_b200_one_hot_u8(const uint8_t* labels, float* dst, int64_t n, int classes, void* stream) {
    DB_TRY
    DB_REQUIRE(labels && dst && n >= 0 && classes > 0, "one_hot: bad argument");
    if (n > 0) {
        db::one_hot_kernel<<<db::stream_grid(n * classes, 256, 8), 256, 0, (cudaStream_t)stream>>>(labels, dst, n, classes);
        DB_LAUNCH_CHECK();
    }
    DB_END
}

int dopt_b200_jitter_sample(dopt_b200_jitter* out, int64_t n, int jitter_x, int jitter_y, int flip_x, int flip_y,
                            uint64_t seed, uint64_t call, void* stream) {
    DB_TRY
    DB_REQUIRE(out && n >= 0 && jitter_x >= 0 && jitter_y >= 0, "jitter_sample: bad argument");
    if (n > 0) {
        db::jitter_sample_kernel<<<db::stream_grid(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(out, n, jitter_x, jitter_y,
                                                                                               flip_x, flip_y, seed, call);
        DB_LAUNCH_CHECK();
    }
    DB_END
}

}  // extern "C"
