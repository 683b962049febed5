// tc_host.cu -- host side of the tcgen05 path: tensor maps, operand staging (fp32 NCHW/KCRS -> bf16 NHWC/packed),
// tile selection and launches for matmul, convolution, convolutionFeaturesGrad, convolutionFiltersGrad.
//
// Data layout in HBM for one convolution call (per-op drop-in mode; the boundary stays NCHW / KCRS fp32 like the
// reference, cuda/source/dopt/cuda/nnet/cudnn7.d:83-90):
//   activations  NCHW fp32  --nchw_to_nhwc_bf16-->  [N][H][W][Cp] bf16   (Cp = C rounded up to 8: TMA needs 16-byte strides)
//   filters      KCRS fp32  --pack_filters-------->  fwd  : [K ][R*S*Cp] bf16, tap-major, FLIPPED (true convolution, survey F1)
//                                                     dgrad: [C ][R*S*Kp] bf16
//   results      fwd / dgrad: the epilogue writes NCHW fp32 directly (32 lanes = 32 consecutive pixels of one channel)
//                wgrad      : fp32 atomics into KCRS (split over the pixel dimension)
#include "common.cuh"
#include "tc.cuh"
#include "flat.cuh"
#include <type_traits>
#include "tc_kernel.cuh"
#include <set>
#include <mutex>

namespace db {

// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess &&
            qr == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    if (!fn) throw Error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return fn;
}

CUresult encode_tiled(CUtensorMap* map, CUtensorMapDataType dtype, uint32_t rank, const void* gaddr,
                      const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                      const uint32_t* elem_strides, CUtensorMapSwizzle swizzle) {
    cuuint64_t d[5], st[5];
    cuuint32_t b[5], es[5];
    for (uint32_t i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        es[i] = elem_strides ? elem_strides[i] : 1;
        if (i + 1 < rank) st[i] = strides_bytes[i];
    }
    return get_encode()(map, dtype, rank, const_cast<void*>(gaddr), d, st, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
}

static void check_cu(CUresult r, const char* what) {
    if (r != CUDA_SUCCESS) throw Error(std::string("cuTensorMapEncodeTiled failed (") + std::to_string((int)r) + ") for " + what);
}

static void make_map_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t ld_elems,
                        uint32_t box_cols, uint32_t box_rows, const char* what) {
    uint64_t dims[2] = {cols, rows};
    uint64_t st[1] = {ld_elems * 2};
    uint32_t box[2] = {box_cols, box_rows};
    check_cu(encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, st, box, nullptr,
                          CU_TENSOR_MAP_SWIZZLE_128B), what);
}
// NHWC tensor [N][H][W][Cp]; box = {64 channels, bw, bh, bn} output pixels, traversal stride (sv, su) along (W, H)
static void make_map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int Cp, int C_valid, int bn, int bh,
                          int bw, int su, int sv, const char* what) {
    uint64_t dims[4] = {(uint64_t)C_valid, (uint64_t)W, (uint64_t)H, (uint64_t)N};
    uint64_t st[3] = {(uint64_t)Cp * 2, (uint64_t)W * Cp * 2, (uint64_t)H * W * Cp * 2};
    uint32_t box[4] = {64, (uint32_t)(bw * sv), (uint32_t)(bh * su), (uint32_t)bn};
    uint32_t es[4] = {1, (uint32_t)sv, (uint32_t)su, 1};
    check_cu(encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, st, box, es, CU_TENSOR_MAP_SWIZZLE_128B),
             what);
}

// the same tensor with a box that carries one halo row above and below the bh output rows (tc_kernel<.., HALO>).  perm: the
// box is laid out [h][n][w] in shared memory -- tensor-map dimensions (c, w, n, h) -- so that with bn > 1 images per tile a
// vertical tap shift is still ONE uniform offset of bn * bw pixel rows.  Returns false when the driver refuses the map.
static bool make_map_nhwc_halo(CUtensorMap* m, const void* base, int N, int H, int W, int Cp, int C_valid, int bn, int bh,
                               int bw, bool perm) {
    uint32_t es[4] = {1, 1, 1, 1};
    if (!perm) {
        uint64_t dims[4] = {(uint64_t)C_valid, (uint64_t)W, (uint64_t)H, (uint64_t)N};
        uint64_t st[3] = {(uint64_t)Cp * 2, (uint64_t)W * Cp * 2, (uint64_t)H * W * Cp * 2};
        uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)(bh + 2), (uint32_t)bn};
        return encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, st, box, es, CU_TENSOR_MAP_SWIZZLE_128B) == CUDA_SUCCESS;
    }
    uint64_t dims[4] = {(uint64_t)C_valid, (uint64_t)W, (uint64_t)N, (uint64_t)H};
    uint64_t st[3] = {(uint64_t)Cp * 2, (uint64_t)H * W * Cp * 2, (uint64_t)W * Cp * 2};
    uint32_t box[4] = {64, (uint32_t)bw, (uint32_t)bn, (uint32_t)(bh + 2)};
    return encode_tiled(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, st, box, es, CU_TENSOR_MAP_SWIZZLE_128B) == CUDA_SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------------
// staging kernels (HBM-bound)
// ---------------------------------------------------------------------------------------------------------------------
// fp32 [rows][cols] -> bf16 [rows][ld] (ld >= cols, pad columns zeroed).  6 B/element.
__global__ void __launch_bounds__(256) cast_pad_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                       int64_t rows, int64_t cols, int64_t ld) {
    int64_t n = rows * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = i / ld, c = i - r * ld;
        out[i] = __float2bfloat16_rn(c < cols ? in[r * cols + c] : 0.f);
    }
}

// NCHW fp32 -> NHWC bf16 with channel padding: per image a [C][HW] -> [HW][Cp] transpose through shared memory.
// Tile = 16 channels x PX pixels (PX = 256, or 64 for small maps): the reads are PX*4-byte contiguous runs (DRAM pages like
// long runs: the earlier 64-channel x 32-pixel tile ran at half the bandwidth), the writes one full 32-byte sector per pixel.
// grid (ceil(HW/PX), ceil(Cp/16), N), 256 threads.  4 B read + 2 B written per element.
template <int PX>
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const float* __restrict__ in,
                                                                __nv_bfloat16* __restrict__ out, int C, int HW, int Cp) {
    constexpr int PITCH = PX + 4;
    constexpr int V4 = PX / 4;
    constexpr int PER_CH = V4 / 32 > 0 ? V4 / 32 : 1;
    __shared__ __align__(16) float tile[16][PITCH];
    const int n = blockIdx.z;
    const int hw0 = blockIdx.x * PX, c0 = blockIdx.y * 16;
    const float* src = in + (int64_t)n * C * HW;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const bool vec = (HW & 3) == 0 && (((uintptr_t)in & 15) == 0);
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int cl = w + half * 8, c = c0 + cl;
        if (vec) {
            float4 v[PER_CH];
#pragma unroll
            for (int u = 0; u < PER_CH; ++u) {
                const int p4 = lane + u * 32;
                const int hw = hw0 + p4 * 4;
                v[u] = (c < C && p4 < V4 && hw < HW) ? *(const float4*)(src + (int64_t)c * HW + hw) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < PER_CH; ++u) {
                const int p4 = lane + u * 32;
                if (p4 < V4) *(float4*)&tile[cl][p4 * 4] = v[u];
            }
        } else {
            for (int p = lane; p < PX; p += 32) {
                const int hw = hw0 + p;
                tile[cl][p] = (c < C && hw < HW) ? src[(int64_t)c * HW + hw] : 0.f;
            }
        }
    }
    __syncthreads();
    __nv_bfloat16* dst = out + (int64_t)n * HW * Cp;
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

// KCRS fp32 -> packed bf16 through a shared-memory tile, so that both the fp32 reads (runs of TC*RS contiguous floats) and
// the bf16 writes (128-byte rows of 64 channels) are coalesced.  6 B/element.
//   MODE 0 (fwd):   out[k][(r*S+s)*Cp + c] = w[k][c][R-1-r][S-1-s]     tile 4 k x 64 c
//   MODE 1 (dgrad): out[c][(r*S+s)*Kp + k] = w[k][c][R-1-r][S-1-s]     tile 16 k x 16 c (reads: 16*RS-float runs; writes: full
//                                                                      32-byte sectors, neighbours complete the lines in L2)
// (small tiles: a 160x160x3x3 filter still gives 100+ CTAs)
// grid (ceil(Kx/TK), ceil(Cx/TCc)) over the padded extents, 256 threads, dynamic smem TK*(TCc*RS+1) floats.
// RS9: the 3x3 case with RS as a compile-time constant (the index arithmetic is full of divisions by RS)
template <int MODE, bool RS9 = false>
__device__ __forceinline__ void pack_filters_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int K, int C,
                                                  int RS_, int Kp, int Cp, int bx, int by, float* pf_tile) {
    const int RS = RS9 ? 9 : RS_;
    constexpr int TK = MODE == 0 ? 4 : 16, TCc = MODE == 0 ? 64 : 16;
    const int ld = TCc * RS + 1;   // pf_tile: [TK k][TCc c * RS (+1)]
    const int k0 = bx * TK, c0 = by * TCc;
    const int run = min(TCc, C - c0) * RS;   // contiguous floats of one k row inside this tile (<= 0: padding tile)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // TK * TCc == 256 == blockDim: the tile has exactly RS elements per thread.  All loads of a thread are issued before the
    // first shared-memory store (registers), otherwise each block pays RS dependent global-memory round trips.
    const int row_len = TCc * RS;
    auto load_tile = [&](auto rs_tag) {
        constexpr int U = decltype(rs_tag)::value;   // elements per thread handled per trip
        for (int i0 = 0; i0 < RS; i0 += U) {
            float v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = (i0 + u) * 256 + (int)threadIdx.x;
                const int kk = e / row_len, j = e - kk * row_len;
                const int k = k0 + kk;
                v[u] = (i0 + u < RS && k < K && j < run) ? w[((int64_t)k * C + c0) * RS + j] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int e = (i0 + u) * 256 + (int)threadIdx.x;
                const int kk = e / row_len, j = e - kk * row_len;
                if (i0 + u < RS) pf_tile[kk * ld + j] = v[u];
            }
        }
    };
    if (RS % 9 == 0) load_tile(std::integral_constant<int, 9>());
    else if (RS % 5 == 0) load_tile(std::integral_constant<int, 5>());
    else load_tile(std::integral_constant<int, 1>());
    __syncthreads();
    if (MODE == 0) {
        // one (k, tap) row of 64 channels = 128 bytes per warp trip, two channels per lane
        for (int row = wid; row < TK * RS; row += 8) {
            const int kk = row / RS, t = row - kk * RS;
            const int k = k0 + kk, c = c0 + lane * 2;
            if (k < Kp && c < Cp) {   // Cp is even
                const float* src = pf_tile + kk * ld + (RS - 1 - t);
                *(__nv_bfloat162*)(out + ((int64_t)k * RS + t) * Cp + c) =
                    __floats2bfloat162_rn(src[(lane * 2) * RS], src[(lane * 2 + 1) * RS]);
            }
        }
    } else {
        // rows (c, tap) of 16 output channels k = 32 bytes, two k per thread
        for (int item = threadIdx.x; item < TCc * RS * (TK / 2); item += 256) {
            const int row = item / (TK / 2), kp = item - row * (TK / 2);
            const int cc = row / RS, t = row - cc * RS;
            const int c = c0 + cc, k = k0 + kp * 2;
            if (c < Cp && k < Kp) {   // Kp is even
                const float* src = pf_tile + cc * RS + (RS - 1 - t);
                *(__nv_bfloat162*)(out + ((int64_t)c * RS + t) * Kp + k) =
                    __floats2bfloat162_rn(src[(kp * 2) * ld], src[(kp * 2 + 1) * ld]);
            }
        }
    }
}
// Both layouts from ONE read of the filter (a parameter is packed for its convolution and for the feature gradient of the
// same layer every step): tile 16 k x 64 c -- 37 KB of fp32 per CTA for a 3x3 filter, read as runs of 64*RS floats, written
// as 128-byte rows of 64 channels (forward layout) and 32-byte rows of 16 output channels (feature-gradient layout).
template <bool RS9>
__device__ __forceinline__ void pack_filters_both_tile(const float* __restrict__ w, __nv_bfloat16* __restrict__ out_f,
                                                       __nv_bfloat16* __restrict__ out_d, int K, int C, int RS_, int Cp, int Kp2,
                                                       int bx, int by, float* pf_tile) {
    const int RS = RS9 ? 9 : RS_;
    constexpr int TK = 16, TCc = 64;
    const int ld = TCc * RS + 1;
    const int k0 = bx * TK, c0 = by * TCc;
    const int run = min(TCc, C - c0) * RS;
    const int row_len = TCc * RS;
    const int total = TK * row_len;
    // the tile has 4 * RS elements per thread; four loads in flight per trip
    for (int e0 = threadIdx.x; e0 < total; e0 += 4 * 256) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * 256;
            const int kk = e / row_len, j = e - kk * row_len;
            v[u] = (e < total && k0 + kk < K && j < run) ? w[((int64_t)(k0 + kk) * C + c0) * RS + j] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = e0 + u * 256;
            const int kk = e / row_len, j = e - kk * row_len;
            if (e < total) pf_tile[kk * ld + j] = v[u];
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // forward layout: out_f[k][(tap)*Cp + c], one (k, tap) row of 64 channels per warp trip
    for (int row = wid; row < TK * RS; row += 8) {
        const int kk = row / RS, t = row - kk * RS;
        const int k = k0 + kk, c = c0 + lane * 2;
        if (k < K && c < Cp) {
            const float* src = pf_tile + kk * ld + (RS - 1 - t);
            *(__nv_bfloat162*)(out_f + ((int64_t)k * RS + t) * Cp + c) = __floats2bfloat162_rn(src[(lane * 2) * RS], src[(lane * 2 + 1) * RS]);
        }
    }
    // feature-gradient layout: out_d[c][(tap)*Kp2 + k], rows (c, tap) of 16 output channels, two per thread
    for (int item = threadIdx.x; item < TCc * RS * (TK / 2); item += 256) {
        const int row = item / (TK / 2), kp = item - row * (TK / 2);
        const int cc = row / RS, t = row - cc * RS;
        const int c = c0 + cc, k = k0 + kp * 2;
        if (c < C && k < Kp2) {
            const float* src = pf_tile + cc * RS + (RS - 1 - t);
            *(__nv_bfloat162*)(out_d + ((int64_t)c * RS + t) * Kp2 + k) = __floats2bfloat162_rn(src[(kp * 2) * ld], src[(kp * 2 + 1) * ld]);
        }
    }
}
template <int MODE>
__global__ void __launch_bounds__(256) pack_filters_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                           int K, int C, int RS, int Kp, int Cp) {
    extern __shared__ float pf_tile[];
    pack_filters_tile<MODE>(w, out, K, C, RS, Kp, Cp, blockIdx.x, blockIdx.y, pf_tile);
}
// every filter of a plan in one launch: block -> (row, tile) through the rows' tile prefix
__global__ void __launch_bounds__(256) pack_filters_multi_kernel(const FilterPack* __restrict__ rows, int n_rows) {
    extern __shared__ float pf_tile[];
    int lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (rows[mid].tile0 <= (int)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const FilterPack r = rows[lo];
    const int t = (int)blockIdx.x - r.tile0;
    const int bx = t % r.tiles_x, by = t / r.tiles_x;
    if (r.mode == 2) {
        if (r.RS == 9) pack_filters_both_tile<true>(r.w, (__nv_bfloat16*)r.out, (__nv_bfloat16*)r.out2, r.K, r.C, 9, r.Cp, r.Kp2, bx, by, pf_tile);
        else pack_filters_both_tile<false>(r.w, (__nv_bfloat16*)r.out, (__nv_bfloat16*)r.out2, r.K, r.C, r.RS, r.Cp, r.Kp2, bx, by, pf_tile);
    } else if (r.RS == 9) {
        if (r.mode == 0) pack_filters_tile<0, true>(r.w, (__nv_bfloat16*)r.out, r.K, r.C, 9, r.Kp, r.Cp, bx, by, pf_tile);
        else pack_filters_tile<1, true>(r.w, (__nv_bfloat16*)r.out, r.K, r.C, 9, r.Kp, r.Cp, bx, by, pf_tile);
    } else if (r.mode == 0) {
        pack_filters_tile<0>(r.w, (__nv_bfloat16*)r.out, r.K, r.C, r.RS, r.Kp, r.Cp, bx, by, pf_tile);
    } else {
        pack_filters_tile<1>(r.w, (__nv_bfloat16*)r.out, r.K, r.C, r.RS, r.Kp, r.Cp, bx, by, pf_tile);
    }
}

// convolutionFiltersGrad: the tensor-core kernel accumulates into a scratch laid out [tap][C][K] (K contiguous = the
// accumulator's lane dimension, so a warp's 32 atomic adds fall into one 128-byte line); this turns it into dopt's KCRS
// with the filter flip.  grid (ceil(K/32), ceil(C/32)), 256 threads, dynamic smem 32*(32*RS+1) floats.
__device__ __forceinline__ void wgrad_finish_tile(const float* __restrict__ scratch, float* __restrict__ dw, int K, int C, int RS,
                                                  int bx, int by, float* pf_tile) {
    const int ld = 32 * RS + 1;          // pf_tile: [32 k][32 c * RS (+1)]
    const int k0 = bx * 32, c0 = by * 32;
    const int kk = threadIdx.x & 31, k = k0 + kk;
    // rows (tap, c) of 32 consecutive k; 4 * RS rows per warp, loaded four at a time into registers before they are stored
    for (int row0 = threadIdx.x >> 5; row0 < 32 * RS; row0 += 32) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = row0 + u * 8;
            const int t = row / 32, cc = row - t * 32;
            const int c = c0 + cc;
            v[u] = (row < 32 * RS && k < K && c < C) ? scratch[((int64_t)t * C + c) * K + k] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int row = row0 + u * 8;
            const int t = row / 32, cc = row - t * 32;
            if (row < 32 * RS) pf_tile[kk * ld + cc * RS + (RS - 1 - t)] = v[u];
        }
    }
    __syncthreads();
    const int run = min(32, C - c0) * RS;
    for (int r = threadIdx.x >> 5; r < 32; r += 8) {
        const int kr = k0 + r;
        if (kr >= K) break;
        float* dst = dw + ((int64_t)kr * C + c0) * RS;
        for (int j = threadIdx.x & 31; j < run; j += 32) dst[j] = pf_tile[r * ld + j];
    }
}
__global__ void __launch_bounds__(256) wgrad_finish_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int K,
                                                           int C, int RS) {
    extern __shared__ float pf_tile[];
    wgrad_finish_tile(scratch, dw, K, C, RS, blockIdx.x, blockIdx.y, pf_tile);
}
// every deferred filter gradient of a plan (or of one gradient bucket) in one launch: block -> (row, tile) through the rows'
// tile prefix, like pack_filters_multi_kernel
__global__ void __launch_bounds__(256) wgrad_finish_multi_kernel(const WgradFinish* __restrict__ rows, int n_rows) {
    extern __shared__ float pf_tile[];
    int lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (rows[mid].tile0 <= (int)blockIdx.x) lo = mid;
        else hi = mid - 1;
    }
    const WgradFinish r = rows[lo];
    const int t = (int)blockIdx.x - r.tile0;
    wgrad_finish_tile(r.scratch, r.dw, r.K, r.C, r.RS, t % r.tiles_x, t / r.tiles_x, pf_tile);
}
void wgrad_finish_layout(WgradFinish* rows, int n, int* total_tiles, size_t* smem_bytes) {
    int tiles = 0;
    size_t smem = 0;
    for (int i = 0; i < n; ++i) {
        WgradFinish& f = rows[i];
        f.tiles_x = (int)ceil_div(f.K, 32);
        f.tile0 = tiles;
        tiles += f.tiles_x * (int)ceil_div(f.C, 32);
        smem = std::max(smem, (size_t)32 * (32 * f.RS + 1) * sizeof(float));
    }
    *total_tiles = tiles;
    *smem_bytes = smem;
}
void wgrad_finish_launch(const WgradFinish* dev_rows, int n, int total_tiles, size_t smem_bytes, cudaStream_t s) {
    if (n <= 0 || total_tiles <= 0) return;
    DB_REQUIRE(smem_bytes <= 200 * 1024, "filter window too large for the wgrad finish kernel");
    static size_t configured = 48 * 1024;
    if (smem_bytes > configured) {
        DB_CUDA(cudaFuncSetAttribute(wgrad_finish_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    wgrad_finish_multi_kernel<<<(unsigned)total_tiles, 256, smem_bytes, s>>>(dev_rows, n);
    DB_LAUNCH_CHECK();
}

static void pack_filters(const float* w, __nv_bfloat16* out, int K, int C, int RS, int Kp, int Cp, int mode, cudaStream_t s) {
    const int TK = mode == 0 ? 4 : 16, TCc = mode == 0 ? 64 : 16;
    const size_t smem = (size_t)TK * (TCc * RS + 1) * sizeof(float);
    DB_REQUIRE(smem <= 200 * 1024, "filter window too large for the packing kernel");
    DB_REQUIRE(Kp % 2 == 0 || mode == 0, "packed Kout extent must be even");
    DB_REQUIRE(Cp % 2 == 0 || mode == 1, "packed Cin extent must be even");
    static size_t configured[2] = {48 * 1024, 48 * 1024};
    if (smem > configured[mode]) {
        if (mode == 0) DB_CUDA(cudaFuncSetAttribute(pack_filters_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        else DB_CUDA(cudaFuncSetAttribute(pack_filters_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[mode] = 200 * 1024;
    }
    dim3 grid((unsigned)ceil_div(Kp, TK), (unsigned)ceil_div(Cp, TCc));
    if (mode == 0) pack_filters_kernel<0><<<grid, 256, smem, s>>>(w, out, K, C, RS, Kp, Cp);
    else pack_filters_kernel<1><<<grid, 256, smem, s>>>(w, out, K, C, RS, Kp, Cp);
    DB_LAUNCH_CHECK();
}

size_t filter_pack_bytes(const FilterPack& f) {
    return ((size_t)(f.mode == 0 ? f.K : f.C) * f.RS * (f.mode == 0 ? f.Cp : f.Kp) * 2 + 1023) / 1024 * 1024;
}
void filter_pack_layout(FilterPack* rows, int n, int* total_tiles, size_t* smem_bytes) {
    int tiles = 0;
    size_t smem = 0;
    for (int i = 0; i < n; ++i) {
        FilterPack& f = rows[i];
        f.tile0 = tiles;
        f.tiles_x = 1;
        if (f.mode < 0) continue;   // merged into another row
        const int TK = f.mode == 0 ? 4 : 16, TCc = f.mode == 1 ? 16 : 64;
        f.tiles_x = (int)ceil_div(f.mode == 2 ? std::max(f.Kp, f.Kp2) : f.Kp, TK);
        tiles += f.tiles_x * (int)ceil_div(f.Cp, TCc);
        smem = std::max(smem, (size_t)TK * (TCc * f.RS + 1) * sizeof(float));
    }
    *total_tiles = tiles;
    *smem_bytes = smem;
}
void filter_pack_launch(const FilterPack* dev_rows, int n, int total_tiles, size_t smem_bytes, cudaStream_t s) {
    if (n <= 0 || total_tiles <= 0) return;
    DB_REQUIRE(smem_bytes <= 200 * 1024, "filter window too large for the packing kernel");
    static size_t configured = 48 * 1024;
    if (smem_bytes > configured) {
        DB_CUDA(cudaFuncSetAttribute(pack_filters_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    pack_filters_multi_kernel<<<(unsigned)total_tiles, 256, smem_bytes, s>>>(dev_rows, n);
    DB_LAUNCH_CHECK();
}

static void nchw_to_nhwc_bf16(const float* in, __nv_bfloat16* out, int N, int C, int HW, int Cp, cudaStream_t s) {
    if (HW > 64) {
        dim3 grid((unsigned)ceil_div(HW, 256), (unsigned)ceil_div(Cp, 16), (unsigned)N);
        nchw_to_nhwc_bf16_kernel<256><<<grid, 256, 0, s>>>(in, out, C, HW, Cp);
    } else {
        dim3 grid((unsigned)ceil_div(HW, 64), (unsigned)ceil_div(Cp, 16), (unsigned)N);
        nchw_to_nhwc_bf16_kernel<64><<<grid, 256, 0, s>>>(in, out, C, HW, Cp);
    }
    DB_LAUNCH_CHECK();
}

// process-wide staging arena (the reference likewise shares one static workspace, cudnn7.d:111).  All library work is
// stream-ordered on the caller's stream, so consecutive kernels can reuse it.
static Scratch g_stage;
static uint64_t g_stage_generation = 0;
uint64_t tc_stage_generation() { return g_stage_generation; }
static uint8_t* stage_get(size_t bytes) {
    void* before = g_stage.ptr;
    void* p = g_stage.get(bytes);
    if (p != before) ++g_stage_generation;
    return (uint8_t*)p;
}
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// optional per-launch timing of the tcgen05 kernel alone (bench.py's roofline): enabled together with plan profiling
static bool g_tc_prof = false;
static std::vector<std::pair<cudaEvent_t, cudaEvent_t>> g_tc_ev;   // event pairs, reused between profiling sessions
static size_t g_tc_ev_used = 0;
void tc_prof_enable(bool on) {
    g_tc_prof = on;
    if (on) g_tc_ev_used = 0;
}
// device time spent in tc_kernel launches since tc_prof_enable(true); synchronises with the recorded events
void tc_prof_read(double* us, int64_t* launches) {
    double total = 0;
    for (size_t i = 0; i < g_tc_ev_used; ++i) {
        float ms = 0;
        DB_CUDA(cudaEventSynchronize(g_tc_ev[i].second));
        DB_CUDA(cudaEventElapsedTime(&ms, g_tc_ev[i].first, g_tc_ev[i].second));
        total += ms * 1000.0;
    }
    *us = total;
    *launches = (int64_t)g_tc_ev_used;
}

// SMs left to a collective that runs beside the tensor-core kernels (plan.cu arms the gate while gradient-bucket all-reduces may
// be in flight).  The persistent grid is statically scheduled -- every CTA owns a fixed share of the tiles -- so one CTA that
// shares its SM with an NCCL channel, or queues behind one, stretches the whole launch (8 % per convolution,
// profiles/r01c_conv_bisect.md section 7).  Round 1 shrank the grid on the host from the first bucket to the end of backward
// (11 % of the SMs lost for ~3 ms with 16 channels); now the grid is always launched in full and the kernel itself drops its last
// `sms x CTAs-per-SM` CTAs only while the collective's completion counter says an all-reduce is still running
// (TcArgs::gate_*, tc_kernel.cuh).  Each CTA asks for the full 227 KB of shared memory so that it cannot be placed next to an
// NCCL CTA: the two kernels run on disjoint SMs.
static TcGate* g_gate = nullptr;
void tc_set_gate(TcGate* g) { g_gate = g; }

namespace {
__global__ void gate_step_begin_kernel(unsigned* w) {   // [0] done | [1] base | [2..5] snapshots of two streams
    const unsigned d = w[0];
    w[1] = d;
    w[2] = w[3] = w[4] = w[5] = d;
}
__global__ void gate_comm_done_kernel(unsigned* w) { w[0] += 1u; }
}  // namespace
void tc_gate_create(TcGate* g, int sms) {
    DB_CUDA(cudaMalloc((void**)&g->dev, 8 * sizeof(unsigned)));
    DB_CUDA(cudaMemset(g->dev, 0, 8 * sizeof(unsigned)));
    g->sms = sms;
    g->need = 0;
    g->idx[0] = g->idx[1] = 0;
}
void tc_gate_destroy(TcGate* g) {
    if (g->dev) cudaFree(g->dev);
    g->dev = nullptr;
}
void tc_gate_step_begin(TcGate* g, cudaStream_t s) {
    g->need = 0;
    g->idx[0] = g->idx[1] = 0;
    gate_step_begin_kernel<<<1, 1, 0, s>>>(g->dev);
    DB_LAUNCH_CHECK();
}
void tc_gate_comm_done(TcGate* g, cudaStream_t comm) {
    gate_comm_done_kernel<<<1, 1, 0, comm>>>(g->dev);
    DB_LAUNCH_CHECK();
}

// one instantiation of the kernel: opt in to the full shared memory once, launch with the cluster size it needs and as a
// programmatic dependent of the previous kernel in the stream
template <typename K>
static void tc_launch_variant(K kernel, int cluster, int n_ctas, uint32_t smem, cudaStream_t s, const CUtensorMap& tmA,
                              const CUtensorMap& tmB, const TcArgs& a) {
    static std::set<const void*> configured;   // (every instantiation of tc_kernel has the same function type K)
    if (configured.insert((const void*)kernel).second)
        DB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)n_ctas);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    int n = 0;
    if (cluster > 1) {
        at[n].id = cudaLaunchAttributeClusterDimension;
        at[n].val.clusterDim.x = (unsigned)cluster;
        at[n].val.clusterDim.y = 1;
        at[n].val.clusterDim.z = 1;
        ++n;
    }
    if (pdl_enabled()) {
        at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    cfg.attrs = at;
    cfg.numAttrs = (unsigned)n;
    DB_CUDA(cudaLaunchKernelEx(&cfg, kernel, tmA, tmB, a));
}

template <int MODE>
static void tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a_in, int n_ctas, cudaStream_t s) {
    // n_ctas on entry = number of work items (m_tiles * n_tiles * splits)
    TcArgs a = a_in;
    {   // multiply-high constants for the kernel's tile decoding
        const int csize = (MODE == TC_MODE_CONV && a.pair) ? 2 : 1;
        a.fd_mgroups = tc_fastdiv(a.m_tiles / csize);
        a.fd_splits = tc_fastdiv(a.splits);
        a.fd_tiles_q = tc_fastdiv(a.tiles_q);
        a.fd_tiles_p = tc_fastdiv(a.tiles_p);
        const int mt = (a.M + TC_BM - 1) / TC_BM;
        a.fd_wg_mg = tc_fastdiv(MODE == TC_MODE_WGRAD ? (mt + std::max(1, a.wg_nm) - 1) / std::max(1, a.wg_nm) : 1);
        const int total = (MODE == TC_MODE_WGRAD) ? a.pix_tiles : a.k_iters;
        a.per_split = (total + std::max(1, a.splits) - 1) / std::max(1, a.splits);
    }
    TcSmemLayout L = tc_smem_layout(a);
    DB_REQUIRE(L.total <= 227 * 1024, "tcgen05 kernel: shared memory budget exceeded");
    const bool gated = g_gate && g_gate->dev && g_gate->need > 0 && g_gate->sms > 0;
    if (gated && a.nacc == 2) L.total = 227 * 1024;   // one CTA per SM: keep the SM to itself
    if (g_tc_prof) {
        if (g_tc_ev_used == g_tc_ev.size()) {
            cudaEvent_t e0, e1;
            DB_CUDA(cudaEventCreate(&e0));
            DB_CUDA(cudaEventCreate(&e1));
            g_tc_ev.emplace_back(e0, e1);
        }
        DB_CUDA(cudaEventRecord(g_tc_ev[g_tc_ev_used].first, s));
    }
    // persistent grid: one CTA (pair) per SM, each looping over its share of the n_ctas work items
    {
        const int cs = a.pair ? 2 : 1;
        const int per_sm = a.nacc == 2 ? 1 : 2;
        const int full = per_sm * (sm_count() / cs * cs) / cs;
        int clusters = std::min(n_ctas / cs, full);
        if (clusters < 1) clusters = 1;
        n_ctas = clusters * cs;
        if (gated && clusters == full) {
            // a full grid: its tail is droppable (decided on the device, see TcGate)
            const int drop = std::min(g_gate->sms, sm_count() / 2) * per_sm / cs * cs;
            if (drop > 0 && drop < n_ctas) {
                const int chain = (s == g_gate->side) ? 1 : 0;
                const int i = g_gate->idx[chain]++;
                a.gate_rd = g_gate->dev + 2 + chain * 2 + (i & 1);
                a.gate_wr = g_gate->dev + 2 + chain * 2 + ((i + 1) & 1);
                a.gate_done = g_gate->dev;
                a.gate_base = g_gate->dev + 1;
                a.gate_need = g_gate->need;
                a.gate_drop = drop;
            }
        }
    }
    // epilogue flavour: 0 plain, 1 statistics of the result, 2 result + companion (residual sum), 3 gated backward statistics
    const int epi = a.ep_src ? (a.ep_coef ? 3 : 2) : (a.st_cols > 0 ? 1 : 0);
    if (a.pair) {
        // cta_group::2 kernels need an even cluster
#define TC_PAIR_CASE(KBOX_, EPI_, HALO_)                                                                                        \
        tc_launch_variant(tc_kernel<TC_MODE_CONV, true, false, KBOX_, EPI_, HALO_>, 2, n_ctas, L.total, s, tmA, tmB, a)
#define TC_PAIR_EPI(KBOX_, HALO_)                                                                                               \
        do {                                                                                                                     \
            if (epi == 0) TC_PAIR_CASE(KBOX_, 0, HALO_);                                                                          \
            else if (epi == 1) TC_PAIR_CASE(KBOX_, 1, HALO_);                                                                     \
            else if (epi == 2) TC_PAIR_CASE(KBOX_, 2, HALO_);                                                                     \
            else TC_PAIR_CASE(KBOX_, 3, HALO_);                                                                                   \
        } while (0)
        if (a.trace || a.dbg) tc_launch_variant(tc_kernel<TC_MODE_CONV, true, true>, 2, n_ctas, L.total, s, tmA, tmB, a);
        else if (a.halo) TC_PAIR_EPI(1, true);
        else if (a.kbox == 2) TC_PAIR_EPI(2, false);
        else TC_PAIR_EPI(1, false);
#undef TC_PAIR_EPI
#undef TC_PAIR_CASE
    } else if (a.trace || a.dbg) {
        tc_launch_variant(tc_kernel<MODE, false, true>, 1, n_ctas, L.total, s, tmA, tmB, a);
    } else if (MODE == TC_MODE_CONV && epi == 1) {
        tc_launch_variant(tc_kernel<TC_MODE_CONV, false, false, 1, 1>, 1, n_ctas, L.total, s, tmA, tmB, a);
    } else if (MODE == TC_MODE_CONV && epi == 2) {
        tc_launch_variant(tc_kernel<TC_MODE_CONV, false, false, 1, 2>, 1, n_ctas, L.total, s, tmA, tmB, a);
    } else if (MODE == TC_MODE_CONV && epi == 3) {
        tc_launch_variant(tc_kernel<TC_MODE_CONV, false, false, 1, 3>, 1, n_ctas, L.total, s, tmA, tmB, a);
    } else if (MODE == TC_MODE_WGRAD && a.halo) {
        tc_launch_variant(tc_kernel<TC_MODE_WGRAD, false, false, 1, 0, true>, 1, n_ctas, L.total, s, tmA, tmB, a);
    } else {
        tc_launch_variant(tc_kernel<MODE>, 1, n_ctas, L.total, s, tmA, tmB, a);
    }
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess)
            throw Error(std::string("tcgen05 kernel launch failed: ") + cudaGetErrorString(e) + " [mode " + std::to_string(MODE) +
                        " grid " + std::to_string(n_ctas) + " cluster " + std::to_string(a.cluster) + " pair " +
                        std::to_string(a.pair) + " m_tiles " + std::to_string(a.m_tiles) + " n_tiles " +
                        std::to_string(a.n_tiles) + " splits " + std::to_string(a.splits) + " BN " + std::to_string(a.BN) +
                        " stages " + std::to_string(a.stages) + " smem " + std::to_string(L.total) + "]");
        count_launch();
    }
    if (g_tc_prof) DB_CUDA(cudaEventRecord(g_tc_ev[g_tc_ev_used++].second, s));
}

static int conv_kbox() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DOPT_B200_KBOX");
        v = e ? (atoi(e) == 2 ? 2 : 1) : 2;
    }
    return v;
}

// Persistent CTAs come in two shapes (measured on the WRN layers, profiles/r01c_conv_bisect.md):
//  * one per SM, two TMEM accumulators, the whole shared memory as operand ring: the epilogue of tile i overlaps the main
//    loop of tile i+1 inside the CTA.  Best when an SM gets many tiles (C=160 layer: 6.9 tiles per SM) and for wgrad.
//  * two per SM, one accumulator each and half the ring: two independent pipelines hide each other's tile boundaries
//    and TMA latency.  Best when an SM only gets a few tiles (C=320/640 layers).
static int g_ctas_per_sm = 0;   // 0 = choose per problem
static int pick_stages(TcArgs& a, int64_t items = 0) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char* e = getenv("DOPT_B200_CTAS_PER_SM")) g_ctas_per_sm = atoi(e);
    }
    int per_sm = g_ctas_per_sm;
    // (halo stages are ~50 KB: two CTAs per SM would get two stages each -- measured, profiles/r02_summary.md: 5.81 ms per step
    // with one CTA per SM against 5.91 ms)
    if (per_sm != 1 && per_sm != 2)
        per_sm = (a.mode == TC_MODE_WGRAD || a.halo || items >= (int64_t)6 * sm_count()) ? 1 : 2;
    a.nacc = per_sm == 1 ? 2 : 1;
    a.stages = 2;
    TcSmemLayout L = tc_smem_layout(a);
    const int extra = 2048 + a.st_cols * (a.ep_coef ? 48 : 32);   // barriers, alignment slack, epilogue statistics (+ coefficients)
    int st = (int)(((per_sm == 1 ? 225 : 110) * 1024 - extra) / L.stage_bytes);
    if (st > 8) st = 8;
    if (st < 2) {
        // the tile does not fit twice: fall back to one CTA per SM
        a.nacc = 2;
        st = (int)((225 * 1024 - extra) / L.stage_bytes);
        if (st > 8) st = 8;
        if (st < 2) st = 2;
    }
    return st;
}

// cta_group::2 pair tiles (256 x BN): each SM ingests only half of the filter tile
static int g_use_pair = 1;
static bool pick_pair(int m_tiles, int bn) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char* e = getenv("DOPT_B200_PAIR")) g_use_pair = atoi(e);
    }
    return g_use_pair && m_tiles >= 4 && bn % 16 == 0 && (bn / 2) % 8 == 0;
}

static int pick_bn(int nout) {
    if (nout <= 256) return (int)align_up(nout, 16);
    for (int bn = 256; bn >= 64; bn -= 16)
        if (nout % bn == 0) return bn;
    return 128;
}

// =====================================================================================================================
// matmul
// =====================================================================================================================
struct TcGemm {
    int64_t M, N, K, Kp, Np;
    TcArgs args{};
    int n_ctas;
};

bool tc_gemm_supported(int64_t M, int64_t N, int64_t K) {
    // worth a 128-row tensor-core tile only when the product is big enough; everything else is latency-bound anyway.
    // (Measured, profiles/r02_summary.md: the dense layer of the WRN configs, [128,640]x[640,100] and its two gradients, takes
    // 14.5 us per launch here -- two operand casts + a one-tile launch -- against 22 us on the split-K SIMT kernel's predecessor;
    // it stays on the fp32 path, which also keeps its results at fp32 accuracy.)
    return M >= 128 && N >= 64 && K >= 64 && M * N * K >= (int64_t)1 << 24 && M < (1 << 30) && N < (1 << 30);
}
TcGemm* tc_gemm_create(int64_t M, int64_t N, int64_t K) {
    auto* g = new TcGemm;
    g->M = M; g->N = N; g->K = K;
    g->Kp = (int64_t)align_up(K, 8);
    g->Np = (int64_t)align_up(N, 8);
    TcArgs& a = g->args;
    a.mode = TC_MODE_GEMM;
    a.BN = pick_bn((int)N);
    a.n_tiles = (int)ceil_div(N, a.BN);
    a.splits = 1;
    a.k_iters = (int)ceil_div(K, TC_BK);
    a.M = (int)M; a.N = (int)N; a.Nout = (int)N;
    a.out_kind = TC_OUT_F32;
    a.o_sn = N; a.o_sc = 1;
    a.stages = pick_stages(a);
    a.m_tiles = (int)ceil_div(M, TC_BM);
    a.cluster = 1;
    g->n_ctas = a.m_tiles * a.n_tiles;
    return g;
}
void tc_gemm_run(TcGemm* g, const float* A, const float* B, float* C, cudaStream_t s) {
    size_t a_bytes = align_up((size_t)g->M * g->Kp * 2, 1024), b_bytes = align_up((size_t)g->K * g->Np * 2, 1024);
    uint8_t* st = stage_get(a_bytes + b_bytes);
    auto* Ab = (__nv_bfloat16*)st;
    auto* Bb = (__nv_bfloat16*)(st + a_bytes);
    cast_pad_kernel<<<stream_grid(g->M * g->Kp, 256, 8), 256, 0, s>>>(A, Ab, g->M, g->K, g->Kp);
    DB_LAUNCH_CHECK();
    cast_pad_kernel<<<stream_grid(g->K * g->Np, 256, 8), 256, 0, s>>>(B, Bb, g->K, g->N, g->Np);
    DB_LAUNCH_CHECK();
    CUtensorMap tmA, tmB;
    make_map_2d(&tmA, Ab, (uint64_t)g->K, (uint64_t)g->M, (uint64_t)g->Kp, 64, 128, "matmul A");
    make_map_2d(&tmB, Bb, (uint64_t)g->N, (uint64_t)g->K, (uint64_t)g->Np, 64, 64, "matmul B");
    TcArgs a = g->args;
    a.out = C;
    tc_launch<TC_MODE_GEMM>(tmA, tmB, a, g->n_ctas, s);
}
void tc_gemm_destroy(TcGemm* g) { delete g; }

// =====================================================================================================================
// convolution
// =====================================================================================================================
struct PixelBox {
    int bn, bh, bw, tiles_n, tiles_p, tiles_q;
};
// cover an (N, P, Q) pixel grid with boxes of at most 128 pixels; `mult16` additionally requires bn*bh*bw % 16 == 0
static int g_wgrad_pix = 64;   // pixels per wgrad k-block: 64 keeps the stages at ~36 KB so the ring is 6 deep
static bool pick_box(int N, int P, int Q, bool mult16, PixelBox& b) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char* e = getenv("DOPT_B200_WG_PIX")) g_wgrad_pix = atoi(e) >= 128 ? 128 : (atoi(e) >= 64 ? 64 : 32);
    }
    const int cap = mult16 ? g_wgrad_pix : 128;
    int bw = Q <= cap ? Q : cap;
    int best = 0;
    for (int bh = std::min(P, cap / bw); bh >= 1; --bh) {
        int maxn = (bh == P) ? std::min(N, cap / (bw * bh)) : 1;
        for (int bn = maxn; bn >= 1; --bn) {
            int pix = bn * bh * bw;
            if (mult16 && pix % 16) continue;
            if (pix > best) {
                best = pix;
                b.bn = bn; b.bh = bh; b.bw = bw;
            }
            break;   // smaller bn only lowers the pixel count for this bh
        }
        if (!mult16 && best) break;
    }
    if (!best) return false;
    if (b.bw * 2 > 256 || b.bh * 2 > 256) return false;
    b.tiles_n = (int)ceil_div(N, b.bn);
    b.tiles_p = (int)ceil_div(P, b.bh);
    b.tiles_q = (int)ceil_div(Q, b.bw);
    return true;
}

struct ConvTc {
    ConvGeom g;
    int kind;
    int Cp, Kp;
    PixelBox box;
    const void* pre[2] = {nullptr, nullptr};   // operands already staged as NHWC bf16 by the plan
    const void* pre_w = nullptr;               // filter already packed by the plan (fwd / dgrad)
    void* out_staged = nullptr;                // fwd / dgrad: write the result as [N][H][W][Cp] bf16 here instead of NCHW fp32
    void* stats_ws = nullptr;                  // fwd + out_staged: statistics workspace of the batch norm reading the result
    float* acc_private = nullptr;              // wgrad: private accumulation scratch, finished later by the plan (deferred)
    bool acc_external = false;                 // ... owned and zeroed by the plan (one arena, one memset per step)
    // epilogue companion (TcArgs::ep_src): fwd + out_staged: the addend of the residual sum this convolution feeds (mode 2);
    // dgrad + out_staged: x and the forward coefficients of the batch norm whose backward pass reads this result (mode 3)
    int ep_mode = 0;
    const void* ep_src = nullptr;
    const float* ep_coef = nullptr;
};

void stage_nchw_to_nhwc_bf16(const float* in, void* out, int N, int C, int64_t HW, cudaStream_t s) {
    nchw_to_nhwc_bf16(in, (__nv_bfloat16*)out, N, C, (int)HW, (int)align_up(C, 8), s);
}
size_t staged_nhwc_bytes(int N, int C, int64_t HW) { return align_up((size_t)N * HW * align_up(C, 8) * 2, 1024); }
bool conv_tc_filter_pack(const ConvTc* c, int input, FilterPack* d) {
    if (!c || input != 1 || c->kind == CONV_WGRAD) return false;
    const ConvGeom& g = c->g;
    d->w = nullptr;
    d->out = nullptr;
    d->K = g.K; d->C = g.C; d->RS = g.R * g.S;
    d->mode = c->kind == CONV_FWD ? 0 : 1;
    d->Kp = c->kind == CONV_FWD ? g.K : c->Kp;
    d->Cp = c->kind == CONV_FWD ? c->Cp : g.C;
    d->tiles_x = d->tile0 = 0;
    d->out2 = nullptr;
    d->Kp2 = 0;
    return true;
}
void conv_tc_set_packed_filter(ConvTc* c, const void* packed) {
    if (c) c->pre_w = packed;
}
void conv_tc_set_staged(ConvTc* c, int input, const void* p) {
    if (c && input >= 0 && input < 2) c->pre[input] = p;
}
bool conv_tc_can_stage_output(const ConvTc* c) { return c && c->kind != CONV_WGRAD; }
void conv_tc_set_stats_workspace(ConvTc* c, void* w) {
    if (c && c->kind != CONV_WGRAD) c->stats_ws = w;
}
bool conv_tc_can_companion(const ConvTc* c, int mode) {
    if (!c) return false;
    const ConvGeom& g = c->g;
    if (!((mode == 2 && c->kind == CONV_FWD) || (mode == 3 && c->kind == CONV_DGRAD))) return false;
    // Only where the launch will be the halo pipeline (try_halo): a 3 x 3 unit-stride convolution on cta_group::2 pair tiles.
    // Those instantiations run one CTA per SM and keep the whole companion row in registers; the others (two CTAs per SM, 80
    // registers) would spill it.  (The backward statistics also need the single launch of a unit-stride feature gradient.)
    if (g.R != 3 || g.S != 3 || g.u != 1 || g.v != 1 || g.ph != 1 || g.pw != 1) return false;
    if (const char* e = getenv("DOPT_B200_HALO"))
        if (atoi(e) < 2) return false;
    const PixelBox& b = c->box;
    const int m_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    return pick_pair(m_tiles, pick_bn(mode == 2 ? g.K : g.C)) && (b.bn * b.bw) % 8 == 0 && b.bn * b.bh * b.bw == TC_BM;
}
void conv_tc_set_companion(ConvTc* c, int mode, const void* src, const float* coef) {
    if (!c) return;
    c->ep_mode = mode;
    c->ep_src = src;
    c->ep_coef = coef;
}
void conv_tc_set_staged_output(ConvTc* c, void* nhwc_bf16) {
    if (c && c->kind != CONV_WGRAD) c->out_staged = nhwc_bf16;
}
size_t conv_tc_staged_bytes(const ConvTc* c, int input) {
    if (!c) return 0;
    const ConvGeom& g = c->g;
    // input 0: x (fwd) / dy (dgrad, wgrad); input 1: filters (fwd, dgrad: never staged by the plan) / x (wgrad)
    if (input == 0) return c->kind == CONV_FWD ? staged_nhwc_bytes(g.N, g.C, (int64_t)g.H * g.W) : staged_nhwc_bytes(g.N, g.K, (int64_t)g.P * g.Q);
    if (input == 1 && c->kind == CONV_WGRAD) return staged_nhwc_bytes(g.N, g.C, (int64_t)g.H * g.W);
    return 0;
}

bool conv_tc_supported(const ConvGeom& g, int kind) {
    if (g.R * g.S > TC_MAX_TAPS) return false;
    if (g.u > 2 || g.v > 2) return false;
    // tiny-channel layers (the 3-channel stem) are bandwidth-bound: fp32 direct kernel.  Exception: the stem's filter
    // gradient, a 131072-pixel reduction the direct kernel spends 0.15 ms on
    if (g.K < 16) return false;
    if (g.C < 16 && (kind != CONV_WGRAD || (int64_t)g.N * g.P * g.Q < 16384)) return false;
    if ((int64_t)g.N * g.P * g.Q < 128) return false;
    PixelBox b;
    if (kind == CONV_FWD) return pick_box(g.N, g.P, g.Q, false, b);
    if (kind == CONV_DGRAD) {
        // phase grids of the input: ceil(H/u) x ceil(W/v)
        return pick_box(g.N, (g.H + g.u - 1) / g.u, (g.W + g.v - 1) / g.v, false, b);
    }
    return pick_box(g.N, g.P, g.Q, true, b);
}

ConvTc* conv_tc_create(const ConvGeom& g, int kind) {
    auto* c = new ConvTc;
    c->g = g;
    c->kind = kind;
    c->Cp = (int)align_up(g.C, 8);
    c->Kp = (int)align_up(g.K, 8);
    bool ok;
    if (kind == CONV_FWD) ok = pick_box(g.N, g.P, g.Q, false, c->box);
    else if (kind == CONV_DGRAD) ok = pick_box(g.N, (g.H + g.u - 1) / g.u, (g.W + g.v - 1) / g.v, false, c->box);
    else ok = pick_box(g.N, g.P, g.Q, true, c->box);
    DB_REQUIRE(ok, "conv_tc_create: unsupported geometry");
    return c;
}
void conv_tc_destroy(ConvTc* c) {
    if (c && c->acc_private && !c->acc_external) cudaFree(c->acc_private);
    delete c;
}
bool conv_tc_defer_finish(ConvTc* c, WgradFinish* row) {
    if (!c || c->kind != CONV_WGRAD) return false;
    const ConvGeom& g = c->g;
    const size_t bytes = (size_t)g.R * g.S * g.K * g.C * sizeof(float);
    if (!c->acc_private) DB_CUDA(cudaMalloc((void**)&c->acc_private, std::max<size_t>(bytes, 16)));
    row->scratch = c->acc_private;
    row->dw = nullptr;
    row->K = g.K; row->C = g.C; row->RS = g.R * g.S;
    row->tiles_x = row->tile0 = 0;
    return true;
}

void conv_tc_set_scratch(ConvTc* c, float* p) {
    if (!c || c->kind != CONV_WGRAD || !p) return;
    if (c->acc_private && !c->acc_external) cudaFree(c->acc_private);
    c->acc_private = p;
    c->acc_external = true;
}

bool conv_tc_side_stream_safe(const ConvTc* c) {
    return c && c->kind == CONV_WGRAD && c->pre[0] && c->pre[1] && c->acc_private;
}

// Switch a pair-tile CONV launch whose taps are exactly the 3 x 3 neighbourhood {-1, 0, 1}^2 at unit stride to the halo
// pipeline (tc_kernel<.., HALO>): stages become (filter column, channel block) and tmA gets the box with halo rows.
// DOPT_B200_HALO=0 switches it off; =1 restricts it to one-image tiles (no permuted tensor map).
static bool try_halo(TcArgs& a, CUtensorMap* tmA, const void* base, int N, int H, int W, int Cp, int C_valid) {
    static int mode = -1;
    if (mode < 0) {
        const char* e = getenv("DOPT_B200_HALO");
        mode = e ? atoi(e) : 2;
    }
    if (mode == 0 || !a.pair || a.taps != 9 || a.a_su != 1 || a.a_sv != 1 || a.trace || a.dbg) return false;
    if ((a.bn * a.bw) % 8 != 0 || a.bn * a.bh * a.bw != TC_BM) return false;
    if (a.bn > 1 && mode < 2) return false;
    int col[3][3];
    for (int j = 0; j < 3; ++j)
        for (int v = 0; v < 3; ++v) col[j][v] = -1;
    for (int t = 0; t < 9; ++t) {
        const int dh = a.tap_dh[t], dw = a.tap_dw[t];
        if (dh < -1 || dh > 1 || dw < -1 || dw > 1 || col[dw + 1][dh + 1] >= 0) return false;
        col[dw + 1][dh + 1] = a.tap_bcol[t];
    }
    CUtensorMap m;
    if (!make_map_nhwc_halo(&m, base, N, H, W, Cp, C_valid, a.bn, a.bh, a.bw, a.bn > 1)) return false;
    *tmA = m;
    a.halo = a.bn > 1 ? 2 : 1;
    a.kbox = 1;
    a.taps = 3;
    for (int j = 0; j < 3; ++j) {
        a.tap_dw[j] = j - 1;
        a.tap_dh[j] = -1;
        for (int v = 0; v < 3; ++v) a.hb_col[j][v] = col[j][v];
    }
    a.k_iters = 3 * a.c_iters;
    return true;
}

static int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static int pos_mod(int a, int b) { return ((a % b) + b) % b; }

static void run_fwd(ConvTc* c, const float* x, const float* w, float* y, cudaStream_t s) {
    const ConvGeom& g = c->g;
    const int RS = g.R * g.S, Cp = c->Cp;
    size_t xb = align_up((size_t)g.N * g.H * g.W * Cp * 2, 1024), wb = align_up((size_t)g.K * RS * Cp * 2, 1024);
    uint8_t* st = stage_get(xb + wb);
    auto* xh = (__nv_bfloat16*)st;
    auto* wp = (__nv_bfloat16*)(st + xb);
    if (c->pre[0]) xh = (__nv_bfloat16*)c->pre[0];
    else nchw_to_nhwc_bf16(x, xh, g.N, g.C, g.H * g.W, Cp, s);
    if (c->pre_w) wp = (__nv_bfloat16*)c->pre_w;
    else pack_filters(w, wp, g.K, g.C, RS, g.K, Cp, 0, s);
    const PixelBox& b = c->box;
    CUtensorMap tmA, tmB;
    make_map_nhwc(&tmA, xh, g.N, g.H, g.W, Cp, g.C, b.bn, b.bh, b.bw, g.u, g.v, "convolution x");
    TcArgs a{};
    a.mode = TC_MODE_CONV;
    a.BN = pick_bn(g.K);
    a.m_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    a.pair = pick_pair(a.m_tiles, a.BN) ? 1 : 0;
    a.cluster = 1;
    a.m_tiles = (int)align_up(a.m_tiles, a.pair ? 2 : 1);
    make_map_2d(&tmB, wp, (uint64_t)RS * Cp, (uint64_t)g.K, (uint64_t)RS * Cp, 64,
                (uint32_t)(a.pair ? a.BN / 2 : a.BN), "convolution w");
    a.n_tiles = (int)ceil_div(g.K, a.BN);
    a.splits = 1;
    a.taps = RS;
    a.c_iters = (int)ceil_div(g.C, TC_BK);
    a.c_valid = g.C;
    a.k_iters = RS * a.c_iters;
    for (int r = 0; r < g.R; ++r)
        for (int q = 0; q < g.S; ++q) {
            int t = r * g.S + q;
            a.tap_dh[t] = r - g.ph;
            a.tap_dw[t] = q - g.pw;
            a.tap_bcol[t] = t * Cp;
        }
    a.bn = b.bn; a.bh = b.bh; a.bw = b.bw;
    a.tiles_p = b.tiles_p; a.tiles_q = b.tiles_q;
    a.a_su = g.u; a.a_sv = g.v;
    a.NI = g.N; a.OP = g.P; a.OQ = g.Q;
    a.Nout = g.K;
    a.out_kind = TC_OUT_F32;
    a.o_off = 0;
    a.o_sn = (long long)g.K * g.P * g.Q; a.o_sc = (long long)g.P * g.Q; a.o_sh = g.Q; a.o_sw = 1;
    a.out = y;
    if (c->out_staged) {   // bf16-interior plans: the next reader takes NHWC bf16 (channel padding stays zero from allocation)
        a.out_kind = TC_OUT_BF16;
        a.o_sn = (long long)g.P * g.Q * c->Kp; a.o_sc = 1; a.o_sh = (long long)g.Q * c->Kp; a.o_sw = c->Kp;
        a.out = c->out_staged;
        if (c->stats_ws && !getenv("DOPT_B200_DBG") && !getenv("DOPT_B200_TRACE")) {
            a.st_cols = a.n_tiles * a.BN;
            a.st_cp = c->Kp;
            flat_stats_sink(c->stats_ws, g.K, &a.st_epoch, &a.st_sums, &a.st_copies);
        }
        if (c->ep_mode == 2) {
            DB_REQUIRE(c->ep_src, "convolution: the addend of its residual sum was not bound");
            a.ep_src = c->ep_src;
        }
    } else {
        DB_REQUIRE(c->ep_mode == 0, "convolution: an epilogue companion needs the NHWC bf16 result");
    }
    a.kbox = (a.pair && conv_kbox() == 2 && !getenv("DOPT_B200_DBG") && !getenv("DOPT_B200_TRACE")) ? 2 : 1;
    if (!getenv("DOPT_B200_DBG") && !getenv("DOPT_B200_TRACE")) try_halo(a, &tmA, xh, g.N, g.H, g.W, Cp, g.C);
    a.stages = pick_stages(a, (int64_t)a.m_tiles * a.n_tiles);
    if (const char* e = getenv("DOPT_B200_DBG")) a.dbg = atoi(e);
    static unsigned long long* trace_dev = nullptr;
    static int trace_runs = 0;
    if (getenv("DOPT_B200_TRACE") && (atoi(getenv("DOPT_B200_TRACE")) <= 1 || atoi(getenv("DOPT_B200_TRACE")) == g.C) && trace_runs < 6) {
        if (!trace_dev) DB_CUDA(cudaMalloc(&trace_dev, 1024 * 8));
        DB_CUDA(cudaMemsetAsync(trace_dev, 0, 1024 * 8, s));
        a.trace = trace_dev;
    }
    if (const char* e = getenv("DOPT_B200_STAGES")) a.stages = atoi(e);
    tc_launch<TC_MODE_CONV>(tmA, tmB, a, a.m_tiles * a.n_tiles, s);
    if (a.trace && ++trace_runs == 5) {
        std::vector<unsigned long long> t(1024);
        DB_CUDA(cudaStreamSynchronize(s));
        DB_CUDA(cudaMemcpy(t.data(), trace_dev, 1024 * 8, cudaMemcpyDeviceToHost));
        unsigned long long t0 = t[0];
        fprintf(stderr, "TRACE k_iters %d stages %d BN %d pair %d\n", a.k_iters, a.stages, a.BN, a.pair);
        fprintf(stderr, "producer (after empty wait) : ");
        for (int i = 0; i < 64; ++i) fprintf(stderr, "%lld ", (long long)(t[i] - t0));
        fprintf(stderr, "\nproducer B (after empty)    : ");
        for (int i = 0; i < 64; ++i) fprintf(stderr, "%lld ", (long long)(t[768 + i] - t0));
        fprintf(stderr, "\nmma (after full wait)       : ");
        for (int i = 0; i < 64; ++i) fprintf(stderr, "%lld ", (long long)(t[256 + i] - t0));
        fprintf(stderr, "\nepilogue (start,end) per tile: ");
        for (int i = 0; i < 8; ++i) fprintf(stderr, "(%lld,%lld) ", (long long)(t[512 + 2 * i] - t0), (long long)(t[512 + 2 * i + 1] - t0));
        fprintf(stderr, "\n");
    }
}

static void run_dgrad(ConvTc* c, const float* dy, const float* w, float* dx, cudaStream_t s) {
    const ConvGeom& g = c->g;
    const int RS = g.R * g.S, Kp = c->Kp;
    size_t yb = align_up((size_t)g.N * g.P * g.Q * Kp * 2, 1024), wb = align_up((size_t)g.C * RS * Kp * 2, 1024);
    uint8_t* st = stage_get(yb + wb);
    auto* dyh = (__nv_bfloat16*)st;
    auto* wp = (__nv_bfloat16*)(st + yb);
    if (c->pre[0]) dyh = (__nv_bfloat16*)c->pre[0];
    else nchw_to_nhwc_bf16(dy, dyh, g.N, g.K, g.P * g.Q, Kp, s);
    if (c->pre_w) wp = (__nv_bfloat16*)c->pre_w;
    else pack_filters(w, wp, g.K, g.C, RS, Kp, g.C, 1, s);
    const PixelBox& b = c->box;
    CUtensorMap tmA, tmB;
    make_map_nhwc(&tmA, dyh, g.N, g.P, g.Q, Kp, g.K, b.bn, b.bh, b.bw, 1, 1, "convolutionFeaturesGrad dy");
    int BN = pick_bn(g.C);
    int m_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    int pair = pick_pair(m_tiles, BN) ? 1 : 0;
    int cluster = 1;
    m_tiles = (int)align_up(m_tiles, pair ? 2 : 1);
    make_map_2d(&tmB, wp, (uint64_t)RS * Kp, (uint64_t)g.C, (uint64_t)RS * Kp, 64,
                (uint32_t)(pair ? BN / 2 : BN), "convolutionFeaturesGrad w");
    bool need_zero = false;
    std::vector<TcArgs> launches;
    for (int pa = 0; pa < g.u; ++pa)
        for (int pb = 0; pb < g.v; ++pb) {
            TcArgs a{};
            a.mode = TC_MODE_CONV;
            a.BN = BN;
            a.m_tiles = m_tiles;
            a.cluster = cluster;
            a.pair = pair;
            a.n_tiles = (int)ceil_div(g.C, BN);
            a.splits = 1;
            a.c_iters = (int)ceil_div(g.K, TC_BK);
            a.c_valid = g.K;
            int nt = 0;
            for (int r = 0; r < g.R; ++r) {
                if (pos_mod(pa + g.ph - r, g.u)) continue;
                for (int q = 0; q < g.S; ++q) {
                    if (pos_mod(pb + g.pw - q, g.v)) continue;
                    a.tap_dh[nt] = floor_div(pa + g.ph - r, g.u);
                    a.tap_dw[nt] = floor_div(pb + g.pw - q, g.v);
                    a.tap_bcol[nt] = (r * g.S + q) * Kp;
                    ++nt;
                }
            }
            int OPh = (g.H - pa + g.u - 1) / g.u, OQh = (g.W - pb + g.v - 1) / g.v;   // rows/cols of this phase
            if (nt == 0 || OPh <= 0 || OQh <= 0) {
                if (OPh > 0 && OQh > 0) need_zero = true;
                continue;
            }
            a.taps = nt;
            a.k_iters = nt * a.c_iters;
            a.bn = b.bn; a.bh = b.bh; a.bw = b.bw;
            a.tiles_p = b.tiles_p; a.tiles_q = b.tiles_q;
            a.a_su = 1; a.a_sv = 1;
            a.NI = g.N; a.OP = OPh; a.OQ = OQh;
            a.Nout = g.C;
            a.out_kind = TC_OUT_F32;
            a.o_off = (long long)pa * g.W + pb;
            a.o_sn = (long long)g.C * g.H * g.W; a.o_sc = (long long)g.H * g.W;
            a.o_sh = (long long)g.u * g.W; a.o_sw = g.v;
            a.out = dx;
            if (c->out_staged) {
                const long long Cp = c->Cp;
                a.out_kind = TC_OUT_BF16;
                a.o_off = ((long long)pa * g.W + pb) * Cp;
                a.o_sn = (long long)g.H * g.W * Cp; a.o_sc = 1;
                a.o_sh = (long long)g.u * g.W * Cp; a.o_sw = (long long)g.v * Cp;
                a.out = c->out_staged;
                if (c->ep_mode == 3) {
                    DB_REQUIRE(g.u == 1 && g.v == 1 && c->ep_src && c->ep_coef && c->stats_ws,
                               "convolutionFeaturesGrad: backward batch-norm statistics need x, the forward coefficients and the sink");
                    a.ep_src = c->ep_src;
                    a.ep_coef = c->ep_coef;
                    a.st_cols = a.n_tiles * a.BN;
                    a.st_cp = (int)Cp;
                    flat_stats_sink(c->stats_ws, g.C, &a.st_epoch, &a.st_sums, &a.st_copies);
                }
            } else {
                DB_REQUIRE(c->ep_mode == 0, "convolutionFeaturesGrad: an epilogue companion needs the NHWC bf16 result");
            }
            a.kbox = (a.pair && conv_kbox() == 2 && !getenv("DOPT_B200_DBG") && !getenv("DOPT_B200_TRACE")) ? 2 : 1;
            if (g.u == 1 && g.v == 1 && !getenv("DOPT_B200_DBG") && !getenv("DOPT_B200_TRACE"))
                try_halo(a, &tmA, dyh, g.N, g.P, g.Q, Kp, g.K);
            a.stages = pick_stages(a, (int64_t)a.m_tiles * a.n_tiles);
            launches.push_back(a);
        }
    if (need_zero) {
        if (c->out_staged) DB_CUDA(cudaMemsetAsync(c->out_staged, 0, (size_t)g.N * g.H * g.W * c->Cp * 2, s));
        else DB_CUDA(cudaMemsetAsync(dx, 0, (size_t)g.N * g.C * g.H * g.W * sizeof(float), s));
        count_launch();
    }
    for (auto& a : launches) tc_launch<TC_MODE_CONV>(tmA, tmB, a, a.m_tiles * a.n_tiles, s);
}

static void run_wgrad(ConvTc* c, const float* dy, const float* x, float* dw, cudaStream_t s) {
    const ConvGeom& g = c->g;
    const int RS = g.R * g.S, Kp = c->Kp, Cp = c->Cp;
    size_t yb = align_up((size_t)g.N * g.P * g.Q * Kp * 2, 1024), xb = align_up((size_t)g.N * g.H * g.W * Cp * 2, 1024);
    const size_t sb = align_up((size_t)RS * g.K * g.C * sizeof(float), 1024);   // [tap][C][K] accumulation scratch
    // with both operands staged by the plan and a private accumulator the shared arena is not touched at all -- which is what
    // lets the plan run this op on its side stream next to other convolutions (conv_tc_side_stream_safe)
    const bool own = c->pre[0] && c->pre[1] && c->acc_private;
    uint8_t* st = own ? nullptr : stage_get(yb + xb + (c->acc_private ? 0 : sb));
    float* acc = c->acc_private ? c->acc_private : (float*)(st + yb + xb);
    auto* dyh = (__nv_bfloat16*)st;
    auto* xh = (__nv_bfloat16*)(st + yb);
    if (c->pre[0]) dyh = (__nv_bfloat16*)c->pre[0];
    else nchw_to_nhwc_bf16(dy, dyh, g.N, g.K, g.P * g.Q, Kp, s);
    if (c->pre[1]) xh = (__nv_bfloat16*)c->pre[1];
    else nchw_to_nhwc_bf16(x, xh, g.N, g.C, g.H * g.W, Cp, s);
    if (!c->acc_external) {
        DB_CUDA(cudaMemsetAsync(acc, 0, (size_t)g.K * g.C * RS * sizeof(float), s));
        count_launch();
    }
    const PixelBox& b = c->box;
    CUtensorMap tmA, tmB;
    make_map_nhwc(&tmA, dyh, g.N, g.P, g.Q, Kp, g.K, b.bn, b.bh, b.bw, 1, 1, "convolutionFiltersGrad dy");
    make_map_nhwc(&tmB, xh, g.N, g.H, g.W, Cp, g.C, b.bn, b.bh, b.bw, g.u, g.v, "convolutionFiltersGrad x");
    TcArgs a{};
    a.mode = TC_MODE_WGRAD;
    a.BN = pick_bn(g.C);
    a.n_tiles = (int)ceil_div(g.C, a.BN);
    a.M = g.K; a.N = g.C; a.Nout = g.C;
    a.taps = RS;
    for (int r = 0; r < g.R; ++r)
        for (int q = 0; q < g.S; ++q) {
            int t = r * g.S + q;
            a.tap_dh[t] = r - g.ph;
            a.tap_dw[t] = q - g.pw;
            a.tap_bcol[t] = t * g.C * g.K;   // this tap's [C][K] plane of the scratch
        }
    a.bn = b.bn; a.bh = b.bh; a.bw = b.bw;
    a.tiles_p = b.tiles_p; a.tiles_q = b.tiles_q;
    a.a_su = g.u; a.a_sv = g.v;
    a.kmma = b.bn * b.bh * b.bw / 16;
    a.pix_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    a.out_kind = TC_OUT_F32_ATOMIC;
    a.o_off = 0;
    // scratch[tap][c][k]: the accumulator row (TMEM lane) is k, so the 32 atomics of a warp instruction hit 32 consecutive
    // floats -- one L2 line instead of 32 (the strided KCRS version was bound by L2 atomic throughput)
    a.o_sn = 1;        // per Kout row
    a.o_sc = g.K;      // per Cin column
    a.out = acc;
    int mt = (int)ceil_div(g.K, TC_BM);
    // Halo variant (3 x 3, unit stride, one image per pixel box): an item = (filter COLUMN, one Kout tile, Cin tile, pixel split);
    // its three vertical taps keep their accumulators side by side in TMEM and share both the dy tile and ONE x box that
    // carries a halo row above and below.  L2 -> SM bytes per MAC drop by a third (profiles/r02_summary.md).
    static const int wg_halo = getenv("DOPT_B200_WG_HALO") ? atoi(getenv("DOPT_B200_WG_HALO")) : 1;
    if (wg_halo && g.R == 3 && g.S == 3 && g.u == 1 && g.v == 1 && b.bn == 1 && b.bw % 8 == 0 && b.bw == g.Q && 3 * a.BN <= 512 &&
        !getenv("DOPT_B200_DBG")) {
        uint64_t dims[4] = {(uint64_t)g.C, (uint64_t)g.W, (uint64_t)g.H, (uint64_t)g.N};
        uint64_t st[3] = {(uint64_t)Cp * 2, (uint64_t)g.W * Cp * 2, (uint64_t)g.H * g.W * Cp * 2};
        uint32_t box[4] = {64, (uint32_t)b.bw, (uint32_t)(b.bh + 2), 1};
        uint32_t es[4] = {1, 1, 1, 1};
        if (encode_tiled(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, xh, dims, st, box, es, CU_TENSOR_MAP_SWIZZLE_128B) == CUDA_SUCCESS)
            a.halo = 1;
    }
    if (a.halo) {
        a.wg_nm = 1;
        const int sms = sm_count();
        const double pix = a.kmma * 16.0, nblk = (a.BN + 63) / 64, pixh = (double)(b.bh + 2) * b.bw;
        const double mma = 3.0 * a.kmma * 115.0 * a.BN / 160.0;
        const double copy = (2.0 * pix + nblk * pixh) * 128.0 / 42.0;   // ~42 B/clk per SM out of L2 with every SM pulling
        const double iter = std::max(mma, copy) + 100.0, fixed = 4000.0 + 3.0 * 4000.0 * a.BN / 160.0;
        const int64_t tiles_ = (int64_t)g.S * mt * a.n_tiles;
        double best_t = 1e300;
        int best_sp = 1;
        for (int sp = 1; sp <= a.pix_tiles && tiles_ * sp <= (int64_t)16 * sms; ++sp) {
            const int64_t per = ceil_div((int64_t)a.pix_tiles, (int64_t)sp);
            if (per < 4 && sp > 1) break;
            const int64_t rounds = ceil_div(tiles_ * sp, (int64_t)sms);
            const double t = (double)rounds * ((double)per * iter + fixed);
            if (t < best_t * 0.98) {
                best_t = t;
                best_sp = sp;
            }
        }
        a.splits = best_sp;
        if (const char* e = getenv("DOPT_B200_WG_SPLITS")) a.splits = std::max(1, std::min(a.pix_tiles, atoi(e)));
        a.stages = pick_stages(a);
        a.m_tiles = g.S * mt;
        a.cluster = 1;
        tc_launch<TC_MODE_WGRAD>(tmA, tmB, a, (int)tiles_ * a.splits, s);
    } else {
    // Work decomposition.  An item = (tap, group of wg_nm Kout tiles, Cin tile, pixel split).  The Kout tiles of a group share
    // the x tile of every stage, which is what keeps the kernel off the L2 -> SM bandwidth limit
    // (profiles/r01c_conv_bisect.md); wg_nm is bounded by the 512 TMEM columns and the shared-memory ring.  The pixel range
    // is split so that the items fill whole rounds of the persistent grid (one CTA per SM).  Both are chosen with a small
    // cost model: time ~ rounds * (iterations * max(MMA, copy) + per-item epilogue).
    int nm_max = std::max(1, std::min(3, std::min(mt, 512 / a.BN)));
    if (const char* e = getenv("DOPT_B200_WG_NM")) nm_max = std::max(1, std::min(atoi(e), nm_max));
    const int sms = sm_count();
    const double pix = a.kmma * 16.0, nblk = (a.BN + 63) / 64;
    double best_t = 1e300;
    int best_nm = 1, best_sp = 1;
    for (int nm = 1; nm <= nm_max; ++nm) {
        a.wg_nm = nm;
        a.stages = 2;
        if (nm > 1 && 3 * tc_smem_layout(a).stage_bytes > 223 * 1024) break;   // keep the ring at least 3 deep
        const int mg_ = (int)ceil_div(mt, nm);
        const double nm_eff = (double)mt / mg_;                                  // average Kout tiles per item
        const double mma = nm_eff * a.kmma * 115.0 * a.BN / 160.0;               // cycles per k-iteration
        const double copy = (2.0 * nm_eff + nblk) * pix * 128.0 / 50.0;          // ~50 B/clk per SM from L2
        const double iter = std::max(mma, copy) + 100.0;
        const double fixed = 4000.0 + 4000.0 * nm_eff;
        const int64_t tiles_ = (int64_t)RS * mg_ * a.n_tiles;
        for (int sp = 1; sp <= a.pix_tiles && tiles_ * sp <= (int64_t)16 * sms; ++sp) {
            const int64_t per = ceil_div((int64_t)a.pix_tiles, (int64_t)sp);
            if (per < 4 && sp > 1) break;
            const int64_t rounds = ceil_div(tiles_ * sp, (int64_t)sms);
            const double t = (double)rounds * ((double)per * iter + fixed);
            if (t < best_t * 0.98) {
                best_t = t;
                best_nm = nm;
                best_sp = sp;
            }
        }
    }
    a.wg_nm = best_nm;
    a.splits = best_sp;
    const int mg = (int)ceil_div(mt, a.wg_nm);
    int tiles = RS * mg * a.n_tiles;
    if (const char* e = getenv("DOPT_B200_WG_SPLITS")) a.splits = std::max(1, std::min(a.pix_tiles, atoi(e)));
    a.stages = pick_stages(a);
    a.m_tiles = RS * mg;
    a.cluster = 1;
    if (const char* e = getenv("DOPT_B200_DBG")) a.dbg = atoi(e);
    tc_launch<TC_MODE_WGRAD>(tmA, tmB, a, tiles * a.splits, s);
    }
    if (!c->acc_private) {
        const size_t smem = (size_t)32 * (32 * RS + 1) * sizeof(float);
        DB_REQUIRE(smem <= 200 * 1024, "filter window too large for the wgrad finish kernel");
        static size_t configured = 48 * 1024;
        if (smem > configured) {
            DB_CUDA(cudaFuncSetAttribute(wgrad_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            configured = 200 * 1024;
        }
        dim3 grid((unsigned)ceil_div(g.K, 32), (unsigned)ceil_div(g.C, 32));
        wgrad_finish_kernel<<<grid, 256, smem, s>>>(acc, dw, g.K, g.C, RS);
        DB_LAUNCH_CHECK();
    }
}

void conv_tc_run(ConvTc* c, const float* a, const float* b, float* out, cudaStream_t s) {
    if (c->kind == CONV_FWD) run_fwd(c, a, b, out, s);
    else if (c->kind == CONV_DGRAD) run_dgrad(c, a, b, out, s);
    else run_wgrad(c, a, b, out, s);
}

}  // namespace db
