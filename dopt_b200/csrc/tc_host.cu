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
#include "tc_kernel.cuh"
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
// grid (ceil(HW/32), ceil(Cp/64), N), 256 threads.  4 B read + 2 B written per element.
__global__ void __launch_bounds__(256) nchw_to_nhwc_bf16_kernel(const float* __restrict__ in,
                                                                __nv_bfloat16* __restrict__ out, int C, int HW, int Cp) {
    __shared__ float tile[64][33];
    const int n = blockIdx.z;
    const int hw0 = blockIdx.x * 32, c0 = blockIdx.y * 64;
    const float* src = in + (int64_t)n * C * HW;
    __nv_bfloat16* dst = out + (int64_t)n * HW * Cp;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
#pragma unroll
    for (int j = 0; j < 64; j += 8) {
        int c = c0 + ty + j, hw = hw0 + tx;
        tile[ty + j][tx] = (c < C && hw < HW) ? src[(int64_t)c * HW + hw] : 0.f;
    }
    __syncthreads();
    // write: each thread packs 2 channels; 32 threads cover 64 channels of one pixel (128 contiguous bytes)
#pragma unroll
    for (int j = 0; j < 32; j += 8) {
        int hw = hw0 + ty + j;
        int c = c0 + tx * 2;
        if (hw < HW && c < Cp) {
            __nv_bfloat162 v = __floats2bfloat162_rn(tile[tx * 2][ty + j], tile[tx * 2 + 1][ty + j]);
            *(__nv_bfloat162*)(dst + (int64_t)hw * Cp + c) = v;
        }
    }
}

// KCRS fp32 -> packed bf16.  MODE 0 (fwd): out[k][(r*S+s)*Cp + c] = w[k][c][R-1-r][S-1-s]
//                            MODE 1 (dgrad): out[c][(r*S+s)*Kp + k] = w[k][c][R-1-r][S-1-s]
__global__ void __launch_bounds__(256) pack_filters_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out,
                                                           int K, int C, int R, int S, int Kp, int Cp, int mode) {
    const int RS = R * S;
    int64_t n = mode == 0 ? (int64_t)K * RS * Cp : (int64_t)C * RS * Kp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        int k, c, t;
        if (mode == 0) {
            c = (int)(i % Cp);
            int64_t j = i / Cp;
            t = (int)(j % RS);
            k = (int)(j / RS);
        } else {
            k = (int)(i % Kp);
            int64_t j = i / Kp;
            t = (int)(j % RS);
            c = (int)(j / RS);
        }
        int r = t / S, s = t - r * S;
        float v = 0.f;
        if (k < K && c < C) v = w[(((int64_t)k * C + c) * R + (R - 1 - r)) * S + (S - 1 - s)];
        out[i] = __float2bfloat16_rn(v);
    }
}

static void nchw_to_nhwc_bf16(const float* in, __nv_bfloat16* out, int N, int C, int HW, int Cp, cudaStream_t s) {
    dim3 grid((unsigned)ceil_div(HW, 32), (unsigned)ceil_div(Cp, 64), (unsigned)N);
    nchw_to_nhwc_bf16_kernel<<<grid, 256, 0, s>>>(in, out, C, HW, Cp);
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
static double g_tc_prof_us = 0;
static int64_t g_tc_prof_launches = 0;
static cudaEvent_t g_tc_e0 = nullptr, g_tc_e1 = nullptr;
void tc_prof_enable(bool on) {
    g_tc_prof = on;
    g_tc_prof_us = 0;
    g_tc_prof_launches = 0;
    if (on && !g_tc_e0) {
        DB_CUDA(cudaEventCreate(&g_tc_e0));
        DB_CUDA(cudaEventCreate(&g_tc_e1));
    }
}
void tc_prof_read(double* us, int64_t* launches) {
    *us = g_tc_prof_us;
    *launches = g_tc_prof_launches;
}

template <int MODE>
static void tc_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcArgs& a, int n_ctas, cudaStream_t s) {
    TcSmemLayout L = tc_smem_layout(a);
    static int configured = 0;
    if (configured < (int)L.total) {
        DB_CUDA(cudaFuncSetAttribute(tc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        configured = 227 * 1024;
    }
    DB_REQUIRE(L.total <= 227 * 1024, "tcgen05 kernel: shared memory budget exceeded");
    if (g_tc_prof) DB_CUDA(cudaEventRecord(g_tc_e0, s));
    if (a.cluster > 1 || a.pair) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((unsigned)n_ctas);
        cfg.blockDim = dim3(TC_THREADS);
        cfg.dynamicSmemBytes = L.total;
        cfg.stream = s;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = (unsigned)(a.pair ? 2 : a.cluster);
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        if (a.pair) {
            static bool pair_configured = false;
            if (!pair_configured) {
                DB_CUDA(cudaFuncSetAttribute(tc_kernel<TC_MODE_CONV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
                pair_configured = true;
            }
            DB_CUDA(cudaLaunchKernelEx(&cfg, tc_kernel<TC_MODE_CONV, true>, tmA, tmB, a));
        } else {
            DB_CUDA(cudaLaunchKernelEx(&cfg, tc_kernel<MODE>, tmA, tmB, a));
        }
    } else {
        tc_kernel<MODE><<<n_ctas, TC_THREADS, L.total, s>>>(tmA, tmB, a);
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
    if (g_tc_prof) {
        DB_CUDA(cudaEventRecord(g_tc_e1, s));
        DB_CUDA(cudaEventSynchronize(g_tc_e1));
        float ms = 0;
        DB_CUDA(cudaEventElapsedTime(&ms, g_tc_e0, g_tc_e1));
        g_tc_prof_us += ms * 1000.0;
        ++g_tc_prof_launches;
    }
}

static int pick_stages(TcArgs& a) {
    a.stages = 2;
    TcSmemLayout L = tc_smem_layout(a);
    int st2 = (int)((110 * 1024 - 2048) / L.stage_bytes);   // two CTAs per SM
    int st1 = (int)((225 * 1024 - 2048) / L.stage_bytes);   // one CTA per SM
    int st = st2 >= 3 ? st2 : st1;
    if (st > 8) st = 8;
    if (st < 2) st = 2;
    return st;
}

// CTAs per cluster that share one filter tile through TMA multicast.  The implicit-GEMM kernel is bound by L2->SM
// bandwidth (profiles/r01a_tc_kernel.md): every CTA of a 128 x BN tile pulls the full BN x 64 filter slab per k-step.
// With a cluster of c CTAs working on c different pixel tiles, each CTA pulls 1/c of it.
static int g_max_cluster = 4;
void tc_set_max_cluster(int c) { g_max_cluster = c; }
static int pick_cluster(int m_tiles, int bn) {
    static bool env_read = false;
    if (!env_read) {
        env_read = true;
        if (const char* e = getenv("DOPT_B200_MAX_CLUSTER")) g_max_cluster = atoi(e);   // tuning knob for experiments
    }
    for (int c = g_max_cluster; c > 1; c >>= 1)
        if (m_tiles >= 2 * c && bn % (8 * c) == 0) return c;
    return 1;
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
    // worth a 128-row tensor-core tile only when the product is big enough; everything else is latency-bound anyway
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
static bool pick_box(int N, int P, int Q, bool mult16, PixelBox& b) {
    int bw = Q <= 128 ? Q : 128;
    int best = 0;
    for (int bh = std::min(P, 128 / bw); bh >= 1; --bh) {
        int maxn = (bh == P) ? std::min(N, 128 / (bw * bh)) : 1;
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
};

void stage_nchw_to_nhwc_bf16(const float* in, void* out, int N, int C, int64_t HW, cudaStream_t s) {
    nchw_to_nhwc_bf16(in, (__nv_bfloat16*)out, N, C, (int)HW, (int)align_up(C, 8), s);
}
size_t staged_nhwc_bytes(int N, int C, int64_t HW) { return align_up((size_t)N * HW * align_up(C, 8) * 2, 1024); }
void conv_tc_set_staged(ConvTc* c, int input, const void* p) {
    if (c && input >= 0 && input < 2) c->pre[input] = p;
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
    if (g.C < 16 || g.K < 16) return false;   // tiny-channel layers (the 3-channel stem) are bandwidth-bound: fp32 direct kernel
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
void conv_tc_destroy(ConvTc* c) { delete c; }

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
    pack_filters_kernel<<<stream_grid((int64_t)g.K * RS * Cp, 256, 8), 256, 0, s>>>(w, wp, g.K, g.C, g.R, g.S, c->Kp, Cp, 0);
    DB_LAUNCH_CHECK();
    const PixelBox& b = c->box;
    CUtensorMap tmA, tmB;
    make_map_nhwc(&tmA, xh, g.N, g.H, g.W, Cp, g.C, b.bn, b.bh, b.bw, g.u, g.v, "convolution x");
    TcArgs a{};
    a.mode = TC_MODE_CONV;
    a.BN = pick_bn(g.K);
    a.m_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    a.pair = pick_pair(a.m_tiles, a.BN) ? 1 : 0;
    a.cluster = a.pair ? 1 : pick_cluster(a.m_tiles, a.BN);
    a.m_tiles = (int)align_up(a.m_tiles, a.pair ? 2 : a.cluster);
    make_map_2d(&tmB, wp, (uint64_t)RS * Cp, (uint64_t)g.K, (uint64_t)RS * Cp, 64,
                (uint32_t)(a.pair ? a.BN / 2 : a.BN / a.cluster), "convolution w");
    a.n_tiles = (int)ceil_div(g.K, a.BN);
    a.splits = 1;
    a.taps = RS;
    a.c_iters = (int)ceil_div(g.C, TC_BK);
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
    a.stages = pick_stages(a);
    tc_launch<TC_MODE_CONV>(tmA, tmB, a, a.m_tiles * a.n_tiles, s);
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
    pack_filters_kernel<<<stream_grid((int64_t)g.C * RS * Kp, 256, 8), 256, 0, s>>>(w, wp, g.K, g.C, g.R, g.S, Kp, c->Cp, 1);
    DB_LAUNCH_CHECK();
    const PixelBox& b = c->box;
    CUtensorMap tmA, tmB;
    make_map_nhwc(&tmA, dyh, g.N, g.P, g.Q, Kp, g.K, b.bn, b.bh, b.bw, 1, 1, "convolutionFeaturesGrad dy");
    int BN = pick_bn(g.C);
    int m_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    int pair = pick_pair(m_tiles, BN) ? 1 : 0;
    int cluster = pair ? 1 : pick_cluster(m_tiles, BN);
    m_tiles = (int)align_up(m_tiles, pair ? 2 : cluster);
    make_map_2d(&tmB, wp, (uint64_t)RS * Kp, (uint64_t)g.C, (uint64_t)RS * Kp, 64,
                (uint32_t)(pair ? BN / 2 : BN / cluster), "convolutionFeaturesGrad w");
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
            a.stages = pick_stages(a);
            launches.push_back(a);
        }
    if (need_zero) {
        DB_CUDA(cudaMemsetAsync(dx, 0, (size_t)g.N * g.C * g.H * g.W * sizeof(float), s));
        count_launch();
    }
    for (auto& a : launches) tc_launch<TC_MODE_CONV>(tmA, tmB, a, a.m_tiles * a.n_tiles, s);
}

static void run_wgrad(ConvTc* c, const float* dy, const float* x, float* dw, cudaStream_t s) {
    const ConvGeom& g = c->g;
    const int RS = g.R * g.S, Kp = c->Kp, Cp = c->Cp;
    size_t yb = align_up((size_t)g.N * g.P * g.Q * Kp * 2, 1024), xb = align_up((size_t)g.N * g.H * g.W * Cp * 2, 1024);
    uint8_t* st = stage_get(yb + xb);
    auto* dyh = (__nv_bfloat16*)st;
    auto* xh = (__nv_bfloat16*)(st + yb);
    if (c->pre[0]) dyh = (__nv_bfloat16*)c->pre[0];
    else nchw_to_nhwc_bf16(dy, dyh, g.N, g.K, g.P * g.Q, Kp, s);
    if (c->pre[1]) xh = (__nv_bfloat16*)c->pre[1];
    else nchw_to_nhwc_bf16(x, xh, g.N, g.C, g.H * g.W, Cp, s);
    DB_CUDA(cudaMemsetAsync(dw, 0, (size_t)g.K * g.C * RS * sizeof(float), s));
    count_launch();
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
            a.tap_bcol[t] = (g.R - 1 - r) * g.S + (g.S - 1 - q);   // flipped position inside the KCRS filter
        }
    a.bn = b.bn; a.bh = b.bh; a.bw = b.bw;
    a.tiles_p = b.tiles_p; a.tiles_q = b.tiles_q;
    a.a_su = g.u; a.a_sv = g.v;
    a.kmma = b.bn * b.bh * b.bw / 16;
    a.pix_tiles = b.tiles_n * b.tiles_p * b.tiles_q;
    a.out_kind = TC_OUT_F32_ATOMIC;
    a.o_off = 0;
    a.o_sn = (long long)g.C * RS;   // per Kout row
    a.o_sc = RS;                    // per Cin column
    a.out = dw;
    int mt = (int)ceil_div(g.K, TC_BM);
    int tiles = RS * mt * a.n_tiles;
    int want = 2 * sm_count();
    a.splits = std::max(1, std::min(a.pix_tiles, (int)ceil_div(want, tiles)));
    a.stages = pick_stages(a);
    a.m_tiles = RS * mt;
    a.cluster = 1;
    tc_launch<TC_MODE_WGRAD>(tmA, tmB, a, tiles * a.splits, s);
}

void conv_tc_run(ConvTc* c, const float* a, const float* b, float* out, cudaStream_t s) {
    if (c->kind == CONV_FWD) run_fwd(c, a, b, out, s);
    else if (c->kind == CONV_DGRAD) run_dgrad(c, a, b, out, s);
    else run_wgrad(c, a, b, out, s);
}

}  // namespace db
