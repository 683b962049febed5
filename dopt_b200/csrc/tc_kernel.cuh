// tc_kernel.cuh -- the one warp-specialised tcgen05 kernel behind matmul and the three convolution ops.
//
//   warp 0   TMA producer   : fills a ring of shared-memory stages (A tile + B tile) with cp.async.bulk.tensor
//   warp 1   MMA issuer     : one elected thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM), tcgen05.commit frees stages
//   warp 2-5 epilogue       : tcgen05.ld the 128 x BN fp32 accumulator out of TMEM and store it
//
// One CTA computes one 128 x BN output tile (optionally one K-split of it).  Up to two CTAs are resident per SM, so the
// epilogue of one overlaps the main loop of the other.  Operands are bf16 in 128-byte-swizzled tiles:
//   MODE_GEMM   A [M][K] K-major (2D map)            B [K][N] N-major (2D map, MN-major descriptor)
//   MODE_CONV   A NHWC activations (4D map, one filter tap per k-iteration, out-of-bounds = zero padding,
//               elementStrides = conv stride)         B [Kout][taps*Cin] K-major (2D map)
//   MODE_WGRAD  A dy NHWC (4D map) as [pixels][Kout] MN-major     B x NHWC (4D map, tap-shifted) as [pixels][Cin] MN-major
#pragma once
#include "tc_common.cuh"
#include <cuda_bf16.h>

namespace db {

enum { TC_MODE_GEMM = 0, TC_MODE_CONV = 1, TC_MODE_WGRAD = 2 };
enum { TC_OUT_F32 = 0, TC_OUT_BF16 = 1, TC_OUT_F32_ATOMIC = 2 };

static constexpr int TC_MAX_TAPS = 25;
static constexpr int TC_BM = 128;
static constexpr int TC_BK = 64;                 // bf16 elements per 128-byte swizzled row
static constexpr int TC_THREADS = 192;

struct TcArgs {
    int mode;
    int BN;             // accumulator columns (multiple of 16, <= 256)
    int stages;
    int n_tiles;        // tiles along N; blockIdx.x = ((n_tile * splits) + split) * m_tiles + m_tile
    int m_tiles;        // tiles along M (padded to a multiple of `cluster`)
    int cluster;        // CTAs per cluster sharing one B tile through TMA multicast (1, 2 or 4; CONV mode)
    int pair;           // CONV mode: cta_group::2 -- two CTAs (one TPC) form a 256 x BN tile, each holding half of the B tile
    int splits;
    int k_iters;        // total k-iterations of the problem (divided over splits)
    // ---- GEMM
    int M, N;
    // ---- CONV / WGRAD geometry
    int taps;
    int tap_dh[TC_MAX_TAPS], tap_dw[TC_MAX_TAPS];   // A coordinate offset of the tap
    int tap_bcol[TC_MAX_TAPS];                      // CONV: first B column of the tap's slab
    int c_iters;        // CONV: 64-channel blocks per tap
    int bn, bh, bw;     // pixel box of one tile: images x rows x cols (bn*bh*bw <= 128)
    int tiles_p, tiles_q;   // tile grid inside one image group: m_tile -> (ng, tp, tq)
    int a_su, a_sv;     // A pixel coordinate = out pixel * stride + tap offset
    int NI, OP, OQ;     // valid extents of the logical output pixel grid (images, rows, cols)
    int Nout;           // valid output channels
    // ---- output addressing: element index = o_off + n*o_sn + c*o_sc + p*o_sh + q*o_sw
    int out_kind;
    long long o_off, o_sn, o_sc, o_sh, o_sw;
    void* out;
    // ---- WGRAD: out index = o_off + m*o_sn + n*o_sc (m = Kout row, n = Cin col); pixel chunks
    int wg_tap;         // unused (tap comes from the tile index)
    int pix_tiles;      // number of pixel boxes (the reduction dimension), split over `splits`
    int kmma;           // MMAs per stage (box pixels / 16)
};

struct TcSmemLayout {
    uint32_t a_bytes, b_bytes, stage_bytes, bar_off, total;
};
__host__ __device__ inline TcSmemLayout tc_smem_layout(const TcArgs& a) {
    TcSmemLayout L;
    if (a.mode == TC_MODE_WGRAD) {
        L.a_bytes = 2u * (uint32_t)(a.kmma * 16) * 128u;                 // two 64-wide Kout blocks
        L.b_bytes = (uint32_t)((a.BN + 63) / 64) * (uint32_t)(a.kmma * 16) * 128u;
    } else if (a.mode == TC_MODE_GEMM) {
        L.a_bytes = TC_BM * 128u;
        L.b_bytes = (uint32_t)((a.BN + 63) / 64) * 64u * 128u;           // [n-block][64 k rows][64 n]
    } else {
        L.a_bytes = TC_BM * 128u;
        L.b_bytes = (uint32_t)(a.pair ? a.BN / 2 : a.BN) * 128u;
    }
    L.b_bytes = (L.b_bytes + 1023u) & ~1023u;
    L.stage_bytes = L.a_bytes + L.b_bytes;
    L.bar_off = L.stage_bytes * (uint32_t)a.stages;
    L.total = L.bar_off + 1024u /* barriers + tmem slot */ + 1024u /* alignment slack */;
    return L;
}

// PAIR = cta_group::2 flavour (CONV mode only).  It is a separate instantiation because a kernel that contains cta_group::2
// instructions can only be launched with an even cluster size.
template <int MODE, bool PAIR = false>
__global__ void __launch_bounds__(TC_THREADS) tc_kernel(const __grid_constant__ CUtensorMap tmA,
                                                        const __grid_constant__ CUtensorMap tmB,
                                                        const __grid_constant__ TcArgs args) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const TcSmemLayout L = tc_smem_layout(args);
    uint64_t* full_bar = (uint64_t*)(smem + L.bar_off);
    uint64_t* empty_bar = full_bar + 16;
    uint64_t* tmem_full_bar = empty_bar + 16;
    uint32_t* tmem_slot = (uint32_t*)(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int BN = args.BN;
    const int stages = args.stages;

    // ---- tile coordinates
    int bid = blockIdx.x;
    const int m_tile = bid % args.m_tiles;   // fastest: the CTAs of one cluster differ only in their M tile
    bid /= args.m_tiles;
    const int split = bid % args.splits;
    const int n_tile = bid / args.splits;
    constexpr bool pair = PAIR && (MODE == TC_MODE_CONV);
    const int csize = pair ? 2 : ((MODE == TC_MODE_CONV) ? args.cluster : 1);
    const uint32_t crank = csize > 1 ? tcg::cluster_ctarank() : 0;
    const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
    int it_begin, it_end;
    {
        int total = (MODE == TC_MODE_WGRAD) ? args.pix_tiles : args.k_iters;
        int per = (total + args.splits - 1) / args.splits;
        it_begin = split * per;
        it_end = min(total, it_begin + per);
    }
    const int n_iters = max(0, it_end - it_begin);

    // conv tile -> pixel box origin
    int img0 = 0, p0 = 0, q0 = 0, wg_tap = 0;
    if (MODE == TC_MODE_CONV) {
        int tq = m_tile % args.tiles_q;
        int t2 = m_tile / args.tiles_q;
        int tp = t2 % args.tiles_p;
        int ng = t2 / args.tiles_p;
        img0 = ng * args.bn;
        p0 = tp * args.bh;
        q0 = tq * args.bw;
    }
    if (MODE == TC_MODE_WGRAD) {
        // m_tile enumerates (tap, Kout tile); Kout tiles = ceil(M / 128)
        int mt = (args.M + TC_BM - 1) / TC_BM;
        wg_tap = m_tile / mt;
    }

    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < BN) tmem_cols <<= 1;

    if (warp == 0 && lane == 0) {
        tcg::tma_prefetch_desc(&tmA);
        tcg::tma_prefetch_desc(&tmB);
        for (int i = 0; i < stages; ++i) {
            tcg::mbar_init(&full_bar[i], 1);
            // multicast: every CTA that receives the data must release the stage; pair: the leader's commit releases both
            tcg::mbar_init(&empty_bar[i], pair ? 1u : (uint32_t)csize);
        }
        tcg::mbar_init(tmem_full_bar, 1);
        tcg::fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (pair) {
            tcg::tmem_alloc_2sm(tmem_slot, tmem_cols);
            tcg::tmem_relinquish_2sm();
        } else {
            tcg::tmem_alloc(tmem_slot, tmem_cols);
            tcg::tmem_relinquish();
        }
    }
    tcg::tc_fence_before();
    if (csize > 1) tcg::cluster_sync();   // peers' barriers must exist before the first remote arrive / multicast write
    else __syncthreads();
    tcg::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (int i = 0; i < n_iters; ++i) {
                const int it = it_begin + i;
                const int st = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                tcg::mbar_wait(&empty_bar[st], ph ^ 1u);
                uint8_t* sa = smem + (size_t)st * L.stage_bytes;
                uint8_t* sb = sa + L.a_bytes;
                if (MODE == TC_MODE_GEMM) {
                    const int nblk = (BN + 63) / 64;
                    tcg::mbar_arrive_expect_tx(&full_bar[st], TC_BM * 128u + (uint32_t)nblk * 8192u);
                    tcg::tma_load_2d(sa, &tmA, &full_bar[st], it * TC_BK, m_tile * TC_BM);
                    for (int b = 0; b < nblk; ++b)
                        tcg::tma_load_2d(sb + b * 8192, &tmB, &full_bar[st], n_tile * BN + b * 64, it * TC_BK);
                } else if constexpr (pair) {
                    // both CTAs of the pair load their own activation tile and their half of the filter tile; all bytes
                    // are credited to the leader's barrier, which the leader arms for the pair
                    const int t = it / args.c_iters, cb = it - t * args.c_iters;
                    const uint32_t a_bytes = (uint32_t)(args.bn * args.bh * args.bw) * 128u;
                    const int rows = BN / 2;
                    if (crank == 0) tcg::mbar_arrive_expect_tx(&full_bar[st], 2u * (a_bytes + (uint32_t)rows * 128u));
                    tcg::tma_load_4d_2sm(sa, &tmA, &full_bar[st], cb * TC_BK, q0 * args.a_sv + args.tap_dw[t],
                                         p0 * args.a_su + args.tap_dh[t], img0);
                    tcg::tma_load_2d_2sm(sb, &tmB, &full_bar[st], args.tap_bcol[t] + cb * TC_BK,
                                         n_tile * BN + (int)crank * rows);
                } else if (MODE == TC_MODE_CONV) {
                    const int t = it / args.c_iters, cb = it - t * args.c_iters;
                    const uint32_t a_bytes = (uint32_t)(args.bn * args.bh * args.bw) * 128u;
                    tcg::mbar_arrive_expect_tx(&full_bar[st], a_bytes + (uint32_t)BN * 128u);
                    tcg::tma_load_4d(sa, &tmA, &full_bar[st], cb * TC_BK, q0 * args.a_sv + args.tap_dw[t],
                                     p0 * args.a_su + args.tap_dh[t], img0);
                    if (csize == 1) {
                        tcg::tma_load_2d(sb, &tmB, &full_bar[st], args.tap_bcol[t] + cb * TC_BK, n_tile * BN);
                    } else {
                        // this CTA fetches 1/csize of the filter tile and multicasts it to the whole cluster
                        const int rows = BN / csize;
                        tcg::tma_load_2d_mcast(sb + (size_t)crank * rows * 128, &tmB, &full_bar[st],
                                               args.tap_bcol[t] + cb * TC_BK, n_tile * BN + (int)crank * rows, cmask);
                    }
                } else {
                    // WGRAD: `it` is a pixel box index -> (image group, row tile, col tile)
                    int tq = it % args.tiles_q;
                    int t2 = it / args.tiles_q;
                    int tp = t2 % args.tiles_p;
                    int ng = t2 / args.tiles_p;
                    const int pix = args.kmma * 16;
                    const int mt = (args.M + TC_BM - 1) / TC_BM;
                    const int m_blk = m_tile % mt;
                    const int nblk = (BN + 63) / 64;
                    tcg::mbar_arrive_expect_tx(&full_bar[st], (uint32_t)(2 + nblk) * (uint32_t)pix * 128u);
                    // A: dy box, two 64-channel column blocks of Kout
                    for (int b = 0; b < 2; ++b)
                        tcg::tma_load_4d(sa + (size_t)b * pix * 128, &tmA, &full_bar[st], m_blk * TC_BM + b * 64,
                                         tq * args.bw, tp * args.bh, ng * args.bn);
                    // B: x box shifted by the tap, conv stride as element stride
                    for (int b = 0; b < nblk; ++b)
                        tcg::tma_load_4d(sb + (size_t)b * pix * 128, &tmB, &full_bar[st], n_tile * BN + b * 64,
                                         tq * args.bw * args.a_sv + args.tap_dw[wg_tap],
                                         tp * args.bh * args.a_su + args.tap_dh[wg_tap], ng * args.bn);
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if constexpr (pair) {
            if (lane == 0 && n_iters > 0 && crank == 0) {
                const uint32_t idesc2 = tcg::make_idesc_bf16(2 * TC_BM, BN, 0, 0);
                for (int i = 0; i < n_iters; ++i) {
                    const int st = i % stages;
                    const uint32_t ph = (uint32_t)(i / stages) & 1u;
                    tcg::mbar_wait(&full_bar[st], ph);
                    tcg::tc_fence_after();
                    const uint32_t sa = tcg::smem_u32(smem + (size_t)st * L.stage_bytes);
                    const uint32_t sb = sa + L.a_bytes;
                    const uint64_t da = tcg::make_smem_desc(sa, 16, 1024, 2);
                    const uint64_t dbb = tcg::make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k)
                        tcg::umma_bf16_2sm(tmem_base, da + (uint64_t)(k * 2), dbb + (uint64_t)(k * 2), idesc2,
                                           (uint32_t)((i | k) != 0));
                    tcg::umma_commit_2sm(&empty_bar[st], 3);   // releases the stage in both CTAs
                }
                tcg::umma_commit_2sm(tmem_full_bar, 3);        // both halves of the accumulator are complete
            }
        } else if (lane == 0 && n_iters > 0) {
            const uint32_t idesc = tcg::make_idesc_bf16(TC_BM, BN, MODE == TC_MODE_WGRAD ? 1 : 0,
                                                        (MODE == TC_MODE_CONV) ? 0 : 1);
            for (int i = 0; i < n_iters; ++i) {
                const int st = i % stages;
                const uint32_t ph = (uint32_t)(i / stages) & 1u;
                tcg::mbar_wait(&full_bar[st], ph);
                tcg::tc_fence_after();
                const uint32_t sa = tcg::smem_u32(smem + (size_t)st * L.stage_bytes);
                const uint32_t sb = sa + L.a_bytes;
                if (MODE == TC_MODE_WGRAD) {
                    const uint32_t pix_bytes = (uint32_t)args.kmma * 16u * 128u;
                    const uint64_t da = tcg::make_smem_desc(sa, pix_bytes, 1024, 2);
                    const uint64_t dbb = tcg::make_smem_desc(sb, pix_bytes, 1024, 2);
                    for (int k = 0; k < args.kmma; ++k)
                        tcg::umma_bf16(tmem_base, da + (uint64_t)(k * 128), dbb + (uint64_t)(k * 128), idesc,
                                       (uint32_t)((i | k) != 0));
                } else {
                    const uint64_t da = tcg::make_smem_desc(sa, 16, 1024, 2);
                    const uint64_t dbb = (MODE == TC_MODE_GEMM) ? tcg::make_smem_desc(sb, 8192, 1024, 2)
                                                                : tcg::make_smem_desc(sb, 16, 1024, 2);
#pragma unroll
                    for (int k = 0; k < TC_BK / 16; ++k) {
                        const uint64_t bk = (MODE == TC_MODE_GEMM) ? (uint64_t)(k * 128) : (uint64_t)(k * 2);
                        tcg::umma_bf16(tmem_base, da + (uint64_t)(k * 2), dbb + bk, idesc, (uint32_t)((i | k) != 0));
                    }
                }
                // frees the stage once these MMAs have read it (cluster-wide when the stage holds multicast data)
                if (csize == 1) tcg::umma_commit(&empty_bar[st]);
                else tcg::umma_commit_mcast(&empty_bar[st], cmask);
            }
            tcg::umma_commit(tmem_full_bar);        // accumulator complete
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
        const int row = quarter * 32 + lane;                // accumulator row == TMEM lane
        if (n_iters > 0) {
            tcg::mbar_wait(tmem_full_bar, 0);
            tcg::tc_fence_after();
        }
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // row -> output coordinates
        bool row_ok;
        long long row_off;
        if (MODE == TC_MODE_CONV) {
            int bwh = args.bw * args.bh;
            int in_ = row / bwh;
            int rem = row - in_ * bwh;
            int ih = rem / args.bw, iw = rem - ih * args.bw;
            int n = img0 + in_, p = p0 + ih, q = q0 + iw;
            row_ok = (row < args.bn * bwh) && n < args.NI && p < args.OP && q < args.OQ;
            row_off = args.o_off + (long long)n * args.o_sn + (long long)p * args.o_sh + (long long)q * args.o_sw;
        } else if (MODE == TC_MODE_GEMM) {
            int m = m_tile * TC_BM + row;
            row_ok = m < args.M;
            row_off = (long long)m * args.o_sn;
        } else {
            const int mt = (args.M + TC_BM - 1) / TC_BM;
            int m = (m_tile % mt) * TC_BM + row;
            row_ok = m < args.M;
            // per-tap output offset is folded into tap_bcol[] by the host (flipped filter position)
            row_off = args.o_off + (long long)m * args.o_sn + (long long)args.tap_bcol[wg_tap];
        }
        const int col0 = n_tile * BN;
        for (int cb = 0; cb < BN; cb += 16) {
            uint32_t r[16];
            if (n_iters > 0) {
                tcg::tmem_ld16(taddr + (uint32_t)cb, r);
                tcg::tmem_ld_wait();
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) r[j] = 0u;
            }
            if (!row_ok) continue;
            if (args.out_kind == TC_OUT_BF16 && args.o_sc == 1 && col0 + cb + 16 <= args.Nout) {
                // NHWC bf16: 16 consecutive channels of one pixel = 32 contiguous bytes
                __nv_bfloat162 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    v[j] = __floats2bfloat162_rn(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                uint4* dst = (uint4*)((__nv_bfloat16*)args.out + row_off + col0 + cb);
                dst[0] = *(uint4*)&v[0];
                dst[1] = *(uint4*)&v[4];
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    int c = col0 + cb + j;
                    if (c < args.Nout) {
                        long long o = row_off + (long long)c * args.o_sc;
                        float val = __uint_as_float(r[j]);
                        if (args.out_kind == TC_OUT_F32) ((float*)args.out)[o] = val;
                        else if (args.out_kind == TC_OUT_BF16) ((__nv_bfloat16*)args.out)[o] = __float2bfloat16_rn(val);
                        else atomicAdd((float*)args.out + o, val);
                    }
                }
            }
        }
    }

    __syncwarp();   // the single-lane producer / issuer loops diverged their warps; the cluster barrier is warp-aligned
    tcg::tc_fence_before();
    if (csize > 1) tcg::cluster_sync();   // nobody leaves while a peer may still write into / arrive on this CTA
    else __syncthreads();
    if (warp == 1) {
        tcg::tc_fence_after();
        if constexpr (pair) tcg::tmem_dealloc_2sm(tmem_base, tmem_cols);
        else tcg::tmem_dealloc(tmem_base, tmem_cols);
    }
}

}  // namespace db
