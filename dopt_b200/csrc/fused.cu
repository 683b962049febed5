// fused.cu -- interpreter kernel for fused pointwise regions (see fused.cuh).
//
// Memory behaviour: one 128-bit load per tensor input and one 128-bit store per region output per four elements; the
// temporaries of the little program live in shared memory (one float4 slot per virtual register per thread, conflict
// free), never in HBM.  HBM-bound: (inputs + outputs) * 4 B per element, e.g. 5 words for the SGD update of a filter
// tensor INCLUDING its weight-decay gradient, where the node-by-node graph moves 30.
#include "fused.cuh"
#include "pointwise.cuh"
#include <cstdio>
#include <cstdlib>

namespace db {

static constexpr int kFzThreads = 256;
static constexpr int kFzChunk = 4096;   // elements per CTA trip

__device__ __forceinline__ float fz_apply(int op, float a, float b) {
    switch (op) {
#define C(OP) case dbk::OP: return dbk::apply<dbk::OP, float>(a, b);
        C(OP_ADD) C(OP_SUB) C(OP_MUL) C(OP_DIV) C(OP_LT) C(OP_LTE) C(OP_GT) C(OP_GTE) C(OP_EQ) C(OP_NEQ) C(OP_MAX)
        C(OP_MIN) C(OP_POW) C(OP_NEG) C(OP_ABS) C(OP_SGN) C(OP_EXP) C(OP_LOG) C(OP_SQRT)
#undef C
    }
    return 0.f;
}
// the same on four lanes with ONE dispatch: the switch is the expensive part of an interpreted instruction
__device__ __forceinline__ float4 fz_apply4(int op, const float4& a, const float4& b) {
    float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
    switch (op) {
#define C(OP)                                         \
    case dbk::OP:                                     \
        d.x = dbk::apply<dbk::OP, float>(a.x, b.x);   \
        d.y = dbk::apply<dbk::OP, float>(a.y, b.y);   \
        d.z = dbk::apply<dbk::OP, float>(a.z, b.z);   \
        d.w = dbk::apply<dbk::OP, float>(a.w, b.w);   \
        break;
        C(OP_ADD) C(OP_SUB) C(OP_MUL) C(OP_DIV) C(OP_LT) C(OP_LTE) C(OP_GT) C(OP_GTE) C(OP_EQ) C(OP_NEQ) C(OP_MAX)
        C(OP_MIN) C(OP_POW) C(OP_NEG) C(OP_ABS) C(OP_SGN) C(OP_EXP) C(OP_LOG) C(OP_SQRT)
#undef C
    }
    return d;
}

__device__ __forceinline__ int fz_find_row(const FzRow* rows, int n_rows, int64_t chunk) {
    int lo = 0, hi = n_rows - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (rows[mid].chunk0 <= chunk) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

__global__ void __launch_bounds__(kFzThreads) fused_kernel(const FzRow* __restrict__ rows, int n_rows, int64_t n_chunks,
                                                           const __grid_constant__ FzProgram prog) {
    extern __shared__ float4 regs[];   // [register][thread]
    const int tid = threadIdx.x;
    const int nt = prog.n_tensors, ns = prog.n_scalars;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const FzRow& r = rows[fz_find_row(rows, n_rows, ch)];
        const int64_t base = (ch - r.chunk0) * kFzChunk;
        const int64_t end = base + kFzChunk < r.n ? base + kFzChunk : r.n;
        uintptr_t align = 0;
        for (int t = 0; t < nt; ++t) align |= (uintptr_t)r.in[t];
        for (int o = 0; o < prog.n_outputs; ++o) align |= (uintptr_t)r.out[o];
        const bool vec = (align & 15) == 0;
        for (int s = 0; s < ns; ++s) {
            float v = r.scalar[s][0];
            regs[(nt + s) * kFzThreads + tid] = make_float4(v, v, v, v);
        }
        // vector lanes: 4 elements per thread per trip; a scalar tail (or the whole row when unaligned) uses lane x only
        const int64_t vend = vec ? (base + ((end - base) & ~(int64_t)3)) : base;
        // the loads of trip i+1 are issued before the program of trip i runs (registers `pre`), so that global-memory latency
        // overlaps the interpreter's shared-memory round trips instead of adding to them
        float4 pre[FZ_MAX_TENSORS];
        int64_t i = base + (int64_t)tid * 4;
        if (i < vend) {
#pragma unroll
            for (int t = 0; t < FZ_MAX_TENSORS; ++t)
                if (t < nt) pre[t] = dbk::ld_stream((const float4*)(r.in[t] + i));
        }
        for (; i < vend; i += kFzThreads * 4) {
#pragma unroll
            for (int t = 0; t < FZ_MAX_TENSORS; ++t)
                if (t < nt) regs[t * kFzThreads + tid] = pre[t];
            const int64_t inext = i + kFzThreads * 4;
            if (inext < vend) {
#pragma unroll
                for (int t = 0; t < FZ_MAX_TENSORS; ++t)
                    if (t < nt) pre[t] = dbk::ld_stream((const float4*)(r.in[t] + inext));
            }
            for (int k = 0; k < prog.n_instr; ++k) {
                const FzInstr ins = prog.instr[k];
                const float4 a = regs[ins.a * kFzThreads + tid];
                const float4 b = regs[ins.b * kFzThreads + tid];
                regs[ins.dst * kFzThreads + tid] = fz_apply4(ins.op, a, b);
            }
            for (int o = 0; o < prog.n_outputs; ++o)
                dbk::st_stream((float4*)(r.out[o] + i), regs[prog.out_reg[o] * kFzThreads + tid]);
        }
        for (int64_t i = vend + tid; i < end; i += kFzThreads) {
            for (int t = 0; t < nt; ++t) regs[t * kFzThreads + tid].x = r.in[t][i];
            for (int k = 0; k < prog.n_instr; ++k) {
                const FzInstr ins = prog.instr[k];
                regs[ins.dst * kFzThreads + tid].x =
                    fz_apply(ins.op, regs[ins.a * kFzThreads + tid].x, regs[ins.b * kFzThreads + tid].x);
            }
            for (int o = 0; o < prog.n_outputs; ++o) r.out[o][i] = regs[prog.out_reg[o] * kFzThreads + tid].x;
        }
    }
}

// ---- pre-compiled programs ------------------------------------------------------------------------------------------------
// The interpreter pays a dispatch and three shared-memory accesses per instruction; for the one region that moves real
// bytes every step -- the parameter update, 5 words per parameter over all 36.5 M parameters of a WRN-28-10 -- that keeps
// the launch at 2.5 TB/s.  A program that matches an entry of this table runs as straight-line code instead: the same
// instruction list, unrolled at compile time over a register array whose indices are all constants, applying the same
// dbk::apply<> routines in the same order, so the result is bit-identical to the interpreter (and to the unfused graph).
// Adding an entry is adding data (the program as the plan dump prints it, DOPT_B200_PLAN_DUMP=1), not code.
template <int ID> struct FzStatic;
// SGD + momentum with the weight-decay gradient (online/source/dopt/online/sgd.d:57-64 + the d/dW of wd*sum(W*W),
// nnet/source/dopt/nnet/layers/conv.d:117): tensors m, W, dW; scalars mu, wd, lr; outputs m', W'
//   r0 = m*mu; r6 = wd*W; r7 = wd*W; r7 = r6+r7; r2 = r7+dW; r2 = lr*r2; r2 = r0+r2 (= m'); r1 = W-r2 (= W')
template <> struct FzStatic<0> {
    static constexpr FzProgram P = {8, 3, 3, 2,
                                    {{2, 0, 3, 0}, {2, 4, 1, 6}, {2, 4, 1, 7}, {0, 6, 7, 7}, {0, 7, 2, 2}, {2, 5, 2, 2},
                                     {0, 0, 2, 2}, {1, 1, 2, 1}},
                                    {2, 1}};
};

constexpr int fz_static_regs(const FzProgram& p) {
    int m = p.n_tensors + p.n_scalars - 1;
    for (int k = 0; k < p.n_instr; ++k) m = p.instr[k].dst > m ? p.instr[k].dst : m;
    return m + 1;
}

template <int OP> __device__ __forceinline__ float4 fz_op4(const float4& a, const float4& b) {
    return make_float4(dbk::apply<OP, float>(a.x, b.x), dbk::apply<OP, float>(a.y, b.y), dbk::apply<OP, float>(a.z, b.z),
                       dbk::apply<OP, float>(a.w, b.w));
}
template <int ID, int K, int NR> struct FzRun {
    static __device__ __forceinline__ void go(float4 (&r)[NR]) {
        if constexpr (K < FzStatic<ID>::P.n_instr) {
            constexpr int op = FzStatic<ID>::P.instr[K].op, a = FzStatic<ID>::P.instr[K].a, b = FzStatic<ID>::P.instr[K].b,
                          d = FzStatic<ID>::P.instr[K].dst;
            r[d] = fz_op4<op>(r[a], r[b]);
            FzRun<ID, K + 1, NR>::go(r);
        }
    }
};

template <int ID, int O, int NR, int NO> struct FzStore {
    static __device__ __forceinline__ void vec(float* const (&out)[NO], int64_t i, const float4 (&r)[NR]) {
        if constexpr (O < NO) {
            constexpr int reg = FzStatic<ID>::P.out_reg[O];
            dbk::st_stream((float4*)(out[O] + i), r[reg]);
            FzStore<ID, O + 1, NR, NO>::vec(out, i, r);
        }
    }
    static __device__ __forceinline__ void lane_x(float* const (&out)[NO], int64_t i, const float4 (&r)[NR]) {
        if constexpr (O < NO) {
            constexpr int reg = FzStatic<ID>::P.out_reg[O];
            out[O][i] = r[reg].x;
            FzStore<ID, O + 1, NR, NO>::lane_x(out, i, r);
        }
    }
};

template <int ID>
__global__ void __launch_bounds__(kFzThreads) fused_static_kernel(const FzRow* __restrict__ rows, int n_rows,
                                                                  int64_t n_chunks) {
    constexpr int NT = FzStatic<ID>::P.n_tensors, NS = FzStatic<ID>::P.n_scalars, NO = FzStatic<ID>::P.n_outputs;
    constexpr int NR = fz_static_regs(FzStatic<ID>::P);
    const int tid = threadIdx.x;
    for (int64_t ch = blockIdx.x; ch < n_chunks; ch += gridDim.x) {
        const FzRow& r = rows[fz_find_row(rows, n_rows, ch)];
        const int64_t base = (ch - r.chunk0) * kFzChunk;
        const int64_t end = base + kFzChunk < r.n ? base + kFzChunk : r.n;
        const float* in[NT];
        float* out[NO];
        float4 sc[NS > 0 ? NS : 1];
        uintptr_t align = 0;
#pragma unroll
        for (int t = 0; t < NT; ++t) { in[t] = r.in[t]; align |= (uintptr_t)in[t]; }
#pragma unroll
        for (int o = 0; o < NO; ++o) { out[o] = r.out[o]; align |= (uintptr_t)out[o]; }
#pragma unroll
        for (int q = 0; q < NS; ++q) { const float v = r.scalar[q][0]; sc[q] = make_float4(v, v, v, v); }
        const bool vec = (align & 15) == 0;
        const int64_t vend = vec ? (base + ((end - base) & ~(int64_t)3)) : base;
        // a chunk is four trips of the CTA: all loads of the chunk are issued before the first result is needed
        constexpr int TRIPS = kFzChunk / (kFzThreads * 4);
        float4 ld[TRIPS][NT];
#pragma unroll
        for (int u = 0; u < TRIPS; ++u) {
            const int64_t i = base + (int64_t)(u * kFzThreads + tid) * 4;
#pragma unroll
            for (int t = 0; t < NT; ++t)
                if (i < vend) ld[u][t] = dbk::ld_stream((const float4*)(in[t] + i));
        }
#pragma unroll
        for (int u = 0; u < TRIPS; ++u) {
            const int64_t i = base + (int64_t)(u * kFzThreads + tid) * 4;
            if (i < vend) {
                float4 R[NR];
#pragma unroll
                for (int t = 0; t < NT; ++t) R[t] = ld[u][t];
#pragma unroll
                for (int q = 0; q < NS; ++q) R[NT + q] = sc[q];
                FzRun<ID, 0, NR>::go(R);
                FzStore<ID, 0, NR, NO>::vec(out, i, R);
            }
        }
        for (int64_t i = vend + tid; i < end; i += kFzThreads) {   // scalar tail / unaligned rows: lane x only
            float4 R[NR];
#pragma unroll
            for (int t = 0; t < NT; ++t) R[t] = make_float4(in[t][i], 0.f, 0.f, 0.f);
#pragma unroll
            for (int q = 0; q < NS; ++q) R[NT + q] = sc[q];
            FzRun<ID, 0, NR>::go(R);
            FzStore<ID, 0, NR, NO>::lane_x(out, i, R);
        }
    }
}

static bool fz_same_program(const FzProgram& a, const FzProgram& b) {
    if (a.n_instr != b.n_instr || a.n_tensors != b.n_tensors || a.n_scalars != b.n_scalars || a.n_outputs != b.n_outputs)
        return false;
    for (int k = 0; k < a.n_instr; ++k)
        if (a.instr[k].op != b.instr[k].op || a.instr[k].a != b.instr[k].a || a.instr[k].b != b.instr[k].b ||
            a.instr[k].dst != b.instr[k].dst)
            return false;
    for (int o = 0; o < a.n_outputs; ++o)
        if (a.out_reg[o] != b.out_reg[o]) return false;
    return true;
}
// index into the table of pre-compiled programs, or -1 (DOPT_B200_NO_STATIC=1, read when the launch is bound: always -1,
// everything is interpreted)
static int fz_static_id(const FzProgram& p) {
    const char* e = getenv("DOPT_B200_NO_STATIC");
    if (e && atoi(e) != 0) return -1;
    if (fz_same_program(p, FzStatic<0>::P)) return 0;
    return -1;
}

void fused_launch(FzLaunch& L, cudaStream_t s) {
    if (L.rows.empty()) return;
    if (L.dirty) {
        // (re)build the chunk prefix and upload the row table; happens outside CUDA-graph capture (the plan runs every
        // launch eagerly once before it captures)
        int64_t chunk = 0;
        for (auto& r : L.rows) {
            r.chunk0 = chunk;
            chunk += ceil_div(std::max<int64_t>(r.n, 1), kFzChunk);
        }
        L.n_chunks = chunk;
        L.static_id = fz_static_id(L.prog);
        if (getenv("DOPT_B200_PLAN_DUMP")) {
            int64_t elems = 0;
            for (auto& r : L.rows) elems += r.n;
            fprintf(stderr, "PLAN fused-bind rows=%zu instr=%d tensors=%d outputs=%d elements=%lld static=%d\n", L.rows.size(),
                    L.prog.n_instr, L.prog.n_tensors, L.prog.n_outputs, (long long)elems, L.static_id);
        }
        if (!L.dev_rows) DB_CUDA(cudaMalloc(&L.dev_rows, L.rows.size() * sizeof(FzRow)));
        DB_CUDA(cudaMemcpy(L.dev_rows, L.rows.data(), L.rows.size() * sizeof(FzRow), cudaMemcpyHostToDevice));
        L.dirty = false;
    }
    if (L.static_id == 0) {
        int grid = (int)std::min<int64_t>(L.n_chunks, (int64_t)sm_count() * 8);
        fused_static_kernel<0><<<grid, kFzThreads, 0, s>>>(L.dev_rows, (int)L.rows.size(), L.n_chunks);
        DB_LAUNCH_CHECK();
        return;
    }
    int max_reg = 0;
    for (int k = 0; k < L.prog.n_instr; ++k) max_reg = std::max<int>(max_reg, L.prog.instr[k].dst);
    max_reg = std::max(max_reg, L.prog.n_tensors + L.prog.n_scalars - 1);
    size_t smem = (size_t)(max_reg + 1) * kFzThreads * sizeof(float4);
    static size_t configured = 48 * 1024;
    if (smem > configured) {
        DB_CUDA(cudaFuncSetAttribute(fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    int ctas_per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
    int grid = (int)std::min<int64_t>(L.n_chunks, (int64_t)sm_count() * ctas_per_sm);
    fused_kernel<<<grid, kFzThreads, smem, s>>>(L.dev_rows, (int)L.rows.size(), L.n_chunks, L.prog);
    DB_LAUNCH_CHECK();
}

void fused_free(FzLaunch& L) {
    if (L.dev_rows) cudaFree(L.dev_rows);
    L.dev_rows = nullptr;
}

}  // namespace db
