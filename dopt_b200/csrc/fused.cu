// fused.cu -- interpreter kernel for fused pointwise regions (see fused.cuh).
//
// Memory behaviour: one 128-bit load per tensor input and one 128-bit store per region output per four elements; the
// temporaries of the little program live in shared memory (one float4 slot per virtual register per thread, conflict
// free), never in HBM.  HBM-bound: (inputs + outputs) * 4 B per element, e.g. 5 words for the SGD update of a filter
// tensor INCLUDING its weight-decay gradient, where the node-by-node graph moves 30.
#include "fused.cuh"
#include "pointwise.cuh"

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
        if (!L.dev_rows) DB_CUDA(cudaMalloc(&L.dev_rows, L.rows.size() * sizeof(FzRow)));
        DB_CUDA(cudaMemcpy(L.dev_rows, L.rows.data(), L.rows.size() * sizeof(FzRow), cudaMemcpyHostToDevice));
        L.dirty = false;
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
