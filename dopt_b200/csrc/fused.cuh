// fused.cuh -- fused pointwise regions.
//
// dopt expresses everything that is not a layer as chains of single pointwise ops over whole tensors: the optimiser
// updates (online/source/dopt/online/sgd.d:57-64, adam.d:52-66), the weight-decay gradient (two `s*W` products and two
// adds per filter tensor, from mulGrad / sumGrad in core/source/dopt/core/grads/math.d:47-76), residual sums, the
// cross-entropy chain.  The reference launches (and synchronises) once per node.  The plan compiler instead collects
// connected pointwise nodes of equal volume into a REGION and runs the whole region in one pass over memory: every
// element is loaded once, pushed through the region's little program in registers/L1, and each value that is needed
// outside the region is stored once.  Regions with identical programs (the 28 filter updates of a WRN, its 50 BN affine
// updates ...) are batched into a single multi-tensor launch.
//
// The program is interpreted (no runtime code generation).  Each instruction applies the same `dbk::apply<>` routine the
// stand-alone pointwise kernels use, in the graph's own order and without FMA contraction, so a fused region is
// bit-identical to running its nodes one by one.
#pragma once
#include "common.cuh"

namespace db {

static constexpr int FZ_MAX_INSTR = 32;
static constexpr int FZ_MAX_REGS = 24;
static constexpr int FZ_MAX_TENSORS = 8;    // tensor inputs
static constexpr int FZ_MAX_SCALARS = 8;    // rank-0 device operands
static constexpr int FZ_MAX_OUTPUTS = 6;

// operand encoding: 0..FZ_MAX_REGS-1 = register; 64+i = tensor input i; 128+i = scalar input i
struct FzInstr {
    uint8_t op, a, b, dst;
};
struct FzProgram {
    int n_instr, n_tensors, n_scalars, n_outputs;
    FzInstr instr[FZ_MAX_INSTR];
    uint8_t out_reg[FZ_MAX_OUTPUTS];
};
// one tensor group of a (possibly multi-tensor) launch
struct FzRow {
    const float* in[FZ_MAX_TENSORS];
    const float* scalar[FZ_MAX_SCALARS];
    float* out[FZ_MAX_OUTPUTS];
    int64_t n;
    int64_t chunk0;
};

struct FzLaunch {
    FzProgram prog;
    std::vector<FzRow> rows;      // host copy
    FzRow* dev_rows = nullptr;    // device copy (uploaded lazily, re-uploaded when pointers change)
    int64_t n_chunks = 0;
    int static_id = -1;           // index into fused.cu's table of pre-compiled programs, -1 = interpreted
    bool dirty = true;
};

void fused_launch(FzLaunch& L, cudaStream_t s);
void fused_free(FzLaunch& L);

}  // namespace db
