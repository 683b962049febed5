// tc.cuh -- host-side interface of the tcgen05 (5th-gen tensor core) kernels in tc_gemm.cu / conv_tc.cu.
#pragma once
#include "common.cuh"
#include "conv.cuh"

namespace db {

// ---- plain GEMM: C[M,N] = A[M,K] * B[K,N], fp32 row-major in/out, bf16 operands, fp32 accumulate ---------------------
struct TcGemm;
bool tc_gemm_supported(int64_t M, int64_t N, int64_t K);
TcGemm* tc_gemm_create(int64_t M, int64_t N, int64_t K);
void tc_gemm_run(TcGemm* g, const float* A, const float* B, float* C, cudaStream_t s);
void tc_gemm_destroy(TcGemm* g);

// ---- convolution (NCHW fp32 at the boundary, NHWC bf16 inside) -------------------------------------------------------
enum ConvKind { CONV_FWD = 0, CONV_DGRAD = 1, CONV_WGRAD = 2 };
struct ConvTc;
bool conv_tc_supported(const ConvGeom& g, int kind);
ConvTc* conv_tc_create(const ConvGeom& g, int kind);
// fwd: a = x, b = w, out = y;  dgrad: a = dy, b = w, out = dx;  wgrad: a = dy, b = x, out = dw  (all NCHW / KCRS fp32)
void conv_tc_run(ConvTc* c, const float* a, const float* b, float* out, cudaStream_t s);
void conv_tc_destroy(ConvTc* c);
void conv_tc_set_staged(ConvTc* c, int input, const void* nhwc_bf16);
size_t conv_tc_staged_bytes(const ConvTc* c, int input);
bool conv_tc_filter_pack(const ConvTc* c, int input, FilterPack* desc);
void conv_tc_set_packed_filter(ConvTc* c, const void* packed);
// fwd / dgrad: write the result as [N][H][W][Cp] bf16 into `nhwc_bf16` instead of NCHW fp32 into `out` (nullptr: back to fp32)
bool conv_tc_can_stage_output(const ConvTc* c);
void conv_tc_set_staged_output(ConvTc* c, void* nhwc_bf16);
// fwd with staged output: also accumulate sum / sum of squares per output channel into the statistics workspace of the flat
// batchNormTrain that reads the result (flat.cuh: flat_stats_sink)
void conv_tc_set_stats_workspace(ConvTc* c, void* bn_workspace);
// epilogue companion, see TcArgs::ep_src.  mode 2 (fwd): result = convolution + src; mode 3 (unit-stride dgrad): src = x and
// coef = forward coefficients of the batch norm whose backward pass reads the result; its statistics go to the workspace given
// with conv_tc_set_stats_workspace.  Both need the NHWC bf16 output.
bool conv_tc_can_companion(const ConvTc* c, int mode);
void conv_tc_set_companion(ConvTc* c, int mode, const void* src, const float* coef);
// wgrad: accumulate into a private scratch (allocated here) and skip the per-op finish kernel; the caller finishes `row` later
bool conv_tc_defer_finish(ConvTc* c, WgradFinish* row);
// ... into a scratch the caller owns and zeroes before every execution (replaces the private one)
void conv_tc_set_scratch(ConvTc* c, float* zeroed_by_caller);
// wgrad that reads only plan-staged operands and writes only its private scratch: it shares no workspace with any other op, so
// the plan may run it on a second stream beside the feature-gradient chain
bool conv_tc_side_stream_safe(const ConvTc* c);


// SM reservation gate of a data-parallel plan (tc_host.cu / tc_kernel.cuh: TcArgs::gate_*): while armed (need > 0), every full
// persistent grid launched through tc_launch decides on the device whether its last `sms x CTAs-per-SM` CTAs take part.
struct TcGate {
    unsigned* dev = nullptr;        // [0] all-reduces completed (ever) | [1] its value at step start | [2..5] snapshots
    cudaStream_t side = nullptr;    // launches on this stream form a chain of their own (the plan's filter-gradient stream)
    int sms = 0;                    // SMs the collective's CTAs occupy (= its channel count)
    int need = 0;                   // all-reduces enqueued so far in this step; 0 = gate not armed
    int idx[2] = {0, 0};            // gated launches so far on the main / side chain
};
void tc_gate_create(TcGate* g, int sms);
void tc_gate_destroy(TcGate* g);
void tc_gate_step_begin(TcGate* g, cudaStream_t s);      // first launch of the step (main stream)
void tc_gate_comm_done(TcGate* g, cudaStream_t comm);    // behind every all-reduce on the communication stream
void tc_set_gate(TcGate* g);                             // the gate tc_launch consults (nullptr: none)

}  // namespace db
