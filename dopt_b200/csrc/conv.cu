// conv.cu -- kernel objects for convolution / convolutionFeaturesGrad / convolutionFiltersGrad
// (cuda/source/dopt/cuda/nnet/cudnn7.d:55-249).  The reference benchmarks cuDNN algorithms at plan build
// (cudnnFind*Algorithm, cudnn7.d:125-141) and shares one static workspace; here the choice is static:
//   MATH_BF16 (default) and a shape the implicit-GEMM kernel tiles  -> tcgen05 path (conv_tc.cu)
//   otherwise                                                          -> fp32 direct kernels (conv_simt.cu)
#include "common.cuh"
#include "conv.cuh"
#include "tc.cuh"
#include "flat.cuh"

namespace db {
namespace {

static ConvGeom make_geom(const dopt_b200_tensor& x, const dopt_b200_tensor& w, const dopt_b200_tensor& y,
                          const dopt_b200_op& d) {
    DB_REQUIRE(x.rank == 4 && w.rank == 4 && y.rank == 4, "convolution: rank-4 tensors required");   // core/ops/nnet.d:43-66
    DB_REQUIRE(x.dtype == DOPT_B200_FLOAT32 && w.dtype == DOPT_B200_FLOAT32, "convolution: float32 only");
    ConvGeom g;
    g.N = (int)x.shape[0]; g.C = (int)x.shape[1]; g.H = (int)x.shape[2]; g.W = (int)x.shape[3];
    g.K = (int)w.shape[0]; g.R = (int)w.shape[2]; g.S = (int)w.shape[3];
    g.P = (int)y.shape[2]; g.Q = (int)y.shape[3];
    g.ph = (int)d.padding[0]; g.pw = (int)d.padding[1];
    g.u = (int)d.stride[0]; g.v = (int)d.stride[1];
    DB_REQUIRE(g.u >= 1 && g.v >= 1, "convolution: stride must be >= 1");
    DB_REQUIRE(w.shape[1] == g.C, "convolution: filter channels must match feature channels");
    DB_REQUIRE(y.shape[0] == g.N && y.shape[1] == g.K, "convolution: output batch/channels mismatch");
    // judgeConvolution, core/ops/nnet.d:68-87
    DB_REQUIRE(g.P == (g.H + 2 * g.ph - g.R) / g.u + 1 && g.Q == (g.W + 2 * g.pw - g.S) / g.v + 1,
               "convolution: output spatial size does not follow (in + 2*pad - filter)/stride + 1");
    return g;
}

struct ConvKernel : Kernel {
    ConvGeom g;
    int kind;
    ConvTc* tc = nullptr;
    ConvKernel(const dopt_b200_op& d, int kind_) : kind(kind_) {
        DB_REQUIRE(d.n_inputs == 2, "convolution ops take two operands");
        if (kind == CONV_FWD) g = make_geom(d.inputs[0], d.inputs[1], d.output, d);            // [x, w] -> y
        else if (kind == CONV_DGRAD) g = make_geom(d.output, d.inputs[1], d.inputs[0], d);     // [dy, w] -> dx (cudnn7.d:167-171)
        else g = make_geom(d.inputs[1], d.output, d.inputs[0], d);                             // [dy, x] -> dw (cudnn7.d:212-216)
        if (resolve_math(d.math) == DOPT_B200_MATH_BF16 && conv_tc_supported(g, kind)) tc = conv_tc_create(g, kind);
    }
    ~ConvKernel() { conv_tc_destroy(tc); }
    size_t staged_bytes(int input) const override { return conv_tc_staged_bytes(tc, input); }
    void set_staged_input(int input, const void* p) override { conv_tc_set_staged(tc, input, p); }
    bool filter_pack(int input, FilterPack* d) const override { return tc && conv_tc_filter_pack(tc, input, d); }
    void set_packed_filter(const void* p) override { conv_tc_set_packed_filter(tc, p); }
    bool can_stage_output() const override { return conv_tc_can_stage_output(tc); }
    void set_staged_output(void* p) override { conv_tc_set_staged_output(tc, p); }
    bool deferred_finish(WgradFinish* row) override { return tc && kind == CONV_WGRAD && conv_tc_defer_finish(tc, row); }
    void set_finish_scratch(float* p) override { conv_tc_set_scratch(tc, p); }
    bool side_stream_safe() const override { return tc && conv_tc_side_stream_safe(tc); }
    int can_produce_stats() const override { return (tc && kind == CONV_FWD) ? 2 : 0; }
    void set_stats_workspace(void* w, int channels) override {
        if (tc && kind == CONV_FWD && channels == g.K) conv_tc_set_stats_workspace(tc, w);
        if (tc && kind == CONV_DGRAD && channels == g.C) conv_tc_set_stats_workspace(tc, w);
    }
    bool can_companion(int mode) const override { return tc && conv_tc_can_companion(tc, mode); }
    void set_companion(int mode, const void* src, const float* coef) override { conv_tc_set_companion(tc, mode, src, coef); }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override {
        DB_REQUIRE(n_in == 2, "convolution ops take two inputs");
        const float* a = (const float*)in[0];
        const float* b = (const float*)in[1];
        if (tc) {
            conv_tc_run(tc, a, b, (float*)out, s);
            return;
        }
        if (kind == CONV_FWD) conv_fwd_simt_launch(a, b, (float*)out, g, s);
        else if (kind == CONV_DGRAD) conv_dgrad_simt_launch(a, b, (float*)out, g, s);
        else conv_wgrad_simt_launch(a, b, (float*)out, g, s);
    }
};
Kernel* make_fwd(const dopt_b200_op& d) { return new ConvKernel(d, CONV_FWD); }
Kernel* make_dgrad(const dopt_b200_op& d) { return new ConvKernel(d, CONV_DGRAD); }
Kernel* make_wgrad(const dopt_b200_op& d) { return new ConvKernel(d, CONV_WGRAD); }
}  // namespace

void register_conv() {
    register_kernel("convolution", make_fwd);
    register_kernel("convolutionFeaturesGrad", make_dgrad);
    register_kernel("convolutionFiltersGrad", make_wgrad);
}

}  // namespace db
