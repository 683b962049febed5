// capi.cpp -- a small C API over the C++ host mirror so that the Python test / bench harness can build dopt graphs,
// differentiate them, export them (for the CPU oracle) and run plans / updaters.  Handles are integer ids into per-process
// tables.  This is harness plumbing, not part of the drop-in boundary (that is include/dopt_b200.h).
#include <cuda_runtime.h>

#include <cmath>
#include <cstring>
#include <sstream>

#include "../../include/dopt_b200.h"
#include "dopt/core.hpp"
#include "dopt/cuda.hpp"
#include "dopt/nnet.hpp"
#include "dopt/online.hpp"

using namespace dopt;

namespace {
std::vector<Operation> g_ops;
std::vector<nnet::LayerPtr> g_layers;
std::vector<std::shared_ptr<nnet::DAGNetwork>> g_nets;
std::vector<PlanPtr> g_plans;
struct UpdaterRec {
    online::Updater fn;
    online::LastUpdate info;
};
std::vector<UpdaterRec> g_updaters;
std::string g_error, g_text;

int addOp(Operation o) {
    g_ops.push_back(std::move(o));
    return (int)g_ops.size() - 1;
}
Operation op(int id) {
    enforce(id >= 0 && id < (int)g_ops.size(), "bad op handle");
    return g_ops[id];
}
std::vector<Operation> opList(const int* ids, int n) {
    std::vector<Operation> v;
    for (int i = 0; i < n; ++i) v.push_back(op(ids[i]));
    return v;
}
std::vector<size_t> sizes(const int64_t* v, int n) { return std::vector<size_t>(v, v + n); }
int findOp(const Operation& o) {
    for (size_t i = 0; i < g_ops.size(); ++i)
        if (g_ops[i] == o) return (int)i;
    return addOp(o);
}
}  // namespace

#define DH_TRY try {
#define DH_CATCH(fail)                   \
    }                                    \
    catch (const std::exception& e) {    \
        g_error = e.what();              \
        return fail;                     \
    }

extern "C" {

const char* dh_last_error() { return g_error.c_str(); }

// returns 1 when the CUDA backend came up (device present), 0 otherwise (graph construction still works)
int dh_init() {
    DH_TRY
    dopt::initialize();
    return cuda::initialize() ? 1 : 0;
    DH_CATCH(-1)
}
const char* dh_init_error() { return cuda::lastInitError().c_str(); }
void dh_reset() {
    g_updaters.clear();
    g_plans.clear();
    g_nets.clear();
    g_layers.clear();
    g_ops.clear();
}
void dh_set_plan_flags(int f) { cuda::setPlanFlags(f); }
int dh_plan_flags() { return cuda::planFlags(); }
void dh_set_math(int m) { cuda::setMath(m); }
void dh_set_stream(void* s) { cuda::setStream(s); }
void dh_seed(uint64_t s) { nnet::seedInitializers(s); }
int dh_init_data_parallel(int rank, int world, const void* id128) {
    DH_TRY
    cuda::initDataParallel(rank, world, id128);
    return 0;
    DH_CATCH(-1)
}
void dh_set_data_parallel_world(int world) { setDataParallelWorld(world); }   // graph construction only (CPU tests)

// ---- graph construction ------------------------------------------------------------------------------------------------
int dh_variable(int dtype, const int64_t* shape, int rank, const void* data) {
    DH_TRY
    return addOp(variable(TensorType(dtype == 0 ? DataType::float32 : DataType::int32, sizes(shape, rank)), data));
    DH_CATCH(-1)
}
int dh_constant(int dtype, const int64_t* shape, int rank, const void* data) {
    DH_TRY
    return addOp(constant(TensorType(dtype == 0 ? DataType::float32 : DataType::int32, sizes(shape, rank)), data));
    DH_CATCH(-1)
}
// generic createOperation: attribute i has name names[i], kind kinds[i] (1 = size_t[], 2 = size_t, 3 = double) and
// takes lens[i] values from ivals (kinds 1, 2) or one value from dvals (kind 3)
int dh_create(const char* type, const int* deps, int ndeps, int nattrs, const char* const* names, const int* kinds,
              const int64_t* ivals, const int* lens, const double* dvals) {
    DH_TRY
    Attributes a;
    int ip = 0, dp = 0;
    for (int i = 0; i < nattrs; ++i) {
        if (kinds[i] == 1) {
            a[names[i]] = Variant(sizes(ivals + ip, lens[i]));
            ip += lens[i];
        } else if (kinds[i] == 2) {
            a[names[i]] = Variant((size_t)ivals[ip]);
            ip += 1;
        } else {
            a[names[i]] = Variant(dvals[dp++]);
        }
    }
    return addOp(createOperation(type, opList(deps, ndeps), a));
    DH_CATCH(-1)
}
// sugar that goes through the same lowering as the D operators / helper functions
int dh_binary(int opchar, int a, int b) {
    DH_TRY
    switch (opchar) {
        case '+': return addOp(op(a) + op(b));
        case '-': return addOp(op(a) - op(b));
        case '*': return addOp(op(a) * op(b));
        case '/': return addOp(op(a) / op(b));
    }
    throw Exception("Unknown binary operation");
    DH_CATCH(-1)
}
int dh_binary_scalar(int opchar, int a, float s, int scalar_left) {
    DH_TRY
    Operation x = op(a);
    switch (opchar) {
        case '+': return addOp(scalar_left ? s + x : x + s);
        case '-': return addOp(scalar_left ? s - x : x - s);
        case '*': return addOp(scalar_left ? s * x : x * s);
        case '/': return addOp(scalar_left ? s / x : x / s);
    }
    throw Exception("Unknown binary operation");
    DH_CATCH(-1)
}
int dh_repeat_n(int a, int64_t n) {
    DH_TRY
    return addOp(repeat(op(a), (size_t)n));
    DH_CATCH(-1)
}
int dh_sum(int a, const int64_t* axes, int n) {
    DH_TRY
    return addOp(sum(op(a), sizes(axes, n)));
    DH_CATCH(-1)
}
int dh_max_element(int a, const int64_t* axes, int n) {
    DH_TRY
    return addOp(maxElement(op(a), sizes(axes, n)));
    DH_CATCH(-1)
}
// batchNormTrain returns three ops: out3 receives their handles
int dh_batch_norm_train(int x, int scale, int bias, int mean, int var, double momentum, int* out3) {
    DH_TRY
    auto r = batchNormTrain(op(x), op(scale), op(bias), op(mean), op(var), momentum);
    for (int i = 0; i < 3; ++i) out3[i] = addOp(r[i]);
    return 0;
    DH_CATCH(-1)
}
int dh_maxpool_grad(int pg, int poolop) {
    DH_TRY
    return addOp(maxpoolGrad(op(pg), op(poolop)));
    DH_CATCH(-1)
}
int dh_convolution_transpose(int f, int w, const int64_t* pad, const int64_t* stride) {
    DH_TRY
    return addOp(convolutionTranspose(op(f), op(w), sizes(pad, 2), sizes(stride, 2)));
    DH_CATCH(-1)
}
int dh_grad(int objective, const int* wrt, int n, int* out) {
    DH_TRY
    auto g = grad(op(objective), opList(wrt, n));
    for (int i = 0; i < n; ++i) out[i] = addOp(g[i]);
    return 0;
    DH_CATCH(-1)
}
int dh_cross_entropy(int hyp, int truth) {
    DH_TRY
    return addOp(nnet::crossEntropy(op(hyp), op(truth)));
    DH_CATCH(-1)
}
int dh_squared_error(int hyp, int truth) {
    DH_TRY
    return addOp(nnet::squaredError(op(hyp), op(truth)));
    DH_CATCH(-1)
}

// ---- introspection -------------------------------------------------------------------------------------------------------
int dh_op_rank(int id) {
    DH_TRY
    return (int)op(id)->rank();
    DH_CATCH(-1)
}
int dh_op_shape(int id, int64_t* out) {
    DH_TRY
    auto& s = op(id)->shape();
    for (size_t i = 0; i < s.size(); ++i) out[i] = (int64_t)s[i];
    return (int)s.size();
    DH_CATCH(-1)
}
int dh_op_dtype(int id) {
    DH_TRY
    return op(id)->elementType() == DataType::float32 ? 0 : 1;
    DH_CATCH(-1)
}
int64_t dh_op_serial(int id) {
    DH_TRY
    return (int64_t)op(id)->id();
    DH_CATCH(-1)
}
int dh_get_value(int id, void* out, size_t bytes) {
    DH_TRY
    auto v = op(id)->value();
    enforce(v != nullptr, "operation has no value buffer");
    v->get(out, bytes);
    return 0;
    DH_CATCH(-1)
}
int dh_set_value(int id, const void* data, size_t bytes) {
    DH_TRY
    auto v = op(id)->value();
    enforce(v != nullptr, "operation has no value buffer");
    v->set(data, bytes);
    return 0;
    DH_CATCH(-1)
}
void* dh_value_device_ptr(int id) {
    try {
        auto cu = std::dynamic_pointer_cast<cuda::CUDABuffer>(op(id)->value());
        return cu ? cu->ptr() : nullptr;
    } catch (...) {
        return nullptr;
    }
}

// Toposorted export of the graph reaching `outputs`, one node per line:
//   serial|opType|dtype|shape,..|dep serials,..|attr=name:kind:v,v;...|handle
// `handle` is the op handle (registered on demand) so that the caller can fetch variable / constant values.
const char* dh_export(const int* outputs, int n) {
    try {
        std::ostringstream ss;
        for (auto& o : topologicalSort(opList(outputs, n))) {
            ss << o->id() << "|" << o->opType() << "|" << (o->elementType() == DataType::float32 ? 0 : 1) << "|";
            for (size_t i = 0; i < o->shape().size(); ++i) ss << (i ? "," : "") << o->shape()[i];
            ss << "|";
            for (size_t i = 0; i < o->deps().size(); ++i) ss << (i ? "," : "") << o->deps()[i]->id();
            ss << "|";
            bool first = true;
            for (auto& kv : o->attributes()) {
                if (kv.second.kind == Variant::Type || kv.second.kind == Variant::Empty) continue;
                ss << (first ? "" : ";") << kv.first << ":";
                first = false;
                if (kv.second.kind == Variant::Sizes) {
                    ss << "1:";
                    for (size_t i = 0; i < kv.second.sizes.size(); ++i) ss << (i ? "," : "") << kv.second.sizes[i];
                } else if (kv.second.kind == Variant::Size) {
                    ss << "2:" << kv.second.size;
                } else {
                    ss.precision(17);
                    ss << "3:" << kv.second.real;
                }
            }
            ss << "|" << findOp(o) << "\n";
        }
        g_text = ss.str();
        return g_text.c_str();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}

// ---- layers / networks ---------------------------------------------------------------------------------------------------
static int addLayer(nnet::LayerPtr l) {
    g_layers.push_back(std::move(l));
    return (int)g_layers.size() - 1;
}
static nnet::LayerPtr layer(int id) {
    enforce(id >= 0 && id < (int)g_layers.size(), "bad layer handle");
    return g_layers[id];
}
int dh_data_source(int var) {
    DH_TRY
    return addLayer(nnet::dataSource(op(var)));
    DH_CATCH(-1)
}
int dh_conv2d(int in, int64_t channels, const int64_t* fdims, const int64_t* pad, const int64_t* stride, float wd,
              int use_bias) {
    DH_TRY
    nnet::Conv2DOptions o;
    o.padding = sizes(pad, 2);
    o.stride = sizes(stride, 2);
    o.weightDecay = wd;
    o.useBias = use_bias != 0;
    return addLayer(nnet::conv2D(layer(in), (size_t)channels, sizes(fdims, 2), o));
    DH_CATCH(-1)
}
int dh_dense(int in, int64_t outputs, float wd, int use_bias) {
    DH_TRY
    nnet::DenseOptions o;
    o.weightDecay = wd;
    o.useBias = use_bias != 0;
    return addLayer(nnet::dense(layer(in), (size_t)outputs, o));
    DH_CATCH(-1)
}
int dh_batch_norm(int in, float momentum) {
    DH_TRY
    nnet::BatchNormOptions o;
    o.momentum = momentum;
    return addLayer(nnet::batchNorm(layer(in), o));
    DH_CATCH(-1)
}
// the same three layers with the regulariser options of conv.d / dense.d / batchnorm.d (infinity / 0 = off)
int dh_conv2d_reg(int in, int64_t channels, const int64_t* fdims, const int64_t* pad, const int64_t* stride, float wd,
                  int use_bias, float maxgain, float spectral_decay) {
    DH_TRY
    nnet::Conv2DOptions o;
    o.padding = sizes(pad, 2);
    o.stride = sizes(stride, 2);
    o.weightDecay = wd;
    o.useBias = use_bias != 0;
    o.maxgain = maxgain;
    o.spectralDecay = spectral_decay;
    return addLayer(nnet::conv2D(layer(in), (size_t)channels, sizes(fdims, 2), o));
    DH_CATCH(-1)
}
int dh_dense_reg(int in, int64_t outputs, float wd, int use_bias, float maxgain, float spectral_decay) {
    DH_TRY
    nnet::DenseOptions o;
    o.weightDecay = wd;
    o.useBias = use_bias != 0;
    o.maxgain = maxgain;
    o.spectralDecay = spectral_decay;
    return addLayer(nnet::dense(layer(in), (size_t)outputs, o));
    DH_CATCH(-1)
}
int dh_batch_norm_reg(int in, float momentum, float maxgain, float lipschitz) {
    DH_TRY
    nnet::BatchNormOptions o;
    o.momentum = momentum;
    o.maxgain = maxgain;
    o.lipschitz = lipschitz;
    return addLayer(nnet::batchNorm(layer(in), o));
    DH_CATCH(-1)
}
int dh_relu(int in) {
    DH_TRY
    return addLayer(nnet::relu(layer(in)));
    DH_CATCH(-1)
}
int dh_max_pool(int in, const int64_t* dims) {
    DH_TRY
    return addLayer(nnet::maxPool(layer(in), sizes(dims, 2)));
    DH_CATCH(-1)
}
int dh_dropout(int in, float drop_prob) {
    DH_TRY
    return addLayer(nnet::dropout(layer(in), drop_prob));
    DH_CATCH(-1)
}
int dh_softmax(int in) {
    DH_TRY
    return addLayer(nnet::softmax(layer(in)));
    DH_CATCH(-1)
}
int dh_wide_resnet(int features, int64_t depth, int64_t width, const int64_t* stride3, float wd) {
    DH_TRY
    nnet::WRNOptions o;
    o.weightDecay = wd;
    for (int i = 0; i < 3; ++i) o.stride[i] = (size_t)stride3[i];
    return addLayer(nnet::wideResNet(op(features), (size_t)depth, (size_t)width, o));
    DH_CATCH(-1)
}
// WRNOptions with the regulariser fields of wrn.d:11-54 (NaN / infinity / 0 = off)
int dh_wide_resnet_reg(int features, int64_t depth, int64_t width, const int64_t* stride3, float wd, int dropout,
                       float maxgain_norm, float lipschitz_norm, float max_norm, float spectral_decay) {
    DH_TRY
    nnet::WRNOptions o;
    o.weightDecay = wd;
    for (int i = 0; i < 3; ++i) o.stride[i] = (size_t)stride3[i];
    o.dropout = dropout != 0;
    o.maxgainNorm = maxgain_norm;
    o.lipschitzNorm = lipschitz_norm;
    o.maxNorm = max_norm;
    o.spectralDecay = spectral_decay;
    return addLayer(nnet::wideResNet(op(features), (size_t)depth, (size_t)width, o));
    DH_CATCH(-1)
}
// VGGOptions with the regulariser fields of vgg.d:12-49; layers = 16 or 19
int dh_vgg_reg(int features, int layers, const int64_t* dense_sizes, int n, int batchnorm, int dropout, float maxgain_norm,
               float lipschitz_norm, float max_norm, float spectral_decay) {
    DH_TRY
    nnet::VGGOptions o;
    o.batchnorm = batchnorm != 0;
    o.dropout = dropout != 0;
    o.maxgainNorm = maxgain_norm;
    o.lipschitzNorm = lipschitz_norm;
    o.maxNorm = max_norm;
    o.spectralDecay = spectral_decay;
    enforce(layers == 16 || layers == 19, "vgg: 16 or 19 layers");
    return addLayer(layers == 16 ? nnet::vgg16(op(features), sizes(dense_sizes, n), o)
                                 : nnet::vgg19(op(features), sizes(dense_sizes, n), o));
    DH_CATCH(-1)
}
int dh_vgg19(int features, const int64_t* dense_sizes, int n, int batchnorm) {
    DH_TRY
    nnet::VGGOptions o;
    o.batchnorm = batchnorm != 0;
    return addLayer(nnet::vgg19(op(features), sizes(dense_sizes, n), o));
    DH_CATCH(-1)
}
int dh_layer_output(int l, int train) {
    DH_TRY
    return findOp(train ? layer(l)->trainOutput() : layer(l)->output());
    DH_CATCH(-1)
}
int dh_network(const int* inputs, int nin, const int* out_layers, int nout) {
    DH_TRY
    std::vector<nnet::LayerPtr> ls;
    for (int i = 0; i < nout; ++i) ls.push_back(layer(out_layers[i]));
    g_nets.push_back(std::make_shared<nnet::DAGNetwork>(opList(inputs, nin), ls));
    return (int)g_nets.size() - 1;
    DH_CATCH(-1)
}
int dh_network_param_loss(int net) {
    DH_TRY
    return findOp(g_nets.at(net)->paramLoss());
    DH_CATCH(-1)
}
int dh_network_params(int net, int* out, int cap) {
    DH_TRY
    auto& p = g_nets.at(net)->params();
    for (size_t i = 0; i < p.size() && (int)i < cap; ++i) out[i] = findOp(p[i]);
    return (int)p.size();
    DH_CATCH(-1)
}
// nnet/lipschitz.d: norms (p_code 1, 2, or 0 for infinity) and the max-norm projection
int dh_matrix_norm(int param, int p_code) {
    DH_TRY
    return addOp(nnet::matrixNorm(op(param), p_code == 0 ? INFINITY : (float)p_code));
    DH_CATCH(-1)
}
int dh_conv_params_norm(int param, const int64_t* in_shape, const int64_t* stride, const int64_t* padding, int p_code) {
    DH_TRY
    return addOp(nnet::convParamsNorm(op(param), sizes(in_shape, 2), sizes(stride, 2), sizes(padding, 2),
                                      p_code == 0 ? INFINITY : (float)p_code));
    DH_CATCH(-1)
}
int dh_max_norm(int param, int norm, int maxval) {
    DH_TRY
    return addOp(nnet::maxNorm(op(param), op(norm), op(maxval)));
    DH_CATCH(-1)
}
int dh_network_save(int net, const char* file) {
    DH_TRY
    g_nets.at(net)->save(file);
    return 0;
    DH_CATCH(-1)
}
int dh_network_load(int net, const char* file) {
    DH_TRY
    g_nets.at(net)->load(file);
    return 0;
    DH_CATCH(-1)
}

// ---- plans -----------------------------------------------------------------------------------------------------------------
// kind 0: defaultCompiler (B200Plan with the current flags); 1: the reference-style node-by-node CUDAPlan
int dh_compile(const int* outputs, int n, int kind) {
    DH_TRY
    PlanPtr p;
    if (kind == 1) p = std::make_shared<cuda::CUDAPlan>(opList(outputs, n));
    else p = compile(opList(outputs, n));
    g_plans.push_back(p);
    return (int)g_plans.size() - 1;
    DH_CATCH(-1)
}
static void runPlan(Plan& plan, const int* arg_ops, const void* const* arg_ptrs, const size_t* arg_bytes, int nargs,
                    void* const* out_ptrs, std::vector<Buffer>* rets_inout) {
    std::map<Operation, Buffer> args;
    for (int i = 0; i < nargs; ++i) args[op(arg_ops[i])] = buffer(arg_ptrs[i], arg_bytes[i]);   // host CPUBuffer-like args
    std::vector<Buffer> rets;
    if (rets_inout) {
        plan.execute(args, *rets_inout);
        rets = *rets_inout;
    } else {
        rets = plan.execute(args);
    }
    for (size_t i = 0; i < plan.outputs().size(); ++i)
        if (out_ptrs && out_ptrs[i]) rets[i]->get(out_ptrs[i], rets[i]->numBytes());
}
int dh_plan_execute(int plan, const int* arg_ops, const void* const* arg_ptrs, const size_t* arg_bytes, int nargs,
                    void* const* out_ptrs) {
    DH_TRY
    runPlan(*g_plans.at(plan), arg_ops, arg_ptrs, arg_bytes, nargs, out_ptrs, nullptr);
    return 0;
    DH_CATCH(-1)
}
int dh_plan_stats(int plan, int64_t* launches, int64_t* bytes, int64_t* nodes) {
    DH_TRY
    auto bp = std::dynamic_pointer_cast<cuda::B200Plan>(g_plans.at(plan));
    enforce(bp != nullptr, "not a B200Plan");
    bp->stats(launches, bytes, nodes);
    return 0;
    DH_CATCH(-1)
}

// ---- updaters ----------------------------------------------------------------------------------------------------------------
// kind: 0 sgd, 1 sgd+nesterov, 2 adam, 3 amsgrad.  hyper[] are op handles (or -1 for the reference defaults):
//   sgd: {learningRate, momentumRate}; adam / amsgrad: {alpha, beta1, beta2, eps}.  net < 0: no projections, wrt given.
int dh_updater(int kind, const int* outputs, int nout, int net, const int* wrt, int nwrt, const int* hyper) {
    DH_TRY
    std::vector<Operation> w;
    std::map<Operation, nnet::Projection> projs;
    if (net >= 0) {
        w = g_nets.at(net)->params();
        projs = g_nets.at(net)->paramProj();
    } else {
        w = opList(wrt, nwrt);
    }
    auto h = [&](int i) -> Operation { return hyper && hyper[i] >= 0 ? op(hyper[i]) : nullptr; };
    online::Updater fn;
    if (kind == 0 || kind == 1) fn = online::sgd(opList(outputs, nout), w, projs, h(0), h(1), kind == 1);
    else if (kind == 2) fn = online::adam(opList(outputs, nout), w, projs, h(0), h(1), h(2), h(3));
    else fn = online::amsgrad(opList(outputs, nout), w, projs, h(0), h(1), h(2), h(3));
    g_updaters.push_back(UpdaterRec{fn, online::lastUpdate()});
    return (int)g_updaters.size() - 1;
    DH_CATCH(-1)
}
// host-buffer step, exactly the call an example makes: updater([features: buffer(fs), labels: buffer(ls)]) then .get
int dh_updater_step(int u, const int* arg_ops, const void* const* arg_ptrs, const size_t* arg_bytes, int nargs,
                    void* const* out_ptrs) {
    DH_TRY
    auto& rec = g_updaters.at(u);
    std::map<Operation, Buffer> args;
    // the caller's (pinned) host memory is wrapped, not copied: the H2D copy reads it directly
    for (int i = 0; i < nargs; ++i)
        args[op(arg_ops[i])] = std::make_shared<HostBuffer>(const_cast<void*>(arg_ptrs[i]), arg_bytes[i]);
    auto rets = rec.fn(args);
    for (size_t i = 0; i < rets.size(); ++i)
        if (out_ptrs && out_ptrs[i]) rets[i]->get(out_ptrs[i], rets[i]->numBytes());
    return 0;
    DH_CATCH(-1)
}
// device-resident step for the kernel-only benchmark leg: arguments are DEVICE pointers, nothing is read back
int dh_updater_step_device(int u, const int* arg_ops, const void* const* dev_ptrs, int nargs) {
    DH_TRY
    auto& rec = g_updaters.at(u);
    auto bp = std::dynamic_pointer_cast<cuda::B200Plan>(rec.info.plan);
    enforce(bp != nullptr, "updater is not backed by a B200Plan");
    std::vector<Operation> ops = opList(arg_ops, nargs);
    std::vector<const void*> ptrs(dev_ptrs, dev_ptrs + nargs);
    std::vector<int> onHost(nargs, 0);
    std::vector<void*> rets;
    for (auto& b : rec.info.newbufs) {
        auto cu = std::dynamic_pointer_cast<cuda::CUDABuffer>(b);
        enforce(cu != nullptr, "updater state is not on the device");
        rets.push_back(cu->ptr());
    }
    bp->executeRaw(ops, ptrs, onHost, rets);
    return 0;
    DH_CATCH(-1)
}
// the plan an updater compiled: outputs ~ newvals ~ state, and for each the op handle it is written back to (-1 = none)
int dh_updater_plan_outputs(int u, int* plan_outputs, int* destinations, int cap) {
    DH_TRY
    auto& info = g_updaters.at(u).info;
    for (size_t i = 0; i < info.planOutputs.size() && (int)i < cap; ++i) {
        plan_outputs[i] = findOp(info.planOutputs[i]);
        destinations[i] = info.destinations[i] ? findOp(info.destinations[i]) : -1;
    }
    return (int)info.planOutputs.size();
    DH_CATCH(-1)
}
// training checkpoint (parameters + optimiser state), see online.hpp
int dh_updater_save_state(int u, const char* file) {
    DH_TRY
    online::saveState(g_updaters.at(u).info, file);
    return 0;
    DH_CATCH(-1)
}
int dh_updater_load_state(int u, const char* file) {
    DH_TRY
    online::loadState(g_updaters.at(u).info, file);
    return 0;
    DH_CATCH(-1)
}
int64_t dh_updater_state_header_bytes(int u) {
    DH_TRY
    return (int64_t)online::stateHeaderBytes(g_updaters.at(u).info);
    DH_CATCH(-1)
}
int dh_updater_stats(int u, int64_t* launches, int64_t* bytes, int64_t* nodes) {
    DH_TRY
    auto bp = std::dynamic_pointer_cast<cuda::B200Plan>(g_updaters.at(u).info.plan);
    enforce(bp != nullptr, "updater is not backed by a B200Plan");
    bp->stats(launches, bytes, nodes);
    return 0;
    DH_CATCH(-1)
}
int dh_updater_replay_class(int u, const char* op_types, int reps, double* usec, int64_t* launches) {
    DH_TRY
    auto bp = std::dynamic_pointer_cast<cuda::B200Plan>(g_updaters.at(u).info.plan);
    enforce(bp != nullptr, "updater is not backed by a B200Plan");
    *usec = bp->replayClass(op_types, reps, launches);
    return 0;
    DH_CATCH(-1)
}
const char* dh_updater_profile(int u, int enable) {
    try {
        auto bp = std::dynamic_pointer_cast<cuda::B200Plan>(g_updaters.at(u).info.plan);
        enforce(bp != nullptr, "updater is not backed by a B200Plan");
        g_text = bp->profile(enable != 0);
        return g_text.c_str();
    } catch (const std::exception& e) {
        g_error = e.what();
        return nullptr;
    }
}

}  // extern "C"
