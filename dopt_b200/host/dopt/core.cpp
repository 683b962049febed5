// dopt/core.cpp -- see core.hpp.  Mirrors core/source/dopt/core/{types,package}.d, ops/{package,basic,math,nnet,random}.d
// and grads/{package,basic,math,nnet}.d of the reference.
#include "core.hpp"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <numeric>

namespace dopt {

void enforce(bool cond, const std::string& msg) {
    if (!cond) throw Exception(msg);
}

// ---- buffers ----------------------------------------------------------------------------------------------------------
void DeviceBuffer::set(const DeviceBuffer& other) {
    enforce(numBytes() == other.numBytes(), "Mismatch in buffer size");
    std::vector<uint8_t> tmp(other.numBytes());
    other.get(tmp.data(), tmp.size());
    set(tmp.data(), tmp.size());
}
void HostBuffer::set(const void* buf, size_t bytes) {
    enforce(bytes == numBytes(), "input buffer is the wrong length.");
    if (bytes) std::memcpy(raw(), buf, bytes);
}
void HostBuffer::get(void* buf, size_t bytes) const {
    enforce(bytes == numBytes(), "output buffer is the wrong length.");
    if (bytes) std::memcpy(buf, raw(), bytes);
}

// ---- Variant ------------------------------------------------------------------------------------------------------------
const std::vector<size_t>& Variant::getSizes() const {
    enforce(kind == Sizes, "Variant: attribute is not a size_t[]");
    return sizes;
}
size_t Variant::getSize() const {
    enforce(kind == Size, "Variant: attribute is not a size_t");
    return size;
}
double Variant::getDouble() const {
    enforce(kind == Double, "Variant: attribute is not a double");
    return real;
}
const TensorType& Variant::getType() const {
    enforce(kind == Type, "Variant: attribute is not a TensorType");
    return type;
}

// ---- registry -----------------------------------------------------------------------------------------------------------
static std::map<std::string, OpDef>& opDefs() {
    static std::map<std::string, OpDef> m;
    return m;
}
static std::map<std::string, Gradient>& gradients() {
    static std::map<std::string, Gradient> m;
    return m;
}
static std::atomic<uint64_t> g_serial{1};

void registerOperation(const std::string& name, OpDef def) {
    enforce(opDefs().find(name) == opDefs().end(), "There is already an operation registered with the name '" + name + "'");
    opDefs()[name] = std::move(def);
}
std::vector<std::string> listAllOperations() {
    std::vector<std::string> r;
    for (auto& kv : opDefs()) r.push_back(kv.first);
    return r;
}

OperationNode::OperationNode(std::string opType, std::vector<Operation> deps, Attributes attribs)
    : mOpType(std::move(opType)), mDeps(std::move(deps)), mAttributes(std::move(attribs)), mId(g_serial++) {
    auto it = opDefs().find(mOpType);
    enforce(it != opDefs().end(), "Cannot make judgement for unknown operation '" + mOpType + "'");
    for (auto& d : mDeps) enforce(d != nullptr, "Operation of type \"" + mOpType + "\" has a null dependency");
    enforce(it->second.verifier(*this), "Operation of type \"" + mOpType + "\" failed verification.");
    mOutputType = it->second.judge(*this);
}

Operation createOperation(const std::string& opType, std::vector<Operation> deps, Attributes attribs) {
    initialize();
    enforce(opDefs().find(opType) != opDefs().end(),
            "Cannot create operation because there is no operation definition registered with the name '" + opType + "'");
    return std::make_shared<OperationNode>(opType, std::move(deps), std::move(attribs));
}

std::vector<Operation> topologicalSort(const std::vector<Operation>& ops) {
    // same order as the reference's recursive post-order walk (ops/package.d:284-311); a visited set replaces its
    // O(n^2) canFind
    std::vector<Operation> sorted;
    std::map<const OperationNode*, bool> seen;
    std::function<void(const Operation&)> visit = [&](const Operation& o) {
        if (seen[o.get()]) return;
        seen[o.get()] = true;
        for (auto& d : o->deps()) visit(d);
        sorted.push_back(o);
    };
    // iterative-safe depth is fine: graphs here are a few thousand nodes deep at most along one chain
    for (auto& o : ops) visit(o);
    return sorted;
}

// ---- helpers --------------------------------------------------------------------------------------------------------------
static const std::vector<size_t>& attrSizes(const OperationNode& op, const char* name) {
    auto it = op.attributes().find(name);
    enforce(it != op.attributes().end(), std::string("missing attribute '") + name + "'");
    return it->second.getSizes();
}
static bool hasSizes(const OperationNode& op, const char* name) {
    auto it = op.attributes().find(name);
    return it != op.attributes().end() && it->second.kind == Variant::Sizes;
}
static size_t prod(const std::vector<size_t>& v, size_t from = 0) {
    size_t p = 1;
    for (size_t i = from; i < v.size(); ++i) p *= v[i];
    return p;
}

// ---- ops/basic.d ----------------------------------------------------------------------------------------------------------
static void initBasic() {
    registerOperation("slice", {[](const OperationNode& op) {   // verifySlice, basic.d:35-58
                                    if (!hasSizes(op, "start") || !hasSizes(op, "stop") || op.deps().size() != 1) return false;
                                    auto& start = attrSizes(op, "start");
                                    auto& stop = attrSizes(op, "stop");
                                    auto& sh = op.deps()[0]->shape();
                                    if (start.size() != stop.size() || start.size() != sh.size()) return false;
                                    for (size_t i = 0; i < sh.size(); ++i)
                                        if (!(start[i] < sh[i] && stop[i] <= sh[i] && start[i] < stop[i])) return false;
                                    return true;
                                },
                                [](const OperationNode& op) {
                                    auto& start = attrSizes(op, "start");
                                    auto& stop = attrSizes(op, "stop");
                                    std::vector<size_t> shape;
                                    for (size_t i = 0; i < start.size(); ++i) shape.push_back(stop[i] - start[i]);
                                    return TensorType(op.deps()[0]->elementType(), shape);
                                }});
    registerOperation("pad", {[](const OperationNode& op) {   // verifyPad, basic.d:77-97
                                  if (!hasSizes(op, "before") || !hasSizes(op, "after") || op.deps().size() != 1) return false;
                                  return attrSizes(op, "before").size() == attrSizes(op, "after").size() &&
                                         attrSizes(op, "before").size() == op.deps()[0]->rank();
                              },
                              [](const OperationNode& op) {
                                  auto& b = attrSizes(op, "before");
                                  auto& a = attrSizes(op, "after");
                                  auto& sh = op.deps()[0]->shape();
                                  std::vector<size_t> shape;
                                  for (size_t i = 0; i < sh.size(); ++i) shape.push_back(b[i] + a[i] + sh[i]);
                                  return TensorType(op.deps()[0]->elementType(), shape);
                              }});
    registerOperation("reshape", {[](const OperationNode& op) {   // basic.d:116-124
                                      return op.deps().size() == 1 && hasSizes(op, "shape") &&
                                             prod(attrSizes(op, "shape")) == op.deps()[0]->volume();
                                  },
                                  [](const OperationNode& op) {
                                      return TensorType(op.deps()[0]->elementType(), attrSizes(op, "shape"));
                                  }});
    registerOperation("transpose", {[](const OperationNode& op) {   // basic.d:131-139
                                        if (op.deps().size() != 1 || !hasSizes(op, "order")) return false;
                                        auto o = attrSizes(op, "order");
                                        std::sort(o.begin(), o.end());
                                        if (o.size() != op.deps()[0]->rank()) return false;
                                        for (size_t i = 0; i < o.size(); ++i)
                                            if (o[i] != i) return false;
                                        return true;
                                    },
                                    [](const OperationNode& op) {
                                        std::vector<size_t> shape;
                                        for (auto x : attrSizes(op, "order")) shape.push_back(op.deps()[0]->shape()[x]);
                                        return TensorType(op.deps()[0]->elementType(), shape);
                                    }});
    registerOperation("repeat", {[](const OperationNode& op) {   // basic.d:155-167
                                     if (!hasSizes(op, "repetitions") || op.deps().size() != 1) return false;
                                     auto& r = attrSizes(op, "repetitions");
                                     if (r.size() != op.deps()[0]->rank()) return false;
                                     for (auto x : r)
                                         if (x == 0) return false;
                                     return true;
                                 },
                                 [](const OperationNode& op) {
                                     auto shape = op.deps()[0]->shape();
                                     auto& r = attrSizes(op, "repetitions");
                                     for (size_t i = 0; i < shape.size(); ++i) shape[i] *= r[i];
                                     return TensorType(op.deps()[0]->elementType(), shape);
                                 }});
    auto verifyVariable = [](const OperationNode& op) {   // basic.d:178-183
        auto it = op.attributes().find("type");
        return op.deps().empty() && it != op.attributes().end() && it->second.kind == Variant::Type;
    };
    auto judgeVariable = [](const OperationNode& op) { return op.attributes().at("type").getType(); };
    registerOperation("variable", {verifyVariable, judgeVariable});
    registerOperation("constant", {verifyVariable, judgeVariable});
}

Operation slice(Operation input, std::vector<size_t> start, std::vector<size_t> stop) {
    return createOperation("slice", {input}, {{"start", Variant(start)}, {"stop", Variant(stop)}});
}
Operation pad(Operation input, std::vector<size_t> before, std::vector<size_t> after) {
    return createOperation("pad", {input}, {{"before", Variant(before)}, {"after", Variant(after)}});
}
Operation reshape(Operation input, std::vector<size_t> shape) {
    return createOperation("reshape", {input}, {{"shape", Variant(shape)}});
}
Operation transpose(Operation input, std::vector<size_t> order) {
    return createOperation("transpose", {input}, {{"order", Variant(order)}});
}
Operation repeat(Operation input, std::vector<size_t> repetitions) {
    enforce(repetitions.size() == input->rank(), "The length of repetitions must be the same as the rank of the input.");
    return createOperation("repeat", {input}, {{"repetitions", Variant(repetitions)}});
}
Operation repeat(Operation input, size_t repetitions) {
    // basic.d:370-381: reshape to a row vector, multiply by a ones column, reshape to [repetitions] ~ shape
    auto vec = reshape(input, {1, input->volume()});
    auto pattern = float32Constant({repetitions, 1}, std::vector<float>(repetitions, 1.0f));
    auto r = matmul(pattern, vec);
    std::vector<size_t> shape{repetitions};
    for (auto s : input->shape()) shape.push_back(s);
    return reshape(r, shape);
}

Operation variable(TensorType type, const void* defaultVal) {
    size_t bufSize = type.volume() * sizeOf(type.elementType);
    auto op = createOperation("variable", {}, {{"type", Variant(type)}});
    auto buf = allocate(bufSize);
    std::vector<uint8_t> zeros;
    if (!defaultVal) {
        zeros.assign(bufSize, 0);
        defaultVal = zeros.data();
    }
    buf->set(defaultVal, bufSize);
    op->setBuffer(buf);
    return op;
}
Operation float32(std::vector<size_t> size, const std::vector<float>& defaultVal) {
    TensorType t(DataType::float32, std::move(size));
    if (!defaultVal.empty())
        enforce(defaultVal.size() == t.volume(), "The length of defaultVal does not match type.volume.");
    return variable(t, defaultVal.empty() ? nullptr : defaultVal.data());
}
Operation float32Scalar(float defaultVal) { return float32(std::vector<size_t>{}, std::vector<float>{defaultVal}); }
Operation int32(std::vector<size_t> size, const std::vector<int>& defaultVal) {
    TensorType t(DataType::int32, std::move(size));
    if (!defaultVal.empty())
        enforce(defaultVal.size() == t.volume(), "The length of defaultVal does not match type.volume.");
    return variable(t, defaultVal.empty() ? nullptr : defaultVal.data());
}
Operation constant(TensorType type, const void* val) {
    size_t bufSize = type.volume() * sizeOf(type.elementType);
    auto op = createOperation("constant", {}, {{"type", Variant(type)}});
    // constants keep their value on the host: plans upload them once (cuda/source/dopt/cuda/package.d:304-307)
    auto buf = std::make_shared<HostBuffer>(bufSize);
    std::vector<uint8_t> zeros;
    if (!val) {
        zeros.assign(bufSize, 0);
        val = zeros.data();
    }
    buf->set(val, bufSize);
    op->setBuffer(buf);
    return op;
}
Operation float32Constant(std::vector<size_t> size, const std::vector<float>& val) {
    TensorType t(DataType::float32, std::move(size));
    enforce(val.size() == t.volume(), "The length of val does not match type.volume.");
    return constant(t, val.data());
}
Operation float32Constant(float val) { return float32Constant({}, {val}); }
Operation int32Constant(std::vector<size_t> size, const std::vector<int>& val) {
    TensorType t(DataType::int32, std::move(size));
    enforce(val.size() == t.volume(), "The length of val does not match type.volume.");
    return constant(t, val.data());
}
Operation int32Constant(int val) { return int32Constant({}, {val}); }

// ---- ops/math.d ---------------------------------------------------------------------------------------------------------
static const char* kBinary[] = {"add", "sub", "mul", "div", "lt", "lte", "gt", "gte", "eq", "neq", "max", "min", "pow"};
static const char* kUnary[] = {"neg", "abs", "sgn", "exp", "log", "sqrt", "sin", "cos", "tan", "asin", "acos", "atan",
                               "atan2", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh"};

static bool verifyReduction(const OperationNode& op) {   // verifySum, math.d:144-160
    if (op.deps().size() != 1 || !hasSizes(op, "axes")) return false;
    auto axes = attrSizes(op, "axes");
    for (auto a : axes)
        if (a >= op.deps()[0]->rank()) return false;
    std::sort(axes.begin(), axes.end());
    return std::unique(axes.begin(), axes.end()) == axes.end();
}
static TensorType judgeReduction(const OperationNode& op) {   // judgeSum, math.d:162-175
    auto& axes = attrSizes(op, "axes");
    std::vector<size_t> shape;
    auto& sh = op.deps()[0]->shape();
    for (size_t i = 0; i < sh.size(); ++i)
        if (std::find(axes.begin(), axes.end(), i) == axes.end()) shape.push_back(sh[i]);
    return TensorType(op.deps()[0]->elementType(), shape);
}

static void initMath() {
    for (auto name : kBinary)
        registerOperation(name, {[](const OperationNode& op) {   // math.d:20-31
                                     return op.deps().size() == 2 && op.deps()[0]->outputType() == op.deps()[1]->outputType();
                                 },
                                 [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
    for (auto name : kUnary)
        registerOperation(name, {[](const OperationNode&) { return true; },
                                 [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
    registerOperation("matmul", {[](const OperationNode& op) {   // math.d:129-136
                                     return op.deps().size() == 2 && op.deps()[0]->rank() == 2 && op.deps()[1]->rank() == 2 &&
                                            op.deps()[0]->elementType() == op.deps()[1]->elementType() &&
                                            op.deps()[0]->shape()[1] == op.deps()[1]->shape()[0];
                                 },
                                 [](const OperationNode& op) {
                                     return TensorType(op.deps()[0]->elementType(),
                                                       {op.deps()[0]->shape()[0], op.deps()[1]->shape()[1]});
                                 }});
    registerOperation("sum", {verifyReduction, judgeReduction});
    registerOperation("maxElement", {verifyReduction, judgeReduction});
    registerOperation("argmin", {[](const OperationNode& op) {   // math.d:177-183
                                     auto it = op.attributes().find("axis");
                                     return op.deps().size() == 1 && it != op.attributes().end() &&
                                            it->second.kind == Variant::Size && it->second.size < op.deps()[0]->rank();
                                 },
                                 [](const OperationNode& op) {
                                     auto shape = op.deps()[0]->shape();
                                     shape[op.attributes().at("axis").getSize()] = 1;
                                     return TensorType(DataType::int32, shape);
                                 }});
}

#define DOPT_DEF_BIN(name) \
    Operation name(Operation a, Operation b) { return createOperation(#name, {a, b}); }
#define DOPT_DEF_UN(name) \
    Operation name(Operation a) { return createOperation(#name, {a}); }
DOPT_DEF_BIN(add) DOPT_DEF_BIN(sub) DOPT_DEF_BIN(mul) DOPT_DEF_BIN(div)
DOPT_DEF_BIN(lt) DOPT_DEF_BIN(lte) DOPT_DEF_BIN(gt) DOPT_DEF_BIN(gte) DOPT_DEF_BIN(eq) DOPT_DEF_BIN(neq)
DOPT_DEF_BIN(max) DOPT_DEF_BIN(min) DOPT_DEF_BIN(pow)
DOPT_DEF_UN(neg) DOPT_DEF_UN(abs) DOPT_DEF_UN(sgn) DOPT_DEF_UN(exp) DOPT_DEF_UN(log) DOPT_DEF_UN(sqrt)
DOPT_DEF_UN(sin) DOPT_DEF_UN(cos) DOPT_DEF_UN(tan) DOPT_DEF_UN(asin) DOPT_DEF_UN(acos) DOPT_DEF_UN(atan)
DOPT_DEF_UN(sinh) DOPT_DEF_UN(cosh) DOPT_DEF_UN(tanh) DOPT_DEF_UN(asinh) DOPT_DEF_UN(acosh) DOPT_DEF_UN(atanh)
#undef DOPT_DEF_BIN
#undef DOPT_DEF_UN

Operation matmul(Operation lhs, Operation rhs) { return createOperation("matmul", {lhs, rhs}); }

static std::vector<size_t> iota(size_t n) {
    std::vector<size_t> v(n);
    std::iota(v.begin(), v.end(), size_t(0));
    return v;
}

Operation sum(Operation op, std::vector<size_t> axes) {
    // math.d:243-282
    if (op->rank() == 0) return reshape(op, op->shape());
    if (axes.empty()) axes = iota(op->rank());
    // "Temporary speed enhancement: use BLAS to do row/col sums of matrices"
    if (op->rank() == 2 && axes.size() == 1) {
        if (axes[0] == 1) {
            auto ones = float32Constant({op->shape()[1], 1}, std::vector<float>(op->shape()[1], 1.0f));
            return reshape(matmul(op, ones), {op->shape()[0]});
        } else if (axes[0] == 0) {
            auto ones = float32Constant({1, op->shape()[0]}, std::vector<float>(op->shape()[0], 1.0f));
            return reshape(matmul(ones, op), {op->shape()[1]});
        } else {
            throw Exception("axes[0] must be less than op.rank");
        }
    }
    return createOperation("sum", {op}, {{"axes", Variant(axes)}});
}
Operation argmin(Operation input, size_t axis) { return createOperation("argmin", {input}, {{"axis", Variant(axis)}}); }
Operation maxElement(Operation op, std::vector<size_t> axes) {
    if (op->rank() == 0) return reshape(op, op->shape());
    if (axes.empty()) axes = iota(op->rank());
    return createOperation("maxElement", {op}, {{"axes", Variant(axes)}});
}

// Operation.opBinary, ops/package.d:94-125
template <class F> static Operation binaryBroadcast(Operation a, Operation b, F f) {
    if (b->rank() == 0 && a->rank() != 0) return f(a, reshape(repeat(b, a->volume()), a->shape()));
    if (a->rank() == 0 && b->rank() != 0) return f(reshape(repeat(a, b->volume()), b->shape()), b);
    return f(a, b);
}
Operation operator+(Operation a, Operation b) { return binaryBroadcast(a, b, [](Operation x, Operation y) { return add(x, y); }); }
Operation operator-(Operation a, Operation b) { return binaryBroadcast(a, b, [](Operation x, Operation y) { return sub(x, y); }); }
Operation operator*(Operation a, Operation b) { return binaryBroadcast(a, b, [](Operation x, Operation y) { return mul(x, y); }); }
Operation operator/(Operation a, Operation b) { return binaryBroadcast(a, b, [](Operation x, Operation y) { return div(x, y); }); }
Operation operator+(Operation a, float b) { return a + float32Constant(b); }   // ops/package.d:134-139
Operation operator-(Operation a, float b) { return a - float32Constant(b); }
Operation operator*(Operation a, float b) { return a * float32Constant(b); }
Operation operator/(Operation a, float b) { return a / float32Constant(b); }
Operation operator+(float a, Operation b) { return b + a; }                    // opBinaryRight: "*" and "+" commute (package.d:143-146)
Operation operator*(float a, Operation b) { return b * a; }
Operation operator-(float a, Operation b) { return float32Constant(a) - b; }   // package.d:147-150
Operation operator/(float a, Operation b) { return float32Constant(a) / b; }
Operation operator-(Operation a) { return neg(a); }

// ---- ops/nnet.d ---------------------------------------------------------------------------------------------------------
static void initNnet() {
    auto yes = [](const OperationNode&) { return true; };
    registerOperation("convolution", {[](const OperationNode& op) {   // verifyConvolution, nnet.d:43-66
                                          if (op.deps().size() != 2) return false;
                                          auto& imgs = op.deps()[0]->outputType();
                                          auto& filters = op.deps()[1]->outputType();
                                          if (imgs.rank() != 4 || filters.rank() != 4) return false;
                                          if (imgs.elementType != filters.elementType) return false;
                                          return imgs.shape[1] == filters.shape[1];
                                      },
                                      [](const OperationNode& op) {   // judgeConvolution, nnet.d:68-87
                                          auto& imgs = op.deps()[0]->shape();
                                          auto& f = op.deps()[1]->shape();
                                          auto& padding = attrSizes(op, "padding");
                                          auto& stride = attrSizes(op, "stride");
                                          size_t h = (imgs[2] + 2 * padding[0] - f[2]) / stride[0] + 1;
                                          size_t w = (imgs[3] + 2 * padding[1] - f[3]) / stride[1] + 1;
                                          return TensorType(op.deps()[0]->elementType(), {imgs[0], f[0], h, w});
                                      }});
    registerOperation("maxpool", {[](const OperationNode& op) {   // nnet.d:89-95
                                      return op.deps().size() == 1 && op.deps()[0]->rank() == 4 && hasSizes(op, "dims") &&
                                             attrSizes(op, "dims").size() == 2;
                                  },
                                  [](const OperationNode& op) {
                                      auto& d = attrSizes(op, "dims");
                                      auto& s = op.deps()[0]->shape();
                                      return TensorType(op.deps()[0]->elementType(), {s[0], s[1], s[2] / d[0], s[3] / d[1]});
                                  }});
    auto judgeFeaturesShape = [](const OperationNode& op) {
        return TensorType(op.deps()[0]->elementType(), attrSizes(op, "featuresShape"));
    };
    registerOperation("convolutionFeaturesGrad", {yes, judgeFeaturesShape});
    registerOperation("convolutionFiltersGrad", {yes, [](const OperationNode& op) {
                                                     return TensorType(op.deps()[0]->elementType(), attrSizes(op, "filtersShape"));
                                                 }});
    registerOperation("maxpoolGrad", {yes, judgeFeaturesShape});
    registerOperation("softmax", {[](const OperationNode& op) { return op.deps().size() == 1; },
                                  [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
    registerOperation("softmaxGrad", {[](const OperationNode& op) { return op.deps().size() == 2; },
                                      [](const OperationNode& op) { return op.deps()[1]->outputType(); }});
    registerOperation("relu", {[](const OperationNode& op) { return op.deps().size() == 1; },
                               [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
    registerOperation("reluGrad", {[](const OperationNode& op) { return op.deps().size() == 3; },
                                   [](const OperationNode& op) { return op.deps()[1]->outputType(); }});
    registerOperation("addBias", {yes, [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
    registerOperation("addBiasGrad", {yes, [](const OperationNode& op) {
                                          return TensorType(op.deps()[0]->elementType(), {op.deps()[0]->shape()[1]});
                                      }});
    registerOperation("batchNormTrain", {yes, [](const OperationNode& op) {   // nnet.d:222-225: packed y | mean | var
                                             return TensorType(op.deps()[0]->elementType(),
                                                               {op.deps()[0]->volume() + 2 * op.deps()[0]->shape()[1]});
                                         }});
    registerOperation("batchNormGrad", {yes, [](const OperationNode& op) {   // nnet.d:232-235 (over-allocated, survey F4)
                                            return TensorType(op.deps()[0]->elementType(),
                                                              {op.deps()[0]->volume() + op.deps()[1]->volume() +
                                                               op.deps()[2]->volume()});
                                        }});
    registerOperation("batchNormInference", {yes, [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
    // ops/random.d:17-36
    registerOperation("uniform", {[](const OperationNode& op) { return op.deps().empty() && hasSizes(op, "shape"); },
                                  [](const OperationNode& op) {
                                      return TensorType(DataType::float32, attrSizes(op, "shape"));
                                  }});
    // data-parallel gradient exchange (new op type registered through the same API; not in the reference)
    registerOperation("allreduce", {[](const OperationNode& op) { return op.deps().size() == 1; },
                                    [](const OperationNode& op) { return op.deps()[0]->outputType(); }});
}

Operation convolution(Operation features, Operation filters, std::vector<size_t> padding, std::vector<size_t> stride) {
    return createOperation("convolution", {features, filters}, {{"padding", Variant(padding)}, {"stride", Variant(stride)}});
}
Operation convolutionTranspose(Operation features, Operation filters, std::vector<size_t> padding,
                               std::vector<size_t> stride) {
    // nnet.d:305-315
    auto outShape = features->shape();
    for (size_t i = 2; i < outShape.size(); ++i) {
        outShape[i] -= 1;
        outShape[i] *= stride[i - 2];
        outShape[i] += filters->shape()[i] - 2 * padding[i - 2];
    }
    outShape[1] = filters->shape()[1];
    return convolutionFeaturesGrad(features, filters, outShape, padding, stride);
}
Operation maxpool(Operation features, std::vector<size_t> dims) {
    return createOperation("maxpool", {features}, {{"dims", Variant(dims)}});
}
Operation convolutionFeaturesGrad(Operation parentGrad, Operation filters, std::vector<size_t> featuresShape,
                                  std::vector<size_t> padding, std::vector<size_t> stride) {
    return createOperation("convolutionFeaturesGrad", {parentGrad, filters},
                           {{"featuresShape", Variant(featuresShape)}, {"padding", Variant(padding)}, {"stride", Variant(stride)}});
}
Operation convolutionFiltersGrad(Operation parentGrad, Operation features, std::vector<size_t> filtersShape,
                                 std::vector<size_t> padding, std::vector<size_t> stride) {
    return createOperation("convolutionFiltersGrad", {parentGrad, features},
                           {{"filtersShape", Variant(filtersShape)}, {"padding", Variant(padding)}, {"stride", Variant(stride)}});
}
Operation maxpoolGrad(Operation parentGrad, Operation op) {
    return createOperation("maxpoolGrad", {parentGrad, op, op->deps()[0]},
                           {{"featuresShape", Variant(op->deps()[0]->shape())}, {"dims", op->attributes().at("dims")}});
}
Operation softmax(Operation inputs) { return createOperation("softmax", {inputs}); }
Operation softmaxGrad(Operation parentGrad, Operation op) { return createOperation("softmaxGrad", {parentGrad, op}); }
Operation relu(Operation inputs) { return createOperation("relu", {inputs}); }
Operation reluGrad(Operation parentGrad, Operation op) {
    return createOperation("reluGrad", {parentGrad, op, op->deps()[0]});
}
Operation addBias(Operation input, Operation bias) { return createOperation("addBias", {input, bias}); }
Operation addBiasGrad(Operation parentGrad) { return createOperation("addBiasGrad", {parentGrad}); }
std::vector<Operation> batchNormTrain(Operation input, Operation scale, Operation bias, Operation mean, Operation var,
                                      double momentum) {
    // nnet.d:476-489: the running mean / variance are packed after the forward value and sliced back out
    auto bnop = createOperation("batchNormTrain", {input, scale, bias, mean, var}, {{"momentum", Variant(momentum)}});
    size_t V = input->volume(), C = input->shape()[1];
    return {reshape(slice(bnop, {0}, {V}), input->shape()), slice(bnop, {V}, {V + C}), slice(bnop, {V + C}, {V + 2 * C})};
}
Operation batchNormGrad(Operation parentGrad, Operation input, Operation scale) {
    return createOperation("batchNormGrad", {parentGrad, input, scale});
}
Operation batchNormInference(Operation input, Operation scale, Operation bias, Operation mean, Operation var) {
    return createOperation("batchNormInference", {input, scale, bias, mean, var});
}
Operation uniformSample(std::vector<size_t> shape) { return createOperation("uniform", {}, {{"shape", Variant(shape)}}); }

// ---- grads ----------------------------------------------------------------------------------------------------------------
void registerGradient(const std::string& opName, Gradient g) {
    enforce(gradients().find(opName) == gradients().end(), "A gradient is already registered for operation '" + opName + "'");
    gradients()[opName] = std::move(g);
}
void deregisterGradient(const std::string& opName) { gradients().erase(opName); }

static int g_dp_world = 1;
void setDataParallelWorld(int w) { g_dp_world = w < 1 ? 1 : w; }
int dataParallelWorld() { return g_dp_world; }

static void initGrads() {
    // grads/basic.d
    registerGradient("transpose", [](Operation op, Operation pg) -> std::vector<Operation> {
        auto& order = attrSizes(*op, "order");
        std::vector<size_t> newOrder(order.size());
        for (size_t x = 0; x < order.size(); ++x)
            newOrder[x] = (size_t)(std::find(order.begin(), order.end(), x) - order.begin());
        return {transpose(pg, newOrder)};
    });
    registerGradient("slice", [](Operation op, Operation pg) -> std::vector<Operation> {
        auto before = attrSizes(*op, "start");
        auto after = op->deps()[0]->shape();
        auto& stop = attrSizes(*op, "stop");
        for (size_t i = 0; i < after.size(); ++i) after[i] -= stop[i];
        return {pad(pg, before, after)};
    });
    registerGradient("pad", [](Operation op, Operation pg) -> std::vector<Operation> {
        auto start = attrSizes(*op, "before");
        auto stop = op->deps()[0]->shape();
        for (size_t i = 0; i < stop.size(); ++i) stop[i] += start[i];
        return {slice(pg, start, stop)};
    });
    registerGradient("reshape", [](Operation op, Operation pg) -> std::vector<Operation> {
        return {reshape(pg, op->deps()[0]->shape())};
    });
    registerGradient("repeat", [](Operation op, Operation pg) -> std::vector<Operation> {
        // grads/basic.d:82-95: interleave (reps, shape) and sum over the repetition axes
        auto& reps = attrSizes(*op, "repetitions");
        auto& sh = op->deps()[0]->shape();
        std::vector<size_t> tmpShape, axes;
        for (size_t i = 0; i < reps.size(); ++i) {
            tmpShape.push_back(reps[i]);
            tmpShape.push_back(sh[i]);
            axes.push_back(2 * i);
        }
        return {sum(reshape(pg, tmpShape), axes)};
    });
    // grads/math.d
    registerGradient("matmul", [](Operation op, Operation pg) -> std::vector<Operation> {
        return {matmul(pg, transpose(op->deps()[1], {1, 0})), matmul(transpose(op->deps()[0], {1, 0}), pg)};
    });
    registerGradient("sum", [](Operation op, Operation pg) -> std::vector<Operation> {
        if (op->volume() == 1) return {reshape(repeat(pg, op->deps()[0]->volume()), op->deps()[0]->shape())};
        auto& axes = attrSizes(*op, "axes");
        auto tmpShape = op->deps()[0]->shape();
        std::vector<size_t> reps(tmpShape.size(), 1);
        for (auto a : axes) {
            reps[a] = tmpShape[a];
            tmpShape[a] = 1;
        }
        return {repeat(reshape(pg, tmpShape), reps)};
    });
    registerGradient("add", [](Operation, Operation pg) -> std::vector<Operation> { return {pg, pg}; });
    registerGradient("sub", [](Operation, Operation pg) -> std::vector<Operation> { return {pg, neg(pg)}; });
    registerGradient("mul", [](Operation op, Operation pg) -> std::vector<Operation> {
        return {pg * op->deps()[1], pg * op->deps()[0]};
    });
    registerGradient("div", [](Operation op, Operation pg) -> std::vector<Operation> {
        return {pg / op->deps()[1], neg(pg * op->deps()[0]) / (op->deps()[1] * op->deps()[1])};
    });
    registerGradient("pow", [](Operation op, Operation pg) -> std::vector<Operation> {
        return {pg * op->deps()[1] * pow(op->deps()[0], op->deps()[1] - 1.0f), pg * op->deps()[1] * log(op->deps()[0])};
    });
    auto minmax = [](Operation op, Operation pg) -> std::vector<Operation> {
        return {eq(op->deps()[0], op) * pg, eq(op->deps()[1], op) * pg};
    };
    registerGradient("min", minmax);
    registerGradient("max", minmax);
    registerGradient("neg", [](Operation, Operation pg) -> std::vector<Operation> { return {neg(pg)}; });
    registerGradient("abs", [](Operation op, Operation pg) -> std::vector<Operation> { return {pg * sgn(op->deps()[0])}; });
    registerGradient("exp", [](Operation op, Operation pg) -> std::vector<Operation> { return {pg * op}; });
    registerGradient("log", [](Operation op, Operation pg) -> std::vector<Operation> { return {pg / op->deps()[0]}; });
    registerGradient("sqrt", [](Operation op, Operation pg) -> std::vector<Operation> { return {pg / op}; });
    // grads/nnet.d
    registerGradient("convolution", [](Operation op, Operation pg) -> std::vector<Operation> {
        auto& padding = attrSizes(*op, "padding");
        auto& stride = attrSizes(*op, "stride");
        return {convolutionFeaturesGrad(pg, op->deps()[1], op->deps()[0]->shape(), padding, stride),
                convolutionFiltersGrad(pg, op->deps()[0], op->deps()[1]->shape(), padding, stride)};
    });
    registerGradient("convolutionFeaturesGrad", [](Operation op, Operation pg) -> std::vector<Operation> {
        auto& padding = attrSizes(*op, "padding");
        auto& stride = attrSizes(*op, "stride");
        return {convolution(pg, op->deps()[1], padding, stride),
                convolutionFiltersGrad(op->deps()[0], pg, op->deps()[1]->shape(), padding, stride)};
    });
    registerGradient("maxpool", [](Operation op, Operation pg) -> std::vector<Operation> { return {maxpoolGrad(pg, op)}; });
    registerGradient("softmax", [](Operation op, Operation pg) -> std::vector<Operation> { return {softmaxGrad(pg, op)}; });
    registerGradient("relu", [](Operation op, Operation pg) -> std::vector<Operation> { return {reluGrad(pg, op)}; });
    registerGradient("addBias", [](Operation, Operation pg) -> std::vector<Operation> { return {pg, addBiasGrad(pg)}; });
    registerGradient("batchNormTrain", [](Operation op, Operation pg) -> std::vector<Operation> {
        // grads/nnet.d:66-83
        auto& d = op->deps();
        auto trimmed = reshape(slice(pg, {0}, {d[0]->volume()}), d[0]->shape());
        auto packed = batchNormGrad(trimmed, d[0], d[1]);
        packed = reshape(packed, {packed->volume()});
        size_t v0 = d[0]->volume(), v1 = d[1]->volume(), v2 = d[2]->volume();
        return {reshape(slice(packed, {0}, {v0}), d[0]->shape()), reshape(slice(packed, {v0}, {v0 + v1}), d[1]->shape()),
                reshape(slice(packed, {v0 + v1}, {v0 + v1 + v2}), d[2]->shape()), float32(d[3]->shape()),
                float32(d[4]->shape())};
    });
}

std::vector<Operation> grad(Operation objective, const std::vector<Operation>& wrt) {
    // grads/package.d:39-110
    initialize();
    enforce(objective->volume() == 1, "The objective must have a volume of one");
    enforce(objective->elementType() == DataType::float32, "The objective must have a floating point type");
    std::vector<Operation> ops = topologicalSort({objective});
    std::map<const OperationNode*, Operation> grads;
    grads[objective.get()] = float32(objective->shape(), {1.0f});
    for (auto it = ops.rbegin(); it != ops.rend(); ++it) {
        const Operation& op = *it;
        auto gf = gradients().find(op->opType());
        auto og = grads.find(op.get());
        if (gf == gradients().end() || og == grads.end()) continue;   // not differentiable: derivative assumed zero
        auto depGrads = gf->second(op, og->second);
        for (size_t i = 0; i < op->deps().size() && i < depGrads.size(); ++i) {
            const OperationNode* d = op->deps()[i].get();
            auto cur = grads.find(d);
            if (cur == grads.end()) grads[d] = depGrads[i];
            else cur->second = cur->second + depGrads[i];
        }
    }
    std::vector<Operation> result;
    for (size_t i = 0; i < wrt.size(); ++i) {
        auto g = grads.find(wrt[i].get());
        enforce(g != grads.end(), "Could not find wrt[" + std::to_string(i) + "] in the operation graph");
        result.push_back(g->second);
    }
    return result;
}

// ---- core/package.d -----------------------------------------------------------------------------------------------------
std::vector<Buffer> Plan::execute(const std::map<Operation, Buffer>& args) {
    std::vector<Buffer> rets;
    for (auto& o : mOutputs) rets.push_back(allocate(o->volume() * sizeOf(o->elementType())));
    executeImpl(args, rets);
    return rets;
}
void Plan::execute(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets) { executeImpl(args, rets); }

Evaluator& defaultEvaluator() {
    static Evaluator e = [](const std::vector<Operation>&, const std::map<Operation, Buffer>&) -> std::vector<Buffer> {
        throw Exception("no backend is loaded: dopt_b200 has no CPU evaluator (call dopt::cuda::initialize() on a B200)");
    };
    return e;
}
namespace {
// compiling without a backend succeeds (so graphs can be built and inspected anywhere); executing does not
struct NoBackendPlan : Plan {
    using Plan::Plan;
    void executeImpl(const std::map<Operation, Buffer>&, std::vector<Buffer>&) override {
        throw Exception("no backend is loaded: dopt_b200 has no CPU evaluator (call dopt::cuda::initialize() on a B200)");
    }
};
}  // namespace
Compiler& defaultCompiler() {
    static Compiler c = [](const std::vector<Operation>& ops) -> PlanPtr { return std::make_shared<NoBackendPlan>(ops); };
    return c;
}
Allocator& defaultVarAllocator() {
    static Allocator a = [](size_t n) -> Buffer { return std::make_shared<HostBuffer>(n); };
    return a;
}
Allocator& defaultArgAllocator() {
    static Allocator a = [](size_t n) -> Buffer { return std::make_shared<HostBuffer>(n); };
    return a;
}
std::vector<Buffer> evaluate(const std::vector<Operation>& ops, const std::map<Operation, Buffer>& args) {
    return defaultEvaluator()(ops, args);
}
Buffer evaluate(Operation op, const std::map<Operation, Buffer>& args) { return evaluate(std::vector<Operation>{op}, args)[0]; }
PlanPtr compile(const std::vector<Operation>& outputs) { return defaultCompiler()(outputs); }
Buffer allocate(size_t numBytes) { return defaultVarAllocator()(numBytes); }
Buffer buffer(const void* vals, size_t bytes) {
    auto b = defaultArgAllocator()(bytes);
    b->set(vals, bytes);
    return b;
}
Buffer buffer(const std::vector<float>& vals) { return buffer(vals.data(), vals.size() * sizeof(float)); }

void initialize() {
    static bool done = false;
    if (done) return;
    done = true;
    initBasic();
    initMath();
    initNnet();
    initGrads();
}

}  // namespace dopt
