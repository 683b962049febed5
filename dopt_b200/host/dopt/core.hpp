// dopt/core.hpp -- C++ mirror of dopt.core (the host side of the hot path): the Operation graph, the operation and
// gradient registries, reverse-mode autodiff and the Plan / DeviceBuffer abstractions.
//
// Why this exists: dopt's host code is D and stays D (see INTEGRATION.md for the glue module).  This environment has no
// D compiler, so this mirror stands in for `dopt.core` with the SAME names, argument meaning and error behaviour, so that
// models built with dopt.nnet and trained through dopt.online can be driven end to end against libdopt_b200.so and the
// parity tests read like the reference's own unit tests.  Each function cites the D source it follows
// (paths relative to the dopt tree).  It contains no kernels and no numerics: every evaluation goes through the C ABI
// of include/dopt_b200.h.
#pragma once
#include <cstddef>
#include <cstdint>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace dopt {

struct Exception : std::runtime_error {
    using std::runtime_error::runtime_error;
};
void enforce(bool cond, const std::string& msg);   // std.exception.enforce

// ---- core/source/dopt/core/types.d ---------------------------------------------------------------------------------
enum class DataType { float32 = 0, int32 = 1 };
inline size_t sizeOf(DataType) { return 4; }

struct TensorType {
    DataType elementType = DataType::float32;
    std::vector<size_t> shape;
    TensorType() {}
    TensorType(DataType t, std::vector<size_t> s) : elementType(t), shape(std::move(s)) {}
    size_t rank() const { return shape.size(); }
    size_t volume() const {
        size_t v = 1;
        for (auto s : shape) v *= s;
        return v;
    }
    bool operator==(const TensorType& o) const { return elementType == o.elementType && shape == o.shape; }
};

// DeviceBuffer, core/source/dopt/core/types.d:59-77
class DeviceBuffer {
public:
    virtual ~DeviceBuffer() {}
    virtual size_t numBytes() const = 0;
    virtual void set(const void* buf, size_t bytes) = 0;   // host -> buffer
    virtual void set(const DeviceBuffer& other);           // buffer -> buffer (default: through the host)
    virtual void get(void* buf, size_t bytes) const = 0;   // buffer -> host
    template <class T> std::vector<T> get() const {
        std::vector<T> v(numBytes() / sizeof(T));
        get(v.data(), v.size() * sizeof(T));
        return v;
    }
};
using Buffer = std::shared_ptr<DeviceBuffer>;

// plain host memory: what dopt.cpu's CPUBuffer is (cpu/source/dopt/cpu/package.d:34-71); `buffer(fs)` makes these
class HostBuffer : public DeviceBuffer {
public:
    explicit HostBuffer(size_t bytes) : mData(bytes, 0), mExt(nullptr), mExtBytes(0) {}
    // non-owning view of caller memory (e.g. a pinned staging buffer): no copy is made
    HostBuffer(void* external, size_t bytes) : mExt((uint8_t*)external), mExtBytes(bytes) {}
    size_t numBytes() const override { return mExt ? mExtBytes : mData.size(); }
    void set(const void* buf, size_t bytes) override;
    void get(void* buf, size_t bytes) const override;
    const void* raw() const { return mExt ? mExt : mData.data(); }
    void* raw() { return mExt ? mExt : mData.data(); }
private:
    std::vector<uint8_t> mData;
    uint8_t* mExt;
    size_t mExtBytes;
};

// ---- attributes (std.variant.Variant in D) ---------------------------------------------------------------------------
struct Variant {
    enum Kind { Empty, Sizes, Size, Double, Type } kind = Empty;
    std::vector<size_t> sizes;
    size_t size = 0;
    double real = 0;
    TensorType type;
    Variant() {}
    Variant(std::vector<size_t> v) : kind(Sizes), sizes(std::move(v)) {}
    Variant(std::initializer_list<size_t> v) : kind(Sizes), sizes(v) {}
    Variant(size_t v) : kind(Size), size(v) {}
    Variant(double v) : kind(Double), real(v) {}
    Variant(TensorType t) : kind(Type), type(std::move(t)) {}
    const std::vector<size_t>& getSizes() const;   // .get!(size_t[]) -- throws on a kind mismatch like Variant does
    size_t getSize() const;
    double getDouble() const;
    const TensorType& getType() const;
};
using Attributes = std::map<std::string, Variant>;

// ---- core/source/dopt/core/ops/package.d:52-246 --------------------------------------------------------------------
class OperationNode;
using Operation = std::shared_ptr<OperationNode>;

class OperationNode : public std::enable_shared_from_this<OperationNode> {
public:
    const std::string& opType() const { return mOpType; }
    const TensorType& outputType() const { return mOutputType; }
    const std::vector<Operation>& deps() const { return mDeps; }
    const Attributes& attributes() const { return mAttributes; }
    const std::vector<size_t>& shape() const { return mOutputType.shape; }
    DataType elementType() const { return mOutputType.elementType; }
    size_t volume() const { return mOutputType.volume(); }
    size_t rank() const { return mOutputType.rank(); }
    Buffer value() const { return mBuffer; }
    void setBuffer(Buffer b) { mBuffer = std::move(b); }
    uint64_t id() const { return mId; }   // creation serial (stable identity for export; D uses the object address)

    OperationNode(std::string opType, std::vector<Operation> deps, Attributes attribs);   // verifies + judges
private:
    std::string mOpType;
    std::vector<Operation> mDeps;
    Attributes mAttributes;
    TensorType mOutputType;
    Buffer mBuffer;
    uint64_t mId;
};

using Verifier = std::function<bool(const OperationNode&)>;
using Judge = std::function<TensorType(const OperationNode&)>;
struct OpDef {
    Verifier verifier;
    Judge judge;
};
void registerOperation(const std::string& name, OpDef def);          // ops/package.d:253-258 (throws if taken)
std::vector<std::string> listAllOperations();                        // ops/package.d:263-266
Operation createOperation(const std::string& opType, std::vector<Operation> deps = {}, Attributes attribs = {});
std::vector<Operation> topologicalSort(const std::vector<Operation>& ops);   // ops/package.d:284-311

// ---- core/source/dopt/core/ops/basic.d ----------------------------------------------------------------------------------
Operation slice(Operation input, std::vector<size_t> start, std::vector<size_t> stop);
Operation pad(Operation input, std::vector<size_t> before, std::vector<size_t> after);
Operation reshape(Operation input, std::vector<size_t> shape);
Operation transpose(Operation input, std::vector<size_t> order);
Operation repeat(Operation input, std::vector<size_t> repetitions);   // per-axis
Operation repeat(Operation input, size_t repetitions);                // new leading axis, lowered to matmul (basic.d:370-381)
Operation variable(TensorType type, const void* defaultVal = nullptr);
Operation float32(std::vector<size_t> size = {}, const std::vector<float>& defaultVal = {});
Operation float32Scalar(float defaultVal);   // D: float32(float) -- renamed: a braced size list must never bind to it
Operation int32(std::vector<size_t> size = {}, const std::vector<int>& defaultVal = {});
Operation constant(TensorType type, const void* val);
Operation float32Constant(std::vector<size_t> size, const std::vector<float>& val);
Operation float32Constant(float val);
Operation int32Constant(std::vector<size_t> size, const std::vector<int>& val);
Operation int32Constant(int val);

// ---- core/source/dopt/core/ops/math.d -----------------------------------------------------------------------------------
#define DOPT_DECL_BIN(name) Operation name(Operation a, Operation b);
#define DOPT_DECL_UN(name) Operation name(Operation a);
DOPT_DECL_BIN(add) DOPT_DECL_BIN(sub) DOPT_DECL_BIN(mul) DOPT_DECL_BIN(div)
DOPT_DECL_BIN(lt) DOPT_DECL_BIN(lte) DOPT_DECL_BIN(gt) DOPT_DECL_BIN(gte) DOPT_DECL_BIN(eq) DOPT_DECL_BIN(neq)
DOPT_DECL_BIN(max) DOPT_DECL_BIN(min) DOPT_DECL_BIN(pow)
DOPT_DECL_UN(neg) DOPT_DECL_UN(abs) DOPT_DECL_UN(sgn) DOPT_DECL_UN(exp) DOPT_DECL_UN(log) DOPT_DECL_UN(sqrt)
DOPT_DECL_UN(sin) DOPT_DECL_UN(cos) DOPT_DECL_UN(tan) DOPT_DECL_UN(asin) DOPT_DECL_UN(acos) DOPT_DECL_UN(atan)
DOPT_DECL_UN(sinh) DOPT_DECL_UN(cosh) DOPT_DECL_UN(tanh) DOPT_DECL_UN(asinh) DOPT_DECL_UN(acosh) DOPT_DECL_UN(atanh)
#undef DOPT_DECL_BIN
#undef DOPT_DECL_UN
Operation matmul(Operation lhs, Operation rhs);
Operation sum(Operation op, std::vector<size_t> axes = {});
Operation argmin(Operation input, size_t axis);
Operation maxElement(Operation op, std::vector<size_t> axes = {});

// Operation.opBinary / opBinaryRight / opUnary (ops/package.d:94-178): rank-0 operands are broadcast through
// repeat(volume).reshape(shape), which is a matmul with a ones column.
Operation operator+(Operation a, Operation b);
Operation operator-(Operation a, Operation b);
Operation operator*(Operation a, Operation b);
Operation operator/(Operation a, Operation b);
Operation operator+(Operation a, float b);
Operation operator-(Operation a, float b);
Operation operator*(Operation a, float b);
Operation operator/(Operation a, float b);
Operation operator+(float a, Operation b);
Operation operator*(float a, Operation b);
Operation operator-(float a, Operation b);
Operation operator/(float a, Operation b);
Operation operator-(Operation a);

// ---- core/source/dopt/core/ops/nnet.d -----------------------------------------------------------------------------------
Operation convolution(Operation features, Operation filters, std::vector<size_t> padding = {0, 0},
                      std::vector<size_t> stride = {1, 1});
Operation convolutionTranspose(Operation features, Operation filters, std::vector<size_t> padding = {0, 0},
                               std::vector<size_t> stride = {1, 1});
Operation maxpool(Operation features, std::vector<size_t> dims);
Operation convolutionFeaturesGrad(Operation parentGrad, Operation filters, std::vector<size_t> featuresShape,
                                  std::vector<size_t> padding, std::vector<size_t> stride);
Operation convolutionFiltersGrad(Operation parentGrad, Operation features, std::vector<size_t> filtersShape,
                                 std::vector<size_t> padding, std::vector<size_t> stride);
Operation maxpoolGrad(Operation parentGrad, Operation op);
Operation softmax(Operation inputs);
Operation softmaxGrad(Operation parentGrad, Operation op);
Operation relu(Operation inputs);
Operation reluGrad(Operation parentGrad, Operation op);
Operation addBias(Operation input, Operation bias);
Operation addBiasGrad(Operation parentGrad);
std::vector<Operation> batchNormTrain(Operation input, Operation scale, Operation bias, Operation mean, Operation var,
                                      double momentum);
Operation batchNormGrad(Operation parentGrad, Operation input, Operation scale);
Operation batchNormInference(Operation input, Operation scale, Operation bias, Operation mean, Operation var);
// core/source/dopt/core/ops/random.d
Operation uniformSample(std::vector<size_t> shape);

// ---- core/source/dopt/core/grads/package.d ----------------------------------------------------------------------------
using Gradient = std::function<std::vector<Operation>(Operation op, Operation parentGrad)>;
void registerGradient(const std::string& opName, Gradient g);     // grads/package.d:131-136
void deregisterGradient(const std::string& opName);               // grads/package.d:138-141
std::vector<Operation> grad(Operation objective, const std::vector<Operation>& wrt);   // grads/package.d:39-110

// data-parallel hook (new): when world_size > 1 every gradient returned to dopt.online is wrapped in an `allreduce`
// node (mean over ranks).  The reference is single-device; see DESIGN.md section "multi-GPU".
void setDataParallelWorld(int world_size);
int dataParallelWorld();

// ---- core/source/dopt/core/package.d ----------------------------------------------------------------------------------
class Plan {
public:
    explicit Plan(std::vector<Operation> outputs) : mOutputs(std::move(outputs)) {}
    virtual ~Plan() {}
    // core/package.d:152-176
    std::vector<Buffer> execute(const std::map<Operation, Buffer>& args = {});
    void execute(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets);
    const std::vector<Operation>& outputs() const { return mOutputs; }
protected:
    virtual void executeImpl(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets) = 0;
    std::vector<Operation> mOutputs;
};
using PlanPtr = std::shared_ptr<Plan>;
using Evaluator = std::function<std::vector<Buffer>(const std::vector<Operation>&, const std::map<Operation, Buffer>&)>;
using Compiler = std::function<PlanPtr(const std::vector<Operation>&)>;
using Allocator = std::function<Buffer(size_t)>;

// process-global defaults, core/package.d:31-69
Evaluator& defaultEvaluator();
Compiler& defaultCompiler();
Allocator& defaultVarAllocator();
Allocator& defaultArgAllocator();

std::vector<Buffer> evaluate(const std::vector<Operation>& ops, const std::map<Operation, Buffer>& args = {});
Buffer evaluate(Operation op, const std::map<Operation, Buffer>& args = {});
PlanPtr compile(const std::vector<Operation>& outputs);
Buffer allocate(size_t numBytes);
Buffer buffer(const void* vals, size_t bytes);            // core/package.d:133-139 (host-side argument buffer)
Buffer buffer(const std::vector<float>& vals);

void initialize();   // == the `shared static this()` chain of dopt.core (core/package.d:71-77)

}  // namespace dopt
