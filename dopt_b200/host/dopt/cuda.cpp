// dopt/cuda.cpp -- see cuda.hpp.  Everything that touches the device goes through include/dopt_b200.h (kernels, plans,
// collectives) or the CUDA runtime's memory API (cudaMalloc / cudaMemcpy for CUDABuffer, like the reference's cuMemAlloc /
// cuMemcpy*).
#include "cuda.hpp"

#include <cuda_runtime.h>

#include <chrono>
#include <cstring>

#include "../../../include/dopt_b200.h"

namespace dopt {
namespace cuda {

static void* g_stream = nullptr;
// bf16 interior is part of the default (production) configuration together with MATH_BF16; setPlanFlags / setMath(MATH_FP32)
// select the fp32-storage and strict-fp32 variants (include/dopt_b200.h documents the numerics of each)
static int g_plan_flags = DOPT_B200_PLAN_FUSE | DOPT_B200_PLAN_CUDA_GRAPH | DOPT_B200_PLAN_BF16_INTERIOR;
static int g_math = DOPT_B200_MATH_DEFAULT;
static std::string g_init_error;

void* currentStream() { return g_stream; }
void setStream(void* s) { g_stream = s; }
void setPlanFlags(int f) { g_plan_flags = f; }
int planFlags() { return g_plan_flags; }
void setMath(int m) { g_math = m; }
const std::string& lastInitError() { return g_init_error; }

static void cudaEnforce(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw Exception(std::string(what) + ": " + cudaGetErrorString(e));
}
static void abiEnforce(int rc) {
    if (rc != 0) throw Exception(dopt_b200_last_error());
}

// ---- CUDABuffer -----------------------------------------------------------------------------------------------------------
std::shared_ptr<CUDABuffer> CUDABuffer::create(size_t numBytes) {
    std::shared_ptr<CUDABuffer> ret(new CUDABuffer());
    if (numBytes == 0) return ret;
    ret->mNumBytes = numBytes;
    cudaEnforce(cudaMalloc(&ret->mPtr, numBytes),
                ("CUDA memory allocation failed: unable to allocate " + std::to_string(numBytes) + " bytes").c_str());
    cudaEnforce(cudaMemset(ret->mPtr, 0, numBytes), "CUDA default buffer initialisation failed");
    return ret;
}
CUDABuffer::~CUDABuffer() {
    if (mPtr) cudaFree(mPtr);
}
void CUDABuffer::set(const void* buf, size_t bytes) {
    enforce(bytes == mNumBytes, "input buffer is the wrong length.");
    if (bytes) cudaEnforce(cudaMemcpy(mPtr, buf, bytes, cudaMemcpyHostToDevice), "Failed to set contents of CUDA buffer");
}
void CUDABuffer::set(const DeviceBuffer& other) {
    enforce(numBytes() == other.numBytes(), "Mismatch in buffer size");
    if (auto cu = dynamic_cast<const CUDABuffer*>(&other)) {
        if (mNumBytes) cudaEnforce(cudaMemcpy(mPtr, cu->ptr(), mNumBytes, cudaMemcpyDeviceToDevice), "cuMemcpyDtoD failed");
    } else if (auto h = dynamic_cast<const HostBuffer*>(&other)) {
        set(h->raw(), h->numBytes());
    } else {
        DeviceBuffer::set(other);
    }
}
void CUDABuffer::get(void* buf, size_t bytes) const {
    enforce(bytes == mNumBytes, "output buffer is the wrong length.");
    if (bytes) cudaEnforce(cudaMemcpy(buf, mPtr, bytes, cudaMemcpyDeviceToHost), "Failed to get contents of CUDA buffer");
}

// ---- op description -----------------------------------------------------------------------------------------------------
static void fillTensor(dopt_b200_tensor& t, const TensorType& ty) {
    enforce(ty.rank() <= DOPT_B200_MAX_RANK, "tensor rank exceeds DOPT_B200_MAX_RANK");
    t.dtype = ty.elementType == DataType::float32 ? DOPT_B200_FLOAT32 : DOPT_B200_INT32;
    t.rank = (int32_t)ty.rank();
    for (size_t i = 0; i < ty.rank(); ++i) t.shape[i] = (int64_t)ty.shape[i];
}
static void copySizes(int64_t* dst, const Attributes& a, const char* name, size_t maxn) {
    auto it = a.find(name);
    if (it == a.end() || it->second.kind != Variant::Sizes) return;
    enforce(it->second.sizes.size() <= maxn, std::string("attribute '") + name + "' too long");
    for (size_t i = 0; i < it->second.sizes.size(); ++i) dst[i] = (int64_t)it->second.sizes[i];
}
// everything the reference kernels read from `op` at construction (op.attributes["padding"] ...), as a POD
static dopt_b200_op describe(const OperationNode& op) {
    dopt_b200_op d;
    std::memset(&d, 0, sizeof(d));
    d.op_type = op.opType().c_str();
    enforce(op.deps().size() <= DOPT_B200_MAX_INPUTS, "too many operands");
    d.n_inputs = (int32_t)op.deps().size();
    for (size_t i = 0; i < op.deps().size(); ++i) fillTensor(d.inputs[i], op.deps()[i]->outputType());
    fillTensor(d.output, op.outputType());
    d.stride[0] = d.stride[1] = 1;
    auto& a = op.attributes();
    copySizes(d.padding, a, "padding", 2);
    copySizes(d.stride, a, "stride", 2);
    copySizes(d.pool_dims, a, "dims", 2);
    copySizes(d.start, a, "start", DOPT_B200_MAX_RANK);
    copySizes(d.stop, a, "stop", DOPT_B200_MAX_RANK);
    copySizes(d.before, a, "before", DOPT_B200_MAX_RANK);
    copySizes(d.after, a, "after", DOPT_B200_MAX_RANK);
    copySizes(d.repetitions, a, "repetitions", DOPT_B200_MAX_RANK);
    copySizes(d.order, a, "order", DOPT_B200_MAX_RANK);
    copySizes(d.axes, a, "axes", DOPT_B200_MAX_RANK);
    auto ax = a.find("axes");
    if (ax != a.end() && ax->second.kind == Variant::Sizes) d.n_axes = (int32_t)ax->second.sizes.size();
    auto axis = a.find("axis");
    if (axis != a.end() && axis->second.kind == Variant::Size) d.axis = (int64_t)axis->second.size;
    auto mom = a.find("momentum");
    if (mom != a.end() && mom->second.kind == Variant::Double) d.momentum = mom->second.real;
    d.math = g_math;
    return d;
}

// ---- kernel registry ------------------------------------------------------------------------------------------------------
static std::map<std::string, CUDAKernelCtr>& kernelCtrs() {
    static std::map<std::string, CUDAKernelCtr> m;
    return m;
}
void registerCUDAKernel(const std::string& opName, CUDAKernelCtr ctr) {
    enforce(kernelCtrs().find(opName) == kernelCtrs().end(),
            "A CUDAKernelCtr is already registered for the operation '" + opName + "'");
    kernelCtrs()[opName] = std::move(ctr);
}
void deregisterCUDAKernel(const std::string& opType) { kernelCtrs().erase(opType); }
std::vector<std::string> listCUDAOperations() {
    std::vector<std::string> r;
    for (auto& kv : kernelCtrs()) r.push_back(kv.first);
    r.push_back("variable");
    r.push_back("reshape");
    return r;
}

// the one CUDAKernel class of the glue: holds a library handle
class B200Kernel : public CUDAKernel {
public:
    explicit B200Kernel(Operation op) {
        dopt_b200_op d = describe(*op);
        abiEnforce(dopt_b200_kernel_create(&d, &mHandle));
    }
    ~B200Kernel() override { dopt_b200_kernel_destroy(mHandle); }
    void execute(const std::vector<const CUDABuffer*>& inputs, CUDABuffer& output) override {
        const void* in[DOPT_B200_MAX_INPUTS];
        for (size_t i = 0; i < inputs.size(); ++i) in[i] = inputs[i]->ptr();
        abiEnforce(dopt_b200_kernel_execute(mHandle, in, (int)inputs.size(), output.ptr(), g_stream));
    }
private:
    dopt_b200_kernel_t mHandle = nullptr;
};

// ---- CUDAPlan: node by node, as the reference (package.d:267-312, 343-424) ---------------------------------------------------
CUDAPlan::CUDAPlan(std::vector<Operation> outputs) : Plan(std::move(outputs)) {
    auto sorted = topologicalSort(mOutputs);
    for (auto& o : sorted) {
        if (o->opType() == "variable" || o->opType() == "reshape" || o->opType() == "constant") continue;
        auto k = kernelCtrs().find(o->opType());
        // the reference falls back to a D2H -> CPU -> H2D wrapper here (package.d:284); this backend refuses instead
        enforce(k != kernelCtrs().end(), "Could not construct a CUDA kernel for operation of type '" + o->opType() + "'");
        mKernels[o.get()] = k->second(o);
    }
    mOps = sorted;
    for (auto& o : mOps) {
        if (o->opType() == "reshape") {
            mResults[o.get()] = mResults[o->deps()[0].get()];
        } else {
            mResults[o.get()] = CUDABuffer::create(o->volume() * sizeOf(o->elementType()));
            if (o->opType() == "constant") mResults[o.get()]->set(*o->value());
        }
    }
}

void CUDAPlan::executeImpl(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets) {
    for (auto& kv : args)
        enforce(kv.first->opType() == "variable",
                "All assignments in args must be for Operations with an opType of 'variable'");
    for (auto& o : mOps) {
        if (o->opType() == "variable" || o->opType() == "constant") continue;
        std::vector<const CUDABuffer*> inputs;
        for (auto& d : o->deps()) {
            if (d->opType() == "variable") {
                std::shared_ptr<CUDABuffer> cubuf;
                auto it = args.find(d);
                Buffer src = it != args.end() ? it->second : d->value();
                cubuf = std::dynamic_pointer_cast<CUDABuffer>(src);
                if (!cubuf) {
                    cubuf = mResults[d.get()];
                    cubuf->set(*src);   // host buffer: H2D into the plan's own copy (package.d:373-381)
                }
                mResults[d.get()] = cubuf;
                inputs.push_back(cubuf.get());
            } else {
                inputs.push_back(mResults[d.get()].get());
            }
        }
        if (o->opType() == "reshape") {
            mResults[o.get()] = mResults[o->deps()[0].get()];
        } else {
            auto t0 = std::chrono::steady_clock::now();
            mKernels[o.get()]->execute(inputs, *mResults[o.get()]);
            profiler[o->opType()] +=
                (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - t0).count();
        }
    }
    enforce(rets.size() == mOutputs.size(), "wrong number of return buffers");
    cudaEnforce(cudaStreamSynchronize((cudaStream_t)g_stream), "stream synchronize");
    for (size_t i = 0; i < mOutputs.size(); ++i) {
        auto& o = mOutputs[i];
        std::shared_ptr<CUDABuffer> res;
        if (o->opType() == "variable") {
            auto it = args.find(o);
            Buffer src = it != args.end() ? it->second : o->value();
            rets[i]->set(*src);
            continue;
        }
        rets[i]->set(*mResults[o.get()]);
    }
}

// ---- B200Plan ---------------------------------------------------------------------------------------------------------------
B200Plan::B200Plan(std::vector<Operation> outputs, int flags) : Plan(std::move(outputs)) {
    dopt_b200_plan_t p = nullptr;
    abiEnforce(dopt_b200_plan_create(&p));
    mPlan = p;
    try {
        auto sorted = topologicalSort(mOutputs);
        for (auto& o : sorted) {
            dopt_b200_op d = describe(*o);
            std::vector<int32_t> deps;
            for (auto& dep : o->deps()) deps.push_back(mIds.at(dep.get()));
            const void* cval = nullptr;
            std::vector<uint8_t> tmp;
            if (o->opType() == "constant") {
                tmp.resize(o->value()->numBytes());
                o->value()->get(tmp.data(), tmp.size());
                cval = tmp.data();
            }
            int id = dopt_b200_plan_add_node(p, &d, deps.data(), (int)deps.size(), cval);
            if (id < 0) throw Exception(dopt_b200_last_error());
            mIds[o.get()] = id;
            if (o->opType() == "variable") mVariables.push_back(o);
        }
        std::vector<int32_t> outs;
        for (auto& o : mOutputs) outs.push_back(mIds.at(o.get()));
        abiEnforce(dopt_b200_plan_set_outputs(p, outs.data(), (int)outs.size()));
        abiEnforce(dopt_b200_plan_finalize(p, flags));
    } catch (...) {
        dopt_b200_plan_destroy(p);
        mPlan = nullptr;
        throw;
    }
}
B200Plan::~B200Plan() {
    if (mPlan) dopt_b200_plan_destroy((dopt_b200_plan_t)mPlan);
}
void B200Plan::stats(int64_t* launches, int64_t* deviceBytes, int64_t* loweredNodes) const {
    abiEnforce(dopt_b200_plan_stats((dopt_b200_plan_t)mPlan, launches, deviceBytes, loweredNodes));
}
std::string B200Plan::profile(bool enable) {
    std::vector<char> buf(1 << 16, 0);
    abiEnforce(dopt_b200_plan_profile((dopt_b200_plan_t)mPlan, enable ? 1 : 0, buf.data(), buf.size()));
    return std::string(buf.data());
}

double B200Plan::replayClass(const std::string& opTypes, int reps, int64_t* launches) {
    double us = 0;
    abiEnforce(dopt_b200_plan_replay_class((dopt_b200_plan_t)mPlan, opTypes.c_str(), reps, &us, launches, nullptr));
    return us;
}

void B200Plan::executeRaw(const std::vector<Operation>& argOps, const std::vector<const void*>& argPtrs,
                          const std::vector<int>& argOnHost, const std::vector<void*>& rets) {
    // variables not named in args are read from their own buffers (package.d:383-392)
    std::vector<int32_t> ids, onHost;
    std::vector<const void*> ptrs;
    std::map<const OperationNode*, size_t> given;
    for (size_t i = 0; i < argOps.size(); ++i) {
        enforce(argOps[i]->opType() == "variable",
                "All assignments in args must be for Operations with an opType of 'variable'");
        given[argOps[i].get()] = i;
    }
    mKeepAlive.clear();
    for (auto& v : mVariables) {
        auto g = given.find(v.get());
        ids.push_back(mIds.at(v.get()));
        if (g != given.end()) {
            ptrs.push_back(argPtrs[g->second]);
            onHost.push_back(argOnHost[g->second]);
        } else {
            Buffer val = v->value();
            if (auto cu = std::dynamic_pointer_cast<CUDABuffer>(val)) {
                ptrs.push_back(cu->ptr());
                onHost.push_back(0);
            } else if (auto h = std::dynamic_pointer_cast<HostBuffer>(val)) {
                ptrs.push_back(h->raw());
                onHost.push_back(1);
            } else {
                throw Exception("variable holds an unknown DeviceBuffer type");
            }
        }
    }
    abiEnforce(dopt_b200_plan_execute((dopt_b200_plan_t)mPlan, ids.data(), ptrs.data(), onHost.data(), (int)ids.size(),
                                      rets.data(), (int)rets.size(), g_stream));
    // data-parallel: the collectives inside the (replayed) step are asynchronous; surface a failed peer / link as an exception
    // at the next step instead of a hang (one host call, no synchronisation)
    if (dopt_b200_comm_world_size() > 1) abiEnforce(dopt_b200_comm_check());
}

void B200Plan::executeImpl(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets) {
    std::vector<Operation> argOps;
    std::vector<const void*> argPtrs;
    std::vector<int> onHost;
    for (auto& kv : args) {
        argOps.push_back(kv.first);
        if (auto cu = std::dynamic_pointer_cast<CUDABuffer>(kv.second)) {
            argPtrs.push_back(cu->ptr());
            onHost.push_back(0);
        } else if (auto h = std::dynamic_pointer_cast<HostBuffer>(kv.second)) {
            argPtrs.push_back(h->raw());
            onHost.push_back(1);
        } else {
            throw Exception("argument holds an unknown DeviceBuffer type");
        }
    }
    enforce(rets.size() == mOutputs.size(), "wrong number of return buffers");
    // rets must be device buffers for the in-library D2D copy; host rets get a device staging buffer
    std::vector<void*> retPtrs;
    std::vector<std::pair<size_t, std::shared_ptr<CUDABuffer>>> staged;
    for (size_t i = 0; i < rets.size(); ++i) {
        if (auto cu = std::dynamic_pointer_cast<CUDABuffer>(rets[i])) {
            retPtrs.push_back(cu->ptr());
        } else {
            auto tmp = CUDABuffer::create(rets[i]->numBytes());
            staged.push_back({i, tmp});
            retPtrs.push_back(tmp->ptr());
        }
    }
    executeRaw(argOps, argPtrs, onHost, retPtrs);
    if (!staged.empty()) {
        cudaEnforce(cudaStreamSynchronize((cudaStream_t)g_stream), "stream synchronize");
        for (auto& s : staged) rets[s.first]->set(*s.second);
    }
}

// ---- module constructor ---------------------------------------------------------------------------------------------------
bool initialize() {
    static int state = 0;   // 0 = not tried, 1 = ok, 2 = failed
    if (state) return state == 1;
    dopt::initialize();
    if (dopt_b200_init() != 0) {
        g_init_error = dopt_b200_last_error();
        state = 2;
        return false;   // like the reference: failures are swallowed and the previous defaults stay (package.d:59-62)
    }
    // dopt.cuda.{basic,math,nnet,random}.initialize(): one registration per op type the library implements
    const char* p = dopt_b200_list_operations();
    while (*p) {
        std::string name(p);
        registerCUDAKernel(name, [](Operation op) { return std::make_shared<B200Kernel>(op); });
        p += name.size() + 1;
    }
    defaultEvaluator() = [](const std::vector<Operation>& ops, const std::map<Operation, Buffer>& args) {
        B200Plan plan(ops, DOPT_B200_PLAN_FUSE);   // one-off evaluation: no point capturing a graph
        return plan.execute(args);
    };
    defaultCompiler() = [](const std::vector<Operation>& ops) -> PlanPtr { return std::make_shared<B200Plan>(ops, g_plan_flags); };
    defaultVarAllocator() = [](size_t n) -> Buffer { return CUDABuffer::create(n); };
    state = 1;
    return true;
}

void initDataParallel(int rank, int worldSize, const void* uniqueId128) {
    abiEnforce(dopt_b200_comm_init(rank, worldSize, uniqueId128));
    setDataParallelWorld(worldSize);
}

}  // namespace cuda
}  // namespace dopt
