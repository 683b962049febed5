// dopt/cuda.hpp -- C++ mirror of dopt.cuda's public interface (cuda/source/dopt/cuda/package.d), backed by the C ABI of
// libdopt_b200.so.  This file is the C++ spelling of the D glue module shown in INTEGRATION.md: a CUDAKernel subclass that
// forwards to dopt_b200_kernel_execute, registered for every op type, and a Plan subclass (B200Plan) that serialises the
// graph into dopt_b200_plan_* and is installed as defaultCompiler.
#pragma once
#include "core.hpp"

namespace dopt {
namespace cuda {

// CUDABuffer, package.d:124-251
class CUDABuffer : public DeviceBuffer {
public:
    static std::shared_ptr<CUDABuffer> create(size_t numBytes);   // cuMemAlloc + zero fill (package.d:135-156)
    ~CUDABuffer() override;
    size_t numBytes() const override { return mNumBytes; }
    void set(const void* buf, size_t bytes) override;              // H2D
    void set(const DeviceBuffer& other) override;                  // D2D when other is a CUDABuffer, else H2D
    void get(void* buf, size_t bytes) const override;              // D2H
    void* ptr() const { return mPtr; }
private:
    CUDABuffer() {}
    size_t mNumBytes = 0;
    void* mPtr = nullptr;
};

// interface CUDAKernel, package.d:68-79
class CUDAKernel {
public:
    virtual ~CUDAKernel() {}
    virtual void execute(const std::vector<const CUDABuffer*>& inputs, CUDABuffer& output) = 0;
};
using CUDAKernelCtr = std::function<std::shared_ptr<CUDAKernel>(Operation op)>;

void registerCUDAKernel(const std::string& opName, CUDAKernelCtr ctr);   // throws if the name is taken (package.d:479-485)
void deregisterCUDAKernel(const std::string& opType);                    // package.d:493-496
std::vector<std::string> listCUDAOperations();                           // package.d:503-506

// The reference's node-by-node executor (package.d:261-424) running B200 kernels through the per-op C ABI.
class CUDAPlan : public Plan {
public:
    explicit CUDAPlan(std::vector<Operation> outputs);
    std::map<std::string, long> profiler;   // microseconds per op type (host stopwatch, like the reference)
protected:
    void executeImpl(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets) override;
private:
    std::vector<Operation> mOps;
    std::map<const OperationNode*, std::shared_ptr<CUDAKernel>> mKernels;
    std::map<const OperationNode*, std::shared_ptr<CUDABuffer>> mResults;
};

// Whole-graph plan: serialises the toposorted graph into dopt_b200_plan_* once; lowering, fusion, buffer planning and
// CUDA-graph capture happen inside the library.
class B200Plan : public Plan {
public:
    B200Plan(std::vector<Operation> outputs, int flags);
    ~B200Plan() override;
    void stats(int64_t* launches, int64_t* deviceBytes, int64_t* loweredNodes) const;
    std::string profile(bool enable);
    // device time (us) and launches of ONE repetition of the kernels booked under `opTypes` (dopt_b200_plan_replay_class)
    double replayClass(const std::string& opTypes, int reps, int64_t* launches);
    // raw execution for benchmarks: device or host pointers, no DeviceBuffer objects (what executeImpl does internally)
    void executeRaw(const std::vector<Operation>& argOps, const std::vector<const void*>& argPtrs,
                    const std::vector<int>& argOnHost, const std::vector<void*>& rets);
protected:
    void executeImpl(const std::map<Operation, Buffer>& args, std::vector<Buffer>& rets) override;
private:
    void* mPlan = nullptr;
    std::vector<Operation> mVariables;                       // every variable node the plan contains
    std::map<const OperationNode*, int> mIds;
    std::vector<Buffer> mKeepAlive;
};

// plan flags used by defaultCompiler (DOPT_B200_PLAN_FUSE | DOPT_B200_PLAN_CUDA_GRAPH by default)
void setPlanFlags(int flags);
int planFlags();
void setMath(int math);   // DOPT_B200_MATH_*

// == dopt.cuda's `shared static this()` (package.d:38-63): registers the kernels, then overrides defaultEvaluator,
// defaultCompiler and defaultVarAllocator.  Returns false (and leaves the defaults alone) when no usable device exists,
// like the reference swallows its init failure (package.d:59-62); lastInitError() tells why.
bool initialize();
const std::string& lastInitError();
void* currentStream();
void setStream(void* stream);

// data-parallel: create the NCCL communicator (one process per GPU) and make grad() wrap gradients in `allreduce`
void initDataParallel(int rank, int worldSize, const void* uniqueId128);

}  // namespace cuda
}  // namespace dopt
