/**
    dopt.b200 -- plugs libdopt_b200.so (hand-written sm_100a kernels) into dopt's CUDA backend.

    Import this module after dopt.cuda.  Its module constructor
      1. replaces every kernel constructor dopt.cuda registered (cuDNN / cuBLAS / NVRTC) by one that forwards to the C ABI,
      2. registers GPU kernels for the ops dopt.cuda leaves to its CPU fallback (sum, maxElement, argmin),
      3. installs B200Plan as defaultCompiler, so `compile()` -- and therefore dopt.online.sgd/adam/amsgrad -- hands the whole
         graph to the library (lowering, fusion, CUDA-graph capture).

    This file is shipped as source: there is no D compiler in the build image.  The C++ mirror in dopt_b200/host/dopt/cuda.cpp
    is a line-for-line transliteration of it and is what the tests exercise.
*/
module dopt.b200;

import std.exception : enforce;
import std.string : fromStringz, toStringz;

import dopt.core;
import dopt.cuda;

extern(C) nothrow @nogc
{
    enum DOPT_B200_MAX_RANK = 8;
    enum DOPT_B200_MAX_INPUTS = 8;

    struct dopt_b200_tensor { int dtype; int rank; long[DOPT_B200_MAX_RANK] shape; }

    struct dopt_b200_op
    {
        const(char)* op_type;
        int n_inputs;
        dopt_b200_tensor[DOPT_B200_MAX_INPUTS] inputs;
        dopt_b200_tensor output;
        long[2] padding, stride, pool_dims;
        long[DOPT_B200_MAX_RANK] start, stop, before, after, repetitions, order, axes;
        int n_axes;
        long axis;
        double momentum;
        ulong seed;
        int math;
        int[7] reserved;
    }

    alias dopt_b200_kernel_t = void*;
    alias dopt_b200_plan_t = void*;

    int dopt_b200_init();
    const(char)* dopt_b200_last_error();
    const(char)* dopt_b200_list_operations();
    int dopt_b200_kernel_create(const(dopt_b200_op)* op, dopt_b200_kernel_t* k);
    int dopt_b200_kernel_execute(dopt_b200_kernel_t k, const(void*)* inputs, int n_inputs, void* output, void* stream);
    int dopt_b200_kernel_destroy(dopt_b200_kernel_t k);
    int dopt_b200_plan_create(dopt_b200_plan_t* p);
    int dopt_b200_plan_add_node(dopt_b200_plan_t p, const(dopt_b200_op)* op, const(int)* deps, int n_deps, const(void)* constValue);
    int dopt_b200_plan_set_outputs(dopt_b200_plan_t p, const(int)* ids, int n);
    int dopt_b200_plan_finalize(dopt_b200_plan_t p, int flags);
    int dopt_b200_plan_execute(dopt_b200_plan_t p, const(int)* varIds, const(void*)* varPtrs, const(int)* varOnHost, int nVars,
                               void** rets, int nRets, void* stream);
    int dopt_b200_plan_destroy(dopt_b200_plan_t p);
    int dopt_b200_plan_stats(dopt_b200_plan_t p, long* launches, long* deviceBytes, long* loweredNodes);
    int dopt_b200_plan_profile(dopt_b200_plan_t p, int enable, char* buf, size_t bufLen);   // CUDAPlan.profiler (package.d:265)
    int dopt_b200_plan_replay_class(dopt_b200_plan_t p, const(char)* opTypes, int reps, double* usecPerRep, long* launchesPerRep,
                                    void* stream);                                         // measurement aid, no counterpart

    const(char)* dopt_b200_version();
    int dopt_b200_device_info(int* smCount, int* ccMajor, int* ccMinor, size_t* totalMem);
    int dopt_b200_has_operation(const(char)* opType);
    // what DOPT_B200_MATH_DEFAULT resolves to: 1 = strict fp32 kernels (the reference's arithmetic), 2 = bf16 tensor cores (default)
    void dopt_b200_set_default_math(int math);

    // the dopt.online update rules as direct entry points (optional: the plan fuses the update graphs of sgd.d / adam.d /
    // amsgrad.d on its own; these serve a host that keeps its own training loop)
    struct dopt_b200_param { float* w; const(float)* g; float* s0; float* s1; float* s2; long n; }
    int dopt_b200_sgd_update(const(dopt_b200_param)* params, int nParams, const(float)* lr, const(float)* momentum, int nesterov,
                             float gradScale, void* stream);
    int dopt_b200_adam_update(const(dopt_b200_param)* params, int nParams, const(float)* alpha, const(float)* beta1,
                              const(float)* beta2, const(float)* eps, float* b1, float* b2, int amsgrad, float gradScale,
                              void* stream);

    // on-device input pipeline (replaces the host loops of nnet/data/cifar.d:50-55 and nnet/data/imagetransformer.d:45-138)
    struct dopt_b200_jitter { int x_off, y_off, flip_x, flip_y; }
    int dopt_b200_image_transform_u8(const(ubyte)* src, float* dst, long n, int c, int h, int w, int jitterX, int jitterY,
                                     const(dopt_b200_jitter)* perImage, void* stream);
    int dopt_b200_image_transform_f32(const(float)* src, float* dst, long n, int c, int h, int w, int jitterX, int jitterY,
                                      const(dopt_b200_jitter)* perImage, void* stream);
    int dopt_b200_one_hot_u8(const(ubyte)* labels, float* dst, long n, int classes, void* stream);
    int dopt_b200_jitter_sample(dopt_b200_jitter* dst, long n, int jitterX, int jitterY, int flipX, int flipY, ulong seed,
                                ulong call, void* stream);

    // data-parallel gradient exchange (new: the reference is single-device, cuda/source/dopt/cuda/package.d:43-45); see
    // INTEGRATION.md section 3 for the `allreduce` operation and the exchange() helper that use these
    int dopt_b200_comm_unique_id(void* id128);
    int dopt_b200_comm_init(int rank, int worldSize, const(void)* id128);
    int dopt_b200_comm_world_size();
    int dopt_b200_comm_rank();
    int dopt_b200_allreduce(float* buf, long n, float scale, void* stream);
    int dopt_b200_comm_check();
    int dopt_b200_comm_destroy();
    // optional: peer-mapped memory + NVSwitch multicast mapping for the gradient buckets (the library's own all-reduce kernel)
    int dopt_b200_comm_set_symmetric(void* localBase, void* multicastBase, size_t bytes, const(void*)* signalPads, int nPads,
                                     size_t signalPadBytes);
}

enum DOPT_B200_PLAN_FUSE = 1;
enum DOPT_B200_PLAN_CUDA_GRAPH = 2;
enum DOPT_B200_PLAN_BF16_INTERIOR = 4;   // activations between tensor-core convolutions stay NHWC bf16 (include/dopt_b200.h)

enum DOPT_B200_MATH_FP32 = 1;
enum DOPT_B200_MATH_BF16 = 2;

/// Arithmetic of convolution* and large matmul: `setMath(DOPT_B200_MATH_FP32)` gives the reference's strict fp32 results (1e-4 of
/// cuDNN / cuBLAS), the default is bf16 operands with fp32 accumulation on the tensor cores (include/dopt_b200.h, "NOTE").
void setMath(int math)
{
    dopt_b200_set_default_math(math);
}

private void check(int rc)
{
    // same convention as cudnnCheck (cuda/source/dopt/cuda/nnet/cudnn7.d:42-48)
    enforce(rc == 0, dopt_b200_last_error().fromStringz.idup);
}

private void fill(ref dopt_b200_tensor t, TensorType type)
{
    enforce(type.rank <= DOPT_B200_MAX_RANK, "tensor rank exceeds DOPT_B200_MAX_RANK");
    t.dtype = type.elementType == DataType.float32 ? 0 : 1;
    t.rank = cast(int)type.rank;
    foreach(i, s; type.shape) t.shape[i] = cast(long)s;
}

private void copySizes(ref long[2] dst, Operation op, string name)
{
    if(auto p = name in op.attributes) if(auto v = p.peek!(size_t[])) foreach(i, x; *v) dst[i] = cast(long)x;
}

private void copySizes(ref long[DOPT_B200_MAX_RANK] dst, Operation op, string name)
{
    if(auto p = name in op.attributes) if(auto v = p.peek!(size_t[])) foreach(i, x; *v) dst[i] = cast(long)x;
}

/// Everything the reference kernels read from `op` at construction, as the POD the C ABI takes.
package dopt_b200_op describe(Operation op)
{
    dopt_b200_op d;
    d.op_type = op.opType.toStringz;
    d.n_inputs = cast(int)op.deps.length;
    foreach(i, dep; op.deps) fill(d.inputs[i], dep.outputType);
    fill(d.output, op.outputType);
    d.stride = [1, 1];
    copySizes(d.padding, op, "padding");
    copySizes(d.stride, op, "stride");
    copySizes(d.pool_dims, op, "dims");
    copySizes(d.start, op, "start");
    copySizes(d.stop, op, "stop");
    copySizes(d.before, op, "before");
    copySizes(d.after, op, "after");
    copySizes(d.repetitions, op, "repetitions");
    copySizes(d.order, op, "order");
    copySizes(d.axes, op, "axes");
    if(auto p = "axes" in op.attributes) if(auto v = p.peek!(size_t[])) d.n_axes = cast(int)v.length;
    if(auto p = "axis" in op.attributes) if(auto v = p.peek!size_t) d.axis = cast(long)*v;
    if(auto p = "momentum" in op.attributes) if(auto v = p.peek!double) d.momentum = *v;
    return d;
}

/// The one CUDAKernel class of the glue (interface: cuda/source/dopt/cuda/package.d:68-79).
class B200Kernel : CUDAKernel
{
    this(Operation op)
    {
        auto d = describe(op);
        check(dopt_b200_kernel_create(&d, &mHandle));
    }

    ~this()
    {
        dopt_b200_kernel_destroy(mHandle);
    }

    void execute(const(CUDABuffer)[] inputs, CUDABuffer output)
    {
        const(void)*[DOPT_B200_MAX_INPUTS] ptrs;
        foreach(i, b; inputs) ptrs[i] = cast(const(void)*)b.ptr;
        // stream null = the legacy default stream, which orders with dopt's synchronous cuMemcpy* calls
        check(dopt_b200_kernel_execute(mHandle, ptrs.ptr, cast(int)inputs.length, cast(void*)output.ptr, null));
    }

    private dopt_b200_kernel_t mHandle;
}

/// Whole-graph plan (replaces CUDAPlan, cuda/source/dopt/cuda/package.d:261-424).
class B200Plan : Plan
{
    this(Operation[] outputs, int flags = DOPT_B200_PLAN_FUSE | DOPT_B200_PLAN_CUDA_GRAPH)
    {
        super(outputs);
        check(dopt_b200_plan_create(&mPlan));

        foreach(o; topologicalSort(outputs))
        {
            auto d = describe(o);
            int[] deps;
            foreach(dep; o.deps) deps ~= mIds[dep];
            const(void)* cval = null;
            ubyte[] tmp;
            if(o.opType == "constant")
            {
                tmp = o.value.get!ubyte;
                cval = tmp.ptr;
            }
            int id = dopt_b200_plan_add_node(mPlan, &d, deps.ptr, cast(int)deps.length, cval);
            enforce(id >= 0, dopt_b200_last_error().fromStringz.idup);
            mIds[o] = id;
            if(o.opType == "variable") mVariables ~= o;
        }

        int[] outs;
        foreach(o; outputs) outs ~= mIds[o];
        check(dopt_b200_plan_set_outputs(mPlan, outs.ptr, cast(int)outs.length));
        check(dopt_b200_plan_finalize(mPlan, flags));
    }

    ~this()
    {
        dopt_b200_plan_destroy(mPlan);
    }

    protected override void executeImpl(DeviceBuffer[Operation] args, DeviceBuffer[] rets)
    {
        import dopt.cpu : CPUBuffer;

        int[] ids, onHost;
        const(void)*[] ptrs;

        // same rules as CUDAPlan.executeImpl (package.d:349-392): args must be variables; variables not in args are read
        // from their own buffers; host buffers are uploaded by the plan
        foreach(v; mVariables)
        {
            DeviceBuffer buf = (v in args) ? args[v] : v.value;
            ids ~= mIds[v];
            if(auto cu = cast(CUDABuffer)buf) { ptrs ~= cast(const(void)*)cu.ptr; onHost ~= 0; }
            else if(auto cpu = cast(CPUBuffer)buf) { ptrs ~= cast(const(void)*)cpu.raw.ptr; onHost ~= 1; }
            else enforce(0, "unknown DeviceBuffer type");
        }
        foreach(o; args.keys) enforce(o.opType == "variable",
            "All assignments in args must be for Operations with an opType of 'variable'");

        void*[] retPtrs;
        foreach(r; rets) retPtrs ~= cast(void*)(cast(CUDABuffer)r).ptr;   // dopt.online passes the variables' own CUDABuffers

        check(dopt_b200_plan_execute(mPlan, ids.ptr, ptrs.ptr, onHost.ptr, cast(int)ids.length, retPtrs.ptr,
                                     cast(int)retPtrs.length, null));
    }

    private
    {
        dopt_b200_plan_t mPlan;
        int[Operation] mIds;
        Operation[] mVariables;
    }
}

/**
    Device-side counterpart of `ImageTransformer` (nnet/source/dopt/nnet/data/imagetransformer.d): same constructor
    arguments, but the batch is uploaded as the dataset's raw bytes and one kernel launch performs normalisation
    (`x / 128.0f - 1.0f`, nnet/data/cifar.d:50), reflect-padding, the random crop and the random mirrors, writing the NCHW
    float tensor straight into a `CUDABuffer`.  Labels become one-hot rows on the device as well.  The buffers are handed to
    the updater as they are (`updater([features: t.features, labels: t.labels])`): `CUDAPlan.executeImpl` uses a `CUDABuffer`
    argument in place instead of uploading it (cuda/source/dopt/cuda/package.d:373-381).
*/
class DeviceImageTransformer
{
    this(size_t batchSize, size_t[] imageShape, size_t numLabels, size_t jitterX, size_t jitterY, bool flipX, bool flipY,
         ulong seed = 0)
    {
        import std.random : unpredictableSeed;

        mBatch = batchSize; mShape = imageShape.dup; mClasses = numLabels;
        mJitterX = jitterX; mJitterY = jitterY; mFlipX = flipX; mFlipY = flipY;
        mSeed = seed == 0 ? unpredictableSeed : seed;       // the reference draws from an unseeded std.random
        auto vol = imageShape[0] * imageShape[1] * imageShape[2];
        mRaw = CUDABuffer.create(batchSize * vol);
        mRawLabels = CUDABuffer.create(batchSize);
        mDraws = CUDABuffer.create(batchSize * dopt_b200_jitter.sizeof);
        features = CUDABuffer.create(batchSize * vol * float.sizeof);
        labels = CUDABuffer.create(batchSize * numLabels * float.sizeof);
    }

    /// pixels: batchSize * C*H*W dataset bytes; labelBytes: batchSize class indices
    void nextBatch(const(ubyte)[] pixels, const(ubyte)[] labelBytes)
    {
        mRaw.set(pixels);
        mRawLabels.set(labelBytes);
        check(dopt_b200_jitter_sample(cast(dopt_b200_jitter*)mDraws.ptr, cast(long)mBatch, cast(int)mJitterX,
                                      cast(int)mJitterY, mFlipX, mFlipY, mSeed, mCall++, null));
        check(dopt_b200_image_transform_u8(cast(const(ubyte)*)mRaw.ptr, cast(float*)features.ptr, cast(long)mBatch,
                                           cast(int)mShape[0], cast(int)mShape[1], cast(int)mShape[2], cast(int)mJitterX,
                                           cast(int)mJitterY, cast(const(dopt_b200_jitter)*)mDraws.ptr, null));
        check(dopt_b200_one_hot_u8(cast(const(ubyte)*)mRawLabels.ptr, cast(float*)labels.ptr, cast(long)mBatch,
                                   cast(int)mClasses, null));
    }

    CUDABuffer features, labels;

    private
    {
        size_t mBatch, mClasses, mJitterX, mJitterY;
        size_t[] mShape;
        bool mFlipX, mFlipY;
        ulong mSeed, mCall;
        CUDABuffer mRaw, mRawLabels, mDraws;
    }
}

shared static this()
{
    // dopt.cuda's own constructor has already run (this module imports it); it swallows its failures
    // (cuda/source/dopt/cuda/package.d:59-62), so check for a usable device ourselves.
    if(dopt_b200_init() != 0)
    {
        return;
    }

    import std.functional : toDelegate;

    CUDAKernel ctor(Operation op) { return new B200Kernel(op); }

    // NUL-separated, double-NUL-terminated list
    auto p = dopt_b200_list_operations();
    while(*p)
    {
        auto name = p.fromStringz.idup;
        deregisterCUDAKernel(name);                       // registerCUDAKernel throws if the name is taken (package.d:481)
        registerCUDAKernel(name, toDelegate(&ctor));
        p += name.length + 1;
    }

    defaultCompiler = (Operation[] ops) { return cast(Plan)new B200Plan(ops); };
    defaultEvaluator = (Operation[] ops, DeviceBuffer[Operation] args)
    {
        return (new B200Plan(ops, DOPT_B200_PLAN_FUSE)).execute(args);
    };
}
