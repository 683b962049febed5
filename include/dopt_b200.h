/*
 * dopt_b200.h -- C ABI of libdopt_b200.so: B200 (sm_100a) kernels for dopt's CUDA training hot path.
 *
 * This is the drop-in boundary.  Every entry point replaces one piece of dopt's `cuda` backend
 * (citations are relative to the dopt source tree):
 *
 *   dopt_b200_kernel_create   <- CUDAKernelCtr  `CUDAKernel delegate(Operation op)`   cuda/source/dopt/cuda/package.d:25
 *                                (called from the CUDAPlan ctor, package.d:284-288)
 *   dopt_b200_kernel_execute  <- CUDAKernel.execute(inputs, output)                   cuda/source/dopt/cuda/package.d:68-79
 *                                (called every step from CUDAPlan.executeImpl, package.d:412)
 *   dopt_b200_list_operations <- listCUDAOperations()                                 cuda/source/dopt/cuda/package.d:503-506
 *   dopt_b200_plan_*          <- CUDAPlan ctor / executeImpl                          cuda/source/dopt/cuda/package.d:267-312,343-424
 *                                (installed through defaultCompiler, core/source/dopt/core/package.d:46-49)
 *   dopt_b200_comm_*          <- (new) data-parallel gradient exchange; the reference is single-device
 *                                (cuda/source/dopt/cuda/package.d:43-45)
 *
 * Conventions
 *   - plain C: pointers, sizes, POD structs.  No C++ or torch types cross this boundary.
 *   - every function returns 0 on success, non-zero on failure; dopt_b200_last_error() gives the message
 *     (the D glue does `enforce(rc == 0, fromStringz(dopt_b200_last_error()))`, mirroring cudnnCheck,
 *     cuda/source/dopt/cuda/nnet/cudnn7.d:42-48).  Nothing throws across the boundary.
 *   - device pointers are raw `void*` (== CUdeviceptr).  The caller owns all tensor memory
 *     (dopt's CUDABuffer, package.d:124-251); the library owns only handles, packed-weight scratch and workspaces.
 *   - `stream` is a cudaStream_t / CUstream passed as void*; NULL = the legacy default stream, which orders
 *     with the reference's synchronous cuMemcpy* calls.
 *   - the library uses the CUDA context that is current on the calling thread (dopt creates its own with
 *     cuCtxCreate, package.d:43-45); it never switches device.
 *   - there is NO CPU fallback: on a machine without a usable sm_100 device every *_create fails.
 */
#ifndef DOPT_B200_H
#define DOPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DOPT_B200_MAX_RANK   8
#define DOPT_B200_MAX_INPUTS 8

/* dopt DataType, core/source/dopt/core/types.d:5-9 */
typedef enum { DOPT_B200_FLOAT32 = 0, DOPT_B200_INT32 = 1 } dopt_b200_dtype;

/* numerics of the dense contractions (convolution*, matmul).  FP32 = SIMT fp32 accumulate in fp32 (tight parity with
 * the reference's cuDNN/cuBLAS fp32 calls); BF16 = tcgen05 tensor cores, bf16 operands, fp32 accumulate in TMEM.
 *
 * NOTE -- deviation from the reference: DEFAULT resolves to BF16 (dopt_b200_set_default_math changes that).  The reference's
 * cuDNN / cuBLAS calls compute strict fp32.  With BF16 a convolution differs from the fp32 result by 2-3e-3 of the tensor's
 * max magnitude (stated bound 2e-2), the filter gradient is accumulated with fp32 atomics and is therefore not
 * bit-reproducible from run to run, and loss curves of the BASELINE configs agree with the fp32 oracle within 3e-2 relative
 * (measured 2.5e-4 on WRN-28-10: DESIGN.md section 5, profiles/r02_parity.md).  Select FP32 for reference-accurate,
 * reproducible results. */
typedef enum { DOPT_B200_MATH_DEFAULT = 0, DOPT_B200_MATH_FP32 = 1, DOPT_B200_MATH_BF16 = 2 } dopt_b200_math;

/* TensorType, core/source/dopt/core/types.d:16-57 */
typedef struct {
    int32_t dtype;                       /* dopt_b200_dtype */
    int32_t rank;
    int64_t shape[DOPT_B200_MAX_RANK];
} dopt_b200_tensor;

/* One graph node as the kernel constructor sees it: op type, operand types (op.deps[i].outputType), result type
 * (op.outputType) and the attributes the reference kernels read from op.attributes. */
typedef struct {
    const char*      op_type;                              /* e.g. "convolution", "add", "batchNormTrain" */
    int32_t          n_inputs;
    dopt_b200_tensor inputs[DOPT_B200_MAX_INPUTS];
    dopt_b200_tensor output;
    /* attributes (only those meaningful for op_type are read) */
    int64_t padding[2];                                    /* convolution*: attributes["padding"]  */
    int64_t stride[2];                                     /* convolution*: attributes["stride"]   */
    int64_t pool_dims[2];                                  /* maxpool*: attributes["dims"]         */
    int64_t start[DOPT_B200_MAX_RANK];                     /* slice: attributes["start"]           */
    int64_t stop[DOPT_B200_MAX_RANK];                      /* slice: attributes["stop"]            */
    int64_t before[DOPT_B200_MAX_RANK];                    /* pad: attributes["before"]            */
    int64_t after[DOPT_B200_MAX_RANK];                     /* pad: attributes["after"]             */
    int64_t repetitions[DOPT_B200_MAX_RANK];               /* repeat: attributes["repetitions"]    */
    int64_t order[DOPT_B200_MAX_RANK];                     /* transpose: attributes["order"]       */
    int64_t axes[DOPT_B200_MAX_RANK];                      /* sum / maxElement: attributes["axes"] */
    int32_t n_axes;
    int64_t axis;                                          /* argmin: attributes["axis"]           */
    double  momentum;                                      /* batchNormTrain: attributes["momentum"] */
    uint64_t seed;                                         /* uniform: 0 = unpredictable seed like the reference */
    int32_t math;                                          /* dopt_b200_math */
    int32_t reserved[7];
} dopt_b200_op;

typedef struct dopt_b200_kernel_s* dopt_b200_kernel_t;
typedef struct dopt_b200_plan_s*   dopt_b200_plan_t;

/* ---- library ---------------------------------------------------------------------------------------------------- */
int         dopt_b200_init(void);                          /* checks for an sm_100 device on the current context   */
const char* dopt_b200_last_error(void);                    /* thread-local message of the last failing call        */
const char* dopt_b200_version(void);
int         dopt_b200_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);
void        dopt_b200_set_default_math(int math);          /* what DOPT_B200_MATH_DEFAULT resolves to (default BF16) */
/* number of kernel launches issued by this library since process start (bench.py's "gpu_launches") */
uint64_t    dopt_b200_launch_count(void);
/* diagnostics: CUDA-event timing of the tcgen05 implicit-GEMM kernel (tc_kernel) alone.  enable=1 starts a session;
 * a later call returns the device time (us) and number of tc_kernel launches since then (bench.py's roofline leg) */
int         dopt_b200_tc_profile(int enable, double* us, int64_t* launches);

/* ---- per-op kernels: registerCUDAKernel / CUDAKernel.execute ----------------------------------------------------- */
/* NUL-separated, double-NUL-terminated list of op types with a registered kernel */
const char* dopt_b200_list_operations(void);
int         dopt_b200_has_operation(const char* op_type);
int         dopt_b200_kernel_create(const dopt_b200_op* op, dopt_b200_kernel_t* out);
/* inputs[i] <-> op.deps[i]; output is op's buffer (volume * sizeof(elem) bytes) */
int         dopt_b200_kernel_execute(dopt_b200_kernel_t k, const void* const* inputs, int n_inputs, void* output,
                                     void* stream);
int         dopt_b200_kernel_destroy(dopt_b200_kernel_t k);

/* ---- fused optimiser updates (dopt.online) ------------------------------------------------------------------------
 * One launch updates a whole list of parameter tensors.  Hyper-parameters are rank-0 DEVICE tensors, exactly as in the
 * reference graphs (online/source/dopt/online/sgd.d:28-29, adam.d:32-34), so they can be changed between steps with
 * `.value.set` and the launch can live in a CUDA graph.  Arithmetic order follows the reference graph op by op
 * (no FMA contraction) so results are bit-identical to the unfused pointwise kernels.
 *   sgd:     m' = m*mu + lr*g ; w' = w - m'                 (sgd.d:57-64;  nesterov: sgd.d:46-55)
 *   adam:    b1' = b1*beta1, b2' = b2*beta2, eta = alpha*sqrt(1-b2')/(1-b1')
 *            m' = beta1*m + (1-beta1)*g ; v' = beta2*v + ((1-beta2)*g)*g ; w' = w - eta*(m'/(sqrt(v')+eps))   (adam.d:46-66)
 *   amsgrad: as adam, plus vhat' = max(vhat, v_old); the update still uses sqrt(v') (amsgrad.d:63-70, survey F11)
 * `grad_scale` (host float) multiplies g first (1/world_size for data-parallel mean); 1.0f leaves g untouched.
 * A tensor with g == NULL is skipped for the gradient term (treated as zeros), matching grads that are zero variables. */
typedef struct {
    float*       w;      /* parameter, updated in place */
    const float* g;      /* gradient                    */
    float*       s0;     /* sgd: momentum; adam/amsgrad: mean */
    float*       s1;     /* adam/amsgrad: var           */
    float*       s2;     /* amsgrad: varhat             */
    int64_t      n;      /* elements                    */
} dopt_b200_param;

int dopt_b200_sgd_update(const dopt_b200_param* params, int n_params, const float* lr, const float* momentum,
                         int nesterov, float grad_scale, void* stream);
int dopt_b200_adam_update(const dopt_b200_param* params, int n_params, const float* alpha, const float* beta1,
                          const float* beta2, const float* eps, float* b1, float* b2, int amsgrad, float grad_scale,
                          void* stream);

/* ---- on-device input pipeline (the caller's side of the hot path) ---------------------------------------------------
 * Replaces the host loops that prepare every batch in the reference: the loaders' byte -> float normalisation
 * `x / 128.0f - 1.0f` and one-hot labels (nnet/source/dopt/nnet/data/cifar.d:50-55) and ImageTransformer.getBatch
 * (nnet/source/dopt/nnet/data/imagetransformer.d:45-138: reflect-pad by jitter, crop at a random offset, random mirror).
 * The batch is uploaded as bytes and ONE launch writes the NCHW fp32 tensor the plan reads.  Bit-exact with the host loops.
 * All pointers are DEVICE pointers.  per_image == NULL means no jitter and no flip (plain normalisation / copy). */
typedef struct {
    int32_t x_off, y_off;     /* crop offset in the reflect-padded image: uniform(0, 2*jitter), imagetransformer.d:101-102 */
    int32_t flip_x, flip_y;   /* mirror rows / columns after cropping, imagetransformer.d:118-137                          */
} dopt_b200_jitter;

int dopt_b200_image_transform_u8(const uint8_t* src, float* dst, int64_t n, int c, int h, int w, int jitter_x, int jitter_y,
                                 const dopt_b200_jitter* per_image, void* stream);
/* the same for an already normalised float batch (what ImageTransformer itself receives); src != dst */
int dopt_b200_image_transform_f32(const float* src, float* dst, int64_t n, int c, int h, int w, int jitter_x, int jitter_y,
                                  const dopt_b200_jitter* per_image, void* stream);
/* dst[n, classes] = one-hot of labels[n] (cifar.d:52-55) */
int dopt_b200_one_hot_u8(const uint8_t* labels, float* dst, int64_t n, int classes, void* stream);
/* fills out[n] on the device with the reference's distributions (Philox-4x32-10 keyed by seed, counter = (image, call));
 * flip_x / flip_y enable the respective mirror (ImageTransformer's flipX / flipY constructor flags) */
int dopt_b200_jitter_sample(dopt_b200_jitter* out, int64_t n, int jitter_x, int jitter_y, int flip_x, int flip_y,
                            uint64_t seed, uint64_t call, void* stream);

/* ---- whole-plan compiler: CUDAPlan ---------------------------------------------------------------------------------
 * The host serialises the topologically sorted graph once (node ids are the order of add_node calls) and the library
 * lowers it: reshape/slice aliasing, BN pack/unpack removal, scalar-broadcast folding, pointwise fusion, buffer
 * planning, CUDA-graph capture.  Semantics of execute() are those of CUDAPlan.executeImpl (package.d:343-424):
 * variables are read from `args` (device or host pointers), outputs are copied to `rets`. */
int dopt_b200_plan_create(dopt_b200_plan_t* out);
/* returns the node id (>= 0) or a negative error.  `deps` are node ids.  For "variable" / "constant" nodes n_deps = 0;
 * `value` (may be NULL) is a HOST pointer to the constant's bytes / the variable's device buffer binding is given at
 * execute time. */
int dopt_b200_plan_add_node(dopt_b200_plan_t p, const dopt_b200_op* op, const int32_t* deps, int n_deps,
                            const void* const_value);
int dopt_b200_plan_set_outputs(dopt_b200_plan_t p, const int32_t* node_ids, int n_outputs);
/* flags */
#define DOPT_B200_PLAN_FUSE        1   /* graph-level lowering and fusion (off = node-by-node like the reference)   */
#define DOPT_B200_PLAN_CUDA_GRAPH  2   /* capture the step in a CUDA graph                                          */
/* bf16 interior (needs PLAN_FUSE; only has an effect where tensor-core convolutions, i.e. MATH_BF16, are in use):
 * activations that only tensor-core convolutions, batch norms and residual adds read are kept as NHWC bf16 and never
 * materialised as NCHW fp32 -- the convolution epilogue writes bf16, batchNormTrain / batchNormGrad / add work on that
 * layout.  Plan inputs, outputs and everything any other op reads stay fp32.  NUMERICS: the stored activations are rounded
 * to bf16 (8 mantissa bits) where the reference stores fp32; arithmetic stays fp32.  Stated tolerance: DESIGN.md section 5. */
#define DOPT_B200_PLAN_BF16_INTERIOR 4
int dopt_b200_plan_finalize(dopt_b200_plan_t p, int flags);
/* var_ids[i] is a "variable" node id, var_ptrs[i] its buffer; var_on_host[i] != 0 means a host pointer that is
 * uploaded first (CUDAPlan does the same for CPUBuffer args, package.d:373-381).  rets[i] is a DEVICE pointer that
 * receives output i (package.d:419-422). */
int dopt_b200_plan_execute(dopt_b200_plan_t p, const int32_t* var_ids, const void* const* var_ptrs,
                           const int32_t* var_on_host, int n_vars, void* const* rets, int n_rets, void* stream);
/* statistics: number of kernel launches per execute, bytes of plan-owned device memory, nodes after lowering */
int dopt_b200_plan_stats(dopt_b200_plan_t p, int64_t* launches, int64_t* device_bytes, int64_t* lowered_nodes);
/* per-op-type accumulated device time in microseconds since the last reset (CUDAPlan.profiler, package.d:265);
 * only filled when profiling is on (it serialises the stream with events).  `buf` receives "opType=usec\n" lines. */
int dopt_b200_plan_profile(dopt_b200_plan_t p, int enable, char* buf, size_t buf_len);
/* measurement aid (no reference counterpart): re-issue ONLY the launches the profiler books under the comma-separated
 * `op_types` ("batchNormTrain", "convolution,convolutionFeaturesGrad,convolutionFiltersGrad", "fusedRegion" ...) with the
 * operands of the last execution, `reps` times back to back on `stream`, timed by ONE pair of CUDA events -- the device
 * time of a kernel class without a per-launch event bracket around every kernel.  It re-runs accumulating kernels out of
 * order (batch-norm statistic sinks, running statistics, in-place parameter updates), so the plan's state is garbage
 * afterwards: call it after the last execution that matters (bench.py does, for its roofline figures). */
int dopt_b200_plan_replay_class(dopt_b200_plan_t p, const char* op_types, int reps, double* usec_per_rep,
                                int64_t* launches_per_rep, void* stream);
int dopt_b200_plan_destroy(dopt_b200_plan_t p);

/* ---- data-parallel gradient exchange (new; one process per GPU) -----------------------------------------------------
 * NCCL over NVLink 5 / NVSwitch.  unique_id is the 128-byte ncclUniqueId produced by rank 0. */
int dopt_b200_comm_unique_id(void* id128);
int dopt_b200_comm_init(int rank, int world_size, const void* id128);
int dopt_b200_comm_world_size(void);
int dopt_b200_comm_rank(void);
/* in-place sum over ranks of a float buffer, then multiply by `scale` (1/world for a mean) */
int dopt_b200_allreduce(float* buf, int64_t n, float scale, void* stream);
/* polls the communicator's asynchronous error state (ncclCommGetAsyncError): collectives are enqueued without waiting, so a
 * failed peer or link only shows up here.  Non-zero = the communicator was aborted; dopt_b200_last_error() says why.
 * (The same poll runs before every all-reduce the library enqueues.)  Call it after synchronising a step. */
int dopt_b200_comm_check(void);
int dopt_b200_comm_destroy(void);
/* Optional: peer-mapped ("symmetric") memory over NVSwitch multicast.  The caller maps one buffer of `bytes` per rank into
 * every process (CUDA VMM: cuMemCreate + shareable handles, one multicast object bound to all of them) and hands over
 *   local_base      this rank's own buffer (unicast mapping),
 *   multicast_base  the multicast mapping that aliases the buffers of ALL ranks (multimem.* instructions),
 *   signal_pads[r]  a zero-initialised, peer-mapped array of 32-bit flags owned by rank r (>= 4 KB each, same layout everywhere).
 * Plans created afterwards carve their gradient-bucket arenas out of the buffer -- every rank must create the same plans in the
 * same order, so that a bucket has the same offset everywhere -- and reduce them with the library's own kernel instead of
 * ncclAllReduce: each rank reduces 1/world of the bucket inside the switch (multimem.ld_reduce), scales it and broadcasts it
 * (multimem.st), between two flag barriers over the signal pads.  Buckets that do not fit fall back to NCCL.  The memory stays
 * owned by the caller and must outlive the plans.  bench.py / tools/dp_check.py obtain it from torch.distributed's
 * symmetric-memory allocator; a D host would use the driver API directly. */
int dopt_b200_comm_set_symmetric(void* local_base, void* multicast_base, size_t bytes, void* const* signal_pads, int n_pads,
                                 size_t signal_pad_bytes);

#ifdef __cplusplus
}
#endif
#endif /* DOPT_B200_H */
