// common.cuh -- shared plumbing for libdopt_b200.so (errors, registry, launch helpers).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <atomic>
#include <stdexcept>
#include <string>
#include <vector>
#include <cstring>
#include <cstdio>

#include "../../include/dopt_b200.h"

namespace db {

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

void set_last_error(const std::string& s);

#define DB_REQUIRE(cond, msg)                                                                         \
    do {                                                                                              \
        if (!(cond)) throw db::Error(std::string(msg) + " [" #cond "] at " __FILE__ ":" + std::to_string(__LINE__)); \
    } while (0)

#define DB_CUDA(expr)                                                                                 \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess)                                                                        \
            throw db::Error(std::string("CUDA error: ") + cudaGetErrorString(_e) + " in " #expr " at " __FILE__ ":" + \
                            std::to_string(__LINE__));                                                \
    } while (0)

extern std::atomic<uint64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

// Post-launch check (cheap: only reads the sticky launch error, no sync).
#define DB_LAUNCH_CHECK()                                                                             \
    do {                                                                                              \
        db::count_launch();                                                                           \
        cudaError_t _e = cudaGetLastError();                                                          \
        if (_e != cudaSuccess)                                                                        \
            throw db::Error(std::string("kernel launch failed: ") + cudaGetErrorString(_e) + " at " __FILE__ ":" + \
                            std::to_string(__LINE__));                                                \
    } while (0)

int sm_count();                 // SMs of the current device (148 on B200)
void require_device();          // throws unless an sm_100 device is current

inline int64_t volume(const dopt_b200_tensor& t) {
    int64_t v = 1;
    for (int i = 0; i < t.rank; ++i) v *= t.shape[i];
    return v;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ---- programmatic dependent launch --------------------------------------------------------------------------------------
// A kernel launched through launch_pdl() may be scheduled while its predecessor in the stream is still running: its CTAs are
// placed as SMs free up and run their prologue, then block in pdl_wait() until the predecessor grid has completed and its
// writes are visible.  Every global-memory access of such a kernel comes after pdl_wait(); pdl_trigger() at the top lets the
// NEXT kernel do the same.  Inside the captured CUDA graph this becomes a programmatic edge; it removes the launch latency
// between the ~250 dependent kernels of a training step.  DOPT_B200_PDL=0 turns it off.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Grid size for a grid-stride streaming kernel: enough CTAs to fill every SM `waves` times over, never more than needed.
inline int stream_grid(int64_t work_items, int threads, int ctas_per_sm = 8) {
    int64_t need = ceil_div(work_items, threads);
    int64_t cap = (int64_t)sm_count() * ctas_per_sm;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// A kernel object == dopt's CUDAKernel (cuda/source/dopt/cuda/package.d:68-79).
// one filter to pack: KCRS fp32 -> bf16, mode 0 = forward layout [K][RS][Cp], mode 1 = feature-gradient layout [C][RS][Kp]
// mode 2 = both layouts from one read of w (`out`/Kp/Cp describe the forward layout, `out2`/Kp2 the feature-gradient one);
// mode -1 = row merged into another one (kept so that row indices stay stable), nothing to do
struct FilterPack {
    const float* w;
    void* out;
    int K, C, RS, Kp, Cp, mode;
    int tiles_x, tile0;      // filled by the launcher: tiles along the first grid dimension, first tile of this row
    void* out2;
    int Kp2;
};
size_t filter_pack_bytes(const FilterPack& f);
// packs rows[0..n) (device copy of the table in dev_rows); total_tiles and smem_bytes come from filter_pack_layout
void filter_pack_layout(FilterPack* rows, int n, int* total_tiles, size_t* smem_bytes);
void filter_pack_launch(const FilterPack* dev_rows, int n, int total_tiles, size_t smem_bytes, cudaStream_t s);

// one filter gradient to finish: the tensor-core kernel's [tap][C][K] accumulation scratch -> dopt's KCRS (flipped) gradient
struct WgradFinish {
    const float* scratch;
    float* dw;
    int K, C, RS;
    int tiles_x, tile0;      // filled by the launcher: 32-wide K tiles, first tile of this row
};
void wgrad_finish_layout(WgradFinish* rows, int n, int* total_tiles, size_t* smem_bytes);
void wgrad_finish_launch(const WgradFinish* dev_rows, int n, int total_tiles, size_t smem_bytes, cudaStream_t s);

// one row of a batched full reduction (msum.cu): *out = sum_i a[i] * b[i]  (b == nullptr: sum_i a[i])
struct MsumRow {
    const float* a;
    const float* b;
    float* out;
    int64_t n;
    int64_t chunk0;   // filled by msum_layout
};
int64_t msum_layout(MsumRow* rows, int n);   // returns the number of partials (floats of scratch)
void msum_launch(const MsumRow* dev_rows, int n, int64_t chunks, float* partial, cudaStream_t s);

struct Absorb {
    bool relu = false;          // apply relu to the result
    float* redirect = nullptr;  // write the fp32 result here instead of the kernel's own output (the relu node's buffer)
    bool skip_fp32 = false;     // nobody reads the fp32 result: do not write it
    void* staged = nullptr;     // also write the result as [N][HW][Cp] bf16 here
    const float* addend = nullptr;   // batchNormGrad: the result is dx + addend (the residual gradient sum that follows it)
};

struct Kernel {
    virtual ~Kernel() {}
    virtual void run(const void* const* in, int n_in, void* out, cudaStream_t s) = 0;
    // Tensor-core convolutions read their activation operands as NHWC bf16.  The plan can stage an activation once and hand
    // the staged copy to every convolution op that reads it (forward + filter gradient share x; feature + filter gradient
    // share dy).  staged_bytes(i) > 0 means input i can be supplied pre-staged; set_staged_input(i, p) supplies it.
    virtual size_t staged_bytes(int /*input*/) const { return 0; }
    virtual void set_staged_input(int /*input*/, const void* /*nhwc_bf16*/) {}
    // Producer-side fusion (plan.cu, pass "absorb"): a kernel that can apply relu to its result and / or also emit the
    // NHWC bf16 copy the tensor-core convolutions read (see struct Absorb).  Only the leading V elements of a packed
    // result are affected.
    virtual bool can_absorb() const { return false; }
    virtual void set_absorbed(const struct Absorb&) {}
    // Filter staging: a tensor-core convolution reads its filter operand in a packed bf16 layout.  When the filter is a
    // plan variable (a parameter), the plan packs ALL filters of the step in one launch at its start (FilterPack rows) and
    // hands every kernel its packed copy instead of letting each op pack on its own.
    virtual bool filter_pack(int /*input*/, struct FilterPack* /*desc*/) const { return false; }
    virtual void set_packed_filter(const void* /*packed*/) {}
    // batchNormGrad whose incoming gradient is gated by the relu that followed `forward` (a batchNormTrain kernel): the gate
    // is recomputed from x and forward->aux_ptr() (its per-channel coefficients), so neither reluGrad nor the stored relu
    // output is needed
    virtual void set_gate_source(const Kernel* /*forward*/) {}
    // batchNormTrain: also write the running statistics (packed tail of the result) straight to these buffers -- the plan
    // passes the caller's return buffers (dopt.online feeds them back as the new `mean` / `var`) and skips the copies
    virtual bool set_stat_outputs(float* /*new_mean*/, float* /*new_var*/) { return false; }
    // bf16-interior plans (plan.cu, pass "residency"; kernels in flat.cu).  A tensor-core convolution can write its result
    // as NHWC bf16 (set_staged_output) instead of NCHW fp32.  A "flat" batchNormTrain / batchNormGrad / add reads its tensor
    // operands from the NHWC bf16 copies given with set_staged_input (batchNormGrad: index 3 = the addend of an absorbed
    // residual add) and writes only Absorb::staged; the fp32 pointers passed to run() are then ignored.
    virtual bool can_stage_output() const { return false; }
    virtual void set_staged_output(void* /*nhwc_bf16*/) {}
    virtual bool can_flat() const { return false; }
    virtual void set_flat(bool /*on*/) {}
    // batch-norm statistics accumulated by the producer of x: a flat batchNormTrain hands out its statistics workspace
    // (stats_workspace; mode 1 = sums pivoted by pixel 0, the residual add; mode 2 = plain sums, the convolution epilogue) and
    // skips its own statistics kernel; the producer (can_produce_stats) accumulates into it while it writes x
    // convolutionFiltersGrad on the tensor cores: leave the result in a private [tap][C][K] scratch and let the plan turn the
    // scratches of many filter gradients into KCRS with ONE multi-tensor launch (28 latency-sized launches per WRN step
    // otherwise).  deferred_finish fills everything of the row but `dw`.
    virtual bool deferred_finish(struct WgradFinish* /*row*/) { return false; }
    // ... and the plan may supply that scratch itself (a slice of one arena holding the scratches of every deferred filter
    // gradient, zeroed by ONE memset per step instead of one per convolution); the kernel then neither owns nor clears it
    virtual void set_finish_scratch(float* /*zeroed_by_caller*/) {}
    // true when run() touches no workspace shared with other ops (only its operands, its result and private scratch): the
    // plan may then issue it on a side stream, concurrently with the ops that follow it in the order
    virtual bool side_stream_safe() const { return false; }
    virtual void* stats_workspace(int /*mode*/) { return nullptr; }
    virtual int can_produce_stats() const { return 0; }   // 0 = no, else the mode it produces
    virtual void set_stats_workspace(void* /*bn_workspace*/, int /*channels*/) {}
    // epilogue companions of a tensor-core convolution writing NHWC bf16 (tc.cuh: conv_tc_set_companion).  mode 2, forward:
    // the residual sum the convolution feeds is written by its epilogue (result + src); mode 3, unit-stride feature gradient:
    // the epilogue also accumulates the backward statistics of the batch norm that reads the result (src = that batch norm's
    // x, coef = its forward coefficients) into the workspace given with set_stats_workspace.  A flat batchNormGrad hands out
    // that workspace through stats_workspace(3) and then skips its own statistics kernel.
    virtual bool can_companion(int /*mode*/) const { return false; }
    virtual void set_companion(int /*mode*/, const void* /*src*/, const float* /*coef*/) {}
    virtual const void* aux_ptr() const { return nullptr; }
};
// NCHW fp32 -> [N][HW][Cp] bf16 (Cp = C rounded up to 8), the staging the tensor-core convolutions use
void stage_nchw_to_nhwc_bf16(const float* in, void* out, int N, int C, int64_t HW, cudaStream_t s);
size_t staged_nhwc_bytes(int N, int C, int64_t HW);
using Factory = Kernel* (*)(const dopt_b200_op&);
void register_kernel(const char* op_type, Factory f);   // == registerCUDAKernel (package.d:479-485)
Factory find_kernel(const char* op_type);
int resolve_math(int math);

// registration entry points, one per translation unit (== dopt.cuda.{math,basic,nnet,random}.initialize)
void register_pointwise();
void register_basic();
void register_reduce();
void register_matmul();
void register_nnet();
void register_batchnorm();
void register_conv();
void register_random();
void register_comm();

// small device scratch that lives for the process (workspaces, packed weights); grows on demand.
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    void* get(size_t need);
    ~Scratch();
};

}  // namespace db

// ---------------------------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------------------------
namespace dbk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// streaming (read-once) 128-bit load / store: keep them out of L1
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 r;
    asm volatile("ld.global.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w)
                 : "memory");
}

}  // namespace dbk
