// comm.cu -- data-parallel gradient exchange over NCCL (NVLink 5 / NVSwitch).
//
// The reference has no multi-device support at all: device ordinal 0 is hard-coded and there is no collective call site
// (cuda/source/dopt/cuda/package.d:43-45).  Training shards by minibatch: one process per GPU, every rank runs the full
// graph on its slice, and parameter gradients are summed across ranks before the optimiser update (mean of the per-rank
// mean losses, nnet/source/dopt/nnet/losses.d:25).  NCCL is loaded with dlopen so that a process that already carries a
// libnccl (PyTorch bundles one) shares it instead of loading a second copy.
#include "common.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <cstdlib>

namespace db {
namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat = 7, ncclSum = 0, ncclAvg = 4 };

struct Nccl {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;
    ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static Nccl g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;
// CTAs (channels) the all-reduce kernels may use = SMs the tensor-core kernels leave free while a bucket is in flight.
// NCCL reads NCCL_MAX_NCHANNELS once per process, when its first communicator is created: a host that creates NCCL
// communicators of its own first (bench.py: torch.distributed) must export the variable before that.
static int g_channels = 0;

static void load_nccl() {
    if (g_nccl.h) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) throw Error(std::string("cannot load libnccl: ") + dlerror());
#define SYM(field, name)                                              \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.h, name);                 \
    if (!g_nccl.field) throw Error(std::string("libnccl lacks ") + name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(AllReduce, "ncclAllReduce")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(CommGetAsyncError, "ncclCommGetAsyncError")
    SYM(CommAbort, "ncclCommAbort")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
}
static void pick_channels() {
    int k = 16;   // measured at 2 GPUs (profiles/r02_summary.md): 82 / 153 / 268 GB/s with 4 / 8 / 16 channels
    if (const char* e = getenv("DOPT_B200_COMM_CHANNELS")) k = atoi(e);
    if (k <= 0) {   // 0: NCCL's own choice, no SM reservation
        g_channels = 0;
        return;
    }
    if (const char* e = getenv("NCCL_MAX_NCHANNELS")) {
        g_channels = std::max(1, std::min(atoi(e), 32));
        return;
    }
    char buf[16];
    snprintf(buf, sizeof(buf), "%d", k);
    setenv("NCCL_MAX_NCHANNELS", buf, 0);
    setenv("NCCL_MIN_NCHANNELS", buf, 0);
    g_channels = k;
}
static void nccl_check(ncclResult_t r, const char* what) {
    if (r != 0) throw Error(std::string("NCCL error in ") + what + ": " + g_nccl.GetErrorString(r));
}

// Collectives are enqueued asynchronously (inside a CUDA graph, even): a failure of a peer or of the fabric only shows up as
// the communicator's asynchronous error state.  It is polled before every enqueue and through dopt_b200_comm_check(); a
// communicator in error is aborted so that no rank keeps waiting in a kernel that can never finish.
static void poll_async_error(const char* where) {
    if (!g_comm) return;
    ncclResult_t async = 0;
    nccl_check(g_nccl.CommGetAsyncError(g_comm, &async), "ncclCommGetAsyncError");
    if (async != 0 && async != 7 /* ncclInProgress */) {
        std::string msg = std::string("NCCL asynchronous error (") + where + "): " + g_nccl.GetErrorString(async);
        g_nccl.CommAbort(g_comm);
        g_comm = nullptr;
        throw Error(msg);
    }
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ p, int64_t n, float s) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        p[i] = __fmul_rn(p[i], s);
}
}  // namespace

int comm_world() { return g_world; }
int comm_reserved_sms() { return g_world > 1 ? g_channels : 0; }

// mean over ranks, in place, one NCCL call (ncclAvg); used by the plan's gradient buckets
void allreduce_mean(float* buf, int64_t n, cudaStream_t s) {
    if (n <= 0 || g_world <= 1) return;
    DB_REQUIRE(g_comm != nullptr, "allreduce: communicator not initialised (dopt_b200_comm_init)");
    poll_async_error("before a gradient-bucket all-reduce");
    nccl_check(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclAvg, g_comm, s), "ncclAllReduce");
    count_launch();
}
void comm_check() {
    if (g_world > 1) poll_async_error("dopt_b200_comm_check");
}

namespace {
// `allreduce` graph node (registered through dopt's registerOperation / registerCUDAKernel like any other op): the mean
// over ranks of its operand.  With one rank it is a copy.
struct AllreduceKernel : Kernel {
    int64_t n;
    AllreduceKernel(const dopt_b200_op& d) {
        DB_REQUIRE(d.n_inputs == 1 && d.output.dtype == DOPT_B200_FLOAT32, "allreduce: one float32 operand");
        n = volume(d.output);
    }
    void run(const void* const* in, int n_in, void* out, cudaStream_t s) override;
};
}  // namespace
void allreduce(float* buf, int64_t n, float scale, cudaStream_t s);
void comm_check();
void AllreduceKernel::run(const void* const* in, int n_in, void* out, cudaStream_t s) {
    DB_REQUIRE(n_in == 1, "allreduce: one input");
    if (n == 0) return;
    if (in[0] != out) {
        DB_CUDA(cudaMemcpyAsync(out, in[0], (size_t)n * 4, cudaMemcpyDeviceToDevice, s));
        count_launch();
    }
    allreduce((float*)out, n, 1.0f / (float)g_world, s);
}
static Kernel* make_allreduce(const dopt_b200_op& d) { return new AllreduceKernel(d); }
void register_comm() { register_kernel("allreduce", make_allreduce); }
int comm_rank() { return g_rank; }

void allreduce(float* buf, int64_t n, float scale, cudaStream_t s) {
    if (n <= 0) return;
    if (g_world > 1) {
        DB_REQUIRE(g_comm != nullptr, "allreduce: communicator not initialised (dopt_b200_comm_init)");
        poll_async_error("before an all-reduce");
        nccl_check(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclSum, g_comm, s), "ncclAllReduce");
        count_launch();
    }
    if (scale != 1.0f) {
        scale_kernel<<<stream_grid(n, 256, 8), 256, 0, s>>>(buf, n, scale);
        DB_LAUNCH_CHECK();
    }
}
}  // namespace db

extern "C" {
int dopt_b200_comm_unique_id(void* id128) {
    try {
        DB_REQUIRE(id128, "null id");
        db::load_nccl();
        db::ncclUniqueId id;
        db::nccl_check(db::g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
        memcpy(id128, &id, sizeof(id));
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
int dopt_b200_comm_init(int rank, int world_size, const void* id128) {
    try {
        DB_REQUIRE(world_size >= 1 && rank >= 0 && rank < world_size, "bad rank / world size");
        db::g_rank = rank;
        db::g_world = world_size;
        if (world_size == 1) return 0;
        DB_REQUIRE(id128, "null id");
        db::require_device();
        db::pick_channels();
        db::load_nccl();
        db::ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        db::nccl_check(db::g_nccl.CommInitRank(&db::g_comm, world_size, id, rank), "ncclCommInitRank");
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
int dopt_b200_comm_world_size(void) { return db::g_world; }
int dopt_b200_comm_rank(void) { return db::g_rank; }
int dopt_b200_allreduce(float* buf, int64_t n, float scale, void* stream) {
    try {
        db::require_device();
        db::allreduce(buf, n, scale, (cudaStream_t)stream);
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
int dopt_b200_comm_check(void) {
    try {
        db::comm_check();
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
int dopt_b200_comm_destroy(void) {
    try {
        if (db::g_comm) {
            db::nccl_check(db::g_nccl.CommDestroy(db::g_comm), "ncclCommDestroy");
            db::g_comm = nullptr;
        }
        db::g_world = 1;
        db::g_rank = 0;
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
}
