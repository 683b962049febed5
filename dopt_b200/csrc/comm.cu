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

// ---- all-reduce over NVSwitch multicast (symmetric memory supplied by the caller, dopt_b200_comm_set_symmetric) -------------
// Two-shot: rank r owns slice r of the buffer.  multimem.ld_reduce fetches the slice from every rank and adds inside the
// switch; the sum is scaled to the mean and multimem.st writes it back to every rank.  Traffic per GPU and direction: one
// buffer length, whatever the world size.  Two flag barriers bracket it: the first so that every rank's bucket is complete,
// the second so that nobody reads its bucket before every slice has landed.  A barrier is per CTA -- CTA b of rank r puts a
// flag into slot [b][r] of every peer's pad (CAS 0 -> 1) and takes the flags of its own slots [b][*] (CAS 1 -> 0) -- so the
// launches need the same grid on every rank and nothing else: no host synchronisation, no NCCL, capturable in a CUDA graph.
namespace {
constexpr int kSymmMaxWorld = 16;
constexpr size_t kSymmPadOffset = 1024;   // bytes: the head of a pad is left to its owner (torch's own collectives use it)
struct Symm {
    char* local = nullptr;
    char* mc = nullptr;
    size_t bytes = 0, used = 0;
    uint32_t* pads[kSymmMaxWorld] = {};
    size_t pad_bytes = 0;
    unsigned* err = nullptr;    // device word: a barrier gave up waiting
    int ctas = 16;
};
Symm g_symm;

struct SymmArgs {
    float4* mc;                  // multicast address of the bucket
    int64_t n4;                  // float4 elements
    uint32_t* pads[kSymmMaxWorld];
    int rank, world;
    float scale;
    unsigned* err;
};

__device__ __forceinline__ uint32_t cas_sys(uint32_t* addr, uint32_t cmp, uint32_t val, int sem) {
    uint32_t old;
    if (sem == 0) asm volatile("atom.global.relaxed.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    else if (sem == 1) asm volatile("atom.global.release.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    else asm volatile("atom.global.acquire.sys.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "l"(addr), "r"(cmp), "r"(val) : "memory");
    return old;
}
// thread t < world handles peer t.  A peer that never shows up (a crashed rank) must not hang the GPU: after ~30 s the barrier
// gives up, raises the error word (dopt_b200_comm_check reports it) and the kernel finishes with garbage.
template <bool ACQ_REL>
__device__ __forceinline__ void symm_barrier(const SymmArgs& a) {
    if ((int)threadIdx.x < a.world) {
        const int peer = (int)threadIdx.x;
        const long long t0 = clock64();
        uint32_t* put = a.pads[peer] + (size_t)blockIdx.x * a.world + a.rank;
        uint32_t* take = a.pads[a.rank] + (size_t)blockIdx.x * a.world + peer;
        bool ok = true;
        while (ok && cas_sys(put, 0u, 1u, ACQ_REL ? 1 : 0) != 0u)
            if (clock64() - t0 > 60000000000ll) ok = false;
        while (ok && cas_sys(take, 1u, 0u, ACQ_REL ? 2 : 0) != 1u)
            if (clock64() - t0 > 60000000000ll) ok = false;
        if (!ok) *a.err = 1u;
    }
}
__global__ void __launch_bounds__(512) symm_allreduce_kernel(const __grid_constant__ SymmArgs a) {
    symm_barrier<false>(a);   // (the buckets were written by kernels that completed before this one started on each rank)
    __syncthreads();
    const int64_t per = a.n4 / a.world;
    const int64_t lo = per * a.rank, hi = a.rank == a.world - 1 ? a.n4 : lo + per;
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += U * stride) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < hi)
                asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                             : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w) : "l"(a.mc + i + u * stride) : "memory");
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride < hi) {
                v[u].x *= a.scale; v[u].y *= a.scale; v[u].z *= a.scale; v[u].w *= a.scale;
                asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a.mc + i + u * stride), "f"(v[u].x),
                             "f"(v[u].y), "f"(v[u].z), "f"(v[u].w) : "memory");
            }
    }
    __syncthreads();
    symm_barrier<true>(a);
}
bool symm_owns(const void* p, size_t bytes) {
    return g_symm.local && (const char*)p >= g_symm.local && (const char*)p + bytes <= g_symm.local + g_symm.bytes;
}
}  // namespace

// carve `bytes` (a multiple of 256) out of the symmetric buffer; nullptr when there is none or it is full
void* comm_symm_alloc(size_t bytes) {
    static const bool off = getenv("DOPT_B200_NO_NVLS") != nullptr;
    if (off || !g_symm.local || g_world <= 1) return nullptr;
    bytes = (bytes + 255) / 256 * 256;
    if (g_symm.used + bytes > g_symm.bytes) return nullptr;
    void* p = g_symm.local + g_symm.used;
    g_symm.used += bytes;
    return p;
}
// give the most recent allocations back (stack discipline: a plan releases its arenas in reverse order when it is destroyed)
void comm_symm_release(void* p, size_t bytes) {
    bytes = (bytes + 255) / 256 * 256;
    if (g_symm.local && (char*)p + bytes == g_symm.local + g_symm.used) g_symm.used -= bytes;
}
bool comm_symm_active() { return g_symm.local != nullptr && g_world > 1; }

int comm_world() { return g_world; }
int comm_reserved_sms() { return g_world > 1 ? g_channels : 0; }

// mean over ranks, in place, one NCCL call (ncclAvg); used by the plan's gradient buckets
void allreduce_mean(float* buf, int64_t n, cudaStream_t s) {
    if (n <= 0 || g_world <= 1) return;
    if (symm_owns(buf, (size_t)n * 4) && n % 4 == 0 && ((uintptr_t)buf & 15) == 0) {
        // the bucket lives in the symmetric buffer: the library's own all-reduce over NVSwitch multicast
        SymmArgs a{};
        a.mc = (float4*)(g_symm.mc + ((char*)buf - g_symm.local));
        a.n4 = n / 4;
        for (int r = 0; r < g_world; ++r) a.pads[r] = (uint32_t*)((char*)g_symm.pads[r] + kSymmPadOffset);
        a.rank = g_rank;
        a.world = g_world;
        a.scale = 1.0f / (float)g_world;
        a.err = g_symm.err;
        const int64_t work = (a.n4 / g_world + 4 * 512 - 1) / (4 * 512);
        const int ctas = (int)std::max<int64_t>(1, std::min<int64_t>(g_symm.ctas, work));
        // (the same grid on every rank: it only depends on the bucket size and the world size)
        symm_allreduce_kernel<<<ctas, 512, 0, s>>>(a);
        DB_LAUNCH_CHECK();
        return;
    }
    DB_REQUIRE(g_comm != nullptr, "allreduce: communicator not initialised (dopt_b200_comm_init)");
    poll_async_error("before a gradient-bucket all-reduce");
    nccl_check(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat, ncclAvg, g_comm, s), "ncclAllReduce");
    count_launch();
}
void comm_check() {
    if (g_world > 1) poll_async_error("dopt_b200_comm_check");
    if (g_symm.err) {
        unsigned e = 0;
        DB_CUDA(cudaMemcpy(&e, g_symm.err, sizeof(e), cudaMemcpyDeviceToHost));
        if (e) throw Error("all-reduce over symmetric memory: a flag barrier timed out (a peer did not reach the collective)");
    }
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
int dopt_b200_comm_set_symmetric(void* local_base, void* multicast_base, size_t bytes, void* const* signal_pads, int n_pads,
                                 size_t signal_pad_bytes) {
    try {
        db::require_device();
        if (!local_base) {   // detach
            db::g_symm.local = db::g_symm.mc = nullptr;
            db::g_symm.bytes = db::g_symm.used = 0;
            return 0;
        }
        DB_REQUIRE(db::g_world > 1 && n_pads == db::g_world, "comm_set_symmetric: call dopt_b200_comm_init first; one signal pad per rank");
        DB_REQUIRE(db::g_world <= db::kSymmMaxWorld, "comm_set_symmetric: world size too large");
        DB_REQUIRE(multicast_base && signal_pads && bytes >= 256, "comm_set_symmetric: multicast mapping and signal pads are required");
        DB_REQUIRE(((uintptr_t)local_base & 255) == 0 && ((uintptr_t)multicast_base & 255) == 0, "comm_set_symmetric: 256-byte alignment");
        int ctas = 16;
        if (const char* e = getenv("DOPT_B200_NVLS_CTAS")) ctas = std::max(1, std::min(64, atoi(e)));
        DB_REQUIRE(signal_pad_bytes >= db::kSymmPadOffset + (size_t)ctas * db::g_world * 4, "comm_set_symmetric: signal pads too small");
        db::g_symm.local = (char*)local_base;
        db::g_symm.mc = (char*)multicast_base;
        db::g_symm.bytes = bytes / 256 * 256;
        db::g_symm.used = 0;
        db::g_symm.pad_bytes = signal_pad_bytes;
        db::g_symm.ctas = ctas;
        for (int r = 0; r < n_pads; ++r) {
            DB_REQUIRE(signal_pads[r], "comm_set_symmetric: null signal pad");
            db::g_symm.pads[r] = (uint32_t*)signal_pads[r];
        }
        if (!db::g_symm.err) {
            DB_CUDA(cudaMalloc((void**)&db::g_symm.err, 16));
            DB_CUDA(cudaMemset(db::g_symm.err, 0, 16));
        }
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return 1;
    }
    return 0;
}
int dopt_b200_comm_destroy(void) {
    try {
        db::g_symm.local = db::g_symm.mc = nullptr;
        db::g_symm.bytes = db::g_symm.used = 0;
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
