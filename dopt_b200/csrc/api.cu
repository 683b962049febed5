// api.cu -- library state, kernel registry and the per-op part of the C ABI (include/dopt_b200.h).
// Mirrors dopt.cuda's registry: registerCUDAKernel / deregisterCUDAKernel / listCUDAOperations
// (cuda/source/dopt/cuda/package.d:479-506) and the kernel lifecycle CUDAPlan drives (package.d:284-288,412).
#include "common.cuh"
#include <cstdlib>
#include <map>
#include <mutex>

namespace db {

static thread_local std::string t_last_error;
void set_last_error(const std::string& s) { t_last_error = s; }

std::atomic<uint64_t> g_launches{0};

static std::map<std::string, Factory>& registry() {
    static std::map<std::string, Factory> r;
    return r;
}
static std::once_flag g_reg_once;
static std::string g_op_list;
static int g_default_math = DOPT_B200_MATH_BF16;

void register_kernel(const char* op_type, Factory f) {
    // same rule as the reference: a second registration for one op type is an error (package.d:481-482)
    DB_REQUIRE(registry().find(op_type) == registry().end(), std::string("kernel already registered for ") + op_type);
    registry()[op_type] = f;
}

static void ensure_registered() {
    std::call_once(g_reg_once, [] {
        register_pointwise();
        register_basic();
        register_reduce();
        register_matmul();
        register_nnet();
        register_batchnorm();
        register_conv();
        register_random();
        register_comm();
        for (auto& kv : registry()) {
            g_op_list += kv.first;
            g_op_list.push_back('\0');
        }
        g_op_list.push_back('\0');
    });
}

Factory find_kernel(const char* op_type) {
    ensure_registered();
    auto it = registry().find(op_type);
    return it == registry().end() ? nullptr : it->second;
}

int resolve_math(int math) { return math == DOPT_B200_MATH_DEFAULT ? g_default_math : math; }

bool pdl_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("DOPT_B200_PDL");
        v = e ? (atoi(e) != 0) : 1;
    }
    return v != 0;
}

static int g_sm_count = 0;
int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            g_sm_count = n;
        else
            g_sm_count = 148;
    }
    return g_sm_count;
}

void require_device() {
    static bool ok = false;
    if (ok) return;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        throw Error(std::string("no CUDA device available (") + cudaGetErrorString(e) +
                    "); libdopt_b200 has no CPU fallback");
    }
    int major = 0, minor = 0;
    DB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    DB_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    if (major != 10)
        throw Error("libdopt_b200 is built for sm_100a only; current device is sm_" + std::to_string(major) +
                    std::to_string(minor));
    ok = true;
}

void* Scratch::get(size_t need) {
    if (need > bytes) {
        if (ptr) DB_CUDA(cudaFree(ptr));
        ptr = nullptr;
        size_t nb = need + need / 4;
        DB_CUDA(cudaMalloc(&ptr, nb));
        bytes = nb;
    }
    return ptr;
}
Scratch::~Scratch() { /* process teardown: the context may already be gone; leak on purpose */ }

}  // namespace db

struct dopt_b200_kernel_s {
    db::Kernel* impl;
};

#define DB_API_BEGIN try {
#define DB_API_END                                    \
    }                                                 \
    catch (const std::exception& e) {                 \
        db::set_last_error(e.what());                 \
        return 1;                                     \
    }                                                 \
    catch (...) {                                     \
        db::set_last_error("unknown C++ exception");  \
        return 2;                                     \
    }                                                 \
    return 0;

namespace db {
void tc_prof_enable(bool on);
void tc_prof_read(double* us, int64_t* launches);
}  // namespace db

extern "C" {

int dopt_b200_init(void) {
    DB_API_BEGIN
    db::require_device();
    db::find_kernel("add");
    DB_API_END
}

const char* dopt_b200_last_error(void) { return db::t_last_error.c_str(); }
const char* dopt_b200_version(void) { return "dopt_b200 0.1 (sm_100a)"; }

int dopt_b200_device_info(int* sms, int* cc_major, int* cc_minor, size_t* total_mem) {
    DB_API_BEGIN
    int dev = 0;
    DB_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    DB_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sms) *sms = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    DB_API_END
}

void dopt_b200_set_default_math(int math) {
    if (math == DOPT_B200_MATH_FP32 || math == DOPT_B200_MATH_BF16) db::g_default_math = math;
}

uint64_t dopt_b200_launch_count(void) { return db::g_launches.load(); }

int dopt_b200_tc_profile(int enable, double* us, int64_t* launches) {
    try {
        double u = 0;
        int64_t l = 0;
        db::tc_prof_read(&u, &l);
        if (us) *us = u;
        if (launches) *launches = l;
        db::tc_prof_enable(enable != 0);
        return 0;
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return -1;
    }
}

const char* dopt_b200_list_operations(void) {
    try {
        db::find_kernel("add");
    } catch (...) {
        return "\0";
    }
    return db::g_op_list.c_str();
}

int dopt_b200_has_operation(const char* op_type) {
    try {
        return db::find_kernel(op_type) != nullptr;
    } catch (...) {
        return 0;
    }
}

int dopt_b200_kernel_create(const dopt_b200_op* op, dopt_b200_kernel_t* out) {
    DB_API_BEGIN
    DB_REQUIRE(op && out && op->op_type, "null argument");
    db::require_device();
    db::Factory f = db::find_kernel(op->op_type);
    if (!f) throw db::Error(std::string("Could not construct a CUDA kernel for operation of type '") + op->op_type + "'");
    db::Kernel* k = f(*op);
    *out = new dopt_b200_kernel_s{k};
    DB_API_END
}

int dopt_b200_kernel_execute(dopt_b200_kernel_t k, const void* const* inputs, int n_inputs, void* output,
                             void* stream) {
    DB_API_BEGIN
    DB_REQUIRE(k && k->impl, "null kernel");
    k->impl->run(inputs, n_inputs, output, (cudaStream_t)stream);
    DB_API_END
}

int dopt_b200_kernel_destroy(dopt_b200_kernel_t k) {
    DB_API_BEGIN
    if (k) {
        delete k->impl;
        delete k;
    }
    DB_API_END
}

}  // extern "C"
