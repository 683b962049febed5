// plan.cu -- the whole-graph executor: libdopt_b200's replacement for CUDAPlan
// (cuda/source/dopt/cuda/package.d:261-424).
//
// Reference behaviour that is kept: topologically ordered nodes, one device buffer per materialised node, `reshape`
// aliases its operand (package.d:295-299,403-406), variables come from `args` (host buffers are uploaded first,
// package.d:373-381), plan outputs are copied into `rets` after all nodes ran (package.d:419-422) -- which is how
// dopt.online overwrites parameters, momenta and BN running statistics in place.
// Reference behaviour that is dropped: the host-side associative-array walk + one launch + one cuCtxSynchronize per
// node, GC.collect() per allocation, and the D2H -> CPU -> H2D fallback for ops without a CUDA kernel (package.d:81-119).
//
// Lowering (DOPT_B200_PLAN_FUSE), all pure graph rewrites that keep results bit-identical to node-by-node execution:
//   * contiguous `slice` becomes a view (pointer + offset).  This removes the batch-norm pack/unpack copies: the three
//     slices of batchNormTrain's packed output and of batchNormGrad's packed output (core/source/dopt/core/ops/nnet.d:476-489,
//     core/source/dopt/core/grads/nnet.d:66-83) cost nothing.
//   * slice(pad(x)) that cuts out exactly x is x (the gradient of the y-slice of the packed BN tensor).
//   * scalar broadcasts `reshape(matmul(ones[V,1], reshape(s,[1,1])))` (core/source/dopt/core/ops/package.d:96-103,
//     core/source/dopt/core/ops/basic.d:370-381) are never materialised when their consumers are pointwise binaries: the
//     pointwise kernel reads the rank-0 operand from device memory.
//   * dead nodes (left over after the rewrites) are dropped.
// DOPT_B200_PLAN_CUDA_GRAPH captures the launch sequence once and replays it.
#include "common.cuh"
#include "pointwise.cuh"
#include <map>
#include <set>
#include <unordered_map>

namespace db {
uint64_t tc_stage_generation();
void tc_prof_enable(bool on);
void tc_prof_read(double* us, int64_t* launches);

namespace {

struct Node {
    std::string type;
    dopt_b200_op op{};
    std::vector<int> deps;
    std::vector<uint8_t> const_value;
    int64_t bytes = 0;
    // lowering
    int alias_of = -1;          // view of another node's buffer
    int64_t alias_off = 0;      // byte offset into it
    int bcast_of = -1;          // this node is a broadcast of the rank-0 node `bcast_of`
    bool folded = false;        // broadcast never materialised
    bool needed = false;
    // pointwise-with-scalar rewrite
    int pw_op = -1, pw_mode = dbk::B_TENSOR;
    int eff_in[2] = {-1, -1};
    // runtime
    void* buf = nullptr;        // plan-owned buffer (or nullptr for views / variables)
    void* ptr = nullptr;        // resolved pointer for this execution
    Kernel* kernel = nullptr;
};

static int64_t dtype_size(int) { return 4; }

}  // namespace
}  // namespace db

struct dopt_b200_plan_s {
    std::vector<db::Node> nodes;
    std::vector<int> outputs;
    bool finalized = false;
    int flags = 0;
    std::vector<int> order;                 // materialised nodes in execution order
    int64_t device_bytes = 0;
    int64_t launches_per_exec = 0;
    std::unordered_map<int, void*> var_stage;   // device staging for variables passed as host pointers
    std::unordered_map<int, void*> var_last;    // last device address seen per variable
    std::set<int> var_moves;                    // variables whose address changed between executions
    // CUDA graph
    cudaStream_t cap_stream = nullptr;
    cudaGraphExec_t graph_exec = nullptr;
    uint64_t graph_key = 0;
    int warm_runs = 0;
    // profiler
    bool profiling = false;
    std::map<std::string, double> prof_us;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    ~dopt_b200_plan_s() {
        for (auto& n : nodes) {
            delete n.kernel;
            if (n.buf) cudaFree(n.buf);
        }
        for (auto& kv : var_stage) cudaFree(kv.second);
        if (graph_exec) cudaGraphExecDestroy(graph_exec);
        if (cap_stream) cudaStreamDestroy(cap_stream);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
    }
};

namespace db {
namespace {

using Plan = dopt_b200_plan_s;

static int root_of(Plan& p, int id, int64_t* off = nullptr) {
    int64_t o = 0;
    while (p.nodes[id].alias_of >= 0) {
        o += p.nodes[id].alias_off;
        id = p.nodes[id].alias_of;
    }
    if (off) *off = o;
    return id;
}

static bool is_ones_column(const Node& n) {
    if (n.type != "constant" || n.op.output.rank != 2 || n.op.output.shape[1] != 1) return false;
    if (n.op.output.dtype != DOPT_B200_FLOAT32) return false;
    size_t cnt = n.const_value.size() / 4;
    if (cnt != (size_t)n.op.output.shape[0]) return false;
    const float* f = (const float*)n.const_value.data();
    for (size_t i = 0; i < cnt; ++i)
        if (f[i] != 1.0f) return false;
    return true;
}

static bool slice_is_contiguous(const dopt_b200_op& d, int64_t* elem_off) {
    const auto& in = d.inputs[0];
    int64_t strides[DOPT_B200_MAX_RANK];
    int64_t s = 1;
    for (int i = in.rank - 1; i >= 0; --i) {
        strides[i] = s;
        s *= in.shape[i];
    }
    for (int i = in.rank - 1; i >= 0; --i) {
        if (d.output.shape[i] != in.shape[i]) {
            for (int j = 0; j < i; ++j)
                if (d.output.shape[j] != 1) return false;
            break;
        }
    }
    int64_t off = 0;
    for (int i = 0; i < in.rank; ++i) off += d.start[i] * strides[i];
    *elem_off = off;
    return true;
}

static void lower(Plan& p) {
    const bool fuse = (p.flags & DOPT_B200_PLAN_FUSE) != 0;
    auto& N = p.nodes;
    for (size_t i = 0; i < N.size(); ++i) {
        Node& n = N[i];
        if (n.type == "reshape") {
            n.alias_of = n.deps[0];
            n.alias_off = 0;
            continue;
        }
        if (!fuse) continue;
        if (n.type == "slice") {
            int64_t eo = 0;
            // slice(pad(x)) == x ?
            int src = n.deps[0];
            int64_t src_off = 0;
            int r = root_of(p, src, &src_off);
            if (N[r].type == "pad" && src_off == 0 && volume(N[src].op.output) == volume(N[r].op.output)) {
                const dopt_b200_op& pd = N[r].op;
                bool same = pd.output.rank == n.op.inputs[0].rank;
                for (int k = 0; same && k < pd.output.rank; ++k)
                    same = (n.op.start[k] == pd.before[k]) && (n.op.output.shape[k] == pd.inputs[0].shape[k]) &&
                           (n.op.inputs[0].shape[k] == pd.output.shape[k]);
                if (same) {
                    n.alias_of = N[r].deps[0];
                    n.alias_off = 0;
                    continue;
                }
            }
            if (slice_is_contiguous(n.op, &eo)) {
                n.alias_of = src;
                n.alias_off = eo * dtype_size(n.op.output.dtype);
                continue;
            }
        }
        if (n.type == "matmul" && n.deps.size() == 2) {
            int a = root_of(p, n.deps[0]), b = n.deps[1];
            if (is_ones_column(N[a]) && volume(N[b].op.output) == 1 && n.op.output.dtype == DOPT_B200_FLOAT32)
                n.bcast_of = b;
        }
    }
    if (fuse) {
        // consumers of broadcast nodes: pointwise binaries take the scalar directly
        std::vector<std::vector<int>> users(N.size());
        for (size_t i = 0; i < N.size(); ++i) {
            if (N[i].alias_of >= 0) continue;
            for (int d : N[i].deps) users[root_of(p, d)].push_back((int)i);
        }
        std::set<int> out_roots;
        for (int o : p.outputs) out_roots.insert(root_of(p, o));
        for (size_t i = 0; i < N.size(); ++i) {
            Node& n = N[i];
            if (n.bcast_of < 0) continue;
            bool ok = !out_roots.count((int)i) && !users[i].empty();
            for (int u : users[i]) {
                const Node& c = N[u];
                int op = pointwise_op_id(c.type.c_str());
                if (op < 0 || pointwise_is_unary(op) || c.op.output.dtype != DOPT_B200_FLOAT32) { ok = false; break; }
                int r0 = root_of(p, c.deps[0]), r1 = root_of(p, c.deps[1]);
                bool b0 = (r0 == (int)i) || (N[r0].bcast_of >= 0), b1 = (r1 == (int)i) || (N[r1].bcast_of >= 0);
                if (b0 && b1) { ok = false; break; }   // scalar (op) scalar broadcast: keep it simple, materialise
            }
            n.folded = ok;
        }
        for (size_t i = 0; i < N.size(); ++i) {
            Node& c = N[i];
            if (c.alias_of >= 0 || c.deps.size() != 2) continue;
            int op = pointwise_op_id(c.type.c_str());
            if (op < 0 || pointwise_is_unary(op)) continue;
            int r0 = root_of(p, c.deps[0]), r1 = root_of(p, c.deps[1]);
            c.pw_op = op;
            c.eff_in[0] = c.deps[0];
            c.eff_in[1] = c.deps[1];
            if (N[r1].folded) {
                c.pw_mode = dbk::B_SCALAR_B;
                c.eff_in[1] = N[r1].bcast_of;
            } else if (N[r0].folded) {
                c.pw_mode = dbk::B_SCALAR_A;
                c.eff_in[0] = N[r0].bcast_of;
            }
        }
    }
    // liveness from the outputs
    std::vector<int> stack(p.outputs.begin(), p.outputs.end());
    while (!stack.empty()) {
        int id = stack.back();
        stack.pop_back();
        if (N[id].needed) continue;
        N[id].needed = true;
        if (N[id].alias_of >= 0) {
            stack.push_back(N[id].alias_of);
            continue;
        }
        if (N[id].pw_op >= 0) {
            stack.push_back(N[id].eff_in[0]);
            stack.push_back(N[id].eff_in[1]);
            continue;
        }
        for (int d : N[id].deps) stack.push_back(d);
    }
}

static void build(Plan& p) {
    auto& N = p.nodes;
    lower(p);
    for (size_t i = 0; i < N.size(); ++i) {
        Node& n = N[i];
        if (!n.needed || n.alias_of >= 0) continue;
        if (n.type == "variable") continue;
        DB_CUDA(cudaMalloc(&n.buf, (size_t)std::max<int64_t>(n.bytes, 16)));
        p.device_bytes += n.bytes;
        if (n.type == "constant") {
            DB_REQUIRE((int64_t)n.const_value.size() == n.bytes, "constant node without a value");
            DB_CUDA(cudaMemcpy(n.buf, n.const_value.data(), (size_t)n.bytes, cudaMemcpyHostToDevice));
            continue;
        }
        // buffers are zeroed once at creation like CUDABuffer.create (package.d:152); batchNormGrad relies on it for the
        // unused tail of its over-allocated result (survey F4)
        DB_CUDA(cudaMemset(n.buf, 0, (size_t)std::max<int64_t>(n.bytes, 16)));
        if (n.pw_op >= 0 && n.pw_mode != dbk::B_TENSOR) {
            p.order.push_back((int)i);   // handled by pointwise_launch with a scalar operand
            continue;
        }
        Factory f = find_kernel(n.type.c_str());
        if (!f) throw Error("Could not construct a CUDA kernel for operation of type '" + n.type + "'");
        n.op.op_type = n.type.c_str();
        n.kernel = f(n.op);
        p.order.push_back((int)i);
    }
}

static void run_nodes(Plan& p, cudaStream_t s) {
    auto& N = p.nodes;
    for (int id : p.order) {
        Node& n = N[id];
        if (p.profiling) DB_CUDA(cudaEventRecord(p.ev0, s));
        if (n.kernel) {
            const void* in[DOPT_B200_MAX_INPUTS];
            for (size_t k = 0; k < n.deps.size(); ++k) in[k] = N[n.deps[k]].ptr;
            n.kernel->run(in, (int)n.deps.size(), n.ptr, s);
        } else {
            pointwise_launch(n.pw_op, n.op.output.dtype, n.pw_mode, N[n.eff_in[0]].ptr, N[n.eff_in[1]].ptr, n.ptr,
                             volume(n.op.output), s);
        }
        if (p.profiling) {
            DB_CUDA(cudaEventRecord(p.ev1, s));
            DB_CUDA(cudaEventSynchronize(p.ev1));
            float ms = 0;
            DB_CUDA(cudaEventElapsedTime(&ms, p.ev0, p.ev1));
            p.prof_us[n.type] += ms * 1000.0;
        }
    }
}

static void execute(Plan& p, const int32_t* var_ids, const void* const* var_ptrs, const int32_t* on_host, int n_vars,
                    void* const* rets, int n_rets, cudaStream_t s) {
    DB_REQUIRE(p.finalized, "plan not finalized");
    DB_REQUIRE(n_rets == (int)p.outputs.size(), "wrong number of return buffers");
    auto& N = p.nodes;
    // bind variables
    for (auto& n : N)
        if (n.type == "variable") n.ptr = nullptr;
    uint64_t key = 1469598103934665603ull;
    auto mix = [&](uint64_t v) { key = (key ^ v) * 1099511628211ull; };
    for (int i = 0; i < n_vars; ++i) {
        int id = var_ids[i];
        DB_REQUIRE(id >= 0 && id < (int)N.size() && N[id].type == "variable",
                   "All assignments in args must be for Operations with an opType of 'variable'");   // package.d:349-353
        if (!N[id].needed) continue;
        if (on_host && on_host[i]) {
            void*& st = p.var_stage[id];
            if (!st) DB_CUDA(cudaMalloc(&st, (size_t)std::max<int64_t>(N[id].bytes, 16)));
            DB_CUDA(cudaMemcpyAsync(st, var_ptrs[i], (size_t)N[id].bytes, cudaMemcpyHostToDevice, s));
            N[id].ptr = st;
        } else {
            // A device argument whose address changes between executions (a fresh input batch each step) is copied into a
            // plan-owned buffer from then on, so the captured CUDA graph keeps seeing one address.  Parameters never
            // move (dopt.online writes new values INTO their buffers, package.d:419-422), so they are read in place.
            void* ptr = const_cast<void*>(var_ptrs[i]);
            auto last = p.var_last.find(id);
            if (last != p.var_last.end() && last->second != ptr && (p.flags & DOPT_B200_PLAN_CUDA_GRAPH))
                p.var_moves.insert(id);
            p.var_last[id] = ptr;
            bool is_ret = false;
            for (int r = 0; r < n_rets && !is_ret; ++r) is_ret = (rets[r] == ptr);
            if (p.var_moves.count(id) && !is_ret) {
                void*& st = p.var_stage[id];
                if (!st) DB_CUDA(cudaMalloc(&st, (size_t)std::max<int64_t>(N[id].bytes, 16)));
                DB_CUDA(cudaMemcpyAsync(st, ptr, (size_t)N[id].bytes, cudaMemcpyDeviceToDevice, s));
                ptr = st;
            }
            N[id].ptr = ptr;
        }
        mix((uint64_t)(uintptr_t)N[id].ptr);
    }
    for (int i = 0; i < n_rets; ++i) mix((uint64_t)(uintptr_t)rets[i]);
    mix(tc_stage_generation());
    // resolve pointers
    for (size_t i = 0; i < N.size(); ++i) {
        Node& n = N[i];
        if (!n.needed) continue;
        if (n.type == "variable") {
            DB_REQUIRE(n.ptr != nullptr, "plan_execute: a variable the plan reads was not bound");
        } else if (n.alias_of < 0) {
            n.ptr = n.buf;
        }
    }
    for (size_t i = 0; i < N.size(); ++i) {
        Node& n = N[i];
        if (!n.needed || n.alias_of < 0) continue;
        int64_t off = 0;
        int r = root_of(p, (int)i, &off);
        n.ptr = (char*)N[r].ptr + off;
    }
    auto body = [&](cudaStream_t st) {
        run_nodes(p, st);
        for (int i = 0; i < n_rets; ++i) {
            const Node& o = N[p.outputs[i]];
            if (o.bytes > 0 && rets[i] != o.ptr) {
                DB_CUDA(cudaMemcpyAsync(rets[i], o.ptr, (size_t)o.bytes, cudaMemcpyDeviceToDevice, st));
                count_launch();
            }
        }
    };
    const bool want_graph = (p.flags & DOPT_B200_PLAN_CUDA_GRAPH) && !p.profiling;
    if (!want_graph) {
        uint64_t l0 = g_launches.load();
        body(s);
        p.launches_per_exec = (int64_t)(g_launches.load() - l0);
        return;
    }
    if (p.graph_exec && p.graph_key == key) {
        DB_CUDA(cudaGraphLaunch(p.graph_exec, s));
        count_launch((int)p.launches_per_exec);
        return;
    }
    if (p.warm_runs < 1) {
        // first execution runs eagerly: it sizes every workspace so that nothing allocates during capture
        uint64_t l0 = g_launches.load();
        body(s);
        p.launches_per_exec = (int64_t)(g_launches.load() - l0);
        ++p.warm_runs;
        return;
    }
    if (p.graph_exec) {
        cudaGraphExecDestroy(p.graph_exec);
        p.graph_exec = nullptr;
    }
    if (!p.cap_stream) DB_CUDA(cudaStreamCreateWithFlags(&p.cap_stream, cudaStreamNonBlocking));
    DB_CUDA(cudaStreamSynchronize(s));
    cudaGraph_t graph = nullptr;
    DB_CUDA(cudaStreamBeginCapture(p.cap_stream, cudaStreamCaptureModeThreadLocal));
    uint64_t l0 = g_launches.load();
    try {
        body(p.cap_stream);
    } catch (...) {
        cudaStreamEndCapture(p.cap_stream, &graph);
        if (graph) cudaGraphDestroy(graph);
        throw;
    }
    p.launches_per_exec = (int64_t)(g_launches.load() - l0);
    DB_CUDA(cudaStreamEndCapture(p.cap_stream, &graph));
    DB_CUDA(cudaGraphInstantiate(&p.graph_exec, graph, 0));
    DB_CUDA(cudaGraphDestroy(graph));
    p.graph_key = key;
    DB_CUDA(cudaGraphLaunch(p.graph_exec, s));
}

}  // namespace
}  // namespace db

#define PLAN_TRY try {
#define PLAN_CATCH                                  \
    }                                               \
    catch (const std::exception& e) {               \
        db::set_last_error(e.what());               \
        return 1;                                   \
    }                                               \
    return 0;

extern "C" {

int dopt_b200_plan_create(dopt_b200_plan_t* out) {
    PLAN_TRY
    DB_REQUIRE(out, "null argument");
    *out = new dopt_b200_plan_s;
    PLAN_CATCH
}

int dopt_b200_plan_add_node(dopt_b200_plan_t p, const dopt_b200_op* op, const int32_t* deps, int n_deps,
                            const void* const_value) {
    try {
        DB_REQUIRE(p && op && op->op_type, "null argument");
        DB_REQUIRE(!p->finalized, "plan already finalized");
        DB_REQUIRE(n_deps >= 0 && n_deps <= DOPT_B200_MAX_INPUTS, "too many deps");
        db::Node n;
        n.type = op->op_type;
        n.op = *op;
        n.op.op_type = nullptr;
        n.bytes = db::volume(op->output) * 4;
        for (int i = 0; i < n_deps; ++i) {
            DB_REQUIRE(deps[i] >= 0 && deps[i] < (int)p->nodes.size(), "dep id out of range (nodes must be added in topological order)");
            n.deps.push_back(deps[i]);
            // operand types come from the graph, not from the caller
            n.op.inputs[i] = p->nodes[deps[i]].op.output;
        }
        n.op.n_inputs = n_deps;
        if (const_value && n.bytes > 0) n.const_value.assign((const uint8_t*)const_value, (const uint8_t*)const_value + n.bytes);
        p->nodes.push_back(std::move(n));
        return (int)p->nodes.size() - 1;
    } catch (const std::exception& e) {
        db::set_last_error(e.what());
        return -1;
    }
}

int dopt_b200_plan_set_outputs(dopt_b200_plan_t p, const int32_t* ids, int n) {
    PLAN_TRY
    DB_REQUIRE(p && (ids || n == 0), "null argument");
    p->outputs.clear();
    for (int i = 0; i < n; ++i) {
        DB_REQUIRE(ids[i] >= 0 && ids[i] < (int)p->nodes.size(), "output id out of range");
        p->outputs.push_back(ids[i]);
    }
    PLAN_CATCH
}

int dopt_b200_plan_finalize(dopt_b200_plan_t p, int flags) {
    PLAN_TRY
    DB_REQUIRE(p && !p->finalized, "null or already finalized plan");
    db::require_device();
    p->flags = flags;
    db::build(*p);
    p->finalized = true;
    PLAN_CATCH
}

int dopt_b200_plan_execute(dopt_b200_plan_t p, const int32_t* var_ids, const void* const* var_ptrs,
                           const int32_t* var_on_host, int n_vars, void* const* rets, int n_rets, void* stream) {
    PLAN_TRY
    DB_REQUIRE(p, "null plan");
    db::execute(*p, var_ids, var_ptrs, var_on_host, n_vars, rets, n_rets, (cudaStream_t)stream);
    PLAN_CATCH
}

int dopt_b200_plan_stats(dopt_b200_plan_t p, int64_t* launches, int64_t* device_bytes, int64_t* lowered_nodes) {
    PLAN_TRY
    DB_REQUIRE(p, "null plan");
    if (launches) *launches = p->launches_per_exec;
    if (device_bytes) *device_bytes = p->device_bytes;
    if (lowered_nodes) *lowered_nodes = (int64_t)p->order.size();
    PLAN_CATCH
}

int dopt_b200_plan_profile(dopt_b200_plan_t p, int enable, char* buf, size_t buf_len) {
    PLAN_TRY
    DB_REQUIRE(p, "null plan");
    if (buf && buf_len) {
        std::string s;
        for (auto& kv : p->prof_us) s += kv.first + "=" + std::to_string((long long)kv.second) + "\n";
        double tus = 0;
        int64_t tl = 0;
        db::tc_prof_read(&tus, &tl);
        s += "tc_kernel=" + std::to_string((long long)tus) + "\ntc_kernel_launches=" + std::to_string((long long)tl) + "\n";
        snprintf(buf, buf_len, "%s", s.c_str());
    }
    if (enable && !p->ev0) {
        DB_CUDA(cudaEventCreate(&p->ev0));
        DB_CUDA(cudaEventCreate(&p->ev1));
    }
    if (enable != (int)p->profiling) {
        p->prof_us.clear();
        db::tc_prof_enable(enable != 0);
    }
    p->profiling = enable != 0;
    PLAN_CATCH
}

int dopt_b200_plan_destroy(dopt_b200_plan_t p) {
    PLAN_TRY
    delete p;
    PLAN_CATCH
}

}  // extern "C"
