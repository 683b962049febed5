// dopt/online.cpp -- see online.hpp.
#include "online.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <set>

namespace dopt {
namespace online {

static LastUpdate g_last;
const LastUpdate& lastUpdate() { return g_last; }

// ---- training checkpoint ---------------------------------------------------------------------------------------------------
static const char kCkptMagic[8] = {'D', 'O', 'P', 'T', 'C', 'K', 'P', 'T'};
static const uint32_t kCkptVersion = 1;

static std::vector<Operation> stateTensors(const LastUpdate& u) {
    std::vector<Operation> v;
    for (auto& d : u.destinations)
        if (d) v.push_back(d);
    return v;
}
size_t stateHeaderBytes(const LastUpdate& u) { return 8 + 4 + 4 + 8 * stateTensors(u).size(); }

void saveState(const LastUpdate& u, const std::string& filename) {
    auto tensors = stateTensors(u);
    for (auto& t : tensors) enforce(t->elementType() == DataType::float32, "checkpoint: only float32 state is supported");
    FILE* f = std::fopen(filename.c_str(), "wb");
    enforce(f != nullptr, "cannot open " + filename);
    bool ok = std::fwrite(kCkptMagic, 1, 8, f) == 8;
    uint32_t n = (uint32_t)tensors.size();
    ok = ok && std::fwrite(&kCkptVersion, 4, 1, f) == 1 && std::fwrite(&n, 4, 1, f) == 1;
    for (auto& t : tensors) {
        uint64_t vol = t->volume();
        ok = ok && std::fwrite(&vol, 8, 1, f) == 1;
    }
    for (auto& t : tensors) {
        auto v = t->value()->get<float>();
        ok = ok && std::fwrite(v.data(), sizeof(float), v.size(), f) == v.size();
    }
    ok = (std::fclose(f) == 0) && ok;
    enforce(ok, "short write to " + filename);
}

void loadState(const LastUpdate& u, const std::string& filename) {
    auto tensors = stateTensors(u);
    FILE* f = std::fopen(filename.c_str(), "rb");
    enforce(f != nullptr, "cannot open " + filename);
    auto fail = [&](const std::string& why) {
        std::fclose(f);
        throw Exception("checkpoint " + filename + ": " + why);
    };
    char magic[8];
    uint32_t version = 0, n = 0;
    if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, kCkptMagic, 8) != 0) fail("not a dopt_b200 checkpoint");
    if (std::fread(&version, 4, 1, f) != 1 || version != kCkptVersion) fail("unsupported version");
    if (std::fread(&n, 4, 1, f) != 1 || n != tensors.size()) fail("tensor count does not match this updater");
    for (auto& t : tensors) {
        uint64_t vol = 0;
        if (std::fread(&vol, 8, 1, f) != 1 || vol != (uint64_t)t->volume()) fail("tensor sizes do not match this updater");
    }
    // read everything before touching any variable: a truncated file must not leave a half-restored model behind
    std::vector<std::vector<float>> data;
    for (auto& t : tensors) {
        data.emplace_back(t->volume());
        if (std::fread(data.back().data(), sizeof(float), data.back().size(), f) != data.back().size()) fail("file is too short");
    }
    std::fclose(f);
    for (size_t i = 0; i < tensors.size(); ++i) tensors[i]->value()->set(data[i].data(), data[i].size() * sizeof(float));
}

// A "gradient" that is a plain variable / constant is the zero tensor the batch-norm gradient rule hands out for the running
// mean and variance (core/source/dopt/core/grads/nnet.d:80-81): such parameters are state, moved only by their projection.
static bool isStateOnly(const Operation& g) { return g->opType() == "variable" || g->opType() == "constant"; }

// Does `op` depend only on the parameters being updated and on constants?  Such a value is the same on every rank: the
// replicas start from identical parameters and apply identical (exchanged) updates.  `same` holds the variables known to be
// identical everywhere: the parameters, and variables that are not part of the objective's own graph -- those were created by
// grad() (its seed float32([], [1.0f]), grads/package.d:52) and nobody holds a handle to feed them.  Anything reached through
// another variable (the minibatch, a hyper-parameter fed per call) or through a random draw is treated as rank-local.
static bool rankInvariant(const Operation& op, const std::set<const OperationNode*>& wrt,
                          std::map<const OperationNode*, bool>& memo) {
    auto it = memo.find(op.get());
    if (it != memo.end()) return it->second;
    bool inv;
    if (op->opType() == "constant") inv = true;
    else if (op->opType() == "variable") inv = true;   // a parameter or grad()'s seed (the caller's variables are pre-marked)
    else if (op->opType() == "uniform" || op->opType() == "allreduce" || op->deps().empty()) inv = false;
    else {
        inv = true;
        for (auto& d : op->deps())
            if (!rankInvariant(d, wrt, memo)) {
                inv = false;
                break;
            }
    }
    memo[op.get()] = inv;
    return inv;
}

// mean over ranks of g.  The mean is linear and a rank-invariant addend is its own mean, so add(a, b) with b rank-invariant
// becomes add(allreduce(a), b): the weight-decay term of a filter gradient (2 * wd * W, nnet/layers/conv.d's weight decay
// through grads/math.d) is not copied into a gradient bucket and sent over NVLink, the filter-gradient kernel writes straight
// into the bucket, and the term is added by the same fused update kernel as on a single GPU.  Rounding differs from
// allreduce(a + b) only by the order of two fp32 additions.
static Operation exchangeOne(const Operation& g, const std::set<const OperationNode*>& wrt,
                             std::map<const OperationNode*, bool>& memo) {
    if (g->opType() == "add" && g->deps().size() == 2 && g->deps()[0]->shape() == g->shape() && g->deps()[1]->shape() == g->shape()) {
        const bool i0 = rankInvariant(g->deps()[0], wrt, memo), i1 = rankInvariant(g->deps()[1], wrt, memo);
        if (i1 && !i0) return exchangeOne(g->deps()[0], wrt, memo) + g->deps()[1];
        if (i0 && !i1) return g->deps()[0] + exchangeOne(g->deps()[1], wrt, memo);
    }
    return createOperation("allreduce", {g});
}

// data-parallel: the mean over ranks of every gradient, as a registered `allreduce` op between grad() and the update rule.
// Zero "gradients" of state-only parameters are identical on every rank and are not exchanged.
static std::vector<Operation> exchange(std::vector<Operation> grads, const std::vector<Operation>& wrt, const Operation& objective) {
    if (dataParallelWorld() <= 1) return grads;
    std::set<const OperationNode*> params;
    for (auto& w : wrt) params.insert(w.get());
    std::map<const OperationNode*, bool> memo;
    for (auto& op : topologicalSort({objective}))
        if (op->opType() == "variable" && !params.count(op.get())) memo[op.get()] = false;   // fed by the caller: rank-local
    const bool split = std::getenv("DOPT_B200_NO_EXCHANGE_SPLIT") == nullptr;
    for (auto& g : grads) {
        if (isStateOnly(g)) continue;
        if (rankInvariant(g, params, memo)) continue;   // (a parameter that only the regulariser touches)
        g = split ? exchangeOne(g, params, memo) : createOperation("allreduce", {g});
    }
    return grads;
}

static Updater finish(const std::vector<Operation>& outputs, std::vector<Operation> planOutputs,
                      const std::vector<Operation>& stateVars) {
    // sgd.d:76-91 / adam.d:75-90
    auto updatePlan = compile(planOutputs);
    std::vector<Buffer> newbufs;
    std::vector<Operation> dests;
    for (auto& o : outputs) {
        newbufs.push_back(allocate(o->volume() * sizeOf(o->elementType())));
        dests.push_back(nullptr);
    }
    for (auto& v : stateVars) {
        newbufs.push_back(v->value());
        dests.push_back(v);
    }
    g_last = LastUpdate{updatePlan, planOutputs, dests, newbufs};
    size_t nOut = outputs.size();
    return [updatePlan, newbufs, nOut](const std::map<Operation, Buffer>& args) mutable {
        updatePlan->execute(args, newbufs);
        return std::vector<Buffer>(newbufs.begin(), newbufs.begin() + nOut);
    };
}

// `rawGrads` are the gradients before the exchange.  Data-parallel: the projection of a state-only parameter (the batch-norm
// running statistics, nnet/layers/batchnorm.d:140-154, computed from the rank's own batch) is averaged over the ranks, so
// every replica -- and any checkpoint or inference plan made from it -- carries the same running mean / variance.
static void applyProjections(const std::vector<Operation>& wrt, const std::map<Operation, Projection>& projs,
                             std::vector<Operation>& newvals, const std::vector<Operation>& rawGrads) {
    for (size_t i = 0; i < newvals.size(); ++i) {
        auto it = projs.find(wrt[i]);
        if (it == projs.end() || !it->second) continue;
        newvals[i] = it->second(newvals[i]);
        if (dataParallelWorld() > 1 && isStateOnly(rawGrads[i])) newvals[i] = createOperation("allreduce", {newvals[i]});
    }
}

Updater sgd(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
            const std::map<Operation, Projection>& projs, Operation learningRate, Operation momentumRate, bool nesterov) {
    // sgd.d:28-94
    if (!learningRate) learningRate = float32({}, {0.01f});
    if (!momentumRate) momentumRate = float32({}, {0.0f});
    auto objective = outputs[0];
    auto rawGrads = grad(objective, wrt);
    auto grads = exchange(rawGrads, wrt, objective);
    std::vector<Operation> momentum, newMomentum, newvals;
    for (auto& g : grads) momentum.push_back(float32(g->shape()));
    if (nesterov) {
        for (size_t i = 0; i < grads.size(); ++i) newMomentum.push_back(momentum[i] * momentumRate - learningRate * grads[i]);
        for (size_t i = 0; i < grads.size(); ++i)
            newvals.push_back(wrt[i] + momentumRate * newMomentum[i] - learningRate * grads[i]);
    } else {
        for (size_t i = 0; i < grads.size(); ++i) newMomentum.push_back(momentum[i] * momentumRate + learningRate * grads[i]);
        for (size_t i = 0; i < grads.size(); ++i) newvals.push_back(wrt[i] - newMomentum[i]);
    }
    applyProjections(wrt, projs, newvals, rawGrads);
    std::vector<Operation> planOutputs(outputs);
    planOutputs.insert(planOutputs.end(), newvals.begin(), newvals.end());
    planOutputs.insert(planOutputs.end(), newMomentum.begin(), newMomentum.end());
    std::vector<Operation> state(wrt);
    state.insert(state.end(), momentum.begin(), momentum.end());
    return finish(outputs, planOutputs, state);
}

static Updater adamImpl(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
                        const std::map<Operation, Projection>& projs, Operation alpha, Operation beta1, Operation beta2,
                        Operation eps, bool ams) {
    // adam.d:32-93, amsgrad.d:32-99
    if (!alpha) alpha = float32({}, {0.001f});
    if (!beta1) beta1 = float32({}, {0.9f});
    if (!beta2) beta2 = float32({}, {0.999f});
    if (!eps) eps = float32({}, {1e-8f});
    auto objective = outputs[0];
    auto rawGrads = grad(objective, wrt);
    auto grads = exchange(rawGrads, wrt, objective);
    std::vector<Operation> means, vars, varhats;
    for (auto& w : wrt) means.push_back(float32(w->shape()));
    for (auto& w : wrt) vars.push_back(float32(w->shape()));
    if (ams)
        for (auto& w : wrt) varhats.push_back(float32(w->shape()));
    auto b1 = float32({}, {1.0f});
    auto b2 = float32({}, {1.0f});
    auto nb1 = b1 * beta1;
    auto nb2 = b2 * beta2;
    auto eta = alpha * sqrt(1.0f - nb2) / (1.0f - nb1);
    std::vector<Operation> newMeans, newVars, newVarhats, newvals;
    for (size_t i = 0; i < wrt.size(); ++i) newMeans.push_back(beta1 * means[i] + (1.0f - beta1) * grads[i]);
    for (size_t i = 0; i < wrt.size(); ++i) newVars.push_back(beta2 * vars[i] + (1.0f - beta2) * grads[i] * grads[i]);
    if (ams)
        for (size_t i = 0; i < wrt.size(); ++i) newVarhats.push_back(max(varhats[i], vars[i]));   // amsgrad.d:63-66 (survey F11)
    for (size_t i = 0; i < wrt.size(); ++i) newvals.push_back(wrt[i] - eta * (newMeans[i] / (sqrt(newVars[i]) + eps)));
    applyProjections(wrt, projs, newvals, rawGrads);
    std::vector<Operation> planOutputs(outputs);
    planOutputs.insert(planOutputs.end(), newvals.begin(), newvals.end());
    planOutputs.insert(planOutputs.end(), newMeans.begin(), newMeans.end());
    planOutputs.insert(planOutputs.end(), newVars.begin(), newVars.end());
    if (ams) planOutputs.insert(planOutputs.end(), newVarhats.begin(), newVarhats.end());
    planOutputs.push_back(nb1);
    planOutputs.push_back(nb2);
    std::vector<Operation> state(wrt);
    state.insert(state.end(), means.begin(), means.end());
    state.insert(state.end(), vars.begin(), vars.end());
    if (ams) state.insert(state.end(), varhats.begin(), varhats.end());
    state.push_back(b1);
    state.push_back(b2);
    return finish(outputs, planOutputs, state);
}

Updater adam(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
             const std::map<Operation, Projection>& projs, Operation alpha, Operation beta1, Operation beta2, Operation eps) {
    return adamImpl(outputs, wrt, projs, alpha, beta1, beta2, eps, false);
}
Updater amsgrad(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
                const std::map<Operation, Projection>& projs, Operation alpha, Operation beta1, Operation beta2, Operation eps) {
    return adamImpl(outputs, wrt, projs, alpha, beta1, beta2, eps, true);
}

}  // namespace online
}  // namespace dopt
