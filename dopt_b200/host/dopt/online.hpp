// dopt/online.hpp -- C++ mirror of dopt.online: sgd / adam / amsgrad build the update rule as graph, compile ONE plan for
// (outputs ~ new values ~ optimiser state) and return a closure that executes it with the parameters' own buffers as the
// destinations (online/source/dopt/online/{package,sgd,adam,amsgrad}.d).
#pragma once
#include "core.hpp"
#include "nnet.hpp"

namespace dopt {
namespace online {

using Projection = nnet::Projection;                                                        // online/package.d:28
using Updater = std::function<std::vector<Buffer>(const std::map<Operation, Buffer>&)>;      // online/package.d:23

Updater sgd(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
            const std::map<Operation, Projection>& projs, Operation learningRate = nullptr, Operation momentumRate = nullptr,
            bool nesterov = false);
Updater adam(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
             const std::map<Operation, Projection>& projs, Operation alpha = nullptr, Operation beta1 = nullptr,
             Operation beta2 = nullptr, Operation eps = nullptr);
Updater amsgrad(const std::vector<Operation>& outputs, const std::vector<Operation>& wrt,
                const std::map<Operation, Projection>& projs, Operation alpha = nullptr, Operation beta1 = nullptr,
                Operation beta2 = nullptr, Operation eps = nullptr);

// what the last sgd / adam / amsgrad call compiled, for tests and benchmarks: the plan, and the list of operations it
// evaluates / the variables it writes back to (same order as the plan outputs)
struct LastUpdate {
    PlanPtr plan;
    std::vector<Operation> planOutputs;    // outputs ~ newvals ~ state
    std::vector<Operation> destinations;   // null for the user-visible outputs, else the variable overwritten
    std::vector<Buffer> newbufs;
};
const LastUpdate& lastUpdate();

// Training checkpoint: everything the updater overwrites every step -- the parameters, the
// optimiser state (momenta / means / variances / varhats) and Adam's running b1, b2 -- in the order of `destinations`.
// The reference only has DAGNetwork.save/load (nnet/networks.d:130-164: raw fp32 of the parameters, no header, no optimiser
// state), so a resumed run restarts its momenta from zero.  File layout:
//   "DOPTCKPT" | u32 version (1) | u32 n_tensors | n_tensors x u64 element count | raw fp32 of every tensor in order
// With a network-built updater the tensors start with the network's parameters in DAGNetwork order, so the first part of
// the body is byte-for-byte the file DAGNetwork::save writes.  loadState refuses a file whose tensor list does not match.
void saveState(const LastUpdate& u, const std::string& filename);
void loadState(const LastUpdate& u, const std::string& filename);
size_t stateHeaderBytes(const LastUpdate& u);

}  // namespace online
}  // namespace dopt
