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

}  // namespace online
}  // namespace dopt
