"""Graph-level CPU oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/dopt_ref.py for the rules).

Restates the reference CPU evaluator `evaluateCPU` (cpu/source/dopt/cpu/package.d:182-280): walk the topologically sorted
graph, evaluate every node with the oracle kernels of dopt_ref.py, return the requested outputs.  The graph itself comes
from the host library's export (dopt_b200.host.export), i.e. it is the graph the product built -- the same thing the
reference's CPU backend would be handed by `dopt.core`.

`UpdaterOracle` replays what the closure returned by dopt.online.{sgd,adam,amsgrad} does (online/source/dopt/online/sgd.d:86-91):
evaluate (outputs ~ new values ~ state) and copy the new values over the variables' buffers.  It keeps its OWN copy of every
variable, so it can be stepped side by side with the GPU updater for loss-curve comparisons.
"""
import numpy as np

from . import dopt_ref as R


def evaluate(nodes, outputs_ids, values, args=None, overrides=None):
    """nodes: dopt_b200.host.export() result; values: {node id: ndarray} for variables / constants; args override.
    overrides: {node id: ndarray} fixes the value of any other node -- used for `uniform` nodes (the reference draws them
    unseeded, cuda/source/dopt/cuda/random.d:56-83, so a comparison has to share the draw)."""
    env = {}
    args = args or {}
    overrides = overrides or {}
    for n in nodes:
        t = n["type"]
        if n["id"] in overrides:
            env[n["id"]] = np.asarray(overrides[n["id"]], dtype=n["dtype"]).reshape(n["shape"])
            continue
        if t in ("variable", "constant"):
            v = args.get(n["id"], values.get(n["id"]))
            if v is None:
                raise KeyError("no value for %s node %d" % (t, n["id"]))
            env[n["id"]] = np.asarray(v, dtype=n["dtype"]).reshape(n["shape"])
            continue
        ins = [env[d] for d in n["deps"]]
        if t == "allreduce":
            out = ins[0]  # single-process oracle: the mean over one rank
        else:
            out = R.evaluate_op(t, ins, n["attrs"], n["shape"])
        env[n["id"]] = np.asarray(out, dtype=n["dtype"]).reshape(n["shape"])
    return [env[i] for i in outputs_ids]


def snapshot_values(nodes):
    """Current contents of every variable / constant the graph reads (fetched through the host API)."""
    return {n["id"]: n["op"].get() for n in nodes if n["type"] in ("variable", "constant")}


def evaluate_ops(host, outputs, args=None):
    """Oracle value of host Ops: export, snapshot, evaluate.  `args` maps host Ops to arrays."""
    nodes = host.export(outputs)
    vals = snapshot_values(nodes)
    a = {}
    for k, v in (args or {}).items():
        a[k.serial] = v
    return evaluate(nodes, [o.serial for o in outputs], vals, a)


class UpdaterOracle(object):
    def __init__(self, updater):
        from dopt_b200 import host
        self.plan_ops, self.dest_ops = updater.plan_outputs()
        self.n_user = sum(1 for d in self.dest_ops if d is None)
        self.nodes = host.export(self.plan_ops)
        self.values = snapshot_values(self.nodes)
        self.out_ids = [o.serial for o in self.plan_ops]
        self.dest_ids = [d.serial if d is not None else None for d in self.dest_ops]

    def step(self, args, overrides=None):
        a = dict((k.serial, v) for k, v in args.items())
        o = dict((k.serial, v) for k, v in (overrides or {}).items())
        outs = evaluate(self.nodes, self.out_ids, self.values, a, o)
        # all reads happen before any write-back, like the D2D copies after the last node (package.d:419-422)
        for val, dest in zip(outs, self.dest_ids):
            if dest is not None:
                self.values[dest] = val.copy()
        return outs[:self.n_user]

    def value_of(self, op):
        return self.values[op.serial]
