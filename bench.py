#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric: Wide-ResNet-28-10 CIFAR-shaped training images/sec on N B200s (one process per GPU).

  python bench.py --gpus N --steps K --warmup W              this repo's sm_100a kernels behind dopt's plugin API
  python bench.py --impl reference --gpus N --steps K ...    the reference path on the host CPU (oracle port), bounded sample

One "step" = one execution of the plan dopt.online.sgd compiles for WRN-28-10 + dense(100) + softmax + cross-entropy +
weight decay (examples/cifar100.d:33-49 with the BASELINE model/batch): forward, backward, SGD-momentum update of all 130
parameter tensors and the BN running-statistics write-back, on 128 synthetic 3x32x32 images per GPU.

Printed JSON (one line, rank 0):
  value        images/s over all ranks with the step's inputs already resident in HBM, device-timed (CUDA events, max over ranks)
  e2e          the same metric through the public call a dopt program makes -- updater([features: buffer(fs), labels: buffer(ls)])
               followed by reading loss and predictions back -- with the H2D copy from pinned host memory and the D2H read
               inside the timed region
  roofline     the dominant kernel (the tcgen05 implicit-GEMM kernel behind the three convolution ops): algorithmic conv
               FLOPs it executes per step / its summed launch time, both measured in this process with CUDA events
  cpu_baseline the oracle (CPU restatement of the reference) timed on this box's host cores on a bounded sample
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if WORLD > 1:
    # the gradient all-reduces get a fixed, small number of CTAs and the tensor-core kernels leave exactly those SMs free while
    # a bucket is in flight (dopt_b200/csrc/comm.cu, tc_host.cu); NCCL reads the variable when the process creates its first
    # communicator -- torch.distributed's, below -- so it has to be in the environment now
    os.environ.setdefault("NCCL_MAX_NCHANNELS", os.environ.get("DOPT_B200_COMM_CHANNELS", "16"))
    os.environ.setdefault("NCCL_MIN_NCHANNELS", os.environ["NCCL_MAX_NCHANNELS"])
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    ids = vis.split(",") if vis else [str(i) for i in range(64)]
    PHYS_GPU = ids[LOCAL_RANK]
    # Gradient buckets in peer-mapped memory reduced by the library's own NVSwitch-multicast kernel (DOPT_B200_SYMM=0: NCCL
    # only).  The symmetric-memory allocator tells peers apart by device index, so every GPU stays visible and rank r works
    # on device r; with DOPT_B200_SYMM=0 each process sees exactly its own device as ordinal 0 -- the reference hard-codes
    # ordinal 0 (cuda/source/dopt/cuda/package.d:44), so that is how the D host would be launched.
    SYMM = os.environ.get("DOPT_B200_SYMM", "1") != "0"
    if SYMM:
        DEV = LOCAL_RANK
    else:
        DEV = 0
        os.environ["CUDA_VISIBLE_DEVICES"] = PHYS_GPU
else:
    PHYS_GPU = (os.environ.get("CUDA_VISIBLE_DEVICES") or "0").split(",")[0]
    SYMM, DEV = False, 0

if "reference" in sys.argv and WORLD > 1 and RANK == 0:
    # the reference arm runs the host-CPU oracle on rank 0 alone: give it every host core -- torchrun exports
    # OMP_NUM_THREADS=1 to its workers, which would time a single-threaded BLAS (the numbers at N >= 2 of round 1)
    for var in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[var] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

MODEL = dict(depth=28, width=10, batch=128, hw=32, classes=100, lr=0.1, momentum=0.9, wd=1e-4)
# algorithmic conv+matmul FLOPs per image for one WRN-28-10 training step (3x forward minus the
# stem's unreachable dgrad); the three convolution ops carry all but ~0.01 % of it
TRAIN_GFLOP_PER_IMAGE = 31.459


# images per step of the bounded CPU sample (cpu_baseline and --impl reference): about 10 s of host work per step
CPU_SAMPLE = 16

# dram__bytes_read.sum + dram__bytes_write.sum of the 88 tc_kernel launches of one training step (ncu, final build of round 2:
# profiles/r02final_tc_dram.csv; the NHWC bf16 results mostly stay in L2 until a later kernel evicts them, hence the small
# write figure).  Re-measure with tools/capture_profiles.sh whenever the tensor-core kernel or the staging changes.
TC_DRAM_BYTES_PER_STEP = 2930144768 + 2820864


def workload_name(depth=MODEL["depth"], width=MODEL["width"], batch=MODEL["batch"]):
    return ("WRN-%d-%d + dense(100) + softmax/cross-entropy + weight decay, 3x32x32, batch %d/GPU, "
            "SGD lr 0.1 momentum 0.9 wd 1e-4 (examples/cifar100.d graph)" % (depth, width, batch))


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}
    try:
        p.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        p["source"] = "measured"
    except Exception:
        pass
    return p


def build_wrn(H, batch, depth=MODEL["depth"], width=MODEL["width"], reset=True):
    if reset:
        H.reset()
    H.seed(1234)
    x = H.float32((batch, 3, MODEL["hw"], MODEL["hw"]))
    y = H.float32((batch, MODEL["classes"]))
    preds = H.wide_resnet(x, depth, width, weight_decay=MODEL["wd"]).dense(MODEL["classes"]).softmax()
    net = H.Network([x], [preds])
    loss = H.cross_entropy(preds.train_output, y) + net.param_loss
    upd = H.Updater(H.SGD, [loss, preds.train_output], network=net,
                    hyper=[H.float32((), [MODEL["lr"]]), H.float32((), [MODEL["momentum"]])])
    net.preds = preds          # the output layer (its `.output` is the test-time graph, examples/cifar100.d:49)
    return x, y, net, upd


def adjacent_rows(torch, db, H, x, net, batch, hbm_gbs):
    """SURVEY.md section 8(f) rows 1 and 2, measured beside the training step (reported, never part of `value`):
      input_pipeline  ImageTransformer.getBatch + byte normalisation as one kernel (csrc/input.cu) over a CIFAR-sized training
                      set resident on the device as bytes, and over one 128-image batch; HBM-bound, 5 B per element
      inference       the test-time plan of the same network (batchNormInference), `testPlan.execute` with host buffers
    Each leg reports its own error instead of raising: the bench line must not depend on them."""
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)
    try:
        res = {}
        for tag, n, reps in (("train_set_50000", 50000, 10), ("batch_%d" % batch, batch, 50)):
            src = torch.randint(0, 256, (n, 3, MODEL["hw"], MODEL["hw"]), dtype=torch.uint8, device="cuda")
            per = db.jitter_sample(n, 4, 4, True, False, seed=7, call=0)          # cifar100.d: ImageTransformer(train, 4, 4, true, false)
            for _ in range(3):
                dst = db.image_transform(src, 4, 4, per)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(reps):
                dst = db.image_transform(src, 4, 4, per)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            alg = src.numel() * 5 + per.numel() * 4                              # 1 B read + 4 B written per element, the draws
            res[tag] = {"us_per_launch": us, "images_per_s": n / (us * 1e-6), "alg_bytes": int(alg),
                        "achieved_gbs": alg / (us * 1e-6) / 1e9, "frac_of_hbm": alg / (us * 1e-6) / 1e9 / hbm_gbs}
            del src, per, dst
        out["input_pipeline"] = res
    except Exception as e:
        out["input_pipeline"] = {"error": repr(e)}
    try:
        plan = H.Plan([net.preds.output])
        fs = synthetic_batch(batch, 99)[0]
        for _ in range(3):
            probs = plan.execute({x: fs})[0]
        torch.cuda.synchronize()
        reps = 10
        t0 = time.perf_counter()
        for _ in range(reps):
            probs = plan.execute({x: fs})[0]
        dt = (time.perf_counter() - t0) / reps
        st = plan.stats()
        out["inference"] = {"images_per_s": batch / dt, "ms_per_batch": dt * 1e3, "batch": batch,
                            "launches": st["launches"], "plan_device_bytes": st["device_bytes"],
                            "rows_sum_to_one": bool(abs(float(probs.sum(axis=1).mean()) - 1.0) < 1e-3),
                            "how": "Plan([preds.output]).execute({features: host array}): H2D of the batch, the test-time "
                                   "graph (batchNormInference with the relu and the NHWC bf16 staging of the next convolution in its apply pass), "
                                   "D2H of the class probabilities, wall clock"}
    except Exception as e:
        out["inference"] = {"error": repr(e)}
    return out


def synthetic_batch(batch, seed):
    rng = np.random.RandomState(seed)
    fs = (rng.rand(batch, 3, MODEL["hw"], MODEL["hw"]) * 2 - 1).astype(np.float32)   # loader normalisation x/128-1
    ls = np.eye(MODEL["classes"], dtype=np.float32)[rng.randint(0, MODEL["classes"], batch)]
    return fs, ls


# ---------------------------------------------------------------------------------------------------------------------
# per-kernel-class rooflines (the north star asks for the memory-bound classes as a fraction of HBM bandwidth)
# ---------------------------------------------------------------------------------------------------------------------
CONV_OPS = "convolution,convolutionFeaturesGrad,convolutionFiltersGrad"


def class_rooflines(nodes, loss_id, prof_us, prof_n, n_params, hbm_gbs, elem=4):
    """ALGORITHMIC bytes per step of the memory-bound op classes, from the exported dopt graph and SURVEY.md section 8(d)'s
    per-element figures (batchNormTrain 2V*s, batchNormGrad 3V*s -- the relu / reluGrad / NHWC staging the plan folds into
    those passes add no algorithmic bytes --, residual add 3V*s, SGD+momentum 5P*4), divided by the device time the plan's
    profiler attributes to that op type inside the training step.  s = `elem`: 4 with fp32 activations, 2 with the bf16
    interior ("element size s = 4 (fp32 drop-in path) or 2 (bf16 interior)") -- the smaller figure is used
    whenever the plan runs with bf16 interior activations, so the fraction is never flattered by bytes that are not moved.

    nodes: H.export(plan outputs); prof_us / prof_n: per-op-type microseconds and launches PER STEP."""
    by_id = dict((n["id"], n) for n in nodes)

    def vol(n):
        v = 1
        for d in n["shape"]:
            v *= d
        return v

    # forward nodes = ancestors of the loss
    fwd, stack = set(), [loss_id]
    while stack:
        i = stack.pop()
        if i in fwd:
            continue
        fwd.add(i)
        stack.extend(by_id[i]["deps"])
    v_bn = sum(vol(by_id[n["deps"][0]]) for n in nodes if n["type"] == "batchNormTrain")
    v_bng = sum(vol(by_id[n["deps"][1]]) for n in nodes if n["type"] == "batchNormGrad")
    res_adds = [n for n in nodes if n["type"] == "add" and len(n["shape"]) == 4 and n["id"] in fwd]
    out = {}

    def put(name, ops, alg_bytes, note):
        us = sum(prof_us.get(o, 0.0) for o in ops)
        if us <= 0 or alg_bytes <= 0:
            return
        gbs = alg_bytes / (us * 1e-6) / 1e9
        out[name] = {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs,
                     "alg_bytes_per_step": int(alg_bytes), "us_per_step": us,
                     "launches_per_step": int(sum(prof_n.get(o, 0) for o in ops)), "what": note}

    put("batchNormTrain", ["batchNormTrain"], 2 * v_bn * elem,
        "2V*%d B; the pass also applies relu and writes the NHWC bf16 operand of the next convolution" % elem)
    put("batchNormGrad", ["batchNormGrad"], 3 * v_bng * elem,
        "3V*%d B; the pass also applies the relu gate and the residual-gradient add" % elem)
    if res_adds and prof_n.get("add", 0) == len(res_adds):
        put("residual_add", ["add"], 3 * sum(vol(n) for n in res_adds) * elem, "3V*%d B, forward residual sums" % elem)
    put("optimiser", ["update"], 5 * n_params * 4,
        "5P*4 B (SGD+momentum): the plan's terminal fused launches, which also compute the weight-decay gradient; the other "
        "pointwise regions of the step (loss chain, scalar products) are booked under fusedRegion")
    return out


# ---------------------------------------------------------------------------------------------------------------------
# CPU reference arm / baseline
# ---------------------------------------------------------------------------------------------------------------------
def cpu_reference_run(steps, warmup, sample_batch, standalone=True):
    """The reference algorithm on the host: the oracle evaluates the very same dopt graph (built by the host mirror, which
    needs no GPU for that) for a bounded sample of `sample_batch` images per step."""
    from dopt_b200 import host as H
    from oracle import graph_eval as G
    H.init()
    x, y, net, upd = build_wrn(H, sample_batch, reset=standalone)
    oracle = G.UpdaterOracle(upd)
    fs, ls = synthetic_batch(sample_batch, 99)
    for _ in range(warmup):
        oracle.step({x: fs, y: ls})
    t0 = time.perf_counter()
    loss = None
    for _ in range(steps):
        loss = oracle.step({x: fs, y: ls})[0]
    dt = (time.perf_counter() - t0) / max(steps, 1)
    if standalone:
        H.reset()
    return sample_batch / dt, dt, float(loss)


def reference_arm(args):
    if RANK != 0:
        return
    sample = CPU_SAMPLE
    steps = max(1, min(args.steps, 2))
    warm = min(args.warmup, 1)
    ips, dt, loss = cpu_reference_run(steps, warm, sample)
    cores = os.cpu_count() or 1
    line = {
        "impl": "reference", "metric": "WRN-28-10 CIFAR train images/sec", "value": ips, "unit": "images/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(), "parallelism": "host cores of rank 0",
                   "sample": "%d images per step (bounded sample of the 128-image batch)" % sample},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port",
                         "sample": "%d steps of %d images, numpy/OpenBLAS oracle over the exported dopt graph" % (steps, sample)},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "the reference is D and cannot be built here (no D compiler); its CPU backend also lacks conv-gradient / "
                "batch-norm kernels, so the port supplies them from the cuDNN definitions (oracle/dopt_ref.py)",
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.f = tempfile.NamedTemporaryFile(prefix="clocks", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.f.name):
            c = [t.strip() for t in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            sm.sort()
            out.update(sm_mhz=sm[len(sm) // 2], sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dopt_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-adjacent", action="store_true", help="skip the input-pipeline / inference legs (SURVEY 8f rows)")
    ap.add_argument("--depth", type=int, default=MODEL["depth"])
    ap.add_argument("--width", type=int, default=MODEL["width"])
    ap.add_argument("--timeline", default=None,
                    help="diagnostics: trace 3 graph-replayed steps with CUPTI (torch.profiler) and write a per-kernel "
                         "summary (device time, launch count, idle gaps) to this file; prints no bench line")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
        return
    args.warmup = max(args.warmup, 3)

    import ctypes as C
    import torch
    import torch.distributed as dist
    import dopt_b200 as db
    from dopt_b200 import host as H

    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU path)"
    torch.cuda.set_device(DEV)
    world = WORLD
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", DEV))
    assert H.init(), H.init_error()
    symm_keep = None
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if RANK == 0:
            buf = C.create_string_buffer(128)
            db.check(db.lib.dopt_b200_comm_unique_id(buf))
            uid.copy_(torch.tensor(list(buf.raw), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        H.init_data_parallel(RANK, world, bytes(uid.cpu().tolist()))
        if SYMM:
            from dopt_b200 import symm
            # 146 MB of gradients + the running statistics, bucket by bucket (256-byte aligned slices)
            symm_keep = symm.attach(192 << 20, torch.device("cuda", DEV))

    B = MODEL["batch"]
    x, y, net, upd = build_wrn(H, B, args.depth, args.width)
    n_params = sum(p.volume for p in net.params)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # rotating synthetic batches (different data every step), identical weights on every rank, rank-specific data
    NB = 4
    batches = [synthetic_batch(B, 1234 + RANK * 100 + i) for i in range(NB)]
    pinned = [(torch.from_numpy(f).pin_memory(), torch.from_numpy(l).pin_memory()) for f, l in batches]
    device = [(f.cuda(), l.cuda()) for f, l in pinned]
    handles = (C.c_int * 2)(x.h, y.h)
    nbytes = (C.c_size_t * 2)(pinned[0][0].numel() * 4, pinned[0][1].numel() * 4)
    loss_out = np.zeros((), np.float32)
    pred_out = np.zeros((B, MODEL["classes"]), np.float32)
    out_ptrs = (C.c_void_p * 2)(loss_out.ctypes.data, pred_out.ctypes.data)

    def step_e2e(i):
        f, l = pinned[i % NB]
        upd.step_raw(handles, (C.c_void_p * 2)(f.data_ptr(), l.data_ptr()), nbytes, out_ptrs)

    def step_dev(i):
        f, l = device[i % NB]
        upd.step_device(handles, (C.c_void_p * 2)(f.data_ptr(), l.data_ptr()))

    # ---- warm-up (sizes workspaces, captures the CUDA graph) ----
    for i in range(args.warmup):
        step_dev(i)
    for i in range(2):
        step_e2e(i)
    barrier()
    first_loss = float(loss_out)

    if args.timeline:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            for i in range(3):
                step_dev(i)
            torch.cuda.synchronize()
        evs = sorted((e for e in prof.events() if e.device_type.name == "CUDA" or "cuda" in str(e.device_type).lower()),
                     key=lambda e: e.time_range.start)
        by, busy, gaps, last_end, last_name, gap_by = {}, 0.0, 0.0, None, None, {}
        for e in evs:
            dur = e.time_range.end - e.time_range.start
            name = e.name.replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "")
            a = by.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += dur
            busy += dur
            if last_end is not None and e.time_range.start > last_end:
                gaps += e.time_range.start - last_end
                g = gap_by.setdefault((last_name.split("<")[0], name.split("<")[0]), [0, 0.0])
                g[0] += 1
                g[1] += e.time_range.start - last_end
            if last_end is None or e.time_range.end >= last_end:
                last_name = name
            last_end = max(last_end or 0, e.time_range.end)
        with open(args.timeline, "w") as f:
            f.write("# 3 graph-replayed steps, CUPTI kernel trace (torch.profiler); times in us per step\n")
            f.write("busy %.1f  idle-between-kernels %.1f  kernels %d\n" % (busy / 3, gaps / 3, len(evs) // 3))
            for k, (n, us) in sorted(by.items(), key=lambda kv: -kv[1][1]):
                f.write("%-70s %5d %10.1f\n" % (k[:70], n // 3, us / 3))
            f.write("idle time by (kernel that ended, kernel that started), us per step:\n")
            for (a_, b_), (n, us) in sorted(gap_by.items(), key=lambda kv: -kv[1][1])[:14]:
                f.write("  %-34s -> %-34s %4d %8.1f\n" % (a_[:34], b_[:34], n // 3, us / 3))
            span = evs[-1].time_range.end - evs[0].time_range.start
            f.write("span of the 3 steps %.1f us (%.1f per step)\n" % (span, span / 3))
            # the tensor-core launches of the last traced step, in launch order (us)
            tc = [(e.name.split("(")[0].replace("void ", "").replace("db::", ""), e.time_range.end - e.time_range.start)
                  for e in evs if "tc_kernel" in e.name]
            tc = tc[2 * len(tc) // 3:]
            f.write("tc_kernel launches of one step, in order:\n")
            f.write(" ".join("%s:%.0f" % (n.replace("tc_kernel", ""), d) for n, d in tc) + "\n")
        # every device activity of the last traced step with its stream: start (us from the step's first kernel), duration, stream
        try:
            kev = [k for k in prof.profiler.kineto_results.events() if "cuda" in str(k.device_type()).lower()]
            kev.sort(key=lambda k: k.start_ns())
            starts = [i for i, k in enumerate(kev) if "pack_filters_multi" in k.name() or "gate_step_begin" in k.name()]
            first = [i for j, i in enumerate(starts) if j == 0 or starts[j - 1] < i - 50]   # first launch of each step
            lo = first[-1] if first else 0
            t0 = kev[lo].start_ns()
            with open(args.timeline + ".raw", "w") as f:
                f.write("# start_us dur_us stream name   (last of 3 traced steps)\n")
                for k in kev[lo:]:
                    nm = k.name().replace("(anonymous namespace)::", "").split("(")[0].replace("void ", "").replace("db::", "")
                    f.write("%9.1f %8.1f %4d %s\n" % ((k.start_ns() - t0) / 1e3, k.duration_ns() / 1e3, k.device_resource_id(), nm[:80]))
        except Exception as ex:   # diagnostics only
            sys.stderr.write("timeline raw dump failed: %r\n" % (ex,))
        return

    # ---- timed: device-resident inputs ----
    sampler = ClockSampler(PHYS_GPU) if RANK == 0 else None
    launches0 = db.lib.dopt_b200_launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        step_dev(i)
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    launches = db.lib.dopt_b200_launch_count() - launches0

    # ---- timed: end to end through the public call (H2D from pinned memory + D2H of loss and predictions) ----
    barrier()
    t0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        step_e2e(i)
    e1.record()
    barrier()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    clocks = sampler.stop() if sampler else None
    last_loss = float(loss_out)

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = float(t[0]), float(t[1])

    # ---- per-op device times: every plan item between two CUDA events (serialised, eager; includes the launch latency the
    # captured graph hides, so it is an upper bound per op type) ----
    roof = None
    # every rank runs the profiled steps (the plan contains the gradient all-reduce); rank 0 reports
    upd.profile(True)
    for i in range(2):
        step_dev(i)
    torch.cuda.synchronize()
    prof = upd.profile(False)
    # ---- rooflines: the launches of ONE kernel class re-issued back to back with the operands of the last step, between one
    # pair of CUDA events (dopt_b200_plan_replay_class) -- the class's device time without an event bracket around every
    # launch.  This scrambles the plan's accumulating state, so it is the last thing this process does with the updater. ----
    replay = {}
    if world == 1:
        for name, ops in (("tc", CONV_OPS), ("batchNormTrain", "batchNormTrain"), ("batchNormGrad", "batchNormGrad"),
                          ("add", "add"), ("update", "update")):
            try:
                replay[name] = upd.replay_class(ops, 3)
            except Exception as e:   # diagnostics: never lose the bench line over it
                replay[name + "_error"] = repr(e)
        torch.cuda.synchronize()
    if RANK == 0:
        pk = peaks()
        tc_us = prof.get("tc_kernel", 0) / 2.0
        tc_us_events = tc_us
        if "tc" in replay:
            # the three convolution ops of the step: 88 tcgen05 launches + the 3-channel stem's direct kernel + the memsets of
            # the filter-gradient accumulators (part of those ops), against the same algorithmic FLOPs
            tc_us = replay["tc"][0]
        conv_flops = TRAIN_GFLOP_PER_IMAGE * 1e9 * B if (args.depth, args.width) == (28, 10) else None
        if tc_us > 0 and conv_flops:
            achieved = conv_flops / (tc_us * 1e-6) / 1e12
            # which measured peak applies is decided by this run's own clock record: the burst figure when the SM clock
            # stayed at its maximum without a power cap (a 0.2 s timed region does), the sustained one otherwise
            burst = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and
                         clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"] and "sw_power_cap" not in clocks.get("reasons", []))
            peak = pk["bf16_tflops"] if burst else pk["bf16_tflops_sustained"]
            roof = {"bound": "tensor", "kernel": "tc_kernel<CONV|WGRAD> (tcgen05 implicit GEMM)", "achieved": achieved,
                    "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                    "frac_of_burst_peak": achieved / pk["bf16_tflops"],
                    "frac_of_sustained_peak": achieved / pk["bf16_tflops_sustained"],
                    "traffic": TC_DRAM_BYTES_PER_STEP if (args.depth, args.width) == (28, 10) else None,
                    "traffic_source": "ncu dram__bytes_read+write summed over the step's tc_kernel launches "
                                      "(profiles/r02final_tc_dram.csv), per step like `achieved`",
                    "peak_source": pk["source"] + (" (burst: SM clock at max, no power cap during the timed region)" if burst
                                                   else " (sustained: SM clock below max or power-capped during the timed region)"),
                    "launches_per_step": prof.get("tc_kernel_launches", 0) / 2.0, "kernel_ms_per_step": tc_us / 1e3,
                    "how": ("the step's convolution launches re-issued back to back between one CUDA-event pair "
                            "(dopt_b200_plan_replay_class; includes the stem's direct kernel and the memset of the filter-gradient scratch arena)"
                            if "tc" in replay else "sum of per-launch CUDA-event brackets around the tcgen05 kernel"),
                    "kernel_ms_per_step_event_brackets": tc_us_events / 1e3}
    classes = None
    if RANK == 0:
        try:
            plan_outs, _ = upd.plan_outputs()
            nodes = H.export(plan_outs)
            loss_id = [n["id"] for n in nodes if n["op"].h == plan_outs[0].h][0]
            per_us = dict((k, v / 2.0) for k, v in prof.items() if "#" not in k)
            per_n = dict((k[:-2], v / 2.0) for k, v in prof.items() if k.endswith("#n"))
            for name in ("batchNormTrain", "batchNormGrad", "add", "update"):
                if name in replay:      # class replay (one event pair per class) instead of per-launch event brackets
                    per_us[name] = replay[name][0]
            interior = bool(H.plan_flags() & db._lib.PLAN_BF16_INTERIOR)
            classes = class_rooflines(nodes, loss_id, per_us, per_n, n_params, peaks()["hbm_gbs"], elem=2 if interior else 4)
        except Exception as e:  # diagnostics only: never lose the bench line over it
            classes = {"error": repr(e)}
    barrier()

    if RANK == 0:
        st = upd.stats()
        cpu = None
        if not args.no_cpu_baseline and world == 1:    # reported at N=1 only (rank 0's host cores)
            ips, dt, _ = cpu_reference_run(1, 0, CPU_SAMPLE, standalone=False)
            cpu = {"value": ips, "unit": "images/s", "cores": os.cpu_count() or 1, "kind": "port",
                   "sample": "1 step of %d images of the same WRN-28-10 train graph, numpy/OpenBLAS oracle (%.1f s)"
                             % (CPU_SAMPLE, dt)}
        ms_step = dev_ms / args.steps
        adjacent = None
        if world == 1 and not args.no_adjacent:
            adjacent = adjacent_rows(torch, db, H, x, net, B, peaks()["hbm_gbs"])
        line = {
            "metric": "WRN-28-10 CIFAR train images/sec", "value": B * world * args.steps / (dev_ms * 1e-3),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_name(args.depth, args.width, B),
                       "parallelism": "dp%d" % world, "params": n_params,
                       "exchange": (None if world == 1 else
                                    "gradient buckets in peer-mapped memory, reduced by the library's multimem kernel over NVSwitch "
                                    "multicast" if (symm_keep is not None and not os.environ.get("DOPT_B200_NO_NVLS")) else
                                    "gradient buckets reduced by ncclAllReduce"),
                       "cache": "inputs larger than L2: one step streams several GB of activations through the 126 MB L2",
                       "precision": "convolutions bf16 operands / fp32 accumulate on tcgen05; activations between tensor-core "
                                    "convolutions stored NHWC bf16 (plan flag BF16_INTERIOR), arithmetic and everything "
                                    "else fp32" if (H.plan_flags() & db._lib.PLAN_BF16_INTERIOR) else
                                    "convolutions bf16 operands / fp32 accumulate on tcgen05; everything else fp32"},
            "e2e": {"value": B * world * args.steps / (e2e_ms * 1e-3), "unit": "images/s",
                    "h2d_bytes_per_step": int(nbytes[0] + nbytes[1]), "d2h_bytes_per_step": int(4 + pred_out.nbytes),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "launches_per_step": st["launches"], "plan_nodes": st["lowered_nodes"],
            "plan_device_bytes": st["device_bytes"], "clocks": clocks, "roofline": roof, "roofline_classes": classes,
            "cpu_baseline": cpu, "adjacent_rows": adjacent,
            "loss_first": first_loss, "loss_last": last_loss,
            "per_op_us_per_step": dict((k, [v / 2.0, prof.get(k + "#n", 0) // 2]) for k, v in
                                       sorted(((k, v) for k, v in prof.items() if "#" not in k and k != "tc_kernel_launches"),
                                              key=lambda kv: -kv[1])[:16]),
        }
        print(json.dumps(line))
    if world > 1:
        db.check(db.lib.dopt_b200_comm_check())   # a flag barrier of the multicast all-reduce that timed out, an NCCL error
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
