"""Run under torchrun: does torch's symmetric memory (peer-mapped buffers + NVSwitch multicast) come up in this environment with
one visible GPU per process, and how fast is its own multimem all-reduce on a gradient-bucket-sized buffer?  Diagnostics."""
import os
import sys
import time

LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
RANK = int(os.environ.get("RANK", "0"))
if os.environ.get("PROBE_ALL_VISIBLE") != "1":
    os.environ["CUDA_VISIBLE_DEVICES"] = str(LOCAL_RANK)
    DEV = 0
else:
    DEV = LOCAL_RANK

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402


def log(*a):
    if RANK == 0:
        print(*a, flush=True)


def main():
    torch.cuda.set_device(DEV)
    dist.init_process_group("nccl", device_id=torch.device("cuda", DEV))
    group = dist.group.WORLD
    n = 36 * (1 << 20) // 4
    try:
        symm_mem.enable_symm_mem_for_group(group.group_name)
    except Exception as e:
        log("enable_symm_mem_for_group:", repr(e))
    t = symm_mem.empty(n, dtype=torch.float32, device=torch.device("cuda", DEV))
    hdl = symm_mem.rendezvous(t, group)
    log("rendezvous ok: world", hdl.world_size, "multicast", hdl.has_multicast_support(DeviceTypeOrNone(), DEV) if False else "?",
        "mc_ptr", hex(hdl.multicast_ptr), "signal_pad_size", hdl.signal_pad_size, "buffer_ptrs", [hex(p) for p in hdl.buffer_ptrs][:4])
    t.fill_(float(RANK + 1))
    torch.cuda.synchronize()
    dist.barrier()
    peer = hdl.get_buffer((RANK + 1) % WORLD, (16,), torch.float32)
    ok = float(peer[0]) == float((RANK + 1) % WORLD + 1)
    log("peer view ok:", ok)
    dist.barrier()
    for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
        try:
            op = getattr(torch.ops.symm_mem, name)
            t.fill_(1.0)
            torch.cuda.synchronize()
            dist.barrier()
            r = op(t, "sum", group.group_name)
            torch.cuda.synchronize()
            val = float((r if r is not None else t)[12345])
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dist.barrier()
            e0.record()
            for _ in range(20):
                op(t, "sum", group.group_name)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            log("%s: value %.1f (want %d), 36 MB in %.0f us = %.0f GB/s" % (name, val, WORLD, us, n * 4 / us / 1e3))
        except Exception as e:
            log(name, "FAILED", repr(e)[:300])
    dist.barrier()
    dist.destroy_process_group()


class DeviceTypeOrNone:
    pass


if __name__ == "__main__":
    main()
