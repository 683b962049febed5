"""Plumbing for dopt_b200_comm_set_symmetric (include/dopt_b200.h): obtains one peer-mapped buffer per rank plus the NVSwitch
multicast mapping over all of them from torch.distributed's symmetric-memory allocator (CUDA VMM + multicast objects underneath)
and hands the raw addresses to the library.  The gradient-bucket arenas of every plan created afterwards live in that buffer
and are reduced by the library's own multimem kernel (csrc/comm.cu) instead of ncclAllReduce.

torch's allocator identifies peers by CUDA device index, so all GPUs must be visible in every process and rank r works on
device LOCAL_RANK (with one visible device per process every rank would claim index 0 and the rendezvous refuses)."""
import ctypes as C


def attach(pool_bytes, device, group=None):
    """Returns an object that keeps the memory alive (drop it only after the last plan is gone and after detach()), or
    None when symmetric memory / multicast is not available here -- the library then keeps using NCCL."""
    import torch
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm_mem

    from . import check, lib

    group = group or dist.group.WORLD
    failure = None
    try:
        n = (int(pool_bytes) + 1023) // 1024 * 256
        t = symm_mem.empty(n, dtype=torch.float32, device=device)
        hdl = symm_mem.rendezvous(t, group)
        off = int(getattr(hdl, "offset", 0) or 0)          # the tensor's offset inside the allocation the handle describes
        mc = int(hdl.multicast_ptr)
        ptrs = [int(p) + off for p in hdl.buffer_ptrs]
        pads = [int(p) for p in hdl.signal_pad_ptrs]
        if mc == 0 or len(pads) != hdl.world_size:
            raise RuntimeError("no multicast mapping")
        mc += off
        if ptrs[hdl.rank] != t.data_ptr():
            raise RuntimeError("symmetric handle does not describe the tensor (%#x vs %#x)" % (ptrs[hdl.rank], t.data_ptr()))
        t.zero_()
        torch.cuda.synchronize()
    except Exception as e:   # no NVSwitch multicast, no fd passing in this container, one visible device per process ...
        failure = e
    # Every rank must take the same exchange: a rank that fell back to NCCL alone would wait in ncclAllReduce for peers that
    # sit in the multicast kernel's flag barrier.  One collective vote (it doubles as the barrier behind the zero fill).
    try:
        agreed = all_agree(failure is None, group, device)
    except Exception as e:   # the vote itself could not run: decide locally, like a single rank would
        import sys
        sys.stderr.write("dopt_b200.symm: exchange vote failed (%r)\n" % (e,))
        agreed = failure is None
        try:
            dist.barrier(group)
        except Exception:
            pass
    if not agreed:
        import sys
        sys.stderr.write("dopt_b200.symm: symmetric memory not available on every rank (here: %r); using NCCL\n" % (failure,))
        return None
    arr = (C.c_void_p * len(pads))(*pads)
    check(lib.dopt_b200_comm_set_symmetric(C.c_void_p(ptrs[hdl.rank]), C.c_void_p(mc), C.c_size_t(n * 4), arr, len(pads),
                                           C.c_size_t(int(hdl.signal_pad_size))))
    return (t, hdl)


def all_agree(ok, group=None, device=None):
    """True when `ok` holds on every rank of the group (one MIN all-reduce of a flag; a collective: every rank calls it)."""
    import torch
    import torch.distributed as dist

    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=device if device is not None else "cpu")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group or dist.group.WORLD)
    return bool(int(flag.item()))


def detach():
    from . import lib
    lib.dopt_b200_comm_set_symmetric(None, None, 0, None, 0, 0)
