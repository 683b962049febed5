"""Data-parallel training through the real NCCL gradient buckets: tools/dp_check.py under torchrun with two ranks (one GPU
each).  Phase 1: a small convolutional net in fp32, four SGD steps -- replicas bit-identical and equal (1e-4) to the
single-process CPU oracle on the concatenated batch (the sharding rule of SURVEY.md section 8(e)).  Phase 2: WRN-10-4 with
batch norm, bf16 tensor cores and bf16 interior activations -- every trainable parameter AND the batch-norm running
statistics bit-identical across ranks, loss finite and falling.  Skips itself on a box with fewer than two GPUs; the log
of the last run is kept in gpurun_out/dp_check.log."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["nccl", "multicast"])
def test_two_rank_nccl_data_parallel_matches_oracle_and_replicas_agree(exchange):
    """exchange = "multicast": the gradient buckets live in peer-mapped memory (dopt_b200_comm_set_symmetric) and are reduced by
    the library's own multimem.ld_reduce / multimem.st kernel between two flag barriers instead of ncclAllReduce; the same
    assertions must hold (the sum is formed once, inside the switch, so replicas stay bit-identical)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    port = 29600 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "dp_check.py")]
    env = dict(os.environ)
    env.pop("CUDA_VISIBLE_DEVICES", None)
    env["DOPT_B200_SYMM"] = "1" if exchange == "multicast" else "0"
    r = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600)
    out = r.stdout.decode("utf-8", "replace")
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "dp_check_%s.log" % exchange), "w") as f:
        f.write(out)
    assert r.returncode == 0, out[-3000:]
    assert "-> PASS" in out and "exchange=%s" % exchange in out, out[-3000:]
