"""Sharded path on real GPUs (needs >= 2): molecule-aligned particle ranges, one process per GPU, the only collective
is the exchange of the kinetic-energy vector (peer-mapped inboxes over NVLink, or an NCCL all-reduce).  The CPU-side logic is covered by
tests/test_host.py::test_sharded_kinetic_energy_reduction_gloo."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_sharded_step_matches_single_gpu(cuda, p2p):
    """p2p=1: kinetic-energy partials through peer-mapped inboxes (the default on an NVLink node); p2p=0: NCCL all-reduce."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs 2 GPUs")
    world = 2
    env = dict(os.environ, TGNH_P2P=p2p)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29541" if p2p == "1" else "29542", os.path.join(ROOT, "tests", "shard_worker.py")],
                         capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and f"SHARD_OK {world}" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    if p2p == "1":
        print(out.stdout[-200:])
