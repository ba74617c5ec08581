"""N > 1: one process per rank.  The host half (halo index exchange, start-up sync, per-rank
preprocessing) runs here on CPU over gloo with world_size 2; the full step with the NCCL halo
exchange runs on the GPU box when it has >= 2 GPUs.

Note on rand(): the reference seeds every MPI rank's libc separately, so a multi-rank run draws
the border bases of each rank from its own stream.  The golden fixtures were produced by the
reference on ONE rank; they match a 2-rank run for flags/faces always, and for values only on
inner zones' nodes away from the border -- so the GPU test compares the multi-rank run with a
single-process run of the product (bit-exact halo path) and the host test with the fixtures."""
import os
import subprocess
import sys

import numpy as np
import pytest

import util_parity as up

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _torchrun(nproc, args, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "mp_worker.py")] + args
    return subprocess.run(cmd, capture_output=True, text=True, timeout=600)


@pytest.mark.parametrize("name", ["zones2", "zones3_jitter"])
def test_two_ranks_host_half_gloo(built, name):
    r = _torchrun(2, ["host", name], 29531)
    assert r.returncode == 0 and "MP_RESULT OK" in r.stdout, (r.stdout[-1500:], r.stderr[-6000:])


@pytest.mark.gpu
def test_two_ranks_host_driven_stages(built):
    """The stage-by-stage path a host drives itself (DataBus.sync_nodes through torch.distributed P2P ops +
    gcm_b200_halo_pack / unpack, then gcm_b200_set_do_next_part_step): no overlapped stage, the shell nodes of
    the pattern kernel run right behind the range launch.  Same 3-step comparison with the single-process run."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    old = os.environ.get("GCM_B200_TORCH_HALO")
    os.environ["GCM_B200_TORCH_HALO"] = "1"
    try:
        r = _torchrun(2, ["gpu", "mp24"], 29536)
    finally:
        if old is None:
            os.environ.pop("GCM_B200_TORCH_HALO", None)
        else:
            os.environ["GCM_B200_TORCH_HALO"] = old
    assert r.returncode == 0 and "MP_RESULT OK" in r.stdout, (r.stdout[-1500:], r.stderr[-6000:])


@pytest.mark.gpu
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_ranks_nccl_halo_exchange(built, p2p):
    """Cross-process halo on the device against a single-process run, three steps: direct peer stores into the
    neighbour's halo slots (p2p = 1, the default) and the pack -> ncclSend/ncclRecv -> unpack fallback (p2p = 0)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    old = os.environ.get("GCM_B200_P2P")
    os.environ["GCM_B200_P2P"] = p2p
    try:
        r = _torchrun(2, ["gpu", "mp24"], 29532 + int(p2p))
    finally:
        if old is None:
            os.environ.pop("GCM_B200_P2P", None)
        else:
            os.environ["GCM_B200_P2P"] = old
    assert r.returncode == 0 and "MP_RESULT OK" in r.stdout, (r.stdout[-1500:], r.stderr[-6000:])
