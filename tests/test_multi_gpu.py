"""GPU test of the frame-sharded NCCL path (needs >= 2 GPUs on the box; skipped otherwise)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_two_rank_parity():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "mgpu_parity.py")]
    env = dict(os.environ)
    env.pop("MCCBA_P2P", None)
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "MGPU_PARITY_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    # default: the NVLink peer-memory exchange when the windows can be mapped, else NCCL
    assert "exchange=peer" in out.stdout or "exchange=nccl" in out.stdout
    # the same run with the peer-memory exchange switched off: ncclAllReduce
    env["MCCBA_P2P"] = "0"
    cmd[cmd.index("29517")] = "29518"
    out2 = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "MGPU_PARITY_OK" in out2.stdout, out2.stdout[-3000:] + out2.stderr[-3000:]
    assert "exchange=nccl" in out2.stdout
    print(out.stdout[-1500:], out2.stdout[-1500:])
