"""-m gpu, needs >= 2 GPUs: the one-process-per-GPU path with real NCCL exchanges against the oracle.
(The driver's 1-GPU box skips it; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`.)"""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


MODES = {
    "default": {},                                           # fused peer stores: bulk stores, two streams, dependent launches
    "direct_stores_one_stream": {"OFFTB_BULK": "0", "OFFTB_PDL": "0", "OFFTB_OVERLAP": "0"},
    "nccl_send_recv": {"OFFTB_EXCHANGE": "nccl"},            # the fallback when peer mapping is unavailable
}


@pytest.mark.parametrize("mode", list(MODES))
def test_nccl_ranks_match_oracle(mode):
    import os
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    n = 8 if n >= 8 else 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29631", str(ROOT / "tests" / "mgpu_worker.py")]
    env = dict(os.environ, OFFTB_FLAG_TIMEOUT_S="30", **MODES[mode])
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    print(res.stdout[-4000:], res.stderr[-2000:])
    assert res.returncode == 0 and "MGPU PARITY PASSED" in res.stdout


def test_lost_peer_times_out_with_a_message_and_the_context_survives():
    """a flag wait must neither hang nor trap - the execute fails with a message, the next plan works"""
    import os
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs at least 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29641", str(ROOT / "tests" / "mgpu_timeout_worker.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=dict(os.environ, OFFTB_FLAG_TIMEOUT_S="3"))
    print(res.stdout[-2000:], res.stderr[-2000:])
    assert res.returncode == 0 and "MGPU TIMEOUT PASSED" in res.stdout
