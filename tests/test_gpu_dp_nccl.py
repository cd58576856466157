"""GPU, >= 2 devices: data-parallel parity of the real CUDA path over NCCL -- see tests/dp_worker.py.
Skipped on a single-GPU box (the gloo test tests/test_dist_cpu.py covers the host logic there); run it with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_dp_nccl.py -q -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("mode", ["recurrent", "feed_forward"])
def test_two_rank_nccl_update_matches_single_process(mode):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dp_worker.py"), mode]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
    assert "dp_nccl_parity ok" in r.stdout
