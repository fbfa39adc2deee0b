"""N > 1 on hardware: B200DDP + GradSync (NCCL all-reduce, AVG) + the packed all-gather + the W x local-slice InfoNCE
backward + the rank-0 weight broadcast, against the fp32 oracle on the global batch (tests/dist_worker.py).
Needs >= 2 GPUs on the box (`gpurun --gpus 2`); the gloo / CPU tests in test_host_cpu.py cover the host logic."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs >= 2 CUDA devices")
def test_two_rank_nccl_step_matches_global_batch_oracle():
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="4")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert "[nccl] all cases passed" in r.stdout
