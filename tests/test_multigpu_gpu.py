"""Frame-sharded U-Net on real GPUs over NCCL (needs >= 2 CUDA devices; skipped on the 1-GPU box).
Launches `torch.distributed.run` on tools/parallel_check.py, which compares the sharded forward with the
single-GPU forward of the same model and with the CPU oracle."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(900)]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,peer", [(2, "1"), (2, "0")])
def test_frame_sharded_unet(world, peer):
    """peer=1: NVLink peer-memory kernels (csrc/peer.cu); peer=0: NCCL collectives."""
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} CUDA devices")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tools", "parallel_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=800, env=dict(os.environ, VC_PEER_COMM=peer))
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "PARALLEL_CHECK_OK" in r.stdout
