"""Decomposed device-resident MD over NCCL equals the single-GPU run (needs >= 2 GPUs, else skipped)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world,peer", [(2, 0), (2, 1), (4, 0), (8, 0)])
def test_decomposed_md_equals_single_gpu(world, peer):
    """peer = 1: the same checks with the reverse halo fused into the force kernel (ghost forces added straight to the
    owner rank's accumulators over NVLink, annp_b200_peer_*) instead of the grouped NCCL exchange."""
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world + 20 * peer), os.path.join(ROOT, "scripts", "multi_gpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, ANNP_B200_PEER=str(peer)))
    assert p.returncode == 0 and "MULTI_GPU_CHECK PASS" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
