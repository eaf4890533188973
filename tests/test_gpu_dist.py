"""GPU: the sub-frame-sharded blurry view (deblurgs_b200.dist.render_blurry_sharded + NCCL all-reduce of the blurred
image, the gradients and the densification statistics) against the unsharded render, launched under torchrun with
one rank per GPU (2 ranks when the box has two GPUs; a single rank otherwise, which still runs the sharded code
path end to end).  The check itself lives in tests/dist_gpu_check.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(600)
def test_subframe_sharded_view_equals_unsharded_under_torchrun():
    n = 2 if torch.cuda.device_count() >= 2 else 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
           "--master-addr", "127.0.0.1", "--master-port", "29571", os.path.join(ROOT, "tests", "dist_gpu_check.py"), "small"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "OK" in r.stdout, r.stdout[-2000:]


@pytest.mark.timeout(300)
def test_direct_nccl_collectives_inside_a_cuda_graph():
    """deblurgs_b200.nccl_direct: all-reduces on a side-stream branch of a captured graph (needs two GPUs)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29573", os.path.join(ROOT, "tests", "nccl_direct_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "nccl_direct OK" in r.stdout, r.stdout[-2000:]
