"""Launches tests/dist_check.py under torchrun on every power-of-two GPU count the box offers (>= 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_gpus():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_register_matches_oracle(world):
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + world), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"dist_check_{world}gpu.log"), "w") as f:  # evidence / post-mortem
        f.write(r.stdout[-200000:] + "\n==== stderr ====\n" + r.stderr[-20000:])
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world", [2, 4])
def test_fused_exchange_matches_oracle(world):
    """Exchanges carried by sweeps (the default since round 2; QSV_DIST_FUSED_SWAP=0 switches it off): exchanges through a second
    buffer, stored by the sweep before them straight into the partner's memory."""
    if _n_gpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29700 + world), os.path.join(ROOT, "tests", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=dict(os.environ, DIST_CHECK_FUSED="1"))
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, f"dist_check_fused_{world}gpu.log"), "w") as f:
        f.write(r.stdout[-200000:] + "\n==== stderr ====\n" + r.stderr[-20000:])
    assert r.returncode == 0 and "DIST_CHECK PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
