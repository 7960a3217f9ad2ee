import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _cuda_device_present() -> bool:
    """Is there a GPU in this machine at all?  Asked WITHOUT the product library: on a GPU box a missing or broken
    libqsv_b200.so must make the gpu tests fail loudly, not skip."""
    import glob

    if glob.glob("/dev/nvidia[0-9]*"):
        return True
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without a usable CUDA device, so plain `pytest tests` works
    everywhere; on the B200 box they run, and the CUDA path fails loudly if its extension is missing."""
    if not any("gpu" in item.keywords for item in items) or _cuda_device_present():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this machine")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


def c_arr(pairs, dtype=np.complex128):
    a = np.asarray(pairs, dtype=np.float64)
    return (a[..., 0] + 1j * a[..., 1]).astype(dtype)


def random_state(n, seed, dtype=np.complex128):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    v /= np.linalg.norm(v)
    return v.astype(dtype)


def obs_from_json(o):
    """JSON list -> oracle observable tuple."""
    kind = o[0]
    if kind == "Named":
        return ("Named", o[1], list(o[2]))
    if kind == "Hermitian":
        return ("Hermitian", c_arr(o[1]), list(o[2]))
    if kind == "TensorProd":
        return ("TensorProd", [obs_from_json(x) for x in o[1]])
    if kind == "Hamiltonian":
        return ("Hamiltonian", list(o[1]), [obs_from_json(x) for x in o[2]])
    raise ValueError(kind)
