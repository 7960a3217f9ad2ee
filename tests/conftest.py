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
