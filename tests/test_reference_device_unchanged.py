"""The drop-in claim for the Python layer: the reference's own `lightning_gpu.py` and `_serialize.py`,
byte-for-byte UNCHANGED, run against `lightning_gpu_qubit_ops` built from this repository.

The reference's sources are never part of this repository.  They are copied at test time into a scratch package from the
first of: $QSV_REFERENCE_PY, /root/reference/pennylane_lightning_gpu (the build container), oracle/_ref/pennylane_lightning_gpu
(git-ignored; `__graft_entry__.build()` fills it when /root/reference is present so that it travels to the GPU box like a
built .so).  PennyLane is not installable here; tests/stubs/ holds a small stand-in for the subset the two files use, and
an empty `cuquantum` module for the reference's import guard (lightning_gpu.py:93-99).  The check itself runs in a child
process (tests/ref_device_check.py), so the second import of the binary module cannot collide with other tests."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILES = ("__init__.py", "_version.py", "lightning_gpu.py", "_serialize.py")


def _reference_dir():
    for d in (os.environ.get("QSV_REFERENCE_PY"), "/root/reference/pennylane_lightning_gpu",
              os.path.join(ROOT, "oracle", "_ref", "pennylane_lightning_gpu")):
        if d and all(os.path.exists(os.path.join(d, f)) for f in FILES):
            return d
    return None


def _run(tmp_path, mode):
    src = _reference_dir()
    if src is None:
        pytest.skip("the reference's Python sources are not available on this machine")
    from pennylane_lightning_gpu_b200 import _build

    so = _build.pybind_module_path()
    assert os.path.exists(so), "lightning_gpu_qubit_ops is not built"
    pkg = tmp_path / "pennylane_lightning_gpu"
    pkg.mkdir()
    for f in FILES:
        shutil.copyfile(os.path.join(src, f), pkg / f)  # unchanged
    shutil.copyfile(so, pkg / os.path.basename(so))
    os.symlink(os.path.join(os.path.dirname(so), "lib"), pkg / "lib")  # the module's rpath is $ORIGIN/lib
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([str(tmp_path), os.path.join(ROOT, "tests", "stubs")])
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_device_check.py"), mode], env=env,
                       capture_output=True, text=True, timeout=600)
    return r


def test_reference_python_imports_every_binding_name(tmp_path):
    """No GPU needed: the reference's import block resolves all class / function names against this repository's module."""
    if os.path.exists("/dev/nvidia0"):
        pytest.skip("GPU box: covered by the gpu test below")
    r = _run(tmp_path, "import")
    assert r.returncode == 0 and "REF_DEVICE IMPORT PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.gpu
def test_reference_device_runs_unchanged_against_the_oracle(tmp_path):
    r = _run(tmp_path, "gpu")
    out_dir = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "ref_device_check.log"), "w") as f:
        f.write(r.stdout[-100000:] + "\n==== stderr ====\n" + r.stderr[-20000:])
    assert r.returncode == 0 and "REF_DEVICE PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
