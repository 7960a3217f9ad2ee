"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads without a GPU and
exports every symbol include/qsv_b200.h declares; compute calls fail loudly (no fallback)."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def qsv():
    from pennylane_lightning_gpu_b200 import _build, _cabi

    _build.build_lib()
    return _cabi


def declared_symbols():
    with open(os.path.join(ROOT, "include", "qsv_b200.h")) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qsv_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported_and_bound(qsv):
    names = declared_symbols()
    assert len(names) >= 45
    lib = qsv.lib()
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/qsv_b200.h but not exported"
        assert name in qsv.SIGNATURES, f"{name} has no ctypes signature"
    assert set(qsv.SIGNATURES) == set(names)
    assert lib.qsv_version() >= 100


def test_library_is_sm100a_only(qsv):
    import subprocess

    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", qsv.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_gpu(qsv):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(qsv.QsvError):
        qsv.StateVector(3, np.complex128)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "pennylane_lightning_gpu_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                with open(os.path.join(d, f)) as fh:
                    src = fh.read()
                assert "np_oracle" not in src and "from oracle" not in src and "import oracle" not in src, f


def test_ops_and_observable_records_build_on_cpu(qsv):
    ops = qsv.Ops([{"name": "RX", "wires": [0], "params": [0.1]}, {"name": "CNOT", "wires": [0, 1]},
                   {"name": "QubitUnitary", "wires": [1], "matrix": np.eye(2)}])
    assert len(ops) == 3
    o = qsv.Observable.from_tuple(("Hamiltonian", [0.5, 1.0], [("Named", "PauliZ", [0]),
                                                                ("TensorProd", [("Named", "PauliX", [0]),
                                                                                ("Hermitian", np.eye(2), [1])])]))
    assert o._h


def test_workload_generators():
    from pennylane_lightning_gpu_b200 import workloads as w

    ops, n_par = w.strongly_entangling_layers(20, 2, 1337)
    assert len(ops) == 160 and n_par == 120 and ops[60]["wires"] == [0, 1] and ops[159]["wires"] == [19, 1]
    ops = w.random_gate_circuit(30, 200, 2024)
    assert len(ops) == 200 and all(max(o["wires"]) < 30 for o in ops)
    assert w.gate_bytes({"name": "RX"}, 30, 16) == 2 * 16 * 2**30
    assert w.gate_bytes({"name": "CNOT"}, 30, 16) == 16 * 2**30
    ops, n_par = w.hardware_efficient_ansatz(24, 4, 11)
    assert n_par == 192 and len(ops) == 4 * (48 + 23)
    words, wires, coeffs = w.random_pauli_hamiltonian(24, 100, 5)
    assert len(words) == 100 and all(len(a) == len(b) for a, b in zip(words, wires))
    m, (w2, ws2, c2) = w.molecular_style_sparse_hamiltonian(8, 40, 5, 3)
    from oracle import np_oracle as orc
    psi = np.random.default_rng(0).normal(size=256) + 1j * np.random.default_rng(1).normal(size=256)
    psi /= np.linalg.norm(psi)
    assert abs(orc.expval_csr(psi, m.indptr, m.indices, m.data) - orc.expval_pauli_words(psi, w2, ws2, c2)) < 1e-12


def test_pybind_module_has_the_reference_names():
    """Python-visible names of bindings/Bindings.cpp:67-877, 890-1720, 1727-1777 of the reference (import only)."""
    from pennylane_lightning_gpu_b200 import lightning_gpu_qubit_ops as m

    names = ["DevPool", "DevTag", "PLException", "MPIManager", "device_reset", "allToAllAccess", "is_gpu_supported",
             "get_gpu_arch"]
    for bits in ("64", "128"):
        for mpi in ("", "MPI"):
            names += [f"LightningGPU{mpi}_C{bits}", f"NamedObsGPU{mpi}_C{bits}", f"HermitianObsGPU{mpi}_C{bits}",
                      f"TensorProdObsGPU{mpi}_C{bits}", f"HamiltonianGPU{mpi}_C{bits}", f"SparseHamiltonianGPU{mpi}_C{bits}",
                      f"OpsStructGPU{mpi}_C{bits}", f"AdjointJacobianGPU{mpi}_C{bits}"]
    missing = [n for n in names if not hasattr(m, n)]
    assert not missing, missing
    for cls in (m.LightningGPU_C128, m.LightningGPUMPI_C128):
        for meth in ("setBasisState", "setStateVector", "apply", "ExpectationValue", "Probability", "GenerateSamples",
                     "DeviceToHost", "HostToDevice", "DeviceToDevice", "RX", "CNOT", "DoubleExcitationPlus", "MultiRZ"):
            assert hasattr(cls, meth), (cls, meth)
    for meth in ("numLocalQubits", "numGlobalQubits", "dataLength", "resetGPU"):
        assert hasattr(m.LightningGPUMPI_C128, meth)
    for meth in ("Barrier", "getRank", "getSize", "getSizeNode", "getTime", "getVendor", "getVersion", "Scatter"):
        assert hasattr(m.MPIManager, meth)
    assert hasattr(m.AdjointJacobianGPUMPI_C128, "adjoint_jacobian_serial")
