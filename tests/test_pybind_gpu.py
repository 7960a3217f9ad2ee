"""The drop-in surface: pybind11 module `lightning_gpu_qubit_ops` with the reference's class and method
names (bindings/Bindings.cpp), driven with exactly the calls lightning_gpu.py makes
(lightning_gpu.py:519-555 apply_cq, :820-897 expval, :899-926 probability, :638-752 adjoint_jacobian),
checked against the oracle.  GPU only."""
import math
import os
import sys

import numpy as np
import pytest

from conftest import random_state
from oracle import np_oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ops_mod():
    from pennylane_lightning_gpu_b200 import lightning_gpu_qubit_ops as m  # one module object per process

    return m


PREC = [("128", np.complex128, 1e-10), ("64", np.complex64, 1e-5)]


@pytest.mark.parametrize("bits,dtype,tol", PREC)
def test_state_io_and_named_gate_methods(ops_mod, bits, dtype, tol):
    SV = getattr(ops_mod, "LightningGPU_C" + bits)
    n = 6
    sv = SV(n)
    assert sv.numQubits() == n and sv.dataLength() == 1 << n and sv.GetNumGPUs() >= 1
    out = np.zeros(1 << n, dtype=dtype)
    sv.DeviceToHost(out, False)
    assert np.array_equal(out, orc.basis_state(n, 0, dtype))
    sv.setBasisState(5, False)
    sv.DeviceToHost(out, False)
    assert np.array_equal(out, orc.basis_state(n, 5, dtype))
    idt = np.int64 if bits == "128" else np.int32
    sv.setStateVector(np.array([1, 7], dtype=idt), np.array([0.6, 0.8j], dtype=dtype), False)
    sv.DeviceToHost(out, False)
    assert np.allclose(out, orc.set_state_vector(n, [1, 7], np.array([0.6, 0.8j]), dtype))
    psi = random_state(n, 3)
    rng = np.random.default_rng(4)
    for name, (nw, npar) in orc.GATE_ARITY.items():
        nw = nw if nw is not None else 3
        wires = [int(w) for w in rng.choice(n, size=nw, replace=False)]
        params = [float(x) for x in rng.uniform(-2, 2, npar)]
        for adj in (False, True):
            sv.HostToDevice(psi.astype(dtype), False)
            getattr(sv, name)(wires, adj, params)          # method = getattr(self._gpu_state, name)
            sv.DeviceToHost(out, False)
            assert np.max(np.abs(out - orc.apply_op(psi, name, wires, params, adj))) < tol, (name, adj)
    # matrix path: apply(name, wires, adjoint, params, matrix) for ops without a kernel
    u = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    sv.HostToDevice(psi.astype(dtype), False)
    sv.apply("QubitUnitary", [4, 1], False, [], u.astype(dtype).reshape(-1))
    sv.DeviceToHost(out, False)
    assert np.max(np.abs(out - orc.apply_matrix(psi, u, [4, 1]))) < tol * 10
    # vector overloads
    sv.HostToDevice(psi.astype(dtype), False)
    sv.apply(["RX", "CNOT", "Hadamard"], [[0], [0, 3], [5]], [False, False, True], [[0.3], [], []])
    sv.DeviceToHost(out, False)
    want = orc.apply_ops(psi, [{"name": "RX", "wires": [0], "params": [0.3]}, {"name": "CNOT", "wires": [0, 3]},
                               {"name": "Hadamard", "wires": [5]}])
    assert np.max(np.abs(out - want)) < tol
    with pytest.raises(ops_mod.PLException, match="Currently unsupported gate"):
        sv.apply("NotAGate", [0], False, [], np.zeros(0, dtype=dtype))
    copy = SV(sv)
    other = np.zeros(1 << n, dtype=dtype)
    copy.DeviceToHost(other, False)
    assert np.array_equal(other, out)
    from_np = SV(psi.astype(dtype))
    from_np.DeviceToHost(other, False)
    assert np.allclose(other, psi.astype(dtype))
    copy.DeviceToDevice(from_np, False)
    copy.DeviceToHost(out, False)
    assert np.array_equal(out, other)
    copy.resetGPU(False)
    copy.DeviceToHost(out, False)
    assert np.array_equal(out, orc.basis_state(n, 0, dtype))


@pytest.mark.parametrize("bits,dtype,tol", PREC)
def test_measurement_overloads(ops_mod, bits, dtype, tol):
    SV = getattr(ops_mod, "LightningGPU_C" + bits)
    n = 7
    psi = random_state(n, 9)
    sv = SV(psi.astype(dtype))
    empty = np.zeros(0, dtype=dtype)
    for name in ("PauliX", "PauliY", "PauliZ", "Hadamard"):
        got = sv.ExpectationValue(name, [2], [], empty)
        assert abs(got - orc.expval_named(psi, name, [2])) < tol
    rng = np.random.default_rng(1)
    h = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    h = h + h.conj().T
    assert abs(sv.ExpectationValue([1, 4], h.astype(dtype).reshape(-1)) - orc.expval_matrix(psi, h, [1, 4]).real) < tol * 10
    assert abs(sv.ExpectationValue("Hermitian", [1, 4], [], h.astype(dtype).reshape(-1))
               - orc.expval_matrix(psi, h, [1, 4]).real) < tol * 10
    words, wires, coeffs = ["XZ", "Y", "ZZI"], [[0, 3], [6], [1, 2, 5]], [0.5, -0.25, 2.0]
    got = sv.ExpectationValue(words, wires, np.array(coeffs, dtype=dtype))
    assert abs(got - orc.expval_pauli_words(psi.astype(dtype), words, wires, coeffs)) < tol * 10
    import scipy.sparse as sp

    m = sp.random(1 << n, 1 << n, density=0.02, random_state=3, dtype=np.float64).tocsr()
    m = (m + m.T).tocsr().astype(np.complex128)
    m.sort_indices()
    idt = np.int64 if bits == "128" else np.int32
    got = sv.ExpectationValue(m.indptr.astype(idt), m.indices.astype(idt), m.data.astype(dtype))
    assert abs(got - orc.expval_csr(psi, m.indptr, m.indices, m.data)) < tol * 100
    p = sv.Probability([1, 3, 6])
    assert np.max(np.abs(p - orc.probs_custatevec_order(psi.astype(dtype), [1, 3, 6]))) < tol
    # the device re-orders exactly like lightning_gpu.py:920-924
    pl = p.reshape([2] * 3).transpose().reshape(-1)
    assert np.max(np.abs(pl - orc.probs(psi.astype(dtype), [1, 3, 6]))) < tol
    s = sv.GenerateSamples(n, 2000)
    assert s.shape == (2000, n) and set(np.unique(s)) <= {0, 1}
    freq = s[:, 0].mean()
    p1 = orc.probs(psi, [0])[1]
    assert abs(freq - p1) < 0.05
    assert np.array_equal(sv.GenerateSamples(n, 100, 42), sv.GenerateSamples(n, 100, 42))


@pytest.mark.parametrize("bits,dtype,tol", PREC)
def test_adjoint_through_bindings(ops_mod, bits, dtype, tol):
    SV = getattr(ops_mod, "LightningGPU_C" + bits)
    Named = getattr(ops_mod, "NamedObsGPU_C" + bits)
    Tensor = getattr(ops_mod, "TensorProdObsGPU_C" + bits)
    Ham = getattr(ops_mod, "HamiltonianGPU_C" + bits)
    Herm = getattr(ops_mod, "HermitianObsGPU_C" + bits)
    Sparse = getattr(ops_mod, "SparseHamiltonianGPU_C" + bits)
    Adj = getattr(ops_mod, "AdjointJacobianGPU_C" + bits)
    from pennylane_lightning_gpu_b200 import workloads

    n = 6
    ops, n_par = workloads.hardware_efficient_ansatz(n, layers=2, seed=3)
    ops.insert(4, {"name": "IsingXX", "wires": [0, 3], "params": [0.4]})
    ops.insert(9, {"name": "CRZ", "wires": [2, 5], "params": [-0.9]})
    n_par += 2
    psi = orc.apply_ops(orc.basis_state(n), ops)
    rdt = np.float64 if bits == "128" else np.float32
    names = [o["name"] for o in ops]
    params = [np.array(o["params"], dtype=rdt) for o in ops]
    wires = [o["wires"] for o in ops]
    invs = [False] * len(ops)
    mats = [np.zeros(0, dtype=dtype) for _ in ops]
    adj = Adj()
    rec = adj.create_ops_list(names, params, wires, invs, mats)
    assert "RY" in repr(rec)
    h1 = np.array([[0.3, 0.1 - 0.2j], [0.1 + 0.2j, -0.7]])
    obs_t = [("Named", "PauliZ", [0]),
             ("TensorProd", [("Named", "PauliX", [1]), ("Named", "PauliY", [4])]),
             ("Hamiltonian", [0.7, -0.2], [("Named", "PauliZ", [2]), ("TensorProd", [("Named", "PauliX", [0]), ("Named", "PauliZ", [5])])]),
             ("Hermitian", h1, [3])]
    obs = [Named("PauliZ", [0]), Tensor([Named("PauliX", [1]), Named("PauliY", [4])]),
           Ham(np.array([0.7, -0.2], dtype=rdt), [Named("PauliZ", [2]), Tensor([Named("PauliX", [0]), Named("PauliZ", [5])])]),
           Herm(h1.astype(dtype).reshape(-1), [3])]
    sv = SV(psi.astype(dtype))
    tp = list(range(n_par))
    jac = adj.adjoint_jacobian(sv, obs, rec, tp)
    want = orc.adjoint_jacobian(psi, ops, obs_t, tp)
    assert jac.shape == want.shape
    assert np.max(np.abs(jac - want)) < (1e-10 if bits == "128" else 2e-4)
    assert np.allclose(adj.adjoint_jacobian_batched(sv, obs, rec, tp), jac)
    with pytest.raises(ops_mod.PLException, match="No trainable parameters provided"):
        adj.adjoint_jacobian(sv, obs, rec, [])
    assert obs[0] == Named("PauliZ", [0]) and not (obs[0] == Named("PauliZ", [1]))
    assert obs[1].get_wires() == [1, 4]
    # sparse Hamiltonian observable in the adjoint
    import scipy.sparse as sp

    m, (w2, ws2, c2) = workloads.molecular_style_sparse_hamiltonian(n, 20, 4, 3)
    idt = np.int64 if bits == "128" else np.int32
    so = Sparse(m.data.astype(dtype), m.indices.astype(idt), m.indptr.astype(idt), list(range(n)))
    j2 = adj.adjoint_jacobian(sv, [so], rec, tp)
    w2j = orc.adjoint_jacobian(psi, ops, [("Sparse", m.indptr, m.indices, m.data)], tp)
    assert np.max(np.abs(j2 - w2j)) < (1e-9 if bits == "128" else 5e-4)


def test_module_functions(ops_mod):
    assert ops_mod.is_gpu_supported(0)
    assert ops_mod.get_gpu_arch(0)[0] == 10
    tag = ops_mod.DevTag(0)
    assert tag.getDeviceID() == 0
    pool = ops_mod.DevPool()
    assert pool.getTotalDevices() >= 1
    d = pool.acquireDevice()
    assert pool.isActive(d)
    pool.releaseDevice(d)
    assert pool.isInactive(d)
    sv = ops_mod.LightningGPU_C128(3, tag)
    assert sv.getCurrentGPU() == 0
