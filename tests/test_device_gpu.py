"""LightningGPU device mirror (pennylane_lightning_gpu_b200/lightning_gpu.py) against the oracle: the same
checks the reference's Python suite makes (tests/test_apply.py, test_expval.py, test_var.py, test_probs.py,
test_sample.py, test_adjoint_jacobian.py), on plain operation records because PennyLane is not installed."""
import math

import numpy as np
import pytest

from conftest import random_state
from oracle import np_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev_mod():
    from pennylane_lightning_gpu_b200 import lightning_gpu as lg

    return lg


def _ops_and_dicts(lg, n, seed):
    rng = np.random.default_rng(seed)
    dicts = []
    for w in range(n):
        dicts.append({"name": "Hadamard", "wires": [w], "params": []})
    for name in ("RX", "RY", "RZ", "CNOT", "CRY", "IsingXX", "Rot", "PhaseShift", "Toffoli", "SingleExcitation", "CZ"):
        nw, npar = orc.GATE_ARITY[name]
        dicts.append({"name": name, "wires": [int(x) for x in rng.choice(n, nw, replace=False)],
                      "params": [float(x) for x in rng.uniform(-2, 2, npar)]})
    u = np.linalg.qr(rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4)))[0]
    dicts.append({"name": "QubitUnitary", "wires": [1, 3], "params": [], "matrix": u})
    ops = [lg.Op(d["name"], d["wires"], d["params"], False, d.get("matrix")) for d in dicts]
    return ops, dicts


@pytest.mark.parametrize("c_dtype,tol", [(np.complex128, 1e-10), (np.complex64, 1e-5)])
def test_apply_state_expval_var_probs(dev_mod, c_dtype, tol):
    lg = dev_mod
    n = 5
    dev = lg.LightningGPU(n, c_dtype=c_dtype)
    ops, dicts = _ops_and_dicts(lg, n, 1)
    dev.apply(ops)
    psi = orc.apply_ops(orc.basis_state(n), dicts)
    assert np.max(np.abs(dev.state - psi)) < tol
    for name in ("PauliX", "PauliY", "PauliZ", "Hadamard"):
        o = lg.Obs(name, [2])
        e = orc.expval_named(psi, name, [2])
        assert abs(dev.expval(o) - e) < tol
        assert abs(dev.var(o) - (1 - e * e)) < tol * 10
    t = lg.Obs("Tensor", terms=[lg.Obs("PauliX", [0]), lg.Obs("PauliZ", [3])])
    want = orc.expval_obs(psi, ("TensorProd", [("Named", "PauliX", [0]), ("Named", "PauliZ", [3])]))
    assert abs(dev.expval(t) - want) < tol
    ham = lg.Obs("Hamiltonian", coeffs=[0.4, -1.3], terms=[t, lg.Obs("PauliY", [4])])
    want_h = 0.4 * want - 1.3 * orc.expval_named(psi, "PauliY", [4])
    assert abs(dev.expval(ham) - want_h) < tol * 10
    h = np.array([[0.2, 1 - 0.5j], [1 + 0.5j, -0.9]])
    ho = lg.Obs("Hermitian", [1], matrix=h)
    assert abs(dev.expval(ho) - orc.expval_matrix(psi, h, [1]).real) < tol * 10
    assert abs(dev.var(ho) - (orc.expval_matrix(psi, h @ h, [1]).real - orc.expval_matrix(psi, h, [1]).real ** 2)) < tol * 50
    # tensor product with non-Pauli factors on unordered wires, and variances of composite observables: through the
    # observable's dense matrix (lightning_gpu.py:884-897, 936-960)
    t2 = lg.Obs("Tensor", terms=[lg.Obs("Hadamard", [3]), ho, lg.Obs("PauliY", [0])])
    w2, m2 = lg.LightningGPU._matrix_of(t2)
    e2 = orc.expval_matrix(psi, m2, w2).real
    assert abs(dev.expval(t2) - e2) < tol * 10
    assert abs(dev.var(t2) - (orc.expval_matrix(psi, m2 @ m2, w2).real - e2 ** 2)) < tol * 50
    wh, mh = lg.LightningGPU._matrix_of(ham)
    assert abs(dev.var(ham) - (orc.expval_matrix(psi, mh @ mh, wh).real - want_h ** 2)) < tol * 50
    for wires in ([0], [1, 3], [0, 2, 4], list(range(n))):
        assert np.max(np.abs(dev.probability(wires) - orc.probs(psi.astype(c_dtype), wires))) < tol
    with pytest.raises(RuntimeError, match="out-of-order"):
        dev.probability([3, 1])
    # state preparation on a subset of wires and basis states
    sub = random_state(2, 5)
    dev.reset()
    dev.apply([lg.Op("StatePrep", [3, 1], [sub])])
    full = np.zeros(1 << n, dtype=complex)
    for k in range(4):
        full[((k >> 1) & 1) << (n - 1 - 3) | (k & 1) << (n - 1 - 1)] = sub[k]
    assert np.max(np.abs(dev.state - full)) < tol
    dev.reset()
    dev.apply([lg.Op("BasisState", [0, 4], [[1, 1]])])
    assert abs(dev.state[(1 << (n - 1)) | 1] - 1) < tol
    with pytest.raises(ValueError, match="cannot be used after other Operations"):
        dev.apply([lg.Op("PauliX", [0]), lg.Op("BasisState", [0], [[1]])])
    with pytest.raises(TypeError):
        lg.LightningGPU(2, c_dtype=np.float64)


def test_samples_and_shot_statistics(dev_mod):
    lg = dev_mod
    dev = lg.LightningGPU(3, shots=20000, seed=7)
    dev.apply([lg.Op("RY", [0], [1.1]), lg.Op("CNOT", [0, 2]), lg.Op("Hadamard", [1])])
    s = dev.generate_samples()
    assert s.shape == (20000, 3) and set(np.unique(s)) <= {0, 1}
    assert np.all(s[:, 0] == s[:, 2])
    psi = orc.apply_ops(orc.basis_state(3), [{"name": "RY", "wires": [0], "params": [1.1]},
                                               {"name": "CNOT", "wires": [0, 2]}, {"name": "Hadamard", "wires": [1]}])
    assert abs(dev.expval(lg.Obs("PauliZ", [0])) - orc.expval_named(psi, "PauliZ", [0])) < 0.03
    assert abs(dev.expval(lg.Obs("PauliX", [1])) - 1.0) < 1e-12  # |+> on wire 1
    assert abs(dev.var(lg.Obs("PauliZ", [0])) - (1 - orc.expval_named(psi, "PauliZ", [0]) ** 2)) < 0.03
    assert np.max(np.abs(dev.probability([0, 1]) - orc.probs(psi, [0, 1]))) < 0.02
    ev = dev.sample(lg.Obs("PauliZ", [0]))
    assert ev.shape == (20000,) and set(np.unique(ev)) <= {-1.0, 1.0}
    counts = dev.sample(lg.Obs("PauliX", [1]), counts=True)
    assert counts == {1.0: 20000}  # |+> on wire 1: every shot gives +1


@pytest.mark.parametrize("c_dtype,tol", [(np.complex128, 1e-10), (np.complex64, 3e-4)])
def test_adjoint_jacobian_and_vjp(dev_mod, c_dtype, tol):
    lg = dev_mod
    n = 5
    dev = lg.LightningGPU(n, c_dtype=c_dtype)
    ops, dicts = _ops_and_dicts(lg, n, 2)
    # the serializer expands Rot into RZ RY RZ; do the same for the oracle
    exp = []
    for d in dicts:
        if d["name"] == "Rot":
            p = d["params"]
            exp += [{"name": "RZ", "wires": d["wires"], "params": [p[0]]}, {"name": "RY", "wires": d["wires"], "params": [p[1]]},
                    {"name": "RZ", "wires": d["wires"], "params": [p[2]]}]
        else:
            exp.append(d)
    psi = orc.apply_ops(orc.basis_state(n), exp)
    obs = [lg.Obs("PauliZ", [0]), lg.Obs("Tensor", terms=[lg.Obs("PauliX", [1]), lg.Obs("PauliY", [2])]),
           lg.Obs("Hamiltonian", coeffs=[0.5, 2.0], terms=[lg.Obs("PauliZ", [3]), lg.Obs("PauliX", [4])])]
    obs_t = [("Named", "PauliZ", [0]), ("TensorProd", [("Named", "PauliX", [1]), ("Named", "PauliY", [2])]),
             ("Hamiltonian", [0.5, 2.0], [("Named", "PauliZ", [3]), ("Named", "PauliX", [4])])]
    n_par = sum(1 for d in exp if d["params"])
    jac = dev.adjoint_jacobian(ops, obs)
    want = orc.adjoint_jacobian(psi, exp, obs_t, list(range(n_par)))
    assert jac.shape == want.shape and np.max(np.abs(jac - want)) < tol
    sub = dev.adjoint_jacobian(ops, obs, trainable_params=[1, 4])
    assert np.max(np.abs(sub - want[:, [1, 4]])) < tol
    dy = np.array([0.3, -1.0, 0.25])
    assert np.max(np.abs(dev.vjp(ops, obs, dy) - dy @ want)) < tol * 10
    assert dev.adjoint_jacobian(ops, []).size == 0
