"""The C/OpenMP CPU baseline (oracle/lq_port.c) against the NumPy oracle, which is itself pinned by the
reference's golden vectors (tests/test_oracle_golden.py).  CPU only."""
import math

import numpy as np
import pytest

from conftest import random_state
from oracle import lq_port as lq
from oracle import np_oracle as orc
from pennylane_lightning_gpu_b200 import workloads


def test_every_gate_matches_numpy_oracle():
    n = 7
    rng = np.random.default_rng(1)
    psi = random_state(n, 2)
    for name, (nw, npar) in orc.GATE_ARITY.items():
        nw = nw if nw is not None else 3
        for rep in range(3):
            wires = [int(w) for w in rng.choice(n, size=nw, replace=False)]
            params = [float(x) for x in rng.uniform(-3, 3, npar)]
            for adj in (False, True):
                st = lq.LQState(n, psi)
                st.apply_op(name, wires, params, adj)
                assert np.allclose(st.sv, orc.apply_op(psi, name, wires, params, adj), atol=1e-12), (name, wires, adj)
    u = workloads.haar_unitary(rng, 8)
    st = lq.LQState(n, psi)
    st.apply_op("QubitUnitary", [5, 0, 3], matrix=u, adjoint=True)
    assert np.allclose(st.sv, orc.apply_op(psi, "QubitUnitary", [5, 0, 3], adjoint=True, matrix=u), atol=1e-12)


def test_golden_gate_vectors(kats):
    from conftest import c_arr

    for case in kats["gates_py"] + kats["gates_cpp"]:
        psi = c_arr(case["input"])
        n = int(math.log2(psi.size))
        st = lq.LQState(n, psi)
        st.apply_op(case["gate"], case["wires"], case["params"], case.get("adjoint", False))
        assert np.allclose(st.sv, c_arr(case["expected"]), atol=max(case.get("atol", 0), 1e-12)), case.get("cite")


def test_measurements_and_adjoint_match_numpy_oracle():
    n = 8
    ops, n_par = workloads.hardware_efficient_ansatz(n, layers=2, seed=11)
    words, wires, coeffs = workloads.random_pauli_hamiltonian(n, 20, seed=5)
    ham = workloads.hamiltonian_tuple(words, wires, coeffs)
    st = lq.LQState(n)
    st.apply_ops(ops)
    psi = orc.apply_ops(orc.basis_state(n), ops)
    assert np.allclose(st.sv, psi, atol=1e-12)
    assert st.expval_pauli_words(words, wires, coeffs) == pytest.approx(orc.expval_obs(psi, ham), abs=1e-12)
    hp = st.apply_pauli_hamiltonian(words, wires, coeffs)
    assert np.allclose(hp.sv, orc.apply_observable(psi, ham), atol=1e-12)
    obs = [ham, ("Named", "PauliZ", [0]), ("TensorProd", [("Named", "PauliX", [1]), ("Named", "PauliY", [3])])]
    jac = lq.adjoint_jacobian(st, ops, obs, list(range(n_par)))
    assert np.allclose(jac, orc.adjoint_jacobian(psi, ops, obs, list(range(n_par))), atol=1e-12)
    assert lq.num_threads() >= 1
