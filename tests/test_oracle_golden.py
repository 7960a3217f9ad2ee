"""Pin the NumPy oracle against the reference's own golden vectors (CPU only).

Every expected value in tests/golden/reference_kats.json comes from the reference's
tests (see tests/golden/make_golden.py); nothing is produced by our own code.
"""
import math

import numpy as np
import pytest

from oracle import np_oracle as orc
from conftest import c_arr, obs_from_json


def _close(a, b, case):
    atol = case.get("atol", 0.0)
    rtol = case.get("rtol", 0.0)
    assert np.allclose(a, b, atol=atol, rtol=rtol), (case.get("cite"), a, b)


def test_gates_python_tables(kats):
    """tests/test_apply.py in/out state pairs (1-, 2- and 3-wire gates)."""
    assert len(kats["gates_py"]) >= 70
    for case in kats["gates_py"]:
        out = orc.apply_op(c_arr(case["input"]), case["gate"], case["wires"], case["params"])
        _close(out, c_arr(case["expected"]), case)


def test_gates_cpp_vectors(kats):
    for case in kats["gates_cpp"]:
        out = orc.apply_op(c_arr(case["input"]), case["gate"], case["wires"], case["params"],
                           adjoint=case["adjoint"])
        _close(out, c_arr(case["expected"]), case)


def test_rz_phaseshift_plus_state():
    """Test_StateVectorCudaManaged_Param.cpp:165-290: |+++> then RZ / PhaseShift on wire i puts
    diag element [bit_i] on amplitude k."""
    plus = np.full(8, 1 / (2 * math.sqrt(2)), dtype=np.complex128)
    for gate, angles in (("RZ", [0.2, 0.7, 2.9]), ("PhaseShift", [0.3, 0.8, 2.4])):
        for w, a in enumerate(angles):
            d = np.diag(orc.gate_matrix(gate, [a]))
            exp = np.array([d[(k >> (2 - w)) & 1] for k in range(8)]) / (2 * math.sqrt(2))
            assert np.allclose(orc.apply_op(plus, gate, [w], [a]), exp, atol=1e-12)
            assert np.allclose(orc.apply_op(plus, gate, [w], [a], adjoint=True), exp.conj(), atol=1e-12)


def test_controlled_phase_shift_plus_state():
    """Test_StateVectorCudaManaged_Param.cpp:292-340."""
    plus = np.full(8, 1 / (2 * math.sqrt(2)), dtype=np.complex128)
    for wires, a, ones in (([0, 1], 0.3, [6, 7]), ([1, 2], 2.4, [3, 7])):
        exp = plus.copy()
        exp[ones] *= np.exp(1j * a)
        assert np.allclose(orc.apply_op(plus, "ControlledPhaseShift", wires, [a]), exp, atol=1e-12)


def test_rot_crot_basis():
    """Test_StateVectorCudaManaged_Param.cpp:342-435."""
    angles = [[0.3, 0.8, 2.4], [0.5, 1.1, 3.0], [2.3, 0.1, 0.4]]
    for i, a in enumerate(angles):
        m = orc.rot(*a)
        exp = np.zeros(8, dtype=complex)
        exp[0] = m[0, 0]
        exp[1 << (3 - i - 1)] = m[1, 0]
        assert np.allclose(orc.apply_op(orc.basis_state(3), "Rot", [i], a), exp, atol=1e-12)
    m = orc.rot(*angles[0])
    st = orc.apply_op(orc.basis_state(3), "PauliX", [0])
    exp = np.zeros(8, dtype=complex)
    exp[4] = m[0, 0]
    exp[6] = m[1, 0]
    assert np.allclose(orc.apply_op(st, "CRot", [0, 1], angles[0]), exp, atol=1e-12)
    assert np.allclose(orc.apply_op(orc.basis_state(3), "CRot", [0, 1], angles[0]), orc.basis_state(3))


def test_expval_matrix(kats):
    k = kats["expval_matrix"]
    r = orc.expval_matrix(c_arr(k["state"]), c_arr(k["matrix"]).reshape(8, 8), k["wires"])
    assert r.real == pytest.approx(k["expected"][0], rel=k["rtol"])
    assert r.imag == pytest.approx(k["expected"][1], rel=k["rtol"])


def test_expval_csr(kats):
    k = kats["expval_csr"]
    r = orc.expval_csr(c_arr(k["state"]), k["indptr"], k["indices"], c_arr(k["values"]))
    assert r == pytest.approx(k["expected"], rel=k["rtol"])


def test_pauli_words(kats):
    k = kats["pauli_words"]
    st = c_arr(k["state"])
    for case in k["cases"]:
        r = orc.expval_pauli_words(st, case["words"], case["tgts"], case["coeffs"])
        assert r == pytest.approx(case["expected"], abs=case["atol"]), case
        # cross-check the mask formulation against dense kron matrices
        tot = 0.0
        for w, t, c in zip(case["words"], case["tgts"], case["coeffs"]):
            m = orc._kron(*[orc.PAULI[ch] for ch in w])
            tot += c * orc.expval_matrix(st, m, t).real
        assert tot == pytest.approx(case["expected"], abs=case["atol"])


def test_sparse_pauli_expvals(kats):
    import scipy.sparse as sp
    k = kats["sparse_pauli"]
    st = orc.apply_ops(orc.basis_state(2), k["ops"])
    for word, exp in k["cases"]:
        m = sp.csr_matrix(orc._kron(*[orc.PAULI[ch] for ch in word]))
        r = orc.expval_csr(st, m.indptr, m.indices, m.data)
        assert r == pytest.approx(exp, abs=k["atol"])
        assert orc.expval_pauli_words(st, [word], [[0, 1]], [1.0]) == pytest.approx(exp, abs=1e-12)


def test_probs(kats):
    k = kats["probs"]
    st = orc.apply_ops(orc.basis_state(k["n"]), k["ops"])
    for wires, exp in k["cases"]:
        assert np.allclose(orc.probs(st, wires), exp, atol=k["atol"])
    # cuStateVec order: first listed wire is the LSB (Managed.hpp:949-967)
    st3 = orc.apply_ops(orc.basis_state(3), [{"name": "PauliX", "wires": [2]}])
    assert np.argmax(orc.probs(st3, [1, 2])) == 1
    assert np.argmax(orc.probs_custatevec_order(st3, [1, 2])) == 2


def test_adjoint_jacobians(kats):
    for case in kats["adjoint"]:
        n = case["n"]
        init = orc.basis_state(n) if case["init"] == "zero" else c_arr(case["init"])
        obs = [obs_from_json(o) for o in case["obs"]]
        jac = orc.adjoint_jacobian(init, case["ops"], obs, case["trainable"], apply_operations=True)
        _close(jac, np.asarray(case["expected"]), case)


def test_adjoint_analytic():
    """Test_AdjointDiffGPU.cpp:44-196: d<Z>/dtheta of RX = -sin, d<X>/dtheta of RY = cos."""
    for p in (-math.pi / 7, math.pi / 5, 2 * math.pi / 3):
        j = orc.adjoint_jacobian(orc.basis_state(1), [{"name": "RX", "wires": [0], "params": [p]}],
                                 [("Named", "PauliZ", [0])], [0], apply_operations=True)
        assert j[0, 0] == pytest.approx(-math.sin(p), abs=1e-12)
        j = orc.adjoint_jacobian(orc.basis_state(1), [{"name": "RY", "wires": [0], "params": [p]}],
                                 [("Named", "PauliX", [0])], [0], apply_operations=True)
        assert j[0, 0] == pytest.approx(math.cos(p), abs=1e-12)


def test_adjoint_hermitian_equals_tensor():
    """Test_AdjointDiffGPU.cpp:547-584."""
    p = [-math.pi / 7, math.pi / 5, 2 * math.pi / 3]
    ops = [{"name": "RX", "wires": [i], "params": [p[i]]} for i in range(3)]
    o1 = ("TensorProd", [("Named", "PauliZ", [0]), ("Named", "PauliZ", [1])])
    o2 = ("Hermitian", np.diag([1, -1, -1, 1]).astype(complex), [0, 1])
    j1 = orc.adjoint_jacobian(orc.basis_state(3), ops, [o1], [0, 2], True)
    j2 = orc.adjoint_jacobian(orc.basis_state(3), ops, [o2], [0, 2], True)
    assert np.allclose(j1, j2, atol=1e-12)


def test_generators_vs_finite_difference():
    """Test_Generators.cpp:19-47 style: U(t) ~ exp(i s t G) for every generator."""
    import scipy.linalg as la
    t = 0.37
    for name, (nw, npar) in orc.GATE_ARITY.items():
        if npar != 1:
            continue
        nw = nw or 3
        g, s = orc.generator(name, nw)
        u = orc.gate_matrix(name, [t], nw)
        assert np.allclose(la.expm(1j * s * t * g), u, atol=1e-12), name


def test_all_gates_unitary_and_adjoint_param_order():
    for name, (nw, npar) in orc.GATE_ARITY.items():
        nw = nw or 3
        u = orc.gate_matrix(name, [0.3, 0.8, 2.4][:npar], nw)
        assert np.allclose(u @ u.conj().T, np.eye(u.shape[0]), atol=1e-12), name
    # Managed.hpp:215-224: adjoint Rot == getRot(p2,p1,p0)^dagger ... equals Rot(p)^dagger
    a = [0.3, 0.8, 2.4]
    assert np.allclose(orc.rot(*a).conj().T, orc.rz(-a[0]) @ orc.ry(-a[1]) @ orc.rz(-a[2]))


def test_strongly_entangling_layers_shape():
    w = np.random.default_rng(1337).uniform(0, 2 * np.pi, (2, 20, 3))
    ops = orc.strongly_entangling_layers(w)
    assert len(ops) == 160 and sum(1 for o in ops if o["params"]) == 120
    assert ops[60]["wires"] == [0, 1] and ops[159]["wires"] == [19, 1]


def test_sample_definition():
    st = orc.apply_ops(orc.basis_state(3), [{"name": "Hadamard", "wires": [0]}, {"name": "CNOT", "wires": [0, 2]}])
    s = orc.sample(st, 1000, seed=7)
    assert s.shape == (1000, 3) and set(np.unique(s)) <= {0, 1}
    assert np.all(s[:, 0] == s[:, 2]) and np.all(s[:, 1] == 0)
    assert 400 < s[:, 0].sum() < 600
    assert np.array_equal(s, orc.sample(st, 1000, seed=7))


def test_closed_form_expvals_and_variances(kats):
    """tests/test_expval.py:38-200, tests/test_var.py:34-130: the closed forms the reference asserts for named and tensor
    observables on 3-wire circuits (48 cases on the tests' own parameter grids)."""
    assert len(kats["closed_forms"]) == 48
    for case in kats["closed_forms"]:
        psi = orc.apply_ops(orc.basis_state(case["n"]), case["ops"])
        obs = obs_from_json(case["obs"])
        e = orc.expval_obs(psi, obs)
        if "expval" in case:
            _close(e, case["expval"], case)
        if "var" in case:
            o_psi = orc.apply_observable(psi.copy(), obs)
            _close(np.vdot(o_psi, o_psi).real - e ** 2, case["var"], case)


def test_cy_and_identity_analytic():
    """The two named gates for which the reference's tests hold no in/out vector.  CY is pinned by the reference's own
    matrix (simulator/cuGates_host.hpp:154-163: rows (1,0,0,0), (0,1,0,0), (0,0,0,-i), (0,0,i,0), first wire = control)
    and by CY = S(t) CNOT S(t)^dagger; Identity is a no-op on any wire (StateVectorCudaManaged.hpp:321-323)."""
    cy = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, -1j], [0, 0, 1j, 0]], dtype=complex)
    assert np.array_equal(orc.gate_matrix("CY", [], 2), cy)
    rng = np.random.default_rng(8)
    n = 4
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    for c, t in ((0, 1), (3, 1), (2, 0)):
        got = orc.apply_op(psi, "CY", [c, t], [])
        alt = orc.apply_op(orc.apply_op(orc.apply_op(psi, "S", [t], [], adjoint=True), "CNOT", [c, t], []), "S", [t], [])
        assert np.max(np.abs(got - alt)) < 1e-15
        # control = first wire: amplitudes with the control bit clear are untouched
        keep = ((np.arange(1 << n) >> (n - 1 - c)) & 1) == 0
        assert np.array_equal(got[keep], psi[keep])
        assert np.max(np.abs(orc.apply_op(got, "CY", [c, t], [], adjoint=True) - psi)) < 1e-15
    for w in range(n):
        assert np.array_equal(orc.apply_op(psi, "Identity", [w], []), psi)
    assert np.array_equal(orc.gate_matrix("Identity", [], 1), np.eye(2))
