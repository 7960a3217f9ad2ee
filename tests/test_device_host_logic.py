"""Host-side logic of the LightningGPU device mirror that needs no GPU: argument validation of vjp and the
operation check of the adjoint method (reference: pennylane_lightning_gpu/lightning_gpu.py:622-636, 771-810)."""
import numpy as np
import pytest

from pennylane_lightning_gpu_b200 import _build


@pytest.fixture(scope="module")
def lg():
    _build.build_all()
    from pennylane_lightning_gpu_b200 import lightning_gpu

    return lightning_gpu


def _bare_device(lg):
    dev = object.__new__(lg.LightningGPU)  # no GPU state: the checks below run before any call into the binary
    dev.R_DTYPE, dev.C_DTYPE, dev.shots = np.float64, np.complex128, None
    return dev


def test_vjp_argument_checks(lg):
    dev = _bare_device(lg)
    ops = [lg.Op("RX", [0], [0.1])]
    obs = [lg.Obs("PauliZ", [0])]
    with pytest.raises(ValueError, match="real-valued dy"):
        dev.vjp(ops, obs, [1j], trainable_params=[0])
    with pytest.raises(ValueError, match="same as the length of dy"):
        dev.vjp(ops, obs, [1.0, 2.0], trainable_params=[0])
    out = dev.vjp(ops, obs, [0.0], trainable_params=[0])  # dy == 0: zeros without touching the device
    assert out.shape == (1,) and out[0] == 0.0


def test_adjoint_operation_check(lg):
    ok = [lg.Op("Rot", [0], [0.1, 0.2, 0.3]), lg.Op("RX", [1], [0.3]), lg.Op("CNOT", [0, 1]),
          lg.Op("QubitStateVector", [0], [np.array([1.0, 0.0]), 0])]
    lg.LightningGPU._check_adjdiff_supported_operations(ok)
    with pytest.raises(ValueError, match='CRot operation is not supported using the "adjoint"'):
        lg.LightningGPU._check_adjdiff_supported_operations([lg.Op("CRot", [0, 1], [0.1, 0.2, 0.3])])


def test_constructor_argument_checks(lg):
    with pytest.raises(TypeError, match="Unsupported complex Type"):
        lg.LightningGPU(2, c_dtype=np.float64)
    with pytest.raises(TypeError, match="mpi_buf_size"):
        lg.LightningGPU(2, mpi_buf_size=-1)
    with pytest.raises(TypeError, match="power of 2"):
        lg.LightningGPU(2, mpi_buf_size=3)


def test_dense_matrix_of_composite_observables(lg):
    """LightningGPU._matrix_of (what qml.matrix(observable) is to the reference's expval / var): tensor products with
    non-Pauli factors on unordered wires and Hamiltonians of them, against factor-by-factor application in the oracle."""
    from oracle import np_oracle as orc

    rng = np.random.default_rng(0)
    h = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    h = h + h.conj().T
    named = lg.LightningGPU._NAMED_MATRIX
    t = lg.Obs("Tensor", terms=[lg.Obs("Hadamard", [2]), lg.Obs("Hermitian", [0, 3], matrix=h), lg.Obs("PauliY", [1])])
    wires, m = lg.LightningGPU._matrix_of(t)
    assert wires == [2, 0, 3, 1] and np.allclose(m, m.conj().T)
    psi = rng.normal(size=16) + 1j * rng.normal(size=16)
    psi /= np.linalg.norm(psi)
    phi = orc.apply_matrix(psi.copy(), named["Hadamard"], [2])
    phi = orc.apply_matrix(phi, h, [0, 3])
    phi = orc.apply_matrix(phi, named["PauliY"], [1])
    assert abs(orc.expval_matrix(psi, m, wires) - np.vdot(psi, phi)) < 1e-12
    ham = lg.Obs("Hamiltonian", coeffs=[0.5, -2.0], terms=[lg.Obs("PauliZ", [3]), t])
    wires2, m2 = lg.LightningGPU._matrix_of(ham)
    phi2 = 0.5 * orc.apply_matrix(psi.copy(), named["PauliZ"], [3]) - 2.0 * phi
    assert sorted(wires2) == [0, 1, 2, 3]
    assert abs(orc.expval_matrix(psi, m2, wires2) - np.vdot(psi, phi2)) < 1e-12
    with pytest.raises(ValueError, match="distinct wires"):
        lg.LightningGPU._matrix_of(lg.Obs("Tensor", terms=[lg.Obs("PauliX", [0]), lg.Obs("Hadamard", [0])]))
    with pytest.raises(NotImplementedError):
        lg.LightningGPU._matrix_of(lg.Obs("Tensor", terms=[lg.Obs("Hadamard", [w]) for w in range(11)]))


def test_capabilities_and_jacobian_processing(lg):
    caps = lg.LightningGPU.capabilities()
    assert caps["model"] == "qubit" and caps["returns_state"] and caps["supports_finite_shots"]
    assert caps["supports_inverse_operations"] and caps["supports_analytic_computation"] and "passthru_devices" not in caps
    proc = lg.LightningGPU._adjoint_jacobian_processing
    assert proc(np.array([[0.5]])).shape == ()                      # one observable, one parameter
    one_obs = proc(np.array([[0.1, 0.2, 0.3]]))
    assert isinstance(one_obs, tuple) and len(one_obs) == 3 and float(one_obs[1]) == 0.2
    many = proc(np.arange(6.0).reshape(2, 3))
    assert isinstance(many, tuple) and isinstance(many[0], tuple) and float(many[1][2]) == 5.0


class _FakeState:
    """Records what the device sends to the binary module (no GPU)."""

    def __init__(self):
        self.seeds, self.applied, self.jac_calls = [], [], []

    def GenerateSamples(self, n, shots, seed=None):
        self.seeds.append(seed)
        return np.zeros((shots, n), dtype=np.uint64)

    def apply(self, *a):
        self.applied.append(a)

    def RX(self, wires, inv, params):
        self.applied.append(("RX", wires, inv, params))


def test_seeded_sampling_draws_a_fresh_sub_seed_per_call(lg):
    """ADVICE (round 1): with `seed` set every generate_samples() call used to re-seed with the same value, so shot noise was
    identical across observables and executions.  Now one stream of sub-seeds per device, reproducible for a given seed."""
    def device(seed):
        dev = _bare_device(lg)
        dev._gpu_state, dev._seed, dev.num_wires, dev.shots = _FakeState(), seed, 3, 10
        return dev

    a, b, c = device(7), device(7), device(8)
    for d in (a, b, c):
        for _ in range(4):
            d.generate_samples()
    assert len(set(a._gpu_state.seeds)) == 4                      # fresh uniforms on every call
    assert a._gpu_state.seeds == b._gpu_state.seeds               # reproducible
    assert a._gpu_state.seeds != c._gpu_state.seeds
    unseeded = device(None)
    unseeded.generate_samples()
    assert unseeded._gpu_state.seeds == [None]                    # the binary draws from random_device


def test_apply_sends_the_whole_tape_in_one_call(lg):
    """apply_cq hands the list of operations to the binary in ONE call (vector form of `apply`, fused executor); a matrix
    operation travels with its matrix; `fuse_ops=False` restores one call per operation; `inverse` is per operation."""
    dev = _bare_device(lg)
    dev._gpu_state, dev._fuse_ops = _FakeState(), True
    u = np.eye(2, dtype=complex)
    ops = [lg.Op("RX", [0], [0.1]), lg.Op("Identity", [1]), lg.Op("QubitUnitary", [1], matrix=u),
           lg.Op("RX", [2], [0.3], inverse=True)]
    dev.apply_cq(ops)
    assert len(dev._gpu_state.applied) == 1
    names, wires, invs, params, mats = dev._gpu_state.applied[0]
    assert names == ["RX", "QubitUnitary", "RX"] and wires == [[0], [1], [2]] and invs == [False, False, True]
    assert params == [[0.1], [], [0.3]] and len(mats[0]) == 0 and np.array_equal(np.asarray(mats[1]).reshape(2, 2), u)
    dev._gpu_state, dev._fuse_ops = _FakeState(), False
    dev.apply_cq(ops)
    assert [a[0] for a in dev._gpu_state.applied] == ["RX", "QubitUnitary", "RX"]


def test_vjp_zero_dy_counts_all_parameters(lg):
    """ADVICE (round 1): dy == 0 with trainable_params=None returned zeros(0); now one zero per parameter of the tape (Rot
    counts three), and the length check comes first."""
    dev = _bare_device(lg)
    ops = [lg.Op("RX", [0], [0.1]), lg.Op("Rot", [1], [0.1, 0.2, 0.3]), lg.Op("CNOT", [0, 1])]
    out = dev.vjp(ops, [lg.Obs("PauliZ", [0])], [0.0])
    assert out.shape == (4,) and not out.any()
    with pytest.raises(ValueError, match="same as the length of dy"):
        dev.vjp(ops, [lg.Obs("PauliZ", [0])], [0.0, 0.0])
    with pytest.raises(ValueError, match="inverse of Rot"):
        lg.LightningGPU._check_adjdiff_supported_operations([lg.Op("Rot", [0], [0.1, 0.2, 0.3], inverse=True)])
