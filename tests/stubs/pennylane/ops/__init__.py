import numpy as np
import scipy.sparse as sp

from ..operation import Observable, Operation, Operator, Tensor, expand_matrix
from ..wires import Wires

I2 = np.eye(2, dtype=complex)
X = np.array([[0, 1], [1, 0]], dtype=complex)
Y = np.array([[0, -1j], [1j, 0]], dtype=complex)
Z = np.array([[1, 0], [0, -1]], dtype=complex)
H = np.array([[1, 1], [1, -1]], dtype=complex) / np.sqrt(2)


def _rot(p, theta):
    return np.cos(theta / 2) * np.eye(len(p)) - 1j * np.sin(theta / 2) * p


def _ctrl(u, n_ctrl=1):
    d = u.shape[0]
    m = np.eye(d << n_ctrl, dtype=complex)
    m[-d:, -d:] = u
    return m


class _ObsOp(Observable, Operation):
    pass


class Identity(_ObsOp):
    def __init__(self, wires=None, id=None):
        super().__init__(wires=wires)

    def compute_matrix(self):
        return np.eye(1 << len(self.wires), dtype=complex)

    def eigvals(self):
        return np.ones(1 << len(self.wires))


def _fixed(name, mat, diag_gates=None, eig=None, obs=False):
    base = (_ObsOp,) if obs else (Operation,)

    def __init__(self, wires=None, id=None):
        Operator.__init__(self, wires=wires)

    def compute_matrix(self):
        return mat

    ns = {"__init__": __init__, "compute_matrix": compute_matrix}
    if diag_gates is not None:
        ns["diagonalizing_gates"] = lambda self: diag_gates(self)
    if eig is not None:
        ns["eigvals"] = lambda self: np.array(eig, dtype=float)
    return type(name, base, ns)


def _param(name, fn):
    def __init__(self, *params, wires=None, id=None):
        Operator.__init__(self, *params, wires=wires)

    def compute_matrix(self, *params):
        return fn(*[float(p) for p in params]) if name != "MultiRZ" else fn(float(params[0]), len(self.wires))

    return type(name, (Operation,), {"__init__": __init__, "compute_matrix": compute_matrix})


S_M = np.diag([1, 1j])
T_M = np.diag([1, np.exp(1j * np.pi / 4)])
SWAP_M = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=complex)

Hadamard = _fixed("Hadamard", H, lambda self: [RY(-np.pi / 4, wires=self.wires)], [1, -1], obs=True)
PauliZ = _fixed("PauliZ", Z, lambda self: [], [1, -1], obs=True)
PauliX = _fixed("PauliX", X, lambda self: [Hadamard(wires=self.wires)], [1, -1], obs=True)
PauliY = _fixed("PauliY", Y, lambda self: [PauliZ(wires=self.wires), S(wires=self.wires), Hadamard(wires=self.wires)],
                [1, -1], obs=True)
S = _fixed("S", S_M)
T = _fixed("T", T_M)
SX = _fixed("SX", 0.5 * np.array([[1 + 1j, 1 - 1j], [1 - 1j, 1 + 1j]]))
CNOT = _fixed("CNOT", _ctrl(X))
CZ = _fixed("CZ", _ctrl(Z))
CY = _fixed("CY", _ctrl(Y))
SWAP = _fixed("SWAP", SWAP_M)
Toffoli = _fixed("Toffoli", _ctrl(X, 2))
CSWAP = _fixed("CSWAP", _ctrl(SWAP_M))

RX = _param("RX", lambda t: _rot(X, t))
RY = _param("RY", lambda t: _rot(Y, t))
RZ = _param("RZ", lambda t: _rot(Z, t))
PhaseShift = _param("PhaseShift", lambda t: np.diag([1, np.exp(1j * t)]))
CRX = _param("CRX", lambda t: _ctrl(_rot(X, t)))
CRY = _param("CRY", lambda t: _ctrl(_rot(Y, t)))
CRZ = _param("CRZ", lambda t: _ctrl(_rot(Z, t)))
ControlledPhaseShift = _param("ControlledPhaseShift", lambda t: np.diag([1, 1, 1, np.exp(1j * t)]))
IsingXX = _param("IsingXX", lambda t: _rot(np.kron(X, X), t))
IsingYY = _param("IsingYY", lambda t: _rot(np.kron(Y, Y), t))
IsingZZ = _param("IsingZZ", lambda t: _rot(np.kron(Z, Z), t))


def _multirz(t, n):
    z = np.array([1.0])
    for _ in range(n):
        z = np.kron(z, np.array([1.0, -1.0]))
    return np.diag(np.exp(-0.5j * t * z))


MultiRZ = _param("MultiRZ", _multirz)


def _single_excitation(t):
    c, s = np.cos(t / 2), np.sin(t / 2)
    return np.array([[1, 0, 0, 0], [0, c, -s, 0], [0, s, c, 0], [0, 0, 0, 1]], dtype=complex)


def _double_excitation(t):
    m = np.eye(16, dtype=complex)
    c, s = np.cos(t / 2), np.sin(t / 2)
    m[3, 3] = c
    m[12, 12] = c
    m[3, 12] = -s
    m[12, 3] = s
    return m


SingleExcitation = _param("SingleExcitation", _single_excitation)
DoubleExcitation = _param("DoubleExcitation", _double_excitation)


def _rot_matrix(phi, theta, omega):
    return _rot(Z, omega) @ _rot(Y, theta) @ _rot(Z, phi)


class Rot(Operation):
    def __init__(self, phi, theta, omega, wires=None, id=None):
        super().__init__(phi, theta, omega, wires=wires)

    def compute_matrix(self, phi, theta, omega):
        return _rot_matrix(float(phi), float(theta), float(omega))

    def expand(self):
        from ..tape import QuantumScript

        phi, theta, omega = self.parameters
        return QuantumScript([RZ(phi, wires=self.wires), RY(theta, wires=self.wires), RZ(omega, wires=self.wires)], [])


class CRot(Operation):
    def __init__(self, phi, theta, omega, wires=None, id=None):
        super().__init__(phi, theta, omega, wires=wires)

    def compute_matrix(self, phi, theta, omega):
        return _ctrl(_rot_matrix(float(phi), float(theta), float(omega)))


class QubitUnitary(Operation):
    def __init__(self, U, wires=None, id=None):
        super().__init__(np.asarray(U, dtype=complex), wires=wires)

    def compute_matrix(self, U):
        return U


class StatePrep(Operation):
    def __init__(self, state, wires=None, id=None):
        super().__init__(np.asarray(state), wires=wires)


QubitStateVector = StatePrep


class BasisState(Operation):
    def __init__(self, n, wires=None, id=None):
        super().__init__(np.asarray(n), wires=wires)


class Hermitian(Observable):
    def __init__(self, A, wires=None, id=None):
        super().__init__(np.asarray(A, dtype=complex), wires=wires)

    def compute_matrix(self, A):
        return A


class Projector(Observable):
    def __init__(self, basis_state, wires=None, id=None):
        super().__init__(np.asarray(basis_state), wires=wires)

    def compute_matrix(self, b):
        idx = int("".join(str(int(x)) for x in b), 2)
        m = np.zeros((1 << len(b), 1 << len(b)), dtype=complex)
        m[idx, idx] = 1
        return m


class Hamiltonian(Observable):
    def __init__(self, coeffs, observables, id=None):
        self._coeffs = np.asarray(coeffs)
        self._ops = list(observables)
        self.data = []
        self._wires = Wires.all_wires([o.wires for o in self._ops])
        self._name = "Hamiltonian"

    @property
    def coeffs(self):
        return self._coeffs

    @property
    def ops(self):
        return self._ops

    @property
    def parameters(self):
        return list(self._coeffs)

    def terms(self):
        return list(self._coeffs), self._ops

    def matrix(self, wire_order=None):
        wire_order = list(self.wires) if wire_order is None else list(wire_order)
        m = np.zeros((1 << len(wire_order),) * 2, dtype=complex)
        for c, o in zip(self._coeffs, self._ops):
            m = m + c * expand_matrix(o.matrix(), o.wires, wire_order)
        return m

    def sparse_matrix(self, wire_order=None):
        return sp.csr_matrix(self.matrix(wire_order))


class SparseHamiltonian(Observable):
    def __init__(self, H, wires=None, id=None):
        self._H = sp.csr_matrix(H)
        self.data = [self._H]
        self._wires = Wires(wires)
        self._name = "SparseHamiltonian"

    def sparse_matrix(self, wire_order=None):
        return self._H

    def matrix(self, wire_order=None):
        return self._H.toarray()


class Sum(Observable):  # only referenced through isinstance checks
    pass
