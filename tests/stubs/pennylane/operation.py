import numpy as np

from .wires import Wires


def expand_matrix(m, wires, wire_order):
    """m on `wires` (first wire = most significant) -> matrix on wire_order (identity elsewhere)."""
    wires, wire_order = list(wires), list(wire_order)
    n, k = len(wire_order), len(wires)
    t = np.asarray(m, dtype=complex).reshape([2] * (2 * k))
    full = np.eye(1 << n, dtype=complex).reshape([2] * (2 * n))
    pos = [wire_order.index(w) for w in wires]
    # contract the operator's input axes with the row axes of the identity at the operator's wires
    out = np.tensordot(t, full, axes=(list(range(k, 2 * k)), pos))
    out = np.moveaxis(out, list(range(k)), pos)
    return out.reshape(1 << n, 1 << n)


class Operator:
    num_wires = None
    num_params = 0
    _pauli_rep = None

    def __init__(self, *params, wires=None, id=None):
        self.data = [p for p in params]
        self._wires = Wires(wires if wires is not None else [])
        self._name = type(self).__name__

    @property
    def name(self):
        return self._name

    @property
    def wires(self):
        return self._wires

    @property
    def parameters(self):
        return list(self.data)

    @property
    def num_params(self):  # noqa: F811
        return len(self.data)

    def matrix(self, wire_order=None):
        m = np.asarray(self.compute_matrix(*self.parameters), dtype=complex)
        if wire_order is None or list(wire_order) == list(self.wires):
            return m
        return expand_matrix(m, self.wires, wire_order)

    def compute_matrix(self, *params):
        raise NotImplementedError(self.name)

    def diagonalizing_gates(self):
        return []

    def eigvals(self):
        return np.linalg.eigvalsh(self.matrix())

    def __repr__(self):
        return f"{self.name}({', '.join(str(p) for p in self.parameters)}, wires={list(self.wires)})"


class Operation(Operator):
    pass


class Observable(Operator):
    return_type = None

    def __matmul__(self, other):
        if isinstance(other, Tensor):
            return Tensor(self, *other.obs)
        return Tensor(self, other)


class Tensor(Observable):
    def __init__(self, *obs):
        self.obs = []
        for o in obs:
            self.obs.extend(o.obs if isinstance(o, Tensor) else [o])
        self.data = []
        self._wires = Wires.all_wires([o.wires for o in self.obs])
        self._name = "Tensor"

    @property
    def name(self):
        return [o.name for o in self.obs]

    @property
    def parameters(self):
        return [o.parameters for o in self.obs]

    @property
    def num_params(self):
        return 0

    @property
    def non_identity_obs(self):
        return [o for o in self.obs if o.name != "Identity"]

    def __matmul__(self, other):
        return Tensor(*self.obs, other)

    def matrix(self, wire_order=None):
        wire_order = list(self.wires) if wire_order is None else list(wire_order)
        m = np.eye(1 << len(wire_order), dtype=complex)
        for o in self.obs:
            m = expand_matrix(o.matrix(), o.wires, wire_order) @ m
        return m

    def diagonalizing_gates(self):
        out = []
        for o in self.obs:
            out.extend(o.diagonalizing_gates())
        return out

    def eigvals(self):
        ev = np.array([1.0])
        for o in self.obs:
            ev = np.kron(ev, o.eigvals())
        return ev
