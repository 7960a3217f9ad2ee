import numpy as np

from .wires import Wires


class QubitDevice:
    """The slice of pennylane.QubitDevice that lightning_gpu.py builds on."""

    operations = set()
    observables = set()
    _asarray = staticmethod(np.asarray)
    _reshape = staticmethod(np.reshape)

    def __init__(self, wires=1, shots=None, *, r_dtype=np.float64, c_dtype=np.complex128, analytic=None):
        self._wires = Wires(range(wires)) if isinstance(wires, int) else Wires(wires)
        self.num_wires = len(self._wires)
        self._wire_map = {w: i for i, w in enumerate(self._wires)}
        self.shots = shots
        self.R_DTYPE = r_dtype
        self.C_DTYPE = c_dtype
        self._samples = None

    @property
    def wires(self):
        return self._wires

    @property
    def wire_map(self):
        return self._wire_map

    def map_wires(self, wires):
        return Wires([self._wire_map[w] for w in Wires(wires)])

    @staticmethod
    def _get_batch_size(tensor, expected_shape, expected_size):
        return None if np.ndim(tensor) == len(expected_shape) else np.shape(tensor)[0]

    @classmethod
    def capabilities(cls):
        return {"model": "qubit", "supports_broadcasting": False, "passthru_devices": {}}

    def supports_operation(self, name):
        return name in self.operations

    def reset(self):
        self._samples = None

    def execute(self, circuit, **kwargs):
        """apply + statistics for expectation values (what adjoint_jacobian needs from it)."""
        from .measurements import Expectation

        self.apply(circuit.operations, rotations=[])
        out = []
        for m in circuit.measurements:
            if m.return_type is Expectation:
                out.append(self.expval(m.obs))
        return np.asarray(out)

    # -- generic pieces the reference falls back on ------------------------------------------------------------
    def expval(self, observable, shot_range=None, bin_size=None):
        psi = np.asarray(self.state).reshape(-1)
        m = observable.matrix(wire_order=list(self.wires)) if len(observable.wires) < self.num_wires else observable.matrix()
        from .operation import expand_matrix

        if m.shape[0] != psi.size:
            m = expand_matrix(m, self.map_wires(observable.wires), list(range(self.num_wires)))
        return float(np.real(np.vdot(psi, m @ psi)))

    def sample(self, observable, shot_range=None, bin_size=None, counts=False):
        device_wires = self.map_wires(observable.wires)
        s = np.asarray(self._samples)[:, list(device_wires)]
        idx = s @ (1 << np.arange(len(device_wires) - 1, -1, -1))
        return np.asarray(observable.eigvals())[idx]

    def estimate_probability(self, wires=None, shot_range=None, bin_size=None):
        wires = self.map_wires(wires or self.wires)
        s = np.asarray(self._samples)[:, list(wires)]
        idx = s @ (1 << np.arange(len(wires) - 1, -1, -1))
        return np.bincount(idx, minlength=1 << len(wires)) / len(idx)

    def statistics(self, circuit, shot_range=None, bin_size=None):
        return [self.expval(m.obs) for m in circuit.measurements]

    def _get_diagonalizing_gates(self, circuit):
        out = []
        for m in circuit.measurements:
            if m.obs is not None:
                out.extend(m.obs.diagonalizing_gates())
        return out
