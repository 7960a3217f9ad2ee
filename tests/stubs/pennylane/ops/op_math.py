import numpy as np

from ..operation import Operation


class Adjoint(Operation):
    def __init__(self, base):
        self.base = base
        self.data = base.data
        self._wires = base.wires
        self._name = f"Adjoint({base.name})"

    def matrix(self, wire_order=None):
        return np.conj(self.base.matrix(wire_order)).T


def adjoint(op):
    return Adjoint(op)
