"""A SMALL stand-in for the parts of PennyLane (>= 0.30, < 0.33 API) that the reference's `lightning_gpu.py` and
`_serialize.py` touch, so that those two files can be imported UNCHANGED against `lightning_gpu_qubit_ops` built from this
repository (tests/test_reference_device_unchanged.py).  PennyLane itself is not installable here (no network).  Test
infrastructure only: semantics follow PennyLane's documented behaviour for the subset (operator matrices in wire order,
`Rot = RZ(omega) RY(theta) RZ(phi)`, `Tensor.name` a list, trainable parameters indexed over the flat parameter list ...).
"""
from __future__ import annotations

import numpy as _np

from . import math  # noqa: F401
from .wires import Wires  # noqa: F401
from .operation import Operation, Observable, Operator, Tensor  # noqa: F401
from .ops import (  # noqa: F401
    Identity, PauliX, PauliY, PauliZ, Hadamard, S, T, SX, RX, RY, RZ, PhaseShift, Rot, CNOT, CZ, CY, SWAP, Toffoli, CSWAP, CRX,
    CRY, CRZ, CRot, ControlledPhaseShift, IsingXX, IsingYY, IsingZZ, MultiRZ, SingleExcitation, DoubleExcitation,
    QubitUnitary, StatePrep, QubitStateVector, BasisState, Hermitian, Projector, Hamiltonian, SparseHamiltonian, Sum,
)
from .ops.op_math import Adjoint, adjoint  # noqa: F401
from .measurements import expval, var, probs, sample, state  # noqa: F401
from . import tape  # noqa: F401
from .devices import QubitDevice  # noqa: F401

__version__ = "0.32.0-stub"


class DeviceError(Exception):
    pass


class QuantumFunctionError(Exception):
    pass


class BooleanFn:
    def __init__(self, fn):
        self.fn = fn

    def __call__(self, *a, **k):
        return self.fn(*a, **k)


def active_return():
    return True


def matrix(op, wire_order=None):
    """Matrix of an operator in the order of its own wires (or of wire_order)."""
    m = op.matrix()
    if wire_order is None or list(wire_order) == list(op.wires):
        return m
    from .operation import expand_matrix

    return expand_matrix(m, list(op.wires), list(wire_order))
