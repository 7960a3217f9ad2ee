import copy as _copy

from .wires import Wires


class QuantumScript:
    """Operations + measurements; trainable parameters are indices into the flat list of all operation parameters
    (state preparations included), all trainable by default -- PennyLane's convention."""

    def __init__(self, ops=(), measurements=(), prep=(), shots=None):
        self._ops = list(prep) + list(ops)
        self._measurements = list(measurements)
        self._par_info = []
        for i, op in enumerate(self._ops):
            for p in range(len(op.parameters)):
                self._par_info.append((op, i, p))
        self.trainable_params = list(range(len(self._par_info)))

    @property
    def operations(self):
        return self._ops

    @property
    def measurements(self):
        return self._measurements

    @property
    def observables(self):
        return [m.obs if m.obs is not None else m for m in self._measurements]

    @property
    def wires(self):
        return Wires.all_wires([o.wires for o in self._ops] + [m.wires for m in self._measurements if m.wires is not None])

    def get_operation(self, idx):
        """(operation, operation index, parameter index within the operation) of the idx-th TRAINABLE parameter."""
        return self._par_info[self.trainable_params[idx]]

    def get_parameters(self, trainable_only=True):
        idx = self.trainable_params if trainable_only else range(len(self._par_info))
        return [self._par_info[i][0].parameters[self._par_info[i][2]] for i in idx]

    def copy(self, copy_operations=False):
        t = _copy.copy(self)
        t._measurements = list(self._measurements)
        return t


QuantumTape = QuantumScript
