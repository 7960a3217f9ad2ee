class Wires(tuple):
    """Ordered collection of unique wire labels."""

    def __new__(cls, wires=()):
        if isinstance(wires, Wires):
            return wires
        if isinstance(wires, (int, str)):
            wires = [wires]
        wires = list(wires)
        if len(set(wires)) != len(wires):
            raise ValueError(f"Wires must be unique; got {wires}.")
        return super().__new__(cls, wires)

    @property
    def labels(self):
        return tuple(self)

    def tolist(self):
        return list(self)

    def toarray(self):
        import numpy as np

        return np.array(list(self))

    def index(self, wire):
        if isinstance(wire, Wires):
            wire = wire[0]
        return tuple.index(self, wire)

    def indices(self, wires):
        return [self.index(w) for w in Wires(wires)]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            return Wires(tuple.__getitem__(self, idx))
        return tuple.__getitem__(self, idx)

    def __eq__(self, other):
        return isinstance(other, (Wires, tuple, list)) and list(self) == list(other)

    def __ne__(self, other):
        return not self == other

    def __hash__(self):
        return tuple.__hash__(self)

    def __add__(self, other):
        return Wires.all_wires([self, Wires(other)])

    def __repr__(self):
        return f"<Wires = {list(self)}>"

    @staticmethod
    def all_wires(list_of_wires):
        out = []
        for ws in list_of_wires:
            for w in Wires(ws):
                if w not in out:
                    out.append(w)
        return Wires(out)
