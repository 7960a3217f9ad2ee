class ObservableReturnTypes:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


Expectation = ObservableReturnTypes("expval")
Variance = ObservableReturnTypes("var")
Sample = ObservableReturnTypes("sample")
Probability = ObservableReturnTypes("probs")
State = ObservableReturnTypes("state")


class MeasurementProcess:
    def __init__(self, return_type, obs=None, wires=None):
        self.return_type = return_type
        self.obs = obs
        self._wires = wires

    @property
    def wires(self):
        return self.obs.wires if self.obs is not None else self._wires


def expval(op):
    return MeasurementProcess(Expectation, obs=op)


def var(op):
    return MeasurementProcess(Variance, obs=op)


def sample(op=None, wires=None):
    return MeasurementProcess(Sample, obs=op, wires=wires)


def probs(wires=None, op=None):
    return MeasurementProcess(Probability, obs=op, wires=wires)


def state():
    return MeasurementProcess(State)
