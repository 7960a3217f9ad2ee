"""Stand-in for pennylane_lightning.lightning_qubit (the reference imports the class as its CPU fall-back base,
lightning_gpu.py:37, 977-998).  It must never be instantiated by the tests: the B200 engine has no CPU fall-back."""


class LightningQubit:  # pragma: no cover
    def __init__(self, *a, **k):
        raise RuntimeError("the CPU fall-back of lightning.gpu was selected: the native module did not load")
