import numpy as np
from numpy import linalg  # noqa: F401


def allclose(a, b, rtol=1e-05, atol=1e-08, **k):
    return bool(np.allclose(np.asarray(a), np.asarray(b), rtol=rtol, atol=atol))


def convert_like(x, like):
    return np.asarray(x)


def is_abstract(x):
    return False


def T(x):
    return np.transpose(x)


def conj(x):
    return np.conj(x)


def dot(a, b):
    return np.dot(a, b)
