"""B200-native state-vector engine for the hot path of PennyLane's ``lightning.gpu``.

Layers (bottom up):
  csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/qsv_b200.h) -> lib/libqsv_b200.so
  _cabi.py         ctypes binding of that ABI (tests, bench, Python device)
  src/             C++ class surface of the reference (StateVectorCudaManaged, AdjointJacobianGPU,
                   ObservablesGPU, ...) over the C ABI + pybind11 module ``lightning_gpu_qubit_ops``
  lightning_gpu.py the ``LightningGPU`` device mirror (PennyLane import-guarded)

There is no CPU fallback anywhere in this package.
"""
from ._cabi import LIB_PATH, Observable, Ops, QsvError, StateVector, device_arch, device_count, lib  # noqa: F401

__version__ = "0.1.0"
