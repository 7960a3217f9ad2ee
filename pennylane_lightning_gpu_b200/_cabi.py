"""ctypes binding of libqsv_b200.so (include/qsv_b200.h) with numpy-friendly wrappers.

This is the thinnest possible host layer over the C ABI: it is what the parity tests, bench.py and
the Python device use.  There is NO CPU fallback: importing works without a GPU (so that the symbol
table can be checked), but every compute call goes to the CUDA library and raises ``QsvError``
(the ``LightningException`` analogue, bindings/Bindings.cpp:1734) when it fails.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Sequence

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# QSV_LIB_PATH: another build of the same library (kernel A/B experiments, tools/regs_variants.py)
LIB_PATH = os.environ.get("QSV_LIB_PATH") or os.path.join(_PKG, "lib", "libqsv_b200.so")

QSV_C64, QSV_C128 = 0, 1


class QsvError(RuntimeError):
    """Raised for every non-zero status from the C ABI (message = qsv_last_error())."""


_lib = None

_P = C.c_void_p
_I = C.c_int
_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)
_I64P = C.POINTER(C.c_int64)
_U64P = C.POINTER(C.c_uint64)

# name -> (restype, argtypes); must list every symbol include/qsv_b200.h declares
SIGNATURES = {
    "qsv_last_error": (C.c_char_p, []),
    "qsv_version": (_I, []),
    "qsv_device_count": (_I, [_IP]),
    "qsv_device_arch": (_I, [_I, _IP, _IP]),
    "qsv_device_mem_info": (_I, [_I, C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "qsv_device_reset": (_I, []),
    "qsv_enable_peer_access": (_I, []),
    "qsv_create": (_I, [_I, _I, _I, C.POINTER(_P)]),
    "qsv_create_external": (_I, [_I, _I, _I, _P, _P, C.POINTER(_P)]),
    "qsv_destroy": (_I, [_P]),
    "qsv_set_stream": (_I, [_P, _P]),
    "qsv_synchronize": (_I, [_P]),
    "qsv_data_ptr": (_P, [_P]),
    "qsv_num_qubits": (_I, [_P]),
    "qsv_dtype": (_I, [_P]),
    "qsv_device": (_I, [_P]),
    "qsv_set_basis_state": (_I, [_P, C.c_uint64]),
    "qsv_set_state_vector": (_I, [_P, _I64P, _P, C.c_size_t]),
    "qsv_h2d": (_I, [_P, _P, C.c_size_t]),
    "qsv_d2h": (_I, [_P, _P, C.c_size_t]),
    "qsv_d2d": (_I, [_P, _P]),
    "qsv_apply_named": (_I, [_P, C.c_char_p, _IP, _I, _I, _DP, _I]),
    "qsv_apply_matrix": (_I, [_P, _DP, _IP, _I, _IP, _I, _I]),
    "qsv_apply_generator": (_I, [_P, C.c_char_p, _IP, _I, _I, _DP]),
    "qsv_ops_create": (_I, [C.POINTER(_P)]),
    "qsv_ops_destroy": (_I, [_P]),
    "qsv_ops_append": (_I, [_P, C.c_char_p, _IP, _I, _DP, _I, _I, _DP, C.c_size_t]),
    "qsv_ops_size": (_I, [_P]),
    "qsv_apply_ops": (_I, [_P, _P, _I]),
    "qsv_ops_plan_sweeps": (_I, [_P, _I, _I, _I, _I64P, _I64P, _I64P, _IP]),
    "qsv_ops_plan_work": (_I, [_P, _I, _I, _DP, _I64P, _I64P]),
    "qsv_last_apply_stats": (_I, [_P, _I64P, _I64P]),
    "qsv_expval_named": (_I, [_P, C.c_char_p, _IP, _I, _DP, _I, _DP]),
    "qsv_expval_matrix": (_I, [_P, _DP, _IP, _I, _DP]),
    "qsv_expval_pauli_words": (_I, [_P, _I, C.c_char_p, _IP, _IP, _DP, _DP, _DP]),
    "qsv_expval_csr": (_I, [_P, _P, _P, _DP, C.c_int64, _I, _DP]),
    "qsv_probs": (_I, [_P, _IP, _I, _DP]),
    "qsv_sample": (_I, [_P, _DP, C.c_int64, _U64P]),
    "qsv_inner_product": (_I, [_P, _P, _DP]),
    "qsv_axpy": (_I, [_DP, _P, _P]),
    "qsv_obs_named": (_I, [C.c_char_p, _IP, _I, _DP, _I, C.POINTER(_P)]),
    "qsv_obs_hermitian": (_I, [_DP, C.c_size_t, _IP, _I, C.POINTER(_P)]),
    "qsv_obs_tensor": (_I, [C.POINTER(_P), _I, C.POINTER(_P)]),
    "qsv_obs_hamiltonian": (_I, [_DP, C.POINTER(_P), _I, C.POINTER(_P)]),
    "qsv_obs_sparse": (_I, [_I64P, C.c_int64, _I64P, _DP, C.c_int64, C.POINTER(_P)]),
    "qsv_obs_destroy": (_I, [_P]),
    "qsv_obs_apply": (_I, [_P, _P]),
    "qsv_obs_expval": (_I, [_P, _P, _DP]),
    "qsv_adjoint_jacobian": (_I, [_P, _P, C.POINTER(_P), _I, _I64P, _I, _I, _DP]),
    "qsv_dist_unique_id": (_I, [_P]),
    "qsv_dist_init": (_I, [_P, _P, _I, _I]),
    "qsv_dist_finalize": (_I, [_P]),
    "qsv_dist_swap_bits": (_I, [_P, _I, _I, C.c_size_t]),
    "qsv_dist_apply_ops": (_I, [_P, _P, _I, C.c_size_t]),
    "qsv_dist_canonicalize": (_I, [_P, C.c_size_t]),
    "qsv_dist_set_lazy_map": (_I, [_P, _I]),
    "qsv_dist_qubit_map": (_I, [_P, _IP, _I]),
    "qsv_dist_expval_pauli_words": (_I, [_P, _I, C.c_char_p, _IP, _IP, _DP, _DP, _DP]),
    "qsv_dist_allreduce_f64": (_I, [_P, _DP, _I]),
    "qsv_dist_last_swap_stats": (_I, [_P, _U64P, C.POINTER(C.c_float)]),
    "qsv_dist_total_swap_stats": (_I, [_P, _IP, _U64P, C.POINTER(C.c_float), _I]),
    "qsv_dist_plan": (_I, [_P, _I, _I, _IP, _I, _IP, _IP]),
    "qsv_dist_fused_exchange_stats": (_I, [_P, _IP, _IP]),
    "qsv_dist_split_exchange_stats": (_I, [_P, _IP, _IP]),
    "qsv_dist_plan_from": (_I, [_P, _I, _I, _IP, _IP, _I, _IP, _IP]),
    "qsv_dist_uses_peer_access": (_I, [_P]),
    "qsv_dist_set_basis_state": (_I, [_P, C.c_uint64]),
    "qsv_dist_set_state_vector": (_I, [_P, _I64P, _P, C.c_size_t]),
    "qsv_dist_expval_named": (_I, [_P, C.c_char_p, _IP, _I, _DP, _I, _DP]),
    "qsv_dist_expval_matrix": (_I, [_P, _DP, _IP, _I, _DP]),
    "qsv_dist_expval_csr": (_I, [_P, _I64P, _I64P, _DP, C.c_int64, _DP]),
    "qsv_dist_probs": (_I, [_P, _IP, _I, _DP]),
    "qsv_dist_sample": (_I, [_P, _DP, C.c_int64, _U64P]),
    "qsv_dist_obs_expval": (_I, [_P, _P, _DP]),
    "qsv_dist_obs_apply": (_I, [_P, _P]),
    "qsv_dist_adjoint_jacobian": (_I, [_P, _P, C.POINTER(_P), _I, _I64P, _I, _I, _DP]),
    "qsv_dist_h2d": (_I, [_P, _P, C.c_size_t]),
    "qsv_dist_d2h": (_I, [_P, _P, C.c_size_t]),
    "qsv_dist_copy": (_I, [_P, _P]),
    "qsv_dist_barrier": (_I, [_P]),
    "qsv_dist_bcast_bytes": (_I, [_P, _P, C.c_size_t, _I]),
    "qsv_dist_scatter_host": (_I, [_P, _P, _P, C.c_size_t, _I]),
    "qsv_dist_nccl_version": (_I, [_IP]),
    "qsv_dist_rank": (_I, [_P]),
    "qsv_dist_world_size": (_I, [_P]),
    "qsv_dist_total_qubits": (_I, [_P]),
}


def lib() -> C.CDLL:
    """Load libqsv_b200.so (built in-tree by _build.py); never falls back to anything else."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise QsvError(
                f"{LIB_PATH} is missing: build it with `python -m pennylane_lightning_gpu_b200._build` "
                "(there is no CPU fallback)"
            )
        l = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def _check(status: int) -> None:
    if status != 0:
        raise QsvError(lib().qsv_last_error().decode())


def _ints(v: Sequence[int]):
    a = np.ascontiguousarray(np.asarray(list(v), dtype=np.int32).reshape(-1))
    return a, a.ctypes.data_as(_IP)


def _dbls(v) -> tuple[np.ndarray, object]:
    a = np.ascontiguousarray(np.asarray(v, dtype=np.float64).reshape(-1))
    return a, a.ctypes.data_as(_DP)


def _cmat(m) -> tuple[np.ndarray, object]:
    a = np.ascontiguousarray(np.asarray(m, dtype=np.complex128).reshape(-1))
    return a, a.ctypes.data_as(_DP)


def device_count() -> int:
    n = C.c_int(0)
    _check(lib().qsv_device_count(C.byref(n)))
    return n.value


def device_arch(device: int = 0) -> tuple[int, int]:
    a, b = C.c_int(0), C.c_int(0)
    _check(lib().qsv_device_arch(device, C.byref(a), C.byref(b)))
    return a.value, b.value


class Ops:
    """Recorded circuit = OpsData<SV> of the reference (bindings/Bindings.cpp:806-840)."""

    def __init__(self, ops: Sequence[dict] = ()):
        self._h = _P()
        _check(lib().qsv_ops_create(C.byref(self._h)))
        for op in ops:
            self.append(op["name"], op["wires"], op.get("params", ()), op.get("adjoint", False),
                        op.get("matrix"))

    def append(self, name, wires, params=(), inverse=False, matrix=None):
        wa, wp = _ints(wires)
        pa, pp = _dbls(params)
        if matrix is not None:
            ma, mp = _cmat(matrix)
            dim = 1 << len(wa)
        else:
            ma, mp, dim = None, None, 0
        _check(lib().qsv_ops_append(self._h, name.encode(), wp, len(wa), pp, len(pa), int(bool(inverse)),
                                    mp, dim))

    def __len__(self):
        return lib().qsv_ops_size(self._h)

    def plan_sweeps(self, n_qubits: int, dag: bool = True, low_bits: int = 0) -> dict:
        """Host-only: how apply_ops(fuse=True) would pack this circuit into HBM sweeps (no device needed)."""
        m, s, g, ok = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        _check(lib().qsv_ops_plan_sweeps(self._h, n_qubits, int(bool(dag)), low_bits, C.byref(m), C.byref(s),
                                         C.byref(g), C.byref(ok)))
        return {"gates_after_merge": m.value, "sweeps": s.value, "max_gates_per_sweep": g.value,
                "order_valid": bool(ok.value)}

    def plan_work(self, n_qubits: int, dtype=np.complex128) -> dict:
        """Host-only: arithmetic of apply_ops(fuse=True) on this circuit -- fused multiply-adds per amplitude, HBM sweeps
        and register passes of the programs the planner builds (no device needed)."""
        f, s, p = C.c_double(), C.c_int64(), C.c_int64()
        code = QSV_C128 if np.dtype(dtype) == np.complex128 else QSV_C64
        _check(lib().qsv_ops_plan_work(self._h, n_qubits, code, C.byref(f), C.byref(s), C.byref(p)))
        return {"fma_per_amplitude": f.value, "sweeps": s.value, "passes": p.value}

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.qsv_ops_destroy(self._h)
            self._h = None


class Observable:
    """Observable tree = ObservableGPU<T> hierarchy (algorithms/ObservablesGPU.hpp)."""

    def __init__(self, handle, keep=()):
        self._h = handle
        self._keep = keep

    @classmethod
    def named(cls, name, wires, params=()):
        h = _P()
        wa, wp = _ints(wires)
        pa, pp = _dbls(params)
        _check(lib().qsv_obs_named(name.encode(), wp, len(wa), pp, len(pa), C.byref(h)))
        return cls(h)

    @classmethod
    def hermitian(cls, matrix, wires):
        h = _P()
        wa, wp = _ints(wires)
        ma, mp = _cmat(matrix)
        _check(lib().qsv_obs_hermitian(mp, 1 << len(wa), wp, len(wa), C.byref(h)))
        return cls(h)

    @classmethod
    def tensor(cls, children: Sequence["Observable"]):
        h = _P()
        arr = (_P * len(children))(*[c._h for c in children])
        _check(lib().qsv_obs_tensor(arr, len(children), C.byref(h)))
        return cls(h)

    @classmethod
    def hamiltonian(cls, coeffs, children: Sequence["Observable"]):
        h = _P()
        ca, cp = _dbls(coeffs)
        arr = (_P * len(children))(*[c._h for c in children])
        _check(lib().qsv_obs_hamiltonian(cp, arr, len(children), C.byref(h)))
        return cls(h)

    @classmethod
    def sparse(cls, indptr, indices, data):
        h = _P()
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int64)
        va, vp = _cmat(data)
        _check(lib().qsv_obs_sparse(ip.ctypes.data_as(_I64P), len(ip), ix.ctypes.data_as(_I64P), vp, len(ix),
                                    C.byref(h)))
        return cls(h)

    @classmethod
    def from_tuple(cls, t):
        """Same plain-tuple encoding as the oracle uses (tests share their fixtures)."""
        kind = t[0]
        if kind == "Named":
            return cls.named(t[1], t[2], t[3] if len(t) > 3 else ())
        if kind == "Hermitian":
            return cls.hermitian(t[1], t[2])
        if kind == "TensorProd":
            return cls.tensor([cls.from_tuple(x) for x in t[1]])
        if kind == "Hamiltonian":
            return cls.hamiltonian(t[1], [cls.from_tuple(x) for x in t[2]])
        if kind == "Sparse":
            return cls.sparse(t[1], t[2], t[3])
        raise ValueError(kind)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.qsv_obs_destroy(self._h)
            self._h = None


class StateVector:
    """One 2^n amplitude array on one GPU (StateVectorCudaManaged<T> analogue over the C ABI)."""

    def __init__(self, n_qubits: int, dtype=np.complex128, device: int = 0, *, external_ptr: int | None = None,
                 stream: int | None = None):
        self.np_dtype = np.dtype(dtype)
        if self.np_dtype not in (np.dtype(np.complex64), np.dtype(np.complex128)):
            raise TypeError("dtype must be complex64 or complex128")
        code = QSV_C128 if self.np_dtype == np.complex128 else QSV_C64
        self._h = _P()
        _check(lib().qsv_create_external(n_qubits, code, device, _P(external_ptr) if external_ptr else None,
                                         _P(stream) if stream else None, C.byref(self._h)))
        self.n = n_qubits

    # -- lifetime / copies -------------------------------------------------------------------
    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.qsv_destroy(self._h)
            self._h = None

    @property
    def data_ptr(self) -> int:
        return lib().qsv_data_ptr(self._h)

    def set_stream(self, stream: int | None):
        _check(lib().qsv_set_stream(self._h, _P(stream) if stream else None))

    def synchronize(self):
        _check(lib().qsv_synchronize(self._h))

    def set_basis_state(self, index: int = 0):
        _check(lib().qsv_set_basis_state(self._h, index))

    def set_state_vector(self, indices, values):
        ia = np.ascontiguousarray(indices, dtype=np.int64)
        va = np.ascontiguousarray(values, dtype=self.np_dtype)
        _check(lib().qsv_set_state_vector(self._h, ia.ctypes.data_as(_I64P), va.ctypes.data_as(_P), len(ia)))

    def h2d(self, state):
        a = np.ascontiguousarray(state, dtype=self.np_dtype).reshape(-1)
        _check(lib().qsv_h2d(self._h, a.ctypes.data_as(_P), a.size))

    def d2h(self, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(1 << self.n, dtype=self.np_dtype)
        _check(lib().qsv_d2h(self._h, out.ctypes.data_as(_P), out.size))
        return out

    def copy_from(self, other: "StateVector"):
        _check(lib().qsv_d2d(self._h, other._h))

    # -- gates -------------------------------------------------------------------------------
    def apply(self, name: str, wires, params=(), adjoint=False, matrix=None):
        """applyOperation(name, wires, adjoint, params, matrix) (Managed.hpp:198-247)."""
        wa, wp = _ints(wires)
        if matrix is not None and name not in GATE_NAMES:
            ma, mp = _cmat(matrix)
            _check(lib().qsv_apply_matrix(self._h, mp, None, 0, wp, len(wa), int(bool(adjoint))))
            return
        pa, pp = _dbls(params)
        _check(lib().qsv_apply_named(self._h, name.encode(), wp, len(wa), int(bool(adjoint)), pp, len(pa)))

    def apply_matrix(self, matrix, wires, ctrls=(), adjoint=False):
        wa, wp = _ints(wires)
        ca, cp = _ints(ctrls)
        ma, mp = _cmat(matrix)
        _check(lib().qsv_apply_matrix(self._h, mp, cp, len(ca), wp, len(wa), int(bool(adjoint))))

    def apply_generator(self, name, wires, adjoint=False) -> float:
        wa, wp = _ints(wires)
        s = C.c_double(0)
        _check(lib().qsv_apply_generator(self._h, name.encode(), wp, len(wa), int(bool(adjoint)), C.byref(s)))
        return s.value

    def apply_ops(self, ops: Ops, fuse: bool = True):
        _check(lib().qsv_apply_ops(self._h, ops._h, int(bool(fuse))))

    def last_apply_stats(self) -> tuple[int, int]:
        a, b = C.c_int64(0), C.c_int64(0)
        _check(lib().qsv_last_apply_stats(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    # -- measurements ------------------------------------------------------------------------
    def expval_named(self, name, wires, params=()) -> complex:
        wa, wp = _ints(wires)
        pa, pp = _dbls(params)
        out = (C.c_double * 2)()
        _check(lib().qsv_expval_named(self._h, name.encode(), wp, len(wa), pp, len(pa), out))
        return complex(out[0], out[1])

    def expval_matrix(self, matrix, wires) -> complex:
        wa, wp = _ints(wires)
        ma, mp = _cmat(matrix)
        out = (C.c_double * 2)()
        _check(lib().qsv_expval_matrix(self._h, mp, wp, len(wa), out))
        return complex(out[0], out[1])

    def expval_pauli_words(self, words: Sequence[str], wires: Sequence[Sequence[int]], coeffs,
                           return_terms: bool = False):
        letters = "".join(words).encode()
        offs = np.zeros(len(words) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(w) for w in words])
        flat = [int(x) for ws in wires for x in ws]
        assert len(flat) == offs[-1], "each word needs one wire per letter"
        wa, wp = _ints(flat)
        ca, cp = _cmat(coeffs)
        terms = np.zeros(max(len(words), 1), dtype=np.float64)
        out = C.c_double(0)
        _check(lib().qsv_expval_pauli_words(self._h, len(words), letters, wp, offs.ctypes.data_as(_IP), cp,
                                            terms.ctypes.data_as(_DP), C.byref(out)))
        return (out.value, terms[: len(words)]) if return_terms else out.value

    def expval_csr(self, indptr, indices, data) -> float:
        idt = np.int64 if self.np_dtype == np.complex128 else np.int32  # Bindings.cpp:108-116
        ip = np.ascontiguousarray(indptr, dtype=idt)
        ix = np.ascontiguousarray(indices, dtype=idt)
        va, vp = _cmat(data)
        out = C.c_double(0)
        _check(lib().qsv_expval_csr(self._h, ip.ctypes.data_as(_P), ix.ctypes.data_as(_P), vp, len(ix),
                                    ip.itemsize, C.byref(out)))
        return out.value

    def expval(self, obs: Observable) -> float:
        out = C.c_double(0)
        _check(lib().qsv_obs_expval(obs._h, self._h, C.byref(out)))
        return out.value

    def apply_observable(self, obs: Observable):
        _check(lib().qsv_obs_apply(obs._h, self._h))

    def probs(self, wires) -> np.ndarray:
        """cuStateVec bit order like the reference: first listed wire = LSB of the output index."""
        wa, wp = _ints(wires)
        out = np.zeros(1 << len(wa), dtype=np.float64)
        _check(lib().qsv_probs(self._h, wp, len(wa), out.ctypes.data_as(_DP)))
        return out

    def sample(self, uniforms) -> np.ndarray:
        ua, up = _dbls(uniforms)
        out = np.zeros((len(ua), self.n), dtype=np.uint64)
        _check(lib().qsv_sample(self._h, up, len(ua), out.ctypes.data_as(_U64P)))
        return out

    def inner_product(self, other: "StateVector") -> complex:
        """<self|other>"""
        out = (C.c_double * 2)()
        _check(lib().qsv_inner_product(self._h, other._h, out))
        return complex(out[0], out[1])

    def axpy(self, alpha: complex, x: "StateVector"):
        a = (C.c_double * 2)(complex(alpha).real, complex(alpha).imag)
        _check(lib().qsv_axpy(a, x._h, self._h))

    def adjoint_jacobian(self, ops: Ops, observables: Sequence[Observable], trainable: Sequence[int],
                         apply_operations: bool = False) -> np.ndarray:
        tp = np.ascontiguousarray(list(trainable), dtype=np.int64)
        jac = np.zeros((len(observables), len(tp)), dtype=np.float64)
        arr = (_P * max(len(observables), 1))(*[o._h for o in observables])
        _check(lib().qsv_adjoint_jacobian(self._h, ops._h, arr, len(observables), tp.ctypes.data_as(_I64P),
                                          len(tp), int(bool(apply_operations)), jac.ctypes.data_as(_DP)))
        return jac


GATE_NAMES = frozenset([
    "Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "T", "RX", "RY", "RZ", "PhaseShift", "Rot",
    "CNOT", "CY", "CZ", "SWAP", "IsingXX", "IsingYY", "IsingZZ", "CRX", "CRY", "CRZ", "CRot",
    "ControlledPhaseShift", "SingleExcitation", "SingleExcitationMinus", "SingleExcitationPlus", "Toffoli",
    "CSWAP", "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus", "MultiRZ",
])
