"""Driver for oracle/lq_port.c  --  TEST INFRASTRUCTURE / CPU BASELINE ONLY.

The C file restates lightning.qubit's in-place OpenMP kernels; this module adds the op-level
dispatch (which kernel for which gate), the Pauli-word / Hamiltonian measurements and the adjoint
loop (same structure as algorithms/AdjointDiffGPU.hpp:499-596 of the reference, which is also how
lightning.qubit's AdjointJacobian is organised), all on top of np_oracle's gate definitions.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from . import np_oracle as orc

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "lq_port.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liblq_port.so")

_lib = None
_CP = C.c_void_p
_IP = C.POINTER(C.c_int)
_DP = C.POINTER(C.c_double)


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC", SRC, "-o", LIB, "-lm"],
                       check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        try:
            _lib = C.CDLL(build())
        except OSError:
            # built on a different host CPU (-march=native): rebuild for this one
            _lib = C.CDLL(build(force=True))
        l = _lib
        l.lq_num_threads.restype = C.c_int
        l.lq_set_num_threads.argtypes = [C.c_int]
        l.lq_apply_1q.argtypes = [_CP, C.c_int, _CP, C.c_int]
        l.lq_apply_dense.argtypes = [_CP, C.c_int, _CP, _IP, C.c_int, C.c_uint64]
        l.lq_apply_diag.argtypes = [_CP, C.c_int, _CP, _IP, C.c_int, C.c_uint64]
        l.lq_apply_parity.argtypes = [_CP, C.c_int, C.c_uint64, _CP, C.c_uint64]
        l.lq_apply_x.argtypes = [_CP, C.c_int, C.c_int, C.c_uint64]
        l.lq_copy.argtypes = [_CP, _CP, C.c_int]
        l.lq_set_basis.argtypes = [_CP, C.c_int, C.c_uint64]
        l.lq_zero.argtypes = [_CP, C.c_int]
        l.lq_inner.argtypes = [_CP, _CP, C.c_int, _DP]
        l.lq_bra_pauli_ket.argtypes = [_CP, _CP, C.c_int, C.c_uint64, C.c_uint64, C.c_int, _DP]
        l.lq_pauli_axpy.argtypes = [_CP, _CP, C.c_int, C.c_uint64, C.c_uint64, C.c_double, C.c_double]
    return _lib


def num_threads() -> int:
    return lib().lq_num_threads()


def set_num_threads(t: int) -> None:
    lib().lq_set_num_threads(int(t))


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_CP)


_CTRL = {"CNOT": 1, "CY": 1, "CZ": 1, "CRX": 1, "CRY": 1, "CRZ": 1, "CRot": 1, "ControlledPhaseShift": 1,
         "Toffoli": 2, "CSWAP": 1}
_CORE = {"CNOT": "PauliX", "CY": "PauliY", "CZ": "PauliZ", "CRX": "RX", "CRY": "RY", "CRZ": "RZ", "CRot": "Rot",
         "ControlledPhaseShift": "PhaseShift", "Toffoli": "PauliX", "CSWAP": "SWAP"}
_DIAG = {"PauliZ", "S", "T", "RZ", "PhaseShift", "IsingZZ", "MultiRZ"}


class LQState:
    """In-place complex128 state vector driven through liblq_port.so."""

    def __init__(self, n: int, state: np.ndarray | None = None):
        self.n = n
        self.sv = np.empty(1 << n, dtype=np.complex128)
        if state is None:
            lib().lq_set_basis(_ptr(self.sv), n, 0)
        else:
            self.sv[:] = state

    def copy(self) -> "LQState":
        out = LQState.__new__(LQState)
        out.n = self.n
        out.sv = np.empty_like(self.sv)
        lib().lq_copy(_ptr(out.sv), _ptr(self.sv), self.n)
        return out

    def bit(self, w: int) -> int:
        return self.n - 1 - w

    def apply_matrix(self, mat: np.ndarray, tgt_wires, ctrl_wires=()):
        mat = np.ascontiguousarray(mat, dtype=np.complex128)
        tb = np.asarray([self.bit(w) for w in tgt_wires], dtype=np.int32)
        cm = 0
        for w in ctrl_wires:
            cm |= 1 << self.bit(w)
        lib().lq_apply_dense(_ptr(self.sv), self.n, _ptr(mat), tb.ctypes.data_as(_IP), len(tb), cm)

    def apply_op(self, name, wires, params=(), adjoint=False, matrix=None):
        if name == "Identity":
            return
        wires = list(wires)
        if name not in orc.GATE_ARITY:
            m = np.asarray(matrix, dtype=np.complex128).reshape(1 << len(wires), -1)
            self.apply_matrix(m.conj().T if adjoint else m, wires)
            return
        nc = _CTRL.get(name, 0)
        core = _CORE.get(name, name)
        ctrls, tgts = wires[:nc], wires[nc:]
        cm = 0
        for w in ctrls:
            cm |= 1 << self.bit(w)
        if core == "PauliX":
            lib().lq_apply_x(_ptr(self.sv), self.n, self.bit(tgts[0]), cm)
            return
        m = orc.gate_matrix(core, params, len(tgts))
        if adjoint:
            m = m.conj().T
        if core in _DIAG:
            if core in ("IsingZZ", "MultiRZ"):
                z = 0
                for w in tgts:
                    z |= 1 << self.bit(w)
                eo = np.array([m[0, 0], m[1, 1]], dtype=np.complex128)
                lib().lq_apply_parity(_ptr(self.sv), self.n, z, _ptr(eo), cm)
            else:
                d = np.ascontiguousarray(np.diag(m), dtype=np.complex128)
                tb = np.asarray([self.bit(w) for w in tgts], dtype=np.int32)
                lib().lq_apply_diag(_ptr(self.sv), self.n, _ptr(d), tb.ctypes.data_as(_IP), len(tb), cm)
            return
        self.apply_matrix(m, tgts, ctrls)

    def apply_ops(self, ops):
        for op in ops:
            self.apply_op(op["name"], op["wires"], op.get("params", ()), op.get("adjoint", False), op.get("matrix"))

    # -- measurements --------------------------------------------------------------------------
    def _masks(self, word, wires):
        x = z = ny = 0
        for c, w in zip(word, wires):
            b = 1 << self.bit(w)
            if c in "XY":
                x |= b
            if c in "ZY":
                z |= b
            if c == "Y":
                ny += 1
        return x, z, ny

    def expval_pauli_word(self, word, wires) -> float:
        x, z, ny = self._masks(word, wires)
        out = (C.c_double * 2)()
        lib().lq_bra_pauli_ket(_ptr(self.sv), _ptr(self.sv), self.n, x, z, ny, out)
        return out[0]

    def expval_pauli_words(self, words, wires, coeffs) -> float:
        return float(sum(c * self.expval_pauli_word(w, ws) for w, ws, c in zip(words, wires, coeffs)))

    def inner(self, other: "LQState") -> complex:
        out = (C.c_double * 2)()
        lib().lq_inner(_ptr(self.sv), _ptr(other.sv), self.n, out)
        return complex(out[0], out[1])

    def apply_pauli_hamiltonian(self, words, wires, coeffs) -> "LQState":
        out = LQState.__new__(LQState)
        out.n = self.n
        out.sv = np.empty_like(self.sv)
        lib().lq_zero(_ptr(out.sv), self.n)
        for w, ws, c in zip(words, wires, coeffs):
            x, z, ny = self._masks(w, ws)
            cc = complex(c) * (1j ** ny)
            lib().lq_pauli_axpy(_ptr(out.sv), _ptr(self.sv), self.n, x, z, cc.real, cc.imag)
        return out

    def apply_observable(self, obs) -> "LQState":
        """obs in np_oracle's tuple encoding -> new state O|self>."""
        kind = obs[0]
        if kind == "Hamiltonian":
            out = LQState.__new__(LQState)
            out.n = self.n
            out.sv = np.zeros_like(self.sv)
            for c, o in zip(obs[1], obs[2]):
                out.sv += c * self.apply_observable(o).sv
            return out
        out = self.copy()
        if kind == "Named":
            out.apply_op(obs[1], obs[2], obs[3] if len(obs) > 3 else ())
        elif kind == "Hermitian":
            out.apply_matrix(np.asarray(obs[1]), obs[2])
        elif kind == "TensorProd":
            for o in obs[1]:
                out = out.apply_observable(o)
        else:
            raise ValueError(kind)
        return out


def adjoint_jacobian(final: LQState, ops, observables, trainable) -> np.ndarray:
    """Reverse sweep with explicit mu, as lightning.qubit / AdjointDiffGPU.hpp:560-595 do."""
    lam = final.copy()
    bras = [lam.apply_observable(o) for o in observables]
    jac = np.zeros((len(observables), len(trainable)))
    n_par = sum(1 for op in ops if len(op.get("params", ())) > 0)
    tp = list(trainable)
    tp_pos = len(tp) - 1
    cur = n_par - 1
    for op in reversed(ops):
        params = op.get("params", ())
        if tp_pos < 0:
            break
        inv = bool(op.get("adjoint", False))
        mu = lam.copy()
        lam.apply_op(op["name"], op["wires"], params, not inv, op.get("matrix"))
        if len(params) > 0:
            if cur == tp[tp_pos]:
                g, s = orc.generator(op["name"], len(op["wires"]))
                mu.apply_matrix(g, op["wires"])
                s = s * (-1.0 if inv else 1.0)
                for i, b in enumerate(bras):
                    jac[i, tp_pos] = -2.0 * s * b.inner(mu).imag
                tp_pos -= 1
            cur -= 1
        for b in bras:
            b.apply_op(op["name"], op["wires"], params, not inv, op.get("matrix"))
    return jac
