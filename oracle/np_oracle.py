"""NumPy oracle for the lightning.gpu hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement of what the reference computes on the path
gate-apply -> measurements -> adjoint Jacobian.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product path (``pennylane_lightning_gpu_b200``)
never does and fails loudly when its CUDA library is missing.

Where the arithmetic lives.  The reference forwards every numerical operation to
closed-source libraries that are NOT under /root/reference (cuStateVec
``custatevec-cu11`` unpinned in requirements.txt:4, cuSPARSE/cuBLAS from the
CUDA toolkit) and compares itself against ``pennylane-lightning>=0.30``
(requirements.txt:7), which is not installed here either.  So this oracle
restates the *published* algorithm (state-vector simulation, matrix on target
wires, adjoint method of arXiv:2009.02823) and is anchored on the reference's own
call sites and golden vectors:

* gate definitions           src/simulator/cuGates_host.hpp:27-1326
* control/target split       src/simulator/StateVectorCudaManaged.hpp:321-560
* dispatch, Rot/adjoint      src/simulator/StateVectorCudaManaged.hpp:198-247
* Pauli rotation convention  src/simulator/StateVectorCudaManaged.hpp:1339-1386
* generators + scaling       src/algorithms/GateGenerators.hpp:58-321,
                             src/algorithms/AdjointDiffGPU.hpp:58-114
* adjoint loop               src/algorithms/AdjointDiffGPU.hpp:499-596, 132-162
* observables                src/algorithms/ObservablesGPU.hpp:56-587
* expval / probs / samples   src/simulator/StateVectorCudaManaged.hpp:702-1148

Pinned by tests/test_oracle_golden.py against tests/golden/reference_kats.json
(machine-extracted / transcribed from the reference's own tests, see
tests/golden/make_golden.py).  Sample parity is pinned by definition only: the
reference has no seed (StateVectorCudaManaged.hpp:1003), see ``sample``.

Conventions: wire w <-> index bit n-1-w (wire 0 is the MSB); gate matrices are
row-major in PennyLane wire order (first listed wire = most significant matrix
bit); ``adjoint=True`` applies M^dagger.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np

# ----------------------------------------------------------------------------
# Gate matrices (PennyLane wire order, row-major)
# ----------------------------------------------------------------------------

_I2 = np.eye(2, dtype=np.complex128)
_X = np.array([[0, 1], [1, 0]], dtype=np.complex128)
_Y = np.array([[0, -1j], [1j, 0]], dtype=np.complex128)
_Z = np.array([[1, 0], [0, -1]], dtype=np.complex128)
_H = np.array([[1, 1], [1, -1]], dtype=np.complex128) / math.sqrt(2.0)
_S = np.diag([1, 1j]).astype(np.complex128)
_T = np.diag([1, np.exp(1j * math.pi / 4)]).astype(np.complex128)
_P1 = np.diag([0, 1]).astype(np.complex128)  # |1><1|, GateGenerators.hpp:32-37

PAULI = {"I": _I2, "X": _X, "Y": _Y, "Z": _Z}


def _kron(*ms):
    out = np.array([[1.0 + 0j]])
    for m in ms:
        out = np.kron(out, m)
    return out


def _controlled(u: np.ndarray, n_ctrl: int = 1) -> np.ndarray:
    """Block matrix diag(1, ..., 1, U): controls are the leading wires."""
    d = u.shape[0]
    dim = d << n_ctrl
    out = np.eye(dim, dtype=np.complex128)
    out[dim - d:, dim - d:] = u
    return out


def rx(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -1j * s], [-1j * s, c]], dtype=np.complex128)


def ry(t):
    c, s = math.cos(t / 2), math.sin(t / 2)
    return np.array([[c, -s], [s, c]], dtype=np.complex128)


def rz(t):
    return np.diag([np.exp(-0.5j * t), np.exp(0.5j * t)]).astype(np.complex128)


def phase_shift(t):
    return np.diag([1, np.exp(1j * t)]).astype(np.complex128)


def rot(phi, theta, omega):
    """Rot(phi, theta, omega) = RZ(omega) RY(theta) RZ(phi); cuGates_host.hpp:376-405."""
    return rz(omega) @ ry(theta) @ rz(phi)


def pauli_rot(word: str, t: float) -> np.ndarray:
    """exp(-i t/2 P); StateVectorCudaManaged.hpp:1372 passes -t/2 to exp(i a P)."""
    p = _kron(*[PAULI[c] for c in word])
    return math.cos(t / 2) * np.eye(p.shape[0]) - 1j * math.sin(t / 2) * p


def single_excitation(t, phase=0):
    """phase = 0: SingleExcitation, -1: ...Minus, +1: ...Plus; cuGates_host.hpp:640-816
    (stored there bit-reversed and applied with un-reversed wires)."""
    c, s = math.cos(t / 2), math.sin(t / 2)
    e = np.exp(1j * phase * t / 2)
    m = np.zeros((4, 4), dtype=np.complex128)
    m[0, 0] = m[3, 3] = e
    m[1, 1] = m[2, 2] = c
    m[1, 2] = -s
    m[2, 1] = s
    return m


def double_excitation(t, phase=0):
    """cuGates_host.hpp:871-1110: rotation in span{|0011>, |1100>}."""
    c, s = math.cos(t / 2), math.sin(t / 2)
    e = np.exp(1j * phase * t / 2)
    m = np.eye(16, dtype=np.complex128) * e
    m[3, 3] = m[12, 12] = c
    m[3, 12] = -s
    m[12, 3] = s
    return m


_SWAP = np.array([[1, 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]], dtype=np.complex128)

# name -> (number of wires or None for variadic, number of parameters)
GATE_ARITY = {
    "Identity": (1, 0), "PauliX": (1, 0), "PauliY": (1, 0), "PauliZ": (1, 0),
    "Hadamard": (1, 0), "S": (1, 0), "T": (1, 0),
    "RX": (1, 1), "RY": (1, 1), "RZ": (1, 1), "PhaseShift": (1, 1), "Rot": (1, 3),
    "CNOT": (2, 0), "CY": (2, 0), "CZ": (2, 0), "SWAP": (2, 0),
    "IsingXX": (2, 1), "IsingYY": (2, 1), "IsingZZ": (2, 1),
    "CRX": (2, 1), "CRY": (2, 1), "CRZ": (2, 1), "CRot": (2, 3),
    "ControlledPhaseShift": (2, 1),
    "SingleExcitation": (2, 1), "SingleExcitationMinus": (2, 1), "SingleExcitationPlus": (2, 1),
    "Toffoli": (3, 0), "CSWAP": (3, 0),
    "DoubleExcitation": (4, 1), "DoubleExcitationMinus": (4, 1), "DoubleExcitationPlus": (4, 1),
    "MultiRZ": (None, 1),
}


def gate_matrix(name: str, params: Sequence[float] = (), n_wires: int | None = None) -> np.ndarray:
    """Full (controls included) matrix of a named gate in PennyLane wire order."""
    p = list(params)
    if name == "Identity":
        return _I2.copy()
    if name == "PauliX":
        return _X.copy()
    if name == "PauliY":
        return _Y.copy()
    if name == "PauliZ":
        return _Z.copy()
    if name == "Hadamard":
        return _H.copy()
    if name == "S":
        return _S.copy()
    if name == "T":
        return _T.copy()
    if name == "RX":
        return rx(p[0])
    if name == "RY":
        return ry(p[0])
    if name == "RZ":
        return rz(p[0])
    if name == "PhaseShift":
        return phase_shift(p[0])
    if name == "Rot":
        return rot(*p[:3])
    if name == "CNOT":
        return _controlled(_X)
    if name == "CY":
        return _controlled(_Y)
    if name == "CZ":
        return _controlled(_Z)
    if name == "SWAP":
        return _SWAP.copy()
    if name == "IsingXX":
        return pauli_rot("XX", p[0])
    if name == "IsingYY":
        return pauli_rot("YY", p[0])
    if name == "IsingZZ":
        return pauli_rot("ZZ", p[0])
    if name == "CRX":
        return _controlled(rx(p[0]))
    if name == "CRY":
        return _controlled(ry(p[0]))
    if name == "CRZ":
        return _controlled(rz(p[0]))
    if name == "CRot":
        return _controlled(rot(*p[:3]))
    if name == "ControlledPhaseShift":
        return _controlled(phase_shift(p[0]))
    if name == "SingleExcitation":
        return single_excitation(p[0], 0)
    if name == "SingleExcitationMinus":
        return single_excitation(p[0], -1)
    if name == "SingleExcitationPlus":
        return single_excitation(p[0], +1)
    if name == "Toffoli":
        return _controlled(_X, 2)
    if name == "CSWAP":
        return _controlled(_SWAP)
    if name == "DoubleExcitation":
        return double_excitation(p[0], 0)
    if name == "DoubleExcitationMinus":
        return double_excitation(p[0], -1)
    if name == "DoubleExcitationPlus":
        return double_excitation(p[0], +1)
    if name == "MultiRZ":
        assert n_wires is not None
        return pauli_rot("Z" * n_wires, p[0])
    raise ValueError(f"Currently unsupported gate: {name}")


# ----------------------------------------------------------------------------
# State evolution
# ----------------------------------------------------------------------------

def apply_matrix(state: np.ndarray, mat: np.ndarray, wires: Sequence[int]) -> np.ndarray:
    """Return (mat on ``wires``) |state>; wire 0 = MSB, mat row-major in wire order."""
    n = int(math.log2(state.size))
    k = len(wires)
    mat = np.asarray(mat, dtype=np.complex128).reshape(1 << k, 1 << k)
    psi = state.reshape([2] * n)
    psi = np.moveaxis(psi, list(wires), list(range(k)))
    shp = psi.shape
    psi = mat @ psi.reshape(1 << k, -1)
    psi = np.moveaxis(psi.reshape(shp), list(range(k)), list(wires))
    return np.ascontiguousarray(psi).reshape(-1)


def apply_op(state, name, wires, params=(), adjoint=False, matrix=None):
    """StateVectorCudaManaged::applyOperation (Managed.hpp:198-247)."""
    if name == "Identity":
        return state
    if matrix is not None and name not in GATE_ARITY:
        m = np.asarray(matrix, dtype=np.complex128).reshape(1 << len(wires), -1)
    else:
        m = gate_matrix(name, params, len(wires))
    if adjoint:
        m = m.conj().T
    return apply_matrix(state, m, wires)


def apply_ops(state, ops):
    """ops: iterable of dicts {name, wires, params, adjoint, matrix}."""
    for op in ops:
        state = apply_op(state, op["name"], op["wires"], op.get("params", ()),
                         op.get("adjoint", False), op.get("matrix"))
    return state


def basis_state(n: int, index: int = 0, dtype=np.complex128) -> np.ndarray:
    s = np.zeros(1 << n, dtype=dtype)
    s[index] = 1
    return s


def set_state_vector(n, indices, values, dtype=np.complex128):
    """Managed.hpp:164-185 / initSV.cu:64-71: zero, then sv[idx[i]] = val[i]."""
    s = np.zeros(1 << n, dtype=dtype)
    s[np.asarray(indices)] = np.asarray(values)
    return s


# ----------------------------------------------------------------------------
# Generators (GateGenerators.hpp) and scaling factors (AdjointDiffGPU.hpp:96-114)
# ----------------------------------------------------------------------------

def generator(name: str, n_wires: int):
    """-> (matrix on the op's wires in PennyLane order, scaling factor)."""
    if name in ("RX", "RY", "RZ"):
        return PAULI[name[1]].copy(), -0.5
    if name in ("IsingXX", "IsingYY", "IsingZZ"):
        c = name[-1]
        return _kron(PAULI[c], PAULI[c]), -0.5
    if name == "PhaseShift":
        return _P1.copy(), 1.0
    if name == "ControlledPhaseShift":
        return _kron(_P1, _P1), 1.0  # P_1111, GateGenerators.hpp:38-44, 208-213
    if name in ("CRX", "CRY", "CRZ"):
        return _kron(_P1, PAULI[name[2]]), -0.5  # GateGenerators.hpp:158-196
    if name.startswith("SingleExcitation"):
        g = np.zeros((4, 4), dtype=np.complex128)
        g[1, 2] = -1j
        g[2, 1] = 1j
        d = {"SingleExcitation": 0, "SingleExcitationMinus": 1, "SingleExcitationPlus": -1}[name]
        g[0, 0] = g[3, 3] = d
        return g, -0.5
    if name.startswith("DoubleExcitation"):
        d = {"DoubleExcitation": 0, "DoubleExcitationMinus": 1, "DoubleExcitationPlus": -1}[name]
        g = np.eye(16, dtype=np.complex128) * d
        g[3, 3] = g[12, 12] = 0
        g[3, 12] = -1j
        g[12, 3] = 1j
        return g, -0.5
    if name == "MultiRZ":
        return _kron(*[_Z] * n_wires), -0.5  # Managed.hpp:680-688
    raise ValueError(f"no generator for {name}")


# ----------------------------------------------------------------------------
# Observables (ObservablesGPU.hpp).  Plain tuples keep the oracle dependency-free:
#   ("Named", name, wires[, params])   ("Hermitian", matrix, wires)
#   ("TensorProd", [obs, ...])         ("Hamiltonian", coeffs, [obs, ...])
#   ("Sparse", indptr, indices, data)
# ----------------------------------------------------------------------------

def apply_observable(state, obs):
    kind = obs[0]
    if kind == "Named":
        params = obs[3] if len(obs) > 3 else ()
        return apply_op(state, obs[1], obs[2], params)
    if kind == "Hermitian":
        return apply_matrix(state, np.asarray(obs[1]), obs[2])
    if kind == "TensorProd":
        for o in obs[1]:
            state = apply_observable(state, o)
        return state
    if kind == "Hamiltonian":
        out = np.zeros_like(state)
        for c, o in zip(obs[1], obs[2]):
            out = out + c * apply_observable(state, o)
        return out
    if kind == "Sparse":
        return csr_matvec(obs[1], obs[2], obs[3], state)
    raise ValueError(kind)


def csr_matvec(indptr, indices, data, x):
    indptr = np.asarray(indptr)
    indices = np.asarray(indices)
    data = np.asarray(data)
    prod = data * x[indices]
    y = np.add.reduceat(np.concatenate([prod, [0]]), np.minimum(indptr[:-1], prod.size))
    y[np.diff(indptr) == 0] = 0
    return y


# ----------------------------------------------------------------------------
# Measurements (Managed.hpp:702-1148)
# ----------------------------------------------------------------------------

def expval_matrix(state, mat, wires) -> complex:
    """<psi|M|psi>, complex (Managed.hpp:755-780, KAT NonParam.cpp:864-894)."""
    return complex(np.vdot(state, apply_matrix(state, mat, wires)))


def expval_named(state, name, wires, params=()) -> float:
    return float(np.vdot(state, apply_op(state, name, wires, params)).real)


def pauli_word_matrix_free(state, word: str, wires: Sequence[int]):
    """P|psi> through x/z masks; independent of apply_matrix on purpose."""
    n = int(math.log2(state.size))
    idx = np.arange(state.size, dtype=np.int64)
    xmask = zmask = 0
    ny = 0
    for c, w in zip(word, wires):
        b = 1 << (n - 1 - w)
        if c in "XY":
            xmask |= b
        if c in "ZY":
            zmask |= b
        if c == "Y":
            ny += 1
    # (P psi)_i = i^{ny} (-1)^{popc((i ^ x) & z)} psi_{i ^ x}
    src = idx ^ xmask
    par = np.zeros(state.size, dtype=np.int64)
    t = src & zmask
    while np.any(t):
        par ^= t & 1
        t >>= 1
    return (1j ** ny) * np.where(par == 1, -1.0, 1.0) * state[src]


def expval_pauli_words(state, words, tgts, coeffs) -> float:
    """getExpectationValuePauliWords (Managed.hpp:1071-1148): per-term value is a
    double; for complex64 it is cast to float before the coefficient dot (:1137-1146)."""
    single = state.dtype == np.complex64
    psi = state.astype(np.complex128)
    tot = 0j
    for w, t, c in zip(words, tgts, coeffs):
        e = float(np.vdot(psi, pauli_word_matrix_free(psi, w, t)).real)
        if single:
            e = float(np.float32(e))
        tot += e * complex(c)
    return float(tot.real)


def expval_csr(state, indptr, indices, data) -> float:
    """getExpectationValueOnSparseSpMV (Managed.hpp:795-922): Re <psi| H psi>."""
    return float(np.vdot(state, csr_matvec(indptr, indices, data, state)).real)


def expval_obs(state, obs) -> float:
    return float(np.vdot(state, apply_observable(state, obs)).real)


def probs(state, wires: Sequence[int]) -> np.ndarray:
    """Marginal |psi|^2 in PennyLane order (first wire = MSB of the output index).
    The reference returns cuStateVec bit order (Managed.hpp:931-970) and the device
    re-transposes it (lightning_gpu.py:920-924); ``probs_custatevec_order`` gives that."""
    n = int(math.log2(state.size))
    p = (np.abs(state.astype(np.complex128)) ** 2).reshape([2] * n)
    other = tuple(i for i in range(n) if i not in wires)
    p = p.sum(axis=other) if other else p
    kept = [w for w in range(n) if w in wires]
    p = np.transpose(p, [kept.index(w) for w in wires])
    return np.ascontiguousarray(p).reshape(-1)


def probs_custatevec_order(state, wires):
    """First listed wire = LSB of the output index (Managed.hpp:949-967)."""
    return probs(state, list(wires)[::-1])


def sample(state, shots: int, seed: int) -> np.ndarray:
    """(shots, n) array of 0/1, column j = wire j (Managed.hpp:1038-1055).

    The reference draws from an unseeded mt19937 (Managed.hpp:1003) so sample parity is
    pinned by DEFINITION here: u_i = numpy default_rng(seed).random(shots) (float64),
    index_i = searchsorted(cumsum(|psi|^2, float64), u_i, side='right') clipped."""
    n = int(math.log2(state.size))
    u = np.random.default_rng(seed).random(shots)
    cdf = np.cumsum(np.abs(state.astype(np.complex128)) ** 2)
    idx = np.minimum(np.searchsorted(cdf, u * cdf[-1], side="right"), state.size - 1)
    bits = (idx[:, None] >> (n - 1 - np.arange(n))[None, :]) & 1
    return bits.astype(np.uint64)


# ----------------------------------------------------------------------------
# Adjoint Jacobian (AdjointDiffGPU.hpp:499-596)
# ----------------------------------------------------------------------------

_STATE_PREPS = ("QubitStateVector", "StatePrep", "BasisState")


def adjoint_jacobian(state, ops, observables, trainable, apply_operations=False):
    """ops: list of dicts {name, wires, params, adjoint(=inverse)[, matrix]}.
    ``trainable`` indexes parametric ops only (hasParams = non-empty params).
    ``state`` is the final state unless ``apply_operations``.
    Returns jac[n_obs][len(trainable)]."""
    if len(trainable) == 0:
        raise ValueError("No trainable parameters provided.")
    lam = np.asarray(state, dtype=np.complex128).copy()
    if apply_operations:
        lam = apply_ops(lam, ops)
    bras = [apply_observable(lam, o) for o in observables]
    jac = np.zeros((len(observables), len(trainable)))
    n_par_ops = sum(1 for op in ops if len(op.get("params", ())) > 0)
    tp = list(trainable)
    tp_pos = len(tp) - 1
    cur = n_par_ops - 1
    for op in reversed(ops):
        params = op.get("params", ())
        if len(params) > 1:
            raise ValueError("The operation is not supported using the adjoint differentiation method")
        if op["name"] in _STATE_PREPS:
            continue
        if tp_pos < 0:
            break
        inv = bool(op.get("adjoint", False))
        mu = lam
        lam = apply_op(lam, op["name"], op["wires"], params, not inv, op.get("matrix"))
        if len(params) > 0:
            if cur == tp[tp_pos]:
                g, s = generator(op["name"], len(op["wires"]))
                gmu = apply_matrix(mu, g, op["wires"])
                s = s * (-1.0 if inv else 1.0)
                for i, b in enumerate(bras):
                    jac[i, tp_pos] = -2.0 * s * np.vdot(b, gmu).imag
                tp_pos -= 1
            cur -= 1
        bras = [apply_op(b, op["name"], op["wires"], params, not inv, op.get("matrix")) for b in bras]
    return jac


# ----------------------------------------------------------------------------
# Circuit templates used by BASELINE.json configs (PennyLane definitions)
# ----------------------------------------------------------------------------

def strongly_entangling_layers(weights: np.ndarray, expand_rot: bool = True):
    """StronglyEntanglingLayers(weights[L, n, 3]); ranges r_l = (l mod (n-1)) + 1.
    Used by the reference only at tests/test_comparison.py:246-253; the definition is
    PennyLane's.  ``expand_rot`` applies _serialize.py:311-312 (Rot -> RZ RY RZ)."""
    L, n, _ = weights.shape
    ops = []
    for l in range(L):
        for i in range(n):
            phi, theta, omega = (float(x) for x in weights[l, i])
            if expand_rot:
                ops.append({"name": "RZ", "wires": [i], "params": [phi]})
                ops.append({"name": "RY", "wires": [i], "params": [theta]})
                ops.append({"name": "RZ", "wires": [i], "params": [omega]})
            else:
                ops.append({"name": "Rot", "wires": [i], "params": [phi, theta, omega]})
        if n > 1:
            r = (l % (n - 1)) + 1
            for i in range(n):
                ops.append({"name": "CNOT", "wires": [i, (i + r) % n], "params": []})
    return ops
