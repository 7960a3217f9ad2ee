"""Synthetic workload generators for BASELINE.json's configs.  Pure Python /
NumPy; shared by bench.py, the parity tests and the fixture generator so that all three see the same
circuits.  Ops are plain dicts {name, wires, params[, adjoint, matrix]}."""
from __future__ import annotations

import math

import numpy as np


def haar_unitary(rng: np.random.Generator, dim: int) -> np.ndarray:
    z = rng.normal(size=(dim, dim)) + 1j * rng.normal(size=(dim, dim))
    q, r = np.linalg.qr(z)
    return q * (np.diag(r) / np.abs(np.diag(r)))


def strongly_entangling_layers(n: int, layers: int = 2, seed: int = 1337):
    """C1: PennyLane's StronglyEntanglingLayers with Rot expanded to RZ RY RZ (_serialize.py:311-312)."""
    w = np.random.default_rng(seed).uniform(0, 2 * math.pi, (layers, n, 3))
    ops = []
    for l in range(layers):
        for i in range(n):
            phi, theta, omega = (float(x) for x in w[l, i])
            ops.append({"name": "RZ", "wires": [i], "params": [phi]})
            ops.append({"name": "RY", "wires": [i], "params": [theta]})
            ops.append({"name": "RZ", "wires": [i], "params": [omega]})
        if n > 1:
            r = (l % (n - 1)) + 1
            for i in range(n):
                ops.append({"name": "CNOT", "wires": [i, (i + r) % n], "params": []})
    return ops, 3 * n * layers


def random_gate_circuit(n: int, n_gates: int = 200, seed: int = 2024):
    """C2(ii): gates drawn uniformly from {RX, RY, RZ, CNOT, CZ, QubitUnitary 1q, QubitUnitary 2q}."""
    rng = np.random.default_rng(seed)
    kinds = ["RX", "RY", "RZ", "CNOT", "CZ", "QubitUnitary1", "QubitUnitary2"]
    ops = []
    for _ in range(n_gates):
        k = kinds[int(rng.integers(len(kinds)))]
        if k in ("RX", "RY", "RZ"):
            ops.append({"name": k, "wires": [int(rng.integers(n))], "params": [float(rng.uniform(-math.pi, math.pi))]})
        elif k in ("CNOT", "CZ"):
            a, b = (int(x) for x in rng.choice(n, size=2, replace=False))
            ops.append({"name": k, "wires": [a, b], "params": []})
        else:
            nq = int(k[-1])
            wires = [int(x) for x in rng.choice(n, size=nq, replace=False)]
            ops.append({"name": "QubitUnitary", "wires": wires, "params": [], "matrix": haar_unitary(rng, 1 << nq)})
    return ops


def random_layer_circuit(n: int, layers: int = 4, seed: int = 99):
    """BASELINE config 5: per layer a random one-qubit rotation (RX / RY / RZ, angle uniform in
    (-pi, pi)) on every wire, then CNOTs on a random perfect matching of the wires (n odd: one wire sits out)."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(layers):
        for w in range(n):
            name = ("RX", "RY", "RZ")[int(rng.integers(3))]
            ops.append({"name": name, "wires": [w], "params": [float(rng.uniform(-math.pi, math.pi))]})
        perm = [int(x) for x in rng.permutation(n)]
        for i in range(0, n - 1, 2):
            ops.append({"name": "CNOT", "wires": [perm[i], perm[i + 1]], "params": []})
    return ops


def gate_bytes(op: dict, n: int, amp_bytes: int) -> int:
    """Algorithmic bytes of one gate sweep: 2 * B * N / 2^c, c = number of
    control wires, explicit or implicit (CNOT, CZ, Toffoli, CR*, ...)."""
    n_ctrl = {"CNOT": 1, "CY": 1, "CZ": 1, "CRX": 1, "CRY": 1, "CRZ": 1, "CRot": 1, "ControlledPhaseShift": 1,
              "Toffoli": 2, "CSWAP": 1}.get(op["name"], 0)
    return 2 * amp_bytes * (1 << n) >> n_ctrl


_DIAGONAL = {"Identity", "PauliZ", "S", "T", "PhaseShift", "RZ", "CZ", "CRZ", "ControlledPhaseShift", "IsingZZ", "MultiRZ"}
_N_CONTROLS = {"CNOT": 1, "CY": 1, "CRX": 1, "CRY": 1, "CRot": 1, "Toffoli": 2, "CSWAP": 1}


def gate_bit_masks(op: dict, n: int) -> tuple[int, int]:
    """(dense, diag): index bits (bit = n - 1 - wire) a gate changes / only looks at (controls, phases).  Two gates
    commute structurally when they share no bit or only bits on which both are diagonal; a sharded register needs
    the dense bits local and nothing else (csrc/dist.cu: gate_bit_masks)."""
    bits = [n - 1 - int(w) for w in op["wires"]]
    if op["name"] in _DIAGONAL:
        return 0, sum(1 << b for b in bits)
    c = _N_CONTROLS.get(op["name"], 0)
    return sum(1 << b for b in bits[c:]), sum(1 << b for b in bits[:c])


def hardware_efficient_ansatz(n: int, layers: int = 4, seed: int = 11):
    """C3: layers x [RY, RZ on every wire; CNOT(i, i+1) ladder]."""
    rng = np.random.default_rng(seed)
    ops = []
    for _ in range(layers):
        for i in range(n):
            ops.append({"name": "RY", "wires": [i], "params": [float(rng.uniform(-math.pi, math.pi))]})
            ops.append({"name": "RZ", "wires": [i], "params": [float(rng.uniform(-math.pi, math.pi))]})
        for i in range(n - 1):
            ops.append({"name": "CNOT", "wires": [i, i + 1], "params": []})
    return ops, 2 * n * layers


def random_pauli_hamiltonian(n: int, n_terms: int = 100, seed: int = 5):
    """C3: words of weight 1..4 on random distinct wires, letters from {X, Y, Z}, normal coefficients."""
    rng = np.random.default_rng(seed)
    words, wires, coeffs = [], [], []
    for _ in range(n_terms):
        k = int(rng.integers(1, 5))
        ws = sorted(int(x) for x in rng.choice(n, size=min(k, n), replace=False))
        words.append("".join(rng.choice(list("XYZ"), size=len(ws))))
        wires.append(ws)
        coeffs.append(float(rng.normal()))
    return words, wires, coeffs


_PAULI_NAME = {"X": "PauliX", "Y": "PauliY", "Z": "PauliZ", "I": "Identity"}


def hamiltonian_tuple(words, wires, coeffs):
    """Oracle / Observable.from_tuple encoding of a Pauli-word Hamiltonian."""
    terms = []
    for w, ws in zip(words, wires):
        factors = [("Named", _PAULI_NAME[c], [x]) for c, x in zip(w, ws)]
        terms.append(factors[0] if len(factors) == 1 else ("TensorProd", factors))
    return ("Hamiltonian", list(coeffs), terms)


def _pauli_sum_csr(n: int, words, wires, coeffs):
    """CSR matrix of sum_t c_t P_t, one entry per (row, distinct flip mask), columns sorted.  Built with torch on the GPU
    when there is one (a second at 22 qubits / 1.3e8 non-zeros) and with the same code on the CPU otherwise (minutes at
    that size; the tests use <= 14 qubits).  (P psi)_i = i^ny (-1)^{popc((i^x)&z)} psi_{i^x}  ->  H[i, i^x] += c i^ny sign."""
    import scipy.sparse as sp
    import torch

    dev = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")
    dim = 1 << n
    idx = torch.arange(dim, dtype=torch.int64, device=dev)
    by_mask: dict[int, object] = {}
    for w, ws, c in zip(words, wires, coeffs):
        x = z = ny = 0
        for ch, q in zip(w, ws):
            b = 1 << (n - 1 - q)
            if ch in "XY":
                x |= b
            if ch in "ZY":
                z |= b
            if ch == "Y":
                ny += 1
        t = (idx ^ x) & z
        for sh in (32, 16, 8, 4, 2, 1):  # parity by xor-folding
            t = t ^ (t >> sh)
        sign = (1 - 2 * (t & 1)).to(torch.float64)
        ph = c * (1j ** ny)
        v = by_mask.get(x)
        if v is None:
            v = by_mask[x] = torch.zeros(dim, 2, dtype=torch.float64, device=dev)
        v[:, 0] += ph.real * sign
        v[:, 1] += ph.imag * sign
    masks = sorted(by_mask)
    k = len(masks)
    cols = torch.stack([idx ^ x for x in masks], dim=1)                    # [dim, k]
    vals = torch.stack([by_mask[x] for x in masks], dim=1)                 # [dim, k, 2]
    del by_mask
    cols, order = torch.sort(cols, dim=1)
    vals = torch.gather(vals, 1, order.unsqueeze(-1).expand(-1, -1, 2))
    indices = cols.reshape(-1).cpu().numpy()
    data = vals.reshape(-1, 2).cpu().numpy().view(np.complex128).reshape(-1)
    indptr = np.arange(dim + 1, dtype=np.int64) * k
    del cols, vals, order
    return sp.csr_matrix((data, indices, indptr), shape=(dim, dim))


def molecular_style_sparse_hamiltonian(n: int, n_terms: int = 400, n_flip_masks: int = 30, seed: int = 3):
    """C4: 2/3 Z-only words, the rest on a small set of even-weight X/Y flip masks; returns scipy CSR
    plus the Pauli-word form (words, wires, coeffs) for the cross-check."""
    import scipy.sparse as sp

    rng = np.random.default_rng(seed)
    masks = []
    for _ in range(n_flip_masks):
        k = int(rng.choice([2, 4]))
        masks.append(sorted(int(x) for x in rng.choice(n, size=k, replace=False)))
    words, wires, coeffs = [], [], []
    for t in range(n_terms):
        if t % 3 != 2:
            k = int(rng.integers(1, 5))
            ws = sorted(int(x) for x in rng.choice(n, size=k, replace=False))
            words.append("Z" * k)
            wires.append(ws)
        else:
            ws = masks[int(rng.integers(len(masks)))]
            # an even number of Y letters keeps the matrix real-symmetric like molecular Hamiltonians
            letters = ["X"] * len(ws)
            for j in rng.choice(len(ws), size=2 * int(rng.integers(0, len(ws) // 2 + 1)), replace=False):
                letters[int(j)] = "Y"
            words.append("".join(letters))
            wires.append(list(ws))
        coeffs.append(float(rng.normal()))
    m = _pauli_sum_csr(n, words, wires, coeffs)
    return m, (words, wires, coeffs)
