"""``LightningGPU`` device mirror over the pybind11 module ``lightning_gpu_qubit_ops``.

The reference device (pennylane_lightning_gpu/lightning_gpu.py:205-960) subclasses PennyLane's
``QubitDevice``; PennyLane is not installable in this image, so this mirror keeps the reference's
constructor keywords (``wires, sync, c_dtype, shots, batch_obs``), method names and call pattern into the
binary module (``getattr(self._gpu_state, op_name)(wires, inverse, params)``, ``ExpectationValue``
overloads, ``Probability`` + re-ordering, ``GenerateSamples``, ``create_ops_list`` +
``adjoint_jacobian``) on plain operation records, and derives from ``QubitDevice`` automatically when
PennyLane is importable.  There is no CPU fallback class (the reference's lightning_gpu.py:975-998 falls back
to lightning.qubit when the binary is missing): a missing binary raises ImportError.

Operation record: anything with ``.name``, ``.wires`` (list of ints), ``.parameters`` (list of floats) and
optionally ``.inverse`` / ``.matrix`` -- ``Op`` below, or a PennyLane operator.
Observable record: ``Obs`` below (``name``, ``wires``, and for composites ``coeffs`` / ``terms`` / ``matrix`` /
``csr``) -- the structure _serialize.py builds from PennyLane observables.
"""
from __future__ import annotations

import os
import sys
from dataclasses import dataclass, field
from typing import Sequence, Union

import numpy as np

try:  # one module object per process: pybind11 types can only be registered once
    if __package__:
        from . import lightning_gpu_qubit_ops as _ops
    else:  # pragma: no cover - file used outside the package
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        import lightning_gpu_qubit_ops as _ops  # noqa: E402
except ImportError as e:  # pragma: no cover - loud failure instead of a CPU fallback
    raise ImportError(
        "lightning_gpu_qubit_ops is not built: run `python -m pennylane_lightning_gpu_b200._build` "
        "(this device has no CPU fallback)"
    ) from e

try:  # PennyLane is optional: with it the class is a real QubitDevice
    from pennylane import QubitDevice as _Base  # type: ignore
except Exception:  # noqa: BLE001
    _Base = object

PLException = _ops.PLException


@dataclass
class Op:
    name: str
    wires: Sequence[int]
    parameters: Sequence[float] = ()
    inverse: bool = False
    matrix: np.ndarray | None = None


@dataclass
class Obs:
    """name in {PauliX, PauliY, PauliZ, Hadamard, Identity, Hermitian, Tensor, Hamiltonian, SparseHamiltonian}."""
    name: str
    wires: Sequence[int] = ()
    matrix: np.ndarray | None = None          # Hermitian
    terms: Sequence["Obs"] = field(default_factory=list)   # Tensor factors / Hamiltonian terms
    coeffs: Sequence[float] = ()              # Hamiltonian
    csr: tuple | None = None                  # SparseHamiltonian: (indptr, indices, data)


_PAULI_LETTER = {"PauliX": "X", "PauliY": "Y", "PauliZ": "Z", "Identity": "I"}
_STATE_PREPS = ("QubitStateVector", "StatePrep", "BasisState")


class LightningGPU(_Base):
    """B200-native ``lightning.gpu`` device: one GPU, or with ``mpi=True`` one process per GPU (torchrun / mpirun)
    sharing one register through NCCL / NVLink peer access."""

    name = "B200-native Lightning GPU device"
    short_name = "lightning.gpu"
    _CPP_BINARY_AVAILABLE = True

    def __init__(self, wires, *, mpi: bool = False, mpi_buf_size: int = 0, sync: bool = False,
                 c_dtype=np.complex128, shots=None, batch_obs: Union[bool, int] = False, seed: int | None = None,
                 fuse_ops: bool = True):
        if c_dtype is np.complex64 or np.dtype(c_dtype) == np.complex64:
            self.use_csingle, self.R_DTYPE, self.C_DTYPE, bits = True, np.float32, np.complex64, "64"
        elif c_dtype is np.complex128 or np.dtype(c_dtype) == np.complex128:
            self.use_csingle, self.R_DTYPE, self.C_DTYPE, bits = False, np.float64, np.complex128, "128"
        else:
            raise TypeError(f"Unsupported complex Type: {c_dtype}")
        if mpi_buf_size < 0:
            raise TypeError(f"Unsupported mpi_buf_size value: {mpi_buf_size}")
        if mpi_buf_size and mpi_buf_size & (mpi_buf_size - 1):
            raise TypeError(f"Unsupported mpi_buf_size value: {mpi_buf_size}. mpi_buf_size should be power of 2.")
        if _Base is not object:
            super().__init__(wires, shots=shots, r_dtype=self.R_DTYPE, c_dtype=self.C_DTYPE)
        else:
            self.num_wires = wires if isinstance(wires, int) else len(wires)
            self.shots = shots
        self._bits = bits
        self._sync = sync
        self._fuse_ops = bool(fuse_ops)
        self._batch_obs = batch_obs
        self._seed = seed
        self._dp = _ops.DevPool()
        self._mpi = bool(mpi)
        if not mpi:
            self._num_local_wires = self.num_wires
            self._gpu_state = getattr(_ops, "LightningGPU_C" + bits)(self.num_wires)
        else:
            # lightning_gpu.py:298-324: one process per GPU (torchrun / mpirun), top log2(P) wires are global
            self._mpi_manager = _ops.MPIManager()
            if self._dp.getTotalDevices() < self._mpi_manager.getSizeNode():
                raise ValueError("Number of devices should be larger than or equal to the number of processes on each node.")
            if self._mpi_manager.getSize() > (1 << (self.num_wires - 1)):
                raise ValueError("Number of processes should be smaller than the number of statevector elements.")
            self._num_global_wires = self._mpi_manager.getSize().bit_length() - 1
            self._num_local_wires = self.num_wires - self._num_global_wires
            deviceid = self._mpi_manager.getRank() % self._mpi_manager.getSizeNode()
            self._devtag = _ops.DevTag(deviceid)
            self._gpu_state = getattr(_ops, "LightningGPUMPI_C" + bits)(self._mpi_manager, self._devtag, mpi_buf_size,
                                                                        self._num_global_wires, self._num_local_wires)
        self._state_host = None
        self._samples = None

    @classmethod
    def capabilities(cls):
        """lightning_gpu.py:485-497."""
        base = getattr(super(), "capabilities", None)
        caps = dict(base()) if callable(base) else {}
        caps.update(model="qubit", supports_inverse_operations=True, supports_analytic_computation=True,
                    supports_finite_shots=True, returns_state=True)
        caps.pop("passthru_devices", None)
        return caps

    @staticmethod
    def _adjoint_jacobian_processing(jac):
        """Post-processing of the Jacobian for the new return type system (lightning_gpu.py:754-769): a scalar for one
        observable and one parameter, a tuple of arrays for one of the two, a tuple of tuples otherwise."""
        jac = np.squeeze(jac)
        if jac.ndim == 0:
            return np.array(jac)
        if jac.ndim == 1:
            return tuple(np.array(j) for j in jac)
        return tuple(tuple(np.array(j_) for j_ in j) for j in jac)

    # ---- state ------------------------------------------------------------------------------------
    def reset(self):
        if _Base is not object:
            super().reset()
        self._gpu_state.resetGPU(False)
        self._state_host = None

    @property
    def state(self) -> np.ndarray:
        """The state vector; with mpi=True this rank's shard (doc/devices.rst:130-151 of the reference)."""
        out = np.zeros(1 << self._num_local_wires, dtype=self.C_DTYPE)
        self._gpu_state.DeviceToHost(out, False)
        return out

    def syncD2H(self, state_vector: np.ndarray, use_async: bool = False):
        self._gpu_state.DeviceToHost(state_vector.ravel(order="C"), use_async)

    def syncH2D(self, state_vector=None, use_async: bool = False):
        sv = self._state_host if state_vector is None else state_vector
        self._gpu_state.HostToDevice(np.ascontiguousarray(sv, dtype=self.C_DTYPE).ravel(order="C"), use_async)

    def _create_basis_state_GPU(self, index: int, use_async: bool = False):
        self._gpu_state.setBasisState(index, use_async)

    def _apply_state_vector_GPU(self, state, device_wires, use_async: bool = False):
        """StatePrep on a subset of wires: the other wires are |0> (lightning_gpu.py:392-447)."""
        state = np.asarray(state, dtype=self.C_DTYPE).reshape(-1)
        device_wires = list(device_wires)
        if len(device_wires) == self.num_wires and device_wires == sorted(device_wires):
            if self._mpi:
                local = np.zeros(1 << self._num_local_wires, dtype=self.C_DTYPE)
                self._mpi_manager.Scatter(state, local, 0)  # lightning_gpu.py:427-431
                state = local
            self.syncH2D(state, use_async)
            return
        n = self.num_wires
        k = len(device_wires)
        idx = np.zeros(1 << k, dtype=np.int64)
        for j, w in enumerate(device_wires):
            idx |= ((np.arange(1 << k) >> (k - 1 - j)) & 1) << (n - 1 - w)
        idt = np.int32 if self.use_csingle else np.int64
        self._gpu_state.setStateVector(idx.astype(idt), state, use_async)

    def _apply_basis_state_GPU(self, bits, wires):
        n = self.num_wires
        index = 0
        for b, w in zip(bits, wires):
            index |= int(b) << (n - 1 - w)
        self._create_basis_state_GPU(index)

    # ---- gates ------------------------------------------------------------------------------------
    def apply_cq(self, operations):
        """The reference makes one call into the binary per operation, dispatched by name (lightning_gpu.py:519-555).  Here
        the whole list goes down in ONE call -- the vector form of `apply` (the reference's own batched overload,
        StateVectorCudaManaged.hpp:279-315, extended by per-operation matrices) -- so that consecutive gates are fused
        into shared HBM sweeps.  `fuse_ops=False` on the device restores the per-operation calls.  The `inverse` flag is
        taken per operation (reference defect Q1 -- a sticky flag -- is not reproduced)."""
        names, wires, invs, params, mats = [], [], [], [], []
        for o in operations:
            name = o.name
            if name in _STATE_PREPS or name == "Identity":
                continue
            inv = bool(getattr(o, "inverse", False))
            w = list(o.wires)
            method = getattr(self._gpu_state, name, None)
            named = method is not None and getattr(o, "matrix", None) is None
            if not self._fuse_ops:
                if named:
                    method(w, inv, [float(p) for p in o.parameters])
                else:
                    self._gpu_state.apply(name, w, inv, [], np.asarray(o.matrix, dtype=self.C_DTYPE).reshape(-1))
                continue
            names.append(name)
            wires.append(w)
            invs.append(inv)
            params.append([float(p) for p in o.parameters] if named else [])
            mats.append([] if named else np.asarray(o.matrix, dtype=self.C_DTYPE).reshape(-1))
        if names:
            self._gpu_state.apply(names, wires, invs, params, mats)

    def apply(self, operations, **kwargs):
        ops = list(operations)
        if ops and ops[0].name in ("QubitStateVector", "StatePrep"):
            self._apply_state_vector_GPU(ops[0].parameters[0], ops[0].wires)
            ops = ops[1:]
        elif ops and ops[0].name == "BasisState":
            self._apply_basis_state_GPU(ops[0].parameters[0], ops[0].wires)
            ops = ops[1:]
        for o in ops:
            if o.name in _STATE_PREPS:
                raise ValueError(f"Operation {o.name} cannot be used after other Operations have already been applied")
        self.apply_cq(ops)
        if self._sync:
            self._state_host = self.state

    # ---- measurements ---------------------------------------------------------------------------
    def _serialize_obs(self, o: Obs):
        b = ("MPI_C" if self._mpi else "_C") + self._bits
        return self._serialize_obs_as(o, b)

    def _serialize_obs_as(self, o: Obs, b: str):
        if o.name in ("PauliX", "PauliY", "PauliZ", "Hadamard", "Identity"):
            return getattr(_ops, "NamedObsGPU" + b)(o.name, list(o.wires))
        if o.name == "Hermitian":
            return getattr(_ops, "HermitianObsGPU" + b)(np.asarray(o.matrix, dtype=self.C_DTYPE).reshape(-1), list(o.wires))
        if o.name == "Tensor":
            return getattr(_ops, "TensorProdObsGPU" + b)([self._serialize_obs_as(t, b) for t in o.terms])
        if o.name == "Hamiltonian":
            return getattr(_ops, "HamiltonianGPU" + b)(np.asarray(o.coeffs, dtype=self.R_DTYPE),
                                                        [self._serialize_obs_as(t, b) for t in o.terms])
        if o.name == "SparseHamiltonian":
            indptr, indices, data = o.csr
            idt = np.int32 if self.use_csingle else np.int64
            return getattr(_ops, "SparseHamiltonianGPU" + b)(np.asarray(data, dtype=self.C_DTYPE),
                                                               np.asarray(indices, dtype=idt),
                                                               np.asarray(indptr, dtype=idt), list(range(self.num_wires)))
        raise ValueError(f"unsupported observable {o.name}")

    @staticmethod
    def _pauli_word(o: Obs):
        if o.name in _PAULI_LETTER:
            return _PAULI_LETTER[o.name], list(o.wires)
        if o.name == "Tensor" and all(t.name in _PAULI_LETTER for t in o.terms):
            return "".join(_PAULI_LETTER[t.name] for t in o.terms), [t.wires[0] for t in o.terms]
        return None

    _NAMED_MATRIX = {
        "PauliX": np.array([[0, 1], [1, 0]], dtype=np.complex128),
        "PauliY": np.array([[0, -1j], [1j, 0]], dtype=np.complex128),
        "PauliZ": np.array([[1, 0], [0, -1]], dtype=np.complex128),
        "Hadamard": np.array([[1, 1], [1, -1]], dtype=np.complex128) / np.sqrt(2.0),
        "Identity": np.eye(2, dtype=np.complex128),
    }
    _MAX_DENSE_WIRES = 10  # 2^10 x 2^10 complex128 = 16 MiB on the host; the reference builds qml.matrix up to 13 wires

    @classmethod
    def _matrix_of(cls, o: Obs):
        """(wires, dense matrix on those wires in PennyLane order) of a named / Hermitian / Tensor / Hamiltonian observable:
        what ``qml.matrix(observable)`` gives the reference (lightning_gpu.py:884-897, 936-960)."""
        if o.name in cls._NAMED_MATRIX:
            return list(o.wires), cls._NAMED_MATRIX[o.name]
        if o.name == "Hermitian":
            k = len(o.wires)
            return list(o.wires), np.asarray(o.matrix, dtype=np.complex128).reshape(1 << k, 1 << k)
        if o.name not in ("Tensor", "Hamiltonian"):
            raise NotImplementedError(f"no dense matrix for observable {o.name}")
        parts = [cls._matrix_of(t) for t in o.terms]
        wires = []
        for w, _ in parts:
            wires += [x for x in w if x not in wires]
        if len(wires) > cls._MAX_DENSE_WIRES:
            raise NotImplementedError(f"dense matrix of an observable on {len(wires)} wires")
        k = len(wires)

        def embed(w, m):  # m on wires w -> the same operator on `wires`
            if o.name == "Tensor" and any(x in seen for x in w):
                raise ValueError("the factors of a tensor product must act on distinct wires")
            rest = [x for x in wires if x not in w]
            full = np.kron(m, np.eye(1 << len(rest), dtype=np.complex128)).reshape([2] * (2 * k))
            order = list(w) + rest                          # current axis order of the row (and column) indices
            perm = [order.index(x) for x in wires]
            return full.transpose(perm + [k + p for p in perm]).reshape(1 << k, 1 << k)

        seen = []
        if o.name == "Tensor":
            out = np.eye(1 << k, dtype=np.complex128)
            for w, m in parts:
                out = out @ embed(w, m)
                seen += list(w)
            return wires, out
        out = np.zeros((1 << k, 1 << k), dtype=np.complex128)
        for c, (w, m) in zip(o.coeffs, parts):
            out = out + complex(c) * embed(w, m)
        return wires, out

    def expval(self, observable: Obs, shot_range=None, bin_size=None) -> float:
        """Routing of lightning_gpu.py:820-897, except that a Hamiltonian of Pauli words always takes the
        fused Pauli-word kernels (the reference builds a dense 2^k x 2^k host matrix below 14 wires)."""
        if self.shots is not None:
            return float(np.squeeze(np.mean(self._sample_observable(observable))))
        if observable.name == "SparseHamiltonian":
            indptr, indices, data = observable.csr
            idt = np.int32 if self.use_csingle else np.int64
            return self._gpu_state.ExpectationValue(np.asarray(indptr, dtype=idt), np.asarray(indices, dtype=idt),
                                                    np.asarray(data, dtype=self.C_DTYPE))
        if observable.name == "Hamiltonian":
            words = [self._pauli_word(t) for t in observable.terms]
            if all(w is not None for w in words):
                return self._gpu_state.ExpectationValue([w[0] for w in words], [w[1] for w in words],
                                                        np.asarray(observable.coeffs, dtype=self.C_DTYPE))
            return float(sum(c * self.expval(t) for c, t in zip(observable.coeffs, observable.terms)))
        if observable.name == "Hermitian":
            return self._gpu_state.ExpectationValue(list(observable.wires),
                                                    np.asarray(observable.matrix, dtype=self.C_DTYPE).reshape(-1))
        word = self._pauli_word(observable)
        if word is not None and observable.name == "Tensor":
            return self._gpu_state.ExpectationValue([word[0]], [word[1]], np.ones(1, dtype=self.C_DTYPE))
        if observable.name == "Tensor":
            # factors other than Pauli letters (Hadamard, Hermitian): the dense matrix of the product, as the reference does
            wires, m = self._matrix_of(observable)
            return self._gpu_state.ExpectationValue(wires, m.astype(self.C_DTYPE).reshape(-1))
        return self._gpu_state.ExpectationValue(observable.name, list(observable.wires), [],
                                                np.zeros(0, dtype=self.C_DTYPE))

    def var(self, observable: Obs, shot_range=None, bin_size=None) -> float:
        """<O^2> - <O>^2 (lightning_gpu.py:936-960); for Pauli words and Hadamard O^2 = 1."""
        if self.shots is not None:
            return float(np.squeeze(np.var(self._sample_observable(observable))))
        mean = self.expval(observable)
        if observable.name in ("PauliX", "PauliY", "PauliZ", "Hadamard", "Identity") or self._pauli_word(observable):
            return 1.0 - mean**2
        # <O^dagger O> - <O>^2 from the dense matrix of the observable (lightning_gpu.py:944-960)
        wires, m = self._matrix_of(observable)
        sq = self._gpu_state.ExpectationValue(wires, (m.conj().T @ m).astype(self.C_DTYPE).reshape(-1))
        return sq - mean**2

    def probability(self, wires=None, shot_range=None, bin_size=None) -> np.ndarray:
        """Marginal probabilities in PennyLane order; the binary returns cuStateVec bit order and is
        re-transposed here exactly as lightning_gpu.py:899-926 does (sorted wires only, like the reference)."""
        wires = list(range(self.num_wires)) if wires is None else list(wires)
        if self.shots is not None:
            s = self.generate_samples()[:, wires]
            idx = s.astype(np.int64) @ (1 << np.arange(len(wires) - 1, -1, -1))
            return np.bincount(idx, minlength=1 << len(wires)) / len(idx)
        if wires != sorted(wires):
            raise RuntimeError("Lightning-GPU does not currently support out-of-order indices for probabilities")
        p = self._gpu_state.Probability(wires)
        k = len(wires)
        return p.reshape([2] * k).transpose().reshape(-1)

    def generate_samples(self) -> np.ndarray:
        if self._seed is None:
            return self._gpu_state.GenerateSamples(self.num_wires, int(self.shots)).astype(int)
        # one stream of sub-seeds per device, seeded once: every call (every observable, every execution) draws fresh
        # uniforms, and the whole sequence is reproducible for a given `seed`
        if getattr(self, "_seed_stream", None) is None:
            self._seed_stream = np.random.default_rng(int(self._seed))
        sub_seed = int(self._seed_stream.integers(0, 2**63 - 1))
        return self._gpu_state.GenerateSamples(self.num_wires, int(self.shots), sub_seed).astype(int)

    def sample(self, observable: Obs, shot_range=None, bin_size=None, counts: bool = False):
        """Eigenvalue samples of an observable (lightning_gpu.py:812-818): the state is rotated into the observable's
        eigenbasis, sampled in the computational basis and restored; counts=True returns {eigenvalue: occurrences}."""
        s = self._sample_observable(observable)
        if shot_range is not None:
            s = s[slice(*shot_range)]
        if counts:
            vals, n = np.unique(s, return_counts=True)
            return {float(v): int(c) for v, c in zip(vals, n)}
        return s if bin_size is None else s.reshape(-1, bin_size)

    def _sample_observable(self, observable: Obs) -> np.ndarray:
        """Eigenvalue samples of a Pauli-word observable measured in the computational basis after the
        diagonalising rotations (QubitDevice.sample)."""
        word = self._pauli_word(observable)
        if word is None:
            raise NotImplementedError("sampling is implemented for Pauli words")
        saved = type(self._gpu_state)(self._gpu_state)
        for letter, w in zip(*word):
            if letter == "X":
                self._gpu_state.Hadamard([w], False, [])
            elif letter == "Y":
                self._gpu_state.S([w], True, [])
                self._gpu_state.Hadamard([w], False, [])
        s = self.generate_samples()
        self._gpu_state.DeviceToDevice(saved, False)
        cols = [w for letter, w in zip(*word) if letter != "I"]
        return 1.0 - 2.0 * (s[:, cols].sum(axis=1) % 2)

    # ---- adjoint differentiation ---------------------------------------------------------------------
    def _serialize_ops(self, operations):
        """(names, params, wires, inverses, matrices) with Rot expanded into RZ RY RZ
        (pennylane_lightning_gpu/_serialize.py:279-336)."""
        names, params, wires, invs, mats = [], [], [], [], []
        for o in operations:
            if o.name in _STATE_PREPS:
                continue
            inv = bool(getattr(o, "inverse", False))
            if o.name == "Rot" and not inv:
                phi, theta, omega = (float(p) for p in o.parameters)
                for n_, p_ in (("RZ", phi), ("RY", theta), ("RZ", omega)):
                    names.append(n_); params.append(np.array([p_], dtype=self.R_DTYPE)); wires.append(list(o.wires))
                    invs.append(False); mats.append(np.zeros(0, dtype=self.C_DTYPE))
                continue
            names.append(o.name)
            params.append(np.asarray([float(p) for p in o.parameters], dtype=self.R_DTYPE))
            wires.append(list(o.wires))
            invs.append(inv)
            m = getattr(o, "matrix", None)
            mats.append(np.zeros(0, dtype=self.C_DTYPE) if m is None else np.asarray(m, dtype=self.C_DTYPE).reshape(-1))
        return names, params, wires, invs, mats

    def adjoint_jacobian(self, operations, observables: Sequence[Obs], trainable_params=None, starting_state=None,
                         use_device_state: bool = False) -> np.ndarray:
        """Jacobian d<O_i>/d theta_j by the adjoint method (lightning_gpu.py:638-752).  Unlike the reference,
        a Hamiltonian observable is ONE row computed from one bra (the reference splits it per term in
        _serialize.py:158-167 and sums the rows afterwards, lightning_gpu.py:743-748)."""
        if self.shots is not None:
            import warnings

            warnings.warn("Requested adjoint differentiation to be computed with finite shots. The derivative is "
                          "always exact when using the adjoint differentiation method.", UserWarning)
        operations = list(operations)
        self._check_adjdiff_supported_operations(operations)
        if not observables:
            return np.array([], dtype=self.R_DTYPE)
        if starting_state is not None:
            self.syncH2D(np.asarray(starting_state, dtype=self.C_DTYPE))
            self.apply_cq(operations)
        elif not use_device_state:
            self.reset()
            self.apply(operations)
        names, params, wires, invs, mats = self._serialize_ops(operations)
        n_par = sum(1 for p in params if len(p))
        tp = list(range(n_par)) if trainable_params is None else sorted(trainable_params)
        # the reverse sweep walks the parameters in ascending order; columns are handed back in the caller's order
        order = None
        if trainable_params is not None and list(trainable_params) != tp:
            if len(set(tp)) != len(tp):
                raise ValueError("trainable_params must not contain duplicates")
            order = [tp.index(int(t)) for t in trainable_params]
        if not tp:
            return np.zeros((len(observables), 0), dtype=self.R_DTYPE)
        adj = getattr(_ops, ("AdjointJacobianGPUMPI_C" if self._mpi else "AdjointJacobianGPU_C") + self._bits)()
        rec = adj.create_ops_list(names, params, wires, invs, mats)
        obs = [self._serialize_obs(o) for o in observables]
        if self._mpi:  # batch_obs = memory-saving one-observable-at-a-time sweep (lightning_gpu.py:704-737)
            fn = adj.adjoint_jacobian_serial if self._batch_obs else adj.adjoint_jacobian
        else:
            fn = adj.adjoint_jacobian_batched if self._batch_obs else adj.adjoint_jacobian
        jac = np.asarray(fn(self._gpu_state, obs, rec, tp))
        return jac if order is None else jac[:, order]

    @staticmethod
    def _check_adjdiff_supported_operations(operations):
        """lightning_gpu.py:622-636: operations with more than one parameter other than Rot cannot be differentiated;
        an inverse Rot is not expanded by the serializer, so it is rejected here instead of in the C++ layer."""
        for op in operations:
            if len(getattr(op, "parameters", ())) > 1 and op.name != "Rot" and op.name not in _STATE_PREPS:
                raise ValueError(f'The {op.name} operation is not supported using the "adjoint" differentiation method')
            if op.name == "Rot" and getattr(op, "inverse", False):
                raise ValueError('The inverse of Rot is not supported using the "adjoint" differentiation method; '
                                 "write it as RZ(-omega) RY(-theta) RZ(-phi)")

    def vjp(self, operations, observables, dy, trainable_params=None, **kw) -> np.ndarray:
        """Vector-Jacobian product dy^T J (lightning_gpu.py:771-810).  As in the reference the observables are combined
        into one Hamiltonian sum_i dy_i O_i first, so the reverse sweep carries one bra instead of one per observable."""
        if np.iscomplexobj(dy):
            raise ValueError("The vjp method only works with a real-valued dy when the tape is returning an expectation value")
        dy = np.asarray(dy, dtype=self.R_DTYPE).reshape(-1)
        if len(dy) != len(observables):
            raise ValueError("Number of observables in the tape must be the same as the length of dy in the vjp method")
        if np.allclose(dy, 0):
            if trainable_params is not None:
                n = len(trainable_params)
            else:  # all parameters of the tape, counted the way the serializer does (Rot = 3)
                n = sum(1 for p in self._serialize_ops(list(operations))[1] if len(p))
            return np.zeros(n, dtype=self.R_DTYPE)
        coeffs, terms = [], []
        for w, o in zip(dy, observables):
            if o.name == "Hamiltonian":
                coeffs += [float(w) * float(c) for c in o.coeffs]
                terms += list(o.terms)
            else:
                coeffs.append(float(w))
                terms.append(o)
        if all(t.name in ("PauliX", "PauliY", "PauliZ", "Hadamard", "Identity", "Hermitian", "Tensor") for t in terms):
            ham = Obs("Hamiltonian", coeffs=coeffs, terms=terms)
            return np.asarray(self.adjoint_jacobian(operations, [ham], trainable_params, **kw)).reshape(-1)
        jac = self.adjoint_jacobian(operations, observables, trainable_params, **kw)  # sparse / nested observables
        return dy @ jac.reshape(len(dy), -1)
