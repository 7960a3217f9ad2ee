"""Sharded state vectors over NCCL: one process per GPU, launched by torchrun.

StateVectorCudaMPI analogue (simulator/StateVectorCudaMPI.hpp of the reference).  torch.distributed is
only the rendezvous that carries the 128-byte NCCL unique id from rank 0 to the other ranks (the role
MPI_Bcast / MPI_Allgather of IPC handles plays at MPIWorker.hpp:306-320); all data-path communication
is NCCL send/recv + all-reduce inside libqsv_b200.so (csrc/dist.cu).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _cabi
from ._cabi import Ops, StateVector, _check, lib


def plan(ops: Ops, n_total: int, n_local: int, initial_map=None):
    """Host-only view of the exchange schedule: -> (steps, final map).  steps = list of
    ("swap", global_phys_bit, local_phys_bit) | ("gate", op_index).  initial_map = physical bit of every logical bit
    before the circuit (None = identity); the map persists between calls of apply_ops, so feeding the final map back
    in gives the schedule of a repeated application."""
    cap = 4 * len(ops) + 16
    steps = (C.c_int * (3 * cap))()
    n_steps = C.c_int(0)
    final = (C.c_int * n_total)()
    first = (C.c_int * n_total)(*[int(x) for x in initial_map]) if initial_map is not None else None
    _check(lib().qsv_dist_plan_from(ops._h, n_total, n_local, first, steps, cap, C.byref(n_steps), final))
    out = []
    for i in range(n_steps.value):
        kind, a, b = steps[3 * i], steps[3 * i + 1], steps[3 * i + 2]
        out.append(("swap", a, b) if kind == 0 else ("gate", a))
    return out, list(final)


class DistributedStateVector:
    """n_total-qubit register sharded over WORLD_SIZE GPUs on the top log2(WORLD_SIZE) index bits
    (PennyLane wires 0..g-1 are global, as in lightning_gpu.py:317-319 / MPI.hpp:240-246)."""

    def __init__(self, n_total: int, dtype=np.complex128, device: int = 0, *, external_ptr: int | None = None,
                 chunk_bytes: int = 0, lazy_map: bool = True):
        """lazy_map=True (this class's default, the throughput setting): the logical->physical qubit map persists
        between calls, and `d2h` / `canonicalize` are COLLECTIVE (every rank must call them).  lazy_map=False is the
        C ABI's own default and what the pybind classes (`LightningGPUMPI_C*`) use: every call returns with the canonical
        layout and `d2h` is a purely local copy, as in the reference."""
        import torch
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed must be initialised (torchrun) before creating a sharded register")
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        g = int(math.log2(self.world))
        if 1 << g != self.world:
            raise ValueError("number of ranks must be a power of two")
        self.n_total, self.n_global, self.n_local = n_total, g, n_total - g
        self.chunk_bytes = chunk_bytes
        self.local = StateVector(self.n_local, dtype, device, external_ptr=external_ptr)
        if external_ptr is None:
            # |0...0>: only rank 0 holds the amplitude 1
            if self.rank != 0:
                self.local.set_state_vector([], np.zeros(0, dtype=dtype))
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            _check(lib().qsv_dist_unique_id(buf))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(torch.device("cuda", device)) if dist.get_backend() == "nccl" else ident
        dist.broadcast(ident, src=0)
        raw = bytes(ident.cpu().tolist())
        self._id = (C.c_ubyte * 128).from_buffer_copy(raw)
        _check(lib().qsv_dist_init(self.local._h, self._id, self.rank, self.world))
        self.lazy_map = bool(lazy_map)
        if self.lazy_map:
            _check(lib().qsv_dist_set_lazy_map(self.local._h, 1))

    def set_lazy_map(self, lazy: bool):
        """Collective: switches the qubit-map policy (see __init__); back to eager restores the canonical layout."""
        _check(lib().qsv_dist_set_lazy_map(self.local._h, int(bool(lazy))))
        self.lazy_map = bool(lazy)

    # -- gates ----------------------------------------------------------------------------------
    def apply_ops(self, ops: Ops, fuse: bool = True):
        _check(lib().qsv_dist_apply_ops(self.local._h, ops._h, int(bool(fuse)), self.chunk_bytes))

    def last_apply_stats(self):
        return self.local.last_apply_stats()

    def swap_stats(self, reset: bool = False):
        n, b, ms = C.c_int(0), C.c_uint64(0), C.c_float(0)
        _check(lib().qsv_dist_total_swap_stats(self.local._h, C.byref(n), C.byref(b), C.byref(ms), int(reset)))
        return n.value, b.value, ms.value

    def fused_exchange_stats(self):
        """(exchanges done through the second buffer, how many of them a gate sweep carried); on by default, QSV_DIST_FUSED_SWAP=0 switches it off."""
        a, b = C.c_int(0), C.c_int(0)
        _check(lib().qsv_dist_fused_exchange_stats(self.local._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def split_exchange_stats(self):
        """(second halves of split exchanges performed, how many of them a gate sweep carried) -- QSV_DIST_SPLIT_XCHG."""
        a, b = C.c_int(0), C.c_int(0)
        _check(lib().qsv_dist_split_exchange_stats(self.local._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    @property
    def uses_peer_access(self) -> bool:
        return bool(lib().qsv_dist_uses_peer_access(self.local._h))

    def qubit_map(self):
        m = (C.c_int * self.n_total)()
        _check(lib().qsv_dist_qubit_map(self.local._h, m, self.n_total))
        return list(m)

    def canonicalize(self):
        _check(lib().qsv_dist_canonicalize(self.local._h, self.chunk_bytes))

    # -- measurements ---------------------------------------------------------------------------
    def expval_pauli_words(self, words, wires, coeffs, return_terms: bool = False):
        letters = "".join(words).encode()
        offs = np.zeros(len(words) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(w) for w in words])
        wa, wp = _cabi._ints([int(x) for ws in wires for x in ws])
        ca, cp = _cabi._cmat(coeffs)
        terms = np.zeros(max(len(words), 1), dtype=np.float64)
        out = C.c_double(0)
        _check(lib().qsv_dist_expval_pauli_words(self.local._h, len(words), letters, wp,
                                                 offs.ctypes.data_as(_cabi._IP), cp,
                                                 terms.ctypes.data_as(_cabi._DP), C.byref(out)))
        return (out.value, terms[: len(words)]) if return_terms else out.value

    def set_basis_state(self, index: int = 0):
        """setBasisState on the whole register (MPI.hpp:332-345); resets the qubit map."""
        _check(lib().qsv_dist_set_basis_state(self.local._h, index))

    def set_state_vector(self, indices, values):
        """setStateVector (MPI.hpp:357-381): global indices, every rank keeps those of its shard."""
        ia = np.ascontiguousarray(indices, dtype=np.int64)
        va = np.ascontiguousarray(values, dtype=self.local.np_dtype)
        _check(lib().qsv_dist_set_state_vector(self.local._h, ia.ctypes.data_as(_cabi._I64P), va.ctypes.data_as(_cabi._P),
                                               len(ia)))

    def h2d(self, shard):
        """This rank's shard from host memory, canonical layout (CopyHostDataToGpu of StateVectorCudaMPI); resets the
        qubit map.  Collective in the sense that every rank has to call it."""
        a = np.ascontiguousarray(shard, dtype=self.local.np_dtype).reshape(-1)
        _check(lib().qsv_dist_h2d(self.local._h, a.ctypes.data_as(_cabi._P), a.size))

    def d2h(self, out: np.ndarray | None = None) -> np.ndarray:
        """This rank's shard in the canonical layout (CopyGpuDataToHost of StateVectorCudaMPI)."""
        if out is None:
            out = np.empty(1 << self.n_local, dtype=self.local.np_dtype)
        _check(lib().qsv_dist_d2h(self.local._h, out.ctypes.data_as(_cabi._P), out.size))
        return out

    def apply(self, name: str, wires, params=(), adjoint=False, matrix=None):
        """One gate on the whole register (applyOperation, MPI.hpp:392-470)."""
        self.apply_ops(Ops([{"name": name, "wires": list(wires), "params": list(params), "adjoint": adjoint,
                             "matrix": matrix}]), fuse=False)

    def expval_named(self, name, wires, params=()) -> complex:
        wa, wp = _cabi._ints(wires)
        pa, pp = _cabi._dbls(params)
        out = (C.c_double * 2)()
        _check(lib().qsv_dist_expval_named(self.local._h, name.encode(), wp, len(wa), pp, len(pa), out))
        return complex(out[0], out[1])

    def expval_matrix(self, matrix, wires) -> complex:
        wa, wp = _cabi._ints(wires)
        ma, mp = _cabi._cmat(matrix)
        out = (C.c_double * 2)()
        _check(lib().qsv_dist_expval_matrix(self.local._h, mp, wp, len(wa), out))
        return complex(out[0], out[1])

    def expval_csr(self, indptr, indices, data) -> float:
        ip = np.ascontiguousarray(indptr, dtype=np.int64)
        ix = np.ascontiguousarray(indices, dtype=np.int64)
        va, vp = _cabi._cmat(data)
        out = C.c_double(0)
        _check(lib().qsv_dist_expval_csr(self.local._h, ip.ctypes.data_as(_cabi._I64P), ix.ctypes.data_as(_cabi._I64P),
                                         vp, len(ix), C.byref(out)))
        return out.value

    def expval(self, obs) -> float:
        out = C.c_double(0)
        _check(lib().qsv_dist_obs_expval(obs._h, self.local._h, C.byref(out)))
        return out.value

    def apply_observable(self, obs):
        _check(lib().qsv_dist_obs_apply(obs._h, self.local._h))

    def probs(self, wires) -> np.ndarray:
        """Marginal probabilities of the whole register; first listed wire = LSB (as StateVector.probs)."""
        wa, wp = _cabi._ints(wires)
        out = np.zeros(1 << len(wa), dtype=np.float64)
        _check(lib().qsv_dist_probs(self.local._h, wp, len(wa), out.ctypes.data_as(_cabi._DP)))
        return out

    def sample(self, uniforms) -> np.ndarray:
        ua, up = _cabi._dbls(uniforms)
        out = np.zeros((len(ua), self.n_total), dtype=np.uint64)
        _check(lib().qsv_dist_sample(self.local._h, up, len(ua), out.ctypes.data_as(_cabi._U64P)))
        return out

    def adjoint_jacobian(self, ops: Ops, observables, trainable, apply_operations: bool = False) -> np.ndarray:
        tp = np.ascontiguousarray(list(trainable), dtype=np.int64)
        jac = np.zeros((len(observables), len(tp)), dtype=np.float64)
        arr = (_cabi._P * max(len(observables), 1))(*[o._h for o in observables])
        _check(lib().qsv_dist_adjoint_jacobian(self.local._h, ops._h, arr, len(observables),
                                               tp.ctypes.data_as(_cabi._I64P), len(tp), int(bool(apply_operations)),
                                               jac.ctypes.data_as(_cabi._DP)))
        return jac

    def norm2(self) -> float:
        v = np.array([self.local.inner_product(self.local).real], dtype=np.float64)
        _check(lib().qsv_dist_allreduce_f64(self.local._h, v.ctypes.data_as(_cabi._DP), 1))
        return float(v[0])

    def local_state(self) -> np.ndarray:
        """This rank's shard in the CURRENT qubit map (call canonicalize() first for the standard layout)."""
        return self.local.d2h()

    def synchronize(self):
        self.local.synchronize()

    def close(self):
        _check(lib().qsv_dist_finalize(self.local._h))
