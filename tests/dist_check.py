"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Same pattern as the reference's src/tests/mpi/Test_StateVectorCudaMPI_Param.cpp:59-125 (apply on the
sharded register with wires chosen to hit global and local bits, compare shard-wise with the
single-device result), with the oracle as the single-device result."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import np_oracle as orc  # noqa: E402
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402
from pennylane_lightning_gpu_b200.distributed import DistributedStateVector  # noqa: E402


def circuit(n, seed, n_gates=60):
    ops = workloads.random_gate_circuit(n, n_gates, seed)
    extra = [{"name": "CZ", "wires": [0, n - 1], "params": []}, {"name": "RZ", "wires": [0], "params": [0.3]},
             {"name": "CNOT", "wires": [0, 3], "params": []}, {"name": "Toffoli", "wires": [1, 0, 2], "params": []},
             {"name": "IsingXX", "wires": [0, 1], "params": [0.7]}, {"name": "MultiRZ", "wires": [0, 2, 5], "params": [0.9]},
             {"name": "CRY", "wires": [4, 0], "params": [1.1]}, {"name": "SWAP", "wires": [0, n - 2], "params": []},
             {"name": "PhaseShift", "wires": [1], "params": [0.4]}, {"name": "DoubleExcitation", "wires": [0, 1, 2, 3], "params": [0.5]},
             {"name": "ControlledPhaseShift", "wires": [0, 1], "params": [0.2]}, {"name": "Hadamard", "wires": [0], "params": []}]
    for i, e in enumerate(extra):
        ops.insert(3 + 4 * i, e)
    return ops


def _obs_zoo(rng, n):
    h2 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
    h2 = h2 + h2.conj().T
    h1 = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
    h1 = h1 + h1.conj().T
    return [
        ("Named", "PauliZ", [0]),
        ("Named", "PauliX", [0]),
        ("Named", "PauliY", [n - 1]),
        ("Named", "Hadamard", [0]),
        ("Hermitian", h1, [1]),
        ("Hermitian", h2, [n - 1, 0]),
        ("TensorProd", [("Named", "PauliX", [0]), ("Named", "PauliY", [2]), ("Named", "PauliZ", [n - 1])]),
        ("TensorProd", [("Named", "PauliZ", [1]), ("Hermitian", h1, [0])]),
        ("Hamiltonian", [0.3, -1.1, 0.7],
         [("Named", "PauliZ", [0]), ("TensorProd", [("Named", "PauliX", [1]), ("Named", "PauliX", [0])]),
          ("TensorProd", [("Named", "PauliY", [0]), ("Named", "PauliZ", [n - 1])])]),
        ("Hamiltonian", [0.9, 0.4], [("Hermitian", h2, [1, 0]), ("Named", "Hadamard", [0])]),
    ]


def measurements_and_adjoint(rank, local_rank, world, g):
    """Sharded measurements, observables and adjoint Jacobian against the oracle (the reference's own pattern:
    src/tests/mpi/Test_StateVectorCudaMPI_NonParam.cpp:836-927, Test_AdjointDiffGPUMPI.cpp)."""
    failures = []
    n = g + 8
    rng = np.random.default_rng(77)
    for dtype, tol in ((np.complex128, 1e-10), (np.complex64, 1e-5)):
        ops = workloads.random_gate_circuit(n, 40, seed=5 + n)
        ops += [{"name": "Hadamard", "wires": [0], "params": []}, {"name": "CNOT", "wires": [0, n - 1], "params": []},
                {"name": "RY", "wires": [1], "params": [0.37]}]
        psi0 = orc.apply_ops(orc.basis_state(n), [{"name": "Hadamard", "wires": [w], "params": []} for w in range(n)])
        want = orc.apply_ops(psi0, ops)
        sv = DistributedStateVector(n, dtype, device=local_rank)
        # state preparation on the whole register, then the circuit
        sv.set_basis_state(5)
        probe = sv.probs([n - 1, n - 3, 0])  # index 5 = wires n-1 and n-3 set; first listed wire = LSB
        if abs(probe[0b011] - 1.0) > 1e-12:
            failures.append(f"set_basis_state/probs: {probe}")
        idx = np.arange(1 << n, dtype=np.int64)
        sv.set_state_vector(idx, psi0.astype(dtype))
        sv.apply_ops(q.Ops(ops), fuse=True)

        def check(what, got, ref):
            # north_star's tolerance, relative to the size of the quantity when that exceeds 1 (Hamiltonian sums)
            err = float(np.max(np.abs(np.asarray(got) - np.asarray(ref))))
            if not err <= tol * max(1.0, float(np.max(np.abs(np.asarray(ref))))):
                failures.append(f"{np.dtype(dtype).name} {what}: err {err:.2e}")
            if rank == 0:
                print(f"[dist_check] {np.dtype(dtype).name} {what}: err={err:.2e}", flush=True)

        for wires in ([0], [n - 1], [0, n - 1], [2, 0, 1], [n - 2, 1, 0, 3], list(range(g))):
            check(f"probs{wires}", sv.probs(wires), orc.probs_custatevec_order(want, wires))
        h1 = rng.normal(size=(2, 2)) + 1j * rng.normal(size=(2, 2))
        h2 = rng.normal(size=(4, 4)) + 1j * rng.normal(size=(4, 4))
        check("expval_named RX(0)", sv.expval_named("RX", [0], [0.4]), orc.expval_matrix(want, orc.gate_matrix("RX", [0.4]), [0]))
        check("expval_matrix 1q", sv.expval_matrix(h1, [0]), orc.expval_matrix(want, h1, [0]))
        check("expval_matrix 2q", sv.expval_matrix(h2, [n - 1, 0]), orc.expval_matrix(want, h2, [n - 1, 0]))
        zoo = _obs_zoo(np.random.default_rng(3), n)
        for i, t in enumerate(zoo):
            check(f"obs[{i}] expval", sv.expval(q.Observable.from_tuple(t)), orc.expval_obs(want, t))
        # sparse Hamiltonian: row blocks + gathers from the other shards
        m, (w2, ws2, c2) = workloads.molecular_style_sparse_hamiltonian(n, n_terms=40, n_flip_masks=6, seed=3)
        ref_sparse = orc.expval_csr(want, m.indptr, m.indices, m.data)
        check("expval_csr", sv.expval_csr(m.indptr, m.indices, m.data), ref_sparse)
        check("expval sparse obs", sv.expval(q.Observable.sparse(m.indptr, m.indices, m.data)), ref_sparse)
        check("pauli words vs csr", sv.expval_pauli_words(w2, ws2, c2), ref_sparse)
        # sampling: same definition as the single-GPU sampler over the whole register
        shots = 2000
        u = np.random.default_rng(1234).random(shots)
        got = sv.sample(u)
        ref = orc.sample(want.astype(dtype), shots, seed=1234)
        cdf = np.cumsum(np.abs(want) ** 2)
        near = np.min(np.abs(cdf[None, :] - (u * cdf[-1])[:, None]), axis=1) < (1e-12 if dtype == np.complex128 else 1e-6)
        same = np.all(got == ref, axis=1)
        if not (np.all(same | near) and same.mean() > 0.99):
            failures.append(f"{np.dtype(dtype).name} sampling: {same.mean():.4f} identical")
        # observable application
        for i in (5, 8, 9):
            a = DistributedStateVector(n, dtype, device=local_rank)
            a.set_state_vector(idx, want.astype(dtype))
            a.apply_observable(q.Observable.from_tuple(zoo[i]))
            a.canonicalize()
            shard = a.local_state().astype(np.complex128)
            ref_shard = orc.apply_observable(want, zoo[i])[rank << (n - g):(rank + 1) << (n - g)]
            check(f"obs[{i}] apply", shard, ref_shard)
            a.close()
        spo = DistributedStateVector(n, dtype, device=local_rank)
        spo.set_state_vector(idx, want.astype(dtype))
        spo.apply_observable(q.Observable.sparse(m.indptr, m.indices, m.data))
        spo.canonicalize()
        ref_shard = orc.csr_matvec(m.indptr, m.indices, m.data, want)[rank << (n - g):(rank + 1) << (n - g)]
        check("sparse apply", spo.local_state().astype(np.complex128), ref_shard)
        spo.close()
        sv.close()

        # adjoint Jacobian: every parametric gate family, wires hitting global and local qubits
        aops = [{"name": "RX", "wires": [0], "params": [0.3]}, {"name": "RY", "wires": [n - 1], "params": [-0.7]},
                {"name": "CNOT", "wires": [0, 1], "params": []}, {"name": "RZ", "wires": [0], "params": [1.1]},
                {"name": "PhaseShift", "wires": [0], "params": [0.2]}, {"name": "CRX", "wires": [0, 2], "params": [0.5]},
                {"name": "CRY", "wires": [3, 0], "params": [-0.4]}, {"name": "CRZ", "wires": [1, 0], "params": [0.9]},
                {"name": "IsingXX", "wires": [0, n - 1], "params": [0.6]}, {"name": "IsingYY", "wires": [1, 0], "params": [0.8]},
                {"name": "IsingZZ", "wires": [0, 2], "params": [-1.2]}, {"name": "Hadamard", "wires": [0], "params": []},
                {"name": "ControlledPhaseShift", "wires": [0, 1], "params": [0.33]},
                {"name": "SingleExcitation", "wires": [0, 3], "params": [0.21]},
                {"name": "SingleExcitationPlus", "wires": [1, 0], "params": [-0.5]},
                {"name": "DoubleExcitation", "wires": [0, 1, 2, 3], "params": [0.44]},
                {"name": "DoubleExcitationMinus", "wires": [n - 1, 0, 2, 1], "params": [0.15]},
                {"name": "MultiRZ", "wires": [0, 1, n - 1], "params": [0.77]}, {"name": "RX", "wires": [0], "params": [0.9], "adjoint": True},
                {"name": "RY", "wires": [2], "params": [0.1]}]
        n_par = sum(1 for o in aops if o["params"])
        obs_t = [zoo[0], zoo[5], zoo[6], zoo[8], zoo[9], ("Sparse", m.indptr, m.indices, m.data)]
        psi_a = orc.apply_ops(psi0, aops)
        jref = orc.adjoint_jacobian(psi_a, aops, obs_t, list(range(n_par)))
        a = DistributedStateVector(n, dtype, device=local_rank)
        a.set_state_vector(idx, psi0.astype(dtype))
        rec = q.Ops(aops)
        a.apply_ops(rec, fuse=False)
        before = a.norm2()
        jac = a.adjoint_jacobian(rec, [q.Observable.from_tuple(t) for t in obs_t], list(range(n_par)))
        check("adjoint jacobian (all params)", jac, jref)
        tp = [0, 3, 7, n_par - 1]
        jac2 = a.adjoint_jacobian(rec, [q.Observable.from_tuple(obs_t[3])], tp)
        check("adjoint jacobian (subset)", jac2[0], jref[3][tp])
        # the register itself is untouched by the adjoint sweep
        a.canonicalize()
        check("state after adjoint", a.local_state().astype(np.complex128), psi_a[rank << (n - g):(rank + 1) << (n - g)])
        if abs(before - 1) > tol:
            failures.append("norm before adjoint")
        a.close()
    return failures


def pybind_twins(rank, local_rank, world, g):
    """The reference's Python-visible MPI classes (bindings/Bindings.cpp:890-1720) driven the way lightning_gpu.py
    does (:286-294, :429, :840-852), against the oracle."""
    from pennylane_lightning_gpu_b200 import lightning_gpu_qubit_ops as m

    failures = []
    n = g + 7
    mgr = m.MPIManager()
    if (mgr.getRank(), mgr.getSize()) != (rank, world) or mgr.getVendor() != "NCCL":
        failures.append("MPIManager rank/size")
    mgr.Barrier()
    rng = np.random.default_rng(21)
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi0 /= np.linalg.norm(psi0)
    for bits, dtype, tol in (("128", np.complex128, 1e-10), ("64", np.complex64, 1e-5)):
        sv = getattr(m, "LightningGPUMPI_C" + bits)(mgr, m.DevTag(local_rank), 0, g, n - g)
        if (sv.numLocalQubits(), sv.numGlobalQubits(), sv.dataLength()) != (n - g, g, 1 << (n - g)):
            failures.append("MPI state-vector sizes")
        # lightning_gpu.py:429: scatter the full state from rank 0, then HostToDevice of the local shard
        local = np.zeros(1 << (n - g), dtype=dtype)
        mgr.Scatter(psi0.astype(dtype) if rank == 0 else np.zeros(1 << n, dtype=dtype), local, 0)
        sv.HostToDevice(local, False)
        ops = [("Hadamard", [0], []), ("RX", [0], [0.3]), ("CNOT", [0, n - 1], []), ("RY", [n - 1], [0.7]),
               ("CRZ", [1, 0], [0.2]), ("IsingXX", [0, 1], [0.5]), ("Toffoli", [0, 1, 2], []), ("Rot", [0], [0.1, 0.2, 0.3])]
        want = psi0.copy()
        for name, wires, params in ops:
            getattr(sv, name)(wires, False, params)
            want = orc.apply_op(want, name, wires, params)
        mat = orc.gate_matrix("RY", [0.4])
        sv.apply("QubitUnitary", [0], False, [], mat.ravel().astype(dtype))
        want = orc.apply_op(want, "QubitUnitary", [0], [], matrix=mat)

        def check(what, got, ref):
            err = float(np.max(np.abs(np.asarray(got) - np.asarray(ref))))
            if not err <= tol * max(1.0, float(np.max(np.abs(np.asarray(ref))))):
                failures.append(f"pybind C{bits} {what}: err {err:.2e}")
            if rank == 0:
                print(f"[dist_check] pybind C{bits} {what}: err={err:.2e}", flush=True)

        out = np.zeros(1 << (n - g), dtype=dtype)
        sv.DeviceToHost(out, False)
        check("state shard", out, want[rank << (n - g):(rank + 1) << (n - g)])
        check("expval PauliZ(0)", sv.ExpectationValue("PauliZ", [0], [], np.zeros(0, dtype=dtype)), orc.expval_named(want, "PauliZ", [0]))
        words, wires, coeffs = ["XZ", "Y", "ZZ"], [[0, 1], [n - 1], [0, 2]], np.array([0.3, -0.5, 0.9], dtype=dtype)
        check("expval Pauli words", sv.ExpectationValue(words, wires, coeffs), orc.expval_pauli_words(want, words, wires, coeffs))
        check("probability", sv.Probability([0, n - 1]), orc.probs_custatevec_order(want, [0, n - 1]))
        samples = sv.GenerateSamples(n, 100)
        if samples.shape != (100, n) or samples.max() > 1:
            failures.append("GenerateSamples shape")
        # sparse Hamiltonian given on rank 0 only (lightning_gpu.py:840-852)
        sp, (w2, ws2, c2) = workloads.molecular_style_sparse_hamiltonian(n, n_terms=30, n_flip_masks=5, seed=3)
        idt = np.int64 if bits == "128" else np.int32
        if rank == 0:
            got = sv.ExpectationValue(sp.indptr.astype(idt), sp.indices.astype(idt), sp.data.astype(dtype))
        else:
            got = sv.ExpectationValue(np.array([0, 1, 2], dtype=idt), np.array([0, 1], dtype=idt), np.ones(2, dtype=dtype))
        check("sparse expval (rank-0 matrix)", got, orc.expval_csr(want, sp.indptr, sp.indices, sp.data))
        # adjoint Jacobian through the MPI twins
        names = ["RX", "CNOT", "RY", "RZ", "CRX"]
        params = [np.array([0.3]), np.array([]), np.array([-0.7]), np.array([1.1]), np.array([0.5])]
        awires = [[0], [0, 1], [n - 1], [0], [0, 2]]
        aops = [{"name": a, "wires": w, "params": list(p)} for a, w, p in zip(names, awires, params)]
        adj = getattr(m, "AdjointJacobianGPUMPI_C" + bits)()
        rdt = np.float64 if bits == "128" else np.float32
        rec = adj.create_ops_list(names, [p.astype(rdt) for p in params], awires, [False] * 5, [np.zeros(0, dtype=dtype)] * 5)
        sv2 = getattr(m, "LightningGPUMPI_C" + bits)(mgr, m.DevTag(local_rank), 0, g, n - g)
        for a, w, p in zip(names, awires, params):
            getattr(sv2, a)(w, False, list(p))
        Named = getattr(m, "NamedObsGPUMPI_C" + bits)
        Tensor = getattr(m, "TensorProdObsGPUMPI_C" + bits)
        Ham = getattr(m, "HamiltonianGPUMPI_C" + bits)
        obs = [Named("PauliZ", [0]), Ham(np.array([0.4, -0.8], dtype=rdt), [Named("PauliX", [0]), Tensor([Named("PauliZ", [0]), Named("PauliY", [n - 1])])])]
        obs_t = [("Named", "PauliZ", [0]), ("Hamiltonian", [0.4, -0.8], [("Named", "PauliX", [0]), ("TensorProd", [("Named", "PauliZ", [0]), ("Named", "PauliY", [n - 1])])])]
        fin = orc.apply_ops(orc.basis_state(n), aops)
        jref = orc.adjoint_jacobian(fin, aops, obs_t, [0, 1, 2, 3])
        check("adjoint_jacobian", adj.adjoint_jacobian(sv2, obs, rec, [0, 1, 2, 3]), jref)
        check("adjoint_jacobian_serial", adj.adjoint_jacobian_serial(sv2, obs, rec, [0, 1, 2, 3]), jref)
        del sv, sv2
    return failures


def device_mirror(rank, world, g):
    """LightningGPU(wires, mpi=True): the reference's Python device call pattern (lightning_gpu.py:255-324, 392-447,
    638-752, 820-960) on the sharded register."""
    from pennylane_lightning_gpu_b200.lightning_gpu import LightningGPU, Obs, Op

    failures = []
    n = g + 6
    rng = np.random.default_rng(5)
    psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi0 /= np.linalg.norm(psi0)
    dev = LightningGPU(n, mpi=True, c_dtype=np.complex128, seed=7)
    ops = [Op("StatePrep", list(range(n)), [psi0]), Op("RX", [0], [0.4]), Op("CNOT", [0, n - 1]), Op("RY", [1], [0.3]),
           Op("CRY", [n - 1, 0], [0.8]), Op("Rot", [0], [0.1, 0.2, 0.3])]
    dev.apply(ops)
    want = psi0.copy()
    for o in ops[1:]:
        want = orc.apply_op(want, o.name, list(o.wires), list(o.parameters))
    sh = 1 << (n - g)

    def check(what, got, ref, tol=1e-10):
        err = float(np.max(np.abs(np.asarray(got) - np.asarray(ref))))
        if not err <= tol:
            failures.append(f"device {what}: err {err:.2e}")
        if rank == 0:
            print(f"[dist_check] device {what}: err={err:.2e}", flush=True)

    check("state shard", dev.state, want[rank * sh:(rank + 1) * sh])
    ham = Obs("Hamiltonian", coeffs=[0.5, -0.25], terms=[Obs("PauliZ", [0]), Obs("Tensor", terms=[Obs("PauliX", [0]), Obs("PauliY", [n - 1])])])
    ham_t = ("Hamiltonian", [0.5, -0.25], [("Named", "PauliZ", [0]), ("TensorProd", [("Named", "PauliX", [0]), ("Named", "PauliY", [n - 1])])])
    check("expval Hamiltonian", dev.expval(ham), orc.expval_obs(want, ham_t))
    check("expval PauliX(0)", dev.expval(Obs("PauliX", [0])), orc.expval_named(want, "PauliX", [0]))
    check("var PauliZ(0)", dev.var(Obs("PauliZ", [0])), 1 - orc.expval_named(want, "PauliZ", [0]) ** 2)
    check("probability", dev.probability([0, 1, n - 1]), orc.probs(want, [0, 1, n - 1]))
    aops = [Op("RX", [0], [0.4]), Op("CNOT", [0, n - 1]), Op("RY", [1], [0.3]), Op("Rot", [0], [0.1, 0.2, 0.3])]
    jac = dev.adjoint_jacobian(aops, [Obs("PauliZ", [0]), ham])
    ser = [{"name": "RX", "wires": [0], "params": [0.4]}, {"name": "CNOT", "wires": [0, n - 1], "params": []},
           {"name": "RY", "wires": [1], "params": [0.3]}, {"name": "RZ", "wires": [0], "params": [0.1]},
           {"name": "RY", "wires": [0], "params": [0.2]}, {"name": "RZ", "wires": [0], "params": [0.3]}]
    fin = orc.apply_ops(orc.basis_state(n), ser)
    check("adjoint_jacobian", jac, orc.adjoint_jacobian(fin, ser, [("Named", "PauliZ", [0]), ham_t], list(range(5))))
    dev_s = LightningGPU(n, mpi=True, shots=500, seed=11)
    dev_s.apply([Op("Hadamard", [0]), Op("CNOT", [0, n - 1])])
    s = dev_s.generate_samples()
    if s.shape != (500, n) or not np.array_equal(s[:, 0], s[:, n - 1]) or not 150 < s[:, 0].sum() < 350:
        failures.append("device samples of a Bell pair on a global and a local wire")
    return failures


def rank_conditional_read(rank, local_rank, world, g):
    """The reference's DeviceToHost is a purely local copy (StateVectorCudaBase.hpp:104-228), so `if rank == 0:
    print(dev.state)` is legal there.  With the C ABI's default qubit-map policy (eager: every call returns with the
    canonical layout) the same holds here: only rank 0 reads, after gates on global qubits, and nothing dead-locks."""
    failures = []
    n = g + 9
    ops = circuit(n, seed=31, n_gates=40)
    want = orc.apply_ops(orc.basis_state(n), ops)
    sv = DistributedStateVector(n, np.complex128, device=local_rank, lazy_map=False)
    sv.apply_ops(q.Ops(ops), fuse=True)
    if sv.qubit_map() != list(range(n)):
        failures.append("eager policy left a permuted qubit map")
    if rank == 0:
        shard = sv.d2h()  # rank 0 only
        err = float(np.max(np.abs(shard - want[: 1 << (n - g)])))
        print(f"[dist_check] rank-0-only read after global gates: err={err:.2e}", flush=True)
        if err > 1e-10:
            failures.append(f"rank-0-only read: err {err:.2e}")
    z = sv.expval_named("PauliX", [0]).real  # a measurement on a global qubit, then again a one-sided read
    if abs(z - orc.expval_named(want, "PauliX", [0])) > 1e-10:
        failures.append("expval after eager apply")
    if rank == world - 1:
        shard = sv.d2h()
        if float(np.max(np.abs(shard - want[rank << (n - g):]))) > 1e-10:
            failures.append("last-rank-only read")
    sv.close()
    return failures


def config5_generator_at_28_qubits(rank, local_rank, world, g):
    """BASELINE config 5's circuit generator (4 layers of a random one-qubit rotation on every wire and CNOTs on
    a random perfect matching, default_rng(99)) at n = 28 on all GPUs of the box vs ONE GPU vs the CPU restatement
    (oracle/lq_port.c, itself pinned to the NumPy oracle and the reference's golden vectors).  Pattern:
    src/tests/mpi/Test_StateVectorCudaMPI_Param.cpp:59-125 (sharded against single device)."""
    failures = []
    n = int(os.environ.get("DIST_CHECK_BIG_N", "28"))
    ops = workloads.random_layer_circuit(n, layers=4, seed=99)
    rec = q.Ops(ops)
    one = q.StateVector(n, np.complex128, device=local_rank)
    one.set_basis_state(0)
    one.apply_ops(rec, fuse=True)
    sharded = DistributedStateVector(n, np.complex128, device=local_rank)
    sharded.apply_ops(rec, fuse=True)
    n_swaps = sharded.swap_stats()[0]
    n_oop, n_carried = sharded.fused_exchange_stats()
    sh = 1 << (n - g)
    got = sharded.d2h()
    ref = one.d2h()
    err = float(np.max(np.abs(got - ref[rank * sh:(rank + 1) * sh])))
    t = torch.tensor([err], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = float(t)
    if err > 1e-10:
        failures.append(f"config-5 generator n={n}: {world} GPUs vs 1 GPU err {err:.2e}")
    err_cpu = -1.0
    if rank == 0:
        from oracle import lq_port as lq

        lq.set_num_threads(os.cpu_count() or 1)
        st = lq.LQState(n)
        st.apply_ops(ops)
        err_cpu = float(np.max(np.abs(st.sv - ref)))
        if err_cpu > 1e-10:
            failures.append(f"config-5 generator n={n}: 1 GPU vs CPU restatement err {err_cpu:.2e}")
        print(f"[dist_check] config-5 generator n={n} ({len(ops)} gates): {world} GPUs vs 1 GPU err={err:.2e}, 1 GPU vs "
              f"lq_port err={err_cpu:.2e}, exchanges in place={n_swaps} through second buffer={n_oop} carried={n_carried}", flush=True)
    del one
    sharded.close()
    return failures


def config3_adjoint_sharded(rank, local_rank, world, g):
    """BASELINE config 3 (hardware-efficient ansatz, 4 layers, 100-term Pauli Hamiltonian) at 26 qubits on the sharded
    register: the layered reverse sweep with the batched generator kernel and fused U^dagger sweeps over lambda and the
    bra (AdjointDiffGPUMPI.hpp:248-334 undoes one gate at a time) against the single-GPU engine on the same circuit."""
    import time

    failures = []
    n = int(os.environ.get("DIST_CHECK_CONFIG3_N", "26"))
    ops, n_par = workloads.hardware_efficient_ansatz(n, layers=4, seed=11)
    words, wires, coeffs = workloads.random_pauli_hamiltonian(n, 100, seed=5)
    ham = q.Observable.from_tuple(workloads.hamiltonian_tuple(words, wires, coeffs))
    rec = q.Ops(ops)
    one = q.StateVector(n, np.complex128, device=local_rank)
    one.apply_ops(rec, fuse=True)
    want = one.adjoint_jacobian(rec, [ham], list(range(n_par)))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    want = one.adjoint_jacobian(rec, [ham], list(range(n_par)))
    torch.cuda.synchronize()
    t_one = time.perf_counter() - t0
    e_one = one.expval(ham)
    del one
    sv = DistributedStateVector(n, np.complex128, device=local_rank)
    sv.apply_ops(rec, fuse=True)
    e = sv.expval(ham)
    jac = sv.adjoint_jacobian(rec, [ham], list(range(n_par)))
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    jac = sv.adjoint_jacobian(rec, [ham], list(range(n_par)))
    torch.cuda.synchronize()
    dist.barrier()
    t_sh = time.perf_counter() - t0
    scale = max(1.0, float(np.sum(np.abs(coeffs))))  # a gradient of a sum of 100 weighted terms
    err = float(np.max(np.abs(jac - want)))
    if err > 1e-10 * scale or abs(e - e_one) > 1e-10 * scale:
        failures.append(f"config-3 adjoint n={n}: Jacobian err {err:.2e}, expval err {abs(e - e_one):.2e}")
    if rank == 0:
        print(f"[dist_check] config-3 adjoint n={n} ({n_par} parameters, 100 terms): {world} GPUs {t_sh * 1e3:.1f} ms vs 1 GPU "
              f"{t_one * 1e3:.1f} ms, Jacobian err={err:.2e} expval err={abs(e - e_one):.2e}", flush=True)
    sv.close()
    return failures


def fused_exchange_check(rank, local_rank, world, g):
    """Exchanges through the second buffer (the default where it fits; QSV_DIST_FUSED_SWAP=0 switches it off), carried by the sweep before them where
    the exchanged bit is not one of its tile bits.  20 local qubits so that sweeps can carry them; 13 so that the
    copy-pass form runs too; repeated application (odd and even numbers of exchanges, register back home each time)."""
    failures = []
    os.environ["QSV_DIST_FUSED_SWAP"] = "1"
    for dtype, tol in ((np.complex128, 1e-10), (np.complex64, 1e-5)):
        for n_local in (13, 20):
            n_total = n_local + g
            ops = circuit(n_total, seed=100 + n_total)
            want = orc.basis_state(n_total)
            sv = DistributedStateVector(n_total, dtype, device=local_rank)
            for rep in range(3):
                sv.apply_ops(q.Ops(ops), fuse=True)
                want = orc.apply_ops(want, ops)
            n_oop, n_carried = sv.fused_exchange_stats()
            nrm = sv.norm2()
            sv.canonicalize()
            shard = torch.from_numpy(sv.local_state().astype(np.complex128).view(np.float64)).cuda()
            parts = [torch.empty_like(shard) for _ in range(world)]
            dist.all_gather(parts, shard)
            full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
            err = float(np.max(np.abs(full - want)))
            tag = f"fused exchange dtype={np.dtype(dtype).name} n_local={n_local}"
            if err > tol * 30 or abs(nrm - 1) > tol * 100 or n_oop == 0:
                failures.append(f"{tag}: state err {err:.2e}, norm {nrm}, out-of-place exchanges {n_oop}")
            if rank == 0:
                print(f"[dist_check] {tag}: err={err:.2e} exchanges={n_oop} carried_by_sweeps={n_carried}", flush=True)
            sv.close()
    os.environ.pop("QSV_DIST_FUSED_SWAP", None)
    return failures


def main():
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = int(math.log2(world))
    failures = []
    if os.environ.get("DIST_CHECK_FUSED") == "1":
        failures = fused_exchange_check(rank, local_rank, world, g)
        ok = torch.tensor([0 if failures else 1], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if rank == 0:
            print("DIST_CHECK", "PASS" if int(ok) == 1 else "FAIL", failures, flush=True)
        dist.destroy_process_group()
        sys.exit(0 if int(ok) == 1 else 1)
    for dtype, tol in ((np.complex128, 1e-10), (np.complex64, 1e-5)):
        for n_total in (g + 6, g + 13):
            n_local = n_total - g
            ops = circuit(n_total, seed=n_total)
            want = orc.apply_ops(orc.basis_state(n_total), ops)
            want2 = orc.apply_ops(want, ops)
            # exchange schedule over the dependency DAG (default) and in program order (QSV_DIST_DAG=0, read per call)
            # ... and with exchanges fused into sweeps through a second buffer (default) or in place (QSV_DIST_FUSED_SWAP=0)
            for fuse, chunk, dag, fx in ((False, 0, "1", "1"), (True, 0, "1", "1"), (False, 1 << 12, "1", "1"),
                                         (True, 1 << 12, "1", "0"), (True, 0, "0", "1"), (False, 1 << 12, "0", "0"),
                                         (True, 0, "1", "0")):
                os.environ["QSV_DIST_DAG"] = dag
                os.environ["QSV_DIST_FUSED_SWAP"] = fx
                sv = DistributedStateVector(n_total, dtype, device=local_rank, chunk_bytes=chunk)
                sv.apply_ops(q.Ops(ops), fuse=fuse)
                n_swaps, nbytes, ms = sv.swap_stats()
                # measurements in the permuted layout
                words = ["X", "Z", "XY", "ZZ", "YXZ"]
                wires = [[0], [0], [0, n_total - 1], [1, 2], [1, 0, 3]]
                coeffs = [0.5, -1.0, 0.3, 0.8, 1.2]
                ev = sv.expval_pauli_words(words, wires, coeffs)
                ev_want = orc.expval_pauli_words(want, words, wires, coeffs)
                nrm = sv.norm2()
                sv.canonicalize()
                assert sv.qubit_map() == list(range(n_total))
                shard = torch.from_numpy(sv.local_state().astype(np.complex128).view(np.float64)).cuda()
                parts = [torch.empty_like(shard) for _ in range(world)]
                dist.all_gather(parts, shard)
                full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
                err = float(np.max(np.abs(full - want)))
                tag = f"dtype={np.dtype(dtype).name} n={n_total} fuse={fuse} chunk={chunk} dag={dag} fused_exchange={fx}"
                if err > tol or abs(ev - ev_want) > tol * max(1.0, abs(ev_want)) or abs(nrm - 1) > tol:
                    failures.append(f"{tag}: state err {err:.2e}, expval {ev} vs {ev_want}, norm {nrm}")
                if rank == 0:
                    print(f"[dist_check] {tag}: err={err:.2e} expval_err={abs(ev - ev_want):.2e} swaps={n_swaps}", flush=True)
                if chunk == 0:
                    # the same circuit twice more: sv was canonicalised above, so the first of them starts from
                    # the identity map again and the second from the map that one leaves behind (as in bench.py)
                    sv.apply_ops(q.Ops(ops), fuse=fuse)
                    sv.apply_ops(q.Ops(ops), fuse=fuse)
                    third = orc.apply_ops(want2, ops)
                    sv.canonicalize()
                    shard = torch.from_numpy(sv.local_state().astype(np.complex128).view(np.float64)).cuda()
                    parts = [torch.empty_like(shard) for _ in range(world)]
                    dist.all_gather(parts, shard)
                    full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
                    err = float(np.max(np.abs(full - third)))
                    if err > tol:
                        failures.append(f"{tag} repeated: state err {err:.2e}")
                    if rank == 0:
                        print(f"[dist_check] {tag} repeated: err={err:.2e}", flush=True)
                sv.close()
            os.environ.pop("QSV_DIST_DAG", None)
            os.environ.pop("QSV_DIST_FUSED_SWAP", None)
    failures += rank_conditional_read(rank, local_rank, world, g)
    if os.environ.get("DIST_CHECK_N28", "1") == "1":
        failures += config5_generator_at_28_qubits(rank, local_rank, world, g)
    failures += measurements_and_adjoint(rank, local_rank, world, g)
    if os.environ.get("DIST_CHECK_CONFIG3", "1") == "1":
        failures += config3_adjoint_sharded(rank, local_rank, world, g)
    failures += pybind_twins(rank, local_rank, world, g)
    failures += device_mirror(rank, world, g)
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(ok) == 1 else "FAIL", failures, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok) == 1 else 1)


if __name__ == "__main__":
    main()
