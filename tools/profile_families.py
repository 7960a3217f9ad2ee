"""One launch of every kernel family that is not the fused sweep, in its CURRENT default launch shape, at sizes that are
HBM-resident (28-30 qubits complex128), bracketed by cudaProfilerStart/Stop:

    ncu --profile-from-start off --set full --clock-control none --import-source on -o gpurun_out/r2_families python tools/profile_families.py

Prints the algorithmic bytes of every step so that the ncu durations turn into fractions of the measured HBM peak."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

n = int(os.environ.get("PROFILE_QUBITS", "30"))
B, N = 16, 1 << n
buf = torch.empty(N * 2, dtype=torch.float64, device="cuda")
buf.normal_()
buf.mul_(1.0 / float(buf.norm()))
sv = q.StateVector(n, np.complex128, external_ptr=buf.data_ptr())
rng = np.random.default_rng(7)
u2 = workloads.haar_unitary(rng, 2)
u4 = workloads.haar_unitary(rng, 4)
u8 = workloads.haar_unitary(rng, 8)
h4 = u4 + u4.conj().T
steps = [
    ("k_apply_dense 1q, bit 12 (RX wire n-13)", 2 * B * N, lambda: sv.apply("RX", [n - 13], [0.3])),
    ("k_apply_dense 1q, bit 29 (RX wire 0)", 2 * B * N, lambda: sv.apply("RX", [0], [0.3])),
    ("k_apply_dense_bit0 (RX wire n-1)", 2 * B * N, lambda: sv.apply("RX", [n - 1], [0.3])),
    ("k_apply_dense 1q + control (CNOT 0 -> 1)", B * N, lambda: sv.apply("CNOT", [0, 1])),
    ("k_apply_dense 2q (QubitUnitary wires 3,17)", 2 * B * N, lambda: sv.apply_matrix(u4, [3, 17])),
    ("k_apply_dense 3q (QubitUnitary wires 2,9,20)", 2 * B * N, lambda: sv.apply_matrix(u8, [2, 9, 20])),
    ("k_apply_diag (RZ wire 5)", 2 * B * N, lambda: sv.apply("RZ", [5], [1.1])),
    ("k_apply_diag (CZ 4,21: quarter of the state)", B * N // 2, lambda: sv.apply("CZ", [4, 21])),
    ("k_bra_dense_ket (expval Hermitian 2q)", B * N, lambda: sv.expval_matrix(h4, [1, 11])),
    ("k_bra_pauli_ket / gens tile (expval 3 Pauli words)", B * N, lambda: sv.expval_pauli_words(["XZ", "Y", "ZZ"], [[0, 5], [n - 1], [2, 3]], [0.3, -0.5, 0.9])),
    ("k_probs_small (3 wires)", B * N, lambda: sv.probs([0, 7, n - 1])),
    ("k_probs_large (14 wires)", B * N, lambda: sv.probs(list(range(14)))),
    ("k_block_mass + k_sample_in_block (1000 shots)", B * N, lambda: sv.sample(np.random.default_rng(1).random(1000))),
    ("k_fill_basis (setBasisState)", B * N, lambda: sv.set_basis_state(5)),
]
for _, _, fn in steps[:-1]:
    fn()  # warm-up (allocations, function attributes)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for name, nbytes, fn in steps:
    fn()
    torch.cuda.synchronize()
    print(f"STEP {name}: algorithmic {nbytes / 1e9:.3f} GB", flush=True)
torch.cuda.cudart().cudaProfilerStop()

# sparse Hamiltonian (config 4 generator at 20 qubits: CSR 0.78 GB device-resident)
n4 = int(os.environ.get("PROFILE_CSR_QUBITS", "20"))
m, _ = workloads.molecular_style_sparse_hamiltonian(n4, 400, 30, seed=3)
s4 = q.StateVector(n4, np.complex128)
s4.apply_ops(q.Ops(workloads.hardware_efficient_ansatz(n4, layers=2, seed=11)[0]), fuse=True)
obs = q.Observable.sparse(m.indptr, m.indices, m.data)
s4.expval(obs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e = s4.expval(obs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
N4 = 1 << n4
print(f"STEP k_csr (expval, {n4} qubits, nnz {m.nnz}): algorithmic {(m.nnz * 24 + (N4 + 1) * 8 + 2 * 16 * N4) / 1e9:.3f} GB  <H> = {e:.6f}")
