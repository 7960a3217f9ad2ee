"""Config-3 adjoint Jacobian once, bracketed by cudaProfilerStart/Stop, for
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file ... python tools/adjoint_profile.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ops, n_par = workloads.hardware_efficient_ansatz(n, layers=4, seed=11)
words, wires, coeffs = workloads.random_pauli_hamiltonian(n, 100, seed=5)
ham = q.Observable.from_tuple(workloads.hamiltonian_tuple(words, wires, coeffs))
rec = q.Ops(ops)
sv = q.StateVector(n, np.complex128)
sv.apply_ops(rec, fuse=True)
sv.adjoint_jacobian(rec, [ham], list(range(n_par)))  # warm-up
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
e = sv.expval(ham)
jac = sv.adjoint_jacobian(rec, [ham], list(range(n_par)))
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("expval", e, "jac norm", float(np.linalg.norm(jac)))
