"""BASELINE config 3 (24-qubit hardware-efficient ansatz, 100-term Pauli Hamiltonian, adjoint Jacobian) timed through
the ctypes layer, for A/B runs by environment (QSV_GENS_TB, QSV_GENS_L12/L13, QSV_ADJOINT_DEFER, QSV_REGS_PERSIST ...);
prints seconds, kernel launches and a finite-difference check of two Jacobian entries."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
ops, n_par = workloads.hardware_efficient_ansatz(n, layers=4, seed=11)
words, wires, coeffs = workloads.random_pauli_hamiltonian(n, 100, seed=5)
ham = q.Observable.from_tuple(workloads.hamiltonian_tuple(words, wires, coeffs))
rec = q.Ops(ops)
sv = q.StateVector(n, np.complex128)


def run():
    sv.set_basis_state(0)
    sv.apply_ops(rec, fuse=True)
    e = sv.expval(ham)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    jac = sv.adjoint_jacobian(rec, [ham], list(range(n_par)))
    torch.cuda.synchronize()
    return e, jac, time.perf_counter() - t0


run()
best = min(run()[2] for _ in range(5))
e, jac, _ = run()
launches = sv.last_apply_stats()
idx_par = [i for i, o in enumerate(ops) if o["params"]]
fd = 0.0
for p in (0, n_par // 2, n_par - 1):
    vals = []
    for sgn in (+1, -1):
        ops2 = [dict(o) for o in ops]
        ops2[idx_par[p]] = dict(ops2[idx_par[p]], params=[ops[idx_par[p]]["params"][0] + sgn * 1e-4])
        sv.set_basis_state(0)
        sv.apply_ops(q.Ops(ops2), fuse=True)
        vals.append(sv.expval(ham))
    fd = max(fd, abs((vals[0] - vals[1]) / 2e-4 - jac[0, p]))
print("adjoint", n, "qubits env", {k: v for k, v in os.environ.items() if k.startswith("QSV_")},
      "jacobian_s %.5f" % best, "launches", launches, "expval %.12f" % e, "jac_norm %.12f" % float(np.linalg.norm(jac)),
      "fd_err %.2e" % fd, "jac_sum %.14f" % float(jac.sum()))
