"""Device-resident CSR expval (config-4 generator) timed for A/B runs by environment (QSV_CSR_LPR)."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 22
m, (words, wires, coeffs) = workloads.molecular_style_sparse_hamiltonian(n, 400, 30, seed=3)
sv = q.StateVector(n, np.complex128)
sv.apply_ops(q.Ops(workloads.hardware_efficient_ansatz(n, layers=4, seed=11)[0]), fuse=True)
obs = q.Observable.sparse(m.indptr, m.indices, m.data)
e = sv.expval(obs)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    t0 = time.perf_counter()
    e = sv.expval(obs)
    torch.cuda.synchronize()
    best = min(best, time.perf_counter() - t0)
N = 1 << n
alg = m.nnz * 24 + (N + 1) * 8 + 2 * 16 * N
print(f"csr expval n={n} nnz={m.nnz} env LPR={os.environ.get('QSV_CSR_LPR', 'auto')}: {best * 1e3:.3f} ms, {alg / best / 1e9:.0f} GB/s, <H>={e:.12f}")
