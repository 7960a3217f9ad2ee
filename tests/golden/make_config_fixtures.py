"""Generate fixtures for BASELINE.json's configs with the NumPy oracle (run here, on CPU):

    python tests/golden/make_config_fixtures.py

config1_sel20.json : 20-qubit StronglyEntanglingLayers (2 layers), <Z0>, 120-parameter adjoint Jacobian,
                     weights = default_rng(1337).uniform(0, 2pi, (2, 20, 3))  
config3_vqe.json   : reduced C3 (hardware-efficient ansatz + random Pauli Hamiltonian) at 14 qubits, same
                     generator as bench/configs.py uses at 24 qubits
The oracle is pinned by the reference's golden vectors (tests/test_oracle_golden.py); these files
only carry its outputs to the GPU box, where /root/reference and long CPU runs are unavailable.
"""
import json
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import np_oracle as orc  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402


def pairs(z):
    z = np.asarray(z)
    return np.stack([z.real, z.imag], -1).tolist()


def config1():
    n, layers, seed = 20, 2, 1337
    w = np.random.default_rng(seed).uniform(0, 2 * math.pi, (layers, n, 3))
    ops = orc.strongly_entangling_layers(w)
    psi = orc.apply_ops(orc.basis_state(n), ops)
    obs = [("Named", "PauliZ", [0])]
    ev = orc.expval_obs(psi, obs[0])
    jac = orc.adjoint_jacobian(psi, ops, obs, list(range(120)))
    idx = np.random.default_rng(0).choice(1 << n, size=64, replace=False)
    out = {"n": n, "layers": layers, "seed": seed, "n_params": 120, "expval": ev, "jacobian": jac[0].tolist(),
           "state_sample_idx": idx.tolist(), "state_sample": pairs(psi[idx]), "norm": float(np.vdot(psi, psi).real)}
    with open(os.path.join(HERE, "config1_sel20.json"), "w") as f:
        json.dump(out, f)
    print("config1: <Z0> =", ev, " |jac| =", float(np.linalg.norm(jac)))


def config3(n=14):
    ops, n_params = workloads.hardware_efficient_ansatz(n, layers=4, seed=11)
    words, wires, coeffs = workloads.random_pauli_hamiltonian(n, 100, seed=5)
    psi = orc.apply_ops(orc.basis_state(n), ops)
    ham = workloads.hamiltonian_tuple(words, wires, coeffs)
    ev = orc.expval_obs(psi, ham)
    jac = orc.adjoint_jacobian(psi, ops, [ham], list(range(n_params)))
    out = {"n": n, "layers": 4, "n_params": n_params, "n_terms": 100, "expval": ev, "jacobian": jac[0].tolist()}
    with open(os.path.join(HERE, "config3_vqe%d.json" % n), "w") as f:
        json.dump(out, f)
    print("config3: <H> =", ev, " |jac| =", float(np.linalg.norm(jac)))


if __name__ == "__main__":
    config1()
    config3()
