"""GPU diagnostic for the register-tile kernel: runs the feature-switch circuit under several environments, twice each,
and reports determinism, the number of wrong amplitudes and which index bits the wrong ones have in common."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from oracle import np_oracle as orc  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

n = 15
ops = [{"name": "CNOT", "wires": [i, i + 1], "params": []} for i in range(n - 1)]
ops += [{"name": "PauliX", "wires": [3], "params": []}, {"name": "SWAP", "wires": [0, n - 1], "params": []},
        {"name": "CNOT", "wires": [n - 1, 0], "params": []}]
ops += workloads.random_gate_circuit(n, 150, 77)
ladder, _ = workloads.hardware_efficient_ansatz(n, layers=2, seed=3)
ops += ladder
ops += [{"name": "SWAP", "wires": [2, 9], "params": []}, {"name": "CNOT", "wires": [9, 2], "params": []},
        {"name": "PauliX", "wires": [n - 1], "params": []}]
rng = np.random.default_rng(21)
psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
psi /= np.linalg.norm(psi)
KEYS = ("QSV_REGS_FOLD", "QSV_REGS_UDIAG", "QSV_REGS_DAG", "QSV_REGS_DIAG1", "QSV_REGS_PREFETCH", "QSV_REGS_RB",
        "QSV_REGS_UCONST", "QSV_REGS_MAX_GATES")


def run(env, sub_ops, dtype=np.complex128):
    for k in KEYS:
        os.environ.pop(k, None)
    os.environ.update(env)
    sv = q.StateVector(n, dtype)
    sv.h2d(psi.astype(dtype))
    sv.apply_ops(q.Ops(sub_ops), fuse=True)
    return sv.d2h(), sv.last_apply_stats()


envs = [{}, {"QSV_REGS_DAG": "0"}, {"QSV_REGS_PREFETCH": "3"}, {"QSV_REGS_RB": "3"}, {"QSV_REGS_UCONST": "0"},
        {"QSV_REGS_DAG": "0", "QSV_REGS_UCONST": "0"}, {"QSV_REGS_DAG": "0", "QSV_REGS_FOLD": "0"},
        {"QSV_REGS_DAG": "0", "QSV_REGS_FOLD": "0", "QSV_REGS_UCONST": "0"},
        {"QSV_REGS_DAG": "0", "QSV_REGS_UDIAG": "0", "QSV_REGS_DIAG1": "0"}]
for env in envs:
    # shortest failing prefix of the circuit
    lo, hi = 0, len(ops)
    want_full = orc.apply_ops(psi.copy(), ops)
    a, st = run(env, ops)
    b, _ = run(env, ops)
    err = np.abs(a - want_full)
    bad = np.nonzero(err > 1e-9)[0]
    print(env, "stats", st, "max err", err.max(), "n_bad", bad.size, "deterministic", bool(np.array_equal(a, b)), flush=True)
    if bad.size == 0:
        continue
    while hi - lo > 1:
        mid = (lo + hi) // 2
        got, _ = run(env, ops[:mid])
        if np.max(np.abs(got - orc.apply_ops(psi.copy(), ops[:mid]))) > 1e-9:
            hi = mid
        else:
            lo = mid
    got, st = run(env, ops[:hi])
    want = orc.apply_ops(psi.copy(), ops[:hi])
    bad = np.nonzero(np.abs(got - want) > 1e-9)[0]
    common1 = np.bitwise_and.reduce(bad) if bad.size else 0
    common0 = np.bitwise_and.reduce(~bad) & ((1 << n) - 1) if bad.size else 0
    print("   shortest failing prefix:", hi, "ops; last op", {k: v for k, v in ops[hi - 1].items() if k != "matrix"}, "stats", st)
    print("   bad amplitudes:", bad.size, "bits always 1: %s, always 0: %s" % (bin(common1), bin(common0)), "first bad", bad[:8])
    for mg in ("1", "2", "4", "8"):
        got, st = run(dict(env, QSV_REGS_MAX_GATES=mg), ops[:hi])
        print("   max gates per sweep", mg, "->", np.max(np.abs(got - want)), st)
