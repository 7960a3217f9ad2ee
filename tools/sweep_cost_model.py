"""Cost model of the fused register-tile executor, fitted to the per-launch times of the 30-qubit complex128 config-2
circuit (profiles/r1_launches_fused.csv: the 8 sweeps of a step, stable to 0.01 ms) against the structure of the programs
the host side builds for them (tests/native/regs_emu.cu: regs_emu_sweep_stats -- CPU only, no GPU needed):

    ms per sweep = 3.12 + 1.88 * passes + 1.03 * (4x4 blocks) + 0.56 * (2x2-type gates) + 0.20 * (diagonal gates)

Residuals <= 0.3 ms per sweep; out of sample: L = 5 low tile bits 112.9 predicted / 113.8 measured, 30-qubit
hardware-efficient ansatz 138.7 / 142.4.  Reading: the marginal cost of a gate is at the FP64 roofline already (2^30
amplitudes x 8 DFMA = 0.46 ms at 64 DFMA/clk/SM for a 2x2, 0.92 ms for a 4x4); what is left is 1.9 ms per register pass
(shared-memory transposition, twice the 0.92 ms its bandwidth needs) and ~3 ms per sweep of HBM time that the arithmetic
does not hide.  Usage: python tools/sweep_cost_model.py [--fit] -- prints the prediction for planner variants."""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

COEF = np.array([3.12, 1.88, 1.03, 0.56, 0.20])  # const, per pass, per 4x4 block, per 2x2-type gate, per diagonal gate
MEASURED_MS = [19.86, 12.25, 12.61, 13.77, 17.25, 14.50, 10.42, 11.14]  # profiles/r1_launches_fused.csv, one step
LONE_GATE_MS = 5.3  # one HBM sweep


def _emu():
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "native", "build_emu.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    lib = C.CDLL(m.build())
    lib.regs_emu_sweep_stats.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.c_int]
    return lib


def sweep_rows(lib, ops, n=30, low=4, dag=1, rb=4):
    rec = q.Ops(ops)
    out = (C.c_int64 * (8 * 512))()
    k = lib.regs_emu_sweep_stats(rec._h, n, 1, rb, low, dag, out, 512)
    return [list(out[8 * i:8 * i + 8]) for i in range(k)]


def predict(rows):
    total = 0.0
    for r in rows:
        total += LONE_GATE_MS if r[0] == 0 else float(COEF @ np.array([1, r[0], r[2], r[3], r[4]]))
    return total


def main():
    lib = _emu()
    ops = workloads.random_gate_circuit(30, 200, 2024)
    rows = sweep_rows(lib, ops)
    if "--fit" in sys.argv:
        a = np.array([[1, r[0], r[2], r[3], r[4]] for r in rows], float)
        coef, *_ = np.linalg.lstsq(a, np.array(MEASURED_MS), rcond=None)
        print("fitted coefficients", np.round(coef, 2), "residuals", np.round(a @ coef - np.array(MEASURED_MS), 2))
    print("config 2, default plan: sweeps", len(rows), "passes", sum(r[0] for r in rows), "predicted ms", round(predict(rows), 1),
          "(measured 111.7)")
    parts = COEF * np.array([len(rows), sum(r[0] for r in rows), sum(r[2] for r in rows), sum(r[3] for r in rows),
                             sum(r[4] for r in rows)])
    print("  of which: per-sweep constant %.1f, passes %.1f, 4x4 blocks %.1f, 2x2 gates %.1f, diagonal gates %.1f ms" % tuple(parts))
    for low in (3, 5, 6):
        r = sweep_rows(lib, ops, low=low)
        print(f"low tile bits {low}: sweeps {len(r)} passes {sum(x[0] for x in r)} predicted {predict(r):.1f}")
    hea, _ = workloads.hardware_efficient_ansatz(30, layers=4, seed=11)
    r = sweep_rows(lib, hea)
    print(f"30-qubit hardware-efficient ansatz: sweeps {len(r)} passes {sum(x[0] for x in r)} predicted {predict(r):.1f} (measured 142.4)")


if __name__ == "__main__":
    main()
