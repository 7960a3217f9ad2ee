"""Cost model of the fused register-tile executor (CPU only: program structure from tests/native/regs_emu.cu,
regs_emu_sweep_stats), fitted to 17 measurements on B200 at 30 qubits complex128: the per-launch times of the 8 sweeps of
the config-2 circuit (profiles/r1_launches_fused.csv, stable to 0.01 ms) and the step times of config 2 / a
hardware-efficient ansatz / StronglyEntanglingLayers with QSV_REGS_MMA = 0, 1, 2 and 5 low tile bits
(profiles/r1_ab_mma.txt, rescaled by 111.7 / 112.9 to the final build):

    ms = 3.39 * sweeps + 0.96 * passes + 0.85 * (passes with a tensor-core block)
         + 1.03 * (4x4 gates on register bits) + 0.56 * (2x2-type gates on register bits) + 0.20 * (diagonal gates)

Largest error 3.5 % on a single sweep, 2.5 % on a circuit.  Reading: a gate on register bits costs what the FP64 pipe needs
(2^30 amplitudes x 8 DFMA = 0.46 ms at 64 DFMA/clk/SM for a 2x2, 0.92 ms for a 4x4); a register pass costs what the
shared-memory bandwidth needs for the transposition (2 x 16 GiB at 128 B/clk/SM = 0.92 ms); a tensor-core block costs
0.85 ms however many gates were multiplied into it; 3.4 ms per sweep is HBM time the arithmetic does not hide (5.3 ms
would be all of it).  The pass scheduler's beam search (csrc/tile_regs.cu, QSV_REGS_BEAM) and the multi-start sweep
packing (csrc/tile_kernels.cu, QSV_REGS_PACK_TRIES) minimise this model.
Usage: python tools/sweep_cost_model.py [--fit]"""
import ctypes as C
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

COEF = np.array([3.39, 0.96, 0.85, 1.03, 0.56, 0.20])  # sweep, pass, tensor-core block, 4x4 gate, 2x2-type gate, diagonal gate
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
        total += LONE_GATE_MS if r[0] == 0 else float(COEF @ np.array([1, r[0], r[5], r[2], r[3], r[4]]))
    return total


def main():
    lib = _emu()
    ops = workloads.random_gate_circuit(30, 200, 2024)
    rows = sweep_rows(lib, ops)
    if "--fit" in sys.argv:
        # the 8 per-sweep times alone do not separate a pass from its tensor-core block (every pass has one); the full
        # fit over the 17 measurements is in the docstring.  Here: residuals of the model on the 8 sweeps.
        a = np.array([[1, r[0], r[5], r[2], r[3], r[4]] for r in rows], float)
        print("per-sweep residuals (ms)", np.round(a @ COEF - np.array(MEASURED_MS), 2))
    print("config 2: sweeps", len(rows), "passes", sum(r[0] for r in rows), "tensor-core blocks", sum(r[5] for r in rows),
          "predicted ms", round(predict(rows), 1), "(measured 111.7 with QSV_REGS_BEAM=1 QSV_REGS_PACK_TRIES=1)")
    parts = COEF * np.array([len(rows), sum(r[0] for r in rows), sum(r[5] for r in rows), sum(r[2] for r in rows),
                             sum(r[3] for r in rows), sum(r[4] for r in rows)])
    print("  of which: sweeps %.1f, passes %.1f, tensor-core blocks %.1f, 4x4 gates %.1f, 2x2 gates %.1f, diagonal gates %.1f ms"
          % tuple(parts))
    for low in (3, 5, 6):
        r = sweep_rows(lib, ops, low=low)
        print(f"low tile bits {low}: sweeps {len(r)} passes {sum(x[0] for x in r)} predicted {predict(r):.1f}")
    hea, _ = workloads.hardware_efficient_ansatz(30, layers=4, seed=11)
    r = sweep_rows(lib, hea)
    print(f"30-qubit hardware-efficient ansatz: sweeps {len(r)} passes {sum(x[0] for x in r)} predicted {predict(r):.1f} "
          "(measured 142.4 with QSV_REGS_BEAM=1 QSV_REGS_PACK_TRIES=1)")
    sel, _ = workloads.strongly_entangling_layers(30, layers=2, seed=1337)
    r = sweep_rows(lib, sel)
    print(f"30-qubit StronglyEntanglingLayers x 2: sweeps {len(r)} passes {sum(x[0] for x in r)} predicted {predict(r):.1f} "
          "(measured 77.5 with QSV_REGS_BEAM=1 QSV_REGS_PACK_TRIES=1)")


if __name__ == "__main__":
    main()
