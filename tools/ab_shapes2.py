"""A/B of the launch shapes of the bit-0, two-target and diagonal gate kernels (QSV_BIT0_SHAPE / QSV_DENSE2_SHAPE /
QSV_DIAG_SHAPE, csrc/apply_kernels.cu) in one process: parity against shape 0 at 18 qubits, timings at 30 qubits."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

KNOBS = {"QSV_BIT0_SHAPE": ["U4 NT256", "U8 NT128", "U2 NT256", "U8 NT256"],
         "QSV_DENSE2_SHAPE": ["U2 NT256", "U4 NT128", "U1 NT256", "U2 NT128"],
         "QSV_DIAG_SHAPE": ["U4 NT256", "U8 NT128", "U8 NT256", "U2 NT256"]}
rng = np.random.default_rng(7)
u2, u4 = workloads.haar_unitary(rng, 2), workloads.haar_unitary(rng, 4)


def circuit(sv, n):
    sv.apply("RX", [n - 1], [0.3]); sv.apply("CNOT", [n - 2, n - 1]); sv.apply("CRY", [2, n - 1], [0.5])
    sv.apply_matrix(u4, [0, 1]); sv.apply_matrix(u4, [n - 2, n - 1]); sv.apply_matrix(u4, [n - 1, 5])
    sv.apply("RZ", [0], [1.1]); sv.apply("RZ", [n - 1], [0.4]); sv.apply("CZ", [3, 4]); sv.apply("IsingZZ", [1, n - 1], [0.2])
    sv.apply("PhaseShift", [n - 2], [0.6]); sv.apply("MultiRZ", [0, 4, 9], [0.8])


def main():
    n = 18
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    ref = {}
    for dt in (np.complex128, np.complex64):
        sv = q.StateVector(n, dt); sv.h2d(psi.astype(dt)); circuit(sv, n); ref[dt] = sv.d2h()
    n = 30
    buf = torch.empty((1 << n) * 2, dtype=torch.float64, device="cuda")
    for s0 in range(0, buf.numel(), 1 << 26):
        buf[s0:s0 + (1 << 26)].normal_()
    buf.mul_(1.0 / np.sqrt(float(1 << (n + 1))))
    big = q.StateVector(n, np.complex128, external_ptr=buf.data_ptr())
    full = 2 * 16 * (1 << n)

    def timed(fn, reps=3):
        fn(); fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    cases = {"QSV_BIT0_SHAPE": [("RX29", lambda: big.apply("RX", [29], [0.3]), full), ("U2_29", lambda: big.apply_matrix(u2, [29]), full),
                                ("CNOT_10_29", lambda: big.apply("CNOT", [10, 29]), full // 2)],
             "QSV_DENSE2_SHAPE": [("U4_0_1", lambda: big.apply_matrix(u4, [0, 1]), full), ("U4_14_15", lambda: big.apply_matrix(u4, [14, 15]), full),
                                  ("U4_3_20", lambda: big.apply_matrix(u4, [3, 20]), full), ("U4_27_28", lambda: big.apply_matrix(u4, [27, 28]), full)],
             "QSV_DIAG_SHAPE": [("RZ0", lambda: big.apply("RZ", [0], [1.1]), full), ("RZ15", lambda: big.apply("RZ", [15], [1.1]), full),
                                ("RZ29", lambda: big.apply("RZ", [29], [1.1]), full), ("CZ_3_4", lambda: big.apply("CZ", [3, 4]), full // 4)]}
    out = {}
    for knob, names in KNOBS.items():
        out[knob] = {}
        for s, label in enumerate(names):
            os.environ[knob] = str(s)
            try:
                err = 0.0
                for dt, tol in ((np.complex128, 1e-13), (np.complex64, 2e-6)):
                    sv = q.StateVector(18, dt); sv.h2d(psi.astype(dt)); circuit(sv, 18)
                    err = max(err, float(np.max(np.abs(sv.d2h() - ref[dt]))) / tol)
                gbs = {c: round(nb / (timed(fn) * 1e-3) / 1e9, 1) for c, fn, nb in cases[knob]}
            except Exception as e:
                print(knob, s, "failed", e, file=sys.stderr)
                continue
            out[knob][s] = {"shape": label, "parity_err_over_tol": err, "gbs": gbs}
            print(knob, s, label, f"parity {err:.2g}", gbs, flush=True)
        os.environ[knob] = "0"
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/ab_shapes2.json", "w"), indent=1)


if __name__ == "__main__":
    main()
