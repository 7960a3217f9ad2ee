"""A/B of the launch shape of the single-target dense kernel (QSV_DENSE1_SHAPE, csrc/apply_kernels.cu), one process:
parity of every shape against shape 0 on a 20-qubit register (complex128 / complex64, plain, controlled, low and high
target bits), then RX / U2 / CNOT timings on a 30-qubit complex128 register.  Writes gpurun_out/ab_dense1.json and
prints the shape with the best median RX bandwidth whose parity is clean."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

SHAPES = {0: "U=4 NT=256 (default until this A/B)", 1: "U=2 NT=256", 2: "U=8 NT=256", 3: "U=4 NT=128", 4: "U=2 NT=512",
          5: "U=8 NT=128", 6: "U=1 NT=512", 7: "gate (x) identity through the 2-target kernel",
          8: "256-bit accesses U=2 NT=256"}  # 9 / 10 (256-bit, U=4 / U=1) were measured once and removed: profiles/r1_ab_dense1.txt


def set_shape(s):
    os.environ["QSV_DENSE1_SHAPE"] = str(s)


def parity():
    n = 20
    rng = np.random.default_rng(3)
    u2 = workloads.haar_unitary(rng, 2)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    gates = [("RX", [0], [0.3]), ("RX", [n - 1], [0.3]), ("RX", [n - 2], [0.7]), ("RY", [7], [0.4]),
             ("CNOT", [3, 12], []), ("CNOT", [n - 2, n - 1], []), ("CNOT", [n - 1, n - 2], []), ("CRX", [0, n - 1], [0.9]),
             ("Toffoli", [1, 5, 9], []), ("Toffoli", [n - 3, n - 1, n - 2], []), ("Hadamard", [n - 2], [])]
    worst = {}
    for dt, tol in ((np.complex128, 1e-13), (np.complex64, 2e-6)):
        ref = None
        for s in SHAPES:
            set_shape(s)
            try:
                sv = q.StateVector(n, dt)
                sv.h2d(psi.astype(dt))
                for name, wires, params in gates:
                    sv.apply(name, wires, params)
                for w in (0, 1, 10, n - 2, n - 1):
                    sv.apply_matrix(u2, [w])
                out = sv.d2h()
            except Exception as e:  # a broken variant must not take the others down
                print(f"shape {s} failed: {e}", file=sys.stderr)
                worst[s] = float("inf")
                continue
            if ref is None:
                ref = out
            err = float(np.max(np.abs(out - ref)))
            worst[s] = max(worst.get(s, 0.0), err / tol)
    return worst  # <= 1 means within tolerance


def timings():
    n = 30
    buf = torch.empty((1 << n) * 2, dtype=torch.float64, device="cuda")
    chunk = 1 << 26
    for s0 in range(0, buf.numel(), chunk):
        buf[s0:s0 + chunk].normal_()
    buf.mul_(1.0 / np.sqrt(float(1 << (n + 1))))
    sv = q.StateVector(n, np.complex128, external_ptr=buf.data_ptr())
    u2 = workloads.haar_unitary(np.random.default_rng(7), 2)
    full = 2 * 16 * (1 << n)

    def timed(fn, reps=4):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    res = {}
    wires = [0, 5, 10, 15, 20, 23, 26, 28]
    for s in SHAPES:
        set_shape(s)
        try:
            sv.apply("RX", [3], [0.1])
            torch.cuda.synchronize()
        except Exception as e:
            print(f"shape {s} failed: {e}", file=sys.stderr)
            continue
        rx = [full / (timed(lambda: sv.apply("RX", [w], [0.3])) * 1e-3) / 1e9 for w in wires]
        u = [full / (timed(lambda: sv.apply_matrix(u2, [w])) * 1e-3) / 1e9 for w in (0, 12, 24)]
        cn = [full / 2 / (timed(lambda: sv.apply("CNOT", [w, w + 1])) * 1e-3) / 1e9 for w in (3, 14, 25)]
        res[s] = {"shape": SHAPES[s], "rx_gbs": [round(x, 1) for x in rx], "rx_median": float(np.median(rx)),
                  "rx_min": min(rx), "u2_gbs": [round(x, 1) for x in u], "cnot_gbs": [round(x, 1) for x in cn]}
    return res, wires


def main():
    worst = parity()
    res, wires = timings()
    ok = [s for s in SHAPES if worst[s] <= 1.0 and s in res]
    best = max(ok, key=lambda s: res[s]["rx_median"] + 0.5 * float(np.median(res[s]["cnot_gbs"])))
    out = {"rx_wires": wires, "parity_err_over_tol": worst, "shapes": res, "best": best}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/ab_dense1.json", "w") as f:
        json.dump(out, f, indent=1)
    for s in res:
        r = res[s]
        print(f"shape {s} [{r['shape']}] parity {worst[s]:.2g}: RX median {r['rx_median']:.0f} min {r['rx_min']:.0f} "
              f"{r['rx_gbs']} U2 {r['u2_gbs']} CNOT {r['cnot_gbs']}", file=sys.stderr)
    print(best)


if __name__ == "__main__":
    main()
