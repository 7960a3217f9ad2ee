"""CUDA-event timings of the measurement kernels at 30 qubits complex128 (probabilities, few-word Pauli expval, sampler),
for A/B runs by environment; prints GB/s against one read of the state."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
N = 1 << n
buf = torch.empty(N * 2, dtype=torch.float64, device="cuda")
buf.normal_()
buf.mul_(1.0 / float(buf.norm()))
sv = q.StateVector(n, np.complex128, external_ptr=buf.data_ptr())


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


cases = [
    ("probs 3 wires (0, 7, n-1)", lambda: sv.probs([0, 7, n - 1])),
    ("probs 1 wire (5)", lambda: sv.probs([5])),
    ("probs 14 wires (0..13)", lambda: sv.probs(list(range(14)))),
    ("probs 12 low wires (n-12..n-1)", lambda: sv.probs(list(range(n - 12, n)))),
    ("expval 3 Pauli words", lambda: sv.expval_pauli_words(["XZ", "Y", "ZZ"], [[0, 5], [n - 1], [2, 3]], [0.3, -0.5, 0.9])),
    ("expval PauliZ(0)", lambda: sv.expval_named("PauliZ", [0])),
    ("sample 1000 shots", lambda: sv.sample(np.random.default_rng(1).random(1000))),
]
for name, fn in cases:
    ms = timed(fn)
    print(f"{name}: {ms:.3f} ms, {16 * N / ms / 1e6:.0f} GB/s of one state read (env {dict((k, v) for k, v in os.environ.items() if k.startswith('QSV_'))})")
p = sv.probs([0, 7, n - 1])
print("probs sum", p.sum())
