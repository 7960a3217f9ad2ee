"""Timing of a D1-heavy circuit (hardware-efficient ansatz: RY, RZ on every wire + CNOT ladder per layer) and of a
StronglyEntanglingLayers circuit through the fused executor, complex128; for A/B runs by environment."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
for name, ops in (("hea4", workloads.hardware_efficient_ansatz(n, layers=4, seed=11)[0]),
                  ("sel2", workloads.strongly_entangling_layers(n, 2, 1337)[0])):
    rec = q.Ops(ops)
    sv = q.StateVector(n, np.complex128)
    for _ in range(3):
        sv.apply_ops(rec, fuse=True)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        sv.apply_ops(rec, fuse=True)
    b.record()
    torch.cuda.synchronize()
    print(name, n, "qubits", len(ops), "ops", "env", {k: v for k, v in os.environ.items() if k.startswith("QSV_")},
          "ms per apply %.2f" % (a.elapsed_time(b) / 3), "launches", sv.last_apply_stats())
    del sv
