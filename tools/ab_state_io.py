"""Host <-> device copy rates of a 28-qubit complex128 state (4 GiB): pageable NumPy memory through the staged pipeline
(csrc/state_io.cu) and through one plain cudaMemcpy (QSV_IO_STAGED=0 in a child process), and pinned memory."""
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import pennylane_lightning_gpu_b200 as q  # noqa: E402

n = int(os.environ.get("AB_IO_QUBITS", "28"))
sv = q.StateVector(n, np.complex128)
host = np.empty(1 << n, dtype=np.complex128)
host.view(np.float64)[:] = 1.0  # touch every page
gb = host.nbytes / 1e9


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return best


mode = "staged" if os.environ.get("QSV_IO_STAGED", "1") != "0" else "plain cudaMemcpy"
t_h2d = timed(lambda: sv.h2d(host))
t_d2h = timed(lambda: sv.d2h(host))
print(f"pageable {mode} threads={os.environ.get('QSV_IO_THREADS', 'default')} cores={os.cpu_count()}: "
      f"H2D {gb / t_h2d:.1f} GB/s, D2H {gb / t_d2h:.1f} GB/s ({gb:.1f} GB)")
if mode == "staged" and "AB_IO_CHILD" not in os.environ:
    pinned = torch.empty(1 << n, dtype=torch.complex128).pin_memory()
    hp = pinned.numpy()
    print(f"pinned: H2D {gb / timed(lambda: sv.h2d(hp)):.1f} GB/s, D2H {gb / timed(lambda: sv.d2h(hp)):.1f} GB/s")
    if "QSV_IO_THREADS" not in os.environ:
        subprocess.run([sys.executable, __file__], env=dict(os.environ, QSV_IO_STAGED="0", AB_IO_CHILD="1"))
