"""Multi-GPU parity check, run under torchrun (one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_check.py

Same pattern as the reference's src/tests/mpi/Test_StateVectorCudaMPI_Param.cpp:59-125 (apply on the
sharded register with wires chosen to hit global and local bits, compare shard-wise with the
single-device result), with the oracle as the single-device result."""
import math
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import np_oracle as orc  # noqa: E402
import pennylane_lightning_gpu_b200 as q  # noqa: E402
from pennylane_lightning_gpu_b200 import workloads  # noqa: E402
from pennylane_lightning_gpu_b200.distributed import DistributedStateVector  # noqa: E402


def circuit(n, seed, n_gates=60):
    ops = workloads.random_gate_circuit(n, n_gates, seed)
    extra = [{"name": "CZ", "wires": [0, n - 1], "params": []}, {"name": "RZ", "wires": [0], "params": [0.3]},
             {"name": "CNOT", "wires": [0, 3], "params": []}, {"name": "Toffoli", "wires": [1, 0, 2], "params": []},
             {"name": "IsingXX", "wires": [0, 1], "params": [0.7]}, {"name": "MultiRZ", "wires": [0, 2, 5], "params": [0.9]},
             {"name": "CRY", "wires": [4, 0], "params": [1.1]}, {"name": "SWAP", "wires": [0, n - 2], "params": []},
             {"name": "PhaseShift", "wires": [1], "params": [0.4]}, {"name": "DoubleExcitation", "wires": [0, 1, 2, 3], "params": [0.5]},
             {"name": "ControlledPhaseShift", "wires": [0, 1], "params": [0.2]}, {"name": "Hadamard", "wires": [0], "params": []}]
    for i, e in enumerate(extra):
        ops.insert(3 + 4 * i, e)
    return ops


def main():
    rank = int(os.environ["RANK"])
    local_rank = int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    g = int(math.log2(world))
    failures = []
    for dtype, tol in ((np.complex128, 1e-10), (np.complex64, 2e-5)):
        for n_total in (g + 6, g + 13):
            n_local = n_total - g
            ops = circuit(n_total, seed=n_total)
            want = orc.apply_ops(orc.basis_state(n_total), ops)
            for fuse in (False, True):
                for chunk in (0, 1 << 12):
                    sv = DistributedStateVector(n_total, dtype, device=local_rank, chunk_bytes=chunk)
                    sv.apply_ops(q.Ops(ops), fuse=fuse)
                    n_swaps, nbytes, ms = sv.swap_stats()
                    # measurements in the permuted layout
                    words = ["X", "Z", "XY", "ZZ", "YXZ"]
                    wires = [[0], [0], [0, n_total - 1], [1, 2], [1, 0, 3]]
                    coeffs = [0.5, -1.0, 0.3, 0.8, 1.2]
                    ev = sv.expval_pauli_words(words, wires, coeffs)
                    ev_want = orc.expval_pauli_words(want, words, wires, coeffs)
                    nrm = sv.norm2()
                    sv.canonicalize()
                    assert sv.qubit_map() == list(range(n_total))
                    shard = torch.from_numpy(sv.local_state().astype(np.complex128).view(np.float64)).cuda()
                    parts = [torch.empty_like(shard) for _ in range(world)]
                    dist.all_gather(parts, shard)
                    full = np.concatenate([p.cpu().numpy().view(np.complex128) for p in parts])
                    err = float(np.max(np.abs(full - want)))
                    tag = f"dtype={np.dtype(dtype).name} n={n_total} fuse={fuse} chunk={chunk}"
                    if err > tol * 10 or abs(ev - ev_want) > tol * 100 or abs(nrm - 1) > tol * 100:
                        failures.append(f"{tag}: state err {err:.2e}, expval {ev} vs {ev_want}, norm {nrm}")
                    if rank == 0:
                        print(f"[dist_check] {tag}: err={err:.2e} expval_err={abs(ev - ev_want):.2e} swaps={n_swaps}", flush=True)
                    sv.close()
    ok = torch.tensor([0 if failures else 1], device="cuda")
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("DIST_CHECK", "PASS" if int(ok) == 1 else "FAIL", failures, flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(ok) == 1 else 1)


if __name__ == "__main__":
    main()
