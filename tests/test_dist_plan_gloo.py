"""N > 1 host logic on CPU: the exchange schedule the library plans for a sharded register
(qsv_dist_plan, host-only) is executed here on NumPy shards by two gloo ranks -- torch.distributed
send/recv standing in for NCCL -- and must reproduce the oracle's full state.  Local gate arithmetic is
the oracle's (this test checks the N > 1 plumbing: qubit map, swap semantics, partner choice)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _circuit(n, seed, n_gates=40):
    from pennylane_lightning_gpu_b200 import workloads

    rng = np.random.default_rng(seed)
    ops = workloads.random_gate_circuit(n, n_gates, seed)
    # add gates that must NOT trigger exchanges on global wires: controls and diagonal gates
    ops.insert(5, {"name": "CZ", "wires": [0, n - 1], "params": []})
    ops.insert(9, {"name": "RZ", "wires": [0], "params": [0.3]})
    ops.insert(12, {"name": "CNOT", "wires": [0, 3], "params": []})
    ops.insert(20, {"name": "Toffoli", "wires": [1, 0, 2], "params": []})
    return ops


def _swap_shard(shard, n_local, rank, gphys, l, send_recv):
    """Reference semantics of qsv_dist_swap_bits on a NumPy shard."""
    gb = gphys - n_local
    peer = rank ^ (1 << gb)
    mybit = (rank >> gb) & 1
    idx = np.arange(shard.size)
    sel = ((idx >> l) & 1) == (mybit ^ 1)
    recv = send_recv(shard[sel].copy(), peer)
    out = shard.copy()
    out[sel] = recv
    return out


def _gather_logical(shards, phys_of, n_total, n_local):
    """Full logical state from all shards under the logical->physical map."""
    full_phys = np.concatenate(shards)  # physical index = rank << n_local | local
    idx = np.arange(1 << n_total)
    phys_idx = np.zeros_like(idx)
    for b in range(n_total):
        phys_idx |= ((idx >> b) & 1) << phys_of[b]
    return full_phys[phys_idx]


def _worker(rank, world, port, n_total, seed, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import np_oracle as orc
    from pennylane_lightning_gpu_b200 import Ops
    from pennylane_lightning_gpu_b200.distributed import plan

    g = int(np.log2(world))
    n_local = n_total - g
    ops = _circuit(n_total, seed)
    steps, final_map = plan(Ops(ops), n_total, n_local)
    shard = np.zeros(1 << n_local, dtype=np.complex128)
    if rank == 0:
        shard[0] = 1.0
    phys_of = list(range(n_total))
    log_of = list(range(n_total))

    def send_recv(buf, peer):
        t_out = torch.from_numpy(np.ascontiguousarray(buf).view(np.float64))
        t_in = torch.empty_like(t_out)
        reqs = dist.batch_isend_irecv([dist.P2POp(dist.isend, t_out, peer), dist.P2POp(dist.irecv, t_in, peer)])
        for r in reqs:
            r.wait()
        return t_in.numpy().view(np.complex128)

    n_swaps = 0
    for st in steps:
        if st[0] == "swap":
            _, gp, l = st
            shard = _swap_shard(shard, n_local, rank, gp, l, send_recv)
            a, b = log_of[gp], log_of[l]
            log_of[gp], log_of[l] = b, a
            phys_of[a], phys_of[b] = l, gp
            n_swaps += 1
        else:
            # apply the op on the logical full state (gathered), then re-scatter: checks the map only
            gathered = [torch.empty(shard.size * 2, dtype=torch.float64) for _ in range(world)]
            dist.all_gather(gathered, torch.from_numpy(shard.view(np.float64).copy()))
            shards = [t.numpy().view(np.complex128) for t in gathered]
            full = _gather_logical(shards, phys_of, n_total, n_local)
            op = ops[st[1]]
            # dense targets must be local at this point
            full = orc.apply_op(full, op["name"], op["wires"], op.get("params", ()), op.get("adjoint", False),
                                op.get("matrix"))
            idx = np.arange(1 << n_total)
            phys_idx = np.zeros_like(idx)
            for bb in range(n_total):
                phys_idx |= ((idx >> bb) & 1) << phys_of[bb]
            full_phys = np.empty_like(full)
            full_phys[phys_idx] = full
            shard = full_phys[rank << n_local:(rank + 1) << n_local].copy()
    assert phys_of == final_map
    gathered = [torch.empty(shard.size * 2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(shard.view(np.float64).copy()))
    full = _gather_logical([t.numpy().view(np.complex128) for t in gathered], phys_of, n_total, n_local)
    want = orc.apply_ops(orc.basis_state(n_total), ops)
    ok = bool(np.allclose(full, want, atol=1e-12))
    if rank == 0:
        ret["ok"] = ok
        ret["n_swaps"] = n_swaps
        ret["n_gates"] = sum(1 for s in steps if s[0] == "gate")
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_swap_plan_reproduces_full_state_with_gloo(world):
    from pennylane_lightning_gpu_b200 import _build

    _build.build_lib()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(world, port, 7, 11, ret), nprocs=world, join=True)
    assert ret["ok"]
    assert ret["n_gates"] == 44
    assert 0 < ret["n_swaps"] < 20


def test_plan_properties():
    """No exchange for controls / diagonal gates on global wires; dense targets are local when applied."""
    from pennylane_lightning_gpu_b200 import Ops, _build
    from pennylane_lightning_gpu_b200.distributed import plan

    _build.build_lib()
    n_total, n_local = 10, 7
    free = [{"name": "CZ", "wires": [0, 9]}, {"name": "RZ", "wires": [1], "params": [0.2]},
            {"name": "CNOT", "wires": [2, 5]}, {"name": "IsingZZ", "wires": [0, 1], "params": [0.1]},
            {"name": "PhaseShift", "wires": [0], "params": [0.4]}, {"name": "Toffoli", "wires": [0, 1, 9]},
            {"name": "MultiRZ", "wires": [0, 1, 2, 3], "params": [0.5]}, {"name": "CRZ", "wires": [5, 0], "params": [0.3]}]
    steps, fm = plan(Ops(free), n_total, n_local)
    assert all(s[0] == "gate" for s in steps) and fm == list(range(n_total))
    # a dense target on global wire 0 (bit 9) costs exactly one exchange, and stays local afterwards
    ops = [{"name": "RX", "wires": [0], "params": [0.1]}, {"name": "RY", "wires": [0], "params": [0.2]},
           {"name": "Hadamard", "wires": [0]}]
    steps, fm = plan(Ops(ops), n_total, n_local)
    assert [s[0] for s in steps] == ["swap", "gate", "gate", "gate"]
    assert steps[0][1] == 9 and fm[9] == steps[0][2] and steps[0][2] >= n_local - 8
    # Belady: the evicted qubit is the one needed latest
    ops = [{"name": "RX", "wires": [0], "params": [0.1]}] + \
          [{"name": "RX", "wires": [w], "params": [0.1]} for w in (3, 4, 5, 6, 7, 8)]  # bits 6..1 used soon
    steps, fm = plan(Ops(ops), n_total, n_local)
    assert steps[0][0] == "swap" and steps[0][2] == 0  # bit 0 (wire 9) is never used again
    assert sum(1 for s in steps if s[0] == "swap") == 1
