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


@pytest.mark.parametrize("world,dag", [(2, 1), (2, 0), (4, 1)])
def test_swap_plan_reproduces_full_state_with_gloo(world, dag, monkeypatch):
    from pennylane_lightning_gpu_b200 import _build

    _build.build_lib()
    monkeypatch.setenv("QSV_DIST_DAG", str(dag))  # inherited by the spawned ranks
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() + 7 * world + dag) % 2000
    mp.spawn(_worker, args=(world, port, 7, 11, ret), nprocs=world, join=True)
    assert ret["ok"]
    assert ret["n_gates"] == 44
    assert 0 < ret["n_swaps"] < 20


def test_plan_properties(monkeypatch):
    """No exchange for controls / diagonal gates on global wires; dense targets are local when applied."""
    from pennylane_lightning_gpu_b200 import Ops, _build
    from pennylane_lightning_gpu_b200.distributed import plan

    _build.build_lib()
    n_total, n_local = 10, 7
    free = [{"name": "CZ", "wires": [0, 9]}, {"name": "RZ", "wires": [1], "params": [0.2]},
            {"name": "CNOT", "wires": [2, 5]}, {"name": "IsingZZ", "wires": [0, 1], "params": [0.1]},
            {"name": "PhaseShift", "wires": [0], "params": [0.4]}, {"name": "Toffoli", "wires": [0, 1, 9]},
            {"name": "MultiRZ", "wires": [0, 1, 2, 3], "params": [0.5]}, {"name": "CRZ", "wires": [5, 0], "params": [0.3]}]
    rx_chain = [{"name": "RX", "wires": [0], "params": [0.1]}, {"name": "RY", "wires": [0], "params": [0.2]},
                {"name": "Hadamard", "wires": [0]}]
    belady = [{"name": "RX", "wires": [0], "params": [0.1]}] + \
             [{"name": "RX", "wires": [w], "params": [0.1]} for w in (3, 4, 5, 6, 7, 8)]  # bits 6..1 used soon
    for dag in ("0", "1"):
        monkeypatch.setenv("QSV_DIST_DAG", dag)
        steps, fm = plan(Ops(free), n_total, n_local)
        assert all(s[0] == "gate" for s in steps) and fm == list(range(n_total))
        assert sorted(s[1] for s in steps) == list(range(len(free)))
        # a dense target on global wire 0 (bit 9) costs exactly one exchange, and stays local afterwards
        steps, fm = plan(Ops(rx_chain), n_total, n_local)
        assert steps == [("swap", 9, steps[0][2]), ("gate", 0), ("gate", 1), ("gate", 2)]
        assert fm[9] == steps[0][2] and steps[0][2] >= n_local - 8
        steps, fm = plan(Ops(belady), n_total, n_local)
        assert sum(1 for s in steps if s[0] == "swap") == 1
        if dag == "0":
            # program order, Belady: the evicted qubit is the one needed latest -- bit 0 (wire 9) is never used again
            assert steps[0][0] == "swap" and steps[0][2] == 0
        else:
            # dependency order: everything that is local runs first, the exchange comes last
            assert [s[0] for s in steps] == ["gate"] * 6 + ["swap", "gate"] and steps[-1] == ("gate", 0)


def _n_swaps(steps):
    return sum(1 for s in steps if s[0] == "swap")


def _check_plan(ops, steps, final_map, n_total, n_local, initial_map=None):
    """Every gate exactly once, dense targets local when it runs, non-commuting pairs in program order."""
    from pennylane_lightning_gpu_b200 import workloads

    phys = list(initial_map) if initial_map is not None else list(range(n_total))
    log = [0] * n_total
    for q, p in enumerate(phys):
        log[p] = q
    pos = {}
    for k, st in enumerate(steps):
        if st[0] == "swap":
            _, gp, l = st
            assert gp >= n_local > l >= 0
            a, b = log[gp], log[l]
            log[gp], log[l] = b, a
            phys[a], phys[b] = l, gp
        else:
            i = st[1]
            assert i not in pos
            pos[i] = k
            dense, _ = workloads.gate_bit_masks(ops[i], n_total)
            for b in range(n_total):
                if dense >> b & 1:
                    assert phys[b] < n_local, (i, ops[i], b)
    assert sorted(pos) == list(range(len(ops))) and phys == final_map
    masks = [workloads.gate_bit_masks(op, n_total) for op in ops]
    for i in range(len(ops)):
        for j in range(i):
            conflict = (masks[i][0] & (masks[j][0] | masks[j][1])) or (masks[i][1] & masks[j][0])
            if conflict:
                assert pos[j] < pos[i], (j, i, ops[j], ops[i])


@pytest.mark.parametrize("n_total,n_local", [(31, 30), (32, 30), (33, 30), (36, 33), (12, 6)])
def test_dependency_order_needs_fewer_exchanges(n_total, n_local, monkeypatch):
    """The bench circuits (config 2 / config 5 sizes): the dependency-ordered schedule is valid and never needs more
    exchanges than the program-order schedule; at 4 and 8 GPUs it needs fewer."""
    from pennylane_lightning_gpu_b200 import Ops, _build, workloads
    from pennylane_lightning_gpu_b200.distributed import plan

    _build.build_lib()
    ops = workloads.random_gate_circuit(n_total, 200, 2024)
    rec = Ops(ops)
    monkeypatch.setenv("QSV_DIST_DAG", "0")
    s0, f0 = plan(rec, n_total, n_local)
    monkeypatch.setenv("QSV_DIST_DAG", "1")
    s1, f1 = plan(rec, n_total, n_local)
    _check_plan(ops, s0, f0, n_total, n_local)
    _check_plan(ops, s1, f1, n_total, n_local)
    assert _n_swaps(s1) <= _n_swaps(s0)
    if n_total - n_local >= 2:
        assert _n_swaps(s1) < _n_swaps(s0)


def test_cyclic_use_of_more_qubits_than_fit():
    """Four qubits used round-robin with three local slots: program order exchanges every round, the dependency
    order (all RX gates commute across wires) exchanges once."""
    from pennylane_lightning_gpu_b200 import Ops, _build
    from pennylane_lightning_gpu_b200.distributed import plan

    _build.build_lib()
    ops = [{"name": "RX", "wires": [w], "params": [0.1 * (r + 1)]} for r in range(4) for w in range(4)]
    steps, _ = plan(Ops(ops), 4, 3)
    assert _n_swaps(steps) == 1


@pytest.mark.parametrize("n_total,n_local,expect", [(31, 30, 1), (32, 30, 3), (36, 33, 3)])
def test_repeated_application_keeps_the_map(n_total, n_local, expect):
    """bench.py applies the same circuit again and again; the qubit map persists between the calls.  Planned exchanges
    per step in the steady state: 1 / 2-3 / 3 at 2 / 4 / 8 GPUs (program order: 2 / 6 / 6)."""
    from pennylane_lightning_gpu_b200 import Ops, _build, workloads
    from pennylane_lightning_gpu_b200.distributed import plan

    _build.build_lib()
    ops = workloads.random_gate_circuit(n_total, 200, 2024)
    rec = Ops(ops)
    m = None
    for _ in range(6):
        steps, m_next = plan(rec, n_total, n_local, m)
        _check_plan(ops, steps, m_next, n_total, n_local, m)
        assert _n_swaps(steps) <= expect
        m = m_next
    with pytest.raises(Exception):
        plan(rec, n_total, n_local, [0] * n_total)  # not a permutation


def test_reordered_plans_are_legal_for_every_gate_family():
    """Random small circuits over one-, two-, three- and four-wire gates (two-level gates like SWAP and the excitations,
    controlled rotations, diagonal families): the gates applied in the planned order give the state of the program order,
    every gate runs exactly once, and the exchanges leave the qubit map the plan reports -- from the identity map and
    from the map the previous application left behind, down to two local qubits."""
    from oracle import np_oracle as orc
    from pennylane_lightning_gpu_b200 import Ops, _build
    from pennylane_lightning_gpu_b200.distributed import plan

    _build.build_lib()
    rng = np.random.default_rng(1)
    names1 = ["RX", "RY", "RZ", "Hadamard", "PauliX", "S", "PhaseShift"]
    names2 = ["CNOT", "CZ", "SWAP", "CRX", "CRZ", "IsingXX", "IsingZZ", "SingleExcitation", "ControlledPhaseShift"]
    with_param = {"RX", "RY", "RZ", "PhaseShift", "CRX", "CRZ", "IsingXX", "IsingZZ", "SingleExcitation",
                  "ControlledPhaseShift", "DoubleExcitation", "MultiRZ"}
    for trial in range(60):
        n_total = int(rng.integers(4, 9))
        n_local = int(rng.integers(2, n_total))
        ops = []
        for _ in range(int(rng.integers(5, 50))):
            r = rng.random()
            if r < 0.4:
                nm, k = names1[rng.integers(len(names1))], 1
            elif r < 0.85:
                nm, k = names2[rng.integers(len(names2))], 2
            elif r < 0.93 and n_local >= 3:
                nm, k = "Toffoli", 3
            elif n_local >= 4:
                nm, k = "DoubleExcitation", 4
            else:
                nm, k = "MultiRZ", 3
            ops.append({"name": nm, "wires": [int(x) for x in rng.choice(n_total, k, replace=False)],
                        "params": [float(rng.uniform(-3, 3))] if nm in with_param else []})
        psi = rng.normal(size=1 << n_total) + 1j * rng.normal(size=1 << n_total)
        psi /= np.linalg.norm(psi)
        want = orc.apply_ops(psi.copy(), ops)
        rec = Ops(ops)
        m = None
        for _ in range(2):
            steps, m_next = plan(rec, n_total, n_local, m)
            phys = list(m) if m is not None else list(range(n_total))
            log = [0] * n_total
            for qb, p in enumerate(phys):
                log[p] = qb
            got, seen = psi.copy(), []
            for st in steps:
                if st[0] == "swap":
                    _, gp, l = st
                    assert gp >= n_local > l >= 0
                    a, b = log[gp], log[l]
                    log[gp], log[l] = b, a
                    phys[a], phys[b] = l, gp
                else:
                    o = ops[st[1]]
                    seen.append(st[1])
                    got = orc.apply_op(got, o["name"], o["wires"], o.get("params", ()))
            assert sorted(seen) == list(range(len(ops))) and phys == m_next
            assert np.max(np.abs(got - want)) < 1e-12, (trial, n_total, n_local)
            m = m_next


@pytest.mark.parametrize("n_total,n_local", [(8, 7), (9, 7), (10, 7), (9, 4), (7, 2), (15, 13), (16, 13)])
@pytest.mark.parametrize("dag", ["1", "0"])
def test_sharded_executor_emulated_on_host_shards(n_total, n_local, dag, monkeypatch):
    """csrc/dist.cu: dist_apply_ops replayed on host shards for all ranks (tests/native/regs_emu.cu: dist_emu_apply_ops) with
    the library's own lowering, exchange schedule and per-rank gate localisation -- controls and diagonal gates on global
    qubits resolved against the rank's bits, two-level gates, 3- and 4-wire gates -- against the oracle's full state, for one
    application and for three in a row over the persisting qubit map."""
    import ctypes as C
    import importlib.util

    from oracle import np_oracle as orc
    from pennylane_lightning_gpu_b200 import Ops, workloads

    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(ROOT, "tests", "native", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    emu = C.CDLL(mod.build())
    emu.dist_emu_apply_ops.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    monkeypatch.setenv("QSV_DIST_DAG", dag)
    # from 13 local qubits the schedule is the one priced with the sweep cost model (csrc/dist.cu: plan_dist_steps_priced;
    # the default threshold of 26 local qubits is lowered here so that the host emulation can afford it)
    monkeypatch.setenv("QSV_DIST_PLAN_MIN_LOCAL", "13")
    rng = np.random.default_rng(5 + n_total)
    ops = workloads.random_gate_circuit(n_total, 50 if n_local < 13 else 120, 40 + n_total)
    extra = [{"name": "CZ", "wires": [0, n_total - 1], "params": []}, {"name": "RZ", "wires": [0], "params": [0.3]},
             {"name": "CNOT", "wires": [0, 3], "params": []}, {"name": "CRY", "wires": [4, 0], "params": [1.1]},
             {"name": "IsingXX", "wires": [0, 1], "params": [0.7]}, {"name": "SWAP", "wires": [0, n_total - 2], "params": []},
             {"name": "ControlledPhaseShift", "wires": [0, 1], "params": [0.2]}, {"name": "PhaseShift", "wires": [1], "params": [0.4]},
             {"name": "SingleExcitation", "wires": [1, 5], "params": [0.6]}, {"name": "Hadamard", "wires": [0], "params": []}]
    if n_local >= 3:
        extra += [{"name": "Toffoli", "wires": [1, 0, 2], "params": []}, {"name": "MultiRZ", "wires": [0, 2, 5], "params": [0.9]}]
    if n_local >= 4:
        extra += [{"name": "DoubleExcitation", "wires": [0, 1, 2, 3], "params": [0.5]}]
    for i, e in enumerate(extra):
        ops.insert(3 + 3 * i, e)
    rec = Ops(ops)
    psi = rng.normal(size=1 << n_total) + 1j * rng.normal(size=1 << n_total)
    psi /= np.linalg.norm(psi)
    want = psi.copy()
    for reps in (1, 3):
        buf = np.ascontiguousarray(psi).view(np.float64).copy()
        n_x = C.c_int(-1)
        rc = emu.dist_emu_apply_ops(rec._h, n_total, n_local, reps, buf.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n_x))
        assert rc == 0
        want = psi.copy()
        for _ in range(reps):
            want = orc.apply_ops(want, ops)
        assert np.max(np.abs(buf.view(np.complex128) - want)) < 1e-12, (reps, n_x.value)
        assert n_x.value >= 1
