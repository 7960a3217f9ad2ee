"""CPU-only check of the HOST side of the fused executor against the oracle: gate merging, DAG sweep packing, pass
scheduling, folded index permutations (PauliX / CNOT / SWAP at pass boundaries), merged thread-uniform diagonal
gates, affine address maps and program encoding.  tests/native/regs_emu.cu emulates the thread / CTA structure of
k_tile_regs and runs the kernel's own per-thread code (csrc/tile_regs_core.cuh) on the very program the GPU would get.
The GPU execution of the same programs is covered by the -m gpu parity tests."""
import ctypes as C
import importlib.util
import os

import numpy as np
import pytest

from oracle import np_oracle as orc
from pennylane_lightning_gpu_b200 import workloads

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    spec = importlib.util.spec_from_file_location("build_emu", os.path.join(HERE, "native", "build_emu.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = C.CDLL(mod.build())
    lib.regs_emu_apply_ops.restype = C.c_int
    lib.regs_emu_apply_ops.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    return lib


def run_emulated(emu, ops, psi0, dtype=np.complex128, rb=4, low=0, dag=True):
    import pennylane_lightning_gpu_b200 as q

    n = int(np.log2(psi0.size))
    rec = q.Ops(ops)
    buf = np.ascontiguousarray(psi0.astype(np.complex128)).view(np.float64).copy()
    stats = (C.c_int64 * 6)()
    rc = emu.regs_emu_apply_ops(rec._h, n, 1 if dtype == np.complex128 else 0, rb, low, int(dag),
                                buf.ctypes.data_as(C.POINTER(C.c_double)), stats)
    assert rc == 0
    names = ("sweeps", "passes", "folded", "merged_diag", "lone", "mma")
    return buf.view(np.complex128), dict(zip(names, list(stats)))


def rand_state(n, seed):
    rng = np.random.default_rng(seed)
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    return v / np.linalg.norm(v)


def dtype_code_ok():
    import pennylane_lightning_gpu_b200._cabi as cabi

    return cabi


def random_mixed_circuit(n, n_gates, seed):
    rng = np.random.default_rng(seed)
    one = ["RX", "RY", "RZ", "Hadamard", "PauliX", "PauliY", "PauliZ", "S", "T", "PhaseShift", "Rot"]
    two = ["CNOT", "CZ", "SWAP", "CY", "CRX", "CRY", "CRZ", "IsingXX", "IsingYY", "IsingZZ", "ControlledPhaseShift",
           "SingleExcitation", "SingleExcitationPlus", "CRot"]
    npar = {"RX": 1, "RY": 1, "RZ": 1, "PhaseShift": 1, "Rot": 3, "CRX": 1, "CRY": 1, "CRZ": 1, "IsingXX": 1, "IsingYY": 1,
            "IsingZZ": 1, "ControlledPhaseShift": 1, "SingleExcitation": 1, "SingleExcitationPlus": 1, "CRot": 3,
            "MultiRZ": 1, "DoubleExcitation": 1}
    ops = []
    for _ in range(n_gates):
        r = rng.random()
        if r < 0.40:
            nm, k = one[rng.integers(len(one))], 1
        elif r < 0.80:
            nm, k = two[rng.integers(len(two))], 2
        elif r < 0.86:
            nm, k = ["Toffoli", "CSWAP", "MultiRZ"][rng.integers(3)], 3
        elif r < 0.93:
            w = [int(x) for x in rng.choice(n, 1 + int(rng.integers(2)), replace=False)]
            ops.append({"name": "QubitUnitary", "wires": w, "params": [],
                        "matrix": workloads.haar_unitary(rng, 1 << len(w))})
            continue
        else:
            nm, k = "DoubleExcitation", 4
        w = [int(x) for x in rng.choice(n, k, replace=False)]
        ops.append({"name": nm, "wires": w, "params": [float(x) for x in rng.uniform(-3, 3, npar.get(nm, 0))],
                    "adjoint": bool(rng.random() < 0.2)})
    return ops


def code_of(dtype):
    return dtype


@pytest.mark.parametrize("n,seed", [(12, 1), (13, 2), (14, 3)])
@pytest.mark.parametrize("rb", [4, 3])
def test_mixed_circuits_match_oracle(emu, n, seed, rb):
    ops = random_mixed_circuit(n, 150, seed)
    psi0 = rand_state(n, 100 + seed)
    want = orc.apply_ops(psi0.copy(), ops)
    got, st = run_emulated(emu, ops, psi0, rb=rb)
    assert np.max(np.abs(got - want)) < 1e-12, st
    assert st["sweeps"] >= 1 and st["passes"] >= st["sweeps"] - st["lone"]


@pytest.mark.parametrize("env", [{}, {"QSV_REGS_FOLD": "0"}, {"QSV_REGS_UDIAG": "0"}, {"QSV_REGS_DAG": "0"},
                                 {"QSV_REGS_FOLD": "0", "QSV_REGS_UDIAG": "0"}, {"QSV_MERGE_1Q": "0"}, {"QSV_REGS_MMA": "0"}])
def test_feature_switches(emu, env, monkeypatch):
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    n = 13
    ops = random_mixed_circuit(n, 120, 7)
    psi0 = rand_state(n, 8)
    want = orc.apply_ops(psi0.copy(), ops)
    got, st = run_emulated(emu, ops, psi0, dag=env.get("QSV_REGS_DAG", "1") == "1")
    assert np.max(np.abs(got - want)) < 1e-12, (env, st)
    if env.get("QSV_REGS_FOLD") == "0":
        assert st["folded"] == 0
    if env.get("QSV_REGS_UDIAG") == "0":
        assert st["merged_diag"] == 0
    if env.get("QSV_REGS_MMA") == "0":
        assert st["mma"] == 0
    elif not env:
        assert st["mma"] > 0


@pytest.mark.parametrize("beam", [2, 4, 16])
@pytest.mark.parametrize("dtype", [np.complex128, np.complex64])
def test_pass_sequence_search(emu, beam, dtype, monkeypatch):
    """csrc/tile_regs.cu: beam search over pass sequences (QSV_REGS_BEAM; the default from 28 qubits up).  The programs it
    builds must give the oracle's state, and must not need more passes than the greedy rule does."""
    passes = {}
    for b in (1, beam):
        monkeypatch.setenv("QSV_REGS_BEAM", str(b))
        total = 0
        for n, seed in ((12, 11), (14, 12), (15, 13)):
            ops = random_mixed_circuit(n, 160, seed) if seed != 12 else workloads.random_gate_circuit(n, 200, 2024)
            psi0 = rand_state(n, 50 + seed)
            want = orc.apply_ops(psi0.copy(), ops)
            got, st = run_emulated(emu, ops, psi0, dtype=dtype)
            assert np.max(np.abs(got - want)) < (1e-12 if dtype == np.complex128 else 3e-5), (b, n, st)
            total += st["passes"]
        ladder, _ = workloads.hardware_efficient_ansatz(13, layers=3, seed=11)
        psi0 = rand_state(13, 99)
        got, st = run_emulated(emu, ladder, psi0, dtype=dtype)
        assert np.max(np.abs(got - orc.apply_ops(psi0.copy(), ladder))) < (1e-12 if dtype == np.complex128 else 3e-5)
        passes[b] = total + st["passes"]
    assert passes[beam] <= passes[1], passes


@pytest.mark.parametrize("tries", [3, 8])
def test_multi_start_sweep_packing(emu, tries, monkeypatch):
    """csrc/tile_kernels.cu: plan_sweeps_regs tries several orders of offering the ready gates to first fit and keeps the
    cheapest plan under the cost model (QSV_REGS_PACK_TRIES; default from 28 qubits up).  Whatever it picks must give the
    oracle's state; on a layered ansatz it must not need more sweeps than program order."""
    sweeps = {}
    for t in (1, tries):
        monkeypatch.setenv("QSV_REGS_PACK_TRIES", str(t))
        for n, seed in ((13, 21), (15, 22)):
            ops = random_mixed_circuit(n, 180, seed)
            psi0 = rand_state(n, 60 + seed)
            got, st = run_emulated(emu, ops, psi0)
            assert np.max(np.abs(got - orc.apply_ops(psi0.copy(), ops))) < 1e-12, (t, n, st)
        ladder, _ = workloads.hardware_efficient_ansatz(16, layers=4, seed=11)
        psi0 = rand_state(16, 98)
        got, st = run_emulated(emu, ladder, psi0)
        assert np.max(np.abs(got - orc.apply_ops(psi0.copy(), ladder))) < 1e-12
        got32, _ = run_emulated(emu, ladder, psi0, dtype=np.complex64)
        assert np.max(np.abs(got32 - got)) < 3e-5
        sweeps[t] = st["sweeps"]
    assert sweeps[tries] <= sweeps[1], sweeps


def test_permutation_only_and_ladders(emu):
    n = 12
    psi0 = rand_state(n, 3)
    # CNOT ladder + X + SWAP chain: everything folds into the load / store address maps, no arithmetic at all
    ops = [{"name": "CNOT", "wires": [i, i + 1], "params": []} for i in range(n - 1)]
    ops += [{"name": "PauliX", "wires": [3], "params": []}, {"name": "SWAP", "wires": [0, 11], "params": []},
            {"name": "SWAP", "wires": [5, 6], "params": []}, {"name": "CNOT", "wires": [11, 0], "params": []}]
    want = orc.apply_ops(psi0.copy(), ops)
    got, st = run_emulated(emu, ops, psi0)
    assert np.max(np.abs(got - want)) == 0.0
    assert st["folded"] == len(ops) and st["sweeps"] == 1 and st["passes"] == 1
    # hardware-efficient ansatz: rotations + ladder per layer
    ops, _ = workloads.hardware_efficient_ansatz(n, layers=3, seed=11)
    want = orc.apply_ops(psi0.copy(), ops)
    got, st = run_emulated(emu, ops, psi0)
    assert np.max(np.abs(got - want)) < 1e-12
    assert st["folded"] > 0


@pytest.mark.parametrize("low", [3, 4, 6])
def test_config2_style_circuit(emu, low):
    n = 14
    ops = workloads.random_gate_circuit(n, 200, 2024)
    psi0 = rand_state(n, 5)
    want = orc.apply_ops(psi0.copy(), ops)
    got, st = run_emulated(emu, ops, psi0, low=low)
    assert np.max(np.abs(got - want)) < 1e-12, st
    got32, _ = run_emulated(emu, ops, psi0, dtype=np.complex64, low=low)
    assert np.max(np.abs(got32 - want)) < 2e-5


@pytest.mark.parametrize("merge2q", ["1", "0"])
def test_two_qubit_blocks_absorb_neighbours(emu, merge2q, monkeypatch):
    """merge_single_qubit_runs: a dense two-qubit gate absorbs the pending single-qubit blocks of its qubits, later
    single-qubit gates on them and later two-qubit gates on the same pair in either wire order."""
    import pennylane_lightning_gpu_b200 as q

    monkeypatch.setenv("QSV_MERGE_2Q", merge2q)
    n = 12
    rng = np.random.default_rng(17)
    ops = []
    for rep in range(40):
        a, b = (int(x) for x in rng.choice(n, 2, replace=False))
        for _ in range(int(rng.integers(0, 3))):
            ops.append({"name": ["RX", "RY", "RZ", "Hadamard", "T"][rng.integers(5)], "wires": [[a, b][rng.integers(2)]],
                        "params": [float(rng.uniform(-3, 3))]})
            if ops[-1]["name"] in ("Hadamard", "T"):
                ops[-1]["params"] = []
        ops.append({"name": "QubitUnitary", "wires": [a, b], "params": [], "matrix": workloads.haar_unitary(rng, 4)})
        if rep % 3 == 0:
            ops.append({"name": "QubitUnitary", "wires": [b, a], "params": [], "matrix": workloads.haar_unitary(rng, 4),
                        "adjoint": True})
        if rep % 4 == 0:
            ops.append({"name": "IsingXX", "wires": [a, b], "params": [0.4]})
        for _ in range(int(rng.integers(0, 3))):
            ops.append({"name": "RY", "wires": [[a, b][rng.integers(2)]], "params": [float(rng.uniform(-3, 3))]})
        if rep % 5 == 0:
            ops.append({"name": "CNOT", "wires": [a, b], "params": []})
    psi0 = rand_state(n, 4)
    want = orc.apply_ops(psi0.copy(), ops)
    got, st = run_emulated(emu, ops, psi0)
    assert np.max(np.abs(got - want)) < 1e-12, st
    plan = q.Ops(ops).plan_sweeps(n)
    assert plan["order_valid"]
    if merge2q == "1":
        monkeypatch.setenv("QSV_MERGE_2Q", "0")
        assert plan["gates_after_merge"] < q.Ops(ops).plan_sweeps(n)["gates_after_merge"]


def test_sweep_fused_with_exchange(emu):
    """k_tile_regs<XCHG = true> (csrc/tile_regs.cu) + dist_apply_ops with QSV_DIST_FUSED_SWAP=1 (csrc/dist.cu), emulated
    for the two ranks of a register sharded on one global bit: a batch of local gates whose last sweep stores out of place
    through xchg_target (this rank's other buffer / the partner's, exchanged bit flipped), or the copy-pass form when that
    sweep cannot carry it.  Must equal: gates on both shards, then global bit <-> local bit exchanged."""
    import ctypes as C

    import pennylane_lightning_gpu_b200 as q

    dp = C.POINTER(C.c_double)
    emu.regs_emu_fused_exchange.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.POINTER(C.c_int)]
    modes = set()
    for n_local, local_bit, low in [(13, 12, 4), (14, 13, 4), (14, 9, 4), (13, 5, 4), (14, 12, 3), (13, 2, 4)]:
        ops = random_mixed_circuit(n_local, 60, 70 + n_local + local_bit)
        rec = q.Ops(ops)
        shards = [rand_state(n_local, 200 + r) / np.sqrt(2) for r in range(2)]
        want = [orc.apply_ops(s.copy(), ops) for s in shards]
        # exchange of the rank bit with local bit `local_bit`: amplitude (rank = x, bit = y, rest) -> (rank = y, bit = x, rest)
        idx = np.arange(1 << n_local)
        bit = (idx >> local_bit) & 1
        expect = [np.empty_like(want[0]), np.empty_like(want[0])]
        for r in range(2):
            for y in range(2):
                src = idx[bit == y]                       # on rank r with the bit = y ...
                dst = src ^ ((y ^ r) << local_bit)        # ... goes to rank y with the bit = r
                expect[y][dst] = want[r][src]
        out = [np.zeros(2 << n_local), np.zeros(2 << n_local)]
        carried = C.c_int(-1)
        ins = [np.ascontiguousarray(s).view(np.float64) for s in shards]
        rc = emu.regs_emu_fused_exchange(rec._h, n_local, low, local_bit, ins[0].ctypes.data_as(dp), ins[1].ctypes.data_as(dp),
                                         out[0].ctypes.data_as(dp), out[1].ctypes.data_as(dp), C.byref(carried))
        assert rc == 0
        for r in range(2):
            got = out[r].view(np.complex128)
            assert np.max(np.abs(got - expect[r])) < 1e-12, (n_local, local_bit, r, carried.value)
        modes.add(carried.value)
    # the last sweep carries the exchange wherever the exchanged bit sits (outside the tile, thread bit, register bit); the
    # copy pass is left for batches that end in a lone unfused gate
    assert 1 in modes and modes <= {0, 1}


def test_split_exchange_push_then_pull(emu):
    """QSV_DIST_SPLIT_XCHG (csrc/dist.cu, csrc/tile_regs.cu): an exchange split between the last sweep of the batch before it
    (xchg_target with a stash bit: push a quarter of the shard, park a quarter) and the first sweep of the batch after it
    (xchg_source: fetch what the partner parked), emulated for the two ranks of a register sharded on one global bit with the
    kernel's own per-thread code and routing rules -- carried by sweeps and in the copy-pass forms.  Must equal: gates of
    batch 1 on both shards, global bit <-> local bit exchanged, gates of batch 2."""
    import ctypes as C

    import pennylane_lightning_gpu_b200 as q

    dp = C.POINTER(C.c_double)
    emu.regs_emu_split_exchange.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, dp, dp, dp, dp, C.POINTER(C.c_int)]
    modes = set()
    cases = [(13, 12, 4, 60, 50), (14, 13, 4, 40, 30), (14, 9, 4, 50, 1), (13, 4, 5, 30, 40), (14, 12, 4, 1, 25), (13, 7, 4, 0, 20),
             (13, 11, 4, 20, 0)]
    for n_local, local_bit, stash, n1, n2 in cases:
        ops1 = random_mixed_circuit(n_local, n1, 170 + n_local + local_bit) if n1 else []
        ops2 = random_mixed_circuit(n_local, n2, 270 + n_local + local_bit) if n2 else []
        shards = [rand_state(n_local, 300 + r) / np.sqrt(2) for r in range(2)]
        mid = [orc.apply_ops(s.copy(), ops1) for s in shards]
        idx = np.arange(1 << n_local)
        bit = (idx >> local_bit) & 1
        swapped = [np.empty_like(mid[0]), np.empty_like(mid[0])]
        for r in range(2):
            for y in range(2):
                src = idx[bit == y]                       # on rank r with the bit = y ...
                dst = src ^ ((y ^ r) << local_bit)        # ... goes to rank y with the bit = r
                swapped[y][dst] = mid[r][src]
        expect = [orc.apply_ops(s, ops2) for s in swapped]
        out = [np.zeros(2 << n_local), np.zeros(2 << n_local)]
        mode = C.c_int(-1)
        ins = [np.ascontiguousarray(s).view(np.float64) for s in shards]
        rec1, rec2 = q.Ops(ops1), q.Ops(ops2)  # kept alive across the call
        rc = emu.regs_emu_split_exchange(rec1._h, rec2._h, n_local, local_bit, stash, ins[0].ctypes.data_as(dp),
                                         ins[1].ctypes.data_as(dp), out[0].ctypes.data_as(dp), out[1].ctypes.data_as(dp),
                                         C.byref(mode))
        assert rc == 0
        for r in range(2):
            got = out[r].view(np.complex128)
            assert np.max(np.abs(got - expect[r])) < 1e-12, (n_local, local_bit, stash, n1, n2, r, mode.value)
        modes.add(mode.value)
    assert 3 in modes and (0 in modes or 1 in modes or 2 in modes)  # both halves carried by sweeps, and copy-pass forms
