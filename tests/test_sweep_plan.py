"""CPU-only checks of the host-side sweep planner of the fused executor (csrc/tile_kernels.cu: plan_sweeps_regs):
packing over the dependency DAG must execute every gate exactly once, keep the order of every pair of gates that
does not commute structurally, and need fewer HBM sweeps than program-order packing."""
import numpy as np
import pytest

from pennylane_lightning_gpu_b200 import _build, workloads


@pytest.fixture(scope="module")
def q():
    _build.build_lib()
    import pennylane_lightning_gpu_b200 as q

    return q


def test_config2_circuit_dag_packing(q):
    ops = workloads.random_gate_circuit(30, 200, 2024)
    rec = q.Ops(ops)
    in_order = rec.plan_sweeps(30, dag=False)
    dag = rec.plan_sweeps(30, dag=True)
    assert in_order["order_valid"] and dag["order_valid"]
    assert in_order["gates_after_merge"] == dag["gates_after_merge"]
    assert dag["sweeps"] < in_order["sweeps"]
    assert dag["sweeps"] <= 9  # 8 with the default tile geometry (4 low bits + 8 arbitrary high bits)


@pytest.mark.parametrize("n,seed", [(12, 0), (14, 1), (20, 2), (33, 3)])
def test_random_circuits_keep_dependencies(q, n, seed):
    rng = np.random.default_rng(seed)
    names1 = ["RX", "RY", "RZ", "Hadamard", "PauliX", "S", "T", "PhaseShift"]
    names2 = ["CNOT", "CZ", "SWAP", "CRX", "CRZ", "IsingXX", "IsingZZ", "ControlledPhaseShift", "SingleExcitation"]
    ops = []
    for _ in range(400):
        r = rng.random()
        if r < 0.5:
            nm = names1[rng.integers(len(names1))]
            w = [int(rng.integers(n))]
        elif r < 0.9:
            nm = names2[rng.integers(len(names2))]
            w = [int(x) for x in rng.choice(n, 2, replace=False)]
        elif r < 0.95:
            nm, w = "Toffoli", [int(x) for x in rng.choice(n, 3, replace=False)]
        else:
            nm, w = "MultiRZ", [int(x) for x in rng.choice(n, 3, replace=False)]
        par = [float(rng.uniform(-3, 3))] if nm[0] in "RCIMS" and nm not in ("CNOT", "CZ", "SWAP", "S") or nm == "PhaseShift" else []
        ops.append({"name": nm, "wires": w, "params": par})
    rec = q.Ops(ops)
    for low in (0, 3, 6):
        for dag in (False, True):
            plan = rec.plan_sweeps(n, dag=dag, low_bits=low)
            assert plan["order_valid"], (n, seed, low, dag, plan)
            assert plan["max_gates_per_sweep"] <= 48


def test_plan_rejects_unknown_gate(q):
    rec = q.Ops([{"name": "NoSuchGate", "wires": [0], "params": []}])
    with pytest.raises(q.QsvError):
        rec.plan_sweeps(20)


def test_plan_cache_is_keyed_by_structure(q, monkeypatch):
    """csrc/tile_kernels.cu: plan_sweeps_cached.  The same gate structure with other angles reuses the cached plan, another
    structure does not, and every plan that comes back is valid for the circuit it was asked for."""
    ops = workloads.random_gate_circuit(30, 200, 2024)
    first = q.Ops(ops).plan_sweeps(30, dag=True)
    again = q.Ops(ops).plan_sweeps(30, dag=True)
    scaled = [dict(o) for o in ops]
    for o in scaled:
        if o.get("params"):
            o["params"] = [0.5 * p + 0.1 for p in o["params"]]
    same_structure = q.Ops(scaled).plan_sweeps(30, dag=True)
    other = q.Ops(workloads.random_gate_circuit(30, 200, 7)).plan_sweeps(30, dag=True)
    for p in (first, again, same_structure, other):
        assert p["order_valid"]
    assert first["sweeps"] == again["sweeps"] == same_structure["sweeps"]
    assert first["gates_after_merge"] == same_structure["gates_after_merge"]
    monkeypatch.setenv("QSV_REGS_PLAN_CACHE", "0")
    uncached = q.Ops(ops).plan_sweeps(30, dag=True)
    assert uncached["order_valid"] and uncached["sweeps"] == first["sweeps"]


def test_plan_work_counts_arithmetic(q):
    """qsv_ops_plan_work: multiply-adds per amplitude of the fused programs.  Two uncontrolled 2x2 gates on different
    qubits cost 8 each (or one tensor-core block of 16 for both); a CNOT is folded into the address maps and costs nothing;
    the config-2 circuit lands between its diagonal-only and all-dense bounds."""
    w = q.Ops([{"name": "RX", "wires": [3], "params": [0.3]}, {"name": "Hadamard", "wires": [7], "params": []},
               {"name": "CNOT", "wires": [1, 2], "params": []}]).plan_work(20)
    assert w["sweeps"] == 1 and w["passes"] == 1 and 8.0 <= w["fma_per_amplitude"] <= 16.0
    ops = workloads.random_gate_circuit(30, 200, 2024)
    w = q.Ops(ops).plan_work(30)
    assert w["sweeps"] <= 9 and w["passes"] >= w["sweeps"]
    assert 4.0 * 60 < w["fma_per_amplitude"] < 16.0 * 200
    w32 = q.Ops(ops).plan_work(30, np.complex64)
    assert w32["sweeps"] == w["sweeps"]
