#!/usr/bin/env python
"""Build tests/golden/reference_kats.json from the reference's OWN tests.

Run in the build container only (needs /root/reference, which does not exist on the
GPU box):   python tests/golden/make_golden.py

Three sources, all cited per entry:
 (1) machine-extracted with ``ast`` from the reference's Python tests
     (tests/test_apply.py class-level ``test_data_*`` tables): in/out state pairs;
 (2) machine-extracted with a regex from the reference's C++ tests (numeric
     initialiser lists at cited line ranges);
 (3) scalar known answers (Jacobians, expvals) transcribed from CHECK(...) lines of
     the C++ tests, each with its file:line.
Nothing here is computed by our oracle: the file pins the oracle, not vice versa.
"""
import ast
import json
import math
import os
import re

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")


def cplx(z):
    z = complex(z)
    return [z.real, z.imag]


def cvec(v):
    return [cplx(z) for z in np.asarray(v, dtype=np.complex128).reshape(-1)]


# ---------------------------------------------------------------- (1) Python tables
class _Qml:
    def __getattr__(self, name):
        return name


def extract_test_apply():
    path = os.path.join(REF, "tests/test_apply.py")
    src = open(path).read()
    tree = ast.parse(src)
    wanted = {
        "test_data_no_parameters": ("in_out", 1),
        "test_data_two_wires_no_parameters": ("in_out", 2),
        "test_data_three_wires_no_parameters": ("in_out", 3),
        "test_data_single_wire_with_parameters": ("in_out_par", 1),
        "test_data_two_wires_with_parameters": ("in_out_par", 2),
    }
    env = {"qml": _Qml(), "math": math, "np": np}
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and \
                isinstance(node.targets[0], ast.Name) and node.targets[0].id in wanted:
            name = node.targets[0].id
            kind, nw = wanted[name]
            rows = eval(compile(ast.Expression(node.value), path, "eval"), env)
            for i, row in enumerate(rows):
                if kind == "in_out":
                    op, inp, exp = row
                    par = []
                else:
                    op, inp, exp, par = row
                out.append({
                    "cite": f"tests/test_apply.py:{node.lineno} ({name}[{i}])",
                    "gate": op, "wires": list(range(nw)), "params": [float(p) for p in par],
                    "input": cvec(inp), "expected": cvec(exp), "atol": 1e-4,
                })
    return out


# ---------------------------------------------------------------- (2) C++ initialiser lists
_PAIR = re.compile(r"\{\s*([-+0-9.eE]+)\s*,\s*([-+0-9.eE]+)\s*\}")


def cpp_pairs(relpath, first, last):
    lines = open(os.path.join(REF, relpath)).read().splitlines()[first - 1:last]
    return [[float(a), float(b)] for a, b in _PAIR.findall("\n".join(lines))]


def extract_cpp():
    T = "pennylane_lightning_gpu/src/tests/"
    kat = {}
    # 16-amplitude state + Pauli-word expvals, mpi/Test_StateVectorCudaMPI_NonParam.cpp:835-927
    sv16 = cpp_pairs(T + "mpi/Test_StateVectorCudaMPI_NonParam.cpp", 835, 851)
    assert len(sv16) == 16, len(sv16)
    kat["pauli_words"] = {
        "cite": T + "mpi/Test_StateVectorCudaMPI_NonParam.cpp:835-927",
        "state": sv16,
        "cases": [
            {"words": ["XYZI", "ZZXX"], "tgts": [[0, 1, 2, 3], [0, 1, 2, 3]],
             "coeffs": [0.1, 0.2], "expected": 0.0014895211, "atol": 1e-7},
            {"words": ["X", "Y", "Z", "I"], "tgts": [[0], [0], [0], [0]],
             "coeffs": [0.1, 0.2, 0.3, 0.4], "expected": 0.4589167637, "atol": 1e-7},
            {"words": ["X", "Y", "Z", "I"], "tgts": [[3], [3], [3], [3]],
             "coeffs": [0.1, 0.2, 0.3, 0.4], "expected": 0.4841317321, "atol": 1e-7},
            {"words": ["X", "XY", "XYZ", "XYZI"], "tgts": [[0], [0, 1], [0, 1, 2], [0, 1, 2, 3]],
             "coeffs": [0.1, 0.2, 0.3, 0.4], "expected": -0.0105768395, "atol": 1e-7},
        ],
    }
    # dense expval (1.263, -1.011), Test_StateVectorCudaManaged_NonParam.cpp:864-894
    st8 = cpp_pairs(T + "Test_StateVectorCudaManaged_NonParam.cpp", 867, 869)
    mat = cpp_pairs(T + "Test_StateVectorCudaManaged_NonParam.cpp", 873, 886)
    assert len(st8) == 8 and len(mat) == 64, (len(st8), len(mat))
    kat["expval_matrix"] = {
        "cite": T + "Test_StateVectorCudaManaged_NonParam.cpp:864-894",
        "state": st8, "wires": [0, 1, 2], "matrix": mat,
        "expected": [1.263, -1.011], "rtol": 1e-7,
    }
    # CSR expval == 1, Test_StateVectorCudaManaged_NonParam.cpp:897-932
    vals = cpp_pairs(T + "Test_StateVectorCudaManaged_NonParam.cpp", 916, 920)
    assert len(vals) == 16
    kat["expval_csr"] = {
        "cite": T + "Test_StateVectorCudaManaged_NonParam.cpp:897-932",
        "state": st8,
        "indptr": [0, 2, 4, 6, 8, 10, 12, 14, 16],
        "indices": [0, 3, 1, 2, 1, 2, 0, 3, 4, 7, 5, 6, 5, 6, 4, 7],
        "values": vals, "expected": 1.0, "rtol": 1e-7,
    }
    # RX / RY single-qubit vectors, Test_StateVectorCudaManaged_Param.cpp:29-163
    P = T + "Test_StateVectorCudaManaged_Param.cpp"
    rx = cpp_pairs(P, 35, 39)
    ry = cpp_pairs(P, 100, 106)
    ry_adj = cpp_pairs(P, 107, 113)
    ry_init = cpp_pairs(P, 115, 116)
    assert len(rx) == 6 and len(ry) == 6 and len(ry_adj) == 6 and len(ry_init) == 2
    gates = []
    # NB: the RX test lists 2 angles but 3 expected vectors; only the first two are used.
    for i, a in enumerate([0.1, 0.6]):
        gates.append({"cite": P + ":29-91", "gate": "RX", "wires": [0], "params": [a], "adjoint": False,
                      "input": [[1, 0], [0, 0]], "expected": rx[2 * i:2 * i + 2], "atol": 1e-7})
    for i, a in enumerate([0.2, 0.7, 2.9]):
        gates.append({"cite": P + ":93-163", "gate": "RY", "wires": [0], "params": [a], "adjoint": False,
                      "input": ry_init, "expected": ry[2 * i:2 * i + 2], "atol": 1e-7})
        gates.append({"cite": P + ":93-163", "gate": "RY", "wires": [0], "params": [a], "adjoint": True,
                      "input": ry_init, "expected": ry_adj[2 * i:2 * i + 2], "atol": 1e-7})

    # sparse-initialised expectations on |0..0>: (index -> value) tables
    def sparse_case(cite, gate, wires, angle, adjoint, n, entries):
        exp = [[0.0, 0.0] for _ in range(1 << n)]
        for k, v in entries.items():
            exp[k] = list(v)
        inp = [[0.0, 0.0] for _ in range(1 << n)]
        inp[0] = [1.0, 0.0]
        gates.append({"cite": cite, "gate": gate, "wires": wires, "params": [angle], "adjoint": adjoint,
                      "input": inp, "expected": exp, "atol": 1e-7})

    c3, s3 = 0.9887710779360422, 0.14943813247359922   # cos/sin(0.15)
    c8, s8 = 0.9210609940028851, 0.3894183423086505    # cos/sin(0.4)
    for ang, c, s in ((0.3, c3, s3), (0.8, c8, s8)):
        # IsingXX, Param.cpp:437-528
        sparse_case(P + ":437-528", "IsingXX", [0, 1], ang, False, 3, {0: (c, 0), 6: (0, -s)})
        sparse_case(P + ":437-528", "IsingXX", [0, 2], ang, False, 3, {0: (c, 0), 5: (0, -s)})
        sparse_case(P + ":437-528", "IsingXX", [0, 1], ang, True, 3, {0: (c, 0), 6: (0, s)})
        # SingleExcitationMinus/Plus, Param.cpp:737-875
        sparse_case(P + ":737-805", "SingleExcitationMinus", [0, 1], ang, False, 3, {0: (c, -s)})
        sparse_case(P + ":737-805", "SingleExcitationMinus", [0, 2], ang, True, 3, {0: (c, s)})
        sparse_case(P + ":807-875", "SingleExcitationPlus", [0, 1], ang, False, 3, {0: (c, s)})
        sparse_case(P + ":807-875", "SingleExcitationPlus", [0, 2], ang, True, 3, {0: (c, -s)})
        # DoubleExcitationMinus/Plus, Param.cpp:909-1005
        sparse_case(P + ":909-957", "DoubleExcitationMinus", [0, 1, 2, 3], ang, False, 4, {0: (c, -s)})
        sparse_case(P + ":959-1005", "DoubleExcitationPlus", [0, 1, 2, 3], ang, False, 4, {0: (c, s)})
        # SingleExcitation / DoubleExcitation leave |0..0> alone, Param.cpp:695-735, 877-907
        sparse_case(P + ":695-735", "SingleExcitation", [0, 1], ang, False, 3, {0: (1, 0)})
        sparse_case(P + ":877-907", "DoubleExcitation", [0, 1, 2, 3], ang, False, 4, {0: (1, 0)})
    kat["gates_cpp"] = gates
    return kat


# ---------------------------------------------------------------- (3) transcribed scalars
def transcribed():
    A = "pennylane_lightning_gpu/src/tests/Test_AdjointDiffGPU.cpp"
    p = [-math.pi / 7, math.pi / 5, 2 * math.pi / 3]
    jac = []
    jac.append({
        "cite": A + ":198-231", "n": 3, "init": "zero",
        "ops": [{"name": "RX", "wires": [i], "params": [p[i]], "adjoint": False} for i in range(3)],
        "obs": [["TensorProd", [["Named", "PauliZ", [0]], ["Named", "PauliZ", [1]], ["Named", "PauliZ", [2]]]]],
        "trainable": [0, 1, 2],
        "expected": [[-0.1755096592645253, 0.26478810666384334, -0.6312451595102775]], "atol": 1e-7,
    })
    names = ["RZ", "RY", "RZ", "CNOT", "CNOT", "RZ", "RY", "RZ"]
    pars = [[p[0]], [p[1]], [p[2]], [], [], [p[0]], [p[1]], [p[2]]]
    wires = [[0], [0], [0], [0, 1], [1, 2], [1], [1], [1]]
    jac.append({
        "cite": A + ":233-279", "n": 3, "init": "zero",
        "ops": [{"name": a, "wires": w, "params": q, "adjoint": False} for a, q, w in zip(names, pars, wires)],
        "obs": [["TensorProd", [["Named", "PauliX", [0]], ["Named", "PauliX", [1]], ["Named", "PauliX", [2]]]]],
        "trainable": [0, 1, 2, 3, 4, 5],
        "expected": [[0.0, -0.674214427, 0.275139672, 0.275139672, -0.0129093062, 0.323846156]], "atol": 1e-7,
    })
    # Decomposed Rot on (|0> - |1>)/sqrt2, thetas = linspace(-2pi, 2pi, 7); A:281-337
    thetas = np.linspace(-2 * math.pi, 2 * math.pi, 7)
    table = [[0.0, -9.90819496e-01, 0.0], [-8.18996553e-01, 1.62526544e-01, 0.0],
             [-0.203949, 0.48593716, 0.0], [0.0, 1.0, 0.0],
             [-2.03948985e-01, 4.85937177e-01, 0.0], [-8.18996598e-01, 1.62526487e-01, 0.0],
             [0.0, -9.90819511e-01, 0.0]]
    for th, row in zip(thetas, table):
        lp = [float(th), float(th ** 3), float(math.sqrt(2) * th)]
        jac.append({
            "cite": A + ":281-337", "n": 1,
            "init": [[1 / math.sqrt(2), 0.0], [-1 / math.sqrt(2), 0.0]],
            "ops": [{"name": g, "wires": [0], "params": [q], "adjoint": False}
                    for g, q in zip(["RZ", "RY", "RZ"], lp)],
            "obs": [["Named", "PauliZ", [0]]], "trainable": [0, 1, 2],
            "expected": [row], "atol": 1e-6,
        })
    # Mixed ops, t_params {1,2,3}; A:339-405
    lp = [0.543, 0.54, 0.1, 0.5, 1.3, -2.3, 0.5, -0.5, 0.5]
    names = ["Hadamard", "RX", "CNOT", "RZ", "RY", "RZ", "RZ", "RY", "RZ", "RZ", "RY", "CNOT"]
    pars = [[], [lp[0]], [], [lp[1]], [lp[2]], [lp[3]], [lp[4]], [lp[5]], [lp[6]], [lp[7]], [lp[8]], []]
    wires = [[0], [0], [0, 1], [0], [0], [0], [0], [0], [0], [0], [1], [0, 1]]
    jac.append({
        "cite": A + ":339-405", "n": 2, "init": "zero",
        "ops": [{"name": a, "wires": w, "params": q, "adjoint": False} for a, q, w in zip(names, pars, wires)],
        "obs": [["TensorProd", [["Named", "PauliX", [0]], ["Named", "PauliZ", [1]]]]],
        "trainable": [1, 2, 3],
        "expected": [[-0.71429188, 0.04998561, -0.71904837]], "rtol": 1e-5,
    })
    # Hamiltonian observables; A:478-545
    jac.append({
        "cite": A + ":478-508", "n": 2, "init": "zero",
        "ops": [{"name": "RX", "wires": [0], "params": [p[0]], "adjoint": False}],
        "obs": [["Hamiltonian", [0.3, 0.7], [["Named", "PauliZ", [0]], ["Named", "PauliZ", [1]]]]],
        "trainable": [0], "expected": [[-0.3 * math.sin(p[0])]], "atol": 1e-7,
    })
    jac.append({
        "cite": A + ":510-545", "n": 3, "init": "zero",
        "ops": [{"name": "RX", "wires": [i], "params": [p[i]], "adjoint": False} for i in range(3)],
        "obs": [["Hamiltonian", [0.47, 0.32, 0.96],
                 [["Named", "PauliZ", [0]], ["Named", "PauliZ", [1]], ["Named", "PauliZ", [2]]]]],
        "trainable": [0, 2],
        "expected": [[-0.47 * math.sin(p[0]), -0.96 * math.sin(p[2])]], "atol": 1e-7,
    })
    # tests/test_hamiltonian_sparse.py:73-101: RX(0.4) w0, RY(-0.2) w1, expval of a 2-wire Pauli word
    sparse = {
        "cite": "tests/test_hamiltonian_sparse.py:73-101",
        "ops": [{"name": "RX", "wires": [0], "params": [0.4]}, {"name": "RY", "wires": [1], "params": [-0.2]}],
        "cases": [["XI", 0.0], ["IX", -0.19866933079506122], ["YI", -0.38941834230865050],
                  ["IY", 0.0], ["ZI", 0.92106099400288520], ["IZ", 0.98006657784124170]],
        "atol": 1e-4,
    }
    # tests/test_probs.py:80-104: RX(0.4) Rot(0.5,0.3,-0.7) RY(-0.2) on wire 0 of 2
    probs = {
        "cite": "tests/test_probs.py:80-104",
        "n": 2,
        "ops": [{"name": "RX", "wires": [0], "params": [0.4]},
                {"name": "Rot", "wires": [0], "params": [0.5, 0.3, -0.7]},
                {"name": "RY", "wires": [0], "params": [-0.2]}],
        "cases": [[[0, 1], [0.9165164490394898, 0.0, 0.08348355096051052, 0.0]],
                  [[0], [0.9165164490394898, 0.08348355096051052]]],
        "atol": 1e-4,
    }
    return {"adjoint": jac, "sparse_pauli": sparse, "probs": probs}


def closed_forms():
    """tests/test_expval.py:38-200 and tests/test_var.py:34-130: 3-wire circuits with the closed-form expectation values and
    variances the reference asserts (evaluated here at the tests' own parameter grids; tolerance = the tests' default
    atol 1e-8 loosened to 1e-7 for the formulas' rounding)."""
    theta_s = np.linspace(0.11, 1, 3)
    phi_s = np.linspace(0.32, 1, 3)
    varphi_s = np.linspace(0.02, 1, 3)
    cos, sin, sqrt = math.cos, math.sin, math.sqrt
    E, V = "tests/test_expval.py", "tests/test_var.py"

    def op(name, wires, *params):
        return {"name": name, "wires": list(wires), "params": [float(x) for x in params]}

    def named(name, w):
        return ["Named", name, [w]]

    out = []
    for th, ph in zip(theta_s, phi_s):
        rx = [op("RX", [0], th), op("RX", [1], ph), op("CNOT", [0, 1])]
        ry = [op("RY", [0], th), op("RY", [1], ph), op("CNOT", [0, 1])]
        out += [
            {"cite": E + ":47-60", "ops": rx, "obs": named("Identity", 0), "expval": 1.0},
            {"cite": E + ":47-60", "ops": rx, "obs": named("Identity", 1), "expval": 1.0},
            {"cite": E + ":62-74", "ops": rx, "obs": named("PauliZ", 0), "expval": cos(th)},
            {"cite": E + ":62-74", "ops": rx, "obs": named("PauliZ", 1), "expval": cos(th) * cos(ph)},
            {"cite": E + ":76-88", "ops": ry, "obs": named("PauliX", 0), "expval": sin(th) * sin(ph)},
            {"cite": E + ":76-88", "ops": ry, "obs": named("PauliX", 1), "expval": sin(ph)},
            {"cite": E + ":90-102", "ops": rx, "obs": named("PauliY", 0), "expval": 0.0},
            {"cite": E + ":90-102", "ops": rx, "obs": named("PauliY", 1), "expval": -cos(th) * sin(ph)},
            {"cite": E + ":104-124", "ops": ry, "obs": named("Hadamard", 0),
             "expval": (sin(th) * sin(ph) + cos(th)) / sqrt(2)},
            {"cite": E + ":104-124", "ops": ry, "obs": named("Hadamard", 1),
             "expval": (cos(th) * cos(ph) + sin(ph)) / sqrt(2)},
            {"cite": V + ":44-63", "ops": [op("RX", [0], ph), op("RY", [0], th)], "obs": named("PauliZ", 0),
             "var": 0.25 * (3 - cos(2 * th) - 2 * cos(th) ** 2 * cos(2 * ph))},
        ]
    for th, ph, vp in zip(theta_s, phi_s, varphi_s):
        c3 = [op("RX", [0], th), op("RX", [1], ph), op("RX", [2], vp), op("CNOT", [0, 1]), op("CNOT", [1, 2])]
        xy = ["TensorProd", [named("PauliX", 0), named("PauliY", 2)]]
        ziz = ["TensorProd", [named("PauliZ", 0), named("Identity", 1), named("PauliZ", 2)]]
        zhy = ["TensorProd", [named("PauliZ", 0), named("Hadamard", 1), named("PauliY", 2)]]
        out += [
            {"cite": E + ":129-149", "ops": c3, "obs": xy, "expval": sin(th) * sin(ph) * sin(vp)},
            {"cite": E + ":151-172", "ops": c3, "obs": ziz, "expval": cos(vp) * cos(ph)},
            {"cite": E + ":174-195", "ops": c3, "obs": zhy,
             "expval": -(cos(vp) * sin(ph) + sin(vp) * cos(th)) / sqrt(2)},
            {"cite": V + ":69-96", "ops": c3, "obs": xy,
             "var": (8 * sin(th) ** 2 * cos(2 * vp) * sin(ph) ** 2 - cos(2 * (th - ph)) - cos(2 * (th + ph))
                     + 2 * cos(2 * th) + 2 * cos(2 * ph) + 14) / 16},
            {"cite": V + ":98-125", "ops": c3, "obs": zhy,
             "var": (3 + cos(2 * ph) * cos(vp) ** 2 - cos(2 * th) * sin(vp) ** 2
                     - 2 * cos(th) * sin(ph) * sin(2 * vp)) / 4},
        ]
    for case in out:
        case["n"] = 3
        case["atol"] = 1e-7
    return {"closed_forms": out}


def main():
    kats = {"_about": "generated by tests/golden/make_golden.py from /root/reference; do not edit"}
    kats["gates_py"] = extract_test_apply()
    kats.update(extract_cpp())
    kats.update(transcribed())
    kats.update(closed_forms())
    with open(OUT, "w") as f:
        json.dump(kats, f, indent=1)
    print("wrote", OUT, {k: (len(v) if isinstance(v, list) else "-") for k, v in kats.items()})


if __name__ == "__main__":
    main()
