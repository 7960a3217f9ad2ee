"""Parity at the sizes BASELINE.json names: config 2 at 24 qubits against the CPU restatement of
lightning.qubit (oracle/lq_port.c -- pinned to the NumPy oracle and through it to the reference's golden vectors by
tests/test_oracle_lq_port.py), and complex64 against complex128 at the full 30 qubits on the device.

Runs on the B200 box only (``-m gpu``); nothing here reads /root/reference."""
import numpy as np
import pytest

from conftest import random_state

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def q():
    import pennylane_lightning_gpu_b200 as q

    return q


def test_config2_24_qubits_vs_cpu_port(q):
    """The bench's own circuit (200 gates drawn from RX/RY/RZ/CNOT/CZ/QubitUnitary 1q/2q, default_rng(2024)) at 24 qubits:
    fused sweeps and one sweep per gate against lq_port, 1e-10 (complex128) / 1e-5 (complex64), no widening."""
    import os

    from oracle import lq_port as lq
    from pennylane_lightning_gpu_b200 import workloads

    n = 24
    ops = workloads.random_gate_circuit(n, 200, 2024)
    psi = random_state(n, 2024)
    lq.set_num_threads(os.cpu_count() or 1)
    st = lq.LQState(n, psi)
    st.apply_ops(ops)
    want = st.sv
    rec = q.Ops(ops)
    for dtype, tol in ((np.complex128, 1e-10), (np.complex64, 1e-5)):
        for fuse in (True, False):
            sv = q.StateVector(n, dtype)
            sv.h2d(psi.astype(dtype))
            sv.apply_ops(rec, fuse=fuse)
            got = sv.d2h()
            err = float(np.max(np.abs(got - want)))
            assert err <= tol, f"{np.dtype(dtype).name} fuse={fuse}: max abs err {err:.3e}"
            # amplitudes are ~2.4e-4 here, so also a relative statement: the state as a whole
            rel = float(np.linalg.norm(got.astype(np.complex128) - want))
            assert rel <= (1e-10 if dtype == np.complex128 else 2e-5), f"{np.dtype(dtype).name} fuse={fuse}: l2 err {rel:.3e}"
            del sv


def test_config2_30_qubits_complex64_vs_complex128(q):
    """'also compare c64 vs c128 GPU at 30q, tol 1e-5' -- the full-size state never leaves the device."""
    import torch

    from pennylane_lightning_gpu_b200 import workloads

    free, _ = torch.cuda.mem_get_info()
    if free < (30 << 30):
        pytest.skip("needs 30 GiB of free device memory")
    n = 30
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234)
    b128 = torch.empty((1 << n) * 2, dtype=torch.float64, device=dev)
    chunk = 1 << 26
    for s in range(0, b128.numel(), chunk):
        b128[s:s + chunk].normal_(generator=gen)
    nrm2 = sum(float(b128[s:s + chunk].square().sum()) for s in range(0, b128.numel(), chunk))
    b128.mul_(1.0 / np.sqrt(nrm2))
    b64 = torch.empty((1 << n) * 2, dtype=torch.float32, device=dev)
    for s in range(0, b128.numel(), chunk):
        b64[s:s + chunk].copy_(b128[s:s + chunk])
    ops = workloads.random_gate_circuit(n, 200, 2024)
    rec = q.Ops(ops)
    sv128 = q.StateVector(n, np.complex128, external_ptr=b128.data_ptr())
    sv64 = q.StateVector(n, np.complex64, external_ptr=b64.data_ptr())
    sv128.apply_ops(rec, fuse=True)
    sv64.apply_ops(rec, fuse=True)
    torch.cuda.synchronize()
    worst, dot_re, dot_im, n64, n128 = 0.0, 0.0, 0.0, 0.0, 0.0
    for s in range(0, b128.numel(), chunk):
        a = b128[s:s + chunk]
        b = b64[s:s + chunk].double()
        worst = max(worst, float((a - b).abs().max()))
        ar, ai, br, bi = a[0::2], a[1::2], b[0::2], b[1::2]
        dot_re += float((ar * br + ai * bi).sum())
        dot_im += float((ar * bi - ai * br).sum())
        n64 += float(b.square().sum())
        n128 += float(a.square().sum())
    assert worst <= 1e-5, f"max |c64 - c128| = {worst:.3e}"
    fidelity = (dot_re ** 2 + dot_im ** 2) / n64
    assert abs(fidelity - 1.0) <= 1e-5 and abs(n64 - 1.0) <= 1e-5, (fidelity, n64)
    # the meaningful statement at amplitudes of 3e-5: l2 distance of the two (unit-norm) states, i.e. a relative error
    dist2 = n128 + n64 - 2.0 * dot_re
    assert abs(n128 - 1.0) <= 1e-10
    assert dist2 <= 1e-8, f"||psi64 - psi128||^2 = {dist2:.3e} (norms^2 {n128:.12f}, {n64:.12f})"


def test_probabilities_at_24_qubits(q):
    """csrc/measure_kernels.cu: k_probs_rows at a size where every branch of its work split is taken (one chunk per group
    without atomics, several chunks per group with atomics, measured wires among the five lane bits, in any order)
    against NumPy marginals of the same device state."""
    from pennylane_lightning_gpu_b200 import workloads

    n = 24
    sv = q.StateVector(n, np.complex128)
    sv.apply_ops(q.Ops(workloads.random_gate_circuit(n, 120, 7)), fuse=True)
    psi = sv.d2h()
    p_full = (np.abs(psi) ** 2).reshape([2] * n)
    for wires in ([0], [n - 1], [3, 17], [n - 1, 0], [5, n - 2, 11], list(range(10)), list(range(n - 8, n)),
                  [n - 3, 2, n - 1, 9, 20], list(range(0, n, 2)), list(range(n))):
        got = sv.probs(wires)
        # first listed wire = least significant bit of the output index (cuStateVec order, StateVectorCudaManaged.hpp:931-970)
        keep = list(wires)
        marg = p_full.sum(axis=tuple(a for a in range(n) if a not in keep)) if len(keep) < n else p_full
        order = sorted(keep)
        marg = np.transpose(marg, [order.index(w) for w in reversed(keep)]).reshape(-1)
        assert got.shape == marg.shape
        assert np.max(np.abs(got - marg)) <= 1e-12, wires
    s32 = q.StateVector(n, np.complex64)
    s32.h2d(psi.astype(np.complex64))
    keep = [1, n - 1, 12]
    got = s32.probs(keep)
    marg = p_full.sum(axis=tuple(a for a in range(n) if a not in keep))
    marg = np.transpose(marg, [sorted(keep).index(w) for w in reversed(keep)]).reshape(-1)
    assert np.max(np.abs(got - marg)) <= 1e-6
