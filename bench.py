#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native lightning.gpu hot path.

Metric (BASELINE.json): "30q gate-apply HBM GB/s".  One step = one pass of BASELINE config 2(ii), the
200-gate random 1/2-qubit circuit {RX, RY, RZ, CNOT, CZ, QubitUnitary 1q/2q} (default_rng(2024)), over a
30-qubit complex128 state (16 GiB, >> 126 MB L2, so no L2 flush is needed between steps).
value = algorithmic bytes of the circuit (sum over gates of 2*B*N/2^c, no credit
for fusion) / device time, state resident in HBM.  e2e = the same circuit through the public API from a
HOST state (pinned H2D of the initial state, op records built from host data, <Z0> read back).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--qubits 30]
                    [--dtype c128|c64] [--fuse 1|0] [--sweeps 1|0]

N > 1 (torchrun, one rank per GPU): weak scaling, 30 local qubits per GPU, n = 30 + log2(N) qubits in
total, the same circuit generator over all n wires; gates on the log2(N) global wires trigger
index-bit swaps over NCCL send/recv.
--impl reference: the CPU restatement of lightning.qubit (oracle/lq_port.c, OpenMP, all host cores) on
a bounded sample of the same circuit.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pennylane_lightning_gpu_b200 import workloads  # noqa: E402

METRIC = "30q gate-apply HBM GB/s"
UNIT = "GB/s"


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi-equivalent sampling through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        if self._nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t is not None:
            self._t.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def circuit_bytes(ops, n, amp_bytes):
    return sum(workloads.gate_bytes(op, n, amp_bytes) for op in ops)


# -------------------------------------------------------------------------------------------------
# CPU arm: lightning.qubit restatement (oracle/lq_port.c) on a bounded sample of the workload
# -------------------------------------------------------------------------------------------------
def cpu_sample(n_target: int, budget_s: float = 15.0):
    """Runs as many gates of the C2 circuit as fit in `budget_s` on all host cores.
    -> (GB/s over those gates, cores, description, ms per gate)"""
    from oracle import lq_port as lq

    cores = os.cpu_count() or 1
    lq.set_num_threads(cores)
    try:
        avail = os.sysconf("SC_AVPHYS_PAGES") * os.sysconf("SC_PAGE_SIZE")
    except (ValueError, OSError):
        avail = 32 << 30
    n = n_target
    while (16 << n) * 2.5 > avail and n > 20:
        n -= 1
    ops = workloads.random_gate_circuit(n, 200, 2024)
    st = lq.LQState(n)
    for w in range(min(n, 4)):  # touch the pages / leave |0...0>
        st.apply_op("Hadamard", [w])
    t0 = time.perf_counter()
    done, nbytes = 0, 0
    for op in ops:
        st.apply_op(op["name"], op["wires"], op.get("params", ()), False, op.get("matrix"))
        done += 1
        nbytes += workloads.gate_bytes(op, n, 16)
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    desc = (f"first {done} of the 200 gates of the config-2 random circuit at {n} qubits complex128 "
            f"(oracle/lq_port.c, OpenMP, {cores} threads, {dt:.1f} s)")
    return nbytes / dt / 1e9, cores, desc, dt / done * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    warm = max(args.warmup, 0)
    budget = max(2.0, min(20.0, 120.0 / (steps + warm)))
    for _ in range(warm):
        cpu_sample(args.qubits, budget_s=budget / 2)
    vals, ms = [], []
    desc, cores = "", 1
    for _ in range(steps):
        v, cores, desc, m = cpu_sample(args.qubits, budget_s=budget)
        vals.append(v)
        ms.append(m)
    v = float(np.mean(vals))
    n_total = args.qubits + int(math.log2(max(args.gpus, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": float(np.mean(ms)) * 200, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config 2(ii): 200-gate random 1/2-qubit circuit, {n_total} qubits, complex128",
                   "note": "ms_per_step extrapolated from the sampled gates to the 200-gate circuit"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch

    import pennylane_lightning_gpu_b200 as q

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    distributed = world > 1
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if distributed:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    n_local = args.qubits
    n_glob = int(math.log2(world))
    assert 1 << n_glob == world, "number of GPUs must be a power of two"
    n_total = n_local + n_glob
    cdtype = np.complex128 if args.dtype == "c128" else np.complex64
    amp_bytes = 16 if args.dtype == "c128" else 8
    tdtype = torch.complex128 if args.dtype == "c128" else torch.complex64
    hbm_peak, peak_src = measured_peaks()

    ops = workloads.random_gate_circuit(n_total, args.gates, 2024)
    alg_bytes_total = circuit_bytes(ops, n_total, amp_bytes)  # whole job (all ranks)

    # state in torch-owned HBM, initialised on the device: normalised random state, seed 1234
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    real_dtype = torch.float64 if args.dtype == "c128" else torch.float32
    buf = torch.empty((1 << n_local) * 2, dtype=real_dtype, device=dev)
    chunk = 1 << 26
    for s in range(0, buf.numel(), chunk):
        buf[s:s + chunk].normal_(generator=gen)
    nrm2 = torch.zeros((), dtype=torch.float64, device=dev)
    for s in range(0, buf.numel(), chunk):
        nrm2 += buf[s:s + chunk].double().square().sum()
    if distributed:
        dist.all_reduce(nrm2)
    buf.mul_(1.0 / math.sqrt(float(nrm2)))

    if distributed:
        from pennylane_lightning_gpu_b200.distributed import DistributedStateVector

        sv = DistributedStateVector(n_total, cdtype, device=local_rank, external_ptr=buf.data_ptr())
    else:
        sv = q.StateVector(n_local, cdtype, device=local_rank, external_ptr=buf.data_ptr())
    rec = q.Ops(ops)

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        sv.apply_ops(rec, fuse=bool(args.fuse))
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = 0
    sweeps = 0
    with ClockSampler(local_rank) as clk:
        barrier()
        e0.record()
        for _ in range(args.steps):
            sv.apply_ops(rec, fuse=bool(args.fuse))
            l, s = sv.last_apply_stats()
            launches += l
            sweeps += s
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms)
    ms_per_step = ms_total / args.steps
    value = alg_bytes_total / (ms_per_step * 1e-3) / 1e9

    # roofline of the dominant kernel: the gate / tile sweep kernels are the only kernels in the
    # timed region, so average launch duration = region time / launches (CUDA events, this stream).
    per_launch_ms = ms_total / max(launches, 1)
    if args.fuse and sweeps < len(ops) * args.steps:
        # fused tile kernel: every launch reads and writes the whole local shard once
        bytes_per_launch = 2 * amp_bytes * (1 << n_local)
        kernel = "k_tile_sweep (fused shared-memory tile kernel), 2*B*N_local bytes per launch"
    else:
        bytes_per_launch = alg_bytes_total / world / max(len(ops), 1)
        kernel = "k_apply_dense / k_apply_diag (one sweep per gate), mean algorithmic bytes per gate"
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": None, "kernel": kernel, "peak_source": peak_src,
                "bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms}

    detail = {}
    if distributed:
        n_swaps, swap_bytes, swap_ms = sv.swap_stats()
        # per-direction NVLink bandwidth of the global<->local index-bit swaps (device time of the
        # NCCL send/recv stream, this rank) over warm-up + timed steps
        detail["nvlink_swaps"] = {
            "n_swaps_per_step": n_swaps / (max(args.warmup, 3) + args.steps),
            "gb_sent_per_swap": (swap_bytes / max(n_swaps, 1)) / 1e9,
            "gbs_per_direction": (swap_bytes / 1e9) / (swap_ms * 1e-3) if swap_ms > 0 else None,
            "frac_of_770_measured_peer_copy": ((swap_bytes / 1e9) / (swap_ms * 1e-3) / 770.0) if swap_ms > 0 else None,
            "swap_ms_per_step": swap_ms / (max(args.warmup, 3) + args.steps),
        }
        kernel_ms = ms_total - detail["nvlink_swaps"]["swap_ms_per_step"] * args.steps
        if launches > 0 and kernel_ms > 0:
            roofline["ms_per_launch"] = kernel_ms / launches
            roofline["achieved"] = bytes_per_launch / (roofline["ms_per_launch"] * 1e-3) / 1e9
            roofline["frac"] = roofline["achieved"] / hbm_peak
            roofline["note"] = "exchange time (NCCL stream) subtracted from the region before dividing by launches"
    # ---- single-gate sweeps: C2(i), every target wire individually ----------------------------
    if args.sweeps and not distributed:
        detail["single_gate_sweeps"] = single_gate_sweeps(torch, q, sv, n_local, amp_bytes, hbm_peak)

    # ---- end to end through the public API with host buffers ----------------------------------
    e2e = None
    if not distributed and args.e2e:
        e2e = end_to_end(torch, q, sv, buf, ops, n_local, cdtype, tdtype, amp_bytes, alg_bytes_total, args)

    line = None
    if rank == 0:
        cpu = None
        if args.cpu_baseline and not distributed:
            v, cores, desc, _ = cpu_sample(n_local, budget_s=15.0)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.dtype == "c128" else "f32", "data": "synthetic",
            "config": {"workload": f"config 2(ii): {args.gates}-gate random 1/2-qubit circuit, {n_total} qubits, "
                                   f"{'complex128' if args.dtype == 'c128' else 'complex64'}",
                       "local_qubits": n_local, "global_qubits": n_glob, "fused": bool(args.fuse),
                       "l2": "state (16 GiB) >> L2 (126 MB): no flush needed",
                       "bytes_rule": "sum over gates of 2*B*N/2^controls (SURVEY.md 8d), no credit for fusion"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "hbm_sweeps": sweeps, "clocks": clk.summary(), "detail": detail,
        }
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return line


def single_gate_sweeps(torch, q, sv, n, amp_bytes, hbm_peak):
    """Every target wire, RX / RZ / generic 2x2 / CNOT(i,i+1) / generic 4x4(i,i+1): GB/s per gate."""
    rng = np.random.default_rng(7)
    u2 = workloads.haar_unitary(rng, 2)
    u4 = workloads.haar_unitary(rng, 4)
    out = {}

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    full = 2 * amp_bytes * (1 << n)
    cases = {
        "RX": (lambda w: sv.apply("RX", [w], [0.3]), full, range(n)),
        "RZ": (lambda w: sv.apply("RZ", [w], [1.1]), full, range(n)),
        "U2": (lambda w: sv.apply_matrix(u2, [w]), full, range(n)),
        "CNOT_adjacent": (lambda w: sv.apply("CNOT", [w, w + 1]), full // 2, range(n - 1)),
        "U4_adjacent": (lambda w: sv.apply_matrix(u4, [w, w + 1]), full, range(n - 1)),
    }
    for name, (fn, nbytes, wires) in cases.items():
        gbs = [nbytes / (timed(lambda: fn(w)) * 1e-3) / 1e9 for w in wires]
        out[name] = {"min_gbs": min(gbs), "median_gbs": float(np.median(gbs)), "max_gbs": max(gbs),
                     "min_frac_of_peak": min(gbs) / hbm_peak, "median_frac_of_peak": float(np.median(gbs)) / hbm_peak,
                     "per_wire_gbs": [round(g, 1) for g in gbs]}
    return out


def end_to_end(torch, q, sv, buf, ops, n, cdtype, tdtype, amp_bytes, alg_bytes, args):
    """StatePrep(host state) -> circuit -> <Z0> on the host, all through the public API."""
    try:
        host = torch.empty(1 << n, dtype=tdtype, pin_memory=True)
    except RuntimeError:
        return None
    host.copy_(torch.view_as_complex(buf.view(-1, 2)))
    torch.cuda.synchronize()
    host_np = host.numpy()
    z0 = q.Observable.named("PauliZ", [0])
    h2d = host_np.nbytes + sum(16 * len(np.atleast_1d(op.get("params", ()))) + (op["matrix"].nbytes if "matrix" in op else 0)
                               for op in ops)

    def step():
        sv.h2d(host_np)                      # HostToDevice of the prepared state (lightning_gpu.py:392-447)
        rec = q.Ops(ops)                     # host op list -> recorded circuit
        sv.apply_ops(rec, fuse=bool(args.fuse))
        return sv.expval(z0)                 # one double back to the host

    step()
    torch.cuda.synchronize()
    k = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(k):
        val = step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / k
    return {"value": alg_bytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8,
            "ms_per_step": dt * 1e3, "result_check": float(val)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=30, help="local qubits per GPU")
    ap.add_argument("--gates", type=int, default=200)
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    ap.add_argument("--fuse", type=int, default=1)
    ap.add_argument("--sweeps", type=int, default=1)
    ap.add_argument("--e2e", type=int, default=1)
    ap.add_argument("--cpu-baseline", dest="cpu_baseline", type=int, default=1)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
