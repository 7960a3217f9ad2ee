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
    -> (GB/s over those gates, cores, description, ms per gate, gates done, seconds)"""
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
    return nbytes / dt / 1e9, cores, desc, dt / done * 1e3, done, dt


def run_reference(args):
    """The reference arm: the CPU restatement of lightning.qubit's pair-strided kernels (oracle/lq_port.c; the reference
    itself cannot be built or imported here) on all host cores.  One step = a bounded SAMPLE of the
    workload (the first gates of the same circuit that fit the time budget); `value` is the rate over the sample, and
    `ms_per_step` is the measured duration of that sample, not an extrapolation."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(args.steps, 1)
    warm = max(args.warmup, 0)
    budget = max(2.0, min(20.0, 120.0 / (steps + warm)))
    for _ in range(warm):
        cpu_sample(args.qubits, budget_s=budget / 2)
    vals, ms, gates = [], [], []
    desc, cores = "", 1
    for _ in range(steps):
        v, cores, desc, _, done, dt = cpu_sample(args.qubits, budget_s=budget)
        vals.append(v)
        ms.append(dt * 1e3)
        gates.append(done)
    v = float(np.mean(vals))
    n_total = args.qubits + int(math.log2(max(args.gpus, 1)))
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": float(np.mean(ms)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"config 2(ii): 200-gate random 1/2-qubit circuit, {n_total} qubits, complex128",
                   "sample_gates_per_step": float(np.mean(gates)),
                   "note": "a step is a bounded sample: the first gates of the circuit that fit the time budget; ms_per_step "
                           "is the measured time of that sample; the full 200-gate step would take ms_per_step * 200 / "
                           "sample_gates_per_step"},
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
    clk_all = ClockSampler(local_rank)  # the whole process, beside the sampler of the timed region below
    clk_all.__enter__()
    if distributed:
        import torch.distributed as dist

        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep stdout to the one JSON line
        dist.init_process_group("nccl", device_id=dev)

    n_local = args.qubits
    n_glob = int(math.log2(world))
    assert 1 << n_glob == world, "number of GPUs must be a power of two"
    n_total = n_local + n_glob
    cdtype = np.complex128 if args.dtype == "c128" else np.complex64
    amp_bytes = 16 if args.dtype == "c128" else 8
    tdtype = torch.complex128 if args.dtype == "c128" else torch.complex64
    hbm_peak, peak_src = measured_peaks()

    circuit = args.circuit if args.circuit != "auto" else ("config5" if n_local >= 32 else "config2")
    if circuit == "config5":
        # BASELINE config 5: 4 layers of [random one-qubit rotation on every wire, CNOTs on a random perfect matching]
        ops = workloads.random_layer_circuit(n_total, layers=4, seed=99)
        workload = f"config 5: 4-layer random-layer circuit ({len(ops)} gates), {n_total} qubits"
    else:
        ops = workloads.random_gate_circuit(n_total, args.gates, 2024)
        workload = f"config 2(ii): {args.gates}-gate random 1/2-qubit circuit, {n_total} qubits"
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

    parity = None
    if distributed:
        from pennylane_lightning_gpu_b200.distributed import DistributedStateVector

        # SCALE carries its own check: the same circuit generator on a reduced register (20 local qubits), sharded
        # engine with the exchange schedule / fused exchanges of this run against the single-GPU engine on every rank
        parity = dist_self_check(torch, q, DistributedStateVector, dist, args, local_rank, rank, n_glob, cdtype)
        # Up to 31 local qubits the register owns its shard, so that exchanges fused into sweeps may leave it in either of
        # its two buffers (no copy back); beyond that there is no room for a second buffer and it borrows torch's.
        owned = 3 * (amp_bytes << n_local) < (150 << 30)
        sv = DistributedStateVector(n_total, cdtype, device=local_rank, external_ptr=None if owned else buf.data_ptr())
        if owned:
            sv.local.copy_from(q.StateVector(n_local, cdtype, device=local_rank, external_ptr=buf.data_ptr()))
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
    if distributed:
        sv.swap_stats(reset=True)  # connection set-up of the first exchanges stays out of the statistics
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
    fused = bool(args.fuse) and sweeps < len(ops) * args.steps
    # algorithmic bytes per launch per SURVEY.md 8d: the gates a launch applies, 2*B*N/2^c each, no credit for
    # fusion -- so a fused sweep can exceed 1.0x of the copy bandwidth; the actual DRAM rate is reported beside it
    bytes_per_launch = alg_bytes_total / world * args.steps / max(launches, 1)
    achieved = bytes_per_launch / (per_launch_ms * 1e-3) / 1e9
    sweep_bytes = 2 * amp_bytes * (1 << n_local)
    regs_kernel = os.environ.get("QSV_TILE_KERNEL", "1") == "1" and n_local >= 12
    fused_kernel = ("k_tile_regs (register-blocked fused tile kernel: 16 amplitudes per thread, DAG-packed sweeps, "
                    "list-scheduled passes, swizzled shared-memory transposes with PauliX/CNOT/SWAP folded into the "
                    "address maps)") if regs_kernel else \
        "k_tile_sweep (first-generation fused shared-memory tile kernel, TMA-staged)"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": sweep_bytes if fused else None,
                "traffic_source": ("ncu dram__bytes_read+write per fused sweep launch = 34.3 GB = 2*B*N at 30q c128 "
                                   "(profiles/r1_ncu_regs_final_summary.txt)") if fused else
                                  "ncu: 34.4 GB per full-state gate launch (profiles/r1_ncu_summary.txt); controlled gates move less",
                "kernel": fused_kernel if fused else "k_apply_dense / k_apply_diag (one sweep per gate)",
                "peak_source": peak_src, "bytes_per_launch": bytes_per_launch, "ms_per_launch": per_launch_ms,
                "gates_per_launch": len(ops) * args.steps / max(launches, 1)}
    roofline["frac_note"] = ("`frac` follows SURVEY.md 8d literally (algorithmic bytes of every gate a launch applies, no credit for "
                             "fusion), so for fused sweeps it is a speed-up over one-sweep-per-gate, NOT a fraction of a roof; "
                             "`frac_physical` = DRAM bytes actually moved per launch / time / measured HBM peak and "
                             "`frac_compute` = executed FP multiply-adds / time / FP pipe peak are the roofline fractions")
    roofline["frac_physical"] = roofline["frac"]
    if fused:
        roofline["hbm_actual_gbs"] = sweep_bytes / (per_launch_ms * 1e-3) / 1e9
        roofline["hbm_actual_frac"] = roofline["hbm_actual_gbs"] / hbm_peak
        roofline["frac_physical"] = roofline["hbm_actual_frac"]
        roofline["note"] = (f"a fused sweep applies ~{len(ops) * args.steps / max(launches, 1):.0f} gates per read+write of the "
                            "state, so it is bound by FP64 issue (64 DFMA/clk/SM) and instruction overhead rather than by HBM: "
                            "see DESIGN.md 4.2; detail.unfused_gate_by_gate and detail.single_gate_sweeps are the HBM-bound "
                            "one-sweep-per-gate kernels") \
            if regs_kernel else \
            ("fused sweeps are shared-memory-bandwidth-bound (1024 clk per gate per 64 KiB tile), not "
             "HBM-bound: see DESIGN.md 4.2; --fuse 0 gives the HBM-bound one-sweep-per-gate kernels")

    if fused and regs_kernel and not distributed:
        # the compute view of the same launches: fused multiply-adds the planner's programs execute (host-side count,
        # qsv_ops_plan_work) against the FP64 / FP32 pipe: 64 DFMA (128 FFMA) per clock and SM, 148 SMs
        try:
            work = rec.plan_work(n_local, cdtype)
            flops = 2.0 * work["fma_per_amplitude"] * float(1 << n_local)
            per_clk = 64 if args.dtype == "c128" else 128
            roofline["compute"] = {"pipe": "fp64" if args.dtype == "c128" else "fp32",
                                   "fma_per_amplitude": work["fma_per_amplitude"], "passes_per_step": work["passes"],
                                   "tflops": flops / (ms_per_step * 1e-3) / 1e12,
                                   "peak_tflops_per_ghz": 148 * per_clk * 2 / 1e3}
        except Exception as e:  # a reporting extra must not cost the bench line
            roofline["compute"] = {"error": str(e)}

    detail = {}
    if distributed:
        detail["parity_check"] = parity
        n_swaps, swap_bytes, swap_ms = sv.swap_stats()
        # per-direction NVLink bandwidth of the global<->local index-bit swaps (device time of the
        # NCCL send/recv stream, this rank) over warm-up + timed steps
        detail["nvlink_swaps"] = {
            "transport": "peer load/store kernel over CUDA IPC mappings" if sv.uses_peer_access else "staged ncclSend/ncclRecv",
            "schedule": "program order, farthest-next-use eviction" if os.environ.get("QSV_DIST_DAG") == "0" else
                        "dependency order (gates run while any ready gate is local)",
            "n_swaps_per_step": n_swaps / args.steps,
            "gb_sent_per_swap": (swap_bytes / max(n_swaps, 1)) / 1e9,
            "gbs_per_direction": (swap_bytes / 1e9) / (swap_ms * 1e-3) if swap_ms > 0 else None,
            "frac_of_770_measured_peer_copy": ((swap_bytes / 1e9) / (swap_ms * 1e-3) / 770.0) if swap_ms > 0 else None,
            "swap_ms_per_step": swap_ms / args.steps,
        }
        try:
            n_oop, n_carried = sv.fused_exchange_stats()
        except Exception:  # a reporting extra must not cost the bench line
            n_oop, n_carried = 0, 0
        if n_oop:  # exchanges carried by the sweeps around them (not part of the swap timing above)
            total_steps = args.steps + max(args.warmup, 3)
            detail["nvlink_swaps"]["out_of_place_exchanges_per_step"] = n_oop / total_steps
            detail["nvlink_swaps"]["carried_by_sweeps_per_step"] = n_carried / total_steps
            try:
                n_pull, n_pull_carried = sv.split_exchange_stats()
                detail["nvlink_swaps"]["split_exchanges_per_step"] = n_pull / total_steps
                detail["nvlink_swaps"]["second_halves_carried_by_sweeps_per_step"] = n_pull_carried / total_steps
            except Exception:  # a reporting extra must not cost the bench line
                pass
        if circuit == "config5":
            sw = detail["nvlink_swaps"]
            detail["config5"] = {
                "n_qubits": n_total, "local_qubits": n_local, "gpus": world, "gates": len(ops), "ms_per_step": ms_per_step,
                "shard_gib": (amp_bytes << n_local) / 2**30, "exchanges_per_step": sw["n_swaps_per_step"],
                "gb_sent_per_exchange_per_gpu": sw["gb_sent_per_swap"], "gbs_per_direction": sw["gbs_per_direction"],
                "frac_of_770_measured_peer_copy": sw["frac_of_770_measured_peer_copy"],
                "frac_of_900_nominal": (sw["gbs_per_direction"] / 900.0) if sw["gbs_per_direction"] else None,
                "target": "global-qubit swaps at >= 0.70 of NVLink bandwidth (north_star)",
                "timing": "CUDA events around handshake + k_peer_swap + handshake on the compute stream, read lazily; this rank",
            }
        kernel_ms = ms_total - detail["nvlink_swaps"]["swap_ms_per_step"] * args.steps
        if launches > 0 and kernel_ms > 0:
            roofline["ms_per_launch"] = kernel_ms / launches
            roofline["achieved"] = bytes_per_launch / (roofline["ms_per_launch"] * 1e-3) / 1e9
            roofline["frac"] = roofline["achieved"] / hbm_peak
            roofline["frac_physical"] = roofline["frac"]
            if fused:
                roofline["hbm_actual_gbs"] = sweep_bytes / (roofline["ms_per_launch"] * 1e-3) / 1e9
                roofline["hbm_actual_frac"] = roofline["hbm_actual_gbs"] / hbm_peak
                roofline["frac_physical"] = roofline["hbm_actual_frac"]
            roofline["exchange_note"] = "exchange time (NCCL stream) subtracted from the region before dividing by launches"
    # ---- the same circuit gate by gate (one HBM sweep per gate) and in complex64 ----------------
    if args.sweeps and not distributed and args.fuse:
        def timed_circuit(vec, fuse, reps=2):
            for _ in range(3):
                vec.apply_ops(rec, fuse=fuse)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                vec.apply_ops(rec, fuse=fuse)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        t = timed_circuit(sv, False)
        detail["unfused_gate_by_gate"] = {"ms_per_step": t, "value_gbs": alg_bytes_total / (t * 1e-3) / 1e9,
                                          "frac_of_peak": alg_bytes_total / (t * 1e-3) / 1e9 / hbm_peak,
                                          "kernel": "k_apply_dense / k_apply_diag, HBM-bound"}
        if args.dtype == "c128":
            buf32 = torch.empty((1 << n_local) * 2, dtype=torch.float32, device=dev)
            for s0 in range(0, buf32.numel(), chunk):
                buf32[s0:s0 + chunk].copy_(buf[s0:s0 + chunk])
            sv32 = q.StateVector(n_local, np.complex64, device=local_rank, external_ptr=buf32.data_ptr())
            b32 = circuit_bytes(ops, n_total, 8)
            t = timed_circuit(sv32, True, reps=3)
            l32, _ = sv32.last_apply_stats()
            detail["config2_complex64"] = {"ms_per_step": t, "value_gbs": b32 / (t * 1e-3) / 1e9, "launches_per_step": l32,
                                           "hbm_actual_frac": 2 * 8 * (1 << n_local) * l32 / (t * 1e-3) / 1e9 / hbm_peak}
            t = timed_circuit(sv32, False)
            detail["config2_complex64"]["unfused_ms_per_step"] = t
            detail["config2_complex64"]["unfused_frac_of_peak"] = b32 / (t * 1e-3) / 1e9 / hbm_peak
            del sv32, buf32
    # ---- single-gate sweeps: C2(i), every target wire individually ----------------------------
    if args.sweeps and not distributed:
        detail["single_gate_sweeps"] = single_gate_sweeps(torch, q, sv, n_local, amp_bytes, hbm_peak)

    if args.adjoint and not distributed:
        del sv
        detail["adjoint_config3"] = adjoint_config3(torch, q, args)
        detail["config1_sel20"] = config1_sel20(torch, q, args)
    if args.config4 and not distributed:
        detail["sparse_config4"] = sparse_config4(torch, q, args)
        detail["state_io"] = state_io(torch, q)
    if args.adjoint and not distributed:
        sv = q.StateVector(n_local, cdtype, device=local_rank, external_ptr=buf.data_ptr())

    # ---- end to end through the public API with host buffers ----------------------------------
    e2e = None
    if args.e2e and n_local <= 31:  # a pinned host copy of a 33-qubit shard (128 GiB per rank) is not a sensible input
        e2e = end_to_end(torch, q, sv, buf, ops, n_local, cdtype, tdtype, amp_bytes, alg_bytes_total, args,
                         dist if distributed else None)

    line = None
    if rank == 0:
        cpu = None
        if args.cpu_baseline and not distributed:
            v, cores, desc = cpu_sample(n_local, budget_s=15.0)[:3]
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": desc}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64" if args.dtype == "c128" else "f32", "data": "synthetic",
            "config": {"workload": f"{workload}, {'complex128' if args.dtype == 'c128' else 'complex64'}",
                       "local_qubits": n_local, "global_qubits": n_glob, "fused": bool(args.fuse),
                       "l2": f"state ({(amp_bytes << n_local) / 2**30:.0f} GiB per GPU) >> L2 (126 MB): no flush needed",
                       "bytes_rule": "sum over gates of 2*B*N/2^controls (SURVEY.md 8d), no credit for fusion"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "hbm_sweeps": sweeps, "clocks": clk.summary(), "detail": detail,
        }
        clk_all.__exit__()
        line["clocks"]["region"] = "the timed steps of `value`"
        line["clocks"]["whole_process"] = clk_all.summary()
        comp = roofline.get("compute")
        if comp and "tflops" in comp and (line["clocks"] or {}).get("sm_mhz"):
            comp["peak_tflops"] = comp.pop("peak_tflops_per_ghz") * line["clocks"]["sm_mhz"] / 1e3
            comp["frac"] = comp["tflops"] / comp["peak_tflops"]
            roofline["frac_compute"] = comp["frac"]
        print(json.dumps(line), flush=True)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return line


def dist_self_check(torch, q, DistributedStateVector, dist, args, local_rank, rank, n_glob, cdtype, n_check_local=20):
    """Before anything is timed at N > 1: the bench's circuit generator on n_check_local + log2(N) qubits, applied twice
    (the second application runs on the qubit map the first one left, as the timed steps do) on a sharded register with
    this run's settings, compared shard by shard with the single-GPU engine on the same full state (which the GPU parity
    tests pin to the oracle).  Raises on a mismatch, so an unverified computation is never timed."""
    n = n_check_local + n_glob
    tol = 1e-10 if cdtype == np.complex128 else 1e-5
    ops = workloads.random_gate_circuit(n, args.gates, 2024)
    rec = q.Ops(ops)
    rng = np.random.default_rng(4321)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi = (psi / np.linalg.norm(psi)).astype(cdtype)
    one = q.StateVector(n, cdtype, device=local_rank)
    one.h2d(psi)
    sharded = DistributedStateVector(n, cdtype, device=local_rank)
    lo, hi = rank << n_check_local, (rank + 1) << n_check_local
    sharded.h2d(psi[lo:hi])
    worst = 0.0
    for _ in range(2):
        one.apply_ops(rec, fuse=bool(args.fuse))
        sharded.apply_ops(rec, fuse=bool(args.fuse))
    n_swaps = sharded.swap_stats()[0]
    try:
        n_oop, n_carried = sharded.fused_exchange_stats()
    except Exception:
        n_oop, n_carried = 0, 0
    want = one.d2h()[lo:hi]
    got = sharded.d2h()  # collective (lazy qubit map): every rank calls it
    worst = float(np.max(np.abs(got - want)))
    t = torch.tensor([worst], dtype=torch.float64, device=torch.device("cuda", local_rank))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    worst = float(t)
    del one, sharded
    out = {"n_qubits": n, "local_qubits": n_check_local, "gates": 2 * len(ops), "max_abs_err_vs_single_gpu": worst, "tol": tol,
           "in_place_exchanges": n_swaps, "exchanges_through_second_buffer": n_oop, "carried_by_sweeps": n_carried,
           "passed": bool(worst <= tol)}
    if not out["passed"]:
        raise SystemExit("bench.py: sharded engine differs from the single-GPU engine: " + json.dumps(out))
    return out


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


def adjoint_config3(torch, q, args):
    """BASELINE config 3: 24-qubit hardware-efficient ansatz (4 layers, 192 parameters), 100-term Pauli
    Hamiltonian, adjoint Jacobian; wall seconds through the public API (device sync on both sides)."""
    n = args.adjoint_qubits
    ops, n_par = workloads.hardware_efficient_ansatz(n, layers=4, seed=11)
    words, wires, coeffs = workloads.random_pauli_hamiltonian(n, 100, seed=5)
    ham = q.Observable.from_tuple(workloads.hamiltonian_tuple(words, wires, coeffs))
    rec = q.Ops(ops)
    sv = q.StateVector(n, np.complex128)
    out = {"n_qubits": n, "n_params": n_par, "n_ops": len(ops), "n_terms": 100, "dtype": "complex128"}

    def run():
        sv.set_basis_state(0)
        sv.apply_ops(rec, fuse=True)
        e = sv.expval(ham)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        jac = sv.adjoint_jacobian(rec, [ham], list(range(n_par)))
        torch.cuda.synchronize()
        return e, jac, time.perf_counter() - t0

    run()
    best = None
    for _ in range(3):
        e, jac, dt = run()
        best = dt if best is None else min(best, dt)
    out["jacobian_s"] = best
    out["expval"] = e
    out["jac_norm"] = float(np.linalg.norm(jac))
    # self-check without an oracle at this size: central finite difference on two parameters
    idx_par = [i for i, o in enumerate(ops) if o["params"]]
    fd_err = 0.0
    for p in (0, n_par - 1):
        vals = []
        for sgn in (+1, -1):
            ops2 = [dict(o) for o in ops]
            ops2[idx_par[p]] = dict(ops2[idx_par[p]], params=[ops[idx_par[p]]["params"][0] + sgn * 1e-4])
            sv.set_basis_state(0)
            sv.apply_ops(q.Ops(ops2), fuse=True)
            vals.append(sv.expval(ham))
        fd_err = max(fd_err, abs((vals[0] - vals[1]) / 2e-4 - jac[0, p]))
    out["finite_difference_max_abs_err"] = fd_err
    # algorithmic bytes per SURVEY.md 8d with n_bra = 1: P * 2BN * 2 + Q * BN * 2
    B, N = 16, 1 << n
    alg = len(ops) * 2 * B * N * 2 + n_par * B * N * 2
    out["algorithmic_gb"] = alg / 1e9
    out["algorithmic_gbs"] = alg / best / 1e9
    if args.cpu_baseline:
        try:
            from oracle import lq_port as lq

            lq.set_num_threads(os.cpu_count() or 1)
            st = lq.LQState(n)
            st.apply_ops(ops)
            ham_t = workloads.hamiltonian_tuple(words, wires, coeffs)
            t0 = time.perf_counter()

            # Hamiltonian bra through the Pauli-sum kernel of the port (one bra, as the GPU path does)
            lam = st.copy()
            bra = lam.apply_pauli_hamiltonian(words, wires, coeffs)
            jac_cpu = _cpu_adjoint(lq, lam, bra, ops, n_par)
            out["cpu_port_jacobian_s"] = time.perf_counter() - t0
            out["cpu_cores"] = os.cpu_count()
            out["gpu_vs_cpu_port_max_abs_diff"] = float(np.max(np.abs(jac_cpu - jac[0])))
        except Exception as ex:  # the CPU leg must never break the GPU numbers
            out["cpu_port_error"] = repr(ex)
    return out


def config1_sel20(torch, q, args):
    """BASELINE config 1: 20-qubit StronglyEntanglingLayers (2 layers, Rot expanded: 160 ops, 120 parameters),
    expval(Z0) + adjoint Jacobian, complex128 -- the reference's own CPU-runnable case, timed on the GPU and on the
    CPU port with all host cores and with one thread."""
    n = 20
    ops, n_par = workloads.strongly_entangling_layers(n, 2, 1337)
    rec = q.Ops(ops)
    z0 = q.Observable.named("PauliZ", [0])
    sv = q.StateVector(n, np.complex128)
    out = {"n_qubits": n, "n_params": n_par, "n_ops": len(ops), "dtype": "complex128"}

    def run():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sv.set_basis_state(0)
        sv.apply_ops(rec, fuse=True)
        e = sv.expval(z0)
        jac = sv.adjoint_jacobian(rec, [z0], list(range(n_par)))
        torch.cuda.synchronize()
        return e, jac, time.perf_counter() - t0

    run()
    best = None
    for _ in range(5):
        e, jac, dt = run()
        best = dt if best is None else min(best, dt)
    out.update({"gpu_circuit_expval_jacobian_s": best, "expval": e, "jac_norm": float(np.linalg.norm(jac))})
    if args.cpu_baseline:
        try:
            from oracle import lq_port as lq

            for threads in (os.cpu_count() or 1, 1):
                lq.set_num_threads(threads)
                t0 = time.perf_counter()
                st = lq.LQState(n)
                st.apply_ops(ops)
                jac_cpu = lq.adjoint_jacobian(st, ops, [("Named", "PauliZ", [0])], list(range(n_par)))
                out[f"cpu_port_{threads}_threads_s"] = time.perf_counter() - t0
            out["gpu_vs_cpu_port_max_abs_diff"] = float(np.max(np.abs(np.asarray(jac_cpu).reshape(-1) - jac[0])))
            lq.set_num_threads(os.cpu_count() or 1)
        except Exception as ex:  # the CPU leg must never break the GPU numbers
            out["cpu_port_error"] = repr(ex)
    return out


def _cpu_adjoint(lq, lam, bra, ops, n_par):
    """lightning.qubit-style reverse sweep with one bra (oracle/lq_port.py's loop, inlined for one bra)."""
    from oracle import np_oracle as orc

    jac = np.zeros(n_par)
    cur = n_par - 1
    for op in reversed(ops):
        params = op.get("params", ())
        mu = lam.copy() if len(params) else None
        lam.apply_op(op["name"], op["wires"], params, True, op.get("matrix"))
        if len(params):
            g, s = orc.generator(op["name"], len(op["wires"]))
            mu.apply_matrix(g, op["wires"])
            jac[cur] = -2.0 * s * bra.inner(mu).imag
            cur -= 1
        bra.apply_op(op["name"], op["wires"], params, True, op.get("matrix"))
    return jac


def sparse_config4(torch, q, args):
    """BASELINE config 4: 22-qubit molecular-style sparse Hamiltonian, <H> by CSR SpMV fused with the dot."""
    n = args.config4_qubits
    t0 = time.perf_counter()
    m, (words, wires, coeffs) = workloads.molecular_style_sparse_hamiltonian(n, 400, 30, seed=3)
    build_s = time.perf_counter() - t0
    ops, _ = workloads.hardware_efficient_ansatz(n, layers=4, seed=11)
    sv = q.StateVector(n, np.complex128)
    sv.apply_ops(q.Ops(ops), fuse=True)
    obs = q.Observable.sparse(m.indptr, m.indices, m.data)
    e_csr = sv.expval_csr(m.indptr, m.indices, m.data)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e_csr = sv.expval_csr(m.indptr, m.indices, m.data)
    torch.cuda.synchronize()
    t_csr = time.perf_counter() - t0
    # the same matrix as a SparseHamiltonian observable: its CSR arrays stay device-resident after the first use
    e_obs = sv.expval(obs)
    torch.cuda.synchronize()
    t_obs = None
    for _ in range(5):
        t0 = time.perf_counter()
        e_obs = sv.expval(obs)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t_obs = dt if t_obs is None else min(t_obs, dt)
    t0 = time.perf_counter()
    e_pw = sv.expval_pauli_words(words, wires, coeffs)
    torch.cuda.synchronize()
    t_pw = time.perf_counter() - t0
    B, I, N = 16, 8, 1 << n
    alg = m.nnz * (B + I) + (N + 1) * I + 2 * B * N
    hbm_peak, _ = measured_peaks()
    return {"n_qubits": n, "nnz": int(m.nnz), "nnz_per_row": m.nnz / N, "csr_gb": m.nnz * 24 / 1e9, "host_build_s": build_s,
            "expval_csr": e_csr, "expval_pauli_words": e_pw, "abs_diff": abs(e_csr - e_pw),
            "csr_call_s_including_h2d": t_csr, "csr_observable_call_s_device_resident": t_obs,
            "csr_observable_gbs": alg / t_obs / 1e9, "csr_observable_frac_of_hbm_peak": alg / t_obs / 1e9 / hbm_peak,
            "expval_csr_observable": e_obs, "pauli_words_call_s": t_pw, "algorithmic_gb": alg / 1e9,
            "bytes_rule": "nnz * (16 + 8) + (N + 1) * 8 + 2 * 16 * N (SURVEY.md 8d), CSR arrays device-resident, wall time of "
                          "the call incl. the read-back of the result"}


def state_io(torch, q, n=28):
    """Whole-state host <-> device copies (lightning_gpu.py:345-381 syncH2D / syncD2H): pageable NumPy memory through the
    multi-threaded pinned staging of csrc/state_io.cu, and pinned memory straight over PCIe."""
    sv = q.StateVector(n, np.complex128)
    host = np.empty(1 << n, dtype=np.complex128)
    host.view(np.float64)[:] = 0.0
    pinned = torch.empty(1 << n, dtype=torch.complex128).pin_memory().numpy()
    gb = host.nbytes / 1e9

    def best(fn):
        fn()
        out = None
        for _ in range(2):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            out = dt if out is None else min(out, dt)
        return gb / out

    return {"n_qubits": n, "gb": gb, "host_threads": min(16, max(1, (os.cpu_count() or 2) // 2)),
            "pageable_h2d_gbs": best(lambda: sv.h2d(host)), "pageable_d2h_gbs": best(lambda: sv.d2h(host)),
            "pinned_h2d_gbs": best(lambda: sv.h2d(pinned)), "pinned_d2h_gbs": best(lambda: sv.d2h(pinned)),
            "plain_cudaMemcpy_of_pageable_gbs": "10.5 H2D / 18.3 D2H on this host (profiles/r2_state_io.txt)"}


def end_to_end(torch, q, sv, buf, ops, n, cdtype, tdtype, amp_bytes, alg_bytes, args, dist=None):
    """StatePrep(host state) -> circuit -> <Z0> on the host, all through the public API.  On a sharded register every
    rank uploads its own shard from its own pinned buffer (CopyHostDataToGpu of StateVectorCudaMPI), the expectation
    value is all-reduced; the time is the slowest rank's."""
    world = dist.get_world_size() if dist is not None else 1
    ok = 1
    try:
        host = torch.empty(1 << n, dtype=tdtype, pin_memory=True)
    except RuntimeError:
        ok = 0
    if dist is not None:  # all ranks or none: a rank that dropped out would leave the others in a collective
        flag = torch.tensor([ok], dtype=torch.int32, device=buf.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        ok = int(flag)
    if not ok:
        return None
    host.copy_(torch.view_as_complex(buf.view(-1, 2)))
    torch.cuda.synchronize()
    host_np = host.numpy()
    z0 = q.Observable.named("PauliZ", [0])
    h2d = world * host_np.nbytes + sum(16 * len(np.atleast_1d(op.get("params", ()))) +
                                       (op["matrix"].nbytes if "matrix" in op else 0) for op in ops)

    def sync():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    api = "ctypes DistributedStateVector (every rank uploads its own shard)"
    if dist is None:
        # the plugin surface: the device class with the reference's call pattern (lightning_gpu.py:392-447 StatePrep ->
        # syncH2D -> HostToDevice; apply; expval) on top of the pybind11 module, which calls the C ABI
        from pennylane_lightning_gpu_b200.lightning_gpu import LightningGPU, Obs, Op

        device = LightningGPU(n, c_dtype=cdtype)
        z0_obs = Obs("PauliZ", [0])
        api = "LightningGPU device (pennylane_lightning_gpu_b200/lightning_gpu.py) -> lightning_gpu_qubit_ops (pybind11) -> C ABI"

        def step():
            recs = [Op("StatePrep", list(range(n)), [host_np])]
            recs += [Op(op["name"], op["wires"], op.get("params", ()), False, op.get("matrix")) for op in ops]
            device.apply(recs)               # StatePrep = HostToDevice of the 2^n amplitudes, then the circuit
            return device.expval(z0_obs)     # one double back to the host
    else:
        def step():
            sv.h2d(host_np)                  # HostToDevice of this rank's shard
            rec = q.Ops(ops)                 # host op list -> recorded circuit
            sv.apply_ops(rec, fuse=bool(args.fuse))
            return sv.expval(z0)             # one double back to the host

    step()
    k = max(1, min(args.steps, 3))
    sync()
    t0 = time.perf_counter()
    for _ in range(k):
        val = step()
    sync()
    dt = (time.perf_counter() - t0) / k
    if dist is not None:
        t = torch.tensor([dt], dtype=torch.float64, device=buf.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t)
    return {"value": alg_bytes / dt / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 8 * world,
            "ms_per_step": dt * 1e3, "result_check": float(val), "api": api,
            "h2d_gbs_if_alone": None, "note": "pinned host state; the 2^n-amplitude upload is PCIe-bound (~55 GB/s measured on "
            "this host, detail.state_io), the circuit itself is `ms_per_step` of the device-resident line"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--qubits", type=int, default=30, help="local qubits per GPU")
    ap.add_argument("--gates", type=int, default=200)
    ap.add_argument("--circuit", default="auto", choices=["auto", "config2", "config5"],
                    help="auto: config 2(ii) random gate circuit; config 5 random-layer circuit from 32 local qubits up")
    ap.add_argument("--dtype", default="c128", choices=["c128", "c64"])
    ap.add_argument("--fuse", type=int, default=1)
    ap.add_argument("--sweeps", type=int, default=1)
    ap.add_argument("--e2e", type=int, default=1)
    ap.add_argument("--cpu-baseline", dest="cpu_baseline", type=int, default=1)
    ap.add_argument("--adjoint", type=int, default=1, help="also time BASELINE config 3 (24q adjoint Jacobian)")
    ap.add_argument("--adjoint-qubits", dest="adjoint_qubits", type=int, default=24)
    ap.add_argument("--config4", type=int, default=1, help="also run BASELINE config 4 (22q sparse Hamiltonian)")
    ap.add_argument("--config4-qubits", dest="config4_qubits", type=int, default=22)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
