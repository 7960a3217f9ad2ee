"""Builds the TEST-ONLY CPU emulator of the fused register-tile kernel (tests/native/regs_emu.cu) into
tests/native/_build/libregs_emu.so, linked against the product library for the host-side planner it checks."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(HERE, "_build", "libregs_emu.so")


def build() -> str:
    from pennylane_lightning_gpu_b200 import _build

    lib = _build.build_lib()
    src = os.path.join(HERE, "regs_emu.cu")
    deps = [src, lib] + [os.path.join(_build.CSRC, f) for f in os.listdir(_build.CSRC) if f.endswith((".h", ".cuh"))]
    if os.path.exists(OUT) and all(os.path.getmtime(d) <= os.path.getmtime(OUT) for d in deps):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nccl_inc, _ = _build._nccl_dirs()
    cmd = [_build.NVCC, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-shared",
           "-gencode", "arch=compute_100a,code=sm_100a"] + (["-I", nccl_inc] if nccl_inc else []) + \
          [src, "-o", OUT, "-L", _build.LIBDIR, "-lqsv_b200", "-Xlinker", "-rpath=" + _build.LIBDIR]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator build failed:\n" + r.stdout + r.stderr)
    return OUT
