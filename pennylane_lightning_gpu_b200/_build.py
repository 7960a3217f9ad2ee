"""In-tree build of the native pieces (no cmake, no JIT cache):

* ``lib/libqsv_b200.so``            -- CUDA kernels + C ABI (include/qsv_b200.h), nvcc, sm_100a only
* ``lightning_gpu_qubit_ops*.so``   -- pybind11 module mirroring bindings/Bindings.cpp of the
                                        reference, g++ only, links libqsv_b200.so

Objects are rebuilt only when a source or header is newer; translation units compile in parallel.
"""
from __future__ import annotations

import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
SRC = os.path.join(PKG, "src")
LIBDIR = os.path.join(PKG, "lib")
OBJDIR = os.path.join(PKG, "build")
LIB = os.path.join(LIBDIR, "libqsv_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
]


def _nccl_dirs() -> tuple[str | None, str | None]:
    """(include dir, lib dir) of the NCCL that PyTorch itself loads (the `nvidia-nccl` wheel): linking the same
    libnccl.so.2 keeps one NCCL in the process whichever of torch / libqsv_b200 is loaded first; the system libnccl
    is the fallback."""
    try:
        import importlib.util

        spec = importlib.util.find_spec("nvidia.nccl")
        for base in (spec.submodule_search_locations if spec else []):
            lib, inc = os.path.join(base, "lib"), os.path.join(base, "include")
            if os.path.exists(os.path.join(lib, "libnccl.so.2")) and os.path.exists(os.path.join(inc, "nccl.h")):
                return inc, lib
    except Exception:
        pass
    return None, None


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    out = (r.stdout + r.stderr).strip()
    if out:
        print(out, file=sys.stderr)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "qsv_b200.h"))
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs = []
    objs = []
    nccl_inc, nccl_lib = _nccl_dirs()
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s[:-3] + ".o")
        objs.append(obj)
        if force or _newer(obj, [src] + headers):
            cmd = [NVCC] + NVCC_FLAGS + (["-I", nccl_inc] if nccl_inc else []) + \
                (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)
    if jobs:
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
            list(ex.map(_run, jobs))
    if jobs or not os.path.exists(LIB):
        link = ["-L", nccl_lib, "-l:libnccl.so.2", "-Xlinker", "-rpath=" + nccl_lib] if nccl_lib else ["-lnccl"]
        _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs + link)
    return LIB


def build_lib_variant(name: str, extra_flags: list[str]) -> str:
    """A/B build of the same library with extra nvcc flags (e.g. -DQSV_PLAIN_SMEM) into lib/variants/libqsv_<name>.so;
    select it at run time with QSV_LIB_PATH.  The default library is left untouched."""
    out_dir = os.path.join(LIBDIR, "variants")
    obj_dir = os.path.join(OBJDIR, "variant_" + name)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)
    out = os.path.join(out_dir, f"libqsv_{name}.so")
    nccl_inc, nccl_lib = _nccl_dirs()
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    jobs, objs = [], []
    for s in sources:
        obj = os.path.join(obj_dir, s[:-3] + ".o")
        objs.append(obj)
        jobs.append([NVCC] + NVCC_FLAGS + list(extra_flags) + (["-I", nccl_inc] if nccl_inc else []) +
                    ["-c", os.path.join(CSRC, s), "-o", obj])
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 4)) as ex:
        list(ex.map(_run, jobs))
    link = ["-L", nccl_lib, "-l:libnccl.so.2", "-Xlinker", "-rpath=" + nccl_lib] if nccl_lib else ["-lnccl"]
    _run([NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", out] + objs + link)
    return out


def pybind_module_path() -> str:
    return os.path.join(PKG, "lightning_gpu_qubit_ops" + sysconfig.get_config_var("EXT_SUFFIX"))


def build_pybind(force: bool = False) -> str | None:
    """pybind11 module with the reference's class names, calling only the C ABI."""
    src = os.path.join(SRC, "bindings", "Bindings.cpp")
    if not os.path.exists(src):
        return None
    import pybind11

    out = pybind_module_path()
    deps = [src, os.path.join(ROOT, "include", "qsv_b200.h")]
    for d, _, files in os.walk(SRC):
        deps += [os.path.join(d, f) for f in files if f.endswith((".hpp", ".h"))]
    if force or _newer(out, deps) or _newer(out, [LIB]):
        cmd = [
            "g++", "-O2", "-std=c++20", "-shared", "-fPIC", "-fvisibility=hidden",
            "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"],
            "-I", os.path.join(ROOT, "include"), "-I", os.path.join(SRC, "simulator"),
            "-I", os.path.join(SRC, "algorithms"), "-I", os.path.join(SRC, "util"),
            src, "-o", out, "-L", LIBDIR, "-lqsv_b200", "-Wl,-rpath,$ORIGIN/lib",
        ]
        _run(cmd)
    return out


def build_all(force: bool = False) -> None:
    build_lib(force)
    build_pybind(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv)
    print(LIB)
