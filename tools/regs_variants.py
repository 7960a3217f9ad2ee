"""Build experimental variants of csrc/tile_regs.cu into lib/variants/libqsv_<name>.so (the default library is left
untouched); run one with QSV_LIB_PATH=... python bench.py.  Used for the A/B runs recorded in profiles/."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pennylane_lightning_gpu_b200 import _build  # noqa: E402

SRC = os.path.join(ROOT, "pennylane_lightning_gpu_b200", "csrc", "tile_regs.cu")

D1_K = '''template <typename T, int B, int KIND, bool PRED, typename A>
__device__ __forceinline__ void reg_d1_k(A (&x)[RT_NS], const T *mp, uint32_t creg) {
    const A q0 = reinterpret_cast<const A *>(mp)[0], q1 = reinterpret_cast<const A *>(mp)[1];
    const A q2 = reinterpret_cast<const A *>(mp)[2], q3 = reinterpret_cast<const A *>(mp)[3];
#pragma unroll
    for (int j = 0; j < RT_NS; ++j) {
        if ((j >> B) & 1) continue;
        if (!PRED || (j & creg) == creg) {
            const A a = x[j], b = x[j | (1 << B)];
            if (KIND == RG_D1_SWAP) {
                x[j] = b;
                x[j | (1 << B)] = a;
            } else if (KIND == RG_D1_REAL) {
                x[j].x = q0.x * a.x + q1.x * b.x;
                x[j].y = q0.x * a.y + q1.x * b.y;
                x[j | (1 << B)].x = q2.x * a.x + q3.x * b.x;
                x[j | (1 << B)].y = q2.x * a.y + q3.x * b.y;
            } else if (KIND == RG_D1_RX) {
                x[j].x = q0.x * a.x - q1.y * b.y;
                x[j].y = q0.x * a.y + q1.y * b.x;
                x[j | (1 << B)].x = q3.x * b.x - q2.y * a.y;
                x[j | (1 << B)].y = q3.x * b.y + q2.y * a.x;
            } else {
                x[j].x = q0.x * a.x - q0.y * a.y + q1.x * b.x - q1.y * b.y;
                x[j].y = q0.x * a.y + q0.y * a.x + q1.x * b.y + q1.y * b.x;
                x[j | (1 << B)].x = q2.x * a.x - q2.y * a.y + q3.x * b.x - q3.y * b.y;
                x[j | (1 << B)].y = q2.x * a.y + q2.y * a.x + q3.x * b.y + q3.y * b.x;
            }
        }
    }
}

template <typename T, int B, typename A>
__device__ __forceinline__ void reg_d1(A (&x)[RT_NS], int kind, const T *mp, uint32_t creg) {
    if (creg == 0) {
        switch (kind) {
        case RG_D1_SWAP: reg_d1_k<T, B, RG_D1_SWAP, false>(x, mp, 0); break;
        case RG_D1_REAL: reg_d1_k<T, B, RG_D1_REAL, false>(x, mp, 0); break;
        case RG_D1_RX: reg_d1_k<T, B, RG_D1_RX, false>(x, mp, 0); break;
        default: reg_d1_k<T, B, RG_D1, false>(x, mp, 0); break;
        }
    } else {
        switch (kind) {
        case RG_D1_SWAP: reg_d1_k<T, B, RG_D1_SWAP, true>(x, mp, creg); break;
        default: reg_d1_k<T, B, RG_D1, true>(x, mp, creg); break;
        }
    }
}

'''


def v4(s):
    """unpredicated D1 fast path when no control sits on a register bit"""
    a = s.index("template <typename T, int B, typename A>\n__device__ __forceinline__ void reg_d1(A (&x)[RT_NS], int kind, const T *mp, uint32_t creg) {")
    b = s.index("// 4x4 block on register bits BA > BB; matrix index = 2 * bit(BA) + bit(BB)")
    return s[:a] + D1_K + s[b:]


def v6(s):
    """v4 + thread-level control as a real branch hoisted in front of every gate (no per-slot thr_on predicate)"""
    s = v4(s)
    old = "            const T *mp = spool + g.mat_off;\n            const uint32_t creg = g.ctrl_reg;\n"
    assert old in s
    s = s.replace(old, old + "            if (__ballot_sync(0xffffffffu, thr_on) == 0u) continue;  // warp-uniform skip\n", 1)
    return s


def v8(s):
    """v4 with 2 CTAs/SM but 512 threads?  no: only lifts the min-blocks hint for double so ptxas may use 168 registers
    (1 CTA of 256 threads per SM would halve occupancy) -- kept as a control experiment"""
    s = v4(s)
    return s.replace("constexpr int MINB = sizeof(T) == 8 ? 2 : 3;", "constexpr int MINB = sizeof(T) == 8 ? 1 : 3;")


VARIANTS = {"v4": v4, "v6": v6, "v8": v8}


def main():
    names = sys.argv[1:] or list(VARIANTS)
    out_dir = os.path.join(ROOT, "pennylane_lightning_gpu_b200", "lib", "variants")
    os.makedirs(out_dir, exist_ok=True)
    orig = open(SRC).read()
    backup = _build.LIB + ".orig"
    shutil.copy(_build.LIB, backup)
    try:
        for n in names:
            open(SRC, "w").write(VARIANTS[n](orig))
            _build.build_lib()
            shutil.copy(_build.LIB, os.path.join(out_dir, f"libqsv_{n}.so"))
            print("built", n)
    finally:
        open(SRC, "w").write(orig)
        shutil.copy(backup, _build.LIB)
        os.remove(backup)
        # make the object file of the restored source current again
        subprocess.run([sys.executable, "-c", "from pennylane_lightning_gpu_b200 import _build; _build.build_lib()"], cwd=ROOT)


if __name__ == "__main__":
    main()
