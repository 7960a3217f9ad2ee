// Program format and per-thread arithmetic of the register-blocked fused tile kernel (tile_regs.cu).
//
// Everything here is __host__ __device__: the kernel and the test-only CPU emulator of the kernel
// (tests/native/regs_emu.cu, never part of the product library) run the very same per-thread code on the very
// same RegProgram, so the host-side planner / program builder can be checked against the oracle without a GPU.
#pragma once

#include <vector>

#include "device_utils.cuh"

namespace qsv {
namespace rt {

constexpr int TB = 12;       // tile bits: a CTA owns 2^12 amplitudes
constexpr int RB_MAX = 4;    // register bits (QSV_REGS_RB=3: 8 amplitudes per thread, 512 threads per CTA)
constexpr int NTB_MAX = 9;   // thread bits, at most
constexpr int MAX_GATES = 48;
constexpr int MAX_PASSES = 48;
constexpr int POOL = 1280;   // doubles

enum : unsigned char { RG_D1 = 1, RG_D1_REAL = 2, RG_D1_RX = 3, RG_D1_SWAP = 4, RG_D2 = 5, RG_DIAG = 6, RG_D1_DIAG = 7 };

struct RegGate {
    unsigned char kind;
    unsigned char ra, rb;        // register-bit numbers (D1: ra; D2: ra = matrix MSB > rb; DIAG: rb = #table bits)
    unsigned char ctrl_reg;      // controls among the register bits (mask over the slot number)
    unsigned char reg_mask[2];   // DIAG: table bit b = parity(slot & reg_mask[b]) ^ parity(tid & thr_mask[b]) ^ ...
    unsigned short ctrl_thr;     // controls among the thread bits (mask over threadIdx.x)
    unsigned short thr_mask[2];
    unsigned short mat_off;      // first double of this gate in the pool
    unsigned short pad0;
    uint64_t out_ctrl;           // controls outside the tile (global bit positions)
    uint64_t out_mask[2];        // ... ^ parity(outside & out_mask[b]); table index = 2 * bit[0] + bit[1]
};

// Addresses are GF(2)-affine in (thread bits, register bits): offset = c ^ XOR_i tid_i * thr[i] ^ XOR_b slot_b * reg[b].
// Shared-memory offsets are in elements and already XOR-swizzled (the swizzle is linear); index permutations
// (PauliX, CNOT, SWAP between passes) are folded into the columns on the host, so they cost nothing here.
struct RegPass {
    unsigned short ld_thr[NTB_MAX], ld_reg[RB_MAX], ld_c;  // load at the start of the pass (passes > 0)
    unsigned short st_thr[NTB_MAX], st_reg[RB_MAX], st_c;  // store at its end (all but the last pass)
    unsigned short gate_begin;  // [gate_begin, udiag_end): thread-uniform diagonal gates, merged into one phase
    unsigned short udiag_end;   // [udiag_end, gate_end): everything else, in order
    unsigned short gate_end;
    unsigned short mma_off;     // != NO_MMA: pool offset of the real 8x8 form of a 4x4 gate on thread bits 1 (MSB) and 0,
                                // applied first in the pass through FP64 tensor-core MMA (complex128 kernels only)
};
constexpr unsigned short NO_MMA = 0xffff;

struct GlobalMap {  // element offsets inside the shard (tile bits only)
    uint64_t thr[NTB_MAX], reg[RB_MAX], c;
};

struct RegProgram {
    int n_passes;
    int n_gates;
    int pool_used;
    int prefetch;                     // > 0: CTA b prefetches the tile of CTA b + prefetch into L2
    uint64_t index_hi;                // value of the index bits above the local shard
    Holes tile_holes;                 // all tile bits, ascending (expands blockIdx.x to the tile base)
    GlobalMap gl_load, gl_store;      // first pass load / last pass store
    RegPass passes[MAX_PASSES];
    RegGate gates[MAX_GATES];
    double pool[POOL];                // gate constants (complex128 kernels read them from here or from shared memory)
    float poolf[POOL];                // the same in single precision for the complex64 kernels
    int uniform_consts;               // != 0: uncontrolled dense gates read their matrix straight from this struct
    int n_folded;                     // host-side statistics: gates folded into pass boundaries ...
    int n_mma_gates;                  // ... and gates multiplied into tensor-core blocks
    unsigned stagger_ns;              // > 0: the second CTA to arrive on each SM waits this long once (see k_tile_regs)
    unsigned epoch;                   // launch number (tags the per-SM arrival counters)
    int pad1;
};

// Sweep fused with a global<->local index-bit exchange (k_tile_regs<..., XCHG = true>, csrc/dist.cu): an amplitude whose
// shard offset has the exchanged bit equal to this rank's value of the global bit (`keep` = bit_mask or 0) stays on this
// rank, at the same offset of its other buffer; every other amplitude goes to the partner's other buffer with that bit
// flipped.  The rule is applied per stored element, so the exchanged bit may be any bit of the shard: outside the tile
// (whole tiles go one way), a thread bit or a register bit of the last pass (a tile is split between the two targets).
struct XchgTarget {
    bool stays;
    uint64_t base;
};
// `stash_mask` != 0 splits the exchange between the sweep before it and the sweep after it (dist.cu): of the amplitudes that
// leave, only those with the stash bit CLEAR are pushed to the partner now; the others are parked at their own offset of this
// rank's other buffer, where the partner's next sweep fetches them (xchg_source) -- each of the two sweeps then carries a
// quarter of the shard over NVLink instead of one sweep carrying half of it at the link's full rate.
__host__ __device__ __forceinline__ XchgTarget xchg_target(uint64_t offset, uint64_t bit_mask, uint64_t keep,
                                                           uint64_t stash_mask = 0) {
    const bool stays = (offset & bit_mask) == keep || (offset & stash_mask) != 0;
    return {stays, stays ? offset : offset ^ bit_mask};
}
// Pull side: where the amplitude of the NEW layout at `offset` is read from -- this rank's buffer (it stayed, or the partner
// pushed it), or the partner's buffer at the offset with the exchanged bit flipped (the partner parked it there).
__host__ __device__ __forceinline__ XchgTarget xchg_source(uint64_t offset, uint64_t bit_mask, uint64_t keep,
                                                           uint64_t stash_mask) {
    const bool local = (offset & bit_mask) == keep || (offset & stash_mask) == 0;
    return {local, local ? offset : offset ^ bit_mask};
}

template <typename T> __host__ __device__ __forceinline__ const T *const_pool(const RegProgram &P);
// where the unpredicated dense paths read their matrix from: the kernel-parameter constant bank (complex128) or the
// shared-memory copy (complex64; -DQSV_PLAIN_SMEM forces it for both, an A/B build)
template <typename T> __host__ __device__ __forceinline__ const T *plain_consts(const RegProgram &P, const T *spool);
template <> __host__ __device__ __forceinline__ const double *const_pool<double>(const RegProgram &P) { return P.pool; }
template <> __host__ __device__ __forceinline__ const float *const_pool<float>(const RegProgram &P) { return P.poolf; }
template <typename T> __host__ __device__ __forceinline__ const T *plain_consts(const RegProgram &P, const T *spool) {
    // measured on B200 (profiles/r1_ab_fold.txt): equal for complex128, shared memory 4 % faster for complex64
#ifdef QSV_PLAIN_SMEM
    (void)P;
    return spool;
#else
    if constexpr (sizeof(T) == 4) {
        (void)P;
        return spool;
    } else {
        (void)spool;
        return const_pool<T>(P);
    }
#endif
}

template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; static constexpr int SW = 3; };
template <> struct Cx<float> { using type = float2; static constexpr int SW = 4; };

// XOR swizzle of the shared-memory tile: the low SW bits (one 128-byte line) are XORed with every higher
// SW-bit field, so that lanes differing in any bits with distinct positions mod SW hit distinct banks.
__host__ __device__ __forceinline__ uint32_t swz(uint32_t e, int sw) {
    const uint32_t m = (1u << sw) - 1u;
    uint32_t r = e;
    for (int s = sw; s < TB; s += sw) r ^= (e >> s) & m;
    return r;
}

__host__ __device__ __forceinline__ int popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
__host__ __device__ __forceinline__ int popc64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}

template <int NTB> __host__ __device__ __forceinline__ uint32_t thread_offset(const unsigned short *cols, unsigned short c, uint32_t tid) {
    uint32_t r = c;
#pragma unroll
    for (int i = 0; i < NTB; ++i) r ^= (0u - ((tid >> i) & 1u)) & (uint32_t)cols[i];
    return r;
}
template <int NTB> __host__ __device__ __forceinline__ uint64_t thread_offset64(const uint64_t *cols, uint64_t c, uint32_t tid) {
    uint64_t r = c;
#pragma unroll
    for (int i = 0; i < NTB; ++i) r ^= (0ull - (uint64_t)((tid >> i) & 1u)) & cols[i];
    return r;
}
// offset of slot j: XOR of the register-bit columns selected by the (compile-time) bits of j
template <int RB, typename U> __host__ __device__ __forceinline__ U slot_offset(U base, const U (&col)[RB_MAX], int j) {
    U o = base;
#pragma unroll
    for (int b = 0; b < RB; ++b)
        if ((j >> b) & 1) o ^= col[b];
    return o;
}

// ---- gates on register-resident amplitudes --------------------------------------------------------------
// Controls among the register bits are a per-slot predicate.  (A variant with separate unpredicated code paths
// plus a scheduling constraint that keeps controls out of the register bits was measured slower on B200:
// 196 ms vs 166 ms for the config-2 circuit -- more code, 128 registers with spills.)
template <typename T, int B, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d1(A (&x)[NS], int kind, const T *mp, uint32_t creg, int tb1) {
    const A q0 = reinterpret_cast<const A *>(mp)[0], q1 = reinterpret_cast<const A *>(mp)[1];
    const A q2 = reinterpret_cast<const A *>(mp)[2], q3 = reinterpret_cast<const A *>(mp)[3];
    if (kind == RG_D1_DIAG) {
        // one-bit phase table whose only register bit is B (RZ, PhaseShift, CRZ, IsingZZ / MultiRZ with the other qubits on
        // thread bits): the slot's phase is known at compile time, the thread part of the parity swaps the two phases
        const A p0 = tb1 ? q1 : q0, p1 = tb1 ? q0 : q1;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            if ((j & creg) == creg) {
                const A a = x[j], b = x[j | (1 << B)];
                x[j].x = p0.x * a.x - p0.y * a.y;
                x[j].y = p0.x * a.y + p0.y * a.x;
                x[j | (1 << B)].x = p1.x * b.x - p1.y * b.y;
                x[j | (1 << B)].y = p1.x * b.y + p1.y * b.x;
            }
        }
    } else if (kind == RG_D1_SWAP) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            if ((j & creg) == creg) {
                const A t = x[j];
                x[j] = x[j | (1 << B)];
                x[j | (1 << B)] = t;
            }
        }
    } else if (kind == RG_D1_REAL) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            if ((j & creg) == creg) {
                const A a = x[j], b = x[j | (1 << B)];
                x[j].x = q0.x * a.x + q1.x * b.x;
                x[j].y = q0.x * a.y + q1.x * b.y;
                x[j | (1 << B)].x = q2.x * a.x + q3.x * b.x;
                x[j | (1 << B)].y = q2.x * a.y + q3.x * b.y;
            }
        }
    } else if (kind == RG_D1_RX) {
        // real diagonal, imaginary off-diagonal (RX and products of RX)
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            if ((j & creg) == creg) {
                const A a = x[j], b = x[j | (1 << B)];
                x[j].x = q0.x * a.x - q1.y * b.y;
                x[j].y = q0.x * a.y + q1.y * b.x;
                x[j | (1 << B)].x = q3.x * b.x - q2.y * a.y;
                x[j | (1 << B)].y = q3.x * b.y + q2.y * a.x;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            if ((j & creg) == creg) {
                const A a = x[j], b = x[j | (1 << B)];
                x[j].x = q0.x * a.x - q0.y * a.y + q1.x * b.x - q1.y * b.y;
                x[j].y = q0.x * a.y + q0.y * a.x + q1.x * b.y + q1.y * b.x;
                x[j | (1 << B)].x = q2.x * a.x - q2.y * a.y + q3.x * b.x - q3.y * b.y;
                x[j | (1 << B)].y = q2.x * a.y + q2.y * a.x + q3.x * b.y + q3.y * b.x;
            }
        }
    }
}

// Uncontrolled dense gates: no predicates, and the matrix is read from the kernel-parameter constant bank with a
// CTA-uniform index -- in convergent code the compiler keeps it in UNIFORM registers (LDCU) and feeds DFMA / FFMA from
// there, so the constants cost neither vector registers nor shared-memory round trips.
// Every output's LAST multiply-add reads the very input it replaces (new a.x ends with q0x * a.x, ...): all other uses of
// that input come earlier, so the result can be written over it in place and the unrolled code needs no register moves
// to bring the amplitudes back to their home registers before the next gate.
template <typename T, int B, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d1_plain_inplace(A (&x)[NS], int kind, const T *cp) {
    const T q0x = cp[0], q0y = cp[1], q1x = cp[2], q1y = cp[3], q2x = cp[4], q2y = cp[5], q3x = cp[6], q3y = cp[7];
    if (kind == RG_D1_REAL) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            const int k = j | (1 << B);
            const T t0 = q1x * x[k].x, t1 = q1x * x[k].y, t2 = q2x * x[j].x, t3 = q2x * x[j].y;
            x[j].x = q0x * x[j].x + t0;
            x[j].y = q0x * x[j].y + t1;
            x[k].x = q3x * x[k].x + t2;
            x[k].y = q3x * x[k].y + t3;
        }
    } else if (kind == RG_D1_RX) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            const int k = j | (1 << B);
            const T t0 = -(q1y * x[k].y), t1 = q1y * x[k].x, t2 = -(q2y * x[j].y), t3 = q2y * x[j].x;
            x[j].x = q0x * x[j].x + t0;
            x[j].y = q0x * x[j].y + t1;
            x[k].x = q3x * x[k].x + t2;
            x[k].y = q3x * x[k].y + t3;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            const int k = j | (1 << B);
            const T t0 = q1x * x[k].x - q1y * x[k].y - q0y * x[j].y;
            const T t1 = q1x * x[k].y + q1y * x[k].x + q0y * x[j].x;
            const T t2 = q2x * x[j].x - q2y * x[j].y - q3y * x[k].y;
            const T t3 = q2x * x[j].y + q2y * x[j].x + q3y * x[k].x;
            x[j].x = q0x * x[j].x + t0;
            x[j].y = q0x * x[j].y + t1;
            x[k].x = q3x * x[k].x + t2;
            x[k].y = q3x * x[k].y + t3;
        }
    }
}

template <typename T, int BA, int BB, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d2_plain_inplace(A (&x)[NS], const T *cp) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        if (((j >> BA) & 1) || ((j >> BB) & 1)) continue;
        const int idx[4] = {j, j | (1 << BB), j | (1 << BA), j | (1 << BA) | (1 << BB)};
        T tr[4], ti[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const T *row = cp + 8 * r;
            // everything but the real-part-of-the-diagonal term, which goes last and in place
            T ar = -(row[2 * r + 1] * x[idx[r]].y), ai = row[2 * r + 1] * x[idx[r]].x;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (c == r) continue;
                ar = ar + row[2 * c] * x[idx[c]].x - row[2 * c + 1] * x[idx[c]].y;
                ai = ai + row[2 * c] * x[idx[c]].y + row[2 * c + 1] * x[idx[c]].x;
            }
            tr[r] = ar;
            ti[r] = ai;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const T d = cp[8 * r + 2 * r];
            x[idx[r]].x = d * x[idx[r]].x + tr[r];
            x[idx[r]].y = d * x[idx[r]].y + ti[r];
        }
    }
}

// out-of-place forms (every output from fresh copies of the inputs): what the complex64 kernels use -- with 72 registers
// per thread (3 CTAs per SM) the temporaries of the in-place forms above spill (ptxas: 460 bytes), while the complex128
// kernels (128 registers) lose two thirds of their register moves with them (static IMAD.MOV 362 -> 135)
template <typename T, int B, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d1_plain_oop(A (&x)[NS], int kind, const T *cp) {
    const T q0x = cp[0], q0y = cp[1], q1x = cp[2], q1y = cp[3], q2x = cp[4], q2y = cp[5], q3x = cp[6], q3y = cp[7];
    if (kind == RG_D1_REAL) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            const A a = x[j], b = x[j | (1 << B)];
            x[j].x = q0x * a.x + q1x * b.x;
            x[j].y = q0x * a.y + q1x * b.y;
            x[j | (1 << B)].x = q2x * a.x + q3x * b.x;
            x[j | (1 << B)].y = q2x * a.y + q3x * b.y;
        }
    } else if (kind == RG_D1_RX) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            const A a = x[j], b = x[j | (1 << B)];
            x[j].x = q0x * a.x - q1y * b.y;
            x[j].y = q0x * a.y + q1y * b.x;
            x[j | (1 << B)].x = q3x * b.x - q2y * a.y;
            x[j | (1 << B)].y = q3x * b.y + q2y * a.x;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if ((j >> B) & 1) continue;
            const A a = x[j], b = x[j | (1 << B)];
            x[j].x = q0x * a.x - q0y * a.y + q1x * b.x - q1y * b.y;
            x[j].y = q0x * a.y + q0y * a.x + q1x * b.y + q1y * b.x;
            x[j | (1 << B)].x = q2x * a.x - q2y * a.y + q3x * b.x - q3y * b.y;
            x[j | (1 << B)].y = q2x * a.y + q2y * a.x + q3x * b.y + q3y * b.x;
        }
    }
}

template <typename T, int BA, int BB, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d2_plain_oop(A (&x)[NS], const T *cp) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        if (((j >> BA) & 1) || ((j >> BB) & 1)) continue;
        const int i0 = j, i1 = j | (1 << BB), i2 = j | (1 << BA), i3 = j | (1 << BA) | (1 << BB);
        const A v0 = x[i0], v1 = x[i1], v2 = x[i2], v3 = x[i3];
        A y[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const T *row = cp + 8 * r;
            y[r].x = row[0] * v0.x - row[1] * v0.y + row[2] * v1.x - row[3] * v1.y + row[4] * v2.x - row[5] * v2.y + row[6] * v3.x - row[7] * v3.y;
            y[r].y = row[0] * v0.y + row[1] * v0.x + row[2] * v1.y + row[3] * v1.x + row[4] * v2.y + row[5] * v2.x + row[6] * v3.y + row[7] * v3.x;
        }
        x[i0] = y[0];
        x[i1] = y[1];
        x[i2] = y[2];
        x[i3] = y[3];
    }
}

template <typename T, int B, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d1_plain(A (&x)[NS], int kind, const T *cp) {
#ifdef QSV_PLAIN_OOP
    reg_d1_plain_oop<T, B, NS>(x, kind, cp);
#else
    if constexpr (sizeof(T) == 8)
        reg_d1_plain_inplace<T, B, NS>(x, kind, cp);
    else
        reg_d1_plain_oop<T, B, NS>(x, kind, cp);
#endif
}
template <typename T, int BA, int BB, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d2_plain(A (&x)[NS], const T *cp) {
#ifdef QSV_PLAIN_OOP
    reg_d2_plain_oop<T, BA, BB, NS>(x, cp);
#else
#ifdef QSV_D2_INPLACE
    if constexpr (sizeof(T) == 8)
        reg_d2_plain_inplace<T, BA, BB, NS>(x, cp);
    else
#endif
        reg_d2_plain_oop<T, BA, BB, NS>(x, cp);
#endif
}

// 4x4 block on register bits BA > BB; matrix index = 2 * bit(BA) + bit(BB)
template <typename T, int BA, int BB, int NS, typename A>
__host__ __device__ __forceinline__ void reg_d2(A (&x)[NS], const T *mp, uint32_t creg) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
        if (((j >> BA) & 1) || ((j >> BB) & 1)) continue;
        if ((j & creg) == creg) {
            const int i0 = j, i1 = j | (1 << BB), i2 = j | (1 << BA), i3 = j | (1 << BA) | (1 << BB);
            const A v0 = x[i0], v1 = x[i1], v2 = x[i2], v3 = x[i3];
            A y[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const A *row = reinterpret_cast<const A *>(mp) + 4 * r;
                const A c0 = row[0], c1 = row[1], c2 = row[2], c3 = row[3];
                y[r].x = c0.x * v0.x - c0.y * v0.y + c1.x * v1.x - c1.y * v1.y + c2.x * v2.x - c2.y * v2.y + c3.x * v3.x - c3.y * v3.y;
                y[r].y = c0.x * v0.y + c0.y * v0.x + c1.x * v1.y + c1.y * v1.x + c2.x * v2.y + c2.y * v2.x + c3.x * v3.y + c3.y * v3.x;
            }
            x[i0] = y[0];
            x[i1] = y[1];
            x[i2] = y[2];
            x[i3] = y[3];
        }
    }
}

// diagonal / parity gates: phase table of NB bits; table bit b of slot j = tb[b] ^ parity(j & q[b])
template <typename T, int NB, int NS, typename A>
__host__ __device__ __forceinline__ void reg_diag(A (&x)[NS], const T *mp, bool thr_on, uint32_t creg, int tb0, int tb1,
                                                  uint32_t q0, uint32_t q1) {
    const A *tab = reinterpret_cast<const A *>(mp);
    if (NB == 0) {
        const A e = tab[0];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (thr_on && (j & creg) == creg) {
                const A a = x[j];
                x[j].x = e.x * a.x - e.y * a.y;
                x[j].y = e.x * a.y + e.y * a.x;
            }
        }
    } else if (NB == 1) {
        // one table bit (RZ, CRZ, IsingZZ, MultiRZ ...): a per-thread pair of phases, slots pick by parity
        const A e0 = tab[0], e1 = tab[1];
        const A pa = tb1 ? e1 : e0, pb = tb1 ? e0 : e1;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (thr_on && (j & creg) == creg) {
                const bool odd = popc32(j & q1) & 1;
                const T pr = odd ? pb.x : pa.x, pi = odd ? pb.y : pa.y;
                const A a = x[j];
                x[j].x = pr * a.x - pi * a.y;
                x[j].y = pr * a.y + pi * a.x;
            }
        }
    } else {
        const A e0 = tab[0], e1 = tab[1], e2 = tab[2], e3 = tab[3];
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            if (thr_on && (j & creg) == creg) {
                const int c0 = tb0 ^ (popc32(j & q0) & 1), c1 = tb1 ^ (popc32(j & q1) & 1);
                const T pr = c0 ? (c1 ? e3.x : e2.x) : (c1 ? e1.x : e0.x);
                const T pi = c0 ? (c1 ? e3.y : e2.y) : (c1 ? e1.y : e0.y);
                const A a = x[j];
                x[j].x = pr * a.x - pi * a.y;
                x[j].y = pr * a.y + pi * a.x;
            }
        }
    }
}

// A complex 4x4 gate on the two lowest LANE bits through mma.sync.m8n8k4.f64: lane 4m+t holds amplitude t of quad m; the
// real 8x8 form R of the gate is the B operand (this lane: R[lane/4][2*(lane%4)+s] for the k-step s = re / im inputs) and
// the D fragment comes back as (re_t, im_t) of the same quad, i.e. in place (tools/probes/dmma_gate_probe.cu).  DMMA runs
// on the same FP64 units as DFMA (profiles/r1_dmma_probe.txt) but needs 8x fewer issue slots per flop.
#ifdef __CUDA_ARCH__
__device__ __forceinline__ void dmma_8x8x4(double &d0, double &d1, double a, double b, double c0, double c1) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};"
        : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int NS> __device__ __forceinline__ void mma_gate_4x4(double2 (&x)[NS], const double *rtab, uint32_t lane) {
    const double2 bb = *reinterpret_cast<const double2 *>(rtab + (lane >> 2) * 8 + 2 * (lane & 3));
    constexpr int G = 4;  // slots in flight: the two MMAs of a slot depend on each other, different slots do not
#pragma unroll
    for (int j = 0; j < NS; j += G) {
        double t0[G], t1[G];
#pragma unroll
        for (int q = 0; q < G; ++q) dmma_8x8x4(t0[q], t1[q], x[j + q].x, bb.x, 0.0, 0.0);
#pragma unroll
        for (int q = 0; q < G; ++q) dmma_8x8x4(x[j + q].x, x[j + q].y, x[j + q].y, bb.y, t0[q], t1[q]);
    }
}
#endif

// All the arithmetic of one pass for one thread: x = the thread's 2^RB amplitudes, `outside` = index bits beyond the
// tile, spool = gate constants in the kernel's precision.
template <typename T, int RB, typename A>
__host__ __device__ __forceinline__ void pass_compute(A (&x)[1 << RB], const RegProgram &P, const RegPass &ps, uint32_t tid,
                                                      uint64_t outside, const T *spool) {
    constexpr int NS = 1 << RB;
#ifdef __CUDA_ARCH__
    // the tensor-core gate of the pass (every warp is converged here); the CPU emulator applies it before calling this
    if constexpr (sizeof(T) == 8) {
        if (ps.mma_off != NO_MMA) mma_gate_4x4<NS>(x, spool + ps.mma_off, tid & 31u);
    }
#endif
    // Thread-uniform diagonal gates (no target or control on a register bit) commute with every other gate of the pass:
    // their phases are multiplied per thread and applied to the 2^RB amplitudes once.
    if (ps.udiag_end > ps.gate_begin) {
        T pr = (T)1, pi = (T)0;
        bool any = false;
        for (int gi = ps.gate_begin; gi < ps.udiag_end; ++gi) {
            const RegGate &g = P.gates[gi];
            if ((outside & g.out_ctrl) != g.out_ctrl) continue;
            if ((tid & g.ctrl_thr) != g.ctrl_thr) continue;
            const int tb0 = (popc32(tid & g.thr_mask[0]) ^ popc64(outside & g.out_mask[0])) & 1;
            const int tb1 = (popc32(tid & g.thr_mask[1]) ^ popc64(outside & g.out_mask[1])) & 1;
            const A e = reinterpret_cast<const A *>(spool + g.mat_off)[2 * tb0 + tb1];
            const T nr = pr * e.x - pi * e.y;
            pi = pr * e.y + pi * e.x;
            pr = nr;
            any = true;
        }
        if (any) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                // two temporaries, results written over their own inputs (no register moves at the merge point)
                const T t = pi * x[j].y, u = pi * x[j].x;
                x[j].x = pr * x[j].x - t;
                x[j].y = pr * x[j].y + u;
            }
        }
    }
    for (int gi = ps.udiag_end; gi < ps.gate_end; ++gi) {
        const RegGate &g = P.gates[gi];
        if ((outside & g.out_ctrl) != g.out_ctrl) continue;  // CTA-uniform
        const bool thr_on = (tid & g.ctrl_thr) == g.ctrl_thr;
        const T *mp = spool + g.mat_off;
        const uint32_t creg = g.ctrl_reg;
        const bool plain = P.uniform_consts != 0 && g.ctrl_thr == 0 && creg == 0;  // CTA-uniform
        if (g.kind == RG_DIAG) {
            const int nb = g.rb;  // table bits in use
            const int tb0 = (popc32(tid & g.thr_mask[0]) ^ popc64(outside & g.out_mask[0])) & 1;
            const int tb1 = (popc32(tid & g.thr_mask[1]) ^ popc64(outside & g.out_mask[1])) & 1;
            const uint32_t q0 = g.reg_mask[0], q1 = g.reg_mask[1];
            if (nb == 0)
                reg_diag<T, 0, NS>(x, mp, thr_on, creg, tb0, tb1, q0, q1);
            else if (nb == 1)
                reg_diag<T, 1, NS>(x, mp, thr_on, creg, tb0, tb1, q0, q1);
            else
                reg_diag<T, 2, NS>(x, mp, thr_on, creg, tb0, tb1, q0, q1);
        } else if (g.kind == RG_D2) {
            if (plain) {
                const T *cp = plain_consts<T>(P, spool) + g.mat_off;
                const int pair = g.ra * 4 + g.rb;
                if (pair == 1 * 4 + 0) {
                    reg_d2_plain<T, 1, 0, NS>(x, cp);
                } else if (pair == 2 * 4 + 0) {
                    reg_d2_plain<T, 2, 0, NS>(x, cp);
                } else if (pair == 2 * 4 + 1) {
                    reg_d2_plain<T, 2, 1, NS>(x, cp);
                } else if constexpr (RB > 3) {
                    if (pair == 3 * 4 + 0)
                        reg_d2_plain<T, 3, 0, NS>(x, cp);
                    else if (pair == 3 * 4 + 1)
                        reg_d2_plain<T, 3, 1, NS>(x, cp);
                    else
                        reg_d2_plain<T, 3, 2, NS>(x, cp);
                }
            } else if (thr_on) {
                const int pair = g.ra * 4 + g.rb;
                if (pair == 1 * 4 + 0) {
                    reg_d2<T, 1, 0, NS>(x, mp, creg);
                } else if (pair == 2 * 4 + 0) {
                    reg_d2<T, 2, 0, NS>(x, mp, creg);
                } else if (pair == 2 * 4 + 1) {
                    reg_d2<T, 2, 1, NS>(x, mp, creg);
                } else if constexpr (RB > 3) {
                    if (pair == 3 * 4 + 0)
                        reg_d2<T, 3, 0, NS>(x, mp, creg);
                    else if (pair == 3 * 4 + 1)
                        reg_d2<T, 3, 1, NS>(x, mp, creg);
                    else
                        reg_d2<T, 3, 2, NS>(x, mp, creg);
                }
            }
        } else if (plain && g.kind != RG_D1_DIAG && g.kind != RG_D1_SWAP) {
            const T *cp = plain_consts<T>(P, spool) + g.mat_off;
            if (g.ra == 0) {
                reg_d1_plain<T, 0, NS>(x, g.kind, cp);
            } else if (g.ra == 1) {
                reg_d1_plain<T, 1, NS>(x, g.kind, cp);
            } else if (g.ra == 2) {
                reg_d1_plain<T, 2, NS>(x, g.kind, cp);
            } else if constexpr (RB > 3) {
                reg_d1_plain<T, 3, NS>(x, g.kind, cp);
            }
        } else {
            if (thr_on) {
                int tb1 = 0;
                if (g.kind == RG_D1_DIAG) tb1 = (popc32(tid & g.thr_mask[1]) ^ popc64(outside & g.out_mask[1])) & 1;
                if (g.ra == 0) {
                    reg_d1<T, 0, NS>(x, g.kind, mp, creg, tb1);
                } else if (g.ra == 1) {
                    reg_d1<T, 1, NS>(x, g.kind, mp, creg, tb1);
                } else if (g.ra == 2) {
                    reg_d1<T, 2, NS>(x, g.kind, mp, creg, tb1);
                } else if constexpr (RB > 3) {
                    reg_d1<T, 3, NS>(x, g.kind, mp, creg, tb1);
                }
            }
        }
    }
}

}  // namespace rt

// Host-side program construction (tile_regs.cu); exported for the test-only emulator.
struct LoweredGate;
void build_reg_program(int n_local, int dtype, uint64_t index_hi, const std::vector<const LoweredGate *> &gates,
                       uint64_t need, int L, int rb, rt::RegProgram &P);

}  // namespace qsv
