// Register-blocked fused circuit execution: many gates per HBM sweep, several gates per shared-memory pass.
//
// Second generation of the fused executor of tile_kernels.cu (north_star items b + c).  A CTA still owns a
// tile of 2^12 amplitudes whose index bits are L low (contiguous) bits plus 12 - L arbitrary high bits, but
// the tile now lives in REGISTERS: each of the 256 threads holds the 16 amplitudes that differ in four
// "register bits" of the tile.  A pass applies every scheduled gate whose non-diagonal target bits are among
// the current register bits without touching memory at all; diagonal / parity gates and controls may sit on
// any bit (register, thread or outside the tile) and ride along for free.  Between passes the tile is
// transposed through XOR-swizzled shared memory so that a different set of four bits becomes register bits.
// The first pass loads straight from HBM into registers and the last one stores straight back, so a sweep
// with P passes costs P - 1 shared-memory round trips instead of one per gate (the first-generation kernel
// is shared-memory-bandwidth-bound at ~1200 clk per gate per tile).
//
// Gates acting on disjoint qubits commute, so the host list-schedules the gates of a sweep into passes
// (dependency = a shared bit that at least one of the two gates touches non-diagonally).
//
// The reference has no counterpart: one custatevecApplyMatrix per gate
// (simulator/StateVectorCudaManaged.hpp:1433-1471), one full sweep each.
#include <algorithm>
#include <cstdlib>

#include "device_utils.cuh"
#include "qsv_internal.h"

namespace qsv {

namespace {

constexpr int RT_TB = 12;   // tile bits
constexpr int RT_RB = 4;    // register bits, at most (QSV_REGS_RB=3: 8 amplitudes per thread, 512 threads per CTA)
constexpr int RT_MAX_GATES = 48;
constexpr int RT_MAX_PASSES = 48;
constexpr int RT_POOL = 1280;  // doubles

enum : unsigned char { RG_D1 = 1, RG_D1_REAL = 2, RG_D1_RX = 3, RG_D1_SWAP = 4, RG_D2 = 5, RG_DIAG = 6 };

struct RegGate {
    unsigned char kind;
    unsigned char ra, rb;        // register-bit numbers (D1: ra; D2: ra = matrix MSB > rb)
    unsigned char ctrl_reg;      // controls among the register bits (mask over the slot number)
    unsigned char reg_mask[2];   // DIAG: table bit b = parity(slot & reg_mask[b]) ^ parity(tid & thr_mask[b]) ^ ...
    unsigned short ctrl_thr;     // controls among the thread bits (mask over threadIdx.x)
    unsigned short thr_mask[2];
    unsigned short mat_off;      // first double of this gate in the pool
    unsigned short pad0;
    uint64_t out_ctrl;           // controls outside the tile (global bit positions)
    uint64_t out_mask[2];        // ... ^ parity(outside & out_mask[b]); table index = 2 * bit[0] + bit[1]
};

struct RegPass {
    unsigned char rbits[RT_RB];          // tile-local position of register bit 0..3
    unsigned char tbits[RT_TB];          // tile-local position of thread bit 0..(TB - RB - 1)
    unsigned short gate_begin, gate_end;
};

struct RegProgram {
    int n_passes;
    int n_gates;
    int pool_used;
    int pad0;
    uint64_t index_hi;                // value of the index bits above the local shard
    unsigned char gpos[16];           // global bit of tile-local position p
    Holes tile_holes;                 // all tile bits, ascending (expands blockIdx.x to the tile base)
    RegPass passes[RT_MAX_PASSES];
    RegGate gates[RT_MAX_GATES];
    double pool[RT_POOL];
};

template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; static constexpr int SW = 3; };
template <> struct Cx<float> { using type = float2; static constexpr int SW = 4; };

// XOR swizzle of the shared-memory tile: the low SW bits (one 128-byte line) are XORed with every higher
// SW-bit field, so that lanes differing in any bits with distinct positions mod SW hit distinct banks.
template <int SW> __device__ __forceinline__ uint32_t swz(uint32_t e) {
    constexpr uint32_t M = (1u << SW) - 1u;
    uint32_t r = e;
#pragma unroll
    for (int s = SW; s < RT_TB; s += SW) r ^= (e >> s) & M;
    return r;
}

// ---- gates on register-resident amplitudes --------------------------------------------------------------
// Controls among the register bits are a per-slot predicate.  (A variant with separate unpredicated code paths
// plus a scheduling constraint that keeps controls out of the register bits was measured slower on B200:
// 196 ms vs 166 ms for the config-2 circuit -- more code, 128 registers with spills.)
template <typename T, int B, int NS, typename A>
__device__ __forceinline__ void reg_d1(A (&x)[NS], int kind, const T *mp, uint32_t creg) {
    const A q0 = reinterpret_cast<const A *>(mp)[0], q1 = reinterpret_cast<const A *>(mp)[1];
    const A q2 = reinterpret_cast<const A *>(mp)[2], q3 = reinterpret_cast<const A *>(mp)[3];
    if (kind == RG_D1_SWAP) {
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

// 4x4 block on register bits BA > BB; matrix index = 2 * bit(BA) + bit(BB)
template <typename T, int BA, int BB, int NS, typename A>
__device__ __forceinline__ void reg_d2(A (&x)[NS], const T *mp, uint32_t creg) {
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
__device__ __forceinline__ void reg_diag(A (&x)[NS], const T *mp, bool thr_on, uint32_t creg, int tb0, int tb1,
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
                const bool odd = __popc(j & q1) & 1;
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
                const int c0 = tb0 ^ (__popc(j & q0) & 1), c1 = tb1 ^ (__popc(j & q1) & 1);
                const T pr = c0 ? (c1 ? e3.x : e2.x) : (c1 ? e1.x : e0.x);
                const T pi = c0 ? (c1 ? e3.y : e2.y) : (c1 ? e1.y : e0.y);
                const A a = x[j];
                x[j].x = pr * a.x - pi * a.y;
                x[j].y = pr * a.y + pi * a.x;
            }
        }
    }
}

template <typename T, int RB, int MINB>
__global__ void __launch_bounds__(1 << (RT_TB - RB), MINB)
    k_tile_regs(void *single, void *const *table, const __grid_constant__ RegProgram P) {
    using A = typename Cx<T>::type;
    constexpr int SW = Cx<T>::SW;
    constexpr int NS = 1 << RB;            // amplitudes per thread
    constexpr int NT = 1 << (RT_TB - RB);  // threads per CTA
    constexpr int NTB = RT_TB - RB;        // thread bits
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *s = reinterpret_cast<A *>(smem_raw);
    A *gbase = reinterpret_cast<A *>(table ? table[blockIdx.y] : single);
    const uint64_t base = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    const uint64_t outside = base | P.index_hi;
    const uint32_t tid = threadIdx.x;
    // Gate constants go to shared memory once per CTA (in the kernel's precision) and are then read with uniform,
    // broadcast LDS: indexed constant-bank loads (LDC) inside the thread-divergent gate code run on the ADU pipe,
    // which ncu showed to be the busiest unit of the first version of this kernel (52 % vs 33 % FP64).
    T *spool = reinterpret_cast<T *>(smem_raw + (sizeof(A) << RT_TB));
    for (int i = tid; i < P.pool_used; i += NT) spool[i] = (T)P.pool[i];
    __syncthreads();

    A x[NS];
    for (int p = 0; p < P.n_passes; ++p) {
        const RegPass &ps = P.passes[p];
        uint32_t lt = 0;  // thread part of the tile-local index
#pragma unroll
        for (int i = 0; i < NTB; ++i) lt |= ((tid >> i) & 1u) << ps.tbits[i];
        uint32_t sr[RB];   // swizzled shared-memory offset of register bit b
        uint64_t gr[RB];   // global offset of register bit b
        const uint32_t st = swz<SW>(lt);
        const bool first = p == 0, last = p == P.n_passes - 1;
        uint64_t gt = 0;
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            sr[b] = swz<SW>(1u << ps.rbits[b]);
            gr[b] = 0;
        }
        if (first || last) {
#pragma unroll
            for (int i = 0; i < NTB; ++i) gt |= (uint64_t)((tid >> i) & 1u) << P.gpos[ps.tbits[i]];
            gt |= base;
#pragma unroll
            for (int b = 0; b < RB; ++b) gr[b] = 1ull << P.gpos[ps.rbits[b]];
        }
        auto goff = [&](int j) {
            uint64_t o = gt;
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if ((j >> b) & 1) o += gr[b];
            return o;
        };
        auto soff = [&](int j) {
            uint32_t o = st;
#pragma unroll
            for (int b = 0; b < RB; ++b)
                if ((j >> b) & 1) o ^= sr[b];
            return o;
        };
        if (first) {
#pragma unroll
            for (int j = 0; j < NS; ++j) x[j] = gbase[goff(j)];
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) x[j] = s[soff(j)];
        }

        for (int gi = ps.gate_begin; gi < ps.gate_end; ++gi) {
            const RegGate &g = P.gates[gi];
            if ((outside & g.out_ctrl) != g.out_ctrl) continue;  // CTA-uniform
            const bool thr_on = (tid & g.ctrl_thr) == g.ctrl_thr;
            const T *mp = spool + g.mat_off;
            const uint32_t creg = g.ctrl_reg;
            if (g.kind == RG_DIAG) {
                const int nb = g.rb;  // table bits in use
                const int tb0 = (__popc(tid & g.thr_mask[0]) ^ __popcll(outside & g.out_mask[0])) & 1;
                const int tb1 = (__popc(tid & g.thr_mask[1]) ^ __popcll(outside & g.out_mask[1])) & 1;
                const uint32_t q0 = g.reg_mask[0], q1 = g.reg_mask[1];
                if (nb == 0)
                    reg_diag<T, 0, NS>(x, mp, thr_on, creg, tb0, tb1, q0, q1);
                else if (nb == 1)
                    reg_diag<T, 1, NS>(x, mp, thr_on, creg, tb0, tb1, q0, q1);
                else
                    reg_diag<T, 2, NS>(x, mp, thr_on, creg, tb0, tb1, q0, q1);
            } else if (g.kind == RG_D2) {
                if (thr_on) {
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
            } else {
                if (thr_on) {
                    if (g.ra == 0) {
                        reg_d1<T, 0, NS>(x, g.kind, mp, creg);
                    } else if (g.ra == 1) {
                        reg_d1<T, 1, NS>(x, g.kind, mp, creg);
                    } else if (g.ra == 2) {
                        reg_d1<T, 2, NS>(x, g.kind, mp, creg);
                    } else if constexpr (RB > 3) {
                        reg_d1<T, 3, NS>(x, g.kind, mp, creg);
                    }
                }
            }
        }

        if (last) {
#pragma unroll
            for (int j = 0; j < NS; ++j) gbase[goff(j)] = x[j];
        } else {
            if (!first) __syncthreads();  // every thread has finished reading the previous layout
#pragma unroll
            for (int j = 0; j < NS; ++j) s[soff(j)] = x[j];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: normalisation of the gates of a sweep, pass scheduling, program construction
// ------------------------------------------------------------------------------------------------
struct NormGate {
    int kind = 0;                 // RG_*
    uint32_t dense = 0;           // tile-local bits touched non-diagonally
    uint32_t bits = 0;            // all tile-local bits the gate looks at
    int da = -1, db = -1;         // local positions of the dense targets (D2: da = matrix MSB)
    uint32_t ctrl_loc = 0;
    uint32_t forbid = 0;          // positions that must not be register bits of this gate's pass
    uint64_t ctrl_out = 0;
    uint32_t mask_loc[2] = {0, 0};  // DIAG table-bit masks
    uint64_t mask_out[2] = {0, 0};
    int n_table_bits = 0;
    std::vector<double> pool;     // what goes into the constant pool
};

uint64_t touched_of(const LoweredGate &g) {
    uint64_t m = 0;
    for (uint64_t o : g.offs) m |= o;
    return m;
}

}  // namespace

// can the register kernel take this gate inside a fused sweep?
bool regs_fusable(const LoweredGate &g, int n_local) {
    switch (g.kind) {
    case LoweredGate::DENSE: {
        const uint64_t t = touched_of(g);
        if ((t >> n_local) != 0) return false;
        if (g.k == 2) return g.tgt_bits.size() == 2;
        if (g.k != 1) return false;
        if (g.tgt_bits.size() == 1) return g.offs[0] == 0;
        // two-level block of a multi-qubit gate: taken as a 4x4 when it spans exactly the two bits it touches
        uint64_t fixed = 0;
        for (int h : g.holes) fixed |= 1ull << h;
        return __builtin_popcountll(t) == 2 && (fixed & ~g.ctrl_mask) == t;
    }
    case LoweredGate::DIAG: return g.k <= 2;
    case LoweredGate::PARITY: return true;
    default: return false;
    }
}

// dense-touched bits >= L that a sweep must make tile bits for this gate
uint64_t regs_need_bits(const LoweredGate &g) { return g.kind == LoweredGate::DENSE ? touched_of(g) : 0; }

namespace {

template <typename T, int RB> void launch_regs_t(State &sv, const RegProgram &P, void *const *table, int n_vecs) {
    // register budget: 16 amplitudes per thread need 2 CTAs of 256 threads (128 registers); 8 amplitudes per thread run as
    // 2 CTAs of 512 threads (64 registers)
    constexpr int MINB = RB == 4 ? (sizeof(T) == 8 ? 2 : 3) : 2;
    const size_t smem = ((size_t)1 << RT_TB) * sizeof(typename Cx<T>::type) + RT_POOL * sizeof(T);
    static bool configured[64] = {false};  // per device: function attributes belong to the device's context
    auto kern = k_tile_regs<T, RB, MINB>;
    if (!configured[sv.device & 63]) {
        QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[sv.device & 63] = true;
    }
    dim3 grid((unsigned)(1ull << (sv.n - RT_TB)), (unsigned)n_vecs);
    kern<<<grid, 1 << (RT_TB - RB), smem, sv.stream>>>(table ? nullptr : sv.data, table, P);
    QSV_CUDA(cudaGetLastError());
}

int regs_rb() {
    const char *v = std::getenv("QSV_REGS_RB");
    const int rb = v ? std::atoi(v) : 4;
    return rb == 3 ? 3 : 4;
}

}  // namespace

// One sweep of the register kernel over `gates` (all regs_fusable, dense-touched bits >= L listed in `need`).
void run_sweep_regs(State &sv, const std::vector<const LoweredGate *> &gates, uint64_t need, int L, void *const *table,
                    int n_vecs) {
    const int n = sv.n;
    const int tb = RT_TB;
    const int rb = regs_rb();
    QSV_CHECK(n >= tb, "internal: register tile kernel needs at least 12 local qubits");
    QSV_CHECK((int)gates.size() <= RT_MAX_GATES, "internal: too many gates in a sweep");
    RegProgram P;
    memset(&P, 0, sizeof(P));
    P.index_hi = sv.index_hi;

    // tile bits: the low L bits, the needed high bits, then the lowest free bits
    std::vector<int> hi;
    for (int b = L; b < n; ++b)
        if (need >> b & 1) hi.push_back(b);
    for (int b = L; b < n && (int)hi.size() < tb - L; ++b)
        if (!(need >> b & 1)) hi.push_back(b);
    std::sort(hi.begin(), hi.end());
    QSV_CHECK((int)hi.size() == tb - L, "internal: tile bit selection");
    int pos[64];
    for (int b = 0; b < 64; ++b) pos[b] = -1;
    std::vector<int> tile_bits;
    for (int b = 0; b < L; ++b) tile_bits.push_back(b);
    for (int b : hi) tile_bits.push_back(b);
    for (int p = 0; p < tb; ++p) {
        pos[tile_bits[p]] = p;
        P.gpos[p] = (unsigned char)tile_bits[p];
    }
    P.tile_holes = make_holes(tile_bits.data(), tb, 0);
    auto map_mask = [&](uint64_t m, uint64_t &outside) {
        uint32_t loc = 0;
        outside = 0;
        for (int b = 0; b < 64; ++b)
            if (m >> b & 1) {
                if (pos[b] >= 0)
                    loc |= 1u << pos[b];
                else
                    outside |= 1ull << b;
            }
        return loc;
    };

    // ---- normalise ---------------------------------------------------------------------------------
    const size_t m = gates.size();
    std::vector<NormGate> ng(m);
    for (size_t i = 0; i < m; ++i) {
        const LoweredGate &g = *gates[i];
        NormGate &o = ng[i];
        o.ctrl_loc = map_mask(g.ctrl_mask, o.ctrl_out);
        o.bits = o.ctrl_loc;
        if (g.kind == LoweredGate::DENSE) {
            uint64_t out = 0;
            o.dense = map_mask(touched_of(g), out);
            QSV_CHECK(out == 0, "internal: dense target outside the tile");
            o.bits |= o.dense;
            if (g.k == 1 && g.tgt_bits.size() == 1) {
                o.da = pos[g.tgt_bits[0]];
                const cplx zero(0.0, 0.0), one(1.0, 0.0);
                bool real = true, rx = true;
                for (const cplx &c : g.mat) real = real && c.imag() == 0.0;
                rx = g.mat[0].imag() == 0.0 && g.mat[3].imag() == 0.0 && g.mat[1].real() == 0.0 && g.mat[2].real() == 0.0;
                if (g.mat[0] == zero && g.mat[3] == zero && g.mat[1] == one && g.mat[2] == one)
                    o.kind = RG_D1_SWAP;
                else if (real)
                    o.kind = RG_D1_REAL;
                else if (rx)
                    o.kind = RG_D1_RX;
                else
                    o.kind = RG_D1;
                for (const cplx &c : g.mat) {
                    o.pool.push_back(c.real());
                    o.pool.push_back(c.imag());
                }
            } else {
                // 4x4 on two bits; the matrix MSB goes to the target with the higher local position
                o.kind = RG_D2;
                cplx M[16];
                int hi_bit, lo_bit;
                if (g.k == 2) {
                    hi_bit = g.tgt_bits[0];
                    lo_bit = g.tgt_bits[1];
                    for (int q = 0; q < 16; ++q) M[q] = g.mat[q];
                } else {
                    const uint64_t t = touched_of(g);
                    lo_bit = __builtin_ctzll(t);
                    hi_bit = 63 - __builtin_clzll(t);
                    for (int q = 0; q < 16; ++q) M[q] = (q / 4 == q % 4) ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
                    int idx[2];
                    for (int e = 0; e < 2; ++e)
                        idx[e] = (int)(((g.offs[e] >> hi_bit) & 1) << 1 | ((g.offs[e] >> lo_bit) & 1));
                    M[idx[0] * 4 + idx[0]] = g.mat[0];
                    M[idx[0] * 4 + idx[1]] = g.mat[1];
                    M[idx[1] * 4 + idx[0]] = g.mat[2];
                    M[idx[1] * 4 + idx[1]] = g.mat[3];
                }
                if (pos[hi_bit] < pos[lo_bit]) {
                    auto sw = [](int q) { return ((q & 1) << 1) | (q >> 1); };
                    cplx M2[16];
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) M2[r * 4 + c] = M[sw(r) * 4 + sw(c)];
                    for (int q = 0; q < 16; ++q) M[q] = M2[q];
                    std::swap(hi_bit, lo_bit);
                }
                o.da = pos[hi_bit];
                o.db = pos[lo_bit];
                for (int q = 0; q < 16; ++q) {
                    o.pool.push_back(M[q].real());
                    o.pool.push_back(M[q].imag());
                }
            }
        } else {
            o.kind = RG_DIAG;
            cplx tab[4];
            o.n_table_bits = g.kind == LoweredGate::PARITY ? 1 : g.k;
            if (g.kind == LoweredGate::PARITY) {
                o.mask_loc[1] = map_mask(g.zmask, o.mask_out[1]);
                tab[0] = tab[2] = g.mat[0];
                tab[1] = tab[3] = g.mat[1];
            } else if (g.k == 0) {
                tab[0] = tab[1] = tab[2] = tab[3] = g.mat[0];
            } else if (g.k == 1) {
                o.mask_loc[1] = map_mask(1ull << g.tgt_bits[0], o.mask_out[1]);
                tab[0] = tab[2] = g.mat[0];
                tab[1] = tab[3] = g.mat[1];
            } else {
                o.mask_loc[0] = map_mask(1ull << g.tgt_bits[0], o.mask_out[0]);
                o.mask_loc[1] = map_mask(1ull << g.tgt_bits[1], o.mask_out[1]);
                for (int q = 0; q < 4; ++q) tab[q] = g.mat[q];
            }
            o.bits |= o.mask_loc[0] | o.mask_loc[1];
            for (int q = 0; q < 4; ++q) {
                o.pool.push_back(tab[q].real());
                o.pool.push_back(tab[q].imag());
            }
        }
    }

    // ---- list-schedule into passes -----------------------------------------------------------------
    // gate j depends on an earlier gate i when they share a tile bit that one of them touches non-diagonally
    std::vector<std::vector<int>> preds(m);
    for (size_t j = 0; j < m; ++j)
        for (size_t i = 0; i < j; ++i)
            if ((ng[i].dense & ng[j].bits) | (ng[i].bits & ng[j].dense)) preds[j].push_back((int)i);
    std::vector<char> done(m, 0);
    size_t n_done = 0;
    auto grow = [&](int seed, std::vector<char> &dn, uint32_t &R, uint32_t &F, std::vector<int> &order) {
        // greedy: keep adding the ready gate that needs the fewest new register bits
        R = 0;
        F = 0;
        order.clear();
        int next = seed;
        while (next >= 0) {
            order.push_back(next);
            dn[next] = 1;
            R |= ng[next].dense;
            F |= ng[next].forbid;
            next = -1;
            int best_new = 99;
            for (size_t j = 0; j < m; ++j) {
                if (dn[j]) continue;
                bool ready = true;
                for (int i : preds[j]) ready = ready && dn[i];
                if (!ready) continue;
                const uint32_t u = R | ng[j].dense;
                if (__builtin_popcount(u) > rb || (u & (F | ng[j].forbid))) continue;
                const int nw = __builtin_popcount(u) - __builtin_popcount(R);
                if (nw < best_new) {
                    best_new = nw;
                    next = (int)j;
                }
            }
        }
    };
    int n_pool = 0;
    const int SW = sv.dtype == QSV_C128 ? 3 : 4;
    while (n_done < m) {
        // try every ready gate as the seed of the next pass, keep the pass that retires the most gates
        std::vector<int> best_order;
        uint32_t best_R = 0, best_F = 0;
        for (size_t sd = 0; sd < m; ++sd) {
            if (done[sd]) continue;
            bool ready = true;
            for (int i : preds[sd]) ready = ready && done[i];
            if (!ready) continue;
            std::vector<char> dn = done;
            std::vector<int> order;
            uint32_t R, F;
            grow((int)sd, dn, R, F, order);
            if (order.size() > best_order.size()) {
                best_order = order;
                best_R = R;
                best_F = F;
            }
        }
        QSV_CHECK(!best_order.empty(), "internal: pass scheduling made no progress");
        QSV_CHECK(P.n_passes < RT_MAX_PASSES, "internal: too many passes in a sweep");
        RegPass &ps = P.passes[P.n_passes++];
        // register bits: the dense bits of the pass, filled up with the highest other positions
        uint32_t R = best_R;
        for (int p = tb - 1; p >= 0 && __builtin_popcount(R) < rb; --p)
            if (!((R | best_F) >> p & 1)) R |= 1u << p;
        QSV_CHECK(__builtin_popcount(R) == rb, "internal: no free register bits for a pass");
        int regbit_of[16];
        for (int p = 0, k = 0; p < tb; ++p) {
            regbit_of[p] = -1;
            if (R >> p & 1) {
                regbit_of[p] = k;
                ps.rbits[k++] = (unsigned char)p;
            }
        }
        // thread bits: lanes take the lowest positions (coalescing); among them, the first SW lane bits get
        // distinct positions mod SW when possible (conflict-free swizzled shared-memory accesses)
        std::vector<int> rest;
        for (int p = 0; p < tb; ++p)
            if (!(R >> p & 1)) rest.push_back(p);
        std::vector<int> lanes(rest.begin(), rest.begin() + 5), ordered;
        std::vector<char> used(5, 0);
        uint32_t seen = 0;
        for (int q = 0; q < 5 && (int)ordered.size() < SW; ++q)
            if (!(seen >> (lanes[q] % SW) & 1)) {
                seen |= 1u << (lanes[q] % SW);
                ordered.push_back(lanes[q]);
                used[q] = 1;
            }
        for (int q = 0; q < 5; ++q)
            if (!used[q]) ordered.push_back(lanes[q]);
        for (size_t q = 5; q < rest.size(); ++q) ordered.push_back(rest[q]);
        int thrbit_of[16];
        for (int p = 0; p < tb; ++p) thrbit_of[p] = -1;
        for (int k = 0; k < tb - rb; ++k) {
            ps.tbits[k] = (unsigned char)ordered[k];
            thrbit_of[ordered[k]] = k;
        }
        auto split = [&](uint32_t loc, unsigned &reg, unsigned &thr) {
            reg = 0;
            thr = 0;
            for (int p = 0; p < tb; ++p)
                if (loc >> p & 1) {
                    if (regbit_of[p] >= 0)
                        reg |= 1u << regbit_of[p];
                    else
                        thr |= 1u << thrbit_of[p];
                }
        };
        ps.gate_begin = (unsigned short)P.n_gates;
        for (int gi : best_order) {
            const NormGate &o = ng[gi];
            RegGate &t = P.gates[P.n_gates++];
            t.kind = (unsigned char)o.kind;
            unsigned reg, thr;
            split(o.ctrl_loc, reg, thr);
            t.ctrl_reg = (unsigned char)reg;
            t.ctrl_thr = (unsigned short)thr;
            t.out_ctrl = o.ctrl_out;
            if (o.kind == RG_DIAG) {
                t.rb = (unsigned char)o.n_table_bits;
                for (int b = 0; b < 2; ++b) {
                    split(o.mask_loc[b], reg, thr);
                    t.reg_mask[b] = (unsigned char)reg;
                    t.thr_mask[b] = (unsigned short)thr;
                    t.out_mask[b] = o.mask_out[b];
                }
            } else {
                t.ra = (unsigned char)regbit_of[o.da];
                if (o.kind == RG_D2) t.rb = (unsigned char)regbit_of[o.db];
            }
            QSV_CHECK(n_pool + (int)o.pool.size() <= RT_POOL, "internal: constant pool overflow");
            t.mat_off = (unsigned short)n_pool;
            for (double v : o.pool) P.pool[n_pool++] = v;
            done[gi] = 1;
            ++n_done;
        }
        ps.gate_end = (unsigned short)P.n_gates;
    }

    P.pool_used = n_pool;
    sv.stat_launches += 1;
    sv.stat_sweeps += 1;
    if (sv.dtype == QSV_C128) {
        if (rb == 4)
            launch_regs_t<double, 4>(sv, P, table, n_vecs);
        else
            launch_regs_t<double, 3>(sv, P, table, n_vecs);
    } else {
        if (rb == 4)
            launch_regs_t<float, 4>(sv, P, table, n_vecs);
        else
            launch_regs_t<float, 3>(sv, P, table, n_vecs);
    }
}

}  // namespace qsv
