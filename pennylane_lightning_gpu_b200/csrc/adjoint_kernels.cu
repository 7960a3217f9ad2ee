// Batched generator inner products for the adjoint method: Im/Re <bra| G_k |ket> for MANY generators G_k in one
// read of (bra, ket).
//
// The reference evaluates one trainable parameter at a time (algorithms/AdjointDiffGPU.hpp:562-592: copy lambda to
// mu, apply the generator to mu, one cuBLAS dot per observable, one D2H per parameter).  Generators of gates that
// act on disjoint wires commute with each other and with each other's gates, so a whole layer of rotations can be
// differentiated against the SAME pair of vectors (see adjoint_jacobian in circuit.cu).  This kernel does that:
// a CTA stages a tile of `ket` in shared memory (the low L index bits plus arbitrary high bits, like the fused gate
// kernels), keeps the matching `bra` amplitudes in registers, and accumulates conj(bra_i) * (G_k ket)_i for every
// generator whose non-diagonal target bit lies inside the tile; diagonal / parity generators can sit on any bit.
// One warp-shuffle reduction per generator and warp, one atomicAdd pair per generator and CTA.
#include <algorithm>

#include "device_utils.cuh"
#include "qsv_internal.h"

namespace qsv {

namespace {

constexpr int GT_JB = 4;                 // the top GT_JB tile bits distinguish the amplitudes of one thread
constexpr int GT_EPT = 1 << GT_JB;       // 16 amplitudes per thread
constexpr int GT_MAX = 48;               // generators per launch
constexpr int GT_TB_MAX = 13;            // tile bits: 2^12 (256 threads) or 2^13 (512 threads) amplitudes of ket in shared memory

struct GenDesc {
    int kind;            // 0: diagonal table (<= 1 table bit: parity of index & zmask), 1: 2x2 block on one tile bit,
                         // 2: Pauli word (X/Y letters on tile bits, Z letters anywhere),
                         // 3: "Z sum": kind 0 without controls whose parity mask has at most one bit (RZ, PhaseShift ...)
    int slot;            // complex output slot
    unsigned tbit;       // kind 1: tile-local position of the target bit; kind 2: tile-local flip mask;
                         // kind 3: tile-local position of the parity bit, ZS_OUTSIDE (sign from xg & base) or ZS_NONE
    unsigned jinfo;      // what the generator sees of the thread's GT_JB amplitude bits j (CTA-uniform, so the per-amplitude
                         // work is branch-free): bits 0-3 parity mask over j, bits 4-7 control mask over j, bit 8 constant
                         // parity flip (kind 2: the flipped j bits under the parity mask)
    uint64_t ctrl;       // global bits that must be 1          } outside the j columns: thread-constant
    uint64_t zmask;      // kind 0 / 2: global parity-mask bits }
    uint64_t xg;         // kind 2: flip mask on global bits; kind 3: the parity bit as a global mask
    double m[8];         // kind 1: row-major complex 2x2; kind 0 / 3: phases for even / odd parity; kind 2: i^ny
};
constexpr unsigned ZS_OUTSIDE = 255u, ZS_NONE = 254u;

struct GenProgram {
    int n_gens;
    int n_first;                // generators [0, n_first) are of kind 1 / 2 (they need bra),
    int n_diag;                 // [n_first, n_diag) generic diagonal ones (kind 0), [n_diag, n_gens) Z sums (kind 3)
    int L;
    unsigned char hi_bits[16];  // global positions of tile bits L..TB-1
    Holes tile_holes;
    GenDesc g[GT_MAX];
};

__device__ __forceinline__ void pick_w(const double (&Wr)[GT_EPT], const double (&Wi)[GT_EPT], unsigned idx, double &wr, double &wi) {
    switch (idx & 15u) {  // CTA-uniform
    case 0: wr = Wr[0]; wi = Wi[0]; break;
    case 1: wr = Wr[1]; wi = Wi[1]; break;
    case 2: wr = Wr[2]; wi = Wi[2]; break;
    case 3: wr = Wr[3]; wi = Wi[3]; break;
    case 4: wr = Wr[4]; wi = Wi[4]; break;
    case 5: wr = Wr[5]; wi = Wi[5]; break;
    case 6: wr = Wr[6]; wi = Wi[6]; break;
    case 7: wr = Wr[7]; wi = Wi[7]; break;
    case 8: wr = Wr[8]; wi = Wi[8]; break;
    case 9: wr = Wr[9]; wi = Wi[9]; break;
    case 10: wr = Wr[10]; wi = Wi[10]; break;
    case 11: wr = Wr[11]; wi = Wi[11]; break;
    case 12: wr = Wr[12]; wi = Wi[12]; break;
    case 13: wr = Wr[13]; wi = Wi[13]; break;
    case 14: wr = Wr[14]; wi = Wi[14]; break;
    default: wr = Wr[15]; wi = Wi[15]; break;
    }
}

// Sums 8 values over the 32 lanes of a warp through a per-warp scratch area (32 x 9 doubles, padded): 8 stores, 8 loads,
// 7 adds and two shuffle levels instead of 8 x 5 shuffle levels.  Afterwards lane 4 * i (and its three neighbours) holds
// the warp total of v[i].
constexpr int RED_PAD = 9;
__device__ __forceinline__ double warp_reduce8(const double (&v)[8], double *scratch, uint32_t lane) {
#pragma unroll
    for (int i = 0; i < 8; ++i) scratch[lane * RED_PAD + i] = v[i];
    __syncwarp();
    const uint32_t i = lane >> 2, part = lane & 3u;
    double a = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) a += scratch[(8u * part + t) * RED_PAD + i];
    a += __shfl_xor_sync(0xffffffffu, a, 1);
    a += __shfl_xor_sync(0xffffffffu, a, 2);
    __syncwarp();
    return a;
}

// One CTA = one tile of 2^TB amplitudes: ket in shared memory, the matching bra amplitudes in registers (16 per thread:
// element e = tid + NT * j, j = the top four tile bits).  Everything a generator needs to know about an amplitude splits
// into a thread-constant part (computed once per generator and thread) and a part that depends on j only, which is
// CTA-uniform and unrolled -- so the inner loops are one shared-memory read and four FP64 multiply-adds per amplitude.
// Cross-lane sums are taken four generators at a time (warp_reduce8).
template <typename T, int TB>
__global__ void __launch_bounds__(1 << (TB - GT_JB))
    k_bra_gens_ket(const void *__restrict__ bra, const void *__restrict__ ket, double *out,
                   const __grid_constant__ GenProgram P) {
    using A = typename VecOf<T, 1>::type;
    constexpr int NT = 1 << (TB - GT_JB), NW = NT / 32;
    constexpr int NZ = 1 + 5 + GT_JB;  // per-warp Z-sum table: total, sign by lane bit 0..4, sign by j bit 0..3
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *s = reinterpret_cast<A *>(smem_raw);
    double(*s_acc)[NW][2] = reinterpret_cast<double(*)[NW][2]>(smem_raw + (sizeof(A) << TB));
    double *s_red = reinterpret_cast<double *>(s_acc + GT_MAX);           // NW x 32 x RED_PAD
    double(*s_z)[NZ][2] = reinterpret_cast<double(*)[NZ][2]>(s_red + NW * 32 * RED_PAD);  // NW x NZ x 2
    const uint64_t base = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    double *scratch = s_red + warp * (32 * RED_PAD);

    // element e of the tile (local index) lives at global index base | deposit(e); g0 = the thread part (j bits clear)
    uint64_t g0 = tid & ((1u << P.L) - 1u);
    for (int q = P.L; q < TB - GT_JB; ++q) g0 |= (uint64_t)((tid >> q) & 1u) << P.hi_bits[q - P.L];
    g0 |= base;
    uint64_t col[GT_JB];
#pragma unroll
    for (int q = 0; q < GT_JB; ++q) col[q] = 1ull << P.hi_bits[TB - GT_JB + q - P.L];
    const bool same = bra == ket;  // expectation values: one read of the vector serves both sides
    A b[GT_EPT];
#pragma unroll
    for (int j = 0; j < GT_EPT; ++j) {
        uint64_t gi = g0;
#pragma unroll
        for (int q = 0; q < GT_JB; ++q)
            if ((j >> q) & 1) gi |= col[q];
        s[tid + (uint32_t)j * NT] = reinterpret_cast<const A *>(ket)[gi];
        if (!same) b[j] = reinterpret_cast<const A *>(bra)[gi];
    }
    __syncthreads();
    if (same) {
#pragma unroll
        for (int j = 0; j < GT_EPT; ++j) b[j] = s[tid + (uint32_t)j * NT];
    }

    // the per-thread results of up to four generators wait in the warp's scratch area for one batched cross-lane sum
    // (straight to shared memory: holding them in registers costs 16 registers this kernel does not have)
    auto put = [&](int q, double re, double im) {
        scratch[lane * RED_PAD + 2 * q] = re;
        scratch[lane * RED_PAD + 2 * q + 1] = im;
    };
    auto flush_group = [&](int k0, int count) {
        __syncwarp();
        const uint32_t i = lane >> 2, part = lane & 3u;
        double a = 0.0;
#pragma unroll
        for (int t = 0; t < 8; ++t) a += scratch[(8u * part + t) * RED_PAD + i];
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
        if (part == 0u && (int)(i >> 1) < count) s_acc[k0 + (i >> 1)][warp][i & 1u] = a;
        __syncwarp();
    };

    // ---- generators with off-diagonal parts: Pauli words and 2x2 blocks --------------------------------------
    for (int k0 = 0; k0 < P.n_first; k0 += 4) {
        const int count = min(4, P.n_first - k0);
        for (int q = 0; q < count; ++q) {
            const GenDesc &g = P.g[k0 + q];
            const unsigned mj = g.jinfo & 15u, cjm = (g.jinfo >> 4) & 15u;
            const bool t_on = (g0 & g.ctrl) == g.ctrl;
            double re = 0.0, im = 0.0;
            if (g.kind == 2) {
                // (P ket)_i = i^ny * (-1)^{popc(i' & z)} * ket_i' with i' = i ^ x (only where the control bits are set:
                // generators |1><1| (x) X / Y of controlled rotations); the constant i^ny is applied after the sum.
                // partner element = (tid ^ low part of the flip mask) + NT * (j ^ its j part): one thread-level XOR per
                // generator, the j part is CTA-uniform arithmetic
                const bool t_odd = ((__popcll((g0 ^ g.xg) & g.zmask) ^ (g.jinfo >> 8)) & 1u) != 0u;
                const A *sp = s + (tid ^ (g.tbit & (uint32_t)(NT - 1)));
                const uint32_t jx = g.tbit >> (TB - GT_JB);
                double ar[2] = {0.0, 0.0}, ai[2] = {0.0, 0.0};  // four independent chains of multiply-adds
                if (t_on) {
                    if ((mj | cjm) == 0u) {
#pragma unroll
                        for (int j = 0; j < GT_EPT; ++j) {
                            const A x = sp[((uint32_t)j ^ jx) * NT];
                            ar[j & 1] = fma((double)b[j].x, (double)x.x, ar[j & 1]);
                            ar[j & 1] = fma((double)b[j].y, (double)x.y, ar[j & 1]);
                            ai[j & 1] = fma((double)b[j].x, (double)x.y, ai[j & 1]);
                            ai[j & 1] = fma(-(double)b[j].y, (double)x.x, ai[j & 1]);
                        }
                    } else if (cjm == 0u) {
                        // the sign depends on the j bits (a Z / Y letter sits on one of them): flip the sign bit of the
                        // partner amplitude with a CTA-uniform mask instead of multiplying
#pragma unroll
                        for (int j = 0; j < GT_EPT; ++j) {
                            const A x = sp[((uint32_t)j ^ jx) * NT];
                            const int flip = (__popc((unsigned)j & mj) & 1) << 31;  // CTA-uniform
                            const double xr = __hiloint2double(__double2hiint((double)x.x) ^ flip, __double2loint((double)x.x));
                            const double xi = __hiloint2double(__double2hiint((double)x.y) ^ flip, __double2loint((double)x.y));
                            ar[j & 1] = fma((double)b[j].x, xr, ar[j & 1]);
                            ar[j & 1] = fma((double)b[j].y, xi, ar[j & 1]);
                            ai[j & 1] = fma((double)b[j].x, xi, ai[j & 1]);
                            ai[j & 1] = fma(-(double)b[j].y, xr, ai[j & 1]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < GT_EPT; ++j) {
                            if (((unsigned)j & cjm) != cjm) continue;
                            const A x = sp[((uint32_t)j ^ jx) * NT];
                            const double sg = (__popc((unsigned)j & mj) & 1) ? -1.0 : 1.0;  // CTA-uniform
                            const double bx = sg * (double)b[j].x, by = sg * (double)b[j].y;
                            ar[j & 1] = fma(bx, (double)x.x, ar[j & 1]);
                            ar[j & 1] = fma(by, (double)x.y, ar[j & 1]);
                            ai[j & 1] = fma(bx, (double)x.y, ai[j & 1]);
                            ai[j & 1] = fma(-by, (double)x.x, ai[j & 1]);
                        }
                    }
                }
                const double r0 = t_odd ? -(ar[0] + ar[1]) : ar[0] + ar[1], i0 = t_odd ? -(ai[0] + ai[1]) : ai[0] + ai[1];
                re = g.m[0] * r0 - g.m[1] * i0;
                im = g.m[0] * i0 + g.m[1] * r0;
            } else {
                const uint32_t bit = 1u << g.tbit;
                const double m0 = g.m[0], m1 = g.m[1], m2 = g.m[2], m3 = g.m[3];
                const double m4 = g.m[4], m5 = g.m[5], m6 = g.m[6], m7 = g.m[7];
                if (t_on) {
#pragma unroll
                    for (int j = 0; j < GT_EPT; ++j) {
                        if (((unsigned)j & cjm) != cjm) continue;
                        const uint32_t e = tid + (uint32_t)j * NT;
                        const A x0 = s[e & ~bit], x1 = s[e | bit];
                        const bool hi = (e & bit) != 0;
                        const double ar = hi ? m4 : m0, ai = hi ? m5 : m1, br = hi ? m6 : m2, bi = hi ? m7 : m3;
                        const double yr = ar * (double)x0.x - ai * (double)x0.y + br * (double)x1.x - bi * (double)x1.y;
                        const double yi = ar * (double)x0.y + ai * (double)x0.x + br * (double)x1.y + bi * (double)x1.x;
                        re += (double)b[j].x * yr + (double)b[j].y * yi;
                        im += (double)b[j].x * yi - (double)b[j].y * yr;
                    }
                }
            }
            put(q, re, im);
        }
        flush_group(k0, count);
    }

    // ---- diagonal / parity generators -------------------------------------------------------------------------
    // conj(b_i) * p(i) * k_i with t0_i = conj(b_i) * k_i shared by all of them; the thread's 16 amplitudes differ in the j
    // bits only.
    if (P.n_first < P.n_gens) {
        double Wr[GT_EPT], Wi[GT_EPT];
#pragma unroll
        for (int j = 0; j < GT_EPT; ++j) {
            const A x = s[tid + (uint32_t)j * NT];
            Wr[j] = fma((double)b[j].x, (double)x.x, (double)b[j].y * (double)x.y);
            Wi[j] = fma((double)b[j].x, (double)x.y, -((double)b[j].y * (double)x.x));
        }
        if (P.n_diag < P.n_gens) {
            // Z sums: every generator of the form  e_even * [bit = 0] + e_odd * [bit = 1]  (RZ, PhaseShift, Z-type
            // observables) needs only M = sum_i t0_i and Z = sum_i (-1)^{bit(i)} t0_i.  One pass yields Z for EVERY bit of
            // the index at once: the four j bits by a pruned Walsh-Hadamard transform in registers, the five lane bits by a
            // butterfly over the warp, the warp bits and the bits outside the tile as signs in the final sum below.
            double zr[8], zi_[8];  // zr/zi_[q] = sum_j (-1)^{j_q} t0_j for q < 4; index 4: plain sum
            {
                double sr[GT_EPT], si[GT_EPT];
#pragma unroll
                for (int j = 0; j < GT_EPT; ++j) {
                    sr[j] = Wr[j];
                    si[j] = Wi[j];
                }
#pragma unroll
                for (int q = 0; q < GT_JB; ++q) {
                    const int half = GT_EPT >> (q + 1);  // values left after this level
                    double dr = 0.0, di = 0.0;
#pragma unroll
                    for (int t = 0; t < half; ++t) {
                        dr += sr[2 * t] - sr[2 * t + 1];
                        di += si[2 * t] - si[2 * t + 1];
                        sr[t] = sr[2 * t] + sr[2 * t + 1];
                        si[t] = si[2 * t] + si[2 * t + 1];
                    }
                    zr[q] = dr;
                    zi_[q] = di;
                }
                zr[4] = sr[0];
                zi_[4] = si[0];
            }
            // j-bit sums over the lanes (8 values at once)
            {
                double w8[8];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    w8[2 * q] = zr[q];
                    w8[2 * q + 1] = zi_[q];
                }
                const double a = warp_reduce8(w8, scratch, lane);
                if ((lane & 3u) == 0u) s_z[warp][6 + (lane >> 3)][(lane >> 2) & 1u] = a;
            }
            // lane bits: Walsh-Hadamard butterfly of the plain per-thread sum over the 32 lanes
            double hr = zr[4], hi = zi_[4];
#pragma unroll
            for (int m = 1; m < 32; m <<= 1) {
                const double orr = __shfl_xor_sync(0xffffffffu, hr, m), oi = __shfl_xor_sync(0xffffffffu, hi, m);
                const double sg = (lane & m) ? -1.0 : 1.0;
                hr = fma(sg, hr, orr);
                hi = fma(sg, hi, oi);
            }
            // lane 0: total; lane 2^q: sign by lane bit q
            if (lane == 0u) {
                s_z[warp][0][0] = hr;
                s_z[warp][0][1] = hi;
            }
#pragma unroll
            for (int q = 0; q < 5; ++q)
                if (lane == (1u << q)) {
                    s_z[warp][1 + q][0] = hr;
                    s_z[warp][1 + q][1] = hi;
                }
        }
        if (P.n_first < P.n_diag) {
            // generic diagonal generators (several parity bits, controls): with W[m] = sum_j (-1)^{popc(j & m)} t0_j
            //   sum over the j that have the control bits c set of (-1)^{popc(j & m)} t0_j = 2^-|c| sum_{s subset of c} (-1)^|s| W[m ^ s]
#pragma unroll
            for (int h = 1; h < GT_EPT; h <<= 1)
#pragma unroll
                for (int j = 0; j < GT_EPT; ++j)
                    if (!(j & h)) {
                        const double ar = Wr[j], ai = Wi[j], br = Wr[j | h], bi = Wi[j | h];
                        Wr[j] = ar + br;
                        Wi[j] = ai + bi;
                        Wr[j | h] = ar - br;
                        Wi[j | h] = ai - bi;
                    }
            for (int k0 = P.n_first; k0 < P.n_diag; k0 += 4) {
                const int count = min(4, P.n_diag - k0);
                for (int q = 0; q < count; ++q) {
                    const GenDesc &g = P.g[k0 + q];
                    const unsigned mj = g.jinfo & 15u, cjm = (g.jinfo >> 4) & 15u;
                    double re = 0.0, im = 0.0;
                    if ((g0 & g.ctrl) == g.ctrl) {
                        // S0 = sum over the controlled j of t0_j, Sm = the same with the parity sign
                        double s0r = 0.0, s0i = 0.0, smr = 0.0, smi = 0.0;
                        unsigned sub = cjm;
                        for (;;) {  // CTA-uniform loop over the subsets of the control mask (one iteration without controls on j)
                            const double sg = (__popc(sub) & 1) ? -1.0 : 1.0;
                            double wr, wi;
                            pick_w(Wr, Wi, sub, wr, wi);
                            s0r += sg * wr;
                            s0i += sg * wi;
                            pick_w(Wr, Wi, mj ^ sub, wr, wi);
                            smr += sg * wr;
                            smi += sg * wi;
                            if (sub == 0u) break;
                            sub = (sub - 1u) & cjm;
                        }
                        const double sc = 0.5 / (double)(1u << __popc(cjm));
                        const double Er = sc * (s0r + smr), Ei = sc * (s0i + smi);  // even parity among the j bits
                        const double Or = sc * (s0r - smr), Oi = sc * (s0i - smi);  // odd
                        const bool odd = __popcll(g0 & g.zmask) & 1;
                        const double ear = odd ? g.m[2] : g.m[0], eai = odd ? g.m[3] : g.m[1];
                        const double ebr = odd ? g.m[0] : g.m[2], ebi = odd ? g.m[1] : g.m[3];
                        re = ear * Er - eai * Ei + ebr * Or - ebi * Oi;
                        im = ear * Ei + eai * Er + ebr * Oi + ebi * Or;
                    }
                    put(q, re, im);
                }
                flush_group(k0, count);
            }
        }
    }
    __syncthreads();
    if ((int)tid < P.n_gens) {
        const GenDesc &g = P.g[tid];
        double re = 0.0, im = 0.0;
        if ((int)tid < P.n_diag) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                re += s_acc[tid][w][0];
                im += s_acc[tid][w][1];
            }
        } else {
            // value = e_even * (M + Z) / 2 + e_odd * (M - Z) / 2
            double Mr = 0.0, Mi = 0.0, Zr = 0.0, Zi = 0.0;
            const unsigned p = g.tbit;
            for (int w = 0; w < NW; ++w) {
                const double mr = s_z[w][0][0], mi = s_z[w][0][1];
                Mr += mr;
                Mi += mi;
                if (p < 5u) {
                    Zr += s_z[w][1 + p][0];
                    Zi += s_z[w][1 + p][1];
                } else if (p < (unsigned)(TB - GT_JB)) {
                    const double sg = ((unsigned)w >> (p - 5u)) & 1u ? -1.0 : 1.0;
                    Zr += sg * mr;
                    Zi += sg * mi;
                } else if (p < (unsigned)TB) {
                    Zr += s_z[w][6 + p - (TB - GT_JB)][0];
                    Zi += s_z[w][6 + p - (TB - GT_JB)][1];
                }
            }
            if (p == ZS_OUTSIDE) {
                const double sg = (base & g.xg) ? -1.0 : 1.0;
                Zr = sg * Mr;
                Zi = sg * Mi;
            } else if (p == ZS_NONE) {
                Zr = Mr;
                Zi = Mi;
            }
            const double Er = 0.5 * (Mr + Zr), Ei = 0.5 * (Mi + Zi), Or = 0.5 * (Mr - Zr), Oi = 0.5 * (Mi - Zi);
            re = g.m[0] * Er - g.m[1] * Ei + g.m[2] * Or - g.m[3] * Oi;
            im = g.m[0] * Ei + g.m[1] * Er + g.m[2] * Oi + g.m[3] * Or;
        }
        atomicAdd(out + 2 * (size_t)g.slot, re);
        atomicAdd(out + 2 * (size_t)g.slot + 1, im);
    }
}

// out = sum_t c_t P_t in   (+ out when `accumulate`)   for Pauli words whose X / Y letters all sit on tile bits: the tile of
// `in` is read once into shared memory, every thread accumulates its 16 output amplitudes in registers over all terms of
// the launch (one shared-memory read, a sign flip and four FP64 multiply-adds per term and amplitude) and writes them
// once.  Replaces the per-term gather of k_pauli_sum_apply (one global read per term and amplitude) when a Hamiltonian
// is applied to a vector -- the bra H|lambda> of the adjoint method (algorithms/ObservablesGPU.hpp:346-362 in the
// reference: per-term copy + apply + axpy).
template <typename T, int TB>
__global__ void __launch_bounds__(1 << (TB - GT_JB))
    k_pauli_sum_tile(const void *__restrict__ in, void *__restrict__ out, int accumulate, const __grid_constant__ GenProgram P) {
    using A = typename VecOf<T, 1>::type;
    constexpr int NT = 1 << (TB - GT_JB);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *s = reinterpret_cast<A *>(smem_raw);
    const uint64_t base = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    const uint32_t tid = threadIdx.x;
    uint64_t g0 = tid & ((1u << P.L) - 1u);
    for (int q = P.L; q < TB - GT_JB; ++q) g0 |= (uint64_t)((tid >> q) & 1u) << P.hi_bits[q - P.L];
    g0 |= base;
    uint64_t col[GT_JB];
#pragma unroll
    for (int q = 0; q < GT_JB; ++q) col[q] = 1ull << P.hi_bits[TB - GT_JB + q - P.L];
#pragma unroll
    for (int j = 0; j < GT_EPT; ++j) {
        uint64_t gi = g0;
#pragma unroll
        for (int q = 0; q < GT_JB; ++q)
            if ((j >> q) & 1) gi |= col[q];
        s[tid + (uint32_t)j * NT] = reinterpret_cast<const A *>(in)[gi];
    }
    __syncthreads();
    double yr[GT_EPT], yi[GT_EPT];
#pragma unroll
    for (int j = 0; j < GT_EPT; ++j) yr[j] = yi[j] = 0.0;
    for (int k = 0; k < P.n_gens; ++k) {
        const GenDesc &g = P.g[k];
        const unsigned mj = g.jinfo & 15u;
        // sign of amplitude i: (-1)^{popc((i ^ x) & z)} = thread part ^ j part; the j part is CTA-uniform
        const int t_flip = (int)(((__popcll((g0 ^ g.xg) & g.zmask) ^ (g.jinfo >> 8)) & 1u) << 31);
        const A *sp = s + (tid ^ (g.tbit & (uint32_t)(NT - 1)));
        const uint32_t jx = g.tbit >> (TB - GT_JB);
        const double cr = g.m[0], ci = g.m[1];
#pragma unroll
        for (int j = 0; j < GT_EPT; ++j) {
            const A x = sp[((uint32_t)j ^ jx) * NT];
            const int flip = t_flip ^ ((__popc((unsigned)j & mj) & 1) << 31);
            const double xr = __hiloint2double(__double2hiint((double)x.x) ^ flip, __double2loint((double)x.x));
            const double xi = __hiloint2double(__double2hiint((double)x.y) ^ flip, __double2loint((double)x.y));
            yr[j] = fma(cr, xr, yr[j]);
            yr[j] = fma(-ci, xi, yr[j]);
            yi[j] = fma(cr, xi, yi[j]);
            yi[j] = fma(ci, xr, yi[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < GT_EPT; ++j) {
        uint64_t gi = g0;
#pragma unroll
        for (int q = 0; q < GT_JB; ++q)
            if ((j >> q) & 1) gi |= col[q];
        A *o = reinterpret_cast<A *>(out) + gi;
        if (accumulate) {
            const A old = *o;
            yr[j] += (double)old.x;
            yi[j] += (double)old.y;
        }
        A y;
        y.x = (T)yr[j];
        y.y = (T)yi[j];
        *o = y;
    }
}

// which generators the tile kernel takes; everything else goes through launch_bra_op_ket one by one
bool tile_gen_kind(const LoweredGate &g, int n, int &kind) {
    if (g.kind == LoweredGate::PARITY) {
        kind = 0;
        return true;
    }
    if (g.kind == LoweredGate::DIAG && g.k <= 1) {
        kind = 0;
        return true;
    }
    if (g.kind == LoweredGate::DENSE && g.k == 1 && g.tgt_bits.size() == 1 && g.offs[0] == 0 && g.tgt_bits[0] < n) {
        kind = 1;
        return true;
    }
    return false;
}

template <typename T, int TB>
void launch_gens_t(State &sv, const void *bra, const void *ket, double *out, const GenProgram &P) {
    constexpr int NT = 1 << (TB - GT_JB);
    constexpr int NW = NT / 32;
    const size_t smem = ((size_t)1 << TB) * sizeof(typename VecOf<T, 1>::type) +
                        ((size_t)GT_MAX * NW * 2 + (size_t)NW * 32 * RED_PAD + (size_t)NW * (1 + 5 + GT_JB) * 2) * sizeof(double);
    const unsigned grid = (unsigned)(1ull << (sv.n - TB));
    static bool configured[64] = {false};  // per device: function attributes belong to the device's context
    auto kern = k_bra_gens_ket<T, TB>;
    if (!configured[sv.device & 63]) {
        QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[sv.device & 63] = true;
    }
    kern<<<grid, NT, smem, sv.stream>>>(bra, ket, out, P);
    QSV_CUDA(cudaGetLastError());
}

int gens_env(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

}  // namespace

// tile size of the batched generator kernel for an n-qubit vector (0: too small, generators go one by one) and its
// contiguous low bits: a 2^13 tile with 2 low bits takes 13 single-qubit generators on arbitrary qubits per launch and the
// 11 others in the next one, so a layer of rotations on 24 qubits needs two reads of (bra, ket)
int gens_tile_bits(int n) {
    static const int want = std::max(12, std::min(gens_env("QSV_GENS_TB", 13), GT_TB_MAX));
    return n >= want ? want : (n >= 12 ? 12 : 0);
}
int gens_low_bits(int tb) {
    static const int l13 = std::max(1, std::min(gens_env("QSV_GENS_L13", 2), 8));
    static const int l12 = std::max(1, std::min(gens_env("QSV_GENS_L12", 4), 8));
    return tb == 13 ? l13 : l12;
}

namespace {

// a generator / Pauli word waiting for a tile: `need` = non-diagonal bits >= L that must be tile bits; tgt / xg are
// global positions that are translated to tile-local ones once the tile is chosen
struct GenItem {
    GenDesc d;
    uint64_t need = 0;
    int tgt = -1;  // kind 1
};

template <typename T, int TB>
void launch_sum_tile_t(State &sv, const void *in, void *out, bool accumulate, const GenProgram &P) {
    constexpr int NT = 1 << (TB - GT_JB);
    const size_t smem = ((size_t)1 << TB) * sizeof(typename VecOf<T, 1>::type);
    const unsigned grid = (unsigned)(1ull << (sv.n - TB));
    static bool configured[64] = {false};
    auto kern = k_pauli_sum_tile<T, TB>;
    if (!configured[sv.device & 63]) {
        QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[sv.device & 63] = true;
    }
    kern<<<grid, NT, smem, sv.stream>>>(in, out, accumulate ? 1 : 0, P);
    QSV_CUDA(cudaGetLastError());
}

// apply_out != nullptr: the items are Pauli words with coefficients and the launches are k_pauli_sum_tile (out = / +=)
void run_items(State &sv, const void *bra, const void *ket, std::vector<GenItem> &todo, double *out_dev,
               void *apply_out = nullptr, bool *apply_accumulate = nullptr) {
    const int n = sv.n;
    // a handful of generators: the 2^12 tile (two CTAs per SM, one loading while the other computes) beats the 2^13 one,
    // whose point is to fit more generators into a launch
    static const int small_tb = gens_env("QSV_GENS_SMALL_TB", 1);
    bool small = small_tb && todo.size() <= 6 && n >= 12 && gens_tile_bits(n) > 12;
    for (const GenItem &it : todo)
        small = small && __builtin_popcountll(it.need & ~((1ull << gens_low_bits(12)) - 1ull)) <= 12 - gens_low_bits(12);
    const int TB = small ? 12 : gens_tile_bits(n);
    const int L = gens_low_bits(TB);
    const int max_hi = TB - L;
    const uint64_t low = (1ull << L) - 1ull;
    while (!todo.empty()) {
        // one launch: up to GT_MAX items whose non-diagonal bits >= L number at most max_hi (first fit over all)
        GenProgram P;
        memset(&P, 0, sizeof(P));
        P.L = L;
        uint64_t need = 0;
        std::vector<GenItem> rest, take;
        for (GenItem &it : todo) {
            if ((int)take.size() < GT_MAX && __builtin_popcountll((need | it.need) & ~low) <= max_hi) {
                need |= it.need & ~low;
                take.push_back(it);
            } else {
                rest.push_back(it);
            }
        }
        QSV_CHECK(!take.empty(), "internal: a generator does not fit the tile");
        // Z sums: diagonal generators without controls whose parity mask has at most one bit
        for (GenItem &it : take)
            if (it.d.kind == 0 && it.d.ctrl == 0 && __builtin_popcountll(it.d.zmask) <= 1) it.d.kind = 3;
        // off-diagonal generators first (they need bra, whose registers the diagonal part then reuses), Z sums last
        auto rank_of = [](const GenItem &a) { return a.d.kind == 3 ? 2 : (a.d.kind == 0 ? 1 : 0); };
        std::stable_sort(take.begin(), take.end(), [&](const GenItem &a, const GenItem &b) { return rank_of(a) < rank_of(b); });
        std::vector<int> hi;
        for (int b = L; b < n; ++b)
            if (need >> b & 1) hi.push_back(b);
        for (int b = L; b < n && (int)hi.size() < max_hi; ++b)
            if (!(need >> b & 1)) hi.push_back(b);
        std::sort(hi.begin(), hi.end());
        int pos[64];
        for (int b = 0; b < 64; ++b) pos[b] = -1;
        std::vector<int> tile_bits;
        for (int b = 0; b < L; ++b) {
            pos[b] = b;
            tile_bits.push_back(b);
        }
        for (int j = 0; j < (int)hi.size(); ++j) {
            pos[hi[j]] = L + j;
            P.hi_bits[j] = (unsigned char)hi[j];
            tile_bits.push_back(hi[j]);
        }
        P.tile_holes = make_holes(tile_bits.data(), (int)tile_bits.size(), 0);
        for (GenItem &it : take) {
            GenDesc &d = P.g[P.n_gens++];
            d = it.d;
            if (d.kind == 1 || d.kind == 2) P.n_first = P.n_gens;
            if (d.kind != 3) P.n_diag = P.n_gens;
            if (d.kind == 3) {
                d.xg = d.zmask;
                d.tbit = d.zmask == 0 ? ZS_NONE : ZS_OUTSIDE;
                for (int b = 0; b < n; ++b)
                    if ((d.zmask >> b & 1) && pos[b] >= 0) d.tbit = (unsigned)pos[b];
                d.zmask = 0;
                d.jinfo = 0;
                continue;
            }
            if (d.kind == 1) {
                d.tbit = (unsigned)pos[it.tgt];
            } else if (d.kind == 2) {
                unsigned xl = 0;
                for (int b = 0; b < n; ++b)
                    if (d.xg >> b & 1) xl |= 1u << pos[b];
                d.tbit = xl;
            }
            // the thread's GT_JB amplitude bits j are the top tile bits (e = tid + NT * j): what the generator sees of them
            unsigned mj = 0, cj = 0, xj = 0;
            for (int q = 0; q < GT_JB; ++q) {
                const int gbit = tile_bits[TB - GT_JB + q];
                mj |= (unsigned)((d.zmask >> gbit) & 1ull) << q;
                cj |= (unsigned)((d.ctrl >> gbit) & 1ull) << q;
                xj |= (unsigned)((d.xg >> gbit) & 1ull) << q;
                d.zmask &= ~(1ull << gbit);
                d.ctrl &= ~(1ull << gbit);
            }
            const unsigned flip0 = d.kind == 2 ? (unsigned)(__builtin_popcount(xj & mj) & 1) : 0u;
            d.jinfo = mj | (cj << 4) | (flip0 << 8);
        }
        sv.stat_launches += 1;
        if (apply_out) {
            if (sv.dtype == QSV_C128) {
                if (TB == 13)
                    launch_sum_tile_t<double, 13>(sv, ket, apply_out, *apply_accumulate, P);
                else
                    launch_sum_tile_t<double, 12>(sv, ket, apply_out, *apply_accumulate, P);
            } else {
                if (TB == 13)
                    launch_sum_tile_t<float, 13>(sv, ket, apply_out, *apply_accumulate, P);
                else
                    launch_sum_tile_t<float, 12>(sv, ket, apply_out, *apply_accumulate, P);
            }
            *apply_accumulate = true;
            todo.swap(rest);
            continue;
        }
        if (sv.dtype == QSV_C128) {
            if (TB == 13)
                launch_gens_t<double, 13>(sv, bra, ket, out_dev, P);
            else
                launch_gens_t<double, 12>(sv, bra, ket, out_dev, P);
        } else {
            if (TB == 13)
                launch_gens_t<float, 13>(sv, bra, ket, out_dev, P);
            else
                launch_gens_t<float, 12>(sv, bra, ket, out_dev, P);
        }
        todo.swap(rest);
    }
}

}  // namespace

// out_dev[2 * slots[k] ..] += <bra| gens[k] |ket> (re, im) for every k, with as few reads of the vectors as the
// tile size allows.  Generators are LoweredGates used as operators (controls = projectors).
void launch_bra_gens_ket(State &sv, const void *bra, const void *ket, const std::vector<LoweredGate> &gens,
                         const std::vector<int> &slots, double *out_dev) {
    sv.use();
    const int n = sv.n;
    std::vector<GenItem> todo;
    for (size_t k = 0; k < gens.size(); ++k) {
        const LoweredGate &g = gens[k];
        int kind;
        if (g.kind == LoweredGate::NOP) continue;  // a projector that is zero on this shard
        if (!(gens_tile_bits(n) > 0 && tile_gen_kind(g, n, kind))) {
            launch_bra_op_ket(sv, bra, ket, g, out_dev, slots[k]);
            continue;
        }
        GenItem it;
        memset(&it.d, 0, sizeof(it.d));
        it.d.kind = kind;
        it.d.slot = slots[k];
        it.d.ctrl = g.ctrl_mask;
        if (kind == 1) {
            it.tgt = g.tgt_bits[0];
            it.need = 1ull << it.tgt;
            const cplx zero(0.0, 0.0), one(1.0, 0.0), pi_(0.0, 1.0), mi_(0.0, -1.0);
            const bool is_x = g.mat[0] == zero && g.mat[1] == one && g.mat[2] == one && g.mat[3] == zero;
            const bool is_y = g.mat[0] == zero && g.mat[1] == mi_ && g.mat[2] == pi_ && g.mat[3] == zero;
            if (is_x || is_y) {
                // a Pauli X / Y generator (RX, RY, CRX, CRY ...) is a one-letter Pauli word: one shared-memory read and six
                // FP64 operations per amplitude instead of two reads and twelve
                it.d.kind = 2;
                it.d.xg = 1ull << it.tgt;
                it.d.zmask = is_y ? 1ull << it.tgt : 0;
                it.d.m[0] = is_y ? 0.0 : 1.0;
                it.d.m[1] = is_y ? 1.0 : 0.0;
                it.tgt = -1;
            } else {
                for (int q = 0; q < 4; ++q) {
                    it.d.m[2 * q] = g.mat[q].real();
                    it.d.m[2 * q + 1] = g.mat[q].imag();
                }
            }
        } else if (g.kind == LoweredGate::PARITY) {
            it.d.zmask = g.zmask;
            it.d.m[0] = g.mat[0].real();
            it.d.m[1] = g.mat[0].imag();
            it.d.m[2] = g.mat[1].real();
            it.d.m[3] = g.mat[1].imag();
        } else {  // DIAG with 0 or 1 table bits
            it.d.zmask = g.k == 1 ? (1ull << g.tgt_bits[0]) : 0;
            it.d.m[0] = g.mat[0].real();
            it.d.m[1] = g.mat[0].imag();
            it.d.m[2] = g.k == 1 ? g.mat[1].real() : g.mat[0].real();
            it.d.m[3] = g.k == 1 ? g.mat[1].imag() : g.mat[0].imag();
        }
        todo.push_back(it);
    }
    run_items(sv, bra, ket, todo, out_dev);
}

// out_dev[2 * (first_slot + t) ..] += <bra| P_t |ket> for Pauli words given by (x, z, #Y) masks: single-pass fused
// expectation values of many words -- every word whose X/Y letters fit one tile shares one read of the vectors
// (replaces custatevecComputeExpectationsOnPauliBasis, Managed.hpp:1117-1128)
void launch_bra_paulis_ket(State &sv, const void *bra, const void *ket, int n_terms, const uint64_t *xmasks,
                           const uint64_t *zmasks, const int *nys, int first_slot, double *out_dev) {
    sv.use();
    const int n = sv.n;
    const int TB = gens_tile_bits(n);
    const uint64_t low = (1ull << gens_low_bits(TB)) - 1ull;
    std::vector<GenItem> todo;
    for (int t = 0; t < n_terms; ++t) {
        const uint64_t need = xmasks[t];
        if (TB == 0 || __builtin_popcountll(need & ~low) > TB - gens_low_bits(TB)) {
            launch_bra_pauli_ket(sv, bra, ket, xmasks[t], zmasks[t], nys[t], out_dev, first_slot + t);
            continue;
        }
        GenItem it;
        memset(&it.d, 0, sizeof(it.d));
        it.d.kind = 2;
        it.d.slot = first_slot + t;
        it.d.xg = xmasks[t];
        it.d.zmask = zmasks[t];
        const double ph[4][2] = {{1, 0}, {0, 1}, {-1, 0}, {0, -1}};
        it.d.m[0] = ph[nys[t] & 3][0];
        it.d.m[1] = ph[nys[t] & 3][1];
        it.need = need;
        todo.push_back(it);
    }
    run_items(sv, bra, ket, todo, out_dev);
}

// out (+)= sum_t coeffs[t] * P_t in through the tile kernel; returns false when the vector is too small for a tile or a
// word's X / Y letters do not fit one (the caller then uses the per-term gather kernel)
bool launch_pauli_sum_apply_tiled(State &sv, const void *in, void *out, int n_terms, const uint64_t *xmasks,
                                  const uint64_t *zmasks, const cplx *coeffs, bool accumulate) {
    sv.use();
    const int n = sv.n;
    const int TB = gens_tile_bits(n);
    if (TB == 0 || n_terms < 2) return false;
    const int L = gens_low_bits(TB);
    const uint64_t low = (1ull << L) - 1ull;
    std::vector<GenItem> todo;
    for (int t = 0; t < n_terms; ++t) {
        if (__builtin_popcountll(xmasks[t] & ~low) > TB - L) return false;
        GenItem it;
        memset(&it.d, 0, sizeof(it.d));
        it.d.kind = 2;
        it.d.xg = xmasks[t];
        it.d.zmask = zmasks[t];
        it.d.m[0] = coeffs[t].real();
        it.d.m[1] = coeffs[t].imag();
        it.need = xmasks[t];
        todo.push_back(it);
    }
    bool acc = accumulate;
    run_items(sv, nullptr, in, todo, nullptr, out, &acc);
    return true;
}

}  // namespace qsv
