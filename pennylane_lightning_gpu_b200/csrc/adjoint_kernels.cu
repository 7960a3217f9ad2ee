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

constexpr int GT_TB = 11;        // tile bits: 2^11 amplitudes of ket in shared memory (32 KiB complex128)
constexpr int GT_NT = 256;
constexpr int GT_EPT = (1 << GT_TB) / GT_NT;  // 8 amplitudes per thread
static_assert(GT_EPT == 8 && GT_NT == 256, "the kernel's index split assumes e = tid + 256 * j with j < 8");
constexpr int GT_MAX = 32;       // generators per launch

struct GenDesc {
    int kind;            // 0: diagonal table (<= 1 table bit: parity of index & zmask), 1: 2x2 block on one tile bit,
                         // 2: Pauli word (X/Y letters on tile bits, Z letters anywhere)
    int slot;            // complex output slot
    unsigned tbit;       // kind 1: tile-local position of the target bit; kind 2: tile-local flip mask
    unsigned pad0;       // kind 0: bits 0-2 = parity mask on the thread's three amplitude-index bits (tile bits 8..10), bit 8 =
                         // a control sits on one of them (generic per-amplitude path)
    uint64_t ctrl;       // global bits that must be 1
    uint64_t zmask;      // kind 0 / 2: global bits whose parity selects the phase / the sign
    uint64_t xg;         // kind 2: flip mask on global bits
    double m[8];         // kind 1: row-major complex 2x2; kind 0: phases for even / odd parity; kind 2: i^ny
};

struct GenProgram {
    int n_gens;
    int L;
    unsigned char hi_bits[16];  // global positions of tile bits L..GT_TB-1
    Holes tile_holes;
    GenDesc g[GT_MAX];
};

template <typename T>
__global__ void __launch_bounds__(GT_NT)
    k_bra_gens_ket(const void *__restrict__ bra, const void *__restrict__ ket, double *out,
                   const __grid_constant__ GenProgram P) {
    using A = typename VecOf<T, 1>::type;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *s = reinterpret_cast<A *>(smem_raw);
    __shared__ double s_acc[GT_MAX][2];
    const uint64_t base = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    const int tid = threadIdx.x, lane = tid & 31;
    for (int i = tid; i < 2 * GT_MAX; i += GT_NT) (&s_acc[0][0])[i] = 0.0;

    // element e of the tile (local index) lives at global index base | deposit(e)
    // (e = tid + 256 * j: the thread part of the deposit is computed once, the three j bits are CTA-uniform columns)
    A b[GT_EPT];
    uint64_t gidx[GT_EPT];
    uint64_t dep_tid = (uint32_t)tid & ((1u << P.L) - 1u);
    for (int q = P.L; q < GT_TB - 3; ++q) dep_tid |= (uint64_t)(((uint32_t)tid >> q) & 1u) << P.hi_bits[q - P.L];
    dep_tid |= base;
    const uint64_t cj0 = 1ull << P.hi_bits[GT_TB - 3 - P.L], cj1 = 1ull << P.hi_bits[GT_TB - 2 - P.L],
                   cj2 = 1ull << P.hi_bits[GT_TB - 1 - P.L];
#pragma unroll
    for (int j = 0; j < GT_EPT; ++j) {
        const uint32_t e = (uint32_t)tid + (uint32_t)j * GT_NT;
        gidx[j] = dep_tid | ((j & 1) ? cj0 : 0ull) | ((j & 2) ? cj1 : 0ull) | ((j & 4) ? cj2 : 0ull);
        s[e] = reinterpret_cast<const A *>(ket)[gidx[j]];
        b[j] = reinterpret_cast<const A *>(bra)[gidx[j]];
    }
    __syncthreads();

    // Diagonal / parity generators: conj(b_i) * p(i) * k_i with t0_i = conj(b_i) * k_i shared by all of them.  The thread's
    // 8 amplitudes differ in tile bits 8..10 (j = 0..7), everything else of the index is thread-constant, so
    //   sum_j p(parity_j) t0_j = e_a * (W[0] + W[m]) / 2 + e_b * (W[0] - W[m]) / 2,   W[m] = sum_j (-1)^{popc(j & m)} t0_j
    // with m = the generator's parity mask restricted to the three j bits and (e_a, e_b) = the two phases ordered by the
    // thread-constant part of the parity: one Walsh-Hadamard transform per thread, then ~12 FP64 operations per generator
    // instead of ~8 per generator AND amplitude.
    double Wr[GT_EPT], Wi[GT_EPT];
#pragma unroll
    for (int j = 0; j < GT_EPT; ++j) {
        const A x = s[(uint32_t)tid + (uint32_t)j * GT_NT];
        Wr[j] = (double)b[j].x * (double)x.x + (double)b[j].y * (double)x.y;
        Wi[j] = (double)b[j].x * (double)x.y - (double)b[j].y * (double)x.x;
    }
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

    for (int k = 0; k < P.n_gens; ++k) {
        const GenDesc &g = P.g[k];
        double re = 0.0, im = 0.0;
        if (g.kind == 0 && (g.pad0 >> 8) == 0u) {
            // no control on a j bit: the control test is thread-constant too (gidx[0] has the j bits clear)
            if ((gidx[0] & g.ctrl) == g.ctrl) {
                const unsigned m = g.pad0 & 7u;
                double wr, wi;
                switch (m) {  // CTA-uniform
                case 0: wr = Wr[0]; wi = Wi[0]; break;
                case 1: wr = Wr[1]; wi = Wi[1]; break;
                case 2: wr = Wr[2]; wi = Wi[2]; break;
                case 3: wr = Wr[3]; wi = Wi[3]; break;
                case 4: wr = Wr[4]; wi = Wi[4]; break;
                case 5: wr = Wr[5]; wi = Wi[5]; break;
                case 6: wr = Wr[6]; wi = Wi[6]; break;
                default: wr = Wr[7]; wi = Wi[7]; break;
                }
                const bool odd = __popcll(gidx[0] & g.zmask) & 1;
                const double ear = odd ? g.m[2] : g.m[0], eai = odd ? g.m[3] : g.m[1];
                const double ebr = odd ? g.m[0] : g.m[2], ebi = odd ? g.m[1] : g.m[3];
                const double Ar = 0.5 * (Wr[0] + wr), Ai = 0.5 * (Wi[0] + wi);
                const double Br = 0.5 * (Wr[0] - wr), Bi = 0.5 * (Wi[0] - wi);
                re = ear * Ar - eai * Ai + ebr * Br - ebi * Bi;
                im = ear * Ai + eai * Ar + ebr * Bi + ebi * Br;
            }
        } else if (g.kind == 0) {
            const double e0r = g.m[0], e0i = g.m[1], e1r = g.m[2], e1i = g.m[3];
#pragma unroll
            for (int j = 0; j < GT_EPT; ++j) {
                if ((gidx[j] & g.ctrl) != g.ctrl) continue;
                const bool odd = __popcll(gidx[j] & g.zmask) & 1;
                const double pr = odd ? e1r : e0r, pi = odd ? e1i : e0i;
                const A x = s[(uint32_t)tid + (uint32_t)j * GT_NT];
                const double yr = pr * (double)x.x - pi * (double)x.y, yi = pr * (double)x.y + pi * (double)x.x;
                re += (double)b[j].x * yr + (double)b[j].y * yi;
                im += (double)b[j].x * yi - (double)b[j].y * yr;
            }
        } else if (g.kind == 2) {
            // (P ket)_i = i^ny * (-1)^{popc(j & z)} * ket_j with j = i ^ x (optionally only where the control bits are set:
            // generators |1><1| (x) X / Y of controlled rotations); the constant i^ny is applied after the sum
#pragma unroll
            for (int j = 0; j < GT_EPT; ++j) {
                if ((gidx[j] & g.ctrl) != g.ctrl) continue;
                const uint32_t e = (uint32_t)tid + (uint32_t)j * GT_NT;
                const A x = s[e ^ g.tbit];
                const bool odd = __popcll((gidx[j] ^ g.xg) & g.zmask) & 1;
                const double ur = (double)b[j].x * (double)x.x + (double)b[j].y * (double)x.y;
                const double ui = (double)b[j].x * (double)x.y - (double)b[j].y * (double)x.x;
                re += odd ? -ur : ur;
                im += odd ? -ui : ui;
            }
            const double cr = g.m[0], ci = g.m[1], r0 = re;
            re = cr * r0 - ci * im;
            im = cr * im + ci * r0;
        } else {
            const uint32_t bit = 1u << g.tbit;
            const double m0 = g.m[0], m1 = g.m[1], m2 = g.m[2], m3 = g.m[3];
            const double m4 = g.m[4], m5 = g.m[5], m6 = g.m[6], m7 = g.m[7];
#pragma unroll
            for (int j = 0; j < GT_EPT; ++j) {
                if ((gidx[j] & g.ctrl) != g.ctrl) continue;
                const uint32_t e = (uint32_t)tid + (uint32_t)j * GT_NT;
                const A x0 = s[e & ~bit], x1 = s[e | bit];
                const bool hi = (e & bit) != 0;
                const double ar = hi ? m4 : m0, ai = hi ? m5 : m1, br = hi ? m6 : m2, bi = hi ? m7 : m3;
                const double yr = ar * (double)x0.x - ai * (double)x0.y + br * (double)x1.x - bi * (double)x1.y;
                const double yi = ar * (double)x0.y + ai * (double)x0.x + br * (double)x1.y + bi * (double)x1.x;
                re += (double)b[j].x * yr + (double)b[j].y * yi;
                im += (double)b[j].x * yi - (double)b[j].y * yr;
            }
        }
        re = warp_sum(re);
        im = warp_sum(im);
        if (lane == 0) {
            atomicAdd(&s_acc[k][0], re);
            atomicAdd(&s_acc[k][1], im);
        }
    }
    __syncthreads();
    if (tid < P.n_gens) {
        atomicAdd(out + 2 * (size_t)P.g[tid].slot, s_acc[tid][0]);
        atomicAdd(out + 2 * (size_t)P.g[tid].slot + 1, s_acc[tid][1]);
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

template <typename T> void launch_gens_t(State &sv, const void *bra, const void *ket, double *out, const GenProgram &P) {
    const size_t smem = ((size_t)1 << GT_TB) * sizeof(typename VecOf<T, 1>::type);
    const unsigned grid = (unsigned)(1ull << (sv.n - GT_TB));
    k_bra_gens_ket<T><<<grid, GT_NT, smem, sv.stream>>>(bra, ket, out, P);
    QSV_CUDA(cudaGetLastError());
}

}  // namespace

namespace {

// a generator / Pauli word waiting for a tile: `need` = non-diagonal bits >= L that must be tile bits; tgt / xg are
// global positions that are translated to tile-local ones once the tile is chosen
struct GenItem {
    GenDesc d;
    uint64_t need = 0;
    int tgt = -1;  // kind 1
};

constexpr int GT_L = 4;

void run_items(State &sv, const void *bra, const void *ket, std::vector<GenItem> &todo, double *out_dev) {
    const int n = sv.n;
    const int max_hi = GT_TB - GT_L;
    while (!todo.empty()) {
        // one launch: up to GT_MAX items whose non-diagonal bits >= L number at most max_hi (first fit over all)
        GenProgram P;
        memset(&P, 0, sizeof(P));
        P.L = GT_L;
        uint64_t need = 0;
        std::vector<GenItem> rest, take;
        for (GenItem &it : todo) {
            if ((int)take.size() < GT_MAX && __builtin_popcountll(need | it.need) <= max_hi) {
                need |= it.need;
                take.push_back(it);
            } else {
                rest.push_back(it);
            }
        }
        std::vector<int> hi;
        for (int b = GT_L; b < n; ++b)
            if (need >> b & 1) hi.push_back(b);
        for (int b = GT_L; b < n && (int)hi.size() < max_hi; ++b)
            if (!(need >> b & 1)) hi.push_back(b);
        std::sort(hi.begin(), hi.end());
        int pos[64];
        for (int b = 0; b < 64; ++b) pos[b] = -1;
        std::vector<int> tile_bits;
        for (int b = 0; b < GT_L; ++b) {
            pos[b] = b;
            tile_bits.push_back(b);
        }
        for (int j = 0; j < (int)hi.size(); ++j) {
            pos[hi[j]] = GT_L + j;
            P.hi_bits[j] = (unsigned char)hi[j];
            tile_bits.push_back(hi[j]);
        }
        P.tile_holes = make_holes(tile_bits.data(), (int)tile_bits.size(), 0);
        static_assert(GT_TB - 3 >= GT_L, "the three per-thread amplitude bits are high tile bits");
        for (GenItem &it : take) {
            GenDesc &d = P.g[P.n_gens++];
            d = it.d;
            if (d.kind == 0) {
                // tile bits GT_TB-3 .. GT_TB-1 distinguish the 8 amplitudes of a thread (e = tid + 256 * j)
                unsigned m = 0, cj = 0;
                for (int q = 0; q < 3; ++q) {
                    const int gbit = tile_bits[GT_TB - 3 + q];
                    m |= (unsigned)((d.zmask >> gbit) & 1ull) << q;
                    cj |= (unsigned)((d.ctrl >> gbit) & 1ull);
                }
                d.pad0 = m | (cj << 8);
            }
            if (d.kind == 1) {
                d.tbit = (unsigned)pos[it.tgt];
            } else if (d.kind == 2) {
                unsigned xl = 0;
                for (int b = 0; b < n; ++b)
                    if (d.xg >> b & 1) xl |= 1u << pos[b];
                d.tbit = xl;
            }
        }
        sv.stat_launches += 1;
        if (sv.dtype == QSV_C128)
            launch_gens_t<double>(sv, bra, ket, out_dev, P);
        else
            launch_gens_t<float>(sv, bra, ket, out_dev, P);
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
        if (!(n >= GT_TB && tile_gen_kind(g, n, kind))) {
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
            if (it.tgt >= GT_L) it.need = 1ull << it.tgt;
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
    const uint64_t low = (1ull << GT_L) - 1ull;
    std::vector<GenItem> todo;
    for (int t = 0; t < n_terms; ++t) {
        const uint64_t need = xmasks[t] & ~low;
        if (n < GT_TB || __builtin_popcountll(need) > GT_TB - GT_L) {
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

}  // namespace qsv
