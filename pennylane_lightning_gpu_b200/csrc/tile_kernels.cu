// Fused circuit execution: many gates per HBM sweep through a shared-memory tile.
//
// This is north_star items (b) + (c): high-target-qubit gates are served by staging, with TMA bulk
// copies (cp.async.bulk + mbarrier), the 2^(TB-L) strided runs that contain all partner amplitudes
// into shared memory, i.e. an index-bit permutation done by the copy engine: local bit j >= L of the
// tile is global bit hi_bits[j-L].  Once a tile is resident, EVERY gate whose non-diagonal target
// bits lie inside the tile is applied to it before it is written back, so a run of gates costs one
// read + one write of the state instead of one per gate.  Controls and diagonal gates may sit on ANY
// bit: bits outside the tile are constant per tile and become a per-CTA predicate / table offset
// (this is also how gates controlled by, or diagonal on, the global qubits of a sharded state run
// without communication).
//
// The reference has no counterpart: it issues one custatevecApplyMatrix per gate
// (simulator/StateVectorCudaManaged.hpp:1433-1471), i.e. one full sweep each, from a Python loop
// (lightning_gpu.py:519-555).
//
// Tile: 64 KiB (2^12 complex128 or 2^13 complex64), 256 threads, 3 CTAs / SM; the loads of one CTA
// overlap the arithmetic and stores of its neighbours, so no intra-CTA pipeline is needed.
#include <algorithm>
#include <mutex>
#include <cstdlib>
#include <map>

#include "device_utils.cuh"
#include "qsv_internal.h"

namespace qsv {

namespace {

constexpr int TILE_NT = 256;
constexpr int TILE_MAX_GATES = 40;
constexpr int TILE_POOL = 1280;  // doubles

enum : unsigned char { TG_DENSE1 = 1, TG_DENSE2 = 2, TG_DIAG = 3, TG_PARITY = 4, TG_SWAP1 = 5, TG_REAL1 = 6 };

struct TileGate {
    unsigned char kind;
    unsigned char n_holes;  // dense: local holes (targets + local controls)
    unsigned char k;        // diag: table bits
    unsigned char pad0;
    uint32_t loc_ctrl;      // control bits inside the tile (local positions)
    uint64_t out_ctrl;      // control bits outside the tile (global positions)
    unsigned char holes[8]; // local hole positions, ascending
    uint16_t offs[4];       // dense: local offsets of the group members
    unsigned char dsrc[4];  // diag: table-bit sources, MSB first; < 64 local position, >= 64 global bit + 64
    uint32_t zmask_loc;     // parity
    uint32_t mat_off;       // first double of this gate in the pool
    uint64_t zmask_out;
    // dense: expand_local(it * TILE_NT) for it = 0..15, so that group (tid + it * TILE_NT) sits at
    // expand_local(tid) | hi_tab[it]  (bit deposit is linear over disjoint bit sets)
    uint16_t hi_tab[16];
};

struct TileProgram {
    int tb;        // tile bits
    int L;         // low contiguous bits
    int n_gates;
    int pad0;
    uint64_t index_hi;           // value of the index bits above the local shard
    unsigned char hi_bits[16];   // global positions of local bits L..tb-1
    Holes tile_holes;            // all tile bits, ascending (expands blockIdx.x to the tile base)
    TileGate gates[TILE_MAX_GATES];
    double pool[TILE_POOL];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t mbar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(mbar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(mbar)
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, uint32_t src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes)
                 : "memory");
}

__device__ __forceinline__ uint32_t expand_local(uint32_t o, const unsigned char *holes, int n) {
    for (int j = 0; j < n; ++j) {
        const int p = holes[j];
        o = ((o >> p) << (p + 1)) | (o & ((1u << p) - 1u));
    }
    return o;
}

template <typename T> struct Cx;
template <> struct Cx<double> { using type = double2; };
template <> struct Cx<float> { using type = float2; };

template <typename T, bool USE_TMA>
__global__ void __launch_bounds__(TILE_NT, 3)
    k_tile_sweep(void *single, void *const *table, const __grid_constant__ TileProgram P) {
    using A = typename Cx<T>::type;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar_storage;
    A *s = reinterpret_cast<A *>(smem_raw);
    char *gbase = reinterpret_cast<char *>(table ? table[blockIdx.y] : single);
    const uint64_t base = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    const int n_runs = 1 << (P.tb - P.L);
    const uint32_t run_amps = 1u << P.L;
    const uint32_t run_bytes = run_amps * (uint32_t)sizeof(A);
    const int lane = threadIdx.x & 31;

    // ---- stage the tile: run r holds the amplitudes whose tile bits L.. equal r --------------------
    if constexpr (USE_TMA) {
        const uint32_t mbar = smem_u32(&mbar_storage);
        if (threadIdx.x == 0) mbar_init(mbar, 1);
        __syncthreads();
        if (threadIdx.x < 32) {
            if (lane == 0) mbar_expect_tx(mbar, (uint32_t)n_runs * run_bytes);
            __syncwarp();
            for (int r = lane; r < n_runs; r += 32) {
                uint64_t off = 0;
                for (int j = 0; j < P.tb - P.L; ++j) off |= (uint64_t)((r >> j) & 1) << P.hi_bits[j];
                bulk_g2s(smem_u32(smem_raw) + (uint32_t)r * run_bytes, gbase + (base | off) * sizeof(A), run_bytes, mbar);
            }
        }
        while (!mbar_try_wait(mbar, 0)) {
        }
    } else {
        const uint32_t vec_per_run = run_bytes / 16;
        const uint32_t total = (uint32_t)n_runs * vec_per_run;
        for (uint32_t v = threadIdx.x; v < total; v += TILE_NT) {
            const uint32_t r = v / vec_per_run, w = v % vec_per_run;
            uint64_t off = 0;
            for (int j = 0; j < P.tb - P.L; ++j) off |= (uint64_t)((r >> j) & 1) << P.hi_bits[j];
            reinterpret_cast<int4 *>(smem_raw)[v] =
                *reinterpret_cast<const int4 *>(gbase + (base | off) * sizeof(A) + (size_t)w * 16);
        }
        __syncthreads();
    }

    // ---- apply the program ---------------------------------------------------------------------
    const uint64_t outside = base | P.index_hi;
    const uint32_t tile_amps = 1u << P.tb;
    for (int gi = 0; gi < P.n_gates; ++gi) {
        const TileGate &g = P.gates[gi];
        if ((outside & g.out_ctrl) == g.out_ctrl) {  // CTA-uniform
            const double *mp = P.pool + g.mat_off;
            if (g.kind == TG_DENSE1 || g.kind == TG_SWAP1 || g.kind == TG_REAL1) {
                const uint32_t n_groups = tile_amps >> g.n_holes;
                const uint32_t o0 = g.offs[0], o1 = g.offs[1];
                const bool active = threadIdx.x < n_groups;
                const uint32_t e_tid = expand_local(threadIdx.x, g.holes, g.n_holes) | g.loc_ctrl;
                const int iters = n_groups >= TILE_NT ? (int)(n_groups / TILE_NT) : 1;
                if (active) {
                    if (g.kind == TG_SWAP1) {
                        // X-type gate (PauliX / CNOT / Toffoli / SWAP / CSWAP): a pure exchange, no FP
#pragma unroll 4
                        for (int it = 0; it < iters; ++it) {
                            const uint32_t i0 = e_tid | g.hi_tab[it];
                            const A a = s[i0 + o0], b = s[i0 + o1];
                            s[i0 + o0] = b;
                            s[i0 + o1] = a;
                        }
                    } else if (g.kind == TG_REAL1) {
                        // real 2x2 (RY, Hadamard, ...): half the FP64 work
                        const T m00 = (T)mp[0], m01 = (T)mp[2], m10 = (T)mp[4], m11 = (T)mp[6];
#pragma unroll 4
                        for (int it = 0; it < iters; ++it) {
                            const uint32_t i0 = e_tid | g.hi_tab[it];
                            const A a = s[i0 + o0], b = s[i0 + o1];
                            A x, y;
                            x.x = m00 * a.x + m01 * b.x;
                            x.y = m00 * a.y + m01 * b.y;
                            y.x = m10 * a.x + m11 * b.x;
                            y.y = m10 * a.y + m11 * b.y;
                            s[i0 + o0] = x;
                            s[i0 + o1] = y;
                        }
                    } else {
                        const T m00r = (T)mp[0], m00i = (T)mp[1], m01r = (T)mp[2], m01i = (T)mp[3];
                        const T m10r = (T)mp[4], m10i = (T)mp[5], m11r = (T)mp[6], m11i = (T)mp[7];
#pragma unroll 4
                        for (int it = 0; it < iters; ++it) {
                            const uint32_t i0 = e_tid | g.hi_tab[it];
                            const A a = s[i0 + o0], b = s[i0 + o1];
                            A x, y;
                            x.x = m00r * a.x - m00i * a.y + m01r * b.x - m01i * b.y;
                            x.y = m00r * a.y + m00i * a.x + m01r * b.y + m01i * b.x;
                            y.x = m10r * a.x - m10i * a.y + m11r * b.x - m11i * b.y;
                            y.y = m10r * a.y + m10i * a.x + m11r * b.y + m11i * b.x;
                            s[i0 + o0] = x;
                            s[i0 + o1] = y;
                        }
                    }
                }
            } else if (g.kind == TG_DENSE2) {
                T mr[16], mi[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    mr[j] = (T)mp[2 * j];
                    mi[j] = (T)mp[2 * j + 1];
                }
                const uint32_t n_groups = tile_amps >> g.n_holes;
                const uint32_t o0 = g.offs[0], o1 = g.offs[1], o2 = g.offs[2], o3 = g.offs[3];
                const bool active = threadIdx.x < n_groups;
                const uint32_t e_tid = expand_local(threadIdx.x, g.holes, g.n_holes) | g.loc_ctrl;
                const int iters = n_groups >= TILE_NT ? (int)(n_groups / TILE_NT) : 1;
                if (active) {
#pragma unroll 2
                    for (int it = 0; it < iters; ++it) {
                        const uint32_t i0 = e_tid | g.hi_tab[it];
                        A x[4];
                        x[0] = s[i0 + o0];
                        x[1] = s[i0 + o1];
                        x[2] = s[i0 + o2];
                        x[3] = s[i0 + o3];
                        A y[4];
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            T yr = T(0), yi = T(0);
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                yr += mr[r * 4 + c] * x[c].x - mi[r * 4 + c] * x[c].y;
                                yi += mr[r * 4 + c] * x[c].y + mi[r * 4 + c] * x[c].x;
                            }
                            y[r].x = yr;
                            y[r].y = yi;
                        }
                        s[i0 + o0] = y[0];
                        s[i0 + o1] = y[1];
                        s[i0 + o2] = y[2];
                        s[i0 + o3] = y[3];
                    }
                }
            } else {
                // DIAG / PARITY: per-amplitude phase; table bits outside the tile are fixed per CTA
                int t_fixed = 0;
                uint32_t loc_sel[4] = {0, 0, 0, 0};
                const bool parity = g.kind == TG_PARITY;
                if (parity) {
                    t_fixed = __popcll(outside & g.zmask_out) & 1;
                } else {
                    for (int b = 0; b < g.k; ++b) {
                        const int src = g.dsrc[b];
                        const int shift = g.k - 1 - b;
                        if (src >= 64)
                            t_fixed |= (int)((outside >> (src - 64)) & 1ull) << shift;
                        else
                            loc_sel[b] = 1u << src;
                    }
                }
                const uint32_t lc = g.loc_ctrl;
                for (uint32_t i = threadIdx.x; i < tile_amps; i += TILE_NT) {
                    if ((i & lc) == lc) {
                        int t = t_fixed;
                        if (parity) {
                            t ^= __popc(i & g.zmask_loc) & 1;
                        } else {
#pragma unroll
                            for (int b = 0; b < 4; ++b)
                                if (b < g.k && (i & loc_sel[b])) t |= 1 << (g.k - 1 - b);
                        }
                        const T pr = (T)mp[2 * t], pi = (T)mp[2 * t + 1];
                        const A a = s[i];
                        A y;
                        y.x = pr * a.x - pi * a.y;
                        y.y = pr * a.y + pi * a.x;
                        s[i] = y;
                    }
                }
            }
        }
        __syncthreads();
    }

    // ---- write the tile back ---------------------------------------------------------------------
    if constexpr (USE_TMA) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32) {
            for (int r = lane; r < n_runs; r += 32) {
                uint64_t off = 0;
                for (int j = 0; j < P.tb - P.L; ++j) off |= (uint64_t)((r >> j) & 1) << P.hi_bits[j];
                bulk_s2g(gbase + (base | off) * sizeof(A), smem_u32(smem_raw) + (uint32_t)r * run_bytes, run_bytes);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
    } else {
        const uint32_t vec_per_run = run_bytes / 16;
        const uint32_t total = (uint32_t)n_runs * vec_per_run;
        for (uint32_t v = threadIdx.x; v < total; v += TILE_NT) {
            const uint32_t r = v / vec_per_run, w = v % vec_per_run;
            uint64_t off = 0;
            for (int j = 0; j < P.tb - P.L; ++j) off |= (uint64_t)((r >> j) & 1) << P.hi_bits[j];
            *reinterpret_cast<int4 *>(gbase + (base | off) * sizeof(A) + (size_t)w * 16) =
                reinterpret_cast<const int4 *>(smem_raw)[v];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// host side: gate merging, sweep packing, program construction
// ------------------------------------------------------------------------------------------------
int env_int(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

uint64_t touched_mask(const LoweredGate &g) {  // bits whose value differs inside a DENSE group
    uint64_t m = 0;
    for (uint64_t o : g.offs) m |= o;
    return m;
}

uint64_t all_bits(const LoweredGate &g) {
    uint64_t m = g.ctrl_mask | g.zmask;
    for (int h : g.holes) m |= 1ull << h;
    for (int b : g.tgt_bits) m |= 1ull << b;
    return m;
}

bool tile_fusable(const LoweredGate &g, int n_local) {
    switch (g.kind) {
    case LoweredGate::DENSE:
        if (g.k > 2 || g.holes.size() > 8) return false;
        return (touched_mask(g) >> n_local) == 0;
    case LoweredGate::DIAG: return g.k <= 4;
    case LoweredGate::PARITY: return true;
    default: return false;
    }
}

struct Sweep {
    std::vector<const LoweredGate *> gates;
    uint64_t need = 0;  // dense-touched bits >= L
    int pool = 0;
};

int pool_need(const LoweredGate &g) {
    if (g.kind == LoweredGate::DENSE) return g.k == 1 ? 8 : 32;
    if (g.kind == LoweredGate::DIAG) return 2 << g.k;
    return 4;
}

template <typename T> void launch_sweep_t(State &sv, const TileProgram &P, void *const *table, int n_vecs, bool tma) {
    const size_t smem = ((size_t)1 << P.tb) * sizeof(typename Cx<T>::type);
    static bool configured[64][2] = {{false, false}};  // per device: function attributes belong to the device's context
    auto kern = tma ? k_tile_sweep<T, true> : k_tile_sweep<T, false>;
    if (!configured[sv.device & 63][tma ? 1 : 0]) {
        QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
        configured[sv.device & 63][tma ? 1 : 0] = true;
    }
    dim3 grid((unsigned)(1ull << (sv.n - P.tb)), (unsigned)n_vecs);
    kern<<<grid, TILE_NT, smem, sv.stream>>>(table ? nullptr : sv.data, table, P);
    QSV_CUDA(cudaGetLastError());
}

void run_sweep(State &sv, const Sweep &sw, int tb, int L, void *const *table, int n_vecs, bool tma) {
    const int n = sv.n;
    TileProgram P;
    memset(&P, 0, sizeof(P));
    P.tb = tb;
    P.L = L;
    P.index_hi = sv.index_hi;
    // tile bits: low L bits, the needed bits, then the lowest free bits
    std::vector<int> hi;
    for (int b = L; b < n; ++b)
        if (sw.need >> b & 1) hi.push_back(b);
    for (int b = L; b < n && (int)hi.size() < tb - L; ++b)
        if (!(sw.need >> b & 1)) hi.push_back(b);
    std::sort(hi.begin(), hi.end());
    QSV_CHECK((int)hi.size() == tb - L, "internal: tile bit selection");
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
    int pool = 0;
    for (const LoweredGate *gp : sw.gates) {
        const LoweredGate &g = *gp;
        TileGate &t = P.gates[P.n_gates++];
        t.mat_off = (uint32_t)pool;
        for (const cplx &c : g.mat) {
            P.pool[pool++] = c.real();
            P.pool[pool++] = c.imag();
        }
        t.loc_ctrl = map_mask(g.ctrl_mask, t.out_ctrl);
        if (g.kind == LoweredGate::DENSE) {
            t.kind = g.k == 1 ? TG_DENSE1 : TG_DENSE2;
            if (g.k == 1) {
                const cplx zero(0.0, 0.0), one(1.0, 0.0);
                bool real = true;
                for (const cplx &c : g.mat) real = real && c.imag() == 0.0;
                if (g.mat[0] == zero && g.mat[3] == zero && g.mat[1] == one && g.mat[2] == one)
                    t.kind = TG_SWAP1;
                else if (real)
                    t.kind = TG_REAL1;
            }
            int nh = 0;
            for (int h : g.holes)
                if (pos[h] >= 0) t.holes[nh++] = (unsigned char)pos[h];  // ascending: pos is monotonic
            t.n_holes = (unsigned char)nh;
            for (int it = 0; it < 16; ++it) {
                uint32_t o = (uint32_t)it * TILE_NT;
                for (int j = 0; j < nh; ++j) {
                    const int hp = t.holes[j];
                    o = ((o >> hp) << (hp + 1)) | (o & ((1u << hp) - 1u));
                }
                t.hi_tab[it] = (uint16_t)(o & 0xffffu);  // entries beyond the tile are never used
            }
            for (size_t j = 0; j < g.offs.size(); ++j) {
                uint64_t dummy;
                t.offs[j] = (uint16_t)map_mask(g.offs[j], dummy);
                QSV_CHECK(dummy == 0, "internal: dense target outside the tile");
            }
        } else if (g.kind == LoweredGate::DIAG) {
            t.kind = TG_DIAG;
            t.k = (unsigned char)g.k;
            for (int b = 0; b < g.k; ++b) {
                const int gb = g.tgt_bits[b];
                t.dsrc[b] = (unsigned char)(pos[gb] >= 0 ? pos[gb] : 64 + gb);
            }
        } else {
            t.kind = TG_PARITY;
            t.zmask_loc = map_mask(g.zmask, t.zmask_out);
        }
    }
    sv.stat_launches += 1;
    sv.stat_sweeps += 1;
    if (sv.dtype == QSV_C128)
        launch_sweep_t<double>(sv, P, table, n_vecs, tma);
    else
        launch_sweep_t<float>(sv, P, table, n_vecs, tma);
}

// ---- merge runs of uncontrolled single-qubit gates on the same qubit into one 2x2 ---------------
bool as_plain_1q(const LoweredGate &g, int &bit, cplx m[4], bool &diag) {
    if (g.kind == LoweredGate::DENSE && g.k == 1 && g.ctrl_mask == 0 && g.holes.size() == 1 && g.offs[0] == 0 &&
        g.offs[1] == (1ull << g.holes[0])) {
        bit = g.holes[0];
        for (int i = 0; i < 4; ++i) m[i] = g.mat[i];
        diag = false;
        return true;
    }
    if (g.kind == LoweredGate::DIAG && g.k == 1 && g.ctrl_mask == 0) {
        bit = g.tgt_bits[0];
        m[0] = g.mat[0];
        m[1] = m[2] = 0.0;
        m[3] = g.mat[1];
        diag = true;
        return true;
    }
    if (g.kind == LoweredGate::DIAG && g.k == 0 && __builtin_popcountll(g.ctrl_mask) == 1) {
        bit = __builtin_ctzll(g.ctrl_mask);
        m[0] = 1.0;
        m[1] = m[2] = 0.0;
        m[3] = g.mat[0];
        diag = true;
        return true;
    }
    return false;
}

// an uncontrolled dense 4x4 on two bits (tgt_bits[0] = matrix MSB)
bool as_plain_2q(const LoweredGate &g) {
    return g.kind == LoweredGate::DENSE && g.k == 2 && g.ctrl_mask == 0 && g.tgt_bits.size() == 2 && g.holes.size() == 2;
}

// Host-side gate fusion ahead of the sweep packer (north_star item c, "dense k-qubit blocks" with k <= 2):
//  * runs of uncontrolled single-qubit gates on the same qubit become one 2x2 (diagonal if all of them are);
//  * an uncontrolled dense two-qubit gate absorbs the pending single-qubit blocks of its two qubits, every later
//    single-qubit gate on them, and later dense two-qubit gates on the same pair, until another gate touches one of
//    the two qubits.  A 4x4 block costs the same 16 FP64 operations per amplitude however many gates it absorbed.
// QSV_MERGE_2Q=0 keeps two-qubit gates as they are.
std::vector<LoweredGate> merge_single_qubit_runs(const std::vector<LoweredGate> &in) {
    struct Pending {
        int k = 1;            // qubits
        int hi = -1, lo = -1; // k = 1: bit = lo;  k = 2: matrix index = 2 * bit(hi) + bit(lo)
        cplx m[16];
        bool diag = false;
        bool live = true;
    };
    const bool merge2 = env_int("QSV_MERGE_2Q", 1) != 0;
    std::vector<Pending> pool;
    std::map<int, int> pending;  // bit -> index into pool
    std::vector<LoweredGate> out;
    auto flush = [&](int bit) {
        auto it = pending.find(bit);
        if (it == pending.end()) return;
        Pending &p = pool[it->second];
        if (p.k == 1) {
            if (p.diag)
                out.push_back(make_diag_gate({p.lo}, 0, {p.m[0], p.m[3]}));
            else
                out.push_back(make_dense_gate({p.lo}, 0, {p.m[0], p.m[1], p.m[2], p.m[3]}));
            pending.erase(it);
        } else {
            out.push_back(make_dense_gate({p.hi, p.lo}, 0, std::vector<cplx>(p.m, p.m + 16)));
            const int hi = p.hi, lo = p.lo;
            pending.erase(hi);
            pending.erase(lo);
        }
        p.live = false;
    };
    // 4x4 = a (x) b in the (hi, lo) basis
    auto kron = [](const cplx a[4], const cplx b[4], cplx r[16]) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) r[i * 4 + j] = a[(i >> 1) * 2 + (j >> 1)] * b[(i & 1) * 2 + (j & 1)];
    };
    auto matmul4 = [](const cplx a[16], const cplx b[16], cplx r[16]) {
        for (int i = 0; i < 4; ++i)
            for (int j = 0; j < 4; ++j) {
                cplx s = 0.0;
                for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
                r[i * 4 + j] = s;
            }
    };
    const cplx id2[4] = {1.0, 0.0, 0.0, 1.0};
    for (const LoweredGate &g : in) {
        if (g.kind == LoweredGate::NOP) continue;
        int bit;
        cplx m[4];
        bool diag;
        if (as_plain_1q(g, bit, m, diag)) {
            auto it = pending.find(bit);
            if (it == pending.end()) {
                Pending p;
                p.k = 1;
                p.lo = bit;
                for (int i = 0; i < 4; ++i) p.m[i] = m[i];
                p.diag = diag;
                pool.push_back(p);
                pending[bit] = (int)pool.size() - 1;
            } else {
                Pending &p = pool[it->second];
                if (p.k == 1) {  // new = m * old
                    cplx r[4] = {m[0] * p.m[0] + m[1] * p.m[2], m[0] * p.m[1] + m[1] * p.m[3],
                                 m[2] * p.m[0] + m[3] * p.m[2], m[2] * p.m[1] + m[3] * p.m[3]};
                    for (int i = 0; i < 4; ++i) p.m[i] = r[i];
                    p.diag = p.diag && diag;
                } else {  // new = (m on this bit) * old
                    cplx full[16], r[16];
                    if (bit == p.hi)
                        kron(m, id2, full);
                    else
                        kron(id2, m, full);
                    matmul4(full, p.m, r);
                    for (int i = 0; i < 16; ++i) p.m[i] = r[i];
                }
            }
            continue;
        }
        if (merge2 && as_plain_2q(g)) {
            const int hi = g.tgt_bits[0], lo = g.tgt_bits[1];
            auto ih = pending.find(hi), il = pending.find(lo);
            if (ih != pending.end() && il != pending.end() && ih->second == il->second && pool[ih->second].k == 2) {
                // same pair: new = G * old, G expressed in the pending block's bit order
                Pending &p = pool[ih->second];
                cplx gm[16], r[16];
                const bool same_order = p.hi == hi;
                for (int i = 0; i < 4; ++i)
                    for (int j = 0; j < 4; ++j) {
                        const int si = same_order ? i : ((i & 1) << 1) | (i >> 1), sj = same_order ? j : ((j & 1) << 1) | (j >> 1);
                        gm[i * 4 + j] = g.mat[si * 4 + sj];
                    }
                matmul4(gm, p.m, r);
                for (int i = 0; i < 16; ++i) p.m[i] = r[i];
                continue;
            }
            // blocks on a different pair end here; single-qubit blocks are absorbed: new = G * (P_hi (x) P_lo)
            for (int b : {hi, lo}) {
                auto it = pending.find(b);
                if (it != pending.end() && pool[it->second].k == 2) flush(b);
            }
            cplx ph[4] = {1.0, 0.0, 0.0, 1.0}, pl[4] = {1.0, 0.0, 0.0, 1.0};
            for (int w = 0; w < 2; ++w) {
                auto it = pending.find(w == 0 ? hi : lo);
                if (it == pending.end()) continue;
                Pending &q1 = pool[it->second];
                for (int i = 0; i < 4; ++i) (w == 0 ? ph : pl)[i] = q1.m[i];
                q1.live = false;
                pending.erase(it);
            }
            Pending p;
            p.k = 2;
            p.hi = hi;
            p.lo = lo;
            cplx pre[16], gm[16];
            kron(ph, pl, pre);
            for (int i = 0; i < 16; ++i) gm[i] = g.mat[i];
            matmul4(gm, pre, p.m);
            pool.push_back(p);
            pending[hi] = pending[lo] = (int)pool.size() - 1;
            continue;
        }
        const uint64_t bits = all_bits(g);
        for (int b = 0; b < 64; ++b)
            if (bits >> b & 1) flush(b);
        out.push_back(g);
    }
    while (!pending.empty()) flush(pending.begin()->first);
    return out;
}

}  // namespace

// Register-blocked executor (tile_regs.cu).  The tile has 12 - L arbitrary high bits with L as small as one
// warp-wide access allows, and the gates of a sweep are list-scheduled into register passes.
//
// Sweep packing works on the dependency DAG of the circuit instead of its program order: two gates commute when
// they share no index bit, or share only bits on which both act diagonally (controls, phase tables, parity masks),
// so a sweep may take any gate whose predecessors are already inside it (or done) as long as the union of the
// high dense-target bits still fits the tile.  On the config-2 circuit (200 random gates, 30 qubits) this packs
// 8 sweeps instead of the 12 of in-order packing.  QSV_REGS_DAG=0 restores program order.
static std::vector<SweepPlan> plan_sweeps_once(int n_local, const std::vector<LoweredGate> &gates, int L, bool dag,
                                               int max_gates, int window, int order_seed) {
    const int max_hi = 12 - L;
    const uint64_t low = (1ull << L) - 1ull;
    std::vector<SweepPlan> plan;
    auto pool_of = [](const LoweredGate &g) {
        return (g.kind == LoweredGate::DENSE && !(g.k == 1 && g.tgt_bits.size() == 1)) ? 66 : 8;  // 4x4: up to a real 8x8 (tensor-core block)
    };
    // one maximal run of fusable gates (indices into `gates`), packed over its dependency DAG
    auto pack_segment = [&](const std::vector<int> &seg) {
        const int m = (int)seg.size();
        std::vector<uint64_t> need(m);
        std::vector<std::vector<int>> succ(m);
        std::vector<int> unsat(m, 0);
        {
            // exact dependencies through per-bit tracking: the last gate that touched the bit non-diagonally and the
            // diagonal touches since then
            int last_dense[64];
            std::vector<int> diag_since[64];
            for (int b = 0; b < 64; ++b) last_dense[b] = -1;
            auto edge = [&](int i, int j) {
                if (i < 0 || (!succ[i].empty() && succ[i].back() == j)) return;
                succ[i].push_back(j);
                ++unsat[j];
            };
            for (int j = 0; j < m; ++j) {
                const LoweredGate &g = gates[seg[j]];
                const uint64_t dense = regs_need_bits(g);
                need[j] = dense & ~low;
                const uint64_t bits = all_bits(g) | dense;
                for (int b = 0; b < 64; ++b) {
                    if (!(bits >> b & 1)) continue;
                    edge(last_dense[b], j);
                    if (dense >> b & 1) {
                        for (int i : diag_since[b]) edge(i, j);
                        diag_since[b].clear();
                        last_dense[b] = j;
                    } else {
                        diag_since[b].push_back(j);
                    }
                }
            }
        }
        std::vector<char> done(m, 0);
        int first = 0, n_done = 0;
        while (n_done < m) {
            while (done[first]) ++first;
            SweepPlan sw;
            int cur_pool = 0;
            auto try_take = [&](int j) {
                const int pn = pool_of(gates[seg[j]]);
                if (__builtin_popcountll(sw.need | need[j]) > max_hi || (int)sw.gates.size() >= max_gates ||
                    cur_pool + pn > 1280)
                    return false;
                sw.gates.push_back(seg[j]);
                sw.need |= need[j];
                cur_pool += pn;
                done[j] = 1;
                ++n_done;
                for (int k : succ[j]) --unsat[k];
                return true;
            };
            if (!dag) {
                // program order: the first gate that does not fit ends the sweep
                for (int j = first; j < m && try_take(j); ++j) {
                }
            } else {
                // first fit over the ready gates of a look-ahead window, repeated until nothing more fits
                // order_seed 0: program order; 1 / 2: by lowest / highest index bit; > 2: a fixed pseudo-random order of the
                // window (multi-start, see below)
                const int stop = std::min(m, first + window);
                std::vector<int> order;
                for (int j = first; j < stop; ++j) order.push_back(j);
                if (order_seed == 1 || order_seed == 2) {
                    // by index bit instead of program order: a layered circuit is then packed along its wires (all layers
                    // of a group of neighbouring qubits in one sweep, as far as the entangling gates allow)
                    auto key = [&](int j) {
                        const LoweredGate &g = gates[seg[j]];
                        const uint64_t bits = all_bits(g) | regs_need_bits(g);
                        const int b = bits ? (order_seed == 1 ? __builtin_ctzll(bits) : 63 - __builtin_clzll(bits)) : 0;
                        return order_seed == 1 ? b : -b;
                    };
                    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return key(a) < key(b); });
                } else if (order_seed > 2) {
                    uint32_t r = 2654435761u * (uint32_t)(order_seed + 17 * (int)plan.size());
                    for (int k = (int)order.size() - 1; k > 0; --k) {
                        r = r * 1664525u + 1013904223u;
                        std::swap(order[k], order[(r >> 8) % (uint32_t)(k + 1)]);
                    }
                }
                for (bool progress = true; progress;) {
                    progress = false;
                    for (int j : order)
                        if (!done[j] && unsat[j] == 0 && try_take(j)) progress = true;
                }
            }
            QSV_CHECK(!sw.gates.empty(), "internal: sweep packing made no progress");
            sw.fused = sw.gates.size() > 1;  // a lone gate: the one-sweep kernel touches only what the gate changes
            plan.push_back(std::move(sw));
        }
    };
    std::vector<int> seg;
    for (int i = 0; i < (int)gates.size(); ++i) {
        const LoweredGate &g = gates[i];
        if (g.kind == LoweredGate::NOP) continue;
        if (!regs_fusable(g, n_local)) {
            pack_segment(seg);
            seg.clear();
            SweepPlan sw;
            sw.gates.push_back(i);
            plan.push_back(std::move(sw));
            continue;
        }
        seg.push_back(i);
    }
    pack_segment(seg);
    return plan;
}

// Multi-start packing (QSV_REGS_PACK_TRIES; default 3 from 28 qubits up -- where its ~3 ms of host time, paid once per
// gate-list structure thanks to the plan cache, are small against the circuit's GPU time -- else 1): first fit depends on
// the order in which the ready gates are offered.  Besides program order the gates are offered by lowest and by highest
// index bit (try 1 and 2; further tries use fixed pseudo-random orders).  Program order is close to the best found for
// random circuits, but layered circuits (rotations on every wire + an entangling ladder) pack much better along their wires
// -- the 30-qubit hardware-efficient ansatz: 9 sweeps / 26 passes instead of 11 / 30.  Candidates are priced with the cost
// model of tools/sweep_cost_model.py on the programs the pass scheduler builds for them.
int regs_pack_tries(int n_local, bool seen_before) {
    // from 28 qubits up the search pays even for a circuit that is applied once; from 20 qubits up it is run when a gate-list
    // structure comes back (plan cache), i.e. for the loops -- variational iterations, adjoint sweeps layer by layer,
    // benchmarks -- that apply one structure many times
    return std::max(1, env_int("QSV_REGS_PACK_TRIES", n_local >= 28 || (seen_before && n_local >= 20) ? 3 : 1));
}

std::vector<SweepPlan> plan_sweeps_regs(int n_local, const std::vector<LoweredGate> &gates, int L, bool dag, int max_gates,
                                        int window, int dtype, int tries_in) {
    std::vector<SweepPlan> best = plan_sweeps_once(n_local, gates, L, dag, max_gates, window, 0);
    const int tries = dag && n_local >= 12 ? (tries_in > 0 ? tries_in : regs_pack_tries(n_local, false)) : 1;
    if (tries <= 1) return best;
    std::vector<const LoweredGate *> cur;
    auto price = [&](const std::vector<SweepPlan> &plan, bool greedy) {
        double c = 0.0;
        for (const SweepPlan &sw : plan) {
            if (!sw.fused) {
                c += 5.3;  // one HBM sweep of a lone gate
                continue;
            }
            cur.clear();
            for (int i : sw.gates) cur.push_back(&gates[i]);
            c += regs_sweep_model_cost(n_local, dtype, cur, sw.need, L, greedy);
        }
        return c;
    };
    // candidates are priced with the greedy pass scheduler (cheap); the winner then has to beat program order under the
    // scheduler that will really build the programs (the pass-sequence search), so the result is never worse than it
    std::vector<SweepPlan> winner;
    double winner_cost = price(best, true);
    bool have_winner = false;
    for (int t = 1; t < tries; ++t) {
        std::vector<SweepPlan> cand = plan_sweeps_once(n_local, gates, L, dag, max_gates, window, t);
        const double c = price(cand, true);
        if (c < winner_cost - 1e-9) {
            winner_cost = c;
            winner = std::move(cand);
            have_winner = true;
        }
    }
    if (have_winner && price(winner, false) < price(best, false) - 1e-9) best = std::move(winner);
    return best;
}

std::vector<LoweredGate> prepare_gates_regs(const std::vector<LoweredGate> &gates_in) {
    return env_int("QSV_MERGE_1Q", 1) != 0 ? merge_single_qubit_runs(gates_in) : gates_in;
}

// commutation rule the packer relies on, pairwise (used by the plan self-check of the C ABI and by the tests)
bool gates_commute_structurally(const LoweredGate &a, const LoweredGate &b) {
    const uint64_t da = a.kind == LoweredGate::DENSE ? touched_mask(a) : 0, db = b.kind == LoweredGate::DENSE ? touched_mask(b) : 0;
    const uint64_t ba = all_bits(a) | da, bb = all_bits(b) | db;
    return ((da & bb) | (ba & db)) == 0;
}

// Sweep plans of the circuits seen last, keyed by the structure of the merged gate list (kinds and index bits, not the
// matrix entries: dependencies and tile needs are structural).  A variational loop or a benchmark applies the same
// structure again and again; the multi-start packing then costs its 2 ms of host time once.
namespace {
struct PlanCacheEntry {
    uint64_t h1 = 0, h2 = 0;
    int tries = 1;  // how hard the packing was searched for this plan
    std::vector<SweepPlan> plan;
};
std::mutex g_plan_cache_mu;
std::vector<PlanCacheEntry> g_plan_cache;  // most recently used first, at most 32 entries

void structure_hash(const std::vector<LoweredGate> &gates, const int *params, int n_params, uint64_t &h1, uint64_t &h2) {
    h1 = 0xcbf29ce484222325ull;
    h2 = 0x9e3779b97f4a7c15ull;
    auto mix = [&](uint64_t v) {
        h1 = (h1 ^ v) * 0x100000001b3ull;
        h2 = (h2 + v) * 0xff51afd7ed558ccdull;
        h2 ^= h2 >> 33;
    };
    for (int i = 0; i < n_params; ++i) mix((uint64_t)(int64_t)params[i]);
    for (const LoweredGate &g : gates) {
        mix(0xabcdefull + (uint64_t)g.kind * 131u + (uint64_t)g.k);
        mix(g.ctrl_mask);
        mix(g.zmask);
        mix(g.holes.size());
        for (int h : g.holes) mix((uint64_t)h);
        mix(g.tgt_bits.size());
        for (int b : g.tgt_bits) mix((uint64_t)b);
        mix(g.offs.size());
        for (uint64_t o : g.offs) mix(o);
    }
}
}  // namespace

std::vector<SweepPlan> plan_sweeps_cached(int n_local, const std::vector<LoweredGate> &merged, int L, bool dag, int max_gates,
                                          int window, int dtype) {
    if (env_int("QSV_REGS_PLAN_CACHE", 1) == 0) return plan_sweeps_regs(n_local, merged, L, dag, max_gates, window, dtype);
    uint64_t h1 = 0, h2 = 0;
    const int key[] = {n_local, L, (int)dag, max_gates, window, dtype, env_int("QSV_REGS_PACK_TRIES", -1), env_int("QSV_REGS_RB", 4),
                       env_int("QSV_REGS_MMA", 2), env_int("QSV_REGS_FOLD", 1), env_int("QSV_REGS_UDIAG", 1),
                       env_int("QSV_REGS_DIAG1", 1)};
    structure_hash(merged, key, (int)(sizeof(key) / sizeof(key[0])), h1, h2);
    bool seen_before = false;
    {
        std::lock_guard<std::mutex> lock(g_plan_cache_mu);
        for (size_t i = 0; i < g_plan_cache.size(); ++i)
            if (g_plan_cache[i].h1 == h1 && g_plan_cache[i].h2 == h2) {
                if (i != 0) std::rotate(g_plan_cache.begin(), g_plan_cache.begin() + i, g_plan_cache.begin() + i + 1);
                if (g_plan_cache[0].tries >= regs_pack_tries(n_local, true)) return g_plan_cache[0].plan;
                seen_before = true;  // a structure that comes back is worth the search it was spared the first time
                break;
            }
    }
    const int tries = regs_pack_tries(n_local, seen_before);
    std::vector<SweepPlan> plan = plan_sweeps_regs(n_local, merged, L, dag, max_gates, window, dtype, tries);
    std::lock_guard<std::mutex> lock(g_plan_cache_mu);
    for (size_t i = 0; i < g_plan_cache.size(); ++i)
        if (g_plan_cache[i].h1 == h1 && g_plan_cache[i].h2 == h2) {
            g_plan_cache.erase(g_plan_cache.begin() + i);
            break;
        }
    PlanCacheEntry e;
    e.h1 = h1;
    e.h2 = h2;
    e.tries = tries;
    e.plan = plan;
    g_plan_cache.insert(g_plan_cache.begin(), std::move(e));
    if (g_plan_cache.size() > 32) g_plan_cache.pop_back();
    return plan;
}

static void apply_gates_regs(State &sv, const std::vector<LoweredGate> &gates_in, void *const *dev_table, int n_vecs,
                             FusedExchange *fx, FusedPull *pull) {
    int L = env_int("QSV_REGS_LOW", 4);  // measured on B200 (profiles/r1_regs_ab.txt)
    L = std::max(1, std::min(L, 11));
    const std::vector<LoweredGate> merged = prepare_gates_regs(gates_in);
    const std::vector<SweepPlan> plan =
        plan_sweeps_cached(sv.n, merged, L, env_int("QSV_REGS_DAG", 1) != 0, std::min(48, env_int("QSV_REGS_MAX_GATES", 48)),
                           std::max(1, env_int("QSV_REGS_WINDOW", 512)), sv.dtype);
    const bool single = dev_table == nullptr && n_vecs == 1;
    auto tile_ok = [&](const SweepPlan &sw) { return sw.fused || regs_fusable(merged[sw.gates[0]], sv.n); };
    // The second half of a split exchange comes first: carried by the first sweep when that sweep is not also the one that
    // has to carry the next exchange (the two halves have fixed buffer roles, whatever the shape of this rank's batch: the
    // pull always moves the register to its other buffer, the push always moves it again), else as a copy pass.
    bool moved = false, pull_pending = false;
    if (pull) {
        QSV_CHECK(single, "internal: a split exchange works on one vector");
        const bool first_carries = !plan.empty() && tile_ok(plan[0]) && regs_pull_supported() && !(fx && plan.size() == 1);
        if (!first_carries) {
            launch_xchg_pull_copy(sv, sv.data, pull->in_peer, pull->out_mine, pull->local_bit, pull->my_value, pull->stash_bit);
            sv.data = pull->out_mine;
            moved = true;
            if (pull->after) pull->after();
        } else {
            pull_pending = true;
        }
    }
    std::vector<const LoweredGate *> cur;
    for (size_t k = 0; k < plan.size(); ++k) {
        const SweepPlan &sw = plan[k];
        // the last sweep of the batch carries the exchange (its stores are routed element by element, xchg_target); a lone
        // gate in that position runs through the tile kernel too when it can: one pass over the shard instead of the
        // gate's own sweep plus a copy pass for the exchange
        const bool carry = fx != nullptr && k + 1 == plan.size() && single && tile_ok(sw);
        const bool pulls = pull_pending && k == 0;
        if (!sw.fused && !carry && !pulls) {
            const LoweredGate &g = merged[sw.gates[0]];
            if (dev_table)
                launch_gate_multi(sv, g, dev_table, n_vecs);
            else
                launch_gate(sv, g);
            continue;
        }
        cur.clear();
        for (int i : sw.gates) cur.push_back(&merged[i]);
        if (carry && fx->before) fx->before();
        run_sweep_regs(sv, cur, sw.need, L, dev_table, n_vecs, carry ? fx : nullptr, moved ? 1 : 0, pulls ? pull : nullptr);
        if (pulls) {
            sv.data = pull->out_mine;
            moved = true;
            pull->carried = true;
            pull_pending = false;
            if (pull->after) pull->after();
        }
        if (carry) fx->done = true;
    }
    if (fx) fx->moved_before = moved;
}

void apply_gates_tiled(State &sv, const std::vector<LoweredGate> &gates_in, void *const *dev_table, int n_vecs,
                       FusedExchange *fx, FusedPull *pull) {
    sv.use();
    if (fx) fx->done = false;
    if (sv.n >= 12 && env_int("QSV_TILE_KERNEL", 1) == 1) {
        apply_gates_regs(sv, gates_in, dev_table, n_vecs, fx, pull);
        return;
    }
    if (pull) {  // the first-generation executor carries nothing: copy pass first
        launch_xchg_pull_copy(sv, sv.data, pull->in_peer, pull->out_mine, pull->local_bit, pull->my_value, pull->stash_bit);
        sv.data = pull->out_mine;
        if (pull->after) pull->after();
        if (fx) fx->moved_before = true;
    }
    const bool f32 = sv.dtype == QSV_C64;
    int tb = env_int("QSV_TILE_BITS", f32 ? 13 : 12);
    int L = env_int("QSV_TILE_LOW", f32 ? 8 : 7);
    const bool tma = env_int("QSV_TILE_TMA", 1) != 0;
    const bool merge = env_int("QSV_MERGE_1Q", 1) != 0;
    tb = std::min(tb, sv.n);
    tb = std::min(tb, f32 ? 13 : 12);
    L = std::max(1, std::min(L, tb));
    if ((1 << L) * sv.amp_bytes() < 16) L = 1;  // bulk copies move multiples of 16 bytes
    const int max_hi = tb - L;

    const std::vector<LoweredGate> merged = merge ? merge_single_qubit_runs(gates_in) : gates_in;
    Sweep cur;
    auto flush = [&]() {
        if (cur.gates.empty()) return;
        if (cur.gates.size() == 1 && !tile_fusable(*cur.gates[0], sv.n)) {
            if (dev_table)
                launch_gate_multi(sv, *cur.gates[0], dev_table, n_vecs);
            else
                launch_gate(sv, *cur.gates[0]);
        } else if (cur.gates.size() == 1) {
            // a lone gate: the register kernel touches only what the gate changes
            if (dev_table)
                launch_gate_multi(sv, *cur.gates[0], dev_table, n_vecs);
            else
                launch_gate(sv, *cur.gates[0]);
        } else {
            run_sweep(sv, cur, tb, L, dev_table, n_vecs, tma);
        }
        cur = Sweep();
    };
    for (const LoweredGate &g : merged) {
        if (g.kind == LoweredGate::NOP) continue;
        if (!tile_fusable(g, sv.n)) {
            flush();
            cur.gates.push_back(&g);
            flush();
            continue;
        }
        uint64_t need = 0;
        if (g.kind == LoweredGate::DENSE) need = touched_mask(g) & ~((1ull << L) - 1ull);
        const int pn = pool_need(g);
        const bool fits = __builtin_popcountll(cur.need | need) <= max_hi &&
                          (int)cur.gates.size() < TILE_MAX_GATES && cur.pool + pn <= TILE_POOL;
        if (!fits) flush();
        QSV_CHECK(__builtin_popcountll(need) <= max_hi, "internal: gate does not fit a tile");
        cur.gates.push_back(&g);
        cur.need |= need;
        cur.pool += pn;
    }
    flush();
}

void apply_ops_fused(State &sv, const std::vector<LoweredGate> &gates) { apply_gates_tiled(sv, gates, nullptr, 1); }

}  // namespace qsv
