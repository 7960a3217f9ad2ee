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
#include <type_traits>
#include <cstdlib>
#include <cstring>

#include <vector>

#include "qsv_internal.h"
#include "tile_regs_core.cuh"

namespace qsv {

using namespace rt;

namespace {

constexpr int RT_TB = TB;
constexpr int RT_MAX_GATES = MAX_GATES;
constexpr int RT_MAX_PASSES = MAX_PASSES;
constexpr int RT_POOL = POOL;

// per-SM arrival counters of the current launch: (epoch << 32) | CTAs that have started on this SM
__device__ unsigned long long g_sm_arrivals[256];

// Sweep fused with a global<->local index-bit exchange of a sharded register (XCHG = true, csrc/dist.cu): the sweep reads
// the shard in place but stores out of place -- tiles whose index bit `bit` equals this rank's value of the exchanged
// global bit go to the same offset of this rank's other buffer, all other tiles to the partner GPU's other buffer (a
// peer mapping: plain stores over NVLink) at the offset with that bit flipped.  The choice is made per stored element
// (xchg_target), so `bit` may also be a tile bit.  The transfer overlaps the arithmetic of the tiles in flight; no
// separate exchange pass, no staging.
struct XchgArgs {
    // push (the sweep in front of an exchange): stores routed by xchg_target; bit_mask = 0 switches it off (everything goes to
    // out_mine at its own offset: a plain out-of-place sweep, as the pull sweep is)
    void *out_mine;
    void *out_peer;
    uint64_t bit_mask;    // 1 << (exchanged local bit)
    uint64_t keep;        // bit_mask when this rank's value of the global bit is 1, else 0
    uint64_t stash_mask;  // != 0: the exchange is split, see xchg_target
    // pull (the sweep behind a split exchange): loads routed by xchg_source; pull_bit_mask = 0 switches it off
    const void *in_peer;
    uint64_t pull_bit_mask, pull_keep, pull_stash_mask;
    // When the exchanged bit is outside the tile, whole tiles go one way.  With `interleave` the LOWEST bit of the block
    // index is the exchanged bit, so staying and leaving tiles alternate instead of all leaving tiles coming in the second
    // half of the sweep.  (No measurable effect on the bench circuits, whose exchanged bits are tile bits.)
    int interleave;
    int bit_pos;
    Holes holes2;         // tile bits + the exchanged bit
};
struct NoXchg {};

template <typename T, int RB, int MINB, bool XCHG = false>
__global__ void __launch_bounds__(1 << (RT_TB - RB), MINB)
    k_tile_regs(void *single, void *const *table, const __grid_constant__ RegProgram P,
                const std::conditional_t<XCHG, XchgArgs, NoXchg> xa) {
    using A = typename Cx<T>::type;
    constexpr int NS = 1 << RB;            // amplitudes per thread
    constexpr int NT = 1 << (RT_TB - RB);  // threads per CTA
    constexpr int NTB = RT_TB - RB;        // thread bits
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *s = reinterpret_cast<A *>(smem_raw);
    A *gbase = reinterpret_cast<A *>(table ? table[blockIdx.y] : single);
    uint64_t base_v;
    if constexpr (XCHG) {
        if (xa.interleave)
            base_v = expand_index((uint64_t)blockIdx.x >> 1, xa.holes2) | ((uint64_t)(blockIdx.x & 1u) << xa.bit_pos);
        else
            base_v = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    } else {
        base_v = expand_index((uint64_t)blockIdx.x, P.tile_holes);
    }
    const uint64_t base = base_v;
    const uint64_t outside = base | P.index_hi;
    const uint32_t tid = threadIdx.x;
    // The tile that the CTA scheduled `prefetch` CTAs after this one will load is requested into L2 now, so that its
    // HBM latency is hidden behind the arithmetic of the tiles in flight.
    if (P.prefetch > 0 && blockIdx.x + (unsigned)P.prefetch < gridDim.x) {
        const uint64_t pbase = expand_index((uint64_t)blockIdx.x + (uint64_t)P.prefetch, P.tile_holes);
        const uint64_t gt = pbase ^ thread_offset64<NTB>(P.gl_load.thr, P.gl_load.c, tid);
        uint64_t gr[RB_MAX];
#pragma unroll
        for (int b = 0; b < RB; ++b) gr[b] = P.gl_load.reg[b];
        if ((tid & (sizeof(A) == 16 ? 7u : 15u)) == 0u) {  // one request per 128-byte line of the warp's row
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const A *p = gbase + slot_offset<RB>(gt, gr, j);
                asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
            }
        }
    }
    // Experiment (off by default): every tile costs the same, so the CTAs resident on an SM -- and on the whole GPU -- might
    // run in lockstep, all of them loading from HBM at the same time, then all of them computing (time per tile looks like
    // HBM time + FP64 time + transposition time).  The k-th CTA to arrive on each SM (k < CTAs per SM)
    // therefore waits k / (CTAs per SM) of a tile time, once; after that the CTA slots of the SM stay out of phase.
    __shared__ unsigned s_arrival;
    if (tid == 0) {
        unsigned mine = 0;
        if (P.stagger_ns > 0 && blockIdx.x < 1024u && blockIdx.y == 0) {
            unsigned smid;
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            unsigned long long *w = &g_sm_arrivals[smid & 255u];
            unsigned long long old = *reinterpret_cast<volatile unsigned long long *>(w);
            for (;;) {
                const bool fresh = (unsigned)(old >> 32) != P.epoch;
                const unsigned long long want = fresh ? ((unsigned long long)P.epoch << 32) | 1ull : old + 1ull;
                const unsigned long long seen = atomicCAS(w, old, want);
                if (seen == old) {
                    mine = fresh ? 0u : (unsigned)(old & 0xffffffffull);
                    break;
                }
                old = seen;
            }
        }
        s_arrival = mine;
    }
    // Gate constants go to shared memory once per CTA (in the kernel's precision) and are then read with uniform,
    // broadcast LDS: indexed constant-bank loads (LDC) inside the thread-divergent gate code run on the ADU pipe,
    // which ncu showed to be the busiest unit of the first version of this kernel (52 % vs 33 % FP64).
    T *spool = reinterpret_cast<T *>(smem_raw + (sizeof(A) << RT_TB));
    for (int i = tid; i < P.pool_used; i += NT) spool[i] = (T)P.pool[i];
    __syncthreads();
    if (s_arrival >= 1u && s_arrival < (unsigned)MINB) __nanosleep(P.stagger_ns * s_arrival);

    A x[NS];
    {
        const uint64_t gt = base ^ thread_offset64<NTB>(P.gl_load.thr, P.gl_load.c, tid);
        uint64_t gr[RB_MAX];
#pragma unroll
        for (int b = 0; b < RB; ++b) gr[b] = P.gl_load.reg[b];
        if constexpr (XCHG) {
            if (xa.pull_bit_mask != 0) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const XchgTarget t = xchg_source(slot_offset<RB>(gt, gr, j), xa.pull_bit_mask, xa.pull_keep, xa.pull_stash_mask);
                    x[j] = (t.stays ? gbase : reinterpret_cast<const A *>(xa.in_peer))[t.base];
                }
            } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) x[j] = gbase[slot_offset<RB>(gt, gr, j)];
            }
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) x[j] = gbase[slot_offset<RB>(gt, gr, j)];
        }
    }
    const int last = P.n_passes - 1;
    for (int pv = 0;; ++pv) {
        // the pass counter may be spilled under the 128-register cap; a broadcast from lane 0 tells the compiler that it is
        // warp-uniform again, so that pass / gate descriptors and gate constants are fetched with uniform loads (LDCU) into
        // uniform registers instead of per-thread LDC + vector registers
        const int p = __shfl_sync(0xffffffffu, pv, 0);
        const RegPass &ps = P.passes[p];
        pass_compute<T, RB>(x, P, ps, tid, outside, spool);
        if (p == last) break;
        // transpose through shared memory: store in this pass's layout (folded permutations included), load in the next one's
        if (p > 0) __syncthreads();  // every thread has finished reading the previous layout
        {
            const uint32_t st = thread_offset<NTB>(ps.st_thr, ps.st_c, tid);
            uint32_t sr[RB_MAX];
#pragma unroll
            for (int b = 0; b < RB; ++b) sr[b] = ps.st_reg[b];
#pragma unroll
            for (int j = 0; j < NS; ++j) s[slot_offset<RB>(st, sr, j)] = x[j];
        }
        __syncthreads();
        {
            const RegPass &pn = P.passes[p + 1];
            const uint32_t st = thread_offset<NTB>(pn.ld_thr, pn.ld_c, tid);
            uint32_t sr[RB_MAX];
#pragma unroll
            for (int b = 0; b < RB; ++b) sr[b] = pn.ld_reg[b];
#pragma unroll
            for (int j = 0; j < NS; ++j) x[j] = s[slot_offset<RB>(st, sr, j)];
        }
    }
    // A single-pass sweep has no barrier between its loads and its stores, and with a folded permutation a thread does
    // not write where it read.
    if (last == 0) __syncthreads();
    {
        // opaque copy of the thread index: otherwise the per-bit masks of the load addresses are kept (spilled) for the
        // whole kernel just to be reused here
        uint32_t tid_s = tid;
        asm volatile("" : "+r"(tid_s));
        const uint64_t gt = base ^ thread_offset64<NTB>(P.gl_store.thr, P.gl_store.c, tid_s);
        uint64_t gr[RB_MAX];
#pragma unroll
        for (int b = 0; b < RB; ++b) gr[b] = P.gl_store.reg[b];
        if constexpr (XCHG) {
#pragma unroll
            for (int j = 0; j < NS; ++j) {
                const XchgTarget t = xchg_target(slot_offset<RB>(gt, gr, j), xa.bit_mask, xa.keep, xa.stash_mask);
                reinterpret_cast<A *>(t.stays ? xa.out_mine : xa.out_peer)[t.base] = x[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < NS; ++j) gbase[slot_offset<RB>(gt, gr, j)] = x[j];
        }
    }
}

// ---- experiment (QSV_REGS_PERSIST=1, off: measured 10 % SLOWER than the kernel above, profiles/r2_ab_persist.txt) ----
// persistent CTAs, the next tile streams into the idle transposition buffer
// The register-tile kernel above loads a tile from HBM into registers, computes, stores: nothing of a CTA overlaps its own
// HBM latency, and with 128 registers there are only two CTAs per SM to cover for each other (ncu, round 1: 3.4 ms per
// sweep that the arithmetic does not hide).  Here a CTA is resident for the whole sweep (grid = CTAs that fit the GPU)
// and walks over tiles blockIdx.x, + gridDim.x, ...  The shared-memory transposition buffer is idle during the LAST
// pass of a tile (its result goes from registers straight to HBM), so as soon as the last transposition has been read
// every thread issues the 2^RB asynchronous copies (cp.async, LDGSTS: HBM -> shared memory, no registers) of ITS OWN
// amplitudes of the CTA's next tile into private slots (element j * NT + tid: conflict-free, and no barrier is needed
// to read them back).  The HBM latency of tile k+1 runs under the arithmetic of the last pass of tile k and under the
// stores; gate constants and descriptors are set up once per CTA instead of once per tile.
template <int BYTES> __device__ __forceinline__ void cp_async_elem(void *smem_dst, const void *gsrc) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void touch_value(double2 &v) { asm volatile("" : "+d"(v.x), "+d"(v.y)); }
__device__ __forceinline__ void touch_value(float2 &v) { asm volatile("" : "+f"(v.x), "+f"(v.y)); }

template <typename T, int RB, int MINB, bool XCHG = false>
__global__ void __launch_bounds__(1 << (RT_TB - RB), MINB)
    k_tile_regs_p(void *single, void *const *table, const __grid_constant__ RegProgram P,
                  const std::conditional_t<XCHG, XchgArgs, NoXchg> xa, int tiles_log2, uint32_t n_items) {
    using A = typename Cx<T>::type;
    constexpr int NS = 1 << RB;
    constexpr int NT = 1 << (RT_TB - RB);
    constexpr int NTB = RT_TB - RB;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    A *s = reinterpret_cast<A *>(smem_raw);
    const uint32_t tid = threadIdx.x;
    T *spool = reinterpret_cast<T *>(smem_raw + (sizeof(A) << RT_TB));
    for (int i = tid; i < P.pool_used; i += NT) spool[i] = (T)P.pool[i];

    const uint32_t tile_mask = (1u << tiles_log2) - 1u;
    auto item_base = [&](uint32_t item, A *&gb) -> uint64_t {
        gb = reinterpret_cast<A *>(table ? table[item >> tiles_log2] : single);
        if constexpr (XCHG) {
            if (xa.interleave)
                return expand_index((uint64_t)(item & tile_mask) >> 1, xa.holes2) | ((uint64_t)(item & 1u) << xa.bit_pos);
        }
        return expand_index((uint64_t)(item & tile_mask), P.tile_holes);
    };
    auto prefetch = [&](uint32_t item) {
        A *gb;
        const uint64_t b = item_base(item, gb);
        uint32_t tid_p = tid;
        asm volatile("" : "+r"(tid_p));  // opaque: the per-bit masks are recomputed here, not kept alive across the passes
        const uint64_t gt = b ^ thread_offset64<NTB>(P.gl_load.thr, P.gl_load.c, tid_p);
        uint64_t gr[RB_MAX];
#pragma unroll
        for (int k = 0; k < RB; ++k) gr[k] = P.gl_load.reg[k];
#pragma unroll
        for (int j = 0; j < NS; ++j) cp_async_elem<(int)sizeof(A)>(s + j * NT + tid_p, gb + slot_offset<RB>(gt, gr, j));
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t item = blockIdx.x;
    if (item < n_items) prefetch(item);
    __syncthreads();  // the constant pool is in place
    const int last = P.n_passes - 1;
    for (; item < n_items; item += gridDim.x) {
        A *gbase;
        const uint64_t base = item_base(item, gbase);
        const uint64_t outside = base | P.index_hi;
        const uint32_t next = item + gridDim.x;
        A x[NS];
        asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int j = 0; j < NS; ++j) x[j] = s[j * NT + tid];
        if (last == 0 && next < n_items) {
            // private slots: nobody else reads or writes them; the copies may only be issued once this thread's own reads
            // of the slots have returned (cp.async is not ordered after earlier loads of the issuing thread)
#pragma unroll
            for (int j = 0; j < NS; ++j) touch_value(x[j]);
            prefetch(next);
        }
        for (int pv = 0;; ++pv) {
            const int p = __shfl_sync(0xffffffffu, pv, 0);
            const RegPass &ps = P.passes[p];
            pass_compute<T, RB>(x, P, ps, tid, outside, spool);
            if (p == last) break;
            __syncthreads();  // every thread has read the previous layout (pass 0: its prefetched slots)
            {
                const uint32_t st = thread_offset<NTB>(ps.st_thr, ps.st_c, tid);
                uint32_t sr[RB_MAX];
#pragma unroll
                for (int b = 0; b < RB; ++b) sr[b] = ps.st_reg[b];
#pragma unroll
                for (int j = 0; j < NS; ++j) s[slot_offset<RB>(st, sr, j)] = x[j];
            }
            __syncthreads();
            {
                const RegPass &pn = P.passes[p + 1];
                const uint32_t st = thread_offset<NTB>(pn.ld_thr, pn.ld_c, tid);
                uint32_t sr[RB_MAX];
#pragma unroll
                for (int b = 0; b < RB; ++b) sr[b] = pn.ld_reg[b];
#pragma unroll
                for (int j = 0; j < NS; ++j) x[j] = s[slot_offset<RB>(st, sr, j)];
            }
            if (p + 1 == last && next < n_items) {
                __syncthreads();  // the buffer is idle from here on
                prefetch(next);
            }
        }
        // single pass: a folded permutation stores where another thread of this tile loaded; all of its copies have landed
        // once every thread is past its wait_group
        if (last == 0) __syncthreads();
        {
            uint32_t tid_s = tid;
            asm volatile("" : "+r"(tid_s));
            const uint64_t gt = base ^ thread_offset64<NTB>(P.gl_store.thr, P.gl_store.c, tid_s);
            uint64_t gr[RB_MAX];
#pragma unroll
            for (int b = 0; b < RB; ++b) gr[b] = P.gl_store.reg[b];
            if constexpr (XCHG) {
#pragma unroll
                for (int j = 0; j < NS; ++j) {
                    const XchgTarget t = xchg_target(slot_offset<RB>(gt, gr, j), xa.bit_mask, xa.keep, xa.stash_mask);
                    reinterpret_cast<A *>(t.stays ? xa.out_mine : xa.out_peer)[t.base] = x[j];
                }
            } else {
#pragma unroll
                for (int j = 0; j < NS; ++j) gbase[slot_offset<RB>(gt, gr, j)] = x[j];
            }
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
    int perm = 0;                 // index permutation folded into a pass boundary: 1 = X / CNOT on da, 2 = SWAP(da, db)
    bool mma_ok = false;          // uncontrolled 4x4 block that may run on two lane bits through FP64 tensor-core MMA
    std::vector<double> pool;     // what goes into the constant pool
};

// GF(2)-affine map on tile-local indices: e -> XOR_p e_p * col[p] ^ c
struct AffineMap {
    uint32_t col[RT_TB];
    uint32_t c = 0;
    AffineMap() {
        for (int p = 0; p < RT_TB; ++p) col[p] = 1u << p;
    }
    // compose with an index permutation applied AFTER this map
    void then(const NormGate &g) {
        auto f = [&](uint32_t v, bool is_const) {
            if (g.perm == 2) {
                const uint32_t a = (v >> g.da) & 1u, b = (v >> g.db) & 1u;
                return (v & ~((1u << g.da) | (1u << g.db))) | (b << g.da) | (a << g.db);
            }
            if (g.ctrl_loc == 0) return is_const ? v ^ (1u << g.da) : v;  // PauliX: a translation
            const uint32_t on = (v & g.ctrl_loc) == g.ctrl_loc ? 1u : 0u;   // single control: linear
            return v ^ (on << g.da);
        };
        for (int p = 0; p < RT_TB; ++p) col[p] = f(col[p], false);
        c = f(c, true);
    }
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

int env_int_regs(const char *name, int dflt);

template <typename T, int RB>
void launch_regs_t(State &sv, const RegProgram &P, void *const *table, int n_vecs, const XchgArgs *xa) {
    // register budget: 16 amplitudes per thread need 2 CTAs of 256 threads (128 registers); 8 amplitudes per thread run as
    // 2 CTAs of 512 threads (64 registers)
    constexpr int MINB = RB == 4 ? (sizeof(T) == 8 ? 2 : 3) : 2;
    const size_t smem = ((size_t)1 << RT_TB) * sizeof(typename Cx<T>::type) + RT_POOL * sizeof(T);
    const int tiles_log2 = sv.n - RT_TB;
    QSV_CHECK(!xa || (table == nullptr && n_vecs == 1), "internal: a sweep fused with an exchange works on one vector");
    static const int persist = env_int_regs("QSV_REGS_PERSIST", 0);  // measured slower on B200 (profiles/r2_ab_persist.txt)
    if (persist) {
        // generation 3: resident CTAs walking over the tiles, next tile prefetched into the idle transposition buffer
        const uint64_t n_items = (uint64_t)n_vecs << tiles_log2;
        static const int ctas_per_sm = env_int_regs("QSV_REGS_PERSIST_CTAS", MINB);
        const unsigned grid = (unsigned)std::min<uint64_t>(n_items, (uint64_t)NUM_SMS * ctas_per_sm);
        auto launch = [&](auto kern, bool &configured, auto xarg) {
            if (!configured) {
                QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                configured = true;
            }
            kern<<<grid, 1 << (RT_TB - RB), smem, sv.stream>>>(table ? nullptr : sv.data, table, P, xarg, tiles_log2,
                                                                   (uint32_t)n_items);
            QSV_CUDA(cudaGetLastError());
        };
        static bool conf_p[64] = {false}, conf_px[64] = {false};  // per device: function attributes belong to its context
        if (xa)
            launch(k_tile_regs_p<T, RB, MINB, true>, conf_px[sv.device & 63], *xa);
        else
            launch(k_tile_regs_p<T, RB, MINB, false>, conf_p[sv.device & 63], NoXchg{});
        return;
    }
    dim3 grid((unsigned)(1ull << tiles_log2), (unsigned)n_vecs);
    if (xa) {
        static bool configured_x[64] = {false};
        auto kern = k_tile_regs<T, RB, MINB, true>;
        if (!configured_x[sv.device & 63]) {
            QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            configured_x[sv.device & 63] = true;
        }
        kern<<<grid, 1 << (RT_TB - RB), smem, sv.stream>>>(sv.data, nullptr, P, *xa);
        QSV_CUDA(cudaGetLastError());
        return;
    }
    static bool configured[64] = {false};  // per device: function attributes belong to the device's context
    auto kern = k_tile_regs<T, RB, MINB>;
    if (!configured[sv.device & 63]) {
        QSV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[sv.device & 63] = true;
    }
    kern<<<grid, 1 << (RT_TB - RB), smem, sv.stream>>>(table ? nullptr : sv.data, table, P, NoXchg{});
    QSV_CUDA(cudaGetLastError());
}

int env_int_regs(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
}

bool env_flag(const char *name, int dflt) {
    const char *v = std::getenv(name);
    return (v ? std::atoi(v) : dflt) != 0;
}

int regs_rb() {
    const char *v = std::getenv("QSV_REGS_RB");
    const int rb = v ? std::atoi(v) : 4;
    return rb == 3 ? 3 : 4;
}

}  // namespace

static thread_local int t_beam_override = 0;  // > 0 while regs_sweep_model_cost prices candidates (greedy scheduler)

// tile bits above the low L ones: the needed high bits, then the lowest free bits (ascending)
std::vector<int> regs_tile_high_bits(int n, uint64_t need, int L) {
    std::vector<int> hi;
    for (int b = L; b < n; ++b)
        if (need >> b & 1) hi.push_back(b);
    for (int b = L; b < n && (int)hi.size() < RT_TB - L; ++b)
        if (!(need >> b & 1)) hi.push_back(b);
    std::sort(hi.begin(), hi.end());
    return hi;
}

bool regs_tile_contains_bit(int n, uint64_t need, int L, int bit) {
    if (bit < L) return true;
    const std::vector<int> hi = regs_tile_high_bits(n, need, L);
    return std::find(hi.begin(), hi.end(), bit) != hi.end();
}

// Program of one sweep of the register kernel over `gates` (all regs_fusable, dense-touched bits >= L listed in `need`).
void build_reg_program(int n, int dtype, uint64_t index_hi, const std::vector<const LoweredGate *> &gates, uint64_t need,
                       int L, int rb, RegProgram &P) {
    const int tb = RT_TB;
    QSV_CHECK(n >= tb, "internal: register tile kernel needs at least 12 local qubits");
    QSV_CHECK((int)gates.size() <= RT_MAX_GATES, "internal: too many gates in a sweep");
    QSV_CHECK(rb == 3 || rb == 4, "internal: 3 or 4 register bits");
    memset(&P, 0, sizeof(P));
    P.index_hi = index_hi;
    const bool fold = env_flag("QSV_REGS_FOLD", 1);
    const bool merge_diag = env_flag("QSV_REGS_UDIAG", 1);
    const bool diag_as_d1 = env_flag("QSV_REGS_DIAG1", 1);

    std::vector<int> hi = regs_tile_high_bits(n, need, L);
    QSV_CHECK((int)hi.size() == tb - L, "internal: tile bit selection");
    int pos[64];
    int gpos[RT_TB];  // global bit of tile-local position p
    for (int b = 0; b < 64; ++b) pos[b] = -1;
    std::vector<int> tile_bits;
    for (int b = 0; b < L; ++b) tile_bits.push_back(b);
    for (int b : hi) tile_bits.push_back(b);
    for (int p = 0; p < tb; ++p) {
        pos[tile_bits[p]] = p;
        gpos[p] = tile_bits[p];
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

    // index permutations that can be folded into a pass boundary: PauliX / CNOT with the control inside the tile, SWAP
    if (fold) {
        for (size_t i = 0; i < m; ++i) {
            NormGate &o = ng[i];
            if (o.ctrl_out != 0) continue;
            if (o.kind == RG_D1_SWAP && __builtin_popcount(o.ctrl_loc) <= 1) {
                o.perm = 1;
            } else if (o.kind == RG_D2 && o.ctrl_loc == 0) {
                static const double swap_m[16] = {1, 0, 0, 0, 0, 0, 1, 0, 0, 1, 0, 0, 0, 0, 0, 1};
                bool is_swap = true;
                for (int q = 0; q < 16; ++q) is_swap = is_swap && o.pool[2 * q] == swap_m[q] && o.pool[2 * q + 1] == 0.0;
                if (is_swap) o.perm = 2;
            }
        }
    }

    // complex128: uncontrolled 4x4 blocks may run on two lane bits through DMMA (one such gate per pass, applied first)
    // QSV_REGS_MMA: 0 = off, 1 = one 4x4 gate per pass, 2 (default) = a 4x4 BLOCK per pass: the pass scheduler multiplies
    // uncontrolled 2x2 and 4x4 gates on the block's two positions into it for as long as nothing else in the pass has touched
    // those positions (a 2x2 alone runs as U (x) 1: twice the flops, a quarter of the instructions)
    const int mma_mode = (dtype == QSV_C128 && rb == 4) ? env_int_regs("QSV_REGS_MMA", 2) : 0;
    const bool use_mma = mma_mode > 0;
    for (size_t i = 0; i < m; ++i) {
        const bool plain = ng[i].ctrl_loc == 0 && ng[i].ctrl_out == 0 && ng[i].perm == 0;
        const bool d1 = ng[i].kind == RG_D1 || ng[i].kind == RG_D1_REAL || ng[i].kind == RG_D1_RX;
        ng[i].mma_ok = use_mma && plain && (ng[i].kind == RG_D2 || (mma_mode >= 2 && d1));
    }

    // ---- list-schedule into passes -----------------------------------------------------------------
    // gate j depends on an earlier gate i when they share a tile bit that one of them touches non-diagonally
    std::vector<std::vector<int>> preds(m);
    for (size_t j = 0; j < m; ++j)
        for (size_t i = 0; i < j; ++i)
            if ((ng[i].dense & ng[j].bits) | (ng[i].bits & ng[j].dense)) preds[j].push_back((int)i);
    std::vector<char> done(m, 0);
    size_t n_done = 0;
    auto ready_in = [&](size_t j, const std::vector<char> &dn) {
        for (int i : preds[j])
            if (!dn[i]) return false;
        return true;
    };
    // folded permutations run at pass boundaries: slot s = before pass s (slot 0: the load from HBM, slot n_passes: the
    // store to HBM)
    std::vector<std::vector<int>> slot_perms;
    auto take_perms = [&]() {
        std::vector<int> v;
        for (bool progress = true; progress;) {
            progress = false;
            for (size_t j = 0; j < m; ++j)
                if (!done[j] && ng[j].perm && ready_in(j, done)) {
                    v.push_back((int)j);
                    done[j] = 1;
                    ++n_done;
                    progress = true;
                }
        }
        return v;
    };
    bool mma_allowed = use_mma;  // switched off for the rest of the sweep if the constant pool could not hold another block
    struct PassPick {
        uint32_t R = 0, F = 0, Pm = 0;  // register bits, forbidden positions, the (one or two) positions of the MMA block
        std::vector<int> mma;           // gates multiplied into the MMA block (applied first in the pass), in order
        std::vector<int> order;
    };
    auto grow = [&](int seed, std::vector<char> &dn, PassPick &pk, bool seed_may_be_block = true, bool with_block = true) {
        // greedy: keep adding the ready gate that needs the fewest new register bits.  Uncontrolled dense gates on positions
        // that nothing else in the pass has touched join the pass's tensor-core block (no register bits at all); the block
        // runs first, and every later gate of the pass that looks at its positions follows it in dependency order anyway.
        pk = PassPick();
        uint32_t touched_other = 0;  // positions looked at by the non-block gates of the pass
        auto mma_fits = [&](int j) {
            if (!with_block || !mma_allowed || !ng[j].mma_ok || (ng[j].dense & (pk.R | pk.F | touched_other))) return false;
            if (mma_mode == 1 && !pk.mma.empty()) return false;
            return __builtin_popcount(pk.Pm | ng[j].dense) <= 2;
        };
        auto add = [&](int j, bool as_mma) {
            pk.order.push_back(j);
            dn[j] = 1;
            if (as_mma) {
                pk.mma.push_back(j);
                pk.Pm |= ng[j].dense;
            } else {
                touched_other |= ng[j].bits;
                pk.R |= ng[j].dense;
                pk.F |= ng[j].forbid;
            }
        };
        add(seed, seed_may_be_block && mma_fits(seed));
        for (;;) {
            int next = -1, best_new = 99;
            bool next_mma = false;
            for (size_t j = 0; j < m; ++j) {
                if (dn[j] || ng[j].perm || !ready_in(j, dn)) continue;
                if (mma_fits((int)j)) {
                    next = (int)j;
                    next_mma = true;
                    break;
                }
                const uint32_t u = pk.R | ng[j].dense;
                if (__builtin_popcount(u) > rb || (u & (pk.F | ng[j].forbid)) || (ng[j].dense & pk.Pm)) continue;
                const int nw = __builtin_popcount(u) - __builtin_popcount(pk.R);
                if (nw < best_new) {
                    best_new = nw;
                    next = (int)j;
                    next_mma = false;
                }
            }
            if (next < 0) break;
            add(next, next_mma);
        }
    };
    int n_pool = 0;
    const int SW = dtype == QSV_C128 ? 3 : 4;
    const int ntb = tb - rb;
    int lay_r[RT_MAX_PASSES][RB_MAX], lay_t[RT_MAX_PASSES][NTB_MAX];
    slot_perms.push_back(take_perms());
    // Beam search over pass sequences (QSV_REGS_BEAM, default 4 from 28 qubits up, where its 0.5 ms of host time per 200 gates are
    // small against the passes it saves even when nothing overlaps them):
    // the greedy rule below -- the pass that retires most gates -- is short-sighted; a register pass costs 1 ms and a
    // tensor-core block another 0.85 ms at 30 qubits (tools/sweep_cost_model.py), so the cheapest sequence under that model
    // is searched for among the shortest ones and those one pass longer.
    std::vector<PassPick> planned;
    size_t plan_pos = 0;
    {
        const int beam = t_beam_override > 0 ? t_beam_override : std::max(1, env_int_regs("QSV_REGS_BEAM", n >= 28 ? 4 : 1));
        // cost of a sequence in 1/100 ms at 30 qubits complex128 (tools/sweep_cost_model.py, fitted to 17 measured
        // sweeps / circuits: a pass 0.96 ms, its tensor-core block 0.85 ms, a 4x4 / 2x2 gate on register bits 1.03 / 0.56 ms)
        auto pass_cost = [&](const PassPick &pk) {
            long c = 96 + (pk.mma.empty() ? 0 : 85);
            for (int j : pk.order) {
                if (std::find(pk.mma.begin(), pk.mma.end(), j) != pk.mma.end()) continue;
                c += ng[j].kind == RG_D2 ? 103 : (ng[j].kind == RG_DIAG ? 0 : (ng[j].kind == RG_D1_SWAP ? 0 : 56));
            }
            return c;
        };
        struct Node {
            std::vector<char> dn;
            std::vector<PassPick> seq;
            size_t retired = 0;
            long cost = 0;
        };
        auto sim_perms = [&](std::vector<char> &dn, size_t &retired) {
            for (bool progress = true; progress;) {
                progress = false;
                for (size_t j = 0; j < m; ++j)
                    if (!dn[j] && ng[j].perm && ready_in(j, dn)) {
                        dn[j] = 1;
                        ++retired;
                        progress = true;
                    }
            }
        };
        if (beam > 1 && n_done < m) {
            std::vector<Node> level(1);
            level[0].dn = done;
            level[0].retired = n_done;
            long best_cost = -1;
            int extra_levels = 1;  // sequences one pass longer than the shortest are still compared by cost
            for (int depth = 0; depth < RT_MAX_PASSES && !level.empty(); ++depth) {
                std::vector<Node> next;
                for (const Node &nd : level) {
                    for (size_t sd3 = 0; sd3 < 3 * m; ++sd3) {
                        // every ready gate as the seed: as the tensor-core block (if it may), on register bits with a block
                        // for others, or in a pass without a block (a block costs more than one 2x2 gate on register bits)
                        const size_t sd = sd3 / 3;
                        const int variant = (int)(sd3 % 3);
                        if (nd.dn[sd] || ng[sd].perm || !ready_in(sd, nd.dn)) continue;
                        if (variant == 1 && !(mma_allowed && ng[sd].mma_ok)) continue;  // same pass as variant 0
                        if (variant == 2 && !mma_allowed) continue;                     // same pass as variant 0
                        Node c;
                        c.dn = nd.dn;
                        PassPick pk;
                        grow((int)sd, c.dn, pk, variant == 0, variant != 2);
                        c.retired = nd.retired + pk.order.size();
                        c.cost = nd.cost + pass_cost(pk);
                        sim_perms(c.dn, c.retired);
                        bool dup = false;
                        for (Node &o : next)
                            if (o.dn == c.dn) {
                                dup = true;
                                if (c.cost < o.cost) {  // same gates retired, cheaper way
                                    o.cost = c.cost;
                                    o.seq = nd.seq;
                                    o.seq.push_back(pk);
                                }
                                break;
                            }
                        if (dup) continue;
                        c.seq = nd.seq;
                        c.seq.push_back(std::move(pk));
                        next.push_back(std::move(c));
                    }
                }
                // finished sequences compete by cost; unfinished ones go on, the most advanced (then cheapest) first
                std::vector<Node> open;
                for (Node &c : next) {
                    if (c.retired == m) {
                        if (best_cost < 0 || c.cost < best_cost) {
                            best_cost = c.cost;
                            planned = c.seq;
                        }
                    } else {
                        open.push_back(std::move(c));
                    }
                }
                if (best_cost >= 0 && extra_levels-- <= 0) break;
                std::stable_sort(open.begin(), open.end(), [](const Node &a, const Node &b) {
                    return a.retired != b.retired ? a.retired > b.retired : a.cost < b.cost;
                });
                if ((int)open.size() > beam) open.resize(beam);
                level = std::move(open);
            }
        }
    }
    while (n_done < m || P.n_passes == 0) {
        PassPick best;
        if (plan_pos < planned.size()) {
            best = planned[plan_pos++];
        } else {
            // try every ready gate as the seed of the next pass, keep the pass that retires the most gates
            for (size_t sd = 0; sd < m; ++sd) {
                if (done[sd] || ng[sd].perm || !ready_in(sd, done)) continue;
                std::vector<char> dn = done;
                PassPick pk;
                grow((int)sd, dn, pk);
                if (pk.order.size() > best.order.size()) best = pk;
            }
        }
        if (!best.mma.empty()) {
            // every remaining gate may still need its own constants: keep room for them
            size_t need_pool = (size_t)n_pool + 66;
            for (size_t j = 0; j < m; ++j)
                if (!done[j] && std::find(best.mma.begin(), best.mma.end(), (int)j) == best.mma.end()) need_pool += ng[j].pool.size();
            if (need_pool > (size_t)RT_POOL) {
                mma_allowed = false;
                planned.clear();  // the searched sequence assumed tensor-core blocks: fall back to the greedy rule
                plan_pos = 0;
                continue;  // choose this pass again without a tensor-core block
            }
        }
        const std::vector<int> &best_order = best.order;
        // the MMA block's two positions (ma = matrix MSB); a block on one position gets the lowest free position as partner
        int ma = -1, mb = -1;
        if (!best.mma.empty()) {
            uint32_t pm = best.Pm;
            for (int p = 0; p < tb && __builtin_popcount(pm) < 2; ++p)
                if (!((pm | best.R | best.F) >> p & 1)) pm |= 1u << p;
            QSV_CHECK(__builtin_popcount(pm) == 2, "internal: no partner position for the tensor-core block");
            best.Pm = pm;
            mb = __builtin_ctz(pm);
            ma = 31 - __builtin_clz(pm);
        }
        const uint32_t best_R = best.R, best_F = best.F | best.Pm;  // register bits stay away from the MMA block's lane bits
        QSV_CHECK(!best_order.empty() || n_done == m, "internal: pass scheduling made no progress");
        QSV_CHECK(P.n_passes < RT_MAX_PASSES, "internal: too many passes in a sweep");
        const int pi = P.n_passes++;
        RegPass &ps = P.passes[pi];
        // register bits: the dense bits of the pass, filled up with the highest other positions -- preferably ones that no
        // diagonal gate of the pass looks at, so that those gates stay thread-uniform
        uint32_t avoid = 0;
        for (int gi : best_order)
            if (ng[gi].kind == RG_DIAG) avoid |= ng[gi].bits;
        uint32_t R = best_R;
        for (int round = 0; round < 2; ++round)
            for (int p = tb - 1; p >= 0 && __builtin_popcount(R) < rb; --p)
                if (!((R | best_F) >> p & 1) && (round == 1 || !(avoid >> p & 1))) R |= 1u << p;
        QSV_CHECK(__builtin_popcount(R) == rb, "internal: no free register bits for a pass");
        int regbit_of[16];
        for (int p = 0, k = 0; p < tb; ++p) {
            regbit_of[p] = -1;
            if (R >> p & 1) {
                regbit_of[p] = k;
                lay_r[pi][k++] = p;
            }
        }
        // thread bits: lanes take the lowest positions (coalescing); among them, the first SW lane bits get
        // distinct positions mod SW when possible (conflict-free swizzled shared-memory accesses)
        std::vector<int> rest;
        if (ma >= 0) {  // thread bit 0 = matrix LSB, thread bit 1 = matrix MSB of the tensor-core block
            rest.push_back(mb);
            rest.push_back(ma);
        }
        for (int p = 0; p < tb; ++p)
            if (!(R >> p & 1) && !(best.Pm >> p & 1)) rest.push_back(p);
        std::vector<int> lanes(rest.begin(), rest.begin() + 5), ordered;
        std::vector<char> used(5, 0);
        uint32_t seen = 0;
        if (ma >= 0)
            for (int q = 0; q < 2; ++q) {
                ordered.push_back(lanes[q]);
                used[q] = 1;
                seen |= 1u << (lanes[q] % SW);
            }
        for (int q = 0; q < 5 && (int)ordered.size() < SW; ++q)
            if (!used[q] && !(seen >> (lanes[q] % SW) & 1)) {
                seen |= 1u << (lanes[q] % SW);
                ordered.push_back(lanes[q]);
                used[q] = 1;
            }
        for (int q = 0; q < 5; ++q)
            if (!used[q]) ordered.push_back(lanes[q]);
        for (size_t q = 5; q < rest.size(); ++q) ordered.push_back(rest[q]);
        int thrbit_of[16];
        for (int p = 0; p < tb; ++p) thrbit_of[p] = -1;
        for (int k = 0; k < ntb; ++k) {
            lay_t[pi][k] = ordered[k];
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
        auto thread_uniform = [&](const NormGate &o) {
            if (!merge_diag || o.kind != RG_DIAG) return false;
            unsigned reg, thr;
            split(o.ctrl_loc | o.mask_loc[0] | o.mask_loc[1], reg, thr);
            return reg == 0;
        };
        ps.mma_off = NO_MMA;
        if (ma >= 0) {
            QSV_CHECK(thrbit_of[mb] == 0 && thrbit_of[ma] == 1, "internal: lane bits of the tensor-core block");
            // product of the member gates as a 4x4 in the (ma, mb) basis
            cplx M4[16];
            for (int q = 0; q < 16; ++q) M4[q] = (q / 4 == q % 4) ? cplx(1.0, 0.0) : cplx(0.0, 0.0);
            for (int gi : best.mma) {
                const NormGate &o = ng[gi];
                cplx G[16];
                if (o.kind == RG_D2) {
                    const bool same = o.da == ma;
                    QSV_CHECK((same && o.db == mb) || (o.da == mb && o.db == ma), "internal: tensor-core block positions");
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) {
                            const int sr = same ? r : ((r & 1) << 1) | (r >> 1), sc = same ? c : ((c & 1) << 1) | (c >> 1);
                            G[r * 4 + c] = cplx(o.pool[2 * (sr * 4 + sc)], o.pool[2 * (sr * 4 + sc) + 1]);
                        }
                } else {
                    cplx U[4];
                    for (int q = 0; q < 4; ++q) U[q] = cplx(o.pool[2 * q], o.pool[2 * q + 1]);
                    const bool on_msb = o.da == ma;
                    QSV_CHECK(on_msb || o.da == mb, "internal: tensor-core block position");
                    for (int r = 0; r < 4; ++r)
                        for (int c = 0; c < 4; ++c) {
                            const int rh = r >> 1, rl = r & 1, ch = c >> 1, cl = c & 1;
                            G[r * 4 + c] = on_msb ? (rl == cl ? U[rh * 2 + ch] : cplx(0.0, 0.0))
                                                  : (rh == ch ? U[rl * 2 + cl] : cplx(0.0, 0.0));
                        }
                }
                cplx T4[16];
                for (int r = 0; r < 4; ++r)
                    for (int c = 0; c < 4; ++c) {
                        cplx acc(0.0, 0.0);
                        for (int k = 0; k < 4; ++k) acc += G[r * 4 + k] * M4[k * 4 + c];
                        T4[r * 4 + c] = acc;
                    }
                for (int q = 0; q < 16; ++q) M4[q] = T4[q];
                done[gi] = 1;
                ++n_done;
                ++P.n_mma_gates;
            }
            // real 8x8 form over (re0, im0, ..., re3, im3), amplitude index t = 2 * bit(ma) + bit(mb)
            QSV_CHECK(n_pool + 66 <= RT_POOL, "internal: constant pool overflow");
            n_pool = (n_pool + 1) & ~1;  // 16-byte aligned pairs
            ps.mma_off = (unsigned short)n_pool;
            for (int t = 0; t < 4; ++t)
                for (int u = 0; u < 4; ++u) {
                    const double re = M4[t * 4 + u].real(), im = M4[t * 4 + u].imag();
                    P.pool[n_pool + (2 * t) * 8 + 2 * u] = re;
                    P.pool[n_pool + (2 * t) * 8 + 2 * u + 1] = -im;
                    P.pool[n_pool + (2 * t + 1) * 8 + 2 * u] = im;
                    P.pool[n_pool + (2 * t + 1) * 8 + 2 * u + 1] = re;
                }
            n_pool += 64;
        }
        ps.gate_begin = (unsigned short)P.n_gates;
        for (int phase = 0; phase < 2; ++phase) {
            // phase 0: the thread-uniform diagonal gates (they commute with every other gate of the pass), phase 1: the rest
            for (int gi : best_order) {
                const NormGate &o = ng[gi];
                if (std::find(best.mma.begin(), best.mma.end(), gi) != best.mma.end()) continue;
                if (thread_uniform(o) != (phase == 0)) continue;
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
                    // a one-bit phase table with exactly one register bit runs as a diagonal 2x2 on that bit
                    if (diag_as_d1 && o.n_table_bits == 1 && __builtin_popcount(t.reg_mask[1]) == 1) {
                        t.kind = RG_D1_DIAG;
                        t.ra = (unsigned char)__builtin_ctz(t.reg_mask[1]);
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
            if (phase == 0) ps.udiag_end = (unsigned short)P.n_gates;
        }
        ps.gate_end = (unsigned short)P.n_gates;
        slot_perms.push_back(take_perms());
    }
    QSV_CHECK((int)slot_perms.size() == P.n_passes + 1, "internal: permutation slots");
    for (const auto &v : slot_perms) P.n_folded += (int)v.size();

    // ---- address maps --------------------------------------------------------------------------------
    auto to_global = [&](uint32_t v) {
        uint64_t o = 0;
        for (int p = 0; p < tb; ++p)
            if (v >> p & 1) o |= 1ull << gpos[p];
        return o;
    };
    // Shared-memory swizzle of one transposition: low SW bits ^= XOR over the set high bits p of lcol[p].  Every choice of
    // lcol is a bijection; a warp-wide 128-bit (complex64: 64-bit) access is conflict-free when the first SW thread bits --
    // the lanes of one access phase -- land on linearly independent low parts.  The store of pass p (with the folded
    // permutations in its columns) and the load of pass p + 1 share the buffer, so one lcol must serve both lane sets.
    struct Swizzle {
        unsigned char lcol[RT_TB];
    };
    auto swz_with = [&](uint32_t v, const Swizzle &sw) {
        uint32_t low = v & ((1u << SW) - 1u);
        for (int p = SW; p < tb; ++p)
            if (v >> p & 1) low ^= sw.lcol[p];
        return (v & ~((1u << SW) - 1u)) | low;
    };
    auto default_swizzle = [&]() {
        Swizzle sw;
        for (int p = 0; p < tb; ++p) sw.lcol[p] = (unsigned char)(p < SW ? 0 : 1u << (p % SW));
        return sw;
    };
    auto low_rank = [&](const uint32_t *vecs, int k, const Swizzle &sw) {  // rank over GF(2) of the swizzled low parts
        uint32_t basis[8] = {0};
        int rank = 0;
        for (int i = 0; i < k; ++i) {
            uint32_t v = swz_with(vecs[i], sw) & ((1u << SW) - 1u);
            for (int b = SW - 1; b >= 0 && v; --b)
                if (v >> b & 1) {
                    if (basis[b])
                        v ^= basis[b];
                    else {
                        basis[b] = v;
                        ++rank;
                        v = 0;
                    }
                }
        }
        return rank;
    };
    const bool tune_swizzle = env_flag("QSV_REGS_SWIZZLE_SEARCH", 1);
    auto choose_swizzle = [&](const AffineMap &store_map, int p_store, int p_load) {
        Swizzle best = default_swizzle();
        if (!tune_swizzle) return best;
        uint32_t sv_[4], lv_[4];
        for (int i = 0; i < SW; ++i) {
            sv_[i] = store_map.col[lay_t[p_store][i]];
            lv_[i] = 1u << lay_t[p_load][i];
        }
        int best_score = low_rank(sv_, SW, best) + low_rank(lv_, SW, best);
        uint32_t rng = 0x9e3779b9u;
        for (int tries = 0; tries < 256 && best_score < 2 * SW; ++tries) {
            Swizzle cand;
            for (int p = 0; p < tb; ++p) {
                rng = rng * 1664525u + 1013904223u;
                cand.lcol[p] = (unsigned char)(p < SW ? 0 : (rng >> 24) & ((1u << SW) - 1u));
            }
            const int score = low_rank(sv_, SW, cand) + low_rank(lv_, SW, cand);
            if (score > best_score) {
                best_score = score;
                best = cand;
            }
        }
        return best;
    };
    auto smem_cols = [&](const AffineMap &mp, int pi, const Swizzle &sw, unsigned short *thr, unsigned short *reg,
                         unsigned short &c) {
        for (int k = 0; k < ntb; ++k) thr[k] = (unsigned short)swz_with(mp.col[lay_t[pi][k]], sw);
        for (int k = 0; k < rb; ++k) reg[k] = (unsigned short)swz_with(mp.col[lay_r[pi][k]], sw);
        c = (unsigned short)swz_with(mp.c, sw);
    };
    auto global_cols = [&](const AffineMap &mp, int pi, GlobalMap &g) {
        for (int k = 0; k < ntb; ++k) g.thr[k] = to_global(mp.col[lay_t[pi][k]]);
        for (int k = 0; k < rb; ++k) g.reg[k] = to_global(mp.col[lay_r[pi][k]]);
        g.c = to_global(mp.c);
    };
    for (int pi = 0; pi < P.n_passes; ++pi) {
        RegPass &ps = P.passes[pi];
        AffineMap after;  // permutations between this pass and the next: the value held for index e is stored at pi(e)
        for (int gi : slot_perms[pi + 1]) after.then(ng[gi]);
        if (pi + 1 < P.n_passes) {
            const Swizzle sw = choose_swizzle(after, pi, pi + 1);
            smem_cols(after, pi, sw, ps.st_thr, ps.st_reg, ps.st_c);
            RegPass &pn = P.passes[pi + 1];
            smem_cols(AffineMap(), pi + 1, sw, pn.ld_thr, pn.ld_reg, pn.ld_c);
        } else {
            global_cols(after, pi, P.gl_store);
        }
    }
    {
        // slot 0: the first pass wants S'[e] = S[pi^-1(e)]; every folded permutation is an involution, so the inverse is the
        // same gates composed in reverse order
        AffineMap inv;
        for (auto it = slot_perms[0].rbegin(); it != slot_perms[0].rend(); ++it) inv.then(ng[*it]);
        global_cols(inv, 0, P.gl_load);
    }
    P.pool_used = n_pool;
    {
        // a rough tile time divided by the CTAs per SM: HBM share of a 64 KiB tile + FP pipe time of the gates + transposes, in SM clocks
        double fp = 0.0;
        for (int gi = 0; gi < P.n_gates; ++gi) {
            const int k = P.gates[gi].kind;
            fp += k == RG_D2 ? 256 : k == RG_D1 ? 128 : k == RG_D1_SWAP ? 0 : 64;
        }
        for (int pi = 0; pi < P.n_passes; ++pi) fp += P.passes[pi].mma_off != NO_MMA ? 256 : 0;
        const double amp = dtype == QSV_C128 ? 1.0 : 0.5;
        const double clk = amp * (5800.0 + 4.0 * fp + 1000.0 * (P.n_passes - 1));
        // QSV_REGS_STAGGER_NS: 0 (default) = off, -1 = the estimate above, > 0 = nanoseconds per CTA slot.  Measured on B200
        // (profiles/r1_ab_mma.txt): no effect at any setting -- the resident CTAs are not in lockstep; kept as an experiment.
        const int forced = env_int_regs("QSV_REGS_STAGGER_NS", 0);
        const int ctas_per_sm = (dtype == QSV_C128 || rb == 3) ? 2 : 3;  // MINB of launch_regs_t
        P.stagger_ns = forced >= 0 ? (unsigned)forced : (unsigned)(clk / ctas_per_sm / 1.9);
    }
    for (int i = 0; i < n_pool; ++i) P.poolf[i] = (float)P.pool[i];
    P.uniform_consts = env_flag("QSV_REGS_UCONST", 1) ? 1 : 0;
    P.prefetch = std::max(0, env_int_regs("QSV_REGS_PREFETCH", 0));
}

double regs_sweep_model_cost(int n, int dtype, const std::vector<const LoweredGate *> &gates, uint64_t need, int L,
                             bool greedy_scheduler) {
    static thread_local RegProgram P;
    t_beam_override = greedy_scheduler ? 1 : 0;
    struct Reset {
        ~Reset() { t_beam_override = 0; }
    } reset;
    build_reg_program(n, dtype, 0, gates, need, L, regs_rb(), P);
    double c = 3.39 + 0.96 * P.n_passes;
    for (int p = 0; p < P.n_passes; ++p) c += P.passes[p].mma_off != NO_MMA ? 0.85 : 0.0;
    for (int g = 0; g < P.n_gates; ++g) {
        const int k = P.gates[g].kind;
        c += k == RG_D2 ? 1.03 : (k == RG_DIAG || k == RG_D1_DIAG) ? 0.20 : (k == RG_D1_SWAP ? 0.0 : 0.56);
    }
    return c;
}

// FP64 (FP32 for complex64) fused multiply-adds per amplitude that the program of one fused sweep executes, and its
// passes: 2x2 gate 8 (real matrix or RX 4), 4x4 gate 16, tensor-core block 16 (the real 8x8 form on 4 amplitudes),
// diagonal gate 4 (one complex multiplication); a controlled gate works on 1 / 2^controls of the amplitudes.
void regs_sweep_work(int n, int dtype, const std::vector<const LoweredGate *> &gates, uint64_t need, int L, double *fma_per_amp,
                     int *passes) {
    static thread_local RegProgram P;
    build_reg_program(n, dtype, 0, gates, need, L, regs_rb(), P);
    double f = 0.0;
    for (int p = 0; p < P.n_passes; ++p) f += P.passes[p].mma_off != NO_MMA ? 16.0 : 0.0;
    for (int g = 0; g < P.n_gates; ++g) {
        const RegGate &t = P.gates[g];
        const double w = t.kind == RG_D2 ? 16.0 : (t.kind == RG_D1 ? 8.0 : (t.kind == RG_D1_SWAP ? 0.0 : 4.0));
        const int n_ctrl = __builtin_popcount(t.ctrl_reg) + __builtin_popcount(t.ctrl_thr) + __builtin_popcountll(t.out_ctrl);
        f += w / (double)(1ull << n_ctrl);
    }
    if (fma_per_amp) *fma_per_amp = f;
    if (passes) *passes = P.n_passes;
}

bool regs_pull_supported() {
    static const bool persist = env_int_regs("QSV_REGS_PERSIST", 0) != 0;
    return !persist;  // the persistent variant streams its tiles with cp.async and has no routed loads
}

void run_sweep_regs(State &sv, const std::vector<const LoweredGate *> &gates, uint64_t need, int L, void *const *table,
                    int n_vecs, const FusedExchange *fx, int fx_idx, const FusedPull *pull) {
    const int rb = regs_rb();
    RegProgram P;
    build_reg_program(sv.n, sv.dtype, sv.index_hi, gates, need, L, rb, P);
    static unsigned launch_epoch = 0;
    P.epoch = ++launch_epoch == 0 ? ++launch_epoch : launch_epoch;  // never 0: the counters start zeroed
    sv.stat_launches += 1;
    sv.stat_sweeps += 1;
    XchgArgs xa_store;
    const XchgArgs *xa = nullptr;
    if (fx || pull) {
        memset(&xa_store, 0, sizeof(xa_store));
        xa_store.holes2 = P.tile_holes;
        if (pull) {
            QSV_CHECK(regs_pull_supported(), "internal: this sweep kernel cannot fetch from the partner");
            xa_store.in_peer = pull->in_peer;
            xa_store.pull_bit_mask = 1ull << pull->local_bit;
            xa_store.pull_keep = pull->my_value ? xa_store.pull_bit_mask : 0ull;
            xa_store.pull_stash_mask = 1ull << pull->stash_bit;
            xa_store.out_mine = pull->out_mine;  // a plain out-of-place sweep unless it pushes as well
        }
        if (fx) {
            xa_store.out_mine = fx->out_mine[fx_idx];
            xa_store.out_peer = fx->out_peer[fx_idx];
            xa_store.bit_mask = 1ull << fx->local_bit;
            xa_store.keep = fx->my_value ? xa_store.bit_mask : 0ull;
            xa_store.stash_mask = fx->stash_bit >= 0 ? 1ull << fx->stash_bit : 0ull;
            xa_store.bit_pos = fx->local_bit;
            static const bool interleave_ok = env_flag("QSV_DIST_XCHG_INTERLEAVE", 1);
            if (interleave_ok && !regs_tile_contains_bit(sv.n, need, L, fx->local_bit) && sv.n > RT_TB &&
                P.tile_holes.n < MAX_HOLES) {
                // tile bits + the exchanged bit, ascending
                int pos[MAX_HOLES + 1], m = 0;
                bool placed = false;
                for (int j = 0; j < P.tile_holes.n; ++j) {
                    if (!placed && fx->local_bit < (int)P.tile_holes.pos[j]) {
                        pos[m++] = fx->local_bit;
                        placed = true;
                    }
                    pos[m++] = P.tile_holes.pos[j];
                }
                if (!placed) pos[m++] = fx->local_bit;
                xa_store.holes2 = make_holes(pos, m, 0);
                xa_store.interleave = 1;
            }
        }
        xa = &xa_store;
    }
    if (sv.dtype == QSV_C128) {
        if (rb == 4)
            launch_regs_t<double, 4>(sv, P, table, n_vecs, xa);
        else
            launch_regs_t<double, 3>(sv, P, table, n_vecs, xa);
    } else {
        if (rb == 4)
            launch_regs_t<float, 4>(sv, P, table, n_vecs, xa);
        else
            launch_regs_t<float, 3>(sv, P, table, n_vecs, xa);
    }
}


// ---- copy-pass forms of the two halves of an exchange (a batch whose first / last sweep cannot carry them) ------------
namespace {
// 16-byte units; masks are in units
template <int U>
__global__ void __launch_bounds__(256)
    k_xchg_push_copy(const uint4 *__restrict__ in, uint4 *__restrict__ out_mine, uint4 *__restrict__ out_peer, uint64_t count,
                     uint64_t bit_mask, uint64_t keep, uint64_t stash_mask) {
    const uint64_t stride = (uint64_t)gridDim.x * 256 * U;
    for (uint64_t i0 = (uint64_t)blockIdx.x * 256 * U + threadIdx.x; i0 < count; i0 += stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t)u * 256;
            if (i < count) v[u] = in[i];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t)u * 256;
            if (i < count) {
                const XchgTarget t = xchg_target(i, bit_mask, keep, stash_mask);
                (t.stays ? out_mine : out_peer)[t.base] = v[u];
            }
        }
    }
}
template <int U>
__global__ void __launch_bounds__(256)
    k_xchg_pull_copy(const uint4 *__restrict__ in_mine, const uint4 *__restrict__ in_peer, uint4 *__restrict__ out, uint64_t count,
                     uint64_t bit_mask, uint64_t keep, uint64_t stash_mask) {
    const uint64_t stride = (uint64_t)gridDim.x * 256 * U;
    for (uint64_t i0 = (uint64_t)blockIdx.x * 256 * U + threadIdx.x; i0 < count; i0 += stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t)u * 256;
            if (i < count) {
                const XchgTarget t = xchg_source(i, bit_mask, keep, stash_mask);
                v[u] = (t.stays ? in_mine : in_peer)[t.base];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = i0 + (uint64_t)u * 256;
            if (i < count) out[i] = v[u];
        }
    }
}
}  // namespace

void launch_xchg_push_copy(State &sv, const void *in, void *out_mine, void *out_peer, int local_bit, int my_value,
                           int stash_bit) {
    sv.use();
    const int shift = sv.dtype == QSV_C128 ? 0 : 1;  // complex64: two amplitudes per 16-byte unit
    QSV_CHECK(local_bit >= shift && (stash_bit < 0 || stash_bit >= shift), "internal: exchanged bit below the copy unit");
    const uint64_t count = sv.length() >> shift;
    const uint64_t bit_mask = 1ull << (local_bit - shift);
    constexpr int U = 4;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((count + 256 * U - 1) / (256 * U), NUM_SMS * 16));
    k_xchg_push_copy<U><<<grid, 256, 0, sv.stream>>>((const uint4 *)in, (uint4 *)out_mine, (uint4 *)out_peer, count, bit_mask,
                                                     my_value ? bit_mask : 0ull, stash_bit >= 0 ? 1ull << (stash_bit - shift) : 0ull);
    QSV_CUDA(cudaGetLastError());
    sv.stat_launches += 1;
    sv.stat_sweeps += 1;
}

void launch_xchg_pull_copy(State &sv, const void *in_mine, const void *in_peer, void *out, int local_bit, int my_value,
                           int stash_bit) {
    sv.use();
    const int shift = sv.dtype == QSV_C128 ? 0 : 1;
    QSV_CHECK(local_bit >= shift && stash_bit >= shift, "internal: exchanged bit below the copy unit");
    const uint64_t count = sv.length() >> shift;
    const uint64_t bit_mask = 1ull << (local_bit - shift);
    constexpr int U = 4;
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((count + 256 * U - 1) / (256 * U), NUM_SMS * 16));
    k_xchg_pull_copy<U><<<grid, 256, 0, sv.stream>>>((const uint4 *)in_mine, (const uint4 *)in_peer, (uint4 *)out, count, bit_mask,
                                                     my_value ? bit_mask : 0ull, 1ull << (stash_bit - shift));
    QSV_CUDA(cudaGetLastError());
    sv.stat_launches += 1;
    sv.stat_sweeps += 1;
}

}  // namespace qsv
