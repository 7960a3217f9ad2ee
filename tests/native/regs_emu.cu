// TEST-ONLY CPU emulator of the register-blocked fused tile kernel (csrc/tile_regs.cu).
//
// It is built into tests/native/_build/libregs_emu.so by tests/native/build_emu.py, is never linked into or loaded by
// the product (pennylane_lightning_gpu_b200/), and exists so that the HOST side of the fused executor -- gate merging,
// DAG sweep packing, pass scheduling, folded index permutations, address maps, program encoding -- can be checked
// against the NumPy oracle on a machine without a GPU.  The per-thread arithmetic is the kernel's own
// (__host__ __device__ code of csrc/tile_regs_core.cuh); only the thread/CTA loops and the memories are emulated.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../pennylane_lightning_gpu_b200/csrc/qsv_internal.h"
#include "../../pennylane_lightning_gpu_b200/csrc/tile_regs_core.cuh"

using namespace qsv;
using namespace qsv::rt;

namespace {

// plain reference application of one lowered gate (lone gates go to the one-sweep kernels on the GPU)
template <typename A> void apply_lowered_host(std::vector<A> &psi, int n, const LoweredGate &g) {
    using T = decltype(A().x);
    const uint64_t N = 1ull << n;
    auto cmul = [](cplx a, A b) { return cplx(a.real() * b.x - a.imag() * b.y, a.real() * b.y + a.imag() * b.x); };
    if (g.kind == LoweredGate::DIAG || g.kind == LoweredGate::PARITY) {
        for (uint64_t i = 0; i < N; ++i) {
            if ((i & g.ctrl_mask) != g.ctrl_mask) continue;
            int t = 0;
            if (g.kind == LoweredGate::PARITY) {
                t = __builtin_popcountll(i & g.zmask) & 1;
            } else {
                for (int b = 0; b < g.k; ++b) t = (t << 1) | (int)((i >> g.tgt_bits[b]) & 1);
            }
            const cplx r = cmul(g.mat[t], psi[i]);
            psi[i].x = (T)r.real();
            psi[i].y = (T)r.imag();
        }
    } else if (g.kind == LoweredGate::DENSE) {
        Holes h = make_holes(g.holes.data(), (int)g.holes.size(), 0);
        const uint64_t groups = N >> g.holes.size();
        const int d = 1 << g.k;
        std::vector<cplx> in(d), out(d);
        for (uint64_t o = 0; o < groups; ++o) {
            const uint64_t base = expand_index(o, h) | g.ctrl_mask;
            for (int a = 0; a < d; ++a) in[a] = cplx(psi[base + g.offs[a]].x, psi[base + g.offs[a]].y);
            for (int r = 0; r < d; ++r) {
                cplx s = 0;
                for (int c = 0; c < d; ++c) s += g.mat[r * d + c] * in[c];
                psi[base + g.offs[r]].x = (T)s.real();
                psi[base + g.offs[r]].y = (T)s.imag();
            }
        }
    }
}

// what the XCHG = true kernel does with the stores of its last pass (null out_mine: the plain in-place kernel)
template <typename A> struct EmuXchg {
    A *out_mine = nullptr;
    A *out_peer = nullptr;
    uint64_t bit_mask = 0, keep = 0, stash_mask = 0;
    // pull side of a split exchange (k_tile_regs: xa.in_peer / pull_*): the first pass reads through xchg_source
    const A *in_peer = nullptr;
    uint64_t pull_bit_mask = 0, pull_keep = 0, pull_stash_mask = 0;
};

template <typename T, int RB>
void emulate_program(std::vector<typename Cx<T>::type> &psi, int n, const RegProgram &P,
                     const EmuXchg<typename Cx<T>::type> &xc = EmuXchg<typename Cx<T>::type>()) {
    using A = typename Cx<T>::type;
    constexpr int NS = 1 << RB, NT = 1 << (TB - RB), NTB = TB - RB;
    std::vector<A> smem(1 << TB);
    std::vector<T> spool(POOL);
    for (int i = 0; i < P.pool_used; ++i) spool[i] = (T)P.pool[i];
    std::vector<A> xs((size_t)NT * NS);
    const uint64_t tiles = 1ull << (n - TB);
    for (uint64_t blk = 0; blk < tiles; ++blk) {
        const uint64_t base = expand_index(blk, P.tile_holes);
        const uint64_t outside = base | P.index_hi;
        for (int p = 0; p < P.n_passes; ++p) {
            const RegPass &ps = P.passes[p];
            const bool first = p == 0, last = p == P.n_passes - 1;
            for (uint32_t tid = 0; tid < (uint32_t)NT; ++tid) {
                A *x = &xs[(size_t)tid * NS];
                if (first) {
                    const uint64_t gt = base ^ thread_offset64<NTB>(P.gl_load.thr, P.gl_load.c, tid);
                    for (int j = 0; j < NS; ++j) {
                        const uint64_t off = slot_offset<RB>(gt, P.gl_load.reg, j);
                        if (xc.pull_bit_mask != 0) {
                            const XchgTarget t = xchg_source(off, xc.pull_bit_mask, xc.pull_keep, xc.pull_stash_mask);
                            x[j] = t.stays ? psi[t.base] : xc.in_peer[t.base];
                        } else {
                            x[j] = psi[off];
                        }
                    }
                } else {
                    const uint32_t st = thread_offset<NTB>(ps.ld_thr, ps.ld_c, tid);
                    uint32_t sr[RB_MAX] = {0, 0, 0, 0};
                    for (int b = 0; b < RB; ++b) sr[b] = ps.ld_reg[b];
                    for (int j = 0; j < NS; ++j) x[j] = smem[slot_offset<RB>(st, sr, j)];
                }
            }
            if (ps.mma_off != NO_MMA) {
                // the tensor-core gate of the pass: threads 4m .. 4m+3 hold amplitudes t = 0..3 of quad m (per slot); the real
                // 8x8 table acts on (re0, im0, ..., re3, im3)
                const double *R = P.pool + ps.mma_off;
                for (uint32_t t0 = 0; t0 < (uint32_t)NT; t0 += 4)
                    for (int j = 0; j < NS; ++j) {
                        double in[8], out[8];
                        for (int u = 0; u < 4; ++u) {
                            in[2 * u] = (double)xs[(size_t)(t0 + u) * NS + j].x;
                            in[2 * u + 1] = (double)xs[(size_t)(t0 + u) * NS + j].y;
                        }
                        for (int r = 0; r < 8; ++r) {
                            double acc = 0.0;
                            for (int k = 0; k < 8; ++k) acc += R[r * 8 + k] * in[k];
                            out[r] = acc;
                        }
                        for (int u = 0; u < 4; ++u) {
                            xs[(size_t)(t0 + u) * NS + j].x = (T)out[2 * u];
                            xs[(size_t)(t0 + u) * NS + j].y = (T)out[2 * u + 1];
                        }
                    }
            }
            for (uint32_t tid = 0; tid < (uint32_t)NT; ++tid) {
                A(&x)[NS] = *reinterpret_cast<A(*)[NS]>(&xs[(size_t)tid * NS]);
                pass_compute<T, RB>(x, P, ps, tid, outside, spool.data());
                if (last && xc.out_mine) {
                    const uint64_t gt = base ^ thread_offset64<NTB>(P.gl_store.thr, P.gl_store.c, tid);
                    for (int j = 0; j < NS; ++j) {
                        const XchgTarget t = xchg_target(slot_offset<RB>(gt, P.gl_store.reg, j), xc.bit_mask, xc.keep, xc.stash_mask);
                        (t.stays ? xc.out_mine : xc.out_peer)[t.base] = x[j];
                    }
                } else if (last) {
                    const uint64_t gt = base ^ thread_offset64<NTB>(P.gl_store.thr, P.gl_store.c, tid);
                    for (int j = 0; j < NS; ++j) psi[slot_offset<RB>(gt, P.gl_store.reg, j)] = x[j];
                } else {
                    const uint32_t st = thread_offset<NTB>(ps.st_thr, ps.st_c, tid);
                    uint32_t sr[RB_MAX] = {0, 0, 0, 0};
                    for (int b = 0; b < RB; ++b) sr[b] = ps.st_reg[b];
                    for (int j = 0; j < NS; ++j) smem[slot_offset<RB>(st, sr, j)] = x[j];
                }
            }
        }
    }
}

template <typename T>
void run_all(std::vector<typename Cx<T>::type> &psi, int n, int dtype, int rb, int L, bool dag, const std::vector<LoweredGate> &merged,
             int64_t *stats) {
    const std::vector<SweepPlan> plan = plan_sweeps_regs(n, merged, L, dag, 48, 512, dtype);
    std::vector<const LoweredGate *> cur;
    static RegProgram P;
    for (const SweepPlan &sw : plan) {
        stats[0] += 1;
        if (!sw.fused) {
            apply_lowered_host(psi, n, merged[sw.gates[0]]);
            stats[4] += 1;
            continue;
        }
        cur.clear();
        for (int i : sw.gates) cur.push_back(&merged[i]);
        build_reg_program(n, dtype, 0, cur, sw.need, L, rb, P);
        stats[1] += P.n_passes;
        stats[2] += P.n_folded;  // gates folded into pass boundaries
        for (int p = 0; p < P.n_passes; ++p) stats[3] += P.passes[p].udiag_end - P.passes[p].gate_begin;
        stats[5] += P.n_mma_gates;
        if (rb == 4)
            emulate_program<T, 4>(psi, n, P);
        else
            emulate_program<T, 3>(psi, n, P);
    }
}

}  // namespace

// host-side cost of one fused application (lowering, merging, sweep planning, program construction), in seconds
extern "C" double regs_emu_time_host_side(const void *ops_handle, int n, int dtype, int reps) {
    const qsv_ops *ops = reinterpret_cast<const qsv_ops *>(ops_handle);
    static RegProgram P;
    const auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        std::vector<LoweredGate> gates;
        for (const auto &op : ops->ops) {
            if (op.name == "Identity") continue;
            if (find_gate(op.name) != nullptr)
                gates.push_back(lower_named(n, op.name, op.wires, op.params, op.inverse));
            else
                gates.push_back(lower_matrix(n, op.matrix.data(), {}, op.wires, op.inverse));
        }
        const std::vector<LoweredGate> merged = prepare_gates_regs(gates);
        const std::vector<SweepPlan> plan = plan_sweeps_regs(n, merged, 4, true, 48, 512);
        std::vector<const LoweredGate *> cur;
        for (const SweepPlan &sw : plan) {
            if (!sw.fused) continue;
            cur.clear();
            for (int i : sw.gates) cur.push_back(&merged[i]);
            build_reg_program(n, dtype, 0, cur, sw.need, 4, 4, P);
        }
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() / reps;
}

// Two ranks of a register sharded on one global bit run the same batch of local gates and then exchange that global bit
// with local bit `local_bit` the way dist_apply_ops does with QSV_DIST_FUSED_SWAP=1: the last sweep of the batch stores out
// of place through xchg_target (carried = 1), or, when the exchanged bit is one of its tile bits or the batch ends in a lone
// gate, the batch runs in place and a copy pass does the same stores (k_xchg_oop; carried = 0).  shard / out: 2^n_local
// interleaved (re, im) doubles per rank.
extern "C" int regs_emu_fused_exchange(const void *ops_handle, int n_local, int low_bits, int local_bit, const double *shard0,
                                       const double *shard1, double *out0, double *out1, int *carried) {
    try {
        const qsv_ops *ops = reinterpret_cast<const qsv_ops *>(ops_handle);
        std::vector<LoweredGate> gates;
        for (const auto &op : ops->ops) {
            if (op.name == "Identity") continue;
            if (find_gate(op.name) != nullptr)
                gates.push_back(lower_named(n_local, op.name, op.wires, op.params, op.inverse));
            else
                gates.push_back(lower_matrix(n_local, op.matrix.data(), {}, op.wires, op.inverse));
        }
        const std::vector<LoweredGate> merged = prepare_gates_regs(gates);
        const int L = low_bits > 0 ? low_bits : 4;
        const uint64_t N = 1ull << n_local;
        const std::vector<SweepPlan> plan = plan_sweeps_regs(n_local, merged, L, true, 48, 512);
        std::vector<double2> out[2] = {std::vector<double2>(N), std::vector<double2>(N)};
        static RegProgram P;
        for (int r = 0; r < 2; ++r) {
            const double *in = r == 0 ? shard0 : shard1;
            std::vector<double2> psi(N);
            for (uint64_t i = 0; i < N; ++i) psi[i] = make_double2(in[2 * i], in[2 * i + 1]);
            EmuXchg<double2> xc;
            xc.out_mine = out[r].data();
            xc.out_peer = out[1 - r].data();
            xc.bit_mask = 1ull << local_bit;
            xc.keep = r ? xc.bit_mask : 0;
            bool done = false;
            std::vector<const LoweredGate *> cur;
            for (size_t k = 0; k < plan.size(); ++k) {
                const SweepPlan &sw = plan[k];
                const bool carry = k + 1 == plan.size() && (sw.fused || regs_fusable(merged[sw.gates[0]], n_local));
                if (!sw.fused && !carry) {
                    apply_lowered_host(psi, n_local, merged[sw.gates[0]]);
                    continue;
                }
                cur.clear();
                for (int i : sw.gates) cur.push_back(&merged[i]);
                build_reg_program(n_local, QSV_C128, 0, cur, sw.need, L, 4, P);
                emulate_program<double, 4>(psi, n_local, P, carry ? xc : EmuXchg<double2>());
                done = done || carry;
            }
            if (!done)  // the copy pass
                for (uint64_t i = 0; i < N; ++i) {
                    if ((i & xc.bit_mask) == xc.keep)
                        xc.out_mine[i] = psi[i];
                    else
                        xc.out_peer[i ^ xc.bit_mask] = psi[i];
                }
            if (carried) *carried = done ? 1 : 0;
        }
        for (uint64_t i = 0; i < N; ++i) {
            out0[2 * i] = out[0][i].x;
            out0[2 * i + 1] = out[0][i].y;
            out1[2 * i] = out[1][i].x;
            out1[2 * i + 1] = out[1][i].y;
        }
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "regs_emu_fused_exchange: %s\n", e.what());
        return 1;
    }
}

// A SPLIT exchange (csrc/dist.cu, QSV_DIST_SPLIT_XCHG): two ranks of a register sharded on one global bit run batch 1, whose
// last sweep pushes the leaving amplitudes with the stash bit clear to the partner and parks the others (xchg_target with a
// stash mask), then batch 2, whose first sweep fetches the parked amplitudes from the partner (xchg_source) -- or the copy
// passes when the batch cannot carry them (mode bit 0: push carried, bit 1: pull carried).  Buffer roles as in the
// library: batch 1 runs in X and stores into Y, the pull reads Y (own and the partner's) and writes X.
static void lower_all(const qsv_ops *ops, int n_local, std::vector<LoweredGate> &gates) {
    for (const auto &op : ops->ops) {
        if (op.name == "Identity") continue;
        if (find_gate(op.name) != nullptr)
            gates.push_back(lower_named(n_local, op.name, op.wires, op.params, op.inverse));
        else
            gates.push_back(lower_matrix(n_local, op.matrix.data(), {}, op.wires, op.inverse));
    }
}
extern "C" int regs_emu_split_exchange(const void *ops1_handle, const void *ops2_handle, int n_local, int local_bit,
                                       int stash_bit, const double *shard0, const double *shard1, double *out0, double *out1,
                                       int *mode) {
    try {
        std::vector<LoweredGate> g1, g2;
        lower_all(reinterpret_cast<const qsv_ops *>(ops1_handle), n_local, g1);
        lower_all(reinterpret_cast<const qsv_ops *>(ops2_handle), n_local, g2);
        const int L = 4;
        const uint64_t N = 1ull << n_local;
        const uint64_t bit_mask = 1ull << local_bit, stash_mask = 1ull << stash_bit;
        std::vector<double2> X[2] = {std::vector<double2>(N), std::vector<double2>(N)};
        std::vector<double2> Y[2] = {std::vector<double2>(N), std::vector<double2>(N)};
        for (uint64_t i = 0; i < N; ++i) {
            X[0][i] = make_double2(shard0[2 * i], shard0[2 * i + 1]);
            X[1][i] = make_double2(shard1[2 * i], shard1[2 * i + 1]);
        }
        static RegProgram P;
        int m = 0;
        // batch 1 on both ranks: X -> Y (own and partner's)
        {
            const std::vector<LoweredGate> merged = prepare_gates_regs(g1);
            const std::vector<SweepPlan> plan = plan_sweeps_regs(n_local, merged, L, true, 48, 512);
            for (int r = 0; r < 2; ++r) {
                EmuXchg<double2> xc;
                xc.out_mine = Y[r].data();
                xc.out_peer = Y[1 - r].data();
                xc.bit_mask = bit_mask;
                xc.keep = r ? bit_mask : 0;
                xc.stash_mask = stash_mask;
                bool done = false;
                std::vector<const LoweredGate *> cur;
                for (size_t k = 0; k < plan.size(); ++k) {
                    const SweepPlan &sw = plan[k];
                    const bool carry = k + 1 == plan.size() && (sw.fused || regs_fusable(merged[sw.gates[0]], n_local));
                    if (!sw.fused && !carry) {
                        apply_lowered_host(X[r], n_local, merged[sw.gates[0]]);
                        continue;
                    }
                    cur.clear();
                    for (int i : sw.gates) cur.push_back(&merged[i]);
                    build_reg_program(n_local, QSV_C128, 0, cur, sw.need, L, 4, P);
                    emulate_program<double, 4>(X[r], n_local, P, carry ? xc : EmuXchg<double2>());
                    done = done || carry;
                }
                if (!done)  // k_xchg_push_copy
                    for (uint64_t i = 0; i < N; ++i) {
                        const XchgTarget t = xchg_target(i, bit_mask, xc.keep, stash_mask);
                        (t.stays ? xc.out_mine : xc.out_peer)[t.base] = X[r][i];
                    }
                if (done) m |= 1;
            }
        }
        // batch 2: the first sweep (or the copy pass) reads Y of both ranks and writes X; the other sweeps run in place on X
        {
            const std::vector<LoweredGate> merged = prepare_gates_regs(g2);
            const std::vector<SweepPlan> plan = plan_sweeps_regs(n_local, merged, L, true, 48, 512);
            for (int r = 0; r < 2; ++r) {
                EmuXchg<double2> xc;
                xc.out_mine = X[r].data();
                xc.out_peer = nullptr;
                xc.in_peer = Y[1 - r].data();
                xc.pull_bit_mask = bit_mask;
                xc.pull_keep = r ? bit_mask : 0;
                xc.pull_stash_mask = stash_mask;
                const bool first_carries =
                    !plan.empty() && (plan[0].fused || regs_fusable(merged[plan[0].gates[0]], n_local));
                if (!first_carries)  // k_xchg_pull_copy
                    for (uint64_t i = 0; i < N; ++i) {
                        const XchgTarget t = xchg_source(i, bit_mask, xc.pull_keep, stash_mask);
                        X[r][i] = t.stays ? Y[r][t.base] : Y[1 - r][t.base];
                    }
                else
                    m |= 2;
                std::vector<const LoweredGate *> cur;
                for (size_t k = 0; k < plan.size(); ++k) {
                    const SweepPlan &sw = plan[k];
                    const bool pulls = first_carries && k == 0;
                    if (!sw.fused && !pulls) {
                        apply_lowered_host(X[r], n_local, merged[sw.gates[0]]);
                        continue;
                    }
                    cur.clear();
                    for (int i : sw.gates) cur.push_back(&merged[i]);
                    build_reg_program(n_local, QSV_C128, 0, cur, sw.need, L, 4, P);
                    if (pulls)
                        emulate_program<double, 4>(Y[r], n_local, P, xc);  // reads Y (and the partner's), stores into X
                    else
                        emulate_program<double, 4>(X[r], n_local, P);
                }
            }
        }
        for (uint64_t i = 0; i < N; ++i) {
            out0[2 * i] = X[0][i].x;
            out0[2 * i + 1] = X[0][i].y;
            out1[2 * i] = X[1][i].x;
            out1[2 * i + 1] = X[1][i].y;
        }
        if (mode) *mode = m;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "regs_emu_split_exchange: %s\n", e.what());
        return 1;
    }
}

// The sharded executor (csrc/dist.cu: dist_apply_ops) replayed on host shards for all 2^(n_total - n_local) ranks: the
// library's own lowering, exchange schedule (QSV_DIST_DAG) and per-rank gate localisation (controls / diagonal bits on
// global qubits resolved against the rank's index bits); only the gate arithmetic (plain reference loops) and the exchange
// (a host permutation between the two shards of a pair) are the emulator's.  state: 2^n_total interleaved (re, im)
// doubles in the canonical layout, updated in place; map_io: logical -> physical bit, in (nullptr = identity) and out;
// n_exchanges: exchanges performed.  `reps` applications in a row share the qubit map, as in bench.py.
extern "C" int dist_emu_apply_ops(const void *ops_handle, int n_total, int n_local, int reps, double *state, int *n_exchanges) {
    try {
        const qsv_ops *ops = reinterpret_cast<const qsv_ops *>(ops_handle);
        const int g = n_total - n_local, world = 1 << g;
        const uint64_t NL = 1ull << n_local;
        std::vector<std::vector<double2>> shard(world, std::vector<double2>(NL));
        for (int r = 0; r < world; ++r)
            for (uint64_t i = 0; i < NL; ++i) {
                const uint64_t k = ((uint64_t)r << n_local) | i;
                shard[r][i] = make_double2(state[2 * k], state[2 * k + 1]);
            }
        std::vector<LoweredGate> lowered;
        for (const auto &op : ops->ops)
            if (op.name != "Identity") lowered.push_back(dist_hook_lower(n_total, op));
        std::vector<int> phys_of(n_total), log_of(n_total);
        for (int b = 0; b < n_total; ++b) phys_of[b] = log_of[b] = b;
        int exchanges = 0;
        for (int rep = 0; rep < reps; ++rep) {
            std::vector<int> plan_phys = phys_of, plan_log = log_of;
            const auto steps = dist_hook_plan(lowered, plan_phys, plan_log, n_local);
            for (const auto &st : steps) {
                if (st[0] == 0) {
                    // exchange physical global bit st[1] with local bit st[2]: rank r keeps the amplitudes whose local
                    // bit equals its value a of the global bit, and swaps the others with rank r ^ (1 << gb)
                    const int gb = st[1] - n_local, l = st[2];
                    for (int r = 0; r < world; ++r) {
                        const int peer = r ^ (1 << gb);
                        if (peer < r) continue;
                        // r has a = 0: its elements with bit l = 1 <-> peer's (a = 1) elements with bit l = 0
                        for (uint64_t i = 0; i < NL; ++i)
                            if (!(i >> l & 1)) std::swap(shard[r][i | (1ull << l)], shard[peer][i]);
                    }
                    const int a = log_of[st[1]], b = log_of[st[2]];
                    log_of[st[1]] = b;
                    log_of[st[2]] = a;
                    phys_of[a] = st[2];
                    phys_of[b] = st[1];
                    ++exchanges;
                    continue;
                }
                for (int r = 0; r < world; ++r) {
                    const LoweredGate gl = dist_hook_localized(lowered[st[1]], phys_of, n_local, (uint64_t)r << n_local);
                    if (gl.kind != LoweredGate::NOP) apply_lowered_host(shard[r], n_local, gl);
                }
            }
            if (phys_of != plan_phys) return 2;
        }
        // back to the canonical layout: amplitude with logical index k sits at physical index p(k)
        const uint64_t N = 1ull << n_total;
        for (uint64_t k = 0; k < N; ++k) {
            uint64_t p = 0;
            for (int b = 0; b < n_total; ++b)
                if (k >> b & 1) p |= 1ull << phys_of[b];
            const double2 v = shard[p >> n_local][p & (NL - 1)];
            state[2 * k] = v.x;
            state[2 * k + 1] = v.y;
        }
        if (n_exchanges) *n_exchanges = exchanges;
        return 0;
    } catch (const std::exception &e) {
        std::fprintf(stderr, "dist_emu_apply_ops: %s\n", e.what());
        return 1;
    }
}

// Per-sweep structure of the programs the host side builds for a circuit (no state needed): for sweep k,
// out[8 k ..] = {passes, gates in the program, D2 blocks, D1-type gates, diagonal gates, tensor-core passes, gates folded
// into pass boundaries, 256 * (distinct dense-target bits of the sweep) + gates of the sweep before normalisation}.  Returns the number of sweeps (lone gates: passes = 0).
extern "C" int regs_emu_sweep_stats(const void *ops_handle, int n, int dtype, int rb, int low_bits, int dag, int64_t *out,
                                    int max_sweeps) {
    const qsv_ops *ops = reinterpret_cast<const qsv_ops *>(ops_handle);
    static RegProgram P;
    std::vector<LoweredGate> gates;
    for (const auto &op : ops->ops) {
        if (op.name == "Identity") continue;
        if (find_gate(op.name) != nullptr)
            gates.push_back(lower_named(n, op.name, op.wires, op.params, op.inverse));
        else
            gates.push_back(lower_matrix(n, op.matrix.data(), {}, op.wires, op.inverse));
    }
    const std::vector<LoweredGate> merged = prepare_gates_regs(gates);
    const int L = low_bits > 0 ? low_bits : 4;
    auto env_or = [](const char *name, int dflt) {
        const char *v = std::getenv(name);
        return v ? std::atoi(v) : dflt;
    };
    const std::vector<SweepPlan> plan = plan_sweeps_regs(n, merged, L, dag != 0, std::min(48, env_or("QSV_REGS_MAX_GATES", 48)),
                                                         std::max(1, env_or("QSV_REGS_WINDOW", 512)), dtype);
    std::vector<const LoweredGate *> cur;
    int k = 0;
    for (const SweepPlan &sw : plan) {
        if (k >= max_sweeps) break;
        int64_t *o = out + 8 * k++;
        for (int i = 0; i < 8; ++i) o[i] = 0;
        uint64_t dense_bits = 0;
        for (int i : sw.gates)
            if (merged[i].kind == LoweredGate::DENSE)
                for (uint64_t off : merged[i].offs) dense_bits |= off;
        o[7] = 256 * (int64_t)__builtin_popcountll(dense_bits) + (int64_t)sw.gates.size();
        if (std::getenv("REGS_EMU_DUMP")) {
            std::printf("sweep %d:", k - 1);
            for (int i : sw.gates) {
                const LoweredGate &g = merged[i];
                uint64_t db = 0;
                if (g.kind == LoweredGate::DENSE)
                    for (uint64_t off : g.offs) db |= off;
                std::printf(" %s[", g.kind == LoweredGate::DENSE ? (g.k == 2 ? "D2" : "D1") : "diag");
                for (int b = 0; b < n; ++b)
                    if (db >> b & 1) std::printf("%d ", b);
                std::printf("|c");
                for (int b = 0; b < n; ++b)
                    if (g.ctrl_mask >> b & 1) std::printf(" %d", b);
                std::printf("]");
            }
            std::printf("\n");
        }
        if (!sw.fused) continue;
        cur.clear();
        for (int i : sw.gates) cur.push_back(&merged[i]);
        build_reg_program(n, dtype, 0, cur, sw.need, L, rb, P);
        o[0] = P.n_passes;
        o[1] = P.n_gates;
        for (int g = 0; g < P.n_gates; ++g) {
            const int kind = P.gates[g].kind;
            if (kind == RG_D2)
                ++o[2];
            else if (kind == RG_DIAG || kind == RG_D1_DIAG)
                ++o[4];
            else
                ++o[3];
        }
        for (int p = 0; p < P.n_passes; ++p) o[5] += P.passes[p].mma_off != NO_MMA ? 1 : 0;
        o[6] = P.n_folded;
    }
    return k;
}

// ops = a qsv_ops handle of libqsv_b200.so; state = 2^n interleaved (re, im) doubles, updated in place (complex64 runs
// the float kernel code on a float copy).  stats[6] = {sweeps, passes, folded permutation gates, merged diagonal gates,
// lone gates, tensor-core gates}.  Returns 0 on success.
extern "C" int regs_emu_apply_ops(const void *ops_handle, int n, int dtype, int rb, int low_bits, int dag, double *state,
                                  int64_t *stats) {
    try {
        const qsv_ops *ops = reinterpret_cast<const qsv_ops *>(ops_handle);
        std::vector<LoweredGate> gates;
        for (const auto &op : ops->ops) {
            if (op.name == "Identity") continue;
            if (find_gate(op.name) != nullptr)
                gates.push_back(lower_named(n, op.name, op.wires, op.params, op.inverse));
            else
                gates.push_back(lower_matrix(n, op.matrix.data(), {}, op.wires, op.inverse));
        }
        const std::vector<LoweredGate> merged = prepare_gates_regs(gates);
        for (int i = 0; i < 6; ++i) stats[i] = 0;
        const uint64_t N = 1ull << n;
        const int L = low_bits > 0 ? low_bits : 4;
        if (dtype == QSV_C128) {
            std::vector<double2> psi(N);
            for (uint64_t i = 0; i < N; ++i) psi[i] = make_double2(state[2 * i], state[2 * i + 1]);
            run_all<double>(psi, n, dtype, rb, L, dag != 0, merged, stats);
            for (uint64_t i = 0; i < N; ++i) {
                state[2 * i] = psi[i].x;
                state[2 * i + 1] = psi[i].y;
            }
        } else {
            std::vector<float2> psi(N);
            for (uint64_t i = 0; i < N; ++i) psi[i] = make_float2((float)state[2 * i], (float)state[2 * i + 1]);
            run_all<float>(psi, n, dtype, rb, L, dag != 0, merged, stats);
            for (uint64_t i = 0; i < N; ++i) {
                state[2 * i] = psi[i].x;
                state[2 * i + 1] = psi[i].y;
            }
        }
        return 0;
    } catch (const std::exception &e) {
        fprintf(stderr, "regs_emu: %s\n", e.what());
        return 1;
    }
}
