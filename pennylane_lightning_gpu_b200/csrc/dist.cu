// Sharded state vectors: one process per GPU, amplitude index bits [n_local, n_total) = rank.
//
// Replaces, in the reference, StateVectorCudaMPI's bit-swap path
//   simulator/StateVectorCudaMPI.hpp:2023-2088 (global-wire detection), :2488-2587 (applyMPI_Dispatcher:
//   scheduler + SVSwapWorker, swap in, apply, swap back, barrier + device sync around every batch),
//   simulator/MPIWorker.hpp:55-91 (createWirePairs), :226-349 (communicator, transfer workspace, IPC)
// and util/MPIManager.hpp's collectives, with
//   * NCCL send/recv over NVLink for the exchange, chunked through two staging buffers; the D2D
//     placement of chunk i overlaps the transfer of chunk i+1 on a second stream;
//   * a LAZY logical->physical qubit map: a swapped-in qubit stays local until evicted (the
//     reference swaps back after every gate), victims chosen by farthest next use (the whole
//     circuit is known in qsv_dist_apply_ops);
//   * ZERO communication for controls and diagonal gates on global qubits: they are resolved against
//     the rank's own index bits on the host (the reference swaps for those too, MPI.hpp:2054-2087);
//   * stream-ordered NCCL, no barriers, no device-wide syncs.
#include <nccl.h>

#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <mutex>

#include "qsv_internal.h"

namespace qsv {

#define QSV_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (expr);                                                                  \
        if (r__ != ncclSuccess)                                                                     \
            ::qsv::fail(std::string("NCCL error: ") + ncclGetErrorString(r__) + " in " #expr);      \
    } while (0)

// a device buffer of shard size together with its CUDA-IPC mappings on every other rank
struct PeerBuf {
    void *local = nullptr;
    std::vector<void *> peer;      // peer[r] = rank r's buffer mapped here (local for r == rank)
    std::vector<void *> map_base;  // what cudaIpcOpenMemHandle returned (for closing)
    bool ok = false;
};

struct DistCtx {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, n_total = 0, n_global = 0;
    std::vector<int> phys_of;  // logical bit -> physical bit
    std::vector<int> log_of;   // physical bit -> logical bit
    cudaStream_t comm_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_xfer[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    // exchange timing for the statistics: event pairs recorded on the stream and read LAZILY (qsv_dist_*_swap_stats) --
    // the host never waits for an exchange on the hot path, so the next batch is queued behind it at once
    struct TimedSwap {
        cudaEvent_t t0, t1;
        uint64_t bytes;
    };
    std::vector<TimedSwap> pending;          // recorded, not read yet
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> free_events;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;  // the pair of the exchange in flight (taken from free_events)
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    uint64_t last_bytes = 0, total_bytes = 0;
    float last_ms = 0.f, total_ms = 0.f;
    int n_swaps = 0;
    double *red_dev = nullptr;
    // direct peer access (CUDA IPC): the partners' shards mapped into this process
    bool p2p = false;
    PeerBuf main;                          // the register's own shard (sv.data at registration time)
    int *hs_dev = nullptr;                 // 2 ints for the handshakes
    // companions: further vectors of shard size that share the qubit map and therefore follow every exchange
    // (lambda and the bras of the adjoint method, temporaries of observables)
    struct Companion {
        void *data;
        PeerBuf pb;
    };
    std::vector<Companion> comp;
    bool skip_main = false;                // adjoint sweep: the register itself stays where it is
    // Qubit-map policy.  lazy_map = false (default, the reference's contract): every collective entry point leaves the
    // register in the canonical layout, so DeviceToHost / `state` stay purely local reads of the shard, as in the
    // reference (StateVectorCudaBase.hpp:104-228), and a rank-conditional read cannot dead-lock.  lazy_map = true
    // (qsv_dist_set_lazy_map; the throughput setting of bench.py): the map persists between calls -- a swapped-in
    // qubit stays local until evicted -- and qsv_dist_d2h / qsv_dist_canonicalize become COLLECTIVE.
    bool lazy_map = false;
    // exchanges carried by sweeps (default; QSV_DIST_FUSED_SWAP=0 switches it off): a second shard-sized buffer, peer-mapped like the first;
    // a fused sweep reads the buffer the register lives in and writes this rank's and the partner's other buffer
    void *shadow = nullptr;
    PeerBuf shadow_pb;
    bool shadow_tried = false;
    int n_fused = 0, n_oop = 0;            // statistics: exchanges through the second buffer, and how many of them a sweep carried
    int n_pull = 0, n_pull_carried = 0;    // split exchanges: second halves, and how many of them a sweep carried
    void **tab_dev = nullptr;              // device table of vector pointers for batched launches
    std::vector<void *> tab_cache;
};

void setup_peer_access(State &sv);
PeerBuf register_peer_buffer(State &sv, void *data);
void release_peer_buffer(PeerBuf &pb);

void setup_peer_access(State &sv);

namespace {

uint64_t touched_mask(const LoweredGate &g) {
    uint64_t m = 0;
    for (uint64_t o : g.offs) m |= o;
    return m;
}

uint64_t remap_mask(uint64_t m, const std::vector<int> &phys_of) {
    uint64_t r = 0;
    for (int b = 0; b < (int)phys_of.size(); ++b)
        if (m >> b & 1) r |= 1ull << phys_of[b];
    return r;
}

// logical-bit gate -> physical-bit gate under the current qubit map
LoweredGate remap_gate(const LoweredGate &g, const std::vector<int> &phys_of) {
    LoweredGate r = g;
    r.ctrl_mask = remap_mask(g.ctrl_mask, phys_of);
    r.zmask = remap_mask(g.zmask, phys_of);
    for (auto &o : r.offs) o = remap_mask(o, phys_of);
    for (auto &b : r.tgt_bits) b = phys_of[b];
    for (auto &h : r.holes) h = phys_of[h];
    std::sort(r.holes.begin(), r.holes.end());
    return r;
}

// resolve everything that refers to physical bits >= n_local against this rank's index bits
LoweredGate localize_gate(const LoweredGate &g, int n_local, uint64_t index_hi) {
    const uint64_t hi_mask = ~((1ull << n_local) - 1ull);
    LoweredGate r = g;
    const uint64_t gc = g.ctrl_mask & hi_mask;
    if ((index_hi & gc) != gc) return LoweredGate{};  // a global control is 0 on this rank
    r.ctrl_mask = g.ctrl_mask & ~hi_mask;
    if (g.kind == LoweredGate::DENSE) {
        QSV_CHECK((touched_mask(g) & hi_mask) == 0, "internal: dense target on a global qubit");
        r.holes.clear();
        for (int h : g.holes)
            if (h < n_local) r.holes.push_back(h);
    } else if (g.kind == LoweredGate::DIAG) {
        std::vector<int> keep;
        int fixed = 0;
        for (int b = 0; b < g.k; ++b) {
            const int bit = g.tgt_bits[b];
            if (bit >= n_local) {
                if (index_hi >> bit & 1) fixed |= 1 << (g.k - 1 - b);
            } else {
                keep.push_back(b);
            }
        }
        if ((int)keep.size() != g.k) {
            const int k2 = (int)keep.size();
            std::vector<cplx> tab(1u << k2);
            for (int t2 = 0; t2 < (1 << k2); ++t2) {
                int t = fixed;
                for (int j = 0; j < k2; ++j)
                    if (t2 >> (k2 - 1 - j) & 1) t |= 1 << (g.k - 1 - keep[j]);
                tab[t2] = g.mat[t];
            }
            std::vector<int> tb;
            for (int j : keep) tb.push_back(g.tgt_bits[j]);
            r.k = k2;
            r.tgt_bits = tb;
            r.mat = tab;
        }
    } else if (g.kind == LoweredGate::PARITY) {
        r.zmask = g.zmask & ~hi_mask;
        if (__builtin_popcountll(index_hi & g.zmask & hi_mask) & 1) std::swap(r.mat[0], r.mat[1]);
    }
    return r;
}

LoweredGate lower_op_total(int n_total, const Op &op, bool extra_adjoint) {
    const bool adj = op.inverse != extra_adjoint;
    if (find_gate(op.name) != nullptr) return lower_named(n_total, op.name, op.wires, op.params, adj);
    QSV_CHECK(!op.matrix.empty(), "Currently unsupported gate: " + op.name);
    const size_t dim = 1ull << op.wires.size();
    QSV_CHECK(op.matrix.size() == dim * dim, "matrix of gate " + op.name + " does not match its wires");
    return lower_matrix(n_total, op.matrix.data(), {}, op.wires, adj);
}

// Eviction policy shared by the executor and the host-only planner: among the local physical bits
// whose logical qubit gate i does not need, evict the one needed again latest (Belady); look at the top
// 8 bits of the shard first so that the exchanged half consists of few, large contiguous blocks.
int pick_victim(const std::vector<uint64_t> &need, size_t i, const std::vector<int> &log_of, int n_local) {
    const int window_lo = std::max(0, n_local - 8);
    int best = -1;
    size_t best_next = 0;
    for (int pass = 0; pass < 2 && best < 0; ++pass) {
        const int lo = pass == 0 ? window_lo : 0;
        for (int l = n_local - 1; l >= lo; --l) {
            const int q = log_of[l];
            if (need[i] >> q & 1) continue;
            size_t next = need.size() + 1;
            for (size_t j = i + 1; j < need.size(); ++j)
                if (need[j] >> q & 1) {
                    next = j;
                    break;
                }
            if (best < 0 || next > best_next) {
                best = l;
                best_next = next;
            }
        }
    }
    QSV_CHECK(best >= 0, "no local qubit can be evicted");
    return best;
}

// ------------------------------------------------------------------------------------------------
// Exchange schedule of a circuit on the sharded register, shared by the executor (dist_apply_ops) and the
// host-only planner (qsv_dist_plan).
//   dense[i] = logical bits gate i changes (they have to be local when it runs); diag[i] = logical bits it only
//   looks at (controls, phase-table / parity bits): fine on global qubits.
// Program order (QSV_DIST_DAG=0): a global qubit is swapped in when the next gate needs it, victim = farthest next use.
// Dependency order (default): two gates commute when they share no bit, or only bits on which both are diagonal.  Gates
// are executed as long as any ready gate has all its dense bits local; only when every ready gate waits for a global
// qubit is one exchanged, and the (incoming qubit, evicted local bit) pair is the one that lets most gates run before
// the next stall (bounded look-ahead), ties broken by the farthest next use of the evicted qubit.  On the 200-gate
// random circuits of the bench this needs 1 / 2 / 4 exchanges at 2 / 4 / 8 GPUs instead of 3 / 6 / 6.
// ------------------------------------------------------------------------------------------------
struct DistStep {
    int kind;  // 0 = exchange global physical bit a with local physical bit b, 1 = gate a
    int a, b;
};

bool dist_dag_enabled() {
    const char *e = std::getenv("QSV_DIST_DAG");
    return !(e && std::atoi(e) == 0);
}

std::vector<DistStep> plan_dist_steps(const std::vector<uint64_t> &dense, const std::vector<uint64_t> &diag,
                                      std::vector<int> &phys_of, std::vector<int> &log_of, int n_local, int depth_arg = 0,
                                      int window_arg = 0) {
    const int n_total = (int)phys_of.size();
    const size_t n = dense.size();
    std::vector<DistStep> steps;
    auto do_swap = [&](int gphys, int l) {
        steps.push_back({0, gphys, l});
        const int a = log_of[gphys], b = log_of[l];
        log_of[gphys] = b;
        log_of[l] = a;
        phys_of[a] = l;
        phys_of[b] = gphys;
    };
    if (!dist_dag_enabled()) {
        for (size_t i = 0; i < n; ++i) {
            for (int lb = 0; lb < n_total; ++lb) {
                if (!(dense[i] >> lb & 1) || phys_of[lb] < n_local) continue;
                do_swap(phys_of[lb], pick_victim(dense, i, log_of, n_local));
            }
            steps.push_back({1, (int)i, 0});
        }
        return steps;
    }
    // dependency graph by per-bit last writer (dense) and the readers (diagonal) since then
    std::vector<std::vector<int>> succ(n);
    std::vector<int> indeg(n, 0);
    {
        std::vector<int> writer(n_total, -1);
        std::vector<std::vector<int>> readers(n_total);
        std::vector<int> preds;
        for (size_t i = 0; i < n; ++i) {
            preds.clear();
            for (int b = 0; b < n_total; ++b) {
                if (dense[i] >> b & 1) {
                    if (writer[b] >= 0) preds.push_back(writer[b]);
                    preds.insert(preds.end(), readers[b].begin(), readers[b].end());
                    writer[b] = (int)i;
                    readers[b].clear();
                } else if (diag[i] >> b & 1) {
                    if (writer[b] >= 0) preds.push_back(writer[b]);
                    readers[b].push_back((int)i);
                }
            }
            std::sort(preds.begin(), preds.end());
            preds.erase(std::unique(preds.begin(), preds.end()), preds.end());
            for (int p : preds) succ[p].push_back((int)i);
            indeg[i] = (int)preds.size();
        }
    }
    std::vector<char> done(n, 0);
    std::vector<int> ready;  // ready and not yet executed, kept sorted (program order among ready gates)
    for (size_t i = 0; i < n; ++i)
        if (indeg[i] == 0) ready.push_back((int)i);
    size_t n_done = 0;  // includes the gates of a simulation in progress
    // run every ready gate whose dense bits avoid `gmask` (the global qubits); returns the number executed.  With
    // record != nullptr the run is a simulation: executed gates are listed there and undone by the caller.
    auto run = [&](uint64_t gmask, std::vector<int> *record, size_t cap) {
        size_t count = 0;
        for (size_t r = 0; r < ready.size() && count < cap;) {
            const int i = ready[r];
            if (dense[i] & gmask) {
                ++r;
                continue;
            }
            ready.erase(ready.begin() + r);
            done[i] = 1;
            ++count;
            ++n_done;
            if (record)
                record->push_back(i);
            else
                steps.push_back({1, i, 0});
            for (int sidx : succ[i])
                if (--indeg[sidx] == 0) ready.insert(std::lower_bound(ready.begin(), ready.end(), sidx), sidx);
            r = 0;  // a successor may have been inserted before r (ready lists are short)
        }
        return count;
    };
    auto undo = [&](const std::vector<int> &record, const std::vector<int> &ready_before) {
        for (auto it = record.rbegin(); it != record.rend(); ++it) {
            done[*it] = 0;
            --n_done;
            for (int sidx : succ[*it]) ++indeg[sidx];
        }
        ready = ready_before;
    };
    const size_t LOOKAHEAD = 2048;
    const long FINISHED = 1l << 40;  // score of an exchange after which the circuit runs to its end
    size_t first_pending = 0;
    auto next_use = [&](int q) {  // program-order position of the next pending gate that changes qubit q
        size_t scanned = 0;
        for (size_t j = first_pending; j < n && scanned < 4 * LOOKAHEAD; ++j) {
            if (done[j]) continue;
            ++scanned;
            if (dense[j] >> q & 1) return j;
        }
        return n + 1;
    };
    struct Choice {
        long score = -1;
        size_t next = 0;
        int q = -1, l = -1;
    };
    // evict from the top bits: few, large contiguous blocks
    const int window_lo = std::max(0, n_local - (window_arg > 0 ? window_arg : 8));
    // best (incoming qubit, evicted local bit) at a stall under the map (phys, log); depth 2 adds the best score of
    // the following stall
    std::function<Choice(const std::vector<int> &, const std::vector<int> &, int)> choose =
        [&](const std::vector<int> &phys, const std::vector<int> &log, int depth) {
            uint64_t gmask = 0;
            for (int q = 0; q < n_total; ++q)
                if (phys[q] >= n_local) gmask |= 1ull << q;
            uint64_t want = 0;  // global qubits the ready gates are waiting for
            for (int i : ready) want |= dense[i] & gmask;
            Choice best;
            std::vector<int> record;
            const std::vector<int> ready_before = ready;
            for (int q = 0; q < n_total; ++q) {
                if (!(want >> q & 1)) continue;
                for (int l = n_local - 1; l >= window_lo; --l) {
                    const int out = log[l];
                    const uint64_t trial = (gmask & ~(1ull << q)) | (1ull << out);
                    record.clear();
                    const size_t count = run(trial, &record, LOOKAHEAD);
                    long score = (long)count;
                    if (n_done == n) {
                        score += FINISHED;
                    } else if (depth > 1 && count > 0 && count < LOOKAHEAD) {
                        std::vector<int> phys2 = phys, log2 = log;
                        const int gp = phys2[q];
                        log2[gp] = out;
                        log2[l] = q;
                        phys2[q] = l;
                        phys2[out] = gp;
                        score += std::max(0l, choose(phys2, log2, depth - 1).score);
                    }
                    undo(record, ready_before);
                    const size_t nu = next_use(out);
                    if (score > best.score || (score == best.score && nu > best.next)) {
                        best.score = score;
                        best.next = nu;
                        best.q = q;
                        best.l = l;
                    }
                }
            }
            return best;
        };
    const char *depth_env = std::getenv("QSV_DIST_PLAN_DEPTH");
    const int depth = depth_arg > 0 ? depth_arg
                                    : (depth_env ? std::max(1, std::min(3, std::atoi(depth_env))) : (n <= 4096 ? 2 : 1));
    while (true) {
        uint64_t gmask = 0;
        for (int q = 0; q < n_total; ++q)
            if (phys_of[q] >= n_local) gmask |= 1ull << q;
        run(gmask, nullptr, n + 1);
        if (n_done == n) break;
        QSV_CHECK(!ready.empty(), "internal: the exchange planner found no ready gate");
        while (first_pending < n && done[first_pending]) ++first_pending;
        Choice best = choose(phys_of, log_of, depth);
        if (best.score <= 0) {
            // the ready gates wait for more than one global qubit each: serve the first of them, never evicting
            // a qubit it needs itself (so its number of global qubits falls with every exchange)
            const int i0 = ready.front();
            best = Choice{};
            for (int q = 0; q < n_total && best.q < 0; ++q)
                if ((dense[i0] & gmask) >> q & 1) best.q = q;
            for (int pass = 0; pass < 2 && best.l < 0; ++pass)
                for (int l = n_local - 1; l >= (pass == 0 ? window_lo : 0); --l) {
                    const int out = log_of[l];
                    if (dense[i0] >> out & 1) continue;
                    const size_t nu = next_use(out);
                    if (best.l < 0 || nu > best.next) {
                        best.l = l;
                        best.next = nu;
                    }
                }
            QSV_CHECK(best.q >= 0 && best.l >= 0, "no local qubit can be evicted");
        }
        do_swap(phys_of[best.q], best.l);
    }
    return steps;
}

// logical bits a lowered gate changes / only looks at
void gate_bit_masks(const LoweredGate &g, uint64_t &dense, uint64_t &diag) {
    dense = g.kind == LoweredGate::DENSE ? touched_mask(g) : 0;
    uint64_t all = g.ctrl_mask | g.zmask;
    for (int h : g.holes) all |= 1ull << h;
    for (int b : g.tgt_bits) all |= 1ull << b;
    diag = all & ~dense;
}

void ensure_stage(State &sv, size_t bytes) {
    DistCtx &d = *sv.dist;
    if (d.stage_bytes >= bytes) return;
    QSV_CUDA(cudaDeviceSynchronize());
    for (int k = 0; k < 2; ++k) {
        if (d.stage[k]) QSV_CUDA(cudaFree(d.stage[k]));
        d.stage[k] = nullptr;
        QSV_CUDA(cudaMalloc(&d.stage[k], bytes));
    }
    d.stage_bytes = bytes;
}

// In-place exchange with the partner GPU through its peer-mapped shard (NVLink load/store, no staging):
//     mine[idx with bit l = !a]  <->  partner[idx with bit l = a]        (a = this rank's value of the global bit)
// Each GPU of the pair runs this kernel on one half of the index range, so both NVLink directions carry
// half of the traffic as remote loads of one kernel and half as remote stores of the other.
template <int U>
__global__ void __launch_bounds__(256)
    k_peer_swap(uint4 *__restrict__ mine, uint4 *__restrict__ theirs, uint64_t first, uint64_t count, int l_vec,
                uint64_t my_bit_vec, uint64_t their_bit_vec) {
    // indices are in 16-byte units; l_vec is the position of the swapped bit in those units
    const uint64_t stride = (uint64_t)gridDim.x * 256 * U;
    const uint64_t low = (1ull << l_vec) - 1ull;
    for (uint64_t i0 = (uint64_t)blockIdx.x * 256 * U + threadIdx.x; i0 < count; i0 += stride) {
        uint4 a[U], b[U];
        uint64_t im[U], it[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = first + i0 + (uint64_t)u * 256;
            const uint64_t e = ((i >> l_vec) << (l_vec + 1)) | (i & low);
            im[u] = e | my_bit_vec;
            it[u] = e | their_bit_vec;
            if (i0 + (uint64_t)u * 256 < count) {
                a[u] = mine[im[u]];
                b[u] = theirs[it[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i0 + (uint64_t)u * 256 < count) {
                mine[im[u]] = b[u];
                theirs[it[u]] = a[u];
            }
        }
    }
}

// (The out-of-place forms of an exchange -- carried by a sweep, or as copy passes -- live in tile_regs.cu:
// k_tile_regs<..., XCHG = true>, launch_xchg_push_copy, launch_xchg_pull_copy.)
void handshake(State &sv, int peer) {
    DistCtx &d = *sv.dist;
    QSV_NCCL(ncclGroupStart());
    QSV_NCCL(ncclSend(d.hs_dev, 1, ncclInt, peer, d.comm, sv.stream));
    QSV_NCCL(ncclRecv(d.hs_dev + 1, 1, ncclInt, peer, d.comm, sv.stream));
    QSV_NCCL(ncclGroupEnd());
}

// the vectors an exchange has to move: the register (unless frozen) and every companion
struct SwapItem {
    char *data;
    const PeerBuf *pb;
};
std::vector<SwapItem> swap_set(State &sv) {
    DistCtx &d = *sv.dist;
    std::vector<SwapItem> v;
    if (!d.skip_main) v.push_back({(char *)sv.data, &d.main});
    for (auto &c : d.comp) v.push_back({(char *)c.data, &c.pb});
    return v;
}

// statistics without a host stall: take an event pair for the exchange about to be queued ...
void begin_timed_swap(DistCtx &d) {
    if (d.free_events.empty()) {
        cudaEvent_t a, b;
        QSV_CUDA(cudaEventCreate(&a));
        QSV_CUDA(cudaEventCreate(&b));
        d.free_events.push_back({a, b});
    }
    d.ev_t0 = d.free_events.back().first;
    d.ev_t1 = d.free_events.back().second;
    d.free_events.pop_back();
}
// ... read every finished (or, with `wait`, every recorded) pair into the totals
void harvest_swaps(DistCtx &d, bool wait) {
    size_t k = 0;
    for (; k < d.pending.size(); ++k) {
        DistCtx::TimedSwap &t = d.pending[k];
        if (wait) {
            QSV_CUDA(cudaEventSynchronize(t.t1));
        } else if (cudaEventQuery(t.t1) != cudaSuccess) {
            cudaGetLastError();
            break;
        }
        float ms = 0.f;
        QSV_CUDA(cudaEventElapsedTime(&ms, t.t0, t.t1));
        d.last_ms = ms;
        d.last_bytes = t.bytes;
        d.total_ms += ms;
        d.total_bytes += t.bytes;
        d.n_swaps += 1;
        d.free_events.push_back({t.t0, t.t1});
    }
    d.pending.erase(d.pending.begin(), d.pending.begin() + k);
}
void end_timed_swap(DistCtx &d, uint64_t bytes) {
    d.pending.push_back({d.ev_t0, d.ev_t1, bytes});
    d.ev_t0 = d.ev_t1 = nullptr;
    if (d.pending.size() > 64) harvest_swaps(d, d.pending.size() > 1024);
}

void swap_p2p(State &sv, const std::vector<SwapItem> &items, int gphys, int l) {
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    const int gb = gphys - n_local;
    const int peer = d.rank ^ (1 << gb);
    const uint64_t a = (d.rank >> gb) & 1;
    // 16-byte units: complex128 = 1 unit, complex64 = half a unit (two amplitudes per unit, so l >= 1)
    const int shift = sv.dtype == QSV_C128 ? 0 : 1;
    QSV_CHECK(l >= shift, "internal: cannot swap index bit 0 of a complex64 shard through 16-byte units");
    const int l_vec = l - shift;
    const uint64_t pairs = (sv.length() >> shift) >> 1;  // 16-byte units in the exchanged half
    const uint64_t half = pairs / 2;
    const uint64_t first = a == 0 ? 0 : half;
    const uint64_t count = a == 0 ? half : pairs - half;
    begin_timed_swap(d);
    QSV_CUDA(cudaEventRecord(d.ev_t0, sv.stream));
    handshake(sv, peer);  // the partner has finished everything queued before its own handshake
    if (count > 0) {
        constexpr int U = 4;
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((count + 256 * U - 1) / (256 * U), 148 * 16));
        for (const SwapItem &it : items) {
            k_peer_swap<U><<<grid, 256, 0, sv.stream>>>((uint4 *)it.data, (uint4 *)it.pb->peer[peer], first, count, l_vec,
                                                        (a ^ 1) << l_vec, a << l_vec);
            QSV_CUDA(cudaGetLastError());
        }
    }
    handshake(sv, peer);  // the partner's kernels have finished writing into this rank's vectors
    QSV_CUDA(cudaEventRecord(d.ev_t1, sv.stream));
}

}  // namespace


// ------------------------------------------------------------------------------------------------
// Hooks for the CPU emulation of the sharded executor (tests/native/regs_emu.cu): the host-side pieces of
// dist_apply_ops -- lowering on the whole register, the exchange schedule, and the gate a rank actually runs under a
// qubit map -- so that the code the GPUs execute is what the test replays on host shards.
// ------------------------------------------------------------------------------------------------
LoweredGate dist_hook_lower(int n_total, const Op &op) { return lower_op_total(n_total, op, false); }

std::vector<DistStep> plan_dist_steps_priced(const std::vector<LoweredGate> &lowered, const std::vector<uint64_t> &dense,
                                             const std::vector<uint64_t> &diag, std::vector<int> &phys_of,
                                             std::vector<int> &log_of, int n_local, int dtype, bool out_of_place);

std::vector<std::array<int, 3>> dist_hook_plan(const std::vector<LoweredGate> &lowered, std::vector<int> &phys_of,
                                               std::vector<int> &log_of, int n_local) {
    std::vector<uint64_t> dense(lowered.size(), 0), diag(lowered.size(), 0);
    for (size_t i = 0; i < lowered.size(); ++i) gate_bit_masks(lowered[i], dense[i], diag[i]);
    std::vector<std::array<int, 3>> out;
    // what dist_run_lowered executes for a fused complex128 circuit with room for the second buffer
    for (const DistStep &st : plan_dist_steps_priced(lowered, dense, diag, phys_of, log_of, n_local, QSV_C128, true))
        out.push_back({st.kind, st.a, st.b});
    return out;
}

LoweredGate dist_hook_localized(const LoweredGate &g, const std::vector<int> &phys_of, int n_local, uint64_t index_hi) {
    return localize_gate(remap_gate(g, phys_of), n_local, index_hi);
}

// physical swap of global bit gphys (>= n_local) with local bit l, applied to the register and its companions
void dist_swap_physical(State &sv, int gphys, int l, size_t chunk_bytes) {
    sv.use();
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    QSV_CHECK(gphys >= n_local && gphys < d.n_total && l >= 0 && l < n_local, "invalid swap bits");
    const int gb = gphys - n_local;
    const int peer = d.rank ^ (1 << gb);
    const int mybit = (d.rank >> gb) & 1;
    const size_t ab = sv.amp_bytes();
    const std::vector<SwapItem> items = swap_set(sv);
    if (items.empty()) return;
    bool all_mapped = d.p2p && !(sv.dtype == QSV_C64 && l == 0);
    for (const SwapItem &it : items) all_mapped = all_mapped && it.pb->ok;
    if (!d.skip_main)
        QSV_CHECK(!d.p2p || sv.data == d.main.local, "the shard was re-allocated after qsv_dist_init; peer mappings are stale");
    if (all_mapped) {
        swap_p2p(sv, items, gphys, l);
        end_timed_swap(d, (uint64_t)(sv.length() / 2) * ab * items.size());
        return;
    }
    const uint64_t block_amps = 1ull << l;
    const uint64_t n_blocks = 1ull << (n_local - 1 - l);
    if (chunk_bytes == 0) chunk_bytes = (size_t)256 << 20;
    const uint64_t chunk_amps = std::max<uint64_t>(1, std::min<uint64_t>(block_amps, chunk_bytes / ab));
    ensure_stage(sv, chunk_amps * ab);

    QSV_CUDA(cudaEventRecord(d.ev_ready, sv.stream));
    QSV_CUDA(cudaStreamWaitEvent(d.comm_stream, d.ev_ready, 0));
    QSV_CUDA(cudaStreamWaitEvent(d.copy_stream, d.ev_ready, 0));
    begin_timed_swap(d);
    QSV_CUDA(cudaEventRecord(d.ev_t0, d.comm_stream));
    uint64_t counter = 0;
    for (const SwapItem &it : items) {
        char *base = it.data;
        for (uint64_t blk = 0; blk < n_blocks; ++blk) {
            const uint64_t start = (blk << (l + 1)) | ((uint64_t)(mybit ^ 1) << l);
            for (uint64_t off = 0; off < block_amps; off += chunk_amps, ++counter) {
                const int k = (int)(counter & 1);
                const size_t bytes = (size_t)std::min<uint64_t>(chunk_amps, block_amps - off) * ab;
                char *ptr = base + (start + off) * ab;
                if (counter >= 2) QSV_CUDA(cudaStreamWaitEvent(d.comm_stream, d.ev_copy[k], 0));
                QSV_NCCL(ncclGroupStart());
                QSV_NCCL(ncclSend(ptr, bytes, ncclChar, peer, d.comm, d.comm_stream));
                QSV_NCCL(ncclRecv(d.stage[k], bytes, ncclChar, peer, d.comm, d.comm_stream));
                QSV_NCCL(ncclGroupEnd());
                QSV_CUDA(cudaEventRecord(d.ev_xfer[k], d.comm_stream));
                QSV_CUDA(cudaStreamWaitEvent(d.copy_stream, d.ev_xfer[k], 0));
                QSV_CUDA(cudaMemcpyAsync(ptr, d.stage[k], bytes, cudaMemcpyDeviceToDevice, d.copy_stream));
                QSV_CUDA(cudaEventRecord(d.ev_copy[k], d.copy_stream));
            }
        }
    }
    QSV_CUDA(cudaEventRecord(d.ev_t1, d.comm_stream));
    QSV_CUDA(cudaStreamWaitEvent(sv.stream, d.ev_t1, 0));
    for (int k = 0; k < 2 && (uint64_t)k < counter; ++k) QSV_CUDA(cudaStreamWaitEvent(sv.stream, d.ev_copy[k], 0));
    end_timed_swap(d, (uint64_t)(sv.length() / 2) * ab * items.size());
}

namespace {

// swap so that logical qubit `lq` (currently global) becomes local, evicting local physical bit l
void swap_logical_in(State &sv, int gphys, int l, size_t chunk_bytes) {
    DistCtx &d = *sv.dist;
    dist_swap_physical(sv, gphys, l, chunk_bytes);
    const int a = d.log_of[gphys], b = d.log_of[l];
    d.log_of[gphys] = b;
    d.log_of[l] = a;
    d.phys_of[a] = l;
    d.phys_of[b] = gphys;
}

}  // namespace

namespace {

// on unless QSV_DIST_FUSED_SWAP=0 (parity on 2 B200s: profiles/r2_dist_check_2gpu_fused.log); needs room for a second
// shard-sized buffer (ensure_shadow), otherwise the exchanges stay in place
bool fused_swap_enabled() {
    const char *e = std::getenv("QSV_DIST_FUSED_SWAP");
    return !(e && std::atoi(e) == 0);
}

// Collective (every rank calls it at the same point of the same plan): allocate and peer-map the second buffer.  Any
// rank short of memory, or any failed mapping, switches the feature off on all ranks.
bool ensure_shadow(State &sv) {
    DistCtx &d = *sv.dist;
    if (d.shadow_tried) return d.shadow != nullptr && d.shadow_pb.ok;
    d.shadow_tried = true;
    if (!d.p2p || !d.main.ok) return false;
    size_t free_b = 0, total_b = 0;
    void *buf = nullptr;
    const size_t bytes = std::max<size_t>(sv.bytes(), (size_t)2 << 20);  // own IPC handle, see TempVec
    if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && free_b > bytes + ((size_t)2 << 30)) {
        if (cudaMalloc(&buf, bytes) != cudaSuccess) buf = nullptr;
    }
    cudaGetLastError();
    d.shadow_pb = register_peer_buffer(sv, buf);  // null on any rank => not ok on every rank
    if (!d.shadow_pb.ok) {
        release_peer_buffer(d.shadow_pb);
        if (buf) cudaFree(buf);
        return false;
    }
    d.shadow = buf;
    return true;
}

}  // namespace

namespace {
void *const *pointer_table(State &sv, const std::vector<void *> &vecs);
}

// ---- exchange schedules priced with the sweep cost model --------------------------------------------------------------
// plan_dist_steps minimises the number of exchanges; what a step costs, though, is exchanges PLUS the fused sweeps of the
// batches between them, and where the exchanges fall decides how well those batches pack.  For registers of 26 local
// qubits and more (where a sweep is milliseconds) a few schedules -- look-ahead depth 1 / 2 / 3, eviction window 8 / 4 / 12
// top bits -- are priced with the model of tools/sweep_cost_model.py (regs_sweep_model_cost on the programs the planner
// builds for every batch, + 6 ms for an exchange a sweep carries, + 13 ms for one that needs its own pass; all at 30 local
// qubits, the comparison does not depend on the size) and the cheapest is executed.  Rank-independent by construction
// (global controls are priced as satisfied), so every rank picks the same schedule.  Results are cached by circuit
// structure and qubit map: a variational loop or a benchmark plans each (circuit, map) once.
struct DistPlanEntry {
    uint64_t h1, h2;
    std::vector<DistStep> steps;
    std::vector<int> phys, log;
};
std::mutex g_dist_plan_mu;
std::vector<DistPlanEntry> g_dist_plan_cache;

double dist_schedule_cost(const std::vector<LoweredGate> &lowered, const std::vector<DistStep> &steps,
                          std::vector<int> phys, std::vector<int> log, int n_local, int dtype, bool out_of_place) {
    const int L = 4;
    double cost = 0.0;
    std::vector<LoweredGate> batch;
    const uint64_t all_hi = ~((1ull << n_local) - 1ull);  // every global control satisfied: the rank with the most work
    auto close_batch = [&](bool exchange_follows) {
        bool carried = false;
        if (!batch.empty()) {
            const std::vector<LoweredGate> merged = prepare_gates_regs(batch);
            const std::vector<SweepPlan> plan = plan_sweeps_regs(n_local, merged, L, true, 48, 512, dtype, 1);
            std::vector<const LoweredGate *> cur;
            for (size_t k = 0; k < plan.size(); ++k) {
                const SweepPlan &sw = plan[k];
                const bool last = k + 1 == plan.size();
                const bool tile = sw.fused || (last && exchange_follows && out_of_place && regs_fusable(merged[sw.gates[0]], n_local));
                if (!tile) {
                    const LoweredGate &g = merged[sw.gates[0]];
                    cost += 5.3 / (double)(1u << std::min(6, __builtin_popcountll(g.ctrl_mask)));
                    continue;
                }
                cur.clear();
                for (int i : sw.gates) cur.push_back(&merged[i]);
                cost += regs_sweep_model_cost(n_local, dtype, cur, sw.need, L, true);
                if (last) carried = true;
            }
            batch.clear();
        }
        if (exchange_follows) cost += out_of_place ? (carried ? 6.0 : 13.0) : 12.6;
    };
    for (const DistStep &st : steps) {
        if (st.kind == 0) {
            close_batch(true);
            const int a = log[st.a], b = log[st.b];
            log[st.a] = b;
            log[st.b] = a;
            phys[a] = st.b;
            phys[b] = st.a;
            continue;
        }
        LoweredGate g = localize_gate(remap_gate(lowered[st.a], phys), n_local, all_hi);
        if (g.kind != LoweredGate::NOP) batch.push_back(std::move(g));
    }
    close_batch(false);
    return cost;
}

std::vector<DistStep> plan_dist_steps_priced(const std::vector<LoweredGate> &lowered, const std::vector<uint64_t> &dense,
                                             const std::vector<uint64_t> &diag, std::vector<int> &phys_of,
                                             std::vector<int> &log_of, int n_local, int dtype, bool out_of_place) {
    static const int tries = [] {
        const char *e = std::getenv("QSV_DIST_PLAN_TRIES");
        return e ? std::max(1, std::min(6, std::atoi(e))) : 5;
    }();
    const char *min_env = std::getenv("QSV_DIST_PLAN_MIN_LOCAL");  // tests lower it to exercise the search on small shards
    const int min_local = std::max(13, min_env ? std::atoi(min_env) : 26);
    if (tries <= 1 || n_local < min_local || !dist_dag_enabled() || lowered.size() < 8 || lowered.size() > 4096)
        return plan_dist_steps(dense, diag, phys_of, log_of, n_local);
    // cache key: gate structure (kinds, bits) + qubit map
    uint64_t h1 = 1469598103934665603ull, h2 = 0x9e3779b97f4a7c15ull;
    auto mix = [&](uint64_t v) {
        h1 = (h1 ^ v) * 1099511628211ull;
        h2 = (h2 + v) * 0xff51afd7ed558ccdull;
        h2 ^= h2 >> 29;
    };
    mix((uint64_t)n_local);
    mix((uint64_t)dtype);
    mix((uint64_t)out_of_place);
    for (size_t i = 0; i < lowered.size(); ++i) {
        mix(dense[i]);
        mix(diag[i]);
        mix((uint64_t)lowered[i].kind * 131u + (uint64_t)lowered[i].k);
    }
    for (int p : phys_of) mix((uint64_t)p);
    {
        std::lock_guard<std::mutex> lock(g_dist_plan_mu);
        for (size_t i = 0; i < g_dist_plan_cache.size(); ++i)
            if (g_dist_plan_cache[i].h1 == h1 && g_dist_plan_cache[i].h2 == h2) {
                if (i != 0) std::rotate(g_dist_plan_cache.begin(), g_dist_plan_cache.begin() + i, g_dist_plan_cache.begin() + i + 1);
                phys_of = g_dist_plan_cache[0].phys;
                log_of = g_dist_plan_cache[0].log;
                return g_dist_plan_cache[0].steps;
            }
    }
    static const int cand[6][2] = {{2, 8}, {1, 8}, {3, 8}, {2, 4}, {2, 12}, {1, 4}};  // (look-ahead depth, eviction window)
    DistPlanEntry best;
    double best_cost = 0.0;
    size_t best_exchanges = 0;
    for (int c = 0; c < tries; ++c) {
        std::vector<int> phys = phys_of, log = log_of;
        std::vector<DistStep> steps = plan_dist_steps(dense, diag, phys, log, n_local, cand[c][0], cand[c][1]);
        size_t n_x = 0;
        for (const DistStep &st : steps) n_x += st.kind == 0 ? 1 : 0;
        const double cost = dist_schedule_cost(lowered, steps, phys_of, log_of, n_local, dtype, out_of_place);
        if (std::getenv("QSV_DIST_PLAN_DEBUG"))
            fprintf(stderr, "[qsv] exchange schedule depth=%d window=%d: %zu exchanges, model cost %.1f ms (30 local qubits)\n",
                    cand[c][0], cand[c][1], n_x, cost);
        if (c == 0 || cost < best_cost - 1e-9 || (cost < best_cost + 1e-9 && n_x < best_exchanges)) {
            best_cost = cost;
            best_exchanges = n_x;
            best.steps = std::move(steps);
            best.phys = std::move(phys);
            best.log = std::move(log);
        }
    }
    best.h1 = h1;
    best.h2 = h2;
    phys_of = best.phys;
    log_of = best.log;
    std::vector<DistStep> out = best.steps;
    std::lock_guard<std::mutex> lock(g_dist_plan_mu);
    g_dist_plan_cache.insert(g_dist_plan_cache.begin(), std::move(best));
    if (g_dist_plan_cache.size() > 64) g_dist_plan_cache.pop_back();
    return out;
}

// Gates on LOGICAL bits of the whole register, applied to the register itself (vecs == nullptr) or to a set of its
// companions (lambda and the bras of the adjoint method): exchanges scheduled over the dependency DAG, everything between
// two exchanges through the fused tile executor (or gate by gate when fuse is false).
void dist_run_lowered(State &sv, const std::vector<LoweredGate> &lowered, bool fuse, size_t chunk_bytes,
                      const std::vector<void *> *vecs) {
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    std::vector<uint64_t> dense(lowered.size(), 0), diag(lowered.size(), 0);
    for (size_t i = 0; i < lowered.size(); ++i) gate_bit_masks(lowered[i], dense[i], diag[i]);
    for (uint64_t m : dense)
        QSV_CHECK(__builtin_popcountll(m) <= n_local, "gate acts on more wires than a shard holds");
    // the schedule is planned on a copy of the qubit map; the exchanges below replay it on the real one
    std::vector<int> plan_phys = d.phys_of, plan_log = d.log_of;
    // will the exchanges of this call go through the second buffer (a sweep can carry them)?  Same test as try_fused below,
    // minus the allocation; rank-independent
    const bool oop_mode = fuse && vecs == nullptr && fused_swap_enabled() && d.comp.empty() && !d.skip_main && d.p2p &&
                          (d.shadow != nullptr || !d.shadow_tried);
    const std::vector<DistStep> steps =
        fuse ? plan_dist_steps_priced(lowered, dense, diag, plan_phys, plan_log, n_local, sv.dtype, oop_mode)
             : plan_dist_steps(dense, diag, plan_phys, plan_log, n_local);

    // Exchanges carried by the sweeps around them (fused_swap_enabled): the register alternates between its own buffer
    // and the second one; `home` is restored when the call ends.
    size_t n_exchanges = 0;
    for (const DistStep &st : steps) n_exchanges += st.kind == 0 ? 1 : 0;
    const bool try_fused = fuse && vecs == nullptr && fused_swap_enabled() && n_exchanges > 0 && d.comp.empty() &&
                           !d.skip_main && sv.data == d.main.local && ensure_shadow(sv);
    void *const home = sv.data;
    struct Home {  // the register's pointer is back in place whatever happens below
        State &sv;
        void *p;
        ~Home() { sv.data = p; }
    } restore_home{sv, home};
    bool in_shadow = false;
    auto buf = [&](bool shadow) -> const PeerBuf & { return shadow ? d.shadow_pb : d.main; };
    // second half of a split exchange, owed to the next batch: the amplitudes the partner parked are still over there
    struct PendingPull {
        bool on = false;
        int peer = 0, bit = 0, val = 0, stash = 0;
    } pp;
    // QSV_DIST_SPLIT_XCHG: 0 = off; n > 0 = split when >= n gates follow the exchange.  Default: 24 on up to 4 GPUs, where the
    // split has been run on hardware (2 and 4 B200s: one and two partners per rank); off beyond that until it has been.
    static const int split_env = [] {
        const char *e = std::getenv("QSV_DIST_SPLIT_XCHG");
        return e ? std::atoi(e) : -1;
    }();
    const int split_min_gates = split_env >= 0 ? split_env : (d.world <= 4 ? 24 : 0);

    std::vector<LoweredGate> batch;
    auto flush = [&](FusedExchange *fx) {
        if (vecs) {
            if (batch.empty()) return;
            void *const *table = pointer_table(sv, *vecs);
            if (fuse) {
                apply_gates_tiled(sv, batch, table, (int)vecs->size(), nullptr);
            } else {
                for (const auto &g : batch) launch_gate_multi(sv, g, table, (int)vecs->size());
            }
        } else if (fuse) {
            FusedPull pull;
            FusedPull *pl = nullptr;
            if (pp.on) {
                // the register lives in buf(in_shadow); what the partner parked sits in ITS buffer of the same role
                pull.in_peer = buf(in_shadow).peer[pp.peer];
                pull.out_mine = buf(!in_shadow).local;
                pull.local_bit = pp.bit;
                pull.my_value = pp.val;
                pull.stash_bit = pp.stash;
                const int peer1 = pp.peer;
                // the partner has read what this rank parked for it: the buffer may be written again
                pull.after = [&sv, peer1] { handshake(sv, peer1); };
                pl = &pull;
            }
            if (!batch.empty() || pl) apply_gates_tiled(sv, batch, nullptr, 1, fx, pl);
            if (pl) {
                in_shadow = !in_shadow;  // the callee has set sv.data = pull.out_mine
                pp.on = false;
                d.n_pull += 1;
                d.n_pull_carried += pull.carried ? 1 : 0;
            }
        } else {
            for (const auto &g : batch) launch_gate(sv, g);
        }
        batch.clear();
    };
    for (size_t si = 0; si < steps.size(); ++si) {
        const DistStep &st = steps[si];
        if (st.kind == 0) {
            // 16-byte units: index bit 0 of a complex64 shard cannot be exchanged out of place (rank-independent test)
            if (try_fused && !(sv.dtype == QSV_C64 && st.b == 0)) {
                const int gb = st.a - n_local;
                const int peer = d.rank ^ (1 << gb);
                // Split the exchange between this batch's last sweep (push) and the next batch's first (pull) when enough
                // gates follow for that batch to have more than one sweep: each carrier then moves a quarter of the shard
                // over NVLink instead of one sweep moving half of it at the link's full rate.  Decided on the global step
                // list, so every rank decides alike.
                size_t following = 0;
                while (si + 1 + following < steps.size() && steps[si + 1 + following].kind == 1) ++following;
                const bool split = split_min_gates > 0 && (int)following >= split_min_gates && n_local >= 12 &&
                                   regs_pull_supported();
                const int stash = st.b == 4 ? 5 : 4;
                // a pull owed to THIS batch runs first and moves the register: the push targets are what is "other" then
                const bool shadow_at_push = pp.on ? !in_shadow : in_shadow;
                FusedExchange fx;
                for (int k = 0; k < 2; ++k) {
                    fx.out_mine[k] = buf(!shadow_at_push).local;
                    fx.out_peer[k] = buf(!shadow_at_push).peer[peer];
                }
                fx.local_bit = st.b;
                fx.my_value = (d.rank >> gb) & 1;
                fx.stash_bit = split ? stash : -1;
                // the partner has finished everything that read or wrote its other buffer
                fx.before = [&sv, peer] { handshake(sv, peer); };
                flush(&fx);
                if (fx.done) {
                    d.n_fused += 1;
                } else {
                    // no sweep to carry it on this rank (empty batch, or a last gate the tile kernel does not take): the same
                    // stores from a copy pass.  The partner cannot tell the difference, so this choice need not be the same
                    // on both sides.
                    fx.before();
                    launch_xchg_push_copy(sv, sv.data, fx.out_mine[0], fx.out_peer[0], st.b, fx.my_value, fx.stash_bit);
                }
                handshake(sv, peer);  // the partner has finished writing into this rank's other buffer
                in_shadow = !in_shadow;
                sv.data = in_shadow ? d.shadow : home;
                const int a = d.log_of[st.a], b = d.log_of[st.b];
                d.log_of[st.a] = b;
                d.log_of[st.b] = a;
                d.phys_of[a] = st.b;
                d.phys_of[b] = st.a;
                d.n_oop += 1;  // not in n_swaps / total_bytes / total_ms: those describe the timed in-place exchanges
                if (split) {
                    pp.on = true;
                    pp.peer = peer;
                    pp.bit = st.b;
                    pp.val = fx.my_value;
                    pp.stash = stash;
                }
                continue;
            }
            flush(nullptr);
            if (in_shadow) {
                // the in-place exchange works on the peer mappings of the buffer the register lives in
                std::swap(d.main, d.shadow_pb);
                struct Unswap {
                    DistCtx &d;
                    ~Unswap() { std::swap(d.main, d.shadow_pb); }
                } unswap{d};
                swap_logical_in(sv, st.a, st.b, chunk_bytes);
            } else {
                swap_logical_in(sv, st.a, st.b, chunk_bytes);
            }
            continue;
        }
        LoweredGate g = localize_gate(remap_gate(lowered[st.a], d.phys_of), n_local, sv.index_hi);
        if (g.kind != LoweredGate::NOP) batch.push_back(std::move(g));
    }
    flush(nullptr);
    if (in_shadow) {
        // an odd number of fused exchanges.  A register that owns its memory simply lives in the other buffer from now on
        // (the two buffers and their peer mappings trade places; qsv_data_ptr follows); a borrowed buffer (a torch
        // tensor) has to hold the state when the call returns: one device-to-device copy of the shard.
        if (sv.owns) {
            std::swap(d.main, d.shadow_pb);
            void *other = d.shadow;
            d.shadow = home;
            restore_home.p = other;
        } else {
            QSV_CUDA(cudaMemcpyAsync(home, d.shadow, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
        }
    }
    QSV_CHECK(d.phys_of == plan_phys, "internal: the executed exchanges do not match the planned qubit map");
}

void dist_apply_ops(State &sv, const Ops &ops, bool fuse, size_t chunk_bytes) {
    sv.use();
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    const int n_total = sv.dist->n_total;
    sv.stat_launches = 0;
    sv.stat_sweeps = 0;
    std::vector<LoweredGate> lowered;
    lowered.reserve(ops.ops.size());
    for (const auto &op : ops.ops) {
        if (op.name == "Identity") continue;
        lowered.push_back(lower_op_total(n_total, op, false));
    }
    dist_run_lowered(sv, lowered, fuse, chunk_bytes, nullptr);
}

void dist_allreduce(State &sv, double *host, int count) {
    sv.use();
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv.dist;
    QSV_CHECK(count >= 0 && count <= 4096, "allreduce supports up to 4096 doubles");
    QSV_CUDA(cudaMemcpyAsync(d.red_dev, host, count * sizeof(double), cudaMemcpyHostToDevice, sv.stream));
    QSV_NCCL(ncclAllReduce(d.red_dev, d.red_dev, count, ncclDouble, ncclSum, d.comm, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(host, d.red_dev, count * sizeof(double), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

// Re sum_t c_t <P_t> over the sharded register; per-term values in per_term (may be null)
void dist_expval_pauli(State &sv, int n_terms, const uint64_t *x_log, const uint64_t *z_log, const int *ny,
                       const double *coeffs, double *per_term, double *out, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    const uint64_t hi_mask = ~((1ull << n_local) - 1ull);
    std::vector<double> vals(2 * (size_t)std::max(n_terms, 1), 0.0);
    // groups of words whose X/Y letters are all on local qubits under one layout: one batched, fused launch sequence per
    // group (launch_bra_paulis_ket); X/Y letters on global qubits make those qubits local first
    std::vector<char> left(n_terms, 1);
    int n_left = n_terms;
    while (n_left > 0) {
        int first = 0;
        while (!left[first]) ++first;
        for (int lb = 0; lb < d.n_total; ++lb) {
            if (!(x_log[first] >> lb & 1) || d.phys_of[lb] < n_local) continue;
            int best = -1;
            for (int l = n_local - 1; l >= 0; --l)
                if (!(x_log[first] >> d.log_of[l] & 1)) {
                    best = l;
                    break;
                }
            QSV_CHECK(best >= 0, "Pauli word flips more qubits than a shard holds");
            swap_logical_in(sv, d.phys_of[lb], best, chunk_bytes);
        }
        std::vector<int> group;
        std::vector<uint64_t> gx, gz;
        std::vector<int> gy;
        std::vector<double> sgn;
        for (int t = first; t < n_terms; ++t) {
            if (!left[t]) continue;
            const uint64_t x = remap_mask(x_log[t], d.phys_of);
            if (x & hi_mask) continue;
            const uint64_t z = remap_mask(z_log[t], d.phys_of);
            group.push_back(t);
            gx.push_back(x);
            gz.push_back(z & ~hi_mask);
            gy.push_back(ny[t]);
            sgn.push_back((__builtin_popcountll(sv.index_hi & z & hi_mask) & 1) ? -1.0 : 1.0);
            left[t] = 0;
            --n_left;
        }
        double *red = sv.reduction_buffer(2 * group.size());
        reduction_zero(sv, red, 2 * group.size());
        launch_bra_paulis_ket(sv, sv.data, sv.data, (int)group.size(), gx.data(), gz.data(), gy.data(), 0, red);
        std::vector<double> h(2 * group.size());
        reduction_read(sv, red, h.data(), h.size());
        for (size_t q = 0; q < group.size(); ++q) {
            vals[2 * group[q]] = sgn[q] * h[2 * q];
            vals[2 * group[q] + 1] = sgn[q] * h[2 * q + 1];
        }
    }
    for (int off = 0; off < 2 * n_terms; off += 4096)
        dist_allreduce(sv, vals.data() + off, std::min(4096, 2 * n_terms - off));
    double tot = 0;
    for (int t = 0; t < n_terms; ++t) {
        if (per_term) per_term[t] = vals[2 * t];
        if (coeffs) tot += coeffs[2 * t] * vals[2 * t];
    }
    if (out) *out = tot;
}

// ------------------------------------------------------------------------------------------------
// Everything below: measurements, observables and the adjoint method on the sharded register
// (StateVectorCudaMPI.hpp:957-1595, ObservablesGPUMPI.hpp, AdjointDiffGPUMPI.hpp:248-437 of the reference).
// ------------------------------------------------------------------------------------------------
namespace {

void allreduce_vec(State &sv, double *host, size_t count) {
    for (size_t off = 0; off < count; off += 4096) dist_allreduce(sv, host + off, (int)std::min<size_t>(4096, count - off));
}

// stream barrier across all ranks: everything queued before it on every rank has completed when it completes
void stream_barrier(State &sv) {
    DistCtx &d = *sv.dist;
    QSV_NCCL(ncclAllReduce(d.hs_dev, d.hs_dev, 1, ncclInt, ncclMax, d.comm, sv.stream));
}

// make the logical qubits in `need` local; victims are local qubits outside `need`, highest physical bit first
void dist_localize(State &sv, uint64_t need, size_t chunk_bytes) {
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    QSV_CHECK(__builtin_popcountll(need) <= n_local, "operation acts on more wires than a shard holds");
    for (int lb = 0; lb < d.n_total; ++lb) {
        if (!(need >> lb & 1) || d.phys_of[lb] < n_local) continue;
        int best = -1;
        for (int l = n_local - 1; l >= 0; --l)
            if (!(need >> d.log_of[l] & 1)) {
                best = l;
                break;
            }
        QSV_CHECK(best >= 0, "no local qubit can be evicted");
        swap_logical_in(sv, d.phys_of[lb], best, chunk_bytes);
    }
}

// device table of vector pointers for launch_gate_multi (re-uploaded only when it changes)
void *const *pointer_table(State &sv, const std::vector<void *> &vecs) {
    DistCtx &d = *sv.dist;
    QSV_CHECK(vecs.size() <= 256, "too many vectors in one batched launch");
    if (!d.tab_dev) QSV_CUDA(cudaMalloc(&d.tab_dev, 256 * sizeof(void *)));
    if (d.tab_cache != vecs) {
        QSV_CUDA(cudaMemcpyAsync(d.tab_dev, vecs.data(), vecs.size() * sizeof(void *), cudaMemcpyHostToDevice, sv.stream));
        d.tab_cache = vecs;
    }
    return d.tab_dev;
}

// gate given on LOGICAL bits of the whole register -> gate on this rank's shard (NOP when a global control is 0)
LoweredGate dist_prepare_gate(State &sv, const LoweredGate &g, size_t chunk_bytes) {
    DistCtx &d = *sv.dist;
    if (g.kind == LoweredGate::NOP) return g;
    if (g.kind == LoweredGate::DENSE) dist_localize(sv, touched_mask(g), chunk_bytes);
    return localize_gate(remap_gate(g, d.phys_of), sv.n, sv.index_hi);
}

void dist_apply_gate(State &sv, const LoweredGate &g_logical, const std::vector<void *> &vecs, size_t chunk_bytes) {
    const LoweredGate g = dist_prepare_gate(sv, g_logical, chunk_bytes);
    if (g.kind == LoweredGate::NOP) return;
    if (vecs.size() == 1 && vecs[0] == sv.data)
        launch_gate(sv, g);
    else
        launch_gate_multi(sv, g, pointer_table(sv, vecs), (int)vecs.size());
}

// a shard-sized temporary that follows every exchange of the register (collective: all ranks create / destroy
// their temporaries in the same order)
struct TempVec {
    State &sv;
    void *data = nullptr;
    TempVec(State &s, const void *copy_from) : sv(s) {
        DistCtx &d = *sv.dist;
        // at least 2 MiB: smaller cudaMalloc blocks are sub-allocated and would share one IPC handle
        QSV_CUDA(cudaMalloc(&data, std::max<size_t>(sv.bytes(), (size_t)2 << 20)));
        if (copy_from)
            QSV_CUDA(cudaMemcpyAsync(data, copy_from, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
        else
            QSV_CUDA(cudaMemsetAsync(data, 0, sv.bytes(), sv.stream));
        DistCtx::Companion c;
        c.data = data;
        if (d.p2p) c.pb = register_peer_buffer(sv, data);
        d.comp.push_back(std::move(c));
    }
    ~TempVec() {
        DistCtx &d = *sv.dist;
        cudaStreamSynchronize(sv.stream);
        for (size_t i = 0; i < d.comp.size(); ++i)
            if (d.comp[i].data == data) {
                release_peer_buffer(d.comp[i].pb);
                d.comp.erase(d.comp.begin() + i);
                break;
            }
        d.tab_cache.clear();
        cudaFree(data);
    }
    TempVec(const TempVec &) = delete;
    TempVec &operator=(const TempVec &) = delete;
};

const PeerBuf *peer_buf_of(State &sv, const void *data) {
    DistCtx &d = *sv.dist;
    if (data == sv.data) return &d.main;
    for (auto &c : d.comp)
        if (c.data == data) return &c.pb;
    return nullptr;
}

void canonicalize_all(State &sv, size_t chunk_bytes);

// rows [rank * 2^n_local, (rank + 1) * 2^n_local) of a CSR matrix on the device, offsets rebased to 0
struct CsrBlock {
    void *indptr = nullptr, *indices = nullptr, *values = nullptr;
    int64_t nnz = 0;
    CsrBlock(State &sv, const int64_t *indptr_h, const int64_t *indices_h, const cplx *values_h) {
        DistCtx &d = *sv.dist;
        const int64_t rows = (int64_t)sv.length();
        const int64_t r0 = (int64_t)d.rank * rows;
        const int64_t lo = indptr_h[r0];
        nnz = indptr_h[r0 + rows] - lo;
        std::vector<int64_t> ptr(rows + 1);
        for (int64_t i = 0; i <= rows; ++i) ptr[i] = indptr_h[r0 + i] - lo;
        QSV_CUDA(cudaMalloc(&indptr, (rows + 1) * 8));
        QSV_CUDA(cudaMalloc(&indices, std::max<int64_t>(nnz, 1) * 8));
        QSV_CUDA(cudaMalloc(&values, std::max<int64_t>(nnz, 1) * 16));
        QSV_CUDA(cudaMemcpyAsync(indptr, ptr.data(), (rows + 1) * 8, cudaMemcpyHostToDevice, sv.stream));
        if (nnz) {
            QSV_CUDA(cudaMemcpyAsync(indices, indices_h + lo, nnz * 8, cudaMemcpyHostToDevice, sv.stream));
            QSV_CUDA(cudaMemcpyAsync(values, values_h + lo, nnz * 16, cudaMemcpyHostToDevice, sv.stream));
        }
        QSV_CUDA(cudaStreamSynchronize(sv.stream));  // ptr goes out of scope
    }
    ~CsrBlock() {
        cudaFree(indptr);
        cudaFree(indices);
        cudaFree(values);
    }
    CsrBlock(const CsrBlock &) = delete;
    CsrBlock &operator=(const CsrBlock &) = delete;
};

// y (may be null) = H x restricted to this rank's rows and / or accumulate <x|Hx> (local part) into red[0..1].
// The register must be in the canonical layout.  x is gathered from the other ranks' shards over NVLink through
// their peer mappings; without peer access the whole vector is all-gathered first (small registers only).
void dist_csr(State &sv, const void *x, void *y, const int64_t *indptr_h, const int64_t *indices_h, const cplx *values_h,
              double *red) {
    DistCtx &d = *sv.dist;
    const int64_t rows = (int64_t)sv.length();
    CsrBlock blk(sv, indptr_h, indices_h, values_h);
    const PeerBuf *pb = peer_buf_of(sv, x);
    if (d.p2p && pb && pb->ok) {
        void **tab = nullptr;
        QSV_CUDA(cudaMalloc(&tab, d.world * sizeof(void *)));
        QSV_CUDA(cudaMemcpyAsync(tab, pb->peer.data(), d.world * sizeof(void *), cudaMemcpyHostToDevice, sv.stream));
        stream_barrier(sv);  // every rank's x is complete
        launch_csr_sharded(sv, tab, sv.n, x, y, blk.indptr, blk.indices, blk.values, rows, blk.nnz, 8, red, 0);
        stream_barrier(sv);  // nobody reads this rank's x any more
        QSV_CUDA(cudaStreamSynchronize(sv.stream));
        QSV_CUDA(cudaFree(tab));
        return;
    }
    void *full = nullptr;
    QSV_CHECK(d.n_total <= 30, "sparse Hamiltonians on a sharded register need peer access between the GPUs");
    QSV_CUDA(cudaMalloc(&full, sv.bytes() * (size_t)d.world));
    QSV_NCCL(ncclAllGather(x, full, sv.bytes(), ncclChar, d.comm, sv.stream));
    void **tab = nullptr;
    std::vector<void *> ptrs(d.world);
    for (int r = 0; r < d.world; ++r) ptrs[r] = (char *)full + (size_t)r * sv.bytes();
    QSV_CUDA(cudaMalloc(&tab, d.world * sizeof(void *)));
    QSV_CUDA(cudaMemcpyAsync(tab, ptrs.data(), d.world * sizeof(void *), cudaMemcpyHostToDevice, sv.stream));
    launch_csr_sharded(sv, tab, sv.n, x, y, blk.indptr, blk.indices, blk.values, rows, blk.nnz, 8, red, 0);
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
    QSV_CUDA(cudaFree(tab));
    QSV_CUDA(cudaFree(full));
}

// target <- O target, where target is the register's shard or one of its companions
void dist_apply_observable(State &sv, const Obs &o, void *target, size_t chunk_bytes) {
    DistCtx &d = *sv.dist;
    const int n_total = d.n_total;
    switch (o.kind) {
    case Obs::NAMED: {
        if (o.name == "Identity") return;
        Op op;
        op.name = o.name;
        op.wires = o.wires;
        op.params = o.params;
        op.matrix = o.matrix;
        dist_apply_gate(sv, lower_op_total(n_total, op, false), {target}, chunk_bytes);
        return;
    }
    case Obs::HERMITIAN: {
        const size_t dim = 1ull << o.wires.size();
        QSV_CHECK(o.matrix.size() == dim * dim, "Hermitian matrix does not match its wires");
        dist_apply_gate(sv, lower_matrix(n_total, o.matrix.data(), {}, o.wires, false), {target}, chunk_bytes);
        return;
    }
    case Obs::TENSOR:
        for (const auto &c : o.children) dist_apply_observable(sv, *c, target, chunk_bytes);
        return;
    case Obs::HAMILTONIAN: {
        std::vector<uint64_t> xs, zs;
        std::vector<cplx> cf;
        TempVec acc(sv, nullptr);
        if (hamiltonian_of_pauli_words(o, n_total, xs, zs, cf)) {
            // groups of terms whose X/Y letters are all on local qubits under one layout; each group is one
            // gather launch accumulating into acc
            const uint64_t hi_mask = ~((1ull << sv.n) - 1ull);
            std::vector<char> left(xs.size(), 1);
            size_t n_left = xs.size();
            while (n_left) {
                size_t first = 0;
                while (!left[first]) ++first;
                dist_localize(sv, xs[first], chunk_bytes);
                std::vector<uint64_t> gx, gz;
                std::vector<cplx> gc;
                for (size_t t = first; t < xs.size(); ++t) {
                    if (!left[t]) continue;
                    const uint64_t x = remap_mask(xs[t], d.phys_of);
                    if (x & hi_mask) continue;
                    const uint64_t z = remap_mask(zs[t], d.phys_of);
                    const double sgn = (__builtin_popcountll(sv.index_hi & z & hi_mask) & 1) ? -1.0 : 1.0;
                    gx.push_back(x);
                    gz.push_back(z & ~hi_mask);
                    gc.push_back(sgn * cf[t]);
                    left[t] = 0;
                    --n_left;
                }
                launch_pauli_sum_apply(sv, target, acc.data, (int)gx.size(), gx.data(), gz.data(), gc.data(), true);
            }
        } else {
            for (size_t i = 0; i < o.children.size(); ++i) {
                TempVec tmp(sv, target);
                dist_apply_observable(sv, *o.children[i], tmp.data, chunk_bytes);
                launch_axpy(sv, cplx(o.coeffs[i], 0.0), tmp.data, acc.data);
            }
        }
        QSV_CUDA(cudaMemcpyAsync(target, acc.data, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
        return;
    }
    case Obs::SPARSE: {
        QSV_CHECK(o.indptr.size() == (1ull << n_total) + 1, "sparse Hamiltonian dimension does not match the register");
        canonicalize_all(sv, chunk_bytes);
        TempVec y(sv, nullptr);
        dist_csr(sv, target, y.data, o.indptr.data(), o.indices.data(), o.values.data(), nullptr);
        QSV_CUDA(cudaMemcpyAsync(target, y.data, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
        return;
    }
    }
}

// <sv| M |sv> for a dense / diagonal operator on logical bits: local part, not yet reduced
void dist_expval_gate_local(State &sv, const LoweredGate &g_logical, double out[2], size_t chunk_bytes) {
    out[0] = out[1] = 0.0;
    const LoweredGate g = dist_prepare_gate(sv, g_logical, chunk_bytes);
    if (g.kind == LoweredGate::NOP) return;
    double *red = sv.reduction_buffer(2);
    reduction_zero(sv, red, 2);
    launch_bra_op_ket(sv, sv.data, sv.data, g, red, 0);
    reduction_read(sv, red, out, 2);
}

}  // namespace

// restore the identity qubit map for the register and its companions
namespace {
void canonicalize_all(State &sv, size_t chunk_bytes) {
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    for (int gp = n_local; gp < d.n_total; ++gp) {
        if (d.log_of[gp] == gp) continue;
        int p = d.phys_of[gp];
        if (p >= n_local) {
            swap_logical_in(sv, p, n_local - 1, chunk_bytes);
            p = d.phys_of[gp];
        }
        swap_logical_in(sv, gp, p, chunk_bytes);
    }
    std::vector<void *> vecs;
    if (!d.skip_main) vecs.push_back(sv.data);
    for (auto &c : d.comp) vecs.push_back(c.data);
    for (int b = 0; b < n_local; ++b) {
        while (d.log_of[b] != b) {
            const int other = d.phys_of[b];
            LoweredGate g;
            g.kind = LoweredGate::DENSE;
            g.k = 1;
            g.holes = {std::min(b, other), std::max(b, other)};
            g.offs = {1ull << b, 1ull << other};
            g.mat = {0.0, 1.0, 1.0, 0.0};
            if (vecs.size() == 1 && vecs[0] == sv.data)
                launch_gate(sv, g);
            else if (!vecs.empty())
                launch_gate_multi(sv, g, pointer_table(sv, vecs), (int)vecs.size());
            const int qa = d.log_of[b], qb = d.log_of[other];
            d.log_of[b] = qb;
            d.log_of[other] = qa;
            d.phys_of[qa] = other;
            d.phys_of[qb] = b;
        }
    }
}
}  // namespace

void dist_canonicalize(State &sv, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    canonicalize_all(sv, chunk_bytes);
}

// Re <sv|O|sv> on the sharded register
double dist_obs_expval(State &sv, const Obs &o, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    DistCtx &d = *sv.dist;
    const int n_total = d.n_total;
    uint64_t x = 0, z = 0;
    int ny = 0;
    if (as_pauli_word(o, n_total, x, z, ny)) {
        const double one[2] = {1.0, 0.0};
        double out = 0;
        dist_expval_pauli(sv, 1, &x, &z, &ny, one, nullptr, &out, chunk_bytes);
        return out;
    }
    switch (o.kind) {
    case Obs::NAMED:
    case Obs::HERMITIAN: {
        std::vector<cplx> m = o.matrix;
        if (o.kind == Obs::NAMED && find_gate(o.name) != nullptr) m = named_gate_matrix(o.name, o.params, (int)o.wires.size());
        QSV_CHECK(!m.empty(), "Currently unsupported observable: " + o.name);
        const size_t dim = 1ull << o.wires.size();
        QSV_CHECK(m.size() == dim * dim, "observable matrix does not match its wires");
        double h[2];
        dist_expval_gate_local(sv, lower_matrix(n_total, m.data(), {}, o.wires, false), h, chunk_bytes);
        allreduce_vec(sv, h, 2);
        return h[0];
    }
    case Obs::HAMILTONIAN: {
        std::vector<uint64_t> xs, zs;
        std::vector<cplx> cf;
        if (hamiltonian_of_pauli_words(o, n_total, xs, zs, cf)) {
            // cf carries i^ny; dist_expval_pauli applies i^ny itself, so pass the plain coefficients
            std::vector<int> nys(xs.size());
            std::vector<double> c2(2 * xs.size(), 0.0);
            for (size_t t = 0; t < xs.size(); ++t) {
                nys[t] = __builtin_popcountll(xs[t] & zs[t]);
                c2[2 * t] = o.coeffs[t];
            }
            double out = 0;
            dist_expval_pauli(sv, (int)xs.size(), xs.data(), zs.data(), nys.data(), c2.data(), nullptr, &out, chunk_bytes);
            return out;
        }
        double tot = 0;
        for (size_t i = 0; i < o.children.size(); ++i) tot += o.coeffs[i] * dist_obs_expval(sv, *o.children[i], chunk_bytes);
        return tot;
    }
    case Obs::TENSOR: {
        TempVec tmp(sv, sv.data);
        dist_apply_observable(sv, o, tmp.data, chunk_bytes);
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        LoweredGate id;
        launch_bra_op_ket(sv, sv.data, tmp.data, id, red, 0);
        double h[2];
        reduction_read(sv, red, h, 2);
        allreduce_vec(sv, h, 2);
        return h[0];
    }
    case Obs::SPARSE: {
        QSV_CHECK(o.indptr.size() == (1ull << n_total) + 1, "sparse Hamiltonian dimension does not match the register");
        canonicalize_all(sv, chunk_bytes);
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        dist_csr(sv, sv.data, nullptr, o.indptr.data(), o.indices.data(), o.values.data(), red);
        double h[2];
        reduction_read(sv, red, h, 2);
        allreduce_vec(sv, h, 2);
        return h[0];
    }
    }
    return 0.0;
}

void dist_obs_apply(State &sv, const Obs &o, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    dist_apply_observable(sv, o, sv.data, chunk_bytes);
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

void dist_expval_matrix(State &sv, const cplx *matrix, const std::vector<int> &wires, double out[2], size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    dist_expval_gate_local(sv, lower_matrix(sv.dist->n_total, matrix, {}, wires, false), out, chunk_bytes);
    allreduce_vec(sv, out, 2);
}

// probability(wires) on the sharded register (MPI.hpp:1187-1290): local marginal over the measured qubits that are
// local, placed at the output positions given by this rank's values of the measured global qubits, all-reduced.
// out has 2^k entries, first listed wire = least significant bit (the reference's cuStateVec bit order).
void dist_probs(State &sv, const std::vector<int> &wires, double *out) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    DistCtx &d = *sv.dist;
    const int k = (int)wires.size();
    QSV_CHECK(k >= 1 && k <= 26, "probability needs between 1 and 26 wires on a sharded register");
    std::vector<int> local_bits, local_pos;
    uint64_t fixed = 0;
    for (int i = 0; i < k; ++i) {
        QSV_CHECK(wires[i] >= 0 && wires[i] < d.n_total, "wire out of range");
        for (int j = 0; j < i; ++j) QSV_CHECK(wires[i] != wires[j], "repeated wire in probability");
        const int p = d.phys_of[d.n_total - 1 - wires[i]];
        if (p < sv.n) {
            local_bits.push_back(p);
            local_pos.push_back(i);
        } else if (sv.index_hi >> p & 1) {
            fixed |= 1ull << i;
        }
    }
    const size_t nb = 1ull << k;
    std::vector<double> full(nb, 0.0);
    if (local_bits.empty()) {
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        LoweredGate id;
        launch_bra_op_ket(sv, sv.data, sv.data, id, red, 0);
        double h[2];
        reduction_read(sv, red, h, 2);
        full[fixed] = h[0];
    } else {
        std::vector<double> loc(1ull << local_bits.size());
        launch_probs(sv, local_bits, loc.data());
        for (size_t t = 0; t < loc.size(); ++t) {
            uint64_t idx = fixed;
            for (size_t q = 0; q < local_pos.size(); ++q) idx |= (uint64_t)((t >> q) & 1) << local_pos[q];
            full[idx] = loc[t];
        }
    }
    allreduce_vec(sv, full.data(), nb);
    for (size_t t = 0; t < nb; ++t) out[t] = full[t];
}

// generate_samples on the sharded register (MPI.hpp:1454-1595): the inverse CDF over the canonical amplitude
// order, i.e. rank by rank; every rank samples the shots whose target mass falls into its own shard.
// out[shot * n_total + w] = bit of wire w (identical on all ranks after the all-reduce).
void dist_sample(State &sv, const double *uniforms, int64_t shots, uint64_t *out, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    DistCtx &d = *sv.dist;
    if (shots <= 0) return;
    canonicalize_all(sv, chunk_bytes);
    std::vector<double> mass(d.world, 0.0);
    mass[d.rank] = state_mass(sv);
    allreduce_vec(sv, mass.data(), d.world);
    std::vector<double> cdf(d.world + 1, 0.0);
    for (int r = 0; r < d.world; ++r) cdf[r + 1] = cdf[r] + mass[r];
    const double total = cdf[d.world];
    std::vector<int64_t> mine;
    std::vector<double> targets;
    for (int64_t s = 0; s < shots; ++s) {
        const double t = uniforms[s] * total;
        int r = (int)(std::upper_bound(cdf.begin() + 1, cdf.end(), t) - (cdf.begin() + 1));
        if (r >= d.world) r = d.world - 1;
        if (r == d.rank) {
            mine.push_back(s);
            targets.push_back(t - cdf[r]);
        }
    }
    std::vector<uint64_t> idx(mine.size());
    if (!mine.empty()) launch_sample_indices(sv, targets.data(), (int64_t)mine.size(), idx.data(), true);
    const int n = d.n_total;
    const size_t count = (size_t)shots * n;
    for (size_t i = 0; i < count; ++i) out[i] = 0;
    for (size_t q = 0; q < mine.size(); ++q) {
        const uint64_t g = ((uint64_t)d.rank << sv.n) | idx[q];
        for (int w = 0; w < n; ++w) out[mine[q] * n + w] = (g >> (n - 1 - w)) & 1ull;
    }
    uint64_t *dev = nullptr;
    QSV_CUDA(cudaMalloc(&dev, count * sizeof(uint64_t)));
    QSV_CUDA(cudaMemcpyAsync(dev, out, count * sizeof(uint64_t), cudaMemcpyHostToDevice, sv.stream));
    QSV_NCCL(ncclAllReduce(dev, dev, count, ncclUint64, ncclSum, d.comm, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(out, dev, count * sizeof(uint64_t), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
    QSV_CUDA(cudaFree(dev));
}

// setBasisState / setStateVector on the sharded register (MPI.hpp:332-381): indices address the whole register;
// the qubit map is reset to the identity
void dist_set_state(State &sv, const int64_t *indices, const void *values, size_t count) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    DistCtx &d = *sv.dist;
    for (int b = 0; b < d.n_total; ++b) d.phys_of[b] = d.log_of[b] = b;
    const size_t ab = sv.amp_bytes();
    std::vector<int64_t> li;
    std::vector<char> lv;
    for (size_t i = 0; i < count; ++i) {
        QSV_CHECK(indices[i] >= 0 && (uint64_t)indices[i] < (1ull << d.n_total), "state-vector index out of range");
        if ((int)((uint64_t)indices[i] >> sv.n) != d.rank) continue;
        li.push_back(indices[i] & (int64_t)(sv.length() - 1));
        lv.insert(lv.end(), (const char *)values + i * ab, (const char *)values + (i + 1) * ab);
    }
    QSV_CUDA(cudaMemsetAsync(sv.data, 0, sv.bytes(), sv.stream));
    if (!li.empty()) {
        const size_t vb = li.size() * ab;
        const size_t vb_al = (vb + 15) / 16 * 16;
        char *scr = (char *)sv.scratch_buffer(vb_al + li.size() * 8);
        QSV_CUDA(cudaMemcpyAsync(scr, lv.data(), vb, cudaMemcpyHostToDevice, sv.stream));
        QSV_CUDA(cudaMemcpyAsync(scr + vb_al, li.data(), li.size() * 8, cudaMemcpyHostToDevice, sv.stream));
        launch_scatter(sv, (const int64_t *)(scr + vb_al), scr, li.size());
    }
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
}

// ------------------------------------------------------------------------------------------------
// Adjoint Jacobian on the sharded register (AdjointDiffGPUMPI.hpp:248-437): lambda and the bras are sharded like
// the register and share its qubit map, so they follow every exchange; the register itself stays where it is.
// Jacobian entries are reduced over the ranks once, at the end (the reference all-reduces per (observable,
// parameter), AdjointDiffGPUMPI.hpp:126-129).
// ------------------------------------------------------------------------------------------------
void dist_adjoint_jacobian(State &sv, const Ops &ops, const std::vector<const Obs *> &obs,
                           const std::vector<int64_t> &trainable, bool apply_operations, double *jac, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    sv.use();
    DistCtx &d = *sv.dist;
    const int n_total = d.n_total;
    QSV_CHECK(!trainable.empty(), "No trainable parameters provided.");
    const size_t n_obs = obs.size(), n_tp = trainable.size();
    for (size_t i = 0; i < n_obs * n_tp; ++i) jac[i] = 0.0;
    if (n_obs == 0) return;
    for (const auto &op : ops.ops)
        QSV_CHECK(op.params.size() <= 1, "The operation is not supported using the adjoint differentiation method");

    const std::vector<int> saved_phys = d.phys_of, saved_log = d.log_of;
    struct Restore {
        DistCtx &d;
        const std::vector<int> &p, &l;
        ~Restore() {
            d.skip_main = false;
            d.phys_of = p;
            d.log_of = l;
        }
    } restore{d, saved_phys, saved_log};
    d.skip_main = true;

    TempVec lambda(sv, sv.data);
    if (apply_operations)
        for (const auto &op : ops.ops) {
            if (op.name == "Identity") continue;
            dist_apply_gate(sv, lower_op_total(n_total, op, false), {lambda.data}, chunk_bytes);
        }
    std::vector<std::unique_ptr<TempVec>> bras;
    for (size_t i = 0; i < n_obs; ++i) {
        bras.push_back(std::make_unique<TempVec>(sv, lambda.data));
        dist_apply_observable(sv, *obs[i], bras.back()->data, chunk_bytes);
    }
    std::vector<void *> all_vecs = {lambda.data};
    for (auto &b : bras) all_vecs.push_back(b->data);

    const size_t n_slots = 2 * n_obs * n_tp;
    double *red = sv.reduction_buffer(2 * n_slots);
    reduction_zero(sv, red, 2 * n_slots);
    std::vector<double> factor(n_tp, 0.0), extra(n_tp, 0.0);
    const char *batch_env = std::getenv("QSV_ADJOINT_BATCH");
    if (!(batch_env && std::atoi(batch_env) == 0)) {
        // The layered reverse sweep of the single-GPU path (circuit.cu: layered_reverse_sweep) on the sharded vectors: the
        // generators of a whole layer are evaluated by the batched tile kernel once their dense target qubits are local,
        // and every batch of U^dagger goes through the exchange scheduler and the fused tile executor over lambda and
        // all bras at once.  The reference undoes one gate at a time with two swaps each
        // (algorithms/AdjointDiffGPUMPI.hpp:248-334).
        ReverseSweepHooks hk;
        hk.n_wires = n_total;
        hk.n_obs = n_obs;
        hk.generator = [&](const Op &op) { return lower_generator(n_total, op.name, op.wires); };
        hk.dagger = [&](const Op &op) { return lower_op_total(n_total, op, true); };
        hk.inner_products = [&](const std::vector<LoweredGate> &gens, const std::vector<int> &slot0) {
            uint64_t need = 0;
            for (const LoweredGate &g : gens)
                if (g.kind == LoweredGate::DENSE) need |= touched_mask(g);
            // the ready operations act on disjoint wires; should they need more qubits than a shard holds, go in groups
            std::vector<size_t> todo(gens.size());
            for (size_t k = 0; k < gens.size(); ++k) todo[k] = k;
            while (!todo.empty()) {
                uint64_t group_need = 0;
                std::vector<size_t> group, rest;
                for (size_t k : todo) {
                    const uint64_t t = gens[k].kind == LoweredGate::DENSE ? touched_mask(gens[k]) : 0;
                    if (__builtin_popcountll(group_need | t) <= sv.n - 1 || group.empty()) {
                        group_need |= t;
                        group.push_back(k);
                    } else {
                        rest.push_back(k);
                    }
                }
                if (group_need) dist_localize(sv, group_need, chunk_bytes);
                std::vector<LoweredGate> local;
                std::vector<int> base_slot;
                for (size_t k : group) {
                    LoweredGate gl = localize_gate(remap_gate(gens[k], d.phys_of), sv.n, sv.index_hi);
                    if (gl.kind == LoweredGate::NOP) continue;  // a projector that is zero on this rank
                    local.push_back(std::move(gl));
                    base_slot.push_back(slot0[k]);
                }
                std::vector<int> slots(local.size());
                for (size_t i = 0; i < n_obs && !local.empty(); ++i) {
                    for (size_t k = 0; k < local.size(); ++k) slots[k] = base_slot[k] + (int)(2 * i);
                    launch_bra_gens_ket(sv, bras[i]->data, lambda.data, local, slots, red);
                }
                todo.swap(rest);
            }
            (void)need;
        };
        hk.identity_inner_product = [&](int64_t tp) {
            LoweredGate id;
            for (size_t i = 0; i < n_obs; ++i)
                launch_bra_op_ket(sv, bras[i]->data, lambda.data, id, red, (int)((tp * n_obs + i) * 2 + 1));
        };
        hk.apply = [&](const std::vector<LoweredGate> &batch) { dist_run_lowered(sv, batch, true, chunk_bytes, &all_vecs); };
        layered_reverse_sweep(ops, trainable, factor, extra, hk);
    } else {
    size_t n_par_ops = 0;
    for (const auto &op : ops.ops) n_par_ops += op.params.empty() ? 0 : 1;
    int64_t tp_pos = (int64_t)n_tp - 1;    int64_t cur = (int64_t)n_par_ops - 1;
    for (int64_t idx = (int64_t)ops.ops.size() - 1; idx >= 0; --idx) {
        const Op &op = ops.ops[idx];
        if (op.name == "QubitStateVector" || op.name == "StatePrep" || op.name == "BasisState") continue;
        if (tp_pos < 0) break;
        if (!op.params.empty()) {
            if (cur == trainable[tp_pos]) {
                LoweredGenerator g = lower_generator(n_total, op.name, op.wires);
                factor[tp_pos] = -2.0 * g.scale * (op.inverse ? -1.0 : 1.0);
                extra[tp_pos] = g.extra_identity;
                const LoweredGate gl = dist_prepare_gate(sv, g.op, chunk_bytes);
                for (size_t i = 0; i < n_obs; ++i) {
                    if (gl.kind != LoweredGate::NOP || g.op.kind == LoweredGate::NOP)
                        launch_bra_op_ket(sv, bras[i]->data, lambda.data, gl, red, (int)((tp_pos * n_obs + i) * 2));
                    if (g.extra_identity != 0.0) {
                        LoweredGate id;
                        launch_bra_op_ket(sv, bras[i]->data, lambda.data, id, red, (int)((tp_pos * n_obs + i) * 2 + 1));
                    }
                }
                --tp_pos;
            }
            --cur;
        }
        if (op.name != "Identity") dist_apply_gate(sv, lower_op_total(n_total, op, true), all_vecs, chunk_bytes);
    }
    }
    std::vector<double> h(2 * n_slots);
    reduction_read(sv, red, h.data(), 2 * n_slots);
    allreduce_vec(sv, h.data(), h.size());
    for (size_t p = 0; p < n_tp; ++p)
        for (size_t i = 0; i < n_obs; ++i) {
            const size_t s = (p * n_obs + i) * 2;
            const double im = h[2 * s + 1] + extra[p] * h[2 * (s + 1) + 1];
            jac[i * n_tp + p] = factor[p] * im;
        }
}

void dist_free(State &sv) {
    if (!sv.dist) return;
    DistCtx *d = sv.dist;
    cudaSetDevice(sv.device);
    cudaDeviceSynchronize();
    if (d->comm) ncclCommDestroy(d->comm);
    for (int k = 0; k < 2; ++k) {
        if (d->stage[k]) cudaFree(d->stage[k]);
        if (d->ev_xfer[k]) cudaEventDestroy(d->ev_xfer[k]);
        if (d->ev_copy[k]) cudaEventDestroy(d->ev_copy[k]);
    }
    if (d->ev_ready) cudaEventDestroy(d->ev_ready);
    for (auto &t : d->pending) {
        cudaEventDestroy(t.t0);
        cudaEventDestroy(t.t1);
    }
    for (auto &e : d->free_events) {
        cudaEventDestroy(e.first);
        cudaEventDestroy(e.second);
    }
    if (d->ev_t0) cudaEventDestroy(d->ev_t0);
    if (d->ev_t1) cudaEventDestroy(d->ev_t1);
    if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->red_dev) cudaFree(d->red_dev);
    if (d->hs_dev) cudaFree(d->hs_dev);
    release_peer_buffer(d->main);
    if (d->shadow) {
        release_peer_buffer(d->shadow_pb);
        cudaFree(d->shadow);
    }
    for (auto &c : d->comp) release_peer_buffer(c.pb);
    if (d->tab_dev) cudaFree(d->tab_dev);
    delete d;
    sv.dist = nullptr;
    sv.index_hi = 0;
}

}  // namespace qsv

namespace qsv {

// Map a shard-sized buffer of every other rank into this process with CUDA IPC (collective: every rank calls it
// for its own buffer, in the same order).  The handles travel through NCCL itself (all-gather of 80 bytes per
// rank), so no other rendezvous is needed.  ok == false on all ranks when any mapping failed.
PeerBuf register_peer_buffer(State &sv, void *data) {
    DistCtx &d = *sv.dist;
    PeerBuf pb;
    pb.local = data;
    struct Msg {
        cudaIpcMemHandle_t handle;
        uint64_t offset;
        uint64_t ok;
    };
    static_assert(sizeof(Msg) == 80, "IPC message layout");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    // base of the allocation that contains the buffer (it may be a view into a torch block)
    typedef int (*GetRangeFn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    unsigned long long base = 0;
    size_t size = 0;
    bool ok = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn != nullptr &&
              ((GetRangeFn)fn)(&base, &size, (unsigned long long)(uintptr_t)data) == 0;
    if (ok) ok = cudaIpcGetMemHandle(&mine.handle, (void *)(uintptr_t)base) == cudaSuccess;
    cudaGetLastError();
    mine.offset = ok ? (uint64_t)((uintptr_t)data - base) : 0;
    mine.ok = ok ? 1 : 0;
    Msg *dev = nullptr;
    QSV_CUDA(cudaMalloc(&dev, sizeof(Msg) * (size_t)d.world));
    QSV_CUDA(cudaMemcpyAsync(dev + d.rank, &mine, sizeof(Msg), cudaMemcpyHostToDevice, sv.stream));
    QSV_NCCL(ncclAllGather(dev + d.rank, dev, sizeof(Msg), ncclChar, d.comm, sv.stream));
    std::vector<Msg> all(d.world);
    QSV_CUDA(cudaMemcpyAsync(all.data(), dev, sizeof(Msg) * (size_t)d.world, cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
    QSV_CUDA(cudaFree(dev));
    bool all_ok = true;
    for (const Msg &m : all) all_ok = all_ok && m.ok == 1;
    pb.peer.assign(d.world, nullptr);
    pb.map_base.assign(d.world, nullptr);
    pb.peer[d.rank] = data;
    int mapped_ok = all_ok ? 1 : 0;
    if (all_ok) {
        for (int r = 0; r < d.world && mapped_ok; ++r) {
            if (r == d.rank) continue;
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[r].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
                cudaGetLastError();
                mapped_ok = 0;
                break;
            }
            pb.map_base[r] = p;
            pb.peer[r] = (char *)p + all[r].offset;
        }
    }
    // agree on the outcome
    int *flag = nullptr;
    QSV_CUDA(cudaMalloc(&flag, sizeof(int)));
    QSV_CUDA(cudaMemcpyAsync(flag, &mapped_ok, sizeof(int), cudaMemcpyHostToDevice, sv.stream));
    QSV_NCCL(ncclAllReduce(flag, flag, 1, ncclInt, ncclMin, d.comm, sv.stream));
    QSV_CUDA(cudaMemcpyAsync(&mapped_ok, flag, sizeof(int), cudaMemcpyDeviceToHost, sv.stream));
    QSV_CUDA(cudaStreamSynchronize(sv.stream));
    QSV_CUDA(cudaFree(flag));
    pb.ok = mapped_ok == 1;
    if (!pb.ok) release_peer_buffer(pb);
    return pb;
}

void release_peer_buffer(PeerBuf &pb) {
    for (void *&p : pb.map_base)
        if (p) {
            cudaIpcCloseMemHandle(p);
            p = nullptr;
        }
    pb.ok = false;
}

void setup_peer_access(State &sv) {
    DistCtx &d = *sv.dist;
    d.main = register_peer_buffer(sv, sv.data);
    d.p2p = d.main.ok;
}

}  // namespace qsv

using namespace qsv;

#define QSV_API_BEGIN try {
#define QSV_API_END                                                                                \
    return 0;                                                                                      \
    }                                                                                              \
    catch (const std::exception &e) {                                                              \
        set_last_error(e.what());                                                                  \
        return 1;                                                                                  \
    }                                                                                              \
    catch (...) {                                                                                  \
        set_last_error("unknown error");                                                           \
        return 1;                                                                                  \
    }
// end of a collective entry point: back to the canonical layout unless the register was switched to a lazy map
#define QSV_DIST_SETTLE(sv)                                                                        \
    do {                                                                                           \
        if (!(sv)->dist->lazy_map) dist_canonicalize(*(sv), 0);                                    \
    } while (0)

extern "C" {

int qsv_dist_unique_id(void *id128) {
    QSV_API_BEGIN
    QSV_CHECK(id128 != nullptr, "null id buffer");
    ncclUniqueId id;
    QSV_NCCL(ncclGetUniqueId(&id));
    static_assert(sizeof(id) == 128, "NCCL unique id size");
    memcpy(id128, &id, 128);
    QSV_API_END
}

int qsv_dist_init(qsv_state *sv, const void *id128, int rank, int world_size) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && id128 != nullptr, "null argument");
    QSV_CHECK(sv->dist == nullptr, "distributed context already initialised");
    QSV_CHECK(world_size >= 1 && (world_size & (world_size - 1)) == 0, "number of ranks must be a power of two");
    QSV_CHECK(rank >= 0 && rank < world_size, "invalid rank");
    sv->use();
    auto d = std::make_unique<DistCtx>();
    d->rank = rank;
    d->world = world_size;
    d->n_global = __builtin_ctz(world_size);
    d->n_total = sv->n + d->n_global;
    d->phys_of.resize(d->n_total);
    d->log_of.resize(d->n_total);
    for (int b = 0; b < d->n_total; ++b) d->phys_of[b] = d->log_of[b] = b;
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    QSV_NCCL(ncclCommInitRank(&d->comm, world_size, id, rank));
    QSV_CUDA(cudaStreamCreateWithFlags(&d->comm_stream, cudaStreamNonBlocking));
    QSV_CUDA(cudaStreamCreateWithFlags(&d->copy_stream, cudaStreamNonBlocking));
    QSV_CUDA(cudaEventCreateWithFlags(&d->ev_ready, cudaEventDisableTiming));
    for (int k = 0; k < 2; ++k) {
        QSV_CUDA(cudaEventCreateWithFlags(&d->ev_xfer[k], cudaEventDisableTiming));
        QSV_CUDA(cudaEventCreateWithFlags(&d->ev_copy[k], cudaEventDisableTiming));
    }
    QSV_CUDA(cudaMalloc(&d->red_dev, 4096 * sizeof(double)));
    QSV_CUDA(cudaMalloc(&d->hs_dev, 2 * sizeof(int)));
    QSV_CUDA(cudaMemset(d->hs_dev, 0, 2 * sizeof(int)));
    sv->index_hi = (uint64_t)rank << sv->n;
    sv->dist = d.release();
    const char *p2p_env = std::getenv("QSV_DIST_P2P");
    if (world_size > 1 && !(p2p_env && std::atoi(p2p_env) == 0)) setup_peer_access(*sv);
    QSV_API_END
}

int qsv_dist_finalize(qsv_state *sv) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr, "null state");
    dist_free(*sv);
    QSV_API_END
}

int qsv_dist_swap_bits(qsv_state *sv, int global_bit, int local_bit, size_t chunk_bytes) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    // physical exchange + bookkeeping of the qubit map
    DistCtx &d = *sv->dist;
    dist_swap_physical(*sv, global_bit, local_bit, chunk_bytes);
    const int a = d.log_of[global_bit], b = d.log_of[local_bit];
    d.log_of[global_bit] = b;
    d.log_of[local_bit] = a;
    d.phys_of[a] = local_bit;
    d.phys_of[b] = global_bit;
    QSV_API_END
}

int qsv_dist_apply_ops(qsv_state *sv, const qsv_ops *ops, int fuse, size_t chunk_bytes) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && ops != nullptr, "null argument");
    dist_apply_ops(*sv, *ops, fuse != 0, chunk_bytes);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_canonicalize(qsv_state *sv, size_t chunk_bytes) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr, "null state");
    dist_canonicalize(*sv, chunk_bytes);
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    QSV_API_END
}

int qsv_dist_set_lazy_map(qsv_state *sv, int lazy) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    sv->dist->lazy_map = lazy != 0;
    QSV_DIST_SETTLE(sv);  // collective when it switches a permuted register back to the canonical contract
    QSV_API_END
}

int qsv_dist_qubit_map(const qsv_state *sv, int *phys_of_logical_bit, int n) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    QSV_CHECK(n == sv->dist->n_total && phys_of_logical_bit != nullptr, "map buffer must hold n_total entries");
    for (int b = 0; b < n; ++b) phys_of_logical_bit[b] = sv->dist->phys_of[b];
    QSV_API_END
}

int qsv_dist_allreduce_f64(qsv_state *sv, double *host_values, int count) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && host_values != nullptr, "null argument");
    dist_allreduce(*sv, host_values, count);
    QSV_API_END
}

int qsv_dist_expval_pauli_words(qsv_state *sv, int n_terms, const char *letters, const int *wires,
                                const int *offsets, const double *coeffs, double *per_term, double *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    const int n = sv->dist->n_total;
    std::vector<uint64_t> xs(n_terms), zs(n_terms);
    std::vector<int> nys(n_terms);
    for (int t = 0; t < n_terms; ++t) {
        uint64_t x = 0, z = 0;
        int ny = 0;
        for (int j = offsets[t]; j < offsets[t + 1]; ++j) {
            QSV_CHECK(wires[j] >= 0 && wires[j] < n, "Pauli word wire out of range");
            const uint64_t b = 1ull << (n - 1 - wires[j]);
            switch (letters[j]) {
            case 'I': break;
            case 'X': x |= b; break;
            case 'Y': x |= b; z |= b; ++ny; break;
            case 'Z': z |= b; break;
            default: fail(std::string("invalid Pauli letter '") + letters[j] + "'");
            }
        }
        xs[t] = x;
        zs[t] = z;
        nys[t] = ny;
    }
    dist_expval_pauli(*sv, n_terms, xs.data(), zs.data(), nys.data(), coeffs, per_term, out, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_set_basis_state(qsv_state *sv, uint64_t index) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    const int64_t idx = (int64_t)index;
    if (sv->dtype == QSV_C128) {
        const double one[2] = {1.0, 0.0};
        dist_set_state(*sv, &idx, one, 1);
    } else {
        const float one[2] = {1.0f, 0.0f};
        dist_set_state(*sv, &idx, one, 1);
    }
    QSV_API_END
}

int qsv_dist_set_state_vector(qsv_state *sv, const int64_t *indices, const void *values, size_t count) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    QSV_CHECK(count == 0 || (indices && values), "null index/value arrays");
    dist_set_state(*sv, indices, values, count);
    QSV_API_END
}

int qsv_dist_expval_named(qsv_state *sv, const char *name, const int *wires, int n_wires, const double *params,
                          int n_params, double *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && name && out, "null argument");
    QSV_CHECK(find_gate(name) != nullptr, std::string("Currently unsupported observable: ") + name);
    std::vector<cplx> m = named_gate_matrix(name, std::vector<double>(params, params + n_params), n_wires);
    dist_expval_matrix(*sv, m.data(), std::vector<int>(wires, wires + n_wires), out, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_expval_matrix(qsv_state *sv, const double *matrix, const int *wires, int n_wires, double *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && matrix && out, "null argument");
    QSV_CHECK(n_wires >= 1 && n_wires <= 10, "dense observables act on 1..10 wires");
    const size_t dim = 1ull << n_wires;
    std::vector<cplx> m(dim * dim);
    for (size_t i = 0; i < dim * dim; ++i) m[i] = cplx(matrix[2 * i], matrix[2 * i + 1]);
    dist_expval_matrix(*sv, m.data(), std::vector<int>(wires, wires + n_wires), out, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_expval_csr(qsv_state *sv, const int64_t *row_offsets, const int64_t *col_indices, const double *values,
                        int64_t nnz, double *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && row_offsets && out, "null argument");
    Obs o;
    o.kind = Obs::SPARSE;
    const size_t rows = 1ull << sv->dist->n_total;
    o.indptr.assign(row_offsets, row_offsets + rows + 1);
    QSV_CHECK(o.indptr[rows] == nnz, "CSR offsets do not match nnz");
    if (nnz) {
        o.indices.assign(col_indices, col_indices + nnz);
        o.values.resize(nnz);
        for (int64_t i = 0; i < nnz; ++i) o.values[i] = cplx(values[2 * i], values[2 * i + 1]);
    }
    *out = dist_obs_expval(*sv, o, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_probs(qsv_state *sv, const int *wires, int n_wires, double *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && wires && out, "null argument");
    dist_probs(*sv, std::vector<int>(wires, wires + n_wires), out);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_sample(qsv_state *sv, const double *uniforms, int64_t shots, uint64_t *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    QSV_CHECK(shots >= 0 && (shots == 0 || (uniforms && out)), "invalid sampling arguments");
    dist_sample(*sv, uniforms, shots, out, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_obs_expval(const qsv_obs *obs, qsv_state *sv, double *out) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && obs && out, "null argument");
    *out = dist_obs_expval(*sv, *obs->p, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_obs_apply(const qsv_obs *obs, qsv_state *sv) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && obs, "null argument");
    dist_obs_apply(*sv, *obs->p, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_adjoint_jacobian(qsv_state *sv, const qsv_ops *ops, qsv_obs *const *observables, int n_obs,
                              const int64_t *trainable, int n_trainable, int apply_operations, double *jac) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && ops, "null argument");
    QSV_CHECK(n_obs >= 0 && n_trainable >= 0, "negative sizes");
    std::vector<const Obs *> o;
    for (int i = 0; i < n_obs; ++i) {
        QSV_CHECK(observables[i] != nullptr, "null observable");
        o.push_back(observables[i]->p.get());
    }
    std::vector<int64_t> tp(trainable, trainable + n_trainable);
    QSV_CHECK(n_trainable == 0 || jac != nullptr, "null Jacobian output");
    dist_adjoint_jacobian(*sv, *ops, o, tp, apply_operations != 0, jac, 0);
    QSV_DIST_SETTLE(sv);
    QSV_API_END
}

int qsv_dist_rank(const qsv_state *sv) { return sv && sv->dist ? sv->dist->rank : -1; }
int qsv_dist_world_size(const qsv_state *sv) { return sv && sv->dist ? sv->dist->world : -1; }
int qsv_dist_total_qubits(const qsv_state *sv) { return sv && sv->dist ? sv->dist->n_total : -1; }


/* local shard <-> host in the canonical layout (CopyHostDataToGpu / CopyGpuDataToHost of StateVectorCudaMPI) */
int qsv_dist_h2d(qsv_state *sv, const void *host, size_t n_amps) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && host, "null argument");
    QSV_CHECK(n_amps <= sv->length(), "host buffer larger than the local shard");
    sv->use();
    DistCtx &d = *sv->dist;
    for (int b = 0; b < d.n_total; ++b) d.phys_of[b] = d.log_of[b] = b;
    copy_state_host(*sv, sv->data, const_cast<void *>(host), n_amps * sv->amp_bytes(), true);
    QSV_API_END
}

int qsv_dist_d2h(qsv_state *sv, void *host, size_t n_amps) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && host, "null argument");
    QSV_CHECK(n_amps <= sv->length(), "host buffer larger than the local shard");
    // A purely local copy (as CopyGpuDataToHost of the reference) whenever the layout is canonical -- always, unless the
    // register runs with a lazy map (qsv_dist_set_lazy_map), where this call is collective and documented as such.
    bool canonical = true;
    for (int b = 0; b < sv->dist->n_total; ++b) canonical = canonical && sv->dist->phys_of[b] == b;
    if (!canonical) {
        QSV_CHECK(sv->dist->lazy_map, "internal: non-canonical layout outside lazy-map mode");
        dist_canonicalize(*sv, 0);
    }
    copy_state_host(*sv, sv->data, host, n_amps * sv->amp_bytes(), false);
    QSV_API_END
}

/* updateData(other) of StateVectorCudaMPI: shard and qubit map */
int qsv_dist_copy(qsv_state *dst, const qsv_state *src) {
    QSV_API_BEGIN
    QSV_CHECK(dst && src && dst->dist && src->dist, "null argument");
    QSV_CHECK(dst->n == src->n && dst->dtype == src->dtype && dst->dist->n_total == src->dist->n_total,
              "sharded registers differ in size or precision");
    dst->use();
    if (src->stream != dst->stream) QSV_CUDA(cudaStreamSynchronize(src->stream));
    QSV_CUDA(cudaMemcpyAsync(dst->data, src->data, dst->bytes(), cudaMemcpyDefault, dst->stream));
    QSV_CUDA(cudaStreamSynchronize(dst->stream));
    dst->dist->phys_of = src->dist->phys_of;
    dst->dist->log_of = src->dist->log_of;
    QSV_API_END
}

int qsv_dist_barrier(qsv_state *sv) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    sv->use();
    stream_barrier(*sv);
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    QSV_API_END
}

/* MPI_Bcast of a small host buffer (MPIManager::Bcast, util/MPIManager.hpp) */
int qsv_dist_bcast_bytes(qsv_state *sv, void *host, size_t bytes, int root) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && host, "null argument");
    sv->use();
    DistCtx &d = *sv->dist;
    void *dev = nullptr;
    QSV_CUDA(cudaMalloc(&dev, std::max<size_t>(bytes, 16)));
    if (d.rank == root) QSV_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, sv->stream));
    QSV_NCCL(ncclBroadcast(dev, dev, bytes, ncclChar, root, d.comm, sv->stream));
    QSV_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, sv->stream));
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    QSV_CUDA(cudaFree(dev));
    QSV_API_END
}

/* MPI_Scatter of a host buffer held by `root` (MPIManager::Scatter, bindings/Bindings.cpp:905-928): rank r receives
 * bytes [r * bytes_per_rank, (r + 1) * bytes_per_rank), staged through device memory in <= 64 MiB pieces */
int qsv_dist_scatter_host(qsv_state *sv, const void *send_host, void *recv_host, size_t bytes_per_rank, int root) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr && recv_host, "null argument");
    sv->use();
    DistCtx &d = *sv->dist;
    QSV_CHECK(d.rank != root || send_host != nullptr, "null send buffer on the root rank");
    const size_t piece = (size_t)64 << 20;
    void *dev = nullptr;
    QSV_CUDA(cudaMalloc(&dev, std::min(piece, std::max<size_t>(bytes_per_rank, 16))));
    for (size_t off = 0; off < bytes_per_rank; off += piece) {
        const size_t nb = std::min(piece, bytes_per_rank - off);
        if (d.rank == root) {
            for (int r = 0; r < d.world; ++r) {
                const char *src = (const char *)send_host + (size_t)r * bytes_per_rank + off;
                if (r == root) {
                    memcpy((char *)recv_host + off, src, nb);
                    continue;
                }
                QSV_CUDA(cudaMemcpyAsync(dev, src, nb, cudaMemcpyHostToDevice, sv->stream));
                QSV_NCCL(ncclSend(dev, nb, ncclChar, r, d.comm, sv->stream));
                QSV_CUDA(cudaStreamSynchronize(sv->stream));
            }
        } else {
            QSV_NCCL(ncclRecv(dev, nb, ncclChar, root, d.comm, sv->stream));
            QSV_CUDA(cudaMemcpyAsync((char *)recv_host + off, dev, nb, cudaMemcpyDeviceToHost, sv->stream));
            QSV_CUDA(cudaStreamSynchronize(sv->stream));
        }
    }
    QSV_CUDA(cudaFree(dev));
    QSV_API_END
}

int qsv_dist_nccl_version(int *version) {
    QSV_API_BEGIN
    QSV_CHECK(version != nullptr, "null argument");
    QSV_NCCL(ncclGetVersion(version));
    QSV_API_END
}


int qsv_dist_last_swap_stats(const qsv_state *sv, uint64_t *bytes_sent, float *ms) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    sv->use();
    harvest_swaps(*sv->dist, true);  // waits for the recorded exchanges; the hot path never does
    if (bytes_sent) *bytes_sent = sv->dist->last_bytes;
    if (ms) *ms = sv->dist->last_ms;
    QSV_API_END
}

int qsv_dist_uses_peer_access(const qsv_state *sv) { return sv && sv->dist && sv->dist->p2p ? 1 : 0; }

int qsv_dist_fused_exchange_stats(const qsv_state *sv, int *n_out_of_place, int *n_carried_by_sweeps) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    if (n_out_of_place) *n_out_of_place = sv->dist->n_oop;
    if (n_carried_by_sweeps) *n_carried_by_sweeps = sv->dist->n_fused;
    QSV_API_END
}

int qsv_dist_split_exchange_stats(const qsv_state *sv, int *n_second_halves, int *n_carried_by_sweeps) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    if (n_second_halves) *n_second_halves = sv->dist->n_pull;
    if (n_carried_by_sweeps) *n_carried_by_sweeps = sv->dist->n_pull_carried;
    QSV_API_END
}

int qsv_dist_total_swap_stats(qsv_state *sv, int *n_swaps, uint64_t *bytes_sent, float *ms, int reset) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv->dist;
    sv->use();
    harvest_swaps(d, true);
    if (n_swaps) *n_swaps = d.n_swaps;
    if (bytes_sent) *bytes_sent = d.total_bytes;
    if (ms) *ms = d.total_ms;
    if (reset) {
        d.n_swaps = 0;
        d.total_bytes = 0;
        d.total_ms = 0.f;
    }
    QSV_API_END
}

/* Host-only planning entry (no GPU, no NCCL): which exchanges a circuit triggers for a register of
 * n_total qubits sharded over 2^(n_total - n_local) ranks.  steps receives triples
 * (kind, a, b): kind 0 = swap of physical global bit a with local bit b, kind 1 = gate number a.
 * Used by the gloo (CPU) tests of the N > 1 path. */
int qsv_dist_plan(const qsv_ops *ops, int n_total, int n_local, int *steps, int max_steps, int *n_steps,
                  int *final_phys_of_logical_bit) {
    return qsv_dist_plan_from(ops, n_total, n_local, nullptr, steps, max_steps, n_steps, final_phys_of_logical_bit);
}

int qsv_dist_plan_from(const qsv_ops *ops, int n_total, int n_local, const int *initial_phys_of_logical_bit, int *steps,
                       int max_steps, int *n_steps, int *final_phys_of_logical_bit) {
    QSV_API_BEGIN
    QSV_CHECK(ops != nullptr && n_steps != nullptr, "null argument");
    QSV_CHECK(n_local >= 1 && n_local <= n_total && n_total <= 48, "invalid register sizes");
    std::vector<int> phys_of(n_total), log_of(n_total, -1);
    for (int b = 0; b < n_total; ++b) {
        phys_of[b] = initial_phys_of_logical_bit ? initial_phys_of_logical_bit[b] : b;
        QSV_CHECK(phys_of[b] >= 0 && phys_of[b] < n_total && log_of[phys_of[b]] < 0, "the initial qubit map is not a permutation");
        log_of[phys_of[b]] = b;
    }
    std::vector<uint64_t> dense, diag;
    std::vector<int> op_index;
    std::vector<LoweredGate> lowered;
    for (size_t i = 0; i < ops->ops.size(); ++i) {
        const Op &op = ops->ops[i];
        if (op.name == "Identity") continue;
        LoweredGate g = lower_op_total(n_total, op, false);
        uint64_t a = 0, b = 0;
        gate_bit_masks(g, a, b);
        QSV_CHECK(__builtin_popcountll(a) <= n_local, "gate acts on more wires than a shard holds");
        dense.push_back(a);
        diag.push_back(b);
        op_index.push_back((int)i);
        lowered.push_back(std::move(g));
    }
    int ns = 0;
    // the schedule qsv_dist_apply_ops executes for a fused complex128 circuit with room for the second buffer
    for (const DistStep &st : plan_dist_steps_priced(lowered, dense, diag, phys_of, log_of, n_local, QSV_C128, true)) {
        if (steps && ns < max_steps) {
            steps[3 * ns] = st.kind;
            steps[3 * ns + 1] = st.kind == 0 ? st.a : op_index[st.a];
            steps[3 * ns + 2] = st.kind == 0 ? st.b : 0;
        }
        ++ns;
    }
    *n_steps = ns;
    if (final_phys_of_logical_bit)
        for (int b = 0; b < n_total; ++b) final_phys_of_logical_bit[b] = phys_of[b];
    QSV_API_END
}

}  // extern "C"
