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
#include <cstdlib>

#include "qsv_internal.h"

namespace qsv {

#define QSV_NCCL(expr)                                                                              \
    do {                                                                                            \
        ncclResult_t r__ = (expr);                                                                  \
        if (r__ != ncclSuccess)                                                                     \
            ::qsv::fail(std::string("NCCL error: ") + ncclGetErrorString(r__) + " in " #expr);      \
    } while (0)

struct DistCtx {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, n_total = 0, n_global = 0;
    std::vector<int> phys_of;  // logical bit -> physical bit
    std::vector<int> log_of;   // physical bit -> logical bit
    cudaStream_t comm_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_ready = nullptr, ev_xfer[2] = {nullptr, nullptr}, ev_copy[2] = {nullptr, nullptr};
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
    void *stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    uint64_t last_bytes = 0, total_bytes = 0;
    float last_ms = 0.f, total_ms = 0.f;
    int n_swaps = 0;
    double *red_dev = nullptr;
    // direct peer access (CUDA IPC): the partner's shard mapped into this process
    bool p2p = false;
    void *registered = nullptr;            // sv.data at registration time
    std::vector<void *> peer;              // peer[r] = rank r's shard, mapped here (null for r == rank)
    std::vector<void *> peer_map_base;     // what cudaIpcOpenMemHandle returned (for closing)
    int *hs_dev = nullptr;                 // 2 ints for the handshakes
};

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

void handshake(State &sv, int peer) {
    DistCtx &d = *sv.dist;
    QSV_NCCL(ncclGroupStart());
    QSV_NCCL(ncclSend(d.hs_dev, 1, ncclInt, peer, d.comm, sv.stream));
    QSV_NCCL(ncclRecv(d.hs_dev + 1, 1, ncclInt, peer, d.comm, sv.stream));
    QSV_NCCL(ncclGroupEnd());
}

void swap_p2p(State &sv, int gphys, int l) {
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    const int gb = gphys - n_local;
    const int peer = d.rank ^ (1 << gb);
    const uint64_t a = (d.rank >> gb) & 1;
    QSV_CHECK(sv.data == d.registered, "the shard was re-allocated after qsv_dist_init; peer mappings are stale");
    // 16-byte units: complex128 = 1 unit, complex64 = half a unit (two amplitudes per unit, so l >= 1)
    const int shift = sv.dtype == QSV_C128 ? 0 : 1;
    QSV_CHECK(l >= shift, "internal: cannot swap index bit 0 of a complex64 shard through 16-byte units");
    const int l_vec = l - shift;
    const uint64_t pairs = (sv.length() >> shift) >> 1;  // 16-byte units in the exchanged half
    const uint64_t half = pairs / 2;
    const uint64_t first = a == 0 ? 0 : half;
    const uint64_t count = a == 0 ? half : pairs - half;
    QSV_CUDA(cudaEventRecord(d.ev_t0, sv.stream));
    handshake(sv, peer);  // the partner has finished everything queued before its own handshake
    if (count > 0) {
        constexpr int U = 4;
        const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((count + 256 * U - 1) / (256 * U), 148 * 16));
        k_peer_swap<U><<<grid, 256, 0, sv.stream>>>((uint4 *)sv.data, (uint4 *)d.peer[peer], first, count, l_vec,
                                                    (a ^ 1) << l_vec, a << l_vec);
        QSV_CUDA(cudaGetLastError());
    }
    handshake(sv, peer);  // the partner's kernel has finished writing into this shard
    QSV_CUDA(cudaEventRecord(d.ev_t1, sv.stream));
}

}  // namespace

// physical swap of global bit gphys (>= n_local) with local bit l
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
    if (d.p2p && !(sv.dtype == QSV_C64 && l == 0)) {
        swap_p2p(sv, gphys, l);
        QSV_CUDA(cudaEventSynchronize(d.ev_t1));
        float ms = 0.f;
        QSV_CUDA(cudaEventElapsedTime(&ms, d.ev_t0, d.ev_t1));
        d.last_ms = ms;
        d.last_bytes = (uint64_t)(sv.length() / 2) * ab;
        d.total_ms += ms;
        d.total_bytes += d.last_bytes;
        d.n_swaps += 1;
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
    QSV_CUDA(cudaEventRecord(d.ev_t0, d.comm_stream));
    uint64_t counter = 0;
    char *base = (char *)sv.data;
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
    QSV_CUDA(cudaEventRecord(d.ev_t1, d.comm_stream));
    QSV_CUDA(cudaStreamWaitEvent(sv.stream, d.ev_t1, 0));
    for (int k = 0; k < 2 && (uint64_t)k < counter; ++k) QSV_CUDA(cudaStreamWaitEvent(sv.stream, d.ev_copy[k], 0));
    // statistics (host sync only here, once per swap: the transfer is tens of milliseconds)
    QSV_CUDA(cudaEventSynchronize(d.ev_t1));
    float ms = 0.f;
    QSV_CUDA(cudaEventElapsedTime(&ms, d.ev_t0, d.ev_t1));
    d.last_ms = ms;
    d.last_bytes = (uint64_t)(sv.length() / 2) * ab;
    d.total_ms += ms;
    d.total_bytes += d.last_bytes;
    d.n_swaps += 1;
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

void dist_apply_ops(State &sv, const Ops &ops, bool fuse, size_t chunk_bytes) {
    sv.use();
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv.dist;
    const int n_local = sv.n, n_total = d.n_total;
    sv.stat_launches = 0;
    sv.stat_sweeps = 0;
    std::vector<LoweredGate> lowered;
    lowered.reserve(ops.ops.size());
    for (const auto &op : ops.ops) {
        if (op.name == "Identity") continue;
        lowered.push_back(lower_op_total(n_total, op, false));
    }
    std::vector<uint64_t> need(lowered.size(), 0);  // logical bits that must be local for gate i
    for (size_t i = 0; i < lowered.size(); ++i)
        if (lowered[i].kind == LoweredGate::DENSE) need[i] = touched_mask(lowered[i]);
    for (uint64_t m : need)
        QSV_CHECK(__builtin_popcountll(m) <= n_local, "gate acts on more wires than a shard holds");

    std::vector<LoweredGate> batch;
    auto flush = [&]() {
        if (batch.empty()) return;
        if (fuse) {
            apply_gates_tiled(sv, batch, nullptr, 1);
        } else {
            for (const auto &g : batch) launch_gate(sv, g);
        }
        batch.clear();
    };
    for (size_t i = 0; i < lowered.size(); ++i) {
        // bring the dense-target qubits of gate i into the local shard
        for (int lb = 0; lb < n_total; ++lb) {
            if (!(need[i] >> lb & 1) || d.phys_of[lb] < n_local) continue;
            flush();
            const int best = pick_victim(need, i, d.log_of, n_local);
            swap_logical_in(sv, d.phys_of[lb], best, chunk_bytes);
        }
        LoweredGate g = localize_gate(remap_gate(lowered[i], d.phys_of), n_local, sv.index_hi);
        if (g.kind != LoweredGate::NOP) batch.push_back(std::move(g));
    }
    flush();
}

// restore the identity qubit map (physical bit b holds logical bit b)
void dist_canonicalize(State &sv, size_t chunk_bytes) {
    QSV_CHECK(sv.dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv.dist;
    const int n_local = sv.n;
    // local-local permutations are done with SWAP gates, global ones with exchanges
    for (int gp = n_local; gp < d.n_total; ++gp) {
        if (d.log_of[gp] == gp) continue;
        // logical qubit gp lives somewhere else: if it is global, first bring it local
        int p = d.phys_of[gp];
        if (p >= n_local) {
            // find a local slot holding a qubit that is not a global-home qubit if possible
            int l = n_local - 1;
            swap_logical_in(sv, p, l, chunk_bytes);
            p = d.phys_of[gp];
        }
        swap_logical_in(sv, gp, p, chunk_bytes);
    }
    // now all global positions are right; fix the local part with local SWAPs
    for (int b = 0; b < n_local; ++b) {
        while (d.log_of[b] != b) {
            const int other = d.phys_of[b];  // where logical b currently sits (local)
            // SWAP physical bits b and other
            LoweredGate g;
            g.kind = LoweredGate::DENSE;
            g.k = 1;
            g.holes = {std::min(b, other), std::max(b, other)};
            g.offs = {1ull << b, 1ull << other};
            g.mat = {0.0, 1.0, 1.0, 0.0};
            launch_gate(sv, g);
            const int qa = d.log_of[b], qb = d.log_of[other];
            d.log_of[b] = qb;
            d.log_of[other] = qa;
            d.phys_of[qa] = other;
            d.phys_of[qb] = b;
        }
    }
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
    double *red = sv.reduction_buffer(2);
    for (int t = 0; t < n_terms; ++t) {
        // X/Y letters on global qubits need the partner amplitudes: make those qubits local first
        for (int lb = 0; lb < d.n_total; ++lb) {
            if (!(x_log[t] >> lb & 1) || d.phys_of[lb] < n_local) continue;
            int best = -1;
            for (int l = n_local - 1; l >= 0; --l)
                if (!(x_log[t] >> d.log_of[l] & 1)) {
                    best = l;
                    break;
                }
            QSV_CHECK(best >= 0, "Pauli word flips more qubits than a shard holds");
            swap_logical_in(sv, d.phys_of[lb], best, chunk_bytes);
        }
        const uint64_t x = remap_mask(x_log[t], d.phys_of), z = remap_mask(z_log[t], d.phys_of);
        reduction_zero(sv, red, 2);
        launch_bra_pauli_ket(sv, sv.data, sv.data, x, z & ~hi_mask, ny[t], red, 0);
        double h[2];
        reduction_read(sv, red, h, 2);
        const double sgn = (__builtin_popcountll(sv.index_hi & z & hi_mask) & 1) ? -1.0 : 1.0;
        vals[2 * t] = sgn * h[0];
        vals[2 * t + 1] = sgn * h[1];
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
    if (d->ev_t0) cudaEventDestroy(d->ev_t0);
    if (d->ev_t1) cudaEventDestroy(d->ev_t1);
    if (d->comm_stream) cudaStreamDestroy(d->comm_stream);
    if (d->copy_stream) cudaStreamDestroy(d->copy_stream);
    if (d->red_dev) cudaFree(d->red_dev);
    if (d->hs_dev) cudaFree(d->hs_dev);
    for (void *p : d->peer_map_base)
        if (p) cudaIpcCloseMemHandle(p);
    delete d;
    sv.dist = nullptr;
    sv.index_hi = 0;
}

}  // namespace qsv

namespace qsv {

// Map every other rank's shard into this process with CUDA IPC.  The handles travel through NCCL itself
// (all-gather of 80 bytes per rank), so no other rendezvous is needed.  Falls back to the NCCL
// send/recv exchange (all ranks together) when any mapping fails.
void setup_peer_access(State &sv) {
    DistCtx &d = *sv.dist;
    struct Msg {
        cudaIpcMemHandle_t handle;
        uint64_t offset;
        uint64_t ok;
    };
    static_assert(sizeof(Msg) == 80, "IPC message layout");
    Msg mine;
    memset(&mine, 0, sizeof(mine));
    // base of the allocation that contains the shard (the shard may be a view into a torch block)
    typedef int (*GetRangeFn)(unsigned long long *, size_t *, unsigned long long);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    unsigned long long base = 0;
    size_t size = 0;
    bool ok = cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn != nullptr &&
              ((GetRangeFn)fn)(&base, &size, (unsigned long long)(uintptr_t)sv.data) == 0;
    if (ok) ok = cudaIpcGetMemHandle(&mine.handle, (void *)(uintptr_t)base) == cudaSuccess;
    cudaGetLastError();
    mine.offset = ok ? (uint64_t)((uintptr_t)sv.data - base) : 0;
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
    d.peer.assign(d.world, nullptr);
    d.peer_map_base.assign(d.world, nullptr);
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
            d.peer_map_base[r] = p;
            d.peer[r] = (char *)p + all[r].offset;
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
    d.p2p = mapped_ok == 1;
    d.registered = sv.data;
}

}  // namespace qsv

using namespace qsv;
struct qsv_state : State {};
struct qsv_ops : Ops {};

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
    QSV_CUDA(cudaEventCreate(&d->ev_t0));
    QSV_CUDA(cudaEventCreate(&d->ev_t1));
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
    QSV_API_END
}

int qsv_dist_canonicalize(qsv_state *sv, size_t chunk_bytes) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr, "null state");
    dist_canonicalize(*sv, chunk_bytes);
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
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
    QSV_API_END
}

int qsv_dist_last_swap_stats(const qsv_state *sv, uint64_t *bytes_sent, float *ms) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    if (bytes_sent) *bytes_sent = sv->dist->last_bytes;
    if (ms) *ms = sv->dist->last_ms;
    QSV_API_END
}

int qsv_dist_uses_peer_access(const qsv_state *sv) { return sv && sv->dist && sv->dist->p2p ? 1 : 0; }

int qsv_dist_total_swap_stats(qsv_state *sv, int *n_swaps, uint64_t *bytes_sent, float *ms, int reset) {
    QSV_API_BEGIN
    QSV_CHECK(sv != nullptr && sv->dist != nullptr, "state vector is not part of a distributed register");
    DistCtx &d = *sv->dist;
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
    QSV_API_BEGIN
    QSV_CHECK(ops != nullptr && n_steps != nullptr, "null argument");
    QSV_CHECK(n_local >= 1 && n_local <= n_total && n_total <= 48, "invalid register sizes");
    std::vector<int> phys_of(n_total), log_of(n_total);
    for (int b = 0; b < n_total; ++b) phys_of[b] = log_of[b] = b;
    std::vector<uint64_t> need;
    std::vector<int> op_index;
    for (size_t i = 0; i < ops->ops.size(); ++i) {
        const Op &op = ops->ops[i];
        if (op.name == "Identity") continue;
        LoweredGate g = lower_op_total(n_total, op, false);
        need.push_back(g.kind == LoweredGate::DENSE ? touched_mask(g) : 0);
        op_index.push_back((int)i);
    }
    int ns = 0;
    auto emit = [&](int kind, int a, int b) {
        if (steps && ns < max_steps) {
            steps[3 * ns] = kind;
            steps[3 * ns + 1] = a;
            steps[3 * ns + 2] = b;
        }
        ++ns;
    };
    for (size_t i = 0; i < need.size(); ++i) {
        for (int lb = 0; lb < n_total; ++lb) {
            if (!(need[i] >> lb & 1) || phys_of[lb] < n_local) continue;
            const int best = pick_victim(need, i, log_of, n_local);
            const int gp = phys_of[lb];
            emit(0, gp, best);
            const int a = log_of[gp], b = log_of[best];
            log_of[gp] = b;
            log_of[best] = a;
            phys_of[a] = best;
            phys_of[b] = gp;
        }
        emit(1, op_index[i], 0);
    }
    *n_steps = ns;
    if (final_phys_of_logical_bit)
        for (int b = 0; b < n_total; ++b) final_phys_of_logical_bit[b] = phys_of[b];
    QSV_API_END
}

}  // extern "C"
