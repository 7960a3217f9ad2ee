// Internal declarations shared by the translation units of libqsv_b200.so.
// Nothing in here is part of the C ABI (see include/qsv_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <complex>
#include <cstdint>
#include <cstring>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <array>
#include <vector>

#include "../../include/qsv_b200.h"

namespace qsv {

using cplx = std::complex<double>;

struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

[[noreturn]] void fail(const std::string &msg);
void set_last_error(const std::string &msg);

#define QSV_CUDA(expr)                                                                             \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            ::qsv::fail(std::string("CUDA error: ") + cudaGetErrorString(e__) + " in " #expr " (" + \
                        __FILE__ + ":" + std::to_string(__LINE__) + ")");                          \
    } while (0)

#define QSV_CHECK(cond, msg)                                                                       \
    do {                                                                                           \
        if (!(cond)) ::qsv::fail(msg);                                                             \
    } while (0)

// ------------------------------------------------------------------------------------------------
// Lowered gate: what the kernels consume.  Bits are amplitude-index bit positions.
// ------------------------------------------------------------------------------------------------
struct LoweredGate {
    enum Kind { NOP, DENSE, DIAG, PARITY } kind = NOP;
    // DENSE : 2^k x 2^k row-major matrix acting on the 2^k amplitudes base + offs[j];
    //         `holes` are all index bits fixed by the group (targets and controls), ascending.
    // DIAG  : amp *= tab[t], t = bits of the index at tgt_bits (tgt_bits[0] = MSB of t)
    // PARITY: amp *= tab[popc(index & zmask) & 1]
    // All kinds: only indices with every bit of ctrl_mask set are touched.
    int k = 0;                     // log2(#amplitudes per group) for DENSE, #table bits for DIAG
    std::vector<int> holes;        // DENSE
    std::vector<uint64_t> offs;    // DENSE, 2^k entries
    std::vector<int> tgt_bits;     // DIAG (MSB first), DENSE (for the tile executor; MSB first or empty)
    uint64_t ctrl_mask = 0;
    uint64_t zmask = 0;            // PARITY
    std::vector<cplx> mat;         // DENSE: 4^k, DIAG: 2^k, PARITY: 2
    int n_ctrl() const { return __builtin_popcountll(ctrl_mask); }
};

// Named-gate table (gate definitions follow simulator/cuGates_host.hpp and the control/target split
// of StateVectorCudaManaged.hpp:321-560; see gates.cu).
struct GateInfo {
    const char *name;
    int n_wires;   // 0 = variadic (MultiRZ)
    int n_params;
};
const GateInfo *find_gate(const std::string &name);
// full matrix (controls included) of a named gate in wire order; used by observables and tests
std::vector<cplx> named_gate_matrix(const std::string &name, const std::vector<double> &params,
                                    int n_wires);
LoweredGate lower_named(int n_qubits, const std::string &name, const std::vector<int> &wires,
                        const std::vector<double> &params, bool adjoint);
LoweredGate lower_matrix(int n_qubits, const cplx *matrix, const std::vector<int> &ctrl_wires,
                         const std::vector<int> &tgt_wires, bool adjoint);
// building blocks on index bits (tgt_bits[0] = most significant matrix / table bit)
LoweredGate make_dense_gate(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, std::vector<cplx> matrix);
LoweredGate make_diag_gate(const std::vector<int> &tgt_bits, uint64_t ctrl_mask, std::vector<cplx> diag);

// Generator of a parametric gate as an operator (GateGenerators.hpp); `scale` per
// AdjointDiffGPU.hpp:96-114.  `extra_identity` is the coefficient d of an additional d * Identity
// term (SingleExcitationMinus/Plus, DoubleExcitationMinus/Plus) so that the remaining part stays
// sparse:  G = extra_identity * I + op.
struct LoweredGenerator {
    LoweredGate op;             // non-unitary operator in LoweredGate form (projectors = ctrl_mask)
    double extra_identity = 0;  // G = op + extra_identity * 1
    double scale = -0.5;
};
LoweredGenerator lower_generator(int n_qubits, const std::string &name, const std::vector<int> &wires);

// ------------------------------------------------------------------------------------------------
// State
// ------------------------------------------------------------------------------------------------
struct DistCtx;  // dist.cu

struct State {
    int n = 0;
    int dtype = QSV_C128;
    int device = 0;
    cudaStream_t stream = nullptr;
    void *data = nullptr;
    bool owns = false;
    bool swap_ok = false;  // borrowed from the workspace cache: may be exchanged for another block of the same size
    // reduction scratch (device) + pinned host mirror, lazily allocated
    double *red_dev = nullptr;
    double *red_host = nullptr;
    size_t red_cap = 0;  // doubles
    // generic device scratch (matrices of big gates, pointer tables, ...)
    void *scratch = nullptr;
    size_t scratch_cap = 0;
    int64_t stat_launches = 0, stat_sweeps = 0;
    // sharded states: value of the index bits above n (rank << n); controls / diagonal gates on those
    // "global" bits are evaluated against it and need no communication
    uint64_t index_hi = 0;
    DistCtx *dist = nullptr;

    size_t amp_bytes() const { return dtype == QSV_C128 ? 16 : 8; }
    uint64_t length() const { return 1ull << n; }
    size_t bytes() const { return length() * amp_bytes(); }
    void use() const { QSV_CUDA(cudaSetDevice(device)); }
    double *reduction_buffer(size_t n_doubles);
    void *scratch_buffer(size_t bytes);
    ~State();
};

struct Op {
    std::string name;
    std::vector<int> wires;
    std::vector<double> params;
    bool inverse = false;
    std::vector<cplx> matrix;
};
struct Ops {
    std::vector<Op> ops;
};

struct Obs {
    enum Kind { NAMED, HERMITIAN, TENSOR, HAMILTONIAN, SPARSE } kind = NAMED;
    std::string name;
    std::vector<int> wires;
    std::vector<double> params;
    std::vector<cplx> matrix;
    std::vector<double> coeffs;
    std::vector<std::shared_ptr<Obs>> children;
    // SPARSE
    std::vector<int64_t> indptr, indices;
    std::vector<cplx> values;
    // device copy of the CSR arrays, made on first use and kept (circuit.cu: csr_on_device)
    mutable std::shared_ptr<void> dev_cache;
    mutable int dev_cache_device = -1;
};

// Layered reverse sweep of the adjoint method (circuit.cu), parameterised over how generators are evaluated and gates
// applied: single-GPU vectors (circuit.cu) or a sharded register with its companions (dist.cu).  Gates and generators
// are in the hooks' own qubit numbering (local index bits / logical bits of the whole register).
struct ReverseSweepHooks {
    int n_wires = 0;
    size_t n_obs = 0;
    std::function<LoweredGenerator(const Op &)> generator;
    std::function<LoweredGate(const Op &)> dagger;
    // <bra_i| gens[k] |lambda> accumulated into complex slot slot0[k] + 2 * i for every observable i
    std::function<void(const std::vector<LoweredGate> &gens, const std::vector<int> &slot0)> inner_products;
    // <bra_i|lambda> into the identity slot of trainable parameter tp (slot (tp * n_obs + i) * 2 + 1)
    std::function<void(int64_t tp)> identity_inner_product;
    std::function<void(const std::vector<LoweredGate> &batch)> apply;  // to lambda and every bra
};
void layered_reverse_sweep(const Ops &ops, const std::vector<int64_t> &trainable, std::vector<double> &factor,
                           std::vector<double> &extra, const ReverseSweepHooks &hooks);

// Whole-state host <-> device copy (state_io.cu): pageable host memory goes through multi-threaded pinned staging
void copy_state_host(State &sv, void *dev, void *host, size_t bytes, bool to_device);

// Per-device cache of released workspace blocks (state.cu): state-sized temporaries without cudaMalloc / cudaFree per call.
void *ws_acquire(int device, size_t bytes, cudaStream_t stream);
void ws_release(int device, void *p, size_t bytes, cudaStream_t stream);
void ws_trim(int device);

// ------------------------------------------------------------------------------------------------
// Kernel launchers (apply_kernels.cu / measure_kernels.cu / tile_kernels.cu).
// `vecs` = device pointers of the vectors the op is applied to (all same n/dtype, same stream).
// ------------------------------------------------------------------------------------------------
void launch_gate(State &sv, const LoweredGate &g);
void launch_gate_multi(State &sv, const LoweredGate &g, void *const *vecs, int n_vecs);

void launch_fill_basis(State &sv, uint64_t index);
void launch_scatter(State &sv, const int64_t *dev_idx, const void *dev_vals, size_t count);
void launch_axpy(State &sv, cplx alpha, const void *x, void *y);

// <bra| op |ket> for a LoweredGate used as an operator (DENSE / DIAG / PARITY / NOP = identity);
// results (re, im) are accumulated as doubles into out_dev[2*slot..] asynchronously.
void launch_bra_op_ket(State &sv, const void *bra, const void *ket, const LoweredGate &op,
                       double *out_dev, int slot);
// out_dev[2 * slots[k] ..] += <bra| gens[k] |ket> for many generators in as few reads of (bra, ket) as possible
// (adjoint_kernels.cu)
void launch_bra_gens_ket(State &sv, const void *bra, const void *ket, const std::vector<LoweredGate> &gens,
                         const std::vector<int> &slots, double *out_dev);
// many Pauli words at once: out_dev[2 * (first_slot + t) ..] += <bra| P_t |ket>
void launch_bra_paulis_ket(State &sv, const void *bra, const void *ket, int n_terms, const uint64_t *xmasks,
                           const uint64_t *zmasks, const int *nys, int first_slot, double *out_dev);
// <bra| P |ket> for a Pauli word given by masks; result *(i^ny) applied on device
void launch_bra_pauli_ket(State &sv, const void *bra, const void *ket, uint64_t xmask, uint64_t zmask,
                          int ny, double *out_dev, int slot);
// out[i] = sum_t coeff_t (P_t in)[i]   (out-of-place, in != out)
bool launch_pauli_sum_apply_tiled(State &sv, const void *in, void *out, int n_terms, const uint64_t *xmasks,
                                  const uint64_t *zmasks, const cplx *coeffs, bool accumulate);
void launch_pauli_sum_apply(State &sv, const void *in, void *out, int n_terms, const uint64_t *xmasks,
                            const uint64_t *zmasks, const cplx *coeffs_with_phase, bool accumulate = false);
void launch_probs(State &sv, const std::vector<int> &bits_lsb_first, double *out_host);
void launch_sample(State &sv, const double *uniforms, int64_t shots, uint64_t *out_host);
// CSR: y = H x (y may be null) and/or accumulate <x|Hx> into out_dev[2*slot..]
void launch_csr(State &sv, const void *x, void *y, const void *dev_indptr, const void *dev_indices,
                const void *dev_values, int64_t n_rows, int64_t nnz, int index_bytes, double *out_dev,
                int slot);
// row block of a sharded CSR product: x is spread over 2^(n_total - n_local) shards whose device pointers are
// x_shards[r] (peer-mapped); this rank's rows are [row_base, row_base + n_rows); column c lives in shard
// c >> n_local at offset c & (2^n_local - 1).  indptr is local (n_rows + 1 entries starting at 0).
void launch_csr_sharded(State &sv, void *const *x_shards_dev, int n_local, const void *x_local, void *y,
                        const void *dev_indptr, const void *dev_indices, const void *dev_values, int64_t n_rows,
                        int64_t nnz, int index_bytes, double *out_dev, int slot);
// index of each sample only (no bit expansion); *total_out receives the probability mass of the vector
void launch_sample_indices(State &sv, const double *targets, int64_t shots, uint64_t *index_host, bool targets_are_mass);
double state_mass(State &sv);
// zero `count` doubles of the reduction buffer / read them back (one sync)
void reduction_zero(State &sv, double *dev, size_t count);
void reduction_read(State &sv, const double *dev, double *host, size_t count);

// circuits
void apply_op(State &sv, const Op &op, bool extra_adjoint);
void apply_ops_fused(State &sv, const std::vector<LoweredGate> &gates);
// A global<->local index-bit exchange of a sharded register, carried by the sweeps of the gate batches around it (dist.cu).
// FusedExchange: the LAST sweep of a batch stores out of place -- amplitudes whose bit `local_bit` equals `my_value` to
// out_mine, the others to out_peer (the partner's buffer, peer-mapped) with the bit flipped; with stash_bit >= 0 only the
// leaving amplitudes whose stash bit is clear are pushed, the others wait in out_mine for the partner's pull.  Index [1] of
// the targets applies when a pull earlier in the same batch has already moved the register to its other buffer.
// `before` runs right before the carrying sweep (the handshake with the partner); done = a sweep carried it.
struct FusedExchange {
    void *out_mine[2] = {nullptr, nullptr};
    void *out_peer[2] = {nullptr, nullptr};
    int local_bit = 0;
    int my_value = 0;
    int stash_bit = -1;
    std::function<void()> before;
    bool done = false;
    bool moved_before = false;  // out: a pull earlier in the batch moved the register, so targets [1] are the ones to use
};
// FusedPull: the second half of a split exchange -- the FIRST sweep of the next batch reads out of place (from sv.data, or
// from in_peer at the flipped offset for what the partner parked) and writes out_mine, where the register lives from then
// on (the callee sets sv.data).  When the first sweep cannot carry it, the callee runs the copy-pass form before the batch.
// `after` runs right after the pull (the handshake that tells the partner its parked amplitudes have been read).
struct FusedPull {
    const void *in_peer = nullptr;
    void *out_mine = nullptr;
    int local_bit = 0;
    int my_value = 0;
    int stash_bit = 0;
    std::function<void()> after;
    bool carried = false;  // statistics: a sweep did it (else the copy pass)
};
// same, on several vectors at once (dev_table = device array of n_vecs pointers, or null for sv.data)
void apply_gates_tiled(State &sv, const std::vector<LoweredGate> &gates, void *const *dev_table, int n_vecs,
                       FusedExchange *fx = nullptr, FusedPull *pull = nullptr);
void launch_xchg_push_copy(State &sv, const void *in, void *out_mine, void *out_peer, int local_bit, int my_value, int stash_bit);
void launch_xchg_pull_copy(State &sv, const void *in_mine, const void *in_peer, void *out, int local_bit, int my_value,
                           int stash_bit);
bool regs_pull_supported();
// register-blocked tile kernel (tile_regs.cu) and its sweep planner (tile_kernels.cu)
struct SweepPlan {
    std::vector<int> gates;  // indices into the (merged) gate list, in execution order
    uint64_t need = 0;       // dense-target bits >= L the tile must contain
    bool fused = false;      // false: a lone gate for the one-sweep-per-gate kernels
};
std::vector<LoweredGate> prepare_gates_regs(const std::vector<LoweredGate> &gates_in);
std::vector<SweepPlan> plan_sweeps_regs(int n_local, const std::vector<LoweredGate> &gates, int L, bool dag,
                                        int max_gates, int window, int dtype = QSV_C128, int tries = 0);
// the same through a small cache keyed by the structure of the gate list (kinds and index bits, not matrix entries)
std::vector<SweepPlan> plan_sweeps_cached(int n_local, const std::vector<LoweredGate> &merged, int L, bool dag, int max_gates,
                                          int window, int dtype);
// arithmetic of one fused sweep: fused multiply-adds per amplitude and register passes of the program built for it
void regs_sweep_work(int n, int dtype, const std::vector<const LoweredGate *> &gates, uint64_t need, int L, double *fma_per_amp,
                     int *passes);
// price of one fused sweep under the cost model of tools/sweep_cost_model.py (ms at 30 qubits complex128; only ratios matter)
double regs_sweep_model_cost(int n, int dtype, const std::vector<const LoweredGate *> &gates, uint64_t need, int L,
                             bool greedy_scheduler);
bool gates_commute_structurally(const LoweredGate &a, const LoweredGate &b);
bool regs_fusable(const LoweredGate &g, int n_local);
uint64_t regs_need_bits(const LoweredGate &g);
void run_sweep_regs(State &sv, const std::vector<const LoweredGate *> &gates, uint64_t need, int L,
                    void *const *table, int n_vecs, const FusedExchange *fx = nullptr, int fx_idx = 0,
                    const FusedPull *pull = nullptr);
bool regs_tile_contains_bit(int n, uint64_t need, int L, int bit);
void apply_observable(State &sv, const Obs &obs);          // sv <- O sv
double observable_expval(State &sv, const Obs &obs);       // Re <sv|O|sv>
void adjoint_jacobian(State &sv, const Ops &ops, const std::vector<const Obs *> &obs,
                      const std::vector<int64_t> &trainable, bool apply_operations, double *jac);

// hooks for the CPU emulation of the sharded executor (dist.cu; used by tests/native only)
LoweredGate dist_hook_lower(int n_total, const Op &op);
std::vector<std::array<int, 3>> dist_hook_plan(const std::vector<LoweredGate> &lowered, std::vector<int> &phys_of,
                                               std::vector<int> &log_of, int n_local);
LoweredGate dist_hook_localized(const LoweredGate &g, const std::vector<int> &phys_of, int n_local, uint64_t index_hi);

// Pauli-word views of observables (circuit.cu)
bool as_pauli_word(const Obs &o, int n, uint64_t &x, uint64_t &z, int &ny);
bool hamiltonian_of_pauli_words(const Obs &o, int n, std::vector<uint64_t> &xs, std::vector<uint64_t> &zs,
                                std::vector<cplx> &cf);

// dist.cu
void dist_free(State &sv);

}  // namespace qsv

// the opaque handles of include/qsv_b200.h
struct qsv_state : qsv::State {};
struct qsv_ops : qsv::Ops {};
struct qsv_obs {
    std::shared_ptr<qsv::Obs> p;
};
