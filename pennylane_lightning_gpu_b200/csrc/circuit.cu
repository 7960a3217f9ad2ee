// Host-side orchestration above the kernels: op dispatch, observables, adjoint Jacobian.
//
// Mirrors (behaviour, not code)
//   applyOperation                     simulator/StateVectorCudaManaged.hpp:198-247
//   ObservableGPU<T>::applyInPlace     algorithms/ObservablesGPU.hpp:56-587
//   AdjointJacobianGPU::adjointJacobian algorithms/AdjointDiffGPU.hpp:499-596, updateJacobian :132-162
// with these structural differences (all on the hot path, none visible in results):
//   * no mu vector: <H_lambda| G |lambda> is one fused read of both vectors (launch_bra_op_ket);
//   * U^dagger is applied to lambda and to all bras by ONE launch (gridDim.y = 1 + n_obs);
//   * Jacobian entries accumulate in a device buffer and are read back once, not once per
//     parameter (AdjointDiffGPU.hpp:153-155 syncs every parameter);
//   * a Hamiltonian of Pauli words is applied by one gather kernel, not by per-term
//     copy + apply + axpy (ObservablesGPU.hpp:346-362).
#include <algorithm>
#include <cstdlib>
#include <functional>

#include "qsv_internal.h"

namespace qsv {

// Pauli word view of an observable: Named X/Y/Z/Identity or a tensor product of those on
// distinct wires
bool as_pauli_word(const Obs &o, int n, uint64_t &x, uint64_t &z, int &ny) {
    if (o.kind == Obs::NAMED) {
        if (o.wires.size() != 1) return false;
        const int w = o.wires[0];
        QSV_CHECK(w >= 0 && w < n, "observable wire out of range");
        const uint64_t b = 1ull << (n - 1 - w);
        if ((x | z) & b) return false;  // same wire twice: not a plain word
        if (o.name == "PauliX") {
            x |= b;
        } else if (o.name == "PauliY") {
            x |= b;
            z |= b;
            ny += 1;
        } else if (o.name == "PauliZ") {
            z |= b;
        } else if (o.name != "Identity") {
            return false;
        }
        return true;
    }
    if (o.kind == Obs::TENSOR) {
        for (const auto &c : o.children)
            if (!as_pauli_word(*c, n, x, z, ny)) return false;
        return true;
    }
    return false;
}

bool hamiltonian_of_pauli_words(const Obs &o, int n, std::vector<uint64_t> &xs, std::vector<uint64_t> &zs,
                                std::vector<cplx> &cf) {
    if (o.kind != Obs::HAMILTONIAN) return false;
    for (size_t t = 0; t < o.children.size(); ++t) {
        uint64_t x = 0, z = 0;
        int ny = 0;
        if (!as_pauli_word(*o.children[t], n, x, z, ny)) return false;
        cplx ph(1.0, 0.0);
        for (int i = 0; i < (ny & 3); ++i) ph *= cplx(0.0, 1.0);
        xs.push_back(x);
        zs.push_back(z);
        cf.push_back(ph * o.coeffs[t]);
    }
    return true;
}

namespace {

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    int device = 0;
    cudaStream_t stream = nullptr;
    // small, short-lived buffers (pointer tables, CSR arrays): plain cudaMalloc
    explicit DevBuf(size_t n) : bytes(n) { QSV_CUDA(cudaMalloc(&p, n)); }
    // state-sized workspace: from the per-device cache
    DevBuf(const State &sv, size_t n) : bytes(n), device(sv.device), stream(sv.stream), cached(true) {
        p = ws_acquire(device, n, stream);
    }
    ~DevBuf() {
        if (!p) return;
        if (cached)
            ws_release(device, p, bytes, stream);
        else
            cudaFree(p);
    }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;

  private:
    bool cached = false;
};

LoweredGate lower_op(const State &sv, const Op &op, bool extra_adjoint) {
    const bool adj = op.inverse != extra_adjoint;
    if (find_gate(op.name) != nullptr) return lower_named(sv.n, op.name, op.wires, op.params, adj);
    if (op.matrix.empty()) fail("Currently unsupported gate: " + op.name);
    const size_t dim = 1ull << op.wires.size();
    QSV_CHECK(op.matrix.size() == dim * dim, "matrix of gate " + op.name + " does not match its wires");
    return lower_matrix(sv.n, op.matrix.data(), {}, op.wires, adj);
}

struct CsrDev {
    DevBuf indptr, indices, values;
    CsrDev(State &sv, const Obs &o)
        : indptr(o.indptr.size() * 8), indices(std::max<size_t>(o.indices.size(), 1) * 8),
          values(std::max<size_t>(o.values.size(), 1) * 16) {
        QSV_CUDA(cudaMemcpyAsync(indptr.p, o.indptr.data(), o.indptr.size() * 8, cudaMemcpyHostToDevice, sv.stream));
        QSV_CUDA(cudaMemcpyAsync(indices.p, o.indices.data(), o.indices.size() * 8, cudaMemcpyHostToDevice, sv.stream));
        QSV_CUDA(cudaMemcpyAsync(values.p, o.values.data(), o.values.size() * 16, cudaMemcpyHostToDevice, sv.stream));
        QSV_CUDA(cudaStreamSynchronize(sv.stream));  // the arrays may be used from other streams later
    }
};

// Device copy of a sparse observable, uploaded once per (observable, device) and kept for the observable's lifetime: a
// SparseHamiltonian is immutable once built, and its CSR arrays (3.1 GB at BASELINE config 4) dwarf the product itself.
// The reference uploads them on every call (simulator/StateVectorCudaManaged.hpp:829-832).
const CsrDev &csr_on_device(State &sv, const Obs &o) {
    if (!o.dev_cache || o.dev_cache_device != sv.device) {
        o.dev_cache.reset();
        o.dev_cache = std::shared_ptr<void>(new CsrDev(sv, o), [](void *p) { delete static_cast<CsrDev *>(p); });
        o.dev_cache_device = sv.device;
    }
    return *static_cast<const CsrDev *>(o.dev_cache.get());
}

void check_sparse_shape(const State &sv, const Obs &o) {
    QSV_CHECK(o.indptr.size() == sv.length() + 1, "sparse Hamiltonian dimension does not match the state vector");
}

// replace the contents of sv by those of tmp (same size)
void adopt(State &sv, DevBuf &tmp) {
    if (sv.owns || sv.swap_ok) {
        std::swap(sv.data, tmp.p);  // the old buffer is released by tmp's destructor
    } else {
        QSV_CUDA(cudaMemcpyAsync(sv.data, tmp.p, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
        QSV_CUDA(cudaStreamSynchronize(sv.stream));
    }
}

double obs_expval_generic(State &sv, const Obs &o) {
    // tmp = O sv ; <sv|tmp>
    DevBuf tmp(sv, sv.bytes());
    State t;
    t.n = sv.n;
    t.dtype = sv.dtype;
    t.device = sv.device;
    t.stream = sv.stream;
    t.data = tmp.p;
    t.owns = false;
    QSV_CUDA(cudaMemcpyAsync(t.data, sv.data, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
    apply_observable(t, o);
    double *red = sv.reduction_buffer(2);
    reduction_zero(sv, red, 2);
    LoweredGate id;
    launch_bra_op_ket(sv, sv.data, t.data, id, red, 0);
    double out[2];
    reduction_read(sv, red, out, 2);
    return out[0];
}

}  // namespace

void apply_op(State &sv, const Op &op, bool extra_adjoint) {
    if (op.name == "Identity") return;
    launch_gate(sv, lower_op(sv, op, extra_adjoint));
}

void apply_observable(State &sv, const Obs &o) {
    sv.use();
    switch (o.kind) {
    case Obs::NAMED: {
        Op op;
        op.name = o.name;
        op.wires = o.wires;
        op.params = o.params;
        op.matrix = o.matrix;
        apply_op(sv, op, false);
        return;
    }
    case Obs::HERMITIAN: {
        const size_t dim = 1ull << o.wires.size();
        QSV_CHECK(o.matrix.size() == dim * dim, "Hermitian matrix does not match its wires");
        launch_gate(sv, lower_matrix(sv.n, o.matrix.data(), {}, o.wires, false));
        return;
    }
    case Obs::TENSOR:
        for (const auto &c : o.children) apply_observable(sv, *c);
        return;
    case Obs::HAMILTONIAN: {
        std::vector<uint64_t> xs, zs;
        std::vector<cplx> cf;
        if (hamiltonian_of_pauli_words(o, sv.n, xs, zs, cf)) {
            DevBuf tmp(sv, sv.bytes());
            launch_pauli_sum_apply(sv, sv.data, tmp.p, (int)xs.size(), xs.data(), zs.data(), cf.data());
            adopt(sv, tmp);
            return;
        }
        // generic terms: acc = sum_t c_t O_t psi
        DevBuf acc(sv, sv.bytes()), tmp(sv, sv.bytes());
        QSV_CUDA(cudaMemsetAsync(acc.p, 0, sv.bytes(), sv.stream));
        State t;
        t.n = sv.n;
        t.dtype = sv.dtype;
        t.device = sv.device;
        t.stream = sv.stream;
        t.owns = false;
        for (size_t i = 0; i < o.children.size(); ++i) {
            t.data = tmp.p;
            QSV_CUDA(cudaMemcpyAsync(t.data, sv.data, sv.bytes(), cudaMemcpyDeviceToDevice, sv.stream));
            apply_observable(t, *o.children[i]);
            launch_axpy(sv, cplx(o.coeffs[i], 0.0), t.data, acc.p);
        }
        QSV_CUDA(cudaStreamSynchronize(sv.stream));
        adopt(sv, acc);
        return;
    }
    case Obs::SPARSE: {
        check_sparse_shape(sv, o);
        const CsrDev &csr = csr_on_device(sv, o);
        DevBuf y(sv, sv.bytes());
        launch_csr(sv, sv.data, y.p, csr.indptr.p, csr.indices.p, csr.values.p, (int64_t)sv.length(),
                   (int64_t)o.values.size(), 8, nullptr, 0);
        QSV_CUDA(cudaStreamSynchronize(sv.stream));
        adopt(sv, y);
        return;
    }
    }
}

double observable_expval(State &sv, const Obs &o) {
    sv.use();
    uint64_t x = 0, z = 0;
    int ny = 0;
    if (as_pauli_word(o, sv.n, x, z, ny)) {
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        launch_bra_pauli_ket(sv, sv.data, sv.data, x, z, ny, red, 0);
        double out[2];
        reduction_read(sv, red, out, 2);
        return out[0];
    }
    switch (o.kind) {
    case Obs::NAMED: {
        // any named gate used as an observable: its full matrix on its wires
        std::vector<cplx> m = o.matrix;
        if (find_gate(o.name) != nullptr) m = named_gate_matrix(o.name, o.params, (int)o.wires.size());
        QSV_CHECK(!m.empty(), "Currently unsupported observable: " + o.name);
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        launch_bra_op_ket(sv, sv.data, sv.data, lower_matrix(sv.n, m.data(), {}, o.wires, false), red, 0);
        double out[2];
        reduction_read(sv, red, out, 2);
        return out[0];
    }
    case Obs::HERMITIAN: {
        const size_t dim = 1ull << o.wires.size();
        QSV_CHECK(o.matrix.size() == dim * dim, "Hermitian matrix does not match its wires");
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        launch_bra_op_ket(sv, sv.data, sv.data, lower_matrix(sv.n, o.matrix.data(), {}, o.wires, false), red, 0);
        double out[2];
        reduction_read(sv, red, out, 2);
        return out[0];
    }
    case Obs::HAMILTONIAN: {
        std::vector<uint64_t> xs, zs;
        std::vector<cplx> cf;
        if (hamiltonian_of_pauli_words(o, sv.n, xs, zs, cf)) {
            const size_t T = xs.size();
            double *red = sv.reduction_buffer(2 * T);
            reduction_zero(sv, red, 2 * T);
            std::vector<int> ny0(T, 0);  // cf already carries i^ny
            launch_bra_paulis_ket(sv, sv.data, sv.data, (int)T, xs.data(), zs.data(), ny0.data(), 0, red);
            std::vector<double> out(2 * T);
            reduction_read(sv, red, out.data(), 2 * T);
            double tot = 0;
            for (size_t t = 0; t < T; ++t) tot += (cf[t] * cplx(out[2 * t], out[2 * t + 1])).real();
            return tot;
        }
        double tot = 0;
        for (size_t i = 0; i < o.children.size(); ++i) tot += o.coeffs[i] * observable_expval(sv, *o.children[i]);
        return tot;
    }
    case Obs::SPARSE: {
        check_sparse_shape(sv, o);
        const CsrDev &csr = csr_on_device(sv, o);
        double *red = sv.reduction_buffer(2);
        reduction_zero(sv, red, 2);
        launch_csr(sv, sv.data, nullptr, csr.indptr.p, csr.indices.p, csr.values.p, (int64_t)sv.length(),
                   (int64_t)o.values.size(), 8, red, 0);
        double out[2];
        reduction_read(sv, red, out, 2);
        return out[0];
    }
    case Obs::TENSOR:
        return obs_expval_generic(sv, o);
    }
    return 0.0;
}

// ------------------------------------------------------------------------------------------------
// Layered reverse sweep of the adjoint method (shared by the single-GPU path below and the sharded register, dist.cu)
// ------------------------------------------------------------------------------------------------
// Gates on disjoint wires commute, and so do their generators.  An op is READY when every later op that shares a wire
// with it has already been undone; all ready ops are pairwise disjoint, so the generator inner products of the
// trainable ones can all be taken against the current (bra, lambda) -- one batched launch per observable -- and their
// U^dagger, followed by every non-trainable op that becomes ready behind them (e.g. a whole CNOT ladder), go into one
// fused multi-gate sweep over lambda and all bras.  Reference: one op at a time, algorithms/AdjointDiffGPU.hpp:562-592.
void layered_reverse_sweep(const Ops &ops, const std::vector<int64_t> &trainable, std::vector<double> &factor,
                           std::vector<double> &extra, const ReverseSweepHooks &hk) {
    const size_t n_tp = trainable.size(), n_obs = hk.n_obs;
    const int64_t n_ops = (int64_t)ops.ops.size();
    std::vector<int64_t> tp_of(n_ops, -1);
    int64_t first_needed = n_ops;
    {
        int64_t cur = 0;
        size_t t = 0;
        for (int64_t idx = 0; idx < n_ops; ++idx) {
            if (ops.ops[idx].params.empty()) continue;
            while (t < n_tp && trainable[t] < cur) ++t;
            if (t < n_tp && trainable[t] == cur) {
                tp_of[idx] = (int64_t)t;
                first_needed = std::min(first_needed, idx);
            }
            ++cur;
        }
    }
    auto skipped = [&](const Op &op) {
        return op.name == "QubitStateVector" || op.name == "StatePrep" || op.name == "BasisState";
    };
    // per wire: the ops that touch it, in program order; an op is ready when it is the last pending op on every one
    // of its wires
    std::vector<std::vector<int64_t>> on_wire(hk.n_wires);
    for (int64_t idx = first_needed; idx < n_ops; ++idx) {
        if (skipped(ops.ops[idx])) continue;
        for (int w : ops.ops[idx].wires) {
            QSV_CHECK(w >= 0 && w < hk.n_wires, "wire out of range");
            on_wire[w].push_back(idx);
        }
    }
    std::vector<char> done(n_ops, 0);
    auto is_ready = [&](int64_t idx) {
        for (int w : ops.ops[idx].wires)
            if (on_wire[w].empty() || on_wire[w].back() != idx) return false;
        return true;
    };
    auto retire = [&](int64_t idx) {
        done[idx] = 1;
        for (int w : ops.ops[idx].wires) on_wire[w].pop_back();
    };
    size_t pending = 0;
    for (int64_t idx = first_needed; idx < n_ops; ++idx) pending += skipped(ops.ops[idx]) ? 0 : 1;
    const char *defer_env = std::getenv("QSV_ADJOINT_DEFER");
    const bool defer_diag = !(defer_env && std::atoi(defer_env) == 0);
    std::vector<LoweredGate> deferred_gens;  // diagonal generators waiting for the next layer's launch
    std::vector<int> deferred_slot0;
    while (pending) {
        std::vector<int64_t> ready;
        for (int w = 0; w < hk.n_wires; ++w)
            if (!on_wire[w].empty() && is_ready(on_wire[w].back()) &&
                std::find(ready.begin(), ready.end(), on_wire[w].back()) == ready.end())
                ready.push_back(on_wire[w].back());
        QSV_CHECK(!ready.empty(), "internal: adjoint scheduling made no progress");
        std::sort(ready.begin(), ready.end(), std::greater<int64_t>());
        // undo the ready ops and everything non-trainable that becomes ready behind them, in one fused batch
        // (collected first: which generators may wait depends on what the batch touches)
        std::vector<LoweredGate> batch;
        std::vector<int64_t> grown;
        auto take = [&](int64_t idx) {
            if (ops.ops[idx].name != "Identity") batch.push_back(hk.dagger(ops.ops[idx]));
            retire(idx);
            --pending;
        };
        for (int64_t idx : ready) take(idx);
        bool grew = true;
        while (grew && pending) {
            grew = false;
            for (int w = 0; w < hk.n_wires; ++w) {
                if (on_wire[w].empty()) continue;
                const int64_t idx = on_wire[w].back();
                if (tp_of[idx] >= 0 || !is_ready(idx)) continue;
                take(idx);
                grown.push_back(idx);
                grew = true;
            }
        }
        uint64_t grown_wires = 0;
        for (int64_t idx : grown)
            for (int w : ops.ops[idx].wires) grown_wires |= 1ull << w;
        // generator inner products of the trainable ready ops, all against the current vectors.  A diagonal generator
        // of a diagonal gate (RZ, PhaseShift, CRZ, IsingZZ, MultiRZ ...) commutes with its own gate and with the other
        // gates of this batch (disjoint wires) unless a gate that became ready behind it shares a wire: it may then be
        // evaluated AFTER the batch just as well, i.e. together with the generators of the next layer -- one read of
        // (bra, lambda) fewer per layer of such gates.
        std::vector<LoweredGate> gens;
        std::vector<int> gen_slot0;
        gens.swap(deferred_gens);
        gen_slot0.swap(deferred_slot0);
        for (int64_t idx : ready) {
            const int64_t tp = tp_of[idx];
            if (tp < 0) continue;
            const Op &op = ops.ops[idx];
            LoweredGenerator g = hk.generator(op);
            factor[tp] = -2.0 * g.scale * (op.inverse ? -1.0 : 1.0);
            extra[tp] = g.extra_identity;
            uint64_t op_wires = 0;
            for (int w : op.wires) op_wires |= 1ull << w;
            const bool gen_diag = g.op.kind == LoweredGate::DIAG || g.op.kind == LoweredGate::PARITY;
            bool gate_diag = false;
            if (defer_diag && gen_diag && g.extra_identity == 0.0 && (op_wires & grown_wires) == 0) {
                const LoweredGate lg = hk.dagger(op);
                gate_diag = lg.kind == LoweredGate::DIAG || lg.kind == LoweredGate::PARITY;
            }
            if (gate_diag) {
                deferred_gens.push_back(std::move(g.op));
                deferred_slot0.push_back((int)(tp * n_obs * 2));
                continue;
            }
            gens.push_back(std::move(g.op));
            gen_slot0.push_back((int)(tp * n_obs * 2));
            if (g.extra_identity != 0.0) hk.identity_inner_product(tp);
        }
        if (!gens.empty()) hk.inner_products(gens, gen_slot0);
        if (!batch.empty()) hk.apply(batch);
        // is any trainable op left?  if not, the remaining daggers are not needed
        bool trainable_left = false;
        for (int64_t idx = first_needed; idx < n_ops && !trainable_left; ++idx)
            trainable_left = !done[idx] && tp_of[idx] >= 0;
        if (!trainable_left) break;
    }
    if (!deferred_gens.empty()) hk.inner_products(deferred_gens, deferred_slot0);  // against the vectors after the last batch
}

// ------------------------------------------------------------------------------------------------
// Adjoint Jacobian
// ------------------------------------------------------------------------------------------------
void adjoint_jacobian(State &sv, const Ops &ops, const std::vector<const Obs *> &obs,
                      const std::vector<int64_t> &trainable, bool apply_operations, double *jac) {
    sv.use();
    QSV_CHECK(!trainable.empty(), "No trainable parameters provided.");
    const size_t n_obs = obs.size();
    const size_t n_tp = trainable.size();
    for (size_t i = 0; i < n_obs * n_tp; ++i) jac[i] = 0.0;
    if (n_obs == 0) return;
    for (const auto &op : ops.ops)
        QSV_CHECK(op.params.size() <= 1, "The operation is not supported using the adjoint differentiation method");

    // lambda and the bras H_lambda[i] = O_i lambda
    const size_t bytes = sv.bytes();
    std::vector<std::unique_ptr<State>> vecs;  // [0] = lambda, [1..] = bras
    std::vector<std::unique_ptr<DevBuf>> vec_mem;  // their memory: cached workspace, released when the call ends
    for (size_t i = 0; i < 1 + n_obs; ++i) {
        vec_mem.push_back(std::make_unique<DevBuf>(sv, bytes));
        auto s = std::make_unique<State>();
        s->n = sv.n;
        s->dtype = sv.dtype;
        s->device = sv.device;
        s->stream = sv.stream;
        s->data = vec_mem.back()->p;
        s->owns = false;
        s->swap_ok = true;  // an observable may exchange the buffer for another cached block of the same size
        vecs.push_back(std::move(s));
    }
    State &lambda = *vecs[0];
    QSV_CUDA(cudaMemcpyAsync(lambda.data, sv.data, bytes, cudaMemcpyDeviceToDevice, sv.stream));
    if (apply_operations)
        for (const auto &op : ops.ops) apply_op(lambda, op, false);
    for (size_t i = 0; i < n_obs; ++i) {
        QSV_CUDA(cudaMemcpyAsync(vecs[1 + i]->data, lambda.data, bytes, cudaMemcpyDeviceToDevice, sv.stream));
        apply_observable(*vecs[1 + i], *obs[i]);
        vec_mem[1 + i]->p = vecs[1 + i]->data;  // follow a buffer exchange made by the observable
    }
    // device table of vector pointers for the batched U^dagger launch
    std::vector<void *> h_table(1 + n_obs);
    for (size_t i = 0; i < 1 + n_obs; ++i) h_table[i] = vecs[i]->data;
    DevBuf d_table(h_table.size() * sizeof(void *));
    QSV_CUDA(cudaMemcpyAsync(d_table.p, h_table.data(), h_table.size() * sizeof(void *), cudaMemcpyHostToDevice,
                             sv.stream));

    // one complex slot per (trainable parameter, observable) for op part, one more for the identity part
    const size_t n_slots = 2 * n_obs * n_tp;
    double *red = sv.reduction_buffer(2 * n_slots);
    reduction_zero(sv, red, 2 * n_slots);
    std::vector<double> factor(n_tp, 0.0), extra(n_tp, 0.0);

    size_t n_par_ops = 0;
    for (const auto &op : ops.ops) n_par_ops += op.params.empty() ? 0 : 1;
    const char *batch_env = std::getenv("QSV_ADJOINT_BATCH");
    if (!(batch_env && std::atoi(batch_env) == 0)) {
        // ---- layered reverse sweep (layered_reverse_sweep below; shared with the sharded register, dist.cu) ----
        ReverseSweepHooks hk;
        hk.n_wires = sv.n;
        hk.n_obs = n_obs;
        hk.generator = [&](const Op &op) { return lower_generator(sv.n, op.name, op.wires); };
        hk.dagger = [&](const Op &op) { return lower_op(sv, op, true); };
        hk.inner_products = [&](const std::vector<LoweredGate> &gens, const std::vector<int> &slot0) {
            std::vector<int> slots(gens.size());
            for (size_t i = 0; i < n_obs; ++i) {
                for (size_t k = 0; k < gens.size(); ++k) slots[k] = slot0[k] + (int)(2 * i);
                launch_bra_gens_ket(sv, vecs[1 + i]->data, lambda.data, gens, slots, red);
            }
        };
        hk.identity_inner_product = [&](int64_t tp) {
            LoweredGate id;
            for (size_t i = 0; i < n_obs; ++i)
                launch_bra_op_ket(sv, vecs[1 + i]->data, lambda.data, id, red, (int)((tp * n_obs + i) * 2 + 1));
        };
        hk.apply = [&](const std::vector<LoweredGate> &batch) {
            apply_gates_tiled(sv, batch, (void *const *)d_table.p, (int)(1 + n_obs));
        };
        layered_reverse_sweep(ops, trainable, factor, extra, hk);
    } else {
    int64_t tp_pos = (int64_t)n_tp - 1;
    int64_t cur = (int64_t)n_par_ops - 1;
    for (int64_t idx = (int64_t)ops.ops.size() - 1; idx >= 0; --idx) {
        const Op &op = ops.ops[idx];
        if (op.name == "QubitStateVector" || op.name == "StatePrep" || op.name == "BasisState") continue;
        if (tp_pos < 0) break;
        if (!op.params.empty()) {
            if (cur == trainable[tp_pos]) {
                LoweredGenerator g = lower_generator(sv.n, op.name, op.wires);
                factor[tp_pos] = -2.0 * g.scale * (op.inverse ? -1.0 : 1.0);
                extra[tp_pos] = g.extra_identity;
                for (size_t i = 0; i < n_obs; ++i) {
                    launch_bra_op_ket(sv, vecs[1 + i]->data, lambda.data, g.op, red, (int)((tp_pos * n_obs + i) * 2));
                    if (g.extra_identity != 0.0) {
                        LoweredGate id;
                        launch_bra_op_ket(sv, vecs[1 + i]->data, lambda.data, id, red,
                                          (int)((tp_pos * n_obs + i) * 2 + 1));
                    }
                }
                --tp_pos;
            }
            --cur;
        }
        // lambda <- U^dagger lambda and H_lambda[i] <- U^dagger H_lambda[i], one launch
        if (op.name != "Identity")
            launch_gate_multi(sv, lower_op(sv, op, true), (void *const *)d_table.p, (int)(1 + n_obs));
    }
    }
    std::vector<double> h(2 * n_slots);
    reduction_read(sv, red, h.data(), 2 * n_slots);
    for (size_t p = 0; p < n_tp; ++p)
        for (size_t i = 0; i < n_obs; ++i) {
            const size_t s = (p * n_obs + i) * 2;
            const double im = h[2 * s + 1] + extra[p] * h[2 * (s + 1) + 1];
            jac[i * n_tp + p] = factor[p] * im;
        }
}

}  // namespace qsv
