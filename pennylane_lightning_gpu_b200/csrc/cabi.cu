// extern "C" surface of libqsv_b200.so (see include/qsv_b200.h for the reference call sites each
// entry replaces).  Every entry converts C++ exceptions into a status code + thread-local message.
#include <algorithm>
#include <cstdio>

#include "qsv_internal.h"

namespace qsv {
const std::string &last_error_ref();
}

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

namespace {

std::vector<int> ivec(const int *p, int n) {
    QSV_CHECK(n >= 0 && (n == 0 || p != nullptr), "null wire list");
    return std::vector<int>(p, p + n);
}
std::vector<double> dvec(const double *p, int n) {
    QSV_CHECK(n >= 0 && (n == 0 || p != nullptr), "null parameter list");
    return std::vector<double>(p, p + n);
}
std::vector<cplx> cvec(const double *p, size_t n) {
    std::vector<cplx> v(n);
    for (size_t i = 0; i < n; ++i) v[i] = cplx(p[2 * i], p[2 * i + 1]);
    return v;
}
void need(const void *p, const char *what) { QSV_CHECK(p != nullptr, std::string("null ") + what); }

void check_device_is_blackwell(int device) {
    cudaDeviceProp prop;
    QSV_CUDA(cudaGetDeviceProperties(&prop, device));
    QSV_CHECK(prop.major == 10, "libqsv_b200 is built for sm_100a only; device " + std::to_string(device) +
                                    " is sm_" + std::to_string(prop.major) + std::to_string(prop.minor));
}

}  // namespace

extern "C" {

const char *qsv_last_error(void) { return last_error_ref().c_str(); }
int qsv_version(void) { return 100; }

int qsv_device_count(int *count) {
    QSV_API_BEGIN
    need(count, "count");
    QSV_CUDA(cudaGetDeviceCount(count));
    QSV_API_END
}

int qsv_device_arch(int device, int *major, int *minor) {
    QSV_API_BEGIN
    cudaDeviceProp prop;
    QSV_CUDA(cudaGetDeviceProperties(&prop, device));
    if (major) *major = prop.major;
    if (minor) *minor = prop.minor;
    QSV_API_END
}

int qsv_device_mem_info(int device, size_t *free_bytes, size_t *total_bytes) {
    QSV_API_BEGIN
    QSV_CUDA(cudaSetDevice(device));
    size_t f = 0, t = 0;
    QSV_CUDA(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = f;
    if (total_bytes) *total_bytes = t;
    QSV_API_END
}

int qsv_device_reset(void) {
    QSV_API_BEGIN
    int count = 0;
    QSV_CUDA(cudaGetDeviceCount(&count));
    for (int d = 0; d < count; ++d) {
        QSV_CUDA(cudaSetDevice(d));
        ws_trim(d);  // cached workspace blocks die with the context
        QSV_CUDA(cudaDeviceReset());
    }
    QSV_API_END
}

int qsv_enable_peer_access(void) {
    QSV_API_BEGIN
    int count = 0;
    QSV_CUDA(cudaGetDeviceCount(&count));
    for (int a = 0; a < count; ++a) {
        QSV_CUDA(cudaSetDevice(a));
        for (int b = 0; b < count; ++b) {
            if (a == b) continue;
            int can = 0;
            QSV_CUDA(cudaDeviceCanAccessPeer(&can, a, b));
            if (!can) continue;
            cudaError_t e = cudaDeviceEnablePeerAccess(b, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled)
                cudaGetLastError();
            else
                QSV_CUDA(e);
        }
    }
    QSV_API_END
}

int qsv_create_external(int n_qubits, int dtype, int device, void *device_ptr, void *cuda_stream,
                        qsv_state **out) {
    QSV_API_BEGIN
    need(out, "output handle");
    QSV_CHECK(n_qubits >= 1 && n_qubits <= 40, "number of qubits must be in [1, 40]");
    QSV_CHECK(dtype == QSV_C64 || dtype == QSV_C128, "dtype must be QSV_C64 or QSV_C128");
    int count = 0;
    QSV_CUDA(cudaGetDeviceCount(&count));
    QSV_CHECK(device >= 0 && device < count, "invalid CUDA device " + std::to_string(device));
    check_device_is_blackwell(device);
    auto sv = std::make_unique<qsv_state>();
    sv->n = n_qubits;
    sv->dtype = dtype;
    sv->device = device;
    sv->stream = (cudaStream_t)cuda_stream;
    sv->use();
    if (device_ptr) {
        sv->data = device_ptr;
        sv->owns = false;
    } else {
        // at least 2 MiB: smaller cudaMalloc blocks are sub-allocated from a shared block, which cannot be exported
        // one by one over CUDA IPC when the state later becomes a shard of a distributed register
        QSV_CUDA(cudaMalloc(&sv->data, std::max<size_t>(sv->bytes(), (size_t)2 << 20)));
        sv->owns = true;
        launch_fill_basis(*sv, 0);
    }
    *out = sv.release();
    QSV_API_END
}

int qsv_create(int n_qubits, int dtype, int device, qsv_state **out) {
    return qsv_create_external(n_qubits, dtype, device, nullptr, nullptr, out);
}

int qsv_destroy(qsv_state *sv) {
    QSV_API_BEGIN
    delete sv;
    QSV_API_END
}

int qsv_set_stream(qsv_state *sv, void *cuda_stream) {
    QSV_API_BEGIN
    need(sv, "state");
    sv->use();
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    sv->stream = (cudaStream_t)cuda_stream;
    QSV_API_END
}

int qsv_synchronize(qsv_state *sv) {
    QSV_API_BEGIN
    need(sv, "state");
    sv->use();
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    QSV_API_END
}

void *qsv_data_ptr(qsv_state *sv) { return sv ? sv->data : nullptr; }
int qsv_num_qubits(const qsv_state *sv) { return sv ? sv->n : -1; }
int qsv_dtype(const qsv_state *sv) { return sv ? sv->dtype : -1; }
int qsv_device(const qsv_state *sv) { return sv ? sv->device : -1; }

int qsv_set_basis_state(qsv_state *sv, uint64_t index) {
    QSV_API_BEGIN
    need(sv, "state");
    launch_fill_basis(*sv, index);
    QSV_API_END
}

int qsv_set_state_vector(qsv_state *sv, const int64_t *indices, const void *values, size_t count) {
    QSV_API_BEGIN
    need(sv, "state");
    QSV_CHECK(count == 0 || (indices && values), "null index/value arrays");
    sv->use();
    for (size_t i = 0; i < count; ++i)
        QSV_CHECK(indices[i] >= 0 && (uint64_t)indices[i] < sv->length(), "state-vector index out of range");
    QSV_CUDA(cudaMemsetAsync(sv->data, 0, sv->bytes(), sv->stream));
    if (count) {
        const size_t vb = count * sv->amp_bytes();
        // values first: they need the 16-byte alignment of the 128-bit loads, the indices only 8
        const size_t vb_al = (vb + 15) / 16 * 16;
        char *scr = (char *)sv->scratch_buffer(vb_al + count * 8);
        QSV_CUDA(cudaMemcpyAsync(scr, values, vb, cudaMemcpyHostToDevice, sv->stream));
        QSV_CUDA(cudaMemcpyAsync(scr + vb_al, indices, count * 8, cudaMemcpyHostToDevice, sv->stream));
        launch_scatter(*sv, (const int64_t *)(scr + vb_al), scr, count);
    }
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    QSV_API_END
}

int qsv_h2d(qsv_state *sv, const void *host, size_t n_amps) {
    QSV_API_BEGIN
    need(sv, "state");
    need(host, "host buffer");
    QSV_CHECK(n_amps <= sv->length(), "host buffer larger than the state vector");
    copy_state_host(*sv, sv->data, const_cast<void *>(host), n_amps * sv->amp_bytes(), true);
    QSV_API_END
}

int qsv_d2h(qsv_state *sv, void *host, size_t n_amps) {
    QSV_API_BEGIN
    need(sv, "state");
    need(host, "host buffer");
    QSV_CHECK(n_amps <= sv->length(), "host buffer larger than the state vector");
    copy_state_host(*sv, sv->data, host, n_amps * sv->amp_bytes(), false);
    QSV_API_END
}

int qsv_d2d(qsv_state *dst, const qsv_state *src) {
    QSV_API_BEGIN
    need(dst, "destination state");
    need(src, "source state");
    QSV_CHECK(dst->n == src->n && dst->dtype == src->dtype, "state vectors differ in size or precision");
    dst->use();
    // order after everything queued on the source stream
    if (src->stream != dst->stream) QSV_CUDA(cudaStreamSynchronize(src->stream));
    QSV_CUDA(cudaMemcpyAsync(dst->data, src->data, dst->bytes(), cudaMemcpyDefault, dst->stream));
    QSV_CUDA(cudaStreamSynchronize(dst->stream));
    QSV_API_END
}

int qsv_apply_named(qsv_state *sv, const char *name, const int *wires, int n_wires, int adjoint,
                    const double *params, int n_params) {
    QSV_API_BEGIN
    need(sv, "state");
    need(name, "gate name");
    Op op;
    op.name = name;
    op.wires = ivec(wires, n_wires);
    op.params = dvec(params, n_params);
    op.inverse = adjoint != 0;
    apply_op(*sv, op, false);
    QSV_API_END
}

int qsv_apply_matrix(qsv_state *sv, const double *matrix, const int *ctrl_wires, int n_ctrls,
                     const int *tgt_wires, int n_tgts, int adjoint) {
    QSV_API_BEGIN
    need(sv, "state");
    need(matrix, "matrix");
    QSV_CHECK(n_tgts >= 1 && n_tgts <= 10, "matrix gates act on 1..10 target wires");
    const size_t dim = 1ull << n_tgts;
    std::vector<cplx> m = cvec(matrix, dim * dim);
    launch_gate(*sv, lower_matrix(sv->n, m.data(), ivec(ctrl_wires, n_ctrls), ivec(tgt_wires, n_tgts), adjoint != 0));
    QSV_API_END
}

int qsv_apply_generator(qsv_state *sv, const char *name, const int *wires, int n_wires, int adjoint,
                        double *scale) {
    QSV_API_BEGIN
    need(sv, "state");
    need(name, "gate name");
    (void)adjoint;  // generators are Hermitian
    LoweredGenerator g = lower_generator(sv->n, name, ivec(wires, n_wires));
    if (scale) *scale = g.scale;
    // G = op + extra * 1, where `op` is ZERO outside the amplitudes it touches (projector
    // semantics for its controls / two-level pairs) while launch_gate leaves them unchanged.
    //   tmp = gate(psi)            [op psi | psi]        (touched | untouched)
    //   tmp += extra * psi         [op psi + e psi | (1 + e) psi]
    //   psi  = zero the touched    [0 | psi]
    //   tmp -= psi                 [op psi + e psi | e psi]  = G psi
    // This entry exists for API parity (Managed.hpp:563-688); the adjoint sweep itself never
    // materialises G psi (see launch_bra_op_ket).
    sv->use();
    void *tmp = nullptr;
    QSV_CUDA(cudaMalloc(&tmp, sv->bytes()));
    try {
        QSV_CUDA(cudaMemcpyAsync(tmp, sv->data, sv->bytes(), cudaMemcpyDeviceToDevice, sv->stream));
        void *orig = sv->data;
        sv->data = tmp;
        launch_gate(*sv, g.op);
        sv->data = orig;
        if (g.extra_identity != 0.0) launch_axpy(*sv, cplx(g.extra_identity, 0.0), sv->data, tmp);
        LoweredGate zero = g.op;
        for (auto &c : zero.mat) c = cplx(0.0, 0.0);
        launch_gate(*sv, zero);
        launch_axpy(*sv, cplx(-1.0, 0.0), sv->data, tmp);
        QSV_CUDA(cudaMemcpyAsync(sv->data, tmp, sv->bytes(), cudaMemcpyDeviceToDevice, sv->stream));
        QSV_CUDA(cudaStreamSynchronize(sv->stream));
    } catch (...) {
        cudaFree(tmp);
        throw;
    }
    QSV_CUDA(cudaFree(tmp));
    QSV_API_END
}

int qsv_ops_create(qsv_ops **out) {
    QSV_API_BEGIN
    need(out, "output handle");
    *out = new qsv_ops();
    QSV_API_END
}

int qsv_ops_destroy(qsv_ops *ops) {
    QSV_API_BEGIN
    delete ops;
    QSV_API_END
}

int qsv_ops_append(qsv_ops *ops, const char *name, const int *wires, int n_wires, const double *params,
                   int n_params, int inverse, const double *matrix, size_t mat_dim) {
    QSV_API_BEGIN
    need(ops, "ops");
    need(name, "gate name");
    Op op;
    op.name = name;
    op.wires = ivec(wires, n_wires);
    op.params = dvec(params, n_params);
    op.inverse = inverse != 0;
    if (matrix && mat_dim) op.matrix = cvec(matrix, mat_dim * mat_dim);
    ops->ops.push_back(std::move(op));
    QSV_API_END
}

int qsv_ops_size(const qsv_ops *ops) { return ops ? (int)ops->ops.size() : -1; }

static std::vector<LoweredGate> lower_all(int n_qubits, const qsv_ops *ops) {
    std::vector<LoweredGate> gates;
    gates.reserve(ops->ops.size());
    for (const auto &op : ops->ops) {
        if (op.name == "Identity") continue;
        if (find_gate(op.name) != nullptr) {
            gates.push_back(lower_named(n_qubits, op.name, op.wires, op.params, op.inverse));
        } else {
            QSV_CHECK(!op.matrix.empty(), "Currently unsupported gate: " + op.name);
            const size_t dim = 1ull << op.wires.size();
            QSV_CHECK(op.matrix.size() == dim * dim, "matrix of gate " + op.name + " does not match its wires");
            gates.push_back(lower_matrix(n_qubits, op.matrix.data(), {}, op.wires, op.inverse));
        }
    }
    return gates;
}

int qsv_apply_ops(qsv_state *sv, const qsv_ops *ops, int fuse) {
    QSV_API_BEGIN
    need(sv, "state");
    need(ops, "ops");
    sv->stat_launches = 0;
    sv->stat_sweeps = 0;
    if (!fuse) {
        for (const auto &op : ops->ops) apply_op(*sv, op, false);
    } else {
        apply_ops_fused(*sv, lower_all(sv->n, ops));
    }
    QSV_API_END
}

int qsv_ops_plan_work(const qsv_ops *ops, int n_qubits, int dtype, double *fma_per_amplitude, int64_t *n_sweeps,
                      int64_t *n_passes) {
    QSV_API_BEGIN
    need(ops, "ops");
    QSV_CHECK(n_qubits >= 12 && n_qubits <= 62, "sweep planning needs 12..62 qubits");
    QSV_CHECK(dtype == QSV_C64 || dtype == QSV_C128, "dtype must be QSV_C64 or QSV_C128");
    const std::vector<LoweredGate> merged = prepare_gates_regs(lower_all(n_qubits, ops));
    const std::vector<SweepPlan> plan = plan_sweeps_cached(n_qubits, merged, 4, true, 48, 512, dtype);
    double fma = 0.0;
    int64_t passes = 0;
    std::vector<const LoweredGate *> cur;
    for (const SweepPlan &sw : plan) {
        if (!sw.fused) {
            // a lone gate through the one-sweep kernels: 2^k x 2^k block = 4 * 2^k multiply-adds per amplitude, diagonal 4
            const LoweredGate &g = merged[sw.gates[0]];
            const double w = g.kind == LoweredGate::DENSE ? 4.0 * (double)(1 << g.k) : 4.0;
            fma += w / (double)(1ull << __builtin_popcountll(g.ctrl_mask));
            continue;
        }
        cur.clear();
        for (int i : sw.gates) cur.push_back(&merged[i]);
        double f = 0.0;
        int p = 0;
        regs_sweep_work(n_qubits, dtype, cur, sw.need, 4, &f, &p);
        fma += f;
        passes += p;
    }
    if (fma_per_amplitude) *fma_per_amplitude = fma;
    if (n_sweeps) *n_sweeps = (int64_t)plan.size();
    if (n_passes) *n_passes = passes;
    QSV_API_END
}

int qsv_ops_plan_sweeps(const qsv_ops *ops, int n_qubits, int dag, int low_bits, int64_t *n_gates_merged,
                        int64_t *n_sweeps, int64_t *max_gates_per_sweep, int *order_valid) {
    QSV_API_BEGIN
    need(ops, "ops");
    QSV_CHECK(n_qubits >= 12 && n_qubits <= 62, "sweep planning needs 12..62 qubits");
    const int L = std::max(1, std::min(low_bits > 0 ? low_bits : 4, 11));
    const std::vector<LoweredGate> merged = prepare_gates_regs(lower_all(n_qubits, ops));
    const std::vector<SweepPlan> plan = plan_sweeps_cached(n_qubits, merged, L, dag != 0, 48, 512, QSV_C128);
    // self-check: every gate exactly once, and every pair that does not commute structurally keeps its order
    std::vector<int64_t> pos(merged.size(), -1);
    int64_t at = 0, biggest = 0, total = 0;
    bool ok = true;
    for (const SweepPlan &sw : plan) {
        biggest = std::max<int64_t>(biggest, (int64_t)sw.gates.size());
        for (int i : sw.gates) {
            ok = ok && i >= 0 && i < (int)merged.size() && pos[i] < 0;
            if (!ok) break;
            pos[i] = at++;
        }
        if (sw.fused) ok = ok && __builtin_popcountll(sw.need) <= 12 - L;
    }
    for (size_t i = 0; ok && i < merged.size(); ++i) {
        if (merged[i].kind == LoweredGate::NOP) continue;
        ++total;
        ok = ok && pos[i] >= 0;
        for (size_t j = i + 1; ok && j < merged.size(); ++j)
            if (merged[j].kind != LoweredGate::NOP && !gates_commute_structurally(merged[i], merged[j]))
                ok = pos[j] >= 0 && pos[i] < pos[j];
    }
    ok = ok && at == total;
    if (n_gates_merged) *n_gates_merged = total;
    if (n_sweeps) *n_sweeps = (int64_t)plan.size();
    if (max_gates_per_sweep) *max_gates_per_sweep = biggest;
    if (order_valid) *order_valid = ok ? 1 : 0;
    QSV_API_END
}

int qsv_last_apply_stats(const qsv_state *sv, int64_t *launches, int64_t *sweeps) {
    QSV_API_BEGIN
    need(sv, "state");
    if (launches) *launches = sv->stat_launches;
    if (sweeps) *sweeps = sv->stat_sweeps;
    QSV_API_END
}

int qsv_expval_named(qsv_state *sv, const char *name, const int *wires, int n_wires, const double *params,
                     int n_params, double *out) {
    QSV_API_BEGIN
    need(sv, "state");
    need(name, "observable name");
    need(out, "output");
    QSV_CHECK(find_gate(name) != nullptr, std::string("Currently unsupported observable: ") + name);
    std::vector<cplx> m = named_gate_matrix(name, dvec(params, n_params), n_wires);
    double *red = sv->reduction_buffer(2);
    reduction_zero(*sv, red, 2);
    launch_bra_op_ket(*sv, sv->data, sv->data, lower_matrix(sv->n, m.data(), {}, ivec(wires, n_wires), false), red, 0);
    reduction_read(*sv, red, out, 2);
    QSV_API_END
}

int qsv_expval_matrix(qsv_state *sv, const double *matrix, const int *wires, int n_wires, double *out) {
    QSV_API_BEGIN
    need(sv, "state");
    need(matrix, "matrix");
    need(out, "output");
    QSV_CHECK(n_wires >= 1 && n_wires <= 10, "dense observables act on 1..10 wires");
    const size_t dim = 1ull << n_wires;
    std::vector<cplx> m = cvec(matrix, dim * dim);
    double *red = sv->reduction_buffer(2);
    reduction_zero(*sv, red, 2);
    launch_bra_op_ket(*sv, sv->data, sv->data, lower_matrix(sv->n, m.data(), {}, ivec(wires, n_wires), false), red, 0);
    reduction_read(*sv, red, out, 2);
    QSV_API_END
}

int qsv_expval_pauli_words(qsv_state *sv, int n_terms, const char *letters, const int *wires, const int *offsets,
                           const double *coeffs, double *per_term, double *out) {
    QSV_API_BEGIN
    need(sv, "state");
    QSV_CHECK(n_terms >= 0, "negative number of terms");
    if (n_terms > 0) {
        need(letters, "letters");
        need(wires, "wires");
        need(offsets, "offsets");
    }
    const int n = sv->n;
    double *red = sv->reduction_buffer(2 * (size_t)std::max(n_terms, 1));
    reduction_zero(*sv, red, 2 * (size_t)std::max(n_terms, 1));
    std::vector<uint64_t> xs(n_terms), zs(n_terms);
    std::vector<int> nys(n_terms);
    for (int t = 0; t < n_terms; ++t) {
        uint64_t x = 0, z = 0;
        int ny = 0;
        for (int j = offsets[t]; j < offsets[t + 1]; ++j) {
            QSV_CHECK(wires[j] >= 0 && wires[j] < n, "Pauli word wire out of range");
            const uint64_t b = 1ull << (n - 1 - wires[j]);
            QSV_CHECK(((x | z) & b) == 0 || letters[j] == 'I', "repeated wire in a Pauli word");
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
    // single-pass fused evaluation: words whose X/Y letters fit one tile share one read of the state
    launch_bra_paulis_ket(*sv, sv->data, sv->data, n_terms, xs.data(), zs.data(), nys.data(), 0, red);
    std::vector<double> h(2 * (size_t)std::max(n_terms, 1));
    reduction_read(*sv, red, h.data(), h.size());
    double tot = 0;
    for (int t = 0; t < n_terms; ++t) {
        double e = h[2 * t];
        if (per_term) per_term[t] = e;
        // Managed.hpp:1137-1146: for complex64 the per-term value is cast to float before the dot
        if (sv->dtype == QSV_C64) e = (double)(float)e;
        if (coeffs) tot += e * coeffs[2 * t];
    }
    if (out) *out = tot;
    QSV_API_END
}

int qsv_expval_csr(qsv_state *sv, const void *row_offsets, const void *col_indices, const double *values,
                   int64_t nnz, int index_bytes, double *out) {
    QSV_API_BEGIN
    need(sv, "state");
    need(row_offsets, "row offsets");
    need(out, "output");
    QSV_CHECK(index_bytes == 4 || index_bytes == 8, "CSR index width must be 4 or 8 bytes");
    QSV_CHECK(nnz >= 0, "negative nnz");
    sv->use();
    const size_t rows = sv->length();
    const size_t ib = (size_t)index_bytes;
    const size_t b_ptr = (rows + 1) * ib, b_idx = (size_t)nnz * ib, b_val = (size_t)nnz * 16;
    const size_t o_idx = (b_ptr + 255) / 256 * 256, o_val = o_idx + (b_idx + 255) / 256 * 256;
    void *dev = nullptr;
    QSV_CUDA(cudaMalloc(&dev, o_val + b_val + 256));
    try {
        char *d = (char *)dev;
        QSV_CUDA(cudaMemcpyAsync(d, row_offsets, b_ptr, cudaMemcpyHostToDevice, sv->stream));
        if (nnz) {
            QSV_CUDA(cudaMemcpyAsync(d + o_idx, col_indices, b_idx, cudaMemcpyHostToDevice, sv->stream));
            QSV_CUDA(cudaMemcpyAsync(d + o_val, values, b_val, cudaMemcpyHostToDevice, sv->stream));
        }
        double *red = sv->reduction_buffer(2);
        reduction_zero(*sv, red, 2);
        launch_csr(*sv, sv->data, nullptr, d, d + o_idx, d + o_val, (int64_t)rows, nnz, index_bytes, red, 0);
        double h[2];
        reduction_read(*sv, red, h, 2);
        *out = h[0];
    } catch (...) {
        cudaFree(dev);
        throw;
    }
    QSV_CUDA(cudaFree(dev));
    QSV_API_END
}

int qsv_probs(qsv_state *sv, const int *wires, int n_wires, double *out) {
    QSV_API_BEGIN
    need(sv, "state");
    need(out, "output");
    std::vector<int> w = ivec(wires, n_wires);
    std::vector<int> bits;
    for (int x : w) {
        QSV_CHECK(x >= 0 && x < sv->n, "wire out of range");
        bits.push_back(sv->n - 1 - x);
    }
    for (size_t i = 0; i < bits.size(); ++i)
        for (size_t j = i + 1; j < bits.size(); ++j) QSV_CHECK(bits[i] != bits[j], "repeated wire in probability");
    launch_probs(*sv, bits, out);
    QSV_API_END
}

int qsv_sample(qsv_state *sv, const double *uniforms, int64_t shots, uint64_t *out) {
    QSV_API_BEGIN
    need(sv, "state");
    QSV_CHECK(shots >= 0, "negative number of shots");
    if (shots > 0) {
        need(uniforms, "uniform random numbers");
        need(out, "output");
    }
    launch_sample(*sv, uniforms, shots, out);
    QSV_API_END
}

int qsv_inner_product(qsv_state *a, qsv_state *b, double *out) {
    QSV_API_BEGIN
    need(a, "state a");
    need(b, "state b");
    need(out, "output");
    QSV_CHECK(a->n == b->n && a->dtype == b->dtype && a->device == b->device,
              "state vectors differ in size, precision or device");
    if (a->stream != b->stream) QSV_CUDA(cudaStreamSynchronize(a->stream));
    double *red = b->reduction_buffer(2);
    reduction_zero(*b, red, 2);
    LoweredGate id;
    launch_bra_op_ket(*b, a->data, b->data, id, red, 0);
    reduction_read(*b, red, out, 2);
    QSV_API_END
}

int qsv_axpy(const double *alpha, const qsv_state *x, qsv_state *y) {
    QSV_API_BEGIN
    need(alpha, "alpha");
    need(x, "x");
    need(y, "y");
    QSV_CHECK(x->n == y->n && x->dtype == y->dtype && x->device == y->device,
              "state vectors differ in size, precision or device");
    if (x->stream != y->stream) QSV_CUDA(cudaStreamSynchronize(x->stream));
    launch_axpy(*y, cplx(alpha[0], alpha[1]), x->data, y->data);
    QSV_CUDA(cudaStreamSynchronize(y->stream));
    QSV_API_END
}

// ---- observables ----
int qsv_obs_named(const char *name, const int *wires, int n_wires, const double *params, int n_params,
                  qsv_obs **out) {
    QSV_API_BEGIN
    need(name, "observable name");
    need(out, "output handle");
    auto o = std::make_shared<Obs>();
    o->kind = Obs::NAMED;
    o->name = name;
    o->wires = ivec(wires, n_wires);
    o->params = dvec(params, n_params);
    *out = new qsv_obs{o};
    QSV_API_END
}

int qsv_obs_hermitian(const double *matrix, size_t mat_dim, const int *wires, int n_wires, qsv_obs **out) {
    QSV_API_BEGIN
    need(matrix, "matrix");
    need(out, "output handle");
    QSV_CHECK(n_wires >= 1 && n_wires <= 10 && mat_dim == (1ull << n_wires), "Hermitian matrix does not match its wires");
    auto o = std::make_shared<Obs>();
    o->kind = Obs::HERMITIAN;
    o->wires = ivec(wires, n_wires);
    o->matrix = cvec(matrix, mat_dim * mat_dim);
    *out = new qsv_obs{o};
    QSV_API_END
}

int qsv_obs_tensor(qsv_obs *const *children, int n_children, qsv_obs **out) {
    QSV_API_BEGIN
    need(out, "output handle");
    QSV_CHECK(n_children >= 1 && children, "tensor product needs at least one factor");
    auto o = std::make_shared<Obs>();
    o->kind = Obs::TENSOR;
    for (int i = 0; i < n_children; ++i) {
        need(children[i], "child observable");
        o->children.push_back(children[i]->p);
    }
    *out = new qsv_obs{o};
    QSV_API_END
}

int qsv_obs_hamiltonian(const double *coeffs, qsv_obs *const *children, int n_children, qsv_obs **out) {
    QSV_API_BEGIN
    need(out, "output handle");
    QSV_CHECK(n_children >= 1 && children && coeffs, "Hamiltonian needs at least one term");
    auto o = std::make_shared<Obs>();
    o->kind = Obs::HAMILTONIAN;
    for (int i = 0; i < n_children; ++i) {
        need(children[i], "child observable");
        o->children.push_back(children[i]->p);
        o->coeffs.push_back(coeffs[i]);
    }
    *out = new qsv_obs{o};
    QSV_API_END
}

int qsv_obs_sparse(const int64_t *row_offsets, int64_t n_rows_plus_1, const int64_t *col_indices,
                   const double *values, int64_t nnz, qsv_obs **out) {
    QSV_API_BEGIN
    need(out, "output handle");
    need(row_offsets, "row offsets");
    QSV_CHECK(n_rows_plus_1 >= 2 && nnz >= 0, "invalid CSR sizes");
    auto o = std::make_shared<Obs>();
    o->kind = Obs::SPARSE;
    o->indptr.assign(row_offsets, row_offsets + n_rows_plus_1);
    if (nnz) {
        need(col_indices, "column indices");
        need(values, "values");
        o->indices.assign(col_indices, col_indices + nnz);
        o->values = cvec(values, (size_t)nnz);
    }
    *out = new qsv_obs{o};
    QSV_API_END
}

int qsv_obs_destroy(qsv_obs *obs) {
    QSV_API_BEGIN
    delete obs;
    QSV_API_END
}

int qsv_obs_apply(const qsv_obs *obs, qsv_state *sv) {
    QSV_API_BEGIN
    need(obs, "observable");
    need(sv, "state");
    apply_observable(*sv, *obs->p);
    QSV_CUDA(cudaStreamSynchronize(sv->stream));
    QSV_API_END
}

int qsv_obs_expval(const qsv_obs *obs, qsv_state *sv, double *out) {
    QSV_API_BEGIN
    need(obs, "observable");
    need(sv, "state");
    need(out, "output");
    *out = observable_expval(*sv, *obs->p);
    QSV_API_END
}

int qsv_adjoint_jacobian(qsv_state *sv, const qsv_ops *ops, qsv_obs *const *observables, int n_obs,
                         const int64_t *trainable, int n_trainable, int apply_operations, double *jac) {
    QSV_API_BEGIN
    need(sv, "state");
    need(ops, "ops");
    QSV_CHECK(n_obs >= 0 && n_trainable >= 0, "negative sizes");
    std::vector<const Obs *> o;
    for (int i = 0; i < n_obs; ++i) {
        need(observables[i], "observable");
        o.push_back(observables[i]->p.get());
    }
    std::vector<int64_t> tp(trainable, trainable + n_trainable);
    QSV_CHECK(n_trainable == 0 || jac != nullptr, "null Jacobian output");
    adjoint_jacobian(*sv, *ops, o, tp, apply_operations != 0, jac);
    QSV_API_END
}

}  // extern "C"
