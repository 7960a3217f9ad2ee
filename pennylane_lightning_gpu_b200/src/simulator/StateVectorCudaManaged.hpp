// StateVectorCudaManaged<PrecisionT>: the reference's single-GPU state-vector class
// (pennylane_lightning_gpu/src/simulator/StateVectorCudaManaged.hpp + StateVectorCudaBase.hpp) with its
// public method names and argument meanings, implemented as a thin shell over the C ABI of
// libqsv_b200.so.  Nothing here includes CUDA, cuStateVec, cuSPARSE or cuBLAS.
//
//   reference method (file:line)                          C-ABI call
//   ctor/dtor (Managed.hpp:83-134)                        qsv_create / qsv_destroy
//   initSV, setBasisState (Base.hpp:234, Managed:144)     qsv_set_basis_state
//   setStateVector (Managed.hpp:164-185)                  qsv_set_state_vector
//   applyOperation / applyOperation_std (:198-267)        qsv_apply_named / qsv_apply_matrix
//   35 named apply* (:321-560)                            qsv_apply_named
//   applyGenerator* (:563-688)                            qsv_apply_generator
//   expval x3 (:702-780)                                  qsv_expval_named / qsv_expval_matrix
//   getExpectationValueOnSparseSpMV (:795-922)            qsv_expval_csr
//   getExpectationValuePauliWords (:1071-1148)            qsv_expval_pauli_words
//   probability (:931-970), generate_samples (:982-1061)  qsv_probs / qsv_sample
//   CopyHostDataToGpu / CopyGpuDataToHost / updateData    qsv_h2d / qsv_d2h / qsv_d2d
#pragma once
#include <complex>
#include <cstddef>
#include <cstdint>
#include <random>
#include <string>
#include <type_traits>
#include <vector>

#include "DevTag.hpp"
#include "Error.hpp"
#include "qsv_b200.h"

namespace Pennylane {

template <class PrecisionT> class StateVectorCudaManaged {
  public:
    using Precision = PrecisionT;
    using ComplexT = std::complex<PrecisionT>;
    using CFP_t = ComplexT;
    static constexpr int dtype_code = std::is_same_v<PrecisionT, double> ? QSV_C128 : QSV_C64;

    StateVectorCudaManaged() = delete;
    explicit StateVectorCudaManaged(std::size_t num_qubits, CUDA::DevTag<int> dev_tag = {0, nullptr},
                                    bool /*alloc*/ = true)
        : num_qubits_(num_qubits), dev_tag_(dev_tag) {
        create();
    }
    StateVectorCudaManaged(const ComplexT *host_data, std::size_t length, CUDA::DevTag<int> dev_tag = {0, nullptr})
        : num_qubits_(log2_exact(length)), dev_tag_(dev_tag) {
        create();
        CopyHostDataToGpu(host_data, length, false);
    }
    StateVectorCudaManaged(const StateVectorCudaManaged &other)
        : num_qubits_(other.num_qubits_), dev_tag_(other.dev_tag_) {
        create();
        Util::check(qsv_d2d(sv_, other.sv_));
    }
    StateVectorCudaManaged &operator=(const StateVectorCudaManaged &) = delete;
    StateVectorCudaManaged(StateVectorCudaManaged &&o) noexcept
        : sv_(o.sv_), num_qubits_(o.num_qubits_), dev_tag_(o.dev_tag_) {
        o.sv_ = nullptr;
    }
    ~StateVectorCudaManaged() {
        if (sv_) qsv_destroy(sv_);
    }

    // ---- sizes / handles -----------------------------------------------------------------------
    std::size_t getNumQubits() const { return num_qubits_; }
    std::size_t getLength() const { return std::size_t{1} << num_qubits_; }
    qsv_state *handle() const { return sv_; }
    const CUDA::DevTag<int> &getDevTag() const { return dev_tag_; }
    void *getData() const { return qsv_data_ptr(sv_); }

    // ---- initialisation and copies ---------------------------------------------------------------
    void initSV(bool /*async*/ = false) { Util::check(qsv_set_basis_state(sv_, 0)); }
    void setBasisState(const ComplexT & /*value*/, std::size_t index, bool /*async*/ = false) {
        Util::check(qsv_set_basis_state(sv_, index));
    }
    template <class index_type>
    void setStateVector(index_type num_indices, const ComplexT *values, const index_type *indices,
                        bool /*async*/ = false) {
        std::vector<int64_t> idx(indices, indices + num_indices);
        Util::check(qsv_set_state_vector(sv_, idx.data(), values, static_cast<std::size_t>(num_indices)));
    }
    void CopyHostDataToGpu(const ComplexT *host, std::size_t length, bool /*async*/ = false) {
        PL_ABORT_IF_NOT(getLength() == length, "Sizes do not match for Host and GPU data");
        Util::check(qsv_h2d(sv_, host, length));
    }
    void CopyHostDataToGpu(const std::vector<ComplexT> &sv, bool async = false) {
        CopyHostDataToGpu(sv.data(), sv.size(), async);
    }
    void CopyGpuDataToHost(ComplexT *host, std::size_t length, bool /*async*/ = false) const {
        PL_ABORT_IF_NOT(getLength() == length, "Sizes do not match for Host and GPU data");
        Util::check(qsv_d2h(sv_, host, length));
    }
    void updateData(const StateVectorCudaManaged &other, bool /*async*/ = false) {
        PL_ABORT_IF_NOT(getLength() == other.getLength(), "Sizes do not match for GPU data");
        Util::check(qsv_d2d(sv_, other.sv_));
    }

    // ---- gates --------------------------------------------------------------------------------------
    void applyOperation(const std::string &opName, const std::vector<std::size_t> &wires, bool adjoint = false,
                        const std::vector<PrecisionT> &params = {0.0},
                        const std::vector<ComplexT> &gate_matrix = {}) {
        const std::vector<int> w = to_int(wires);
        if (!gate_matrix.empty() && !is_named(opName)) {
            const std::vector<double> m = to_doubles(gate_matrix);
            Util::check(qsv_apply_matrix(sv_, m.data(), nullptr, 0, w.data(), static_cast<int>(w.size()), adjoint));
            return;
        }
        const std::vector<double> p(params.begin(), params.end());
        Util::check(qsv_apply_named(sv_, opName.c_str(), w.data(), static_cast<int>(w.size()), adjoint, p.data(),
                                    static_cast<int>(p.size())));
    }
    void applyOperation_std(const std::string &opName, const std::vector<std::size_t> &wires, bool adjoint = false,
                            const std::vector<PrecisionT> &params = {0.0},
                            const std::vector<ComplexT> &gate_matrix = {}) {
        applyOperation(opName, wires, adjoint, params, gate_matrix);
    }
    // The vector overloads (Managed.hpp:279-315 in the reference: a loop over the single-gate form) hand the whole list to
    // the fused executor as ONE recorded circuit: consecutive gates share HBM sweeps.  The five-argument form adds the
    // matrices of operations without a named kernel, like create_ops_list (bindings/Bindings.cpp:806-840).
    void applyOperation(const std::vector<std::string> &ops, const std::vector<std::vector<std::size_t>> &wires,
                        const std::vector<bool> &adjoints, const std::vector<std::vector<PrecisionT>> &params,
                        const std::vector<std::vector<ComplexT>> &matrices) {
        PL_ABORT_IF(ops.size() != wires.size() || ops.size() != adjoints.size() || ops.size() != params.size() ||
                        (!matrices.empty() && matrices.size() != ops.size()),
                    "Invalid arguments: number of operations, wires, inverses and parameters must all be equal");
        qsv_ops *rec = nullptr;
        Util::check(qsv_ops_create(&rec));
        int st = 0;
        for (std::size_t i = 0; i < ops.size() && st == 0; ++i) {
            const std::vector<int> w(wires[i].begin(), wires[i].end());
            const std::vector<double> p(params[i].begin(), params[i].end());
            std::vector<double> m;
            if (!matrices.empty() && !matrices[i].empty()) m = to_doubles(matrices[i]);
            st = qsv_ops_append(rec, ops[i].c_str(), w.data(), static_cast<int>(w.size()), p.data(), static_cast<int>(p.size()),
                                adjoints[i], m.empty() ? nullptr : m.data(), m.empty() ? 0 : (std::size_t{1} << w.size()));
        }
        if (st == 0) st = qsv_apply_ops(sv_, rec, 1);
        qsv_ops_destroy(rec);
        Util::check(st);
    }
    void applyOperation(const std::vector<std::string> &ops, const std::vector<std::vector<std::size_t>> &wires,
                        const std::vector<bool> &adjoints, const std::vector<std::vector<PrecisionT>> &params) {
        applyOperation(ops, wires, adjoints, params, {});
    }
    void applyOperation(const std::vector<std::string> &ops, const std::vector<std::vector<std::size_t>> &wires,
                        const std::vector<bool> &adjoints) {
        applyOperation(ops, wires, adjoints, std::vector<std::vector<PrecisionT>>(ops.size()), {});
    }
    // matrix given explicitly (the reference's applyDeviceMatrixGate / applyHostMatrixGate pair)
    void applyMatrix(const std::vector<ComplexT> &matrix, const std::vector<std::size_t> &ctrls,
                     const std::vector<std::size_t> &tgts, bool adjoint = false) {
        const std::vector<int> c = to_int(ctrls), t = to_int(tgts);
        const std::vector<double> m = to_doubles(matrix);
        Util::check(qsv_apply_matrix(sv_, m.data(), c.data(), static_cast<int>(c.size()), t.data(),
                                     static_cast<int>(t.size()), adjoint));
    }

#define QSV_GATE0(NAME)                                                                            \
    void apply##NAME(const std::vector<std::size_t> &wires, bool adjoint) { applyOperation(#NAME, wires, adjoint, {}); }
#define QSV_GATE1(NAME)                                                                            \
    void apply##NAME(const std::vector<std::size_t> &wires, bool adjoint, PrecisionT param) {      \
        applyOperation(#NAME, wires, adjoint, {param});                                            \
    }
#define QSV_GATE3(NAME)                                                                            \
    void apply##NAME(const std::vector<std::size_t> &wires, bool adjoint, PrecisionT p0, PrecisionT p1,  \
                     PrecisionT p2) {                                                              \
        applyOperation(#NAME, wires, adjoint, {p0, p1, p2});                                       \
    }                                                                                              \
    void apply##NAME(const std::vector<std::size_t> &wires, bool adjoint, const std::vector<PrecisionT> &p) {  \
        applyOperation(#NAME, wires, adjoint, p);                                                  \
    }
    QSV_GATE0(Identity) QSV_GATE0(PauliX) QSV_GATE0(PauliY) QSV_GATE0(PauliZ) QSV_GATE0(Hadamard) QSV_GATE0(S)
    QSV_GATE0(T) QSV_GATE0(CNOT) QSV_GATE0(CY) QSV_GATE0(CZ) QSV_GATE0(SWAP) QSV_GATE0(Toffoli) QSV_GATE0(CSWAP)
    QSV_GATE1(RX) QSV_GATE1(RY) QSV_GATE1(RZ) QSV_GATE1(PhaseShift) QSV_GATE1(IsingXX) QSV_GATE1(IsingYY)
    QSV_GATE1(IsingZZ) QSV_GATE1(CRX) QSV_GATE1(CRY) QSV_GATE1(CRZ) QSV_GATE1(ControlledPhaseShift)
    QSV_GATE1(SingleExcitation) QSV_GATE1(SingleExcitationMinus) QSV_GATE1(SingleExcitationPlus)
    QSV_GATE1(DoubleExcitation) QSV_GATE1(DoubleExcitationMinus) QSV_GATE1(DoubleExcitationPlus) QSV_GATE1(MultiRZ)
    QSV_GATE3(Rot) QSV_GATE3(CRot)
#undef QSV_GATE0
#undef QSV_GATE1
#undef QSV_GATE3

    // state <- G state; returns the scaling factor of AdjointDiffGPU.hpp:96-114
    PrecisionT applyGenerator(const std::string &opName, const std::vector<std::size_t> &wires, bool adjoint = false) {
        const std::vector<int> w = to_int(wires);
        double scale = 0;
        Util::check(qsv_apply_generator(sv_, opName.c_str(), w.data(), static_cast<int>(w.size()), adjoint, &scale));
        return static_cast<PrecisionT>(scale);
    }
#define QSV_GEN(NAME)                                                                              \
    void applyGenerator##NAME(const std::vector<std::size_t> &wires, bool adjoint = false) {       \
        applyGenerator(#NAME, wires, adjoint);                                                     \
    }
    QSV_GEN(RX) QSV_GEN(RY) QSV_GEN(RZ) QSV_GEN(IsingXX) QSV_GEN(IsingYY) QSV_GEN(IsingZZ) QSV_GEN(PhaseShift)
    QSV_GEN(CRX) QSV_GEN(CRY) QSV_GEN(CRZ) QSV_GEN(ControlledPhaseShift) QSV_GEN(SingleExcitation)
    QSV_GEN(SingleExcitationMinus) QSV_GEN(SingleExcitationPlus) QSV_GEN(DoubleExcitation)
    QSV_GEN(DoubleExcitationMinus) QSV_GEN(DoubleExcitationPlus) QSV_GEN(MultiRZ)
#undef QSV_GEN

    // ---- measurements ---------------------------------------------------------------------------------
    // <psi|O|psi> for a named observable, or for `gate_matrix` on `wires` when the name is not a gate
    ComplexT expval(const std::string &obsName, const std::vector<std::size_t> &wires,
                    const std::vector<PrecisionT> &params = {0.0}, const std::vector<ComplexT> &gate_matrix = {}) {
        const std::vector<int> w = to_int(wires);
        double out[2] = {0, 0};
        if (is_named(obsName)) {
            const std::vector<double> p(params.begin(), params.end());
            Util::check(qsv_expval_named(sv_, obsName.c_str(), w.data(), static_cast<int>(w.size()), p.data(),
                                         static_cast<int>(p.size()), out));
        } else {
            PL_ABORT_IF(gate_matrix.empty(), std::string("Currently unsupported observable: ") + obsName);
            const std::vector<double> m = to_doubles(gate_matrix);
            Util::check(qsv_expval_matrix(sv_, m.data(), w.data(), static_cast<int>(w.size()), out));
        }
        return {static_cast<PrecisionT>(out[0]), static_cast<PrecisionT>(out[1])};
    }
    ComplexT expval(const std::vector<std::size_t> &wires, const std::vector<ComplexT> &gate_matrix) {
        const std::vector<int> w = to_int(wires);
        const std::vector<double> m = to_doubles(gate_matrix);
        double out[2] = {0, 0};
        Util::check(qsv_expval_matrix(sv_, m.data(), w.data(), static_cast<int>(w.size()), out));
        return {static_cast<PrecisionT>(out[0]), static_cast<PrecisionT>(out[1])};
    }
    template <class index_type>
    PrecisionT getExpectationValueOnSparseSpMV(const index_type *csrOffsets, index_type /*csrOffsets_size*/,
                                               const index_type *columns, const ComplexT *values, index_type numNNZ) {
        std::vector<double> v(2 * static_cast<std::size_t>(numNNZ));
        for (std::size_t i = 0; i < static_cast<std::size_t>(numNNZ); ++i) {
            v[2 * i] = values[i].real();
            v[2 * i + 1] = values[i].imag();
        }
        double out = 0;
        Util::check(qsv_expval_csr(sv_, csrOffsets, columns, v.data(), static_cast<int64_t>(numNNZ),
                                   static_cast<int>(sizeof(index_type)), &out));
        return static_cast<PrecisionT>(out);
    }
    PrecisionT getExpectationValuePauliWords(const std::vector<std::string> &pauli_words,
                                             const std::vector<std::vector<std::size_t>> &tgts,
                                             const ComplexT *coeffs) {
        std::string letters;
        std::vector<int> wires, offsets{0};
        std::vector<double> c;
        for (std::size_t t = 0; t < pauli_words.size(); ++t) {
            PL_ABORT_IF(pauli_words[t].size() != tgts[t].size(), "Pauli word and target wires differ in length");
            letters += pauli_words[t];
            for (auto w : tgts[t]) wires.push_back(static_cast<int>(w));
            offsets.push_back(static_cast<int>(letters.size()));
            c.push_back(coeffs[t].real());
            c.push_back(coeffs[t].imag());
        }
        double out = 0;
        Util::check(qsv_expval_pauli_words(sv_, static_cast<int>(pauli_words.size()), letters.c_str(), wires.data(),
                                           offsets.data(), c.data(), nullptr, &out));
        return static_cast<PrecisionT>(out);
    }
    std::vector<double> probability(const std::vector<std::size_t> &wires) {
        const std::vector<int> w = to_int(wires);
        std::vector<double> p(std::size_t{1} << w.size());
        Util::check(qsv_probs(sv_, w.data(), static_cast<int>(w.size()), p.data()));
        return p;
    }
    // shots x num_qubits matrix of 0/1; the uniform numbers come from a host mt19937 like the reference's
    // (Managed.hpp:1003-1007); pass a seed for reproducible samples
    std::vector<std::size_t> generate_samples(std::size_t num_samples) {
        return generate_samples(num_samples, std::random_device{}());
    }
    std::vector<std::size_t> generate_samples(std::size_t num_samples, std::uint64_t seed) {
        std::mt19937_64 gen(seed);
        std::uniform_real_distribution<double> dis(0.0, 1.0);
        std::vector<double> u(num_samples);
        for (auto &x : u) x = dis(gen);
        std::vector<uint64_t> out(num_samples * num_qubits_);
        Util::check(qsv_sample(sv_, u.data(), static_cast<int64_t>(num_samples), out.data()));
        return std::vector<std::size_t>(out.begin(), out.end());
    }

  private:
    void create() {
        PL_ABORT_IF(num_qubits_ == 0, "a state vector needs at least one qubit");
        Util::check(qsv_create_external(static_cast<int>(num_qubits_), dtype_code, dev_tag_.getDeviceID(), nullptr,
                                        dev_tag_.getStreamID(), &sv_));
    }
    static std::size_t log2_exact(std::size_t length) {
        std::size_t n = 0;
        while ((std::size_t{1} << n) < length) ++n;
        PL_ABORT_IF((std::size_t{1} << n) != length, "state-vector length must be a power of two");
        return n;
    }
    static std::vector<int> to_int(const std::vector<std::size_t> &w) { return std::vector<int>(w.begin(), w.end()); }
    static std::vector<double> to_doubles(const std::vector<ComplexT> &m) {
        std::vector<double> d(2 * m.size());
        for (std::size_t i = 0; i < m.size(); ++i) {
            d[2 * i] = m[i].real();
            d[2 * i + 1] = m[i].imag();
        }
        return d;
    }
    static bool is_named(const std::string &name) {
        static const char *names[] = {"Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "T", "RX", "RY", "RZ",
                                      "PhaseShift", "Rot", "CNOT", "CY", "CZ", "SWAP", "IsingXX", "IsingYY", "IsingZZ",
                                      "CRX", "CRY", "CRZ", "CRot", "ControlledPhaseShift", "SingleExcitation",
                                      "SingleExcitationMinus", "SingleExcitationPlus", "Toffoli", "CSWAP",
                                      "DoubleExcitation", "DoubleExcitationMinus", "DoubleExcitationPlus", "MultiRZ"};
        for (const char *n : names)
            if (name == n) return true;
        return false;
    }

    qsv_state *sv_{nullptr};
    std::size_t num_qubits_;
    CUDA::DevTag<int> dev_tag_;
};

}  // namespace Pennylane
