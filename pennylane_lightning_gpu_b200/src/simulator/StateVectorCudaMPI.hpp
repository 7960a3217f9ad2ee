// StateVectorCudaMPI<PrecisionT>: the reference's sharded state-vector class
// (pennylane_lightning_gpu/src/simulator/StateVectorCudaMPI.hpp) with its public method names and argument
// meanings, as a thin shell over the qsv_dist_* entries of libqsv_b200.so.  One process per GPU; the top
// num_global_qubits index bits are the rank (MPI.hpp:240-246).  No MPI, no cuStateVec: the exchange is NCCL /
// direct NVLink peer access inside the library (csrc/dist.cu).
//
//   reference method (file:line)                                   C-ABI call
//   ctor (MPI.hpp:100-215)                                         qsv_create + qsv_dist_init
//   setBasisState / setStateVector (:332-381), initSV_MPI          qsv_dist_set_basis_state / _set_state_vector
//   applyOperation + named apply* (:392-900, :2023-2587)           qsv_dist_apply_ops (lazy qubit map)
//   expval x3 (:957-1035)                                          qsv_dist_expval_named / _matrix
//   getExpectationValueOnSparseSpMV (:1050-1176)                   qsv_dist_expval_csr
//   getExpectationValuePauliWords (:1296-1445)                     qsv_dist_expval_pauli_words
//   probability (:1187-1290), generate_samples (:1454-1595)        qsv_dist_probs / qsv_dist_sample
//   CopyHostDataToGpu / CopyGpuDataToHost / updateData             qsv_dist_h2d / _d2h / _copy
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
#include "MPIManager.hpp"
#include "qsv_b200.h"

namespace Pennylane {

template <class PrecisionT> class StateVectorCudaMPI {
  public:
    using Precision = PrecisionT;
    using ComplexT = std::complex<PrecisionT>;
    using CFP_t = ComplexT;
    static constexpr int dtype_code = std::is_same_v<PrecisionT, double> ? QSV_C128 : QSV_C64;

    StateVectorCudaMPI() = delete;
    // mpi_buf_size: size in MiB of the staging buffers of the NCCL send/recv fallback (the reference's transfer
    // workspace, MPIWorker.hpp:276-294); 0 = default.  The direct peer-access exchange needs no buffer.
    StateVectorCudaMPI(MPI::MPIManager &mpi_manager, const CUDA::DevTag<int> &dev_tag, std::size_t mpi_buf_size,
                       std::size_t num_global_qubits, std::size_t num_local_qubits)
        : mpi_manager_(&mpi_manager), dev_tag_(dev_tag), mpi_buf_size_(mpi_buf_size),
          num_global_qubits_(num_global_qubits), num_local_qubits_(num_local_qubits) {
        PL_ABORT_IF((std::size_t{1} << num_global_qubits) != static_cast<std::size_t>(mpi_manager.getSize()),
                    "number of global qubits does not match the number of processes");
        Util::check(qsv_create_external(static_cast<int>(num_local_qubits_), dtype_code, dev_tag_.getDeviceID(), nullptr,
                                        dev_tag_.getStreamID(), &sv_));
        const auto id = mpi_manager.newUniqueId();
        Util::check(qsv_dist_init(sv_, id.data(), mpi_manager.getRank(), mpi_manager.getSize()));
        initSV_MPI();
    }
    StateVectorCudaMPI(const StateVectorCudaMPI &other)
        : StateVectorCudaMPI(*other.mpi_manager_, other.dev_tag_, other.mpi_buf_size_, other.num_global_qubits_,
                             other.num_local_qubits_) {
        Util::check(qsv_dist_copy(sv_, other.sv_));
    }
    StateVectorCudaMPI &operator=(const StateVectorCudaMPI &) = delete;
    ~StateVectorCudaMPI() {
        if (sv_) {
            qsv_dist_finalize(sv_);
            qsv_destroy(sv_);
        }
    }

    // ---- sizes / handles -----------------------------------------------------------------------
    std::size_t getNumGlobalQubits() const { return num_global_qubits_; }
    std::size_t getNumLocalQubits() const { return num_local_qubits_; }
    std::size_t getTotalNumQubits() const { return num_global_qubits_ + num_local_qubits_; }
    std::size_t getNumQubits() const { return num_local_qubits_; }
    std::size_t getLength() const { return std::size_t{1} << num_local_qubits_; }
    qsv_state *handle() const { return sv_; }
    const CUDA::DevTag<int> &getDevTag() const { return dev_tag_; }
    MPI::MPIManager &getMPIManager() const { return *mpi_manager_; }
    void *getData() const { return qsv_data_ptr(sv_); }
    bool usesPeerAccess() const { return qsv_dist_uses_peer_access(sv_) != 0; }

    // ---- initialisation and copies ---------------------------------------------------------------
    void initSV_MPI(bool /*async*/ = false) { Util::check(qsv_dist_set_basis_state(sv_, 0)); }
    void setBasisState(const ComplexT & /*value*/, std::size_t index, bool /*async*/ = false) {
        Util::check(qsv_dist_set_basis_state(sv_, index));
    }
    template <class index_type>
    void setStateVector(index_type num_indices, const ComplexT *values, const index_type *indices,
                        bool /*async*/ = false) {
        std::vector<int64_t> idx(indices, indices + num_indices);
        Util::check(qsv_dist_set_state_vector(sv_, idx.data(), values, static_cast<std::size_t>(num_indices)));
    }
    void CopyHostDataToGpu(const ComplexT *host, std::size_t length, bool /*async*/ = false) {
        PL_ABORT_IF_NOT(getLength() == length, "Sizes do not match for Host and GPU data");
        Util::check(qsv_dist_h2d(sv_, host, length));
    }
    void CopyHostDataToGpu(const std::vector<ComplexT> &sv, bool async = false) {
        CopyHostDataToGpu(sv.data(), sv.size(), async);
    }
    void CopyGpuDataToHost(ComplexT *host, std::size_t length, bool /*async*/ = false) const {
        PL_ABORT_IF_NOT(getLength() == length, "Sizes do not match for Host and GPU data");
        Util::check(qsv_dist_d2h(sv_, host, length));
    }
    void updateData(const StateVectorCudaMPI &other, bool /*async*/ = false) {
        PL_ABORT_IF_NOT(getLength() == other.getLength(), "Sizes do not match for GPU data");
        Util::check(qsv_dist_copy(sv_, other.sv_));
    }

    // ---- gates: wires address the whole register ----------------------------------------------------
    void applyOperation(const std::string &opName, const std::vector<std::size_t> &wires, bool adjoint = false,
                        const std::vector<PrecisionT> &params = {0.0},
                        const std::vector<ComplexT> &gate_matrix = {}) {
        qsv_ops *ops = nullptr;
        Util::check(qsv_ops_create(&ops));
        const std::vector<int> w(wires.begin(), wires.end());
        const std::vector<double> p(params.begin(), params.end());
        const std::vector<double> m = to_doubles(gate_matrix);
        int st = qsv_ops_append(ops, opName.c_str(), w.data(), static_cast<int>(w.size()), p.data(),
                                static_cast<int>(p.size()), adjoint, m.empty() ? nullptr : m.data(),
                                m.empty() ? 0 : (std::size_t{1} << w.size()));
        if (st == 0) st = qsv_dist_apply_ops(sv_, ops, 0, chunk_bytes());
        qsv_ops_destroy(ops);
        Util::check(st);
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
        if (st == 0) st = qsv_dist_apply_ops(sv_, rec, 1, chunk_bytes());
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
    // a whole recorded circuit in one call: lazy qubit map + fused local sweeps between the exchanges
    void applyOperations(qsv_ops *ops, bool fuse = true) { Util::check(qsv_dist_apply_ops(sv_, ops, fuse, chunk_bytes())); }

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

    // ---- measurements: results are identical on all ranks -------------------------------------------
    ComplexT expval(const std::string &obsName, const std::vector<std::size_t> &wires,
                    const std::vector<PrecisionT> &params = {0.0}, const std::vector<ComplexT> &gate_matrix = {}) {
        const std::vector<int> w(wires.begin(), wires.end());
        double out[2] = {0, 0};
        if (gate_matrix.empty()) {
            const std::vector<double> p(params.begin(), params.end());
            Util::check(qsv_dist_expval_named(sv_, obsName.c_str(), w.data(), static_cast<int>(w.size()), p.data(),
                                              static_cast<int>(p.size()), out));
        } else {
            const std::vector<double> m = to_doubles(gate_matrix);
            Util::check(qsv_dist_expval_matrix(sv_, m.data(), w.data(), static_cast<int>(w.size()), out));
        }
        return {static_cast<PrecisionT>(out[0]), static_cast<PrecisionT>(out[1])};
    }
    ComplexT expval(const std::vector<std::size_t> &wires, const std::vector<ComplexT> &gate_matrix) {
        return expval("", wires, {}, gate_matrix);
    }
    // Every rank passes the whole CSR matrix, or only rank 0 does (the reference's Python passes an identity
    // placeholder on the other ranks, lightning_gpu.py:840-852): sizes and arrays are then broadcast from rank 0.
    template <class index_type>
    PrecisionT getExpectationValueOnSparseSpMV(const index_type *csrOffsets, index_type csrOffsets_size,
                                               const index_type *columns, const ComplexT *values, index_type numNNZ) {
        const std::size_t rows = std::size_t{1} << getTotalNumQubits();
        int64_t have = static_cast<std::size_t>(csrOffsets_size) == rows + 1 ? 1 : 0;
        int64_t meta[2] = {have, static_cast<int64_t>(numNNZ)};
        mpi_manager_->Bcast(meta, 2, 0);
        PL_ABORT_IF(meta[0] != 1, "the sparse Hamiltonian on rank 0 does not match the size of the register");
        const std::size_t nnz = static_cast<std::size_t>(meta[1]);
        std::vector<int64_t> offs(rows + 1), cols(nnz);
        std::vector<double> v(2 * nnz);
        if (mpi_manager_->getRank() == 0) {
            for (std::size_t i = 0; i <= rows; ++i) offs[i] = static_cast<int64_t>(csrOffsets[i]);
            for (std::size_t i = 0; i < nnz; ++i) {
                cols[i] = static_cast<int64_t>(columns[i]);
                v[2 * i] = values[i].real();
                v[2 * i + 1] = values[i].imag();
            }
        }
        mpi_manager_->Bcast(offs, 0);
        mpi_manager_->Bcast(cols, 0);
        mpi_manager_->Bcast(v, 0);
        double out = 0;
        Util::check(qsv_dist_expval_csr(sv_, offs.data(), cols.data(), v.data(), static_cast<int64_t>(nnz), &out));
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
        Util::check(qsv_dist_expval_pauli_words(sv_, static_cast<int>(pauli_words.size()), letters.c_str(), wires.data(),
                                                offsets.data(), c.data(), nullptr, &out));
        return static_cast<PrecisionT>(out);
    }
    std::vector<double> probability(const std::vector<std::size_t> &wires) {
        const std::vector<int> w(wires.begin(), wires.end());
        std::vector<double> p(std::size_t{1} << w.size());
        Util::check(qsv_dist_probs(sv_, w.data(), static_cast<int>(w.size()), p.data()));
        return p;
    }
    // shots x total-qubits matrix of 0/1, identical on all ranks: the uniform numbers are drawn on rank 0 (the
    // reference's unseeded host generator, Managed.hpp:1003) and broadcast
    std::vector<std::size_t> generate_samples(std::size_t num_samples) {
        uint64_t seed = std::random_device{}();
        mpi_manager_->Bcast(&seed, 1, 0);
        return generate_samples(num_samples, seed);
    }
    std::vector<std::size_t> generate_samples(std::size_t num_samples, std::uint64_t seed) {
        std::mt19937_64 gen(seed);
        std::uniform_real_distribution<double> dis(0.0, 1.0);
        std::vector<double> u(num_samples);
        for (auto &x : u) x = dis(gen);
        std::vector<uint64_t> out(num_samples * getTotalNumQubits());
        Util::check(qsv_dist_sample(sv_, u.data(), static_cast<int64_t>(num_samples), out.data()));
        return std::vector<std::size_t>(out.begin(), out.end());
    }
    // NVLink bytes sent by this rank and device milliseconds of all exchanges so far
    void getSwapStatistics(int &n_swaps, uint64_t &bytes_sent, float &ms, bool reset = false) {
        Util::check(qsv_dist_total_swap_stats(sv_, &n_swaps, &bytes_sent, &ms, reset));
    }

  private:
    std::size_t chunk_bytes() const { return mpi_buf_size_ << 20; }
    static std::vector<double> to_doubles(const std::vector<ComplexT> &m) {
        std::vector<double> d(2 * m.size());
        for (std::size_t i = 0; i < m.size(); ++i) {
            d[2 * i] = m[i].real();
            d[2 * i + 1] = m[i].imag();
        }
        return d;
    }

    qsv_state *sv_{nullptr};
    MPI::MPIManager *mpi_manager_;
    CUDA::DevTag<int> dev_tag_;
    std::size_t mpi_buf_size_;
    std::size_t num_global_qubits_, num_local_qubits_;
};

}  // namespace Pennylane
