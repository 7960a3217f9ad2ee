// AdjointJacobianGPU<T> and OpsData<SV> with the interface of the reference's
// algorithms/AdjointDiffGPU.hpp:392-596 and pennylane-lightning's JacobianData.hpp (OpsData), over
// qsv_adjoint_jacobian.  The reverse sweep itself runs inside libqsv_b200.so (csrc/circuit.cu).
#pragma once
#include <complex>
#include <algorithm>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "Error.hpp"
#include "ObservablesGPU.hpp"
#include "StateVectorCudaManaged.hpp"
#include "qsv_b200.h"

namespace Pennylane::Algorithms {

// Recorded operations of a tape: names, parameters, wires, inverse flags and (for ops without a
// dedicated kernel) matrices.  num_par_ops = number of ops with a non-empty parameter list.
template <class SVType> class OpsData {
  public:
    using PrecisionT = typename SVType::Precision;
    OpsData(const std::vector<std::string> &ops_name, const std::vector<std::vector<PrecisionT>> &ops_params,
            const std::vector<std::vector<std::size_t>> &ops_wires, const std::vector<bool> &ops_inverses,
            const std::vector<std::vector<std::complex<PrecisionT>>> &ops_matrices)
        : names_(ops_name), params_(ops_params), wires_(ops_wires), inverses_(ops_inverses), matrices_(ops_matrices) {
        PL_ABORT_IF(names_.size() != params_.size() || names_.size() != wires_.size() ||
                        names_.size() != inverses_.size(),
                    "operation names, parameters, wires and inverses must have the same length");
        matrices_.resize(names_.size());
        Util::check(qsv_ops_create(&handle_));
        for (std::size_t i = 0; i < names_.size(); ++i) {
            num_par_ops_ += params_[i].empty() ? 0 : 1;
            num_nonpar_ops_ += params_[i].empty() ? 1 : 0;
            const std::vector<int> w(wires_[i].begin(), wires_[i].end());
            const std::vector<double> p(params_[i].begin(), params_[i].end());
            std::vector<double> m(2 * matrices_[i].size());
            for (std::size_t j = 0; j < matrices_[i].size(); ++j) {
                m[2 * j] = matrices_[i][j].real();
                m[2 * j + 1] = matrices_[i][j].imag();
            }
            Util::check(qsv_ops_append(handle_, names_[i].c_str(), w.data(), static_cast<int>(w.size()), p.data(),
                                       static_cast<int>(p.size()), inverses_[i], m.empty() ? nullptr : m.data(),
                                       m.empty() ? 0 : (std::size_t{1} << w.size())));
        }
    }
    OpsData(const std::vector<std::string> &ops_name, const std::vector<std::vector<PrecisionT>> &ops_params,
            const std::vector<std::vector<std::size_t>> &ops_wires, const std::vector<bool> &ops_inverses)
        : OpsData(ops_name, ops_params, ops_wires, ops_inverses, {}) {}
    OpsData(const OpsData &o) : OpsData(o.names_, o.params_, o.wires_, o.inverses_, o.matrices_) {}
    OpsData &operator=(const OpsData &) = delete;
    ~OpsData() {
        if (handle_) qsv_ops_destroy(handle_);
    }
    std::size_t getSize() const { return names_.size(); }
    const std::vector<std::string> &getOpsName() const { return names_; }
    const std::vector<std::vector<PrecisionT>> &getOpsParams() const { return params_; }
    const std::vector<std::vector<std::size_t>> &getOpsWires() const { return wires_; }
    const std::vector<bool> &getOpsInverses() const { return inverses_; }
    const std::vector<std::vector<std::complex<PrecisionT>>> &getOpsMatrices() const { return matrices_; }
    bool hasParams(std::size_t index) const { return !params_[index].empty(); }
    std::size_t getNumParOps() const { return num_par_ops_; }
    std::size_t getNumNonParOps() const { return num_nonpar_ops_; }
    std::size_t getTotalNumParams() const {
        std::size_t n = 0;
        for (const auto &p : params_) n += p.size();
        return n;
    }
    qsv_ops *handle() const { return handle_; }

  private:
    std::vector<std::string> names_;
    std::vector<std::vector<PrecisionT>> params_;
    std::vector<std::vector<std::size_t>> wires_;
    std::vector<bool> inverses_;
    std::vector<std::vector<std::complex<PrecisionT>>> matrices_;
    std::size_t num_par_ops_{0}, num_nonpar_ops_{0};
    qsv_ops *handle_{nullptr};
};

template <class T = double> class AdjointJacobianGPU {
  public:
    using SV = StateVectorCudaManaged<T>;
    using ObsPtr = std::shared_ptr<ObservableGPU<T>>;

    // jac[obs][param]; `sv` holds the final state (or the initial one with apply_operations = true)
    void adjointJacobian(const SV &sv, std::vector<std::vector<T>> &jac, const std::vector<ObsPtr> &obs,
                         const OpsData<SV> &ops, const std::vector<std::size_t> &trainableParams,
                         bool apply_operations = false) {
        PL_ABORT_IF(trainableParams.empty(), "No trainable parameters provided.");
        std::vector<qsv_obs *> hs;
        for (const auto &o : obs) hs.push_back(o->handle());
        const std::vector<int64_t> tp(trainableParams.begin(), trainableParams.end());
        std::vector<double> flat(obs.size() * tp.size(), 0.0);
        Util::check(qsv_adjoint_jacobian(sv.handle(), ops.handle(), hs.data(), static_cast<int>(hs.size()), tp.data(),
                                         static_cast<int>(tp.size()), apply_operations, flat.data()));
        jac.assign(obs.size(), std::vector<T>(tp.size(), 0));
        for (std::size_t i = 0; i < obs.size(); ++i)
            for (std::size_t p = 0; p < tp.size(); ++p) jac[i][p] = static_cast<T>(flat[i * tp.size() + p]);
    }
    // Observable batching over all visible GPUs (AdjointDiffGPU.hpp:392-473): replicas, no communication.  The
    // observables are split into contiguous chunks (ceil division, :416-419), one std::thread per GPU copies the
    // state to its device (peer copy) and runs the sweep on its chunk.  On one GPU this is adjointJacobian.
    void batchAdjointJacobian(const SV &sv, std::vector<std::vector<T>> &jac, const std::vector<ObsPtr> &obs,
                              const OpsData<SV> &ops, const std::vector<std::size_t> &trainableParams,
                              bool apply_operations = false) {
        int n_dev = 0;
        Util::check(qsv_device_count(&n_dev));
        const std::size_t n_chunks = std::min<std::size_t>(static_cast<std::size_t>(std::max(n_dev, 1)), obs.size());
        if (n_chunks <= 1) {
            adjointJacobian(sv, jac, obs, ops, trainableParams, apply_operations);
            return;
        }
        PL_ABORT_IF(trainableParams.empty(), "No trainable parameters provided.");
        jac.assign(obs.size(), std::vector<T>(trainableParams.size(), 0));
        const std::size_t per = (obs.size() + n_chunks - 1) / n_chunks;
        std::vector<std::thread> threads;
        std::vector<std::string> errors(n_chunks);
        for (std::size_t c = 0; c < n_chunks; ++c) {
            const std::size_t first = c * per, last = std::min(obs.size(), first + per);
            if (first >= last) break;
            threads.emplace_back([&, c, first, last]() {
                try {
                    const int dev = static_cast<int>(c);
                    std::vector<ObsPtr> mine(obs.begin() + first, obs.begin() + last);
                    std::vector<std::vector<T>> part;
                    if (dev == sv.getDevTag().getDeviceID()) {
                        adjointJacobian(sv, part, mine, ops, trainableParams, apply_operations);
                    } else {
                        SV local(sv.getNumQubits(), CUDA::DevTag<int>(dev));
                        Util::check(qsv_d2d(local.handle(), sv.handle()));
                        adjointJacobian(local, part, mine, ops, trainableParams, apply_operations);
                    }
                    for (std::size_t i = 0; i < part.size(); ++i) jac[first + i] = part[i];
                } catch (const std::exception &e) {
                    errors[c] = e.what();
                }
            });
        }
        for (auto &t : threads) t.join();
        for (const auto &e : errors) PL_ABORT_IF(!e.empty(), e);
    }
};

}  // namespace Pennylane::Algorithms
