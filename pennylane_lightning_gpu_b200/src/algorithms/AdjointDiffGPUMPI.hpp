// AdjointJacobianGPUMPI<T> with the interface of the reference's algorithms/AdjointDiffGPUMPI.hpp:248-437 over
// qsv_dist_adjoint_jacobian: lambda and the bras are sharded like the register and follow its exchanges inside
// libqsv_b200.so (csrc/dist.cu); the Jacobian is all-reduced once and is identical on all ranks.
#pragma once
#include <memory>
#include <vector>

#include "AdjointDiffGPU.hpp"
#include "ObservablesGPUMPI.hpp"
#include "StateVectorCudaMPI.hpp"

namespace Pennylane::Algorithms {

template <class T = double> class AdjointJacobianGPUMPI {
  public:
    using SV = StateVectorCudaMPI<T>;
    using ObsPtr = std::shared_ptr<ObservableGPUMPI<T>>;

    void adjointJacobian(const SV &sv, std::vector<std::vector<T>> &jac, const std::vector<ObsPtr> &obs,
                         const OpsData<SV> &ops, const std::vector<std::size_t> &trainableParams,
                         bool apply_operations = false) {
        PL_ABORT_IF(trainableParams.empty(), "No trainable parameters provided.");
        std::vector<qsv_obs *> hs;
        for (const auto &o : obs) hs.push_back(o->handle());
        const std::vector<int64_t> tp(trainableParams.begin(), trainableParams.end());
        std::vector<double> flat(obs.size() * tp.size(), 0.0);
        Util::check(qsv_dist_adjoint_jacobian(sv.handle(), ops.handle(), hs.data(), static_cast<int>(hs.size()),
                                              tp.data(), static_cast<int>(tp.size()), apply_operations, flat.data()));
        jac.assign(obs.size(), std::vector<T>(tp.size(), 0));
        for (std::size_t i = 0; i < obs.size(); ++i)
            for (std::size_t p = 0; p < tp.size(); ++p) jac[i][p] = static_cast<T>(flat[i * tp.size() + p]);
    }
    // the reference's memory-saving variant (one observable at a time, AdjointDiffGPUMPI.hpp:340-437)
    void adjointJacobian_serial(const SV &sv, std::vector<std::vector<T>> &jac, const std::vector<ObsPtr> &obs,
                                const OpsData<SV> &ops, const std::vector<std::size_t> &trainableParams,
                                bool apply_operations = false) {
        jac.clear();
        for (const auto &o : obs) {
            std::vector<std::vector<T>> row;
            adjointJacobian(sv, row, {o}, ops, trainableParams, apply_operations);
            jac.push_back(row[0]);
        }
    }
};

}  // namespace Pennylane::Algorithms
