// ObservableGPUMPI<T> hierarchy with the interface of the reference's algorithms/ObservablesGPUMPI.hpp
// (applyInPlace on a StateVectorCudaMPI, getObsName, getWires, operator==).  Each object wraps the single-GPU
// observable of the same kind: the observable record (qsv_obs) is the same, only the entry points that evaluate
// it on a sharded register differ (qsv_dist_obs_apply / qsv_dist_obs_expval).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "ObservablesGPU.hpp"
#include "StateVectorCudaMPI.hpp"

namespace Pennylane::Algorithms {

template <typename T> class ObservableGPUMPI {
  public:
    virtual ~ObservableGPUMPI() = default;
    // sv <- O sv  (ObservablesGPUMPI.hpp:57)
    void applyInPlace(StateVectorCudaMPI<T> &sv) const { Util::check(qsv_dist_obs_apply(inner_->handle(), sv.handle())); }
    // Re <sv|O|sv>, identical on all ranks
    T expval(StateVectorCudaMPI<T> &sv) const {
        double out = 0;
        Util::check(qsv_dist_obs_expval(inner_->handle(), sv.handle(), &out));
        return static_cast<T>(out);
    }
    std::string getObsName() const { return inner_->getObsName(); }
    std::vector<std::size_t> getWires() const { return inner_->getWires(); }
    bool operator==(const ObservableGPUMPI<T> &other) const { return *inner_ == *other.inner_; }
    bool operator!=(const ObservableGPUMPI<T> &other) const { return !(*this == other); }
    qsv_obs *handle() const { return inner_->handle(); }
    const std::shared_ptr<ObservableGPU<T>> &inner() const { return inner_; }

  protected:
    ObservableGPUMPI() = default;
    std::shared_ptr<ObservableGPU<T>> inner_;
};

template <typename T> class NamedObsGPUMPI final : public ObservableGPUMPI<T> {
  public:
    NamedObsGPUMPI(std::string obs_name, std::vector<std::size_t> wires, std::vector<T> params = {}) {
        this->inner_ = std::make_shared<NamedObsGPU<T>>(std::move(obs_name), std::move(wires), std::move(params));
    }
};

template <typename T> class HermitianObsGPUMPI final : public ObservableGPUMPI<T> {
  public:
    using MatrixT = std::vector<std::complex<T>>;
    HermitianObsGPUMPI(MatrixT matrix, std::vector<std::size_t> wires) {
        this->inner_ = std::make_shared<HermitianObsGPU<T>>(std::move(matrix), std::move(wires));
    }
};

template <typename T> class TensorProdObsGPUMPI final : public ObservableGPUMPI<T> {
  public:
    using ObsPtr = std::shared_ptr<ObservableGPUMPI<T>>;
    explicit TensorProdObsGPUMPI(const std::vector<ObsPtr> &obs) {
        std::vector<std::shared_ptr<ObservableGPU<T>>> in;
        for (const auto &o : obs) in.push_back(o->inner());
        this->inner_ = std::make_shared<TensorProdObsGPU<T>>(in);
    }
    static auto create(std::vector<ObsPtr> obs) { return std::make_shared<TensorProdObsGPUMPI<T>>(obs); }
};

template <typename T> class HamiltonianGPUMPI final : public ObservableGPUMPI<T> {
  public:
    using ObsPtr = std::shared_ptr<ObservableGPUMPI<T>>;
    HamiltonianGPUMPI(const std::vector<T> &coeffs, const std::vector<ObsPtr> &obs) {
        std::vector<std::shared_ptr<ObservableGPU<T>>> in;
        for (const auto &o : obs) in.push_back(o->inner());
        this->inner_ = std::make_shared<HamiltonianGPU<T>>(coeffs, in);
    }
    static auto create(std::vector<T> coeffs, std::vector<ObsPtr> obs) {
        return std::make_shared<HamiltonianGPUMPI<T>>(coeffs, obs);
    }
};

template <typename T> class SparseHamiltonianGPUMPI final : public ObservableGPUMPI<T> {
  public:
    using IdxT = typename SparseHamiltonianGPU<T>::IdxT;
    SparseHamiltonianGPUMPI(std::vector<std::complex<T>> data, std::vector<IdxT> indices, std::vector<IdxT> offsets,
                            std::vector<std::size_t> wires) {
        this->inner_ = std::make_shared<SparseHamiltonianGPU<T>>(std::move(data), std::move(indices), std::move(offsets),
                                                                 std::move(wires));
    }
};

}  // namespace Pennylane::Algorithms
