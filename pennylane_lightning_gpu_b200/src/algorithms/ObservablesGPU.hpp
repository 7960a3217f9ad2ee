// ObservableGPU<T> hierarchy with the interface of the reference's algorithms/ObservablesGPU.hpp:36-612
// (applyInPlace, getObsName, getWires, operator==); each object owns a qsv_obs handle of the C ABI.
#pragma once
#include <algorithm>
#include <complex>
#include <functional>
#include <memory>
#include <sstream>
#include <string>
#include <unordered_set>
#include <vector>

#include "Error.hpp"
#include "StateVectorCudaManaged.hpp"
#include "qsv_b200.h"

namespace Pennylane::Algorithms {

template <typename T> class ObservableGPU {
  public:
    virtual ~ObservableGPU() {
        if (handle_) qsv_obs_destroy(handle_);
    }
    ObservableGPU(const ObservableGPU &) = delete;
    ObservableGPU &operator=(const ObservableGPU &) = delete;

    // sv <- O sv   (ObservablesGPU.hpp:56)
    virtual void applyInPlace(StateVectorCudaManaged<T> &sv) const { Util::check(qsv_obs_apply(handle_, sv.handle())); }
    // Re <sv|O|sv> without touching sv
    T expval(StateVectorCudaManaged<T> &sv) const {
        double out = 0;
        Util::check(qsv_obs_expval(handle_, sv.handle(), &out));
        return static_cast<T>(out);
    }
    virtual std::string getObsName() const = 0;
    virtual std::vector<std::size_t> getWires() const = 0;
    bool operator==(const ObservableGPU<T> &other) const {
        return typeid(*this) == typeid(other) && isEqual(other);
    }
    bool operator!=(const ObservableGPU<T> &other) const { return !(*this == other); }
    qsv_obs *handle() const { return handle_; }

  protected:
    ObservableGPU() = default;
    virtual bool isEqual(const ObservableGPU<T> &other) const = 0;
    qsv_obs *handle_{nullptr};
};

template <typename T> class NamedObsGPU final : public ObservableGPU<T> {
  public:
    NamedObsGPU(std::string obs_name, std::vector<std::size_t> wires, std::vector<T> params = {})
        : obs_name_(std::move(obs_name)), wires_(std::move(wires)), params_(std::move(params)) {
        const std::vector<int> w(wires_.begin(), wires_.end());
        const std::vector<double> p(params_.begin(), params_.end());
        Util::check(qsv_obs_named(obs_name_.c_str(), w.data(), static_cast<int>(w.size()), p.data(),
                                  static_cast<int>(p.size()), &this->handle_));
    }
    std::string getObsName() const override {
        std::ostringstream s;
        s << obs_name_ << "[";
        for (std::size_t i = 0; i < wires_.size(); ++i) s << (i ? ", " : "") << wires_[i];
        s << "]";
        return s.str();
    }
    std::vector<std::size_t> getWires() const override { return wires_; }

  private:
    bool isEqual(const ObservableGPU<T> &other) const override {
        const auto &o = static_cast<const NamedObsGPU<T> &>(other);
        return obs_name_ == o.obs_name_ && wires_ == o.wires_ && params_ == o.params_;
    }
    std::string obs_name_;
    std::vector<std::size_t> wires_;
    std::vector<T> params_;
};

template <typename T> class HermitianObsGPU final : public ObservableGPU<T> {
  public:
    using MatrixT = std::vector<std::complex<T>>;
    HermitianObsGPU(MatrixT matrix, std::vector<std::size_t> wires) : matrix_(std::move(matrix)), wires_(std::move(wires)) {
        const std::size_t dim = std::size_t{1} << wires_.size();
        PL_ABORT_IF(matrix_.size() != dim * dim, "The matrix size does not match the number of wires");
        std::vector<double> m(2 * matrix_.size());
        for (std::size_t i = 0; i < matrix_.size(); ++i) {
            m[2 * i] = matrix_[i].real();
            m[2 * i + 1] = matrix_[i].imag();
        }
        const std::vector<int> w(wires_.begin(), wires_.end());
        Util::check(qsv_obs_hermitian(m.data(), dim, w.data(), static_cast<int>(w.size()), &this->handle_));
    }
    const MatrixT &getMatrix() const { return matrix_; }
    std::string getObsName() const override {
        std::size_t h = 0;
        for (const auto &c : matrix_) {
            h ^= std::hash<T>()(c.real()) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
            h ^= std::hash<T>()(c.imag()) + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
        }
        return "Hermitian" + std::to_string(h);
    }
    std::vector<std::size_t> getWires() const override { return wires_; }

  private:
    bool isEqual(const ObservableGPU<T> &other) const override {
        const auto &o = static_cast<const HermitianObsGPU<T> &>(other);
        return matrix_ == o.matrix_ && wires_ == o.wires_;
    }
    MatrixT matrix_;
    std::vector<std::size_t> wires_;
};

template <typename T> class TensorProdObsGPU final : public ObservableGPU<T> {
  public:
    using ObsPtr = std::shared_ptr<ObservableGPU<T>>;
    template <typename... Ts> explicit TensorProdObsGPU(Ts &&...args) : obs_{std::forward<Ts>(args)...} { build(); }
    static auto create(std::initializer_list<ObsPtr> obs) { return std::make_shared<TensorProdObsGPU<T>>(std::vector<ObsPtr>(obs)); }
    static auto create(std::vector<ObsPtr> obs) { return std::make_shared<TensorProdObsGPU<T>>(std::move(obs)); }
    std::size_t getSize() const { return obs_.size(); }
    std::vector<std::size_t> getWires() const override { return all_wires_; }
    std::string getObsName() const override {
        std::ostringstream s;
        for (std::size_t i = 0; i < obs_.size(); ++i) s << (i ? " @ " : "") << obs_[i]->getObsName();
        return s.str();
    }

  private:
    void build() {
        std::unordered_set<std::size_t> seen;
        std::vector<qsv_obs *> hs;
        for (const auto &o : obs_) {
            for (auto w : o->getWires()) {
                PL_ABORT_IF(seen.count(w) != 0, "All wires in observables must be disjoint.");
                seen.insert(w);
            }
            hs.push_back(o->handle());
        }
        all_wires_.assign(seen.begin(), seen.end());
        std::sort(all_wires_.begin(), all_wires_.end());
        Util::check(qsv_obs_tensor(hs.data(), static_cast<int>(hs.size()), &this->handle_));
    }
    bool isEqual(const ObservableGPU<T> &other) const override {
        const auto &o = static_cast<const TensorProdObsGPU<T> &>(other);
        if (obs_.size() != o.obs_.size()) return false;
        for (std::size_t i = 0; i < obs_.size(); ++i)
            if (*obs_[i] != *o.obs_[i]) return false;
        return true;
    }
    std::vector<ObsPtr> obs_;
    std::vector<std::size_t> all_wires_;
};

template <typename T> class HamiltonianGPU final : public ObservableGPU<T> {
  public:
    using ObsPtr = std::shared_ptr<ObservableGPU<T>>;
    template <typename T1, typename T2> HamiltonianGPU(T1 &&coeffs, T2 &&obs) : coeffs_{std::forward<T1>(coeffs)}, obs_{std::forward<T2>(obs)} {
        PL_ABORT_IF(coeffs_.size() != obs_.size(), "number of coefficients and observables must match");
        std::vector<double> c(coeffs_.begin(), coeffs_.end());
        std::vector<qsv_obs *> hs;
        for (const auto &o : obs_) hs.push_back(o->handle());
        Util::check(qsv_obs_hamiltonian(c.data(), hs.data(), static_cast<int>(hs.size()), &this->handle_));
    }
    static auto create(std::initializer_list<T> coeffs, std::initializer_list<ObsPtr> obs) {
        return std::make_shared<HamiltonianGPU<T>>(std::vector<T>(coeffs), std::vector<ObsPtr>(obs));
    }
    std::vector<std::size_t> getWires() const override {
        std::unordered_set<std::size_t> s;
        for (const auto &o : obs_)
            for (auto w : o->getWires()) s.insert(w);
        std::vector<std::size_t> w(s.begin(), s.end());
        std::sort(w.begin(), w.end());
        return w;
    }
    std::string getObsName() const override {
        std::ostringstream s;
        s << "Hamiltonian: { 'coeffs' : [";
        for (std::size_t i = 0; i < coeffs_.size(); ++i) s << (i ? ", " : "") << coeffs_[i];
        s << "], 'observables' : [";
        for (std::size_t i = 0; i < obs_.size(); ++i) s << (i ? ", " : "") << obs_[i]->getObsName();
        s << "]}";
        return s.str();
    }
    const std::vector<T> &getCoeffs() const { return coeffs_; }
    const std::vector<ObsPtr> &getObs() const { return obs_; }

  private:
    bool isEqual(const ObservableGPU<T> &other) const override {
        const auto &o = static_cast<const HamiltonianGPU<T> &>(other);
        if (coeffs_ != o.coeffs_ || obs_.size() != o.obs_.size()) return false;
        for (std::size_t i = 0; i < obs_.size(); ++i)
            if (*obs_[i] != *o.obs_[i]) return false;
        return true;
    }
    std::vector<T> coeffs_;
    std::vector<ObsPtr> obs_;
};

template <typename T> class SparseHamiltonianGPU final : public ObservableGPU<T> {
  public:
    // the reference uses int32 indices for complex64 and int64 for complex128 (ObservablesGPU.hpp:381-384)
    using IdxT = typename std::conditional<std::is_same<T, float>::value, int32_t, int64_t>::type;
    template <typename T1, typename T2, typename T3 = T2, typename T4>
    SparseHamiltonianGPU(T1 &&data, T2 &&indices, T3 &&offsets, T4 &&wires)
        : data_{std::forward<T1>(data)}, indices_{std::forward<T2>(indices)}, offsets_{std::forward<T3>(offsets)},
          wires_{std::forward<T4>(wires)} {
        PL_ABORT_IF(data_.size() != indices_.size(), "sparse data and indices differ in length");
        std::vector<int64_t> ip(offsets_.begin(), offsets_.end()), ix(indices_.begin(), indices_.end());
        std::vector<double> v(2 * data_.size());
        for (std::size_t i = 0; i < data_.size(); ++i) {
            v[2 * i] = data_[i].real();
            v[2 * i + 1] = data_[i].imag();
        }
        Util::check(qsv_obs_sparse(ip.data(), static_cast<int64_t>(ip.size()), ix.data(), v.data(),
                                   static_cast<int64_t>(ix.size()), &this->handle_));
    }
    static auto create(std::initializer_list<std::complex<T>> data, std::initializer_list<IdxT> indices,
                       std::initializer_list<IdxT> offsets, std::initializer_list<std::size_t> wires) {
        return std::make_shared<SparseHamiltonianGPU<T>>(std::vector<std::complex<T>>(data), std::vector<IdxT>(indices),
                                                         std::vector<IdxT>(offsets), std::vector<std::size_t>(wires));
    }
    std::string getObsName() const override {
        std::ostringstream s;
        s << "SparseHamiltonian: {\n'data' : ";
        for (const auto &d : data_) s << "{" << d.real() << ", " << d.imag() << "}, ";
        s << "\n'indices' : ";
        for (const auto &i : indices_) s << i << ", ";
        s << "\n'offsets' : ";
        for (const auto &o : offsets_) s << o << ", ";
        s << "\n}";
        return s.str();
    }
    std::vector<std::size_t> getWires() const override { return wires_; }

  private:
    bool isEqual(const ObservableGPU<T> &other) const override {
        const auto &o = static_cast<const SparseHamiltonianGPU<T> &>(other);
        return data_ == o.data_ && indices_ == o.indices_ && offsets_ == o.offsets_;
    }
    std::vector<std::complex<T>> data_;
    std::vector<IdxT> indices_;
    std::vector<IdxT> offsets_;
    std::vector<std::size_t> wires_;
};

}  // namespace Pennylane::Algorithms
