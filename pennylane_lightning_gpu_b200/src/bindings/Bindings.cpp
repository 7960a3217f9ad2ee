// pybind11 module `lightning_gpu_qubit_ops` with the Python-visible names of the reference's
// bindings/Bindings.cpp (classes LightningGPU_C64/C128, *ObsGPU_C*, OpsStructGPU_C*,
// AdjointJacobianGPU_C*, DevPool, DevTag; functions device_reset, allToAllAccess, is_gpu_supported,
// get_gpu_arch; exception PLException), so that lightning_gpu.py / _serialize.py import unchanged
// (lightning_gpu.py:52-90 of the reference).  Every method forwards to the C++ shells in
// ../simulator and ../algorithms, which only call the C ABI of libqsv_b200.so.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <complex>
#include <sstream>
#include <string>
#include <vector>

#include "AdjointDiffGPU.hpp"
#include "AdjointDiffGPUMPI.hpp"
#include "DevTag.hpp"
#include "DevicePool.hpp"
#include "Error.hpp"
#include "ObservablesGPU.hpp"
#include "MPIManager.hpp"
#include "ObservablesGPUMPI.hpp"
#include "StateVectorCudaMPI.hpp"
#include "StateVectorCudaManaged.hpp"

namespace py = pybind11;
using namespace Pennylane;
using namespace Pennylane::Algorithms;
using Pennylane::CUDA::DevicePool;
using Pennylane::CUDA::DevTag;
using Pennylane::MPI::MPIManager;
using Pennylane::Util::LightningException;

namespace {

template <class T> std::vector<std::complex<T>> to_vec(const py::array_t<std::complex<T>, py::array::c_style | py::array::forcecast> &a) {
    const auto info = a.request();
    const auto *p = static_cast<const std::complex<T> *>(info.ptr);
    return info.size ? std::vector<std::complex<T>>(p, p + info.size) : std::vector<std::complex<T>>{};
}

// one registration per named gate: method(wires, adjoint, params)
template <class SV, class PyClass> void register_gates(PyClass &cls) {
    static const char *gates[] = {"Identity", "PauliX", "PauliY", "PauliZ", "Hadamard", "S", "T", "CNOT", "SWAP", "CY", "CZ",
                                  "Toffoli", "CSWAP", "PhaseShift", "ControlledPhaseShift", "RX", "RY", "RZ", "Rot", "CRX",
                                  "CRY", "CRZ", "CRot", "IsingXX", "IsingYY", "IsingZZ", "SingleExcitation",
                                  "SingleExcitationMinus", "SingleExcitationPlus", "DoubleExcitation",
                                  "DoubleExcitationMinus", "DoubleExcitationPlus", "MultiRZ"};
    using P = typename SV::Precision;
    for (const char *g : gates) {
        const std::string name(g);
        cls.def(
            g,
            [name](SV &sv, const std::vector<std::size_t> &wires, bool adjoint, const std::vector<P> &params) {
                sv.applyOperation(name, wires, adjoint, params);
            },
            ("Apply the " + name + " gate.").c_str());
    }
}

template <class PrecisionT> void register_precision(py::module_ &m) {
    using SV = StateVectorCudaManaged<PrecisionT>;
    using ParamT = PrecisionT;
    using np_arr_r = py::array_t<ParamT, py::array::c_style | py::array::forcecast>;
    using np_arr_c = py::array_t<std::complex<ParamT>, py::array::c_style | py::array::forcecast>;
    using index_type = typename std::conditional<std::is_same<ParamT, float>::value, int32_t, int64_t>::type;
    using np_arr_idx = py::array_t<index_type, py::array::c_style | py::array::forcecast>;
    const std::string bits = std::to_string(sizeof(std::complex<PrecisionT>) * 8);

    auto sv_cls = py::class_<SV>(m, ("LightningGPU_C" + bits).c_str());
    sv_cls.def(py::init<std::size_t>())
        .def(py::init<std::size_t, DevTag<int>>())
        .def(py::init<const SV &>())
        .def(py::init([](const np_arr_c &arr) {
            const auto info = arr.request();
            return new SV(static_cast<const std::complex<PrecisionT> *>(info.ptr), static_cast<std::size_t>(arr.size()));
        }))
        .def(
            "setBasisState",
            [](SV &sv, std::size_t index, bool use_async) { sv.setBasisState({1, 0}, index, use_async); },
            "Create Basis State on GPU.")
        .def(
            "setStateVector",
            [](SV &sv, const np_arr_idx &indices, const np_arr_c &state, bool use_async) {
                sv.template setStateVector<index_type>(static_cast<index_type>(indices.request().size),
                                                       static_cast<const std::complex<PrecisionT> *>(state.request().ptr),
                                                       static_cast<const index_type *>(indices.request().ptr), use_async);
            },
            "Set State Vector on GPU with values and their corresponding indices for the state vector on device")
        .def("apply",
             py::overload_cast<const std::vector<std::string> &, const std::vector<std::vector<std::size_t>> &,
                               const std::vector<bool> &, const std::vector<std::vector<PrecisionT>> &>(&SV::applyOperation))
        .def("apply",
             py::overload_cast<const std::vector<std::string> &, const std::vector<std::vector<std::size_t>> &,
                               const std::vector<bool> &, const std::vector<std::vector<PrecisionT>> &,
                               const std::vector<std::vector<std::complex<PrecisionT>>> &>(&SV::applyOperation),
             "Whole list of operations as one recorded circuit (fused sweeps); matrices for operations without a kernel")
        .def("apply", py::overload_cast<const std::vector<std::string> &, const std::vector<std::vector<std::size_t>> &,
                                        const std::vector<bool> &>(&SV::applyOperation))
        .def("apply", py::overload_cast<const std::string &, const std::vector<std::size_t> &, bool,
                                        const std::vector<PrecisionT> &, const std::vector<std::complex<PrecisionT>> &>(
                          &SV::applyOperation_std))
        .def(
            "ExpectationValue",
            [](SV &sv, const std::string &obsName, const std::vector<std::size_t> &wires, const std::vector<ParamT> &params,
               const np_arr_c &gate_matrix) { return sv.expval(obsName, wires, params, to_vec<ParamT>(gate_matrix)).real(); },
            "Calculate the expectation value of the given observable.")
        .def(
            "ExpectationValue",
            [](SV &sv, const std::vector<std::string> &obsName, const std::vector<std::size_t> &wires,
               const std::vector<std::vector<ParamT>> &, const np_arr_c &gate_matrix) {
                std::string concat{"#"};
                for (const auto &s : obsName) concat += s;
                return sv.expval(concat, wires, std::vector<ParamT>{}, to_vec<ParamT>(gate_matrix)).real();
            },
            "Calculate the expectation value of the given observable.")
        .def(
            "ExpectationValue",
            [](SV &sv, const std::vector<std::size_t> &wires, const np_arr_c &gate_matrix) {
                return sv.expval(wires, to_vec<ParamT>(gate_matrix)).real();
            },
            "Calculate the expectation value of a dense Hamiltonian matrix on the given wires.")
        .def(
            "ExpectationValue",
            [](SV &sv, const np_arr_idx &csrOffsets, const np_arr_idx &columns, const np_arr_c values) {
                return sv.template getExpectationValueOnSparseSpMV<index_type>(
                    static_cast<const index_type *>(csrOffsets.request().ptr), static_cast<index_type>(csrOffsets.request().size),
                    static_cast<const index_type *>(columns.request().ptr),
                    static_cast<const std::complex<PrecisionT> *>(values.request().ptr),
                    static_cast<index_type>(values.request().size));
            },
            "Calculate the expectation value of a sparse Hamiltonian.")
        .def(
            "ExpectationValue",
            [](SV &sv, const std::vector<std::string> &pauli_words, const std::vector<std::vector<std::size_t>> &target_wires,
               const np_arr_c &coeffs) {
                return sv.getExpectationValuePauliWords(pauli_words, target_wires,
                                                        static_cast<const std::complex<PrecisionT> *>(coeffs.request().ptr));
            },
            "Calculate the expectation value of a Hamiltonian composed solely from sums of Pauli-words")
        .def(
            "Probability",
            [](SV &sv, const std::vector<std::size_t> &wires) { return py::array_t<ParamT>(py::cast(sv.probability(wires))); },
            "Calculate the probabilities for given wires. Results returned in Col-major order.")
        .def("GenerateSamples",
             [](SV &sv, std::size_t num_wires, std::size_t num_shots) {
                 auto result = sv.generate_samples(num_shots);
                 py::array_t<std::size_t> out({num_shots, num_wires});
                 std::copy(result.begin(), result.end(), out.mutable_data());
                 return out;
             })
        .def("GenerateSamples",
             [](SV &sv, std::size_t num_wires, std::size_t num_shots, std::uint64_t seed) {
                 auto result = sv.generate_samples(num_shots, seed);
                 py::array_t<std::size_t> out({num_shots, num_wires});
                 std::copy(result.begin(), result.end(), out.mutable_data());
                 return out;
             })
        .def(
            "DeviceToDevice", [](SV &sv, const SV &other, bool async) { sv.updateData(other, async); },
            "Synchronize data from another GPU device to current device.")
        .def(
            "DeviceToHost",
            [](const SV &gpu_sv, np_arr_c &cpu_sv, bool) {
                auto info = cpu_sv.request();
                if (cpu_sv.size()) gpu_sv.CopyGpuDataToHost(static_cast<std::complex<PrecisionT> *>(info.ptr), cpu_sv.size());
            },
            "Synchronize data from the GPU device to host.")
        .def(
            "HostToDevice",
            [](SV &gpu_sv, const np_arr_c &cpu_sv, bool async) {
                const auto info = cpu_sv.request();
                const auto length = static_cast<std::size_t>(info.shape[0]);
                if (length) gpu_sv.CopyHostDataToGpu(static_cast<const std::complex<PrecisionT> *>(info.ptr), length, async);
            },
            "Synchronize data from the host device to GPU.")
        .def("GetNumGPUs", [](SV &) { return DevicePool<int>::getTotalDevices(); }, "Get the number of available GPUs.")
        .def("getCurrentGPU", [](SV &sv) { return sv.getDevTag().getDeviceID(); }, "Get the GPU index for the statevector data.")
        .def("numQubits", &SV::getNumQubits)
        .def("dataLength", &SV::getLength)
        .def("resetGPU", &SV::initSV);
    register_gates<SV>(sv_cls);

    // ---- observables ----
    using Obs = ObservableGPU<PrecisionT>;
    using ObsPtr = std::shared_ptr<Obs>;
    py::class_<Obs, ObsPtr>(m, ("ObservableGPU_C" + bits).c_str(), py::module_local());

#define QSV_OBS_COMMON(CLS)                                                                        \
    .def("__repr__", &CLS::getObsName)                                                             \
        .def("get_wires", &CLS::getWires, "Get wires of observables")                              \
        .def(                                                                                      \
            "__eq__",                                                                              \
            [](const CLS &self, py::handle other) -> bool {                                        \
                if (!py::isinstance<CLS>(other)) return false;                                     \
                return self == *other.cast<std::shared_ptr<CLS>>();                                \
            },                                                                                     \
            "Compare two observables")

    using Named = NamedObsGPU<PrecisionT>;
    py::class_<Named, std::shared_ptr<Named>, Obs>(m, ("NamedObsGPU_C" + bits).c_str(), py::module_local())
        .def(py::init([](const std::string &name, const std::vector<std::size_t> &wires) {
            return std::make_shared<Named>(name, wires);
        })) QSV_OBS_COMMON(Named);

    using Herm = HermitianObsGPU<PrecisionT>;
    py::class_<Herm, std::shared_ptr<Herm>, Obs>(m, ("HermitianObsGPU_C" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_c &matrix, const std::vector<std::size_t> &wires) {
            return std::make_shared<Herm>(to_vec<ParamT>(matrix), wires);
        })) QSV_OBS_COMMON(Herm);

    using Tensor = TensorProdObsGPU<PrecisionT>;
    py::class_<Tensor, std::shared_ptr<Tensor>, Obs>(m, ("TensorProdObsGPU_C" + bits).c_str(), py::module_local())
        .def(py::init([](const std::vector<ObsPtr> &obs) { return std::make_shared<Tensor>(obs); })) QSV_OBS_COMMON(Tensor);

    using Ham = HamiltonianGPU<PrecisionT>;
    py::class_<Ham, std::shared_ptr<Ham>, Obs>(m, ("HamiltonianGPU_C" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_r &coeffs, const std::vector<ObsPtr> &obs) {
            const auto info = coeffs.request();
            const auto *p = static_cast<const ParamT *>(info.ptr);
            return std::make_shared<Ham>(std::vector<ParamT>(p, p + info.size), obs);
        })) QSV_OBS_COMMON(Ham);

    using Sparse = SparseHamiltonianGPU<PrecisionT>;
    using SpIDX = typename Sparse::IdxT;
    py::class_<Sparse, std::shared_ptr<Sparse>, Obs>(m, ("SparseHamiltonianGPU_C" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_c &data, const np_arr_idx &indices, const np_arr_idx &offsets,
                         const std::vector<std::size_t> &wires) {
            const auto *ip = static_cast<const SpIDX *>(indices.request().ptr);
            const auto *op = static_cast<const SpIDX *>(offsets.request().ptr);
            return std::make_shared<Sparse>(to_vec<ParamT>(data), std::vector<SpIDX>(ip, ip + indices.size()),
                                            std::vector<SpIDX>(op, op + offsets.size()), wires);
        })) QSV_OBS_COMMON(Sparse);
#undef QSV_OBS_COMMON

    // ---- operations record + adjoint Jacobian ----
    using Ops = OpsData<SV>;
    py::class_<Ops>(m, ("OpsStructGPU_C" + bits).c_str(), py::module_local())
        .def(py::init<const std::vector<std::string> &, const std::vector<std::vector<ParamT>> &,
                      const std::vector<std::vector<std::size_t>> &, const std::vector<bool> &,
                      const std::vector<std::vector<std::complex<PrecisionT>>> &>())
        .def("__repr__", [](const Ops &ops) {
            std::ostringstream s;
            for (std::size_t op = 0; op < ops.getSize(); ++op) {
                s << "{'name': " << ops.getOpsName()[op] << ", 'params': [";
                for (std::size_t j = 0; j < ops.getOpsParams()[op].size(); ++j) s << (j ? ", " : "") << ops.getOpsParams()[op][j];
                s << "], 'inv': " << ops.getOpsInverses()[op] << "}" << (op + 1 < ops.getSize() ? "," : "");
            }
            return "Operations: [" + s.str() + "]";
        });

    using Adj = AdjointJacobianGPU<PrecisionT>;
    auto jacobian = [](Adj &adj, const SV &sv, const std::vector<ObsPtr> &observables, const Ops &operations,
                       const std::vector<std::size_t> &trainableParams) {
        std::vector<std::vector<PrecisionT>> jac;
        adj.adjointJacobian(sv, jac, observables, operations, trainableParams, false);
        py::array_t<ParamT> out({observables.size(), trainableParams.size()});
        for (std::size_t i = 0; i < jac.size(); ++i) std::copy(jac[i].begin(), jac[i].end(), out.mutable_data(i, 0));
        return out;
    };
    py::class_<Adj>(m, ("AdjointJacobianGPU_C" + bits).c_str(), py::module_local())
        .def(py::init<>())
        .def("create_ops_list",
             [](Adj &, const std::vector<std::string> &ops_name, const std::vector<np_arr_r> &ops_params,
                const std::vector<std::vector<std::size_t>> &ops_wires, const std::vector<bool> &ops_inverses,
                const std::vector<np_arr_c> &ops_matrices) {
                 std::vector<std::vector<PrecisionT>> params(ops_name.size());
                 std::vector<std::vector<std::complex<PrecisionT>>> mats(ops_name.size());
                 for (std::size_t op = 0; op < ops_name.size(); ++op) {
                     if (op < ops_params.size()) {
                         const auto info = ops_params[op].request();
                         const auto *p = static_cast<const ParamT *>(info.ptr);
                         if (info.size) params[op].assign(p, p + info.size);
                     }
                     if (op < ops_matrices.size()) mats[op] = to_vec<ParamT>(ops_matrices[op]);
                 }
                 return Ops{ops_name, params, ops_wires, ops_inverses, mats};
             })
        .def("adjoint_jacobian", jacobian)
        .def("adjoint_jacobian_batched",
             [](Adj &adj, const SV &sv, const std::vector<ObsPtr> &observables, const Ops &operations,
                const std::vector<std::size_t> &trainableParams) {
                 std::vector<std::vector<PrecisionT>> jac;
                 adj.batchAdjointJacobian(sv, jac, observables, operations, trainableParams, false);
                 py::array_t<ParamT> out({observables.size(), trainableParams.size()});
                 for (std::size_t i = 0; i < jac.size(); ++i) std::copy(jac[i].begin(), jac[i].end(), out.mutable_data(i, 0));
                 return out;
             });
}


// ---- sharded twins: LightningGPUMPI_C*, *ObsGPUMPI_C*, OpsStructGPUMPI_C*, AdjointJacobianGPUMPI_C*
// (bindings/Bindings.cpp:930-1720 of the reference) ---------------------------------------------------------
template <class PrecisionT> void register_precision_mpi(py::module_ &m) {
    using SV = StateVectorCudaMPI<PrecisionT>;
    using ParamT = PrecisionT;
    using np_arr_r = py::array_t<ParamT, py::array::c_style | py::array::forcecast>;
    using np_arr_c = py::array_t<std::complex<ParamT>, py::array::c_style | py::array::forcecast>;
    using index_type = typename std::conditional<std::is_same<ParamT, float>::value, int32_t, int64_t>::type;
    using np_arr_idx = py::array_t<index_type, py::array::c_style | py::array::forcecast>;
    const std::string bits = std::to_string(sizeof(std::complex<PrecisionT>) * 8);

    auto sv_cls = py::class_<SV>(m, ("LightningGPUMPI_C" + bits).c_str());
    sv_cls
        .def(py::init([](MPIManager &mpi_manager, const DevTag<int> devtag_local, std::size_t mpi_buf_size,
                         std::size_t num_global_qubits, std::size_t num_local_qubits) {
                 return new SV(mpi_manager, devtag_local, mpi_buf_size, num_global_qubits, num_local_qubits);
             }),
             py::keep_alive<1, 2>())
        .def(py::init<const SV &>())
        .def(
            "setBasisState", [](SV &sv, std::size_t index, bool use_async) { sv.setBasisState({1, 0}, index, use_async); },
            "Create Basis State on GPU.")
        .def(
            "setStateVector",
            [](SV &sv, const np_arr_idx &indices, const np_arr_c &state, bool use_async) {
                sv.template setStateVector<index_type>(static_cast<index_type>(indices.request().size),
                                                       static_cast<const std::complex<PrecisionT> *>(state.request().ptr),
                                                       static_cast<const index_type *>(indices.request().ptr), use_async);
            },
            "Set State Vector on GPU with values and their corresponding indices for the state vector on device")
        .def("apply",
             py::overload_cast<const std::vector<std::string> &, const std::vector<std::vector<std::size_t>> &,
                               const std::vector<bool> &, const std::vector<std::vector<PrecisionT>> &>(&SV::applyOperation))
        .def("apply",
             py::overload_cast<const std::vector<std::string> &, const std::vector<std::vector<std::size_t>> &,
                               const std::vector<bool> &, const std::vector<std::vector<PrecisionT>> &,
                               const std::vector<std::vector<std::complex<PrecisionT>>> &>(&SV::applyOperation),
             "Whole list of operations as one recorded circuit (fused sweeps); matrices for operations without a kernel")
        .def("apply", py::overload_cast<const std::vector<std::string> &, const std::vector<std::vector<std::size_t>> &,
                                        const std::vector<bool> &>(&SV::applyOperation))
        .def("apply", py::overload_cast<const std::string &, const std::vector<std::size_t> &, bool,
                                        const std::vector<PrecisionT> &, const std::vector<std::complex<PrecisionT>> &>(
                          &SV::applyOperation_std))
        .def(
            "ExpectationValue",
            [](SV &sv, const std::string &obsName, const std::vector<std::size_t> &wires, const std::vector<ParamT> &params,
               const np_arr_c &gate_matrix) { return sv.expval(obsName, wires, params, to_vec<ParamT>(gate_matrix)).real(); },
            "Calculate the expectation value of the given observable.")
        .def(
            "ExpectationValue",
            [](SV &sv, const std::vector<std::string> &obsName, const std::vector<std::size_t> &wires,
               const std::vector<std::vector<ParamT>> &, const np_arr_c &gate_matrix) {
                std::string concat{"#"};
                for (const auto &s : obsName) concat += s;
                return sv.expval(concat, wires, std::vector<ParamT>{}, to_vec<ParamT>(gate_matrix)).real();
            },
            "Calculate the expectation value of the given observable.")
        .def(
            "ExpectationValue",
            [](SV &sv, const std::vector<std::size_t> &wires, const np_arr_c &gate_matrix) {
                return sv.expval(wires, to_vec<ParamT>(gate_matrix)).real();
            },
            "Calculate the expectation value of a dense Hamiltonian matrix on the given wires.")
        .def(
            "ExpectationValue",
            [](SV &sv, const np_arr_idx &csrOffsets, const np_arr_idx &columns, const np_arr_c values) {
                return sv.template getExpectationValueOnSparseSpMV<index_type>(
                    static_cast<const index_type *>(csrOffsets.request().ptr), static_cast<index_type>(csrOffsets.request().size),
                    static_cast<const index_type *>(columns.request().ptr),
                    static_cast<const std::complex<PrecisionT> *>(values.request().ptr),
                    static_cast<index_type>(values.request().size));
            },
            "Calculate the expectation value of a sparse Hamiltonian (whole matrix on every rank, or on rank 0 only).")
        .def(
            "ExpectationValue",
            [](SV &sv, const std::vector<std::string> &pauli_words, const std::vector<std::vector<std::size_t>> &target_wires,
               const np_arr_c &coeffs) {
                return sv.getExpectationValuePauliWords(pauli_words, target_wires,
                                                        static_cast<const std::complex<PrecisionT> *>(coeffs.request().ptr));
            },
            "Calculate the expectation value of a Hamiltonian composed solely from sums of Pauli-words")
        .def(
            "Probability",
            [](SV &sv, const std::vector<std::size_t> &wires) { return py::array_t<ParamT>(py::cast(sv.probability(wires))); },
            "Calculate the probabilities for given wires. Results returned in Col-major order.")
        .def("GenerateSamples",
             [](SV &sv, std::size_t num_wires, std::size_t num_shots) {
                 auto result = sv.generate_samples(num_shots);
                 py::array_t<std::size_t> out({num_shots, num_wires});
                 std::copy(result.begin(), result.end(), out.mutable_data());
                 return out;
             })
        .def("GenerateSamples",
             [](SV &sv, std::size_t num_wires, std::size_t num_shots, std::uint64_t seed) {
                 auto result = sv.generate_samples(num_shots, seed);
                 py::array_t<std::size_t> out({num_shots, num_wires});
                 std::copy(result.begin(), result.end(), out.mutable_data());
                 return out;
             })
        .def(
            "DeviceToDevice", [](SV &sv, const SV &other, bool async) { sv.updateData(other, async); },
            "Synchronize data from another GPU device to current device.")
        .def(
            "DeviceToHost",
            [](const SV &gpu_sv, np_arr_c &cpu_sv, bool) {
                auto info = cpu_sv.request();
                if (cpu_sv.size()) gpu_sv.CopyGpuDataToHost(static_cast<std::complex<PrecisionT> *>(info.ptr), cpu_sv.size());
            },
            "Synchronize the local shard from the GPU device to host.")
        .def(
            "HostToDevice",
            [](SV &gpu_sv, const np_arr_c &cpu_sv, bool async) {
                const auto info = cpu_sv.request();
                const auto length = static_cast<std::size_t>(cpu_sv.size());
                if (length) gpu_sv.CopyHostDataToGpu(static_cast<const std::complex<PrecisionT> *>(info.ptr), length, async);
            },
            "Synchronize the local shard from the host to the GPU.")
        .def("GetNumGPUs", [](SV &) { return DevicePool<int>::getTotalDevices(); }, "Get the number of available GPUs.")
        .def("getCurrentGPU", [](SV &sv) { return sv.getDevTag().getDeviceID(); }, "Get the GPU index for the statevector data.")
        .def("numLocalQubits", &SV::getNumLocalQubits)
        .def("numGlobalQubits", &SV::getNumGlobalQubits)
        .def("dataLength", &SV::getLength)
        .def("usesPeerAccess", &SV::usesPeerAccess)
        .def("resetGPU", &SV::initSV_MPI);
    register_gates<SV>(sv_cls);

    using Obs = ObservableGPUMPI<PrecisionT>;
    using ObsPtr = std::shared_ptr<Obs>;
    py::class_<Obs, ObsPtr>(m, ("ObservableGPUMPI_C" + bits).c_str(), py::module_local());
#define QSV_OBS_COMMON(CLS)                                                                        \
    .def("__repr__", &CLS::getObsName)                                                             \
        .def("get_wires", &CLS::getWires, "Get wires of observables")                              \
        .def(                                                                                      \
            "__eq__",                                                                              \
            [](const CLS &self, py::handle other) -> bool {                                        \
                if (!py::isinstance<CLS>(other)) return false;                                     \
                return self == *other.cast<std::shared_ptr<CLS>>();                                \
            },                                                                                     \
            "Compare two observables")
    using Named = NamedObsGPUMPI<PrecisionT>;
    py::class_<Named, std::shared_ptr<Named>, Obs>(m, ("NamedObsGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init([](const std::string &name, const std::vector<std::size_t> &wires) {
            return std::make_shared<Named>(name, wires);
        })) QSV_OBS_COMMON(Named);
    using Herm = HermitianObsGPUMPI<PrecisionT>;
    py::class_<Herm, std::shared_ptr<Herm>, Obs>(m, ("HermitianObsGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_c &matrix, const std::vector<std::size_t> &wires) {
            return std::make_shared<Herm>(to_vec<ParamT>(matrix), wires);
        })) QSV_OBS_COMMON(Herm);
    using Tensor = TensorProdObsGPUMPI<PrecisionT>;
    py::class_<Tensor, std::shared_ptr<Tensor>, Obs>(m, ("TensorProdObsGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init([](const std::vector<ObsPtr> &obs) { return std::make_shared<Tensor>(obs); })) QSV_OBS_COMMON(Tensor);
    using Ham = HamiltonianGPUMPI<PrecisionT>;
    py::class_<Ham, std::shared_ptr<Ham>, Obs>(m, ("HamiltonianGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_r &coeffs, const std::vector<ObsPtr> &obs) {
            const auto info = coeffs.request();
            const auto *p = static_cast<const ParamT *>(info.ptr);
            return std::make_shared<Ham>(std::vector<ParamT>(p, p + info.size), obs);
        })) QSV_OBS_COMMON(Ham);
    using Sparse = SparseHamiltonianGPUMPI<PrecisionT>;
    using SpIDX = typename Sparse::IdxT;
    py::class_<Sparse, std::shared_ptr<Sparse>, Obs>(m, ("SparseHamiltonianGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init([](const np_arr_c &data, const np_arr_idx &indices, const np_arr_idx &offsets,
                         const std::vector<std::size_t> &wires) {
            const auto *ip = static_cast<const SpIDX *>(indices.request().ptr);
            const auto *op = static_cast<const SpIDX *>(offsets.request().ptr);
            return std::make_shared<Sparse>(to_vec<ParamT>(data), std::vector<SpIDX>(ip, ip + indices.size()),
                                            std::vector<SpIDX>(op, op + offsets.size()), wires);
        })) QSV_OBS_COMMON(Sparse);
#undef QSV_OBS_COMMON

    using Ops = OpsData<SV>;
    py::class_<Ops>(m, ("OpsStructGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init<const std::vector<std::string> &, const std::vector<std::vector<ParamT>> &,
                      const std::vector<std::vector<std::size_t>> &, const std::vector<bool> &,
                      const std::vector<std::vector<std::complex<PrecisionT>>> &>())
        .def("__repr__", [](const Ops &ops) { return "Operations: " + std::to_string(ops.getSize()); });

    using Adj = AdjointJacobianGPUMPI<PrecisionT>;
    auto jacobian = [](Adj &adj, const SV &sv, const std::vector<ObsPtr> &observables, const Ops &operations,
                       const std::vector<std::size_t> &trainableParams) {
        std::vector<std::vector<PrecisionT>> jac;
        adj.adjointJacobian(sv, jac, observables, operations, trainableParams, false);
        py::array_t<ParamT> out({observables.size(), trainableParams.size()});
        for (std::size_t i = 0; i < jac.size(); ++i) std::copy(jac[i].begin(), jac[i].end(), out.mutable_data(i, 0));
        return out;
    };
    auto jacobian_serial = [](Adj &adj, const SV &sv, const std::vector<ObsPtr> &observables, const Ops &operations,
                              const std::vector<std::size_t> &trainableParams) {
        std::vector<std::vector<PrecisionT>> jac;
        adj.adjointJacobian_serial(sv, jac, observables, operations, trainableParams, false);
        py::array_t<ParamT> out({observables.size(), trainableParams.size()});
        for (std::size_t i = 0; i < jac.size(); ++i) std::copy(jac[i].begin(), jac[i].end(), out.mutable_data(i, 0));
        return out;
    };
    py::class_<Adj>(m, ("AdjointJacobianGPUMPI_C" + bits).c_str(), py::module_local())
        .def(py::init<>())
        .def("create_ops_list",
             [](Adj &, const std::vector<std::string> &ops_name, const std::vector<np_arr_r> &ops_params,
                const std::vector<std::vector<std::size_t>> &ops_wires, const std::vector<bool> &ops_inverses,
                const std::vector<np_arr_c> &ops_matrices) {
                 std::vector<std::vector<PrecisionT>> params(ops_name.size());
                 std::vector<std::vector<std::complex<PrecisionT>>> mats(ops_name.size());
                 for (std::size_t op = 0; op < ops_name.size(); ++op) {
                     if (op < ops_params.size()) {
                         const auto info = ops_params[op].request();
                         const auto *p = static_cast<const ParamT *>(info.ptr);
                         if (info.size) params[op].assign(p, p + info.size);
                     }
                     if (op < ops_matrices.size()) mats[op] = to_vec<ParamT>(ops_matrices[op]);
                 }
                 return Ops{ops_name, params, ops_wires, ops_inverses, mats};
             })
        .def("adjoint_jacobian", jacobian)
        .def("adjoint_jacobian_serial", jacobian_serial);
}

void register_mpi_manager(py::module_ &m) {
    using np_arr_c64 = py::array_t<std::complex<float>, py::array::c_style | py::array::forcecast>;
    using np_arr_c128 = py::array_t<std::complex<double>, py::array::c_style | py::array::forcecast>;
    py::class_<MPIManager>(m, "MPIManager")
        .def(py::init<>())
        .def("Barrier", &MPIManager::Barrier)
        .def("getRank", &MPIManager::getRank)
        .def("getSize", &MPIManager::getSize)
        .def("getSizeNode", &MPIManager::getSizeNode)
        .def("getTime", &MPIManager::getTime)
        .def("getVendor", &MPIManager::getVendor)
        .def("getVersion", &MPIManager::getVersion)
        .def(
            "Scatter",
            [](MPIManager &mgr, np_arr_c64 &sendBuf, np_arr_c64 &recvBuf, int root) {
                mgr.Scatter<std::complex<float>>(static_cast<std::complex<float> *>(sendBuf.request().ptr),
                                                 static_cast<std::complex<float> *>(recvBuf.request().ptr),
                                                 static_cast<std::size_t>(recvBuf.request().size), root);
            },
            "Scatter of a host array held by `root` (NCCL send/recv through device staging).")
        .def(
            "Scatter",
            [](MPIManager &mgr, np_arr_c128 &sendBuf, np_arr_c128 &recvBuf, int root) {
                mgr.Scatter<std::complex<double>>(static_cast<std::complex<double> *>(sendBuf.request().ptr),
                                                  static_cast<std::complex<double> *>(recvBuf.request().ptr),
                                                  static_cast<std::size_t>(recvBuf.request().size), root);
            },
            "Scatter of a host array held by `root` (NCCL send/recv through device staging).");
}

}  // namespace

PYBIND11_MODULE(lightning_gpu_qubit_ops, m) {
    py::options options;
    options.disable_function_signatures();
    py::register_exception<LightningException>(m, "PLException");

    m.def("device_reset", []() { Util::check(qsv_device_reset()); }, "Reset all GPU devices and contexts.");
    m.def("allToAllAccess", []() { Util::check(qsv_enable_peer_access()); });
    m.def(
        "is_gpu_supported",
        [](int device_number) {
            int major = 0, minor = 0;
            Util::check(qsv_device_arch(device_number, &major, &minor));
            return major == 10;  // the kernels are built for sm_100a only
        },
        py::arg("device_number") = 0, "Checks if the given GPU device is a Blackwell (sm_100) part.");
    m.def(
        "get_gpu_arch",
        [](int device_number) {
            int major = 0, minor = 0;
            Util::check(qsv_device_arch(device_number, &major, &minor));
            return std::make_pair(major, minor);
        },
        py::arg("device_number") = 0, "Returns the given GPU major and minor GPU support.");

    py::class_<DevicePool<int>>(m, "DevPool")
        .def(py::init<>())
        .def("getActiveDevices", &DevicePool<int>::getActiveDevices)
        .def("isActive", &DevicePool<int>::isActive)
        .def("isInactive", &DevicePool<int>::isInactive)
        .def("acquireDevice", &DevicePool<int>::acquireDevice)
        .def("releaseDevice", &DevicePool<int>::releaseDevice)
        .def("syncDevice", &DevicePool<int>::syncDevice)
        .def_static("getTotalDevices", &DevicePool<int>::getTotalDevices)
        .def_static("getDeviceUIDs", &DevicePool<int>::getDeviceUIDs)
        .def_static("setDeviceID", &DevicePool<int>::setDeviceIdx);

    py::class_<DevTag<int>>(m, "DevTag")
        .def(py::init<>())
        .def(py::init<int>())
        .def(py::init([](int device_id, void *stream_id) { return new DevTag<int>(device_id, stream_id); }))
        .def(py::init<const DevTag<int> &>())
        .def("getDeviceID", &DevTag<int>::getDeviceID)
        .def("getStreamID", [](DevTag<int> &t) { return t.getStreamID(); })
        .def("refresh", &DevTag<int>::refresh);

    register_precision<float>(m);
    register_precision<double>(m);
    register_mpi_manager(m);
    register_precision_mpi<float>(m);
    register_precision_mpi<double>(m);
}
