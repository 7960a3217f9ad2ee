// State container and error plumbing.
// Owns the 2^n amplitude array the way util/DataBuffer.hpp:31-121 does in the reference (RAII
// cudaMalloc buffer + device/stream tag), plus two lazily grown scratch areas so that no kernel
// launch on the hot path ever calls cudaMalloc/cudaFree (reference defect Q5: per-gate workspace
// allocation at simulator/StateVectorCudaManaged.hpp:1447-1471).
#include "qsv_internal.h"

namespace qsv {

namespace {
thread_local std::string g_last_error;
}

void set_last_error(const std::string &msg) { g_last_error = msg; }
const std::string &last_error_ref() { return g_last_error; }

[[noreturn]] void fail(const std::string &msg) { throw Error(msg); }

double *State::reduction_buffer(size_t n_doubles) {
    use();
    if (n_doubles > red_cap) {
        // make sure nothing in flight still uses the old buffer
        QSV_CUDA(cudaStreamSynchronize(stream));
        if (red_dev) QSV_CUDA(cudaFree(red_dev));
        if (red_host) QSV_CUDA(cudaFreeHost(red_host));
        red_dev = nullptr;
        red_host = nullptr;
        size_t cap = 256;
        while (cap < n_doubles) cap *= 2;
        QSV_CUDA(cudaMalloc(&red_dev, cap * sizeof(double)));
        QSV_CUDA(cudaMallocHost(&red_host, cap * sizeof(double)));
        red_cap = cap;
    }
    return red_dev;
}

void *State::scratch_buffer(size_t bytes) {
    use();
    if (bytes > scratch_cap) {
        QSV_CUDA(cudaStreamSynchronize(stream));
        if (scratch) QSV_CUDA(cudaFree(scratch));
        scratch = nullptr;
        size_t cap = 1 << 16;
        while (cap < bytes) cap *= 2;
        QSV_CUDA(cudaMalloc(&scratch, cap));
        scratch_cap = cap;
    }
    return scratch;
}

State::~State() {
    cudaSetDevice(device);
    if (dist) dist_free(*this);
    if (stream) cudaStreamSynchronize(stream);
    if (owns && data) cudaFree(data);
    if (red_dev) cudaFree(red_dev);
    if (red_host) cudaFreeHost(red_host);
    if (scratch) cudaFree(scratch);
}

}  // namespace qsv
