// State container and error plumbing.
// Owns the 2^n amplitude array the way util/DataBuffer.hpp:31-121 does in the reference (RAII
// cudaMalloc buffer + device/stream tag), plus two lazily grown scratch areas so that no kernel
// launch on the hot path ever calls cudaMalloc/cudaFree (reference defect Q5: per-gate workspace
// allocation at simulator/StateVectorCudaManaged.hpp:1447-1471).
#include <cstdlib>
#include <mutex>

#include "qsv_internal.h"

namespace qsv {

namespace {
thread_local std::string g_last_error;

// ---- workspace cache -------------------------------------------------------------------------------------------
// State-sized temporaries (lambda and the bras of the adjoint Jacobian, O|psi> of an observable, the y of a sparse
// product) are needed again and again by a VQE loop; cudaMalloc / cudaFree of hundreds of MiB cost milliseconds each and
// cudaFree synchronises the device (measured: the 24-qubit adjoint Jacobian took 38 ms instead of 21 ms when other
// allocations were alive).  Released blocks are kept per device, up to QSV_WORKSPACE_CACHE_MB (default 16384), and handed
// out again for requests of exactly the same size.  The reference allocates and frees per call
// (algorithms/AdjointDiffGPU.hpp:513-530, simulator/StateVectorCudaManaged.hpp:829-918).
struct WsBlock {
    void *p;
    size_t cap;
    cudaStream_t stream;  // work queued on this stream may still use the block
};
struct WsCache {
    std::mutex mu;
    std::vector<WsBlock> free_blocks[64];
    size_t cached_bytes[64] = {0};
} g_ws;

size_t ws_limit() {
    static const size_t lim = [] {
        const char *v = std::getenv("QSV_WORKSPACE_CACHE_MB");
        return (size_t)(v ? std::max(0l, std::atol(v)) : 16384l) << 20;
    }();
    return lim;
}
size_t ws_round(size_t bytes) { return std::max<size_t>(bytes, (size_t)2 << 20); }  // >= 2 MiB: own allocation, IPC-exportable
}  // namespace

void ws_trim(int device) {
    std::vector<WsBlock> victims;
    {
        std::lock_guard<std::mutex> lk(g_ws.mu);
        victims.swap(g_ws.free_blocks[device & 63]);
        g_ws.cached_bytes[device & 63] = 0;
    }
    for (const WsBlock &b : victims) cudaFree(b.p);
}

void *ws_acquire(int device, size_t bytes, cudaStream_t stream) {
    const size_t cap = ws_round(bytes);
    {
        std::unique_lock<std::mutex> lk(g_ws.mu);
        auto &v = g_ws.free_blocks[device & 63];
        for (size_t i = 0; i < v.size(); ++i)
            if (v[i].cap == cap) {
                const WsBlock b = v[i];
                v.erase(v.begin() + i);
                g_ws.cached_bytes[device & 63] -= cap;
                lk.unlock();
                if (b.stream != stream) QSV_CUDA(cudaStreamSynchronize(b.stream));
                return b.p;
            }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, cap);
    if (e == cudaErrorMemoryAllocation) {
        cudaGetLastError();
        ws_trim(device);
        e = cudaMalloc(&p, cap);
    }
    QSV_CUDA(e);
    return p;
}

void ws_release(int device, void *p, size_t bytes, cudaStream_t stream) {
    if (!p) return;
    const size_t cap = ws_round(bytes);
    {
        std::lock_guard<std::mutex> lk(g_ws.mu);
        if (g_ws.cached_bytes[device & 63] + cap <= ws_limit()) {
            g_ws.free_blocks[device & 63].push_back({p, cap, stream});
            g_ws.cached_bytes[device & 63] += cap;
            return;
        }
    }
    cudaFree(p);
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
