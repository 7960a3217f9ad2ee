// Whole-state copies between HOST memory and the device (CopyHostDataToGpu / CopyGpuDataToHost of
// simulator/StateVectorCudaBase.hpp:104-228, reached from Python through syncH2D / syncD2H,
// pennylane_lightning_gpu/lightning_gpu.py:345-381).
//
// The reference hands the caller's pointer to one cudaMemcpy.  A NumPy array is pageable memory, for which the driver
// stages the copy through a small internal pinned buffer on one thread (10-25 GB/s on these hosts); a 16 GiB 30-qubit
// state then spends most of a second in the copy.  Here a copy of pageable memory is cut into chunks that several host
// threads move through their own pairs of pinned staging buffers, each on its own CUDA stream: the memcpy of chunk i+1
// into one buffer overlaps the DMA of chunk i out of the other, and the threads together keep the PCIe link busy.
// Pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) or managed memory goes straight to cudaMemcpyAsync.
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "qsv_internal.h"

namespace qsv {

namespace {

constexpr size_t IO_CHUNK = (size_t)16 << 20;   // bytes per staged chunk
constexpr size_t IO_DIRECT_BELOW = (size_t)64 << 20;
constexpr int IO_MAX_WORKERS = 16;

struct IoLane {  // one host thread's staging: two pinned buffers, one stream, one event per buffer
    void *pinned[2] = {nullptr, nullptr};
    cudaStream_t stream = nullptr;
    cudaEvent_t done[2] = {nullptr, nullptr};
};
struct IoPool {
    std::mutex mu;  // one staged copy per device at a time
    std::vector<IoLane> lanes;
};
IoPool g_io[64];

int io_workers() {
    static const int n = [] {
        const char *v = std::getenv("QSV_IO_THREADS");
        int t = v ? std::atoi(v) : (int)std::max(1u, std::thread::hardware_concurrency() / 2);
        return std::max(1, std::min(t, IO_MAX_WORKERS));
    }();
    return n;
}

void ensure_lanes(IoPool &pool, int n) {
    while ((int)pool.lanes.size() < n) {
        IoLane l;
        for (int k = 0; k < 2; ++k) {
            QSV_CUDA(cudaHostAlloc(&l.pinned[k], IO_CHUNK, cudaHostAllocDefault));
            QSV_CUDA(cudaEventCreateWithFlags(&l.done[k], cudaEventDisableTiming));
        }
        QSV_CUDA(cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking));
        pool.lanes.push_back(l);
    }
}

bool is_pageable(const void *host) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

}  // namespace

// dev <-> host, `bytes` long; synchronous (returns when the data has arrived), ordered after the work queued on sv.stream
void copy_state_host(State &sv, void *dev, void *host, size_t bytes, bool to_device) {
    sv.use();
    static const bool staged_ok = [] {
        const char *v = std::getenv("QSV_IO_STAGED");
        return !(v && std::atoi(v) == 0);
    }();
    if (bytes < IO_DIRECT_BELOW || !staged_ok || !is_pageable(host)) {
        QSV_CUDA(cudaMemcpyAsync(to_device ? dev : host, to_device ? host : dev, bytes,
                                 to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, sv.stream));
        QSV_CUDA(cudaStreamSynchronize(sv.stream));
        return;
    }
    QSV_CUDA(cudaStreamSynchronize(sv.stream));  // everything that produced / still reads the device buffer is done
    IoPool &pool = g_io[sv.device & 63];
    std::lock_guard<std::mutex> lk(pool.mu);
    const size_t n_chunks = (bytes + IO_CHUNK - 1) / IO_CHUNK;
    const int n_workers = (int)std::min<size_t>((size_t)io_workers(), n_chunks);
    ensure_lanes(pool, n_workers);
    std::atomic<size_t> next{0};
    std::atomic<int> failed{0};
    auto work = [&](int w) {
        if (cudaSetDevice(sv.device) != cudaSuccess) {
            failed = 1;
            return;
        }
        IoLane &l = pool.lanes[w];
        bool used[2] = {false, false};
        size_t pending_chunk[2] = {0, 0};  // device -> host: the chunk whose DMA into pinned[k] is in flight
        auto drain = [&](int k) {            // device -> host: staging buffer k -> caller's memory
            if (!used[k]) return;
            if (cudaEventSynchronize(l.done[k]) != cudaSuccess) failed = 1;
            const size_t off = pending_chunk[k] * IO_CHUNK, len = std::min(IO_CHUNK, bytes - off);
            std::memcpy((char *)host + off, l.pinned[k], len);
            used[k] = false;
        };
        int k = 0;
        for (;;) {
            const size_t c = next.fetch_add(1);  // chunks are handed out dynamically: the threads stay balanced
            if (c >= n_chunks || failed) break;
            const size_t off = c * IO_CHUNK, len = std::min(IO_CHUNK, bytes - off);
            if (to_device) {
                if (used[k] && cudaEventSynchronize(l.done[k]) != cudaSuccess) failed = 1;  // buffer k is free again
                std::memcpy(l.pinned[k], (const char *)host + off, len);
                if (cudaMemcpyAsync((char *)dev + off, l.pinned[k], len, cudaMemcpyHostToDevice, l.stream) != cudaSuccess ||
                    cudaEventRecord(l.done[k], l.stream) != cudaSuccess)
                    failed = 1;
                used[k] = true;
            } else {
                drain(k);
                if (cudaMemcpyAsync(l.pinned[k], (const char *)dev + off, len, cudaMemcpyDeviceToHost, l.stream) != cudaSuccess ||
                    cudaEventRecord(l.done[k], l.stream) != cudaSuccess)
                    failed = 1;
                used[k] = true;
                pending_chunk[k] = c;
            }
            k ^= 1;
        }
        if (!to_device) {
            drain(k);
            drain(k ^ 1);
        }
        if (cudaStreamSynchronize(l.stream) != cudaSuccess) failed = 1;
    };
    std::vector<std::thread> threads;
    for (int w = 1; w < n_workers; ++w) threads.emplace_back(work, w);
    work(0);
    for (auto &t : threads) t.join();
    if (failed) {
        const cudaError_t e = cudaGetLastError();
        fail(std::string("staged host copy failed: ") + cudaGetErrorString(e));
    }
}

}  // namespace qsv
