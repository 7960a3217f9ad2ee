// Thread-safe pool of GPU ids, interface of the reference's util/DevicePool.hpp:15-130 (used by the
// observable-batched adjoint, AdjointDiffGPU.hpp:392-473).
#pragma once
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "Error.hpp"

namespace Pennylane::CUDA {

template <class DeviceIndexType = int> class DevicePool {
  public:
    DevicePool() {
        for (DeviceIndexType i = 0; i < static_cast<DeviceIndexType>(getTotalDevices()); ++i) free_.push_back(i);
    }
    std::size_t getActiveDevices() const {
        std::lock_guard<std::mutex> lk(m_);
        return active_.size();
    }
    bool isActive(const DeviceIndexType &idx) const {
        std::lock_guard<std::mutex> lk(m_);
        return active_.count(idx) != 0;
    }
    bool isInactive(const DeviceIndexType &idx) const { return !isActive(idx); }
    int acquireDevice() {
        std::lock_guard<std::mutex> lk(m_);
        PL_ABORT_IF(free_.empty(), "No free GPU device available");
        const DeviceIndexType d = free_.back();
        free_.pop_back();
        active_.insert(d);
        return d;
    }
    void releaseDevice(DeviceIndexType idx) {
        std::lock_guard<std::mutex> lk(m_);
        if (active_.erase(idx)) free_.push_back(idx);
    }
    void syncDevice(DeviceIndexType) {}
    static int getTotalDevices() {
        int n = 0;
        Util::check(qsv_device_count(&n));
        return n;
    }
    static std::vector<std::string> getDeviceUIDs() {
        std::vector<std::string> out;
        for (int i = 0; i < getTotalDevices(); ++i) out.push_back("GPU-" + std::to_string(i));
        return out;
    }
    static void setDeviceIdx(DeviceIndexType) {}

  private:
    mutable std::mutex m_;
    std::vector<DeviceIndexType> free_;
    std::unordered_set<DeviceIndexType> active_;
};

}  // namespace Pennylane::CUDA
