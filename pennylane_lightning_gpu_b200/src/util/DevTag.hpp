// (device id, stream) tag with the interface of the reference's util/DevTag.hpp:14-65; the stream is an
// opaque pointer here because nothing above the C ABI includes CUDA headers.
#pragma once
#include <utility>

namespace Pennylane::CUDA {

template <class DevId = int> class DevTag {
  public:
    DevTag() = default;
    explicit DevTag(DevId device_id) : device_id_(device_id) {}
    DevTag(DevId device_id, void *stream_id) : device_id_(device_id), stream_id_(stream_id) {}
    DevTag(const DevTag &) = default;
    DevTag &operator=(const DevTag &) = default;

    DevId getDeviceID() const { return device_id_; }
    void *getStreamID() const { return stream_id_; }
    // the reference calls cudaSetDevice here; every C-ABI entry already sets the handle's device
    void refresh() const {}
    bool operator==(const DevTag &o) const { return device_id_ == o.device_id_ && stream_id_ == o.stream_id_; }

  private:
    DevId device_id_{0};
    void *stream_id_{nullptr};
};

}  // namespace Pennylane::CUDA
