// ref_shim stub (test infrastructure): the two members of core::HashMap the voxel-grid CPU kernels touch.
#pragma once
#include <memory>

#include "open3d/core/Tensor.h"

namespace open3d {
namespace core {
using buf_index_t = uint32_t;   // core/hashmap/HashBackendBuffer.h

class DeviceHashBackend {
public:
    virtual ~DeviceHashBackend() {}
};

class HashMap {
public:
    explicit HashMap(std::shared_ptr<DeviceHashBackend> backend) : backend_(backend) {}
    std::shared_ptr<DeviceHashBackend> GetDeviceHashBackend() const { return backend_; }
    Device GetDevice() const { return Device(); }
private:
    std::shared_ptr<DeviceHashBackend> backend_;
};
}  // namespace core
}  // namespace open3d
