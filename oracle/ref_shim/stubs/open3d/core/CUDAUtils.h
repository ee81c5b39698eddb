// ref_shim stub (test infrastructure): host-only build of the reference's own
// header-inline math.  Mirrors the non-CUDA branch of core/CUDAUtils.h:40-52.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#define OPEN3D_FORCE_INLINE inline
#define OPEN3D_HOST_DEVICE
#define OPEN3D_DEVICE
#define OPEN3D_ASSERT_HOST_DEVICE_LAMBDA(type)
#define OPEN3D_CUDA_CHECK(err)
#define OPEN3D_GET_LAST_CUDA_ERROR(message)
#define CUDA_CALL(cuda_function, ...) throw std::runtime_error("no CUDA in ref_shim")
#include "open3d/utility/Logging.h"

#define OPEN3D_ASSERT_MSG(cond, msg) do { if (!(cond)) throw std::runtime_error(msg); } while (0)
#define OPEN3D_ASSERT(cond) do { if (!(cond)) throw std::runtime_error(#cond); } while (0)

namespace open3d {
namespace core {
class Device;
namespace cuda {
inline void Synchronize() {}
inline void Synchronize(const Device&) {}
}  // namespace cuda
}  // namespace core
}  // namespace open3d
