// ref_shim stub (test infrastructure): core/Dispatch.h:30-65 restricted to the dtypes the
// image kernels are exercised with here.
#pragma once
#include <cstdint>

#include "open3d/utility/Logging.h"

#include "open3d/core/Tensor.h"
#define DISPATCH_DTYPE_TO_TEMPLATE(DTYPE, ...)        \
    [&] {                                             \
        if (DTYPE == open3d::core::Float32) {         \
            using scalar_t = float;                   \
            return __VA_ARGS__();                     \
        } else if (DTYPE == open3d::core::UInt16) {   \
            using scalar_t = uint16_t;                \
            return __VA_ARGS__();                     \
        } else {                                      \
            open3d::utility::LogError("Unsupported data type."); \
        }                                             \
    }()

#define DISPATCH_FLOAT_DTYPE_TO_TEMPLATE(DTYPE, ...)  \
    [&] {                                             \
        if (DTYPE == open3d::core::Float32) {         \
            using scalar_t = float;                   \
            return __VA_ARGS__();                     \
        } else if (DTYPE == open3d::core::Float64) {  \
            using scalar_t = double;                  \
            return __VA_ARGS__();                     \
        } else {                                      \
            open3d::utility::LogError("Unsupported data type."); \
        }                                             \
    }()
