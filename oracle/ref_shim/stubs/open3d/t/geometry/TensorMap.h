// ref_shim stub (test infrastructure): TensorMap as the kernels use it (Contains / at / operator[]).
#pragma once
#include <string>
#include <unordered_map>

#include "open3d/core/Tensor.h"

namespace open3d {
namespace t {
namespace geometry {
class TensorMap : public std::unordered_map<std::string, core::Tensor> {
public:
    explicit TensorMap(const std::string& primary_key = "") : primary_key_(primary_key) {}
    bool Contains(const std::string& key) const { return count(key) != 0; }
private:
    std::string primary_key_;
};
}  // namespace geometry
}  // namespace t
}  // namespace open3d
