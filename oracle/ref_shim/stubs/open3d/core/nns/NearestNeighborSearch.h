// ref_shim stub (test infrastructure): the interface PointCloudImpl.h's host wrappers name; never run in the shim
// (the real one needs nanoflann).
#pragma once
#include <tuple>
#include <utility>

#include "open3d/core/Tensor.h"
namespace open3d {
namespace core {
namespace nns {
class NearestNeighborSearch {
public:
    NearestNeighborSearch(const Tensor&, const Dtype& = Int64) {}
    bool KnnIndex() { unsupported(); }
    bool FixedRadiusIndex(double = 0) { unsupported(); }
    bool HybridIndex(double = 0) { unsupported(); }
    std::pair<Tensor, Tensor> KnnSearch(const Tensor&, int) { unsupported(); }
    std::tuple<Tensor, Tensor, Tensor> FixedRadiusSearch(const Tensor&, double, bool = true) { unsupported(); }
    std::tuple<Tensor, Tensor, Tensor> HybridSearch(const Tensor&, double, int) const { unsupported(); }
private:
    [[noreturn]] static void unsupported() { utility::LogError("ref_shim: NearestNeighborSearch is not available"); }
};
}  // namespace nns
}  // namespace core
}  // namespace open3d
