// ref_shim stub (test infrastructure): core::ParallelFor (upstream: tbb::parallel_for over blocked ranges,
// core/ParallelFor.h:76-110; here: an OpenMP static parallel for — TBB is absent) and the element-wise Indexer
// of image::ToCPU.
#pragma once
#include <cstdint>
#include <initializer_list>

#include "open3d/core/Tensor.h"

namespace open3d {
namespace core {
enum class DtypePolicy { NONE };
// element-wise indexer of contiguous same-shape tensors (only what image::ToCPU touches; never run here)
class Indexer {
public:
    Indexer(std::initializer_list<Tensor> in, const Tensor& out, DtypePolicy) : in_(*in.begin()), out_(out) {}
    int64_t NumWorkloads() const { return out_.NumElements(); }
    template <typename T>
    T* GetInputPtr(int, int64_t i) const { return const_cast<T*>(in_.GetDataPtr<T>()) + i; }
    template <typename T>
    T* GetOutputPtr(int64_t i) const { return const_cast<T*>(out_.GetDataPtr<T>()) + i; }
private:
    Tensor in_, out_;
};

template <typename func_t>
void ParallelFor(const Device&, int64_t n, const func_t& func) {
#pragma omp parallel for schedule(static) if (n > 4096)
    for (int64_t i = 0; i < n; ++i) func(i);
}
}  // namespace core
}  // namespace open3d
