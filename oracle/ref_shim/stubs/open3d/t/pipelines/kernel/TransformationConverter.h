// ref_shim stub (test infrastructure): declarations only.  DecodeAndSolve6x6 is DEFINED by ref_shim_reg.cpp as a
// probe that captures the 29 reduced scalars the reference's CPU kernels hand to it (the real one needs LAPACK).
#pragma once
#include "open3d/core/Tensor.h"
namespace open3d {
namespace t {
namespace pipelines {
namespace kernel {
void DecodeAndSolve6x6(const core::Tensor& A_reduction, core::Tensor& delta, float& inlier_residual, int& inlier_count);
core::Tensor PoseToTransformation(const core::Tensor& pose);
}  // namespace kernel
}  // namespace pipelines
}  // namespace t
}  // namespace open3d
