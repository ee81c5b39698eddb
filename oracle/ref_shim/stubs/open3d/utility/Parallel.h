#pragma once
namespace open3d {
namespace utility {
inline int EstimateMaxThreads() { return 1; }
inline bool InParallel() { return false; }
}  // namespace utility
}  // namespace open3d
