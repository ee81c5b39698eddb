// ref_shim stub (test infrastructure): serial execution, std::unordered_set suffices.
#pragma once
#include <unordered_set>
namespace tbb {
template <typename K, typename H = std::hash<K>, typename E = std::equal_to<K>>
using concurrent_unordered_set = std::unordered_set<K, H, E>;
}  // namespace tbb
