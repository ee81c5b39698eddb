// ref_shim stub (test infrastructure): one range, one body call — the serial order of the reduction.
#pragma once
#include "tbb/blocked_range.h"
namespace tbb {
template <typename Range, typename Value, typename Body, typename Join>
Value parallel_reduce(const Range& range, const Value& identity, const Body& body, const Join&) {
    return body(range, identity);
}
}  // namespace tbb
