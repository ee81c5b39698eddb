// ref_shim stub (test infrastructure)
#pragma once
#include <cstddef>
namespace tbb {
template <typename T>
class blocked_range {
public:
    blocked_range(T b, T e, size_t grain = 1) : b_(b), e_(e) { (void)grain; }
    T begin() const { return b_; }
    T end() const { return e_; }
private:
    T b_, e_;
};
}  // namespace tbb
