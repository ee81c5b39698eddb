// ref_shim stub (test infrastructure): the shim runs core::ParallelFor serially, so the lock is a no-op.
#pragma once
namespace tbb {
class spin_mutex {
public:
    class scoped_lock {
    public:
        explicit scoped_lock(spin_mutex&) {}
    };
};
namespace profiling {
template <typename T>
inline void set_name(T&, const char*) {}
}  // namespace profiling
}  // namespace tbb
