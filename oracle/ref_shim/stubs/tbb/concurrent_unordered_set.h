// ref_shim stub (test infrastructure): std::unordered_set behind a lock — thread-safe emplace is all
// DepthTouchCPU needs from tbb::concurrent_unordered_set (iteration happens after the parallel loop).
#pragma once
#include <mutex>
#include <unordered_set>
namespace tbb {
template <typename K, typename H = std::hash<K>, typename E = std::equal_to<K>>
class concurrent_unordered_set {
public:
    template <typename... Args>
    void emplace(Args&&... args) {
        K k(std::forward<Args>(args)...);
        std::lock_guard<std::mutex> lock(mu_);
        set_.insert(k);
    }
    size_t size() const { return set_.size(); }
    auto begin() const { return set_.begin(); }
    auto end() const { return set_.end(); }
private:
    std::unordered_set<K, H, E> set_;
    std::mutex mu_;
};
}  // namespace tbb
