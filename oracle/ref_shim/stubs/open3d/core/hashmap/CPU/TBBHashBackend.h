// ref_shim stub (test infrastructure): TBBHashBackend::GetImpl() is a tbb::concurrent_unordered_map upstream
// (core/hashmap/CPU/TBBHashBackend.h); RayCastCPU only calls find()/end() on it, which std::unordered_map provides.
#pragma once
#include <memory>
#include <unordered_map>

#include "open3d/core/hashmap/HashMap.h"

namespace open3d {
namespace core {
template <typename Key, typename Hash, typename Eq>
class TBBHashBackend : public DeviceHashBackend {
public:
    using Map = std::unordered_map<Key, buf_index_t, Hash, Eq>;
    TBBHashBackend() : impl_(std::make_shared<Map>()) {}
    std::shared_ptr<Map> GetImpl() const { return impl_; }
private:
    std::shared_ptr<Map> impl_;
};
}  // namespace core
}  // namespace open3d
