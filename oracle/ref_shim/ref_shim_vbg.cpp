// ref_shim_vbg.cpp — TEST INFRASTRUCTURE.  Compiles the reference's own CPU implementation of the voxel-block-grid
// kernels — t/geometry/kernel/VoxelBlockGridCPU.cpp (DepthTouchCPU) together with the VoxelBlockGridImpl.h templates it
// instantiates (IntegrateCPU, EstimateRangeCPU, RayCastCPU) — unmodified, from where the file lies under /root/reference,
// against the stub core::Tensor / HashMap / TBB headers in stubs/ (OpenMP ParallelFor as upstream, std:: containers
// behind locks instead of tbb::), and exports them through a small C ABI so that the CPU oracle (oracle/tsdf_oracle.c) can be checked against
// the real thing bit for bit.  No reference source is copied into this repository.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#ifdef _OPENMP
#include <omp.h>
#endif

using std::abs;
using std::max;
using std::min;

#include "open3d/t/geometry/kernel/VoxelBlockGridCPU.cpp"   // from -I /root/reference/cpp

namespace o3c = open3d::core;
namespace o3v = open3d::t::geometry::kernel::voxel_grid;
using open3d::t::geometry::TensorMap;

namespace {
using Key = open3d::utility::MiniVec<int, 3>;
using Hash = open3d::utility::MiniVecHash<int, 3>;
using Eq = open3d::utility::MiniVecEq<int, 3>;

o3c::Tensor K_tensor(const double* K) { return o3c::Tensor((void*)K, {3, 3}, o3c::Float64); }
o3c::Tensor E_tensor(const double* E) { return o3c::Tensor((void*)E, {4, 4}, o3c::Float64); }
}  // namespace

// IntegrateCPU<u16,u8,f32,W,C> / <f32,f32,f32,W,C> (VoxelBlockGridImpl.h:151-308) for both value layouts the
// reference instantiates: (W, C) = (uint16_t, uint16_t) — the slam::Model layout — and (float, float).
template <typename W>
static void integrate_layout(const void* depth, const void* color, int inputs_f32, int rows, int cols, const int32_t* buf_indices,
                             int64_t n_blocks, const int32_t* block_keys, int64_t capacity, float* tsdf, W* weight, W* color_buf,
                             const double dK[9], const double cK[9], const double E[16], int resolution, float voxel_size,
                             float sdf_trunc, float depth_scale, float depth_max) {
    const o3c::Dtype vdt = sizeof(W) == 2 ? o3c::UInt16 : o3c::Float32;
    const int64_t r3 = (int64_t)resolution * resolution * resolution;
    o3c::Tensor d((void*)depth, {rows, cols, 1}, inputs_f32 ? o3c::Float32 : o3c::UInt16);
    o3c::Tensor c = color ? o3c::Tensor((void*)color, {rows, cols, 3}, inputs_f32 ? o3c::Float32 : o3c::UInt8) : o3c::Tensor();
    o3c::Tensor idx((void*)buf_indices, {n_blocks}, o3c::Int32);
    o3c::Tensor keys((void*)block_keys, {capacity, 3}, o3c::Int32);
    TensorMap vm("tsdf");
    vm["tsdf"] = o3c::Tensor(tsdf, {capacity * r3, 1}, o3c::Float32);
    vm["weight"] = o3c::Tensor(weight, {capacity * r3, 1}, vdt);
    if (color_buf) vm["color"] = o3c::Tensor(color_buf, {capacity * r3, 3}, vdt);
    if (inputs_f32)
        o3v::IntegrateCPU<float, float, float, W, W>(d, c, idx, keys, vm, K_tensor(dK), K_tensor(cK), E_tensor(E), resolution,
                                                     voxel_size, sdf_trunc, depth_scale, depth_max);
    else
        o3v::IntegrateCPU<uint16_t, uint8_t, float, W, W>(d, c, idx, keys, vm, K_tensor(dK), K_tensor(cK), E_tensor(E), resolution,
                                                          voxel_size, sdf_trunc, depth_scale, depth_max);
}

extern "C" {

int ref_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// DepthTouchCPU (VoxelBlockGridCPU.cpp:117-201): returns the number of touched blocks, keys in arbitrary order.
int64_t ref_depth_touch(const void* depth, int is_f32, int rows, int cols, const double K[9], const double E[16],
                        int resolution, float voxel_size, float sdf_trunc, float depth_scale, float depth_max,
                        int stride, int32_t* keys_out, int64_t max_keys) {
    o3c::Tensor d((void*)depth, {rows, cols, 1}, is_f32 ? o3c::Float32 : o3c::UInt16);
    std::shared_ptr<o3c::HashMap> hm;
    o3c::Tensor coords;
    o3v::DepthTouchCPU(hm, d, K_tensor(K), E_tensor(E), coords, resolution, voxel_size, sdf_trunc, depth_scale,
                       depth_max, stride);
    const int64_t n = coords.GetLength();
    memcpy(keys_out, coords.GetDataPtr<int32_t>(), (size_t)std::min(n, max_keys) * 3 * sizeof(int32_t));
    return n;
}

void ref_integrate(const void* depth, const void* color, int inputs_f32, int rows, int cols, const int32_t* buf_indices,
                   int64_t n_blocks, const int32_t* block_keys, int64_t capacity, float* tsdf, uint16_t* weight,
                   uint16_t* color_buf, const double dK[9], const double cK[9], const double E[16], int resolution,
                   float voxel_size, float sdf_trunc, float depth_scale, float depth_max) {
    integrate_layout<uint16_t>(depth, color, inputs_f32, rows, cols, buf_indices, n_blocks, block_keys, capacity, tsdf, weight,
                               color_buf, dK, cK, E, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
}
void ref_integrate_f32_values(const void* depth, const void* color, int inputs_f32, int rows, int cols,
                              const int32_t* buf_indices, int64_t n_blocks, const int32_t* block_keys, int64_t capacity,
                              float* tsdf, float* weight, float* color_buf, const double dK[9], const double cK[9],
                              const double E[16], int resolution, float voxel_size, float sdf_trunc, float depth_scale,
                              float depth_max) {
    integrate_layout<float>(depth, color, inputs_f32, rows, cols, buf_indices, n_blocks, block_keys, capacity, tsdf, weight,
                            color_buf, dK, cK, E, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
}

// EstimateRangeCPU (VoxelBlockGridImpl.h:310-555).  frag_capacity <= 0: upstream's own heuristic allocation.
// Returns upstream's fragment_buffer length after the call (needed fragments if it had to grow).
int64_t ref_estimate_range(const int32_t* block_keys, int64_t n, const double K[9], const double E[16], int h, int w,
                           int down_factor, int resolution, float voxel_size, float depth_min, float depth_max,
                           int64_t frag_capacity, float* range_out) {
    o3c::Tensor keys((void*)block_keys, {n, 3}, o3c::Int32);
    o3c::Tensor range, frags;
    if (frag_capacity > 0) frags = o3c::Tensor({frag_capacity, 6}, o3c::Float32);
    o3v::EstimateRangeCPU(keys, range, K_tensor(K), E_tensor(E), h, w, down_factor, resolution, voxel_size, depth_min,
                          depth_max, frags);
    memcpy(range_out, range.GetDataPtr<float>(), (size_t)(h / down_factor) * (w / down_factor) * 2 * sizeof(float));
    return frags.GetLength();
}

// RayCastCPU<float, uint16_t, uint16_t> (VoxelBlockGridImpl.h:578-1120).  Output pointers may be null.
void ref_ray_cast(const int32_t* table_keys, int64_t size, const float* tsdf, const uint16_t* weight,
                  const uint16_t* color_buf, const float* range, const double K[9], const double E[16], int h, int w,
                  int resolution, float voxel_size, float depth_scale, float depth_min, float depth_max,
                  float weight_threshold, float trunc_voxel_multiplier, int down, float* depth_out, float* vertex_out,
                  float* color_out, float* normal_out, int64_t* index_out, uint8_t* mask_out, float* ratio_out,
                  float* ratio_dx_out, float* ratio_dy_out, float* ratio_dz_out) {
    const int64_t r3 = (int64_t)resolution * resolution * resolution;
    auto backend = std::make_shared<o3c::TBBHashBackend<Key, Hash, Eq>>();
    for (int64_t s = 0; s < size; ++s)
        (*backend->GetImpl())[Key(table_keys[3 * s], table_keys[3 * s + 1], table_keys[3 * s + 2])] = (o3c::buf_index_t)s;
    auto hm = std::make_shared<o3c::HashMap>(backend);
    TensorMap vm("tsdf");
    vm["tsdf"] = o3c::Tensor((void*)tsdf, {size * r3, 1}, o3c::Float32);
    vm["weight"] = o3c::Tensor((void*)weight, {size * r3, 1}, o3c::UInt16);
    if (color_buf) vm["color"] = o3c::Tensor((void*)color_buf, {size * r3, 3}, o3c::UInt16);
    o3c::Tensor rng((void*)range, {h / down, w / down, 2}, o3c::Float32);
    TensorMap out("range");
    if (depth_out) out["depth"] = o3c::Tensor(depth_out, {h, w, 1}, o3c::Float32);
    if (vertex_out) out["vertex"] = o3c::Tensor(vertex_out, {h, w, 3}, o3c::Float32);
    if (color_out) out["color"] = o3c::Tensor(color_out, {h, w, 3}, o3c::Float32);
    if (normal_out) out["normal"] = o3c::Tensor(normal_out, {h, w, 3}, o3c::Float32);
    if (index_out) out["index"] = o3c::Tensor(index_out, {h, w, 8}, o3c::Int64);
    if (mask_out) out["mask"] = o3c::Tensor(mask_out, {h, w, 8}, o3c::Bool);
    if (ratio_out) out["interp_ratio"] = o3c::Tensor(ratio_out, {h, w, 8}, o3c::Float32);
    if (ratio_dx_out) out["interp_ratio_dx"] = o3c::Tensor(ratio_dx_out, {h, w, 8}, o3c::Float32);
    if (ratio_dy_out) out["interp_ratio_dy"] = o3c::Tensor(ratio_dy_out, {h, w, 8}, o3c::Float32);
    if (ratio_dz_out) out["interp_ratio_dz"] = o3c::Tensor(ratio_dz_out, {h, w, 8}, o3c::Float32);
    o3v::RayCastCPU<float, uint16_t, uint16_t>(hm, vm, rng, out, K_tensor(K), E_tensor(E), h, w, resolution, voxel_size,
                                               depth_scale, depth_min, depth_max, weight_threshold,
                                               trunc_voxel_multiplier, down);
}

}  // extern "C"
