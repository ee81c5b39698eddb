// ref_shim.cpp — TEST INFRASTRUCTURE.  Compiles the reference's OWN header-inline
// math for the hot path, unmodified, from where it lies under /root/reference
// (include path: stubs/ first, then /root/reference/cpp), and exports it through a
// small C ABI so that tests can validate the CPU oracle (oracle/*.c) against the
// real reference code.  Output: oracle/_ref/libo3dref.so (git-ignored, travels to
// the GPU box).  No reference source is copied into this repository.
//
// Reference functions exposed (all header-inline, zero third-party dependencies):
//   GetJacobianPointToPlane<float>, GetJacobianColoredICP<float>   t/pipelines/kernel/RegistrationImpl.h:251-287, 413-493
//   PoseToTransformationImpl<double>                               t/pipelines/kernel/TransformationConverterImpl.h:22-42
//   TransformPointsKernel<float>, TransformNormalsKernel<float>    t/geometry/kernel/TransformImpl.h:20-62
//   SpatialHash, ComputeVoxelIndex                                 core/nns/NeighborSearchCommon.h:31-52
//   utility::MiniVecHash<int,3>                                    core/hashmap/Dispatch.h:67-81
//   DISPATCH_ROBUST_KERNEL_FUNCTION                                t/pipelines/registration/RobustKernelImpl.h:35-115
//   TransformIndexer, ArrayIndexer                                 t/geometry/kernel/GeometryIndexer.h:25-144, 160-420
//   solve_svd3x3<float> / <double>                                 core/linalg/kernel/SVD3x3.h:2170-2215
//   odometry::GetJacobianPointToPlane, HuberDeriv, HuberLoss         t/pipelines/kernel/RGBDOdometryJacobianImpl.h:29-160
//   image::ClipTransformCPU, PyrDownDepthCPU, CreateVertexMapCPU,
//          CreateNormalMapCPU (whole functions, serial ParallelFor)  t/geometry/kernel/ImageImpl.h:86-315
#include <cmath>
#include <cstdint>
#include <cstring>

using std::abs;
using std::exp;
using std::max;
using std::min;
using std::pow;

#include "open3d/core/hashmap/Dispatch.h"
#include "open3d/core/linalg/kernel/SVD3x3.h"
#include "open3d/core/nns/NeighborSearchCommon.h"
#include "open3d/t/geometry/kernel/GeometryIndexer.h"
#define OPEN3D_SKIP_TRANSFORM_MAIN   // TransformImpl.h:90: keep only the per-point kernels
#include "open3d/t/geometry/kernel/TransformImpl.h"
#include "open3d/t/geometry/kernel/ImageImpl.h"
#include "open3d/t/pipelines/kernel/RGBDOdometryJacobianImpl.h"
#include "open3d/t/pipelines/kernel/RegistrationImpl.h"
#include "open3d/t/pipelines/kernel/TransformationConverterImpl.h"
#include "open3d/t/pipelines/registration/RobustKernelImpl.h"

namespace o3k = open3d::t::pipelines::kernel;
namespace o3r = open3d::t::pipelines::registration;
namespace o3g = open3d::t::geometry::kernel;

extern "C" {

int ref_jacobian_p2plane_f32(int64_t i, const float* src, const float* tgt, const float* nrm,
                             const int64_t* corr, float J[6], float* r) {
    return o3k::GetJacobianPointToPlane<float>(i, src, tgt, nrm, corr, J, *r) ? 1 : 0;
}

int ref_jacobian_colored_f32(int64_t i, const float* src, const float* src_c, const float* tgt,
                             const float* nrm, const float* tgt_c, const float* tgt_g, const int64_t* corr,
                             float sqrt_lg, float sqrt_lp, float JG[6], float JI[6], float* rG, float* rI) {
    return o3k::GetJacobianColoredICP<float>(i, src, src_c, tgt, nrm, tgt_c, tgt_g, corr, sqrt_lg, sqrt_lp, JG,
                                             JI, *rG, *rI)
                   ? 1
                   : 0;
}

void ref_pose_to_transformation(const double pose[6], double T[16]) {
    // kernel/TransformationConverter.cpp:81-104: identity, rotation from Impl, translation copied
    for (int i = 0; i < 16; ++i) T[i] = (i % 5 == 0) ? 1.0 : 0.0;
    o3k::PoseToTransformationImpl<double>(T, pose);
    T[3] = pose[3];
    T[7] = pose[4];
    T[11] = pose[5];
}

void ref_transform_points_f32(const float T[16], float* pts, int64_t n) {
    for (int64_t i = 0; i < n; ++i) open3d::t::geometry::kernel::transform::TransformPointsKernel<float>(T, pts + 3 * i);
}

void ref_transform_normals_f32(const float T[16], float* nrm, int64_t n) {
    for (int64_t i = 0; i < n; ++i) open3d::t::geometry::kernel::transform::TransformNormalsKernel<float>(T, nrm + 3 * i);
}

uint64_t ref_spatial_hash(int x, int y, int z) { return (uint64_t)open3d::core::nns::SpatialHash(x, y, z); }

void ref_compute_voxel_index_f32(const float pos[3], float inv_voxel_size, int out[3]) {
    open3d::utility::MiniVec<float, 3> p(pos);
    auto v = open3d::core::nns::ComputeVoxelIndex(p, inv_voxel_size);
    out[0] = v[0];
    out[1] = v[1];
    out[2] = v[2];
}

uint64_t ref_minivec_hash_i32x3(int x, int y, int z) {
    open3d::utility::MiniVec<int, 3> k;
    k[0] = x;
    k[1] = y;
    k[2] = z;
    return open3d::utility::MiniVecHash<int, 3>()(k);
}

double ref_robust_weight_f64(int method, double scaling_p, double shape_p, double residual) {
    using o3r::RobustKernelMethod;
    using namespace open3d;
    double out = 0;
    typedef double scalar_t;
    DISPATCH_ROBUST_KERNEL_FUNCTION(static_cast<RobustKernelMethod>(method), scalar_t, scaling_p, shape_p,
                                    [&]() { out = GetWeightFromRobustKernel(residual); });
    return out;
}

float ref_robust_weight_f32(int method, double scaling_p, double shape_p, float residual) {
    using o3r::RobustKernelMethod;
    using namespace open3d;
    float out = 0;
    typedef float scalar_t;
    DISPATCH_ROBUST_KERNEL_FUNCTION(static_cast<RobustKernelMethod>(method), scalar_t, scaling_p, shape_p,
                                    [&]() { out = GetWeightFromRobustKernel(residual); });
    return out;
}

// TransformIndexer round trip: which = 0 RigidTransform, 1 Project, 2 Unproject
void ref_transform_indexer(const double K[9], const double E[16], float scale, int which, const float in[3],
                           float out[3]) {
    open3d::core::Tensor Kt((void*)K, {3, 3}, open3d::core::Float64);
    open3d::core::Tensor Et((void*)E, {4, 4}, open3d::core::Float64);
    o3g::TransformIndexer ti(Kt, Et, scale);
    out[0] = out[1] = out[2] = 0;
    if (which == 0) ti.RigidTransform(in[0], in[1], in[2], &out[0], &out[1], &out[2]);
    else if (which == 1) ti.Project(in[0], in[1], in[2], &out[0], &out[1]);
    else ti.Unproject(in[0], in[1], in[2], &out[0], &out[1], &out[2]);
}

// ArrayIndexer: voxel workload -> (x,y,z) and InBoundary, as IntegrateCPU uses them
void ref_workload_to_coord3(int res, int workload, int out[3]) {
    o3g::TArrayIndexer<int> idx(open3d::core::SizeVector{res, res, res});
    idx.WorkloadToCoord(workload, &out[0], &out[1], &out[2]);
}
int ref_in_boundary2(int rows, int cols, float x, float y) {
    o3g::TArrayIndexer<int> idx(open3d::core::SizeVector{rows, cols});
    return idx.InBoundary(x, y) ? 1 : 0;
}

// x = pinv(A) b through the reference's own fast 3x3 SVD (used by EstimateColorGradients)
void ref_solve_svd3x3_f32(const float A[9], const float b[3], float x[3]) {
    open3d::core::linalg::kernel::solve_svd3x3<float>(A, b, x);
}
void ref_solve_svd3x3_f64(const double A[9], const double b[3], double x[3]) {
    open3d::core::linalg::kernel::solve_svd3x3<double>(A, b, x);
}

// ---- RGB-D odometry
float ref_huber_deriv(float r, float delta) { return o3k::odometry::HuberDeriv(r, delta); }
float ref_huber_loss(float r, float delta) { return o3k::odometry::HuberLoss(r, delta); }

// One pixel of ComputeOdometryResultPointToPlane: vertex / normal maps are [rows][cols][3] f32.
int ref_odometry_jacobian_p2plane(int x, int y, float depth_outlier_trunc, const float* source_vertex,
                                  const float* target_vertex, const float* target_normal, int rows, int cols,
                                  const double K[9], const double T[16], float J[6], float* r) {
    using open3d::core::Tensor;
    Tensor sv((void*)source_vertex, {rows, cols, 3}, open3d::core::Float32);
    Tensor tv((void*)target_vertex, {rows, cols, 3}, open3d::core::Float32);
    Tensor tn((void*)target_normal, {rows, cols, 3}, open3d::core::Float32);
    Tensor Kt((void*)K, {3, 3}, open3d::core::Float64), Tt((void*)T, {4, 4}, open3d::core::Float64);
    o3g::NDArrayIndexer svi(sv, 2), tvi(tv, 2), tni(tn, 2);
    o3g::TransformIndexer ti(Kt, Tt);
    return o3k::odometry::GetJacobianPointToPlane(x, y, depth_outlier_trunc, svi, tvi, tni, ti, J, *r) ? 1 : 0;
}

// Image kernels, whole functions.  depth_dtype: 0 = u16, 1 = f32 (ClipTransform's source).
void ref_clip_transform(const void* src, int src_is_f32, int rows, int cols, float scale, float min_value,
                        float max_value, float clip_fill, float* dst) {
    open3d::core::Tensor s((void*)src, {rows, cols, 1}, src_is_f32 ? open3d::core::Float32 : open3d::core::UInt16);
    open3d::core::Tensor d((void*)dst, {rows, cols, 1}, open3d::core::Float32);
    o3g::image::ClipTransformCPU(s, d, scale, min_value, max_value, clip_fill);
}
void ref_pyr_down_depth(const float* src, int rows, int cols, float depth_diff, float invalid_fill, float* dst) {
    open3d::core::Tensor s((void*)src, {rows, cols, 1}, open3d::core::Float32);
    open3d::core::Tensor d((void*)dst, {rows / 2, cols / 2, 1}, open3d::core::Float32);
    o3g::image::PyrDownDepthCPU(s, d, depth_diff, invalid_fill);
}
void ref_create_vertex_map(const float* depth, int rows, int cols, const double K[9], float invalid_fill,
                           float* vertex) {
    open3d::core::Tensor s((void*)depth, {rows, cols, 1}, open3d::core::Float32);
    open3d::core::Tensor d((void*)vertex, {rows, cols, 3}, open3d::core::Float32);
    open3d::core::Tensor Kt((void*)K, {3, 3}, open3d::core::Float64);
    o3g::image::CreateVertexMapCPU(s, d, Kt, invalid_fill);
}
void ref_create_normal_map(const float* vertex, int rows, int cols, float invalid_fill, float* normal) {
    open3d::core::Tensor s((void*)vertex, {rows, cols, 3}, open3d::core::Float32);
    open3d::core::Tensor d((void*)normal, {rows, cols, 3}, open3d::core::Float32);
    o3g::image::CreateNormalMapCPU(s, d, invalid_fill);
}

}  // extern "C"
