/*
 * oracle.h — CPU restatement of the Open3D ICP + TSDF hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load this library, and only as the checker or as
 * the timed CPU baseline.  The product path (open3d_b200/csrc) never links,
 * imports or calls it.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the Open3D tree, cpp/open3d/...).  Parity status is stated per function:
 * "pinned: <reference KAT>" or "parity unpinned".
 *
 * Plain C99 + OpenMP (the reference CPU path is TBB; TBB is absent here).
 */
#ifndef O3D_ORACLE_H_
#define O3D_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ hashes */

/* core/hashmap/Dispatch.h:67-81 (MiniVecHash<int,3>), int32 sign-extended to
 * uint64, FNV-1a style.  Pinned indirectly (set semantics) by
 * tests/t/geometry/VoxelBlockGrid.cpp:211-219; the hash VALUE is pinned against
 * the reference header itself through oracle/_ref (ref_shim). */
uint64_t orc_minivec_hash_i32x3(int32_t x, int32_t y, int32_t z);

/* core/nns/NeighborSearchCommon.h:31-37 SpatialHash: int arithmetic (wraps),
 * widened to size_t. */
uint64_t orc_spatial_hash(int32_t x, int32_t y, int32_t z);

/* core/nns/NeighborSearchCommon.h:44-52 ComputeVoxelIndex (f32). */
void orc_compute_voxel_index_f32(const float pos[3], float inv_voxel_size,
                                 int32_t out[3]);

/* ------------------------------------------------------------ robust kernel */

/* t/pipelines/registration/RobustKernelImpl.h:35-115, enum order of
 * RobustKernel.h:15-23: 0 L2, 1 L1, 2 Huber, 3 Cauchy, 4 GM, 5 Tukey,
 * 6 Generalized.  Pinned: tests/t/pipelines/registration/Registration.cpp:411-490.
 */
double orc_robust_weight_f64(int method, double scale, double shape,
                             double residual);
float orc_robust_weight_f32(int method, double scale, double shape,
                            float residual);

/* ------------------------------------------------------------------- search */

/* Hybrid search semantics (radius + max_knn), restating
 * core/nns/FixedRadiusSearchImpl.cuh:514-631 (CUDA) and
 * core/nns/NanoFlannImpl.h:306-370 (CPU, nanoflann v1.5.0 radiusSearch sorted,
 * truncated to max_knn): for every query the max_knn nearest points with
 * dist^2 <= radius^2, ascending; idx padded with -1, dist^2 with 0.
 * dist^2 is computed in f32 as fma(dz,dz,fma(dy,dy,dx*dx)); exact ties go to
 * the lower point index (the reference leaves tie order undefined).
 * Uniform-grid accelerated, exact.  Pinned: tests/core/NearestNeighborSearch.cpp:321-383.
 */
void orc_hybrid_search_f32(const float* points, int64_t num_points,
                           const float* queries, int64_t num_queries,
                           double radius, int max_knn, int32_t* idx /*N*k*/,
                           float* dist2 /*N*k*/, int32_t* counts /*N*/);

/* O(N*M) brute force twin of the above; used to validate the grid version. */
void orc_hybrid_search_bruteforce_f32(const float* points, int64_t num_points,
                                      const float* queries,
                                      int64_t num_queries, double radius,
                                      int max_knn, int32_t* idx, float* dist2,
                                      int32_t* counts);

/* -------------------------------------------------------- pose estimation */

/* t/pipelines/kernel/RegistrationImpl.h:251-287 (Jacobian, f32 per-term math)
 * + t/pipelines/kernel/RegistrationCPU.cpp:30-90 (29-slot packing).
 * sums64: accumulated in f64 (the precise target); sums32: accumulated in f32
 * in index order (what a single-threaded reference build would produce);
 * abs64: sum of |term| per slot (the natural error scale).  Any may be NULL.
 * corr == -1 means "no correspondence".
 * Pinned: tests/t/pipelines/registration/TransformationEstimation.cpp:33-84,148,176.
 */
void orc_pose_p2plane_sums_f32(const float* src, const float* tgt,
                               const float* tgt_normals, const int64_t* corr,
                               int64_t n, int robust_method,
                               double robust_scale, double robust_shape,
                               double sums64[29], float sums32[29],
                               double abs64[29]);

/* Same for f64 clouds (all math in f64). */
void orc_pose_p2plane_sums_f64(const double* src, const double* tgt,
                               const double* tgt_normals, const int64_t* corr,
                               int64_t n, int robust_method,
                               double robust_scale, double robust_shape,
                               double sums64[29]);

/* t/pipelines/kernel/RegistrationImpl.h:413-493 + RegistrationCPU.cpp
 * (ComputePoseColoredICPKernelCPU): two residual rows per correspondence;
 * slot 27 = sum(r_G^2 + r_I^2).  parity unpinned offline (the reference's
 * ColoredICP tests need a downloaded dataset); validated against the reference
 * header itself via oracle/_ref. */
void orc_pose_colored_sums_f32(const float* src, const float* src_colors,
                               const float* tgt, const float* tgt_normals,
                               const float* tgt_colors,
                               const float* tgt_color_gradients,
                               const int64_t* corr, int64_t n,
                               double lambda_geometric, int robust_method,
                               double robust_scale, double robust_shape,
                               double sums64[29], double abs64[29]);

/* t/pipelines/kernel/TransformationConverter.cpp:189-226 DecodeAndSolve6x6:
 * unpack lower-tri, solve AtA x = -Atb by LU with partial pivoting (LAPACK
 * dgesv semantics).  Returns 0, or 1 if a pivot is exactly zero (singular). */
int orc_decode_and_solve_6x6(const double sums[29], double pose[6],
                             double* residual, int* inlier_count);

/* t/pipelines/kernel/TransformationConverterImpl.h:22-42 + .cpp:81-104:
 * pose (alpha,beta,gamma,tx,ty,tz) -> row-major 4x4 (f64). */
void orc_pose_to_transformation(const double pose[6], double T[16]);

/* t/geometry/kernel/TransformImpl.h:20-45: in-place p <- (T p)/w, f32, T cast
 * to f32 first (kernel/Transform.cpp).  No FMA contraction. */
void orc_transform_points_f32(const double T[16], float* points, int64_t n);
/* TransformImpl.h:47-62 */
void orc_transform_normals_f32(const double T[16], float* normals, int64_t n);

/* TransformationEstimation.cpp:161-194 ComputeRMSE (point to plane). */
double orc_rmse_p2plane_f32(const float* src, const float* tgt,
                            const float* tgt_normals, const int64_t* corr,
                            int64_t n);

/* ---------------------------------------------------------------- ICP loop */

typedef struct {
    int num_iterations;    /* iterations executed (RegistrationResult::num_iterations_) */
    int converged;         /* RegistrationResult::converged_ */
    double fitness;        /* final fitness_ */
    double inlier_rmse;    /* final inlier_rmse_ */
    double transformation[16]; /* row-major 4x4 f64 */
    double loop_seconds;   /* wall time of the iteration loop only (no index build / final evaluation) */
    double build_seconds;  /* wall time of the search-index build */
} orc_icp_result;

/* t/pipelines/registration/Registration.cpp:24-62, 275-360, 362-444 for one
 * scale, PointToPlane estimation, no down-sampling (voxel_size = -1).
 * source/target are not modified (the reference clones the source).
 * per_iter (may be NULL) receives 2*max_iteration doubles (fitness, rmse of
 * every executed iteration, as seen by the callback at Registration.cpp:330-345);
 * corr_out (may be NULL) receives the final N correspondences (int64).
 * accumulate_f64 != 0: 29 sums accumulated in f64 (precise); 0: f32. */
int orc_icp_p2plane_f32(const float* source, int64_t n, const float* target,
                        const float* target_normals, int64_t m,
                        double max_corr_dist, const double init_T[16],
                        int max_iteration, double rel_fitness, double rel_rmse,
                        int robust_method, double robust_scale,
                        double robust_shape, int accumulate_f64,
                        orc_icp_result* result, double* per_iter,
                        int64_t* corr_out);

/* registration::GetInformationMatrix's reduction (Registration.cpp:446-485 -> kernel::ComputeInformationMatrix):
 * 6x6 f64 GTG over the target points matched by `corr` ([n] int64, -1 = none). */
void orc_information_matrix_f32(const float* tgt, const int64_t* corr, int64_t n, double info36[36]);

/* ------------------------------------------------------------ ColoredICP */

/* t/geometry/kernel/PointCloudImpl.h:1066-1165 EstimatePointWiseColorGradientKernel driven by
 * EstimateColorGradientsUsingHybridSearch (:1167-1222): hybrid search of the cloud on itself
 * (radius, max_nn), skip neighbour 0, project neighbours on the tangent plane, 3x3 normal
 * equations + the orthogonality row (all f32, upstream's order), then the 3x3 solve:
 *   ORC_GRADIENT_SOLVER_REFERENCE (default): upstream's solve_svd3x3<float> (core/linalg/kernel/SVD3x3.h — the
 *     4-sweep fast f32 SVD) restated in svd3_oracle.c, bit-identical to the reference's own function compiled in
 *     oracle/_ref (tests/test_oracle_vs_ref.py) — so the whole per-point kernel equals the reference's bit for bit;
 *   ORC_GRADIENT_SOLVER_EXACT: the exact pseudo-inverse of the same f32 system (f64 Jacobi), the option the
 *     product also offers (the fast SVD is off by a median 12 % on these condition-1e5 systems).
 * SURVEY.md 8f #3. */
#define ORC_GRADIENT_SOLVER_REFERENCE 0
#define ORC_GRADIENT_SOLVER_EXACT 1
void orc_estimate_color_gradients_f32(const float* points, const float* normals, const float* colors,
                                      int64_t n, double radius, int max_nn, float* gradients_out);
void orc_estimate_color_gradients_solver_f32(const float* points, const float* normals, const float* colors,
                                             int64_t n, double radius, int max_nn, int solver,
                                             float* gradients_out);

/* pinv(A) b for a symmetric 3x3 A (row-major), f64 Jacobi; singular values < 1e-10 dropped. */
void orc_solve_sym3x3_pinv(const double A[9], const double b[3], double x[3]);

/* svd3_oracle.c: core/linalg/kernel/SVD3x3.h:1131-2168 svd3x3<float> (A = U diag(S) V^T, row-major) and
 * :2170-2215 solve_svd3x3<float>, the host build's operation sequence. */
void orc_svd3x3_f32(const float A[9], float U[9], float S[3], float V[9]);
void orc_solve_svd3x3_f32(const float A[9], const float b[3], float x[3]);

/* Registration.cpp:275-444 with TransformationEstimationForColoredICP
 * (TransformationEstimation.cpp:382-432 -> ComputePoseColoredICP, kernel/Registration.cpp:137-191). */
int orc_icp_colored_f32(const float* source, const float* source_colors, int64_t n, const float* target,
                        const float* target_normals, const float* target_colors,
                        const float* target_color_gradients, int64_t m, double max_corr_dist,
                        const double init_T[16], int max_iteration, double rel_fitness, double rel_rmse,
                        double lambda_geometric, int robust_method, double robust_scale,
                        double robust_shape, orc_icp_result* result, double* per_iter, int64_t* corr_out);

/* ------------------------------------------------------- VoxelDownSample */

/* t/geometry/PointCloud.cpp:496-560 PointCloud::VoxelDownSample(voxel_size, "mean"):
 * voxel = floor(p / voxel_size) evaluated in f32 (the division is a Float32 tensor op),
 * one output point per occupied voxel = mean of every attribute (positions, and normals /
 * colors when given; normals are NOT re-normalised upstream).  The reference accumulates
 * in f32 with IndexAdd_ (order undefined); this restatement accumulates in f64 and rounds
 * once.  Output order upstream is the hash map's slot order (undefined); here voxels are
 * sorted lexicographically by (vx, vy, vz) and voxel_keys_out (may be NULL) receives them.
 * Returns the number of voxels.  SURVEY.md 8f #1; parity unpinned offline (the reference's
 * tests compare against a downloaded cloud), set-of-voxels exact by construction. */
int64_t orc_voxel_down_sample_f32(const float* positions, const float* normals, const float* colors,
                                  int64_t n, double voxel_size, float* positions_out,
                                  float* normals_out, float* colors_out, int32_t* voxel_keys_out);

/* ------------------------------------------------------------------- TSDF */

/* t/geometry/kernel/GeometryIndexer.h:136-143 stores intrinsics/extrinsics as
 * float; t/geometry/Utility.h:77-115 InverseTransformation in f64. */
void orc_inverse_transformation(const double T[16], double Tinv[16]);

/* t/geometry/kernel/VoxelBlockGridCPU.cpp:117-201 DepthTouchCPU.
 * depth: H*W, u16 (is_f32 == 0) or f32.  Writes the UNIQUE block keys
 * (int32 triples) sorted lexicographically (x,y,z) into keys_out (capacity
 * max_keys triples) and returns their number, or -1 if capacity is exceeded.
 * The reference's output order is undefined; sorted here for set comparison.
 * PINNED: identical block set to the reference's own DepthTouchCPU compiled unmodified
 * (oracle/ref_shim/ref_shim_vbg.cpp; tests/test_oracle_vs_ref_vbg.py). */
int64_t orc_depth_touch(const void* depth, int is_f32, int rows, int cols,
                        const double K[9], const double extrinsic[16],
                        int resolution, float voxel_size, float sdf_trunc,
                        float depth_scale, float depth_max, int stride,
                        int32_t* keys_out, int64_t max_keys);

/* t/geometry/kernel/VoxelBlockGridImpl.h:151-308 IntegrateCPU for the
 * slam::Model layout (tsdf f32, weight u16, color u16x3;  Model.cpp:28-35) with
 * u16 depth / u8 color or f32 depth / f32 color inputs.
 * block_keys: capacity x 3 int32, buf_indices: n_blocks int32 into the value
 * buffers (tsdf [cap][res^3], weight [cap][res^3], color [cap][res^3][3]).
 * color / color_buf may be NULL (depth-only overload, VoxelBlockGrid.cpp:269-278).
 * PINNED bit-exactly (tsdf, weight, colour; both input dtypes; with and without colour)
 * against the reference's own IntegrateCPU compiled unmodified (tests/test_oracle_vs_ref_vbg.py). */
void orc_tsdf_integrate(const void* depth, const void* color, int inputs_f32,
                        int rows, int cols, const int32_t* buf_indices,
                        int64_t n_blocks, const int32_t* block_keys,
                        float* tsdf_buf, uint16_t* weight_buf,
                        uint16_t* color_buf, const double depth_K[9],
                        const double color_K[9], const double extrinsic[16],
                        int resolution, float voxel_size, float sdf_trunc,
                        float depth_scale, float depth_max);
/* The same for the reference's second value layout (weight Float32, colour Float32: the other two IntegrateCPU /
 * IntegrateCUDA instantiations); pinned bit-exactly against IntegrateCPU<.., float, float> in oracle/_ref. */
void orc_tsdf_integrate_f32_values(const void* depth, const void* color, int inputs_f32,
                                   int rows, int cols, const int32_t* buf_indices,
                                   int64_t n_blocks, const int32_t* block_keys,
                                   float* tsdf_buf, float* weight_buf,
                                   float* color_buf, const double depth_K[9],
                                   const double color_K[9], const double extrinsic[16],
                                   int resolution, float voxel_size, float sdf_trunc,
                                   float depth_scale, float depth_max);

/* Set-semantics activate, core/hashmap/HashMap.cpp:166-197 +
 * CPU/TBBHashBackend.h:173-227: keys n x 3; masks[i] = 1 for exactly one
 * inserter of every key not yet present; buf_indices[i] = slot of key i
 * (slots assigned in first-seen order starting at *size).  Slot ORDER is
 * implementation-defined in the reference (parity unpinned for buf values).
 * table_keys: capacity x 3 key buffer; *size updated.  Returns 0 or -1 if
 * capacity would be exceeded.  Pinned (set semantics):
 * tests/t/geometry/VoxelBlockGrid.cpp:211-219, tests/core/HashMap.cpp:92-360. */
int orc_hashmap_activate(int32_t* table_keys, int64_t capacity, int64_t* size,
                         const int32_t* keys, int64_t n, int32_t* buf_indices,
                         uint8_t* masks);

/* VoxelBlockGrid ray casting (t/geometry/kernel/VoxelBlockGridImpl.h:310-555,
 * 578-1120; VoxelBlockGrid.cpp:328-402).  range: [h/down][w/down][2] f32 (min, max).
 * PINNED bit-exactly against the reference's own EstimateRangeCPU / RayCastCPU<float,u16,u16>, compiled
 * unmodified from /root/reference (oracle/ref_shim/ref_shim_vbg.cpp; tests/test_oracle_vs_ref_vbg.py),
 * and through analytic properties of the rendered scene (tests/test_oracle_raycast.py). */
void orc_estimate_range(const int32_t* block_keys, int64_t n, const double K[9],
                        const double E[16], int h, int w, int down_factor,
                        int resolution, float voxel_size, float depth_min,
                        float depth_max, float* range);
/* Outputs (any may be NULL): depth [h][w], vertex/color/normal [h][w][3], index (int64) /
 * mask (u8) / interp_ratio(_dx,_dy,_dz) [h][w][8].  table_keys: [size][3] block keys in slot
 * order; tsdf/weight/color buffers in the slam::Model layout (see orc_tsdf_integrate). */
void orc_ray_cast(const int32_t* table_keys, int64_t size, const float* tsdf_buf,
                  const uint16_t* weight_buf, const uint16_t* color_buf,
                  const float* range, const double K[9], const double E[16], int h,
                  int w, int resolution, float voxel_size, float depth_scale,
                  float depth_min, float depth_max, float weight_threshold,
                  float trunc_voxel_multiplier, int range_map_down_factor,
                  float* depth_out, float* vertex_out, float* color_out,
                  float* normal_out, int64_t* index_out, uint8_t* mask_out,
                  float* ratio_out, float* ratio_dx_out, float* ratio_dy_out,
                  float* ratio_dz_out);

/* ---------------------------------------------------------- RGB-D odometry
 * t/pipelines/odometry/RGBDOdometry.cpp (driver), t/pipelines/kernel/RGBDOdometry{CPU.cpp,JacobianImpl.h}
 * (PointToPlane method), t/geometry/kernel/ImageImpl.h (pyramid kernels); see odometry_oracle.c.
 * Images are row-major [rows][cols](x3).  orc_filter_bilateral_f32 is PARITY UNPINNED (NPP/IPP). */
void orc_clip_transform(const void* src, int src_is_f32, int rows, int cols, float scale,
                        float min_value, float max_value, float clip_fill, float* dst);
void orc_pyr_down_depth(const float* src, int rows, int cols, float depth_diff,
                        float invalid_fill, float* dst /* [rows/2][cols/2] */);
void orc_create_vertex_map(const float* depth, int rows, int cols, const double K[9],
                           float invalid_fill, float* vertex);
void orc_create_normal_map(const float* vertex, int rows, int cols, float invalid_fill, float* normal);
void orc_filter_bilateral_f32(const float* src, int rows, int cols, int kernel_size,
                              float value_sigma, float dist_sigma, float* dst);
float orc_huber_deriv(float r, float delta);
float orc_huber_loss(float r, float delta);
int orc_odometry_jacobian_p2plane(int x, int y, float depth_outlier_trunc, const float* source_vertex,
                                  const float* target_vertex, const float* target_normal, int rows, int cols,
                                  const double K[9], const double T[16], float J[6], float* r);
void orc_odometry_p2plane_sums(const float* source_vertex, const float* target_vertex,
                               const float* target_normal, int rows, int cols, const double K[9],
                               const double T[16], float depth_outlier_trunc, float depth_huber_delta,
                               double sums29[29], double abs29[29]);
int orc_compute_odometry_result_p2plane(const float* source_vertex, const float* target_vertex,
                                        const float* target_normal, int rows, int cols,
                                        const double K[9], const double T[16],
                                        float depth_outlier_trunc, float depth_huber_delta,
                                        double delta_T[16], double* inlier_rmse, double* fitness);
int orc_rgbd_odometry_multi_scale_p2plane(const void* source_depth, const void* target_depth, int depth_is_f32,
                                          int rows, int cols, const double K[9], const double init_T[16],
                                          float depth_scale, float depth_max, int n_levels,
                                          const int* max_iteration, const double* relative_rmse,
                                          const double* relative_fitness, float depth_outlier_trunc,
                                          float depth_huber_delta, double T_out[16], double* inlier_rmse,
                                          double* fitness, double* per_iter, int* executed);

int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
