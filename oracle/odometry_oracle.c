/*
 * odometry_oracle.c — CPU restatement of Open3D's RGB-D odometry, PointToPlane method
 * (SURVEY.md 8f #2): the depth-image pyramid kernels, the per-pixel Jacobian/29-sum
 * reduction and the multi-scale driver that slam::Model::TrackFrameToModel runs.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Never imported by the product.
 * Compile with -ffp-contract=off.  Reference paths are relative to cpp/open3d/.
 *
 * Pinned bit-exactly against the reference's own code through oracle/ref_shim:
 * GetJacobianPointToPlane / HuberDeriv / HuberLoss (RGBDOdometryJacobianImpl.h) and the
 * whole functions ClipTransformCPU, PyrDownDepthCPU, CreateVertexMapCPU, CreateNormalMapCPU
 * (t/geometry/kernel/ImageImpl.h).  NOT pinnable: Image::FilterBilateral, which upstream
 * forwards to NPP (CUDA) / IPP (CPU) — closed third-party libraries absent here; see
 * orc_filter_bilateral_f32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"
#include "geometry_indexer.h"

/* t/geometry/kernel/ImageImpl.h:86-120 ClipTransform: out = in / scale; <= min or >= max -> fill */
void orc_clip_transform(const void* src, int src_is_f32, int rows, int cols, float scale,
                        float min_value, float max_value, float clip_fill, float* dst) {
    const int64_t n = (int64_t)rows * cols;
    for (int64_t i = 0; i < n; ++i) {
        const float in = src_is_f32 ? ((const float*)src)[i] : (float)((const uint16_t*)src)[i];
        float out = in / scale;
        out = out <= min_value ? clip_fill : out;
        out = out >= max_value ? clip_fill : out;
        dst[i] = out;
    }
}

/* ImageImpl.h:122-198 PyrDownDepth: 5x5 Gaussian {.375,.25,.0625} over the neighbours within
 * depth_diff of the centre; comparisons are written as upstream so that NaN fills behave the same
 * (NaN == NaN is false: a NaN centre falls through to w_sum == 0 -> invalid_fill). */
void orc_pyr_down_depth(const float* src, int rows, int cols, float depth_diff,
                        float invalid_fill, float* dst) {
    const int rows_down = rows / 2, cols_down = cols / 2;
    const float gweights[3] = {0.375f, 0.25f, 0.0625f};
    for (int y = 0; y < rows_down; ++y)
        for (int x = 0; x < cols_down; ++x) {
            const int y_src = 2 * y, x_src = 2 * x;
            const float v_center = src[(int64_t)y_src * cols + x_src];
            float* out = dst + (int64_t)y * cols_down + x;
            if (v_center == invalid_fill) {
                *out = invalid_fill;
                continue;
            }
            const int x_min = x_src - 2 > 0 ? x_src - 2 : 0, y_min = y_src - 2 > 0 ? y_src - 2 : 0;
            const int x_max = x_src + 2 < cols - 1 ? x_src + 2 : cols - 1;
            const int y_max = y_src + 2 < rows - 1 ? y_src + 2 : rows - 1;
            float v_sum = 0, w_sum = 0;
            for (int yk = y_min; yk <= y_max; ++yk)
                for (int xk = x_min; xk <= x_max; ++xk) {
                    const float v = src[(int64_t)yk * cols + xk];
                    const int dy = abs(yk - y_src), dx = abs(xk - x_src);
                    if (v != invalid_fill && fabsf(v - v_center) < depth_diff) {
                        const float w = gweights[dx] * gweights[dy];
                        v_sum += w * v;
                        w_sum += w;
                    }
                }
            *out = w_sum == 0 ? invalid_fill : v_sum / w_sum;
        }
}

static inline int is_invalid(float v, float invalid_fill) { /* ImageImpl.h:227-231 */
    if (isinf(invalid_fill)) return isinf(v);
    if (isnan(invalid_fill)) return isnan(v);
    return v == invalid_fill;
}

/* ImageImpl.h:200-248 CreateVertexMap (TransformIndexer(intrinsics, Eye)) */
void orc_create_vertex_map(const float* depth, int rows, int cols, const double K[9],
                           float invalid_fill, float* vertex) {
    const double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    xform_indexer ti;
    xi_init(&ti, K, eye, 1.0f);
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const float d = depth[(int64_t)y * cols + x];
            float* v = vertex + 3 * ((int64_t)y * cols + x);
            if (!is_invalid(d, invalid_fill)) xi_unproject(&ti, (float)x, (float)y, d, v + 0, v + 1, v + 2);
            else v[0] = v[1] = v[2] = invalid_fill;
        }
}

/* ImageImpl.h:249-315 CreateNormalMap: cross product of the forward differences, last row / column invalid */
void orc_create_normal_map(const float* vertex, int rows, int cols, float invalid_fill, float* normal) {
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float* n = normal + 3 * ((int64_t)y * cols + x);
            if (y < rows - 1 && x < cols - 1) {
                const float* v00 = vertex + 3 * ((int64_t)y * cols + x);
                const float* v10 = vertex + 3 * ((int64_t)y * cols + x + 1);
                const float* v01 = vertex + 3 * ((int64_t)(y + 1) * cols + x);
                if ((v00[0] == invalid_fill && v00[1] == invalid_fill && v00[2] == invalid_fill) ||
                    (v01[0] == invalid_fill && v01[1] == invalid_fill && v01[2] == invalid_fill) ||
                    (v10[0] == invalid_fill && v10[1] == invalid_fill && v10[2] == invalid_fill)) {
                    n[0] = n[1] = n[2] = invalid_fill;
                    continue;
                }
                const float dx0 = v01[0] - v00[0], dy0 = v01[1] - v00[1], dz0 = v01[2] - v00[2];
                const float dx1 = v10[0] - v00[0], dy1 = v10[1] - v00[1], dz1 = v10[2] - v00[2];
                n[0] = dy0 * dz1 - dz0 * dy1;
                n[1] = dz0 * dx1 - dx0 * dz1;
                n[2] = dx0 * dy1 - dy0 * dx1;
                float norm = sqrtf(n[0] * n[0] + n[1] * n[1] + n[2] * n[2]);
                norm = norm > 1e-5f ? norm : 1e-5f;
                n[0] /= norm;
                n[1] /= norm;
                n[2] /= norm;
            } else {
                n[0] = n[1] = n[2] = invalid_fill;
            }
        }
}

/* Image::FilterBilateral (t/geometry/Image.cpp:248-285) -> npp::FilterBilateral
 * (t/geometry/kernel/NPPImage.cpp:319-376): nppiFilterBilateralGaussBorder_32f_C1R with radius
 * kernel_size/2, step 1, nValSquareSigma = value_sigma^2, nPosSquareSigma = dist_sigma^2, replicated
 * border.  PARITY UNPINNED — NPP is closed source and absent here; this restates NPP's documented
 * definition: w = exp(-(dx^2+dy^2)/(2 nPosSquareSigma)) * exp(-(v - v_c)^2/(2 nValSquareSigma)),
 * out = sum(w v)/sum(w), evaluated in f32 in row-major window order (NaN inputs propagate). */
void orc_filter_bilateral_f32(const float* src, int rows, int cols, int kernel_size,
                              float value_sigma, float dist_sigma, float* dst) {
    const int r = kernel_size / 2;
    const float val2 = 2.0f * (value_sigma * value_sigma), pos2 = 2.0f * (dist_sigma * dist_sigma);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            const float vc = src[(int64_t)y * cols + x];
            float v_sum = 0, w_sum = 0;
            for (int dy = -r; dy <= r; ++dy)
                for (int dx = -r; dx <= r; ++dx) {
                    int yy = y + dy, xx = x + dx;
                    yy = yy < 0 ? 0 : (yy > rows - 1 ? rows - 1 : yy);
                    xx = xx < 0 ? 0 : (xx > cols - 1 ? cols - 1 : xx);
                    const float v = src[(int64_t)yy * cols + xx];
                    const float dv = v - vc;
                    const float w = expf(-((float)(dx * dx + dy * dy)) / pos2) * expf(-(dv * dv) / val2);
                    v_sum += w * v;
                    w_sum += w;
                }
            dst[(int64_t)y * cols + x] = v_sum / w_sum;
        }
}

/* t/pipelines/kernel/RGBDOdometryJacobianImpl.h:29-37.  Sign() takes an int (GeometryMacros.h:92):
 * the residual is truncated to an integer first, so for |r| < 1 the derivative beyond delta is 0. */
static inline int isign_(int x) { return (x > 0) ? 1 : ((x < 0) ? -1 : 0); }
float orc_huber_deriv(float r, float delta) {
    const float abs_r = fabsf(r);
    return abs_r < delta ? r : delta * isign_((int)r);
}
float orc_huber_loss(float r, float delta) {
    const float abs_r = fabsf(r);
    return abs_r < delta ? 0.5 * r * r : delta * abs_r - 0.5 * delta * delta; /* double intermediates, as written */
}

/* RGBDOdometryJacobianImpl.h:106-160 GetJacobianPointToPlane */
int orc_odometry_jacobian_p2plane(int x, int y, float depth_outlier_trunc, const float* source_vertex,
                                  const float* target_vertex, const float* target_normal, int rows, int cols,
                                  const double K[9], const double T[16], float J[6], float* r) {
    xform_indexer ti;
    xi_init(&ti, K, T, 1.0f);
    const float* sv = source_vertex + 3 * ((int64_t)y * cols + x);
    if (isnan(sv[0])) return 0;
    float p[3], u, v;
    xi_rigid(&ti, sv[0], sv[1], sv[2], &p[0], &p[1], &p[2]);
    xi_project(&ti, p[0], p[1], p[2], &u, &v);
    u = roundf(u);
    v = roundf(v);
    if (p[2] < 0 || !in_boundary(u, v, rows, cols)) return 0;
    const int ui = (int)u, vi = (int)v;
    const float* tv = target_vertex + 3 * ((int64_t)vi * cols + ui);
    const float* tn = target_normal + 3 * ((int64_t)vi * cols + ui);
    if (isnan(tv[0]) || isnan(tn[0])) return 0;
    *r = (p[0] - tv[0]) * tn[0] + (p[1] - tv[1]) * tn[1] + (p[2] - tv[2]) * tn[2];
    if (fabsf(*r) > depth_outlier_trunc) return 0;
    J[0] = -p[2] * tn[1] + p[1] * tn[2];
    J[1] = p[2] * tn[0] - p[0] * tn[2];
    J[2] = -p[1] * tn[0] + p[0] * tn[1];
    J[3] = tn[0];
    J[4] = tn[1];
    J[5] = tn[2];
    return 1;
}

/* RGBDOdometryCPU.cpp:290-362 / RGBDOdometryCUDA.cu:37-86: 21 J J^T (NOT Huber-weighted, as upstream),
 * 6 J * HuberDeriv(r), HuberLoss(r), inlier count.  f32 terms accumulated in f64 (the reference's order
 * is a TBB / atomic tree); abs29 (optional) = sum of |term| per slot, the scale of the 1e-5 tolerance. */
void orc_odometry_p2plane_sums(const float* source_vertex, const float* target_vertex,
                               const float* target_normal, int rows, int cols, const double K[9],
                               const double T[16], float depth_outlier_trunc, float depth_huber_delta,
                               double sums29[29], double abs29[29]) {
    double acc[29], aab[29];
    for (int k = 0; k < 29; ++k) acc[k] = aab[k] = 0.0;
    for (int y = 0; y < rows; ++y)
        for (int x = 0; x < cols; ++x) {
            float J[6], r;
            if (!orc_odometry_jacobian_p2plane(x, y, depth_outlier_trunc, source_vertex, target_vertex,
                                               target_normal, rows, cols, K, T, J, &r))
                continue;
            const float d_huber = orc_huber_deriv(r, depth_huber_delta);
            const float r_huber = orc_huber_loss(r, depth_huber_delta);
            int i = 0;
            for (int j = 0; j < 6; ++j) {
                for (int k = 0; k <= j; ++k) {
                    const float t = J[j] * J[k];
                    acc[i] += t;
                    aab[i] += fabsf(t);
                    ++i;
                }
                const float t = J[j] * d_huber;
                acc[21 + j] += t;
                aab[21 + j] += fabsf(t);
            }
            acc[27] += r_huber;
            aab[27] += fabsf(r_huber);
            acc[28] += 1;
            aab[28] += 1;
        }
    memcpy(sums29, acc, sizeof(acc));
    if (abs29) memcpy(abs29, aab, sizeof(aab));
}

/* t/pipelines/odometry/RGBDOdometry.cpp:432-459 ComputeOdometryResultPointToPlane: returns 0, or 1 when
 * inlier_count <= 0 / the system is singular (upstream LogError's). */
int orc_compute_odometry_result_p2plane(const float* source_vertex, const float* target_vertex,
                                        const float* target_normal, int rows, int cols,
                                        const double K[9], const double T[16],
                                        float depth_outlier_trunc, float depth_huber_delta,
                                        double delta_T[16], double* inlier_rmse, double* fitness) {
    double s[29], pose[6];
    orc_odometry_p2plane_sums(source_vertex, target_vertex, target_normal, rows, cols, K, T,
                              depth_outlier_trunc, depth_huber_delta, s, NULL);
    /* the 29 sums reach DecodeAndSolve6x6 as a Float32 tensor (RGBDOdometryCPU.cpp:359-361) */
    for (int k = 0; k < 29; ++k) s[k] = (double)(float)s[k];
    double residual = 0;
    int count = 0;
    if (orc_decode_and_solve_6x6(s, pose, &residual, &count) != 0) return 1;
    if (count <= 0) return 1;
    orc_pose_to_transformation(pose, delta_T);
    *inlier_rmse = (float)residual / count; /* float& inlier_residual / int, as written (:455) */
    *fitness = (double)count / (double)((int64_t)rows * cols);
    return 0;
}

static void matmul4_(const double A[16], const double B[16], double C[16]) {
    double R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            R[i * 4 + j] = s;
        }
    memcpy(C, R, sizeof(R));
}

/* RGBDOdometry.cpp:56-206 RGBDOdometryMultiScale, Method::PointToPlane.  criteria: n_levels entries,
 * coarse-to-fine order as upstream's criteria list (criteria[0] runs on the coarsest level).
 * per_iter (optional): 2 doubles (inlier_rmse, fitness of the delta) per executed iteration, -1 separators
 * are not written; *executed receives the count.  Returns 0, or 1 if an iteration failed (no inliers /
 * singular). */
int orc_rgbd_odometry_multi_scale_p2plane(const void* source_depth, const void* target_depth, int depth_is_f32,
                                          int rows, int cols, const double K[9], const double init_T[16],
                                          float depth_scale, float depth_max, int n_levels,
                                          const int* max_iteration, const double* relative_rmse,
                                          const double* relative_fitness, float depth_outlier_trunc,
                                          float depth_huber_delta, double T_out[16], double* inlier_rmse,
                                          double* fitness, double* per_iter, int* executed) {
    const int64_t n = (int64_t)rows * cols;
    float* src_d = (float*)malloc(n * sizeof(float));
    float* tgt_d = (float*)malloc(n * sizeof(float));
    float* tmp = (float*)malloc(n * sizeof(float));
    float **sv = calloc(n_levels, sizeof(float*)), **tv = calloc(n_levels, sizeof(float*)),
          **tn = calloc(n_levels, sizeof(float*));
    double (*Ks)[9] = malloc(sizeof(double[9]) * n_levels);
    int *lr = malloc(sizeof(int) * n_levels), *lc = malloc(sizeof(int) * n_levels);
    if (!src_d || !tgt_d || !tmp || !sv || !tv || !tn || !Ks || !lr || !lc) abort();
    /* :84-88 ClipTransform(depth_scale, 0, depth_max, NAN) */
    orc_clip_transform(source_depth, depth_is_f32, rows, cols, depth_scale, 0.0f, depth_max, NAN, src_d);
    orc_clip_transform(target_depth, depth_is_f32, rows, cols, depth_scale, 0.0f, depth_max, NAN, tgt_d);
    double Kp[9];
    memcpy(Kp, K, sizeof(Kp));
    int r = rows, c = cols;
    for (int i = 0; i < n_levels; ++i) { /* :132-163 */
        const int L = n_levels - 1 - i;
        const int64_t m = (int64_t)r * c;
        sv[L] = malloc(3 * m * sizeof(float));
        tv[L] = malloc(3 * m * sizeof(float));
        tn[L] = malloc(3 * m * sizeof(float));
        float* tsm = malloc(3 * m * sizeof(float));
        if (!sv[L] || !tv[L] || !tn[L] || !tsm) abort();
        orc_create_vertex_map(src_d, r, c, Kp, NAN, sv[L]);
        orc_create_vertex_map(tgt_d, r, c, Kp, NAN, tv[L]);
        orc_filter_bilateral_f32(tgt_d, r, c, 5, 5.0f, 10.0f, tmp);
        orc_create_vertex_map(tmp, r, c, Kp, NAN, tsm);
        orc_create_normal_map(tsm, r, c, NAN, tn[L]);
        free(tsm);
        memcpy(Ks[L], Kp, sizeof(Kp));
        lr[L] = r;
        lc[L] = c;
        if (i != n_levels - 1) {
            orc_pyr_down_depth(src_d, r, c, depth_outlier_trunc * 2, NAN, tmp);
            memcpy(src_d, tmp, (size_t)(r / 2) * (c / 2) * sizeof(float));
            orc_pyr_down_depth(tgt_d, r, c, depth_outlier_trunc * 2, NAN, tmp);
            memcpy(tgt_d, tmp, (size_t)(r / 2) * (c / 2) * sizeof(float));
            r /= 2;
            c /= 2;
            for (int k = 0; k < 9; ++k) Kp[k] /= 2; /* :159-160 intrinsics_pyr /= 2; [-1][-1] = 1 */
            Kp[8] = 1;
        }
    }
    double T[16];
    memcpy(T, init_T, sizeof(T));
    double res_rmse = 0.0, res_fitness = 1.0; /* :165 */
    int rc = 0, done = 0;
    for (int i = 0; i < n_levels && rc == 0; ++i)
        for (int iter = 0; iter < max_iteration[i]; ++iter) {
            double dT[16], d_rmse = 0, d_fit = 0;
            if (orc_compute_odometry_result_p2plane(sv[i], tv[i], tn[i], lr[i], lc[i], Ks[i], T, depth_outlier_trunc,
                                                    depth_huber_delta, dT, &d_rmse, &d_fit) != 0) {
                rc = 1;
                break;
            }
            matmul4_(dT, T, T); /* :175-176 */
            if (per_iter) {
                per_iter[2 * done] = d_rmse;
                per_iter[2 * done + 1] = d_fit;
            }
            ++done;
            if (fabs(res_fitness - d_fit) / res_fitness < relative_fitness[i] &&
                fabs(res_rmse - d_rmse) / res_rmse < relative_rmse[i]) /* :181-189 */
                break;
            res_rmse = d_rmse;
            res_fitness = d_fit;
        }
    memcpy(T_out, T, sizeof(T));
    *inlier_rmse = res_rmse;
    *fitness = res_fitness;
    if (executed) *executed = done;
    for (int L = 0; L < n_levels; ++L) {
        free(sv[L]);
        free(tv[L]);
        free(tn[L]);
    }
    free(sv);
    free(tv);
    free(tn);
    free(Ks);
    free(lr);
    free(lc);
    free(src_d);
    free(tgt_d);
    free(tmp);
    return rc;
}
