/*
 * icp_oracle.c — CPU restatement of Open3D's tensor ICP hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Never imported by the product.
 * Compile with -ffp-contract=off: every f32 expression below is evaluated
 * operation by operation, in the source order of the reference; the only fused
 * operations are the explicit fmaf() calls in dist2_f32().
 *
 * Reference paths are relative to cpp/open3d/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* bench.py's CPU legs call this with the host's core count: launchers such as torchrun export
 * OMP_NUM_THREADS=1, which would otherwise time the "all host threads" baseline on one core. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* ------------------------------------------------------------------ hashes */

/* core/hashmap/Dispatch.h:67-81 */
uint64_t orc_minivec_hash_i32x3(int32_t x, int32_t y, int32_t z) {
    const int32_t key[3] = {x, y, z};
    uint64_t hash = UINT64_C(14695981039346656037);
    for (int i = 0; i < 3; ++i) {
        hash ^= (uint64_t)(int64_t)key[i]; /* static_cast<uint64_t>(int) sign-extends */
        hash *= UINT64_C(1099511628211);
    }
    return hash;
}

/* core/nns/NeighborSearchCommon.h:31-37: `x * 73856096 ^ y * 193649663 ^
 * z * 83492791` evaluated in int (wrapping; done here in uint32 to avoid UB),
 * then converted int -> size_t (sign extension). */
uint64_t orc_spatial_hash(int32_t x, int32_t y, int32_t z) {
    uint32_t h = ((uint32_t)x * 73856096u) ^ ((uint32_t)y * 193649663u) ^
                 ((uint32_t)z * 83492791u);
    return (uint64_t)(int64_t)(int32_t)h;
}

/* core/nns/NeighborSearchCommon.h:44-52 */
void orc_compute_voxel_index_f32(const float pos[3], float inv_voxel_size,
                                 int32_t out[3]) {
    for (int i = 0; i < 3; ++i) {
        float ref = pos[i] * inv_voxel_size;
        out[i] = (int32_t)floorf(ref);
    }
}

/* ------------------------------------------------------------ robust kernel */

/* t/pipelines/registration/RobustKernelImpl.h:35-115.  The reference lambdas
 * take and return scalar_t; double literals promote sub-expressions to double
 * exactly as written there. */
double orc_robust_weight_f64(int method, double scale, double shape,
                             double r) {
    switch (method) {
        case 0: /* L2Loss :41-46 */
            return 1.0;
        case 1: /* L1Loss :47-52 */
            return 1.0 / fabs(r);
        case 2: /* HuberLoss :53-58 */
            return scale / fmax(fabs(r), scale);
        case 3: /* CauchyLoss :59-64 */
            return 1.0 / (1.0 + (r / scale) * (r / scale));
        case 4: { /* GMLoss :65-70 */
            double s = scale + r * r;
            return scale / (s * s);
        }
        case 5: { /* TukeyLoss :71-77 */
            double q = fmin(1.0, fabs(r) / scale);
            double v = 1.0 - q * q;
            return v * v;
        }
        case 6: /* GeneralizedLoss :78-112.  open3d::IsClose (GeometryMacros.h:58-63) is a
                 * RELATIVE test: x > (1-rtol) y && x < (1+rtol) y, which can never hold for
                 * y == 0, so the "shape ~ 0" branch of the reference is dead code. */
            if (shape > (1.0 - 1e-3) * 2.0 && shape < (1.0 + 1e-3) * 2.0) {
                return 1.0 / (scale * scale);
            } else if (shape > (1.0 - 1e-3) * 0.0 && shape < (1.0 + 1e-3) * 0.0) {
                return 2.0 / (r * r + 2 * scale * scale);
            } else if (shape < -1e7) {
                return exp(((r / scale) * (r / scale)) / (-2.0)) /
                       (scale * scale);
            } else {
                return pow(((r / scale) * (r / scale)) / fabs(shape - 2.0) + 1,
                           (shape / 2.0) - 1.0) /
                       (scale * scale);
            }
        default:
            return NAN;
    }
}

float orc_robust_weight_f32(int method, double scale_d, double shape,
                            float r) {
    const float scale = (float)scale_d; /* :38 static_cast<scalar_t> */
    switch (method) {
        case 0:
            return (float)1.0;
        case 1:
            return (float)(1.0 / fabsf(r));
        case 2:
            return scale / fmaxf(fabsf(r), scale);
        case 3: {
            float q = r / scale;
            return (float)(1.0 / (1.0 + (double)(q * q)));
        }
        case 4: {
            float s = scale + r * r;
            return scale / (s * s);
        }
        case 5: {
            float q = fminf((float)1.0, fabsf(r) / scale);
            double v = 1.0 - (double)(q * q);
            return (float)(v * v);
        }
        case 6:
            if (shape > (1.0 - 1e-3) * 2.0 && shape < (1.0 + 1e-3) * 2.0) {
                double const_val = 1.0 / (double)(scale * scale);
                return (float)const_val;
            } else if (shape > (1.0 - 1e-3) * 0.0 && shape < (1.0 + 1e-3) * 0.0) {
                return (float)(2.0 / (r * r + 2 * (scale * scale)));
            } else if (shape < -1e7) {
                float q = r / scale;
                return (float)(exp((double)(q * q) / (-2.0)) /
                               (double)(scale * scale));
            } else {
                float q = r / scale;
                return (float)(pow(((double)(q * q) / fabs(shape - 2.0) + 1),
                                   ((shape / 2.0) - 1.0)) /
                               (double)(scale * scale));
            }
        default:
            return NAN;
    }
}

/* ------------------------------------------------------------------- search */

/* FixedRadiusSearchImpl.cuh:41-60 NeighborTest (L2): d = p1 - p2; d.dot(d).
 * Canonical arithmetic for this repo: fma(dz,dz, fma(dy,dy, dx*dx)) with
 * d = point - query (the CUDA reference is compiled with -fmad=true and would
 * contract similarly; nanoflann's own order is unobservable offline). */
static inline float dist2_f32(const float* p, const float* q) {
    const float dx = p[0] - q[0];
    const float dy = p[1] - q[1];
    const float dz = p[2] - q[2];
    return fmaf(dz, dz, fmaf(dy, dy, dx * dx));
}

/* Keep the k best (dist asc, ties -> lower index) in a tiny sorted array. */
static inline void knn_insert(int32_t* bi, float* bd, int* cnt, int k,
                              int32_t idx, float d) {
    int n = *cnt;
    if (n == k) {
        if (d > bd[n - 1] || (d == bd[n - 1] && idx > bi[n - 1])) return;
        n = k - 1;
    }
    int pos = n;
    while (pos > 0 &&
           (bd[pos - 1] > d || (bd[pos - 1] == d && bi[pos - 1] > idx))) {
        bd[pos] = bd[pos - 1];
        bi[pos] = bi[pos - 1];
        --pos;
    }
    bd[pos] = d;
    bi[pos] = idx;
    *cnt = n + 1;
}

#define ORC_MAX_KNN 64

void orc_hybrid_search_bruteforce_f32(const float* points, int64_t M,
                                      const float* queries, int64_t N,
                                      double radius, int max_knn, int32_t* idx,
                                      float* dist2, int32_t* counts) {
    /* FixedRadiusSearchImpl.cuh:692 / NanoFlannImpl.h:327: T(radius) * T(radius) in T */
    const float threshold = (float)radius * (float)radius;
    if (max_knn > ORC_MAX_KNN) max_knn = ORC_MAX_KNN;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        int32_t bi[ORC_MAX_KNN];
        float bd[ORC_MAX_KNN];
        int cnt = 0;
        const float* q = queries + 3 * i;
        for (int64_t j = 0; j < M; ++j) {
            float d = dist2_f32(points + 3 * j, q);
            if (d <= threshold) knn_insert(bi, bd, &cnt, max_knn, (int32_t)j, d);
        }
        for (int k = 0; k < max_knn; ++k) {
            idx[i * max_knn + k] = k < cnt ? bi[k] : -1;
            dist2[i * max_knn + k] = k < cnt ? bd[k] : 0.0f;
        }
        counts[i] = cnt;
    }
}

typedef struct {
    double ox, oy, oz, inv_c;
    int64_t nx, ny, nz;
    int64_t* cell_start; /* nx*ny*nz + 1 */
    int32_t* order;      /* point ids sorted by cell */
    float* sorted;       /* xyz sorted by cell (3 floats per point) */
} orc_grid;

static inline int64_t cell_of(double x, double o, double inv_c, int64_t n) {
    int64_t c = (int64_t)floor((x - o) * inv_c);
    if (c < 0) c = 0;
    if (c > n - 1) c = n - 1;
    return c;
}

static int grid_build(orc_grid* g, const float* pts, int64_t M, double radius) {
    double lo[3] = {INFINITY, INFINITY, INFINITY};
    double hi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = 0; i < M; ++i) {
        for (int a = 0; a < 3; ++a) {
            double v = pts[3 * i + a];
            if (v < lo[a]) lo[a] = v;
            if (v > hi[a]) hi[a] = v;
        }
    }
    double c = radius;
    if (!(c > 0)) c = 1.0;
    for (;;) {
        g->nx = (int64_t)floor((hi[0] - lo[0]) / c) + 1;
        g->ny = (int64_t)floor((hi[1] - lo[1]) / c) + 1;
        g->nz = (int64_t)floor((hi[2] - lo[2]) / c) + 1;
        if ((double)g->nx * (double)g->ny * (double)g->nz <= 134217728.0) break;
        c *= 2.0;
    }
    g->ox = lo[0];
    g->oy = lo[1];
    g->oz = lo[2];
    g->inv_c = 1.0 / c;
    const int64_t ncell = g->nx * g->ny * g->nz;
    g->cell_start = (int64_t*)calloc((size_t)ncell + 1, sizeof(int64_t));
    g->order = (int32_t*)malloc((size_t)(M > 0 ? M : 1) * sizeof(int32_t));
    g->sorted = (float*)malloc((size_t)(M > 0 ? M : 1) * 3 * sizeof(float));
    int64_t* key = (int64_t*)malloc((size_t)(M > 0 ? M : 1) * sizeof(int64_t));
    if (!g->cell_start || !g->order || !g->sorted || !key) return -1;
    for (int64_t i = 0; i < M; ++i) {
        int64_t ix = cell_of(pts[3 * i + 0], g->ox, g->inv_c, g->nx);
        int64_t iy = cell_of(pts[3 * i + 1], g->oy, g->inv_c, g->ny);
        int64_t iz = cell_of(pts[3 * i + 2], g->oz, g->inv_c, g->nz);
        key[i] = (iz * g->ny + iy) * g->nx + ix;
        g->cell_start[key[i] + 1]++;
    }
    for (int64_t cidx = 0; cidx < ncell; ++cidx)
        g->cell_start[cidx + 1] += g->cell_start[cidx];
    int64_t* cursor = (int64_t*)malloc((size_t)ncell * sizeof(int64_t));
    if (!cursor) return -1;
    memcpy(cursor, g->cell_start, (size_t)ncell * sizeof(int64_t));
    for (int64_t i = 0; i < M; ++i) { /* stable: ascending index inside a cell */
        int64_t p = cursor[key[i]]++;
        g->order[p] = (int32_t)i;
        g->sorted[3 * p + 0] = pts[3 * i + 0];
        g->sorted[3 * p + 1] = pts[3 * i + 1];
        g->sorted[3 * p + 2] = pts[3 * i + 2];
    }
    free(cursor);
    free(key);
    return 0;
}

static void grid_free(orc_grid* g) {
    free(g->cell_start);
    free(g->order);
    free(g->sorted);
}

static void grid_query(const orc_grid* g, const float* q, double radius,
                       float threshold, int max_knn, int32_t* bi, float* bd,
                       int* cnt_out) {
    /* Monotone binning: every point with |p - q| <= r (per axis) lies in a
     * cell of [cell(q - r'), cell(q + r')], r' = r (1 + 1e-6). */
    const double rr = radius * (1.0 + 1e-6);
    int64_t x0 = cell_of(q[0] - rr, g->ox, g->inv_c, g->nx);
    int64_t x1 = cell_of(q[0] + rr, g->ox, g->inv_c, g->nx);
    int64_t y0 = cell_of(q[1] - rr, g->oy, g->inv_c, g->ny);
    int64_t y1 = cell_of(q[1] + rr, g->oy, g->inv_c, g->ny);
    int64_t z0 = cell_of(q[2] - rr, g->oz, g->inv_c, g->nz);
    int64_t z1 = cell_of(q[2] + rr, g->oz, g->inv_c, g->nz);
    int cnt = 0;
    for (int64_t iz = z0; iz <= z1; ++iz) {
        for (int64_t iy = y0; iy <= y1; ++iy) {
            const int64_t row = (iz * g->ny + iy) * g->nx;
            const int64_t s = g->cell_start[row + x0];
            const int64_t e = g->cell_start[row + x1 + 1];
            for (int64_t j = s; j < e; ++j) {
                float d = dist2_f32(g->sorted + 3 * j, q);
                if (d <= threshold)
                    knn_insert(bi, bd, &cnt, max_knn, g->order[j], d);
            }
        }
    }
    *cnt_out = cnt;
}

void orc_hybrid_search_f32(const float* points, int64_t M,
                           const float* queries, int64_t N, double radius,
                           int max_knn, int32_t* idx, float* dist2,
                           int32_t* counts) {
    if (max_knn > ORC_MAX_KNN) max_knn = ORC_MAX_KNN;
    if (M <= 0) {
        for (int64_t i = 0; i < N; ++i) {
            for (int k = 0; k < max_knn; ++k) {
                idx[i * max_knn + k] = -1;
                dist2[i * max_knn + k] = 0.0f;
            }
            counts[i] = 0;
        }
        return;
    }
    orc_grid g;
    if (grid_build(&g, points, M, radius) != 0) abort();
    const float threshold = (float)radius * (float)radius;
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < N; ++i) {
        int32_t bi[ORC_MAX_KNN];
        float bd[ORC_MAX_KNN];
        int cnt = 0;
        grid_query(&g, queries + 3 * i, radius, threshold, max_knn, bi, bd, &cnt);
        for (int k = 0; k < max_knn; ++k) {
            idx[i * max_knn + k] = k < cnt ? bi[k] : -1;
            dist2[i * max_knn + k] = k < cnt ? bd[k] : 0.0f;
        }
        counts[i] = cnt;
    }
    grid_free(&g);
}

/* -------------------------------------------------------- pose estimation */

/* RegistrationImpl.h:251-287 GetJacobianPointToPlane<float>. */
static inline int jacobian_p2plane_f32(int64_t i, const float* src,
                                       const float* tgt, const float* nrm,
                                       const int64_t* corr, float J[6],
                                       float* r) {
    if (corr[i] == -1) return 0;
    const int64_t t = 3 * corr[i];
    const int64_t s = 3 * i;
    const float sx = src[s + 0], sy = src[s + 1], sz = src[s + 2];
    const float tx = tgt[t + 0], ty = tgt[t + 1], tz = tgt[t + 2];
    const float nx = nrm[t + 0], ny = nrm[t + 1], nz = nrm[t + 2];
    *r = (sx - tx) * nx + (sy - ty) * ny + (sz - tz) * nz;
    J[0] = nz * sy - ny * sz;
    J[1] = nx * sz - nz * sx;
    J[2] = ny * sx - nx * sy;
    J[3] = nx;
    J[4] = ny;
    J[5] = nz;
    return 1;
}

static inline int jacobian_p2plane_f64(int64_t i, const double* src,
                                       const double* tgt, const double* nrm,
                                       const int64_t* corr, double J[6],
                                       double* r) {
    if (corr[i] == -1) return 0;
    const int64_t t = 3 * corr[i];
    const int64_t s = 3 * i;
    const double sx = src[s + 0], sy = src[s + 1], sz = src[s + 2];
    const double tx = tgt[t + 0], ty = tgt[t + 1], tz = tgt[t + 2];
    const double nx = nrm[t + 0], ny = nrm[t + 1], nz = nrm[t + 2];
    *r = (sx - tx) * nx + (sy - ty) * ny + (sz - tz) * nz;
    J[0] = nz * sy - ny * sz;
    J[1] = nx * sz - nz * sx;
    J[2] = ny * sx - nx * sy;
    J[3] = nx;
    J[4] = ny;
    J[5] = nz;
    return 1;
}

/* RegistrationCPU.cpp:30-90: per valid i
 *   A[idx(j,k)] += J[j]*w*J[k] (k<=j, row-major packed), A[21+j] += J[j]*w*r,
 *   A[27] += r, A[28] += 1.  Products are evaluated in scalar_t (f32). */
void orc_pose_p2plane_sums_f32(const float* src, const float* tgt,
                               const float* nrm, const int64_t* corr, int64_t n,
                               int method, double scale, double shape,
                               double sums64[29], float sums32[29],
                               double abs64[29]) {
    double acc[29], aabs[29];
    float acc32[29];
    for (int k = 0; k < 29; ++k) acc[k] = aabs[k] = 0.0, acc32[k] = 0.0f;

#pragma omp parallel
    {
        double la[29], lb[29];
        for (int k = 0; k < 29; ++k) la[k] = lb[k] = 0.0;
#pragma omp for schedule(static) nowait
        for (int64_t i = 0; i < n; ++i) {
            float J[6], r = 0;
            int valid = jacobian_p2plane_f32(i, src, tgt, nrm, corr, J, &r);
            if (!valid) continue;
            float w = orc_robust_weight_f32(method, scale, shape, r);
            int s = 0;
            for (int j = 0; j < 6; ++j) {
                for (int k = 0; k <= j; ++k) {
                    float term = J[j] * w * J[k];
                    la[s] += term;
                    lb[s] += fabs((double)term);
                    ++s;
                }
                float tb = J[j] * w * r;
                la[21 + j] += tb;
                lb[21 + j] += fabs((double)tb);
            }
            la[27] += r;
            lb[27] += fabs((double)r);
            la[28] += 1;
            lb[28] += 1;
        }
#pragma omp critical
        for (int k = 0; k < 29; ++k) acc[k] += la[k], aabs[k] += lb[k];
    }
    if (sums32) { /* sequential f32 accumulation in index order */
        for (int64_t i = 0; i < n; ++i) {
            float J[6], r = 0;
            if (!jacobian_p2plane_f32(i, src, tgt, nrm, corr, J, &r)) continue;
            float w = orc_robust_weight_f32(method, scale, shape, r);
            int s = 0;
            for (int j = 0; j < 6; ++j) {
                for (int k = 0; k <= j; ++k) {
                    acc32[s] += J[j] * w * J[k];
                    ++s;
                }
                acc32[21 + j] += J[j] * w * r;
            }
            acc32[27] += r;
            acc32[28] += 1;
        }
        memcpy(sums32, acc32, sizeof(acc32));
    }
    if (sums64) memcpy(sums64, acc, sizeof(acc));
    if (abs64) memcpy(abs64, aabs, sizeof(aabs));
}

void orc_pose_p2plane_sums_f64(const double* src, const double* tgt,
                               const double* nrm, const int64_t* corr,
                               int64_t n, int method, double scale,
                               double shape, double sums64[29]) {
    double acc[29];
    for (int k = 0; k < 29; ++k) acc[k] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        double J[6], r = 0;
        if (!jacobian_p2plane_f64(i, src, tgt, nrm, corr, J, &r)) continue;
        double w = orc_robust_weight_f64(method, scale, shape, r);
        int s = 0;
        for (int j = 0; j < 6; ++j) {
            for (int k = 0; k <= j; ++k) {
                acc[s] += J[j] * w * J[k];
                ++s;
            }
            acc[21 + j] += J[j] * w * r;
        }
        acc[27] += r;
        acc[28] += 1;
    }
    memcpy(sums64, acc, sizeof(acc));
}

/* RegistrationImpl.h:413-493 GetJacobianColoredICP<float> and
 * RegistrationCPU.cpp ComputePoseColoredICPKernelCPU:
 *   A[idx(j,k)] += J_G[j]*w_G*J_G[k] + J_I[j]*w_I*J_I[k]
 *   A[21+j]     += J_G[j]*w_G*r_G   + J_I[j]*w_I*r_I
 *   A[27]       += r_G*r_G + r_I*r_I ;  A[28] += 1 */
void orc_pose_colored_sums_f32(const float* src, const float* src_colors,
                               const float* tgt, const float* nrm,
                               const float* tgt_colors, const float* tgt_grad,
                               const int64_t* corr, int64_t n,
                               double lambda_geometric, int method,
                               double scale, double shape, double sums64[29],
                               double abs64[29]) {
    /* RegistrationCPU.cpp:311-314: sqrt in f64, then cast to scalar_t */
    const float sqrt_lg = (float)sqrt(lambda_geometric);
    const float sqrt_lp = (float)sqrt(1.0 - lambda_geometric);
    double acc[29], aabs[29];
    for (int k = 0; k < 29; ++k) acc[k] = aabs[k] = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        if (corr[i] == -1) continue;
        const int64_t t = 3 * corr[i];
        const int64_t s = 3 * i;
        const float vs[3] = {src[s], src[s + 1], src[s + 2]};
        const float vt[3] = {tgt[t], tgt[t + 1], tgt[t + 2]};
        const float nt[3] = {nrm[t], nrm[t + 1], nrm[t + 2]};
        const float d = (vs[0] - vt[0]) * nt[0] + (vs[1] - vt[1]) * nt[1] +
                        (vs[2] - vt[2]) * nt[2];
        float JG[6], JI[6];
        JG[0] = sqrt_lg * (-vs[2] * nt[1] + vs[1] * nt[2]);
        JG[1] = sqrt_lg * (vs[2] * nt[0] - vs[0] * nt[2]);
        JG[2] = sqrt_lg * (-vs[1] * nt[0] + vs[0] * nt[1]);
        JG[3] = sqrt_lg * nt[0];
        JG[4] = sqrt_lg * nt[1];
        JG[5] = sqrt_lg * nt[2];
        const float rG = sqrt_lg * d;
        const float vsp[3] = {vs[0] - d * nt[0], vs[1] - d * nt[1],
                              vs[2] - d * nt[2]};
        const float is = (float)((src_colors[s] + src_colors[s + 1] +
                                  src_colors[s + 2]) /
                                 3.0);
        const float it = (float)((tgt_colors[t] + tgt_colors[t + 1] +
                                  tgt_colors[t + 2]) /
                                 3.0);
        const float dit[3] = {tgt_grad[t], tgt_grad[t + 1], tgt_grad[t + 2]};
        const float is_proj = dit[0] * (vsp[0] - vt[0]) +
                              dit[1] * (vsp[1] - vt[1]) +
                              dit[2] * (vsp[2] - vt[2]) + it;
        const float sdot = dit[0] * nt[0] + dit[1] * nt[1] + dit[2] * nt[2];
        const float dM[3] = {sdot * nt[0] - dit[0], sdot * nt[1] - dit[1],
                             sdot * nt[2] - dit[2]};
        JI[0] = sqrt_lp * (-vs[2] * dM[1] + vs[1] * dM[2]);
        JI[1] = sqrt_lp * (vs[2] * dM[0] - vs[0] * dM[2]);
        JI[2] = sqrt_lp * (-vs[1] * dM[0] + vs[0] * dM[1]);
        JI[3] = sqrt_lp * dM[0];
        JI[4] = sqrt_lp * dM[1];
        JI[5] = sqrt_lp * dM[2];
        const float rI = sqrt_lp * (is - is_proj);
        const float wG = orc_robust_weight_f32(method, scale, shape, rG);
        const float wI = orc_robust_weight_f32(method, scale, shape, rI);
        int p = 0;
        for (int j = 0; j < 6; ++j) {
            for (int k = 0; k <= j; ++k) {
                float term = JG[j] * wG * JG[k] + JI[j] * wI * JI[k];
                acc[p] += term;
                aabs[p] += fabs((double)term);
                ++p;
            }
            float tb = JG[j] * wG * rG + JI[j] * wI * rI;
            acc[21 + j] += tb;
            aabs[21 + j] += fabs((double)tb);
        }
        float rr = rG * rG + rI * rI;
        acc[27] += rr;
        aabs[27] += rr;
        acc[28] += 1;
        aabs[28] += 1;
    }
    if (sums64) memcpy(sums64, acc, sizeof(acc));
    if (abs64) memcpy(abs64, aabs, sizeof(aabs));
}

/* TransformationConverter.cpp:189-226.  AtA.Solve(-Atb) is LAPACK dgesv
 * (core/linalg/SolveCPU.cpp:24): LU with partial (row) pivoting; info > 0
 * (exact zero pivot) => Open3D raises. */
int orc_decode_and_solve_6x6(const double A[29], double pose[6],
                             double* residual, int* inlier_count) {
    double M[6][7];
    for (int j = 0; j < 6; ++j) {
        const int base = (j * (j + 1)) / 2;
        for (int k = 0; k <= j; ++k) {
            M[j][k] = A[base + k];
            M[k][j] = A[base + k];
        }
        M[j][6] = -A[21 + j];
    }
    for (int c = 0; c < 6; ++c) {
        int piv = c;
        double best = fabs(M[c][c]);
        for (int r = c + 1; r < 6; ++r) {
            if (fabs(M[r][c]) > best) {
                best = fabs(M[r][c]);
                piv = r;
            }
        }
        if (best == 0.0) {
            for (int j = 0; j < 6; ++j) pose[j] = 0.0;
            if (residual) *residual = 0.0;
            if (inlier_count) *inlier_count = 0;
            return 1;
        }
        if (piv != c) {
            for (int k = 0; k < 7; ++k) {
                double tmp = M[c][k];
                M[c][k] = M[piv][k];
                M[piv][k] = tmp;
            }
        }
        for (int r = c + 1; r < 6; ++r) {
            double f = M[r][c] / M[c][c];
            for (int k = c; k < 7; ++k) M[r][k] -= f * M[c][k];
        }
    }
    for (int r = 5; r >= 0; --r) {
        double s = M[r][6];
        for (int k = r + 1; k < 6; ++k) s -= M[r][k] * pose[k];
        pose[r] = s / M[r][r];
    }
    if (residual) *residual = A[27];
    if (inlier_count) *inlier_count = (int)A[28];
    return 0;
}

/* TransformationConverterImpl.h:22-42 (rotation) + TransformationConverter.cpp:81-104
 * (identity start, translation = pose[3:6]). */
void orc_pose_to_transformation(const double p[6], double T[16]) {
    for (int i = 0; i < 16; ++i) T[i] = 0.0;
    T[0] = T[5] = T[10] = T[15] = 1.0;
    T[0] = cos(p[2]) * cos(p[1]);
    T[1] = -1 * sin(p[2]) * cos(p[0]) + cos(p[2]) * sin(p[1]) * sin(p[0]);
    T[2] = sin(p[2]) * sin(p[0]) + cos(p[2]) * sin(p[1]) * cos(p[0]);
    T[4] = sin(p[2]) * cos(p[1]);
    T[5] = cos(p[2]) * cos(p[0]) + sin(p[2]) * sin(p[1]) * sin(p[0]);
    T[6] = -1 * cos(p[2]) * sin(p[0]) + sin(p[2]) * sin(p[1]) * cos(p[0]);
    T[8] = -1 * sin(p[1]);
    T[9] = cos(p[1]) * sin(p[0]);
    T[10] = cos(p[1]) * cos(p[0]);
    T[3] = p[3];
    T[7] = p[4];
    T[11] = p[5];
}

/* TransformImpl.h:20-45 TransformPointsKernel<float>; T cast to f32 first
 * (t/geometry/kernel/Transform.cpp:29-31). */
void orc_transform_points_f32(const double Td[16], float* pts, int64_t n) {
    float T[16];
    for (int i = 0; i < 16; ++i) T[i] = (float)Td[i];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float* p = pts + 3 * i;
        float x[4] = {T[0] * p[0] + T[1] * p[1] + T[2] * p[2] + T[3],
                      T[4] * p[0] + T[5] * p[1] + T[6] * p[2] + T[7],
                      T[8] * p[0] + T[9] * p[1] + T[10] * p[2] + T[11],
                      T[12] * p[0] + T[13] * p[1] + T[14] * p[2] + T[15]};
        p[0] = x[0] / x[3];
        p[1] = x[1] / x[3];
        p[2] = x[2] / x[3];
    }
}

/* TransformImpl.h:47-62 */
void orc_transform_normals_f32(const double Td[16], float* nrm, int64_t n) {
    float T[16];
    for (int i = 0; i < 16; ++i) T[i] = (float)Td[i];
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float* p = nrm + 3 * i;
        float x[3] = {T[0] * p[0] + T[1] * p[1] + T[2] * p[2],
                      T[4] * p[0] + T[5] * p[1] + T[6] * p[2],
                      T[8] * p[0] + T[9] * p[1] + T[10] * p[2]};
        p[0] = x[0];
        p[1] = x[1];
        p[2] = x[2];
    }
}

/* TransformationEstimation.cpp:161-194: error = sum(((s - t) * n)^2) over the
 * three components separately (element-wise Mul_, then squared and summed),
 * rmse = sqrt(error / #valid).  NOTE: this is sum over components of
 * ((s_c - t_c) n_c)^2, not (dot)^2 — restated as written. */
double orc_rmse_p2plane_f32(const float* src, const float* tgt,
                            const float* nrm, const int64_t* corr, int64_t n) {
    double err = 0.0;
    int64_t cnt = 0;
    for (int64_t i = 0; i < n; ++i) {
        if (corr[i] == -1) continue;
        const int64_t t = 3 * corr[i];
        for (int c = 0; c < 3; ++c) {
            float e = (src[3 * i + c] - tgt[t + c]) * nrm[t + c];
            e = e * e;
            err += e;
        }
        ++cnt;
    }
    return sqrt(err / (double)cnt);
}

/* ---------------------------------------------------------------- ICP loop */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

static void matmul4(const double A[16], const double B[16], double C[16]) {
    double R[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += A[i * 4 + k] * B[k * 4 + j];
            R[i * 4 + j] = s;
        }
    memcpy(C, R, sizeof(R));
}

static void eye4(double T[16]) {
    for (int i = 0; i < 16; ++i) T[i] = 0.0;
    T[0] = T[5] = T[10] = T[15] = 1.0;
}

/* Registration.cpp:24-62 ComputeRegistrationResult on a prebuilt grid. */
static void compute_registration_result(const orc_grid* g, const float* src,
                                        int64_t n, double radius,
                                        int64_t* corr, double* fitness,
                                        double* rmse, int64_t* count_out) {
    const float threshold = (float)radius * (float)radius;
    double sq = 0.0;
    int64_t cnt = 0;
#pragma omp parallel for schedule(dynamic, 1024) reduction(+ : sq, cnt)
    for (int64_t i = 0; i < n; ++i) {
        int32_t bi[2];
        float bd[2];
        int c = 0;
        grid_query(g, src + 3 * i, radius, threshold, 1, bi, bd, &c);
        corr[i] = c ? (int64_t)bi[0] : -1;
        if (c) {
            sq += (double)bd[0];
            cnt += 1;
        }
    }
    *count_out = cnt;
    if (cnt != 0) {
        *fitness = (double)cnt / (double)n;
        *rmse = sqrt(sq / (double)cnt);
    } else {
        *fitness = 0.0;
        *rmse = 0.0;
    }
}

int orc_icp_p2plane_f32(const float* source, int64_t n, const float* target,
                        const float* nrm, int64_t m, double max_corr_dist,
                        const double init_T[16], int max_iteration,
                        double rel_fitness, double rel_rmse, int method,
                        double scale, double shape, int accumulate_f64,
                        orc_icp_result* res, double* per_iter,
                        int64_t* corr_out) {
    /* Registration.cpp:398-404: clone source, transform by the initial guess */
    float* src = (float*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(float));
    int64_t* corr = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    if (!src || !corr) return -1;
    memcpy(src, source, (size_t)n * 3 * sizeof(float));
    double T[16];
    memcpy(T, init_T, sizeof(T));
    orc_transform_points_f32(T, src, n);

    orc_grid g;
    const double t_build0 = now_s();
    if (grid_build(&g, target, m, max_corr_dist) != 0) return -1;
    const double t_loop0 = now_s();

    double fitness = 0, rmse = 0, prev_fitness = 0, prev_rmse = 0;
    int64_t cnt = 0;
    int converged = 0;
    int it = 0;
    int early = 0;
    for (it = 0; it < max_iteration; ++it) {
        compute_registration_result(&g, src, n, max_corr_dist, corr, &fitness,
                                    &rmse, &cnt);
        if (cnt == 0) eye4(T); /* Registration.cpp:51-60 */
        if (fitness <= 2.2250738585072014e-308) { /* :300-306 */
            converged = 0;
            early = 1;
            break;
        }
        double s64[29];
        float s32[29];
        orc_pose_p2plane_sums_f32(src, target, nrm, corr, n, method, scale,
                                  shape, s64, accumulate_f64 ? NULL : s32, NULL);
        if (!accumulate_f64)
            for (int k = 0; k < 29; ++k) s64[k] = (double)s32[k];
        double pose[6], U[16];
        if (orc_decode_and_solve_6x6(s64, pose, NULL, NULL) != 0) {
            grid_free(&g);
            free(src);
            free(corr);
            return 1; /* singular: the reference raises */
        }
        orc_pose_to_transformation(pose, U);
        matmul4(U, T, T);                    /* :319 */
        orc_transform_points_f32(U, src, n); /* :322 */
        if (per_iter) {
            per_iter[2 * it + 0] = fitness;
            per_iter[2 * it + 1] = rmse;
        }
        if (it != 0 && fabs(prev_fitness - fitness) < rel_fitness &&
            fabs(prev_rmse - rmse) < rel_rmse) { /* :348-355 */
            converged = 1;
            break;
        }
        prev_fitness = fitness;
        prev_rmse = rmse;
    }
    /* Registration.cpp:358: the tuple carries iteration_count, which after a
     * `break` has not been incremented. */
    int iterations = it;
    (void)early;
    res->loop_seconds = now_s() - t_loop0;
    res->build_seconds = t_loop0 - t_build0;
    /* :424-431 final evaluation on the last scale */
    compute_registration_result(&g, src, n, max_corr_dist, corr, &fitness, &rmse,
                                &cnt);
    if (cnt == 0) {
        eye4(T);
        converged = 0;
    }
    res->num_iterations = iterations;
    res->converged = converged;
    res->fitness = fitness;
    res->inlier_rmse = rmse;
    memcpy(res->transformation, T, sizeof(T));
    if (corr_out) memcpy(corr_out, corr, (size_t)n * sizeof(int64_t));
    grid_free(&g);
    free(src);
    free(corr);
    return 0;
}

/* ------------------------------------------------------- VoxelDownSample */

typedef struct {
    int32_t k[3];
    int64_t i;
} vds_entry;

static int vds_cmp(const void* a, const void* b) {
    const vds_entry* x = (const vds_entry*)a;
    const vds_entry* y = (const vds_entry*)b;
    for (int c = 0; c < 3; ++c) {
        if (x->k[c] < y->k[c]) return -1;
        if (x->k[c] > y->k[c]) return 1;
    }
    return x->i < y->i ? -1 : (x->i > y->i ? 1 : 0);
}

/* t/geometry/PointCloud.cpp:496-560 */
int64_t orc_voxel_down_sample_f32(const float* pos, const float* nrm, const float* col, int64_t n,
                                  double voxel_size, float* pos_out, float* nrm_out, float* col_out,
                                  int32_t* keys_out) {
    if (n <= 0) return 0;
    vds_entry* e = (vds_entry*)malloc((size_t)n * sizeof(vds_entry));
    if (!e) return -1;
    const float vs = (float)voxel_size; /* scalar -> tensor dtype (Float32) */
    for (int64_t i = 0; i < n; ++i) {
        for (int c = 0; c < 3; ++c) e[i].k[c] = (int32_t)floorf(pos[3 * i + c] / vs); /* :506-507 */
        e[i].i = i;
    }
    qsort(e, (size_t)n, sizeof(vds_entry), vds_cmp);
    int64_t m = 0;
    for (int64_t s = 0; s < n;) {
        int64_t t = s;
        double ap[3] = {0, 0, 0}, an[3] = {0, 0, 0}, ac[3] = {0, 0, 0};
        while (t < n && e[t].k[0] == e[s].k[0] && e[t].k[1] == e[s].k[1] && e[t].k[2] == e[s].k[2]) {
            const int64_t i = e[t].i;
            for (int c = 0; c < 3; ++c) {
                ap[c] += pos[3 * i + c];
                if (nrm) an[c] += nrm[3 * i + c];
                if (col) ac[c] += col[3 * i + c];
            }
            ++t;
        }
        const double cnt = (double)(t - s);
        for (int c = 0; c < 3; ++c) {
            pos_out[3 * m + c] = (float)(ap[c] / cnt);
            if (nrm && nrm_out) nrm_out[3 * m + c] = (float)(an[c] / cnt);
            if (col && col_out) col_out[3 * m + c] = (float)(ac[c] / cnt);
            if (keys_out) keys_out[3 * m + c] = e[s].k[c];
        }
        ++m;
        s = t;
    }
    free(e);
    return m;
}

/* ------------------------------------------------- GetInformationMatrix */

/* kernel::ComputeInformationMatrix (kernel/Registration.cpp:406-436, RegistrationCPU.cpp:655-735) with
 * GetInformationJacobians (RegistrationImpl.h:686-715): GTG = sum over matched target points of
 * Jx Jx^T + Jy Jy^T + Jz Jz^T, the 21 lower-triangle terms evaluated in f32 exactly as upstream writes them
 * (J_x[j] * J_x[k] + J_y[j] * J_y[k] + J_z[j] * J_z[k]) and accumulated here in f64 (upstream: f32 partial sums). */
void orc_information_matrix_f32(const float* tgt, const int64_t* corr, int64_t n, double info36[36]) {
    double sum[21] = {0};
#pragma omp parallel
    {
        double loc[21] = {0};
#pragma omp for schedule(static) nowait
        for (int64_t w = 0; w < n; ++w) {
            if (corr[w] == -1) continue;
            const float* p = tgt + 3 * corr[w];
            const float Jx[6] = {0.f, p[2], -p[1], 1.f, 0.f, 0.f};
            const float Jy[6] = {-p[2], 0.f, p[0], 0.f, 1.f, 0.f};
            const float Jz[6] = {p[1], -p[0], 0.f, 0.f, 0.f, 1.f};
            int i = 0;
            for (int j = 0; j < 6; ++j)
                for (int k = 0; k <= j; ++k) loc[i++] += (double)(Jx[j] * Jx[k] + Jy[j] * Jy[k] + Jz[j] * Jz[k]);
        }
#pragma omp critical
        for (int i = 0; i < 21; ++i) sum[i] += loc[i];
    }
    int i = 0;
    for (int j = 0; j < 6; ++j)
        for (int k = 0; k <= j; ++k) info36[j * 6 + k] = info36[k * 6 + j] = sum[i++];
}

/* ------------------------------------------------------------ ColoredICP */

void orc_solve_sym3x3_pinv(const double Ain[9], const double b[3], double x[3]) {
    double A[3][3], V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A[i][j] = 0.5 * (Ain[3 * i + j] + Ain[3 * j + i]);
    for (int sweep = 0; sweep < 64; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(A[p][q]) < 1e-300) continue;
                const double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < 3; ++k) {
                    const double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - sn * akq;
                    A[k][q] = sn * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    const double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - sn * aqk;
                    A[q][k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - sn * vkq;
                    V[k][q] = sn * vkp + c * vkq;
                }
            }
    }
    x[0] = x[1] = x[2] = 0.0;
    for (int i = 0; i < 3; ++i) {
        const double lam = A[i][i];
        if (fabs(lam) < 1e-10) continue; /* SVD3x3.h:2184-2187 */
        const double proj = (V[0][i] * b[0] + V[1][i] * b[1] + V[2][i] * b[2]) / lam;
        for (int k = 0; k < 3; ++k) x[k] += V[k][i] * proj;
    }
}

/* PointCloudImpl.h:1066-1165 for one point; idx: its neighbour list (count valid entries). */
static void color_gradient_point(const float* pts, const float* nrm, const float* col, int64_t i,
                                 const int32_t* idx, int count, int solver, float* out) {
    const int64_t o = 3 * i;
    if (count < 4) {
        out[o] = out[o + 1] = out[o + 2] = 0;
        return;
    }
    const float vt[3] = {pts[o], pts[o + 1], pts[o + 2]};
    const float nt[3] = {nrm[o], nrm[o + 1], nrm[o + 2]};
    const float it = (float)((col[o] + col[o + 1] + col[o + 2]) / 3.0);
    float AtA[9] = {0}, Atb[3] = {0};
    const float s = vt[0] * nt[0] + vt[1] * nt[1] + vt[2] * nt[2];
    int k = 1;
    for (; k < count; ++k) {
        const int64_t a = 3 * (int64_t)idx[k];
        if (a == -1) break; /* as written upstream (:1106-1108) */
        const float va[3] = {pts[a], pts[a + 1], pts[a + 2]};
        const float d = va[0] * nt[0] + va[1] * nt[1] + va[2] * nt[2] - s;
        const float vp[3] = {va[0] - d * nt[0], va[1] - d * nt[1], va[2] - d * nt[2]};
        const float ia = (float)((col[a] + col[a + 1] + col[a + 2]) / 3.0);
        const float A[3] = {vp[0] - vt[0], vp[1] - vt[1], vp[2] - vt[2]};
        AtA[0] += A[0] * A[0];
        AtA[1] += A[1] * A[0];
        AtA[2] += A[2] * A[0];
        AtA[4] += A[1] * A[1];
        AtA[5] += A[2] * A[1];
        AtA[8] += A[2] * A[2];
        const float b = ia - it;
        Atb[0] += A[0] * b;
        Atb[1] += A[1] * b;
        Atb[2] += A[2] * b;
    }
    const float A[3] = {(k - 1) * nt[0], (k - 1) * nt[1], (k - 1) * nt[2]};
    AtA[0] += A[0] * A[0];
    AtA[1] += A[0] * A[1];
    AtA[2] += A[0] * A[2];
    AtA[4] += A[1] * A[1];
    AtA[5] += A[1] * A[2];
    AtA[8] += A[2] * A[2];
    AtA[3] = AtA[1];
    AtA[6] = AtA[2];
    AtA[7] = AtA[5];
    if (solver == ORC_GRADIENT_SOLVER_REFERENCE) { /* PointCloudImpl.h:1163: solve_svd3x3 (svd3_oracle.c) */
        orc_solve_svd3x3_f32(AtA, Atb, out + o);
        return;
    }
    double Ad[9], bd[3], xd[3]; /* option: exact pseudo-inverse of the same f32 system */
    for (int q = 0; q < 9; ++q) Ad[q] = AtA[q];
    for (int q = 0; q < 3; ++q) bd[q] = Atb[q];
    orc_solve_sym3x3_pinv(Ad, bd, xd);
    out[o] = (float)xd[0];
    out[o + 1] = (float)xd[1];
    out[o + 2] = (float)xd[2];
}

void orc_estimate_color_gradients_f32(const float* pts, const float* nrm, const float* col, int64_t n,
                                      double radius, int max_nn, float* out) {
    orc_estimate_color_gradients_solver_f32(pts, nrm, col, n, radius, max_nn, ORC_GRADIENT_SOLVER_REFERENCE, out);
}

void orc_estimate_color_gradients_solver_f32(const float* pts, const float* nrm, const float* col, int64_t n,
                                             double radius, int max_nn, int solver, float* out) {
    if (max_nn > ORC_MAX_KNN) max_nn = ORC_MAX_KNN;
    int32_t* idx = (int32_t*)malloc((size_t)n * max_nn * sizeof(int32_t));
    float* d2 = (float*)malloc((size_t)n * max_nn * sizeof(float));
    int32_t* cnt = (int32_t*)malloc((size_t)n * sizeof(int32_t));
    if (!idx || !d2 || !cnt) abort();
    orc_hybrid_search_f32(pts, n, pts, n, radius, max_nn, idx, d2, cnt);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) color_gradient_point(pts, nrm, col, i, idx + i * max_nn, cnt[i], solver, out);
    free(idx);
    free(d2);
    free(cnt);
}

int orc_icp_colored_f32(const float* source, const float* source_colors, int64_t n, const float* target,
                        const float* nrm, const float* tcol, const float* tgrad, int64_t m,
                        double max_corr_dist, const double init_T[16], int max_iteration,
                        double rel_fitness, double rel_rmse, double lambda_geometric, int method,
                        double scale, double shape, orc_icp_result* res, double* per_iter,
                        int64_t* corr_out) {
    float* src = (float*)malloc((size_t)(n > 0 ? n : 1) * 3 * sizeof(float));
    int64_t* corr = (int64_t*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int64_t));
    if (!src || !corr) return -1;
    memcpy(src, source, (size_t)n * 3 * sizeof(float));
    double T[16];
    memcpy(T, init_T, sizeof(T));
    orc_transform_points_f32(T, src, n);
    orc_grid g;
    const double t_build0 = now_s();
    if (grid_build(&g, target, m, max_corr_dist) != 0) return -1;
    const double t_loop0 = now_s();
    double fitness = 0, rmse = 0, prev_fitness = 0, prev_rmse = 0;
    int64_t cnt = 0;
    int converged = 0, it = 0;
    for (it = 0; it < max_iteration; ++it) {
        compute_registration_result(&g, src, n, max_corr_dist, corr, &fitness, &rmse, &cnt);
        if (cnt == 0) eye4(T);
        if (fitness <= 2.2250738585072014e-308) {
            converged = 0;
            break;
        }
        double s64[29], pose[6], U[16];
        orc_pose_colored_sums_f32(src, source_colors, target, nrm, tcol, tgrad, corr, n, lambda_geometric,
                                  method, scale, shape, s64, NULL);
        if (orc_decode_and_solve_6x6(s64, pose, NULL, NULL) != 0) {
            grid_free(&g);
            free(src);
            free(corr);
            return 1;
        }
        orc_pose_to_transformation(pose, U);
        matmul4(U, T, T);
        orc_transform_points_f32(U, src, n);
        if (per_iter) {
            per_iter[2 * it + 0] = fitness;
            per_iter[2 * it + 1] = rmse;
        }
        if (it != 0 && fabs(prev_fitness - fitness) < rel_fitness && fabs(prev_rmse - rmse) < rel_rmse) {
            converged = 1;
            break;
        }
        prev_fitness = fitness;
        prev_rmse = rmse;
    }
    res->loop_seconds = now_s() - t_loop0;
    res->build_seconds = t_loop0 - t_build0;
    const int iterations = it;
    compute_registration_result(&g, src, n, max_corr_dist, corr, &fitness, &rmse, &cnt);
    if (cnt == 0) {
        eye4(T);
        converged = 0;
    }
    res->num_iterations = iterations;
    res->converged = converged;
    res->fitness = fitness;
    res->inlier_rmse = rmse;
    memcpy(res->transformation, T, sizeof(T));
    if (corr_out) memcpy(corr_out, corr, (size_t)n * sizeof(int64_t));
    grid_free(&g);
    free(src);
    free(corr);
    return 0;
}
