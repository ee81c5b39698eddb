/*
 * tsdf_oracle.c — CPU restatement of Open3D's VoxelBlockGrid TSDF hot path
 * (frustum block discovery, set-semantics activate, per-voxel integration).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Never imported by the product.
 * Compile with -ffp-contract=off: all f32 expressions are evaluated operation
 * by operation in the reference's source order, so block keys and pixel
 * selections (floor / truncation of f32 values) are reproducible bit for bit.
 *
 * Reference paths are relative to cpp/open3d/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

/* t/geometry/Utility.h:77-115 */
void orc_inverse_transformation(const double T[16], double Ti[16]) {
    Ti[0] = T[0];
    Ti[1] = T[4];
    Ti[2] = T[8];
    Ti[4] = T[1];
    Ti[5] = T[5];
    Ti[6] = T[9];
    Ti[8] = T[2];
    Ti[9] = T[6];
    Ti[10] = T[10];
    Ti[3] = -(Ti[0] * T[3] + Ti[1] * T[7] + Ti[2] * T[11]);
    Ti[7] = -(Ti[4] * T[3] + Ti[5] * T[7] + Ti[6] * T[11]);
    Ti[11] = -(Ti[8] * T[3] + Ti[9] * T[7] + Ti[10] * T[11]);
    Ti[12] = 0;
    Ti[13] = 0;
    Ti[14] = 0;
    Ti[15] = 1;
}

/* t/geometry/kernel/GeometryIndexer.h:25-144 TransformIndexer: everything is
 * stored as float (:47-59). */
typedef struct {
    float e[3][4];
    float fx, fy, cx, cy;
    float scale;
} xform_indexer;

static void xi_init(xform_indexer* t, const double K[9], const double E[16],
                    float scale) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) t->e[i][j] = (float)E[i * 4 + j];
    t->fx = (float)K[0];
    t->fy = (float)K[4];
    t->cx = (float)K[2];
    t->cy = (float)K[5];
    t->scale = scale;
}

/* :62-78 */
static inline void xi_rigid(const xform_indexer* t, float x, float y, float z,
                            float* xo, float* yo, float* zo) {
    x *= t->scale;
    y *= t->scale;
    z *= t->scale;
    *xo = x * t->e[0][0] + y * t->e[0][1] + z * t->e[0][2] + t->e[0][3];
    *yo = x * t->e[1][0] + y * t->e[1][1] + z * t->e[1][2] + t->e[1][3];
    *zo = x * t->e[2][0] + y * t->e[2][1] + z * t->e[2][2] + t->e[2][3];
}

/* :100-108 */
static inline void xi_project(const xform_indexer* t, float x, float y,
                              float z, float* u, float* v) {
    float inv_z = 1.0f / z;
    *u = t->fx * x * inv_z + t->cx;
    *v = t->fy * y * inv_z + t->cy;
}

/* :111-120 */
static inline void xi_unproject(const xform_indexer* t, float u, float v,
                                float d, float* x, float* y, float* z) {
    *x = (u - t->cx) * d / t->fx;
    *y = (v - t->cy) * d / t->fy;
    *z = d;
}

/* :294-297 ArrayIndexer::InBoundary(x, y), shape = (rows, cols) */
static inline int in_boundary(float x, float y, int rows, int cols) {
    return y >= 0 && x >= 0 && y <= rows - 1.0f && x <= cols - 1.0f;
}

static int cmp_key3(const void* a, const void* b) {
    const int32_t* x = (const int32_t*)a;
    const int32_t* y = (const int32_t*)b;
    for (int i = 0; i < 3; ++i) {
        if (x[i] < y[i]) return -1;
        if (x[i] > y[i]) return 1;
    }
    return 0;
}

/* t/geometry/kernel/VoxelBlockGridCPU.cpp:117-201 */
int64_t orc_depth_touch(const void* depth, int is_f32, int rows, int cols,
                        const double K[9], const double extrinsic[16],
                        int resolution, float voxel_size, float sdf_trunc,
                        float depth_scale, float depth_max, int stride,
                        int32_t* keys_out, int64_t max_keys) {
    double pose[16];
    orc_inverse_transformation(extrinsic, pose); /* :131 */
    xform_indexer ti;
    xi_init(&ti, K, pose, 1.0f); /* :132 */

    const int rows_s = rows / stride;
    const int cols_s = cols / stride;
    const int64_t n = (int64_t)rows_s * cols_s;
    const float block_size = voxel_size * resolution; /* :140 */

    int32_t* cand = (int32_t*)malloc((size_t)(n > 0 ? n : 1) * 4 * 3 * sizeof(int32_t));
    uint8_t* used = (uint8_t*)calloc((size_t)(n > 0 ? n : 1), 1);
    if (!cand || !used) return -2;

#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < n; ++w) {
        const int y = (int)(w / cols_s) * stride;
        const int x = (int)(w % cols_s) * stride;
        float d;
        if (is_f32)
            d = ((const float*)depth)[(int64_t)y * cols + x] / depth_scale;
        else
            d = ((const uint16_t*)depth)[(int64_t)y * cols + x] / depth_scale;
        if (d > 0 && d < depth_max) {
            float xc, yc, zc, xg, yg, zg;
            xi_unproject(&ti, (float)x, (float)y, 1.0f, &xc, &yc, &zc);
            xi_rigid(&ti, xc, yc, zc, &xg, &yg, &zg);
            const float xo = ti.e[0][3], yo = ti.e[1][3], zo = ti.e[2][3];
            const float xd = xg - xo, yd = yg - yo, zd = zg - zo;
            const int step_size = 3;
            const float t_min = fmaxf(d - sdf_trunc, 0.0f);
            const float t_max = fminf(d + sdf_trunc, depth_max);
            const float t_step = (t_max - t_min) / step_size;
            float t = t_min;
            for (int step = 0; step <= step_size; ++step) {
                int32_t* k = cand + (w * 4 + step) * 3;
                k[0] = (int32_t)floorf((xo + t * xd) / block_size);
                k[1] = (int32_t)floorf((yo + t * yd) / block_size);
                k[2] = (int32_t)floorf((zo + t * zd) / block_size);
                t += t_step;
            }
            used[w] = 1;
        }
    }
    int64_t m = 0;
    for (int64_t w = 0; w < n; ++w) {
        if (!used[w]) continue;
        if (m != w) memmove(cand + m * 12, cand + w * 12, 12 * sizeof(int32_t));
        ++m;
    }
    m *= 4;
    qsort(cand, (size_t)m, 3 * sizeof(int32_t), cmp_key3);
    int64_t u = 0;
    for (int64_t i = 0; i < m; ++i) {
        if (i == 0 || cmp_key3(cand + 3 * i, cand + 3 * (i - 1)) != 0) {
            if (u >= max_keys) {
                free(cand);
                free(used);
                return -1;
            }
            memcpy(keys_out + 3 * u, cand + 3 * i, 3 * sizeof(int32_t));
            ++u;
        }
    }
    free(cand);
    free(used);
    return u;
}

/* t/geometry/kernel/VoxelBlockGridImpl.h:151-308 */
void orc_tsdf_integrate(const void* depth, const void* color, int inputs_f32,
                        int rows, int cols, const int32_t* buf_indices,
                        int64_t n_blocks, const int32_t* block_keys,
                        float* tsdf_buf, uint16_t* weight_buf,
                        uint16_t* color_buf, const double depth_K[9],
                        const double color_K[9], const double extrinsic[16],
                        int resolution, float voxel_size, float sdf_trunc,
                        float depth_scale, float depth_max) {
    const int res = resolution;
    const int res2 = res * res;
    const int res3 = res2 * res;
    xform_indexer ti, ci;
    xi_init(&ti, depth_K, extrinsic, voxel_size); /* :184 */
    const double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
    xi_init(&ci, color_K ? color_K : depth_K, eye, 1.0f); /* :185-187 */
    const int integrate_color = color != NULL && color_buf != NULL;
    const float color_multiplier = (integrate_color && inputs_f32) ? 255.0f : 1.0f;

    const int64_t n = n_blocks * res3;
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < n; ++w) {
        const int block_idx = buf_indices[w / res3];
        const int voxel_idx = (int)(w % res3);
        const int32_t* key = block_keys + 3 * (int64_t)block_idx;
        const int xb = key[0], yb = key[1], zb = key[2];
        /* GeometryIndexer.h:270-278 WorkloadToCoord (3D) */
        int rem = voxel_idx;
        const int xv = rem % res;
        rem = (rem - xv) / res;
        const int yv = rem % res;
        const int zv = rem / res;
        const int x = xb * res + xv;
        const int y = yb * res + yv;
        const int z = zb * res + zv;
        float xc, yc, zc, u, v;
        xi_rigid(&ti, (float)x, (float)y, (float)z, &xc, &yc, &zc);
        xi_project(&ti, xc, yc, zc, &u, &v);
        if (!in_boundary(u, v, rows, cols)) continue;
        int ui = (int)u;
        int vi = (int)v;
        float dep;
        if (inputs_f32)
            dep = ((const float*)depth)[(int64_t)vi * cols + ui] / depth_scale;
        else
            dep = ((const uint16_t*)depth)[(int64_t)vi * cols + ui] / depth_scale;
        float sdf = dep - zc;
        if (dep <= 0 || dep > depth_max || zc <= 0 || sdf < -sdf_trunc) continue;
        sdf = sdf < sdf_trunc ? sdf : sdf_trunc;
        sdf /= sdf_trunc;

        const int64_t lin = (int64_t)block_idx * res3 + voxel_idx;
        float* tsdf_ptr = tsdf_buf + lin;
        uint16_t* weight_ptr = weight_buf + lin;
        float inv_wsum = 1.0f / (*weight_ptr + 1);
        float weight = *weight_ptr;
        *tsdf_ptr = (weight * (*tsdf_ptr) + sdf) * inv_wsum;

        if (integrate_color) {
            uint16_t* color_ptr = color_buf + 3 * lin;
            float px, py, pz, uf, vf;
            xi_unproject(&ti, (float)ui, (float)vi, 1.0f, &px, &py, &pz);
            xi_project(&ci, px, py, pz, &uf, &vf);
            if (in_boundary(uf, vf, rows, cols)) {
                ui = (int)roundf(uf);
                vi = (int)roundf(vf);
                const int64_t off = ((int64_t)vi * cols + ui) * 3;
                for (int i = 0; i < 3; ++i) {
                    float in = inputs_f32 ? ((const float*)color)[off + i]
                                          : (float)((const uint8_t*)color)[off + i];
                    color_ptr[i] = (uint16_t)((weight * color_ptr[i] +
                                               in * color_multiplier) *
                                              inv_wsum);
                }
            }
        }
        *weight_ptr = (uint16_t)(weight + 1);
    }
}

/* core/hashmap/HashMap.cpp:166-197 + CPU/TBBHashBackend.h:173-227: set insert.
 * A throw-away chained hash over the key buffer, rebuilt per call. */
int orc_hashmap_activate(int32_t* table_keys, int64_t capacity, int64_t* size,
                         const int32_t* keys, int64_t n, int32_t* buf_indices,
                         uint8_t* masks) {
    int64_t nb = 16;
    while (nb < 2 * (*size + n)) nb <<= 1;
    int64_t* head = (int64_t*)malloc((size_t)nb * sizeof(int64_t));
    int64_t* next = (int64_t*)malloc((size_t)(capacity > 0 ? capacity : 1) * sizeof(int64_t));
    if (!head || !next) return -2;
    for (int64_t i = 0; i < nb; ++i) head[i] = -1;
    for (int64_t s = 0; s < *size; ++s) {
        const int32_t* k = table_keys + 3 * s;
        uint64_t h = orc_minivec_hash_i32x3(k[0], k[1], k[2]) & (uint64_t)(nb - 1);
        next[s] = head[h];
        head[h] = s;
    }
    int rc = 0;
    for (int64_t i = 0; i < n; ++i) {
        const int32_t* k = keys + 3 * i;
        uint64_t h = orc_minivec_hash_i32x3(k[0], k[1], k[2]) & (uint64_t)(nb - 1);
        int64_t s = head[h];
        while (s >= 0) {
            const int32_t* t = table_keys + 3 * s;
            if (t[0] == k[0] && t[1] == k[1] && t[2] == k[2]) break;
            s = next[s];
        }
        if (s >= 0) {
            buf_indices[i] = (int32_t)s;
            masks[i] = 0;
        } else {
            if (*size >= capacity) {
                rc = -1;
                buf_indices[i] = -1;
                masks[i] = 0;
                continue;
            }
            s = (*size)++;
            memcpy(table_keys + 3 * s, k, 3 * sizeof(int32_t));
            next[s] = head[h];
            head[h] = s;
            buf_indices[i] = (int32_t)s;
            masks[i] = 1;
        }
    }
    free(head);
    free(next);
    return rc;
}
