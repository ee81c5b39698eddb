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

#include "geometry_indexer.h"

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
/* VoxelBlockGridImpl.h:151-308 for one value layout (the reference instantiates weight_t / color_t as
 * uint16_t / uint16_t and float / float, VoxelBlockGridCPU.cpp / VoxelBlockGridCUDA.cu:238-244). */
#define ORC_DEFINE_TSDF_INTEGRATE(NAME, WEIGHT_T, COLOR_T) \
static void NAME(const void* depth, const void* color, int inputs_f32, \
                        int rows, int cols, const int32_t* buf_indices, \
                        int64_t n_blocks, const int32_t* block_keys, \
                        float* tsdf_buf, WEIGHT_T* weight_buf, \
                        COLOR_T* color_buf, const double depth_K[9], \
                        const double color_K[9], const double extrinsic[16], \
                        int resolution, float voxel_size, float sdf_trunc, \
                        float depth_scale, float depth_max) { \
    const int res = resolution; \
    const int res2 = res * res; \
    const int res3 = res2 * res; \
    xform_indexer ti, ci; \
    xi_init(&ti, depth_K, extrinsic, voxel_size); /* :184 */ \
    const double eye[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1}; \
    xi_init(&ci, color_K ? color_K : depth_K, eye, 1.0f); /* :185-187 */ \
    const int integrate_color = color != NULL && color_buf != NULL; \
    const float color_multiplier = (integrate_color && inputs_f32) ? 255.0f : 1.0f; \
    const int64_t n = n_blocks * res3; \
    _Pragma("omp parallel for schedule(static)") \
    for (int64_t w = 0; w < n; ++w) { \
        const int block_idx = buf_indices[w / res3]; \
        const int voxel_idx = (int)(w % res3); \
        const int32_t* key = block_keys + 3 * (int64_t)block_idx; \
        const int xb = key[0], yb = key[1], zb = key[2]; \
        /* GeometryIndexer.h:270-278 WorkloadToCoord (3D) */ \
        int rem = voxel_idx; \
        const int xv = rem % res; \
        rem = (rem - xv) / res; \
        const int yv = rem % res; \
        const int zv = rem / res; \
        const int x = xb * res + xv; \
        const int y = yb * res + yv; \
        const int z = zb * res + zv; \
        float xc, yc, zc, u, v; \
        xi_rigid(&ti, (float)x, (float)y, (float)z, &xc, &yc, &zc); \
        xi_project(&ti, xc, yc, zc, &u, &v); \
        if (!in_boundary(u, v, rows, cols)) continue; \
        int ui = (int)u; \
        int vi = (int)v; \
        float dep; \
        if (inputs_f32) \
            dep = ((const float*)depth)[(int64_t)vi * cols + ui] / depth_scale; \
        else \
            dep = ((const uint16_t*)depth)[(int64_t)vi * cols + ui] / depth_scale; \
        float sdf = dep - zc; \
        if (dep <= 0 || dep > depth_max || zc <= 0 || sdf < -sdf_trunc) continue; \
        sdf = sdf < sdf_trunc ? sdf : sdf_trunc; \
        sdf /= sdf_trunc; \
        const int64_t lin = (int64_t)block_idx * res3 + voxel_idx; \
        float* tsdf_ptr = tsdf_buf + lin; \
        WEIGHT_T* weight_ptr = weight_buf + lin; \
        float inv_wsum = 1.0f / (*weight_ptr + 1); \
        float weight = *weight_ptr; \
        *tsdf_ptr = (weight * (*tsdf_ptr) + sdf) * inv_wsum; \
        if (integrate_color) { \
            COLOR_T* color_ptr = color_buf + 3 * lin; \
            float px, py, pz, uf, vf; \
            xi_unproject(&ti, (float)ui, (float)vi, 1.0f, &px, &py, &pz); \
            xi_project(&ci, px, py, pz, &uf, &vf); \
            if (in_boundary(uf, vf, rows, cols)) { \
                ui = (int)roundf(uf); \
                vi = (int)roundf(vf); \
                const int64_t off = ((int64_t)vi * cols + ui) * 3; \
                for (int i = 0; i < 3; ++i) { \
                    float in = inputs_f32 ? ((const float*)color)[off + i] \
                                          : (float)((const uint8_t*)color)[off + i]; \
                    color_ptr[i] = (COLOR_T)((weight * color_ptr[i] + \
                                               in * color_multiplier) * \
                                              inv_wsum); \
                } \
            } \
        } \
        *weight_ptr = (WEIGHT_T)(weight + 1); \
    } \
}

ORC_DEFINE_TSDF_INTEGRATE(tsdf_integrate_u16, uint16_t, uint16_t)
ORC_DEFINE_TSDF_INTEGRATE(tsdf_integrate_f32, float, float)

void orc_tsdf_integrate(const void* depth, const void* color, int inputs_f32,
                        int rows, int cols, const int32_t* buf_indices,
                        int64_t n_blocks, const int32_t* block_keys,
                        float* tsdf_buf, uint16_t* weight_buf,
                        uint16_t* color_buf, const double depth_K[9],
                        const double color_K[9], const double extrinsic[16],
                        int resolution, float voxel_size, float sdf_trunc,
                        float depth_scale, float depth_max) {
    tsdf_integrate_u16(depth, color, inputs_f32, rows, cols, buf_indices, n_blocks, block_keys, tsdf_buf, weight_buf,
                       color_buf, depth_K, color_K, extrinsic, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
}

void orc_tsdf_integrate_f32_values(const void* depth, const void* color, int inputs_f32,
                                   int rows, int cols, const int32_t* buf_indices,
                                   int64_t n_blocks, const int32_t* block_keys,
                                   float* tsdf_buf, float* weight_buf,
                                   float* color_buf, const double depth_K[9],
                                   const double color_K[9], const double extrinsic[16],
                                   int resolution, float voxel_size, float sdf_trunc,
                                   float depth_scale, float depth_max) {
    tsdf_integrate_f32(depth, color, inputs_f32, rows, cols, buf_indices, n_blocks, block_keys, tsdf_buf, weight_buf,
                       color_buf, depth_K, color_K, extrinsic, resolution, voxel_size, sdf_trunc, depth_scale, depth_max);
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

/* ------------------------------------------------- EstimateRange + RayCast */

/* GeometryIndexer.h:81-97 Rotate */
static inline void xi_rotate(const xform_indexer* t, float x, float y, float z,
                             float* xo, float* yo, float* zo) {
    x *= t->scale;
    y *= t->scale;
    z *= t->scale;
    *xo = x * t->e[0][0] + y * t->e[0][1] + z * t->e[0][2];
    *yo = x * t->e[1][0] + y * t->e[1][1] + z * t->e[1][2];
    *zo = x * t->e[2][0] + y * t->e[2][1] + z * t->e[2][2];
}

/* float -> int conversion of an out-of-range value is undefined in C++; the reference's CUDA build
 * saturates (cvt.rzi.s32.f32), x86 yields INT_MIN.  A block corner just in front of the camera plane
 * projects to |u|, |v| ~ 1e9+, so the choice matters for EstimateRange: the CUDA behaviour is restated. */
static inline int sat_int(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return 2147483647;
    if (x <= -2147483648.0f) return (-2147483647 - 1);
    return (int)x;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* t/geometry/kernel/VoxelBlockGridImpl.h:310-555 EstimateRangeCPU.  The 16x16 "fragments" of
 * passes 0/1 only partition each block's screen rectangle; the min/max they scatter equals the
 * min/max over the whole rectangle, which is what is computed here.  The reference's fragment
 * buffer can overflow (:431-434, :458-468: a warning and a partial map); that failure mode is
 * not restated — the map is always complete. */
void orc_estimate_range(const int32_t* block_keys, int64_t n, const double K[9],
                        const double E[16], int h, int w, int down_factor,
                        int resolution, float voxel_size, float depth_min,
                        float depth_max, float* range) {
    const int h_down = h / down_factor, w_down = w / down_factor;
    xform_indexer ti;
    xi_init(&ti, K, E, 1.0f);
    for (int64_t i = 0; i < (int64_t)h_down * w_down; ++i) { /* pass 0.5 (:472-484) */
        range[2 * i] = depth_max;
        range[2 * i + 1] = depth_min;
    }
    const int64_t block_resolution = resolution;
    for (int64_t b = 0; b < n; ++b) {
        const int32_t* key = block_keys + 3 * b;
        int u_min = w_down - 1, v_min = h_down - 1, u_max = 0, v_max = 0;
        float z_min = depth_max, z_max = depth_min;
        for (int i = 0; i < 8; ++i) { /* :389-412 */
            float xw = (key[0] + ((i & 1) > 0)) * block_resolution * voxel_size;
            float yw = (key[1] + ((i & 2) > 0)) * block_resolution * voxel_size;
            float zw = (key[2] + ((i & 4) > 0)) * block_resolution * voxel_size;
            float xc, yc, zc, u, v;
            xi_rigid(&ti, xw, yw, zw, &xc, &yc, &zc);
            if (zc <= 0) continue;
            xi_project(&ti, xc, yc, zc, &u, &v);
            u /= down_factor;
            v /= down_factor;
            v_min = imin(sat_int(floorf(v)), v_min);
            v_max = imax(sat_int(ceilf(v)), v_max);
            u_min = imin(sat_int(floorf(u)), u_min);
            u_max = imax(sat_int(ceilf(u)), u_max);
            z_min = z_min < zc ? z_min : zc;
            z_max = z_max > zc ? z_max : zc;
        }
        v_min = imax(0, v_min);
        v_max = imin(h_down - 1, v_max);
        u_min = imax(0, u_min);
        u_max = imin(w_down - 1, u_max);
        if (v_min >= v_max || u_min >= u_max || z_min >= z_max) continue; /* :420 */
        for (int v = v_min; v <= v_max; ++v)
            for (int u = u_min; u <= u_max; ++u) { /* pass 1 (:497-541) */
                float* r = range + 2 * ((int64_t)v * w_down + u);
                r[0] = z_min < r[0] ? z_min : r[0];
                r[1] = z_max > r[1] ? z_max : r[1];
            }
    }
}

typedef struct {
    const int32_t* keys;
    int64_t* head;
    int64_t* next;
    int64_t nb;
} block_lookup;

static int lookup_build(block_lookup* m, const int32_t* keys, int64_t size) {
    m->keys = keys;
    m->nb = 16;
    while (m->nb < 2 * size) m->nb <<= 1;
    m->head = (int64_t*)malloc((size_t)m->nb * sizeof(int64_t));
    m->next = (int64_t*)malloc((size_t)(size > 0 ? size : 1) * sizeof(int64_t));
    if (!m->head || !m->next) return -1;
    for (int64_t i = 0; i < m->nb; ++i) m->head[i] = -1;
    for (int64_t s = 0; s < size; ++s) {
        const int32_t* k = keys + 3 * s;
        uint64_t hsh = orc_minivec_hash_i32x3(k[0], k[1], k[2]) & (uint64_t)(m->nb - 1);
        m->next[s] = m->head[hsh];
        m->head[hsh] = s;
    }
    return 0;
}

static inline int64_t lookup_find(const block_lookup* m, int x, int y, int z) {
    uint64_t hsh = orc_minivec_hash_i32x3(x, y, z) & (uint64_t)(m->nb - 1);
    for (int64_t s = m->head[hsh]; s >= 0; s = m->next[s]) {
        const int32_t* k = m->keys + 3 * s;
        if (k[0] == x && k[1] == y && k[2] == z) return s;
    }
    return -1;
}

static inline int isign(int x) { return (x > 0) ? 1 : ((x < 0) ? -1 : 0); } /* GeometryMacros.h:92-94 */

/* t/geometry/kernel/VoxelBlockGridImpl.h:578-1120 RayCastCPU for <float, uint16_t, uint16_t>.
 * The 1-entry MiniVecCache (:557-576) only short-cuts hash lookups and is omitted.
 * Bit-exact against RayCastCPU<float,u16,u16> compiled from the reference (tests/test_oracle_vs_ref_vbg.py).
 * One deliberate deviation: the voxel coordinate inside the block, index_t((x_g - x_b*block_size)
 * / voxel_size) (:826-828), can round up to `resolution` when x_g sits a rounding error below a
 * block face; upstream then indexes the next row of the block (or past the buffer).  It is
 * clamped to resolution-1 here and in the CUDA kernel. */
void orc_ray_cast(const int32_t* table_keys, int64_t size, const float* tsdf_buf,
                  const uint16_t* weight_buf, const uint16_t* color_buf,
                  const float* range, const double K[9], const double E[16], int h,
                  int w, int resolution, float voxel_size, float depth_scale,
                  float depth_min, float depth_max, float weight_threshold,
                  float trunc_voxel_multiplier, int range_map_down_factor,
                  float* depth_out, float* vertex_out, float* color_out,
                  float* normal_out, int64_t* index_out, uint8_t* mask_out,
                  float* ratio_out, float* ratio_dx_out, float* ratio_dy_out,
                  float* ratio_dz_out) {
    (void)depth_min;
    (void)depth_max;
    block_lookup map;
    if (lookup_build(&map, table_keys, size) != 0) abort();
    double Einv[16];
    orc_inverse_transformation(E, Einv);
    xform_indexer c2w, w2c;
    xi_init(&c2w, K, Einv, 1.0f);
    xi_init(&w2c, K, E, 1.0f);
    const int block_resolution = resolution;
    const float block_size = voxel_size * block_resolution;
    const int resolution2 = block_resolution * block_resolution;
    const int resolution3 = resolution2 * block_resolution;
    const int w_down = w / range_map_down_factor, h_down = h / range_map_down_factor;
    const int render_color = color_buf != NULL && color_out != NULL;
    const int visit_neighbors = render_color || normal_out || mask_out || index_out || ratio_out ||
                                ratio_dx_out || ratio_dy_out || ratio_dz_out;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t workload = 0; workload < (int64_t)h * w; ++workload) {
        const int y = (int)(workload / w), x = (int)(workload % w);
        /* :851-852 range_indexer(x / down, y / down): when the image size is not a multiple of the down
         * factor the last partial row / column of cells does not exist (h_down = h / down) and upstream
         * reads past the map; clamped to the last cell here and in the CUDA kernel. */
        const float* rng = range + 2 * ((int64_t)imin(y / range_map_down_factor, h_down - 1) * w_down +
                                        imin(x / range_map_down_factor, w_down - 1));
        float* depth_ptr = depth_out ? depth_out + workload : NULL;
        float* vertex_ptr = vertex_out ? vertex_out + 3 * workload : NULL;
        float* color_ptr = render_color ? color_out + 3 * workload : NULL;
        float* normal_ptr = normal_out ? normal_out + 3 * workload : NULL;
        int64_t* index_ptr = index_out ? index_out + 8 * workload : NULL;
        uint8_t* mask_ptr = mask_out ? mask_out + 8 * workload : NULL;
        float* ratio_ptr = ratio_out ? ratio_out + 8 * workload : NULL;
        float* rdx_ptr = ratio_dx_out ? ratio_dx_out + 8 * workload : NULL;
        float* rdy_ptr = ratio_dy_out ? ratio_dy_out + 8 * workload : NULL;
        float* rdz_ptr = ratio_dz_out ? ratio_dz_out + 8 * workload : NULL;
        if (vertex_ptr) vertex_ptr[0] = vertex_ptr[1] = vertex_ptr[2] = 0;
        if (depth_ptr) depth_ptr[0] = 0;
        if (normal_ptr) normal_ptr[0] = normal_ptr[1] = normal_ptr[2] = 0;
        if (color_out) color_out[3 * workload] = color_out[3 * workload + 1] = color_out[3 * workload + 2] = 0;
        for (int i = 0; i < 8; ++i) {
            if (mask_ptr) mask_ptr[i] = 0;
            if (index_ptr) index_ptr[i] = 0;
            if (ratio_ptr) ratio_ptr[i] = 0;
            if (rdx_ptr) rdx_ptr[i] = 0;
            if (rdy_ptr) rdy_ptr[i] = 0;
            if (rdz_ptr) rdz_ptr[i] = 0;
        }
        float t = rng[0];
        const float t_max = rng[1];
        if (t >= t_max) continue;

        float x_c = 0, y_c = 0, z_c = 0, x_g = 0, y_g = 0, z_g = 0, x_o = 0, y_o = 0, z_o = 0;
        float t_prev = t;
        float tsdf_prev = -1.0f;
        float tsdf = 1.0;
        const float sdf_trunc = voxel_size * trunc_voxel_multiplier;
        float wgt = 0.0;
        xi_rigid(&c2w, 0, 0, 0, &x_o, &y_o, &z_o);
        xi_unproject(&c2w, (float)x, (float)y, 1.0f, &x_c, &y_c, &z_c);
        xi_rigid(&c2w, x_c, y_c, z_c, &x_g, &y_g, &z_g);
        const float x_d = (x_g - x_o), y_d = (y_g - y_o), z_d = (z_g - z_o);

        int surface_found = 0;
        while (t < t_max) {
            /* GetLinearIdxAtT (:795-833) */
            int64_t linear_idx = -1;
            {
                const float xg = x_o + t * x_d, yg = y_o + t * y_d, zg = z_o + t * z_d;
                const int x_b = (int)floorf(xg / block_size);
                const int y_b = (int)floorf(yg / block_size);
                const int z_b = (int)floorf(zg / block_size);
                const int64_t blk = lookup_find(&map, x_b, y_b, z_b);
                if (blk >= 0) {
                    int x_v = (int)((xg - x_b * block_size) / voxel_size);
                    int y_v = (int)((yg - y_b * block_size) / voxel_size);
                    int z_v = (int)((zg - z_b * block_size) / voxel_size);
                    x_v = imin(x_v, block_resolution - 1);
                    y_v = imin(y_v, block_resolution - 1);
                    z_v = imin(z_v, block_resolution - 1);
                    linear_idx = blk * resolution3 + z_v * resolution2 + y_v * block_resolution + x_v;
                }
            }
            if (linear_idx < 0) {
                t_prev = t;
                t += block_size;
            } else {
                tsdf_prev = tsdf;
                tsdf = tsdf_buf[linear_idx];
                wgt = weight_buf[linear_idx];
                if (tsdf_prev > 0 && wgt >= weight_threshold && tsdf <= 0) {
                    surface_found = 1;
                    break;
                }
                t_prev = t;
                float delta = tsdf * sdf_trunc;
                t += delta < voxel_size ? voxel_size : delta;
            }
        }
        if (!surface_found) continue;

        float t_intersect = (t * tsdf_prev - t_prev * tsdf) / (tsdf_prev - tsdf);
        x_g = x_o + t_intersect * x_d;
        y_g = y_o + t_intersect * y_d;
        z_g = z_o + t_intersect * z_d;
        if (depth_ptr) *depth_ptr = t_intersect * depth_scale;
        if (vertex_ptr) xi_rigid(&w2c, x_g, y_g, z_g, vertex_ptr + 0, vertex_ptr + 1, vertex_ptr + 2);
        if (!visit_neighbors) continue;

        const int x_b = (int)floorf(x_g / block_size);
        const int y_b = (int)floorf(y_g / block_size);
        const int z_b = (int)floorf(z_g / block_size);
        const float x_v = (x_g - (float)x_b * block_size) / voxel_size;
        const float y_v = (y_g - (float)y_b * block_size) / voxel_size;
        const float z_v = (z_g - (float)z_b * block_size) / voxel_size;
        const int64_t block_buf_idx = lookup_find(&map, x_b, y_b, z_b);
        if (block_buf_idx < 0) continue;
        const int x_v_floor = (int)floorf(x_v), y_v_floor = (int)floorf(y_v), z_v_floor = (int)floorf(z_v);
        const float ratio_x = x_v - (float)x_v_floor;
        const float ratio_y = y_v - (float)y_v_floor;
        const float ratio_z = z_v - (float)z_v_floor;
        float sum_r = 0.0;
        for (int k = 0; k < 8; ++k) {
            const int dx_v = (k & 1) > 0 ? 1 : 0, dy_v = (k & 2) > 0 ? 1 : 0, dz_v = (k & 4) > 0 ? 1 : 0;
            /* GetLinearIdxAtP (:748-793) */
            int64_t lin;
            {
                const int xv = x_v_floor + dx_v, yv = y_v_floor + dy_v, zv = z_v_floor + dz_v;
                const int x_vn = (xv + block_resolution) % block_resolution;
                const int y_vn = (yv + block_resolution) % block_resolution;
                const int z_vn = (zv + block_resolution) % block_resolution;
                const int dx_b = isign(xv - x_vn), dy_b = isign(yv - y_vn), dz_b = isign(zv - z_vn);
                if (dx_b == 0 && dy_b == 0 && dz_b == 0) {
                    lin = block_buf_idx * resolution3 + zv * resolution2 + yv * block_resolution + xv;
                } else {
                    const int64_t nb = lookup_find(&map, x_b + dx_b, y_b + dy_b, z_b + dz_b);
                    lin = nb < 0 ? -1 : nb * resolution3 + z_vn * resolution2 + y_vn * block_resolution + x_vn;
                }
            }
            if (lin >= 0 && weight_buf[lin] > 0) {
                const float rx = dx_v * (ratio_x) + (1 - dx_v) * (1 - ratio_x);
                const float ry = dy_v * (ratio_y) + (1 - dy_v) * (1 - ratio_y);
                const float rz = dz_v * (ratio_z) + (1 - dz_v) * (1 - ratio_z);
                const float r = rx * ry * rz;
                if (ratio_ptr) ratio_ptr[k] = r;
                if (mask_ptr) mask_ptr[k] = 1;
                if (index_ptr) index_ptr[k] = lin;
                const float tsdf_k = tsdf_buf[lin];
                const float idx_ = ry * rz * (2 * dx_v - 1);
                const float idy_ = rx * rz * (2 * dy_v - 1);
                const float idz_ = rx * ry * (2 * dz_v - 1);
                if (rdx_ptr) rdx_ptr[k] = idx_;
                if (rdy_ptr) rdy_ptr[k] = idy_;
                if (rdz_ptr) rdz_ptr[k] = idz_;
                if (normal_ptr) {
                    normal_ptr[0] += idx_ * tsdf_k;
                    normal_ptr[1] += idy_ * tsdf_k;
                    normal_ptr[2] += idz_ * tsdf_k;
                }
                if (color_ptr) {
                    const int64_t c = lin * 3;
                    color_ptr[0] += r * color_buf[c + 0];
                    color_ptr[1] += r * color_buf[c + 1];
                    color_ptr[2] += r * color_buf[c + 2];
                }
                sum_r += r;
            }
        }
        if (sum_r > 0) {
            sum_r *= 255.0;
            if (color_ptr) {
                color_ptr[0] /= sum_r;
                color_ptr[1] /= sum_r;
                color_ptr[2] /= sum_r;
            }
            if (normal_ptr) {
                const float EPSILON = 1e-5f;
                float norm = sqrtf(normal_ptr[0] * normal_ptr[0] + normal_ptr[1] * normal_ptr[1] +
                                   normal_ptr[2] * normal_ptr[2]);
                norm = norm > EPSILON ? norm : EPSILON;
                xi_rotate(&w2c, -normal_ptr[0] / norm, -normal_ptr[1] / norm, -normal_ptr[2] / norm,
                          normal_ptr + 0, normal_ptr + 1, normal_ptr + 2);
            }
        }
    }
    free(map.head);
    free(map.next);
}
