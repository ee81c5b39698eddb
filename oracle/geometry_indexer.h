/*
 * geometry_indexer.h — t/geometry/kernel/GeometryIndexer.h:25-144 TransformIndexer and
 * ArrayIndexer::InBoundary restated for the oracle (TEST INFRASTRUCTURE, see oracle.h).
 * Shared by tsdf_oracle.c and odometry_oracle.c; pinned bit-exactly against the reference's
 * own header through oracle/ref_shim (tests/test_oracle_vs_ref.py).
 */
#ifndef ORC_GEOMETRY_INDEXER_H
#define ORC_GEOMETRY_INDEXER_H

/* everything is stored as float (:47-59) */
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

#endif
