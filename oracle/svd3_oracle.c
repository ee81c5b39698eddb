/* TEST INFRASTRUCTURE — CPU restatement of the reference's Float32 3x3 SVD least-squares solver.
 *
 * core/linalg/kernel/SVD3x3.h:1131-2168 (svd3x3<float>) and :2170-2215 (solve_svd3x3): the fast SVD of
 * McAdams, Selle, Tamstorf, Teran, Sifakis (UW-Madison TR1690, 2011): A^T A, 4 fixed cyclic Jacobi sweeps
 * with the approximate Givens half-angle accumulated in a quaternion, V from the quaternion, B = A V,
 * columns sorted by norm, three Givens rotations for the QR of B (U, singular values on the diagonal),
 * then x = V diag(1/s_i for |s_i| >= 1e-10) U^T b.
 *
 * The reference is written as ~1000 lines of unrolled scalar statements; here the three conjugations, the
 * three column swaps and the three Givens rotations are one routine each, indexed by axis — the f32
 * operation sequence per scalar is the reference's (host build: __frsqrt_rn(x) = float(1.0 / double(sqrtf(x))),
 * SVD3x3.h:58-72; the double literals 1e-20 and 4 gamma^2 promote their expressions to double).  Pinned
 * bit for bit to the reference's own solve_svd3x3<float> compiled in oracle/_ref
 * (tests/test_oracle_vs_ref.py::test_svd3_solver_bit_exact_vs_reference).
 * Compiled with -ffp-contract=off (oracle/Makefile). */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "oracle.h"

typedef struct {
    float a[3][3], v[3][3], u[3][3];
    float s[3][3]; /* symmetric, [i][j] with i >= j live */
    float qs, qv[3];
} svd3_t;

static inline float* sym(svd3_t* st, int i, int j) { return i >= j ? &st->s[i][j] : &st->s[j][i]; }

static inline float rsqrt_ref(float x) { return (float)(1.0 / (double)sqrtf(x)); }

static inline float rsqrt_refined(float x) {
    const float r = rsqrt_ref(x);
    const float h = r * 0.5f;
    float t = r * h;
    t = r * t;
    t = x * t;
    return (r + h) - t;
}

static inline float bits(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

/* SVD3x3.h:1205-1311 (x,y,z = 0,1,2), :1317-1416 (1,2,0), :1422-1514 (2,0,1) */
static void jacobi_conjugate(svd3_t* st, int X, int Y, int Z) {
    float *sxx = sym(st, X, X), *syy = sym(st, Y, Y), *szz = sym(st, Z, Z);
    float *syx = sym(st, Y, X), *szx = sym(st, Z, X), *szy = sym(st, Z, Y);
    float sh = *syx * 0.5f;
    float t5 = *sxx - *syy;
    const int big = (double)(sh * sh) >= 1.e-20;
    sh = big ? sh : 0.f;
    float ch = big ? t5 : 1.f;
    float t1 = sh * sh, t2 = ch * ch;
    const float r = rsqrt_ref(t1 + t2);
    sh = r * sh;
    ch = r * ch;
    t1 = (float)(5.8284273147583007813 * (double)t1);
    if (t2 <= t1) {
        sh = bits(1053028117u); /* sin(pi/8) */
        ch = bits(1064076127u); /* cos(pi/8) */
    }
    t1 = sh * sh;
    t2 = ch * ch;
    const float c = t2 - t1;
    float s = ch * sh;
    s = s + s;
    const float nrm = t1 + t2;
    *szz = *szz * nrm;
    *szx = *szx * nrm;
    *szy = *szy * nrm;
    *szz = *szz * nrm;
    t1 = s * *szx;
    t2 = s * *szy;
    *szx = t2 + c * *szx;
    *szy = c * *szy - t1;
    const float ss = s * s, cc = c * c;
    t1 = *syy * ss;
    const float t3 = *sxx * ss;
    *sxx = *sxx * cc + t1;
    *syy = *syy * cc + t3;
    const float two_yx = *syx + *syx;
    *syx = *syx * (cc - ss);
    const float cs = c * s;
    t2 = two_yx * cs;
    t5 = t5 * cs;
    *sxx = *sxx + t2;
    *syx = *syx - t5;
    *syy = *syy - t2;
    t1 = sh * st->qv[X];
    t2 = sh * st->qv[Y];
    const float tz = sh * st->qv[Z];
    sh = sh * st->qs;
    st->qs = ch * st->qs;
    for (int k = 0; k < 3; ++k) st->qv[k] = ch * st->qv[k];
    st->qv[Z] = st->qv[Z] + sh;
    st->qs = st->qs - tz;
    st->qv[X] = st->qv[X] + t2;
    st->qv[Y] = st->qv[Y] - t1;
}

/* SVD3x3.h:1655-1810 */
static void sort_columns(svd3_t* st, float n2[3], int P, int Q, int NEG) {
    const int sw = n2[P] < n2[Q];
    if (sw) {
        for (int i = 0; i < 3; ++i) {
            float t = st->a[i][P];
            st->a[i][P] = st->a[i][Q];
            st->a[i][Q] = t;
            t = st->v[i][P];
            st->v[i][P] = st->v[i][Q];
            st->v[i][Q] = t;
        }
        const float t = n2[P];
        n2[P] = n2[Q];
        n2[Q] = t;
    }
    const float f = 1.f + (sw ? -2.f : 0.f);
    for (int i = 0; i < 3; ++i) {
        st->a[i][NEG] = st->a[i][NEG] * f;
        st->v[i][NEG] = st->v[i][NEG] * f;
    }
}

/* SVD3x3.h:1830-1930, 1934-2033, 2037-2138 */
static void qr_givens(svd3_t* st, int P, int Q) {
    const float small = 1.e-12f;
    const float app = st->a[P][P], aqp = st->a[Q][P];
    float sh = (aqp * aqp >= small) ? aqp : 0.f;
    float ch = 0.f - app;
    ch = ch < app ? app : ch;
    ch = ch < small ? small : ch;
    const int pos = app >= 0.f;
    float n = ch * ch + sh * sh;
    ch = ch + rsqrt_refined(n) * n;
    if (!pos) {
        const float t = ch;
        ch = sh;
        sh = t;
    }
    n = ch * ch + sh * sh;
    const float r = rsqrt_refined(n);
    ch = ch * r;
    sh = sh * r;
    const float c = ch * ch - sh * sh;
    float s = sh * ch;
    s = s + s;
    for (int j = 0; j < 3; ++j) {
        const float t1 = s * st->a[P][j], t2 = s * st->a[Q][j];
        st->a[P][j] = c * st->a[P][j] + t2;
        st->a[Q][j] = c * st->a[Q][j] - t1;
    }
    for (int i = 0; i < 3; ++i) {
        const float t1 = s * st->u[i][P], t2 = s * st->u[i][Q];
        st->u[i][P] = c * st->u[i][P] + t2;
        st->u[i][Q] = c * st->u[i][Q] - t1;
    }
}

void orc_svd3x3_f32(const float A[9], float U[9], float S[3], float V[9]) {
    svd3_t st;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) st.a[i][j] = A[3 * i + j];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j <= i; ++j)
            st.s[i][j] = st.a[2][i] * st.a[2][j] + (st.a[1][i] * st.a[1][j] + st.a[0][i] * st.a[0][j]);
    st.qs = 1.f;
    st.qv[0] = st.qv[1] = st.qv[2] = 0.f;
    for (int sweep = 0; sweep < 4; ++sweep) {
        jacobi_conjugate(&st, 0, 1, 2);
        jacobi_conjugate(&st, 1, 2, 0);
        jacobi_conjugate(&st, 2, 0, 1);
    }
    float n = st.qs * st.qs;
    for (int k = 0; k < 3; ++k) n = st.qv[k] * st.qv[k] + n;
    const float r = rsqrt_refined(n);
    const float w = st.qs * r, x = st.qv[0] * r, y = st.qv[1] * r, z = st.qv[2] * r;
    const float xx = x * x, yy = y * y, zz = z * z, ww = w * w;
    const float d = ww - xx;
    st.v[2][2] = (d - yy) + zz;
    st.v[1][1] = (d + yy) - zz;
    st.v[0][0] = ((ww + xx) - yy) - zz;
    const float x2 = x + x, y2 = y + y, z2 = z + z;
    const float wx = w * x2, wy = w * y2, wz = w * z2;
    const float xy = y * x2, yz = z * y2, zx = x * z2;
    st.v[0][1] = xy - wz;
    st.v[1][2] = yz - wx;
    st.v[2][0] = zx - wy;
    st.v[1][0] = xy + wz;
    st.v[2][1] = yz + wx;
    st.v[0][2] = zx + wy;
    for (int i = 0; i < 3; ++i) {
        const float a0 = st.a[i][0], a1 = st.a[i][1], a2 = st.a[i][2];
        for (int j = 0; j < 3; ++j) st.a[i][j] = (st.v[0][j] * a0 + st.v[1][j] * a1) + st.v[2][j] * a2;
    }
    float n2[3];
    for (int j = 0; j < 3; ++j) n2[j] = (st.a[0][j] * st.a[0][j] + st.a[1][j] * st.a[1][j]) + st.a[2][j] * st.a[2][j];
    sort_columns(&st, n2, 0, 1, 1);
    sort_columns(&st, n2, 0, 2, 0);
    sort_columns(&st, n2, 1, 2, 2);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) st.u[i][j] = i == j ? 1.f : 0.f;
    qr_givens(&st, 0, 1);
    qr_givens(&st, 0, 2);
    qr_givens(&st, 1, 2);
    for (int i = 0; i < 3; ++i) {
        S[i] = st.a[i][i];
        for (int j = 0; j < 3; ++j) {
            U[3 * i + j] = st.u[i][j];
            V[3 * i + j] = st.v[i][j];
        }
    }
}

/* SVD3x3.h:2170-2215; products through core/linalg/kernel/Matrix.h:33-59 */
void orc_solve_svd3x3_f32(const float A[9], const float b[3], float x[3]) {
    float U[9], S[3], V[9], sut[9], ainv[9];
    orc_svd3x3_f32(A, U, S, V);
    for (int i = 0; i < 3; ++i) {
        const float inv = fabsf(S[i]) < 1e-10f ? 0.f : (float)(1.0 / (double)S[i]);
        for (int j = 0; j < 3; ++j) sut[3 * i + j] = U[3 * j + i] * inv;
    }
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            ainv[3 * i + j] = (V[3 * i] * sut[j] + V[3 * i + 1] * sut[3 + j]) + V[3 * i + 2] * sut[6 + j];
    for (int i = 0; i < 3; ++i) x[i] = (ainv[3 * i] * b[0] + ainv[3 * i + 1] * b[1]) + ainv[3 * i + 2] * b[2];
}
