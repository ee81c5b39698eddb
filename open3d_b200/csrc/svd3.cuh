// 3x3 least-squares solve with the reference's Float32 solver semantics.
//
// PointCloud::EstimateColorGradients (t/geometry/kernel/PointCloudImpl.h:1068-1165) ends in
// core::linalg::kernel::solve_svd3x3 (core/linalg/kernel/SVD3x3.h:2170-2215): the fast SVD of
// McAdams, Selle, Tamstorf, Teran, Sifakis, "Computing the Singular Value Decomposition of 3x3
// matrices with minimal branching and elementary floating point operations" (UW-Madison TR1690,
// 2011) — 4 fixed cyclic Jacobi sweeps on A^T A with the approximate Givens angle (quaternion
// accumulation), column sort, Givens QR — followed by x = V diag(1/s_i, |s_i| >= 1e-10) U^T b.
// With 4 sweeps and f32 arithmetic the decomposition is only approximate for the ill-conditioned
// systems this path produces (condition ~1e5), so "what the reference computes" is defined by its
// exact operation sequence, not by the mathematical solution.  This file restates that sequence
// (indexed by cyclic axis triples instead of the reference's unrolled scalars) with every float
// operation pinned to round-to-nearest and no FMA contraction, and with the reference's HOST
// semantics for the mixed-precision spots (SVD3x3.h:58-72: rsqrt(x) = float(1.0 / double(sqrtf(x))),
// the double constants 1e-20 and 4*gamma^2 promote their expressions to double), so the result is
// bit-identical to the reference's CPU path, which is what the oracle is pinned to
// (tests/test_oracle_vs_ref.py::test_svd3_solver_*).
#pragma once
#include <cuda_runtime.h>

namespace o3db {
namespace svd3 {

#define S3_MUL(a, b) __fmul_rn((a), (b))
#define S3_ADD(a, b) __fadd_rn((a), (b))
#define S3_SUB(a, b) __fsub_rn((a), (b))

__device__ __forceinline__ float rsqrt_ref(float x) {   // SVD3x3.h:64 (host definition of __frsqrt_rn)
    return (float)__ddiv_rn(1.0, (double)__fsqrt_rn(x));
}
// one Newton step on r = rsqrt(x):  r + r/2 - x r^3 / 2, in the reference's association (SVD3x3.h:1531-1536)
__device__ __forceinline__ float rsqrt_refined(float x) {
    const float r = rsqrt_ref(x);
    const float h = S3_MUL(r, 0.5f);
    float t = S3_MUL(r, h);
    t = S3_MUL(r, t);
    t = S3_MUL(x, t);
    return S3_SUB(S3_ADD(r, h), t);
}

struct State {
    float a[3][3];   // A, later B = A V, later R of the QR
    float v[3][3];
    float u[3][3];
    float s[3][3];   // symmetric A^T A; only s[i][j], i >= j, is live
    float qs, qv[3];
};

__device__ __forceinline__ float& sym(State& st, int i, int j) { return i >= j ? st.s[i][j] : st.s[j][i]; }

// One Jacobi conjugation on the (X, Y) plane, Z the remaining axis; (X,Y,Z) runs through the cyclic
// triples (0,1,2), (1,2,0), (2,0,1) (SVD3x3.h:1205-1311, 1317-1416, 1422-1514 are the three instances).
template <int X, int Y, int Z>
__device__ __forceinline__ void jacobi_conjugate(State& st) {
    const float kSinPi8 = __uint_as_float(1053028117u), kCosPi8 = __uint_as_float(1064076127u);
    float& sxx = sym(st, X, X);
    float& syy = sym(st, Y, Y);
    float& szz = sym(st, Z, Z);
    float& syx = sym(st, Y, X);
    float& szx = sym(st, Z, X);
    float& szy = sym(st, Z, Y);
    // approximate Givens half-angle (ch, sh)
    float sh = S3_MUL(syx, 0.5f);
    float t5 = S3_SUB(sxx, syy);
    const bool big = (double)S3_MUL(sh, sh) >= 1.e-20;
    sh = big ? sh : 0.f;
    float ch = big ? t5 : 1.f;
    float t1 = S3_MUL(sh, sh), t2 = S3_MUL(ch, ch);
    const float r = rsqrt_ref(S3_ADD(t1, t2));
    sh = S3_MUL(r, sh);
    ch = S3_MUL(r, ch);
    t1 = (float)__dmul_rn(5.8284273147583007813, (double)t1);
    const bool clamp = t2 <= t1;
    sh = clamp ? kSinPi8 : sh;
    ch = clamp ? kCosPi8 : ch;
    t1 = S3_MUL(sh, sh);
    t2 = S3_MUL(ch, ch);
    const float c = S3_SUB(t2, t1);
    float s = S3_MUL(ch, sh);
    s = S3_ADD(s, s);
    // conjugation S <- Q^T S Q (unnormalised: the Z row carries (sh^2 + ch^2))
    const float nrm = S3_ADD(t1, t2);
    szz = S3_MUL(szz, nrm);
    szx = S3_MUL(szx, nrm);
    szy = S3_MUL(szy, nrm);
    szz = S3_MUL(szz, nrm);
    t1 = S3_MUL(s, szx);
    t2 = S3_MUL(s, szy);
    szx = S3_ADD(t2, S3_MUL(c, szx));
    szy = S3_SUB(S3_MUL(c, szy), t1);
    const float ss = S3_MUL(s, s), cc = S3_MUL(c, c);
    t1 = S3_MUL(syy, ss);
    const float t3 = S3_MUL(sxx, ss);
    sxx = S3_ADD(S3_MUL(sxx, cc), t1);
    syy = S3_ADD(S3_MUL(syy, cc), t3);
    const float two_yx = S3_ADD(syx, syx);
    syx = S3_MUL(syx, S3_SUB(cc, ss));
    const float cs = S3_MUL(c, s);
    t2 = S3_MUL(two_yx, cs);
    t5 = S3_MUL(t5, cs);
    sxx = S3_ADD(sxx, t2);
    syx = S3_SUB(syx, t5);
    syy = S3_SUB(syy, t2);
    // cumulative rotation as a quaternion
    t1 = S3_MUL(sh, st.qv[X]);
    t2 = S3_MUL(sh, st.qv[Y]);
    const float tz = S3_MUL(sh, st.qv[Z]);
    sh = S3_MUL(sh, st.qs);
    st.qs = S3_MUL(ch, st.qs);
    st.qv[0] = S3_MUL(ch, st.qv[0]);
    st.qv[1] = S3_MUL(ch, st.qv[1]);
    st.qv[2] = S3_MUL(ch, st.qv[2]);
    st.qv[Z] = S3_ADD(st.qv[Z], sh);
    st.qs = S3_SUB(st.qs, tz);
    st.qv[X] = S3_ADD(st.qv[X], t2);
    st.qv[Y] = S3_SUB(st.qv[Y], t1);
}

// Swap columns P and Q of B and V when |b_P|^2 < |b_Q|^2 and negate column NEG (keeps det V = +1)
// (SVD3x3.h:1655-1810).
template <int P, int Q, int NEG>
__device__ __forceinline__ void sort_columns(State& st, float n2[3]) {
    const bool sw = n2[P] < n2[Q];
    if (sw) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float t = st.a[i][P];
            st.a[i][P] = st.a[i][Q];
            st.a[i][Q] = t;
            t = st.v[i][P];
            st.v[i][P] = st.v[i][Q];
            st.v[i][Q] = t;
        }
        const float t = n2[P];
        n2[P] = n2[Q];
        n2[Q] = t;
    }
    const float f = S3_ADD(1.f, sw ? -2.f : 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        st.a[i][NEG] = S3_MUL(st.a[i][NEG], f);
        st.v[i][NEG] = S3_MUL(st.v[i][NEG], f);
    }
}

// Givens rotation of rows (P, Q) zeroing b_QP, applied to B and accumulated into U
// (SVD3x3.h:1830-1930, 1934-2033, 2037-2138).
template <int P, int Q>
__device__ __forceinline__ void qr_givens(State& st) {
    const float kSmall = 1.e-12f;
    const float app = st.a[P][P], aqp = st.a[Q][P];
    float sh = S3_MUL(aqp, aqp) >= kSmall ? aqp : 0.f;
    float ch = S3_SUB(0.f, app);
    ch = ch < app ? app : ch;         // std::max(ch, app)
    ch = ch < kSmall ? kSmall : ch;   // std::max(ch, small)
    const bool pos = app >= 0.f;
    float n = S3_ADD(S3_MUL(ch, ch), S3_MUL(sh, sh));
    ch = S3_ADD(ch, S3_MUL(rsqrt_refined(n), n));
    if (!pos) {
        const float t = ch;
        ch = sh;
        sh = t;
    }
    n = S3_ADD(S3_MUL(ch, ch), S3_MUL(sh, sh));
    const float r = rsqrt_refined(n);
    ch = S3_MUL(ch, r);
    sh = S3_MUL(sh, r);
    const float c = S3_SUB(S3_MUL(ch, ch), S3_MUL(sh, sh));
    float s = S3_MUL(sh, ch);
    s = S3_ADD(s, s);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float t1 = S3_MUL(s, st.a[P][j]), t2 = S3_MUL(s, st.a[Q][j]);
        st.a[P][j] = S3_ADD(S3_MUL(c, st.a[P][j]), t2);
        st.a[Q][j] = S3_SUB(S3_MUL(c, st.a[Q][j]), t1);
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float t1 = S3_MUL(s, st.u[i][P]), t2 = S3_MUL(s, st.u[i][Q]);
        st.u[i][P] = S3_ADD(S3_MUL(c, st.u[i][P]), t2);
        st.u[i][Q] = S3_SUB(S3_MUL(c, st.u[i][Q]), t1);
    }
}

// A = U diag(S) V^T, row-major 3x3 (SVD3x3.h:1131-2168).
__device__ __forceinline__ void decompose(const float A[9], State& st) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) st.a[i][j] = A[3 * i + j];
    // normal equations, lower triangle: s_ij = sum_k a_ki a_kj, k ascending
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j <= i; ++j)
            st.s[i][j] = S3_ADD(S3_MUL(st.a[2][i], st.a[2][j]),
                                S3_ADD(S3_MUL(st.a[1][i], st.a[1][j]), S3_MUL(st.a[0][i], st.a[0][j])));
    st.qs = 1.f;
    st.qv[0] = st.qv[1] = st.qv[2] = 0.f;
#pragma unroll 1
    for (int sweep = 0; sweep < 4; ++sweep) {
        jacobi_conjugate<0, 1, 2>(st);
        jacobi_conjugate<1, 2, 0>(st);
        jacobi_conjugate<2, 0, 1>(st);
    }
    // normalise the quaternion (one Newton step on the reciprocal square root), SVD3x3.h:1521-1540
    float n = S3_MUL(st.qs, st.qs);
    n = S3_ADD(S3_MUL(st.qv[0], st.qv[0]), n);
    n = S3_ADD(S3_MUL(st.qv[1], st.qv[1]), n);
    n = S3_ADD(S3_MUL(st.qv[2], st.qv[2]), n);
    const float r = rsqrt_refined(n);
    const float w = S3_MUL(st.qs, r), x = S3_MUL(st.qv[0], r), y = S3_MUL(st.qv[1], r), z = S3_MUL(st.qv[2], r);
    // quaternion -> V, SVD3x3.h:1546-1572
    const float xx = S3_MUL(x, x), yy = S3_MUL(y, y), zz = S3_MUL(z, z), ww = S3_MUL(w, w);
    const float d = S3_SUB(ww, xx);
    st.v[2][2] = S3_ADD(S3_SUB(d, yy), zz);
    st.v[1][1] = S3_SUB(S3_ADD(d, yy), zz);
    st.v[0][0] = S3_SUB(S3_SUB(S3_ADD(ww, xx), yy), zz);
    const float x2 = S3_ADD(x, x), y2 = S3_ADD(y, y), z2 = S3_ADD(z, z);
    const float wx = S3_MUL(w, x2), wy = S3_MUL(w, y2), wz = S3_MUL(w, z2);
    const float xy = S3_MUL(y, x2), yz = S3_MUL(z, y2), zx = S3_MUL(x, z2);
    st.v[0][1] = S3_SUB(xy, wz);
    st.v[1][2] = S3_SUB(yz, wx);
    st.v[2][0] = S3_SUB(zx, wy);
    st.v[1][0] = S3_ADD(xy, wz);
    st.v[2][1] = S3_ADD(yz, wx);
    st.v[0][2] = S3_ADD(zx, wy);
    // B = A V, SVD3x3.h:1578-1630
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float a0 = st.a[i][0], a1 = st.a[i][1], a2 = st.a[i][2];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            st.a[i][j] = S3_ADD(S3_ADD(S3_MUL(st.v[0][j], a0), S3_MUL(st.v[1][j], a1)), S3_MUL(st.v[2][j], a2));
    }
    // sort the columns by norm, SVD3x3.h:1636-1814
    float n2[3];
#pragma unroll
    for (int j = 0; j < 3; ++j)
        n2[j] = S3_ADD(S3_ADD(S3_MUL(st.a[0][j], st.a[0][j]), S3_MUL(st.a[1][j], st.a[1][j])), S3_MUL(st.a[2][j], st.a[2][j]));
    sort_columns<0, 1, 1>(st, n2);
    sort_columns<0, 2, 0>(st, n2);
    sort_columns<1, 2, 2>(st, n2);
    // QR of B by three Givens rotations, U accumulated from the identity
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) st.u[i][j] = i == j ? 1.f : 0.f;
    qr_givens<0, 1>(st);
    qr_givens<0, 2>(st);
    qr_givens<1, 2>(st);
}

// x = V Sigma^+ U^T b, singular values below 1e-10 dropped (SVD3x3.h:2170-2215; products
// through core/linalg/kernel/Matrix.h:33-59: left-to-right sums).
__device__ __forceinline__ void solve(const float A[9], const float b[3], float x[3]) {
    State st;
    decompose(A, st);
    float sut[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float sv = st.a[i][i];
        const float inv = fabsf(sv) < 1e-10f ? 0.f : (float)__ddiv_rn(1.0, (double)sv);
#pragma unroll
        for (int j = 0; j < 3; ++j) sut[i][j] = S3_MUL(st.u[j][i], inv);
    }
    float ainv[3][3];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            ainv[i][j] = S3_ADD(S3_ADD(S3_MUL(st.v[i][0], sut[0][j]), S3_MUL(st.v[i][1], sut[1][j])), S3_MUL(st.v[i][2], sut[2][j]));
#pragma unroll
    for (int i = 0; i < 3; ++i)
        x[i] = S3_ADD(S3_ADD(S3_MUL(ainv[i][0], b[0]), S3_MUL(ainv[i][1], b[1])), S3_MUL(ainv[i][2], b[2]));
}

#undef S3_MUL
#undef S3_ADD
#undef S3_SUB

}  // namespace svd3
}  // namespace o3db
