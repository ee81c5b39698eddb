// grid.cuh — dense uniform search grid over the target cloud and the exact
// 1-NN-within-radius scan used by both the stand-alone hybrid search and the
// fused ICP iteration kernel.
//
// Replaces the reference's hashed CSR table (core/nns/FixedRadiusSearchImpl.cuh:
// 63-134 build, :514-631 query).  Results (index, dist^2, count) are identical
// to an exhaustive search; the table layout is private to this library.
//
// Layout in HBM (built once per target / scale):
//   pts4      float4[M]   target points sorted by cell key, .w = original index bits
//   nrm4      float4[M]   target normals in the same order (optional)
//   cell_start u32[ncell+1] CSR offsets, key = (iz*ny + iy)*nx + ix (x fastest, so
//              every (iy,iz) row of cells is ONE contiguous run of pts4)
#pragma once

#include "common.cuh"

namespace o3db {

struct Grid {
    float ox, oy, oz;   // origin = bbox min
    float inv_c, c;     // cell size (>= search radius) and reciprocal
    float tol;          // pruning slack covering binning round-off
    int nx, ny, nz;
    float bmin[3], bmax[3];
};

// Monotone binning: x <= y  =>  cell1(x) <= cell1(y)  (sub, mul by a positive
// constant, floor and clamp are all monotone under round-to-nearest).  Coverage
// of the radius search relies only on this property, never on exact cell bounds.
__device__ __forceinline__ int cell1(float x, float o, float inv_c, int n) {
    int i = __float2int_rd((x - o) * inv_c);
    return min(max(i, 0), n - 1);
}

__device__ __forceinline__ unsigned cell_key(const Grid& g, float x, float y, float z) {
    const int ix = cell1(x, g.ox, g.inv_c, g.nx);
    const int iy = cell1(y, g.oy, g.inv_c, g.ny);
    const int iz = cell1(z, g.oz, g.inv_c, g.nz);
    return (unsigned)((iz * g.ny + iy) * g.nx + ix);
}

// Lower / upper end of the interval that must be binned to cover |t - q| <= r
// given that dist^2 was accepted in f32: r' = r (1 + 1e-6) plus 2 ulp of |q|.
__device__ __forceinline__ float lo_bound(float q, float rr) { return (q - rr) - fabsf(q) * 2.4e-7f; }
__device__ __forceinline__ float hi_bound(float q, float rr) { return (q + rr) + fabsf(q) * 2.4e-7f; }

struct Best {
    float d;    // best dist^2 so far (starts at the threshold: accepts d <= thr)
    int j;      // position in the sorted arrays, -1 = none
    int idx;    // original index of the best point (tie-break: lower wins)
    float x, y, z;
};

__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, unsigned s, unsigned e,
                                           float qx, float qy, float qz, Best& b) {
    for (unsigned j = s; j < e; ++j) {
        const float4 t = __ldg(&pts[j]);
        const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
        // canonical arithmetic (bit-identical to oracle/icp_oracle.c dist2_f32)
        const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
        const int idx = __float_as_int(t.w);
        if (d < b.d || (d == b.d && idx < b.idx)) {
            b.d = d;
            b.j = (int)j;
            b.idx = idx;
            b.x = t.x;
            b.y = t.y;
            b.z = t.z;
        }
    }
}

// Exact nearest neighbour of (qx,qy,qz) among points with dist^2 <= thr.
// rr = radius * (1 + 1e-6).
template <bool PRUNE>
__device__ __forceinline__ void nn_search(const Grid& g, const float4* __restrict__ pts,
                                          const unsigned* __restrict__ cs, float qx, float qy,
                                          float qz, float rr, float thr, Best& b) {
    b.d = thr;
    b.j = -1;
    b.idx = 0x7fffffff;
    b.x = b.y = b.z = 0.f;
    const float lx = lo_bound(qx, rr), hx = hi_bound(qx, rr);
    const float ly = lo_bound(qy, rr), hy = hi_bound(qy, rr);
    const float lz = lo_bound(qz, rr), hz = hi_bound(qz, rr);
    // entirely outside the (slightly inflated) bounding box: no candidate can pass
    if (hx < g.bmin[0] || lx > g.bmax[0] || hy < g.bmin[1] || ly > g.bmax[1] || hz < g.bmin[2] ||
        lz > g.bmax[2] || !(qx == qx) || !(qy == qy) || !(qz == qz))
        return;
    const int x0 = cell1(lx, g.ox, g.inv_c, g.nx), x1 = cell1(hx, g.ox, g.inv_c, g.nx);
    const int y0 = cell1(ly, g.oy, g.inv_c, g.ny), y1 = cell1(hy, g.oy, g.inv_c, g.ny);
    const int z0 = cell1(lz, g.oz, g.inv_c, g.nz), z1 = cell1(hz, g.oz, g.inv_c, g.nz);
    const int cy = cell1(qy, g.oy, g.inv_c, g.ny), cz = cell1(qz, g.oz, g.inv_c, g.nz);
    {   // the query's own row first: it almost always holds the winner, which
        // then prunes most of the other rows
        const int row = (cz * g.ny + cy) * g.nx;
        scan_range(pts, cs[row + x0], cs[row + x1 + 1], qx, qy, qz, b);
    }
    for (int iz = z0; iz <= z1; ++iz) {
        float gz = 0.f;
        if (PRUNE) {
            if (iz > cz) gz = (g.oz + (float)iz * g.c) - qz - g.tol;
            else if (iz < cz) gz = qz - (g.oz + (float)(iz + 1) * g.c) - g.tol;
            gz = fmaxf(gz, 0.f);
            gz *= gz;
        }
        for (int iy = y0; iy <= y1; ++iy) {
            if (iy == cy && iz == cz) continue;
            if (PRUNE) {
                float gy = 0.f;
                if (iy > cy) gy = (g.oy + (float)iy * g.c) - qy - g.tol;
                else if (iy < cy) gy = qy - (g.oy + (float)(iy + 1) * g.c) - g.tol;
                gy = fmaxf(gy, 0.f);
                if (fmaf(gy, gy, gz) > b.d) continue;  // strict: keeps exact ties reachable
            }
            const int row = (iz * g.ny + iy) * g.nx;
            scan_range(pts, cs[row + x0], cs[row + x1 + 1], qx, qy, qz, b);
        }
    }
}

}  // namespace o3db
