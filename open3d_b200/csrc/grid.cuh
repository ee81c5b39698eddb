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

// All fields are in GRID axis order: grid-x is the fastest-varying axis of the CSR table.
// ax[k] names the real coordinate (0 = x, 1 = y, 2 = z) that grid axis k follows; the host
// makes the axis with the smallest extent the fastest one, so that for surface-like clouds a
// whole "column" of cells along the thin direction is one short contiguous run.
struct Grid {
    float ox, oy, oz;   // origin = bbox min
    float inv_c, c;     // cell size and reciprocal (grid-y, grid-z)
    float inv_cx, cx;   // cell size along grid-x, the thin axis: kThinFactor x coarser — slab scans cross
                        // all grid-x cells anyway, and the CSR table shrinks by that factor (L2 residency)
    float tol;          // pruning slack covering binning round-off
    int nx, ny, nz;
    float bmin[3], bmax[3];
    int ax[3];
};

__device__ __forceinline__ float pick_axis(int a, float x, float y, float z) { return a == 0 ? x : (a == 1 ? y : z); }
// real (x,y,z) -> grid-ordered (gx,gy,gz)
__device__ __forceinline__ void to_grid(const Grid& g, float x, float y, float z, float& gx, float& gy, float& gz) {
    gx = pick_axis(g.ax[0], x, y, z);
    gy = pick_axis(g.ax[1], x, y, z);
    gz = pick_axis(g.ax[2], x, y, z);
}

// Monotone binning: x <= y  =>  cell1(x) <= cell1(y)  (sub, mul by a positive
// constant, floor and clamp are all monotone under round-to-nearest).  Coverage
// of the radius search relies only on this property, never on exact cell bounds.
__device__ __forceinline__ int cell1(float x, float o, float inv_c, int n) {
    int i = __float2int_rd((x - o) * inv_c);
    return min(max(i, 0), n - 1);
}

__device__ __forceinline__ unsigned cell_key(const Grid& g, float rx, float ry, float rz) {
    float x, y, z;
    to_grid(g, rx, ry, rz, x, y, z);
    const int ix = cell1(x, g.ox, g.inv_cx, g.nx);
    const int iy = cell1(y, g.oy, g.inv_c, g.ny);
    const int iz = cell1(z, g.oz, g.inv_c, g.nz);
    return (unsigned)((iz * g.ny + iy) * g.nx + ix);
}

// Lower / upper end of the interval that must be binned to cover |t - q| <= r
// given that dist^2 was accepted in f32: r' = r (1 + 1e-6) plus 2 ulp of |q|.
__device__ __forceinline__ float lo_bound(float q, float rr) { return (q - rr) - fabsf(q) * 2.4e-7f; }
__device__ __forceinline__ float hi_bound(float q, float rr) { return (q + rr) + fabsf(q) * 2.4e-7f; }

static constexpr unsigned kSlabMax = 32;   // longest slab scanned whole (see nn_search_two_pass)

struct Best {
    float d;    // best dist^2 so far (starts at the threshold: accepts d <= thr)
    int j;      // position in the sorted arrays, -1 = none
    int idx;    // original index of the best point (tie-break: lower wins)
    float x, y, z;
};

// Scans pts[s, e): 4 independent 128-bit loads in flight per step (the tail re-reads the last
// candidate, which cannot change the result).
__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, unsigned s, unsigned e,
                                           float qx, float qy, float qz, Best& b) {
    for (unsigned j = s; j < e; j += 4) {
        float4 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = __ldg(&pts[min(j + k, e - 1)]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dx = t[k].x - qx, dy = t[k].y - qy, dz = t[k].z - qz;
            // canonical arithmetic (bit-identical to oracle/icp_oracle.c dist2_f32)
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const int idx = __float_as_int(t[k].w);
            if (d < b.d || (d == b.d && idx < b.idx)) {
                b.d = d;
                b.j = (int)min(j + k, e - 1);
                b.idx = idx;
                b.x = t[k].x;
                b.y = t[k].y;
                b.z = t[k].z;
            }
        }
    }
}

// One (iy, iz) row of cells: visit only the x cells that can still hold a point at
// distance <= best, given that every point of the row is at least sqrt(gap2) away in (y, z).
// skip_cx >= 0: that cell was scanned already.
__device__ __forceinline__ void scan_row(const Grid& g, const float4* __restrict__ pts,
                                         const unsigned* __restrict__ cs, int row, int x0, int x1, int skip_cx,
                                         float gap2, float gqx, float qx, float qy, float qz, Best& b) {
    // admissible |dx|: dx^2 <= best - gap2 (1e-6 best covers the rounding of the subtraction)
    const float ex = sqrtf(fmaf(b.d, 1e-6f, b.d - gap2)) * 1.00001f;
    const int xa = max(x0, cell1(lo_bound(gqx, ex), g.ox, g.inv_cx, g.nx));
    const int xb = min(x1, cell1(hi_bound(gqx, ex), g.ox, g.inv_cx, g.nx));
    if (skip_cx < xa || skip_cx > xb) {
        if (xa <= xb) scan_range(pts, cs[row + xa], cs[row + xb + 1], qx, qy, qz, b);
    } else {
        if (xa < skip_cx) scan_range(pts, cs[row + xa], cs[row + skip_cx], qx, qy, qz, b);
        if (skip_cx < xb) scan_range(pts, cs[row + skip_cx + 1], cs[row + xb + 1], qx, qy, qz, b);
    }
}

// Exact nearest neighbour of (qx,qy,qz) among points with dist^2 <= thr.
// rr = radius * (1 + 1e-6).  Visiting order: the query's own cell, the rest of its row, then the
// other rows of the neighbourhood; a row (and the x cells inside it) is skipped as soon as its
// distance lower bound exceeds the best distance found so far.  PRUNE = false visits everything
// (used to validate the pruning).
template <bool PRUNE>
__device__ __forceinline__ void nn_search(const Grid& g, const float4* __restrict__ pts,
                                          const unsigned* __restrict__ cs, float qx, float qy,
                                          float qz, float rr, float thr, Best& b) {
    b.d = thr;
    b.j = -1;
    b.idx = 0x7fffffff;
    b.x = b.y = b.z = 0.f;
    float gx, gy, gz;                       // the query in grid axis order (cells / gaps only)
    to_grid(g, qx, qy, qz, gx, gy, gz);
    const float lx = lo_bound(gx, rr), hx = hi_bound(gx, rr);
    const float ly = lo_bound(gy, rr), hy = hi_bound(gy, rr);
    const float lz = lo_bound(gz, rr), hz = hi_bound(gz, rr);
    // entirely outside the bounding box (or NaN): no candidate can pass
    if (hx < g.bmin[0] || lx > g.bmax[0] || hy < g.bmin[1] || ly > g.bmax[1] || hz < g.bmin[2] ||
        lz > g.bmax[2] || !(qx == qx) || !(qy == qy) || !(qz == qz))
        return;
    const int x0 = cell1(lx, g.ox, g.inv_cx, g.nx), x1 = cell1(hx, g.ox, g.inv_cx, g.nx);
    const int y0 = cell1(ly, g.oy, g.inv_c, g.ny), y1 = cell1(hy, g.oy, g.inv_c, g.ny);
    const int z0 = cell1(lz, g.oz, g.inv_c, g.nz), z1 = cell1(hz, g.oz, g.inv_c, g.nz);
    const int cx = cell1(gx, g.ox, g.inv_cx, g.nx);
    const int cy = cell1(gy, g.oy, g.inv_c, g.ny), cz = cell1(gz, g.oz, g.inv_c, g.nz);
    if (!PRUNE) {
        for (int iz = z0; iz <= z1; ++iz)
            for (int iy = y0; iy <= y1; ++iy) {
                const int row = (iz * g.ny + iy) * g.nx;
                scan_range(pts, cs[row + x0], cs[row + x1 + 1], qx, qy, qz, b);
            }
        return;
    }
    {
        const int row = (cz * g.ny + cy) * g.nx;
        scan_range(pts, cs[row + cx], cs[row + cx + 1], qx, qy, qz, b);   // own cell
        scan_row(g, pts, cs, row, x0, x1, cx, 0.f, gx, qx, qy, qz, b);     // rest of the own row
    }
    for (int iz = z0; iz <= z1; ++iz) {
        float dz = 0.f;
        if (iz > cz) dz = (g.oz + (float)iz * g.c) - gz - g.tol;
        else if (iz < cz) dz = gz - (g.oz + (float)(iz + 1) * g.c) - g.tol;
        dz = fmaxf(dz, 0.f);
        dz *= dz;
        if (dz > b.d) continue;
        for (int iy = y0; iy <= y1; ++iy) {
            if (iy == cy && iz == cz) continue;
            float dy = 0.f;
            if (iy > cy) dy = (g.oy + (float)iy * g.c) - gy - g.tol;
            else if (iy < cy) dy = gy - (g.oy + (float)(iy + 1) * g.c) - g.tol;
            dy = fmaxf(dy, 0.f);
            const float gap2 = fmaf(dy, dy, dz);
            if (gap2 > b.d) continue;  // strict: keeps exact ties reachable
            scan_row(g, pts, cs, (iz * g.ny + iy) * g.nx, x0, x1, -1, gap2, gx, qx, qy, qz, b);
        }
    }
}

// Two-pass search for fine grids (cell < radius), the fused ICP kernel's default:
//   pass 1  every point of the 3x3x3-ish box [q - r1, q + r1] (r1 = cell size), no pruning:
//           all lanes run the same short loops (little divergence); if the best point
//           found is within r1 it is the exact nearest neighbour (nothing closer can lie
//           outside the box);
//   pass 2  only for lanes that found nothing within r1: the general pruned search.
// r1_accept2 = (r1 (1 - 1e-4))^2.
__device__ __forceinline__ void nn_search_two_pass(const Grid& g, const float4* __restrict__ pts,
                                                   const unsigned* __restrict__ cs, float qx, float qy, float qz,
                                                   float r1, float r1_accept2, float rr, float thr, Best& b) {
    b.d = thr;
    b.j = -1;
    b.idx = 0x7fffffff;
    b.x = b.y = b.z = 0.f;
    float gx, gy, gz;
    to_grid(g, qx, qy, qz, gx, gy, gz);
    if (hi_bound(gx, rr) < g.bmin[0] || lo_bound(gx, rr) > g.bmax[0] || hi_bound(gy, rr) < g.bmin[1] ||
        lo_bound(gy, rr) > g.bmax[1] || hi_bound(gz, rr) < g.bmin[2] || lo_bound(gz, rr) > g.bmax[2] ||
        !(qx == qx) || !(qy == qy) || !(qz == qz))
        return;
    const int x0 = cell1(lo_bound(gx, r1), g.ox, g.inv_cx, g.nx), x1 = cell1(hi_bound(gx, r1), g.ox, g.inv_cx, g.nx);
    const int y0 = cell1(lo_bound(gy, r1), g.oy, g.inv_c, g.ny), y1 = cell1(hi_bound(gy, r1), g.oy, g.inv_c, g.ny);
    const int z0 = cell1(lo_bound(gz, r1), g.oz, g.inv_c, g.nz), z1 = cell1(hi_bound(gz, r1), g.oz, g.inv_c, g.nz);
    if (z1 - z0 <= 2) {
        // SLABS.  For every grid-z of the box, the cells [all grid-x] x [y0..y1] are ONE contiguous
        // run of the table (x fastest, then y).  Grid-x is the cloud's thin direction, so for
        // surface-like data the run holds just the handful of points of 3 cell columns — the same
        // candidates as nine per-row runs, in a third of the loops and a third of the CSR loads.
        // A slab that turns out long (volumetric data) is scanned row by row over [x0..x1] instead.
        unsigned ss[3], se[3];
#pragma unroll
        for (int dz = 0; dz < 3; ++dz) {
            const bool on = z0 + dz <= z1;
            const int plane = min(z0 + dz, z1) * g.ny;
            const unsigned s = __ldg(&cs[(plane + y0) * g.nx]), e = __ldg(&cs[(plane + y1 + 1) * g.nx]);
            ss[dz] = s;
            se[dz] = on ? e : s;
        }
        const unsigned n0 = se[0] - ss[0], n1 = se[1] - ss[1], n2 = se[2] - ss[2];
        if (n0 <= kSlabMax && n1 <= kSlabMax && n2 <= kSlabMax) {
            // ONE loop over the concatenation of the three slabs: a warp then runs
            // max_lane(n0+n1+n2)/4 trips instead of sum_slabs(max_lane(n_slab)/4) — the lanes'
            // slabs fill differently, and padding every slab separately to the warp's worst lane
            // left two thirds of the candidate slots idle (ncu).  The running best
            // is one 64-bit key (dist^2 bits : index) so that "closer, ties to the lower index"
            // is a single unsigned comparison; non-negative floats order like their bit patterns.
            const unsigned total = n0 + n1 + n2;
            const unsigned o1 = ss[1] - n0, o2 = ss[2] - n0 - n1;   // virtual index -> global index offsets
            unsigned long long best = ((unsigned long long)__float_as_uint(thr) << 32) | 0x7fffffffull;
            unsigned bj = 0xffffffffu;
            for (unsigned v = 0; v < total; v += 4) {
                unsigned jj[4];
                float4 t[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const unsigned w = min(v + k, total - 1);
                    jj[k] = w < n0 ? ss[0] + w : (w < n0 + n1 ? o1 + w : o2 + w);
                    t[k] = __ldg(&pts[jj[k]]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float dx = t[k].x - qx, dy = t[k].y - qy, dz = t[k].z - qz;
                    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));   // canonical (see scan_range)
                    const unsigned long long key =
                            ((unsigned long long)__float_as_uint(d) << 32) | (unsigned)__float_as_int(t[k].w);
                    if (key <= best) {   // d >= 0, so the bit patterns order like the values; NaN sorts last
                        best = key;
                        bj = jj[k];
                    }
                }
            }
            if (bj != 0xffffffffu) {
                const float4 w = __ldg(&pts[bj]);
                b.d = __uint_as_float((unsigned)(best >> 32));
                b.idx = (int)(unsigned)(best & 0xffffffffull);
                b.j = (int)bj;
                b.x = w.x;
                b.y = w.y;
                b.z = w.z;
            }
        } else {
#pragma unroll
            for (int dz = 0; dz < 3; ++dz) {
                if (se[dz] - ss[dz] <= kSlabMax) {
                    scan_range(pts, ss[dz], se[dz], qx, qy, qz, b);
                } else {
                    const int plane = (z0 + dz) * g.ny;
                    for (int iy = y0; iy <= y1; ++iy) {
                        const int row = (plane + iy) * g.nx;
                        scan_range(pts, cs[row + x0], cs[row + x1 + 1], qx, qy, qz, b);
                    }
                }
            }
        }
    } else {
        for (int iz = z0; iz <= z1; ++iz)
            for (int iy = y0; iy <= y1; ++iy) {
                const int row = (iz * g.ny + iy) * g.nx;
                scan_range(pts, cs[row + x0], cs[row + x1 + 1], qx, qy, qz, b);
            }
    }
    if (b.j >= 0 && b.d <= r1_accept2) return;
    nn_search<true>(g, pts, cs, qx, qy, qz, rr, thr, b);
}

}  // namespace o3db
