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
                                         float gap2, float qx, float qy, float qz, Best& b) {
    // admissible |dx|: dx^2 <= best - gap2 (1e-6 best covers the rounding of the subtraction)
    const float ex = sqrtf(fmaf(b.d, 1e-6f, b.d - gap2)) * 1.00001f;
    const int xa = max(x0, cell1(lo_bound(qx, ex), g.ox, g.inv_c, g.nx));
    const int xb = min(x1, cell1(hi_bound(qx, ex), g.ox, g.inv_c, g.nx));
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
    const float lx = lo_bound(qx, rr), hx = hi_bound(qx, rr);
    const float ly = lo_bound(qy, rr), hy = hi_bound(qy, rr);
    const float lz = lo_bound(qz, rr), hz = hi_bound(qz, rr);
    // entirely outside the bounding box (or NaN): no candidate can pass
    if (hx < g.bmin[0] || lx > g.bmax[0] || hy < g.bmin[1] || ly > g.bmax[1] || hz < g.bmin[2] ||
        lz > g.bmax[2] || !(qx == qx) || !(qy == qy) || !(qz == qz))
        return;
    const int x0 = cell1(lx, g.ox, g.inv_c, g.nx), x1 = cell1(hx, g.ox, g.inv_c, g.nx);
    const int y0 = cell1(ly, g.oy, g.inv_c, g.ny), y1 = cell1(hy, g.oy, g.inv_c, g.ny);
    const int z0 = cell1(lz, g.oz, g.inv_c, g.nz), z1 = cell1(hz, g.oz, g.inv_c, g.nz);
    const int cx = cell1(qx, g.ox, g.inv_c, g.nx);
    const int cy = cell1(qy, g.oy, g.inv_c, g.ny), cz = cell1(qz, g.oz, g.inv_c, g.nz);
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
        scan_row(g, pts, cs, row, x0, x1, cx, 0.f, qx, qy, qz, b);         // rest of the own row
    }
    for (int iz = z0; iz <= z1; ++iz) {
        float gz = 0.f;
        if (iz > cz) gz = (g.oz + (float)iz * g.c) - qz - g.tol;
        else if (iz < cz) gz = qz - (g.oz + (float)(iz + 1) * g.c) - g.tol;
        gz = fmaxf(gz, 0.f);
        gz *= gz;
        if (gz > b.d) continue;
        for (int iy = y0; iy <= y1; ++iy) {
            if (iy == cy && iz == cz) continue;
            float gy = 0.f;
            if (iy > cy) gy = (g.oy + (float)iy * g.c) - qy - g.tol;
            else if (iy < cy) gy = qy - (g.oy + (float)(iy + 1) * g.c) - g.tol;
            gy = fmaxf(gy, 0.f);
            const float gap2 = fmaf(gy, gy, gz);
            if (gap2 > b.d) continue;  // strict: keeps exact ties reachable
            scan_row(g, pts, cs, (iz * g.ny + iy) * g.nx, x0, x1, -1, gap2, qx, qy, qz, b);
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
    if (hi_bound(qx, rr) < g.bmin[0] || lo_bound(qx, rr) > g.bmax[0] || hi_bound(qy, rr) < g.bmin[1] ||
        lo_bound(qy, rr) > g.bmax[1] || hi_bound(qz, rr) < g.bmin[2] || lo_bound(qz, rr) > g.bmax[2] ||
        !(qx == qx) || !(qy == qy) || !(qz == qz))
        return;
    const int x0 = cell1(lo_bound(qx, r1), g.ox, g.inv_c, g.nx), x1 = cell1(hi_bound(qx, r1), g.ox, g.inv_c, g.nx);
    const int y0 = cell1(lo_bound(qy, r1), g.oy, g.inv_c, g.ny), y1 = cell1(hi_bound(qy, r1), g.oy, g.inv_c, g.ny);
    const int z0 = cell1(lo_bound(qz, r1), g.oz, g.inv_c, g.nz), z1 = cell1(hi_bound(qz, r1), g.oz, g.inv_c, g.nz);
    if (y1 - y0 <= 2 && z1 - z0 <= 2) {
        // the usual case (r1 == cell size): at most 3 x 3 rows.  All 18 CSR offsets are
        // requested before the first candidate is touched (memory-level parallelism instead
        // of nine serialised L2 round trips); rows outside the box get an empty range.
        unsigned rs[9], re[9];
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const bool on = (z0 + dz <= z1) && (y0 + dy <= y1);
                const int row = (min(z0 + dz, z1) * g.ny + min(y0 + dy, y1)) * g.nx;
                const unsigned s = __ldg(&cs[row + x0]), e = __ldg(&cs[row + x1 + 1]);
                rs[dz * 3 + dy] = s;
                re[dz * 3 + dy] = on ? e : s;
            }
#pragma unroll
        for (int k = 0; k < 9; ++k) scan_range(pts, rs[k], re[k], qx, qy, qz, b);
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

// Same two-pass search, with pass 1 FLATTENED: the (up to) nine row ranges of the box are
// parked in shared memory (column threadIdx.x of s_rng[18][blockDim]) and walked by ONE loop
// per lane.  A warp then runs max_lane(total candidates) iterations instead of
// sum_rows(max_lane(row candidates)) — the lanes of a warp look at different rows, so the
// row-by-row version leaves most lanes idle most of the time (ncu: 14 of 32 lanes active).
template <int BLOCK>
__device__ __forceinline__ void nn_search_two_pass_flat(const Grid& g, const float4* __restrict__ pts,
                                                        const unsigned* __restrict__ cs, float qx, float qy,
                                                        float qz, float r1, float r1_accept2, float rr, float thr,
                                                        unsigned (*s_rng)[BLOCK], Best& b) {
    b.d = thr;
    b.j = -1;
    b.idx = 0x7fffffff;
    b.x = b.y = b.z = 0.f;
    if (hi_bound(qx, rr) < g.bmin[0] || lo_bound(qx, rr) > g.bmax[0] || hi_bound(qy, rr) < g.bmin[1] ||
        lo_bound(qy, rr) > g.bmax[1] || hi_bound(qz, rr) < g.bmin[2] || lo_bound(qz, rr) > g.bmax[2] ||
        !(qx == qx) || !(qy == qy) || !(qz == qz))
        return;
    const int x0 = cell1(lo_bound(qx, r1), g.ox, g.inv_c, g.nx), x1 = cell1(hi_bound(qx, r1), g.ox, g.inv_c, g.nx);
    const int y0 = cell1(lo_bound(qy, r1), g.oy, g.inv_c, g.ny), y1 = cell1(hi_bound(qy, r1), g.oy, g.inv_c, g.ny);
    const int z0 = cell1(lo_bound(qz, r1), g.oz, g.inv_c, g.nz), z1 = cell1(hi_bound(qz, r1), g.oz, g.inv_c, g.nz);
    if (y1 - y0 <= 2 && z1 - z0 <= 2) {
        const int t = threadIdx.x;
        unsigned rs[9], re[9];
#pragma unroll
        for (int dz = 0; dz < 3; ++dz)
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
                const bool on = (z0 + dz <= z1) && (y0 + dy <= y1);
                const int row = (min(z0 + dz, z1) * g.ny + min(y0 + dy, y1)) * g.nx;
                const unsigned s = __ldg(&cs[row + x0]), e = __ldg(&cs[row + x1 + 1]);
                rs[dz * 3 + dy] = s;
                re[dz * 3 + dy] = on ? e : s;
            }
#pragma unroll
        for (int k = 1; k < 9; ++k) {   // row 0 stays in registers
            s_rng[2 * k][t] = rs[k];
            s_rng[2 * k + 1][t] = re[k];
        }
        unsigned j = rs[0], e = re[0];
        int row = 0;
        for (;;) {
            while (j >= e) {
                if (++row >= 9) goto pass1_done;
                j = s_rng[2 * row][t];
                e = s_rng[2 * row + 1][t];
            }
            const float4 c = __ldg(&pts[j]);
            const float dx = c.x - qx, dy = c.y - qy, dz = c.z - qz;
            const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
            const int idx = __float_as_int(c.w);
            if (d < b.d || (d == b.d && idx < b.idx)) {
                b.d = d;
                b.j = (int)j;
                b.idx = idx;
                b.x = c.x;
                b.y = c.y;
                b.z = c.z;
            }
            ++j;
        }
    pass1_done:;
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

// ---------------------------------------------------------------------------
// Group search: G consecutive lanes (spatially neighbouring queries of the cell-sorted
// source) share ONE candidate list, so every lane of the group runs the same loops with the
// same addresses (no intra-group divergence, broadcast loads):
//   stage A  all points of the cells that contain the group's queries ("core box")
//   stage B  whatever else lies within sqrt(best) of any lane (union of the lanes' boxes
//            for their current best distance), minus the core box
// Every point within a lane's best distance is visited, so the result is the exact nearest
// neighbour (ties -> lower index), identical to nn_search().  If the group is not compact
// (its union box is much larger than a single lane's), each lane falls back to nn_search().
// Must be called by all 32 lanes (inactive lanes pass valid = false).
// Out-of-line copy of the per-lane search for the rare non-compact group (keeps the hot
// path's register budget).
__device__ __noinline__ void nn_search_slow(const Grid& g, const float4* __restrict__ pts,
                                            const unsigned* __restrict__ cs, float qx, float qy, float qz,
                                            float rr, float thr, Best& b) {
    nn_search<true>(g, pts, cs, qx, qy, qz, rr, thr, b);
}

// Shuffles name only the G lanes of the group in their mask: groups of one warp take
// different paths (early return of non-compact groups, different trip counts), so a
// full-warp mask would dead-lock.
template <int G>
__device__ __forceinline__ unsigned group_mask() {
    return G >= 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
}

template <int G>
__device__ __forceinline__ void group_minmax(int& lo, int& hi) {
    const unsigned m = group_mask<G>();
#pragma unroll
    for (int o = 1; o < G; o <<= 1) {
        lo = min(lo, __shfl_xor_sync(m, lo, o));
        hi = max(hi, __shfl_xor_sync(m, hi, o));
    }
}

template <int G>
__device__ __forceinline__ void nn_search_group(const Grid& g, const float4* __restrict__ pts,
                                                const unsigned* __restrict__ cs, bool valid, float qx, float qy,
                                                float qz, float rr, float thr, Best& b) {
    b.d = thr;
    b.j = -1;
    b.idx = 0x7fffffff;
    b.x = b.y = b.z = 0.f;
    const float lx = lo_bound(qx, rr), hx = hi_bound(qx, rr);
    const float ly = lo_bound(qy, rr), hy = hi_bound(qy, rr);
    const float lz = lo_bound(qz, rr), hz = hi_bound(qz, rr);
    const bool inside = valid && !(hx < g.bmin[0] || lx > g.bmax[0] || hy < g.bmin[1] || ly > g.bmax[1] ||
                                   hz < g.bmin[2] || lz > g.bmax[2] || !(qx == qx) || !(qy == qy) || !(qz == qz));
    const int kBig = 0x3fffffff;
    // ---- stage A: the cells holding the group's queries
    const int cx = cell1(qx, g.ox, g.inv_c, g.nx), cy = cell1(qy, g.oy, g.inv_c, g.ny),
              cz = cell1(qz, g.oz, g.inv_c, g.nz);
    int CX0 = inside ? cx : kBig, CX1 = inside ? cx : -kBig;
    int CY0 = inside ? cy : kBig, CY1 = inside ? cy : -kBig;
    int CZ0 = inside ? cz : kBig, CZ1 = inside ? cz : -kBig;
    group_minmax<G>(CX0, CX1);
    group_minmax<G>(CY0, CY1);
    group_minmax<G>(CZ0, CZ1);
    // compactness guard (group-uniform): a group straddling the end of a cell row would drag
    // in a huge box; such groups use the per-lane search instead
    const bool compact = CX0 <= CX1 && (CX1 - CX0) <= 6 && (CY1 - CY0) <= 3 && (CZ1 - CZ0) <= 3;
    if (!compact) {
        if (inside) nn_search<true>(g, pts, cs, qx, qy, qz, rr, thr, b);
        return;
    }
    for (int iz = CZ0; iz <= CZ1; ++iz)
        for (int iy = CY0; iy <= CY1; ++iy) {
            const int row = (iz * g.ny + iy) * g.nx;
            scan_range(pts, cs[row + CX0], cs[row + CX1 + 1], qx, qy, qz, b);
        }
    // ---- stage B: union of the lanes' boxes for their current best distance
    const float e = sqrtf(b.d) * 1.00001f;   // b.d <= thr; the slack covers sqrt/rounding
    int UX0 = kBig, UX1 = -kBig, UY0 = kBig, UY1 = -kBig, UZ0 = kBig, UZ1 = -kBig;
    if (inside) {
        UX0 = cell1(lo_bound(qx, e), g.ox, g.inv_c, g.nx);
        UX1 = cell1(hi_bound(qx, e), g.ox, g.inv_c, g.nx);
        UY0 = cell1(lo_bound(qy, e), g.oy, g.inv_c, g.ny);
        UY1 = cell1(hi_bound(qy, e), g.oy, g.inv_c, g.ny);
        UZ0 = cell1(lo_bound(qz, e), g.oz, g.inv_c, g.nz);
        UZ1 = cell1(hi_bound(qz, e), g.oz, g.inv_c, g.nz);
    }
    group_minmax<G>(UX0, UX1);
    group_minmax<G>(UY0, UY1);
    group_minmax<G>(UZ0, UZ1);
    for (int iz = UZ0; iz <= UZ1; ++iz)
        for (int iy = UY0; iy <= UY1; ++iy) {
            const int row = (iz * g.ny + iy) * g.nx;
            const bool core_row = iz >= CZ0 && iz <= CZ1 && iy >= CY0 && iy <= CY1;
            if (!core_row) {
                scan_range(pts, cs[row + UX0], cs[row + UX1 + 1], qx, qy, qz, b);
            } else {
                if (UX0 < CX0) scan_range(pts, cs[row + UX0], cs[row + CX0], qx, qy, qz, b);
                if (UX1 > CX1) scan_range(pts, cs[row + CX1 + 1], cs[row + UX1 + 1], qx, qy, qz, b);
            }
        }
}

}  // namespace o3db
