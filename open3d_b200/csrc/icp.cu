// icp.cu — correspondence search, pose reduction and the fused device-resident
// point-to-plane ICP loop for sm_100a.  See include/open3d_b200.h for the
// reference interfaces each entry point replaces and DESIGN.md for the design.
//
// No CPU fallback: every entry point needs a CUDA device.
#include <cfloat>
#include <cstddef>
#include <algorithm>
#include <climits>
#include <cmath>
#include <new>
#include <vector>

#include "comm.h"
#include "common.cuh"
#include "grid.cuh"
#include "reduce.cuh"

namespace o3db {

static constexpr int64_t kMaxCells = int64_t(1) << 26;  // 256 MB of u32 CSR offsets at most
static constexpr int kMaxCellsPerAxis = 4096;
#ifndef ICP_THIN_FACTOR
#define ICP_THIN_FACTOR 8
#endif
static constexpr double kThinFactor = ICP_THIN_FACTOR;   // grid-x (thin axis) cells are this much coarser
#ifndef ICP_PREFETCH_SRC
#define ICP_PREFETCH_SRC 0   // 1: prefetch the next chunk's source point (round-2 candidate; measure with profiles/tune_icp.sh)
#endif
#ifndef ICP_CELL_SCALE
#define ICP_CELL_SCALE 0.5
#endif
static constexpr double kDefaultCellScale = ICP_CELL_SCALE;
#ifndef ICP_TWO_PASS
#define ICP_TWO_PASS 1   // 1: slab/two-pass search on the fine grid (default), 0: pruned single-pass search
#endif
static constexpr bool kTwoPass = ICP_TWO_PASS != 0;
#ifndef ICP_MIN_BLOCKS
#define ICP_MIN_BLOCKS 3
#endif
#ifndef ICP_ACC_SMEM
#define ICP_ACC_SMEM 0
#endif
#ifndef ICP_DEFAULT_VARIANT
#define ICP_DEFAULT_VARIANT 1
#endif

// --------------------------------------------------------------------- bbox

__global__ void bbox_kernel(const float* __restrict__ pts, int64_t n, unsigned* __restrict__ bbox /*6*/) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const float v = pts[3 * i + a];
            mn[a] = fminf(mn[a], v);
            mx[a] = fmaxf(mx[a], v);
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(0xffffffffu, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(0xffffffffu, mx[a], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            atomicMin(&bbox[a], float_to_ordered(mn[a]));
            atomicMax(&bbox[3 + a], float_to_ordered(mx[a]));
        }
    }
}

// ------------------------------------------------------------ counting sort

struct Affine {  // row-major 3x4 + projective row, f32 (TransformImpl.h:20-45)
    float m[16];
};

__device__ __forceinline__ void apply_transform(const float* __restrict__ T, float& x, float& y, float& z) {
    const float px = x, py = y, pz = z;
    const float ox = T[0] * px + T[1] * py + T[2] * pz + T[3];
    const float oy = T[4] * px + T[5] * py + T[6] * pz + T[7];
    const float oz = T[8] * px + T[9] * py + T[10] * pz + T[11];
    const float ow = T[12] * px + T[13] * py + T[14] * pz + T[15];
    if (ow == 1.0f) {   // rigid transforms: x / 1 == x exactly, skip three IEEE divisions
        x = ox;
        y = oy;
        z = oz;
    } else {
        x = ox / ow;
        y = oy / ow;
        z = oz / ow;
    }
}

// Sort key of the working source: tile-major (16 x 16 x 4 cells per tile, cells row-major
// inside), so that 256 consecutive sorted queries occupy a compact 3-D box whose candidate
// rows can be staged in shared memory.  Only locality depends on it, never correctness.
static constexpr int kTileX = 16, kTileY = 16, kTileZ = 4;
__host__ __device__ inline int64_t tiled_key_space(int nx, int ny, int nz) {
    return (int64_t)((nx + kTileX - 1) / kTileX) * ((ny + kTileY - 1) / kTileY) * ((nz + kTileZ - 1) / kTileZ) *
           (kTileX * kTileY * kTileZ);
}
__device__ __forceinline__ unsigned cell_key_tiled(const Grid& g, float rx, float ry, float rz) {
    float x, y, z;
    to_grid(g, rx, ry, rz, x, y, z);
    const int ix = cell1(x, g.ox, g.inv_cx, g.nx), iy = cell1(y, g.oy, g.inv_c, g.ny), iz = cell1(z, g.oz, g.inv_c, g.nz);
    const int ntx = (g.nx + kTileX - 1) / kTileX, nty = (g.ny + kTileY - 1) / kTileY;
    const int tile = ((iz / kTileZ) * nty + iy / kTileY) * ntx + ix / kTileX;
    const int local = ((iz % kTileZ) * kTileY + iy % kTileY) * kTileX + ix % kTileX;
    return (unsigned)tile * (unsigned)(kTileX * kTileY * kTileZ) + (unsigned)local;
}

// key[i] = cell of (optionally transformed) point i, rank[i] = arrival order in the cell.
// TRANSFORM = true is the source path (tile-major key); false the target (row-major key).
template <bool TRANSFORM>
__global__ void count_kernel(const float* __restrict__ pts, int64_t n, Grid g, Affine T,
                             unsigned* __restrict__ count, unsigned* __restrict__ key,
                             unsigned* __restrict__ rank) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (TRANSFORM) apply_transform(T.m, x, y, z);
    const unsigned k = TRANSFORM ? cell_key_tiled(g, x, y, z) : cell_key(g, x, y, z);
    key[i] = k;
    rank[i] = atomicAdd(&count[k], 1u);
}

template <bool TRANSFORM>
__global__ void scatter_kernel(const float* __restrict__ pts, const float* __restrict__ nrm, int64_t n,
                               Affine T, const unsigned* __restrict__ start,
                               const unsigned* __restrict__ key, const unsigned* __restrict__ rank,
                               float4* __restrict__ pts4, float4* __restrict__ nrm4) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = pts[3 * i], y = pts[3 * i + 1], z = pts[3 * i + 2];
    if (TRANSFORM) apply_transform(T.m, x, y, z);
    const unsigned p = start[key[i]] + rank[i];
    pts4[p] = make_float4(x, y, z, __int_as_float((int)i));
    if (nrm4) nrm4[p] = make_float4(nrm[3 * i], nrm[3 * i + 1], nrm[3 * i + 2], 0.f);
}

// Exclusive scan of u32 data[0..n) in place; data[n] receives the total.
static constexpr int kScanItems = 8;
static constexpr int kScanTile = kThreads * kScanItems;

__device__ __forceinline__ unsigned block_exclusive_scan(unsigned v, unsigned& total) {
    __shared__ unsigned s_w[kThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned x = lane < kThreads / 32 ? s_w[lane] : 0u;
#pragma unroll
        for (int o = 1; o < kThreads / 32; o <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += t;
        }
        if (lane < kThreads / 32) s_w[lane] = x;  // inclusive over warps
    }
    __syncthreads();
    total = s_w[kThreads / 32 - 1];
    const unsigned base = w ? s_w[w - 1] : 0u;
    __syncthreads();
    return base + inc - v;
}

__global__ void scan_tile_sums(const unsigned* __restrict__ data, int64_t n, unsigned* __restrict__ tile_sums) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k)
        if (base + k < n) s += data[base + k];
    unsigned total;
    block_exclusive_scan(s, total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void scan_tile_offsets(unsigned* __restrict__ tile_sums, int64_t ntiles, unsigned* __restrict__ grand_total) {
    unsigned carry = 0;
    for (int64_t b = 0; b < ntiles; b += kThreads) {
        const int64_t i = b + threadIdx.x;
        const unsigned v = i < ntiles ? tile_sums[i] : 0u;
        unsigned total;
        const unsigned ex = block_exclusive_scan(v, total);
        if (i < ntiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

__global__ void scan_apply(unsigned* __restrict__ data, int64_t n, const unsigned* __restrict__ tile_offsets) {
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    unsigned v[kScanItems];
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        v[k] = base + k < n ? data[base + k] : 0u;
        s += v[k];
    }
    unsigned total;
    unsigned ex = block_exclusive_scan(s, total) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) data[base + k] = ex;
        ex += v[k];
    }
}

// data has n+1 entries; scratch has ceil(n/kScanTile) entries.
static int exclusive_scan_u32(unsigned* data, int64_t n, unsigned* scratch, cudaStream_t st) {
    const int64_t ntiles = ceil_div(n, kScanTile);
    scan_tile_sums<<<(unsigned)ntiles, kThreads, 0, st>>>(data, n, scratch);
    O3DB_LAUNCH_CHECK();
    scan_tile_offsets<<<1, kThreads, 0, st>>>(scratch, ntiles, data + n);
    O3DB_LAUNCH_CHECK();
    scan_apply<<<(unsigned)ntiles, kThreads, 0, st>>>(data, n, scratch);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}


// ------------------------------------- reference-layout CSR table (interop)

// core/nns/NeighborSearchCommon.h:31-52 + FixedRadiusSearchImpl.cuh:63-134.
__device__ __forceinline__ unsigned ref_bucket(const float* __restrict__ p, float inv_voxel, unsigned table_size) {
    const int vx = (int)floorf(p[0] * inv_voxel), vy = (int)floorf(p[1] * inv_voxel), vz = (int)floorf(p[2] * inv_voxel);
    const unsigned h32 = ((unsigned)vx * 73856096u) ^ ((unsigned)vy * 193649663u) ^ ((unsigned)vz * 83492791u);
    const uint64_t h = (uint64_t)(int64_t)(int)h32;   // int -> size_t sign-extends
    return (unsigned)(h % (uint64_t)table_size);
}

__global__ void ref_count_kernel(const float* __restrict__ pts, int64_t n, float inv_voxel, unsigned table_size,
                                 unsigned* __restrict__ count, unsigned* __restrict__ rank) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    rank[i] = atomicAdd(&count[ref_bucket(pts + 3 * i, inv_voxel, table_size)], 1u);
}

__global__ void ref_scatter_kernel(const float* __restrict__ pts, int64_t n, float inv_voxel, unsigned table_size,
                                   const unsigned* __restrict__ start, const unsigned* __restrict__ rank,
                                   unsigned* __restrict__ index) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    index[start[ref_bucket(pts + 3 * i, inv_voxel, table_size)] + rank[i]] = (unsigned)i;
}

// ------------------------------------------------------------------ index

}  // namespace o3db

using namespace o3db;

struct o3db_nns {
    Grid g{};
    double radius = 0;
    int64_t m = 0;
    int64_t ncell = 0;
    float4* pts4 = nullptr;
    float4* nrm4 = nullptr;
    unsigned* cell_start = nullptr;
};

namespace o3db {

static void nns_free(o3db_nns* s, cudaStream_t st) {
    if (!s) return;
    if (s->pts4) cudaFreeAsync(s->pts4, st);
    if (s->nrm4) cudaFreeAsync(s->nrm4, st);
    if (s->cell_start) cudaFreeAsync(s->cell_start, st);
    s->pts4 = s->nrm4 = nullptr;
    s->cell_start = nullptr;
}

static int grid_from_bbox(const float mn[3], const float mx[3], double radius, double cell_scale, Grid* g,
                          int64_t* ncell) {
    double c = radius * (cell_scale > 0 ? cell_scale : 1.0) * (1.0 + 1e-4);
    if (!(c > 0) || !std::isfinite(c)) c = 1.0;
    double ext[3];
    double maxabs = 0;
    for (int a = 0; a < 3; ++a) {
        ext[a] = std::max(0.0, (double)mx[a] - (double)mn[a]);
        maxabs = std::max(maxabs, std::max(std::fabs((double)mn[a]), std::fabs((double)mx[a])));
    }
    // grid axis order: fastest = the real axis with the smallest extent (ties: z, then y),
    // slowest = the one with the largest
    int order[3] = {2, 1, 0};
    std::stable_sort(order, order + 3, [&](int a, int b) { return ext[a] < ext[b]; });
    for (int k = 0; k < 3; ++k) g->ax[k] = order[k];
    for (;;) {
        double n[3];
        bool ok = true;
        double prod = 1;
        for (int k = 0; k < 3; ++k) {
            n[k] = std::floor(ext[order[k]] / (k == 0 ? c * kThinFactor : c)) + 1;
            ok = ok && n[k] <= kMaxCellsPerAxis;
            prod *= n[k];
        }
        if (ok && prod <= (double)kMaxCells) {
            g->nx = (int)n[0];
            g->ny = (int)n[1];
            g->nz = (int)n[2];
            break;
        }
        c *= 1.25;
    }
    g->c = (float)c;
    g->inv_c = 1.0f / g->c;
    g->cx = (float)(c * kThinFactor);
    g->inv_cx = 1.0f / g->cx;
    g->ox = mn[order[0]];
    g->oy = mn[order[1]];
    g->oz = mn[order[2]];
    g->tol = (float)(c * 2e-3 + maxabs * 1e-6);
    for (int k = 0; k < 3; ++k) {
        g->bmin[k] = mn[order[k]];
        g->bmax[k] = mx[order[k]];
    }
    *ncell = (int64_t)g->nx * g->ny * g->nz;
    return O3DB_OK;
}

// Builds pts4 / nrm4 / cell_start for `pts` (device).  Synchronises `st` once
// (the grid dimensions depend on the bounding box).
static int nns_build(o3db_nns* s, const float* pts, const float* nrm, int64_t m, double radius,
                     double cell_scale, cudaStream_t st) {
    configure_memory_pool();
    s->m = m;
    s->radius = radius;
    unsigned* d_bbox = nullptr;
    O3DB_CUDA_CHECK(cudaMallocAsync(&d_bbox, 6 * sizeof(unsigned), st));
    const unsigned init[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    O3DB_CUDA_CHECK(cudaMemcpyAsync(d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
    const int blocks = (int)std::min<int64_t>(ceil_div(m, kThreads), (int64_t)num_sms() * 8);
    bbox_kernel<<<blocks, kThreads, 0, st>>>(pts, m, d_bbox);
    O3DB_LAUNCH_CHECK();
    unsigned h_bbox[6];
    O3DB_CUDA_CHECK(cudaMemcpyAsync(h_bbox, d_bbox, sizeof(h_bbox), cudaMemcpyDeviceToHost, st));
    O3DB_CUDA_CHECK(cudaStreamSynchronize(st));
    O3DB_CUDA_CHECK(cudaFreeAsync(d_bbox, st));
    float mn[3], mx[3];
    for (int a = 0; a < 3; ++a) {
        mn[a] = ordered_to_float(h_bbox[a]);
        mx[a] = ordered_to_float(h_bbox[3 + a]);
        if (!(mn[a] <= mx[a]) || !std::isfinite(mn[a]) || !std::isfinite(mx[a])) {
            set_last_error("target point cloud has non-finite coordinates");
            return O3DB_ERR_INVALID;
        }
    }
    grid_from_bbox(mn, mx, radius, cell_scale, &s->g, &s->ncell);

    unsigned *key = nullptr, *rank = nullptr, *scratch = nullptr;
    O3DB_CUDA_CHECK(cudaMallocAsync(&s->cell_start, (s->ncell + 1) * sizeof(unsigned), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&s->pts4, m * sizeof(float4), st));
    if (nrm) O3DB_CUDA_CHECK(cudaMallocAsync(&s->nrm4, m * sizeof(float4), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&key, m * sizeof(unsigned), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&rank, m * sizeof(unsigned), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&scratch, ceil_div(s->ncell, kScanTile) * sizeof(unsigned), st));
    O3DB_CUDA_CHECK(cudaMemsetAsync(s->cell_start, 0, (s->ncell + 1) * sizeof(unsigned), st));
    Affine id{};
    const unsigned nb = (unsigned)ceil_div(m, kThreads);
    count_kernel<false><<<nb, kThreads, 0, st>>>(pts, m, s->g, id, s->cell_start, key, rank);
    O3DB_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(s->cell_start, s->ncell, scratch, st);
    if (rc) return rc;
    scatter_kernel<false><<<nb, kThreads, 0, st>>>(pts, nrm, m, id, s->cell_start, key, rank, s->pts4, s->nrm4);
    O3DB_LAUNCH_CHECK();
    O3DB_CUDA_CHECK(cudaFreeAsync(key, st));
    O3DB_CUDA_CHECK(cudaFreeAsync(rank, st));
    O3DB_CUDA_CHECK(cudaFreeAsync(scratch, st));
    return O3DB_OK;
}

// ------------------------------------------------------- stand-alone search

template <bool PRUNE>
__global__ void __launch_bounds__(kThreads)
hybrid_search_k1_kernel(Grid g, const float4* __restrict__ pts, const unsigned* __restrict__ cs,
                        const float* __restrict__ q, int64_t n, float rr, float thr,
                        int32_t* __restrict__ idx, float* __restrict__ dist, int32_t* __restrict__ cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Best b;
    nn_search<PRUNE>(g, pts, cs, q[3 * i], q[3 * i + 1], q[3 * i + 2], rr, thr, b);
    if (idx) idx[i] = b.j >= 0 ? b.idx : -1;
    if (dist) dist[i] = b.j >= 0 ? b.d : 0.f;
    if (cnt) cnt[i] = b.j >= 0 ? 1 : 0;
}

static constexpr int kMaxKnn = 32;

// General max_knn (small): per-thread sorted list, same visiting scheme without pruning.
__global__ void __launch_bounds__(kThreads)
hybrid_search_knn_kernel(Grid g, const float4* __restrict__ pts, const unsigned* __restrict__ cs,
                         const float* __restrict__ q, int64_t n, float rr, float thr, int k,
                         int32_t* __restrict__ idx, float* __restrict__ dist, int32_t* __restrict__ cnt) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float qx = q[3 * i], qy = q[3 * i + 1], qz = q[3 * i + 2];
    int bi[kMaxKnn];
    float bd[kMaxKnn];
    int c = 0;
    float gx, gy, gz;
    to_grid(g, qx, qy, qz, gx, gy, gz);
    const float lx = lo_bound(gx, rr), hx = hi_bound(gx, rr);
    const float ly = lo_bound(gy, rr), hy = hi_bound(gy, rr);
    const float lz = lo_bound(gz, rr), hz = hi_bound(gz, rr);
    const bool outside = hx < g.bmin[0] || lx > g.bmax[0] || hy < g.bmin[1] || ly > g.bmax[1] ||
                         hz < g.bmin[2] || lz > g.bmax[2] || !(qx == qx) || !(qy == qy) || !(qz == qz);
    if (!outside) {
        const int x0 = cell1(lx, g.ox, g.inv_cx, g.nx), x1 = cell1(hx, g.ox, g.inv_cx, g.nx);
        const int y0 = cell1(ly, g.oy, g.inv_c, g.ny), y1 = cell1(hy, g.oy, g.inv_c, g.ny);
        const int z0 = cell1(lz, g.oz, g.inv_c, g.nz), z1 = cell1(hz, g.oz, g.inv_c, g.nz);
        for (int iz = z0; iz <= z1; ++iz)
            for (int iy = y0; iy <= y1; ++iy) {
                const int row = (iz * g.ny + iy) * g.nx;
                const unsigned s = cs[row + x0], e = cs[row + x1 + 1];
                for (unsigned j = s; j < e; ++j) {
                    const float4 t = __ldg(&pts[j]);
                    const float dx = t.x - qx, dy = t.y - qy, dz = t.z - qz;
                    const float d = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                    if (!(d <= thr)) continue;
                    const int id = __float_as_int(t.w);
                    int m = c;
                    if (m == k) {
                        if (d > bd[m - 1] || (d == bd[m - 1] && id > bi[m - 1])) continue;
                        m = k - 1;
                    }
                    int pos = m;
                    while (pos > 0 && (bd[pos - 1] > d || (bd[pos - 1] == d && bi[pos - 1] > id))) {
                        bd[pos] = bd[pos - 1];
                        bi[pos] = bi[pos - 1];
                        --pos;
                    }
                    bd[pos] = d;
                    bi[pos] = id;
                    c = m + 1;
                }
            }
    }
    for (int j = 0; j < k; ++j) {
        if (idx) idx[i * k + j] = j < c ? bi[j] : -1;
        if (dist) dist[i * k + j] = j < c ? bd[j] : 0.f;
    }
    if (cnt) cnt[i] = c;
}

// --------------------------------------------------------- robust kernels

struct Robust {
    int method;
    float scale;
    double shape;
};

// RobustKernelImpl.h:35-115 for scalar_t = float (double literals promote as upstream).
__device__ __forceinline__ float robust_weight(const Robust& k, float r) {
    switch (k.method) {
        case O3DB_ROBUST_L2: return 1.0f;
        case O3DB_ROBUST_L1: return (float)(1.0 / fabsf(r));
        case O3DB_ROBUST_HUBER: return k.scale / fmaxf(fabsf(r), k.scale);
        case O3DB_ROBUST_CAUCHY: {
            const float q = r / k.scale;
            return (float)(1.0 / (1.0 + (double)(q * q)));
        }
        case O3DB_ROBUST_GM: {
            const float s = k.scale + r * r;
            return k.scale / (s * s);
        }
        case O3DB_ROBUST_TUKEY: {
            const float q = fminf(1.0f, fabsf(r) / k.scale);
            const double v = 1.0 - (double)(q * q);
            return (float)(v * v);
        }
        default: {  // generalized
            const float s2 = k.scale * k.scale;
            // open3d::IsClose (GeometryMacros.h:58-63) is relative: the shape ~ 0 branch of
            // RobustKernelImpl.h:85-91 can never be taken; only shape ~ 2 is special-cased.
            if (k.shape > (1.0 - 1e-3) * 2.0 && k.shape < (1.0 + 1e-3) * 2.0) return (float)(1.0 / (double)s2);
            const float q = r / k.scale;
            if (k.shape < -1e7) return (float)(exp((double)(q * q) / (-2.0)) / (double)s2);
            return (float)(pow((double)(q * q) / fabs(k.shape - 2.0) + 1, (k.shape / 2.0) - 1.0) / (double)s2);
        }
    }
}

// ------------------------------------------------- 29(+1)-scalar reduction (reduce.cuh)

__device__ __forceinline__ void flush_acc_smem(float (*s_acc)[kThreads], double (*s_warp)[kSumStride]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) {
        const double v = warp_sum((double)s_acc[k][threadIdx.x]);
        if (lane == 0) s_warp[w][k] += v;
        s_acc[k][threadIdx.x] = 0.f;
    }
}

// RegistrationImpl.h:251-287 + RegistrationCUDA.cu:29-79 slot layout.
template <bool L2LOSS>
__device__ __forceinline__ void accumulate_p2plane(float (&acc)[kNumSums], const Robust& rk, float sx, float sy,
                                                   float sz, float tx, float ty, float tz, float nx, float ny,
                                                   float nz) {
    // r and J are evaluated without FMA contraction, in the reference's operation order, so
    // that residual-dependent robust weights (e.g. L1: 1/|r|) see the same r as the CPU path
    // even when (s - t).n cancels to ~0; the accumulation below may fuse.
    const float r = __fadd_rn(__fadd_rn(__fmul_rn(__fsub_rn(sx, tx), nx), __fmul_rn(__fsub_rn(sy, ty), ny)),
                              __fmul_rn(__fsub_rn(sz, tz), nz));
    float J[6];
    J[0] = __fsub_rn(__fmul_rn(nz, sy), __fmul_rn(ny, sz));
    J[1] = __fsub_rn(__fmul_rn(nx, sz), __fmul_rn(nz, sx));
    J[2] = __fsub_rn(__fmul_rn(ny, sx), __fmul_rn(nx, sy));
    J[3] = nx;
    J[4] = ny;
    J[5] = nz;
    const float w = L2LOSS ? 1.0f : robust_weight(rk, r);
    int s = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float jw = J[j] * w;
#pragma unroll
        for (int k = 0; k <= j; ++k) acc[s++] += jw * J[k];
        acc[21 + j] += jw * r;
    }
    acc[27] += r;
    acc[28] += 1.0f;
}

// ComputePoseColoredICP per-correspondence body (RegistrationImpl.h:337-425 / RegistrationCUDA.cu:119-183):
// geometric + photometric Jacobian rows, two robust weights, same 29-slot layout.
template <bool L2LOSS>
__device__ __forceinline__ void accumulate_colored(float (&acc)[kNumSums], const Robust& rk, const float (&vs)[3],
                                                   const float (&vt)[3], const float (&nt)[3], float is, float it,
                                                   const float (&dit)[3], float sqrt_lg, float sqrt_lp) {
    const float d = (vs[0] - vt[0]) * nt[0] + (vs[1] - vt[1]) * nt[1] + (vs[2] - vt[2]) * nt[2];
    float JG[6], JI[6];
    JG[0] = sqrt_lg * (-vs[2] * nt[1] + vs[1] * nt[2]);
    JG[1] = sqrt_lg * (vs[2] * nt[0] - vs[0] * nt[2]);
    JG[2] = sqrt_lg * (-vs[1] * nt[0] + vs[0] * nt[1]);
    JG[3] = sqrt_lg * nt[0];
    JG[4] = sqrt_lg * nt[1];
    JG[5] = sqrt_lg * nt[2];
    const float rG = sqrt_lg * d;
    const float vp[3] = {vs[0] - d * nt[0], vs[1] - d * nt[1], vs[2] - d * nt[2]};
    const float is_proj = dit[0] * (vp[0] - vt[0]) + dit[1] * (vp[1] - vt[1]) + dit[2] * (vp[2] - vt[2]) + it;
    const float sd = dit[0] * nt[0] + dit[1] * nt[1] + dit[2] * nt[2];
    const float dM[3] = {sd * nt[0] - dit[0], sd * nt[1] - dit[1], sd * nt[2] - dit[2]};
    JI[0] = sqrt_lp * (-vs[2] * dM[1] + vs[1] * dM[2]);
    JI[1] = sqrt_lp * (vs[2] * dM[0] - vs[0] * dM[2]);
    JI[2] = sqrt_lp * (-vs[1] * dM[0] + vs[0] * dM[1]);
    JI[3] = sqrt_lp * dM[0];
    JI[4] = sqrt_lp * dM[1];
    JI[5] = sqrt_lp * dM[2];
    const float rI = sqrt_lp * (is - is_proj);
    const float wG = L2LOSS ? 1.0f : robust_weight(rk, rG);
    const float wI = L2LOSS ? 1.0f : robust_weight(rk, rI);
    int p = 0;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
#pragma unroll
        for (int k = 0; k <= j; ++k) acc[p++] += JG[j] * wG * JG[k] + JI[j] * wI * JI[k];
        acc[21 + j] += JG[j] * wG * rG + JI[j] * wI * rI;
    }
    acc[27] += rG * rG + rI * rI;
    acc[28] += 1.0f;
}

// intensity of an RGB triple exactly as upstream: float sum, divided by the double literal 3.0
__device__ __forceinline__ float color_intensity(float r, float g, float b) { return (float)((r + g + b) / 3.0); }

// -------------------------------------------- stand-alone pose reductions

struct PoseOut {
    double* sums29;
    double* pose;
    int* status;  // device int: 0 ok, 1 singular
};

// ComputePosePointToPlaneKernelCUDA (RegistrationCUDA.cu:29-79) with int64 correspondences.
template <bool L2LOSS>
__global__ void __launch_bounds__(kThreads)
pose_p2plane_kernel(const float* __restrict__ src, const float* __restrict__ tgt, const float* __restrict__ nrm,
                    const int64_t* __restrict__ corr, int64_t n, Robust rk, double* __restrict__ partials,
                    unsigned* ticket, PoseOut out) {
    __shared__ double s_warp[kThreads / 32][kSumStride];
    __shared__ double s_final[kSumStride];
    for (int k = threadIdx.x; k < (kThreads / 32) * kSumStride; k += kThreads) (&s_warp[0][0])[k] = 0.0;
    __syncthreads();
    float acc[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) acc[k] = 0.f;
    int since = 0;
    for (int64_t base = (int64_t)blockIdx.x * kThreads; base < n; base += (int64_t)gridDim.x * kThreads) {
        const int64_t i = base + threadIdx.x;
        if (i < n) {
            const int64_t c = corr[i];
            if (c != -1) {
                const float* t = tgt + 3 * c;
                const float* m = nrm + 3 * c;
                accumulate_p2plane<L2LOSS>(acc, rk, src[3 * i], src[3 * i + 1], src[3 * i + 2], t[0], t[1], t[2],
                                           m[0], m[1], m[2]);
            }
        }
        if (++since == kFlushEvery) {
            flush_acc(acc, s_warp);
            since = 0;
        }
    }
    flush_acc(acc, s_warp);
    if (!block_reduce_to_global(s_warp, partials, ticket, s_final)) return;
    if (threadIdx.x < 29 && out.sums29) out.sums29[threadIdx.x] = s_final[threadIdx.x];
    if (threadIdx.x == 0) {
        double pose[6] = {0, 0, 0, 0, 0, 0};
        const bool ok = solve6x6(s_final, pose);
        if (!ok)
            for (int k = 0; k < 6; ++k) pose[k] = 0.0;
        if (out.pose)
            for (int k = 0; k < 6; ++k) out.pose[k] = pose[k];
        if (out.status) *out.status = ok ? 0 : 1;
    }
}

// ComputePoseColoredICPKernelCUDA (RegistrationCUDA.cu:119-180), RegistrationImpl.h:413-493.
template <bool L2LOSS>
__global__ void __launch_bounds__(kThreads)
pose_colored_kernel(const float* __restrict__ src, const float* __restrict__ src_c, const float* __restrict__ tgt,
                    const float* __restrict__ nrm, const float* __restrict__ tgt_c,
                    const float* __restrict__ tgt_g, const int64_t* __restrict__ corr, int64_t n, float sqrt_lg,
                    float sqrt_lp, Robust rk, double* __restrict__ partials, unsigned* ticket, PoseOut out) {
    __shared__ double s_warp[kThreads / 32][kSumStride];
    __shared__ double s_final[kSumStride];
    for (int k = threadIdx.x; k < (kThreads / 32) * kSumStride; k += kThreads) (&s_warp[0][0])[k] = 0.0;
    __syncthreads();
    float acc[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) acc[k] = 0.f;
    int since = 0;
    for (int64_t base = (int64_t)blockIdx.x * kThreads; base < n; base += (int64_t)gridDim.x * kThreads) {
        const int64_t i = base + threadIdx.x;
        if (i < n && corr[i] != -1) {
            const int64_t t = 3 * corr[i], s = 3 * i;
            const float vs[3] = {src[s], src[s + 1], src[s + 2]};
            const float vt[3] = {tgt[t], tgt[t + 1], tgt[t + 2]};
            const float nt[3] = {nrm[t], nrm[t + 1], nrm[t + 2]};
            const float is = color_intensity(src_c[s], src_c[s + 1], src_c[s + 2]);
            const float it = color_intensity(tgt_c[t], tgt_c[t + 1], tgt_c[t + 2]);
            const float dit[3] = {tgt_g[t], tgt_g[t + 1], tgt_g[t + 2]};
            accumulate_colored<L2LOSS>(acc, rk, vs, vt, nt, is, it, dit, sqrt_lg, sqrt_lp);
        }
        if (++since == kFlushEvery) {
            flush_acc(acc, s_warp);
            since = 0;
        }
    }
    flush_acc(acc, s_warp);
    if (!block_reduce_to_global(s_warp, partials, ticket, s_final)) return;
    if (threadIdx.x < 29 && out.sums29) out.sums29[threadIdx.x] = s_final[threadIdx.x];
    if (threadIdx.x == 0) {
        double pose[6] = {0, 0, 0, 0, 0, 0};
        const bool ok = solve6x6(s_final, pose);
        if (!ok)
            for (int k = 0; k < 6; ++k) pose[k] = 0.0;
        if (out.pose)
            for (int k = 0; k < 6; ++k) out.pose[k] = pose[k];
        if (out.status) *out.status = ok ? 0 : 1;
    }
}

// ------------------------------------------------------ transform kernels

__global__ void transform_points_kernel(float* __restrict__ p, int64_t n, Affine T) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    apply_transform(T.m, x, y, z);
    p[3 * i] = x;
    p[3 * i + 1] = y;
    p[3 * i + 2] = z;
}

__global__ void transform_normals_kernel(float* __restrict__ p, int64_t n, Affine T) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = p[3 * i], y = p[3 * i + 1], z = p[3 * i + 2];
    p[3 * i] = T.m[0] * x + T.m[1] * y + T.m[2] * z;
    p[3 * i + 1] = T.m[4] * x + T.m[5] * y + T.m[6] * z;
    p[3 * i + 2] = T.m[8] * x + T.m[9] * y + T.m[10] * z;
}

// ColoredICP side arrays, in the sort order of the working clouds (orig index = .w of the float4)
__global__ void pack_target_color_kernel(const float4* __restrict__ pts4, const float* __restrict__ col,
                                         const float* __restrict__ grad, int64_t m, float4* __restrict__ tcg) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t o = 3 * (int64_t)__float_as_int(pts4[j].w);
    tcg[j] = make_float4(grad[o], grad[o + 1], grad[o + 2], color_intensity(col[o], col[o + 1], col[o + 2]));
}

__global__ void pack_source_intensity_kernel(const float4* __restrict__ src4, const float* __restrict__ col,
                                             int64_t n, float* __restrict__ sint) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = 3 * (int64_t)__float_as_int(src4[i].w);
    sint[i] = color_intensity(col[o], col[o + 1], col[o + 2]);
}

// ------------------------------------------------------------ fused ICP loop

struct IcpState {        // lives in device memory; read back once at the end
    double T[16];        // cumulative source->target transformation (f64)
    double sums[kSumStride];
    double prev_fitness, prev_rmse, fitness, rmse;
    double count;        // correspondences of the last search
    float Uf[16];        // pending update to apply to the working source (f32)
    int iter;            // == reference's iteration_count
    int executed;        // iterations whose (fitness, rmse) were logged
    int done;            // loop left (converged / no correspondences / singular)
    int converged;
    int status;          // 0 ok, 1 singular
    unsigned ticket;
};

struct IcpArgs {
    Grid g;
    const float4* tgt;
    const float4* nrm;
    const unsigned* cs;
    float4* src;          // working source, sorted, .w = original index bits
    int64_t n;            // local source points
    double n_total;       // source points over all ranks (fitness denominator)
    float rr, thr;
    float r1, r1_accept2;   // pass-1 radius of the two-pass search (= cell size) and its acceptance bound
    Robust rk;
    double* partials;
    IcpState* st;
    double* per_iter;
    int64_t* corr_out;    // evaluate mode only
    double rel_fitness, rel_rmse;
    int max_iteration;
    int fuse_finalize;    // 0: leave the totals in st->sums (multi-GPU all-reduce follows)
    // ColoredICP only (null otherwise)
    const float4* tcg;    // per sorted target point: colour gradient xyz, .w = intensity
    const float* sint;    // per sorted source point: intensity
    float sqrt_lg, sqrt_lp;
};

__device__ void set_identity(double* T, float* Uf) {
    for (int i = 0; i < 16; ++i) {
        const double v = (i % 5 == 0) ? 1.0 : 0.0;
        if (T) T[i] = v;
        if (Uf) Uf[i] = (float)v;
    }
}

// Host part of DoSingleScaleICPIterations (Registration.cpp:293-358), on device,
// run by one thread once per iteration.
__device__ void icp_finalize_iteration(const IcpArgs& a, const double* sums) {
    IcpState* st = a.st;
    const double count = sums[28];
    const double fitness = count / a.n_total;                       // Registration.cpp:47-50
    const double rmse = count > 0 ? sqrt(sums[29] / count) : 0.0;
    st->fitness = fitness;
    st->rmse = rmse;
    st->count = count;
    if (!(fitness > DBL_MIN)) {  // :51-60, :300-306 — no correspondences
        set_identity(st->T, st->Uf);
        st->converged = 0;
        st->done = 1;
        return;
    }
    double pose[6];
    if (!solve6x6(sums, pose)) {  // TransformationConverter.cpp:219-225 — the reference raises
        set_identity(nullptr, st->Uf);
        st->status = 1;
        st->done = 1;
        return;
    }
    double U[16], R[16];
    pose_to_T(pose, U);
    for (int i = 0; i < 4; ++i)          // :319  T <- U * T
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += U[i * 4 + k] * st->T[k * 4 + j];
            R[i * 4 + j] = s;
        }
    for (int i = 0; i < 16; ++i) {
        st->T[i] = R[i];
        st->Uf[i] = (float)U[i];          // :322 applied by the next kernel's load
    }
    if (a.per_iter) {
        a.per_iter[2 * st->executed] = fitness;
        a.per_iter[2 * st->executed + 1] = rmse;
    }
    st->executed += 1;
    if (st->iter != 0 && fabs(st->prev_fitness - fitness) < a.rel_fitness &&
        fabs(st->prev_rmse - rmse) < a.rel_rmse) {  // :348-355
        st->converged = 1;
        st->done = 1;
        return;
    }
    st->prev_fitness = fitness;
    st->prev_rmse = rmse;
    st->iter += 1;
    if (st->iter >= a.max_iteration) st->done = 1;
}

// Final ComputeRegistrationResult (Registration.cpp:424-431).
__device__ void icp_finalize_evaluate(const IcpArgs& a, const double* sums) {
    IcpState* st = a.st;
    const double count = sums[28];
    st->count = count;
    if (count > 0) {
        st->fitness = count / a.n_total;
        st->rmse = sqrt(sums[29] / count);
    } else {
        st->fitness = 0.0;
        st->rmse = 0.0;
        set_identity(st->T, nullptr);
        st->converged = 0;
    }
    set_identity(nullptr, st->Uf);
}

// One ICP iteration in ONE kernel: apply the pending update to the working
// source in place, exact 1-NN within the radius through the grid, Jacobian,
// 30-scalar reduction, and (last block) solve + pose update + convergence test.
// MODE 0 = iterate, MODE 1 = evaluate (no Jacobian; writes correspondences).
template <bool L2LOSS, int MODE, bool COLORED = false>
__global__ void __launch_bounds__(kThreads, ICP_MIN_BLOCKS)
icp_iteration_kernel(IcpArgs a) {
    __shared__ double s_warp[kThreads / 32][kSumStride];
    __shared__ double s_final[kSumStride];
    __shared__ float s_U[16];
    __shared__ int s_done;
    if (threadIdx.x == 0) s_done = *(volatile int*)&a.st->done;
    if (threadIdx.x < 16) s_U[threadIdx.x] = a.st->Uf[threadIdx.x];
    for (int k = threadIdx.x; k < (kThreads / 32) * kSumStride; k += kThreads) (&s_warp[0][0])[k] = 0.0;
    __syncthreads();
    if (MODE == 0 && s_done) return;

#if ICP_ACC_SMEM
    // per-thread accumulators live in shared memory (column tid of s_acc): the 30 values are
    // only in registers while one correspondence is being expanded, which leaves the search
    // the whole register budget and lets one more block fit per SM
    __shared__ float s_acc[kNumSums][kThreads];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) s_acc[k][threadIdx.x] = 0.f;
#else
    float acc[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) acc[k] = 0.f;
#endif
    int since = 0;
#if ICP_PREFETCH_SRC
    // experimental (default off, unmeasured): issue the NEXT chunk's source load before this chunk's search so
    // that its latency (10 % of the stall samples) hides behind the search; same values, same results
    float4 p_next = make_float4(0.f, 0.f, 0.f, 0.f);
    {
        const int64_t i0 = (int64_t)blockIdx.x * kThreads + threadIdx.x;
        if (i0 < a.n) p_next = a.src[i0];
    }
#endif
    for (int64_t base = (int64_t)blockIdx.x * kThreads; base < a.n; base += (int64_t)gridDim.x * kThreads) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < a.n;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
#if ICP_PREFETCH_SRC
        p = p_next;
        {
            const int64_t in = i + (int64_t)gridDim.x * kThreads;
            if (in < a.n) p_next = a.src[in];
        }
        if (live) {
            apply_transform(s_U, p.x, p.y, p.z);
            a.src[i] = p;
        }
#else
        if (live) {
            p = a.src[i];
            apply_transform(s_U, p.x, p.y, p.z);   // Registration.cpp:322 (PointCloud::Transform), fused
            a.src[i] = p;
        }
#endif
        Best b;
        b.j = -1;
        if (live) {
            if (kTwoPass) nn_search_two_pass(a.g, a.tgt, a.cs, p.x, p.y, p.z, a.r1, a.r1_accept2, a.rr, a.thr, b);
            else nn_search<true>(a.g, a.tgt, a.cs, p.x, p.y, p.z, a.rr, a.thr, b);
        }
        if (live) {
            if (b.j >= 0) {
#if ICP_ACC_SMEM
                if (MODE == 0) {
                    float term[kNumSums];
#pragma unroll
                    for (int k = 0; k < kNumSums; ++k) term[k] = 0.f;
                    const float4 nn = __ldg(&a.nrm[b.j]);
                    accumulate_p2plane<L2LOSS>(term, a.rk, p.x, p.y, p.z, b.x, b.y, b.z, nn.x, nn.y, nn.z);
                    term[29] = b.d;
#pragma unroll
                    for (int k = 0; k < kNumSums; ++k) s_acc[k][threadIdx.x] += term[k];
                } else {
                    s_acc[28][threadIdx.x] += 1.0f;
                    s_acc[29][threadIdx.x] += b.d;
                }
#else
                if (MODE == 0) {
                    const float4 nn = __ldg(&a.nrm[b.j]);
                    if (COLORED) {
                        const float4 cg = __ldg(&a.tcg[b.j]);
                        const float vs[3] = {p.x, p.y, p.z}, vt[3] = {b.x, b.y, b.z}, nt[3] = {nn.x, nn.y, nn.z};
                        const float dit[3] = {cg.x, cg.y, cg.z};
                        accumulate_colored<L2LOSS>(acc, a.rk, vs, vt, nt, __ldg(&a.sint[i]), cg.w, dit, a.sqrt_lg,
                                                   a.sqrt_lp);
                    } else {
                        accumulate_p2plane<L2LOSS>(acc, a.rk, p.x, p.y, p.z, b.x, b.y, b.z, nn.x, nn.y, nn.z);
                    }
                } else {
                    acc[28] += 1.0f;
                }
                acc[29] += b.d;
#endif
            }
            if (MODE == 1 && a.corr_out) a.corr_out[__float_as_int(p.w)] = b.j >= 0 ? (int64_t)b.idx : (int64_t)-1;
        }
        if (++since == kFlushEvery) {
#if ICP_ACC_SMEM
            flush_acc_smem(s_acc, s_warp);
#else
            flush_acc(acc, s_warp);
#endif
            since = 0;
        }
    }
#if ICP_ACC_SMEM
    flush_acc_smem(s_acc, s_warp);
#else
    flush_acc(acc, s_warp);
#endif
    if (!block_reduce_to_global(s_warp, a.partials, &a.st->ticket, s_final)) return;
    if (a.fuse_finalize) {
        if (threadIdx.x == 0) {
            if (MODE == 0) icp_finalize_iteration(a, s_final);
            else icp_finalize_evaluate(a, s_final);
        }
    } else if (threadIdx.x < kNumSums) {
        a.st->sums[threadIdx.x] = s_final[threadIdx.x];
    }
}


// ------------------------------------------------ TMA-staged tile variant

// Shared-memory tile of one 256-query chunk: the candidate rows of the chunk's cell box are
// brought in with cp.async.bulk (TMA, one bulk copy per non-empty (y,z) row — a row of cells
// is one contiguous run of float4 points) and the matching cell_start segments with coalesced
// loads; the per-query scan then touches shared memory only.
static constexpr int kTilePts = 2560;   // float4 candidates  (40 KB)
static constexpr int kTileCs = 4096;    // staged cell_start entries (16 KB)
static constexpr int kTileRows = 256;   // (y,z) rows of the box (one per thread)

struct __align__(16) TileSmem {
    float4 pts[kTilePts];
    unsigned cs[kTileCs];
    int row_delta[kTileRows];            // smem index minus global index of the row's points
    unsigned long long mbar;
    int bbox[6];
    int ok;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done;
    do {
        asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(b)), "r"(parity)
                : "memory");
    } while (!done);
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// scan_range() over candidates staged in shared memory; `base` is the shared address that
// global index 0 of this row would have (may wrap below the window; only [s, e) is touched).
__device__ __forceinline__ void scan_range_smem(unsigned base, unsigned s, unsigned e, float qx, float qy, float qz,
                                                Best& b) {
    for (unsigned j = s; j < e; j += 4) {
        float4 t[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) t[k] = lds128(base + min(j + k, e - 1) * 16u);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float dx = t[k].x - qx, dy = t[k].y - qy, dz = t[k].z - qz;
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

template <bool L2LOSS, int MODE>
__global__ void __launch_bounds__(kThreads, ICP_MIN_BLOCKS)
icp_iteration_tile_kernel(IcpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileSmem& sm = *reinterpret_cast<TileSmem*>(smem_raw);
    __shared__ double s_warp[kThreads / 32][kSumStride];
    __shared__ double s_final[kSumStride];
    __shared__ float s_U[16];
    __shared__ int s_done;
    if (threadIdx.x == 0) {
        s_done = *(volatile int*)&a.st->done;
        mbar_init(&sm.mbar, 1);
    }
    if (threadIdx.x < 16) s_U[threadIdx.x] = a.st->Uf[threadIdx.x];
    for (int k = threadIdx.x; k < (kThreads / 32) * kSumStride; k += kThreads) (&s_warp[0][0])[k] = 0.0;
    __syncthreads();
    if (MODE == 0 && s_done) return;

    const Grid& g = a.g;
    const int kBig = 0x3fffffff;
    unsigned phase = 0;
    float acc[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) acc[k] = 0.f;
    int since = 0;
    for (int64_t base = (int64_t)blockIdx.x * kThreads; base < a.n; base += (int64_t)gridDim.x * kThreads) {
        const int64_t i = base + threadIdx.x;
        const bool live = i < a.n;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (live) {
            p = a.src[i];
            apply_transform(s_U, p.x, p.y, p.z);   // Registration.cpp:322 (PointCloud::Transform), fused
            a.src[i] = p;
        }
        // ---- this query's pass-1 cell box and the block's union of them
        float gx, gy, gz;   // the query in grid axis order
        to_grid(g, p.x, p.y, p.z, gx, gy, gz);
        const bool inside = live && !(hi_bound(gx, a.rr) < g.bmin[0] || lo_bound(gx, a.rr) > g.bmax[0] ||
                                      hi_bound(gy, a.rr) < g.bmin[1] || lo_bound(gy, a.rr) > g.bmax[1] ||
                                      hi_bound(gz, a.rr) < g.bmin[2] || lo_bound(gz, a.rr) > g.bmax[2] ||
                                      !(p.x == p.x) || !(p.y == p.y) || !(p.z == p.z));
        int x0 = kBig, x1 = -kBig, y0 = kBig, y1 = -kBig, z0 = kBig, z1 = -kBig;
        if (inside) {
            x0 = cell1(lo_bound(gx, a.r1), g.ox, g.inv_cx, g.nx);
            x1 = cell1(hi_bound(gx, a.r1), g.ox, g.inv_cx, g.nx);
            y0 = cell1(lo_bound(gy, a.r1), g.oy, g.inv_c, g.ny);
            y1 = cell1(hi_bound(gy, a.r1), g.oy, g.inv_c, g.ny);
            z0 = cell1(lo_bound(gz, a.r1), g.oz, g.inv_c, g.nz);
            z1 = cell1(hi_bound(gz, a.r1), g.oz, g.inv_c, g.nz);
        }
        if (threadIdx.x < 6) sm.bbox[threadIdx.x] = (threadIdx.x & 1) ? -kBig : kBig;
        __syncthreads();
        {
            int mn[3] = {x0, y0, z0}, mx[3] = {x1, y1, z1};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    mn[k] = min(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
                    mx[k] = max(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
                }
            }
            if ((threadIdx.x & 31) == 0) {
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    atomicMin(&sm.bbox[2 * k], mn[k]);
                    atomicMax(&sm.bbox[2 * k + 1], mx[k]);
                }
            }
        }
        __syncthreads();
        const int BX0 = sm.bbox[0], BX1 = sm.bbox[1], BY0 = sm.bbox[2], BY1 = sm.bbox[3], BZ0 = sm.bbox[4],
                  BZ1 = sm.bbox[5];
        const int ty = BY1 - BY0 + 1, tz = BZ1 - BZ0 + 1, W = BX1 - BX0 + 2;   // W entries per row
        const int rows = ty * tz;
        bool tile_ok = BX0 <= BX1 && rows <= kTileRows && (int64_t)rows * W <= kTileCs;
        // ---- stage the cell_start segments of every row of the box (coalesced)
        if (tile_ok) {
            for (int e = threadIdx.x; e < rows * W; e += kThreads) {
                const int r = e / W, k = e - r * W;
                const int iy = BY0 + r % ty, iz = BZ0 + r / ty;
                sm.cs[e] = __ldg(&a.cs[(iz * g.ny + iy) * g.nx + BX0 + k]);
            }
        }
        __syncthreads();
        // ---- per-row candidate runs -> shared offsets; one bulk copy per non-empty row
        unsigned cnt = 0, gstart = 0;
        if (tile_ok && threadIdx.x < rows) {
            gstart = sm.cs[threadIdx.x * W];
            cnt = sm.cs[threadIdx.x * W + W - 1] - gstart;
        }
        unsigned total;
        const unsigned off = block_exclusive_scan(cnt, total);
        tile_ok = tile_ok && total <= (unsigned)kTilePts;
        if (tile_ok) {
            if (threadIdx.x < rows) sm.row_delta[threadIdx.x] = (int)off - (int)gstart;
            if (threadIdx.x == 0) mbar_expect_tx(&sm.mbar, total * 16u);
        }
        __syncthreads();
        if (tile_ok) {
            if (cnt > 0) bulk_g2s(&sm.pts[off], a.tgt + gstart, cnt * 16u, &sm.mbar);
            mbar_wait(&sm.mbar, phase);
            phase ^= 1u;
        }
        // ---- search
        Best b;
        b.d = a.thr;
        b.j = -1;
        b.idx = 0x7fffffff;
        b.x = b.y = b.z = 0.f;
        if (inside) {
            bool need_global = !tile_ok;
            if (tile_ok) {
                const unsigned pts_base = smem_u32(sm.pts);
                for (int iz = z0; iz <= z1; ++iz)
                    for (int iy = y0; iy <= y1; ++iy) {
                        const int r = (iz - BZ0) * ty + (iy - BY0);
                        const unsigned s = sm.cs[r * W + (x0 - BX0)], e = sm.cs[r * W + (x1 - BX0) + 1];
                        scan_range_smem(pts_base + (unsigned)sm.row_delta[r] * 16u, s, e, p.x, p.y, p.z, b);
                    }
                need_global = !(b.j >= 0 && b.d <= a.r1_accept2) && a.r1 < a.rr;
            }
            if (need_global) nn_search_two_pass(g, a.tgt, a.cs, p.x, p.y, p.z, a.r1, a.r1_accept2, a.rr, a.thr, b);
        }
        if (live) {
            if (b.j >= 0) {
                if (MODE == 0) {
                    const float4 nn = __ldg(&a.nrm[b.j]);
                    accumulate_p2plane<L2LOSS>(acc, a.rk, p.x, p.y, p.z, b.x, b.y, b.z, nn.x, nn.y, nn.z);
                } else {
                    acc[28] += 1.0f;
                }
                acc[29] += b.d;
            }
            if (MODE == 1 && a.corr_out) a.corr_out[__float_as_int(p.w)] = b.j >= 0 ? (int64_t)b.idx : (int64_t)-1;
        }
        if (++since == kFlushEvery) {
            flush_acc(acc, s_warp);
            since = 0;
        }
        __syncthreads();   // everyone is done with the tile before the next chunk overwrites it
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    flush_acc(acc, s_warp);
    if (!block_reduce_to_global(s_warp, a.partials, &a.st->ticket, s_final)) return;
    if (a.fuse_finalize) {
        if (threadIdx.x == 0) {
            if (MODE == 0) icp_finalize_iteration(a, s_final);
            else icp_finalize_evaluate(a, s_final);
        }
    } else if (threadIdx.x < kNumSums) {
        a.st->sums[threadIdx.x] = s_final[threadIdx.x];
    }
}

// Multi-GPU: runs after the all-reduce of st->sums.
template <int MODE>
__global__ void icp_finalize_kernel(IcpArgs a) {
    if (threadIdx.x != 0) return;
    if (MODE == 0) {
        if (a.st->done) return;
        icp_finalize_iteration(a, a.st->sums);
    } else {
        icp_finalize_evaluate(a, a.st->sums);
    }
}

// Gather the caller's source into the sorted working copy (clone + initial transform).
// (scatter_kernel<true> does the work; this is the reset path reusing key/rank.)

}  // namespace o3db

struct o3db_icp {
    o3db_nns nns;
    o3db_icp_options opt{};
    const float* src_user = nullptr;
    int64_t n = 0;
    double n_total = 0;
    double init_T[16];
    float4* src4 = nullptr;
    unsigned* src_key = nullptr;     // cell key of every source point (sort order)
    unsigned* src_rank = nullptr;
    unsigned* src_start = nullptr;   // CSR offsets of the source sort
    double* partials = nullptr;
    double* per_iter = nullptr;
    IcpState* st = nullptr;
    IcpState* h_st = nullptr;        // pinned
    o3db_comm* comm = nullptr;
    int grid_blocks = 0;
    int variant = 2;                 // 1 = direct global search, 2 = TMA-staged tiles
    int64_t src_keys = 0;            // size of the source sort key space
    int launched = 0;
    bool l2loss = true;
    // ColoredICP (TransformationEstimationForColoredICP); null / unused for point-to-plane
    bool colored = false;
    const float* src_colors_user = nullptr;
    float4* tcg4 = nullptr;          // sorted target: colour gradient xyz + intensity
    float* sint = nullptr;           // sorted source: intensity
    double lambda_geometric = 0.968;
};

namespace o3db {

static IcpArgs make_args(o3db_icp* c) {
    IcpArgs a{};
    a.g = c->nns.g;
    a.tgt = c->nns.pts4;
    a.nrm = c->nns.nrm4;
    a.cs = c->nns.cell_start;
    a.src = c->src4;
    a.n = c->n;
    a.n_total = c->n_total;
    const float r = (float)c->opt.max_correspondence_distance;
    a.thr = r * r;                       // FixedRadiusSearchImpl.cuh:692: T(radius) * T(radius)
    a.rr = r * (1.0f + 1e-6f);
    a.r1 = fminf(c->nns.g.c, a.rr);
    a.r1_accept2 = (a.r1 * (1.0f - 1e-4f)) * (a.r1 * (1.0f - 1e-4f));
    a.rk.method = c->opt.kernel.method;
    a.rk.scale = (float)c->opt.kernel.scale;
    a.rk.shape = c->opt.kernel.shape;
    a.partials = c->partials;
    a.st = c->st;
    a.per_iter = c->per_iter;
    a.corr_out = nullptr;
    a.rel_fitness = c->opt.relative_fitness;
    a.rel_rmse = c->opt.relative_rmse;
    a.max_iteration = c->opt.max_iteration;
    a.fuse_finalize = c->comm ? 0 : 1;
    a.tcg = c->tcg4;
    a.sint = c->sint;
    a.sqrt_lg = (float)sqrt(c->lambda_geometric);          // RegistrationCUDA.cu:205-208
    a.sqrt_lp = (float)sqrt(1.0 - c->lambda_geometric);
    return a;
}

static int icp_init_state(o3db_icp* c, cudaStream_t st) {
    IcpState h{};
    for (int i = 0; i < 16; ++i) {
        h.T[i] = c->init_T[i];
        h.Uf[i] = (i % 5 == 0) ? 1.f : 0.f;   // the initial transform is applied by the gather
    }
    memcpy(c->h_st, &h, sizeof(h));
    O3DB_CUDA_CHECK(cudaMemcpyAsync(c->st, c->h_st, sizeof(IcpState), cudaMemcpyHostToDevice, st));
    c->launched = 0;
    return O3DB_OK;
}

static int icp_gather_source(o3db_icp* c, cudaStream_t st) {
    Affine T0;
    for (int i = 0; i < 16; ++i) T0.m[i] = (float)c->init_T[i];   // Transform.cpp:29-31: T cast to the point dtype
    if (c->n == 0) return O3DB_OK;
    scatter_kernel<true><<<(unsigned)ceil_div(c->n, kThreads), kThreads, 0, st>>>(
            c->src_user, nullptr, c->n, T0, c->src_start, c->src_key, c->src_rank, c->src4, nullptr);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

}  // namespace o3db

extern "C" {

int o3db_nns_create(const float* points_dev, int64_t num_points, double radius, void* stream, o3db_nns** out) {
    O3DB_REQUIRE(out != nullptr, "o3db_nns_create: out is null");
    *out = nullptr;
    O3DB_REQUIRE(points_dev != nullptr && num_points > 0, "o3db_nns_create: empty point set");
    O3DB_REQUIRE(num_points < INT_MAX, "o3db_nns_create: too many points");
    O3DB_REQUIRE(radius > 0 && std::isfinite(radius), "o3db_nns_create: radius must be positive");
    o3db_nns* s = new (std::nothrow) o3db_nns();
    O3DB_REQUIRE(s != nullptr, "out of host memory");
    const int rc = nns_build(s, points_dev, nullptr, num_points, radius, 1.0, (cudaStream_t)stream);
    if (rc != O3DB_OK) {
        nns_free(s, (cudaStream_t)stream);
        delete s;
        return rc;
    }
    *out = s;
    return O3DB_OK;
}

void o3db_nns_destroy(o3db_nns* nns) {
    if (!nns) return;
    nns_free(nns, 0);
    delete nns;
}

int o3db_nns_hybrid_search(const o3db_nns* nns, const float* queries_dev, int64_t num_queries, double radius,
                           int max_knn, int32_t* indices_dev, float* distances_dev, int32_t* counts_dev,
                           void* stream) {
    O3DB_REQUIRE(nns != nullptr, "o3db_nns_hybrid_search: null index");
    O3DB_REQUIRE(max_knn >= 1 && max_knn <= kMaxKnn, "o3db_nns_hybrid_search: max_knn must be in 1..%d", kMaxKnn);
    O3DB_REQUIRE(radius > 0 && radius <= nns->radius * (1 + 1e-12),
                 "o3db_nns_hybrid_search: radius %g exceeds the index radius %g", radius, nns->radius);
    if (num_queries == 0) return O3DB_OK;
    O3DB_REQUIRE(queries_dev != nullptr && num_queries > 0, "o3db_nns_hybrid_search: bad queries");
    const float r = (float)radius;
    const float thr = r * r, rr = r * (1.0f + 1e-6f);
    const unsigned nb = (unsigned)ceil_div(num_queries, kThreads);
    cudaStream_t st = (cudaStream_t)stream;
    if (max_knn == 1)
        hybrid_search_k1_kernel<true><<<nb, kThreads, 0, st>>>(nns->g, nns->pts4, nns->cell_start, queries_dev,
                                                               num_queries, rr, thr, indices_dev, distances_dev,
                                                               counts_dev);
    else
        hybrid_search_knn_kernel<<<nb, kThreads, 0, st>>>(nns->g, nns->pts4, nns->cell_start, queries_dev,
                                                          num_queries, rr, thr, max_knn, indices_dev,
                                                          distances_dev, counts_dev);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

int o3db_build_spatial_hash_table(const float* points_dev, int64_t num_points, double radius,
                                  uint32_t hash_table_size, uint32_t* hash_table_index_dev,
                                  uint32_t* hash_table_cell_splits_dev, void* stream) {
    O3DB_REQUIRE(points_dev && hash_table_index_dev && hash_table_cell_splits_dev && num_points > 0 &&
                         num_points < INT_MAX && hash_table_size > 0 && radius > 0,
                 "o3db_build_spatial_hash_table: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const float voxel = 2 * (float)radius;          // FixedRadiusSearchImpl.cuh:760-761 (T arithmetic)
    const float inv_voxel = 1 / voxel;
    unsigned *rank = nullptr, *scratch = nullptr;
    O3DB_CUDA_CHECK(cudaMallocAsync(&rank, num_points * sizeof(unsigned), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&scratch, ceil_div(hash_table_size, kScanTile) * sizeof(unsigned), st));
    O3DB_CUDA_CHECK(cudaMemsetAsync(hash_table_cell_splits_dev, 0, ((size_t)hash_table_size + 1) * sizeof(unsigned), st));
    const unsigned nb = (unsigned)ceil_div(num_points, kThreads);
    ref_count_kernel<<<nb, kThreads, 0, st>>>(points_dev, num_points, inv_voxel, hash_table_size,
                                              hash_table_cell_splits_dev, rank);
    O3DB_LAUNCH_CHECK();
    int rc = exclusive_scan_u32(hash_table_cell_splits_dev, hash_table_size, scratch, st);
    if (rc) return rc;
    ref_scatter_kernel<<<nb, kThreads, 0, st>>>(points_dev, num_points, inv_voxel, hash_table_size,
                                                hash_table_cell_splits_dev, rank, hash_table_index_dev);
    O3DB_LAUNCH_CHECK();
    O3DB_CUDA_CHECK(cudaFreeAsync(rank, st));
    O3DB_CUDA_CHECK(cudaFreeAsync(scratch, st));
    return O3DB_OK;
}

void o3db_pose_to_transformation(const double pose_host[6], double transformation_host[16]) {
    pose_to_T(pose_host, transformation_host);
}

int o3db_transform_points(const double T[16], float* points_dev, int64_t n, void* stream) {
    O3DB_REQUIRE(T != nullptr && (points_dev != nullptr || n == 0) && n >= 0, "o3db_transform_points: bad arguments");
    if (n == 0) return O3DB_OK;
    Affine A;
    for (int i = 0; i < 16; ++i) A.m[i] = (float)T[i];
    transform_points_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(points_dev, n, A);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

int o3db_transform_normals(const double T[16], float* normals_dev, int64_t n, void* stream) {
    O3DB_REQUIRE(T != nullptr && (normals_dev != nullptr || n == 0) && n >= 0, "o3db_transform_normals: bad arguments");
    if (n == 0) return O3DB_OK;
    Affine A;
    for (int i = 0; i < 16; ++i) A.m[i] = (float)T[i];
    transform_normals_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, (cudaStream_t)stream>>>(normals_dev, n, A);
    O3DB_LAUNCH_CHECK();
    return O3DB_OK;
}

}  // extern "C"

namespace o3db {

struct PoseScratch {
    double* partials = nullptr;
    unsigned* ticket = nullptr;
    int* status = nullptr;
    double* sums = nullptr;
    int blocks = 0;
};

static int pose_scratch_alloc(PoseScratch* s, int64_t n, cudaStream_t st) {
    s->blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, kThreads), (int64_t)num_sms() * 4));
    char* base = nullptr;
    const size_t bytes = (size_t)s->blocks * kSumStride * sizeof(double) + kSumStride * sizeof(double) + 64;
    O3DB_CUDA_CHECK(cudaMallocAsync(&base, bytes, st));
    O3DB_CUDA_CHECK(cudaMemsetAsync(base, 0, bytes, st));
    s->partials = (double*)base;
    s->sums = s->partials + (size_t)s->blocks * kSumStride;
    s->ticket = (unsigned*)(s->sums + kSumStride);
    s->status = (int*)(s->ticket + 4);
    return O3DB_OK;
}

static int pose_finish(PoseScratch* s, double* sums29_dev, float* residual_host, int* inlier_count_host,
                       cudaStream_t st) {
    int rc = O3DB_OK;
    if (sums29_dev)
        O3DB_CUDA_CHECK(cudaMemcpyAsync(sums29_dev, s->sums, 29 * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (residual_host || inlier_count_host) {
        double h[kSumStride];
        int status = 0;
        O3DB_CUDA_CHECK(cudaMemcpyAsync(h, s->sums, sizeof(h), cudaMemcpyDeviceToHost, st));
        O3DB_CUDA_CHECK(cudaMemcpyAsync(&status, s->status, sizeof(int), cudaMemcpyDeviceToHost, st));
        O3DB_CUDA_CHECK(cudaStreamSynchronize(st));
        if (status) {  // TransformationConverter.cpp:219-225
            set_last_error("Singular 6x6 linear system detected, tracking failed.");
            if (residual_host) *residual_host = 0;
            if (inlier_count_host) *inlier_count_host = 0;
            rc = O3DB_ERR_SINGULAR;
        } else {
            if (residual_host) *residual_host = (float)h[27];
            if (inlier_count_host) *inlier_count_host = (int)h[28];
        }
    }
    O3DB_CUDA_CHECK(cudaFreeAsync(s->partials, st));
    return rc;
}

}  // namespace o3db

extern "C" {

int o3db_compute_pose_point_to_plane(const float* source_dev, const float* target_dev,
                                     const float* target_normals_dev, const int64_t* correspondences_dev,
                                     int64_t n, const o3db_robust_kernel* kernel, double* sums29_dev,
                                     double* pose_dev, float* residual_host, int* inlier_count_host,
                                     void* stream) {
    O3DB_REQUIRE(n >= 0 && (n == 0 || (source_dev && target_dev && target_normals_dev && correspondences_dev)),
                 "o3db_compute_pose_point_to_plane: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PoseScratch s;
    int rc = pose_scratch_alloc(&s, n, st);
    if (rc) return rc;
    Robust rk{kernel ? kernel->method : 0, kernel ? (float)kernel->scale : 1.f, kernel ? kernel->shape : 1.0};
    PoseOut out{s.sums, pose_dev, s.status};
    if (rk.method == O3DB_ROBUST_L2)
        pose_p2plane_kernel<true><<<s.blocks, kThreads, 0, st>>>(source_dev, target_dev, target_normals_dev,
                                                                 correspondences_dev, n, rk, s.partials, s.ticket, out);
    else
        pose_p2plane_kernel<false><<<s.blocks, kThreads, 0, st>>>(source_dev, target_dev, target_normals_dev,
                                                                  correspondences_dev, n, rk, s.partials, s.ticket, out);
    O3DB_LAUNCH_CHECK();
    return pose_finish(&s, sums29_dev, residual_host, inlier_count_host, st);
}

int o3db_compute_pose_colored_icp(const float* source_dev, const float* source_colors_dev, const float* target_dev,
                                  const float* target_normals_dev, const float* target_colors_dev,
                                  const float* target_color_gradients_dev, const int64_t* correspondences_dev,
                                  int64_t n, const o3db_robust_kernel* kernel, double lambda_geometric,
                                  double* sums29_dev, double* pose_dev, float* residual_host,
                                  int* inlier_count_host, void* stream) {
    O3DB_REQUIRE(n >= 0 && (n == 0 || (source_dev && source_colors_dev && target_dev && target_normals_dev &&
                                       target_colors_dev && target_color_gradients_dev && correspondences_dev)),
                 "o3db_compute_pose_colored_icp: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    PoseScratch s;
    int rc = pose_scratch_alloc(&s, n, st);
    if (rc) return rc;
    Robust rk{kernel ? kernel->method : 0, kernel ? (float)kernel->scale : 1.f, kernel ? kernel->shape : 1.0};
    PoseOut out{s.sums, pose_dev, s.status};
    const float sl = (float)sqrt(lambda_geometric), sp = (float)sqrt(1.0 - lambda_geometric);  // RegistrationCUDA.cu:205-208
    if (rk.method == O3DB_ROBUST_L2)
        pose_colored_kernel<true><<<s.blocks, kThreads, 0, st>>>(
                source_dev, source_colors_dev, target_dev, target_normals_dev, target_colors_dev,
                target_color_gradients_dev, correspondences_dev, n, sl, sp, rk, s.partials, s.ticket, out);
    else
        pose_colored_kernel<false><<<s.blocks, kThreads, 0, st>>>(
                source_dev, source_colors_dev, target_dev, target_normals_dev, target_colors_dev,
                target_color_gradients_dev, correspondences_dev, n, sl, sp, rk, s.partials, s.ticket, out);
    O3DB_LAUNCH_CHECK();
    return pose_finish(&s, sums29_dev, residual_host, inlier_count_host, st);
}

// ------------------------------------------------------------- fused ICP API

void o3db_icp_destroy(o3db_icp* c) {
    if (!c) return;
    nns_free(&c->nns, 0);
    if (c->src4) cudaFreeAsync(c->src4, 0);
    if (c->src_key) cudaFreeAsync(c->src_key, 0);
    if (c->src_rank) cudaFreeAsync(c->src_rank, 0);
    if (c->src_start) cudaFreeAsync(c->src_start, 0);
    if (c->partials) cudaFreeAsync(c->partials, 0);
    if (c->per_iter) cudaFreeAsync(c->per_iter, 0);
    if (c->st) cudaFreeAsync(c->st, 0);
    if (c->tcg4) cudaFreeAsync(c->tcg4, 0);
    if (c->sint) cudaFreeAsync(c->sint, 0);
    if (c->h_st) pinned_release(c->h_st);
    delete c;
}

}  // extern "C"

namespace o3db {
struct ColoredInputs {   // all device pointers; null source_colors = plain point-to-plane
    const float* source_colors = nullptr;
    const float* target_colors = nullptr;
    const float* target_color_gradients = nullptr;
    double lambda_geometric = 0.968;
};
}  // namespace o3db

static int icp_create_impl(const float* source_dev, int64_t n, const float* target_dev, const float* target_normals_dev,
                           int64_t m, const double init_T[16], const o3db_icp_options* options, o3db_comm* comm,
                           const o3db::ColoredInputs& col, void* stream, o3db_icp** out) {
    O3DB_REQUIRE(out != nullptr, "o3db_icp_create: out is null");
    *out = nullptr;
    O3DB_REQUIRE(options != nullptr && init_T != nullptr, "o3db_icp_create: null options / init");
    // Registration.cpp:119-219 AssertInputMultiScaleICP
    O3DB_REQUIRE(source_dev && target_dev && n > 0 && m > 0, "Source and/or Target pointcloud is empty.");
    O3DB_REQUIRE(target_normals_dev != nullptr, "Target pointcloud missing normals attribute.");
    O3DB_REQUIRE(n < INT_MAX && m < INT_MAX, "o3db_icp_create: too many points");
    O3DB_REQUIRE(options->max_correspondence_distance > 0, "max_correspondence_distance must be positive");
    O3DB_REQUIRE(options->max_iteration >= 0, "max_iteration must be non-negative");
    cudaStream_t st = (cudaStream_t)stream;
    o3db_icp* c = new (std::nothrow) o3db_icp();
    O3DB_REQUIRE(c != nullptr, "out of host memory");
    c->opt = *options;
    c->src_user = source_dev;
    c->n = n;
    c->comm = comm;
    c->l2loss = options->kernel.method == O3DB_ROBUST_L2;
    c->colored = col.source_colors != nullptr;
    c->src_colors_user = col.source_colors;
    c->lambda_geometric = col.lambda_geometric;
    memcpy(c->init_T, init_T, sizeof(c->init_T));
    // fine cells (half the radius) by default: pass 1 of the two-pass search then covers the
    // +-1 cell box, which holds the nearest neighbour of every already roughly aligned point
    int rc = nns_build(&c->nns, target_dev, target_normals_dev, m, options->max_correspondence_distance,
                       options->cell_scale > 0 ? options->cell_scale : kDefaultCellScale, st);
    if (rc) {
        o3db_icp_destroy(c);
        return rc;
    }
#define ICP_TRY(expr)                         \
    do {                                      \
        int rc__ = (expr);                    \
        if (rc__ != O3DB_OK) {                \
            o3db_icp_destroy(c);              \
            return rc__;                      \
        }                                     \
    } while (0)
#define ICP_CUDA(expr)                                                                           \
    do {                                                                                         \
        cudaError_t e__ = (expr);                                                                \
        if (e__ != cudaSuccess) {                                                                \
            set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
            o3db_icp_destroy(c);                                                                 \
            return O3DB_ERR_CUDA;                                                                \
        }                                                                                        \
    } while (0)
    // fitness denominator over all ranks
    c->n_total = (double)n;
    // 1 = direct (two-pass search straight from global/L1), 2 = TMA-staged shared-memory tiles.
    // Default: whichever measured faster on B200 — currently the direct kernel.
    c->variant = options->search_variant == 2 ? 2 : (options->search_variant == 1 ? 1 : ICP_DEFAULT_VARIANT);
    if (c->colored) c->variant = 1;   // the tile-staged kernel has no coloured instantiation
    int occ = 1;
    if (c->colored) {
        if (c->l2loss)
            ICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_iteration_kernel<true, 0, true>, kThreads, 0));
        else
            ICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_iteration_kernel<false, 0, true>, kThreads, 0));
    } else if (c->variant == 2) {
        const int smem = (int)sizeof(TileSmem);
        ICP_CUDA(cudaFuncSetAttribute(icp_iteration_tile_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ICP_CUDA(cudaFuncSetAttribute(icp_iteration_tile_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        ICP_CUDA(cudaFuncSetAttribute(icp_iteration_tile_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        if (c->l2loss)
            ICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_iteration_tile_kernel<true, 0>, kThreads, smem));
        else
            ICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_iteration_tile_kernel<false, 0>, kThreads, smem));
    } else if (c->l2loss) {
        ICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_iteration_kernel<true, 0>, kThreads, 0));
    } else {
        ICP_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, icp_iteration_kernel<false, 0>, kThreads, 0));
    }
    c->grid_blocks = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(n, kThreads), (int64_t)num_sms() * std::max(occ, 1)));
    const int64_t ncell = tiled_key_space(c->nns.g.nx, c->nns.g.ny, c->nns.g.nz);   // tile-major source keys
    c->src_keys = ncell;
    ICP_CUDA(cudaMallocAsync(&c->src4, n * sizeof(float4), st));
    ICP_CUDA(cudaMallocAsync(&c->src_key, n * sizeof(unsigned), st));
    ICP_CUDA(cudaMallocAsync(&c->src_rank, n * sizeof(unsigned), st));
    ICP_CUDA(cudaMallocAsync(&c->src_start, (ncell + 1) * sizeof(unsigned), st));
    ICP_CUDA(cudaMallocAsync(&c->partials, (size_t)c->grid_blocks * kSumStride * sizeof(double), st));
    ICP_CUDA(cudaMallocAsync(&c->per_iter, (size_t)std::max(1, options->max_iteration) * 2 * sizeof(double), st));
    ICP_CUDA(cudaMallocAsync(&c->st, sizeof(IcpState), st));
    static_assert(sizeof(IcpState) <= 4096, "IcpState must fit a pinned block");
    c->h_st = (IcpState*)pinned_acquire(sizeof(IcpState));
    if (!c->h_st) {
        set_last_error("pinned host allocation failed");
        o3db_icp_destroy(c);
        return O3DB_ERR_CUDA;
    }
    ICP_CUDA(cudaMemsetAsync(c->partials, 0, (size_t)c->grid_blocks * kSumStride * sizeof(double), st));
    ICP_CUDA(cudaMemsetAsync(c->per_iter, 0, (size_t)std::max(1, options->max_iteration) * 2 * sizeof(double), st));
    // sort order of the source: target-grid cell of the initially transformed point
    unsigned* scratch = nullptr;
    ICP_CUDA(cudaMallocAsync(&scratch, ceil_div(ncell, kScanTile) * sizeof(unsigned), st));
    ICP_CUDA(cudaMemsetAsync(c->src_start, 0, (ncell + 1) * sizeof(unsigned), st));
    Affine T0;
    for (int i = 0; i < 16; ++i) T0.m[i] = (float)init_T[i];
    count_kernel<true><<<(unsigned)ceil_div(n, kThreads), kThreads, 0, st>>>(source_dev, n, c->nns.g, T0, c->src_start,
                                                                            c->src_key, c->src_rank);
    count_launch();
    ICP_CUDA(cudaGetLastError());
    ICP_TRY(exclusive_scan_u32(c->src_start, ncell, scratch, st));
    ICP_CUDA(cudaFreeAsync(scratch, st));
    ICP_TRY(icp_gather_source(c, st));
    if (c->colored) {
        ICP_CUDA(cudaMallocAsync(&c->tcg4, m * sizeof(float4), st));
        ICP_CUDA(cudaMallocAsync(&c->sint, n * sizeof(float), st));
        pack_target_color_kernel<<<(unsigned)ceil_div(m, kThreads), kThreads, 0, st>>>(
                c->nns.pts4, col.target_colors, col.target_color_gradients, m, c->tcg4);
        count_launch();
        ICP_CUDA(cudaGetLastError());
        // the source sort order is fixed at creation (o3db_icp_reset re-gathers into the same slots)
        pack_source_intensity_kernel<<<(unsigned)ceil_div(n, kThreads), kThreads, 0, st>>>(c->src4, col.source_colors, n,
                                                                                          c->sint);
        count_launch();
        ICP_CUDA(cudaGetLastError());
    }
    ICP_TRY(icp_init_state(c, st));
    if (comm) {
        double* d = nullptr;
        ICP_CUDA(cudaMallocAsync(&d, sizeof(double), st));
        ICP_CUDA(cudaMemcpyAsync(d, &c->n_total, sizeof(double), cudaMemcpyHostToDevice, st));
        ICP_TRY(o3db_comm_allreduce_f64(comm, d, 1, st));
        ICP_CUDA(cudaMemcpyAsync(&c->n_total, d, sizeof(double), cudaMemcpyDeviceToHost, st));
        ICP_CUDA(cudaStreamSynchronize(st));
        ICP_CUDA(cudaFreeAsync(d, st));
    }
    *out = c;
    return O3DB_OK;
}

extern "C" {

int o3db_icp_create(const float* source_dev, int64_t n, const float* target_dev, const float* target_normals_dev,
                    int64_t m, const double init_T[16], const o3db_icp_options* options, o3db_comm* comm,
                    void* stream, o3db_icp** out) {
    return icp_create_impl(source_dev, n, target_dev, target_normals_dev, m, init_T, options, comm,
                           o3db::ColoredInputs{}, stream, out);
}

int o3db_icp_create_colored(const float* source_dev, const float* source_colors_dev, int64_t n,
                            const float* target_dev, const float* target_normals_dev, const float* target_colors_dev,
                            const float* target_color_gradients_dev, int64_t m, const double init_T[16],
                            const o3db_icp_options* options, double lambda_geometric, o3db_comm* comm, void* stream,
                            o3db_icp** out) {
    O3DB_REQUIRE(out != nullptr, "o3db_icp_create_colored: out is null");
    *out = nullptr;
    // ColoredICP.. TransformationEstimationForColoredICP::ComputeTransformation (TransformationEstimation.cpp:226-262)
    O3DB_REQUIRE(source_colors_dev != nullptr, "Source pointcloud missing colors attribute.");
    O3DB_REQUIRE(target_colors_dev != nullptr, "Target pointcloud missing colors attribute.");
    O3DB_REQUIRE(target_color_gradients_dev != nullptr,
                 "Target pointcloud missing color_gradients attribute (o3db_estimate_color_gradients).");
    O3DB_REQUIRE(lambda_geometric >= 0.0 && lambda_geometric <= 1.0, "lambda_geometric must be in [0, 1]");
    o3db::ColoredInputs col;
    col.source_colors = source_colors_dev;
    col.target_colors = target_colors_dev;
    col.target_color_gradients = target_color_gradients_dev;
    col.lambda_geometric = lambda_geometric;
    return icp_create_impl(source_dev, n, target_dev, target_normals_dev, m, init_T, options, comm, col, stream, out);
}

int o3db_icp_reset(o3db_icp* c, void* stream) {
    O3DB_REQUIRE(c != nullptr, "o3db_icp_reset: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    int rc = icp_gather_source(c, st);
    if (rc) return rc;
    return icp_init_state(c, st);
}

int o3db_icp_iterate(o3db_icp* c, int iterations, void* stream) {
    O3DB_REQUIRE(c != nullptr, "o3db_icp_iterate: null handle");
    cudaStream_t st = (cudaStream_t)stream;
    const int todo = std::min(iterations, c->opt.max_iteration - c->launched);
    IcpArgs a = make_args(c);
    for (int k = 0; k < todo; ++k) {
        if (c->variant == 2) {
            if (c->l2loss) icp_iteration_tile_kernel<true, 0><<<c->grid_blocks, kThreads, sizeof(TileSmem), st>>>(a);
            else icp_iteration_tile_kernel<false, 0><<<c->grid_blocks, kThreads, sizeof(TileSmem), st>>>(a);
        } else if (c->colored) {
            if (c->l2loss) icp_iteration_kernel<true, 0, true><<<c->grid_blocks, kThreads, 0, st>>>(a);
            else icp_iteration_kernel<false, 0, true><<<c->grid_blocks, kThreads, 0, st>>>(a);
        } else {
            if (c->l2loss) icp_iteration_kernel<true, 0><<<c->grid_blocks, kThreads, 0, st>>>(a);
            else icp_iteration_kernel<false, 0><<<c->grid_blocks, kThreads, 0, st>>>(a);
        }
        O3DB_LAUNCH_CHECK();
        if (c->comm) {
            int rc = o3db_comm_allreduce_f64(c->comm, (double*)((char*)c->st + offsetof(IcpState, sums)), kNumSums, st);
            if (rc) return rc;
            icp_finalize_kernel<0><<<1, 32, 0, st>>>(a);
            O3DB_LAUNCH_CHECK();
        }
        c->launched += 1;
    }
    return O3DB_OK;
}

int o3db_icp_finish(o3db_icp* c, o3db_icp_result* result, int64_t* correspondences_dev, double* per_iteration_host,
                    void* stream) {
    O3DB_REQUIRE(c != nullptr && result != nullptr, "o3db_icp_finish: null argument");
    cudaStream_t st = (cudaStream_t)stream;
    IcpArgs a = make_args(c);
    a.corr_out = correspondences_dev;
    if (c->variant == 2) icp_iteration_tile_kernel<true, 1><<<c->grid_blocks, kThreads, sizeof(TileSmem), st>>>(a);
    else icp_iteration_kernel<true, 1><<<c->grid_blocks, kThreads, 0, st>>>(a);
    O3DB_LAUNCH_CHECK();
    if (c->comm) {
        int rc = o3db_comm_allreduce_f64(c->comm, (double*)((char*)c->st + offsetof(IcpState, sums)), kNumSums, st);
        if (rc) return rc;
        icp_finalize_kernel<1><<<1, 32, 0, st>>>(a);
        O3DB_LAUNCH_CHECK();
    }
    O3DB_CUDA_CHECK(cudaMemcpyAsync(c->h_st, c->st, sizeof(IcpState), cudaMemcpyDeviceToHost, st));
    O3DB_CUDA_CHECK(cudaStreamSynchronize(st));
    const IcpState& h = *c->h_st;
    memcpy(result->transformation, h.T, sizeof(h.T));
    result->fitness = h.fitness;
    result->inlier_rmse = h.rmse;
    result->converged = h.converged;
    result->num_iterations = h.iter;
    result->num_correspondences = (int64_t)h.count;
    result->status = h.status ? O3DB_ERR_SINGULAR : O3DB_OK;
    if (per_iteration_host && h.executed > 0) {
        O3DB_CUDA_CHECK(cudaMemcpyAsync(per_iteration_host, c->per_iter, (size_t)h.executed * 2 * sizeof(double),
                                        cudaMemcpyDeviceToHost, st));
        O3DB_CUDA_CHECK(cudaStreamSynchronize(st));
    }
    if (h.status) {
        set_last_error("Singular 6x6 linear system detected, tracking failed.");
        return O3DB_ERR_SINGULAR;
    }
    return O3DB_OK;
}

int o3db_icp_point_to_plane(const float* source_dev, int64_t n, const float* target_dev,
                            const float* target_normals_dev, int64_t m, const double init_T[16],
                            const o3db_icp_options* options, o3db_icp_result* result, int64_t* correspondences_dev,
                            double* per_iteration_host, void* stream) {
    o3db_icp* c = nullptr;
    int rc = o3db_icp_create(source_dev, n, target_dev, target_normals_dev, m, init_T, options, nullptr, stream, &c);
    if (rc) return rc;
    rc = o3db_icp_iterate(c, options->max_iteration, stream);
    if (rc == O3DB_OK) rc = o3db_icp_finish(c, result, correspondences_dev, per_iteration_host, stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    o3db_icp_destroy(c);
    return rc;
}

int o3db_icp_colored(const float* source_dev, const float* source_colors_dev, int64_t n, const float* target_dev,
                     const float* target_normals_dev, const float* target_colors_dev,
                     const float* target_color_gradients_dev, int64_t m, const double init_T[16],
                     const o3db_icp_options* options, double lambda_geometric, o3db_icp_result* result,
                     int64_t* correspondences_dev, double* per_iteration_host, void* stream) {
    o3db_icp* c = nullptr;
    int rc = o3db_icp_create_colored(source_dev, source_colors_dev, n, target_dev, target_normals_dev,
                                     target_colors_dev, target_color_gradients_dev, m, init_T, options,
                                     lambda_geometric, nullptr, stream, &c);
    if (rc) return rc;
    rc = o3db_icp_iterate(c, options->max_iteration, stream);
    if (rc == O3DB_OK) rc = o3db_icp_finish(c, result, correspondences_dev, per_iteration_host, stream);
    cudaStreamSynchronize((cudaStream_t)stream);
    o3db_icp_destroy(c);
    return rc;
}

int o3db_icp_point_to_plane_host(const float* source_host, int64_t n, const float* target_host,
                                 const float* target_normals_host, int64_t m, const double init_T[16],
                                 const o3db_icp_options* options, o3db_icp_result* result,
                                 int64_t* correspondences_host, double* per_iteration_host) {
    O3DB_REQUIRE(source_host && target_host && n > 0 && m > 0, "Source and/or Target pointcloud is empty.");
    O3DB_REQUIRE(target_normals_host != nullptr, "Target pointcloud missing normals attribute.");
    configure_memory_pool();
    cudaStream_t st = 0;
    float *d_src = nullptr, *d_tgt = nullptr, *d_nrm = nullptr;
    int64_t* d_corr = nullptr;
    O3DB_CUDA_CHECK(cudaMallocAsync(&d_src, n * 3 * sizeof(float), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&d_tgt, m * 3 * sizeof(float), st));
    O3DB_CUDA_CHECK(cudaMallocAsync(&d_nrm, m * 3 * sizeof(float), st));
    if (correspondences_host) O3DB_CUDA_CHECK(cudaMallocAsync(&d_corr, n * sizeof(int64_t), st));
    O3DB_CUDA_CHECK(cudaMemcpyAsync(d_src, source_host, n * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    O3DB_CUDA_CHECK(cudaMemcpyAsync(d_tgt, target_host, m * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    O3DB_CUDA_CHECK(cudaMemcpyAsync(d_nrm, target_normals_host, m * 3 * sizeof(float), cudaMemcpyHostToDevice, st));
    int rc = o3db_icp_point_to_plane(d_src, n, d_tgt, d_nrm, m, init_T, options, result, d_corr, per_iteration_host, st);
    if (rc == O3DB_OK && correspondences_host) {
        cudaError_t e = cudaMemcpyAsync(correspondences_host, d_corr, n * sizeof(int64_t), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            set_last_error("D2H copy of correspondences failed: %s", cudaGetErrorString(e));
            rc = O3DB_ERR_CUDA;
        }
    }
    cudaFreeAsync(d_src, st);
    cudaFreeAsync(d_tgt, st);
    cudaFreeAsync(d_nrm, st);
    if (d_corr) cudaFreeAsync(d_corr, st);
    return rc;
}

}  // extern "C"
