// pointcloud.cu — t::geometry::PointCloud::VoxelDownSample for sm_100a (the
// pyramid build that sits directly in front of the ICP loop, Registration.cpp:237-240, 266-269).
//
// Reference (t/geometry/PointCloud.cpp:496-560): voxel = floor(p / voxel_size) (f32), a HashSet
// insert + find to map points to dense voxel ids, then one IndexAdd_ pass per attribute and a
// division by the per-voxel count.  Here: one kernel hashes every point's voxel key into a
// lock-free table (the first point of a voxel claims a dense id), one kernel accumulates all
// attributes with f32 atomics (the reference's IndexAdd_ is f32 too), one divides.
#include <climits>
#include <cmath>

#include "common.cuh"
#include "hash.cuh"

namespace o3db {

static constexpr int kVT = 256;

struct VdsArgs {
    const float* pos;
    const float* nrm;   // may be null
    const float* col;   // may be null
    int n;
    float voxel_size;
    int* keys;          // [n,3] voxel key of every point
    Table tab;          // committed keys = vkeys
    int* vkeys;         // [n,3] key of every voxel id
    int* counter;       // number of voxels
    int* vid;           // [n] voxel id of the point (filled by the winners in pass 1, by all in pass 2)
    float* pos_out;
    float* nrm_out;
    float* col_out;
    float* cnt;         // [n] points per voxel (f32, as voxel_num_points upstream)
};

// pass 1: voxel keys + claim.  A winner takes the next dense id and publishes it at once.
__global__ void vds_claim_kernel(VdsArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    // PointCloud.cpp:506-507: (positions / voxel_size).Floor().To(Int64) — Float32 division
    const int kx = (int)floorf(__fdiv_rn(a.pos[3 * i], a.voxel_size));
    const int ky = (int)floorf(__fdiv_rn(a.pos[3 * i + 1], a.voxel_size));
    const int kz = (int)floorf(__fdiv_rn(a.pos[3 * i + 2], a.voxel_size));
    a.keys[3 * i] = kx;
    a.keys[3 * i + 1] = ky;
    a.keys[3 * i + 2] = kz;
    __threadfence();   // the key must be visible before a marker can point at it
    unsigned bucket = 0;
    const int r = probe<true>(a.tab, a.keys, i, kx, ky, kz, &bucket);
    int id = -1;
    if (r == kResInserted) {
        id = atomicAdd(a.counter, 1);
        a.vkeys[3 * id] = kx;
        a.vkeys[3 * id + 1] = ky;
        a.vkeys[3 * id + 2] = kz;
        __threadfence();
        atomicExch(&a.tab.table[bucket], id);   // provisional marker -> committed id
    } else if (r >= 0) {
        id = r;
    }
    a.vid[i] = id;   // -1: the voxel was still provisional when this point looked; resolved in pass 2
}

// pass 2: every point adds its attributes to its voxel (IndexAdd_, PointCloud.cpp:536-552).
__global__ void vds_accumulate_kernel(VdsArgs a) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    int id = a.vid[i];
    if (id < 0) {
        unsigned bucket;
        id = probe<false>(a.tab, a.keys, 0, a.keys[3 * i], a.keys[3 * i + 1], a.keys[3 * i + 2], &bucket);
    }
    if (id < 0) return;   // cannot happen: every key was inserted in pass 1
    atomicAdd(&a.cnt[id], 1.0f);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        atomicAdd(&a.pos_out[3 * (size_t)id + c], a.pos[3 * (size_t)i + c]);
        if (a.nrm) atomicAdd(&a.nrm_out[3 * (size_t)id + c], a.nrm[3 * (size_t)i + c]);
        if (a.col) atomicAdd(&a.col_out[3 * (size_t)id + c], a.col[3 * (size_t)i + c]);
    }
}

// pass 3: voxel_attr /= voxel_num_points (PointCloud.cpp:549)
__global__ void vds_divide_kernel(VdsArgs a) {
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= *a.counter) return;
    const float c = a.cnt[v];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        a.pos_out[3 * (size_t)v + k] = a.pos_out[3 * (size_t)v + k] / c;
        if (a.nrm) a.nrm_out[3 * (size_t)v + k] = a.nrm_out[3 * (size_t)v + k] / c;
        if (a.col) a.col_out[3 * (size_t)v + k] = a.col_out[3 * (size_t)v + k] / c;
    }
}

}  // namespace o3db

using namespace o3db;

extern "C" int o3db_voxel_down_sample(const float* positions_dev, const float* normals_dev, const float* colors_dev,
                                      int64_t n, double voxel_size, float* positions_out_dev, float* normals_out_dev,
                                      float* colors_out_dev, int64_t* num_out_host, void* stream) {
    O3DB_REQUIRE(voxel_size > 0, "voxel_size must be positive.");   // PointCloud.cpp:498-500
    O3DB_REQUIRE(n >= 0 && n < INT_MAX / 4, "o3db_voxel_down_sample: bad point count");
    O3DB_REQUIRE(num_out_host != nullptr, "o3db_voxel_down_sample: num_out_host is null");
    *num_out_host = 0;
    if (n == 0) return O3DB_OK;
    O3DB_REQUIRE(positions_dev && positions_out_dev, "o3db_voxel_down_sample: null positions");
    O3DB_REQUIRE((normals_dev == nullptr) == (normals_out_dev == nullptr) &&
                         (colors_dev == nullptr) == (colors_out_dev == nullptr),
                 "o3db_voxel_down_sample: attribute in/out buffers must come in pairs");
    configure_memory_pool();
    cudaStream_t st = (cudaStream_t)stream;
    unsigned nb = 16;
    while ((int64_t)nb < 2 * n) nb <<= 1;
    char* base = nullptr;
    const size_t b_keys = (size_t)n * 3 * sizeof(int), b_tab = (size_t)nb * sizeof(int), b_vid = (size_t)n * sizeof(int),
                 b_cnt = (size_t)n * sizeof(float);
    O3DB_CUDA_CHECK(cudaMallocAsync(&base, 2 * b_keys + b_tab + b_vid + b_cnt + 64, st));
    VdsArgs a{};
    a.pos = positions_dev;
    a.nrm = normals_dev;
    a.col = colors_dev;
    a.n = (int)n;
    a.voxel_size = (float)voxel_size;   // scalar operand takes the tensor's dtype (Float32)
    a.keys = (int*)base;
    a.vkeys = (int*)(base + b_keys);
    int* table = (int*)(base + 2 * b_keys);
    a.vid = (int*)(base + 2 * b_keys + b_tab);
    a.cnt = (float*)(base + 2 * b_keys + b_tab + b_vid);
    a.counter = (int*)(base + 2 * b_keys + b_tab + b_vid + b_cnt);
    a.tab = Table{table, nb - 1, a.vkeys};
    a.pos_out = positions_out_dev;
    a.nrm_out = normals_out_dev;
    a.col_out = colors_out_dev;
    O3DB_CUDA_CHECK(cudaMemsetAsync(table, 0xff, b_tab, st));
    O3DB_CUDA_CHECK(cudaMemsetAsync(a.cnt, 0, b_cnt + 64, st));   // counts + counter
    O3DB_CUDA_CHECK(cudaMemsetAsync(positions_out_dev, 0, (size_t)n * 3 * sizeof(float), st));
    if (normals_out_dev) O3DB_CUDA_CHECK(cudaMemsetAsync(normals_out_dev, 0, (size_t)n * 3 * sizeof(float), st));
    if (colors_out_dev) O3DB_CUDA_CHECK(cudaMemsetAsync(colors_out_dev, 0, (size_t)n * 3 * sizeof(float), st));
    const unsigned grid = (unsigned)ceil_div(n, kVT);
    vds_claim_kernel<<<grid, kVT, 0, st>>>(a);
    O3DB_LAUNCH_CHECK();
    vds_accumulate_kernel<<<grid, kVT, 0, st>>>(a);
    O3DB_LAUNCH_CHECK();
    vds_divide_kernel<<<grid, kVT, 0, st>>>(a);
    O3DB_LAUNCH_CHECK();
    int m = 0;
    O3DB_CUDA_CHECK(cudaMemcpyAsync(&m, a.counter, sizeof(int), cudaMemcpyDeviceToHost, st));
    O3DB_CUDA_CHECK(cudaStreamSynchronize(st));
    O3DB_CUDA_CHECK(cudaFreeAsync(base, st));
    *num_out_host = m;
    return O3DB_OK;
}
