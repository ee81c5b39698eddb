// hash.cuh — lock-free open-addressing hash of int32x3 keys shared by the voxel block grid
// (tsdf.cu) and PointCloud::VoxelDownSample (pointcloud.cu).  Hash function == the reference's
// utility::MiniVecHash<int,3> (core/hashmap/Dispatch.h:67-81), bit for bit.
#pragma once

#include <climits>

#include "common.cuh"

namespace o3db {

static constexpr int kEmpty = -1;
static constexpr int kTomb = INT_MIN;

// utility::MiniVecHash<int,3> (core/hashmap/Dispatch.h:67-81): FNV-1a style over
// the elements, int32 sign-extended to uint64.
__host__ __device__ __forceinline__ uint64_t minivec_hash(int x, int y, int z) {
    uint64_t h = 14695981039346656037ull;
    h ^= (uint64_t)(int64_t)x;
    h *= 1099511628211ull;
    h ^= (uint64_t)(int64_t)y;
    h *= 1099511628211ull;
    h ^= (uint64_t)(int64_t)z;
    h *= 1099511628211ull;
    return h;
}

__device__ __forceinline__ unsigned bucket_of(uint64_t h, unsigned mask) {
    return ((unsigned)h ^ (unsigned)(h >> 32)) & mask;
}

// Open addressing, linear probing.  table[b] is
//   kEmpty            free
//   v >= 0            committed: slot v of the key/value buffers
//   v <= -2           provisional (inserted by the running call): candidate -(v+2)
//   kTomb             dead bucket (only after an overflow)
struct Table {
    int* table;
    unsigned mask;          // nbuckets - 1
    const int* keys;        // committed keys [capacity,3]
};

enum { kResInserted = -2, kResDuplicate = -3, kResFull = -4, kResMiss = -5 };

// Looks `k` up; if absent and INSERT, claims a bucket with the provisional marker
// of candidate `cand` (whose key must already be globally visible in cand_keys).
// Returns slot >= 0 (committed), kResInserted (+ *bucket), kResDuplicate (another
// candidate of this call holds the key), kResMiss or kResFull.
template <bool INSERT>
__device__ __forceinline__ int probe(const Table& t, const int* __restrict__ cand_keys, int cand, int kx, int ky,
                                     int kz, unsigned* bucket) {
    unsigned b = bucket_of(minivec_hash(kx, ky, kz), t.mask);
    for (unsigned n = 0; n <= t.mask; ++n, b = (b + 1) & t.mask) {
        int v = __ldcg(&t.table[b]);
        if (v == kEmpty) {
            if (!INSERT) return kResMiss;
            const int old = atomicCAS(&t.table[b], kEmpty, -(cand + 2));
            if (old == kEmpty) {
                *bucket = b;
                return kResInserted;
            }
            v = old;
        }
        if (v == kTomb) continue;
        if (v < 0 && cand_keys == nullptr) continue;   // a provisional entry seen by a reader without the candidate array
                                                       // (ray cast / find): not a committed key — keep probing
        const int* kk = v >= 0 ? t.keys + 3 * (size_t)v : cand_keys + 3 * (size_t)(-(v + 2));
        if (__ldcg(kk) == kx && __ldcg(kk + 1) == ky && __ldcg(kk + 2) == kz) return v >= 0 ? v : kResDuplicate;
    }
    return kResFull;
}


}  // namespace o3db
