// gem_submap.cuh -- loop-closure re-fusion of submaps: what ElevationMapping::updateGlobalMap
// (ElevationMapping.cpp:773-905) does with pcl::transformPointCloud, two std::unordered_map's and a pairwise formula,
// on the device: rigid re-transform of a submap's PointXYZRGBICT records, and the pairwise fusion of two submaps through
// open-addressing hash tables of their cells.  DESIGN.md section "f4" lists, item by item, what is reproduced literally
// and what is DEFINED here because the reference leaves it to unordered_map iteration order or to uninitialised memory.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gem {

struct SubPoint { // PointXYZRGBICT.hpp:26-48, 32 bytes
    float x, y, z, w;
    uint32_t bgra;
    float covariance, intensity, travers;
};

static_assert(sizeof(SubPoint) == 32, "PointXYZRGBICT is 32 bytes");

// pcl::transformPointCloud (ElevationMapping.cpp:805): x' = t00 x + t01 y + t02 z + t03, left to right in float (the
// scalar code of PCL <= 1.9; PCL is an unpinned dependency of the reference).  T: row-major 4 x 4.
struct Rigid { float t[12]; };
__global__ void __launch_bounds__(256) k_transform_cloud(SubPoint *p, int n, const __grid_constant__ Rigid T)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float x = p[i].x, y = p[i].y, z = p[i].z;
    p[i].x = ((T.t[0] * x + T.t[1] * y) + T.t[2] * z) + T.t[3];
    p[i].y = ((T.t[4] * x + T.t[5] * y) + T.t[6] * z) + T.t[7];
    p[i].z = ((T.t[8] * x + T.t[9] * y) + T.t[10] * z) + T.t[11];
}

// pointCloudtoHash (ElevationMapping.cpp:1180-1192): the cell of a point is the float pair
// (ceil(x / res) * res - res / 2, same for y), evaluated in double (resolution_ is a double) and stored to float;
// GridPointEqual compares the floats.
__device__ __forceinline__ unsigned long long cell_key(float x, float y, double res, float &rx, float &ry)
{
    rx = (float)(ceil((double)x / res) * res - res / 2.0);
    ry = (float)(ceil((double)y / res) * res - res / 2.0);
    if (rx == 0.0f) rx = 0.0f; // -0 == +0 for GridPointEqual
    if (ry == 0.0f) ry = 0.0f;
    return ((unsigned long long)__float_as_uint(rx) << 32) | (unsigned long long)__float_as_uint(ry);
}
constexpr unsigned long long HASH_EMPTY = 0xffffffffffffffffull; // NaN/NaN pattern: no real cell has it

__device__ __forceinline__ unsigned hash_slot(unsigned long long k, unsigned mask)
{
    k ^= k >> 33; k *= 0xff51afd7ed558ccdull; k ^= k >> 33; k *= 0xc4ceb9fe1a85ec53ull; k ^= k >> 33;
    return (unsigned)k & mask;
}
// umap::insert keeps the FIRST point of a cell: the table stores, per cell, the smallest point index
__global__ void __launch_bounds__(256) k_hash_insert(const SubPoint *p, int n, double res, unsigned long long *keys, int *first, unsigned mask)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float rx, ry;
    const unsigned long long k = cell_key(p[i].x, p[i].y, res, rx, ry);
    if (k == HASH_EMPTY || rx != rx || ry != ry) return; // a NaN position equals nothing, itself included: such points never meet another
    unsigned s = hash_slot(k, mask);
    for (;;) {
        const unsigned long long prev = atomicCAS(&keys[s], HASH_EMPTY, k);
        if (prev == HASH_EMPTY || prev == k) { atomicMin(&first[s], i); return; }
        s = (s + 1) & mask;
    }
}
__device__ __forceinline__ int hash_find(const unsigned long long *keys, const int *first, unsigned mask, unsigned long long k)
{
    unsigned s = hash_slot(k, mask);
    for (;;) {
        const unsigned long long q = keys[s];
        if (q == k) return first[s];
        if (q == HASH_EMPTY) return -1;
        s = (s + 1) & mask;
    }
}

// One thread per point of the NEW submap (the neighbour, "out_new"); `old` is submap i ("out_old").  keep_*[i] = 1 iff
// point i is the first of its cell (those are the points the hash maps hold and localHashtoPointCloud emits).
// ElevationMapping.cpp:847-870, with every cell present in both maps fused exactly ONCE (DEFINITION: the reference
// erases and re-inserts while iterating, so what it visits twice depends on libstdc++'s bucket order).
// compat != 0: the fused values as the reference's expression evaluates (C operator precedence, :862-863);
// compat == 0: the weighting the expression was written for.
__global__ void __launch_bounds__(256)
k_refuse_pair(SubPoint *pn, int nn, SubPoint *po, int no, double res, const unsigned long long *kn, const int *fn, unsigned mn,
              const unsigned long long *ko, const int *fo, unsigned mo, unsigned char *keep_n, unsigned char *keep_o, int compat, int *count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < no) { // old map: which points it holds; their position becomes the cell's (localHashtoPointCloud :1129-1130)
        float rx, ry;
        const unsigned long long k = cell_key(po[i].x, po[i].y, res, rx, ry);
        const bool nan = rx != rx || ry != ry;
        const bool first = nan || hash_find(ko, fo, mo, k) == i;
        keep_o[i] = first ? 1 : 0;
        if (first) { po[i].x = rx; po[i].y = ry; po[i].w = 1.0f; }
    }
    if (i >= nn) return;
    float rx, ry;
    const unsigned long long k = cell_key(pn[i].x, pn[i].y, res, rx, ry);
    const bool nan = rx != rx || ry != ry;
    const bool first = nan || hash_find(kn, fn, mn, k) == i;
    keep_n[i] = first ? 1 : 0;
    if (!first) return;
    SubPoint a = pn[i];
    a.x = rx; a.y = ry; a.w = 1.0f;
    const int j = nan ? -1 : hash_find(ko, fo, mo, k);
    if (j >= 0) {
        const float vo = po[j].covariance, eo = po[j].z, vn = a.covariance, en = a.z;
        if (vo > 0.0f && vo < 1.0f) { // :857
            const double vn2 = (double)vn * (double)vn, vo2 = (double)vo * (double)vo; // pow(float, 2) in double: exact squares
            float ef, vf;
            if (compat) { // :862-863 as C parses them
                ef = (float)((vn2 * (double)eo + vo2 * (double)en / vo2) + vn2);
                vf = (float)(vo2 * vn2 / vo2 + vn2);
            } else {
                ef = (float)((vn2 * (double)eo + vo2 * (double)en) / (vo2 + vn2));
                vf = (float)(vo2 * vn2 / (vo2 + vn2));
            }
            a.z = ef;
            a.covariance = vf;
            SubPoint b = a; // both maps get the fused cell, with the NEW map's colour / intensity / travers (:856)
            po[j] = b;      // (x, y are the cell's position in both)
            atomicAdd(count, 1);
        }
    }
    pn[i] = a;
}

// order-preserving compaction of the kept points (one block; a loop-closure event is rare and a submap has < 1e6 points)
__global__ void __launch_bounds__(1024) k_compact_points(const SubPoint *in, const unsigned char *keep, int n, SubPoint *out, int *n_out)
{
    __shared__ int s_w[32];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u;
    const int w = threadIdx.x >> 5;
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const int f = (i < n && keep[i]) ? 1 : 0;
        const unsigned b = __ballot_sync(0xffffffffu, f);
        if (lane == 0u) s_w[w] = __popc(b);
        __syncthreads();
        int before = s_carry;
        for (int q = 0; q < w; q++) before += s_w[q];
        if (f) out[before + __popc(b & ((1u << lane) - 1u))] = in[i];
        __syncthreads();
        if (threadIdx.x == 0) { int t = 0; for (int q = 0; q < 32; q++) t += s_w[q]; s_carry += t; }
        __syncthreads();
    }
    if (threadIdx.x == 0) *n_out = s_carry;
}

} // namespace gem
