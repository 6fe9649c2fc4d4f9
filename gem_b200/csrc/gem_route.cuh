// gem_route.cuh -- point routing for spatially tiled maps (BASELINE configs 4/5).
//
// The reference is single-GPU; a map that outgrows one GPU is cut into geographic tiles, one
// per rank.  Every rank transforms its share of the input, buckets the accepted in-grid
// points STABLY by owning tile, and the host side exchanges the buckets with one NCCL
// all-to-all (gem_b200/tiled.py).  Concatenating received buckets in (source rank, source
// order) reproduces the global point order, so the tiled result is bit-identical to the
// single-GPU result on the concatenated cloud.
#pragma once
#include "gem_add.cuh"

namespace gem {

constexpr int ROUTE_MAX_OWNERS = 64;
constexpr int ROUTE_BLOCK = 256;

// pass 1: transform, owner id, per-block owner histogram
__global__ void __launch_bounds__(ROUTE_BLOCK)
k_route_count(MapGeom g, FrameParams f, const float4 *xyzi, int n, int tile_h, int tile_w, int tiles_c,
              int n_owners, int *owner_out, int *gkey_out, float *h_out, float *hv_out, int *blockCounts /* [owners][blocks] */)
{
    __shared__ int s_cnt[ROUTE_MAX_OWNERS];
    for (int o = threadIdx.x; o < n_owners; o += blockDim.x) s_cnt[o] = 0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int owner = -1;
    if (i < n) {
        const float4 p = ld_stream_f4(xyzi + i);
        const PtRes r = transform_point(g, f, p.x, p.y, p.z);
        int gkey = -1;
        if (r.ingrid) {
            owner = (r.gx / tile_h) * tiles_c + (r.gy / tile_w);
            gkey = r.gx * g.L + r.gy;
        }
        owner_out[i] = owner;
        gkey_out[i] = gkey;
        h_out[i] = r.h;
        hv_out[i] = r.hv;
    }
    if (owner >= 0) atomicAdd(&s_cnt[owner], 1);
    __syncthreads();
    for (int o = threadIdx.x; o < n_owners; o += blockDim.x) blockCounts[o * gridDim.x + blockIdx.x] = s_cnt[o];
}

// pass 2: one block scans blockCounts owner-major -> exclusive offsets; owner totals
__global__ void __launch_bounds__(1024) k_route_scan(int *blockCounts, int n_owners, int nblocks, int *counts_out)
{
    __shared__ int s_part[1024];
    __shared__ int s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int total = n_owners * nblocks;
    for (int base = 0; base < total; base += 1024) {
        const int idx = base + threadIdx.x;
        const int v = idx < total ? blockCounts[idx] : 0;
        s_part[threadIdx.x] = v;
        __syncthreads();
        for (int d = 1; d < 1024; d <<= 1) { // Hillis-Steele inclusive scan
            const int t = threadIdx.x >= d ? s_part[threadIdx.x - d] : 0;
            __syncthreads();
            s_part[threadIdx.x] += t;
            __syncthreads();
        }
        const int incl = s_part[threadIdx.x] + s_carry;
        if (idx < total) blockCounts[idx] = incl - v; // exclusive, global over (owner, block)
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    // owner totals from consecutive owner starts
    __syncthreads();
    for (int o = threadIdx.x; o < n_owners; o += blockDim.x) {
        const int begin = blockCounts[o * nblocks];
        const int end = (o + 1 < n_owners) ? blockCounts[(o + 1) * nblocks] : s_carry;
        counts_out[o] = end - begin;
    }
}

// pass 3: stable in-block rank and record write
__global__ void __launch_bounds__(ROUTE_BLOCK)
k_route_write(const float4 *xyzi, const uchar4 *rgba, int n, int n_owners, const int *owner_in, const int *gkey_in,
              const float *h_in, const float *hv_in, const int *blockOffsets, RouteRec *out, int bucket_stride)
{
    __shared__ int s_wcnt[ROUTE_BLOCK / 32][ROUTE_MAX_OWNERS];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int w = threadIdx.x >> 5;
    for (int o = (int)lane; o < n_owners; o += 32) s_wcnt[w][o] = 0;
    __syncwarp();
    const int owner = (i < n) ? owner_in[i] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, owner);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (owner >= 0 && rank_in_warp == 0) s_wcnt[w][owner] = __popc(peers);
    __syncthreads();
    if (owner >= 0) {
        int before = 0;
        for (int ww = 0; ww < w; ww++) before += s_wcnt[ww][owner];
        int pos = blockOffsets[owner * gridDim.x + blockIdx.x] + before + rank_in_warp;
        // fixed-stride layout: bucket o starts at o*stride (padded all-to-all, no host-side split sizes)
        if (bucket_stride > 0) pos = pos - blockOffsets[owner * gridDim.x] + owner * bucket_stride;
        RouteRec r;
        r.gkey = gkey_in[i];
        r.h = h_in[i];
        r.var = hv_in[i];
        r.rgb = 0u;
        if (rgba) {
            const uchar4 c = rgba[i];
            r.rgb = pack_rgb(c.x, c.y, c.z);
        }
        r.intensity = xyzi[i].w;
        out[pos] = r;
    }
}

// ---- peer-memory routing: compute + "collective" in one kernel ------------------------------
// Instead of bucketing locally and calling an all-to-all, every record is stored straight into the
// OWNING rank's receive buffer through a peer mapping (NVLink / NVSwitch; torch symmetric memory
// provides the mappings, gem_b200/tiled.py).  Rank r's records for owner o land in bucket r of o's
// buffer, in source order, so the receiver sees (source rank, source order) exactly as with the
// all-to-all.  The per-(source,owner) count goes to the owner's count array the same way.
struct PeerTable {
    unsigned long long recv[ROUTE_MAX_OWNERS];   // RouteRec* of every rank's receive buffer (current parity)
    unsigned long long counts[ROUTE_MAX_OWNERS]; // int* of every rank's per-source count array
};

__global__ void __launch_bounds__(ROUTE_BLOCK)
k_route_write_peer(const float4 *xyzi, const uchar4 *rgba, int n, int n_owners, const int *owner_in, const int *gkey_in,
                   const float *h_in, const float *hv_in, const int *blockOffsets, const int *counts_local,
                   const __grid_constant__ PeerTable pt, int my_rank, int bucket_stride)
{
    __shared__ int s_wcnt[ROUTE_BLOCK / 32][ROUTE_MAX_OWNERS];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned lane = threadIdx.x & 31u;
    const int w = threadIdx.x >> 5;
    if (blockIdx.x == 0 && (int)threadIdx.x < n_owners) // publish my bucket sizes to their owners
        ((int *)pt.counts[threadIdx.x])[my_rank] = counts_local[threadIdx.x];
    for (int o = (int)lane; o < n_owners; o += 32) s_wcnt[w][o] = 0;
    __syncwarp();
    const int owner = (i < n) ? owner_in[i] : -1;
    const unsigned peers = __match_any_sync(0xffffffffu, owner);
    const int rank_in_warp = __popc(peers & ((1u << lane) - 1u));
    if (owner >= 0 && rank_in_warp == 0) s_wcnt[w][owner] = __popc(peers);
    __syncthreads();
    if (owner >= 0) {
        int before = 0;
        for (int ww = 0; ww < w; ww++) before += s_wcnt[ww][owner];
        const int pos = blockOffsets[owner * gridDim.x + blockIdx.x] - blockOffsets[owner * gridDim.x] + before + rank_in_warp;
        RouteRec r;
        r.gkey = gkey_in[i];
        r.h = h_in[i];
        r.var = hv_in[i];
        r.rgb = 0u;
        if (rgba) {
            const uchar4 c = rgba[i];
            r.rgb = pack_rgb(c.x, c.y, c.z);
        }
        r.intensity = xyzi[i].w;
        RouteRec *dst = (RouteRec *)pt.recv[owner] + (size_t)my_rank * bucket_stride + pos; // peer store over NVLink
        *dst = r;
    }
}

// per-point scratch of the routing passes
struct RouteScratch {
    int *owner, *gkey;
    float *h, *hv;
    int *blockCounts; // [owners][blocks]
    size_t blockCounts_capacity;
};
inline cudaError_t route_points(cudaStream_t st, const MapGeom &g, const FrameParams &fp, const float4 *xyzi,
                                const uchar4 *rgba, int n, int tiles_r, int tiles_c, const RouteScratch &sc,
                                RouteRec *out, int *counts_out, int bucket_stride,
                                const PeerTable *peer = nullptr, int my_rank = 0)
{
    const int n_owners = tiles_r * tiles_c;
    const int nblocks = n > 0 ? (n + ROUTE_BLOCK - 1) / ROUTE_BLOCK : 1;
    if ((size_t)n_owners * nblocks > sc.blockCounts_capacity) return cudaErrorInvalidValue;
    const int tile_h = (g.L + tiles_r - 1) / tiles_r, tile_w = (g.L + tiles_c - 1) / tiles_c;
    k_route_count<<<nblocks, ROUTE_BLOCK, 0, st>>>(g, fp, xyzi, n, tile_h, tile_w, tiles_c, n_owners, sc.owner, sc.gkey,
                                                  sc.h, sc.hv, sc.blockCounts);
    k_route_scan<<<1, 1024, 0, st>>>(sc.blockCounts, n_owners, nblocks, counts_out);
    if (peer)
        k_route_write_peer<<<nblocks, ROUTE_BLOCK, 0, st>>>(xyzi, rgba, n, n_owners, sc.owner, sc.gkey, sc.h, sc.hv, sc.blockCounts,
                                                          counts_out, *peer, my_rank, bucket_stride);
    else
        k_route_write<<<nblocks, ROUTE_BLOCK, 0, st>>>(xyzi, rgba, n, n_owners, sc.owner, sc.gkey, sc.h, sc.hv, sc.blockCounts, out, bucket_stride);
    return cudaGetLastError();
}

} // namespace gem
