// gem_add.cuh -- the add path of libgem_b200: Process_points + Fuse (gpu.cu:384-455, 477-537) in TWO kernels.
//
//   k_bin   1 thread/point : float4 load, SE(3), filters, sensor variance, cell key, per-cell arrival rank via one
//                            L2 atomic on the cell's own 32-byte record, record slot allocated IN the kernel
//                            (touched-list slot = first chunk of 8 records; geometric overflow chunks from a bump
//                            pool), 16-byte record {point index, h, var, rgb} stored straight to its slot
//   k_fold  1 thread or 1 warp/touched cell : order the cell's records by point index (== the order in which
//                            G_fuse's per-cell loop visits them), sequential Kalman fold with the 5-sigma gate,
//                            lowest-scan update, one 16 B write-back per cell
//
// Round 1 had four kernels (transform+bin, per-cell allocation, scatter, fold) and five per-cell arrays; every
// touched cell cost five random 32-byte sectors per call.  Here a cell IS one sector, and the allocation and scatter
// kernels are gone: a point learns its slot inside k_bin by waiting for a pointer that a point with a LOWER arrival
// rank in the same cell publishes (that thread has already executed its atomic, so it is resident and running, and
// it publishes without waiting for anything of equal or higher level -- the wait-for graph is acyclic, see bin_points).
#pragma once
#include "gem_kernels.cuh"

namespace gem {

// ---------------------------------------------------------------------------------------
// chunk geometry: ranks 0..7 live in chunk 0 (indexed by the cell's touched-list slot), ranks >= 8 in chunks of
// 32, 128, 512, ... records (x4 per level) bump-allocated from the pool by the point that draws the level's first
// rank.  A pool chunk is one header slot {next-level offset, 0, 0, 0} followed by its records.
// ---------------------------------------------------------------------------------------
constexpr int CHUNK0 = 8;
__host__ __device__ __forceinline__ int level_base(int j) { return (CHUNK0 * ((1 << (2 * j)) - 1)) / 3; } // 8, 40, 168, 680, ...
__host__ __device__ __forceinline__ int level_cap(int j) { return CHUNK0 << (2 * j); }                    // 32, 128, 512, ...
__device__ __forceinline__ int level_of(int rank) // rank >= CHUNK0
{
    int j = 1;
    while (rank >= level_base(j + 1)) j++;
    return j;
}
constexpr int FOLD_LARGE_FROM = CHUNK0;  // a cell whose arrival counter reaches this rank joins the "large" list (k > 8)
constexpr int FOLD_LONG_FROM = 40;       // = level_base(2): joins the "long" list (k > 40), folded first

struct BinCounters { // one hot counter per 128-byte line
    int ntouched; int pad0[31];
    int pool;     int pad1[31];  // pool slots handed out (offset of the next chunk - 1)
    int nlarge;   int pad2[31];
    int nlong;    int pad3[31];
    int total;    int pad4[31];  // points binned (accepted AND inside the grid / tile)
    int maxk;     int pad5[31];  // longest list among the large cells (written by the fold)
};

struct BinScratch { // one set per call parity
    int *touched;      // [T]   keys of the touched cells in slot order
    uint4 *chunk0;     // [8 T] records of rank 0..7 of the cell in touched slot t
    int *ovf1;         // [T]   pool offset of the level-1 chunk of slot t, 0 = not published; zero between calls
    uint4 *pool;       // [pool_cap + 1] overflow chunks; offset 0 is never handed out
    int4 *tlarge;      // {key, slot, p1, 0}  cells that reached rank 8
    int4 *tlong;       // {key, slot, p1, p2} cells that reached rank 40
    BinCounters *ctr;      // counters of this call (zero when it starts)
    BinCounters *ctr_next; // zeroed by this call's bin kernel for the call after
    int par;           // which {counter, slot} pair of the cell records this call uses
    int pool_cap;
};

// acquire / release at GPU scope on 32-bit words (pointer publication between running blocks)
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int spin_nonzero(const int *p)
{
    int v = ld_acquire(p);
    while (v == 0) {
        __nanosleep(32);
        v = ld_acquire(p);
    }
    return v;
}

// ---------------------------------------------------------------------------------------
// point sources of the bin kernel
// ---------------------------------------------------------------------------------------
enum { SRC_XYZI = 0, SRC_SOA = 1, SRC_PCL32 = 2, SRC_KEYS = 3, SRC_RECORDS = 4 };

struct RouteRec { // 20 bytes on the wire between tiles (gem_route.cuh)
    int gkey;     // global geographic linear index gx*L+gy
    float h, var;
    uint32_t rgb;
    float intensity;
};

struct BinSource {
    // SRC_XYZI: float4 {x,y,z,intensity} + optional uchar4 rgba; SRC_PCL32: 2 x float4 per point (PointXYZRGBICT.hpp:26-48)
    const float4 *xyzi;
    const uchar4 *rgba;
    const float4 *pcl;
    // SRC_SOA (Process_points, gpu.cu:1085): inputs x, y, z; optional outputs = what Process_points returns
    const float *x, *y, *z;
    int *key_out;
    float *h_out, *hv_out, *xt_out, *yt_out;
    // SRC_KEYS (Fuse, gpu.cu:1154): map_index, height, var, R, G, B, intensity arrays
    const int *key_in, *R, *G, *B;
    const float *h_in, *hv_in, *inten_in;
    int ncells;
    // SRC_RECORDS (tiled maps): records received from the other tiles, buckets of `stride` slots filled up to
    // src_counts[bucket] (src_counts == nullptr: all n slots are records; gkey < 0 = padding)
    const RouteRec *rec;
    const int *src_counts;
    int stride;
};

struct PointOut {
    int key;       // layer key or -1
    float h, hv;
    uint32_t rgbf; // rgb | REC_COLOUR_OK
};

template <int SRC>
__device__ __forceinline__ PointOut bin_source_point(const MapGeom &g, const FrameParams &f, const BinSource &in, int i,
                                                     const SegTable *segs, const FrameParams *frames)
{
    PointOut o;
    o.key = -1; o.h = -1.0f; o.hv = -1.0f; o.rgbf = 0u;
    if (SRC == SRC_XYZI) {
        const float4 p = ld_stream_f4(in.xyzi + i);
        const PtRes r = segs ? transform_point(g, frames[find_segment(*segs, i)], p.x, p.y, p.z) : transform_point(g, f, p.x, p.y, p.z);
        if (r.ingrid) o.key = local_key(g, r.gx, r.gy);
        o.h = r.h; o.hv = r.hv;
        uint32_t rgb = 0u;
        if (in.rgba) {
            const uchar4 c = in.rgba[i];
            rgb = pack_rgb(c.x, c.y, c.z);
        }
        o.rgbf = with_colour_flag(rgb, p.w);
    } else if (SRC == SRC_PCL32) {
        const float4 p = ld_stream_f4(in.pcl + 2 * (size_t)i);
        const float4 q = ld_stream_f4(in.pcl + 2 * (size_t)i + 1); // {rgb(b,g,r,a bytes), covariance, intensity, travers}
        const PtRes r = transform_point(g, f, p.x, p.y, p.z);
        if (r.ingrid) o.key = local_key(g, r.gx, r.gy);
        o.h = r.h; o.hv = r.hv;
        const uint32_t bgra = __float_as_uint(q.x);
        o.rgbf = with_colour_flag(pack_rgb((bgra >> 16) & 255, (bgra >> 8) & 255, bgra & 255), q.z);
    } else if (SRC == SRC_SOA) {
        const PtRes r = transform_point(g, f, in.x[i], in.y[i], in.z[i]);
        if (r.ingrid) o.key = local_key(g, r.gx, r.gy);
        o.h = r.h; o.hv = r.hv;
        if (in.key_out) in.key_out[i] = o.key;
        if (in.h_out) in.h_out[i] = r.h;
        if (in.hv_out) in.hv_out[i] = r.hv;
        if (in.xt_out) { in.xt_out[i] = r.xt; in.yt_out[i] = r.yt; }
    } else if (SRC == SRC_KEYS) {
        int key = in.key_in[i];
        if (key < 0 || key >= in.ncells) key = -1; // no G_fuse thread has such a map_index
        o.key = key;
        o.h = in.h_in[i]; o.hv = in.hv_in[i];
        // the reference tests R,G,B != 0 on int values; channels are 8-bit by construction (PointXYZRGBICT r/g/b are
        // uint8, SPB.cpp:164-166).  A non-zero int whose low byte is zero is mapped to 255 in that byte so "!= 0" holds.
        const int r = in.R ? in.R[i] : 0, gg = in.G ? in.G[i] : 0, b = in.B ? in.B[i] : 0;
        const uint32_t rgb = pack_rgb((r != 0 && (r & 255) == 0) ? 255 : r, (gg != 0 && (gg & 255) == 0) ? 255 : gg,
                                      (b != 0 && (b & 255) == 0) ? 255 : b);
        o.rgbf = with_colour_flag(rgb, in.inten_in ? in.inten_in[i] : 0.0f);
    } else { // SRC_RECORDS
        bool valid = true;
        if (in.src_counts) {
            const int s = i / in.stride;
            valid = (i - s * in.stride) < in.src_counts[s];
        }
        if (valid) {
            const RouteRec r = in.rec[i];
            if (r.gkey >= 0) {
                const int gx = r.gkey / g.L, gy = r.gkey - gx * g.L;
                o.key = local_key(g, gx, gy);
            }
            o.h = r.h; o.hv = r.var;
            o.rgbf = with_colour_flag(r.rgb, r.intensity);
        }
    }
    return o;
}

// ---------------------------------------------------------------------------------------
// the bin phase: U points per thread and iteration; every slot u of an iteration is a coalesced row of points
// (index base + u*nthreads + tid).  Must be entered by whole blocks (block barriers inside).
//
// Deadlock freedom.  Within an iteration a thread goes through: (A) loads, arithmetic, arrival atomics -- no
// waiting; (B) block-aggregated reservation of touched slots / pool space / list slots -- block barriers, reached by
// every thread of the block without waiting on another block; (C0) publication of the touched slot by rank-0 points
// -- no waiting; (C1..Cj) publication of level-j chunk pointers, level by level over all of the thread's points: a
// level-j publisher waits only for pointers of levels < j of the same cell; (D) record stores: wait for the
// pointers of the point's own cell.  Whoever is waited for has drawn a lower rank in that cell, i.e. has executed
// its atomic in phase A of ITS current iteration, and needs no barrier after B to publish.  The trailing barrier
// keeps a block's threads in the same iteration, so a block waiting in B never has a member spinning in D.
// ---------------------------------------------------------------------------------------
template <int SRC, int U>
__device__ __forceinline__ void bin_points(const MapGeom &g, const FrameParams &f, const BinSource &in, int n, Cell *cells,
                                           const BinScratch &sc, int tid, int nthreads, const SegTable *segs,
                                           const FrameParams *frames)
{
    __shared__ int s_w[5][ADD_BLOCK_MAX / 32]; // per-warp totals: touched, pool slots, large, long, binned
    __shared__ int s_b[5];
    const unsigned lane = threadIdx.x & 31u;
    const int w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    const int par = sc.par;
    for (int base = 0; base < n; base += U * nthreads) { // trip count identical for every thread of the grid
        int key[U], rank[U];
        uint4 rec[U];
        // ---- A: load, transform, arrival rank ------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = base + u * nthreads + tid;
            key[u] = -1;
            rec[u] = make_uint4(0u, 0u, 0u, 0u);
            if (i < n) {
                const PointOut o = bin_source_point<SRC>(g, f, in, i, segs, frames);
                key[u] = o.key;
                rec[u] = make_uint4((uint32_t)i, __float_as_uint(o.h), __float_as_uint(o.hv), o.rgbf);
            }
        }
#pragma unroll
        for (int u = 0; u < U; u++) rank[u] = (key[u] >= 0) ? atomicAdd(&cells[key[u]].bin[par].x, 1) : -1;
        // ---- B: block-aggregated reservations --------------------------------------------------------------
        int lvl[U], mypool = 0;
        uint32_t packed = 0; // byte 0: firsts, 1: rank == 8, 2: rank == 40, 3: binned (each <= U <= 4 per thread)
#pragma unroll
        for (int u = 0; u < U; u++) {
            lvl[u] = 0; // 0 = not an allocator
            if (rank[u] >= 0) {
                packed += 1u << 24;
                if (rank[u] == 0) packed += 1u;
                else if (rank[u] >= CHUNK0) {
                    const int j = level_of(rank[u]);
                    if (rank[u] == level_base(j)) {
                        lvl[u] = j;
                        mypool += level_cap(j) + 1;
                        if (j == 1) packed += 1u << 8;
                        if (j == 2) packed += 1u << 16;
                    }
                }
            }
        }
        uint32_t incl = packed; // warp inclusive scan of the four byte-wide counts (warp sums <= 128: no carry between bytes)
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
            if ((int)lane >= d) incl += t;
        }
        int pincl = mypool;
        const bool anypool = __any_sync(0xffffffffu, mypool != 0);
        if (anypool) {
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, pincl, d);
                if ((int)lane >= d) pincl += t;
            }
        }
        if (lane == 31u) {
            s_w[0][w] = (int)(incl & 255u);
            s_w[1][w] = pincl;
            s_w[2][w] = (int)((incl >> 8) & 255u);
            s_w[3][w] = (int)((incl >> 16) & 255u);
            s_w[4][w] = (int)(incl >> 24);
        }
        __syncthreads();
        if (threadIdx.x < 5) { // one thread per counter: exclusive scan over the warps + one global atomic
            int tot = 0;
            for (int i = 0; i < nw; i++) { const int v = s_w[threadIdx.x][i]; s_w[threadIdx.x][i] = tot; tot += v; }
            int *ctr = threadIdx.x == 0 ? &sc.ctr->ntouched : threadIdx.x == 1 ? &sc.ctr->pool : threadIdx.x == 2 ? &sc.ctr->nlarge
                     : threadIdx.x == 3 ? &sc.ctr->nlong : &sc.ctr->total;
            s_b[threadIdx.x] = tot ? atomicAdd(ctr, tot) : 0;
        }
        __syncthreads();
        const uint32_t excl = incl - packed;
        int tslot = s_b[0] + s_w[0][w] + (int)(excl & 255u);                      // next touched slot of this thread
        int ppos = 1 + s_b[1] + s_w[1][w] + (pincl - mypool);                     // next pool offset (0 is reserved)
        int lpos = s_b[2] + s_w[2][w] + (int)((excl >> 8) & 255u);                // next large-list slot
        int gpos = s_b[3] + s_w[3][w] + (int)((excl >> 16) & 255u);               // next long-list slot
        // ---- C0: rank-0 points publish the cell's touched slot ----------------------------------------------
        int slot[U], myp[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            slot[u] = -1;
            myp[u] = 0;
            if (rank[u] == 0) {
                slot[u] = tslot++;
                sc.touched[slot[u]] = key[u];
                st_release(&cells[key[u]].bin[par].y, slot[u] + 1);
            }
            if (lvl[u]) { myp[u] = ppos; ppos += level_cap(lvl[u]) + 1; }
        }
        __syncwarp(); // every lane's slot publication is issued before any lane of the warp starts waiting
        // ---- C1..: chunk pointers, level by level ---------------------------------------------------------------
        int maxl = 0;
#pragma unroll
        for (int u = 0; u < U; u++) maxl = max(maxl, lvl[u]);
        for (int l = 1; l <= maxl; l++) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                if (lvl[u] != l) continue;
                slot[u] = spin_nonzero(&cells[key[u]].bin[par].y) - 1;
                sc.pool[myp[u]] = make_uint4(0u, 0u, 0u, 0u); // fresh header: "next level not published"
                if (l == 1) {
                    sc.tlarge[lpos++] = make_int4(key[u], slot[u], myp[u], 0);
                    st_release(&sc.ovf1[slot[u]], myp[u]);
                } else {
                    const int p1 = spin_nonzero(&sc.ovf1[slot[u]]);
                    int q = p1;
                    for (int k = 2; k < l; k++) q = spin_nonzero(reinterpret_cast<const int *>(&sc.pool[q]));
                    if (l == 2) sc.tlong[gpos++] = make_int4(key[u], slot[u], p1, myp[u]);
                    st_release(reinterpret_cast<int *>(&sc.pool[q]), myp[u]);
                }
            }
        }
        // ---- D: store the records ------------------------------------------------------------------------------
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (rank[u] < 0) continue;
            if (slot[u] < 0) slot[u] = spin_nonzero(&cells[key[u]].bin[par].y) - 1;
            uint4 *dst;
            if (rank[u] < CHUNK0) {
                dst = sc.chunk0 + (size_t)CHUNK0 * slot[u] + rank[u];
            } else {
                const int j = level_of(rank[u]);
                int q = (lvl[u] == 1) ? myp[u] : spin_nonzero(&sc.ovf1[slot[u]]);
                for (int k = 2; k <= j; k++) q = (lvl[u] == k) ? myp[u] : spin_nonzero(reinterpret_cast<const int *>(&sc.pool[q]));
                dst = sc.pool + q + 1 + (rank[u] - level_base(j));
            }
            *dst = rec[u];
        }
        if (base + U * nthreads < n) __syncthreads(); // keep the block's threads in the same iteration (see above)
    }
}

__device__ __forceinline__ void zero_next_counters(const BinScratch &sc, int tid)
{
    if (tid < (int)(sizeof(BinCounters) / sizeof(int))) ((int *)sc.ctr_next)[tid] = 0; // 192 ints: the bin grids have >= 256 threads
}

template <int SRC, int U>
__global__ void __launch_bounds__(ADD_BLOCK)
k_bin(MapGeom g, MapLayers ml, FrameParams f, BinSource in, int n, BinScratch sc, RegionOps ro, int point_blocks,
      const __grid_constant__ SegTable segs, const FrameParams *frames)
{
    if ((int)blockIdx.x < point_blocks) {
        zero_next_counters(sc, blockIdx.x * blockDim.x + threadIdx.x);
        bin_points<SRC, U>(g, f, in, n, ml.cell, sc, blockIdx.x * blockDim.x + threadIdx.x, point_blocks * blockDim.x,
                           frames ? &segs : nullptr, frames);
    } else { // extra blocks: deferred scroll clears + variance floor (only when no fold is in flight, see gem_api.cu)
        const size_t rb = gridDim.x - point_blocks;
        phase_regions(g, ml, ro, (size_t)(blockIdx.x - point_blocks) * blockDim.x + threadIdx.x, rb * blockDim.x);
    }
}

// =========================================================================================
// fold
// =========================================================================================
// where the fold finds the intensity of the point a cell finally takes its colour from: the record carries the
// point index, not the 4-byte intensity (a 16-byte record is one vector store and a cell's first 8 records are
// one 128-byte line); the input array is read once per cell instead
struct FoldSrc {
    int kind;        // SRC_* of the call; SRC_SOA = no attributes (Process_points only updates `lowest`)
    const void *a;   // xyzi / pcl / intensity array / RouteRec array
};
__device__ __forceinline__ float fetch_intensity(const FoldSrc &s, uint32_t idx)
{
    switch (s.kind) {
    case SRC_XYZI: return reinterpret_cast<const float4 *>(s.a)[idx].w;
    case SRC_PCL32: return reinterpret_cast<const float4 *>(s.a)[2 * (size_t)idx + 1].z;
    case SRC_KEYS: return s.a ? reinterpret_cast<const float *>(s.a)[idx] : 0.0f;
    case SRC_RECORDS: return reinterpret_cast<const RouteRec *>(s.a)[idx].intensity;
    default: return 0.0f;
    }
}

struct CellState {
    float elev, var;
    uint32_t src;  // point index the cell last took intensity + colour from
    uint32_t rgb;
    bool ci_dirty;
    float minh, minhv; // lowest-scan: min height and variance of the first point attaining it
    bool any;
    float low_old;     // lowest[cell] before this call, fetched with the cell state (off the tail of the cell)
};

// lowest-scan of gpu.cu:432-438 (ORACLE DEFINITION): running minimum height of the call's points in
// the cell and the variance of the FIRST index attaining it.  Records must be offered in index order.
__device__ __forceinline__ void lowest_step(CellState &s, float h, float v)
{
    if (!s.any || h < s.minh) {
        s.minh = h;
        s.minhv = v;
        s.any = true;
    }
}

__device__ __forceinline__ void fold_step(CellState &s, float h, float v, uint32_t rgb, uint32_t idx, bool do_fuse)
{
    if (!do_fuse) return;
    const bool skip = (h == -1.0f); // gpu.cu:482
    const bool colour_ok = (rgb & REC_COLOUR_OK) != 0u; // gpu.cu:488, precomputed by the bin kernel
    const bool first = (s.elev == -10.0f); // gpu.cu:484
    // gpu.cu:500-501: `var < 0.0001` compares in double; (float)0.0001 is the largest float below
    // the double literal, so the test is exactly `var <= 1e-4f`
    const float ov = (s.var <= 1e-4f) ? 1e-4f : s.var;
    const float oe = s.elev;
    // gpu.cu:502-504: gate = RN(|h-e| / RN(sqrt(var))) > 5.  The fold is a serial dependency
    // chain per cell, so the IEEE sqrt and divide are kept off it: the two roundings move the
    // quotient by < 2.5e-7 relative, hence comparing d^2 with 25*var decides every case outside
    // a +-1e-5 band exactly like the reference expression; inside the band, and for huge or
    // non-finite values, the literal expression is evaluated.
    const float d = fabsf(h - oe);
    const float dd = d * d, tv = 25.0f * ov;
    const bool hi = dd > tv * 1.00001f, lo = dd < tv * 0.99999f;
    bool gate = hi;
    if (!(dd < 1e30f && tv < 1e30f) || !(hi || lo)) gate = (d / sqrtf(ov)) > 5.0f; // rare
    // gpu.cu:518-519, computed speculatively (selected below)
    float qe, qv;
    div2_rn(ov * h + v * oe, v * ov, ov + v, qe, qv);
    const bool higher = oe < h; // gpu.cu:505
    const float ne = first ? h : (gate ? (higher ? h : oe) : qe);
    const float nv = first ? v : (gate ? (higher ? v : ov) : qv);
    const bool take = first || !gate || higher;
    if (!skip) {
        s.elev = ne;
        s.var = nv;
        if (take && colour_ok) {
            s.src = idx;
            s.rgb = rgb & 0xffffffu;
            s.ci_dirty = true;
        }
    }
}

// Branch-free twin of fold_step for the serial tail of long lists.  Same arithmetic, but no control flow inside the
// step: one warp folding one cell is in-order, so every branch of fold_step (gate band, division guard) puts the
// elevation-dependent gate chain IN FRONT of the variance-dependent reciprocal chain instead of beside it.  Here the
// step always takes the common path and only reports (returns true) when fold_step would have left it: gate inside
// the +-1e-5 band or non-finite, or division operands outside the guarded range.  The caller then redoes the chunk
// with fold_step from the saved state, so results are fold_step's bit for bit.
__device__ __forceinline__ bool fold_step_fast(CellState &s, float h, float v, uint32_t rgb, uint32_t idx)
{
    const bool skip = (h == -1.0f);
    const bool colour_ok = (rgb & REC_COLOUR_OK) != 0u;
    const bool first = (s.elev == -10.0f);
    const float ov = (s.var <= 1e-4f) ? 1e-4f : s.var;
    const float oe = s.elev;
    const float d = fabsf(h - oe);
    const float dd = d * d, tv = 25.0f * ov;
    const bool hi = dd > tv * 1.00001f, lo = dd < tv * 0.99999f;
    const bool rare_gate = !(dd < 1e30f && tv < 1e30f) | !(hi | lo);
    const float n0 = ov * h + v * oe, n1 = v * ov, den = ov + v;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
    const float t = __fmaf_rn(-den, r, 1.0f);
    r = __fmaf_rn(r, t, r);
    const float p0 = __fmaf_rn(n0, r, 0.0f), p1 = __fmaf_rn(n1, r, 0.0f);
    const float e0 = __fmaf_rn(-den, p0, n0), e1 = __fmaf_rn(-den, p1, n1);
    const float qe = __fmaf_rn(r, e0, p0), qv = __fmaf_rn(r, e1, p1);
    const bool rare_div = !div2_fast_ok(n0, n1, den);
    const bool higher = oe < h;
    const float ne = first ? h : (hi ? (higher ? h : oe) : qe);
    const float nv = first ? v : (hi ? (higher ? v : ov) : qv);
    const bool take = (first | !hi | higher) & colour_ok & !skip;
    s.elev = skip ? s.elev : ne;
    s.var = skip ? s.var : nv;
    s.src = take ? idx : s.src;
    s.rgb = take ? (rgb & 0xffffffu) : s.rgb;
    s.ci_dirty = s.ci_dirty | take;
    return !skip & !first & (rare_gate | (!hi & rare_div));
}

__device__ __forceinline__ void cell_begin(CellState &s, const MapGeom &g, const MapLayers &ml, int key, bool do_lowest)
{
    s.low_old = do_lowest ? ml.lowest[key_to_lowest(g, key)] : 0.0f;
    const float2 ev = load_ev(ml.cell, key);
    s.elev = ev.x; s.var = ev.y; s.src = 0u; s.rgb = 0u; s.ci_dirty = false;
    s.minh = 0.0f; s.minhv = 0.0f; s.any = false;
}

// is the cell inside a scroll clear that the NEXT add call's Move has already decided?  (pipelined mode: this fold
// runs concurrently with the next call's bin kernel and carries that call's row / column clears, see gem_api.cu)
__device__ __forceinline__ bool in_clear_region(const MapGeom &g, const RegionOps &ro, int key)
{
    if (ro.count == 0) return false;
    const int row = key / g.cols, col = key - row * g.cols;
    bool hit = false;
    for (int r = 0; r < ro.count; r++) {
        const RegionOp op = ro.op[r];
        if (op.kind == 1) hit |= (row >= op.start && row < op.start + op.n);
        else if (op.kind == 2) hit |= (col >= op.start && col < op.start + op.n);
    }
    return hit;
}

__device__ __forceinline__ void cell_end(CellState &s, const MapGeom &g, const MapLayers &ml, const BinScratch &sc, const FoldSrc &src,
                                         const RegionOps &ro_next, int key, bool do_fuse, bool do_lowest)
{
    if (do_fuse) {
        if (in_clear_region(g, ro_next, key)) {
            // the cell scrolls out before anything can observe this fold: write what the clear writes (the region
            // blocks of this launch store the same bits, in either order)
            store_ev(ml.cell, key, make_float2(-10.0f, (float)0.0001));
            store_ci(ml.cell, key, make_uint2(0u, 0u));
        } else {
            if (s.var <= 1e-4f) s.var = 1e-4f; // gpu.cu:533-534 (same double-compare equivalence)
            store_ev(ml.cell, key, make_float2(s.elev, s.var));
            if (s.ci_dirty) store_ci(ml.cell, key, make_uint2(__float_as_uint(fetch_intensity(src, s.src)), s.rgb));
        }
    }
    if (do_lowest && s.any) {
        // ORACLE DEFINITION of the racy gpu.cu:434-438: with m = min h of this
        // call's points in the cell and i* the first index attaining it,
        // lowest = m + 3*hv[i*] iff m <= lowest_old.
        if (s.minh <= s.low_old) ml.lowest[key_to_lowest(g, key)] = s.minh + 3.0f * s.minhv;
    }
    ml.cell[key].bin[sc.par] = make_int2(0, 0); // restore the all-zero invariant of this parity's {counter, slot}
}

// short lists (k <= 8): one thread per touched slot, records of chunk 0 held in registers, selection in index order
__device__ __forceinline__ void fold_small(const MapGeom &g, const MapLayers &ml, const BinScratch &sc, const FoldSrc &src,
                                           const RegionOps &ro_next, bool do_fuse, bool do_lowest, int tid, int nthreads)
{
    const int nt = sc.ctr->ntouched;
    for (int j = tid; j < nt; j += nthreads) {
        const int key = sc.touched[j];
        const uint4 *c0 = sc.chunk0 + (size_t)CHUNK0 * j;
        // the cell's record and its first records are fetched together (the slot IS j): one round trip
        const int k = ml.cell[key].bin[sc.par].x;
        uint4 rr[CHUNK0];
#pragma unroll
        for (int e = 0; e < 2; e++) rr[e] = c0[e];
        if (k > CHUNK0) continue; // folded by a warp (fold_large)
        CellState s;
        cell_begin(s, g, ml, key, do_lowest);
#pragma unroll
        for (int e = 2; e < CHUNK0; e++) {
            rr[e] = make_uint4(0u, 0u, 0u, 0u);
            if (e < k) rr[e] = c0[e];
        }
        int last = -1;
        for (int it = 0; it < k; it++) {
            int best = 0x7fffffff;
            uint4 b = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int e = 0; e < CHUNK0; e++) {
                const int ie = (e < k) ? (int)rr[e].x : 0x7fffffff;
                const bool c = ie > last && ie < best;
                if (c) { best = ie; b = rr[e]; }
            }
            lowest_step(s, __uint_as_float(b.y), __uint_as_float(b.z));
            fold_step(s, __uint_as_float(b.y), __uint_as_float(b.z), b.w, b.x, do_fuse);
            last = best;
        }
        cell_end(s, g, ml, sc, src, ro_next, key, do_fuse, do_lowest);
    }
}

constexpr int FOLD_KMAX = 1024;  // list length one warp sorts in shared memory
constexpr int FOLD_SLOT_BITS = 10; // sort key = (point index << 10) | rank: needs index < 2^22
constexpr int FOLD_INDEX_BITS = 32 - FOLD_SLOT_BITS; // = the largest launch (gem_create caps max_points)

// Warp-wide bitonic sort of 32*R keys held in registers: element i lives in lane i%32,
// register i/32.  Partners less than 32 apart are exchanged with one shuffle, the rest are in
// the same lane.  (A shared-memory network costs ~450 cycles per stage on B200, a shuffle
// stage ~30.)
template <int R>
__device__ __forceinline__ void warp_bitonic(uint32_t (&key)[R], unsigned lane)
{
#pragma unroll
    for (int size = 2; size <= 32 * R; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {
                const int rs = stride >> 5;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if ((r & rs) == 0) {
                        const uint32_t a = key[r], b2 = key[r | rs];
                        const bool up = (((int)lane + 32 * r) & size) == 0;
                        const bool sw = (a > b2) == up;
                        key[r] = sw ? b2 : a;
                        key[r | rs] = sw ? a : b2;
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const uint32_t a = key[r];
                    const uint32_t o = __shfl_xor_sync(0xffffffffu, a, stride);
                    const bool up = (((int)lane + 32 * r) & size) == 0;
                    const bool lower = ((int)lane & stride) == 0;
                    const uint32_t mn = min(a, o), mx = max(a, o);
                    key[r] = (lower == up) ? mn : mx;
                }
            }
        }
    }
}

// order-preserving map float -> uint32 (for a warp min-reduction); -0 is folded onto +0
__device__ __forceinline__ uint32_t float_order_key(float f)
{
    const uint32_t u = __float_as_uint(f == 0.0f ? 0.0f : f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// fold 32 records (one per lane, already in index order: r = {idx, h, var, rgb}) into the cell state
__device__ __forceinline__ void fold_chunk(CellState &s, const uint4 &r, int m, bool do_fuse)
{
    {   // lowest-scan of the chunk, off the serial chain: warp minimum of h, first lane attaining it
        // (lanes hold the records in index order), then the same strict-< update as lowest_step
        const unsigned lane = threadIdx.x & 31u;
        const uint32_t k = ((int)lane < m) ? float_order_key(__uint_as_float(r.y)) : 0xffffffffu;
        const uint32_t kmin = __reduce_min_sync(0xffffffffu, k);
        const int src = __ffs(__ballot_sync(0xffffffffu, k == kmin)) - 1;
        const float ch = __uint_as_float(__shfl_sync(0xffffffffu, r.y, src));
        const float cv = __uint_as_float(__shfl_sync(0xffffffffu, r.z, src));
        lowest_step(s, ch, cv);
    }
    if (!do_fuse) return;
    const CellState s0 = s;
    // broadcast record t+1 while record t is folded (in-order issue: keeps the shuffle latency
    // off the serial chain)
    uint32_t nh = __shfl_sync(0xffffffffu, r.y, 0), nv = __shfl_sync(0xffffffffu, r.z, 0);
    uint32_t nc = __shfl_sync(0xffffffffu, r.w, 0), ni = __shfl_sync(0xffffffffu, r.x, 0);
    bool rare = false;
    for (int t = 0; t < m; t++) {
        const float h = __uint_as_float(nh), v = __uint_as_float(nv);
        const uint32_t rgb = nc, idx = ni;
        const int tn = (t + 1) & 31;
        nh = __shfl_sync(0xffffffffu, r.y, tn);
        nv = __shfl_sync(0xffffffffu, r.z, tn);
        nc = __shfl_sync(0xffffffffu, r.w, tn);
        ni = __shfl_sync(0xffffffffu, r.x, tn);
        rare |= fold_step_fast(s, h, v, rgb, idx);
    }
    if (__any_sync(0xffffffffu, rare)) { // some step left the common path: redo the chunk literally
        s = s0;
        for (int t = 0; t < m; t++) {
            const float h = __uint_as_float(__shfl_sync(0xffffffffu, r.y, t)), v = __uint_as_float(__shfl_sync(0xffffffffu, r.z, t));
            const uint32_t rgb = __shfl_sync(0xffffffffu, r.w, t), idx = __shfl_sync(0xffffffffu, r.x, t);
            fold_step(s, h, v, rgb, idx, true);
        }
    }
}

// chunk pointers of one cell, warp-uniform.  p[j] = pool offset of the level-j chunk (j >= 1), t = touched slot.
struct ChunkRefs {
    int t;
    int p[5]; // levels 1..4 cached (ranks < 2728); deeper levels are walked
};
__device__ __forceinline__ const uint4 *record_ptr(const BinScratch &sc, const ChunkRefs &c, int rank)
{
    if (rank < CHUNK0) return sc.chunk0 + (size_t)CHUNK0 * c.t + rank;
    if (rank < level_base(2)) return sc.pool + c.p[1] + 1 + (rank - level_base(1));
    if (rank < level_base(3)) return sc.pool + c.p[2] + 1 + (rank - level_base(2));
    if (rank < level_base(4)) return sc.pool + c.p[3] + 1 + (rank - level_base(3));
    if (rank < level_base(5)) return sc.pool + c.p[4] + 1 + (rank - level_base(4));
    int j = 5, q = (int)sc.pool[c.p[4]].x; // very long lists: walk the chain
    while (rank >= level_base(j + 1)) { q = (int)sc.pool[q].x; j++; }
    return sc.pool + q + 1 + (rank - level_base(j));
}

// sort a list of k <= 32*R records in registers and fold it.  The records are read ONCE, coalesced in rank order,
// together with the sort keys: the record goes to the warp's shared scratch (16 B x 256 slots = the 4 KB of s_key),
// the keys (index << 10 | rank) are sorted in registers, and each chunk then picks its records by rank from shared
// memory instead of a second dependent global gather.
template <int R>
__device__ __forceinline__ void fold_list_regs(CellState &s, const BinScratch &sc, const ChunkRefs &c, int k, unsigned lane,
                                               bool do_fuse, uint32_t *s_key)
{
    static_assert(32 * R * 16 <= FOLD_KMAX * 4, "records of a register-sorted list must fit the warp's scratch");
    uint4 *s_rec = reinterpret_cast<uint4 *>(s_key);
    uint32_t key[R];
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int e = (int)lane + 32 * r;
        key[r] = 0xffffffffu;
        if (e < k) {
            const uint4 rec = *record_ptr(sc, c, e);
            key[r] = (rec.x << FOLD_SLOT_BITS) | (uint32_t)e;
            s_rec[e] = rec;
        }
    }
    __syncwarp();
    warp_bitonic<R>(key, lane);
#pragma unroll
    for (int r = 0; r < R; r++) {
        const int c0 = 32 * r;
        if (c0 < k) {
            uint4 rec = make_uint4(0, 0, 0, 0);
            if (c0 + (int)lane < k) rec = s_rec[key[r] & ((1u << FOLD_SLOT_BITS) - 1u)];
            fold_chunk(s, rec, min(32, k - c0), do_fuse);
        }
    }
    __syncwarp(); // the scratch is reused by this warp's next cell
}

// long lists: one warp per cell.  s_key: per-warp shared scratch of FOLD_KMAX words.
// Cells are dealt to warps statically, the long lists (k > 40) first so that their serial chains start with the
// first wave of blocks, and in boustrophedon order over the rounds so that a warp that drew a long list in one
// round draws from the short end in the next.  (Measured on B200: a ticket counter instead of the static deal costs
// an atomic round trip per cell and is slower.)
__device__ __forceinline__ void fold_large(const MapGeom &g, const MapLayers &ml, const BinScratch &sc, const FoldSrc &src,
                                           const RegionOps &ro_next, bool do_fuse, bool do_lowest, uint32_t *s_key, int gwarp, int nwarps)
{
    const int nlong = sc.ctr->nlong;
    const int nl = nlong + sc.ctr->nlarge;
    const unsigned lane = threadIdx.x & 31u;
    for (int round = 0; round * nwarps < nl; round++) {
        const int j = round * nwarps + ((round & 1) ? nwarps - 1 - gwarp : gwarp);
        if (j >= nl) continue;
        const int4 info = j < nlong ? sc.tlong[j] : sc.tlarge[j - nlong];
        const int key = info.x;
        const int k = ml.cell[key].bin[sc.par].x;
        if (j >= nlong && k > FOLD_LONG_FROM) continue; // also on the long list: folded from there
        ChunkRefs c;
        c.t = info.y; c.p[1] = info.z; c.p[2] = info.w; c.p[3] = 0; c.p[4] = 0;
        if (k > level_base(3)) c.p[3] = (int)sc.pool[c.p[2]].x;
        if (k > level_base(4)) c.p[4] = (int)sc.pool[c.p[3]].x;
        CellState s;
        cell_begin(s, g, ml, key, do_lowest);
        // order the records by point index (== the visiting order of G_fuse's per-cell loop)
        if (k <= 32) fold_list_regs<1>(s, sc, c, k, lane, do_fuse, s_key);
        else if (k <= 64) fold_list_regs<2>(s, sc, c, k, lane, do_fuse, s_key);
        else if (k <= 128) fold_list_regs<4>(s, sc, c, k, lane, do_fuse, s_key);
        else if (k <= 256) fold_list_regs<8>(s, sc, c, k, lane, do_fuse, s_key);
        else if (k <= FOLD_KMAX) {
            // bitonic sort of packed (index, rank) keys in shared memory
            int P = 512;
            while (P < k) P <<= 1;
            for (int e = (int)lane; e < P; e += 32)
                s_key[e] = (e < k) ? ((record_ptr(sc, c, e)->x << FOLD_SLOT_BITS) | (uint32_t)e) : 0xffffffffu;
            __syncwarp();
            for (int size = 2; size <= P; size <<= 1) {
                for (int stride = size >> 1; stride > 0; stride >>= 1) {
                    for (int i = (int)lane; i < (P >> 1); i += 32) {
                        const int lo = ((i & ~(stride - 1)) << 1) | (i & (stride - 1));
                        const int hi = lo + stride;
                        const uint32_t a = s_key[lo], b2 = s_key[hi];
                        const bool up = (lo & size) == 0;
                        if ((a > b2) == up) { s_key[lo] = b2; s_key[hi] = a; }
                    }
                    __syncwarp();
                }
            }
            for (int c0 = 0; c0 < k; c0 += 32) {
                const int sidx = c0 + (int)lane;
                uint4 r = make_uint4(0, 0, 0, 0);
                if (sidx < k) r = *record_ptr(sc, c, (int)(s_key[sidx] & ((1u << FOLD_SLOT_BITS) - 1u)));
                fold_chunk(s, r, min(32, k - c0), do_fuse);
            }
            __syncwarp();
        } else {
            // very long lists: repeated selection of the next smallest index from global memory
            uint32_t last = 0;
            bool have_last = false;
            for (int it = 0; it < k; it++) {
                uint32_t best = 0xffffffffu;
                int beste = -1;
                for (int e = (int)lane; e < k; e += 32) {
                    const uint32_t v = record_ptr(sc, c, e)->x;
                    if ((!have_last || v > last) && v < best) { best = v; beste = e; }
                }
                const uint32_t wbest = __reduce_min_sync(0xffffffffu, best);
                const unsigned who = __ballot_sync(0xffffffffu, best == wbest && beste >= 0);
                const int srcl = __ffs(who) - 1;
                const int e = __shfl_sync(0xffffffffu, beste, srcl);
                const uint4 r = *record_ptr(sc, c, e);
                lowest_step(s, __uint_as_float(r.y), __uint_as_float(r.z));
                fold_step(s, __uint_as_float(r.y), __uint_as_float(r.z), r.w, r.x, do_fuse);
                last = wbest;
                have_last = true;
            }
        }
        if (lane == 0u) {
            cell_end(s, g, ml, sc, src, ro_next, key, do_fuse, do_lowest);
            sc.ovf1[c.t] = 0; // restore the all-zero invariant of the level-1 pointers
            atomicMax(&sc.ctr->maxk, k);
        }
    }
}

// fold_blocks blocks fold; blocks beyond them execute `ro` (row / column clears of the NEXT call's Move, pipelined
// mode only: a cell inside such a region is written with the cleared value by whoever touches it, see cell_end)
__global__ void __launch_bounds__(ADD_BLOCK)
k_fold(MapGeom g, MapLayers ml, BinScratch sc, FoldSrc src, RegionOps ro, int fold_blocks, int do_fuse, int do_lowest)
{
    __shared__ __align__(16) uint32_t s_key[ADD_BLOCK / 32][FOLD_KMAX];
    if ((int)blockIdx.x >= fold_blocks) {
        const size_t rb = gridDim.x - fold_blocks;
        phase_regions(g, ml, ro, (size_t)(blockIdx.x - fold_blocks) * blockDim.x + threadIdx.x, rb * blockDim.x);
        return;
    }
    const int w = threadIdx.x >> 5;
    const int gw = blockIdx.x * (ADD_BLOCK / 32) + w, nw = fold_blocks * (ADD_BLOCK / 32);
    // long lists first (they are the critical path), then the short ones
    fold_large(g, ml, sc, src, ro, do_fuse != 0, do_lowest != 0, s_key[w], gw, nw);
    // the warps that drew a list longer than 40 ARE the tail of this kernel: they sit the short lists out
    const int nlong = min(sc.ctr->nlong, nw / 2);
    if (gw >= nlong)
        fold_small(g, ml, sc, src, ro, do_fuse != 0, do_lowest != 0, (gw - nlong) * 32 + (threadIdx.x & 31), (nw - nlong) * 32);
}

} // namespace gem
