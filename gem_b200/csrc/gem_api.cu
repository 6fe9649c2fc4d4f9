// gem_api.cu -- host side of libgem_b200.so: the extern "C" ABI of include/gem_b200.h.
//
// Replaces the host wrappers of the reference's gpu_process.cu (Init_GPU_elevationmap :940,
// Move :1004, Process_points :1085, Mapvar_update :1146, Fuse :1154, Map_optmove :1215,
// Map_closeloop :1235, Map_feature :1256, Raytracing :1304).  Differences by design:
// per-handle state instead of __device__ globals, one stream per handle, zero per-call
// cudaMalloc/cudaFree (the reference does 8+7+9 per frame), geometry passed as kernel
// parameters instead of cudaMemcpyTo/FromSymbol round trips, int status codes.
#include <cuda_runtime.h>

#include <algorithm>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gem_b200.h"
#include "gem_kernels.cuh"
#include "gem_route.cuh"

using namespace gem;

namespace {

thread_local std::string g_create_error;

} // namespace

struct gem_map {
    gem_config cfg{};
    int dev = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int L = 0;
    size_t nc = 0; // cells held by this handle
    int P = 0;     // per-launch point capacity
    MapGeom geom{};
    MapLayers ml{};
    Scratch sc{};
    float sensorZ = 0.0f;
    // deferred region operations (scroll clears of Move, the every-cell variance floor of
    // G_fuse): executed by the next add/fuse launch, or flushed before anything observes the map
    std::vector<RegionOp> pending;
    Counters *ctr_buf[2] = {nullptr, nullptr};
    int ctr_cur = 0;          // which counter buffer the NEXT call uses (it is zero)
    Counters *ctr_last = nullptr; // counters of the last finished call
    bool pdl = false;         // programmatic dependent launch between the add-path kernels (opt-in)
    bool pdl_front = false;   // stream mode only: PDL between transform -> alloc -> scatter on the front stream (opt-in)
    int coop_blocks = 0;      // co-resident grid size of the fused kernel (0 = unavailable)
    int fused_max_points = 1 << 20;
    // staging (device), lazily allocated
    void *d_xyzi = nullptr, *d_rgba = nullptr, *d_pcl = nullptr;
    float *d_x = nullptr, *d_y = nullptr, *d_z = nullptr, *d_xt = nullptr, *d_yt = nullptr;
    int *d_keyin = nullptr, *d_R = nullptr, *d_G = nullptr, *d_B = nullptr;
    float *d_int = nullptr;
    float *d_out = nullptr; // 9 * nc floats read-out staging
    int *d_owner_cnt = nullptr;
    float2 *prev_ev = nullptr;     // gem_snapshot_shown: prevMap_ (ElevationMapping.cpp:422) on the device
    uint2 *prev_ci = nullptr;
    float *prev_tr = nullptr;
    MapGeom prev_geom{};
    bool prev_valid = false;
    int *d_viscnt = nullptr;       // visual-cloud export: per (column, row chunk) counts / offsets
    uint32_t *d_gbitmap = nullptr; // tiled ray clean-up: validity bitmap of the map-wide lowest layer
    Counters *h_ctr = nullptr; // pinned
    // pipelined host ingest (gem_add_points_host_async)
    cudaStream_t copy_stream = nullptr;
    void *d_axyzi[2] = {nullptr, nullptr}, *d_argba[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr};
    Counters *h_ctr_ring = nullptr; // pinned, 2 entries
    unsigned async_calls = 0;
    // gem_add_points_stream: frame-pipelined mode (own scratch sets, front stream, events)
    bool pipe_ready = false;
    // frame pipeline of gem_add_points_stream: three stages on three streams (transform+bin | alloc+scatter | fold),
    // three scratch sets (one per frame in flight), four counter buffers (the transform kernel of frame i clears
    // the buffer of frame i+1, last used by frame i-3)
    Scratch pipe_sc[3];
    Counters *pipe_ctr[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaStream_t front_stream = nullptr, mid_stream = nullptr;
    bool mid_owned = false;
    cudaEvent_t ev_bin[3] = {nullptr, nullptr, nullptr}, ev_front[3] = {nullptr, nullptr, nullptr}, ev_fold[3] = {nullptr, nullptr, nullptr};
    unsigned pipe_calls = 0;
    // gem_add_points_multi: ring of per-call FrameParams tables (pinned host + device)
    FrameParams *h_frames = nullptr, *d_frames = nullptr;
    cudaEvent_t ev_frames[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned multi_calls = 0;
    gem_stats stats{};
    std::string err;
    std::vector<void *> allocs;
    // launch accounting / optional per-kernel CUDA-event timing (gem_profile_*)
    long long launches = 0;
    bool profiling = false;
    struct Span { int cls; cudaEvent_t e0, e1; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> free_events;
    double prof_ms[GEM_PROF_CLASSES] = {0};
    long long prof_count[GEM_PROF_CLASSES] = {0};
};

namespace {

int fail(gem_map *m, int code, const std::string &msg)
{
    if (m) m->err = msg; else g_create_error = msg;
    return code;
}

cudaEvent_t prof_event(gem_map *m)
{
    if (!m->free_events.empty()) {
        cudaEvent_t e = m->free_events.back();
        m->free_events.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}
// every kernel launch of the library goes through this macro: counts the launch and, when
// profiling is on, brackets it with CUDA events on the handle's stream
#define GEM_LAUNCH_ON(m, st, cls, ...)                           \
    do {                                                         \
        (m)->launches++;                                         \
        if ((m)->profiling) {                                    \
            gem_map::Span sp__{(cls), prof_event(m), prof_event(m)}; \
            cudaEventRecord(sp__.e0, (st));                      \
            __VA_ARGS__;                                         \
            cudaEventRecord(sp__.e1, (st));                      \
            (m)->spans.push_back(sp__);                          \
        } else {                                                 \
            __VA_ARGS__;                                         \
        }                                                        \
    } while (0)
#define GEM_LAUNCH(m, cls, ...)                                  \
    do {                                                         \
        (m)->launches++;                                         \
        if ((m)->profiling) {                                    \
            gem_map::Span sp__{(cls), prof_event(m), prof_event(m)}; \
            cudaEventRecord(sp__.e0, (m)->stream);               \
            __VA_ARGS__;                                         \
            cudaEventRecord(sp__.e1, (m)->stream);               \
            (m)->spans.push_back(sp__);                          \
        } else {                                                 \
            __VA_ARGS__;                                         \
        }                                                        \
    } while (0)

// launch with programmatic stream serialization (PDL); falls back to a plain launch when disabled
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

#define GEM_CUDA(m, expr)                                                                      \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            return fail((m), GEM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
        }                                                                                      \
    } while (0)

template <typename T> int dev_alloc(gem_map *m, T **p, size_t count)
{
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) + 16);
    if (e != cudaSuccess) return fail(m, GEM_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    m->allocs.push_back(q);
    *p = (T *)q;
    return GEM_OK;
}

// capacity of a per-call list of the cells holding MORE than k of a call's <= P records: at most P / (k + 1) such
// cells exist, and never more than the map has
inline size_t list_cap(size_t P, size_t nc, int k) { return std::min(nc, P / (size_t)(k + 1)) + 1; }
inline int blocks_for(size_t n, int bs, int cap = 148 * 16)
{
    size_t b = (n + bs - 1) / bs;
    if (b < 1) b = 1;
    if ((size_t)cap < b) b = cap;
    return (int)b;
}

struct SetDev {
    int prev = -1;
    explicit SetDev(int d) { cudaGetDevice(&prev); if (prev != d) cudaSetDevice(d); else prev = -1; }
    ~SetDev() { if (prev >= 0) cudaSetDevice(prev); }
};

FrameParams make_frame(const gem_frame *f)
{
    FrameParams p;
    memset(&p, 0, sizeof p);
    for (int i = 0; i < 12; i++) p.T[i] = f->T[i];
    for (int i = 0; i < 3; i++) { p.sJ[i] = f->sensor_jacobian[i]; p.P[i] = f->P_mul_C_BM_transpose[i]; }
    p.has_rot = 0;
    for (int i = 0; i < 9; i++) {
        p.rotVar[i] = f->rotation_variance[i];
        p.CSBT[i] = f->C_SB_transpose[i];
        p.Bskew[i] = f->B_r_BS_skew[i];
        if (f->rotation_variance[i] != 0.0f) p.has_rot = 1;
    }
    p.lo = f->rel_lower;
    p.hi = f->rel_upper;
    p.sensor_type = f->sensor.type;
    p.min_r = f->sensor.min_radius;
    p.beam_a = f->sensor.beam_angle;
    p.beam_c = f->sensor.beam_constant;
    p.nf_a = f->sensor.normal_factor_a;
    p.nf_b = f->sensor.normal_factor_b;
    p.nf_c = f->sensor.normal_factor_c;
    p.nf_d = f->sensor.normal_factor_d;
    p.nf_e = f->sensor.normal_factor_e;
    p.lat = f->sensor.lateral_factor;
    p.cut_lo = (float)f->sensor.cutoff_min_depth; // pcl::PassThrough::setFilterLimits takes floats
    p.cut_hi = (float)f->sensor.cutoff_max_depth;
    return p;
}

size_t region_cells(const gem_map *m, const RegionOp &r)
{
    if (r.kind == 0) return m->nc;
    if (r.kind == 1) return (size_t)r.n * m->geom.cols;
    return (size_t)r.n * m->geom.rows;
}

int launch_regions(gem_map *m, const RegionOp *ops, int count)
{
    for (int i = 0; i < count; i += MAX_REGION_OPS) {
        RegionOps ro{};
        size_t cells = 0;
        ro.count = (count - i < MAX_REGION_OPS) ? (count - i) : MAX_REGION_OPS;
        for (int k = 0; k < ro.count; k++) { ro.op[k] = ops[i + k]; cells += region_cells(m, ops[i + k]); }
        GEM_LAUNCH(m, GEM_PROF_CLEAR, k_regions<<<blocks_for(cells, ADD_BLOCK), ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, ro));
    }
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

// something is about to read the layers: execute deferred clears now (a clear is visible in
// the reference as soon as Move returns); floors stay pending until the next Fuse
int flush_for_observer(gem_map *m)
{
    std::vector<RegionOp> now;
    for (RegionOp &r : m->pending)
        if (r.clear) {
            RegionOp c = r;
            c.floor_ = 0;
            now.push_back(c);
            r.clear = 0;
        }
    if (now.empty()) return GEM_OK;
    return launch_regions(m, now.data(), (int)now.size());
}

// a Fuse-type call is starting: hand the pending operations to its first kernel.  Returns the
// number of extra blocks that kernel should carry.
int take_region_ops(gem_map *m, RegionOps &ro, int &region_blocks)
{
    ro.count = 0;
    region_blocks = 0;
    const int np = (int)m->pending.size();
    if (np > MAX_REGION_OPS) { // rare: many moves without an add
        int rc = launch_regions(m, m->pending.data(), np - MAX_REGION_OPS);
        if (rc) return rc;
        m->pending.erase(m->pending.begin(), m->pending.end() - MAX_REGION_OPS);
    }
    size_t cells = 0;
    for (const RegionOp &r : m->pending) {
        ro.op[ro.count++] = r;
        cells += region_cells(m, r);
    }
    m->pending.clear();
    if (ro.count) region_blocks = blocks_for(cells, ADD_BLOCK * 4, 148 * 2);
    return GEM_OK;
}

// a Fuse with nothing to fold still applies clears + floor
int flush_all_pending(gem_map *m)
{
    if (m->pending.empty()) return GEM_OK;
    int rc = launch_regions(m, m->pending.data(), (int)m->pending.size());
    m->pending.clear();
    return rc;
}

void pend_all_floor(gem_map *m)
{
    m->pending.clear();
    m->pending.push_back(RegionOp{0, 0, 0, 0, 1});
}

Scratch cur_scratch(gem_map *m)
{
    Scratch sc = m->sc;
    sc.ctr = m->ctr_buf[m->ctr_cur];
    sc.ctr_next = m->ctr_buf[m->ctr_cur ^ 1];
    return sc;
}
void call_done(gem_map *m)
{
    m->ctr_last = m->ctr_buf[m->ctr_cur];
    m->ctr_cur ^= 1;
}

// points per thread in the transform / scatter kernels: one point per thread keeps a frame-sized call
// (1e5 points, < 1 wave) latency-optimal; large calls get 2 or 4 points per thread so that a thread has
// several independent DRAM/L2 round trips in flight instead of running 3-4 waves of serial chains
inline int points_per_thread(int n) { return n >= 600000 ? 4 : (n >= 250000 ? 2 : 1); }

// K2..K4 after the binning kernel of the current chunk
template <int ATTR>
int run_group_fold(gem_map *m, const Scratch &sc, const AttrInput &a, int n, bool do_fuse, bool do_lowest)
{
    GEM_LAUNCH(m, GEM_PROF_ALLOC, launch_pdl(m->pdl, k_alloc_cells, blocks_for((size_t)n, ADD_BLOCK, 148 * 4), ADD_BLOCK, m->stream, sc));
    const int U = points_per_thread(n);
    const int sb = blocks_for((size_t)(n + U - 1) / U, ADD_BLOCK, 148 * 16);
    if (U == 4) GEM_LAUNCH(m, GEM_PROF_SCATTER, launch_pdl(m->pdl, k_scatter<ATTR, 4>, sb, ADD_BLOCK, m->stream, a, n, sc));
    else if (U == 2) GEM_LAUNCH(m, GEM_PROF_SCATTER, launch_pdl(m->pdl, k_scatter<ATTR, 2>, sb, ADD_BLOCK, m->stream, a, n, sc));
    else GEM_LAUNCH(m, GEM_PROF_SCATTER, launch_pdl(m->pdl, k_scatter<ATTR, 1>, sb, ADD_BLOCK, m->stream, a, n, sc));
    GEM_LAUNCH(m, GEM_PROF_FOLD, launch_pdl(m->pdl, k_fold, blocks_for((size_t)n, ADD_BLOCK, 148 * 8), ADD_BLOCK, m->stream, m->geom, m->ml, sc, do_fuse ? 1 : 0, do_lowest ? 1 : 0));
    GEM_CUDA(m, cudaGetLastError());
    call_done(m);
    return GEM_OK;
}

int read_counters(gem_map *m, long long n_in, bool accumulate)
{
    GEM_CUDA(m, cudaMemcpyAsync(m->h_ctr, m->ctr_last, sizeof(Counters), cudaMemcpyDeviceToHost, m->stream));
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    if (!accumulate) memset(&m->stats, 0, sizeof m->stats);
    m->stats.points_in += n_in;
    m->stats.points_binned += m->h_ctr->total;
    m->stats.cells_touched += m->h_ctr->ntouched;
    int mk = m->h_ctr->maxk; // only lists longer than FOLD_SMALL_K are tracked exactly
    if (mk < 1 && m->h_ctr->ntouched > 0) mk = (m->h_ctr->total > m->h_ctr->ntouched) ? FOLD_SMALL_K : 1;
    if (mk > m->stats.max_points_per_cell) m->stats.max_points_per_cell = mk;
    return GEM_OK;
}

int ensure_host_staging(gem_map *m)
{
    if (m->d_xyzi) return GEM_OK;
    int rc;
    if ((rc = dev_alloc(m, (float4 **)&m->d_xyzi, (size_t)m->P))) return rc;
    if ((rc = dev_alloc(m, (uchar4 **)&m->d_rgba, (size_t)m->P))) return rc;
    return GEM_OK;
}
int ensure_pcl_staging(gem_map *m)
{
    if (m->d_pcl) return GEM_OK;
    return dev_alloc(m, (float4 **)&m->d_pcl, (size_t)m->P * 2);
}
int ensure_compat_staging(gem_map *m)
{
    if (m->d_x) return GEM_OK;
    int rc;
    const size_t P = (size_t)m->P;
    if ((rc = dev_alloc(m, &m->d_x, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_y, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_z, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_xt, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_yt, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_keyin, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_R, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_G, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_B, P))) return rc;
    if ((rc = dev_alloc(m, &m->d_int, P))) return rc;
    return GEM_OK;
}
int ensure_out_staging(gem_map *m)
{
    if (m->d_out) return GEM_OK;
    return dev_alloc(m, &m->d_out, m->nc * 9);
}

// one chunk of the fused path on device-resident input
template <int IN, int ATTR>
int add_chunk(gem_map *m, const PointInput &in, const AttrInput &a, int n, const FrameParams &fp)
{
    RegionOps ro;
    int rb = 0;
    int rc = take_region_ops(m, ro, rb);
    if (rc) return rc;
    const Scratch sc = cur_scratch(m);
    if (m->coop_blocks > 0 && n <= m->fused_max_points) {
        // frame-sized call: one cooperative launch, grid barriers between the phases
        MapGeom g = m->geom;
        MapLayers ml = m->ml;
        FrameParams f = fp;
        PointInput pin = in;
        AttrInput at = a;
        int nn = n, do_fuse = 1, do_lowest = 1;
        Scratch s2 = sc;
        void *args[] = {&g, &ml, &f, &pin, &at, &nn, &s2, &ro, &do_fuse, &do_lowest};
        int blocks = blocks_for((size_t)(n > 0 ? n : 1), ADD_BLOCK, m->coop_blocks);
        if (ro.count && blocks < m->coop_blocks) blocks = (blocks + rb < m->coop_blocks) ? blocks + rb : m->coop_blocks;
        GEM_LAUNCH(m, GEM_PROF_FUSED,
                   cudaLaunchCooperativeKernel((const void *)k_add_fused<IN, ATTR>, dim3(blocks), dim3(ADD_BLOCK), args, 0, m->stream));
        GEM_CUDA(m, cudaGetLastError());
        call_done(m);
        return GEM_OK;
    }
    const int U = points_per_thread(n);
    const int pb = blocks_for((size_t)(n + U - 1) / U, ADD_BLOCK, 148 * 16);
    if (U == 4)
        GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, launch_pdl(m->pdl, k_transform_bin<IN, 4>, pb + rb, ADD_BLOCK, m->stream, m->geom, m->ml, fp, in, n, sc, ro, pb, (float *)nullptr, (float *)nullptr));
    else if (U == 2)
        GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, launch_pdl(m->pdl, k_transform_bin<IN, 2>, pb + rb, ADD_BLOCK, m->stream, m->geom, m->ml, fp, in, n, sc, ro, pb, (float *)nullptr, (float *)nullptr));
    else
        GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, launch_pdl(m->pdl, k_transform_bin<IN, 1>, pb + rb, ADD_BLOCK, m->stream, m->geom, m->ml, fp, in, n, sc, ro, pb, (float *)nullptr, (float *)nullptr));
    return run_group_fold<ATTR>(m, sc, a, n, true, true);
}

} // namespace

// =========================================================================================
static GridMapFrame grid_frame(const gem_map *m, float cx, float cy, int sx, int sy)
{
    GridMapFrame f;
    f.res = m->cfg.grid_resolution > 0.0 ? m->cfg.grid_resolution : (double)m->cfg.resolution;
    f.half = 0.5 * ((double)m->L * f.res) - 0.5 * f.res;
    f.cx = (double)cx; f.cy = (double)cy;
    f.L = m->L; f.sx = sx; f.sy = sy;
    return f;
}

// count -> scan -> write of the cells Src takes, in GridMapIterator order; returns the total through *total_out
template <class Src> static int compact_cells(gem_map *m, const Src &src, int capacity, int *total_out)
{
    const int L = m->L, nch = (L + 31) / 32;
    int rc;
    const int n = L * nch, nseg = (n + SCAN_SEG - 1) / SCAN_SEG;
    if (!m->d_viscnt) { if ((rc = dev_alloc(m, &m->d_viscnt, (size_t)n + nseg + 1))) return rc; }
    int *d_segtot = m->d_viscnt + n, *d_total = d_segtot + nseg;
    const dim3 grid(nch, nch);
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_compact_count<Src><<<grid, 1024, 0, m->stream>>>(src, L, nch, m->d_viscnt));
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_compact_scan<<<nseg, SCAN_SEG, 0, m->stream>>>(m->d_viscnt, n, d_segtot));
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_compact_write<Src><<<grid, 1024, 0, m->stream>>>(src, L, nch, m->d_viscnt, d_segtot, nseg, d_total, capacity));
    GEM_CUDA(m, cudaGetLastError());
    GEM_CUDA(m, cudaMemcpyAsync(total_out, d_total, sizeof(int), cudaMemcpyDeviceToHost, m->stream));
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

extern "C" {

int gem_version(void) { return GEM_B200_VERSION; }

const char *gem_last_error(const gem_map *m) { return m ? m->err.c_str() : g_create_error.c_str(); }

int gem_create(const gem_config *cfg, gem_map **out)
{
    if (!cfg || !out) return fail(nullptr, GEM_ERR_INVALID, "gem_create: null argument");
    *out = nullptr;
    if (cfg->length < 1 || cfg->length > 46340 || !(cfg->resolution > 0.0f))
        return fail(nullptr, GEM_ERR_INVALID, "gem_create: length must be in [1,46340] and resolution > 0");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev < 1)
        return fail(nullptr, GEM_ERR_NO_DEVICE,
                    std::string("gem_create: no CUDA device (libgem_b200 has no CPU fallback): ") +
                        (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
    int dev = cfg->device;
    if (dev < 0) {
        e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return fail(nullptr, GEM_ERR_NO_DEVICE, cudaGetErrorString(e));
    }
    if (dev >= ndev) return fail(nullptr, GEM_ERR_INVALID, "gem_create: device ordinal out of range");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return fail(nullptr, GEM_ERR_NO_DEVICE, cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, GEM_ERR_NO_DEVICE,
                    "gem_create: this library carries sm_100a code only (found compute capability " +
                        std::to_string(prop.major) + "." + std::to_string(prop.minor) + ")");

    gem_map *m = new gem_map();
    m->cfg = *cfg;
    m->dev = dev;
    SetDev sd(dev);
    m->L = cfg->length;
    const bool tiled = cfg->tile_rows > 0 && cfg->tile_cols > 0;
    if (tiled) {
        if (cfg->tile_row0 < 0 || cfg->tile_col0 < 0 || cfg->tile_row0 + cfg->tile_rows > m->L ||
            cfg->tile_col0 + cfg->tile_cols > m->L) {
            delete m;
            return fail(nullptr, GEM_ERR_INVALID, "gem_create: tile outside the map");
        }
    }
    m->geom.L = m->L;
    m->geom.res = cfg->resolution;
    m->geom.cx = m->geom.cy = 0.0f; // gpu.cu:942
    m->geom.sx = m->geom.sy = 0;    // gpu.cu:943
    m->geom.box_filter = cfg->compat_box_filter ? 1 : 0;
    m->geom.tiled = tiled ? 1 : 0;
    m->geom.r0 = tiled ? cfg->tile_row0 : 0;
    m->geom.rows = tiled ? cfg->tile_rows : m->L;
    m->geom.c0 = tiled ? cfg->tile_col0 : 0;
    m->geom.cols = tiled ? cfg->tile_cols : m->L;
    m->nc = (size_t)m->geom.rows * m->geom.cols;
    m->P = cfg->max_points > 0 ? cfg->max_points : (1 << 21);
    if (m->P > (1 << 22)) m->P = 1 << 22; // the fold's sort key packs the point index into 22 bits
    if ((size_t)m->P < m->nc / 32 + 1) m->P = (int)(m->nc / 32 + 1); // per-point scratch doubles as the ray bitmap

    int rc = GEM_OK;
    auto bail = [&](int code) {
        std::string msg = m->err;
        gem_destroy(m);
        g_create_error = msg;
        return code;
    };
    if (cfg->stream) {
        m->stream = (cudaStream_t)cfg->stream;
    } else {
        e = cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { m->err = cudaGetErrorString(e); return bail(GEM_ERR_CUDA); }
        m->own_stream = true;
    }
    const size_t nc = m->nc, P = (size_t)m->P;
    if ((rc = dev_alloc(m, &m->ml.ev, nc)) || (rc = dev_alloc(m, &m->ml.ci, nc)) ||
        (rc = dev_alloc(m, &m->ml.traver, nc)) || (rc = dev_alloc(m, &m->ml.lowest, nc)) ||
        (rc = dev_alloc(m, &m->ml.rough, nc)) || (rc = dev_alloc(m, &m->ml.slope, nc)) ||
        (rc = dev_alloc(m, &m->ml.traver_out, nc)) || (rc = dev_alloc(m, &m->sc.cnt, nc)) ||
        (rc = dev_alloc(m, &m->sc.cellBase, nc)) || (rc = dev_alloc(m, &m->sc.touched, P < nc ? P : nc)) ||
        (rc = dev_alloc(m, &m->ctr_buf[0], 2)) || (rc = dev_alloc(m, &m->sc.key, P)) ||
        (rc = dev_alloc(m, &m->sc.tsmall, P < nc ? P : nc)) || (rc = dev_alloc(m, &m->sc.tlarge, list_cap(P, nc, FOLD_SMALL_K))) || (rc = dev_alloc(m, &m->sc.tlong, list_cap(P, nc, FOLD_LONG_K))) ||
        (rc = dev_alloc(m, &m->sc.rank, P)) || (rc = dev_alloc(m, &m->sc.h, P)) ||
        (rc = dev_alloc(m, &m->sc.hv, P)) || (rc = dev_alloc(m, &m->sc.recA, P)) ||
        (rc = dev_alloc(m, &m->sc.recI, P)))
        return bail(rc);
    e = cudaHostAlloc((void **)&m->h_ctr, sizeof(Counters), cudaHostAllocDefault);
    if (e != cudaSuccess) { m->err = cudaGetErrorString(e); return bail(GEM_ERR_CUDA); }
    // G_Init_map gpu.cu:198-214
    GEM_LAUNCH(m, GEM_PROF_CLEAR, k_clear_range<<<blocks_for(nc, 256), 256, 0, m->stream>>>(m->ml, 0, nc, 2));
    GEM_LAUNCH(m, GEM_PROF_CLEAR, k_fill<<<blocks_for(nc, 256), 256, 0, m->stream>>>(m->ml.rough, nc, 0.0f));
    GEM_LAUNCH(m, GEM_PROF_CLEAR, k_fill<<<blocks_for(nc, 256), 256, 0, m->stream>>>(m->ml.slope, nc, 0.0f));
    GEM_LAUNCH(m, GEM_PROF_CLEAR, k_fill<<<blocks_for(nc, 256), 256, 0, m->stream>>>(m->ml.traver_out, nc, -10.0f));
    e = cudaMemsetAsync(m->sc.cnt, 0, nc * sizeof(int), m->stream);
    m->ctr_buf[1] = m->ctr_buf[0] + 1;
    m->ctr_last = m->ctr_buf[0];
    if (e == cudaSuccess) e = cudaMemsetAsync(m->ctr_buf[0], 0, 2 * sizeof(Counters), m->stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->stream);
    if (e != cudaSuccess) {
        m->err = std::string("gem_create: init kernels failed (is the device sm_100?): ") + cudaGetErrorString(e);
        return bail(GEM_ERR_NO_DEVICE);
    }
    pend_all_floor(m); // first Fuse floors every cell (gpu.cu:533-534)
    {   // the fused add kernel needs a co-resident grid (cooperative launch)
        int coop = 0, per_sm = 0;
        cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
        if (coop && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_add_fused<IN_XYZI, ATTR_XYZI>, ADD_BLOCK, 0) == cudaSuccess)
            m->coop_blocks = per_sm * prop.multiProcessorCount;
        // measured on B200 (profiles/): four stream-ordered launches (32.8 us/frame) beat the
        // single cooperative launch with three grid barriers (36.5 us/frame), so the fused
        // kernel is opt-in
        const char *env = getenv("GEM_B200_FUSED");
        if (!(env && atoi(env) == 1)) m->coop_blocks = 0;
        // measured on B200: with PDL the frame takes 101 us instead of 32.9 us (waiting dependent
        // CTAs occupy the SMs the predecessor's serial fold tail needs), so it is opt-in
        const char *envp = getenv("GEM_B200_PDL");
        if (envp && atoi(envp) == 1) m->pdl = true;
        const char *envf = getenv("GEM_B200_PDL_FRONT");
        if (envf && atoi(envf) == 1) m->pdl_front = true;
        const char *envn = getenv("GEM_B200_FUSED_MAX_POINTS");
        if (envn && atoi(envn) > 0) m->fused_max_points = atoi(envn);
        cudaGetLastError();
    }
    *out = m;
    return GEM_OK;
}

int gem_destroy(gem_map *m)
{
    if (!m) return GEM_OK;
    SetDev sd(m->dev);
    if (m->stream) cudaStreamSynchronize(m->stream);
    for (auto &sp : m->spans) { cudaEventDestroy(sp.e0); cudaEventDestroy(sp.e1); }
    for (cudaEvent_t e : m->free_events) cudaEventDestroy(e);
    for (void *p : m->allocs) cudaFree(p);
    if (m->h_ctr) cudaFreeHost(m->h_ctr);
    if (m->h_ctr_ring) cudaFreeHost(m->h_ctr_ring);
    if (m->h_frames) cudaFreeHost(m->h_frames);
    for (int i = 0; i < 3; i++) {
        if (m->ev_bin[i]) cudaEventDestroy(m->ev_bin[i]);
        if (m->ev_front[i]) cudaEventDestroy(m->ev_front[i]);
        if (m->ev_fold[i]) cudaEventDestroy(m->ev_fold[i]);
    }
    if (m->front_stream) { cudaStreamSynchronize(m->front_stream); cudaStreamDestroy(m->front_stream); }
    if (m->mid_owned && m->mid_stream) { cudaStreamSynchronize(m->mid_stream); cudaStreamDestroy(m->mid_stream); }
    for (int i = 0; i < 4; i++) if (m->ev_frames[i]) cudaEventDestroy(m->ev_frames[i]);
    for (int i = 0; i < 2; i++) {
        if (m->ev_h2d[i]) cudaEventDestroy(m->ev_h2d[i]);
        if (m->ev_done[i]) cudaEventDestroy(m->ev_done[i]);
    }
    if (m->copy_stream) { cudaStreamSynchronize(m->copy_stream); cudaStreamDestroy(m->copy_stream); }
    if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
    delete m;
    return GEM_OK;
}

void *gem_get_stream(gem_map *m) { return m ? (void *)m->stream : nullptr; }

int gem_debug_phase_stamps(gem_map *m, int enable, unsigned long long out[16])
{
    if (!m) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    if (enable && !m->sc.tstamp) {
        int rc = dev_alloc(m, &m->sc.tstamp, 16);
        if (rc) return rc;
        GEM_CUDA(m, cudaMemsetAsync(m->sc.tstamp, 0, 16 * 8, m->stream));
    }
    if (out && m->sc.tstamp) {
        GEM_CUDA(m, cudaMemcpyAsync(out, m->sc.tstamp, 16 * 8, cudaMemcpyDeviceToHost, m->stream));
        GEM_CUDA(m, cudaStreamSynchronize(m->stream));
        GEM_CUDA(m, cudaMemsetAsync(m->sc.tstamp, 0, 16 * 8, m->stream));
        GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    }
    return GEM_OK;
}

int gem_sync(gem_map *m)
{
    if (!m) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

// ---- Move gpu.cu:1004-1083 -----------------------------------------------------------------
static int index_to_range(int index, int L)
{ // gpu.cu:916-921
    if (index < 0) index += ((-index / L) + 1) * L;
    return index % L;
}
static int d2i_host(double d)
{ // cvt.rzi semantics for the host-side casts of gpu.cu:897,998-999
    if (d != d) return 0;
    if (d >= 2147483648.0) return 2147483647;
    if (d <= -2147483649.0) return -2147483647 - 1;
    return (int)d;
}
static float position_to_range(float p, float shift, float resolution)
{ // gpu.cu:996-1002
    const int p_index = d2i_host((double)roundf(p / resolution));
    const int shift_index = d2i_host((double)roundf(shift / resolution));
    return (float)(p_index + shift_index) * resolution;
}
// scroll clears are deferred: the next add/fuse launch executes them (plus the variance
// floor that G_fuse would apply to the cleared cells), or flush_for_observer does
static void clear_rows(gem_map *m, int start, int n) { m->pending.push_back(RegionOp{1, start, n, 1, 1}); }
static void clear_cols(gem_map *m, int start, int n) { m->pending.push_back(RegionOp{2, start, n, 1, 1}); }

int gem_move(gem_map *m, const float pos[3], float centre_out[2], int start_out[2], float shift_out[2])
{
    if (!m || !pos) return fail(m, GEM_ERR_INVALID, "gem_move: null argument");
    SetDev sd(m->dev);
    const int L = m->L;
    m->sensorZ = pos[2]; // gpu.cu:1011-1012
    float aligned[2] = {0.0f, 0.0f};
    if (m->geom.tiled) {
        // tiled (multi-GPU) maps are global, non-scrolling maps
        if (centre_out) { centre_out[0] = m->geom.cx; centre_out[1] = m->geom.cy; }
        if (start_out) { start_out[0] = 0; start_out[1] = 0; }
        if (shift_out) { shift_out[0] = 0.0f; shift_out[1] = 0.0f; }
        return GEM_OK;
    }
    float centre[2] = {m->geom.cx, m->geom.cy};
    int start[2] = {m->geom.sx, m->geom.sy};
    int indexShift[2];
    for (int i = 0; i < 2; i++) {
        const float ps = pos[i] - centre[i];
        indexShift[i] = d2i_host((double)(ps / m->geom.res) + 0.5 * (ps > 0 ? 1 : -1)); // gpu.cu:897
        aligned[i] = (float)indexShift[i] * m->geom.res;                                   // gpu.cu:909
    }
    for (int i = 0; i < 2; i++) {
        if (indexShift[i] != 0) {
            // |shift| >= L clears everything (the reference tests only the positive side,
            // gpu.cu:1033, and would write out of bounds for shift <= -L)
            if (indexShift[i] >= L || indexShift[i] <= -L) {
                GEM_LAUNCH(m, GEM_PROF_CLEAR, k_clear_range<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, 0, m->nc, 1));
                pend_all_floor(m);
            } else {
                const int sign = indexShift[i] > 0 ? 1 : -1;
                const int startIndex = start[i] - (sign > 0 ? 1 : 0);
                const int endIndex = startIndex + sign - indexShift[i];
                const int nCells = std::abs(indexShift[i]);
                int index = sign < 0 ? startIndex : endIndex;
                index = index_to_range(index, L);
                if (index + nCells <= L) {
                    if (i == 0) clear_rows(m, index, nCells); else clear_cols(m, index, nCells);
                } else {
                    const int firstn = L - index, secondn = nCells - firstn;
                    if (i == 0) { clear_rows(m, index, firstn); clear_rows(m, 0, secondn); }
                    else { clear_cols(m, index, firstn); clear_cols(m, 0, secondn); }
                }
            }
        }
        start[i] = index_to_range(start[i] - indexShift[i], L);
        centre[i] = position_to_range(centre[i], aligned[i], m->geom.res);
    }
    m->geom.cx = centre[0]; m->geom.cy = centre[1];
    m->geom.sx = start[0]; m->geom.sy = start[1];
    if (centre_out) { centre_out[0] = centre[0]; centre_out[1] = centre[1]; }
    if (start_out) { start_out[0] = start[0]; start_out[1] = start[1]; }
    if (shift_out) { shift_out[0] = aligned[0]; shift_out[1] = aligned[1]; }
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

// ---- fused add -------------------------------------------------------------------------------
int gem_add_points(gem_map *m, const void *xyzi, const void *rgba, int n, const gem_frame *frame)
{
    if (!m || !frame || n < 0 || (n > 0 && !xyzi)) return fail(m, GEM_ERR_INVALID, "gem_add_points: bad argument");
    SetDev sd(m->dev);
    int rc = GEM_OK;
    const FrameParams fp = make_frame(frame);
    memset(&m->stats, 0, sizeof m->stats);
    if (n == 0) return flush_all_pending(m);
    for (int off = 0; off < n; off += m->P) {
        const int cn = (n - off < m->P) ? (n - off) : m->P;
        PointInput in{};
        in.xyzi = (const float4 *)xyzi + off;
        in.rgba = rgba ? (const uchar4 *)rgba + off : nullptr;
        AttrInput a{};
        a.xyzi = in.xyzi;
        a.rgba = in.rgba;
        if ((rc = add_chunk<IN_XYZI, ATTR_XYZI>(m, in, a, cn, fp))) return rc;
        if (n > m->P && (rc = read_counters(m, cn, true))) return rc; // chunked: keep totals
    }
    if (n <= m->P) m->stats.points_in = n; // counters are fetched lazily by gem_get_stats
    return GEM_OK;
}

int gem_add_points_host(gem_map *m, const void *xyzi, const void *rgba, int n, const gem_frame *frame)
{
    if (!m || !frame || n < 0 || (n > 0 && !xyzi)) return fail(m, GEM_ERR_INVALID, "gem_add_points_host: bad argument");
    SetDev sd(m->dev);
    int rc = ensure_host_staging(m);
    if (rc) return rc;
    const FrameParams fp = make_frame(frame);
    memset(&m->stats, 0, sizeof m->stats);
    if (n == 0) { if ((rc = flush_all_pending(m))) return rc; return gem_sync(m); }
    for (int off = 0; off < n; off += m->P) {
        const int cn = (n - off < m->P) ? (n - off) : m->P;
        if (cn > 0) {
            GEM_CUDA(m, cudaMemcpyAsync(m->d_xyzi, (const float4 *)xyzi + off, (size_t)cn * 16, cudaMemcpyHostToDevice, m->stream));
            if (rgba)
                GEM_CUDA(m, cudaMemcpyAsync(m->d_rgba, (const uchar4 *)rgba + off, (size_t)cn * 4, cudaMemcpyHostToDevice, m->stream));
        }
        PointInput in{};
        in.xyzi = (const float4 *)m->d_xyzi;
        in.rgba = rgba ? (const uchar4 *)m->d_rgba : nullptr;
        AttrInput a{};
        a.xyzi = in.xyzi;
        a.rgba = in.rgba;
        if ((rc = add_chunk<IN_XYZI, ATTR_XYZI>(m, in, a, cn, fp))) return rc;
        if ((rc = read_counters(m, cn, true))) return rc; // also the host-visible completion point
    }
    return GEM_OK;
}

static int pipe_setup(gem_map *m)
{
    if (m->pipe_ready) return GEM_OK;
    int rc;
    const size_t nc = m->nc, P = (size_t)m->P, T = P < nc ? P : nc;
    for (int i = 0; i < 3; i++) {
        Scratch &sc = m->pipe_sc[i];
        memset(&sc, 0, sizeof sc);
        if ((rc = dev_alloc(m, &sc.cnt, nc)) || (rc = dev_alloc(m, &sc.cellBase, nc)) || (rc = dev_alloc(m, &sc.touched, T)) ||
            (rc = dev_alloc(m, &sc.tsmall, T)) || (rc = dev_alloc(m, &sc.tlarge, list_cap(P, nc, FOLD_SMALL_K))) || (rc = dev_alloc(m, &sc.tlong, list_cap(P, nc, FOLD_LONG_K))) ||
            (rc = dev_alloc(m, &sc.key, P)) || (rc = dev_alloc(m, &sc.rank, P)) || (rc = dev_alloc(m, &sc.h, P)) ||
            (rc = dev_alloc(m, &sc.hv, P)) || (rc = dev_alloc(m, &sc.recA, P)) || (rc = dev_alloc(m, &sc.recI, P)))
            return rc;
        GEM_CUDA(m, cudaMemsetAsync(sc.cnt, 0, nc * sizeof(int), m->stream));
        GEM_CUDA(m, cudaEventCreateWithFlags(&m->ev_bin[i], cudaEventDisableTiming));
        GEM_CUDA(m, cudaEventCreateWithFlags(&m->ev_front[i], cudaEventDisableTiming));
        GEM_CUDA(m, cudaEventCreateWithFlags(&m->ev_fold[i], cudaEventDisableTiming));
    }
    if ((rc = dev_alloc(m, &m->pipe_ctr[0], 4))) return rc;
    for (int i = 1; i < 4; i++) m->pipe_ctr[i] = m->pipe_ctr[0] + i;
    GEM_CUDA(m, cudaMemsetAsync(m->pipe_ctr[0], 0, 4 * sizeof(Counters), m->stream));
    GEM_CUDA(m, cudaStreamCreateWithFlags(&m->front_stream, cudaStreamNonBlocking));
    {   // GEM_B200_STREAM_STAGES=2 keeps alloc+scatter on the transform stream (the two-stage pipeline)
        const char *env = getenv("GEM_B200_STREAM_STAGES");
        if (env && atoi(env) == 2) m->mid_stream = m->front_stream;
        else {
            GEM_CUDA(m, cudaStreamCreateWithFlags(&m->mid_stream, cudaStreamNonBlocking));
            m->mid_owned = true;
        }
    }
    for (int i = 0; i < 3; i++) GEM_CUDA(m, cudaEventRecord(m->ev_fold[i], m->stream)); // sets are free once init is done
    m->pipe_ready = true;
    return GEM_OK;
}

int gem_add_points_stream(gem_map *m, const void *xyzi, const void *rgba, int n, const gem_frame *frame)
{
    if (!m || !frame || n < 0 || (n > 0 && !xyzi)) return fail(m, GEM_ERR_INVALID, "gem_add_points_stream: bad argument");
    if (n > m->P) return fail(m, GEM_ERR_INVALID, "gem_add_points_stream: n exceeds max_points (use gem_add_points)");
    SetDev sd(m->dev);
    int rc = pipe_setup(m);
    if (rc) return rc;
    if (n == 0) return flush_all_pending(m);
    const unsigned i = m->pipe_calls++;
    const int par = (int)(i % 3u), c = (int)(i & 3u);
    Scratch sc = m->pipe_sc[par];
    sc.ctr = m->pipe_ctr[c];
    sc.ctr_next = m->pipe_ctr[(c + 1) & 3];
    sc.tstamp = nullptr;
    const FrameParams fp = make_frame(frame);
    PointInput in{};
    in.xyzi = (const float4 *)xyzi;
    in.rgba = (const uchar4 *)rgba;
    AttrInput a{};
    a.xyzi = in.xyzi;
    a.rgba = in.rgba;
    // stage 1 (front stream): transform+bin of THIS frame -- overlaps alloc+scatter of the previous frame and the
    // fold of the one before; it only waits for the fold that last used this scratch set (three calls ago)
    GEM_CUDA(m, cudaStreamWaitEvent(m->front_stream, m->ev_fold[par], 0));
    RegionOps none{};
    const int pb = blocks_for((size_t)n, ADD_BLOCK, 148 * 16);
    GEM_LAUNCH_ON(m, m->front_stream, GEM_PROF_TRANSFORM_BIN,
                  k_transform_bin<IN_XYZI><<<pb, ADD_BLOCK, 0, m->front_stream>>>(m->geom, m->ml, fp, in, n, sc, none, pb, nullptr, nullptr));
    // stage 2 (mid stream): alloc + scatter
    if (m->mid_stream != m->front_stream) {
        GEM_CUDA(m, cudaEventRecord(m->ev_bin[par], m->front_stream));
        GEM_CUDA(m, cudaStreamWaitEvent(m->mid_stream, m->ev_bin[par], 0));
    }
    GEM_LAUNCH_ON(m, m->mid_stream, GEM_PROF_ALLOC,
                  launch_pdl(m->pdl_front, k_alloc_cells, blocks_for((size_t)n, ADD_BLOCK, 148 * 4), ADD_BLOCK, m->mid_stream, sc));
    GEM_LAUNCH_ON(m, m->mid_stream, GEM_PROF_SCATTER,
                  launch_pdl(m->pdl_front, k_scatter<ATTR_XYZI, 1>, blocks_for((size_t)n, ADD_BLOCK, 148 * 16), ADD_BLOCK, m->mid_stream, a, n, sc));
    GEM_CUDA(m, cudaEventRecord(m->ev_front[par], m->mid_stream));
    // stage 3 (main stream): deferred scroll clears / floors, then the fold (the only kernel that touches the layers)
    if (!m->pending.empty() && (rc = flush_all_pending(m))) return rc;
    GEM_CUDA(m, cudaStreamWaitEvent(m->stream, m->ev_front[par], 0));
    GEM_LAUNCH(m, GEM_PROF_FOLD, k_fold<<<blocks_for((size_t)n, ADD_BLOCK, 148 * 8), ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, sc, 1, 1));
    GEM_CUDA(m, cudaEventRecord(m->ev_fold[par], m->stream));
    GEM_CUDA(m, cudaGetLastError());
    m->ctr_last = m->pipe_ctr[c];
    memset(&m->stats, 0, sizeof m->stats);
    m->stats.points_in = n;
    return GEM_OK;
}

int gem_add_points_multi(gem_map *m, const void *xyzi, const void *rgba, int n_segments, const int *offsets,
                         const gem_frame *frames)
{
    if (!m || !xyzi || !offsets || !frames || n_segments < 1 || n_segments > MAX_SEGMENTS)
        return fail(m, GEM_ERR_INVALID, "gem_add_points_multi: bad argument");
    const int n = offsets[n_segments] - offsets[0];
    if (offsets[0] != 0 || n < 0 || n > m->P) return fail(m, GEM_ERR_INVALID, "gem_add_points_multi: offsets must start at 0 and n <= max_points");
    for (int s = 0; s < n_segments; s++)
        if (offsets[s + 1] < offsets[s]) return fail(m, GEM_ERR_INVALID, "gem_add_points_multi: offsets not monotone");
    SetDev sd(m->dev);
    int rc = GEM_OK;
    if (!m->h_frames) {
        GEM_CUDA(m, cudaHostAlloc((void **)&m->h_frames, 4 * MAX_SEGMENTS * sizeof(FrameParams), cudaHostAllocDefault));
        if ((rc = dev_alloc(m, &m->d_frames, (size_t)4 * MAX_SEGMENTS))) return rc;
        for (int i = 0; i < 4; i++) GEM_CUDA(m, cudaEventCreateWithFlags(&m->ev_frames[i], cudaEventDisableTiming));
    }
    if (n == 0) return flush_all_pending(m);
    const int slot = (int)(m->multi_calls++ & 3u);
    if (m->multi_calls > 4) GEM_CUDA(m, cudaEventSynchronize(m->ev_frames[slot])); // pinned slot free again
    FrameParams *hf = m->h_frames + (size_t)slot * MAX_SEGMENTS, *df = m->d_frames + (size_t)slot * MAX_SEGMENTS;
    SegTable st{};
    st.n = n_segments;
    for (int s = 0; s <= n_segments; s++) st.off[s] = offsets[s];
    for (int s = 0; s < n_segments; s++) hf[s] = make_frame(&frames[s]);
    GEM_CUDA(m, cudaMemcpyAsync(df, hf, (size_t)n_segments * sizeof(FrameParams), cudaMemcpyHostToDevice, m->stream));
    GEM_CUDA(m, cudaEventRecord(m->ev_frames[slot], m->stream));
    RegionOps ro;
    int rb = 0;
    if ((rc = take_region_ops(m, ro, rb))) return rc;
    const Scratch sc = cur_scratch(m);
    PointInput in{};
    in.xyzi = (const float4 *)xyzi;
    in.rgba = (const uchar4 *)rgba;
    AttrInput a{};
    a.xyzi = in.xyzi;
    a.rgba = in.rgba;
    const int U = points_per_thread(n);
    const int pb = blocks_for((size_t)(n + U - 1) / U, ADD_BLOCK, 148 * 32);
    if (U == 4) GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, k_transform_bin_multi<4><<<pb + rb, ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, st, df, in, n, sc, ro, pb));
    else if (U == 2) GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, k_transform_bin_multi<2><<<pb + rb, ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, st, df, in, n, sc, ro, pb));
    else GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, k_transform_bin_multi<1><<<pb + rb, ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, st, df, in, n, sc, ro, pb));
    if ((rc = run_group_fold<ATTR_XYZI>(m, sc, a, n, true, true))) return rc;
    memset(&m->stats, 0, sizeof m->stats);
    m->stats.points_in = n;
    return GEM_OK;
}

int gem_add_points_host_async(gem_map *m, const void *xyzi, const void *rgba, int n, const gem_frame *frame)
{
    if (!m || !frame || n < 0 || (n > 0 && !xyzi)) return fail(m, GEM_ERR_INVALID, "gem_add_points_host_async: bad argument");
    if (n > m->P) return fail(m, GEM_ERR_INVALID, "gem_add_points_host_async: n exceeds max_points (use gem_add_points_host)");
    SetDev sd(m->dev);
    int rc = GEM_OK;
    if (!m->copy_stream) { // lazy set-up: copy stream, two staging sets, events, counter ring
        GEM_CUDA(m, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            if ((rc = dev_alloc(m, (float4 **)&m->d_axyzi[i], (size_t)m->P))) return rc;
            if ((rc = dev_alloc(m, (uchar4 **)&m->d_argba[i], (size_t)m->P))) return rc;
            GEM_CUDA(m, cudaEventCreateWithFlags(&m->ev_h2d[i], cudaEventDisableTiming));
            GEM_CUDA(m, cudaEventCreateWithFlags(&m->ev_done[i], cudaEventDisableTiming));
            GEM_CUDA(m, cudaEventRecord(m->ev_done[i], m->stream));
        }
        GEM_CUDA(m, cudaHostAlloc((void **)&m->h_ctr_ring, 2 * sizeof(Counters), cudaHostAllocDefault));
    }
    if (n == 0) return flush_all_pending(m);
    const int b = (int)(m->async_calls++ & 1u);
    // copy stream: wait until the kernels that last read staging set b are done, then H2D
    GEM_CUDA(m, cudaStreamWaitEvent(m->copy_stream, m->ev_done[b], 0));
    GEM_CUDA(m, cudaMemcpyAsync(m->d_axyzi[b], xyzi, (size_t)n * 16, cudaMemcpyHostToDevice, m->copy_stream));
    if (rgba) GEM_CUDA(m, cudaMemcpyAsync(m->d_argba[b], rgba, (size_t)n * 4, cudaMemcpyHostToDevice, m->copy_stream));
    GEM_CUDA(m, cudaEventRecord(m->ev_h2d[b], m->copy_stream));
    // compute stream: wait for the copy, run the add, read the counters back, mark set b free
    GEM_CUDA(m, cudaStreamWaitEvent(m->stream, m->ev_h2d[b], 0));
    const FrameParams fp = make_frame(frame);
    PointInput in{};
    in.xyzi = (const float4 *)m->d_axyzi[b];
    in.rgba = rgba ? (const uchar4 *)m->d_argba[b] : nullptr;
    AttrInput a{};
    a.xyzi = in.xyzi;
    a.rgba = in.rgba;
    if ((rc = add_chunk<IN_XYZI, ATTR_XYZI>(m, in, a, n, fp))) return rc;
    GEM_CUDA(m, cudaMemcpyAsync(&m->h_ctr_ring[b], m->ctr_last, sizeof(Counters), cudaMemcpyDeviceToHost, m->stream));
    GEM_CUDA(m, cudaEventRecord(m->ev_done[b], m->stream));
    memset(&m->stats, 0, sizeof m->stats);
    m->stats.points_in = n; // the rest is fetched by gem_get_stats
    return GEM_OK;
}

int gem_add_cloud_pcl_host(gem_map *m, const void *pts, int n, const gem_frame *frame)
{
    if (!m || !frame || n < 0 || (n > 0 && !pts)) return fail(m, GEM_ERR_INVALID, "gem_add_cloud_pcl_host: bad argument");
    SetDev sd(m->dev);
    int rc = ensure_pcl_staging(m);
    if (rc) return rc;
    const FrameParams fp = make_frame(frame);
    memset(&m->stats, 0, sizeof m->stats);
    if (n == 0) { if ((rc = flush_all_pending(m))) return rc; return gem_sync(m); }
    for (int off = 0; off < n; off += m->P) {
        const int cn = (n - off < m->P) ? (n - off) : m->P;
        if (cn > 0)
            GEM_CUDA(m, cudaMemcpyAsync(m->d_pcl, (const char *)pts + (size_t)off * 32, (size_t)cn * 32, cudaMemcpyHostToDevice, m->stream));
        PointInput in{};
        in.pcl = (const float4 *)m->d_pcl;
        AttrInput a{};
        a.pcl = in.pcl;
        if ((rc = add_chunk<IN_PCL32, ATTR_PCL32>(m, in, a, cn, fp))) return rc;
        if ((rc = read_counters(m, cn, true))) return rc;
    }
    return GEM_OK;
}

// ---- unfused reference calls ---------------------------------------------------------------
int gem_process_points(gem_map *m, int *map_index, const float *x, const float *y, const float *z, float *var,
                       float *x_ts, float *y_ts, float *z_ts, int n, const gem_frame *frame)
{
    if (!m || !frame || n < 0 || (n > 0 && (!x || !y || !z)))
        return fail(m, GEM_ERR_INVALID, "gem_process_points: bad argument");
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_process_points: not available on tiled handles");
    SetDev sd(m->dev);
    int rc = ensure_compat_staging(m);
    if (rc) return rc;
    const FrameParams fp = make_frame(frame);
    memset(&m->stats, 0, sizeof m->stats);
    for (int off = 0; off < n; off += m->P) {
        const int cn = (n - off < m->P) ? (n - off) : m->P;
        const size_t b = (size_t)cn * 4;
        GEM_CUDA(m, cudaMemcpyAsync(m->d_x, x + off, b, cudaMemcpyHostToDevice, m->stream));
        GEM_CUDA(m, cudaMemcpyAsync(m->d_y, y + off, b, cudaMemcpyHostToDevice, m->stream));
        GEM_CUDA(m, cudaMemcpyAsync(m->d_z, z + off, b, cudaMemcpyHostToDevice, m->stream));
        PointInput in{};
        in.x = m->d_x; in.y = m->d_y; in.z = m->d_z;
        const Scratch sc = cur_scratch(m);
        RegionOps ro{}; // Process_points does not fuse: clears/floors stay pending
        const int pb = blocks_for((size_t)cn, ADD_BLOCK, 148 * 16);
        GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, k_transform_bin<IN_SOA><<<pb, ADD_BLOCK, 0, m->stream>>>(
            m->geom, m->ml, fp, in, cn, sc, ro, pb, m->d_xt, m->d_yt));
        AttrInput a{};
        if ((rc = run_group_fold<ATTR_NONE>(m, sc, a, cn, false, true))) return rc; // lowest-scan only
        if (map_index) GEM_CUDA(m, cudaMemcpyAsync(map_index + off, m->sc.key, b, cudaMemcpyDeviceToHost, m->stream));
        if (var) GEM_CUDA(m, cudaMemcpyAsync(var + off, m->sc.hv, b, cudaMemcpyDeviceToHost, m->stream));
        if (z_ts) GEM_CUDA(m, cudaMemcpyAsync(z_ts + off, m->sc.h, b, cudaMemcpyDeviceToHost, m->stream));
        if (x_ts) GEM_CUDA(m, cudaMemcpyAsync(x_ts + off, m->d_xt, b, cudaMemcpyDeviceToHost, m->stream));
        if (y_ts) GEM_CUDA(m, cudaMemcpyAsync(y_ts + off, m->d_yt, b, cudaMemcpyDeviceToHost, m->stream));
        if ((rc = read_counters(m, cn, true))) return rc;
    }
    return GEM_OK;
}

int gem_fuse(gem_map *m, int n, const int *index, const int *R, const int *G, const int *B, const float *intensity,
             const float *height, const float *var)
{
    if (!m || n < 0 || (n > 0 && (!index || !height || !var))) return fail(m, GEM_ERR_INVALID, "gem_fuse: bad argument");
    SetDev sd(m->dev);
    int rc = ensure_compat_staging(m);
    if (rc) return rc;
    memset(&m->stats, 0, sizeof m->stats);
    if (n == 0) { if ((rc = flush_all_pending(m))) return rc; return gem_sync(m); }
    for (int off = 0; off < n; off += m->P) {
        const int cn = (n - off < m->P) ? (n - off) : m->P;
        const size_t b = (size_t)cn * 4;
        GEM_CUDA(m, cudaMemcpyAsync(m->d_keyin, index + off, b, cudaMemcpyHostToDevice, m->stream));
        GEM_CUDA(m, cudaMemcpyAsync(m->sc.h, height + off, b, cudaMemcpyHostToDevice, m->stream));
        GEM_CUDA(m, cudaMemcpyAsync(m->sc.hv, var + off, b, cudaMemcpyHostToDevice, m->stream));
        AttrInput a{};
        if (R) { GEM_CUDA(m, cudaMemcpyAsync(m->d_R, R + off, b, cudaMemcpyHostToDevice, m->stream)); a.R = m->d_R; }
        if (G) { GEM_CUDA(m, cudaMemcpyAsync(m->d_G, G + off, b, cudaMemcpyHostToDevice, m->stream)); a.G = m->d_G; }
        if (B) { GEM_CUDA(m, cudaMemcpyAsync(m->d_B, B + off, b, cudaMemcpyHostToDevice, m->stream)); a.B = m->d_B; }
        if (intensity) {
            GEM_CUDA(m, cudaMemcpyAsync(m->d_int, intensity + off, b, cudaMemcpyHostToDevice, m->stream));
            a.intensity = m->d_int;
        }
        RegionOps ro;
        int rb = 0;
        if ((rc = take_region_ops(m, ro, rb))) return rc;
        const Scratch sc = cur_scratch(m);
        const int pb = blocks_for((size_t)cn, ADD_BLOCK, 148 * 16);
        GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, k_count_keys<<<pb + rb, ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, m->d_keyin, cn, (int)m->nc, sc, ro, pb));
        if ((rc = run_group_fold<ATTR_INT_ARRAYS>(m, sc, a, cn, true, false))) return rc;
        if ((rc = read_counters(m, cn, true))) return rc;
    }
    return GEM_OK;
}

int gem_var_update(gem_map *m, float dv)
{
    if (!m) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    // x + 0.0f == x for every non-NaN x: the GEM node always passes 0 (ElevationMapping.cpp:944-945)
    if (dv == 0.0f) return GEM_OK;
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_var_update<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, dv));
    if (dv < 0.0f) pend_all_floor(m); // variances may drop below the floor: next Fuse floors every cell
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

int gem_compute_features(gem_map *m)
{
    if (!m) return GEM_ERR_INVALID;
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_compute_features: tiled handles take the halo-padded tile: use gem_compute_features_tiled");
    SetDev sd(m->dev);
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    GEM_LAUNCH(m, GEM_PROF_FEATURES, k_features<false><<<blocks_for(m->nc, 256, 1 << 30), 256, 0, m->stream>>>(m->geom, m->ml, nullptr));
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

static int copy_layer_out(gem_map *m, int layer, void *host, int slot)
{
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    float *dst = m->d_out + (size_t)slot * m->nc;
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_unpack_layer<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, layer, dst));
    GEM_CUDA(m, cudaGetLastError());
    GEM_CUDA(m, cudaMemcpyAsync(host, dst, m->nc * 4, cudaMemcpyDeviceToHost, m->stream));
    return GEM_OK;
}

int gem_map_feature(gem_map *m, float *elevation, float *var, int *R, int *G, int *B, float *rough, float *slope,
                    float *traver, float *intensity)
{
    if (!m) return GEM_ERR_INVALID;
    int rc = gem_compute_features(m);
    if (rc) return rc;
    SetDev sd(m->dev);
    if ((rc = ensure_out_staging(m))) return rc;
    struct { void *p; int layer; } outs[9] = {{elevation, 0}, {var, 1}, {R, 3}, {G, 4}, {B, 5},
                                              {rough, 8}, {slope, 9}, {traver, 10}, {intensity, 2}};
    for (int k = 0; k < 9; k++)
        if (outs[k].p && (rc = copy_layer_out(m, outs[k].layer, outs[k].p, k))) return rc;
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

int gem_get_layer_device(gem_map *m, int layer, void *out_device)
{
    if (!m || !out_device || layer < 0 || layer > 10) return fail(m, GEM_ERR_INVALID, "gem_get_layer_device: bad argument");
    SetDev sd(m->dev);
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_unpack_layer<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, layer, out_device));
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

int gem_compute_features_tiled(gem_map *m, const float *padded_elevation)
{
    if (!m || !padded_elevation) return fail(m, GEM_ERR_INVALID, "gem_compute_features_tiled: bad argument");
    if (!m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_compute_features_tiled: handle is not tiled");
    SetDev sd(m->dev);
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    GEM_LAUNCH(m, GEM_PROF_FEATURES, k_features<true><<<blocks_for(m->nc, 256, 1 << 30), 256, 0, m->stream>>>(m->geom, m->ml, padded_elevation));
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

int gem_raytracing_tiled(gem_map *m, const float *global_lowest)
{
    if (!m || !global_lowest) return fail(m, GEM_ERR_INVALID, "gem_raytracing_tiled: bad argument");
    if (!m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_raytracing_tiled: handle is not tiled");
    SetDev sd(m->dev);
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    const size_t ng = (size_t)m->L * m->L;
    if (!m->d_gbitmap) { int rc = dev_alloc(m, &m->d_gbitmap, ng / 32 + 1); if (rc) return rc; }
    int *ray_count = &m->ctr_buf[m->ctr_cur ^ 1]->pad4[0];
    GEM_CUDA(m, cudaMemsetAsync(ray_count, 0, sizeof(int), m->stream));
    MapLayers mlg = m->ml;
    mlg.lowest = const_cast<float *>(global_lowest); // rays probe the replicated, map-wide lowest layer
    GEM_LAUNCH(m, GEM_PROF_RAYTRACE, k_ray_collect<<<blocks_for(m->nc, 256, 1 << 30), 256, 0, m->stream>>>(m->geom, m->ml, m->cfg.obstacle_threshold, m->sc.cellBase, ray_count));
    GEM_LAUNCH(m, GEM_PROF_RAYTRACE, k_lowest_bitmap<<<blocks_for(ng, 256, 1 << 30), 256, 0, m->stream>>>(global_lowest, (int)ng, m->d_gbitmap));
    GEM_LAUNCH(m, GEM_PROF_RAYTRACE, k_ray_trace<<<148 * 8, 256, 0, m->stream>>>(m->geom, mlg, m->d_gbitmap, m->sensorZ, m->sc.cellBase, ray_count));
    GEM_LAUNCH(m, GEM_PROF_CLEAR, k_fill<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml.lowest, m->nc, 10.0f)); // own tile's lowest
    GEM_CUDA(m, cudaGetLastError());
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

int gem_raytracing(gem_map *m)
{
    if (!m) return GEM_ERR_INVALID;
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_raytracing: tiled handles take the map-wide lowest layer: use gem_raytracing_tiled");
    SetDev sd(m->dev);
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    // ray list lives in cellBase (free between add calls), its length in the spare counter buffer
    int *ray_count = &m->ctr_buf[m->ctr_cur ^ 1]->pad4[0];
    GEM_CUDA(m, cudaMemsetAsync(ray_count, 0, sizeof(int), m->stream));
    GEM_LAUNCH(m, GEM_PROF_RAYTRACE, k_ray_collect<<<blocks_for(m->nc, 256, 1 << 30), 256, 0, m->stream>>>(m->geom, m->ml, m->cfg.obstacle_threshold, m->sc.cellBase, ray_count));
    uint32_t *bitmap = (uint32_t *)m->sc.rank; // nc/32 words <= max_points (checked at create)
    GEM_LAUNCH(m, GEM_PROF_RAYTRACE, k_lowest_bitmap<<<blocks_for(m->nc, 256, 1 << 30), 256, 0, m->stream>>>(m->ml.lowest, (int)m->nc, bitmap));
    GEM_LAUNCH(m, GEM_PROF_RAYTRACE, k_ray_trace<<<148 * 8, 256, 0, m->stream>>>(m->geom, m->ml, bitmap, m->sensorZ, m->sc.cellBase, ray_count));
    GEM_LAUNCH(m, GEM_PROF_CLEAR, k_fill<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml.lowest, m->nc, 10.0f)); // G_Clear_maplowest
    GEM_CUDA(m, cudaGetLastError());
    GEM_CUDA(m, cudaStreamSynchronize(m->stream)); // gpu.cu:1312
    return GEM_OK;
}

int gem_opt_move(gem_map *m, const float opt_p[2], float height_update, float aligned_out[2])
{
    if (!m || !opt_p) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    float c[2] = {m->geom.cx, m->geom.cy};
    for (int i = 0; i < 2; i++) { // alignedPosition gpu.cu:1203-1213
        const float ps = opt_p[i] - c[i];
        const int is = d2i_host((double)(ps / m->geom.res) + 0.5 * (ps > 0 ? 1 : -1));
        c[i] = c[i] + m->geom.res * (float)is;
        if (aligned_out) aligned_out[i] = c[i];
    }
    m->geom.cx = c[0]; m->geom.cy = c[1];
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_add_height<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, height_update));
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

int gem_closeloop(gem_map *m, const float up[2], float height_update)
{
    if (!m || !up) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    float c[2] = {m->geom.cx, m->geom.cy};
    for (int i = 0; i < 2; i++) { // gpu.cu:1242-1247
        const float ps = up[i] - c[i];
        const int is = d2i_host((double)(ps / m->geom.res) + 0.5 * (ps > 0 ? 1 : -1));
        const float aligned = (float)is * m->geom.res;
        c[i] = position_to_range(c[i], aligned, m->geom.res);
    }
    m->geom.cx = c[0]; m->geom.cy = c[1];
    { int rcf = flush_for_observer(m); if (rcf) return rcf; }
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_add_height<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, height_update));
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

int gem_colourise_points(gem_map *m, void *xyzi, int n, const double Tc[12], const double Tl[16], const unsigned char *bgr,
                         int width, int height, int row_stride, void *rgba_out)
{
    if (!m || n < 0 || (n > 0 && (!xyzi || !rgba_out)) || !Tc || !Tl || !bgr || width < 1 || height < 1 || row_stride < 3 * width)
        return fail(m, GEM_ERR_INVALID, "gem_colourise_points: bad argument");
    SetDev sd(m->dev);
    ProjParams pp;
    for (int i = 0; i < 3; i++) // P_lidar2img = Tcamera * TLidar (ElevationMapping.cpp:347), double, left-to-right sums
        for (int j = 0; j < 4; j++) {
            double a = Tc[4 * i + 0] * Tl[0 + j];
            for (int k = 1; k < 4; k++) a = a + Tc[4 * i + k] * Tl[4 * k + j];
            pp.P[4 * i + j] = a;
        }
    pp.width = width; pp.height = height; pp.row_stride = row_stride;
    if (n > 0)
        GEM_LAUNCH(m, GEM_PROF_OTHER, k_colourise<<<blocks_for((size_t)n, 256), 256, 0, m->stream>>>((float4 *)xyzi, n, pp, bgr, (uchar4 *)rgba_out));
    GEM_CUDA(m, cudaGetLastError());
    return GEM_OK;
}

int gem_export_layers(gem_map *m, float *host_layers[9])
{
    if (!m || !host_layers) return GEM_ERR_INVALID;
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_export_layers: not available on tiled handles");
    SetDev sd(m->dev);
    int rc = ensure_out_staging(m);
    if (rc) return rc;
    if ((rc = flush_for_observer(m))) return rc;
    dim3 grid((m->L + 31) / 32, (m->L + 31) / 32);
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_export_colmajor<<<grid, 256, 0, m->stream>>>(m->ml, m->L, m->d_out));
    GEM_CUDA(m, cudaGetLastError());
    for (int k = 0; k < 9; k++)
        if (host_layers[k])
            GEM_CUDA(m, cudaMemcpyAsync(host_layers[k], m->d_out + (size_t)k * m->nc, m->nc * 4, cudaMemcpyDeviceToHost, m->stream));
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

int gem_export_orthomosaic(gem_map *m, unsigned char *host_bgr)
{
    if (!m || !host_bgr) return fail(m, GEM_ERR_INVALID, "gem_export_orthomosaic: null argument");
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_export_orthomosaic: not available on tiled handles");
    SetDev sd(m->dev);
    int rc = ensure_out_staging(m);
    if (rc) return rc;
    if ((rc = flush_for_observer(m))) return rc;
    unsigned char *d_img = reinterpret_cast<unsigned char *>(m->d_out);
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_orthomosaic<<<blocks_for((m->nc + 3) / 4, 256, 1 << 30), 256, 0, m->stream>>>(m->geom, m->ml, d_img));
    GEM_CUDA(m, cudaGetLastError());
    GEM_CUDA(m, cudaMemcpyAsync(host_bgr, d_img, m->nc * 3, cudaMemcpyDeviceToHost, m->stream));
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

int gem_export_visual_points(gem_map *m, float *host_xyz, unsigned char *host_rgb, int capacity, int *count_out)
{
    if (!m || !count_out || capacity < 0 || (capacity > 0 && (!host_xyz || !host_rgb)))
        return fail(m, GEM_ERR_INVALID, "gem_export_visual_points: bad argument");
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_export_visual_points: not available on tiled handles");
    SetDev sd(m->dev);
    int rc = ensure_out_staging(m);
    if (rc) return rc;
    if ((rc = flush_for_observer(m))) return rc;
    // staging (9 floats per cell): xyz = 3 floats per cell, rgb = 3 bytes per cell behind it
    VisualSrc src;
    src.ml = m->ml;
    src.f = grid_frame(m, m->geom.cx, m->geom.cy, m->geom.sx, m->geom.sy);
    src.xyz = m->d_out;
    src.rgb = reinterpret_cast<unsigned char *>(m->d_out + 3 * m->nc);
    const int cap = (int)std::min<size_t>((size_t)capacity, m->nc);
    int total = 0;
    if ((rc = compact_cells(m, src, cap, &total))) return rc;
    *count_out = total; // the number of shown cells; only min(total, capacity) points are written
    const size_t n = (size_t)std::min(total, cap);
    if (n) {
        GEM_CUDA(m, cudaMemcpyAsync(host_xyz, src.xyz, n * 12, cudaMemcpyDeviceToHost, m->stream));
        GEM_CUDA(m, cudaMemcpyAsync(host_rgb, src.rgb, n * 3, cudaMemcpyDeviceToHost, m->stream));
        GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    }
    return GEM_OK;
}

int gem_snapshot_shown(gem_map *m)
{
    if (!m) return GEM_ERR_INVALID;
    if (m->geom.tiled) return fail(m, GEM_ERR_INVALID, "gem_snapshot_shown: not available on tiled handles");
    SetDev sd(m->dev);
    int rc;
    if ((rc = flush_for_observer(m))) return rc;
    if (!m->prev_ev) {
        if ((rc = dev_alloc(m, &m->prev_ev, m->nc)) || (rc = dev_alloc(m, &m->prev_ci, m->nc)) || (rc = dev_alloc(m, &m->prev_tr, m->nc))) return rc;
    }
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_snapshot_shown<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, m->prev_ev, m->prev_ci, m->prev_tr));
    GEM_CUDA(m, cudaGetLastError());
    m->prev_geom = m->geom;
    m->prev_valid = true;
    return GEM_OK;
}

int gem_harvest_scrolled_out(gem_map *m, const float current_xy[2], const float shift_xy[2], void *host_points32,
                             int capacity, int *count_out)
{
    if (!m || !current_xy || !shift_xy || !count_out || capacity < 0 || (capacity > 0 && !host_points32))
        return fail(m, GEM_ERR_INVALID, "gem_harvest_scrolled_out: bad argument");
    if (!m->prev_valid) return fail(m, GEM_ERR_INVALID, "gem_harvest_scrolled_out: no snapshot (call gem_snapshot_shown first)");
    SetDev sd(m->dev);
    int rc = ensure_out_staging(m);
    if (rc) return rc;
    HarvestSrc src;
    src.pev = m->prev_ev; src.pci = m->prev_ci; src.ptr = m->prev_tr;
    src.f = grid_frame(m, m->prev_geom.cx, m->prev_geom.cy, m->prev_geom.sx, m->prev_geom.sy);
    // :727-734: current_x (float) -+ length_ * resolution_ / 2 with the node's double resolution_
    const double halfwin = (double)m->L * src.f.res / 2;
    src.lox = (double)current_xy[0] - halfwin; src.hix = (double)current_xy[0] + halfwin;
    src.loy = (double)current_xy[1] - halfwin; src.hiy = (double)current_xy[1] + halfwin;
    src.dx = shift_xy[0]; src.dy = shift_xy[1];
    src.out = reinterpret_cast<float4 *>(m->d_out); // 8 of the 9 staging floats per cell
    const int cap = (int)std::min<size_t>((size_t)capacity, m->nc);
    int total = 0;
    if ((rc = compact_cells(m, src, cap, &total))) return rc;
    *count_out = total;
    const size_t n = (size_t)std::min(total, cap);
    if (n) {
        GEM_CUDA(m, cudaMemcpyAsync(host_points32, src.out, n * 32, cudaMemcpyDeviceToHost, m->stream));
        GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    }
    return GEM_OK;
}

int gem_get_layer(gem_map *m, int layer, void *host_out)
{
    if (!m || !host_out || layer < 0 || layer > 9) return fail(m, GEM_ERR_INVALID, "gem_get_layer: bad argument");
    SetDev sd(m->dev);
    int rc = ensure_out_staging(m);
    if (rc) return rc;
    if ((rc = copy_layer_out(m, layer, host_out, 0))) return rc;
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    return GEM_OK;
}

int gem_set_layer(gem_map *m, int layer, const void *host_in)
{
    if (!m || !host_in || layer < 0 || layer > 9) return fail(m, GEM_ERR_INVALID, "gem_set_layer: bad argument");
    SetDev sd(m->dev);
    int rc = ensure_out_staging(m);
    if (rc) return rc;
    if ((rc = flush_for_observer(m))) return rc;
    GEM_CUDA(m, cudaMemcpyAsync(m->d_out, host_in, m->nc * 4, cudaMemcpyHostToDevice, m->stream));
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_pack_layer<<<blocks_for(m->nc, 256), 256, 0, m->stream>>>(m->ml, m->nc, layer, m->d_out));
    GEM_CUDA(m, cudaGetLastError());
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    if (layer == GEM_LAYER_VARIANCE) pend_all_floor(m);
    return GEM_OK;
}

int gem_get_state(gem_map *m, float centre[2], int start[2], float *sensor_z)
{
    if (!m) return GEM_ERR_INVALID;
    if (centre) { centre[0] = m->geom.cx; centre[1] = m->geom.cy; }
    if (start) { start[0] = m->geom.sx; start[1] = m->geom.sy; }
    if (sensor_z) *sensor_z = m->sensorZ;
    return GEM_OK;
}

int gem_get_stats(gem_map *m, gem_stats *out)
{
    if (!m || !out) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    if (m->stats.points_in > 0 && m->stats.cells_touched == 0 && m->stats.points_binned == 0) {
        // device-pointer call: counters not fetched yet
        const long long n_in = m->stats.points_in;
        int rc = read_counters(m, 0, false);
        if (rc) return rc;
        m->stats.points_in = n_in;
    }
    *out = m->stats;
    return GEM_OK;
}

int gem_profile_enable(gem_map *m, int on)
{
    if (!m) return GEM_ERR_INVALID;
    m->profiling = on != 0;
    return GEM_OK;
}

int gem_profile_read(gem_map *m, gem_profile *out, int reset)
{
    if (!m || !out) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    for (auto &sp : m->spans) {
        float ms = 0.0f;
        if (cudaEventElapsedTime(&ms, sp.e0, sp.e1) == cudaSuccess) {
            m->prof_ms[sp.cls] += ms;
            m->prof_count[sp.cls]++;
        }
        m->free_events.push_back(sp.e0);
        m->free_events.push_back(sp.e1);
    }
    m->spans.clear();
    out->launches = m->launches;
    for (int i = 0; i < GEM_PROF_CLASSES; i++) {
        out->ms[i] = m->prof_ms[i];
        out->count[i] = m->prof_count[i];
    }
    if (reset) {
        m->launches = 0;
        for (int i = 0; i < GEM_PROF_CLASSES; i++) { m->prof_ms[i] = 0.0; m->prof_count[i] = 0; }
    }
    return GEM_OK;
}

int gem_selftest_division(gem_map *m, unsigned long long seed, unsigned long long n, unsigned long long *mismatches_out,
                          unsigned long long *fast_out)
{
    if (!m || !mismatches_out) return GEM_ERR_INVALID;
    SetDev sd(m->dev);
    unsigned long long *d = nullptr;
    GEM_CUDA(m, cudaMalloc((void **)&d, 16));
    GEM_CUDA(m, cudaMemsetAsync(d, 0, 16, m->stream));
    GEM_LAUNCH(m, GEM_PROF_OTHER, k_div_selftest<<<148 * 8, 256, 0, m->stream>>>(seed, (size_t)n, d, d + 1));
    unsigned long long h[2] = {0, 0};
    GEM_CUDA(m, cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, m->stream));
    GEM_CUDA(m, cudaStreamSynchronize(m->stream));
    cudaFree(d);
    *mismatches_out = h[0];
    if (fast_out) *fast_out = h[1];
    return GEM_OK;
}

int gem_host_alloc(void **out, unsigned long long bytes)
{
    if (!out) return GEM_ERR_INVALID;
    return cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? GEM_OK : GEM_ERR_NOMEM;
}
int gem_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? GEM_OK : GEM_ERR_CUDA; }

// ---- multi-GPU routing -----------------------------------------------------------------------
int gem_route_points(gem_map *m, const void *xyzi, const void *rgba, int n, const gem_frame *frame, int tiles_r,
                     int tiles_c, void *rec_out, int *counts_out, int bucket_stride)
{
    if (!m || !frame || n < 0 || tiles_r < 1 || tiles_c < 1 || !rec_out || !counts_out || (n > 0 && !xyzi))
        return fail(m, GEM_ERR_INVALID, "gem_route_points: bad argument");
    if (n > m->P) return fail(m, GEM_ERR_INVALID, "gem_route_points: n exceeds max_points");
    if (tiles_r * tiles_c > ROUTE_MAX_OWNERS) return fail(m, GEM_ERR_INVALID, "gem_route_points: too many tiles");
    SetDev sd(m->dev);
    const FrameParams fp = make_frame(frame);
    MapGeom gg = m->geom;
    gg.tiled = 0; // routing works on global geographic indices
    if (bucket_stride > 0) // padded layout: unused slots must read as "no record" (gkey = -1)
        GEM_CUDA(m, cudaMemsetAsync(rec_out, 0xff, (size_t)tiles_r * tiles_c * bucket_stride * sizeof(RouteRec), m->stream));
    const cudaError_t e = route_points(m->stream, gg, fp, (const float4 *)xyzi, (const uchar4 *)rgba, n, tiles_r,
                                       tiles_c, cur_scratch(m), m->nc, (RouteRec *)rec_out, counts_out, bucket_stride);
    m->launches += 3;
    if (e != cudaSuccess) return fail(m, GEM_ERR_CUDA, std::string("gem_route_points: ") + cudaGetErrorString(e));
    return GEM_OK;
}

static int fuse_records_impl(gem_map *m, const void *rec, int n, const int *src_counts, int stride);

int gem_fuse_records(gem_map *m, const void *rec, int n) { return fuse_records_impl(m, rec, n, nullptr, 1); }

int gem_fuse_records_counted(gem_map *m, const void *rec, const int *src_counts, int n_sources, int bucket_stride)
{
    if (!m || !rec || !src_counts || n_sources < 1 || bucket_stride < 1) return fail(m, GEM_ERR_INVALID, "gem_fuse_records_counted: bad argument");
    if ((long long)n_sources * bucket_stride > m->P) return fail(m, GEM_ERR_INVALID, "gem_fuse_records_counted: n_sources*bucket_stride exceeds max_points");
    return fuse_records_impl(m, rec, n_sources * bucket_stride, src_counts, bucket_stride);
}

int gem_route_points_peer(gem_map *m, const void *xyzi, const void *rgba, int n, const gem_frame *frame, int tiles_r, int tiles_c,
                          const unsigned long long *peer_recv, const unsigned long long *peer_counts, int my_rank, int bucket_stride)
{
    const int no = tiles_r * tiles_c;
    if (!m || !frame || n < 0 || tiles_r < 1 || tiles_c < 1 || !peer_recv || !peer_counts || (n > 0 && !xyzi) || bucket_stride < n ||
        my_rank < 0 || my_rank >= no)
        return fail(m, GEM_ERR_INVALID, "gem_route_points_peer: bad argument (bucket_stride must be >= n)");
    if (n > m->P) return fail(m, GEM_ERR_INVALID, "gem_route_points_peer: n exceeds max_points");
    if (no > ROUTE_MAX_OWNERS) return fail(m, GEM_ERR_INVALID, "gem_route_points_peer: too many tiles");
    SetDev sd(m->dev);
    int rc = GEM_OK;
    if (!m->d_owner_cnt && (rc = dev_alloc(m, &m->d_owner_cnt, ROUTE_MAX_OWNERS))) return rc;
    const FrameParams fp = make_frame(frame);
    MapGeom gg = m->geom;
    gg.tiled = 0;
    PeerTable pt;
    memset(&pt, 0, sizeof pt);
    for (int o = 0; o < no; o++) { pt.recv[o] = peer_recv[o]; pt.counts[o] = peer_counts[o]; }
    const cudaError_t e = route_points(m->stream, gg, fp, (const float4 *)xyzi, (const uchar4 *)rgba, n, tiles_r, tiles_c,
                                       cur_scratch(m), m->nc, nullptr, m->d_owner_cnt, bucket_stride, &pt, my_rank);
    m->launches += 3;
    if (e != cudaSuccess) return fail(m, GEM_ERR_CUDA, std::string("gem_route_points_peer: ") + cudaGetErrorString(e));
    return GEM_OK;
}

static int fuse_records_impl(gem_map *m, const void *rec, int n, const int *src_counts, int stride)
{
    if (!m || n < 0 || (n > 0 && !rec)) return fail(m, GEM_ERR_INVALID, "gem_fuse_records: bad argument");
    SetDev sd(m->dev);
    int rc = flush_all_pending(m);
    if (rc) return rc;
    memset(&m->stats, 0, sizeof m->stats);
    for (int off = 0; off < n; off += m->P) {
        const int cn = (n - off < m->P) ? (n - off) : m->P;
        const Scratch sc = cur_scratch(m);
        GEM_LAUNCH(m, GEM_PROF_TRANSFORM_BIN, k_count_records<<<blocks_for((size_t)cn, ADD_BLOCK, 1 << 30), ADD_BLOCK, 0, m->stream>>>(m->geom, (const RouteRec *)rec + off, cn, sc, src_counts, stride));
        GEM_LAUNCH(m, GEM_PROF_ALLOC, k_alloc_cells<<<blocks_for((size_t)cn, ADD_BLOCK, 148 * 4), ADD_BLOCK, 0, m->stream>>>(sc));
        GEM_LAUNCH(m, GEM_PROF_SCATTER, k_scatter_records<<<blocks_for((size_t)cn, ADD_BLOCK, 1 << 30), ADD_BLOCK, 0, m->stream>>>((const RouteRec *)rec + off, cn, sc));
        GEM_LAUNCH(m, GEM_PROF_FOLD, k_fold<<<blocks_for((size_t)cn, ADD_BLOCK, 148 * 8), ADD_BLOCK, 0, m->stream>>>(m->geom, m->ml, sc, 1, 1));
        GEM_CUDA(m, cudaGetLastError());
        call_done(m);
        if (n > m->P && (rc = read_counters(m, cn, true))) return rc;
    }
    if (n <= m->P) m->stats.points_in = n;
    return GEM_OK;
}

} // extern "C"
