/*
 * gem_b200.h -- C ABI of libgem_b200.so: the B200-native (sm_100a) replacement for GEM's
 * GPU map library `libgpu.so` (reference: elevation_mapping/elevation_mapping/cuda/
 * gpu_process.cu, "gpu.cu" below; ZJU-Robotics-Lab/GEM @ d7ec953).
 *
 * The reference boundary is 9 C++-mangled free functions declared ad hoc by their callers
 * (ElevationMapping.cpp:44-50, SensorProcessorBase.cpp:34, RobotMotionMapUpdater.cpp:18)
 * with Eigen types by value and one process-global map.  This header is the C-ABI they bind
 * to instead: plain pointers and sizes, an opaque per-map handle, int status codes.
 * compat/gpu_process_shim.cpp re-exports the 9 original symbols on top of it (needs Eigen,
 * compiled inside the catkin workspace), see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns GEM_OK (0) or a GEM_ERR_* code; gem_last_error() gives text;
 *   - "device" pointers are CUDA device pointers on the handle's device, "host" pointers
 *     are ordinary (pageable or pinned) host memory;
 *   - all work of one handle is ordered on one CUDA stream; functions taking device
 *     pointers are asynchronous on that stream, functions taking host pointers return
 *     after the result is visible to the host (like the reference wrappers, which are all
 *     host-synchronous);
 *   - a handle is thread-safe: every entry point takes the handle's (recursive) mutex, so the
 *     reference node's three threads (processpoints: Process_points OUTSIDE MapMutex_ and Fuse
 *     inside, ElevationMapping.cpp:271-282; processmapcells :286-300; the spinner's Move /
 *     Map_feature / Raytracing :388-421) may enter concurrently; calls are serialised, and
 *     what one call enqueued is ordered before the next on the handle's stream;
 *   - there is NO CPU fallback: gem_create fails with GEM_ERR_NO_DEVICE without a GPU.
 *
 * Layer layout seen through this ABI is the reference's: row-major L*L arrays, index
 * x*L+y; elevation/variance/intensity/colour/traver are indexed by STORAGE index (circular
 * buffer), lowest by GEOGRAPHIC index (gpu.cu:430-431).  Empty cell
 * sentinels: elevation == -10, variance == -10, traver == -10 (gpu.cu:203-210).
 */
#ifndef GEM_B200_H
#define GEM_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define GEM_B200_VERSION 100

enum {
    GEM_OK = 0,
    GEM_ERR_INVALID = 1,   /* bad argument                                   */
    GEM_ERR_CUDA = 2,      /* CUDA runtime error (text in gem_last_error)    */
    GEM_ERR_NO_DEVICE = 3, /* no usable CUDA device / kernels not loadable   */
    GEM_ERR_NOMEM = 4
};

typedef struct gem_map gem_map; /* opaque */

/* Replaces the arguments of Init_GPU_elevationmap (gpu.cu:940) + the hard-coded knobs. */
typedef struct gem_config {
    int length;                  /* cells per side L (gpu.cu:35)                          */
    float resolution;            /* metres per cell (gpu.cu:36)                           */
    float mahalanobis_threshold; /* uploaded but unused by the reference (gate is 5)      */
    float obstacle_threshold;    /* ElevationMapping.cpp:194 hard-codes 0.7               */
    int compat_box_filter;       /* 1: apply the sensor-frame box filter of gpu.cu:393    */
    int max_points;              /* per-launch capacity; larger calls are chunked. 0=auto */
    int device;                  /* CUDA ordinal, -1 = current device                     */
    void *stream;                /* cudaStream_t to run on; NULL = library-owned stream   */
    /* spatial tile owned by this handle (multi-GPU tiling).  All zero = whole
     * map.  Tiled handles do not scroll (gem_move keeps start index 0). */
    int tile_row0, tile_rows, tile_col0, tile_cols;
    /* The node keeps its resolution as a double (ElevationMapping.hpp:314) and grid_map computes cell-centre
     * positions with it, while the CUDA side gets the float.  Used only for the positions emitted by
     * gem_export_visual_points / gem_harvest_scrolled_out; 0 = (double)resolution. */
    double grid_resolution;
} gem_config;

enum { GEM_SENSOR_LASER = 0, GEM_SENSOR_STRUCTURED_LIGHT = 1 };

/* Sensor noise model: laser = gpu.cu:410-411 (C_min_r, C_beam_a, C_beam_c);
 * structured light = StructuredLightSensorProcessor.cpp:129-139 (doubles), plus the depth
 * pass-through of its cleanPointCloud (:51-66). */
typedef struct gem_sensor_model {
    int type;
    float min_radius, beam_angle, beam_constant;
    double normal_factor_a, normal_factor_b, normal_factor_c, normal_factor_d, normal_factor_e;
    double lateral_factor;
    /* structured light only: cleanPointCloud's pcl::PassThrough on the sensor-frame z
     * (StructuredLightSensorProcessor.cpp:51-66, realsense_d435.yaml 0.2 / 3.25; the node's defaults are
     * DBL_MIN / DBL_MAX, :39-40).  PCL converts the limits to float and drops a point when it is not finite or
     * z < min || z > max.  Always applied for GEM_SENSOR_STRUCTURED_LIGHT by the fused add calls (a dropped point
     * is a rejected point: the order of the remaining ones is unchanged); ignored for the laser model, whose
     * cleanPointCloud only removes non-finite points (LaserSensorProcessor.cpp:50-59) -- those never pass the
     * height window of gpu.cu:397 anyway. */
    double cutoff_min_depth, cutoff_max_depth;
} gem_sensor_model;

/* Per-frame constants = the by-value arguments of Process_points (gpu.cu:1085), derived by
 * SensorProcessorBase::GPUPointCloudprocess / readcomputerparam (SPB.cpp:171-206,270-290).
 * Matrices are row-major. */
typedef struct gem_frame {
    float T[16];                 /* map <- sensor (Eigen::Matrix4f transform)             */
    float sensor_jacobian[3];    /* row 3 of R_map<-sensor (SPB.cpp:275)                  */
    float rotation_variance[9];  /* Sigma_q, all zero in GEM (SPB.cpp:202-204)            */
    float C_SB_transpose[9];     /* SPB.cpp:283                                           */
    float P_mul_C_BM_transpose[3]; /* SPB.cpp:282                                         */
    float B_r_BS_skew[9];        /* SPB.cpp:284                                           */
    double rel_lower, rel_upper; /* height window, double compare (gpu.cu:397)            */
    gem_sensor_model sensor;
} gem_frame;

typedef struct gem_stats {
    long long points_in;      /* points offered by the last add/process call              */
    long long points_binned;  /* accepted by the filters AND inside the grid              */
    long long cells_touched;  /* distinct cells updated by the last add/fuse call         */
    int max_points_per_cell;  /* longest per-cell list of the last call (exact above 8)       */
} gem_stats;

/* layer ids for gem_get_layer / gem_set_layer */
enum {
    GEM_LAYER_ELEVATION = 0, GEM_LAYER_VARIANCE = 1, GEM_LAYER_INTENSITY = 2,
    GEM_LAYER_COLOR_R = 3, GEM_LAYER_COLOR_G = 4, GEM_LAYER_COLOR_B = 5,
    GEM_LAYER_TRAVER = 6, GEM_LAYER_LOWEST = 7, GEM_LAYER_ROUGH = 8, GEM_LAYER_SLOPE = 9
};

int gem_version(void);
const char *gem_last_error(const gem_map *m); /* m may be NULL: last create error */

/* Init_GPU_elevationmap (gpu.cu:940-994): allocate layers + scratch, init sentinels. */
int gem_create(const gem_config *cfg, gem_map **out);
int gem_destroy(gem_map *m);
int gem_sync(gem_map *m); /* wait for the handle's stream */
/* the cudaStream_t all work of this handle is ordered on (record your own events there) */
void *gem_get_stream(gem_map *m);
/* enqueue everything the pipelined add calls have deferred (the fold of the last gem_add_points_stream /
 * _multi / _host_async call) on the handle's stream without waiting for it; gem_sync and every call that
 * reads or changes the map do this implicitly */
int gem_flush(gem_map *m);
/* debug: %globaltimer marks (ns) of the add kernels' phases, accumulated since the last call (min of the starts,
 * max of the marks): [0] k_bin start, [1] ranks drawn, [2] slots reserved, [3] pointers published, [4] records
 * stored; [8] k_fold start, [9] long + large lists done, [10] short lists done, [11] longest lists loaded,
 * [12] longest lists folded.  enable != 0 (re)arms the marks, 0 disarms them; out may be NULL. */
int gem_debug_stamps(gem_map *m, int enable, unsigned long long out[16]);

/* Move (gpu.cu:1004-1083): scroll the circular buffer to follow pos[0..1], record
 * pos[2] as sensorZatLowestScan.  Outputs may be NULL. */
int gem_move(gem_map *m, const float pos[3], float centre_out[2], int start_out[2],
             float aligned_shift_out[2]);

/* ---- fused hot path: Process_points + Fuse with device-resident intermediates --------
 * xyzi: n x float4 {x, y, z, intensity} in the sensor frame; rgba: n x uchar4 {r,g,b,-}
 * or NULL (colour path off).  NOTE (reference behaviour, gpu.cu:488): a cell takes a point's
 * intensity AND colour only when R, G, B and intensity are ALL non-zero, so with rgba == NULL
 * (or for points whose colour has a zero channel) the intensity layer is not written either --
 * LiDAR-only clouds that want the intensity layer must pass a non-zero dummy colour.  Equivalent to SensorProcessorBase::process +
 * ElevationMapping::processpoints (ElevationMapping.cpp:254-283). */
int gem_add_points(gem_map *m, const void *xyzi_device, const void *rgba_device, int n,
                   const gem_frame *frame);
int gem_add_points_host(gem_map *m, const void *xyzi_host, const void *rgba_host, int n,
                        const gem_frame *frame);
/* Stream mode: same result as gem_add_points, but consecutive calls are software-pipelined: call
 * i+1 issues ONE two-node CUDA graph {fold of frame i || bin of frame i+1} on the handle's
 * stream (the fold is mostly a serial tail, the bin kernel is throughput work; per-cell scratch
 * is double-buffered inside the 32-byte cell record).  The fold of the last frame is issued by
 * the next call of any kind that reads or changes the map, by gem_flush or by gem_sync.
 * Contract: the device inputs are read by the bin kernel of this call AND (intensities) by the
 * deferred fold: they must stay valid until two further stream calls have COMPLETED on the
 * stream, or until gem_sync.  n <= max_points.  GEM_B200_PIPE=stream selects two streams +
 * events instead of the graph, GEM_B200_PIPE=off makes this call identical to gem_add_points. */
int gem_add_points_stream(gem_map *m, const void *xyzi_device, const void *rgba_device, int n,
                          const gem_frame *frame);
/* Several clouds in one launch (multi-sensor rigs, BASELINE config 5): the device buffers hold
 * n_segments clouds back to back, cloud s = points [offsets[s], offsets[s+1]) with its own
 * per-frame constants frames[s] (both host arrays, offsets has n_segments+1 entries,
 * n_segments <= 64).  Equivalent to n_segments gem_add_points calls in order (the per-cell
 * order is the global point index), except that `lowest` is updated once for the whole call. */
int gem_add_points_multi(gem_map *m, const void *xyzi_device, const void *rgba_device, int n_segments,
                         const int *offsets, const gem_frame *frames);
/* Pipelined host ingest: like gem_add_points_host but returns without waiting.  The copy runs on
 * a second stream into one of three staging buffers, so frame i+1's H2D overlaps frame i's kernels
 * (which run pipelined like gem_add_points_stream); every call also reads back the counters of the
 * newest frame whose fold has been issued (gem_get_stats gives the last frame's after a drain).  The
 * host buffers must be pinned (gem_host_alloc / cudaHostRegister) and stay untouched until the call
 * after next on this handle, or gem_sync().  n must not exceed max_points. */
int gem_add_points_host_async(gem_map *m, const void *xyzi_pinned, const void *rgba_pinned, int n,
                              const gem_frame *frame);
/* PCL record ingest: n x 32-byte PointXYZRGBICT {x,y,z,pad, b,g,r,a, covariance, intensity,
 * travers} (PointXYZRGBICT.hpp:26-48), host memory, e.g. cloud->points.data(). */
int gem_add_cloud_pcl_host(gem_map *m, const void *points32_host, int n, const gem_frame *frame);

/* ---- unfused reference calls (host arrays, exactly the reference argument meaning) ----
 * Process_points (gpu.cu:1085-1144): outputs key (storage index or -1), var, x_ts, y_ts,
 * z_ts; rejected points give -1 in every output (gpu.cu:443-450).  Also updates `lowest`. */
int gem_process_points(gem_map *m, int *map_index, const float *x, const float *y,
                       const float *z, float *var, float *x_ts, float *y_ts, float *z_ts,
                       int n, const gem_frame *frame);
/* Fuse (gpu.cu:1154-1193) */
int gem_fuse(gem_map *m, int n, const int *index, const int *R, const int *G, const int *B,
             const float *intensity, const float *height, const float *var);

/* Mapvar_update (gpu.cu:1146-1152) */
int gem_var_update(gem_map *m, float var_update);

/* Map_feature (gpu.cu:1256-1302): computes traversability into the map and copies 9
 * row-major storage-indexed layers to host arrays (any may be NULL). */
int gem_map_feature(gem_map *m, float *elevation, float *var, int *R, int *G, int *B,
                    float *rough, float *slope, float *traver, float *intensity);
/* same computation, no host copies (results stay in the device layers) */
int gem_compute_features(gem_map *m);

/* Raytracing (gpu.cu:1304-1318): visibility clean-up + reset of `lowest` */
int gem_raytracing(gem_map *m);

/* Map_optmove (gpu.cu:1215-1233), Map_closeloop (gpu.cu:1235-1254) */
int gem_opt_move(gem_map *m, const float opt_p[2], float height_update, float aligned_out[2]);
int gem_closeloop(gem_map *m, const float update_position[2], float height_update);

/* ---- colourisation of the cloud from the camera image (ElevationMapping.cpp:331-381), the step
 * right before the fusion path.  T_camera: row-major 3x4 "T.camera", T_lidar: row-major 4x4 "T.lidar"
 * (kitti_intrinsic.yaml / yq_intrinsic.yaml), bgr: device BGR8 image.  Writes rgba_out (r,g,b,255 or
 * 0,0,0,0) and zeroes the intensity of points that do not project into the image, exactly like the
 * reference loop; the reference's debug circle drawing into the image (:372) is not reproduced. */
int gem_colourise_points(gem_map *m, void *xyzi_device, int n, const double T_camera[12], const double T_lidar[16],
                         const unsigned char *bgr_device, int width, int height, int row_stride_bytes,
                         void *rgba_out_device);

/* ---- write-back replacing ElevationMap::show's L*L CPU loop (ElevationMap.cpp:85-149) --
 * Emits 9 float32 layers {elevation, variance, rough, slope, traver, color_r, color_g,
 * color_b, intensity} (ElevationMap.cpp:44) in grid_map::Matrix layout: COLUMN-major,
 * storage indexed, NaN where the reference leaves the cell cleared (elevation == -10 or
 * traver == -10 or traver is NaN, ElevationMap.cpp:101).  host_layers[k] may be NULL. */
int gem_export_layers(gem_map *m, float *host_layers[9]);
/* the same in two halves: _begin returns once the copies are under way (pinned host memory!), _end waits for them.
 * Calls that do not change the exported state may be made in between -- the node's next call after show() is
 * Raytracing (ElevationMapping.cpp:404-421) -- so the 36 * L^2 bytes cross PCIe under the ray clean-up. */
int gem_export_layers_begin(gem_map *m, float *host_layers[9]);
int gem_export_layers_end(gem_map *m);

/* The other two products of ElevationMap::show, from the same pass's state (call after gem_compute_features):
 * gem_export_orthomosaic: the bgr8 image of ElevationMap.cpp:87,123-125, L x L x 3 bytes row-major; a shown cell
 *   (ix, iy) is drawn at pixel ((ix + L - start_x) % L, (iy + L - start_y) % L), everything else is black.
 * gem_export_visual_points: the pcl::PointXYZRGB cloud of ElevationMap.cpp:112-121, one point per shown cell in
 *   GridMapIterator order (linear index ix + iy*L): xyz (3 floats/point: grid_map cell-centre position, elevation)
 *   and rgb (3 bytes/point).  *count_out = number of shown cells; min(count, capacity) points are written. */
int gem_export_orthomosaic(gem_map *m, unsigned char *host_bgr);
int gem_export_visual_points(gem_map *m, float *host_xyz, unsigned char *host_rgb, int capacity, int *count_out);

/* ---- scroll-out capture into the submap store (ElevationMapping.cpp:609-765) ----
 * gem_snapshot_shown: prevMap_ = map_.visualMap_ (:422) on the device: the shown state of this frame (after
 *   gem_compute_features, before gem_raytracing) with its geometry; ~20 B/cell device-to-device, no host copy.
 * gem_harvest_scrolled_out: the "L-shape" loop of :716-765 over that snapshot: every cell with traver >= 0 whose
 *   centre lies outside the window current_xy +- length*resolution/2 on the side(s) selected by the signs of
 *   shift_xy (both as returned by the gem_move that followed the snapshot) becomes one 32-byte PointXYZRGBICT
 *   record {x, y, elevation, 1 | bgra, variance, intensity, traver} (:748-759; the same values GridPointData
 *   stores, :736-737), in GridMapIterator order.  *count_out = number of such cells; min(count, capacity) written.
 *   The host-side gate of :716 (|shift| >= resolution, init / jump flags) stays with the caller. */
int gem_snapshot_shown(gem_map *m);
int gem_harvest_scrolled_out(gem_map *m, const float current_xy[2], const float shift_xy[2], void *host_points32,
                             int capacity, int *count_out);

/* raw layer access (row-major L*L, float or int32 for the colour ids) for tests and
 * checkpoint/restore (the dead G_get_mapinfo/G_set_mapinfo of gpu.cu:457-475). */
int gem_get_layer(gem_map *m, int layer, void *host_out);
int gem_set_layer(gem_map *m, int layer, const void *host_in);
int gem_get_state(gem_map *m, float centre[2], int start[2], float *sensor_z);
int gem_get_stats(gem_map *m, gem_stats *out);

/* ---- launch accounting and per-kernel device timing ------------------------------------
 * The library counts every kernel it launches.  With profiling enabled each launch is also
 * bracketed by CUDA events on the handle's stream (costs ~2 us per launch: use a separate
 * pass, not the timed one) and the pipelined add calls fall back to the serial schedule
 * (bin, then fold, nothing overlapped), so the per-kernel durations are uncontended.
 * gem_profile_read synchronises the stream. */
enum {
    GEM_PROF_BIN = 0,       /* k_bin: transform + bin + record store */
    GEM_PROF_FOLD_LONG = 1, /* k_fold_long: the cells with more than 40 records of the call */
    GEM_PROF_UNUSED = 2,
    GEM_PROF_FOLD = 3,      /* k_fold: all other cells */
    GEM_PROF_CLEAR = 4, GEM_PROF_FEATURES = 5, GEM_PROF_RAYTRACE = 6, GEM_PROF_OTHER = 7,
    GEM_PROF_ROUTE = 8, /* tiled maps: the routing kernel */
    GEM_PROF_CLASSES = 9
};
typedef struct gem_profile {
    long long launches;                  /* kernels launched since the last reset          */
    double ms[GEM_PROF_CLASSES];         /* summed device time per kernel class            */
    long long count[GEM_PROF_CLASSES];   /* timed launches per class                       */
} gem_profile;
int gem_profile_enable(gem_map *m, int on);
int gem_profile_read(gem_map *m, gem_profile *out, int reset);

/* self-test: compares the fold's shared-reciprocal division (div2_rn) with the IEEE `/` operator
 * on n pseudo-random operand triples; mismatches must come back 0.  fast_out = how many triples
 * took the fast path. */
int gem_selftest_division(gem_map *m, unsigned long long seed, unsigned long long n,
                          unsigned long long *mismatches_out, unsigned long long *fast_out);

/* pinned host memory helpers for callers that want async-capable staging */
int gem_host_alloc(void **out, unsigned long long bytes);
int gem_host_free(void *p);

/* ---- multi-GPU tiling helpers ---------------------------------------------
 * gem_route_points: transform n device points like gem_add_points but do not fuse; emit
 * routed records {key(global geographic linear index), h, var, rgba, intensity} = 20 B
 * stably bucketed by owning tile (owner = (gx / tile_rows) * tiles_per_row + gy / tile_cols)
 * into rec_out_device, and the per-owner counts into counts_out_device[n_owners].
 * bucket_stride == 0: buckets are packed back to back (split sizes come from the counts);
 * bucket_stride  > 0: bucket o starts at record o*bucket_stride and unused slots hold gkey = -1,
 * so a fixed-size all-to-all needs no host-side split sizes (no stream synchronisation).
 * gem_fuse_records: fold received records (any owner order, already in global order)
 * into this handle's tile. */
int gem_route_points(gem_map *m, const void *xyzi_device, const void *rgba_device, int n,
                     const gem_frame *frame, int tiles_r, int tiles_c, void *rec_out_device,
                     int *counts_out_device, int bucket_stride);
int gem_fuse_records(gem_map *m, const void *rec_device, int n);
/* ---- features / ray clean-up on tiled handles (halo + replicated lowest) ---------------
 * gem_get_layer_device: dense rows*cols copy of one layer into device memory (float, or int32 for the
 *   colour ids; id 10 = the traversability output of the last feature pass), e.g. to cut halo strips.
 * gem_compute_features_tiled: Map_feature's kernel on a tile; padded_elevation_device is the tile's
 *   elevation with a 2-cell halo from the neighbouring tiles, (tile_rows+4) x (tile_cols+4) row-major,
 *   -10 outside the map.  Results equal the untiled map's, cell for cell.
 * gem_raytracing_tiled: Raytracing on a tile; global_lowest_device is the map-wide L x L lowest layer
 *   (every rank's tile of it gathered); resets the OWN tile's lowest to 10 afterwards. */
int gem_get_layer_device(gem_map *m, int layer, void *out_device);
int gem_compute_features_tiled(gem_map *m, const float *padded_elevation_device);
int gem_raytracing_tiled(gem_map *m, const float *global_lowest_device);

/* Peer-memory routing (no collective library on the data path): like gem_route_points with a bucket
 * stride, but every record is stored directly into the OWNING rank's receive buffer through a peer
 * mapping (NVLink/NVSwitch): rank r's records for owner o go to peer_recv[o] + r*bucket_stride, its
 * bucket size to ((int*)peer_counts[o])[r].  peer_recv / peer_counts: n_owners device addresses
 * valid on this handle's device (host arrays).  The caller synchronises the ranks (one barrier)
 * before folding with gem_fuse_records_counted. */
int gem_route_points_peer(gem_map *m, const void *xyzi_device, const void *rgba_device, int n,
                          const gem_frame *frame, int tiles_r, int tiles_c,
                          const unsigned long long *peer_recv, const unsigned long long *peer_counts,
                          int my_rank, int bucket_stride);
/* fold a receive buffer of n_sources buckets of bucket_stride slots, bucket s filled up to
 * src_counts_device[s] */
int gem_fuse_records_counted(gem_map *m, const void *rec_device, const int *src_counts_device, int n_sources,
                             int bucket_stride);

/* ---- loop-closure re-fusion of submaps (ElevationMapping::updateGlobalMap, ElevationMapping.cpp:773-905) ----
 * Submaps are arrays of 32-byte PointXYZRGBICT records in device memory (what gem_harvest_scrolled_out produces).
 * gem_transform_cloud: the rigid re-transform of :805 (pcl::transformPointCloud with T = optimised pose * old pose^-1,
 *   row-major 4 x 4; x' = t00 x + t01 y + t02 z + t03 evaluated left to right in float), in place.
 * gem_refuse_submaps: one pass of the pairwise loop :847-883 for the pair (new = the neighbour, old = submap i): both clouds
 *   are reduced to one point per cell (pointCloudtoHash :1180-1192: cell = (ceil(x / res) * res - res / 2, same for y) in
 *   double -> float, the FIRST point of a cell wins), every cell present in both whose OLD variance lies in (0, 1) gets the
 *   fused elevation / variance in both clouds together with the new cloud's colour, intensity and traversability, and both
 *   clouds come back compacted in place (x, y = the cell's position, w = 1; *n_new / *n_old updated; first-occurrence order).
 *   compat != 0 evaluates the fused values exactly as the reference's expression parses (:862-863: var_n^2 e_o + (var_o^2 e_n) /
 *   var_o^2 + var_n^2, and var_o^2 var_n^2 / var_o^2 + var_n^2), compat == 0 the weighting it was written for
 *   ((var_n^2 e_o + var_o^2 e_n) / (var_o^2 + var_n^2), var_o^2 var_n^2 / (var_o^2 + var_n^2)), all in double like pow().
 *   Definitions where the reference is implementation-defined (unordered_map iteration while erasing / inserting, uninitialised
 *   point fields): DESIGN.md "f4".  Host-synchronous.  The kd-tree neighbour selection of :821-838 stays with the caller
 *   (a handful of submap centres). */
int gem_transform_cloud(gem_map *m, void *points32_device, int n, const float T[16]);
int gem_refuse_submaps(gem_map *m, void *new_points32_device, int *n_new, void *old_points32_device, int *n_old, double resolution,
                       int compat, int *fused_out);

/* ---- tiled maps, peer path: one kernel routes AND exchanges (no collective library, no barrier kernel) ----------
 * The caller allocates, on every rank, four peer-accessible buffers (e.g. CUDA IPC / torch symmetric memory; the
 * library does no inter-process plumbing) and passes the addresses under which THIS device sees every rank's copy:
 *   recv_records   uint4 [5][world * cap]   {global geographic key, height, variance, rgb}
 *   recv_intensity float [5][world * cap]
 *   recv_counts    int   [5][world * cap / 256]
 *   flags          int   [world], zero-initialised before the first step
 * with cap = bucket_capacity rounded up to a multiple of 256 (>= the largest cloud any rank adds per step; world * cap
 * <= max_points).  gem_tiled_step(r) = transform rank r's cloud, store every in-grid point into the OWNING rank's
 * buffer over NVLink (slot = (r * cap / 256 + source block) * 256 + position in the block: deterministic, source
 * order), raise rank r's flag on every peer; then, once every peer's flag of this step is up, bin and fold what
 * arrived.  The result equals the single-GPU map of the rank-by-rank concatenated clouds bit for bit.  Steps are
 * pipelined like gem_add_points_stream: call j issues ONE graph {folds of step j-1 || route -> bin of step j}; the cloud
 * of a call is consumed by that call's graph, and the map contains a step one call later or after gem_flush / gem_sync /
 * any reading call (which issue what is outstanding).  Every rank must make the same sequence of gem_tiled_step calls (a
 * bin waits on the device for every peer's flag of its step).  GEM_B200_TILED_DEPTH=3 selects a three-deep schedule
 * {folds of step j-2 || bin of step j-1 || route of step j} (bit-identical, measured slower; it is what the five
 * buffers are sized for). */
typedef struct gem_tiled_peers {
    int tiles_r, tiles_c, my_rank, bucket_capacity;
    unsigned long long recv_records[64], recv_intensity[64], recv_counts[64], flags[64];
} gem_tiled_peers;
int gem_tiled_attach(gem_map *m, const gem_tiled_peers *peers);
int gem_tiled_step(gem_map *m, const void *xyzi_device, const void *rgba_device, int n, const gem_frame *frame);

#ifdef __cplusplus
}
#endif
#endif /* GEM_B200_H */
