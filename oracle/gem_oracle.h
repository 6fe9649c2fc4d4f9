/*
 * gem_oracle.h -- CPU ORACLE for the GEM point-cloud -> elevation-grid fusion path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, link, import or run it.
 * The product (gem_b200/, include/) never touches anything under oracle/.
 *
 * PARITY STATUS: PINNED AGAINST THE REFERENCE ITSELF.  The reference (ZJU-Robotics-Lab/GEM
 * @ d7ec953) ships no tests, no golden vectors and no CPU implementation of this path
 * (SURVEY.md section 0, 4, 8c); its only statement of the algorithm is the CUDA file
 * elevation_mapping/elevation_mapping/cuda/gpu_process.cu ("gpu.cu" below).  That file needs
 * Eigen, which this image lacks, so oracle/build_ref.py compiles it UNMODIFIED from
 * /root/reference against a ~120-line stand-in for the Eigen features it uses
 * (oracle/mini_eigen) plus an extern "C" harness (oracle/ref_harness.cu) into oracle/_ref/.
 * The oracle is checked against those reference kernels run on a B200:
 *   - tests/test_reference_pin.py (GPU): bit-identical map_index / height / variance /
 *     transformed x,y / fused layers / Move outputs vs the -fmad=false build; within 1e-5
 *     relative of the build with the reference's own flags (FMA contraction);
 *   - tests/golden/gem_golden_v1.npz: outputs of the reference itself, generated on a B200 by
 *     tests/golden/make_golden.py; tests/test_golden.py (CPU) replays them against the oracle.
 * Not pinned by the reference (defined here instead, each marked ORACLE DEFINITION): the racy
 * `lowest` update of gpu.cu:434-438, the last bit of CUDA libm's trig in the feature kernel,
 * output values the reference leaves uninitialised, Move for shifts <= -length.
 * Further cross-checks: a literal O(C*N) twin of G_fuse vs the O(N) form, and an independent
 * numpy float32 re-derivation (tests/test_oracle_vs_numpy.py).
 *
 * Every function cites the reference lines it follows.  Conscious definitions where
 * the reference is racy / undefined are marked "ORACLE DEFINITION".
 */
#ifndef GEM_ORACLE_H
#define GEM_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_map {
    int L;                    /* cells per side (gpu.cu:35 Length)            */
    float res;                /* gpu.cu:36 Resolution                         */
    float obstacle_threshold; /* gpu.cu:37                                    */
    float mahalanobis;        /* gpu.cu:56, uploaded but unused (gate is 5)   */
    /* layers, row-major L*L (gpu.cu:20-28) */
    float *lowest;            /* indexed by GEOGRAPHIC linear index           */
    float *elevation, *variance, *intensity, *traver; /* indexed by STORAGE   */
    int *colorR, *colorG, *colorB;
    float centre[2];          /* gpu.cu:30 central_coordinate                 */
    int start[2];             /* gpu.cu:31 start_indice                       */
    float sensorZ;            /* gpu.cu:33 sensorZatLowestScan                */
    int compat_box_filter;    /* 1 = hard-coded filter of gpu.cu:393          */
} orc_map;

/* sensor model selector */
enum { ORC_SENSOR_LASER = 0, ORC_SENSOR_STRUCTURED_LIGHT = 1 };

typedef struct orc_sensor {
    int type;
    /* laser (gpu.cu:410-411, Laser.cpp:134-163) */
    float min_r, beam_a, beam_c;
    /* structured light (SL.cpp:129-141), parameters are double in the reference */
    double nf_a, nf_b, nf_c, nf_d, nf_e, lateral;
    /* structured light cleanPointCloud pass-through limits (SL.cpp:39-40,51-66) */
    double cutoff_min, cutoff_max;
} orc_sensor;

/* SensorProcessorBase::process -> cleanPointCloud (SPB.cpp:90), run before Process_points.
 * laser: pcl::removeNaNFromPointCloud (Laser.cpp:50-59) drops points with a non-finite x, y or z;
 * structured light: pcl::PassThrough on field "z" with limits (float)cutoff_min/max (SL.cpp:51-66)
 * drops non-finite points and points with z < min || z > max.
 * PCL is an un-vendored, unpinned dependency of the reference: both filters are restated from
 * PCL's published algorithm (filters/impl/passthrough.hpp, filters/impl/filter.hpp) -> this
 * function is PARITY UNPINNED.  Compacts xyzi (n x 4 floats) and rgba (n x 4 bytes, may be NULL)
 * in place, order preserved; returns the number of points kept. */
int orc_clean_point_cloud(const orc_sensor *sensor, int n, float *xyzi, unsigned char *rgba);

/* gpu.cu:940-994 + G_Init_map :198-214 */
orc_map *orc_create(int length, float resolution, float mahalanobis, float obstacle_threshold);
void orc_destroy(orc_map *m);

/* gpu.cu:1004-1083 Move (+ :893-938, :996-1002) */
void orc_move(orc_map *m, const float pos[3], float centre_out[2], int start_out[2],
              float shift_out[2]);

/* gpu.cu:309-358: returns geographic linear index or -1; *storage gets PointsToMapIndex */
int orc_points_to_index(const orc_map *m, float px, float py, int *storage);

/* gpu.cu:1085-1144 Process_points + G_pointsprocess :384-455.
 * T: row-major 4x4 map<-sensor.  rotVar, C_SB_T, B_skew: row-major 3x3.
 * Outputs (length n): key (storage index or -1), var, x_ts, y_ts, z_ts (any may be NULL).
 * Updates m->lowest (ORACLE DEFINITION, see .c). */
void orc_process_points(orc_map *m, int n, const float *x, const float *y, const float *z,
                        const float T[16], double relLower, double relUpper,
                        const orc_sensor *sensor, const float sJ[3], const float rotVar[9],
                        const float C_SB_T[9], const float P_C_BM_T[3], const float B_skew[9],
                        int *key, float *var, float *x_ts, float *y_ts, float *z_ts);

/* gpu.cu:1154-1193 Fuse + G_fuse :477-537, O(N) in-order scatter form */
void orc_fuse(orc_map *m, int n, const int *key, const int *R, const int *G, const int *B,
              const float *intensity, const float *h, const float *var);
/* literal O(C*N) twin of G_fuse (one "thread" per cell scanning all points) */
void orc_fuse_literal(orc_map *m, int n, const int *key, const int *R, const int *G,
                      const int *B, const float *intensity, const float *h, const float *var);

/* gpu.cu:1146-1152 + :540-547 */
void orc_var_update(orc_map *m, float dv);

/* gpu.cu:1256-1302 + G_Mapfeature :549-670 + computerEigenvalue :66-187.
 * Output arrays are STORAGE indexed, length L*L.  For empty cells the reference leaves
 * rough/slope/traver outputs uninitialised; ORACLE DEFINITION: rough=0, slope=0, traver=-10
 * in the OUTPUT arrays (map state traver is left stale exactly like the reference). */
void orc_map_feature(orc_map *m, float *elevation, float *var, int *R, int *G, int *B,
                     float *rough, float *slope, float *traver, float *intensity);

/* gpu.cu:1304-1318 + G_Raytracing :708-891 + G_Clear_maplowest :232-239 */
void orc_raytracing(orc_map *m);

/* gpu.cu:1215-1233, :1235-1254 */
void orc_optmove(orc_map *m, const float opt_p[2], float height_update, float aligned_out[2]);
void orc_closeloop(orc_map *m, const float update_position[2], float height_update);

/* ElevationMapping.cpp:331-381: colourise the cloud from a BGR8 image.  xyzi is n x {x,y,z,intensity}
 * (intensity zeroed for points that do not project into the image), rgba_out n x {r,g,b,a}.
 * NOT pinned by reference-generated vectors (the reference loop needs OpenCV/ROS); the debug circle the
 * reference draws into the image after each lookup (:372) is deliberately not restated. */
void orc_colourise(float *xyzi, int n, const double T_camera[12], const double T_lidar[16], const unsigned char *bgr,
                   int width, int height, int row_stride, unsigned char *rgba_out);

/* ElevationMap.cpp:85-149 show(): bgr8 orthomosaic (L*L*3) + visual cloud (xyz 3 floats, rgb 3 bytes per shown
 * cell, GridMapIterator order) from Map_feature's outputs; any output may be NULL.  The cell-centre position follows
 * grid_map's getPositionFromIndex (un-vendored, unpinned: ORACLE DEFINITION, see .c). */
void orc_show(const orc_map *m, double grid_res /* the node's double resolution_, 0 = (double)m->res */,
              const float *elevation, const float *traver, const int *R, const int *G, const int *B,
              unsigned char *bgr, float *xyz, unsigned char *rgb, int *count);

/* ElevationMapping.cpp:716-765: cells of the previous frame's shown map (Map_feature outputs + that frame's centre and
 * start index) that lie outside the current window -> n x 8 floats (PointXYZRGBICT records), GridMapIterator order. */
void orc_harvest(int L, double grid_res, const float centre_prev[2], const int start_prev[2], const float *elevation,
                 const float *variance, const float *traver, const int *R, const int *G, const int *B,
                 const float *intensity, const float current[2], const float shift[2], float *out, int *count);

/* deterministic float trig used by the feature kernel restatement (see .c) */
float orc_sinf(float a);
float orc_cosf(float a);
float orc_atan2f(float y, float x);
float orc_acosf(float x);

/* multi-threaded twin used only as the CPU baseline timer (bench.py cpu_baseline):
 * process_points + fuse over nthreads, cells partitioned by storage row range so the
 * per-cell order is preserved.  Result identical to the single-thread path. */
void orc_add_points_mt(orc_map *m, int n, const float *xyzi /* n*4 */, const unsigned char *rgba /* n*4 or NULL */,
                       const float T[16], double relLower, double relUpper,
                       const orc_sensor *sensor, const float sJ[3], int nthreads);

/* ElevationMapping::updateGlobalMap (ElevationMapping.cpp:773-905), PARITY UNPINNED: rigid re-transform of a submap and one
 * pairwise re-fusion (new = the neighbour, old = submap i); points are 8 floats {x,y,z,w,bgra,covariance,intensity,travers};
 * returns the number of fused cells, clouds compacted in place */
void orc_transform_cloud(float *pts, int n, const float T[16]);
int orc_refuse_submaps(float *pn, int *n_new, float *po, int *n_old, double res, int compat);

/* pooled variant of orc_add_points_mt: persistent worker threads, per-band point lists, dynamic band assignment
 * (bench.py's CPU baseline; same results) */
typedef struct orc_pool orc_pool;
orc_pool *orc_pool_create(int nthreads);
void orc_pool_destroy(orc_pool *p);
void orc_add_points_pool(orc_pool *p, orc_map *m, int n, const float *xyzi, const unsigned char *rgba, const float T[16],
                         double relLower, double relUpper, const orc_sensor *sensor, const float sJ[3]);

#ifdef __cplusplus
}
#endif
#endif
