// gpu_process_shim.cpp -- source-level drop-in for GEM's cuda/gpu_process.cu.
//
// Defines the nine C++-mangled free functions the ROS node links against
// (declared ad hoc at ElevationMapping.cpp:44-50, SensorProcessorBase.cpp:34,
// RobotMotionMapUpdater.cpp:18) with their original signatures (gpu_process.cu:940, 1004, 1085,
// 1146, 1154, 1215, 1235, 1256, 1304) and forwards them to libgem_b200.so through the C ABI.
// Needs the Eigen headers, so it is compiled inside the catkin workspace in place of
// cuda/gpu_process.cu; in this repository it is only syntax-checked against
// the stand-in Eigen header of oracle/mini_eigen (tests/test_abi.py).
//
// Like the reference, the map is one per process (the reference keeps it in __device__ globals).
#include <cstdio>
#include <cstring>

#include <Eigen/Core>

#include "gem_b200.h"

namespace {
gem_map *g_map = nullptr;
float g_resolution = 0.0f;

void report(int rc, const char *what)
{ // the reference prints CUDA errors to stderr and carries on (gpu_process.cu:987-992)
    if (rc != GEM_OK) std::fprintf(stderr, "%s failed: %s\n", what, gem_last_error(g_map));
}
} // namespace

void Init_GPU_elevationmap(int length, float resolution, float h_mahalanobisDistanceThreshold_, float h_obstacle_threshold)
{
    gem_config c;
    std::memset(&c, 0, sizeof c);
    c.length = length;
    c.resolution = resolution;
    c.mahalanobis_threshold = h_mahalanobisDistanceThreshold_;
    c.obstacle_threshold = h_obstacle_threshold;
    c.compat_box_filter = 1; // gpu_process.cu:393
    c.device = -1;
    if (g_map) gem_destroy(g_map);
    g_map = nullptr;
    g_resolution = resolution;
    if (gem_create(&c, &g_map) != GEM_OK) std::fprintf(stderr, "Init_GPU_elevationmap failed: %s\n", gem_last_error(nullptr));
}

void Move(float *current_Position, float /*resolution*/, int /*length*/, float *Central_coordinate, int *Start_indice,
          float *alignedPositionShift)
{
    report(gem_move(g_map, current_Position, Central_coordinate, Start_indice, alignedPositionShift), "Move");
}

int Process_points(int *map_index, float *point_x, float *point_y, float *point_z, float *point_var, float *point_x_ts,
                   float *point_y_ts, float *point_z_ts, Eigen::Matrix4f transform, int point_num,
                   double relativeLowerThreshold, double relativeUpperThreshold, float min_r, float beam_a, float beam_c,
                   Eigen::RowVector3f sensorJacobian, Eigen::Matrix3f rotationVariance, Eigen::Matrix3f C_SB_transpose,
                   Eigen::RowVector3f P_mul_C_BM_transpose, Eigen::Matrix3f B_r_BS_skew)
{
    gem_frame f;
    std::memset(&f, 0, sizeof f);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) f.T[4 * i + j] = transform(i, j);
    for (int i = 0; i < 3; i++) {
        f.sensor_jacobian[i] = sensorJacobian(0, i);
        f.P_mul_C_BM_transpose[i] = P_mul_C_BM_transpose(0, i);
        for (int j = 0; j < 3; j++) {
            f.rotation_variance[3 * i + j] = rotationVariance(i, j);
            f.C_SB_transpose[3 * i + j] = C_SB_transpose(i, j);
            f.B_r_BS_skew[3 * i + j] = B_r_BS_skew(i, j);
        }
    }
    f.rel_lower = relativeLowerThreshold;
    f.rel_upper = relativeUpperThreshold;
    f.sensor.type = GEM_SENSOR_LASER;
    f.sensor.min_radius = min_r;
    f.sensor.beam_angle = beam_a;
    f.sensor.beam_constant = beam_c;
    f.sensor.normal_factor_e = 1.0;
    report(gem_process_points(g_map, map_index, point_x, point_y, point_z, point_var, point_x_ts, point_y_ts, point_z_ts,
                              point_num, &f),
           "Process_points");
    return 0; // gpu_process.cu:1143
}

void Fuse(int /*length*/, int point_num, int *point_index, int *point_colorR, int *point_colorG, int *point_colorB,
          float *point_intensity, float *point_height, float *point_var)
{
    report(gem_fuse(g_map, point_num, point_index, point_colorR, point_colorG, point_colorB, point_intensity, point_height,
                    point_var),
           "Fuse");
}

void Mapvar_update(int /*length*/, float var_update) { report(gem_var_update(g_map, var_update), "Mapvar_update"); }

void Map_feature(int /*length*/, float *elevation, float *var, int *colorR, int *colorG, int *colorB, float *rough,
                 float *slope, float *traver, float *intensity)
{
    report(gem_map_feature(g_map, elevation, var, colorR, colorG, colorB, rough, slope, traver, intensity), "Map_feature");
}

void Raytracing(int /*length*/) { report(gem_raytracing(g_map), "Raytracing"); }

void Map_optmove(float *opt_p, float height_update, float /*resolution*/, int /*length*/, float *opt_alignedPosition)
{
    report(gem_opt_move(g_map, opt_p, height_update, opt_alignedPosition), "Map_optmove");
}

void Map_closeloop(float *update_position, float height_update, int /*length*/, float /*resolution*/)
{
    report(gem_closeloop(g_map, update_position, height_update), "Map_closeloop");
}
