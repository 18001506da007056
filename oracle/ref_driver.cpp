// oracle/ref_driver.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Thin C-ABI around the UNMODIFIED reference class so Python (ctypes) can call it.  Linked
// together with the reference's own translation units, compiled in place from /root/reference
// (oracle/Makefile), into oracle/_ref/libref_{strict,fast}.so.  Nothing from the reference is
// copied into this repository; this file only *calls* its public interface:
//   recursive_patchwork::RecursivePatchwork::filterGroundPoints
//       RP/include/recursive_patchwork.hpp:53-54, RP/src/recursive_patchwork.cpp:310-426
//   recursive_patchwork::PatchworkConfig        RP/include/recursive_patchwork.hpp:25-36
//   recursive_patchwork::RecursivePatchwork::sampleGroundAndObstacles
//       RP/include/recursive_patchwork.hpp:56-59, RP/src/recursive_patchwork.cpp:428-465
//   recursive_patchwork::LidarFusion::fuseLidarPointClouds   RP/src/lidar_fusion.cpp:42-86
//   recursive_patchwork::Visualization::createBEVImage / createGroundNonGroundImage
//       RP/include/visualization.hpp:16-27, RP/src/visualization.cpp:18-80
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load the resulting library.
#include "recursive_patchwork.hpp"
#include "lidar_fusion.hpp"
#include "visualization.hpp"  // against tests/ref_build/opencv2/opencv.hpp (cv::Mat stand-in; OpenCV's C++ headers are absent here)

#include <chrono>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <vector>

using recursive_patchwork::PatchworkConfig;
using recursive_patchwork::Point3D;
using recursive_patchwork::RecursivePatchwork;

extern "C" {

// Plain-C mirror of PatchworkConfig (RP/include/recursive_patchwork.hpp:25-36), field for field.
struct rpwref_config {
    float sensor_height;
    float max_range;
    int32_t num_sectors;
    int32_t max_iter;
    int32_t adaptive_seed_height;
    float th_seeds;
    float th_dist;
    float th_outlier;
    float filtering_radius;
    int32_t max_split_depth;
};

static PatchworkConfig to_cfg(const rpwref_config* c) {
    PatchworkConfig cfg;
    if (c) {
        cfg.sensor_height = c->sensor_height;
        cfg.max_range = c->max_range;
        cfg.num_sectors = c->num_sectors;
        cfg.max_iter = c->max_iter;
        cfg.adaptive_seed_height = c->adaptive_seed_height != 0;
        cfg.th_seeds = c->th_seeds;
        cfg.th_dist = c->th_dist;
        cfg.th_outlier = c->th_outlier;
        cfg.filtering_radius = c->filtering_radius;
        cfg.max_split_depth = c->max_split_depth;
    }
    return cfg;
}

// The reference prints 3-5 lines per cuda::ops call.  Putting std::cout into a
// failed state turns every operator<< into a no-op; nothing else in the process is affected.
void rpwref_silence(int on) {
    if (on) std::cout.setstate(std::ios_base::failbit);
    else std::cout.clear();
}

void rpwref_default_config(rpwref_config* out) {
    PatchworkConfig d;
    out->sensor_height = d.sensor_height;
    out->max_range = d.max_range;
    out->num_sectors = d.num_sectors;
    out->max_iter = d.max_iter;
    out->adaptive_seed_height = d.adaptive_seed_height ? 1 : 0;
    out->th_seeds = d.th_seeds;
    out->th_dist = d.th_dist;
    out->th_outlier = d.th_outlier;
    out->filtering_radius = d.filtering_radius;
    out->max_split_depth = d.max_split_depth;
}

static inline bool same_bits(const Point3D& a, const float* b) {
    return std::memcmp(&a.x, b, 4) == 0 && std::memcmp(&a.y, b + 1, 4) == 0 && std::memcmp(&a.z, b + 2, 4) == 0;
}

// Runs the reference on one cloud.  xyz: n points, `stride` floats apart (3 or 4).
// ground_out / nonground_out: caller buffers of 3*n floats (may be NULL).
// labels_out (may be NULL): per INPUT index 0 non-ground in zone, 1 ground, 2 beyond the
// filtering radius, 3 dropped as non-finite; reconstructed by walking the two returned clouds
// (their order is fixed by RP/src/recursive_patchwork.cpp:402-419).  *ambiguous counts input
// points whose xyz bits matched the heads of both clouds (duplicates; labelled ground then).
// Returns 0, or -1 if the returned clouds are inconsistent with the input.
int rpwref_filter_ground(const rpwref_config* c, const float* xyz, size_t n, size_t stride,
                         float* ground_out, size_t* n_ground, float* nonground_out, size_t* n_nonground,
                         uint8_t* labels_out, size_t* ambiguous) {
    std::vector<Point3D> pts(n);
    for (size_t i = 0; i < n; ++i) {
        pts[i].x = xyz[i * stride + 0];
        pts[i].y = xyz[i * stride + 1];
        pts[i].z = xyz[i * stride + 2];
    }
    RecursivePatchwork rp(to_cfg(c));
    auto result = rp.filterGroundPoints(pts);
    const auto& g = result.first;
    const auto& ng = result.second;
    if (n_ground) *n_ground = g.size();
    if (n_nonground) *n_nonground = ng.size();
    if (ground_out) for (size_t i = 0; i < g.size(); ++i) { ground_out[3*i] = g[i].x; ground_out[3*i+1] = g[i].y; ground_out[3*i+2] = g[i].z; }
    if (nonground_out) for (size_t i = 0; i < ng.size(); ++i) { nonground_out[3*i] = ng[i].x; nonground_out[3*i+1] = ng[i].y; nonground_out[3*i+2] = ng[i].z; }
    if (ambiguous) *ambiguous = 0;
    if (!labels_out) return 0;

    // Beyond-radius points form the tail of the non-ground cloud, in input order
    // (RP/src/recursive_patchwork.cpp:415-419).  The walk below needs to know which input
    // points those are; it re-evaluates the reference's own predicate sqrt(x*x+y*y) > R on
    // finite points (RP/cuda/cuda_interface.cu:590, RP/src/recursive_patchwork.cpp:416).  This
    // file is always compiled with strict IEEE flags (oracle/Makefile), and every re-evaluated
    // decision is verified against the returned clouds bit-for-bit (-1 on any mismatch).
    size_t gi = 0, ni = 0, amb = 0;
    const float R = to_cfg(c).filtering_radius;
    size_t n_tail = 0;
    for (size_t i = 0; i < n; ++i) {
        const float x = pts[i].x, y = pts[i].y, z = pts[i].z;
        if (!(std::isfinite(x) && std::isfinite(y) && std::isfinite(z))) continue;
        const float d = std::sqrt(x * x + y * y);
        if (d > R) n_tail++;
    }
    if (n_tail > ng.size()) return -1;
    size_t ti = ng.size() - n_tail;  // cursor into the beyond-radius tail
    const size_t zone_ng_end = ti;
    // Degenerate return ({}, cleaned): fewer than 3 in-zone points (RP/src/recursive_patchwork.cpp:339-341)
    // — then ng holds ALL cleaned points in input order and there is no separate tail.
    size_t n_clean = 0;
    for (size_t i = 0; i < n; ++i)
        if (std::isfinite(pts[i].x) && std::isfinite(pts[i].y) && std::isfinite(pts[i].z)) n_clean++;
    const bool degenerate = g.empty() && ng.size() == n_clean && (n_clean - n_tail) < 3;
    for (size_t i = 0; i < n; ++i) {
        const float* p = xyz + i * stride;
        if (!(std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]))) { labels_out[i] = 3; continue; }
        const bool out = std::sqrt(p[0] * p[0] + p[1] * p[1]) > R;
        if (degenerate) { labels_out[i] = out ? 2 : 0; continue; }
        if (out) {
            if (ti >= ng.size() || !same_bits(ng[ti], p)) return -1;
            ti++;
            labels_out[i] = 2;
            continue;
        }
        const bool mg = gi < g.size() && same_bits(g[gi], p);
        const bool mn = ni < zone_ng_end && same_bits(ng[ni], p);
        if (mg && mn) amb++;
        if (mg) { labels_out[i] = 1; gi++; }
        else if (mn) { labels_out[i] = 0; ni++; }
        else return -1;
    }
    if (!degenerate && (gi != g.size() || ni != zone_ng_end || ti != ng.size())) return -1;
    if (ambiguous) *ambiguous = amb;
    return 0;
}

// The reference's multi-LiDAR fusion front end: LidarFusion::fuseLidarPointClouds
// (RP/include/lidar_fusion.hpp:21-22, RP/src/lidar_fusion.cpp:42-86) with one LidarConfig per sensor.
// fused_xyz: caller buffer of 3 * (sum of n) floats.  Returns the number of fused points.
struct rpwref_sensor {
    const float* xyz;
    size_t n;
    float rotation_deg;
    float ego_radius;
};

size_t rpwref_fuse(const rpwref_sensor* sensors, size_t n_sensors, size_t stride, float* fused_xyz) {
    recursive_patchwork::LidarFusion fusion;
    fusion.clearLidars();
    std::vector<std::vector<Point3D>> clouds(n_sensors);
    for (size_t s = 0; s < n_sensors; ++s) {
        recursive_patchwork::LidarConfig lc{static_cast<int>(s + 1), "", sensors[s].rotation_deg, sensors[s].ego_radius};
        fusion.addLidar(lc);
        clouds[s].resize(sensors[s].n);
        for (size_t i = 0; i < sensors[s].n; ++i) {
            clouds[s][i].x = sensors[s].xyz[i * stride];
            clouds[s][i].y = sensors[s].xyz[i * stride + 1];
            clouds[s][i].z = sensors[s].xyz[i * stride + 2];
        }
    }
    const std::vector<Point3D> fused = fusion.fuseLidarPointClouds(clouds);
    for (size_t i = 0; i < fused.size(); ++i) { fused_xyz[3 * i] = fused[i].x; fused_xyz[3 * i + 1] = fused[i].y; fused_xyz[3 * i + 2] = fused[i].z; }
    return fused.size();
}

// The CLI's post-filter, RecursivePatchwork::sampleGroundAndObstacles (RP/src/recursive_patchwork.cpp:428-465):
// the result is [random ground context sample (min(2000, #ground) points, unseeded std::mt19937) | obstacles].
// out_xyz: 3 * cap floats.  Returns the number of points (or (size_t)-1 if cap is too small).
size_t rpwref_sample_ground_and_obstacles(const rpwref_config* c, const float* xyz, size_t n, size_t stride,
                                          float target_height, float base_tol, float* out_xyz, size_t cap) {
    std::vector<Point3D> pts(n);
    for (size_t i = 0; i < n; ++i) { pts[i].x = xyz[i * stride]; pts[i].y = xyz[i * stride + 1]; pts[i].z = xyz[i * stride + 2]; }
    RecursivePatchwork rp(to_cfg(c));
    const std::vector<Point3D> res = rp.sampleGroundAndObstacles(pts, target_height, base_tol);
    if (res.size() > cap) return (size_t)-1;
    for (size_t i = 0; i < res.size(); ++i) { out_xyz[3 * i] = res[i].x; out_xyz[3 * i + 1] = res[i].y; out_xyz[3 * i + 2] = res[i].z; }
    return res.size();
}

// The CLI's rasters (RP/src/visualization.cpp:18-80) through the cv::Mat stand-in.  mode 0:
// createGroundNonGroundImage(a, b); mode 1: createBEVImage(a) (b ignored).  bgr_out: height * width * 3 bytes.
int rpwref_bev(int mode, const float* a, size_t na, const float* b, size_t nb, int width, int height,
               float x_min, float y_min, float x_max, float y_max, uint8_t* bgr_out) {
    auto cloud = [](const float* p, size_t n) {
        std::vector<Point3D> v(n);
        for (size_t i = 0; i < n; ++i) { v[i].x = p[3 * i]; v[i].y = p[3 * i + 1]; v[i].z = p[3 * i + 2]; }
        return v;
    };
    recursive_patchwork::Visualization viz;
    cv::Mat img = mode == 0 ? viz.createGroundNonGroundImage(cloud(a, na), cloud(b, nb), width, height, x_min, y_min, x_max, y_max)
                            : viz.createBEVImage(cloud(a, na), width, height, x_min, y_min, x_max, y_max);
    if (img.rows != height || img.cols != width) return -1;
    std::memcpy(bgr_out, img.data(), (size_t)width * height * 3);
    return 0;
}

// Times `reps` back-to-back calls of filterGroundPoints on one cloud (seconds per call,
// wall clock, conversion to std::vector<Point3D> excluded).  Thread-safe: every call builds
// its own RecursivePatchwork (the class holds only config_, RP/include/recursive_patchwork.hpp:70).
double rpwref_time_scan(const rpwref_config* c, const float* xyz, size_t n, size_t stride, int reps,
                        size_t* n_ground_last) {
    std::vector<Point3D> pts(n);
    for (size_t i = 0; i < n; ++i) {
        pts[i].x = xyz[i * stride + 0];
        pts[i].y = xyz[i * stride + 1];
        pts[i].z = xyz[i * stride + 2];
    }
    RecursivePatchwork rp(to_cfg(c));
    size_t ng = 0;
    auto t0 = std::chrono::steady_clock::now();
    for (int r = 0; r < reps; ++r) {
        auto res = rp.filterGroundPoints(pts);
        ng = res.first.size();
    }
    auto t1 = std::chrono::steady_clock::now();
    if (n_ground_last) *n_ground_last = ng;
    return std::chrono::duration<double>(t1 - t0).count() / (reps > 0 ? reps : 1);
}

}  // extern "C"
