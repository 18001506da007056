// host/recursive_patchwork.hpp — drop-in C++ front end for the B200 segmentation path.
//
// Source-compatible with the public surface of the reference's
// src/recursive_patchwork/include/recursive_patchwork.hpp (namespace, type names, member
// functions, argument meaning, default values), so the reference's callers — main.cpp:193,268,286,
// recursive_patchwork_node.cpp:40,91 and test_recursive_patchwork.cpp:65-68,86-89,151-155 — build
// against it unchanged.  It REPLACES that file in place (the reference's other headers include it by
// quote, `#include "recursive_patchwork.hpp"`, which resolves to their own directory before any -I
// path): tests/test_gpu_refbuild.py builds the reference's own
// test/test_recursive_patchwork.cpp and src/main.cpp, unmodified, against it.  Differences:
//   * the class holds an opaque C-ABI handle (include/rpw_b200.h) instead of running the algorithm
//     on the host; librpw_b200.so must be linked.  Nothing here needs Eigen, but the reference's
//     header includes <Eigen/Dense> and its dependants rely on that (lidar_fusion.hpp:36,50-51 uses
//     Eigen::Matrix4f without including it), so it is included when the build has it;
//   * copies of a RecursivePatchwork share nothing: a copy carries the configuration (all the
//     reference's class holds, recursive_patchwork.hpp:70) and creates its own device handle on
//     first use;
//   * failures (no CUDA device, CUDA error, capacity) surface as std::runtime_error — the ROS2
//     node already wraps its callback in try/catch(std::exception) (recursive_patchwork_node.cpp:66,105);
//     there is no CPU fallback;
//   * filterGroundPoints has a labels-returning sibling (the north-star addition);
//   * nothing is printed on the hot path (the reference logs 3-5 lines per cuda::ops call).
// Header-only on purpose: the host side stays trivial and the whole product is behind the C-ABI.
#pragma once

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#if defined(__has_include)
#if __has_include(<Eigen/Dense>)
#include <Eigen/Dense>  // transitive include the reference's dependants rely on (lidar_fusion.hpp, cuda_interface.hpp)
#endif
#endif

#include "rpw_b200.h"

namespace recursive_patchwork {

class PointCloudProcessor;
class LidarFusion;
class RosbagLoader;
class Visualization;

// 12-byte AoS point, the layout the C-ABI takes with stride_bytes = 12.
struct Point3D {
    float x, y, z;
    Point3D() : x(0), y(0), z(0) {}
    Point3D(float px, float py, float pz) : x(px), y(py), z(pz) {}
};
static_assert(sizeof(Point3D) == 12, "Point3D must stay a packed xyz triple");

// Same fields, order and defaults as the reference struct; convertible to rpw_config.
struct PatchworkConfig {
    float sensor_height = 1.2f;
    float max_range = 150.0f;         // carried, never read by the path (as in the reference)
    int num_sectors = 10;
    int max_iter = 100;
    bool adaptive_seed_height = true;
    float th_seeds = 0.15f;
    float th_dist = 0.2f;
    float th_outlier = 0.08f;         // carried, never read by the path
    float filtering_radius = 150.0f;
    int max_split_depth = 1000;

    rpw_config toC() const {
        rpw_config c;
        c.sensor_height = sensor_height; c.max_range = max_range; c.num_sectors = num_sectors; c.max_iter = max_iter;
        c.adaptive_seed_height = adaptive_seed_height ? 1 : 0; c.th_seeds = th_seeds; c.th_dist = th_dist;
        c.th_outlier = th_outlier; c.filtering_radius = filtering_radius; c.max_split_depth = max_split_depth;
        return c;
    }
};

struct LidarConfig {
    int lidar_id;
    std::string topic_name;
    float rotation_angle;  // degrees
    float ego_radius = 2.5f;
};

class RecursivePatchwork {
public:
    RecursivePatchwork(const PatchworkConfig& config = PatchworkConfig{}, int device = 0)
        : config_(config), device_(device) {}
    ~RecursivePatchwork() { release(); }
    // Copyable like the reference's class, whose only state is the configuration: a copy gets the
    // configuration and the device index and builds its own handle when it is first used.
    RecursivePatchwork(const RecursivePatchwork& o) : config_(o.config_), device_(o.device_) {}
    RecursivePatchwork& operator=(const RecursivePatchwork& o) {
        if (this != &o) {
            if (device_ != o.device_) release();
            device_ = o.device_;
            setConfig(o.config_);
        }
        return *this;
    }
    RecursivePatchwork(RecursivePatchwork&& o) noexcept
        : config_(o.config_), device_(o.device_), handle_(o.handle_), capacity_(o.capacity_) { o.handle_ = nullptr; o.capacity_ = 0; }
    RecursivePatchwork& operator=(RecursivePatchwork&& o) noexcept {
        if (this != &o) {
            release();
            config_ = o.config_; device_ = o.device_; handle_ = o.handle_; capacity_ = o.capacity_;
            o.handle_ = nullptr; o.capacity_ = 0;
        }
        return *this;
    }

    // ---- main processing --------------------------------------------------------------------
    // (ground, non-ground) exactly as the reference orders them: ground in input order;
    // non-ground in input order followed by the beyond-radius points in input order.
    std::pair<std::vector<Point3D>, std::vector<Point3D>> filterGroundPoints(const std::vector<Point3D>& points) {
        return segmentClouds(points, nullptr);
    }

    // Same, also returning one label per input point (RPW_LABEL_*).
    std::pair<std::vector<Point3D>, std::vector<Point3D>> filterGroundPoints(const std::vector<Point3D>& points,
                                                                             std::vector<std::uint8_t>& labels) {
        labels.resize(points.size());
        return segmentClouds(points, labels.empty() ? nullptr : labels.data());
    }

    // Multi-LiDAR frame with the fusion front end on the device (what main.cpp:245-268 does with
    // LidarFusion::fuseLidarPointClouds followed by filterGroundPoints): clouds[i] is sensor i's cloud in
    // its own frame, configs[i] its LidarConfig (rotation_angle in degrees, ego_radius).  Returns the
    // ground / non-ground clouds of the MERGED frame in vehicle coordinates, in the reference's order;
    // labels (optional) receives one vector per sensor (RPW_LABEL_*; RPW_LABEL_EGO = removed).
    std::pair<std::vector<Point3D>, std::vector<Point3D>> filterGroundPointsFused(
        const std::vector<std::vector<Point3D>>& clouds, const std::vector<LidarConfig>& configs,
        std::vector<std::vector<std::uint8_t>>* labels = nullptr) {
        std::pair<std::vector<Point3D>, std::vector<Point3D>> out;
        const std::size_t k = std::min(clouds.size(), configs.size());
        std::vector<std::vector<std::uint8_t>> local(k);
        std::vector<rpw_sensor_cloud> sens(k);
        std::vector<std::uint8_t*> lp(k);
        std::size_t total = 0;
        for (std::size_t i = 0; i < k; ++i) {
            local[i].assign(clouds[i].size(), RPW_LABEL_DROPPED);
            sens[i].xyz = clouds[i].empty() ? nullptr : &clouds[i][0].x;
            sens[i].n = clouds[i].size();
            sens[i].rotation_deg = configs[i].rotation_angle;
            sens[i].ego_radius = configs[i].ego_radius;
            lp[i] = local[i].data();
            total += clouds[i].size();
        }
        if (total) {
            ensure(total);
            check(rpw_segment_fused(handle_, sens.data(), k, sizeof(Point3D), lp.data(), nullptr));
        }
        // the two clouds in vehicle-frame coordinates, assembled on the device (rotation with the
        // operations LidarFusion::applyRotation2D uses, ego points removed)
        if (total) {
            static_assert(sizeof(Point3D) == 3 * sizeof(float), "Point3D must be packed xyz");
            out.first.resize(total);
            out.second.resize(total);
            std::uint64_t counts[2] = {0, 0};
            check(rpw_last_clouds(handle_, &out.first[0].x, &out.second[0].x, 0, counts));
            out.first.resize(static_cast<std::size_t>(counts[0]));
            out.second.resize(static_cast<std::size_t>(counts[1]));
        }
        if (labels) *labels = std::move(local);
        return out;
    }

    // Labels only (no cloud assembly): the cheapest call for a caller that indexes its own data.
    std::vector<std::uint8_t> segmentLabels(const std::vector<Point3D>& points) {
        std::vector<std::uint8_t> labels(points.size());
        if (points.empty()) return labels;
        ensure(points.size());
        check(rpw_segment(handle_, &points[0].x, points.size(), sizeof(Point3D), labels.data(), nullptr));
        return labels;
    }

    // Post-filter of the standalone app (reference recursive_patchwork.cpp:428-465): non-ground
    // points outside the ego radius whose height is within base_tol of target_height, preceded by
    // a random subsample of at most 2000 ground points.  Host-side; the subsample is unseeded in
    // the reference too, so only sizes are reproducible.
    std::vector<Point3D> sampleGroundAndObstacles(const std::vector<Point3D>& points, float target_height = 1.1f,
                                                  float base_tol = 0.5f) {
        // segmentation, cloud assembly, ego / height-band filter and the 2000-point ground context sample all on
        // the device (RP/src/recursive_patchwork.cpp:428-465); the host only draws the sample's indices
        std::vector<Point3D> result;
        if (points.empty()) return result;
        ensure(points.size());
        std::vector<std::uint8_t> labels(points.size());
        check(rpw_segment(handle_, &points[0].x, points.size(), sizeof(Point3D), labels.data(), nullptr));
        static_assert(sizeof(Point3D) == 3 * sizeof(float), "Point3D must be packed xyz");
        result.resize(points.size() + 2000);
        std::size_t n_ground = 0, n_obstacles = 0;
        check(rpw_sample_ground_and_obstacles(handle_, target_height, base_tol, 2.5f, 2000, 0, &result[0].x, result.size(),
                                              &n_ground, &n_obstacles));
        result.resize(n_ground + n_obstacles);
        return result;
    }

    // ---- utilities (host side, same behaviour as the reference's) ---------------------------
    std::vector<Point3D> cleanPoints(const std::vector<Point3D>& points) {
        std::vector<Point3D> kept;
        kept.reserve(points.size());
        for (const Point3D& p : points)
            if (std::isfinite(p.x) && std::isfinite(p.y) && std::isfinite(p.z)) kept.push_back(p);
        return kept;
    }

    std::vector<Point3D> rotatePoints2D(const std::vector<Point3D>& points, float angle_degrees) {
        // the reference evaluates angle_degrees * M_PI / 180.0f in double and rounds once (recursive_patchwork.cpp:38)
        const float a = static_cast<float>(static_cast<double>(angle_degrees) * M_PI / static_cast<double>(180.0f));
        const float c = std::cos(a), s = std::sin(a);
        std::vector<Point3D> out;
        out.reserve(points.size());
        for (const Point3D& p : points) out.emplace_back(p.x * c - p.y * s, p.x * s + p.y * c, p.z);
        return out;
    }

    // ---- configuration ---------------------------------------------------------------------
    void setConfig(const PatchworkConfig& config) {
        config_ = config;
        if (handle_) { const rpw_config c = config_.toC(); check(rpw_set_config(handle_, &c)); }
    }
    const PatchworkConfig& getConfig() const { return config_; }

    // Pre-size the device buffers (otherwise they grow on demand).
    void reserve(std::size_t max_points) { ensure(max_points); }

private:
    // labels and both clouds come from the device: the result assembly of recursive_patchwork.cpp:402-419 (ground in
    // input order; non-ground in input order, then the beyond-radius points) is a stable compaction kernel; the call
    // hands back views of its pinned staging memory, which become the two vectors (one synchronisation per scan)
    std::pair<std::vector<Point3D>, std::vector<Point3D>> segmentClouds(const std::vector<Point3D>& points, std::uint8_t* labels) {
        std::pair<std::vector<Point3D>, std::vector<Point3D>> out;
        if (points.empty()) return out;  // empty in, two empty clouds out
        ensure(points.size());
        static_assert(sizeof(Point3D) == 3 * sizeof(float), "Point3D must be packed xyz");
        const float *g = nullptr, *ng = nullptr;
        std::size_t n_ground = 0, n_non = 0;
        check(rpw_segment_clouds_view(handle_, &points[0].x, points.size(), sizeof(Point3D), labels, &g, &n_ground, &ng, &n_non));
        const Point3D* gp = reinterpret_cast<const Point3D*>(g);
        const Point3D* np_ = reinterpret_cast<const Point3D*>(ng);
        out.first.assign(gp, gp + n_ground);
        out.second.assign(np_, np_ + n_non);
        return out;
    }

    PatchworkConfig config_;
    int device_ = 0;
    rpw_handle* handle_ = nullptr;
    std::size_t capacity_ = 0;

    void release() {
        if (handle_) rpw_destroy(handle_);
        handle_ = nullptr;
        capacity_ = 0;
    }
    void ensure(std::size_t n) {
        if (handle_ && n <= capacity_) return;
        const std::size_t cap = std::max<std::size_t>(n + n / 4, std::size_t(1) << 18);
        if (handle_) {
            // a larger cloud than any before: the handle grows in place (streams, configuration and
            // solver choice stay; only the capacity-sized buffers are replaced)
            check(rpw_reserve(handle_, cap, 1));
        } else {
            const rpw_config c = config_.toC();
            const int rc = rpw_create(&c, device_, cap, 1, &handle_);
            if (rc != RPW_OK) {
                handle_ = nullptr;
                throw std::runtime_error(std::string("RecursivePatchwork (B200): ") + rpw_last_error(nullptr));
            }
        }
        capacity_ = cap;
    }
    void check(int rc) const {
        if (rc != RPW_OK) throw std::runtime_error(std::string("RecursivePatchwork (B200): ") + rpw_last_error(handle_));
    }
};

// Wall-clock helper with the reference's interface.
class Timer {
public:
    Timer() : start_(clock::now()) {}
    double elapsed() const { return std::chrono::duration<double>(clock::now() - start_).count(); }
    void reset() { start_ = clock::now(); }

private:
    using clock = std::chrono::steady_clock;
    clock::time_point start_;
};

}  // namespace recursive_patchwork
