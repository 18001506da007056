// tests/cpp/test_dropin.cpp — the reference's own smoke tests, re-stated against the drop-in header.
//
// Follows RP/test/test_recursive_patchwork.cpp: testBasicFunctionality (:51-79, 5,000 points,
// R = 50, 8 sectors, max_iter = 50), testEnhancedFiltering (:81-98, 3,000 points, defaults) and
// testPerformance (:146-164, 10,000 points, defaults), with the reference's assertions, plus an
// exact check of the returned clouds against labels written by the caller (argv[1] = cloud file of
// float32 xyz triples, argv[2] = expected label file, optional).  Built and run by
// tests/test_gpu_dropin.py on the GPU box.
#include "recursive_patchwork.hpp"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>
#include <limits>
#include <fstream>
#include <iostream>
#include <random>

using namespace recursive_patchwork;

#define REQUIRE(cond)                                                            \
    do {                                                                         \
        if (!(cond)) { std::fprintf(stderr, "FAILED: %s (%s:%d)\n", #cond, __FILE__, __LINE__); std::exit(1); } \
    } while (0)

static std::vector<Point3D> synthetic_cloud(size_t n, unsigned seed) {
    std::vector<Point3D> pts;
    pts.reserve(n);
    std::mt19937 gen(seed);
    std::normal_distribution<float> ground_z(0.0f, 0.05f);
    std::uniform_real_distribution<float> ground_xy(-50.0f, 50.0f), obstacle_xy(-30.0f, 30.0f), obstacle_z(0.5f, 3.0f);
    const size_t n_ground = static_cast<size_t>(n * 0.7);
    for (size_t i = 0; i < n_ground; ++i) { Point3D p; p.x = ground_xy(gen); p.y = ground_xy(gen); p.z = ground_z(gen); pts.push_back(p); }
    for (size_t i = n_ground; i < n; ++i) { Point3D p; p.x = obstacle_xy(gen); p.y = obstacle_xy(gen); p.z = obstacle_z(gen); pts.push_back(p); }
    return pts;
}

int main(int argc, char** argv) {
    {   // testBasicFunctionality
        auto points = synthetic_cloud(5000, 43);
        PatchworkConfig config;
        config.sensor_height = 1.2f;
        config.filtering_radius = 50.0f;
        config.num_sectors = 8;
        config.max_iter = 50;
        RecursivePatchwork patchwork(config);
        auto [ground_points, non_ground_points] = patchwork.filterGroundPoints(points);
        REQUIRE(ground_points.size() + non_ground_points.size() <= points.size());
        REQUIRE(ground_points.size() > 0);
        REQUIRE(non_ground_points.size() > 0);
        std::printf("basic: ground=%zu non_ground=%zu\n", ground_points.size(), non_ground_points.size());
    }
    {   // testEnhancedFiltering
        auto points = synthetic_cloud(3000, 42);
        RecursivePatchwork patchwork{PatchworkConfig{}};
        auto filtered = patchwork.sampleGroundAndObstacles(points);
        REQUIRE(filtered.size() > 0);
        REQUIRE(filtered.size() <= points.size());
        std::printf("enhanced: filtered=%zu\n", filtered.size());
    }
    {   // testPerformance (prints points/second like the reference; asserts nothing about speed)
        auto points = synthetic_cloud(10000, 42);
        RecursivePatchwork patchwork{PatchworkConfig{}};
        patchwork.filterGroundPoints(points);  // first call creates the handle
        Timer timer;
        auto [ground_points, non_ground_points] = patchwork.filterGroundPoints(points);
        const double dt = timer.elapsed();
        std::printf("performance: %zu points in %.6f s = %.0f points/second, ground=%zu non_ground=%zu\n", points.size(), dt,
                    points.size() / dt, ground_points.size(), non_ground_points.size());
        REQUIRE(ground_points.size() + non_ground_points.size() == points.size());
        // degenerate calls keep the reference's behaviour
        auto empty = patchwork.filterGroundPoints(std::vector<Point3D>{});
        REQUIRE(empty.first.empty() && empty.second.empty());
        std::vector<Point3D> two = {Point3D(3, 4, 0), Point3D(5, 1, 0.1f)};
        auto few = patchwork.filterGroundPoints(two);
        REQUIRE(few.first.empty() && few.second.size() == 2);
        // fewer than three points inside the radius: ({}, cleaned points) in plain input order, beyond-radius points
        // interleaved where they stood (RP/src/recursive_patchwork.cpp:339-341)
        const float inf = std::numeric_limits<float>::infinity();
        std::vector<Point3D> mixed = {Point3D(900, 0, 1), Point3D(3, 4, 0), Point3D(0, -700, 2), Point3D(inf, 0, 0), Point3D(5, 1, 0.1f), Point3D(400, 400, 3)};
        auto deg = patchwork.filterGroundPoints(mixed);
        REQUIRE(deg.first.empty() && deg.second.size() == 5);
        const int order[5] = {0, 1, 2, 4, 5};
        for (int k = 0; k < 5; ++k) REQUIRE(std::memcmp(&deg.second[k], &mixed[order[k]], 12) == 0);
        // setConfig / getConfig
        PatchworkConfig c2 = patchwork.getConfig();
        c2.num_sectors = 16;
        patchwork.setConfig(c2);
        REQUIRE(patchwork.getConfig().num_sectors == 16);
        auto again = patchwork.filterGroundPoints(points);
        REQUIRE(again.first.size() + again.second.size() == points.size());
    }
    {   // the class is copyable and movable like the reference's (its only state there is the configuration); a cloud
        // larger than any before grows the handle in place
        auto small = synthetic_cloud(2000, 5), large = synthetic_cloud(400000, 6);
        PatchworkConfig cfg;
        cfg.num_sectors = 12;
        RecursivePatchwork a(cfg);
        auto ra = a.filterGroundPoints(small);
        RecursivePatchwork b = a;             // copy: same configuration, own handle on first use
        REQUIRE(b.getConfig().num_sectors == 12);
        auto rb = b.filterGroundPoints(small);
        REQUIRE(ra.first.size() == rb.first.size() && ra.second.size() == rb.second.size());
        REQUIRE(std::memcmp(ra.first.data(), rb.first.data(), ra.first.size() * 12) == 0);
        RecursivePatchwork c;
        c = a;                                // copy assignment
        REQUIRE(c.getConfig().num_sectors == 12);
        RecursivePatchwork d = std::move(b);  // move keeps the handle
        auto rd = d.filterGroundPoints(small);
        REQUIRE(rd.first.size() == ra.first.size());
        std::vector<RecursivePatchwork> pool(2, a);  // containers of the class work as with the reference's
        REQUIRE(pool[1].filterGroundPoints(small).first.size() == ra.first.size());
        auto big = a.filterGroundPoints(large);      // 400,000 points > the initial 262,144-point capacity
        REQUIRE(big.first.size() + big.second.size() == large.size());
        auto after = a.filterGroundPoints(small);    // and the grown handle still answers the small cloud identically
        REQUIRE(after.first.size() == ra.first.size() && std::memcmp(after.first.data(), ra.first.data(), ra.first.size() * 12) == 0);
        std::printf("copy/move/grow: ground=%zu, large cloud ground=%zu\n", ra.first.size(), big.first.size());
    }
    {   // fused multi-LiDAR frame: three sensors, default 0 / +120 / -120 degree yaws, ego radius 2.5
        std::vector<std::vector<Point3D>> clouds = {synthetic_cloud(4000, 7), synthetic_cloud(3000, 8), synthetic_cloud(3000, 9)};
        std::vector<LidarConfig> cfgs = {{1, "/lidar_front", 0.0f, 2.5f}, {2, "/lidar_left", 120.0f, 2.5f}, {3, "/lidar_right", -120.0f, 2.5f}};
        RecursivePatchwork patchwork{PatchworkConfig{}};
        std::vector<std::vector<std::uint8_t>> labels;
        auto fused = patchwork.filterGroundPointsFused(clouds, cfgs, &labels);
        size_t ego = 0, total = 0;
        for (auto& l : labels) { total += l.size(); for (auto v : l) ego += v == RPW_LABEL_EGO; }
        REQUIRE(total == 10000 && fused.first.size() > 0 && fused.second.size() > 0);
        REQUIRE(fused.first.size() + fused.second.size() + ego == total);
        std::printf("fused: ground=%zu non_ground=%zu ego_removed=%zu\n", fused.first.size(), fused.second.size(), ego);
    }
    if (argc >= 3) {  // exact clouds from a cloud file + expected labels (written by the python test from the oracle)
        std::ifstream fc(argv[1], std::ios::binary), fl(argv[2], std::ios::binary);
        std::vector<char> cb((std::istreambuf_iterator<char>(fc)), std::istreambuf_iterator<char>());
        std::vector<char> lb((std::istreambuf_iterator<char>(fl)), std::istreambuf_iterator<char>());
        const size_t n = cb.size() / 12;
        REQUIRE(lb.size() == n);
        std::vector<Point3D> points(n);
        std::memcpy(static_cast<void*>(points.data()), cb.data(), n * 12);
        PatchworkConfig cfg;
        cfg.filtering_radius = argc >= 4 ? std::atof(argv[3]) : 150.0f;
        RecursivePatchwork patchwork(cfg);
        std::vector<std::uint8_t> labels;
        auto clouds = patchwork.filterGroundPoints(points, labels);
        size_t agree = 0, g = 0, ng = 0;
        for (size_t i = 0; i < n; ++i) agree += labels[i] == static_cast<std::uint8_t>(lb[i]);
        for (size_t i = 0; i < n; ++i) {
            if (labels[i] == 1) { REQUIRE(std::memcmp(&clouds.first[g], &points[i], 12) == 0); ++g; }
            else if (labels[i] == 0) { REQUIRE(std::memcmp(&clouds.second[ng], &points[i], 12) == 0); ++ng; }
        }
        for (size_t i = 0; i < n; ++i) if (labels[i] == 2) { REQUIRE(std::memcmp(&clouds.second[ng], &points[i], 12) == 0); ++ng; }
        REQUIRE(g == clouds.first.size() && ng == clouds.second.size());
        std::printf("file: n=%zu label_agreement=%.6f ground=%zu non_ground=%zu\n", n, double(agree) / n, g, ng);
        REQUIRE(double(agree) / n >= 0.999);
    }
    std::printf("ALL DROP-IN TESTS PASSED\n");
    return 0;
}
