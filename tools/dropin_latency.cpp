// tools/dropin_latency.cpp — latency of the C++ drop-in as the reference's callers use it: pageable
// std::vector<Point3D> in, the two clouds out (RecursivePatchwork::filterGroundPoints,
// host/recursive_patchwork.hpp), one scan per call like the ROS2 callback
// (RP/src/recursive_patchwork_node.cpp:61-108).  Built by __graft_entry__.build(), run by bench.py.
//   dropin_latency <frames.bin> <n_frames> <filtering_radius> <reps>
// frames.bin: n_frames records of [uint32 n][n * 3 float32 xyz].  Prints one JSON line.
#include "recursive_patchwork.hpp"

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <vector>

using namespace recursive_patchwork;

int main(int argc, char** argv) {
    if (argc < 5) { std::fprintf(stderr, "usage: %s frames.bin n_frames radius reps\n", argv[0]); return 2; }
    const int n_frames = std::atoi(argv[2]), reps = std::atoi(argv[4]);
    std::ifstream f(argv[1], std::ios::binary);
    std::vector<std::vector<Point3D>> frames(n_frames);
    for (auto& fr : frames) {
        std::uint32_t n = 0;
        f.read(reinterpret_cast<char*>(&n), 4);
        fr.resize(n);
        f.read(reinterpret_cast<char*>(fr.data()), std::streamsize(n) * 12);
    }
    if (!f) { std::fprintf(stderr, "short read\n"); return 2; }
    PatchworkConfig cfg;
    cfg.filtering_radius = float(std::atof(argv[3]));
    try {
        RecursivePatchwork rp(cfg);
        std::size_t ground = 0;
        for (int r = 0; r < 16; ++r) ground = rp.filterGroundPoints(frames[r % n_frames]).first.size();  // warm-up: handle, graph
        std::vector<double> clouds_ms, labels_ms;
        for (int r = 0; r < reps; ++r) {
            const auto& fr = frames[r % n_frames];
            auto t0 = std::chrono::steady_clock::now();
            auto res = rp.filterGroundPoints(fr);
            auto t1 = std::chrono::steady_clock::now();
            auto lab = rp.segmentLabels(fr);
            auto t2 = std::chrono::steady_clock::now();
            ground += res.first.size() + lab.size();
            clouds_ms.push_back(std::chrono::duration<double, std::milli>(t1 - t0).count());
            labels_ms.push_back(std::chrono::duration<double, std::milli>(t2 - t1).count());
        }
        auto q = [](std::vector<double> v, double p) { std::sort(v.begin(), v.end()); return v[std::min(v.size() - 1, std::size_t(p * v.size()))]; };
        std::printf("{\"clouds_p50_ms\": %.4f, \"clouds_p99_ms\": %.4f, \"labels_p50_ms\": %.4f, \"labels_p99_ms\": %.4f, \"reps\": %d, \"checksum\": %zu}\n",
                    q(clouds_ms, 0.5), q(clouds_ms, 0.99), q(labels_ms, 0.5), q(labels_ms, 0.99), reps, ground);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
