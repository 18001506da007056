// csrc/rpw_kernels.h — interface between the kernels (rpw_kernels.cu) and the C-ABI host code
// (rpw_capi.cu).  Internal; the public boundary is include/rpw_b200.h.
#pragma once

#include "rpw_b200.h"
#include "rpw_device.cuh"

// (Experiment, measured and not adopted, code removed: K2 scattering input INDICES -- 4 B per point instead of reading and
// writing the 16-byte record -- with the level-0 fit gathering its points from the input through the index list.  Scatter
// 0.349 -> 0.205 ms per 512 scans, but the fit 1.204 -> 1.391 ms: a patch's points lie in runs of a few records per azimuth
// step, so every 16-byte lane load of the gather is its own sector request (32 per warp instruction instead of 4), on top
// of a second dependent memory round trip per patch.  Step 1.846 -> 1.893 ms.  Labels identical.)

namespace rpw {

typedef rpw_node rpw_node_rec;

// One fitPlaneAndSplit node on the device worklist: a contiguous run of its level's point buffer.
struct __align__(16) NodeRef {
    uint32_t start;  // first slot (global index into the sorted / partition buffers)
    uint32_t n;
    uint32_t root;   // scan * P + patch
    uint32_t pad;
};

struct FitArgs {
    const float4* sortedA;     // level 0: (x, y, z, bits of global input index), patch-major, input order inside a patch
    float4* bufB;              // odd levels: partitioned children (x, y, z, -)
    float4* bufC;              // even levels >= 2
    uint8_t* gmask;            // per-slot mask scratch for nodes that do not fit in shared memory
    uint8_t* labels;           // per input point
    const uint32_t* patch_start;  // [batch][P + 1]
    const uint32_t* cls_count;    // [kNumFitClasses] non-empty root patches per size class (this launch group)
    const uint4* cls_list;        // [kNumFitClasses][cls_cap] (start, n, scan * P + patch, -) per class, in arrival order
    uint32_t* host_counts;        // mapped host memory: the class counts, [kClsWords-1] the group's scan count (grid estimate
                                  // of the next launch group; never needed for correctness)
    uint32_t cls_cap;
    float* root_mean;          // [batch * P] mean range of the root patch (Q4)
    NodeRef* queue[2];         // next-level queues, ping-pong by level parity
    uint32_t* q_count;         // [levels_cap] nodes enqueued for level l
    uint32_t* fetch_ctr;       // [levels_cap] dynamic fetch cursor of level l
    uint32_t* stats;           // [0] levels run (out) [1] nodes processed (accumulator) [2] block arrival [3] nodes (out)
    uint32_t* overflow;        // set if a queue would overflow
    unsigned long long* timing; // optional [16] cycle accounting (rpw_debug_fit_timing)
    rpw_trace_rec* trace;      // optional per-node timeline (rpw_debug_fit_trace)
    uint32_t* trace_count;
    uint32_t trace_cap;
    rpw_node_rec* dbg_nodes;   // optional
    uint32_t* dbg_count;
    uint32_t dbg_cap;
    uint32_t q_cap;
    int n_roots;               // batch * P
    int n_scans;               // scans in this launch group
    int P;
    int smem_cap;              // points a block can hold in shared memory
    int profile;               // class table of the level-0 fit: 0 throughput, 1 latency (thread-block clusters for big patches)
    uint32_t scan_base;        // index of the launch group's first scan inside the call's batch (debug records)
    FitParams fp;
};

constexpr int kNumFitClasses = 7;  // size classes of the level-0 fit kernel (rpw_kernels.cu: kFitClasses)
constexpr int kClsWords = 16;      // words of a class-count array: counts [0, kNumFitClasses), scan count in the last
struct ClassBounds { uint32_t hi[kNumFitClasses]; };  // class c holds patches with hi[c-1] < n <= hi[c]
ClassBounds fit_class_bounds(int profile);  // 0: throughput table (batches), 1: latency table (one or two scans per call; rpw_fit.cuh)
size_t fit_smem_bytes(int smem_cap, int threads);
cudaError_t fit_configure(int smem_cap, int* blocks_per_sm);

cudaError_t launch_bin(cudaStream_t st, const PointLayout& lay, const float* pts, const uint64_t* scan_off, const uint32_t* chunk_base,
                       const ZoneModel& zm, uint16_t* keys, uint8_t* labels, uint32_t* blk_hist, uint32_t* cls_count,
                       const FusionTable* fusion, int max_chunks, int batch, int threads);
cudaError_t launch_offsets(cudaStream_t st, const uint64_t* scan_off, const uint32_t* chunk_base, uint32_t* blk_hist,
                           uint32_t* patch_start, uint32_t* cls_count, uint4* cls_list, uint32_t cls_cap, int P, int batch, int profile);
cudaError_t launch_scatter(cudaStream_t st, const PointLayout& lay, const float* pts, const uint64_t* scan_off, const uint32_t* chunk_base,
                           const uint16_t* keys, const uint32_t* blk_hist, const uint32_t* patch_start, float4* sorted,
                           int P, const FusionTable* fusion, int max_chunks, int batch, int threads);
cudaError_t launch_compact(cudaStream_t st, const PointLayout& lay, const float* pts, const uint8_t* labels, const uint64_t* scan_off,
                           const uint32_t* chunk_base, uint32_t* cnt, const FusionTable* fusion, float* ground, float* nonground,
                           uint32_t* scan_counts, int max_chunks, int batch, int packed = 0, int threads = kBinThreads);
cudaError_t launch_bev(cudaStream_t st, int mode, const float* a, uint32_t n_a, const float* b, uint32_t n_b, int width, int height,
                       float x_min, float y_min, float x_scale, float y_scale, uint32_t* owner, uint8_t* bgr);
cudaError_t launch_obstacles(cudaStream_t st, const float* nonground_xyz, uint32_t n, float target, float tol, float ego, uint32_t* cnt,
                             float* out, uint32_t* total);
cudaError_t launch_gather_xyz(cudaStream_t st, const float* xyz, const uint32_t* idx, uint32_t k, float* out);
cudaError_t launch_fit_roots(cudaStream_t st, const FitArgs& args, int size_class, unsigned grid_blocks);
cudaError_t launch_fit_levels(cudaStream_t st, const FitArgs& args, int grid_blocks);
cudaError_t launch_eig3(cudaStream_t st, const float* mats, size_t count, float* evals, float* evecs);
cudaError_t launch_normal(cudaStream_t st, const float* sc, size_t count, int mode, float* normals, uint32_t* cycles);
cudaError_t launch_atan2(cudaStream_t st, const float* y, const float* x, size_t count, float* out);

}  // namespace rpw
