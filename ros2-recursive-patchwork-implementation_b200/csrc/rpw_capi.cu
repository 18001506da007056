// csrc/rpw_capi.cu — host side of the C-ABI declared in include/rpw_b200.h: handle, buffers,
// stream plumbing, host<->device copies and the launch sequence K1 -> K1b -> K2 -> K3.
// There is no CPU fallback in this file or anywhere in the library: without a CUDA device
// rpw_create fails, and every entry point reports CUDA errors instead of hiding them.
#include "rpw_kernels.h"

#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost nothing unless a profiler is attached

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <random>
#include <string>
#include <vector>

using namespace rpw;

static thread_local std::string g_create_error;

struct rpw_handle {
    rpw_config cfg;
    ZoneModel zm;
    FitParams fp;
    int device = 0;
    int num_sms = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // Launch groups of one call alternate over two lanes so that the long-tail patches of one group
    // overlap the bulk of the next.  A lane has its own streams and its own small worklist state; the
    // big per-point buffers are shared (groups touch disjoint index ranges).
    struct Lane {
        cudaStream_t main = nullptr;                          // K1, K1b, K2, K3b
        cudaStream_t side[kNumFitClasses] = {};   // K3a size classes run concurrently
        cudaEvent_t ev_fork = nullptr, ev_join[kNumFitClasses] = {}, ev_start = nullptr, ev_done = nullptr;
        NodeRef* d_queue[2] = {nullptr, nullptr};
        uint32_t* d_counters = nullptr;     // fetch_ctr[levels_cap] | q_count[levels_cap] | stats[8] | overflow
        uint32_t* d_cls_count = nullptr;    // [8] root patches per size class over the launch group
        uint4* d_cls_list = nullptr;        // [kNumFitClasses][cls_cap] the classes' work lists
        uint32_t* h_counts = nullptr;       // mapped host copy of the last finished group's class counts ([kClsWords-1]: its scan count)
        uint32_t* d_counts_map = nullptr;   // device address of h_counts
    };
    static constexpr int kLanes = 2;
    Lane lane[kLanes];
    cudaEvent_t ev_call = nullptr;
    uint32_t* d_dbg_count = nullptr;
    int n_waves = 1;  // launch groups per call (1: within one call more groups only add tails)
    size_t cap_points = 0, cap_batch = 0;
    int P = 0;
    uint32_t cls_cap = 0;
    int levels_cap = 0;
    uint32_t q_cap = 0;
    int smem_cap = 0;
    int fit_blocks = 0;
    int solver = RPW_SOLVER_HYBRID;
    int exact_replay = -1;  // rpw_set_exact_replay
    int fit_profile = -1;   // class table: -1 by batch size, 0 throughput, 1 latency (RPW_FIT_PROFILE)

    // device buffers
    float* d_in = nullptr;       // staged input records of host-path calls
    size_t d_in_bytes = 0;
    size_t h_stage_in_bytes = 0;
    size_t field_off[3] = {0, 4, 8};  // byte offsets of x, y, z inside a record for the next host-path call
    size_t src_pitch = 0, src_off = 0;  // PointCloud2 with adjacent x, y, z: only those 12 bytes of every src_pitch-byte record are copied
    int pc2_pack = 0;                   // RPW_PC2_PACK=1 switches the strided copy on: it LOST (0.273 against 0.178 ms per 120 k-point scan of 32-byte records: the copy engine spends ~1.6 ns per 12-byte row)
    uint16_t* d_keys = nullptr;
    uint8_t* d_labels = nullptr;
    float4* d_sortedA = nullptr;       // the patches: (x, y, z, bits of the input index), patch-major, input order inside a patch
    float4* d_bufB = nullptr;
    float4* d_bufC = nullptr;
    uint8_t* d_gmask = nullptr;
    uint32_t* d_blk_hist = nullptr;
    uint32_t* d_patch_start = nullptr;
    float* d_root_mean = nullptr;
    uint64_t* d_scan_off = nullptr;
    uint32_t* d_chunk_base = nullptr;
    FusionTable* d_fusion = nullptr;    // device copy of the sensor table of a fused frame
    const FusionTable* fusion_arg = nullptr;  // non-null while a fused frame is being enqueued
    FusionTable* h_fusion = nullptr;    // pinned
    rpw_node* d_dbg_nodes = nullptr;
    unsigned long long* d_timing = nullptr;  // [16], allocated by rpw_debug_fit_timing
    rpw_trace_rec* d_trace = nullptr;        // rpw_debug_fit_trace
    // result assembly (rpw_last_clouds): what the last call ran on, and lazily allocated buffers
    const float* last_pts = nullptr;
    PointLayout last_lay{};
    uint8_t* last_labels = nullptr;
    bool last_fused = false;
    uint32_t* d_cmp_cnt = nullptr;      // [chunk rows][4] label counts per 4096-point chunk
    uint32_t* d_scan_counts = nullptr;  // [cap_batch][2] ground / non-ground points per scan
    uint32_t* h_scan_counts = nullptr;  // pinned copy
    float* d_cloud_g = nullptr;         // 3 floats x cap_points each, only for host-bound clouds
    float* d_cloud_ng = nullptr;
    float* h_cloud_stage = nullptr;     // pinned, 3 floats x cap_points: both clouds of a single scan, ground first (rpw_segment_clouds_view)
    uint32_t* d_sample_idx = nullptr;   // rpw_sample_ground_and_obstacles: indices of the ground context sample
    uint32_t* d_bev_owner = nullptr;    // rpw_bev_image: winning draw index per pixel, then the BGR image behind it
    size_t bev_pixels_cap = 0;
    size_t sample_idx_cap = 0;
    uint32_t* d_trace_count = nullptr;
    uint32_t trace_cap = 0;
    bool timing_enabled = false;
    uint32_t dbg_cap = 0;
    bool dbg_enabled = false;

    // The pipeline of a single host-path scan as a CUDA graph (the reference's real use: one scan per ROS2 callback):
    // kernels with fixed parameters and grids sized for `chunks` 4096-point chunks; a call only copies its points and
    // its 24-byte meta block in, launches the graph and copies the labels out.  Re-captured when the configuration,
    // the solver, the record layout, the stream or the capacity changes (cfg_epoch) or a scan needs more chunks.
    struct ScanGraph {
        cudaGraphExec_t exec = nullptr;
        int chunks = 0;
        uint64_t epoch = 0;
        PointLayout lay{};
        cudaStream_t stream = nullptr;
    } graph1;
    uint64_t cfg_epoch = 1;
    bool graphs_enabled = true;
    uint64_t graph_launches = 0, graph_captures = 0;

    // host staging
    static constexpr int kMetaSlots = 8;
    cudaEvent_t meta_ev[kMetaSlots] = {};  // the upload from ring slot s has been consumed
    size_t meta_slot_bytes = 0;
    int meta_next = 0;
    uint64_t* h_meta = nullptr;  // pinned ring of kMetaSlots blocks: scan_off[batch+1] (u64) then chunk_base[batch+1] (u32)
    void* h_stage_in = nullptr;  // pinned, lazily allocated, cap_points*16
    uint8_t* h_stage_labels = nullptr;
    uint32_t* h_stats = nullptr;  // pinned 8 words

    // last call
    std::vector<uint64_t> last_off;
    size_t last_batch = 0;
    size_t last_total = 0;
    std::vector<const float*> pend_src;
    std::vector<uint8_t*> pend_labels;
    bool pend_labels_staged = false;
    uint64_t launches = 0;
    uint64_t launches_call = 0;

    // optional per-kernel timing (rpw_profile_*): CUDA events bracketing every launch
    bool prof_enabled = false;
    std::vector<cudaEvent_t> prof_ev;   // pairs
    std::vector<int> prof_kind;         // kernel id per pair
    size_t prof_used = 0;
    double prof_ms[RPW_PROF_KERNELS] = {0, 0, 0, 0};
    uint64_t prof_launches[RPW_PROF_KERNELS] = {0, 0, 0, 0};
    std::string err;
};

#define RPW_FAIL(h, code, ...)                                   \
    do {                                                         \
        char _b[512];                                            \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                   \
        if (h) (h)->err = _b; else g_create_error = _b;          \
        return code;                                             \
    } while (0)

#define RPW_CUDA(h, call)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (call);                                                                  \
        if (_e != cudaSuccess) RPW_FAIL(h, RPW_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

extern "C" {

int rpw_abi_version(void) { return RPW_ABI_VERSION; }

void rpw_default_config(rpw_config* c) {
    if (!c) return;
    // defaults of PatchworkConfig, RP/include/recursive_patchwork.hpp:25-36
    c->sensor_height = 1.2f;
    c->max_range = 150.0f;
    c->num_sectors = 10;
    c->max_iter = 100;
    c->adaptive_seed_height = 1;
    c->th_seeds = 0.15f;
    c->th_dist = 0.2f;
    c->th_outlier = 0.08f;
    c->filtering_radius = 150.0f;
    c->max_split_depth = 1000;
}

int rpw_zone_model(const rpw_config* cfg, float* ring_edges9, float* sector_angle) {
    if (!cfg || !ring_edges9 || !sector_angle) return RPW_ERR_BAD_ARG;
    // RP/src/recursive_patchwork.cpp:344-352: powf on floats, sector angle in double rounded once.
    const float r_min = 1.0f, r_max = cfg->filtering_radius;
    for (int i = 0; i <= RPW_NUM_RINGS; ++i) ring_edges9[i] = r_min * powf(r_max / r_min, (float)i / RPW_NUM_RINGS);
    *sector_angle = (float)((double)2.0f * 3.14159265358979323846 / (double)cfg->num_sectors);
    return RPW_OK;
}

static int check_config(const rpw_config* c, std::string& why) {
    if (c->num_sectors < 1 || c->num_sectors > kMaxSectors) { why = "num_sectors must be in [1, 128]"; return RPW_ERR_BAD_ARG; }
    if (!(c->filtering_radius == c->filtering_radius)) { why = "filtering_radius is NaN"; return RPW_ERR_BAD_ARG; }
    return RPW_OK;
}

static void apply_solver(rpw_handle* h) {
    h->cfg_epoch++;
    h->fp.exact_eig = (h->solver == RPW_SOLVER_EIGEN_QR || h->solver == RPW_SOLVER_REFERENCE) ? 1 : 0;
    h->fp.hybrid = h->solver == RPW_SOLVER_HYBRID ? 1 : 0;
    h->fp.exact_replay = h->solver == RPW_SOLVER_REFERENCE ? 0 : h->exact_replay;
}

static void apply_config(rpw_handle* h, const rpw_config* c) {
    h->cfg = *c;
    rpw_zone_model(c, h->zm.ring_edges, &h->zm.sector_angle);
    h->zm.inv_sector_angle = 1.0f / h->zm.sector_angle;
    h->zm.rings_increasing = 1;
    for (int i = 0; i < RPW_NUM_RINGS; ++i) if (!(h->zm.ring_edges[i] < h->zm.ring_edges[i + 1])) h->zm.rings_increasing = 0;
    h->zm.radius = c->filtering_radius;
    h->zm.num_sectors = c->num_sectors;
    h->zm.num_patches = RPW_NUM_RINGS * c->num_sectors;
    h->fp.sensor_height = c->sensor_height;
    h->fp.th_seeds = c->th_seeds;
    h->fp.th_dist = c->th_dist;
    h->fp.radius = c->filtering_radius;
    h->fp.max_iter = c->max_iter;
    h->fp.adaptive_seed_height = c->adaptive_seed_height;
    h->fp.max_split_depth = c->max_split_depth;
    apply_solver(h);
}

static void free_patch_buffers(rpw_handle* h) {
    cudaFree(h->d_blk_hist); h->d_blk_hist = nullptr;
    cudaFree(h->d_patch_start); h->d_patch_start = nullptr;
    cudaFree(h->d_root_mean); h->d_root_mean = nullptr;
    for (auto& L : h->lane) {
        cudaFree(L.d_cls_count); L.d_cls_count = nullptr;
        cudaFree(L.d_cls_list); L.d_cls_list = nullptr;
    }
}

static int alloc_patch_buffers(rpw_handle* h) {
    const int P = RPW_NUM_RINGS * h->cfg.num_sectors;
    h->P = P;
    const size_t rows = h->cap_points / kBinChunk + h->cap_batch + 1;
    RPW_CUDA(h, cudaMalloc(&h->d_blk_hist, rows * P * sizeof(uint32_t)));
    RPW_CUDA(h, cudaMalloc(&h->d_patch_start, h->cap_batch * (size_t)(P + 1) * sizeof(uint32_t)));
    RPW_CUDA(h, cudaMalloc(&h->d_root_mean, h->cap_batch * (size_t)P * sizeof(float)));
    const size_t pairs = h->cap_batch * (size_t)P;  // a listed patch holds at least one point
    h->cls_cap = (uint32_t)(pairs < h->cap_points ? pairs : h->cap_points);
    for (auto& L : h->lane) {
        RPW_CUDA(h, cudaMalloc(&L.d_cls_count, kClsWords * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMemset(L.d_cls_count, 0, kClsWords * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMalloc(&L.d_cls_list, (size_t)kNumFitClasses * h->cls_cap * sizeof(uint4)));
        if (!L.h_counts) {
            RPW_CUDA(h, cudaHostAlloc(&L.h_counts, kClsWords * sizeof(uint32_t), cudaHostAllocMapped));
            RPW_CUDA(h, cudaHostGetDevicePointer(&L.d_counts_map, L.h_counts, 0));
        }
        memset(L.h_counts, 0, kClsWords * sizeof(uint32_t));
    }
    return RPW_OK;
}

static int alloc_level_buffers(rpw_handle* h) {
    long long lv = (long long)h->cap_points / 10 + 1;
    if (h->cfg.max_split_depth >= 0 && h->cfg.max_split_depth < lv) lv = h->cfg.max_split_depth;
    if (lv < 0) lv = 0;
    h->levels_cap = (int)lv + 4;
    const size_t words = (size_t)h->levels_cap * 2 + 16;
    for (auto& L : h->lane) {
        cudaFree(L.d_counters);
        L.d_counters = nullptr;
        RPW_CUDA(h, cudaMalloc(&L.d_counters, words * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMemset(L.d_counters, 0, words * sizeof(uint32_t)));
    }
    return RPW_OK;
}

}  // extern "C"

// Everything whose size follows the handle's capacity (max_total_points, max_batch): the per-point
// buffers, the worklists, the per-scan meta blocks.  rpw_create and rpw_reserve call it.
static void free_capacity_buffers(rpw_handle* h) {
    cudaFree(h->d_in); h->d_in = nullptr; h->d_in_bytes = 0;
    cudaFree(h->d_keys); h->d_keys = nullptr;
    cudaFree(h->d_labels); h->d_labels = nullptr;
    cudaFree(h->d_sortedA); h->d_sortedA = nullptr;
    cudaFree(h->d_bufB); h->d_bufB = nullptr;
    cudaFree(h->d_bufC); h->d_bufC = nullptr;
    cudaFree(h->d_gmask); h->d_gmask = nullptr;
    for (auto& L : h->lane) {
        cudaFree(L.d_queue[0]); cudaFree(L.d_queue[1]); L.d_queue[0] = L.d_queue[1] = nullptr;
        cudaFree(L.d_counters); L.d_counters = nullptr;
    }
    cudaFree(h->d_scan_off); h->d_scan_off = nullptr;
    h->d_chunk_base = nullptr;  // (inside the d_scan_off block)
    if (h->graph1.exec) { cudaGraphExecDestroy(h->graph1.exec); h->graph1.exec = nullptr; }
    if (h->h_meta) { cudaFreeHost(h->h_meta); h->h_meta = nullptr; }
    free_patch_buffers(h);
    // lazily allocated, sized by the capacity: dropped here, re-created on demand
    cudaFree(h->d_cmp_cnt); h->d_cmp_cnt = nullptr;
    cudaFree(h->d_scan_counts); h->d_scan_counts = nullptr;
    if (h->h_scan_counts) { cudaFreeHost(h->h_scan_counts); h->h_scan_counts = nullptr; }
    cudaFree(h->d_cloud_g); h->d_cloud_g = nullptr;
    cudaFree(h->d_cloud_ng); h->d_cloud_ng = nullptr;
    if (h->h_cloud_stage) { cudaFreeHost(h->h_cloud_stage); h->h_cloud_stage = nullptr; }
    if (h->h_stage_in) { cudaFreeHost(h->h_stage_in); h->h_stage_in = nullptr; h->h_stage_in_bytes = 0; }
    if (h->h_stage_labels) { cudaFreeHost(h->h_stage_labels); h->h_stage_labels = nullptr; }
    cudaFree(h->d_dbg_nodes); h->d_dbg_nodes = nullptr; h->dbg_cap = 0;
}

static int alloc_capacity_buffers(rpw_handle* h) {
    const size_t N = h->cap_points, B = h->cap_batch;
#define RPW_ALLOC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) RPW_FAIL(h, _e == cudaErrorMemoryAllocation ? RPW_ERR_ALLOC : RPW_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(_e)); } while (0)
    RPW_ALLOC(cudaMalloc(&h->d_in, N * 16));
    h->d_in_bytes = N * 16;
    RPW_ALLOC(cudaMalloc(&h->d_keys, N * sizeof(uint16_t)));
    RPW_ALLOC(cudaMalloc(&h->d_labels, N));
    RPW_ALLOC(cudaMalloc(&h->d_sortedA, N * sizeof(float4)));
    RPW_ALLOC(cudaMalloc(&h->d_bufB, N * sizeof(float4)));
    RPW_ALLOC(cudaMalloc(&h->d_bufC, N * sizeof(float4)));
    RPW_ALLOC(cudaMalloc(&h->d_gmask, N));
    h->q_cap = (uint32_t)(N / 25 + 64);
    for (auto& L : h->lane) {
        RPW_ALLOC(cudaMalloc(&L.d_queue[0], (size_t)h->q_cap * sizeof(NodeRef)));
        RPW_ALLOC(cudaMalloc(&L.d_queue[1], (size_t)h->q_cap * sizeof(NodeRef)));
    }
    // one device block: scan_off[batch + 1] (u64) directly followed by chunk_base[batch + 1] (u32) of the CURRENT call,
    // so that a call uploads its meta data with a single copy
    h->meta_slot_bytes = ((B + 1) * (sizeof(uint64_t) + sizeof(uint32_t)) + 63) / 64 * 64;
    RPW_ALLOC(cudaMalloc(&h->d_scan_off, h->meta_slot_bytes));
    h->d_chunk_base = reinterpret_cast<uint32_t*>(h->d_scan_off + B + 1);
    RPW_ALLOC(cudaMallocHost(&h->h_meta, h->meta_slot_bytes * rpw_handle::kMetaSlots));
    h->meta_next = 0;
    h->cfg_epoch++;
#undef RPW_ALLOC
    int rc = alloc_patch_buffers(h);
    if (rc != RPW_OK) return rc;
    rc = alloc_level_buffers(h);
    if (rc != RPW_OK) return rc;
    h->last_off.clear(); h->last_batch = 0; h->last_total = 0;
    h->last_pts = nullptr; h->last_labels = nullptr;
    h->pend_labels.clear();
    return RPW_OK;
}

extern "C" {

void rpw_destroy(rpw_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    free_capacity_buffers(h);
    cudaFree(h->d_dbg_count); cudaFree(h->d_trace); cudaFree(h->d_trace_count); cudaFree(h->d_sample_idx); cudaFree(h->d_bev_owner);
    cudaFree(h->d_timing); cudaFree(h->d_fusion);
    if (h->h_fusion) cudaFreeHost(h->h_fusion);
    if (h->h_stats) cudaFreeHost(h->h_stats);
    for (auto& L : h->lane) if (L.h_counts) { cudaFreeHost(L.h_counts); L.h_counts = nullptr; }
    for (cudaEvent_t e : h->prof_ev) cudaEventDestroy(e);
    for (auto& L : h->lane) {
        for (int k = 0; k < kNumFitClasses; ++k) { if (L.side[k]) cudaStreamDestroy(L.side[k]); if (L.ev_join[k]) cudaEventDestroy(L.ev_join[k]); }
        if (L.ev_fork) cudaEventDestroy(L.ev_fork);
        if (L.ev_start) cudaEventDestroy(L.ev_start);
        if (L.ev_done) cudaEventDestroy(L.ev_done);
        if (L.main) cudaStreamDestroy(L.main);
    }
    if (h->ev_call) cudaEventDestroy(h->ev_call);
    for (auto& e : h->meta_ev) if (e) cudaEventDestroy(e);
    if (h->own_stream) cudaStreamDestroy(h->own_stream);
    delete h;
}

int rpw_create(const rpw_config* cfg, int device, size_t max_total_points, size_t max_batch, rpw_handle** out) {
    if (!out) { g_create_error = "out is NULL"; return RPW_ERR_BAD_ARG; }
    *out = nullptr;
    rpw_config c;
    if (cfg) c = *cfg; else rpw_default_config(&c);
    std::string why;
    if (check_config(&c, why) != RPW_OK) { g_create_error = why; return RPW_ERR_BAD_ARG; }
    if (max_total_points == 0 || max_batch == 0) { g_create_error = "capacity must be non-zero"; return RPW_ERR_BAD_ARG; }
    if (max_total_points >= 0xFFFF0000ull) { g_create_error = "max_total_points must be below 2^32 - 65536"; return RPW_ERR_BAD_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (this library has no CPU fallback)";
        return RPW_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= ndev) { g_create_error = "device index out of range"; return RPW_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10) {
        g_create_error = "device is not sm_100 (this library carries sm_100a code only)";
        return RPW_ERR_NO_DEVICE;
    }
    if (!prop.cooperativeLaunch) { g_create_error = "device lacks cooperative launch"; return RPW_ERR_NO_DEVICE; }
    rpw_handle* h = new (std::nothrow) rpw_handle();
    if (!h) { g_create_error = "out of host memory"; return RPW_ERR_ALLOC; }
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    h->cap_points = max_total_points;
    h->cap_batch = max_batch;
    if (const char* s = getenv("RPW_PLANE_SOLVER")) {
        const int v = atoi(s);
        h->solver = v == RPW_SOLVER_CLOSED_FORM || v == RPW_SOLVER_EIGEN_QR || v == RPW_SOLVER_REFERENCE ? v : RPW_SOLVER_HYBRID;
    }
    if (const char* s = getenv("RPW_EXACT_REPLAY")) h->exact_replay = atoi(s) < 0 ? -1 : atoi(s);
    apply_config(h, &c);
    int rc = RPW_OK;
    auto fail = [&](int code) { g_create_error = h->err; rpw_destroy(h); return code; };
#define TRY(x) do { rc = (x); if (rc != RPW_OK) return fail(rc); } while (0)
#define TRYC(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(_e); return fail(_e == cudaErrorMemoryAllocation ? RPW_ERR_ALLOC : RPW_ERR_CUDA); } } while (0)
    TRYC(cudaSetDevice(device));
    TRYC(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    // The size classes of the level-0 fit run on side streams whose priority falls with the patch
    // size (side[0] is the largest class): the block scheduler then starts the big, long-iterating
    // patches first and back-fills the shared memory they leave with the small ones, instead of
    // draining the kernels in whatever order their stream dependencies happen to resolve.
    int prio_least = 0, prio_greatest = 0;
    TRYC(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    const int prio_levels = prio_least - prio_greatest + 1;
    for (auto& L : h->lane) {
        TRYC(cudaStreamCreateWithFlags(&L.main, cudaStreamNonBlocking));
        for (int k = 0; k < kNumFitClasses; ++k) {
            int lev = k * prio_levels / kNumFitClasses;
            if (lev > prio_levels - 1) lev = prio_levels - 1;
            TRYC(cudaStreamCreateWithPriority(&L.side[k], cudaStreamNonBlocking, prio_greatest + lev));
            TRYC(cudaEventCreateWithFlags(&L.ev_join[k], cudaEventDisableTiming));
        }
        TRYC(cudaEventCreateWithFlags(&L.ev_fork, cudaEventDisableTiming));
        TRYC(cudaEventCreateWithFlags(&L.ev_start, cudaEventDisableTiming));
        TRYC(cudaEventCreateWithFlags(&L.ev_done, cudaEventDisableTiming));
    }
    TRYC(cudaEventCreateWithFlags(&h->ev_call, cudaEventDisableTiming));
    for (auto& e : h->meta_ev) TRYC(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char* s = getenv("RPW_NO_GRAPH")) h->graphs_enabled = atoi(s) == 0;
    if (const char* s = getenv("RPW_FIT_PROFILE")) h->fit_profile = atoi(s) < 0 ? -1 : (atoi(s) ? 1 : 0);
    if (const char* s = getenv("RPW_PC2_PACK")) h->pc2_pack = atoi(s) != 0;
    TRYC(cudaMalloc(&h->d_dbg_count, sizeof(uint32_t)));
    TRYC(cudaMemset(h->d_dbg_count, 0, sizeof(uint32_t)));
    TRY(alloc_capacity_buffers(h));
    TRYC(cudaMalloc(&h->d_fusion, sizeof(FusionTable)));
    TRYC(cudaMallocHost(&h->h_fusion, sizeof(FusionTable)));
    TRYC(cudaMallocHost(&h->h_stats, 16 * rpw_handle::kLanes * sizeof(uint32_t)));
    // shared-memory budget of the fit kernel: points a block keeps resident
    // (the levels kernel: kLevelBlocksPerSm persistent blocks per SM share the shared memory; nodes of depth >= 1
    // are mostly small, larger ones stream from L2)
    int cap = ((220 * 1024 / kLevelBlocksPerSm) - 4096) / 13 / 256 * 256;
    if (cap > 8192) cap = 8192;
    if (const char* s = getenv("RPW_FIT_SMEM_CAP")) { const int v = atoi(s); if (v >= 256 && v <= 16384) cap = v; }
    if (const char* s = getenv("RPW_WAVES")) { const int v = atoi(s); if (v >= 0) h->n_waves = v; }
    h->smem_cap = cap;
    int bps = 0;
    TRYC(fit_configure(cap, &bps));
    if (bps < 1) { h->err = "fit kernel does not fit on an SM"; return fail(RPW_ERR_CUDA); }
    if (bps > kLevelBlocksPerSm) bps = kLevelBlocksPerSm;
    h->fit_blocks = h->num_sms * bps;  // the levels kernel: persistent blocks, all resident (cooperative launch)
#undef TRY
#undef TRYC
    *out = h;
    return RPW_OK;
}

int rpw_set_config(rpw_handle* h, const rpw_config* cfg) {
    if (!h || !cfg) return RPW_ERR_BAD_ARG;
    std::string why;
    if (check_config(cfg, why) != RPW_OK) RPW_FAIL(h, RPW_ERR_BAD_ARG, "%s", why.c_str());
    RPW_CUDA(h, cudaSetDevice(h->device));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    const bool sectors_changed = cfg->num_sectors != h->cfg.num_sectors;
    const bool depth_changed = cfg->max_split_depth != h->cfg.max_split_depth;
    apply_config(h, cfg);
    if (sectors_changed) {
        free_patch_buffers(h);
        int rc = alloc_patch_buffers(h);
        if (rc != RPW_OK) return rc;
    }
    if (depth_changed) {
        int rc = alloc_level_buffers(h);
        if (rc != RPW_OK) return rc;
    }
    return RPW_OK;
}

int rpw_reserve(rpw_handle* h, size_t max_total_points, size_t max_batch) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (max_total_points <= h->cap_points && max_batch <= h->cap_batch) return RPW_OK;
    if (max_total_points >= 0xFFFF0000ull) RPW_FAIL(h, RPW_ERR_BAD_ARG, "max_total_points must be below 2^32 - 65536");
    RPW_CUDA(h, cudaSetDevice(h->device));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    // the handle, its streams, events, configuration and debug switches stay; only the buffers whose
    // size follows the capacity are replaced
    free_capacity_buffers(h);
    if (max_total_points > h->cap_points) h->cap_points = max_total_points;
    if (max_batch > h->cap_batch) h->cap_batch = max_batch;
    const int rc = alloc_capacity_buffers(h);
    if (rc != RPW_OK) return rc;
    return h->dbg_enabled ? rpw_debug_enable_nodes(h, 1) : RPW_OK;
}

int rpw_capacity(const rpw_handle* h, size_t* max_total_points, size_t* max_batch) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (max_total_points) *max_total_points = h->cap_points;
    if (max_batch) *max_batch = h->cap_batch;
    return RPW_OK;
}

int rpw_set_plane_solver(rpw_handle* h, int solver) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (solver != RPW_SOLVER_EIGEN_QR && solver != RPW_SOLVER_CLOSED_FORM && solver != RPW_SOLVER_HYBRID && solver != RPW_SOLVER_REFERENCE)
        RPW_FAIL(h, RPW_ERR_BAD_ARG, "unknown plane solver %d", solver);
    h->solver = solver;
    apply_solver(h);
    return RPW_OK;
}

int rpw_set_exact_replay(rpw_handle* h, int max_fast_iterations) {
    if (!h) return RPW_ERR_BAD_ARG;
    h->exact_replay = max_fast_iterations < 0 ? -1 : max_fast_iterations;
    apply_solver(h);
    return RPW_OK;
}

int rpw_get_config(const rpw_handle* h, rpw_config* out) {
    if (!h || !out) return RPW_ERR_BAD_ARG;
    *out = h->cfg;
    return RPW_OK;
}

int rpw_set_stream(rpw_handle* h, void* cuda_stream) {
    if (!h) return RPW_ERR_BAD_ARG;
    cudaStream_t next = cuda_stream ? (cudaStream_t)cuda_stream : h->own_stream;
    if (next != h->stream) {
        // work still queued on the old stream uses the handle's buffers: the new stream waits for it
        RPW_CUDA(h, cudaSetDevice(h->device));
        RPW_CUDA(h, cudaEventRecord(h->ev_call, h->stream));
        RPW_CUDA(h, cudaStreamWaitEvent(next, h->ev_call, 0));
        h->stream = next;
        h->cfg_epoch++;
    }
    return RPW_OK;
}

const char* rpw_last_error(const rpw_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

uint64_t rpw_kernel_launches(const rpw_handle* h) { return h ? h->launches : 0; }

int rpw_scan_graph(rpw_handle* h, int enable, uint64_t* launches, uint64_t* captures) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (enable >= 0) h->graphs_enabled = enable != 0;
    if (launches) *launches = h->graph_launches;
    if (captures) *captures = h->graph_captures;
    return RPW_OK;
}

void* rpw_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) return nullptr;
    return p;
}

void rpw_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// per-kernel timing
// ---------------------------------------------------------------------------------------------
static int prof_fold(rpw_handle* h) {
    if (h->prof_used == 0) return RPW_OK;
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < h->prof_used; ++i) {
        float ms = 0.f;
        RPW_CUDA(h, cudaEventElapsedTime(&ms, h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
        h->prof_ms[h->prof_kind[i]] += ms;
        h->prof_launches[h->prof_kind[i]]++;
    }
    h->prof_used = 0;
    return RPW_OK;
}

static const char* const kProfNames[RPW_PROF_KERNELS] = {"rpw K1 bin", "rpw K1b offsets", "rpw K2 scatter", "rpw K3 fit"};

struct ProfScope {
    rpw_handle* h;
    size_t slot = (size_t)-1;
    ProfScope(rpw_handle* h_, int kind) : h(h_) {
        nvtxRangePushA(kProfNames[kind]);  // NVTX range around the launch(es) of every stage
        if (!h->prof_enabled) return;
        if (h->prof_used * 2 >= h->prof_ev.size()) {
            if (h->prof_ev.size() >= 2 * 4096) { if (prof_fold(h) != RPW_OK) return; }
            else {
                for (int k = 0; k < 128; ++k) { cudaEvent_t e; if (cudaEventCreate(&e) != cudaSuccess) return; h->prof_ev.push_back(e); }
                h->prof_kind.resize(h->prof_ev.size() / 2);
            }
        }
        slot = h->prof_used++;
        h->prof_kind[slot] = kind;
        cudaEventRecord(h->prof_ev[2 * slot], h->stream);
    }
    ~ProfScope() {
        if (slot != (size_t)-1) cudaEventRecord(h->prof_ev[2 * slot + 1], h->stream);
        nvtxRangePop();
    }
};

// ---------------------------------------------------------------------------------------------
// launch sequence
// ---------------------------------------------------------------------------------------------
// Fills a slot of the pinned meta ring (scan offsets relative to the first scan, chunk bases) and uploads it with one
// copy, unless the call has the same offsets as the previous one.  No synchronisation: a slot is reused eight calls
// later, after its own upload has been consumed (an event per slot).
static int upload_meta(rpw_handle* h, const uint64_t* off, size_t batch) {
    if (batch == 0 || batch > h->cap_batch) RPW_FAIL(h, RPW_ERR_CAPACITY, "batch %zu exceeds the handle's max_batch %zu", batch, h->cap_batch);
    if (off[batch] - off[0] > h->cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%llu points exceed the handle's capacity %zu", (unsigned long long)(off[batch] - off[0]), h->cap_points);
    bool same = h->last_batch == batch && h->last_off.size() == batch + 1;
    for (size_t i = 0; same && i <= batch; ++i) same = h->last_off[i] == off[i] - off[0];
    if (same) return RPW_OK;
    const int slot = h->meta_next;
    h->meta_next = (slot + 1) % rpw_handle::kMetaSlots;
    RPW_CUDA(h, cudaEventSynchronize(h->meta_ev[slot]));  // (an event never recorded counts as complete)
    uint64_t* so = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(h->h_meta) + (size_t)slot * h->meta_slot_bytes);
    uint32_t* cb = reinterpret_cast<uint32_t*>(so + batch + 1);
    uint32_t run = 0;
    for (size_t i = 0; i <= batch; ++i) {
        so[i] = off[i] - off[0];
        cb[i] = run;
        if (i < batch) {
            if (off[i + 1] < off[i]) RPW_FAIL(h, RPW_ERR_BAD_ARG, "scan offsets must be non-decreasing");
            run += (uint32_t)((off[i + 1] - off[i] + kBinChunk - 1) / kBinChunk);
        }
    }
    h->d_chunk_base = reinterpret_cast<uint32_t*>(h->d_scan_off + batch + 1);
    RPW_CUDA(h, cudaMemcpyAsync(h->d_scan_off, so, (batch + 1) * (sizeof(uint64_t) + sizeof(uint32_t)), cudaMemcpyHostToDevice, h->stream));
    RPW_CUDA(h, cudaEventRecord(h->meta_ev[slot], h->stream));
    h->last_off.assign(so, so + batch + 1);
    h->last_batch = batch;
    h->last_total = (size_t)so[batch];
    return RPW_OK;
}

// ProfScope records on h->stream; launch groups run on lane streams, so point it there for a while.
struct StreamSwap {
    rpw_handle* h; cudaStream_t saved;
    StreamSwap(rpw_handle* h_, cudaStream_t s) : h(h_), saved(h_->stream) { h->stream = s; }
    ~StreamSwap() { h->stream = saved; }
};

// One launch group: scans [b0, b0 + nb) of the call on lane `L`, stream `st`.
// graph_chunks > 0: the launches are being captured into the single-scan graph, whose grids must hold for every scan of
// up to graph_chunks chunks (blocks past a scan's end leave at once) and must not depend on the previous call.
static int run_group(rpw_handle* h, rpw_handle::Lane& L, cudaStream_t st, const float* d_pts, const PointLayout& lay, uint8_t* d_labels,
                     size_t b0, size_t nb, int graph_chunks = 0) {
    const uint64_t* so = h->last_off.data();
    uint64_t max_n = 0;
    for (size_t i = b0; i < b0 + nb; ++i) max_n = so[i + 1] - so[i] > max_n ? so[i + 1] - so[i] : max_n;
    if (so[b0 + nb] == so[b0] && !graph_chunks) return RPW_OK;  // nothing but empty scans
    const int max_chunks = graph_chunks ? graph_chunks : (int)((max_n + kBinChunk - 1) / kBinChunk);
    // class table of the level-0 fit: calls of one or two scans are latency-bound (most SMs stay empty, the call ends with
    // its longest-iterating patch) and spread big patches over thread-block clusters; the reference-order kernels do not
    const int profile = h->fp.exact_replay >= 0 ? 0 : (h->fit_profile < 0 ? (nb <= 2 ? 1 : 0) : h->fit_profile);
    StreamSwap swap(h, st);
    // The kernels index scans relative to the pointers they are given.
    const uint64_t* d_so = h->d_scan_off + b0;
    const uint32_t* d_cb = h->d_chunk_base + b0;
    uint32_t* d_ps = h->d_patch_start + b0 * (size_t)(h->P + 1);
    // calls of one or two scans: 1024-thread blocks for K1 / K2 (four points per thread instead of sixteen), as long as the
    // scatter's per-warp offset tables (32 warps x P words) stay within the default shared-memory limit
    static const char* bin_wide_env = getenv("RPW_BIN_WIDE");
    const bool bin_wide = bin_wide_env ? atoi(bin_wide_env) != 0 : (profile == 1 && h->P <= 256);
    const int bin_threads = bin_wide ? 1024 : kBinThreads;
    { ProfScope ps(h, 0);
      RPW_CUDA(h, launch_bin(st, lay, d_pts, d_so, d_cb, h->zm, h->d_keys, d_labels, h->d_blk_hist, L.d_cls_count, h->fusion_arg, max_chunks, (int)nb, bin_threads)); }
    { ProfScope ps(h, 1);
      RPW_CUDA(h, launch_offsets(st, d_so, d_cb, h->d_blk_hist, d_ps, L.d_cls_count, L.d_cls_list, h->cls_cap, h->P, (int)nb, profile)); }
    { ProfScope ps(h, 2);
      RPW_CUDA(h, launch_scatter(st, lay, d_pts, d_so, d_cb, h->d_keys, h->d_blk_hist, d_ps, h->d_sortedA,
                                 h->P, h->fusion_arg, max_chunks, (int)nb, bin_threads)); }
    FitArgs A;
    A.sortedA = h->d_sortedA;
    A.bufB = h->d_bufB; A.bufC = h->d_bufC; A.gmask = h->d_gmask;
    A.labels = d_labels;
    A.patch_start = d_ps;
    A.cls_count = L.d_cls_count;
    A.cls_list = L.d_cls_list;
    A.cls_cap = h->cls_cap;
    A.host_counts = L.d_counts_map;
    A.root_mean = h->d_root_mean + b0 * (size_t)h->P;
    A.queue[0] = L.d_queue[0]; A.queue[1] = L.d_queue[1];
    A.fetch_ctr = L.d_counters;
    A.q_count = L.d_counters + h->levels_cap;
    A.stats = L.d_counters + 2 * (size_t)h->levels_cap;
    A.overflow = A.stats + 8;
    A.timing = h->timing_enabled ? h->d_timing : nullptr;
    A.trace = h->trace_cap ? h->d_trace : nullptr;
    A.trace_count = h->d_trace_count;
    A.trace_cap = h->trace_cap;
    A.dbg_nodes = h->dbg_enabled ? h->d_dbg_nodes : nullptr;
    A.dbg_count = h->d_dbg_count;
    A.dbg_cap = h->dbg_cap;
    A.q_cap = h->q_cap;
    A.n_roots = (int)(nb * (size_t)h->P);
    A.n_scans = (int)nb;
    A.P = h->P;
    A.smem_cap = h->smem_cap;
    A.profile = profile;
    A.scan_base = (uint32_t)b0;
    A.fp = h->fp;
    {
        ProfScope ps(h, 3);
        // level 0: the size classes on side streams (largest first), joined back before the
        // persistent kernel that walks the deeper levels
        static const char* dbg_mask = getenv("RPW_DBG_CLASS_MASK");  // profiling aid: bit k = run size class k only
        const int mask = dbg_mask ? atoi(dbg_mask) : (1 << kNumFitClasses) - 1;
        // Grid of class c: the class's patch count in the last finished launch group of this lane
        // (scaled to this group's scan count, plus slack), capped by what the class can hold at all.
        // Too small only makes some blocks take a second patch, too large starts blocks that leave
        // after one load: the estimate never affects results.
        static const char* no_est = getenv("RPW_NO_GRID_ESTIMATE");
        const ClassBounds cb = fit_class_bounds(A.profile);
        const uint64_t group_pts = graph_chunks ? (uint64_t)graph_chunks * kBinChunk : so[b0 + nb] - so[b0];
        unsigned grids[kNumFitClasses];
        const uint32_t est_scans = graph_chunks ? 0u : L.h_counts[kClsWords - 1];
        for (int c = 0; c < kNumFitClasses; ++c) {
            const uint64_t lo = c == 0 ? 0 : cb.hi[c - 1];
            uint64_t bound = group_pts / (lo + 1);
            if (bound > (uint64_t)nb * (uint64_t)h->P) bound = (uint64_t)nb * (uint64_t)h->P;
            if (bound > h->cls_cap) bound = h->cls_cap;
            uint64_t g = bound;
            if (est_scans && !no_est) {
                const uint64_t est = ((uint64_t)L.h_counts[c] * nb + est_scans - 1) / est_scans;
                g = est + est / 32 + 2;
                // never fewer than two blocks per SM: if the scene changes (a class that was empty fills up)
                // the list is still walked in parallel; surplus blocks cost one load each
                if (g < 2u * (uint64_t)h->num_sms) g = 2u * (uint64_t)h->num_sms;
                if (g > bound) g = bound;
            }
            grids[c] = (unsigned)(g ? g : 1);
        }
        RPW_CUDA(h, cudaEventRecord(L.ev_fork, st));
        // side[k] has priority k (0 = highest): strictly by size, largest class first (other orders, e.g. the
        // smallest class ahead of the second smallest so that it is not left alone at the end, measured 2 % slower)
        // (launching the largest resident class ahead of the streamed one, whose 512-thread blocks take half an SM's
        // registers each: nothing on C2, 8 % slower on the big-patch shapes C4 and C5)
        for (int k = 0; k < kNumFitClasses; ++k) {
            const int cls = kNumFitClasses - 1 - k;
            RPW_CUDA(h, cudaStreamWaitEvent(L.side[k], L.ev_fork, 0));
            if (mask & (1 << cls)) RPW_CUDA(h, launch_fit_roots(L.side[k], A, cls, grids[cls]));
            RPW_CUDA(h, cudaEventRecord(L.ev_join[k], L.side[k]));
        }
        for (int k = 0; k < kNumFitClasses; ++k) RPW_CUDA(h, cudaStreamWaitEvent(st, L.ev_join[k], 0));
        RPW_CUDA(h, launch_fit_levels(st, A, h->fit_blocks));
    }
    h->launches += 4 + kNumFitClasses;
    h->launches_call += 4 + kNumFitClasses;
    return RPW_OK;
}

static bool same_layout(const PointLayout& a, const PointLayout& b) {
    return a.vec4 == b.vec4 && a.stride == b.stride && a.ox == b.ox && a.oy == b.oy && a.oz == b.oz;
}

// One host-path scan through the captured graph (see rpw_handle::ScanGraph).
static int run_scan_graph(rpw_handle* h, const PointLayout& lay) {
    rpw_handle::ScanGraph& g = h->graph1;
    const int chunks = (int)((h->last_total + kBinChunk - 1) / kBinChunk);
    if (!g.exec || g.epoch != h->cfg_epoch || g.stream != h->stream || !same_layout(g.lay, lay) || chunks > g.chunks) {
        if (g.exec) { cudaGraphExecDestroy(g.exec); g.exec = nullptr; }
        int want = chunks + chunks / 4 + 1;  // head room: frames of a stream vary by a few percent
        const int cap_chunks = (int)(h->cap_points / kBinChunk + 1);
        if (want > cap_chunks) want = cap_chunks;
        if (want < chunks) want = chunks;
        const uint64_t launches_before = h->launches;
        cudaGraph_t graph = nullptr;
        RPW_CUDA(h, cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        int rc = RPW_OK;
        for (auto& L : h->lane) {
            uint32_t* st = L.d_counters + 2 * (size_t)h->levels_cap;
            if (cudaMemsetAsync(st, 0, sizeof(uint32_t), h->stream) != cudaSuccess || cudaMemsetAsync(st + 3, 0, sizeof(uint32_t), h->stream) != cudaSuccess) rc = RPW_ERR_CUDA;
        }
        if (rc == RPW_OK) rc = run_group(h, h->lane[0], h->stream, h->d_in, lay, h->d_labels, 0, 1, want);
        const cudaError_t ec = cudaStreamEndCapture(h->stream, &graph);
        h->launches = launches_before;  // capturing launches nothing
        cudaError_t ei = cudaSuccess;
        if (rc == RPW_OK && ec == cudaSuccess && graph) ei = cudaGraphInstantiate(&g.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        if (rc != RPW_OK || ec != cudaSuccess || ei != cudaSuccess || !g.exec) {
            // not fatal: the same launches work outside a graph (the caller falls back to them)
            g.exec = nullptr;
            cudaGetLastError();
            h->graphs_enabled = false;
            h->err = std::string("single-scan graph unavailable (") + cudaGetErrorString(ec != cudaSuccess ? ec : ei) + "), using plain launches";
            return -1;
        }
        g.chunks = want; g.epoch = h->cfg_epoch; g.lay = lay; g.stream = h->stream;
        h->graph_captures++;
    }
    RPW_CUDA(h, cudaGraphLaunch(g.exec, h->stream));
    h->graph_launches++;
    h->launches += 4 + kNumFitClasses;
    h->launches_call = 4 + kNumFitClasses;
    return RPW_OK;
}

// Enqueues K1..K3 for scans [0, batch) whose points are device resident at `d_pts`.
static int run_pipeline(rpw_handle* h, const float* d_pts, const PointLayout& lay, uint8_t* d_labels, size_t batch) {
    h->launches_call = 0;
    h->last_pts = d_pts; h->last_lay = lay; h->last_labels = d_labels; h->last_fused = h->fusion_arg != nullptr;
    if (batch == 1 && h->graphs_enabled && h->last_total > 0 && d_pts == h->d_in && d_labels == h->d_labels && !h->fusion_arg &&
        !h->dbg_enabled && !h->timing_enabled && !h->trace_cap && !h->prof_enabled) {
        const int rc = run_scan_graph(h, lay);
        if (rc >= 0) return rc;  // (-1: no graph on this driver, fall through to the plain launches)
    }
    // stats[0] (levels) and stats[3] (nodes) accumulate over the call's launch groups
    for (auto& L : h->lane) {
        uint32_t* st = L.d_counters + 2 * (size_t)h->levels_cap;
        RPW_CUDA(h, cudaMemsetAsync(st, 0, sizeof(uint32_t), h->stream));
        RPW_CUDA(h, cudaMemsetAsync(st + 3, 0, sizeof(uint32_t), h->stream));
    }
    int waves = h->n_waves > 0 ? h->n_waves : 1;
    if ((size_t)waves > batch) waves = (int)batch;
    if (h->dbg_enabled || h->timing_enabled) waves = 1;  // debug records are easier to read in one group
    if (waves <= 1) return run_group(h, h->lane[0], h->stream, d_pts, lay, d_labels, 0, batch);
    // fork: both lanes wait for everything already enqueued on the handle's stream (inputs, meta)
    RPW_CUDA(h, cudaEventRecord(h->ev_call, h->stream));
    for (auto& L : h->lane) RPW_CUDA(h, cudaStreamWaitEvent(L.main, h->ev_call, 0));
    for (int w = 0; w < waves; ++w) {
        const size_t b0 = batch * (size_t)w / waves, b1 = batch * (size_t)(w + 1) / waves;
        if (b1 <= b0) continue;
        rpw_handle::Lane& L = h->lane[w % rpw_handle::kLanes];
        int rc = run_group(h, L, L.main, d_pts, lay, d_labels, b0, b1 - b0);
        if (rc != RPW_OK) return rc;
    }
    // join
    for (auto& L : h->lane) {
        RPW_CUDA(h, cudaEventRecord(L.ev_done, L.main));
        RPW_CUDA(h, cudaStreamWaitEvent(h->stream, L.ev_done, 0));
    }
    return RPW_OK;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static int ensure_stage(rpw_handle* h, size_t in_bytes) {
    if (in_bytes < h->cap_points * 16) in_bytes = h->cap_points * 16;
    if (!h->h_stage_in || h->h_stage_in_bytes < in_bytes) {
        if (h->h_stage_in) cudaFreeHost(h->h_stage_in);
        h->h_stage_in = nullptr;
        RPW_CUDA(h, cudaMallocHost(&h->h_stage_in, in_bytes));
        h->h_stage_in_bytes = in_bytes;
    }
    if (!h->h_stage_labels) RPW_CUDA(h, cudaMallocHost(&h->h_stage_labels, h->cap_points));
    return RPW_OK;
}

// Wide records (a PointCloud2 point_step above 16 bytes) need a larger device staging area.
static int ensure_d_in(rpw_handle* h, size_t bytes) {
    if (bytes <= h->d_in_bytes) return RPW_OK;
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(h->d_in);
    h->d_in = nullptr;
    h->d_in_bytes = 0;
    RPW_CUDA(h, cudaMalloc(&h->d_in, bytes));
    h->d_in_bytes = bytes;
    return RPW_OK;
}

static bool stride_ok(size_t stride_bytes) { return stride_bytes >= 12 && stride_bytes <= 1024 && stride_bytes % 4 == 0; }

static int reset_dbg(rpw_handle* h) {
    if (h->dbg_enabled) RPW_CUDA(h, cudaMemsetAsync(h->d_dbg_count, 0, sizeof(uint32_t), h->stream));
    return RPW_OK;
}

// A device worklist that overflowed dropped children (their points keep stale labels): the levels kernel reports it
// through mapped host memory, and every entry point that synchronises checks it, with or without a stats request.
static int check_overflow(rpw_handle* h) {
    bool seen = false;
    for (auto& L : h->lane)
        if (L.h_counts && L.h_counts[kClsWords - 2]) { L.h_counts[kClsWords - 2] = 0; seen = true; }
    if (seen) RPW_FAIL(h, RPW_ERR_CAPACITY, "device worklist overflowed (q_cap %u): labels of this call are incomplete", h->q_cap);
    return RPW_OK;
}

static int fill_stats(rpw_handle* h, rpw_stats* st, uint8_t* const* labels, const size_t* n, size_t batch) {
    if (!st) return check_overflow(h);
    memset(st, 0, sizeof(*st));
    for (int l = 0; l < rpw_handle::kLanes; ++l)
        RPW_CUDA(h, cudaMemcpyAsync(h->h_stats + 16 * l, h->lane[l].d_counters + 2 * (size_t)h->levels_cap, 10 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    st->kernel_launches = h->launches_call;
    for (int l = 0; l < rpw_handle::kLanes; ++l) {
        const uint32_t* s = h->h_stats + 16 * l;
        st->n_levels = s[0] > st->n_levels ? s[0] : st->n_levels;
        st->n_nodes += s[3];
    }
    { const int rc = check_overflow(h); if (rc != RPW_OK) return rc; }
    uint64_t cnt[4] = {0, 0, 0, 0};
    for (size_t b = 0; b < batch; ++b) {
        const uint8_t* l = labels[b];
        for (size_t i = 0; i < n[b]; ++i) cnt[l[i] & 3]++;  // (label 4, ego, only occurs in fused frames; rpw_segment_fused recounts)
        st->n_points += n[b];
    }
    st->n_nonground = cnt[0]; st->n_ground = cnt[1]; st->n_beyond = cnt[2]; st->n_dropped = cnt[3];
    return RPW_OK;
}

extern "C" {

int rpw_segment_batch_async(rpw_handle* h, const float* const* clouds, const size_t* n, size_t batch, size_t stride_bytes,
                            uint8_t* const* labels_out) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!clouds || !n || !labels_out || batch == 0) RPW_FAIL(h, RPW_ERR_BAD_ARG, "NULL argument or empty batch");
    if (!stride_ok(stride_bytes)) RPW_FAIL(h, RPW_ERR_BAD_ARG, "stride_bytes must be a multiple of 4 in [12, 1024], got %zu", stride_bytes);
    RPW_CUDA(h, cudaSetDevice(h->device));
    std::vector<uint64_t> off(batch + 1);
    off[0] = 0;
    for (size_t b = 0; b < batch; ++b) {
        if (n[b] && (!clouds[b] || !labels_out[b])) RPW_FAIL(h, RPW_ERR_BAD_ARG, "scan %zu has a NULL buffer", b);
        off[b + 1] = off[b] + n[b];
    }
    if (batch > h->cap_batch) RPW_FAIL(h, RPW_ERR_CAPACITY, "batch %zu exceeds the handle's max_batch %zu", batch, h->cap_batch);
    if (off[batch] > h->cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%llu points exceed the handle's capacity %zu", (unsigned long long)off[batch], h->cap_points);
    if (off[batch] == 0) { h->last_batch = 0; h->last_total = 0; h->last_off.clear(); h->pend_labels.clear(); return RPW_OK; }
    int rc = upload_meta(h, off.data(), batch);
    if (rc != RPW_OK) return rc;
    rc = reset_dbg(h);
    if (rc != RPW_OK) return rc;
    rc = ensure_d_in(h, (size_t)off[batch] * stride_bytes);
    if (rc != RPW_OK) return rc;
    // H2D: one copy per run of scans that are contiguous in host memory
    char* d_in = reinterpret_cast<char*>(h->d_in);
    if (h->src_pitch) {
        // wide records with adjacent x, y, z (a PointCloud2 buffer): a strided copy moves the 12 xyz bytes of every record
        // and leaves the rest (intensity, ring, time stamps ...) on the host: 12 instead of point_step bytes per point
        for (size_t b = 0; b < batch; ++b)
            if (n[b]) RPW_CUDA(h, cudaMemcpy2DAsync(d_in + off[b] * 12, 12, reinterpret_cast<const char*>(clouds[b]) + h->src_off, h->src_pitch, 12, n[b],
                                                    cudaMemcpyHostToDevice, h->stream));
    } else
    for (size_t b = 0; b < batch;) {
        size_t e = b + 1;
        const char* base = reinterpret_cast<const char*>(clouds[b]);
        while (e < batch && reinterpret_cast<const char*>(clouds[e]) == base + (off[e] - off[b]) * stride_bytes) ++e;
        const size_t bytes = (size_t)(off[e] - off[b]) * stride_bytes;
        if (bytes) RPW_CUDA(h, cudaMemcpyAsync(d_in + off[b] * stride_bytes, base, bytes, cudaMemcpyHostToDevice, h->stream));
        b = e;
    }
    PointLayout lay;
    lay.stride = (int)(stride_bytes / 4);
    lay.ox = (int)(h->field_off[0] / 4); lay.oy = (int)(h->field_off[1] / 4); lay.oz = (int)(h->field_off[2] / 4);
    lay.vec4 = (stride_bytes == 16 && lay.ox == 0 && lay.oy == 1 && lay.oz == 2) ? 1 : 0;
    rc = run_pipeline(h, h->d_in, lay, h->d_labels, batch);
    if (rc != RPW_OK) return rc;
    for (size_t b = 0; b < batch;) {
        size_t e = b + 1;
        while (e < batch && labels_out[e] == labels_out[b] + (off[e] - off[b])) ++e;
        const size_t bytes = (size_t)(off[e] - off[b]);
        if (bytes) RPW_CUDA(h, cudaMemcpyAsync(labels_out[b], h->d_labels + off[b], bytes, cudaMemcpyDeviceToHost, h->stream));
        b = e;
    }
    h->pend_labels.assign(labels_out, labels_out + batch);
    return RPW_OK;
}

int rpw_wait(rpw_handle* h, rpw_stats* stats) {
    if (!h) return RPW_ERR_BAD_ARG;
    RPW_CUDA(h, cudaSetDevice(h->device));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    if (stats) {
        if (h->last_batch == 0) { memset(stats, 0, sizeof(*stats)); return RPW_OK; }
        std::vector<size_t> n(h->last_batch);
        for (size_t b = 0; b < h->last_batch; ++b) n[b] = (size_t)(h->last_off[b + 1] - h->last_off[b]);
        if (h->pend_labels.size() != h->last_batch) RPW_FAIL(h, RPW_ERR_BAD_ARG, "no host labels pending for stats");
        return fill_stats(h, stats, h->pend_labels.data(), n.data(), h->last_batch);
    }
    return check_overflow(h);
}

int rpw_segment_batch(rpw_handle* h, const float* const* clouds, const size_t* n, size_t batch, size_t stride_bytes,
                      uint8_t* const* labels_out, rpw_stats* stats) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!clouds || !n || !labels_out || batch == 0) RPW_FAIL(h, RPW_ERR_BAD_ARG, "NULL argument or empty batch");
    if (!stride_ok(stride_bytes)) RPW_FAIL(h, RPW_ERR_BAD_ARG, "stride_bytes must be a multiple of 4 in [12, 1024], got %zu", stride_bytes);
    RPW_CUDA(h, cudaSetDevice(h->device));
    size_t total = 0;
    for (size_t b = 0; b < batch; ++b) total += n[b];
    if (batch > h->cap_batch) RPW_FAIL(h, RPW_ERR_CAPACITY, "batch %zu exceeds the handle's max_batch %zu", batch, h->cap_batch);
    if (total > h->cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%zu points exceed the handle's capacity %zu", total, h->cap_points);
    // Pageable host memory is staged through the handle's pinned buffers so that the DMA engine
    // can run at full rate; pinned caller buffers are used in place.
    // (only the first non-empty scan is probed: a pageable buffer handed to cudaMemcpyAsync is
    // still copied correctly, just not at full rate)
    bool all_pinned = true, probed = false;
    for (size_t b = 0; b < batch; ++b) {
        if (n[b] && (!clouds[b] || !labels_out[b])) RPW_FAIL(h, RPW_ERR_BAD_ARG, "scan %zu has a NULL buffer", b);
        if (n[b] && !probed) { all_pinned = is_pinned(clouds[b]) && is_pinned(labels_out[b]); probed = true; }
    }
    int rc;
    if (all_pinned) {
        rc = rpw_segment_batch_async(h, clouds, n, batch, stride_bytes, labels_out);
        if (rc != RPW_OK) return rc;
        RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    } else {
        rc = ensure_stage(h, total * stride_bytes);
        if (rc != RPW_OK) return rc;
        RPW_CUDA(h, cudaStreamSynchronize(h->stream));
        std::vector<const float*> src(batch);
        std::vector<uint8_t*> dst(batch);
        size_t o = 0;
        const size_t pitch = h->src_pitch, poff = h->src_off;
        for (size_t b = 0; b < batch; ++b) {
            char* s = reinterpret_cast<char*>(h->h_stage_in) + o * stride_bytes;
            if (n[b] && pitch) {  // pack the xyz triples while staging
                const char* from = reinterpret_cast<const char*>(clouds[b]) + poff;
                for (size_t i = 0; i < n[b]; ++i) memcpy(s + i * 12, from + i * pitch, 12);
            } else if (n[b]) memcpy(s, clouds[b], n[b] * stride_bytes);
            src[b] = reinterpret_cast<const float*>(s);
            dst[b] = h->h_stage_labels + o;
            o += n[b];
        }
        h->src_pitch = 0;  // the staged records are packed
        rc = rpw_segment_batch_async(h, src.data(), n, batch, stride_bytes, dst.data());
        h->src_pitch = pitch;
        if (rc != RPW_OK) return rc;
        RPW_CUDA(h, cudaStreamSynchronize(h->stream));
        for (size_t b = 0; b < batch; ++b) if (n[b]) memcpy(labels_out[b], dst[b], n[b]);
    }
    h->pend_labels.assign(labels_out, labels_out + batch);
    if (total == 0) { if (stats) memset(stats, 0, sizeof(*stats)); return RPW_OK; }
    return fill_stats(h, stats, labels_out, n, batch);
}

int rpw_segment(rpw_handle* h, const float* xyz, size_t n, size_t stride_bytes, uint8_t* labels_out, rpw_stats* stats) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (n == 0) {  // empty cloud -> empty clouds (RP/src/recursive_patchwork.cpp:316-318)
        if (stats) memset(stats, 0, sizeof(*stats));
        return RPW_OK;
    }
    const float* c[1] = {xyz};
    uint8_t* l[1] = {labels_out};
    return rpw_segment_batch(h, c, &n, 1, stride_bytes, l, stats);
}

int rpw_segment_pc2(rpw_handle* h, const void* data, size_t n_points, size_t point_step, size_t off_x, size_t off_y, size_t off_z,
                    uint8_t* labels_out, rpw_stats* stats) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!stride_ok(point_step)) RPW_FAIL(h, RPW_ERR_BAD_ARG, "point_step must be a multiple of 4 in [12, 1024], got %zu", point_step);
    const size_t offs[3] = {off_x, off_y, off_z};
    for (size_t o : offs)
        if (o % 4 != 0 || o + 4 > point_step) RPW_FAIL(h, RPW_ERR_BAD_ARG, "field offset %zu does not address a float32 inside a %zu-byte record", o, point_step);
    int rc;
    if (h->pc2_pack && off_y == off_x + 4 && off_z == off_x + 8 && point_step > 12) {
        h->src_pitch = point_step; h->src_off = off_x;
        rc = rpw_segment(h, reinterpret_cast<const float*>(data), n_points, 12, labels_out, stats);
        h->src_pitch = 0; h->src_off = 0;
    } else {
        h->field_off[0] = off_x; h->field_off[1] = off_y; h->field_off[2] = off_z;
        rc = rpw_segment(h, reinterpret_cast<const float*>(data), n_points, point_step, labels_out, stats);
        h->field_off[0] = 0; h->field_off[1] = 4; h->field_off[2] = 8;
    }
    return rc;
}

int rpw_segment_fused(rpw_handle* h, const rpw_sensor_cloud* sensors, size_t n_sensors, size_t stride_bytes,
                      uint8_t* const* labels_out, rpw_stats* stats) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!sensors || !labels_out || n_sensors == 0 || n_sensors > (size_t)kMaxSensors)
        RPW_FAIL(h, RPW_ERR_BAD_ARG, "need 1..%d sensors", kMaxSensors);
    if (!stride_ok(stride_bytes)) RPW_FAIL(h, RPW_ERR_BAD_ARG, "stride_bytes must be a multiple of 4 in [12, 1024], got %zu", stride_bytes);
    RPW_CUDA(h, cudaSetDevice(h->device));
    size_t total = 0;
    for (size_t s = 0; s < n_sensors; ++s) {
        if (sensors[s].n && (!sensors[s].xyz || !labels_out[s])) RPW_FAIL(h, RPW_ERR_BAD_ARG, "sensor %zu has a NULL buffer", s);
        total += sensors[s].n;
    }
    if (stats) memset(stats, 0, sizeof(*stats));
    if (total == 0) return RPW_OK;
    if (total > h->cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%zu points exceed the handle's capacity %zu", total, h->cap_points);
    int rc = ensure_stage(h, total * stride_bytes);
    if (rc != RPW_OK) return rc;
    rc = ensure_d_in(h, total * stride_bytes);
    if (rc != RPW_OK) return rc;
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));  // staging buffers and the sensor table are reused
    // sensor table: angle in radians and its cosine / sine exactly as LidarFusion::applyRotation2D
    // computes them on the host (lidar_fusion.cpp:111-113): float(deg * M_PI / 180.0f), cosf, sinf
    FusionTable& ft = *h->h_fusion;
    memset(&ft, 0, sizeof(ft));
    ft.n = (int)n_sensors;
    size_t o = 0;
    char* stage = reinterpret_cast<char*>(h->h_stage_in);
    for (size_t s = 0; s < n_sensors; ++s) {
        ft.start[s] = (uint32_t)o;
        const float deg = sensors[s].rotation_deg;
        const float rad = (float)((double)deg * 3.14159265358979323846 / (double)180.0f);
        ft.cos_a[s] = cosf(rad);
        ft.sin_a[s] = sinf(rad);
        ft.rotate[s] = fabsf(deg) > 1e-6f ? 1 : 0;
        ft.ego[s] = sensors[s].ego_radius;
        if (sensors[s].n) memcpy(stage + o * stride_bytes, sensors[s].xyz, sensors[s].n * stride_bytes);
        o += sensors[s].n;
    }
    for (size_t s = n_sensors; s <= (size_t)kMaxSensors; ++s) ft.start[s] = (uint32_t)total;
    RPW_CUDA(h, cudaMemcpyAsync(h->d_fusion, &ft, sizeof(ft), cudaMemcpyHostToDevice, h->stream));
    const uint64_t off[2] = {0, (uint64_t)total};
    rc = upload_meta(h, off, 1);
    if (rc != RPW_OK) return rc;
    rc = reset_dbg(h);
    if (rc != RPW_OK) return rc;
    RPW_CUDA(h, cudaMemcpyAsync(h->d_in, stage, total * stride_bytes, cudaMemcpyHostToDevice, h->stream));
    PointLayout lay;
    lay.stride = (int)(stride_bytes / 4); lay.ox = 0; lay.oy = 1; lay.oz = 2;
    lay.vec4 = stride_bytes == 16 ? 1 : 0;
    h->fusion_arg = h->d_fusion;
    rc = run_pipeline(h, h->d_in, lay, h->d_labels, 1);
    h->fusion_arg = nullptr;
    if (rc != RPW_OK) return rc;
    RPW_CUDA(h, cudaMemcpyAsync(h->h_stage_labels, h->d_labels, total, cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    { const int rc2 = check_overflow(h); if (rc2 != RPW_OK) return rc2; }
    o = 0;
    for (size_t s = 0; s < n_sensors; ++s) {
        if (sensors[s].n) memcpy(labels_out[s], h->h_stage_labels + o, sensors[s].n);
        o += sensors[s].n;
    }
    h->pend_labels.clear();
    if (stats) {
        uint8_t* one[1] = {h->h_stage_labels};
        const size_t nn[1] = {total};
        rc = fill_stats(h, stats, one, nn, 1);
        if (rc != RPW_OK) return rc;
        // label 4 (ego) was folded into "dropped" by the 2-bit histogram; count it properly
        uint64_t ego = 0, dropped = 0;
        for (size_t i = 0; i < total; ++i) { ego += h->h_stage_labels[i] == RPW_LABEL_EGO; dropped += h->h_stage_labels[i] == RPW_LABEL_DROPPED; }
        stats->n_dropped = dropped;
        stats->n_nonground = total - stats->n_ground - stats->n_beyond - dropped - ego;
    }
    return RPW_OK;
}

int rpw_last_clouds(rpw_handle* h, float* ground_xyz, float* nonground_xyz, int on_device, uint64_t* counts) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!counts) RPW_FAIL(h, RPW_ERR_BAD_ARG, "counts must not be NULL");
    const size_t batch = h->last_batch;
    for (size_t b = 0; b < 2 * batch; ++b) counts[b] = 0;
    if (batch == 0 || h->last_total == 0) return RPW_OK;
    if (!h->last_pts || !h->last_labels) RPW_FAIL(h, RPW_ERR_BAD_ARG, "no segmented scans to assemble clouds from");
    if (on_device && (!ground_xyz || !nonground_xyz)) RPW_FAIL(h, RPW_ERR_BAD_ARG, "device cloud buffers must not be NULL");
    RPW_CUDA(h, cudaSetDevice(h->device));
    if (!h->d_cmp_cnt) {
        const size_t rows = h->cap_points / kBinChunk + h->cap_batch + 1;
        RPW_CUDA(h, cudaMalloc(&h->d_cmp_cnt, rows * 4 * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMalloc(&h->d_scan_counts, h->cap_batch * 2 * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMallocHost(&h->h_scan_counts, h->cap_batch * 2 * sizeof(uint32_t)));
    }
    float* dg = ground_xyz;
    float* dng = nonground_xyz;
    if (!on_device) {
        if (!h->d_cloud_g) {
            RPW_CUDA(h, cudaMalloc(&h->d_cloud_g, h->cap_points * 3 * sizeof(float)));
            RPW_CUDA(h, cudaMalloc(&h->d_cloud_ng, h->cap_points * 3 * sizeof(float)));
        }
        dg = h->d_cloud_g; dng = h->d_cloud_ng;
    }
    const uint64_t* so = h->last_off.data();
    uint64_t max_n = 0;
    for (size_t i = 0; i < batch; ++i) max_n = so[i + 1] - so[i] > max_n ? so[i + 1] - so[i] : max_n;
    const int max_chunks = (int)((max_n + kBinChunk - 1) / kBinChunk);
    nvtxRangePushA("rpw K4 result assembly");
    const cudaError_t ek4 = launch_compact(h->stream, h->last_lay, h->last_pts, h->last_labels, h->d_scan_off, h->d_chunk_base, h->d_cmp_cnt,
                                           h->last_fused ? h->d_fusion : nullptr, dg, dng, h->d_scan_counts, max_chunks, (int)batch, 0, batch <= 2 ? 1024 : kBinThreads);
    nvtxRangePop();
    RPW_CUDA(h, ek4);
    h->launches += 2;
    RPW_CUDA(h, cudaMemcpyAsync(h->h_scan_counts, h->d_scan_counts, batch * 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    { const int rc = check_overflow(h); if (rc != RPW_OK) return rc; }
    for (size_t b = 0; b < 2 * batch; ++b) counts[b] = h->h_scan_counts[b];
    if (!on_device) {
        // exactly the bytes of the clouds: scan b's clouds start at record (scan offset b) of the caller's buffers
        for (size_t b = 0; b < batch; ++b) {
            const size_t o = (size_t)(so[b] - so[0]) * 3;
            if (ground_xyz && counts[2 * b])
                RPW_CUDA(h, cudaMemcpyAsync(ground_xyz + o, dg + o, counts[2 * b] * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
            if (nonground_xyz && counts[2 * b + 1])
                RPW_CUDA(h, cudaMemcpyAsync(nonground_xyz + o, dng + o, counts[2 * b + 1] * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        }
        RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RPW_OK;
}

int rpw_sample_ground_and_obstacles(rpw_handle* h, float target_height, float base_tol, float ego_radius, size_t sample_size,
                                    uint64_t seed, float* out_xyz, size_t out_cap_points, size_t* n_ground_sample, size_t* n_obstacles) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (n_ground_sample) *n_ground_sample = 0;
    if (n_obstacles) *n_obstacles = 0;
    if (!out_xyz) RPW_FAIL(h, RPW_ERR_BAD_ARG, "out_xyz must not be NULL");
    if (h->last_batch > 1) RPW_FAIL(h, RPW_ERR_BAD_ARG, "the last call segmented %zu scans; this post-filter takes one", h->last_batch);
    if (h->last_batch == 0 || h->last_total == 0) return RPW_OK;
    uint64_t counts[2] = {0, 0};
    int rc = rpw_last_clouds(h, nullptr, nullptr, 0, counts);  // the two clouds, left on the device
    if (rc != RPW_OK) return rc;
    const size_t n_g = (size_t)counts[0], n_ng = (size_t)counts[1];
    if (n_ng == 0) {  // :435-437: without non-ground points the whole ground cloud is returned
        if (n_g > out_cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%zu points do not fit the output buffer (%zu)", n_g, out_cap_points);
        if (n_g) RPW_CUDA(h, cudaMemcpyAsync(out_xyz, h->d_cloud_g, n_g * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        RPW_CUDA(h, cudaStreamSynchronize(h->stream));
        if (n_ground_sample) *n_ground_sample = n_g;
        return RPW_OK;
    }
    // ground context: `sample_size` distinct indices drawn as PointCloudProcessor::randomSubsample draws them
    // (uniform draws, repeats rejected, order of the draws kept; point_cloud_processor.cpp:122-148)
    std::vector<uint32_t> idx;
    if (n_g <= sample_size) {
        idx.resize(n_g);
        for (size_t i = 0; i < n_g; ++i) idx[i] = (uint32_t)i;
    } else {
        std::mt19937 gen(seed ? (std::mt19937::result_type)seed : std::random_device{}());
        std::uniform_int_distribution<size_t> dis(0, n_g - 1);
        std::vector<bool> selected(n_g, false);
        idx.reserve(sample_size);
        while (idx.size() < sample_size) {
            const size_t i = dis(gen);
            if (!selected[i]) { selected[i] = true; idx.push_back((uint32_t)i); }
        }
    }
    const size_t k = idx.size();
    if (3 * (k + n_ng) > 4 * h->cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "sample of %zu points exceeds the device scratch", k);
    if (k > h->sample_idx_cap) {
        cudaFree(h->d_sample_idx);
        h->d_sample_idx = nullptr;
        h->sample_idx_cap = 0;
        RPW_CUDA(h, cudaMalloc(&h->d_sample_idx, k * sizeof(uint32_t)));
        h->sample_idx_cap = k;
    }
    float* d_out = reinterpret_cast<float*>(h->d_bufB);  // a level buffer of the fit: free between calls
    if (k) {
        RPW_CUDA(h, cudaMemcpyAsync(h->d_sample_idx, idx.data(), k * sizeof(uint32_t), cudaMemcpyHostToDevice, h->stream));
        RPW_CUDA(h, launch_gather_xyz(h->stream, h->d_cloud_g, h->d_sample_idx, (uint32_t)k, d_out));
    }
    RPW_CUDA(h, launch_obstacles(h->stream, h->d_cloud_ng, (uint32_t)n_ng, target_height, base_tol, ego_radius, h->d_cmp_cnt,
                                 d_out + 3 * k, h->d_scan_counts));
    h->launches += 3;
    RPW_CUDA(h, cudaMemcpyAsync(h->h_scan_counts, h->d_scan_counts, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));  // (also orders the index upload before `idx` goes out of scope)
    const size_t n_obs = h->h_scan_counts[0];
    if (k + n_obs > out_cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%zu points do not fit the output buffer (%zu)", k + n_obs, out_cap_points);
    if (k + n_obs) RPW_CUDA(h, cudaMemcpyAsync(out_xyz, d_out, (k + n_obs) * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    if (n_ground_sample) *n_ground_sample = k;
    if (n_obstacles) *n_obstacles = n_obs;
    return RPW_OK;
}

int rpw_bev_image(rpw_handle* h, int mode, int width, int height, float x_min, float y_min, float x_max, float y_max, uint8_t* bgr_out) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!bgr_out || width <= 0 || height <= 0 || (long long)width * height > (1ll << 28))
        RPW_FAIL(h, RPW_ERR_BAD_ARG, "bad image buffer or size %d x %d", width, height);
    if (mode != RPW_BEV_CLASSES && mode != RPW_BEV_HEIGHT_NONGROUND && mode != RPW_BEV_HEIGHT_ALL) RPW_FAIL(h, RPW_ERR_BAD_ARG, "unknown BEV mode %d", mode);
    if (h->last_batch > 1) RPW_FAIL(h, RPW_ERR_BAD_ARG, "the last call segmented %zu scans; the raster takes one", h->last_batch);
    const size_t n_pixels = (size_t)width * (size_t)height;
    if (h->last_batch == 0 || h->last_total == 0) { memset(bgr_out, 0, n_pixels * 3); return RPW_OK; }
    uint64_t counts[2] = {0, 0};
    int rc = rpw_last_clouds(h, nullptr, nullptr, 0, counts);  // the two clouds, left on the device
    if (rc != RPW_OK) return rc;
    if (n_pixels > h->bev_pixels_cap) {
        cudaFree(h->d_bev_owner);
        h->d_bev_owner = nullptr;
        h->bev_pixels_cap = 0;
        RPW_CUDA(h, cudaMalloc(&h->d_bev_owner, n_pixels * (sizeof(uint32_t) + 3)));
        h->bev_pixels_cap = n_pixels;
    }
    uint8_t* d_img = reinterpret_cast<uint8_t*>(h->d_bev_owner + n_pixels);
    // scale factors as the reference computes them (visualization.cpp:28-29)
    const float x_scale = static_cast<float>(width) / (x_max - x_min), y_scale = static_cast<float>(height) / (y_max - y_min);
    const uint32_t n_g = (uint32_t)counts[0], n_ng = (uint32_t)counts[1];
    if (mode == RPW_BEV_HEIGHT_NONGROUND)
        RPW_CUDA(h, launch_bev(h->stream, 1, h->d_cloud_ng, n_ng, nullptr, 0, width, height, x_min, y_min, x_scale, y_scale, h->d_bev_owner, d_img));
    else
        RPW_CUDA(h, launch_bev(h->stream, mode == RPW_BEV_CLASSES ? 0 : 1, h->d_cloud_g, n_g, h->d_cloud_ng, n_ng, width, height, x_min, y_min,
                               x_scale, y_scale, h->d_bev_owner, d_img));
    h->launches += 3;
    RPW_CUDA(h, cudaMemcpyAsync(bgr_out, d_img, n_pixels * 3, cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPW_OK;
}

// One scan with both clouds, everything enqueued before the only synchronisation: input copy, the pipeline (its graph),
// the result-assembly kernels writing both clouds into one buffer (ground first), then labels, counts and the cloud
// buffer back into pinned staging.  This is what the C++ drop-in's filterGroundPoints costs per ROS2 callback.
int rpw_segment_clouds_view(rpw_handle* h, const float* xyz, size_t n, size_t stride_bytes, uint8_t* labels_out,
                            const float** ground_xyz, size_t* n_ground, const float** nonground_xyz, size_t* n_nonground) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (n_ground) *n_ground = 0;
    if (n_nonground) *n_nonground = 0;
    if (ground_xyz) *ground_xyz = nullptr;
    if (nonground_xyz) *nonground_xyz = nullptr;
    if (n == 0) return RPW_OK;
    if (!xyz) RPW_FAIL(h, RPW_ERR_BAD_ARG, "NULL cloud");
    if (!stride_ok(stride_bytes)) RPW_FAIL(h, RPW_ERR_BAD_ARG, "stride_bytes must be a multiple of 4 in [12, 1024], got %zu", stride_bytes);
    if (n > h->cap_points) RPW_FAIL(h, RPW_ERR_CAPACITY, "%zu points exceed the handle's capacity %zu", n, h->cap_points);
    RPW_CUDA(h, cudaSetDevice(h->device));
    int rc = ensure_stage(h, n * stride_bytes);
    if (rc != RPW_OK) return rc;
    if (!h->d_cmp_cnt) {
        const size_t rows = h->cap_points / kBinChunk + h->cap_batch + 1;
        RPW_CUDA(h, cudaMalloc(&h->d_cmp_cnt, rows * 4 * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMalloc(&h->d_scan_counts, h->cap_batch * 2 * sizeof(uint32_t)));
        RPW_CUDA(h, cudaMallocHost(&h->h_scan_counts, h->cap_batch * 2 * sizeof(uint32_t)));
    }
    if (!h->d_cloud_g) {
        RPW_CUDA(h, cudaMalloc(&h->d_cloud_g, h->cap_points * 3 * sizeof(float)));
        RPW_CUDA(h, cudaMalloc(&h->d_cloud_ng, h->cap_points * 3 * sizeof(float)));
    }
    if (!h->h_cloud_stage) RPW_CUDA(h, cudaMallocHost(&h->h_cloud_stage, h->cap_points * 3 * sizeof(float)));
    // the staging buffers are about to be rewritten: the previous call's copies must be done (they are, unless the
    // caller mixed in asynchronous calls)
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    const uint64_t off[2] = {0, (uint64_t)n};
    rc = upload_meta(h, off, 1);
    if (rc != RPW_OK) return rc;
    rc = reset_dbg(h);
    if (rc != RPW_OK) return rc;
    rc = ensure_d_in(h, n * stride_bytes);
    if (rc != RPW_OK) return rc;
    if (is_pinned(xyz)) {
        RPW_CUDA(h, cudaMemcpyAsync(h->d_in, xyz, n * stride_bytes, cudaMemcpyHostToDevice, h->stream));
    } else {
        // pageable input (the std::vector of the reference's callers): staged through pinned memory in four pieces, the
        // copy engine moving one piece while the CPU stages the next
        const size_t bytes = n * stride_bytes, piece = (bytes / 4 + 4095) / 4096 * 4096;
        for (size_t o = 0; o < bytes; o += piece) {
            const size_t len = bytes - o < piece ? bytes - o : piece;
            memcpy(reinterpret_cast<char*>(h->h_stage_in) + o, reinterpret_cast<const char*>(xyz) + o, len);
            RPW_CUDA(h, cudaMemcpyAsync(reinterpret_cast<char*>(h->d_in) + o, reinterpret_cast<char*>(h->h_stage_in) + o, len, cudaMemcpyHostToDevice, h->stream));
        }
    }
    PointLayout lay;
    lay.stride = (int)(stride_bytes / 4);
    lay.ox = (int)(h->field_off[0] / 4); lay.oy = (int)(h->field_off[1] / 4); lay.oz = (int)(h->field_off[2] / 4);
    lay.vec4 = (stride_bytes == 16 && lay.ox == 0 && lay.oy == 1 && lay.oz == 2) ? 1 : 0;
    rc = run_pipeline(h, h->d_in, lay, h->d_labels, 1);
    if (rc != RPW_OK) return rc;
    const int max_chunks = (int)((n + kBinChunk - 1) / kBinChunk);
    RPW_CUDA(h, launch_compact(h->stream, lay, h->d_in, h->d_labels, h->d_scan_off, h->d_chunk_base, h->d_cmp_cnt, nullptr,
                               h->d_cloud_g, h->d_cloud_g, h->d_scan_counts, max_chunks, 1, /*packed*/1, /*threads*/1024));
    h->launches += 2;
    uint8_t* lab_dst = labels_out && is_pinned(labels_out) ? labels_out : h->h_stage_labels;
    if (labels_out) RPW_CUDA(h, cudaMemcpyAsync(lab_dst, h->d_labels, n, cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaMemcpyAsync(h->h_scan_counts, h->d_scan_counts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaMemcpyAsync(h->h_cloud_stage, h->d_cloud_g, n * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    rc = check_overflow(h);
    if (rc != RPW_OK) return rc;
    if (labels_out && lab_dst != labels_out) memcpy(labels_out, lab_dst, n);
    h->pend_labels.clear();
    const size_t ng = h->h_scan_counts[0], nn = h->h_scan_counts[1];
    if (n_ground) *n_ground = ng;
    if (n_nonground) *n_nonground = nn;
    if (ground_xyz) *ground_xyz = h->h_cloud_stage;
    if (nonground_xyz) *nonground_xyz = h->h_cloud_stage + 3 * ng;
    return RPW_OK;
}

int rpw_segment_clouds(rpw_handle* h, const float* xyz, size_t n, size_t stride_bytes, uint8_t* labels_out,
                       float* ground_xyz, size_t* n_ground, float* nonground_xyz, size_t* n_nonground) {
    const float *g = nullptr, *ng = nullptr;
    size_t cg = 0, cn = 0;
    const int rc = rpw_segment_clouds_view(h, xyz, n, stride_bytes, labels_out, &g, &cg, &ng, &cn);
    if (n_ground) *n_ground = cg;
    if (n_nonground) *n_nonground = cn;
    if (rc != RPW_OK) return rc;
    if (ground_xyz && cg) memcpy(ground_xyz, g, cg * 3 * sizeof(float));
    if (nonground_xyz && cn) memcpy(nonground_xyz, ng, cn * 3 * sizeof(float));
    return RPW_OK;
}

int rpw_segment_device(rpw_handle* h, const void* d_points, const uint64_t* scan_offsets, size_t batch, void* d_labels) {
    if (!h) return RPW_ERR_BAD_ARG;
    if (!d_points || !scan_offsets || !d_labels || batch == 0) RPW_FAIL(h, RPW_ERR_BAD_ARG, "NULL argument or empty batch");
    RPW_CUDA(h, cudaSetDevice(h->device));
    int rc = upload_meta(h, scan_offsets, batch);
    if (rc != RPW_OK) return rc;
    if (h->last_total == 0) return RPW_OK;
    rc = reset_dbg(h);
    if (rc != RPW_OK) return rc;
    h->pend_labels.clear();
    const float* pts = reinterpret_cast<const float*>(d_points) + scan_offsets[0] * 4;
    PointLayout lay;
    lay.vec4 = 1; lay.stride = 4; lay.ox = 0; lay.oy = 1; lay.oz = 2;
    return run_pipeline(h, pts, lay, reinterpret_cast<uint8_t*>(d_labels) + scan_offsets[0], batch);
}

int rpw_debug_keys(rpw_handle* h, uint16_t* keys_out, size_t n_total) {
    if (!h || !keys_out) return RPW_ERR_BAD_ARG;
    if (n_total > h->last_total) RPW_FAIL(h, RPW_ERR_BAD_ARG, "last call had %zu points", h->last_total);
    RPW_CUDA(h, cudaSetDevice(h->device));
    RPW_CUDA(h, cudaMemcpyAsync(keys_out, h->d_keys, n_total * sizeof(uint16_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    return RPW_OK;
}

int rpw_debug_enable_nodes(rpw_handle* h, int enable) {
    if (!h) return RPW_ERR_BAD_ARG;
    RPW_CUDA(h, cudaSetDevice(h->device));
    if (enable && !h->d_dbg_nodes) {
        size_t cap = h->cap_points / 8 + h->cap_batch * (size_t)h->P * 2 + 1024;
        if (cap > (1u << 22)) cap = 1u << 22;
        RPW_CUDA(h, cudaMalloc(&h->d_dbg_nodes, cap * sizeof(rpw_node)));
        h->dbg_cap = (uint32_t)cap;
    }
    h->dbg_enabled = enable != 0;
    return RPW_OK;
}

int rpw_debug_nodes(rpw_handle* h, rpw_node* out, size_t cap, size_t* count) {
    if (!h || !count) return RPW_ERR_BAD_ARG;
    if (!h->dbg_enabled) RPW_FAIL(h, RPW_ERR_BAD_ARG, "node recording is not enabled");
    RPW_CUDA(h, cudaSetDevice(h->device));
    uint32_t c = 0;
    RPW_CUDA(h, cudaMemcpyAsync(&c, h->d_dbg_count, sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    *count = c;
    size_t take = c < h->dbg_cap ? c : h->dbg_cap;
    if (take > cap) take = cap;
    if (out && take) {
        RPW_CUDA(h, cudaMemcpyAsync(out, h->d_dbg_nodes, take * sizeof(rpw_node), cudaMemcpyDeviceToHost, h->stream));
        RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    }
    return RPW_OK;
}

int rpw_debug_eig3(rpw_handle* h, const float* mats, size_t count, float* evals, float* evecs) {
    if (!h || !mats || !evals || !evecs) return RPW_ERR_BAD_ARG;
    if (count == 0) return RPW_OK;
    RPW_CUDA(h, cudaSetDevice(h->device));
    float *dm = nullptr, *dv = nullptr, *dq = nullptr;
    RPW_CUDA(h, cudaMalloc(&dm, count * 9 * sizeof(float)));
    RPW_CUDA(h, cudaMalloc(&dv, count * 3 * sizeof(float)));
    RPW_CUDA(h, cudaMalloc(&dq, count * 9 * sizeof(float)));
    RPW_CUDA(h, cudaMemcpyAsync(dm, mats, count * 9 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RPW_CUDA(h, launch_eig3(h->stream, dm, count, dv, dq));
    h->launches++;
    RPW_CUDA(h, cudaMemcpyAsync(evals, dv, count * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaMemcpyAsync(evecs, dq, count * 9 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dm); cudaFree(dv); cudaFree(dq);
    return RPW_OK;
}

int rpw_debug_normal(rpw_handle* h, const float* scatter6, size_t count, int mode, float* normals3, uint32_t* cycles) {
    if (!h || !scatter6 || !normals3 || !cycles) return RPW_ERR_BAD_ARG;
    if (count == 0) return RPW_OK;
    RPW_CUDA(h, cudaSetDevice(h->device));
    float *ds = nullptr, *dn = nullptr;
    uint32_t* dc = nullptr;
    RPW_CUDA(h, cudaMalloc(&ds, count * 6 * sizeof(float)));
    RPW_CUDA(h, cudaMalloc(&dn, count * 3 * sizeof(float)));
    RPW_CUDA(h, cudaMalloc(&dc, count * sizeof(uint32_t)));
    RPW_CUDA(h, cudaMemcpyAsync(ds, scatter6, count * 6 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RPW_CUDA(h, launch_normal(h->stream, ds, count, mode, dn, dc));
    h->launches++;
    RPW_CUDA(h, cudaMemcpyAsync(normals3, dn, count * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaMemcpyAsync(cycles, dc, count * sizeof(uint32_t), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(ds); cudaFree(dn); cudaFree(dc);
    return RPW_OK;
}

int rpw_debug_atan2(rpw_handle* h, const float* y, const float* x, size_t count, float* out) {
    if (!h || !y || !x || !out) return RPW_ERR_BAD_ARG;
    if (count == 0) return RPW_OK;
    RPW_CUDA(h, cudaSetDevice(h->device));
    float *dy = nullptr, *dx = nullptr, *dout = nullptr;
    RPW_CUDA(h, cudaMalloc(&dy, count * sizeof(float)));
    RPW_CUDA(h, cudaMalloc(&dx, count * sizeof(float)));
    RPW_CUDA(h, cudaMalloc(&dout, count * sizeof(float)));
    RPW_CUDA(h, cudaMemcpyAsync(dy, y, count * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RPW_CUDA(h, cudaMemcpyAsync(dx, x, count * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    RPW_CUDA(h, launch_atan2(h->stream, dy, dx, count, dout));
    h->launches++;
    RPW_CUDA(h, cudaMemcpyAsync(out, dout, count * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    cudaFree(dy); cudaFree(dx); cudaFree(dout);
    return RPW_OK;
}

int rpw_debug_fit_timing(rpw_handle* h, int enable, uint64_t* cycles16) {
    if (!h) return RPW_ERR_BAD_ARG;
    RPW_CUDA(h, cudaSetDevice(h->device));
    if (!h->d_timing) {
        RPW_CUDA(h, cudaMalloc(&h->d_timing, 16 * sizeof(unsigned long long)));
        RPW_CUDA(h, cudaMemset(h->d_timing, 0, 16 * sizeof(unsigned long long)));
    }
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    if (cycles16) RPW_CUDA(h, cudaMemcpy(cycles16, h->d_timing, 16 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    RPW_CUDA(h, cudaMemset(h->d_timing, 0, 16 * sizeof(unsigned long long)));
    h->timing_enabled = enable != 0;
    return RPW_OK;
}

int rpw_debug_fit_trace(rpw_handle* h, rpw_trace_rec* out, size_t cap, size_t* count) {
    if (!h) return RPW_ERR_BAD_ARG;
    RPW_CUDA(h, cudaSetDevice(h->device));
    RPW_CUDA(h, cudaStreamSynchronize(h->stream));
    if (!h->d_trace_count) {
        RPW_CUDA(h, cudaMalloc(&h->d_trace_count, sizeof(uint32_t)));
        RPW_CUDA(h, cudaMemset(h->d_trace_count, 0, sizeof(uint32_t)));
    }
    if (!out) {  // (re)arm
        if (cap > 0xFFFFFFFFull) RPW_FAIL(h, RPW_ERR_BAD_ARG, "trace capacity %zu too large", cap);
        if (cap > h->trace_cap) {
            cudaFree(h->d_trace);
            h->d_trace = nullptr;
            h->trace_cap = 0;
            RPW_CUDA(h, cudaMalloc(&h->d_trace, cap * sizeof(rpw_trace_rec)));
        }
        h->trace_cap = (uint32_t)cap;
        RPW_CUDA(h, cudaMemset(h->d_trace_count, 0, sizeof(uint32_t)));
        if (count) *count = 0;
        return RPW_OK;
    }
    uint32_t seen = 0;
    RPW_CUDA(h, cudaMemcpy(&seen, h->d_trace_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    size_t take = seen < h->trace_cap ? seen : h->trace_cap;
    if (take > cap) take = cap;
    if (take) RPW_CUDA(h, cudaMemcpy(out, h->d_trace, take * sizeof(rpw_trace_rec), cudaMemcpyDeviceToHost));
    RPW_CUDA(h, cudaMemset(h->d_trace_count, 0, sizeof(uint32_t)));
    if (count) *count = seen;
    return RPW_OK;
}

int rpw_profile_enable(rpw_handle* h, int enable) {
    if (!h) return RPW_ERR_BAD_ARG;
    RPW_CUDA(h, cudaSetDevice(h->device));
    int rc = prof_fold(h);
    if (rc != RPW_OK) return rc;
    h->prof_enabled = enable != 0;
    for (int k = 0; k < RPW_PROF_KERNELS; ++k) { h->prof_ms[k] = 0; h->prof_launches[k] = 0; }
    return RPW_OK;
}

int rpw_copy_probe(int device, size_t bytes, int reps, int flags, double* seconds) {
    if (!seconds || bytes == 0 || reps <= 0) return RPW_ERR_BAD_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return RPW_ERR_NO_DEVICE;
    const bool d2h = flags & 1, both = flags & 4;
    const size_t back = both ? (bytes / 12 ? bytes / 12 : 1) : 0;  // labels: one byte back per 12 bytes in
    void *hbuf = nullptr, *dbuf = nullptr, *hbuf2 = nullptr, *dbuf2 = nullptr;
    cudaStream_t st = nullptr, st2 = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
    int rc = RPW_ERR_CUDA;
    if (cudaHostAlloc(&hbuf, bytes, (flags & 2) ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess &&
        cudaMalloc(&dbuf, bytes) == cudaSuccess && cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess &&
        cudaStreamCreateWithFlags(&st2, cudaStreamNonBlocking) == cudaSuccess &&
        cudaEventCreate(&e0) == cudaSuccess && cudaEventCreate(&e1) == cudaSuccess && cudaEventCreateWithFlags(&e2, cudaEventDisableTiming) == cudaSuccess &&
        (!both || (cudaHostAlloc(&hbuf2, back, cudaHostAllocDefault) == cudaSuccess && cudaMalloc(&dbuf2, back) == cudaSuccess))) {
        memset(hbuf, 1, bytes);
        cudaMemcpyAsync(d2h ? hbuf : dbuf, d2h ? dbuf : hbuf, bytes, d2h ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, st);  // warm-up
        cudaStreamSynchronize(st);
        cudaEventRecord(e0, st);
        if (both) cudaStreamWaitEvent(st2, e0, 0);
        for (int r = 0; r < reps; ++r) {
            cudaMemcpyAsync(d2h ? hbuf : dbuf, d2h ? dbuf : hbuf, bytes, d2h ? cudaMemcpyDeviceToHost : cudaMemcpyHostToDevice, st);
            if (both) cudaMemcpyAsync(hbuf2, dbuf2, back, cudaMemcpyDeviceToHost, st2);  // the other copy engine, at the same time
        }
        if (both) { cudaEventRecord(e2, st2); cudaStreamWaitEvent(st, e2, 0); }
        cudaEventRecord(e1, st);
        float ms = 0.f;
        if (cudaStreamSynchronize(st) == cudaSuccess && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) { *seconds = ms * 1e-3; rc = RPW_OK; }
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    if (e2) cudaEventDestroy(e2);
    if (st) cudaStreamDestroy(st);
    if (st2) cudaStreamDestroy(st2);
    cudaFree(dbuf); cudaFree(dbuf2);
    if (hbuf) cudaFreeHost(hbuf);
    if (hbuf2) cudaFreeHost(hbuf2);
    return rc;
}

int rpw_profile_read(rpw_handle* h, rpw_profile* out) {
    if (!h || !out) return RPW_ERR_BAD_ARG;
    RPW_CUDA(h, cudaSetDevice(h->device));
    int rc = prof_fold(h);
    if (rc != RPW_OK) return rc;
    for (int k = 0; k < RPW_PROF_KERNELS; ++k) { out->ms[k] = h->prof_ms[k]; out->launches[k] = h->prof_launches[k]; }
    out->fit_grid_blocks = (uint32_t)h->fit_blocks;
    out->fit_smem_points = (uint32_t)h->smem_cap;
    return RPW_OK;
}

}  // extern "C"
