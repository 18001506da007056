// csrc/rpw_fit.cuh — K3, the plane fit of the per-scan ground-segmentation path (fitPlaneAndSplit,
// RP/src/recursive_patchwork.cpp:109-308): device code shared by the two translation units that instantiate
// it, rpw_fit_fast.cu (the default path) and rpw_fit_replay.cu (the same kernels carrying the reference-order
// arithmetic of RPW_SOLVER_REFERENCE / rpw_set_exact_replay).  Two units so that the default kernels stay free
// of the replay code, and so that the two halves compile side by side.
#pragma once

#include "rpw_kernels.h"

#include <cooperative_groups.h>

#ifndef RPW_LB128
#define RPW_LB128 5  // resident blocks per SM the 128-thread fit kernels are compiled for (5 x 40 KB slots fill an SM)
#endif
#ifndef RPW_NEWTON_TOL_HYBRID
// Last Newton step of the closed-form solve under the hybrid solver.  The hybrid solver only keeps the closed form
// where the gap to the second eigenvalue is above 2 % of the matrix scale; there Newton converges quadratically, so
// a step below 1e-7 leaves an error of ~1e-12 (one step fewer than running to 1e-13; labels unchanged).
#define RPW_NEWTON_TOL_HYBRID 1e-7
#endif
#ifndef RPW_LB32
#define RPW_LB32 16  // resident blocks per SM the one-warp fit kernel (smallest patches) is compiled for
#endif
#ifndef RPW_T0
#define RPW_T0 64    // threads of the <= 1024-point class
#endif
#ifndef RPW_T1
#define RPW_T1 64    // threads of the <= 2048-point class
#endif
#ifndef RPW_LT0
#define RPW_LT0 RPW_T0  // threads of the latency table's classes <= 1024 / <= 2048 / <= 3072 and <= 4096 points
#endif
#ifndef RPW_LT1
#define RPW_LT1 RPW_T1
#endif
#ifndef RPW_LT2
#define RPW_LT2 256  // (128 -> 256 threads for the 2049..4096-point classes of single-scan calls: C2 p50 0.143 -> 0.134 ms)
#endif
#ifndef RPW_LB64
#define RPW_LB64 10  // resident blocks per SM the 64-thread fit kernels are compiled for (8 / 9 / 10 / 11 / 12: 1.830 / 1.821 / 1.814 / 1.813 / 1.811 ms per 512 C2 scans)
#endif
#ifndef RPW_STREAM_THREADS
// (256 threads, four blocks per SM: C4 1.51 -> 1.56 ms per 64 scans, C5 1.52 -> 1.54, C2 the same; 1024 threads, one block:
// C4 1.54, C5 1.50, C2 1.832 -> 1.877)
#define RPW_STREAM_THREADS 512  // block size of the class whose patches stream from L2 (no shared-memory slot)
#endif
#ifndef RPW_STREAM_TMA
// 1: the streamed class's passes read their records through a ring of TMA bulk copies (cp.async.bulk + mbarrier) instead of
// per-thread LDG.128.  Built and measured (labels bit-identical), and it LOSES: fit per 64 C4 scans 1.18 ms with plain loads,
// 1.80 ms with 3 stages x 2048 records, 2.14 ms with 6 x 1024, 2.68 ms with 8 x 512; C5 1.22 -> 1.31; C2 1.185 -> 1.196
// (profiles/r02_fit_issue_bound.md).  Compiled out.
#define RPW_STREAM_TMA 0
#endif
#ifndef RPW_TMA_ROWS
#define RPW_TMA_ROWS 4    // records per thread in one ring tile
#endif
#ifndef RPW_TMA_STAGES
#define RPW_TMA_STAGES 3  // ring depth
#endif

namespace rpw {

// =============================================================================================
// K3: fit
//   K3a rpw_fit_roots_kernel  fitPlaneAndSplit at depth 0 (RP/src/recursive_patchwork.cpp:109-308), one
//                           block per listed ring/sector patch, seven size classes (block shape and shared-
//                           memory slot per class) launched on concurrent, prioritised streams;
//                           per node: early-outs, seeds, iterated PCA plane fit with a register
//                           3x3 eigensolve, residual mask, split (variance axis, exact radix-select
//                           median, stable partition), child enqueue, label scatter
//   K3b rpw_fit_levels_kernel persistent cooperative kernel (grid barrier in global memory):
//                           level-synchronous device worklist over the children of split nodes
//                           (depth >= 1), no host round trips
// =============================================================================================
struct FitSmem {
    float* x; float* y; float* z; uint8_t* m;
    float* red;        // 2 * (TT / 32) * kRedMax floats (ping-pong)
    uint32_t* hist;    // 256
    uint32_t* misc;    // small broadcast area
    float4* ring;      // streamed class only: RPW_TMA_STAGES tiles of RPW_TMA_ROWS * RPW_STREAM_THREADS records (else nullptr)
    uint64_t* bar;     // ... and one mbarrier per tile
};

constexpr int kRedMax = 16;
// FitSmem::misc, in words: [0, 1] radix select, [2..4] heap-select replay, [8..10] plane normal, [16..24] sequential sums, [32..79] cluster exchange,
// [80] phase bits of the ring's mbarriers
constexpr int kMiscWords = 96;
constexpr int kCapTiny = 1024;   // points a 64-thread block keeps in shared memory
constexpr int kCapSmall = 4096;  // points a 128-thread block keeps in shared memory
constexpr int kCapLarge = 8192;  // largest shared-memory slot (256-thread block); larger patches stream from L2 ...
#ifndef RPW_CAP_STREAM
// (Giving the 512-thread class a resident slot -- 12288 or 16384 points, one block per SM instead of two streaming ones:
// C2 unchanged, 1.838 -> 1.835 ms per 512 scans; C4 1.50 -> 1.66 / 1.68 per 64; C5 1.52 -> 1.46 / 1.41.)
#define RPW_CAP_STREAM 256
#endif
constexpr int kCapStream = RPW_CAP_STREAM;  // ... in 512-thread blocks that keep nothing resident

// Sum of K floats over the block; every thread receives the totals (bitwise identical in all
// threads).  One __syncthreads per call; the scratch area alternates so that back-to-back calls do
// not race.  Inside a warp the K running sums are reduce-scattered (at every butterfly step a lane
// hands half of its values to its partner and keeps the other half), so K values cost about K + 4
// shuffles instead of 5 K; after the barrier every warp folds the 8 per-warp partials and
// broadcasts the totals with one shuffle each.
template <int TT, int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* red, int& phase) {
    static_assert(K <= 16, "block_sum handles at most 16 values");
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if constexpr (TT == 32 && K <= 2) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
        }
        __syncwarp();
        return;
    } else if constexpr (TT == 32) {
        // a block of one warp (the small-patch classes): the same reduce-scatter butterfly, then every lane fetches
        // the totals from the lanes that hold them; no shared memory, no barrier
        float w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k] = k < K ? v[k] : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool up = lane & 16;
            const float send = up ? w[j] : w[j + 8];
            const float keep = up ? w[j + 8] : w[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool up = lane & 8;
            const float send = up ? w[j] : w[j + 4];
            const float keep = up ? w[j + 4] : w[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const bool up = lane & 4;
            const float send = up ? w[j] : w[j + 2];
            const float keep = up ? w[j + 2] : w[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        {
            const bool up = lane & 2;
            const float send = up ? w[0] : w[1];
            const float keep = up ? w[1] : w[0];
            w[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
        // value k sits in the lanes whose bits 4..1 spell k (bit4 -> 8, bit3 -> 4, bit2 -> 2, bit1 -> 1)
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = __shfl_sync(0xffffffffu, w[0], ((k >> 3) & 1) * 16 + ((k >> 2) & 1) * 8 + ((k >> 1) & 1) * 4 + (k & 1) * 2);
        __syncwarp();
        return;
    } else {
    float* r = red + phase * ((TT / 32) * 16);
    phase ^= 1;
    if (K <= 2) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) r[warp * 16 + k] = v[k];
        }
    } else {
        float w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k] = k < K ? v[k] : 0.f;
        // 16 -> 8 -> 4 -> 2 -> 1 values per lane; lane bit (4,3,2,1) selects the half that is kept
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool up = lane & 16;
            const float send = up ? w[j] : w[j + 8];
            const float keep = up ? w[j + 8] : w[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const bool up = lane & 8;
            const float send = up ? w[j] : w[j + 4];
            const float keep = up ? w[j + 4] : w[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const bool up = lane & 4;
            const float send = up ? w[j] : w[j + 2];
            const float keep = up ? w[j + 2] : w[j];
            w[j] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        {
            const bool up = lane & 2;
            const float send = up ? w[0] : w[1];
            const float keep = up ? w[1] : w[0];
            w[0] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
        w[0] += __shfl_xor_sync(0xffffffffu, w[0], 1);
        // the value index this lane ended up with: bit4 -> 8, bit3 -> 4, bit2 -> 2, bit1 -> 1
        const int idx = ((lane >> 4) & 1) * 8 + ((lane >> 3) & 1) * 4 + ((lane >> 2) & 1) * 2 + ((lane >> 1) & 1);
        if ((lane & 1) == 0) r[warp * 16 + idx] = w[0];
    }
    __syncthreads();
    // fold the per-warp partials: lane l sums value (l & 15) over warps (l >> 4) * 4 .. + 3
    const int val = lane & 15, half = lane >> 4;
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < (TT / 32) / 2; ++q) t += r[(half * ((TT / 32) / 2) + q) * 16 + val];
    t += __shfl_xor_sync(0xffffffffu, t, 16);
#pragma unroll
    for (int k = 0; k < K; ++k) v[k] = __shfl_sync(0xffffffffu, t, k);
    }
}

template <int TT, int K>
__device__ __forceinline__ void block_min(float (&v)[K], float* red, int& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], d));
    }
    if (TT == 32) { __syncwarp(); return; }  // the xor butterfly left the minima in every lane
    float* r = red + phase * ((TT / 32) * kRedMax);
    phase ^= 1;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) r[warp * kRedMax + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = r[(lane & ((TT / 32) - 1)) * kRedMax + k];
#pragma unroll
        for (int d = (TT / 32) / 2; d > 0; d >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, d));
        v[k] = t;
    }
}

// min over the block of a 64-bit key; all threads get the result.
template <int TT>
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, float* red, int& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o < v ? o : v;
    }
    if (TT == 32) return v;
    unsigned long long* r = reinterpret_cast<unsigned long long*>(red + phase * ((TT / 32) * kRedMax));
    phase ^= 1;
    if (lane == 0) r[warp] = v;
    __syncthreads();
    unsigned long long t = r[lane & ((TT / 32) - 1)];
#pragma unroll
    for (int d = (TT / 32) / 2; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, t, d);
        t = o < t ? o : t;
    }
    return t;
}

// ---------------------------------------------------------------------------------------------
// A node spread over the CL thread blocks of a cluster (the single-scan path, where most SMs would otherwise idle and
// a call ends with its longest-iterating patch): block `rank` keeps the points [lo, lo + nl) of the node in ITS shared
// memory and runs the same control flow as every other block of the cluster; wherever a block-wide reduction decides
// something, the blocks' partial results are combined through distributed shared memory in rank order, so every block
// holds bit-identical totals and takes the same branches.  A pass then costs n / CL points per SM instead of n.
// CL == 1: a node on one block; everything below compiles away.
// ---------------------------------------------------------------------------------------------
namespace cg = cooperative_groups;

template <int CL>
struct Clu {
    unsigned rank = 0;
    uint32_t lo = 0;       // first point of this block's slice
    float* xch = nullptr;  // [2][16] this block's partial values (two exchanges in flight at most)
    float* tot = nullptr;  // [16] combined values
    int xphase = 0;
    __device__ __forceinline__ bool lead() const { return CL == 1 || rank == 0; }
};

template <int K, int CL, typename Op>
__device__ __forceinline__ void cluster_combine(float (&v)[K], Clu<CL>& cc, Op op) {
    if constexpr (CL > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        float* my = cc.xch + cc.xphase * 16;
        cc.xphase ^= 1;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < K; ++k) my[k] = v[k];
        }
        cluster.sync();
        if (threadIdx.x < K) {
            float t = cluster.map_shared_rank(my, 0)[threadIdx.x];
#pragma unroll
            for (int r = 1; r < CL; ++r) t = op(t, cluster.map_shared_rank(my, r)[threadIdx.x]);  // rank order: the same bits in every block
            cc.tot[threadIdx.x] = t;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < K; ++k) v[k] = cc.tot[k];
    }
}

template <int TT, int K, int CL>
__device__ __forceinline__ void node_sum(float (&v)[K], const FitSmem& S, int& phase, Clu<CL>& cc) {
    block_sum<TT, K>(v, S.red, phase);
    cluster_combine<K, CL>(v, cc, [](float a, float b) { return a + b; });
}
template <int TT, int K, int CL>
__device__ __forceinline__ void node_min(float (&v)[K], const FitSmem& S, int& phase, Clu<CL>& cc) {
    block_min<TT, K>(v, S.red, phase);
    cluster_combine<K, CL>(v, cc, [](float a, float b) { return fminf(a, b); });
}

// Point access: shared-memory resident (node fits) or streamed from the L2-resident segment.
template <bool SMEM>
struct NodeView {
    const float4* src;  // node's first record in its level buffer
    uint8_t* gmask;     // node's first byte of the streaming mask scratch
    FitSmem s;
    __device__ __forceinline__ void get(uint32_t i, float& x, float& y, float& z) const {
        if (SMEM) { x = s.x[i]; y = s.y[i]; z = s.z[i]; }
        else { const float4 v = __ldcg(src + i); x = v.x; y = v.y; z = v.z; }
    }
    __device__ __forceinline__ float coord(uint32_t i, int axis) const {
        if (SMEM) return axis == 0 ? s.x[i] : (axis == 1 ? s.y[i] : s.z[i]);
        const float4 v = __ldcg(src + i);
        return axis == 0 ? v.x : (axis == 1 ? v.y : v.z);
    }
    __device__ __forceinline__ uint8_t mask(uint32_t i) const { return SMEM ? s.m[i] : gmask[i]; }
    __device__ __forceinline__ void set_mask(uint32_t i, uint8_t v) const { if (SMEM) s.m[i] = v; else gmask[i] = v; }
};

// k-th smallest (0-based) of one coordinate over the node: exact 4x8-bit radix select.
// Reference: std::sort + index (RP/src/recursive_patchwork.cpp:156-159, :259-260, :267-268).
// n: the points this block holds; k: rank over the whole node (CL > 1: the histograms of the cluster's blocks are added).
template <int TT, bool SMEM, int CL = 1>
__device__ float radix_select(const NodeView<SMEM>& nv, uint32_t n, int axis, uint32_t k) {
    uint32_t* hist = nv.s.hist;
    uint32_t* misc = nv.s.misc;
    uint32_t prefix = 0, pmask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += TT) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += TT) {
            const uint32_t u = f2ord(nv.coord(i, axis));
            if ((u & pmask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
        if constexpr (CL > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            cluster.sync();  // every block's histogram is complete
            uint32_t sum[(256 + TT - 1) / TT];
#pragma unroll
            for (int j = 0; j < (256 + TT - 1) / TT; ++j) {
                const int b = threadIdx.x + j * TT;
                sum[j] = 0;
                if (b < 256)
                    for (int r = 0; r < CL; ++r) sum[j] += cluster.map_shared_rank(hist, r)[b];
            }
            cluster.sync();  // every block has read every histogram
#pragma unroll
            for (int j = 0; j < (256 + TT - 1) / TT; ++j) {
                const int b = threadIdx.x + j * TT;
                if (b < 256) hist[b] = sum[j];
            }
            __syncthreads();
        }
        if (threadIdx.x < 32) {
            // lane l owns bins [8l, 8l+8)
            uint32_t c[8], tot = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { c[j] = hist[threadIdx.x * 8 + j]; tot += c[j]; }
            uint32_t inc = tot;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int)threadIdx.x >= d) inc += t;
            }
            uint32_t before = inc - tot;
            if (k >= before && k < inc) {
                uint32_t kk = k - before;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (kk < c[j]) { misc[0] = threadIdx.x * 8 + j; misc[1] = kk; kk = 0xFFFFFFFFu; }
                    else if (kk != 0xFFFFFFFFu) kk -= c[j];
                }
            }
        }
        __syncthreads();
        prefix |= misc[0] << shift;
        pmask |= 255u << shift;
        k = misc[1];
        __syncthreads();
    }
    return ord2f(prefix);
}

__device__ __forceinline__ void dbg_record(const FitArgs& A, const NodeRef& nd, int depth, int outcome, int iters, int n_in,
                                           int axis, float cx, float cy, float cz, float nx, float ny, float nz, float res,
                                           float median, float mean_dist) {
    if (A.dbg_nodes == nullptr) return;
    const uint32_t slot = atomicAdd(A.dbg_count, 1u);
    if (slot >= A.dbg_cap) return;
    rpw_node_rec& r = A.dbg_nodes[slot];
    const uint32_t lscan = nd.root / (uint32_t)A.P;
    r.scan = (int32_t)(A.scan_base + lscan);
    r.root = (int32_t)(nd.root % (uint32_t)A.P);
    r.depth = depth;
    r.start = (int32_t)(nd.start - A.patch_start[(size_t)lscan * (A.P + 1) + r.root]);
    r.n = (int32_t)nd.n;
    r.outcome = outcome; r.iters = iters; r.n_inliers = n_in; r.split_axis = axis;
    r.centroid[0] = cx; r.centroid[1] = cy; r.centroid[2] = cz;
    r.normal[0] = nx; r.normal[1] = ny; r.normal[2] = nz;
    r.residual = res; r.median = median; r.mean_dist = mean_dist;
}

// Labels of a whole node set to one value (early-outs; RP/src/recursive_patchwork.cpp:111-113,
// :126-129, :138-140).  Slot j of the node labels input point sortedA[start + j].w — the
// positional read-back of SURVEY Q1.
// Input index of patch slot `slot` (positional read-back).
__device__ __forceinline__ uint32_t slot_input_index(const FitArgs& A, uint32_t slot) {
    return __float_as_uint(__ldcg(&A.sortedA[slot].w));
}

template <int TT>
__device__ __forceinline__ void label_const(const FitArgs& A, uint32_t start, uint32_t n, uint8_t v) {
    for (uint32_t i = threadIdx.x; i < n; i += TT) A.labels[slot_input_index(A, start + i)] = v;
}
template <int TT>
__device__ __forceinline__ void label_const(const FitArgs& A, const NodeRef& nd, uint8_t v) { label_const<TT>(A, nd.start, nd.n, v); }

// Strided loop over a node's points, four rows per trip: the loads of a trip are issued together
// before any of its arithmetic (memory-level parallelism instead of one dependent chain per row).
// body(i, x, y, z, m) sees point i with its current mask byte.  (A layout where a thread owns four
// CONSECUTIVE points and loads them with three LDS.128 needs a quarter of the load instructions but
// measured 6 % slower end to end: the passes are bound by the dependent latency of a thread's own
// instruction stream at the low occupancy shared memory allows, not by issue slots.)
#ifndef RPW_UNROLL
#define RPW_UNROLL 4
#endif
#ifndef RPW_WIDE
#define RPW_WIDE 8
#endif
#ifndef RPW_UNROLL_STREAM
// rows per trip of a STREAMED node's passes (records from L2; the 512-thread kernel has 64 registers per thread): two.
// Per 64 C4 scans 1.81 / 1.42 / 1.69 / 1.50 / 1.73 ms with 1 / 2 / 3 / 4 / 8; C5 1.54 / 1.49 / 1.48 / 1.52 / 1.52; C2 unchanged.
#define RPW_UNROLL_STREAM 2
#endif

// ---- streamed nodes: records through a ring of TMA bulk copies --------------------------------------------------------
// A node too large for a shared-memory slot is re-read from L2 on every pass.  With plain loads each thread has
// kUnroll 16-byte requests in flight and waits for them before its arithmetic starts.  Here one thread keeps
// RPW_TMA_STAGES - 1 tiles of RPW_TMA_ROWS * TT records in flight ahead of the tile being consumed (cp.async.bulk,
// completion counted in bytes on the tile's mbarrier; SASS: UBLKCP + SYNCS), and the block reads a landed tile with one
// LDS.128 per record.  Point i is still handled by thread i mod TT in ascending order, so every sum is bit-identical to the
// plain-load path.  The mask bytes stay in global memory (a node's mask range is not 16-byte aligned, which a bulk copy
// needs); their loads are issued before the wait on the tile.
// Why it loses: a pass is bound by instruction issue, not by load latency (two resident 512-thread blocks already keep
// 32 warps x 4 LDG.128 in flight per SM), and LDS.128 replaces LDG.128 one for one, so no instruction is saved; what the
// ring adds is one block barrier and one mbarrier wait per tile, which put all 16 warps in lockstep with the slowest one
// where the plain loop lets them run free -- the smaller the tile, the more often.
constexpr int kRingTile = RPW_TMA_ROWS * RPW_STREAM_THREADS;
constexpr size_t kRingBytes = RPW_STREAM_TMA ? (size_t)RPW_TMA_STAGES * kRingTile * 16 + 128 + 8 * RPW_TMA_STAGES : 0;
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ring_init(const FitSmem& S) {
    if (S.ring == nullptr) return;
    if (threadIdx.x == 0) {
        for (int s = 0; s < RPW_TMA_STAGES; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_addr(S.bar + s)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        S.misc[80] = 0;
    }
    __syncthreads();
}
template <int TT, bool WITH_MASK, typename F>
__device__ __forceinline__ void for_points_ring(const float4* src, const uint8_t* gmask, const FitSmem& S, uint32_t n, F body) {
    constexpr int R = RPW_TMA_ROWS, NS = RPW_TMA_STAGES;
    constexpr uint32_t TILE = R * TT;
    const uint32_t tiles = (n + TILE - 1) / TILE;
    const uint32_t ring0 = smem_addr(S.ring), bar0 = smem_addr(S.bar);
    __syncthreads();  // the previous call's phase word is written, its last tile read
    uint32_t ph = S.misc[80];
    auto issue = [&](uint32_t t) {
        const uint32_t s = t % NS, rows = min(TILE, n - t * TILE), bar = bar0 + 8 * s;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(rows * 16u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ring0 + s * TILE * 16u), "l"(src + (size_t)t * TILE), "r"(rows * 16u), "r"(bar) : "memory");
    };
    if (threadIdx.x == 0)
        for (uint32_t t = 0; t < tiles && t < (uint32_t)NS; ++t) issue(t);
    uint32_t s = 0;
    for (uint32_t t = 0; t < tiles; ++t) {
        const uint32_t base = t * TILE + threadIdx.x;
        const bool full = (t + 1) * TILE <= n;
        uint8_t m[R];
#pragma unroll
        for (int u = 0; u < R; ++u) m[u] = (WITH_MASK && (full || base + u * TT < n)) ? gmask[base + u * TT] : (uint8_t)1;
        const uint32_t bar = bar0 + 8 * s, parity = (ph >> s) & 1u;
        uint32_t ok;
        do {
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                         : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        } while (!ok);
        ph ^= 1u << s;
        const float4* tile = S.ring + s * TILE + threadIdx.x;
        float4 v[R];
#pragma unroll
        for (int u = 0; u < R; ++u) v[u] = tile[u * TT];
        if (full) {
#pragma unroll
            for (int u = 0; u < R; ++u) body(base + u * TT, v[u].x, v[u].y, v[u].z, m[u]);
        } else {
#pragma unroll
            for (int u = 0; u < R; ++u)
                if (base + u * TT < n) body(base + u * TT, v[u].x, v[u].y, v[u].z, m[u]);
        }
        __syncthreads();  // the tile is consumed: its stage may be refilled
        if (threadIdx.x == 0 && t + NS < tiles) issue(t + NS);
        s = (s + 1 == (uint32_t)NS) ? 0u : s + 1;
    }
    if (threadIdx.x == 0) S.misc[80] = ph;
}

template <int TT, bool SMEM, bool WITH_MASK, typename F>
__device__ __forceinline__ void for_points(const NodeView<SMEM>& nv, uint32_t n, uint32_t first, F body) {
#if RPW_STREAM_TMA
    if constexpr (!SMEM && TT == RPW_STREAM_THREADS) {
        if (first == 0 && nv.s.ring != nullptr) {  // (block-uniform)
            for_points_ring<TT, WITH_MASK>(nv.src, nv.gmask, nv.s, n, body);
            return;
        }
    }
#endif
    constexpr int kUnroll = SMEM ? RPW_UNROLL : RPW_UNROLL_STREAM;
    uint32_t i = first + threadIdx.x;
    for (; i + (kUnroll - 1) * TT < n; i += kUnroll * TT) {
        float x[kUnroll], y[kUnroll], z[kUnroll];
        uint8_t m[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            nv.get(i + u * TT, x[u], y[u], z[u]);
            m[u] = WITH_MASK ? nv.mask(i + u * TT) : (uint8_t)1;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) body(i + u * TT, x[u], y[u], z[u], m[u]);
    }
    for (; i < n; i += TT) {
        float x, y, z;
        nv.get(i, x, y, z);
        body(i, x, y, z, WITH_MASK ? nv.mask(i) : (uint8_t)1);
    }
}
template <int TT, bool SMEM, bool WITH_MASK, typename F>
__device__ __forceinline__ void for_points(const NodeView<SMEM>& nv, uint32_t n, F body) {
    for_points<TT, SMEM, WITH_MASK>(nv, n, 0u, body);
}

// (Experiment, measured and not adopted, code removed: the distance / mask / moments pass on Blackwell's packed FP32
// instructions -- FADD2, FMUL2, FFMA2, two IEEE single-precision operations per instruction, each rounded exactly like the
// scalar form; a thread takes its rows two at a time, one per half of every register pair.  15 floating-point instructions
// per point instead of 24, and 0.7 % SLOWER end to end, 1.897 against 1.884 ms per 512 scans, labels identical: the packed
// instructions occupy the FMA pipe for two cycles.  One trap for whoever retries it: ptxas 12.9 contracts mul.rn.f32x2
// followed by add.rn.f32x2 into FFMA2 even under -fmad=false, which rounds the distance differently from the reference;
// packed products feeding SCALAR additions are left alone.)

// Optional cycle accounting (rpw_debug_fit_timing): thread 0 of every block adds the cycles it spent
// in each section of process_node to A.timing[section].  Sections: 0 load+bbox, 1 seeds, 2 covariance
// pass + reduce, 3 eigensolve + broadcast, 4 distance/mask pass + reduce, 5 final fit, 6 leaf label
// write, 7 split, 8 fetch / between nodes, 9 grid barrier, 10 nodes, 11 plane-fit iterations.
struct Tick {
    unsigned long long* t;
    long long last;
    __device__ __forceinline__ Tick(unsigned long long* timing) : t(timing), last(0) { if (t && threadIdx.x == 0) last = clock64(); }
    __device__ __forceinline__ void operator()(int slot) {
        if (t && threadIdx.x == 0) { const long long now = clock64(); atomicAdd(t + slot, (unsigned long long)(now - last)); last = now; }
    }
    __device__ __forceinline__ void count(int slot, unsigned v) { if (t && threadIdx.x == 0) atomicAdd(t + slot, (unsigned long long)v); }
};

// Plane normal from the scatter sums of the current inliers (fitPlanePCA, :86-95): smallest-
// eigenvalue eigenvector, flipped to z >= 0.  Computed by warp 0, broadcast through shared memory.
// (Which warp runs the solve makes no measurable difference: first, last, or spread round-robin over
// the SM's sub-partitions all give the same throughput.)
template <bool EXACT, int TT>
__device__ __forceinline__ void plane_normal(const float (&cv)[6], float cnt, float* bc, float& nx, float& ny, float& nz,
                                             bool hybrid, unsigned long long* timing = nullptr) {
    if (TT == 32 || threadIdx.x < 32) {
        float ax, ay, az;
        if (EXACT) {
            long long t0 = 0;
            if (timing && threadIdx.x == 0) t0 = clock64();
            plane_normal_exact(cv, cnt - 1.f, ax, ay, az);  // computeCovariance divides by n-1 (point_cloud_processor.cpp:84)
            if (timing && threadIdx.x == 0) atomicAdd(timing + 14, (unsigned long long)(clock64() - t0));
        } else {
            bool small_gap;
            smallest_eigvec_psd(cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], ax, ay, az, &small_gap, hybrid ? RPW_NEWTON_TOL_HYBRID : 1e-13);
            // hybrid solver: where the eigenvector is ill-conditioned only the reference's own operation
            // sequence reproduces the reference's answer
            if (hybrid && small_gap) {
                // out of line: taken by a few percent of the solves, and ~1100 instructions that would otherwise sit in
                // the middle of the plane-fit loop (fit phase -2 %)
                const Normal3 r = plane_normal_exact_cold(cv[0], cv[1], cv[2], cv[3], cv[4], cv[5], cnt - 1.f);
                ax = r.x; ay = r.y; az = r.z;
                if (timing && threadIdx.x == 0) atomicAdd(timing + 14, 1ull);  // (hybrid: slot 14 counts the QR solves)
            }
        }
        if (az < 0.f) { ax = -ax; ay = -ay; az = -az; }  // :93-95
        if (TT == 32) { nx = ax; ny = ay; nz = az; return; }  // a block of one warp: every lane holds the result
        if (threadIdx.x == 0) { bc[0] = ax; bc[1] = ay; bc[2] = az; }
    }
    __syncthreads();
    nx = bc[0]; ny = bc[1]; nz = bc[2];
}

// Timeline record of one node (thread 0; debugging aid, off unless rpw_debug_fit_trace armed it).
struct TraceScope {
    const FitArgs& A; uint64_t t0; uint32_t n; uint16_t depth, cls;
    __device__ __forceinline__ static uint64_t now() { uint64_t t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
    __device__ __forceinline__ TraceScope(const FitArgs& a, uint32_t n_, int depth_, int cls_) : A(a), t0(0), n(n_), depth((uint16_t)depth_), cls((uint16_t)cls_) {
        if (A.trace && threadIdx.x == 0) t0 = now();
    }
    __device__ __forceinline__ void done(int iters) {
        if (A.trace && threadIdx.x == 0) {
            const uint32_t slot = atomicAdd(A.trace_count, 1u);
            if (slot < A.trace_cap) {
                uint32_t smid;
                asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
                rpw_trace_rec r;
                r.t_start_ns = t0; r.t_end_ns = now(); r.sm = smid; r.n = n; r.depth = depth; r.size_class = cls; r.iters = (uint32_t)iters;
                A.trace[slot] = r;
            }
        }
    }
};

// =============================================================================================
// Reference-order arithmetic (rpw_set_exact_replay / RPW_SOLVER_REFERENCE).
//
// The reference adds floats one after the other -- computeCentroid and computeCovariance
// (RP/src/point_cloud_processor.cpp:58-86), the residual (RP/src/recursive_patchwork.cpp:98-104), the root patch's
// mean range (:383-387), the split statistics (:240-249) -- and a float sum depends on its order.  The fast path
// above sums in trees; the two differ in the last bits of every moment, which a patch whose fit sits between two
// fixed points (or creeps until max_iter) amplifies into different masks.  Here the sums are taken in the
// reference's order: NV running sums are NV dependent chains of FADDs, one lane each; the warp's 32 lanes first
// compute the addends of 32 consecutive points side by side and hand them over through shared memory, so the chain
// lanes only load and add (four to five cycles per point, the latency of a dependent FADD, against ~0.5 for the tree).  Masked-out points contribute
// +0.0f, which leaves a sum's bits unchanged (a sum that starts at +0 never becomes -0).  Everything else of a plane
// fit is element-wise and already the reference's operations.  With these sums, the QR eigensolver and the exact
// medians / percentiles, every decision of fitPlaneAndSplit is taken on the reference's bits.
// =============================================================================================
#ifndef RPW_LB512_REPLAY
#define RPW_LB512_REPLAY 1
#endif
#ifndef RPW_LB256_REPLAY
#define RPW_LB256_REPLAY 3
#endif
#ifndef RPW_LB128_REPLAY
#define RPW_LB128_REPLAY RPW_LB128
#endif
#ifndef RPW_SEQ_AHEAD
#define RPW_SEQ_AHEAD 4  // chunks a streamed node's sequential sums keep in flight
#endif
// (Tried on top of this, labels identical each time, per 512 C2 / 64 C4 / 64 C5 scans in the reference-order mode, from 14.8 / 36.2 /
// 18.3 ms: every class above 2048 points as 64-thread STREAMING blocks, eight chains to an SM instead of the 2-5 a shared-memory
// slot allows: 15.9 / 37.0 / 18.0 -- the batches end with single long nodes, residency is not what limits them.  A producer
// warp filling two row buffers and a consumer warp that only loads and adds, handing buffers over with named barriers: 17.2 /
// 44.7 / 20.9, two barrier round trips per 32-point chunk cost more than the work they move out of the chain's warp.  The
// chain written ahead of the next chunk's addend computation in program order: 17.3 / 40.4 / 19.3.)
constexpr int kSeqStride = 36;  // floats between the addend rows of two sums (36: the chain lanes' 16-byte loads spread over the banks)

// Sequential sums over the node's points i = 0 .. n-1, in order.  produce(i, x, y, z, m, v) fills the NV addends of
// point i.  scratch: NV * kSeqStride floats (FitSmem::hist).  Every thread of the block receives the sums.
template <int TT, bool SMEM, int NV, typename F>
__device__ __forceinline__ void seq_sums(const NodeView<SMEM>& nv, uint32_t n, float* scratch, float* bc, float (&out)[NV], F produce) {
    static_assert(NV * kSeqStride <= 256, "the addend rows live in the 256-word histogram area");
    const int lane = threadIdx.x & 31;
    if (TT > 32) __syncthreads();  // scratch and bc may still be read by the previous user
    if (TT == 32 || threadIdx.x < 32) {
        // Software pipeline over chunks of 32 points, one basic block per chunk so that the compiler can fill the chain's
        // latency (32 dependent FADDs, 4 cycles each) with the other chunks' work: while chunk c is added, chunk c+1's
        // addends are computed and stored and chunk c+2's points are loaded.  Every lane runs the chain (lanes >= NV repeat
        // the last row and are ignored): no divergence.  Chunks past the end store zeros, which are never read.
        float acc = 0.f;
        const float4* rowp = reinterpret_cast<const float4*>(scratch + (lane < NV ? lane : NV - 1) * kSeqStride);
        float x, y, z;
        uint8_t m;
        auto fetch = [&](uint32_t base) {
            const uint32_t i = min(base + (uint32_t)lane, n - 1u);
            nv.get(i, x, y, z);
            m = nv.mask(i);
        };
        auto put = [&](uint32_t base) {
            float v[NV];
            produce(base + lane, x, y, z, m, v);
            const bool live = base + lane < n;
#pragma unroll
            for (int k = 0; k < NV; ++k) scratch[k * kSeqStride + lane] = live ? v[k] : 0.f;
        };
        if constexpr (!SMEM) {
            // Streamed node: a point comes from L2 (~700 cycles), so one chunk of look-ahead leaves the chain waiting for its
            // addends (measured: 20 cycles per point against 10 for a resident node).  kAhead chunks are kept in flight in a
            // register ring; the loop is unrolled over the ring so that its slots are compile-time registers.  Groups of
            // kAhead chunks: the chunks past the end add +0.0f, which leaves the sums' bits unchanged.
            constexpr int kAhead = RPW_SEQ_AHEAD;
            float fx[kAhead], fy[kAhead], fz[kAhead];
            uint8_t fm[kAhead];
            auto fetch_slot = [&](int k, uint32_t base) {
                const uint32_t i = min(base + (uint32_t)lane, n - 1u);
                nv.get(i, fx[k], fy[k], fz[k]);
                fm[k] = nv.mask(i);
            };
            auto put_slot = [&](int k, uint32_t base) {
                float v[NV];
                produce(base + lane, fx[k], fy[k], fz[k], fm[k], v);
                const bool live = base + lane < n;
#pragma unroll
                for (int q = 0; q < NV; ++q) scratch[q * kSeqStride + lane] = live ? v[q] : 0.f;
            };
#pragma unroll
            for (int k = 0; k < kAhead; ++k) fetch_slot(k, 32u * k);
            put_slot(0, 0);  // rows = chunk 0; slots 1 .. kAhead-1 hold the chunks after it, slot 0 is free
            __syncwarp();
            for (uint32_t base = 0; base < n; base += 32 * kAhead) {
#pragma unroll
                for (int u = 0; u < kAhead; ++u) {
                    const uint32_t cb = base + 32u * u;  // the chunk whose addends are in the rows
                    float4 t[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) t[q] = rowp[q];
                    __syncwarp();
                    fetch_slot(u, cb + 32u * kAhead);               // slot u was consumed when this chunk's rows were written
                    put_slot((u + 1) % kAhead, cb + 32u);          // the next chunk's addends
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        acc = __fadd_rn(acc, t[q].x); acc = __fadd_rn(acc, t[q].y); acc = __fadd_rn(acc, t[q].z); acc = __fadd_rn(acc, t[q].w);
                    }
                    __syncwarp();
                }
            }
        } else {
        fetch(0);
        put(0);
        fetch(32);
        __syncwarp();
        for (uint32_t base = 0; base < n; base += 32) {
            float4 t[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) t[q] = rowp[q];
            __syncwarp();       // every lane holds its row of chunk `base`: the rows can be overwritten
            put(base + 32);     // addends of the next chunk (its points arrived during the previous chain)
            fetch(base + 64);   // points of the chunk after that, in flight during this chain
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc = __fadd_rn(acc, t[q].x); acc = __fadd_rn(acc, t[q].y); acc = __fadd_rn(acc, t[q].z); acc = __fadd_rn(acc, t[q].w);
            }
            __syncwarp();       // the next chunk's rows are complete
        }
        }
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = __shfl_sync(0xffffffffu, acc, k);
        if (TT > 32 && lane == 0) {
#pragma unroll
            for (int k = 0; k < NV; ++k) bc[k] = out[k];
        }
    }
    if (TT > 32) {
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NV; ++k) out[k] = bc[k];
    }
}

struct FitState {
    float cx, cy, cz, nx, ny, nz, residual, cnt;
    int iters;
};

// fitPlanePCA (RP/src/recursive_patchwork.cpp:77-95) of the points whose mask byte is set, reference order.
template <int TT, bool SMEM>
__device__ __forceinline__ void plane_of_mask_seq(const NodeView<SMEM>& nv, uint32_t n, float cnt, FitSmem S, float& cx, float& cy, float& cz,
                                                  float& nx, float& ny, float& nz) {
    float* bc = reinterpret_cast<float*>(S.misc + 8);
    float* sbc = reinterpret_cast<float*>(S.misc + 16);
    float s3[3];
    seq_sums<TT, SMEM, 3>(nv, n, reinterpret_cast<float*>(S.hist), sbc, s3, [](uint32_t, float x, float y, float z, uint8_t m, float (&v)[3]) {
        v[0] = (m & 1) ? x : 0.f; v[1] = (m & 1) ? y : 0.f; v[2] = (m & 1) ? z : 0.f;  // (bit 0: the current mask; exact_refit keeps older ones above it)
    });
    cx = s3[0] / cnt; cy = s3[1] / cnt; cz = s3[2] / cnt;  // centroid /= points.size()
    const float ccx = cx, ccy = cy, ccz = cz;
    float cv[6];  // xx yx yy zx zy zz (diff * diff^T is symmetric bit for bit: float products commute)
    seq_sums<TT, SMEM, 6>(nv, n, reinterpret_cast<float*>(S.hist), sbc, cv, [=](uint32_t, float x, float y, float z, uint8_t m, float (&v)[6]) {
        const float d0 = x - ccx, d1 = y - ccy, d2 = z - ccz;
        const bool in = m & 1;
        v[0] = in ? d0 * d0 : 0.f; v[1] = in ? d1 * d0 : 0.f; v[2] = in ? d1 * d1 : 0.f;
        v[3] = in ? d2 * d0 : 0.f; v[4] = in ? d2 * d1 : 0.f; v[5] = in ? d2 * d2 : 0.f;
    });
    plane_normal<true, TT>(cv, cnt, bc, nx, ny, nz, false);  // cov /= (size - 1), Eigen's QR sequence, z-up flip
}

// The iterated fit of fitPlaneAndSplit (:185-228) from the seed mask, in the reference's arithmetic order.
// c0..c2: the three lowest points when fewer than three seeds lie below z_th (ascending index), else unused.
template <int TT, bool SMEM>
static __device__ __noinline__ void exact_refit(const NodeView<SMEM> nv, const uint32_t n, const float z_th, const float tau, const int max_iter,
                                                const bool seeds_by_height, const uint32_t c0, const uint32_t c1, const uint32_t c2,
                                                FitSmem S, FitState* out) {
    const int tid = threadIdx.x;
    int phase = 0;
    float cntv[1] = {0.f};
    for (uint32_t i = tid; i < n; i += TT) {
        const bool m = seeds_by_height ? nv.coord(i, 2) < z_th : (i == c0 || i == c1 || i == c2);
        nv.set_mask(i, m ? 1 : 0);
        cntv[0] += m ? 1.f : 0.f;
    }
    block_sum<TT, 1>(cntv, S.red, phase);
    float cnt = cntv[0];
    float cx = 0.f, cy = 0.f, cz = 0.f, nx = 0.f, ny = 0.f, nz = 1.f;
    bool plane_is_final = false;
    int iters = 0;
    // In this arithmetic a fit is a pure function of the mask (centroid and covariance are recomputed from the inliers every
    // time, as in the reference), so the sequence of masks M_0, M_1 = F(M_0), ... is determined by M_0 alone: once a mask
    // repeats, the rest of the max_iter iterations is known.  The mask byte keeps the last four masks above the current one
    // (bit j = the mask j iterations ago); a new mask equal to the one p iterations back (p = 2 .. 4) closes a cycle of period
    // p, and the mask the reference is left with after max_iter iterations is read out of the history instead of being
    // iterated to.  (Bistable patches: the 100-iteration tails of this mode.)
    for (int iter = 0; iter < max_iter; ++iter) {
        if (cnt < 3.f) break;  // :196
        plane_of_mask_seq<TT, SMEM>(nv, n, cnt, S, cx, cy, cz, nx, ny, nz);
        iters++;
        float st[5] = {0.f, 0.f, 0.f, 0.f, 0.f};  // new count; differs from the current mask, from the masks 1, 2, 3 iterations before it
        for (uint32_t i = tid; i < n; i += TT) {
            float x, y, z;
            nv.get(i, x, y, z);
            const uint32_t b = nv.mask(i);
            const uint32_t nm = plane_dist(x, y, z, cx, cy, cz, nx, ny, nz) < tau ? 1u : 0u;
            st[0] += nm ? 1.f : 0.f;
            st[1] += (nm != (b & 1u)) ? 1.f : 0.f;
            st[2] += (nm != ((b >> 1) & 1u)) ? 1.f : 0.f;
            st[3] += (nm != ((b >> 2) & 1u)) ? 1.f : 0.f;
            st[4] += (nm != ((b >> 3) & 1u)) ? 1.f : 0.f;
            nv.set_mask(i, (uint8_t)(nm | ((b << 1) & 0x1Eu)));  // in place: a point's test does not read other masks
        }
        block_sum<TT, 5>(st, S.red, phase);
        // (the history now reads: bit 0 = M_{iter+1}, bit 1 = M_iter, bit 2 = M_{iter-1}, ...)
        if (st[1] == 0.f) { plane_is_final = true; break; }  // :215
        cnt = st[0];
        int period = 0;
        if (st[2] == 0.f && iter >= 1) period = 2;
        else if (st[3] == 0.f && iter >= 2) period = 3;
        else if (st[4] == 0.f && iter >= 3) period = 4;
        if (period && iter + 1 < max_iter) {
            // M_{iter+1} == M_{iter+1-period}: from there on the masks repeat with that period.  The reference stops after
            // max_iter iterations holding M_{max_iter} = the mask at position (max_iter - (iter + 1)) mod period of the cycle
            // that starts at M_{iter+1-period}, which is bit (period - that position) of the history (position 0: bit 0).
            const int pos = (max_iter - (iter + 1)) % period;
            const int bit = pos == 0 ? 0 : period - pos;
            float c2v[1] = {0.f};
            for (uint32_t i = tid; i < n; i += TT) {
                const uint32_t b = nv.mask(i);
                const uint32_t fm = (b >> bit) & 1u;
                nv.set_mask(i, (uint8_t)fm);
                c2v[0] += fm ? 1.f : 0.f;
            }
            block_sum<TT, 1>(c2v, S.red, phase);
            cnt = c2v[0];
            iters = max_iter;  // what the reference's loop counts
            break;
        }
    }
    // the history bits go: from here on the byte is the mask (the leaf's labels are copied from it)
    for (uint32_t i = tid; i < n; i += TT) nv.set_mask(i, nv.mask(i) & 1u);
    if (TT > 32) __syncthreads();
    float residual = FLT_MAX;
    if (cnt >= 3.f) {  // :220-228, fitPlanePCA on the final mask
        if (!plane_is_final) plane_of_mask_seq<TT, SMEM>(nv, n, cnt, S, cx, cy, cz, nx, ny, nz);
        float rs[1];
        const float fcx = cx, fcy = cy, fcz = cz, fnx = nx, fny = ny, fnz = nz;
        seq_sums<TT, SMEM, 1>(nv, n, reinterpret_cast<float*>(S.hist), reinterpret_cast<float*>(S.misc + 16), rs,
                              [=](uint32_t, float x, float y, float z, uint8_t m, float (&v)[1]) {
                                  v[0] = (m & 1) ? plane_dist(x, y, z, fcx, fcy, fcz, fnx, fny, fnz) : 0.f;
                              });
        residual = rs[0] / cnt;
    } else {
        cx = cy = cz = 0.f; nx = ny = 0.f; nz = 1.f;  // :78-80
    }
    if (TT > 32) __syncthreads();
    out->cx = cx; out->cy = cy; out->cz = cz; out->nx = nx; out->ny = ny; out->nz = nz;
    out->residual = residual; out->cnt = cnt; out->iters = iters;
}

// CL > 1: the node is spread over the CL blocks of a cluster (see Clu); this block holds the points [lo, lo + nl).
// n is the node's size (what the reference's formulas see), nl what this block loops over.
// (Moving the cold paths out of line -- the seed fallback, the split, the radix select and the node record as __noinline__
// functions, to shorten the instruction stream the resident patches' warps fetch -- made the fit 11 % SLOWER, 1.34 against
// 1.20 ms per 512 scans: the node view and the shared-memory carve-out they take by reference then live on the stack, and
// the hot loop pays for it.  Only the QR fallback, which takes scalars, is out of line.)
template <int TT, bool SMEM, bool EXACT, bool REPLAY, int CL = 1>
__device__ int process_node(const FitArgs& A, const NodeRef nd, const int depth, FitSmem S, Clu<CL> cc = Clu<CL>()) {
    static_assert(CL == 1 || (SMEM && !REPLAY), "cluster nodes are shared-memory resident and take the default arithmetic");
    const FitParams& fp = A.fp;
    const uint32_t n = nd.n;
    const int tid = threadIdx.x;
    int phase = 0;
    Tick tick(A.timing);
    uint32_t lo = 0, nl = n;
    if constexpr (CL > 1) {
        const uint32_t per = ((n + CL - 1) / CL + 3u) & ~3u;
        lo = min(n, cc.rank * per);
        nl = min(n - lo, per);
        cc.lo = lo;
    }
    const bool lead = cc.lead();  // the block that speaks for the node (records, queues, per-node stores)

    if (n < 3 || depth > fp.max_split_depth) {  // :111-113
        label_const<TT>(A, nd.start + lo, nl, 0);
        if (tid == 0 && lead) dbg_record(A, nd, depth, RPW_NODE_SMALL, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, 0);
        return 0;
    }
    NodeView<SMEM> nv;
    nv.s = S;
    nv.src = (depth == 0 ? A.sortedA : ((depth & 1) ? A.bufB : A.bufC)) + nd.start + lo;
    nv.gmask = A.gmask + nd.start + lo;

    // ---- pass 1: load, bounding box, (root only) mean range ------------------------------
    float mm[6] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};  // min x,y,z, min -x,-y,-z
    float sd[2] = {0.f, 0.f};  // sum of ranges; [1]: a square root left the branch-free sequence's range
    {
        // sqrtf carries a branch to its slow path, which would keep the rows' 40-cycle chains from
        // overlapping; the branch-free copy (ArithSpec, bit-identical in range) is checked once at the end
        ArithSpec ar;
        auto take = [&](uint32_t i, const float4 v) {
            if (SMEM) { S.x[i] = v.x; S.y[i] = v.y; S.z[i] = v.z; }
            mm[0] = fminf(mm[0], v.x); mm[1] = fminf(mm[1], v.y); mm[2] = fminf(mm[2], v.z);
            mm[3] = fminf(mm[3], -v.x); mm[4] = fminf(mm[4], -v.y); mm[5] = fminf(mm[5], -v.z);
            if (depth == 0) sd[0] += ar.sqrt(v.x * v.x + v.y * v.y);  // range2d
        };
        // many independent 16-byte loads in flight per thread: the pass is DRAM/L2-latency bound, every trip costs a
        // full memory round trip.  Eight per thread: sixteen were better while the QR fallback sat inside the plane-fit
        // loop, and are 2 % worse since it moved out of line (fit 1.24 -> 1.20 ms per 512 scans with eight; four: 1.22)
        constexpr int kWide = TT <= 256 ? RPW_WIDE : 8;
        // (.ca loads for level 0, so that the leaf's label write would find the input indices in L1: no gain)
        auto ld_rec = [&](const float4* p) { return __ldcg(p); };
        // (A bare predicated copy loop followed by a short second loop over shared memory for the bounding box and the
        // range sum -- 550 instructions less code -- measured the same: 1.846 against 1.838 ms per 512 scans.)
        uint32_t i = tid;
        for (; i + (kWide - 1) * TT < nl; i += kWide * TT) {
            float4 v[kWide];
#pragma unroll
            for (int u = 0; u < kWide; ++u) v[u] = ld_rec(nv.src + i + u * TT);
#pragma unroll
            for (int u = 0; u < kWide; ++u) take(i + u * TT, v[u]);
        }
        for (; i + 3 * TT < nl; i += 4 * TT) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = ld_rec(nv.src + i + u * TT);
#pragma unroll
            for (int u = 0; u < 4; ++u) take(i + u * TT, v[u]);
        }
        for (; i < nl; i += TT) take(i, ld_rec(nv.src + i));
        sd[1] = ar.ok() ? 0.f : 1.f;
    }
    tick(12);
    node_min<TT, 6, CL>(mm, S, phase, cc);
    float mean_dist;
    // < 0: off; 0: every fit in the reference's arithmetic order; K > 0: fits of more than K iterations.  A template
    // parameter so that the kernels of the default path do not carry the reference-order code (it cost them 1.8 %)
    const int replay = REPLAY ? fp.exact_replay : -1;
    if (depth == 0) {
        node_sum<TT, 2, CL>(sd, S, phase, cc);
        if (sd[1] != 0.f) {  // (never for patch points, whose range is at least 1 m: kept for safety)
            sd[0] = 0.f;
            for (uint32_t i = tid; i < nl; i += TT) sd[0] += range2d(nv.coord(i, 0), nv.coord(i, 1));
            node_sum<TT, 2, CL>(sd, S, phase, cc);
        }
        if (replay >= 0) {  // the reference's running sum (:383-387), so that z_th and the distance threshold carry its bits
            float sr[1];
            seq_sums<TT, SMEM, 1>(nv, n, reinterpret_cast<float*>(S.hist), reinterpret_cast<float*>(S.misc + 16), sr,
                                  [](uint32_t, float x, float y, float, uint8_t, float (&v)[1]) { v[0] = range2d(x, y); });
            sd[0] = sr[0];
        }
        mean_dist = sd[0] / (float)n;  // :383-387
        if (tid == 0 && lead) A.root_mean[nd.root] = mean_dist;
    } else {
        mean_dist = __ldcg(A.root_mean + nd.root);  // Q4: inherited unchanged
    }
    tick(0);
    const float x_min = mm[0], x_max = -mm[3], y_min = mm[1], y_max = -mm[4], z_min = mm[2], z_max = -mm[5];
    const float area = (x_max - x_min) * (y_max - y_min);
    if (area < 25.0f && depth > 0) {  // :126-129
        label_const<TT>(A, nd.start + lo, nl, 1);
        if (tid == 0 && lead) dbg_record(A, nd, depth, RPW_NODE_AREA, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, mean_dist);
        return 0;
    }
    if ((z_max - z_min) < 0.05f && n > 10) {  // :138-140
        label_const<TT>(A, nd.start + lo, nl, 1);
        if (tid == 0 && lead) dbg_record(A, nd, depth, RPW_NODE_FLAT, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, mean_dist);
        return 0;
    }
    // (the barrier inside block_min already made every thread's shared-memory stores visible)

    // ---- seed threshold (:149-160) ----------------------------------------------------------
    const float rel_dist = mean_dist / fp.radius;
    float z_th;
    if (fp.adaptive_seed_height) {
        z_th = fp.sensor_height + 0.2f * rel_dist;
    } else {
        const uint32_t idx = (uint32_t)(0.1f * (float)n);
        z_th = radix_select<TT, SMEM, CL>(nv, nl, 2, idx) + fp.th_seeds;
    }
    const float tau = fp.th_dist * (1.0f + 0.2f * rel_dist);  // :203

    // ---- seeds (:163-182) -------------------------------------------------------------------
    // One pass builds the seed mask AND its first/second moments about a fixed pivot (bounding-box
    // centre, lowest z): the seed centroid is pivot + s/n and its scatter S' - s s^T / n, so the first
    // plane fit needs no separate covariance pass.  The subtraction is harmless while the pivot sits
    // inside the seed cloud; a compact seed set far from the pivot (S' >> scatter) falls back to the
    // covariance pass about the centroid.
    const float px = 0.5f * (x_min + x_max), py = 0.5f * (y_min + y_max), pz = z_min;
    float acc[10] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // count, s, S' (xx yx yy zx zy zz)
    for_points<TT, SMEM, false>(nv, nl, [&](uint32_t i, float x, float y, float z, uint8_t) {
        const bool m = z < z_th;
        nv.set_mask(i, m ? 1 : 0);
        const float dx = m ? x - px : 0.f, dy = m ? y - py : 0.f, dz = m ? z - pz : 0.f;
        acc[0] += m ? 1.f : 0.f; acc[1] += dx; acc[2] += dy; acc[3] += dz;
        acc[4] = fmaf(dx, dx, acc[4]); acc[5] = fmaf(dy, dx, acc[5]); acc[6] = fmaf(dy, dy, acc[6]);
        acc[7] = fmaf(dz, dx, acc[7]); acc[8] = fmaf(dz, dy, acc[8]); acc[9] = fmaf(dz, dz, acc[9]);
    });
    node_sum<TT, 10, CL>(acc, S, phase, cc);
    bool seeds_by_height = true;
    uint32_t chosen[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (acc[0] < 3.f) {
        // the three lowest-z points (std::partial_sort over indices, :173-181).  Parallel pick by
        // (z, index); if z ties reach across the cut the SET libstdc++'s heap-select keeps depends on
        // its heap history, so in that (rare) case one thread replays the heap exactly.
        // (indices below are positions in the NODE; with a cluster a point of another block is read through
        // distributed shared memory: rare path, three points or one sequential replay)
        auto gcoord = [&](uint32_t g, int axis) -> float {
            if constexpr (CL > 1) {
                const uint32_t per = ((n + CL - 1) / CL + 3u) & ~3u, r = g / per;
                float* base = axis == 0 ? S.x : (axis == 1 ? S.y : S.z);
                return cg::this_cluster().map_shared_rank(base, r)[g - r * per];
            } else {
                return nv.coord(g, axis);
            }
        };
        for (int r = 0; r < 3; ++r) {
            unsigned long long best = ~0ull;
            for (uint32_t i = tid; i < nl; i += TT) {
                const uint32_t g = lo + i;
                if (g == chosen[0] || g == chosen[1]) continue;
                const unsigned long long key = ((unsigned long long)f2ord(nv.coord(i, 2)) << 32) | g;
                best = key < best ? key : best;
            }
            best = block_min_u64<TT>(best, S.red, phase);
            if constexpr (CL > 1) {  // the smallest key of the cluster
                cg::cluster_group cluster = cg::this_cluster();
                unsigned long long* my = reinterpret_cast<unsigned long long*>(cc.xch + cc.xphase * 16);
                cc.xphase ^= 1;
                if (tid == 0) my[0] = best;
                cluster.sync();
                for (int q = 0; q < CL; ++q) { const unsigned long long o = cluster.map_shared_rank(my, q)[0]; best = o < best ? o : best; }
            }
            chosen[r] = (uint32_t)(best & 0xFFFFFFFFu);
        }
        {
            const float v3 = gcoord(chosen[2], 2);
            float ties[1] = {0.f};
            for (uint32_t i = tid; i < nl; i += TT) ties[0] += (nv.coord(i, 2) <= v3) ? 1.f : 0.f;
            node_sum<TT, 1, CL>(ties, S, phase, cc);
            if (ties[0] > 3.f) {
                // replay of std::__heap_select(first, first + 3, last, z[a] < z[b]) on a 3-element heap
                // (make_heap with one __adjust_heap, then __pop_heap for every later element that is
                // strictly below the root); same code as oracle/rpw_oracle.c:lowest3
                if (tid == 0 && lead) {
                    uint32_t hp[3] = {0, 1, 2};
                    auto zz = [&](uint32_t i) { return gcoord(i, 2); };
                    auto adjust = [&](uint32_t value) {
                        uint32_t hole = 0;
                        uint32_t second = 2;
                        if (zz(hp[2]) < zz(hp[1])) second = 1;
                        hp[0] = hp[second];
                        hole = second;
                        if (zz(hp[0]) < zz(value)) { hp[hole] = hp[0]; hole = 0; }  // __push_heap: parent is the root
                        hp[hole] = value;
                    };
                    adjust(hp[0]);
                    for (uint32_t i = 3; i < n; ++i)
                        if (zz(i) < zz(hp[0])) adjust(i);
                    S.misc[2] = hp[0]; S.misc[3] = hp[1]; S.misc[4] = hp[2];
                }
                if constexpr (CL > 1) {
                    cg::cluster_group cluster = cg::this_cluster();
                    cluster.sync();
                    const uint32_t* m0 = cluster.map_shared_rank(S.misc, 0);
                    chosen[0] = m0[2]; chosen[1] = m0[3]; chosen[2] = m0[4];
                    cluster.sync();
                } else {
                    __syncthreads();
                    chosen[0] = S.misc[2]; chosen[1] = S.misc[3]; chosen[2] = S.misc[4];
                    __syncthreads();
                }
            }
        }
        // ascending index so that the 3-term sums follow the reference's order
        if (chosen[0] > chosen[1]) { const uint32_t t = chosen[0]; chosen[0] = chosen[1]; chosen[1] = t; }
        if (chosen[1] > chosen[2]) { const uint32_t t = chosen[1]; chosen[1] = chosen[2]; chosen[2] = t; }
        if (chosen[0] > chosen[1]) { const uint32_t t = chosen[0]; chosen[0] = chosen[1]; chosen[1] = t; }
        for (uint32_t i = tid; i < nl; i += TT) nv.set_mask(i, (lo + i == chosen[0] || lo + i == chosen[1] || lo + i == chosen[2]) ? 1 : 0);
        acc[0] = 3.f; acc[1] = acc[2] = acc[3] = 0.f;
        for (int r = 0; r < 3; ++r) { acc[1] += gcoord(chosen[r], 0); acc[2] += gcoord(chosen[r], 1); acc[3] += gcoord(chosen[r], 2); }
        seeds_by_height = false;
    }

    tick(1);
    // ---- iterated plane fit (:185-217) -------------------------------------------------------
    // One pass per iteration: the distance pass that builds the next mask also accumulates that
    // mask's first and second moments about the CURRENT centroid; the next centroid is mu + s/n and
    // the next scatter matrix is S' - s s^T / n (the shift s/n is small, so nothing cancels badly).
    // Only the very first fit (seeds) needs a separate covariance pass.
    float cnt = acc[0];
    float cx = acc[1] / cnt, cy = acc[2] / cnt, cz = acc[3] / cnt;  // computeCentroid (seeds by height: relative to the pivot)
    float nx = 0.f, ny = 0.f, nz = 1.f, residual = FLT_MAX;
    float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // scatter of the current mask about (cx, cy, cz): xx yx yy zx zy zz
    bool have_cv = false;
    if (seeds_by_height) {
        cv[0] = fmaf(-acc[1], cx, acc[4]); cv[1] = fmaf(-acc[2], cx, acc[5]); cv[2] = fmaf(-acc[2], cy, acc[6]);
        cv[3] = fmaf(-acc[3], cx, acc[7]); cv[4] = fmaf(-acc[3], cy, acc[8]); cv[5] = fmaf(-acc[3], cz, acc[9]);
        cx += px; cy += py; cz += pz;
        // more than four bits lost to the pivot offset: take the covariance pass about the centroid instead
        have_cv = !((acc[4] + acc[6] + acc[9]) > 16.f * (cv[0] + cv[2] + cv[5]));
    }
    int iters = 0;
    bool have_final = false;  // final plane (:220-228) already known
    float* bc = reinterpret_cast<float*>(S.misc + 8);
    if (replay != 0) {  // (replay == 0: every fit runs in the reference's order below, the tree-sum fit is skipped)
    auto covariance_pass = [&]() {  // computeCovariance about the centroid (point_cloud_processor.cpp:72-86)
#pragma unroll
        for (int k = 0; k < 6; ++k) cv[k] = 0.f;
        for_points<TT, SMEM, true>(nv, nl, [&](uint32_t, float x, float y, float z, uint8_t m) {
            // masked-out points contribute exact zeros
            const float dx = m ? x - cx : 0.f, dy = m ? y - cy : 0.f, dz = m ? z - cz : 0.f;
            cv[0] = fmaf(dx, dx, cv[0]); cv[1] = fmaf(dy, dx, cv[1]); cv[2] = fmaf(dy, dy, cv[2]);
            cv[3] = fmaf(dz, dx, cv[3]); cv[4] = fmaf(dz, dy, cv[4]); cv[5] = fmaf(dz, dz, cv[5]);
        });
        node_sum<TT, 6, CL>(cv, S, phase, cc);
        have_cv = true;
    };
    for (int iter = 0; iter < fp.max_iter; ++iter) {
        if (cnt < 3.f) break;  // :196 — collapsed mask is kept (Q3)
        if (!have_cv) covariance_pass();
        tick(2);
        plane_normal<EXACT, TT>(cv, cnt, bc, nx, ny, nz, fp.hybrid != 0, A.timing);
        iters++;
        tick(3);
        // distances, new mask, convergence, residual of the fit just made, moments of the new mask
        float st[12] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for_points<TT, SMEM, true>(nv, nl, [&](uint32_t i, float x, float y, float z, uint8_t om) {
            const float dx = x - cx, dy = y - cy, dz = z - cz;
            const float p0 = dx * nx, p1 = dy * ny, p2 = dz * nz;
            const float dist = fabsf(p0 + (p1 + p2));  // Eigen's dot order, see plane_dist
            const bool nm = dist < tau;
            st[5] += om ? dist : 0.f;
            if (nm != (om != 0)) { st[4] = 1.f; nv.set_mask(i, nm ? 1 : 0); }
            const float ex = nm ? dx : 0.f, ey = nm ? dy : 0.f, ez = nm ? dz : 0.f;
            st[0] += nm ? 1.f : 0.f; st[1] += ex; st[2] += ey; st[3] += ez;
            st[6] = fmaf(ex, ex, st[6]); st[7] = fmaf(ey, ex, st[7]); st[8] = fmaf(ey, ey, st[8]);
            st[9] = fmaf(ez, ex, st[9]); st[10] = fmaf(ez, ey, st[10]); st[11] = fmaf(ez, ez, st[11]);
            // (predicated accumulation, "if (nm) { st[..] += .. }", instead of adding selected zeros: 6 % slower, 1.276
            // against 1.202 ms per 512 scans -- ptxas turns the block into a branch)
        });
        tick(4);
        node_sum<TT, 12, CL>(st, S, phase, cc);
        tick(13);
        if (st[4] == 0.f) {  // :215 converged: the final fit repeats this one
            residual = st[5] / cnt;
            have_final = true;
            break;
        }
        cnt = st[0];
        if (cnt >= 3.f) {
            const float mx = st[1] / cnt, my = st[2] / cnt, mz = st[3] / cnt;  // centroid shift
            cv[0] = fmaf(-st[1], mx, st[6]); cv[1] = fmaf(-st[2], mx, st[7]); cv[2] = fmaf(-st[2], my, st[8]);
            cv[3] = fmaf(-st[3], mx, st[9]); cv[4] = fmaf(-st[3], my, st[10]); cv[5] = fmaf(-st[3], mz, st[11]);
            cx += mx; cy += my; cz += mz;
            have_cv = true;
        }
    }
    // ---- final fit (:220-228) ----------------------------------------------------------------
    if (!have_final) {
        if (cnt >= 3.f) {
            if (!have_cv) covariance_pass();
            plane_normal<EXACT, TT>(cv, cnt, bc, nx, ny, nz, fp.hybrid != 0, A.timing);
            float rs[1] = {0.f};
            for_points<TT, SMEM, true>(nv, nl, [&](uint32_t, float x, float y, float z, uint8_t m) {
                rs[0] += m ? plane_dist(x, y, z, cx, cy, cz, nx, ny, nz) : 0.f;
            });
            node_sum<TT, 1, CL>(rs, S, phase, cc);
            residual = rs[0] / cnt;
        } else {
            cx = cy = cz = 0.f; nx = ny = 0.f; nz = 1.f; residual = FLT_MAX;  // :78-80
        }
    }
    }  // replay != 0
    if (replay >= 0 && (replay == 0 || iters > replay)) {
        // Reference-order refit: this node's fit again from its seeds with the reference's sequential sums and the QR
        // solver; the fast fit above only decided that the node is worth it (long runs are the ones that amplify the
        // last bits of the moments: bistable or creeping masks).
        FitState fs;
        exact_refit<TT, SMEM>(nv, n, z_th, tau, fp.max_iter, seeds_by_height, chosen[0], chosen[1], chosen[2], S, &fs);
        cx = fs.cx; cy = fs.cy; cz = fs.cz; nx = fs.nx; ny = fs.ny; nz = fs.nz; residual = fs.residual; cnt = fs.cnt; iters = fs.iters;
        tick.count(15, 1);
    }
    const int n_in = (int)cnt;
    tick(5);
    tick.count(10, 1);
    tick.count(11, (unsigned)iters);

    // ---- split decision (:231-235) -----------------------------------------------------------
    const float split_threshold = fp.th_dist * (1.0f + 1.5f * (float)depth);
    const uint32_t min_patch = (uint32_t)(50 + 10 * depth);
    if (!(residual > split_threshold && depth < fp.max_split_depth && n >= min_patch)) {
        // leaf: slot j labels input point sortedA[start + j].w (positional read-back, Q1)
        {
            const uint32_t slot0 = nd.start + lo;
            constexpr int kWide = TT <= 256 ? RPW_WIDE : 8;
            uint32_t i = tid;
            for (; i + (kWide - 1) * TT < nl; i += kWide * TT) {
                uint32_t w[kWide];
#pragma unroll
                for (int u = 0; u < kWide; ++u) w[u] = slot_input_index(A, slot0 + i + u * TT);
#pragma unroll
                for (int u = 0; u < kWide; ++u) A.labels[w[u]] = nv.mask(i + u * TT);
            }
            for (; i + 3 * TT < nl; i += 4 * TT) {
                uint32_t w[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) w[u] = slot_input_index(A, slot0 + i + u * TT);
#pragma unroll
                for (int u = 0; u < 4; ++u) A.labels[w[u]] = nv.mask(i + u * TT);
            }
            for (; i < nl; i += TT) A.labels[slot_input_index(A, slot0 + i)] = nv.mask(i);
        }
        if (tid == 0 && lead) dbg_record(A, nd, depth, RPW_NODE_FIT, iters, n_in, -1, cx, cy, cz, nx, ny, nz, residual, 0, mean_dist);
        tick(6);
        return iters;
    }

    // ---- split (:238-283) --------------------------------------------------------------------
    float sxy[2] = {0.f, 0.f}, var[2] = {0.f, 0.f};
    float ccx, ccy;
    if (replay >= 0) {  // :240-249 in the reference's order
        seq_sums<TT, SMEM, 2>(nv, n, reinterpret_cast<float*>(S.hist), reinterpret_cast<float*>(S.misc + 16), sxy,
                              [](uint32_t, float x, float y, float, uint8_t, float (&v)[2]) { v[0] = x; v[1] = y; });
        ccx = sxy[0] / (float)n; ccy = sxy[1] / (float)n;
        const float qx = ccx, qy = ccy;
        seq_sums<TT, SMEM, 2>(nv, n, reinterpret_cast<float*>(S.hist), reinterpret_cast<float*>(S.misc + 16), var,
                              [=](uint32_t, float x, float y, float, uint8_t, float (&v)[2]) {
                                  const float dx = x - qx, dy = y - qy;
                                  v[0] = dx * dx; v[1] = dy * dy;
                              });
        if (TT > 32) __syncthreads();  // the histogram area goes back to the radix select
    } else {
    for (uint32_t i = tid; i < nl; i += TT) {
        float x, y, z;
        nv.get(i, x, y, z);
        sxy[0] += x; sxy[1] += y;
    }
    node_sum<TT, 2, CL>(sxy, S, phase, cc);
    ccx = sxy[0] / (float)n; ccy = sxy[1] / (float)n;
    for (uint32_t i = tid; i < nl; i += TT) {
        float x, y, z;
        nv.get(i, x, y, z);
        const float dx = x - ccx, dy = y - ccy;
        var[0] = fmaf(dx, dx, var[0]); var[1] = fmaf(dy, dy, var[1]);
    }
    node_sum<TT, 2, CL>(var, S, phase, cc);
    }
    const int axis = (var[0] / (float)n > var[1] / (float)n) ? 0 : 1;  // :250
    const float median = radix_select<TT, SMEM, CL>(nv, nl, axis, n / 2);   // upper median (Q7)

    // stable partition: thread t owns the contiguous run [t*rper, (t+1)*rper) of this block's points
    const uint32_t rper = (nl + TT - 1) / TT;
    const uint32_t rlo = min(nl, (uint32_t)tid * rper), rhi = min(nl, rlo + rper);
    uint32_t nleft = 0;
    for (uint32_t i = rlo; i < rhi; ++i) nleft += nv.coord(i, axis) <= median;
    // block exclusive scan of nleft
    uint32_t* wsum = S.hist;  // reuse (256 words)
    const int lane = tid & 31, warp = tid >> 5;
    uint32_t inc = nleft;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
    }
    __syncthreads();
    if (lane == 31) wsum[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0, total_left = 0;
#pragma unroll
    for (int w = 0; w < (TT / 32); ++w) {
        const uint32_t v = wsum[w];
        if (w < warp) wbase += v;
        total_left += v;
    }
    uint32_t lefts_before_block = 0;
    if constexpr (CL > 1) {
        // the blocks of the cluster in rank order: lefts of the blocks before this one, lefts of the whole node
        cg::cluster_group cluster = cg::this_cluster();
        uint32_t* my = reinterpret_cast<uint32_t*>(cc.xch + cc.xphase * 16);
        cc.xphase ^= 1;
        if (tid == 0) my[0] = total_left;
        cluster.sync();
        uint32_t all = 0;
        for (int q = 0; q < CL; ++q) {
            const uint32_t v = cluster.map_shared_rank(my, q)[0];
            if (q < (int)cc.rank) lefts_before_block += v;
            all += v;
        }
        total_left = all;
    }
    uint32_t lpos = lefts_before_block + wbase + inc - nleft;  // lefts of the node before my run
    uint32_t rpos = total_left + (lo + rlo - lpos);             // rights before my run = points before it - lefts before it
    float4* dst = (((depth + 1) & 1) ? A.bufB : A.bufC) + nd.start;
    for (uint32_t i = rlo; i < rhi; ++i) {
        float x, y, z;
        nv.get(i, x, y, z);
        const float v = axis == 0 ? x : y;
        if (v <= median) dst[lpos++] = make_float4(x, y, z, 0.f);
        else dst[rpos++] = make_float4(x, y, z, 0.f);
    }
    // children (:286-287) go to the next level's queue; the parent range [start, start+n) is
    // simply cut in two, which IS the reference's concatenated return order (Q1).
    NodeRef L, R;
    L.start = nd.start; L.n = total_left; L.root = nd.root; L.pad = 0;
    R.start = nd.start + total_left; R.n = n - total_left; R.root = nd.root; R.pad = 0;
    if (L.n < 3 && lead) {
        label_const<TT>(A, L, 0);
        if (tid == 0) dbg_record(A, L, depth + 1, RPW_NODE_SMALL, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, 0);
    }
    if (R.n < 3 && lead) {
        label_const<TT>(A, R, 0);
        if (tid == 0) dbg_record(A, R, depth + 1, RPW_NODE_SMALL, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, 0);
    }
    if (tid == 0 && lead) {
        const uint32_t k = (L.n >= 3) + (R.n >= 3);
        if (k) {
            uint32_t slot = atomicAdd(A.q_count + depth + 1, k);
            NodeRef* q = A.queue[(depth + 1) & 1];
            if (slot + k <= A.q_cap) {
                if (L.n >= 3) q[slot++] = L;
                if (R.n >= 3) q[slot] = R;
            } else {
                atomicExch(A.overflow, 1u);
            }
        }
        dbg_record(A, nd, depth, RPW_NODE_SPLIT, iters, n_in, axis, cx, cy, cz, nx, ny, nz, residual, median, mean_dist);
    }
    tick(7);
    return iters;
}

__device__ __forceinline__ FitSmem carve_smem(unsigned char* raw, int cap, int warps, bool with_ring = false) {
    FitSmem S;
    S.x = reinterpret_cast<float*>(raw);
    S.y = S.x + cap;
    S.z = S.y + cap;
    S.red = S.z + cap;
    S.hist = reinterpret_cast<uint32_t*>(S.red + 2 * warps * kRedMax);
    S.misc = S.hist + 256;
    S.m = reinterpret_cast<uint8_t*>(S.misc + kMiscWords);
    S.ring = nullptr; S.bar = nullptr;
    if (with_ring) {
        const uintptr_t at = (reinterpret_cast<uintptr_t>(S.m) + (size_t)cap + 127) & ~(uintptr_t)127;
        S.ring = reinterpret_cast<float4*>(at);
        S.bar = reinterpret_cast<uint64_t*>(S.ring + (size_t)RPW_TMA_STAGES * kRingTile);
    }
    return S;
}

// ---------------------------------------------------------------------------------------------
// K3a: level 0 — the ring/sector patches themselves, one block per patch, no grid-wide barrier.
// Patches come in very different sizes (a few points near the sensor, >10k in the far rings), so
// the launch is split into size classes, each with its own block size and shared-memory carve-out:
//   64 threads: <= 1024 points (15 KB) and <= 2048 (28 KB) : many resident blocks, their eigensolves overlap
//   128 threads: <= 3072 (41 KB) and <= 4096 (55 KB)
//   256 threads: <= 5632 (74 KB, three per SM)
//   512 threads: everything larger: <= 8192 points shared-memory resident (108 KB), beyond that
//                streamed from L2; the far-ring patches that iterate longest get the most threads
// The class kernels run concurrently on separate streams and the hardware block scheduler packs
// whatever mix fits an SM.  blockIdx.x walks (patch rank, scan) largest patch first; a block whose
// patch belongs to another class exits at once.
// EXACT = true: plane normals from Eigen's QR sequence (bit-comparable with the CPU reference);
// EXACT = false: closed-form FP64 smallest eigenvector (faster, ~1e-6 rad away from the reference's
// float solver, which is enough to tip chaotic two-layer patches the other way).
// ---------------------------------------------------------------------------------------------
template <int TT, bool EXACT, bool REPLAY>
// (the 512-thread kernel of the reference-order build is compiled for one block per SM: its sequential sums run in ONE warp,
// whose register ring of chunks in flight spilled at the 64 registers two resident blocks allow)
__global__ void __launch_bounds__(TT, (TT <= 32 ? RPW_LB32 : TT <= 64 ? (REPLAY ? 8 : RPW_LB64) : TT <= 128 ? (REPLAY ? RPW_LB128_REPLAY : RPW_LB128) : TT <= 256 ? (REPLAY ? RPW_LB256_REPLAY : 3) : (REPLAY ? RPW_LB512_REPLAY : (TT > 512 ? 1 : 2))))
rpw_fit_roots_kernel(FitArgs A, int cls, int cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // The class's work list and its length are read together (independent addresses, one latency).
    // The host sizes the grid from the previous launch group's counts; a block whose index is past
    // the list leaves at once, and a grid shorter than the list strides over it.
    const uint4* list = A.cls_list + (size_t)cls * A.cls_cap;
    uint32_t i = blockIdx.x;
    uint4 item = __ldcg(list + (i < A.cls_cap ? i : 0));
    const uint32_t count = min(__ldcg(A.cls_count + cls), A.cls_cap);
    if (i >= count) return;
    constexpr bool kRing = RPW_STREAM_TMA && TT == RPW_STREAM_THREADS;
    FitSmem S = carve_smem(smem_raw, cap, TT / 32, kRing && cap == kCapStream);
    if constexpr (kRing) ring_init(S);
    for (;;) {
        NodeRef nd;
        nd.start = item.x; nd.n = item.y; nd.root = item.z; nd.pad = 0;
        const uint32_t next = i + gridDim.x;
        if (next < count) item = __ldcg(list + next);  // in flight while this node is processed
        TraceScope trace(A, nd.n, 0, cls);
        int iters = 0;
        if (nd.n <= (uint32_t)cap) iters = process_node<TT, true, EXACT, REPLAY>(A, nd, 0, S);
        else if constexpr (TT == RPW_STREAM_THREADS) iters = process_node<TT, false, EXACT, REPLAY>(A, nd, 0, S);
        trace.done(iters);
        if (threadIdx.x == 0) atomicAdd(A.stats + 1, 1u);
        if (next >= count) break;
        i = next;
        __syncthreads();  // the next node reuses the shared-memory arrays
    }
}

// ---------------------------------------------------------------------------------------------
// K3a, cluster form (calls of one or two scans): one thread-block CLUSTER per listed root patch, the patch's points
// spread over the shared memory of its CL blocks (see Clu).  A call of a single scan leaves most SMs empty and ends with
// its longest-iterating patch; an iteration of a big patch is bounded by the pass, which one SM issues at n x ~38
// instructions / 4 schedulers: CL SMs issue it CL times faster, the serial rest of the iteration (reduction, eigensolve)
// stays, plus one cluster barrier per reduction.
// ---------------------------------------------------------------------------------------------
template <int TT, bool EXACT, int CL>
__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(TT, 1) rpw_fit_cluster_kernel(FitArgs A, int cls, int cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint4* list = A.cls_list + (size_t)cls * A.cls_cap;
    const uint32_t count = min(__ldcg(A.cls_count + cls), A.cls_cap);
    const uint32_t stride = gridDim.x / CL;
    FitSmem S = carve_smem(smem_raw, cap, TT / 32);
    Clu<CL> cc;
    cc.rank = cluster.block_rank();
    cc.xch = reinterpret_cast<float*>(S.misc + 32);
    cc.tot = reinterpret_cast<float*>(S.misc + 64);
    FitArgs Aq = A;  // the per-node timeline is written by the cluster's first block only
    if (cc.rank != 0) { Aq.trace = nullptr; Aq.timing = nullptr; }
    for (uint32_t i = blockIdx.x / CL; i < count; i += stride) {  // the same trip count in every block of a cluster
        const uint4 item = __ldcg(list + i);
        NodeRef nd;
        nd.start = item.x; nd.n = item.y; nd.root = item.z; nd.pad = 0;
        TraceScope trace(Aq, nd.n, 0, cls);
        const int iters = process_node<TT, true, EXACT, false, CL>(Aq, nd, 0, S, cc);
        trace.done(iters);
        if (threadIdx.x == 0 && cc.rank == 0) atomicAdd(A.stats + 1, 1u);
        cluster.sync();  // nobody reads this block's shared memory any more: the next node may overwrite it
    }
}

// ---------------------------------------------------------------------------------------------
// K3b: levels >= 1 — persistent kernel, level-synchronous device worklist with no host round trip:
// every split at level l pushed its children to the queue of level l+1; a grid-wide barrier
// separates levels; the kernel ends when a level enqueued nothing.  Typical scans never split: every
// block then reads an empty queue and leaves without touching the barrier.
//
// The grid is one block per SM, launched cooperatively (all blocks resident, so the barrier cannot
// deadlock), and the barrier itself is a counter in global memory (arrive + spin with nanosleep).
// All cross-block data (children, queues, counters) is read with ld.global.cg, so no stale L1 lines
// are involved.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(uint32_t* ctr, uint32_t target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*reinterpret_cast<volatile uint32_t*>(ctr) < target) __nanosleep(64);
        __threadfence();
    }
    __syncthreads();
}

// (Fitting the children inside the block that split the parent -- depth first, a small stack in shared memory, no queue, no
// level kernel work for the few splits of a C2 batch -- was built three ways, labels identical each time.  One call site with
// a run-time depth: the roots' own fits slow down 5 % (the depth-0 specialisation is lost), net C2 1.830 -> 1.843 ms per 512
// scans, C5 1.470 -> 1.379 per 64, C4 1.486 -> 1.713 (two halves of a 20 k-point patch are better off on two SMs).  A second
// inlined call site for the children: C2 2.197.  The children out of line behind a __grid_constant__ parameter, resident
// parents only: C2 1.817, C4 1.437 -> 1.472, C5 1.471 -> 1.466.  Nothing worth its code; the queue stays.)
// (A second form for calls of one or two scans -- 256 threads, one block per SM with an 8192-point slot -- changed nothing:
// C5 single scan 0.408 -> 0.405 ms; the deeper levels of a scan are ~15 us each, mostly the grid barrier and a few short nodes.)
template <bool EXACT, bool REPLAY>
__global__ void __launch_bounds__(kFitThreads, kLevelBlocksPerSm) rpw_fit_levels_kernel(FitArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int TT = kFitThreads;
    FitSmem S = carve_smem(smem_raw, A.smem_cap, TT / 32);
    __shared__ uint32_t s_fetch;
    uint32_t n_done = 0, pre = 0, bar_target = 0;
    // grid estimate for the next launch group (zero-copy store); word kClsWords - 2 is the overflow report below
    if (blockIdx.x == 0 && threadIdx.x < kClsWords && threadIdx.x != kClsWords - 2 && A.host_counts)
        A.host_counts[threadIdx.x] = threadIdx.x < kNumFitClasses ? __ldcg(A.cls_count + threadIdx.x) : (uint32_t)A.n_scans;
    Tick ktick(A.timing);
    int level = 0;
    for (;;) {
        const uint32_t cnt = min(__ldcg(A.q_count + level + 1), A.q_cap);
        if (cnt == 0) break;
        level++;
        const NodeRef* q = A.queue[level & 1];
        // the cursor of the NEXT node is fetched while the current one is processed
        if (threadIdx.x == 0) pre = atomicAdd(A.fetch_ctr + level, 1u);
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_fetch = pre;
            __syncthreads();
            const uint32_t id = s_fetch;
            ktick(8);
            if (id >= cnt) break;
            if (threadIdx.x == 0) pre = atomicAdd(A.fetch_ctr + level, 1u);
            NodeRef nd;
            const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(q + id));
            nd.start = raw.x; nd.n = raw.y; nd.root = raw.z; nd.pad = 0;
            TraceScope trace(A, nd.n, level, 0xFFFF);
            int iters;
            if (nd.n <= (uint32_t)A.smem_cap) iters = process_node<TT, true, EXACT, REPLAY>(A, nd, level, S);
            else iters = process_node<TT, false, EXACT, REPLAY>(A, nd, level, S);
            trace.done(iters);
            n_done++;
            if (A.timing && threadIdx.x == 0) ktick.last = clock64();
        }
        bar_target += gridDim.x;
        grid_barrier(A.stats + 4, bar_target);
        ktick(9);
    }
    // bookkeeping + self-cleaning: the last block to arrive publishes the totals and zeroes the
    // per-level counters, so the next call needs no memset
    if (threadIdx.x == 0) {
        if (n_done) atomicAdd(A.stats + 1, n_done);
        __threadfence();
        const uint32_t arrived = atomicAdd(A.stats + 2, 1u);
        if (arrived == gridDim.x - 1) {
            A.stats[0] = max(A.stats[0], (uint32_t)level + 1);
            A.stats[3] += atomicExch(A.stats + 1, 0u);
            A.stats[2] = 0;
            A.stats[4] = 0;
            for (int l = 0; l <= level + 1; ++l) { A.fetch_ctr[l] = 0; A.q_count[l] = 0; }
            // a worklist that overflowed dropped children: reported to the host through mapped memory, so that every
            // entry point sees it after its synchronisation, with or without a stats request
            if (atomicExch(A.overflow, 0u) != 0u && A.host_counts) A.host_counts[kClsWords - 2] = 1u;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// host-side pieces shared by the instantiating units
// ---------------------------------------------------------------------------------------------
// Size classes of the level-0 kernel: (threads, largest patch, shared-memory capacity in points).
// Finer steps waste less shared memory per resident patch (a 2100-point patch in a 4096-point slot
// blocks twice the memory it needs for its whole 40 us life), which is what bounds the fit phase.
struct FitClass { int threads; uint32_t hi; int cap; int cluster; };
// Two tables.  [0] throughput: batches keep every SM full of patches, so a patch gets just enough threads and the
// smallest shared-memory slot that holds it.  [1] latency (calls of one or two scans, most SMs empty, the call ends with its
// longest-iterating patch): patches above 4096 points go to thread-block clusters of 4 (<= 8192 points, 2048 per block)
// or 8 blocks (<= 65536 points, 8192 per block), which issue a pass 4 / 8 times faster than one SM can.
// (What did NOT work for the latency table: more threads per patch on ONE SM -- twice to four times the threads per class,
// a 1024-thread class keeping up to 16384 points resident: C2 fit 102 -> 127 us, C5 397 -> 480 us, C4 596 -> 537 us.  One SM's
// four schedulers issue a pass at the same rate however many warps share it, and larger blocks pay more per barrier.)
static const FitClass kFitClasses[2][kNumFitClasses] = {
    {{RPW_T0, 1024, 1024, 1}, {RPW_T1, 2048, 2048, 1}, {128, 3072, 3072, 1}, {128, 4096, 4096, 1}, {256, 5632, 5632, 1},
     {256, kCapLarge, kCapLarge, 1}, {RPW_STREAM_THREADS, 0xFFFFFFFFu, kCapStream, 1}},
    {{RPW_LT0, 1024, 1024, 1}, {RPW_LT1, 2048, 2048, 1}, {RPW_LT2, 3072, 3072, 1}, {RPW_LT2, 4096, 4096, 1}, {256, 8192, 2048, 4},
     {256, 65536, 8192, 8}, {RPW_STREAM_THREADS, 0xFFFFFFFFu, kCapStream, 1}},
};

inline size_t fit_smem_bytes_inl(int smem_cap, int threads) {
    return (size_t)smem_cap * 13 + (2 * (threads / 32) * kRedMax + 256 + kMiscWords) * 4 + 16;
}
// the streamed class of the roots kernel also carries the TMA ring
inline size_t roots_smem_bytes(int cap, int threads) {
    return fit_smem_bytes_inl(cap, threads) + ((threads == RPW_STREAM_THREADS && cap == kCapStream) ? kRingBytes : 0);
}

template <typename KernelT>
static cudaError_t set_smem(KernelT k, size_t bytes) {
    return cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

template <int TT, bool REPLAY>
static cudaError_t configure_roots() {
    size_t need = 0;
    for (const auto& table : kFitClasses)
        for (const FitClass& c : table) if (c.threads == TT && c.cluster == 1) need = roots_smem_bytes(c.cap, TT) > need ? roots_smem_bytes(c.cap, TT) : need;
    cudaError_t e = set_smem(rpw_fit_roots_kernel<TT, true, REPLAY>, need);
    if (e != cudaSuccess) return e;
    return set_smem(rpw_fit_roots_kernel<TT, false, REPLAY>, need);
}

template <bool REPLAY>
static cudaError_t fit_configure_t(int smem_cap, int* blocks_per_sm) {
    cudaError_t e;
    if ((e = configure_roots<RPW_T0, REPLAY>()) != cudaSuccess) return e;
    if ((e = configure_roots<RPW_T1, REPLAY>()) != cudaSuccess) return e;
    if ((e = configure_roots<128, REPLAY>()) != cudaSuccess) return e;
    if ((e = configure_roots<256, REPLAY>()) != cudaSuccess) return e;
    if ((e = configure_roots<RPW_STREAM_THREADS, REPLAY>()) != cudaSuccess) return e;
    if constexpr (!REPLAY) {
        if ((e = set_smem(rpw_fit_cluster_kernel<256, true, 4>, fit_smem_bytes_inl(2048, 256))) != cudaSuccess) return e;
        if ((e = set_smem(rpw_fit_cluster_kernel<256, false, 4>, fit_smem_bytes_inl(2048, 256))) != cudaSuccess) return e;
        if ((e = set_smem(rpw_fit_cluster_kernel<256, true, 8>, fit_smem_bytes_inl(8192, 256))) != cudaSuccess) return e;
        if ((e = set_smem(rpw_fit_cluster_kernel<256, false, 8>, fit_smem_bytes_inl(8192, 256))) != cudaSuccess) return e;
    }
    const size_t smem = fit_smem_bytes_inl(smem_cap, kFitThreads);
    if ((e = set_smem(rpw_fit_levels_kernel<true, REPLAY>, smem)) != cudaSuccess) return e;
    if ((e = set_smem(rpw_fit_levels_kernel<false, REPLAY>, smem)) != cudaSuccess) return e;
    int a = 0, b = 0;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, rpw_fit_levels_kernel<true, REPLAY>, kFitThreads, smem)) != cudaSuccess) return e;
    if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, rpw_fit_levels_kernel<false, REPLAY>, kFitThreads, smem)) != cudaSuccess) return e;
    *blocks_per_sm = a < b ? a : b;
    return cudaSuccess;
}

template <int TT, bool REPLAY>
static void launch_roots_tt(cudaStream_t st, const FitArgs& args, int cls, int cap, unsigned grid) {
    const size_t sm = roots_smem_bytes(cap, TT);
    if (args.fp.exact_eig) rpw_fit_roots_kernel<TT, true, REPLAY><<<grid, TT, sm, st>>>(args, cls, cap);
    else rpw_fit_roots_kernel<TT, false, REPLAY><<<grid, TT, sm, st>>>(args, cls, cap);
}

// size class cls in [0, kNumFitClasses): patches with kFitClasses[cls-1].hi < n <= kFitClasses[cls].hi
template <bool REPLAY>
static cudaError_t launch_fit_roots_t(cudaStream_t st, const FitArgs& args, int cls, unsigned grid) {
    const FitClass& c = kFitClasses[(!REPLAY && args.profile) ? 1 : 0][cls];
    if (grid == 0) grid = 1;
    if constexpr (!REPLAY) {
        if (c.cluster > 1) {  // one cluster per listed patch
            const size_t sm = fit_smem_bytes_inl(c.cap, 256);
            const unsigned g = grid * (unsigned)c.cluster;
            if (c.cluster == 4) {
                if (args.fp.exact_eig) rpw_fit_cluster_kernel<256, true, 4><<<g, 256, sm, st>>>(args, cls, c.cap);
                else rpw_fit_cluster_kernel<256, false, 4><<<g, 256, sm, st>>>(args, cls, c.cap);
            } else {
                if (args.fp.exact_eig) rpw_fit_cluster_kernel<256, true, 8><<<g, 256, sm, st>>>(args, cls, c.cap);
                else rpw_fit_cluster_kernel<256, false, 8><<<g, 256, sm, st>>>(args, cls, c.cap);
            }
            return cudaGetLastError();
        }
    }
    if (c.threads == RPW_T0) launch_roots_tt<RPW_T0, REPLAY>(st, args, cls, c.cap, grid);
    else if (c.threads == RPW_T1) launch_roots_tt<RPW_T1, REPLAY>(st, args, cls, c.cap, grid);
    else if (c.threads == 128) launch_roots_tt<128, REPLAY>(st, args, cls, c.cap, grid);
    else if (c.threads == 256) launch_roots_tt<256, REPLAY>(st, args, cls, c.cap, grid);
    else launch_roots_tt<RPW_STREAM_THREADS, REPLAY>(st, args, cls, c.cap, grid);
    return cudaGetLastError();
}

// Launched with cudaLaunchCooperativeKernel: the driver only starts the grid when every block can be
// resident, which is what makes the spin barrier safe however many handles share the device (two
// plain launches from different handles could otherwise each hold half of the SMs and wait for the
// other half forever).
template <bool REPLAY>
static cudaError_t launch_fit_levels_t(cudaStream_t st, const FitArgs& args, int grid_blocks) {
    FitArgs a = args;
    void* params[] = {&a};
    void* fn = args.fp.exact_eig ? (void*)rpw_fit_levels_kernel<true, REPLAY> : (void*)rpw_fit_levels_kernel<false, REPLAY>;
    return cudaLaunchCooperativeKernel(fn, dim3(grid_blocks), dim3(kFitThreads), params, fit_smem_bytes_inl(args.smem_cap, kFitThreads), st);
}

}  // namespace rpw
