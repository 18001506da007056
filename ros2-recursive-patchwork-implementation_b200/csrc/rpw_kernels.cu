// csrc/rpw_kernels.cu — the sm_100a kernels of the per-scan ground-segmentation path.
//
//   K1  rpw_bin_kernel      clean + range + radius + angle + ring/sector key, per-block patch
//                           histogram (RP/src/recursive_patchwork.cpp:315-378 steps 1-6a)
//   K1b rpw_offsets_kernel  per-scan exclusive scan of the block histograms -> stable offsets
//   K2  rpw_scatter_kernel  stable counting-sort scatter of (x, y, z, input index) into
//                           ring/sector patch segments (input order inside every patch, Q1)
//   K3  rpw_fit_kernel      persistent cooperative kernel: level-synchronous device worklist over
//                           fitPlaneAndSplit nodes (RP/src/recursive_patchwork.cpp:109-308);
//                           per node: early-outs, seeds, iterated PCA plane fit with a register
//                           3x3 eigensolve, residual mask, split (variance axis, exact radix-select
//                           median, stable partition), child enqueue, label scatter
//   dbg rpw_eig3_kernel / rpw_atan2_kernel   unit-test entry points for the device math
//
// All of it is HBM/L2/shared-memory bound integer-and-float SIMT work; there is no dense
// contraction, so no tensor-core path.  Compiled with -fmad=false (see rpw_device.cuh).
#include "rpw_kernels.h"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace rpw {

// =============================================================================================
// K1: binning
// =============================================================================================
template <int STRIDE>
__device__ __forceinline__ void load_xyz(const float* __restrict__ pts, uint64_t i, float& x, float& y, float& z) {
    if (STRIDE == 4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(pts) + i);
        x = v.x; y = v.y; z = v.z;
    } else {
        const float* p = pts + i * 3;
        x = __ldg(p); y = __ldg(p + 1); z = __ldg(p + 2);
    }
}

template <int STRIDE>
__global__ void __launch_bounds__(kBinThreads) rpw_bin_kernel(const float* __restrict__ pts, const uint64_t* __restrict__ scan_off,
                                                             const uint32_t* __restrict__ chunk_base, ZoneModel zm,
                                                             uint16_t* __restrict__ keys, uint8_t* __restrict__ labels,
                                                             uint32_t* __restrict__ blk_hist, uint32_t* __restrict__ patch_total) {
    extern __shared__ uint32_t s_hist[];
    const int b = blockIdx.y, chunk = blockIdx.x;
    // batch-wide points-per-patch totals (scheduling order of the fit kernel) start from zero;
    // the offsets kernel, which runs after this one, accumulates them
    if (b == 0 && chunk == 0)
        for (int p = threadIdx.x; p < zm.num_patches; p += kBinThreads) patch_total[p] = 0;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const uint32_t base = (uint32_t)chunk * kBinChunk;
    if (base >= n) return;
    const int P = zm.num_patches;
    for (int p = threadIdx.x; p < P; p += kBinThreads) s_hist[p] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31;
#pragma unroll 4
    for (int k = 0; k < kBinChunk / kBinThreads; ++k) {
        const uint32_t i = base + k * kBinThreads + threadIdx.x;
        uint16_t key = kKeyDropped;
        const bool valid = i < n;
        if (valid) {
            float x, y, z;
            load_xyz<STRIDE>(pts, off + i, x, y, z);
            key = bin_key(x, y, z, zm);
            keys[off + i] = key;
            // points that never enter a patch get their final label here; patch points are
            // labelled by the fit kernel when their leaf finishes.
            if (key >= kKeyUnbinned) labels[off + i] = key == kKeyDropped ? 3 : (key == kKeyBeyond ? 2 : 0);
        }
        const uint32_t kk = (valid && key < kKeyUnbinned) ? key : 0xFFFFFFFFu;
        const unsigned peers = __match_any_sync(0xffffffffu, kk);
        if (kk != 0xFFFFFFFFu && lane == __ffs(peers) - 1) atomicAdd(&s_hist[kk], __popc(peers));
    }
    __syncthreads();
    uint32_t* out = blk_hist + ((size_t)chunk_base[b] + chunk) * P;
    for (int p = threadIdx.x; p < P; p += kBinThreads) out[p] = s_hist[p];
}

// =============================================================================================
// K1b: per-scan offsets.  blk_hist[b][c][p] becomes the exclusive prefix over chunks c;
// patch_start[b][p] = scan base + exclusive prefix over patches; patch_start[b][P] = end.
// =============================================================================================
__global__ void __launch_bounds__(256) rpw_offsets_kernel(const uint64_t* __restrict__ scan_off, const uint32_t* __restrict__ chunk_base,
                                                         uint32_t* __restrict__ blk_hist, uint32_t* __restrict__ patch_start,
                                                         uint32_t* __restrict__ patch_total, int P) {
    extern __shared__ uint32_t s_cnt[];  // P + 1
    const int b = blockIdx.x;
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const int chunks = (int)((n + kBinChunk - 1) / kBinChunk);
    uint32_t* h = blk_hist + (size_t)chunk_base[b] * P;
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
        uint32_t run = 0;
        int c = 0;
        for (; c + 4 <= chunks; c += 4) {
            const uint32_t v0 = h[(size_t)(c + 0) * P + p], v1 = h[(size_t)(c + 1) * P + p];
            const uint32_t v2 = h[(size_t)(c + 2) * P + p], v3 = h[(size_t)(c + 3) * P + p];
            h[(size_t)(c + 0) * P + p] = run; run += v0;
            h[(size_t)(c + 1) * P + p] = run; run += v1;
            h[(size_t)(c + 2) * P + p] = run; run += v2;
            h[(size_t)(c + 3) * P + p] = run; run += v3;
        }
        for (; c < chunks; ++c) {
            const uint32_t v = h[(size_t)c * P + p];
            h[(size_t)c * P + p] = run;
            run += v;
        }
        s_cnt[p] = run;
        if (run) atomicAdd(&patch_total[p], run);
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t carry = (uint32_t)off;  // sorted segment of scan b starts at its input offset
        for (int p0 = 0; p0 < P; p0 += 32) {
            const int p = p0 + threadIdx.x;
            const uint32_t v = p < P ? s_cnt[p] : 0;
            uint32_t inc = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
                if ((int)threadIdx.x >= d) inc += t;
            }
            if (p < P) patch_start[(size_t)b * (P + 1) + p] = carry + inc - v;
            carry += __shfl_sync(0xffffffffu, inc, 31);
        }
        if (threadIdx.x == 0) patch_start[(size_t)b * (P + 1) + P] = carry;
    }
}

// =============================================================================================
// K2: stable scatter.  Inside a block every warp owns a contiguous run of kBinChunk/8 points and
// walks it 32 at a time, so (block, warp, round, lane) order == input order; ranks inside a
// 32-group come from __match_any_sync.  No atomics claim slots, so the result is deterministic
// and stable regardless of scheduling (SURVEY Q1 needs that).
// =============================================================================================
template <int STRIDE>
__global__ void __launch_bounds__(kBinThreads) rpw_scatter_kernel(const float* __restrict__ pts, const uint64_t* __restrict__ scan_off,
                                                                 const uint32_t* __restrict__ chunk_base,
                                                                 const uint16_t* __restrict__ keys, const uint32_t* __restrict__ blk_hist,
                                                                 const uint32_t* __restrict__ patch_start, float4* __restrict__ sorted,
                                                                 const uint32_t* __restrict__ patch_total, uint32_t* __restrict__ patch_order,
                                                                 int P) {
    extern __shared__ uint32_t s_off[];  // [warps][P]
    constexpr int kWarps = kBinThreads / 32;
    constexpr int kPerWarp = kBinChunk / kWarps;
    const int b = blockIdx.y, chunk = blockIdx.x;
    if (b == 0 && chunk == 0) {
        // Longest-processing-time-first order for the fit kernel's worklist: patches ranked by
        // their batch-wide point totals, largest first (ties by patch index).  Big far-ring
        // patches are also the ones whose plane fit iterates longest, so they must start early.
        for (int p = threadIdx.x; p < P; p += kBinThreads) s_off[p] = patch_total[p];
        __syncthreads();
        for (int p = threadIdx.x; p < P; p += kBinThreads) {
            const uint32_t t = s_off[p];
            uint32_t rank = 0;
            for (int q = 0; q < P; ++q) { const uint32_t u = s_off[q]; rank += (u > t) || (u == t && q < p); }
            patch_order[rank] = (uint32_t)p;
        }
        __syncthreads();
    }
    const uint64_t off = scan_off[b];
    const uint32_t n = (uint32_t)(scan_off[b + 1] - off);
    const uint32_t base = (uint32_t)chunk * kBinChunk;
    if (base >= n) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kWarps * P; i += kBinThreads) s_off[i] = 0;
    __syncthreads();
    uint32_t* my = s_off + warp * P;
    const uint32_t wbase = base + warp * kPerWarp;
    // phase 1: per-warp counts
    for (int r = 0; r < kPerWarp / 32; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        uint32_t kk = 0xFFFFFFFFu;
        if (i < n) { const uint16_t key = keys[off + i]; if (key < kKeyUnbinned) kk = key; }
        const unsigned peers = __match_any_sync(0xffffffffu, kk);
        if (kk != 0xFFFFFFFFu && lane == __ffs(peers) - 1) my[kk] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    // phase 2: exclusive prefix over warps + block offset + patch offset
    const uint32_t* bh = blk_hist + ((size_t)chunk_base[b] + chunk) * P;
    const uint32_t* ps = patch_start + (size_t)b * (P + 1);
    for (int p = threadIdx.x; p < P; p += kBinThreads) {
        uint32_t run = ps[p] + bh[p];
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
            const uint32_t c = s_off[w * P + p];
            s_off[w * P + p] = run;
            run += c;
        }
    }
    __syncthreads();
    // phase 3: ranks and scatter
    for (int r = 0; r < kPerWarp / 32; ++r) {
        const uint32_t i = wbase + r * 32 + lane;
        uint32_t kk = 0xFFFFFFFFu;
        if (i < n) { const uint16_t key = keys[off + i]; if (key < kKeyUnbinned) kk = key; }
        const unsigned peers = __match_any_sync(0xffffffffu, kk);
        if (kk != 0xFFFFFFFFu) {
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            const uint32_t pos = my[kk] + rank;
            float x, y, z;
            load_xyz<STRIDE>(pts, off + i, x, y, z);
            sorted[pos] = make_float4(x, y, z, __uint_as_float((uint32_t)(off + i)));
        }
        __syncwarp();
        if (kk != 0xFFFFFFFFu && lane == __ffs(peers) - 1) my[kk] += __popc(peers);
        __syncwarp();
    }
}

// =============================================================================================
// K3: fit
// =============================================================================================
struct FitSmem {
    float* x; float* y; float* z; uint8_t* m;
    float* red;        // 2 * kFitWarps * kRedMax floats (ping-pong)
    uint32_t* hist;    // 256
    uint32_t* misc;    // small broadcast area
};

constexpr int kRedMax = 8;

// Sum of K floats over the block; every thread receives the totals (bitwise identical in all
// threads: butterfly shuffles of a commutative op).  One __syncthreads per call; the scratch
// area alternates so that back-to-back calls do not race.
template <int K>
__device__ __forceinline__ void block_sum(float (&v)[K], float* red, int& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], d);
    }
    float* r = red + phase * (kFitWarps * kRedMax);
    phase ^= 1;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) r[warp * kRedMax + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = r[(lane & (kFitWarps - 1)) * kRedMax + k];
#pragma unroll
        for (int d = kFitWarps / 2; d > 0; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        v[k] = t;
    }
}

template <int K>
__device__ __forceinline__ void block_min(float (&v)[K], float* red, int& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v[k] = fminf(v[k], __shfl_xor_sync(0xffffffffu, v[k], d));
    }
    float* r = red + phase * (kFitWarps * kRedMax);
    phase ^= 1;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) r[warp * kRedMax + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; ++k) {
        float t = r[(lane & (kFitWarps - 1)) * kRedMax + k];
#pragma unroll
        for (int d = kFitWarps / 2; d > 0; d >>= 1) t = fminf(t, __shfl_xor_sync(0xffffffffu, t, d));
        v[k] = t;
    }
}

// min over the block of a 64-bit key; all threads get the result.
__device__ __forceinline__ unsigned long long block_min_u64(unsigned long long v, float* red, int& phase) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, v, d);
        v = o < v ? o : v;
    }
    unsigned long long* r = reinterpret_cast<unsigned long long*>(red + phase * (kFitWarps * kRedMax));
    phase ^= 1;
    if (lane == 0) r[warp] = v;
    __syncthreads();
    unsigned long long t = r[lane & (kFitWarps - 1)];
#pragma unroll
    for (int d = kFitWarps / 2; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, t, d);
        t = o < t ? o : t;
    }
    return t;
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
template <bool SMEM>
__device__ float radix_select(const NodeView<SMEM>& nv, uint32_t n, int axis, uint32_t k) {
    uint32_t* hist = nv.s.hist;
    uint32_t* misc = nv.s.misc;
    uint32_t prefix = 0, pmask = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += kFitThreads) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < n; i += kFitThreads) {
            const uint32_t u = f2ord(nv.coord(i, axis));
            if ((u & pmask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
        }
        __syncthreads();
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
__device__ __forceinline__ void label_const(const FitArgs& A, const NodeRef& nd, uint8_t v) {
    for (uint32_t i = threadIdx.x; i < nd.n; i += kFitThreads)
        A.labels[__float_as_uint(A.sortedA[nd.start + i].w)] = v;
}

template <bool SMEM>
__device__ void process_node(const FitArgs& A, const NodeRef nd, const int depth, FitSmem S) {
    const FitParams& fp = A.fp;
    const uint32_t n = nd.n;
    const int tid = threadIdx.x;
    int phase = 0;

    if (n < 3 || depth > fp.max_split_depth) {  // :111-113
        label_const(A, nd, 0);
        if (tid == 0) dbg_record(A, nd, depth, RPW_NODE_SMALL, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, 0);
        return;
    }
    NodeView<SMEM> nv;
    nv.s = S;
    nv.src = (depth == 0 ? A.sortedA : ((depth & 1) ? A.bufB : A.bufC)) + nd.start;
    nv.gmask = A.gmask + nd.start;

    // ---- pass 1: load, bounding box, (root only) mean range ------------------------------
    float mm[6] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};  // min x,y,z, min -x,-y,-z
    float sd[1] = {0.f};
    for (uint32_t i = tid; i < n; i += kFitThreads) {
        const float4 v = depth == 0 ? __ldg(nv.src + i) : __ldcg(nv.src + i);
        if (SMEM) { S.x[i] = v.x; S.y[i] = v.y; S.z[i] = v.z; }
        mm[0] = fminf(mm[0], v.x); mm[1] = fminf(mm[1], v.y); mm[2] = fminf(mm[2], v.z);
        mm[3] = fminf(mm[3], -v.x); mm[4] = fminf(mm[4], -v.y); mm[5] = fminf(mm[5], -v.z);
        if (depth == 0) sd[0] += range2d(v.x, v.y);
    }
    block_min<6>(mm, S.red, phase);
    float mean_dist;
    if (depth == 0) {
        block_sum<1>(sd, S.red, phase);
        mean_dist = sd[0] / (float)n;  // :383-387
        if (tid == 0) A.root_mean[nd.root] = mean_dist;
    } else {
        mean_dist = __ldcg(A.root_mean + nd.root);  // Q4: inherited unchanged
    }
    const float x_min = mm[0], x_max = -mm[3], y_min = mm[1], y_max = -mm[4], z_min = mm[2], z_max = -mm[5];
    const float area = (x_max - x_min) * (y_max - y_min);
    if (area < 25.0f && depth > 0) {  // :126-129
        label_const(A, nd, 1);
        if (tid == 0) dbg_record(A, nd, depth, RPW_NODE_AREA, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, mean_dist);
        return;
    }
    if ((z_max - z_min) < 0.05f && n > 10) {  // :138-140
        label_const(A, nd, 1);
        if (tid == 0) dbg_record(A, nd, depth, RPW_NODE_FLAT, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, mean_dist);
        return;
    }
    // (the barrier inside block_min already made every thread's shared-memory stores visible)

    // ---- seed threshold (:149-160) ----------------------------------------------------------
    const float rel_dist = mean_dist / fp.radius;
    float z_th;
    if (fp.adaptive_seed_height) {
        z_th = fp.sensor_height + 0.2f * rel_dist;
    } else {
        const uint32_t idx = (uint32_t)(0.1f * (float)n);
        z_th = radix_select<SMEM>(nv, n, 2, idx) + fp.th_seeds;
    }
    const float tau = fp.th_dist * (1.0f + 0.2f * rel_dist);  // :203

    // ---- seeds (:163-182) -------------------------------------------------------------------
    float acc[4] = {0.f, 0.f, 0.f, 0.f};  // count, sum x, sum y, sum z over the mask
    for (uint32_t i = tid; i < n; i += kFitThreads) {
        float x, y, z;
        nv.get(i, x, y, z);
        const uint8_t m = z < z_th;
        nv.set_mask(i, m);
        if (m) { acc[0] += 1.f; acc[1] += x; acc[2] += y; acc[3] += z; }
    }
    block_sum<4>(acc, S.red, phase);
    if (acc[0] < 3.f) {
        // the three lowest-z points, lowest index first among equal z (std::partial_sort at
        // :175-176 leaves ties unspecified)
        uint32_t chosen[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        for (int r = 0; r < 3; ++r) {
            unsigned long long best = ~0ull;
            for (uint32_t i = tid; i < n; i += kFitThreads) {
                if (i == chosen[0] || i == chosen[1]) continue;
                const unsigned long long key = ((unsigned long long)f2ord(nv.coord(i, 2)) << 32) | i;
                best = key < best ? key : best;
            }
            best = block_min_u64(best, S.red, phase);
            chosen[r] = (uint32_t)(best & 0xFFFFFFFFu);
        }
        // ascending index so that the 3-term sums follow the reference's order
        if (chosen[0] > chosen[1]) { const uint32_t t = chosen[0]; chosen[0] = chosen[1]; chosen[1] = t; }
        if (chosen[1] > chosen[2]) { const uint32_t t = chosen[1]; chosen[1] = chosen[2]; chosen[2] = t; }
        if (chosen[0] > chosen[1]) { const uint32_t t = chosen[0]; chosen[0] = chosen[1]; chosen[1] = t; }
        for (uint32_t i = tid; i < n; i += kFitThreads) nv.set_mask(i, (i == chosen[0] || i == chosen[1] || i == chosen[2]) ? 1 : 0);
        acc[0] = 3.f; acc[1] = acc[2] = acc[3] = 0.f;
        for (int r = 0; r < 3; ++r) {
            float x, y, z;
            nv.get(chosen[r], x, y, z);
            acc[1] += x; acc[2] += y; acc[3] += z;
        }
    }

    // ---- iterated plane fit (:185-217) -------------------------------------------------------
    float cnt = acc[0];
    float cx = acc[1] / cnt, cy = acc[2] / cnt, cz = acc[3] / cnt;  // computeCentroid
    float nx = 0.f, ny = 0.f, nz = 1.f, residual = FLT_MAX;
    int iters = 0;
    bool have_final = false;  // final plane (:220-228) already known
    float* bc = reinterpret_cast<float*>(S.misc + 8);
    for (int iter = 0; iter < fp.max_iter; ++iter) {
        if (cnt < 3.f) break;  // :196 — collapsed mask is kept (Q3)
        // computeCovariance about the centroid (point_cloud_processor.cpp:72-86)
        float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (uint32_t i = tid; i < n; i += kFitThreads) {
            if (nv.mask(i)) {
                float x, y, z;
                nv.get(i, x, y, z);
                const float dx = x - cx, dy = y - cy, dz = z - cz;
                cv[0] = fmaf(dx, dx, cv[0]); cv[1] = fmaf(dy, dx, cv[1]); cv[2] = fmaf(dy, dy, cv[2]);
                cv[3] = fmaf(dz, dx, cv[3]); cv[4] = fmaf(dz, dy, cv[4]); cv[5] = fmaf(dz, dz, cv[5]);
            }
        }
        block_sum<6>(cv, S.red, phase);
        if (tid < 32) {
            const float d = cnt - 1.f;
            const Eig3 E = eig3_sym(cv[0] / d, cv[1] / d, cv[2] / d, cv[3] / d, cv[4] / d, cv[5] / d);
            float ax = E.vec[0][0], ay = E.vec[1][0], az = E.vec[2][0];
            if (az < 0.f) { ax = -ax; ay = -ay; az = -az; }  // :93-95
            if (tid == 0) { bc[0] = ax; bc[1] = ay; bc[2] = az; }
        }
        __syncthreads();
        nx = bc[0]; ny = bc[1]; nz = bc[2];
        iters++;
        // distances, new mask, convergence, next centroid, residual of the fit just made
        float st[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // new count, sum x, y, z, changed, sum |dist| over old mask
        for (uint32_t i = tid; i < n; i += kFitThreads) {
            float x, y, z;
            nv.get(i, x, y, z);
            const float dist = plane_dist(x, y, z, cx, cy, cz, nx, ny, nz);
            const uint8_t om = nv.mask(i);
            const uint8_t nm = dist < tau;
            if (om) st[5] += dist;
            if (nm != om) { st[4] = 1.f; nv.set_mask(i, nm); }
            if (nm) { st[0] += 1.f; st[1] += x; st[2] += y; st[3] += z; }
        }
        block_sum<6>(st, S.red, phase);
        if (st[4] == 0.f) {  // :215 converged: the final fit repeats this one
            residual = st[5] / cnt;
            have_final = true;
            break;
        }
        cnt = st[0];
        if (cnt >= 3.f) { cx = st[1] / cnt; cy = st[2] / cnt; cz = st[3] / cnt; }
    }
    // ---- final fit (:220-228) ----------------------------------------------------------------
    if (!have_final) {
        if (cnt >= 3.f) {
            float cv[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (uint32_t i = tid; i < n; i += kFitThreads) {
                if (nv.mask(i)) {
                    float x, y, z;
                    nv.get(i, x, y, z);
                    const float dx = x - cx, dy = y - cy, dz = z - cz;
                    cv[0] = fmaf(dx, dx, cv[0]); cv[1] = fmaf(dy, dx, cv[1]); cv[2] = fmaf(dy, dy, cv[2]);
                    cv[3] = fmaf(dz, dx, cv[3]); cv[4] = fmaf(dz, dy, cv[4]); cv[5] = fmaf(dz, dz, cv[5]);
                }
            }
            block_sum<6>(cv, S.red, phase);
            if (tid < 32) {
                const float d = cnt - 1.f;
                const Eig3 E = eig3_sym(cv[0] / d, cv[1] / d, cv[2] / d, cv[3] / d, cv[4] / d, cv[5] / d);
                float ax = E.vec[0][0], ay = E.vec[1][0], az = E.vec[2][0];
                if (az < 0.f) { ax = -ax; ay = -ay; az = -az; }
                if (tid == 0) { bc[0] = ax; bc[1] = ay; bc[2] = az; }
            }
            __syncthreads();
            nx = bc[0]; ny = bc[1]; nz = bc[2];
            float rs[1] = {0.f};
            for (uint32_t i = tid; i < n; i += kFitThreads) {
                if (nv.mask(i)) {
                    float x, y, z;
                    nv.get(i, x, y, z);
                    rs[0] += plane_dist(x, y, z, cx, cy, cz, nx, ny, nz);
                }
            }
            block_sum<1>(rs, S.red, phase);
            residual = rs[0] / cnt;
        } else {
            cx = cy = cz = 0.f; nx = ny = 0.f; nz = 1.f; residual = FLT_MAX;  // :78-80
        }
    }
    const int n_in = (int)cnt;

    // ---- split decision (:231-235) -----------------------------------------------------------
    const float split_threshold = fp.th_dist * (1.0f + 1.5f * (float)depth);
    const uint32_t min_patch = (uint32_t)(50 + 10 * depth);
    if (!(residual > split_threshold && depth < fp.max_split_depth && n >= min_patch)) {
        // leaf: slot j labels input point sortedA[start + j].w (positional read-back, Q1)
        for (uint32_t i = tid; i < n; i += kFitThreads)
            A.labels[__float_as_uint(A.sortedA[nd.start + i].w)] = nv.mask(i);
        if (tid == 0) dbg_record(A, nd, depth, RPW_NODE_FIT, iters, n_in, -1, cx, cy, cz, nx, ny, nz, residual, 0, mean_dist);
        return;
    }

    // ---- split (:238-283) --------------------------------------------------------------------
    float sxy[2] = {0.f, 0.f};
    for (uint32_t i = tid; i < n; i += kFitThreads) {
        float x, y, z;
        nv.get(i, x, y, z);
        sxy[0] += x; sxy[1] += y;
    }
    block_sum<2>(sxy, S.red, phase);
    const float ccx = sxy[0] / (float)n, ccy = sxy[1] / (float)n;
    float var[2] = {0.f, 0.f};
    for (uint32_t i = tid; i < n; i += kFitThreads) {
        float x, y, z;
        nv.get(i, x, y, z);
        const float dx = x - ccx, dy = y - ccy;
        var[0] = fmaf(dx, dx, var[0]); var[1] = fmaf(dy, dy, var[1]);
    }
    block_sum<2>(var, S.red, phase);
    const int axis = (var[0] / (float)n > var[1] / (float)n) ? 0 : 1;  // :250
    const float median = radix_select<SMEM>(nv, n, axis, n / 2);       // upper median (Q7)

    // stable partition: thread t owns the contiguous run [t*per, (t+1)*per)
    const uint32_t per = (n + kFitThreads - 1) / kFitThreads;
    const uint32_t lo = min(n, (uint32_t)tid * per), hi = min(n, lo + per);
    uint32_t nleft = 0;
    for (uint32_t i = lo; i < hi; ++i) nleft += nv.coord(i, axis) <= median;
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
    for (int w = 0; w < kFitWarps; ++w) {
        const uint32_t v = wsum[w];
        if (w < warp) wbase += v;
        total_left += v;
    }
    uint32_t lpos = wbase + inc - nleft;  // lefts before my run
    uint32_t rpos = total_left + (lo - lpos);
    float4* dst = (((depth + 1) & 1) ? A.bufB : A.bufC) + nd.start;
    for (uint32_t i = lo; i < hi; ++i) {
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
    if (L.n < 3) {
        label_const(A, L, 0);
        if (tid == 0) dbg_record(A, L, depth + 1, RPW_NODE_SMALL, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, 0);
    }
    if (R.n < 3) {
        label_const(A, R, 0);
        if (tid == 0) dbg_record(A, R, depth + 1, RPW_NODE_SMALL, 0, 0, -1, 0, 0, 0, 0, 0, 1, FLT_MAX, 0, 0);
    }
    if (tid == 0) {
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
}

__device__ __forceinline__ void run_node(const FitArgs& A, const NodeRef nd, int depth, FitSmem S) {
    if (nd.n <= (uint32_t)A.smem_cap) process_node<true>(A, nd, depth, S);
    else process_node<false>(A, nd, depth, S);
}

__global__ void __launch_bounds__(kFitThreads, 2) rpw_fit_kernel(FitArgs A) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cg::grid_group grid = cg::this_grid();
    FitSmem S;
    S.x = reinterpret_cast<float*>(smem_raw);
    S.y = S.x + A.smem_cap;
    S.z = S.y + A.smem_cap;
    S.red = S.z + A.smem_cap;
    S.hist = reinterpret_cast<uint32_t*>(S.red + 2 * kFitWarps * kRedMax);
    S.misc = S.hist + 256;
    S.m = reinterpret_cast<uint8_t*>(S.misc + 16);
    __shared__ uint32_t s_fetch;

    // level 0: the ring/sector patches themselves
    const uint32_t n_roots = (uint32_t)A.n_roots;
    uint32_t n_done = 0;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_fetch = atomicAdd(A.fetch_ctr + 0, 1u);
        __syncthreads();
        const uint32_t id = s_fetch;
        if (id >= n_roots) break;
        // largest patches first: id walks (patch rank, scan)
        const uint32_t b = id % (uint32_t)A.n_scans, p = A.patch_order[id / (uint32_t)A.n_scans];
        const uint32_t* ps = A.patch_start + (size_t)b * (A.P + 1) + p;
        NodeRef nd;
        nd.start = ps[0]; nd.n = ps[1] - ps[0]; nd.root = b * (uint32_t)A.P + p; nd.pad = 0;
        if (nd.n == 0) continue;  // :380
        run_node(A, nd, 0, S);
        n_done++;
    }
    // deeper levels: level-synchronous, no host round trip
    int level = 0;
    for (;;) {
        grid.sync();
        const uint32_t cnt = min(__ldcg(A.q_count + level + 1), A.q_cap);
        if (cnt == 0) break;
        level++;
        const NodeRef* q = A.queue[level & 1];
        for (;;) {
            __syncthreads();
            if (threadIdx.x == 0) s_fetch = atomicAdd(A.fetch_ctr + level, 1u);
            __syncthreads();
            const uint32_t id = s_fetch;
            if (id >= cnt) break;
            NodeRef nd;
            const uint4 raw = __ldcg(reinterpret_cast<const uint4*>(q + id));
            nd.start = raw.x; nd.n = raw.y; nd.root = raw.z; nd.pad = 0;
            run_node(A, nd, level, S);
            n_done++;
        }
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
            for (int l = 0; l <= level + 1; ++l) { A.fetch_ctr[l] = 0; A.q_count[l] = 0; }
        }
    }
}

// =============================================================================================
// unit-test entry points
// =============================================================================================
__global__ void rpw_eig3_kernel(const float* __restrict__ mats, size_t count, float* __restrict__ evals, float* __restrict__ evecs) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const float* a = mats + i * 9;
    const Eig3 E = eig3_sym(a[0], a[3], a[4], a[6], a[7], a[8]);
    for (int k = 0; k < 3; ++k) evals[i * 3 + k] = E.val[k];
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) evecs[i * 9 + r * 3 + c] = E.vec[r][c];
}

__global__ void rpw_atan2_kernel(const float* __restrict__ y, const float* __restrict__ x, size_t count, float* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = atan2f_libm(y[i], x[i]);
}

// =============================================================================================
// host-side launchers
// =============================================================================================
size_t fit_smem_bytes(int smem_cap) {
    return (size_t)smem_cap * 13 + (2 * kFitWarps * kRedMax + 256 + 16) * 4 + 16;
}

cudaError_t launch_bin(cudaStream_t st, int stride_floats, const float* pts, const uint64_t* scan_off, const uint32_t* chunk_base,
                       const ZoneModel& zm, uint16_t* keys, uint8_t* labels, uint32_t* blk_hist, uint32_t* patch_total,
                       int max_chunks, int batch) {
    dim3 grid(max_chunks, batch);
    const size_t smem = (size_t)zm.num_patches * 4;
    if (stride_floats == 4) rpw_bin_kernel<4><<<grid, kBinThreads, smem, st>>>(pts, scan_off, chunk_base, zm, keys, labels, blk_hist, patch_total);
    else rpw_bin_kernel<3><<<grid, kBinThreads, smem, st>>>(pts, scan_off, chunk_base, zm, keys, labels, blk_hist, patch_total);
    return cudaGetLastError();
}

cudaError_t launch_offsets(cudaStream_t st, const uint64_t* scan_off, const uint32_t* chunk_base, uint32_t* blk_hist,
                           uint32_t* patch_start, uint32_t* patch_total, int P, int batch) {
    rpw_offsets_kernel<<<batch, 256, (size_t)(P + 1) * 4, st>>>(scan_off, chunk_base, blk_hist, patch_start, patch_total, P);
    return cudaGetLastError();
}

cudaError_t launch_scatter(cudaStream_t st, int stride_floats, const float* pts, const uint64_t* scan_off, const uint32_t* chunk_base,
                           const uint16_t* keys, const uint32_t* blk_hist, const uint32_t* patch_start, float4* sorted,
                           const uint32_t* patch_total, uint32_t* patch_order, int P, int max_chunks, int batch) {
    dim3 grid(max_chunks, batch);
    const size_t smem = (size_t)(kBinThreads / 32) * P * 4;
    if (stride_floats == 4) rpw_scatter_kernel<4><<<grid, kBinThreads, smem, st>>>(pts, scan_off, chunk_base, keys, blk_hist, patch_start, sorted, patch_total, patch_order, P);
    else rpw_scatter_kernel<3><<<grid, kBinThreads, smem, st>>>(pts, scan_off, chunk_base, keys, blk_hist, patch_start, sorted, patch_total, patch_order, P);
    return cudaGetLastError();
}

cudaError_t fit_configure(int smem_cap, int* blocks_per_sm) {
    const size_t smem = fit_smem_bytes(smem_cap);
    cudaError_t e = cudaFuncSetAttribute(rpw_fit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, rpw_fit_kernel, kFitThreads, smem);
}

cudaError_t launch_fit(cudaStream_t st, const FitArgs& args, int grid_blocks) {
    FitArgs a = args;
    void* params[] = {&a};
    return cudaLaunchCooperativeKernel((void*)rpw_fit_kernel, dim3(grid_blocks), dim3(kFitThreads), params,
                                       fit_smem_bytes(args.smem_cap), st);
}

cudaError_t launch_eig3(cudaStream_t st, const float* mats, size_t count, float* evals, float* evecs) {
    rpw_eig3_kernel<<<(unsigned)((count + 127) / 128), 128, 0, st>>>(mats, count, evals, evecs);
    return cudaGetLastError();
}

cudaError_t launch_atan2(cudaStream_t st, const float* y, const float* x, size_t count, float* out) {
    rpw_atan2_kernel<<<(unsigned)((count + 255) / 256), 256, 0, st>>>(y, x, count, out);
    return cudaGetLastError();
}

}  // namespace rpw
